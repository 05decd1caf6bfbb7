// step_kernel.cuh -- the fused per-time-step kernel.
//
// One launch = one explicit time step of CardiacModel.run's loop body
// (finitewave/core/model/cardiac_model.py:170-174):
//     u_new[n]  = sum_k u[n + off_k] * w[n, k]          (diffusion_kernel_*)
//     state, u_new[n] (+/-)= dt * I(u[n], state)        (ionic_kernel_*)
//     act_t / ECG trackers on the steps whose gate passes
// for every node the solver updates.  u is read-only, u_new and the state are
// written only by the thread that owns the node, so the kernel is race free.
//
// Thread mapping: warp = 32 consecutive flat nodes (one "chunk"), lane = node.
// A node's weights and state live at its COMPACT index (rank among updated
// nodes = position in the reference's myo_indexes), so warps read contiguous,
// fully used lines even on 30 %-fibrotic or shell-shaped tissues, and chunks
// without tissue cost one 4-byte load.  Neighbour values of u come from
// global memory through L1 (read-only path); the block's tile shape (see
// Grid) makes the neighbour lines L1 hits.
#pragma once
#include "fwb_common.cuh"
#include "models.cuh"

namespace fwb {

struct StepCommon {
    Grid g;
    const double *u;
    double *u_new;
    const double *w;      // compact SoA [K][ld]
    double *state;        // compact SoA [S][ld]
    // trackers (only read by the TRACK instantiation)
    double *act_t;
    double act_thr;
    double t;
    int do_act;
    int do_ecg;
    int n_leads;
    const double *ecg_coords;   // (n_leads, 3)
    double dr;
    double *ecg_partial;        // [n_blocks][n_leads]
};

// "no model": diffusion only (fwb_diffuse, ECG re-application)
struct NoModel {
    static constexpr int NS = 0, NP = 0;
    static constexpr uint32_t READ_MASK = 0, WRITE_MASK = 0;
    struct Consts { double dt; };
    static void derive(const double *, double dt, Consts &c) { c.dt = dt; }
    FWB_HD static void ionic(double, double &, double *, const Consts &) {}
};

template <class M> struct StepArgs {
    StepCommon k;
    typename M::Consts c;
};

// ---------------------------------------------------------------------------
// Stencil slot tables in the reference's slot order.
//   2D iso   isotropic_stencil_2d.py:126-130      2D aniso  asymmetric_stencil_2d.py:192-200
//   3D iso   isotropic_stencil_3d.py:103-109      3D aniso  asymmetric_stencil_3d.py:141-159
// entries are (d_plane, d_row, d_line); in 2D (i, j) = (row, line).
// ---------------------------------------------------------------------------
struct Off { int p, r, l; };

template <int DIM, int ST> struct Stencil;
template <> struct Stencil<2, FWB_STENCIL_ISO> {
    static constexpr int K = 5;
    FWB_HD static constexpr Off at(int k)
    {
        constexpr Off T[5] = {{0, -1, 0}, {0, 0, -1}, {0, 0, 0}, {0, 0, 1}, {0, 1, 0}};
        return T[k];
    }
};
template <> struct Stencil<2, FWB_STENCIL_ANISO> {
    static constexpr int K = 9;
    FWB_HD static constexpr Off at(int k)
    {
        constexpr Off T[9] = {{0, -1, -1}, {0, -1, 0}, {0, -1, 1}, {0, 0, -1}, {0, 0, 0},
                              {0, 0, 1},   {0, 1, -1}, {0, 1, 0},  {0, 1, 1}};
        return T[k];
    }
};
template <> struct Stencil<3, FWB_STENCIL_ISO> {
    static constexpr int K = 7;
    FWB_HD static constexpr Off at(int k)
    {
        constexpr Off T[7] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 0},
                              {0, 0, 1},  {0, 1, 0},  {1, 0, 0}};
        return T[k];
    }
};
template <> struct Stencil<3, FWB_STENCIL_ANISO> {
    static constexpr int K = 19;
    FWB_HD static constexpr Off at(int k)
    {
        constexpr Off T[19] = {
            {-1, -1, 0}, {-1, 0, 0},  {-1, 1, 0}, {0, -1, 0}, {0, 0, 0},  {0, 1, 0}, {1, -1, 0},
            {1, 0, 0},   {1, 1, 0},   {0, -1, -1}, {0, -1, 1}, {0, 0, -1}, {0, 0, 1}, {0, 1, -1},
            {0, 1, 1},   {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1}, {1, 0, 1}};
        return T[k];
    }
};

#ifdef __CUDACC__

template <class M, int DIM, int ST, bool TRACK>
__global__ void __launch_bounds__(BLOCK_THREADS)
step_kernel(const __grid_constant__ StepArgs<M> A)
{
    using S = Stencil<DIM, ST>;
    constexpr int K = S::K;
    const StepCommon &P = A.k;
    const Grid &g = P.g;
    const int lane = threadIdx.x & 31;

    const int64_t chunk = warp_chunk(g);
    uint32_t bits = 0;
    if (chunk >= 0) bits = __ldg(g.chunk_bits + chunk);
    if (!TRACK && bits == 0) return;

    const int64_t n = chunk * 32 + lane;
    const bool myo = (bits >> lane) & 1u;

    double diff = 0.0;   // u_tr - u of this node (ECG)
    if (TRACK && P.do_act && chunk >= 0 && n < g.n_nodes) {
        // ActivationTime{2,3}DTracker._track: every grid node, strict >
        const double a = P.act_t[n];
        const double uu = __ldg(P.u + n);
        if (a < 0 && uu > P.act_thr) P.act_t[n] = P.t;
    }

    if (myo) {
        const int64_t c = (int64_t)__ldg(g.chunk_base + chunk) +
                          __popc(bits & ((1u << lane) - 1u));
        const double *__restrict__ u = P.u;
        const double *__restrict__ w = P.w + c;
        const int64_t ld = g.ld;

        // issue every load of this node up front
        double un[K], wn[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const Off o = S::at(k);
            const int64_t off = (DIM == 3 ? (int64_t)o.p * g.s_plane : 0) +
                                (int64_t)o.r * g.s_row + o.l;
            un[k] = __ldg(u + n + off);
            wn[k] = ld_stream(w + (int64_t)k * ld);
        }
        double s[M::NS > 0 ? M::NS : 1];
        double *__restrict__ st = P.state + c;
#pragma unroll
        for (int q = 0; q < M::NS; ++q)
            if (M::READ_MASK & (1u << q)) s[q] = ld_stream(st + (int64_t)q * ld);

        // diffusion: left-to-right sum in slot order, no FMA contraction
        double acc = mul(un[0], wn[0]);
#pragma unroll
        for (int k = 1; k < K; ++k) acc = add(acc, mul(un[k], wn[k]));

        // centre value of u is one of the slots
        double uc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (S::at(k).p == 0 && S::at(k).r == 0 && S::at(k).l == 0) uc = un[k];

        if (TRACK) diff = acc - uc;

        M::ionic(uc, acc, s, A.c);

        P.u_new[n] = acc;
#pragma unroll
        for (int q = 0; q < M::NS; ++q)
            if (M::WRITE_MASK & (1u << q)) st_stream(st + (int64_t)q * ld, s[q]);
    }

    if (TRACK && P.do_ecg) {
        // ECG{2,3}DTracker: per lead sum over updated nodes of
        // (u_tr - u) / (d * dr), d = squared index distance, d == 0 skipped.
        // Deterministic: warp tree -> fixed-order sum over warps -> per-block
        // partial; ecg_finalize sums the partials in a fixed order.
        __shared__ double red[WARPS_PER_BLOCK][32];
        const int warp = threadIdx.x >> 5;
        double ci = 0, cj = 0, ck = 0;
        if (myo) {
            if (DIM == 3) {
                const int64_t i = n / g.s_plane, rem = n % g.s_plane;
                ci = (double)i; cj = (double)(rem / g.s_row); ck = (double)(rem % g.s_row);
            } else {
                ci = (double)(n / g.s_row); cj = (double)(n % g.s_row);
            }
        }
        for (int l0 = 0; l0 < P.n_leads; l0 += 32) {
            const int nl = min(32, P.n_leads - l0);
            for (int l = 0; l < nl; ++l) {
                const double x = __ldg(P.ecg_coords + 3 * (l0 + l));
                const double y = __ldg(P.ecg_coords + 3 * (l0 + l) + 1);
                const double z = __ldg(P.ecg_coords + 3 * (l0 + l) + 2);
                double v = 0.0;
                if (myo) {
                    const double d = (DIM == 3)
                        ? (x - ci) * (x - ci) + (y - cj) * (y - cj) + (z - ck) * (z - ck)
                        : (x - ci) * (x - ci) + (y - cj) * (y - cj) + z * z;
                    if (d > 0) v = diff / (d * P.dr);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp][l] = v;
            }
            __syncthreads();
            if (threadIdx.x < nl) {
                double sum = 0.0;
                for (int wq = 0; wq < WARPS_PER_BLOCK; ++wq) sum += red[wq][threadIdx.x];
                P.ecg_partial[(int64_t)blockIdx.x * P.n_leads + l0 + threadIdx.x] = sum;
            }
            __syncthreads();
        }
    }
}

template <class M, int DIM, int ST, bool TRACK>
static int launch_one(const StepCommon &k, const void *consts, cudaStream_t s)
{
    StepArgs<M> a;
    a.k = k;
    a.c = *reinterpret_cast<const typename M::Consts *>(consts);
    if (k.g.n_blocks <= 0) return 0;
    step_kernel<M, DIM, ST, TRACK><<<(unsigned)k.g.n_blocks, BLOCK_THREADS, 0, s>>>(a);
    FWB_KERNEL_CHECK("step_kernel");
    return 0;
}

template <class M>
static int launch_model(int dim, int stencil, bool track, const StepCommon &k,
                        const void *consts, cudaStream_t s)
{
#define FWB_CASE(D, ST)                                                         \
    if (dim == D && stencil == ST)                                              \
        return track ? launch_one<M, D, ST, true>(k, consts, s)                 \
                     : launch_one<M, D, ST, false>(k, consts, s);
    FWB_CASE(2, FWB_STENCIL_ISO)
    FWB_CASE(2, FWB_STENCIL_ANISO)
    FWB_CASE(3, FWB_STENCIL_ISO)
    FWB_CASE(3, FWB_STENCIL_ANISO)
#undef FWB_CASE
    set_error("unsupported dim/stencil %d/%d", dim, stencil);
    return FWB_E_UNSUPPORTED;
}

template <class M> static void derive_model(const double *p, double dt, void *out)
{
    M::derive(p, dt, *reinterpret_cast<typename M::Consts *>(out));
}
#endif  // __CUDACC__

// per-model entry points (one translation unit each)
typedef int (*LaunchFn)(int dim, int stencil, bool track, const StepCommon &k,
                        const void *consts, cudaStream_t s);
typedef void (*DeriveFn)(const double *p, double dt, void *consts_out);
struct ModelEntry {
    int n_state, n_params;
    uint32_t read_mask, write_mask;
    LaunchFn launch;
    DeriveFn derive;
};
constexpr int CONSTS_BYTES = 1024;

const ModelEntry *model_entry(int model);   // NULL if unknown; index FWB_N_MODELS = NoModel

#define FWB_DEFINE_MODEL_ENTRY(NAME, M)                                                   \
    static_assert(sizeof(M::Consts) <= fwb::CONSTS_BYTES, "Consts too large");            \
    extern const fwb::ModelEntry NAME = {M::NS, M::NP, M::READ_MASK, M::WRITE_MASK,       \
                                         &fwb::launch_model<M>, &fwb::derive_model<M>};

}  // namespace fwb

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
// Thread mapping: warp = 32 consecutive flat nodes (one "chunk", taken from the
// work list), lane = node.  A node's weights and state live at its COMPACT
// index (rank among updated nodes = position in the reference's myo_indexes),
// so warps read contiguous, fully used lines even on 30 %-fibrotic or
// shell-shaped tissues.  Neighbour values of u come from global memory through
// L1 (read-only path); the work list's tile order makes the neighbour lines L1
// hits.  State values are loaded where the model needs them and stored as soon
// as they are final, which keeps the 19-state TP06 kernel at 2 blocks per SM.
//
// Slab runs (one process per GPU, slabs along the slowest axis): the blocks that
// own the two slab-boundary slices come first in the work list.  They wait for
// the neighbour's "previous step done" flag, compute, store u_new both locally
// and straight into the neighbour's ghost slice through a peer-mapped pointer
// (NVLink), and the last of them to finish raises the neighbour's flag for the
// next step.  Interior blocks never wait: the exchange overlaps the interior.
#pragma once
#include "fwb_common.cuh"
#include "models.cuh"

namespace fwb {

// one side (lo = towards slice 0, hi = towards the last slice) of a slab halo
struct HaloSide {
    int on;
    int64_t first;            // first flat node of the owned boundary slice
    double *peer_dst;         // neighbour's ghost slice inside ITS u_new buffer (peer pointer)
    unsigned *peer_flag;      // neighbour's flag that this side raises (peer pointer)
    const unsigned *flag;     // local flag the neighbour raises
    unsigned *counter;        // local count of finished boundary blocks
    unsigned n_blocks;        // boundary blocks of this side
    unsigned first_block;     // first block index of this side in the work list
};

struct Halo {
    int on;
    int64_t slice;            // nodes per slice
    unsigned epoch;           // this step's number; flags must be >= epoch to start
    HaloSide lo, hi;
};

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
    Halo halo;
};

// "no model": diffusion only (fwb_diffuse, ECG re-application)
struct NoModel {
    static constexpr int NS = 0, NP = 0, MIN_BLOCKS = 4;
    static constexpr uint32_t READ_MASK = 0, WRITE_MASK = 0;
    struct Consts { double dt; };
    static bool derive(const double *, double dt, Consts &c) { c.dt = dt; return true; }
    template <class IO> FWB_HD static void ionic(double, double &, IO &, const Consts &) {}
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

// state accessor of one node: slot q lives at state[q * ld + c]
struct StateIO {
    double *base;
    int64_t stride;
    __device__ __forceinline__ double ld(int q) const { return ld_stream(base + (int64_t)q * stride); }
    __device__ __forceinline__ void st(int q, double v) const { st_stream(base + (int64_t)q * stride, v); }
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class M, int DIM, int ST, bool TRACK, bool HALO>
__global__ void __launch_bounds__(BLOCK_THREADS, M::MIN_BLOCKS)
step_kernel(const __grid_constant__ StepArgs<M> A)
{
    using S = Stencil<DIM, ST>;
    constexpr int K = S::K;
    const StepCommon &P = A.k;
    const Grid &g = P.g;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // which slab boundary (if any) this block belongs to
    const HaloSide *side = nullptr;
    if (HALO) {
        const unsigned b = blockIdx.x;
        if (P.halo.lo.on && b - P.halo.lo.first_block < P.halo.lo.n_blocks) side = &P.halo.lo;
        else if (P.halo.hi.on && b - P.halo.hi.first_block < P.halo.hi.n_blocks) side = &P.halo.hi;
        if (side) {
            // the neighbour has finished the previous step on this interface: its
            // stores into our ghost slice have landed and it no longer reads the
            // ghost slice (in its other buffer) that we are about to overwrite
            if (threadIdx.x == 0)
                while (ld_acquire_sys(side->flag) < P.halo.epoch) __nanosleep(64);
            __syncthreads();
        }
    }

    const int64_t slot = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    const int64_t chunk = slot < g.n_work ? (int64_t)__ldg(g.worklist + slot) : -1;
    uint32_t bits = 0;
    if (chunk >= 0) bits = __ldg(g.chunk_bits + chunk);

    const int64_t n = chunk * 32 + lane;
    const bool myo = (bits >> lane) & 1u;

    double diff = 0.0;   // u_tr - u of this node (ECG)
    if (TRACK && P.do_act && chunk >= 0 && n < g.n_nodes) {
        // ActivationTime{2,3}DTracker._track: every grid node, strict >
        const double a = P.act_t[n];
        const double uu = P.u[n];
        if (a < 0 && uu > P.act_thr) P.act_t[n] = P.t;
    }

    if (myo) {
        const int64_t c = (int64_t)__ldg(g.chunk_base + chunk) +
                          __popc(bits & ((1u << lane) - 1u));
        const double *__restrict__ u = P.u + n;
        const double *__restrict__ w = P.w + c;
        const int64_t ld = g.ld;

        // diffusion: issue the 2K loads, then the left-to-right sum in slot order
        // (no FMA contraction: -fmad=false)
        double un[K], wn[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const Off o = S::at(k);
            const int64_t off = (DIM == 3 ? (int64_t)o.p * g.s_plane : 0) +
                                (int64_t)o.r * g.s_row + o.l;
            // boundary blocks of a slab read ghost values a peer GPU wrote during the
            // previous step: they must not come from the non-coherent path
            un[k] = (HALO && side) ? __ldcg(u + off) : __ldg(u + off);
            wn[k] = ld_stream(w);
            w += ld;
        }
        double acc = mul(un[0], wn[0]);
#pragma unroll
        for (int k = 1; k < K; ++k) acc = add(acc, mul(un[k], wn[k]));

        // centre value of u is one of the slots
        double uc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (S::at(k).p == 0 && S::at(k).r == 0 && S::at(k).l == 0) uc = un[k];

        if (TRACK) diff = acc - uc;

        StateIO io{P.state + c, ld};
        M::ionic(uc, acc, io, A.c);

        P.u_new[n] = acc;
        if (HALO && side) side->peer_dst[n - side->first] = acc;
    }

    if (HALO && side) {
        // publish: all of this block's peer stores are visible system-wide before the
        // block is counted; the last block of the side raises the neighbour's flag
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(side->counter, 1u) + 1u;
            if (done == side->n_blocks) {
                *side->counter = 0;
                __threadfence_system();
                st_release_sys(side->peer_flag, P.halo.epoch + 1u);
            }
        }
    }

    if (TRACK && P.do_ecg) {
        // ECG{2,3}DTracker: per lead sum over updated nodes of
        // (u_tr - u) / (d * dr), d = squared index distance, d == 0 skipped.
        // Deterministic: warp tree -> fixed-order sum over warps -> per-block
        // partial; ecg_finalize sums the partials in a fixed order.
        __shared__ double red[WARPS_PER_BLOCK][32];
        double ci = 0, cj = 0, ck = 0;
        if (myo) {
            if (DIM == 3) {
                const int64_t i = n / g.s_plane, rem = n % g.s_plane;
                ci = (double)(i + g.slow_offset);
                cj = (double)(rem / g.s_row); ck = (double)(rem % g.s_row);
            } else {
                ci = (double)(n / g.s_row + g.slow_offset); cj = (double)(n % g.s_row);
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

inline int64_t step_blocks(const Grid &g) { return (g.n_work + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; }

template <class M, int DIM, int ST, bool TRACK, bool HALO>
static int launch_one(const StepCommon &k, const void *consts, cudaStream_t s)
{
    StepArgs<M> a;
    a.k = k;
    a.c = *reinterpret_cast<const typename M::Consts *>(consts);
    const int64_t blocks = step_blocks(k.g);
    if (blocks <= 0) return 0;
    step_kernel<M, DIM, ST, TRACK, HALO><<<(unsigned)blocks, BLOCK_THREADS, 0, s>>>(a);
    FWB_KERNEL_CHECK("step_kernel");
    return 0;
}

template <class M>
static int launch_model(int dim, int stencil, bool track, const StepCommon &k,
                        const void *consts, cudaStream_t s)
{
    const bool halo = k.halo.on != 0;
#define FWB_CASE(D, ST)                                                              \
    if (dim == D && stencil == ST) {                                                 \
        if (halo)                                                                    \
            return track ? launch_one<M, D, ST, true, true>(k, consts, s)            \
                         : launch_one<M, D, ST, false, true>(k, consts, s);          \
        return track ? launch_one<M, D, ST, true, false>(k, consts, s)               \
                     : launch_one<M, D, ST, false, false>(k, consts, s);             \
    }
    FWB_CASE(2, FWB_STENCIL_ISO)
    FWB_CASE(2, FWB_STENCIL_ANISO)
    FWB_CASE(3, FWB_STENCIL_ISO)
    FWB_CASE(3, FWB_STENCIL_ANISO)
#undef FWB_CASE
    set_error("unsupported dim/stencil %d/%d", dim, stencil);
    return FWB_E_UNSUPPORTED;
}

template <class M> static bool derive_model(const double *p, double dt, void *out)
{
    return M::derive(p, dt, *reinterpret_cast<typename M::Consts *>(out));
}
#endif  // __CUDACC__

// per-model entry points (one translation unit each)
typedef int (*LaunchFn)(int dim, int stencil, bool track, const StepCommon &k,
                        const void *consts, cudaStream_t s);
typedef bool (*DeriveFn)(const double *p, double dt, void *consts_out);
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

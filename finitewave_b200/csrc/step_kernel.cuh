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
// hits.  State values are read where the model needs them and stored as soon
// as they are final, which keeps the 19-state TP06 kernel at 64-80 registers.
//
// Two kernels implement this step:
//   step_kernel      one block per tile (all models; the FP64-bound LR91 / TP06 /
//                    Courtemanche always use it).  STAGED = 0: every operand through LDG;
//                    STAGED = 1 / 2: the tile's state rows (and weight rows) arrive by TMA
//                    bulk copies in shared memory, so the FP64 work never waits on HBM
//   step_kernel_tma  persistent blocks (grid = resident blocks of the GPU) that walk
//                    the tile list; the tile's weight rows and state rows -- the
//                    HBM streams, contiguous in the tile-ordered compact layout --
//                    are fetched by TMA bulk copies (cp.async.bulk -> UBLKCP) into a
//                    2-4 stage shared-memory ring guarded by mbarriers, so tens of KB
//                    per SM are in flight independent of register occupancy while
//                    the warps compute the previous tile (light, HBM-bound models).
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
    const uint32_t *tile_base;  // [n_work/8 + 1] compact index of each tile's first node, or NULL
    const uint4 *records;       // [n_work] {chunk, bits, chunk base, tile base} (TMA kernel)
};

// "no model": diffusion only (fwb_diffuse, ECG re-application)
struct NoModel {
    static constexpr int NS = 0, NP = 0, MIN_BLOCKS = 4;
    static constexpr bool USE_TMA = true;
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

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
constexpr int tma_popc(uint32_t m) { return m ? (int)(m & 1u) + tma_popc(m >> 1) : 0; }
// position of state slot q among the staged (read) slots
constexpr int tma_slot(uint32_t mask, int q) { return tma_popc(mask & ((1u << q) - 1u)); }
// q-th staged slot -> state slot
__host__ __device__ constexpr int tma_nth(uint32_t mask, int n)
{
    int q = 0;
    for (; q < 32; ++q)
        if ((mask >> q) & 1u) { if (n == 0) break; --n; }
    return q;
}

// Models with many state arrays (LR91, TP06, Courtemanche) cannot hold their state in
// registers and load it where it is used -- each such load would pay the full HBM latency
// in the middle of the FP64 work.  With the tile-ordered compact layout a tile's state rows
// are contiguous, so one thread of the block fetches them with TMA bulk copies into shared
// memory while all threads wait for the stencil operands anyway; the model then reads its
// state at shared-memory latency (step_kernel<..., STAGED>).  The two switches exist for A/B
// measurements (scripts/gpu_ab.sh).
#ifdef FWB_NO_STAGE
template <class M> constexpr bool stage_state() { return false; }
#else
template <class M> constexpr bool stage_state() { return M::NS > 4; }
#endif
#ifdef FWB_NO_STAGE_W
template <class M> constexpr bool stage_weights() { return false; }
#else
template <class M> constexpr bool stage_weights() { return true; }
#endif

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase)
{
    const uint32_t a = smem_u32(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(phase)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy (TMA, 1-D); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

constexpr int TMA_REC_BYTES = WARPS_PER_BLOCK * 16;   // the tile's 8 work records
constexpr int TMA_SEG = 258;   // doubles per staged row: 256 nodes + 16-B alignment slack

// state accessor of the TMA kernel: reads from the staged rows, writes to global
template <class M> struct StateIOTma {
    const double *sm;     // this node's column in the staged state rows
    double *gp;           // this node's column in the global compact state
    int64_t stride;
    // (__popc of a constant folds; the constexpr recursion of tma_slot() does not when q
    // only becomes a constant after inlining, and costs a 150-instruction loop per load)
    __device__ __forceinline__ double ld(int q) const
    {
        return sm[__popc(M::READ_MASK & ((1u << q) - 1u)) * TMA_SEG];
    }
    __device__ __forceinline__ void st(int q, double v) const { st_stream(gp + (int64_t)q * stride, v); }
};

// state accessor of one node: slot q lives at state[q * ld + c]
struct StateIO {
    double *base;
    int64_t stride;
    __device__ __forceinline__ double ld(int q) const { return ld_stream(base + (int64_t)q * stride); }
    __device__ __forceinline__ void st(int q, double v) const { st_stream(base + (int64_t)q * stride, v); }
    // pull every row this node will read towards the SM without holding registers: the
    // models with many state arrays load them where they are used (register budget), and
    // would otherwise pay the full HBM latency at each of those points
    template <uint32_t MASK> __device__ __forceinline__ void prefetch() const
    {
#pragma unroll
        for (int q = 0; q < 32; ++q)
            if ((MASK >> q) & 1u) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (int64_t)q * stride));
            }
    }
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

// STAGED: 0 = every operand through LDG; 1 = the tile's state rows by TMA into shared memory;
// 2 = state rows AND weight rows (3 blocks per SM: 76 KB each).  1 and 2 need the tile-ordered
// compact layout (P.tile_base / P.records): block b owns tile b = compact range
// [tile_base[b], tile_base[b + 1]).
template <class M, int K, int STAGED> struct StageCfg {
    static constexpr int NSR = tma_popc(M::READ_MASK);
    static constexpr int ROWS = STAGED == 2 ? K + NSR : (STAGED == 1 ? NSR : 0);
    static constexpr size_t SMEM = (size_t)ROWS * TMA_SEG * sizeof(double);
    // blocks per SM the launch bounds ask for: the model's target, capped by what the staged
    // rows leave of the 228 KB of shared memory per SM (1 KB per block is reserved)
    static constexpr int FIT = SMEM ? (int)((228 * 1024) / (SMEM + 1024 + 128)) : 32;
    static constexpr int BLOCKS = M::MIN_BLOCKS < FIT ? M::MIN_BLOCKS : (FIT < 1 ? 1 : FIT);
};

template <class M, int DIM, int ST, bool TRACK, bool HALO, int STAGED = 0>
__global__ void __launch_bounds__(BLOCK_THREADS, (StageCfg<M, Stencil<DIM, ST>::K, STAGED>::BLOCKS))
step_kernel(const __grid_constant__ StepArgs<M> A)
{
    using S = Stencil<DIM, ST>;
    constexpr int K = S::K;
    const StepCommon &P = A.k;
    const Grid &g = P.g;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int NSR = tma_popc(M::READ_MASK);
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    double *const staged_w = reinterpret_cast<double *>(dyn_smem);              // [K][TMA_SEG] (mode 2)
    double *const staged = staged_w + (STAGED == 2 ? K * TMA_SEG : 0);          // [NSR][TMA_SEG]
    __shared__ uint64_t staged_full[2];                                         // state, weights
#ifdef FWB_EXP_SMEM
    exp_table_to_smem();
    if (!STAGED) __syncthreads();          // STAGED: the barrier after mbar_init below
#endif
    if (STAGED) {
        if (threadIdx.x == 0) {
            mbar_init(&staged_full[0], 1);
            mbar_init(&staged_full[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t c0 = __ldg(P.tile_base + blockIdx.x), c1 = __ldg(P.tile_base + blockIdx.x + 1);
            const uint32_t c0a = c0 & ~1u;
            const unsigned bytes = c1 > c0 ? (((c1 - c0a) + 1u) & ~1u) * 8u : 0u;
            if (STAGED == 2) {
                mbar_arrive_expect_tx(&staged_full[1], bytes * K);
                if (bytes) {
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        tma_load_1d(staged_w + k * TMA_SEG, P.w + (int64_t)k * g.ld + c0a, bytes,
                                    &staged_full[1]);
                }
            }
            mbar_arrive_expect_tx(&staged_full[0], bytes * NSR);
            if (bytes) {
#pragma unroll
                for (int q = 0; q < NSR; ++q)
                    tma_load_1d(staged + q * TMA_SEG,
                                P.state + (int64_t)tma_nth(M::READ_MASK, q) * g.ld + c0a, bytes,
                                &staged_full[0]);
            }
        }
    }

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

    // the warp's work record {chunk, bits, compact base, .}: one 16-B load when the engine
    // built the records (tile-ordered layout), else three dependent loads
    const int64_t slot = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    int64_t chunk = -1;
    uint32_t bits = 0, cbase = 0, tbase = 0;
    if (slot < g.n_work) {
        if (STAGED || P.records) {
            const uint4 rec = __ldg(P.records + slot);
            chunk = (int64_t)(int32_t)rec.x;
            if (chunk >= 0) { bits = rec.y; cbase = rec.z; tbase = rec.w; }
        } else {
            chunk = (int64_t)__ldg(g.worklist + slot);
            if (chunk >= 0) { bits = __ldg(g.chunk_bits + chunk); cbase = __ldg(g.chunk_base + chunk); }
        }
    }

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
        const int64_t c = (int64_t)cbase + __popc(bits & ((1u << lane) - 1u));
        const double *__restrict__ u = P.u + n;
        const double *__restrict__ w = P.w + c;
        const int64_t ld = g.ld;
        if constexpr (STAGED == 0) {
#ifndef FWB_NO_PREFETCH
            if (M::NS > 4) StateIO{P.state + c, ld}.template prefetch<M::READ_MASK>();
#endif
        }

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
            if (STAGED != 2) wn[k] = ld_stream(w);
            w += ld;
        }
        if (STAGED == 2) {
            mbar_wait(&staged_full[1], 0);
            const double *sw = staged_w + (c - (int64_t)(tbase & ~1u));
#pragma unroll
            for (int k = 0; k < K; ++k) wn[k] = sw[k * TMA_SEG];
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

        if constexpr (STAGED != 0) {
            mbar_wait(&staged_full[0], 0);
            StateIOTma<M> io{staged + (c - (int64_t)(tbase & ~1u)), P.state + c, ld};
            M::ionic(uc, acc, io, A.c);
        } else {
            StateIO io{P.state + c, ld};
            M::ionic(uc, acc, io, A.c);
        }

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


// ---------------------------------------------------------------------------
// step_kernel_tma: persistent, TMA-fed variant (see the header comment)
// ---------------------------------------------------------------------------

template <class M, int K> struct TmaCfg {
    static constexpr int NSR = tma_popc(M::READ_MASK);
    static constexpr int NARR = K + NSR;
    static constexpr int NSTAGE = 2;
    // blocks per SM the launch bounds ask for: as many as the ring allows (227 KB / SM)
    static constexpr size_t STAGE = (size_t)NARR * TMA_SEG * sizeof(double) + TMA_REC_BYTES;
    static constexpr size_t SMEM = NSTAGE * STAGE + 64;
    static constexpr int BLOCKS = SMEM * 4 <= 224 * 1024 ? 4 : (SMEM * 3 <= 224 * 1024 ? 3 : 2);
};

template <class M, int DIM, int ST, bool TRACK, bool HALO>
__global__ void __launch_bounds__(BLOCK_THREADS, (TmaCfg<M, Stencil<DIM, ST>::K>::BLOCKS))
step_kernel_tma(const __grid_constant__ StepArgs<M> A)
{
    using S = Stencil<DIM, ST>;
    constexpr int K = S::K;
    using C = TmaCfg<M, K>;
    constexpr int NARR = C::NARR, NSTAGE = C::NSTAGE;
    constexpr int STAGE_D = (int)(C::STAGE / sizeof(double));   // doubles per ring stage
    const StepCommon &P = A.k;
    const Grid &g = P.g;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);     // [NSTAGE]{[NARR][SEG] rows, 8 records}
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)NSTAGE * STAGE_D);

    const int64_t n_tiles = g.n_work / WARPS_PER_BLOCK;
    exp_table_to_smem();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer (warp 0): start the bulk copies of tile t (compact range [c0, c1)) into
    // ring stage s
    auto issue = [&](int64_t t, int s, uint32_t c0, uint32_t c1) {
        const uint32_t c0a = c0 & ~1u;
        const unsigned bytes = c1 > c0 ? (((c1 - c0a) + 1u) & ~1u) * 8u : 0u;
        if (lane == 0) mbar_arrive_expect_tx(full + s, bytes * NARR + TMA_REC_BYTES);
        __syncwarp();
        double *stage = ring + (size_t)s * STAGE_D;
        if (lane == 31)
            tma_load_1d(stage + NARR * TMA_SEG, P.records + t * WARPS_PER_BLOCK, TMA_REC_BYTES,
                        full + s);
        if (bytes) {
            for (int a = lane; a < NARR; a += 32) {
                const double *src = a < K
                    ? P.w + (int64_t)a * g.ld + c0a
                    : P.state + (int64_t)tma_nth(M::READ_MASK, a - K) * g.ld + c0a;
                tma_load_1d(stage + a * TMA_SEG, src, bytes, full + s);
            }
        }
    };
    // compact range of the tile that will be issued next (loaded one iteration early so
    // that the producer warp never waits on it)
    uint32_t nx0 = 0, nx1 = 0;
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE - 1; ++s) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (t < n_tiles) issue(t, s, __ldg(P.tile_base + t), __ldg(P.tile_base + t + 1));
        }
        const int64_t t = (int64_t)blockIdx.x + (int64_t)(NSTAGE - 1) * gridDim.x;
        if (t < n_tiles) { nx0 = __ldg(P.tile_base + t); nx1 = __ldg(P.tile_base + t + 1); }
    }

    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % NSTAGE;
        const unsigned phase = (unsigned)(it / NSTAGE) & 1u;
        if (warp == 0) {
            // the stage consumed in the previous iteration is free (barrier at loop end)
            const int64_t t2 = t + (int64_t)(NSTAGE - 1) * gridDim.x;
            if (t2 < n_tiles) issue(t2, (it + NSTAGE - 1) % NSTAGE, nx0, nx1);
            const int64_t t3 = t2 + gridDim.x;
            if (t3 < n_tiles) { nx0 = __ldg(P.tile_base + t3); nx1 = __ldg(P.tile_base + t3 + 1); }
        }

        const HaloSide *side = nullptr;
        if (HALO) {
            const unsigned b = (unsigned)t;
            if (P.halo.lo.on && b - P.halo.lo.first_block < P.halo.lo.n_blocks) side = &P.halo.lo;
            else if (P.halo.hi.on && b - P.halo.hi.first_block < P.halo.hi.n_blocks) side = &P.halo.hi;
            if (side) {
                if (threadIdx.x == 0)
                    while (ld_acquire_sys(side->flag) < P.halo.epoch) __nanosleep(64);
                __syncthreads();
            }
        }

        // everything this tile needs from HBM (work records, weight rows, state rows) was
        // requested one iteration ago; normally the wait returns at once
        mbar_wait(full + s, phase);
        const double *stage = ring + (size_t)s * STAGE_D;
        const uint4 rec = reinterpret_cast<const uint4 *>(stage + NARR * TMA_SEG)[warp];
        const int64_t chunk = (int64_t)(int32_t)rec.x;
        const uint32_t bits = chunk >= 0 ? rec.y : 0u;
        const int64_t n = chunk * 32 + lane;
        const bool myo = (bits >> lane) & 1u;

        if (TRACK && P.do_act && chunk >= 0 && n < g.n_nodes) {
            const double a = P.act_t[n];
            const double uu = P.u[n];
            if (a < 0 && uu > P.act_thr) P.act_t[n] = P.t;
        }

        double un[K];
        if (myo) {
            const double *__restrict__ u = P.u + n;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const Off o = S::at(k);
                const int64_t off = (DIM == 3 ? (int64_t)o.p * g.s_plane : 0) +
                                    (int64_t)o.r * g.s_row + o.l;
                un[k] = (HALO && side) ? __ldcg(u + off) : __ldg(u + off);
            }
        }

        if (myo) {
            const uint32_t c = rec.z + __popc(bits & ((1u << lane) - 1u));
            const double *row = stage + (c - (rec.w & ~1u));

            double acc = mul(un[0], row[0]);
#pragma unroll
            for (int k = 1; k < K; ++k) acc = add(acc, mul(un[k], row[k * TMA_SEG]));
            double uc = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (S::at(k).p == 0 && S::at(k).r == 0 && S::at(k).l == 0) uc = un[k];

            StateIOTma<M> io{row + K * TMA_SEG, P.state + c, g.ld};
            M::ionic(uc, acc, io, A.c);

            P.u_new[n] = acc;
            if (HALO && side) side->peer_dst[n - side->first] = acc;
        }

        if (HALO && side) {
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
        __syncthreads();   // every warp is done with stage s before it is refilled
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
    if constexpr (stage_state<M>()) {
        if (k.tile_base && k.records) {
            // everything by TMA when three such blocks fit an SM and no ECG sample is due
            // (its reduction buffer would cost the third block); else the state rows only
            constexpr int K = Stencil<DIM, ST>::K;
            constexpr bool FULL = !TRACK && stage_weights<M>() && StageCfg<M, K, 2>::FIT >= 3;
            if constexpr (FULL) {
                auto kern = step_kernel<M, DIM, ST, TRACK, HALO, 2>;
                static bool attr = false;
                if (!attr) {
                    FWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)StageCfg<M, K, 2>::SMEM));
                    attr = true;
                }
                kern<<<(unsigned)blocks, BLOCK_THREADS, StageCfg<M, K, 2>::SMEM, s>>>(a);
                note_step_variant(2);
            } else {
                step_kernel<M, DIM, ST, TRACK, HALO, 1>
                    <<<(unsigned)blocks, BLOCK_THREADS, StageCfg<M, K, 1>::SMEM, s>>>(a);
                note_step_variant(1);
            }
            FWB_KERNEL_CHECK("step_kernel (staged)");
            return 0;
        }
    }
    step_kernel<M, DIM, ST, TRACK, HALO><<<(unsigned)blocks, BLOCK_THREADS, 0, s>>>(a);
    note_step_variant(0);
    FWB_KERNEL_CHECK("step_kernel");
    return 0;
}

template <class M, int DIM, int ST, bool TRACK, bool HALO>
static int launch_tma(const StepCommon &k, const void *consts, cudaStream_t s)
{
    using C = TmaCfg<M, Stencil<DIM, ST>::K>;
    auto kern = step_kernel_tma<M, DIM, ST, TRACK, HALO>;
    static int resident = 0;      // blocks of this instantiation the device holds at once
    if (resident == 0) {
        FWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C::SMEM));
        int dev = 0, sms = 0, per_sm = 0;
        FWB_CUDA(cudaGetDevice(&dev));
        FWB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK_THREADS,
                                                               C::SMEM));
        if (per_sm < 1) { set_error("step_kernel_tma does not fit on an SM"); return FWB_E_UNSUPPORTED; }
        resident = sms * per_sm;
    }
    StepArgs<M> a;
    a.k = k;
    a.c = *reinterpret_cast<const typename M::Consts *>(consts);
    const int64_t tiles = step_blocks(k.g);
    if (tiles <= 0) return 0;
    int64_t blocks = tiles < resident ? tiles : resident;
    if (HALO) {
        // boundary tiles wait on a neighbour GPU: they must all be resident in the first round
        const int64_t nb = (int64_t)k.halo.lo.n_blocks + k.halo.hi.n_blocks;
        if (nb > blocks) { set_error("slab boundary does not fit the resident grid"); return FWB_E_UNSUPPORTED; }
    }
    kern<<<(unsigned)blocks, BLOCK_THREADS, C::SMEM, s>>>(a);
    note_step_variant(3);
    FWB_KERNEL_CHECK("step_kernel_tma");
    return 0;
}

template <class M, int DIM, int ST, bool TRACK, bool HALO>
static int launch_pick(const StepCommon &k, const void *consts, cudaStream_t s)
{

    if constexpr (M::USE_TMA) {
        if (k.tile_base && !k.do_ecg) return launch_tma<M, DIM, ST, TRACK, HALO>(k, consts, s);
    }
    return launch_one<M, DIM, ST, TRACK, HALO>(k, consts, s);
}

template <class M>
static int launch_model(int dim, int stencil, bool track, const StepCommon &k,
                        const void *consts, cudaStream_t s)
{
    const bool halo = k.halo.on != 0;
#define FWB_CASE(D, ST)                                                              \
    if (dim == D && stencil == ST) {                                                 \
        if (halo)                                                                    \
            return track ? launch_pick<M, D, ST, true, true>(k, consts, s)           \
                         : launch_pick<M, D, ST, false, true>(k, consts, s);         \
        return track ? launch_pick<M, D, ST, true, false>(k, consts, s)              \
                     : launch_pick<M, D, ST, false, false>(k, consts, s);            \
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

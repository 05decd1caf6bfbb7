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
#include <map>
#include <mutex>

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
    // compact-lane tile kernel (step_kernel_tile)
    const uint4 *tile_rec;      // [n_work/8] {compact base, count, chunk id of slot 0, boundary}
    const uint8_t *pos_of;      // [ld] (slot << 5) | lane of each compact node inside its tile
    uint32_t *defer_list;       // [n_work/8] tiles handed to the reference-statement kernel
    uint32_t *defer_ctr;        // [0] entries in defer_list, [1] finished blocks of that kernel
    const int32_t *node_of;     // PACKED mode (sparse tissue): flat node of each compact index;
    int64_t n_packed;           // block t owns compact nodes [256 t, 256 t + 256) of n_packed
    int copy_idle;              // ring kernel: lanes of listed chunks that own no updated node
                                // store u -> u_new (the value that is there already), so that
                                // every 32-byte sector of u_new is written whole
    int brick;                  // the launch carries a tensor map of u (interior tiles use it)
    const void *tmap_host;      // HOST pointer to that CUtensorMap (read by the launcher only)
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
}  // namespace fwb
#include <cuda.h>      // CUtensorMap (driver API type only; the encoder is fetched at run time)
namespace fwb {

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
    bool on = true;       // false: a shadow lane of a partially filled warp (computes, never stores)
    // (__popc of a constant folds; the constexpr recursion of tma_slot() does not when q
    // only becomes a constant after inlining, and costs a 150-instruction loop per load)
    __device__ __forceinline__ double ld(int q) const
    {
        return sm[__popc(M::READ_MASK & ((1u << q) - 1u)) * TMA_SEG];
    }
    __device__ __forceinline__ void st(int q, double v) const
    {
        if (on) st_stream(gp + (int64_t)q * stride, v);
    }
};

// state accessor of one node: slot q lives at state[q * ld + c]
struct StateIO {
    double *base;
    int64_t stride;
    bool on = true;
    __device__ __forceinline__ double ld(int q) const { return ld_stream(base + (int64_t)q * stride); }
    __device__ __forceinline__ void st(int q, double v) const
    {
        if (on) st_stream(base + (int64_t)q * stride, v);
    }
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
// (LR91 / TP06 / Courtemanche normally run step_kernel_tile; this kernel is their fallback for
// callers without the tile layout and carries both of their paths: 2 blocks per SM, no spills)
__global__ void __launch_bounds__(BLOCK_THREADS, (stage_state<M>() ? 2 : StageCfg<M, Stencil<DIM, ST>::K, STAGED>::BLOCKS))
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
// step_kernel_tile: LR91 / TP06 / Courtemanche -- the FP64-heavy models
//
// One block per SPATIAL tile of the tile-granular work list (3D: 2 planes x 4 rows x 32 nodes,
// 2D: 8 rows x 32 nodes; slab-boundary tiles 1 x 8 x 32 / 1 x 256).
//  * compact lanes: thread r of the block owns the r-th updated node of the tile (`pos_of`
//    gives its slot and lane), so a warp is 32 myocytes whatever the tissue looks like --
//    a ventricle wall or 30 % fibrosis no longer leaves lanes idle while the warp issues the
//    whole FP64 instruction stream of the model
//  * the tile's u neighbourhood arrives as ONE brick ((2+2) x (4+2) x (32+4) doubles in 3D,
//    (8+2) x (32+4) in 2D) by a tensor-TMA copy (cp.async.bulk.tensor -> UTMALDG) into shared
//    memory; the 5 ... 19 stencil operands are shared-memory reads at compile-time offsets.
//    Slab-boundary tiles (which read ghost values a peer GPU stored) and grids whose line
//    length is not a multiple of 32 take plain loads instead
//  * the tile's state rows (contiguous in the tile-ordered compact layout) arrive by 1-D TMA
//    bulk copies on the same mbarrier; the weight rows are coalesced streaming loads straight
//    into registers (they are consumed once, right away)
//  * the hot kernel carries the model's rearranged fast path ONLY.  A tile with a node that
//    is not eligible for it (|u| >= 300 mV, a non-positive or non-finite concentration, NaN)
//    is appended to `defer_list` and left untouched; step_kernel_tile<..., SLOW = true>, launched
//    right after, recomputes those tiles from scratch with the reference statement.  The
//    cold code no longer sets the hot kernel's register count (TP06: 64 registers, 4 blocks
//    per SM, against 80 / 3 with both paths in one kernel)
// ---------------------------------------------------------------------------
template <int DIM> struct Brick;
// (the box of a tensor-TMA copy must START on a 16-byte boundary of the line -- an odd fp64
// column is an "illegal instruction" -- so the brick starts two columns left of the tile and
// is 36 columns wide: X0 = 2)
template <> struct Brick<3> {
    static constexpr int P = 4, R = 6, L = 36, X0 = 2, N = P * R * L;
    // brick index of the node at tile slot q, lane ln (interior tile: 2 planes x 4 rows)
    __device__ static int centre(int q, int ln) { return (((q >> 2) + 1) * R + (q & 3) + 1) * L + ln + X0; }
    __host__ __device__ static constexpr int off(const Off o) { return (o.p * R + o.r) * L + o.l; }
};
template <> struct Brick<2> {
    static constexpr int P = 1, R = 10, L = 36, X0 = 2, N = R * L;
    __device__ static int centre(int q, int ln) { return (q + 1) * L + ln + X0; }
    __host__ __device__ static constexpr int off(const Off o) { return o.r * L + o.l; }
};
constexpr int BRICK_BYTES_MAX = ((Brick<3>::N * 8 + 127) / 128) * 128;

template <class M, bool BRICK> struct TileCfg {
    static constexpr int NSR = tma_popc(M::READ_MASK);
    static constexpr size_t STATE = (size_t)NSR * TMA_SEG * sizeof(double);
    static constexpr size_t SMEM = STATE + (BRICK ? BRICK_BYTES_MAX : 0);
};

// flat node of tile slot q, lane ln (mirrors worklist_mark_kernel)
template <int DIM>
__device__ __forceinline__ int64_t tile_node(const Grid &g, bool tiled, bool bnd, int64_t n0, int q, int ln)
{
    if (!tiled) return n0 + q * 32 + ln;
    if (DIM == 3) return bnd ? n0 + (int64_t)q * g.s_row + ln
                             : n0 + (int64_t)(q >> 2) * g.s_plane + (int64_t)(q & 3) * g.s_row + ln;
    return bnd ? n0 + q * 32 + ln : n0 + (int64_t)q * g.s_row + ln;
}

__device__ __forceinline__ void tma_load_brick3(void *dst, const CUtensorMap *map, int c0, int c1,
                                                int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_brick2(void *dst, const CUtensorMap *map, int c0, int c1,
                                                uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// one tile: everything between fetching the tile record and the tracker reductions
template <class M, int DIM, int ST, bool TRACK, bool HALO, bool BRICK, bool SLOW>
__device__ __forceinline__ void tile_body(const StepArgs<M> &A, const CUtensorMap *tmap,
                                          const uint32_t t, double *brick, double *staged,
                                          uint64_t *full_p, double (*red)[32])
{
    using S = Stencil<DIM, ST>;
    using B = Brick<DIM>;
    constexpr int K = S::K;
    constexpr int NSR = tma_popc(M::READ_MASK);
    const StepCommon &P = A.k;
    const Grid &g = P.g;
    const int tid = threadIdx.x;
    const bool tiled = (g.line & 31) == 0;
    uint64_t &full = *full_p;
    // tile mode: block t owns spatial tile t.  Packed mode (tissue that fills its tiles
    // poorly -- a ventricle wall, heavy fibrosis; single GPU): block t owns the 256
    // consecutive compact nodes [256 t, 256 t + 256), every warp is full, u by plain loads.
    const bool packed = !HALO && P.node_of != nullptr;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (packed) {
        rec.x = t * (uint32_t)BLOCK_THREADS;
        const uint32_t np = (uint32_t)P.n_packed;                 // (< 2^31: node ids are int32)
        rec.y = np > rec.x ? min(np - rec.x, (uint32_t)BLOCK_THREADS) : 0u;
    } else {
        rec = __ldg(P.tile_rec + t);
    }
    const uint32_t c0 = rec.x, cnt = rec.y;
    const int64_t n0 = (int64_t)rec.z * 32;
    const bool bnd = HALO && rec.w != 0;
    const bool use_brick = BRICK && !SLOW && !bnd;   // (BRICK kernels: tiled grids, tile mode)
    const uint32_t c0a = c0 & ~1u;

    if (!SLOW && tid == 0) {
        const unsigned bytes = cnt ? (((c0 + cnt - c0a) + 1u) & ~1u) * 8u : 0u;
        mbar_arrive_expect_tx(&full, bytes * NSR + (use_brick ? (unsigned)(B::N * 8) : 0u));
        if (use_brick) {
            // tile origin (plane, row, column) from the chunk id of its slot 0
            const int64_t cpl = g.line >> 5;
            const int64_t o = rec.z;
            const int col0 = (int)(o % cpl) * 32;
            if (DIM == 3) {
                const int row0 = (int)((o / cpl) % g.rows), pl0 = (int)(o / (cpl * g.rows));
                tma_load_brick3(brick, tmap, col0 - B::X0, row0 - 1, pl0 - 1, &full);
            } else {
                tma_load_brick2(brick, tmap, col0 - B::X0, (int)(o / cpl) - 1, &full);
            }
        }
        if (bytes) {
#pragma unroll
            for (int q = 0; q < NSR; ++q)
                tma_load_1d(staged + q * TMA_SEG,
                            P.state + (int64_t)tma_nth(M::READ_MASK, q) * g.ld + c0a, bytes, &full);
        }
    }

    // which slab boundary (if any) this tile belongs to
    const HaloSide *side = nullptr;
    if (HALO) {
        if (P.halo.lo.on && t - P.halo.lo.first_block < P.halo.lo.n_blocks) side = &P.halo.lo;
        else if (P.halo.hi.on && t - P.halo.hi.first_block < P.halo.hi.n_blocks) side = &P.halo.hi;
        if (side) {
            if (tid == 0)
                while (ld_acquire_sys(side->flag) < P.halo.epoch) __nanosleep(64);
            __syncthreads();
        }
    }

    // compact lanes: thread r owns the r-th updated node of the tile.  A partially
    // filled warp runs whole (its spare lanes shadow the tile's last node and never
    // store); warps beyond the node count skip the arithmetic altogether.
    const bool act = (uint32_t)tid < cnt;
    const bool wact = (uint32_t)(tid & ~31) < cnt;
    const uint32_t c = c0 + (act ? (uint32_t)tid : cnt - 1u);
    int64_t n = 0;
    int q = 0, ln = 0;
    double acc = 0.0, uc = 0.0;
    if (wact) {
        if (packed) {
            n = __ldg(P.node_of + c);
        } else {
            const unsigned pos = __ldg(P.pos_of + c);
            q = pos >> 5; ln = pos & 31;
            n = tile_node<DIM>(g, tiled, bnd, n0, q, ln);
        }
        // diffusion: left-to-right sum in slot order, no FMA contraction (-fmad=false).
        // Brick tiles: the weights are requested first (one coalesced streaming load per
        // slot, all in flight together), the u operands are shared-memory reads once the
        // brick has landed.  Other tiles: both operands are plain loads, issued ahead by
        // the compiler as far as the registers allow.
        const double *__restrict__ w = P.w + c;
        if (use_brick) {
            double wn[K];
#pragma unroll
            for (int k = 0; k < K; ++k) wn[k] = ld_stream(w + (int64_t)k * g.ld);
            mbar_wait(&full, 0u);     // (the fast kernel owns a single tile: phase 0)
            const double *bc = brick + B::centre(q, ln);
            acc = mul(bc[B::off(S::at(0))], wn[0]);
#pragma unroll
            for (int k = 1; k < K; ++k) acc = add(acc, mul(bc[B::off(S::at(k))], wn[k]));
            uc = bc[0];
        } else {
            const double *__restrict__ u = P.u + n;
            const bool ghost = HALO && side;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const Off o = S::at(k);
                const int64_t off = (DIM == 3 ? (int64_t)o.p * g.s_plane : 0) +
                                    (int64_t)o.r * g.s_row + o.l;
                // ghost values were stored by a peer GPU: not through the non-coherent path
                const double uk = ghost ? __ldcg(u + off) : __ldg(u + off);
                const double pk = mul(uk, ld_stream(w + (int64_t)k * g.ld));
                acc = k == 0 ? pk : add(acc, pk);
                if (o.p == 0 && o.r == 0 && o.l == 0) uc = uk;
            }
            if (!SLOW) mbar_wait(&full, 0u);
        }
    }
    const double diff = acc - uc;      // u_tr - u of this node (ECG)

    bool handled = true;
    if (SLOW) {
        if (wact) {
            StateIO io{P.state + c, g.ld, act};
            M::ionic_ref(uc, acc, io, A.c);
        }
    } else {
        StateIOTma<M> io{staged + (c - (int64_t)c0a), P.state + c, g.ld, act};
        const bool ok = !wact || M::fast_ok(uc, io, A.c);
        if (__syncthreads_or(!ok)) {
            // not this kernel's business: the reference-statement kernel redoes the tile
            if (tid == 0) P.defer_list[atomicAdd(P.defer_ctr, 1u)] = t;
            handled = false;
        } else if (wact) {
            M::ionic_fastpath(uc, acc, io, A.c);
        }
    }
    if (!handled) return;

    if (act) {
        if (TRACK && P.do_act) {
            // ActivationTime{2,3}DTracker._track on the updated nodes (strict >); nodes the
            // solver never updates are covered by the full-grid samples (sim.cu)
            const double a = P.act_t[n];
            if (a < 0 && uc > P.act_thr) P.act_t[n] = P.t;
        }
        P.u_new[n] = acc;
        if (HALO && side) side->peer_dst[n - side->first] = acc;
    }

    if (HALO && side) {
        // publish: all of this block's peer stores are visible system-wide before the
        // block is counted; the last block of the side raises the neighbour's flag
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const unsigned done = atomicAdd(side->counter, 1u) + 1u;
            if (done == side->n_blocks) {
                *side->counter = 0;
                __threadfence_system();
                st_release_sys(side->peer_flag, P.halo.epoch + 1u);
            }
        }
    }

    if (TRACK && P.do_ecg) {
        // ECG{2,3}DTracker: per lead sum over updated nodes of (u_tr - u) / (d * dr),
        // d = squared index distance, d == 0 skipped; deterministic order (warp tree ->
        // fixed-order sum over warps -> per-tile partial, summed by ecg_finalize)
        const int lane = tid & 31, warp = tid >> 5;
        double ci = 0, cj = 0, ck = 0;
        if (act) {
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
                if (act) {
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
            if (tid < nl) {
                double sum = 0.0;
                for (int wq = 0; wq < WARPS_PER_BLOCK; ++wq) sum += red[wq][tid];
                P.ecg_partial[(int64_t)t * P.n_leads + l0 + tid] = sum;
            }
            __syncthreads();
        }
    }
}

template <class M, int DIM, int ST, bool TRACK, bool HALO, bool BRICK, bool SLOW>
__global__ void __launch_bounds__(BLOCK_THREADS, (SLOW ? 1 : M::MIN_BLOCKS))
step_kernel_tile(const __grid_constant__ StepArgs<M> A, const __grid_constant__ CUtensorMap tmap)
{
    const StepCommon &P = A.k;
    const int tid = threadIdx.x;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    double *const brick = reinterpret_cast<double *>(dyn_smem);                       // [Brick::N]
    double *const staged = reinterpret_cast<double *>(dyn_smem + (BRICK ? BRICK_BYTES_MAX : 0));
    __shared__ uint64_t full;
    __shared__ double red[TRACK ? WARPS_PER_BLOCK : 1][32];

    if (!SLOW) {
        // the fast kernel: block b owns tile b
#ifdef FWB_EXP_SMEM
        exp_table_to_smem();
#endif
        if (tid == 0) {
            mbar_init(&full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        tile_body<M, DIM, ST, TRACK, HALO, BRICK, SLOW>(A, &tmap, blockIdx.x, brick, staged, &full, red);
    } else {
        // the reference-statement kernel walks the list of handed-over tiles (normally empty)
        const uint32_t n_def = *reinterpret_cast<volatile const uint32_t *>(P.defer_ctr);
        for (uint32_t it = blockIdx.x; it < n_def; it += gridDim.x) {
            tile_body<M, DIM, ST, TRACK, HALO, BRICK, SLOW>(A, &tmap, P.defer_list[it], brick, staged,
                                                            &full, red);
            __syncthreads();
        }
        // the last block to finish empties the list for the next step
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned done = atomicAdd(P.defer_ctr + 1, 1u) + 1u;
            if (done == gridDim.x) {
                P.defer_ctr[0] = 0;
                P.defer_ctr[1] = 0;
                __threadfence();
            }
        }
    }
}


// ---------------------------------------------------------------------------
// step_kernel_tma: persistent, TMA-fed variant (see the header comment)
// ---------------------------------------------------------------------------

template <class M, int K, int DIM> struct TmaCfg {
    static constexpr int NSR = tma_popc(M::READ_MASK);
    static constexpr int NARR = K + NSR;
    static constexpr int NSTAGE = 2;
    // one ring stage: the tile's u brick (tensor TMA; first, 128-byte aligned), its weight and
    // state rows, its 8 work records -- rounded up to a multiple of 128 bytes
    static constexpr size_t BRICK = ((size_t)Brick<DIM>::N * sizeof(double) + 127) / 128 * 128;
    static constexpr size_t ROWS = (size_t)NARR * TMA_SEG * sizeof(double);
    static constexpr size_t STAGE = (BRICK + ROWS + TMA_REC_BYTES + 127) / 128 * 128;
    static constexpr size_t STAGE_NB = (ROWS + TMA_REC_BYTES + 127) / 128 * 128;   // no brick
    static constexpr size_t SMEM = NSTAGE * STAGE + 64;
    static constexpr size_t SMEM_NB = NSTAGE * STAGE_NB + 64;
    // blocks per SM the launch bounds ask for: as many as the ring allows (227 KB / SM)
    static constexpr int BLOCKS = SMEM * 4 <= 224 * 1024 ? 4 : (SMEM * 3 <= 224 * 1024 ? 3 : 2);
};

template <class M, int DIM, int ST, bool TRACK, bool HALO>
__global__ void __launch_bounds__(BLOCK_THREADS, (TmaCfg<M, Stencil<DIM, ST>::K, DIM>::BLOCKS))
step_kernel_tma(const __grid_constant__ StepArgs<M> A, const __grid_constant__ CUtensorMap tmap)
{
    using S = Stencil<DIM, ST>;
    using B = Brick<DIM>;
    constexpr int K = S::K;
    using C = TmaCfg<M, K, DIM>;
    constexpr int NARR = C::NARR, NSTAGE = C::NSTAGE;

    const StepCommon &P = A.k;
    const Grid &g = P.g;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // a ring stage holds the tile's rows and records, preceded -- only when bricks are on, so
    // that the L1 keeps the shared memory it would take otherwise -- by the tile's u brick
    const bool bricks = P.brick != 0 && P.tile_rec != nullptr;
    const int BRICK_D = bricks ? (int)(C::BRICK / sizeof(double)) : 0;
    const int STAGE_D = (int)((bricks ? C::STAGE : C::STAGE_NB) / sizeof(double));

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
    // tiles of a tiled grid that are not slab-boundary tiles get their u neighbourhood as a
    // brick by tensor TMA (tile_rec: origin chunk + boundary flag); the others use plain loads
    auto issue = [&](int64_t t, int s, uint32_t c0, uint32_t c1) {
        const uint32_t c0a = c0 & ~1u;
        const unsigned bytes = c1 > c0 ? (((c1 - c0a) + 1u) & ~1u) * 8u : 0u;
        uint4 trec = make_uint4(0u, 0u, 0u, 1u);
        if (bricks && lane == 0) trec = __ldg(P.tile_rec + t);
        const bool brick_t = bricks && trec.w == 0u;        // (meaningful in lane 0 only)
        if (lane == 0)
            mbar_arrive_expect_tx(full + s, bytes * NARR + TMA_REC_BYTES +
                                                (brick_t ? (unsigned)(B::N * 8) : 0u));
        __syncwarp();
        double *stage = ring + (size_t)s * STAGE_D;
        if (lane == 0 && brick_t) {
            const int64_t cpl = g.line >> 5;
            const int64_t o = trec.z;
            const int col0 = (int)(o % cpl) * 32;
            if (DIM == 3) {
                const int row0 = (int)((o / cpl) % g.rows), pl0 = (int)(o / (cpl * g.rows));
                tma_load_brick3(stage, &tmap, col0 - B::X0, row0 - 1, pl0 - 1, full + s);
            } else {
                tma_load_brick2(stage, &tmap, col0 - B::X0, (int)(o / cpl) - 1, full + s);
            }
        }
        if (lane == 31)
            tma_load_1d(stage + BRICK_D + NARR * TMA_SEG, P.records + t * WARPS_PER_BLOCK,
                        TMA_REC_BYTES, full + s);
        if (bytes) {
            for (int a = lane; a < NARR; a += 32) {
                const double *src = a < K
                    ? P.w + (int64_t)a * g.ld + c0a
                    : P.state + (int64_t)tma_nth(M::READ_MASK, a - K) * g.ld + c0a;
                tma_load_1d(stage + BRICK_D + a * TMA_SEG, src, bytes, full + s);
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
        const double *stage = ring + (size_t)s * STAGE_D + BRICK_D;    // the tile's rows
        const double *brick = ring + (size_t)s * STAGE_D;
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
        // Whole sectors: a partially written 32-byte sector makes L2 fetch the rest from DRAM
        // (30 % fibrosis: three sectors in four).  Where the host has checked that both
        // potential buffers agree on the nodes the solver does not update, the idle lanes of a
        // listed chunk carry the value that is there already and the warp stores its 256 bytes
        // with ONE instruction below (two partial stores per sector would still fill).
        const bool idle = P.copy_idle && !myo && chunk >= 0 && n < g.n_nodes;
        double out = 0.0;
        if (idle) out = __ldg(P.u + n);

        double un[K];
        if (myo) {
            if (bricks && !(HALO && side)) {
                // warp = tile slot, lane = column: the brick's compile-time offsets
                const double *bc = brick + B::centre(warp, lane);
#pragma unroll
                for (int k = 0; k < K; ++k) un[k] = bc[B::off(S::at(k))];
            } else {
                const double *__restrict__ u = P.u + n;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const Off o = S::at(k);
                    const int64_t off = (DIM == 3 ? (int64_t)o.p * g.s_plane : 0) +
                                        (int64_t)o.r * g.s_row + o.l;
                    un[k] = (HALO && side) ? __ldcg(u + off) : __ldg(u + off);
                }
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

            out = acc;
            if (HALO && side) side->peer_dst[n - side->first] = acc;
        }
        if (myo || idle) P.u_new[n] = out;

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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device property of a kernel: set
// it once per (kernel, device), whichever thread or GPU the caller drives
inline int ensure_dyn_smem_impl(const void *kern, size_t bytes)
{
    // keyed by the kernel's ADDRESS (instantiations of one template share a pointer type)
    static std::mutex mu;
    static std::map<const void *, unsigned long long> done;   // bit d: set on device d
    int dev = 0;
    FWB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    unsigned long long &mask = done[kern];
    if (dev < 64 && ((mask >> dev) & 1ull)) return 0;
    FWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (dev < 64) mask |= 1ull << dev;
    return 0;
}
template <class Kern> static int ensure_dyn_smem(Kern kern, size_t bytes)
{
    return ensure_dyn_smem_impl(reinterpret_cast<const void *>(kern), bytes);
}

template <class M, int DIM, int ST, bool TRACK, bool HALO>
static int launch_one(const StepCommon &k, const void *consts, cudaStream_t s)
{
    StepArgs<M> a;
    a.k = k;
    a.c = *reinterpret_cast<const typename M::Consts *>(consts);
    int64_t blocks = step_blocks(k.g);
    if (blocks <= 0) return 0;
    if constexpr (stage_state<M>()) {
        if (k.tile_rec && k.pos_of && k.defer_list && k.defer_ctr) {
            const bool packed = !HALO && k.node_of != nullptr;
            if (packed) blocks = (k.n_packed + BLOCK_THREADS - 1) / BLOCK_THREADS;
            if (blocks <= 0) return 0;
            // compact-lane tile kernel (fast path only) + the reference-statement kernel for
            // the tiles it hands over (normally none: that launch returns at once)
            CUtensorMap none;
            memset(&none, 0, sizeof(none));
            const bool brick = !packed && k.brick && k.tmap_host && (k.g.line & 31) == 0;
            int rc;
            if (brick) {
                auto kern = step_kernel_tile<M, DIM, ST, TRACK, HALO, true, false>;
                if ((rc = ensure_dyn_smem(kern, TileCfg<M, true>::SMEM))) return rc;
                kern<<<(unsigned)blocks, BLOCK_THREADS, TileCfg<M, true>::SMEM, s>>>(
                    a, *reinterpret_cast<const CUtensorMap *>(k.tmap_host));
                note_step_variant(5);
            } else {
                auto kern = step_kernel_tile<M, DIM, ST, TRACK, HALO, false, false>;
                if ((rc = ensure_dyn_smem(kern, TileCfg<M, false>::SMEM))) return rc;
                kern<<<(unsigned)blocks, BLOCK_THREADS, TileCfg<M, false>::SMEM, s>>>(a, none);
                note_step_variant(4);
            }
            FWB_KERNEL_CHECK("step_kernel_tile");
            const int64_t slow_blocks = blocks < 296 ? blocks : 296;
            step_kernel_tile<M, DIM, ST, TRACK, HALO, false, true>
                <<<(unsigned)slow_blocks, BLOCK_THREADS, 0, s>>>(a, none);
            FWB_KERNEL_CHECK("step_kernel_tile (reference statement)");
            return 0;
        }
    }
    step_kernel<M, DIM, ST, TRACK, HALO><<<(unsigned)blocks, BLOCK_THREADS, 0, s>>>(a);
    note_step_variant(0);
    FWB_KERNEL_CHECK("step_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// step_kernel_small: tissues of a few thousand nodes (the README quick start is 100 x 100)
//
// At this size one time step is ~2 us of launch latency around ~0.2 us of work.  This kernel
// runs MANY steps in one launch: a single thread-block cluster (<= 8 CTAs x 1024 threads on 8
// SMs) owns the whole tissue.  The two potential buffers live in DISTRIBUTED SHARED MEMORY for
// the whole run -- CTA k holds the dense nodes [k * chunk, (k + 1) * chunk), chunk a power of
// two -- every thread keeps the state of its (<= 4) nodes in its CTA's shared memory, the
// stencil operands are `ld.shared::cluster` reads (mapa: node -> owning CTA), u_new is a
// `st.shared::cluster` store, and the steps are separated by a hardware cluster barrier
// (release / acquire at cluster scope) instead of a kernel boundary or an L2 round trip.
// The arithmetic is the per-step kernels' own (same slot order, same Model::ionic), so the
// result is bit-identical; the host (fwb_sim_run) only uses it for runs of steps in which no
// stimulus fires and only the activation-time tracker samples.  t advances by the same
// repeated fp64 addition as on the host.
// ---------------------------------------------------------------------------
struct SmallArgs {
    double *buf[2];            // the two u buffers; buf[cur] is u at the first step
    int cur;
    int n_steps;
    int npt;                   // nodes per thread
    int chunk;                 // dense nodes per CTA (a multiple of 32, <= 4096)
    int64_t n_myo;
    int64_t step0;
    double t0;
    // activation-time tracker (at most one, already primed)
    double *act_t;
    double act_thr, act_start, act_end;
    int64_t act_every;
};
constexpr int SMALL_THREADS = 1024, SMALL_MAX_CTAS = 16, SMALL_MAX_NPT = 4, SMALL_MAX_CHUNK = 4096;
constexpr size_t SMALL_SMEM_MAX = 220 * 1024;     // dynamic shared memory of one CTA

// state of one node in shared memory: slot q of node-slot j of thread tid
struct StateIOSmem {
    double *s;                 // &smem[(j * NS + 0) * SMALL_THREADS + tid]
    __device__ __forceinline__ double ld(int q) const { return s[q * SMALL_THREADS]; }
    __device__ __forceinline__ void st(int q, double v) const { s[q * SMALL_THREADS] = v; }
};

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// distributed shared memory: the word at local shared address `laddr` of CTA `rank`
__device__ __forceinline__ double dsmem_ld(uint32_t laddr, uint32_t rank)
{
    uint32_t ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(laddr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void dsmem_st(uint32_t laddr, uint32_t rank, double v)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

template <class M, int DIM, int ST>
__global__ void __launch_bounds__(SMALL_THREADS, 1)
step_kernel_small(const __grid_constant__ StepArgs<M> A, const __grid_constant__ SmallArgs sa)
{
    using S = Stencil<DIM, ST>;
    constexpr int K = S::K;
    constexpr int NS = M::NS > 0 ? M::NS : 1;
    const StepCommon &P = A.k;
    const Grid &g = P.g;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    const int chunk = sa.chunk;                                       // <= npt * SMALL_THREADS
    double *const ub = reinterpret_cast<double *>(dyn_smem);          // [2][chunk] this CTA's nodes
    double *const st_sm = ub + 2 * chunk;                             // [npt][NS][SMALL_THREADS]
    // (the cluster barrier invalidates L1, so everything a step needs lives in shared memory)
    double *const w_sm = st_sm + (size_t)sa.npt * NS * SMALL_THREADS;  // [npt][K][SMALL_THREADS]
    const int tid = threadIdx.x;
    const uint32_t rank = blockIdx.x;                                 // one cluster: rank = block
    const uint32_t ub_addr = smem_u32(ub);
    const int64_t base = (int64_t)rank * chunk;
    exp_table_to_smem();

    // CTA `rank` owns the dense nodes [base, base + chunk): both potential buffers go to its
    // shared memory, thread tid owns the nodes base + tid + j * 1024 (the updated ones among
    // them: their state and weight rows go to shared memory too)
    for (int i = tid; i < chunk; i += SMALL_THREADS) {
        const int64_t m = base + i;
        ub[i] = m < g.n_nodes ? sa.buf[0][m] : 0.0;
        ub[chunk + i] = m < g.n_nodes ? sa.buf[1][m] : 0.0;
    }
    uint32_t live = 0;                 // bit j: node-slot j is an updated node
    for (int j = 0; j < sa.npt; ++j) {
        const int loc0 = tid + j * SMALL_THREADS;
        const int64_t m = base + loc0;
        if (loc0 >= chunk || m >= g.n_nodes) continue;
        const uint32_t bits = g.chunk_bits[m >> 5];
        const int ln = (int)(m & 31);
        if (!((bits >> ln) & 1u)) continue;
        const int64_t c = (int64_t)g.chunk_base[m >> 5] + __popc(bits & ((1u << ln) - 1u));
        live |= 1u << j;
#pragma unroll
        for (int q = 0; q < M::NS; ++q)
            st_sm[(j * NS + q) * SMALL_THREADS + tid] = P.state[(int64_t)q * g.ld + c];
#pragma unroll
        for (int k = 0; k < K; ++k)
            w_sm[(j * K + k) * SMALL_THREADS + tid] = P.w[(int64_t)k * g.ld + c];
    }
    cluster_sync_all();

    double t = sa.t0;
    int cur = sa.cur;
    int64_t act_rem = sa.act_every > 0 ? sa.step0 % sa.act_every : 0;
    for (int it = 0; it < sa.n_steps; ++it) {
        const double *usrc = ub + (size_t)cur * chunk;
        double *udst = ub + (size_t)(cur ^ 1) * chunk;
        const uint32_t src = ub_addr + (uint32_t)cur * (uint32_t)chunk * 8u;
        const bool do_act = sa.act_t && !(sa.act_start > t || t > sa.act_end) && act_rem == 0;
        act_rem = act_rem + 1 == sa.act_every ? 0 : act_rem + 1;
#ifdef FWB_SMALL_NOWORK          // (timing experiments only)
        for (int j = 0; j < 0; ++j) {
#else
        for (int j = 0; j < sa.npt; ++j) {
#endif
            if (!((live >> j) & 1u)) continue;
            const int32_t loc = tid + j * SMALL_THREADS;              // node inside this CTA's range
            const double *w = w_sm + (j * K) * SMALL_THREADS + tid;
            double un[K], wn[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const Off o = S::at(k);
                const int32_t off = (DIM == 3 ? o.p * (int32_t)g.s_plane : 0) +
                                    o.r * (int32_t)g.s_row + o.l;
                const int32_t nb = loc + off;
                // most neighbours are this CTA's own: plain shared memory; the others (near
                // the ends of the range) come from the neighbour CTA's shared memory
                if ((uint32_t)nb < (uint32_t)chunk) {
                    un[k] = usrc[nb];
                } else {
                    const uint32_t m = (uint32_t)((int32_t)base + nb);
                    const uint32_t owner = m / (uint32_t)chunk;
                    un[k] = dsmem_ld(src + (m - owner * (uint32_t)chunk) * 8u, owner);
                }
                wn[k] = w[k * SMALL_THREADS];
            }
            double acc = mul(un[0], wn[0]);
#pragma unroll
            for (int k = 1; k < K; ++k) acc = add(acc, mul(un[k], wn[k]));
            double uc = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (S::at(k).p == 0 && S::at(k).r == 0 && S::at(k).l == 0) uc = un[k];
            if (do_act) {
                const int64_t n = base + loc;
                const double a = sa.act_t[n];
                if (a < 0 && uc > sa.act_thr) sa.act_t[n] = t;
            }
            StateIOSmem io{st_sm + (j * NS) * SMALL_THREADS + tid};
            M::ionic(uc, acc, io, A.c);
            udst[loc] = acc;
        }
        t += A.c.dt;
        cur ^= 1;
#ifndef FWB_SMALL_NOBAR          // (timing experiments only)
        cluster_sync_all();
#endif
    }

    // both buffers and the state back to global memory
    for (int i = tid; i < chunk; i += SMALL_THREADS) {
        const int64_t m = base + i;
        if (m < g.n_nodes) {
            sa.buf[0][m] = ub[i];
            sa.buf[1][m] = ub[chunk + i];
        }
    }
    for (int j = 0; j < sa.npt; ++j) {
        if (!((live >> j) & 1u)) continue;
        const int64_t m = base + tid + j * SMALL_THREADS;
        const uint32_t bits = g.chunk_bits[m >> 5];
        const int64_t c = (int64_t)g.chunk_base[m >> 5] + __popc(bits & ((1u << (int)(m & 31)) - 1u));
#pragma unroll
        for (int q = 0; q < M::NS; ++q)
            if ((M::WRITE_MASK >> q) & 1u)
                P.state[(int64_t)q * g.ld + c] = st_sm[(j * NS + q) * SMALL_THREADS + tid];
    }
}

template <class M, int DIM, int ST>
static int launch_small_one(const StepCommon &k, const void *consts, const SmallArgs &sa_in,
                            cudaStream_t s)
{
    StepArgs<M> a;
    a.k = k;
    a.c = *reinterpret_cast<const typename M::Consts *>(consts);
    SmallArgs sa = sa_in;
    // CTA r owns the dense nodes [r * chunk, (r + 1) * chunk).  As many CTAs (= SMs: the step
    // is FP64-throughput-bound inside each of them) as one cluster may have -- 16 where the
    // device schedules a non-portable cluster of this kernel, else 8 -- but not more than one
    // per 256 nodes
    auto kern = step_kernel_small<M, DIM, ST>;
    int rc;
    if ((rc = ensure_dyn_smem(kern, SMALL_SMEM_MAX))) return rc;
    static std::mutex mu;
    static int max_ctas_of[64] = {0};
    int dev = 0;
    FWB_CUDA(cudaGetDevice(&dev));
    int max_ctas = 8;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 64 && max_ctas_of[dev]) {
            max_ctas = max_ctas_of[dev];
        } else {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t q;
                memset(&q, 0, sizeof(q));
                q.gridDim = dim3(SMALL_MAX_CTAS); q.blockDim = dim3(SMALL_THREADS);
                q.dynamicSmemBytes = SMALL_SMEM_MAX;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = SMALL_MAX_CTAS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                q.attrs = qa; q.numAttrs = 1;
                int n_clusters = 0;
                if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q) == cudaSuccess && n_clusters >= 1)
                    max_ctas = SMALL_MAX_CTAS;
            }
            cudaGetLastError();
            if (dev < 64) max_ctas_of[dev] = max_ctas;
        }
    }
    int ctas = (int)((k.g.n_nodes + 255) / 256);
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    int chunk = (int)(((k.g.n_nodes + ctas - 1) / ctas + 31) / 32 * 32);
    if (chunk > SMALL_MAX_CHUNK) { set_error("step_kernel_small: tissue too large"); return FWB_E_UNSUPPORTED; }
    ctas = (int)((k.g.n_nodes + chunk - 1) / chunk);
    sa.chunk = chunk;
    sa.npt = (chunk + SMALL_THREADS - 1) / SMALL_THREADS;
    constexpr int NS = M::NS > 0 ? M::NS : 1;
    constexpr int K = Stencil<DIM, ST>::K;
    const size_t smem = (size_t)2 * chunk * sizeof(double) +
                        (size_t)sa.npt * (NS + K) * SMALL_THREADS * sizeof(double);
    if (smem > SMALL_SMEM_MAX) {
        set_error("step_kernel_small: tissue too large for shared memory");
        return FWB_E_UNSUPPORTED;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(SMALL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FWB_CUDA(cudaLaunchKernelEx(&cfg, kern, a, sa));
    note_step_variant(6);
    return 0;
}

template <class M>
static int launch_small_model(int dim, int stencil, const StepCommon &k, const void *consts,
                              const SmallArgs &sa, cudaStream_t s)
{
    if constexpr (M::NS <= 4) {
        if (dim == 2 && stencil == FWB_STENCIL_ISO) return launch_small_one<M, 2, FWB_STENCIL_ISO>(k, consts, sa, s);
        if (dim == 2 && stencil == FWB_STENCIL_ANISO) return launch_small_one<M, 2, FWB_STENCIL_ANISO>(k, consts, sa, s);
        if (dim == 3 && stencil == FWB_STENCIL_ISO) return launch_small_one<M, 3, FWB_STENCIL_ISO>(k, consts, sa, s);
        if (dim == 3 && stencil == FWB_STENCIL_ANISO) return launch_small_one<M, 3, FWB_STENCIL_ANISO>(k, consts, sa, s);
    }
    set_error("step_kernel_small: unsupported model / stencil");
    return FWB_E_UNSUPPORTED;
}

template <class M, int DIM, int ST, bool TRACK, bool HALO>
static int launch_tma(const StepCommon &k, const void *consts, cudaStream_t s)
{
    using C = TmaCfg<M, Stencil<DIM, ST>::K, DIM>;
    auto kern = step_kernel_tma<M, DIM, ST, TRACK, HALO>;
    // u bricks in the ring: on for 3D tiled grids, off in 2D (A/B on a B200, profiles/:
    // C3 +2 %, C2 -7 %); FWB_RING_BRICK=0/1 overrides
    bool brick = k.brick && k.tmap_host && k.tile_rec && (k.g.line & 31) == 0 && DIM == 3;
    {
        const char *e = getenv("FWB_RING_BRICK");
        if (e && e[0] == '0') brick = false;
        if (e && e[0] == '1') brick = k.brick && k.tmap_host && k.tile_rec && (k.g.line & 31) == 0;
    }
    const size_t smem = brick ? C::SMEM : C::SMEM_NB;
    // blocks of this instantiation each device holds at once (per device and per mode: a
    // process may drive several GPUs, from several threads)
    static std::mutex mu;
    static int resident_of[64][2] = {{0, 0}};
    int dev = 0;
    FWB_CUDA(cudaGetDevice(&dev));
    int resident = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || resident_of[dev][brick] == 0) {
            FWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)C::SMEM));
            int sms = 0, per_sm = 0;
            FWB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            FWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK_THREADS,
                                                                   smem));
            if (per_sm < 1) { set_error("step_kernel_tma does not fit on an SM"); return FWB_E_UNSUPPORTED; }
            resident = sms * per_sm;
            if (dev < 64) resident_of[dev][brick] = resident;
        } else {
            resident = resident_of[dev][brick];
        }
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
    CUtensorMap map;
    if (brick) map = *reinterpret_cast<const CUtensorMap *>(k.tmap_host);
    else memset(&map, 0, sizeof(map));
    a.k.brick = brick ? 1 : 0;
    kern<<<(unsigned)blocks, BLOCK_THREADS, smem, s>>>(a, map);
    note_step_variant(brick ? 7 : 3);
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

// load every instantiation the step loop can launch for (dim, stencil) -- see
// preload_aux_kernels() in aux_kernels.cu for why
template <class K> static int preload_kernel(K kern)
{
    cudaFuncAttributes a;
    FWB_CUDA(cudaFuncGetAttributes(&a, kern));
    return 0;
}
template <class M, int DIM, int ST, bool TRACK, bool HALO> static int preload_variants()
{
    int rc;
    if ((rc = preload_kernel(step_kernel<M, DIM, ST, TRACK, HALO, 0>))) return rc;
    if constexpr (M::USE_TMA) {
        if ((rc = preload_kernel(step_kernel_tma<M, DIM, ST, TRACK, HALO>))) return rc;
    }
    if constexpr (stage_state<M>()) {
        if ((rc = preload_kernel(step_kernel_tile<M, DIM, ST, TRACK, HALO, true, false>))) return rc;
        if ((rc = preload_kernel(step_kernel_tile<M, DIM, ST, TRACK, HALO, false, false>))) return rc;
        if ((rc = preload_kernel(step_kernel_tile<M, DIM, ST, TRACK, HALO, false, true>))) return rc;
    }
    return 0;
}
template <class M, int DIM, int ST> static int preload_stencil()
{
    int rc;
    if ((rc = preload_variants<M, DIM, ST, false, false>())) return rc;
    if ((rc = preload_variants<M, DIM, ST, true, false>())) return rc;
    if ((rc = preload_variants<M, DIM, ST, false, true>())) return rc;
    if ((rc = preload_variants<M, DIM, ST, true, true>())) return rc;
    return 0;
}
template <class M> static int preload_model(int dim, int stencil)
{
    if (dim == 2 && stencil == FWB_STENCIL_ISO) return preload_stencil<M, 2, FWB_STENCIL_ISO>();
    if (dim == 2 && stencil == FWB_STENCIL_ANISO) return preload_stencil<M, 2, FWB_STENCIL_ANISO>();
    if (dim == 3 && stencil == FWB_STENCIL_ISO) return preload_stencil<M, 3, FWB_STENCIL_ISO>();
    if (dim == 3 && stencil == FWB_STENCIL_ANISO) return preload_stencil<M, 3, FWB_STENCIL_ANISO>();
    return 0;
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
struct SmallArgs;
typedef int (*LaunchSmallFn)(int dim, int stencil, const StepCommon &k, const void *consts,
                             const SmallArgs &sa, cudaStream_t s);
struct ModelEntry {
    int n_state, n_params;
    uint32_t read_mask, write_mask;
    LaunchFn launch;
    DeriveFn derive;
    LaunchSmallFn launch_small;    // multi-step cluster kernel for tiny tissues (light models)
    int (*preload)(int dim, int stencil);   // force-load every step kernel of (dim, stencil)
};
constexpr int CONSTS_BYTES = 1024;

const ModelEntry *model_entry(int model);   // NULL if unknown; index FWB_N_MODELS = NoModel

#define FWB_DEFINE_MODEL_ENTRY(NAME, M)                                                   \
    static_assert(sizeof(M::Consts) <= fwb::CONSTS_BYTES, "Consts too large");            \
    extern const fwb::ModelEntry NAME = {M::NS, M::NP, M::READ_MASK, M::WRITE_MASK,       \
                                         &fwb::launch_model<M>, &fwb::derive_model<M>,     \
                                         &fwb::launch_small_model<M>, &fwb::preload_model<M>};

}  // namespace fwb

// aux_kernels.cu -- index structures, layout conversions, stimuli, tracker helpers.
#include "aux_kernels.cuh"
#define FWB_NO_EXP_SMEM   // devmath_kernel measures the helpers with the table in global memory
#include "fexp.cuh"

namespace fwb {

// ---------------------------------------------------------------------------
// chunk bits + exclusive prefix popcount (replaces the int64 myo_indexes list of
// finitewave/core/tissue/cardiac_tissue.py:54-63)
// ---------------------------------------------------------------------------
__global__ void chunk_bits_kernel(const uint8_t *mask, int64_t n_nodes, uint32_t *bits,
                                  int64_t n_chunks)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_chunks) return;
    const int64_t n = gw * 32 + lane;
    const bool on = n < n_nodes && mask[n] != 0;
    const uint32_t b = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[gw] = b;
}

constexpr int SCAN_ITEMS = 4096;   // chunks per scan block (1024 threads x 4)

// pass 1: per-block popcount totals
__global__ void __launch_bounds__(1024) scan_totals_kernel(const uint32_t *bits, int64_t n_chunks,
                                                           unsigned long long *block_tot)
{
    __shared__ unsigned long long sh[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
    unsigned long long v = 0;
    for (int q = 0; q < 4; ++q) {
        const int64_t i = base + q * 1024 + threadIdx.x;
        if (i < n_chunks) v += __popc(bits[i]);
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_tot[blockIdx.x] = v;
    }
}

// pass 2: exclusive scan of the block totals (single thread block, serial over
// at most a few thousand entries per thread-strided segment)
__global__ void scan_blocks_kernel(unsigned long long *block_tot, int n_blocks,
                                   unsigned long long *total)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < n_blocks; ++i) {
            const unsigned long long v = block_tot[i];
            block_tot[i] = run;
            run += v;
        }
        *total = run;
    }
}

// pass 3: exclusive scan inside each block + block offset
__global__ void __launch_bounds__(1024) scan_final_kernel(const uint32_t *bits, int64_t n_chunks,
                                                          const unsigned long long *block_off,
                                                          uint32_t *base_out)
{
    __shared__ unsigned int warp_tot[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
    // thread handles 4 consecutive chunks
    const int64_t i0 = base + (int64_t)threadIdx.x * 4;
    unsigned int p[4], sum = 0;
    for (int q = 0; q < 4; ++q) {
        p[q] = (i0 + q < n_chunks) ? __popc(bits[i0 + q]) : 0;
        sum += p[q];
    }
    // inclusive warp scan of per-thread sums
    unsigned int inc = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_tot[lane];
        unsigned int winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;   // exclusive
    }
    __syncthreads();
    unsigned long long run = block_off[blockIdx.x] + warp_tot[warp] + (inc - sum);
    for (int q = 0; q < 4; ++q) {
        if (i0 + q < n_chunks) base_out[i0 + q] = (uint32_t)run;
        run += p[q];
    }
}

// ---------------------------------------------------------------------------
// work list: ids of the chunks that contain active nodes, in tile order
// ---------------------------------------------------------------------------
struct WlSeg {
    int64_t s0, s1;        // slice range of the segment (slowest axis)
    int64_t R, cpl;        // rows per slice, chunks per line
    int ts, tr, tc;        // tile = ts slices x tr rows x tc column segments (ts*tr*tc = 8)
    int64_t nTc, nTr;      // tiles along columns / rows
    int64_t n_pos;         // positions = tiles * 8
    int linear;            // 1: position == chunk id (line % 32 != 0)
    int64_t n_chunks, n_nodes;
};

__global__ void worklist_mark_kernel(WlSeg sg, const uint8_t *mask, int32_t *cand, uint32_t *flag)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= sg.n_pos) return;
    int64_t chunk = -1;
    if (sg.linear) {
        chunk = q < sg.n_chunks ? q : -1;
    } else {
        const int w = (int)(q & 7);
        const int64_t T = q >> 3;
        const int ws = w / (sg.tr * sg.tc), wr = (w / sg.tc) % sg.tr, wc = w % sg.tc;
        const int64_t Tc = T % sg.nTc, Tr = (T / sg.nTc) % sg.nTr, Ts = T / (sg.nTc * sg.nTr);
        const int64_t slice = sg.s0 + Ts * sg.ts + ws, row = Tr * sg.tr + wr, seg = Tc * sg.tc + wc;
        if (slice < sg.s1 && row < sg.R && seg < sg.cpl) chunk = (slice * sg.R + row) * sg.cpl + seg;
    }
    uint32_t on = 0;
    if (chunk >= 0) {
        const int64_t n0 = chunk * 32;
        const int cnt = (int)min((int64_t)32, sg.n_nodes - n0);
        for (int l = 0; l < cnt; ++l) on |= mask[n0 + l];
        on = on ? 1u : 0u;
    }
    // TILE granularity: the 8 positions of a tile (8 consecutive q, hence 8 neighbouring
    // lanes) are listed together as soon as one of them holds tissue; positions without
    // tissue (or outside the grid) are listed as -1.  A block of the step kernel therefore
    // owns exactly one spatial tile, which is what lets it fetch the tile's u brick with one
    // tensor-TMA copy and number its nodes compactly.
    cand[q] = on ? (int32_t)chunk : -1;
    flag[q] = on;
}
__global__ void worklist_tileflag_kernel(int64_t n_pos, uint32_t *flag)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t on = q < n_pos ? flag[q] : 0u;
    const unsigned b = __ballot_sync(0xffffffffu, on != 0);
    const unsigned lane = threadIdx.x & 31u;
    if (q < n_pos) flag[q] = ((b >> (lane & ~7u)) & 0xffu) ? 1u : 0u;
}

__global__ void worklist_scatter_kernel(int64_t n_pos, const int32_t *cand, const uint32_t *flag,
                                        const uint32_t *base, int32_t *out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pos) return;
    if (flag[q]) out[base[q]] = cand[q];
}

// compact order = work-list order (tile-contiguous weights / state rows)
__global__ void order_gather_kernel(const uint32_t *bits, const int32_t *wl, int64_t n_work,
                                    uint32_t *tmp)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_work) return;
    const int32_t c = wl[i];
    tmp[i] = c >= 0 ? bits[c] : 0u;
}
__global__ void order_scatter_kernel(const int32_t *wl, int64_t n_work, const uint32_t *ord,
                                     const uint32_t *tmp_bits, uint32_t total,
                                     uint32_t *chunk_base, uint32_t *tile_base, uint4 *records)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_work) return;
    if (i == n_work) { tile_base[n_work / 8] = total; return; }
    const int32_t c = wl[i];
    if (c >= 0) chunk_base[c] = ord[i];
    if ((i & 7) == 0) tile_base[i >> 3] = ord[i];
    // what the TMA kernel needs per work-list entry: chunk id, update bits, compact
    // index of the chunk's first node, compact index of the tile's first node
    if (records) records[i] = make_uint4((uint32_t)c, tmp_bits[i], ord[i], ord[i & ~(int64_t)7]);
}

// per tile {compact base, node count, chunk id of the tile's slot 0, 0} and per node its
// position inside the tile (slot << 5 | lane): what the compact-lane tile kernel needs to
// map thread -> node (step_kernel_tile in step_kernel.cuh)
struct TileGeo {
    int dim, tiled, halo_lo, halo_hi;
    int64_t S, R, cpl;           // slices, rows per slice (1 in 2D), chunks per line
    int64_t n_lo_tiles, n_hi_tiles;
};
__device__ __forceinline__ int64_t tile_slot_offset(const TileGeo &tg, bool bnd, int w)
{
    // chunk-id offset of slot w from slot 0 (mirrors worklist_mark_kernel / fwb_build_worklist)
    if (!tg.tiled) return w;
    int ts, tr, tc;
    if (tg.dim == 3) { if (bnd) { ts = 1; tr = 8; tc = 1; } else { ts = 2; tr = 4; tc = 1; } }
    else { if (bnd) { ts = 1; tr = 1; tc = 8; } else { ts = 8; tr = 1; tc = 1; } }
    (void)ts;
    const int ws = w / (tr * tc), wr = (w / tc) % tr, wc = w % tc;
    return (int64_t)ws * tg.R * tg.cpl + (int64_t)wr * tg.cpl + wc;
}
__global__ void order_tiles_kernel(TileGeo tg, const int32_t *wl, int64_t n_tiles,
                                   const uint32_t *bits, const uint32_t *chunk_base,
                                   const uint32_t *tile_base, uint4 *tile_rec, uint8_t *pos_of)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp per tile
    const int lane = threadIdx.x & 31;
    if (gw >= n_tiles) return;
    const bool bnd = gw < tg.n_lo_tiles + tg.n_hi_tiles;
    int64_t origin = -1;
    for (int w = 0; w < 8; ++w) {
        const int32_t c = wl[gw * 8 + w];
        if (c < 0) continue;
        if (origin < 0) origin = (int64_t)c - tile_slot_offset(tg, bnd, w);
        const uint32_t b = bits[c];
        if ((b >> lane) & 1u)
            pos_of[chunk_base[c] + __popc(b & ((1u << lane) - 1u))] = (uint8_t)((w << 5) | lane);
    }
    if (lane == 0) {
        const uint32_t c0 = tile_base[gw], c1 = tile_base[gw + 1];
        tile_rec[gw] = make_uint4(c0, c1 - c0, (uint32_t)(origin < 0 ? 0 : origin), bnd ? 1u : 0u);
    }
}

// slab halo: hold the stream until both neighbours have finished step `epoch - 1`
__global__ void halo_wait_kernel(const unsigned *flag_lo, const unsigned *flag_hi, unsigned epoch)
{
    if (threadIdx.x == 0) {
        if (flag_lo) while (*(volatile const unsigned *)flag_lo < epoch) __nanosleep(64);
        if (flag_hi) while (*(volatile const unsigned *)flag_hi < epoch) __nanosleep(64);
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------
// dense <-> compact
// ---------------------------------------------------------------------------
__global__ void gather_kernel(const double *dense, double *compact, int64_t n_chunks,
                              const uint32_t *bits, const uint32_t *base)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_chunks) return;
    const uint32_t b = bits[gw];
    if ((b >> lane) & 1u)
        compact[(int64_t)base[gw] + __popc(b & ((1u << lane) - 1u))] = dense[gw * 32 + lane];
}

__global__ void scatter_kernel(const double *compact, double *dense, double fill, int keep,
                               int64_t n_nodes, int64_t n_chunks, const uint32_t *bits,
                               const uint32_t *base)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_chunks) return;
    const uint32_t b = bits[gw];
    const int64_t n = gw * 32 + lane;
    if (n >= n_nodes) return;
    if ((b >> lane) & 1u)
        dense[n] = compact[(int64_t)base[gw] + __popc(b & ((1u << lane) - 1u))];
    else if (!keep)
        dense[n] = fill;
}

// how many nodes the solver does NOT update hold a value different from `fill`
__global__ void count_offfill_kernel(const double *dense, double fill, int64_t n_nodes,
                                     int64_t n_chunks, const uint32_t *bits,
                                     unsigned long long *count)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_chunks) return;
    const int64_t n = gw * 32 + lane;
    const bool off = n < n_nodes && !((bits[gw] >> lane) & 1u) &&
                     __double_as_longlong(dense[n]) != __double_as_longlong(fill);
    const unsigned m = __ballot_sync(0xffffffffu, off);
    if (lane == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

__global__ void weights_pack_kernel(const double *aos, double *soa, int K, int64_t ld,
                                    int64_t n_chunks, const uint32_t *bits, const uint32_t *base)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_chunks) return;
    const uint32_t b = bits[gw];
    if (!((b >> lane) & 1u)) return;
    const int64_t c = (int64_t)base[gw] + __popc(b & ((1u << lane) - 1u));
    const int64_t n = gw * 32 + lane;
    for (int k = 0; k < K; ++k) soa[(int64_t)k * ld + c] = aos[n * K + k];
}

__global__ void weights_unpack_kernel(const double *soa, double *aos, int K, int64_t ld,
                                      int64_t n_nodes, int64_t n_chunks, const uint32_t *bits,
                                      const uint32_t *base)
{
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_chunks) return;
    const uint32_t b = bits[gw];
    const int64_t n = gw * 32 + lane;
    if (n >= n_nodes) return;
    const bool on = (b >> lane) & 1u;
    const int64_t c = (int64_t)base[gw] + __popc(b & ((1u << lane) - 1u));
    // nodes the solver does not update show the reference's untouched row: all zero
    // except the +1 on the centre slot (isotropic_stencil_2d.py:66-67 and twins)
    const int centre = K == 5 ? 2 : K == 7 ? 3 : 4;
    for (int k = 0; k < K; ++k)
        aos[n * K + k] = on ? soa[(int64_t)k * ld + c] : (k == centre ? 1.0 : 0.0);
}

// ---------------------------------------------------------------------------
// stimuli (cpuwave{2D,3D}/stimulation/*.py `stimulate`), ROI sized
// ---------------------------------------------------------------------------
__device__ __forceinline__ void stim_apply(double *u, int64_t n, int mode, double value,
                                           double dt_value, int has_u_max, double u_max)
{
    if (mode == FWB_STIM_CURRENT) {
        double v = u[n] + dt_value;                 // u += dt * curr_value
        if (has_u_max && v > u_max) v = u_max;      // np.where(u > u_max, u_max, u)
        u[n] = v;
    } else {
        u[n] = value;
    }
}

__global__ void stim_box_kernel(double *u, const uint8_t *tissue, StimBox b, int mode,
                                double value, double dt_value, int has_u_max, double u_max)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= b.count) return;
    const int64_t l = q % b.e2, r = (q / b.e2) % b.e1, p = q / (b.e2 * b.e1);
    const int64_t n = (b.o0 + p) * b.s0 + (b.o1 + r) * b.s1 + (b.o2 + l);
    if (tissue[n] != 1) return;
    stim_apply(u, n, mode, value, dt_value, has_u_max, u_max);
}

__global__ void stim_nodes_kernel(double *u, const int64_t *nodes, int64_t count, int mode,
                                  double value, double dt_value, int has_u_max, double u_max)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    stim_apply(u, nodes[q], mode, value, dt_value, has_u_max, u_max);
}

// ---------------------------------------------------------------------------
// tracker helpers
// ---------------------------------------------------------------------------
// extra ActivationTime trackers beyond the fused one
__global__ void act_kernel(double *act_t, const double *u, int64_t n_nodes, double thr, double t)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    if (act_t[n] < 0 && u[n] > thr) act_t[n] = t;
}

// deterministic sum of the per-block ECG partials: one block per lead
__global__ void __launch_bounds__(256) ecg_finalize_kernel(const double *partial, int64_t n_blocks,
                                                           int n_leads, double *out)
{
    __shared__ double sh[256];
    const int lead = blockIdx.x;
    double v = 0.0;
    for (int64_t b = threadIdx.x; b < n_blocks; b += 256) v += partial[b * n_leads + lead];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[lead] = sh[0];
}

// point samplers: item = (var, flat node, compact index)
__global__ void point_gather_kernel(const int64_t *items, const double *fill, int n_items,
                                    const double *u, const double *state, int64_t ld, double *out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_items) return;
    const int64_t var = items[3 * q], n = items[3 * q + 1], c = items[3 * q + 2];
    double v;
    if (var == 0) v = u[n];
    else v = (c >= 0) ? state[(var - 1) * ld + c] : fill[q];
    out[q] = v;
}

static inline unsigned warp_blocks(int64_t n_chunks, int threads)
{
    return (unsigned)((n_chunks * 32 + threads - 1) / threads);
}

// CUDA loads a kernel lazily at its first launch, and that load can wait for the device to go
// idle.  In a slab run a kernel of one slab may be spinning on a flag that only a later launch
// of its neighbour raises -- if that launch is a kernel's FIRST (two slabs in one process share
// the context), the load never returns.  Everything the step loop can launch is therefore
// loaded up front when a halo is wired (cudaFuncGetAttributes forces the load).
template <class K> static int preload_one(K kern)
{
    cudaFuncAttributes a;
    FWB_CUDA(cudaFuncGetAttributes(&a, kern));
    return 0;
}
int preload_aux_kernels()
{
    int rc;
    if ((rc = preload_one(stim_box_kernel))) return rc;
    if ((rc = preload_one(stim_nodes_kernel))) return rc;
    if ((rc = preload_one(act_kernel))) return rc;
    if ((rc = preload_one(ecg_finalize_kernel))) return rc;
    if ((rc = preload_one(point_gather_kernel))) return rc;
    if ((rc = preload_one(halo_wait_kernel))) return rc;
    return 0;
}

int launch_stim_box(double *u, const uint8_t *tissue, const StimBox &b, int mode, double value,
                    double dt_value, int has_u_max, double u_max, cudaStream_t s)
{
    if (b.count <= 0) return 0;
    stim_box_kernel<<<(unsigned)((b.count + 255) / 256), 256, 0, s>>>(u, tissue, b, mode, value,
                                                                        dt_value, has_u_max, u_max);
    FWB_KERNEL_CHECK("stim_box_kernel");
    return 0;
}

int launch_stim_nodes(double *u, const int64_t *nodes, int64_t count, int mode, double value,
                      double dt_value, int has_u_max, double u_max, cudaStream_t s)
{
    if (count <= 0) return 0;
    stim_nodes_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(u, nodes, count, mode, value,
                                                                        dt_value, has_u_max, u_max);
    FWB_KERNEL_CHECK("stim_nodes_kernel");
    return 0;
}

int launch_act(double *act_t, const double *u, int64_t n_nodes, double thr, double t,
               cudaStream_t s)
{
    if (n_nodes <= 0) return 0;
    act_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s>>>(act_t, u, n_nodes, thr, t);
    FWB_KERNEL_CHECK("act_kernel");
    return 0;
}

int launch_ecg_finalize(const double *partial, int64_t n_blocks, int n_leads, double *out,
                        cudaStream_t s)
{
    if (n_leads <= 0) return 0;
    ecg_finalize_kernel<<<n_leads, 256, 0, s>>>(partial, n_blocks, n_leads, out);
    FWB_KERNEL_CHECK("ecg_finalize_kernel");
    return 0;
}

int launch_point_gather(const int64_t *items, const double *fill, int n_items, const double *u,
                        const double *state, int64_t ld, double *out, cudaStream_t s)
{
    if (n_items <= 0) return 0;
    point_gather_kernel<<<(n_items + 63) / 64, 64, 0, s>>>(items, fill, n_items, u, state, ld, out);
    FWB_KERNEL_CHECK("point_gather_kernel");
    return 0;
}

int launch_halo_wait(const unsigned *flag_lo, const unsigned *flag_hi, unsigned epoch,
                     cudaStream_t s)
{
    halo_wait_kernel<<<1, 32, 0, s>>>(flag_lo, flag_hi, epoch);
    FWB_KERNEL_CHECK("halo_wait_kernel");
    return 0;
}

// exclusive scan of popcount(bits[i]) into base[i]; total to the host (synchronises)
static int scan_popc(const uint32_t *bits, int64_t n, uint32_t *base, unsigned long long *total,
                     cudaStream_t s)
{
    const int n_sb = (int)((n + SCAN_ITEMS - 1) / SCAN_ITEMS);
    unsigned long long *tot = nullptr;
    FWB_CUDA(cudaMallocAsync((void **)&tot, sizeof(unsigned long long) * (n_sb + 1), s));
    scan_totals_kernel<<<n_sb, 1024, 0, s>>>(bits, n, tot);
    FWB_KERNEL_CHECK("scan_totals_kernel");
    scan_blocks_kernel<<<1, 32, 0, s>>>(tot, n_sb, tot + n_sb);
    FWB_KERNEL_CHECK("scan_blocks_kernel");
    scan_final_kernel<<<n_sb, 1024, 0, s>>>(bits, n, tot, base);
    FWB_KERNEL_CHECK("scan_final_kernel");
    FWB_CUDA(cudaMemcpyAsync(total, tot + n_sb, sizeof(*total), cudaMemcpyDeviceToHost, s));
    FWB_CUDA(cudaStreamSynchronize(s));
    FWB_CUDA(cudaFreeAsync(tot, s));
    return 0;
}

}  // namespace fwb

using namespace fwb;

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

extern "C" int64_t fwb_worklist_capacity(int dim, const int64_t *shape)
{
    if (!shape || (dim != 2 && dim != 3)) return FWB_E_ARG;
    int64_t n = 1;
    for (int d = 0; d < dim; ++d) n *= shape[d];
    // every chunk once + padding of each of the three segments and of partial tiles
    const int64_t line = shape[dim - 1];
    const int64_t S = shape[0], R = dim == 3 ? shape[1] : 1;
    if (line % 32 != 0) return cdiv(n, 32) + 8;
    const int64_t cpl = line / 32;
    return (cdiv(S, 2) * 2 + 4) * (cdiv(R, 8) * 8) * (cdiv(cpl, 8) * 8) + 64;
}

extern "C" int fwb_build_worklist(int dim, const int64_t *shape, const uint8_t *active,
                                  int halo_lo, int halo_hi, int32_t *worklist,
                                  int64_t capacity, int64_t *n_work, int64_t *n_lo_blocks,
                                  int64_t *n_hi_blocks, fwb_stream_t stream)
{
    if (!shape || (dim != 2 && dim != 3) || !active || !worklist || !n_work) {
        set_error("fwb_build_worklist: bad argument");
        return FWB_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t n_nodes = 1;
    for (int d = 0; d < dim; ++d) n_nodes *= shape[d];
    const int64_t line = shape[dim - 1];
    const int64_t S = shape[0], R = dim == 3 ? shape[1] : 1;
    const bool tiled = line % 32 == 0;
    if ((halo_lo || halo_hi) && (!tiled || S < 4)) {
        set_error("fwb_build_worklist: a slab halo needs a contiguous axis that is a multiple "
                  "of 32 nodes and at least 4 slices");
        return FWB_E_UNSUPPORTED;
    }
    FWB_CUDA(cudaMemsetAsync(worklist, 0xff, sizeof(int32_t) * capacity, s));

    WlSeg segs[3];
    int n_segs = 0;
    auto make = [&](int64_t s0, int64_t s1, bool boundary) {
        WlSeg g;
        memset(&g, 0, sizeof(g));
        g.s0 = s0; g.s1 = s1; g.R = R; g.cpl = tiled ? line / 32 : 0;
        g.n_chunks = cdiv(n_nodes, 32); g.n_nodes = n_nodes;
        if (!tiled) { g.linear = 1; g.n_pos = g.n_chunks; return g; }
        if (dim == 3) { if (boundary) { g.ts = 1; g.tr = 8; g.tc = 1; } else { g.ts = 2; g.tr = 4; g.tc = 1; } }
        else { if (boundary) { g.ts = 1; g.tr = 1; g.tc = 8; } else { g.ts = 8; g.tr = 1; g.tc = 1; } }
        g.nTc = cdiv(g.cpl, g.tc); g.nTr = cdiv(R, g.tr);
        g.n_pos = cdiv(s1 - s0, g.ts) * g.nTr * g.nTc * 8;
        return g;
    };
    int64_t lo0 = 0, hi1 = S;
    if (halo_lo) { segs[n_segs++] = make(1, 2, true); lo0 = 2; }
    if (halo_hi) { segs[n_segs++] = make(S - 2, S - 1, true); hi1 = S - 2; }
    // ghost slices (0 / S-1 on a halo side) are never listed
    segs[n_segs++] = make(halo_lo ? lo0 : 0, halo_hi ? hi1 : S, false);

    int64_t off = 0, blocks[3] = {0, 0, 0};
    for (int i = 0; i < n_segs; ++i) {
        const WlSeg &g = segs[i];
        if (g.n_pos <= 0) continue;
        int32_t *cand = nullptr;
        uint32_t *flag = nullptr, *base = nullptr;
        FWB_CUDA(cudaMallocAsync((void **)&cand, sizeof(int32_t) * g.n_pos, s));
        FWB_CUDA(cudaMallocAsync((void **)&flag, sizeof(uint32_t) * g.n_pos, s));
        FWB_CUDA(cudaMallocAsync((void **)&base, sizeof(uint32_t) * g.n_pos, s));
        const unsigned nb = (unsigned)cdiv(g.n_pos, 256);
        worklist_mark_kernel<<<nb, 256, 0, s>>>(g, active, cand, flag);
        FWB_KERNEL_CHECK("worklist_mark_kernel");
        worklist_tileflag_kernel<<<nb, 256, 0, s>>>(g.n_pos, flag);
        FWB_KERNEL_CHECK("worklist_tileflag_kernel");
        unsigned long long total = 0;
        int rc = scan_popc(flag, g.n_pos, base, &total, s);
        if (rc) return rc;
        if (off + (int64_t)total > capacity) {
            set_error("fwb_build_worklist: capacity %lld too small", (long long)capacity);
            return FWB_E_ARG;
        }
        worklist_scatter_kernel<<<nb, 256, 0, s>>>(g.n_pos, cand, flag, base, worklist + off);
        FWB_KERNEL_CHECK("worklist_scatter_kernel");
        FWB_CUDA(cudaFreeAsync(cand, s));
        FWB_CUDA(cudaFreeAsync(flag, s));
        FWB_CUDA(cudaFreeAsync(base, s));
        blocks[i] = cdiv((int64_t)total, WARPS_PER_BLOCK);
        off += blocks[i] * WARPS_PER_BLOCK;      // segments start on a block boundary
    }
    FWB_CUDA(cudaStreamSynchronize(s));
    *n_work = off;
    int bi = 0;
    if (n_lo_blocks) *n_lo_blocks = halo_lo ? blocks[bi] : 0;
    if (halo_lo) ++bi;
    if (n_hi_blocks) *n_hi_blocks = halo_hi ? blocks[bi] : 0;
    return 0;
}

extern "C" int fwb_order_compact(const uint32_t *chunk_bits, int64_t n_chunks,
                                 const int32_t *worklist, int64_t n_work, uint32_t *chunk_base,
                                 uint32_t *tile_base, uint32_t *records, int64_t *n_myo,
                                 fwb_stream_t stream)
{
    if (!chunk_bits || !worklist || !chunk_base || !tile_base || n_work < 0 || n_work % 8 != 0) {
        set_error("fwb_order_compact: bad argument");
        return FWB_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    FWB_CUDA(cudaMemsetAsync(chunk_base, 0, sizeof(uint32_t) * n_chunks, s));
    unsigned long long total = 0;
    if (n_work > 0) {
        uint32_t *tmp = nullptr, *ord = nullptr;
        FWB_CUDA(cudaMallocAsync((void **)&tmp, sizeof(uint32_t) * n_work, s));
        FWB_CUDA(cudaMallocAsync((void **)&ord, sizeof(uint32_t) * n_work, s));
        const unsigned nb = (unsigned)cdiv(n_work + 1, 256);
        order_gather_kernel<<<nb, 256, 0, s>>>(chunk_bits, worklist, n_work, tmp);
        FWB_KERNEL_CHECK("order_gather_kernel");
        int rc = scan_popc(tmp, n_work, ord, &total, s);
        if (rc) return rc;
        order_scatter_kernel<<<nb, 256, 0, s>>>(worklist, n_work, ord, tmp, (uint32_t)total,
                                                chunk_base, tile_base,
                                                reinterpret_cast<uint4 *>(records));
        FWB_KERNEL_CHECK("order_scatter_kernel");
        FWB_CUDA(cudaFreeAsync(tmp, s));
        FWB_CUDA(cudaFreeAsync(ord, s));
    } else {
        FWB_CUDA(cudaMemsetAsync(tile_base, 0, sizeof(uint32_t), s));
    }
    FWB_CUDA(cudaStreamSynchronize(s));
    if (n_myo) *n_myo = (int64_t)total;
    return 0;
}

extern "C" int fwb_order_tiles(int dim, const int64_t *shape, int halo_lo, int halo_hi,
                               int64_t n_lo_blocks, int64_t n_hi_blocks,
                               const uint32_t *chunk_bits, const uint32_t *chunk_base,
                               const int32_t *worklist, int64_t n_work, const uint32_t *tile_base,
                               uint32_t *tile_rec, uint8_t *pos_of, fwb_stream_t stream)
{
    if (!shape || (dim != 2 && dim != 3) || !chunk_bits || !chunk_base || !worklist || !tile_base ||
        !tile_rec || !pos_of || n_work < 0 || n_work % 8 != 0) {
        set_error("fwb_order_tiles: bad argument");
        return FWB_E_ARG;
    }
    if (n_work == 0) return 0;
    TileGeo tg;
    memset(&tg, 0, sizeof(tg));
    const int64_t line = shape[dim - 1];
    tg.dim = dim; tg.tiled = line % 32 == 0;
    tg.halo_lo = halo_lo; tg.halo_hi = halo_hi;
    tg.S = shape[0]; tg.R = dim == 3 ? shape[1] : 1; tg.cpl = tg.tiled ? line / 32 : 0;
    tg.n_lo_tiles = halo_lo ? n_lo_blocks : 0; tg.n_hi_tiles = halo_hi ? n_hi_blocks : 0;
    const int64_t n_tiles = n_work / 8;
    const unsigned nb = (unsigned)cdiv(n_tiles * 32, 256);
    order_tiles_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(
        tg, worklist, n_tiles, chunk_bits, chunk_base, tile_base,
        reinterpret_cast<uint4 *>(tile_rec), pos_of);
    FWB_KERNEL_CHECK("order_tiles_kernel");
    return 0;
}

// ---- IPC-exportable device memory for slab halos ---------------------------
extern "C" int fwb_dev_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) { set_error("fwb_dev_alloc: bad argument"); return FWB_E_ARG; }
    FWB_CUDA(cudaMalloc(ptr, (size_t)bytes));
    FWB_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return 0;
}
extern "C" int fwb_dev_free(void *ptr)
{
    if (ptr) FWB_CUDA(cudaFree(ptr));
    return 0;
}
extern "C" int fwb_ipc_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }
extern "C" int fwb_ipc_get_handle(const void *ptr, void *handle_out)
{
    if (!ptr || !handle_out) { set_error("fwb_ipc_get_handle: bad argument"); return FWB_E_ARG; }
    cudaIpcMemHandle_t h;
    FWB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}
extern "C" int fwb_ipc_open_handle(const void *handle, void **ptr_out)
{
    if (!handle || !ptr_out) { set_error("fwb_ipc_open_handle: bad argument"); return FWB_E_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    FWB_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int fwb_ipc_close_handle(void *ptr)
{
    if (ptr) FWB_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
// slabs of ONE process on several devices: let kernels on `device` store into / read flags in
// memory of `peer` (NVLink peer mapping); a pair that is already enabled is not an error
extern "C" int fwb_enable_peer_access(int device, int peer)
{
    if (device == peer) return 0;
    int can = 0;
    FWB_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) { set_error("device %d cannot access device %d", device, peer); return FWB_E_UNSUPPORTED; }
    int prev = 0;
    FWB_CUDA(cudaGetDevice(&prev));
    FWB_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fwb::cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return 0;
}

extern "C" int fwb_build_chunks(const uint8_t *update_mask, int64_t n_nodes,
                                uint32_t *chunk_bits, uint32_t *chunk_base,
                                int64_t *n_myo, fwb_stream_t stream)
{
    if (!update_mask || !chunk_bits || !chunk_base || n_nodes <= 0) {
        set_error("fwb_build_chunks: bad argument");
        return FWB_E_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_chunks = (n_nodes + 31) / 32;
    chunk_bits_kernel<<<warp_blocks(n_chunks, 256), 256, 0, s>>>(update_mask, n_nodes, chunk_bits,
                                                                  n_chunks);
    FWB_KERNEL_CHECK("chunk_bits_kernel");
    unsigned long long total = 0;
    {
        int rc = scan_popc(chunk_bits, n_chunks, chunk_base, &total, s);
        if (rc) return rc;
    }
    if (total > 0xffffffffULL) {
        set_error("fwb_build_chunks: more than 2^32 updated nodes on one device");
        return FWB_E_UNSUPPORTED;
    }
    if (n_myo) *n_myo = (int64_t)total;
    return 0;
}

extern "C" int fwb_gather_compact(const double *dense, double *compact, int64_t n_nodes,
                                  const uint32_t *chunk_bits, const uint32_t *chunk_base,
                                  fwb_stream_t stream)
{
    if (!dense || !compact || !chunk_bits || !chunk_base) { set_error("fwb_gather_compact: bad argument"); return FWB_E_ARG; }
    const int64_t n_chunks = (n_nodes + 31) / 32;
    gather_kernel<<<warp_blocks(n_chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        dense, compact, n_chunks, chunk_bits, chunk_base);
    FWB_KERNEL_CHECK("gather_kernel");
    return 0;
}

extern "C" int fwb_scatter_compact(const double *compact, double *dense, double fill,
                                   int64_t n_nodes, const uint32_t *chunk_bits,
                                   const uint32_t *chunk_base, fwb_stream_t stream)
{
    if (!dense || !compact || !chunk_bits || !chunk_base) { set_error("fwb_scatter_compact: bad argument"); return FWB_E_ARG; }
    const int64_t n_chunks = (n_nodes + 31) / 32;
    scatter_kernel<<<warp_blocks(n_chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        compact, dense, fill, 0, n_nodes, n_chunks, chunk_bits, chunk_base);
    FWB_KERNEL_CHECK("scatter_kernel");
    return 0;
}

extern "C" int fwb_scatter_compact_keep(const double *compact, double *dense, int64_t n_nodes,
                                        const uint32_t *chunk_bits, const uint32_t *chunk_base,
                                        fwb_stream_t stream)
{
    if (!dense || !compact || !chunk_bits || !chunk_base) { set_error("fwb_scatter_compact_keep: bad argument"); return FWB_E_ARG; }
    const int64_t n_chunks = (n_nodes + 31) / 32;
    scatter_kernel<<<warp_blocks(n_chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        compact, dense, 0.0, 1, n_nodes, n_chunks, chunk_bits, chunk_base);
    FWB_KERNEL_CHECK("scatter_kernel");
    return 0;
}

extern "C" int fwb_count_offfill(const double *dense, double fill, int64_t n_nodes,
                                 const uint32_t *chunk_bits, unsigned long long *count,
                                 fwb_stream_t stream)
{
    if (!dense || !chunk_bits || !count) { set_error("fwb_count_offfill: bad argument"); return FWB_E_ARG; }
    const int64_t n_chunks = (n_nodes + 31) / 32;
    count_offfill_kernel<<<warp_blocks(n_chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        dense, fill, n_nodes, n_chunks, chunk_bits, count);
    FWB_KERNEL_CHECK("count_offfill_kernel");
    return 0;
}

extern "C" int fwb_weights_pack(const double *dense_aos, double *compact_soa, int K, int64_t ld,
                                int64_t n_nodes, const uint32_t *chunk_bits,
                                const uint32_t *chunk_base, fwb_stream_t stream)
{
    if (!dense_aos || !compact_soa || K <= 0) { set_error("fwb_weights_pack: bad argument"); return FWB_E_ARG; }
    const int64_t n_chunks = (n_nodes + 31) / 32;
    weights_pack_kernel<<<warp_blocks(n_chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        dense_aos, compact_soa, K, ld, n_chunks, chunk_bits, chunk_base);
    FWB_KERNEL_CHECK("weights_pack_kernel");
    return 0;
}

extern "C" int fwb_weights_unpack(const double *compact_soa, double *dense_aos, int K, int64_t ld,
                                  int64_t n_nodes, const uint32_t *chunk_bits,
                                  const uint32_t *chunk_base, fwb_stream_t stream)
{
    if (!dense_aos || !compact_soa || K <= 0) { set_error("fwb_weights_unpack: bad argument"); return FWB_E_ARG; }
    const int64_t n_chunks = (n_nodes + 31) / 32;
    weights_unpack_kernel<<<warp_blocks(n_chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        compact_soa, dense_aos, K, ld, n_nodes, n_chunks, chunk_bits, chunk_base);
    FWB_KERNEL_CHECK("weights_unpack_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// LocalActivationTime / Period trackers:
// finitewave/cpuwave2D/tracker/local_activation_time_2d_tracker.py:58-90 (_track,
// cross_threshold), period_2d_tracker.py:38-52.  Element-wise over every grid node.
// ---------------------------------------------------------------------------
namespace fwb {
// cross = (u >= thr) & !activated; activated |= cross; then nodes with (u < thr) & activated
// are re-armed (cross_threshold :71-90).  `layer` (may be NULL) is the tracker's current
// activation layer: *flag is raised if a crossing node already holds a time in it (_track :66)
__global__ void lat_cross_kernel(const double *u, int64_t n, double thr, uint8_t *activated,
                                 uint8_t *cross, const double *layer, int *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (i < n) {
        const double v = u[i];
        const uint8_t a = activated[i];
        const bool c = (v >= thr) && a == 0;
        uint8_t a2 = c ? 1 : a;
        if (v < thr && a2 == 1) a2 = 0;
        activated[i] = a2;
        cross[i] = c ? 1 : 0;
        hit = c && layer && layer[i] > -1.0;
    }
    if (__any_sync(0xffffffffu, hit) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
// layer = where(cross, t, layer)   (_track :69)
__global__ void lat_write_kernel(const uint8_t *cross, int64_t n, double t, double *layer)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && cross[i]) layer[i] = t;
}
__global__ void gather_u8_kernel(const uint8_t *src, const int64_t *idx, int64_t k, uint8_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = src[idx[i]];
}
}  // namespace fwb

extern "C" int fwb_lat_cross(const double *u, int64_t n_nodes, double threshold, uint8_t *activated,
                             uint8_t *cross, const double *layer, int *flag, fwb_stream_t stream)
{
    if (!u || !activated || !cross || n_nodes < 0 || (layer && !flag)) {
        set_error("fwb_lat_cross: bad argument");
        return FWB_E_ARG;
    }
    if (n_nodes == 0) return 0;
    fwb::lat_cross_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        u, n_nodes, threshold, activated, cross, layer, flag);
    FWB_KERNEL_CHECK("lat_cross_kernel");
    return 0;
}

extern "C" int fwb_lat_write(const uint8_t *cross, int64_t n_nodes, double t, double *layer,
                             fwb_stream_t stream)
{
    if (!cross || !layer || n_nodes < 0) { set_error("fwb_lat_write: bad argument"); return FWB_E_ARG; }
    if (n_nodes == 0) return 0;
    fwb::lat_write_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        cross, n_nodes, t, layer);
    FWB_KERNEL_CHECK("lat_write_kernel");
    return 0;
}

extern "C" int fwb_gather_u8(const uint8_t *src, const int64_t *idx, int64_t k, uint8_t *out,
                             fwb_stream_t stream)
{
    if (!src || !idx || !out || k < 0) { set_error("fwb_gather_u8: bad argument"); return FWB_E_ARG; }
    if (k == 0) return 0;
    fwb::gather_u8_kernel<<<(unsigned)((k + 127) / 128), 128, 0, (cudaStream_t)stream>>>(src, idx, k, out);
    FWB_KERNEL_CHECK("gather_u8_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// Fibrosis patterns on the device: Diffuse{2,3}DPattern and
// Structural{2,3}DPattern (finitewave/cpuwave2D/fibrosis/diffuse_2d_pattern.py:63-87,
// structural_2d_pattern.py:71-120, cpuwave3D/fibrosis/*.py).  Inside the box every node
// (diffuse) or every block anchored at the box origin and clipped at its end (structural)
// becomes 2 with probability `density` and 1 otherwise -- the box is overwritten, as in the
// reference.  The reference draws from numpy's / random's global state; here the draw is a
// hash of (seed, global block id), so a tissue is reproducible and identical however it is
// cut into slabs.
// ---------------------------------------------------------------------------
namespace fwb {
struct PatternArgs {
    int dim;
    int64_t shape[3];     // stored (local) shape, padded with 1
    int64_t lo[3], hi[3]; // box in GLOBAL indices, half open
    int64_t len[3];       // block edge lengths (1 = diffuse)
    int64_t nb[3];        // blocks per axis
    int64_t off[3];       // global index of local index 0 per axis (slab runs: slowest real axis)
    double density;
    unsigned long long seed;
};
__host__ __device__ inline double hash_u01(unsigned long long id, unsigned long long seed)
{
    unsigned long long h = id + seed * 0x9E3779B97F4A7C15ULL;     // splitmix64 finaliser
    h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ULL;
    h ^= h >> 27; h *= 0x94D049BB133111EBULL;
    h ^= h >> 31;
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void pattern_kernel(int8_t *mesh, const __grid_constant__ PatternArgs A)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = A.shape[0] * A.shape[1] * A.shape[2];
    if (n >= total) return;
    int64_t g[3];
    g[2] = n % A.shape[2] + A.off[2];
    g[1] = (n / A.shape[2]) % A.shape[1] + A.off[1];
    g[0] = n / (A.shape[2] * A.shape[1]) + A.off[0];
    int64_t b[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (g[a] < A.lo[a] || g[a] >= A.hi[a]) return;
        b[a] = (g[a] - A.lo[a]) / A.len[a];
    }
    const unsigned long long id = (unsigned long long)((b[0] * A.nb[1] + b[1]) * A.nb[2] + b[2]);
    mesh[n] = hash_u01(id, A.seed) <= A.density ? 2 : 1;
}
}  // namespace fwb

extern "C" int fwb_pattern_fibrosis(int8_t *mesh, int dim, const int64_t *shape, const int64_t *box,
                                    const int64_t *block, double density, unsigned long long seed,
                                    int64_t slow_offset, fwb_stream_t stream)
{
    if (!mesh || (dim != 2 && dim != 3) || !shape || !box || !block) {
        set_error("fwb_pattern_fibrosis: bad argument");
        return FWB_E_ARG;
    }
    fwb::PatternArgs A;
    const int pad = 3 - dim;          // 2D grids are (1, n_i, n_j)
    for (int a = 0; a < 3; ++a) { A.shape[a] = 1; A.lo[a] = 0; A.hi[a] = 1; A.len[a] = 1; }
    for (int a = 0; a < dim; ++a) {
        A.shape[a + pad] = shape[a];
        A.lo[a + pad] = box[2 * a];
        A.hi[a + pad] = box[2 * a + 1];
        A.len[a + pad] = block[a];
        if (block[a] < 1 || shape[a] < 0) { set_error("fwb_pattern_fibrosis: bad block / shape"); return FWB_E_ARG; }
    }
    for (int a = 0; a < 3; ++a) {
        const int64_t ext = A.hi[a] > A.lo[a] ? A.hi[a] - A.lo[a] : 0;
        A.nb[a] = (ext + A.len[a] - 1) / A.len[a];
        if (A.nb[a] < 1) A.nb[a] = 1;
    }
    A.dim = dim; A.density = density; A.seed = seed;
    A.off[0] = A.off[1] = A.off[2] = 0;
    A.off[pad] = slow_offset;         // slabs are cut along the slowest real axis
    const int64_t total = A.shape[0] * A.shape[1] * A.shape[2];
    if (total == 0) return 0;
    fwb::pattern_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mesh, A);
    FWB_KERNEL_CHECK("pattern_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// SpiralWaveCore{2,3}DTracker:
// finitewave/cpuwave2D/tracker/spiral_wave_core_2d_tracker.py:108-248 (_correct_tip_pos,
// _apply_threshold, _track_tip_line), cpuwave3D/tracker/spiral_wave_core_3d_tracker.py:30-48
// (one scan per slice of the LAST axis).  One thread per 2x2 cell; the arithmetic is the
// reference's, operation by operation (the file is compiled without FMA contraction), so
// the tip coordinates are bit-identical.  Output rows {x, y, cell key}; the key
// ((k * n_i + i) * n_j + j) restores the reference's scan order on the host.
// ---------------------------------------------------------------------------
namespace fwb {
struct TipArgs {
    const double *u_prev, *u;
    int64_t n_i, n_j, n_k;      // n_k = 1 in 2D
    int64_t s_i, s_j;           // flat strides of axis i and j (2D: n_j, 1; 3D: n_j*n_k, n_k)
    double thr;
    double *out;                // [capacity][3]
    unsigned *count;
    unsigned capacity;
};
__device__ __forceinline__ int tip_threshold(const double *p, int64_t s_i, int64_t s_j, double thr)
{
    const double a = p[0], b = p[s_i], c = p[s_j], d = p[s_i + s_j];
    if (a >= thr && (b < thr || c < thr || d < thr)) return 1;
    if (a < thr && (b >= thr || c >= thr || d >= thr)) return 1;
    return 0;
}
__global__ void tip_scan_kernel(const __grid_constant__ TipArgs A)
{
    constexpr int64_t delta = 5;                    // safety margin of _track_tip_line
    const int64_t wi = A.n_i - 2 * delta, wj = A.n_j - 2 * delta;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi <= 0 || wj <= 0 || q >= wi * wj * A.n_k) return;
    const int64_t j = q % wj + delta, i = (q / wj) % wi + delta, k = q / (wj * wi);
    const int64_t n = i * A.s_i + j * A.s_j + k;    // the last axis has stride 1 in 3D;
                                                    // in 2D n_k = 1, k = 0, s_j = 1
    const double *u = A.u_prev + n, *v = A.u + n;
    const int64_t si = A.s_i, sj = A.s_j;
    if (tip_threshold(u, si, sj, A.thr) != 1 || tip_threshold(v, si, sj, A.thr) != 1) return;
    const double thr = A.thr;
    // _correct_tip_pos :124-170
    const double AC = add(add(add(u[0], -u[sj]), u[si + sj]), -u[si]);
    const double GC = add(u[sj], -u[0]);
    const double BC = add(u[si], -u[0]);
    const double DC = add(u[0], -thr);
    const double AD = add(add(add(v[0], -v[sj]), v[si + sj]), -v[si]);
    const double GD = add(v[sj], -v[0]);
    const double BD = add(v[si], -v[0]);
    const double DD = add(v[0], -thr);
    const double Q = add(mul(BC, AD), -mul(BD, AC));
    const double R = add(mul(GC, AD), -mul(GD, AC));
    const double S = add(mul(DC, AD), -mul(DD, AC));
    const double QOnR = __ddiv_rn(Q, R), SOnR = __ddiv_rn(S, R);
    const double T = mul(AC, QOnR);
    const double U = add(add(mul(AC, SOnR), -BC), mul(GC, QOnR));
    const double V = add(mul(GC, SOnR), -DC);
    const double disc = add(mul(U, U), -mul(mul(4., T), V));
    if (disc < 0) return;
    const double T2 = mul(2., T);
    if (T2 == 0.) return;
    const double sq = __dsqrt_rn(disc);
    const double xn = __ddiv_rn(add(-U, -sq), T2), xp = __ddiv_rn(add(-U, sq), T2);
    const double yn = add(mul(-QOnR, xn), -SOnR), yp = add(mul(-QOnR, xp), -SOnR);
    double cx, cy;
    if (0 <= xn && xn <= 1 && 0 <= yn && yn <= 1) { cx = xn; cy = yn; }
    else if (0 <= xp && xp <= 1 && 0 <= yp && yp <= 1) { cx = xp; cy = yp; }
    else return;
    const unsigned slot = atomicAdd(A.count, 1u);
    if (slot < A.capacity) {
        A.out[3 * slot] = add((double)j, cy);        // out.append([j + correction[1], i + correction[0]])
        A.out[3 * slot + 1] = add((double)i, cx);
        A.out[3 * slot + 2] = (double)((k * A.n_i + i) * A.n_j + j);
    }
}
}  // namespace fwb

extern "C" int fwb_tip_scan(const double *u_prev, const double *u, int dim, const int64_t *shape,
                            double threshold, double *out, unsigned capacity, unsigned *count,
                            fwb_stream_t stream)
{
    if (!u_prev || !u || (dim != 2 && dim != 3) || !shape || !out || !count) {
        set_error("fwb_tip_scan: bad argument");
        return FWB_E_ARG;
    }
    fwb::TipArgs A;
    A.u_prev = u_prev; A.u = u; A.thr = threshold; A.out = out; A.count = count; A.capacity = capacity;
    A.n_i = shape[0]; A.n_j = shape[1]; A.n_k = dim == 3 ? shape[2] : 1;
    A.s_i = A.n_j * A.n_k; A.s_j = A.n_k;
    const int64_t wi = A.n_i - 10, wj = A.n_j - 10;
    if (wi <= 0 || wj <= 0 || A.n_k <= 0) return 0;
    const int64_t total = wi * wj * A.n_k;
    fwb::tip_scan_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    FWB_KERNEL_CHECK("tip_scan_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// Device evaluation of the fast-path math helpers (fexp.cuh) for the accuracy tests: the
// host builds of the same sources are checked against libm on the CPU; this is the check of
// what the GPU actually computes (MUFU seeds, FMA contraction).
// ---------------------------------------------------------------------------
namespace fwb {
__global__ void devmath_kernel(int op, const double *x, double *y, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r;
    switch (op) {
    case 0: r = fexp(v); break;
    case 1: r = fexp_fast(v); break;
    case 2: r = flog(v); break;
    case 3: r = frcp(v); break;
    case 4: r = frcp3(v); break;
    case 5: r = fsqrt(v); break;
    default: r = v;
    }
    y[i] = r;
}
}  // namespace fwb

extern "C" int fwb_devmath(int op, const double *x, double *y, int64_t n, fwb_stream_t stream)
{
    if (!x || !y || n < 0 || op < 0 || op > 5) { set_error("fwb_devmath: bad argument"); return FWB_E_ARG; }
    if (n == 0) return 0;
    fwb::devmath_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(op, x, y, n);
    FWB_KERNEL_CHECK("devmath_kernel");
    return 0;
}

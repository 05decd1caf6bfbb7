// fwb_common.cuh -- shared definitions of the sm_100a step backend.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/finitewave_b200.h"

namespace fwb {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK_THREADS = WARPS_PER_BLOCK * 32;
constexpr int MAX_PARAMS = 64;
constexpr int MAX_STATE = 19;

// thread-local last error (fwb_last_error)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define FWB_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return fwb::cuda_fail(e__, #call); \
    } while (0)

#define FWB_KERNEL_CHECK(what)                           \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) return fwb::cuda_fail(e__, what); \
    } while (0)

// ---------------------------------------------------------------------------
// Grid description shared by all kernels.
//
// The dense grid is C order.  Internally every grid is (planes, rows, line):
//   3D: planes = n_i, rows = n_j, line = n_k      2D: planes = 1, rows = n_i, line = n_j
// `line` is the contiguous axis.  A chunk is 32 consecutive FLAT nodes; a warp
// owns one chunk, lane l owns node 32*chunk + l.
//
// Block -> chunk mapping (tile_* > 0, used when line % 32 == 0): a block of
// WARPS_PER_BLOCK warps covers tile_p planes x tile_r rows of one 32-node
// column segment, so that the stencil's neighbour lines are shared through L1.
// Otherwise (cpl == 0) blocks take WARPS_PER_BLOCK consecutive chunks.
// ---------------------------------------------------------------------------
struct Grid {
    int dim;
    int64_t n_nodes;
    int64_t n_chunks;
    int64_t s_plane;   // flat stride of axis i (3D) ; 0 in 2D
    int64_t s_row;     // flat stride of the row axis (j in 3D, i in 2D) = line length
    int planes, rows, line;
    int cpl;           // chunks per line (line / 32) or 0 for the linear mapping
    int tile_p, tile_r;
    int tiles_r;       // ceil(rows / tile_r)
    int tiles_p;       // ceil(planes / tile_p)
    int64_t n_blocks;
    const uint32_t *chunk_bits;
    const uint32_t *chunk_base;
    int64_t ld;
};

inline Grid make_grid(int dim, const int64_t *shape, const uint32_t *bits,
                      const uint32_t *base, int64_t ld)
{
    Grid g;
    memset(&g, 0, sizeof(g));
    g.dim = dim;
    if (dim == 3) {
        g.planes = (int)shape[0]; g.rows = (int)shape[1]; g.line = (int)shape[2];
        g.s_plane = shape[1] * shape[2];
        g.s_row = shape[2];
    } else {
        g.planes = 1; g.rows = (int)shape[0]; g.line = (int)shape[1];
        g.s_plane = 0;
        g.s_row = shape[1];
    }
    g.n_nodes = (int64_t)g.planes * g.rows * g.line;
    g.n_chunks = (g.n_nodes + 31) / 32;
    g.chunk_bits = bits;
    g.chunk_base = base;
    g.ld = ld;
    if (g.line % 32 == 0) {
        g.cpl = g.line / 32;
        if (dim == 3 && g.planes >= 2) { g.tile_p = 2; g.tile_r = WARPS_PER_BLOCK / 2; }
        else { g.tile_p = 1; g.tile_r = WARPS_PER_BLOCK; }
        g.tiles_r = (g.rows + g.tile_r - 1) / g.tile_r;
        g.tiles_p = (g.planes + g.tile_p - 1) / g.tile_p;
        g.n_blocks = (int64_t)g.cpl * g.tiles_r * g.tiles_p;
    } else {
        g.cpl = 0;
        g.n_blocks = (g.n_chunks + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    }
    return g;
}

#ifdef __CUDACC__
// chunk owned by this warp, or -1
__device__ __forceinline__ int64_t warp_chunk(const Grid &g)
{
    const int warp = threadIdx.x >> 5;
    if (g.cpl == 0) {
        const int64_t c = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
        return c < g.n_chunks ? c : -1;
    }
    const unsigned b = blockIdx.x;
    const unsigned seg = b % (unsigned)g.cpl;
    const unsigned b2 = b / (unsigned)g.cpl;
    const unsigned tr = b2 % (unsigned)g.tiles_r;
    const unsigned tp = b2 / (unsigned)g.tiles_r;
    const int r = (int)tr * g.tile_r + warp % g.tile_r;
    const int p = (int)tp * g.tile_p + warp / g.tile_r;
    if (r >= g.rows || p >= g.planes) return -1;
    return ((int64_t)p * g.rows + r) * g.cpl + seg;
}

// exact IEEE multiply/add that the compiler may never contract into an FMA
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

// streaming loads/stores for data touched once per step
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
#endif

}  // namespace fwb

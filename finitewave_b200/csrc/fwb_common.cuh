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
constexpr int MAX_STATE = 21;

// thread-local last error (fwb_last_error)
void set_error(const char *fmt, ...);
// which step kernel the last launch used (fwb_last_step_variant): 0 = one block per tile,
// plain loads; 3 = persistent TMA ring; 4 = compact-lane tile kernel, plain u loads; 5 = the
// same with the u brick by tensor TMA; 6 = multi-step cluster kernel; 7 = ring + u brick
void note_step_variant(int v);
int last_step_launches();
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
// Which chunk a warp takes comes from the WORK LIST (fwb_build_worklist): the ids
// of all chunks that contain tissue, in tile order (a block's 8 consecutive
// entries are 2 planes x 4 rows, or 8 rows in 2D, of one 32-node column segment,
// so the stencil's neighbour lines are shared through L1), padded with -1.  The
// list removes all index arithmetic from the step kernel, skips empty space
// (ventricle-shaped meshes) at block granularity, and fixes the block order:
// with a halo the chunks of the two slab-boundary planes come first.
// ---------------------------------------------------------------------------
struct Grid {
    int dim;
    int planes, rows, line;
    int64_t n_nodes;
    int64_t n_chunks;
    int64_t s_plane;   // flat stride of axis i (3D) ; 0 in 2D
    int64_t s_row;     // flat stride of the row axis (j in 3D, i in 2D) = line length
    int64_t slow_offset;   // global index of local slice 0 of the slowest axis (slab runs)
    const uint32_t *chunk_bits;
    const uint32_t *chunk_base;
    int64_t ld;
    const int32_t *worklist;   // [n_work], -1 = padding
    int64_t n_work;
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
    return g;
}

// nodes per slice of the slowest axis (the slab / halo unit): a plane in 3D, a row in 2D
inline int64_t slow_stride(const Grid &g) { return g.dim == 3 ? g.s_plane : g.s_row; }
inline int64_t slow_extent(const Grid &g) { return g.dim == 3 ? g.planes : g.rows; }

#ifdef __CUDACC__
// exact IEEE multiply/add that the compiler may never contract into an FMA
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

// streaming loads/stores for data touched once per step
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
#endif

}  // namespace fwb

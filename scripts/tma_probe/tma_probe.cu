// stand-alone probe: which tensor-map box shapes / data types does UTMALDG accept for a dense
// fp64 grid?  (run on the GPU box; prints one line per configuration)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap_p, const CUtensorMap *tmap_g, int c0, int c1, int c2, int bytes, double *out, int n_out, int fence_kind)
{
    const CUtensorMap *tmapp = tmap_g ? tmap_g : &tmap_p;
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        if (fence_kind) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(sm)), "l"(tmapp), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sm)), "l"(tmapp), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const double *d = reinterpret_cast<const double *>(sm);
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = d[i];
}

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const int P = 16, R = 64, L = 128;
    double *u, *out;
    cudaMalloc(&u, sizeof(double) * P * R * L);
    cudaMalloc(&out, sizeof(double) * 4096);
    double *h = new double[P * R * L];
    for (int i = 0; i < P * R * L; ++i) h[i] = i;
    cudaMemcpy(u, h, sizeof(double) * P * R * L, cudaMemcpyHostToDevice);
    struct Cfg { const char *name; int rank; CUtensorMapDataType dt; int esz; int b0, b1, b2; };
    Cfg cfgs[] = {
        {"f64 3d 34x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 34, 6, 4},
        {"f64 3d 32x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 32, 6, 4},
        {"f64 3d 36x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 36, 6, 4},
        {"f64 2d 34x10", 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 34, 10, 1},
        {"u64 3d 34x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 34, 6, 4},
        {"i32 3d 68x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 68, 6, 4},
        {"f32 3d 68x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 68, 6, 4},
        {"f64 3d 32x4x2", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 32, 4, 2},
        {"f64 3d 16x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 16, 6, 4},
        {"f64 2d 32x8", 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 32, 8, 1},
        {"i32 3d 64x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 64, 6, 4},
        {"u8 3d 256x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 6, 4},
        {"i32 3d 72x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 72, 6, 4},
        {"f64 3d 40x6x4", 3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 40, 6, 4},
    };
    for (const Cfg &c : cfgs) {
        ++idx;
        if (only >= 0 && idx != only) continue;
        CUtensorMap map;
        const int per = 8 / c.esz;
        cuuint64_t dims[3] = {(cuuint64_t)L * per, (cuuint64_t)R, (cuuint64_t)P};
        cuuint64_t strides[2] = {(cuuint64_t)L * 8, (cuuint64_t)L * R * 8};
        cuuint32_t box[3] = {(cuuint32_t)c.b0, (cuuint32_t)c.b1, (cuuint32_t)c.b2};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&map, c.dt, c.rank, u, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, (getenv("L2P") ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-16s encode failed %d\n", c.name, (int)r); continue; }
        { const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&map); printf("desc:"); for (int i = 0; i < 16; ++i) printf(" %016llx", w[i]); printf("\n  u=%p q=%d fn=%p\n", (void*)u, (int)q, fn); }
        CUtensorMap *gmap = nullptr;
        if (argc > 2) { cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice); printf("(descriptor in global memory) "); }
        const int bytes = c.b0 * c.b1 * c.b2 * c.esz;
        const int n_out = bytes / 8;
        for (int neg = 0; neg < 2; ++neg) {
            const int c0 = neg ? -per : 31 * per, c1 = neg ? -1 : 3, c2 = neg ? -1 : 1;
            cudaMemset(out, 0, sizeof(double) * 4096);
            if (c.rank == 3) probe<3><<<1, 128, 16384>>>(map, gmap, c0, c1, c2, bytes, out, n_out, getenv("FENCE") ? 1 : 0);
            else probe<2><<<1, 128, 16384>>>(map, gmap, c0, c1, 0, bytes, out, n_out, getenv("FENCE") ? 1 : 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%-16s coords(%d,%d,%d): %s\n", c.name, c0, c1, c2, cudaGetErrorString(e)); return 1; }
            double ho[4096];
            cudaMemcpy(ho, out, sizeof(double) * n_out, cudaMemcpyDeviceToHost);
            // expected element (0,0,0) of the box
            const int x0 = c0 / per, y0 = c1, z0 = c.rank == 3 ? c2 : 0;
            int bad = 0;
            const int bx = c.b0 / per;
            for (int z = 0; z < c.b2; ++z) for (int y = 0; y < c.b1; ++y) for (int x = 0; x < bx; ++x) {
                const int gx = x0 + x, gy = y0 + y, gz = z0 + z;
                const bool in = gx >= 0 && gx < L && gy >= 0 && gy < R && gz >= 0 && gz < P;
                const double want = in ? (double)((gz * R + gy) * L + gx) : 0.0;
                if (ho[(z * c.b1 + y) * bx + x] != want) ++bad;
            }
            printf("%-16s coords(%d,%d,%d): ok, %d mismatches of %d\n", c.name, c0, c1, c2, bad, n_out);
        }
    }
    return 0;
}

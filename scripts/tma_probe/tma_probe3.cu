// parametric TMA probe: tma_probe3 <dtype i|d> <rank> <b0> <b1> <b2> <smem s|y>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK, bool DYN>
__global__ void kern(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, int bytes, unsigned char *out)
{
    __shared__ alignas(128) unsigned char sbuf[DYN ? 16 : 16384];
    extern __shared__ __align__(128) unsigned char dbuf[];
    unsigned char *buf = DYN ? dbuf : sbuf;
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(buf)), "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(buf)), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char **argv)
{
    const char dt = argv[1][0]; const int rank = atoi(argv[2]);
    const int b0 = atoi(argv[3]), b1 = atoi(argv[4]), b2 = atoi(argv[5]); const bool dyn = argv[6][0] == 'y';
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const int esz = dt == 'i' ? 4 : 8;
    const int P = 16, R = 64, Lb = 1024;          // bytes per line
    unsigned char *d, *out; cudaMalloc(&d, (size_t)P * R * Lb); cudaMalloc(&out, 16384);
    cudaMemset(d, 1, (size_t)P * R * Lb);
    CUtensorMap map;
    cuuint64_t size[3] = {(cuuint64_t)Lb / esz, R, P}; cuuint64_t stride[2] = {(cuuint64_t)Lb, (cuuint64_t)Lb * R};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&map, dt == 'i' ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, d, size, stride, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = b0 * b1 * (rank == 3 ? b2 : 1) * esz;
    if (rank == 2) { if (dyn) kern<2, true><<<1, 128, 16384>>>(map, 8, 3, 1, bytes, out); else kern<2, false><<<1, 128>>>(map, 8, 3, 1, bytes, out); }
    else { if (dyn) kern<3, true><<<1, 128, 16384>>>(map, 8, 3, 1, bytes, out); else kern<3, false><<<1, 128>>>(map, 8, 3, 1, bytes, out); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("dtype %c rank %d box %dx%dx%d %s smem, %d bytes: encode %d, %s\n", dt, rank, b0, b1, b2, dyn ? "dynamic" : "static", bytes, (int)r, cudaGetErrorString(e));
    return 0;
}

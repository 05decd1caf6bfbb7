// the CUDA programming guide's own TMA example (libcu++ API), 2-D int tensor
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdio.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GH = 256, GW = 256, SH = 16, SW = 64;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void kernel_c(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SH][SW];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((int)sizeof(smem_buffer)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem_buffer)), "l"(&tensor_map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) { out[0] = smem_buffer[0][0]; out[1] = smem_buffer[SH - 1][SW - 1]; }
}
__global__ void kernel_b(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem_buffer)), "l"(&tensor_map), "r"(x), "r"(y), "r"(smem_u32(cuda::device::barrier_native_handle(bar))) : "memory");
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    if (threadIdx.x == 0) { out[0] = smem_buffer[0][0]; out[1] = smem_buffer[SH - 1][SW - 1]; }
}
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    if (threadIdx.x == 0) { out[0] = smem_buffer[0][0]; out[1] = smem_buffer[SH - 1][SW - 1]; }
}
int main(int argc, char **argv)
{
    const char mode = argc > 1 ? argv[1][0] : 'a';
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    int *d, *out; cudaMalloc(&d, GH * GW * 4); cudaMalloc(&out, 8);
    int *h = new int[GH * GW]; for (int i = 0; i < GH * GW; ++i) h[i] = i;
    cudaMemcpy(d, h, GH * GW * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * sizeof(int)};
    cuuint32_t box[2] = {SW, SH}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    if (mode == 'a') kernel<<<1, 128>>>(map, 64, 32, out);
    else if (mode == 'b') kernel_b<<<1, 128>>>(map, 64, 32, out);
    else kernel_c<<<1, 128>>>(map, 64, 32, out);
    cudaError_t e = cudaDeviceSynchronize();
    int ho[2] = {0, 0}; cudaMemcpy(ho, out, 8, cudaMemcpyDeviceToHost);
    printf("mode %c: %s  got %d %d want %d %d\n", mode, cudaGetErrorString(e), ho[0], ho[1], 32 * GW + 64, (32 + SH - 1) * GW + 64 + SW - 1);
    return 0;
}

// bisecting probe 3: u8 2-D map; argv: x y bw bh promo pitch
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int bytes, uint8_t* out)
{
    __shared__ alignas(128) uint8_t smem_buffer[128 * 64];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char** argv)
{
    const int x = atoi(argv[1]), y = atoi(argv[2]), BW = atoi(argv[3]), BH = atoi(argv[4]), promo = atoi(argv[5]), pitch = atoi(argv[6]);
    const int rows = 518;
    std::vector<uint8_t> h((size_t)pitch * rows);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *d_out; cudaMalloc(&d, h.size()); cudaMalloc(&d_out, 128 * 64);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm{};
    cuuint64_t size[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}; cuuint64_t stride[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    kernel<<<1, 32>>>(tm, x, y, BW * BH, d_out);
    cudaError_t ce = cudaDeviceSynchronize();
    std::vector<uint8_t> o(BW * BH); cudaMemcpy(o.data(), d_out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < BH * BW; i++) bad += o[i] != h[(size_t)(y + i / BW) * pitch + x + i % BW];
    printf("x=%d y=%d box=%dx%d promo=%d pitch=%d: encode %d sync=%d (%s) mismatches=%d\n", x, y, BW, BH, promo, pitch, (int)r, (int)ce, cudaGetErrorString(ce), bad);
    return 0;
}

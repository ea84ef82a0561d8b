// bisecting probe: canonical cuda::barrier + cde:: wrappers with (mode 1) a u8 2-D box 48x38, (mode 2) a u8 3-D box 48x38x1,
// (mode 3) the same 3-D box through raw PTX with a count-1 barrier and a parity wait
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 48, BH = 38;
__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int mode, int x, int y, int z, uint8_t* out)
{
    __shared__ alignas(128) uint8_t smem_buffer[BH * BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    __shared__ alignas(8) unsigned long long rawbar;
    if (mode < 3) {
        if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
        __syncthreads();
        barrier::arrival_token token;
        if (threadIdx.x == 0) {
            if (mode == 1) cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
            else cde::cp_async_bulk_tensor_3d_global_to_shared(&smem_buffer, &tensor_map, x, y, z, bar);
            token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
        } else token = bar.arrive();
        bar.wait(std::move(token));
    } else {
        const uint32_t b = sa(&rawbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (mode == 3) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BW * BH) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(sa(smem_buffer)), "l"(&tensor_map), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
            } else {
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(sa(smem_buffer)), "l"(&tensor_map), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BW * BH) : "memory");
            }
        }
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(b), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < BH * BW; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 1;
    const int pitch = 704, rows = 518, frames = 3;
    const size_t fbytes = ((size_t)pitch * rows + 255) & ~(size_t)255;
    std::vector<uint8_t> h(fbytes * frames);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *d_out; cudaMalloc(&d, h.size()); cudaMalloc(&d_out, BW * BH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm{};
    cuuint64_t size[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames}; cuuint64_t stride[2] = {(cuuint64_t)pitch, (cuuint64_t)fbytes};
    cuuint32_t box[3] = {BW, BH, 1}; cuuint32_t es[3] = {1, 1, 1};
    const int rank = mode == 1 ? 2 : 3;
    CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int x = 45, y = 33, z = mode == 1 ? 0 : 2;
    kernel<<<1, 32>>>(tm, mode, x, y, z, d_out);
    cudaError_t ce = cudaDeviceSynchronize();
    std::vector<uint8_t> o(BW * BH); cudaMemcpy(o.data(), d_out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < BH * BW; i++) bad += o[i] != h[(size_t)z * fbytes + (size_t)(y + i / BW) * pitch + x + i % BW];
    printf("mode %d: encode %d sync=%d (%s) mismatches=%d\n", mode, (int)r, (int)ce, cudaGetErrorString(ce), bad);
    return 0;
}

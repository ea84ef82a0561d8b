// tma_probe.cu — standalone check of the 3-D u8 tensor-map load used by the FAST / describe kernels (GPU box only):
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/probe/tma_probe.cu && /tmp/tma_probe <variant>
// variant bits: 1 = descriptor in global memory (else __grid_constant__ param), 2 = box width 64 (else 48),
//               4 = fence.proxy.async after mbarrier.init (else fence.mbarrier_init), 8 = static index 0 into the map array
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
struct alignas(64) Maps { CUtensorMap m[4]; };
__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ Maps maps, const CUtensorMap* gmaps, int which, int x, int y, int z, uint8_t* out, int variant, int bytes)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem + ((128u - (sa(smem) & 127u)) & 127u);
    __shared__ __align__(8) unsigned long long bar_s;
    const uint32_t bar = sa(&bar_s);
    if (threadIdx.x == 0) {
        const CUtensorMap* mp = (variant & 1) ? gmaps + which : (variant & 8) ? &maps.m[0] : &maps.m[which];
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        if (!(variant & 4)) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sa(tile)), "l"(mp), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    __syncwarp();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = tile[i];
}
int main(int argc, char** argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int pitch = 704, rows = 518, frames = 3;
    const size_t fbytes = ((size_t)pitch * rows + 255) & ~(size_t)255;
    std::vector<uint8_t> h(fbytes * frames);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *d_out; cudaMalloc(&d, h.size()); cudaMalloc(&d_out, 128 * 128);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    enc_t enc = (enc_t)fn;
    Maps maps; memset(&maps, 0, sizeof maps);
    const int BW = (variant & 2) ? 64 : 48, BH = 38;
    for (int l = 0; l < 4; l++) {
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fbytes};
        const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1}; const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) printf("encode level %d -> %d\n", l, (int)r);
    }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof maps); cudaMemcpy(gm, &maps, sizeof maps, cudaMemcpyHostToDevice);
    for (int t = 0; t < 3; t++) {
        const int x = t == 0 ? 45 : t == 1 ? 690 : 13, y = t == 0 ? 33 : t == 1 ? 500 : 0, z = t;
        cudaMemset(d_out, 0xEE, 128 * 128);
        probe<<<1, 32, BW * BH + 128>>>(maps, gm, t, x, y, z, d_out, variant, BW * BH);
        cudaError_t ce = cudaDeviceSynchronize();
        std::vector<uint8_t> o(BW * BH);
        cudaMemcpy(o.data(), d_out, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < BH; r++)
            for (int c = 0; c < BW; c++) {
                const int gx = x + c, gy = y + r;
                const uint8_t want = (gx < pitch && gy < rows) ? h[(size_t)z * fbytes + (size_t)gy * pitch + gx] : 0;
                bad += o[r * BW + c] != want;
            }
        printf("variant %d test %d: sync=%d (%s) mismatches=%d\n", variant, t, (int)ce, cudaGetErrorString(ce), bad);
    }
    return 0;
}

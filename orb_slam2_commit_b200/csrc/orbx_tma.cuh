// orbx_tma.cuh — TMA (cp.async.bulk.tensor) staging of pyramid tiles into shared memory, sm_100a.
//
// The raw pyramid of a working set is described to the TMA unit by one 3-D tensor map per level
// (x = byte column of the level buffer, y = buffer row incl. the 19-px apron, z = frame of the working set; u8
// elements) with a fixed box per consumer:
//   FAST      box = (tile pitch, tallest cell tile of the level, 1)     — the cell ROI of ORBextractor.cc:892-898 + 3-px ring
//   describe  box = (64, 43, 1)                                         — the IC_Angle / rBRIEF patch of :77-105, :110-152 + blur halo
// The innermost box coordinate has to be a multiple of 16 BYTES (measured on B200: any other x raises an illegal-instruction
// fault, tools/probe/tma_probe3.cu), so a consumer loads the box that starts at the aligned column at or left of its
// tile — hence boxes 15 columns wider than the tile — and offsets its base pointer by x & 15.
// One elected lane of a warp arms the warp's mbarrier with the box's byte count and issues the bulk copy; the warp then
// waits on the barrier's phase. No per-thread address arithmetic, no register staging, no st.shared.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "orbx_internal.cuh"

struct alignas(64) OrbxTmaps { CUtensorMap m[ORBX_MAX_LEVELS]; };

// host: encodes the per-level maps of a working set (orbx_capi.cu, at orbx_reserve time). Returns false (and leaves
// *err) when the driver entry point is missing or rejects a descriptor.
bool orbx_encode_level_maps(OrbxTmaps* out, const OrbxLevelGeom* lvl, int nlevels, uint8_t* raw, size_t frame_raw_bytes,
                            int frames, int box_w, const int* box_h, const char** err);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t orbx_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void orbx_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy before the copy uses it
}
__device__ __forceinline__ void orbx_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity. try_wait suspends the thread for a hardware-defined time slice per call; a
// copy that never completes (a broken descriptor, a lost arm) ends in a trap after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void orbx_mbar_wait(uint32_t bar, uint32_t parity)
{
    for (int spins = 0;; spins++) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1 << 22)) __trap();                 // every try_wait sleeps for a time slice: seconds, not a busy loop
    }
}
// box origin (x, y, z) in elements of the map's three dimensions
__device__ __forceinline__ void orbx_tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// L2 prefetch of a box (no shared-memory destination, no barrier)
__device__ __forceinline__ void orbx_tma_prefetch_3d(const CUtensorMap* map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}
// generic-proxy accesses to a shared-memory buffer are ordered before a later bulk copy into the same buffer
__device__ __forceinline__ void orbx_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// orbx_pyramid.cu — image pyramid resident in HBM (replaces ORBextractor::ComputePyramid,
// ORBextractor.cc:1215-1250 = cv::resize INTER_LINEAR 8U chain + cv::copyMakeBorder REFLECT_101).
//
// One launch per level (level l is resampled from level l-1, so levels are dependent). Each thread produces four
// horizontally adjacent bytes of the level buffer INCLUDING the 19-px apron: an apron byte is the payload byte at
// the reflected coordinate, recomputed instead of re-read, so every level is written exactly once and its border
// needs no second pass. Arithmetic is OpenCV's fixed-point bilinear (11-bit coefficients, tables built on the host
// in OpenCV's float/double sequence): bit-exact, integer only.
#include "orbx_internal.cuh"

#define PYR_TX 64
#define PYR_TY 4

__device__ __forceinline__ int reflect101(int p, int len)
{
    // |p| never exceeds len by more than the 19-px apron + 3 and len >= 62, so one fold per side is enough
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

#define PYR_RPT 8        // buffer rows per thread

// level 0: copy of the input frame (+apron)
__global__ void __launch_bounds__(PYR_TX * PYR_TY) pyr_level0_kernel(OrbxFrameLayout L, const uint8_t* __restrict__ img, int stride,
                                                         size_t frame_pitch)
{
    const OrbxLevelGeom g = L.lvl[0];
    const int t = blockIdx.x * PYR_TX + threadIdx.x;        // group of 4 buffer columns starting at column 12
    const int rb0 = (blockIdx.y * PYR_TY + threadIdx.y) * PYR_RPT;
    const int cb = 12 + 4 * t;
    const int rows = g.h + 2 * ORBX_EDGE;
    if (cb >= ORBX_XOFF + g.w + ORBX_EDGE || rb0 >= rows) return;
    int xr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) xr[k] = reflect101(cb + k - ORBX_XOFF, g.w);
    const uint8_t* fimg = img + (size_t)blockIdx.z * frame_pitch;
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + g.raw_off + (size_t)rb0 * g.pitch + cb;
#pragma unroll
    for (int rr = 0; rr < PYR_RPT; rr++) {
        if (rb0 + rr >= rows) break;
        const uint8_t* src = fimg + (size_t)reflect101(rb0 + rr - ORBX_EDGE, g.h) * stride;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) out |= (uint32_t)__ldg(src + xr[k]) << (8 * k);
        *reinterpret_cast<uint32_t*>(dst + (size_t)rr * g.pitch) = out;
    }
}

// level 0 from an interleaved colour frame (3 or 4 channels): cv::cvtColor(.., CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY /
// CV_BGRA2GRAY) of Tracking::GrabImage* (Tracking.cc:174-199) fused into the level-0 copy. OpenCV 4.x 8-bit arithmetic:
// (R*9798 + G*19235 + B*3735 + 2^14) >> 15.  `rgb` != 0 <=> the first channel is R (mbRGB).
__global__ void __launch_bounds__(PYR_TX * PYR_TY) pyr_level0_color_kernel(OrbxFrameLayout L, const uint8_t* __restrict__ img,
                                                                            int stride, size_t frame_pitch, int channels, int rgb)
{
    const OrbxLevelGeom g = L.lvl[0];
    const int t = blockIdx.x * PYR_TX + threadIdx.x;
    const int rb0 = (blockIdx.y * PYR_TY + threadIdx.y) * PYR_RPT;
    const int cb = 12 + 4 * t;
    const int rows = g.h + 2 * ORBX_EDGE;
    if (cb >= ORBX_XOFF + g.w + ORBX_EDGE || rb0 >= rows) return;
    int xr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) xr[k] = reflect101(cb + k - ORBX_XOFF, g.w) * channels;
    const int c0 = rgb ? 9798 : 3735, c2 = rgb ? 3735 : 9798;
    const uint8_t* fimg = img + (size_t)blockIdx.z * frame_pitch;
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + g.raw_off + (size_t)rb0 * g.pitch + cb;
#pragma unroll
    for (int rr = 0; rr < PYR_RPT; rr++) {
        if (rb0 + rr >= rows) break;
        const uint8_t* src = fimg + (size_t)reflect101(rb0 + rr - ORBX_EDGE, g.h) * stride;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint8_t* px = src + xr[k];
            const int v = (__ldg(px) * c0 + __ldg(px + 1) * 19235 + __ldg(px + 2) * c2 + (1 << 14)) >> 15;
            out |= (uint32_t)v << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)rr * g.pitch) = out;
    }
}

// level 0 through a rectification map: cv::remap(im, imRect, M1, M2, INTER_LINEAR) of Examples/Stereo/stereo_euroc.cc:136-137
// fused into the level-0 write (the rectified frame never exists outside the pyramid). `map` holds OpenCV's fixed-point
// form of the CV_32FC1 map pair (built once on the host in OpenCV's sequence): x = (int16 ix) | (int16 iy) << 16,
// y = a | b << 5 with ix = sx >> 5, a = sx & 31, sx = cvRound(map1 * 32). Bilinear taps outside the source read 0
// (BORDER_CONSTANT), dst = (sum p*w + 2^14) >> 15 with w = {(32-a)(32-b), a(32-b), (32-a)b, ab} * 32.
__global__ void __launch_bounds__(PYR_TX * PYR_TY) pyr_level0_remap_kernel(OrbxFrameLayout L, const uint8_t* __restrict__ img,
                                                                            int stride, size_t frame_pitch,
                                                                            const uint2* __restrict__ map, int sw, int sh)
{
    const OrbxLevelGeom g = L.lvl[0];
    const int t = blockIdx.x * PYR_TX + threadIdx.x;
    const int rb0 = (blockIdx.y * PYR_TY + threadIdx.y) * PYR_RPT;
    const int cb = 12 + 4 * t;
    const int rows = g.h + 2 * ORBX_EDGE;
    if (cb >= ORBX_XOFF + g.w + ORBX_EDGE || rb0 >= rows) return;
    int xr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) xr[k] = reflect101(cb + k - ORBX_XOFF, g.w);
    const uint8_t* fimg = img + (size_t)blockIdx.z * frame_pitch;
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + g.raw_off + (size_t)rb0 * g.pitch + cb;
#pragma unroll 2
    for (int rr = 0; rr < PYR_RPT; rr++) {
        if (rb0 + rr >= rows) break;
        const uint2* mrow = map + (size_t)reflect101(rb0 + rr - ORBX_EDGE, g.h) * g.w;
        uint2 m[4];
#pragma unroll
        for (int k = 0; k < 4; k++) m[k] = __ldg(mrow + xr[k]);
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ix = (short)(m[k].x & 0xffffu), iy = (int)m[k].x >> 16;
            const int a = m[k].y & 31, b = (m[k].y >> 5) & 31;
            const bool x0 = (unsigned)ix < (unsigned)sw, x1 = (unsigned)(ix + 1) < (unsigned)sw;
            const bool y0 = (unsigned)iy < (unsigned)sh, y1 = (unsigned)(iy + 1) < (unsigned)sh;
            const uint8_t* p = fimg + (ptrdiff_t)iy * stride + ix;
            const int p00 = (x0 && y0) ? __ldg(p) : 0, p01 = (x1 && y0) ? __ldg(p + 1) : 0;
            const int p10 = (x0 && y1) ? __ldg(p + stride) : 0, p11 = (x1 && y1) ? __ldg(p + stride + 1) : 0;
            const int v = (p00 * ((32 - a) * (32 - b)) + p01 * (a * (32 - b)) + p10 * ((32 - a) * b) + p11 * (a * b) + 512) >> 10;
            out |= (uint32_t)v << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)rr * g.pitch) = out;
    }
}

// level l > 0 from level l-1. Each thread produces 4 adjacent bytes of PYR_RPT consecutive buffer rows: the four
// x-taps are fetched once and up to 16*PYR_RPT independent source-pixel loads are in flight per thread.
__global__ void __launch_bounds__(PYR_TX * PYR_TY, 8) pyr_resize_kernel(OrbxFrameLayout L, int level)
{
    const OrbxLevelGeom g = L.lvl[level];
    const OrbxLevelGeom s = L.lvl[level - 1];
    const int t = blockIdx.x * PYR_TX + threadIdx.x;
    const int rb0 = (blockIdx.y * PYR_TY + threadIdx.y) * PYR_RPT;
    const int cb = 12 + 4 * t;
    const int rows = g.h + 2 * ORBX_EDGE;
    if (cb >= ORBX_XOFF + g.w + ORBX_EDGE || rb0 >= rows) return;
    const uint8_t* sbase = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + s.raw_off + ORBX_XOFF;
    int sx[4], a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const OrbxResizeTap tx = L.taps[g.xtab_off + reflect101(cb + k - ORBX_XOFF, g.w)];
        sx[k] = tx.ofs; a0[k] = tx.c0; a1[k] = tx.c1;
    }
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + g.raw_off + (size_t)rb0 * g.pitch + cb;
#pragma unroll
    for (int rr = 0; rr < PYR_RPT; rr++) {
        if (rb0 + rr >= rows) break;
        const OrbxResizeTap ty = L.taps[g.ytab_off + reflect101(rb0 + rr - ORBX_EDGE, g.h)];
        // rows sy and sy+1 of the source payload; when sy is the last row its coefficient c1 is 0 and row sy+1 is
        // the (valid) apron row, so no clamp is needed — same for columns
        const uint8_t* r0 = sbase + (size_t)(ty.ofs + ORBX_EDGE) * s.pitch;
        const uint8_t* r1 = r0 + s.pitch;
        const int b0 = ty.c0, b1 = ty.c1;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int S0 = r0[sx[k]] * a0[k] + r0[sx[k] + 1] * a1[k];
            const int S1 = r1[sx[k]] * a0[k] + r1[sx[k] + 1] * a1[k];
            const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xff) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)rr * g.pitch) = out;
    }
}

void orbx_launch_pyramid(const OrbxFrameLayout& L, const OrbxLevelGeom* h_lvl, const uint8_t* d_img, int w, int h,
                         int stride, size_t frame_pitch, int nframes, cudaStream_t st, int channels, int rgb,
                         const uint2* d_remap, int src_w, int src_h)
{
    (void)w; (void)h;
    for (int l = 0; l < L.nlevels; l++) {
        const OrbxLevelGeom& g = h_lvl[l];
        const int groups = (ORBX_XOFF + g.w + ORBX_EDGE - 12 + 3) / 4;
        const int rows = g.h + 2 * ORBX_EDGE;
        dim3 block(PYR_TX, PYR_TY);
        dim3 grid((groups + PYR_TX - 1) / PYR_TX, (rows + PYR_TY * PYR_RPT - 1) / (PYR_TY * PYR_RPT), nframes);
        if (l == 0 && d_remap) pyr_level0_remap_kernel<<<grid, block, 0, st>>>(L, d_img, stride, frame_pitch, d_remap, src_w, src_h);
        else if (l == 0 && channels > 1) pyr_level0_color_kernel<<<grid, block, 0, st>>>(L, d_img, stride, frame_pitch, channels, rgb);
        else if (l == 0) pyr_level0_kernel<<<grid, block, 0, st>>>(L, d_img, stride, frame_pitch);
        else pyr_resize_kernel<<<grid, block, 0, st>>>(L, l);
    }
}

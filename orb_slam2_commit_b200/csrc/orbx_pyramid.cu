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

// level l > 0 from level l-1. Each thread produces 4 adjacent bytes of PYR_RPT consecutive buffer rows.
// Generic path (threads on the apron: reflected, non-monotonic x-taps): byte loads per tap. Kept out of line so that its
// registers do not burden the fast path below.
__device__ __noinline__ void pyr_resize_rows_generic(const OrbxResizeTap* __restrict__ xtab, const OrbxResizeTap* __restrict__ ytab,
                                                     const int gw, const int gh, const int gpitch, const int spitch,
                                                     const uint8_t* __restrict__ sbase, uint8_t* __restrict__ dst,
                                                     const int cb, const int rb0, const int rows)
{
    int sx[4], a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const OrbxResizeTap tx = xtab[reflect101(cb + k - ORBX_XOFF, gw)];
        sx[k] = tx.ofs; a0[k] = tx.c0; a1[k] = tx.c1;
    }
#pragma unroll 2
    for (int rr = 0; rr < PYR_RPT; rr++) {
        if (rb0 + rr >= rows) break;
        const OrbxResizeTap ty = ytab[reflect101(rb0 + rr - ORBX_EDGE, gh)];
        // rows sy and sy+1 of the source payload; when sy is the last row its coefficient c1 is 0 and row sy+1 is
        // the (valid) apron row, so no clamp is needed — same for columns
        const uint8_t* r0 = sbase + (size_t)(ty.ofs + ORBX_EDGE) * spitch;
        const uint8_t* r1 = r0 + spitch;
        const int b0 = ty.c0, b1 = ty.c1;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int S0 = r0[sx[k]] * a0[k] + r0[sx[k] + 1] * a1[k];
            const int S1 = r1[sx[k]] * a0[k] + r1[sx[k] + 1] * a1[k];
            const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xff) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)rr * gpitch) = out;
    }
}

// Fast path for threads whose four columns lie in the payload (no reflection): the eight x-taps of a source row sit
// within 8 consecutive bytes, so each source row is three aligned 32-bit loads, two funnel shifts that bring byte
// sx[0] to the front, two byte permutes that lay the (left, right) tap pairs of two outputs side by side, and one
// 2-way dot product per output against the packed coefficient pair — instead of eight byte loads with 64-bit
// address arithmetic.
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
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + g.raw_off + (size_t)rb0 * g.pitch + cb;
    // taps are stored as (ofs, c0, c1, pad) shorts: one 64-bit load each; c0 | c1 << 16 is bytes 2..5. On the apron the
    // reflected columns give the same taps in descending order: the window starts at the smallest offset either way.
    unsigned cf[4];
    int x[4];
    const uint2* tp = reinterpret_cast<const uint2*>(L.taps + g.xtab_off);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint2 tk = tp[reflect101(cb + k - ORBX_XOFF, g.w)];
        x[k] = (short)(tk.x & 0xffff);
        cf[k] = __byte_perm(tk.x, tk.y, 0x5432);
    }
    const int x0 = min(min(x[0], x[1]), min(x[2], x[3]));
    const int d0 = x[0] - x0, d1 = x[1] - x0, d2 = x[2] - x0, d3 = x[3] - x0;
    const bool fast = max(max(d0, d1), max(d2, d3)) <= 6;
    const int ab = x0 & ~3, s8 = (x0 & 3) * 8;
    const unsigned sel01 = (unsigned)(d0 | ((d0 + 1) << 4) | (d1 << 8) | ((d1 + 1) << 12));
    const unsigned sel23 = (unsigned)(d2 | ((d2 + 1) << 4) | (d3 << 8) | ((d3 + 1) << 12));
    if (!fast) {
        pyr_resize_rows_generic(L.taps + g.xtab_off, L.taps + g.ytab_off, g.w, g.h, g.pitch, s.pitch, sbase, dst, cb, rb0, rows);
        return;
    }
    const uint8_t* sb = sbase + ab;
#pragma unroll
    for (int rr = 0; rr < PYR_RPT; rr++) {
        if (rb0 + rr >= rows) break;
        const OrbxResizeTap ty = L.taps[g.ytab_off + reflect101(rb0 + rr - ORBX_EDGE, g.h)];
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(sb + (size_t)(ty.ofs + ORBX_EDGE) * s.pitch);
        const uint32_t* p1 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(p0) + s.pitch);
        const int b0 = ty.c0, b1 = ty.c1;
        const uint32_t u0 = p0[0], u1 = p0[1], u2 = p0[2], v0 = p1[0], v1 = p1[1], v2 = p1[2];
        const uint32_t A0 = __funnelshift_r(u0, u1, s8), A1 = __funnelshift_r(u1, u2, s8);
        const uint32_t B0 = __funnelshift_r(v0, v1, s8), B1 = __funnelshift_r(v1, v2, s8);
        const uint32_t qa01 = __byte_perm(A0, A1, sel01), qa23 = __byte_perm(A0, A1, sel23);
        const uint32_t qb01 = __byte_perm(B0, B1, sel01), qb23 = __byte_perm(B0, B1, sel23);
        int S0[4], S1[4];
        S0[0] = (int)__dp2a_lo(cf[0], qa01, 0u); S0[1] = (int)__dp2a_hi(cf[1], qa01, 0u);
        S0[2] = (int)__dp2a_lo(cf[2], qa23, 0u); S0[3] = (int)__dp2a_hi(cf[3], qa23, 0u);
        S1[0] = (int)__dp2a_lo(cf[0], qb01, 0u); S1[1] = (int)__dp2a_hi(cf[1], qb01, 0u);
        S1[2] = (int)__dp2a_lo(cf[2], qb23, 0u); S1[3] = (int)__dp2a_hi(cf[3], qb23, 0u);
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = (((b0 * (S0[k] >> 4)) >> 16) + ((b1 * (S1[k] >> 4)) >> 16) + 2) >> 2;
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

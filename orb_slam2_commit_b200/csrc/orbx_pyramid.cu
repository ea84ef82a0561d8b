// orbx_pyramid.cu — image pyramid resident in HBM (replaces ORBextractor::ComputePyramid,
// ORBextractor.cc:1215-1250 = cv::resize INTER_LINEAR 8U chain + cv::copyMakeBorder REFLECT_101).
//
// One launch per level (level l is resampled from level l-1, so levels are dependent). Each thread produces four
// horizontally adjacent bytes of the level buffer INCLUDING the 19-px apron: an apron byte is the payload byte at
// the reflected coordinate, recomputed instead of re-read, so every level is written exactly once and its border
// needs no second pass. Arithmetic is OpenCV's fixed-point bilinear (11-bit coefficients, tables built on the host
// in OpenCV's float/double sequence): bit-exact, integer only.
#include "orbx_internal.cuh"
#include "orbx_tma.cuh"
#include <algorithm>
#include <vector>

#define PYR_TX 64
#define PYR_TY 4

__device__ __forceinline__ int reflect101(int p, int len)
{
    // |p| never exceeds len by more than the 19-px apron + 3 and len >= 62, so one fold per side is enough
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

#define PYR_RPT 8        // buffer rows per thread

// level 0: copy of the input frame (+apron). Thread = 16 adjacent bytes of PYR_L0_ROWS buffer rows, written with one 128-bit
// store per row. Groups that lie inside the payload and whose source is 16-byte aligned (frame base, stride) read with one
// 128-bit load: they are the "interior" CTAs (blockIdx.x < nxb_in, lane = group). The other groups (apron columns:
// BORDER_REFLECT_101, or every group of an unaligned frame) gather bytes, ~15x the instructions — they are packed densely into
// their own CTAs (blockIdx.x == nxb_in: thread -> (row quad, edge group) by division), so that no warp of the interior runs the
// gather path for the sake of two lanes (with lane = group across the whole row every warp did: 155 k instead of 20 k
// warp-instructions per VGA frame).
#define PYR_L0_ROWS 4
__global__ void __launch_bounds__(256) pyr_level0_kernel(OrbxFrameLayout L, const uint8_t* __restrict__ img, int stride,
                                                         size_t frame_pitch, int nxb_in, int gi1, int n_edge)
{
    const OrbxLevelGeom* __restrict__ gp = L.lvl;
    const int gw = gp->w, gh = gp->h, gpitch = gp->pitch;
    const int rows = gh + 2 * ORBX_EDGE;
    const uint8_t* fimg = img + (size_t)blockIdx.z * frame_pitch;
    if ((int)blockIdx.x < nxb_in) {
        const int grp = 2 + blockIdx.x * 32 + threadIdx.x;
        const int rb0 = (blockIdx.y * 8 + threadIdx.y) * PYR_L0_ROWS;
        if (grp >= gi1 || rb0 >= rows) return;
        const int cb = 16 * grp;
        uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + gp->raw_off + (size_t)rb0 * gpitch + cb;
        const uint8_t* src = fimg + (cb - ORBX_XOFF);
        uint4 v[PYR_L0_ROWS];
#pragma unroll
        for (int rr = 0; rr < PYR_L0_ROWS; rr++)
            v[rr] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)reflect101(min(rb0 + rr, rows - 1) - ORBX_EDGE, gh) * stride));
#pragma unroll
        for (int rr = 0; rr < PYR_L0_ROWS; rr++)
            if (rb0 + rr < rows) *reinterpret_cast<uint4*>(dst + (size_t)rr * gpitch) = v[rr];
        return;
    }
    const int t = blockIdx.y * 256 + threadIdx.y * 32 + threadIdx.x;
    const int rq = t / n_edge, e = t - rq * n_edge;
    const int rb0 = rq * PYR_L0_ROWS;
    if (rb0 >= rows) return;
    const int cb = 16 * (e < 2 ? e : gi1 + e - 2);           // buffer column of the group (columns 13..31 hold the left apron)
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + gp->raw_off + (size_t)rb0 * gpitch + cb;
    int xr[16];
#pragma unroll
    for (int k = 0; k < 16; k++) xr[k] = reflect101(min(max(cb + k, ORBX_XOFF - ORBX_EDGE), ORBX_XOFF + gw + ORBX_EDGE - 1) - ORBX_XOFF, gw);
    for (int rr = 0; rr < PYR_L0_ROWS; rr++) {
        if (rb0 + rr >= rows) break;
        const uint8_t* src = fimg + (size_t)reflect101(rb0 + rr - ORBX_EDGE, gh) * stride;
        uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 16; k++) o[k >> 2] |= (uint32_t)__ldg(src + xr[k]) << (8 * (k & 3));
        *reinterpret_cast<uint4*>(dst + (size_t)rr * gpitch) = make_uint4(o[0], o[1], o[2], o[3]);
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
// Generic path: byte loads per tap. Used (as its own kernel) for levels where some group of four outputs spreads its x-taps
// over more than 8 source bytes (scale factors above ~1.7), which the host finds out from the tap table.
__device__ __forceinline__ void pyr_resize_rows_generic(const OrbxResizeTap* __restrict__ xtab, const OrbxResizeTap* __restrict__ ytab,
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

// Fast path (all columns whose four x-taps lie within 8 consecutive source bytes — every thread at the reference's scale
// factors, apron included: reflected columns give the same taps in descending order). A source row costs three aligned
// 32-bit loads, two funnel shifts that bring the first tap to byte 0, two byte permutes that lay the (left, right) tap
// pairs of two outputs side by side and one 2-way dot product (dp2a) per output against the packed coefficient pair.
// HORIZONTAL RESULTS ARE SHARED BETWEEN OUTPUT ROWS: at a scale of 1.2 the source rows (s, s+1) of consecutive output rows
// overlap — the next row's s is the previous row's s+1 five times out of six — so a thread that walks PYR_RRPT output rows
// keeps the four horizontal sums of its last source row and computes 1.2 instead of 2 source rows per output row.
// The vertical pass takes (b * (S >> 4)) >> 16 as one multiply-high against b << 16.
// The kernel above the generic one: a CTA produces a PYR_TW x PYR_TH tile of the level buffer (apron included). ONE bulk
// tensor copy (TMA, orbx_tma.cuh) brings the source rectangle the tile needs — all x-taps and both y-taps of every output,
// found per tile column / tile row on the host (orbx_pyr_tiles) — from level l-1 into shared memory; the threads then work
// from shared memory (the global-load version of this kernel sat at 48 % long-scoreboard stalls: two dependent L2 round
// trips per output row). Thread = 4 adjacent output bytes x PYR_RRPT rows; the four horizontal sums of a source row are
// shared between output rows: at a scale of 1.2 the rows (s, s+1) of consecutive outputs overlap — the next output's s is the
// previous one's s+1 five times out of six — so a thread computes about 1.2 instead of 2 source rows per output row.
#define PYR_TW 128       // tile width in bytes (32 threads x 4)
#define PYR_RRPT 16      // output rows per thread
#define PYR_TH (PYR_TY2 * PYR_RRPT)
#define PYR_TY2 4        // thread rows (warps) per CTA
// Everything a thread needs about its four columns and about each of its rows is a pure function of the level geometry, so
// the host lays it out once (orbx_pyr_tiles): per group of four buffer columns {word-aligned source column, funnel-shift
// amount, the two byte-permute selectors, the four packed coefficient pairs} (two 128-bit loads), per buffer row {source row
// x box pitch, b0 << 16, b1 << 16} (one 128-bit load, the same address for the whole warp) — reflections, clamps and selector
// arithmetic were 30 % of the kernel's instructions when every thread derived them itself.
__global__ void __launch_bounds__(32 * PYR_TY2, 8) pyr_resize_kernel(OrbxFrameLayout L, int level, const __grid_constant__ OrbxTmaps maps)
{
    extern __shared__ __align__(128) uint8_t pyr_smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(16) uint4 s_yr[PYR_TY2][PYR_RRPT];    // the row entries of each warp (a warp's lanes share their rows)
    const OrbxLevelGeom* __restrict__ gp = L.lvl + level;
    const int gw = gp->w, gh = gp->h, gpitch = gp->pitch;
    const int bw = gp->pyr_box_w;
    const int2 tx0 = reinterpret_cast<const int2*>(L.pyr_tiles + gp->pyr_tile_off)[blockIdx.x];                 // (source buffer column of the box, -)
    const int2 ty0 = reinterpret_cast<const int2*>(L.pyr_tiles + gp->pyr_tile_off)[gridDim.x + blockIdx.y];     // (source buffer row of the box, that row x box pitch)
    uint8_t* tile = pyr_smem + ((128u - (orbx_smem_addr(pyr_smem) & 127u)) & 127u);
    const uint32_t bar = orbx_smem_addr(&s_bar);
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) {
        orbx_mbar_init(bar, 1);
        orbx_mbar_expect_tx(bar, (uint32_t)(bw * gp->pyr_box_h));
        orbx_tma_load_3d(orbx_smem_addr(tile), &maps.m[level], tx0.x, ty0.x, L.frame0 + blockIdx.z, bar);
    }
    __syncthreads();                                         // the barrier is initialised before anyone waits on it
    const int gi = 32 * blockIdx.x + threadIdx.x;            // group of four buffer columns: cb = 12 + 4 * gi
    const int cb = 12 + 4 * gi;
    const int rb0 = PYR_TH * blockIdx.y + threadIdx.y * PYR_RRPT;
    const int rows = gh + 2 * ORBX_EDGE;
    const bool live = cb < ORBX_XOFF + gw + ORBX_EDGE && rb0 < rows;
    const uint4* xg = reinterpret_cast<const uint4*>(L.pyr_tiles + gp->pyr_xg_off) + 2 * gi;
    const uint4 xa = xg[0], xc = xg[1];                     // (source column, s8, sel01, sel23), coefficient pairs
    // the 16 row entries go through shared memory (one coalesced load per warp while the tile is in flight): read straight
    // from global inside the row loop they were the kernel's main stall (long scoreboard, 4-6 per issue)
    if (threadIdx.x < PYR_RRPT) s_yr[threadIdx.y][threadIdx.x] = (reinterpret_cast<const uint4*>(L.pyr_tiles + gp->pyr_yr_off) + rb0)[threadIdx.x];
    __syncwarp();
    const uint4* yr = s_yr[threadIdx.y];
    const unsigned s8 = xa.y, sel01 = xa.z, sel23 = xa.w;
    const uint8_t* sb = tile + ((int)xa.x - tx0.x) - ty0.y;  // + (source buffer row x box pitch) = the thread's window in that row
    orbx_mbar_wait(bar, 0);
    if (!live) return;
    uint8_t* dst = L.raw + (size_t)blockIdx.z * L.frame_raw_bytes + gp->raw_off + (size_t)rb0 * gpitch + cb;
    // (S >> 4) of the four outputs for one source row: three words from the thread's aligned window, two funnel shifts that
    // bring the first tap to byte 0, two byte permutes that lay the (left, right) tap pairs of two outputs side by side, one
    // 2-way dot product per output against the packed coefficient pair
    auto hrow = [&](const int off, unsigned (&t)[4]) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(sb + off);
        const uint32_t u0 = p[0], u1 = p[1], u2 = p[2];
        const uint32_t A0 = __funnelshift_r(u0, u1, s8), A1 = __funnelshift_r(u1, u2, s8);
        const uint32_t q01 = __byte_perm(A0, A1, sel01), q23 = __byte_perm(A0, A1, sel23);
        t[0] = __dp2a_lo(xc.x, q01, 0u) >> 4; t[1] = __dp2a_hi(xc.y, q01, 0u) >> 4;
        t[2] = __dp2a_lo(xc.z, q23, 0u) >> 4; t[3] = __dp2a_hi(xc.w, q23, 0u) >> 4;
    };
    int have = -0x7fffffff;                                  // source row (x box pitch) whose sums `tb` holds
    unsigned ta[4], tb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int rr = 0; rr < PYR_RRPT; rr++) {
        if (rb0 + rr >= rows) break;
        // rows sy and sy+1 of the source payload (buffer rows sy + 19, sy + 20); when sy is the last row its coefficient c1
        // is 0 and row sy+1 is the (valid) apron row, so no clamp is needed — same for columns
        const uint4 y = yr[rr];                              // (source row x box pitch, b0 << 16, b1 << 16, -)
        const int off = (int)y.x;
        if (have == off) {                                   // warp-uniform: a warp's lanes share their rows
#pragma unroll
            for (int k = 0; k < 4; k++) ta[k] = tb[k];
        } else hrow(off, ta);
        hrow(off + bw, tb);
        have = off + bw;
        // ((b0 * (S0 >> 4)) >> 16) as one multiply-high against b0 << 16
        unsigned v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (__umulhi(y.y, ta[k]) + __umulhi(y.z, tb[k]) + 2u) >> 2;
        *reinterpret_cast<uint32_t*>(dst + rr * gpitch) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    }
}

// Host: the tables of the resize tiles of level `g` (from level l-1): per tile column the 16-byte aligned source buffer column
// where its box starts, per tile row the source buffer row (and that row x box pitch); box_w / box_h = the largest extent any
// tile needs (+ the alignment slack and the 12-byte windows the threads read); then the per-column-group and per-row thread
// tables described above the kernel. Layout of `out` from g.pyr_tile_off: ntx int2, nty int2; from g.pyr_xg_off: 8 ints per
// column group; from g.pyr_yr_off: 4 ints per buffer row (both 16-byte aligned, padded to whole tiles).
void orbx_pyr_tiles(OrbxLevelGeom& g, const OrbxResizeTap* h_taps, std::vector<int>& out)
{
    auto refl = [](int p, int len) { p = p < 0 ? -p : p; return p >= len ? 2 * (len - 1) - p : p; };
    const int cols = ORBX_XOFF + g.w + ORBX_EDGE, rows = g.h + 2 * ORBX_EDGE;
    const int ntx = (cols - 12 + PYR_TW - 1) / PYR_TW, nty = (rows + PYR_TH - 1) / PYR_TH;
    while (out.size() & 3) out.push_back(0);
    g.pyr_tile_off = (int)out.size(); g.pyr_ntx = ntx; g.pyr_nty = nty;
    int bw = 16, bh = 2;
    for (int bx = 0; bx < ntx; bx++) {
        int lo = 1 << 30, hi = -1;
        for (int c = 12 + PYR_TW * bx; c < std::min(12 + PYR_TW * (bx + 1), cols); c++) {
            const int o = h_taps[g.xtab_off + refl(c - ORBX_XOFF, g.w)].ofs; lo = std::min(lo, o); hi = std::max(hi, o + 1);
        }
        const int start = (lo + ORBX_XOFF) & ~15;
        bw = std::max(bw, ((hi + ORBX_XOFF) & ~3) + 12 - start);          // a thread reads three words from its aligned window start
        out.push_back(start); out.push_back(0);
    }
    const size_t ytiles = out.size();
    for (int by = 0; by < nty; by++) {
        int lo = 1 << 30, hi = -1;
        for (int r = PYR_TH * by; r < std::min(PYR_TH * (by + 1), rows); r++) {
            const int o = h_taps[g.ytab_off + refl(r - ORBX_EDGE, g.h)].ofs; lo = std::min(lo, o); hi = std::max(hi, o + 1);
        }
        bh = std::max(bh, hi - lo + 1);
        out.push_back(lo + ORBX_EDGE); out.push_back(0);
    }
    bw = (bw + 15) & ~15;
    g.pyr_box_w = bw; g.pyr_box_h = bh;
    for (int by = 0; by < nty; by++) out[ytiles + 2 * by + 1] = out[ytiles + 2 * by] * bw;
    while (out.size() & 3) out.push_back(0);
    g.pyr_xg_off = (int)out.size();
    for (int gi = 0; gi < ntx * 32; gi++) {
        const int cb = 12 + 4 * gi;
        int x[4]; unsigned cf[4];
        for (int k = 0; k < 4; k++) {
            const OrbxResizeTap& t = h_taps[g.xtab_off + refl(std::min(cb + k, cols - 1) - ORBX_XOFF, g.w)];
            x[k] = t.ofs; cf[k] = (unsigned)(unsigned short)t.c0 | ((unsigned)(unsigned short)t.c1 << 16);
        }
        const int x0 = std::min(std::min(x[0], x[1]), std::min(x[2], x[3]));
        const int d0 = x[0] - x0, d1 = x[1] - x0, d2 = x[2] - x0, d3 = x[3] - x0;
        out.push_back((x0 + ORBX_XOFF) & ~3);
        out.push_back((x0 & 3) * 8);
        out.push_back(d0 | ((d0 + 1) << 4) | (d1 << 8) | ((d1 + 1) << 12));
        out.push_back(d2 | ((d2 + 1) << 4) | (d3 << 8) | ((d3 + 1) << 12));
        for (int k = 0; k < 4; k++) out.push_back((int)cf[k]);
    }
    g.pyr_yr_off = (int)out.size();
    for (int r = 0; r < nty * PYR_TH; r++) {
        const OrbxResizeTap& t = h_taps[g.ytab_off + refl(std::min(r, rows - 1) - ORBX_EDGE, g.h)];
        out.push_back((t.ofs + ORBX_EDGE) * bw);
        out.push_back((int)((unsigned)(unsigned short)t.c0 << 16));
        out.push_back((int)((unsigned)(unsigned short)t.c1 << 16));
        out.push_back(0);
    }
}

__global__ void __launch_bounds__(PYR_TX * PYR_TY) pyr_resize_generic_kernel(OrbxFrameLayout L, int level)
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
    pyr_resize_rows_generic(L.taps + g.xtab_off, L.taps + g.ytab_off, g.w, g.h, g.pitch, s.pitch, sbase, dst, cb, rb0, rows);
}

// true when every group of four adjacent buffer columns of level `g` (apron included) keeps its eight x-taps within 8
// consecutive source bytes: the condition of the fast kernel
bool orbx_pyr_fast_ok(const OrbxLevelGeom& g, const OrbxResizeTap* h_taps)
{
    auto refl = [&](int p) { p = p < 0 ? -p : p; return p >= g.w ? 2 * (g.w - 1) - p : p; };
    for (int cb = 12; cb < ORBX_XOFF + g.w + ORBX_EDGE; cb += 4) {
        int lo = 1 << 30, hi = -1;
        for (int k = 0; k < 4; k++) { const int x = h_taps[g.xtab_off + refl(cb + k - ORBX_XOFF)].ofs; lo = std::min(lo, x); hi = std::max(hi, x); }
        if (hi - lo > 6) return false;
    }
    return true;
}

void orbx_launch_pyramid(const OrbxFrameLayout& L, const OrbxTmaps& maps, const OrbxLevelGeom* h_lvl, const uint8_t* d_img, int w, int h,
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
        else if (l == 0) {
            const int ngroups = (ORBX_XOFF + g.w + ORBX_EDGE + 15) / 16, rq = (rows + PYR_L0_ROWS - 1) / PYR_L0_ROWS;
            const bool aligned = ((reinterpret_cast<uintptr_t>(d_img) | (uintptr_t)stride | (uintptr_t)frame_pitch) & 15) == 0;
            const int gi1 = aligned ? std::max((ORBX_XOFF + g.w) / 16, 2) : 2;      // interior groups: [2, gi1)
            const int n_edge = 2 + ngroups - gi1, nxb_in = (gi1 - 2 + 31) / 32;
            const int gy = std::max(nxb_in ? (rq + 7) / 8 : 0, (rq * n_edge + 255) / 256);
            pyr_level0_kernel<<<dim3(nxb_in + 1, gy, nframes), dim3(32, 8), 0, st>>>(L, d_img, stride, frame_pitch, nxb_in, gi1, n_edge);
        }
        else if (g.resize_fast) {
            static OrbxSmemMark mk = {};
            const size_t smem = (size_t)g.pyr_box_w * g.pyr_box_h + 128;
            orbx_need_smem(pyr_resize_kernel, mk, smem);
            pyr_resize_kernel<<<dim3(g.pyr_ntx, g.pyr_nty, nframes), dim3(32, PYR_TY2), smem, st>>>(L, l, maps);
        }
        else pyr_resize_generic_kernel<<<grid, block, 0, st>>>(L, l);
    }
}

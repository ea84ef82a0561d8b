// orbx_blur.cu — cv::GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level
// (ORBextractor.cc:1188-1190), dense, into a second pyramid in HBM that only the descriptor kernel reads.
//
// Why dense: a keypoint's 256 test pairs sample a 37x37 window of the blurred level, and 1000 keypoints per frame need
// 1.37 M blurred pixels where the whole pyramid has 0.95 M — blurring patch by patch inside the descriptor kernel (the
// round-1 design) cost 750 of its 1500 warp-instructions per keypoint. The step is instruction-bound, not HBM-bound, so the
// extra 2 x 0.95 MB per frame of traffic is the cheaper currency.
//
// OpenCV's 8-bit path is fixed point: kernel [18,34,48,56,48,34,18]/256, horizontal sums exact in 16 bits, vertical
// (sum + 2^15) >> 16. Border = REFLECT_101 of the level, which is exactly what the 19-px apron of the raw level buffer holds,
// so the kernel never reflects: it blurs the apron-extended buffer (the blurred buffer has the raw buffer's geometry; its
// outermost 3 apron pixels are not written and never read: samples stay within 18 px of a keypoint that is at least 16 px
// inside the level).
//
// ONE HALF-WARP PER UNIT of 64 columns x R rows (R <= 32, even), all levels and frames in one launch, no block barrier (a
// full-warp unit of 128 columns left 17 % of the lanes beyond the right edge of the levels; half-warp units 6 %). The two
// halves of a warp take two consecutive units of the table — any two: each half has its own 80 x (R + 6) source box, copied
// into its shared-memory slot by the TMA unit on the warp's mbarrier. Lane = 4 adjacent output bytes, walked top to bottom:
//   horizontal: the 7 taps of each of the 4 outputs lie in 3 consecutive words (u0 u1 u2): 6 funnel shifts line them up, two
//               4-way byte dot products (dp4a) per output against (18,34,48,56) and (48,34,18,0)            3.5 ops / pixel
//   vertical:   the 16-bit sums of two consecutive rows are packed per column (one byte permute per pair); an output row is
//               four 2-way dot products (dp2a) of the packed pairs against the weight pairs for its parity, on top of the
//               rounding constant; rows are produced two at a time from a rotating window of four pairs     4.5 ops / pixel
//   byte 2 of each accumulator is the pixel: three byte permutes pack a lane's four outputs into one 32-bit store.
#include "orbx_internal.cuh"
#include "orbx_tma.cuh"

#define BLUR_WARPS 4
#define BLUR_SLOT (ORBX_BLUR_BOX_W * (ORBX_BLUR_MAX_ROWS + 6) + 32)      // 80 x 38 = 3040 -> 3072 (multiple of 128), one per half-warp
#define BLUR_PW (ORBX_BLUR_BOX_W / 4)                                    // box pitch in words

__global__ void __launch_bounds__(BLUR_WARPS * 32) blur_units_kernel(OrbxFrameLayout L, const OrbxBlurUnit* __restrict__ units, int nunits,
                                                                      const __grid_constant__ OrbxTmaps maps)
{
    extern __shared__ __align__(128) uint8_t blur_smem[];
    __shared__ __align__(8) unsigned long long s_bar[BLUR_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    const int unit0 = 2 * (blockIdx.x * BLUR_WARPS + wid), frame = blockIdx.y;
    if (unit0 >= nunits) return;                                         // whole warp
    const bool have = unit0 + half < nunits;                             // an odd table leaves the last warp's upper half idle
    const OrbxBlurUnit u = units[have ? unit0 + half : unit0];
    uint8_t* slot = blur_smem + ((128u - (orbx_smem_addr(blur_smem) & 127u)) & 127u) + (size_t)(2 * wid + half) * BLUR_SLOT;
    const uint32_t bar = orbx_smem_addr(&s_bar[wid]);
    const int rows_other = __shfl_xor_sync(0xffffffffu, (int)u.rows, 16);
    const int rows_max = max((int)u.rows, rows_other);
    if (lane == 0) {
        orbx_mbar_init(bar, 1);
        orbx_mbar_expect_tx(bar, (uint32_t)(ORBX_BLUR_BOX_W * (u.rows + 6 + (unit0 + 1 < nunits ? rows_other + 6 : 0))));
    }
    __syncwarp();                                                        // the barrier is initialised and armed before the copies and the wait
    if (hl == 0 && have) orbx_tma_load_3d(orbx_smem_addr(slot), &maps.m[u.level], u.c0, u.r0, L.frame0 + frame, bar);
    const OrbxLevelGeom* __restrict__ gp = L.lvl + u.level;
    const int gpitch = gp->pitch, brows = gp->h + 2 * ORBX_EDGE;
    const int col = u.c0 + 4 + 4 * hl;                                   // buffer column of the lane's first output byte
    uint8_t* dst = L.blur + (size_t)frame * L.frame_raw_bytes + gp->raw_off + (size_t)(u.r0 + 3) * gpitch + col;
    const bool live = have && col < gpitch;
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(slot) + hl;     // (u0, u1, u2) of box row i = tw[i * BLUR_PW + 0..2]
    orbx_mbar_wait(bar, 0);

    const unsigned WA = 18u | (34u << 8) | (48u << 16) | (56u << 24), WB = 48u | (34u << 8) | (18u << 16);
    auto hrow = [&](const int i, unsigned (&h)[4]) {
        const uint32_t u0 = tw[i * BLUR_PW], u1 = tw[i * BLUR_PW + 1], u2 = tw[i * BLUR_PW + 2];
        // output byte j of u1 takes box bytes j+1 .. j+7 counted from u0
        h[0] = __dp4a(__funnelshift_r(u1, u2, 8), WB, __dp4a(__funnelshift_r(u0, u1, 8), WA, 0u));
        h[1] = __dp4a(__funnelshift_r(u1, u2, 16), WB, __dp4a(__funnelshift_r(u0, u1, 16), WA, 0u));
        h[2] = __dp4a(__funnelshift_r(u1, u2, 24), WB, __dp4a(__funnelshift_r(u0, u1, 24), WA, 0u));
        h[3] = __dp4a(u2, WB, __dp4a(u1, WA, 0u));
    };
    // P[k][c] = (h of box row 2k, h of box row 2k+1) of column c, k taken modulo 4
    unsigned P[4][4];
    auto pair = [&](const int k) {
        unsigned he[4], ho[4];
        hrow(2 * k, he); hrow(2 * k + 1, ho);
#pragma unroll
        for (int c = 0; c < 4; c++) P[k & 3][c] = __byte_perm(he[c], ho[c], 0x5410);
    };
    pair(0); pair(1); pair(2);
#pragma unroll
    for (int a = 0; a < ORBX_BLUR_MAX_ROWS / 2; a++) {
        if (2 * a >= rows_max) break;                                     // warp-uniform
        pair(a + 3);
        const unsigned(&p0)[4] = P[a & 3], (&p1)[4] = P[(a + 1) & 3], (&p2)[4] = P[(a + 2) & 3], (&p3)[4] = P[(a + 3) & 3];
        unsigned e[4], o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // output row 2a: box rows 2a .. 2a+6; output row 2a+1: box rows 2a+1 .. 2a+7
            e[c] = __dp2a_lo(p3[c], 18u, __dp2a_lo(p2[c], 48u | (34u << 8), __dp2a_lo(p1[c], 48u | (56u << 8), __dp2a_lo(p0[c], 18u | (34u << 8), 32768u))));
            o[c] = __dp2a_lo(p3[c], 34u | (18u << 8), __dp2a_lo(p2[c], 56u | (48u << 8), __dp2a_lo(p1[c], 34u | (48u << 8), __dp2a_lo(p0[c], 18u << 8, 32768u))));
        }
        const uint32_t we = __byte_perm(__byte_perm(e[0], e[1], 0x0062), __byte_perm(e[2], e[3], 0x0062), 0x5410);
        const uint32_t wo = __byte_perm(__byte_perm(o[0], o[1], 0x0062), __byte_perm(o[2], o[3], 0x0062), 0x5410);
        const int r = 2 * a < u.rows ? u.r0 + 3 + 2 * a : brows;
        if (live && r < brows) *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * a) * gpitch) = we;
        if (live && r + 1 < brows) *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * a + 1) * gpitch) = wo;
    }
}

// Host: the units of one level. Blurred pixels are needed for payload coordinates [-2, w+2) x [-2, h+2); units start at
// buffer column 20 (box column 16) and buffer row 16 (box row 13).
void orbx_blur_units(const OrbxLevelGeom& g, int level, std::vector<OrbxBlurUnit>& out, int* rows_per_unit)
{
    const int c_first = 16, c_end = ORBX_XOFF + g.w + 2;                 // box columns; outputs start 4 further right
    const int r_first = ORBX_EDGE - 3, r_end = ORBX_EDGE + g.h + 2;      // first output row 16
    const int need = r_end - r_first;
    const int ny = (need + ORBX_BLUR_MAX_ROWS - 1) / ORBX_BLUR_MAX_ROWS;
    const int R = std::min(ORBX_BLUR_MAX_ROWS, (((need + ny - 1) / ny) + 1) & ~1);
    *rows_per_unit = R;
    for (int y = 0; y < ny; y++)
        for (int c0 = c_first; c0 + 4 < c_end; c0 += 64) {
            OrbxBlurUnit u;
            u.level = (short)level; u.c0 = (short)c0; u.r0 = (short)(r_first - 3 + y * R); u.rows = (short)R;
            out.push_back(u);
        }
}

void orbx_launch_blur(const OrbxFrameLayout& L, const OrbxTmaps& maps, const OrbxBlurUnit* d_units, int nunits, int nframes, cudaStream_t st)
{
    static OrbxSmemMark mk = {};
    const size_t smem = (size_t)BLUR_SLOT * 2 * BLUR_WARPS + 128;
    orbx_need_smem(blur_units_kernel, mk, smem);
    blur_units_kernel<<<dim3((nunits + 2 * BLUR_WARPS - 1) / (2 * BLUR_WARPS), nframes), BLUR_WARPS * 32, smem, st>>>(L, d_units, nunits, maps);
}

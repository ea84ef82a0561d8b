// orbx_fast.cu — per-cell FAST-9/16 with the iniThFAST -> minThFAST retry (replaces the cell loop of
// ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:849-914, and the cv::FAST(..., nonmaxSuppression=true)
// calls inside it).
//
// ONE WARP PER 30-px CELL, all levels and all frames of the batch in one launch; no block-level barrier anywhere.
// The warp stages the cell tile (+3-px ring) in shared memory with aligned 32-bit loads, then
//   1. quick test (four compass pixels) on every pixel; survivors are appended, in row-major order, to a per-warp
//      list with ballot + popc (dense work for the expensive steps, no divergence waste);
//   2. exact corner score of the listed pixels (one polarity per pixel, chosen from the compass pixels), written to a
//      zero-initialised u8 score map; the list is re-compacted to the true corners (score >= T):
//        score(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min d, -max d) - 1,  d_k = I(p) - I(ring_k)
//        p is a corner at threshold T  <=>  score(p) >= T      (cv::cornerScore<16>; independent of T)
//   3. cv::FAST's strict 3x3 non-max suppression over the list (still row-major), survivors written to the cell's
//      slot with ballot-prefix offsets — the reference's output order.
// The score map is zero outside the cell's own scored rectangle: in the reference every cell is a separate
// cv::FAST call on a sub-image, so NMS never sees a neighbouring cell. If step 3 leaves nothing at iniThFAST the
// cell is redone at minThFAST (ORBextractor.cc:894-900). The score map never touches HBM.
#include "orbx_internal.cuh"

#define FAST_MAX_WARPS 8     // warps (= cells) per CTA are chosen at launch: whatever packs most warps into an SM

struct FastSmemCfg { int tpw, th, sp, srows, list_cap, tile_off, score_off, list_off, per_warp; };

// max over the 16 arcs of 9 contiguous ring pixels of min(e[k..k+8]) with e_k = sg * (I(p) - I(ring_k)):
// sg = +1 scores arcs of DARKER pixels (A = max_k min d), sg = -1 arcs of BRIGHTER pixels (-B = max_k min(-d)).
// Two ring pixels ride in one register: P[k] = (e[k] + 1000) | (e[k+8] + 1000) << 16 (the bias keeps both halves
// positive, so each P[k] is two exact multiply-adds and needs no packing step), and the sliding-window minimum
// (log-step doubling: windows of 2, 4, then 4+4+4 overlapping = 9) runs on the packed 16-bit min/max instructions of
// sm_100a (VIMNMX.S16x2 / VIMNMX3.S16x2). Index k+8 of a packed array is the same register with its halves swapped.
// NOTE: an earlier scalar version that folded the dark and the bright result into a single running accumulator was
// mis-compiled by ptxas 12.9 for sm_100a (VIMNMX3 fusion) — tests/test_gpu_parity.py::test_stages_match_oracle pins
// the scores.
__device__ __forceinline__ unsigned swap16(const unsigned x) { return __byte_perm(x, 0, 0x1032); }
// returns max_k min(e[k..k+8]); when `other` is set also the opposite polarity, max_k min(-e[k..k+8]) = -min_k max(e[k..k+8]),
// from the same packed registers (sliding-window maximum), and the larger of the two
__device__ __forceinline__ int fast_arc_score(const uint8_t* __restrict__ c, const int tp, const int v, const int sg, const bool other)
{
    const unsigned mlo = (unsigned)(-sg), mhi = (unsigned)(-sg) << 16;
    const unsigned bias = (unsigned)(sg * v + 1000) * 0x10001u;
    unsigned P[9], L2[10], L4[13];
#define ORBX_PK(j, off) P[j] = (unsigned)c[off] * mlo + ((unsigned)c[-(off)] * mhi + bias);
    ORBX_PK(0, 3 * tp) ORBX_PK(1, 3 * tp + 1) ORBX_PK(2, 2 * tp + 2) ORBX_PK(3, tp + 3)
    ORBX_PK(4, 3) ORBX_PK(5, -tp + 3) ORBX_PK(6, -2 * tp + 2) ORBX_PK(7, -3 * tp + 1)
#undef ORBX_PK
    P[8] = swap16(P[0]);
#pragma unroll
    for (int k = 0; k < 8; k++) L2[k] = __vmins2(P[k], P[k + 1]);
    L2[8] = swap16(L2[0]); L2[9] = swap16(L2[1]);
#pragma unroll
    for (int k = 0; k < 8; k++) L4[k] = __vmins2(L2[k], L2[k + 2]);
#pragma unroll
    for (int k = 0; k < 5; k++) L4[8 + k] = swap16(L4[k]);
    unsigned lo9[8];
#pragma unroll
    for (int k = 0; k < 8; k++) lo9[k] = __vimin3_s16x2(L4[k], L4[k + 4], L4[k + 5]);   // e[k..k+3], e[k+4..k+7], e[k+5..k+8]
    const unsigned m1 = __vimax3_s16x2(lo9[0], lo9[1], lo9[2]), m2 = __vimax3_s16x2(lo9[3], lo9[4], lo9[5]);
    const unsigned m = __vmaxs2(__vimax3_s16x2(lo9[6], lo9[7], m1), m2);
    int best = max((int)(m & 0xffffu), (int)(m >> 16)) - 1000;
    if (other) {
        unsigned H2[10], H4[13], hi9[8];
#pragma unroll
        for (int k = 0; k < 8; k++) H2[k] = __vmaxs2(P[k], P[k + 1]);
        H2[8] = swap16(H2[0]); H2[9] = swap16(H2[1]);
#pragma unroll
        for (int k = 0; k < 8; k++) H4[k] = __vmaxs2(H2[k], H2[k + 2]);
#pragma unroll
        for (int k = 0; k < 5; k++) H4[8 + k] = swap16(H4[k]);
#pragma unroll
        for (int k = 0; k < 8; k++) hi9[k] = __vimax3_s16x2(H4[k], H4[k + 4], H4[k + 5]);
        const unsigned n1 = __vimin3_s16x2(hi9[0], hi9[1], hi9[2]), n2 = __vimin3_s16x2(hi9[3], hi9[4], hi9[5]);
        const unsigned n = __vmins2(__vimin3_s16x2(hi9[6], hi9[7], n1), n2);
        best = max(best, 1000 - min((int)(n & 0xffffu), (int)(n >> 16)));
    }
    return best;
}

// exact corner score of a pixel that passed the quick test at threshold T; 0 when it is not a corner at T:
//   score(p) = max(A, -B) - 1,   p is a corner at T  <=>  score(p) >= T      (cv::cornerScore<16>; independent of T)
// A score >= T needs an arc whose nine pixels are all beyond the threshold, hence two ADJACENT compass pixels (ring
// 0, 4, 8, 12) beyond it on that side: only a polarity with such a pair can reach T, so the second polarity is scanned
// only for pixels that have a dark and a bright compass pair (edge-like pixels).
__device__ __forceinline__ int fast_score_T(const uint8_t* __restrict__ c, const int tp, const int T)
{
    const int v = c[0];
    const int d0 = v - c[3 * tp], d4 = v - c[3], d8 = v - c[-3 * tp], d12 = v - c[-3];
    const int dp = ((T - d4) | (T - d12)) & ((T - d0) | (T - d8));     // sign set <=> a dark arc is possible at T
    const int bp = ((d4 + T) | (d12 + T)) & ((d0 + T) | (d8 + T));     // sign set <=> a bright arc is possible at T
    const int s = fast_arc_score(c, tp, v, dp < 0 ? 1 : -1, (dp & bp) < 0) - 1;
    return s >= T ? s : 0;
}

// Quick test of one pixel, result in the SIGN BIT: set iff two adjacent compass pixels (ring 0, 4, 8, 12) are both darker
// than v - T or both brighter than v + T. Both compares of a ring pixel r ride in one multiply-add:
//   X = r * 0xFFFF0001 + C,  C = v * 0xFFFF + KT,  KT = (T + 0x8000) << 16 | (0x8000 + T)
// leaves r - (v - T) + 0x8000 in the low half and (v + T) - r + 0x8000 in the high half (neither half can carry into
// the other), so bit 15 is CLEAR iff r is dark and bit 31 is CLEAR iff r is bright. With Y = (X4 & X12) | (X0 & X8),
// bit 15 of Y is clear iff (D4|D12)&(D0|D8) == (D0&D4)|(D4&D8)|(D8&D12)|(D12&D0), bit 31 likewise for bright.
__device__ __forceinline__ unsigned quick_test(const uint8_t* __restrict__ qp, const int tp, const unsigned KT)
{
    const unsigned C = (unsigned)qp[0] * 0xFFFFu + KT;
    const unsigned X0 = (unsigned)qp[3 * tp] * 0xFFFF0001u + C, X4 = (unsigned)qp[3] * 0xFFFF0001u + C;
    const unsigned X8 = (unsigned)qp[-3 * tp] * 0xFFFF0001u + C, X12 = (unsigned)qp[-3] * 0xFFFF0001u + C;
    const unsigned Y = (X4 & X12) | (X0 & X8);
    return ~(Y & (Y << 16));
}

// TPC / SPC: compile-time tile / score-map pitches in bytes (all ring and NMS offsets become immediates);
// 0 = take them from cfg (cells wider than the common 30..46 px)
template <int TPC, int SPC>
// (288 threads x 4 CTAs as the bound: at most 56 registers per thread = 7 allocation units per warp, 36 warps per SM by registers)
__global__ void __launch_bounds__(288, 4) fast_cells_kernel(OrbxFrameLayout L, FastSmemCfg cfg)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cell_id = blockIdx.x * (blockDim.x >> 5) + wid;
    const int frame = blockIdx.y;
    if (cell_id >= L.ncells) return;                      // whole warp
    const OrbxCell c = L.cells[cell_id];
    int* cell_count = L.cell_count + (size_t)frame * L.ncells + cell_id;
    const int ew = c.ex1 - c.ex0, eh = c.ey1 - c.ey0;
    if (ew <= 0 || eh <= 0) { if (lane == 0) *cell_count = 0; return; }
    const OrbxLevelGeom g = L.lvl[c.level];

    uint8_t* wbase = smem + (size_t)wid * cfg.per_warp;
    uint32_t* tile32 = reinterpret_cast<uint32_t*>(wbase + cfg.tile_off);
    uint8_t* score = wbase + cfg.score_off;
    unsigned short* list = reinterpret_cast<unsigned short*>(wbase + cfg.list_off);
    const int tp = TPC ? TPC : cfg.tpw * 4, sp = SPC ? SPC : cfg.sp, tpw = tp >> 2;

    // ---- stage the tile rows [ey0-3, ey1+3) x [ex0-3, ex1+3) with aligned 32-bit loads
    const int tw = ew + 6, th = eh + 6;
    const uint8_t* p0 = L.raw + (size_t)frame * L.frame_raw_bytes + g.raw_off +
                        (size_t)(c.ey0 - 3 + ORBX_EDGE) * g.pitch + (c.ex0 - 3 + ORBX_XOFF);
    const int sh = (int)(reinterpret_cast<uintptr_t>(p0) & 3);
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(p0 - sh);
    const int nw = (sh + tw + 3) >> 2;                    // words per row
    const int pitch_w = g.pitch >> 2;
    {
        // flattened (row, word) index, four independent loads in flight per lane; i / nw by reciprocal multiply
        // (exact for i < 2048, nw < 40 with a 20-bit reciprocal: checked exhaustively on the host)
        const int nwords = th * nw, inv = (1 << 20) / nw + 1;   // th*nw < 2048
        for (int i0 = lane; i0 < nwords; i0 += 128) {
            uint32_t v[4]; int dst[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + 32 * u;
                const int r = (int)(((unsigned)i * (unsigned)inv) >> 20), w = i - r * nw;
                dst[u] = r * tpw + w;
                v[u] = i < nwords ? __ldg(pa + r * pitch_w + w) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) if (i0 + 32 * u < nwords) tile32[dst[u]] = v[u];
        }
    }
    const uint8_t* tile = reinterpret_cast<const uint8_t*>(tile32) + sh;   // tile[ty*tp + tx]

    uint32_t* slot = L.slots + (size_t)frame * L.slot_total + c.slot_off;
    int total = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int T = pass ? L.min_th : L.ini_th;
        // zero the score map (frame included)
        {
            uint32_t* s32 = reinterpret_cast<uint32_t*>(score);
            const int nwords = ((eh + 2) * sp) >> 2;
            for (int i = lane; i < nwords; i += 32) s32[i] = 0;
        }
        __syncwarp();
        // 1. quick test on every pixel. Every arc of 9 contiguous ring pixels holds two ADJACENT compass pixels
        //    (ring 0,4,8,12), so a dark (bright) arc needs an adjacent compass pair that is dark (bright).
        //    Row-major list entry = py<<7 | px.
        int cnt = 0;
        const unsigned lt_mask = (1u << lane) - 1;
        {
            // lane = ROW, pixels walked left to right; the pass bit of every pixel is shifted into the lane's row mask
            // (funnel shift pulls the sign bit of quick_test() in: one instruction, no ballot in the arithmetic loop).
            // A lane's mask IS its row of the row-major list, so no transposition is needed: row offsets come from one
            // warp scan, then every lane writes the entries of its own row. The tile pitch is an odd number of words, so
            // the 32 rows a warp touches per load sit in 32 different banks.
            const unsigned KT = ((unsigned)(T + 0x8000) << 16) + (unsigned)(0x8000 + T);
            const int nlo = min(ew, 32), nhi = ew - nlo;         // up to two 32-column chunks (cells are < 60 px wide)
            for (int rbase = 0; rbase < eh; rbase += 32) {       // and two 32-row halves (< 60 px tall)
                const int py = rbase + lane;
                const uint8_t* row = tile + (min(py, eh - 1) + 3) * tp + 3;   // clamped: loads stay inside the tile
                unsigned acc0 = 0, acc1 = 0;
#pragma unroll 4
                for (int px = 0; px < nlo; px++) acc0 = __funnelshift_l(quick_test(row + px, tp, KT), acc0, 1);
#pragma unroll 4
                for (int px = 32; px < ew; px++) acc1 = __funnelshift_l(quick_test(row + px, tp, KT), acc1, 1);
                // MSB-first accumulation: pixel px of an n-pixel chunk sits at bit n-1-px -> bit px after reversal
                const unsigned rm0 = py < eh ? __brev(acc0) >> (32 - nlo) : 0u;
                const unsigned rm1 = (py < eh && nhi > 0) ? __brev(acc1) >> (32 - nhi) : 0u;
                const int c = __popc(rm0) + __popc(rm1);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                int off = cnt + incl - c;
                const int ebase = py * tp;                       // list entry = py * tp + px: offset into the tile AND the score map
                // (the list holds one slot per pixel of the largest cell, so it cannot overflow)
                for (unsigned m = rm0; m; m &= m - 1) list[off++] = (unsigned short)(ebase + (__ffs(m) - 1));
                for (unsigned m = rm1; m; m &= m - 1) list[off++] = (unsigned short)(ebase + (32 + __ffs(m) - 1));
                cnt += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        cnt = min(cnt, cfg.list_cap);
        __syncwarp();
        // 2. exact score of the listed pixels, written to the score map; the list is re-compacted in place to the true
        //    corners (score >= T). In-place is safe: a chunk of 32 entries is read before it is written and the write
        //    position never passes the read position.
        int ncorner = 0;
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            int e = 0, sc = 0;
            if (k < cnt) {
                e = list[k];
                sc = fast_score_T(tile + e + (3 * tp + 3), tp, T);
            }
            const unsigned m = __ballot_sync(0xffffffffu, sc != 0);
            __syncwarp();
            if (sc) {
                list[ncorner + __popc(m & lt_mask)] = (unsigned short)e;
                score[e + (sp + 1)] = (uint8_t)sc;           // sp == tp
            }
            ncorner += __popc(m);
        }
        __syncwarp();
        // 4. strict 3x3 NMS over the corners (row-major) + ordered write
        for (int k0 = 0; k0 < ncorner; k0 += 32) {
            const int k = k0 + lane;
            int keep = 0, s = 0, e = 0;
            if (k < ncorner) {
                e = list[k];
                const uint8_t* q = score + e + (sp + 1);
                s = q[0];
                // branch-free: strictly greater than the largest of the eight neighbours (list entries have s >= T > 0)
                const int nmax = max(max(max(max((int)q[-1], (int)q[1]), (int)q[-sp - 1]), max((int)q[-sp], (int)q[-sp + 1])),
                                     max(max((int)q[sp - 1], (int)q[sp]), (int)q[sp + 1]));
                keep = s > nmax;
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int off = total + __popc(m & ((1u << lane) - 1));
                const int py = e / tp, px = e - py * tp;        // survivors only
                if (off < c.slot_cap)
                    slot[off] = ((uint32_t)s << 24) | ((uint32_t)(c.ey0 + py - ORBX_MINB) << 12) | (uint32_t)(c.ex0 + px - ORBX_MINB);
            }
            total += __popc(m);
        }
        if (total > 0) break;
        __syncwarp();
    }
    if (lane == 0) *cell_count = total < c.slot_cap ? total : c.slot_cap;
}

void orbx_launch_fast(const OrbxFrameLayout& L, int max_tile_w, int max_tile_h, int nframes, cudaStream_t st)
{
    FastSmemCfg cfg;
    // compile-time pitches (11, 13 or 15 words: odd, see the quick-test loop) for the tile and, so that one list entry
    // addresses both, for the score map; the narrowest that holds the widest cell (+3-px ring, +3 bytes of alignment
    // slack) keeps shared memory per warp — and with it the number of resident warps — as good as it gets
    const int need = max_tile_w + 3;
    const int variant = need <= 44 ? 0 : need <= 52 ? 1 : need <= 60 ? 2 : 3;
    cfg.tpw = variant == 0 ? 11 : variant == 1 ? 13 : variant == 2 ? 15 : ((3 + max_tile_w + 3) / 4 + 1) | 1;
    cfg.th = max_tile_h;
    cfg.sp = cfg.tpw * 4;                               // score pitch == tile pitch
    cfg.srows = max_tile_h - 6 + 2;
    cfg.list_cap = (max_tile_w - 6) * (max_tile_h - 6);
    cfg.tile_off = 0;
    cfg.score_off = (cfg.tpw * 4 * cfg.th + 15) & ~15;
    cfg.list_off = (cfg.score_off + cfg.sp * cfg.srows + 15) & ~15;
    cfg.per_warp = (cfg.list_off + 2 * cfg.list_cap + 15) & ~15;
    typedef void (*kern_t)(OrbxFrameLayout, FastSmemCfg);
    static const kern_t kerns[4] = {fast_cells_kernel<44, 44>, fast_cells_kernel<52, 52>, fast_cells_kernel<60, 60>, fast_cells_kernel<0, 0>};
    const kern_t kern = kerns[variant];
    // The kernel is issue-bound and gains from every extra resident warp; shared memory per warp (tile + score map + list)
    // decides how many fit, and the CTA size decides how well they pack: ask the occupancy calculator for each size.
    static int best_fw[64][4] = {}; static int best_pw[64][4] = {};
    static OrbxSmemMark mk[4] = {};
    int dev = 0; cudaGetDevice(&dev); dev &= 63;
    if (best_pw[dev][variant] != cfg.per_warp) {
        int bw = 0, bfw = 4;
        for (int fw = FAST_MAX_WARPS; fw >= 2; fw--) {
            const size_t sm = (size_t)cfg.per_warp * fw;
            orbx_need_smem(kern, mk[variant], sm);
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, fw * 32, sm) == cudaSuccess && nb * fw > bw) { bw = nb * fw; bfw = fw; }
        }
        cudaGetLastError();
        best_fw[dev][variant] = bfw; best_pw[dev][variant] = cfg.per_warp;
    }
    const int fwarps = best_fw[dev][variant];
    const size_t smem = (size_t)cfg.per_warp * fwarps;
    dim3 grid((L.ncells + fwarps - 1) / fwarps, nframes);
    kern<<<grid, fwarps * 32, smem, st>>>(L, cfg);
}

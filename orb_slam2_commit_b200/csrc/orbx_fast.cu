// orbx_fast.cu — per-cell FAST-9/16 with the iniThFAST -> minThFAST retry (replaces the cell loop of
// ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:849-914, and the cv::FAST(..., nonmaxSuppression=true)
// calls inside it).
//
// ONE WARP PER 30-px CELL, all levels and all frames of the batch in one launch; no block-level barrier anywhere.
// One elected lane has the TMA unit copy the cell tile (+3-px ring) from the HBM pyramid into the warp's shared-memory
// slot (cp.async.bulk.tensor through the level's tensor map, completion on the warp's mbarrier: orbx_tma.cuh); while the
// copy is in flight the warp clears its score map. Then
//   1. quick test (four compass pixels) on every pixel, LANE = COLUMN, rows walked top to bottom (conflict-free for any
//      tile pitch, so the pitch is simply the TMA box width: 64 bytes, or 80 for cells wider than 43 px — the box must start
//      at a 16-byte aligned column of the level buffer, so it is up to 15 columns wider than the tile); survivors are appended, in row-major order, to a per-warp
//      list with ballot + popc, each entry carrying which polarity (dark / bright arc) its compass pixels allow;
//   2. exact corner score of the listed pixels (the polarity the entry names; both only for the few edge-like pixels
//      that allow both), written to a zero-initialised u8 score map; the list is re-compacted to the true corners
//      (score >= T):
//        score(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min d, -max d) - 1,  d_k = I(p) - I(ring_k)
//        p is a corner at threshold T  <=>  score(p) >= T      (cv::cornerScore<16>; independent of T)
//   3. cv::FAST's strict 3x3 non-max suppression over the list (still row-major), survivors written to the cell's
//      slot with ballot-prefix offsets — the reference's output order.
// The score map is zero outside the cell's own scored rectangle: in the reference every cell is a separate
// cv::FAST call on a sub-image, so NMS never sees a neighbouring cell. If step 3 leaves nothing at iniThFAST the
// cell is redone at minThFAST (ORBextractor.cc:894-900). The score map never touches HBM.
#include "orbx_internal.cuh"
#include "orbx_tma.cuh"
#include <cstdlib>

#define FAST_MAX_WARPS 8     // warps (= cells) per CTA are chosen at launch: whatever packs most warps into an SM

struct FastSmemCfg { int list_cap, score_off, score_bytes, list_off, bar_off, per_warp; };

// max over the 16 arcs of 9 contiguous ring pixels of min(e[k..k+8]) with e_k = sg * (I(p) - I(ring_k)):
// sg = +1 scores arcs of DARKER pixels (A = max_k min d), sg = -1 arcs of BRIGHTER pixels (-B = max_k min(-d)).
// Two ring pixels ride in one register: P[k] = (e[k] + 1000) | (e[k+8] + 1000) << 16 (the bias keeps both halves
// positive, so each P[k] is two exact multiply-adds and needs no packing step), and the sliding-window minimum
// (windows of 3, then 3+3+3 = 9) runs on the packed 16-bit 3-input min/max instructions of sm_100a (VIMNMX3.S16x2).
// Index k+8 of a packed array is the same register with its halves swapped.
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
    unsigned P[10], W3[14], W9[8];
#define ORBX_PK(j, off) P[j] = (unsigned)c[off] * mlo + ((unsigned)c[-(off)] * mhi + bias);
    ORBX_PK(0, 3 * tp) ORBX_PK(1, 3 * tp + 1) ORBX_PK(2, 2 * tp + 2) ORBX_PK(3, tp + 3)
    ORBX_PK(4, 3) ORBX_PK(5, -tp + 3) ORBX_PK(6, -2 * tp + 2) ORBX_PK(7, -3 * tp + 1)
#undef ORBX_PK
    P[8] = swap16(P[0]); P[9] = swap16(P[1]);
    // a window of nine = three windows of three: e[k..k+2], e[k+3..k+5], e[k+6..k+8] (3-input packed min, VIMNMX3.S16x2)
#pragma unroll
    for (int k = 0; k < 8; k++) W3[k] = __vimin3_s16x2(P[k], P[k + 1], P[k + 2]);
#pragma unroll
    for (int k = 0; k < 6; k++) W3[8 + k] = swap16(W3[k]);
#pragma unroll
    for (int k = 0; k < 8; k++) W9[k] = __vimin3_s16x2(W3[k], W3[k + 3], W3[k + 6]);
    const unsigned m1 = __vimax3_s16x2(W9[0], W9[1], W9[2]), m2 = __vimax3_s16x2(W9[3], W9[4], W9[5]);
    const unsigned m = __vmaxs2(__vimax3_s16x2(W9[6], W9[7], m1), m2);
    int best = max((int)(m & 0xffffu), (int)(m >> 16)) - 1000;
    if (other) {
        unsigned H3[14], H9[8];
#pragma unroll
        for (int k = 0; k < 8; k++) H3[k] = __vimax3_s16x2(P[k], P[k + 1], P[k + 2]);
#pragma unroll
        for (int k = 0; k < 6; k++) H3[8 + k] = swap16(H3[k]);
#pragma unroll
        for (int k = 0; k < 8; k++) H9[k] = __vimax3_s16x2(H3[k], H3[k + 3], H3[k + 6]);
        const unsigned n1 = __vimin3_s16x2(H9[0], H9[1], H9[2]), n2 = __vimin3_s16x2(H9[3], H9[4], H9[5]);
        const unsigned n = __vmins2(__vimin3_s16x2(H9[6], H9[7], n1), n2);
        best = max(best, 1000 - min((int)(n & 0xffffu), (int)(n >> 16)));
    }
    return best;
}

// Quick test of one pixel: both compares of a ring pixel r ride in one multiply-add,
//   X = r * 0xFFFF0001 + C,  C = v * 0xFFFF + KT,  KT = (T + 0x4000) << 16 | (0x4000 + T)
// leaves r - (v - T) + 0x4000 in the low half and (v + T) - r + 0x4000 in the high half (neither half can carry into
// the other), so bit 14 is CLEAR iff r is dark and bit 30 is CLEAR iff r is bright. With Y = (X4 & X12) | (X0 & X8),
// bit 14 of Y is clear iff (D4|D12)&(D0|D8) == (D0&D4)|(D4&D8)|(D8&D12)|(D12&D0): two ADJACENT compass pixels (ring 0, 4,
// 8, 12) are darker than v - T — every arc of 9 contiguous ring pixels holds such a pair; bit 30 likewise for bright.
// Returns ~Y: bit 14 SET <=> a dark arc is possible at T, bit 30 SET <=> a bright arc is possible at T.
// (v, up, down) are the lane's own column at rows y, y-3, y+3 and come from the caller's register window.
__device__ __forceinline__ unsigned quick_test(const unsigned v, const unsigned up, const unsigned down, const unsigned left,
                                               const unsigned right, const unsigned KT)
{
    const unsigned C = v * 0xFFFFu + KT;
    const unsigned X0 = down * 0xFFFF0001u + C, X4 = right * 0xFFFF0001u + C;
    const unsigned X8 = up * 0xFFFF0001u + C, X12 = left * 0xFFFF0001u + C;
    return ~((X4 & X12) | (X0 & X8));
}

// predicated shared-memory stores (one instruction, no branch, no dummy-slot select)
__device__ __forceinline__ void sts_u16_if(const unsigned p, const void* addr, const unsigned v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.u16 [%1], %2; }" ::"r"(p), "r"(orbx_smem_addr(addr)), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts_u8_if(const unsigned p, const void* addr, const unsigned v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.u8 [%1], %2; }" ::"r"(p), "r"(orbx_smem_addr(addr)), "r"(v) : "memory");
}

// One row of the quick-test append: nY = ~Y & mask (the lanes that pass, with their polarity bits), ballot, prefix count,
// predicated 16-bit store of `base + nY + (nY >> 15)` (entry | polarity flags in bits 14 / 15) at list_addr + 2 * prefix.
// Returns the ballot. One predicate serves the vote and the store.
__device__ __forceinline__ unsigned fast_append_row(const unsigned Y, const unsigned mask, const unsigned lt_mask, const uint32_t list_addr,
                                                    const unsigned base)
{
    unsigned m;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .b32 ny, t, a, v;\n"
        "lop3.b32 ny, %1, %2, 0, 0x0c;\n"            // ~Y & mask
        "setp.ne.u32 q, ny, 0;\n"
        "vote.sync.ballot.b32 %0, q, 0xffffffff;\n"
        "and.b32 t, %0, %3;\n"
        "popc.b32 t, t;\n"
        "mad.lo.u32 a, t, 2, %4;\n"
        "shr.u32 v, ny, 15;\n"
        "add.u32 v, v, ny;\n"
        "add.u32 v, v, %5;\n"
        "@q st.shared.u16 [a], v;\n"
        "}\n" : "=r"(m) : "r"(Y), "r"(mask), "r"(lt_mask), "r"(list_addr), "r"(base) : "memory");
    return m;
}

// The corner append of the score phase in one block: predicate (k < cnt && s >= T), ballot, prefix count, predicated stores of
// the list entry and of the score byte. The vote is also the warp-level ordering the in-place compaction needs (every lane has
// consumed its own entry before any lane overwrites one).
__device__ __forceinline__ unsigned fast_append_corner(const int s, const int T, const int k, const int cnt, const unsigned lt_mask,
                                                       const uint32_t list_addr, const unsigned e, const uint32_t score_addr)
{
    unsigned m;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .b32 t, a;\n"
        "setp.ge.s32 q, %1, %2;\n"
        "setp.lt.and.s32 q, %3, %4, q;\n"
        "vote.sync.ballot.b32 %0, q, 0xffffffff;\n"
        "and.b32 t, %0, %5;\n"
        "popc.b32 t, t;\n"
        "mad.lo.u32 a, t, 2, %6;\n"
        "@q st.shared.u16 [a], %7;\n"
        "@q st.shared.u8 [%8], %1;\n"
        "}\n" : "=r"(m) : "r"(s), "r"(T), "r"(k), "r"(cnt), "r"(lt_mask), "r"(list_addr), "h"((unsigned short)e), "r"(score_addr) : "memory");
    return m;
}

#define FAST_DARK 0x4000u      // list entry flags (bits 14 / 15); bits 0..13 = py * TP + px
#define FAST_BRIGHT 0x8000u
#define FAST_QMASK 0x40004000u // the two result bits of quick_test()

// TP: tile pitch = TMA box width in bytes (64 or 80). The TMA unit wants the box to start at a 16-byte aligned byte
// column of the level buffer (measured: any other origin raises an illegal-instruction fault, tools/probe/), so the box
// starts up to 15 columns left of the tile and `sh` shifts the tile's base pointer; all ring offsets stay immediates.
// SP: score-map pitch (>= widest cell + 2; separate from TP so that the u8 score map stays small).
// (288 threads x 4 CTAs as the bound: at most 56 registers per thread = 7 allocation units per warp, 36 warps per SM by registers)
template <int TP, int SP>
__global__ void __launch_bounds__(288, 4) fast_cells_kernel(OrbxFrameLayout L, FastSmemCfg cfg, const __grid_constant__ OrbxTmaps maps)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cell_id = blockIdx.x * (blockDim.x >> 5) + wid;
    const int frame = blockIdx.y;
    if (cell_id >= L.ncells) return;                      // whole warp
    const OrbxCell c = L.cells[cell_id];
    int* cell_count = L.cell_count + (size_t)frame * L.ncells + cell_id;
    const int ew = c.ex1 - c.ex0, eh = c.ey1 - c.ey0;
    if (ew <= 0 || eh <= 0) { if (lane == 0) *cell_count = 0; return; }

    // the dynamic shared-memory window starts 128-byte aligned only by convention: align it here (the launch adds 128 B)
    uint8_t* wbase = smem + ((128u - (orbx_smem_addr(smem) & 127u)) & 127u) + (size_t)wid * cfg.per_warp;
    uint8_t* score = wbase + cfg.score_off;
    unsigned short* list = reinterpret_cast<unsigned short*>(wbase + cfg.list_off);
    const uint32_t bar = orbx_smem_addr(wbase + cfg.bar_off);

    // ---- stage the tile rows [ey0-3, ey1+3) x [ex0-3, ex1+3): one bulk tensor copy of the box that starts at the 16-byte
    // aligned column at or left of the tile; rows / columns beyond the level buffer are zero-filled by the TMA unit and
    // never read for an emitted pixel
    const int x0 = c.ex0 - 3 + ORBX_XOFF;                 // byte column of the tile's first pixel in the level buffer
    const uint8_t* tile = wbase + (x0 & 15);              // tile[ty * TP + tx], (tx, ty) = (0, 0) at payload (ex0 - 3, ey0 - 3)
    if (lane == 0) {
        orbx_mbar_init(bar, 1);
        orbx_mbar_expect_tx(bar, (uint32_t)(TP * c.box_h));
        orbx_tma_load_3d(orbx_smem_addr(wbase), &maps.m[c.level], x0 & ~15, c.ey0 - 3 + ORBX_EDGE, L.frame0 + frame, bar);
    }
    // meanwhile: zero the score map (frame included). It stays all-zero if the first pass finds no corner, so the
    // second pass needs no second clearing.
    {
        uint4* s128 = reinterpret_cast<uint4*>(score);
        const int n16 = cfg.score_bytes >> 4;
        for (int i = lane; i < n16; i += 32) s128[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();                                         // the barrier is initialised before anyone waits on it
    orbx_mbar_wait(bar, 0);

    uint32_t* slot = L.slots + (size_t)frame * L.slot_total + c.slot_off;
    const unsigned lt_mask = (1u << lane) - 1;
    int total = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int T = pass ? L.min_th : L.ini_th;
        // 1. quick test on every pixel, lane = column: the 32 lanes read 32 consecutive bytes of a tile row per load, and
        //    a lane keeps its own column in a six-row register window (row y+3 is loaded once and serves as "down", three
        //    rows later as centre, six rows later as "up"). Row-major list entry = py * TP + px | polarity flags; lanes
        //    that do not pass store to a dummy slot behind the list, so the append is branch-free.
        int cnt = 0;
        {
            const unsigned KT = ((unsigned)(T + 0x4000) << 16) + (unsigned)(0x4000 + T);
            if (ew <= 32) {
                const unsigned lanemask = lane < ew ? FAST_QMASK : 0u;
                const uint8_t* q = tile + 3 + min(lane, ew - 1);           // clamped: loads stay inside the tile; q[ty * TP] = own column
                unsigned w0 = q[0], w1 = q[TP], w2 = q[2 * TP], w3 = q[3 * TP], w4 = q[4 * TP], w5 = q[5 * TP];   // tile rows py .. py+5
                q += 3 * TP;                                                // q[0] = centre pixel of row py
                unsigned e = lane;
                uint32_t la = orbx_smem_addr(list);                         // address of list[cnt]
#pragma unroll 6
                for (int py = 0; py < eh; py++) {
                    const unsigned w6 = q[3 * TP];
                    const unsigned Y = ~quick_test(w3, w0, w6, q[-3], q[3], KT);
                    const unsigned m = fast_append_row(Y, lanemask, lt_mask, la, e);
                    la += 2 * __popc(m);
                    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6;
                    q += TP; e += TP;
                }
                cnt = (int)(la - orbx_smem_addr(list)) >> 1;
            } else {                                                        // cells wider than 32 px (< 64): two column chunks per row
                const unsigned lanemask1 = lane + 32 < ew ? FAST_QMASK : 0u;
                const uint8_t* q0 = tile + 3 * TP + 3 + lane;
                const uint8_t* q1 = tile + 3 * TP + 3 + min(lane + 32, ew - 1);
                int e = lane;
#pragma unroll 2
                for (int py = 0; py < eh; py++) {
                    const unsigned nY0 = quick_test(q0[0], q0[-3 * TP], q0[3 * TP], q0[-3], q0[3], KT) & FAST_QMASK;
                    const unsigned nY1 = quick_test(q1[0], q1[-3 * TP], q1[3 * TP], q1[-3], q1[3], KT) & lanemask1;
                    const unsigned m0 = __ballot_sync(0xffffffffu, nY0 != 0), m1 = __ballot_sync(0xffffffffu, nY1 != 0);
                    sts_u16_if(nY0, list + cnt + __popc(m0 & lt_mask), (unsigned)e + ((nY0 | (nY0 >> 15)) & (FAST_DARK | FAST_BRIGHT)));
                    cnt += __popc(m0);
                    sts_u16_if(nY1, list + cnt + __popc(m1 & lt_mask), (unsigned)(e + 32) + ((nY1 | (nY1 >> 15)) & (FAST_DARK | FAST_BRIGHT)));
                    cnt += __popc(m1);
                    q0 += TP; q1 += TP; e += TP;
                }
            }
        }
        __syncwarp();
        // 2. exact score of the listed pixels, written to the score map; the list is re-compacted in place to the true
        //    corners (score >= T). In-place is safe: a chunk of 32 entries is read before it is written and the write
        //    position never passes the read position.
        //      score(p) = max(A, -B) - 1,   p is a corner at T  <=>  score(p) >= T      (cv::cornerScore<16>)
        //    A score >= T needs an arc whose nine pixels are all beyond the threshold, hence two ADJACENT compass pixels
        //    beyond it on that side: only a polarity the entry's flags name can reach T.
        int ncorner = 0;
        uint32_t lc = orbx_smem_addr(list);                              // address of list[ncorner]
        const uint32_t score_base = orbx_smem_addr(score) + (SP + 1);
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            const unsigned ev = k < cnt ? list[k] : FAST_DARK;           // idle lanes score tile pixel (0,0): valid memory, result dropped
            const int e = ev & 0x3fff;
            const uint8_t* cp = tile + e + (3 * TP + 3);
            const int s = fast_arc_score(cp, TP, cp[0], (ev & FAST_DARK) ? 1 : -1, (ev & (FAST_DARK | FAST_BRIGHT)) == (FAST_DARK | FAST_BRIGHT)) - 1;
            const unsigned m = fast_append_corner(s, T, k, cnt, lt_mask, lc, (unsigned)e,
                                                  score_base + (unsigned)(e - (e / TP) * (TP - SP)));   // (py, px) -> py * SP + px
            lc += 2 * __popc(m);
        }
        ncorner = (int)(lc - orbx_smem_addr(list)) >> 1;
        __syncwarp();
        // 3. strict 3x3 NMS over the corners (row-major) + ordered write
        for (int k0 = 0; k0 < ncorner; k0 += 32) {
            const int k = k0 + lane;
            // branch-free: lanes past the end redo the last corner and drop the result
            const int e = list[min(k, ncorner - 1)];
            const int py = e / TP, px = e - py * TP;
            const uint8_t* q = score + py * SP + px + (SP + 1);
            const int s = q[0];
            // strictly greater than the largest of the eight neighbours (list entries have s >= T > 0)
            const int nmax = max(max(max(max((int)q[-1], (int)q[1]), (int)q[-SP - 1]), max((int)q[-SP], (int)q[-SP + 1])),
                                 max(max((int)q[SP - 1], (int)q[SP]), (int)q[SP + 1]));
            const bool keep = k < ncorner && s > nmax;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int off = total + __popc(m & lt_mask);
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

// box width: the tile (widest cell + 6) plus up to 15 columns of alignment slack
int orbx_fast_tile_pitch(int max_tile_w) { return (max_tile_w + 15 <= 64 && !getenv("ORBX_FAST_TP80")) ? 64 : 80; }

void orbx_launch_fast(const OrbxFrameLayout& L, const OrbxTmaps& maps, int max_tile_w, int max_tile_h, int nframes, cudaStream_t st)
{
    FastSmemCfg cfg;
    const int TP = orbx_fast_tile_pitch(max_tile_w);
    const int ew = max_tile_w - 6, eh = max_tile_h - 6;
    const int SP = TP == 80 ? 80 : ew + 2 <= 40 ? 40 : ew + 2 <= 48 ? 48 : 64;
    const int variant = TP == 80 ? 3 : SP == 40 ? 0 : SP == 48 ? 1 : 2;
    cfg.list_cap = ew * eh;
    cfg.score_off = (TP * max_tile_h + 127) & ~127;                         // tile: TP x (tallest box), 128-byte aligned for the TMA unit
    cfg.score_bytes = (SP * (eh + 2) + 16 + 15) & ~15;                      // (eh + 2) rows + the (SP + 1) origin shift + a spare byte
    cfg.list_off = cfg.score_off + cfg.score_bytes;
    cfg.bar_off = (cfg.list_off + 2 * (cfg.list_cap + 1) + 15) & ~15;    // + the dummy slot of the branch-free appends
    cfg.per_warp = (cfg.bar_off + 16 + 127) & ~127;
    typedef void (*kern_t)(OrbxFrameLayout, FastSmemCfg, const OrbxTmaps);
    static const kern_t kerns[4] = {fast_cells_kernel<64, 40>, fast_cells_kernel<64, 48>, fast_cells_kernel<64, 64>, fast_cells_kernel<80, 80>};
    const kern_t kern = kerns[variant];
    // The kernel is issue-bound and gains from every extra resident warp; shared memory per warp (tile + score map + list)
    // decides how many fit, and the CTA size decides how well they pack: ask the occupancy calculator for each size.
    static int best_fw[64][4] = {}; static int best_pw[64][4] = {};
    static OrbxSmemMark mk[4] = {};
    int dev = 0; cudaGetDevice(&dev); dev &= 63;
    if (best_pw[dev][variant] != cfg.per_warp) {
        int bw = 0, bfw = 4;
        const char* fwe = getenv("ORBX_FAST_WARPS");
        for (int fw = fwe ? atoi(fwe) : FAST_MAX_WARPS; fw >= (fwe ? atoi(fwe) : 2); fw--) {
            const size_t sm = (size_t)cfg.per_warp * fw + 128;
            orbx_need_smem(kern, mk[variant], sm);
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, fw * 32, sm) == cudaSuccess && nb * fw > bw) { bw = nb * fw; bfw = fw; }
        }
        cudaGetLastError();
        best_fw[dev][variant] = bfw; best_pw[dev][variant] = cfg.per_warp;
    }
    const int fwarps = best_fw[dev][variant];
    const size_t smem = (size_t)cfg.per_warp * fwarps + 128;
    dim3 grid((L.ncells + fwarps - 1) / fwarps, nframes);
    kern<<<grid, fwarps * 32, smem, st>>>(L, cfg, maps);
}

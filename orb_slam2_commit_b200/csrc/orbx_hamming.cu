// orbx_hamming.cu — 256-bit Hamming distance + best / second-best search (replaces
// ORBmatcher::DescriptorDistance, ORBmatcher.cc:1844-1860, and the top-2 loop idiom around it, :84-126;
// the row-band best-match stage of Frame::ComputeStereoMatches, Frame.cc:554-663).
//
// Brute force (BASELINE config 4): every thread keeps HT_QPT queries in registers (8 x u32 each), the CTA streams
// its slice of the train set through shared memory with 128-bit loads that are broadcast to the whole warp, and the
// inner loop is XOR + __popc + add with a branch-free top-2 update on packed keys (dist << 22 | local index):
//   t = max(key, best); best = min(best, key); second = min(second, t)
// which is exactly  if(d<best){second=best;best=d;idx=j}else if(d<second)second=d  with "first index wins" because
// keys of equal distance order by index. Partial results of the train slices are merged with a 64-bit CAS into
// one packed word per query; the same merge rule joins the per-GPU results after the NCCL all-gather.
// This kernel is bound by the integer/popc pipes, not by HBM (32 B per train row are reused by every query).
#include "orbx_internal.cuh"
#include <algorithm>

#define HT_THREADS 256
#define HT_QPT 2                 // queries per thread
#define HT_TILE 256              // train rows per shared-memory tile (8 KB)
#ifndef HT_GROUP
#define HT_GROUP 4                // train rows per top-2 update group
#endif
#define HT_SHIFT 22              // local train index bits inside a key (slice <= 4M rows)
#define HT_SLICE_MAX (1 << HT_SHIFT)

__host__ __device__ __forceinline__ unsigned long long ht_pack(int d1, int d2, unsigned idx)
{
    return ((unsigned long long)(unsigned)d1 << 48) | ((unsigned long long)(unsigned)d2 << 32) | idx;
}
// merge rule: best = smaller distance, lower index on ties; second = 2nd smallest of the multiset {d1a,d2a,d1b,d2b}
__device__ __forceinline__ unsigned long long ht_merge(unsigned long long a, unsigned long long b)
{
    const int d1a = (int)(a >> 48), d2a = (int)((a >> 32) & 0xffff), d1b = (int)(b >> 48), d2b = (int)((b >> 32) & 0xffff);
    const unsigned ia = (unsigned)a, ib = (unsigned)b;
    const bool a_first = d1a < d1b || (d1a == d1b && ia <= ib);
    const int d1 = a_first ? d1a : d1b;
    const unsigned idx = a_first ? ia : ib;
    const int d2 = min(max(d1a, d1b), min(d2a, d2b));
    return ht_pack(d1, d2, idx);
}

__global__ void hamming_init_kernel(unsigned long long* packed, int nq)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) packed[i] = ht_pack(256, 256, 0xffffffffu);
}

// 256-bit Hamming distance. The popcount unit issues 16 lanes per clock per SM, a quarter of the logic pipe, and eight
// POPC per pair made it the binding unit (0.91 of its ceiling, profiles/). A carry-save adder tree over the eight XOR
// words (Harley-Seal) moves the work to 3-input logic ops: two full adders compress x0..x5 into two "ones" and two "twos"
// words, a third takes the ones with x6, a fourth the three twos, so
//   dist = popc(s3) + popc(x7) + 2 * popc(s4) + 4 * popc(c4)
// — four POPC and eight extra LOP3 (sum = a ^ b ^ c, carry = majority(a, b, c)) instead of eight POPC.
// (spelled as PTX so that the compiler keeps the adder tree as written: left to itself it re-associates the XORs through the
// tree and ends up with 23 LOP3 per pair instead of 16)
__device__ __forceinline__ unsigned ht_xor(unsigned a, unsigned b) { unsigned r; asm("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned ht_xor3(unsigned a, unsigned b, unsigned c) { unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ unsigned ht_maj(unsigned a, unsigned b, unsigned c) { unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ int ht_mad(int a, int b, int c) { int r; asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// key = dist << HT_SHIFT | idx, with the weighted popcount sum and the shift folded into multiply-adds (FMA pipe): the logic
// pipe is the binding one (16 LOP3 + 3 min/max per pair), every op taken off it counts
__device__ __forceinline__ int ht_key(const uint4 qa, const uint4 qb, const uint4 ta, const uint4 tb, const int idx)
{
    const unsigned x0 = ht_xor(qa.x, ta.x), x1 = ht_xor(qa.y, ta.y), x2 = ht_xor(qa.z, ta.z), x3 = ht_xor(qa.w, ta.w);
    const unsigned x4 = ht_xor(qb.x, tb.x), x5 = ht_xor(qb.y, tb.y), x6 = ht_xor(qb.z, tb.z), x7 = ht_xor(qb.w, tb.w);
    const unsigned s1 = ht_xor3(x0, x1, x2), c1 = ht_maj(x0, x1, x2);
    const unsigned s2 = ht_xor3(x3, x4, x5), c2 = ht_maj(x3, x4, x5);
    const unsigned s3 = ht_xor3(s1, s2, x6), c3 = ht_maj(s1, s2, x6);
    const unsigned s4 = ht_xor3(c1, c2, c3), c4 = ht_maj(c1, c2, c3);
    int k = ht_mad(__popc(c4), 4 << HT_SHIFT, idx);
    k = ht_mad(__popc(s4), 2 << HT_SHIFT, k);
    k = ht_mad(__popc(s3), 1 << HT_SHIFT, k);
    return ht_mad(__popc(x7), 1 << HT_SHIFT, k);
}
__device__ __forceinline__ int ht_dist(const uint4 qa, const uint4 qb, const uint4 ta, const uint4 tb)
{
    const unsigned x0 = ht_xor(qa.x, ta.x), x1 = ht_xor(qa.y, ta.y), x2 = ht_xor(qa.z, ta.z), x3 = ht_xor(qa.w, ta.w);
    const unsigned x4 = ht_xor(qb.x, tb.x), x5 = ht_xor(qb.y, tb.y), x6 = ht_xor(qb.z, tb.z), x7 = ht_xor(qb.w, tb.w);
#ifdef ORBX_HAMMING_PLAIN_POPC
    return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
#else
    const unsigned s1 = ht_xor3(x0, x1, x2), c1 = ht_maj(x0, x1, x2);
    const unsigned s2 = ht_xor3(x3, x4, x5), c2 = ht_maj(x3, x4, x5);
    const unsigned s3 = ht_xor3(s1, s2, x6), c3 = ht_maj(s1, s2, x6);
    const unsigned s4 = ht_xor3(c1, c2, c3), c4 = ht_maj(c1, c2, c3);
    return __popc(s3) + __popc(x7) + 2 * __popc(s4) + 4 * __popc(c4);
#endif
}

__global__ void __launch_bounds__(HT_THREADS) hamming_top2_kernel(const uint4* __restrict__ q, int nq,
                                                                  const uint4* __restrict__ t, int nt, int slice,
                                                                  long long index_base,
                                                                  unsigned long long* __restrict__ packed,
                                                                  const OrbxHtPeer peer)
{
    __shared__ uint4 s_t[2][HT_TILE * 2];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * slice;                       // this CTA's train slice [t0, t1)
    const int t1 = min(nt, t0 + slice);
    const int qbase = blockIdx.y * (HT_THREADS * HT_QPT);
    uint4 qa[HT_QPT], qb[HT_QPT];
    int best[HT_QPT], second[HT_QPT];
#pragma unroll
    for (int k = 0; k < HT_QPT; k++) {
        const int qi = qbase + k * HT_THREADS + tid;
        if (qi < nq) { qa[k] = q[2 * (size_t)qi]; qb[k] = q[2 * (size_t)qi + 1]; }
        else { qa[k] = make_uint4(0, 0, 0, 0); qb[k] = qa[k]; }
        best[k] = 256 << HT_SHIFT;                           // reference init: bestDist = bestDist2 = 256
        second[k] = 256 << HT_SHIFT;
    }
    const int ntiles = (t1 - t0 + HT_TILE - 1) / HT_TILE;
    // prologue: tile 0
    if (ntiles > 0) {
        for (int i = tid; i < HT_TILE * 2; i += HT_THREADS) {
            const int row = t0 + (i >> 1);
            s_t[0][i] = row < t1 ? t[2 * (size_t)t0 + i] : make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    for (int tile = 0; tile < ntiles; tile++) {
        const int buf = tile & 1;
        // prefetch the next tile into registers, store after the compute loop
        uint4 pre[2];
        const int nbase = t0 + (tile + 1) * HT_TILE;
        const bool has_next = tile + 1 < ntiles;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int i = tid + r * HT_THREADS;
            const int row = nbase + (i >> 1);
            pre[r] = (has_next && row < t1) ? t[2 * (size_t)nbase + i] : make_uint4(0, 0, 0, 0);
        }
        const int rows = min(HT_TILE, t1 - (t0 + tile * HT_TILE));
        const int jbase = tile * HT_TILE;
        // Rows go HT_GROUP at a time: the top-2 update (three min/max per pair on the binding logic pipe) only runs when some
        // lane's group minimum beats its current second best — after the first few thousand rows that is under 1 % of the
        // groups (a running minimum improves ~ln(n) times). A skipped update would have been a no-op for every lane
        // (key >= second >= best leaves both unchanged), so the result is the same.
        int j = 0;
#ifndef ORBX_HAMMING_PLAIN_POPC
        for (; j + HT_GROUP <= rows; j += HT_GROUP) {
            int key[HT_GROUP][HT_QPT];
            bool improve = false;
#pragma unroll
            for (int r = 0; r < HT_GROUP; r++) {
                const uint4 ta = s_t[buf][2 * (j + r)], tb = s_t[buf][2 * (j + r) + 1];
#pragma unroll
                for (int k = 0; k < HT_QPT; k++) key[r][k] = ht_key(qa[k], qb[k], ta, tb, jbase + j + r);
            }
#pragma unroll
            for (int k = 0; k < HT_QPT; k++) {
                int g = key[0][k];
#pragma unroll
                for (int r = 1; r < HT_GROUP; r++) g = min(g, key[r][k]);
                improve = improve || g < second[k];
            }
            if (__any_sync(0xffffffffu, improve)) {
#pragma unroll
                for (int r = 0; r < HT_GROUP; r++)
#pragma unroll
                    for (int k = 0; k < HT_QPT; k++) {
                        const int hi = max(key[r][k], best[k]);
                        best[k] = min(best[k], key[r][k]);
                        second[k] = min(second[k], hi);
                    }
            }
        }
#endif
        for (; j < rows; j++) {
            const uint4 ta = s_t[buf][2 * j], tb = s_t[buf][2 * j + 1];
#pragma unroll
            for (int k = 0; k < HT_QPT; k++) {
#ifdef ORBX_HAMMING_PLAIN_POPC
                const int key = (ht_dist(qa[k], qb[k], ta, tb) << HT_SHIFT) | (jbase + j);
#else
                const int key = ht_key(qa[k], qb[k], ta, tb, jbase + j);
#endif
                const int hi = max(key, best[k]);
                best[k] = min(best[k], key);
                second[k] = min(second[k], hi);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) s_t[buf ^ 1][tid + r * HT_THREADS] = pre[r];
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < HT_QPT; k++) {
        const int qi = qbase + k * HT_THREADS + tid;
        if (qi >= nq) continue;
        const int d1 = best[k] >> HT_SHIFT, d2 = min(second[k] >> HT_SHIFT, 256);
        if (d1 >= 256) continue;                              // nothing closer than 256: leave the init value
        const unsigned idx = (unsigned)(index_base + t0 + (best[k] & (HT_SLICE_MAX - 1)));
        const unsigned long long mine = ht_pack(d1, d2, idx);
        unsigned long long old = packed[qi];
        while (true) {
            const unsigned long long want = ht_merge(old, mine);
            if (want == old) break;
            const unsigned long long prev = atomicCAS(&packed[qi], old, want);
            if (prev == old) break;
            old = prev;
        }
    }
    // ---- fused exchange (multi-GPU): the LAST slice-CTA of a query tile pushes the tile's merged local result into
    // every peer's landing buffer with plain 64-bit stores over NVLink and then bumps the peers' arrival counters.
    // This replaces all-gather + a separate collective launch; peers merge after their counter reaches the target.
    if (peer.world > 0) {
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&peer.tile_done[blockIdx.y], 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
#pragma unroll
            for (int k = 0; k < HT_QPT; k++) {
                const int qi = qbase + k * HT_THREADS + tid;
                if (qi >= nq) continue;
                const unsigned long long v = __ldcg(&packed[qi]);
                for (int p = 0; p < peer.world; p++) peer.parts[p][(size_t)peer.rank * peer.nq_max + qi] = v;
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                peer.tile_done[blockIdx.y] = 0;
                for (int p = 0; p < peer.world; p++) atomicAdd_system(peer.arrive[p], 1u);
            }
        }
    }
}

// Waits (bounded) until every rank's every query tile has landed, then merges the world partial results.
__global__ void __launch_bounds__(1024) hamming_wait_merge_kernel(const unsigned long long* parts, const unsigned* arrive,
                                                                  unsigned target, int world, int nq, int nq_max,
                                                                  int* __restrict__ idx, int* __restrict__ d1,
                                                                  int* __restrict__ d2, int* __restrict__ status)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int ok = 1;
        // wrap-safe compare; ~4 s at 2 GHz is far beyond any legitimate wait and keeps a lost peer from hanging the GPU
        while ((int)(*reinterpret_cast<const volatile unsigned*>(arrive) - target) < 0) {
            if (clock64() - t0 > 8000000000LL) { ok = 0; break; }
            __nanosleep(500);
        }
        __threadfence_system();
        s_ok = ok;
        if (status) *status = ok ? 0 : 1;
    }
    __syncthreads();
    if (!s_ok) {
        // a peer never arrived: the landing buffer may be partly written, so report "no match" for every query instead of
        // leaving the caller's buffers uninitialised; *status = 1 tells the host that this matcher must be recreated
        for (int i = threadIdx.x; i < nq; i += blockDim.x) {
            if (idx) idx[i] = -1;
            if (d1) d1[i] = 256;
            if (d2) d2[i] = 256;
        }
        return;
    }
    for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        unsigned long long acc = ht_pack(256, 256, 0xffffffffu);
        for (int p = 0; p < world; p++) acc = ht_merge(acc, __ldcv(&parts[(size_t)p * nq_max + i]));
        const int b = (int)(acc >> 48);
        if (idx) idx[i] = b >= 256 ? -1 : (int)(unsigned)acc;
        if (d1) d1[i] = b;
        if (d2) d2[i] = (int)((acc >> 32) & 0xffff);
    }
}

__global__ void hamming_merge_kernel(const unsigned long long* __restrict__ parts, int nparts, int nq,
                                     int* __restrict__ idx, int* __restrict__ d1, int* __restrict__ d2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    unsigned long long acc = ht_pack(256, 256, 0xffffffffu);
    for (int p = 0; p < nparts; p++) acc = ht_merge(acc, parts[(size_t)p * nq + i]);
    const int b = (int)(acc >> 48);
    if (idx) idx[i] = b >= 256 ? -1 : (int)(unsigned)acc;
    if (d1) d1[i] = b;
    if (d2) d2[i] = (int)((acc >> 32) & 0xffff);
}

void orbx_launch_hamming_init(uint64_t* d_packed, int nq, cudaStream_t st)
{
    if (nq > 0) hamming_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(d_packed), nq);
}

int orbx_hamming_qtiles(int nq) { return (nq + HT_THREADS * HT_QPT - 1) / (HT_THREADS * HT_QPT); }

void orbx_launch_hamming_top2_peer(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, long long index_base,
                                   uint64_t* d_packed, const OrbxHtPeer& peer, cudaStream_t st)
{
    if (nq <= 0) return;
    const int qtiles = orbx_hamming_qtiles(nq);
    // size the grid to ~4 waves of 148 SMs x 2 resident CTAs, slices a multiple of the tile
    int want = (148 * 2 * 4 + qtiles - 1) / qtiles;
    int slice = (std::max(nt, 1) + want - 1) / want;
    slice = ((slice + HT_TILE - 1) / HT_TILE) * HT_TILE;
    if (slice > HT_SLICE_MAX) slice = HT_SLICE_MAX;
    const int nslices = std::max(1, (nt + slice - 1) / slice);      // an empty shard still has to signal its peers
    dim3 grid(nslices, qtiles);
    hamming_top2_kernel<<<grid, HT_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(d_q), nq,
                                                     reinterpret_cast<const uint4*>(d_t), nt, slice, index_base,
                                                     reinterpret_cast<unsigned long long*>(d_packed), peer);
}

void orbx_launch_hamming_top2(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, long long index_base,
                              uint64_t* d_packed, cudaStream_t st)
{
    if (nq <= 0 || nt <= 0) return;
    OrbxHtPeer none{};
    orbx_launch_hamming_top2_peer(d_q, nq, d_t, nt, index_base, d_packed, none, st);
}

void orbx_launch_hamming_wait_merge(const uint64_t* d_parts_local, const unsigned* d_arrive_local, unsigned target,
                                    int world, int nq, int nq_max, int* d_idx, int* d_d1, int* d_d2, int* d_status,
                                    cudaStream_t st)
{
    hamming_wait_merge_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_parts_local), d_arrive_local,
                                                  target, world, nq, nq_max, d_idx, d_d1, d_d2, d_status);
}

void orbx_launch_hamming_merge(const uint64_t* d_parts, int nparts, int nq, int* d_idx, int* d_d1, int* d_d2,
                               cudaStream_t st)
{
    if (nq > 0)
        hamming_merge_kernel<<<(nq + 255) / 256, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_parts),
                                                               nparts, nq, d_idx, d_d1, d_d2);
}

// ---------------------------------------------------------------- stereo row-band matching (Frame.cc:596-663)
// One warp per left keypoint; lanes stride over the candidates of row (int)vL (CSR table built on the host in the
// reference's push_back order = ascending right index). Warp-reduced best with (distance, candidate position) keys
// so the first candidate attaining the minimum wins, as with the reference's strict '<'.
__global__ void __launch_bounds__(256) stereo_hamming_kernel(const OrbxKp28* __restrict__ kl, const uint4* __restrict__ dl,
                                                             int nl, const OrbxKp28* __restrict__ kr,
                                                             const uint4* __restrict__ dr, const int* __restrict__ row_start,
                                                             const int* __restrict__ row_tab, int rows, float minD,
                                                             float maxD, int* __restrict__ best_idx,
                                                             int* __restrict__ best_dist)
{
    const int lane = threadIdx.x & 31;
    const int iL = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (iL >= nl) return;
    const OrbxKp28 k = kl[iL];
    const int row = (int)k.y;
    int key = (100 << 20) | 0xfffff;                          // bestDist = TH_HIGH, no index yet
    if (row >= 0 && row < rows) {
        const float minU = __fsub_rn(k.x, maxD), maxU = __fsub_rn(k.x, minD);
        if (!(maxU < 0.f)) {
            const uint4 qa = dl[2 * (size_t)iL], qb = dl[2 * (size_t)iL + 1];
            const int c0 = row_start[row], c1 = row_start[row + 1];
            for (int c = c0 + lane; c < c1; c += 32) {
                const int iR = row_tab[c];
                const OrbxKp28 r = kr[iR];
                if (r.octave < k.octave - 1 || r.octave > k.octave + 1) continue;
                if (r.x >= minU && r.x <= maxU) {
                    const int d = ht_dist(qa, qb, dr[2 * (size_t)iR], dr[2 * (size_t)iR + 1]);
                    if (d < 100) key = min(key, (d << 20) | min(c - c0, 0xffffe));
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    if (lane == 0) {
        const int pos = key & 0xfffff;
        best_dist[iL] = key >> 20;
        best_idx[iL] = pos == 0xfffff ? -1 : row_tab[row_start[row] + pos];
    }
}

void orbx_launch_stereo_hamming(const OrbxKp28* kl, const uint8_t* dl, int nl, const OrbxKp28* kr, const uint8_t* dr,
                                int nr, const int* row_start, const int* row_tab, int rows, float minD, float maxD,
                                int* best_idx, int* best_dist, cudaStream_t st)
{
    (void)nr;
    if (nl <= 0) return;
    const int warps_per_block = 8;
    stereo_hamming_kernel<<<(nl + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(
        kl, reinterpret_cast<const uint4*>(dl), nl, kr, reinterpret_cast<const uint4*>(dr), row_start, row_tab, rows,
        minD, maxD, best_idx, best_dist);
}

// ---------------------------------------------------------------- full stereo matching (Frame.cc:547-788)
// Device-resident and batched over stereo pairs. Kernel 1: one warp per left keypoint. The reference's row table
// (vRowIndices, Frame.cc:564-590) is not materialised: the CTA stages (band, octave, x) of every right keypoint of the
// pair in shared memory and each warp tests "row (int)vL lies in the band" directly while striding over the right
// keypoints in ascending index — the order the table would have given — then takes the warp-min of
// (distance << 20 | right index): first candidate attaining the minimum, strict '<' against TH_HIGH = 100.
// For a match below 75 the 11x11 SAD of centre-normalised windows is slid over +-5 columns on the level pyramids of
// the two extractors, a parabola is fitted through the three SADs around the minimum and disparity / depth follow
// Frame.cc:733-760. Window arithmetic is integer (the reference converts to float only to call cv::norm); the float
// steps use un-contracted IEEE operations. Kernel 2: one CTA per pair finds the median SAD with a two-pass radix
// select and applies the 1.5*1.4*median cut (Frame.cc:774-787).
struct StereoR { float x; short minr, maxr; int octave; };   // 12 bytes per right keypoint
#define STEREO_LPC 64                                        // left keypoints per CTA

// BANDS: the right keypoints are additionally bucketed by 8-row bands (a keypoint enters every band its row interval
// touches; unordered lists, built with shared-memory atomics), so a left keypoint only looks at the right keypoints
// of its own band — the (distance << 20 | right index) key restores the ascending-index tie-break of vRowIndices.
#define STEREO_MAXBANDS 512
template <bool BANDS>
__global__ void __launch_bounds__(256) stereo_match_batch_kernel(OrbxStereoBatch A, int band_cap)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    StereoR* sr = reinterpret_cast<StereoR*>(s_raw);
    __shared__ int s_bstart[BANDS ? STEREO_MAXBANDS + 1 : 1], s_bfill[BANDS ? STEREO_MAXBANDS + 1 : 1];
    unsigned short* bent = reinterpret_cast<unsigned short*>(sr + A.cap);      // [band_cap] right-keypoint indices by band
    const int pair = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int nl = min(A.nl[pair], A.cap), nr = min(A.nr[pair], A.cap);
    if ((int)(blockIdx.x * STEREO_LPC) >= nl) return;         // whole CTA
    const OrbxKp28* kl = A.kl + (size_t)pair * A.cap;
    const OrbxKp28* kr = A.kr + (size_t)pair * A.cap;
    const uint4* dl = reinterpret_cast<const uint4*>(A.dl + (size_t)pair * A.cap * 32);
    const uint4* dr = reinterpret_cast<const uint4*>(A.dr + (size_t)pair * A.cap * 32);
    for (int i = threadIdx.x; i < nr; i += blockDim.x) {
        const OrbxKp28 r = kr[i];
        const float rad = __fmul_rn(2.0f, A.lvl[r.octave].scale);           // r = 2*mvScaleFactors[octave]
        StereoR e;
        e.x = r.x; e.octave = r.octave;
        e.maxr = (short)(int)ceilf(__fadd_rn(r.y, rad));
        e.minr = (short)(int)floorf(__fsub_rn(r.y, rad));
        sr[i] = e;
    }
    __syncthreads();
    if (BANDS) {
        const int nb = (A.rows + 7) >> 3;
        for (int b = threadIdx.x; b <= nb; b += blockDim.x) s_bfill[b] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < nr; i += blockDim.x) {
            const int b0 = max((int)sr[i].minr, 0) >> 3, b1 = min((int)sr[i].maxr, A.rows - 1) >> 3;
            for (int b = b0; b <= b1; b++) atomicAdd(&s_bfill[b], 1);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                   // exclusive scan of up to 512 band counts by one warp
            int carry = 0;
            for (int b0 = 0; b0 <= nb; b0 += 32) {
                const int b = b0 + lane;
                const int c = b < nb ? s_bfill[b] : 0;
                int x = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                if (b <= nb) s_bstart[b] = carry + x - c;
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b <= nb; b += blockDim.x) s_bfill[b] = s_bstart[b];
        __syncthreads();
        for (int i = threadIdx.x; i < nr; i += blockDim.x) {
            const int b0 = max((int)sr[i].minr, 0) >> 3, b1 = min((int)sr[i].maxr, A.rows - 1) >> 3;
            for (int b = b0; b <= b1; b++) { const int pos = atomicAdd(&s_bfill[b], 1); if (pos < band_cap) bent[pos] = (unsigned short)i; }
        }
        __syncthreads();
    }
    // a CTA stages the right keypoints once and serves STEREO_LPC left keypoints (8 warps, several keypoints each)
    for (int iL = blockIdx.x * STEREO_LPC + (threadIdx.x >> 5); iL < min(nl, (int)(blockIdx.x + 1) * STEREO_LPC); iL += 8) {
    const OrbxKp28 k = kl[iL];
    const int row = (int)k.y;
    int key = (100 << 20) | 0xfffff;
    if (row >= 0 && row < A.rows) {
        const float minU = __fsub_rn(k.x, A.maxD), maxU = __fsub_rn(k.x, A.minD);
        if (!(maxU < 0.f)) {
            const uint4 qa = dl[2 * (size_t)iL], qb = dl[2 * (size_t)iL + 1];
            const int j0 = BANDS ? s_bstart[row >> 3] : 0, j1 = BANDS ? min(s_bstart[(row >> 3) + 1], band_cap) : nr;
            for (int j = j0 + lane; j < j1; j += 32) {
                const int iR = BANDS ? (int)bent[j] : j;
                const StereoR r = sr[iR];
                if (row < r.minr || row > r.maxr) continue;
                if (r.octave < k.octave - 1 || r.octave > k.octave + 1) continue;
                if (r.x >= minU && r.x <= maxU) {
                    const int d = ht_dist(qa, qb, dr[2 * (size_t)iR], dr[2 * (size_t)iR + 1]);
                    if (d < 100) key = min(key, (d << 20) | iR);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    float out_u = -1.f, out_d = -1.f;
    int out_sad = -1;
    const int iR = key & 0xfffff;
    if (iR != 0xfffff && (key >> 20) < 75) {                  // thOrbDist = (TH_HIGH + TH_LOW) / 2  (warp-uniform)
        const OrbxLevelGeom g = A.lvl[k.octave];
        const float uR0 = sr[iR].x;
        const float scaleduL = roundf(__fmul_rn(k.x, g.inv_scale));
        const float scaledvL = roundf(__fmul_rn(k.y, g.inv_scale));
        const float scaleduR0 = roundf(__fmul_rn(uR0, g.inv_scale));
        const float iniu = __fsub_rn(__fadd_rn(scaleduR0, 5.f), 5.f);          // scaleduR0 + L - w  (sic)
        const float endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, 5.f), 5.f), 1.f);
        if (!(iniu < 0.f || endu >= (float)g.w)) {
            const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
            const uint8_t* pL = A.raw_left + (size_t)pair * A.frame_raw_bytes + g.raw_off + (size_t)(cv + ORBX_EDGE) * g.pitch + (cu + ORBX_XOFF);
            const uint8_t* pR = A.raw_right + (size_t)pair * A.frame_raw_bytes + g.raw_off + (size_t)(cv + ORBX_EDGE) * g.pitch + (cr + ORBX_XOFF);
            const int cL = pL[0];
            int cR[11], acc[11];
#pragma unroll
            for (int s = 0; s < 11; s++) { cR[s] = pR[s - 5]; acc[s] = 0; }
            for (int i = lane; i < 121; i += 32) {
                const int dy = i / 11 - 5, dx = i - (dy + 5) * 11 - 5;
                const int a = (int)pL[dy * g.pitch + dx] - cL;
                const uint8_t* q = pR + dy * g.pitch + dx;
#pragma unroll
                for (int s = 0; s < 11; s++) acc[s] += abs(a - ((int)q[s - 5] - cR[s]));
            }
#pragma unroll
            for (int s = 0; s < 11; s++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
            int bestDist = 2147483647, bestinc = 0;
#pragma unroll
            for (int s = 0; s < 11; s++) if ((float)acc[s] < (float)bestDist) { bestDist = acc[s]; bestinc = s - 5; }
            if (bestinc != -5 && bestinc != 5) {
                float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int s = 1; s < 10; s++) if (s - 5 == bestinc) { d1 = (float)acc[s - 1]; d2 = (float)acc[s]; d3 = (float)acc[s + 1]; }
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                if (!(deltaR < -1.f || deltaR > 1.f)) {
                    float bestuR = __fmul_rn(g.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(k.x, bestuR);
                    if (disparity >= A.minD && disparity < A.maxD) {
                        if (disparity <= 0.f) { disparity = 0.01f; bestuR = (float)((double)k.x - 0.01); }
                        out_d = __fdiv_rn(A.mbf, disparity);
                        out_u = bestuR;
                        out_sad = bestDist;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        const size_t o = (size_t)pair * A.cap + iL;
        A.u_right[o] = out_u; A.depth[o] = out_d; A.sad[o] = out_sad;
    }
    }
}

// median of the matched SADs (the element at index size/2 of the sorted list, Frame.cc:775) by a two-pass radix
// select on the 16-bit values, then mvuRight = mvDepth = -1 for every match with SAD >= 1.5f*1.4f*median
__global__ void __launch_bounds__(256) stereo_median_cut_kernel(OrbxStereoBatch A)
{
    __shared__ int hist[256];
    __shared__ int s_sel[3];
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int nl = min(A.nl[pair], A.cap);
    const int* sad = A.sad + (size_t)pair * A.cap;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += 256) { const int v = sad[i]; if (v >= 0) atomicAdd(&hist[min(v >> 8, 255)], 1); }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int b = 0; b < 256; b++) m += hist[b];
        int k = m / 2, cum = 0, bin = -1;
        for (int b = 0; b < 256 && bin < 0; b++) { if (cum + hist[b] > k) bin = b; else cum += hist[b]; }
        s_sel[0] = m; s_sel[1] = bin; s_sel[2] = k - cum;
    }
    __syncthreads();
    const int m = s_sel[0], bin = s_sel[1], k2 = s_sel[2];
    if (m == 0) return;                                       // the reference indexes an empty vector here
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += 256) { const int v = sad[i]; if (v >= 0 && min(v >> 8, 255) == bin) atomicAdd(&hist[v & 255], 1); }
    __syncthreads();
    if (tid == 0) {
        int cum = 0, lo = 0;
        for (int b = 0; b < 256; b++) { if (cum + hist[b] > k2) { lo = b; break; } cum += hist[b]; }
        s_sel[1] = (bin << 8) | lo;
    }
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)s_sel[1]);
    for (int i = tid; i < nl; i += 256) {
        const int v = sad[i];
        if (v >= 0 && !((float)v < thDist)) {
            A.u_right[(size_t)pair * A.cap + i] = -1.f;
            A.depth[(size_t)pair * A.cap + i] = -1.f;
        }
    }
}

void orbx_launch_stereo_batch(const OrbxStereoBatch& a, cudaStream_t st)
{
    if (a.pairs <= 0 || a.cap <= 0) return;
    // band lists: a right keypoint's row interval [y - 2 s, y + 2 s] touches at most (4 s + 2) / 8 + 2 bands; 8 entries
    // per keypoint cover every pyramid this library accepts (beyond scale 10 the plain path runs)
    const size_t smem_plain = (size_t)a.cap * sizeof(StereoR);
    const int band_cap = a.cap * 8;
    const size_t smem_bands = smem_plain + (size_t)band_cap * 2;
    const bool bands = smem_bands <= 200 * 1024 && a.rows <= 8 * STEREO_MAXBANDS && a.max_scale <= 10.0f;   // interval width <= 4 s + 3 = 43 rows -> at most 7 bands
    static OrbxSmemMark mark[2] = {};
    dim3 grid((a.cap + STEREO_LPC - 1) / STEREO_LPC, a.pairs);
    if (bands) {
        orbx_need_smem(stereo_match_batch_kernel<true>, mark[0], smem_bands);
        stereo_match_batch_kernel<true><<<grid, 256, smem_bands, st>>>(a, band_cap);
    } else {
        orbx_need_smem(stereo_match_batch_kernel<false>, mark[1], smem_plain);
        stereo_match_batch_kernel<false><<<grid, 256, smem_plain, st>>>(a, 0);
    }
    stereo_median_cut_kernel<<<a.pairs, 256, 0, st>>>(a);
}

// ---------------------------------------------------------------- windowed top-2 on the Frame grid
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:46-142): per projected map point the
// candidates of Frame::GetFeaturesInArea (Frame.cc:388-444) are scanned in grid order (cell column, cell row, then
// ascending keypoint index inside a cell — the order AssignFeaturesToGrid filled the cells in) keeping best / second
// best with strict '<'. Here every CTA rebuilds Frame::mGrid in shared memory as a CSR (counting sort + per-cell
// insertion sort for the ascending order), then one THREAD per query walks the cells of its window in exactly that
// order and runs the reference's update rule verbatim — no tie-break keys needed.
struct WinKp { float x, y; int octave; };
#define WIN_CELLS (64 * 48)
#define WIN_THREADS 512

__global__ void __launch_bounds__(WIN_THREADS) window_top2_kernel(OrbxWindowArgs A, int q_per_cta)
{
    extern __shared__ __align__(16) unsigned char s_raw2[];
    WinKp* sk = reinterpret_cast<WinKp*>(s_raw2);                                    // [n]
    unsigned short* order = reinterpret_cast<unsigned short*>(sk + A.n);            // [n] keypoint indices sorted by cell
    unsigned short* cstart = order + ((A.n + 1) & ~1);                              // [WIN_CELLS + 1]
    unsigned short* cfill = cstart + WIN_CELLS + 2;                                 // [WIN_CELLS]
    __shared__ int s_w[17];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int c = tid; c < WIN_CELLS; c += WIN_THREADS) cfill[c] = 0;
    __syncthreads();
    for (int i = tid; i < A.n; i += WIN_THREADS) {
        const OrbxKp28 k = A.kps[i];
        WinKp e; e.x = k.x; e.y = k.y; e.octave = k.octave;
        sk[i] = e;
        const int posX = (int)roundf(__fmul_rn(__fsub_rn(k.x, A.minX), A.invW));     // Frame::PosInGrid
        const int posY = (int)roundf(__fmul_rn(__fsub_rn(k.y, A.minY), A.invH));
        if (!(posX < 0 || posX >= 64 || posY < 0 || posY >= 48)) {
            const int c = posX * 48 + posY;
            atomicAdd(reinterpret_cast<unsigned*>(cfill) + (c >> 1), (c & 1) ? 0x10000u : 1u);   // two 16-bit counters per word
        }
    }
    __syncthreads();
    {
        const int c0 = tid * 6;                                                      // 512 threads x 6 cells = 3072
        int loc[6], sum = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) { loc[j] = sum; sum += cfill[c0 + j]; }
        int x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            const int t = lane < 16 ? s_w[lane] : 0;
            int z = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            if (lane < 16) s_w[lane] = z - t;
            if (lane == 15) s_w[16] = z;
        }
        __syncthreads();
        const int base = s_w[wid] + x - sum;
#pragma unroll
        for (int j = 0; j < 6; j++) { cstart[c0 + j] = (unsigned short)(base + loc[j]); cfill[c0 + j] = (unsigned short)(base + loc[j]); }
        if (tid == 0) cstart[WIN_CELLS] = (unsigned short)s_w[16];
    }
    __syncthreads();
    for (int i = tid; i < A.n; i += WIN_THREADS) {
        const WinKp k = sk[i];
        const int posX = (int)roundf(__fmul_rn(__fsub_rn(k.x, A.minX), A.invW));
        const int posY = (int)roundf(__fmul_rn(__fsub_rn(k.y, A.minY), A.invH));
        if (!(posX < 0 || posX >= 64 || posY < 0 || posY >= 48)) {
            const int c = posX * 48 + posY;
            const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(cfill) + (c >> 1), (c & 1) ? 0x10000u : 1u);
            order[(c & 1) ? (old >> 16) : (old & 0xffffu)] = (unsigned short)i;
        }
    }
    __syncthreads();
    for (int c = tid; c < WIN_CELLS; c += WIN_THREADS) {
        const int b0 = cstart[c], b1 = cstart[c + 1];
        for (int i = b0 + 1; i < b1; i++) {
            const unsigned short v = order[i];
            int j = i - 1;
            while (j >= b0 && order[j] > v) { order[j + 1] = order[j]; j--; }
            order[j + 1] = v;
        }
    }
    __syncthreads();

    const uint4* dsc = reinterpret_cast<const uint4*>(A.desc);
    const int q_end = min(A.nq, (int)(blockIdx.x + 1) * q_per_cta);
    for (int qi = blockIdx.x * q_per_cta + tid; qi < q_end; qi += WIN_THREADS) {
        const OrbxWinQuery q = A.q[qi];
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.x, A.minX), q.r), A.invW)));
        const int cx1 = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.x, A.minX), q.r), A.invW)));
        const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.y, A.minY), q.r), A.invH)));
        const int cy1 = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.y, A.minY), q.r), A.invH)));
        if (!(cx0 >= 64 || cx1 < 0 || cy0 >= 48 || cy1 < 0)) {
            const bool check_levels = q.min_level > 0 || q.max_level >= 0;
            const uint4 qa = reinterpret_cast<const uint4*>(A.qdesc)[2 * (size_t)qi], qb = reinterpret_cast<const uint4*>(A.qdesc)[2 * (size_t)qi + 1];
            for (int ix = cx0; ix <= cx1; ix++) {
                // rows cy0..cy1 of one cell column are one contiguous CSR range
                const int j0 = cstart[ix * 48 + cy0], j1 = cstart[ix * 48 + cy1 + 1];
                for (int j = j0; j < j1; j++) {
                    const int i = order[j];
                    const WinKp k = sk[i];
                    if (check_levels) {
                        if (k.octave < q.min_level) continue;
                        if (q.max_level >= 0 && k.octave > q.max_level) continue;
                    }
                    if (!(fabsf(__fsub_rn(k.x, q.x)) < q.r && fabsf(__fsub_rn(k.y, q.y)) < q.r)) continue;
                    if (A.occupied && A.occupied[i]) continue;
                    if (A.u_right) { const float ur = A.u_right[i]; if (ur > 0.f && fabsf(__fsub_rn(q.xr, ur)) > q.r) continue; }
                    const int d = ht_dist(qa, qb, dsc[2 * (size_t)i], dsc[2 * (size_t)i + 1]);
                    if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = k.octave; bestIdx = i; }
                    else if (d < bestDist2) { bestLevel2 = k.octave; bestDist2 = d; }
                }
            }
        }
        A.best_idx[qi] = bestIdx; A.best_dist[qi] = bestDist; A.best_level[qi] = bestLevel;
        A.best_dist2[qi] = bestDist2; A.best_level2[qi] = bestLevel2;
    }
}

void orbx_launch_window_top2(const OrbxWindowArgs& a, cudaStream_t st)
{
    if (a.nq <= 0) return;
    const size_t n1 = (size_t)(a.n > 0 ? a.n : 1);
    const size_t smem = n1 * sizeof(WinKp) + ((n1 + 1) & ~(size_t)1) * 2 + (size_t)(2 * WIN_CELLS + 4) * 2 + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(window_top2_kernel, mark[0], smem);
    // every CTA rebuilds the grid, so give each one at least a full round of queries
    const int q_per_cta = std::max(WIN_THREADS, (a.nq + 147) / 148);
    window_top2_kernel<<<(a.nq + q_per_cta - 1) / q_per_cta, WIN_THREADS, smem, st>>>(a, q_per_cta);
}

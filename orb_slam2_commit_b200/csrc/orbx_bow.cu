// orbx_bow.cu — bag-of-words rows behind the extractor:
//   * Frame::ComputeBoW (Frame.cc:462-469) = ORBVocabulary::transform(descriptors, mBowVec, mFeatVec, 4)
//     (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1196, 1216-1260): every descriptor walks the vocabulary tree
//     (k Hamming distances per level, first minimum wins), then the BowVector (word -> accumulated, normalised weight)
//     and the FeatureVector (node at level L-levelsup -> feature indices) are assembled.
//   * L1Scoring::score between BowVectors (KeyFrameDatabase.cc:145,274; LoopClosing.cc:152).
//   * ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (ORBmatcher.cc:175-325): node-constrained best / second-best
//     matching with the rotation-histogram check (ComputeThreeMaxima, :1797-1839).
// All of it batched over frames (one launch for a whole batch of extractor outputs that are still in HBM).
// Floating-point sums are taken in the reference's order with un-contracted f64 adds, so values are bit-identical.
#include "orbx_internal.cuh"
#include <algorithm>

__device__ __forceinline__ int bow_dist(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---------------------------------------------------------------------------------------------------- tree descent
// One warp per descriptor. Children of a node sit in consecutive SLOTS: lane j reads the 32-byte descriptor of child j
// (one coalesced 32*k-byte segment), the warp takes the minimum of (distance << 8 | j) — the lowest j wins ties, which
// is the reference's strict `d < best_d` scan — and the winner's (first slot, child count, node id) come from the slot
// arrays: two dependent memory round trips per level.
__global__ void __launch_bounds__(256) bow_descend_kernel(OrbxVocabDev V, const uint8_t* __restrict__ desc,
                                                          const int* __restrict__ d_n, int cap, int levelsup,
                                                          int* __restrict__ leaf_out, int* __restrict__ nid_out)
{
    const int lane = threadIdx.x & 31;
    const int feat = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int frame = blockIdx.y;
    const int n = min(d_n[frame], cap);
    if (feat >= n) return;
    const uint4* fd = reinterpret_cast<const uint4*>(desc + ((size_t)frame * cap + feat) * 32);
    const uint4 f0 = __ldg(fd), f1 = __ldg(fd + 1);
    const int nid_level = V.L - levelsup;
    int first = 0, cnt = V.root_children, node = 0, nid = 0, level = 0;
    while (cnt > 0) {
        ++level;
        unsigned key = 0xffffffffu;
        if (lane < cnt) {
            const uint4* cd = V.slot_desc + 2 * (size_t)(first + lane);
            key = ((unsigned)bow_dist(f0, f1, __ldg(cd), __ldg(cd + 1)) << 8) | (unsigned)lane;
        }
        key = __reduce_min_sync(0xffffffffu, key);
        const int s = first + (int)(key & 255u);
        const int2 kd = __ldg(V.slot_kids + s);
        node = __ldg(V.slot_node + s);
        first = kd.x; cnt = kd.y;
        if (level == nid_level) nid = node;
    }
    if (lane == 0) {
        leaf_out[(size_t)frame * cap + feat] = node;
        nid_out[(size_t)frame * cap + feat] = nid;
    }
}

// ---------------------------------------------------------------------------------------------------- per-frame assembly
// One CTA per frame. Both maps of the reference (std::map<WordId, WordValue>, std::map<NodeId, vector<unsigned>>) are
// ordered by key with values in feature order: sort (key << 32 | feature) in shared memory, segment heads are the map
// entries.
__device__ void bow_bitonic_sort(unsigned long long* a, int P)
{
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

// exclusive block scan of one int per thread-strided element is overkill here: heads are counted with a ballot-based
// two-level scan over P elements (P <= 16384, blockDim = 512)
__device__ int bow_block_excl_scan(int v, int* warp_sums, int* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = lane < nw ? warp_sums[lane] : 0, si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += y; }
        if (lane < nw) warp_sums[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    const int r = warp_sums[wid] + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(512) bow_assemble_kernel(OrbxVocabDev V, OrbxBowOut O, const int* __restrict__ d_n, int cap, int P)
{
    extern __shared__ __align__(16) unsigned long long keys[];          // P entries
    __shared__ int warp_sums[32];
    __shared__ int s_total, s_nvalid;
    __shared__ double s_norm;
    const int frame = blockIdx.x;
    const int n = min(d_n[frame], cap);
    const int* leaf = O.leaf + (size_t)frame * cap;
    const int* nid = O.nid + (size_t)frame * cap;
    int* word_out = O.word + (size_t)frame * cap;
    const bool tf = V.weighting == 0 || V.weighting == 1;                // TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist
    const bool must = V.scoring != 5;                                    // DOT_PRODUCT does not normalise

    // ---- BowVector: key = word << 32 | feature; stopped words (weight <= 0) and padding sort to the end
    if (threadIdx.x == 0) s_nvalid = 0;
    __syncthreads();
    int myvalid = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n) {
            const int lf = leaf[i];
            const int w = __ldg(V.word + lf);
            word_out[i] = w;
            if (__ldg(V.weight + lf) > 0.0) { key = ((unsigned long long)(unsigned)w << 32) | (unsigned)i; myvalid++; }
        }
        keys[i] = key;
    }
    atomicAdd(&s_nvalid, myvalid);
    __syncthreads();
    const int nvalid = s_nvalid;
    bow_bitonic_sort(keys, P);
    // heads -> entries. Every thread owns the positions p = threadIdx.x + r*blockDim.x; rounds keep the scan simple.
    int base = 0;
    int* bow_id = O.bow_id + (size_t)frame * cap;
    double* bow_val = O.bow_val + (size_t)frame * cap;
    for (int p0 = 0; p0 < nvalid; p0 += blockDim.x) {
        const int p = p0 + threadIdx.x;
        int head = 0, w = 0;
        if (p < nvalid) {
            w = (int)(keys[p] >> 32);
            head = p == 0 || (int)(keys[p - 1] >> 32) != w;
        }
        const int idx = base + bow_block_excl_scan(head, warp_sums, &s_total);
        if (head) {
            int c = 1;
            while (p + c < nvalid && (int)(keys[p + c] >> 32) == w) c++;
            const double wt = __ldg(V.weight + leaf[(unsigned)keys[p]]);   // every feature of a word carries the word's weight
            double v = wt;
            if (tf) for (int t = 1; t < c; t++) v = __dadd_rn(v, wt);    // BowVector::addWeight, once per feature, in order
            bow_id[idx] = w; bow_val[idx] = v;
        }
        base += s_total;
        __syncthreads();
    }
    const int nb = base;
    __syncthreads();
    if (threadIdx.x == 0) {
        O.n_bow[frame] = nb;
        double norm = 0.0;
        if (must) {                                                      // BowVector::normalize: serial sum in word order
            if (V.scoring != 1) for (int j = 0; j < nb; j++) norm = __dadd_rn(norm, fabs(bow_val[j]));
            else { for (int j = 0; j < nb; j++) norm = __dadd_rn(norm, __dmul_rn(bow_val[j], bow_val[j])); norm = __dsqrt_rn(norm); }
        } else if (tf && nb > 0) norm = (double)nb;
        s_norm = norm;
    }
    __syncthreads();
    if (s_norm > 0.0) for (int j = threadIdx.x; j < nb; j += blockDim.x) bow_val[j] = __ddiv_rn(bow_val[j], s_norm);
    __syncthreads();

    // ---- FeatureVector: key = node << 32 | feature
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n && __ldg(V.weight + leaf[i]) > 0.0) key = ((unsigned long long)(unsigned)nid[i] << 32) | (unsigned)i;
        keys[i] = key;
    }
    __syncthreads();
    bow_bitonic_sort(keys, P);
    int* fv_node = O.fv_node + (size_t)frame * cap;
    int* fv_off = O.fv_off + (size_t)frame * (cap + 1);
    int* fv_feat = O.fv_feat + (size_t)frame * cap;
    base = 0;
    for (int p0 = 0; p0 < nvalid; p0 += blockDim.x) {
        const int p = p0 + threadIdx.x;
        int head = 0, nd = 0;
        if (p < nvalid) {
            nd = (int)(keys[p] >> 32);
            head = p == 0 || (int)(keys[p - 1] >> 32) != nd;
            fv_feat[p] = (int)(unsigned)keys[p];
        }
        const int idx = base + bow_block_excl_scan(head, warp_sums, &s_total);
        if (head) { fv_node[idx] = nd; fv_off[idx] = p; }
        base += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { fv_off[base] = nvalid; O.n_fv[frame] = base; }
}

// ---------------------------------------------------------------------------------------------------- L1 score
// One warp per (query vector, database vector) pair. score = -1/2 * sum over common words of (|a-b| - |a| - |b|), summed
// in ascending word order like the reference's merge walk: lanes look their words up by binary search, the hits of
// every 32-word chunk are folded into the running sum serially in lane order.
__global__ void __launch_bounds__(128) bow_score_l1_kernel(OrbxBowOut O, int cap, const int* __restrict__ qa, const int* __restrict__ qb,
                                                           int npairs, double* __restrict__ score)
{
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= npairs) return;
    const int fa = qa[pair], fb = qb[pair];
    const int* ida = O.bow_id + (size_t)fa * cap; const double* va = O.bow_val + (size_t)fa * cap; const int na = O.n_bow[fa];
    const int* idb = O.bow_id + (size_t)fb * cap; const double* vb = O.bow_val + (size_t)fb * cap; const int nb = O.n_bow[fb];
    double acc = 0.0;
    for (int i0 = 0; i0 < na; i0 += 32) {
        const int i = i0 + lane;
        double term = 0.0; bool hit = false;
        if (i < na) {
            const int w = ida[i];
            int lo = 0, hi = nb;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (idb[mid] < w) lo = mid + 1; else hi = mid; }
            if (lo < nb && idb[lo] == w) {
                const double a = va[i], b = vb[lo];
                term = __dsub_rn(__dsub_rn(fabs(__dsub_rn(a, b)), fabs(a)), fabs(b));
                hit = true;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const double t = __shfl_sync(0xffffffffu, term, src);
            acc = __dadd_rn(acc, t);
        }
    }
    if (lane == 0) score[pair] = __ddiv_rn(-acc, 2.0);
}

// ---------------------------------------------------------------------------------------------------- SearchByBoW
// One warp per FeatureVector node of the keyframe. The keyframe's features of the node are taken in order (a frame
// feature matched by an earlier one is skipped by later ones, ORBmatcher.cc:228-229 — the dependency stays inside the
// node because every frame feature belongs to exactly one node); the frame's features of the node are spread over the
// lanes, key = dist << 16 | list position, so the first minimum in list order wins and the second key's distance is
// the reference's bestDist2.
__global__ void __launch_bounds__(128) bow_match_kernel(OrbxBowOut O, OrbxBowMatchArgs A, int cap)
{
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const int fk = A.kf_frame[pair], ff = A.f_frame[pair];
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= O.n_fv[fk]) return;
    const int node = O.fv_node[(size_t)fk * cap + a];
    const int* fnode = O.fv_node + (size_t)ff * cap;
    int lo = 0, hi = O.n_fv[ff];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (fnode[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= O.n_fv[ff] || fnode[lo] != node) return;
    const int* koff = O.fv_off + (size_t)fk * (cap + 1);
    const int* foff = O.fv_off + (size_t)ff * (cap + 1);
    const int* kfeat = O.fv_feat + (size_t)fk * cap;
    const int* ffeat = O.fv_feat + (size_t)ff * cap;
    const int k0 = koff[a], k1 = koff[a + 1], j0 = foff[lo], j1 = foff[lo + 1];
    const uint8_t* kdesc = A.desc + (size_t)fk * cap * 32;
    const uint8_t* fdesc = A.desc + (size_t)ff * cap * 32;
    const OrbxKp28* kkp = A.kps + (size_t)fk * cap;
    const OrbxKp28* fkp = A.kps + (size_t)ff * cap;
    const uint8_t* valid = A.kf_valid ? A.kf_valid + (size_t)pair * cap : nullptr;
    const uint8_t* fvalid = A.f_valid ? A.f_valid + (size_t)pair * cap : nullptr;
    // kf_mode (KeyFrame-KeyFrame form, ORBmatcher.cc:589-736): the output is indexed by the FIRST keyframe's feature
    // (vpMatches12[idx1]) and the second side's "already matched" flags (vbMatched2) live in `taken`
    int* match = A.match + (size_t)pair * cap;
    int* taken = A.kf_mode ? A.taken + (size_t)pair * cap : match;
    int* bin_of = A.bin_of + (size_t)pair * cap;
    for (int ik = k0; ik < k1; ik++) {
        const int ri = kfeat[ik];
        if (valid && !valid[ri]) continue;                                   // warp-uniform
        const uint4* kd = reinterpret_cast<const uint4*>(kdesc + (size_t)ri * 32);
        const uint4 q0 = __ldg(kd), q1 = __ldg(kd + 1);
        unsigned b1 = 0xffffffffu, b2 = 0xffffffffu;                         // per-lane best / second keys
        for (int j = j0 + lane; j < j1; j += 32) {
            const int rf = ffeat[j];
            if (reinterpret_cast<volatile int*>(taken)[rf] >= 0 || (fvalid && !fvalid[rf])) continue;
            const uint4* fd = reinterpret_cast<const uint4*>(fdesc + (size_t)rf * 32);
            const unsigned key = ((unsigned)bow_dist(q0, q1, fd[0], fd[1]) << 16) | (unsigned)(j - j0);
            const unsigned t = max(key, b1); b1 = min(b1, key); b2 = min(b2, t);
        }
        // warp top-2 of the union of the lanes' top-2
        const unsigned m1 = __reduce_min_sync(0xffffffffu, b1);
        const unsigned c2 = b1 == m1 ? b2 : b1;                              // keys are unique (list position)
        const unsigned m2 = __reduce_min_sync(0xffffffffu, c2);
        const int bestDist1 = m1 == 0xffffffffu ? 256 : (int)(m1 >> 16);
        const int bestDist2 = m2 == 0xffffffffu ? 256 : (int)(m2 >> 16);
        // `<= TH_LOW` in the KeyFrame-Frame form (:267), `< TH_LOW` in the KeyFrame-KeyFrame form (:671)
        if ((A.kf_mode ? bestDist1 < A.th_low : bestDist1 <= A.th_low) && (float)bestDist1 < __fmul_rn(A.nnratio, (float)bestDist2)) {
            const int bestIdxF = ffeat[j0 + (int)(m1 & 0xffffu)];
            if (lane == 0) {
                taken[bestIdxF] = ri;
                if (A.kf_mode) match[ri] = bestIdxF;
                if (A.check_orientation) {
                    float rot = __fsub_rn(kkp[ri].angle, fkp[bestIdxF].angle);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                    if (bin == 30) bin = 0;
                    bin_of[A.kf_mode ? ri : bestIdxF] = bin;
                    atomicAdd(A.hist + pair * 32 + bin, 1);
                }
                atomicAdd(A.nmatches + pair, 1);
            }
            __syncwarp();
            __threadfence_block();
        }
    }
}

// ------------------------------------------------------------------------------------------- SearchForTriangulation
// ORBmatcher::SearchForTriangulation (ORBmatcher.cc:738-916): features WITHOUT a map point of two keyframes, matched inside
// common FeatureVector nodes under the epipolar constraint of F12. The reference never sets vbMatched2, so a keyframe-1
// feature's result depends on nothing but the candidates themselves: the winner is the candidate with the smallest
// distance <= TH_LOW among those that pass the stateless gates (map-point / stereo flags, distance to the epipole,
// CheckDistEpipolarLine :153-173), and the LAST such candidate in list order on ties (`dist > bestDist` skips, an equal
// distance replaces). One warp per node of keyframe 1, candidates over the lanes, key = dist << 16 | (0xffff - position).
__global__ void __launch_bounds__(128) bow_triangulation_kernel(OrbxBowOut O, OrbxBowTriArgs T, int cap)
{
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const int f1 = T.kf1_frame[pair], f2 = T.kf2_frame[pair];
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= O.n_fv[f1]) return;
    const int node = O.fv_node[(size_t)f1 * cap + a];
    const int* node2 = O.fv_node + (size_t)f2 * cap;
    int lo = 0, hi = O.n_fv[f2];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (node2[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= O.n_fv[f2] || node2[lo] != node) return;
    const int* off1 = O.fv_off + (size_t)f1 * (cap + 1);
    const int* off2 = O.fv_off + (size_t)f2 * (cap + 1);
    const int* feat1 = O.fv_feat + (size_t)f1 * cap;
    const int* feat2 = O.fv_feat + (size_t)f2 * cap;
    const int k0 = off1[a], k1 = off1[a + 1], j0 = off2[lo], j1 = off2[lo + 1];
    const uint8_t* desc1 = T.desc + (size_t)f1 * cap * 32;
    const uint8_t* desc2 = T.desc + (size_t)f2 * cap * 32;
    const OrbxKp28* kp1s = T.kps + (size_t)f1 * cap;
    const OrbxKp28* kp2s = T.kps + (size_t)f2 * cap;
    const uint8_t* mp1 = T.has_mp ? T.has_mp + (size_t)f1 * cap : nullptr;
    const uint8_t* mp2 = T.has_mp ? T.has_mp + (size_t)f2 * cap : nullptr;
    const float* ur1 = T.u_right ? T.u_right + (size_t)f1 * cap : nullptr;
    const float* ur2 = T.u_right ? T.u_right + (size_t)f2 * cap : nullptr;
    const float* g = T.geom + (size_t)pair * 28;                             // F12 (9), Cw1 (3), R2w (9), t2w (3), K2 (4)
    // epipole of camera 1 in image 2 (:744-754): C2 = R2w*Cw + t2w as one f32 gemm, invz = 1.0f / C2.z
    float C2[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float s = __fmul_rn(g[12 + 3 * r], g[9]);
        s = __fadd_rn(s, __fmul_rn(g[12 + 3 * r + 1], g[10]));
        s = __fadd_rn(s, __fmul_rn(g[12 + 3 * r + 2], g[11]));
        C2[r] = __fadd_rn(s, g[21 + r]);
    }
    const float invz = __fdiv_rn(1.0f, C2[2]);
    const float ex = __fadd_rn(__fmul_rn(__fmul_rn(g[24], C2[0]), invz), g[26]);
    const float ey = __fadd_rn(__fmul_rn(__fmul_rn(g[25], C2[1]), invz), g[27]);
    int* match = T.match + (size_t)pair * cap;
    int* bin_of = T.bin_of + (size_t)pair * cap;
    for (int ik = k0; ik < k1; ik++) {
        const int idx1 = feat1[ik];
        if (mp1 && mp1[idx1]) continue;                                       // warp-uniform
        const bool stereo1 = ur1 ? ur1[idx1] >= 0.f : false;
        if (T.only_stereo && !stereo1) continue;
        const OrbxKp28 kp1 = kp1s[idx1];
        // the epipolar line of kp1 in image 2
        const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g[0]), __fmul_rn(kp1.y, g[3])), g[6]);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g[1]), __fmul_rn(kp1.y, g[4])), g[7]);
        const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g[2]), __fmul_rn(kp1.y, g[5])), g[8]);
        const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        const uint4* d1 = reinterpret_cast<const uint4*>(desc1 + (size_t)idx1 * 32);
        const uint4 q0 = __ldg(d1), q1 = __ldg(d1 + 1);
        unsigned best = 0xffffffffu;
        for (int j = j0 + lane; j < j1; j += 32) {
            const int idx2 = feat2[j];
            if (mp2 && mp2[idx2]) continue;
            const bool stereo2 = ur2 ? ur2[idx2] >= 0.f : false;
            if (T.only_stereo && !stereo2) continue;
            const uint4* d2 = reinterpret_cast<const uint4*>(desc2 + (size_t)idx2 * 32);
            const int dist = bow_dist(q0, q1, d2[0], d2[1]);
            if (dist > T.th_low) continue;
            const OrbxKp28 kp2 = kp2s[idx2];
            if (!stereo1 && !stereo2) {
                const float dx = __fsub_rn(ex, kp2.x), dy = __fsub_rn(ey, kp2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, T.scale_factors[kp2.octave])) continue;
            }
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
            if (den == 0.f) continue;
            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
            if (!((double)dsqr < __dmul_rn(3.84, (double)T.level_sigma2[kp2.octave]))) continue;
            best = min(best, ((unsigned)dist << 16) | (unsigned)(0xffff - (j - j0)));
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best != 0xffffffffu && lane == 0) {
            const int idx2 = feat2[j0 + (0xffff - (int)(best & 0xffffu))];
            match[idx1] = idx2;
            if (T.check_orientation) {
                float rot = __fsub_rn(kp1.angle, kp2s[idx2].angle);
                if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                if (bin == 30) bin = 0;
                bin_of[idx1] = bin;
                atomicAdd(T.hist + pair * 32 + bin, 1);
            }
            atomicAdd(T.nmatches + pair, 1);
        }
    }
}

// ComputeThreeMaxima + rejection of the matches outside the three dominant rotation bins (ORBmatcher.cc:296-322, 1797-1839)
__global__ void __launch_bounds__(256) bow_rotation_kernel(OrbxBowMatchArgs A, const int* __restrict__ d_n, int cap)
{
    __shared__ int ind[3];
    __shared__ int removed;
    const int pair = blockIdx.x;
    if (threadIdx.x == 0) {
        const int* h = A.hist + pair * 32;
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < 30; i++) {
            const int s = h[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
        ind[0] = i1; ind[1] = i2; ind[2] = i3; removed = 0;
    }
    __syncthreads();
    const int n = min(d_n[A.kf_mode ? A.kf_frame[pair] : A.f_frame[pair]], cap);   // rotHist holds idx1 in kf_mode (:687)
    int* match = A.match + (size_t)pair * cap;
    const int* bin_of = A.bin_of + (size_t)pair * cap;
    int mine = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        if (match[j] >= 0) {
            const int b = bin_of[j];
            if (b != ind[0] && b != ind[1] && b != ind[2]) { match[j] = -1; mine++; }
        }
    if (mine) atomicAdd(&removed, mine);
    __syncthreads();
    if (threadIdx.x == 0) A.nmatches[pair] -= removed;
}

__global__ void bow_match_init_kernel(OrbxBowMatchArgs A, int cap, int npairs)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)npairs * cap) { A.match[i] = -1; if (A.kf_mode) A.taken[i] = -1; }
    if (i < (size_t)npairs * 32) A.hist[i] = 0;
    if (i < (size_t)npairs) A.nmatches[i] = 0;
}

// ---------------------------------------------------------------------------------------------------- launchers
void orbx_launch_bow_transform(const OrbxVocabDev& V, const uint8_t* d_desc, const int* d_n, int frames, int cap, int levelsup,
                               const OrbxBowOut& O, cudaStream_t st)
{
    if (frames <= 0 || cap <= 0) return;
    dim3 g1((cap + 7) / 8, frames);
    bow_descend_kernel<<<g1, 256, 0, st>>>(V, d_desc, d_n, cap, levelsup, O.leaf, O.nid);
    int P = 32; while (P < cap) P <<= 1;
    const size_t smem = (size_t)P * sizeof(unsigned long long);
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(bow_assemble_kernel, mark[0], smem);
    bow_assemble_kernel<<<frames, 512, smem, st>>>(V, O, d_n, cap, P);
}

void orbx_launch_bow_score(const OrbxBowOut& O, int cap, const int* d_qa, const int* d_qb, int npairs, double* d_score, cudaStream_t st)
{
    if (npairs <= 0) return;
    bow_score_l1_kernel<<<(npairs + 3) / 4, 128, 0, st>>>(O, cap, d_qa, d_qb, npairs, d_score);
}

void orbx_launch_bow_triangulation(const OrbxBowOut& O, const OrbxBowTriArgs& T, int* d_taken, const int* d_n, int cap, int npairs,
                                   cudaStream_t st)
{
    if (npairs <= 0 || cap <= 0) return;
    OrbxBowMatchArgs A = {};                                                 // init / rotation kernels of the KeyFrame-KeyFrame form
    A.kf_frame = T.kf1_frame; A.f_frame = T.kf2_frame; A.kf_mode = 1; A.check_orientation = T.check_orientation;
    A.match = T.match; A.bin_of = T.bin_of; A.taken = d_taken; A.hist = T.hist; A.nmatches = T.nmatches;
    const size_t tot = (size_t)npairs * std::max(cap, 32);
    bow_match_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(A, cap, npairs);
    dim3 g((cap + 3) / 4, npairs);
    bow_triangulation_kernel<<<g, 128, 0, st>>>(O, T, cap);
    if (T.check_orientation) bow_rotation_kernel<<<npairs, 256, 0, st>>>(A, d_n, cap);
}

void orbx_launch_bow_match(const OrbxBowOut& O, const OrbxBowMatchArgs& A, const int* d_n, int cap, int npairs, cudaStream_t st)
{
    if (npairs <= 0 || cap <= 0) return;
    const size_t tot = (size_t)npairs * std::max(cap, 32);
    bow_match_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(A, cap, npairs);
    dim3 g((cap + 3) / 4, npairs);
    bow_match_kernel<<<g, 128, 0, st>>>(O, A, cap);
    if (A.check_orientation) bow_rotation_kernel<<<npairs, 256, 0, st>>>(A, d_n, cap);
}

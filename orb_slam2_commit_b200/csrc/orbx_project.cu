// orbx_project.cu — ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
// (ORBmatcher.cc:1489-1646), the matcher of Tracking::TrackWithMotionModel: every map point seen in the last frame is
// projected into the current frame, Frame::GetFeaturesInArea (Frame.cc:388-444) collects the keypoints of the window on
// the 64x48 grid and the nearest descriptor (<= TH_HIGH) takes the map point; matches outside the three dominant
// rotation bins are dropped (:1613-1642, ComputeThreeMaxima :1797-1839).
//
// The current frame's grid (Frame::mGrid) is rebuilt in shared memory as a CSR by a counting sort, so a query only
// visits the cells of its window, in GetFeaturesInArea's own order; one thread per last-frame keypoint.
// The reference loop is sequential in one respect: a current keypoint that already holds a map point with
// Observations() > 0 is skipped by LATER last-frame keypoints (:1574-1576). One CTA per frame pair resolves that with a
// fixed-point iteration: every round all queries pick their best keypoint among those not taken by an EARLIER query
// (taker[k] = lowest query index holding k with an observed map point, rebuilt from the previous round's picks). A
// round that changes nothing is the sequential result (induction over the query index), and query i is final after
// round i at the latest; on real frames two or three rounds suffice.
// Projection arithmetic: OpenCV 4.13's small-matrix gemm for `Rcw*x3Dw+tcw` (f32, left to right, addend last), a
// double division for 1.0/z, un-contracted f32 everywhere else (the reference is built without FMA contraction).
#include "orbx_internal.cuh"
#include <algorithm>

struct ProjKp { float x, y; int octave; };
typedef OrbxProjQuery ProjQuery;                               // r < 0: this last-frame keypoint makes no query

#define PROJ_CELLS (64 * 48)                                   // FRAME_GRID_COLS x FRAME_GRID_ROWS (Frame.h)

__device__ __forceinline__ int proj_dist(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// One CTA per frame pair. Shared memory holds Frame::mGrid of the current frame as a CSR (cell -> keypoint indices in
// ascending order, exactly what AssignFeaturesToGrid builds), so a query visits only the cells of its window, in
// GetFeaturesInArea's order (cell column, cell row, index): one THREAD per query, strict '<' keeps the first minimum.
__global__ void __launch_bounds__(512) search_projection_kernel(const OrbxProjPairDev* __restrict__ pairs, OrbxProjCam cam,
                                                                const float* __restrict__ scale_factors, float th,
                                                                int check_orientation, int th_high)
{
    extern __shared__ __align__(16) unsigned char s_raw3[];
    const OrbxProjPairDev P = pairs[blockIdx.x];
    ProjKp* sk = reinterpret_cast<ProjKp*>(s_raw3);                              // [n_cur]
    int* taker = reinterpret_cast<int*>(sk + P.n_cur);                          // [n_cur]
    unsigned short* order = reinterpret_cast<unsigned short*>(taker + P.n_cur); // [n_cur] keypoint indices sorted by cell
    unsigned short* cstart = order + ((P.n_cur + 1) & ~1);                      // [PROJ_CELLS + 1]
    unsigned short* cfill = cstart + PROJ_CELLS + 2;                            // [PROJ_CELLS] counts, then fill cursors
    __shared__ int s_changed, s_hist[32], s_ind[3], s_success, s_removed, s_w[17];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float invW = __fdiv_rn(64.0f, __fsub_rn(cam.maxX, cam.minX)), invH = __fdiv_rn(48.0f, __fsub_rn(cam.maxY, cam.minY));

    // ---- current frame: Frame::AssignFeaturesToGrid / PosInGrid as a counting sort
    for (int c = tid; c < PROJ_CELLS; c += blockDim.x) cfill[c] = 0;
    __syncthreads();
    for (int i = tid; i < P.n_cur; i += blockDim.x) {
        const OrbxKp28 k = P.cur_kps[i];
        ProjKp e; e.x = k.x; e.y = k.y; e.octave = k.octave;
        sk[i] = e;
        taker[i] = 0x7fffffff;
        P.match[i] = -1;
        const int posX = (int)roundf(__fmul_rn(__fsub_rn(k.x, cam.minX), invW));
        const int posY = (int)roundf(__fmul_rn(__fsub_rn(k.y, cam.minY), invH));
        if (!(posX < 0 || posX >= 64 || posY < 0 || posY >= 48)) {
            // 16-bit counters packed two per word: atomicAdd on the containing word
            const int c = posX * 48 + posY;
            atomicAdd(reinterpret_cast<unsigned*>(cfill) + (c >> 1), (c & 1) ? 0x10000u : 1u);
        }
    }
    __syncthreads();
    {   // exclusive scan of the cell counts: 6 cells per thread, then a block scan of the per-thread sums
        const int c0 = tid * 6;
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
        if (tid == 0) cstart[PROJ_CELLS] = (unsigned short)s_w[16];
    }
    __syncthreads();
    for (int i = tid; i < P.n_cur; i += blockDim.x) {
        const ProjKp k = sk[i];
        const int posX = (int)roundf(__fmul_rn(__fsub_rn(k.x, cam.minX), invW));
        const int posY = (int)roundf(__fmul_rn(__fsub_rn(k.y, cam.minY), invH));
        if (!(posX < 0 || posX >= 64 || posY < 0 || posY >= 48)) {
            const int c = posX * 48 + posY;
            const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(cfill) + (c >> 1), (c & 1) ? 0x10000u : 1u);
            order[(c & 1) ? (old >> 16) : (old & 0xffffu)] = (unsigned short)i;
        }
    }
    __syncthreads();
    // the atomics filled every cell in arbitrary order: the reference's cells hold ascending indices
    for (int c = tid; c < PROJ_CELLS; c += blockDim.x) {
        const int b0 = cstart[c], b1 = cstart[c + 1];
        for (int i = b0 + 1; i < b1; i++) {
            const unsigned short v = order[i];
            int j = i - 1;
            while (j >= b0 && order[j] > v) { order[j + 1] = order[j]; j--; }
            order[j + 1] = v;
        }
    }
    // ---- last frame: projection of every map point (:1524-1567)
    for (int i = tid; i < P.n_last; i += blockDim.x) {
        ProjQuery q; q.r = -1.f; q.u = q.v = q.ur = 0.f; q.min_level = q.max_level = -1;
        if (P.last_flags[i] & 1) {
            const float X = P.last_xyz[3 * i], Y = P.last_xyz[3 * i + 1], Z = P.last_xyz[3 * i + 2];
            float c3[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                float s = __fmul_rn(P.Tcw[3 * r], X);
                s = __fadd_rn(s, __fmul_rn(P.Tcw[3 * r + 1], Y));
                s = __fadd_rn(s, __fmul_rn(P.Tcw[3 * r + 2], Z));
                c3[r] = __fadd_rn(s, P.Tcw[9 + r]);
            }
            const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)c3[2]));
            if (!(invzc < 0)) {
                const float u = __fadd_rn(__fmul_rn(__fmul_rn(cam.fx, c3[0]), invzc), cam.cx);
                const float v = __fadd_rn(__fmul_rn(__fmul_rn(cam.fy, c3[1]), invzc), cam.cy);
                if (!(u < cam.minX || u > cam.maxX) && !(v < cam.minY || v > cam.maxY)) {
                    const int oct = P.last_kps[i].octave;
                    q.u = u; q.v = v; q.r = __fmul_rn(th, scale_factors[oct]);
                    q.ur = __fsub_rn(u, __fmul_rn(cam.mbf, invzc));
                    if (P.mode == 1) { q.min_level = oct; q.max_level = -1; }
                    else if (P.mode == 2) { q.min_level = 0; q.max_level = oct; }
                    else { q.min_level = oct - 1; q.max_level = oct + 1; }
                }
            }
        }
        P.query[i] = q;
        P.assign[i] = -1;
    }
    if (tid < 32) s_hist[tid] = 0;
    if (tid == 0) { s_success = 0; s_removed = 0; }
    __syncthreads();

    const uint4* cdesc = reinterpret_cast<const uint4*>(P.cur_desc);
    const uint4* ldesc = reinterpret_cast<const uint4*>(P.last_desc);
    for (int round = 0; round <= P.n_last; round++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int qi = tid; qi < P.n_last; qi += blockDim.x) {
            const ProjQuery q = P.query[qi];
            if (q.r < 0.f) continue;
            const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.u, cam.minX), q.r), invW)));
            const int cx1 = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.u, cam.minX), q.r), invW)));
            const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.v, cam.minY), q.r), invH)));
            const int cy1 = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.v, cam.minY), q.r), invH)));
            int bestDist = 256, bestIdx = -1;
            if (!(cx0 >= 64 || cx1 < 0 || cy0 >= 48 || cy1 < 0)) {
                const bool check_levels = q.min_level > 0 || q.max_level >= 0;
                const uint4 qa = ldesc[2 * (size_t)qi], qb = ldesc[2 * (size_t)qi + 1];
                for (int ix = cx0; ix <= cx1; ix++) {
                    // the cells of one column are adjacent in the CSR: rows cy0..cy1 form one contiguous range
                    const int j0 = cstart[ix * 48 + cy0], j1 = cstart[ix * 48 + cy1 + 1];
                    for (int j = j0; j < j1; j++) {
                        const int i = order[j];
                        const ProjKp k = sk[i];
                        if (check_levels) {
                            if (k.octave < q.min_level) continue;
                            if (q.max_level >= 0 && k.octave > q.max_level) continue;
                        }
                        if (!(fabsf(__fsub_rn(k.x, q.u)) < q.r && fabsf(__fsub_rn(k.y, q.v)) < q.r)) continue;
                        if ((P.cur_occupied && P.cur_occupied[i]) || taker[i] < qi) continue;
                        if (P.cur_u_right) { const float ur = P.cur_u_right[i]; if (ur > 0.f && fabsf(__fsub_rn(q.ur, ur)) > q.r) continue; }
                        const int d = proj_dist(qa, qb, cdesc[2 * (size_t)i], cdesc[2 * (size_t)i + 1]);
                        if (d < bestDist) { bestDist = d; bestIdx = i; }
                    }
                }
            }
            const int a = bestDist <= th_high ? bestIdx : -1;
            if (a != P.assign[qi]) { P.assign[qi] = a; s_changed = 1; }
        }
        __syncthreads();
        if (!s_changed) break;
        for (int i = tid; i < P.n_cur; i += blockDim.x) taker[i] = 0x7fffffff;
        __syncthreads();
        for (int qi = tid; qi < P.n_last; qi += blockDim.x) {
            const int a = P.assign[qi];
            if (a >= 0 && (P.last_flags[qi] & 2)) atomicMin(&taker[a], qi);
        }
        __syncthreads();
    }

    // ---- CurrentFrame.mvpMapPoints[bestIdx2] = pMP in query order (the last writer stays), rotation histogram
    for (int qi = threadIdx.x; qi < P.n_last; qi += blockDim.x) {
        const int a = P.assign[qi];
        if (a < 0) continue;
        atomicMax(&P.match[a], qi);
        atomicAdd(&s_success, 1);
        if (check_orientation) {
            float rot = __fsub_rn(P.last_kps[qi].angle, P.cur_kps[a].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
            if (bin == 30) bin = 0;
            P.query[qi].min_level = bin;                                      // reuse: rotation bin of this match
            atomicAdd(&s_hist[bin], 1);
        }
    }
    __syncthreads();
    if (check_orientation) {
        if (threadIdx.x == 0) {
            int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < 30; i++) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
                else if (s > max3) { max3 = s; i3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
            s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3;
        }
        __syncthreads();
        for (int qi = threadIdx.x; qi < P.n_last; qi += blockDim.x) {
            const int a = P.assign[qi];
            if (a < 0) continue;
            const int bin = P.query[qi].min_level;
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { P.match[a] = -1; atomicAdd(&s_removed, 1); }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *P.nmatches = s_success - s_removed;
}

void orbx_launch_search_projection(const OrbxProjPairDev* d_pairs, int npairs, int max_n_cur, const OrbxProjCam& cam,
                                   const float* d_scale_factors, float th, int check_orientation, cudaStream_t st)
{
    if (npairs <= 0) return;
    const size_t n1 = (size_t)std::max(max_n_cur, 1);
    const size_t smem = n1 * (sizeof(ProjKp) + sizeof(int)) + ((n1 + 1) & ~(size_t)1) * 2 + (size_t)(2 * PROJ_CELLS + 4) * 2 + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(search_projection_kernel, mark[0], smem);
    search_projection_kernel<<<npairs, 512, smem, st>>>(d_pairs, cam, d_scale_factors, th, check_orientation, 100 /* TH_HIGH */);
}

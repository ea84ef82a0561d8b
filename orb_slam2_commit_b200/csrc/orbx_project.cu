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
// double division for 1.0/z, un-contracted f32 everywhere else (the pinned reference semantics).
#include "orbx_grid.cuh"
#include <algorithm>

typedef GridKp ProjKp;
typedef OrbxProjQuery ProjQuery;                               // r < 0: this last-frame keypoint makes no query

// One CTA per frame pair. Shared memory holds Frame::mGrid of the current frame as a CSR (cell -> keypoint indices in
// ascending order, exactly what AssignFeaturesToGrid builds), so a query visits only the cells of its window, in
// GetFeaturesInArea's order (cell column, cell row, index): one THREAD per query, strict '<' keeps the first minimum.
__global__ void __launch_bounds__(512) search_projection_kernel(const OrbxProjPairDev* __restrict__ pairs, OrbxProjCam cam,
                                                                const float* __restrict__ scale_factors, float th,
                                                                int check_orientation, int th_high)
{
    extern __shared__ __align__(16) unsigned char s_raw3[];
    const OrbxProjPairDev P = pairs[blockIdx.x];
    GridSmem G;
    int* taker = reinterpret_cast<int*>(grid_carve(s_raw3, P.n_cur, G));        // [n_cur]
    const ProjKp* sk = G.kp;
    const unsigned short* order = G.order;
    const unsigned short* cstart = G.cstart;
    __shared__ int s_changed, s_hist[32], s_ind[3], s_success, s_removed, s_w[17];
    const int tid = threadIdx.x;
    const float invW = __fdiv_rn(64.0f, __fsub_rn(cam.maxX, cam.minX)), invH = __fdiv_rn(48.0f, __fsub_rn(cam.maxY, cam.minY));

    // ---- current frame: Frame::AssignFeaturesToGrid / PosInGrid as a counting sort
    for (int i = tid; i < P.n_cur; i += blockDim.x) { taker[i] = 0x7fffffff; P.match[i] = -1; }
    grid_build(P.cur_kps, P.n_cur, cam.minX, cam.minY, invW, invH, G, s_w);
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
                        const int d = grid_hamming(qa, qb, cdesc[2 * (size_t)i], cdesc[2 * (size_t)i + 1]);
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
    const size_t smem = grid_smem_bytes(n1) + n1 * sizeof(int) + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(search_projection_kernel, mark[0], smem);
    search_projection_kernel<<<npairs, 512, smem, st>>>(d_pairs, cam, d_scale_factors, th, check_orientation, 100 /* TH_HIGH */);
}

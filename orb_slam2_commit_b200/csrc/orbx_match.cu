// orbx_match.cu — the map-point-to-keypoint matchers of ORBmatcher that walk GetFeaturesInArea windows:
//   * SearchByProjection(Frame &F, const vector<MapPoint*>&, th)  (ORBmatcher.cc:46-142), the matcher of
//     Tracking::SearchLocalPoints: one query per local map point that Frame::isInFrustum marked mbTrackInView;
//   * Fuse(KeyFrame*, const vector<MapPoint*>&, th) (ORBmatcher.cc:918-1092) and the Sim3 form
//     Fuse(KeyFrame*, cv::Mat Scw, ...) (:1094-1236): projection of map points into a keyframe, the distance / viewing
//     angle / scale-prediction gates and the nearest descriptor in the window — the search half; the map surgery that
//     follows (Replace / AddObservation) stays with the caller.
// Both rebuild the frame's 64x48 grid in shared memory as a CSR (orbx_grid.cuh) and give one THREAD per map point its
// window walk in GetFeaturesInArea's order, so the reference's strict-'<' update rules run verbatim.
//
// SearchByProjection is sequential in one respect: a keypoint that holds a map point with Observations() > 0 — from
// before the call or assigned EARLIER in the loop (:90-92, :139) — is skipped. One CTA per frame resolves that with the
// fixed-point iteration of orbx_project.cu: taker[k] = lowest query index that took k with an observed map point; every
// round all queries pick among the keypoints no earlier query holds; a round that changes nothing is the sequential
// result (induction over the query index).
#include "orbx_grid.cuh"
#include <algorithm>

// ------------------------------------------------------------------------------------------- SearchByProjection(F, MPs)
__global__ void __launch_bounds__(512) local_points_kernel(const OrbxLocalFrameDev* __restrict__ frames, float minX, float maxX,
                                                           float minY, float maxY, const float* __restrict__ scale_factors,
                                                           int nlevels, float th, float nnratio, int th_high)
{
    extern __shared__ __align__(16) unsigned char s_raw4[];
    const OrbxLocalFrameDev P = frames[blockIdx.x];
    GridSmem G;
    int* taker = reinterpret_cast<int*>(grid_carve(s_raw4, P.n, G));
    __shared__ int s_changed, s_success, s_w[17];
    const int tid = threadIdx.x;
    const float invW = __fdiv_rn(64.0f, __fsub_rn(maxX, minX)), invH = __fdiv_rn(48.0f, __fsub_rn(maxY, minY));
    const bool bFactor = (double)th != 1.0;                                      // :50

    for (int i = tid; i < P.n; i += blockDim.x) { taker[i] = 0x7fffffff; P.match[i] = -1; }
    for (int i = tid; i < P.nq; i += blockDim.x) P.assign[i] = -1;
    if (tid == 0) s_success = 0;
    grid_build(P.kps, P.n, minX, minY, invW, invH, G, s_w);

    const uint4* kdesc = reinterpret_cast<const uint4*>(P.desc);
    const uint4* qdesc = reinterpret_cast<const uint4*>(P.qdesc);
    for (int round = 0; round <= P.nq; round++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int qi = tid; qi < P.nq; qi += blockDim.x) {
            if (!(P.qflags[qi] & 1)) continue;                                   // !mbTrackInView || isBad()  (:55-59)
            const OrbxTrackQueryDev q = P.q[qi];
            const int level = min(max(q.level, 0), nlevels - 1);
            float r = (double)q.view_cos > 0.998 ? 2.5f : 4.0f;                  // RadiusByViewingCos (:144-150)
            if (bFactor) r = __fmul_rn(r, th);
            r = __fmul_rn(r, scale_factors[level]);
            const int min_level = level - 1, max_level = level;
            int cx0, cx1, cy0, cy1;
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            if (grid_window(q.x, q.y, r, minX, minY, invW, invH, cx0, cx1, cy0, cy1)) {
                const bool check_levels = min_level > 0 || max_level >= 0;
                const uint4 qa = qdesc[2 * (size_t)qi], qb = qdesc[2 * (size_t)qi + 1];
                for (int ix = cx0; ix <= cx1; ix++) {
                    const int j0 = G.cstart[ix * GRID_ROWS + cy0], j1 = G.cstart[ix * GRID_ROWS + cy1 + 1];
                    for (int j = j0; j < j1; j++) {
                        const int i = G.order[j];
                        const GridKp k = G.kp[i];
                        if (check_levels) {
                            if (k.octave < min_level) continue;
                            if (max_level >= 0 && k.octave > max_level) continue;
                        }
                        if (!(fabsf(__fsub_rn(k.x, q.x)) < r && fabsf(__fsub_rn(k.y, q.y)) < r)) continue;
                        if ((P.occupied && P.occupied[i]) || taker[i] < qi) continue;                       // :90-92
                        if (P.u_right) { const float ur = P.u_right[i]; if (ur > 0.f && fabsf(__fsub_rn(q.xr, ur)) > r) continue; }
                        const int d = grid_hamming(qa, qb, kdesc[2 * (size_t)i], kdesc[2 * (size_t)i + 1]);
                        if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = k.octave; bestIdx = i; }
                        else if (d < bestDist2) { bestLevel2 = k.octave; bestDist2 = d; }
                    }
                }
            }
            int a = -1;
            if (bestDist <= th_high && !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2))) a = bestIdx;
            if (a != P.assign[qi]) { P.assign[qi] = a; s_changed = 1; }
        }
        __syncthreads();
        if (!s_changed) break;
        for (int i = tid; i < P.n; i += blockDim.x) taker[i] = 0x7fffffff;
        __syncthreads();
        for (int qi = tid; qi < P.nq; qi += blockDim.x) {
            const int a = P.assign[qi];
            if (a >= 0 && (P.qflags[qi] & 2)) atomicMin(&taker[a], qi);
        }
        __syncthreads();
    }
    // F.mvpMapPoints[bestIdx] = pMP in query order: the last writer stays (:139)
    int mine = 0;
    for (int qi = tid; qi < P.nq; qi += blockDim.x) {
        const int a = P.assign[qi];
        if (a >= 0) { atomicMax(&P.match[a], qi); mine++; }
    }
    if (mine) atomicAdd(&s_success, mine);
    __syncthreads();
    if (tid == 0) *P.nmatches = s_success;
}

void orbx_launch_local_points(const OrbxLocalFrameDev* d_frames, int nframes, int max_n, const float* bounds4,
                              const float* d_scale_factors, int nlevels, float th, float nnratio, cudaStream_t st)
{
    if (nframes <= 0) return;
    const size_t n1 = (size_t)std::max(max_n, 1);
    const size_t smem = grid_smem_bytes(n1) + n1 * sizeof(int) + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(local_points_kernel, mark[0], smem);
    local_points_kernel<<<nframes, 512, smem, st>>>(d_frames, bounds4[0], bounds4[1], bounds4[2], bounds4[3], d_scale_factors,
                                                    nlevels, th, nnratio, 100 /* TH_HIGH */);
}

// ------------------------------------------------------------------------------------------------------------- Fuse
// One CTA per (keyframe, map-point list). Thread per map point: projection with OpenCV's small-matrix gemm order
// (`Rcw*p3Dw + tcw`: f32 products summed left to right, the addend last), `1/z` as an f32 division, cv::norm and
// Mat::dot accumulated in f64 in element order, un-contracted f32 elsewhere (the reference is built without FMA
// contraction). MapPoint::PredictScale's `ceil(logf(ratio) / mfLogScaleFactor)` is evaluated through a threshold table
// built on the host with the host's own logf (level_ratio[n] = smallest ratio whose predicted level exceeds n), so the
// level equals the host libm's bit for bit.
__global__ void __launch_bounds__(512) fuse_search_kernel(const OrbxFuseDev* __restrict__ jobs, OrbxFuseCam cam, int th_low)
{
    extern __shared__ __align__(16) unsigned char s_raw5[];
    const OrbxFuseDev P = jobs[blockIdx.x];
    GridSmem G;
    grid_carve(s_raw5, P.n, G);
    __shared__ int s_w[17];
    __shared__ int s_found;
    const int tid = threadIdx.x;
    const float invW = __fdiv_rn(64.0f, __fsub_rn(cam.maxX, cam.minX)), invH = __fdiv_rn(48.0f, __fsub_rn(cam.maxY, cam.minY));
    if (tid == 0) s_found = 0;
    grid_build(P.kps, P.n, cam.minX, cam.minY, invW, invH, G, s_w);
    const uint4* kdesc = reinterpret_cast<const uint4*>(P.desc);
    const uint4* pdesc = reinterpret_cast<const uint4*>(P.pt_desc);
    int mine = 0;
    for (int pi = tid; pi < P.npts; pi += blockDim.x) {
        int bestDist = P.mode == 0 ? 256 : 0x7fffffff, bestIdx = -1;                 // :993 / :1180 (INT_MAX in the Sim3 form)
        do {
            if (!(P.pt_flags[pi] & 1)) break;                                    // NULL, isBad() or IsInKeyFrame(pKF)
            const float X = P.pt_xyz[3 * pi], Y = P.pt_xyz[3 * pi + 1], Z = P.pt_xyz[3 * pi + 2];
            float c3[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                float s = __fmul_rn(P.Tcw[3 * r], X);
                s = __fadd_rn(s, __fmul_rn(P.Tcw[3 * r + 1], Y));
                s = __fadd_rn(s, __fmul_rn(P.Tcw[3 * r + 2], Z));
                c3[r] = __fadd_rn(s, P.Tcw[9 + r]);
            }
            if (c3[2] < 0.0f) break;
            // `1/z` (:954, f32 division) in Fuse(pKF, vpMapPoints); `1.0/z` (:1146, f64 division rounded to f32) in the Sim3 form
            const float invz = P.mode == 0 ? __fdiv_rn(1.0f, c3[2]) : __double2float_rn(__ddiv_rn(1.0, (double)c3[2]));
            const float x = __fmul_rn(c3[0], invz), y = __fmul_rn(c3[1], invz);
            const float u = __fadd_rn(__fmul_rn(cam.fx, x), cam.cx), v = __fadd_rn(__fmul_rn(cam.fy, y), cam.cy);
            if (!(u >= cam.minX && u < cam.maxX && v >= cam.minY && v < cam.maxY)) break;      // KeyFrame::IsInImage
            const float ur = __fsub_rn(u, __fmul_rn(cam.bf, invz));
            const float po0 = __fsub_rn(X, P.Ow[0]), po1 = __fsub_rn(Y, P.Ow[1]), po2 = __fsub_rn(Z, P.Ow[2]);
            double s2 = __dmul_rn((double)po0, (double)po0);
            s2 = __dadd_rn(s2, __dmul_rn((double)po1, (double)po1));
            s2 = __dadd_rn(s2, __dmul_rn((double)po2, (double)po2));
            const float dist3D = __double2float_rn(__dsqrt_rn(s2));
            if (dist3D < P.pt_dist[3 * pi] || dist3D > P.pt_dist[3 * pi + 1]) break;            // min / max distance invariance
            const float* nrm = P.pt_normal + 3 * pi;
            double dot = __dmul_rn((double)po0, (double)nrm[0]);
            dot = __dadd_rn(dot, __dmul_rn((double)po1, (double)nrm[1]));
            dot = __dadd_rn(dot, __dmul_rn((double)po2, (double)nrm[2]));
            if (dot < __dmul_rn(0.5, (double)dist3D)) break;
            const float ratio = __fdiv_rn(P.pt_dist[3 * pi + 2], dist3D);                         // mfMaxDistance / currentDist
            int level = 0;
            for (int n = 0; n < cam.nlevels - 1; n++) level += (ratio >= cam.level_ratio[n]) ? 1 : 0;
            const float radius = __fmul_rn(P.th, cam.scale_factors[level]);
            int cx0, cx1, cy0, cy1;
            if (!grid_window(u, v, radius, cam.minX, cam.minY, invW, invH, cx0, cx1, cy0, cy1)) break;
            const uint4 qa = pdesc[2 * (size_t)pi], qb = pdesc[2 * (size_t)pi + 1];
            for (int ix = cx0; ix <= cx1; ix++) {
                const int j0 = G.cstart[ix * GRID_ROWS + cy0], j1 = G.cstart[ix * GRID_ROWS + cy1 + 1];
                for (int j = j0; j < j1; j++) {
                    const int i = G.order[j];
                    const GridKp k = G.kp[i];
                    if (!(fabsf(__fsub_rn(k.x, u)) < radius && fabsf(__fsub_rn(k.y, v)) < radius)) continue;
                    if (k.octave < level - 1 || k.octave > level) continue;
                    if (P.mode == 0) {                                           // chi-square gate of Fuse(pKF, vpMapPoints) (:1016-1046)
                        const float ex = __fsub_rn(u, k.x), ey = __fsub_rn(v, k.y);
                        const float kr = P.u_right ? P.u_right[i] : -1.0f;
                        const float inv_s2 = cam.inv_level_sigma2[k.octave];
                        if (kr >= 0.f) {
                            const float er = __fsub_rn(ur, kr);
                            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
                            if ((double)__fmul_rn(e2, inv_s2) > 7.8) continue;
                        } else {
                            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                            if ((double)__fmul_rn(e2, inv_s2) > 5.99) continue;
                        }
                    }
                    const int d = grid_hamming(qa, qb, kdesc[2 * (size_t)i], kdesc[2 * (size_t)i + 1]);
                    if (d < bestDist) { bestDist = d; bestIdx = i; }
                }
            }
        } while (0);
        const bool ok = bestIdx >= 0 && bestDist <= th_low;
        P.best_idx[pi] = ok ? bestIdx : -1;
        P.best_dist[pi] = bestDist;
        mine += ok ? 1 : 0;
    }
    if (mine) atomicAdd(&s_found, mine);
    __syncthreads();
    if (tid == 0) *P.nfound = s_found;
}

void orbx_launch_fuse_search(const OrbxFuseDev* d_jobs, int njobs, int max_n, const OrbxFuseCam& cam, cudaStream_t st)
{
    if (njobs <= 0) return;
    const size_t smem = grid_smem_bytes((size_t)std::max(max_n, 1)) + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(fuse_search_kernel, mark[0], smem);
    fuse_search_kernel<<<njobs, 512, smem, st>>>(d_jobs, cam, 50 /* TH_LOW */);
}

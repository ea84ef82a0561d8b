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
    if (P.result_out) {                                                          // the atomics above are this CTA's own: visible after the barrier
        for (int i = tid; i < P.n; i += blockDim.x) P.result_out[i] = P.match[i];
        if (tid == 0) P.result_out[P.n] = s_success;
    }
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

// ------------------------------------------------------------------------------------------------------ isInFrustum
// Frame::isInFrustum (Frame.cc:315-378), the per-map-point prologue of Tracking::SearchLocalPoints: projection with
// mRcw / mtcw, image-bounds, distance-invariance and viewing-angle tests, MapPoint::PredictScale; writes the five mTrack*
// fields the matcher above reads. Thread per map point, arithmetic as in fuse_search_kernel below.
__global__ void __launch_bounds__(256) in_frustum_kernel(OrbxFrustumArgs A, OrbxFuseCam cam)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.npts) return;
    bool ok = false;
    do {
        const float X = A.pt_xyz[3 * i], Y = A.pt_xyz[3 * i + 1], Z = A.pt_xyz[3 * i + 2];
        float c3[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            float s = __fmul_rn(A.Tcw[3 * r], X);
            s = __fadd_rn(s, __fmul_rn(A.Tcw[3 * r + 1], Y));
            s = __fadd_rn(s, __fmul_rn(A.Tcw[3 * r + 2], Z));
            c3[r] = __fadd_rn(s, A.Tcw[9 + r]);
        }
        if (c3[2] < 0.0f) break;
        const float invz = __fdiv_rn(1.0f, c3[2]);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(cam.fx, c3[0]), invz), cam.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(cam.fy, c3[1]), invz), cam.cy);
        if (u < cam.minX || u > cam.maxX) break;
        if (v < cam.minY || v > cam.maxY) break;
        const float po0 = __fsub_rn(X, A.Ow[0]), po1 = __fsub_rn(Y, A.Ow[1]), po2 = __fsub_rn(Z, A.Ow[2]);
        double s2 = __dmul_rn((double)po0, (double)po0);
        s2 = __dadd_rn(s2, __dmul_rn((double)po1, (double)po1));
        s2 = __dadd_rn(s2, __dmul_rn((double)po2, (double)po2));
        const float dist = __double2float_rn(__dsqrt_rn(s2));
        if (dist < A.pt_dist[3 * i] || dist > A.pt_dist[3 * i + 1]) break;
        const float* nrm = A.pt_normal + 3 * i;
        double dot = __dmul_rn((double)po0, (double)nrm[0]);
        dot = __dadd_rn(dot, __dmul_rn((double)po1, (double)nrm[1]));
        dot = __dadd_rn(dot, __dmul_rn((double)po2, (double)nrm[2]));
        const float viewCos = __double2float_rn(__ddiv_rn(dot, (double)dist));
        if (viewCos < A.view_cos_limit) break;
        if (!(u == u) || !(v == v) || !(viewCos == viewCos)) break;          // NaN: out of the function's domain (a point at the camera centre)
        const float ratio = __fdiv_rn(A.pt_dist[3 * i + 2], dist);
        int level = 0;
        for (int n = 0; n < cam.nlevels - 1; n++) level += (ratio >= cam.level_ratio[n]) ? 1 : 0;
        OrbxTrackQueryDev q;
        q.x = u; q.y = v; q.xr = __fsub_rn(u, __fmul_rn(cam.bf, invz)); q.view_cos = viewCos; q.level = level;
        A.q[i] = q;
        ok = true;
    } while (0);
    A.in_view[i] = ok ? 1 : 0;
}

void orbx_launch_in_frustum(const OrbxFrustumArgs& a, const OrbxFuseCam& cam, cudaStream_t st)
{
    if (a.npts <= 0) return;
    in_frustum_kernel<<<(a.npts + 255) / 256, 256, 0, st>>>(a, cam);
}

// ------------------------------------------------------------------------------------------------------------- Fuse
// One CTA per (keyframe, map-point list). Thread per map point: projection with OpenCV's small-matrix gemm order
// (`Rcw*p3Dw + tcw`: f32 products summed left to right, the addend last), `1/z` as an f32 division, cv::norm and
// Mat::dot accumulated in f64 in element order, un-contracted f32 elsewhere (the pinned reference semantics,
// DESIGN.md §3). MapPoint::PredictScale's `ceil(logf(ratio) / mfLogScaleFactor)` is evaluated through a threshold table
// built on the host with the host's own logf (level_ratio[n] = smallest ratio whose predicted level exceeds n), so the
// level equals the host libm's bit for bit.
__global__ void __launch_bounds__(512) fuse_search_kernel(const OrbxFuseDev* __restrict__ jobs, OrbxFuseCam cam, int th_low, int th_high)
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
    // KeyFrame keeps the image bounds as `const int` (KeyFrame.h:236-239, initialised from the Frame's floats): its grid is
    // the Frame's (float bounds, above), but IsInImage and GetFeaturesInArea compare against the truncated values
    const float kminX = (float)(int)cam.minX, kmaxX = (float)(int)cam.maxX, kminY = (float)(int)cam.minY, kmaxY = (float)(int)cam.maxY;
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
            if (P.mode == 2) {                                                   // SearchBySim3: p3Dc2 = sR21*p3Dc1 + t21 (:1296-1297)
                float d3[3];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    float s = __fmul_rn(P.T2[3 * r], c3[0]);
                    s = __fadd_rn(s, __fmul_rn(P.T2[3 * r + 1], c3[1]));
                    s = __fadd_rn(s, __fmul_rn(P.T2[3 * r + 2], c3[2]));
                    d3[r] = __fadd_rn(s, P.T2[9 + r]);
                }
                c3[0] = d3[0]; c3[1] = d3[1]; c3[2] = d3[2];
            }
            if (c3[2] < 0.0f) break;
            // `1/z` (:954, f32 division) in Fuse(pKF, vpMapPoints); `1.0/z` (:1146, f64 division rounded to f32) in the Sim3 form
            const float invz = P.mode == 0 ? __fdiv_rn(1.0f, c3[2]) : __double2float_rn(__ddiv_rn(1.0, (double)c3[2]));
            const float x = __fmul_rn(c3[0], invz), y = __fmul_rn(c3[1], invz);
            const float u = __fadd_rn(__fmul_rn(cam.fx, x), cam.cx), v = __fadd_rn(__fmul_rn(cam.fy, y), cam.cy);
            if (!(u >= kminX && u < kmaxX && v >= kminY && v < kmaxY)) break;                  // KeyFrame::IsInImage
            const float ur = __fsub_rn(u, __fmul_rn(cam.bf, invz));
            // PO = p3Dw - Ow; SearchBySim3 measures the point in the target camera instead: cv::norm(p3Dc2) (:1318)
            const float po0 = P.mode == 2 ? c3[0] : __fsub_rn(X, P.Ow[0]), po1 = P.mode == 2 ? c3[1] : __fsub_rn(Y, P.Ow[1]),
                        po2 = P.mode == 2 ? c3[2] : __fsub_rn(Z, P.Ow[2]);
            double s2 = __dmul_rn((double)po0, (double)po0);
            s2 = __dadd_rn(s2, __dmul_rn((double)po1, (double)po1));
            s2 = __dadd_rn(s2, __dmul_rn((double)po2, (double)po2));
            const float dist3D = __double2float_rn(__dsqrt_rn(s2));
            if (dist3D < P.pt_dist[3 * pi] || dist3D > P.pt_dist[3 * pi + 1]) break;            // min / max distance invariance
            if (P.mode != 2) {                                                   // viewing angle (:977-980 / :1163-1166)
                const float* nrm = P.pt_normal + 3 * pi;
                double dot = __dmul_rn((double)po0, (double)nrm[0]);
                dot = __dadd_rn(dot, __dmul_rn((double)po1, (double)nrm[1]));
                dot = __dadd_rn(dot, __dmul_rn((double)po2, (double)nrm[2]));
                if (dot < __dmul_rn(0.5, (double)dist3D)) break;
            }
            const float ratio = __fdiv_rn(P.pt_dist[3 * pi + 2], dist3D);                         // mfMaxDistance / currentDist
            int level = 0;
            for (int n = 0; n < cam.nlevels - 1; n++) level += (ratio >= cam.level_ratio[n]) ? 1 : 0;
            const float radius = __fmul_rn(P.th, cam.scale_factors[level]);
            int cx0, cx1, cy0, cy1;
            if (!grid_window(u, v, radius, kminX, kminY, invW, invH, cx0, cx1, cy0, cy1)) break;
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
        const bool ok = bestIdx >= 0 && bestDist <= (P.mode == 2 ? th_high : th_low);
        P.best_idx[pi] = ok ? bestIdx : -1;
        if (P.best_dist) P.best_dist[pi] = bestDist;
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
    fuse_search_kernel<<<njobs, 512, smem, st>>>(d_jobs, cam, 50 /* TH_LOW */, 100 /* TH_HIGH */);
}

// ---------------------------------------------------------------------------------------- SearchBySim3: mutual check
// vnMatch1 / vnMatch2 come from two fuse_search_kernel jobs in mode 2; a pair survives when both directions agree
// (ORBmatcher.cc:1468-1484). One block per keyframe pair.
__global__ void __launch_bounds__(256) sim3_mutual_kernel(const int* __restrict__ match1, int n1, const int* __restrict__ match2, int n2,
                                                          int* __restrict__ match12, int* __restrict__ nfound)
{
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int mine = 0;
    for (int i1 = threadIdx.x; i1 < n1; i1 += blockDim.x) {
        const int idx2 = match1[i1];
        const bool ok = idx2 >= 0 && idx2 < n2 && match2[idx2] == i1;
        match12[i1] = ok ? idx2 : -1;
        mine += ok ? 1 : 0;
    }
    if (mine) atomicAdd(&s_n, mine);
    __syncthreads();
    if (threadIdx.x == 0) *nfound = s_n;
}

void orbx_launch_sim3_mutual(const int* d_match1, int n1, const int* d_match2, int n2, int* d_match12, int* d_nfound, cudaStream_t st)
{
    sim3_mutual_kernel<<<1, 256, 0, st>>>(d_match1, n1, d_match2, n2, d_match12, d_nfound);
}

// --------------------------------------------------------------- SearchByProjection(CurrentFrame, pKF, sAlreadyFound, ...)
//                                                                 SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)
// The relocalisation matcher (ORBmatcher.cc:1648-1795, mode 0) and the loop-closing matcher (:327-440, mode 1). Both
// project map points, gate them like Fuse, and give each the nearest FREE keypoint of its window: a keypoint that holds
// a map point — from before the call or assigned earlier in the loop — is skipped (:1714-1715 / :412-413), so every
// assignment blocks later queries. Same fixed-point iteration as above with taker[k] = lowest query index that took k.
__global__ void __launch_bounds__(512) seq_projection_kernel(const OrbxSeqProjDev* __restrict__ jobs, OrbxFuseCam cam)
{
    extern __shared__ __align__(16) unsigned char s_raw6[];
    const OrbxSeqProjDev P = jobs[blockIdx.x];
    GridSmem G;
    int* taker = reinterpret_cast<int*>(grid_carve(s_raw6, P.n, G));
    __shared__ int s_changed, s_hist[32], s_ind[3], s_success, s_removed, s_w[17];
    const int tid = threadIdx.x;
    const float invW = __fdiv_rn(64.0f, __fsub_rn(cam.maxX, cam.minX)), invH = __fdiv_rn(48.0f, __fsub_rn(cam.maxY, cam.minY));
    for (int i = tid; i < P.n; i += blockDim.x) { taker[i] = 0x7fffffff; P.match[i] = -1; }
    if (tid < 32) s_hist[tid] = 0;
    if (tid == 0) { s_success = 0; s_removed = 0; }
    grid_build(P.kps, P.n, cam.minX, cam.minY, invW, invH, G, s_w);
    // mode 1 searches a KeyFrame, whose bounds are the Frame's truncated to int (see fuse_search_kernel)
    const bool kf = P.mode == 1;
    const float wminX = kf ? (float)(int)cam.minX : cam.minX, wmaxX = kf ? (float)(int)cam.maxX : cam.maxX,
                wminY = kf ? (float)(int)cam.minY : cam.minY, wmaxY = kf ? (float)(int)cam.maxY : cam.maxY;
    // ---- projection and gates of every map point
    for (int pi = tid; pi < P.npts; pi += blockDim.x) {
        OrbxProjQuery q; q.r = -1.f; q.u = q.v = q.ur = 0.f; q.min_level = q.max_level = -1;
        do {
            if (!(P.pt_flags[pi] & 1)) break;                                    // NULL, isBad() or already found
            const float X = P.pt_xyz[3 * pi], Y = P.pt_xyz[3 * pi + 1], Z = P.pt_xyz[3 * pi + 2];
            float c3[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                float s = __fmul_rn(P.Tcw[3 * r], X);
                s = __fadd_rn(s, __fmul_rn(P.Tcw[3 * r + 1], Y));
                s = __fadd_rn(s, __fmul_rn(P.Tcw[3 * r + 2], Z));
                c3[r] = __fadd_rn(s, P.Tcw[9 + r]);
            }
            float u, v;
            if (P.mode == 0) {                                                   // :1673-1684
                const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)c3[2]));
                u = __fadd_rn(__fmul_rn(__fmul_rn(cam.fx, c3[0]), invzc), cam.cx);
                v = __fadd_rn(__fmul_rn(__fmul_rn(cam.fy, c3[1]), invzc), cam.cy);
                if (u < cam.minX || u > cam.maxX) break;
                if (v < cam.minY || v > cam.maxY) break;
                if (!(u == u) || !(v == v)) break;                               // NaN passes every `<` above; its window is empty anyway
            } else {                                                             // :352-366
                if (c3[2] < 0.0f) break;
                const float invz = __fdiv_rn(1.0f, c3[2]);
                const float x = __fmul_rn(c3[0], invz), y = __fmul_rn(c3[1], invz);
                u = __fadd_rn(__fmul_rn(cam.fx, x), cam.cx); v = __fadd_rn(__fmul_rn(cam.fy, y), cam.cy);
                if (!(u >= wminX && u < wmaxX && v >= wminY && v < wmaxY)) break;
            }
            const float po0 = __fsub_rn(X, P.Ow[0]), po1 = __fsub_rn(Y, P.Ow[1]), po2 = __fsub_rn(Z, P.Ow[2]);
            double s2 = __dmul_rn((double)po0, (double)po0);
            s2 = __dadd_rn(s2, __dmul_rn((double)po1, (double)po1));
            s2 = __dadd_rn(s2, __dmul_rn((double)po2, (double)po2));
            const float dist3D = __double2float_rn(__dsqrt_rn(s2));
            if (dist3D < P.pt_dist[3 * pi] || dist3D > P.pt_dist[3 * pi + 1]) break;
            if (P.mode == 1) {
                const float* nrm = P.pt_normal + 3 * pi;
                double dot = __dmul_rn((double)po0, (double)nrm[0]);
                dot = __dadd_rn(dot, __dmul_rn((double)po1, (double)nrm[1]));
                dot = __dadd_rn(dot, __dmul_rn((double)po2, (double)nrm[2]));
                if (dot < __dmul_rn(0.5, (double)dist3D)) break;
            }
            const float ratio = __fdiv_rn(P.pt_dist[3 * pi + 2], dist3D);
            int level = 0;
            for (int n = 0; n < cam.nlevels - 1; n++) level += (ratio >= cam.level_ratio[n]) ? 1 : 0;
            q.u = u; q.v = v; q.r = __fmul_rn(P.th, cam.scale_factors[level]);
            q.min_level = level - 1; q.max_level = P.mode == 0 ? level + 1 : level;
        } while (0);
        P.query[pi] = q;
        P.assign[pi] = -1;
    }
    __syncthreads();
    const uint4* kdesc = reinterpret_cast<const uint4*>(P.desc);
    const uint4* pdesc = reinterpret_cast<const uint4*>(P.pt_desc);
    for (int round = 0; round <= P.npts; round++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int qi = tid; qi < P.npts; qi += blockDim.x) {
            const OrbxProjQuery q = P.query[qi];
            if (q.r < 0.f) continue;
            int cx0, cx1, cy0, cy1;
            int bestDist = 256, bestIdx = -1;
            if (grid_window(q.u, q.v, q.r, wminX, wminY, invW, invH, cx0, cx1, cy0, cy1)) {
                const uint4 qa = pdesc[2 * (size_t)qi], qb = pdesc[2 * (size_t)qi + 1];
                for (int ix = cx0; ix <= cx1; ix++) {
                    const int j0 = G.cstart[ix * GRID_ROWS + cy0], j1 = G.cstart[ix * GRID_ROWS + cy1 + 1];
                    for (int j = j0; j < j1; j++) {
                        const int i = G.order[j];
                        const GridKp k = G.kp[i];
                        if (k.octave < q.min_level || k.octave > q.max_level) continue;
                        if (!(fabsf(__fsub_rn(k.x, q.u)) < q.r && fabsf(__fsub_rn(k.y, q.v)) < q.r)) continue;
                        if ((P.occupied && P.occupied[i]) || taker[i] < qi) continue;
                        const int d = grid_hamming(qa, qb, kdesc[2 * (size_t)i], kdesc[2 * (size_t)i + 1]);
                        if (d < bestDist) { bestDist = d; bestIdx = i; }
                    }
                }
            }
            const int a = bestDist <= P.th_dist ? bestIdx : -1;
            if (a != P.assign[qi]) { P.assign[qi] = a; s_changed = 1; }
        }
        __syncthreads();
        if (!s_changed) break;
        for (int i = tid; i < P.n; i += blockDim.x) taker[i] = 0x7fffffff;
        __syncthreads();
        for (int qi = tid; qi < P.npts; qi += blockDim.x) {
            const int a = P.assign[qi];
            if (a >= 0) atomicMin(&taker[a], qi);
        }
        __syncthreads();
    }
    const bool ori = P.mode == 0 && P.check_orientation;
    for (int qi = tid; qi < P.npts; qi += blockDim.x) {
        const int a = P.assign[qi];
        if (a < 0) continue;
        P.match[a] = qi;                                                         // unique: every assignment blocks the keypoint
        atomicAdd(&s_success, 1);
        if (ori) {
            float rot = __fsub_rn(P.pt_angle[qi], P.kps[a].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
            if (bin == 30) bin = 0;
            P.query[qi].min_level = bin;
            atomicAdd(&s_hist[bin], 1);
        }
    }
    __syncthreads();
    if (ori) {
        if (tid == 0) {
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
        for (int qi = tid; qi < P.npts; qi += blockDim.x) {
            const int a = P.assign[qi];
            if (a < 0) continue;
            const int bin = P.query[qi].min_level;
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { P.match[a] = -1; atomicAdd(&s_removed, 1); }
        }
        __syncthreads();
    }
    if (tid == 0) *P.nmatches = s_success - s_removed;
}

void orbx_launch_seq_projection(const OrbxSeqProjDev* d_jobs, int njobs, int max_n, const OrbxFuseCam& cam, cudaStream_t st)
{
    if (njobs <= 0) return;
    const size_t n1 = (size_t)std::max(max_n, 1);
    const size_t smem = grid_smem_bytes(n1) + n1 * sizeof(int) + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(seq_projection_kernel, mark[0], smem);
    seq_projection_kernel<<<njobs, 512, smem, st>>>(d_jobs, cam);
}

// ------------------------------------------------------------------------------------------ SearchForInitialization
// ORBmatcher::SearchForInitialization (ORBmatcher.cc:442-587), the matcher of the monocular initialiser: level-0 keypoints
// of F1 against the level-0 keypoints of F2 inside a square window around vbPrevMatched[i1]. The loop is sequential
// through vMatchedDistance / vnMatches21 (a later F1 keypoint may take an F2 keypoint away from an earlier one if it is
// strictly closer). All warps first collect, per F1 keypoint, the four nearest window candidates in the reference's order
// (key = dist << 16 | candidate ordinal reproduces `dist < bestDist`: first minimum, multiset second best); one warp then
// walks the F1 keypoints in order and decides from those four, rescanning the window only when fewer than two of them are
// still open and more candidates exist.
__global__ void __launch_bounds__(512) init_match_kernel(const OrbxInitPairDev* __restrict__ pairs, float minX, float maxX,
                                                         float minY, float maxY, float nnratio, int check_orientation, int th_low)
{
    extern __shared__ __align__(16) unsigned char s_raw7[];
    const OrbxInitPairDev P = pairs[blockIdx.x];
    GridSmem G;
    int* mdist = reinterpret_cast<int*>(grid_carve(s_raw7, P.n2, G));       // vMatchedDistance [n2]
    int* m21 = mdist + P.n2;                                                  // vnMatches21 [n2]
    __shared__ int s_w[17], s_hist[32], s_ind[3], s_nm;
    const int tid = threadIdx.x, lane = tid & 31;
    const float invW = __fdiv_rn(64.0f, __fsub_rn(maxX, minX)), invH = __fdiv_rn(48.0f, __fsub_rn(maxY, minY));
    for (int i = tid; i < P.n2; i += blockDim.x) { mdist[i] = 0x7fffffff; m21[i] = -1; }
    for (int i = tid; i < P.n1; i += blockDim.x) { P.match12[i] = -1; P.bin_of[i] = -1; }
    if (tid < 32) s_hist[tid] = 0;
    if (tid == 0) s_nm = 0;
    grid_build(P.kps2, P.n2, minX, minY, invW, invH, G, s_w);
    const uint4* d1 = reinterpret_cast<const uint4*>(P.desc1);
    const uint4* d2 = reinterpret_cast<const uint4*>(P.desc2);
    const float r = (float)P.window;
    // ---- phase 1, all warps, one F1 keypoint per warp at a time: the FOUR nearest window candidates in the reference's
    //      order (key = dist << 16 | candidate ordinal) without the vMatchedDistance filter, and the candidate count.
    //      The sequential phase below decides from these four whenever they suffice, which is exact: any candidate
    //      beyond them has a larger key than all four.
    const int wid = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i1 = wid; i1 < P.n1; i1 += nwarps) {
        unsigned k0 = 0xffffffffu, k1 = 0xffffffffu, k2 = 0xffffffffu, k3 = 0xffffffffu;
        int x0 = -1, x1 = -1, x2 = -1, x3 = -1, ncand = 0;
        const OrbxKp28 kp1 = P.kps1[i1];
        int cx0, cx1, cy0, cy1;
        const float x = P.prev[2 * i1], y = P.prev[2 * i1 + 1];
        if (kp1.octave <= 0 && grid_window(x, y, r, minX, minY, invW, invH, cx0, cx1, cy0, cy1)) {
            const uint4 qa = d1[2 * (size_t)i1], qb = d1[2 * (size_t)i1 + 1];
            int base = 0;
            for (int ix = cx0; ix <= cx1; ix++) {
                const int j0 = G.cstart[ix * GRID_ROWS + cy0], j1 = G.cstart[ix * GRID_ROWS + cy1 + 1];
                for (int j = j0 + lane; j < j1; j += 32) {
                    const int i2 = G.order[j];
                    const GridKp k = G.kp[i2];
                    if (k.octave > 0) continue;                                  // GetFeaturesInArea(.., level1, level1) with level1 == 0
                    if (!(fabsf(__fsub_rn(k.x, x)) < r && fabsf(__fsub_rn(k.y, y)) < r)) continue;
                    const unsigned key = ((unsigned)grid_hamming(qa, qb, d2[2 * (size_t)i2], d2[2 * (size_t)i2 + 1]) << 16) | (unsigned)(base + j - j0);
                    ncand++;
                    if (key < k3) {                                              // insertion into the lane's sorted four
                        k3 = key; x3 = i2;
                        if (k3 < k2) { const unsigned t = k2; k2 = k3; k3 = t; const int u = x2; x2 = x3; x3 = u; }
                        if (k2 < k1) { const unsigned t = k1; k1 = k2; k2 = t; const int u = x1; x1 = x2; x2 = u; }
                        if (k1 < k0) { const unsigned t = k0; k0 = k1; k1 = t; const int u = x0; x0 = x1; x1 = u; }
                    }
                }
                base += j1 - j0;
            }
        }
        ncand = __reduce_add_sync(0xffffffffu, ncand);
        unsigned outk = 0xffffffffu; int outx = -1;                              // lane e (< 4) ends up with the e-th smallest
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const unsigned m = __reduce_min_sync(0xffffffffu, k0);
            const unsigned own = __ballot_sync(0xffffffffu, k0 == m && m != 0xffffffffu);
            const int src = own ? __ffs(own) - 1 : 0;
            const int xi = __shfl_sync(0xffffffffu, x0, src);
            if (lane == e) { outk = m; outx = m != 0xffffffffu ? xi : -1; }
            if (own && lane == src) { k0 = k1; x0 = x1; k1 = k2; x1 = x2; k2 = k3; x2 = x3; k3 = 0xffffffffu; }
        }
        if (lane < 4) { P.top_key[4 * (size_t)i1 + lane] = outk; P.top_idx[4 * (size_t)i1 + lane] = outx; }
        if (lane == 0) P.ncand[i1] = ncand;
    }
    __syncthreads();
    // ---- phase 2, one warp: the reference loop in order (:463-546)
    if (tid < 32) {
        int nmatches = 0;
        for (int i1 = 0; i1 < P.n1; i1++) {
            const int ncand = P.ncand[i1];
            if (ncand == 0) continue;                                            // level > 0, window off the grid, or no candidate
            const unsigned key = lane < 4 ? P.top_key[4 * (size_t)i1 + lane] : 0xffffffffu;
            const int idx = lane < 4 ? P.top_idx[4 * (size_t)i1 + lane] : -1;
            const bool valid = key != 0xffffffffu;
            const bool open = valid && !(mdist[idx] <= (int)(key >> 16));        // :494-495
            const unsigned U = __ballot_sync(0xffffffffu, open);
            unsigned m1, m2; int bestIdx2;
            if (__popc(U) >= 2 || ncand <= 4) {
                if (!U) continue;
                const int l1 = __ffs(U) - 1, l2 = __ffs(U & (U - 1)) - 1;
                m1 = __shfl_sync(0xffffffffu, key, l1); bestIdx2 = __shfl_sync(0xffffffffu, idx, l1);
                m2 = l2 >= 0 ? __shfl_sync(0xffffffffu, key, l2) : 0xffffffffu;
            } else {
                // fewer than two of the four are still open and there are more candidates: rescan the window with the filter
                const float x = P.prev[2 * i1], y = P.prev[2 * i1 + 1];
                int cx0, cx1, cy0, cy1;
                grid_window(x, y, r, minX, minY, invW, invH, cx0, cx1, cy0, cy1);
                const uint4 qa = d1[2 * (size_t)i1], qb = d1[2 * (size_t)i1 + 1];
                unsigned b1 = 0xffffffffu, b2 = 0xffffffffu;
                int base = 0;
                for (int ix = cx0; ix <= cx1; ix++) {
                    const int j0 = G.cstart[ix * GRID_ROWS + cy0], j1 = G.cstart[ix * GRID_ROWS + cy1 + 1];
                    for (int j = j0 + lane; j < j1; j += 32) {
                        const int i2 = G.order[j];
                        const GridKp k = G.kp[i2];
                        if (k.octave > 0) continue;
                        if (!(fabsf(__fsub_rn(k.x, x)) < r && fabsf(__fsub_rn(k.y, y)) < r)) continue;
                        const int dist = grid_hamming(qa, qb, d2[2 * (size_t)i2], d2[2 * (size_t)i2 + 1]);
                        if (mdist[i2] <= dist) continue;
                        const unsigned kk = ((unsigned)dist << 16) | (unsigned)(base + j - j0);
                        const unsigned t = max(kk, b1); b1 = min(b1, kk); b2 = min(b2, t);
                    }
                    base += j1 - j0;
                }
                m1 = __reduce_min_sync(0xffffffffu, b1);
                const unsigned c2 = b1 == m1 ? b2 : b1;
                m2 = __reduce_min_sync(0xffffffffu, c2);
                if (m1 == 0xffffffffu) continue;
                int ord = (int)(m1 & 0xffffu);                                   // the ordinal back to the keypoint index
                bestIdx2 = -1;
                for (int ix = cx0; ix <= cx1; ix++) {
                    const int j0 = G.cstart[ix * GRID_ROWS + cy0], j1 = G.cstart[ix * GRID_ROWS + cy1 + 1];
                    if (ord < j1 - j0) { bestIdx2 = G.order[j0 + ord]; break; }
                    ord -= j1 - j0;
                }
            }
            const int bestDist = (int)(m1 >> 16);
            const float bestDist2 = m2 == 0xffffffffu ? 2147483648.0f /* (float)INT_MAX */ : (float)(int)(m2 >> 16);
            if (bestDist <= th_low && (float)bestDist < __fmul_rn(bestDist2, nnratio)) {
                if (lane == 0) {
                    if (m21[bestIdx2] >= 0) { P.match12[m21[bestIdx2]] = -1; nmatches--; }
                    P.match12[i1] = bestIdx2; m21[bestIdx2] = i1; mdist[bestIdx2] = bestDist;
                    nmatches++;
                    if (check_orientation) {
                        float rot = __fsub_rn(P.kps1[i1].angle, P.kps2[bestIdx2].angle);
                        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                        int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                        if (bin == 30) bin = 0;
                        P.bin_of[i1] = bin; s_hist[bin]++;
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) s_nm = nmatches;
    }
    __syncthreads();
    if (check_orientation) {
        if (tid == 0) {
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
        int removed = 0;
        for (int i1 = tid; i1 < P.n1; i1 += blockDim.x) {
            const int bin = P.bin_of[i1];
            if (bin >= 0 && bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2] && P.match12[i1] >= 0) { P.match12[i1] = -1; removed++; }
        }
        if (removed) atomicSub(&s_nm, removed);
        __syncthreads();
    }
    // vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt (:582-584)
    for (int i1 = tid; i1 < P.n1; i1 += blockDim.x) {
        const int m = P.match12[i1];
        float px = P.prev[2 * i1], py = P.prev[2 * i1 + 1];
        if (m >= 0) { px = G.kp[m].x; py = G.kp[m].y; }
        P.prev_out[2 * i1] = px; P.prev_out[2 * i1 + 1] = py;
    }
    if (tid == 0) *P.nmatches = s_nm;
}

void orbx_launch_init_match(const OrbxInitPairDev* d_pairs, int npairs, int max_n2, const float* bounds4, float nnratio,
                            int check_orientation, cudaStream_t st)
{
    if (npairs <= 0) return;
    const size_t n1 = (size_t)std::max(max_n2, 1);
    const size_t smem = grid_smem_bytes(n1) + 2 * n1 * sizeof(int) + 16;
    static OrbxSmemMark mark[1] = {};
    orbx_need_smem(init_match_kernel, mark[0], smem);
    init_match_kernel<<<npairs, 512, smem, st>>>(d_pairs, bounds4[0], bounds4[1], bounds4[2], bounds4[3], nnratio, check_orientation,
                                                 50 /* TH_LOW */);
}

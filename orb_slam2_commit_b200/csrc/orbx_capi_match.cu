// orbx_capi_match.cu — C ABI of the map-point matchers of orbx_match.cu (include/orbx.h): SearchByProjection(F, vpMapPoints),
// the search half of Fuse, the relocalisation / loop-closing SearchByProjection, SearchBySim3, SearchForInitialization and
// Frame::isInFrustum. Host forms pack their arrays into one stream-ordered allocation and call the device forms.
#include "orbx_capi_common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-142) — orbx_match.cu
extern "C" int orbx_search_local_points_device(const OrbxLocalPointsFrame* frames, int nframes, const float* bounds4,
                                               const float* scale_factors, int nlevels, float th, float nnratio, int device,
                                               void* cuda_stream)
{
    if (nframes <= 0) return ORBX_OK;
    if (!frames || !bounds4 || !scale_factors || nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    size_t tot_q = 0; int max_n = 0;
    for (int p = 0; p < nframes; p++) {
        const OrbxLocalPointsFrame& f = frames[p];
        if (f.n < 0 || f.nq < 0 || !f.match || !f.nmatches) return fail(ORBX_ERR_INVALID, "bad frame");
        if (f.n > 10000) return fail(ORBX_ERR_UNSUPPORTED, "more than 10000 keypoints in the frame");
        if ((f.n > 0 && (!f.keypoints || !f.descriptors)) || (f.nq > 0 && (!f.queries || !f.query_descriptors || !f.query_flags)))
            return fail(ORBX_ERR_INVALID, "NULL array in frame");
        tot_q += (size_t)f.nq; max_n = std::max(max_n, f.n);
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_f = al((size_t)nframes * sizeof(OrbxLocalFrameDev)), b_a = al(std::max<size_t>(tot_q, 1) * 4), b_sf = al((size_t)nlevels * 4);
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, b_f + b_a + b_sf, st));
    int* d_a = (int*)(pool + b_f);
    float* d_sf = (float*)(pool + b_f + b_a);
    std::vector<OrbxLocalFrameDev> hf(nframes);
    size_t off = 0;
    for (int p = 0; p < nframes; p++) {
        const OrbxLocalPointsFrame& f = frames[p];
        OrbxLocalFrameDev& d = hf[p];
        d.kps = (const OrbxKp28*)f.keypoints; d.desc = f.descriptors; d.u_right = f.u_right; d.occupied = f.occupied; d.n = f.n;
        d.q = (const OrbxTrackQueryDev*)f.queries; d.qdesc = f.query_descriptors; d.qflags = f.query_flags; d.nq = f.nq;
        d.match = f.match; d.nmatches = f.nmatches; d.assign = d_a + off; off += (size_t)f.nq; d.result_out = nullptr;
    }
    cudaError_t e;
    do {
        if ((e = cudaMemcpyAsync(pool, hf.data(), (size_t)nframes * sizeof(OrbxLocalFrameDev), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(d_sf, scale_factors, (size_t)nlevels * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        orbx_launch_local_points((const OrbxLocalFrameDev*)pool, nframes, max_n, bounds4, d_sf, nlevels, th, nnratio, st);
        e = cudaGetLastError();
    } while (0);
    cudaFreeAsync(pool, st);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}


extern "C" int orbx_search_local_points(const OrbxLocalPointsFrame* frame, const float* bounds4, const float* scale_factors,
                                        int nlevels, float th, float nnratio, int device)
{
    if (!frame || !frame->match || !frame->nmatches) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    const int n = frame->n, nq = frame->nq;
    if (n < 0 || nq < 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    orbx_keep_mempool(device);
    HostPack P;
    const size_t i_kp = P.add((size_t)n * 28), i_d = P.add((size_t)n * 32), i_ur = P.add((size_t)n * 4), i_oc = P.add(n),
                 i_q = P.add((size_t)nq * sizeof(OrbxTrackQuery)), i_qd = P.add((size_t)nq * 32), i_qf = P.add(nq),
                 i_m = P.add((size_t)n * 4), i_nm = P.add(4);
    CK(cudaMallocAsync(&P.pool, P.tot, 0));
    OrbxLocalPointsFrame d = *frame;
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        auto up = [&](size_t i, const void* src, size_t bytes) { return (!src || !bytes) ? cudaSuccess : cudaMemcpyAsync(P.at(i), src, bytes, cudaMemcpyHostToDevice, 0); };
        if ((e = up(i_kp, frame->keypoints, (size_t)n * 28)) != cudaSuccess) break;
        if ((e = up(i_d, frame->descriptors, (size_t)n * 32)) != cudaSuccess) break;
        if ((e = up(i_ur, frame->u_right, (size_t)n * 4)) != cudaSuccess) break;
        if ((e = up(i_oc, frame->occupied, (size_t)n)) != cudaSuccess) break;
        if ((e = up(i_q, frame->queries, (size_t)nq * sizeof(OrbxTrackQuery))) != cudaSuccess) break;
        if ((e = up(i_qd, frame->query_descriptors, (size_t)nq * 32)) != cudaSuccess) break;
        if ((e = up(i_qf, frame->query_flags, (size_t)nq)) != cudaSuccess) break;
        d.keypoints = (const OrbxKeyPoint*)P.at(i_kp); d.descriptors = P.at(i_d);
        d.u_right = frame->u_right ? (const float*)P.at(i_ur) : nullptr; d.occupied = frame->occupied ? P.at(i_oc) : nullptr;
        d.queries = (const OrbxTrackQuery*)P.at(i_q); d.query_descriptors = P.at(i_qd); d.query_flags = P.at(i_qf);
        d.match = (int32_t*)P.at(i_m); d.nmatches = (int32_t*)P.at(i_nm);
        rc = orbx_search_local_points_device(&d, 1, bounds4, scale_factors, nlevels, th, nnratio, device, nullptr);
        if (rc != ORBX_OK) break;
        if (n > 0 && (e = cudaMemcpy(frame->match, P.at(i_m), (size_t)n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(frame->nmatches, P.at(i_nm), 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(P.pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// ORBmatcher::Fuse, search half (ORBmatcher.cc:918-1092, 1094-1236) — orbx_match.cu
// MapPoint::PredictScale (MapPoint.cc:407-422) is `ceil(logf(ratio) / mfLogScaleFactor)` clamped to [0, nlevels-1]. The
// device takes it from a threshold table instead of its own logf: level_ratio[n] = the smallest positive float whose
// predicted level exceeds n, found by bisection over the float bit patterns with the HOST's logf (monotonic), so the
// device level is the host libm's level for every ratio.
static void predict_scale_thresholds(float log_scale_factor, int nlevels, float* level_ratio)
{
    for (int n = 0; n < ORBX_MAX_LEVELS; n++) level_ratio[n] = INFINITY;
    auto exceeds = [&](uint32_t bits, int n) { float r; memcpy(&r, &bits, 4); return logf(r) / log_scale_factor > (float)n; };
    for (int n = 0; n + 1 < nlevels; n++) {
        uint32_t lo = 1u, hi = 0x7f7fffffu;                   // smallest denormal .. FLT_MAX
        if (!exceeds(hi, n)) continue;
        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (exceeds(mid, n)) hi = mid; else lo = mid + 1; }
        memcpy(&level_ratio[n], &lo, 4);
    }
}

extern "C" int orbx_fuse_search_device(const OrbxFuseJob* jobs, int njobs, const float* camera9, const float* scale_factors,
                                       const float* inv_level_sigma2, int nlevels, float log_scale_factor, int device,
                                       void* cuda_stream)
{
    if (njobs <= 0) return ORBX_OK;
    if (!jobs || !camera9 || !scale_factors || !inv_level_sigma2 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !(log_scale_factor > 0.f))
        return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int max_n = 0;
    std::vector<OrbxFuseDev> hj(njobs);
    for (int p = 0; p < njobs; p++) {
        const OrbxFuseJob& j = jobs[p];
        if (j.n < 0 || j.npts < 0 || !j.nfused || (j.npts > 0 && (!j.best_idx || !j.best_dist)) || j.mode < 0 || j.mode > 1)
            return fail(ORBX_ERR_INVALID, "bad job");
        if (j.n > 10000) return fail(ORBX_ERR_UNSUPPORTED, "more than 10000 keypoints in the keyframe");
        if ((j.n > 0 && (!j.keypoints || !j.descriptors)) ||
            (j.npts > 0 && (!j.pt_xyz || !j.pt_normal || !j.pt_dist || !j.pt_descriptors || !j.pt_flags))) return fail(ORBX_ERR_INVALID, "NULL array in job");
        OrbxFuseDev& d = hj[p];
        d.kps = (const OrbxKp28*)j.keypoints; d.desc = j.descriptors; d.u_right = j.u_right; d.n = j.n;
        memcpy(d.Tcw, j.Tcw, sizeof d.Tcw); memcpy(d.Ow, j.Ow, sizeof d.Ow); d.th = j.th; d.mode = j.mode;
        d.pt_xyz = j.pt_xyz; d.pt_normal = j.pt_normal; d.pt_dist = j.pt_dist; d.pt_desc = j.pt_descriptors; d.pt_flags = j.pt_flags;
        d.npts = j.npts; d.best_idx = j.best_idx; d.best_dist = j.best_dist; d.nfound = j.nfused;
        max_n = std::max(max_n, j.n);
    }
    OrbxFuseCam cam = {};
    cam.fx = camera9[0]; cam.fy = camera9[1]; cam.cx = camera9[2]; cam.cy = camera9[3]; cam.bf = camera9[4];
    cam.minX = camera9[5]; cam.maxX = camera9[6]; cam.minY = camera9[7]; cam.maxY = camera9[8]; cam.nlevels = nlevels;
    for (int l = 0; l < nlevels; l++) { cam.scale_factors[l] = scale_factors[l]; cam.inv_level_sigma2[l] = inv_level_sigma2[l]; }
    predict_scale_thresholds(log_scale_factor, nlevels, cam.level_ratio);
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, (size_t)njobs * sizeof(OrbxFuseDev), st));
    cudaError_t e;
    do {
        if ((e = cudaMemcpyAsync(pool, hj.data(), (size_t)njobs * sizeof(OrbxFuseDev), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        orbx_launch_fuse_search((const OrbxFuseDev*)pool, njobs, max_n, cam, st);
        e = cudaGetLastError();
    } while (0);
    cudaFreeAsync(pool, st);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_fuse_search(const OrbxFuseJob* job, const float* camera9, const float* scale_factors,
                                const float* inv_level_sigma2, int nlevels, float log_scale_factor, int device)
{
    if (!job || !job->nfused) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    const int n = job->n, np = job->npts;
    if (n < 0 || np < 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    orbx_keep_mempool(device);
    HostPack P;
    const size_t i_kp = P.add((size_t)n * 28), i_d = P.add((size_t)n * 32), i_ur = P.add((size_t)n * 4),
                 i_x = P.add((size_t)np * 12), i_nr = P.add((size_t)np * 12), i_ds = P.add((size_t)np * 12), i_pd = P.add((size_t)np * 32),
                 i_pf = P.add(np), i_bi = P.add((size_t)np * 4), i_bd = P.add((size_t)np * 4), i_nf = P.add(4);
    CK(cudaMallocAsync(&P.pool, P.tot, 0));
    OrbxFuseJob d = *job;
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        auto up = [&](size_t i, const void* src, size_t bytes) { return (!src || !bytes) ? cudaSuccess : cudaMemcpyAsync(P.at(i), src, bytes, cudaMemcpyHostToDevice, 0); };
        if ((e = up(i_kp, job->keypoints, (size_t)n * 28)) != cudaSuccess) break;
        if ((e = up(i_d, job->descriptors, (size_t)n * 32)) != cudaSuccess) break;
        if ((e = up(i_ur, job->u_right, (size_t)n * 4)) != cudaSuccess) break;
        if ((e = up(i_x, job->pt_xyz, (size_t)np * 12)) != cudaSuccess) break;
        if ((e = up(i_nr, job->pt_normal, (size_t)np * 12)) != cudaSuccess) break;
        if ((e = up(i_ds, job->pt_dist, (size_t)np * 12)) != cudaSuccess) break;
        if ((e = up(i_pd, job->pt_descriptors, (size_t)np * 32)) != cudaSuccess) break;
        if ((e = up(i_pf, job->pt_flags, (size_t)np)) != cudaSuccess) break;
        d.keypoints = (const OrbxKeyPoint*)P.at(i_kp); d.descriptors = P.at(i_d); d.u_right = job->u_right ? (const float*)P.at(i_ur) : nullptr;
        d.pt_xyz = (const float*)P.at(i_x); d.pt_normal = (const float*)P.at(i_nr); d.pt_dist = (const float*)P.at(i_ds);
        d.pt_descriptors = P.at(i_pd); d.pt_flags = P.at(i_pf);
        d.best_idx = (int32_t*)P.at(i_bi); d.best_dist = (int32_t*)P.at(i_bd); d.nfused = (int32_t*)P.at(i_nf);
        rc = orbx_fuse_search_device(&d, 1, camera9, scale_factors, inv_level_sigma2, nlevels, log_scale_factor, device, nullptr);
        if (rc != ORBX_OK) break;
        if (np > 0 && (e = cudaMemcpy(job->best_idx, P.at(i_bi), (size_t)np * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (np > 0 && (e = cudaMemcpy(job->best_dist, P.at(i_bd), (size_t)np * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(job->nfused, P.at(i_nf), 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(P.pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, ..) / SearchByProjection(pKF, Scw, ..), SearchBySim3,
// SearchForInitialization — orbx_match.cu
static int make_fuse_cam(const float* camera9, const float* scale_factors, const float* inv_level_sigma2, int nlevels,
                         float log_scale_factor, OrbxFuseCam& cam)
{
    if (!camera9 || !scale_factors || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !(log_scale_factor > 0.f)) return fail(ORBX_ERR_INVALID, "bad camera / level tables");
    memset(&cam, 0, sizeof cam);
    cam.fx = camera9[0]; cam.fy = camera9[1]; cam.cx = camera9[2]; cam.cy = camera9[3]; cam.bf = camera9[4];
    cam.minX = camera9[5]; cam.maxX = camera9[6]; cam.minY = camera9[7]; cam.maxY = camera9[8]; cam.nlevels = nlevels;
    for (int l = 0; l < nlevels; l++) { cam.scale_factors[l] = scale_factors[l]; cam.inv_level_sigma2[l] = inv_level_sigma2 ? inv_level_sigma2[l] : 0.f; }
    predict_scale_thresholds(log_scale_factor, nlevels, cam.level_ratio);
    return ORBX_OK;
}

extern "C" int orbx_search_by_projection_kf_device(const OrbxProjectionJob* jobs, int njobs, const float* camera9,
                                                   const float* scale_factors, int nlevels, float log_scale_factor,
                                                   int check_orientation, int device, void* cuda_stream)
{
    if (njobs <= 0) return ORBX_OK;
    if (!jobs) return fail(ORBX_ERR_INVALID, "bad argument");
    OrbxFuseCam cam;
    int rc = make_fuse_cam(camera9, scale_factors, nullptr, nlevels, log_scale_factor, cam);
    if (rc != ORBX_OK) return rc;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    size_t tot = 0; int max_n = 0;
    for (int p = 0; p < njobs; p++) {
        const OrbxProjectionJob& j = jobs[p];
        if (j.n < 0 || j.npts < 0 || !j.match || !j.nmatches || j.mode < 0 || j.mode > 1) return fail(ORBX_ERR_INVALID, "bad job");
        if (j.n > 10000) return fail(ORBX_ERR_UNSUPPORTED, "more than 10000 keypoints");
        if ((j.n > 0 && (!j.keypoints || !j.descriptors)) || (j.npts > 0 && (!j.pt_xyz || !j.pt_dist || !j.pt_descriptors || !j.pt_flags)) ||
            (j.npts > 0 && j.mode == 1 && !j.pt_normal) || (j.npts > 0 && j.mode == 0 && check_orientation && !j.pt_angle))
            return fail(ORBX_ERR_INVALID, "NULL array in job");
        tot += (size_t)j.npts; max_n = std::max(max_n, j.n);
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_j = al((size_t)njobs * sizeof(OrbxSeqProjDev)), b_q = al(std::max<size_t>(tot, 1) * sizeof(OrbxProjQuery)), b_a = al(std::max<size_t>(tot, 1) * 4);
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, b_j + b_q + b_a, st));
    OrbxProjQuery* d_q = (OrbxProjQuery*)(pool + b_j);
    int* d_a = (int*)(pool + b_j + b_q);
    std::vector<OrbxSeqProjDev> hj(njobs);
    size_t off = 0;
    for (int p = 0; p < njobs; p++) {
        const OrbxProjectionJob& j = jobs[p];
        OrbxSeqProjDev& d = hj[p];
        d.kps = (const OrbxKp28*)j.keypoints; d.desc = j.descriptors; d.occupied = j.occupied; d.n = j.n;
        memcpy(d.Tcw, j.Tcw, sizeof d.Tcw); memcpy(d.Ow, j.Ow, sizeof d.Ow); d.th = j.th; d.mode = j.mode; d.th_dist = j.max_dist;
        d.check_orientation = check_orientation;
        d.pt_xyz = j.pt_xyz; d.pt_normal = j.pt_normal; d.pt_dist = j.pt_dist; d.pt_desc = j.pt_descriptors; d.pt_flags = j.pt_flags;
        d.pt_angle = j.pt_angle; d.npts = j.npts; d.match = j.match; d.nmatches = j.nmatches;
        d.query = d_q + off; d.assign = d_a + off; off += (size_t)j.npts;
    }
    cudaError_t e;
    do {
        if ((e = cudaMemcpyAsync(pool, hj.data(), (size_t)njobs * sizeof(OrbxSeqProjDev), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        orbx_launch_seq_projection((const OrbxSeqProjDev*)pool, njobs, max_n, cam, st);
        e = cudaGetLastError();
    } while (0);
    cudaFreeAsync(pool, st);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_search_by_projection_kf(const OrbxProjectionJob* job, const float* camera9, const float* scale_factors,
                                            int nlevels, float log_scale_factor, int check_orientation, int device)
{
    if (!job || !job->match || !job->nmatches) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    const int n = job->n, np = job->npts;
    if (n < 0 || np < 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    orbx_keep_mempool(device);
    HostPack P;
    const size_t i_kp = P.add((size_t)n * 28), i_d = P.add((size_t)n * 32), i_oc = P.add(n), i_x = P.add((size_t)np * 12),
                 i_nr = P.add((size_t)np * 12), i_ds = P.add((size_t)np * 12), i_pd = P.add((size_t)np * 32), i_pf = P.add(np),
                 i_pa = P.add((size_t)np * 4), i_m = P.add((size_t)n * 4), i_nm = P.add(4);
    CK(cudaMallocAsync(&P.pool, P.tot, 0));
    OrbxProjectionJob d = *job;
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        auto up = [&](size_t i, const void* src, size_t bytes) { return (!src || !bytes) ? cudaSuccess : cudaMemcpyAsync(P.at(i), src, bytes, cudaMemcpyHostToDevice, 0); };
        if ((e = up(i_kp, job->keypoints, (size_t)n * 28)) != cudaSuccess) break;
        if ((e = up(i_d, job->descriptors, (size_t)n * 32)) != cudaSuccess) break;
        if ((e = up(i_oc, job->occupied, (size_t)n)) != cudaSuccess) break;
        if ((e = up(i_x, job->pt_xyz, (size_t)np * 12)) != cudaSuccess) break;
        if ((e = up(i_nr, job->pt_normal, (size_t)np * 12)) != cudaSuccess) break;
        if ((e = up(i_ds, job->pt_dist, (size_t)np * 12)) != cudaSuccess) break;
        if ((e = up(i_pd, job->pt_descriptors, (size_t)np * 32)) != cudaSuccess) break;
        if ((e = up(i_pf, job->pt_flags, (size_t)np)) != cudaSuccess) break;
        if ((e = up(i_pa, job->pt_angle, (size_t)np * 4)) != cudaSuccess) break;
        d.keypoints = (const OrbxKeyPoint*)P.at(i_kp); d.descriptors = P.at(i_d); d.occupied = job->occupied ? P.at(i_oc) : nullptr;
        d.pt_xyz = (const float*)P.at(i_x); d.pt_normal = job->pt_normal ? (const float*)P.at(i_nr) : nullptr;
        d.pt_dist = (const float*)P.at(i_ds); d.pt_descriptors = P.at(i_pd); d.pt_flags = P.at(i_pf);
        d.pt_angle = job->pt_angle ? (const float*)P.at(i_pa) : nullptr;
        d.match = (int32_t*)P.at(i_m); d.nmatches = (int32_t*)P.at(i_nm);
        rc = orbx_search_by_projection_kf_device(&d, 1, camera9, scale_factors, nlevels, log_scale_factor, check_orientation, device, nullptr);
        if (rc != ORBX_OK) break;
        if (n > 0 && (e = cudaMemcpy(job->match, P.at(i_m), (size_t)n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(job->nmatches, P.at(i_nm), 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(P.pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

extern "C" int orbx_search_by_sim3_device(const OrbxSim3KeyFrame* kf1, const OrbxSim3KeyFrame* kf2, const float* S12, const float* S21,
                                          const float* camera9, const float* scale_factors, int nlevels, float log_scale_factor,
                                          float th, int32_t* d_match12, int32_t* d_nfound, int32_t* d_scratch, int device,
                                          void* cuda_stream)
{
    if (!kf1 || !kf2 || !S12 || !S21 || !d_match12 || !d_nfound || !d_scratch) return fail(ORBX_ERR_INVALID, "bad argument");
    OrbxFuseCam cam;
    int rc = make_fuse_cam(camera9, scale_factors, nullptr, nlevels, log_scale_factor, cam);
    if (rc != ORBX_OK) return rc;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    const OrbxSim3KeyFrame* K[2] = {kf1, kf2};
    for (const OrbxSim3KeyFrame* k : K) {
        if (k->n < 0 || k->n > 10000) return fail(ORBX_ERR_UNSUPPORTED, "keyframe feature count out of range (0..10000)");
        if (k->n > 0 && (!k->keypoints || !k->descriptors || !k->mp_xyz || !k->mp_dist || !k->mp_descriptors || !k->mp_flags))
            return fail(ORBX_ERR_INVALID, "NULL array in keyframe");
    }
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int32_t* d_m1 = d_scratch; int32_t* d_m2 = d_scratch + kf1->n; int32_t* d_cnt = d_m2 + kf2->n;
    // direction 1: the points of keyframe 1 into keyframe 2 (:1281-1359); direction 2: the reverse (:1362-1440)
    OrbxFuseDev hj[2] = {};
    for (int dir = 0; dir < 2; dir++) {
        const OrbxSim3KeyFrame* src = K[dir]; const OrbxSim3KeyFrame* dst = K[1 - dir];
        OrbxFuseDev& d = hj[dir];
        d.kps = (const OrbxKp28*)dst->keypoints; d.desc = dst->descriptors; d.u_right = nullptr; d.n = dst->n;
        memcpy(d.Tcw, src->Tcw, sizeof d.Tcw); memcpy(d.T2, dir == 0 ? S21 : S12, sizeof d.T2);
        d.th = th; d.mode = 2;
        d.pt_xyz = src->mp_xyz; d.pt_normal = nullptr; d.pt_dist = src->mp_dist; d.pt_desc = src->mp_descriptors; d.pt_flags = src->mp_flags;
        d.npts = src->n; d.best_idx = dir == 0 ? d_m1 : d_m2; d.best_dist = nullptr; d.nfound = d_cnt + dir;
    }
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, sizeof hj, st));
    cudaError_t e;
    do {
        if ((e = cudaMemcpyAsync(pool, hj, sizeof hj, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        orbx_launch_fuse_search((const OrbxFuseDev*)pool, 2, std::max(kf1->n, kf2->n), cam, st);
        orbx_launch_sim3_mutual(d_m1, kf1->n, d_m2, kf2->n, d_match12, d_nfound, st);
        e = cudaGetLastError();
    } while (0);
    cudaFreeAsync(pool, st);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_search_by_sim3(const OrbxSim3KeyFrame* kf1, const OrbxSim3KeyFrame* kf2, const float* S12, const float* S21,
                                   const float* camera9, const float* scale_factors, int nlevels, float log_scale_factor, float th,
                                   int32_t* match12, int32_t* nfound, int device)
{
    if (!kf1 || !kf2 || !nfound || (kf1->n > 0 && !match12)) return fail(ORBX_ERR_INVALID, "bad argument");
    if (kf1->n < 0 || kf2->n < 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    HostPack P;
    size_t idx[2][6];
    const OrbxSim3KeyFrame* K[2] = {kf1, kf2};
    for (int k = 0; k < 2; k++) {
        const size_t n = (size_t)K[k]->n;
        idx[k][0] = P.add(n * 28); idx[k][1] = P.add(n * 32); idx[k][2] = P.add(n * 12); idx[k][3] = P.add(n * 12); idx[k][4] = P.add(n * 32); idx[k][5] = P.add(n);
    }
    const size_t i_m = P.add((size_t)kf1->n * 4), i_nf = P.add(4), i_s = P.add(((size_t)kf1->n + kf2->n + 2) * 4);
    CK(cudaMallocAsync(&P.pool, P.tot, 0));
    OrbxSim3KeyFrame d[2] = {*kf1, *kf2};
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        auto up = [&](size_t i, const void* src, size_t bytes) { return (!src || !bytes) ? cudaSuccess : cudaMemcpyAsync(P.at(i), src, bytes, cudaMemcpyHostToDevice, 0); };
        for (int k = 0; k < 2 && e == cudaSuccess; k++) {
            const size_t n = (size_t)K[k]->n;
            if ((e = up(idx[k][0], K[k]->keypoints, n * 28)) != cudaSuccess) break;
            if ((e = up(idx[k][1], K[k]->descriptors, n * 32)) != cudaSuccess) break;
            if ((e = up(idx[k][2], K[k]->mp_xyz, n * 12)) != cudaSuccess) break;
            if ((e = up(idx[k][3], K[k]->mp_dist, n * 12)) != cudaSuccess) break;
            if ((e = up(idx[k][4], K[k]->mp_descriptors, n * 32)) != cudaSuccess) break;
            if ((e = up(idx[k][5], K[k]->mp_flags, n)) != cudaSuccess) break;
            d[k].keypoints = (const OrbxKeyPoint*)P.at(idx[k][0]); d[k].descriptors = P.at(idx[k][1]); d[k].mp_xyz = (const float*)P.at(idx[k][2]);
            d[k].mp_dist = (const float*)P.at(idx[k][3]); d[k].mp_descriptors = P.at(idx[k][4]); d[k].mp_flags = P.at(idx[k][5]);
        }
        if (e != cudaSuccess) break;
        rc = orbx_search_by_sim3_device(&d[0], &d[1], S12, S21, camera9, scale_factors, nlevels, log_scale_factor, th,
                                        (int32_t*)P.at(i_m), (int32_t*)P.at(i_nf), (int32_t*)P.at(i_s), device, nullptr);
        if (rc != ORBX_OK) break;
        if (kf1->n > 0 && (e = cudaMemcpy(match12, P.at(i_m), (size_t)kf1->n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(nfound, P.at(i_nf), 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(P.pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

extern "C" int orbx_search_for_initialization_device(const OrbxInitPair* pairs, int npairs, const float* bounds4, float nnratio,
                                                     int check_orientation, int device, void* cuda_stream)
{
    if (npairs <= 0) return ORBX_OK;
    if (!pairs || !bounds4) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    size_t tot = 0; int max_n2 = 0;
    for (int p = 0; p < npairs; p++) {
        const OrbxInitPair& q = pairs[p];
        if (q.n1 < 0 || q.n2 < 0 || !q.nmatches || (q.n1 > 0 && (!q.match12 || !q.prev_matched || !q.prev_matched_out))) return fail(ORBX_ERR_INVALID, "bad pair");
        // init_match_kernel keeps F2's grid CSR + vMatchedDistance / vnMatches21 in shared memory: 22 B per F2 keypoint + 12.5 KB,
        // which passes the 227 KB opt-in limit just below 10000 keypoints
        if (q.n2 > 9900) return fail(ORBX_ERR_UNSUPPORTED, "more than 9900 keypoints in F2");
        if ((q.n1 > 0 && (!q.keypoints1 || !q.descriptors1)) || (q.n2 > 0 && (!q.keypoints2 || !q.descriptors2))) return fail(ORBX_ERR_INVALID, "NULL array in pair");
        tot += (size_t)q.n1; max_n2 = std::max(max_n2, q.n2);
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_p = al((size_t)npairs * sizeof(OrbxInitPairDev));
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, b_p + al(std::max<size_t>(tot, 1) * 4 * 10), st));
    int* d_bin = (int*)(pool + b_p);                                          // per F1 keypoint: bin, count, 4 keys, 4 indices
    int* d_cnt = d_bin + std::max<size_t>(tot, 1); int* d_key = d_cnt + std::max<size_t>(tot, 1); int* d_idx = d_key + 4 * std::max<size_t>(tot, 1);
    std::vector<OrbxInitPairDev> hp(npairs);
    size_t off = 0;
    for (int p = 0; p < npairs; p++) {
        const OrbxInitPair& q = pairs[p];
        OrbxInitPairDev& d = hp[p];
        d.kps1 = (const OrbxKp28*)q.keypoints1; d.desc1 = q.descriptors1; d.n1 = q.n1;
        d.kps2 = (const OrbxKp28*)q.keypoints2; d.desc2 = q.descriptors2; d.n2 = q.n2;
        d.prev = q.prev_matched; d.prev_out = q.prev_matched_out; d.window = q.window_size;
        d.match12 = q.match12; d.nmatches = q.nmatches; d.bin_of = d_bin + off;
        d.ncand = d_cnt + off; d.top_key = (unsigned*)d_key + 4 * off; d.top_idx = d_idx + 4 * off; off += (size_t)q.n1;
    }
    cudaError_t e;
    do {
        if ((e = cudaMemcpyAsync(pool, hp.data(), (size_t)npairs * sizeof(OrbxInitPairDev), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        orbx_launch_init_match((const OrbxInitPairDev*)pool, npairs, max_n2, bounds4, nnratio, check_orientation, st);
        e = cudaGetLastError();
    } while (0);
    cudaFreeAsync(pool, st);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_search_for_initialization(const OrbxInitPair* pair, const float* bounds4, float nnratio, int check_orientation,
                                              int device)
{
    if (!pair || !pair->nmatches) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    const int n1 = pair->n1, n2 = pair->n2;
    if (n1 < 0 || n2 < 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    orbx_keep_mempool(device);
    HostPack P;
    const size_t i_k1 = P.add((size_t)n1 * 28), i_d1 = P.add((size_t)n1 * 32), i_k2 = P.add((size_t)n2 * 28), i_d2 = P.add((size_t)n2 * 32),
                 i_pv = P.add((size_t)n1 * 8), i_m = P.add((size_t)n1 * 4), i_nm = P.add(4);
    CK(cudaMallocAsync(&P.pool, P.tot, 0));
    OrbxInitPair d = *pair;
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        auto up = [&](size_t i, const void* src, size_t bytes) { return (!src || !bytes) ? cudaSuccess : cudaMemcpyAsync(P.at(i), src, bytes, cudaMemcpyHostToDevice, 0); };
        if ((e = up(i_k1, pair->keypoints1, (size_t)n1 * 28)) != cudaSuccess) break;
        if ((e = up(i_d1, pair->descriptors1, (size_t)n1 * 32)) != cudaSuccess) break;
        if ((e = up(i_k2, pair->keypoints2, (size_t)n2 * 28)) != cudaSuccess) break;
        if ((e = up(i_d2, pair->descriptors2, (size_t)n2 * 32)) != cudaSuccess) break;
        if ((e = up(i_pv, pair->prev_matched, (size_t)n1 * 8)) != cudaSuccess) break;
        d.keypoints1 = (const OrbxKeyPoint*)P.at(i_k1); d.descriptors1 = P.at(i_d1); d.keypoints2 = (const OrbxKeyPoint*)P.at(i_k2);
        d.descriptors2 = P.at(i_d2); d.prev_matched = (const float*)P.at(i_pv); d.prev_matched_out = (float*)P.at(i_pv);
        d.match12 = (int32_t*)P.at(i_m); d.nmatches = (int32_t*)P.at(i_nm);
        rc = orbx_search_for_initialization_device(&d, 1, bounds4, nnratio, check_orientation, device, nullptr);
        if (rc != ORBX_OK) break;
        if (n1 > 0 && (e = cudaMemcpy(pair->match12, P.at(i_m), (size_t)n1 * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (n1 > 0 && (e = cudaMemcpy(pair->prev_matched_out, P.at(i_pv), (size_t)n1 * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(pair->nmatches, P.at(i_nm), 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(P.pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}


// Frame::isInFrustum (Frame.cc:315-378) — orbx_match.cu
extern "C" int orbx_is_in_frustum_device(const float* Tcw12, const float* Ow3, const float* camera9, int nlevels, float log_scale_factor,
                                         const float* d_pt_xyz, const float* d_pt_normal, const float* d_pt_dist, int npts,
                                         float viewing_cos_limit, OrbxTrackQuery* d_queries, uint8_t* d_in_view, int device,
                                         void* cuda_stream)
{
    if (npts <= 0) return ORBX_OK;
    if (!Tcw12 || !Ow3 || !d_pt_xyz || !d_pt_normal || !d_pt_dist || !d_queries || !d_in_view) return fail(ORBX_ERR_INVALID, "bad argument");
    std::vector<float> sf1((size_t)std::max(nlevels, 1), 1.f);
    OrbxFuseCam cam;
    int rc = make_fuse_cam(camera9, sf1.data(), nullptr, nlevels, log_scale_factor, cam);
    if (rc != ORBX_OK) return rc;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    OrbxFrustumArgs a;
    memcpy(a.Tcw, Tcw12, sizeof a.Tcw); memcpy(a.Ow, Ow3, sizeof a.Ow); a.view_cos_limit = viewing_cos_limit;
    a.pt_xyz = d_pt_xyz; a.pt_normal = d_pt_normal; a.pt_dist = d_pt_dist; a.npts = npts;
    a.q = (OrbxTrackQueryDev*)d_queries; a.in_view = d_in_view;
    orbx_launch_in_frustum(a, cam, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_is_in_frustum(const float* Tcw12, const float* Ow3, const float* camera9, int nlevels, float log_scale_factor,
                                  const float* pt_xyz, const float* pt_normal, const float* pt_dist, int npts, float viewing_cos_limit,
                                  OrbxTrackQuery* queries, uint8_t* in_view, int device)
{
    if (npts <= 0) return ORBX_OK;
    if (!pt_xyz || !pt_normal || !pt_dist || !queries || !in_view) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    HostPack P;
    const size_t i_x = P.add((size_t)npts * 12), i_n = P.add((size_t)npts * 12), i_d = P.add((size_t)npts * 12),
                 i_q = P.add((size_t)npts * sizeof(OrbxTrackQuery)), i_v = P.add(npts);
    CK(cudaMallocAsync(&P.pool, P.tot, 0));
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        if ((e = cudaMemcpyAsync(P.at(i_x), pt_xyz, (size_t)npts * 12, cudaMemcpyHostToDevice, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(P.at(i_n), pt_normal, (size_t)npts * 12, cudaMemcpyHostToDevice, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(P.at(i_d), pt_dist, (size_t)npts * 12, cudaMemcpyHostToDevice, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(P.at(i_q), queries, (size_t)npts * sizeof(OrbxTrackQuery), cudaMemcpyHostToDevice, 0)) != cudaSuccess) break;
        rc = orbx_is_in_frustum_device(Tcw12, Ow3, camera9, nlevels, log_scale_factor, (const float*)P.at(i_x), (const float*)P.at(i_n),
                                       (const float*)P.at(i_d), npts, viewing_cos_limit, (OrbxTrackQuery*)P.at(i_q), P.at(i_v), device, nullptr);
        if (rc != ORBX_OK) break;
        if ((e = cudaMemcpy(queries, P.at(i_q), (size_t)npts * sizeof(OrbxTrackQuery), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(in_view, P.at(i_v), (size_t)npts, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(P.pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}


// orbx_capi_frame.cu — a device-resident Frame (include/orbx.h, orbx_frame_*): what Frame::Frame (Frame.cc:62-123) builds
// from the extractor's output — mvKeys, mDescriptors, mvKeysUn (UndistortKeyPoints), mvuRight — stays in HBM, in one arena
// per handle together with the per-call scratch of the matchers, so a matcher call moves only what the SLAM thread really
// produces (the projected map points) up and the assignments down: ONE packed H2D copy from a pinned staging buffer, one
// kernel, ONE D2H copy into a pinned mirror, one stream synchronisation; no allocation on the call path.
// Every handle owns a stream: calls on different handles (host threads, frames in flight) run concurrently on the GPU.
#include "orbx_capi_common.cuh"

struct orbx_frame {
    int device = 0, nmax = 0, qmax = 0, n = 0;
    cudaStream_t st = nullptr;
    uint8_t* d_arena = nullptr;      // device
    uint8_t* h_up = nullptr;         // pinned: the packed upload of one call (mirrors the device layout from o_up on)
    uint8_t* h_dn = nullptr;         // pinned: match[nmax] + nmatches
    void* h_dn_dev = nullptr;        // the device's address of h_dn (the matcher kernel stores its result there)
    // device offsets
    size_t o_kp = 0, o_kpun = 0, o_desc = 0, o_ur = 0, o_match = 0, o_assign = 0, o_up = 0;
    // offsets inside the upload block
    size_t u_fd = 0, u_sf = 0, u_occ = 0, u_q = 0, u_qd = 0, u_qf = 0, up_bytes = 0;
    bool has_stereo = false;
    OrbxKp28* kp() const { return (OrbxKp28*)(d_arena + o_kp); }
    OrbxKp28* kpun() const { return (OrbxKp28*)(d_arena + o_kpun); }
    uint8_t* desc() const { return d_arena + o_desc; }
    float* ur() const { return (float*)(d_arena + o_ur); }
};

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" void orbx_frame_destroy(orbx_frame* f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->st) { cudaStreamSynchronize(f->st); cudaStreamDestroy(f->st); }
    cudaFree(f->d_arena);
    if (f->h_up) cudaFreeHost(f->h_up);
    if (f->h_dn) cudaFreeHost(f->h_dn);
    cudaGetLastError();
    delete f;
}

extern "C" int orbx_frame_create(int device, int max_keypoints, int max_queries, orbx_frame** out)
{
    if (!out) return fail(ORBX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (max_keypoints <= 0 || max_queries < 0) return fail(ORBX_ERR_INVALID, "max_keypoints > 0 and max_queries >= 0 required");
    if (max_keypoints > 10000) return fail(ORBX_ERR_UNSUPPORTED, "more than 10000 keypoints in a frame (the matchers' shared-memory grid)");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_frame* f = new orbx_frame();
    f->device = device; f->nmax = max_keypoints; f->qmax = std::max(max_queries, 1);
    const size_t N = (size_t)f->nmax, Q = (size_t)f->qmax;
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += al256(bytes); return o; };
    f->o_kp = carve(N * 28); f->o_kpun = carve(N * 28); f->o_desc = carve(N * 32); f->o_ur = carve(N * 4);
    f->o_match = carve(N * 4 + 256); f->o_assign = carve(Q * 4);
    f->o_up = off;
    size_t u = 0;
    auto ucarve = [&](size_t bytes) { const size_t o = u; u += al256(bytes); return o; };
    f->u_fd = ucarve(sizeof(OrbxLocalFrameDev)); f->u_sf = ucarve(ORBX_MAX_LEVELS * 4); f->u_occ = ucarve(N);
    f->u_q = ucarve(Q * sizeof(OrbxTrackQuery)); f->u_qd = ucarve(Q * 32); f->u_qf = ucarve(Q);
    f->up_bytes = u;
    off += u;
    cudaError_t e;
    if ((e = cudaMalloc(&f->d_arena, off)) != cudaSuccess || (e = cudaMallocHost(&f->h_up, f->up_bytes)) != cudaSuccess ||
        (e = cudaMallocHost(&f->h_dn, N * 4 + 256)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking)) != cudaSuccess) {
        orbx_frame_destroy(f);
        return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    }
    if ((e = cudaHostGetDevicePointer(&f->h_dn_dev, f->h_dn, 0)) != cudaSuccess) { orbx_frame_destroy(f); return fail(ORBX_ERR_CUDA, cudaGetErrorString(e)); }
    orbx_keep_mempool(device);
    *out = f;
    return ORBX_OK;
}

static int frame_fill(orbx_frame* f, const OrbxKp28* d_kps, const uint8_t* d_desc, int n, const float* K4, const float* dist, int ndist,
                      cudaStream_t producer)
{
    if (n < 0 || n > f->nmax) return fail(ORBX_ERR_CAPACITY, "more keypoints than the frame handle was created for");
    CK(cudaSetDevice(f->device));
    f->n = n; f->has_stereo = false;
    if (n == 0) return ORBX_OK;
    // the producer's results are complete once its stream has drained up to here: order this handle's stream behind it
    if (producer != f->st) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, producer);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(f->st, ev, 0);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(f->kp(), d_kps, (size_t)n * 28, cudaMemcpyDeviceToDevice, f->st));
    CK(cudaMemcpyAsync(f->desc(), d_desc, (size_t)n * 32, cudaMemcpyDeviceToDevice, f->st));
    if (ndist <= 0 || !dist || dist[0] == 0.0f) {                              // Frame.cc:474-478: mvKeysUn = mvKeys
        CK(cudaMemcpyAsync(f->kpun(), f->kp(), (size_t)n * 28, cudaMemcpyDeviceToDevice, f->st));
        return ORBX_OK;
    }
    return orbx_undistort_keypoints_device((const OrbxKeyPoint*)f->kp(), n, K4, dist, ndist, (OrbxKeyPoint*)f->kpun(), f->st);
}

extern "C" int orbx_frame_from_extract(orbx_frame* f, orbx_extractor* ex, int frame_index, int n, const float* K4, const float* dist, int ndist)
{
    if (!f || !ex || (ndist > 0 && (!K4 || !dist))) return fail(ORBX_ERR_INVALID, "NULL argument");
    const OrbxKp28* d_kps = nullptr; const uint8_t* d_desc = nullptr; cudaStream_t st = nullptr; int dev = 0;
    const int rc = orbx_internal_results(ex, frame_index, &d_kps, &d_desc, &st, &dev);
    if (rc != ORBX_OK) return rc;
    if (dev != f->device) return fail(ORBX_ERR_INVALID, "extractor and frame handle live on different devices");
    return frame_fill(f, d_kps, d_desc, n, K4, dist, ndist, st);
}

extern "C" int orbx_frame_from_device(orbx_frame* f, const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, int n, const float* K4,
                                      const float* dist, int ndist, void* producer_stream)
{
    if (!f || (n > 0 && (!d_keypoints || !d_descriptors)) || (ndist > 0 && (!K4 || !dist))) return fail(ORBX_ERR_INVALID, "NULL argument");
    return frame_fill(f, (const OrbxKp28*)d_keypoints, d_descriptors, n, K4, dist, ndist, (cudaStream_t)producer_stream);
}

extern "C" int orbx_frame_set_stereo(orbx_frame* f, const float* u_right, int on_device)
{
    if (!f) return fail(ORBX_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(f->device));
    f->has_stereo = u_right != nullptr;
    if (u_right && f->n > 0)
        CK(cudaMemcpyAsync(f->ur(), u_right, (size_t)f->n * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, f->st));
    if (u_right && !on_device) CK(cudaStreamSynchronize(f->st));        // the caller's (pageable) array may go away
    return ORBX_OK;
}

extern "C" int orbx_frame_size(const orbx_frame* f) { return f ? f->n : 0; }

extern "C" int orbx_frame_keypoints(orbx_frame* f, OrbxKeyPoint* keypoints_un, uint8_t* descriptors, int cap, int* n)
{
    if (!f || !n) return fail(ORBX_ERR_INVALID, "NULL argument");
    *n = f->n;
    if (f->n > cap) return fail(ORBX_ERR_CAPACITY, "buffer too small");
    CK(cudaSetDevice(f->device));
    if (f->n > 0 && keypoints_un) CK(cudaMemcpyAsync(keypoints_un, f->kpun(), (size_t)f->n * 28, cudaMemcpyDeviceToHost, f->st));
    if (f->n > 0 && descriptors) CK(cudaMemcpyAsync(descriptors, f->desc(), (size_t)f->n * 32, cudaMemcpyDeviceToHost, f->st));
    CK(cudaStreamSynchronize(f->st));
    return ORBX_OK;
}

extern "C" int orbx_frame_device_arrays(orbx_frame* f, const OrbxKeyPoint** d_keypoints, const OrbxKeyPoint** d_keypoints_un,
                                        const uint8_t** d_descriptors, const float** d_u_right, void** cuda_stream)
{
    if (!f) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (d_keypoints) *d_keypoints = (const OrbxKeyPoint*)f->kp();
    if (d_keypoints_un) *d_keypoints_un = (const OrbxKeyPoint*)f->kpun();
    if (d_descriptors) *d_descriptors = f->desc();
    if (d_u_right) *d_u_right = f->has_stereo ? f->ur() : nullptr;
    if (cuda_stream) *cuda_stream = f->st;
    return ORBX_OK;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-142) against the resident frame
extern "C" int orbx_frame_search_local_points(orbx_frame* f, const OrbxTrackQuery* queries, const uint8_t* query_descriptors,
                                              const uint8_t* query_flags, int nq, const uint8_t* occupied, const float* bounds4,
                                              const float* scale_factors, int nlevels, float th, float nnratio, int32_t* match,
                                              int32_t* nmatches)
{
    if (!f || !bounds4 || !scale_factors || !nmatches || nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "bad argument");
    if (nq < 0 || nq > f->qmax || (nq > 0 && (!queries || !query_descriptors || !query_flags)) || (f->n > 0 && !match))
        return fail(ORBX_ERR_INVALID, "bad query arrays (or more queries than the handle was created for)");
    *nmatches = 0;
    const int n = f->n;
    for (int i = 0; i < n; i++) match[i] = -1;
    if (n == 0 || nq == 0) return ORBX_OK;
    CK(cudaSetDevice(f->device));
    // pack the call's upload: the frame descriptor for the kernel, the scale factors, the occupancy flags, the queries
    uint8_t* d_up = f->d_arena + f->o_up;
    OrbxLocalFrameDev fd;
    fd.kps = f->kpun(); fd.desc = f->desc(); fd.u_right = f->has_stereo ? f->ur() : nullptr;
    fd.occupied = occupied ? d_up + f->u_occ : nullptr; fd.n = n;
    fd.q = (const OrbxTrackQueryDev*)(d_up + f->u_q); fd.qdesc = d_up + f->u_qd; fd.qflags = d_up + f->u_qf; fd.nq = nq;
    fd.match = (int*)(f->d_arena + f->o_match); fd.nmatches = fd.match + n; fd.assign = (int*)(f->d_arena + f->o_assign);
    fd.result_out = (int*)f->h_dn_dev;
    memcpy(f->h_up + f->u_fd, &fd, sizeof fd);
    memcpy(f->h_up + f->u_sf, scale_factors, (size_t)nlevels * 4);
    if (occupied) memcpy(f->h_up + f->u_occ, occupied, (size_t)n);
    memcpy(f->h_up + f->u_q, queries, (size_t)nq * sizeof(OrbxTrackQuery));
    memcpy(f->h_up + f->u_qd, query_descriptors, (size_t)nq * 32);
    memcpy(f->h_up + f->u_qf, query_flags, (size_t)nq);
    // one copy covers the block up to the end of the flags actually used
    const size_t used = f->u_qf + (size_t)nq;
    CK(cudaMemcpyAsync(d_up, f->h_up, used, cudaMemcpyHostToDevice, f->st));
    orbx_launch_local_points((const OrbxLocalFrameDev*)(d_up + f->u_fd), 1, n, bounds4, (const float*)(d_up + f->u_sf), nlevels, th, nnratio, f->st);
    CK(cudaGetLastError());
    // the kernel stores match[0..n) and the count straight into the pinned host mirror (result_out): the call is one upload,
    // one launch and one synchronisation — with 16 host threads the driver's per-call lock bounds the rate, not the GPU
    CK(cudaStreamSynchronize(f->st));
    memcpy(match, f->h_dn, (size_t)n * 4);
    memcpy(nmatches, f->h_dn + (size_t)n * 4, 4);
    return ORBX_OK;
}

// orbx_capi.cu — host side of liborbx.so: the C ABI of include/orbx.h, the per-instance HBM working set,
// and the geometry tables (pyramid sizes, FAST cell grid, resize coefficients, quadtree roots, per-level quotas)
// built with the reference's own float/double expressions (ORBextractor.cc:416-490, 829-879, 567-570, 1221-1225).
// No OpenCV, no torch, no CPU fallback: without a CUDA device every compute entry point fails with ORBX_ERR_CUDA.
#include "orbx_capi_common.cuh"
#include "orbx_tma.cuh"
#include <nvtx3/nvToolsExt.h>      // header-only: loads the profiler's injection library at run time, a no-op without one

#include <mutex>

thread_local std::string g_orbx_err;

struct orbx_extractor {
    int nfeatures, nlevels, ini_th, min_th, device;
    double scale_factor;                          // ORBextractor.h keeps it as double
    std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[16];
    // reserved geometry
    int W = 0, H = 0, max_batch = 0;
    std::vector<OrbxLevelGeom> lvl;
    std::vector<OrbxCell> cells;
    std::vector<OrbxResizeTap> taps;
    int max_tile_w = 0, max_tile_h = 0;
    OrbxFrameLayout L{};
    OrbxTmaps tm_fast{}, tm_desc{};               // per-level tensor maps of the raw pyramid (TMA staging of FAST tiles / describe patches)
    OrbxTmaps tm_pyr{};                           // m[l] describes level l-1 with the source box of level l's resize tiles
    OrbxTmaps tm_blur_src{}, tm_desc_blur{};      // raw pyramid with the blur units' boxes / blurred pyramid with the descriptor's box
    std::vector<OrbxBlurUnit> blur_units; OrbxBlurUnit* d_blur_units = nullptr;
    int blur_rows[ORBX_MAX_LEVELS] = {};          // output rows per blur unit of a level
    std::vector<int> pyr_tiles; int* d_pyr_tiles = nullptr;
    std::vector<uint16_t> qt_path; uint16_t* d_qt_path = nullptr;   // quadtree path-code tables (orbx_qt_path_tables)
    // device memory
    void* d_pool = nullptr;                       // one allocation carved into the arrays of L
    OrbxLevelGeom* d_lvl = nullptr;
    OrbxCell* d_cells = nullptr;
    OrbxResizeTap* d_taps = nullptr;
    uint8_t* d_in = nullptr;                      // [B][H][W(*channels)] staging of host frames
    size_t d_in_bytes = 0;
    OrbxKp28* d_kps = nullptr;                    // [B][kp_cap_total]
    uint8_t* d_desc = nullptr;
    int* d_nkp = nullptr;
    cudaStream_t stream = nullptr;
    static constexpr int MAX_SLOTS = 8;
    cudaStream_t slot_stream[MAX_SLOTS] = {};   // host-path double buffering (copy/compute overlap)
    cudaStream_t copy_stream = nullptr;          // host batch path: all uploads of a call, in order, ahead of the kernels
    cudaStream_t side_stream = nullptr;          // the dense blur of a pipeline run, forked next to its quadtree (run_pipeline)
    static constexpr int FORK_RING = 32;
    cudaEvent_t fork_ev[FORK_RING][2] = {}; unsigned fork_ring = 0;
    struct Pending { int n = 0, cap = 0, nslots = 0; size_t fbytes = 0; const int* nkp = nullptr; const int* nkp2 = nullptr; cudaEvent_t done[MAX_SLOTS] = {}; };
    Pending pending[2]; int npending = 0;        // orbx_extract_batch_begin / _end: batches in flight, oldest first
    // a stereo batch in flight is recorded on the LEFT handle but runs on the right handle's working set too: the right
    // handle points at the left one while that is so, and every entry point on either handle completes the batch first
    orbx_extractor* stereo_owner = nullptr;      // set on the right handle
    orbx_extractor* stereo_partner = nullptr;    // set on the left handle
    std::vector<cudaEvent_t> in_ready, in_free;  // per input buffer: upload finished / kernels that read it finished
    // single-frame host calls are launch-bound (11 small kernels): after the first call with a given input form the
    // kernel sequence is replayed from a CUDA graph (one launch instead of eleven)
    cudaGraphExec_t g1 = nullptr;
    int g1_channels = -1, g1_rgb = -1, g1_rect = -1, g1_stride = -1, g1_seen = 0;
    cudaStream_t g1_stream = nullptr;
    uint2* d_remap = nullptr;                     // fixed-point rectification map (orbx_set_rectify_maps)
    int map_w = 0, map_h = 0, map_src_w = 0, map_src_h = 0;
    int pyr_base = 0;                             // first working-set frame of the last pipeline run
    void* stereo_scratch = nullptr; size_t stereo_scratch_bytes = 0;   // SAD per left keypoint (stereo matcher)
    float* stereo_out = nullptr; size_t stereo_out_floats = 0;         // [2][B][kc] mvuRight / mvDepth staging (batched host path)
    int last_frames = 0;                          // frames of the last extract (for the pyramid accessors)
    // host mirror of ONE frame's raw pyramid block (pinned): what the C++ shim's mvImagePyramid views. With mirror_on a
    // single-frame host call downloads the block in the same stream as its kernels (one asynchronous copy, no extra sync)
    uint8_t* mirror = nullptr; size_t mirror_bytes = 0; bool mirror_on = false; int mirror_frame = -1;
    int map_chunk = 0, map_slots = 1;             // host batch path: frame f sits at ((f/chunk) % slots)*chunk + f%chunk
    // working-set index of frame `frame` of the last call, or -1 when a later chunk has reused its slot
    int ws_index(int frame) const
    {
        if (frame < 0 || frame >= last_frames) return -1;
        if (!map_chunk) return pyr_base + frame;
        const int k = frame / map_chunk, nchunks = (last_frames + map_chunk - 1) / map_chunk;
        if (k + map_slots < nchunks) return -1;
        return (k % map_slots) * map_chunk + frame % map_chunk;
    }
    bool constants_ready = false;
    bool timing = false;
    static constexpr int RING = 64;
    cudaEvent_t ev[RING][5] = {};
    long long runs = 0;                          // pipeline runs recorded since timing was enabled
};

// ---- TMA descriptors (orbx_tma.cuh). cuTensorMapEncodeTiled is a driver-API symbol: fetched through the runtime so that
// the library does not link libcuda and still loads on a machine without a driver (tests/test_abi.py).
bool orbx_encode_level_maps_wh(OrbxTmaps* out, const OrbxLevelGeom* lvl, int nlevels, uint8_t* raw, size_t frame_raw_bytes,
                               int frames, const int* box_w, const int* box_h, const char** err);
bool orbx_encode_level_maps(OrbxTmaps* out, const OrbxLevelGeom* lvl, int nlevels, uint8_t* raw, size_t frame_raw_bytes,
                            int frames, int box_w, const int* box_h, const char** err)
{
    int bw[ORBX_MAX_LEVELS];
    for (int l = 0; l < ORBX_MAX_LEVELS; l++) bw[l] = box_w;
    return orbx_encode_level_maps_wh(out, lvl, nlevels, raw, frame_raw_bytes, frames, bw, box_h, err);
}
bool orbx_encode_level_maps_wh(OrbxTmaps* out, const OrbxLevelGeom* lvl, int nlevels, uint8_t* raw, size_t frame_raw_bytes,
                               int frames, const int* box_w, const int* box_h, const char** err)
{
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            if (err) *err = "cuTensorMapEncodeTiled is not available from this driver";
            return false;
        }
        encode = (encode_t)fn;
    }
    for (int l = 0; l < nlevels; l++) {
        const OrbxLevelGeom& g = lvl[l];
        // x = byte column of the level buffer, y = buffer row (apron included), z = frame of the working set
        const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)(g.h + 2 * ORBX_EDGE), (cuuint64_t)frames};
        const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)frame_raw_bytes};     // bytes, dims 1 and 2 (multiples of 16)
        const cuuint32_t box[3] = {(cuuint32_t)box_w[l], (cuuint32_t)box_h[l], 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode(&out->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, raw + g.raw_off, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled rejected a pyramid level descriptor"; return false; }
    }
    return true;
}

static int round_half_even(float v) { return (int)lrintf(v); }   // cvRound
static int finish_all_pending(orbx_extractor* h);                // batches begun with orbx_*_batch_begin (defined below)

extern "C" const char* orbx_last_error(void) { return g_orbx_err.c_str(); }
extern "C" int orbx_abi_version(void) { return ORBX_ABI_VERSION; }
extern "C" int orbx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int orbx_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device,
                           orbx_extractor** out)
{
    if (!out) return fail(ORBX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "nlevels must be in [1,16]");
    if (nfeatures < 0 || !(scaleFactor > 1.0f)) return fail(ORBX_ERR_INVALID, "nfeatures >= 0 and scaleFactor > 1 required");
    orbx_extractor* h = new orbx_extractor();
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->ini_th = iniThFAST; h->min_th = minThFAST; h->device = device;
    h->scale_factor = scaleFactor;
    // scale tables (ORBextractor.cc:421-441)
    h->sf.assign(nlevels, 1.0f); h->sigma2.assign(nlevels, 1.0f);
    for (int i = 1; i < nlevels; i++) {
        h->sf[i] = (float)(h->sf[i - 1] * h->scale_factor);
        h->sigma2[i] = h->sf[i] * h->sf[i];
    }
    h->inv_sf.resize(nlevels); h->inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) { h->inv_sf[i] = 1.0f / h->sf[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    // per-level quota: geometric series, remainder to the last level (ORBextractor.cc:446-457)
    h->quota.assign(nlevels, 0);
    const float factor = (float)(1.0f / h->scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        h->quota[l] = round_half_even(desired);
        sum += h->quota[l];
        desired *= factor;
    }
    h->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // IC_Angle disc (ORBextractor.cc:473-489)
    {
        const int HP = 15;
        int v, v0, vmax = (int)floor(HP * sqrtf(2.f) / 2 + 1), vmin = (int)ceil(HP * sqrtf(2.f) / 2);
        const double hp2 = HP * HP;
        for (v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = HP, v0 = 0; v >= vmin; --v) {
            while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
            h->umax[v] = v0;
            ++v0;
        }
    }
    *out = h;
    return ORBX_OK;
}

static void release_device(orbx_extractor* h)
{
    if (h->device >= 0 && (h->d_pool || h->stream)) cudaSetDevice(h->device);
    if (h->g1) { cudaGraphExecDestroy(h->g1); h->g1 = nullptr; }
    h->g1_seen = 0;
    cudaFree(h->d_pool); h->d_pool = nullptr;
    if (h->mirror) { cudaFreeHost(h->mirror); h->mirror = nullptr; h->mirror_bytes = 0; } h->mirror_frame = -1;
    cudaFree(h->stereo_scratch); h->stereo_scratch = nullptr; h->stereo_scratch_bytes = 0;
    cudaFree(h->stereo_out); h->stereo_out = nullptr; h->stereo_out_floats = 0;
    cudaFree(h->d_lvl); cudaFree(h->d_cells); cudaFree(h->d_taps); cudaFree(h->d_pyr_tiles); cudaFree(h->d_blur_units); cudaFree(h->d_qt_path);
    h->d_qt_path = nullptr; h->d_lvl = nullptr; h->d_cells = nullptr; h->d_taps = nullptr; h->d_pyr_tiles = nullptr; h->d_blur_units = nullptr;
    cudaFree(h->d_in); cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_nkp);
    h->d_in = nullptr; h->d_kps = nullptr; h->d_desc = nullptr; h->d_nkp = nullptr; h->d_in_bytes = 0;
    h->W = h->H = h->max_batch = 0;
}

extern "C" void orbx_destroy(orbx_extractor* h)
{
    if (!h) return;
    if (h->stereo_owner) { finish_all_pending(h); }
    if (h->npending > 0) { cudaSetDevice(h->device); cudaDeviceSynchronize(); h->npending = 0; }   // batches begun and never ended
    if (h->stereo_partner) { h->stereo_partner->stereo_owner = nullptr; h->stereo_partner = nullptr; }
    release_device(h);
    cudaFree(h->d_remap);
    for (int r = 0; r < orbx_extractor::RING; r++)
        for (int i = 0; i < 5; i++) if (h->ev[r][i]) cudaEventDestroy(h->ev[r][i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int j = 0; j < orbx_extractor::MAX_SLOTS; j++) if (h->slot_stream[j]) cudaStreamDestroy(h->slot_stream[j]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    for (auto& e : h->fork_ev) { if (e[0]) cudaEventDestroy(e[0]); if (e[1]) cudaEventDestroy(e[1]); }
    for (auto& pd : h->pending) for (cudaEvent_t e : pd.done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->in_ready) cudaEventDestroy(e);
    for (cudaEvent_t e : h->in_free) cudaEventDestroy(e);
    cudaGetLastError();
    delete h;
}

extern "C" int orbx_get_levels(const orbx_extractor* h) { return h ? h->nlevels : 0; }
extern "C" float orbx_get_scale_factor(const orbx_extractor* h) { return h ? (float)h->scale_factor : 0.f; }
extern "C" int orbx_get_tables(const orbx_extractor* h, float* sf, float* inv_sf, float* s2, float* inv_s2,
                               int32_t* fpl, int32_t* umax16)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    for (int i = 0; i < h->nlevels; i++) {
        if (sf) sf[i] = h->sf[i];
        if (inv_sf) inv_sf[i] = h->inv_sf[i];
        if (s2) s2[i] = h->sigma2[i];
        if (inv_s2) inv_s2[i] = h->inv_sigma2[i];
        if (fpl) fpl[i] = h->quota[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = h->umax[i];
    return ORBX_OK;
}

// cv::resize INTER_LINEAR 8U coefficient table for one axis (OpenCV imgproc/resize.cpp)
static void build_taps(int ssize, int dsize, OrbxResizeTap* out)
{
    const double scale = (double)ssize / (double)dsize;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        out[d].ofs = (short)s;
        out[d].c0 = (short)round_half_even((1.f - f) * 2048.f);
        out[d].c1 = (short)round_half_even(f * 2048.f);
        out[d].pad = 0;
    }
}

static int build_geometry(orbx_extractor* h, int W, int H)
{
    if (W > ORBX_MAX_DIM || H > ORBX_MAX_DIM) return fail(ORBX_ERR_UNSUPPORTED, "image larger than 4096 px");
    const int nl = h->nlevels;
    h->lvl.assign(nl, OrbxLevelGeom{});
    h->cells.clear(); h->taps.clear(); h->pyr_tiles.clear(); h->blur_units.clear(); h->qt_path.clear();
    size_t raw = 0; int slot = 0, cand = 0, kpc = 0, qtcap = 0, hist_ints = 0;
    h->max_tile_w = h->max_tile_h = 8;
    for (int l = 0; l < nl; l++) {
        OrbxLevelGeom& g = h->lvl[l];
        g.w = round_half_even((float)W * h->inv_sf[l]);          // from the ORIGINAL size (ORBextractor.cc:1223)
        g.h = round_half_even((float)H * h->inv_sf[l]);
        // FAST cell grid (ORBextractor.cc:829-847)
        const int maxBX = g.w - ORBX_MINB, maxBY = g.h - ORBX_MINB;
        const float width = (float)(maxBX - ORBX_MINB), height = (float)(maxBY - ORBX_MINB);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        if (nCols < 1 || nRows < 1)
            return fail(ORBX_ERR_UNSUPPORTED, "pyramid level smaller than 62 px: the reference divides by zero here");
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        // quadtree roots (ORBextractor.cc:567-570)
        g.nini = (int)roundf((float)(maxBX - ORBX_MINB) / (float)(maxBY - ORBX_MINB));
        if (g.nini < 1) return fail(ORBX_ERR_UNSUPPORTED, "portrait pyramid level: nIni == 0 in the reference");
        g.hx = (float)(maxBX - ORBX_MINB) / (float)g.nini;
        g.quota = h->quota[l];
        if (g.quota > ORBX_MAX_QUOTA) return fail(ORBX_ERR_UNSUPPORTED, "more than 2040 features on one level");
        g.scale = h->sf[l];
        g.inv_scale = h->inv_sf[l];
        g.kp_size = (float)(int)(31 * h->sf[l]);                  // scaledPatchSize (ORBextractor.cc:925)
        g.pitch = (ORBX_XOFF + g.w + ORBX_EDGE + 3 + 31) & ~31;
        g.raw_off = (int)raw;
        raw += (size_t)g.pitch * (g.h + 2 * ORBX_EDGE);
        raw = (raw + 255) & ~(size_t)255;
        orbx_blur_units(g, l, h->blur_units, &h->blur_rows[l]);
        g.cell0 = (int)h->cells.size(); g.ncols = nCols; g.nrows = nRows;
        g.cand_off = cand;
        int lvl_slots = 0;
        for (int i = 0; i < nRows; i++)
            for (int j = 0; j < nCols; j++) {
                // ROI [iniX,maxX) x [iniY,maxY); cv::FAST scores columns 3..cols-4 of it (ORBextractor.cc:855-879)
                const int iniY = ORBX_MINB + i * hCell, iniX = ORBX_MINB + j * wCell;
                const int maxY = std::min(iniY + hCell + 6, maxBY), maxX = std::min(iniX + wCell + 6, maxBX);
                OrbxCell c{};
                c.level = (short)l;
                c.ex0 = (short)(iniX + 3); c.ex1 = (short)(maxX - 3);
                c.ey0 = (short)(iniY + 3); c.ey1 = (short)(maxY - 3);
                if (iniY >= maxBY - 3 || iniX >= maxBX - 6 || c.ex1 <= c.ex0 || c.ey1 <= c.ey0) { c.ex1 = c.ex0; c.ey1 = c.ey0; }
                const int ew = c.ex1 - c.ex0, eh = c.ey1 - c.ey0;
                c.slot_off = slot;
                c.slot_cap = ((ew + 1) / 2) * ((eh + 1) / 2);    // strict 3x3 NMS: no two survivors are 8-adjacent
                slot += c.slot_cap; lvl_slots += c.slot_cap;
                h->max_tile_w = std::max(h->max_tile_w, ew + 6);
                h->max_tile_h = std::max(h->max_tile_h, eh + 6);
                h->cells.push_back(c);
            }
        {
            int bh = 8;
            for (size_t ci = (size_t)g.cell0; ci < h->cells.size(); ci++) bh = std::max(bh, h->cells[ci].ey1 - h->cells[ci].ey0 + 6);
            for (size_t ci = (size_t)g.cell0; ci < h->cells.size(); ci++) h->cells[ci].box_h = (short)bh;
        }
        g.cand_cap = std::min(lvl_slots, (1 << 23) - 1);
        cand += g.cand_cap;
        g.kp_cap = std::max(g.quota + 3, 4 * g.nini) + 1;
        // quadtree fast path: count pyramid one level deeper than a uniform spread of `quota` leaves needs; bounded by
        // shared memory (<= 12288 ints) and the 16-bit path code the keypoints carry
        {
            int d = 1;
            while ((g.nini << (2 * d)) < 4 * std::max(g.quota, 1)) d++;
            auto ints = [&](int dd) { return (long long)g.nini * (((1ll << (2 * (dd + 1))) - 1) / 3); };
            while (d > 0 && (ints(d) > 12288 || ((long long)g.nini << (2 * d)) > 65536)) d--;
            if (getenv("ORBX_QT_SWEEP_ONLY")) d = 0;
            g.qt_depth = d;
            orbx_qt_path_tables(g, h->qt_path);
            hist_ints = std::max(hist_ints, (int)ints(d));
        }
        qtcap = std::max(qtcap, g.kp_cap);
        h->L.lvl_kp_off[l] = kpc;
        kpc += g.kp_cap;
        // resize taps from level l-1
        g.xtab_off = (int)h->taps.size();
        if (l > 0) {
            h->taps.resize(h->taps.size() + g.w + g.h);
            build_taps(h->lvl[l - 1].w, g.w, h->taps.data() + g.xtab_off);
            g.ytab_off = g.xtab_off + g.w;
            build_taps(h->lvl[l - 1].h, g.h, h->taps.data() + g.ytab_off);
            g.resize_fast = orbx_pyr_fast_ok(g, h->taps.data()) ? 1 : 0;
            orbx_pyr_tiles(g, h->taps.data(), h->pyr_tiles);
            if (g.pyr_box_w > 256 || g.pyr_box_h > 256) g.resize_fast = 0;       // TMA box limit (scale factors above ~1.7)
        } else g.ytab_off = g.xtab_off;
    }
    OrbxFrameLayout& L = h->L;
    L.nlevels = nl; L.ncells = (int)h->cells.size(); L.slot_total = slot; L.cand_total = cand; L.kp_cap_total = kpc;
    L.ini_th = h->ini_th; L.min_th = h->min_th;
    L.frame_raw_bytes = raw;
    L.qt_cap = (qtcap + 31) & ~31;
    L.qt_hist_ints = hist_ints;
    return ORBX_OK;
}

extern "C" int orbx_reserve(orbx_extractor* h, int width, int height, int max_batch)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (width <= 0 || height <= 0 || max_batch <= 0) return fail(ORBX_ERR_INVALID, "width, height, max_batch must be positive");
    if (width == h->W && height == h->H && max_batch <= h->max_batch) return ORBX_OK;
    { const int rc = finish_all_pending(h); if (rc != ORBX_OK) return rc; }   // also a stereo batch begun on the partner handle
    // geometry is pure host arithmetic: reject what the reference cannot run before touching the device
    {
        orbx_extractor probe = *h;          // tables only; device members of the copy are never used or freed
        int rc0 = build_geometry(&probe, width, height);
        if (rc0 != ORBX_OK) return rc0;
    }
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(h->device));
    if (!h->stream) CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (!h->side_stream) {
        CK(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        for (auto& e : h->fork_ev) { CK(cudaEventCreateWithFlags(&e[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e[1], cudaEventDisableTiming)); }
    }
    CK(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < orbx_extractor::MAX_SLOTS; j++) if (h->slot_stream[j]) CK(cudaStreamSynchronize(h->slot_stream[j]));
    release_device(h);
    int rc = build_geometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    OrbxFrameLayout& L = h->L;
    const size_t B = (size_t)max_batch;
    // one pool for the per-frame arrays
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_raw = carve(B * L.frame_raw_bytes);
    const size_t o_blur = carve(B * L.frame_raw_bytes);
    const size_t o_slots = carve(B * L.slot_total * sizeof(uint32_t));
    const size_t o_cc = carve(B * L.ncells * sizeof(int));
    const size_t o_cand = carve(B * L.cand_total * sizeof(uint32_t));
    const size_t o_node = carve(B * L.cand_total * sizeof(uint16_t));
    const size_t o_candc = carve(B * L.nlevels * sizeof(int));
    const size_t o_lkp = carve(B * L.kp_cap_total * sizeof(uint32_t));
    const size_t o_lkpc = carve(B * L.nlevels * sizeof(int));
    CK(cudaMalloc(&h->d_pool, off));
    uint8_t* base = (uint8_t*)h->d_pool;
    L.raw = base + o_raw; L.blur = base + o_blur; L.slots = (uint32_t*)(base + o_slots); L.cell_count = (int*)(base + o_cc);
    L.cand = (uint32_t*)(base + o_cand); L.cand_node = (uint16_t*)(base + o_node); L.cand_count = (int*)(base + o_candc);
    L.lvl_kp = (uint32_t*)(base + o_lkp); L.lvl_kp_count = (int*)(base + o_lkpc);
    CK(cudaMalloc(&h->d_lvl, h->lvl.size() * sizeof(OrbxLevelGeom)));
    CK(cudaMalloc(&h->d_cells, h->cells.size() * sizeof(OrbxCell)));
    CK(cudaMalloc(&h->d_taps, std::max<size_t>(h->taps.size(), 1) * sizeof(OrbxResizeTap)));
    CK(cudaMemcpy(h->d_lvl, h->lvl.data(), h->lvl.size() * sizeof(OrbxLevelGeom), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_cells, h->cells.data(), h->cells.size() * sizeof(OrbxCell), cudaMemcpyHostToDevice));
    if (!h->taps.empty())
        CK(cudaMemcpy(h->d_taps, h->taps.data(), h->taps.size() * sizeof(OrbxResizeTap), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_pyr_tiles, std::max<size_t>(h->pyr_tiles.size(), 2) * sizeof(int)));
    if (!h->pyr_tiles.empty()) CK(cudaMemcpy(h->d_pyr_tiles, h->pyr_tiles.data(), h->pyr_tiles.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_blur_units, h->blur_units.size() * sizeof(OrbxBlurUnit)));
    CK(cudaMemcpy(h->d_blur_units, h->blur_units.data(), h->blur_units.size() * sizeof(OrbxBlurUnit), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_qt_path, std::max<size_t>(h->qt_path.size(), 1) * sizeof(uint16_t)));
    if (!h->qt_path.empty()) CK(cudaMemcpy(h->d_qt_path, h->qt_path.data(), h->qt_path.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    L.lvl = h->d_lvl; L.cells = h->d_cells; L.taps = h->d_taps; L.pyr_tiles = h->d_pyr_tiles; L.qt_path = h->d_qt_path;
    {
        int box_h[ORBX_MAX_LEVELS], box_d[ORBX_MAX_LEVELS], box_b[ORBX_MAX_LEVELS], box_s[ORBX_MAX_LEVELS];
        for (int l = 0; l < h->nlevels; l++) { box_h[l] = h->cells[h->lvl[l].cell0].box_h; box_d[l] = 31; box_b[l] = 37; box_s[l] = h->blur_rows[l] + 6; }
        const char* why = nullptr;
        if (!orbx_encode_level_maps(&h->tm_fast, h->lvl.data(), h->nlevels, L.raw, L.frame_raw_bytes, max_batch,
                                    orbx_fast_tile_pitch(h->max_tile_w), box_h, &why) ||
            !orbx_encode_level_maps(&h->tm_desc, h->lvl.data(), h->nlevels, L.raw, L.frame_raw_bytes, max_batch, 48, box_d, &why) ||
            !orbx_encode_level_maps(&h->tm_desc_blur, h->lvl.data(), h->nlevels, L.blur, L.frame_raw_bytes, max_batch, 64, box_b, &why) ||
            !orbx_encode_level_maps(&h->tm_blur_src, h->lvl.data(), h->nlevels, L.raw, L.frame_raw_bytes, max_batch, ORBX_BLUR_BOX_W, box_s, &why))
            return fail(ORBX_ERR_CUDA, why ? why : "cuTensorMapEncodeTiled failed");
        // the resize tiles of level l read level l-1: describe the source levels, shifted by one, each with its own box
        std::vector<OrbxLevelGeom> src(h->nlevels, h->lvl[0]);
        int pbw[ORBX_MAX_LEVELS], pbh[ORBX_MAX_LEVELS];
        bool any = false;
        for (int l = 0; l < h->nlevels; l++) {
            src[l] = h->lvl[l > 0 ? l - 1 : 0];
            const bool on = l > 0 && h->lvl[l].resize_fast;
            pbw[l] = on ? h->lvl[l].pyr_box_w : 16; pbh[l] = on ? h->lvl[l].pyr_box_h : 1; any = any || on;
        }
        if (any && !orbx_encode_level_maps_wh(&h->tm_pyr, src.data(), h->nlevels, L.raw, L.frame_raw_bytes, max_batch, pbw, pbh, &why))
            return fail(ORBX_ERR_CUDA, why ? why : "cuTensorMapEncodeTiled failed");
    }
    // staging for the host entry points
    const size_t in_bytes = B * (size_t)width * height;
    CK(cudaMalloc(&h->d_in, in_bytes));
    h->d_in_bytes = in_bytes;
    CK(cudaMalloc(&h->d_kps, B * L.kp_cap_total * sizeof(OrbxKp28)));
    CK(cudaMalloc(&h->d_desc, B * L.kp_cap_total * 32));
    CK(cudaMalloc(&h->d_nkp, B * sizeof(int)));
    if (!h->constants_ready) { orbx_upload_constants(); CK(cudaGetLastError()); h->constants_ready = true; }
    h->W = width; h->H = height; h->max_batch = max_batch; h->last_frames = 0;
    return ORBX_OK;
}

extern "C" int orbx_max_keypoints(const orbx_extractor* h) { return (h && h->W) ? h->L.kp_cap_total : 0; }

// the whole device pipeline for n frames that already sit in HBM
// `base` = first frame of the reserved working set to use (the host path runs two chunks concurrently on two
// streams, each in its own half of the working set)
static int run_pipeline(orbx_extractor* h, const uint8_t* d_img, int n, int stride, size_t frame_pitch,
                        OrbxKp28* d_kps, uint8_t* d_desc, int cap, int* d_nkp, cudaStream_t st, int base = 0,
                        int channels = 1, int rgb = 0, bool rectify = false, bool fork_blur = true)
{
    OrbxFrameLayout Lb = h->L;
    Lb.frame0 = base;
    if (base) {
        Lb.raw += (size_t)base * Lb.frame_raw_bytes;
        Lb.blur += (size_t)base * Lb.frame_raw_bytes;
        Lb.slots += (size_t)base * Lb.slot_total;
        Lb.cell_count += (size_t)base * Lb.ncells;
        Lb.cand += (size_t)base * Lb.cand_total;
        Lb.cand_node += (size_t)base * Lb.cand_total;
        Lb.cand_count += (size_t)base * Lb.nlevels;
        Lb.lvl_kp += (size_t)base * Lb.kp_cap_total;
        Lb.lvl_kp_count += (size_t)base * Lb.nlevels;
    }
    const bool tm = h->timing;
    cudaEvent_t* ev = h->ev[h->runs % orbx_extractor::RING];
    // NVTX ranges around the four launch groups (ORBX_NVTX=1): they show up in Nsight Systems / ncu --nvtx timelines
    static const bool nvtx = getenv("ORBX_NVTX") != nullptr;
    struct Range { bool on; explicit Range(bool o, const char* n) : on(o) { if (on) nvtxRangePushA(n); } ~Range() { if (on) nvtxRangePop(); } };
    Range r_all(nvtx, "orbx extract (ORBextractor::operator())");
    if (tm) cudaEventRecord(ev[0], st);
    if (nvtx) nvtxRangePushA("ComputePyramid");
    orbx_launch_pyramid(Lb, h->tm_pyr, h->lvl.data(), d_img, h->W, h->H, stride, frame_pitch, n, st, channels, rgb,
                        rectify ? h->d_remap : nullptr, h->map_src_w, h->map_src_h);
    if (nvtx) { nvtxRangePop(); nvtxRangePushA("FAST cells"); }
    if (tm) cudaEventRecord(ev[1], st);
    orbx_launch_fast(Lb, h->tm_fast, h->max_tile_w, h->max_tile_h, n, st);
    // The dense blur only needs the pyramid, the quadtree only the FAST cells: the blur runs on a side stream next to the
    // quadtree (a latency-bound kernel that leaves issue slots idle) and joins before the descriptors. Inside a stream
    // capture the fork / join become two parallel branches of the graph. The chunked host paths keep the blur in line: they
    // already run up to eight pipelines side by side, which fill each other's gaps, and one side stream would chain them
    // (measured: 165.2 k -> 162.9 k frames/s end to end with the fork, 192.5 k -> 194.6 k resident).
    static const bool blur_serial = getenv("ORBX_BLUR_SERIAL") != nullptr;
    cudaEvent_t* fe = nullptr;
    if (fork_blur && !blur_serial && h->side_stream) {
        fe = h->fork_ev[h->fork_ring++ % orbx_extractor::FORK_RING];
        cudaEventRecord(fe[0], st); cudaStreamWaitEvent(h->side_stream, fe[0], 0);
    }
    if (nvtx) { nvtxRangePop(); nvtxRangePushA("DistributeOctTree"); }
    if (tm) cudaEventRecord(ev[2], st);
    // small frames: 256-thread CTAs so that several (level, frame) trees share an SM and hide each other's barriers
    // ... unless there are so few trees that every one gets an SM to itself anyway (single frames): then the wide CTA
    // finishes a tree sooner
    const bool few = n * h->nlevels <= 148;
    orbx_launch_quadtree(Lb, n, ((size_t)h->W * h->H <= (size_t)1 << 20 && !few) ? 256 : 1024, st);
    if (nvtx) { nvtxRangePop(); nvtxRangePushA("IC_Angle + blur + rBRIEF"); }
    if (tm) cudaEventRecord(ev[3], st);
    if (fe) {
        orbx_launch_blur(Lb, h->tm_blur_src, h->d_blur_units, (int)h->blur_units.size(), n, h->side_stream);
        cudaEventRecord(fe[1], h->side_stream); cudaStreamWaitEvent(st, fe[1], 0);
    } else orbx_launch_blur(Lb, h->tm_blur_src, h->d_blur_units, (int)h->blur_units.size(), n, st);
    orbx_launch_describe(Lb, h->tm_desc, h->tm_desc_blur, n, d_kps, d_desc, cap, d_nkp, st);
    if (nvtx) nvtxRangePop();
    if (tm) cudaEventRecord(ev[4], st);
    if (tm) h->runs++;
    CK(cudaGetLastError());
    h->last_frames = n;
    h->pyr_base = base;
    h->map_chunk = 0;
    return ORBX_OK;
}

// batches begun with orbx_extract_batch_begin share the working set with every other call on the handle: complete them first
static int finish_oldest_pending(orbx_extractor* h);
static int finish_all_pending(orbx_extractor* h)
{
    if (h && h->stereo_owner && h->stereo_owner != h) {          // right handle of a stereo batch in flight
        const int rc = finish_all_pending(h->stereo_owner);
        if (rc != ORBX_OK) return rc;
    }
    while (h && h->npending > 0) { const int rc = finish_oldest_pending(h); if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc; }
    return ORBX_OK;
}

extern "C" int orbx_extract_device(orbx_extractor* h, const uint8_t* d_images, int n, int width, int height, int stride,
                                   size_t frame_pitch_bytes, OrbxKeyPoint* d_keypoints, int cap, int32_t* d_nkp,
                                   uint8_t* d_descriptors, void* cuda_stream)
{
    if (!h || !d_images || !d_keypoints || !d_nkp || !d_descriptors) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (n <= 0 || width <= 0 || height <= 0 || stride < width || cap <= 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    { const int rc = finish_all_pending(h); if (rc != ORBX_OK) return rc; }
    if (width != h->W || height != h->H || n > h->max_batch) {
        int rc = orbx_reserve(h, width, height, std::max(n, h->max_batch));
        if (rc != ORBX_OK) return rc;
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    return run_pipeline(h, d_images, n, stride, frame_pitch_bytes, (OrbxKp28*)d_keypoints, d_descriptors, cap, d_nkp, st);
}

extern "C" int orbx_extract_device_rectified(orbx_extractor* h, const uint8_t* d_images, int n, int stride,
                                             size_t frame_pitch_bytes, OrbxKeyPoint* d_keypoints, int cap, int32_t* d_nkp,
                                             uint8_t* d_descriptors, void* cuda_stream)
{
    if (!h || !d_images || !d_keypoints || !d_nkp || !d_descriptors) return fail(ORBX_ERR_INVALID, "NULL argument");
    if (!h->d_remap) return fail(ORBX_ERR_STATE, "orbx_set_rectify_maps has not been called");
    if (n <= 0 || stride < h->map_src_w || cap <= 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    { const int rc = finish_all_pending(h); if (rc != ORBX_OK) return rc; }
    if (h->map_w != h->W || h->map_h != h->H || n > h->max_batch) {
        int rc = orbx_reserve(h, h->map_w, h->map_h, std::max(n, h->max_batch));
        if (rc != ORBX_OK) return rc;
    }
    CK(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    return run_pipeline(h, d_images, n, stride, frame_pitch_bytes, (OrbxKp28*)d_keypoints, d_descriptors, cap, d_nkp, st, 0, 1, 0, true);
}

extern "C" int orbx_synchronize(orbx_extractor* h)
{
    if (!h || !h->stream) return ORBX_OK;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_enable_timing(orbx_extractor* h, int on)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (on && !h->ev[0][0]) {
        if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible");
        CK(cudaSetDevice(h->device));
        for (int r = 0; r < orbx_extractor::RING; r++)
            for (int i = 0; i < 5; i++) CK(cudaEventCreate(&h->ev[r][i]));
    }
    h->timing = on != 0;
    h->runs = 0;
    return ORBX_OK;
}

extern "C" int orbx_get_stage_ms(orbx_extractor* h, float* ms4, int* nruns)
{
    if (!h || !ms4 || !h->ev[0][0]) return fail(ORBX_ERR_STATE, "timing was never enabled");
    CK(cudaSetDevice(h->device));
    const int n = (int)std::min<long long>(h->runs, orbx_extractor::RING);
    double acc[4] = {0, 0, 0, 0};
    for (int k = 0; k < n; k++) {
        cudaEvent_t* ev = h->ev[(h->runs - 1 - k) % orbx_extractor::RING];
        CK(cudaEventSynchronize(ev[4]));
        for (int i = 0; i < 4; i++) { float ms; CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); acc[i] += ms; }
    }
    for (int i = 0; i < 4; i++) ms4[i] = n ? (float)(acc[i] / n) : 0.f;
    if (nruns) *nruns = n;
    return ORBX_OK;
}

// completes the oldest batch begun with orbx_extract_batch_begin: waits for its last operation on every slot stream
static int finish_oldest_pending(orbx_extractor* h)
{
    if (h->npending <= 0) return ORBX_OK;
    orbx_extractor::Pending& pd = h->pending[0];
    for (int j = 0; j < pd.nslots; j++) CK(cudaEventSynchronize(pd.done[j]));
    int status = ORBX_OK;
    for (int i = 0; i < pd.n; i++) if (pd.nkp[i] > pd.cap || (pd.nkp2 && pd.nkp2[i] > pd.cap)) status = ORBX_ERR_CAPACITY;
    std::swap(h->pending[0], h->pending[1]);                   // keeps both event sets alive
    h->npending--;
    if (h->npending == 0 && h->stereo_partner) { h->stereo_partner->stereo_owner = nullptr; h->stereo_partner = nullptr; }
    if (status == ORBX_ERR_CAPACITY) return fail(status, "keypoint buffer too small (see nkp for the required size)");
    return status;
}

static int extract_batch_impl(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height,
                              int stride, int channels, int rgb, OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors,
                              bool rectify = false, bool begin_only = false)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (n <= 0 || !images || width <= 0 || height <= 0) return ORBX_OK;        // empty input: silent, like :1141
    // with a rectification map the frames have the map's SOURCE size and the pyramid has the map's own size
    const int gw = rectify ? h->map_w : width, gh = rectify ? h->map_h : height;
    if (channels != 1 && channels != 3 && channels != 4) return fail(ORBX_ERR_INVALID, "channels must be 1, 3 or 4 (Tracking.cc:174-199)");
    if (!keypoints || !nkp || !descriptors || cap < 0 || stride < width * channels) return fail(ORBX_ERR_INVALID, "bad output buffers");
    if (h->stereo_owner) { const int rc = finish_all_pending(h); if (rc != ORBX_OK) return rc; }   // right handle of a stereo batch in flight
    if (gw != h->W || gh != h->H || h->max_batch < 1)
        while (h->npending > 0) { const int rc = finish_oldest_pending(h); if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc; }
    if (gw != h->W || gh != h->H || h->max_batch < 1) {
        int rc = orbx_reserve(h, gw, gh, std::max(1, std::min(n, std::max(h->max_batch, 64))));
        if (rc != ORBX_OK) return rc;
    }
    CK(cudaSetDevice(h->device));
    const int B = h->max_batch, kc = h->L.kp_cap_total;
    const size_t fbytes = (size_t)width * height * channels;
    const int rowbytes = width * channels;
    // batches still in flight: a synchronous call, a batch of another shape or a third batch first completes them
    while (h->npending > 0 && (!begin_only || h->npending >= 2 || h->pending[0].n != n || h->pending[0].fbytes != fbytes ||
                               h->pending[0].cap != cap)) {
        const int rc = finish_oldest_pending(h);
        if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc;
    }
    if ((size_t)B * fbytes > h->d_in_bytes) {           // colour frames need a wider staging buffer than gray ones
        CK(cudaDeviceSynchronize());
        cudaFree(h->d_in); h->d_in = nullptr; h->d_in_bytes = 0;
        CK(cudaMalloc(&h->d_in, (size_t)B * fbytes));
        h->d_in_bytes = (size_t)B * fbytes;
    }
    int status = ORBX_OK;
    // Up to eight chunks in flight on as many streams, each in its own slice of the reserved working set: the H2D copy
    // of later chunks and the D2H copy of earlier ones overlap the kernels of the current ones, and short chunks keep
    // the pipeline's fill and drain (one chunk's copy time each) small. A slot's stream serialises its own
    // H2D -> kernels -> D2H, so reusing a slot needs no extra event. ORBX_CHUNK / ORBX_SLOTS override the defaults
    // (measured on B200, 1024 VGA frames per call, tools/exp_chunks.py: 8 x 128 -> 125.8 k frames/s, 8 x 64 -> 132.2 k,
    // 8 x 32 -> 133.7 k, 8 x 16 -> 111.8 k).
    int chunk = B >= 8 ? std::min(32, std::max(8, B / 8)) : B;
    if (const char* e = getenv("ORBX_CHUNK")) { const int v = atoi(e); if (v > 0 && B >= 8) chunk = std::min(v, B / 2); }
    int nslots = B >= 8 ? std::max(1, std::min(orbx_extractor::MAX_SLOTS, B / chunk)) : 1;
    if (const char* e = getenv("ORBX_SLOTS")) { const int v = atoi(e); if (v > 0 && B >= 8) nslots = std::max(1, std::min(std::min(v, orbx_extractor::MAX_SLOTS), B / chunk)); }
    for (int j = 0; j < nslots; j++)
        if (!h->slot_stream[j]) CK(cudaStreamCreateWithFlags(&h->slot_stream[j], cudaStreamNonBlocking));
    CK(cudaStreamSynchronize(h->stream));
    // Uploads run ahead of the kernels: the staging buffer holds B frames, i.e. more chunks than there are compute slots,
    // so every upload of a call goes onto ONE copy stream in order and a slot's stream only waits for its own chunk's
    // event. (With the upload on the slot's stream, chunk k + nslots could not start its copy before chunk k had left the
    // slot; the slots tend to finish together, and their uploads then queued up behind each other while the SMs idled:
    // 7.89 -> 7.74 ms per 1024 VGA frames.) An input buffer is reused once the kernels that read it are done.
    // What remains above the 6.7 ms of the kernels: uploads arrive about as fast as the kernels consume them, so few chunks
    // are runnable at once and a chunk running alone fills the GPU less well than eight do (tools/exp_streams2.py).
    const bool ahead = n > chunk && !getenv("ORBX_NO_COPY_STREAM");
    const int nin = std::max(nslots, B / chunk);
    if (ahead) {
        if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        while ((int)h->in_ready.size() < nin) {
            cudaEvent_t a, b;
            CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            h->in_ready.push_back(a); h->in_free.push_back(b);
        }
    }
    int k = 0;
    for (int f0 = 0; f0 < n; f0 += chunk, k++) {
        const int m = std::min(chunk, n - f0);
        const int slot = k % nslots, base = slot * chunk;
        cudaStream_t st = h->slot_stream[slot];
        const int in_slot = ahead ? k % nin : slot;
        cudaStream_t cst = ahead ? h->copy_stream : st;
        uint8_t* d_in = h->d_in + (size_t)in_slot * chunk * fbytes;
        if (ahead) CK(cudaStreamWaitEvent(cst, h->in_free[in_slot], 0));   // no-op until the buffer has had a reader (also across begun batches)
        OrbxKp28* d_kps = h->d_kps + (size_t)base * kc;
        uint8_t* d_desc = h->d_desc + (size_t)base * kc * 32;
        int* d_nkp = h->d_nkp + base;
        // frames that are contiguous in host memory go up in one copy, otherwise one (strided) copy per frame
        bool contiguous = stride == rowbytes;
        for (int i = 0; i < m && contiguous; i++) {
            if (!images[f0 + i]) return fail(ORBX_ERR_INVALID, "images[i] is NULL");
            contiguous = images[f0 + i] == images[f0] + (size_t)i * fbytes;
        }
        if (contiguous) CK(cudaMemcpyAsync(d_in, images[f0], (size_t)m * fbytes, cudaMemcpyHostToDevice, cst));
        else
            for (int i = 0; i < m; i++) {
                if (!images[f0 + i]) return fail(ORBX_ERR_INVALID, "images[i] is NULL");
                CK(cudaMemcpy2DAsync(d_in + (size_t)i * fbytes, rowbytes, images[f0 + i], stride, rowbytes, height,
                                     cudaMemcpyHostToDevice, cst));
            }
        if (ahead) { CK(cudaEventRecord(h->in_ready[in_slot], cst)); CK(cudaStreamWaitEvent(st, h->in_ready[in_slot], 0)); }
        int rc = ORBX_OK;
        const bool single = n == 1 && base == 0 && !h->timing && !getenv("ORBX_NO_GRAPH");
        const bool same_form = h->g1_channels == channels && h->g1_rgb == rgb && h->g1_rect == (int)rectify &&
                               h->g1_stride == rowbytes && h->g1_stream == st;
        if (single && h->g1 && same_form) {
            CK(cudaGraphLaunch(h->g1, st));
            h->last_frames = 1; h->pyr_base = 0;
        } else if (single && same_form && h->g1_seen >= 1 && !h->g1) {
            // second call of this form: record the kernel sequence once (all arguments are fixed device addresses)
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = run_pipeline(h, d_in, 1, rowbytes, fbytes, d_kps, d_desc, kc, d_nkp, st, 0, channels, rgb, rectify);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != ORBX_OK || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc != ORBX_OK ? rc : fail(ORBX_ERR_CUDA, cudaGetErrorString(ce)); }
            ce = cudaGraphInstantiate(&h->g1, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { h->g1 = nullptr; return fail(ORBX_ERR_CUDA, cudaGetErrorString(ce)); }
            CK(cudaGraphLaunch(h->g1, st));
        } else {
            if (single && !same_form) {
                if (h->g1) { cudaGraphExecDestroy(h->g1); h->g1 = nullptr; }
                h->g1_channels = channels; h->g1_rgb = rgb; h->g1_rect = (int)rectify; h->g1_stride = rowbytes; h->g1_stream = st;
                h->g1_seen = 0;
            }
            rc = run_pipeline(h, d_in, m, rowbytes, fbytes, d_kps, d_desc, kc, d_nkp, st, base, channels, rgb, rectify, n <= chunk);
            if (single) h->g1_seen++;
        }
        if (rc != ORBX_OK) return rc;
        h->mirror_frame = -1;
        if (n == 1 && h->mirror_on) {
            if (!h->mirror) { CK(cudaMallocHost(&h->mirror, h->L.frame_raw_bytes)); h->mirror_bytes = h->L.frame_raw_bytes; }
            CK(cudaMemcpyAsync(h->mirror, h->L.raw + (size_t)base * h->L.frame_raw_bytes, h->L.frame_raw_bytes, cudaMemcpyDeviceToHost, st));
            h->mirror_frame = 0;
        }
        if (ahead) CK(cudaEventRecord(h->in_free[in_slot], st));
        CK(cudaMemcpyAsync(nkp + f0, d_nkp, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (cap == kc) {
            // caller sized its buffers with orbx_max_keypoints(): results land in place with two bulk copies
            CK(cudaMemcpyAsync(keypoints + (size_t)f0 * cap, d_kps, (size_t)m * kc * sizeof(OrbxKp28), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(descriptors + (size_t)f0 * cap * 32, d_desc, (size_t)m * kc * 32, cudaMemcpyDeviceToHost, st));
        } else {
            CK(cudaStreamSynchronize(st));
            for (int i = 0; i < m; i++) {
                const int c = std::min(nkp[f0 + i], cap);
                if (c <= 0) continue;
                CK(cudaMemcpyAsync(keypoints + (size_t)(f0 + i) * cap, d_kps + (size_t)i * kc, (size_t)c * sizeof(OrbxKp28), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(descriptors + (size_t)(f0 + i) * cap * 32, d_desc + (size_t)i * kc * 32, (size_t)c * 32, cudaMemcpyDeviceToHost, st));
            }
        }
    }
    h->last_frames = n; h->pyr_base = 0; h->map_chunk = chunk; h->map_slots = nslots;   // for the pyramid / debug accessors
    if (begin_only) {
        orbx_extractor::Pending& pd = h->pending[h->npending];
        pd.n = n; pd.cap = cap; pd.nslots = nslots; pd.fbytes = fbytes; pd.nkp = nkp; pd.nkp2 = nullptr;
        for (int j = 0; j < nslots; j++) {
            if (!pd.done[j]) CK(cudaEventCreateWithFlags(&pd.done[j], cudaEventDisableTiming));
            CK(cudaEventRecord(pd.done[j], h->slot_stream[j]));
        }
        h->npending++;
        return ORBX_OK;
    }
    if (ahead) CK(cudaStreamSynchronize(h->copy_stream));
    for (int j = 0; j < nslots; j++) CK(cudaStreamSynchronize(h->slot_stream[j]));
    for (int i = 0; i < n; i++) if (nkp[i] > cap) status = ORBX_ERR_CAPACITY;
    if (status == ORBX_ERR_CAPACITY) return fail(status, "keypoint buffer too small (see nkp for the required size)");
    return status;
}

extern "C" int orbx_extract_batch(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height,
                                  int stride, OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors)
{
    return extract_batch_impl(h, images, n, width, height, stride, 1, 0, keypoints, cap, nkp, descriptors);
}

extern "C" int orbx_extract_batch_begin(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height,
                                        int stride, OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors)
{
    if (h && cap != orbx_max_keypoints(h) && h->W == width && h->H == height)
        return fail(ORBX_ERR_INVALID, "orbx_extract_batch_begin needs cap == orbx_max_keypoints() (results are downloaded in bulk)");
    return extract_batch_impl(h, images, n, width, height, stride, 1, 0, keypoints, cap, nkp, descriptors, false, true);
}

extern "C" int orbx_extract_batch_end(orbx_extractor* h)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (h->npending <= 0) return fail(ORBX_ERR_STATE, "no batch in flight");
    CK(cudaSetDevice(h->device));
    return finish_oldest_pending(h);
}

extern "C" int orbx_extract_batch_color(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height,
                                        int stride, int channels, int rgb, OrbxKeyPoint* keypoints, int cap, int* nkp,
                                        uint8_t* descriptors)
{
    return extract_batch_impl(h, images, n, width, height, stride, channels, rgb, keypoints, cap, nkp, descriptors);
}

// cv::initUndistortRectifyMap's CV_32FC1 map pair -> OpenCV's fixed-point form (imgwarp.cpp, INTER_BITS = 5), once.
extern "C" int orbx_set_rectify_maps(orbx_extractor* h, const float* map1, const float* map2, int map_width, int map_height,
                                     int map_stride, int src_width, int src_height)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible");
    { const int rc = finish_all_pending(h); if (rc != ORBX_OK) return rc; }
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    cudaFree(h->d_remap); h->d_remap = nullptr; h->map_w = h->map_h = h->map_src_w = h->map_src_h = 0;
    if (h->g1) { cudaGraphExecDestroy(h->g1); h->g1 = nullptr; }      // a recorded single-frame graph holds the old map
    h->g1_seen = 0; h->g1_rect = -1;
    if (!map1 && !map2) return ORBX_OK;                                        // clear
    if (!map1 || !map2 || map_width <= 0 || map_height <= 0 || map_stride < map_width || src_width <= 0 || src_height <= 0)
        return fail(ORBX_ERR_INVALID, "bad rectification maps");
    if (src_width > 32766 || src_height > 32766) return fail(ORBX_ERR_UNSUPPORTED, "source frame too large for 16-bit map coordinates");
    std::vector<uint2> fx((size_t)map_width * map_height);
    auto sat = [](int v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; };
    for (int y = 0; y < map_height; y++)
        for (int x = 0; x < map_width; x++) {
            const int sx = round_half_even(map1[(size_t)y * map_stride + x] * 32.0f);
            const int sy = round_half_even(map2[(size_t)y * map_stride + x] * 32.0f);
            const int ix = sat(sx >> 5), iy = sat(sy >> 5);
            fx[(size_t)y * map_width + x] = make_uint2(((uint32_t)ix & 0xffffu) | ((uint32_t)iy << 16),
                                                       (uint32_t)(sx & 31) | ((uint32_t)(sy & 31) << 5));
        }
    CK(cudaMalloc(&h->d_remap, fx.size() * sizeof(uint2)));
    CK(cudaMemcpy(h->d_remap, fx.data(), fx.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    h->map_w = map_width; h->map_h = map_height; h->map_src_w = src_width; h->map_src_h = src_height;
    return ORBX_OK;
}

// host side of cv::initUndistortRectifyMap: iR = (P[:, :3] * R)^-1 with cv::Matx's product (running sum) and closed 3x3 cofactor
// inverse, in plain doubles
static int rectify_args(const double* K9, const double* D, int nD, const double* R9, const double* P, int p_cols, OrbxRectifyArgs* out)
{
    if (!K9 || !R9 || !P || (nD > 0 && !D) || (p_cols != 3 && p_cols != 4)) return fail(ORBX_ERR_INVALID, "K (3x3), R (3x3), P (3x3 or 3x4) required");
    if (nD != 0 && nD != 4 && nD != 5 && nD != 8 && nD != 12) return fail(ORBX_ERR_UNSUPPORTED, "0, 4, 5, 8 or 12 distortion coefficients (no tilt)");
    OrbxRectifyArgs a{};
    for (int i = 0; i < nD; i++) a.k[i] = D[i];
    a.fx = K9[0]; a.fy = K9[4]; a.u0 = K9[2]; a.v0 = K9[5];
    double m[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int q = 0; q < 3; q++) s = s + P[p_cols * r + q] * R9[3 * q + c];
            m[3 * r + c] = s;
        }
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0) return fail(ORBX_ERR_INVALID, "P * R is singular");
    d = 1.0 / d;
    a.ir[0] = (m[4] * m[8] - m[5] * m[7]) * d; a.ir[1] = (m[2] * m[7] - m[1] * m[8]) * d; a.ir[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    a.ir[3] = (m[5] * m[6] - m[3] * m[8]) * d; a.ir[4] = (m[0] * m[8] - m[2] * m[6]) * d; a.ir[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    a.ir[6] = (m[3] * m[7] - m[4] * m[6]) * d; a.ir[7] = (m[1] * m[6] - m[0] * m[7]) * d; a.ir[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    *out = a;
    return ORBX_OK;
}

extern "C" int orbx_init_undistort_rectify_map(const double* K9, const double* D, int nD, const double* R9, const double* P, int p_cols,
                                               int width, int height, float* map1, float* map2, int device)
{
    if (width <= 0 || height <= 0 || !map1 || !map2) return fail(ORBX_ERR_INVALID, "bad argument");
    OrbxRectifyArgs a;
    int rc = rectify_args(K9, D, nD, R9, P, p_cols, &a);
    if (rc != ORBX_OK) return rc;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    float* d = nullptr;
    const size_t n = (size_t)width * height;
    CK(cudaMalloc(&d, 2 * n * sizeof(float)));
    orbx_launch_rectify_map(a, width, height, d, d + n, nullptr, 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(map1, d, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(map2, d + n, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_set_rectify_camera(orbx_extractor* h, const double* K9, const double* D, int nD, const double* R9, const double* P,
                                       int p_cols, int map_width, int map_height, int src_width, int src_height)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (map_width <= 0 || map_height <= 0 || src_width <= 0 || src_height <= 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    if (src_width > 32766 || src_height > 32766) return fail(ORBX_ERR_UNSUPPORTED, "source frame too large for 16-bit map coordinates");
    OrbxRectifyArgs a;
    int rc = rectify_args(K9, D, nD, R9, P, p_cols, &a);
    if (rc != ORBX_OK) return rc;
    rc = orbx_set_rectify_maps(h, nullptr, nullptr, 0, 0, 0, 0, 0);             // drains batches in flight, drops the old map and graph
    if (rc != ORBX_OK) return rc;
    CK(cudaSetDevice(h->device));
    CK(cudaMalloc(&h->d_remap, (size_t)map_width * map_height * sizeof(uint2)));
    orbx_launch_rectify_map(a, map_width, map_height, nullptr, nullptr, h->d_remap, 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    h->map_w = map_width; h->map_h = map_height; h->map_src_w = src_width; h->map_src_h = src_height;
    return ORBX_OK;
}

extern "C" int orbx_rectify_map_size(const orbx_extractor* h, int* map_width, int* map_height)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (!h->d_remap) return fail(ORBX_ERR_STATE, "orbx_set_rectify_maps has not been called");
    if (map_width) *map_width = h->map_w;
    if (map_height) *map_height = h->map_h;
    return ORBX_OK;
}

extern "C" int orbx_extract_batch_rectified(orbx_extractor* h, const uint8_t* const* images, int n, int stride,
                                            OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (!h->d_remap) return fail(ORBX_ERR_STATE, "orbx_set_rectify_maps has not been called");
    return extract_batch_impl(h, images, n, h->map_src_w, h->map_src_h, stride, 1, 0, keypoints, cap, nkp, descriptors, true);
}

extern "C" int orbx_extract(orbx_extractor* h, const uint8_t* image, int width, int height, int stride,
                            OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    if (nkp) *nkp = 0;
    if (!image || width <= 0 || height <= 0) return ORBX_OK;   // ORBextractor.cc:1141-1142
    const uint8_t* imgs[1] = {image};
    return orbx_extract_batch(h, imgs, 1, width, height, stride, keypoints, cap, nkp, descriptors);
}

int orbx_internal_results(orbx_extractor* h, int frame_index, const OrbxKp28** d_kps, const uint8_t** d_desc, cudaStream_t* st, int* device)
{
    if (!h || !h->W || h->last_frames <= 0) return fail(ORBX_ERR_STATE, "no extract has run yet");
    const int rc = finish_all_pending(h);
    if (rc != ORBX_OK) return rc;
    if (!h->map_chunk) return fail(ORBX_ERR_STATE, "the last call was orbx_extract_device: its results live in the caller's buffers (use orbx_frame_from_device)");
    const int wsi = h->ws_index(frame_index);
    if (wsi < 0) return fail(ORBX_ERR_STATE, "frame out of range, or its working-set slot was reused by a later chunk of the same call");
    const int kc = h->L.kp_cap_total;
    *d_kps = h->d_kps + (size_t)wsi * kc; *d_desc = h->d_desc + (size_t)wsi * kc * 32;
    *st = h->slot_stream[(frame_index / h->map_chunk) % h->map_slots]; *device = h->device;
    return ORBX_OK;
}

extern "C" int orbx_level_size(const orbx_extractor* h, int level, int* width, int* height)
{
    if (!h || !h->W) return fail(ORBX_ERR_STATE, "no geometry reserved yet");
    if (level < 0 || level >= h->nlevels) return fail(ORBX_ERR_INVALID, "level out of range");
    if (width) *width = h->lvl[level].w;
    if (height) *height = h->lvl[level].h;
    return ORBX_OK;
}

extern "C" int orbx_pyramid_level_device(orbx_extractor* h, int frame, int level, const uint8_t** d_payload, int* pitch)
{
    if (!h || !h->W || h->last_frames <= 0) return fail(ORBX_ERR_STATE, "no extract has run yet");
    if (level < 0 || level >= h->nlevels || frame < 0 || frame >= h->last_frames) return fail(ORBX_ERR_INVALID, "frame/level out of range");
    const int wsi = h->ws_index(frame);
    if (wsi < 0) return fail(ORBX_ERR_STATE, "the pyramid of this frame has been overwritten by a later chunk of the same call");
    const OrbxLevelGeom& g = h->lvl[level];
    if (d_payload) *d_payload = h->L.raw + (size_t)wsi * h->L.frame_raw_bytes + g.raw_off + (size_t)ORBX_EDGE * g.pitch + ORBX_XOFF;
    if (pitch) *pitch = g.pitch;
    return ORBX_OK;
}

extern "C" int orbx_set_pyramid_mirror(orbx_extractor* h, int on)
{
    if (!h) return fail(ORBX_ERR_INVALID, "handle is NULL");
    h->mirror_on = on != 0;
    return ORBX_OK;
}

extern "C" int orbx_pyramid_level_layout(const orbx_extractor* h, int level, size_t* payload_offset, int* pitch, size_t* block_bytes)
{
    if (!h || !h->W) return fail(ORBX_ERR_STATE, "no geometry reserved yet");
    if (level < 0 || level >= h->nlevels) return fail(ORBX_ERR_INVALID, "level out of range");
    const OrbxLevelGeom& g = h->lvl[level];
    if (payload_offset) *payload_offset = (size_t)g.raw_off + (size_t)ORBX_EDGE * g.pitch + ORBX_XOFF;
    if (pitch) *pitch = g.pitch;
    if (block_bytes) *block_bytes = h->L.frame_raw_bytes;
    return ORBX_OK;
}

extern "C" int orbx_pyramid_mirror(orbx_extractor* h, int frame, const uint8_t** host_block)
{
    if (!h || !host_block) return fail(ORBX_ERR_INVALID, "NULL argument");
    *host_block = nullptr;
    int rc = finish_all_pending(h);
    if (rc != ORBX_OK) return rc;
    if (!h->W || h->last_frames <= 0) return fail(ORBX_ERR_STATE, "no extract has run yet");
    if (frame < 0 || frame >= h->last_frames) return fail(ORBX_ERR_INVALID, "frame out of range");
    if (h->mirror && h->mirror_frame == frame) { *host_block = h->mirror; return ORBX_OK; }   // downloaded by the extract call itself
    const int wsi = h->ws_index(frame);
    if (wsi < 0) return fail(ORBX_ERR_STATE, "the pyramid of this frame has been overwritten by a later chunk of the same call");
    CK(cudaSetDevice(h->device));
    if (!h->mirror) { CK(cudaMallocHost(&h->mirror, h->L.frame_raw_bytes)); h->mirror_bytes = h->L.frame_raw_bytes; }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyAsync(h->mirror, h->L.raw + (size_t)wsi * h->L.frame_raw_bytes, h->L.frame_raw_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->mirror_frame = frame;
    *host_block = h->mirror;
    return ORBX_OK;
}

extern "C" int orbx_pyramid_level(orbx_extractor* h, int frame, int level, uint8_t* dst, int dst_stride)
{
    const uint8_t* pay; int pitch;
    int rc = finish_all_pending(h);                      // the pyramid of a batch still in flight is not there yet
    if (rc != ORBX_OK) return rc;
    rc = orbx_pyramid_level_device(h, frame, level, &pay, &pitch);
    if (rc != ORBX_OK) return rc;
    const OrbxLevelGeom& g = h->lvl[level];
    if (!dst || dst_stride < g.w + 2 * ORBX_EDGE) return fail(ORBX_ERR_INVALID, "dst too small");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy2D(dst, dst_stride, pay - (size_t)ORBX_EDGE * pitch - ORBX_EDGE, pitch, g.w + 2 * ORBX_EDGE,
                    g.h + 2 * ORBX_EDGE, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" int orbx_debug_blurred_level(orbx_extractor* h, int frame, int level, uint8_t* dst, int dst_stride)
{
    const uint8_t* pay; int pitch;
    int rc = finish_all_pending(h);
    if (rc != ORBX_OK) return rc;
    rc = orbx_pyramid_level_device(h, frame, level, &pay, &pitch);
    if (rc != ORBX_OK) return rc;
    const OrbxLevelGeom& g = h->lvl[level];
    if (!dst || dst_stride < g.w) return fail(ORBX_ERR_INVALID, "dst too small");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy2D(dst, dst_stride, h->L.blur + (pay - h->L.raw), pitch, g.w, g.h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" int orbx_debug_level_counts(orbx_extractor* h, int frame, int32_t* counts)
{
    if (!h || !h->W || h->last_frames <= 0) return fail(ORBX_ERR_STATE, "no extract has run yet");
    if (frame < 0 || frame >= h->last_frames || !counts) return fail(ORBX_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->ws_index(frame) < 0) return fail(ORBX_ERR_STATE, "frame overwritten by a later chunk");
    CK(cudaMemcpy(counts, h->L.lvl_kp_count + (size_t)h->ws_index(frame) * h->nlevels, h->nlevels * sizeof(int), cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" int orbx_debug_candidates(orbx_extractor* h, int frame, int level, OrbxKeyPoint* out, int cap, int* n)
{
    if (!h || !h->W || h->last_frames <= 0) return fail(ORBX_ERR_STATE, "no extract has run yet");
    if (level < 0 || level >= h->nlevels || frame < 0 || frame >= h->last_frames || !n) return fail(ORBX_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    int cnt = 0;
    if (h->ws_index(frame) < 0) return fail(ORBX_ERR_STATE, "frame overwritten by a later chunk");
    CK(cudaMemcpy(&cnt, h->L.cand_count + (size_t)h->ws_index(frame) * h->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    const int m = std::min(cnt, cap);
    if (m > 0 && out) {
        std::vector<uint32_t> tmp(m);
        CK(cudaMemcpy(tmp.data(), h->L.cand + (size_t)h->ws_index(frame) * h->L.cand_total + h->lvl[level].cand_off, (size_t)m * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; i++) {
            OrbxKeyPoint k;
            k.x = (float)(tmp[i] & 0xfff); k.y = (float)((tmp[i] >> 12) & 0xfff);
            k.size = 7.f; k.angle = -1.f; k.response = (float)(tmp[i] >> 24); k.octave = 0; k.class_id = -1;
            out[i] = k;
        }
    }
    return ORBX_OK;
}

// ------------------------------------------------------------------------------------------------ Hamming
extern "C" int orbx_hamming_init_device(uint64_t* d_packed, int nq, void* cuda_stream)
{
    if (!d_packed || nq < 0) return fail(ORBX_ERR_INVALID, "bad argument");
    orbx_launch_hamming_init(d_packed, nq, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}
extern "C" int orbx_hamming_top2_device(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, int64_t index_base,
                                        uint64_t* d_packed, void* cuda_stream)
{
    if (nq < 0 || nt < 0 || (nq > 0 && (!d_query || !d_packed)) || (nt > 0 && !d_train)) return fail(ORBX_ERR_INVALID, "bad argument");
    if (((uintptr_t)d_query | (uintptr_t)d_train) & 15) return fail(ORBX_ERR_INVALID, "descriptor arrays must be 16-byte aligned on the device");
    orbx_launch_hamming_top2(d_query, nq, d_train, nt, index_base, d_packed, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}
extern "C" int orbx_hamming_merge_device(const uint64_t* d_parts, int nparts, int nq, int32_t* d_idx1, int32_t* d_dist1,
                                         int32_t* d_dist2, void* cuda_stream)
{
    if (!d_parts || nparts < 1 || nq < 0) return fail(ORBX_ERR_INVALID, "bad argument");
    orbx_launch_hamming_merge(d_parts, nparts, nq, d_idx1, d_dist1, d_dist2, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_hamming_top2(const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* idx1, int32_t* dist1,
                                 int32_t* dist2, int device)
{
    if (nq < 0 || nt < 0 || (nq > 0 && (!query || !idx1 || !dist1 || !dist2)) || (nt > 0 && !train))
        return fail(ORBX_ERR_INVALID, "bad argument");
    if (nq == 0) return ORBX_OK;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    uint8_t *dq = nullptr, *dt = nullptr; uint64_t* dp = nullptr; int* dout = nullptr;
    int rc = ORBX_OK;
    cudaError_t e;
    do {
        if ((e = cudaMalloc(&dq, (size_t)nq * 32)) != cudaSuccess) break;
        if ((e = cudaMalloc(&dt, std::max<size_t>((size_t)nt * 32, 32))) != cudaSuccess) break;
        if ((e = cudaMalloc(&dp, (size_t)nq * 8)) != cudaSuccess) break;
        if ((e = cudaMalloc(&dout, (size_t)nq * 12)) != cudaSuccess) break;
        if ((e = cudaMemcpy(dq, query, (size_t)nq * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (nt > 0 && (e = cudaMemcpy(dt, train, (size_t)nt * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        orbx_launch_hamming_init(dp, nq, 0);
        orbx_launch_hamming_top2(dq, nq, dt, nt, 0, dp, 0);
        orbx_launch_hamming_merge(dp, 1, nq, dout, dout + nq, dout + 2 * (size_t)nq, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpy(idx1, dout, (size_t)nq * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(dist1, dout + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(dist2, dout + 2 * (size_t)nq, (size_t)nq * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) rc = fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    cudaFree(dq); cudaFree(dt); cudaFree(dp); cudaFree(dout);
    return rc;
}

extern "C" int orbx_stereo_hamming(const OrbxKeyPoint* kl, const uint8_t* dl, int nl, const OrbxKeyPoint* kr,
                                   const uint8_t* dr, int nr, int rows, const float* scale_factors, int nlevels,
                                   float minD, float maxD, int32_t* best_idx_r, int32_t* best_dist, int device)
{
    if (nl < 0 || nr < 0 || rows <= 0 || !scale_factors || (nl > 0 && (!kl || !dl || !best_idx_r || !best_dist)) || (nr > 0 && (!kr || !dr)))
        return fail(ORBX_ERR_INVALID, "bad argument");
    if (nl == 0) return ORBX_OK;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    // vRowIndices (Frame.cc:564-590) as a CSR table; filling in ascending iR keeps every row's push_back order.
    // This is O(nr * band) host bookkeeping on a few thousand keypoints, exactly as in the reference.
    std::vector<int> start(rows + 1, 0);
    std::vector<int> lo(nr), hi(nr);
    for (int i = 0; i < nr; i++) {
        if (kr[i].octave < 0 || kr[i].octave >= nlevels) return fail(ORBX_ERR_INVALID, "right keypoint octave out of range");
        const float r = 2.0f * scale_factors[kr[i].octave];
        hi[i] = std::min((int)ceilf(kr[i].y + r), rows - 1);
        lo[i] = std::max((int)floorf(kr[i].y - r), 0);
        for (int y = lo[i]; y <= hi[i]; y++) start[y + 1]++;
    }
    for (int y = 0; y < rows; y++) start[y + 1] += start[y];
    std::vector<int> tab(std::max(start[rows], 1)), fill(rows, 0);
    for (int i = 0; i < nr; i++)
        for (int y = lo[i]; y <= hi[i]; y++) tab[start[y] + fill[y]++] = i;
    CK(cudaSetDevice(device));
    void* pool = nullptr;
    const size_t b_kl = (size_t)nl * 28, b_dl = (size_t)nl * 32, b_kr = (size_t)std::max(nr, 1) * 28, b_dr = (size_t)std::max(nr, 1) * 32;
    const size_t b_st = (size_t)(rows + 1) * 4, b_tab = tab.size() * 4, b_out = (size_t)nl * 8;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t total = al(b_kl) + al(b_dl) + al(b_kr) + al(b_dr) + al(b_st) + al(b_tab) + al(b_out);
    CK(cudaMalloc(&pool, total));
    uint8_t* p = (uint8_t*)pool;
    uint8_t *p_kl = p; p += al(b_kl);
    uint8_t *p_dl = p; p += al(b_dl);
    uint8_t *p_kr = p; p += al(b_kr);
    uint8_t *p_dr = p; p += al(b_dr);
    uint8_t *p_st = p; p += al(b_st);
    uint8_t *p_tab = p; p += al(b_tab);
    uint8_t *p_out = p;
    int rc = ORBX_OK;
    cudaError_t e;
    do {
        if ((e = cudaMemcpy(p_kl, kl, b_kl, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_dl, dl, b_dl, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (nr > 0 && (e = cudaMemcpy(p_kr, kr, (size_t)nr * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (nr > 0 && (e = cudaMemcpy(p_dr, dr, (size_t)nr * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_st, start.data(), b_st, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_tab, tab.data(), b_tab, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        orbx_launch_stereo_hamming((const OrbxKp28*)p_kl, p_dl, nl, (const OrbxKp28*)p_kr, p_dr, nr, (const int*)p_st,
                                   (const int*)p_tab, rows, minD, maxD, (int*)p_out, (int*)p_out + nl, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpy(best_idx_r, p_out, (size_t)nl * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(best_dist, p_out + (size_t)nl * 4, (size_t)nl * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) rc = fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    cudaFree(pool);
    return rc;
}

// ------------------------------------------------------------------------------------------------ full stereo
static int stereo_check_pair(orbx_extractor* left, orbx_extractor* right, int pairs)
{
    if (left->last_frames < pairs || right->last_frames < pairs)
        return fail(ORBX_ERR_STATE, "both extractors must have extracted the pair(s) first");
    if ((left->map_chunk && pairs > left->map_chunk) || (right->map_chunk && pairs > right->map_chunk))
        return fail(ORBX_ERR_STATE, "batched stereo matching needs the pyramids of orbx_extract_device (contiguous frames)");
    if (left->device != right->device || left->W != right->W || left->H != right->H || left->nlevels != right->nlevels ||
        left->scale_factor != right->scale_factor)
        return fail(ORBX_ERR_INVALID, "left and right extractors must share device, image size and pyramid settings");
    return ORBX_OK;
}

extern "C" int orbx_stereo_match_device(orbx_extractor* left, orbx_extractor* right, int pairs,
                                        const OrbxKeyPoint* d_kl, const uint8_t* d_dl, const int32_t* d_nl,
                                        const OrbxKeyPoint* d_kr, const uint8_t* d_dr, const int32_t* d_nr, int cap,
                                        float mbf, float fx, float* d_u_right, float* d_depth, void* cuda_stream)
{
    if (!left || !right || pairs <= 0 || cap <= 0 || !d_kl || !d_dl || !d_nl || !d_kr || !d_dr || !d_nr || !d_u_right || !d_depth)
        return fail(ORBX_ERR_INVALID, "bad argument");
    if (cap > 18000) return fail(ORBX_ERR_UNSUPPORTED, "more than 18000 keypoints per image");
    int rc = finish_all_pending(left);
    if (rc == ORBX_OK) rc = finish_all_pending(right);
    if (rc != ORBX_OK) return rc;
    rc = stereo_check_pair(left, right, pairs);
    if (rc != ORBX_OK) return rc;
    CK(cudaSetDevice(left->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : left->stream;
    const size_t need = (size_t)pairs * cap * sizeof(int);
    if (need > left->stereo_scratch_bytes) {
        CK(cudaStreamSynchronize(st));
        cudaFree(left->stereo_scratch); left->stereo_scratch = nullptr; left->stereo_scratch_bytes = 0;
        CK(cudaMalloc(&left->stereo_scratch, need));
        left->stereo_scratch_bytes = need;
    }
    const float mb = mbf / fx;                      // Frame.cc:120
    OrbxStereoBatch a;
    a.kl = (const OrbxKp28*)d_kl; a.dl = d_dl; a.nl = d_nl; a.kr = (const OrbxKp28*)d_kr; a.dr = d_dr; a.nr = d_nr;
    a.cap = cap; a.pairs = pairs; a.rows = left->lvl[0].h;
    a.raw_left = left->L.raw + (size_t)left->pyr_base * left->L.frame_raw_bytes;
    a.raw_right = right->L.raw + (size_t)right->pyr_base * right->L.frame_raw_bytes;
    a.frame_raw_bytes = left->L.frame_raw_bytes;
    a.lvl = left->d_lvl;
    a.minD = 0.f; a.maxD = mbf / mb; a.mbf = mbf; a.max_scale = left->sf.back();  // Frame.cc:592-595
    a.u_right = d_u_right; a.depth = d_depth; a.sad = (int*)left->stereo_scratch;
    orbx_launch_stereo_batch(a, st);
    CK(cudaGetLastError());
    return ORBX_OK;
}

// Batched host form of Frame's stereo constructor hot part (Frame.cc:80-117): left + right ExtractORB and
// ComputeStereoMatches for n pairs, chunks pipelined over the slot streams exactly like orbx_extract_batch (H2D of later
// chunks and D2H of earlier ones overlap the kernels). Both extractors run on the left extractor's slot stream so that
// the matcher can follow them without an event.
static int stereo_extract_batch_impl(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                                     const uint8_t* const* images_right, int n, int width, int height, int stride,
                                     float mbf, float fx, OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left,
                                     OrbxKeyPoint* kp_right, uint8_t* desc_right, int32_t* n_right, int cap,
                                     float* u_right, float* depth, bool begin_only, bool rectify = false)
{
    if (!left || !right || left == right) return fail(ORBX_ERR_INVALID, "two extractor instances required (Tracking.cc:120-123)");
    if (n <= 0 || !images_left || !images_right || width <= 0 || height <= 0) return ORBX_OK;
    // with rectification (stereo_euroc.cc:136-137 fused into level 0) the frames have the maps' SOURCE size and the pyramids
    // the maps' own size; both cameras must rectify onto the same image size
    if (rectify && (!left->d_remap || !right->d_remap || left->map_w != right->map_w || left->map_h != right->map_h ||
                    left->map_src_w != right->map_src_w || left->map_src_h != right->map_src_h))
        return fail(ORBX_ERR_STATE, "both extractors need rectification maps of the same size (orbx_set_rectify_maps / orbx_set_rectify_camera)");
    const int gw = rectify ? left->map_w : width, gh = rectify ? left->map_h : height;
    // the right handle may still serve another left handle's batch, or hold a mono batch of its own
    if (right->stereo_owner && right->stereo_owner != left) { const int rc = finish_all_pending(right); if (rc != ORBX_OK) return rc; }
    while (right->npending > 0) { const int rc = finish_oldest_pending(right); if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc; }
    // batches in flight live in `left`: a blocking call, another shape or a third batch first completes them
    while (left->npending > 0 && (!begin_only || left->npending >= 2 || left->pending[0].n != n || left->pending[0].cap != cap ||
                                  left->pending[0].fbytes != (size_t)width * height || gw != left->W || gh != left->H)) {
        const int rc = finish_oldest_pending(left);
        if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc;
    }
    if (!kp_left || !desc_left || !n_left || !kp_right || !desc_right || !n_right || !u_right || !depth || stride < width)
        return fail(ORBX_ERR_INVALID, "bad output buffers");
    if (left->device != right->device || left->nlevels != right->nlevels || left->scale_factor != right->scale_factor)
        return fail(ORBX_ERR_INVALID, "left and right extractors must share device and pyramid settings");
    for (orbx_extractor* h : {left, right})
        if (gw != h->W || gh != h->H || h->max_batch < 1) {
            int rc = orbx_reserve(h, gw, gh, std::max(1, std::min(n, std::max(h->max_batch, 64))));
            if (rc != ORBX_OK) return rc;
        }
    if (left->max_batch != right->max_batch) {
        const int B2 = std::max(left->max_batch, right->max_batch);
        for (orbx_extractor* h : {left, right}) { int rc = orbx_reserve(h, gw, gh, B2); if (rc != ORBX_OK) return rc; }
    }
    const int B = left->max_batch, kc = left->L.kp_cap_total;
    if (cap != kc || right->L.kp_cap_total != kc) return fail(ORBX_ERR_INVALID, "cap must equal orbx_max_keypoints() of both extractors");
    if (kc > 18000) return fail(ORBX_ERR_UNSUPPORTED, "more than 18000 keypoints per image");
    CK(cudaSetDevice(left->device));
    const size_t fbytes = (size_t)width * height;
    for (orbx_extractor* h : {left, right})
        if ((size_t)B * fbytes > h->d_in_bytes) {
            CK(cudaDeviceSynchronize());
            cudaFree(h->d_in); h->d_in = nullptr; h->d_in_bytes = 0;
            CK(cudaMalloc(&h->d_in, (size_t)B * fbytes));
            h->d_in_bytes = (size_t)B * fbytes;
        }
    if (left->stereo_out_floats < (size_t)2 * B * kc || left->stereo_scratch_bytes < (size_t)B * kc * sizeof(int)) {
        CK(cudaDeviceSynchronize());
        cudaFree(left->stereo_out); left->stereo_out = nullptr; left->stereo_out_floats = 0;
        cudaFree(left->stereo_scratch); left->stereo_scratch = nullptr; left->stereo_scratch_bytes = 0;
        CK(cudaMalloc(&left->stereo_out, (size_t)2 * B * kc * sizeof(float)));
        left->stereo_out_floats = (size_t)2 * B * kc;
        CK(cudaMalloc(&left->stereo_scratch, (size_t)B * kc * sizeof(int)));
        left->stereo_scratch_bytes = (size_t)B * kc * sizeof(int);
    }
    int chunk = B >= 8 ? std::min(32, std::max(4, B / 8)) : B;            // a chunk carries two frames per pair
    int nslots = B >= 8 ? std::max(1, std::min(orbx_extractor::MAX_SLOTS, B / chunk)) : 1;
    for (int j = 0; j < nslots; j++)
        if (!left->slot_stream[j]) CK(cudaStreamCreateWithFlags(&left->slot_stream[j], cudaStreamNonBlocking));
    CK(cudaStreamSynchronize(left->stream));
    CK(cudaStreamSynchronize(right->stream));
    const float mb = mbf / fx;
    int k = 0;
    for (int f0 = 0; f0 < n; f0 += chunk, k++) {
        const int m = std::min(chunk, n - f0);
        const int slot = k % nslots, base = slot * chunk;
        cudaStream_t st = left->slot_stream[slot];
        struct Side { orbx_extractor* h; const uint8_t* const* imgs; OrbxKeyPoint* kp; uint8_t* desc; int32_t* nk; };
        const Side sides[2] = {{left, images_left, kp_left, desc_left, n_left}, {right, images_right, kp_right, desc_right, n_right}};
        for (const Side& sd : sides) {
            uint8_t* d_in = sd.h->d_in + (size_t)base * fbytes;
            bool contiguous = stride == width;
            for (int i = 0; i < m && contiguous; i++) {
                if (!sd.imgs[f0 + i]) return fail(ORBX_ERR_INVALID, "images[i] is NULL");
                contiguous = sd.imgs[f0 + i] == sd.imgs[f0] + (size_t)i * fbytes;
            }
            if (contiguous) CK(cudaMemcpyAsync(d_in, sd.imgs[f0], (size_t)m * fbytes, cudaMemcpyHostToDevice, st));
            else
                for (int i = 0; i < m; i++) {
                    if (!sd.imgs[f0 + i]) return fail(ORBX_ERR_INVALID, "images[i] is NULL");
                    CK(cudaMemcpy2DAsync(d_in + (size_t)i * fbytes, width, sd.imgs[f0 + i], stride, width, height, cudaMemcpyHostToDevice, st));
                }
            int rc = run_pipeline(sd.h, d_in, m, width, fbytes, sd.h->d_kps + (size_t)base * kc, sd.h->d_desc + (size_t)base * kc * 32, kc,
                                  sd.h->d_nkp + base, st, base, 1, 0, rectify, false);
            if (rc != ORBX_OK) return rc;
        }
        OrbxStereoBatch a;
        a.kl = left->d_kps + (size_t)base * kc; a.dl = left->d_desc + (size_t)base * kc * 32; a.nl = left->d_nkp + base;
        a.kr = right->d_kps + (size_t)base * kc; a.dr = right->d_desc + (size_t)base * kc * 32; a.nr = right->d_nkp + base;
        a.cap = kc; a.pairs = m; a.rows = left->lvl[0].h;
        a.raw_left = left->L.raw + (size_t)base * left->L.frame_raw_bytes;
        a.raw_right = right->L.raw + (size_t)base * right->L.frame_raw_bytes;
        a.frame_raw_bytes = left->L.frame_raw_bytes;
        a.lvl = left->d_lvl;
        a.minD = 0.f; a.maxD = mbf / mb; a.mbf = mbf; a.max_scale = left->sf.back();
        float* d_ur = left->stereo_out + (size_t)base * kc;
        float* d_dp = left->stereo_out + (size_t)B * kc + (size_t)base * kc;
        a.u_right = d_ur; a.depth = d_dp; a.sad = (int*)left->stereo_scratch + (size_t)base * kc;
        orbx_launch_stereo_batch(a, st);
        CK(cudaGetLastError());
        for (const Side& sd : sides) {
            CK(cudaMemcpyAsync(sd.nk + f0, sd.h->d_nkp + base, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(sd.kp + (size_t)f0 * kc, sd.h->d_kps + (size_t)base * kc, (size_t)m * kc * sizeof(OrbxKp28), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(sd.desc + (size_t)f0 * kc * 32, sd.h->d_desc + (size_t)base * kc * 32, (size_t)m * kc * 32, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaMemcpyAsync(u_right + (size_t)f0 * kc, d_ur, (size_t)m * kc * sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(depth + (size_t)f0 * kc, d_dp, (size_t)m * kc * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (orbx_extractor* h : {left, right}) { h->last_frames = n; h->pyr_base = 0; h->map_chunk = chunk; h->map_slots = nslots; }
    if (begin_only) {
        orbx_extractor::Pending& pd = left->pending[left->npending];
        pd.n = n; pd.cap = cap; pd.nslots = nslots; pd.fbytes = fbytes; pd.nkp = n_left; pd.nkp2 = n_right;
        for (int j = 0; j < nslots; j++) {
            if (!pd.done[j]) CK(cudaEventCreateWithFlags(&pd.done[j], cudaEventDisableTiming));
            CK(cudaEventRecord(pd.done[j], left->slot_stream[j]));
        }
        left->npending++;
        left->stereo_partner = right; right->stereo_owner = left;
        return ORBX_OK;
    }
    for (int j = 0; j < nslots; j++) CK(cudaStreamSynchronize(left->slot_stream[j]));
    for (int i = 0; i < n; i++) if (n_left[i] > kc || n_right[i] > kc) return fail(ORBX_ERR_CAPACITY, "keypoint buffer too small");
    return ORBX_OK;
}

extern "C" int orbx_stereo_extract_batch(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                                         const uint8_t* const* images_right, int n, int width, int height, int stride,
                                         float mbf, float fx, OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left,
                                         OrbxKeyPoint* kp_right, uint8_t* desc_right, int32_t* n_right, int cap,
                                         float* u_right, float* depth)
{
    return stereo_extract_batch_impl(left, right, images_left, images_right, n, width, height, stride, mbf, fx, kp_left, desc_left, n_left,
                                     kp_right, desc_right, n_right, cap, u_right, depth, false);
}

extern "C" int orbx_stereo_extract_batch_begin(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                                               const uint8_t* const* images_right, int n, int width, int height, int stride,
                                               float mbf, float fx, OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left,
                                               OrbxKeyPoint* kp_right, uint8_t* desc_right, int32_t* n_right, int cap,
                                               float* u_right, float* depth)
{
    return stereo_extract_batch_impl(left, right, images_left, images_right, n, width, height, stride, mbf, fx, kp_left, desc_left, n_left,
                                     kp_right, desc_right, n_right, cap, u_right, depth, true);
}

// the same with the rectification of both cameras fused into level 0: UNRECTIFIED frames in (the maps' source size)
extern "C" int orbx_stereo_extract_batch_rectified(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                                                   const uint8_t* const* images_right, int n, int stride, float mbf, float fx,
                                                   OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left, OrbxKeyPoint* kp_right,
                                                   uint8_t* desc_right, int32_t* n_right, int cap, float* u_right, float* depth)
{
    if (!left || !left->d_remap) return fail(ORBX_ERR_STATE, "orbx_set_rectify_maps / orbx_set_rectify_camera has not been called on the left extractor");
    return stereo_extract_batch_impl(left, right, images_left, images_right, n, left->map_src_w, left->map_src_h, stride, mbf, fx, kp_left,
                                     desc_left, n_left, kp_right, desc_right, n_right, cap, u_right, depth, false, true);
}

extern "C" int orbx_stereo_extract_batch_end(orbx_extractor* left)
{
    return orbx_extract_batch_end(left);
}

extern "C" int orbx_stereo_match(orbx_extractor* left, orbx_extractor* right, const OrbxKeyPoint* kl, const uint8_t* dl,
                                 int nl, const OrbxKeyPoint* kr, const uint8_t* dr, int nr, float mbf, float fx,
                                 float* u_right, float* depth)
{
    if (!left || !right || nl < 0 || nr < 0 || (nl > 0 && (!kl || !dl || !u_right || !depth)) || (nr > 0 && (!kr || !dr)))
        return fail(ORBX_ERR_INVALID, "bad argument");
    for (int i = 0; i < nl; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    if (nl == 0) return ORBX_OK;
    int rc = stereo_check_pair(left, right, 1);
    if (rc != ORBX_OK) return rc;
    for (int i = 0; i < nl; i++) if (kl[i].octave < 0 || kl[i].octave >= left->nlevels) return fail(ORBX_ERR_INVALID, "left keypoint octave out of range");
    for (int i = 0; i < nr; i++) if (kr[i].octave < 0 || kr[i].octave >= left->nlevels) return fail(ORBX_ERR_INVALID, "right keypoint octave out of range");
    CK(cudaSetDevice(left->device));
    CK(cudaStreamSynchronize(left->stream));
    CK(cudaStreamSynchronize(right->stream));
    const int cap = std::max(std::max(nl, nr), 1);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_k = al((size_t)cap * 28), b_d = al((size_t)cap * 32), b_f = al((size_t)cap * 4);
    uint8_t* pool = nullptr;
    CK(cudaMalloc(&pool, 2 * b_k + 2 * b_d + 2 * b_f + 256));
    uint8_t *p_kl = pool, *p_kr = pool + b_k, *p_dl = pool + 2 * b_k, *p_dr = pool + 2 * b_k + b_d;
    uint8_t *p_u = pool + 2 * b_k + 2 * b_d, *p_dp = p_u + b_f, *p_n = p_dp + b_f;
    const int counts[2] = {nl, nr};
    int out = ORBX_OK;
    cudaError_t e;
    do {
        if ((e = cudaMemcpy(p_kl, kl, (size_t)nl * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_dl, dl, (size_t)nl * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (nr > 0 && (e = cudaMemcpy(p_kr, kr, (size_t)nr * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (nr > 0 && (e = cudaMemcpy(p_dr, dr, (size_t)nr * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_n, counts, sizeof(counts), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        out = orbx_stereo_match_device(left, right, 1, (const OrbxKeyPoint*)p_kl, p_dl, (const int32_t*)p_n,
                                       (const OrbxKeyPoint*)p_kr, p_dr, (const int32_t*)p_n + 1, cap, mbf, fx,
                                       (float*)p_u, (float*)p_dp, left->stream);
        if (out != ORBX_OK) break;
        if ((e = cudaStreamSynchronize(left->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpy(u_right, p_u, (size_t)nl * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(depth, p_dp, (size_t)nl * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    cudaFree(pool);
    if (out != ORBX_OK) return out;
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

// ------------------------------------------------------------------------------------------------ fused peer exchange
struct orbx_peer_matcher {
    int nq_max = 0, world = 0, rank = 0, device = 0;
    uint8_t* local = nullptr;                      // [2 epochs][world][nq_max] u64 landing buffer, then the arrival counter
    size_t parts_bytes = 0;                        // bytes of ONE epoch's landing buffer
    uint8_t* peer_base[ORBX_MAX_PEERS] = {};       // mapped bases of every rank's `local` (own entry = local)
    bool opened[ORBX_MAX_PEERS] = {};
    uint64_t* d_packed = nullptr;                  // local slice-merge target
    unsigned* d_tile_done = nullptr;
    unsigned long long epoch = 0, expected = 0;    // calls so far; arrivals expected so far (world x query tiles per call)
    bool connected = false;
};

extern "C" int orbx_peer_create(int nq_max, int world, int rank, int device, orbx_peer_matcher** out, uint8_t* handle_out)
{
    if (!out || !handle_out || nq_max <= 0 || world < 1 || world > ORBX_MAX_PEERS || rank < 0 || rank >= world)
        return fail(ORBX_ERR_INVALID, "bad argument");
    *out = nullptr;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_peer_matcher* m = new orbx_peer_matcher();
    m->nq_max = nq_max; m->world = world; m->rank = rank; m->device = device;
    m->parts_bytes = ((size_t)world * nq_max * 8 + 255) & ~(size_t)255;
    const size_t total = 2 * m->parts_bytes + 256;
    cudaError_t e = cudaMalloc(&m->local, total);           // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    if (e == cudaSuccess) e = cudaMemset(m->local, 0, total);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_packed, (size_t)nq_max * 8);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_tile_done, 4096);
    if (e == cudaSuccess) e = cudaMemset(m->d_tile_done, 0, 4096);
    cudaIpcMemHandle_t hnd;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hnd, m->local);
    if (e != cudaSuccess) { orbx_peer_destroy(m); return fail(ORBX_ERR_CUDA, cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == ORBX_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &hnd, ORBX_IPC_HANDLE_BYTES);
    CK(cudaDeviceSynchronize());
    *out = m;
    return ORBX_OK;
}

extern "C" int orbx_peer_connect(orbx_peer_matcher* m, const uint8_t* all_handles)
{
    if (!m || !all_handles) return fail(ORBX_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(m->device));
    for (int p = 0; p < m->world; p++) {
        if (p == m->rank) { m->peer_base[p] = m->local; continue; }
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, all_handles + (size_t)p * ORBX_IPC_HANDLE_BYTES, ORBX_IPC_HANDLE_BYTES);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
        m->peer_base[p] = (uint8_t*)ptr; m->opened[p] = true;
    }
    m->connected = true;
    return ORBX_OK;
}

extern "C" int orbx_peer_hamming_top2(orbx_peer_matcher* m, const uint8_t* d_query, int nq, const uint8_t* d_train, int nt,
                                      int64_t index_base, int32_t* d_idx1, int32_t* d_dist1, int32_t* d_dist2,
                                      int32_t* d_status, void* cuda_stream)
{
    if (!m || !m->connected) return fail(ORBX_ERR_STATE, "orbx_peer_connect has not run");
    if (nq <= 0 || nq > m->nq_max || nt < 0 || !d_query || (nt > 0 && !d_train)) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_hamming_qtiles(nq) > 1024) return fail(ORBX_ERR_UNSUPPORTED, "too many query tiles");
    CK(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int slot = (int)(m->epoch & 1);          // landing buffers alternate per call
    OrbxHtPeer peer{};
    peer.world = m->world; peer.rank = m->rank; peer.nq_max = m->nq_max;
    for (int p = 0; p < m->world; p++) {
        peer.parts[p] = (unsigned long long*)(m->peer_base[p] + (size_t)slot * m->parts_bytes);
        peer.arrive[p] = (unsigned*)(m->peer_base[p] + 2 * m->parts_bytes);
    }
    peer.tile_done = m->d_tile_done;
    orbx_launch_hamming_init(m->d_packed, nq, st);
    orbx_launch_hamming_top2_peer(d_query, nq, d_train, nt, index_base, m->d_packed, peer, st);
    m->epoch++;
    m->expected += (unsigned long long)m->world * (unsigned long long)orbx_hamming_qtiles(nq);
    const unsigned target = (unsigned)m->expected;   // the counter is compared modulo 2^32
    orbx_launch_hamming_wait_merge((const uint64_t*)(m->local + (size_t)slot * m->parts_bytes),
                                   (const unsigned*)(m->local + 2 * m->parts_bytes), target, m->world, nq, m->nq_max,
                                   d_idx1, d_dist1, d_dist2, d_status, st);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" void orbx_peer_destroy(orbx_peer_matcher* m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < m->world; p++) if (m->opened[p]) cudaIpcCloseMemHandle(m->peer_base[p]);
    cudaFree(m->local); cudaFree(m->d_packed); cudaFree(m->d_tile_done);
    cudaGetLastError();
    delete m;
}

// ------------------------------------------------------------------------------------------------ windowed top-2
extern "C" int orbx_window_top2(const OrbxKeyPoint* kps, const uint8_t* desc, int n, const uint8_t* occupied, const float* u_right,
                                float minX, float minY, float invW, float invH, const OrbxWindowQuery* q, const uint8_t* qdesc,
                                int nq, int32_t* best_idx, int32_t* best_dist, int32_t* best_level, int32_t* best_dist2,
                                int32_t* best_level2, int device)
{
    if (n < 0 || nq < 0 || (n > 0 && (!kps || !desc)) || (nq > 0 && (!q || !qdesc || !best_idx || !best_dist || !best_level || !best_dist2 || !best_level2)))
        return fail(ORBX_ERR_INVALID, "bad argument");
    if (nq == 0) return ORBX_OK;
    if (n > 14000) return fail(ORBX_ERR_UNSUPPORTED, "more than 14000 keypoints per frame");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t nn = (size_t)std::max(n, 1);
    const size_t b_k = al(nn * 28), b_d = al(nn * 32), b_o = al(nn), b_u = al(nn * 4), b_q = al((size_t)nq * sizeof(OrbxWindowQuery)),
                 b_qd = al((size_t)nq * 32), b_out = al((size_t)nq * 4);
    uint8_t* pool = nullptr;
    CK(cudaMalloc(&pool, b_k + b_d + b_o + b_u + b_q + b_qd + 5 * b_out));
    uint8_t* p = pool;
    uint8_t *p_k = p; p += b_k;
    uint8_t *p_d = p; p += b_d;
    uint8_t *p_o = p; p += b_o;
    uint8_t *p_u = p; p += b_u;
    uint8_t *p_q = p; p += b_q;
    uint8_t *p_qd = p; p += b_qd;
    uint8_t *p_out = p;
    int32_t* outs[5] = {best_idx, best_dist, best_level, best_dist2, best_level2};
    cudaError_t e;
    do {
        if (n > 0 && (e = cudaMemcpy(p_k, kps, (size_t)n * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (n > 0 && (e = cudaMemcpy(p_d, desc, (size_t)n * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (n > 0 && occupied && (e = cudaMemcpy(p_o, occupied, (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (n > 0 && u_right && (e = cudaMemcpy(p_u, u_right, (size_t)n * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_q, q, (size_t)nq * sizeof(OrbxWindowQuery), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_qd, qdesc, (size_t)nq * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        OrbxWindowArgs a;
        a.kps = (const OrbxKp28*)p_k; a.desc = p_d; a.n = n;
        a.occupied = (occupied && n > 0) ? p_o : nullptr; a.u_right = (u_right && n > 0) ? (const float*)p_u : nullptr;
        a.minX = minX; a.minY = minY; a.invW = invW; a.invH = invH;
        static_assert(sizeof(OrbxWinQuery) == sizeof(OrbxWindowQuery), "query layout");
        a.q = (const OrbxWinQuery*)p_q; a.qdesc = p_qd; a.nq = nq;
        a.best_idx = (int*)p_out; a.best_dist = (int*)(p_out + b_out); a.best_level = (int*)(p_out + 2 * b_out);
        a.best_dist2 = (int*)(p_out + 3 * b_out); a.best_level2 = (int*)(p_out + 4 * b_out);
        orbx_launch_window_top2(a, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        for (int k = 0; k < 5; k++)
            if ((e = cudaMemcpy(outs[k], p_out + (size_t)k * b_out, (size_t)nq * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    cudaFree(pool);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Frame::UndistortKeyPoints / ComputeImageBounds (Frame.cc:471-538)
static int undistort_args(const float* K4, const float* dist, int ndist, OrbxUndistortArgs* a)
{
    if (!K4 || ndist < 0 || ndist > 5 || (ndist > 0 && !dist)) return fail(ORBX_ERR_INVALID, "K4 = (fx, fy, cx, cy) and up to 5 distortion coefficients required");
    a->fx = K4[0]; a->fy = K4[1]; a->cx = K4[2]; a->cy = K4[3];
    a->ifx = 1.0 / a->fx; a->ify = 1.0 / a->fy;
    for (int i = 0; i < 5; i++) a->k[i] = i < ndist ? (double)dist[i] : 0.0;
    return ORBX_OK;
}

extern "C" int orbx_undistort_keypoints_device(const OrbxKeyPoint* d_in, int n, const float* K4, const float* dist, int ndist,
                                               OrbxKeyPoint* d_out, void* cuda_stream)
{
    if (n < 0 || (n > 0 && (!d_in || !d_out))) return fail(ORBX_ERR_INVALID, "bad argument");
    OrbxUndistortArgs a;
    int rc = undistort_args(K4, dist, ndist, &a);
    if (rc != ORBX_OK) return rc;
    if (n == 0) return ORBX_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (ndist == 0 || dist[0] == 0.0f) {                                       // Frame.cc:474-478: mvKeysUn = mvKeys
        if (d_out != d_in) CK(cudaMemcpyAsync(d_out, d_in, (size_t)n * sizeof(OrbxKp28), cudaMemcpyDeviceToDevice, st));
        return ORBX_OK;
    }
    orbx_launch_undistort((const OrbxKp28*)d_in, (OrbxKp28*)d_out, n, a, st);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_undistort_keypoints(const OrbxKeyPoint* in, int n, const float* K4, const float* dist, int ndist,
                                        OrbxKeyPoint* out, int device)
{
    if (n < 0 || (n > 0 && (!in || !out))) return fail(ORBX_ERR_INVALID, "bad argument");
    OrbxUndistortArgs a;
    int rc = undistort_args(K4, dist, ndist, &a);
    if (rc != ORBX_OK) return rc;
    if (n == 0) return ORBX_OK;
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    OrbxKp28* d = nullptr;
    CK(cudaMalloc(&d, (size_t)n * sizeof(OrbxKp28)));
    cudaError_t e;
    do {
        if ((e = cudaMemcpy(d, in, (size_t)n * sizeof(OrbxKp28), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (ndist > 0 && dist[0] != 0.0f) orbx_launch_undistort(d, d, n, a, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        e = cudaMemcpy(out, d, (size_t)n * sizeof(OrbxKp28), cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_image_bounds(int width, int height, const float* K4, const float* dist, int ndist, float* bounds4, int device)
{
    if (!bounds4 || width <= 0 || height <= 0) return fail(ORBX_ERR_INVALID, "bad argument");
    if (ndist == 0 || !dist || dist[0] == 0.0f) {                              // Frame.cc:530-536
        bounds4[0] = 0.f; bounds4[1] = (float)width; bounds4[2] = 0.f; bounds4[3] = (float)height;
        return ORBX_OK;
    }
    OrbxKeyPoint c[4] = {};
    c[1].x = (float)width; c[2].y = (float)height; c[3].x = (float)width; c[3].y = (float)height;
    int rc = orbx_undistort_keypoints(c, 4, K4, dist, ndist, c, device);
    if (rc != ORBX_OK) return rc;
    bounds4[0] = std::min(c[0].x, c[2].x); bounds4[1] = std::max(c[1].x, c[3].x);    // mnMinX, mnMaxX (Frame.cc:524-527)
    bounds4[2] = std::min(c[0].y, c[1].y); bounds4[3] = std::max(c[2].y, c[3].y);    // mnMinY, mnMaxY
    return ORBX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Bag of words: ORBVocabulary::transform / L1 score / ORBmatcher::SearchByBoW (orbx_bow.cu)
struct orbx_vocabulary {
    int k = 0, L = 0, scoring = 0, weighting = 0, nodes = 0, words = 0, device = 0;
    OrbxVocabDev V{};
    void* d_tree = nullptr;                      // one allocation: slot arrays + node arrays
    // workspace of the last transform
    int frames = 0, cap = 0;
    size_t ws_frames = 0, ws_cap = 0;
    void* d_ws = nullptr;
    OrbxBowOut O{};
    const int* d_n = nullptr;                    // counts of the last transform (caller-owned or d_n_own)
    int* d_n_own = nullptr;
    // match workspace
    void* d_mws = nullptr; size_t mws_pairs = 0, mws_cap = 0;
    int *d_bin_of = nullptr, *d_taken = nullptr, *d_hist = nullptr;
    // The tree is immutable, but the workspace of the last transform (O, d_n, frames / cap) and the match workspace are
    // per-handle mutable state, and ORB-SLAM2 shares ONE ORBVocabulary between Tracking, LocalMapping and LoopClosing:
    // every entry point holds this (recursive) mutex, the host one-shot forms across their whole transform + match
    // sequence, and orbx_vocab_lock / _unlock let a caller hold it across transform -> get (host/FrameOps.h).
    std::recursive_mutex mu;
};
#define VLOCK(v) std::lock_guard<std::recursive_mutex> vlock_((v)->mu)

extern "C" int orbx_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                                 const uint8_t* is_leaf, const uint8_t* descriptors, const double* weights, int device,
                                 orbx_vocabulary** out)
{
    if (!out) return fail(ORBX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_nodes <= 0 || !parent || !is_leaf || !descriptors || !weights) return fail(ORBX_ERR_INVALID, "empty vocabulary");
    // the limits of TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1359)
    if (k < 1 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3)
        return fail(ORBX_ERR_INVALID, "k in [1,20], L in [1,10], scoring in [0,5], weighting in [0,3] required");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    const int n = n_nodes + 1;
    std::vector<int> cnt(n + 1, 0), word(n, -1);
    std::vector<double> wt(n, 0.0);
    int words = 0;
    for (int i = 0; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] > i) return fail(ORBX_ERR_INVALID, "parent id must precede the node (text-file order)");
        cnt[parent[i] + 1]++;
        wt[i + 1] = weights[i];
        if (is_leaf[i]) word[i + 1] = words++;
    }
    for (int i = 0; i < n; i++) {
        if (cnt[i + 1] > 32) return fail(ORBX_ERR_UNSUPPORTED, "more than 32 children per node");
        if (i > 0 && is_leaf[i - 1] && cnt[i + 1]) return fail(ORBX_ERR_INVALID, "a leaf has children");
        if (i > 0 && !is_leaf[i - 1] && !cnt[i + 1]) return fail(ORBX_ERR_INVALID, "an inner node has no children");
        cnt[i + 1] += cnt[i];
    }
    // slot of node id = position in its parent's child list (children in push_back = ascending id order)
    std::vector<int> fill(n, 0), slot_node(n_nodes);
    std::vector<int2> slot_kids(n_nodes);
    std::vector<uint8_t> slot_desc((size_t)n_nodes * 32);
    for (int i = 0; i < n_nodes; i++) {
        const int s = cnt[parent[i]] + fill[parent[i]]++;
        slot_node[s] = i + 1;
        slot_kids[s] = make_int2(cnt[i + 1], cnt[i + 2] - cnt[i + 1]);
        memcpy(&slot_desc[(size_t)s * 32], descriptors + (size_t)i * 32, 32);
    }
    orbx_vocabulary* v = new orbx_vocabulary();
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nodes = n; v->words = words; v->device = device;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_desc = al((size_t)n_nodes * 32), b_kids = al((size_t)n_nodes * sizeof(int2)), b_node = al((size_t)n_nodes * 4),
                 b_wt = al((size_t)n * 8), b_word = al((size_t)n * 4);
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaMalloc(&v->d_tree, b_desc + b_kids + b_node + b_wt + b_word)) != cudaSuccess) {
        delete v; return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    }
    uint8_t* p = (uint8_t*)v->d_tree;
    cudaMemcpy(p, slot_desc.data(), (size_t)n_nodes * 32, cudaMemcpyHostToDevice); v->V.slot_desc = (const uint4*)p; p += b_desc;
    cudaMemcpy(p, slot_kids.data(), (size_t)n_nodes * sizeof(int2), cudaMemcpyHostToDevice); v->V.slot_kids = (const int2*)p; p += b_kids;
    cudaMemcpy(p, slot_node.data(), (size_t)n_nodes * 4, cudaMemcpyHostToDevice); v->V.slot_node = (const int*)p; p += b_node;
    cudaMemcpy(p, wt.data(), (size_t)n * 8, cudaMemcpyHostToDevice); v->V.weight = (const double*)p; p += b_wt;
    cudaMemcpy(p, word.data(), (size_t)n * 4, cudaMemcpyHostToDevice); v->V.word = (const int*)p;
    if ((e = cudaGetLastError()) != cudaSuccess) { cudaFree(v->d_tree); delete v; return fail(ORBX_ERR_CUDA, cudaGetErrorString(e)); }
    v->V.L = L; v->V.scoring = scoring; v->V.weighting = weighting; v->V.root_children = cnt[1] - cnt[0];
    *out = v;
    return ORBX_OK;
}

extern "C" void orbx_vocab_destroy(orbx_vocabulary* v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d_tree); cudaFree(v->d_ws); cudaFree(v->d_mws); cudaFree(v->d_n_own);
    cudaGetLastError();
    delete v;
}
extern "C" int orbx_vocab_lock(orbx_vocabulary* v) { if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL"); v->mu.lock(); return ORBX_OK; }
extern "C" int orbx_vocab_unlock(orbx_vocabulary* v) { if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL"); v->mu.unlock(); return ORBX_OK; }
extern "C" int orbx_vocab_words(const orbx_vocabulary* v) { return v ? v->words : 0; }
extern "C" int orbx_vocab_nodes(const orbx_vocabulary* v) { return v ? v->nodes : 0; }

static int bow_workspace(orbx_vocabulary* v, int frames, int cap)
{
    if ((size_t)frames <= v->ws_frames && (size_t)cap == v->ws_cap) return ORBX_OK;
    CK(cudaDeviceSynchronize());
    cudaFree(v->d_ws); v->d_ws = nullptr; v->ws_frames = v->ws_cap = 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t fc = al((size_t)frames * cap * 4), fc1 = al((size_t)frames * (cap + 1) * 4), fd = al((size_t)frames * cap * 8), f1 = al((size_t)frames * 4);
    CK(cudaMalloc(&v->d_ws, 6 * fc + fc1 + fd + 2 * f1));
    uint8_t* p = (uint8_t*)v->d_ws;
    v->O.leaf = (int*)p; p += fc; v->O.nid = (int*)p; p += fc; v->O.word = (int*)p; p += fc;
    v->O.bow_id = (int*)p; p += fc; v->O.fv_node = (int*)p; p += fc; v->O.fv_feat = (int*)p; p += fc;
    v->O.fv_off = (int*)p; p += fc1; v->O.bow_val = (double*)p; p += fd;
    v->O.n_bow = (int*)p; p += f1; v->O.n_fv = (int*)p;
    v->ws_frames = frames; v->ws_cap = cap;
    return ORBX_OK;
}

extern "C" int orbx_bow_transform_device(orbx_vocabulary* v, const uint8_t* d_descriptors, const int32_t* d_counts, int frames,
                                         int cap, int levelsup, void* cuda_stream)
{
    if (!v || !d_descriptors || !d_counts) return fail(ORBX_ERR_INVALID, "NULL argument");
    VLOCK(v);
    if (frames <= 0 || cap <= 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    if (cap > 16384) return fail(ORBX_ERR_UNSUPPORTED, "more than 16384 features per frame");
    CK(cudaSetDevice(v->device));
    int rc = bow_workspace(v, frames, cap);
    if (rc != ORBX_OK) return rc;
    orbx_launch_bow_transform(v->V, d_descriptors, d_counts, frames, cap, levelsup, v->O, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    v->frames = frames; v->cap = cap; v->d_n = d_counts;
    return ORBX_OK;
}

extern "C" int orbx_bow_get(orbx_vocabulary* v, int frame, int32_t* word, int32_t* node, int32_t* bow_id, double* bow_val,
                            int32_t* n_bow, int32_t* fv_node, int32_t* fv_off, int32_t* fv_feat, int32_t* n_fv)
{
    if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL");
    VLOCK(v);
    if (v->frames <= 0) return fail(ORBX_ERR_STATE, "no transform has run yet");
    if (frame < 0 || frame >= v->frames) return fail(ORBX_ERR_INVALID, "frame out of range");
    CK(cudaSetDevice(v->device));
    CK(cudaDeviceSynchronize());
    const int cap = v->cap;
    int n = 0, nb = 0, nf = 0;
    CK(cudaMemcpy(&n, v->d_n + frame, 4, cudaMemcpyDeviceToHost));
    n = std::min(n, cap);
    CK(cudaMemcpy(&nb, v->O.n_bow + frame, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&nf, v->O.n_fv + frame, 4, cudaMemcpyDeviceToHost));
    const size_t o = (size_t)frame * cap;
    if (word && n) CK(cudaMemcpy(word, v->O.word + o, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (node && n) CK(cudaMemcpy(node, v->O.nid + o, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (bow_id && nb) CK(cudaMemcpy(bow_id, v->O.bow_id + o, (size_t)nb * 4, cudaMemcpyDeviceToHost));
    if (bow_val && nb) CK(cudaMemcpy(bow_val, v->O.bow_val + o, (size_t)nb * 8, cudaMemcpyDeviceToHost));
    if (fv_node && nf) CK(cudaMemcpy(fv_node, v->O.fv_node + o, (size_t)nf * 4, cudaMemcpyDeviceToHost));
    if (fv_off) CK(cudaMemcpy(fv_off, v->O.fv_off + (size_t)frame * (cap + 1), (size_t)(nf + 1) * 4, cudaMemcpyDeviceToHost));
    if (fv_feat && nf) {
        int tot = 0;
        CK(cudaMemcpy(&tot, v->O.fv_off + (size_t)frame * (cap + 1) + nf, 4, cudaMemcpyDeviceToHost));
        if (tot) CK(cudaMemcpy(fv_feat, v->O.fv_feat + o, (size_t)tot * 4, cudaMemcpyDeviceToHost));
    }
    if (n_bow) *n_bow = nb;
    if (n_fv) *n_fv = nf;
    return ORBX_OK;
}

// host convenience: `frames` descriptor sets of up to `cap` rows ([frames][cap][32]) -> transform; results via orbx_bow_get
extern "C" int orbx_bow_transform(orbx_vocabulary* v, const uint8_t* descriptors, const int32_t* counts, int frames, int cap,
                                  int levelsup)
{
    if (!v || !descriptors || !counts || frames <= 0 || cap <= 0) return fail(ORBX_ERR_INVALID, "bad argument");
    VLOCK(v);
    CK(cudaSetDevice(v->device));
    CK(cudaDeviceSynchronize());
    cudaFree(v->d_n_own); v->d_n_own = nullptr;
    uint8_t* d_desc = nullptr;
    CK(cudaMalloc(&v->d_n_own, (size_t)frames * 4));
    CK(cudaMalloc(&d_desc, (size_t)frames * cap * 32));
    cudaError_t e;
    int rc = ORBX_OK;
    do {
        if ((e = cudaMemcpy(v->d_n_own, counts, (size_t)frames * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_desc, descriptors, (size_t)frames * cap * 32, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        rc = orbx_bow_transform_device(v, d_desc, v->d_n_own, frames, cap, levelsup, nullptr);
        e = cudaDeviceSynchronize();
    } while (0);
    cudaFree(d_desc);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

extern "C" int orbx_bow_score_device(orbx_vocabulary* v, const int32_t* d_frame_a, const int32_t* d_frame_b, int npairs,
                                     double* d_score, void* cuda_stream)
{
    if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL");
    VLOCK(v);
    if (v->frames <= 0) return fail(ORBX_ERR_STATE, "no transform has run yet");
    if (npairs < 0 || (npairs > 0 && (!d_frame_a || !d_frame_b || !d_score))) return fail(ORBX_ERR_INVALID, "bad argument");
    if (v->scoring != 0) return fail(ORBX_ERR_UNSUPPORTED, "only L1_NORM scoring (ORBvoc's) is implemented");
    CK(cudaSetDevice(v->device));
    orbx_launch_bow_score(v->O, v->cap, d_frame_a, d_frame_b, npairs, d_score, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_bow_score(orbx_vocabulary* v, const int32_t* frame_a, const int32_t* frame_b, int npairs, double* score)
{
    if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL");
    VLOCK(v);
    if (v->frames <= 0) return fail(ORBX_ERR_STATE, "no transform has run yet");
    if (npairs <= 0) return ORBX_OK;
    if (!frame_a || !frame_b || !score) return fail(ORBX_ERR_INVALID, "bad argument");
    for (int i = 0; i < npairs; i++)
        if (frame_a[i] < 0 || frame_a[i] >= v->frames || frame_b[i] < 0 || frame_b[i] >= v->frames) return fail(ORBX_ERR_INVALID, "frame out of range");
    CK(cudaSetDevice(v->device));
    int* d = nullptr; double* ds = nullptr;
    CK(cudaMalloc(&d, (size_t)npairs * 8));
    cudaError_t e = cudaMalloc(&ds, (size_t)npairs * 8);
    int rc = ORBX_OK;
    if (e == cudaSuccess) do {
        if ((e = cudaMemcpy(d, frame_a, (size_t)npairs * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d + npairs, frame_b, (size_t)npairs * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        rc = orbx_bow_score_device(v, d, d + npairs, npairs, ds, nullptr);
        e = cudaMemcpy(score, ds, (size_t)npairs * 8, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(d); cudaFree(ds);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

static int search_by_bow_device_impl(orbx_vocabulary* v, int npairs, const int32_t* d_kf_frame, const int32_t* d_f_frame,
                                     const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_kf_valid,
                                     const uint8_t* d_f_valid, int kf_mode, float nnratio, int check_orientation,
                                     int32_t* d_match, int32_t* d_nmatches, void* cuda_stream)
{
    if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL");
    VLOCK(v);
    if (v->frames <= 0) return fail(ORBX_ERR_STATE, "no transform has run yet");
    if (npairs <= 0) return ORBX_OK;
    if (!d_kf_frame || !d_f_frame || !d_keypoints || !d_descriptors || !d_match || !d_nmatches) return fail(ORBX_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(v->device));
    const int cap = v->cap;
    if ((size_t)npairs > v->mws_pairs || (size_t)cap != v->mws_cap) {
        CK(cudaDeviceSynchronize());
        cudaFree(v->d_mws); v->d_mws = nullptr;
        CK(cudaMalloc(&v->d_mws, (size_t)npairs * cap * 8 + (size_t)npairs * 32 * 4));
        v->d_bin_of = (int*)v->d_mws; v->d_taken = v->d_bin_of + (size_t)npairs * cap; v->d_hist = v->d_taken + (size_t)npairs * cap;
        v->mws_pairs = npairs; v->mws_cap = cap;
    }
    OrbxBowMatchArgs A;
    A.kf_frame = d_kf_frame; A.f_frame = d_f_frame; A.kps = (const OrbxKp28*)d_keypoints; A.desc = d_descriptors;
    A.kf_valid = d_kf_valid; A.f_valid = d_f_valid; A.kf_mode = kf_mode;
    A.nnratio = nnratio; A.check_orientation = check_orientation; A.th_low = 50;   // ORBmatcher::TH_LOW
    A.match = d_match; A.bin_of = v->d_bin_of; A.taken = v->d_taken; A.hist = v->d_hist; A.nmatches = d_nmatches;
    orbx_launch_bow_match(v->O, A, v->d_n, cap, npairs, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_search_by_bow_device(orbx_vocabulary* v, int npairs, const int32_t* d_kf_frame, const int32_t* d_f_frame,
                                         const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_kf_valid,
                                         float nnratio, int check_orientation, int32_t* d_match, int32_t* d_nmatches,
                                         void* cuda_stream)
{
    return search_by_bow_device_impl(v, npairs, d_kf_frame, d_f_frame, d_keypoints, d_descriptors, d_kf_valid, nullptr, 0, nnratio,
                                     check_orientation, d_match, d_nmatches, cuda_stream);
}

extern "C" int orbx_search_by_bow_kf_device(orbx_vocabulary* v, int npairs, const int32_t* d_kf1_frame, const int32_t* d_kf2_frame,
                                            const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_valid1,
                                            const uint8_t* d_valid2, float nnratio, int check_orientation, int32_t* d_match12,
                                            int32_t* d_nmatches, void* cuda_stream)
{
    return search_by_bow_device_impl(v, npairs, d_kf1_frame, d_kf2_frame, d_keypoints, d_descriptors, d_valid1, d_valid2, 1, nnratio,
                                     check_orientation, d_match12, d_nmatches, cuda_stream);
}

// host convenience for one (keyframe, frame) pair: transforms both descriptor sets, then matches
static int search_by_bow_host(orbx_vocabulary* v, const OrbxKeyPoint* kf_keypoints, const uint8_t* kf_descriptors, int n_kf,
                              const uint8_t* kf_valid, const OrbxKeyPoint* f_keypoints, const uint8_t* f_descriptors, int n_f,
                              const uint8_t* f_valid, int kf_mode, int levelsup, float nnratio, int check_orientation,
                              int32_t* match_f, int32_t* nmatches)
{
    const int n_out = kf_mode ? n_kf : n_f;
    if (!v || n_kf < 0 || n_f < 0 || !nmatches || (n_out > 0 && !match_f)) return fail(ORBX_ERR_INVALID, "bad argument");
    VLOCK(v);                                    // transform + match are one critical section
    *nmatches = 0;
    for (int j = 0; j < n_out; j++) match_f[j] = -1;
    if (n_kf == 0 || n_f == 0) return ORBX_OK;
    if (!kf_keypoints || !kf_descriptors || !f_keypoints || !f_descriptors) return fail(ORBX_ERR_INVALID, "NULL argument");
    const int cap = std::max(n_kf, n_f);
    std::vector<uint8_t> desc((size_t)2 * cap * 32, 0);
    memcpy(desc.data(), kf_descriptors, (size_t)n_kf * 32);
    memcpy(desc.data() + (size_t)cap * 32, f_descriptors, (size_t)n_f * 32);
    const int32_t counts[2] = {n_kf, n_f};
    int rc = orbx_bow_transform(v, desc.data(), counts, 2, cap, levelsup);
    if (rc != ORBX_OK) return rc;
    uint8_t* pool = nullptr;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_kp = al((size_t)2 * cap * 28), b_d = al((size_t)2 * cap * 32), b_v = al((size_t)cap), b_m = al((size_t)cap * 4);
    CK(cudaMalloc(&pool, b_kp + b_d + 2 * b_v + b_m + 256 + 256));
    uint8_t *p_kp = pool, *p_d = p_kp + b_kp, *p_v = p_d + b_d, *p_v2 = p_v + b_v, *p_m = p_v2 + b_v, *p_idx = p_m + b_m, *p_nm = p_idx + 256;
    cudaError_t e;
    do {
        const int32_t idx[2] = {0, 1};
        if ((e = cudaMemcpy(p_kp, kf_keypoints, (size_t)n_kf * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_kp + (size_t)cap * 28, f_keypoints, (size_t)n_f * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_d, desc.data(), desc.size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (kf_valid && (e = cudaMemcpy(p_v, kf_valid, (size_t)n_kf, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (f_valid && (e = cudaMemcpy(p_v2, f_valid, (size_t)n_f, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(p_idx, idx, 8, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        rc = search_by_bow_device_impl(v, 1, (const int32_t*)p_idx, (const int32_t*)p_idx + 1, (const OrbxKeyPoint*)p_kp, p_d,
                                       kf_valid ? p_v : nullptr, f_valid ? p_v2 : nullptr, kf_mode, nnratio, check_orientation,
                                       (int32_t*)p_m, (int32_t*)p_nm, nullptr);
        if (rc != ORBX_OK) { e = cudaSuccess; break; }
        if ((e = cudaMemcpy(match_f, p_m, (size_t)n_out * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(nmatches, p_nm, 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(pool);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1489-1646)
// stream-ordered allocations are used for per-call scratch: keep freed blocks in the pool instead of returning them to
// the driver at every synchronisation (the default release threshold is 0)
void orbx_keep_mempool(int device)
{
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done[device] = true;
}

extern "C" int orbx_search_by_projection_device(const OrbxProjectionPair* pairs, int npairs, const float* camera9,
                                                const float* scale_factors, int nlevels, float th, int check_orientation,
                                                int device, void* cuda_stream)
{
    if (npairs <= 0) return ORBX_OK;
    if (!pairs || !camera9 || !scale_factors || nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    orbx_keep_mempool(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    size_t tot_last = 0; int max_cur = 0;
    for (int p = 0; p < npairs; p++) {
        const OrbxProjectionPair& q = pairs[p];
        if (q.n_cur < 0 || q.n_last < 0 || !q.match || !q.nmatches || q.mode < 0 || q.mode > 2) return fail(ORBX_ERR_INVALID, "bad pair");
        if (q.n_cur > 10000) return fail(ORBX_ERR_UNSUPPORTED, "more than 10000 keypoints in the current frame");
        if ((q.n_cur > 0 && (!q.cur_keypoints || !q.cur_descriptors)) ||
            (q.n_last > 0 && (!q.last_keypoints || !q.last_xyz || !q.last_descriptors || !q.last_flags))) return fail(ORBX_ERR_INVALID, "NULL array in pair");
        tot_last += (size_t)q.n_last; max_cur = std::max(max_cur, q.n_cur);
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_pairs = al((size_t)npairs * sizeof(OrbxProjPairDev)), b_q = al(std::max<size_t>(tot_last, 1) * sizeof(OrbxProjQuery)),
                 b_a = al(std::max<size_t>(tot_last, 1) * 4), b_sf = al((size_t)nlevels * 4);
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, b_pairs + b_q + b_a + b_sf, st));
    OrbxProjQuery* d_q = (OrbxProjQuery*)(pool + b_pairs);
    int* d_a = (int*)(pool + b_pairs + b_q);
    float* d_sf = (float*)(pool + b_pairs + b_q + b_a);
    std::vector<OrbxProjPairDev> hp(npairs);
    size_t off = 0;
    for (int p = 0; p < npairs; p++) {
        const OrbxProjectionPair& q = pairs[p];
        OrbxProjPairDev& d = hp[p];
        d.cur_kps = (const OrbxKp28*)q.cur_keypoints; d.cur_desc = q.cur_descriptors; d.cur_u_right = q.cur_u_right;
        d.cur_occupied = q.cur_occupied; d.n_cur = q.n_cur;
        d.last_kps = (const OrbxKp28*)q.last_keypoints; d.last_xyz = q.last_xyz; d.last_desc = q.last_descriptors;
        d.last_flags = q.last_flags; d.n_last = q.n_last;
        memcpy(d.Tcw, q.Tcw, sizeof d.Tcw); d.mode = q.mode;
        d.match = q.match; d.nmatches = q.nmatches;
        d.query = d_q + off; d.assign = d_a + off; off += (size_t)q.n_last;
    }
    OrbxProjCam cam;
    cam.fx = camera9[0]; cam.fy = camera9[1]; cam.cx = camera9[2]; cam.cy = camera9[3]; cam.mbf = camera9[4];
    cam.minX = camera9[5]; cam.maxX = camera9[6]; cam.minY = camera9[7]; cam.maxY = camera9[8];
    cudaError_t e;
    do {
        // pageable staging copies: cudaMemcpyAsync from pageable memory returns after the data has been staged
        if ((e = cudaMemcpyAsync(pool, hp.data(), (size_t)npairs * sizeof(OrbxProjPairDev), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(d_sf, scale_factors, (size_t)nlevels * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        orbx_launch_search_projection((const OrbxProjPairDev*)pool, npairs, max_cur, cam, d_sf, th, check_orientation, st);
        e = cudaGetLastError();
    } while (0);
    cudaFreeAsync(pool, st);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return ORBX_OK;
}

extern "C" int orbx_search_by_projection(const OrbxProjectionPair* pair, const float* camera9, const float* scale_factors,
                                         int nlevels, float th, int check_orientation, int device)
{
    if (!pair || !pair->match || !pair->nmatches) return fail(ORBX_ERR_INVALID, "bad argument");
    if (orbx_device_count() <= 0) return fail(ORBX_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    const int nc = pair->n_cur, nl = pair->n_last;
    if (nc < 0 || nl < 0) return fail(ORBX_ERR_INVALID, "bad sizes");
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t c1 = std::max(nc, 1), l1 = std::max(nl, 1);
    const size_t b[] = {al(c1 * 28), al(c1 * 32), al(c1 * 4), al(c1), al(l1 * 28), al(l1 * 12), al(l1 * 32), al(l1), al(c1 * 4), 256};
    size_t tot = 0; for (size_t x : b) tot += x;
    orbx_keep_mempool(device);
    uint8_t* pool = nullptr;
    CK(cudaMallocAsync(&pool, tot, 0));
    uint8_t* p[10]; { uint8_t* q = pool; for (int i = 0; i < 10; i++) { p[i] = q; q += b[i]; } }
    OrbxProjectionPair d = *pair;
    cudaError_t e = cudaSuccess;
    int rc = ORBX_OK;
    do {
        auto up = [&](uint8_t* dst, const void* src, size_t n) { return (!src || !n) ? cudaSuccess : cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, 0); };
        if ((e = up(p[0], pair->cur_keypoints, (size_t)nc * 28)) != cudaSuccess) break;
        if ((e = up(p[1], pair->cur_descriptors, (size_t)nc * 32)) != cudaSuccess) break;
        if ((e = up(p[2], pair->cur_u_right, (size_t)nc * 4)) != cudaSuccess) break;
        if ((e = up(p[3], pair->cur_occupied, (size_t)nc)) != cudaSuccess) break;
        if ((e = up(p[4], pair->last_keypoints, (size_t)nl * 28)) != cudaSuccess) break;
        if ((e = up(p[5], pair->last_xyz, (size_t)nl * 12)) != cudaSuccess) break;
        if ((e = up(p[6], pair->last_descriptors, (size_t)nl * 32)) != cudaSuccess) break;
        if ((e = up(p[7], pair->last_flags, (size_t)nl)) != cudaSuccess) break;
        d.cur_keypoints = (const OrbxKeyPoint*)p[0]; d.cur_descriptors = p[1];
        d.cur_u_right = pair->cur_u_right ? (const float*)p[2] : nullptr; d.cur_occupied = pair->cur_occupied ? p[3] : nullptr;
        d.last_keypoints = (const OrbxKeyPoint*)p[4]; d.last_xyz = (const float*)p[5]; d.last_descriptors = p[6]; d.last_flags = p[7];
        d.match = (int32_t*)p[8]; d.nmatches = (int32_t*)p[9];
        rc = orbx_search_by_projection_device(&d, 1, camera9, scale_factors, nlevels, th, check_orientation, device, nullptr);
        if (rc != ORBX_OK) break;
        if (nc > 0 && (e = cudaMemcpy(pair->match, p[8], (size_t)nc * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(pair->nmatches, p[9], 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFreeAsync(pool, 0);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

extern "C" int orbx_search_by_bow(orbx_vocabulary* v, const OrbxKeyPoint* kf_keypoints, const uint8_t* kf_descriptors, int n_kf,
                                  const uint8_t* kf_valid, const OrbxKeyPoint* f_keypoints, const uint8_t* f_descriptors, int n_f,
                                  int levelsup, float nnratio, int check_orientation, int32_t* match_f, int32_t* nmatches)
{
    return search_by_bow_host(v, kf_keypoints, kf_descriptors, n_kf, kf_valid, f_keypoints, f_descriptors, n_f, nullptr, 0, levelsup,
                              nnratio, check_orientation, match_f, nmatches);
}

extern "C" int orbx_search_by_bow_kf(orbx_vocabulary* v, const OrbxKeyPoint* kf1_keypoints, const uint8_t* kf1_descriptors, int n1,
                                     const uint8_t* valid1, const OrbxKeyPoint* kf2_keypoints, const uint8_t* kf2_descriptors, int n2,
                                     const uint8_t* valid2, int levelsup, float nnratio, int check_orientation, int32_t* match12,
                                     int32_t* nmatches)
{
    return search_by_bow_host(v, kf1_keypoints, kf1_descriptors, n1, valid1, kf2_keypoints, kf2_descriptors, n2, valid2, 1, levelsup,
                              nnratio, check_orientation, match12, nmatches);
}

// ---------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchForTriangulation (ORBmatcher.cc:738-916) — orbx_bow.cu
extern "C" int orbx_search_for_triangulation_device(orbx_vocabulary* v, int npairs, const int32_t* d_kf1_frame, const int32_t* d_kf2_frame,
                                                    const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_has_mp,
                                                    const float* d_u_right, const float* d_geom, const float* scale_factors,
                                                    const float* level_sigma2, int nlevels, int only_stereo, int check_orientation,
                                                    int32_t* d_match12, int32_t* d_nmatches, void* cuda_stream)
{
    if (!v) return fail(ORBX_ERR_INVALID, "vocabulary is NULL");
    VLOCK(v);
    if (v->frames <= 0) return fail(ORBX_ERR_STATE, "no transform has run yet");
    if (npairs <= 0) return ORBX_OK;
    if (!d_kf1_frame || !d_kf2_frame || !d_keypoints || !d_descriptors || !d_geom || !scale_factors || !level_sigma2 || !d_match12 ||
        !d_nmatches || nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(v->device));
    const int cap = v->cap;
    if ((size_t)npairs > v->mws_pairs || (size_t)cap != v->mws_cap) {
        CK(cudaDeviceSynchronize());
        cudaFree(v->d_mws); v->d_mws = nullptr;
        CK(cudaMalloc(&v->d_mws, (size_t)npairs * cap * 8 + (size_t)npairs * 32 * 4));
        v->d_bin_of = (int*)v->d_mws; v->d_taken = v->d_bin_of + (size_t)npairs * cap; v->d_hist = v->d_taken + (size_t)npairs * cap;
        v->mws_pairs = npairs; v->mws_cap = cap;
    }
    OrbxBowTriArgs T = {};
    T.kf1_frame = d_kf1_frame; T.kf2_frame = d_kf2_frame; T.kps = (const OrbxKp28*)d_keypoints; T.desc = d_descriptors;
    T.has_mp = d_has_mp; T.u_right = d_u_right; T.geom = d_geom;
    for (int l = 0; l < nlevels; l++) { T.scale_factors[l] = scale_factors[l]; T.level_sigma2[l] = level_sigma2[l]; }
    T.only_stereo = only_stereo; T.check_orientation = check_orientation; T.th_low = 50;   // ORBmatcher::TH_LOW
    T.match = d_match12; T.bin_of = v->d_bin_of; T.hist = v->d_hist; T.nmatches = d_nmatches;
    orbx_launch_bow_triangulation(v->O, T, v->d_taken, v->d_n, cap, npairs, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbx_search_for_triangulation(orbx_vocabulary* v, const OrbxKeyPoint* kf1_keypoints, const uint8_t* kf1_descriptors, int n1,
                                             const uint8_t* has_mp1, const float* u_right1, const OrbxKeyPoint* kf2_keypoints,
                                             const uint8_t* kf2_descriptors, int n2, const uint8_t* has_mp2, const float* u_right2,
                                             const float* geom28, const float* scale_factors, const float* level_sigma2, int nlevels,
                                             int levelsup, int only_stereo, int check_orientation, int32_t* match12, int32_t* nmatches)
{
    if (!v || n1 < 0 || n2 < 0 || !nmatches || (n1 > 0 && !match12) || !geom28) return fail(ORBX_ERR_INVALID, "bad argument");
    VLOCK(v);                                    // transform + match are one critical section
    if ((u_right1 == nullptr) != (u_right2 == nullptr) || (has_mp1 == nullptr) != (has_mp2 == nullptr))
        return fail(ORBX_ERR_INVALID, "u_right / has_mp must be given for both keyframes or for neither");
    *nmatches = 0;
    for (int j = 0; j < n1; j++) match12[j] = -1;
    if (n1 == 0 || n2 == 0) return ORBX_OK;
    if (!kf1_keypoints || !kf1_descriptors || !kf2_keypoints || !kf2_descriptors) return fail(ORBX_ERR_INVALID, "NULL argument");
    const int cap = std::max(n1, n2);
    std::vector<uint8_t> desc((size_t)2 * cap * 32, 0);
    memcpy(desc.data(), kf1_descriptors, (size_t)n1 * 32);
    memcpy(desc.data() + (size_t)cap * 32, kf2_descriptors, (size_t)n2 * 32);
    const int32_t counts[2] = {n1, n2};
    int rc = orbx_bow_transform(v, desc.data(), counts, 2, cap, levelsup);
    if (rc != ORBX_OK) return rc;
    HostPack P;
    const size_t i_kp = P.add((size_t)2 * cap * 28), i_d = P.add((size_t)2 * cap * 32), i_mp = P.add((size_t)2 * cap), i_ur = P.add((size_t)2 * cap * 4),
                 i_g = P.add(28 * 4), i_m = P.add((size_t)cap * 4), i_idx = P.add(8), i_nm = P.add(4);
    CK(cudaMalloc(&P.pool, P.tot));
    cudaError_t e;
    do {
        const int32_t idx[2] = {0, 1};
        if ((e = cudaMemset(P.pool, 0, P.tot)) != cudaSuccess) break;
        if ((e = cudaMemcpy(P.at(i_kp), kf1_keypoints, (size_t)n1 * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(P.at(i_kp) + (size_t)cap * 28, kf2_keypoints, (size_t)n2 * 28, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(P.at(i_d), desc.data(), desc.size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (has_mp1 && (e = cudaMemcpy(P.at(i_mp), has_mp1, (size_t)n1, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (has_mp2 && (e = cudaMemcpy(P.at(i_mp) + cap, has_mp2, (size_t)n2, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (u_right1 && (e = cudaMemcpy(P.at(i_ur), u_right1, (size_t)n1 * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (u_right2 && (e = cudaMemcpy(P.at(i_ur) + (size_t)cap * 4, u_right2, (size_t)n2 * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(P.at(i_g), geom28, 28 * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = cudaMemcpy(P.at(i_idx), idx, 8, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        rc = orbx_search_for_triangulation_device(v, 1, (const int32_t*)P.at(i_idx), (const int32_t*)P.at(i_idx) + 1,
                                                  (const OrbxKeyPoint*)P.at(i_kp), P.at(i_d), has_mp1 ? P.at(i_mp) : nullptr,
                                                  u_right1 ? (const float*)P.at(i_ur) : nullptr, (const float*)P.at(i_g), scale_factors,
                                                  level_sigma2, nlevels, only_stereo, check_orientation, (int32_t*)P.at(i_m),
                                                  (int32_t*)P.at(i_nm), nullptr);
        if (rc != ORBX_OK) { e = cudaSuccess; break; }
        if ((e = cudaMemcpy(match12, P.at(i_m), (size_t)n1 * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(nmatches, P.at(i_nm), 4, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(P.pool);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

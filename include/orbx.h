/*
 * orbx.h — C ABI of the B200-native ORB front-end (liborbx.so, hand-written sm_100a CUDA kernels).
 *
 * This is the drop-in boundary for ORB-SLAM2's data-parallel hot path. The reference has no FFI: its seam is the
 * C++ class ORB_SLAM2::ORBextractor (include/ORBextractor.h:51-145) and the static function
 * ORBmatcher::DescriptorDistance (include/ORBmatcher.h:50). The C++ shim in orb_slam2_commit_b200/host/ keeps that
 * class surface and calls only the functions below; INTEGRATION.md shows the bindings.
 * Every entry point cites the reference interface it replaces (file:line relative to the reference tree).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns an orbx_status (0 = ok);
 * there is NO CPU fallback — without a CUDA device every compute call returns ORBX_ERR_CUDA.
 * An orbx_extractor is not re-entrant (like the reference instance, which mutates mvImagePyramid); two different
 * instances may be driven concurrently from two host threads (Frame.cc:80-84 does exactly that for stereo).
 */
#ifndef ORBX_H
#define ORBX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define ORBX_ABI_VERSION 1

typedef enum {
    ORBX_OK = 0,
    ORBX_ERR_INVALID = 1,      /* bad argument (null pointer, non-positive size, nlevels out of range) */
    ORBX_ERR_UNSUPPORTED = 2,  /* geometry the reference itself cannot handle (a level < 62 px, portrait level
                                  => nIni == 0, ORBextractor.cc:567,846) or beyond this build's limits (> 4096 px) */
    ORBX_ERR_CAPACITY = 3,     /* caller buffer too small for the keypoints produced */
    ORBX_ERR_CUDA = 4,         /* CUDA runtime error / no device; see orbx_last_error() */
    ORBX_ERR_STATE = 5         /* call order (e.g. pyramid requested before any extract) */
} orbx_status;

/* Bit-identical to cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct OrbxKeyPoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} OrbxKeyPoint;

typedef struct orbx_extractor orbx_extractor;

const char* orbx_last_error(void);          /* thread-local text of the last failure */
int orbx_abi_version(void);
int orbx_device_count(void);                /* 0 when no CUDA device is visible */

/* ---- ORBextractor::ORBextractor(int,float,int,int,int)  (ORBextractor.h:61, ORBextractor.cc:416-490) ---- */
int orbx_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                int device, orbx_extractor** out);
/* ---- ~ORBextractor  (ORBextractor.h:64) ---- */
void orbx_destroy(orbx_extractor* h);

/* ---- GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 *      GetInverseScaleSigmaSquares  (ORBextractor.h:81-101); arrays hold nlevels entries, any may be NULL.
 *      features_per_level = mnFeaturesPerLevel (ORBextractor.cc:446-457), umax16 = the IC_Angle disc table. ---- */
int   orbx_get_levels(const orbx_extractor* h);
float orbx_get_scale_factor(const orbx_extractor* h);
int   orbx_get_tables(const orbx_extractor* h, float* scale_factors, float* inv_scale_factors,
                      float* level_sigma2, float* inv_level_sigma2, int32_t* features_per_level, int32_t* umax16);

/* Pre-size the HBM-resident working set for frames of width x height, up to max_batch frames per call.
 * Optional: the extract calls do it lazily. Returns ORBX_ERR_UNSUPPORTED for geometry the reference cannot run. */
int orbx_reserve(orbx_extractor* h, int width, int height, int max_batch);
/* Upper bound of keypoints per frame for the reserved geometry (sum over levels of quota + slack). */
int orbx_max_keypoints(const orbx_extractor* h);

/* ---- ORBextractor::operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray descriptors)
 *      (ORBextractor.h:77, ORBextractor.cc:1138-1211). Host buffers in, host buffers out, synchronous.
 *      image: 8-bit single channel, `stride` bytes between rows. mask is ignored by the reference (ORBextractor.h:68).
 *      keypoints[cap], descriptors[cap*32] (row i <-> keypoint i, CV_8U 32-byte rows), *nkp = number produced.
 *      Empty image (NULL / w<=0 / h<=0): returns ORBX_OK with *nkp = 0 and outputs untouched (ORBextractor.cc:1141). ---- */
int orbx_extract(orbx_extractor* h, const uint8_t* image, int width, int height, int stride,
                 OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors);

/* Frame-batched form of the same call (frames are independent; BASELINE configs 3 and 5): n images of identical
 * size; images[i] points at frame i. keypoints[n*cap], descriptors[n*cap*32], nkp[n]. Pinned host memory makes the
 * copies asynchronous; pageable memory works too. */
int orbx_extract_batch(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height, int stride,
                       OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors);

/* The same call for a STREAM of batches (a video, a dataset): begin() enqueues the uploads, kernels and downloads of one batch
 * and returns; end() waits for the OLDEST batch begun and returns its status (ORBX_ERR_CAPACITY as above). Two batches may be
 * in flight, so that the uploads of batch i+1 overlap the kernels of batch i (a third begin(), a batch of another shape or any
 * other extract / stereo / pyramid call on the handle first completes what is in flight; a capacity overflow of a batch that
 * was completed that way is still visible in its nkp). Host buffers must be distinct per batch in flight, stay valid
 * until its end(), and cap must equal orbx_max_keypoints() after orbx_reserve(). Results are those of orbx_extract_batch. */
int orbx_extract_batch_begin(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height, int stride,
                             OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors);
int orbx_extract_batch_end(orbx_extractor* h);

/* Colour input: Tracking::GrabImage{Stereo,RGBD,Monocular} converts 3- and 4-channel frames with cv::cvtColor
 * (CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY, Tracking.cc:174-199, 215-237, 246-261) before the extractor
 * sees them. Here that conversion (OpenCV 4.x 8-bit arithmetic) is fused into the level-0 kernel: `images[i]` are
 * interleaved frames with `channels` in {1, 3, 4} and `stride` bytes per row; rgb != 0 <=> channel 0 is R (mbRGB). */
int orbx_extract_batch_color(orbx_extractor* h, const uint8_t* const* images, int n, int width, int height, int stride,
                             int channels, int rgb, OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors);

/* Stereo rectification in front of the extractor: Examples/Stereo/stereo_euroc.cc:97-98 builds a CV_32FC1 map pair with
 * cv::initUndistortRectifyMap once, and :136-137 runs cv::remap(im, imRect, M1, M2, cv::INTER_LINEAR) on every frame.
 * orbx_set_rectify_maps takes that map pair (map_stride in floats; source frames are src_width x src_height) and converts
 * it to OpenCV's fixed-point form once; orbx_extract_batch_rectified then takes UNRECTIFIED frames and fuses the remap
 * (OpenCV 4.x 8-bit arithmetic, BORDER_CONSTANT 0) into the level-0 kernel: pyramid level 0 IS the rectified frame
 * (map_width x map_height; read it back with orbx_pyramid_level). Passing two NULL maps clears the state.
 * Map values must stay below 2^26 in magnitude. */
int orbx_set_rectify_maps(orbx_extractor* h, const float* map1, const float* map2, int map_width, int map_height,
                          int map_stride, int src_width, int src_height);
/* cv::initUndistortRectifyMap(K, D, R, P, size, CV_32F, M1, M2) itself (stereo_euroc.cc:96-97), on the device in OpenCV 4.x's
 * double arithmetic (bit-exact against cv2 4.13): K9 / R9 row-major 3x3 doubles, D = 0, 4, 5, 8 or 12 distortion coefficients
 * (k1, k2, p1, p2[, k3[, k4, k5, k6[, s1, s2, s3, s4]]]), P row-major with p_cols = 3 or 4 columns (only the left 3x3 block is
 * read). orbx_init_undistort_rectify_map returns the f32 map pair (width x height each); orbx_set_rectify_camera builds the
 * fixed-point map straight in HBM and installs it — the device-side equivalent of the two calls :96-97 + orbx_set_rectify_maps. */
int orbx_init_undistort_rectify_map(const double* K9, const double* D, int nD, const double* R9, const double* P, int p_cols,
                                    int width, int height, float* map1, float* map2, int device);
int orbx_set_rectify_camera(orbx_extractor* h, const double* K9, const double* D, int nD, const double* R9, const double* P,
                            int p_cols, int map_width, int map_height, int src_width, int src_height);
/* size of the rectified frame (= the map's own size, which may differ from the source size): what orbx_reserve and the
 * keypoint capacity of the rectified calls are based on */
int orbx_rectify_map_size(const orbx_extractor* h, int* map_width, int* map_height);
int orbx_extract_batch_rectified(orbx_extractor* h, const uint8_t* const* images, int n, int stride,
                                 OrbxKeyPoint* keypoints, int cap, int* nkp, uint8_t* descriptors);

/* device-resident form of the rectified call (see orbx_extract_device) */
int orbx_extract_device_rectified(orbx_extractor* h, const uint8_t* d_images, int n, int stride, size_t frame_pitch_bytes,
                                  OrbxKeyPoint* d_keypoints, int cap, int32_t* d_nkp, uint8_t* d_descriptors,
                                  void* cuda_stream);

/* Device-resident form: frames already in HBM (frame i at d_images + i*frame_pitch_bytes), outputs stay in HBM.
 * Asynchronous on `cuda_stream` (a cudaStream_t passed as void*; NULL = the extractor's own stream).
 * d_keypoints[n*cap], d_descriptors[n*cap*32], d_nkp[n] (device pointers). Keypoints beyond cap are dropped and
 * d_nkp still reports the true count, so the caller can detect ORBX_ERR_CAPACITY after synchronising. */
int orbx_extract_device(orbx_extractor* h, const uint8_t* d_images, int n, int width, int height, int stride,
                        size_t frame_pitch_bytes, OrbxKeyPoint* d_keypoints, int cap, int32_t* d_nkp,
                        uint8_t* d_descriptors, void* cuda_stream);
int orbx_synchronize(orbx_extractor* h);

/* Measurement hook (bench.py): when enabled, CUDA events are recorded on the launching stream around the four
 * stages of every extract (pyramid, FAST cells, quadtree, orientation+descriptor).
 * orbx_get_stage_ms waits for the last run and returns the mean device time of each stage, in milliseconds, over
 * the pipeline runs recorded since timing was enabled (the most recent 64 at most); *nruns = how many. */
int orbx_enable_timing(orbx_extractor* h, int on);
int orbx_get_stage_ms(orbx_extractor* h, float* ms4, int* nruns);

/* ---- std::vector<cv::Mat> mvImagePyramid  (ORBextractor.h:104; read by Frame.cc:556,681-700) ----
 * Level geometry of the last extract, and a copy of level `level` of frame `frame` INCLUDING its 19-px
 * BORDER_REFLECT_101 apron into dst ((h+38) rows of (w+38) bytes, dst_stride >= w+38) — the layout of the reference's
 * backing `temp` buffer (ORBextractor.cc:1225-1229); the payload starts at dst + 19*dst_stride + 19. */
int orbx_level_size(const orbx_extractor* h, int level, int* width, int* height);
int orbx_pyramid_level(orbx_extractor* h, int frame, int level, uint8_t* dst, int dst_stride);
/* The whole pyramid of one frame in ONE copy: the frame's raw block (all levels, each (h+38) rows of `pitch` bytes with
 * the apron physically around the payload) lands in a pinned host mirror owned by the handle; level `level`'s payload
 * pixel (0,0) sits at host_block + payload_offset and rows are `pitch` bytes apart, so
 *   cv::Mat(h, w, CV_8UC1, (void*)(host_block + payload_offset), pitch)
 * IS the reference's mvImagePyramid[level] (a view at (19,19) of `temp`, ORBextractor.cc:1225-1229): the pointer
 * arithmetic of Frame.cc:681-700 finds the 19-px apron on every side. After orbx_set_pyramid_mirror(h, 1) a single-frame
 * orbx_extract downloads the block itself, asynchronously in the stream of its kernels (no extra synchronisation);
 * orbx_pyramid_mirror then just returns the pointer. The mirror stays valid until the next call on the handle. */
int orbx_set_pyramid_mirror(orbx_extractor* h, int on);
int orbx_pyramid_level_layout(const orbx_extractor* h, int level, size_t* payload_offset, int* pitch, size_t* block_bytes);
int orbx_pyramid_mirror(orbx_extractor* h, int frame, const uint8_t** host_block);
/* Device view of the same level: payload origin, pitch in bytes (apron lies around it in HBM). */
int orbx_pyramid_level_device(orbx_extractor* h, int frame, int level, const uint8_t** d_payload, int* pitch);

/* Stage tap for parity tests: the pre-quadtree candidate list of (frame, level) in the reference's order
 * (cell-row-major, then FAST's row-major; ORBextractor.cc:903-912), coordinates relative to minBorder. */
int orbx_debug_candidates(orbx_extractor* h, int frame, int level, OrbxKeyPoint* out, int cap, int* n);
int orbx_debug_level_counts(orbx_extractor* h, int frame, int32_t* counts /* nlevels */);
/* Stage tap: the payload (w x h) of level `level` after cv::GaussianBlur(.., Size(7,7), 2, 2, BORDER_REFLECT_101) —
 * `workingMat` of ORBextractor.cc:1188-1190, which computeDescriptors samples. */
int orbx_debug_blurred_level(orbx_extractor* h, int frame, int level, uint8_t* dst, int dst_stride);

/* ---- ORBmatcher::DescriptorDistance (ORBmatcher.h:50, ORBmatcher.cc:1844-1860) + the best / second-best search
 *      idiom around it (ORBmatcher.cc:84-126): for every query row the FIRST train index attaining the minimum
 *      distance, that distance, and the second-smallest distance counted with multiplicity; all three are 256 / -1 /
 *      256 when no train row is closer than 256 (the reference's initial values). 32-byte rows, 4-byte aligned. ---- */
int orbx_hamming_top2(const uint8_t* query, int nq, const uint8_t* train, int nt,
                      int32_t* idx1, int32_t* dist1, int32_t* dist2, int device);
/* Device-resident, asynchronous. Result per query packed as (dist1 << 48) | (dist2 << 32) | uint32(idx1);
 * `d_packed[nq]` must be initialised with orbx_hamming_init_device (or hold a previous partial result: results
 * are MERGED into it, which is how train shards — on one GPU or across GPUs — combine). index_base is added to
 * the local train index. */
int orbx_hamming_init_device(uint64_t* d_packed, int nq, void* cuda_stream);
int orbx_hamming_top2_device(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, int64_t index_base,
                             uint64_t* d_packed, void* cuda_stream);
/* Merge `nparts` packed partial results (part p at d_parts + p*nq) — e.g. the all-gathered per-GPU results — and
 * unpack into idx1/dist1/dist2 (device pointers, any may be NULL). */
int orbx_hamming_merge_device(const uint64_t* d_parts, int nparts, int nq,
                              int32_t* d_idx1, int32_t* d_dist1, int32_t* d_dist2, void* cuda_stream);

/* ---- Frame::ComputeStereoMatches, Hamming stage (Frame.cc:554-663): for every left keypoint the best right
 *      keypoint among the row-band candidates with octave within +-1 and uR in [uL-maxD, uL-minD], initial best
 *      TH_HIGH = 100, strict '<' in ascending right index. Host buffers, synchronous.
 *      best_idx_r[i] = -1 and best_dist[i] = 100 when nothing beat 100. ---- */
int orbx_stereo_hamming(const OrbxKeyPoint* kp_left, const uint8_t* desc_left, int n_left,
                        const OrbxKeyPoint* kp_right, const uint8_t* desc_right, int n_right,
                        int rows, const float* scale_factors, int nlevels, float minD, float maxD,
                        int32_t* best_idx_r, int32_t* best_dist, int device);

/* ---- Windowed best / second-best search: the candidate loop of ORBmatcher::SearchByProjection(Frame&,
 *      vector<MapPoint*>&, th) (ORBmatcher.cc:46-142) over Frame::GetFeaturesInArea (Frame.cc:388-444) on the 64x48
 *      grid of AssignFeaturesToGrid / PosInGrid (Frame.cc:254-271, 446-460). One query = one projected map point:
 *      window centre (x, y), radius r (already scaled by mvScaleFactors[level]), level range [min_level, max_level]
 *      (the reference passes nPredictedLevel-1, nPredictedLevel), xr = mTrackProjXR (only used when u_right is given).
 *      occupied[i] != 0 <=> keypoint i already has an observed map point (:90-92); u_right = mvuRight (:95-100);
 *      both may be NULL. minX/minY/invW/invH = mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv.
 *      Outputs per query are the five values the reference loop ends with: bestIdx (-1 = none), bestDist, bestLevel,
 *      bestDist2, bestLevel2; the caller applies `bestDist <= TH_HIGH` and the NN-ratio rule (:129-137).
 *      Host buffers, synchronous; at most 14000 keypoints per frame. ---- */
typedef struct OrbxWindowQuery { float x, y, r; int32_t min_level, max_level; float xr; } OrbxWindowQuery;
int orbx_window_top2(const OrbxKeyPoint* keypoints, const uint8_t* descriptors, int n,
                     const uint8_t* occupied, const float* u_right,
                     float minX, float minY, float invW, float invH,
                     const OrbxWindowQuery* queries, const uint8_t* query_descriptors, int nq,
                     int32_t* best_idx, int32_t* best_dist, int32_t* best_level, int32_t* best_dist2, int32_t* best_level2,
                     int device);

/* ---- Train-sharded matching across GPUs with the exchange FUSED into the matcher (no NCCL call): the last CTA of
 *      every query tile stores the rank's merged top-2 straight into every peer's landing buffer over NVLink
 *      (CUDA IPC peer mappings) and bumps the peers' arrival counters; a one-block kernel waits for the counters
 *      (bounded) and merges. One process per GPU; all ranks call in the same order.
 *      Setup: orbx_peer_create on every rank -> exchange the 64-byte handles by any host channel (e.g.
 *      torch.distributed.all_gather_object) -> orbx_peer_connect with all `world` handles in rank order.
 *      d_status (device int, may be NULL) is set to 1 if a peer never arrived within the bounded wait. ---- */
typedef struct orbx_peer_matcher orbx_peer_matcher;
#define ORBX_IPC_HANDLE_BYTES 64
int orbx_peer_create(int nq_max, int world, int rank, int device, orbx_peer_matcher** out,
                     uint8_t handle_out[ORBX_IPC_HANDLE_BYTES]);
int orbx_peer_connect(orbx_peer_matcher* m, const uint8_t* all_handles /* world x 64 bytes, rank order */);
int orbx_peer_hamming_top2(orbx_peer_matcher* m, const uint8_t* d_query, int nq, const uint8_t* d_train_shard, int nt,
                           int64_t index_base, int32_t* d_idx1, int32_t* d_dist1, int32_t* d_dist2, int32_t* d_status,
                           void* cuda_stream);
void orbx_peer_destroy(orbx_peer_matcher* m);

/* ---- Frame::ComputeStereoMatches, complete (Frame.cc:547-788): row-band Hamming as above, then the 11x11 SAD of
 *      centre-normalised windows slid over +-5 columns on the level pyramids of BOTH extractors (left / right must
 *      have just extracted the left / right image of the pair, same geometry, same device — their pyramids are read
 *      where they lie in HBM), the parabola sub-pixel fit, disparity/depth, and the 1.5*1.4*median outlier cut.
 *      mbf = Camera.bf, fx = Camera.fx (mb = mbf/fx, Frame.cc:120). Outputs: mvuRight / mvDepth, n_left floats each,
 *      -1 where there is no stereo match. Host buffers, synchronous. ---- */
int orbx_stereo_match(orbx_extractor* left, orbx_extractor* right,
                      const OrbxKeyPoint* kp_left, const uint8_t* desc_left, int n_left,
                      const OrbxKeyPoint* kp_right, const uint8_t* desc_right, int n_right,
                      float mbf, float fx, float* u_right, float* depth);
/* Batched host form of the stereo Frame constructor's hot part (Frame.cc:80-117): for n pairs, ExtractORB on the left and
 * the right image and ComputeStereoMatches, pipelined in chunks over several streams (copies overlap kernels). All
 * output arrays are [n][cap] with cap == orbx_max_keypoints() (identical for both extractors after orbx_reserve with
 * the same geometry); entries i >= n_left[pair] of u_right / depth are unspecified. Pinned host memory recommended. */
int orbx_stereo_extract_batch(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                              const uint8_t* const* images_right, int n, int width, int height, int stride,
                              float mbf, float fx, OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left,
                              OrbxKeyPoint* kp_right, uint8_t* desc_right, int32_t* n_right, int cap,
                              float* u_right, float* depth);
/* begin / end form of the same call for streams of batches (see orbx_extract_batch_begin): up to two batches of pairs in
 * flight, end() completes the oldest and returns its status. The batches in flight are tracked in `left`. */
int orbx_stereo_extract_batch_begin(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                                    const uint8_t* const* images_right, int n, int width, int height, int stride,
                                    float mbf, float fx, OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left,
                                    OrbxKeyPoint* kp_right, uint8_t* desc_right, int32_t* n_right, int cap,
                                    float* u_right, float* depth);
int orbx_stereo_extract_batch_end(orbx_extractor* left);
/* EuRoC-style stereo from RAW frames to mvuRight / mvDepth in one call: both extractors carry the rectification of their camera
 * (orbx_set_rectify_camera = initUndistortRectifyMap on the device, or orbx_set_rectify_maps), cv::remap (stereo_euroc.cc:136-137)
 * is fused into the level-0 kernels, then left / right extraction and ComputeStereoMatches as in orbx_stereo_extract_batch.
 * Frames have the maps' source size; cap = orbx_max_keypoints() after orbx_reserve(h, map_width, map_height, ..). */
int orbx_stereo_extract_batch_rectified(orbx_extractor* left, orbx_extractor* right, const uint8_t* const* images_left,
                                        const uint8_t* const* images_right, int n, int stride, float mbf, float fx,
                                        OrbxKeyPoint* kp_left, uint8_t* desc_left, int32_t* n_left, OrbxKeyPoint* kp_right,
                                        uint8_t* desc_right, int32_t* n_right, int cap, float* u_right, float* depth);
/* Device-resident, batched form: `pairs` stereo pairs whose left / right frames were the frames 0..pairs-1 of the last
 * orbx_extract_device call on `left` / `right`. Keypoints, descriptors and counts are those calls' device outputs
 * ([pairs][cap] layout). d_u_right / d_depth: [pairs][cap] floats, entries i < nl[pair] are written. Asynchronous on
 * cuda_stream (NULL = left's stream). No host table, no host sort: the row band test runs against all right
 * keypoints of the pair in shared memory and the median cut is a radix select on the device. */
int orbx_stereo_match_device(orbx_extractor* left, orbx_extractor* right, int pairs,
                             const OrbxKeyPoint* d_kp_left, const uint8_t* d_desc_left, const int32_t* d_n_left,
                             const OrbxKeyPoint* d_kp_right, const uint8_t* d_desc_right, const int32_t* d_n_right,
                             int cap, float mbf, float fx, float* d_u_right, float* d_depth, void* cuda_stream);

/* ---- Frame::UndistortKeyPoints (Frame.cc:471-506) and Frame::ComputeImageBounds (Frame.cc:508-538):
 *      cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) on the keypoint coordinates. K4 = (fx, fy, cx, cy) and
 *      dist = (k1, k2, p1, p2[, k3]) as the f32 values of mK / mDistCoef (Tracking.cc:60-84); OpenCV 4.x arithmetic
 *      (five fixed-point iterations in f64), so `out` is bit-identical to mvKeysUn. dist[0] == 0 copies the keypoints
 *      unchanged (:474-478). in == out is allowed. Host buffers, synchronous / device buffers, asynchronous. ---- */
int orbx_undistort_keypoints(const OrbxKeyPoint* in, int n, const float* K4, const float* dist, int ndist,
                             OrbxKeyPoint* out, int device);
int orbx_undistort_keypoints_device(const OrbxKeyPoint* d_in, int n, const float* K4, const float* dist, int ndist,
                                    OrbxKeyPoint* d_out, void* cuda_stream);
/* bounds4 = (mnMinX, mnMaxX, mnMinY, mnMaxY) of Frame::ComputeImageBounds for a width x height frame */
int orbx_image_bounds(int width, int height, const float* K4, const float* dist, int ndist, float* bounds4, int device);

/* ---- Bag of words (DBoW2 through ORBVocabulary): Frame::ComputeBoW (Frame.cc:462-469), the L1 score of
 *      KeyFrameDatabase.cc:145,274 / LoopClosing.cc:152, and ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)
 *      (ORBmatcher.cc:175-325) incl. the rotation-histogram check (ComputeThreeMaxima, :1797-1839).
 *      The vocabulary is given in the order of TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1338-1420):
 *      non-root node i+1 has parent[i] (0 = root; a parent precedes its children), is_leaf[i], a 32-byte descriptor and
 *      a weight; words are numbered in order of appearance. scoring: 0 L1_NORM .. 5 DOT_PRODUCT, weighting: 0 TF_IDF,
 *      1 TF, 2 IDF, 3 BINARY (BowVector.h:33-54); ORBvoc.txt is "10 6 0 0". ---- */
typedef struct orbx_vocabulary orbx_vocabulary;
int  orbx_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                       const uint8_t* descriptors, const double* weights, int device, orbx_vocabulary** out);
void orbx_vocab_destroy(orbx_vocabulary* v);
/* Thread safety: ORB-SLAM2 shares one ORBVocabulary between Tracking (Frame::ComputeBoW, Frame.cc:462-469), LocalMapping
 * (KeyFrame::ComputeBoW, SearchForTriangulation) and LoopClosing (score / SearchByBoW), and DBoW2's transform() is const.
 * Here the tree is immutable but the handle keeps the results of the LAST transform, so every entry point that takes an
 * orbx_vocabulary* holds the handle's (recursive) mutex for its duration, the host one-shot forms (orbx_search_by_bow*,
 * orbx_search_for_triangulation) across their whole transform + match sequence. A caller that issues a multi-call
 * sequence itself (orbx_bow_transform followed by orbx_bow_get / orbx_bow_score) brackets it with these two. */
int  orbx_vocab_lock(orbx_vocabulary* v);
int  orbx_vocab_unlock(orbx_vocabulary* v);
int  orbx_vocab_words(const orbx_vocabulary* v);
int  orbx_vocab_nodes(const orbx_vocabulary* v);
/* transform(features, mBowVec, mFeatVec, levelsup) for a batch: frame f has min(counts[f], cap) descriptors at
 * descriptors + f*cap*32 — exactly the [frames][cap] layout orbx_extract_device leaves in HBM. Results stay on the
 * device inside the vocabulary object (used by the score / SearchByBoW calls below) until the next transform;
 * orbx_bow_get copies one frame's results out: word / node per feature, the BowVector (ascending word id, normalised
 * value), the FeatureVector as CSR (ascending node id; fv_off has n_fv+1 entries; fv_feat lists feature indices in
 * ascending order per node). Any output pointer may be NULL. At most 16384 features per frame. */
int  orbx_bow_transform_device(orbx_vocabulary* v, const uint8_t* d_descriptors, const int32_t* d_counts, int frames, int cap,
                               int levelsup, void* cuda_stream);
int  orbx_bow_transform(orbx_vocabulary* v, const uint8_t* descriptors, const int32_t* counts, int frames, int cap, int levelsup);
int  orbx_bow_get(orbx_vocabulary* v, int frame, int32_t* word, int32_t* node, int32_t* bow_id, double* bow_val, int32_t* n_bow,
                  int32_t* fv_node, int32_t* fv_off, int32_t* fv_feat, int32_t* n_fv);
/* L1Scoring::score(BowVector of frame_a[i], BowVector of frame_b[i]) for npairs pairs of frames of the last transform */
int  orbx_bow_score(orbx_vocabulary* v, const int32_t* frame_a, const int32_t* frame_b, int npairs, double* score);
int  orbx_bow_score_device(orbx_vocabulary* v, const int32_t* d_frame_a, const int32_t* d_frame_b, int npairs, double* d_score,
                           void* cuda_stream);
/* SearchByBoW for npairs (keyframe, frame) pairs taken from the last transformed batch: keypoints / descriptors are the
 * batch the transform ran on ([frames][cap]); kf_valid[pair][cap] != 0 <=> that keyframe feature has a map point that is
 * not bad (:212-218), NULL = all. match[pair][j] = keyframe feature matched to frame feature j or -1 (stands for
 * vpMapPointMatches), nmatches[pair] = the return value. TH_LOW = 50. */
int  orbx_search_by_bow_device(orbx_vocabulary* v, int npairs, const int32_t* d_kf_frame, const int32_t* d_f_frame,
                               const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_kf_valid,
                               float nnratio, int check_orientation, int32_t* d_match, int32_t* d_nmatches, void* cuda_stream);
/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (ORBmatcher.cc:589-736): both
 * sides carry map-point flags (valid1 / valid2, NULL = all), the acceptance test is `bestDist1 < TH_LOW`, and the output
 * is indexed by the FIRST keyframe: match12[pair][idx1] = feature of the second keyframe, -1 = none. */
int  orbx_search_by_bow_kf_device(orbx_vocabulary* v, int npairs, const int32_t* d_kf1_frame, const int32_t* d_kf2_frame,
                                  const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_valid1,
                                  const uint8_t* d_valid2, float nnratio, int check_orientation, int32_t* d_match12,
                                  int32_t* d_nmatches, void* cuda_stream);
int  orbx_search_by_bow_kf(orbx_vocabulary* v, const OrbxKeyPoint* kf1_keypoints, const uint8_t* kf1_descriptors, int n1,
                           const uint8_t* valid1, const OrbxKeyPoint* kf2_keypoints, const uint8_t* kf2_descriptors, int n2,
                           const uint8_t* valid2, int levelsup, float nnratio, int check_orientation, int32_t* match12,
                           int32_t* nmatches);
/* one pair from host buffers (transforms both descriptor sets first) */
int  orbx_search_by_bow(orbx_vocabulary* v, const OrbxKeyPoint* kf_keypoints, const uint8_t* kf_descriptors, int n_kf,
                        const uint8_t* kf_valid, const OrbxKeyPoint* f_keypoints, const uint8_t* f_descriptors, int n_f,
                        int levelsup, float nnratio, int check_orientation, int32_t* match_f, int32_t* nmatches);

/* ---- ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1489-1646),
 *      the matcher of Tracking::TrackWithMotionModel: projection of the last frame's map points with CurrentFrame.mTcw,
 *      Frame::GetFeaturesInArea on the 64x48 grid, nearest descriptor <= TH_HIGH, the "already holds an observed map
 *      point" skip in the reference's sequential order, rotation histogram + ComputeThreeMaxima.
 *      camera9 = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY); scale_factors = mvScaleFactors.
 *      OpenCV 4.13 f32 arithmetic for the cv::Mat expressions (see oracle/orb_oracle.c). ---- */
typedef struct OrbxProjectionPair {
    const OrbxKeyPoint* cur_keypoints;   /* CurrentFrame.mvKeysUn, n_cur */
    const uint8_t* cur_descriptors;      /* CurrentFrame.mDescriptors */
    const float* cur_u_right;            /* CurrentFrame.mvuRight, or NULL (monocular) */
    const uint8_t* cur_occupied;         /* != 0 <=> CurrentFrame.mvpMapPoints[i] has Observations() > 0 before the call; NULL = none */
    int32_t n_cur;
    const OrbxKeyPoint* last_keypoints;  /* LastFrame.mvKeysUn (octave and angle are used), n_last */
    const float* last_xyz;               /* pMP->GetWorldPos(): 3 floats per last-frame keypoint */
    const uint8_t* last_descriptors;     /* pMP->GetDescriptor(): 32 bytes per last-frame keypoint */
    const uint8_t* last_flags;           /* bit 0: has a map point and !mvbOutlier[i]; bit 1: pMP->Observations() > 0 */
    int32_t n_last;
    float Tcw[12];                       /* CurrentFrame.mTcw: Rcw row-major (9 floats), then tcw (3) */
    int32_t mode;                        /* 0: octaves nLastOctave-1..+1, 1: bForward, 2: bBackward (:1510-1511) */
    int32_t* match;                      /* out, n_cur: last-frame keypoint whose map point the keypoint now holds, -1 = none */
    int32_t* nmatches;                   /* out: the return value */
} OrbxProjectionPair;
/* host buffers, one pair, synchronous */
int orbx_search_by_projection(const OrbxProjectionPair* pair, const float* camera9, const float* scale_factors, int nlevels,
                              float th, int check_orientation, int device);
/* `pairs` is a HOST array whose pointers are DEVICE pointers; one CTA per pair, asynchronous on cuda_stream */
int orbx_search_by_projection_device(const OrbxProjectionPair* pairs, int npairs, const float* camera9,
                                     const float* scale_factors, int nlevels, float th, int check_orientation, int device,
                                     void* cuda_stream);

/* ---- ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (ORBmatcher.cc:46-142),
 *      the matcher of Tracking::SearchLocalPoints, complete: RadiusByViewingCos (:144-150), Frame::GetFeaturesInArea on the
 *      64x48 grid with the level range [level-1, level], best / second-best with their levels, `bestDist <= TH_HIGH`, the
 *      NN-ratio rule (:129-137) and the reference's SEQUENTIAL rule — a keypoint that holds a map point with
 *      Observations() > 0, from before the call or assigned earlier in the loop, is skipped (:90-92).
 *      One query per entry of vpMapPoints, carrying what Frame::isInFrustum (Frame.cc:315-384) left in the map point. ---- */
typedef struct OrbxTrackQuery {
    float proj_x, proj_y, proj_xr;      /* mTrackProjX, mTrackProjY, mTrackProjXR */
    float view_cos;                     /* mTrackViewCos */
    int32_t level;                      /* mnTrackScaleLevel */
} OrbxTrackQuery;
/* Frame::isInFrustum(MapPoint *pMP, float viewingCosLimit) (Frame.cc:315-378) for npts map points at once — what
 * Tracking::SearchLocalPoints (Tracking.cc:1409-1470) runs on every local map point before the matcher: Tcw12 = (mRcw
 * row-major, mtcw), Ow3 = mOw, camera9 = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY), pt_dist = 3 floats per
 * point (GetMinDistanceInvariance(), GetMaxDistanceInvariance(), mfMaxDistance). in_view[i] = the return value
 * (= mbTrackInView); queries[i] is written only where in_view[i] != 0 and feeds orbx_search_local_points directly. */
int orbx_is_in_frustum(const float* Tcw12, const float* Ow3, const float* camera9, int nlevels, float log_scale_factor,
                       const float* pt_xyz, const float* pt_normal, const float* pt_dist, int npts, float viewing_cos_limit,
                       OrbxTrackQuery* queries, uint8_t* in_view, int device);
/* device pointers for the point arrays and outputs; asynchronous on cuda_stream */
int orbx_is_in_frustum_device(const float* Tcw12, const float* Ow3, const float* camera9, int nlevels, float log_scale_factor,
                              const float* d_pt_xyz, const float* d_pt_normal, const float* d_pt_dist, int npts,
                              float viewing_cos_limit, OrbxTrackQuery* d_queries, uint8_t* d_in_view, int device, void* cuda_stream);
typedef struct OrbxLocalPointsFrame {
    const OrbxKeyPoint* keypoints;      /* F.mvKeysUn, n */
    const uint8_t* descriptors;         /* F.mDescriptors */
    const float* u_right;               /* F.mvuRight, or NULL (monocular) */
    const uint8_t* occupied;            /* != 0 <=> F.mvpMapPoints[i] has Observations() > 0 before the call; NULL = none */
    int32_t n;
    const OrbxTrackQuery* queries;      /* nq = vpMapPoints.size() */
    const uint8_t* query_descriptors;   /* pMP->GetDescriptor(), 32 bytes each */
    const uint8_t* query_flags;         /* bit 0: mbTrackInView && !isBad(); bit 1: Observations() > 0 */
    int32_t nq;
    int32_t* match;                     /* out, n: index into vpMapPoints the keypoint holds after the call, -1 = untouched */
    int32_t* nmatches;                  /* out: the return value */
} OrbxLocalPointsFrame;
/* bounds4 = (mnMinX, mnMaxX, mnMinY, mnMaxY); scale_factors = F.mvScaleFactors; nnratio = mfNNratio. Host buffers, synchronous */
int orbx_search_local_points(const OrbxLocalPointsFrame* frame, const float* bounds4, const float* scale_factors, int nlevels,
                             float th, float nnratio, int device);
/* `frames` is a HOST array whose pointers are DEVICE pointers; one CTA per frame, asynchronous on cuda_stream */
int orbx_search_local_points_device(const OrbxLocalPointsFrame* frames, int nframes, const float* bounds4,
                                    const float* scale_factors, int nlevels, float th, float nnratio, int device,
                                    void* cuda_stream);

/* ---- A device-resident Frame (Frame::Frame, Frame.cc:62-123: ExtractORB -> UndistortKeyPoints -> ComputeStereoMatches ->
 *      AssignFeaturesToGrid all work on the SAME keypoints). An orbx_frame keeps mvKeys, mDescriptors, mvKeysUn and mvuRight
 *      of one frame in HBM, taken straight from the extractor's device results, together with the scratch of the matchers:
 *      a matcher call uploads only the projected map points (one packed copy from a pinned staging buffer) and downloads
 *      the assignments (one copy); nothing is allocated on the call path. Every handle owns a CUDA stream, so calls on
 *      different handles — host threads, frames in flight — overlap on the GPU. A handle is not re-entrant. ---- */
typedef struct orbx_frame orbx_frame;
int  orbx_frame_create(int device, int max_keypoints, int max_queries, orbx_frame** out);
void orbx_frame_destroy(orbx_frame* f);
/* Frame `frame_index` (n keypoints, as reported in nkp) of the extractor's last HOST call (orbx_extract / orbx_extract_batch*):
 * device-to-device copy of keypoints + descriptors, then Frame::UndistortKeyPoints (Frame.cc:471-506) with K4 = (fx, fy, cx,
 * cy) and `ndist` distortion coefficients (0 = mvKeysUn = mvKeys). Ordered behind the extractor's stream, asynchronous. */
int  orbx_frame_from_extract(orbx_frame* f, orbx_extractor* ex, int frame_index, int n, const float* K4, const float* dist, int ndist);
/* the same from device arrays (e.g. the outputs of orbx_extract_device); producer_stream = the stream that wrote them */
int  orbx_frame_from_device(orbx_frame* f, const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, int n, const float* K4,
                            const float* dist, int ndist, void* producer_stream);
/* mvuRight (n floats; host array, or a device array when on_device != 0, e.g. the output of orbx_stereo_match_device); NULL = monocular */
int  orbx_frame_set_stereo(orbx_frame* f, const float* u_right, int on_device);
int  orbx_frame_size(const orbx_frame* f);
/* mvKeysUn / mDescriptors back on the host (either may be NULL) */
int  orbx_frame_keypoints(orbx_frame* f, OrbxKeyPoint* keypoints_un, uint8_t* descriptors, int cap, int* n);
/* the resident arrays and the handle's stream, for chaining the *_device matcher forms */
int  orbx_frame_device_arrays(orbx_frame* f, const OrbxKeyPoint** d_keypoints, const OrbxKeyPoint** d_keypoints_un,
                              const uint8_t** d_descriptors, const float** d_u_right, void** cuda_stream);
/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-142) against the resident frame:
 * arguments as orbx_search_local_points, host arrays; occupied (n bytes) may be NULL. Synchronous. */
int  orbx_frame_search_local_points(orbx_frame* f, const OrbxTrackQuery* queries, const uint8_t* query_descriptors,
                                    const uint8_t* query_flags, int nq, const uint8_t* occupied, const float* bounds4,
                                    const float* scale_factors, int nlevels, float th, float nnratio, int32_t* match,
                                    int32_t* nmatches);

/* ---- The search half of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th) (ORBmatcher.cc:918-1092;
 *      mode 0) and of Fuse(KeyFrame *pKF, cv::Mat Scw, vpPoints, th, vpReplacePoint) (:1094-1236; mode 1, with Rcw / tcw / Ow
 *      taken out of Scw by the caller as :1101-1106 do): projection into the keyframe, KeyFrame::IsInImage, the distance-
 *      invariance and viewing-angle gates, MapPoint::PredictScale (MapPoint.cc:407-422), KeyFrame::GetFeaturesInArea
 *      (KeyFrame.cc:708-747), the level and chi-square gates and the nearest descriptor. best_idx[i] = keyframe feature the
 *      point fuses with (bestDist <= TH_LOW) or -1; the Replace / AddObservation surgery that follows stays with the map. ---- */
typedef struct OrbxFuseJob {
    const OrbxKeyPoint* keypoints;      /* pKF->mvKeysUn, n */
    const uint8_t* descriptors;         /* pKF->mDescriptors */
    const float* u_right;               /* pKF->mvuRight, or NULL (all < 0) */
    int32_t n;
    float Tcw[12];                      /* Rcw row-major (9 floats), then tcw (3) */
    float Ow[3];                        /* camera centre */
    const float* pt_xyz;                /* GetWorldPos(): 3 floats per point */
    const float* pt_normal;             /* GetNormal(): 3 floats per point */
    const float* pt_dist;               /* 3 floats per point: GetMinDistanceInvariance(), GetMaxDistanceInvariance(), mfMaxDistance */
    const uint8_t* pt_descriptors;      /* GetDescriptor() */
    const uint8_t* pt_flags;            /* bit 0: non-NULL, !isBad() and not already in the keyframe (:941-945 / :1133-1136) */
    int32_t npts;
    float th;
    int32_t mode;                       /* 0: Fuse(pKF, vpMapPoints, th), 1: Fuse(pKF, Scw, ...) */
    int32_t* best_idx;                  /* out, npts */
    int32_t* best_dist;                 /* out, npts: bestDist (256 / INT_MAX when no candidate was compared) */
    int32_t* nfused;                    /* out: the return value */
} OrbxFuseJob;
/* camera9 = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY); scale_factors = mvScaleFactors, inv_level_sigma2 =
 * mvInvLevelSigma2, log_scale_factor = mfLogScaleFactor. Host buffers, one job, synchronous. */
int orbx_fuse_search(const OrbxFuseJob* job, const float* camera9, const float* scale_factors, const float* inv_level_sigma2,
                     int nlevels, float log_scale_factor, int device);
/* `jobs` is a HOST array whose pointers are DEVICE pointers; one CTA per job, asynchronous on cuda_stream */
int orbx_fuse_search_device(const OrbxFuseJob* jobs, int njobs, const float* camera9, const float* scale_factors,
                            const float* inv_level_sigma2, int nlevels, float log_scale_factor, int device, void* cuda_stream);

/* ---- ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:738-916), the matcher
 *      of LocalMapping::CreateNewMapPoints, for npairs keyframe pairs of the last transformed batch. has_mp[frame][cap] != 0
 *      <=> GetMapPoint(idx) is non-NULL (NULL = none has); u_right[frame][cap] = mvuRight (NULL = monocular keyframes).
 *      geom: 28 floats per pair — F12 row-major (9), pKF1->GetCameraCenter() (3), pKF2->GetRotation() row-major (9),
 *      pKF2->GetTranslation() (3), pKF2's fx, fy, cx, cy. scale_factors / level_sigma2 = pKF2->mvScaleFactors / mvLevelSigma2.
 *      match12[pair][idx1] = idx2 or -1: vMatchedPairs is the list of non-negative entries in ascending idx1.
 *      (The reference never sets vbMatched2, so a keyframe-2 feature may appear in several pairs; reproduced.) ---- */
int  orbx_search_for_triangulation_device(orbx_vocabulary* v, int npairs, const int32_t* d_kf1_frame, const int32_t* d_kf2_frame,
                                          const OrbxKeyPoint* d_keypoints, const uint8_t* d_descriptors, const uint8_t* d_has_mp,
                                          const float* d_u_right, const float* d_geom, const float* scale_factors,
                                          const float* level_sigma2, int nlevels, int only_stereo, int check_orientation,
                                          int32_t* d_match12, int32_t* d_nmatches, void* cuda_stream);
/* one pair from host buffers (transforms both descriptor sets first) */
int  orbx_search_for_triangulation(orbx_vocabulary* v, const OrbxKeyPoint* kf1_keypoints, const uint8_t* kf1_descriptors, int n1,
                                   const uint8_t* has_mp1, const float* u_right1, const OrbxKeyPoint* kf2_keypoints,
                                   const uint8_t* kf2_descriptors, int n2, const uint8_t* has_mp2, const float* u_right2,
                                   const float* geom28, const float* scale_factors, const float* level_sigma2, int nlevels,
                                   int levelsup, int only_stereo, int check_orientation, int32_t* match12, int32_t* nmatches);

/* ---- ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 *      (ORBmatcher.cc:1648-1795; mode 0, Tracking::Relocalization) and ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw,
 *      vpPoints, vpMatched, int th) (:327-440; mode 1, LoopClosing — Rcw / tcw / Ow taken out of Scw by the caller as :333-339
 *      do, max_dist = TH_LOW = 50). Map points are projected, gated (distance invariance, viewing angle in mode 1,
 *      MapPoint::PredictScale) and each takes the nearest FREE feature of its window in the reference's sequential order:
 *      a feature that holds a map point before the call (`occupied`) or received one earlier in the loop is skipped. ---- */
typedef struct OrbxProjectionJob {
    const OrbxKeyPoint* keypoints;      /* mvKeysUn of the frame (mode 0) / keyframe (mode 1), n */
    const uint8_t* descriptors;
    const uint8_t* occupied;            /* != 0 <=> CurrentFrame.mvpMapPoints[i] / vpMatched[i] is non-NULL before the call; NULL = none */
    int32_t n;
    float Tcw[12];                      /* Rcw row-major (9 floats), then tcw (3) */
    float Ow[3];                        /* -Rcw^T * tcw */
    const float* pt_xyz;                /* GetWorldPos(), 3 floats per point */
    const float* pt_normal;             /* GetNormal() (mode 1 only; may be NULL in mode 0) */
    const float* pt_dist;               /* 3 floats per point: GetMinDistanceInvariance(), GetMaxDistanceInvariance(), mfMaxDistance */
    const uint8_t* pt_descriptors;      /* GetDescriptor() */
    const uint8_t* pt_flags;            /* bit 0: non-NULL, !isBad(), not in sAlreadyFound / spAlreadyFound */
    const float* pt_angle;              /* mode 0 with check_orientation: pKF->mvKeysUn[i].angle */
    int32_t npts;
    float th;
    int32_t max_dist;                   /* ORBdist (mode 0) / TH_LOW (mode 1) */
    int32_t mode;
    int32_t* match;                     /* out, n: point index the feature holds after the call, -1 = untouched */
    int32_t* nmatches;                  /* out: the return value */
} OrbxProjectionJob;
int orbx_search_by_projection_kf(const OrbxProjectionJob* job, const float* camera9, const float* scale_factors, int nlevels,
                                 float log_scale_factor, int check_orientation, int device);
/* `jobs` is a HOST array whose pointers are DEVICE pointers; one CTA per job, asynchronous on cuda_stream */
int orbx_search_by_projection_kf_device(const OrbxProjectionJob* jobs, int njobs, const float* camera9, const float* scale_factors,
                                        int nlevels, float log_scale_factor, int check_orientation, int device, void* cuda_stream);

/* ---- ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (ORBmatcher.cc:1238-1487): the map points of each
 *      keyframe are carried into the other camera through the Sim3 and matched in a window (TH_HIGH); pairs found in both
 *      directions survive. Per keyframe: its features and, per feature, the map point it holds (mp_flags bit 0: the point
 *      exists, is not bad and the feature is not in vbAlreadyMatched). S12 = (sR12 row-major, t12), S21 = (sR21, t21) as
 *      :1253-1255 build them. match12[i1] = feature of keyframe 2 or -1. ---- */
typedef struct OrbxSim3KeyFrame {
    const OrbxKeyPoint* keypoints; const uint8_t* descriptors; int32_t n;
    const float* mp_xyz;                /* 3 floats per feature */
    const float* mp_dist;               /* 3 floats per feature (see OrbxProjectionJob.pt_dist) */
    const uint8_t* mp_descriptors;      /* 32 bytes per feature */
    const uint8_t* mp_flags;
    float Tcw[12];                      /* GetRotation() row-major, GetTranslation() */
} OrbxSim3KeyFrame;
int orbx_search_by_sim3(const OrbxSim3KeyFrame* kf1, const OrbxSim3KeyFrame* kf2, const float* S12, const float* S21,
                        const float* camera9, const float* scale_factors, int nlevels, float log_scale_factor, float th,
                        int32_t* match12, int32_t* nfound, int device);
/* device pointers inside kf1 / kf2, d_match12 (kf1->n ints) and d_nfound; d_scratch: (kf1->n + kf2->n + 2) ints */
int orbx_search_by_sim3_device(const OrbxSim3KeyFrame* kf1, const OrbxSim3KeyFrame* kf2, const float* S12, const float* S21,
                               const float* camera9, const float* scale_factors, int nlevels, float log_scale_factor, float th,
                               int32_t* d_match12, int32_t* d_nfound, int32_t* d_scratch, int device, void* cuda_stream);

/* ---- ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:442-587), the matcher
 *      of Tracking::MonocularInitialization: level-0 keypoints only, the sequential vMatchedDistance / vnMatches21 rule, NN
 *      ratio, rotation histogram. prev_matched = vbPrevMatched (2 floats per F1 keypoint), updated like :582-584. ---- */
typedef struct OrbxInitPair {
    const OrbxKeyPoint* keypoints1; const uint8_t* descriptors1; int32_t n1;   /* F1.mvKeysUn / mDescriptors */
    const OrbxKeyPoint* keypoints2; const uint8_t* descriptors2; int32_t n2;   /* F2 */
    const float* prev_matched;          /* in, 2 * n1 */
    float* prev_matched_out;            /* out, 2 * n1 (may alias prev_matched) */
    int32_t window_size;
    int32_t* match12;                   /* out, n1: vnMatches12 */
    int32_t* nmatches;                  /* out */
} OrbxInitPair;
int orbx_search_for_initialization(const OrbxInitPair* pair, const float* bounds4, float nnratio, int check_orientation, int device);
/* `pairs` is a HOST array whose pointers are DEVICE pointers; one CTA per pair, asynchronous on cuda_stream */
int orbx_search_for_initialization_device(const OrbxInitPair* pairs, int npairs, const float* bounds4, float nnratio,
                                          int check_orientation, int device, void* cuda_stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */

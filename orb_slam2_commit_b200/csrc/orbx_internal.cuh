// orbx_internal.cuh — shared declarations of the sm_100a ORB front-end kernels (not part of the C ABI).
//
// HBM layout of one extractor instance (sized by orbx_reserve for W x H frames, B frames per call):
//   blurred pyramid: a second block of the same geometry (orbx_blur.cu)
//   raw pyramid   [B] frames x frame_raw_bytes; level l of a frame starts at lvl[l].raw_off and is
//                 (h+38) rows of `pitch` bytes; the payload pixel (x,y) sits at row y+19, column x+ORBX_XOFF,
//                 so payload rows start 32-byte aligned and the 19-px REFLECT_101 apron of the reference's
//                 `temp` buffer (ORBextractor.cc:1225-1229) physically surrounds every level.
//   cell slots    [B][slot_total] u32  — NMS survivors of every 30-px FAST cell, packed (score<<24 | y<<12 | x),
//                 x,y relative to minBorder (16), row-major inside the cell
//   cell counts   [B][ncells] i32
//   candidates    [B][cand_total] u32 + u16 node ids — per level compacted in the reference's list order
//   level kps     [B][nlevels][kp_cap] u32 + counts — keypoints chosen by the quadtree, list order
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE 19          // EDGE_THRESHOLD (ORBextractor.cc:70)
#define ORBX_XOFF 32          // payload column origin inside a pyramid row (apron occupies columns 13..31)
#define ORBX_MINB 16          // minBorder = EDGE_THRESHOLD-3 (ORBextractor.cc:829)
#define ORBX_MAX_DIM 4096     // coordinates are packed in 12 bits
#define ORBX_QT_THREADS 1024
#define ORBX_MAX_QUOTA 2040   // per-level quadtree target N supported by the shared-memory node table

struct OrbxLevelGeom {
    int w, h;                 // payload size of the level
    int pitch;                // bytes per row of the level buffer
    int raw_off;              // byte offset of the level inside a frame's raw block (row 0 of the apron)
    int quota;                // mnFeaturesPerLevel[l]
    int nini;                 // quadtree roots (ORBextractor.cc:567)
    float hx;                 // root width (ORBextractor.cc:570)
    float scale;              // mvScaleFactor[l]
    float inv_scale;          // mvInvScaleFactor[l]
    float kp_size;            // (float)(int)(31*scale)
    int cell0, ncols, nrows;  // first cell id of the level, cell grid
    int cand_off;             // offset (elements) of the level's compact candidate array inside a frame
    int cand_cap;
    int kp_cap;               // capacity of the level's keypoint list
    int qt_depth;             // quadtree fast path: depth of the count pyramid (0 = always use the sweep path)
    int qt_xs_off, qt_ys_off; // per-coordinate halves of the path code (orbx_qt_path_tables), offsets into OrbxFrameLayout::qt_path
    int xtab_off, ytab_off;   // offsets into the resize coefficient tables (elements)
    int resize_fast;          // every group of four outputs keeps its x-taps within 8 source bytes (orbx_pyr_fast_ok)
    int pyr_tile_off, pyr_ntx, pyr_nty, pyr_box_w, pyr_box_h;   // resize tiles of this level: table offset (ints), grid, TMA box of level l-1
    int pyr_xg_off, pyr_yr_off;                                 // per column group / per buffer row thread tables (ints, orbx_pyr_tiles)
};

struct OrbxCell {             // one FAST cell (ORBextractor.cc:849-914); emission rectangle in payload coords
    short level;
    short ex0, ey0, ex1, ey1; // pixels that can be emitted: [ex0,ex1) x [ey0,ey1)
    short box_h;              // rows of the level's TMA box (tallest cell tile of the level = eh + 6)
    int slot_off;             // offset into the frame's slot array
    int slot_cap;
};

struct OrbxBlurUnit {         // one half-warp's share of the dense Gaussian blur (orbx_blur.cu): 64 columns x `rows` rows of a level
    short level;
    short c0, r0;             // buffer column (a multiple of 16) and buffer row of the source box; outputs start at (c0 + 4, r0 + 3)
    short rows;               // output rows (even, <= ORBX_BLUR_MAX_ROWS); the box has rows + 6
};
#define ORBX_BLUR_BOX_W 80    // 4 + 64 + 4 columns, rounded up to the TMA unit's 16-byte granule
#define ORBX_BLUR_MAX_ROWS 32

struct OrbxResizeTap {        // cv::resize INTER_LINEAR 8U coefficients for one destination coordinate
    short ofs, c0, c1, pad;
};

struct OrbxFrameLayout {      // everything the kernels need, passed by value
    int nlevels, ncells, slot_total, cand_total, kp_cap_total;
    int ini_th, min_th;
    int frame0;                     // working-set index of frame 0 of this launch (z coordinate of the tensor maps; `raw` etc. are already offset)
    size_t frame_raw_bytes;
    const OrbxLevelGeom* lvl;       // device
    const OrbxCell* cells;          // device
    const OrbxResizeTap* taps;      // device
    const int* pyr_tiles;           // device: source boxes of the resize tiles (orbx_pyr_tiles)
    const uint16_t* qt_path;        // device: quadtree path-code tables of every level (orbx_qt_path_tables)
    uint8_t* raw;                   // [B]
    uint8_t* blur;                  // [B] GaussianBlur of every level, same geometry as `raw` (read by the descriptor kernel only)
    uint32_t* slots;                // [B][slot_total]
    int* cell_count;                // [B][ncells]
    uint32_t* cand;                 // [B][cand_total]
    uint16_t* cand_node;            // [B][cand_total]
    int* cand_count;                // [B][nlevels]
    uint32_t* lvl_kp;               // [B][kp_cap_total]   (level l at lvl_kp_off[l])
    int* lvl_kp_count;              // [B][nlevels]
    int lvl_kp_off[ORBX_MAX_LEVELS];
    int qt_cap;                     // node-table capacity of the quadtree kernel
    int qt_hist_ints;               // shared-memory ints of the largest count pyramid
};

struct OrbxKp28 { float x, y, size, angle, response; int octave, class_id; };

// Opt-in for more than 48 KB of dynamic shared memory. The attribute belongs to the (function, device) pair and the
// library may serve several devices and host threads from one process, so the high-water mark is kept per device.
struct OrbxSmemMark { size_t bytes[64]; };
template <typename F>
static inline void orbx_need_smem(F kernel, OrbxSmemMark& mark, size_t bytes)
{
    if (bytes <= 48 * 1024) return;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (bytes > mark.bytes[dev]) {
        // on failure the mark stays where it was: the launch that follows fails with a launch error that the caller's
        // cudaGetLastError() check turns into ORBX_ERR_CUDA, and the next call tries again
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess) mark.bytes[dev] = bytes;
    }
}

// kernel launchers (implemented in the .cu files; all asynchronous on `st`)
struct OrbxTmaps;               // orbx_tma.cuh: one tensor map per pyramid level
void orbx_launch_pyramid(const OrbxFrameLayout& L, const OrbxTmaps& maps, const OrbxLevelGeom* h_lvl, const uint8_t* d_img, int w, int h,
                         int stride, size_t frame_pitch, int nframes, cudaStream_t st, int channels = 1, int rgb = 0,
                         const uint2* d_remap = nullptr, int src_w = 0, int src_h = 0);
bool orbx_pyr_fast_ok(const OrbxLevelGeom& g, const OrbxResizeTap* h_taps);
#include <vector>
void orbx_pyr_tiles(OrbxLevelGeom& g, const OrbxResizeTap* h_taps, std::vector<int>& out);
void orbx_launch_fast(const OrbxFrameLayout& L, const OrbxTmaps& maps, int max_tile_w, int max_tile_h, int nframes, cudaStream_t st);
int orbx_fast_tile_pitch(int max_tile_w);
void orbx_qt_path_tables(OrbxLevelGeom& g, std::vector<uint16_t>& out);
void orbx_launch_quadtree(const OrbxFrameLayout& L, int nframes, int threads, cudaStream_t st);
void orbx_blur_units(const OrbxLevelGeom& g, int level, std::vector<OrbxBlurUnit>& out, int* rows_per_unit);
void orbx_launch_blur(const OrbxFrameLayout& L, const OrbxTmaps& maps, const OrbxBlurUnit* d_units, int nunits, int nframes, cudaStream_t st);
void orbx_launch_describe(const OrbxFrameLayout& L, const OrbxTmaps& maps_raw, const OrbxTmaps& maps_blur, int nframes, OrbxKp28* d_kps,
                          uint8_t* d_desc, int cap, int* d_nkp, cudaStream_t st);
void orbx_upload_constants();  // pattern + umax tables

#define ORBX_MAX_PEERS 16
struct OrbxHtPeer {            // fused exchange of the per-rank top-2 over NVLink peer memory (no NCCL call)
    int world, rank, nq_max;
    unsigned long long* parts[ORBX_MAX_PEERS];   // peer p's landing buffer for the current epoch: [world][nq_max]
    unsigned* arrive[ORBX_MAX_PEERS];            // peer p's arrival counter
    unsigned* tile_done;                         // local: one counter per query tile
};
void orbx_launch_hamming_init(uint64_t* d_packed, int nq, cudaStream_t st);
void orbx_launch_hamming_top2_peer(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, long long index_base,
                                   uint64_t* d_packed, const OrbxHtPeer& peer, cudaStream_t st);
int orbx_hamming_qtiles(int nq);
void orbx_launch_hamming_wait_merge(const uint64_t* d_parts_local, const unsigned* d_arrive_local, unsigned target,
                                    int world, int nq, int nq_max, int* d_idx, int* d_d1, int* d_d2, int* d_status,
                                    cudaStream_t st);
void orbx_launch_hamming_top2(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, long long index_base,
                              uint64_t* d_packed, cudaStream_t st);
void orbx_launch_hamming_merge(const uint64_t* d_parts, int nparts, int nq, int* d_idx, int* d_d1, int* d_d2,
                               cudaStream_t st);
struct OrbxStereoBatch {       // full Frame::ComputeStereoMatches (Frame.cc:547-788) for `pairs` stereo pairs at once
    const OrbxKp28* kl; const uint8_t* dl; const int* nl;   // left  keypoints / descriptors / counts: [pairs][cap]
    const OrbxKp28* kr; const uint8_t* dr; const int* nr;   // right
    int cap, pairs, rows;
    const uint8_t* raw_left; const uint8_t* raw_right;      // HBM pyramids of the two extractors (frame p = pair p)
    size_t frame_raw_bytes;
    const OrbxLevelGeom* lvl;
    float minD, maxD, mbf;
    float max_scale;                                        // mvScaleFactors[nlevels - 1] (bounds the row band of a right keypoint)
    float* u_right; float* depth; int* sad;                 // [pairs][cap]; sad = -1 when unmatched
};
void orbx_launch_stereo_batch(const OrbxStereoBatch& a, cudaStream_t st);
struct OrbxWinQuery { float x, y, r; int min_level, max_level; float xr; };   // == OrbxWindowQuery of include/orbx.h
struct OrbxWindowArgs {        // windowed top-2 on the Frame grid (Frame.cc:388-444, ORBmatcher.cc:46-142)
    const OrbxKp28* kps; const uint8_t* desc; int n;
    const uint8_t* occupied; const float* u_right;          // optional (NULL)
    float minX, minY, invW, invH;
    const OrbxWinQuery* q; const uint8_t* qdesc; int nq;
    int *best_idx, *best_dist, *best_level, *best_dist2, *best_level2;
};
void orbx_launch_window_top2(const OrbxWindowArgs& a, cudaStream_t st);

// ---- bag of words (orbx_bow.cu). Vocabulary on the device: children of a node occupy consecutive slots.
struct OrbxVocabDev {
    int L, scoring, weighting, root_children;
    const uint4* slot_desc;      // [slots][2]   descriptor of the node in the slot
    const int2* slot_kids;       // [slots]      (first slot, count) of that node's own children
    const int* slot_node;        // [slots]      node id
    const double* weight;        // [nodes]      WordValue (0 for inner nodes)
    const int* word;             // [nodes]      word id, -1 for inner nodes
};
struct OrbxBowOut {              // per frame, `cap` entries each ([frames][cap], fv_off [frames][cap+1])
    int *leaf, *nid, *word;      // per feature: leaf node, node at level L - levelsup, word id
    int* bow_id; double* bow_val; int* n_bow;            // BowVector: ascending word id
    int *fv_node, *fv_off, *fv_feat, *n_fv;              // FeatureVector as CSR: ascending node id, features in order
};
struct OrbxBowMatchArgs {        // ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) for `pairs` (keyframe, frame) pairs
    const int *kf_frame, *f_frame;                       // [pairs] frame indices into the transformed batch
    const OrbxKp28* kps; const uint8_t* desc;            // the batch the transform ran on ([frames][cap])
    const uint8_t* kf_valid;                             // [pairs][cap] or NULL (all keyframe features have a good map point)
    const uint8_t* f_valid;                              // [pairs][cap] or NULL: second side's map-point flags (KeyFrame-KeyFrame form)
    float nnratio; int check_orientation, th_low;
    int kf_mode;                                         // 0: SearchByBoW(KeyFrame*, Frame&) ; 1: SearchByBoW(KeyFrame*, KeyFrame*)
    int *match, *bin_of, *taken;                         // [pairs][cap]; taken only in kf_mode
    int *hist, *nmatches;                                // [pairs][32], [pairs]
};
struct OrbxBowTriArgs {          // ORBmatcher::SearchForTriangulation for `pairs` (keyframe 1, keyframe 2) pairs
    const int *kf1_frame, *kf2_frame;                    // [pairs] frame indices into the transformed batch
    const OrbxKp28* kps; const uint8_t* desc;            // [frames][cap]
    const uint8_t* has_mp;                               // [frames][cap] or NULL: the feature already has a map point
    const float* u_right;                                // [frames][cap] or NULL (monocular keyframes)
    const float* geom;                                   // [pairs][28]: F12 (9), Cw1 (3), R2w (9), t2w (3), K2 = fx fy cx cy
    float scale_factors[ORBX_MAX_LEVELS], level_sigma2[ORBX_MAX_LEVELS];
    int only_stereo, check_orientation, th_low;
    int *match, *bin_of, *hist, *nmatches;               // [pairs][cap] x2, [pairs][32], [pairs]
};
void orbx_launch_bow_triangulation(const OrbxBowOut& O, const OrbxBowTriArgs& T, int* d_taken, const int* d_n, int cap, int npairs,
                                   cudaStream_t st);
void orbx_launch_bow_transform(const OrbxVocabDev& V, const uint8_t* d_desc, const int* d_n, int frames, int cap, int levelsup,
                               const OrbxBowOut& O, cudaStream_t st);
void orbx_launch_bow_score(const OrbxBowOut& O, int cap, const int* d_qa, const int* d_qb, int npairs, double* d_score, cudaStream_t st);
void orbx_launch_bow_match(const OrbxBowOut& O, const OrbxBowMatchArgs& A, const int* d_n, int cap, int npairs, cudaStream_t st);

// ---- ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (orbx_project.cu)
struct OrbxProjQuery { float u, v, r, ur; int min_level, max_level; };
struct OrbxProjCam { float fx, fy, cx, cy, mbf, minX, maxX, minY, maxY; };
struct OrbxProjPairDev {
    const OrbxKp28* cur_kps; const uint8_t* cur_desc; const float* cur_u_right; const uint8_t* cur_occupied; int n_cur;
    const OrbxKp28* last_kps; const float* last_xyz; const uint8_t* last_desc; const uint8_t* last_flags; int n_last;
    float Tcw[12]; int mode;
    int* match; int* nmatches;
    OrbxProjQuery* query; int* assign;            // scratch, n_last entries each
};
void orbx_launch_search_projection(const OrbxProjPairDev* d_pairs, int npairs, int max_n_cur, const OrbxProjCam& cam,
                                   const float* d_scale_factors, float th, int check_orientation, cudaStream_t st);

// ---- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) and ORBmatcher::Fuse (orbx_match.cu)
struct OrbxTrackQueryDev { float x, y, xr, view_cos; int level; };   // == OrbxTrackQuery of include/orbx.h
struct OrbxLocalFrameDev {
    const OrbxKp28* kps; const uint8_t* desc; const float* u_right; const uint8_t* occupied; int n;
    const OrbxTrackQueryDev* q; const uint8_t* qdesc; const uint8_t* qflags; int nq;
    int* match; int* nmatches;
    int* assign;                                  // scratch, nq entries
    int* result_out;                              // optional (NULL): the kernel also stores match[0..n) and the count at [n] here —
                                                  // pinned host memory the device can write (the frame handle's one-frame calls: no
                                                  // device-to-host copy call, and the driver's per-call lock is what bounds them)
};
void orbx_launch_local_points(const OrbxLocalFrameDev* d_frames, int nframes, int max_n, const float* bounds4,
                              const float* d_scale_factors, int nlevels, float th, float nnratio, cudaStream_t st);
struct OrbxFuseCam {
    float fx, fy, cx, cy, bf, minX, maxX, minY, maxY;
    int nlevels;
    float scale_factors[ORBX_MAX_LEVELS], inv_level_sigma2[ORBX_MAX_LEVELS];
    float level_ratio[ORBX_MAX_LEVELS];           // PredictScale thresholds: level = #{n < nlevels-1 : ratio >= level_ratio[n]}
};
struct OrbxFuseDev {
    const OrbxKp28* kps; const uint8_t* desc; const float* u_right; int n;
    float Tcw[12], T2[12], Ow[3], th; int mode;
    const float* pt_xyz; const float* pt_normal; const float* pt_dist; const uint8_t* pt_desc; const uint8_t* pt_flags; int npts;
    int* best_idx; int* best_dist; int* nfound;
};
void orbx_launch_fuse_search(const OrbxFuseDev* d_jobs, int njobs, int max_n, const OrbxFuseCam& cam, cudaStream_t st);

struct OrbxFrustumArgs {       // Frame::isInFrustum for npts map points
    float Tcw[12], Ow[3], view_cos_limit;
    const float* pt_xyz; const float* pt_normal; const float* pt_dist; int npts;
    OrbxTrackQueryDev* q; uint8_t* in_view;
};
void orbx_launch_in_frustum(const OrbxFrustumArgs& a, const OrbxFuseCam& cam, cudaStream_t st);
void orbx_launch_sim3_mutual(const int* d_match1, int n1, const int* d_match2, int n2, int* d_match12, int* d_nfound, cudaStream_t st);
struct OrbxSeqProjDev {        // SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (mode 0) / (pKF, Scw, vpPoints, vpMatched, th) (mode 1)
    const OrbxKp28* kps; const uint8_t* desc; const uint8_t* occupied; int n;
    float Tcw[12], Ow[3], th; int mode, th_dist, check_orientation;
    const float* pt_xyz; const float* pt_normal; const float* pt_dist; const uint8_t* pt_desc; const uint8_t* pt_flags;
    const float* pt_angle; int npts;
    int* match; int* nmatches;
    OrbxProjQuery* query; int* assign;            // scratch, npts entries each
};
void orbx_launch_seq_projection(const OrbxSeqProjDev* d_jobs, int njobs, int max_n, const OrbxFuseCam& cam, cudaStream_t st);
struct OrbxInitPairDev {       // ORBmatcher::SearchForInitialization
    const OrbxKp28* kps1; const uint8_t* desc1; int n1;
    const OrbxKp28* kps2; const uint8_t* desc2; int n2;
    const float* prev; float* prev_out;           // vbPrevMatched in / out: 2 floats per F1 keypoint
    int window;
    int* match12; int* nmatches;
    int* bin_of;                                  // scratch, n1 entries
    unsigned* top_key; int* top_idx; int* ncand;  // scratch: 4, 4 and 1 entries per F1 keypoint
};
void orbx_launch_init_match(const OrbxInitPairDev* d_pairs, int npairs, int max_n2, const float* bounds4, float nnratio,
                            int check_orientation, cudaStream_t st);

// cv::undistortPoints(src, dst, K, D, Mat(), K): intrinsics and (k1, k2, p1, p2, k3) widened to f64 on the host
struct OrbxUndistortArgs { double fx, fy, cx, cy, ifx, ify, k[5]; };
void orbx_launch_undistort(const OrbxKp28* d_in, OrbxKp28* d_out, int n, const OrbxUndistortArgs& a, cudaStream_t st);
// cv::initUndistortRectifyMap: ir = (P[:, :3] * R)^-1 row-major, k = (k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4)
struct OrbxRectifyArgs { double ir[9], k[12], fx, fy, u0, v0; };
void orbx_launch_rectify_map(const OrbxRectifyArgs& a, int w, int h, float* d_map1, float* d_map2, uint2* d_fixed, cudaStream_t st);
void orbx_launch_stereo_hamming(const OrbxKp28* kl, const uint8_t* dl, int nl, const OrbxKp28* kr, const uint8_t* dr,
                                int nr, const int* row_start, const int* row_tab, int rows, float minD, float maxD,
                                int* best_idx, int* best_dist, cudaStream_t st);

/*
 * orb_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's ORB front-end hot path
 * (qpc001/ORB_SLAM2_Commit: src/ORBextractor.cc, src/ORBmatcher.cc:1844-1860, src/Frame.cc:547-663)
 * with the OpenCV primitives it calls (resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 /
 * cvRound) restated from OpenCV 4.13's 8-bit arithmetic.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * Pinning: tests/golden/ (cv2 4.13 primitive outputs, tools/gen_golden.py) and oracle/_ref (the verbatim
 * reference ORBextractor.cc compiled against oracle/cvshim).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* identical to cv::KeyPoint (28 bytes) */
typedef struct { float x, y, size, angle, response; int octave, class_id; } OcKeyPoint;

/* ---- OpenCV primitive restatements ---- */
int   oc_round_f(float v);                      /* cvRound: round-half-even */
float oc_fast_atan2(float y, float x);          /* cv::fastAtan2 scalar path */
float oc_cosf(float x);                         /* glibc 2.39 x86_64 (FMA variant) cosf, restated */
float oc_sinf(float x);
void  oc_resize_linear_8u(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride);
void  oc_border_reflect101(uint8_t* whole, int w, int h, int stride, int border);
void  oc_gaussian7x7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
void  oc_cvt_gray(const uint8_t* src, int w, int h, int sstride, int channels, int rgb,
                  uint8_t* dst, int dstride);                 /* cv::cvtColor(.., CV_{RGB,BGR,RGBA,BGRA}2GRAY), 8U */
void  oc_remap_linear_8u(const uint8_t* src, int sw, int sh, int sstride, const float* map1, const float* map2,
                         int map_stride, int dw, int dh, uint8_t* dst, int dstride);  /* cv::remap INTER_LINEAR, BORDER_CONSTANT 0 */
void  oc_undistort_points(const float* xy, int n, const float* K4, const float* dist, int ndist,
                          float* out_xy);                      /* cv::undistortPoints(src,dst,K,D,Mat(),K), Frame.cc:471-538 */
void  oc_init_undistort_rectify_map(const double* K9, const double* D, int nD, const double* R9, const double* Ar9, int w, int h,
                                    float* map1, float* map2);   /* cv::initUndistortRectifyMap(.., CV_32F, ..), stereo_euroc.cc:96-97 */
int   oc_fast_score(const uint8_t* p, int stride);           /* cornerScore<16> with threshold floor 0 */
int   oc_fast9_16(const uint8_t* roi, int w, int h, int stride, int threshold, int nms,
                  OcKeyPoint* out, int cap);                 /* cv::FAST(roi, kps, threshold, nms) */

/* ---- reference functions ---- */
float oc_ic_angle(const uint8_t* center, int stride, const int* umax);          /* ORBextractor.cc:77-105 */
void  oc_orb_descriptor(const uint8_t* center, int stride, float angle_deg, uint8_t* desc32); /* :110-152 */
int   oc_distribute_octtree(const OcKeyPoint* in, int n, int minX, int maxX, int minY, int maxY,
                            int N, OcKeyPoint* out, int cap);                   /* :562-815 */
int   oc_descriptor_distance(const uint8_t* a, const uint8_t* b);               /* ORBmatcher.cc:1844-1860 */
void  oc_hamming_top2(const uint8_t* q, int nq, const uint8_t* t, int nt,
                      int32_t* idx1, int32_t* d1, int32_t* d2, int nthreads);   /* ORBmatcher.cc:84-126 idiom */

/* stereo row-band Hamming stage of Frame::ComputeStereoMatches (Frame.cc:554-663) */
void  oc_stereo_hamming(const OcKeyPoint* kl, const uint8_t* dl, int nl,
                        const OcKeyPoint* kr, const uint8_t* dr, int nr,
                        int rows, const float* scale_factors, float minD, float maxD,
                        int32_t* best_idx_r, int32_t* best_dist);

/* ---- extractor object (ORBextractor.cc:416-490, 1138-1250) ---- */
typedef struct OcExtractor OcExtractor;
OcExtractor* oc_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void  oc_destroy(OcExtractor*);
int   oc_levels(const OcExtractor*);
void  oc_tables(const OcExtractor*, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2,
                int* features_per_level, int* umax16);
/* returns number of keypoints (<= cap) or -1 on overflow; 0 for an empty image */
int   oc_extract(OcExtractor*, const uint8_t* img, int w, int h, int stride,
                 OcKeyPoint* kps, int cap, uint8_t* desc);
/* stage taps, valid after oc_extract */
int   oc_level_size(const OcExtractor*, int level, int* w, int* h, int* stride);
const uint8_t* oc_level_ptr(const OcExtractor*, int level);      /* payload origin (apron is around it) */
const uint8_t* oc_level_blur_ptr(const OcExtractor*, int level); /* blurred level, stride = w ; NULL if skipped */
int   oc_level_candidates(const OcExtractor*, int level, OcKeyPoint* out, int cap); /* pre-quadtree list */
int   oc_level_nkeypoints(const OcExtractor*, int level);

/* Full Frame::ComputeStereoMatches (Frame.cc:547-788) on the outputs and pyramids of two extractors that have just
 * processed the left / right image: row-band Hamming, 11x11 SAD slide +-5 on the level pyramids, parabola
 * sub-pixel, 1.5*1.4*median cut. mb = mbf / fx (Frame.cc:120). u_right / depth: n_left floats, -1 = no match. */
void  oc_stereo_match(const OcExtractor* left, const OcExtractor* right,
                      const OcKeyPoint* kl, const uint8_t* dl, int nl,
                      const OcKeyPoint* kr, const uint8_t* dr, int nr,
                      float mbf, float fx, float* u_right, float* depth);

/* Windowed best / second-best search of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)
 * (ORBmatcher.cc:46-142) over Frame::GetFeaturesInArea (Frame.cc:388-444) on the 64x48 grid built by
 * AssignFeaturesToGrid / PosInGrid (Frame.cc:254-271, 446-460). One query = one projected map point:
 * window centre (x,y), radius r (already multiplied by the level scale), level range, optional stereo coordinate xr
 * (< 0 = none). occupied[i] != 0 stands for "keypoint i already has an observed map point" (:90-92). */
typedef struct { float x, y, r; int32_t min_level, max_level; float xr; } OcWindowQuery;
void  oc_window_top2(const OcKeyPoint* kps, const uint8_t* desc, int n, const uint8_t* occupied, const float* u_right,
                     float minX, float minY, float invW, float invH,
                     const OcWindowQuery* q, const uint8_t* qdesc, int nq,
                     int32_t* best_idx, int32_t* best_dist, int32_t* best_level, int32_t* best_dist2, int32_t* best_level2);

/* ---- bag of words: DBoW2 as used through ORBVocabulary (Frame.cc:462-469, ORBmatcher.cc:175-325,
 *      KeyFrameDatabase.cc:145). DBoW2's .cpp files are absent from the reference snapshot; transform / score are pinned to the
 *      verbatim TemplatedVocabulary.h header by oracle/_ref/libbow_ref.so (oracle/bow_glue.cc, tests/test_matcher_ref.py). ----
 * Vocabulary nodes in text-file order (TemplatedVocabulary.h:1338-1420): node i+1 has parent[i] (0 = root),
 * is_leaf[i], a 32-byte descriptor and a weight; words are numbered in order of appearance.
 * scoring: 0 L1_NORM 1 L2_NORM 2 CHI_SQUARE 3 KL 4 BHATTACHARYYA 5 DOT_PRODUCT; weighting: 0 TF_IDF 1 TF 2 IDF 3 BINARY. */
typedef struct OcVocabulary OcVocabulary;
OcVocabulary* oc_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                              const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
void  oc_vocab_destroy(OcVocabulary*);
int   oc_vocab_words(const OcVocabulary*);
void  oc_vocab_transform(const OcVocabulary* v, const uint8_t* desc, int n, int levelsup,
                         int32_t* word, int32_t* node, int32_t* bow_id, double* bow_val, int32_t* n_bow,
                         int32_t* fv_node, int32_t* fv_off, int32_t* fv_feat, int32_t* n_fv);
double oc_bow_score_l1(const int32_t* id1, const double* v1, int n1, const int32_t* id2, const double* v2, int n2);
int   oc_search_by_bow(const int32_t* kf_fv_node, const int32_t* kf_fv_off, const int32_t* kf_fv_feat, int kf_nfv,
                       const int32_t* f_fv_node, const int32_t* f_fv_off, const int32_t* f_fv_feat, int f_nfv,
                       const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                       const uint8_t* f_desc, const float* f_angle, int f_n,
                       float nnratio, int check_orientation, int32_t* match_f);

int   oc_search_by_bow_kf(const int32_t* fv1_node, const int32_t* fv1_off, const int32_t* fv1_feat, int nfv1,
                          const int32_t* fv2_node, const int32_t* fv2_off, const int32_t* fv2_feat, int nfv2,
                          const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                          const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2,
                          float nnratio, int check_orientation, int32_t* match12);   /* ORBmatcher.cc:589-736 */

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1489-1646);
 * argument meaning as include/orbx.h OrbxProjectionPair. Direct restatement, cv::Mat arithmetic pinned to cv2 4.13's
 * gemm; pinned to the unmodified ORBmatcher.cc by oracle/_ref/libmatcher_ref.so (tests/test_matcher_ref.py). Returns nmatches. */
int   oc_search_by_projection_frame(const OcKeyPoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* cur_u_right,
                                    const uint8_t* cur_occupied, const float* Tcw12, const float* cam9,
                                    const float* scale_factors,
                                    const OcKeyPoint* last_kps, const float* last_xyz, const uint8_t* last_desc,
                                    const uint8_t* last_flags, int n_last, float th, int mode, int check_orientation,
                                    int32_t* match_cur);

/* ---- the remaining window / BoW matchers of ORBmatcher; see the definitions in orb_oracle.c for argument meaning.
 *      Direct restatements, cv::Mat arithmetic as pinned for oc_search_by_projection_frame; pinned to the UNMODIFIED
 *      ORBmatcher.cc by oracle/_ref/libmatcher_ref.so (oracle/matcher_glue.cc, tests/test_matcher_ref.py). ---- */
typedef struct { float x, y, xr, view_cos; int32_t level; } OcTrackQuery;
int   oc_search_local_points(const OcKeyPoint* kps, const uint8_t* desc, int n, const float* u_right, const uint8_t* occupied,
                             const float* bounds4, const float* scale_factors, int nlevels,
                             const OcTrackQuery* q, const uint8_t* qdesc, const uint8_t* qflags, int nq,
                             float th, float nnratio, int32_t* match);          /* ORBmatcher.cc:46-142 */
int   oc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int nlevels);   /* MapPoint.cc:407-422 */
int   oc_fuse_search(const OcKeyPoint* kps, const uint8_t* desc, int n, const float* u_right,
                     const float* Tcw12, const float* Ow3, const float* cam9, const float* scale_factors,
                     const float* inv_level_sigma2, int nlevels, float log_scale_factor,
                     const float* pt_xyz, const float* pt_normal, const float* pt_dist, const uint8_t* pt_desc,
                     const uint8_t* pt_flags, int npts, float th, int mode, int32_t* best_idx, int32_t* best_dist);
                                                                                /* ORBmatcher.cc:918-1092, 1094-1236 */
int   oc_search_for_triangulation(const int32_t* fv1_node, const int32_t* fv1_off, const int32_t* fv1_feat, int nfv1,
                                  const int32_t* fv2_node, const int32_t* fv2_off, const int32_t* fv2_feat, int nfv2,
                                  const OcKeyPoint* kps1, const uint8_t* desc1, const uint8_t* skip1, const float* u_right1, int n1,
                                  const OcKeyPoint* kps2, const uint8_t* desc2, const uint8_t* skip2, const float* u_right2, int n2,
                                  const float* F12, const float* Cw1, const float* pose2, const float* K2,
                                  const float* scale_factors2, const float* level_sigma2_2,
                                  int only_stereo, int check_orientation, int32_t* match12);   /* ORBmatcher.cc:738-916 */

int   oc_search_by_projection_seq(const OcKeyPoint* kps, const uint8_t* desc, int n, const uint8_t* occupied,
                                  const float* Tcw12, const float* Ow3, const float* cam9, const float* scale_factors,
                                  int nlevels, float log_scale_factor,
                                  const float* pt_xyz, const float* pt_normal, const float* pt_dist, const uint8_t* pt_desc,
                                  const uint8_t* pt_flags, const float* pt_angle, int npts,
                                  float th, int th_dist, int mode, int check_orientation, int32_t* match);
                                                                                /* ORBmatcher.cc:1648-1795 (mode 0), 327-440 (mode 1) */
int   oc_search_by_sim3(const OcKeyPoint* kps1, const uint8_t* desc1, int n1, const float* xyz1, const float* dist1,
                        const uint8_t* mpdesc1, const uint8_t* flags1,
                        const OcKeyPoint* kps2, const uint8_t* desc2, int n2, const float* xyz2, const float* dist2,
                        const uint8_t* mpdesc2, const uint8_t* flags2,
                        const float* T1w, const float* T2w, const float* S12, const float* S21,
                        const float* cam9, const float* scale_factors, int nlevels, float log_scale_factor, float th,
                        int32_t* match12);                                      /* ORBmatcher.cc:1238-1487 */
int   oc_search_for_initialization(const OcKeyPoint* kps1, const uint8_t* desc1, int n1,
                                   const OcKeyPoint* kps2, const uint8_t* desc2, int n2, const float* bounds4,
                                   float* prev, int window, float nnratio, int check_orientation, int32_t* match12);
                                                                                /* ORBmatcher.cc:442-587 */

void  oc_is_in_frustum(const float* Tcw12, const float* Ow3, const float* cam9, int nlevels, float log_scale_factor,
                       const float* pt_xyz, const float* pt_normal, const float* pt_dist, int npts, float viewingCosLimit,
                       OcTrackQuery* q, uint8_t* in_view);                      /* Frame.cc:315-378 */

#ifdef __cplusplus
}
#endif
#endif

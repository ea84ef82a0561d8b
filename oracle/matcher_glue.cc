/*
 * oracle/matcher_glue.cc — C entry points around the UNMODIFIED /root/reference/src/ORBmatcher.cc (oracle/_ref build,
 * TEST INFRASTRUCTURE ONLY). The build recipe (oracle/Makefile, target ref_matcher) compiles ORBmatcher.cc where it lies
 * against oracle/slamshim (a small cv::Mat and stub Frame / KeyFrame / MapPoint classes) and takes the few member
 * functions the matcher calls — Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid (and Frame::ComputeStereoMatches), KeyFrame::GetFeaturesInArea /
 * IsInImage, MapPoint::PredictScale / Get{Min,Max}DistanceInvariance — verbatim from the reference sources by line range
 * into oracle/_ref/gen/*.inc. Each mref_* function takes the same flattened arrays as the oc_* restatement in
 * orb_oracle.c, builds the object graph the reference function expects, runs it, and flattens the result back, so that
 * tests/test_matcher_ref.py can compare restatement and reference on the same scenes.
 * Where the reference derives a quantity from its arguments with cv::Mat arithmetic (Ow from mTcw, Rcw / tcw from Scw,
 * sR21 / t21 from s12, R12, t12) the glue derives it with the same shim operators and hands it back, and the test feeds
 * exactly those values to the restatement.
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <vector>
using namespace std;                       /* as DBoW2/TemplatedVocabulary.h:36 does for every reference translation unit */
#include "ORBmatcher.h"

namespace DBoW2 {                          /* DBoW2's .cpp files are absent from the reference snapshot */
FeatureVector::FeatureVector(void) {}
FeatureVector::~FeatureVector(void) {}
}

namespace ORB_SLAM2 {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
#include "gen/frame_assign.inc"
#include "gen/frame_area.inc"
#include "gen/frame_posingrid.inc"
#include "gen/frame_frustum.inc"
#include "gen/frame_stereo.inc"
#include "gen/keyframe_area.inc"
#include "gen/mappoint_dist.inc"
void KeyFrame::AddMapPoint(MapPoint* pMP, const size_t& idx) { pMP->fused = (int)idx; }   /* records Fuse's pick, performs nothing */
}
using namespace ORB_SLAM2;

namespace {
struct KP28 { float x, y, size, angle, response; int32_t octave, class_id; };

cv::Mat vec3(const float* p) { cv::Mat m(3, 1, CV_32F); for (int i = 0; i < 3; i++) m.at<float>(i) = p[i]; return m; }
cv::Mat mat33(const float* p) { cv::Mat m(3, 3, CV_32F); for (int i = 0; i < 9; i++) m.at<float>(i / 3, i % 3) = p[i]; return m; }
cv::Mat desc_rows(const uint8_t* d, int n)
{
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    for (int i = 0; i < n; i++) memcpy(m.ptr<uint8_t>(i), d + 32 * (size_t)i, 32);
    return m;
}
void keys(const KP28* k, int n, vector<cv::KeyPoint>& out)
{
    out.resize((size_t)n);
    if (n) memcpy(static_cast<void*>(out.data()), k, (size_t)n * 28);
}
void set_frame_statics(const float* cam9)
{
    Frame::fx = cam9[0]; Frame::fy = cam9[1]; Frame::cx = cam9[2]; Frame::cy = cam9[3];
    Frame::mnMinX = cam9[5]; Frame::mnMaxX = cam9[6]; Frame::mnMinY = cam9[7]; Frame::mnMaxY = cam9[8];
    /* Frame.cc:66-67 */
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}
void fill_frame(Frame& F, const KP28* kps, const uint8_t* desc, int n, const float* u_right, const float* cam9,
                const float* scale_factors, int nlevels, float log_sf)
{
    set_frame_statics(cam9);
    F.N = n; keys(kps, n, F.mvKeysUn); F.mvKeys = F.mvKeysUn; F.mDescriptors = desc_rows(desc, n);
    F.mvuRight.assign((size_t)n, -1.f);
    if (u_right) F.mvuRight.assign(u_right, u_right + n);
    F.mvpMapPoints.assign((size_t)n, static_cast<MapPoint*>(NULL)); F.mvbOutlier.assign((size_t)n, false);
    F.mbf = cam9[4]; F.mb = cam9[4] / cam9[0];
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels); F.mnScaleLevels = nlevels; F.mfLogScaleFactor = log_sf;
    F.AssignFeaturesToGrid();
}
void fill_keyframe(KeyFrame& K, const KP28* kps, const uint8_t* desc, int n, const float* u_right, const float* cam9,
                   const float* scale_factors, const float* level_sigma2, const float* inv_level_sigma2, int nlevels, float log_sf)
{
    Frame F;                                                        /* KeyFrame::KeyFrame(Frame &F, ...) copies the Frame's grid */
    fill_frame(F, kps, desc, n, u_right, cam9, scale_factors, nlevels, log_sf);
    K.N = n; K.mvKeysUn = F.mvKeysUn; K.mDescriptors = F.mDescriptors; K.mvuRight = F.mvuRight;
    K.fx = cam9[0]; K.fy = cam9[1]; K.cx = cam9[2]; K.cy = cam9[3]; K.mbf = cam9[4];
    K.mnMinX = F.mnMinX; K.mnMinY = F.mnMinY; K.mnMaxX = F.mnMaxX; K.mnMaxY = F.mnMaxY;   /* float -> const int (KeyFrame.cc:41-42) */
    K.mfGridElementWidthInv = F.mfGridElementWidthInv; K.mfGridElementHeightInv = F.mfGridElementHeightInv;
    K.mvScaleFactors = F.mvScaleFactors; K.mnScaleLevels = nlevels; K.mfLogScaleFactor = log_sf;
    if (level_sigma2) K.mvLevelSigma2.assign(level_sigma2, level_sigma2 + nlevels);
    if (inv_level_sigma2) K.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
    K.mvpMapPoints.assign((size_t)n, static_cast<MapPoint*>(NULL));
    K.mGrid.resize(K.mnGridCols);                                    /* KeyFrame.cc:52-57 */
    for (int i = 0; i < K.mnGridCols; i++) {
        K.mGrid[i].resize(K.mnGridRows);
        for (int j = 0; j < K.mnGridRows; j++) K.mGrid[i][j] = F.mGrid[i][j];
    }
}
struct Points {
    vector<unique_ptr<MapPoint> > own;
    vector<MapPoint*> v;
    MapPoint* add()
    {
        own.push_back(unique_ptr<MapPoint>(new MapPoint()));
        own.back()->index = (int)v.size(); v.push_back(own.back().get());
        return v.back();
    }
};
/* pt_dist rows are (GetMinDistanceInvariance, GetMaxDistanceInvariance, mfMaxDistance); the stub needs mfMinDistance with
 * 0.8f * mfMinDistance == row[0] and 1.2f * mfMaxDistance == row[1]: the test scenes are generated so that both hold. */
void fill_point(MapPoint* p, const float* xyz, const float* normal, const float* dist3, const uint8_t* desc)
{
    if (xyz) p->mWorldPos = vec3(xyz);
    if (normal) p->mNormalVector = vec3(normal);
    if (desc) p->mDescriptor = desc_rows(desc, 1);
    if (dist3) { p->mfMaxDistance = dist3[2]; p->mfMinDistance = dist3[3]; }
}
void featvec(DBoW2::FeatureVector& fv, const int32_t* node, const int32_t* off, const int32_t* feat, int nfv)
{
    for (int a = 0; a < nfv; a++) {
        vector<unsigned int>& v = fv[(unsigned int)node[a]];
        for (int i = off[a]; i < off[a + 1]; i++) v.push_back((unsigned int)feat[i]);
    }
}
MapPoint g_dummy_observed;       /* stands for "this feature already holds a map point with observations" */
}

extern "C" {

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) */
int mref_search_local_points(const KP28* kps, const uint8_t* desc, int n, const float* u_right, const uint8_t* occupied,
                             const float* bounds4, const float* scale_factors, int nlevels,
                             const float* q5 /* proj_x, proj_y, proj_xr, view_cos per point */, const int32_t* qlevel,
                             const uint8_t* qdesc, const uint8_t* qflags, int nq, float th, float nnratio, int32_t* match)
{
    const float cam9[9] = {1, 1, 0, 0, 1, bounds4[0], bounds4[1], bounds4[2], bounds4[3]};
    Frame F;
    fill_frame(F, kps, desc, n, u_right, cam9, scale_factors, nlevels, 0.f);
    g_dummy_observed.nObs = 1; g_dummy_observed.index = -1;
    for (int i = 0; i < n; i++) if (occupied && occupied[i]) F.mvpMapPoints[i] = &g_dummy_observed;
    Points P;
    for (int i = 0; i < nq; i++) {
        MapPoint* p = P.add();
        p->mTrackProjX = q5[4 * i]; p->mTrackProjY = q5[4 * i + 1]; p->mTrackProjXR = q5[4 * i + 2]; p->mTrackViewCos = q5[4 * i + 3];
        p->mnTrackScaleLevel = qlevel[i]; p->mbTrackInView = (qflags[i] & 1) != 0; p->nObs = (qflags[i] & 2) ? 1 : 0;
        fill_point(p, 0, 0, 0, qdesc + 32 * (size_t)i);
    }
    ORBmatcher matcher(nnratio, true);
    const int r = matcher.SearchByProjection(F, P.v, th);
    for (int i = 0; i < n; i++) match[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &g_dummy_observed) ? F.mvpMapPoints[i]->index : -1;
    return r;
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono).
 * tlw_z steers bForward / bBackward: LastFrame.mTcw = [I | (0, 0, tlw_z)]. */
int mref_search_by_projection_frame(const KP28* cur_kps, const uint8_t* cur_desc, int n_cur, const float* cur_u_right,
                                    const uint8_t* cur_occupied, const float* Tcw12, const float* cam9, const float* scale_factors,
                                    int nlevels, const KP28* last_kps, const float* last_xyz, const uint8_t* last_desc,
                                    const uint8_t* last_flags, int n_last, float th, int mono, float tlw_z, int check_orientation,
                                    int32_t* match_cur, int32_t* mode_out)
{
    Frame C, L;
    fill_frame(C, cur_kps, cur_desc, n_cur, cur_u_right, cam9, scale_factors, nlevels, 0.f);
    fill_frame(L, last_kps, last_desc, n_last, 0, cam9, scale_factors, nlevels, 0.f);
    C.mTcw = cv::Mat(4, 4, CV_32F); L.mTcw = cv::Mat(4, 4, CV_32F);
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { C.mTcw.at<float>(r, c) = Tcw12[3 * r + c]; L.mTcw.at<float>(r, c) = r == c; } C.mTcw.at<float>(r, 3) = Tcw12[9 + r]; }
    C.mTcw.at<float>(3, 3) = 1; L.mTcw.at<float>(3, 3) = 1; L.mTcw.at<float>(2, 3) = tlw_z;
    g_dummy_observed.nObs = 1; g_dummy_observed.index = -1;
    for (int i = 0; i < n_cur; i++) if (cur_occupied && cur_occupied[i]) C.mvpMapPoints[i] = &g_dummy_observed;
    Points P;
    for (int i = 0; i < n_last; i++) {
        MapPoint* p = P.add();
        fill_point(p, last_xyz + 3 * (size_t)i, 0, 0, last_desc + 32 * (size_t)i);
        p->nObs = (last_flags[i] & 2) ? 1 : 0;
        L.mvpMapPoints[i] = p; L.mvbOutlier[i] = !(last_flags[i] & 1);
    }
    /* the mode the reference will pick (:1505-1511), recomputed with the same operators for the restatement */
    const cv::Mat Rcw = C.mTcw.rowRange(0, 3).colRange(0, 3), tcw = C.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat tlc = L.mTcw.rowRange(0, 3).colRange(0, 3) * twc + L.mTcw.rowRange(0, 3).col(3);
    *mode_out = (tlc.at<float>(2) > C.mb && !mono) ? 1 : (-tlc.at<float>(2) > C.mb && !mono) ? 2 : 0;
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int r = matcher.SearchByProjection(C, L, th, mono != 0);
    for (int i = 0; i < n_cur; i++) match_cur[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &g_dummy_observed) ? C.mvpMapPoints[i]->index : -1;
    return r;
}

/* Fuse(pKF, vpMapPoints, th) (mode 0; Tcw12 / Ow3 are the keyframe's pose) and Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)
 * (mode 1; Scw12 = the 3x4 top of Scw, derived Rcw / tcw / Ow returned in Tcw12_out / Ow3_out). pt_dist4 rows: min-invariance,
 * max-invariance, mfMaxDistance, mfMinDistance. best_idx[i] = the feature the point was added to (all keyframe features are
 * free, so every fused point takes the AddObservation / AddMapPoint branch, which the stub records in MapPoint::fused). */
int mref_fuse(const KP28* kps, const uint8_t* desc, int n, const float* u_right, const float* T12, const float* Ow3,
              const float* cam9, const float* scale_factors, const float* inv_level_sigma2, int nlevels, float log_sf,
              const float* pt_xyz, const float* pt_normal, const float* pt_dist4, const uint8_t* pt_desc, const uint8_t* pt_flags,
              int npts, float th, int mode, int32_t* best_idx, float* Tcw12_out, float* Ow3_out)
{
    KeyFrame K;
    fill_keyframe(K, kps, desc, n, u_right, cam9, scale_factors, 0, inv_level_sigma2, nlevels, log_sf);
    Points P;
    for (int i = 0; i < npts; i++) {
        MapPoint* p = P.add();
        fill_point(p, pt_xyz + 3 * (size_t)i, pt_normal + 3 * (size_t)i, pt_dist4 + 4 * (size_t)i, pt_desc + 32 * (size_t)i);
        p->mbBad = !(pt_flags[i] & 1); p->nObs = 1;
    }
    ORBmatcher matcher(0.8f, true);
    int r;
    if (mode == 0) {
        K.Rcw = mat33(T12); K.tcw = vec3(T12 + 9); K.Ow = vec3(Ow3);
        memcpy(Tcw12_out, T12, 48); memcpy(Ow3_out, Ow3, 12);
        r = matcher.Fuse(&K, P.v, th);
    } else {
        cv::Mat Scw(4, 4, CV_32F);
        for (int rr = 0; rr < 3; rr++) { for (int c = 0; c < 3; c++) Scw.at<float>(rr, c) = T12[3 * rr + c]; Scw.at<float>(rr, 3) = T12[9 + rr]; }
        Scw.at<float>(3, 3) = 1;
        cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);                        /* :1101-1106, same operators */
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        cv::Mat Rcw = sRcw / scw, tcw = Scw.rowRange(0, 3).col(3) / scw, Ow = -Rcw.t() * tcw;
        for (int rr = 0; rr < 3; rr++) { for (int c = 0; c < 3; c++) Tcw12_out[3 * rr + c] = Rcw.at<float>(rr, c); Tcw12_out[9 + rr] = tcw.at<float>(rr); Ow3_out[rr] = Ow.at<float>(rr); }
        vector<MapPoint*> vpReplace((size_t)npts, static_cast<MapPoint*>(NULL));
        r = matcher.Fuse(&K, Scw, P.v, th, vpReplace);
    }
    for (int i = 0; i < npts; i++) best_idx[i] = P.v[i]->fused;
    return r;
}

/* SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (mode 0) / SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)
 * (mode 1; T12 = top of Scw). Derived Tcw / Ow returned for the restatement. */
int mref_search_by_projection_seq(const KP28* kps, const uint8_t* desc, int n, const uint8_t* occupied, const float* T12,
                                  const float* cam9, const float* scale_factors, int nlevels, float log_sf,
                                  const float* pt_xyz, const float* pt_normal, const float* pt_dist4, const uint8_t* pt_desc,
                                  const uint8_t* pt_flags, const float* pt_angle, int npts, float th, int th_dist, int mode,
                                  int check_orientation, int32_t* match, float* Tcw12_out, float* Ow3_out)
{
    Points P;
    for (int i = 0; i < npts; i++) {
        MapPoint* p = P.add();
        fill_point(p, pt_xyz + 3 * (size_t)i, pt_normal ? pt_normal + 3 * (size_t)i : 0, pt_dist4 + 4 * (size_t)i, pt_desc + 32 * (size_t)i);
        p->mbBad = !(pt_flags[i] & 1); p->nObs = 1;
    }
    g_dummy_observed.nObs = 1; g_dummy_observed.index = -1;
    ORBmatcher matcher(0.9f, check_orientation != 0);
    int r;
    if (mode == 0) {
        Frame C;
        fill_frame(C, kps, desc, n, 0, cam9, scale_factors, nlevels, log_sf);
        C.mTcw = cv::Mat(4, 4, CV_32F);
        for (int rr = 0; rr < 3; rr++) { for (int c = 0; c < 3; c++) C.mTcw.at<float>(rr, c) = T12[3 * rr + c]; C.mTcw.at<float>(rr, 3) = T12[9 + rr]; }
        C.mTcw.at<float>(3, 3) = 1;
        for (int i = 0; i < n; i++) if (occupied && occupied[i]) C.mvpMapPoints[i] = &g_dummy_observed;
        const cv::Mat Rcw = C.mTcw.rowRange(0, 3).colRange(0, 3), tcw = C.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -Rcw.t() * tcw;                                          /* :1655 */
        memcpy(Tcw12_out, T12, 48);
        for (int rr = 0; rr < 3; rr++) Ow3_out[rr] = Ow.at<float>(rr);
        KeyFrame K;                                                                 /* only its map points and angles are read */
        K.mvpMapPoints = P.v; K.mvKeysUn.resize((size_t)npts);
        for (int i = 0; i < npts; i++) K.mvKeysUn[i].angle = pt_angle ? pt_angle[i] : 0.f;
        set<MapPoint*> found;
        r = matcher.SearchByProjection(C, &K, found, th, th_dist);
        for (int i = 0; i < n; i++) match[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &g_dummy_observed) ? C.mvpMapPoints[i]->index : -1;
    } else {
        KeyFrame K;
        fill_keyframe(K, kps, desc, n, 0, cam9, scale_factors, 0, 0, nlevels, log_sf);
        cv::Mat Scw(4, 4, CV_32F);
        for (int rr = 0; rr < 3; rr++) { for (int c = 0; c < 3; c++) Scw.at<float>(rr, c) = T12[3 * rr + c]; Scw.at<float>(rr, 3) = T12[9 + rr]; }
        Scw.at<float>(3, 3) = 1;
        cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);                            /* :333-339 */
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        cv::Mat Rcw = sRcw / scw, tcw = Scw.rowRange(0, 3).col(3) / scw, Ow = -Rcw.t() * tcw;
        for (int rr = 0; rr < 3; rr++) { for (int c = 0; c < 3; c++) Tcw12_out[3 * rr + c] = Rcw.at<float>(rr, c); Tcw12_out[9 + rr] = tcw.at<float>(rr); Ow3_out[rr] = Ow.at<float>(rr); }
        vector<MapPoint*> vpMatched((size_t)n, static_cast<MapPoint*>(NULL));
        for (int i = 0; i < n; i++) if (occupied && occupied[i]) vpMatched[i] = &g_dummy_observed;
        r = matcher.SearchByProjection(&K, Scw, P.v, vpMatched, (int)th);
        for (int i = 0; i < n; i++) match[i] = (vpMatched[i] && vpMatched[i] != &g_dummy_observed) ? vpMatched[i]->index : -1;
    }
    return r;
}

/* SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th); S12_out / S21_out = (sR12, t12) / (sR21, t21) as :1253-1255 derive them */
int mref_search_by_sim3(const KP28* kps1, const uint8_t* desc1, int n1, const float* xyz1, const float* dist1_4, const uint8_t* mpdesc1, const uint8_t* flags1,
                        const KP28* kps2, const uint8_t* desc2, int n2, const float* xyz2, const float* dist2_4, const uint8_t* mpdesc2, const uint8_t* flags2,
                        const float* T1w, const float* T2w, float s12, const float* R12_9, const float* t12_3,
                        const float* cam9, const float* scale_factors, int nlevels, float log_sf, float th,
                        int32_t* match12, float* S12_out, float* S21_out)
{
    KeyFrame K1, K2;
    fill_keyframe(K1, kps1, desc1, n1, 0, cam9, scale_factors, 0, 0, nlevels, log_sf);
    fill_keyframe(K2, kps2, desc2, n2, 0, cam9, scale_factors, 0, 0, nlevels, log_sf);
    K1.Rcw = mat33(T1w); K1.tcw = vec3(T1w + 9); K2.Rcw = mat33(T2w); K2.tcw = vec3(T2w + 9);
    Points P1, P2;
    for (int i = 0; i < n1; i++) { MapPoint* p = P1.add(); fill_point(p, xyz1 + 3 * (size_t)i, 0, dist1_4 + 4 * (size_t)i, mpdesc1 + 32 * (size_t)i); if (flags1[i] & 1) K1.mvpMapPoints[i] = p; }
    for (int i = 0; i < n2; i++) { MapPoint* p = P2.add(); fill_point(p, xyz2 + 3 * (size_t)i, 0, dist2_4 + 4 * (size_t)i, mpdesc2 + 32 * (size_t)i); if (flags2[i] & 1) K2.mvpMapPoints[i] = p; }
    const cv::Mat R12 = mat33(R12_9), t12 = vec3(t12_3);
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) { S12_out[3 * r + c] = sR12.at<float>(r, c); S21_out[3 * r + c] = sR21.at<float>(r, c); }
        S12_out[9 + r] = t12.at<float>(r); S21_out[9 + r] = t21.at<float>(r);
    }
    vector<MapPoint*> vpMatches12((size_t)n1, static_cast<MapPoint*>(NULL));
    ORBmatcher matcher(0.75f, true);
    const int r = matcher.SearchBySim3(&K1, &K2, vpMatches12, s12, R12, t12, th);
    for (int i = 0; i < n1; i++) match12[i] = vpMatches12[i] ? vpMatches12[i]->index : -1;
    return r;
}

/* SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) */
int mref_search_for_triangulation(const int32_t* fv1_node, const int32_t* fv1_off, const int32_t* fv1_feat, int nfv1,
                                  const int32_t* fv2_node, const int32_t* fv2_off, const int32_t* fv2_feat, int nfv2,
                                  const KP28* kps1, const uint8_t* desc1, const uint8_t* skip1, const float* u_right1, int n1,
                                  const KP28* kps2, const uint8_t* desc2, const uint8_t* skip2, const float* u_right2, int n2,
                                  const float* geom28, const float* scale_factors2, const float* level_sigma2_2, int nlevels,
                                  int only_stereo, int check_orientation, int32_t* match12)
{
    const float cam9[9] = {geom28[24], geom28[25], geom28[26], geom28[27], 40.f, 0, 640, 0, 480};
    KeyFrame K1, K2;
    fill_keyframe(K1, kps1, desc1, n1, u_right1, cam9, scale_factors2, level_sigma2_2, 0, nlevels, 0.f);
    fill_keyframe(K2, kps2, desc2, n2, u_right2, cam9, scale_factors2, level_sigma2_2, 0, nlevels, 0.f);
    featvec(K1.mFeatVec, fv1_node, fv1_off, fv1_feat, nfv1); featvec(K2.mFeatVec, fv2_node, fv2_off, fv2_feat, nfv2);
    g_dummy_observed.index = -1;
    for (int i = 0; i < n1; i++) if (skip1 && skip1[i]) K1.mvpMapPoints[i] = &g_dummy_observed;
    for (int i = 0; i < n2; i++) if (skip2 && skip2[i]) K2.mvpMapPoints[i] = &g_dummy_observed;
    K1.Ow = vec3(geom28 + 9); K2.Rcw = mat33(geom28 + 12); K2.tcw = vec3(geom28 + 21);
    vector<pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6f, check_orientation != 0);
    const int r = matcher.SearchForTriangulation(&K1, &K2, mat33(geom28), pairs, only_stereo != 0);
    for (int i = 0; i < n1; i++) match12[i] = -1;
    for (size_t i = 0; i < pairs.size(); i++) match12[pairs[i].first] = (int32_t)pairs[i].second;
    return r;
}

/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) */
int mref_search_for_initialization(const KP28* kps1, const uint8_t* desc1, int n1, const KP28* kps2, const uint8_t* desc2, int n2,
                                   const float* bounds4, float* prev, int window, float nnratio, int check_orientation, int32_t* match12)
{
    const float cam9[9] = {1, 1, 0, 0, 1, bounds4[0], bounds4[1], bounds4[2], bounds4[3]};
    const float sf1[1] = {1.f};
    Frame F1, F2;
    fill_frame(F1, kps1, desc1, n1, 0, cam9, sf1, 1, 0.f);
    fill_frame(F2, kps2, desc2, n2, 0, cam9, sf1, 1, 0.f);
    vector<cv::Point2f> vbPrev((size_t)n1);
    for (int i = 0; i < n1; i++) { vbPrev[i].x = prev[2 * i]; vbPrev[i].y = prev[2 * i + 1]; }
    vector<int> vn;
    ORBmatcher matcher(nnratio, check_orientation != 0);
    const int r = matcher.SearchForInitialization(F1, F2, vbPrev, vn, window);
    for (int i = 0; i < n1; i++) { match12[i] = vn[i]; prev[2 * i] = vbPrev[i].x; prev[2 * i + 1] = vbPrev[i].y; }
    return r;
}

/* SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (kf_mode 0, match indexed by the frame's features) and
 * SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (kf_mode 1, indexed by keyframe 1's features) */
int mref_search_by_bow(const int32_t* fv1_node, const int32_t* fv1_off, const int32_t* fv1_feat, int nfv1,
                       const int32_t* fv2_node, const int32_t* fv2_off, const int32_t* fv2_feat, int nfv2,
                       const KP28* kps1, const uint8_t* desc1, const uint8_t* valid1, int n1,
                       const KP28* kps2, const uint8_t* desc2, const uint8_t* valid2, int n2,
                       float nnratio, int check_orientation, int kf_mode, int32_t* match)
{
    const float cam9[9] = {1, 1, 0, 0, 1, 0, 640, 0, 480};
    const float sf1[1] = {1.f};
    KeyFrame K1;
    fill_keyframe(K1, kps1, desc1, n1, 0, cam9, sf1, 0, 0, 1, 0.f);
    featvec(K1.mFeatVec, fv1_node, fv1_off, fv1_feat, nfv1);
    Points P1, P2;
    for (int i = 0; i < n1; i++) { MapPoint* p = P1.add(); if (!valid1 || valid1[i]) K1.mvpMapPoints[i] = p; }
    ORBmatcher matcher(nnratio, check_orientation != 0);
    int r;
    if (kf_mode == 0) {
        Frame F;
        fill_frame(F, kps2, desc2, n2, 0, cam9, sf1, 1, 0.f);
        featvec(F.mFeatVec, fv2_node, fv2_off, fv2_feat, nfv2);
        vector<MapPoint*> out;
        r = matcher.SearchByBoW(&K1, F, out);
        for (int j = 0; j < n2; j++) match[j] = out[j] ? out[j]->index : -1;
    } else {
        KeyFrame K2;
        fill_keyframe(K2, kps2, desc2, n2, 0, cam9, sf1, 0, 0, 1, 0.f);
        featvec(K2.mFeatVec, fv2_node, fv2_off, fv2_feat, nfv2);
        for (int i = 0; i < n2; i++) { MapPoint* p = P2.add(); if (!valid2 || valid2[i]) K2.mvpMapPoints[i] = p; }
        vector<MapPoint*> out;
        r = matcher.SearchByBoW(&K1, &K2, out);
        for (int i = 0; i < n1; i++) match[i] = out[i] ? out[i]->index : -1;
    }
    return r;
}

/* Frame::ComputeStereoMatches (Frame.cc:547-788) on two pyramids: lv_ptr[l] = payload origin of level l, lv_whs = (w, h, stride) */
void mref_stereo_match(const KP28* kl, const uint8_t* dl, int nl, const KP28* kr, const uint8_t* dr, int nr,
                       const uint8_t* const* lvL, const uint8_t* const* lvR, const int32_t* lv_whs, int nlevels,
                       const float* scale_factors, const float* inv_scale_factors, float mbf, float fx, float* u_right, float* depth)
{
    Frame F;
    F.N = nl; keys(kl, nl, F.mvKeys); keys(kr, nr, F.mvKeysRight);
    F.mDescriptors = desc_rows(dl, nl); F.mDescriptorsRight = desc_rows(dr, nr);
    F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels); F.mvInvScaleFactors.assign(inv_scale_factors, inv_scale_factors + nlevels);
    F.mbf = mbf; F.mb = mbf / fx;                                           /* Frame.cc:120 */
    ORBextractor L, R;
    for (int l = 0; l < nlevels; l++) {
        L.mvImagePyramid.push_back(cv::Mat(lv_whs[3 * l + 1], lv_whs[3 * l], CV_8U, const_cast<uint8_t*>(lvL[l]), (size_t)lv_whs[3 * l + 2]));
        R.mvImagePyramid.push_back(cv::Mat(lv_whs[3 * l + 1], lv_whs[3 * l], CV_8U, const_cast<uint8_t*>(lvR[l]), (size_t)lv_whs[3 * l + 2]));
    }
    F.mpORBextractorLeft = &L; F.mpORBextractorRight = &R;
    F.ComputeStereoMatches();
    for (int i = 0; i < nl; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

/* Frame::isInFrustum (Frame.cc:315-378) per point; q5 rows = (mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos), level */
void mref_is_in_frustum(const float* Tcw12, const float* Ow3, const float* cam9, int nlevels, float log_sf, const float* pt_xyz,
                        const float* pt_normal, const float* pt_dist4, int npts, float limit, float* q4, int32_t* level, uint8_t* in_view)
{
    set_frame_statics(cam9);
    Frame F;
    F.mRcw = mat33(Tcw12); F.mtcw = vec3(Tcw12 + 9); F.mOw = vec3(Ow3); F.mbf = cam9[4];
    F.mnScaleLevels = nlevels; F.mfLogScaleFactor = log_sf;
    for (int i = 0; i < npts; i++) {
        MapPoint p;
        fill_point(&p, pt_xyz + 3 * (size_t)i, pt_normal + 3 * (size_t)i, pt_dist4 + 4 * (size_t)i, 0);
        in_view[i] = F.isInFrustum(&p, limit) ? 1 : 0;
        q4[4 * i] = p.mTrackProjX; q4[4 * i + 1] = p.mTrackProjY; q4[4 * i + 2] = p.mTrackProjXR; q4[4 * i + 3] = p.mTrackViewCos;
        level[i] = p.mnTrackScaleLevel;
    }
}
}

// MatcherOps.h — header-only C++ wrappers over liborbx.so for the remaining ORBmatcher functions. Every wrapper keeps the
// reference's name, return value and result containers; what the reference reads through MapPoint* / KeyFrame* pointers
// (world position, normal, distance invariance, descriptor, flags) is passed flattened, because the map graph stays on
// the host. The helper MapPointTable shows the flattening loop a call site runs once per call.
#ifndef ORB_SLAM2_MATCHER_OPS_H
#define ORB_SLAM2_MATCHER_OPS_H

#include <cstring>
#include <utility>
#include <vector>
#include <opencv2/core/core.hpp>
#include "../../include/orbx.h"

namespace ORB_SLAM2
{
// Flattened view of a vector<MapPoint*>: entry i describes vpMapPoints[i].
//   xyz      pMP->GetWorldPos()                       3 floats
//   normal   pMP->GetNormal()                         3 floats
//   dist     GetMinDistanceInvariance(), GetMaxDistanceInvariance(), mfMaxDistance      3 floats
//   desc     pMP->GetDescriptor()                     32 bytes
//   flags    caller-defined per function (see each wrapper)
struct MapPointTable {
    std::vector<float> xyz, normal, dist;
    std::vector<unsigned char> desc, flags;
    size_t size() const { return flags.size(); }
    void resize(size_t n) { xyz.resize(3 * n); normal.resize(3 * n); dist.resize(3 * n); desc.resize(32 * n); flags.assign(n, 0); }
};

struct KeyFrameCamera {        // fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY + the level tables of Frame / KeyFrame
    float camera9[9];
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float mfLogScaleFactor;
};

inline void PackDescriptors(const cv::Mat& m, std::vector<unsigned char>& out)
{
    out.resize((size_t)m.rows * 32);
    for (int i = 0; i < m.rows; i++) std::memcpy(&out[(size_t)i * 32], m.ptr(i), 32);
}

// Frame::isInFrustum(pMP, viewingCosLimit) (Frame.cc:315-378) for all local map points at once, as the loop of
// Tracking::SearchLocalPoints (Tracking.cc:1409-1470) needs it. Tcw12 = (mRcw row-major, mtcw), Ow3 = mOw. queries[i] receives
// the five mTrack* fields where inView[i] != 0 (= the return value = mbTrackInView) and feeds SearchByProjectionGPU below.
inline bool IsInFrustumGPU(const float Tcw12[12], const float Ow3[3], const KeyFrameCamera& cam, const MapPointTable& pts,
                           float viewingCosLimit, std::vector<OrbxTrackQuery>& queries, std::vector<unsigned char>& inView, int device = 0)
{
    queries.resize(pts.size()); inView.assign(pts.size(), 0);
    if (pts.size() == 0) return true;
    return orbx_is_in_frustum(Tcw12, Ow3, cam.camera9, (int)cam.mvScaleFactors.size(), cam.mfLogScaleFactor, pts.xyz.data(),
                              pts.normal.data(), pts.dist.data(), (int)pts.size(), viewingCosLimit, queries.data(), inView.data(),
                              device) == ORBX_OK;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (ORBmatcher.cc:46-142).
// queries[i] = (mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mnTrackScaleLevel) of vpMapPoints[i];
// flags bit 0: mbTrackInView && !isBad(), bit 1: Observations() > 0. occupied[k]: F.mvpMapPoints[k] has observations.
// matchOfKeypoint[k] >= 0 means F.mvpMapPoints[k] = vpMapPoints[matchOfKeypoint[k]]. Returns nmatches, -1 on error.
inline int SearchByProjectionGPU(const std::vector<cv::KeyPoint>& mvKeysUn, const cv::Mat& mDescriptors, const std::vector<float>* mvuRight,
                                 const std::vector<unsigned char>* occupied, const float bounds4[4],
                                 const std::vector<float>& mvScaleFactors, const std::vector<OrbxTrackQuery>& queries,
                                 const std::vector<unsigned char>& queryDescriptors, const std::vector<unsigned char>& queryFlags,
                                 float th, float mfNNratio, std::vector<int>& matchOfKeypoint, int device = 0)
{
    matchOfKeypoint.assign(mvKeysUn.size(), -1);
    if (mvKeysUn.empty() || queries.empty()) return 0;
    std::vector<unsigned char> d; PackDescriptors(mDescriptors, d);
    OrbxLocalPointsFrame f;
    f.keypoints = reinterpret_cast<const OrbxKeyPoint*>(mvKeysUn.data()); f.descriptors = d.data();
    f.u_right = mvuRight ? mvuRight->data() : 0; f.occupied = occupied ? occupied->data() : 0; f.n = (int32_t)mvKeysUn.size();
    f.queries = queries.data(); f.query_descriptors = queryDescriptors.data(); f.query_flags = queryFlags.data(); f.nq = (int32_t)queries.size();
    int32_t nmatches = 0;
    f.match = matchOfKeypoint.data(); f.nmatches = &nmatches;
    if (orbx_search_local_points(&f, bounds4, mvScaleFactors.data(), (int)mvScaleFactors.size(), th, mfNNratio, device) != ORBX_OK) return -1;
    return nmatches;
}

// Search half of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th) (ORBmatcher.cc:918-1092; sim3 =
// false) and of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:1094-1236; sim3 = true, Tcw12 / Ow3 from Scw as :1101-1106).
// pts.flags bit 0: non-NULL, !isBad(), !IsInKeyFrame(pKF). bestIdx[i] >= 0: fuse vpMapPoints[i] with keyframe feature
// bestIdx[i] (the caller runs :1070-1087 / :1213-1229 on it). Returns nFused, -1 on error.
inline int FuseSearchGPU(const std::vector<cv::KeyPoint>& mvKeysUn, const cv::Mat& mDescriptors, const std::vector<float>* mvuRight,
                         const float Tcw12[12], const float Ow3[3], const KeyFrameCamera& cam, const MapPointTable& pts, float th,
                         bool sim3, std::vector<int>& bestIdx, int device = 0)
{
    bestIdx.assign(pts.size(), -1);
    if (pts.size() == 0) return 0;
    std::vector<unsigned char> d; PackDescriptors(mDescriptors, d);
    std::vector<int32_t> bestDist(pts.size());
    OrbxFuseJob j;
    j.keypoints = reinterpret_cast<const OrbxKeyPoint*>(mvKeysUn.data()); j.descriptors = d.data(); j.u_right = mvuRight ? mvuRight->data() : 0;
    j.n = (int32_t)mvKeysUn.size();
    std::memcpy(j.Tcw, Tcw12, sizeof j.Tcw); std::memcpy(j.Ow, Ow3, sizeof j.Ow);
    j.pt_xyz = pts.xyz.data(); j.pt_normal = pts.normal.data(); j.pt_dist = pts.dist.data(); j.pt_descriptors = pts.desc.data();
    j.pt_flags = pts.flags.data(); j.npts = (int32_t)pts.size(); j.th = th; j.mode = sim3 ? 1 : 0;
    int32_t nFused = 0;
    j.best_idx = bestIdx.data(); j.best_dist = bestDist.data(); j.nfused = &nFused;
    if (orbx_fuse_search(&j, cam.camera9, cam.mvScaleFactors.data(), cam.mvInvLevelSigma2.data(), (int)cam.mvScaleFactors.size(),
                         cam.mfLogScaleFactor, device) != ORBX_OK) return -1;
    return nFused;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
// (ORBmatcher.cc:1648-1795; loopClosing = false, pts = pKF->GetMapPointMatches(), ptAngle[i] = pKF->mvKeysUn[i].angle) and
// ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw, vpPoints, vpMatched, int th) (:327-440; loopClosing = true,
// maxDist = TH_LOW). pts.flags bit 0: non-NULL, !isBad(), not already found. occupied[k]: the feature holds a map point.
// matchOfFeature[k] >= 0: the feature receives pts[matchOfFeature[k]]. Returns nmatches, -1 on error.
inline int SearchByProjectionGPU(const std::vector<cv::KeyPoint>& mvKeysUn, const cv::Mat& mDescriptors,
                                 const std::vector<unsigned char>* occupied, const float Tcw12[12], const float Ow3[3],
                                 const KeyFrameCamera& cam, const MapPointTable& pts, const std::vector<float>* ptAngle, float th,
                                 int maxDist, bool loopClosing, bool checkOrientation, std::vector<int>& matchOfFeature, int device = 0)
{
    matchOfFeature.assign(mvKeysUn.size(), -1);
    if (mvKeysUn.empty() || pts.size() == 0) return 0;
    std::vector<unsigned char> d; PackDescriptors(mDescriptors, d);
    OrbxProjectionJob j;
    j.keypoints = reinterpret_cast<const OrbxKeyPoint*>(mvKeysUn.data()); j.descriptors = d.data(); j.occupied = occupied ? occupied->data() : 0;
    j.n = (int32_t)mvKeysUn.size();
    std::memcpy(j.Tcw, Tcw12, sizeof j.Tcw); std::memcpy(j.Ow, Ow3, sizeof j.Ow);
    j.pt_xyz = pts.xyz.data(); j.pt_normal = pts.normal.data(); j.pt_dist = pts.dist.data(); j.pt_descriptors = pts.desc.data();
    j.pt_flags = pts.flags.data(); j.pt_angle = ptAngle ? ptAngle->data() : 0; j.npts = (int32_t)pts.size();
    j.th = th; j.max_dist = maxDist; j.mode = loopClosing ? 1 : 0;
    int32_t nmatches = 0;
    j.match = matchOfFeature.data(); j.nmatches = &nmatches;
    if (orbx_search_by_projection_kf(&j, cam.camera9, cam.mvScaleFactors.data(), (int)cam.mvScaleFactors.size(), cam.mfLogScaleFactor,
                                     checkOrientation ? 1 : 0, device) != ORBX_OK) return -1;
    return nmatches;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:738-916).
// hasMapPoint*[i]: pKF->GetMapPoint(i) != NULL. geom28 = F12 row-major (9), pKF1->GetCameraCenter() (3), pKF2->GetRotation()
// (9), pKF2->GetTranslation() (3), pKF2's fx, fy, cx, cy. Returns nmatches and fills vMatchedPairs, -1 on error.
inline int SearchForTriangulationGPU(orbx_vocabulary* voc, const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& descriptors1,
                                     const std::vector<unsigned char>& hasMapPoint1, const std::vector<float>* mvuRight1,
                                     const std::vector<cv::KeyPoint>& keysUn2, const cv::Mat& descriptors2,
                                     const std::vector<unsigned char>& hasMapPoint2, const std::vector<float>* mvuRight2,
                                     const float geom28[28], const KeyFrameCamera& cam2, bool bOnlyStereo, bool checkOrientation,
                                     std::vector<std::pair<size_t, size_t> >& vMatchedPairs, int levelsup = 4)
{
    vMatchedPairs.clear();
    if (keysUn1.empty() || keysUn2.empty()) return 0;
    std::vector<unsigned char> d1, d2; PackDescriptors(descriptors1, d1); PackDescriptors(descriptors2, d2);
    std::vector<int32_t> match12(keysUn1.size(), -1);
    int32_t nmatches = 0;
    if (orbx_search_for_triangulation(voc, reinterpret_cast<const OrbxKeyPoint*>(keysUn1.data()), d1.data(), (int)keysUn1.size(),
                                      hasMapPoint1.data(), mvuRight1 ? mvuRight1->data() : 0,
                                      reinterpret_cast<const OrbxKeyPoint*>(keysUn2.data()), d2.data(), (int)keysUn2.size(),
                                      hasMapPoint2.data(), mvuRight2 ? mvuRight2->data() : 0, geom28, cam2.mvScaleFactors.data(),
                                      cam2.mvLevelSigma2.data(), (int)cam2.mvScaleFactors.size(), levelsup, bOnlyStereo ? 1 : 0,
                                      checkOrientation ? 1 : 0, match12.data(), &nmatches) != ORBX_OK) return -1;
    vMatchedPairs.reserve((size_t)nmatches);
    for (size_t i = 0; i < match12.size(); i++)
        if (match12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)match12[i]));
    return nmatches;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (ORBmatcher.cc:1238-1487). pts1 / pts2: the map point
// of every feature of the keyframe (flags bit 0: exists, !isBad(), feature not in vbAlreadyMatched); S12 = (s12*R12, t12),
// S21 = ((1.0/s12)*R12.t(), -sR21*t12) as :1253-1255 compute them. match12[i1] >= 0: vpMatches12[i1] =
// pKF2->GetMapPointMatches()[match12[i1]]. Returns nFound, -1 on error.
inline int SearchBySim3GPU(const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& descriptors1, const MapPointTable& pts1, const float T1w[12],
                           const std::vector<cv::KeyPoint>& keysUn2, const cv::Mat& descriptors2, const MapPointTable& pts2, const float T2w[12],
                           const float S12[12], const float S21[12], const KeyFrameCamera& cam, float th, std::vector<int>& match12,
                           int device = 0)
{
    match12.assign(keysUn1.size(), -1);
    std::vector<unsigned char> d1, d2; PackDescriptors(descriptors1, d1); PackDescriptors(descriptors2, d2);
    OrbxSim3KeyFrame k1, k2;
    k1.keypoints = reinterpret_cast<const OrbxKeyPoint*>(keysUn1.data()); k1.descriptors = d1.data(); k1.n = (int32_t)keysUn1.size();
    k1.mp_xyz = pts1.xyz.data(); k1.mp_dist = pts1.dist.data(); k1.mp_descriptors = pts1.desc.data(); k1.mp_flags = pts1.flags.data();
    std::memcpy(k1.Tcw, T1w, sizeof k1.Tcw);
    k2.keypoints = reinterpret_cast<const OrbxKeyPoint*>(keysUn2.data()); k2.descriptors = d2.data(); k2.n = (int32_t)keysUn2.size();
    k2.mp_xyz = pts2.xyz.data(); k2.mp_dist = pts2.dist.data(); k2.mp_descriptors = pts2.desc.data(); k2.mp_flags = pts2.flags.data();
    std::memcpy(k2.Tcw, T2w, sizeof k2.Tcw);
    int32_t nFound = 0;
    if (orbx_search_by_sim3(&k1, &k2, S12, S21, cam.camera9, cam.mvScaleFactors.data(), (int)cam.mvScaleFactors.size(),
                            cam.mfLogScaleFactor, th, match12.data(), &nFound, device) != ORBX_OK) return -1;
    return nFound;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:442-587).
inline int SearchForInitializationGPU(const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& descriptors1,
                                      const std::vector<cv::KeyPoint>& keysUn2, const cv::Mat& descriptors2, const float bounds4[4],
                                      std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize,
                                      float mfNNratio, bool checkOrientation, int device = 0)
{
    vnMatches12.assign(keysUn1.size(), -1);
    if (keysUn1.empty()) return 0;
    std::vector<unsigned char> d1, d2; PackDescriptors(descriptors1, d1); PackDescriptors(descriptors2, d2);
    OrbxInitPair p;
    p.keypoints1 = reinterpret_cast<const OrbxKeyPoint*>(keysUn1.data()); p.descriptors1 = d1.data(); p.n1 = (int32_t)keysUn1.size();
    p.keypoints2 = reinterpret_cast<const OrbxKeyPoint*>(keysUn2.data()); p.descriptors2 = d2.data(); p.n2 = (int32_t)keysUn2.size();
    p.prev_matched = reinterpret_cast<const float*>(vbPrevMatched.data());       // cv::Point2f = two packed floats
    p.prev_matched_out = reinterpret_cast<float*>(vbPrevMatched.data());
    p.window_size = windowSize;
    int32_t nmatches = 0;
    p.match12 = vnMatches12.data(); p.nmatches = &nmatches;
    if (orbx_search_for_initialization(&p, bounds4, mfNNratio, checkOrientation ? 1 : 0, device) != ORBX_OK) return -1;
    return nmatches;
}
} // namespace ORB_SLAM2
#endif

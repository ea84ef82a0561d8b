// FrameOps.h — header-only C++ wrappers over liborbx.so for the Frame / ORBmatcher rows around the extractor
// (SURVEY.md §8 f-2..f-4). Each function names the reference code it replaces; signatures use the reference's own
// container types (std::vector<cv::KeyPoint>, cv::Mat CV_8U N x 32, DBoW2's std::map bases) so the call sites in
// Frame.cc / ORBmatcher.cc / stereo_euroc.cc change by one line each.
#ifndef ORB_SLAM2_FRAME_OPS_H
#define ORB_SLAM2_FRAME_OPS_H

#include <cstring>
#include <map>
#include <vector>
#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "../../include/orbx.h"

namespace ORB_SLAM2
{
// Frame::UndistortKeyPoints (Frame.cc:471-506). K4 = (fx, fy, cx, cy) = mK.at<float>(0,0), (1,1), (0,2), (1,2);
// dist = mDistCoef.ptr<float>() with ndist = 4 or 5 coefficients.
inline bool UndistortKeyPointsGPU(const std::vector<cv::KeyPoint>& mvKeys, const float K4[4], const float* dist, int ndist,
                                  std::vector<cv::KeyPoint>& mvKeysUn, int device = 0)
{
    mvKeysUn.resize(mvKeys.size());
    if (mvKeys.empty()) return true;
    return orbx_undistort_keypoints(reinterpret_cast<const OrbxKeyPoint*>(mvKeys.data()), (int)mvKeys.size(), K4, dist, ndist,
                                    reinterpret_cast<OrbxKeyPoint*>(mvKeysUn.data()), device) == ORBX_OK;
}

// Frame::ComputeImageBounds (Frame.cc:508-538)
inline bool ComputeImageBoundsGPU(int cols, int rows, const float K4[4], const float* dist, int ndist,
                                  float& mnMinX, float& mnMaxX, float& mnMinY, float& mnMaxY, int device = 0)
{
    float b[4];
    if (orbx_image_bounds(cols, rows, K4, dist, ndist, b, device) != ORBX_OK) return false;
    mnMinX = b[0]; mnMaxX = b[1]; mnMinY = b[2]; mnMaxY = b[3];
    return true;
}

// Examples/Stereo/stereo_euroc.cc:97-98: after cv::initUndistortRectifyMap(..., CV_32F, M1, M2) hand the maps over once ...
inline bool SetRectifyMaps(ORBextractor* extractor, const float* M1, const float* M2, int map_cols, int map_rows,
                           int map_stride_floats, int src_cols, int src_rows)
{
    return orbx_set_rectify_maps(extractor->Handle(), M1, M2, map_cols, map_rows, map_stride_floats, src_cols, src_rows) == ORBX_OK;
}
// ... then :136-137 + Frame::ExtractORB (Frame.cc:273-279) become one call on the UNRECTIFIED image
inline bool ExtractRectified(ORBextractor* extractor, const cv::Mat& imUnrectified, std::vector<cv::KeyPoint>& keypoints,
                             cv::Mat& descriptors)
{
    orbx_extractor* h = extractor->Handle();
    if (imUnrectified.empty()) return true;
    // the pyramid (and with it the keypoint capacity) has the MAP's size, which may differ from the source frame's
    int mw = 0, mh = 0;
    if (orbx_rectify_map_size(h, &mw, &mh) != ORBX_OK || orbx_reserve(h, mw, mh, 1) != ORBX_OK) return false;
    const int cap = orbx_max_keypoints(h);
    std::vector<OrbxKeyPoint> kp((size_t)cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int n = 0;
    const unsigned char* imgs[1] = {imUnrectified.data};
    if (orbx_extract_batch_rectified(h, imgs, 1, (int)imUnrectified.step, kp.data(), cap, &n, desc.data()) != ORBX_OK) return false;
    keypoints.resize((size_t)n);
    if (n == 0) { descriptors.release(); return true; }
    std::memcpy(static_cast<void*>(keypoints.data()), kp.data(), (size_t)n * sizeof(OrbxKeyPoint));
    descriptors.create(n, 32, CV_8U);
    for (int i = 0; i < n; i++) std::memcpy(descriptors.ptr(i), &desc[(size_t)i * 32], 32);
    return true;
}

// Frame::ComputeBoW (Frame.cc:462-469): mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4).
// DBoW2::BowVector is a std::map<unsigned, double>, DBoW2::FeatureVector a std::map<unsigned, std::vector<unsigned> >.
inline bool ComputeBoWGPU(orbx_vocabulary* voc, const cv::Mat& mDescriptors, std::map<unsigned int, double>& mBowVec,
                          std::map<unsigned int, std::vector<unsigned int> >& mFeatVec, int levelsup = 4)
{
    mBowVec.clear(); mFeatVec.clear();
    const int n = mDescriptors.rows;
    if (n == 0) return true;
    std::vector<unsigned char> d((size_t)n * 32);
    for (int i = 0; i < n; i++) std::memcpy(&d[(size_t)i * 32], mDescriptors.ptr(i), 32);
    const int32_t counts[1] = {n};
    // transform -> get is one critical section on the shared vocabulary handle (Tracking, LocalMapping and LoopClosing
    // all call ComputeBoW on the same ORBVocabulary)
    struct VocLock { orbx_vocabulary* v; explicit VocLock(orbx_vocabulary* v_) : v(v_) { orbx_vocab_lock(v); } ~VocLock() { orbx_vocab_unlock(v); } } lock(voc);
    if (orbx_bow_transform(voc, d.data(), counts, 1, n, levelsup) != ORBX_OK) return false;
    std::vector<int32_t> bow_id(n), fv_node(n), fv_off(n + 1), fv_feat(n);
    std::vector<double> bow_val(n);
    int32_t nb = 0, nf = 0;
    if (orbx_bow_get(voc, 0, 0, 0, bow_id.data(), bow_val.data(), &nb, fv_node.data(), fv_off.data(), fv_feat.data(), &nf) != ORBX_OK)
        return false;
    for (int j = 0; j < nb; j++) mBowVec.insert(mBowVec.end(), std::make_pair((unsigned int)bow_id[j], bow_val[j]));
    for (int j = 0; j < nf; j++) {
        std::vector<unsigned int>& v = mFeatVec.insert(mFeatVec.end(), std::make_pair((unsigned int)fv_node[j], std::vector<unsigned int>()))->second;
        v.assign(fv_feat.begin() + fv_off[j], fv_feat.begin() + fv_off[j + 1]);
    }
    return true;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1489-1646).
// The caller flattens what the loop reads: per last-frame keypoint the map point's world position / descriptor / flags
// (bit 0: pMP && !mvbOutlier[i]; bit 1: pMP->Observations() > 0), the pose CurrentFrame.mTcw (Rcw row-major, tcw),
// camera9 = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY) and mode (0, 1 = bForward, 2 = bBackward).
// matchOfCurrent[k] >= 0 means CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[matchOfCurrent[k]].
inline int SearchByProjectionGPU(const std::vector<cv::KeyPoint>& curKeysUn, const cv::Mat& curDescriptors,
                                 const std::vector<float>* curURight, const std::vector<unsigned char>* curOccupied,
                                 const float Tcw12[12], const float camera9[9], const std::vector<float>& scaleFactors,
                                 const std::vector<cv::KeyPoint>& lastKeysUn, const std::vector<float>& lastXYZ,
                                 const std::vector<unsigned char>& lastDescriptors, const std::vector<unsigned char>& lastFlags,
                                 float th, int mode, bool checkOrientation, std::vector<int>& matchOfCurrent, int device = 0)
{
    matchOfCurrent.assign(curKeysUn.size(), -1);
    if (curKeysUn.empty() || lastKeysUn.empty()) return 0;
    std::vector<unsigned char> cd(curKeysUn.size() * 32);
    for (size_t i = 0; i < curKeysUn.size(); i++) std::memcpy(&cd[i * 32], curDescriptors.ptr((int)i), 32);
    OrbxProjectionPair p;
    p.cur_keypoints = reinterpret_cast<const OrbxKeyPoint*>(curKeysUn.data()); p.cur_descriptors = cd.data();
    p.cur_u_right = curURight ? curURight->data() : 0; p.cur_occupied = curOccupied ? curOccupied->data() : 0;
    p.n_cur = (int32_t)curKeysUn.size();
    p.last_keypoints = reinterpret_cast<const OrbxKeyPoint*>(lastKeysUn.data()); p.last_xyz = lastXYZ.data();
    p.last_descriptors = lastDescriptors.data(); p.last_flags = lastFlags.data(); p.n_last = (int32_t)lastKeysUn.size();
    std::memcpy(p.Tcw, Tcw12, sizeof p.Tcw); p.mode = mode;
    int32_t nmatches = 0;
    p.match = matchOfCurrent.data(); p.nmatches = &nmatches;
    if (orbx_search_by_projection(&p, camera9, scaleFactors.data(), (int)scaleFactors.size(), th, checkOrientation ? 1 : 0, device) != ORBX_OK)
        return -1;
    return nmatches;
}
} // namespace ORB_SLAM2
#endif

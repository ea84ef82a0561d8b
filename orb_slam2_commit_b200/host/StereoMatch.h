// StereoMatch.h — GPU form of Frame::ComputeStereoMatches (reference: src/Frame.cc:547-788), header-only wrapper over
// liborbx.so. Drop-in use inside Frame::ComputeStereoMatches():
//     ORB_SLAM2::ComputeStereoMatchesGPU(mpORBextractorLeft, mpORBextractorRight, mvKeys, mDescriptors,
//                                        mvKeysRight, mDescriptorsRight, mbf, fx, mvuRight, mvDepth);
// The level pyramids are read where they lie in HBM (both extractors have just processed the pair), so with this
// call in place the extractors' mbDownloadPyramid can be switched off.
#ifndef ORB_SLAM2_STEREO_MATCH_H
#define ORB_SLAM2_STEREO_MATCH_H

#include <vector>
#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "../../include/orbx.h"

namespace ORB_SLAM2
{
inline bool ComputeStereoMatchesGPU(ORBextractor* left, ORBextractor* right,
                                    const std::vector<cv::KeyPoint>& keysLeft, const cv::Mat& descLeft,
                                    const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descRight,
                                    float mbf, float fx, std::vector<float>& mvuRight, std::vector<float>& mvDepth)
{
    const int N = (int)keysLeft.size();
    mvuRight.assign(N, -1.0f);                     // Frame.cc:549-550
    mvDepth.assign(N, -1.0f);
    if (N == 0) return true;
    return orbx_stereo_match(left->Handle(), right->Handle(),
                             reinterpret_cast<const OrbxKeyPoint*>(keysLeft.data()), descLeft.data, N,
                             reinterpret_cast<const OrbxKeyPoint*>(keysRight.data()), descRight.data, (int)keysRight.size(),
                             mbf, fx, mvuRight.data(), mvDepth.data()) == ORBX_OK;
}
} // namespace ORB_SLAM2
#endif

// ORBextractor.h — drop-in replacement for ORB-SLAM2's include/ORBextractor.h (reference :51-145).
//
// Same namespace, class name, constructor, operator(), getters and public mvImagePyramid member as the reference, so
// Frame.cc (ExtractORB :273-279, scale-table copies :70-76, ComputeStereoMatches :556,681-700) and Tracking.cc
// (:120-126) compile and link against it unchanged. All work is done by liborbx.so (include/orbx.h, hand-written
// sm_100a CUDA kernels); this header needs only OpenCV's core types.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <opencv2/core/core.hpp>

struct orbx_extractor;

namespace ORB_SLAM2
{

class ORBextractor
{
public:
    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    // reference: ORBextractor.h:61. `device` is the only addition (CUDA ordinal, default 0).
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // reference: ORBextractor.h:77. Mask is ignored (as in the reference). An empty image returns silently with the
    // outputs untouched (ORBextractor.cc:1141); a non-8UC1 image asserts (ORBextractor.cc:1146).
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints,
      cv::OutputArray descriptors);

    int inline GetLevels(){ return nlevels; }
    float inline GetScaleFactor(){ return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors(){ return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors(){ return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares(){ return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares(){ return mvInvLevelSigma2; }

    // reference: ORBextractor.h:104. Filled after every operator() call when mbDownloadPyramid is true (default):
    // each level is a cv::Mat header over the handle's pinned mirror of the frame's raw pyramid block (one asynchronous
    // device-to-host copy per frame); its step is the device pitch and the BORDER_REFLECT_101 apron lies around the
    // payload, exactly as around the reference's view at (19,19) of `temp`. Valid until the next operator() call.
    // Monocular / RGB-D users can switch the download off (nothing reads the pyramid there).
    std::vector<cv::Mat> mvImagePyramid;
    bool mbDownloadPyramid;

    // The underlying C handle, e.g. to reach the device-resident pyramid (orbx_pyramid_level_device).
    orbx_extractor* Handle() { return mpHandle; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel;
    orbx_extractor* mpHandle;
};

} //namespace ORB_SLAM

#endif

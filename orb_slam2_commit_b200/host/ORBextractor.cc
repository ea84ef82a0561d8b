// ORBextractor.cc — the reference class surface (include/ORBextractor.h:51-145) over the C ABI of liborbx.so.
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "../../include/orbx.h"

namespace ORB_SLAM2
{

static void die(const char* what)
{
    // the reference has no error channel (no exceptions, no return codes): a failing device is fatal, never a
    // silent CPU fallback
    std::fprintf(stderr, "ORBextractor(B200): %s failed: %s\n", what, orbx_last_error());
    std::abort();
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST, int device):
    mbDownloadPyramid(true), nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels),
    iniThFAST(_iniThFAST), minThFAST(_minThFAST), mpHandle(0)
{
    if (orbx_create(nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, device, &mpHandle) != ORBX_OK) die("orbx_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    if (orbx_get_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                        mvInvLevelSigma2.data(), mnFeaturesPerLevel.data(), 0) != ORBX_OK) die("orbx_get_tables");
    mvImagePyramid.resize(nlevels);
    mvWhole.resize(nlevels);
}

ORBextractor::~ORBextractor() { orbx_destroy(mpHandle); }

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                      cv::OutputArray _descriptors)
{
    if(_image.empty())
        return;

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1 );

    if (orbx_reserve(mpHandle, image.cols, image.rows, 1) != ORBX_OK) die("orbx_reserve");
    const int cap = orbx_max_keypoints(mpHandle);
    static_assert(sizeof(cv::KeyPoint) == sizeof(OrbxKeyPoint), "cv::KeyPoint must be the 28-byte POD of OpenCV 2.4/3.x/4.x");
    _keypoints.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    if (orbx_extract(mpHandle, image.data, image.cols, image.rows, (int)image.step,
                     reinterpret_cast<OrbxKeyPoint*>(_keypoints.data()), cap, &n, desc.data) != ORBX_OK) die("orbx_extract");
    _keypoints.resize(n);
    if (n == 0)
        _descriptors.release();
    else
    {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (int i = 0; i < n; i++) std::copy(desc.ptr(i), desc.ptr(i) + 32, out.ptr(i));
    }

    if (mbDownloadPyramid)
        for (int level = 0; level < nlevels; ++level)
        {
            int w = 0, h = 0;
            if (orbx_level_size(mpHandle, level, &w, &h) != ORBX_OK) die("orbx_level_size");
            mvWhole[level].create(h + 38, w + 38, CV_8UC1);
            if (orbx_pyramid_level(mpHandle, 0, level, mvWhole[level].data, (int)mvWhole[level].step) != ORBX_OK) die("orbx_pyramid_level");
            mvImagePyramid[level] = mvWhole[level](cv::Rect(19, 19, w, h));
        }
}

} //namespace ORB_SLAM

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
    if (orbx_set_pyramid_mirror(mpHandle, mbDownloadPyramid ? 1 : 0) != ORBX_OK) die("orbx_set_pyramid_mirror");
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
    {
        // ONE asynchronous copy (issued by orbx_extract itself, in the stream of its kernels) brought the frame's whole
        // raw pyramid block into the handle's pinned mirror; every level is a cv::Mat header over it with the device
        // pitch as step, so the 19-px apron lies around the payload exactly as around the reference's ROI of `temp`
        const unsigned char* block = 0;
        if (orbx_pyramid_mirror(mpHandle, 0, &block) != ORBX_OK) die("orbx_pyramid_mirror");
        for (int level = 0; level < nlevels; ++level)
        {
            int w = 0, h = 0, pitch = 0; size_t off = 0;
            if (orbx_level_size(mpHandle, level, &w, &h) != ORBX_OK ||
                orbx_pyramid_level_layout(mpHandle, level, &off, &pitch, 0) != ORBX_OK) die("orbx_pyramid_level_layout");
            mvImagePyramid[level] = cv::Mat(h, w, CV_8UC1, const_cast<unsigned char*>(block + off), (size_t)pitch);
        }
    }
}

} //namespace ORB_SLAM

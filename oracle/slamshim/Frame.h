/* oracle/slamshim/Frame.h — stub of ORB_SLAM2::Frame with the members ORBmatcher.cc reads (test infrastructure).
 * AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid come verbatim from /root/reference/src/Frame.cc:254-271, 388-444,
 * 446-459 through the build recipe (oracle/Makefile, target ref_matcher). */
#ifndef SLAMSHIM_FRAME_H
#define SLAMSHIM_FRAME_H
#include <vector>
#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
using namespace std;   /* the real Frame.h gets this through ORBVocabulary.h -> DBoW2/TemplatedVocabulary.h:36; ORBmatcher.h relies on it */
namespace ORB_SLAM2
{
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
class MapPoint;
class ORBextractor { public: std::vector<cv::Mat> mvImagePyramid; };   /* ORBextractor.h:104, all ComputeStereoMatches reads of it */
class Frame
{
public:
    Frame() : N(0), mb(0), mbf(0), mnScaleLevels(0), mfLogScaleFactor(0), mpORBextractorLeft(0), mpORBextractorRight(0) {}
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);   /* verbatim from Frame.cc:315-378 */
    cv::Mat mRcw, mtcw, mOw;
    void ComputeStereoMatches();                  /* verbatim from Frame.cc:547-788 through the build recipe */
    std::vector<cv::KeyPoint> mvKeysRight;
    cv::Mat mDescriptorsRight;
    std::vector<float> mvDepth, mvInvScaleFactors;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    void AssignFeaturesToGrid();

    static float fx, fy, cx, cy;
    float mb, mbf;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTcw;
    int mnScaleLevels;
    float mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
}
#endif

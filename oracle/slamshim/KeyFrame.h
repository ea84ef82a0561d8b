/* oracle/slamshim/KeyFrame.h — stub of ORB_SLAM2::KeyFrame with the members ORBmatcher.cc reads (test infrastructure).
 * GetFeaturesInArea / IsInImage come verbatim from /root/reference/src/KeyFrame.cc:708-752 through the build recipe. */
#ifndef SLAMSHIM_KEYFRAME_H
#define SLAMSHIM_KEYFRAME_H
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "Frame.h"
namespace ORB_SLAM2
{
class MapPoint;
class KeyFrame
{
public:
    KeyFrame() : mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS), mfGridElementWidthInv(0), mfGridElementHeightInv(0),
                 fx(0), fy(0), cx(0), cy(0), mbf(0), N(0), mnScaleLevels(0), mfLogScaleFactor(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0) {}
    cv::Mat GetRotation() { return Rcw; }
    cv::Mat GetTranslation() { return tcw; }
    cv::Mat GetCameraCenter() { return Ow; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (size_t i = 0; i < mvpMapPoints.size(); i++) if (mvpMapPoints[i]) s.insert(mvpMapPoints[i]); return s; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx);   /* matcher_glue.cc: records the pick in pMP->fused, performs nothing */
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const;

    int mnGridCols, mnGridRows;
    float mfGridElementWidthInv, mfGridElementHeightInv;
    float fx, fy, cx, cy, mbf;
    int N;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    int mnScaleLevels;
    float mfLogScaleFactor;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnMinX, mnMinY, mnMaxX, mnMaxY;               /* KeyFrame.h:236-239: const int in the reference */
    cv::Mat Rcw, tcw, Ow;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
};
}
#endif

/* oracle/slamshim/MapPoint.h — stub of ORB_SLAM2::MapPoint with the members ORBmatcher.cc reads (test infrastructure).
 * PredictScale / Get{Min,Max}DistanceInvariance are NOT restated here: the build recipe (oracle/Makefile, target
 * ref_matcher) takes their definitions verbatim from /root/reference/src/MapPoint.cc:395-439. */
#ifndef SLAMSHIM_MAPPOINT_H
#define SLAMSHIM_MAPPOINT_H
#include <map>
#include <mutex>
#include <opencv2/core/core.hpp>
#include "KeyFrame.h"
#include "Frame.h"
namespace ORB_SLAM2
{
class KeyFrame;
class Frame;
class MapPoint
{
public:
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0), mTrackViewCos(0),
                 mnLastFrameSeen(0), mbBad(false), nObs(0), mfMinDistance(0), mfMaxDistance(0), index(-1), fused(-1) {}
    cv::Mat GetWorldPos() { return mWorldPos; }
    cv::Mat GetNormal() { return mNormalVector; }
    cv::Mat GetDescriptor() { return mDescriptor; }
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    /* the map surgery of Fuse is not performed; KeyFrame::AddMapPoint records the pick */
    void AddObservation(KeyFrame*, size_t) {}
    void Replace(MapPoint*) {}
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);

    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    long unsigned int mnLastFrameSeen;
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    bool mbBad;
    int nObs;
    std::map<KeyFrame*, size_t> mObservations;
    float mfMinDistance, mfMaxDistance;
    std::mutex mMutexPos;
    int index;                                  /* position in the flattened point table of the glue */
    int fused;                                  /* keyframe feature Fuse added this point to (KeyFrame::AddMapPoint stub) */
};
}
#endif

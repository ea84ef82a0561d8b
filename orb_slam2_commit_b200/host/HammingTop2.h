// HammingTop2.h — batched replacement for the best / second-best loops around ORBmatcher::DescriptorDistance
// (reference: include/ORBmatcher.h:50, src/ORBmatcher.cc:84-126, :1844-1860). Header-only wrapper over liborbx.so.
#ifndef ORB_SLAM2_HAMMING_TOP2_H
#define ORB_SLAM2_HAMMING_TOP2_H

#include <vector>
#include <opencv2/core/core.hpp>
#include "../../include/orbx.h"

namespace ORB_SLAM2
{
// For every row of `queries` (CV_8U, 32 columns, continuous): first train index attaining the minimum distance,
// that distance and the second-smallest distance, i.e. (bestIdx, bestDist, bestDist2) of ORBmatcher.cc:84-126 when the
// candidate set is all of `train`. Returns false when the device call fails (orbx_last_error() has the reason).
inline bool HammingBestSecondBest(const cv::Mat& queries, const cv::Mat& train, std::vector<int>& bestIdx,
                                  std::vector<int>& bestDist, std::vector<int>& bestDist2, int device = 0)
{
    const int nq = queries.rows, nt = train.rows;
    bestIdx.assign(nq, -1); bestDist.assign(nq, 256); bestDist2.assign(nq, 256);
    if (nq == 0) return true;
    return orbx_hamming_top2(queries.data, nq, train.data, nt, bestIdx.data(), bestDist.data(), bestDist2.data(), device) == ORBX_OK;
}
} // namespace ORB_SLAM2
#endif

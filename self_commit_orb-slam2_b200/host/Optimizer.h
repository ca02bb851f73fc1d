// Optimizer.h — mirror of Optimizer::LocalBundleAdjustment (/root/reference/include/Optimizer.h:112) on a flattened
// window.  The adapter that walks KeyFrame/MapPoint (src/Optimizer.cc:633-854) and writes the results back under
// Map::mMutexMapUpdate (:961-996) is listed in INTEGRATION.md.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/b200slam.h"

namespace ORB_SLAM2 {

struct LocalBAWindow {
  int nLocal = 0;                    // lLocalKeyFrames.size(); local keyframes come first in Tcw/fixed
  std::vector<float> Tcw;            // nKF x 16 (KeyFrame::GetPose(), row-major)
  std::vector<uint8_t> fixed;        // nKF: mnId==0 or lFixedCameras
  std::vector<float> points;         // nMP x 3
  std::vector<b2s_ba_edge> edges;    // insertion order of :770-853
  float fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
};

struct LocalBAResult {
  std::vector<float> Tcw;            // nLocal x 16 -> KeyFrame::SetPose
  std::vector<float> points;         // nMP x 3    -> MapPoint::SetWorldPos
  std::vector<uint8_t> outlier;      // per edge   -> vToErase
  bool aborted = false;              // stop flag was set before round 1: nothing to write back (:858-860)
};

struct PoseProblem {                 // Optimizer::PoseOptimization(Frame*) on the flattened frame
  float Tcw[16];                     // pFrame->mTcw
  std::vector<uint8_t> hasMapPoint;  // N: mvpMapPoints[i] != NULL
  std::vector<float> Xw;             // N x 3: pMP->GetWorldPos() (ignored where hasMapPoint == 0)
  std::vector<float> kpx, kpy;       // N: mvKeysUn[i].pt
  std::vector<float> uRight;         // N: mvuRight[i] (< 0: monocular observation)
  std::vector<float> invSigma2;      // N: mvInvLevelSigma2[mvKeysUn[i].octave]
  float fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
};

class Optimizer {
 public:
  // int static PoseOptimization(Frame* pFrame) (include/Optimizer.h:100, src/Optimizer.cc:363-605): fills TcwOut (the
  // pose for pFrame->SetPose) and outlier (pFrame->mvbOutlier), returns nInitialCorrespondences - nBad
  static int PoseOptimization(const PoseProblem& f, float TcwOut[16], std::vector<uint8_t>& outlier);
  // void static LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap) on the flattened window
  static void LocalBundleAdjustment(const LocalBAWindow& w, bool* pbStopFlag, LocalBAResult& out);
};

}  // namespace ORB_SLAM2

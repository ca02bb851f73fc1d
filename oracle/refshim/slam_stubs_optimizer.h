// TEST INFRASTRUCTURE — data-holder stand-ins for Map / KeyFrame / MapPoint / Frame / LoopClosing so that the reference's
// src/Optimizer.cc and src/Converter.cc compile IN PLACE, unmodified, against the vendored g2o (also compiled in place) and
// the Eigen stand-in (refshim/eigen/refshim_eigen.h).  Pre-included with `-include`; it defines the reference headers'
// include guards so include/Optimizer.h opens the real Map.h / KeyFrame.h / ... and skips them.
//
// Only what Optimizer.cc touches is here; nothing has behaviour beyond storing / returning values.  Mutations the
// optimizer performs on the map (EraseMapPointMatch / EraseObservation) are recorded for the glue to read back.
//
// One intervention, for observability only: `OptimizationAlgorithmLevenberg` is aliased to a subclass that calls the
// unmodified g2o solve() (core/optimization_algorithm_levenberg.cpp:61-164, compiled in place) and then records how many
// LM trials that iteration took, whether the last one was accepted (ν == 2), λ and the robustified χ².  That is the
// accept / reject trace the oracle and the CUDA kernel are compared with.
#ifndef B2S_REF_SLAM_STUBS_OPTIMIZER_H
#define B2S_REF_SLAM_STUBS_OPTIMIZER_H
#define MAP_H
#define MAPPOINT_H
#define KEYFRAME_H
#define LOOPCLOSING_H
#define FRAME_H

#include <opencv2/core/core.hpp>

#include <list>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/sparse_optimizer.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"

using namespace std;  // the reference headers rely on it

namespace g2o {
struct B2sLmIteration {
  int trials;        // LM trials of this outer iteration (levenbergIteration())
  int last_accepted; // 1: the last trial was accepted (earlier ones were rejected)
  int result;        // OptimizationAlgorithm::SolverResult
  double lambda;     // currentLambda() after the iteration
  double chi2;       // activeRobustChi2() of the state the iteration left
};
inline std::vector<B2sLmIteration>& b2s_lm_trace() {
  static thread_local std::vector<B2sLmIteration> t;
  return t;
}
class B2sTracedLevenberg : public OptimizationAlgorithmLevenberg {
 public:
  explicit B2sTracedLevenberg(Solver* solver) : OptimizationAlgorithmLevenberg(solver) {}
  virtual SolverResult solve(int iteration, bool online = false) {
    SolverResult r = OptimizationAlgorithmLevenberg::solve(iteration, online);
    B2sLmIteration it;
    it.trials = _levenbergIterations;
    it.last_accepted = (_ni == 2.) ? 1 : 0;
    it.result = (int)r;
    it.lambda = _currentLambda;
    it.chi2 = 0;
    b2s_lm_trace().push_back(it);
    return r;
  }
};
}  // namespace g2o
#define OptimizationAlgorithmLevenberg B2sTracedLevenberg

namespace ORB_SLAM2 {

class KeyFrame;
class MapPoint;
class Map;

struct OptEraseLog {
  std::vector<std::pair<KeyFrame*, MapPoint*> > matches;       // KeyFrame::EraseMapPointMatch(pMP)
  std::vector<std::pair<MapPoint*, KeyFrame*> > observations;  // MapPoint::EraseObservation(pKF)
};
inline OptEraseLog& opt_erase_log() {
  static thread_local OptEraseLog l;
  return l;
}

class MapPoint {
 public:
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = ~0ul, mnBAGlobalForKF = ~0ul, mnCorrectedByKF = 0, mnCorrectedReference = 0;
  cv::Mat mWorldPos, mPosGBA;
  bool mbBad = false;
  std::map<KeyFrame*, size_t> mObservations;
  KeyFrame* mpRefKF = nullptr;
  int nUpdates = 0;
  static std::mutex mGlobalMutex;

  bool isBad() { return mbBad; }
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  void SetWorldPos(const cv::Mat& Pos) { Pos.copyTo(mWorldPos); }
  std::map<KeyFrame*, size_t> GetObservations() { return mObservations; }
  void UpdateNormalAndDepth() { ++nUpdates; }
  void EraseObservation(KeyFrame* pKF) { opt_erase_log().observations.push_back(std::make_pair(this, pKF)); }
  KeyFrame* GetReferenceKeyFrame() { return mpRefKF; }
  int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
};

class KeyFrame {
 public:
  long unsigned int mnId = 0;
  long unsigned int mnBALocalForKF = ~0ul, mnBAFixedForKF = ~0ul, mnBAGlobalForKF = ~0ul;
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
  cv::Mat mK, Tcw, mTcwGBA, mTcwBefGBA;
  bool mbBad = false;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight, mvInvLevelSigma2;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<KeyFrame*> mvCovisible;
  KeyFrame* mpParent = nullptr;
  std::set<KeyFrame*> mspChildren, mspLoopEdges;
  std::map<KeyFrame*, int> mWeights;

  bool isBad() { return mbBad; }
  cv::Mat GetPose() { return Tcw.clone(); }
  void SetPose(const cv::Mat& T) { T.copyTo(Tcw); }
  cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
  cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
  std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvCovisible; }
  std::vector<KeyFrame*> GetCovisiblesByWeight(const int& w) {
    std::vector<KeyFrame*> v;
    for (KeyFrame* k : mvCovisible)
      if (GetWeight(k) >= w) v.push_back(k);
    return v;
  }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  void EraseMapPointMatch(MapPoint* pMP) { opt_erase_log().matches.push_back(std::make_pair(this, pMP)); }
  KeyFrame* GetParent() { return mpParent; }
  bool hasChild(KeyFrame* k) { return mspChildren.count(k) != 0; }
  std::set<KeyFrame*> GetLoopEdges() { return mspLoopEdges; }
  int GetWeight(KeyFrame* k) { return mWeights.count(k) ? mWeights[k] : 0; }
};

class Frame {
 public:
  int N = 0;
  static float fx, fy, cx, cy;
  float mbf = 0;
  cv::Mat mTcw;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight, mvInvLevelSigma2;
  void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); }
};

class Map {
 public:
  std::mutex mMutexMapUpdate;
  std::vector<KeyFrame*> mvKFs;
  std::vector<MapPoint*> mvMPs;
  std::vector<KeyFrame*> GetAllKeyFrames() { return mvKFs; }
  std::vector<MapPoint*> GetAllMapPoints() { return mvMPs; }
  long unsigned int GetMaxKFid() {
    long unsigned int m = 0;
    for (KeyFrame* k : mvKFs) m = std::max(m, k->mnId);
    return m;
  }
};

class LoopClosing {
 public:
  // include/LoopClosing.h:62-66
  typedef map<KeyFrame*, g2o::Sim3, std::less<KeyFrame*>, Eigen::aligned_allocator<std::pair<const KeyFrame*, g2o::Sim3> > >
      KeyFrameAndPose;
};

}  // namespace ORB_SLAM2
#endif

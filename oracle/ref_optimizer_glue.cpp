// TEST INFRASTRUCTURE — drives the reference's OWN Optimizer::LocalBundleAdjustment / Optimizer::PoseOptimization
// (/root/reference/src/Optimizer.cc:363-605, 629-997, compiled in place, unmodified, together with the vendored g2o and
// src/Converter.cc) on the flattened problems the oracle and the CUDA path take (orb_oracle.h: orc_ba_problem,
// orc_pose_problem).  The SLAM objects are the data holders of refshim/slam_stubs_optimizer.h; Eigen is the stand-in of
// refshim/eigen/refshim_eigen.h (not in this image).  Built by `make -C oracle ref` into _ref/libref_optimizer.so.
//
// Mapping of a flattened window onto the object graph LocalBundleAdjustment walks:
//   * keyframe k -> KeyFrame with mnId = k, allocated in one array (std::map<KeyFrame*, size_t> iterates in index order);
//     keyframe 0 is the triggering pKF, keyframes [1, n_local) are its covisible neighbours, [n_local, n_kf) are only
//     reached through observations and therefore become lFixedCameras (:672-689); fixed[k] must be 1 exactly for k == 0
//     (the mnId == 0 rule, :722) and k >= n_local;
//   * map point m -> MapPoint with mnId = m; pKF's match list holds every point in index order so that lLocalMapPoints
//     (:650-668), hence the edge insertion order, equals the flattened edge order (points ascending, keyframes ascending);
//   * edge e -> keypoint number e's position inside its keyframe; mvInvLevelSigma2 of a keyframe is indexed by that same
//     number (octave = local keypoint index), so mvInvLevelSigma2[kpUn.octave] returns the edge's inv_sigma2.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "Optimizer.h"
#include "orb_oracle.h"

namespace ORB_SLAM2 {
std::mutex MapPoint::mGlobalMutex;
float Frame::fx = 0, Frame::fy = 0, Frame::cx = 0, Frame::cy = 0;
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

static cv::Mat mat44(const float* T) {
  cv::Mat m(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m.at<float>(i, j) = T[i * 4 + j];
  return m;
}

// The same glue also drives the product's drop-in bodies (self_commit_orb-slam2_b200/host/adapters/Optimizer_b200.cc): built
// with -DB2S_ADAPTER_BUILD into _ref/libadapter_optimizer.so, the entry points are adp_local_ba / adp_pose_optimization and
// Optimizer:: resolves to the adapter (which calls libb200slam.so) instead of the reference's Optimizer.cc + g2o.
#ifdef B2S_ADAPTER_BUILD
#define ref_local_ba adp_local_ba
#define ref_pose_optimization adp_pose_optimization
#define ref_lm_iterations adp_lm_iterations
#endif

// LM iterations of this thread -> accept(1) / reject(0) per trial, -1 terminated
static int flatten_trace(int32_t* trace, int cap) {
  std::vector<g2o::B2sLmIteration>& it = g2o::b2s_lm_trace();
  int n = 0;
  for (size_t i = 0; i < it.size(); i++)
    for (int k = 0; k < it[i].trials; k++) {
      int acc = (k == it[i].trials - 1) ? it[i].last_accepted : 0;
      if (trace && n < cap - 1) trace[n] = acc;
      n++;
    }
  if (trace) trace[n < cap - 1 ? n : cap - 1] = -1;
  return n;
}

extern "C" {

// per outer LM iteration (both rounds, in order): trials, accepted flag of the last trial, SolverResult, lambda after it
int ref_lm_iterations(int32_t* trials, int32_t* last_accepted, int32_t* result, double* lambda, int cap) {
  std::vector<g2o::B2sLmIteration>& it = g2o::b2s_lm_trace();
  int n = (int)it.size();
  for (int i = 0; i < n && i < cap; i++) {
    if (trials) trials[i] = it[i].trials;
    if (last_accepted) last_accepted[i] = it[i].last_accepted;
    if (result) result[i] = it[i].result;
    if (lambda) lambda[i] = it[i].lambda;
  }
  return n;
}

// 0 ok, 1 aborted before round 1 (no write-back, src/Optimizer.cc:858-860), < 0 the window cannot be expressed
int ref_local_ba(const orc_ba_problem* p, const volatile uint8_t* stop, orc_ba_result* r) {
  if (p->its1 != 5 || p->its2 != 10) return -2;  // hard-coded in the reference (:863, :915)
  for (int k = 0; k < p->n_kf; k++) {
    const bool want = (k == 0) || (k >= p->n_local);
    if ((p->fixed[k] != 0) != want) return -3;
  }
  std::vector<KeyFrame> kfs((size_t)p->n_kf);
  std::vector<MapPoint> mps((size_t)p->n_mp);
  Map map;
  for (int k = 0; k < p->n_kf; k++) {
    KeyFrame& kf = kfs[(size_t)k];
    kf.mnId = (unsigned long)k;
    kf.fx = p->fx;
    kf.fy = p->fy;
    kf.cx = p->cx;
    kf.cy = p->cy;
    kf.mbf = p->bf;
    kf.Tcw = mat44(p->Tcw + 16 * k);
  }
  for (int k = 1; k < p->n_local; k++) kfs[0].mvCovisible.push_back(&kfs[(size_t)k]);
  for (int m = 0; m < p->n_mp; m++) {
    MapPoint& mp = mps[(size_t)m];
    mp.mnId = (unsigned long)m;
    mp.mWorldPos = cv::Mat(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) mp.mWorldPos.at<float>(i) = p->points[3 * m + i];
    kfs[0].mvpMapPoints.push_back(&mp);
  }
  std::map<std::pair<int, int>, int> edge_of;
  for (int e = 0; e < p->n_edges; e++) {
    const orc_ba_edge& ed = p->edges[e];
    if (ed.kf < 0 || ed.kf >= p->n_kf || ed.mp < 0 || ed.mp >= p->n_mp) return -4;
    // edges grouped by point, points ascending (the reference then walks a point's observations in keyframe order, whatever
    // their order here: that only permutes floating-point sums); one observation per (keyframe, point)
    if (e > 0 && ed.mp < p->edges[e - 1].mp) return -5;
    if (edge_of.count(std::make_pair(ed.kf, ed.mp))) return -6;
    KeyFrame& kf = kfs[(size_t)ed.kf];
    const size_t idx = kf.mvKeysUn.size();
    cv::KeyPoint kp(ed.obs[0], ed.obs[1], 31.f);
    kp.octave = (int)idx;
    kf.mvKeysUn.push_back(kp);
    kf.mvuRight.push_back(ed.obs[2]);
    kf.mvInvLevelSigma2.push_back(ed.inv_sigma2);
    mps[(size_t)ed.mp].mObservations[&kf] = idx;
    edge_of[std::make_pair(ed.kf, ed.mp)] = e;
  }
  g2o::b2s_lm_trace().clear();
  opt_erase_log().matches.clear();
  opt_erase_log().observations.clear();
  int n_updates_before = p->n_mp ? mps[0].nUpdates : 0;

  Optimizer::LocalBundleAdjustment(&kfs[0], (bool*)const_cast<uint8_t*>(stop), &map);

  const bool wrote_back = p->n_mp ? (mps[0].nUpdates != n_updates_before) : true;
  if (!wrote_back) return 1;
  for (int k = 0; k < p->n_local; k++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) r->Tcw_out[16 * k + 4 * i + j] = kfs[(size_t)k].Tcw.at<float>(i, j);
  for (int m = 0; m < p->n_mp; m++)
    for (int i = 0; i < 3; i++) r->points_out[3 * m + i] = mps[(size_t)m].mWorldPos.at<float>(i);
  std::memset(r->edge_outlier, 0, (size_t)p->n_edges);
  for (size_t i = 0; i < opt_erase_log().matches.size(); i++) {
    const int k = (int)(opt_erase_log().matches[i].first - &kfs[0]);
    const int m = (int)(opt_erase_log().matches[i].second - &mps[0]);
    r->edge_outlier[edge_of[std::make_pair(k, m)]] = 1;
  }
  r->n_trials = flatten_trace(r->trace, 256);
  r->chi2_final = std::nan("");
  return 0;
}

// returns nInitialCorrespondences - nBad (src/Optimizer.cc:605)
int ref_pose_optimization(const orc_pose_problem* p, orc_pose_result* r) {
  std::vector<MapPoint> mps((size_t)p->n);
  Frame f;
  f.N = p->n;
  Frame::fx = p->fx;
  Frame::fy = p->fy;
  Frame::cx = p->cx;
  Frame::cy = p->cy;
  f.mbf = p->bf;
  f.mTcw = mat44(p->Tcw);
  f.mvpMapPoints.assign((size_t)p->n, nullptr);
  f.mvbOutlier.assign((size_t)p->n, false);
  for (int i = 0; i < p->n; i++) {
    cv::KeyPoint kp(p->kpx[i], p->kpy[i], 31.f);
    kp.octave = i;
    f.mvKeysUn.push_back(kp);
    f.mvuRight.push_back(p->uright[i]);
    f.mvInvLevelSigma2.push_back(p->inv_sigma2[i]);
    if (p->has_mp[i]) {
      mps[(size_t)i].mWorldPos = cv::Mat(3, 1, CV_32F);
      for (int c = 0; c < 3; c++) mps[(size_t)i].mWorldPos.at<float>(c) = p->Xw[3 * i + c];
      f.mvpMapPoints[(size_t)i] = &mps[(size_t)i];
    }
  }
  g2o::b2s_lm_trace().clear();
  const int inliers = Optimizer::PoseOptimization(&f);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r->Tcw_out[4 * i + j] = f.mTcw.at<float>(i, j);
  for (int i = 0; i < p->n; i++) r->outlier[i] = f.mvbOutlier[(size_t)i] ? 1 : 0;
  r->n_trials = flatten_trace(r->trace, 256);
  return inliers;
}

}  // extern "C"

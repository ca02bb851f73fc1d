// Optimizer_b200.cc — drop-in bodies of
//     void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap)   include/Optimizer.h:112
//     int  Optimizer::PoseOptimization(Frame* pFrame)                                      include/Optimizer.h:100
// with the reference's exact signatures, for the LocalMapping / Tracking threads to call unchanged
// (src/LocalMapping.cc:123, src/Tracking.cc:1219, 1398, 1834, 2078, 2152 ...).  Compile this file INSIDE the reference
// tree (it includes the reference's own include/Optimizer.h) in place of those two function bodies of src/Optimizer.cc
// (:363-605, :629-997) and link libb200slam.so; the other Optimizer functions (global BA, essential graph, Sim3) are
// outside the hot path and stay as they are.
//
// What stays on the host is what touches the object graph: the window selection (local keyframes = pKF + covisibles,
// local map points = their matches, fixed cameras = other observers; the mnBALocalForKF / mnBAFixedForKF marks other
// threads rely on), and the write-back under Map::mMutexMapUpdate (erase the outlier observations, SetPose,
// SetWorldPos, UpdateNormalAndDepth).  Everything numeric — the graph, 5 + 10 Levenberg iterations, the outlier
// re-classification — is b2s_local_ba / b2s_pose_optimization on the GPU.  No exceptions, no return codes beyond the
// reference's: a library error is reported on stderr and the map is left untouched.
//
// tests: oracle/ref_optimizer_glue.cpp builds the same stub object graph for this file (oracle/_ref/libadapter_optimizer.so)
// and for the reference's own Optimizer.cc (libref_optimizer.so); tests/test_adapters_gpu.py requires identical erase
// lists and poses / landmarks within the parity bar.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <map>
#include <mutex>
#include <vector>

#include "Optimizer.h"

#include "b200slam.h"

namespace ORB_SLAM2 {
namespace {

int device_index() {
  const char* e = getenv("B2S_DEVICE");
  return e ? atoi(e) : 0;
}

// One solver handle per calling thread role, grown on demand and kept for the process lifetime (handles own device
// buffers and a stream; creating one per call would put cudaMalloc / cudaFree on the mapping thread's critical path).
struct SolverSlot {
  b2s_ba_solver* h = nullptr;
  int kf = 0, mp = 0, edges = 0;
  std::mutex m;
  b2s_ba_solver* get(int nKF, int nMP, int nE) {
    if (h && nKF <= kf && nMP <= mp && nE <= edges) return h;
    b2s_ba_destroy(h);
    h = nullptr;
    kf = nKF > 64 ? nKF : 64;
    mp = nMP > 8192 ? nMP : 8192;
    edges = nE > 65536 ? nE : 65536;
    if (b2s_ba_create(kf, mp, edges, 1, device_index(), &h) != B2S_OK) {
      fprintf(stderr, "Optimizer (b200): %s\n", b2s_last_error());
      h = nullptr;
      kf = mp = edges = 0;
    }
    return h;
  }
};
SolverSlot g_localBA, g_pose;

void pose_to_floats(const cv::Mat& T, float* out) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) out[4 * i + j] = T.at<float>(i, j);
}

}  // namespace

void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap) {
  // ---- window selection (what src/Optimizer.cc:633-694 leaves behind: three lists and the marks on the objects)
  std::vector<KeyFrame*> localKFs(1, pKF);
  pKF->mnBALocalForKF = pKF->mnId;
  {
    const std::vector<KeyFrame*> neigh = pKF->GetVectorCovisibleKeyFrames();
    for (size_t i = 0; i < neigh.size(); i++) {
      neigh[i]->mnBALocalForKF = pKF->mnId;
      if (!neigh[i]->isBad()) localKFs.push_back(neigh[i]);
    }
  }
  std::vector<MapPoint*> localMPs;
  for (size_t k = 0; k < localKFs.size(); k++) {
    const std::vector<MapPoint*> matches = localKFs[k]->GetMapPointMatches();
    for (size_t i = 0; i < matches.size(); i++) {
      MapPoint* p = matches[i];
      if (!p || p->isBad() || p->mnBALocalForKF == pKF->mnId) continue;
      p->mnBALocalForKF = pKF->mnId;
      localMPs.push_back(p);
    }
  }
  std::vector<KeyFrame*> fixedKFs;
  std::vector<std::map<KeyFrame*, size_t> > observations(localMPs.size());
  for (size_t m = 0; m < localMPs.size(); m++) {
    observations[m] = localMPs[m]->GetObservations();
    for (std::map<KeyFrame*, size_t>::const_iterator it = observations[m].begin(); it != observations[m].end(); ++it) {
      KeyFrame* k = it->first;
      if (k->mnBALocalForKF == pKF->mnId || k->mnBAFixedForKF == pKF->mnId) continue;
      k->mnBAFixedForKF = pKF->mnId;
      if (!k->isBad()) fixedKFs.push_back(k);
    }
  }
  if (pbStopFlag && *pbStopFlag) return;  // :858-860 (nothing has been changed yet)

  // ---- flatten: local keyframes first, then the fixed cameras; one edge per observation of a non-bad keyframe
  std::map<KeyFrame*, int> kfIndex;
  std::vector<KeyFrame*> allKFs(localKFs);
  allKFs.insert(allKFs.end(), fixedKFs.begin(), fixedKFs.end());
  const int nKF = (int)allKFs.size(), nLocal = (int)localKFs.size(), nMP = (int)localMPs.size();
  std::vector<float> Tcw((size_t)nKF * 16), pts((size_t)nMP * 3);
  std::vector<uint8_t> fixed((size_t)nKF, 0);
  for (int k = 0; k < nKF; k++) {
    kfIndex[allKFs[k]] = k;
    pose_to_floats(allKFs[k]->GetPose(), &Tcw[(size_t)16 * k]);
    fixed[k] = (k >= nLocal) || (allKFs[k]->mnId == 0);  // :722, :736
  }
  std::vector<b2s_ba_edge> edges;
  std::vector<std::pair<KeyFrame*, MapPoint*> > edgeOwner;
  for (int m = 0; m < nMP; m++) {
    const cv::Mat X = localMPs[m]->GetWorldPos();
    for (int c = 0; c < 3; c++) pts[(size_t)3 * m + c] = X.at<float>(c);
    for (std::map<KeyFrame*, size_t>::const_iterator it = observations[m].begin(); it != observations[m].end(); ++it) {
      KeyFrame* k = it->first;
      if (k->isBad()) continue;
      std::map<KeyFrame*, int>::const_iterator ki = kfIndex.find(k);
      if (ki == kfIndex.end()) continue;  // (cannot happen: every non-bad observer is local or fixed)
      const cv::KeyPoint& kp = k->mvKeysUn[it->second];
      b2s_ba_edge e;
      e.kf = ki->second;
      e.mp = m;
      e.obs[0] = kp.pt.x;
      e.obs[1] = kp.pt.y;
      e.obs[2] = k->mvuRight[it->second];  // < 0: monocular observation (:794)
      e.inv_sigma2 = k->mvInvLevelSigma2[kp.octave];
      edges.push_back(e);
      edgeOwner.push_back(std::make_pair(k, localMPs[m]));
    }
  }
  const int nE = (int)edges.size();
  std::vector<float> TcwOut((size_t)nLocal * 16), ptsOut((size_t)nMP * 3);
  std::vector<uint8_t> outlier((size_t)(nE > 0 ? nE : 1), 0);
  b2s_ba_problem prob;
  prob.n_kf = nKF;
  prob.n_local = nLocal;
  prob.Tcw = Tcw.data();
  prob.fixed = fixed.data();
  prob.n_mp = nMP;
  prob.points = pts.data();
  prob.n_edges = nE;
  prob.edges = edges.data();
  prob.fx = pKF->fx;
  prob.fy = pKF->fy;
  prob.cx = pKF->cx;
  prob.cy = pKF->cy;
  prob.bf = pKF->mbf;
  prob.its1 = 5;   // optimizer.optimize(5)  (:864)
  prob.its2 = 10;  // optimizer.optimize(10) (:917)
  b2s_ba_result res;
  res.Tcw_out = TcwOut.data();
  res.points_out = ptsOut.data();
  res.edge_outlier = outlier.data();
  res.trace = nullptr;
  res.chi2_final = 0;
  res.n_trials = 0;
  {
    std::lock_guard<std::mutex> lock(g_localBA.m);
    b2s_ba_solver* solver = g_localBA.get(nKF, nMP, nE);
    if (!solver) return;
    static_assert(sizeof(bool) == 1, "the stop flag is polled as one byte");
    const int rc = b2s_local_ba(solver, &prob, reinterpret_cast<const volatile uint8_t*>(pbStopFlag), &res);
    if (rc == B2S_ERR_ABORTED) return;
    if (rc != B2S_OK) {
      fprintf(stderr, "Optimizer::LocalBundleAdjustment (b200): %s\n", b2s_last_error());
      return;
    }
  }

  // ---- write-back (:921-996): monocular outliers first, then the stereo ones, each in edge order
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
  for (int pass = 0; pass < 2; pass++)
    for (int e = 0; e < nE; e++) {
      const bool stereo = edges[e].obs[2] >= 0;
      if (!outlier[e] || stereo != (pass == 1)) continue;
      edgeOwner[e].first->EraseMapPointMatch(edgeOwner[e].second);
      edgeOwner[e].second->EraseObservation(edgeOwner[e].first);
    }
  for (int k = 0; k < nLocal; k++) {
    cv::Mat T(4, 4, CV_32F);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) T.at<float>(i, j) = TcwOut[(size_t)16 * k + 4 * i + j];
    localKFs[k]->SetPose(T);
  }
  for (int m = 0; m < nMP; m++) {
    cv::Mat X(3, 1, CV_32F);
    for (int c = 0; c < 3; c++) X.at<float>(c) = ptsOut[(size_t)3 * m + c];
    localMPs[m]->SetWorldPos(X);
    localMPs[m]->UpdateNormalAndDepth();
  }
}

int Optimizer::PoseOptimization(Frame* pFrame) {
  const int N = pFrame->N;
  std::vector<uint8_t> hasMP((size_t)(N > 0 ? N : 1), 0), outl((size_t)(N > 0 ? N : 1), 0);
  std::vector<float> Xw((size_t)3 * (N > 0 ? N : 1), 0.f), kpx((size_t)(N > 0 ? N : 1)), kpy(kpx.size()), ur(kpx.size()),
      w(kpx.size());
  float Tcw[16], TcwOut[16];
  pose_to_floats(pFrame->mTcw, Tcw);
  int nCorr = 0;
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);  // :410
    for (int i = 0; i < N; i++) {
      MapPoint* p = pFrame->mvpMapPoints[i];
      const cv::KeyPoint& kp = pFrame->mvKeysUn[i];
      kpx[i] = kp.pt.x;
      kpy[i] = kp.pt.y;
      ur[i] = pFrame->mvuRight[i];
      w[i] = pFrame->mvInvLevelSigma2[kp.octave];
      if (!p) continue;
      hasMP[i] = 1;
      nCorr++;
      pFrame->mvbOutlier[i] = false;  // :421, :458
      const cv::Mat X = p->GetWorldPos();
      for (int c = 0; c < 3; c++) Xw[(size_t)3 * i + c] = X.at<float>(c);
    }
  }
  if (nCorr < 3) return 0;  // :492-493 (pose untouched)
  b2s_pose_problem prob;
  prob.Tcw = Tcw;
  prob.n = N;
  prob.has_mp = hasMP.data();
  prob.Xw = Xw.data();
  prob.kpx = kpx.data();
  prob.kpy = kpy.data();
  prob.uright = ur.data();
  prob.inv_sigma2 = w.data();
  prob.fx = pFrame->fx;
  prob.fy = pFrame->fy;
  prob.cx = pFrame->cx;
  prob.cy = pFrame->cy;
  prob.bf = pFrame->mbf;
  b2s_pose_result res;
  res.Tcw_out = TcwOut;
  res.outlier = outl.data();
  res.trace = nullptr;
  res.n_inliers = 0;
  res.n_trials = 0;
  {
    std::lock_guard<std::mutex> lock(g_pose.m);
    b2s_ba_solver* solver = g_pose.get(4, 16, 64);
    if (!solver) return 0;
    if (b2s_pose_optimization(solver, &prob, &res) != B2S_OK) {
      fprintf(stderr, "Optimizer::PoseOptimization (b200): %s\n", b2s_last_error());
      return 0;
    }
  }
  for (int i = 0; i < N; i++)
    if (hasMP[i]) pFrame->mvbOutlier[i] = outl[i] != 0;
  cv::Mat T(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) T.at<float>(i, j) = TcwOut[4 * i + j];
  pFrame->SetPose(T);
  return res.n_inliers;
}

}  // namespace ORB_SLAM2

#include "Optimizer.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>

namespace ORB_SLAM2 {

// LocalBA runs on the LocalMapping thread only (src/LocalMapping.cc:123): one solver, grown on demand.
static b2s_ba_solver* g_solver = nullptr;
static int g_kf = 0, g_mp = 0, g_e = 0;
static std::mutex g_mutex;

void Optimizer::LocalBundleAdjustment(const LocalBAWindow& w, bool* pbStopFlag, LocalBAResult& out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  const int nKF = (int)w.fixed.size(), nMP = (int)w.points.size() / 3, nE = (int)w.edges.size();
  if (!g_solver || nKF > g_kf || nMP > g_mp || nE > g_e) {
    b2s_ba_destroy(g_solver);
    g_solver = nullptr;
    g_kf = nKF > 64 ? nKF : 64;
    g_mp = nMP > 8192 ? nMP : 8192;
    g_e = nE > 65536 ? nE : 65536;
    const char* e = getenv("B2S_DEVICE");
    int rc = b2s_ba_create(g_kf, g_mp, g_e, 1, e ? atoi(e) : 0, &g_solver);
    if (rc != B2S_OK) {
      fprintf(stderr, "Optimizer::LocalBundleAdjustment: libb200slam error %d: %s\n", rc, b2s_last_error());
      throw std::runtime_error(b2s_last_error());
    }
  }
  b2s_ba_problem p;
  p.n_kf = nKF;
  p.n_local = w.nLocal;
  p.Tcw = w.Tcw.data();
  p.fixed = w.fixed.data();
  p.n_mp = nMP;
  p.points = w.points.data();
  p.n_edges = nE;
  p.edges = w.edges.data();
  p.fx = w.fx; p.fy = w.fy; p.cx = w.cx; p.cy = w.cy; p.bf = w.bf;
  p.its1 = 5;   // optimizer.optimize(5)  :864
  p.its2 = 10;  // optimizer.optimize(10) :917
  out.Tcw.assign((size_t)w.nLocal * 16, 0.f);
  out.points.assign((size_t)nMP * 3, 0.f);
  out.outlier.assign(nE, 0);
  out.aborted = false;
  b2s_ba_result r;
  r.Tcw_out = out.Tcw.data();
  r.points_out = out.points.data();
  r.edge_outlier = out.outlier.data();
  r.trace = nullptr;
  r.chi2_final = 0;
  r.n_trials = 0;
  static_assert(sizeof(bool) == 1, "bool stop flag is read as a byte");
  int rc = b2s_local_ba(g_solver, &p, (const volatile uint8_t*)pbStopFlag, &r);
  if (rc == B2S_ERR_ABORTED) {
    out.aborted = true;
    return;
  }
  if (rc != B2S_OK) {
    fprintf(stderr, "Optimizer::LocalBundleAdjustment: libb200slam error %d: %s\n", rc, b2s_last_error());
    throw std::runtime_error(b2s_last_error());
  }
}

// PoseOptimization runs on the Tracking thread: its own solver handle (handles are not shared between threads)
static b2s_ba_solver* g_poseSolver = nullptr;
static std::mutex g_poseMutex;

int Optimizer::PoseOptimization(const PoseProblem& f, float TcwOut[16], std::vector<uint8_t>& outlier) {
  std::lock_guard<std::mutex> lock(g_poseMutex);
  if (!g_poseSolver) {
    const char* e = getenv("B2S_DEVICE");
    int rc = b2s_ba_create(4, 16, 64, 1, e ? atoi(e) : 0, &g_poseSolver);
    if (rc != B2S_OK) {
      fprintf(stderr, "Optimizer::PoseOptimization: libb200slam error %d: %s\n", rc, b2s_last_error());
      throw std::runtime_error(b2s_last_error());
    }
  }
  const int N = (int)f.hasMapPoint.size();
  outlier.assign(N > 0 ? N : 1, 0);
  b2s_pose_problem p;
  p.Tcw = f.Tcw; p.n = N; p.has_mp = f.hasMapPoint.data(); p.Xw = f.Xw.data(); p.kpx = f.kpx.data(); p.kpy = f.kpy.data();
  p.uright = f.uRight.data(); p.inv_sigma2 = f.invSigma2.data();
  p.fx = f.fx; p.fy = f.fy; p.cx = f.cx; p.cy = f.cy; p.bf = f.bf;
  b2s_pose_result r;
  r.Tcw_out = TcwOut; r.outlier = outlier.data(); r.trace = nullptr; r.n_inliers = 0; r.n_trials = 0;
  int rc = b2s_pose_optimization(g_poseSolver, &p, &r);
  if (rc != B2S_OK) {
    fprintf(stderr, "Optimizer::PoseOptimization: libb200slam error %d: %s\n", rc, b2s_last_error());
    throw std::runtime_error(b2s_last_error());
  }
  outlier.resize(N);
  return r.n_inliers;
}

}  // namespace ORB_SLAM2

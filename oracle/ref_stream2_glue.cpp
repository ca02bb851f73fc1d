// BENCH INFRASTRUCTURE — the CPU arm of bench.py (`--impl reference`, `cpu_baseline`): one step of the stream workload on
// the host cores with the REFERENCE'S OWN code for every stage:
//   * per stereo frame the reference's stereo Frame constructor (src/Frame.cc:343-458, compiled in place): two ORBextractor
//     threads (src/ORBextractor.cc) + Frame::ComputeStereoMatches + AssignFeaturesToGrid;
//   * temporal matching of every left image against its predecessor with the reference's ORBmatcher::SearchByBoW
//     (KeyFrame*, Frame&) (src/ORBmatcher.cc:230), one vocabulary node = the 2000 x 2000 brute-force case;
//   * optionally (cfg.project) the motion-model matcher of Tracking::TrackWithMotionModel: every stereo point of a frame
//     becomes a MapPoint at Frame::UnprojectStereo (what Tracking::UpdateLastFrame / StereoInitialization do), the next
//     frame gets the pose cfg.Tcl and the reference's ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
//     (src/ORBmatcher.cc:1569) runs on a per-task copy of it (Tracking copies the frame once per frame as well);
//   * Optimizer::PoseOptimization per frame and Optimizer::LocalBundleAdjustment per window through function pointers
//     (the reference's own Optimizer.cc + g2o live in libref_optimizer.so because their stand-in Map / KeyFrame objects
//     differ from the ones Frame.cc needs; bench.py passes ref_pose_optimization / ref_local_ba, or the oracle port when
//     that is faster on this box — whichever is passed is reported).
// Work is dealt dynamically to `threads` host threads, longest tasks first; every frame task runs the two extractor
// threads the reference itself spawns (src/Frame.cc:159-167).  Per-stage busy time is accumulated so that the bench line
// can report utilisation and per-stage milliseconds.  Built into oracle/_ref/libref_stream2.so; never part of the product.
#include <malloc.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "Frame.h"
#include "ORBmatcher.h"
#include "orb_oracle.h"

using namespace ORB_SLAM2;

namespace {
typedef int (*ba_fn_t)(const orc_ba_problem*, const volatile uint8_t*, orc_ba_result*);
typedef int (*pose_fn_t)(const orc_pose_problem*, orc_pose_result*);
ORBVocabulary g_voc;
std::mutex g_exMutex;
std::vector<std::unique_ptr<ORBextractor> > g_extractors;  // two per worker thread, kept across steps

cv::Mat mat_f(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) m.at<float>(i, j) = p[i * c + j];
  return m;
}
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

extern "C" {

typedef struct {
  int nfeatures, nlevels, iniTh, minTh;
  float scaleFactor, fx, fy, cx, cy, bf, thDepth;
  int project;     /* 1: also run SearchByProjection(CurrentFrame, LastFrame) per frame */
  float projTh;    /* its window (7 for stereo, src/Tracking.cc:892-897) */
  float Tcl[12];   /* pose of the current camera relative to the last one, 3 x 4 row major */
} ref_stream2_cfg;

/* imgs: S left images then S right images (dense, w*h each).  ba / pose: arrays of problems (nBa windows, nPose frames;
 * either may be 0).  stats (8 doubles): busy seconds in [0] frame construction (extract L+R + stereo), [1] SearchByBoW,
 * [2] PoseOptimization, [3] LocalBA; [4] keypoints per left image (mean), [5] stereo matches per frame (mean),
 * [6] BoW matches per frame (mean), [7] wall seconds, [8] busy seconds in SearchByProjection(Cur, Last) (+ the MapPoint
 * creation it needs), [9] its matches per frame (mean); 12 doubles in all.  Returns the wall time of the step. */
double ref_stream2_step(const ref_stream2_cfg* cfg, const uint8_t* imgs, int S, int w, int h, int threads,
                        const orc_ba_problem* ba, int nBa, ba_fn_t ba_fn, const orc_pose_problem* pose, int nPose,
                        pose_fn_t pose_fn, double* stats) {
  if (threads < 1) threads = 1;
  {
    // The reference allocates its pyramids / descriptor matrices per frame; with one frame per host thread in flight, glibc's
    // default of serving every block above 128 KB with mmap / munmap serialises all threads on the kernel's address-space
    // lock (measured: 5-9x longer tasks at 128 threads).  Keep big blocks in the per-thread arenas instead.
    static bool tuned = false;
    if (!tuned) {
      mallopt(M_MMAP_THRESHOLD, 1 << 30);
      mallopt(M_TRIM_THRESHOLD, 1 << 30);
      tuned = true;
    }
  }
  {
    std::lock_guard<std::mutex> lock(g_exMutex);
    while ((int)g_extractors.size() < 2 * threads)
      g_extractors.emplace_back(new ORBextractor(cfg->nfeatures, cfg->scaleFactor, cfg->nlevels, cfg->iniTh, cfg->minTh));
  }
  const float k[9] = {cfg->fx, 0, cfg->cx, 0, cfg->fy, cfg->cy, 0, 0, 1};
  const float d[4] = {0, 0, 0, 0};
  cv::Mat K = mat_f(k, 3, 3), D = mat_f(d, 4, 1);  // (the constructor takes non-const references; each Frame clones them)
  std::vector<std::unique_ptr<Frame> > frames((size_t)S);
  const size_t img = (size_t)w * h;
  if (Frame::mbInitialComputations && S > 0) {  // the image bounds / grid constants are computed by the first Frame (:424-446)
    cv::Mat imL(h, w, CV_8UC1, (void*)imgs, (size_t)w), imR(h, w, CV_8UC1, (void*)(imgs + (size_t)S * img), (size_t)w);
    Frame warm(imL, imR, 0.0, g_extractors[0].get(), g_extractors[1].get(), &g_voc, K, D, cfg->bf, cfg->thDepth);
  }
  std::vector<double> busy((size_t)threads * 5, 0.0);
  std::vector<int> nBow((size_t)S, 0), nProj((size_t)S, 0);
  std::vector<std::vector<MapPoint> > framePts((size_t)S);
  const bool project = cfg->project != 0;
  cv::Mat Tcl = cv::Mat::eye(4, 4, CV_32F), Tid = cv::Mat::eye(4, 4, CV_32F);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) Tcl.at<float>(i, j) = cfg->Tcl[i * 4 + j];
  const double t0 = now();
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t]() {
        for (;;) {
          const int task = next.fetch_add(1);
          if (task >= nBa + S) break;
          const double a = now();
          if (task < nBa) {  // longest tasks first
            const orc_ba_problem& P = ba[task];
            std::vector<float> T((size_t)P.n_local * 16), X((size_t)P.n_mp * 3);
            std::vector<uint8_t> o((size_t)P.n_edges);
            orc_ba_result r;
            std::memset(&r, 0, sizeof(r));
            r.Tcw_out = T.data();
            r.points_out = X.data();
            r.edge_outlier = o.data();
            ba_fn(&P, nullptr, &r);
            busy[(size_t)t * 5 + 3] += now() - a;
          } else {
            const int f = task - nBa;
            cv::Mat imL(h, w, CV_8UC1, (void*)(imgs + (size_t)f * img), (size_t)w);
            cv::Mat imR(h, w, CV_8UC1, (void*)(imgs + (size_t)(S + f) * img), (size_t)w);
            frames[(size_t)f].reset(new Frame(imL, imR, (double)f, g_extractors[2 * t].get(), g_extractors[2 * t + 1].get(),
                                              &g_voc, K, D, cfg->bf, cfg->thDepth));
            Frame& fr = *frames[(size_t)f];
            // every feature in vocabulary node 0 (the 2000 x 2000 brute-force case); set here so that the matching phase only reads
            std::vector<unsigned int>& allF = fr.mFeatVec[0];
            allF.resize((size_t)fr.N);
            for (int i = 0; i < fr.N; i++) allF[(size_t)i] = (unsigned)i;
            const double b = now();
            busy[(size_t)t * 5 + 0] += b - a;
            if (project) {  // the frame as a LAST frame: pose = origin, one MapPoint per stereo point
              fr.SetPose(Tid);
              std::vector<MapPoint>& pts = framePts[(size_t)f];
              pts.resize((size_t)fr.N);
              for (int i = 0; i < fr.N; i++) {
                if (!(fr.mvDepth[(size_t)i] > 0)) continue;
                MapPoint& p = pts[(size_t)i];
                p.mWorldPos = fr.UnprojectStereo(i);
                p.mDescriptor = fr.mDescriptors.row(i).clone();
                p.nObs = 1;
                fr.mvpMapPoints[(size_t)i] = &p;
              }
              busy[(size_t)t * 5 + 4] += now() - b;
            }
          }
        }
      });
    for (auto& t : th) t.join();
  }
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t]() {
        for (;;) {
          const int task = next.fetch_add(1);
          if (task >= nPose + S + (project ? S : 0)) break;
          const double a = now();
          if (task < nPose) {
            const orc_pose_problem& P = pose[task];
            float T[16];
            std::vector<uint8_t> o((size_t)(P.n > 0 ? P.n : 1));
            orc_pose_result r;
            std::memset(&r, 0, sizeof(r));
            r.Tcw_out = T;
            r.outlier = o.data();
            pose_fn(&P, &r);
            busy[(size_t)t * 5 + 2] += now() - a;
          } else if (task >= nPose + S) {
            const int f = task - nPose - S;
            Frame cur(*frames[(size_t)f]);  // own copy: the shared one is some other task's last frame
            const Frame& last = *frames[(size_t)((f + S - 1) % S)];
            std::fill(cur.mvpMapPoints.begin(), cur.mvpMapPoints.end(), static_cast<MapPoint*>(NULL));
            cur.SetPose(Tcl);
            ORBmatcher m(0.9f, true);
            nProj[(size_t)f] = m.SearchByProjection(cur, last, cfg->projTh, false);
            busy[(size_t)t * 5 + 4] += now() - a;
          } else {
            const int f = task - nPose;
            Frame& cur = *frames[(size_t)f];
            Frame& prev = *frames[(size_t)((f + S - 1) % S)];
            KeyFrame kf;  // the predecessor plays the reference keyframe: every feature holds a valid map point
            kf.N = prev.N;
            kf.mvKeysUn = prev.mvKeysUn;
            kf.mvKeys = prev.mvKeys;
            kf.mDescriptors = prev.mDescriptors.clone();  // own buffer: cv::Mat::row() bumps the buffer's atomic reference
                                                          // count per call, which two tasks must not share
            std::vector<MapPoint> pts((size_t)prev.N);
            kf.mvpMapPoints.resize((size_t)prev.N);
            std::vector<unsigned int>& allK = kf.mFeatVec[0];
            allK.resize((size_t)prev.N);
            for (int i = 0; i < prev.N; i++) {
              kf.mvpMapPoints[(size_t)i] = &pts[(size_t)i];
              allK[(size_t)i] = (unsigned)i;
            }
            ORBmatcher m(0.7f, true);
            std::vector<MapPoint*> matches;
            nBow[(size_t)f] = m.SearchByBoW(&kf, cur, matches);
            busy[(size_t)t * 5 + 1] += now() - a;
          }
        }
      });
    for (auto& t : th) t.join();
  }
  const double wall = now() - t0;
  if (stats) {
    for (int s = 0; s < 4; s++) {
      stats[s] = 0;
      for (int t = 0; t < threads; t++) stats[s] += busy[(size_t)t * 5 + s];
    }
    stats[8] = 0;
    for (int t = 0; t < threads; t++) stats[8] += busy[(size_t)t * 5 + 4];
    double pj = 0;
    for (int f = 0; f < S; f++) pj += nProj[(size_t)f];
    stats[9] = S ? pj / S : 0;
    stats[10] = stats[11] = 0;
    double kp = 0, st = 0, bw = 0;
    for (int f = 0; f < S; f++) {
      kp += frames[(size_t)f]->N;
      for (int i = 0; i < frames[(size_t)f]->N; i++) st += frames[(size_t)f]->mvuRight[(size_t)i] >= 0;
      bw += nBow[(size_t)f];
    }
    stats[4] = S ? kp / S : 0;
    stats[5] = S ? st / S : 0;
    stats[6] = S ? bw / S : 0;
    stats[7] = wall;
  }
  return wall;
}
}

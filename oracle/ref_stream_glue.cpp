// TEST / BENCH INFRASTRUCTURE — the reference arm of bench.py (`--impl reference`, `cpu_baseline`): one step of the bench
// workload on the host cores with the REFERENCE'S OWN code wherever it can be built here — src/ORBextractor.cc for the
// extraction and src/ORBmatcher.cc (SearchByBoW(KeyFrame*, Frame&, ...)) for the temporal matching, both compiled in
// place, unmodified, against oracle/refshim — and the oracle port of Optimizer::LocalBundleAdjustment (src/Optimizer.cc
// needs g2o + Eigen, which are not in this image).  Same task layout as orc_stream_step (oracle/orb_misc.cpp): LocalBA
// windows and images are dealt to `threads` host threads (one image per task like the two extractor threads of
// src/Frame.cc:159-167, LocalBA single-threaded per window like g2o without OpenMP), then the S frame pairs are matched.
// Built into oracle/_ref/libref_stream.so (git-ignored).  Never part of the product.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "orb_oracle.h"

using namespace ORB_SLAM2;

float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

namespace {
struct Extracted {
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
};

void fill(FeatureHolder& h, const Extracted& e) {
  h.N = (int)e.kps.size();
  h.mvKeys = e.kps;
  h.mvKeysUn = e.kps;
  h.mvuRight.assign(h.N, -1.0f);
  h.mDescriptors = e.desc;
  h.mvpMapPoints.assign(h.N, (MapPoint*)NULL);
  std::vector<unsigned int>& all = h.mFeatVec[0];  // one vocabulary node: the 2000 x 2000 brute-force case of the workload
  all.resize(h.N);
  for (int i = 0; i < h.N; i++) all[i] = (unsigned)i;
}
}  // namespace

extern "C" double ref_stream_step(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs, int S,
                                  int w, int h, const orc_ba_problem* ba, int ba_every, int threads, int* out_counts) {
  if (threads < 1) threads = 1;
  const int nImg = 2 * S;
  const int nBa = (ba && ba_every > 0) ? (S + ba_every - 1) / ba_every : 0;
  std::vector<Extracted> ex(nImg);
  const auto t0 = std::chrono::steady_clock::now();
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&]() {
        std::unique_ptr<ORBextractor> e;
        for (;;) {
          const int task = next.fetch_add(1);
          if (task >= nBa + nImg) break;
          if (task < nBa) {  // longest tasks first
            std::vector<float> T((size_t)ba->n_local * 16), P((size_t)ba->n_mp * 3);
            std::vector<uint8_t> o(ba->n_edges);
            orc_ba_result r;
            std::memset(&r, 0, sizeof(r));
            r.Tcw_out = T.data();
            r.points_out = P.data();
            r.edge_outlier = o.data();
            orc_local_ba(ba, nullptr, &r);
          } else {
            const int i = task - nBa;
            if (!e) e.reset(new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh));
            cv::Mat image(h, w, CV_8UC1, (void*)(imgs + (size_t)i * w * h), (size_t)w);
            cv::Mat mask;
            (*e)(image, mask, ex[i].kps, ex[i].desc);
          }
        }
      });
    for (auto& t : th) t.join();
  }
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&]() {
        for (;;) {
          const int f = next.fetch_add(1);
          if (f >= S) break;
          const int a = (f + S - 1) % S, b = f;  // left images are 0..S-1: frame f against its predecessor
          std::unique_ptr<KeyFrame> K(new KeyFrame());
          std::unique_ptr<Frame> F(new Frame());
          fill(*K, ex[a]);
          fill(*F, ex[b]);
          std::vector<MapPoint> pts(K->N);  // every keyframe feature holds a valid map point
          for (int i = 0; i < K->N; i++) K->mvpMapPoints[i] = &pts[i];
          ORBmatcher m(0.7f, true);
          std::vector<MapPoint*> matches;
          const int n = m.SearchByBoW(K.get(), *F, matches);
          if (out_counts) out_counts[f] = n;
        }
      });
    for (auto& t : th) t.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

/*
 * oracle/orb_misc.cpp — helpers around the oracle (TEST INFRASTRUCTURE, see orb_oracle.h):
 *  - glibc sinf/cosf over a batch (the values the reference gets at src/ORBextractor.cc:181), used to check the
 *    device restatement of glibc's sincosf exhaustively;
 *  - a multi-threaded driver that runs the extractor oracle over many images (CPU baseline: one thread per image,
 *    frames spread over the host cores, like the two extraction threads of src/Frame.cc:159-167).
 */
#include <cmath>
#include <thread>
#include <vector>

#include "orb_oracle.h"

extern "C" void orc_sincosf_batch(const float* in, long n, float* s, float* c, int threads) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([=]() {
      const long beg = n * t / threads, end = n * (t + 1) / threads;
      for (long i = beg; i < end; i++) {
        s[i] = sinf(in[i]);
        c[i] = cosf(in[i]);
      }
    });
  for (auto& t : th) t.join();
}

/* Extract `count` images (each w*h, tightly packed, consecutive) with `threads` worker threads; every worker owns
 * its own extractor.  Outputs per image: kps[i*cap..], desc[i*cap*32..], n_out[i]. */
extern "C" int orc_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs,
                                int count, int w, int h, orc_keypoint* kps, uint8_t* desc, int cap, int* n_out,
                                int threads) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  std::vector<int> rc(threads, 0);
  for (int t = 0; t < threads; t++)
    th.emplace_back([=, &rc]() {
      void* e = orc_extractor_create(nfeatures, scaleFactor, nlevels, iniTh, minTh);
      for (int i = t; i < count; i += threads) {
        int n = orc_extract(e, imgs + (size_t)i * w * h, w, h, w, kps + (size_t)i * cap, desc + (size_t)i * cap * 32, cap);
        if (n < 0) rc[t] = -1;
        n_out[i] = n;
      }
      orc_extractor_destroy(e);
    });
  for (auto& t : th) t.join();
  for (int t = 0; t < threads; t++)
    if (rc[t]) return -1;
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CPU reference arm of bench.py: one "step" of the hot path over S stereo frames on `threads` host threads.
 *   extract  : ORBextractor::operator() on the 2S images (one image per task, like the L/R threads of src/Frame.cc:159)
 *   match    : SearchByBoW(left(t-1), left(t)) with every feature in one vocabulary node (2000x2000 brute force,
 *              BASELINE.json configs[2]); frame 0 is matched against the last frame of the batch (ring)
 *   LocalBA  : one window per `ba_every` frames (single-threaded each, like g2o without OpenMP)
 * Returns wall seconds; out_counts (optional, S ints) receives the per-frame match counts for cross-checking.
 * ------------------------------------------------------------------------------------------------ */
#include <atomic>
#include <chrono>
#include <cstring>

extern "C" double orc_stream_step(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs,
                                  int S, int w, int h, const orc_ba_problem* ba, int ba_every, int threads,
                                  int* out_counts) {
  if (threads < 1) threads = 1;
  const int cap = nfeatures + 4 * nlevels + 16;
  const int nImg = 2 * S;
  const int nBa = (ba && ba_every > 0) ? (S + ba_every - 1) / ba_every : 0;
  std::vector<orc_keypoint> kps((size_t)nImg * cap);
  std::vector<uint8_t> desc((size_t)nImg * cap * 32);
  std::vector<int> cnt(nImg, 0);
  const auto t0 = std::chrono::steady_clock::now();
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&]() {
        void* e = nullptr;
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
            if (!e) e = orc_extractor_create(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            cnt[i] = orc_extract(e, imgs + (size_t)i * w * h, w, h, w, &kps[(size_t)i * cap], &desc[(size_t)i * cap * 32], cap);
          }
        }
        if (e) orc_extractor_destroy(e);
      });
    for (auto& t : th) t.join();
  }
  {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&]() {
        std::vector<int32_t> node, match;
        std::vector<float> angA, angB;
        std::vector<uint8_t> valid;
        for (;;) {
          const int f = next.fetch_add(1);
          if (f >= S) break;
          const int a = (f + S - 1) % S, b = f;  // left images are 0..S-1
          const int nA = cnt[a], nB = cnt[b];
          node.assign(std::max(nA, nB), 0);
          valid.assign(nA, 1);
          match.assign(nB, -1);
          angA.resize(nA);
          angB.resize(nB);
          for (int i = 0; i < nA; i++) angA[i] = kps[(size_t)a * cap + i].angle;
          for (int i = 0; i < nB; i++) angB[i] = kps[(size_t)b * cap + i].angle;
          const int n = orc_search_by_bow(&desc[(size_t)a * cap * 32], node.data(), valid.data(), angA.data(), nA,
                                          &desc[(size_t)b * cap * 32], node.data(), nullptr, angB.data(), nB, 50, 0.7f, 0, 1,
                                          match.data());
          if (out_counts) out_counts[f] = n;
        }
      });
    for (auto& t : th) t.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

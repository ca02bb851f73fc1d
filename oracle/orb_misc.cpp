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

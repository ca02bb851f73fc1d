// oracle/ref_glue.cpp — TEST INFRASTRUCTURE ONLY: C entry point around the reference's own ORB_SLAM2::ORBextractor
// (compiled in place from /root/reference/src/ORBextractor.cc against oracle/refshim, see Makefile target _ref).
#include <cstdlib>
#include <new>
#include <sys/mman.h>
#include <vector>

#include "ORBextractor.h"

// DistributeOctTree sorts (size, ExtractorNode*) pairs (src/ORBextractor.cc:948): ties between equally large nodes are
// broken by the HEAP ADDRESS of the list nodes, i.e. by the allocator.  With g_monotonic set, every allocation made by
// this library comes from a bump arena (addresses strictly increase with creation order, nothing is reused), which is
// the rule the oracle restates ("later-created node first"); without it the reference runs on the system malloc.
static int g_monotonic = 0;
static char* g_arena = nullptr;
static size_t g_off = 0;
static const size_t kArena = (size_t)3 << 30;
static void* arena_alloc(size_t n) {
  if (!g_arena) g_arena = (char*)mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  const size_t a = (g_off + 15) & ~(size_t)15;
  if (a + n > kArena) std::abort();
  g_off = a + n;
  return g_arena + a;
}
void* operator new(size_t n) {
  void* p = g_monotonic ? arena_alloc(n) : std::malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
static inline bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + kArena; }
void operator delete(void* p) noexcept {
  if (p && !in_arena(p)) std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

extern "C" void ref_set_monotonic_allocator(int on) {
  g_monotonic = on;
  g_off = 0;  // (callers switch between extractions only: no live arena objects)
}

extern "C" int ref_extract(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, const uint8_t* img, int w,
                           int h, float* kps /* n x 7: x y size angle response octave class_id */, uint8_t* desc, int cap) {
  if (g_monotonic) g_off = 0;
  ORB_SLAM2::ORBextractor ex(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)w);
  cv::Mat mask, descriptors;
  std::vector<cv::KeyPoint> keypoints;
  ex(image, mask, keypoints, descriptors);
  const int n = (int)keypoints.size();
  if (n > cap) return -1;
  for (int i = 0; i < n; i++) {
    const cv::KeyPoint& k = keypoints[i];
    float* o = kps + (size_t)i * 7;
    o[0] = k.pt.x; o[1] = k.pt.y; o[2] = k.size; o[3] = k.angle; o[4] = k.response; o[5] = (float)k.octave; o[6] = (float)k.class_id;
    for (int b = 0; b < 32; b++) desc[(size_t)i * 32 + b] = descriptors.ptr(i)[b];
  }
  return n;
}

/*
 * oracle/orb_extract.cpp — CPU restatement of ORBextractor (TEST INFRASTRUCTURE, see orb_oracle.h).
 *
 * Follows /root/reference/src/ORBextractor.cc function by function (file:line cited per function)
 * and restates the OpenCV primitives it calls (resize / FAST / GaussianBlur / fastAtan2 / cvRound)
 * with the integer / float recipes of SURVEY.md §8c.  Build with -ffp-contract=off: the
 * reference arithmetic must not be FMA-contracted.
 */
#include "orb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

const int PATCH_SIZE = 31;       // src/ORBextractor.cc:82
const int HALF_PATCH_SIZE = 15;  // :84
const int EDGE_THRESHOLD = 19;   // :86

static const int kPattern[256 * 4] = {
#include "../data/orb_pattern_31.inc"
};

inline int cv_round(double v) { return (int)lrint(v); }  // cvRound: round-half-to-even
inline int cv_roundf(float v) { return (int)lrintf(v); }
inline int cv_floor(double v) {
  int i = (int)v;
  return i - (i > v);
}
inline int cv_ceil(double v) {
  int i = (int)v;
  return i + (i < v);
}

struct Image {
  int w = 0, h = 0;
  std::vector<uint8_t> px;
  void alloc(int w_, int h_) {
    w = w_;
    h = h_;
    px.assign((size_t)w * h, 0);
  }
  const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
  uint8_t* row(int y) { return px.data() + (size_t)y * w; }
};

}  // namespace

/* ------------------------------------------------------------------------------------------
 * cv::resize(src, dst, sz, 0, 0, INTER_LINEAR) for CV_8UC1 — called at src/ORBextractor.cc:1696.
 * OpenCV generic fixed-point path (INTER_RESIZE_COEF_BITS = 11): per-axis offsets/weights, horizontal
 * pass into int32, vertical pass with the (>>4, >>16, +2, >>2) rounding chain.
 * ------------------------------------------------------------------------------------------ */
extern "C" void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                                     int dstride) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ialpha(dw * 2), ibeta(dh * 2);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor(fx);
    fx -= sx;
    if (sx < 0) {
      fx = 0;
      sx = 0;
    }
    if (sx >= sw - 1) {
      fx = 0;
      sx = sw - 1;
    }
    xofs[dx] = sx;
    ialpha[dx * 2] = (short)cv_roundf((1.f - fx) * 2048.f);
    ialpha[dx * 2 + 1] = (short)cv_roundf(fx * 2048.f);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    ibeta[dy * 2] = (short)cv_roundf((1.f - fy) * 2048.f);
    ibeta[dy * 2 + 1] = (short)cv_roundf(fy * 2048.f);
  }
  std::vector<int> r0(dw), r1(dw);
  auto hrow = [&](int sy, std::vector<int>& out) {
    sy = std::min(std::max(sy, 0), sh - 1);  // clip(sy, 0, ssize.height)
    const uint8_t* S = src + (size_t)sy * sstride;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx];
      int s1 = (sx + 1 < sw) ? S[sx + 1] : S[sx];
      out[dx] = S[sx] * ialpha[dx * 2] + s1 * ialpha[dx * 2 + 1];
    }
  };
  for (int dy = 0; dy < dh; dy++) {
    hrow(yofs[dy], r0);
    hrow(yofs[dy] + 1, r1);
    const int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
    uint8_t* D = dst + (size_t)dy * dstride;
    for (int x = 0; x < dw; x++) D[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

/* ------------------------------------------------------------------------------------------
 * cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) on a contiguous CV_8UC1 clone —
 * src/ORBextractor.cc:1626-1634.  OpenCV-4 bit-exact fixed-point kernel: Q8 taps
 * [18,34,48,56,48,34,18], horizontal then vertical, single final rounding (V + 2^15) >> 16.
 * ------------------------------------------------------------------------------------------ */
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

extern "C" void orc_gaussian_blur7_s2_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint16_t> H((size_t)w * h);
  std::vector<uint8_t> pad(w + 6);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src + (size_t)y * sstride;
    for (int i = 0; i < 3; i++) {
      pad[i] = S[reflect101(i - 3, w)];
      pad[w + 3 + i] = S[reflect101(w + i, w)];
    }
    std::memcpy(pad.data() + 3, S, w);
    uint16_t* Hr = &H[(size_t)y * w];
    const uint8_t* p = pad.data();
    for (int x = 0; x < w; x++)
      Hr[x] = (uint16_t)(k[0] * p[x] + k[1] * p[x + 1] + k[2] * p[x + 2] + k[3] * p[x + 3] + k[4] * p[x + 4] +
                         k[5] * p[x + 5] + k[6] * p[x + 6]);  // <= 255*256
  }
  for (int y = 0; y < h; y++) {
    uint8_t* D = dst + (size_t)y * dstride;
    const uint16_t* r[7];
    for (int t = 0; t < 7; t++) r[t] = &H[(size_t)reflect101(y + t - 3, h) * w];
    for (int x = 0; x < w; x++) {
      uint32_t acc = (uint32_t)k[0] * r[0][x] + (uint32_t)k[1] * r[1][x] + (uint32_t)k[2] * r[2][x] +
                     (uint32_t)k[3] * r[3][x] + (uint32_t)k[4] * r[4][x] + (uint32_t)k[5] * r[5][x] +
                     (uint32_t)k[6] * r[6][x];
      D[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * cv::FAST(roi, keypoints, threshold, nonmaxSuppression=true) — called per cell at
 * src/ORBextractor.cc:1126,1135.  FAST-9/16: corner <=> 9 contiguous circle pixels all brighter than
 * v+th or all darker than v-th; score = max threshold for which it stays a corner
 * = max over the 16 arcs of min|v-p| (same sign) - 1; NMS strict '>' over the 8 neighbours'
 * scores (non-corners 0); pixels closer than 3 to the ROI edge are never tested; row-major output.
 * ------------------------------------------------------------------------------------------ */
static const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static inline int fast_arc_strength(const uint8_t* p, int stride) {
  // M = max_s max( min_{j<9} d_{s+j}, min_{j<9} -d_{s+j} ), d_k = v - p_k ; clamped at 0
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = v - p[kCircle[k][1] * stride + kCircle[k][0]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int M = 0;
  for (int s = 0; s < 16; s++) {
    int mn = d[s], mx = d[s];
    for (int j = 1; j < 9; j++) {
      mn = std::min(mn, d[s + j]);
      mx = std::max(mx, d[s + j]);
    }
    M = std::max(M, std::max(mn, -mx));
  }
  return M;
}

// Quick reject (exact): every 9-arc contains one pixel of each opposite pair (k, k+8), so a dark corner at
// threshold t needs max(d_k, d_{k+8}) > t for all 8 pairs, a bright one max(-d_k, -d_{k+8}) > t.
static inline bool fast_can_be_corner(const uint8_t* p, int stride, int t) {
  const int v = p[0];
  bool dark = true, bright = true;
  for (int k = 0; k < 8 && (dark || bright); k++) {
    const int kk = (k * 4) % 8 + (k / 2);  // visit 0,4,1,5,2,6,3,7: far-apart pairs first
    const int d0 = v - p[kCircle[kk][1] * stride + kCircle[kk][0]];
    const int d1 = v - p[kCircle[kk + 8][1] * stride + kCircle[kk + 8][0]];
    dark = dark && (std::max(d0, d1) > t);
    bright = bright && (std::max(-d0, -d1) > t);
  }
  return dark || bright;
}

// prune_th >= 0: pixels that cannot be a corner at prune_th get 0 (they would never pass `M > th` for
// th >= prune_th, and count as 0 in the NMS either way) — a speed-up that leaves results unchanged.
static void fast_score_map(const uint8_t* roi, int w, int h, int stride, uint8_t* score, int score_stride, int prune_th) {
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      int M = 0;
      if (y >= 3 && y < h - 3 && x >= 3 && x < w - 3) {
        const uint8_t* p = roi + (size_t)y * stride + x;
        if (prune_th < 0 || fast_can_be_corner(p, stride, prune_th)) M = fast_arc_strength(p, stride);
      }
      score[(size_t)y * score_stride + x] = (uint8_t)std::min(M, 255);
    }
  }
}

extern "C" void orc_fast_score_map(const uint8_t* roi, int w, int h, int stride, uint8_t* score, int score_stride) {
  fast_score_map(roi, w, h, stride, score, score_stride, -1);
}

extern "C" int orc_fast9_16_nms(const uint8_t* roi, int w, int h, int stride, int th, int* xy, int* resp, int cap) {
  if (w < 7 || h < 7) return 0;
  std::vector<uint8_t> M((size_t)w * h);
  fast_score_map(roi, w, h, stride, M.data(), w, th);
  int n = 0;
  for (int y = 3; y < h - 3; y++) {
    for (int x = 3; x < w - 3; x++) {
      const int m = M[(size_t)y * w + x];
      if (m <= th) continue;  // corner <=> M > th  (score = M-1 >= th)
      const int s = m - 1;
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          const int mm = M[(size_t)(y + dy) * w + (x + dx)];  // 0 on the rim
          const int ns = (mm > th) ? mm - 1 : 0;              // neighbour's score in OpenCV's buffer
          if (!(s > ns)) {
            keep = false;
            break;
          }
        }
      if (!keep) continue;
      if (n < cap) {
        xy[2 * n] = x;
        xy[2 * n + 1] = y;
        resp[n] = s;
      }
      n++;
    }
  }
  return n;
}

/* cv::fastAtan2(y, x) in degrees — called at src/ORBextractor.cc:160. OpenCV scalar polynomial. */
extern "C" float orc_fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
              p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

/* ------------------------------------------------------------------------------------------
 * DistributeOctTree — src/ORBextractor.cc:706-1049, DivideNode :635-703.
 * Deviation: the reference sorts pair<int,ExtractorNode*> so ties on the keypoint count
 * are broken by heap address, which is not reproducible.  The oracle breaks ties by node creation
 * sequence number (later-created node = "larger pointer" = expanded first).
 * ------------------------------------------------------------------------------------------ */
namespace {
struct Node {
  std::vector<orc_keypoint> keys;
  int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
  std::list<Node>::iterator lit;
  bool noMore = false;
  long seq = 0;
};

void divide_node(const Node& p, Node& n1, Node& n2, Node& n3, Node& n4) {
  const int halfX = (int)std::ceil((float)(p.URx - p.ULx) / 2);
  const int halfY = (int)std::ceil((float)(p.BRy - p.ULy) / 2);
  n1.ULx = p.ULx; n1.ULy = p.ULy;
  n1.URx = p.ULx + halfX; n1.URy = p.ULy;
  n1.BLx = p.ULx; n1.BLy = p.ULy + halfY;
  n1.BRx = p.ULx + halfX; n1.BRy = p.ULy + halfY;
  n2.ULx = n1.URx; n2.ULy = n1.URy;
  n2.URx = p.URx; n2.URy = p.URy;
  n2.BLx = n1.BRx; n2.BLy = n1.BRy;
  n2.BRx = p.URx; n2.BRy = p.ULy + halfY;
  n3.ULx = n1.BLx; n3.ULy = n1.BLy;
  n3.URx = n1.BRx; n3.URy = n1.BRy;
  n3.BLx = p.BLx; n3.BLy = p.BLy;
  n3.BRx = n1.BRx; n3.BRy = p.BLy;
  n4.ULx = n3.URx; n4.ULy = n3.URy;
  n4.URx = n2.BRx; n4.URy = n2.BRy;
  n4.BLx = n3.BRx; n4.BLy = n3.BRy;
  n4.BRx = p.BRx; n4.BRy = p.BRy;
  for (const orc_keypoint& kp : p.keys) {
    if (kp.x < n1.URx) {
      if (kp.y < n1.BRy) n1.keys.push_back(kp);
      else n3.keys.push_back(kp);
    } else if (kp.y < n1.BRy)
      n2.keys.push_back(kp);
    else
      n4.keys.push_back(kp);
  }
  if (n1.keys.size() == 1) n1.noMore = true;
  if (n2.keys.size() == 1) n2.noMore = true;
  if (n3.keys.size() == 1) n3.noMore = true;
  if (n4.keys.size() == 1) n4.noMore = true;
}

std::vector<orc_keypoint> distribute_octtree(const std::vector<orc_keypoint>& in, int minX, int maxX, int minY, int maxY,
                                             int N) {
  std::vector<orc_keypoint> result;
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));  // :719 (C round)
  if (nIni <= 0 || in.empty()) {
    // nIni==0 would divide by zero in the reference (never happens for landscape images); be safe.
    if (in.empty()) return result;
  }
  const float hX = (float)(maxX - minX) / nIni;
  std::list<Node> nodes;
  std::vector<Node*> ini(nIni);
  long seq = 0;
  for (int i = 0; i < nIni; i++) {
    Node ni;
    ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
    ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
    ni.BLx = ni.ULx; ni.BLy = maxY - minY;
    ni.BRx = ni.URx; ni.BRy = maxY - minY;
    ni.seq = seq++;
    nodes.push_back(ni);
    ini[i] = &nodes.back();
  }
  for (const orc_keypoint& kp : in) ini[(int)(kp.x / hX)]->keys.push_back(kp);  // :766
  for (auto lit = nodes.begin(); lit != nodes.end();) {
    if (lit->keys.size() == 1) {
      lit->noMore = true;
      ++lit;
    } else if (lit->keys.empty())
      lit = nodes.erase(lit);
    else
      ++lit;
  }
  bool finish = false;
  typedef std::pair<int, Node*> SizeNode;
  auto cmp = [](const SizeNode& a, const SizeNode& b) {
    if (a.first != b.first) return a.first < b.first;
    return a.second->seq < b.second->seq;  // deterministic stand-in for the pointer compare (:948)
  };
  std::vector<SizeNode> sizeAndNode;
  auto push_child = [&](Node& c, bool count, int& nToExpand) {
    if (c.keys.empty()) return;
    c.seq = seq++;
    nodes.push_front(c);
    if (c.keys.size() > 1) {
      if (count) nToExpand++;
      sizeAndNode.push_back(std::make_pair((int)c.keys.size(), &nodes.front()));
      nodes.front().lit = nodes.begin();
    }
  };
  while (!finish) {
    int prevSize = (int)nodes.size();
    auto lit = nodes.begin();
    int nToExpand = 0;
    sizeAndNode.clear();
    while (lit != nodes.end()) {
      if (lit->noMore) {
        ++lit;
        continue;
      }
      Node n1, n2, n3, n4;
      divide_node(*lit, n1, n2, n3, n4);
      push_child(n1, true, nToExpand);
      push_child(n2, true, nToExpand);
      push_child(n3, true, nToExpand);
      push_child(n4, true, nToExpand);
      lit = nodes.erase(lit);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
      finish = true;
    } else if (((int)nodes.size() + nToExpand * 3) > N) {
      while (!finish) {
        prevSize = (int)nodes.size();
        std::vector<SizeNode> prev = sizeAndNode;
        sizeAndNode.clear();
        std::sort(prev.begin(), prev.end(), cmp);
        for (int j = (int)prev.size() - 1; j >= 0; j--) {
          Node n1, n2, n3, n4;
          divide_node(*prev[j].second, n1, n2, n3, n4);
          int dummy = 0;
          push_child(n1, false, dummy);
          push_child(n2, false, dummy);
          push_child(n3, false, dummy);
          push_child(n4, false, dummy);
          nodes.erase(prev[j].second->lit);
          if ((int)nodes.size() >= N) break;
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
      }
    }
  }
  result.reserve(nodes.size());
  for (auto& nd : nodes) {
    const orc_keypoint* best = &nd.keys[0];
    float maxResponse = best->response;
    for (size_t k = 1; k < nd.keys.size(); k++)
      if (nd.keys[k].response > maxResponse) {
        best = &nd.keys[k];
        maxResponse = nd.keys[k].response;
      }
    result.push_back(*best);
  }
  return result;
}
}  // namespace

extern "C" int orc_distribute_octtree(const orc_keypoint* in, int n, int minX, int maxX, int minY, int maxY, int N,
                                      orc_keypoint* out, int cap) {
  std::vector<orc_keypoint> v(in, in + n);
  std::vector<orc_keypoint> r = distribute_octtree(v, minX, maxX, minY, maxY, N);
  if ((int)r.size() > cap) return -1;
  std::copy(r.begin(), r.end(), out);
  return (int)r.size();
}

/* ------------------------------------------------------------------------------------------ */
namespace {
struct Extractor {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // include/ORBextractor.h: `double scaleFactor` initialised from the float ctor argument
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nFeat;
  int umax[HALF_PATCH_SIZE + 1];
  std::vector<Image> pyr, blurred;
  std::vector<std::vector<orc_keypoint>> cand, kps;

  // ORBextractor::ORBextractor — src/ORBextractor.cc:492-609
  Extractor(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
    scale.resize(nl);
    sigma2.resize(nl);
    scale[0] = 1.0f;
    sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
      scale[i] = (float)(scale[i - 1] * scaleFactor);
      sigma2[i] = scale[i] * scale[i];
    }
    invScale.resize(nl);
    invSigma2.resize(nl);
    for (int i = 0; i < nl; i++) {
      invScale[i] = 1.0f / scale[i];
      invSigma2[i] = 1.0f / sigma2[i];
    }
    nFeat.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      nFeat[l] = cv_roundf(nDesired);
      sum += nFeat[l];
      nDesired *= factor;
    }
    nFeat[nl - 1] = std::max(nfeatures - sum, 0);
    // umax :579-608
    int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= HALF_PATCH_SIZE; v++) umax[v] = 0;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    pyr.resize(nl);
    blurred.resize(nl);
    cand.resize(nl);
    kps.resize(nl);
  }

  // ComputePyramid — :1674-1734 (the 19-px border is never read downstream; not materialised)
  void compute_pyramid(const uint8_t* img, int w, int h, int stride) {
    for (int l = 0; l < nlevels; l++) {
      float s = invScale[l];
      int lw = cv_roundf((float)w * s), lh = cv_roundf((float)h * s);
      pyr[l].alloc(lw, lh);
      if (l == 0) {
        for (int y = 0; y < h; y++) std::memcpy(pyr[0].row(y), img + (size_t)y * stride, w);
      } else {
        orc_resize_linear_u8(pyr[l - 1].px.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w, pyr[l].px.data(), lw, lh, lw);
      }
    }
  }

  // IC_Angle — :108-161
  float ic_angle(const Image& im, float ptx, float pty) const {
    int m_01 = 0, m_10 = 0;
    const int step = im.w;
    const uint8_t* center = im.row(cv_roundf(pty)) + cv_roundf(ptx);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      int v_sum = 0;
      int d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int val_plus = center[u + v * step], val_minus = center[u - v * step];
        v_sum += (val_plus - val_minus);
        m_10 += u * (val_plus + val_minus);
      }
      m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
  }

  // ComputeKeyPointsOctTree — :1052-1199
  void compute_keypoints() {
    const float W = 30;
    for (int level = 0; level < nlevels; ++level) {
      const Image& im = pyr[level];
      const int minBorderX = EDGE_THRESHOLD - 3;
      const int minBorderY = minBorderX;
      const int maxBorderX = im.w - EDGE_THRESHOLD + 3;
      const int maxBorderY = im.h - EDGE_THRESHOLD + 3;
      std::vector<orc_keypoint>& toDist = cand[level];
      toDist.clear();
      kps[level].clear();
      const float width = (float)(maxBorderX - minBorderX);
      const float height = (float)(maxBorderY - minBorderY);
      const int nCols = (int)(width / W);
      const int nRows = (int)(height / W);
      if (nCols <= 0 || nRows <= 0) continue;  // degenerate tiny level (reference would divide by zero)
      const int wCell = (int)std::ceil(width / nCols);
      const int hCell = (int)std::ceil(height / nRows);
      std::vector<int> xy, resp;
      for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
          const float iniX = (float)(minBorderX + j * wCell);
          float maxX = iniX + wCell + 6;
          if (iniX >= maxBorderX - 6) continue;
          if (maxX > maxBorderX) maxX = (float)maxBorderX;
          const int rw = (int)maxX - (int)iniX, rh = (int)maxY - (int)iniY;
          const uint8_t* roi = im.row((int)iniY) + (int)iniX;
          const int cap = rw * rh;
          xy.resize(2 * cap + 2);
          resp.resize(cap + 1);
          int n = orc_fast9_16_nms(roi, rw, rh, im.w, iniTh, xy.data(), resp.data(), cap);
          if (n == 0) n = orc_fast9_16_nms(roi, rw, rh, im.w, minTh, xy.data(), resp.data(), cap);
          for (int k = 0; k < n; k++) {
            orc_keypoint kp;
            kp.x = (float)xy[2 * k] + j * wCell;
            kp.y = (float)xy[2 * k + 1] + i * hCell;
            kp.size = 7.f;
            kp.angle = -1.f;
            kp.response = (float)resp[k];
            kp.octave = 0;
            kp.class_id = -1;
            toDist.push_back(kp);
          }
        }
      }
      std::vector<orc_keypoint> sel = distribute_octtree(toDist, minBorderX, maxBorderX, minBorderY, maxBorderY, nFeat[level]);
      const int scaledPatchSize = (int)(PATCH_SIZE * scale[level]);
      for (orc_keypoint& kp : sel) {
        kp.x += minBorderX;
        kp.y += minBorderY;
        kp.octave = level;
        kp.size = (float)scaledPatchSize;
      }
      kps[level] = sel;
    }
    for (int level = 0; level < nlevels; ++level)
      for (orc_keypoint& kp : kps[level]) kp.angle = ic_angle(pyr[level], kp.x, kp.y);
  }

  // computeOrbDescriptor — :173-227
  void descriptor(const orc_keypoint& kpt, const Image& img, uint8_t* desc) const {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = (float)kpt.angle * factorPI;
    float a = (float)std::cos(angle), b = (float)std::sin(angle);  // cos/sin(float) -> cosf/sinf
    const uint8_t* center = img.row(cv_roundf(kpt.y)) + cv_roundf(kpt.x);
    const int step = img.w;
    const int* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
      int val = 0;
      for (int k = 0; k < 8; k++) {
        const int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
        int t0 = center[cv_roundf(x0 * b + y0 * a) * step + cv_roundf(x0 * a - y0 * b)];
        int t1 = center[cv_roundf(x1 * b + y1 * a) * step + cv_roundf(x1 * a - y1 * b)];
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // operator() — :1544-1668
  int extract(const uint8_t* img, int w, int h, int stride, orc_keypoint* out, uint8_t* desc, int cap) {
    if (!img || w <= 0 || h <= 0) return 0;
    compute_pyramid(img, w, h, stride);
    compute_keypoints();
    int total = 0;
    for (int l = 0; l < nlevels; l++) total += (int)kps[l].size();
    if (total > cap) return -1;
    int offset = 0;
    for (int l = 0; l < nlevels; l++) {
      blurred[l].w = blurred[l].h = 0;
      blurred[l].px.clear();
      std::vector<orc_keypoint>& k = kps[l];
      if (k.empty()) continue;
      blurred[l].alloc(pyr[l].w, pyr[l].h);
      orc_gaussian_blur7_s2_u8(pyr[l].px.data(), pyr[l].w, pyr[l].h, pyr[l].w, blurred[l].px.data(), pyr[l].w);
      for (size_t i = 0; i < k.size(); i++) descriptor(k[i], blurred[l], desc + (size_t)(offset + i) * 32);
      for (size_t i = 0; i < k.size(); i++) {
        orc_keypoint kp = k[i];
        if (l != 0) {
          kp.x *= scale[l];
          kp.y *= scale[l];
        }
        out[offset + i] = kp;
      }
      offset += (int)k.size();
    }
    return total;
  }
};
}  // namespace

extern "C" void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
  return new Extractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
extern "C" void orc_extractor_destroy(void* h) { delete (Extractor*)h; }
extern "C" int orc_extract(void* h, const uint8_t* img, int w, int hgt, int stride, orc_keypoint* kps, uint8_t* desc,
                           int cap) {
  return ((Extractor*)h)->extract(img, w, hgt, stride, kps, desc, cap);
}
extern "C" void orc_extractor_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                                     int* nfeat_per_level, int* umax16) {
  Extractor* e = (Extractor*)h;
  for (int i = 0; i < e->nlevels; i++) {
    if (scale) scale[i] = e->scale[i];
    if (inv_scale) inv_scale[i] = e->invScale[i];
    if (sigma2) sigma2[i] = e->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = e->invSigma2[i];
    if (nfeat_per_level) nfeat_per_level[i] = e->nFeat[i];
  }
  if (umax16)
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}
extern "C" int orc_level_dims(void* h, int level, int* w, int* hgt) {
  Extractor* e = (Extractor*)h;
  if (level < 0 || level >= e->nlevels) return -1;
  *w = e->pyr[level].w;
  *hgt = e->pyr[level].h;
  return 0;
}
extern "C" const uint8_t* orc_level_image(void* h, int level) { return ((Extractor*)h)->pyr[level].px.data(); }
extern "C" const uint8_t* orc_level_blurred(void* h, int level) {
  Extractor* e = (Extractor*)h;
  return e->blurred[level].px.empty() ? nullptr : e->blurred[level].px.data();
}
extern "C" int orc_level_candidates(void* h, int level, orc_keypoint* out, int cap) {
  Extractor* e = (Extractor*)h;
  int n = (int)e->cand[level].size();
  if (out) std::copy(e->cand[level].begin(), e->cand[level].begin() + std::min(n, cap), out);
  return n;
}
extern "C" int orc_level_keypoints(void* h, int level, orc_keypoint* out, int cap) {
  Extractor* e = (Extractor*)h;
  int n = (int)e->kps[level].size();
  if (out) std::copy(e->kps[level].begin(), e->kps[level].begin() + std::min(n, cap), out);
  return n;
}

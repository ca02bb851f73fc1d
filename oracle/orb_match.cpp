/*
 * oracle/orb_match.cpp — CPU restatement of ORBmatcher (TEST INFRASTRUCTURE, see orb_oracle.h).
 * Follows /root/reference/src/ORBmatcher.cc on flattened arrays (the pointer graph of
 * Frame/KeyFrame/MapPoint is gathered by the host shim).
 */
#include "orb_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {
const int HISTO_LENGTH = 30;  // src/ORBmatcher.cc:51

// ComputeThreeMaxima — src/ORBmatcher.cc:1866-1908
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) {
      max3 = max2; max2 = max1; max1 = s;
      ind3 = ind2; ind2 = ind1; ind1 = i;
    } else if (s > max2) {
      max3 = max2; max2 = s;
      ind3 = ind2; ind2 = i;
    } else if (s > max3) {
      max3 = s;
      ind3 = i;
    }
  }
  if (max2 < 0.1f * (float)max1) {
    ind2 = -1;
    ind3 = -1;
  } else if (max3 < 0.1f * (float)max1) {
    ind3 = -1;
  }
}

inline int rot_bin(float angA, float angB) {
  const float factor = HISTO_LENGTH / 360.0f;  // :246 (this fork; == 1.0f/HISTO_LENGTH bug of upstream is fixed here)
  float rot = angA - angB;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);  // C round(): half away from zero
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}
}  // namespace

// DescriptorDistance — src/ORBmatcher.cc:1913-1933 (bit-twiddling popcount over 8 int32 words)
extern "C" int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t wa, wb;
    std::memcpy(&wa, a + 4 * i, 4);
    std::memcpy(&wb, b + 4 * i, 4);
    unsigned int v = wa ^ wb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// SearchByBoW — :230-382 (strict_lt=0) and :656-799 (strict_lt=1, validB used as "F side needs a MapPoint")
extern "C" int orc_search_by_bow(const uint8_t* descA, const int32_t* nodeA, const uint8_t* validA, const float* angA,
                                 int nA, const uint8_t* descB, const int32_t* nodeB, const uint8_t* validB,
                                 const float* angB, int nB, int th_low, float nnratio, int strict_lt, int check_ori,
                                 int32_t* matchB) {
  // DBoW2::FeatureVector = std::map<NodeId, vector<unsigned>>; addFeature push_back in feature order
  std::map<int, std::vector<int>> fvA, fvB;
  for (int i = 0; i < nA; i++) fvA[nodeA[i]].push_back(i);
  for (int j = 0; j < nB; j++) fvB[nodeB[j]].push_back(j);
  for (int j = 0; j < nB; j++) matchB[j] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  auto ita = fvA.begin();
  auto itb = fvB.begin();
  while (ita != fvA.end() && itb != fvB.end()) {
    if (ita->first == itb->first) {
      const std::vector<int>& ia = ita->second;
      const std::vector<int>& ib = itb->second;
      for (size_t k = 0; k < ia.size(); k++) {
        const int realA = ia[k];
        if (validA && !validA[realA]) continue;  // no MapPoint / bad
        const uint8_t* dA = descA + (size_t)realA * 32;
        int best1 = 256, bestIdx = -1, best2 = 256;
        for (size_t m = 0; m < ib.size(); m++) {
          const int realB = ib[m];
          if (matchB[realB] >= 0) continue;  // already matched (:288 / :717)
          if (validB && !validB[realB]) continue;  // KF-KF variant: F side needs a good MapPoint (:722-728)
          const int dist = orc_descriptor_distance(dA, descB + (size_t)realB * 32);
          if (dist < best1) {
            best2 = best1;
            best1 = dist;
            bestIdx = realB;
          } else if (dist < best2) {
            best2 = dist;
          }
        }
        const bool pass = strict_lt ? (best1 < th_low) : (best1 <= th_low);
        if (pass) {
          if ((float)best1 < nnratio * (float)best2) {
            matchB[bestIdx] = realA;
            if (check_ori) rotHist[rot_bin(angA[realA], angB[bestIdx])].push_back(bestIdx);
            nmatches++;
          }
        }
      }
      ++ita;
      ++itb;
    } else if (ita->first < itb->first) {
      ita = fvA.lower_bound(itb->first);
    } else {
      itb = fvB.lower_bound(ita->first);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (size_t j = 0; j < rotHist[i].size(); j++) {
        matchB[rotHist[i][j]] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

/* ------------------------------------------------------------------------------------------
 * Frame grid (src/Frame.cc:461-491 AssignFeaturesToGrid, :863-877 PosInGrid, :741-850 GetFeaturesInArea)
 * ------------------------------------------------------------------------------------------ */
namespace {
const int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:55,60

struct Grid {
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  float minX, minY, invW, invH;
  void build(const float* kpx, const float* kpy, int n, const orc_frame_geom* g) {
    minX = g->mnMinX;
    minY = g->mnMinY;
    invW = (float)GRID_COLS / (g->mnMaxX - g->mnMinX);  // src/Frame.cc:213-214
    invH = (float)GRID_ROWS / (g->mnMaxY - g->mnMinY);
    for (int i = 0; i < n; i++) {
      int px = (int)std::round((kpx[i] - minX) * invW);
      int py = (int)std::round((kpy[i] - minY) * invH);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      cell[px][py].push_back(i);
    }
  }
  void in_area(float x, float y, float r, int minLevel, int maxLevel, const float* kpx, const float* kpy,
               const int32_t* octave, std::vector<int>& out) const {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
        const std::vector<int>& v = cell[ix][iy];
        for (size_t j = 0; j < v.size(); j++) {
          const int id = v[j];
          if (bCheckLevels) {
            if (octave[id] < minLevel) continue;
            if (maxLevel >= 0)
              if (octave[id] > maxLevel) continue;
          }
          const float distx = kpx[id] - x;
          const float disty = kpy[id] - y;
          if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(id);
        }
      }
  }
};
}  // namespace

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — src/ORBmatcher.cc:515-643 (monocular
// initialisation): level-0 keypoints of F1 search a fixed window around their previous match in F2; a candidate already
// matched with a distance <= the new one is skipped (vMatchedDistance), a better match steals the feature.
extern "C" int orc_search_for_initialization(const float* prevx, const float* prevy, const int32_t* octave1, const float* angle1,
                                             const uint8_t* desc1, int n1, const float* kpx2, const float* kpy2,
                                             const int32_t* octave2, const float* angle2, const uint8_t* desc2, int n2,
                                             const orc_frame_geom* g, int window, int th_low, float nnratio, int check_ori,
                                             int32_t* match12) {
  Grid* grid = new Grid();
  grid->build(kpx2, kpy2, n2, g);
  int nmatches = 0;
  for (int i = 0; i < n1; i++) match12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> vMatchedDistance(n2, INT_MAX), vnMatches21(n2, -1), cand;
  for (int i1 = 0; i1 < n1; i1++) {
    const int level1 = octave1[i1];
    if (level1 > 0) continue;  // :537
    grid->in_area(prevx[i1], prevy[i1], (float)window, level1, level1, kpx2, kpy2, octave2, cand);
    if (cand.empty()) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (size_t c = 0; c < cand.size(); c++) {
      const int i2 = cand[c];
      const int dist = orc_descriptor_distance(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
      if (vMatchedDistance[i2] <= dist) continue;  // :563
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestIdx2 = i2;
      } else if (dist < bestDist2) {
        bestDist2 = dist;
      }
    }
    if (bestDist <= th_low) {
      if (bestDist < (float)bestDist2 * nnratio) {  // :581
        if (vnMatches21[bestIdx2] >= 0) {
          match12[vnMatches21[bestIdx2]] = -1;
          nmatches--;
        }
        match12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (check_ori) rotHist[rot_bin(angle1[i1], angle2[bestIdx2])].push_back(i1);
      }
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (size_t j = 0; j < rotHist[i].size(); j++) {
        const int idx1 = rotHist[i][j];
        if (match12[idx1] >= 0) {  // a stolen match is already gone (:624-628)
          match12[idx1] = -1;
          nmatches--;
        }
      }
    }
  }
  delete grid;
  return nmatches;  // the caller updates vbPrevMatched[i1] = F2.mvKeysUn[match12[i1]].pt (:636-638)
}

// Frame::GetFeaturesInArea (src/Frame.cc:741-852) on the grid of Frame::AssignFeaturesToGrid (:461-491): test hook for the
// candidate enumeration every projection matcher above relies on (order included).
extern "C" int orc_features_in_area(const float* kpx, const float* kpy, const int32_t* octave, int nf, const orc_frame_geom* g,
                                    float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
  Grid* grid = new Grid();
  grid->build(kpx, kpy, nf, g);
  std::vector<int> cand;
  grid->in_area(x, y, r, min_level, max_level, kpx, kpy, octave, cand);
  delete grid;
  if ((int)cand.size() > cap) return -1;
  std::copy(cand.begin(), cand.end(), out);
  return (int)cand.size();
}

// SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) — src/ORBmatcher.cc:1569-1728, from the
// point where the last frame's map points have been projected (u, v, invzc computed by the shim at :1607-1626).
extern "C" int orc_search_by_projection_last(const orc_proj_query* q, int nq, const float* kpx, const float* kpy,
                                             const int32_t* octave, const float* angle, const float* uright,
                                             const uint8_t* occupied_in, const uint8_t* desc, int nf,
                                             const orc_frame_geom* g, float th, int mode, int th_high, int check_ori,
                                             int32_t* match_cur) {
  Grid* grid = new Grid();
  grid->build(kpx, kpy, nf, g);
  std::vector<uint8_t> occupied(nf, 0);
  if (occupied_in) std::copy(occupied_in, occupied_in + nf, occupied.begin());
  for (int j = 0; j < nf; j++) match_cur[j] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> cand;
  for (int i = 0; i < nq; i++) {
    const float u = q[i].u, v = q[i].v, invzc = q[i].invz;
    if (invzc < 0) continue;
    if (u < g->mnMinX || u > g->mnMaxX) continue;
    if (v < g->mnMinY || v > g->mnMaxY) continue;
    const int nLastOctave = q[i].octave;
    const float radius = th * g->scale_factors[nLastOctave];
    if (mode == 1) grid->in_area(u, v, radius, nLastOctave, -1, kpx, kpy, octave, cand);
    else if (mode == 2) grid->in_area(u, v, radius, 0, nLastOctave, kpx, kpy, octave, cand);
    else grid->in_area(u, v, radius, nLastOctave - 1, nLastOctave + 1, kpx, kpy, octave, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (size_t c = 0; c < cand.size(); c++) {
      const int i2 = cand[c];
      if (occupied[i2]) continue;
      if (uright[i2] > 0) {
        const float ur = u - g->bf * invzc;
        const float er = std::fabs(ur - uright[i2]);
        if (er > radius) continue;
      }
      const int dist = orc_descriptor_distance(q[i].desc, desc + (size_t)i2 * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx2 = i2;
      }
    }
    if (bestDist <= th_high) {
      match_cur[bestIdx2] = i;
      occupied[bestIdx2] = q[i].has_obs ? 1 : 0;
      nmatches++;
      if (check_ori) rotHist[rot_bin(q[i].angle, angle[bestIdx2])].push_back(bestIdx2);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i != ind1 && i != ind2 && i != ind3) {
        for (size_t j = 0; j < rotHist[i].size(); j++) {
          match_cur[rotHist[i][j]] = -1;
          nmatches--;
        }
      }
    }
  }
  delete grid;
  return nmatches;
}

// SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) — src/ORBmatcher.cc:70-175.
extern "C" int orc_search_by_projection_map(const orc_map_query* q, int nq, const float* kpx, const float* kpy,
                                            const int32_t* octave, const float* uright, const uint8_t* occupied_in,
                                            const uint8_t* desc, int nf, const orc_frame_geom* g, float th, int th_high,
                                            float nnratio, int32_t* match_cur) {
  Grid* grid = new Grid();
  grid->build(kpx, kpy, nf, g);
  std::vector<uint8_t> occupied(nf, 0);
  if (occupied_in) std::copy(occupied_in, occupied_in + nf, occupied.begin());
  for (int j = 0; j < nf; j++) match_cur[j] = -1;
  int nmatches = 0;
  const bool bFactor = th != 1.0;  // :76 (float compared with a double literal)
  std::vector<int> cand;
  for (int i = 0; i < nq; i++) {
    if (!q[i].in_view) continue;  // :83-87
    const int nPredictedLevel = q[i].level;
    float r = (q[i].view_cos > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos :178-185 (float vs double literal)
    if (bFactor) r *= th;
    const float rwin = r * g->scale_factors[nPredictedLevel];
    grid->in_area(q[i].u, q[i].v, rwin, nPredictedLevel - 1, nPredictedLevel, kpx, kpy, octave, cand);  // :99-103
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (size_t c = 0; c < cand.size(); c++) {
      const int idx = cand[c];
      if (occupied[idx]) continue;  // :123-125
      if (uright[idx] > 0) {        // :128-139
        const float er = std::fabs(q[i].ur - uright[idx]);
        if (er > r * g->scale_factors[nPredictedLevel]) continue;
      }
      const int dist = orc_descriptor_distance(q[i].desc, desc + (size_t)idx * 32);
      if (dist < bestDist) {  // :147-160
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = octave[idx];
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = octave[idx];
        bestDist2 = dist;
      }
    }
    if (bestDist <= th_high) {  // :164-171
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      match_cur[bestIdx] = i;
      occupied[bestIdx] = q[i].has_obs ? 1 : 0;
      nmatches++;
    }
  }
  delete grid;
  return nmatches;
}

// Common search core of Fuse (:1020-1174, :1179-1310) and SearchByProjection(KeyFrame*, Scw, ...) (:388-512); see orb_oracle.h.
extern "C" int orc_search_windows(const orc_win_query* q, int nq, const float* kpx, const float* kpy, const int32_t* octave,
                                  const float* uright, const float* inv_level_sigma2, const uint8_t* occupied_in,
                                  const uint8_t* desc, int nf, const orc_frame_geom* g, int flags, int th_dist,
                                  int32_t* best_idx, int32_t* best_dist) {
  Grid* grid = new Grid();
  grid->build(kpx, kpy, nf, g);
  std::vector<uint8_t> occupied(nf, 0);
  if (occupied_in) std::copy(occupied_in, occupied_in + nf, occupied.begin());
  const bool chi2 = flags & ORC_WIN_CHI2, greedy = flags & ORC_WIN_GREEDY;
  int n = 0;
  std::vector<int> cand;
  for (int i = 0; i < nq; i++) {
    best_idx[i] = -1;
    if (best_dist) best_dist[i] = 256;
    if (!q[i].valid) continue;
    const float u = q[i].u, v = q[i].v, ur = q[i].ur;
    grid->in_area(u, v, q[i].radius, -1, -1, kpx, kpy, octave, cand);  // KeyFrame::GetFeaturesInArea: no level filter
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx = -1;
    for (size_t c = 0; c < cand.size(); c++) {
      const int idx = cand[c];
      if (greedy && occupied[idx]) continue;  // :462-463
      const int kpLevel = octave[idx];
      if (kpLevel < q[i].min_level || kpLevel > q[i].max_level) continue;  // :1093-1094
      if (chi2) {  // :1097-1124
        if (uright[idx] >= 0) {
          const float ex = u - kpx[idx], ey = v - kpy[idx], er = ur - uright[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
        } else {
          const float ex = u - kpx[idx], ey = v - kpy[idx];
          const float e2 = ex * ex + ey * ey;
          if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
        }
      }
      const int dist = orc_descriptor_distance(q[i].desc, desc + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx = idx;
      }
    }
    if (best_dist) best_dist[i] = bestDist;
    if (bestDist <= th_dist) {  // :1137 / :498
      best_idx[i] = bestIdx;
      if (greedy) occupied[bestIdx] = 1;
      n++;
    }
  }
  delete grid;
  return n;
}

// CheckDistEpipolarLine (src/ORBmatcher.cc:186-215)
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, int oct2, const float* F12,
                                     const float* level_sigma2) {
  const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
  const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
  const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
  const float num = a * x2 + b * y2 + c;
  const float den = a * a + b * b;
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * level_sigma2[oct2];
}

// SearchForTriangulation (src/ORBmatcher.cc:810-1009)
extern "C" int orc_search_for_triangulation(const orc_kf_features* kf1, const orc_kf_features* kf2, const float* F12,
                                            float ex, float ey, const float* scale_factors, const float* level_sigma2,
                                            int only_stereo, int check_ori, int32_t* match12) {
  const int N1 = kf1->n, N2 = kf2->n;
  const int TH_LOW = 50;  // include/ORBmatcher.h:199
  // FeatureVector = std::map<NodeId, std::vector<unsigned>> filled in ascending feature index
  std::map<int, std::vector<int> > fv1, fv2;
  for (int i = 0; i < N1; i++) fv1[kf1->node[i]].push_back(i);
  for (int i = 0; i < N2; i++) fv2[kf2->node[i]].push_back(i);
  int nmatches = 0;
  std::vector<bool> vbMatched2(N2, false);
  std::vector<int> vMatches12(N1, -1);
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = HISTO_LENGTH / 360.0f;
  auto f1it = fv1.begin(), f2it = fv2.begin();
  while (f1it != fv1.end() && f2it != fv2.end()) {
    if (f1it->first == f2it->first) {
      for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
        const int idx1 = f1it->second[i1];
        if (kf1->has_mp[idx1]) continue;  // :852-855
        const bool bStereo1 = kf1->stereo[idx1] != 0;
        if (only_stereo && !bStereo1) continue;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (size_t i2 = 0; i2 < f2it->second.size(); i2++) {
          const int idx2 = f2it->second[i2];
          if (vbMatched2[idx2] || kf2->has_mp[idx2]) continue;  // :877-878
          const bool bStereo2 = kf2->stereo[idx2] != 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = orc_descriptor_distance(kf1->desc + (size_t)idx1 * 32, kf2->desc + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > bestDist) continue;  // :893
          if (!bStereo1 && !bStereo2) {                    // :899-906
            const float distex = ex - kf2->x[idx2];
            const float distey = ey - kf2->y[idx2];
            if (distex * distex + distey * distey < 100 * scale_factors[kf2->octave[idx2]]) continue;
          }
          if (check_dist_epipolar_line(kf1->x[idx1], kf1->y[idx1], kf2->x[idx2], kf2->y[idx2], kf2->octave[idx2], F12,
                                       level_sigma2)) {
            bestIdx2 = idx2;
            bestDist = dist;
          }
        }
        if (bestIdx2 >= 0) {
          vMatches12[idx1] = bestIdx2;
          vbMatched2[bestIdx2] = true;
          nmatches++;
          if (check_ori) {
            float rot = kf1->angle[idx1] - kf2->angle[bestIdx2];
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(idx1);
          }
        }
      }
      ++f1it;
      ++f2it;
    } else if (f1it->first < f2it->first) {
      f1it = fv1.lower_bound(f2it->first);
    } else {
      f2it = fv2.lower_bound(f1it->first);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (size_t j = 0; j < rotHist[i].size(); j++) {
        vMatches12[rotHist[i][j]] = -1;
        nmatches--;
      }
    }
  }
  for (int i = 0; i < N1; i++) match12[i] = vMatches12[i];
  return nmatches;
}

// FORB::distance (Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101): the bit-trick popcount of the 8 XOR words
static int forb_distance(const uint8_t* a, const uint8_t* b) {
  const int32_t* pa = reinterpret_cast<const int32_t*>(a);
  const int32_t* pb = reinterpret_cast<const int32_t*>(b);
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    unsigned int v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) (:1211-1256) for every feature
extern "C" int orc_bow_transform(const orc_vocabulary* v, const uint8_t* features, int n, int levelsup, int32_t* word_id,
                                 double* weight, int32_t* node_id) {
  const int N = v->n_nodes;
  std::vector<std::vector<int> > children(N);
  std::vector<int> word(N, 0);
  int nWords = 0;
  for (int nid = 1; nid < N; nid++) {  // loadFromTextFile :1378-1418
    children[v->parent[nid]].push_back(nid);
    if (v->leaf_flag[nid]) word[nid] = nWords++;
  }
  const int nid_level = v->L - levelsup;
  for (int i = 0; i < n; i++) {
    const uint8_t* feature = features + (size_t)i * 32;
    int nid = 0;  // if(nid_level <= 0 && nid != NULL) *nid = 0
    int final_id = 0, current_level = 0;
    if (children[0].empty()) {  // empty vocabulary
      word_id[i] = 0;
      weight[i] = 0;
      node_id[i] = 0;
      continue;
    }
    do {
      ++current_level;
      const std::vector<int>& nodes = children[final_id];
      final_id = nodes[0];
      double best_d = forb_distance(feature, v->desc + (size_t)final_id * 32);
      for (size_t c = 1; c < nodes.size(); c++) {
        const int id = nodes[c];
        const double d = forb_distance(feature, v->desc + (size_t)id * 32);
        if (d < best_d) {
          best_d = d;
          final_id = id;
        }
      }
      if (current_level == nid_level) nid = final_id;
    } while (!children[final_id].empty());
    word_id[i] = word[final_id];
    weight[i] = v->weight[final_id];
    node_id[i] = nid;
  }
  return nWords;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-440)
extern "C" void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx) {
  for (int p = 0; p < n_points; p++) {
    const int beg = offsets[p];
    const size_t N = (size_t)(offsets[p + 1] - beg);
    best_idx[p] = -1;
    if (N == 0) continue;  // :378-379
    std::vector<std::vector<float> > Distances(N, std::vector<float>(N, 0));
    for (size_t i = 0; i < N; i++) {
      Distances[i][i] = 0;
      for (size_t j = i + 1; j < N; j++) {
        const int distij = orc_descriptor_distance(desc + (size_t)(beg + i) * 32, desc + (size_t)(beg + j) * 32);
        Distances[i][j] = (float)distij;
        Distances[j][i] = (float)distij;
      }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (size_t i = 0; i < N; i++) {
      std::vector<int> vDists(Distances[i].begin(), Distances[i].end());
      std::sort(vDists.begin(), vDists.end());
      const int median = vDists[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) {
        BestMedian = median;
        BestIdx = (int)i;
      }
    }
    best_idx[p] = BestIdx;
  }
}

// TEST INFRASTRUCTURE — C entry points around the reference's own ORBmatcher (src/ORBmatcher.cc compiled in place against
// oracle/refshim: the cv stand-in and the Frame / KeyFrame / MapPoint data holders of slam_stubs.h).  Each call builds the
// holder objects from flat arrays, runs the UNMODIFIED reference function and flattens its result; where the reference
// projects map points itself (Fuse, Sim3, SearchByProjection) the glue also exports the post-projection queries, computed
// with the same expression sequence on the same cv stand-in, which is what the oracle / the CUDA path take as input.
// Built into oracle/_ref/libref_matcher.so (git-ignored); used only by tests/test_oracle_reference_matcher.py.
#include <cstdint>
#include <cstring>
#include <memory>

#include "ORBmatcher.h"

// The same glue also drives the product's drop-in ORBmatcher (self_commit_orb-slam2_b200/host/adapters/ORBmatcher_b200.cc):
// built with -DB2S_ADAPTER_BUILD into _ref/libadapter_matcher.so, the entry points are adp_* and ORBmatcher:: resolves to
// the adapter (gather -> libb200slam.so -> scatter) instead of the reference's src/ORBmatcher.cc.
#ifdef B2S_ADAPTER_BUILD
#define ref_search_by_bow_kf_f adp_search_by_bow_kf_f
#define ref_search_by_bow_kf_kf adp_search_by_bow_kf_kf
#define ref_search_by_projection_map adp_search_by_projection_map
#define ref_search_by_projection_last adp_search_by_projection_last
#define ref_search_by_projection_reloc adp_search_by_projection_reloc
#define ref_search_for_initialization adp_search_for_initialization
#define ref_search_for_triangulation adp_search_for_triangulation
#define ref_fuse adp_fuse
#define ref_search_by_projection_scw adp_search_by_projection_scw
#define ref_search_by_sim3 adp_search_by_sim3
#define ref_descriptor_distance adp_descriptor_distance
#endif

using namespace ORB_SLAM2;

float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

extern "C" {

typedef struct {
  float fx, fy, cx, cy, bf, b;
  float minX, minY, maxX, maxY;
  int32_t nlevels;
  float scaleFactor;
} ref_cam;

typedef struct {
  int32_t n;
  const float *x, *y, *angle;
  const int32_t* octave;
  const float* uright;
  const uint8_t* desc;
  const int32_t* node;    /* DBoW2 feature-vector node per feature, -1: none (may be NULL) */
  const int32_t* mp;      /* map point index per feature, -1: none (may be NULL) */
  const uint8_t* outlier; /* Frame::mvbOutlier (may be NULL) */
  const float* Tcw;       /* 16 floats row-major (may be NULL) */
} ref_feats;

typedef struct {
  int32_t n;
  const float *pos, *normal; /* n x 3 */
  const uint8_t* desc;       /* n x 32 */
  const uint8_t* bad;
  const int32_t* nobs;
  const float *minDist, *maxDist;                    /* mfMinDistance, mfMaxDistance */
  const float *trackX, *trackY, *trackXR, *viewCos; /* Tracking::SearchLocalPoints scratch (may be NULL) */
  const int32_t* trackLevel;
  const uint8_t* inView;
} ref_points;

/* post-projection query, layout of orc_win_query (oracle/orb_oracle.h) */
typedef struct {
  float u, v, ur, radius;
  int32_t min_level, max_level;
  uint8_t valid, pad[3];
  uint8_t desc[32];
} ref_win_query;

/* layout of orc_proj_query */
typedef struct {
  float u, v, invz, angle;
  int32_t octave, has_obs;
  uint8_t desc[32];
} ref_proj_query;
}

namespace {

cv::Mat mat_f(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) m.at<float>(i, j) = p[i * c + j];
  return m;
}

struct World {
  std::vector<std::unique_ptr<MapPoint> > pts;
  std::vector<MapPoint*> list;
  void build(const ref_points* p) {
    for (int i = 0; i < p->n; i++) {
      std::unique_ptr<MapPoint> m(new MapPoint());
      m->id = i;
      m->mWorldPos = mat_f(p->pos + 3 * i, 3, 1);
      m->mNormal = p->normal ? mat_f(p->normal + 3 * i, 3, 1) : cv::Mat(3, 1, CV_32F);
      m->mDescriptor = cv::Mat(1, 32, CV_8U);
      std::memcpy(m->mDescriptor.data, p->desc + 32 * (size_t)i, 32);
      m->mbBad = p->bad ? p->bad[i] != 0 : false;
      m->nObs = p->nobs ? p->nobs[i] : 1;
      m->mfMinDistance = p->minDist ? p->minDist[i] : 0.f;
      m->mfMaxDistance = p->maxDist ? p->maxDist[i] : 0.f;
      if (p->trackX) {
        m->mTrackProjX = p->trackX[i];
        m->mTrackProjY = p->trackY[i];
        m->mTrackProjXR = p->trackXR[i];
        m->mTrackViewCos = p->viewCos[i];
        m->mnTrackScaleLevel = p->trackLevel[i];
        m->mbTrackInView = p->inView[i] != 0;
      }
      list.push_back(m.get());
      pts.push_back(std::move(m));
    }
  }
};

void fill_holder(FeatureHolder& h, const ref_feats* f, const ref_cam* c, const World* w) {
  h.N = f->n;
  h.mvKeys.resize(f->n);
  h.mvuRight.assign(f->uright, f->uright + f->n);
  h.mDescriptors = cv::Mat(f->n, 32, CV_8U);
  if (f->n) std::memcpy(h.mDescriptors.data, f->desc, (size_t)f->n * 32);
  for (int i = 0; i < f->n; i++) {
    cv::KeyPoint& k = h.mvKeys[i];
    k.pt.x = f->x[i];
    k.pt.y = f->y[i];
    k.angle = f->angle[i];
    k.octave = f->octave[i];
    if (f->node && f->node[i] >= 0) h.mFeatVec[(DBoW2::NodeId)f->node[i]].push_back((unsigned)i);
  }
  h.mvKeysUn = h.mvKeys;  // no distortion: UndistortKeyPoints copies (src/Frame.cc:899-907)
  h.mvpMapPoints.assign(f->n, (MapPoint*)NULL);
  if (f->mp && w)
    for (int i = 0; i < f->n; i++)
      if (f->mp[i] >= 0) h.mvpMapPoints[i] = w->list[f->mp[i]];
  // scale tables as ORBextractor builds them (src/ORBextractor.cc:468-491) and Frame copies them (src/Frame.cc:376-384)
  h.mnScaleLevels = c->nlevels;
  h.mfScaleFactor = c->scaleFactor;
  h.mfLogScaleFactor = log(h.mfScaleFactor);
  h.mvScaleFactors.resize(c->nlevels);
  h.mvLevelSigma2.resize(c->nlevels);
  h.mvInvScaleFactors.resize(c->nlevels);
  h.mvInvLevelSigma2.resize(c->nlevels);
  h.mvScaleFactors[0] = 1.0f;
  h.mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < c->nlevels; i++) {
    h.mvScaleFactors[i] = h.mvScaleFactors[i - 1] * c->scaleFactor;
    h.mvLevelSigma2[i] = h.mvScaleFactors[i] * h.mvScaleFactors[i];
  }
  for (int i = 0; i < c->nlevels; i++) {
    h.mvInvScaleFactors[i] = 1.0f / h.mvScaleFactors[i];
    h.mvInvLevelSigma2[i] = 1.0f / h.mvLevelSigma2[i];
  }
  h.AssignFeaturesToGrid(c->minX, c->minY, c->maxX, c->maxY);
}

void set_statics(const ref_cam* c) {
  Frame::fx = c->fx;
  Frame::fy = c->fy;
  Frame::cx = c->cx;
  Frame::cy = c->cy;
  Frame::invfx = 1.0f / c->fx;
  Frame::invfy = 1.0f / c->fy;
  Frame::mnMinX = c->minX;
  Frame::mnMinY = c->minY;
  Frame::mnMaxX = c->maxX;
  Frame::mnMaxY = c->maxY;
}

void fill_frame(Frame& F, const ref_feats* f, const ref_cam* c, const World* w) {
  set_statics(c);
  fill_holder(F, f, c, w);
  F.mbf = c->bf;
  F.mb = c->b;
  F.mvbOutlier.assign(f->n, false);
  if (f->outlier)
    for (int i = 0; i < f->n; i++) F.mvbOutlier[i] = f->outlier[i] != 0;
  if (f->Tcw) F.mTcw = mat_f(f->Tcw, 4, 4);
}

void fill_keyframe(KeyFrame& K, const ref_feats* f, const ref_cam* c, World* w, long id) {
  fill_holder(K, f, c, w);
  K.id = id;
  K.fx = c->fx;
  K.fy = c->fy;
  K.cx = c->cx;
  K.cy = c->cy;
  K.invfx = 1.0f / c->fx;
  K.invfy = 1.0f / c->fy;
  K.mbf = c->bf;
  K.mb = c->b;
  K.mnMinX = (int)c->minX;
  K.mnMinY = (int)c->minY;
  K.mnMaxX = (int)c->maxX;
  K.mnMaxY = (int)c->maxY;
  if (f->Tcw) {
    // KeyFrame::SetPose (src/KeyFrame.cc:105-125)
    K.Tcw = mat_f(f->Tcw, 4, 4);
    cv::Mat Rcw = K.Tcw.rowRange(0, 3).colRange(0, 3);
    cv::Mat tcw = K.Tcw.rowRange(0, 3).col(3);
    cv::Mat Rwc = Rcw.t();
    K.Ow = -Rwc * tcw;
  }
  if (w)
    for (int i = 0; i < f->n; i++)
      if (K.mvpMapPoints[i]) K.mvpMapPoints[i]->mObservations[&K] = (size_t)i;
}

// the window query the reference forms inside Fuse / SearchByProjection(KF, Scw) for one map point
// (src/ORBmatcher.cc:1046-1092, 1214-1262, 424-466): same expressions, same order, same cv stand-in
void window_query(MapPoint* pMP, KeyFrame* pKF, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, float th,
                  bool with_ur, ref_win_query* q) {
  std::memset(q, 0, sizeof(*q));
  std::memcpy(q->desc, pMP->mDescriptor.data, 32);
  cv::Mat p3Dw = pMP->GetWorldPos();
  cv::Mat p3Dc = Rcw * p3Dw + tcw;
  if (p3Dc.at<float>(2) < 0.0f) return;
  const float invz = 1.0 / p3Dc.at<float>(2);
  const float x = p3Dc.at<float>(0) * invz;
  const float y = p3Dc.at<float>(1) * invz;
  const float u = pKF->fx * x + pKF->cx;
  const float v = pKF->fy * y + pKF->cy;
  if (!pKF->IsInImage(u, v)) return;
  const float ur = u - pKF->mbf * invz;
  const float maxDistance = pMP->GetMaxDistanceInvariance();
  const float minDistance = pMP->GetMinDistanceInvariance();
  cv::Mat PO = p3Dw - Ow;
  const float dist3D = cv::norm(PO);
  if (dist3D < minDistance || dist3D > maxDistance) return;
  cv::Mat Pn = pMP->GetNormal();
  if (PO.dot(Pn) < 0.5 * dist3D) return;
  const int lvl = pMP->PredictScale(dist3D, pKF);
  q->u = u;
  q->v = v;
  q->ur = with_ur ? ur : 0.f;
  q->radius = th * pKF->mvScaleFactors[lvl];
  q->min_level = lvl - 1;
  q->max_level = lvl;
  q->valid = 1;
}

void sim3_parts(const float* Scw16, cv::Mat& Rcw, cv::Mat& tcw, cv::Mat& Ow) {
  cv::Mat Scw = mat_f(Scw16, 4, 4);
  cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
  const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
  Rcw = sRcw / scw;
  tcw = Scw.rowRange(0, 3).col(3) / scw;
  Ow = -Rcw.t() * tcw;
}

}  // namespace

extern "C" {

/* SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) — src/ORBmatcher.cc:230-383.  matchF[j] = keyframe feature whose map
 * point the frame feature j received, -1: none. */
int ref_search_by_bow_kf_f(const ref_cam* c, const ref_feats* kf, const ref_points* pts, const ref_feats* f, float nnratio,
                           int check_ori, int32_t* matchF) {
  World w;
  w.build(pts);
  std::unique_ptr<KeyFrame> K(new KeyFrame());
  std::unique_ptr<Frame> F(new Frame());
  fill_keyframe(*K, kf, c, &w, 0);
  fill_frame(*F, f, c, NULL);
  ORBmatcher m(nnratio, check_ori != 0);
  std::vector<MapPoint*> out;
  int n = m.SearchByBoW(K.get(), *F, out);
  for (int j = 0; j < f->n; j++) matchF[j] = out[j] ? (int)out[j]->mObservations[K.get()] : -1;
  return n;
}

/* SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) — :656-805.  match12[i1] = feature of kf2, -1: none. */
int ref_search_by_bow_kf_kf(const ref_cam* c, const ref_feats* kf1, const ref_feats* kf2, const ref_points* pts, float nnratio,
                            int check_ori, int32_t* match12) {
  World w;
  w.build(pts);
  std::unique_ptr<KeyFrame> K1(new KeyFrame()), K2(new KeyFrame());
  fill_keyframe(*K1, kf1, c, &w, 0);
  fill_keyframe(*K2, kf2, c, &w, 1);
  ORBmatcher m(nnratio, check_ori != 0);
  std::vector<MapPoint*> out;
  int n = m.SearchByBoW(K1.get(), K2.get(), out);
  for (int i = 0; i < kf1->n; i++) match12[i] = out[i] ? (int)out[i]->mObservations[K2.get()] : -1;
  return n;
}

/* SearchByProjection(Frame&, vpMapPoints, th) — :70-175.  The frame's features hold map point `f->mp` beforehand (the
 * occupied ones); match_cur[j] = index into pts of the map point feature j holds afterwards if the call changed it, else -1. */
int ref_search_by_projection_map(const ref_cam* c, const ref_feats* f, const ref_points* pts, const int32_t* query_points,
                                 int nq, float th, float nnratio, int32_t* match_cur) {
  World w;
  w.build(pts);
  std::unique_ptr<Frame> F(new Frame());
  fill_frame(*F, f, c, &w);
  std::vector<MapPoint*> before = F->mvpMapPoints, vp;
  for (int i = 0; i < nq; i++) vp.push_back(w.list[query_points[i]]);
  ORBmatcher m(nnratio, true);
  int n = m.SearchByProjection(*F, vp, th);
  for (int j = 0; j < f->n; j++)
    match_cur[j] = (F->mvpMapPoints[j] && F->mvpMapPoints[j] != before[j]) ? (int)F->mvpMapPoints[j]->id : -1;
  return n;
}

/* SearchByProjection(CurrentFrame, LastFrame, th, bMono) — :1569-1728.  match_cur[j] = last-frame feature index whose map
 * point feature j received, -1: none.  queries[i] (one per last-frame feature, octave = -1 where the reference skips it
 * before projecting) and *mode are what the oracle takes. */
int ref_search_by_projection_last(const ref_cam* c, const ref_feats* cur, const ref_feats* last, const ref_points* pts, float th,
                                  int mono, int check_ori, int32_t* match_cur, ref_proj_query* queries, int32_t* mode) {
  World w;
  w.build(pts);
  std::unique_ptr<Frame> C(new Frame()), L(new Frame());
  fill_frame(*C, cur, c, &w);
  fill_frame(*L, last, c, &w);
  // queries: :1582-1626
  {
    const cv::Mat Rcw = C->mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = C->mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = L->mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = L->mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > C->mb && !mono;
    const bool bBackward = -tlc.at<float>(2) > C->mb && !mono;
    *mode = bForward ? 1 : (bBackward ? 2 : 0);
    for (int i = 0; i < last->n; i++) {
      ref_proj_query* q = queries + i;
      std::memset(q, 0, sizeof(*q));
      q->octave = -1;
      MapPoint* pMP = L->mvpMapPoints[i];
      if (!pMP || L->mvbOutlier[i]) continue;
      cv::Mat x3Dw = pMP->GetWorldPos();
      cv::Mat x3Dc = Rcw * x3Dw + tcw;
      const float xc = x3Dc.at<float>(0);
      const float yc = x3Dc.at<float>(1);
      const float invzc = 1.0 / x3Dc.at<float>(2);
      q->u = Frame::fx * xc * invzc + Frame::cx;
      q->v = Frame::fy * yc * invzc + Frame::cy;
      q->invz = invzc;
      q->angle = L->mvKeysUn[i].angle;
      q->octave = L->mvKeys[i].octave;
      q->has_obs = pMP->Observations() > 0;
      std::memcpy(q->desc, pMP->mDescriptor.data, 32);
    }
  }
  ORBmatcher m(0.9f, check_ori != 0);
  std::vector<MapPoint*> before = C->mvpMapPoints;
  int n = m.SearchByProjection(*C, *L, th, mono != 0);
  for (int j = 0; j < cur->n; j++) {
    MapPoint* p = C->mvpMapPoints[j];
    match_cur[j] = -1;
    if (p && p != before[j])
      for (int i = 0; i < last->n; i++)
        if (L->mvpMapPoints[i] == p) {
          match_cur[j] = i;
          break;
        }
  }
  return n;
}

/* SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist) — :1731-1862 (Tracking::Relocalization).
 * already_found[i] != 0: the keyframe's i-th map point is in sAlreadyFound.  match_cur[j] = keyframe feature index whose
 * map point the frame feature j received, -1.  queries[i] (one per keyframe feature; octave = predicted level, -1 where
 * the reference skips the point before the search) are what orc_search_by_projection_last takes with mode 0, every
 * query has_obs = 1, occupied = "holds any map point", no stereo gate (uright = -1) and th_high = ORBdist. */
int ref_search_by_projection_reloc(const ref_cam* c, const ref_feats* cur, const ref_feats* kf, const ref_points* pts,
                                   const uint8_t* already_found, float th, int orb_dist, int check_ori, int32_t* match_cur,
                                   ref_proj_query* queries) {
  World w;
  w.build(pts);
  std::unique_ptr<Frame> C(new Frame());
  std::unique_ptr<KeyFrame> K(new KeyFrame());
  fill_frame(*C, cur, c, &w);
  fill_keyframe(*K, kf, c, &w, 0);
  std::set<MapPoint*> found;
  for (int i = 0; i < kf->n; i++)
    if (already_found[i] && K->mvpMapPoints[i]) found.insert(K->mvpMapPoints[i]);
  {
    const cv::Mat Rcw = C->mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = C->mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    for (int i = 0; i < kf->n; i++) {
      ref_proj_query* q = queries + i;
      std::memset(q, 0, sizeof(*q));
      q->octave = -1;
      MapPoint* pMP = K->mvpMapPoints[i];
      if (!pMP || pMP->isBad() || found.count(pMP)) continue;
      cv::Mat x3Dw = pMP->GetWorldPos();
      cv::Mat x3Dc = Rcw * x3Dw + tcw;
      const float xc = x3Dc.at<float>(0);
      const float yc = x3Dc.at<float>(1);
      const float invzc = 1.0 / x3Dc.at<float>(2);
      const float u = Frame::fx * xc * invzc + Frame::cx;
      const float v = Frame::fy * yc * invzc + Frame::cy;
      cv::Mat PO = x3Dw - Ow;
      float dist3D = cv::norm(PO);
      if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
      q->u = u;
      q->v = v;
      q->invz = 1.0f;  // the reference neither tests the sign of the depth nor uses it here
      q->angle = K->mvKeysUn[i].angle;
      q->octave = pMP->PredictScale(dist3D, C.get());
      q->has_obs = 1;
      std::memcpy(q->desc, pMP->mDescriptor.data, 32);
    }
  }
  ORBmatcher m(0.75f, check_ori != 0);
  int n = m.SearchByProjection(*C, K.get(), found, th, orb_dist);
  for (int j = 0; j < cur->n; j++) {
    MapPoint* p = C->mvpMapPoints[j];
    match_cur[j] = -1;
    if (p && (cur->mp == NULL || cur->mp[j] < 0)) match_cur[j] = (int)p->mObservations.at(K.get());
  }
  return n;
}

/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — :515-643.  prev: n1 x 2, updated in place. */
int ref_search_for_initialization(const ref_cam* c, const ref_feats* f1, const ref_feats* f2, float* prev, int window,
                                  float nnratio, int check_ori, int32_t* match12) {
  std::unique_ptr<Frame> F1(new Frame()), F2(new Frame());
  fill_frame(*F1, f1, c, NULL);
  fill_frame(*F2, f2, c, NULL);
  std::vector<cv::Point2f> vprev(f1->n);
  for (int i = 0; i < f1->n; i++) vprev[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher m(nnratio, check_ori != 0);
  int n = m.SearchForInitialization(*F1, *F2, vprev, m12, window);
  for (int i = 0; i < f1->n; i++) {
    match12[i] = m12[i];
    prev[2 * i] = vprev[i].x;
    prev[2 * i + 1] = vprev[i].y;
  }
  return n;
}

/* SearchForTriangulation — :810-1009.  match12[i1] = i2 or -1; epipole[2] = (ex, ey) of :821-823. */
int ref_search_for_triangulation(const ref_cam* c, const ref_feats* kf1, const ref_feats* kf2, const ref_points* pts,
                                 const float* F12, int only_stereo, int check_ori, int32_t* match12, float* epipole) {
  World w;
  w.build(pts);
  std::unique_ptr<KeyFrame> K1(new KeyFrame()), K2(new KeyFrame());
  fill_keyframe(*K1, kf1, c, &w, 0);
  fill_keyframe(*K2, kf2, c, &w, 1);
  {
    cv::Mat Cw = K1->GetCameraCenter();
    cv::Mat R2w = K2->GetRotation();
    cv::Mat t2w = K2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    epipole[0] = K2->fx * C2.at<float>(0) * invz + K2->cx;
    epipole[1] = K2->fy * C2.at<float>(1) * invz + K2->cy;
  }
  ORBmatcher m(0.6f, check_ori != 0);
  std::vector<std::pair<size_t, size_t> > pairs;
  int n = m.SearchForTriangulation(K1.get(), K2.get(), mat_f(F12, 3, 3), pairs, only_stereo != 0);
  for (int i = 0; i < kf1->n; i++) match12[i] = -1;
  for (size_t k = 0; k < pairs.size(); k++) match12[pairs[k].first] = (int)pairs[k].second;
  return n;
}

/* Fuse(KeyFrame*, vpMapPoints, th) — :1020-1174 (Scw == NULL) and Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) —
 * :1179-1310.  query_points index pts (-1: a NULL entry).  best_idx[i] = keyframe feature the i-th point was fused with
 * (observation added, or replacement requested with the point that feature holds), -1: none.  queries: post-projection. */
int ref_fuse(const ref_cam* c, const ref_feats* kf, const ref_points* pts, const int32_t* query_points, int nq, const float* Scw,
             float th, int32_t* best_idx, ref_win_query* queries) {
  World w;
  w.build(pts);
  std::unique_ptr<KeyFrame> K(new KeyFrame());
  fill_keyframe(*K, kf, c, &w, 0);
  std::vector<MapPoint*> vp;
  for (int i = 0; i < nq; i++) vp.push_back(query_points[i] >= 0 ? w.list[query_points[i]] : (MapPoint*)NULL);
  cv::Mat Rcw, tcw, Ow;
  if (Scw)
    sim3_parts(Scw, Rcw, tcw, Ow);
  else {
    Rcw = K->GetRotation();
    tcw = K->GetTranslation();
    Ow = K->GetCameraCenter();
  }
  const std::set<MapPoint*> inKF = K->GetMapPoints();
  for (int i = 0; i < nq; i++) {
    std::memset(&queries[i], 0, sizeof(ref_win_query));
    MapPoint* p = vp[i];
    if (!p || p->isBad()) continue;
    if (Scw ? inKF.count(p) != 0 : p->IsInKeyFrame(K.get())) continue;
    window_query(p, K.get(), Rcw, tcw, Ow, th, Scw == NULL, &queries[i]);
  }
  stub_log().clear();
  ORBmatcher m(0.6f, true);
  int n;
  std::vector<MapPoint*> repl(nq, (MapPoint*)NULL);
  if (Scw)
    n = m.Fuse(K.get(), mat_f(Scw, 4, 4), vp, th, repl);
  else
    n = m.Fuse(K.get(), vp, th);
  for (int i = 0; i < nq; i++) best_idx[i] = -1;
  std::map<const void*, int> qof;
  for (int i = 0; i < nq; i++)
    if (vp[i]) qof[vp[i]] = i;
  for (const StubLogEntry& e : stub_log()) {
    if (e.kind == 0) best_idx[qof[e.a]] = (int)e.idx;  // AddObservation(pKF, bestIdx)
    if (e.kind == 2) {                                  // a->Replace(b): one is the query point, the other sits in the keyframe
      const MapPoint* a = (const MapPoint*)e.a;
      const MapPoint* b = (const MapPoint*)e.b;
      const bool a_is_query = qof.count(a) && !a->mObservations.count(K.get());
      const MapPoint* qp = a_is_query ? a : b;
      const MapPoint* kp = a_is_query ? b : a;
      best_idx[qof[qp]] = (int)kp->mObservations.at(K.get());
    }
  }
  for (int i = 0; i < nq; i++)
    if (repl[i]) best_idx[i] = (int)repl[i]->mObservations.at(K.get());
  return n;
}

/* SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) — :388-512.  matched_in[j] = index into pts of the point feature j
 * already holds (-1: none); best_idx[i] = feature the i-th point was matched to, -1: none. */
int ref_search_by_projection_scw(const ref_cam* c, const ref_feats* kf, const ref_points* pts, const int32_t* query_points, int nq,
                                 const int32_t* matched_in, const float* Scw, int th, int32_t* best_idx, ref_win_query* queries) {
  World w;
  w.build(pts);
  std::unique_ptr<KeyFrame> K(new KeyFrame());
  fill_keyframe(*K, kf, c, &w, 0);
  std::vector<MapPoint*> vp, matched(kf->n, (MapPoint*)NULL);
  for (int i = 0; i < nq; i++) vp.push_back(w.list[query_points[i]]);
  for (int j = 0; j < kf->n; j++)
    if (matched_in[j] >= 0) matched[j] = w.list[matched_in[j]];
  cv::Mat Rcw, tcw, Ow;
  sim3_parts(Scw, Rcw, tcw, Ow);
  std::set<MapPoint*> found(matched.begin(), matched.end());
  for (int i = 0; i < nq; i++) {
    std::memset(&queries[i], 0, sizeof(ref_win_query));
    if (vp[i]->isBad() || found.count(vp[i])) continue;
    window_query(vp[i], K.get(), Rcw, tcw, Ow, (float)th, false, &queries[i]);
  }
  std::vector<MapPoint*> before = matched;
  ORBmatcher m(0.75f, true);
  int n = m.SearchByProjection(K.get(), mat_f(Scw, 4, 4), vp, matched, th);
  for (int i = 0; i < nq; i++) best_idx[i] = -1;
  for (int j = 0; j < kf->n; j++)
    if (matched[j] && matched[j] != before[j])
      for (int i = 0; i < nq; i++)
        if (vp[i] == matched[j]) best_idx[i] = j;
  return n;
}

/* SearchBySim3 — :1314-1566.  matches12_in[i1] = feature of kf2 whose point is already matched to i1 (-1: none);
 * match12[i1] = kf2 feature newly matched, -1.  q12: kf1's points in kf2; q21: kf2's points in kf1 (post-projection). */
int ref_search_by_sim3(const ref_cam* c, const ref_feats* kf1, const ref_feats* kf2, const ref_points* pts,
                       const int32_t* matches12_in, float s12, const float* R12, const float* t12, float th, int32_t* match12,
                       ref_win_query* q12, ref_win_query* q21) {
  World w;
  w.build(pts);
  std::unique_ptr<KeyFrame> K1(new KeyFrame()), K2(new KeyFrame());
  fill_keyframe(*K1, kf1, c, &w, 0);
  fill_keyframe(*K2, kf2, c, &w, 1);
  std::vector<MapPoint*> vp12(kf1->n, (MapPoint*)NULL);
  for (int i = 0; i < kf1->n; i++)
    if (matches12_in[i] >= 0) vp12[i] = K2->mvpMapPoints[matches12_in[i]];
  cv::Mat mR12 = mat_f(R12, 3, 3), mt12 = mat_f(t12, 3, 1);
  {
    cv::Mat R1w = K1->GetRotation(), t1w = K1->GetTranslation(), R2w = K2->GetRotation(), t2w = K2->GetTranslation();
    cv::Mat sR12 = s12 * mR12;
    cv::Mat sR21 = (1.0 / s12) * mR12.t();
    cv::Mat t21 = -sR21 * mt12;
    std::vector<bool> done1(kf1->n, false), done2(kf2->n, false);
    for (int i = 0; i < kf1->n; i++)
      if (vp12[i]) {
        done1[i] = true;
        int idx2 = vp12[i]->GetIndexInKeyFrame(K2.get());
        if (idx2 >= 0 && idx2 < kf2->n) done2[idx2] = true;
      }
    for (int dir = 0; dir < 2; dir++) {
      KeyFrame* src = dir ? K2.get() : K1.get();
      KeyFrame* dst = dir ? K1.get() : K2.get();
      ref_win_query* out = dir ? q21 : q12;
      for (int i = 0; i < src->N; i++) {
        ref_win_query* q = out + i;
        std::memset(q, 0, sizeof(*q));
        MapPoint* pMP = src->mvpMapPoints[i];
        if (!pMP || (dir ? done2[i] : done1[i]) || pMP->isBad()) continue;
        std::memcpy(q->desc, pMP->mDescriptor.data, 32);
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat pd;
        if (!dir) {
          cv::Mat p3Dc1 = R1w * p3Dw + t1w;
          pd = sR21 * p3Dc1 + t21;
        } else {
          cv::Mat p3Dc2 = R2w * p3Dw + t2w;
          pd = sR12 * p3Dc2 + mt12;
        }
        if (pd.at<float>(2) < 0.0) continue;
        const float invz = 1.0 / pd.at<float>(2);
        const float x = pd.at<float>(0) * invz;
        const float y = pd.at<float>(1) * invz;
        const float u = K1->fx * x + K1->cx;
        const float v = K1->fy * y + K1->cy;
        if (!dst->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        const float dist3D = cv::norm(pd);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int lvl = pMP->PredictScale(dist3D, dst);
        q->u = u;
        q->v = v;
        q->radius = th * dst->mvScaleFactors[lvl];
        q->min_level = lvl - 1;
        q->max_level = lvl;
        q->valid = 1;
      }
    }
  }
  std::vector<MapPoint*> before = vp12;
  ORBmatcher m(0.75f, true);
  int n = m.SearchBySim3(K1.get(), K2.get(), vp12, s12, mR12, mt12, th);
  for (int i = 0; i < kf1->n; i++)
    match12[i] = (vp12[i] && vp12[i] != before[i]) ? (int)vp12[i]->mObservations.at(K2.get()) : -1;
  return n;
}

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
  std::memcpy(ma.data, a, 32);
  std::memcpy(mb.data, b, 32);
  return ORBmatcher::DescriptorDistance(ma, mb);
}
}

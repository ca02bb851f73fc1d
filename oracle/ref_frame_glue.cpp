// TEST INFRASTRUCTURE — C entry points around the reference's own Frame (src/Frame.cc compiled in place, unmodified, with
// src/ORBextractor.cc and src/ORBmatcher.cc, against oracle/refshim; MapPoint / KeyFrame / Converter / ORBVocabulary are
// the stand-ins of slam_stubs.h in B2S_STUB_REAL_FRAME mode).  A stereo Frame is built by the reference's constructor
// (src/Frame.cc:343-458: two extractor threads, UndistortKeyPoints, ComputeStereoMatches, AssignFeaturesToGrid); the glue
// exposes what it computed.  Built into oracle/_ref/libref_frame.so (git-ignored); used only by
// tests/test_oracle_reference_frame.py.
#include <cstdint>
#include <cstring>
#include <memory>

#include "Frame.h"
#include "ORBmatcher.h"

using namespace ORB_SLAM2;

namespace {
std::unique_ptr<ORBextractor> g_exL, g_exR;
std::unique_ptr<Frame> g_frame;
ORBVocabulary g_voc;

cv::Mat mat_f(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) m.at<float>(i, j) = p[i * c + j];
  return m;
}
void put_kp(const cv::KeyPoint& k, float* o) {
  o[0] = k.pt.x; o[1] = k.pt.y; o[2] = k.size; o[3] = k.angle; o[4] = k.response; o[5] = (float)k.octave; o[6] = (float)k.class_id;
}
}  // namespace

extern "C" {

/* Frame(imLeft, imRight, ...) — returns N (left keypoints); *n_right = right keypoints. */
int ref_frame_stereo(const uint8_t* L, const uint8_t* R, int w, int h, int nfeatures, float scaleFactor, int nlevels, int iniTh,
                     int minTh, float fx, float fy, float cx, float cy, float bf, float thDepth, int* n_right) {
  g_frame.reset();
  g_exL.reset(new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh));
  g_exR.reset(new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh));
  Frame::mbInitialComputations = true;  // recompute the image bounds / grid constants for this image size (:424-446)
  cv::Mat imL(h, w, CV_8UC1, (void*)L, (size_t)w), imR(h, w, CV_8UC1, (void*)R, (size_t)w);
  const float k[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  const float d[4] = {0, 0, 0, 0};
  cv::Mat K = mat_f(k, 3, 3), D = mat_f(d, 4, 1);
  g_frame.reset(new Frame(imL, imR, 0.0, g_exL.get(), g_exR.get(), &g_voc, K, D, bf, thDepth));
  *n_right = (int)g_frame->mvKeysRight.size();
  return g_frame->N;
}

/* kps: n x 7 floats (x y size angle response octave class_id); mb: the baseline member after construction */
void ref_frame_get(float* kpsL, uint8_t* descL, float* kpsR, uint8_t* descR, float* uright, float* depth, float* mb) {
  Frame& F = *g_frame;
  for (int i = 0; i < F.N; i++) {
    put_kp(F.mvKeys[i], kpsL + 7 * (size_t)i);
    std::memcpy(descL + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
    uright[i] = F.mvuRight[i];
    depth[i] = F.mvDepth[i];
  }
  for (size_t i = 0; i < F.mvKeysRight.size(); i++) {
    put_kp(F.mvKeysRight[i], kpsR + 7 * i);
    std::memcpy(descR + 32 * i, F.mDescriptorsRight.ptr((int)i), 32);
  }
  *mb = F.mb;
}

/* ORBextractor::mvImagePyramid[level] of the left / right extractor, copied densely (stride = width) */
int ref_frame_level(int right, int level, uint8_t* out, int* w, int* h) {
  const cv::Mat& m = (right ? g_exR : g_exL)->mvImagePyramid[level];
  *w = m.cols;
  *h = m.rows;
  if (out)
    for (int r = 0; r < m.rows; r++) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), m.cols);
  return 0;
}

/* Frame::GetFeaturesInArea (src/Frame.cc:741-852) on the frame built last */
int ref_frame_features_in_area(float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap) {
  std::vector<size_t> v = g_frame->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
  if ((int)v.size() > cap) return -1;
  for (size_t i = 0; i < v.size(); i++) out[i] = (int32_t)v[i];
  return (int)v.size();
}

/* Frame::SetPose + Frame::isInFrustum (src/Frame.cc:608-735) for n map points.  out: n x 4 (mTrackProjX, mTrackProjY,
 * mTrackProjXR, mTrackViewCos); level: mnTrackScaleLevel; in_view: the return value. */
void ref_frame_is_in_frustum(const float* Tcw16, const float* pos, const float* normal, const float* minDist, const float* maxDist,
                             int n, float cosLimit, float* out, int32_t* level, uint8_t* in_view) {
  g_frame->SetPose(mat_f(Tcw16, 4, 4));
  for (int i = 0; i < n; i++) {
    MapPoint p;
    p.mWorldPos = mat_f(pos + 3 * i, 3, 1);
    p.mNormal = mat_f(normal + 3 * i, 3, 1);
    p.mfMinDistance = minDist[i];
    p.mfMaxDistance = maxDist[i];
    in_view[i] = g_frame->isInFrustum(&p, cosLimit) ? 1 : 0;
    out[4 * i + 0] = p.mTrackProjX;
    out[4 * i + 1] = p.mTrackProjY;
    out[4 * i + 2] = p.mTrackProjXR;
    out[4 * i + 3] = p.mTrackViewCos;
    level[i] = p.mnTrackScaleLevel;
  }
}

/* Tracking::SearchLocalPoints on the frame built last (src/Tracking.cc:1166-1234): isInFrustum(pMP, 0.5) for every
 * point, then ORBmatcher(0.8).SearchByProjection(Frame, points, th).  match[j] = point index feature j received, -1. */
int ref_frame_search_local_points(const float* Tcw16, const float* pos, const float* normal, const float* minDist,
                                  const float* maxDist, const uint8_t* desc, int n, float th, int32_t* match, float* track,
                                  int32_t* level, uint8_t* in_view) {
  Frame& F = *g_frame;
  F.SetPose(mat_f(Tcw16, 4, 4));
  std::vector<std::unique_ptr<MapPoint> > own;
  std::vector<MapPoint*> vp;
  for (int i = 0; i < n; i++) {
    std::unique_ptr<MapPoint> p(new MapPoint());
    p->id = i;
    p->nObs = 1;
    p->mWorldPos = mat_f(pos + 3 * i, 3, 1);
    p->mNormal = mat_f(normal + 3 * i, 3, 1);
    p->mfMinDistance = minDist[i];
    p->mfMaxDistance = maxDist[i];
    p->mDescriptor = cv::Mat(1, 32, CV_8U);
    std::memcpy(p->mDescriptor.data, desc + 32 * (size_t)i, 32);
    in_view[i] = F.isInFrustum(p.get(), 0.5) ? 1 : 0;
    track[4 * i + 0] = p->mTrackProjX;
    track[4 * i + 1] = p->mTrackProjY;
    track[4 * i + 2] = p->mTrackProjXR;
    track[4 * i + 3] = p->mTrackViewCos;
    level[i] = p->mnTrackScaleLevel;
    vp.push_back(p.get());
    own.push_back(std::move(p));
  }
  F.mvpMapPoints.assign(F.N, (MapPoint*)NULL);
  ORBmatcher matcher(0.8);
  const int nm = matcher.SearchByProjection(F, vp, th);
  for (int j = 0; j < F.N; j++) match[j] = F.mvpMapPoints[j] ? (int)F.mvpMapPoints[j]->id : -1;
  F.mvpMapPoints.assign(F.N, (MapPoint*)NULL);
  return nm;
}

/* Frame::UnprojectStereo (src/Frame.cc:1478-1500) with the pose set by the last ref_frame_is_in_frustum call */
int ref_frame_unproject_stereo(int i, float* xyz) {
  cv::Mat m = g_frame->UnprojectStereo(i);
  if (m.empty()) return 0;
  for (int k = 0; k < 3; k++) xyz[k] = m.at<float>(k);
  return 1;
}
}

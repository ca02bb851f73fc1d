// TEST INFRASTRUCTURE — data-holder stand-ins for the reference's Frame / KeyFrame / MapPoint / DBoW2::FeatureVector so
// that /root/reference/src/ORBmatcher.cc compiles IN PLACE, unmodified, without the rest of the SLAM system (Map, Tracking,
// g2o, Pangolin, DBoW2 sources ...).  Pre-included with `-include`: it defines the reference headers' include guards
// (MAPPOINT_H, KEYFRAME_H, FRAME_H, __D_T_FEATURE_VECTOR__), so the real headers are opened and skipped.
//
// Only what ORBmatcher.cc touches is here.  The few methods with behaviour are restated from the reference:
//   Frame::PosInGrid / AssignFeaturesToGrid / GetFeaturesInArea        src/Frame.cc:461-491, 741-877
//   KeyFrame::GetFeaturesInArea / IsInImage                            src/KeyFrame.cc:752-802
//   MapPoint::PredictScale / Get{Min,Max}DistanceInvariance            src/MapPoint.cc:523-586
// Mutating calls (AddObservation, AddMapPoint, Replace) are recorded in a log the glue reads back.
#ifndef B2S_REF_SLAM_STUBS_H
#define B2S_REF_SLAM_STUBS_H
//
// Two modes.  Default: Frame is a stand-in too (ORBmatcher.cc alone).  With -DB2S_STUB_REAL_FRAME the reference's own
// Frame.h / Frame.cc and the vendored DBoW2 BowVector / FeatureVector are used, and only MapPoint, KeyFrame, Converter
// and ORBVocabulary are stand-ins (src/Frame.cc compiled in place: ComputeStereoMatches, isInFrustum, the feature grid).
#define MAPPOINT_H
#define KEYFRAME_H
#ifdef B2S_STUB_REAL_FRAME
#define CONVERTER_H
#define ORBVOCABULARY_H
#else
#define FRAME_H
#define __D_T_FEATURE_VECTOR__
#endif

#include <opencv2/core/core.hpp>

#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

using namespace std;  // the reference headers rely on it (ORBmatcher.h uses unqualified vector / pair)

#ifdef B2S_STUB_REAL_FRAME
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#else
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2
#endif

namespace ORB_SLAM2 {

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class KeyFrame;
class Frame;

struct StubLogEntry {
  int kind;  // 0 AddObservation, 1 AddMapPoint, 2 Replace
  const void* a;
  const void* b;
  long idx;
};
inline std::vector<StubLogEntry>& stub_log() {
  static std::vector<StubLogEntry> l;
  return l;
}

class MapPoint {
 public:
  cv::Mat mWorldPos, mNormal, mDescriptor;
  bool mbBad = false;
  int nObs = 0;
  float mfMinDistance = 0, mfMaxDistance = 0;
  std::map<KeyFrame*, size_t> mObservations;
  // Tracking::SearchLocalPoints scratch (include/MapPoint.h)
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0;
  bool mbTrackInView = false;
  int mnTrackScaleLevel = 0;
  long id = 0;

  bool isBad() { return mbBad; }
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  cv::Mat GetNormal() { return mNormal.clone(); }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  int Observations() { return nObs; }
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
  bool IsInKeyFrame(KeyFrame* kf) { return mObservations.count(kf) != 0; }
  int GetIndexInKeyFrame(KeyFrame* kf) { return mObservations.count(kf) ? (int)mObservations[kf] : -1; }
  void AddObservation(KeyFrame* kf, size_t idx) { stub_log().push_back({0, this, kf, (long)idx}); }
  void Replace(MapPoint* other) { stub_log().push_back({2, this, other, -1}); }
  template <class F>  // F = KeyFrame or Frame (instantiated at the call, when the type is complete)
  int PredictScale(const float& currentDist, F* f) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / f->mfLogScaleFactor);
    if (nScale < 0)
      nScale = 0;
    else if (nScale >= f->mnScaleLevels)
      nScale = f->mnScaleLevels - 1;
    return nScale;
  }
};

// what Frame and KeyFrame share for the matcher
class FeatureHolder {
 public:
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors;
  DBoW2::FeatureVector mFeatVec;
  std::vector<MapPoint*> mvpMapPoints;
  int mnScaleLevels = 0;
  float mfScaleFactor = 0, mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  std::vector<size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

  // src/Frame.cc:461-491 + 854-877 (KeyFrame copies the frame's grid, src/KeyFrame.cc:60-70)
  void AssignFeaturesToGrid(float minX, float minY, float maxX, float maxY) {
    mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(maxX - minX);
    mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(maxY - minY);
    for (int i = 0; i < N; i++) {
      const cv::KeyPoint& kp = mvKeysUn[i];
      int px = round((kp.pt.x - minX) * mfGridElementWidthInv);
      int py = round((kp.pt.y - minY) * mfGridElementHeightInv);
      if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
      mGrid[px][py].push_back(i);
    }
  }
  vector<size_t> Area(float minX, float minY, const float& x, const float& y, const float& r, const int minLevel,
                      const int maxLevel) const {
    vector<size_t> vIndices;
    const int nMinCellX = max(0, (int)floor((x - minX - r) * mfGridElementWidthInv));
    if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
    const int nMaxCellX = min((int)FRAME_GRID_COLS - 1, (int)ceil((x - minX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = max(0, (int)floor((y - minY - r) * mfGridElementHeightInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
    const int nMaxCellY = min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - minY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
        const vector<size_t>& cell = mGrid[ix][iy];
        for (size_t j = 0; j < cell.size(); j++) {
          const cv::KeyPoint& kpUn = mvKeysUn[cell[j]];
          if (bCheckLevels) {
            if (kpUn.octave < minLevel) continue;
            if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
          }
          const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
          if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(cell[j]);
        }
      }
    return vIndices;
  }
};

#ifdef B2S_STUB_REAL_FRAME
// Converter::toDescriptorVector (src/Converter.cc:37-47) and an ORBVocabulary whose transform records nothing: Frame.cc
// only needs them to compile (ComputeBoW is pinned separately through the vendored DBoW2, ref_dbow2_glue.cpp)
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> v;
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
  }
};
class ORBVocabulary {
 public:
  void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) {}
};
#else
class Frame : public FeatureHolder {
 public:
  static float fx, fy, cx, cy, invfx, invfy;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  float mbf = 0, mb = 0;
  cv::Mat mTcw;
  std::vector<bool> mvbOutlier;
  vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                   const int maxLevel = -1) const {
    return Area(mnMinX, mnMinY, x, y, r, minLevel, maxLevel);
  }
};
#endif

class KeyFrame : public FeatureHolder {
 public:
  float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;  // const int in include/KeyFrame.h
  cv::Mat Tcw, Ow;
  long id = 0;
  vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints)
      if (p && !p->isBad()) s.insert(p);
    return s;
  }
  cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
  cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
  void AddMapPoint(MapPoint* p, const size_t& idx) { stub_log().push_back({1, this, p, (long)idx}); }
  vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const {
    return Area((float)mnMinX, (float)mnMinY, x, y, r, -1, -1);
  }
};

}  // namespace ORB_SLAM2
#endif

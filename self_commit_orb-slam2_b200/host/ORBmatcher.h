// ORBmatcher.h — mirror of the part of /root/reference/include/ORBmatcher.h that is on the hot path, on flattened
// views (the Frame/KeyFrame/MapPoint pointer graph is gathered by the caller-side adapters shown in INTEGRATION.md).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

#include "../../include/b200slam.h"

namespace ORB_SLAM2 {

// What the adapters extract from a KeyFrame / Frame for SearchByBoW (src/ORBmatcher.cc:230-382, :656-799)
struct BowSide {
  const uint8_t* descriptors = nullptr;  // mDescriptors, N x 32
  const int32_t* node = nullptr;         // DBoW2 FeatureVector node id of every feature (mFeatVec flattened)
  const uint8_t* valid = nullptr;        // 1 if the feature has a MapPoint that is not bad (NULL: all valid)
  const float* angle = nullptr;          // mvKeysUn[i].angle / mvKeys[i].angle
  int n = 0;
};

class ORBmatcher;

// Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc:461-491, 741-852) kept on the device (b2s_frame_grid_*):
// built once per Frame / KeyFrame, it lets the consecutive projection matchers upload only their queries.
class FrameGrid {
 public:
  FrameGrid(ORBmatcher& matcher, const float* kpx, const float* kpy, const int32_t* octave, const float* angle,
            const float* uright, const uint8_t* descriptors, int nFeatures, const b2s_frame_geom& geom,
            const float* invLevelSigma2 = nullptr);
  ~FrameGrid();
  FrameGrid(const FrameGrid&) = delete;
  FrameGrid& operator=(const FrameGrid&) = delete;
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                        const int maxLevel = -1) const;
  int N() const { return mN; }
  b2s_frame_grid* handle() const { return mpGrid; }

 private:
  b2s_frame_grid* mpGrid = nullptr;
  int mN = 0;
};

class ORBmatcher {
 public:
  static const int TH_LOW = 50;        // src/ORBmatcher.cc:49-51
  static const int TH_HIGH = 100;
  static const int HISTO_LENGTH = 30;

  ORBmatcher(float nnratio = 0.6, bool checkOri = true);
  ~ORBmatcher();

  // static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) — src/ORBmatcher.cc:1913 (host popcount: a single
  // pair is not worth a launch; the batched device version is b2s_descriptor_distance)
  static int DescriptorDistance(const uint8_t* a, const uint8_t* b);

  // SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&): matchF[j] = keyframe feature index or -1; returns nmatches
  int SearchByBoW(const BowSide& kf, const BowSide& frame, std::vector<int32_t>& matchF);
  // SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&): strict '<' TH_LOW, both sides need MapPoints
  int SearchByBoWKF(const BowSide& kf1, const BowSide& kf2, std::vector<int32_t>& match2);
  // SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) after the projection of the last frame's map points
  int SearchByProjection(const std::vector<b2s_proj_query>& queries, const float* kpx, const float* kpy,
                         const int32_t* octave, const float* angle, const float* uright, const uint8_t* occupied,
                         const uint8_t* descriptors, int nFeatures, const b2s_frame_geom& geom, float th, int mode,
                         std::vector<int32_t>& matchCur);
  // SearchByProjection(Frame& Cur, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1731,
  // Tracking::Relocalization) after the projection of the keyframe's map points that are neither bad nor already found
  // and lie in the distance-invariance range: query.octave = MapPoint::PredictScale(dist3D, &Cur), query.angle =
  // pKF->mvKeysUn[i].angle.  It is the last-frame search with levels [l-1, l+1], no stereo gate, every assignment
  // occupying its feature and occupied[j] = (Cur.mvpMapPoints[j] != NULL); matchCur[j] = query index or -1.
  int SearchByProjectionReloc(std::vector<b2s_proj_query> queries, const float* kpx, const float* kpy, const int32_t* octave,
                              const float* angle, const uint8_t* occupied, const uint8_t* descriptors, int nFeatures,
                              const b2s_frame_geom& geom, float th, int ORBdist, std::vector<int32_t>& matchCur);
  // the same three searches on a resident FrameGrid (only queries and occupancy flags are uploaded)
  int SearchByProjection(const FrameGrid& grid, const std::vector<b2s_proj_query>& queries, const uint8_t* occupied, float th,
                         int mode, std::vector<int32_t>& matchCur);
  int SearchByProjection(const FrameGrid& grid, const std::vector<b2s_map_query>& mapPoints, const uint8_t* occupied, float th,
                         std::vector<int32_t>& matchF);
  int SearchWindows(const FrameGrid& grid, const std::vector<b2s_win_query>& mapPoints, const uint8_t* occupied, int flags,
                    std::vector<int32_t>& bestIdx);
  // SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, float th) (src/ORBmatcher.cc:70) after
  // Frame::isInFrustum filled the track fields of the local map points
  int SearchByProjection(const std::vector<b2s_map_query>& mapPoints, const float* kpx, const float* kpy,
                         const int32_t* octave, const float* uright, const uint8_t* occupied, const uint8_t* descriptors,
                         int nFeatures, const b2s_frame_geom& geom, float th, std::vector<int32_t>& matchF);
  // SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:515-643) on
  // flattened frames; vbPrevMatched (x, y interleaved) is updated in place like the reference does (:636-638)
  int SearchForInitialization(const float* kpx1, const float* kpy1, const int32_t* octave1, const float* angle1,
                              const uint8_t* descriptors1, int n1, const float* kpx2, const float* kpy2, const int32_t* octave2,
                              const float* angle2, const uint8_t* descriptors2, int n2, const b2s_frame_geom& geom,
                              std::vector<float>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
  // search core of Fuse(KeyFrame*, vpMapPoints, th) (B2S_WIN_CHI2), Fuse(KeyFrame*, Scw, ...) (no flag) and
  // SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (B2S_WIN_GREEDY): best keyframe feature per map point
  int SearchWindows(const std::vector<b2s_win_query>& mapPoints, const float* kpx, const float* kpy, const int32_t* octave,
                    const float* uright, const float* invLevelSigma2, const uint8_t* occupied, const uint8_t* descriptors,
                    int nFeatures, const b2s_frame_geom& geom, int flags, std::vector<int32_t>& bestIdx);
  // SearchForTriangulation(KeyFrame*, KeyFrame*, F12, vMatchedPairs, bOnlyStereo) on flattened keyframes; returns the
  // pairs (idx1, idx2) in ascending idx1 like the reference's vMatchedPairs
  int SearchForTriangulation(const b2s_kf_features& kf1, const b2s_kf_features& kf2, const float F12[9], float ex, float ey,
                             const float* scaleFactors2, const float* levelSigma2_2, int nLevels, bool bOnlyStereo,
                             std::vector<std::pair<size_t, size_t> >& vMatchedPairs);

  b2s_matcher* handle(int nFeatures) {  // (FrameGrid needs the handle sized for the frame)
    Ensure(nFeatures);
    return mpHandle;
  }

 protected:
  void Ensure(int n);
  float mfNNratio;
  bool mbCheckOrientation;
  b2s_matcher* mpHandle = nullptr;
  int mCap = 0;
};

}  // namespace ORB_SLAM2

// TEST INFRASTRUCTURE — C entry points around the reference's own MapPoint (src/MapPoint.cc compiled in place, unmodified,
// with the real MapPoint.h; KeyFrame / Frame / Map are the stand-ins of refshim/slam_stubs_mappoint.h).
// ORBmatcher::DescriptorDistance is not compiled here (src/ORBmatcher.cc needs the full KeyFrame); it forwards to
// orc_descriptor_distance, which tests/test_oracle_reference_matcher.py::test_descriptor_distance pins to the reference's.
// Built into oracle/_ref/libref_mappoint.so (git-ignored); used only by tests/test_oracle_reference_mappoint.py.
#include <cstdint>
#include <cstring>
#include <memory>

#include "MapPoint.h"
#include "ORBmatcher.h"
#include "orb_oracle.h"

namespace ORB_SLAM2 {
const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orc_descriptor_distance(a.ptr(), b.ptr()); }
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

namespace {
cv::Mat mat_f(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) m.at<float>(i, j) = p[i * c + j];
  return m;
}
}  // namespace

extern "C" {
/* For every map point p: observations [offsets[p], offsets[p+1]) = (keyframe id, feature index) pairs.  Keyframe k holds
 * descriptors kf_desc[kf_off[k] .. kf_off[k+1]) (32 B each), camera centre kf_ow[k], bad flag kf_bad[k]; its features'
 * octaves are kf_octave.  Runs AddObservation for every pair, then ComputeDistinctiveDescriptors and UpdateNormalAndDepth
 * (src/MapPoint.cc:359-440, 476-520).  Outputs per point: chosen descriptor, normal, min / max distance invariance. */
void ref_mappoint_update(int n_kf, const int32_t* kf_off, const uint8_t* kf_desc, const int32_t* kf_octave, const float* kf_ow,
                         const uint8_t* kf_bad, int n_points, const int32_t* offsets, const int32_t* obs_kf,
                         const int32_t* obs_idx, const float* pos, int nlevels, float scaleFactor, uint8_t* out_desc,
                         float* out_normal, float* out_minmax) {
  std::vector<KeyFrame> kfs(n_kf);  // one array: std::map<KeyFrame*, size_t> iterates in keyframe-id order
  for (int k = 0; k < n_kf; k++) {
    KeyFrame& K = kfs[k];
    const int n = kf_off[k + 1] - kf_off[k];
    K.mnId = k;
    K.mbBad = kf_bad[k] != 0;
    K.mDescriptors = cv::Mat(n, 32, CV_8U);
    std::memcpy(K.mDescriptors.data, kf_desc + 32 * (size_t)kf_off[k], 32 * (size_t)n);
    K.mvKeysUn.resize(n);
    K.mvuRight.assign(n, -1.0f);
    for (int i = 0; i < n; i++) K.mvKeysUn[i].octave = kf_octave[kf_off[k] + i];
    K.mnScaleLevels = nlevels;
    K.mfLogScaleFactor = log(scaleFactor);
    K.mvScaleFactors.resize(nlevels);
    K.mvScaleFactors[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) K.mvScaleFactors[i] = K.mvScaleFactors[i - 1] * scaleFactor;
    K.mOw = mat_f(kf_ow + 3 * k, 3, 1);
  }
  Map map;
  for (int p = 0; p < n_points; p++) {
    const int b = offsets[p], e = offsets[p + 1];
    std::memset(out_desc + 32 * (size_t)p, 0, 32);
    out_normal[3 * p] = out_normal[3 * p + 1] = out_normal[3 * p + 2] = 0;
    out_minmax[2 * p] = out_minmax[2 * p + 1] = 0;
    if (e == b) continue;
    MapPoint mp(mat_f(pos + 3 * p, 3, 1), &kfs[obs_kf[b]], &map);  // reference keyframe = first observation
    for (int q = b; q < e; q++) mp.AddObservation(&kfs[obs_kf[q]], (size_t)obs_idx[q]);
    mp.ComputeDistinctiveDescriptors();
    mp.UpdateNormalAndDepth();
    cv::Mat d = mp.GetDescriptor(), nrm = mp.GetNormal();
    if (!d.empty()) std::memcpy(out_desc + 32 * (size_t)p, d.data, 32);
    if (!nrm.empty())
      for (int c = 0; c < 3; c++) out_normal[3 * p + c] = nrm.at<float>(c);
    out_minmax[2 * p] = mp.GetMinDistanceInvariance();
    out_minmax[2 * p + 1] = mp.GetMaxDistanceInvariance();
  }
}
}

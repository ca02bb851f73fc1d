// ORBextractor.h — header-compatible mirror of /root/reference/include/ORBextractor.h (class ORB_SLAM2::ORBextractor):
// same constructor, operator(), getters and public mvImagePyramid, so src/Frame.cc:494-515 and src/Tracking.cc:179-192
// compile against it unchanged.  The body runs on the B200 through libb200slam.so (include/b200slam.h).
#pragma once
#include <vector>

#include "../../include/b200slam.h"
#include "cv_compat.h"

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
  // include/ORBextractor.h:92
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // include/ORBextractor.h:110 — mask is ignored, exactly like the reference (src/ORBextractor.cc:1544-1668)
  void operator()(b2s_cv::InputArray image, b2s_cv::InputArray mask, std::vector<b2s_cv::KeyPoint>& keypoints,
                  b2s_cv::OutputArray descriptors);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<b2s_cv::Mat> mvImagePyramid;  // filled after every operator() (Frame::ComputeStereoMatches reads it)

  // Device selection happens before the first image; B2S_DEVICE env var or SetDevice() (default 0).
  static void SetDevice(int device);
  // Skip the pyramid download when the caller does not need mvImagePyramid on the host (monocular / RGB-D frames).
  void SetDownloadPyramid(bool on) { mbDownloadPyramid = on; }

 protected:
  void EnsureHandle(int width, int height);
  int nfeatures;
  double scaleFactor;
  int nlevels, iniThFAST, minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  b2s_extractor* mpHandle = nullptr;
  int mMaxW = 0, mMaxH = 0;
  bool mbDownloadPyramid = true;
  std::vector<b2s_keypoint> mvTmpKeys;
};

}  // namespace ORB_SLAM2

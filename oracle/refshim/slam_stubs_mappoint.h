// TEST INFRASTRUCTURE — stand-ins for KeyFrame / Frame / Map so that /root/reference/src/MapPoint.cc compiles IN PLACE,
// unmodified, with the reference's own MapPoint.h (pre-included with `-include`; defines the KEYFRAME_H / FRAME_H / MAP_H
// include guards so the real headers are skipped).  Only what MapPoint.cc touches is here.
#ifndef B2S_REF_SLAM_STUBS_MAPPOINT_H
#define B2S_REF_SLAM_STUBS_MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H
#define MAP_H
#include <opencv2/core/core.hpp>

#include <climits>
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>

using namespace std;  // the reference headers rely on it

namespace ORB_SLAM2 {
class MapPoint;

struct FeatureTables {
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight, mvScaleFactors;
  int mnScaleLevels = 0;
  float mfLogScaleFactor = 0;
  cv::Mat mDescriptors, mOw;
  cv::Mat GetCameraCenter() { return mOw.clone(); }
};

class KeyFrame : public FeatureTables {
 public:
  long unsigned int mnId = 0, mnFrameId = 0;
  bool mbBad = false;
  std::vector<std::pair<size_t, MapPoint*> > log;  // EraseMapPointMatch / ReplaceMapPointMatch calls
  bool isBad() { return mbBad; }
  void EraseMapPointMatch(const size_t& idx) { log.push_back(std::make_pair(idx, (MapPoint*)NULL)); }
  void ReplaceMapPointMatch(const size_t& idx, MapPoint* p) { log.push_back(std::make_pair(idx, p)); }
};

class Frame : public FeatureTables {
 public:
  long unsigned int mnId = 0;
};

class Map {
 public:
  std::mutex mMutexPointCreation;
  std::vector<MapPoint*> erased;
  void EraseMapPoint(MapPoint* p) { erased.push_back(p); }
};
}  // namespace ORB_SLAM2
#endif

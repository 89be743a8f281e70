// ORACLE — TEST INFRASTRUCTURE ONLY.
// MapPoint::PredictScale (src/MapPoint.cc:491-509) and MapPoint::ComputeDistinctiveDescriptors (:314-378) of the REFERENCE compiled
// UNCHANGED: the function definitions are cut out of the source by name at build time (oracle/_ref/gen/mappoint_fns.inc); this file
// supplies the members they read (mfMaxDistance, mMutexPos, FrameBase::scalepyrinfo_; mObservations, mbBad, mDescriptor, the
// keyframes' descriptor rows, ORBmatcher::DescriptorDistance — itself cut from src/ORBmatcher.cc) and the `using namespace std`
// context of the reference's translation unit, which is what makes the unqualified log(ratio) / ceil(...) resolve to the FLOAT
// overloads.  ComputeDistinctiveDescriptors walks a std::map keyed by KeyFrame POINTERS: the order of the observed descriptors —
// and with it which of several rows with the same least median wins — follows the keyframes' addresses; the wrapper allocates the
// keyframes in one array so that address order == the order the caller lists them in.
#include <limits.h>
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>
using namespace std;
namespace cvmp {
class Mat {  // descriptor rows only
 public:
  const uint8_t* data = nullptr;
  Mat() {}
  explicit Mat(const uint8_t* d) : data(d) {}
  Mat row(int r) const { return Mat(data + 32 * (size_t)r); }
  Mat clone() const { return *this; }
  template <class T> const T* ptr() const { return (const T*)data; }
};
}  // namespace cvmp
#define cv cvmp
namespace VIEO_SLAM_MP {
class ORBmatcher {
 public:
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
#include "orbmatcher_fns.inc"
class KeyFrame {
 public:
  cv::Mat mDescriptors;
  bool isBad() { return false; }
};
struct FrameBase {
  struct _ScalePyramidInfo {  // include/FrameBase.h:183-188
    vector<float> vscalefactor_;
    float flogscalefactor_;
  } scalepyrinfo_;
};
class MapPoint {
 public:
  float mfMaxDistance = 0;
  mutex mMutexPos, mMutexFeatures;
  bool mbBad = false;
  std::map<KeyFrame*, std::set<size_t>> mObservations;
  cv::Mat mDescriptor;
  int PredictScale(const float& currentDist, FrameBase* pfb);
  void ComputeDistinctiveDescriptors();
};
#include "mappoint_fns.inc"
}  // namespace VIEO_SLAM_MP
#undef cv
namespace VIEO_SLAM = VIEO_SLAM_MP;

extern "C" int ref_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels) {
  VIEO_SLAM::FrameBase fb;
  fb.scalepyrinfo_.vscalefactor_.assign(n_levels, 1.0f);
  fb.scalepyrinfo_.flogscalefactor_ = log_scale_factor;
  VIEO_SLAM::MapPoint mp;
  mp.mfMaxDistance = max_distance;
  return mp.PredictScale(current_dist, &fb);
}

// same arguments as orc_distinctive_descriptors: per point the keyframe observations rows[ptr[p] .. ptr[p + 1]) of desc_pool, one
// keyframe per row in that order; best[p] = position of the chosen row in the point's list (-1: no observations)
extern "C" void ref_distinctive_descriptors(const uint8_t* desc_pool, const int32_t* rows, const int32_t* ptr, int n_points, int32_t* best) {
  using namespace VIEO_SLAM_MP;
  for (int p = 0; p < n_points; ++p) {
    const int b = ptr[p], N = ptr[p + 1] - b;
    best[p] = -1;
    std::vector<KeyFrame> kfs(N > 0 ? N : 0);  // one array: address order == list order
    MapPoint mp;
    for (int i = 0; i < N; ++i) {
      kfs[i].mDescriptors = cvmp::Mat(desc_pool + 32 * (size_t)(rows ? rows[b + i] : b + i));
      mp.mObservations[&kfs[i]].insert(0);
    }
    mp.ComputeDistinctiveDescriptors();
    for (int i = 0; i < N && best[p] < 0; ++i)
      if (mp.mDescriptor.data == kfs[i].mDescriptors.data) best[p] = i;
  }
}

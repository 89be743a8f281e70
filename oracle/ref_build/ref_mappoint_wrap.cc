// ORACLE — TEST INFRASTRUCTURE ONLY.
// MapPoint::PredictScale (src/MapPoint.cc:491-509) of the REFERENCE compiled UNCHANGED: the function definition is cut out of the
// source by name at build time (oracle/_ref/gen/mappoint_fns.inc); this file supplies the members it reads (mfMaxDistance, mMutexPos,
// FrameBase::scalepyrinfo_) and the `using namespace std` context of the reference's translation unit, which is what makes the
// unqualified log(ratio) / ceil(...) resolve to the FLOAT overloads.
#include <cmath>
#include <mutex>
#include <vector>
using namespace std;
namespace VIEO_SLAM {
struct FrameBase {
  struct _ScalePyramidInfo {  // include/FrameBase.h:183-188
    vector<float> vscalefactor_;
    float flogscalefactor_;
  } scalepyrinfo_;
};
class MapPoint {
 public:
  float mfMaxDistance = 0;
  mutex mMutexPos;
  int PredictScale(const float& currentDist, FrameBase* pfb);
};
#include "mappoint_fns.inc"
}  // namespace VIEO_SLAM

extern "C" int ref_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels) {
  VIEO_SLAM::FrameBase fb;
  fb.scalepyrinfo_.vscalefactor_.assign(n_levels, 1.0f);
  fb.scalepyrinfo_.flogscalefactor_ = log_scale_factor;
  VIEO_SLAM::MapPoint mp;
  mp.mfMaxDistance = max_distance;
  return mp.PredictScale(current_dist, &fb);
}

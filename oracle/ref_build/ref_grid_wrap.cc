// ORACLE — TEST INFRASTRUCTURE ONLY.
// FrameBase::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea / IsInImage (src/FrameBase.cpp:95-174) of the REFERENCE compiled
// UNCHANGED: the 64 x 48 cell grid behind every guided search — which cells a window touches, the column-major cell walk, the level
// band, the |dx| < r && |dy| < r box — i.e. the ORDER in which candidates reach the Hamming arg-min, which decides ties.  The four
// member-function definitions are cut out of the source by name at build time (oracle/_ref/gen/grid_fns.inc); this file supplies
// the members of FrameBase they read (include/FrameBase.h:105-233) and a TU-local cv::KeyPoint (`#define cv cvst_grid`).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <array>
#include <cassert>
#include <cmath>
#include <tuple>
#include <vector>
using namespace std;

namespace cvst_grid {
struct Point2f {
  float x, y;
};
struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
}  // namespace cvst_grid
#define cv cvst_grid
namespace VIEO_SLAM_GRID {
class FrameBase {
 public:
  typedef struct _GridInfo {  // include/FrameBase.h:221-231
    vector<float> fgrids_widthinv_;
    vector<float> fgrids_heightinv_;
    const int FRAME_GRID_ROWS = 48;
    const int FRAME_GRID_COLS = 64;
    vector<array<float, 4>> minmax_xy_;
  } GridInfo;
  GridInfo gridinfo_;
  vector<vector<vector<size_t>>> vgrids_;
  int N = 0;
  bool usedistort_ = false;
  vector<int> mpCameras;  // only emptiness / size are asked
  vector<cv::KeyPoint> mvKeys, mvKeysUn;
  vector<pair<size_t, size_t>> mapn2in_;
  vector<size_t> GetFeaturesInArea(uint8_t cami, const float& x, const float& y, const float& r, const int minlevel = -1,
                                   const int maxlevel = -1) const;
  void AssignFeaturesToGrid();
  bool PosInGrid(uint8_t cami, const cv::KeyPoint& kp, int& posX, int& posY);
  bool IsInImage(uint8_t cami, const float& x, const float& y) const;
};
#include "grid_fns.inc"
}  // namespace VIEO_SLAM_GRID
#undef cv

struct RefKp {  // == OrcKeyPoint
  float x, y, size, angle, response;
  int32_t octave;
};
// One frame's grid (bounds + inverse cell sizes as FrameBase::ComputeImageBounds leaves them) queried n_q times:
// q = (x, y, r, minlevel, maxlevel) rows; out_ptr [n_q + 1] / out_idx: the candidate lists in the reference's order.
extern "C" int ref_features_in_area(const RefKp* kps, int n_kp, float minx, float maxx, float miny, float maxy, float winv, float hinv,
                                    const float* q_xyr, const int32_t* q_levels, int n_q, int32_t* out_ptr, int32_t* out_idx, int cap,
                                    uint8_t* in_image) {
  VIEO_SLAM_GRID::FrameBase F;
  F.mpCameras.push_back(0);
  F.gridinfo_.fgrids_widthinv_ = {winv};
  F.gridinfo_.fgrids_heightinv_ = {hinv};
  F.gridinfo_.minmax_xy_.push_back({minx, maxx, miny, maxy});
  F.N = n_kp;
  F.mvKeysUn.resize(n_kp);
  for (int i = 0; i < n_kp; ++i) {
    F.mvKeysUn[i].pt.x = kps[i].x; F.mvKeysUn[i].pt.y = kps[i].y; F.mvKeysUn[i].octave = kps[i].octave;
  }
  F.mvKeys = F.mvKeysUn;
  F.AssignFeaturesToGrid();
  int total = 0;
  out_ptr[0] = 0;
  for (int q = 0; q < n_q; ++q) {
    const std::vector<size_t> v = F.GetFeaturesInArea(0, q_xyr[3 * q], q_xyr[3 * q + 1], q_xyr[3 * q + 2], q_levels[2 * q], q_levels[2 * q + 1]);
    for (size_t k = 0; k < v.size(); ++k) {
      if (total < cap) out_idx[total] = (int32_t)v[k];
      ++total;
    }
    out_ptr[q + 1] = total;
    if (in_image) in_image[q] = F.IsInImage(0, q_xyr[3 * q], q_xyr[3 * q + 1]) ? 1 : 0;
  }
  return total;
}

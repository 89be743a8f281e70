// ORACLE — TEST INFRASTRUCTURE ONLY.
// Frame::isInFrustum (src/Frame.cc:335-416) of the REFERENCE compiled UNCHANGED against the Eigen stand-in (eigstub/): the function
// definition is cut out of the source by name at build time (oracle/_ref/gen/frustum_fns.inc), together with
// MapPoint::GetMinDistanceInvariance / GetMaxDistanceInvariance (src/MapPoint.cc:481-489).  PredictScale and the cameras' Project()
// are the reference's own compiled functions of this library (ref_mappoint_wrap.cc, ref_camera_wrap.cc).  Declared by hand: the
// members the body reads — the pose as the three cv::Mat the reference keeps (a holder type whose rowRange / colRange / Converter
// conversions hand back the float values), the camera rig (GetTcr / GetTrc as a float quaternion + translation acting like
// Sophus::SE3f, toK of camera_pinhole.h:63-67), the grid bounds, the stereo baseline, the point's tracking-info lists.
#include <math.h>
#include <array>
#include <cassert>
#include <cmath>
#include <list>
#include <memory>
#include <mutex>
#include <vector>

#include "eigstub/mini_eigen.h"
#include "../oracle.h"  // OrcFrustumRigFrame / OrcFrustumCam layouts only

extern "C" void ref_cam_project(int model, const float* params, int n_params, const double P[3], float uv[2], double* J, double* Jp);
extern "C" int ref_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);

using namespace std;
namespace VIEO_SLAM_FRUSTUM {
using namespace Eigen;
using Vector2img = Eigen::Matrix<float, 2, 1>;

struct PoseMat {  // Tcw_ / mtcw / mOw: float values, read through Converter like the reference does
  Matrix3d R;
  Vector3d t;
  PoseMat rowRange(int, int) const { return *this; }
  PoseMat colRange(int, int) const { return *this; }
};
struct Converter {
  static Matrix3d toMatrix3d(const PoseMat& m) { return m.R; }
  static Vector3d toVector3d(const PoseMat& m) { return m.t; }
};
struct SE3fStandin {  // Sophus::SE3f: so3() * p + translation(), the rotation applied as Eigen's float _transformVector
  Quaternionf q;
  Vector3f t;
  Vector3f operator*(const Vector3f& p) const { return q._transformVector(p) + t; }
  const Vector3f& translation() const { return t; }
};
namespace camm {
class Camera {
 public:
  SE3fStandin Tcr, Trc;
  int model = 0;
  vector<float> params;
  const SE3fStandin& GetTcr() const { return Tcr; }
  const SE3fStandin& GetTrc() const { return Trc; }
  Matrix3f toK() const {  // camera_pinhole.h:63-67
    Matrix3f K;
    K << params[0], 0.f, params[2], 0.f, params[1], params[3], 0.f, 0.f, 1.f;
    return K;
  }
  void Project(const Vector3d& p, Vector2img* img) const {
    float uv[2];
    const double P[3] = {p(0), p(1), p(2)};
    ref_cam_project(model == 2 ? 2 : 0, params.data(), (int)params.size(), P, uv, nullptr, nullptr);
    (*img)(0) = uv[0], (*img)(1) = uv[1];
  }
};
}  // namespace camm

class Frame;
class MapPoint {
 public:
  using Vector3data = Eigen::Matrix<float, 3, 1>;
  struct TrackFastMatchInfo {
    float track_depth_ = INFINITY;
    bool btrack_inview_ = false;
    static constexpr int NUM_PROJ = 3;
    list<float> vtrack_proj_[NUM_PROJ];
    list<size_t> vtrack_cami_;
    list<float> vtrack_viewcos_;
    list<int> vtrack_scalelevel_;
    void Reset(Frame* = nullptr) {
      btrack_inview_ = false;
      for (auto& l : vtrack_proj_) l.clear();
      vtrack_cami_.clear(), vtrack_viewcos_.clear(), vtrack_scalelevel_.clear();
    }
  } trackinfo_;
  TrackFastMatchInfo& GetTrackInfoRef() { return trackinfo_; }
  Vector3data mWorldPos, mNormal;
  float mfMaxDistance = 0, mfMinDistance = 0;
  mutex mMutexPos;
  Vector3data GetWorldPos() { return mWorldPos; }
  Vector3f GetNormal() { return mNormal; }
  float GetMinDistanceInvariance();
  float GetMaxDistanceInvariance();
  int PredictScale(const float& currentDist, Frame* pF);
};
class Frame {
 public:
  PoseMat Tcw_, mtcw, mOw;
  vector<shared_ptr<camm::Camera>> mpCameras;
  bool usedistort_ = false;
  struct {
    vector<array<float, 4>> minmax_xy_;
  } gridinfo_;
  struct {
    float baseline_bf_[2] = {0, 0};
  } stereoinfo_;
  float log_scale_factor = 0;
  int n_levels = 0;
  bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
};
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
  return ref_predict_scale(mfMaxDistance, currentDist, pF->log_scale_factor, pF->n_levels);
}
#include "frustum_fns.inc"
}  // namespace VIEO_SLAM_FRUSTUM

// same arguments and outputs as orc_is_in_frustum_rig
extern "C" int ref_is_in_frustum_rig(const OrcFrustumRigFrame* f, int n, const float* wP, const float* Pn, const float* max_dist,
                                     const float* min_dist, uint8_t* inview, uint8_t* cam_mask, float* proj, int32_t* level,
                                     float* viewcos, float* depth) {
  using namespace VIEO_SLAM_FRUSTUM;
  Frame F;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) F.Tcw_.R(r, c) = f->Rcw[3 * r + c];
    F.mtcw.t(r) = f->tcw[r];
    F.mOw.t(r) = f->Ow[r];
  }
  F.stereoinfo_.baseline_bf_[1] = f->bf;
  F.log_scale_factor = f->log_scale_factor;
  F.n_levels = f->n_levels;
  for (int ci = 0; ci < f->n_cams; ++ci) {
    const OrcFrustumCam& c = f->cam[ci];
    auto cam = make_shared<camm::Camera>();
    cam->Tcr.q = Quaternionf(c.q_cr[3], c.q_cr[0], c.q_cr[1], c.q_cr[2]);
    cam->Tcr.t = Vector3f(c.t_cr[0], c.t_cr[1], c.t_cr[2]);
    cam->Trc.t = Vector3f(c.t_rc[0], c.t_rc[1], c.t_rc[2]);
    cam->model = c.model;
    cam->params = {c.fx, c.fy, c.cx, c.cy};
    if (c.model == 2) cam->params.insert(cam->params.end(), c.k, c.k + 4);
    if (c.model != 0) F.usedistort_ = true;
    F.mpCameras.push_back(cam);
    F.gridinfo_.minmax_xy_.push_back({c.minx, c.maxx, c.miny, c.maxy});
  }
  int n_in = 0;
  for (int i = 0; i < n; ++i) {
    MapPoint mp;
    mp.mWorldPos = Vector3f(wP[3 * i], wP[3 * i + 1], wP[3 * i + 2]);
    mp.mNormal = Vector3f(Pn[3 * i], Pn[3 * i + 1], Pn[3 * i + 2]);
    mp.mfMaxDistance = max_dist[i], mp.mfMinDistance = min_dist[i];
    const bool in = F.isInFrustum(&mp, f->cos_limit);
    inview[i] = in, cam_mask[i] = 0, depth[i] = in ? mp.trackinfo_.track_depth_ : 0.0f;
    for (int c = 0; c < 4; ++c) {
      level[4 * i + c] = -1, viewcos[4 * i + c] = 0;
      proj[12 * i + 3 * c] = proj[12 * i + 3 * c + 1] = proj[12 * i + 3 * c + 2] = 0;
    }
    auto& ti = mp.trackinfo_;
    assert(ti.btrack_inview_ == in);
    auto iu = ti.vtrack_proj_[0].begin(), iv = ti.vtrack_proj_[1].begin(), ir = ti.vtrack_proj_[2].begin();
    auto il = ti.vtrack_scalelevel_.begin();
    auto ic = ti.vtrack_viewcos_.begin();
    for (size_t c : ti.vtrack_cami_) {
      cam_mask[i] |= (uint8_t)(1u << c);
      proj[12 * i + 3 * c] = *iu++, proj[12 * i + 3 * c + 1] = *iv++, proj[12 * i + 3 * c + 2] = *ir++;
      level[4 * i + c] = *il++, viewcos[4 * i + c] = *ic++;
    }
    n_in += in;
  }
  return n_in;
}

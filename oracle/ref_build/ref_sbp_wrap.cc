// ORACLE — TEST INFRASTRUCTURE ONLY.
// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono, th_far_pts) (src/ORBmatcher.cc:1303-1467),
// SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist, th_far_pts) (:1471-1606), SearchByBoW(KeyFrame*, Frame&, ...) (:344-505),
// SearchByProjectionBase (:26-227) and ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th, th_far_pts) (:230-335) + RadiusByViewingCos (:337-342) of
// the REFERENCE compiled UNCHANGED, on top of the reference's own grid functions (FrameBase::GetFeaturesInArea / AssignFeaturesToGrid
// / PosInGrid / IsInImage) and ORBmatcher::DescriptorDistance / ComputeThreeMaxima, all cut out of the sources by name at build time.
// What is pinned: the control flow of the two tracking searches — forward / backward / +-1 level bands, the th_far and depth gates,
// the stereo ur gate, the "keypoint already holds an observed map point" rule, best / second-best with the same-level ratio test,
// the claim order, the rotation histogram and its three maxima.  What this file supplies (stand-ins, stated for what they are):
// the members of Frame / MapPoint / Camera the bodies touch, a TU-local cv::Mat / KeyPoint (`#define cv cvst_sbp`), three-element
// vectors, a 3 x 3 float matrix-vector product, and an SE3 (unit quaternion + translation) whose product / inverse / action are
// Sophus' formulas (so3 * p = Eigen's _transformVector; T1 * T2 = (q1 q2, t1 + q1 * t2); T^-1 = (q*, q* * (-t))) — the same
// formulas the oracle restates, so for these few lines the comparison is between two copies of one formula.
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <array>
#include <cassert>
#include <cmath>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <vector>
using namespace std;

#define PRINT_DEBUG_FILE_MUTEX(...)
#define PRINT_DEBUG_FILE(...)
#define CV_Assert(x) assert(x)

namespace cvst_sbp {
struct Point2f {
  float x, y;
};
struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
class Mat {  // descriptor rows, or (SearchByProjectionBase's Rcrw / tcrw / camera centre) a float rotation / vector read through Converter
 public:
  const uint8_t* data = nullptr;
  int rows = 0;
  float poseR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, poset[3] = {0, 0, 0};
  Mat() {}
  Mat(const uint8_t* d, int r) : data(d), rows(r) {}
  Mat row(int r) const { return Mat(data + 32 * (size_t)r, 1); }
  template <class T> const T* ptr() const { return (const T*)data; }
  // the CV_32F pose algebra of SearchBySim3 (a 3 x 3 block in poseR or, is_t, a 3 x 1 block in poset): scaling through a double like
  // cv::Mat::convertTo, products with a double accumulator rounded to float like OpenCV's gemm for CV_32F, sums in float
  bool is_t = false;
  Mat t() const {
    Mat o = *this;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o.poseR[3 * i + j] = poseR[3 * j + i];
    return o;
  }
  friend Mat operator*(double s, const Mat& m) {
    Mat o = m;
    for (int i = 0; i < 9; ++i) o.poseR[i] = (float)(s * (double)m.poseR[i]);
    for (int i = 0; i < 3; ++i) o.poset[i] = (float)(s * (double)m.poset[i]);
    return o;
  }
  friend Mat operator-(const Mat& m) { return -1.0 * m; }
  friend Mat operator*(const Mat& a, const Mat& b) {
    Mat o;
    o.is_t = b.is_t;
    if (b.is_t) {
      for (int i = 0; i < 3; ++i) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += (double)a.poseR[3 * i + k] * (double)b.poset[k];
        o.poset[i] = (float)acc;
      }
    } else {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double acc = 0;
          for (int k = 0; k < 3; ++k) acc += (double)a.poseR[3 * i + k] * (double)b.poseR[3 * k + j];
          o.poseR[3 * i + j] = (float)acc;
        }
    }
    return o;
  }
  friend Mat operator+(const Mat& a, const Mat& b) {
    Mat o = a;
    for (int i = 0; i < 3; ++i) o.poset[i] = a.poset[i] + b.poset[i];
    return o;
  }
};
}  // namespace cvst_sbp
#define cv cvst_sbp

namespace Eigen {
template <class T>
struct Vec3 {
  T v[3];
  Vec3() : v{0, 0, 0} {}
  Vec3(T a, T b, T c) : v{a, b, c} {}
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  template <class U> Vec3<U> cast() const { return Vec3<U>((U)v[0], (U)v[1], (U)v[2]); }
  Vec3& operator+=(const Vec3& o) {
    for (int i = 0; i < 3; ++i) v[i] += o.v[i];
    return *this;
  }
  Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  T dot(const Vec3& o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
  T norm() const { return std::sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])); }  // Eigen's unrolled redux: t0 + (t1 + t2)
};
using Vector3d = Vec3<double>;
using Vector3f = Vec3<float>;
struct Matrix3f {
  float m[9];
  template <class U> Matrix3f cast() const { return *this; }
  Matrix3f transpose() const { return Matrix3f{{m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]}}; }
};
inline Vector3f operator*(const Matrix3f& K, const Vector3f& p) {  // coefficients in the order of Eigen's unrolled redux: t0 + (t1 + t2)
  Vector3f o;
  for (int r = 0; r < 3; ++r) o.v[r] = K.m[3 * r] * p.v[0] + (K.m[3 * r + 1] * p.v[1] + K.m[3 * r + 2] * p.v[2]);
  return o;
}
inline Vector3f operator*(const Matrix3f& R, const Vec3<double>& p) { return R * p.cast<float>(); }  // float rig offsets kept as doubles here
struct Matrix3d {
  double m[9];
};
inline Vector3d operator*(const Matrix3d& R, const Vector3d& p) {
  Vector3d o;
  for (int r = 0; r < 3; ++r) o.v[r] = R.m[3 * r] * p.v[0] + (R.m[3 * r + 1] * p.v[1] + R.m[3 * r + 2] * p.v[2]);
  return o;
}
}  // namespace Eigen
using Eigen::Vector3d;
using Eigen::Vector3f;

namespace Sophus {
struct SE3d {  // unit quaternion (w, x, y, z) + translation
  double q[4] = {1, 0, 0, 0}, t[3] = {0, 0, 0};
  static void rot(const double q[4], const double v[3], double o[3]) {  // Eigen::QuaternionBase::_transformVector
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    double uv[3] = {y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0]};
    for (int i = 0; i < 3; ++i) uv[i] += uv[i];
    const double c[3] = {y * uv[2] - z * uv[1], z * uv[0] - x * uv[2], x * uv[1] - y * uv[0]};
    for (int i = 0; i < 3; ++i) o[i] = v[i] + w * uv[i] + c[i];
  }
  SE3d inverse() const {
    SE3d o;
    o.q[0] = q[0]; o.q[1] = -q[1]; o.q[2] = -q[2]; o.q[3] = -q[3];
    const double nt[3] = {t[0] * -1.0, t[1] * -1.0, t[2] * -1.0};
    rot(o.q, nt, o.t);
    return o;
  }
  SE3d operator*(const SE3d& b) const {
    SE3d o;
    o.q[0] = q[0] * b.q[0] - q[1] * b.q[1] - q[2] * b.q[2] - q[3] * b.q[3];
    o.q[1] = q[0] * b.q[1] + q[1] * b.q[0] + q[2] * b.q[3] - q[3] * b.q[2];
    o.q[2] = q[0] * b.q[2] + q[2] * b.q[0] + q[3] * b.q[1] - q[1] * b.q[3];
    o.q[3] = q[0] * b.q[3] + q[3] * b.q[0] + q[1] * b.q[2] - q[2] * b.q[1];
    double r[3];
    rot(q, b.t, r);
    for (int i = 0; i < 3; ++i) o.t[i] = t[i] + r[i];
    return o;
  }
  Vector3d operator*(const Vector3d& p) const {
    double r[3];
    rot(q, p.v, r);
    return Vector3d(r[0] + t[0], r[1] + t[1], r[2] + t[2]);
  }
  Vector3f operator*(const Vector3f& p) const {  // the rig transform is an SE3f in the reference: same formula in float
    const float w = (float)q[0], x = (float)q[1], y = (float)q[2], z = (float)q[3];
    float uv[3] = {y * p.v[2] - z * p.v[1], z * p.v[0] - x * p.v[2], x * p.v[1] - y * p.v[0]};
    for (int i = 0; i < 3; ++i) uv[i] += uv[i];
    const float c[3] = {y * uv[2] - z * uv[1], z * uv[0] - x * uv[2], x * uv[1] - y * uv[0]};
    return Vector3f(((p.v[0] + w * uv[0]) + c[0]) + (float)t[0], ((p.v[1] + w * uv[1]) + c[1]) + (float)t[1], ((p.v[2] + w * uv[2]) + c[2]) + (float)t[2]);
  }
  Vector3d translation() const { return Vector3d(t[0], t[1], t[2]); }
  Eigen::Matrix3d rotationMatrix() const {  // Eigen::QuaternionBase::toRotationMatrix
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x,
                 tyy = ty * y, tyz = tz * y, tzz = tz * z;
    return Eigen::Matrix3d{{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)}};
  }
  template <class U> SE3d cast() const { return *this; }
};
}  // namespace Sophus

namespace DBoW2 {
typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;  // DBoW2/FeatureVector.h: node id -> feature indices
}

namespace VIEO_SLAM_SBP {
struct Vector2img {
  float v[2];
  float& operator[](int i) { return v[i]; }
};
namespace camm {
struct Camera {
  using Tio = double;
  using Ptr = std::shared_ptr<Camera>;
  Sophus::SE3d Tcr, Trc;
  Eigen::Matrix3f K;
  const Sophus::SE3d& GetTcr() const { return Tcr; }
  const Sophus::SE3d& GetTrc() const { return Trc; }
  Eigen::Matrix3f toK() const { return K; }
  void Project(const Vector3d&, Vector2img*) const {}  // usedistort_ is false in this wrapper
};
}  // namespace camm

class MapPoint {
 public:
  Vector3f pos;
  const uint8_t* desc = nullptr;
  int obs = 0;
  long unsigned int mnId = 0;
  // the tracking info Frame::isInFrustum leaves (include/MapPoint.h _TrackFastMatchInfo): what the local-map search reads
  struct TrackInfo {
    static const int NUM_PROJ = 3;
    bool btrack_inview_ = false;
    list<float> vtrack_proj_[NUM_PROJ];
    list<int> vtrack_scalelevel_;
    list<float> vtrack_viewcos_;
    list<size_t> vtrack_cami_;
    float track_depth_ = 0;
  } trackinfo;
  bool bad = false;
  float mfMaxDistance = 0, mfMinDistance = 0;
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }  // src/MapPoint.cc:481-489 (compiled themselves in ref_frustum_wrap.cc)
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
  int PredictScale(const float& currentDist, class Frame* pF);
  int PredictScale(const float& currentDist, class KeyFrame* pKF);
  using Vector3data = Vector3f;
  Vector3f normal;
  Vector3f GetNormal() { return normal; }
  bool IsInKeyFrame(class KeyFrame*) { return false; }
  class KeyFrame* in_kf = nullptr;  // the keyframe observing this point and the keypoints it is observed at (SearchBySim3)
  set<size_t> in_idx;
  set<size_t> GetIndexInKeyFrame(class KeyFrame* pKF) { return pKF == in_kf ? in_idx : set<size_t>(); }
  Vector3f GetWorldPos() { return pos; }
  cv::Mat GetDescriptor() { return cv::Mat(desc, 1); }
  int Observations() { return obs; }
  bool isBad() { return bad; }
  TrackInfo& GetTrackInfoRef() { return trackinfo; }
};

class FrameBase {  // as in ref_grid_wrap.cc
 public:
  typedef struct _GridInfo {
    vector<float> fgrids_widthinv_;
    vector<float> fgrids_heightinv_;
    const int FRAME_GRID_ROWS = 48;
    const int FRAME_GRID_COLS = 64;
    vector<array<float, 4>> minmax_xy_;
  } GridInfo;
  GridInfo gridinfo_;
  vector<vector<vector<size_t>>> vgrids_;
  int N = 0;
  static bool usedistort_;
  vector<camm::Camera::Ptr> mpCameras;
  vector<cv::KeyPoint> mvKeys, mvKeysUn;
  vector<pair<size_t, size_t>> mapn2in_;
  vector<size_t> GetFeaturesInArea(uint8_t cami, const float& x, const float& y, const float& r, const int minlevel = -1,
                                   const int maxlevel = -1) const;
  void AssignFeaturesToGrid();
  bool PosInGrid(uint8_t cami, const cv::KeyPoint& kp, int& posX, int& posY);
  bool IsInImage(uint8_t cami, const float& x, const float& y) const;
};
bool FrameBase::usedistort_ = false;
#include "grid_fns.inc"

class Frame : public FrameBase {
 public:
  Sophus::SE3d Tcw;
  const Sophus::SE3d GetTcwCst() const { return Tcw; }
  struct {
    float baseline_bf_[2];
    vector<float> vuright_;
  } stereoinfo_;
  struct {
    vector<float> vscalefactor_;
    float flogscalefactor_ = 0;
  } scalepyrinfo_;
  vector<MapPoint*> mvpMapPoints;
  vector<bool> mvbOutlier;
  cv::Mat mDescriptors;
  DBoW2::FeatureVector mFeatVec;
  const vector<MapPoint*>& GetMapPointMatches() const { return mvpMapPoints; }
  void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
  void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = nullptr; }
};

extern "C" int ref_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);  // ref_mappoint_wrap.cc
}  // namespace VIEO_SLAM_SBP
namespace VIEO_SLAM_SBP {
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
  return ref_predict_scale(mfMaxDistance, currentDist, pF->scalepyrinfo_.flogscalefactor_, (int)pF->scalepyrinfo_.vscalefactor_.size());
}

struct Mat3Holder {
  Eigen::Matrix3f m;
  template <class U> Eigen::Matrix3f cast() const { return m; }
};
struct Converter {
  static Mat3Holder toMatrix3d(const cv::Mat& p) {
    Mat3Holder h;
    for (int i = 0; i < 9; ++i) h.m.m[i] = p.poseR[i];
    return h;
  }
  static Vector3d toVector3d(const cv::Mat& p) { return Vector3d(p.poset[0], p.poset[1], p.poset[2]); }
};
class KeyFrame : public FrameBase {  // the members SearchByBoW(KeyFrame*, Frame&, ...), the relocalisation search and SearchByProjectionBase read
 public:
  vector<MapPoint*> mvpMapPoints;
  vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  std::vector<std::pair<size_t, MapPoint*>> fused;  // what Fuse hands to the map bookkeeping, recorded
  void FuseMP(size_t idx, MapPoint* pMP) { fused.emplace_back(idx, pMP); }
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mDescriptors;
  cv::Mat Ow, Rcw_m, tcw_m;
  cv::Mat GetCameraCenter() { return Ow; }
  cv::Mat GetRotation() { return Rcw_m; }
  cv::Mat GetTranslation() { return tcw_m; }
  struct {
    vector<float> vuright_;
    float baseline_bf_[2] = {0, 0};
  } stereoinfo_;
  struct {
    vector<float> vscalefactor_, vinvlevelsigma2_;
    float flogscalefactor_ = 0;
  } scalepyrinfo_;
};
int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
  return ref_predict_scale(mfMaxDistance, currentDist, pKF->scalepyrinfo_.flogscalefactor_, (int)pKF->scalepyrinfo_.vscalefactor_.size());
}

class ORBmatcher {
 public:
  enum ModeSBP { SBPFuseLater = 0x1, SBPMatchMultiCam = 0x2 };  // include/ORBmatcher.h:27
  void SearchByProjectionBase(const vector<MapPoint*>& vpMapPoints1, cv::Mat Rcrw_cv, cv::Mat tcrw_cv, KeyFrame* pKF, const float th_radius,
                              const float th_bestdist, bool bCheckViewingAngle = false, const float* pbf = nullptr, int* pnfused = nullptr,
                              char mode = (char)SBPMatchMultiCam, vector<vector<bool>>* pvbAlreadyMatched1 = nullptr,
                              vector<set<int>>* pvnMatch1 = nullptr);
  int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0);
  int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12,
                   const float th);
  int SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches);
  int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist,
                         const float th_far_pts = 0);
  ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, const float th_far_pts = 0);
  int SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th = 3, const float th_far_pts = 0);
  float RadiusByViewingCos(const float& viewCos);
  float mfNNratio;
  bool mbCheckOrientation;
};
const int ORBmatcher::TH_HIGH = 100;  // src/ORBmatcher.cc:20-22
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
#include "orbmatcher_fns.inc"
#include "sbp_fns.inc"
#include "bow_fns.inc"
#include "reloc_fns.inc"
#include "sbpbase_fns.inc"
#include "fuse_fns.inc"
#include "sim3search_fns.inc"
}  // namespace VIEO_SLAM_SBP
#undef cv

struct RefKp {
  float x, y, size, angle, response;
  int32_t octave;
};
struct RefSbpFrame {  // == OrcSbpFrame
  int32_t kp_begin, n_kp, q_begin, n_q;
  float minx, maxx, miny, maxy, grid_winv, grid_hinv, bf, b, fx, fy, cx, cy, th, th_far, nn_ratio;
  int32_t mono, check_orientation, n_levels;
  float scale[16];
  double qcw[4], tcw[3], qlw[4], tlw[3];
};
static void fill_frame(VIEO_SLAM_SBP::Frame& F, const RefSbpFrame* f, const RefKp* kps, const float* uright, const uint8_t* desc) {
  using namespace VIEO_SLAM_SBP;
  auto cam = std::make_shared<camm::Camera>();
  const float K[9] = {f->fx, 0.f, f->cx, 0.f, f->fy, f->cy, 0.f, 0.f, 1.f};
  for (int i = 0; i < 9; ++i) cam->K.m[i] = K[i];
  F.mpCameras.push_back(cam);
  F.gridinfo_.fgrids_widthinv_ = {f->grid_winv};
  F.gridinfo_.fgrids_heightinv_ = {f->grid_hinv};
  F.gridinfo_.minmax_xy_.push_back({f->minx, f->maxx, f->miny, f->maxy});
  F.N = f->n_kp;
  F.mvKeysUn.resize(f->n_kp);
  for (int i = 0; i < f->n_kp; ++i) {
    F.mvKeysUn[i].pt.x = kps[i].x; F.mvKeysUn[i].pt.y = kps[i].y; F.mvKeysUn[i].octave = kps[i].octave; F.mvKeysUn[i].angle = kps[i].angle;
  }
  F.mvKeys = F.mvKeysUn;
  F.AssignFeaturesToGrid();
  F.stereoinfo_.baseline_bf_[0] = f->b;
  F.stereoinfo_.baseline_bf_[1] = f->bf;
  F.stereoinfo_.vuright_.assign(uright, uright + f->n_kp);
  F.scalepyrinfo_.vscalefactor_.assign(f->scale, f->scale + f->n_levels);
  F.mvpMapPoints.assign(f->n_kp, nullptr);
  F.mDescriptors = cvst_sbp::Mat(desc, f->n_kp);
  for (int k = 0; k < 4; ++k) F.Tcw.q[k] = f->qcw[k];
  for (int k = 0; k < 3; ++k) F.Tcw.t[k] = f->tcw[k];
}

// same arguments as orc_sbp_last_frame; kp_match[k] = the query that ended up in CurrentFrame.mvpMapPoints[k] (-1 none)
extern "C" int ref_sbp_last_frame(const RefSbpFrame* f, const RefKp* kps, const float* uright, const uint8_t* desc, const double* q_Xw,
                                  const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc, const uint8_t* q_flags,
                                  const uint8_t* kp_blocked, int32_t* kp_match) {
  using namespace VIEO_SLAM_SBP;
  Frame cur, last;
  fill_frame(cur, f, kps, uright, desc);
  MapPoint blocker;
  blocker.obs = 1;
  if (kp_blocked)
    for (int k = 0; k < f->n_kp; ++k)
      if (kp_blocked[k]) cur.mvpMapPoints[k] = &blocker;
  std::vector<MapPoint> mps(f->n_q);
  last.N = f->n_q;
  last.mvKeys.resize(f->n_q);
  last.mvpMapPoints.resize(f->n_q);
  last.mvbOutlier.assign(f->n_q, false);
  for (int i = 0; i < f->n_q; ++i) {
    // GetWorldPos() is float in the reference (MapPoint::Vector3data) and cast back to double: the caller passes float-valued doubles
    mps[i].pos = Vector3f((float)q_Xw[3 * i], (float)q_Xw[3 * i + 1], (float)q_Xw[3 * i + 2]);
    mps[i].desc = q_desc + 32 * (size_t)i;
    mps[i].obs = (q_flags[i] & 1) ? 1 : 0;
    mps[i].mnId = i;
    last.mvpMapPoints[i] = &mps[i];
    last.mvKeys[i].octave = q_octave[i];
    last.mvKeys[i].angle = q_angle[i];
  }
  for (int k = 0; k < 4; ++k) last.Tcw.q[k] = f->qlw[k];
  for (int k = 0; k < 3; ++k) last.Tcw.t[k] = f->tlw[k];
  ORBmatcher m(f->nn_ratio, f->check_orientation != 0);
  const int n = m.SearchByProjection(cur, last, f->th, f->mono != 0, f->th_far);
  for (int k = 0; k < f->n_kp; ++k) {
    MapPoint* p = cur.mvpMapPoints[k];
    kp_match[k] = (p && p != &blocker) ? (int32_t)(p - mps.data()) : -1;
  }
  return n;
}

// same arguments as orc_sbp_local_map: queries = the local map points in view with the tracking info Frame::isInFrustum left
extern "C" int ref_sbp_local_map(const RefSbpFrame* f, const RefKp* kps, const float* uright, const uint8_t* desc, const float* q_proj,
                                 const int32_t* q_level, const float* q_viewcos, const float* q_depth, const uint8_t* q_desc,
                                 const uint8_t* q_flags, const uint8_t* kp_blocked, int32_t* kp_match) {
  using namespace VIEO_SLAM_SBP;
  Frame cur;
  fill_frame(cur, f, kps, uright, desc);
  MapPoint blocker;
  blocker.obs = 1;
  if (kp_blocked)
    for (int k = 0; k < f->n_kp; ++k)
      if (kp_blocked[k]) cur.mvpMapPoints[k] = &blocker;
  std::vector<MapPoint> mps(f->n_q);
  std::vector<MapPoint*> vp(f->n_q);
  for (int i = 0; i < f->n_q; ++i) {
    MapPoint& m = mps[i];
    m.desc = q_desc + 32 * (size_t)i;
    m.obs = (q_flags[i] & 1) ? 1 : 0;
    m.mnId = i;
    m.trackinfo.btrack_inview_ = true;
    for (int k = 0; k < 3; ++k) m.trackinfo.vtrack_proj_[k].push_back(q_proj[3 * i + k]);
    m.trackinfo.vtrack_scalelevel_.push_back(q_level[i]);
    m.trackinfo.vtrack_viewcos_.push_back(q_viewcos[i]);
    m.trackinfo.vtrack_cami_.push_back(0);
    m.trackinfo.track_depth_ = q_depth[i];
    vp[i] = &m;
  }
  ORBmatcher m(f->nn_ratio, f->check_orientation != 0);
  const int n = m.SearchByProjection(cur, vp, f->th, f->th_far);
  for (int k = 0; k < f->n_kp; ++k) {
    MapPoint* p = cur.mvpMapPoints[k];
    kp_match[k] = (p && p != &blocker) ? (int32_t)(p - mps.data()) : -1;
  }
  return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:344-505), same arguments as orc_search_by_bow:
// match_f[k] = id of the keyframe map point assigned to frame keypoint k, or -1
extern "C" int ref_search_by_bow(const RefKp* kp_kf, const uint8_t* desc_kf, const int32_t* mp_id, const int32_t* fv1_node,
                                 const int32_t* fv1_ptr, const int32_t* fv1_idx, int n_nodes1, const RefKp* kp_f, const uint8_t* desc_f,
                                 int n_f, const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int n_nodes2,
                                 float nn_ratio, int check_orientation, int32_t* match_f) {
  using namespace VIEO_SLAM_SBP;
  int n_kf = 0, max_id = -1;
  for (int k = 0; k < fv1_ptr[n_nodes1]; ++k) n_kf = std::max(n_kf, fv1_idx[k] + 1);
  for (int k = 0; k < n_kf; ++k) max_id = std::max(max_id, mp_id[k]);
  std::vector<MapPoint> mps(max_id + 1);
  for (int k = 0; k <= max_id; ++k) mps[k].mnId = k;
  KeyFrame kf;
  kf.mvpMapPoints.assign(n_kf, nullptr);
  kf.mvKeys.resize(n_kf);
  for (int k = 0; k < n_kf; ++k) {
    if (mp_id[k] >= 0) kf.mvpMapPoints[k] = &mps[mp_id[k]];
    kf.mvKeys[k].angle = kp_kf[k].angle;
  }
  kf.mDescriptors = cvst_sbp::Mat(desc_kf, n_kf);
  for (int a = 0; a < n_nodes1; ++a)
    kf.mFeatVec[(unsigned)fv1_node[a]].assign(fv1_idx + fv1_ptr[a], fv1_idx + fv1_ptr[a + 1]);
  Frame F;
  F.N = n_f;
  F.mvKeys.resize(n_f);
  for (int k = 0; k < n_f; ++k) F.mvKeys[k].angle = kp_f[k].angle;
  F.mDescriptors = cvst_sbp::Mat(desc_f, n_f);
  for (int a = 0; a < n_nodes2; ++a) F.mFeatVec[(unsigned)fv2_node[a]].assign(fv2_idx + fv2_ptr[a], fv2_idx + fv2_ptr[a + 1]);
  ORBmatcher m(nn_ratio, check_orientation != 0);
  std::vector<MapPoint*> out;
  const int n = m.SearchByBoW(&kf, F, out);
  for (int k = 0; k < n_f; ++k) match_f[k] = out[k] ? (int32_t)out[k]->mnId : -1;
  return n;
}

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist, th_far_pts) (src/ORBmatcher.cc:1471-1606),
// same inputs as orc_sbp_reloc: the queries are the keyframe's map points (already-found ones are simply not passed)
extern "C" int ref_sbp_reloc(const RefSbpFrame* f, int orb_dist, float log_scale_factor, const RefKp* kps, const uint8_t* desc,
                             const double* q_Xw, const float* q_angle, const float* q_max_dist, const float* q_min_dist,
                             const uint8_t* q_desc, const uint8_t* kp_blocked, int32_t* kp_match) {
  using namespace VIEO_SLAM_SBP;
  Frame cur;
  std::vector<float> no_right(f->n_kp, -1.0f);
  fill_frame(cur, f, kps, no_right.data(), desc);
  cur.scalepyrinfo_.flogscalefactor_ = log_scale_factor;
  MapPoint blocker;
  if (kp_blocked)
    for (int k = 0; k < f->n_kp; ++k)
      if (kp_blocked[k]) cur.mvpMapPoints[k] = &blocker;
  std::vector<MapPoint> mps(f->n_q);
  KeyFrame kf;
  kf.mvKeys.resize(f->n_q);
  for (int i = 0; i < f->n_q; ++i) {
    mps[i].pos = Vector3f((float)q_Xw[3 * i], (float)q_Xw[3 * i + 1], (float)q_Xw[3 * i + 2]);
    mps[i].desc = q_desc + 32 * (size_t)i;
    mps[i].mnId = i;
    mps[i].mfMaxDistance = q_max_dist[i], mps[i].mfMinDistance = q_min_dist[i];
    kf.mvpMapPoints.push_back(&mps[i]);
    kf.mvKeys[i].angle = q_angle[i];
  }
  ORBmatcher m(f->nn_ratio, f->check_orientation != 0);
  const int n = m.SearchByProjection(cur, &kf, std::set<MapPoint*>(), f->th, orb_dist, f->th_far);
  for (int k = 0; k < f->n_kp; ++k) {
    MapPoint* p = cur.mvpMapPoints[k];
    kp_match[k] = (p && p != &blocker) ? (int32_t)(p - mps.data()) : -1;
  }
  return n;
}

// ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227), the search half behind Fuse / SearchBySim3: same inputs as orc_sbp_base.
// Run in the MatchMultiCam | FuseLater mode with th_bestdist = 256 so that pvnMatch1 reports the arg-min keypoint of every map point
// that found one; best_dist is the Hamming distance to it.
struct RefProjSearchFrame {  // == OrcProjSearchFrame
  int32_t kp_begin, n_kp, q_begin, n_q;
  float Rcw[9], tcw[3], Ow[3];
  float fx, fy, cx, cy, minx, maxx, miny, maxy, grid_winv, grid_hinv, bf;
  int32_t use_bf, check_viewing_angle;
  float th_radius;
  int32_t n_levels;
  float log_scale_factor, scale[16], inv_level_sigma2[16], level_ratio[16];
};
extern "C" void ref_sbp_base(const RefProjSearchFrame* f, const RefKp* kps, const float* uright, const uint8_t* desc, const float* wP,
                             const float* Pn, const float* max_dist, const float* min_dist, const uint8_t* q_desc, const uint8_t* q_skip,
                             int32_t* best_idx, int32_t* best_dist) {
  using namespace VIEO_SLAM_SBP;
  KeyFrame kf;
  auto cam = std::make_shared<camm::Camera>();
  const float K[9] = {f->fx, 0.f, f->cx, 0.f, f->fy, f->cy, 0.f, 0.f, 1.f};
  for (int i = 0; i < 9; ++i) cam->K.m[i] = K[i];
  kf.mpCameras.push_back(cam);
  kf.gridinfo_.fgrids_widthinv_ = {f->grid_winv};
  kf.gridinfo_.fgrids_heightinv_ = {f->grid_hinv};
  kf.gridinfo_.minmax_xy_.push_back({f->minx, f->maxx, f->miny, f->maxy});
  kf.N = f->n_kp;
  kf.mvKeysUn.resize(f->n_kp);
  for (int i = 0; i < f->n_kp; ++i) kf.mvKeysUn[i].pt.x = kps[i].x, kf.mvKeysUn[i].pt.y = kps[i].y, kf.mvKeysUn[i].octave = kps[i].octave;
  kf.mvKeys = kf.mvKeysUn;
  kf.AssignFeaturesToGrid();
  kf.stereoinfo_.vuright_.assign(uright, uright + f->n_kp);
  kf.scalepyrinfo_.vscalefactor_.assign(f->scale, f->scale + f->n_levels);
  kf.scalepyrinfo_.vinvlevelsigma2_.assign(f->inv_level_sigma2, f->inv_level_sigma2 + f->n_levels);
  kf.scalepyrinfo_.flogscalefactor_ = f->log_scale_factor;
  kf.mvpMapPoints.assign(f->n_kp, nullptr);
  kf.mDescriptors = cvst_sbp::Mat(desc, f->n_kp);
  cvst_sbp::Mat Rm, tm;
  for (int i = 0; i < 9; ++i) Rm.poseR[i] = f->Rcw[i];
  for (int i = 0; i < 3; ++i) tm.poset[i] = f->tcw[i], kf.Ow.poset[i] = f->Ow[i];
  std::vector<MapPoint> mps(f->n_q);
  std::vector<MapPoint*> vp(f->n_q, nullptr);
  for (int i = 0; i < f->n_q; ++i) {
    best_idx[i] = -1, best_dist[i] = 256;
    if (q_skip && q_skip[i]) continue;
    MapPoint& m = mps[i];
    m.pos = Vector3f(wP[3 * i], wP[3 * i + 1], wP[3 * i + 2]);
    m.normal = Vector3f(Pn[3 * i], Pn[3 * i + 1], Pn[3 * i + 2]);
    m.mfMaxDistance = max_dist[i], m.mfMinDistance = min_dist[i];
    m.desc = q_desc + 32 * (size_t)i;
    vp[i] = &m;
  }
  ORBmatcher matcher(0.6f, true);
  std::vector<std::set<int>> found;
  const float bf = f->bf;
  matcher.SearchByProjectionBase(vp, Rm, tm, &kf, f->th_radius, 256.0f, f->check_viewing_angle != 0, f->use_bf ? &bf : nullptr, nullptr,
                                 (char)(ORBmatcher::SBPMatchMultiCam | ORBmatcher::SBPFuseLater), nullptr, &found);
  for (int i = 0; i < f->n_q; ++i)
    if (!found[i].empty() && *found[i].begin() >= 0) {
      best_idx[i] = *found[i].begin();
      best_dist[i] = ORBmatcher::DescriptorDistance(cvst_sbp::Mat(q_desc + 32 * (size_t)i, 1), cvst_sbp::Mat(desc + 32 * (size_t)best_idx[i], 1));
    }
}

// ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:1152-1165) compiled unchanged over the compiled
// SearchByProjectionBase: same inputs as ref_sbp_base (the viewing-cone test and the stereo gate are what Fuse itself passes: on, with
// the keyframe's bf).  fused_idx [n_q] = the keypoint KeyFrame::FuseMP was called with for each map point (-1: none); returns nFused.
extern "C" int ref_fuse(const RefProjSearchFrame* f, const RefKp* kps, const float* uright, const uint8_t* desc, const float* wP,
                        const float* Pn, const float* max_dist, const float* min_dist, const uint8_t* q_desc, const uint8_t* q_skip,
                        int32_t* fused_idx) {
  using namespace VIEO_SLAM_SBP;
  KeyFrame kf;
  auto cam = std::make_shared<camm::Camera>();
  const float K[9] = {f->fx, 0.f, f->cx, 0.f, f->fy, f->cy, 0.f, 0.f, 1.f};
  for (int i = 0; i < 9; ++i) cam->K.m[i] = K[i];
  kf.mpCameras.push_back(cam);
  kf.gridinfo_.fgrids_widthinv_ = {f->grid_winv};
  kf.gridinfo_.fgrids_heightinv_ = {f->grid_hinv};
  kf.gridinfo_.minmax_xy_.push_back({f->minx, f->maxx, f->miny, f->maxy});
  kf.N = f->n_kp;
  kf.mvKeysUn.resize(f->n_kp);
  for (int i = 0; i < f->n_kp; ++i) kf.mvKeysUn[i].pt.x = kps[i].x, kf.mvKeysUn[i].pt.y = kps[i].y, kf.mvKeysUn[i].octave = kps[i].octave;
  kf.mvKeys = kf.mvKeysUn;
  kf.AssignFeaturesToGrid();
  kf.stereoinfo_.vuright_.assign(uright, uright + f->n_kp);
  kf.stereoinfo_.baseline_bf_[1] = f->bf;
  kf.scalepyrinfo_.vscalefactor_.assign(f->scale, f->scale + f->n_levels);
  kf.scalepyrinfo_.vinvlevelsigma2_.assign(f->inv_level_sigma2, f->inv_level_sigma2 + f->n_levels);
  kf.scalepyrinfo_.flogscalefactor_ = f->log_scale_factor;
  kf.mvpMapPoints.assign(f->n_kp, nullptr);
  kf.mDescriptors = cvst_sbp::Mat(desc, f->n_kp);
  for (int i = 0; i < 9; ++i) kf.Rcw_m.poseR[i] = f->Rcw[i];
  for (int i = 0; i < 3; ++i) kf.tcw_m.poset[i] = f->tcw[i], kf.Ow.poset[i] = f->Ow[i];
  std::vector<MapPoint> mps(f->n_q);
  std::vector<MapPoint*> vp(f->n_q, nullptr);
  for (int i = 0; i < f->n_q; ++i) {
    fused_idx[i] = -1;
    if (q_skip && q_skip[i]) continue;
    MapPoint& m = mps[i];
    m.pos = Vector3f(wP[3 * i], wP[3 * i + 1], wP[3 * i + 2]);
    m.normal = Vector3f(Pn[3 * i], Pn[3 * i + 1], Pn[3 * i + 2]);
    m.mfMaxDistance = max_dist[i], m.mfMinDistance = min_dist[i];
    m.desc = q_desc + 32 * (size_t)i;
    vp[i] = &m;
  }
  ORBmatcher matcher(0.6f, true);
  const int n = matcher.Fuse(&kf, vp, f->th_radius);
  for (auto& pr : kf.fused) fused_idx[pr.second - mps.data()] = (int32_t)pr.first;
  return n;
}

// ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1222-1302, LoopClosing::ComputeSim3) compiled unchanged over the compiled
// SearchByProjectionBase: two keyframes (f1 / f2: camera, grid, keypoints; their own map point per keypoint in the q arrays, q_skip =
// the keypoint holds none; Rcw / tcw / Ow = the keyframe's pose), the similarity (s12, R12 row-major, t12), the search radius th and
// prior12 [n_kp1] (-1 or the keyframe-2 keypoint already matched).  match12 [n_kp1] = keyframe-2 keypoint whose map point
// vpMatches12[i1] holds afterwards (-1 none); returns nFound.  pose21 / pose12 [15] = (R | t | unused) the two searches ran with
// (sR21 R1w, sR21 t1w + t21) and (sR12 R2w, sR12 t2w + t12): what the caller of the flat interface hands over as the frames' poses.
namespace {
struct Sim3Side {
  VIEO_SLAM_SBP::KeyFrame kf;
  std::vector<VIEO_SLAM_SBP::MapPoint> mps;
};
void fill_sim3_side(Sim3Side& S, const RefProjSearchFrame* f, const RefKp* kps, const float* uright, const uint8_t* desc, const float* wP,
                    const float* Pn, const float* max_dist, const float* min_dist, const uint8_t* q_desc, const uint8_t* q_skip) {
  using namespace VIEO_SLAM_SBP;
  KeyFrame& kf = S.kf;
  auto cam = std::make_shared<camm::Camera>();
  const float K[9] = {f->fx, 0.f, f->cx, 0.f, f->fy, f->cy, 0.f, 0.f, 1.f};
  for (int i = 0; i < 9; ++i) cam->K.m[i] = K[i];
  kf.mpCameras.push_back(cam);
  kf.gridinfo_.fgrids_widthinv_ = {f->grid_winv};
  kf.gridinfo_.fgrids_heightinv_ = {f->grid_hinv};
  kf.gridinfo_.minmax_xy_.push_back({f->minx, f->maxx, f->miny, f->maxy});
  kf.N = f->n_kp;
  kf.mvKeysUn.resize(f->n_kp);
  for (int i = 0; i < f->n_kp; ++i) kf.mvKeysUn[i].pt.x = kps[i].x, kf.mvKeysUn[i].pt.y = kps[i].y, kf.mvKeysUn[i].octave = kps[i].octave;
  kf.mvKeys = kf.mvKeysUn;
  kf.AssignFeaturesToGrid();
  kf.stereoinfo_.vuright_.assign(uright, uright + f->n_kp);
  kf.stereoinfo_.baseline_bf_[1] = f->bf;
  kf.scalepyrinfo_.vscalefactor_.assign(f->scale, f->scale + f->n_levels);
  kf.scalepyrinfo_.vinvlevelsigma2_.assign(f->inv_level_sigma2, f->inv_level_sigma2 + f->n_levels);
  kf.scalepyrinfo_.flogscalefactor_ = f->log_scale_factor;
  kf.mDescriptors = cvst_sbp::Mat(desc, f->n_kp);
  for (int i = 0; i < 9; ++i) kf.Rcw_m.poseR[i] = f->Rcw[i];
  kf.tcw_m.is_t = true;
  kf.Ow.is_t = true;
  for (int i = 0; i < 3; ++i) kf.tcw_m.poset[i] = f->tcw[i], kf.Ow.poset[i] = f->Ow[i];
  S.mps.resize(f->n_kp);
  kf.mvpMapPoints.assign(f->n_kp, nullptr);
  for (int i = 0; i < f->n_kp; ++i) {
    if (q_skip[i]) continue;
    MapPoint& m = S.mps[i];
    m.pos = Vector3f(wP[3 * i], wP[3 * i + 1], wP[3 * i + 2]);
    m.normal = Vector3f(Pn[3 * i], Pn[3 * i + 1], Pn[3 * i + 2]);
    m.mfMaxDistance = max_dist[i], m.mfMinDistance = min_dist[i];
    m.desc = q_desc + 32 * (size_t)i;
    m.in_kf = &kf;
    m.in_idx = {(size_t)i};
    kf.mvpMapPoints[i] = &m;
  }
}
}  // namespace
extern "C" int ref_search_by_sim3(const RefProjSearchFrame* f1, const RefKp* kps1, const float* ur1, const uint8_t* desc1, const float* wP1,
                                  const float* Pn1, const float* maxd1, const float* mind1, const uint8_t* qdesc1, const uint8_t* skip1,
                                  const RefProjSearchFrame* f2, const RefKp* kps2, const float* ur2, const uint8_t* desc2, const float* wP2,
                                  const float* Pn2, const float* maxd2, const float* mind2, const uint8_t* qdesc2, const uint8_t* skip2,
                                  float s12, const float* R12, const float* t12, float th, const int32_t* prior12, int32_t* match12,
                                  float* pose21, float* pose12) {
  using namespace VIEO_SLAM_SBP;
  Sim3Side A, B;
  fill_sim3_side(A, f1, kps1, ur1, desc1, wP1, Pn1, maxd1, mind1, qdesc1, skip1);
  fill_sim3_side(B, f2, kps2, ur2, desc2, wP2, Pn2, maxd2, mind2, qdesc2, skip2);
  cvst_sbp::Mat Rm, tm;
  tm.is_t = true;
  for (int i = 0; i < 9; ++i) Rm.poseR[i] = R12[i];
  for (int i = 0; i < 3; ++i) tm.poset[i] = t12[i];
  std::vector<MapPoint*> m12(f1->n_kp, nullptr);
  for (int i = 0; i < f1->n_kp; ++i)
    if (prior12 && prior12[i] >= 0) m12[i] = B.kf.mvpMapPoints[prior12[i]];
  ORBmatcher matcher(0.75f, true);
  const int n = matcher.SearchBySim3(&A.kf, &B.kf, m12, s12, Rm, tm, th);
  for (int i = 0; i < f1->n_kp; ++i) match12[i] = m12[i] ? (int32_t)(m12[i] - B.mps.data()) : -1;
  {  // the same expressions as :1233-1238, 1264, 1274 for the flat interface's frame poses
    const cvst_sbp::Mat sR12 = s12 * Rm, sR21 = (1.0 / s12) * Rm.t(), t21 = -sR21 * tm;
    const cvst_sbp::Mat Ra = sR21 * A.kf.Rcw_m, ta = sR21 * A.kf.tcw_m + t21, Rb = sR12 * B.kf.Rcw_m, tb = sR12 * B.kf.tcw_m + tm;
    for (int i = 0; i < 9; ++i) pose21[i] = Ra.poseR[i], pose12[i] = Rb.poseR[i];
    for (int i = 0; i < 3; ++i) pose21[9 + i] = ta.poset[i], pose12[9 + i] = tb.poset[i];
  }
  return n;
}

// ORACLE — TEST INFRASTRUCTURE ONLY.
// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:896-1150) of the REFERENCE compiled UNCHANGED, together with the two
// GeometricCamera members it calls — epipolarConstrain (common/camera_models/camera_base.h:287-406, the fundamental-matrix branch the
// build selects) and FillMatchesFromPair (:408-585, USE_STRATEGY_MIN_DIST as common/config.h:10-13 sets it) — and the reference's own
// ORBmatcher::DescriptorDistance / ComputeThreeMaxima, common/so3_extra.h (hat) and common/unordered_hash.h (PairHash), all taken from
// where they lie (function bodies cut out by name at build time into oracle/_ref/gen/sft_*.inc).
// What is pinned: the FeatureVector walk with its lower_bound jumps, the map-point / only-stereo / injection ("multi2one") skips, the
// per-image best distance, the epipole gate for monocular pairs, the float / double mix of the epipolar test, the bookkeeping of
// vidxs_matches / goodmatches / mapcamidx2idxs / lastdists, the rotation histogram and the final pair list in creation order.
// What this file supplies (stand-ins, stated for what they are): the members of KeyFrame / camera classes the bodies touch, a TU-local
// cv::Mat / KeyPoint (`#define cv cvst_sft`), a Sophus::SE3 over the SO3 stand-in (product / inverse / cast as Sophus defines them), and
// the Eigen stand-in of eigstub/ (3 x 3 inverse by cofactors like Eigen's fixed-size path).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "mini_eigen.h"
#include "sophus/se3.hpp"
#include "common/so3_extra.h"       // the reference's, unchanged
#include "common/unordered_hash.h"  // the reference's, unchanged

using namespace std;

#define PRINT_DEBUG_FILE_MUTEX(...)
#define PRINT_DEBUG_FILE(...)
#define CV_Assert(x) assert(x)
#define USE_STRATEGY_MIN_DIST  // common/config.h:10-13: defined unless USE_STRATEGY_ABANDON is
typedef float FLT_CAMM;        // common/config.h:23
typedef double FLT_CALC_CAMM;  // common/config.h:24

namespace Eigen {
template <class T>
using aligned_vector = std::vector<T>;  // common/eigen_utils.h: std::vector with Eigen's aligned allocator
}

namespace cvst_sft {
struct Point2f {
  float x, y;
};
struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
// descriptor rows, or a CV_32F pose block (3 x 3 / 3 x 1): product with a double accumulator rounded to float (OpenCV's gemm for
// CV_32F accumulates in double), sum in float
class Mat {
 public:
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0;
  float f[9] = {0};
  Mat() {}
  Mat(const uint8_t* d, int r) : data(d), rows(r), cols(32) {}
  Mat row(int r) const { return Mat(data + 32 * (size_t)r, 1); }
  template <class T>
  const T* ptr() const {
    return (const T*)data;
  }
  Mat operator*(const Mat& b) const {
    Mat o;
    o.rows = rows;
    o.cols = b.cols;
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < b.cols; ++j) {
        double s = 0;
        for (int k = 0; k < cols; ++k) s += (double)f[i * cols + k] * (double)b.f[k * b.cols + j];
        o.f[i * b.cols + j] = (float)s;
      }
    return o;
  }
  Mat operator+(const Mat& b) const {
    Mat o = *this;
    for (int i = 0; i < rows * cols; ++i) o.f[i] = f[i] + b.f[i];
    return o;
  }
};
}  // namespace cvst_sft
#define cv cvst_sft

namespace DBoW2 {
typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;  // DBoW2/FeatureVector.h: node id -> feature indices
}

namespace VIEO_SLAM_SFT {
using Eigen::Vector2f;
using Eigen::Vector3f;
using Vector2img = Eigen::Matrix<FLT_CAMM, 2, 1>;  // src/Odom/g2otypes.h:31

namespace camm {
// the declarations of common/camera_models/camera_base.h the two compiled bodies see (type aliases as there)
class GeometricCamera {
 protected:
  using Tcalc = FLT_CALC_CAMM;

 public:
  using Tdata = FLT_CAMM;
  using Tio = double;
  using Ptr = std::shared_ptr<GeometricCamera>;
  using size_t = std::size_t;
  template <typename _Tp>
  using vector = std::vector<_Tp>;
  template <typename _Tp>
  using aligned_vector = Eigen::aligned_vector<_Tp>;
  template <typename _T1, typename _T2>
  using pair = std::pair<_T1, _T2>;
  using MapCamIdx2Idx = std::unordered_map<pair<size_t, size_t>, size_t, PairHash>;
  using Vec2data = Eigen::Matrix<Tdata, 2, 1>;
  using Vec2calc = Eigen::Matrix<Tcalc, 2, 1>;
  using Vec3calc = Eigen::Matrix<Tcalc, 3, 1>;
  using Mat3data = Eigen::Matrix<Tdata, 3, 3>;
  using Mat3calc = Eigen::Matrix<Tcalc, 3, 3>;
  using Vec3io = Eigen::Matrix<Tio, 3, 1>;
  using Mat3io = Eigen::Matrix3d;
  using SE3data = Sophus::SE3<Tdata>;
  using SE3io = Sophus::SE3<Tio>;

  virtual ~GeometricCamera() {}
  float fx = 0, fy = 0, cx = 0, cy = 0;
  SE3data Trc_, Tcr_;
  const SE3data& GetTrc() const { return Trc_; }
  const SE3data& GetTcr() const { return Tcr_; }
  virtual Mat3data toK() const = 0;
  virtual void Project(const Vec3io&, Vec2data*) const { abort(); }    // usedistort_ is false in this wrapper
  virtual void UnProject(const Vec2data&, Vec3io*) const { abort(); }  // bkp_distort is false in this wrapper
  virtual vector<Tdata> TriangulateMatches(const vector<const GeometricCamera*>&, const aligned_vector<Vec2data>&, const vector<float>&,
                                           Vec3io* = nullptr, float = 0.9998f) const {
    abort();  // reached only with psigmas / pkpts, which SearchForTriangulation passes as nullptr
  }
  virtual bool epipolarConstrain(GeometricCamera* otherCamera, const Vec2data& kp1, const Vec2data& kp2, const Mat3io& R12,
                                 const Vec3io& t12, const float sigmaLevel, const float unc, bool bkp_distort = true) const;
  virtual bool FillMatchesFromPair(const vector<const GeometricCamera*>& pcams, size_t n_cams_tot,
                                   const vector<pair<size_t, size_t>>& vcamidx, float dist, vector<vector<size_t>>& vidxsmatches,
                                   vector<bool>& goodmatches_, MapCamIdx2Idx& mapcamidx2idxs, const float thresh_cosdisparity = 1. - 1.e-6,
                                   aligned_vector<Vec3io>* pv3dpoints = nullptr, aligned_vector<Vec2data>* pkpts = nullptr,
                                   vector<float>* psigmas = nullptr, vector<vector<float>>* plastdists = nullptr,
                                   int* pcount_descmatch = nullptr) const;
};
#include "sft_camera_fns.inc"
using Camera = GeometricCamera;  // camera_base.h:611

class PinholeCamera : public GeometricCamera {  // camera_pinhole.h:16-66: toK() as there, the copy-from-pointer constructor
 public:
  using Ptr = std::shared_ptr<PinholeCamera>;
  PinholeCamera() {}
  PinholeCamera(const PinholeCamera* o) : PinholeCamera(*o) {}
  Mat3data toK() const override {
    Mat3data K;
    K << fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f;
    return K;
  }
};
}  // namespace camm

class MapPoint {};

struct Converter {
  static Eigen::Vector3d toVector3d(const cv::Mat& m) { return Eigen::Vector3d((double)m.f[0], (double)m.f[1], (double)m.f[2]); }
};

class KeyFrame {  // the members SearchForTriangulation reads
 public:
  vector<camm::Camera::Ptr> mpCameras;
  bool usedistort_ = false;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat Rcw, tcw, Ow;  // CV_32F, as KeyFrame keeps them
  Sophus::SE3d Tcw, Twc;
  cv::Mat GetCameraCenter() { return Ow; }
  cv::Mat GetRotation() { return Rcw; }
  cv::Mat GetTranslation() { return tcw; }
  const Sophus::SE3d GetTcw() { return Tcw; }
  const Sophus::SE3d GetTwc() { return Twc; }
  vector<MapPoint*> mvpMapPoints;
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  struct {
    vector<float> vuright_;
  } stereoinfo_;
  struct {
    vector<float> vscalefactor_, vlevelsigma2_;
  } scalepyrinfo_;
  vector<cv::KeyPoint> mvKeys, mvKeysUn;
  vector<pair<size_t, size_t>> mapn2in_;
  cv::Mat mDescriptors;
};

class ORBmatcher {
 public:
  ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<vector<vector<size_t>>>& vMatchedPairs, const bool bOnlyStereo);
  float mfNNratio;
  bool mbCheckOrientation;
};
const int ORBmatcher::TH_HIGH = 100;  // src/ORBmatcher.cc:20-22
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
#include "orbmatcher_fns.inc"
#include "sft_fns.inc"
}  // namespace VIEO_SLAM_SFT
#undef cv

struct RefSftKp {  // == OrcKeyPoint
  float x, y, size, angle, response;
  int32_t octave;
};

namespace {
using namespace VIEO_SLAM_SFT;
Sophus::SE3d make_se3(const double q_wxyz[4], const double t[3]) {
  Eigen::Quaternion<double> q(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]);
  return Sophus::SE3d(Sophus::SO3<double>(q), Eigen::Vector3d(t[0], t[1], t[2]));
}
void fill_kf(KeyFrame& kf, MapPoint* some_mp, const float K[4], const double q_cw[4], const double t_cw[3], const RefSftKp* kp, const float* ur,
             const uint8_t* desc, const uint8_t* has_mp, int n_kp, const int32_t* fv_node, const int32_t* fv_ptr, const int32_t* fv_idx,
             int n_nodes, const float* scale_factor, const float* level_sigma2, int n_levels) {
  auto cam = std::make_shared<camm::PinholeCamera>();
  cam->fx = K[0]; cam->fy = K[1]; cam->cx = K[2]; cam->cy = K[3];
  kf.mpCameras.push_back(cam);
  kf.Tcw = make_se3(q_cw, t_cw);
  kf.Twc = kf.Tcw.inverse();
  // the CV_32F copies KeyFrame::SetPose keeps (Rcw, tcw, Ow = -Rcw^T tcw computed in double here and rounded once)
  const Eigen::Matrix3d R = kf.Tcw.rotationMatrix();
  kf.Rcw.rows = 3; kf.Rcw.cols = 3; kf.tcw.rows = 3; kf.tcw.cols = 1; kf.Ow.rows = 3; kf.Ow.cols = 1;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) kf.Rcw.f[3 * i + j] = (float)R(i, j);
    kf.tcw.f[i] = (float)kf.Tcw.translation()(i);
    kf.Ow.f[i] = (float)kf.Twc.translation()(i);
  }
  kf.mvKeysUn.resize(n_kp);
  for (int i = 0; i < n_kp; ++i) {
    kf.mvKeysUn[i].pt.x = kp[i].x; kf.mvKeysUn[i].pt.y = kp[i].y; kf.mvKeysUn[i].octave = kp[i].octave; kf.mvKeysUn[i].angle = kp[i].angle;
  }
  kf.mvKeys = kf.mvKeysUn;
  kf.stereoinfo_.vuright_.assign(ur, ur + n_kp);
  kf.mvpMapPoints.assign(n_kp, nullptr);
  for (int i = 0; i < n_kp; ++i)
    if (has_mp[i]) kf.mvpMapPoints[i] = some_mp;
  kf.mDescriptors = cvst_sft::Mat(desc, n_kp);
  for (int a = 0; a < n_nodes; ++a) {
    auto& v = kf.mFeatVec[(unsigned)fv_node[a]];
    for (int i = fv_ptr[a]; i < fv_ptr[a + 1]; ++i) v.push_back((unsigned)fv_idx[i]);
  }
  kf.scalepyrinfo_.vscalefactor_.assign(scale_factor, scale_factor + n_levels);
  kf.scalepyrinfo_.vlevelsigma2_.assign(level_sigma2, level_sigma2 + n_levels);
}
}  // namespace

// Two single-pinhole keyframes (K = fx, fy, cx, cy; pose as unit quaternion w, x, y, z + translation, camera <- world), keypoints,
// right coordinates, descriptors, "has a map point" flags and FeatureVectors flattened like orc_search_for_triangulation's.
// pairs [cap][2] = (idx1, idx2) of vMatchedPairs in order; returns nmatches, *n_pairs = vMatchedPairs.size().
extern "C" int ref_search_for_triangulation(const float K1[4], const double q1[4], const double t1[3], const RefSftKp* kp1, const float* ur1,
                                            const uint8_t* desc1, const uint8_t* has_mp1, int n_kp1, const int32_t* fv1_node,
                                            const int32_t* fv1_ptr, const int32_t* fv1_idx, int n_nodes1, const float K2[4], const double q2[4],
                                            const double t2[3], const RefSftKp* kp2, const float* ur2, const uint8_t* desc2,
                                            const uint8_t* has_mp2, int n_kp2, const int32_t* fv2_node, const int32_t* fv2_ptr,
                                            const int32_t* fv2_idx, int n_nodes2, const float* scale_factor, const float* level_sigma2,
                                            int n_levels, int only_stereo, int check_orientation, int32_t* pairs, int cap, int32_t* n_pairs) {
  MapPoint mp;
  KeyFrame a, b;
  fill_kf(a, &mp, K1, q1, t1, kp1, ur1, desc1, has_mp1, n_kp1, fv1_node, fv1_ptr, fv1_idx, n_nodes1, scale_factor, level_sigma2, n_levels);
  fill_kf(b, &mp, K2, q2, t2, kp2, ur2, desc2, has_mp2, n_kp2, fv2_node, fv2_ptr, fv2_idx, n_nodes2, scale_factor, level_sigma2, n_levels);
  ORBmatcher m(0.6f, check_orientation != 0);
  vector<vector<vector<size_t>>> matched;
  const int n = m.SearchForTriangulation(&a, &b, matched, only_stereo != 0);
  *n_pairs = (int32_t)matched.size();
  for (size_t i = 0; i < matched.size() && (int)i < cap; ++i) {
    pairs[2 * i] = (int32_t)matched[i][0][0];
    pairs[2 * i + 1] = (int32_t)matched[i][1][0];
  }
  return n;
}

// The geometry the caller of the flat interfaces (oracle, C ABI) hands over, formed from the same poses with the same stand-in
// operations in the order of src/ORBmatcher.cc:906-925, 1040-1045 and camera_base.h:290-293, 352: the epipole of camera 1 in image 2
// and F12 = K1^-T [t12]x R12 K2^-1 with T12 = (Tc1w * Twc2) rounded to float.
extern "C" void ref_sft_geometry(const float K1[4], const double q1[4], const double t1[3], const float K2[4], const double q2[4],
                                 const double t2[3], float* ex, float* ey, double F12[9]) {
  MapPoint mp;
  KeyFrame a, b;
  const float one = 1.f;
  fill_kf(a, &mp, K1, q1, t1, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, &one, &one, 1);
  fill_kf(b, &mp, K2, q2, t2, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, &one, &one, 1);
  const Vector3f C2 = Converter::toVector3d(b.GetRotation() * a.GetCameraCenter() + b.GetTranslation()).cast<float>();
  const float invz = 1.0f / C2(2);
  const Vector3f pn(C2(0) * invz, C2(1) * invz, 1);
  const Vector3f uv = b.mpCameras[0]->toK().cast<float>() * pn;
  *ex = uv[0];
  *ey = uv[1];
  const auto Tr1r2 = (a.GetTcw() * b.GetTwc()).cast<float>();
  const auto T12 = a.mpCameras[0]->GetTcr() * Tr1r2 * b.mpCameras[0]->GetTrc();
  const Eigen::Matrix3d R12 = T12.rotationMatrix().cast<double>();
  const Eigen::Vector3d t12 = T12.translation().cast<double>();
  const Eigen::Matrix3d Ka = a.mpCameras[0]->toK().cast<double>(), Kb = b.mpCameras[0]->toK().cast<double>();
  const Eigen::Matrix3d F = Ka.transpose().inverse() * Sophus::SO3ex<double>::hat(t12) * R12 * Kb.inverse();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F12[3 * i + j] = F(i, j);
}

// ORACLE — TEST INFRASTRUCTURE ONLY.
// The inertial arithmetic of the REFERENCE compiled UNCHANGED against a minimal stand-in for Eigen / Sophus::SO3 (eigstub/, both are
// third-party dependencies absent from this image):
//   * common/so3_extra.h, whole, from where it lies (SO3ex::exp / log / Log / Exp / JacobianR / JacobianRInv / normalizeRotationM);
//   * src/Odom/NavState.h, whole (the state and its IncSmall updates = the vertices' oplus);
//   * IMUPreIntegratorBase::update (src/Odom/OdomPreIntegrator.h:431-506): the covariance / bias-Jacobian / delta recurrence, cut out
//     by name at build time into oracle/_ref/gen/inertial_*.inc, like the vertex and edge classes below;
//   * VertexNavState<D>, VertexNavStateBias, VertexGThetaXYRwI and the edges EdgeNavStateI<NV> (PVR, PRV, PRVG), EdgeNavStateBias,
//     EdgeNavStatePriorPRVBias, EdgeNavStatePriorPVRBias, EdgeGyrBias (src/Odom/g2otypes.h, g2otypes.cpp): class definitions and
//     computeError / linearizeOplus bodies.
// What is declared here by hand is only what those texts name from g2o (vertex / edge base classes holding _estimate, _vertices,
// _error, _measurement and the Jacobian slots) and the data members of IMUDataBase / IMUPreIntegratorBase that update() touches.
// The C entry points mirror the oracle's (orc_edge_navstate, orc_navstate_oplus, ...) with the same struct layouts.
#include <math.h>
#include <stdlib.h>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include "common/so3_extra.h"    // the reference's, unchanged
#include "src/Odom/NavState.h"   // the reference's, unchanged

#include "../ba_oracle.h"  // struct layouts only (OrcNavState, OrcImuPreint, OrcImuNoise)

extern "C" void ref_cam_project(int model, const float* params, int n_params, const double P[3], float uv[2], double* J, double* Jp);

namespace G2O_INERTIAL {
template <int D>
struct JacDyn;
}
namespace Eigen {
template <int D>
struct traits<G2O_INERTIAL::JacDyn<D>> {
  typedef double Scalar;
  enum { Rows = D, Cols = 15 };
};
}  // namespace Eigen

namespace VIEO_SLAM_INERTIAL {  // own namespaces: other wrappers of this library declare stand-ins under the reference's names
using namespace VIEO_SLAM;
typedef Eigen::Matrix<double, 9, 9> Matrix9d;
typedef NavState NavStated;

struct IMUDataBase {  // src/Odom/OdomData.h:22-36: the statics update() reads
  static Matrix3d mSigmag, mSigmaa;
  static int mdt_cov_noise_fixed;
  static double mFreqRef;
};
Matrix3d IMUDataBase::mSigmag, IMUDataBase::mSigmaa;
int IMUDataBase::mdt_cov_noise_fixed = 0;
double IMUDataBase::mFreqRef = 0;

template <class IMUDataBase>
class IMUPreIntegratorBase {  // data members of OdomPreIntegrator.h:105-150
 public:
  typedef double Tcalc;
  using SO3calc = Sophus::SO3ex<Tcalc>;
  double mdeltatij = 0;
  Matrix3d mRij;
  Vector3d mvij, mpij;
  Matrix9d mSigmaijPRV, mSigmaij;
  Matrix3d mJgpij, mJapij, mJgvij, mJavij, mJgRij;
  IMUPreIntegratorBase() { mRij.setIdentity(); }
  void update(const Vector3d& omega, const Vector3d& acc, const double& dt);
};
template <class IMUDataBase>
#include "inertial_update.inc"

typedef IMUPreIntegratorBase<IMUDataBase> IMUPreintegrator;

// what EdgeReproject names from common/camera_models: a camera whose Project() is the reference's own (the three models' Project
// bodies are compiled unchanged in ref_camera_wrap.cc of this library) and a camera-to-reference-camera transform (identity here)
typedef float FLT_CAMM;
namespace camm {
struct SE3Standin {
  Eigen::Matrix<FLT_CAMM, 3, 3> R = Eigen::Matrix<FLT_CAMM, 3, 3>::Identity();
  Eigen::Matrix<FLT_CAMM, 3, 1> t;
  const Eigen::Matrix<FLT_CAMM, 3, 3>& rotationMatrix() const { return R; }
  const Eigen::Matrix<FLT_CAMM, 3, 1>& translation() const { return t; }
};
class Camera {
 public:
  int model = 0;
  std::vector<float> params;
  SE3Standin Tcr;
  const SE3Standin& GetTcr() const { return Tcr; }
  void Project(const Vector3d& p, Eigen::Matrix<FLT_CAMM, 2, 1>* img, Eigen::Matrix<double, 2, 3>* J = nullptr) const {
    float uv[2];
    double j[6];
    const double P[3] = {p(0), p(1), p(2)};
    ref_cam_project(model, params.data(), (int)params.size(), P, uv, J ? j : nullptr, nullptr);
    if (img) (*img)(0) = uv[0], (*img)(1) = uv[1];
    if (J)
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) (*J)(r, c) = j[3 * r + c];
  }
};
}  // namespace camm
}  // namespace VIEO_SLAM_INERTIAL

namespace G2O_INERTIAL {
using namespace VIEO_SLAM_INERTIAL;
using namespace Eigen;

struct OptimizableGraph {
  struct Vertex {
    virtual ~Vertex() {}
  };
};
template <int D, class T>
class BaseVertex : public OptimizableGraph::Vertex {
 public:
  static const int Dimension = D;
  const T& estimate() const { return _estimate; }
  void setEstimate(const T& e) { _estimate = e; }

 protected:
  T _estimate;
};
class VertexSBAPointXYZ : public BaseVertex<3, Vector3d> {};

// what the fork's getHessian() members return for a multi edge (Eigen::MatrixXd there): a small row-major dynamic matrix with the
// three products `jac.transpose() * rinfo * jac` needs, every sum in index order
struct DynMat : Eigen::DynamicRhsTag {
  int r = 0, c = 0;
  std::vector<double> v;
  DynMat() {}
  DynMat(int r_, int c_) : r(r_), c(c_), v((size_t)r_ * c_, 0.0) {}
  template <class O>
  DynMat(const Eigen::MatrixBase<O>& o) : r(O::Rows), c(O::Cols), v((size_t)O::Rows * O::Cols) {
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) v[(size_t)i * c + j] = o.coeff(i, j);
  }
  int rows() const { return r; }
  int cols() const { return c; }
  double coeff(int i, int j) const { return v[(size_t)i * c + j]; }
  double& coeffRef(int i, int j) { return v[(size_t)i * c + j]; }
  DynMat transpose() const {
    DynMat t(c, r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) t.coeffRef(j, i) = coeff(i, j);
    return t;
  }
};
template <class B>
DynMat operator*(const DynMat& a, const Eigen::MatrixBase<B>& b) {
  assert(a.c == (int)B::Rows);
  DynMat o(a.r, (int)B::Cols);
  for (int i = 0; i < a.r; ++i)
    for (int j = 0; j < (int)B::Cols; ++j) {
      double s = 0;
      for (int k = 0; k < a.c; ++k) s += a.coeff(i, k) * b.coeff(k, j);
      o.coeffRef(i, j) = s;
    }
  return o;
}
template <int D>
struct JacDyn;
template <int D>
DynMat operator*(const DynMat& a, const JacDyn<D>& b);
// one _jacobianOplus[i] of a multi edge (g2o: a map with D rows and as many columns as the vertex has dimensions), with the
// handful of operations EdgeReproject::linearizeOplus applies to it
template <int D>
struct JacDyn {
  int cols = 0;
  double v[D * 15] = {};  // column-major
  double coeff(int i, int j) const { return v[j * D + i]; }
  double& coeffRef(int i, int j) { return v[j * D + i]; }
  template <class O>
  JacDyn& operator=(const Eigen::MatrixBase<O>& o) {
    static_assert((int)O::Rows == D, "rows");
    cols = O::Cols;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < cols; ++j) coeffRef(i, j) = o.coeff(i, j);
    return *this;
  }
  template <int C>
  operator Matrix<double, D, C>() const {
    assert(cols == C);
    Matrix<double, D, C> m;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < C; ++j) m(i, j) = coeff(i, j);
    return m;
  }
  template <int BR, int BC>
  Eigen::Block<JacDyn, BR, BC> block(int i, int j) {
    assert(j + BC <= cols);
    return Eigen::Block<JacDyn, BR, BC>(*this, i, j);
  }
  template <class B>
  JacDyn& operator*=(const Eigen::MatrixBase<B>& b) {
    static_assert((int)B::Rows == (int)B::Cols, "square");
    const Matrix<double, D, B::Cols> r = static_cast<Matrix<double, D, B::Rows>>(*this) * b;
    return *this = r;
  }
  JacDyn& operator*=(double s) {
    for (int k = 0; k < D * cols; ++k) v[k] *= s;
    return *this;
  }
  DynMat transpose() const {
    DynMat t(cols, D);
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < cols; ++j) t.coeffRef(j, i) = coeff(i, j);
    return t;
  }
};
template <int D>
DynMat operator*(const DynMat& a, const JacDyn<D>& b) {
  assert(a.c == D);
  DynMat o(a.r, b.cols);
  for (int i = 0; i < a.r; ++i)
    for (int j = 0; j < b.cols; ++j) {
      double s = 0;
      for (int k = 0; k < D; ++k) s += a.coeff(i, k) * b.coeff(k, j);
      o.coeffRef(i, j) = s;
    }
  return o;
}
template <int D, class B>
Matrix<double, D, Eigen::traits<B>::Cols> operator*(const JacDyn<D>& a, const Eigen::MatrixBase<B>& b) {
  return static_cast<Matrix<double, D, Eigen::traits<B>::Rows>>(a) * b;
}
template <int D, class B>
Matrix<double, D, Eigen::traits<B>::Cols> operator+(const JacDyn<D>& a, const Eigen::MatrixBase<B>& b) {
  return static_cast<Matrix<double, D, Eigen::traits<B>::Cols>>(a) + b;
}
// g2o's robust kernel interface and this fork's Huber kernel (core/robust_kernel_impl.cpp, cut out by name like in ref_kernel_wrap.cc)
class RobustKernel {
 public:
  virtual ~RobustKernel() {}
  virtual void robustify(double e2, Vector3d& rho) const = 0;
  double _delta = 1.;
};
class RobustKernelHuber : public RobustKernel {
 public:
  virtual void setDelta(double delta);
  virtual void setDeltaSqr(const double& delta, const double& deltaSqr);
  virtual void robustify(double e2, Vector3d& rho) const;
  float dsqr;
};
#include "kernel_fns.inc"
namespace internal {
inline int computeUpperTriangleIndex(int i, int j) { return j * (j - 1) / 2 + i; }  // core/base_multi_edge.h
}
// what g2o's BaseEdge gives every edge: error, information, chi2() = e^T Omega e, the robust kernel, the level
template <int D>
class EdgeCommon {
 public:
  typedef Matrix<double, D, D> InformationType;
  virtual ~EdgeCommon() {}
  const InformationType& information() const { return _information; }
  InformationType& information() { return _information; }
  void setInformation(const InformationType& i) { _information = i; }
  double chi2() const { return _error.dot(_information * _error); }
  RobustKernel* robustKernel() const { return _robustKernel; }
  void setRobustKernel(RobustKernel* k) { _robustKernel = k; }
  InformationType robustInformation(const Vector3d& rho) { return rho[1] * _information; }
  int level() const { return _level; }
  void setLevel(int l) { _level = l; }
  Matrix<double, D, 1> _error;
  InformationType _information = InformationType::Identity();
  RobustKernel* _robustKernel = nullptr;
  int _level = 0;
};
template <int D, class E>
class BaseMultiEdge : public EdgeCommon<D> {
 public:
  typedef JacDyn<D> JacobianType;
  struct HessianHelper {
    DynMat matrix;
    bool transposed = false;
  };
  void resize(size_t n) {
    _vertices.assign(n, nullptr);
    _jacobianOplus.resize(n);
  }
  virtual void computeError() = 0;
  virtual void linearizeOplus() = 0;
  std::vector<OptimizableGraph::Vertex*>& vertices() { return _vertices; }
  std::vector<OptimizableGraph::Vertex*> _vertices;
  using EdgeCommon<D>::_error;
  E _measurement;
  std::vector<JacDyn<D>> _jacobianOplus;
  std::vector<HessianHelper> _hessian;
};
template <int D, class E, class Xi, class Xj>
class BaseBinaryEdge : public EdgeCommon<D> {
 public:
  typedef Matrix<double, D, Xi::Dimension> JacobianXiOplusType;
  typedef Matrix<double, D, Xj::Dimension> JacobianXjOplusType;
  typedef Matrix<double, Xi::Dimension, Xj::Dimension> HessianBlockType;
  typedef Matrix<double, Xj::Dimension, Xi::Dimension> HessianBlockTransposedType;
  virtual void computeError() = 0;
  virtual void linearizeOplus() = 0;
  OptimizableGraph::Vertex* _vertices[2] = {nullptr, nullptr};
  using EdgeCommon<D>::_error;
  E _measurement;
  JacobianXiOplusType _jacobianOplusXi;
  JacobianXjOplusType _jacobianOplusXj;
  HessianBlockType _hessian;
  HessianBlockTransposedType _hessianTransposed;
  bool _hessianRowMajor = false;
};
template <int D, class E, class Xi>
class BaseUnaryEdge : public EdgeCommon<D> {
 public:
  typedef Matrix<double, D, Xi::Dimension> JacobianXiOplusType;
  virtual void computeError() = 0;
  virtual void linearizeOplus() = 0;
  const E& measurement() const { return _measurement; }
  OptimizableGraph::Vertex** vertices() { return _vertices; }
  OptimizableGraph::Vertex* _vertices[1] = {nullptr};
  using EdgeCommon<D>::_error;
  E _measurement;
  JacobianXiOplusType _jacobianOplusXi;
};
// the fork's extension of the three edge bases (src/Odom/g2otypes.h:33-254: getRho, getHessian*), cut out by name
typedef DynMat MatrixXd;
typedef enum HessianExactMode { kExactNoRobust, kExactRobust, kNotExact } eHessianExactMode;  // g2otypes.h:34
template <int D, typename E, typename VertexXi>
#include "inertial_base_unary_ex.inc"
template <int D, typename E, typename VertexXi, typename VertexXj>
#include "inertial_base_binary_ex.inc"
template <int D, typename E>
#include "inertial_base_multi_ex.inc"
template <int D, class M, class I>
bool readEdge(std::istream&, M&, I&) {
  return true;
}
template <int D, class M, class I>
bool writeEdge(std::ostream&, const M&, const I&) {
  return true;
}

// class definitions and member bodies of the reference, in the order of its header
template <int D>
#include "inertial_vertex_navstate.inc"
typedef VertexNavState<6> VertexNavStatePR;
typedef VertexNavState<3> VertexNavStateV;
typedef VertexNavState<9> VertexNavStatePVR;
#include "inertial_vertex_bias.inc"
#include "inertial_vertex_gtheta.inc"
template <int NV>
#include "inertial_edge_navstate_class.inc"
template <int NV>
#include "inertial_edge_navstate_error.inc"
template <int NV>
#include "inertial_edge_navstate_jac.inc"
typedef EdgeNavStateI<5> EdgeNavStatePRV;
typedef EdgeNavStateI<6> EdgeNavStatePRVG;
typedef EdgeNavStateI<3> EdgeNavStatePVR;
#include "inertial_edge_classes.inc"
typedef VertexSBAPointXYZ VertexGyrBias;
#include "inertial_edge_gyrbias_class.inc"
#include "inertial_edge_fns.inc"

// the visual edge: VertexScale, EdgeReproject<DE, DV, NV, MODE_OPT_VAR> (class with GetTcw_wX / computeError) and its linearizeOplus
using Vector2img = Eigen::Matrix<FLT_CAMM, 2, 1>;
#include "visual_vertex_scale.inc"
template <int DE, int DV, int NV, int MODE_OPT_VAR = 0>
#include "visual_edge_reproject_class.inc"
template <int DE, int DV, int NV, int MODE_OPT_VAR>
#include "visual_edge_reproject_jac.inc"
typedef EdgeReproject<2, 9, 2> EdgeReprojectPVR;  // g2otypes.h:546-547
typedef EdgeReproject<3, 9, 2> EdgeReprojectPVRStereo;
struct EdgeEncNavStatePVR {  // the wheel-encoder edge: never present on this path (FillCovInv receives nullptr)
  void linearizeOplus() {}
  Matrix<double, 9, 9> getHessianXi(bool = true) const { return Matrix<double, 9, 9>(); }
  Matrix<double, 9, 9> getHessianXj(bool = true) const { return Matrix<double, 9, 9>(); }
  Matrix<double, 9, 9> getHessianXji(int8_t = 0) const { return Matrix<double, 9, 9>(); }
};
}  // namespace G2O_INERTIAL

namespace g2o = G2O_INERTIAL;
// Optimizer::FillCovInv (include/Optimizer.h:126-206), the explicit J^T (rho' Omega) J assembly of PoseOptimization's marginal, cut
// out by name
namespace VIEO_SLAM_INERTIAL {
using namespace Eigen;
using std::vector;
class Optimizer {
 public:
  template <class MatrixNVd>
  static void FillCovInv(g2o::EdgeNavStatePVR* eNSPVR, g2o::EdgeNavStateBias* eNSBias, g2o::EdgeEncNavStatePVR* eEnc, const int8_t schur_bec,
                         const vector<g2o::EdgeReprojectPVR*>* pvpEdgesMono, const vector<g2o::EdgeReprojectPVRStereo*>* pvpEdgesStereo,
                         MatrixNVd& cov_inv, g2o::EdgeNavStatePriorPVRBias* eNSPrior = nullptr,
                         const int8_t exact_mode = (int8_t)g2o::kExactRobust);
};
template <class MatrixNVd>
#include "fill_cov_inv.inc"
}  // namespace VIEO_SLAM_INERTIAL
namespace {
using namespace VIEO_SLAM_INERTIAL;
using Eigen::Matrix;
using Eigen::Quaterniond;
using Sophus::SO3exd;

NavState to_ns(const OrcNavState& s) {
  NavState n;
  n.mpwb = Vector3d(s.p);
  n.mRwb = SO3exd(Quaterniond(s.q[0], s.q[1], s.q[2], s.q[3]));
  n.mvwb = Vector3d(s.v);
  n.mbg = Vector3d(s.bg), n.mba = Vector3d(s.ba), n.mdbg = Vector3d(s.dbg), n.mdba = Vector3d(s.dba);
  return n;
}
void from_ns(const NavState& n, OrcNavState* s) {
  const Quaterniond& q = n.mRwb.unit_quaternion();
  s->q[0] = q.w(), s->q[1] = q.x(), s->q[2] = q.y(), s->q[3] = q.z();
  for (int k = 0; k < 3; ++k) {
    s->p[k] = n.mpwb(k), s->v[k] = n.mvwb(k);
    s->bg[k] = n.mbg(k), s->ba[k] = n.mba(k), s->dbg[k] = n.mdbg(k), s->dba[k] = n.mdba(k);
  }
}
template <int R, int C>
Matrix<double, R, C> from_rm(const double* p) {
  Matrix<double, R, C> m;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) m(i, j) = p[i * C + j];
  return m;
}
template <class M>
void to_rm(const Eigen::MatrixBase<M>& m, double* p) {
  for (int i = 0; i < (int)M::Rows; ++i)
    for (int j = 0; j < (int)M::Cols; ++j) p[i * (int)M::Cols + j] = m.coeff(i, j);
}
IMUPreintegrator to_pre(const OrcImuPreint& o) {
  IMUPreintegrator p;
  p.mRij = from_rm<3, 3>(o.Rij);
  p.mvij = Vector3d(o.vij), p.mpij = Vector3d(o.pij);
  p.mSigmaijPRV = from_rm<9, 9>(o.SigmaPRV), p.mSigmaij = from_rm<9, 9>(o.SigmaPVR);
  p.mJgpij = from_rm<3, 3>(o.Jgp), p.mJapij = from_rm<3, 3>(o.Jap), p.mJgvij = from_rm<3, 3>(o.Jgv);
  p.mJavij = from_rm<3, 3>(o.Jav), p.mJgRij = from_rm<3, 3>(o.JgR);
  p.mdeltatij = o.dt;
  return p;
}
// the fork's *EdgeEx classes re-declare the Jacobian members protected: read them through the (stand-in) g2o base
template <int D, class E>
g2o::BaseMultiEdge<D, E>& base_of(g2o::BaseMultiEdge<D, E>& e) {
  return e;
}
template <int D, class E, class Xi, class Xj>
g2o::BaseBinaryEdge<D, E, Xi, Xj>& base_of(g2o::BaseBinaryEdge<D, E, Xi, Xj>& e) {
  return e;
}
template <int D, class E, class Xi>
g2o::BaseUnaryEdge<D, E, Xi>& base_of(g2o::BaseUnaryEdge<D, E, Xi>& e) {
  return e;
}
// copies a slot (rows x cols) into dst [rows][ld] at column c0
template <int D>
void put_cols(const g2o::JacDyn<D>& s, double* dst, int ld, int c0) {
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < s.cols; ++j) dst[i * ld + c0 + j] = s.coeff(i, j);
}
}  // namespace

// op: 0 exp(w) -> quaternion (w, x, y, z); 1 Exp(w) -> R; 2 SO3ex(q).log(); 3 Log(R); 4 JacobianR(w); 5 JacobianRInv(w);
//     6 normalizeRotationM(R); 7 the Sophus base class' log() of q (what a product of two SO3ex yields in this build)
extern "C" void ref_so3(int op, const double* in, double* out) {
  if (op == 0) {
    const Quaterniond q = SO3exd::exp(Vector3d(in)).unit_quaternion();
    out[0] = q.w(), out[1] = q.x(), out[2] = q.y(), out[3] = q.z();
  } else if (op == 1) {
    to_rm(SO3exd::Exp(Vector3d(in)), out);
  } else if (op == 2) {
    to_rm(SO3exd(Quaterniond(in[0], in[1], in[2], in[3])).log(), out);
  } else if (op == 3) {
    to_rm(SO3exd::Log(from_rm<3, 3>(in)), out);
  } else if (op == 4) {
    to_rm(SO3exd::JacobianR(Vector3d(in)), out);
  } else if (op == 5) {
    to_rm(SO3exd::JacobianRInv(Vector3d(in)), out);
  } else if (op == 6) {
    to_rm(SO3exd::normalizeRotationM(from_rm<3, 3>(in)), out);
  } else if (op == 7) {
    const Sophus::SO3<double> s = SO3exd(Quaterniond(in[0], in[1], in[2], in[3])) * SO3exd();
    to_rm(s.log(), out);
  }
}

// reset state, then update(omega, acc, dt) for every row of trace [n][7]
extern "C" void ref_imu_update_sequence(const OrcImuNoise* nz, const double* trace, int n, OrcImuPreint* out) {
  IMUDataBase::mSigmag = Matrix3d::Identity() * nz->sigma_g;
  IMUDataBase::mSigmaa = Matrix3d::Identity() * nz->sigma_a;
  IMUDataBase::mdt_cov_noise_fixed = nz->dt_cov_noise_fixed;
  IMUDataBase::mFreqRef = nz->freq_ref;
  IMUPreintegrator p;
  for (int i = 0; i < n; ++i) p.update(Vector3d(trace + 7 * i), Vector3d(trace + 7 * i + 3), trace[7 * i + 6]);
  std::memset(out, 0, sizeof(*out));
  to_rm(p.mRij, out->Rij), to_rm(p.mvij, out->vij), to_rm(p.mpij, out->pij);
  to_rm(p.mSigmaijPRV, out->SigmaPRV), to_rm(p.mSigmaij, out->SigmaPVR);
  to_rm(p.mJgpij, out->Jgp), to_rm(p.mJapij, out->Jap), to_rm(p.mJgvij, out->Jgv), to_rm(p.mJavij, out->Jav), to_rm(p.mJgRij, out->JgR);
  out->dt = p.mdeltatij;
}

// EdgeNavStateI<3> (order 0, PVR), <5> (order 1, PRV) or <6> (q_wI != nullptr: PRVG with gw = GI): e [9], Ji / Jj [9][9] with
// the state columns in the residual's order, Jb [9][6], JG [9][2]
extern "C" void ref_edge_navstate(const OrcNavState* nsi, const OrcNavState* nsj, const OrcImuPreint* pre, const double gw[3], int order,
                                  const double* q_wI, double e[9], double* Ji, double* Jj, double* Jb, double* JG) {
  using namespace g2o;
  const NavState a = to_ns(*nsi), b = to_ns(*nsj);
  VertexNavStateBias vb;
  vb.setEstimate(a);
  if (order == 0) {
    VertexNavStatePVR vi, vj;
    vi.setEstimate(a), vj.setEstimate(b);
    EdgeNavStatePVR ed;
    ed._vertices[0] = &vi, ed._vertices[1] = &vj, ed._vertices[2] = &vb;
    ed._measurement = to_pre(*pre);
    ed.SetParams(Vector3d(gw));
    ed.computeError();
    to_rm(ed._error, e);
    if (Ji) {
      ed.linearizeOplus();
      put_cols(base_of(ed)._jacobianOplus[0], Ji, 9, 0), put_cols(base_of(ed)._jacobianOplus[1], Jj, 9, 0), put_cols(base_of(ed)._jacobianOplus[2], Jb, 6, 0);
    }
    return;
  }
  VertexNavStatePR pi, pj;
  VertexNavStateV wi, wj;
  pi.setEstimate(a), pj.setEstimate(b), wi.setEstimate(a), wj.setEstimate(b);
  auto run = [&](auto& ed) {
    ed._vertices[0] = &pi, ed._vertices[1] = &pj, ed._vertices[2] = &wi, ed._vertices[3] = &wj, ed._vertices[4] = &vb;
    ed._measurement = to_pre(*pre);
    ed.SetParams(Vector3d(gw));
    ed.computeError();
    to_rm(ed._error, e);
    if (Ji) {
      ed.linearizeOplus();
      put_cols(base_of(ed)._jacobianOplus[0], Ji, 9, 0), put_cols(base_of(ed)._jacobianOplus[2], Ji, 9, 6);
      put_cols(base_of(ed)._jacobianOplus[1], Jj, 9, 0), put_cols(base_of(ed)._jacobianOplus[3], Jj, 9, 6);
      put_cols(base_of(ed)._jacobianOplus[4], Jb, 6, 0);
    }
  };
  if (!q_wI) {
    EdgeNavStatePRV ed;
    run(ed);
  } else {
    VertexGThetaXYRwI vg;
    vg.setEstimate(SO3exd(Quaterniond(q_wI[0], q_wI[1], q_wI[2], q_wI[3])));
    EdgeNavStatePRVG ed;
    ed._vertices[5] = &vg;
    run(ed);
    if (Ji && JG) put_cols(base_of(ed)._jacobianOplus[5], JG, 2, 0);
  }
}

// kind 0 = PR (6), 1 = PVR (9), 2 = V (3), 3 = Bias (6): the vertices' oplusImpl
extern "C" void ref_navstate_oplus(OrcNavState* ns, int kind, const double* dx) {
  using namespace g2o;
  auto run = [&](auto& v) {
    v.setEstimate(to_ns(*ns));
    v.oplusImpl(dx);
    from_ns(v.estimate(), ns);
  };
  if (kind == 0) {
    VertexNavStatePR v;
    run(v);
  } else if (kind == 1) {
    VertexNavStatePVR v;
    run(v);
  } else if (kind == 2) {
    VertexNavStateV v;
    run(v);
  } else {
    VertexNavStateBias v;
    run(v);
  }
}

// VertexGThetaXYRwI::setToOriginImpl(gw) / oplusImpl
extern "C" void ref_gdir(int op, const double* in, double q_wI[4]) {
  g2o::VertexGThetaXYRwI v;
  if (op == 0) {
    Vector3d gw(in);
    v.setToOriginImpl(gw);
  } else {
    v.setEstimate(SO3exd(Quaterniond(q_wI[0], q_wI[1], q_wI[2], q_wI[3])));
    v.oplusImpl(in);
  }
  const Quaterniond& q = v.estimate().unit_quaternion();
  q_wI[0] = q.w(), q_wI[1] = q.x(), q_wI[2] = q.y(), q_wI[3] = q.z();
}

// EdgeNavStatePriorPVRBias (form 0: e [15], Jpvr [15][9], Jbias [15][6]) / EdgeNavStatePriorPRVBias (form 1: Jpvr holds
// [PR | V] = 15 x 9 in P R V order)
extern "C" void ref_edge_prior(int form, const OrcNavState* ns, const OrcNavState* prior, double e[15], double* Jpvr, double* Jbias) {
  using namespace g2o;
  const NavState s = to_ns(*ns);
  VertexNavStateBias vb;
  vb.setEstimate(s);
  if (form == 0) {
    VertexNavStatePVR v;
    v.setEstimate(s);
    EdgeNavStatePriorPVRBias ed;
    ed._vertices[0] = &v, ed._vertices[1] = &vb;
    ed._measurement = to_ns(*prior);
    ed.computeError();
    to_rm(ed._error, e);
    if (Jpvr) {
      ed.linearizeOplus();
      to_rm(base_of(ed)._jacobianOplusXi, Jpvr);
      if (Jbias) to_rm(base_of(ed)._jacobianOplusXj, Jbias);
    }
  } else {
    VertexNavStatePR vp;
    VertexNavStateV vv;
    vp.setEstimate(s), vv.setEstimate(s);
    EdgeNavStatePriorPRVBias ed;
    ed._vertices[0] = &vp, ed._vertices[1] = &vv, ed._vertices[2] = &vb;
    ed._measurement = to_ns(*prior);
    ed.computeError();
    to_rm(ed._error, e);
    if (Jpvr) {
      ed.linearizeOplus();
      put_cols(ed._jacobianOplus[0], Jpvr, 9, 0), put_cols(ed._jacobianOplus[1], Jpvr, 9, 6);
      if (Jbias) put_cols(ed._jacobianOplus[2], Jbias, 6, 0);
    }
  }
}

// EdgeNavStateBias: e [6] = (bg_j + dbg_j) - (bg_i + dbg_i), same for ba; Ji / Jj [6][6]
extern "C" void ref_edge_bias(const OrcNavState* nsi, const OrcNavState* nsj, double e[6], double* Ji, double* Jj) {
  using namespace g2o;
  VertexNavStateBias vi, vj;
  vi.setEstimate(to_ns(*nsi)), vj.setEstimate(to_ns(*nsj));
  EdgeNavStateBias ed;
  ed._vertices[0] = &vi, ed._vertices[1] = &vj;
  ed.computeError();
  to_rm(ed._error, e);
  if (Ji) {
    ed.linearizeOplus();
    to_rm(base_of(ed)._jacobianOplusXi, Ji), to_rm(base_of(ed)._jacobianOplusXj, Jj);
  }
}

// EdgeGyrBias (Optimizer::OptimizeInitialGyroBias' edge): matrices row-major
extern "C" void ref_edge_gyr_bias(const double* dRij, const double* JgRij, const double* Rwbi, const double* Rwbj, const double bg[3],
                                  double e[3], double* J) {
  using namespace g2o;
  VertexGyrBias v;
  v.setEstimate(Vector3d(bg));
  EdgeGyrBias ed;
  ed._vertices[0] = &v;
  ed.deltaRij = from_rm<3, 3>(dRij), ed.JgRij = from_rm<3, 3>(JgRij), ed.Rwbi = from_rm<3, 3>(Rwbi), ed.Rwbj = from_rm<3, 3>(Rwbj);
  ed.computeError();
  to_rm(ed._error, e);
  if (J) {
    ed.linearizeOplus();
    to_rm(ed._jacobianOplusXi, J);
  }
}

// EdgeReproject<DE, DV, NV, MODE>: form 0 PR (2,6,2), 1 PRStereo (3,6,2), 2 PVR (2,9,2), 3 PVRStereo (3,9,2), 4 PRS (2,6,3,1),
// 5 PRSStereo (3,6,3,1), 6 PRSInv (2,6,3,2).  X = the point vertex' estimate, scale = the VertexScale estimate (forms >= 4).
// e [DE], J_pose [DE][DV], J_point [DE][3], J_scale [DE] (forms >= 4), depth = GetDepth()
namespace {
template <int DE, int DV, int NV, int MODE>
void run_reproject(const OrcCamera* cam, const OrcNavState* ns, const double X[3], const float* obs, double scale, double* e,
                   double* J_pose, double* J_point, double* J_scale, double* depth) {
  using namespace g2o;
  camm::Camera c;
  c.model = cam->model;
  c.params = {cam->fx, cam->fy, cam->cx, cam->cy};
  const int nd = cam->model == 1 ? cam->num_k + 2 : cam->model == 2 ? 4 : 0;
  for (int k = 0; k < nd; ++k) c.params.push_back(cam->dist[k]);
  VertexSBAPointXYZ vx;
  vx.setEstimate(Vector3d(X));
  VertexNavState<DV> vn;
  vn.setEstimate(to_ns(*ns));
  VertexScale vs;
  vs.setEstimate(scale);
  EdgeReproject<DE, DV, NV, MODE> edge;
  BaseMultiEdge<DE, Matrix<double, DE, 1>>& ed = edge;  // the class re-declares the base members protected
  ed._vertices[0] = &vx, ed._vertices[1] = &vn;
  if (NV > 2) ed._vertices[2] = &vs;
  ed._jacobianOplus[0].cols = 3, ed._jacobianOplus[1].cols = DV;
  if (NV > 2) ed._jacobianOplus[2].cols = 1;
  const float bf = cam->bf;
  edge.SetParams(&c, from_rm<3, 3>(cam->Rcb), Vector3d(cam->tcb), &bf);
  for (int k = 0; k < DE; ++k) ed._measurement(k) = (double)obs[k];
  ed.computeError();
  to_rm(ed._error, e);
  if (depth) *depth = edge.GetDepth();
  if (J_pose) {
    ed.linearizeOplus();
    put_cols(ed._jacobianOplus[1], J_pose, DV, 0);
    put_cols(ed._jacobianOplus[0], J_point, 3, 0);
    if (NV > 2) put_cols(ed._jacobianOplus[2], J_scale, 1, 0);
  }
}
}  // namespace
extern "C" void ref_edge_reproject(int form, const OrcCamera* cam, const OrcNavState* ns, const double X[3], const float* obs, double scale,
                                   double* e, double* J_pose, double* J_point, double* J_scale, double* depth) {
  if (form == 0) run_reproject<2, 6, 2, 0>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
  if (form == 1) run_reproject<3, 6, 2, 0>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
  if (form == 2) run_reproject<2, 9, 2, 0>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
  if (form == 3) run_reproject<3, 9, 2, 0>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
  if (form == 4) run_reproject<2, 6, 3, 1>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
  if (form == 5) run_reproject<3, 6, 3, 1>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
  if (form == 6) run_reproject<2, 6, 3, 2>(cam, ns, X, obs, scale, e, J_pose, J_point, J_scale, depth);
}

// Optimizer::FillCovInv on the edges of an inertial PoseOptimization at its final estimate: the inertial edge (last PVR, current PVR,
// last bias), the bias random-walk edge, the prior edge of the last keyframe (when it is free) and the visual PVR edges with their
// levels, information matrices and Huber kernels as given (delta < 0: no kernel).  Errors are computed first, as the caller does
// (include/Optimizer.h:671-676).  C / CL / CCL [15][15] row-major = cov_inv for schur_bec 0 / 2 / 1 (CL, CCL only when prior != NULL).
extern "C" void ref_fill_cov_inv(const OrcCamera* cam, const OrcNavState* cur, const OrcNavState* last, const OrcImuPreint* pre,
                                 const double gw[3], const double* info_imu, double delta_imu, const double* info_bias, double delta_bias,
                                 const OrcNavState* prior, const double* info_prior, double delta_prior, int n_vis, const double* X,
                                 const float* obs, const uint8_t* stereo, const double* w, const int32_t* level, const double* delta,
                                 double* C, double* CL, double* CCL) {
  using namespace g2o;
  camm::Camera c;
  c.model = cam->model;
  c.params = {cam->fx, cam->fy, cam->cx, cam->cy};
  const int nd = cam->model == 1 ? cam->num_k + 2 : cam->model == 2 ? 4 : 0;
  for (int k = 0; k < nd; ++k) c.params.push_back(cam->dist[k]);
  const float bf = cam->bf;
  VertexNavStatePVR vCur, vLast;
  VertexNavStateBias vbCur, vbLast;
  vCur.setEstimate(to_ns(*cur)), vbCur.setEstimate(to_ns(*cur));
  vLast.setEstimate(to_ns(*last)), vbLast.setEstimate(to_ns(*last));
  std::vector<std::unique_ptr<RobustKernelHuber>> kernels;
  auto kernel = [&](double d) -> RobustKernel* {
    if (d < 0) return nullptr;
    kernels.emplace_back(new RobustKernelHuber);
    kernels.back()->setDelta(d);
    return kernels.back().get();
  };
  std::unique_ptr<EdgeNavStatePVR> eI;
  if (pre) {
    eI.reset(new EdgeNavStatePVR);
    auto& b = base_of(*eI);
    b._vertices[0] = &vLast, b._vertices[1] = &vCur, b._vertices[2] = &vbLast;
    b._measurement = to_pre(*pre);
    eI->SetParams(Vector3d(gw));
    b.setInformation(from_rm<9, 9>(info_imu));
    b.setRobustKernel(kernel(delta_imu));
    b.computeError();
  }
  EdgeNavStateBias eB;
  {
    auto& b = base_of(eB);
    b._vertices[0] = &vbLast, b._vertices[1] = &vbCur;
    b.setInformation(from_rm<6, 6>(info_bias));
    b.setRobustKernel(kernel(delta_bias));
    b.computeError();
  }
  std::unique_ptr<EdgeNavStatePriorPVRBias> eP;
  if (prior) {
    eP.reset(new EdgeNavStatePriorPVRBias);
    auto& b = base_of(*eP);
    b._vertices[0] = &vLast, b._vertices[1] = &vbLast;
    b._measurement = to_ns(*prior);
    b.setInformation(from_rm<15, 15>(info_prior));
    b.setRobustKernel(kernel(delta_prior));
    b.computeError();
  }
  std::vector<VertexSBAPointXYZ> pts(n_vis);
  std::vector<std::unique_ptr<EdgeReprojectPVR>> mono_own;
  std::vector<std::unique_ptr<EdgeReprojectPVRStereo>> stereo_own;
  std::vector<EdgeReprojectPVR*> mono;
  std::vector<EdgeReprojectPVRStereo*> ster;
  const Matrix3d Rcb = from_rm<3, 3>(cam->Rcb);
  const Vector3d tcb(cam->tcb);
  for (int i = 0; i < n_vis; ++i) {
    pts[i].setEstimate(Vector3d(X + 3 * i));
    if (stereo[i]) {
      stereo_own.emplace_back(new EdgeReprojectPVRStereo);
      auto& e = *stereo_own.back();
      auto& b = base_of(e);
      b._vertices[0] = &pts[i], b._vertices[1] = &vCur;
      b._jacobianOplus[0].cols = 3, b._jacobianOplus[1].cols = 9;
      e.SetParams(&c, Rcb, tcb, &bf);
      for (int k = 0; k < 3; ++k) b._measurement(k) = (double)obs[3 * i + k];
      b.setInformation(Matrix<double, 3, 3>::Identity() * w[i]);
      b.setRobustKernel(kernel(delta[i]));
      b.setLevel(level[i]);
      b.computeError();
      ster.push_back(&e);
    } else {
      mono_own.emplace_back(new EdgeReprojectPVR);
      auto& e = *mono_own.back();
      auto& b = base_of(e);
      b._vertices[0] = &pts[i], b._vertices[1] = &vCur;
      b._jacobianOplus[0].cols = 3, b._jacobianOplus[1].cols = 9;
      e.SetParams(&c, Rcb, tcb, &bf);
      for (int k = 0; k < 2; ++k) b._measurement(k) = (double)obs[3 * i + k];
      b.setInformation(Matrix<double, 2, 2>::Identity() * w[i]);
      b.setRobustKernel(kernel(delta[i]));
      b.setLevel(level[i]);
      b.computeError();
      mono.push_back(&e);
    }
  }
  typedef Matrix<double, 15, 15> Matrix15d;
  Matrix15d c0, cl, ccl;
  Optimizer::FillCovInv(eI.get(), &eB, (EdgeEncNavStatePVR*)nullptr, 0, &mono, &ster, c0, (EdgeNavStatePriorPVRBias*)nullptr, (int8_t)kExactRobust);
  to_rm(c0, C);
  if (prior) {
    Optimizer::FillCovInv(eI.get(), &eB, (EdgeEncNavStatePVR*)nullptr, 2, &mono, &ster, cl, eP.get(), (int8_t)kExactRobust);
    Optimizer::FillCovInv(eI.get(), &eB, (EdgeEncNavStatePVR*)nullptr, 1, &mono, &ster, ccl, (EdgeNavStatePriorPVRBias*)nullptr, (int8_t)kExactRobust);
    to_rm(cl, CL), to_rm(ccl, CCL);
  }
}

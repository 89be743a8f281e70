// ORACLE — TEST INFRASTRUCTURE ONLY.
// C interface over the three Project() member functions of the REFERENCE's camera models, compiled UNCHANGED:
// camm::PinholeCamera::Project (common/camera_models/camera_pinhole.h:70-106), camm::RadtanCamera::Project
// (camera_radtan.h:61-129) and camm::KB8Camera::Project (camera_kb8.h:68-157).  Their headers pull in Eigen and Sophus
// (absent here), so the recipe (Makefile) cuts exactly these three function definitions out of the headers into
// oracle/_ref/gen/camera_project_fns.inc (git-ignored build output) and this file supplies the class declarations and the few
// vector / matrix operations the bodies use (element access, setZero, resize, the comma initialiser, col() = col() * s,
// conservativeResize, block().setZero()) as a minimal stand-in for the Eigen types named in camera_base.h:60-89.
#include <math.h>
#include <stdint.h>
#include <cmath>
#include <cstring>
#include <vector>

namespace VIEO_SLAM {
namespace camm {

struct Vec3ioS {  // Eigen::Matrix<double, 3, 1>
  double v[3];
  Vec3ioS() : v{0, 0, 0} {}
  Vec3ioS(double a, double b, double c) : v{a, b, c} {}
  const double& operator[](int i) const { return v[i]; }
};
struct Vec2dataS {  // Eigen::Matrix<float, 2, 1>
  float v[2];
  float& operator[](int i) { return v[i]; }
};
struct CommaInit {
  double* p;
  CommaInit& operator,(double x) {
    *p++ = x;
    return *this;
  }
};
struct Mat23ioS {  // Eigen::Matrix<double, 2, 3>; the comma initialiser fills row by row
  double m[6];
  void setZero() { std::memset(m, 0, sizeof(m)); }
  double& operator()(int r, int c) { return m[3 * r + c]; }
  CommaInit operator<<(double x) {
    m[0] = x;
    return CommaInit{m + 1};
  }
};
struct Mat2XioS;
struct ColProd {
  const Mat2XioS* m;
  int c;
  double s;
};
struct ColRef {
  Mat2XioS* m;
  int c;
  ColProd operator*(double s) const;
  void operator=(const ColProd& p);
};
struct BlockRef {
  Mat2XioS* m;
  int r0, c0, nr, nc;
  void setZero();
};
struct Mat2XioS {  // Eigen::Matrix<double, 2, Dynamic> (column-major like Eigen, which only matters to this stand-in)
  std::vector<double> d;
  int cols = 0;
  void resize(int r, int c) {
    (void)r;
    cols = c;
    d.assign(2 * (size_t)c, 0.0);
  }
  void conservativeResize(int r, int c) {
    (void)r;
    d.resize(2 * (size_t)c, 0.0);
    cols = c;
  }
  void setZero() { std::fill(d.begin(), d.end(), 0.0); }
  double& operator()(int r, int c) { return d[2 * (size_t)c + r]; }
  ColRef col(int c) { return ColRef{this, c}; }
  BlockRef block(int r0, int c0, int nr, int nc) { return BlockRef{this, r0, c0, nr, nc}; }
};
inline ColProd ColRef::operator*(double s) const { return ColProd{m, c, s}; }
inline void ColRef::operator=(const ColProd& p) {
  for (int r = 0; r < 2; ++r) m->d[2 * (size_t)c + r] = p.m->d[2 * (size_t)p.c + r] * p.s;
}
inline void BlockRef::setZero() {
  for (int c = c0; c < c0 + nc; ++c)
    for (int r = r0; r < r0 + nr; ++r) (*m)(r, c) = 0.0;
}

class PinholeCamera {  // declaration subset of camera_pinhole.h:14-68 / camera_base.h:58-150
 public:
  using Tdata = float;   // FLT_CAMM (common/config.h:23)
  using Tcalc = double;  // FLT_CALC_CAMM (:24)
  using Vec3io = Vec3ioS;
  using Vec2data = Vec2dataS;
  using Mat23io = Mat23ioS;
  using Mat2Xio = Mat2XioS;
  std::vector<Tdata> parameters_;
  virtual ~PinholeCamera() {}
  inline virtual void Project(const Vec3io& p_3d, Vec2data* p_img, Mat23io* d_img_d_p3d = nullptr,
                              Mat2Xio* d_img_d_param = nullptr) const;
};
class RadtanCamera : public PinholeCamera {
  using Base = PinholeCamera;

 public:
  int num_k_ = 2;
  inline void Project(const Vec3io& p_3d, Vec2data* p_img, Mat23io* d_img_d_p3d = nullptr,
                      Mat2Xio* d_img_d_param = nullptr) const override;
};
class KB8Camera : public PinholeCamera {
  using Base = PinholeCamera;
  const float precision_r_ = 1e-5;  // camera_kb8.h:61

 public:
  inline void Project(const Vec3io& p_3d, Vec2data* p_img, Mat23io* d_img_d_p3d = nullptr,
                      Mat2Xio* d_img_d_param = nullptr) const override;
};

#include "camera_project_fns.inc"

}  // namespace camm
}  // namespace VIEO_SLAM

extern "C" {
// model: 0 pinhole (params fx fy cx cy), 1 radtan (fx fy cx cy k1..k_n p1 p2), 2 KB8 (fx fy cx cy k1..k4)
// uv: the float pixel; J: d(img)/d(p3d) 2x3 row-major or NULL; Jp: d(img)/d(param) 2 x n_params column-major or NULL
void ref_cam_project(int model, const float* params, int n_params, const double P[3], float uv[2], double* J, double* Jp) {
  using namespace VIEO_SLAM::camm;
  PinholeCamera pin;
  RadtanCamera rad;
  KB8Camera kb;
  PinholeCamera* cam = model == 1 ? (PinholeCamera*)&rad : model == 2 ? (PinholeCamera*)&kb : &pin;
  cam->parameters_.assign(params, params + n_params);
  if (model == 1) rad.num_k_ = n_params - 6;
  Vec3ioS p(P[0], P[1], P[2]);
  Vec2dataS img{};
  Mat23ioS j23;
  j23.setZero();
  Mat2XioS jp;
  cam->Project(p, &img, J ? &j23 : nullptr, Jp ? &jp : nullptr);
  uv[0] = img[0];
  uv[1] = img[1];
  if (J) std::memcpy(J, j23.m, sizeof(j23.m));
  if (Jp)
    for (size_t k = 0; k < jp.d.size() && k < 2 * (size_t)n_params; ++k) Jp[k] = jp.d[k];
}
}

// g2o::Sim3 arithmetic for the essential-graph optimiser (posegraph.cu): the exponential constructor, log, inverse, product
// and map of optimizer/g2o/g2o/types/sim3.h:40-270, VertexSim3Expmap::oplusImpl and EdgeSim3::computeError
// (types_seven_dof_expmap.h:30-42, 107-115), with Eigen 3.3's quaternion formulas spelled out (Quaternion(Matrix3),
// quaternion product, _transformVector, toRotationMatrix, 3x3 partial-pivot LU).  fp64 throughout; -fmad=false.
// The functions are __host__ __device__ so that tools/sim3_host_check.cc can run the very same text on the CPU against the
// oracle before GPU time is spent; the library only ever calls them from kernels.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define VIEO_HD __host__ __device__ __forceinline__
#else
#define VIEO_HD inline
#endif

namespace vieo {

struct Sim3d {   // byte-identical to VieoSim3: r = (x, y, z, w), t, s
  double q[4];
  double t[3];
  double s;
};

VIEO_HD void s3_cross(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
VIEO_HD void s3_qrot(const double q[4], const double v[3], double o[3]) {
  double uv[3], c[3];
  s3_cross(q, v, uv);
  for (int k = 0; k < 3; ++k) uv[k] += uv[k];
  s3_cross(q, uv, c);
  for (int k = 0; k < 3; ++k) o[k] = v[k] + q[3] * uv[k] + c[k];
}
VIEO_HD void s3_qmul(const double a[4], const double b[4], double o[4]) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
VIEO_HD void s3_q2R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
VIEO_HD void s3_R2q(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    const double h = 0.5 / t;
    const double qi = 0.5 * t, qw = (R[3 * k + j] - R[3 * j + k]) * h, qj = (R[3 * j + i] + R[3 * i + j]) * h,
                 qk = (R[3 * k + i] + R[3 * i + k]) * h;
    // static indexing only (a dynamically indexed local array would live in local memory on the device)
    q[0] = i == 0 ? qi : (j == 0 ? qj : qk);
    q[1] = i == 1 ? qi : (j == 1 ? qj : qk);
    q[2] = i == 2 ? qi : (j == 2 ? qj : qk);
    q[3] = qw;
  }
}
VIEO_HD void s3_skew(const double w[3], double O[9]) {
  O[0] = 0; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
VIEO_HD void s3_mm3(const double A[9], const double B[9], double C[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}

// the coefficients A, B, C of W = A Omega + B Omega^2 + C I shared by the constructor and log (sim3.h:78-130, 158-203)
VIEO_HD void s3_abc(double sigma, double s, double theta, bool small_theta, double& A, double& B, double& C) {
  const double eps = 0.00001;
  if (fabs(sigma) < eps) {
    C = 1;
    if (small_theta) {
      A = 1. / 2.;
      B = 1. / 6.;
    } else {
      const double theta2 = theta * theta;
      A = (1 - cos(theta)) / theta2;
      B = (theta - sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (s - 1) / sigma;
    if (small_theta) {
      const double sigma2 = sigma * sigma;
      A = ((sigma - 1) * s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * s - 1) / (sigma2 * sigma);
    } else {
      const double a = s * sin(theta), b = s * cos(theta);
      const double theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
}

// Sim3(const Vector7d&) (sim3.h:61-136): u = (omega, upsilon, sigma)
VIEO_HD Sim3d s3_exp(const double u[7]) {
  Sim3d o;
  const double sigma = u[6];
  const double theta = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  double Om[9], Om2[9], R[9];
  s3_skew(u, Om);
  o.s = exp(sigma);
  s3_mm3(Om, Om, Om2);
  const bool small_theta = theta < 0.00001;
  double A, B, C;
  s3_abc(sigma, o.s, theta, small_theta, A, B, C);
  if (small_theta) {
    for (int k = 0; k < 9; ++k) R[k] = ((k % 4 == 0 ? 1.0 : 0.0) + Om[k]) + Om2[k] / 2;
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta);
    for (int k = 0; k < 9; ++k) R[k] = ((k % 4 == 0 ? 1.0 : 0.0) + a * Om[k]) + b * Om2[k];
  }
  s3_R2q(R, o.q);
  for (int r = 0; r < 3; ++r) {
    double acc = 0;
    double w[3];
    for (int c = 0; c < 3; ++c) w[c] = (A * Om[3 * r + c] + B * Om2[3 * r + c]) + C * (r == c ? 1.0 : 0.0);
    acc = w[0] * u[3] + w[1] * u[4] + w[2] * u[5];
    o.t[r] = acc;
  }
  return o;
}

// 3x3 partial-pivot LU solve (Eigen PartialPivLU: largest |.| at or below the diagonal, first wins ties); rows are kept in
// named registers, swaps are value swaps
VIEO_HD void s3_lu3(const double W[9], const double t[3], double x[3]) {
  double a0[4] = {W[0], W[1], W[2], t[0]}, a1[4] = {W[3], W[4], W[5], t[1]}, a2[4] = {W[6], W[7], W[8], t[2]};
#define VIEO_SWAP4(p, q)      \
  for (int c = 0; c < 4; ++c) { \
    const double tmp = p[c];  \
    p[c] = q[c];              \
    q[c] = tmp;               \
  }
  {  // column 0
    int p = 0;
    double best = fabs(a0[0]);
    if (fabs(a1[0]) > best) { p = 1; best = fabs(a1[0]); }
    if (fabs(a2[0]) > best) p = 2;
    if (p == 1) { VIEO_SWAP4(a0, a1) }
    else if (p == 2) { VIEO_SWAP4(a0, a2) }
    const double f1 = a1[0] / a0[0], f2 = a2[0] / a0[0];
    for (int c = 1; c < 4; ++c) {
      a1[c] -= f1 * a0[c];
      a2[c] -= f2 * a0[c];
    }
  }
  {  // column 1
    if (fabs(a2[1]) > fabs(a1[1])) { VIEO_SWAP4(a1, a2) }
    const double f = a2[1] / a1[1];
    for (int c = 2; c < 4; ++c) a2[c] -= f * a1[c];
  }
#undef VIEO_SWAP4
  x[2] = a2[3] / a2[2];
  x[1] = (a1[3] - a1[2] * x[2]) / a1[1];
  x[0] = ((a0[3] - a0[1] * x[1]) - a0[2] * x[2]) / a0[0];
}

// Sim3::log (sim3.h:143-216)
VIEO_HD void s3_log(const Sim3d& S, double res[7]) {
  const double sigma = log(S.s);
  double R[9];
  s3_q2R(S.q, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  const bool small_theta = d > 1 - 0.00001;
  double omega[3], theta = 0;
  if (small_theta) {
    for (int k = 0; k < 3; ++k) omega[k] = 0.5 * dR[k];
  } else {
    theta = acos(d);
    const double f = theta / (2 * sqrt(1 - d * d));
    for (int k = 0; k < 3; ++k) omega[k] = f * dR[k];
  }
  double A, B, C;
  s3_abc(sigma, S.s, theta, small_theta, A, B, C);
  double Om[9], Om2[9], W[9], ups[3];
  s3_skew(omega, Om);
  s3_mm3(Om, Om, Om2);
  for (int k = 0; k < 9; ++k) W[k] = (A * Om[k] + B * Om2[k]) + C * (k % 4 == 0 ? 1.0 : 0.0);
  s3_lu3(W, S.t, ups);
  for (int k = 0; k < 3; ++k) {
    res[k] = omega[k];
    res[3 + k] = ups[k];
  }
  res[6] = sigma;
}

VIEO_HD Sim3d s3_inv(const Sim3d& S) {  // (sim3.h:218-221)
  Sim3d o;
  o.q[0] = -S.q[0]; o.q[1] = -S.q[1]; o.q[2] = -S.q[2]; o.q[3] = S.q[3];
  const double f = -1. / S.s;
  const double v[3] = {f * S.t[0], f * S.t[1], f * S.t[2]};
  s3_qrot(o.q, v, o.t);
  o.s = 1. / S.s;
  return o;
}
VIEO_HD Sim3d s3_mul(const Sim3d& a, const Sim3d& b) {  // (sim3.h:245-251)
  Sim3d o;
  s3_qmul(a.q, b.q, o.q);
  double rt[3];
  s3_qrot(a.q, b.t, rt);
  for (int k = 0; k < 3; ++k) o.t[k] = a.s * rt[k] + a.t[k];
  o.s = a.s * b.s;
  return o;
}
VIEO_HD void s3_map(const Sim3d& S, const double x[3], double o[3]) {  // (sim3.h:138-140)
  double r[3];
  s3_qrot(S.q, x, r);
  for (int k = 0; k < 3; ++k) o[k] = S.s * r[k] + S.t[k];
}
// EdgeSim3::computeError: log(meas * v0 * v1^-1)
VIEO_HD void s3_edge_error(const Sim3d& meas, const Sim3d& v0, const Sim3d& v1, double e[7]) {
  s3_log(s3_mul(s3_mul(meas, v0), s3_inv(v1)), e);
}
// VertexSim3Expmap::oplusImpl with a single non-zero component d of value h (the numeric Jacobian's perturbation), or the
// full update
VIEO_HD Sim3d s3_oplus(const Sim3d& est, const double upd[7], bool fix_scale) {
  double u[7];
  for (int k = 0; k < 7; ++k) u[k] = upd[k];
  if (fix_scale) u[6] = 0;
  return s3_mul(s3_exp(u), est);
}

}  // namespace vieo

// Device-side fp64 3x3 / SO(3) helpers shared by the IMU pre-integrator and the BA kernels.
// Semantics follow common/so3_extra.h:121-288 of the reference (quaternion exp/log with 1e-5 small-angle
// branches, Jr, Jr^-1, normalizeRotationM) and Eigen's quaternion<->matrix formulas.  The library is compiled
// with -fmad=false, so every expression rounds exactly as written.
#pragma once
#include <math.h>

namespace vieo {

constexpr double kSo3Eps = 1e-5;

struct Mat3 {
  double m[9];  // row-major
  __device__ __forceinline__ double& operator()(int r, int c) { return m[3 * r + c]; }
  __device__ __forceinline__ double operator()(int r, int c) const { return m[3 * r + c]; }
};
struct Vec3 {
  double x, y, z;
};

__device__ __forceinline__ Mat3 m3_identity() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
__device__ __forceinline__ Mat3 m3_zero() { return {{0, 0, 0, 0, 0, 0, 0, 0, 0}}; }
__device__ __forceinline__ Mat3 m3_mul(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += a.m[3 * i + k] * b.m[3 * k + j];
      r.m[3 * i + j] = s;
    }
  return r;
}
__device__ __forceinline__ Mat3 m3_t(const Mat3& a) {
  return {{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}};
}
__device__ __forceinline__ Mat3 m3_scale(const Mat3& a, double s) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s;
  return r;
}
__device__ __forceinline__ Mat3 m3_add(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i];
  return r;
}
__device__ __forceinline__ Mat3 m3_sub(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i];
  return r;
}
__device__ __forceinline__ Mat3 m3_hat(const Vec3& w) { return {{0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0}}; }
__device__ __forceinline__ Vec3 m3_mulv(const Mat3& a, const Vec3& v) {
  return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
          a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
__device__ __forceinline__ Vec3 m3_tmulv(const Mat3& a, const Vec3& v) {  // a^T v
  return {a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
          a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}
__device__ __forceinline__ Vec3 v3_add(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec3 v3_sub(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 v3_scale(const Vec3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double v3_norm(const Vec3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

struct Quat {
  double w, x, y, z;
};
__device__ __forceinline__ Quat q_normalized(Quat q) {
  const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (n2 > 0) {
    const double n = sqrt(n2);
    q.w /= n;
    q.x /= n;
    q.y /= n;
    q.z /= n;
  }
  return q;
}
__device__ __forceinline__ Mat3 q_matrix(const Quat& q) {  // Eigen::Quaternion::toRotationMatrix
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  return {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
           1 - (txx + tyy)}};
}
__device__ __forceinline__ Quat q_from_matrix(const Mat3& R) {  // Eigen::Quaternion(Matrix3)
  double c[4];
  double t = R(0, 0) + R(1, 1) + R(2, 2);
  if (t > 0) {
    t = sqrt(t + 1.0);
    c[3] = 0.5 * t;
    t = 0.5 / t;
    c[0] = (R(2, 1) - R(1, 2)) * t;
    c[1] = (R(0, 2) - R(2, 0)) * t;
    c[2] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    c[i] = 0.5 * t;
    t = 0.5 / t;
    c[3] = (R(k, j) - R(j, k)) * t;
    c[j] = (R(j, i) + R(i, j)) * t;
    c[k] = (R(k, i) + R(i, k)) * t;
  }
  return {c[3], c[0], c[1], c[2]};
}
__device__ __forceinline__ Quat q_mul(const Quat& a, const Quat& b) {  // Eigen quaternion product
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
// SO3ex::exp -> normalised quaternion (so3_extra.h:121-142 + the normalising constructor :67-75)
__device__ __forceinline__ Quat so3_exp_q(const Vec3& w) {
  const double theta = v3_norm(w);
  double imag, real;
  if (theta < kSo3Eps) {
    const double t2 = theta * theta;
    imag = 0.5 - t2 / 48.;
    real = 1.0 - t2 / 8.;
  } else {
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  return q_normalized({real, imag * w.x, imag * w.y, imag * w.z});
}
__device__ __forceinline__ Mat3 so3_Exp(const Vec3& w) { return q_matrix(so3_exp_q(w)); }
// SO3ex::log (so3_extra.h:152-190)
__device__ __forceinline__ Vec3 so3_log_q(const Quat& q) {
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z), w = q.w, sw = w * w;
  double f;
  if (n < kSo3Eps) {
    f = 2. / w - 2. / 3 * (n * n) / (w * sw);
  } else if (fabs(w) < kSo3Eps) {
    f = (w > 0 ? M_PI : -M_PI) / n;
    const double n2 = n * n, n4 = n2 * n2;
    f -= 2 * w / n2 - 2. / 3 * (w * sw) / n4;
  } else {
    f = 2 * atan(n / w) / n;
  }
  return {f * q.x, f * q.y, f * q.z};
}
// SO3ex(R).log(): matrix -> quaternion -> normalise -> log (so3_extra.h:59-64, 299-302)
__device__ __forceinline__ Vec3 so3_Log(const Mat3& R) { return so3_log_q(q_normalized(q_from_matrix(R))); }
// SO3ex::JacobianR (so3_extra.h:255-270)
__device__ __forceinline__ Mat3 so3_Jr(const Vec3& w) {
  const double theta = v3_norm(w);
  if (theta < kSo3Eps) {
    const Mat3 O = m3_hat(w), O2 = m3_mul(O, O);
    return m3_add(m3_sub(m3_identity(), m3_scale(O, 0.5)), m3_scale(O2, 1. / 6.));
  }
  const Mat3 K = m3_hat({w.x / theta, w.y / theta, w.z / theta});
  return m3_add(m3_sub(m3_identity(), m3_scale(K, (1 - cos(theta)) / theta)),
                m3_scale(m3_mul(K, K), 1 - sin(theta) / theta));
}
// SO3ex::JacobianRInv (so3_extra.h:271-288)
__device__ __forceinline__ Mat3 so3_JrInv(const Vec3& w) {
  const double theta = v3_norm(w);
  const Mat3 O = m3_hat(w);
  if (theta < kSo3Eps) return m3_add(m3_add(m3_identity(), m3_scale(O, 0.5)), m3_scale(m3_mul(O, O), 1. / 12.));
  const Mat3 K = m3_hat({w.x / theta, w.y / theta, w.z / theta});
  return m3_add(m3_add(m3_identity(), m3_scale(O, 0.5)),
                m3_scale(m3_mul(K, K), 1.0 - (1.0 + cos(theta)) * theta / (2.0 * sin(theta))));
}
// SO3ex::normalizeRotationM (so3_extra.h:218-229)
__device__ __forceinline__ Mat3 so3_normalize(const Mat3& R) {
  Quat q = q_from_matrix(R);
  if (q.w < 0) {
    q.w = -q.w;
    q.x = -q.x;
    q.y = -q.y;
    q.z = -q.z;
  }
  return q_matrix(q_normalized(q));
}

}  // namespace vieo

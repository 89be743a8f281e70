// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  Fixed-size fp64 algebra and the SO(3) helpers of
// common/so3_extra.h:121-288 (quaternion exp / log, Jr, Jr^-1, normalizeRotationM) shared by imu_oracle.cc and
// ba_oracle.cc.  Eigen / Sophus are absent here ("parity unpinned" for their last-bit rounding): quaternion<->matrix
// conversions follow Eigen 3.3.7's published formulas.  common/so3_extra.h itself, compiled unchanged against a stand-in for the
// two libraries, agrees with these helpers to 1 ulp (oracle/_ref, tests/test_oracle_ref.py::test_so3_helpers_equal_reference).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {


constexpr double kEps = 1e-5;  // SO3ex::SMALL_EPS

struct M3 {
  double m[9];  // row-major
};
inline M3 ident() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a.m[3 * i + k] * b.m[3 * k + j];
      r.m[3 * i + j] = s;
    }
  return r;
}
inline M3 tr(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * j + i];
  return r;
}
inline M3 scale(const M3& a, double s) {
  M3 r;
  for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s;
  return r;
}
inline M3 add(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i];
  return r;
}
inline M3 sub(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i];
  return r;
}
inline M3 hat(const double w[3]) { return {{0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}}; }
inline void mulv(const M3& a, const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = a.m[3 * i] * v[0] + a.m[3 * i + 1] * v[1] + a.m[3 * i + 2] * v[2];
}

struct Quat {
  double w, x, y, z;
};
inline Quat qnormalized(Quat q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (n2 > 0) {
    double n = std::sqrt(n2);
    q.w /= n; q.x /= n; q.y /= n; q.z /= n;
  }
  return q;
}
// Eigen::Quaternion::toRotationMatrix
inline M3 qmat(const Quat& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  return {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
           1 - (txx + tyy)}};
}
// Eigen::Quaternion(Matrix3)
inline Quat mquat(const M3& R) {
  auto m = [&](int i, int j) { return R.m[3 * i + j]; };
  double c[4];  // x y z w
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    c[3] = 0.5 * t;
    t = 0.5 / t;
    c[0] = (m(2, 1) - m(1, 2)) * t;
    c[1] = (m(0, 2) - m(2, 0)) * t;
    c[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    c[i] = 0.5 * t;
    t = 0.5 / t;
    c[3] = (m(k, j) - m(j, k)) * t;
    c[j] = (m(j, i) + m(i, j)) * t;
    c[k] = (m(k, i) + m(i, k)) * t;
  }
  return {c[3], c[0], c[1], c[2]};
}
// SO3ex::exp (so3_extra.h:121-142) followed by the normalising constructor and .matrix()
inline M3 so3_Exp(const double w[3]) {
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double imag, real;
  if (theta < kEps) {
    const double t2 = theta * theta;
    imag = 0.5 - t2 / 48.;
    real = 1.0 - t2 / 8.;
  } else {
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  return qmat(qnormalized({real, imag * w[0], imag * w[1], imag * w[2]}));
}
// SO3ex::JacobianR (so3_extra.h:255-270)
inline M3 so3_Jr(const double w[3]) {
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (theta < kEps) {
    M3 O = hat(w), O2 = mul(O, O);
    return add(sub(ident(), scale(O, 0.5)), scale(O2, 1. / 6.));
  }
  const double k[3] = {w[0] / theta, w[1] / theta, w[2] / theta};
  M3 K = hat(k);
  return add(sub(ident(), scale(K, (1 - std::cos(theta)) / theta)), scale(mul(K, K), 1 - std::sin(theta) / theta));
}
// SO3ex::normalizeRotationM (so3_extra.h:218-229)
inline M3 normalize_rot(const M3& R) {
  Quat q = mquat(R);
  if (q.w < 0) {
    q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z;
  }
  return qmat(qnormalized(q));
}


// quaternion product (Eigen::Quaternion::operator*)
inline Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat qconj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }
// SO3ex::exp as a unit quaternion (so3_extra.h:121-142)
inline Quat so3_exp_q(const double w[3]) {
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double imag, real;
  if (theta < kEps) {
    const double t2 = theta * theta;
    imag = 0.5 - t2 / 48.;
    real = 1.0 - t2 / 8.;
  } else {
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  return qnormalized({real, imag * w[0], imag * w[1], imag * w[2]});
}
// SO3ex::log (so3_extra.h:152-190): angle in [-pi, pi)
inline void so3_log_q(const Quat& q, double out[3]) {
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  const double w = q.w, w2 = w * w;
  double f;
  if (n < kEps) {
    f = 2. / w - 2. / 3 * (n * n) / (w * w2);
  } else if (std::fabs(w) < kEps) {
    f = (w > 0 ? M_PI : -M_PI) / n;
    const double n2 = n * n, n4 = n2 * n2;
    f -= 2 * w / n2 - 2. / 3 * (w * w2) / n4;
  } else {
    f = 2 * std::atan(n / w) / n;
  }
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}
// SO3ex::JacobianRInv (so3_extra.h:271-288)
inline M3 so3_JrInv(const double w[3]) {
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const M3 O = hat(w);
  if (theta < kEps) return add(add(ident(), scale(O, 0.5)), scale(mul(O, O), 1. / 12.));
  const double k[3] = {w[0] / theta, w[1] / theta, w[2] / theta};
  const M3 K = hat(k);
  return add(add(ident(), scale(O, 0.5)),
             scale(mul(K, K), 1.0 - (1.0 + std::cos(theta)) * theta / (2.0 * std::sin(theta))));
}
}  // namespace orc

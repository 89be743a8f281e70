// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  CPU restatement of Optimizer::OptimizeEssentialGraph
// (src/Optimizer.cc:2309-2688): g2o::Sim3 (optimizer/g2o/g2o/types/sim3.h:40-270), VertexSim3Expmap::oplusImpl and EdgeSim3
// (types_seven_dof_expmap.h:30-129; the edge has NO analytic Jacobian, BaseBinaryEdge::linearizeOplus differentiates it
// numerically with delta = 1e-9, core/base_binary_edge.hpp:123-187), the quadratic form of a binary edge without a robust
// kernel (base_binary_edge.hpp:60-115), Levenberg-Marquardt with setUserLambdaInit(1e-16) over optimize(20)
// (optimization_algorithm_levenberg.cpp:61-189, sparse_optimizer.cpp:354-419), the SE3 recovery and the map-point correction
// (src/Optimizer.cc:2618-2680).  The Eigen pieces are restated from Eigen 3.3's published formulas: Quaternion(Matrix3)
// (Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>), quaternion product, _transformVector, toRotationMatrix,
// PartialPivLU 3x3.  The reference solves the pose system with SimplicialLDLT (linear_solver_eigen.h); here it is the skyline
// Cholesky of ba_oracle.cc — tolerance parity, like F5.
// Parity pins (tests/test_oracle_posegraph.py): exp / log against scipy.linalg.expm of the 4x4 similarity generator, the edge
// error against numpy matrix algebra, the numeric Jacobian against an independent central difference, one damped step
// against numpy's dense normal equations, drift recovery on a consistent loop.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#include "ba_oracle.h"

bool orc_chol_solve_skyline(std::vector<double>& A, int n, const double* b, double* x);

namespace {

struct S3 {  // g2o::Sim3: r (Eigen::Quaterniond, coeffs x y z w), t, s
  double q[4];
  double t[3];
  double s;
};

inline void cross(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// Eigen QuaternionBase::_transformVector: uv = vec x v; uv += uv; v + w uv + vec x uv
inline void qrot(const double q[4], const double v[3], double o[3]) {
  double uv[3], c[3];
  cross(q, v, uv);
  for (int k = 0; k < 3; ++k) uv[k] += uv[k];
  cross(q, uv, c);
  for (int k = 0; k < 3; ++k) o[k] = v[k] + q[3] * uv[k] + c[k];
}
// Eigen quat_product
inline void qmul(const double a[4], const double b[4], double o[4]) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
// Eigen QuaternionBase::toRotationMatrix (row-major R[3*r + c])
inline void q2R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen quaternionbase_assign_impl<Matrix3, 3, 3>
inline void R2q(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
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
    t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}
inline void skew(const double w[3], double O[9]) {
  O[0] = 0; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
inline void mm3(const double A[9], const double B[9], double C[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}

// Sim3(const Vector7d& update) (sim3.h:61-136): update = (omega, upsilon, sigma)
S3 s3_exp(const double u[7]) {
  S3 o;
  const double* omega = u;
  const double* ups = u + 3;
  const double sigma = u[6];
  const double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double Om[9], Om2[9], R[9];
  skew(omega, Om);
  o.s = std::exp(sigma);
  mm3(Om, Om, Om2);
  const double eps = 0.00001;
  double A, B, C;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (theta < eps) {
      A = 1. / 2.;
      B = 1. / 6.;
      for (int k = 0; k < 9; ++k) R[k] = (I[k] + Om[k]) + Om2[k] / 2;
    } else {
      const double theta2 = theta * theta;
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
      const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
      for (int k = 0; k < 9; ++k) R[k] = (I[k] + a * Om[k]) + b * Om2[k];
    }
  } else {
    C = (o.s - 1) / sigma;
    if (theta < eps) {
      const double sigma2 = sigma * sigma;
      A = ((sigma - 1) * o.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * o.s - 1) / (sigma2 * sigma);
      for (int k = 0; k < 9; ++k) R[k] = (I[k] + Om[k]) + Om2[k] / 2;
    } else {
      const double sa = std::sin(theta) / theta, sb = (1 - std::cos(theta)) / (theta * theta);
      for (int k = 0; k < 9; ++k) R[k] = (I[k] + sa * Om[k]) + sb * Om2[k];
      const double a = o.s * std::sin(theta), b = o.s * std::cos(theta);
      const double theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
  R2q(R, o.q);
  double W[9];
  for (int k = 0; k < 9; ++k) W[k] = (A * Om[k] + B * Om2[k]) + C * I[k];
  for (int r = 0; r < 3; ++r) o.t[r] = W[3 * r] * ups[0] + W[3 * r + 1] * ups[1] + W[3 * r + 2] * ups[2];
  return o;
}

// PartialPivLU of a 3x3 + solve (Eigen: row of the largest |.| in the column at or below the diagonal, first wins ties)
void lu3_solve(const double W[9], const double t[3], double x[3]) {
  double a[3][4];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a[r][c] = W[3 * r + c];
    a[r][3] = t[r];
  }
  for (int k = 0; k < 3; ++k) {
    int p = k;
    for (int r = k + 1; r < 3; ++r)
      if (std::fabs(a[r][k]) > std::fabs(a[p][k])) p = r;
    if (p != k)
      for (int c = 0; c < 4; ++c) std::swap(a[p][c], a[k][c]);
    for (int r = k + 1; r < 3; ++r) {
      const double f = a[r][k] / a[k][k];
      for (int c = k + 1; c < 4; ++c) a[r][c] -= f * a[k][c];
    }
  }
  for (int r = 2; r >= 0; --r) {
    double s = a[r][3];
    for (int c = r + 1; c < 3; ++c) s -= a[r][c] * x[c];
    x[r] = s / a[r][r];
  }
}

// Sim3::log (sim3.h:143-216)
void s3_log(const S3& S, double res[7]) {
  const double sigma = std::log(S.s);
  double R[9];
  q2R(S.q, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};  // deltaR
  double omega[3], Om[9], Om2[9];
  const double eps = 0.00001;
  double A, B, C;
  if (std::fabs(sigma) < eps) {
    C = 1;
    if (d > 1 - eps) {
      for (int k = 0; k < 3; ++k) omega[k] = 0.5 * dR[k];
      A = 1. / 2.;
      B = 1. / 6.;
    } else {
      const double theta = std::acos(d), theta2 = theta * theta;
      const double f = theta / (2 * std::sqrt(1 - d * d));
      for (int k = 0; k < 3; ++k) omega[k] = f * dR[k];
      A = (1 - std::cos(theta)) / theta2;
      B = (theta - std::sin(theta)) / (theta2 * theta);
    }
  } else {
    C = (S.s - 1) / sigma;
    if (d > 1 - eps) {
      const double sigma2 = sigma * sigma;
      for (int k = 0; k < 3; ++k) omega[k] = 0.5 * dR[k];
      A = ((sigma - 1) * S.s + 1) / sigma2;
      B = ((0.5 * sigma2 - sigma + 1) * S.s - 1) / (sigma2 * sigma);
    } else {
      const double theta = std::acos(d);
      const double f = theta / (2 * std::sqrt(1 - d * d));
      for (int k = 0; k < 3; ++k) omega[k] = f * dR[k];
      const double theta2 = theta * theta;
      const double a = S.s * std::sin(theta), b = S.s * std::cos(theta), c = theta2 + sigma * sigma;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
    }
  }
  skew(omega, Om);
  mm3(Om, Om, Om2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double W[9], ups[3];
  for (int k = 0; k < 9; ++k) W[k] = (A * Om[k] + B * Om2[k]) + C * I[k];
  lu3_solve(W, S.t, ups);
  for (int k = 0; k < 3; ++k) {
    res[k] = omega[k];
    res[3 + k] = ups[k];
  }
  res[6] = sigma;
}

// Sim3::inverse (sim3.h:218-221): (r*, r* ((-1/s) t), 1/s)
S3 s3_inv(const S3& S) {
  S3 o;
  o.q[0] = -S.q[0]; o.q[1] = -S.q[1]; o.q[2] = -S.q[2]; o.q[3] = S.q[3];
  const double f = -1. / S.s;
  const double v[3] = {f * S.t[0], f * S.t[1], f * S.t[2]};
  qrot(o.q, v, o.t);
  o.s = 1. / S.s;
  return o;
}
// Sim3::operator* (sim3.h:245-251)
S3 s3_mul(const S3& a, const S3& b) {
  S3 o;
  qmul(a.q, b.q, o.q);
  double rt[3];
  qrot(a.q, b.t, rt);
  for (int k = 0; k < 3; ++k) o.t[k] = a.s * rt[k] + a.t[k];
  o.s = a.s * b.s;
  return o;
}
// Sim3::map (sim3.h:138-140)
void s3_map(const S3& S, const double x[3], double o[3]) {
  double r[3];
  qrot(S.q, x, r);
  for (int k = 0; k < 3; ++k) o[k] = S.s * r[k] + S.t[k];
}
// EdgeSim3::computeError (types_seven_dof_expmap.h:107-115): log(meas * v0 * v1^-1)
void edge_error(const S3& meas, const S3& v0, const S3& v1, double e[7]) { s3_log(s3_mul(s3_mul(meas, v0), s3_inv(v1)), e); }
// VertexSim3Expmap::oplusImpl (types_seven_dof_expmap.h:30-42)
S3 oplus(const S3& est, const double upd[7], bool fix_scale) {
  double u[7];
  memcpy(u, upd, sizeof(u));
  if (fix_scale) u[6] = 0;
  return s3_mul(s3_exp(u), est);
}
// BaseBinaryEdge::linearizeOplus (base_binary_edge.hpp:123-187): central differences, delta 1e-9, one vertex at a time;
// a fixed vertex's Jacobian is not computed (left zero here, never used)
void edge_jacobians(const S3& meas, const S3& v0, const S3& v1, bool fix0, bool fix1, bool fix_scale, double Ji[49], double Jj[49]) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  memset(Ji, 0, 49 * sizeof(double));
  memset(Jj, 0, 49 * sizeof(double));
  for (int side = 0; side < 2; ++side) {
    if (side == 0 ? fix0 : fix1) continue;
    double* J = side == 0 ? Ji : Jj;
    for (int d = 0; d < 7; ++d) {
      double add[7] = {0, 0, 0, 0, 0, 0, 0}, e1[7], e2[7];
      add[d] = delta;
      if (side == 0) edge_error(meas, oplus(v0, add, fix_scale), v1, e1);
      else edge_error(meas, v0, oplus(v1, add, fix_scale), e1);
      add[d] = -delta;
      if (side == 0) edge_error(meas, oplus(v0, add, fix_scale), v1, e2);
      else edge_error(meas, v0, oplus(v1, add, fix_scale), e2);
      for (int r = 0; r < 7; ++r) J[7 * r + d] = scalar * (e1[r] - e2[r]);
    }
  }
}

inline S3 from_c(const OrcSim3& c) {
  S3 s;
  memcpy(s.q, c.q, sizeof(s.q));
  memcpy(s.t, c.t, sizeof(s.t));
  s.s = c.s;
  return s;
}
inline void to_c(const S3& s, OrcSim3* c) {
  memcpy(c->q, s.q, sizeof(s.q));
  memcpy(c->t, s.t, sizeof(s.t));
  c->s = s.s;
}

}  // namespace

extern "C" {

void orc_sim3_exp(const double u[7], OrcSim3* out) { to_c(s3_exp(u), out); }
void orc_sim3_log(const OrcSim3* S, double out[7]) { s3_log(from_c(*S), out); }
void orc_sim3_mul(const OrcSim3* a, const OrcSim3* b, OrcSim3* out) { to_c(s3_mul(from_c(*a), from_c(*b)), out); }
void orc_sim3_inv(const OrcSim3* a, OrcSim3* out) { to_c(s3_inv(from_c(*a)), out); }
void orc_sim3_from_Rt(const double R[9], const double t[3], double s, OrcSim3* out) {  // Sim3(Matrix3d, Vector3d, double) (sim3.h:59)
  S3 o;
  R2q(R, o.q);
  memcpy(o.t, t, sizeof(o.t));
  o.s = s;
  to_c(o, out);
}
void orc_edge_sim3_graph(const OrcSim3* meas, const OrcSim3* v0, const OrcSim3* v1, int fix0, int fix1, int fix_scale, double e[7],
                         double Ji[49], double Jj[49]) {
  edge_error(from_c(*meas), from_c(*v0), from_c(*v1), e);
  if (Ji && Jj) edge_jacobians(from_c(*meas), from_c(*v0), from_c(*v1), fix0 != 0, fix1 != 0, fix_scale != 0, Ji, Jj);
}

// The optimisation of OptimizeEssentialGraph once the graph is collected: vertices Scw[K] (fixed[k] != 0: the loop keyframe),
// edges (ei = vertex 0, ej = vertex 1, meas = Sji, info = 7x7 row-major or NULL for identity), optimize(iterations) with
// the user's initial lambda (<= 0: g2o's 1e-5 max diag).  single_step != 0: build the system at the initial estimate, apply ONE
// damped step with lambda_init and return (test hook; H / b copied out when given).  stats = {chi2 before, chi2 after,
// iterations run, final lambda, LM trials}.
static int essential_graph_impl(int K, const OrcSim3* Scw, const uint8_t* fixed, int fix_scale, int E, const int32_t* ei,
                                const int32_t* ej, const OrcSim3* meas_c, const double* info, int iterations, double lambda_init,
                                int single_step, OrcLmDriver lm, OrcSim3* out, double* stats, double* H_out, double* b_out) {
  std::vector<S3> est(K), meas(E);
  for (int k = 0; k < K; ++k) est[k] = from_c(Scw[k]);
  for (int e = 0; e < E; ++e) meas[e] = from_c(meas_c[e]);
  // active set (sparse_optimizer.cpp:199-267): edges with a free vertex; index mapping over the free vertices that have one
  std::vector<uint8_t> act(E, 0);
  std::vector<int> hidx(K, -1);
  std::vector<uint8_t> touched(K, 0);
  for (int e = 0; e < E; ++e) {
    if (ei[e] < 0 || ei[e] >= K || ej[e] < 0 || ej[e] >= K) return -1;
    act[e] = !(fixed[ei[e]] && fixed[ej[e]]);
    if (act[e]) touched[ei[e]] = touched[ej[e]] = 1;
  }
  int nv = 0;
  for (int k = 0; k < K; ++k)
    if (touched[k] && !fixed[k]) hidx[k] = nv++;
  const int n = 7 * nv;
  for (int k = 0; k < K; ++k) to_c(est[k], out + k);
  if (stats) std::fill(stats, stats + 5, 0.0);
  if (n == 0) return 0;
  const bool fs = fix_scale != 0;
  std::vector<double> err(7 * (size_t)E), H((size_t)n * n), b(n), x(n, 0.0);
  const double I7[49] = {1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1,
                         0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1};
  auto omega = [&](int e) { return info ? info + 49 * (size_t)e : I7; };
  auto errors = [&]() {  // computeActiveErrors + activeRobustChi2 (no kernel: chi2 = e' Omega e)
    double tot = 0;
    for (int e = 0; e < E; ++e) {
      if (!act[e]) continue;
      double* r = &err[7 * (size_t)e];
      edge_error(meas[e], est[ei[e]], est[ej[e]], r);
      const double* O = omega(e);
      double c = 0;
      for (int a = 0; a < 7; ++a) {
        double s = 0;
        for (int k = 0; k < 7; ++k) s += O[7 * a + k] * r[k];
        c += r[a] * s;
      }
      tot += c;
    }
    return tot;
  };
  auto build = [&]() {  // linearizeOplus + constructQuadraticForm (base_binary_edge.hpp:60-85), errors current
    std::fill(H.begin(), H.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    for (int e = 0; e < E; ++e) {
      if (!act[e]) continue;
      const int vi = ei[e], vj = ej[e], hi = hidx[vi], hj = hidx[vj];
      double Ji[49], Jj[49];
      edge_jacobians(meas[e], est[vi], est[vj], hi < 0, hj < 0, fs, Ji, Jj);
      const double* O = omega(e);
      const double* r = &err[7 * (size_t)e];
      double orr[7];  // omega_r = -Omega e
      for (int a = 0; a < 7; ++a) {
        double s = 0;
        for (int k = 0; k < 7; ++k) s += O[7 * a + k] * r[k];
        orr[a] = -s;
      }
      auto AtO = [&](const double* J, double* out49) {  // J' Omega
        for (int a = 0; a < 7; ++a)
          for (int c = 0; c < 7; ++c) {
            double s = 0;
            for (int k = 0; k < 7; ++k) s += J[7 * k + a] * O[7 * k + c];
            out49[7 * a + c] = s;
          }
      };
      auto add_block = [&](int r0, int c0, const double* L, const double* J) {  // H[r0.., c0..] += L J
        for (int a = 0; a < 7; ++a)
          for (int c = 0; c < 7; ++c) {
            double s = 0;
            for (int k = 0; k < 7; ++k) s += L[7 * a + k] * J[7 * k + c];
            H[(size_t)(r0 + a) * n + c0 + c] += s;
          }
      };
      double AO[49], BO[49];
      if (hi >= 0) {
        AtO(Ji, AO);
        for (int a = 0; a < 7; ++a) {
          double s = 0;
          for (int k = 0; k < 7; ++k) s += Ji[7 * k + a] * orr[k];
          b[7 * hi + a] += s;
        }
        add_block(7 * hi, 7 * hi, AO, Ji);
        if (hj >= 0) {  // off-diagonal block, mirrored into the full storage
          double blk[49];
          for (int a = 0; a < 7; ++a)
            for (int c = 0; c < 7; ++c) {
              double s = 0;
              for (int k = 0; k < 7; ++k) s += AO[7 * a + k] * Jj[7 * k + c];
              blk[7 * a + c] = s;
            }
          for (int a = 0; a < 7; ++a)
            for (int c = 0; c < 7; ++c) {
              if (hi == hj) continue;
              H[(size_t)(7 * hi + a) * n + 7 * hj + c] += blk[7 * a + c];
              H[(size_t)(7 * hj + c) * n + 7 * hi + a] += blk[7 * a + c];
            }
        }
      }
      if (hj >= 0) {
        AtO(Jj, BO);
        for (int a = 0; a < 7; ++a) {
          double s = 0;
          for (int k = 0; k < 7; ++k) s += Jj[7 * k + a] * orr[k];
          b[7 * hj + a] += s;
        }
        add_block(7 * hj, 7 * hj, BO, Jj);
      }
    }
  };
  auto update = [&]() {
    for (int k = 0; k < K; ++k)
      if (hidx[k] >= 0) est[k] = oplus(est[k], &x[7 * hidx[k]], fs);
  };
  double chi_first = 0, chi_last = 0;
  if (single_step) {
    chi_first = errors();
    build();
    if (H_out) memcpy(H_out, H.data(), sizeof(double) * (size_t)n * n);
    if (b_out) memcpy(b_out, b.data(), sizeof(double) * n);
    std::vector<double> A(H);
    for (int j = 0; j < n; ++j) A[(size_t)j * n + j] += lambda_init;
    const bool ok2 = orc_chol_solve_skyline(A, n, b.data(), x.data());
    if (ok2) update();
    chi_last = errors();
    for (int k = 0; k < K; ++k) to_c(est[k], out + k);
    if (stats) {
      stats[0] = chi_first; stats[1] = chi_last; stats[2] = 1; stats[3] = lambda_init; stats[4] = 1;
    }
    return ok2 ? n : -2;
  }
  // optimize(iterations): g2o's Levenberg-Marquardt through the callback driver (lm_oracle.cc, or the reference's own compiled
  // solve() when the caller passes oracle/_ref's driver)
  struct Ctx {
    std::function<double()> errors;
    std::function<void()> build, update;
    std::function<bool(double)> solve;
    std::vector<std::vector<S3>> stack;
    std::vector<S3>* est;
    std::vector<double>*H, *b, *x;
    int n;
  } ctx;
  ctx.errors = errors;
  ctx.build = build;
  ctx.update = update;
  ctx.solve = [&](double lambda) {
    std::vector<double> A(H);
    for (int j = 0; j < n; ++j) A[(size_t)j * n + j] += lambda;
    return orc_chol_solve_skyline(A, n, b.data(), x.data());
  };
  ctx.est = &est; ctx.H = &H; ctx.b = &b; ctx.x = &x; ctx.n = n;
  OrcLmCallbacks cb;
  cb.ctx = &ctx;
  cb.n = n;
  cb.errors = [](void* c) { return ((Ctx*)c)->errors(); };
  cb.build = [](void* c) { ((Ctx*)c)->build(); };
  cb.solve = [](void* c, double lambda) { return ((Ctx*)c)->solve(lambda) ? 1 : 0; };
  cb.update = [](void* c) { ((Ctx*)c)->update(); };
  cb.push = [](void* c) { ((Ctx*)c)->stack.push_back(*((Ctx*)c)->est); };
  cb.pop = [](void* c) { *((Ctx*)c)->est = ((Ctx*)c)->stack.back(); ((Ctx*)c)->stack.pop_back(); };
  cb.discard_top = [](void* c) { ((Ctx*)c)->stack.pop_back(); };
  cb.x = [](void* c) { return (const double*)((Ctx*)c)->x->data(); };
  cb.b = [](void* c) { return (const double*)((Ctx*)c)->b->data(); };
  cb.hessian_diag = [](void* c, int j) { return (*((Ctx*)c)->H)[(size_t)j * ((Ctx*)c)->n + j]; };
  cb.terminate = nullptr;
  double st5[5] = {0, 0, 0, 0, 0};
  (lm ? lm : orc_lm_optimize)(&cb, iterations, lambda_init, st5);
  for (int k = 0; k < K; ++k) to_c(est[k], out + k);
  if (stats) std::copy(st5, st5 + 5, stats);
  return n;
}

int orc_essential_graph(int K, const OrcSim3* Scw, const uint8_t* fixed, int fix_scale, int E, const int32_t* ei, const int32_t* ej,
                        const OrcSim3* meas_c, const double* info, int iterations, double lambda_init, int single_step, OrcSim3* out,
                        double* stats, double* H_out, double* b_out) {
  return essential_graph_impl(K, Scw, fixed, fix_scale, E, ei, ej, meas_c, info, iterations, lambda_init, single_step, nullptr, out,
                              stats, H_out, b_out);
}
int orc_essential_graph_lm(int K, const OrcSim3* Scw, const uint8_t* fixed, int fix_scale, int E, const int32_t* ei, const int32_t* ej,
                           const OrcSim3* meas_c, const double* info, int iterations, double lambda_init, OrcLmDriver lm, OrcSim3* out,
                           double* stats) {
  return essential_graph_impl(K, Scw, fixed, fix_scale, E, ei, ej, meas_c, info, iterations, lambda_init, 0, lm, out, stats, nullptr,
                              nullptr);
}

// "SE3 Pose Recovering" (src/Optimizer.cc:2624-2642): Tiw = [R | t / s] of the optimised Siw -> Tcw[12] row-major 3x4
void orc_essential_graph_recover_se3(int K, const OrcSim3* S, double* Tcw) {
  for (int k = 0; k < K; ++k) {
    double R[9];
    q2R(S[k].q, R);
    const double f = 1. / S[k].s;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Tcw[12 * k + 4 * r + c] = R[3 * r + c];
      Tcw[12 * k + 4 * r + 3] = S[k].t[r] * f;
    }
  }
}

// "Correct points" (src/Optimizer.cc:2645-2676): Pw' = correctedSwr.map(Srw.map(Pw)), positions are float (MapPoint::Tdata)
void orc_essential_graph_correct_points(int n, const float* Pw, const int32_t* ref, const OrcSim3* Scw_before, const OrcSim3* Scw_after,
                                        float* out) {
  for (int i = 0; i < n; ++i) {
    const S3 Srw = from_c(Scw_before[ref[i]]);
    const S3 Swr = s3_inv(from_c(Scw_after[ref[i]]));
    const double P[3] = {(double)Pw[3 * i], (double)Pw[3 * i + 1], (double)Pw[3 * i + 2]};
    double Pr[3], Pc[3];
    s3_map(Srw, P, Pr);
    s3_map(Swr, Pr, Pc);
    for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)Pc[k];
  }
}

}  // extern "C"

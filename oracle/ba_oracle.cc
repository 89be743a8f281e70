// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  CPU restatement of the bundle-adjustment path:
//   edges       src/Odom/g2otypes.h:321-541 (EdgeReproject), :725-884 (EdgeNavStateI), g2otypes.cpp:14-124
//               (EdgeNavStateBias, EdgeNavStatePriorPVRBias), NavState::IncSmall (src/Odom/NavState.h:47-82),
//               pinhole projection rounded to float (common/camera_models/camera_pinhole.h:70-106)
//   solver      g2o semantics: active set (sparse_optimizer.cpp:199-267), Levenberg-Marquardt
//               (optimization_algorithm_levenberg.cpp:61-189), Huber with float delta^2 (robust_kernel_impl.cpp:65-91,
//               base_edge.h:96-102), Schur complement (block_solver.hpp:353-486)
//   drivers     Optimizer::PoseOptimization visual (src/Optimizer.cc:1611-1874) and IMU/PVR incl. the kExactRobust
//               marginal prior (include/Optimizer.h:126-816), LocalBundleAdjustmentNavStatePRV (src/Optimizer.cc:21-769)
// Eigen is absent here: LDLT / inverse() / JacobiSVD are restated as Cholesky / Gauss-Jordan ("parity unpinned" for
// their last-bit rounding).  Pinned by the reference itself where it compiles (oracle/_ref, tests/test_oracle_ref.py): every edge
// and vertex class of g2otypes compiled unchanged against an Eigen / Sophus stand-in (residuals, Jacobians, updates to 1e-10), the
// camera models' Project() and the Huber kernel bit for bit, and g2o's own Levenberg-Marquardt solve() / optimize() driving every
// driver of this file bit-identically; plus numeric Jacobians, a scipy solve of the full normal equations and closed-form cases.
#include "ba_oracle.h"

#include <algorithm>
#include <functional>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "so3_oracle.h"

namespace {
using namespace orc;

struct NS {
  double p[3];
  Quat q;
  double v[3], bg[3], ba[3], dbg[3], dba[3];
};
NS from_c(const OrcNavState& s) {
  NS n;
  memcpy(n.p, s.p, 24);
  n.q = {s.q[0], s.q[1], s.q[2], s.q[3]};
  memcpy(n.v, s.v, 24); memcpy(n.bg, s.bg, 24); memcpy(n.ba, s.ba, 24); memcpy(n.dbg, s.dbg, 24); memcpy(n.dba, s.dba, 24);
  return n;
}
void to_c(const NS& n, OrcNavState* s) {
  memcpy(s->p, n.p, 24);
  s->q[0] = n.q.w; s->q[1] = n.q.x; s->q[2] = n.q.y; s->q[3] = n.q.z;
  memcpy(s->v, n.v, 24); memcpy(s->bg, n.bg, 24); memcpy(s->ba, n.ba, 24); memcpy(s->dbg, n.dbg, 24); memcpy(s->dba, n.dba, 24);
}
struct Cam {
  float fx, fy, cx, cy, bf;
  int model, num_k;
  float dist[8];
  M3 Rcb;
  double tcb[3];
};
Cam cam_from_c(const OrcCamera& c) {
  Cam k;
  k.fx = c.fx; k.fy = c.fy; k.cx = c.cx; k.cy = c.cy; k.bf = c.bf;
  k.model = c.model; k.num_k = c.num_k;
  memcpy(k.dist, c.dist, sizeof(k.dist));
  memcpy(k.Rcb.m, c.Rcb, 72);
  memcpy(k.tcb, c.tcb, 24);
  return k;
}

// ---- NavState::IncSmall (NavState.h:47-82), USE_P_PLUS_RDP ----------------------------------------------------
void inc_pr(NS& s, const double* d) {
  const M3 R = qmat(s.q);
  double Rd[3];
  mulv(R, d, Rd);
  for (int i = 0; i < 3; ++i) s.p[i] += Rd[i];
  s.q = qnormalized(qmul(s.q, so3_exp_q(d + 3)));
}
void inc_pvr(NS& s, const double* d) {
  const M3 R = qmat(s.q);
  double Rd[3];
  mulv(R, d, Rd);
  for (int i = 0; i < 3; ++i) s.p[i] += Rd[i];
  for (int i = 0; i < 3; ++i) s.v[i] += d[3 + i];
  s.q = qnormalized(qmul(s.q, so3_exp_q(d + 6)));
}
void inc_v(NS& s, const double* d) {
  for (int i = 0; i < 3; ++i) s.v[i] += d[i];
}
void inc_bias(NS& s, const double* d) {
  for (int i = 0; i < 3; ++i) s.dbg[i] += d[i];
  for (int i = 0; i < 3; ++i) s.dba[i] += d[3 + i];
}

// ---- camera models: Project() rounded to float pixels + d(img)/d(p3d) --------------------------------------------
// pinhole camera_pinhole.h:70-106, radtan camera_radtan.h:61-129, KB8 camera_kb8.h:68-157 (Jacobian formulas kept as
// the reference writes them, including radtan's radial-derivative term that starts at k3)
void cam_project(const Cam& c, const double P[3], float* u, float* v, double* J /*2x3 or null*/) {
  const double fx = (double)c.fx, fy = (double)c.fy;
  if (c.model == 1) {
    const float* k = c.dist;
    const float* p = k + c.num_k;
    const double invz = 1 / P[2];
    double x = P[0] * invz, y = P[1] * invz;
    const double x2 = x * x, y2 = y * y, xy = x * y, r2 = x2 + y2;
    double fd = 1, term_r = 1;
    for (int i = 0; i < c.num_k; ++i) {
      term_r *= r2;
      fd += k[i] * term_r;
    }
    if (J) {
      double fd2 = 0, coeff2 = 0;
      term_r = 1;
      for (int i = 2; i < c.num_k; ++i) {
        coeff2 += 2;
        fd2 += coeff2 * k[i] * term_r;
        term_r *= r2;
      }
      const double du_dx = fx * invz * (fd + fd2 * x2 + 2 * (p[0] * y + 3 * p[1] * x));
      const double du_dy = fx * invz * (fd2 * xy + 2 * (p[0] * x + p[1] * y));
      const double du_dz = -(x * du_dx + y * du_dy);
      const double dv_dx = du_dy * fy / fx;
      const double dv_dy = fy * invz * (fd + fd2 * y2 + 2 * (p[1] * x + 3 * p[0] * y));
      const double dv_dz = -(x * dv_dx + y * dv_dy);
      J[0] = du_dx; J[1] = du_dy; J[2] = du_dz; J[3] = dv_dx; J[4] = dv_dy; J[5] = dv_dz;
    }
    const double xd = x * fd + 2 * p[0] * xy + p[1] * (r2 + 2 * x2);
    const double yd = y * fd + 2 * p[1] * xy + p[0] * (r2 + 2 * y2);
    *u = (float)(fx * xd * 1.0 + c.cx);
    *v = (float)(fy * yd * 1.0 + c.cy);
    return;
  }
  if (c.model == 2) {
    const double x = P[0], y = P[1];
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r = std::sqrt(r2);
    if (r > (double)1e-5f) {
      const double k1 = c.dist[0], k2 = c.dist[1], k3 = c.dist[2], k4 = c.dist[3];
      const double z = P[2];
      const double theta = std::atan2(r, z), theta2 = theta * theta;
      double thetad = k4 * theta2;
      thetad += k3; thetad *= theta2; thetad += k2; thetad *= theta2; thetad += k1; thetad *= theta2; thetad += 1;
      thetad *= theta;
      const double mx = x * thetad / r, my = y * thetad / r;
      *u = (float)(fx * mx * 1.0 + c.cx);
      *v = (float)(fy * my * 1.0 + c.cy);
      if (J) {
        const double invr = 1. / r, d_r_d_x = x * invr, d_r_d_y = y * invr;
        const double tmp = 1. / (z * z + r2);
        const double d_thetad_x = d_r_d_x * z * tmp, d_thetad_y = d_r_d_y * z * tmp;
        double dd = 9.0 * k4 * theta2;
        dd += 7.0 * k3; dd *= theta2; dd += 5.0 * k2; dd *= theta2; dd += 3.0 * k1; dd *= theta2; dd += 1.0;
        const double invr2 = invr * invr;
        J[0] = fx * (x * r * dd * d_thetad_x + y2 * thetad / r) * invr2;
        J[1] = fx * x * (dd * d_thetad_y * r - y * thetad / r) * invr2;
        J[2] = -fx * x * dd * tmp;
        J[3] = J[1] * fy / fx;
        J[4] = fy * (y * r * dd * d_thetad_y + x2 * thetad / r) * invr2;
        J[5] = -fy * y * dd * tmp;
      }
      return;
    }
  }
  const double invz = 1. / P[2];
  *u = (float)(fx * P[0] * invz + c.cx);
  *v = (float)(fy * P[1] * invz + c.cy);
  if (J) {
    const double invz2 = invz * invz;
    J[0] = fx * invz; J[1] = 0; J[2] = -fx * P[0] * invz2;
    J[3] = 0; J[4] = fy * invz; J[5] = -fy * P[1] * invz2;
  }
}

// ---- EdgeReproject (g2otypes.h:338-541), NV = 2 ---------------------------------------------------------------
// e = obs - pi(Rcw Xw + tcw), pi rounded to float (camera_pinhole.h:81-82); returns depth (GetDepth, :431-436)
double reproj_error(const Cam& c, const NS& s, const double Xw[3], const float obs[3], bool stereo, double e[3]) {
  const M3 Rwb = qmat(s.q), Rcw = mul(c.Rcb, tr(Rwb));
  double t[3], Pc[3];
  mulv(Rcw, s.p, t);
  for (int i = 0; i < 3; ++i) t[i] = -t[i] + c.tcb[i];
  mulv(Rcw, Xw, Pc);
  for (int i = 0; i < 3; ++i) Pc[i] += t[i];
  float u, v;
  cam_project(c, Pc, &u, &v, nullptr);
  e[0] = (double)obs[0] - (double)u;
  e[1] = (double)obs[1] - (double)v;
  e[2] = stereo ? (double)obs[2] - ((double)u - (double)c.bf / Pc[2]) : 0.0;
  // GetDepth: Rcw.row(2) * wX + tcw(2)
  return Rcw.m[6] * Xw[0] + Rcw.m[7] * Xw[1] + Rcw.m[8] * Xw[2] + t[2];
}
// Jp (de/d dp), Jr (de/d dphi), JX (de/dX): rows 0..DE-1 of 3x3 row-major blocks
void reproj_jac(const Cam& c, const NS& s, const double Xw[3], bool stereo, double Jp[9], double Jr[9], double JX[9]) {
  const M3 Rwb = qmat(s.q), Rcw = mul(c.Rcb, tr(Rwb));
  double t[3], Pc[3];
  mulv(Rcw, s.p, t);
  for (int i = 0; i < 3; ++i) t[i] = -t[i] + c.tcb[i];
  mulv(Rcw, Xw, Pc);
  for (int i = 0; i < 3; ++i) Pc[i] += t[i];
  const double invz = 1 / Pc[2], invz_2 = invz * invz;
  M3 Jproj = {{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  {
    float u, v;
    double Jc[6];
    cam_project(c, Pc, &u, &v, Jc);
    for (int i = 0; i < 6; ++i) Jproj.m[i] = -Jc[i];
  }
  if (stereo) {
    Jproj.m[6] = Jproj.m[0];
    Jproj.m[7] = Jproj.m[1];
    Jproj.m[8] = Jproj.m[2] - (double)c.bf * invz_2;
  }
  const M3 JdP = mul(Jproj, scale(c.Rcb, -1.0));
  double d[3] = {Xw[0] - s.p[0], Xw[1] - s.p[1], Xw[2] - s.p[2]}, Paux[3];
  mulv(tr(Rwb), d, Paux);
  const M3 JdR = mul(mul(Jproj, c.Rcb), hat(Paux));
  const M3 JdX = mul(Jproj, Rcw);
  memcpy(Jp, JdP.m, 72);
  memcpy(Jr, JdR.m, 72);
  memcpy(JX, JdX.m, 72);
}

// ---- EdgeNavStateI<NV> (g2otypes.h:725-884).  prv = true: residual / column order P,R,V (NV=5), else P,V,R (NV=3)
void navstate_error(const NS& si, const NS& sj, const OrcImuPreint& m, const double gw[3], bool prv, double e[9]) {
  const M3 RiT = tr(qmat(si.q));
  const int idR = prv ? 3 : 6, idV = 9 - idR;
  const double dt = m.dt;
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) a[k] = sj.p[k] - si.p[k] - si.v[k] * dt - gw[k] * (dt * dt / 2);
  mulv(RiT, a, b);
  M3 Jgp, Jap, Jgv, Jav, JgR;
  memcpy(Jgp.m, m.Jgp, 72); memcpy(Jap.m, m.Jap, 72); memcpy(Jgv.m, m.Jgv, 72); memcpy(Jav.m, m.Jav, 72); memcpy(JgR.m, m.JgR, 72);
  double t1[3], t2[3];
  mulv(Jgp, si.dbg, t1);
  mulv(Jap, si.dba, t2);
  for (int k = 0; k < 3; ++k) e[k] = b[k] - (m.pij[k] + t1[k] + t2[k]);
  // eR = Log((dRij Exp(JgR dbg))^-1 * (Ri^-1 Rj))
  M3 Rij;
  memcpy(Rij.m, m.Rij, 72);
  double w[3];
  mulv(JgR, si.dbg, w);
  const Quat qm = qnormalized(qmul(qnormalized(mquat(Rij)), so3_exp_q(w)));
  const Quat qij = qnormalized(qmul(qconj(si.q), sj.q));
  so3_log_q(qnormalized(qmul(qconj(qm), qij)), e + idR);
  for (int k = 0; k < 3; ++k) a[k] = sj.v[k] - si.v[k] - gw[k] * dt;
  mulv(RiT, a, b);
  mulv(Jgv, si.dbg, t1);
  mulv(Jav, si.dba, t2);
  for (int k = 0; k < 3; ++k) e[idV + k] = b[k] - (m.vij[k] + t1[k] + t2[k]);
}
inline void setb(double* A, int ld, int r, int c, const M3& b) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[ld * (r + i) + c + j] = b.m[3 * i + j];
}
// Ji, Jj: 9x9 (state columns in residual order), Jb: 9x6.  `e` must hold the current error (eR is read from it).
void navstate_jac(const NS& si, const NS& sj, const OrcImuPreint& m, const double gw[3], bool prv, const double e[9],
                  double Ji[81], double Jj[81], double Jb[54]) {
  const M3 Ri = qmat(si.q), RiT = tr(Ri), Rj = qmat(sj.q);
  const int idR = prv ? 3 : 6, idV = 9 - idR;
  const double dt = m.dt;
  memset(Ji, 0, 81 * 8); memset(Jj, 0, 81 * 8); memset(Jb, 0, 54 * 8);
  M3 Jgp, Jap, Jgv, Jav, JgR;
  memcpy(Jgp.m, m.Jgp, 72); memcpy(Jap.m, m.Jap, 72); memcpy(Jgv.m, m.Jgv, 72); memcpy(Jav.m, m.Jav, 72); memcpy(JgR.m, m.JgR, 72);
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) a[k] = sj.p[k] - si.p[k] - si.v[k] * dt - gw[k] * (dt * dt / 2);
  mulv(RiT, a, b);
  setb(Ji, 9, 0, idR, hat(b));
  setb(Ji, 9, 0, 0, scale(ident(), -1.0));
  setb(Ji, 9, 0, idV, scale(scale(RiT, -1.0), dt));
  setb(Jb, 6, 0, 0, scale(Jgp, -1.0));
  setb(Jb, 6, 0, 3, scale(Jap, -1.0));
  setb(Jj, 9, 0, 0, mul(RiT, Rj));
  for (int k = 0; k < 3; ++k) a[k] = sj.v[k] - si.v[k] - gw[k] * dt;
  mulv(RiT, a, b);
  setb(Ji, 9, idV, idR, hat(b));
  setb(Ji, 9, idV, idV, scale(RiT, -1.0));
  setb(Jb, 6, idV, 0, scale(Jgv, -1.0));
  setb(Jb, 6, idV, 3, scale(Jav, -1.0));
  setb(Jj, 9, idV, idV, RiT);
  const double* eR = e + idR;
  const M3 Jrinv = so3_JrInv(eR);
  // -Jrinv * (Rj^-1 Ri).matrix()
  const M3 RjTRi = qmat(qnormalized(qmul(qconj(sj.q), si.q)));
  setb(Ji, 9, idR, idR, scale(mul(Jrinv, RjTRi), -1.0));
  double neR[3] = {-eR[0], -eR[1], -eR[2]}, w[3];
  mulv(JgR, si.dbg, w);
  const M3 T = mul(mul(mul(scale(Jrinv, -1.0), so3_Exp(neR)), so3_Jr(w)), JgR);
  setb(Jb, 6, idR, 0, T);
  setb(Jj, 9, idR, idR, Jrinv);
}

// ---- EdgeNavStatePriorPVRBias (g2otypes.cpp:84-124): order P V R bg ba ----------------------------------------
void prior_error(const NS& s, const NS& sb, const NS& pr, double e[15]) {
  const Quat qbar_inv = qconj(pr.q);
  const M3 Rbw = qmat(qbar_inv);
  double d[3] = {s.p[0] - pr.p[0], s.p[1] - pr.p[1], s.p[2] - pr.p[2]};
  mulv(Rbw, d, e);
  for (int k = 0; k < 3; ++k) e[3 + k] = s.v[k] - pr.v[k];
  so3_log_q(qnormalized(qmul(qbar_inv, s.q)), e + 6);
  for (int k = 0; k < 3; ++k) e[9 + k] = sb.bg[k] + sb.dbg[k] - (pr.bg[k] + pr.dbg[k]);
  for (int k = 0; k < 3; ++k) e[12 + k] = sb.ba[k] + sb.dba[k] - (pr.ba[k] + pr.dba[k]);
}
void prior_jac(const NS& s, const NS& pr, const double e[15], double Jpvr[135], double Jb[90]) {
  memset(Jpvr, 0, 135 * 8);
  memset(Jb, 0, 90 * 8);
  setb(Jpvr, 9, 0, 0, mul(tr(qmat(pr.q)), qmat(s.q)));
  setb(Jpvr, 9, 3, 3, ident());
  setb(Jpvr, 9, 6, 6, so3_JrInv(e + 6));
  setb(Jb, 6, 9, 0, ident());
  setb(Jb, 6, 12, 3, ident());
}

// ---- small dense linear algebra --------------------------------------------------------------------------------
// Gauss-Jordan inverse with partial pivoting (stands in for Eigen's inverse() / JacobiSVD pseudo-inverse without clamping)
bool inverse(const double* A, int n, double* Ai) {
  std::vector<double> M(A, A + n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Ai[i * n + j] = i == j;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(M[r * n + c]) > std::fabs(M[piv * n + c])) piv = r;
    if (M[piv * n + c] == 0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        std::swap(M[piv * n + j], M[c * n + j]);
        std::swap(Ai[piv * n + j], Ai[c * n + j]);
      }
    const double d = 1.0 / M[c * n + c];
    for (int j = 0; j < n; ++j) {
      M[c * n + j] *= d;
      Ai[c * n + j] *= d;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[r * n + c];
      if (f == 0) continue;
      for (int j = 0; j < n; ++j) {
        M[r * n + j] -= f * M[c * n + j];
        Ai[r * n + j] -= f * Ai[c * n + j];
      }
    }
  }
  return true;
}
// 3x3 inverse by cofactors (Eigen's fixed-size inverse, block_solver.hpp:389)
void inv3(const double* D, double* I) {
  const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
  const double det = D[0] * c00 + D[1] * c01 + D[2] * c02, id = 1.0 / det;
  I[0] = c00 * id; I[1] = (D[2] * D[7] - D[1] * D[8]) * id; I[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  I[3] = c01 * id; I[4] = (D[0] * D[8] - D[2] * D[6]) * id; I[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  I[6] = c02 * id; I[7] = (D[1] * D[6] - D[0] * D[7]) * id; I[8] = (D[0] * D[4] - D[1] * D[3]) * id;
}
// Cholesky solve A x = b (A symmetric, full storage); false when a pivot is not positive (LDLT::isPositive)
// Dense Cholesky (lower triangle, row-major) + two triangular solves.  Every entry is the scalar recurrence
//   L[i][j] = (A[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j],  k ascending,
// exactly as before; two things make map-sized systems (global BA, n ~ 6000) practical without changing one bit of it:
//  * skyline: first[i] = first structurally non-zero column of row i (fill-in never moves it left), sums start at
//    max(first[i], first[j]) — the skipped terms are products with exact zeros;
//  * column blocks of 64: once the diagonal block is done, the rows below are independent and are spread over the host
//    threads (each entry still sums in the same order).
bool chol_solve(std::vector<double>& A, int n, const double* b, double* x) {
  const size_t N = (size_t)n;
  std::vector<int> first(n);
  for (int i = 0; i < n; ++i) {
    int f = 0;
    while (f < i && A[i * N + f] == 0.0) ++f;
    first[i] = f;
  }
  auto entry = [&](int i, int j) {  // L[i][j] for i > j, given L[j][j]
    double s = A[i * N + j];
    const double* ri = &A[i * N];
    const double* rj = &A[j * N];
    for (int k = std::max(first[i], first[j]); k < j; ++k) s -= ri[k] * rj[k];
    A[i * N + j] = s / rj[j];
  };
  const int NB = 64;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int j1 = std::min(j0 + NB, n);
    for (int j = j0; j < j1; ++j) {  // diagonal block, serial
      double d = A[j * N + j];
      for (int k = first[j]; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
      if (!(d > 0) || !std::isfinite(d)) return false;
      A[j * N + j] = std::sqrt(d);
      for (int i = j + 1; i < j1; ++i)
        if (first[i] <= j) entry(i, j);
    }
    const int rows = n - j1;
    if (rows <= 0) continue;
    auto work = [&](int r0, int r1) {
      for (int i = r0; i < r1; ++i)
        for (int j = std::max(j0, first[i]); j < j1; ++j) entry(i, j);
    };
    const int nt = (int)std::min<unsigned>(hw, (unsigned)std::max(1, rows / 64));
    if (nt <= 1) {
      work(j1, n);
    } else {
      std::vector<std::thread> th;
      // interleaved chunks: the rows' costs grow with i
      for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
          for (int c = j1 + 16 * t; c < n; c += 16 * nt) work(c, std::min(c + 16, n));
        });
      for (auto& q : th) q.join();
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i * n + k] * y[k];
    y[i] = s / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * x[k];
    x[i] = s / A[i * n + i];
  }
  return true;
}

// ---- RobustKernelHuber (robust_kernel_impl.cpp:65-91): delta double, delta^2 stored as float -----------------------
struct Huber {
  bool on = false;
  double delta = 0;
  float dsqr = 0;
  void set(double d) {
    on = true;
    delta = d;
    dsqr = (float)(d * d);
  }
  void rho(double e, double r[2]) const {
    if (!on || e <= (double)dsqr) {
      r[0] = e;
      r[1] = 1.;
    } else {
      const double sq = std::sqrt(e);
      r[0] = 2 * sq * delta - (double)dsqr;
      r[1] = delta / sq;
    }
  }
};

// ================================================================================================================
// Graph engine.  A "state" carries up to three vertices: slot 0 = PR(6) or PVR(9), slot 1 = V(3) (PR-V-B layout
// only), slot 2 = Bias(6).  Hessian index order = states in the given order, slots ascending (the reference's vertex
// ids 3k, 3k+1, 3k+2 resp. 0..3; sparse_optimizer.cpp:166-190), marginalised points after.
struct DenseEdge {  // IMU / bias / prior factors
  int type;         // 0 IMU, 1 bias, 2 prior (PVR-Bias)
  int si, sj;       // states (prior: si only)
  const OrcImuPreint* pre = nullptr;
  NS prior;
  int D;
  std::vector<double> info;  // D x D
  Huber rk;
  double err[15];
  double chi2 = 0;
};
struct VisEdge {
  int s, p;
  float obs[3];
  double w;  // invSigma2
  bool stereo, close;
  int level;
  Huber rk;
  double err[3];
  double chi2 = 0;
};

// LM driver used by Graph::optimize (nullptr: orc_lm_optimize); a test hook, see orc_set_lm_driver
OrcLmDriver g_lm_driver = nullptr;
OrcMargDump* g_marg_dump = nullptr;  // test hook, see orc_set_marg_dump

struct Graph {
  bool pvr;  // slot 0 is PVR(9) (pose optimisation) instead of PR(6)
  Cam cam;
  double gw[3];
  std::vector<NS> st;
  std::vector<uint8_t> fix0, has_vb, fix_vb;  // per state
  std::vector<double> X;                      // points
  bool points_free = false;
  std::vector<VisEdge> vis;
  std::vector<DenseEdge> den;
  volatile const bool* stop = nullptr;
  // index mapping
  std::vector<int> off0, off1, off2;
  int np = 0;
  std::vector<uint8_t> vis_active, den_active, pt_active;
  // system
  std::vector<double> H, b, x, Hll, bl, W, xl;  // H np x np, W per visual edge [d0][3]
  // LM
  double lambda = 0, ni = 2, user_lambda = 0;
  int nBad = 0;
  int total_iters = 0;
  std::vector<NS> st_bak;
  std::vector<double> X_bak;
  // VertexScale (g2otypes.h:294-311) of GlobalBundleAdjustmentNavStatePRV with bScaleOpt (src/Optimizer.cc:843-851): the
  // visual edges become EdgeReprojectPRS[Stereo] (Xw = sc * X, :1132-1200).  One extra row / column of the reduced system
  // after every keyframe vertex (id_scale = maxKFid + 1).  Off by default: nothing below changes the arithmetic then.
  bool has_scale = false;
  double sc = 1.0, sc_bak = 1.0;
  int off_s = -1;
  std::vector<double> Ws;  // per visual edge [3]: Js^T (w Omega) JX
  // VertexGThetaXYRwI (g2otypes.h:674-698) of the IMU initialiser's call (pimu_initiator, src/Optimizer.cc:852-865): the
  // inertial edges become EdgeNavStatePRVG with gw = RwI * GI; two more rows / columns after the scale (id_g =
  // maxKFid + 2).  Off by default.
  bool has_g = false;
  Quat qwI{1, 0, 0, 0}, qwI_bak{1, 0, 0, 0};
  double GI[3] = {0, 0, 0};
  int off_g = -1;
  double JGbuf[18];
  void gravity(double g[3]) const {
    if (!has_g) {
      g[0] = gw[0]; g[1] = gw[1]; g[2] = gw[2];
      return;
    }
    mulv(qmat(qwI), GI, g);
  }

  int d0() const { return pvr ? 9 : 6; }
  void world_point(const VisEdge& e, double Xw[3]) const {
    for (int k = 0; k < 3; ++k) Xw[k] = has_scale ? X[3 * e.p + k] * sc : X[3 * e.p + k];
  }

  void initialize() {  // initializeOptimization(0) + buildIndexMapping
    const int K = (int)st.size();
    off0.assign(K, -1); off1.assign(K, -1); off2.assign(K, -1);
    np = 0;
    for (int k = 0; k < K; ++k) {
      if (!fix0[k]) { off0[k] = np; np += d0(); }
      if (has_vb[k] && !fix_vb[k]) {
        if (!pvr) { off1[k] = np; np += 3; }
        off2[k] = np; np += 6;
      }
    }
    off_s = off_g = -1;
    if (has_scale) { off_s = np; np += 1; }
    if (has_g) { off_g = np; np += 2; }
    vis_active.assign(vis.size(), 0);
    pt_active.assign(X.size() / 3, 0);
    for (size_t i = 0; i < vis.size(); ++i) {
      const VisEdge& e = vis[i];
      const bool anyfree = off0[e.s] >= 0 || points_free || has_scale;
      vis_active[i] = e.level == 0 && anyfree;
      if (vis_active[i] && points_free) pt_active[e.p] = 1;
    }
    den_active.assign(den.size(), 0);
    for (size_t i = 0; i < den.size(); ++i) {
      const DenseEdge& e = den[i];
      bool anyfree = false;
      if (e.type == 0) anyfree = off0[e.si] >= 0 || off0[e.sj] >= 0 || off1[e.si] >= 0 || off1[e.sj] >= 0 || off2[e.si] >= 0 || has_g;
      if (e.type == 3) anyfree = off2[e.si] >= 0;
      if (e.type == 1) anyfree = off2[e.si] >= 0 || off2[e.sj] >= 0;
      if (e.type == 2) anyfree = off0[e.si] >= 0 || off2[e.si] >= 0;
      den_active[i] = anyfree;
    }
  }
  void vis_error(VisEdge& e) {
    double Xw[3];
    world_point(e, Xw);
    reproj_error(cam, st[e.s], Xw, e.obs, e.stereo, e.err);
    double c = 0;
    for (int k = 0; k < (e.stereo ? 3 : 2); ++k) c += e.err[k] * (e.w * e.err[k]);
    e.chi2 = c;
  }
  double vis_depth(const VisEdge& e) {
    double t[3], Xw[3];
    world_point(e, Xw);
    return reproj_error(cam, st[e.s], Xw, e.obs, e.stereo, t);
  }
  void den_error(DenseEdge& e) {
    if (e.type == 0) {
      double g[3];
      gravity(g);
      navstate_error(st[e.si], st[e.sj], *e.pre, g, !pvr, e.err);
    }
    if (e.type == 3) {  // EdgeNavStateBias between the fixed prior-bias vertex (e.prior) and keyframe si's bias
      const NS &a = e.prior, &c = st[e.si];
      for (int k = 0; k < 3; ++k) e.err[k] = (c.bg[k] + c.dbg[k]) - (a.bg[k] + a.dbg[k]);
      for (int k = 0; k < 3; ++k) e.err[3 + k] = (c.ba[k] + c.dba[k]) - (a.ba[k] + a.dba[k]);
    }
    if (e.type == 1) {
      const NS &a = st[e.si], &c = st[e.sj];
      for (int k = 0; k < 3; ++k) e.err[k] = (c.bg[k] + c.dbg[k]) - (a.bg[k] + a.dbg[k]);
      for (int k = 0; k < 3; ++k) e.err[3 + k] = (c.ba[k] + c.dba[k]) - (a.ba[k] + a.dba[k]);
    }
    if (e.type == 2) prior_error(st[e.si], st[e.si], e.prior, e.err);
    double c = 0;
    for (int i = 0; i < e.D; ++i) {
      double s = 0;
      for (int j = 0; j < e.D; ++j) s += e.info[i * e.D + j] * e.err[j];
      c += e.err[i] * s;
    }
    e.chi2 = c;
  }
  void compute_active_errors() {
    for (size_t i = 0; i < vis.size(); ++i)
      if (vis_active[i]) vis_error(vis[i]);
    for (size_t i = 0; i < den.size(); ++i)
      if (den_active[i]) den_error(den[i]);
  }
  double active_robust_chi2() const {
    double chi = 0, r[2];
    for (size_t i = 0; i < den.size(); ++i)
      if (den_active[i]) {
        den[i].rk.rho(den[i].chi2, r);
        chi += r[0];
      }
    for (size_t i = 0; i < vis.size(); ++i)
      if (vis_active[i]) {
        vis[i].rk.rho(vis[i].chi2, r);
        chi += r[0];
      }
    return chi;
  }
  // H[oa.., ob..] += Ja^T (w Omega) Jb ; b[oa] += Ja^T (-w Omega e) for every pair of free vertex blocks
  struct Blk {
    int off, dim;
    const double* J;  // D x ld, columns [c0, c0+dim)
    int ld, c0;
  };
  void add_dense(const DenseEdge& e, const std::vector<Blk>& blks) {
    double r[2];
    e.rk.rho(e.chi2, r);
    const int D = e.D;
    std::vector<double> Om(D * D), oe(D);
    for (int i = 0; i < D * D; ++i) Om[i] = r[1] * e.info[i];
    for (int i = 0; i < D; ++i) {
      double s = 0;
      for (int j = 0; j < D; ++j) s += e.info[i * D + j] * e.err[j];
      oe[i] = -s * r[1];
    }
    for (const Blk& A : blks) {
      if (A.off < 0) continue;
      std::vector<double> AtO(A.dim * D);
      for (int a = 0; a < A.dim; ++a)
        for (int j = 0; j < D; ++j) {
          double s = 0;
          for (int i = 0; i < D; ++i) s += A.J[i * A.ld + A.c0 + a] * Om[i * D + j];
          AtO[a * D + j] = s;
        }
      for (int a = 0; a < A.dim; ++a) {
        double s = 0;
        for (int i = 0; i < D; ++i) s += A.J[i * A.ld + A.c0 + a] * oe[i];
        b[A.off + a] += s;
      }
      for (const Blk& B : blks) {
        if (B.off < 0) continue;
        for (int a = 0; a < A.dim; ++a)
          for (int c = 0; c < B.dim; ++c) {
            double s = 0;
            for (int j = 0; j < D; ++j) s += AtO[a * D + j] * B.J[j * B.ld + B.c0 + c];
            H[(A.off + a) * np + B.off + c] += s;
          }
      }
    }
  }
  std::vector<Blk> dense_blocks(const DenseEdge& e, double* Ji, double* Jj, double* Jb) {
    std::vector<Blk> blks;
    if (e.type == 3) {
      for (int i = 0; i < 36; ++i) Jj[i] = 0;
      for (int i = 0; i < 6; ++i) Jj[7 * i] = 1;
      blks = {{off2[e.si], 6, Jj, 6, 0}};
      return blks;
    }
    if (e.type == 0) {
      double g[3];
      gravity(g);
      navstate_jac(st[e.si], st[e.sj], *e.pre, g, !pvr, e.err, Ji, Jj, Jb);
      if (has_g) {  // JG (g2otypes.h:868-876): P rows RiT dt^2/2 RwI GI^[:, 0:2], V rows RiT dt RwI GI^[:, 0:2], R rows 0
        const M3 A = mul(tr(qmat(st[e.si].q)), mul(qmat(qwI), hat(GI)));
        const double dt = e.pre->dt;
        const int idR = pvr ? 6 : 3, idV = 9 - idR;
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 2; ++c) {
            JGbuf[2 * r + c] = A.m[3 * r + c] * (dt * dt / 2.0);
            JGbuf[2 * (idR + r) + c] = 0.0;
            JGbuf[2 * (idV + r) + c] = A.m[3 * r + c] * dt;
          }
      }
      if (pvr) {
        blks = {{off0[e.si], 9, Ji, 9, 0}, {off0[e.sj], 9, Jj, 9, 0}, {off2[e.si], 6, Jb, 6, 0}};
      } else {
        blks = {{off0[e.si], 6, Ji, 9, 0}, {off0[e.sj], 6, Jj, 9, 0}, {off1[e.si], 3, Ji, 9, 6},
                {off1[e.sj], 3, Jj, 9, 6}, {off2[e.si], 6, Jb, 6, 0}};
      }
      if (has_g) blks.push_back({off_g, 2, JGbuf, 2, 0});
    } else if (e.type == 1) {
      for (int i = 0; i < 36; ++i) Ji[i] = Jj[i] = 0;
      for (int i = 0; i < 6; ++i) {
        Ji[7 * i] = -1;
        Jj[7 * i] = 1;
      }
      blks = {{off2[e.si], 6, Ji, 6, 0}, {off2[e.sj], 6, Jj, 6, 0}};
    } else {
      prior_jac(st[e.si], e.prior, e.err, Ji, Jb);
      blks = {{off0[e.si], 9, Ji, 9, 0}, {off2[e.si], 6, Jb, 6, 0}};
    }
    return blks;
  }
  void build_system() {
    const int P = (int)X.size() / 3, dv = d0();
    H.assign((size_t)np * np, 0.0);
    b.assign(np, 0.0);
    if (points_free) {
      Hll.assign((size_t)P * 9, 0.0);
      bl.assign((size_t)P * 3, 0.0);
      W.assign(vis.size() * (size_t)dv * 3, 0.0);
    }
    if (has_scale) Ws.assign(vis.size() * (size_t)3, 0.0);
    double Ji[135], Jj[81], Jb[90];
    for (size_t i = 0; i < den.size(); ++i)
      if (den_active[i]) add_dense(den[i], dense_blocks(den[i], Ji, Jj, Jb));
    for (size_t i = 0; i < vis.size(); ++i) {
      if (!vis_active[i]) continue;
      const VisEdge& e = vis[i];
      const int DE = e.stereo ? 3 : 2;
      double Jp[9], Jr[9], JX[9], r[2], Xw[3], Js[3] = {0, 0, 0};
      world_point(e, Xw);
      reproj_jac(cam, st[e.s], Xw, e.stereo, Jp, Jr, JX);
      if (has_scale) {  // J_scale = (Jproj Rcw) Xh, J_point = (Jproj Rcw) sc (g2otypes.h:517-521)
        for (int k = 0; k < DE; ++k) Js[k] = JX[3 * k] * X[3 * e.p] + JX[3 * k + 1] * X[3 * e.p + 1] + JX[3 * k + 2] * X[3 * e.p + 2];
        for (int k = 0; k < 9; ++k) JX[k] *= sc;
      }
      e.rk.rho(e.chi2, r);
      const double w = r[1] * e.w;
      double J[3][9];  // pose Jacobian DE x dv: dp | (dv) | dphi
      memset(J, 0, sizeof(J));
      for (int k = 0; k < DE; ++k)
        for (int c = 0; c < 3; ++c) {
          J[k][c] = Jp[3 * k + c];
          J[k][dv - 3 + c] = Jr[3 * k + c];
        }
      double oe[3];
      for (int k = 0; k < DE; ++k) oe[k] = -(e.w * e.err[k]) * r[1];
      const int o = off0[e.s];
      if (o >= 0) {
        for (int a = 0; a < dv; ++a) {
          double s = 0;
          for (int k = 0; k < DE; ++k) s += J[k][a] * oe[k];
          b[o + a] += s;
          for (int c = 0; c < dv; ++c) {
            double h = 0;
            for (int k = 0; k < DE; ++k) h += (J[k][a] * w) * J[k][c];
            H[(size_t)(o + a) * np + o + c] += h;
          }
        }
      }
      if (has_scale) {
        double sb = 0, sh = 0;
        for (int k = 0; k < DE; ++k) {
          sb += Js[k] * oe[k];
          sh += (Js[k] * w) * Js[k];
        }
        b[off_s] += sb;
        H[(size_t)off_s * np + off_s] += sh;
        if (o >= 0)
          for (int a = 0; a < dv; ++a) {
            double h = 0;
            for (int k = 0; k < DE; ++k) h += (J[k][a] * w) * Js[k];
            H[(size_t)(o + a) * np + off_s] += h;
            H[(size_t)off_s * np + o + a] += h;
          }
        if (points_free)
          for (int c = 0; c < 3; ++c) {
            double h = 0;
            for (int k = 0; k < DE; ++k) h += (Js[k] * w) * JX[3 * k + c];
            Ws[i * 3 + c] = h;
          }
      }
      if (points_free) {
        double* Hl = &Hll[(size_t)9 * e.p];
        double* bp = &bl[(size_t)3 * e.p];
        for (int a = 0; a < 3; ++a) {
          double s = 0;
          for (int k = 0; k < DE; ++k) s += JX[3 * k + a] * oe[k];
          bp[a] += s;
          for (int c = 0; c < 3; ++c) {
            double h = 0;
            for (int k = 0; k < DE; ++k) h += (JX[3 * k + a] * w) * JX[3 * k + c];
            Hl[3 * a + c] += h;
          }
        }
        if (o >= 0) {
          double* Wp = &W[i * (size_t)dv * 3];
          for (int a = 0; a < dv; ++a)
            for (int c = 0; c < 3; ++c) {
              double h = 0;
              for (int k = 0; k < DE; ++k) h += (J[k][a] * w) * JX[3 * k + c];
              Wp[3 * a + c] = h;
            }
        }
      }
    }
  }
  double lambda_init() const {
    if (user_lambda > 0) return user_lambda;
    double mx = 0;
    for (int i = 0; i < np; ++i) mx = std::max(std::fabs(H[(size_t)i * np + i]), mx);
    if (points_free)
      for (size_t p = 0; p < pt_active.size(); ++p)
        if (pt_active[p])
          for (int k = 0; k < 3; ++k) mx = std::max(std::fabs(Hll[9 * p + 4 * k]), mx);
    return 1e-5 * mx;
  }
  // (H + lambda I) x = b, with the Schur complement over the free points (block_solver.hpp:353-486)
  double lambda_pose = -1;            // sharded form: lambda on the pose diagonal (rank 0: lambda, others: 0); < 0: = lambda
  std::vector<double> S_out, bs_out;  // reduced camera system of the last solve_system() (before factorisation)
  bool solve_system() {
    const int P = (int)X.size() / 3, dv = d0();
    std::vector<double> S(H);
    for (int i = 0; i < np; ++i) S[(size_t)i * np + i] += lambda_pose >= 0 ? lambda_pose : lambda;
    std::vector<double> bs(b);
    std::vector<double> Dinv;
    if (points_free) {
      Dinv.assign((size_t)P * 9, 0.0);
      xl.assign((size_t)P * 3, 0.0);
      // edges are sorted by point: walk each point's range
      size_t i0 = 0;
      while (i0 < vis.size()) {
        size_t i1 = i0;
        const int p = vis[i0].p;
        while (i1 < vis.size() && vis[i1].p == p) ++i1;
        if (pt_active[p]) {
          double D[9];
          memcpy(D, &Hll[(size_t)9 * p], 72);
          D[0] += lambda; D[4] += lambda; D[8] += lambda;
          double* Di = &Dinv[(size_t)9 * p];
          inv3(D, Di);
          double db[3];
          for (int a = 0; a < 3; ++a) db[a] = Di[3 * a] * bl[3 * p] + Di[3 * a + 1] * bl[3 * p + 1] + Di[3 * a + 2] * bl[3 * p + 2];
          for (size_t a = i0; a < i1; ++a) {
            if (!vis_active[a] || off0[vis[a].s] < 0) continue;
            const double* Wa = &W[a * (size_t)dv * 3];
            const int oa = off0[vis[a].s];
            double WD[27];
            for (int r = 0; r < dv; ++r)
              for (int c = 0; c < 3; ++c) WD[3 * r + c] = Wa[3 * r] * Di[c] + Wa[3 * r + 1] * Di[3 + c] + Wa[3 * r + 2] * Di[6 + c];
            for (int r = 0; r < dv; ++r) bs[oa + r] -= Wa[3 * r] * db[0] + Wa[3 * r + 1] * db[1] + Wa[3 * r + 2] * db[2];
            for (size_t c2 = i0; c2 < i1; ++c2) {
              if (!vis_active[c2] || off0[vis[c2].s] < 0) continue;
              const double* Wb = &W[c2 * (size_t)dv * 3];
              const int ob = off0[vis[c2].s];
              for (int r = 0; r < dv; ++r)
                for (int c = 0; c < dv; ++c)
                  S[(size_t)(oa + r) * np + ob + c] -= WD[3 * r] * Wb[3 * c] + WD[3 * r + 1] * Wb[3 * c + 1] + WD[3 * r + 2] * Wb[3 * c + 2];
            }
          }
          if (has_scale) {  // the scale row / column: every active edge of the point carries a 1x3 block Ws
            for (size_t a = i0; a < i1; ++a) {
              if (!vis_active[a]) continue;
              const double* Wa = &Ws[a * 3];
              double WD[3];
              for (int c = 0; c < 3; ++c) WD[c] = Wa[0] * Di[c] + Wa[1] * Di[3 + c] + Wa[2] * Di[6 + c];
              bs[off_s] -= Wa[0] * db[0] + Wa[1] * db[1] + Wa[2] * db[2];
              for (size_t c2 = i0; c2 < i1; ++c2) {
                if (!vis_active[c2]) continue;
                const double* Wc = &Ws[c2 * 3];
                S[(size_t)off_s * np + off_s] -= WD[0] * Wc[0] + WD[1] * Wc[1] + WD[2] * Wc[2];
                const int ob = off0[vis[c2].s];
                if (ob < 0) continue;
                const double* Wb = &W[c2 * (size_t)dv * 3];
                for (int c = 0; c < dv; ++c) {
                  const double v = WD[0] * Wb[3 * c] + WD[1] * Wb[3 * c + 1] + WD[2] * Wb[3 * c + 2];
                  S[(size_t)off_s * np + ob + c] -= v;
                  S[(size_t)(ob + c) * np + off_s] -= v;
                }
              }
            }
          }
        }
        i0 = i1;
      }
    }
    if ((int)x.size() != np) x.assign(np, 0.0);
    S_out = S;
    bs_out = bs;
    bool ok = np == 0 ? true : chol_solve(S, np, bs.data(), x.data());
    if (!ok) return false;
    if (points_free) {
      size_t i0 = 0;
      while (i0 < vis.size()) {
        size_t i1 = i0;
        const int p = vis[i0].p;
        while (i1 < vis.size() && vis[i1].p == p) ++i1;
        if (pt_active[p]) {
          double c[3] = {bl[3 * p], bl[3 * p + 1], bl[3 * p + 2]};
          for (size_t a = i0; a < i1; ++a) {
            if (!vis_active[a] || off0[vis[a].s] < 0) continue;
            const double* Wa = &W[a * (size_t)dv * 3];
            const int oa = off0[vis[a].s];
            for (int k = 0; k < 3; ++k)
              for (int r = 0; r < dv; ++r) c[k] -= Wa[3 * r + k] * x[oa + r];
          }
          if (has_scale)
            for (size_t a = i0; a < i1; ++a)
              if (vis_active[a])
                for (int k = 0; k < 3; ++k) c[k] -= Ws[a * 3 + k] * x[off_s];
          const double* Di = &Dinv[(size_t)9 * p];
          for (int a = 0; a < 3; ++a) xl[3 * p + a] = Di[3 * a] * c[0] + Di[3 * a + 1] * c[1] + Di[3 * a + 2] * c[2];
        }
        i0 = i1;
      }
    }
    return true;
  }
  void apply_update() {
    for (size_t k = 0; k < st.size(); ++k) {
      if (off0[k] >= 0) {
        if (pvr) inc_pvr(st[k], &x[off0[k]]);
        else inc_pr(st[k], &x[off0[k]]);
      }
      if (off1[k] >= 0) inc_v(st[k], &x[off1[k]]);
      if (off2[k] >= 0) inc_bias(st[k], &x[off2[k]]);
    }
    if (points_free)
      for (size_t p = 0; p < pt_active.size(); ++p)
        if (pt_active[p])
          for (int k = 0; k < 3; ++k) X[3 * p + k] += xl[3 * p + k];
    if (has_scale) sc += x[off_s];  // VertexScale::oplusImpl
    if (has_g) {                    // VertexGThetaXYRwI::oplusImpl: RwI <- RwI Exp((dx, dy, 0))
      const double w[3] = {x[off_g], x[off_g + 1], 0.0};
      qwI = qnormalized(qmul(qwI, so3_exp_q(w)));
    }
  }
  double compute_scale() const {
    double s = 0;
    for (int j = 0; j < np; ++j) s += x[j] * (lambda * x[j] + b[j]);
    if (points_free)
      for (size_t p = 0; p < pt_active.size(); ++p)
        if (pt_active[p])
          for (int k = 0; k < 3; ++k) s += xl[3 * p + k] * (lambda * xl[3 * p + k] + bl[3 * p + k]);
    return s;
  }
  bool terminate() const { return stop && *stop; }
  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve: the control flow lives in ONE place, the callback driver of
  // lm_oracle.cc (orc_lm_optimize) — or the driver a test installs with orc_set_lm_driver, i.e. the REFERENCE's own solve() /
  // optimize() compiled unchanged in oracle/_ref (tests/test_oracle_ref.py runs every BA driver both ways, bit for bit).  The
  // callbacks are this graph's operations; g2o's solution vector is [poses | free landmarks], which is what x() / b() /
  // hessian_diag() present to computeScale / computeLambdaInit.
  std::vector<int> lm_pts;
  std::vector<double> lm_x, lm_b;
  int optimize(int iterations) {
    x.assign(np, 0.0);
    lm_pts.clear();
    if (points_free)
      for (size_t p = 0; p < pt_active.size(); ++p)
        if (pt_active[p]) lm_pts.push_back((int)p);
    OrcLmCallbacks cb;
    cb.ctx = this;
    cb.n = np + 3 * (int)lm_pts.size();
    cb.errors = [](void* c) {
      Graph* g = (Graph*)c;
      g->compute_active_errors();
      return g->active_robust_chi2();
    };
    cb.build = [](void* c) { ((Graph*)c)->build_system(); };
    cb.solve = [](void* c, double l) {
      Graph* g = (Graph*)c;
      g->lambda = l;
      const bool ok2 = g->solve_system();
      if ((int)g->x.size() != g->np) g->x.assign(g->np, 0.0);
      if (g->points_free && g->xl.size() != g->X.size()) g->xl.assign(g->X.size(), 0.0);
      return ok2 ? 1 : 0;
    };
    cb.update = [](void* c) { ((Graph*)c)->apply_update(); };
    cb.push = [](void* c) {
      Graph* g = (Graph*)c;
      g->st_bak = g->st;
      g->X_bak = g->X;
      g->sc_bak = g->sc;
      g->qwI_bak = g->qwI;
    };
    cb.pop = [](void* c) {
      Graph* g = (Graph*)c;
      g->st = g->st_bak;
      g->X = g->X_bak;
      g->sc = g->sc_bak;
      g->qwI = g->qwI_bak;
    };
    cb.discard_top = [](void*) {};
    cb.x = [](void* c) {
      Graph* g = (Graph*)c;
      g->lm_x.assign(g->x.begin(), g->x.begin() + g->np);
      for (int p : g->lm_pts)
        for (int k = 0; k < 3; ++k) g->lm_x.push_back(g->xl[3 * (size_t)p + k]);
      return (const double*)g->lm_x.data();
    };
    cb.b = [](void* c) {
      Graph* g = (Graph*)c;
      g->lm_b.assign(g->b.begin(), g->b.begin() + g->np);
      for (int p : g->lm_pts)
        for (int k = 0; k < 3; ++k) g->lm_b.push_back(g->bl[3 * (size_t)p + k]);
      return (const double*)g->lm_b.data();
    };
    cb.hessian_diag = [](void* c, int j) {
      Graph* g = (Graph*)c;
      if (j < g->np) return g->H[(size_t)j * g->np + j];
      const int q = j - g->np;
      return g->Hll[9 * (size_t)g->lm_pts[q / 3] + 4 * (q % 3)];
    };
    cb.terminate = [](void* c) { return ((Graph*)c)->terminate() ? 1 : 0; };
    double st5[5] = {0, 0, 0, 0, 0};
    const int n = (g_lm_driver ? g_lm_driver : orc_lm_optimize)(&cb, iterations, user_lambda, st5);
    if (n > 0) lambda = st5[3];
    total_iters += n;
    return n;
  }
};


// Sigma^-1 (IMUPreIntegratorBase::GetProcessedInfoij / InfoijPRV, OdomPreIntegrator.h:119-138)
std::vector<double> info_from_sigma(const double* S) {
  std::vector<double> I(81);
  if (!inverse(S, 9, I.data())) std::fill(I.begin(), I.end(), std::numeric_limits<double>::quiet_NaN());
  return I;
}

// J^T (w Omega) J' for the explicit marginal (g2otypes.h:36-254)
void jtoj(const double* Ja, int lda, int ca, int na, const double* Om, int D, double w, const double* Jb, int ldb, int cb,
          int nb, double* out, int ldo, int r0, int c0, bool add) {
  for (int a = 0; a < na; ++a)
    for (int c = 0; c < nb; ++c) {
      double s = 0;
      for (int i = 0; i < D; ++i) {
        double t = 0;
        for (int j = 0; j < D; ++j) t += (w * Om[i * D + j]) * Jb[j * ldb + cb + c];
        s += Ja[i * lda + ca + a] * t;
      }
      if (add) out[(r0 + a) * ldo + c0 + c] += s;
      else out[(r0 + a) * ldo + c0 + c] = s;
    }
}

}  // namespace

// RobustKernelHuber as the edges use it (delta^2 kept in float): rho(e) and rho'(e)
extern "C" void orc_huber(double delta, double e, double rho[2]) {
  Huber k;
  k.set(delta);
  k.rho(e, rho);
}

// GraphOperator::Chi2LargeSetLevel's decision (optimizer/optimizer_ba/g2o_graph_operator.h:13-31): the edge goes to level 1 when its
// chi2 exceeds rat_th_chi2 * chi2_sig5_[dim_freedom], a FLOAT product of the 5 % chi-square table compared against the double chi2
extern "C" int orc_chi2_large_level(double chi2, int dim_freedom, float rat_th_chi2) {
  static const float chi2_sig5[16] = {0,       3.841f,  5.991f,  7.815f,  9.488f,  11.070f, 12.592f, 14.067f,
                                      15.507f, 16.919f, 18.307f, 19.675f, 21.026f, 22.362f, 23.685f, 24.996f};
  return chi2 > rat_th_chi2 * chi2_sig5[dim_freedom] ? 1 : 0;
}

// the SO3 helpers of so3_oracle.h for the tests (oracle/_ref compiles common/so3_extra.h itself): op 0 exp(w) -> unit quaternion
// (w, x, y, z); 1 Exp(w) -> R; 2 log of a quaternion; 3 Log(R); 4 JacobianR(w); 5 JacobianRInv(w); 6 normalizeRotationM(R)
extern "C" void orc_so3(int op, const double* in, double* out) {
  auto put = [&](const M3& m) {
    for (int k = 0; k < 9; ++k) out[k] = m.m[k];
  };
  M3 R;
  if (op == 3 || op == 6)
    for (int k = 0; k < 9; ++k) R.m[k] = in[k];
  if (op == 0) {
    const Quat q = so3_exp_q(in);
    out[0] = q.w, out[1] = q.x, out[2] = q.y, out[3] = q.z;
  } else if (op == 1) {
    put(so3_Exp(in));
  } else if (op == 2) {
    so3_log_q(qnormalized({in[0], in[1], in[2], in[3]}), out);
  } else if (op == 3) {
    so3_log_q(qnormalized(mquat(R)), out);
  } else if (op == 4) {
    put(so3_Jr(in));
  } else if (op == 5) {
    put(so3_JrInv(in));
  } else if (op == 6) {
    put(normalize_rot(R));
  }
}

// test hook: run every BA driver of this file (pose optimisation, local / global BA) through another LM driver with the
// OrcLmDriver signature — oracle/_ref's ref_lm_optimize, the reference's own solve() / optimize() compiled unchanged
extern "C" void orc_set_lm_driver(OrcLmDriver d) { g_lm_driver = d; }
extern "C" void orc_set_marg_dump(OrcMargDump* d) { g_marg_dump = d; }

// camera Project() alone (pinned against the reference's own function text compiled in oracle/_ref, tests/test_oracle_ref.py)
extern "C" void orc_cam_project(const OrcCamera* cam, const double P[3], float uv[2], double* J /* 2x3 or NULL */) {
  cam_project(cam_from_c(*cam), P, &uv[0], &uv[1], J);
}

// the skyline Cholesky above, for the other oracle translation units (posegraph_oracle.cc)
bool orc_chol_solve_skyline(std::vector<double>& A, int n, const double* b, double* x) { return chol_solve(A, n, b, x); }

extern "C" {

int orc_inverse(const double* A, int n, double* Ainv) { return inverse(A, n, Ainv) ? 0 : -1; }

void orc_edge_reproject(const OrcCamera* cam, const OrcNavState* ns, const double Xw[3], const float obs[3], int stereo,
                        double e[3], double* J_pose, double* J_point, double* depth) {
  const Cam c = cam_from_c(*cam);
  const NS s = from_c(*ns);
  const double d = reproj_error(c, s, Xw, obs, stereo != 0, e);
  if (depth) *depth = d;
  if (J_pose || J_point) {
    double Jp[9], Jr[9], JX[9];
    reproj_jac(c, s, Xw, stereo != 0, Jp, Jr, JX);
    if (J_pose)
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) {
          J_pose[6 * k + j] = Jp[3 * k + j];
          J_pose[6 * k + 3 + j] = Jr[3 * k + j];
        }
    if (J_point) memcpy(J_point, JX, 72);
  }
}

void orc_edge_navstate(const OrcNavState* nsi, const OrcNavState* nsj, const OrcImuPreint* pre, const double gw[3],
                       int order, double e[9], double* Ji, double* Jj, double* Jb) {
  const NS a = from_c(*nsi), b = from_c(*nsj);
  navstate_error(a, b, *pre, gw, order == 1, e);
  if (Ji && Jj && Jb) navstate_jac(a, b, *pre, gw, order == 1, e, Ji, Jj, Jb);
}

void orc_navstate_oplus(OrcNavState* ns, int kind, const double* dx) {
  NS s = from_c(*ns);
  if (kind == 0) inc_pr(s, dx);
  if (kind == 1) inc_pvr(s, dx);
  if (kind == 2) inc_v(s, dx);
  if (kind == 3) inc_bias(s, dx);
  to_c(s, ns);
}

// EdgeReprojectPRS / PRSStereo (g2otypes.h:321-541, MODE_OPT_VAR = 1, NV = 3): the point vertex holds the UNSCALED
// position Xh, Xw = scale * Xh; J_point = (Jproj Rcw) * scale, J_scale = (Jproj Rcw) * Xh (:517-521), the pose block is
// the PR edge's at Xw.  Restated for the next round's device side (GlobalBundleAdjustmentNavStatePRV with bScaleOpt,
// src/Optimizer.cc:843-851, 1132-1200); no device code uses it yet.
void orc_edge_reproject_scale(const OrcCamera* cam, const OrcNavState* ns, const double Xh[3], double scale_est,
                              const float obs[3], int stereo, double e[3], double* J_pose, double* J_point, double* J_scale) {
  const double Xw[3] = {Xh[0] * scale_est, Xh[1] * scale_est, Xh[2] * scale_est};
  double JX[9];
  orc_edge_reproject(cam, ns, Xw, obs, stereo, e, J_pose, (J_point || J_scale) ? JX : nullptr, nullptr);
  if (J_scale)
    for (int r = 0; r < 3; ++r) J_scale[r] = JX[3 * r] * Xh[0] + JX[3 * r + 1] * Xh[1] + JX[3 * r + 2] * Xh[2];
  if (J_point)
    for (int k = 0; k < 9; ++k) J_point[k] = JX[k] * scale_est;
}

// VertexGThetaXYRwI (g2otypes.h:674-698): RwI such that gw = RwI * GI, GI = (0, 0, |gw|); 2-dim update on the right.
void orc_gdir_init(const double gw[3], double q_wI[4]) {
  const double n = std::sqrt(gw[0] * gw[0] + gw[1] * gw[1] + gw[2] * gw[2]);
  const double g[3] = {gw[0] / n, gw[1] / n, gw[2] / n};
  double a[3] = {-g[1], g[0], 0.0};  // (0,0,1) x g
  const double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const double th = std::acos(g[2]);
  // Eigen 3.3 normalized() returns a zero vector unchanged: gw parallel OR anti-parallel to the z axis gives RwI = I
  // (for the anti-parallel case RwI * GI = -gw: the reference's behaviour, reproduced)
  const double inv = na > 0 ? 1.0 / na : 1.0;
  const double w[3] = {a[0] * inv * th, a[1] * inv * th, a[2] * inv * th};
  const Quat q = so3_exp_q(w);
  q_wI[0] = q.w; q_wI[1] = q.x; q_wI[2] = q.y; q_wI[3] = q.z;
}
void orc_gdir_oplus(double q_wI[4], const double d[2]) {
  const double w[3] = {d[0], d[1], 0.0};
  const Quat q = qnormalized(qmul({q_wI[0], q_wI[1], q_wI[2], q_wI[3]}, so3_exp_q(w)));
  q_wI[0] = q.w; q_wI[1] = q.x; q_wI[2] = q.y; q_wI[3] = q.z;
}
// EdgeNavStatePRVG (EdgeNavStateI<6>, g2otypes.h:725-884): the PRV edge with gw = RwI * GI and the extra 9x2 block
// JG (:868-876): rows P = RiT dt^2/2 RwI GI^[:, 0:2], rows V = RiT dt RwI GI^[:, 0:2], rows R = 0.
void orc_edge_navstate_g(const OrcNavState* nsi, const OrcNavState* nsj, const OrcImuPreint* pre, const double q_wI[4],
                         const double GI[3], double e[9], double* Ji, double* Jj, double* Jb, double* JG) {
  const M3 RwI = qmat({q_wI[0], q_wI[1], q_wI[2], q_wI[3]});
  double gw[3];
  mulv(RwI, GI, gw);
  orc_edge_navstate(nsi, nsj, pre, gw, 1, e, Ji, Jj, Jb);
  if (JG) {
    const NS a = from_c(*nsi);
    const M3 RiT = tr(qmat(a.q));
    const M3 A = mul(RiT, mul(RwI, hat(GI)));  // 3x3, first two columns used
    const double dt = pre->dt;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 2; ++c) {
        JG[2 * r + c] = A.m[3 * r + c] * (dt * dt / 2.0);      // rows 0..2: P
        JG[2 * (3 + r) + c] = 0.0;                             // rows 3..5: R
        JG[2 * (6 + r) + c] = A.m[3 * r + c] * dt;             // rows 6..8: V
      }
  }
}

void orc_edge_prior_pvr(const OrcNavState* ns, const OrcNavState* prior, double e[15], double* Jpvr) {
  const NS s = from_c(*ns), p = from_c(*prior);
  prior_error(s, s, p, e);
  if (Jpvr) {
    double Jb[90];
    prior_jac(s, p, e, Jpvr, Jb);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Optimizer::PoseOptimization — mode 0: src/Optimizer.cc:1611-1874, mode 1: include/Optimizer.h:208-816
int orc_pose_optimization(const OrcPoseOptProblem* pb, const OrcCamera* cam, const double* Xw, const float* obs,
                          const float* inv_sigma2, const uint8_t* flags, OrcPoseOptResult* res, uint8_t* outlier,
                          double* chi2_out) {
  memset(res, 0, sizeof(*res));
  res->cur = pb->cur;
  res->last = pb->last;
  const int E = pb->edge_end - pb->edge_begin;
  const bool imu_mode = pb->mode == 1;
  const bool fixed_last = !pb->last_has_prior;
  Graph g;
  g.pvr = imu_mode;
  g.cam = cam_from_c(*cam);
  memcpy(g.gw, pb->gw, 24);
  const NS nsj0 = from_c(pb->cur), nsl0 = from_c(pb->last);
  g.st = {nsj0};
  g.fix0 = {0};
  g.has_vb = {(uint8_t)imu_mode};
  g.fix_vb = {0};
  if (imu_mode) {
    g.st.push_back(nsl0);
    g.fix0.push_back(fixed_last);
    g.has_vb.push_back(1);
    g.fix_vb.push_back(fixed_last);
  }
  bool bodom_edge = false;
  int i_imu = -1, i_bias = -1, i_prior = -1;
  if (imu_mode) {
    if (pb->preint.dt != 0) {
      bodom_edge = true;
      DenseEdge e;
      e.type = 0; e.si = 1; e.sj = 0; e.pre = &pb->preint; e.D = 9;
      e.info = info_from_sigma(pb->preint.SigmaPVR);
      if (fixed_last) {
        for (double& v : e.info) v *= 1e-2;
        e.rk.set(std::sqrt(16.919));
      }
      i_imu = (int)g.den.size();
      g.den.push_back(e);
    }
    {
      DenseEdge e;
      e.type = 1; e.si = 1; e.sj = 0; e.D = 6;
      e.info.assign(36, 0.0);
      const double dtij = pb->preint.dt != 0 ? pb->preint.dt : pb->dt_frames;
      for (int k = 0; k < 6; ++k) {
        const double w = (k < 3 ? pb->inv_sigma_bg2 : pb->inv_sigma_ba2) / dtij;
        e.info[7 * k] = fixed_last ? w * 1e-2 : w;
      }
      if (fixed_last) e.rk.set(std::sqrt(12.592));
      i_bias = (int)g.den.size();
      g.den.push_back(e);
    }
    if (!fixed_last) {
      DenseEdge e;
      e.type = 2; e.si = 1; e.sj = 1; e.D = 15;
      e.prior = from_c(pb->prior);
      e.info.assign(pb->prior_info, pb->prior_info + 225);
      e.rk.set(std::sqrt(25.0));
      i_prior = (int)g.den.size();
      g.den.push_back(e);
    }
  }
  const float deltaMono = (float)std::sqrt(5.991), deltaStereo = (float)std::sqrt(7.815);
  g.X.resize((size_t)3 * E);
  g.vis.resize(E);
  for (int i = 0; i < E; ++i) {
    const int gi = pb->edge_begin + i;
    memcpy(&g.X[3 * i], Xw + 3 * gi, 24);
    VisEdge& e = g.vis[i];
    e.s = 0; e.p = i;
    memcpy(e.obs, obs + 3 * gi, 12);
    e.w = (double)inv_sigma2[gi];
    e.stereo = flags[gi] & ORC_EDGE_STEREO;
    e.close = flags[gi] & ORC_EDGE_CLOSE;
    e.level = 0;
    e.rk.set(e.stereo ? (double)deltaStereo : (double)deltaMono);
    outlier[gi] = 0;
    chi2_out[gi] = 0;
  }
  const int nInitial = E;
  res->n_initial = nInitial;
  if (nInitial < 3 && !(imu_mode && pb->no_mps)) {
    res->n_inliers = 0;
    return 0;
  }
  const float chi2Mono = 5.991f, chi2Stereo = 7.815f;
  const size_t n_edges_total = g.vis.size() + g.den.size();
  int nBad = 0;
  for (int it = 0; it < 4; ++it) {
    if (!imu_mode || !bodom_edge) {
      g.st[0] = nsj0;
      if (imu_mode && !fixed_last) g.st[1] = nsl0;
    }
    g.initialize();
    g.optimize(10);
    const float chi2close = 1.5 * chi2Mono;
    nBad = 0;
    for (int i = 0; i < E; ++i) {
      VisEdge& e = g.vis[i];
      const int gi = pb->edge_begin + i;
      if (imu_mode || outlier[gi]) g.vis_error(e);
      const float chi2 = (float)e.chi2;
      bool bad;
      if (e.stereo) bad = chi2 > chi2Stereo;
      else if (imu_mode) bad = chi2 > (e.close ? chi2close : chi2Mono) || !(g.vis_depth(e) > 0.);
      else bad = chi2 > chi2Mono;
      outlier[gi] = bad;
      e.level = bad ? 1 : 0;
      nBad += bad;
      if (it == 2) e.rk.on = false;
    }
    if (n_edges_total < 10) break;
  }
  if (imu_mode && nInitial - nBad < 30) {  // rescue pass (include/Optimizer.h:619-648)
    nBad = 0;
    for (int i = 0; i < E; ++i) {
      VisEdge& e = g.vis[i];
      const int gi = pb->edge_begin + i;
      g.vis_error(e);
      if (e.chi2 < (e.stereo ? (double)24.f : (double)18.f)) {
        e.level = 0;
        outlier[gi] = 0;
      } else
        nBad++;
    }
  }
  for (int i = 0; i < E; ++i) chi2_out[pb->edge_begin + i] = g.vis[i].chi2;
  to_c(g.st[0], &res->cur);
  if (imu_mode) to_c(g.st[1], &res->last);
  res->n_inliers = nInitial - nBad;
  res->iterations = g.total_iters;
  res->lambda_final = g.lambda;
  g.initialize();  // chi2_final: robust chi2 of the final level-0 set with the stored errors
  res->chi2_final = g.active_robust_chi2();
  if (imu_mode && pb->compute_marg) {
    // kExactRobust: recompute errors, re-linearise every edge at the final estimate (include/Optimizer.h:126-206, 671-728)
    if (i_imu >= 0) g.den_error(g.den[i_imu]);
    g.den_error(g.den[i_bias]);
    double C[225];
    memset(C, 0, sizeof(C));
    double Ji[81], Jj[81], Jb[54], r[2];
    const DenseEdge* eI = i_imu >= 0 ? &g.den[i_imu] : nullptr;
    const DenseEdge& eB = g.den[i_bias];
    double wI = 1;
    if (eI) {
      navstate_jac(g.st[1], g.st[0], *eI->pre, g.gw, false, eI->err, Ji, Jj, Jb);
      eI->rk.rho(eI->chi2, r);
      wI = r[1];
      jtoj(Jj, 9, 0, 9, eI->info.data(), 9, wI, Jj, 9, 0, 9, C, 15, 0, 0, false);
    }
    eB.rk.rho(eB.chi2, r);
    const double wB = r[1];
    for (int k = 0; k < 6; ++k) C[(9 + k) * 15 + 9 + k] = wB * eB.info[7 * k];  // Xj^T (w Omega) Xj, Xj = I
    for (int i = 0; i < E; ++i) {
      VisEdge& e = g.vis[i];
      if (e.level) continue;
      double Jp[9], Jr[9], JX[9];
      reproj_jac(g.cam, g.st[0], &g.X[3 * e.p], e.stereo, Jp, Jr, JX);
      e.rk.rho(e.chi2, r);
      const double w = r[1] * e.w;
      const int DE = e.stereo ? 3 : 2;
      double J[3][9];
      memset(J, 0, sizeof(J));
      for (int k = 0; k < DE; ++k)
        for (int c = 0; c < 3; ++c) {
          J[k][c] = Jp[3 * k + c];
          J[k][6 + c] = Jr[3 * k + c];
        }
      for (int a = 0; a < 9; ++a)
        for (int c = 0; c < 9; ++c) {
          double s = 0;
          for (int k = 0; k < DE; ++k) s += J[k][a] * (w * J[k][c]);
          C[a * 15 + c] += s;
        }
    }
    OrcMargDump* dump = g_marg_dump;
    if (dump) {
      dump->filled = 1;
      dump->has_imu = eI ? 1 : 0;
      dump->fixed_last = fixed_last ? 1 : 0;
      dump->n_vis = E;
      if (eI) memcpy(dump->info_imu, eI->info.data(), sizeof(dump->info_imu));
      dump->delta_imu = eI && eI->rk.on ? eI->rk.delta : -1.0;
      memcpy(dump->info_bias, eB.info.data(), sizeof(dump->info_bias));
      dump->delta_bias = eB.rk.on ? eB.rk.delta : -1.0;
      dump->delta_prior = -1.0;
      for (int i = 0; i < E && i < dump->cap; ++i) {
        dump->level[i] = g.vis[i].level;
        dump->delta[i] = g.vis[i].rk.on ? g.vis[i].rk.delta : -1.0;
      }
      memcpy(dump->C, C, sizeof(C));
    }
    if (!fixed_last) {
      DenseEdge& eP = g.den[i_prior];
      g.den_error(eP);
      double CL[225], CCL[225];
      memset(CL, 0, sizeof(CL));
      memset(CCL, 0, sizeof(CCL));
      if (eI) {
        jtoj(Ji, 9, 0, 9, eI->info.data(), 9, wI, Ji, 9, 0, 9, CL, 15, 0, 0, false);
        jtoj(Ji, 9, 0, 9, eI->info.data(), 9, wI, Jb, 6, 0, 6, CL, 15, 0, 9, false);
        jtoj(Jb, 6, 0, 6, eI->info.data(), 9, wI, Jb, 6, 0, 6, CL, 15, 9, 9, false);
        for (int a = 0; a < 9; ++a)
          for (int c = 0; c < 6; ++c) CL[(9 + c) * 15 + a] = CL[a * 15 + 9 + c];
      }
      for (int k = 0; k < 6; ++k) CL[(9 + k) * 15 + 9 + k] += wB * eB.info[7 * k];  // Xi^T (w Omega) Xi, Xi = -I
      double Jpp[135], Jpb[90];
      prior_jac(g.st[1], eP.prior, eP.err, Jpp, Jpb);
      eP.rk.rho(eP.chi2, r);
      const double wP = r[1];
      jtoj(Jpp, 9, 0, 9, eP.info.data(), 15, wP, Jpp, 9, 0, 9, CL, 15, 0, 0, true);
      jtoj(Jpb, 6, 0, 6, eP.info.data(), 15, wP, Jpb, 6, 0, 6, CL, 15, 9, 9, true);
      jtoj(Jpp, 9, 0, 9, eP.info.data(), 15, wP, Jpb, 6, 0, 6, CL, 15, 0, 9, true);
      for (int a = 0; a < 9; ++a)
        for (int c = 0; c < 6; ++c) CL[(9 + c) * 15 + a] = CL[a * 15 + 9 + c];
      if (eI) {
        jtoj(Jj, 9, 0, 9, eI->info.data(), 9, wI, Ji, 9, 0, 9, CCL, 15, 0, 0, false);
        jtoj(Jj, 9, 0, 9, eI->info.data(), 9, wI, Jb, 6, 0, 6, CCL, 15, 0, 9, false);
      }
      for (int k = 0; k < 6; ++k) CCL[(9 + k) * 15 + 9 + k] = -(wB * eB.info[7 * k]);  // Xj^T (w Omega) Xi = -w Omega
      if (dump) {
        memcpy(dump->info_prior, eP.info.data(), sizeof(dump->info_prior));
        dump->delta_prior = eP.rk.on ? eP.rk.delta : -1.0;
        memcpy(dump->CL, CL, sizeof(CL));
        memcpy(dump->CCL, CCL, sizeof(CCL));
      }
      double Cinv[225], T[225];
      if (!inverse(CL, 15, Cinv))
        for (double& v : Cinv) v = std::numeric_limits<double>::quiet_NaN();
      for (int a = 0; a < 15; ++a)
        for (int c = 0; c < 15; ++c) {
          double s = 0;
          for (int k = 0; k < 15; ++k) s += CCL[a * 15 + k] * Cinv[k * 15 + c];
          T[a * 15 + c] = s;
        }
      for (int a = 0; a < 15; ++a)
        for (int c = 0; c < 15; ++c) {
          double s = 0;
          for (int k = 0; k < 15; ++k) s += T[a * 15 + k] * CCL[c * 15 + k];
          C[a * 15 + c] -= s;
        }
    }
    memcpy(res->marg_cov_inv, C, sizeof(C));
    res->prior_set = 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Optimizer::LocalBundleAdjustmentNavStatePRV (src/Optimizer.cc:133-666): optimiser set-up to the err/err_end guard
}  // extern "C"
namespace {
// Graph of Optimizer::LocalBundleAdjustmentNavStatePRV (src/Optimizer.cc:133-520): vertices, IMU + bias edges, visual edges
bool build_lba_graph(const OrcBaProblem* pb, const OrcCamera* cam, Graph& g, int optit[2]) {
  g.pvr = false;
  g.cam = cam_from_c(*cam);
  memcpy(g.gw, pb->gw, 24);
  g.points_free = true;
  if (pb->visual_only) { optit[0] = 5; optit[1] = 10; }
  else if (pb->large) { optit[0] = 2; optit[1] = 2; g.user_lambda = 1e-2; }
  else { optit[0] = 4; optit[1] = 6; g.user_lambda = 1e0; }
  const int K = pb->n_states, P = pb->n_points, E = pb->n_edges;
  g.st.resize(K); g.fix0.resize(K); g.has_vb.resize(K); g.fix_vb.resize(K);
  bool anyfree = false;
  for (int k = 0; k < K; ++k) {
    g.st[k] = from_c(pb->states[k]);
    g.fix0[k] = pb->state_flags[k] & 1;
    g.has_vb[k] = (pb->state_flags[k] & 2) && !pb->visual_only;
    g.fix_vb[k] = (pb->state_flags[k] & 4) != 0;
    anyfree |= !g.fix0[k];
  }
  if (!anyfree) return false;
  const float thHuberPRV = (float)std::sqrt(16.919), thHuberBias = (float)std::sqrt(12.592);
  // GlobalBundleAdjustmentNavStatePRV (src/Optimizer.cc:903-983): information x 1e-2 where the previous keyframe's BIAS
  // vertex is fixed, kernels on every inertial edge iff bRobust; the local BA keys both on a fixed previous keyframe
  const bool global = pb->global_ba & 1, g_robust = pb->global_ba & 2;
  if (global) g.user_lambda = 0;
  for (int m = 0; m < (pb->visual_only ? 0 : pb->n_imu); ++m) {
    const int i = pb->imu_i[m], j = pb->imu_j[m];
    const bool bfixedkf = global ? (!g.has_vb[i] || g.fix_vb[i]) : (bool)g.fix0[i];
    const bool kernel = global ? g_robust : (bfixedkf || pb->rec_init);
    const OrcImuPreint& pre = pb->preint[m];
    if (pre.dt != 0) {
      DenseEdge e;
      e.type = 0; e.si = i; e.sj = j; e.pre = &pre; e.D = 9;
      e.info = info_from_sigma(pre.SigmaPRV);
      if (bfixedkf) for (double& v : e.info) v *= 1e-2;
      if (kernel) e.rk.set((double)thHuberPRV);
      g.den.push_back(e);
    }
    DenseEdge e;
    e.type = 1; e.si = i; e.sj = j; e.D = 6;
    double dtij = pre.dt != 0 ? pre.dt : pb->imu_dt_kf[m];
    if (dtij <= (double)1e-6f) dtij = 15;
    e.info.assign(36, 0.0);
    for (int k = 0; k < 6; ++k) {
      const double w = (k < 3 ? pb->inv_sigma_bg2 : pb->inv_sigma_ba2) / dtij;
      e.info[7 * k] = bfixedkf ? w * 1e-2 : w;
    }
    if (kernel) e.rk.set((double)thHuberBias);
    g.den.push_back(e);
  }
  g.X.assign(pb->points, pb->points + (size_t)3 * P);
  const float chi2Mono = 5.991f;
  // src/Optimizer.cc:361-362 (PRV: sqrt of the float 5.991f) vs :2069-2070 (visual LocalBundleAdjustment: sqrt(5.991))
  // global BA: thHuber2D = sqrt(5.99) (:1042)
  const float thHuberMono = global ? (float)std::sqrt(5.99) : pb->visual_only ? (float)std::sqrt(5.991) : std::sqrt(chi2Mono);
  const float thHuberStereo = (float)std::sqrt(7.815);
  g.vis.resize(E);
  for (int i = 0; i < E; ++i) {
    VisEdge& e = g.vis[i];
    e.s = pb->edge_state[i]; e.p = pb->edge_point[i];
    memcpy(e.obs, pb->obs + 3 * i, 12);
    e.w = (double)pb->inv_sigma2[i];
    e.stereo = pb->edge_flags[i] & ORC_EDGE_STEREO;
    e.close = pb->edge_flags[i] & ORC_EDGE_CLOSE;
    e.level = (pb->edge_flags[i] & ORC_EDGE_LEVEL1) ? 1 : 0;
    if (!(pb->edge_flags[i] & ORC_EDGE_NOKERNEL) && !(global && !g_robust))
      e.rk.set(e.stereo ? (double)thHuberStereo : (double)thHuberMono);
  }
  return true;
}
}  // namespace
extern "C" {

// Optimizer::GlobalBundleAdjustmentNavStatePRV (src/Optimizer.cc:771-1342) with bScaleOpt = false and no IMU initiator:
// one optimize(nIterations) from g2o's own initial lambda (1e-5 max diag), no level classification.
int orc_global_ba_prv(const OrcBaProblem* pb_in, const OrcCamera* cam, int n_iterations, int robust, OrcNavState* states_out,
                      double* points_out, double* edge_chi2, OrcBaResult* res) {
  const int K = pb_in->n_states, P = pb_in->n_points, E = pb_in->n_edges;
  memset(res, 0, sizeof(*res));
  for (int k = 0; k < K; ++k) states_out[k] = pb_in->states[k];
  if (points_out) memcpy(points_out, pb_in->points, sizeof(double) * 3 * (size_t)P);
  OrcBaProblem pb = *pb_in;
  pb.global_ba = 1 | (robust ? 2 : 0);
  pb.large = 0; pb.rec_init = 0;
  Graph g;
  int optit[2];
  if (!build_lba_graph(&pb, cam, g, optit)) return 0;
  g.initialize();
  if (g.np == 0) return 0;
  g.compute_active_errors();
  res->err0 = g.active_robust_chi2();
  const int it = g.optimize(n_iterations);
  res->iterations[0] = it;
  g.compute_active_errors();
  res->err_end = g.active_robust_chi2();
  res->lambda_final = g.lambda;
  res->accepted = 1;
  if (edge_chi2)
    for (int i = 0; i < E; ++i) edge_chi2[i] = g.vis[i].chi2;
  for (int k = 0; k < K; ++k) to_c(g.st[k], &states_out[k]);
  if (points_out) memcpy(points_out, g.X.data(), sizeof(double) * 3 * (size_t)P);
  return it;
}

// The same with bScaleOpt = true (System::FinalGBA, src/System.cc:28-29): VertexScale seeded with 1, visual edges
// EdgeReprojectPRS[Stereo]; on return the points are multiplied by the recovered scale (src/Optimizer.cc:1258-1336).
// Restated ahead of its device side; scale_out receives the vertex estimate.
int orc_global_ba_prv_scale(const OrcBaProblem* pb_in, const OrcCamera* cam, int n_iterations, int robust,
                            OrcNavState* states_out, double* points_out, double* edge_chi2, OrcBaResult* res,
                            double* scale_out) {
  const int K = pb_in->n_states, P = pb_in->n_points, E = pb_in->n_edges;
  memset(res, 0, sizeof(*res));
  *scale_out = 1.0;
  for (int k = 0; k < K; ++k) states_out[k] = pb_in->states[k];
  if (points_out) memcpy(points_out, pb_in->points, sizeof(double) * 3 * (size_t)P);
  OrcBaProblem pb = *pb_in;
  pb.global_ba = 1 | (robust ? 2 : 0);
  pb.large = 0; pb.rec_init = 0;
  Graph g;
  int optit[2];
  if (!build_lba_graph(&pb, cam, g, optit)) return 0;
  g.has_scale = true;
  g.sc = 1.0;
  g.initialize();
  g.compute_active_errors();
  res->err0 = g.active_robust_chi2();
  const int it = g.optimize(n_iterations);
  res->iterations[0] = it;
  g.compute_active_errors();
  res->err_end = g.active_robust_chi2();
  res->lambda_final = g.lambda;
  res->accepted = 1;
  *scale_out = g.sc;
  if (edge_chi2)
    for (int i = 0; i < E; ++i) edge_chi2[i] = g.vis[i].chi2;
  for (int k = 0; k < K; ++k) to_c(g.st[k], &states_out[k]);
  if (points_out)
    for (size_t k = 0; k < (size_t)3 * P; ++k) points_out[k] = g.sc * g.X[k];
  return it;
}
// The IMU initialiser's call (pimu_initiator != nullptr, src/Odom/IMUInitialization.cpp:475: bRobust = false, bScaleOpt =
// false): keyframe 0 keeps only PR fixed (V / Bias free, :825-831 — the caller's state_flags say so), gravity-direction
// vertex seeded from gw (:852-865), EdgeNavStatePRVG on every pair (:955-957), one prior-bias edge from a fixed copy of the
// earliest keyframe's bias with information invSigma / sum(dt) (:866-900, 1026-1054).  states[0] is the earliest keyframe.
// gw_io: in = the initialiser's gravity, out = RwI * GI.
int orc_global_ba_prv_init(const OrcBaProblem* pb_in, const OrcCamera* cam, int n_iterations, double gw_io[3],
                           OrcNavState* states_out, double* points_out, double* edge_chi2, OrcBaResult* res) {
  const int K = pb_in->n_states, P = pb_in->n_points, E = pb_in->n_edges;
  memset(res, 0, sizeof(*res));
  for (int k = 0; k < K; ++k) states_out[k] = pb_in->states[k];
  if (points_out) memcpy(points_out, pb_in->points, sizeof(double) * 3 * (size_t)P);
  OrcBaProblem pb = *pb_in;
  pb.global_ba = 1;
  pb.large = 0; pb.rec_init = 0;
  memcpy(pb.gw, gw_io, 24);
  Graph g;
  int optit[2];
  if (!build_lba_graph(&pb, cam, g, optit)) return 0;
  g.has_g = true;
  const double gn = std::sqrt(gw_io[0] * gw_io[0] + gw_io[1] * gw_io[1] + gw_io[2] * gw_io[2]);
  g.GI[0] = 0; g.GI[1] = 0; g.GI[2] = gn;
  double q[4];
  orc_gdir_init(gw_io, q);
  g.qwI = {q[0], q[1], q[2], q[3]};
  // prior-bias edge on the earliest keyframe; sum_dt over every keyframe pair as the reference accumulates it (:940-947)
  double sum_dt = 0;
  for (int m = 0; m < pb.n_imu; ++m) {
    double dtij = pb.preint[m].dt != 0 ? pb.preint[m].dt : pb.imu_dt_kf[m];
    if (dtij <= (double)1e-6f) dtij = 15;
    sum_dt += dtij;
  }
  if (K > 0 && g.has_vb[0]) {
    DenseEdge e;
    e.type = 3; e.si = 0; e.sj = 0; e.D = 6;
    e.prior = g.st[0];
    e.info.assign(36, 0.0);
    for (int k = 0; k < 6; ++k) e.info[7 * k] = (k < 3 ? pb.inv_sigma_bg2 : pb.inv_sigma_ba2) / sum_dt;
    g.den.push_back(e);
  }
  g.initialize();
  g.compute_active_errors();
  res->err0 = g.active_robust_chi2();
  const int it = g.optimize(n_iterations);
  res->iterations[0] = it;
  g.compute_active_errors();
  res->err_end = g.active_robust_chi2();
  res->lambda_final = g.lambda;
  res->accepted = 1;
  g.gravity(gw_io);
  if (edge_chi2)
    for (int i = 0; i < E; ++i) edge_chi2[i] = g.vis[i].chi2;
  for (int k = 0; k < K; ++k) to_c(g.st[k], &states_out[k]);
  if (points_out) memcpy(points_out, g.X.data(), sizeof(double) * 3 * (size_t)P);
  return it;
}
// One damped step of the LBA graph with a gravity-direction vertex seeded from gw: x_pose [np] (its two entries LAST).
int orc_ba_debug_step_gdir(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, const double gw[3], double* x_pose,
                           double* x_points, double* chi2) {
  Graph g;
  int optit[2];
  if (!build_lba_graph(pb, cam, g, optit)) return -1;
  g.has_g = true;
  g.GI[0] = 0; g.GI[1] = 0; g.GI[2] = std::sqrt(gw[0] * gw[0] + gw[1] * gw[1] + gw[2] * gw[2]);
  double q[4];
  orc_gdir_init(gw, q);
  g.qwI = {q[0], q[1], q[2], q[3]};
  g.initialize();
  g.compute_active_errors();
  if (chi2) *chi2 = g.active_robust_chi2();
  g.build_system();
  g.lambda = lambda;
  if (!g.solve_system()) return -2;
  memcpy(x_pose, g.x.data(), sizeof(double) * g.np);
  if (x_points) memcpy(x_points, g.xl.data(), sizeof(double) * g.xl.size());
  return g.np;
}
// One damped step of the LBA graph with a scale vertex at estimate scale0: x_pose [np] (the scale is the LAST entry).
int orc_ba_debug_step_scale(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, double scale0, double* x_pose,
                            double* x_points, double* chi2) {
  Graph g;
  int optit[2];
  if (!build_lba_graph(pb, cam, g, optit)) return -1;
  g.has_scale = true;
  g.sc = scale0;
  g.initialize();
  g.compute_active_errors();
  if (chi2) *chi2 = g.active_robust_chi2();
  g.build_system();
  g.lambda = lambda;
  if (!g.solve_system()) return -2;
  memcpy(x_pose, g.x.data(), sizeof(double) * g.np);
  if (x_points) memcpy(x_points, g.xl.data(), sizeof(double) * g.xl.size());
  return g.np;
}

// One damped Gauss-Newton step of the LBA graph at the input estimate (build + Schur solve with the given lambda):
// x_pose [np] in Hessian index order, x_points [P][3].  For the tests' cross-check against a dense solve of the
// full normal equations.  Returns np.
int orc_ba_debug_step(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, double* x_pose, double* x_points,
                      double* chi2) {
  Graph g;
  int optit[2];
  if (!build_lba_graph(pb, cam, g, optit)) return -1;
  g.initialize();
  g.compute_active_errors();
  if (chi2) *chi2 = g.active_robust_chi2();
  g.build_system();
  g.lambda = lambda;
  if (!g.solve_system()) return -2;
  memcpy(x_pose, g.x.data(), g.np * 8);
  memcpy(x_points, g.xl.data(), g.xl.size() * 8);
  return g.np;
}

// Reduced camera system [S | bschur] and pose rhs b of a (partial) problem: lambda is added to the point blocks always
// and to the pose diagonal only when lambda_on_poses (rank 0 of a landmark-sharded run).  Returns np.
int orc_ba_debug_system(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, int lambda_on_poses, double* S,
                        double* bs, double* b, double* chi2) {
  Graph g;
  int optit[2];
  if (!build_lba_graph(pb, cam, g, optit)) return -1;
  g.initialize();
  g.compute_active_errors();
  if (chi2) *chi2 = g.active_robust_chi2();
  g.build_system();
  g.lambda = lambda;
  g.lambda_pose = lambda_on_poses ? lambda : 0.0;
  g.solve_system();
  memcpy(S, g.S_out.data(), g.S_out.size() * 8);
  memcpy(bs, g.bs_out.data(), g.bs_out.size() * 8);
  memcpy(b, g.b.data(), g.b.size() * 8);
  return g.np;
}

int orc_local_ba_prv(const OrcBaProblem* pb, const OrcCamera* cam, OrcNavState* states_out, double* points_out,
                     double* edge_chi2, uint8_t* erase, OrcBaResult* res) {
  memset(res, 0, sizeof(*res));
  const int K = pb->n_states, P = pb->n_points, E = pb->n_edges;
  for (int k = 0; k < K; ++k) states_out[k] = pb->states[k];
  memcpy(points_out, pb->points, (size_t)P * 24);
  memset(erase, 0, E);
  Graph g;
  int optit[2];
  if (!build_lba_graph(pb, cam, g, optit)) return 0;
  const float chi2Mono = 5.991f;
  const bool vo = pb->visual_only != 0;  // Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1876-2307)
  // GraphOperator::Chi2LargeSetLevel (g2o_graph_operator.h:23-40), rat 100 — PRV version only (src/Optimizer.cc:534-536)
  if (!vo)
    for (VisEdge& e : g.vis) {
      g.vis_error(e);
      if (orc_chi2_large_level(e.chi2, e.stereo ? 3 : 2, 100.f)) e.level = 1;
    }
  g.initialize();
  g.compute_active_errors();
  const float err = (float)g.active_robust_chi2();
  res->err0 = err;
  res->iterations[0] = g.optimize(optit[0]);
  auto is_bad = [&](VisEdge& e) {
    if (e.stereo) return e.chi2 > 7.815 || !(g.vis_depth(e) > 0.);
    if (vo) return e.chi2 > 5.991 || !(g.vis_depth(e) > 0.);  // :2198, 2238: no "close point" relaxation
    return e.chi2 > (e.close ? 1.5 * chi2Mono : (double)chi2Mono) || !(g.vis_depth(e) > 0.);
  };
  {  // bDoMore
    for (VisEdge& e : g.vis) {
      if (is_bad(e)) e.level = 1;
      e.rk.on = false;
    }
    g.initialize();
    res->iterations[1] = g.optimize(optit[1]);
  }
  const float err_end = (float)g.active_robust_chi2();
  res->err_end = err_end;
  res->lambda_final = g.lambda;
  for (int i = 0; i < E; ++i) edge_chi2[i] = g.vis[i].chi2;
  // "FAIL LOCAL-INERTIAL BA" guard of the PRV version (:663-666); the visual version has none
  if (!vo && (2 * err < err_end || std::isnan(err) || std::isnan(err_end)) && !pb->large) {
    res->accepted = 0;
    return 0;
  }
  res->accepted = 1;
  int n_erase = 0;
  for (int i = 0; i < E; ++i) {
    const bool bad = is_bad(g.vis[i]);
    erase[i] = bad;
    n_erase += bad;
  }
  res->n_erase = n_erase;
  for (int k = 0; k < K; ++k) to_c(g.st[k], &states_out[k]);
  memcpy(points_out, g.X.data(), (size_t)P * 24);
  return 0;
}

}  // extern "C"

// ================================================================================================================
// Optimizer::OptimizeSim3 (src/Optimizer.cc:2689-2920): EdgeReproject<2, 6, 3, MODE> with MODE 1 (PRS) / 2 (PRSInv)
// (src/Odom/g2otypes.h:346-399 GetTcw_wX, :400-406 computeError, :439-541 linearizeOplus; USE_P_PLUS_RDP defined)
namespace {
struct Sim3Edge {
  double e[2], Jp[12], Js[2];  // J_pose 2 x 6 (dp | dphi), J_scale 2 x 1
};
void sim3_edge(const Cam& c, const NS& ns, double sc, const double Xh[3], const float obs[2], bool inverse, bool jac, Sim3Edge& o) {
  M3 Rwb = qmat(ns.q);
  double twb[3] = {ns.p[0], ns.p[1], ns.p[2]};
  const M3 Rbw_var = Rwb;  // "*pRbw = ns.getRwb()" (:361): the un-transposed rotation
  double sfac = sc;
  if (inverse) {
    sfac = 1. / sfac;
    Rwb = tr(Rwb);
    double t[3];
    mulv(Rwb, twb, t);
    for (int i = 0; i < 3; ++i) twb[i] = -t[i];
  }
  const M3 Rcw = mul(c.Rcb, tr(Rwb));
  double tcw[3];
  mulv(Rcw, twb, tcw);
  for (int i = 0; i < 3; ++i) tcw[i] = -tcw[i] + c.tcb[i];
  if (inverse)
    for (int i = 0; i < 3; ++i) tcw[i] *= sfac;
  const double Xw[3] = {Xh[0] * sfac, Xh[1] * sfac, Xh[2] * sfac};
  double Pc[3];
  mulv(Rcw, Xw, Pc);
  for (int i = 0; i < 3; ++i) Pc[i] += tcw[i];
  float u, v;
  double Jc[6];
  cam_project(c, Pc, &u, &v, jac ? Jc : nullptr);
  o.e[0] = (double)obs[0] - (double)u;
  o.e[1] = (double)obs[1] - (double)v;
  if (!jac) return;
  M3 Jproj = {{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  for (int i = 0; i < 6; ++i) Jproj.m[i] = -Jc[i];
  M3 JdP = mul(Jproj, scale(c.Rcb, -1.0));
  M3 JdR;
  if (!inverse) {
    const double d[3] = {Xw[0] - ns.p[0], Xw[1] - ns.p[1], Xw[2] - ns.p[2]};
    double Paux[3];
    mulv(tr(qmat(ns.q)), d, Paux);
    JdR = mul(mul(Jproj, c.Rcb), hat(Paux));
  } else {
    JdR = mul(mul(Jproj, scale(Rcw, -1.0)), hat(Xw));
  }
  const M3 JX = mul(Jproj, Rcw);  // _jacobianOplus[0] before the chain factors
  double Js[2];
  for (int r = 0; r < 2; ++r) Js[r] = JX.m[3 * r] * Xh[0] + JX.m[3 * r + 1] * Xh[1] + JX.m[3 * r + 2] * Xh[2];
  if (inverse) {
    JdP = mul(JdP, scale(Rbw_var, -1.0));  // J_twb_tbw (:524)
    for (int r = 0; r < 2; ++r) {
      const double jt = Jproj.m[3 * r] * tcw[0] + Jproj.m[3 * r + 1] * tcw[1] + Jproj.m[3 * r + 2] * tcw[2];
      Js[r] = (Js[r] + jt) * (-sfac * sfac);  // (:538-539)
    }
  }
  for (int r = 0; r < 2; ++r) {
    for (int k = 0; k < 3; ++k) {
      o.Jp[6 * r + k] = JdP.m[3 * r + k];
      o.Jp[6 * r + 3 + k] = JdR.m[3 * r + k];
    }
    o.Js[r] = Js[r];
  }
}
}  // namespace

extern "C" {

void orc_edge_sim3(const OrcCamera* cam, const OrcNavState* ns, double scale, const double Xh[3], const float obs[2],
                   int inverse, double e[2], double* J_pose, double* J_scale) {
  Sim3Edge o;
  sim3_edge(cam_from_c(*cam), from_c(*ns), scale, Xh, obs, inverse != 0, J_pose || J_scale, o);
  e[0] = o.e[0];
  e[1] = o.e[1];
  if (J_pose) memcpy(J_pose, o.Jp, sizeof(o.Jp));
  if (J_scale) memcpy(J_scale, o.Js, sizeof(o.Js));
}

int orc_optimize_sim3(const OrcSim3Problem* pb, const OrcCamera* cam_c, const double* Xc1, const double* Xc2, const float* obs1,
                      const float* obs2, const float* inv_sigma2_1, const float* inv_sigma2_2, OrcSim3Result* res,
                      uint8_t* keep, double* chi2_12, double* chi2_21) {
  memset(res, 0, sizeof(*res));
  res->ns = pb->ns;
  res->scale = pb->scale;
  const Cam cam = cam_from_c(*cam_c);
  const int M = pb->m_end - pb->m_begin, b0 = pb->m_begin;
  const int n = pb->fix_scale ? 6 : 7;
  NS ns = from_c(pb->ns);
  double sc = pb->scale;
  Huber rk;
  rk.set((double)std::sqrt(pb->th2));  // const float deltaHuber = sqrt(th2) (:2752)
  const double th2 = (double)pb->th2;
  std::vector<uint8_t> active(M, 1);
  std::vector<double> c12(M, 0.0), c21(M, 0.0);
  for (int i = 0; i < M; ++i) keep[b0 + i] = 1;
  res->n_corr = M;
  double lambda = 0;
  int total_iters = 0;
  auto errors = [&]() {  // computeActiveErrors + activeRobustChi2
    double tot = 0;
    for (int i = 0; i < M; ++i) {
      if (!active[i]) continue;
      const int g = b0 + i;
      Sim3Edge o;
      double r[2];
      sim3_edge(cam, ns, sc, Xc2 + 3 * g, obs1 + 2 * g, false, false, o);
      const double w1 = (double)inv_sigma2_1[g];
      c12[i] = o.e[0] * (w1 * o.e[0]) + o.e[1] * (w1 * o.e[1]);
      rk.rho(c12[i], r);
      tot += r[0];
      sim3_edge(cam, ns, sc, Xc1 + 3 * g, obs2 + 2 * g, true, false, o);
      const double w2 = (double)inv_sigma2_2[g];
      c21[i] = o.e[0] * (w2 * o.e[0]) + o.e[1] * (w2 * o.e[1]);
      rk.rho(c21[i], r);
      tot += r[0];
    }
    return tot;
  };
  std::vector<double> H(49), bvec(7), x(7);
  auto add_edge = [&](const Sim3Edge& o, double w, double chi) {
    double r[2];
    rk.rho(chi, r);
    double J[2][7];
    for (int k = 0; k < 2; ++k) {
      for (int a = 0; a < 6; ++a) J[k][a] = o.Jp[6 * k + a];
      J[k][6] = o.Js[k];
    }
    const double ww = r[1] * w;
    for (int a = 0; a < n; ++a) {
      double sb = 0;
      for (int k = 0; k < 2; ++k) sb += J[k][a] * (-(w * o.e[k]) * r[1]);
      bvec[a] += sb;
      for (int cc = 0; cc < n; ++cc) {
        double h = 0;
        for (int k = 0; k < 2; ++k) h += (J[k][a] * ww) * J[k][cc];
        H[a * n + cc] += h;
      }
    }
  };
  auto build = [&]() {  // linearizeOplus + constructQuadraticForm at the current estimate (errors must be current)
    std::fill(H.begin(), H.end(), 0.0);
    std::fill(bvec.begin(), bvec.end(), 0.0);
    for (int i = 0; i < M; ++i) {
      if (!active[i]) continue;
      const int g = b0 + i;
      Sim3Edge o;
      sim3_edge(cam, ns, sc, Xc2 + 3 * g, obs1 + 2 * g, false, true, o);
      add_edge(o, (double)inv_sigma2_1[g], c12[i]);
      sim3_edge(cam, ns, sc, Xc1 + 3 * g, obs2 + 2 * g, true, true, o);
      add_edge(o, (double)inv_sigma2_2[g], c21[i]);
    }
  };
  // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve through the callback driver (lm_oracle.cc / the test hook)
  struct Ctx {
    std::function<double()> errors;
    std::function<void()> build, update;
    std::function<bool(double)> solve;
    NS* ns;
    double* sc;
    NS ns_bak;
    double sc_bak;
    std::vector<double>*H, *bvec, *x;
    int n;
  } ctx;
  ctx.errors = errors;
  ctx.build = build;
  ctx.update = [&]() {
    inc_pr(ns, x.data());
    if (n == 7) sc += x[6];
  };
  ctx.solve = [&](double lam) {
    std::vector<double> A(H.begin(), H.begin() + n * n);
    for (int j = 0; j < n; ++j) A[j * n + j] += lam;
    return chol_solve(A, n, bvec.data(), x.data());
  };
  ctx.ns = &ns; ctx.sc = &sc; ctx.H = &H; ctx.bvec = &bvec; ctx.x = &x; ctx.n = n;
  OrcLmCallbacks cb;
  cb.ctx = &ctx;
  cb.n = n;
  cb.errors = [](void* c) { return ((Ctx*)c)->errors(); };
  cb.build = [](void* c) { ((Ctx*)c)->build(); };
  cb.solve = [](void* c, double l) { return ((Ctx*)c)->solve(l) ? 1 : 0; };
  cb.update = [](void* c) { ((Ctx*)c)->update(); };
  cb.push = [](void* c) { ((Ctx*)c)->ns_bak = *((Ctx*)c)->ns; ((Ctx*)c)->sc_bak = *((Ctx*)c)->sc; };
  cb.pop = [](void* c) { *((Ctx*)c)->ns = ((Ctx*)c)->ns_bak; *((Ctx*)c)->sc = ((Ctx*)c)->sc_bak; };
  cb.discard_top = [](void*) {};
  cb.x = [](void* c) { return (const double*)((Ctx*)c)->x->data(); };
  cb.b = [](void* c) { return (const double*)((Ctx*)c)->bvec->data(); };
  cb.hessian_diag = [](void* c, int j) { return (*((Ctx*)c)->H)[j * ((Ctx*)c)->n + j]; };
  cb.terminate = nullptr;
  auto optimize = [&](int iterations) {
    std::fill(x.begin(), x.end(), 0.0);
    double st5[5] = {0, 0, 0, 0, 0};
    const int its = (g_lm_driver ? g_lm_driver : orc_lm_optimize)(&cb, iterations, 0.0, st5);
    if (its > 0) lambda = st5[3];
    total_iters += its;
  };
  optimize(5);
  int nBad = 0;
  for (int i = 0; i < M; ++i)
    if (c12[i] > th2 || c21[i] > th2) {  // (:2858-2868)
      keep[b0 + i] = 0;
      active[i] = 0;
      nBad++;
    }
  res->n_bad = nBad;
  const int nMore = nBad > 0 ? 10 : 5;
  auto store_chi = [&]() {
    for (int i = 0; i < M; ++i) {
      if (chi2_12) chi2_12[b0 + i] = c12[i];
      if (chi2_21) chi2_21[b0 + i] = c21[i];
    }
  };
  res->iterations = total_iters;
  res->lambda_final = lambda;
  if (M - nBad < 10) {  // (:2878) the Sim3 estimate is NOT written back
    store_chi();
    res->n_inliers = 0;
    return 0;
  }
  optimize(nMore);
  int nIn = 0;
  for (int i = 0; i < M; ++i) {
    if (!active[i]) continue;
    if (c12[i] > th2 || c21[i] > th2) keep[b0 + i] = 0;
    else nIn++;
  }
  store_chi();
  {
    double tot = 0, r[2];
    for (int i = 0; i < M; ++i)
      if (active[i]) {
        rk.rho(c12[i], r); tot += r[0];
        rk.rho(c21[i], r); tot += r[0];
      }
    res->chi2_final = tot;
  }
  to_c(ns, &res->ns);
  res->scale = sc;
  res->n_inliers = nIn;
  res->iterations = total_iters;
  res->lambda_final = lambda;
  return nIn;
}

}  // extern "C"

// Device-side edges of the bundle-adjustment path (fp64), shared by poseopt.cu and ba.cu.
//   EdgeReproject<DE,DV,2>      src/Odom/g2otypes.h:400-541  (pinhole projection rounded to float,
//                               common/camera_models/camera_pinhole.h:70-106)
//   EdgeNavStateI<NV>           src/Odom/g2otypes.h:725-884
//   EdgeNavStatePriorPVRBias    src/Odom/g2otypes.cpp:84-124
//   NavState::IncSmall          src/Odom/NavState.h:47-82 (USE_P_PLUS_RDP: p <- p + R dp, R <- R Exp(dphi))
//   RobustKernelHuber           optimizer/g2o/g2o/core/robust_kernel_impl.cpp:65-91 (delta^2 kept as float)
// Expressions are written in the same order as the CPU oracle so that, with -fmad=false, results differ only by
// the device libm's last-ulp differences and by reduction order.
#pragma once
#include "common.cuh"
#include "so3.cuh"

namespace vieo {

struct NavS {
  Vec3 p;
  Quat q;
  Vec3 v, bg, ba, dbg, dba;
};
__device__ __forceinline__ Vec3 ld3(const double* a) { return {a[0], a[1], a[2]}; }
__device__ __forceinline__ void st3(double* a, const Vec3& v) {
  a[0] = v.x;
  a[1] = v.y;
  a[2] = v.z;
}
__device__ __forceinline__ NavS ns_load(const VieoNavState& s) {
  NavS n;
  n.p = ld3(s.p);
  n.q = {s.q[0], s.q[1], s.q[2], s.q[3]};
  n.v = ld3(s.v);
  n.bg = ld3(s.bg);
  n.ba = ld3(s.ba);
  n.dbg = ld3(s.dbg);
  n.dba = ld3(s.dba);
  return n;
}
__device__ __forceinline__ void ns_store(const NavS& n, VieoNavState& s) {
  st3(s.p, n.p);
  s.q[0] = n.q.w; s.q[1] = n.q.x; s.q[2] = n.q.y; s.q[3] = n.q.z;
  st3(s.v, n.v);
  st3(s.bg, n.bg);
  st3(s.ba, n.ba);
  st3(s.dbg, n.dbg);
  st3(s.dba, n.dba);
}
__device__ __forceinline__ Quat q_conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }

__device__ __forceinline__ void ns_inc_pr(NavS& s, const double* d) {
  const Vec3 Rd = m3_mulv(q_matrix(s.q), ld3(d));
  s.p = v3_add(s.p, Rd);
  s.q = q_normalized(q_mul(s.q, so3_exp_q(ld3(d + 3))));
}
__device__ __forceinline__ void ns_inc_pvr(NavS& s, const double* d) {
  const Vec3 Rd = m3_mulv(q_matrix(s.q), ld3(d));
  s.p = v3_add(s.p, Rd);
  s.v = v3_add(s.v, ld3(d + 3));
  s.q = q_normalized(q_mul(s.q, so3_exp_q(ld3(d + 6))));
}
__device__ __forceinline__ void ns_inc_v(NavS& s, const double* d) { s.v = v3_add(s.v, ld3(d)); }
__device__ __forceinline__ void ns_inc_bias(NavS& s, const double* d) {
  s.dbg = v3_add(s.dbg, ld3(d));
  s.dba = v3_add(s.dba, ld3(d + 3));
}

// Huber kernel state packed as (delta double, delta^2 float); delta == 0: no kernel.
__device__ __forceinline__ void huber_rho(double delta, double e, double& rho0, double& rho1) {
  const double dsqr = (double)(float)(delta * delta);
  if (delta == 0 || e <= dsqr) {
    rho0 = e;
    rho1 = 1.;
  } else {
    const double sq = sqrt(e);
    rho0 = 2 * sq * delta - dsqr;
    rho1 = delta / sq;
  }
}

// Camera pose of a body state: Rcw = Rcb Rwb^T, tcw = -(Rcw pwb) + tcb (g2otypes.h:352-366)
struct CamPose {
  Mat3 Rcw, Rwb;
  Vec3 tcw, pwb;
};
struct CamK {
  double fx, fy, cx, cy, bf;  // float intrinsics promoted to double
  int model, num_k;           // VIEO_CAM_*
  double dist[8];             // float distortion coefficients promoted to double
  Mat3 Rcb;
  Vec3 tcb;
};
__host__ __device__ __forceinline__ void cam_set(CamK& k, const VieoCamera& c) {
  k.fx = (double)c.fx; k.fy = (double)c.fy; k.cx = (double)c.cx; k.cy = (double)c.cy; k.bf = (double)c.bf;
  k.model = c.model; k.num_k = c.num_k;
  for (int i = 0; i < 8; ++i) k.dist[i] = (double)c.dist[i];
  for (int i = 0; i < 9; ++i) k.Rcb.m[i] = c.Rcb[i];
  k.tcb = {c.tcb[0], c.tcb[1], c.tcb[2]};
}
// Camera::Project rounded to float pixels and its 2x3 Jacobian (J may be null): pinhole camera_pinhole.h:70-106,
// radtan camera_radtan.h:61-129, KB8 camera_kb8.h:68-157 — formulas as the reference writes them.
__device__ __forceinline__ void cam_project(const CamK& c, const Vec3& P, float& u, float& v, double* J) {
  if (c.model == VIEO_CAM_RADTAN) {
    const double* k = c.dist;
    const double* p = k + c.num_k;
    const double invz = 1 / P.z;
    const double x = P.x * invz, y = P.y * invz;
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
      const double du_dx = c.fx * invz * (fd + fd2 * x2 + 2 * (p[0] * y + 3 * p[1] * x));
      const double du_dy = c.fx * invz * (fd2 * xy + 2 * (p[0] * x + p[1] * y));
      const double du_dz = -(x * du_dx + y * du_dy);
      const double dv_dx = du_dy * c.fy / c.fx;
      const double dv_dy = c.fy * invz * (fd + fd2 * y2 + 2 * (p[1] * x + 3 * p[0] * y));
      const double dv_dz = -(x * dv_dx + y * dv_dy);
      J[0] = du_dx; J[1] = du_dy; J[2] = du_dz; J[3] = dv_dx; J[4] = dv_dy; J[5] = dv_dz;
    }
    const double xd = x * fd + 2 * p[0] * xy + p[1] * (r2 + 2 * x2);
    const double yd = y * fd + 2 * p[1] * xy + p[0] * (r2 + 2 * y2);
    u = (float)(c.fx * xd * 1.0 + c.cx);
    v = (float)(c.fy * yd * 1.0 + c.cy);
    return;
  }
  if (c.model == VIEO_CAM_KB8) {
    const double x = P.x, y = P.y;
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r = sqrt(r2);
    if (r > (double)1e-5f) {
      const double k1 = c.dist[0], k2 = c.dist[1], k3 = c.dist[2], k4 = c.dist[3];
      const double z = P.z;
      const double theta = atan2(r, z), theta2 = theta * theta;
      double thetad = k4 * theta2;
      thetad += k3; thetad *= theta2; thetad += k2; thetad *= theta2; thetad += k1; thetad *= theta2; thetad += 1;
      thetad *= theta;
      const double mx = x * thetad / r, my = y * thetad / r;
      u = (float)(c.fx * mx * 1.0 + c.cx);
      v = (float)(c.fy * my * 1.0 + c.cy);
      if (J) {
        const double invr = 1. / r, d_r_d_x = x * invr, d_r_d_y = y * invr;
        const double tmp = 1. / (z * z + r2);
        const double d_thetad_x = d_r_d_x * z * tmp, d_thetad_y = d_r_d_y * z * tmp;
        double dd = 9.0 * k4 * theta2;
        dd += 7.0 * k3; dd *= theta2; dd += 5.0 * k2; dd *= theta2; dd += 3.0 * k1; dd *= theta2; dd += 1.0;
        const double invr2 = invr * invr;
        J[0] = c.fx * (x * r * dd * d_thetad_x + y2 * thetad / r) * invr2;
        J[1] = c.fx * x * (dd * d_thetad_y * r - y * thetad / r) * invr2;
        J[2] = -c.fx * x * dd * tmp;
        J[3] = J[1] * c.fy / c.fx;
        J[4] = c.fy * (y * r * dd * d_thetad_y + x2 * thetad / r) * invr2;
        J[5] = -c.fy * y * dd * tmp;
      }
      return;
    }
  }
  const double invz = 1. / P.z;
  u = (float)(c.fx * P.x * invz + c.cx);
  v = (float)(c.fy * P.y * invz + c.cy);
  if (J) {
    const double invz2 = invz * invz;
    J[0] = c.fx * invz; J[1] = 0; J[2] = -c.fx * P.x * invz2;
    J[3] = 0; J[4] = c.fy * invz; J[5] = -c.fy * P.y * invz2;
  }
}
__device__ __forceinline__ CamK cam_load(const VieoCamera& c) {
  CamK k;
  k.fx = (double)c.fx; k.fy = (double)c.fy; k.cx = (double)c.cx; k.cy = (double)c.cy; k.bf = (double)c.bf;
  k.model = c.model; k.num_k = c.num_k;
#pragma unroll
  for (int i = 0; i < 8; ++i) k.dist[i] = (double)c.dist[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) k.Rcb.m[i] = c.Rcb[i];
  k.tcb = ld3(c.tcb);
  return k;
}
__device__ __forceinline__ CamPose cam_pose(const CamK& c, const NavS& s) {
  CamPose P;
  P.Rwb = q_matrix(s.q);
  P.Rcw = m3_mul(c.Rcb, m3_t(P.Rwb));
  const Vec3 t = m3_mulv(P.Rcw, s.p);
  P.tcw = {-t.x + c.tcb.x, -t.y + c.tcb.y, -t.z + c.tcb.z};
  P.pwb = s.p;
  return P;
}
// e = obs - pi(Pc) with the projection rounded to float; returns depth (EdgeReproject::GetDepth)
__device__ __forceinline__ double reproj_error(const CamK& c, const CamPose& P, const Vec3& X, const float* obs, bool stereo,
                                               double e[3]) {
  Vec3 Pc = m3_mulv(P.Rcw, X);
  Pc = v3_add(Pc, P.tcw);
  float u, v;
  cam_project(c, Pc, u, v, nullptr);
  e[0] = (double)obs[0] - (double)u;
  e[1] = (double)obs[1] - (double)v;
  e[2] = stereo ? (double)obs[2] - ((double)u - c.bf / Pc.z) : 0.0;
  return P.Rcw.m[6] * X.x + P.Rcw.m[7] * X.y + P.Rcw.m[8] * X.z + P.tcw.z;
}
// Jacobians (rows 0..DE-1): Jp = de/d(dp), Jr = de/d(dphi), JX = de/dX
__device__ __forceinline__ void reproj_jac(const CamK& c, const CamPose& P, const Vec3& X, bool stereo, Mat3& Jp, Mat3& Jr,
                                           Mat3& JX) {
  Vec3 Pc = m3_mulv(P.Rcw, X);
  Pc = v3_add(Pc, P.tcw);
  const double invz = 1 / Pc.z, invz_2 = invz * invz;
  Mat3 Jproj = m3_zero();
  {
    float u, v;
    double Jc[6];
    cam_project(c, Pc, u, v, Jc);
#pragma unroll
    for (int i = 0; i < 6; ++i) Jproj.m[i] = -Jc[i];
  }
  if (stereo) {
    Jproj.m[6] = Jproj.m[0];
    Jproj.m[7] = Jproj.m[1];
    Jproj.m[8] = Jproj.m[2] - c.bf * invz_2;
  }
  Jp = m3_mul(Jproj, m3_scale(c.Rcb, -1.0));
  const Vec3 Paux = m3_tmulv(P.Rwb, v3_sub(X, P.pwb));
  Jr = m3_mul(m3_mul(Jproj, c.Rcb), m3_hat(Paux));
  JX = m3_mul(Jproj, P.Rcw);
}

// ---- EdgeNavStateI<NV>: prv = residual/column order P,R,V (NV=5) else P,V,R (NV=3) ---------------------------
__device__ __forceinline__ Mat3 ld_m3(const double* a) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = a[i];
  return r;
}
// the fields of VieoImuPreint the edges read, without the two 9x9 covariances (for staging in shared memory)
struct VieoImuPreintLite {
  double Rij[9], vij[3], pij[3], Jgp[9], Jap[9], Jgv[9], Jav[9], JgR[9], dt;
};
template <class Pre>
__device__ inline void navstate_error(const NavS& si, const NavS& sj, const Pre& m, const Vec3& gw, bool prv, double e[9]) {
  const Mat3 RiT = m3_t(q_matrix(si.q));
  const int idR = prv ? 3 : 6, idV = 9 - idR;
  const double dt = m.dt;
  Vec3 a = {sj.p.x - si.p.x - si.v.x * dt - gw.x * (dt * dt / 2), sj.p.y - si.p.y - si.v.y * dt - gw.y * (dt * dt / 2),
            sj.p.z - si.p.z - si.v.z * dt - gw.z * (dt * dt / 2)};
  Vec3 b = m3_mulv(RiT, a);
  Vec3 t1 = m3_mulv(ld_m3(m.Jgp), si.dbg), t2 = m3_mulv(ld_m3(m.Jap), si.dba);
  e[0] = b.x - (m.pij[0] + t1.x + t2.x);
  e[1] = b.y - (m.pij[1] + t1.y + t2.y);
  e[2] = b.z - (m.pij[2] + t1.z + t2.z);
  const Vec3 w = m3_mulv(ld_m3(m.JgR), si.dbg);
  const Quat qm = q_normalized(q_mul(q_normalized(q_from_matrix(ld_m3(m.Rij))), so3_exp_q(w)));
  const Quat qij = q_normalized(q_mul(q_conj(si.q), sj.q));
  const Vec3 eR = so3_log_q(q_normalized(q_mul(q_conj(qm), qij)));
  st3(e + idR, eR);
  a = {sj.v.x - si.v.x - gw.x * dt, sj.v.y - si.v.y - gw.y * dt, sj.v.z - si.v.z - gw.z * dt};
  b = m3_mulv(RiT, a);
  t1 = m3_mulv(ld_m3(m.Jgv), si.dbg);
  t2 = m3_mulv(ld_m3(m.Jav), si.dba);
  e[idV + 0] = b.x - (m.vij[0] + t1.x + t2.x);
  e[idV + 1] = b.y - (m.vij[1] + t1.y + t2.y);
  e[idV + 2] = b.z - (m.vij[2] + t1.z + t2.z);
}
__device__ __forceinline__ void setb(double* A, int ld, int r, int c, const Mat3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[ld * (r + i) + c + j] = b.m[3 * i + j];
}
// Ji, Jj: 9x9 row-major (state columns in residual order), Jb: 9x6; e = current error
template <class Pre>
__device__ inline void navstate_jac(const NavS& si, const NavS& sj, const Pre& m, const Vec3& gw, bool prv,
                                    const double e[9], double* Ji, double* Jj, double* Jb) {
  const Mat3 RiT = m3_t(q_matrix(si.q)), Rj = q_matrix(sj.q);
  const int idR = prv ? 3 : 6, idV = 9 - idR;
  const double dt = m.dt;
  for (int i = 0; i < 81; ++i) Ji[i] = Jj[i] = 0;
  for (int i = 0; i < 54; ++i) Jb[i] = 0;
  const Mat3 JgR = ld_m3(m.JgR);
  Vec3 a = {sj.p.x - si.p.x - si.v.x * dt - gw.x * (dt * dt / 2), sj.p.y - si.p.y - si.v.y * dt - gw.y * (dt * dt / 2),
            sj.p.z - si.p.z - si.v.z * dt - gw.z * (dt * dt / 2)};
  Vec3 b = m3_mulv(RiT, a);
  setb(Ji, 9, 0, idR, m3_hat(b));
  setb(Ji, 9, 0, 0, m3_scale(m3_identity(), -1.0));
  setb(Ji, 9, 0, idV, m3_scale(m3_scale(RiT, -1.0), dt));
  setb(Jb, 6, 0, 0, m3_scale(ld_m3(m.Jgp), -1.0));
  setb(Jb, 6, 0, 3, m3_scale(ld_m3(m.Jap), -1.0));
  setb(Jj, 9, 0, 0, m3_mul(RiT, Rj));
  a = {sj.v.x - si.v.x - gw.x * dt, sj.v.y - si.v.y - gw.y * dt, sj.v.z - si.v.z - gw.z * dt};
  b = m3_mulv(RiT, a);
  setb(Ji, 9, idV, idR, m3_hat(b));
  setb(Ji, 9, idV, idV, m3_scale(RiT, -1.0));
  setb(Jb, 6, idV, 0, m3_scale(ld_m3(m.Jgv), -1.0));
  setb(Jb, 6, idV, 3, m3_scale(ld_m3(m.Jav), -1.0));
  setb(Jj, 9, idV, idV, RiT);
  const Vec3 eR = ld3(e + idR);
  const Mat3 Jrinv = so3_JrInv(eR);
  const Mat3 RjTRi = q_matrix(q_normalized(q_mul(q_conj(sj.q), si.q)));
  setb(Ji, 9, idR, idR, m3_scale(m3_mul(Jrinv, RjTRi), -1.0));
  const Vec3 w = m3_mulv(JgR, si.dbg);
  const Mat3 T = m3_mul(m3_mul(m3_mul(m3_scale(Jrinv, -1.0), so3_Exp({-eR.x, -eR.y, -eR.z})), so3_Jr(w)), JgR);
  setb(Jb, 6, idR, 0, T);
  setb(Jj, 9, idR, idR, Jrinv);
}

// ---- EdgeNavStatePriorPVRBias: order P V R bg ba ---------------------------------------------------------------
__device__ inline void prior_error(const NavS& s, const NavS& pr, double e[15]) {
  const Quat qbi = q_conj(pr.q);
  const Vec3 ep = m3_mulv(q_matrix(qbi), v3_sub(s.p, pr.p));
  st3(e, ep);
  st3(e + 3, v3_sub(s.v, pr.v));
  st3(e + 6, so3_log_q(q_normalized(q_mul(qbi, s.q))));
  e[9] = s.bg.x + s.dbg.x - (pr.bg.x + pr.dbg.x);
  e[10] = s.bg.y + s.dbg.y - (pr.bg.y + pr.dbg.y);
  e[11] = s.bg.z + s.dbg.z - (pr.bg.z + pr.dbg.z);
  e[12] = s.ba.x + s.dba.x - (pr.ba.x + pr.dba.x);
  e[13] = s.ba.y + s.dba.y - (pr.ba.y + pr.dba.y);
  e[14] = s.ba.z + s.dba.z - (pr.ba.z + pr.dba.z);
}
// Jpvr 15x9, Jb 15x6
__device__ inline void prior_jac(const NavS& s, const NavS& pr, const double e[15], double* Jpvr, double* Jb) {
  for (int i = 0; i < 135; ++i) Jpvr[i] = 0;
  for (int i = 0; i < 90; ++i) Jb[i] = 0;
  setb(Jpvr, 9, 0, 0, m3_mul(m3_t(q_matrix(pr.q)), q_matrix(s.q)));
  setb(Jpvr, 9, 3, 3, m3_identity());
  setb(Jpvr, 9, 6, 6, so3_JrInv(ld3(e + 6)));
  setb(Jb, 6, 9, 0, m3_identity());
  setb(Jb, 6, 12, 3, m3_identity());
}

// ---- small dense algebra on shared / local arrays (single thread) --------------------------------------------
// Gauss-Jordan inverse with partial pivoting; M (n x n) is destroyed.  false on a zero pivot.
__device__ inline bool dense_inverse(double* M, int n, double* Ai) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Ai[i * n + j] = i == j;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(M[r * n + c]) > fabs(M[piv * n + c])) piv = r;
    if (M[piv * n + c] == 0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        double t = M[piv * n + j]; M[piv * n + j] = M[c * n + j]; M[c * n + j] = t;
        t = Ai[piv * n + j]; Ai[piv * n + j] = Ai[c * n + j]; Ai[c * n + j] = t;
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
// out[r0.., c0..] (+)= Ja^T (w Om) Jb, Ja: D x lda (cols ca..ca+na), Jb: D x ldb (cols cb..cb+nb)
__device__ inline void jtoj(const double* Ja, int lda, int ca, int na, const double* Om, int D, double w, const double* Jb,
                            int ldb, int cb, int nb, double* out, int ldo, int r0, int c0, bool add) {
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

}  // namespace vieo

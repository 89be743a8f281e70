// Motion-only bundle adjustment on the device — Optimizer::PoseOptimization.
//   mode 0  visual, one VertexNavStatePR (6)                          src/Optimizer.cc:1611-1874
//   mode 1  IMU, VertexNavStatePVR (9) + Bias (6) for the frame and, when the last frame carries a prior, for the
//           last frame too (30-dim), EdgeNavStatePVR + EdgeNavStateBias + EdgeNavStatePriorPVRBias, and the
//           kExactRobust marginal prior                                include/Optimizer.h:126-816
// One thread block per frame runs the reference's whole schedule — 4 rounds x optimize(10) of g2o's
// Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:61-189), inlier re-classification between rounds,
// rescue pass, marginalisation — without leaving the SM.  Batched over frames (grid = number of frames) this is the
// throughput form of the tracking thread's hot loop.
//
// Inside a block the work of one LM trial is laid out to keep the dependent fp64 chain short:
//   * evaluate(): warps 0..5 take the visual edges (one edge per thread: residual, chi2, Huber weight, Jacobian,
//     28 partial sums reduced with warp shuffles in a fixed order); warp 7 lane 0 does the inertial edge (residual +
//     9x24 Jacobian strip), warp 6 lane 0 the bias-walk and prior edges — concurrently.  Then all threads form
//     Omega e, J^T (rho' Omega) and finally every entry of H / b (<= 30 x 30) in a fixed summation order.
//   * evaluate() at the trial estimate IS the next iteration's linearisation when the trial is accepted (g2o recomputes
//     the same numbers): H / b are double-buffered in shared memory, a rejected trial simply keeps the old set.
//   * the <= 30-dim Cholesky and the triangular solves run on warp 0 with lanes over matrix entries.
#include <algorithm>

#include "ba_edges.cuh"

namespace vieo {

#ifdef VIEO_PROF
__device__ long long g_po_prof[16];
#define PO_T(var) const long long var = clock64()
#define PO_ACC(slot, t0, t1) \
  if (blockIdx.x == 0) atomicAdd((unsigned long long*)&g_po_prof[slot], (unsigned long long)((t1) - (t0)))
#else
#define PO_T(var)
#define PO_ACC(slot, t0, t1)
#endif

#ifndef VIEO_PO_THREADS
#define VIEO_PO_THREADS 256
#endif
#ifndef VIEO_PO_MIN_CTAS
// two resident CTAs per SM (128 registers, ~6 KB of L1-resident spills per thread in the single-lane inertial / prior
// paths): measured on B200 for 256 problems — 3.37 ms at 254 registers / 1 CTA per SM, 3.08 ms here, 3.20 ms with 128
// threads, 3.35 / 3.74 ms with 128 threads at 168 / 128 registers (profiles/r02e_poseopt_variants.txt) — the kernel is
// bound by the latency of its serial fp64 chains, so the smaller footprint costs nothing and leaves half of every SM's
// register file to the kernels of the other streams
#define VIEO_PO_MIN_CTAS 2
#endif
constexpr int kPoThreads = VIEO_PO_THREADS;
constexpr int kPoWarps = kPoThreads / 32;
constexpr int kPoVisWarps = kPoWarps - 2;  // the first warps: visual edges; then one warp for the bias + prior edges and
constexpr int kPoImuWarp = kPoWarps - 1;   // the last one for the inertial edge
static_assert(kPoThreads >= 96 && kPoThreads % 32 == 0, "k_pose_opt needs at least one visual warp and 70 threads");
constexpr int kPoN = 30;          // largest system: cur PVR 9 + bias 6 + last PVR 9 + bias 6
constexpr int kPoMaxEdges = 4096; // per frame (the reference has N <= nfeatures + a few per camera)

struct PoLin {  // one linearisation: normal equations + robust chi2
  double H[kPoN * kPoN], b[kPoN];
  double chi;
};
struct PoSmem {
  PoLin lin[2];
  double S[kPoN * kPoN];
  double x[kPoN], y[kPoN];
  double red[kPoWarps][32];
  double tot[32];
  double Jimu[9 * 24];   // [Ji (last PVR) | Jj (frame PVR) | Jb (last bias)]
  double Jpri[15 * 15];  // [Jpvr 15x9 | Jb 15x6]
  double AtOimu[24 * 9], AtOpri[15 * 15];
  double oe_imu[9], oe_pri[15], oe_bias[6];
  double info_imu[81], info_prior[225], info_bias[6];
  double err_imu[9], err_bias[6], err_prior[15];
  double chi2_imu, chi2_bias, chi2_prior, r1_imu, r1_bias, r1_prior, rho_imu, rho_bias, rho_prior;
  double C[225], CL[225], CCL[225];
  NavS st[2], bak[2], ini[2];
  CamPose cp;
  VieoImuPreintLite pre;  // the frame's pre-integration staged once (the Sigma blocks are left out)
  double lambda, ni, inv_d;
  int nBad, ok, total_iters, cur, inv_piv;
  uint8_t eflag[kPoMaxEdges];  // bit0: level 1, bit1: kernel removed
};

struct PoCtx {
  const VieoPoseOptProblem* pb;
  CamK cam;
  const double* Xw;
  const float* obs;
  const float* w;
  const uint8_t* flags;
  double* chi2;
  int E, n, dv;
  bool imu_mode, fixed_last, has_imu;
  double delta_mono, delta_stereo, delta_imu, delta_bias, delta_prior;
  Vec3 gw;
  NavS prior;
};

// sum of NV doubles per thread over the first `nwarps` warps, fixed order: lanes by xor tree, then warps in sequence.
// Every thread of the block must call it (two barriers); result in sm.tot[0..NV).
template <int NV>
__device__ __forceinline__ void block_sum(PoSmem& sm, double (&v)[NV], int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    v[k] = a;
  }
  if (lane == 0 && warp < nwarps) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm.red[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double a = 0;
    for (int w = 0; w < nwarps; ++w) a += sm.red[w][threadIdx.x];
    sm.tot[threadIdx.x] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ double edge_delta(const PoCtx& c, const PoSmem& sm, int i) {
  if (sm.eflag[i] & 2) return 0.0;
  return (c.flags[i] & VIEO_EDGE_STEREO) ? c.delta_stereo : c.delta_mono;
}
// e->computeError() + chi2 for visual edge i at the current estimate
__device__ __forceinline__ double vis_chi2(const PoCtx& c, const PoSmem& sm, int i, double e[3], double* depth = nullptr) {
  const bool stereo = c.flags[i] & VIEO_EDGE_STEREO;
  const double d = reproj_error(c.cam, sm.cp, ld3(c.Xw + 3 * (size_t)i), c.obs + 3 * (size_t)i, stereo, e);
  if (depth) *depth = d;
  const double w = (double)c.w[i];
  double chi = 0;
  for (int k = 0; k < (stereo ? 3 : 2); ++k) chi += e[k] * (w * e[k]);
  return chi;
}

// EdgeNavStatePVR Jacobians straight into the 9 x 24 strip [Ji | Jj | Jb] (same arithmetic as navstate_jac, prv = false)
template <class Pre>
__device__ void navstate_jac_pvr24(const NavS& si, const NavS& sj, const Pre& m, const Vec3& gw, const double e[9],
                                   double* J) {
  const Mat3 RiT = m3_t(q_matrix(si.q)), Rj = q_matrix(sj.q);
  const double dt = m.dt;
  const Mat3 JgR = ld_m3(m.JgR);
  // rows / state columns in P, V, R order; column bases: Ji 0, Jj 9, Jb 18.  The strip is zeroed ONCE per problem by the
  // whole block (the blocks written here are always the same ones)
  Vec3 a = {sj.p.x - si.p.x - si.v.x * dt - gw.x * (dt * dt / 2), sj.p.y - si.p.y - si.v.y * dt - gw.y * (dt * dt / 2),
            sj.p.z - si.p.z - si.v.z * dt - gw.z * (dt * dt / 2)};
  Vec3 b = m3_mulv(RiT, a);
  setb(J, 24, 0, 6, m3_hat(b));
  setb(J, 24, 0, 0, m3_scale(m3_identity(), -1.0));
  setb(J, 24, 0, 3, m3_scale(m3_scale(RiT, -1.0), dt));
  setb(J, 24, 0, 18, m3_scale(ld_m3(m.Jgp), -1.0));
  setb(J, 24, 0, 21, m3_scale(ld_m3(m.Jap), -1.0));
  setb(J, 24, 0, 9, m3_mul(RiT, Rj));
  a = {sj.v.x - si.v.x - gw.x * dt, sj.v.y - si.v.y - gw.y * dt, sj.v.z - si.v.z - gw.z * dt};
  b = m3_mulv(RiT, a);
  setb(J, 24, 3, 6, m3_hat(b));
  setb(J, 24, 3, 3, m3_scale(RiT, -1.0));
  setb(J, 24, 3, 18, m3_scale(ld_m3(m.Jgv), -1.0));
  setb(J, 24, 3, 21, m3_scale(ld_m3(m.Jav), -1.0));
  setb(J, 24, 3, 12, RiT);
  const Vec3 eR = ld3(e + 6);
  const Mat3 Jrinv = so3_JrInv(eR);
  const Mat3 RjTRi = q_matrix(q_normalized(q_mul(q_conj(sj.q), si.q)));
  setb(J, 24, 6, 6, m3_scale(m3_mul(Jrinv, RjTRi), -1.0));
  const Vec3 w = m3_mulv(JgR, si.dbg);
  const Mat3 Tm = m3_mul(m3_mul(m3_mul(m3_scale(Jrinv, -1.0), so3_Exp({-eR.x, -eR.y, -eR.z})), so3_Jr(w)), JgR);
  setb(J, 24, 6, 18, Tm);
  setb(J, 24, 6, 15, Jrinv);
}

__device__ __forceinline__ void bias_error(PoSmem& sm) {
  const NavS &a = sm.st[1], &d = sm.st[0];
  sm.err_bias[0] = (d.bg.x + d.dbg.x) - (a.bg.x + a.dbg.x);
  sm.err_bias[1] = (d.bg.y + d.dbg.y) - (a.bg.y + a.dbg.y);
  sm.err_bias[2] = (d.bg.z + d.dbg.z) - (a.bg.z + a.dbg.z);
  sm.err_bias[3] = (d.ba.x + d.dba.x) - (a.ba.x + a.dba.x);
  sm.err_bias[4] = (d.ba.y + d.dba.y) - (a.ba.y + a.dba.y);
  sm.err_bias[5] = (d.ba.z + d.dba.z) - (a.ba.z + a.dba.z);
}
// [Jpvr 15x9 | Jb 15x6] of the prior edge (ld 15) at the current estimate; err_prior must be current.  sm.Jpri is zeroed
// once per problem by the whole block; only the five diagonal blocks are ever written
__device__ __forceinline__ void prior_jac15(const PoCtx& c, PoSmem& sm) {
  setb(sm.Jpri, 15, 0, 0, m3_mul(m3_t(q_matrix(c.prior.q)), q_matrix(sm.st[1].q)));
  setb(sm.Jpri, 15, 3, 3, m3_identity());
  setb(sm.Jpri, 15, 6, 6, so3_JrInv(ld3(sm.err_prior + 6)));
  setb(sm.Jpri, 15, 9, 9, m3_identity());
  setb(sm.Jpri, 15, 12, 12, m3_identity());
}

// Gauss-Jordan inverse with partial pivoting by the whole block: the same operation sequence per entry as the
// single-thread dense_inverse (ba_edges.cuh), so the result is bit-identical; 3 barriers per column instead of n^2
// dependent shared-memory round trips on one lane.  M (n x n, shared) is destroyed, Ai (n x n, shared) receives the
// inverse.  Columns of M already eliminated are left stale (nothing reads them again).  All threads call it; the return
// value is uniform.
__device__ bool block_inverse(PoSmem& sm, double* M, int n, double* Ai) {
  const int t = threadIdx.x;
  for (int i = t; i < n * n; i += kPoThreads) Ai[i] = (i / n == i % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    if (t == 0) {
      int piv = c;
      for (int r = c + 1; r < n; ++r)
        if (fabs(M[r * n + c]) > fabs(M[piv * n + c])) piv = r;
      const double p = M[piv * n + c];
      sm.inv_piv = p == 0 ? -1 : piv;
      sm.inv_d = p == 0 ? 0.0 : 1.0 / p;
    }
    __syncthreads();
    const int piv = sm.inv_piv;
    if (piv < 0) return false;
    const double d = sm.inv_d;
    if (t < 2 * n) {  // swap rows piv and c of [M | Ai], scale the new row c
      double* X = t < n ? M : Ai;
      const int j = t < n ? t : t - n;
      const double vp = X[piv * n + j], vc = X[c * n + j];
      if (piv != c) X[piv * n + j] = vc;
      X[c * n + j] = vp * d;
    }
    __syncthreads();
    for (int i = t; i < n * 2 * n; i += kPoThreads) {
      const int r = i / (2 * n), jj = i % (2 * n);
      if (r == c || jj == c) continue;
      const double f = M[r * n + c];
      if (f == 0) continue;
      double* X = jj < n ? M : Ai;
      const int j = jj < n ? jj : jj - n;
      X[r * n + j] -= f * X[c * n + j];
    }
    __syncthreads();
  }
  return true;
}

// Omega e, chi2 and Huber weights of the inertial / bias / prior edges from the residuals in shared memory.
// All threads call it (contains barriers); the residuals must be visible (barrier before).
__device__ void dense_chi2(const PoCtx& c, PoSmem& sm) {
  const int t = threadIdx.x;
  if (t < 9) {
    double q = 0;
    if (c.has_imu)
      for (int j = 0; j < 9; ++j) q += sm.info_imu[t * 9 + j] * sm.err_imu[j];
    sm.oe_imu[t] = q;
  } else if (t >= 32 && t < 47) {
    double q = 0;
    if (!c.fixed_last)
      for (int j = 0; j < 15; ++j) q += sm.info_prior[(t - 32) * 15 + j] * sm.err_prior[j];
    sm.oe_pri[t - 32] = q;
  } else if (t >= 64 && t < 70) {
    sm.oe_bias[t - 64] = sm.info_bias[t - 64] * sm.err_bias[t - 64];
  }
  __syncthreads();
  if (t == 0) {
    double s = 0;
    for (int i = 0; i < 9; ++i) s += sm.err_imu[i] * sm.oe_imu[i];
    sm.chi2_imu = c.has_imu ? s : 0.0;
    huber_rho(c.delta_imu, sm.chi2_imu, sm.rho_imu, sm.r1_imu);
  } else if (t == 32) {
    double s = 0;
    for (int i = 0; i < 15; ++i) s += sm.err_prior[i] * sm.oe_pri[i];
    sm.chi2_prior = c.fixed_last ? 0.0 : s;
    huber_rho(c.delta_prior, sm.chi2_prior, sm.rho_prior, sm.r1_prior);
  } else if (t == 64) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += sm.err_bias[i] * sm.oe_bias[i];
    sm.chi2_bias = s;
    huber_rho(c.delta_bias, s, sm.rho_bias, sm.r1_bias);
  }
  __syncthreads();
}
__device__ __forceinline__ double dense_rho_sum(const PoCtx& c, const PoSmem& sm) {
  double tot = 0;
  if (c.imu_mode) {
    if (c.has_imu) tot += sm.rho_imu;
    tot += sm.rho_bias;
    if (!c.fixed_last) tot += sm.rho_prior;
  }
  return tot;
}

// computeActiveErrors + buildSystem at the current estimate into sm.lin[set]; chi2 per visual edge stored in c.chi2.
// sm.cp must hold the camera pose of the current estimate.  All threads call it.
__device__ void evaluate(const PoCtx& c, PoSmem& sm, int set) {
  const int n = c.n, dv = c.dv, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PoLin& L = sm.lin[set];
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0;
  PO_T(t_a);
  if (warp < kPoVisWarps) {
    for (int i = threadIdx.x; i < c.E; i += kPoVisWarps * 32) {
      if (sm.eflag[i] & 1) continue;
      const bool stereo = c.flags[i] & VIEO_EDGE_STEREO;
      const int DE = stereo ? 3 : 2;
      const Vec3 X = ld3(c.Xw + 3 * (size_t)i);
      double e[3];
      reproj_error(c.cam, sm.cp, X, c.obs + 3 * (size_t)i, stereo, e);
      const double wi = (double)c.w[i];
      double chi = 0;
      for (int k = 0; k < DE; ++k) chi += e[k] * (wi * e[k]);
      c.chi2[i] = chi;
      double r0, r1;
      huber_rho(edge_delta(c, sm, i), chi, r0, r1);
      acc[27] += r0;
      Mat3 Jp, Jr, JX;
      reproj_jac(c.cam, sm.cp, X, stereo, Jp, Jr, JX);
      const double w = r1 * wi;
      double J[3][6];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          J[k][cc] = Jp.m[3 * k + cc];
          J[k][3 + cc] = Jr.m[3 * k + cc];
        }
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double s = 0;
        for (int k = 0; k < DE; ++k) s += J[k][a] * (-(wi * e[k]) * r1);
        acc[21 + a] += s;
#pragma unroll
        for (int cc = a; cc < 6; ++cc) {
          double h = 0;
          for (int k = 0; k < DE; ++k) h += (J[k][a] * w) * J[k][cc];
          acc[q++] += h;
        }
      }
    }
  } else if (c.imu_mode && lane == 0) {
    if (warp == kPoImuWarp) {
      if (c.has_imu) {
        navstate_error(sm.st[1], sm.st[0], sm.pre, c.gw, false, sm.err_imu);
        navstate_jac_pvr24(sm.st[1], sm.st[0], sm.pre, c.gw, sm.err_imu, sm.Jimu);
      }
    } else {
      bias_error(sm);
      if (!c.fixed_last) {
        prior_error(sm.st[1], c.prior, sm.err_prior);
        prior_jac15(c, sm);
      }
    }
  }
  PO_T(t_b);
  if (threadIdx.x == 0) PO_ACC(0, t_a, t_b);        // visual edges, thread 0
  if (threadIdx.x == 32 * kPoImuWarp) PO_ACC(1, t_a, t_b);      // inertial edge, last warp lane 0
  if (threadIdx.x == 32 * (kPoImuWarp - 1)) PO_ACC(2, t_a, t_b);  // bias + prior, lane 0 of the warp before
  block_sum<28>(sm, acc, kPoVisWarps);  // its barriers also publish the dense residuals / Jacobians
  PO_T(t_c);
  if (threadIdx.x == 0) PO_ACC(3, t_b, t_c);        // wait + reduction
  if (c.imu_mode) {
    dense_chi2(c, sm);
    // J^T (rho' Omega): AtO[a][j] = sum_i J[i][a] (r1 Omega[i][j])
    for (int t = threadIdx.x; t < 216 + 225; t += kPoThreads) {
      if (t < 216) {
        if (!c.has_imu) continue;
        const int a = t / 9, j = t % 9;
        double q = 0;
        for (int i = 0; i < 9; ++i) q += sm.Jimu[i * 24 + a] * (sm.r1_imu * sm.info_imu[i * 9 + j]);
        sm.AtOimu[t] = q;
      } else {
        if (c.fixed_last) continue;
        const int u = t - 216, a = u / 15, j = u % 15;
        double q = 0;
        for (int i = 0; i < 15; ++i) q += sm.Jpri[i * 15 + a] * (sm.r1_prior * sm.info_prior[i * 15 + j]);
        sm.AtOpri[u] = q;
      }
    }
    // oe <- -(Omega e) rho'
    if (threadIdx.x < 9) sm.oe_imu[threadIdx.x] = -sm.oe_imu[threadIdx.x] * sm.r1_imu;
    else if (threadIdx.x >= 32 && threadIdx.x < 47) sm.oe_pri[threadIdx.x - 32] = -sm.oe_pri[threadIdx.x - 32] * sm.r1_prior;
    else if (threadIdx.x >= 64 && threadIdx.x < 70) sm.oe_bias[threadIdx.x - 64] = -sm.oe_bias[threadIdx.x - 64] * sm.r1_bias;
    __syncthreads();
  }
  PO_T(t_d);
  if (threadIdx.x == 0) PO_ACC(4, t_c, t_d);        // Omega e, chi2, AtO
  // every entry of H (b in the extra column) in a fixed order: inertial, bias, prior, visual.
  // Hessian index order: frame PVR/PR 0.., frame bias 9..14, last PVR 15..23, last bias 24..29
  for (int t = threadIdx.x; t < n * (n + 1); t += kPoThreads) {
    const int r = t / (n + 1), cc = t % (n + 1);
    const bool rhs = cc == n;
    double h = 0;
    if (c.imu_mode) {
      // local column of the inertial strip: frame PVR -> 9.., last PVR -> 0.., last bias -> 18.., frame bias: none
      auto imu_col = [](int g) { return g < 9 ? 9 + g : g < 15 ? -1 : g < 24 ? g - 15 : 18 + (g - 24); };
      if (c.has_imu) {
        const int lr = imu_col(r);
        if (lr >= 0) {
          if (rhs) {
            double s = 0;
            for (int i = 0; i < 9; ++i) s += sm.Jimu[i * 24 + lr] * sm.oe_imu[i];
            h += s;
          } else {
            const int lc = imu_col(cc);
            if (lc >= 0) {
              double s = 0;
              for (int j = 0; j < 9; ++j) s += sm.AtOimu[lr * 9 + j] * sm.Jimu[j * 24 + lc];
              h += s;
            }
          }
        }
      }
      {  // bias walk: J_i = -I on the last bias (24..29), J_j = +I on the frame bias (9..14)
        const int kr = (r >= 9 && r < 15) ? r - 9 : (r >= 24 ? r - 24 : -1);
        if (kr >= 0) {
          const double sgn_r = r < 15 ? 1.0 : -1.0;
          if (rhs) h += sgn_r * sm.oe_bias[kr];
          else {
            const int kc = (cc >= 9 && cc < 15) ? cc - 9 : (cc >= 24 ? cc - 24 : -1);
            if (kc == kr) h += (sgn_r * (cc < 15 ? 1.0 : -1.0)) * (sm.r1_bias * sm.info_bias[kr]);
          }
        }
      }
      if (!c.fixed_last && r >= 15) {  // prior: last PVR 15..23 -> 0..8, last bias 24..29 -> 9..14
        const int lr = r - 15;
        if (rhs) {
          double s = 0;
          for (int i = 0; i < 15; ++i) s += sm.Jpri[i * 15 + lr] * sm.oe_pri[i];
          h += s;
        } else if (cc >= 15) {
          const int lc = cc - 15;
          double s = 0;
          for (int j = 0; j < 15; ++j) s += sm.AtOpri[lr * 15 + j] * sm.Jpri[j * 15 + lc];
          h += s;
        }
      }
    }
    // visual block: pose columns (dp 0..2, dphi dv-3..dv-1) of the frame vertex
    const int va = r < 3 ? r : (r >= dv - 3 && r < dv ? r - (dv - 6) : -1);
    if (va >= 0) {
      if (rhs) h += sm.tot[21 + va];
      else {
        const int vc = cc < 3 ? cc : (cc >= dv - 3 && cc < dv ? cc - (dv - 6) : -1);
        if (vc >= 0) {
          const int lo = min(va, vc), hi = max(va, vc);
          h += sm.tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];  // upper-triangle packing of the 6x6 block
        }
      }
    }
    if (rhs) L.b[r] = h;
    else L.H[r * n + cc] = h;
  }
  if (threadIdx.x == 0) L.chi = dense_rho_sum(c, sm) + sm.tot[27];
  __syncthreads();
  PO_T(t_e);
  if (threadIdx.x == 0) {
    PO_ACC(5, t_d, t_e);                            // H / b entries
    PO_ACC(6, t_a, t_e);                            // evaluate total
    PO_ACC(7, 0, 1);                                // evaluate calls
  }
}

// (H + lambda I) x = b by warp 0 (N <= 30 <= 32).  Lane i keeps row i of the lower triangle in REGISTERS (fully
// unrolled over the compile-time dimension, so every index is static): per column one broadcast of the pivot, one
// reciprocal square root, and the rank-1 update of the trailing rows with the column entries exchanged by shuffles —
// no shared-memory round trips and no __syncwarp inside the factorisation (the round-1 form kept A in shared memory:
// 17 k cycles per solve at N = 30).  The arithmetic per entry is the same sequence as before (A[i][k] -= f_i * f_k in
// increasing j), so the factor is bit-identical.  Sets sm.ok (LDLT::isPositive).  Other warps wait at the caller's barrier.
template <int N>
__device__ __forceinline__ void solve_system_n(PoSmem& sm, const PoLin& L) {
  const int i = threadIdx.x;
  double a[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a[k] = (i < N && k <= i) ? L.H[i * N + k] + (k == i ? sm.lambda : 0.0) : 0.0;
  bool good = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double d2 = __shfl_sync(0xffffffffu, a[j], j);
    if (!(d2 > 0) || !isfinite(d2)) {
      good = false;
      break;
    }
    const double rd = rsqrt(d2);
    double f = 0;
    if (i == j) a[j] = d2 * rd;  // sqrt(d2)
    else if (i > j && i < N) {
      f = a[j] * rd;
      a[j] = f;
    }
#pragma unroll
    for (int k = j + 1; k < N; ++k) {
      // unconditional: on lanes i < k this only touches a[k] above the diagonal, which nothing reads (a predicate here
      // costs ptxas 3 KB of spill code under the 128-register cap)
      a[k] -= f * __shfl_sync(0xffffffffu, f, k);
    }
  }
  if (i == 0) sm.ok = good ? 1 : 0;
  if (!good) return;
  // forward: L y = b with the row in registers; the pivot entry is broadcast every step
  double dii = 1.0;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k == i) dii = a[k];
  const double rd_i = i < N ? 1.0 / dii : 0.0;
  double yi = i < N ? L.b[i] : 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double yj = __shfl_sync(0xffffffffu, yi * rd_i, j);
    if (i == j) yi = yj;
    else if (i > j && i < N) yi -= a[j] * yj;
  }
  // backward: L^T x = y needs column i of L on lane i: the factor goes through shared memory once (sm.S, row-major)
  double* A = sm.S;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (i < N && k <= i) A[i * N + k] = a[k];
  __syncwarp();
  double xi = yi;
  for (int j = N - 1; j >= 0; --j) {
    const double xj = __shfl_sync(0xffffffffu, xi * rd_i, j);
    if (i == j) xi = xj;
    else if (i < j) xi -= A[j * N + i] * xj;
  }
  if (i < N) {
    sm.y[i] = yi;
    sm.x[i] = xi;
  }
}
__device__ void solve_system(PoSmem& sm, const PoLin& L, int n) {
  if (threadIdx.x >= 32) return;
  if (n == 30) solve_system_n<30>(sm, L);
  else if (n == 15) solve_system_n<15>(sm, L);
  else solve_system_n<6>(sm, L);
}

__device__ void apply_update(const PoCtx& c, PoSmem& sm) {
  if (!c.imu_mode) {
    ns_inc_pr(sm.st[0], sm.x);
  } else {
    ns_inc_pvr(sm.st[0], sm.x);
    ns_inc_bias(sm.st[0], sm.x + 9);
    if (!c.fixed_last) {
      ns_inc_pvr(sm.st[1], sm.x + 15);
      ns_inc_bias(sm.st[1], sm.x + 24);
    }
  }
  sm.cp = cam_pose(c.cam, sm.st[0]);
}

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg::solve inlined
__device__ void optimize(const PoCtx& c, PoSmem& sm, int iterations) {
  const int n = c.n;
  if (threadIdx.x < n) sm.x[threadIdx.x] = 0;
  if (threadIdx.x == 0) sm.cp = cam_pose(c.cam, sm.st[0]);
  __syncthreads();
  int cur = 0;
  evaluate(c, sm, cur);  // levels / kernels / estimate may have changed since the last call
  bool ok = true;
  for (int it = 0; it < iterations && ok; ++it) {
    double currentChi = sm.lin[cur].chi;
    const double iniChi = currentChi;
    if (threadIdx.x == 0) {
      if (it == 0) {  // computeLambdaInit: tau * max |diag(H)|
        double mx = 0;
        for (int i = 0; i < n; ++i) mx = fmax(fabs(sm.lin[cur].H[i * n + i]), mx);
        sm.lambda = 1e-5 * mx;
        sm.ni = 2;
        sm.nBad = 0;
      }
      sm.total_iters++;
    }
    double rho = 0;
    int qmax = 0;
    do {
      __syncthreads();
      PO_T(t_s0);
      solve_system(sm, sm.lin[cur], n);
      PO_T(t_s1);
      if (threadIdx.x == 0) {
        sm.bak[0] = sm.st[0];
        sm.bak[1] = sm.st[1];
        if (sm.ok) apply_update(c, sm);
      }
      PO_T(t_s2);
      if (threadIdx.x == 0) {
        PO_ACC(8, t_s0, t_s1);                      // Cholesky + solves
        PO_ACC(9, t_s1, t_s2);                      // oplus + camera pose
      }
      __syncthreads();
      evaluate(c, sm, 1 - cur);  // errors at the trial estimate + speculative linearisation
      double tempChi = sm.lin[1 - cur].chi;
      if (!sm.ok) tempChi = 1.7976931348623157e308;
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < n; ++j) scale += sm.x[j] * (sm.lambda * sm.x[j] + sm.lin[cur].b[j]);
      scale += 1e-3;
      rho /= scale;
      __syncthreads();  // every thread has read lambda / x before thread 0 changes them
      if (rho > 0 && isfinite(tempChi)) {
        if (threadIdx.x == 0) {
          double alpha = 1. - pow((2 * rho - 1), 3);
          alpha = fmin(alpha, 2. / 3.);
          sm.lambda *= fmax(1. / 3., alpha);
          sm.ni = 2;
        }
        currentChi = tempChi;
        cur = 1 - cur;
      } else {
        if (threadIdx.x == 0) {
          sm.lambda *= sm.ni;
          sm.ni *= 2;
          sm.st[0] = sm.bak[0];
          sm.st[1] = sm.bak[1];
          sm.cp = cam_pose(c.cam, sm.st[0]);
        }
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    __syncthreads();
    if (qmax == 10 || rho == 0) {
      ok = false;
    } else {
      // Raul's stop criterion: 3 consecutive iterations with < 0.1 % gain
      int nb = sm.nBad;
      nb = ((iniChi - currentChi) * 1e3 < iniChi) ? nb + 1 : 0;
      __syncthreads();
      if (threadIdx.x == 0) sm.nBad = nb;
      if (nb >= 3) ok = false;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPoThreads, VIEO_PO_MIN_CTAS) k_pose_opt(const VieoPoseOptProblem* __restrict__ pbs,
                                                         const VieoCamera* __restrict__ camp,
                                                         const double* __restrict__ Xw, const float* __restrict__ obs,
                                                         const float* __restrict__ inv_sigma2,
                                                         const uint8_t* __restrict__ flags,
                                                         VieoPoseOptResult* __restrict__ res,
                                                         uint8_t* __restrict__ outlier, double* __restrict__ chi2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PoSmem& sm = *reinterpret_cast<PoSmem*>(smem_raw);
  const VieoPoseOptProblem& pb = pbs[blockIdx.x];
  VieoPoseOptResult& R = res[blockIdx.x];
  PoCtx c;
  c.pb = &pb;
  c.cam = cam_load(*camp);
  const int e0 = pb.edge_begin;
  c.E = min(pb.edge_end - pb.edge_begin, kPoMaxEdges);
  c.Xw = Xw + 3 * (size_t)e0;
  c.obs = obs + 3 * (size_t)e0;
  c.w = inv_sigma2 + e0;
  c.flags = flags + e0;
  c.chi2 = chi2 + e0;
  uint8_t* outl = outlier + e0;
  c.imu_mode = pb.mode == 1;
  c.fixed_last = !pb.last_has_prior;
  c.has_imu = c.imu_mode && pb.preint.dt != 0;
  c.dv = c.imu_mode ? 9 : 6;
  c.n = c.imu_mode ? (c.fixed_last ? 15 : 30) : 6;
  c.delta_mono = (double)(float)sqrt(5.991);
  c.delta_stereo = (double)(float)sqrt(7.815);
  c.delta_imu = c.fixed_last ? sqrt(16.919) : 0.0;   // USE_ZZH_IMU_EDGE_FEBA (include/Optimizer.h:293-303)
  c.delta_bias = c.fixed_last ? sqrt(12.592) : 0.0;  // :326-333
  c.delta_prior = sqrt(25.0);                        // :351
  c.gw = ld3(pb.gw);
  c.prior = ns_load(pb.prior);
  const int E = c.E;
  PO_T(t_k0);

  for (int i = threadIdx.x; i < E; i += kPoThreads) {
    sm.eflag[i] = 0;
    outl[i] = 0;
    c.chi2[i] = 0;
  }
  if (threadIdx.x < 9) {
    const int t = threadIdx.x;
    sm.pre.Rij[t] = pb.preint.Rij[t]; sm.pre.Jgp[t] = pb.preint.Jgp[t]; sm.pre.Jap[t] = pb.preint.Jap[t];
    sm.pre.Jgv[t] = pb.preint.Jgv[t]; sm.pre.Jav[t] = pb.preint.Jav[t]; sm.pre.JgR[t] = pb.preint.JgR[t];
    if (t < 3) {
      sm.pre.vij[t] = pb.preint.vij[t];
      sm.pre.pij[t] = pb.preint.pij[t];
    }
    if (t == 0) sm.pre.dt = pb.preint.dt;
  }
  if (threadIdx.x == 0) {
    sm.st[0] = sm.ini[0] = ns_load(pb.cur);
    sm.st[1] = sm.ini[1] = ns_load(pb.last);
    sm.total_iters = 0;
    sm.lambda = 0;
    sm.ok = 1;
    for (int i = 0; i < 9; ++i) sm.err_imu[i] = 0;
    for (int i = 0; i < 15; ++i) sm.err_prior[i] = 0;
    for (int i = 0; i < 6; ++i) sm.err_bias[i] = 0;
    sm.chi2_imu = sm.chi2_bias = sm.chi2_prior = 0;
    sm.rho_imu = sm.rho_bias = sm.rho_prior = 0;
    sm.r1_imu = sm.r1_bias = sm.r1_prior = 1;
    if (c.imu_mode) {
      const double dtij = pb.preint.dt != 0 ? pb.preint.dt : pb.dt_frames;
      for (int k = 0; k < 6; ++k) {
        const double w = (k < 3 ? pb.inv_sigma_bg2 : pb.inv_sigma_ba2) / dtij;
        sm.info_bias[k] = c.fixed_last ? w * 1e-2 : w;
      }
    }
  }
  if (c.imu_mode) {
    for (int i = threadIdx.x; i < 216; i += kPoThreads) sm.Jimu[i] = 0;
    for (int i = threadIdx.x; i < 225; i += kPoThreads) sm.Jpri[i] = 0;
    if (!c.fixed_last)
      for (int i = threadIdx.x; i < 225; i += kPoThreads) sm.info_prior[i] = pb.prior_info[i];
    if (c.has_imu) {  // GetProcessedInfoij = mSigmaij.inverse() (OdomPreIntegrator.h:129-138), x 1e-2 when last is fixed
      for (int i = threadIdx.x; i < 81; i += kPoThreads) sm.S[i] = pb.preint.SigmaPVR[i];
      __syncthreads();
      const bool inv_ok = block_inverse(sm, sm.S, 9, sm.info_imu);
      for (int i = threadIdx.x; i < 81; i += kPoThreads) {
        double v = inv_ok ? sm.info_imu[i] : nan("");
        if (c.fixed_last) v *= 1e-2;
        sm.info_imu[i] = v;
      }
    }
  }
  __syncthreads();
  PO_T(t_k1);
  if (threadIdx.x == 0) PO_ACC(10, t_k0, t_k1);
  const int nInitial = pb.edge_end - pb.edge_begin;
  if (nInitial < 3 && !(c.imu_mode && pb.no_mps)) {  // src/Optimizer.cc:1787, include/Optimizer.h:499-503
    if (threadIdx.x == 0) {
      R.cur = pb.cur;
      R.last = pb.last;
      for (int i = 0; i < 225; ++i) R.marg_cov_inv[i] = 0;
      R.chi2_final = 0; R.lambda_final = 0; R.n_inliers = 0; R.n_initial = nInitial; R.iterations = 0; R.prior_set = 0;
    }
    return;
  }
  const float chi2Mono = 5.991f, chi2Stereo = 7.815f;
  const float chi2close = 1.5 * chi2Mono;
  const int n_dense = c.imu_mode ? (c.has_imu ? 1 : 0) + 1 + (c.fixed_last ? 0 : 1) : 0;
  const int n_edges_total = nInitial + n_dense;
  int nBad = 0;
  for (int it = 0; it < 4; ++it) {
    if (!c.imu_mode || !c.has_imu) {  // reset the estimate (src/Optimizer.cc:1804, include/Optimizer.h:538-545)
      if (threadIdx.x == 0) {
        sm.st[0] = sm.ini[0];
        if (c.imu_mode && !c.fixed_last) sm.st[1] = sm.ini[1];
      }
    }
    __syncthreads();
    optimize(c, sm, 10);
    PO_T(t_r0);
    double bad[1] = {0};
    for (int i = threadIdx.x; i < E; i += kPoThreads) {
      double e[3], depth = 1;
      if (c.imu_mode || outl[i]) c.chi2[i] = vis_chi2(c, sm, i, e, &depth);
      const float chi = (float)c.chi2[i];
      bool isbad;
      if (c.flags[i] & VIEO_EDGE_STEREO) isbad = chi > chi2Stereo;
      else if (c.imu_mode) isbad = chi > ((c.flags[i] & VIEO_EDGE_CLOSE) ? chi2close : chi2Mono) || !(depth > 0.);
      else isbad = chi > chi2Mono;
      outl[i] = isbad;
      uint8_t f = (sm.eflag[i] & ~1) | (isbad ? 1 : 0);
      if (it == 2) f |= 2;
      sm.eflag[i] = f;
      bad[0] += isbad;
    }
    block_sum<1>(sm, bad, kPoWarps);
    nBad = (int)sm.tot[0];
    PO_T(t_r1);
    if (threadIdx.x == 0) PO_ACC(11, t_r0, t_r1);
    if (n_edges_total < 10) break;
  }
  if (c.imu_mode && nInitial - nBad < 30) {  // rescue pass (include/Optimizer.h:619-648)
    double bad[1] = {0};
    for (int i = threadIdx.x; i < E; i += kPoThreads) {
      double e[3];
      const double chi = vis_chi2(c, sm, i, e);
      c.chi2[i] = chi;
      if (chi < ((c.flags[i] & VIEO_EDGE_STEREO) ? (double)24.f : (double)18.f)) {
        sm.eflag[i] &= ~1;
        outl[i] = 0;
      } else
        bad[0] += 1;
    }
    block_sum<1>(sm, bad, kPoWarps);
    nBad = (int)sm.tot[0];
  }
  // activeRobustChi2 of the final level-0 set with the stored errors
  {
    double acc[1] = {0};
    for (int i = threadIdx.x; i < E; i += kPoThreads) {
      if (sm.eflag[i] & 1) continue;
      double r0, r1;
      huber_rho(edge_delta(c, sm, i), c.chi2[i], r0, r1);
      acc[0] += r0;
    }
    block_sum<1>(sm, acc, kPoWarps);
  }
  if (threadIdx.x == 0) {
    ns_store(sm.st[0], R.cur);
    ns_store(c.imu_mode ? sm.st[1] : sm.ini[1], R.last);
    R.chi2_final = sm.tot[0] + dense_rho_sum(c, sm);
    R.lambda_final = sm.lambda;
    R.n_inliers = nInitial - nBad;
    R.n_initial = nInitial;
    R.iterations = sm.total_iters;
    R.prior_set = 0;
  }
  if (!(c.imu_mode && pb.compute_marg)) {
    for (int i = threadIdx.x; i < 225; i += kPoThreads) R.marg_cov_inv[i] = 0;
    return;
  }
  // ---- marginal prior, exact_mode = kExactRobust (include/Optimizer.h:126-206, 671-728) ----------------------
  // errors of the inertial / bias / prior edges recomputed and every edge re-linearised at the final estimate
  PO_T(t_m0);
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0 && warp == kPoImuWarp && c.has_imu) {
      navstate_error(sm.st[1], sm.st[0], sm.pre, c.gw, false, sm.err_imu);
      navstate_jac_pvr24(sm.st[1], sm.st[0], sm.pre, c.gw, sm.err_imu, sm.Jimu);
    }
    if (lane == 0 && warp == kPoImuWarp - 1) {
      bias_error(sm);
      if (!c.fixed_last) {
        prior_error(sm.st[1], c.prior, sm.err_prior);
        prior_jac15(c, sm);
      }
    }
    if (threadIdx.x == 0) sm.cp = cam_pose(c.cam, sm.st[0]);
  }
  __syncthreads();
  dense_chi2(c, sm);
  const double wI = sm.r1_imu, wB = sm.r1_bias, wP = sm.r1_prior;
  {
    double acc[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) acc[k] = 0;
    for (int i = threadIdx.x; i < E; i += kPoThreads) {
      if (sm.eflag[i] & 1) continue;
      const bool stereo = c.flags[i] & VIEO_EDGE_STEREO;
      const int DE = stereo ? 3 : 2;
      Mat3 Jp, Jr, JX;
      reproj_jac(c.cam, sm.cp, ld3(c.Xw + 3 * (size_t)i), stereo, Jp, Jr, JX);
      double r0, r1;
      huber_rho(edge_delta(c, sm, i), c.chi2[i], r0, r1);
      const double w = r1 * (double)c.w[i];
      double J[3][6];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          J[k][cc] = Jp.m[3 * k + cc];
          J[k][3 + cc] = Jr.m[3 * k + cc];
        }
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int cc = a; cc < 6; ++cc) {
          double s = 0;
          for (int k = 0; k < DE; ++k) s += J[k][a] * (w * J[k][cc]);
          acc[q++] += s;
        }
    }
    block_sum<21>(sm, acc, kPoWarps);
  }
  // Every product below is formed by the whole block, one output entry per thread, with the summation order of the
  // single-thread jtoj() it replaces (W = (w Omega) J first, then J^T W), so the entries are bit-identical:
  //   G  (24 x 24) = Jimu^T (wI Omega_imu) Jimu     over the strip [Ji | Jj | Jb]
  //   Gp (15 x 15) = Jpri^T (wP Omega_prior) Jpri
  double* Wi = sm.AtOimu;        // 9 x 24
  double* Wp = sm.AtOpri;        // 15 x 15
  double* G = sm.lin[1].H;       // 576 of 900
  double* Gp = sm.lin[1].H + 576;  // 225
  for (int t = threadIdx.x; t < 216 + 225; t += kPoThreads) {
    if (t < 216) {
      const int i = t / 24, cc = t % 24;
      double q = 0;
      if (c.has_imu)
        for (int j = 0; j < 9; ++j) q += (wI * sm.info_imu[i * 9 + j]) * sm.Jimu[j * 24 + cc];
      Wi[t] = q;
    } else if (!c.fixed_last) {
      const int u = t - 216, i = u / 15, cc = u % 15;
      double q = 0;
      for (int j = 0; j < 15; ++j) q += (wP * sm.info_prior[i * 15 + j]) * sm.Jpri[j * 15 + cc];
      Wp[u] = q;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 576 + 225; t += kPoThreads) {
    if (t < 576) {
      const int a = t / 24, cc = t % 24;
      double q = 0;
      if (c.has_imu)
        for (int i = 0; i < 9; ++i) q += sm.Jimu[i * 24 + a] * Wi[i * 24 + cc];
      G[t] = q;
    } else if (!c.fixed_last) {
      const int u = t - 576, a = u / 15, cc = u % 15;
      double q = 0;
      for (int i = 0; i < 15; ++i) q += sm.Jpri[i * 15 + a] * Wp[i * 15 + cc];
      Gp[u] = q;
    }
  }
  __syncthreads();
  // C: frame block; CL: last-frame block (inertial + bias walk + prior); CCL: coupling.  Strip columns: Ji 0.., Jj 9.., Jb 18..
  if (threadIdx.x < 225) {
    const int a = threadIdx.x / 15, cc = threadIdx.x % 15;
    const double wb = (a >= 9 && a == cc) ? wB * sm.info_bias[a - 9] : 0.0;
    double v = (a < 9 && cc < 9 && c.has_imu) ? G[(9 + a) * 24 + 9 + cc] : 0.0;
    if (a >= 9 && a == cc) v = wb;
    // visual 6 x 6 (dp, dphi) -> state rows {0, 1, 2, 6, 7, 8}, upper-triangle packing of block_sum<21>
    const int va = a < 3 ? a : (a >= 6 && a < 9 ? a - 3 : -1), vc = cc < 3 ? cc : (cc >= 6 && cc < 9 ? cc - 3 : -1);
    if (va >= 0 && vc >= 0) {
      const int lo = min(va, vc), hi = max(va, vc);
      v += sm.tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
    }
    sm.C[threadIdx.x] = v;
    if (!c.fixed_last) {
      // upper-right block first, the lower-left one is its mirror (as the reference fills it)
      const int ua = (a >= 9 && cc < 9) ? cc : a, uc = (a >= 9 && cc < 9) ? a : cc;
      double l = 0;
      if (c.has_imu) l = G[(ua < 9 ? ua : 18 + ua - 9) * 24 + (uc < 9 ? uc : 18 + uc - 9)];
      if (ua >= 9 && ua == uc) l += wB * sm.info_bias[ua - 9];
      l += Gp[ua * 15 + uc];
      sm.CL[threadIdx.x] = l;
      double k = 0;
      if (a < 9) k = c.has_imu ? G[(9 + a) * 24 + (cc < 9 ? cc : 18 + cc - 9)] : 0.0;
      else if (a == cc) k = -wb;
      sm.CCL[threadIdx.x] = k;
    }
  }
  __syncthreads();
  if (!c.fixed_last) {
    // cov_inv -= E C^-1 E^T (JacobiSVD inverse without clamping in the reference, :709-728)
    double* Cinv = sm.lin[0].H;  // 225 <= 900
    double* T = sm.S;
    const bool inv_ok = block_inverse(sm, sm.CL, 15, Cinv);
    if (threadIdx.x < 225) {
      const int a = threadIdx.x / 15, cc = threadIdx.x % 15;
      double q = 0;
      for (int k = 0; k < 15; ++k) q += sm.CCL[a * 15 + k] * (inv_ok ? Cinv[k * 15 + cc] : nan(""));
      T[threadIdx.x] = q;
    }
    __syncthreads();
    if (threadIdx.x < 225) {
      const int a = threadIdx.x / 15, cc = threadIdx.x % 15;
      double q = 0;
      for (int k = 0; k < 15; ++k) q += T[a * 15 + k] * sm.CCL[cc * 15 + k];
      sm.C[threadIdx.x] -= q;
    }
  }
  if (threadIdx.x < 225) R.marg_cov_inv[threadIdx.x] = sm.C[threadIdx.x];
  if (threadIdx.x == 0) R.prior_set = 1;
  PO_T(t_m1);
  if (threadIdx.x == 0) {
    PO_ACC(12, t_m0, t_m1);
    PO_ACC(13, t_k0, t_m1);
  }
}

}  // namespace vieo

using namespace vieo;

extern "C" {

#ifdef VIEO_PROF
int vieo_debug_po_prof(long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, g_po_prof, sizeof(long long) * 16);
  if (reset) {
    long long z[16] = {};
    cudaMemcpyToSymbol(g_po_prof, z, sizeof(z));
  }
  return 0;
}
#endif

int vieo_pose_opt_batch_dev(const VieoPoseOptProblem* pbs_dev, int n, const VieoCamera* cam_dev, const double* Xw_dev,
                            const float* obs_dev, const float* inv_sigma2_dev, const uint8_t* flags_dev,
                            VieoPoseOptResult* res_dev, uint8_t* outlier_dev, double* chi2_dev, void* stream) {
  VIEO_ARG(n >= 0, "bad argument");
  if (n == 0) return VIEO_OK;
  VIEO_ARG(pbs_dev && cam_dev && res_dev && outlier_dev && chi2_dev, "null argument");
  static SmemOptIn opt_in;
  VIEO_CK(smem_opt_in(k_pose_opt, sizeof(PoSmem), opt_in));
  k_pose_opt<<<n, kPoThreads, sizeof(PoSmem), (cudaStream_t)stream>>>(pbs_dev, cam_dev, Xw_dev, obs_dev, inv_sigma2_dev,
                                                                      flags_dev, res_dev, outlier_dev, chi2_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_pose_opt_batch(const VieoPoseOptProblem* pbs, int n, const VieoCamera* cam, const double* Xw, const float* obs,
                        const float* inv_sigma2, const uint8_t* flags, int n_edges, VieoPoseOptResult* res,
                        uint8_t* outlier, double* chi2, int device) {
  VIEO_ARG(n >= 0 && n_edges >= 0, "bad argument");
  if (n == 0) return VIEO_OK;
  VIEO_ARG(pbs && cam && res && (n_edges == 0 || (Xw && obs && inv_sigma2 && flags && outlier && chi2)), "null argument");
  VIEO_ARG(cam->model >= 0 && cam->model <= 2 && cam->num_k >= 0 && cam->num_k <= 6, "unsupported camera model");
  for (int k = 0; k < n; ++k) {
    VIEO_ARG(pbs[k].edge_begin >= 0 && pbs[k].edge_end >= pbs[k].edge_begin && pbs[k].edge_end <= n_edges, "edge range");
    if (pbs[k].edge_end - pbs[k].edge_begin > kPoMaxEdges) {
      set_error("vieo_pose_opt_batch: frame %d has more than %d correspondences", k, kPoMaxEdges);
      return VIEO_E_CAPACITY;
    }
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  const size_t ne = std::max(n_edges, 1);
  void* d_pb = cs->get(0, sizeof(VieoPoseOptProblem) * n);
  void* d_cam = cs->get(1, sizeof(VieoCamera));
  void* d_X = cs->get(2, 24 * ne);
  void* d_o = cs->get(3, 12 * ne);
  void* d_w = cs->get(4, 4 * ne);
  void* d_f = cs->get(5, ne);
  void* d_r = cs->get(6, sizeof(VieoPoseOptResult) * n);
  void* d_out = cs->get(7, ne);
  void* d_chi = cs->get(8, 8 * ne);
  if (!d_pb || !d_cam || !d_X || !d_o || !d_w || !d_f || !d_r || !d_out || !d_chi) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  VIEO_CK(cudaMemcpyAsync(d_pb, pbs, sizeof(VieoPoseOptProblem) * n, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(d_cam, cam, sizeof(VieoCamera), cudaMemcpyHostToDevice, st));
  if (n_edges) {
    VIEO_CK(cudaMemcpyAsync(d_X, Xw, 24 * ne, cudaMemcpyHostToDevice, st));
    VIEO_CK(cudaMemcpyAsync(d_o, obs, 12 * ne, cudaMemcpyHostToDevice, st));
    VIEO_CK(cudaMemcpyAsync(d_w, inv_sigma2, 4 * ne, cudaMemcpyHostToDevice, st));
    VIEO_CK(cudaMemcpyAsync(d_f, flags, ne, cudaMemcpyHostToDevice, st));
  }
  // edges not covered by any problem's range read back as 0 (the scratch buffers are reused across calls)
  VIEO_CK(cudaMemsetAsync(d_out, 0, ne, st));
  VIEO_CK(cudaMemsetAsync(d_chi, 0, 8 * ne, st));
  rc = vieo_pose_opt_batch_dev((const VieoPoseOptProblem*)d_pb, n, (const VieoCamera*)d_cam, (const double*)d_X,
                               (const float*)d_o, (const float*)d_w, (const uint8_t*)d_f, (VieoPoseOptResult*)d_r,
                               (uint8_t*)d_out, (double*)d_chi, st);
  if (rc) return rc;
  VIEO_CK(cudaMemcpyAsync(res, d_r, sizeof(VieoPoseOptResult) * n, cudaMemcpyDeviceToHost, st));
  if (n_edges) {
    VIEO_CK(cudaMemcpyAsync(outlier, d_out, ne, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(chi2, d_chi, 8 * ne, cudaMemcpyDeviceToHost, st));
  }
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

}  // extern "C"

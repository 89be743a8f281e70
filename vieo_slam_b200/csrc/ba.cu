// Local bundle adjustment on the device — the g2o graph behind Optimizer::LocalBundleAdjustmentNavStatePRV
// (src/Optimizer.cc:21-769; PR(6) + V(3) + Bias(6) vertices per keyframe, marginalised map points) and the visual
// Optimizer::LocalBundleAdjustment (:1876-2307; PR vertices only).
//
// Data layout in HBM (all fp64 unless noted; E edges sorted by point, P points, K keyframes, np pose dims <= 15 K):
//   states K x 176 B, camera poses K x 192 B (Rcw, Rwb, tcw, pwb: recomputed once per error evaluation),
//   points P x 24 B, edges SoA {state i32, point i32, obs 3 x f32, invSigma2 f32, flags u8, level u8, chi2 f64},
//   per edge W = Hpl block 6x3 (144 B) and A = pose-side contribution (21 upper-triangle + 6 rhs, 216 B),
//   per point Hll (72 B), bl (24 B), Dinv (72 B), db (24 B);  sys = [S np^2 | bschur np | b np | chi2] contiguous so
//   that the sharded form all-reduces it with ONE collective per LM trial.
// Kernels per LM trial (block_solver.hpp:353-486, 501-560; optimization_algorithm_levenberg.cpp:61-149):
//   k_ba_linearize   one warp per map point: each lane evaluates one reprojection edge (residual, Jacobians, Huber
//                    weight), the 3x3 Hll / bl are reduced with warp shuffles, W and A are stored per edge
//   k_ba_pose_reduce one block per free keyframe: fixed-order sum of its edges' A blocks -> Hpp diagonal block, b
//   k_ba_dense_build inertial + bias-walk edges (<= 2 per keyframe pair) into Hpp
//   k_ba_schur       one block per free keyframe row: S = Hpp + lambda I - sum_l Hpl Dinv Hpl^T, fixed order
//   k_ba_chol        dense Cholesky + triangular solves of the reduced camera system (np <= 390)
//   k_ba_backsub     one warp per point: xl = Dinv (bl - Hpl^T xp), point update, landmark part of the gain ratio
// All reductions run in a fixed order (no atomics): results are bit-reproducible run to run.
// The LM accept/reject logic stays on the host (one 64-byte read-back per trial) because it must poll the caller's
// abort flag (pbStopFlag, sparse_optimizer.cpp:376) anyway.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <chrono>
#include <thread>
#include <vector>

#include "ba_edges.cuh"

namespace vieo {

constexpr int kBaWarps = 8;
constexpr int kLinPts = 4, kLinLanes = 32 / kLinPts;  // k_ba_linearize: map points per warp / lanes per point

constexpr int kDJ = 26;  // columns of an inertial edge's Jacobian strip: 24 keyframe columns + 2 gravity-direction ones
struct BaDense {  // one inertial (EdgeNavStatePRV[G]) or bias random-walk (EdgeNavStateBias) factor
  int type;       // 0 IMU, 1 bias walk, 3 prior-bias edge of the IMU initialiser's global BA (src/Optimizer.cc:1026-1054):
                  //   EdgeNavStateBias from a FIXED copy of keyframe si's bias (info[6..12) = its bg + dbg, ba + dba)
  int si, sj, pre;
  int color, pad_;  // big path: edges of one colour share no keyframe and are accumulated concurrently
  double delta;     // Huber delta, 0 = none
  double info[81];  // IMU: 9x9 information; bias: [0..6) diagonal
};
struct BaDenseWork {
  double J[9 * kDJ];    // IMU: [Ji(PR) | Jj(PR) | Ji(V) | Jj(V) | Jb | JG], 9 x 26 row-major (JG: EdgeNavStatePRVG only)
  double AtO[kDJ * 9];  // J^T (rho' Omega), 26 x 9
  double oe[9], err[9];
  double chi2, r1, rho0;
  double gg[6];  // gravity-direction block of this edge: JG^T rho' Omega JG (2 x 2) | rhs (2); summed over edges in order
};

// ---------------------------------------------------------------------------------------------------------------
// Device-resident optimiser state.  Every kernel of the LM trial reads its sizes and the current linearisation set from
// here, so the trial sequence has constant launch parameters and is replayed as ONE CUDA graph launch per trial; the
// Levenberg-Marquardt bookkeeping (gain ratio, lambda schedule, accept / restore, stop criteria) runs in k_ba_control.
struct BaParams {
  int K, P, E, np, nfree, n_den, n_pblk, has_dup;
  int n_colors;         // big path: colours of the inertial edges (see BaDense::color)
  int rank, world;      // landmark sharding (1 = single GPU)
  int big;              // global-BA sized handle: H is zeroed by a memset node, multi-CTA Schur / Cholesky kernels
  int lambda_on_poses;  // sharded runs add lambda to the pose diagonal on rank 0 only
  int cur;              // linearisation set (0/1) that belongs to the current estimate
  int done;             // optimize() finished: trial kernels return immediately
  int stop;             // host abort flag seen (pbStopFlag)
  int iteration, iters_target, iters_done, qmax, nBad, ok;
  int trials;           // LM trials run by the current optimize()
  int iters_hist[2];    // iterations of the earlier optimize() stages of an asynchronous LocalBA call (k_ba_begin)
  // the LocalBA routine as ONE graph (vieo_local_ba_prv_begin): what the by-value launch arguments of the direct path carry
  int visual_only;      // Optimizer::LocalBundleAdjustment (no inertial / bias edges, no Chi2LargeSetLevel pass)
  int plan_it[2];       // iterations of the two optimize() stages
  double plan_lambda0;  // their initial lambda (0: g2o's own)
  double lambda, ni, user_lambda;
  double chi_cur, ini_chi;   // activeRobustChi2 of the current estimate / at the start of the iteration
  double pair[4];            // [robust chi2 of the last linearisation, landmark part of computeScale, abort flag seen by
                             //  this rank, -] — all-reduced when sharded, so every rank sees the same sums
  double pscale, chi_dense;  // pose part of computeScale; dense-edge chi2 of the last linearisation
  CamK cam;                  // camera, gravity and Huber deltas of the problem
  Vec3 gw;
  double dm, ds;
  // VertexScale (g2otypes.h:294-311; GlobalBundleAdjustmentNavStatePRV with bScaleOpt, src/Optimizer.cc:843-851) and
  // VertexGThetaXYRwI (g2otypes.h:674-698; the IMU initialiser's call, :852-865): dense border rows / columns of the
  // reduced camera system AFTER every keyframe vertex (ids maxKFid + 1 / + 2): np = keyframe dims + 1 + 2.
  int has_scale, off_s, has_g, off_g;
  // Velocity / bias elimination of the global BA (see k_vb_factor): the dense factorisation runs on the cn = 6 nfree +
  // border dimensional system of the PR vertices; Ky = keyframes whose V / Bias vertices (9 dims) form the eliminated chain
  int vb_elim, cn, Ky;
  double sc, sc_bak;          // VertexScale estimate (1 without the vertex: X * 1.0 is exact) and its push() copy
  double qwI[4], qwI_bak[4];  // RwI (w, x, y, z)
  double GI[3];               // (0, 0, |gw|)
};
// gravity in the world frame as the inertial edges see it: gw, or RwI * GI with the gravity-direction vertex
__device__ __forceinline__ Vec3 ba_gravity(const BaParams& prm) {
  if (!prm.has_g) return prm.gw;
  return m3_mulv(q_matrix({prm.qwI[0], prm.qwI[1], prm.qwI[2], prm.qwI[3]}), {prm.GI[0], prm.GI[1], prm.GI[2]});
}

__global__ void k_ba_campose(CamK cam, const VieoNavState* __restrict__ st, int K, CamPose* __restrict__ cp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) cp[k] = cam_pose(cam, ns_load(st[k]));
}

// graph form: sizes and camera from the device-resident problem header (constant launch parameters, capacity grid)
__global__ void k_ba_campose_p(const BaParams* __restrict__ prm, const VieoNavState* __restrict__ st, CamPose* __restrict__ cp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < prm->K) cp[k] = cam_pose(prm->cam, ns_load(st[k]));
}

// world position of a map point as the visual edges see it: X, or sc * X with the scale vertex (EdgeReprojectPRS:
// "unscaled Xw but scaled pwb", g2otypes.h:400-406 with MODE 1); sc == 1.0 without the vertex, and x * 1.0 == x exactly
__device__ __forceinline__ Vec3 ba_world_point(const double* __restrict__ X, size_t p, double sc) {
  return {X[3 * p] * sc, X[3 * p + 1] * sc, X[3 * p + 2] * sc};
}

__device__ __forceinline__ double edge_huber_delta(uint8_t flags, uint8_t lvl, double dm, double ds) {
  if (lvl & 2) return 0.0;
  return (flags & VIEO_EDGE_STEREO) ? ds : dm;
}

// computeActiveErrors over the visual edges (all == 1: every edge regardless of level, for Chi2LargeSetLevel) and
// per-block partial sums of the robust chi2 of the active ones.
__global__ void __launch_bounds__(256) k_ba_errors(CamK cam, const CamPose* __restrict__ cp, const double* __restrict__ X,
                                                   const int* __restrict__ es, const int* __restrict__ ep,
                                                   const float* __restrict__ obs, const float* __restrict__ w,
                                                   const uint8_t* __restrict__ flags, const uint8_t* __restrict__ lvl,
                                                   const uint8_t* __restrict__ sfix, int points_free, int E, int all,
                                                   double dm, double ds, double* __restrict__ chi2,
                                                   double* __restrict__ partial, const BaParams* __restrict__ prmq,
                                                   int skip_if_visual = 0) {
  __shared__ double s_w[8];
  if (E < 0) {  // graph form (capacity grid): sizes, camera and Huber deltas from the device-resident problem header
    if (skip_if_visual && prmq->visual_only) return;
    E = prmq->E;
    cam = prmq->cam;
    dm = prmq->dm;
    ds = prmq->ds;
    if ((int)blockIdx.x * 256 >= E) return;  // k_ba_dense_errors sums the first ceil(E / 256) partials only
  }
  const double sc = prmq->sc;
  const int i = blockIdx.x * 256 + threadIdx.x;
  double r0 = 0;
  if (i < E) {
    const bool active = !(lvl[i] & 1) && (points_free || !sfix[es[i]]);
    if (all == 2) {  // activeRobustChi2 over the stored errors
      if (active) {
        double r1;
        huber_rho(edge_huber_delta(flags[i], lvl[i], dm, ds), chi2[i], r0, r1);
      }
    } else if (active || all) {
      const bool stereo = flags[i] & VIEO_EDGE_STEREO;
      double e[3];
      reproj_error(cam, cp[es[i]], ba_world_point(X, ep[i], sc), obs + 3 * (size_t)i, stereo, e);
      const double wi = (double)w[i];
      double c = 0;
      for (int k = 0; k < (stereo ? 3 : 2); ++k) c += e[k] * (wi * e[k]);
      chi2[i] = c;
      if (active) {
        double r1;
        huber_rho(edge_huber_delta(flags[i], lvl[i], dm, ds), c, r0, r1);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r0 += __shfl_xor_sync(0xffffffffu, r0, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = r0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int k = 0; k < 8; ++k) a += s_w[k];
    partial[blockIdx.x] = a;
  }
}

// EdgeNavStateBias between the fixed prior-bias vertex (values in d.info[6..12)) and keyframe si's bias vertex
__device__ __forceinline__ void ba_prior_bias_error(const BaDense& d, const NavS& c, double* err) {
  err[0] = (c.bg.x + c.dbg.x) - d.info[6];
  err[1] = (c.bg.y + c.dbg.y) - d.info[7];
  err[2] = (c.bg.z + c.dbg.z) - d.info[8];
  err[3] = (c.ba.x + c.dba.x) - d.info[9];
  err[4] = (c.ba.y + c.dba.y) - d.info[10];
  err[5] = (c.ba.z + c.dba.z) - d.info[11];
}

// errors of the inertial / bias edges (one thread each) and the total robust chi2 (dense first, then the visual
// partial sums in block order) -> *out.  Also sums the landmark part of the gain-ratio denominator when asked.
__global__ void __launch_bounds__(128) k_ba_dense_errors(const BaDense* __restrict__ den, int n_den, const VieoNavState* __restrict__ st,
                                  const VieoImuPreint* __restrict__ pre, const BaParams* __restrict__ prmq, BaDenseWork* __restrict__ wk,
                                  const double* __restrict__ partial, int n_partial, double* __restrict__ out,
                                  const double* __restrict__ scale_part, int n_scale, double* __restrict__ scale_out,
                                  int skip_if_visual = 0) {
  if (n_den < 0) {  // graph form: sizes from the device-resident problem header
    if (skip_if_visual && prmq->visual_only) return;
    n_den = prmq->n_den;
    n_partial = (prmq->E + 255) / 256;
  }
  const Vec3 gw = ba_gravity(*prmq);
  const bool sum_only = n_scale < 0;  // keep the stored errors (activeRobustChi2 without computeActiveErrors)
  for (int m = threadIdx.x; m < n_den && !sum_only; m += blockDim.x) {
    const BaDense& d = den[m];
    BaDenseWork& W = wk[m];
    const NavS a = ns_load(st[d.si]), b = ns_load(st[d.sj]);
    double c = 0;
    if (d.type == 0) {
      navstate_error(a, b, pre[d.pre], gw, true, W.err);
      for (int i = 0; i < 9; ++i) {
        double t = 0;
        for (int j = 0; j < 9; ++j) t += d.info[i * 9 + j] * W.err[j];
        c += W.err[i] * t;
      }
    } else if (d.type == 3) {
      ba_prior_bias_error(d, a, W.err);
      for (int i = 0; i < 6; ++i) c += W.err[i] * (d.info[i] * W.err[i]);
    } else {
      W.err[0] = (b.bg.x + b.dbg.x) - (a.bg.x + a.dbg.x);
      W.err[1] = (b.bg.y + b.dbg.y) - (a.bg.y + a.dbg.y);
      W.err[2] = (b.bg.z + b.dbg.z) - (a.bg.z + a.dbg.z);
      W.err[3] = (b.ba.x + b.dba.x) - (a.ba.x + a.dba.x);
      W.err[4] = (b.ba.y + b.dba.y) - (a.ba.y + a.dba.y);
      W.err[5] = (b.ba.z + b.dba.z) - (a.ba.z + a.dba.z);
      for (int i = 0; i < 6; ++i) c += W.err[i] * (d.info[i] * W.err[i]);
    }
    W.chi2 = c;
    double r0;
    huber_rho(d.delta, c, r0, W.r1);
    W.rho0 = r0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int k = 0; k < n_den; ++k) tot += wk[k].rho0;
    for (int k = 0; k < n_partial; ++k) tot += partial[k];
    *out = tot;
    if (scale_out) {
      double s = 0;
      for (int k = 0; k < n_scale; ++k) s += scale_part[k];
      *scale_out = s;
    }
  }
}

// EdgeNavStatePRV Jacobians written straight into the edge's 9 x 24 strip [Ji(PR) | Jj(PR) | Ji(V) | Jj(V) | Jb]
// (same arithmetic as navstate_jac with prv = true)
template <class Pre>
__device__ void navstate_jac24(const NavS& si, const NavS& sj, const Pre& m, const Vec3& gw, const double e[9], double* J,
                               const BaParams* gprm = nullptr) {
  const Mat3 RiT = m3_t(q_matrix(si.q)), Rj = q_matrix(sj.q);
  const double dt = m.dt;
  for (int i = 0; i < 9 * kDJ; ++i) J[i] = 0;
  const Mat3 JgR = ld_m3(m.JgR);
  // column bases: Ji P 0, R 3, V 12; Jj P 6, R 9, V 15; Jb 18
  Vec3 a = {sj.p.x - si.p.x - si.v.x * dt - gw.x * (dt * dt / 2), sj.p.y - si.p.y - si.v.y * dt - gw.y * (dt * dt / 2),
            sj.p.z - si.p.z - si.v.z * dt - gw.z * (dt * dt / 2)};
  Vec3 b = m3_mulv(RiT, a);
  setb(J, kDJ, 0, 3, m3_hat(b));
  setb(J, kDJ, 0, 0, m3_scale(m3_identity(), -1.0));
  setb(J, kDJ, 0, 12, m3_scale(m3_scale(RiT, -1.0), dt));
  setb(J, kDJ, 0, 18, m3_scale(ld_m3(m.Jgp), -1.0));
  setb(J, kDJ, 0, 21, m3_scale(ld_m3(m.Jap), -1.0));
  setb(J, kDJ, 0, 6, m3_mul(RiT, Rj));
  a = {sj.v.x - si.v.x - gw.x * dt, sj.v.y - si.v.y - gw.y * dt, sj.v.z - si.v.z - gw.z * dt};
  b = m3_mulv(RiT, a);
  setb(J, kDJ, 6, 3, m3_hat(b));
  setb(J, kDJ, 6, 12, m3_scale(RiT, -1.0));
  setb(J, kDJ, 6, 18, m3_scale(ld_m3(m.Jgv), -1.0));
  setb(J, kDJ, 6, 21, m3_scale(ld_m3(m.Jav), -1.0));
  setb(J, kDJ, 6, 15, RiT);
  const Vec3 eR = ld3(e + 3);
  const Mat3 Jrinv = so3_JrInv(eR);
  const Mat3 RjTRi = q_matrix(q_normalized(q_mul(q_conj(sj.q), si.q)));
  setb(J, kDJ, 3, 3, m3_scale(m3_mul(Jrinv, RjTRi), -1.0));
  const Vec3 w = m3_mulv(JgR, si.dbg);
  const Mat3 Tm = m3_mul(m3_mul(m3_mul(m3_scale(Jrinv, -1.0), so3_Exp({-eR.x, -eR.y, -eR.z})), so3_Jr(w)), JgR);
  setb(J, kDJ, 3, 18, Tm);
  setb(J, kDJ, 3, 9, Jrinv);
  if (gprm && gprm->has_g) {
    // EdgeNavStatePRVG, JG (g2otypes.h:868-876): P rows RiT dt^2/2 RwI GI^[:, 0:2], V rows RiT dt RwI GI^[:, 0:2], R rows 0
    const Mat3 RwI = q_matrix({gprm->qwI[0], gprm->qwI[1], gprm->qwI[2], gprm->qwI[3]});
    const Mat3 A = m3_mul(RiT, m3_mul(RwI, m3_hat({gprm->GI[0], gprm->GI[1], gprm->GI[2]})));
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 2; ++c) {
        J[r * kDJ + 24 + c] = A.m[3 * r + c] * (dt * dt / 2.0);
        J[(6 + r) * kDJ + 24 + c] = A.m[3 * r + c] * dt;
      }
  }
}

// accumulate one inertial / bias edge's (J^T rho' Omega) J block into H / b with the threads [t0, t0 + tn) of the block
// off_g >= 0 (gravity-direction vertex): its cross blocks with the edge's keyframes go to H like the others (edges of a
// colour share no keyframe), its own 2 x 2 block and rhs — shared by EVERY inertial edge — into W.gg, summed afterwards.
__device__ __forceinline__ void ba_dense_add_edge(const BaDense& d, BaDenseWork& W, const int* __restrict__ off0,
                                                  const int* __restrict__ off1, const int* __restrict__ off2, int np,
                                                  double* __restrict__ H, double* __restrict__ b, int t0, int tn,
                                                  int off_g = -1) {
    const int tl = (int)threadIdx.x - t0;
    if (tl < 0 || tl >= tn) return;
    if (d.type == 0) {
      const int offs[6] = {off0[d.si], off0[d.sj], off1[d.si], off1[d.sj], off2[d.si], off_g};
      const int base[7] = {0, 6, 12, 15, 18, 24, 26};
      auto gcol = [&](int lc) {
        int blk = lc < 6 ? 0 : lc < 12 ? 1 : lc < 15 ? 2 : lc < 18 ? 3 : lc < 24 ? 4 : 5;
        return offs[blk] < 0 ? -1 : offs[blk] + (lc - base[blk]);
      };
      const int nc = off_g >= 0 ? kDJ : 24;
      for (int t = tl; t < nc * (nc + 1); t += tn) {
        const int r = t / (nc + 1), c = t % (nc + 1);
        const int gr = gcol(r);
        if (gr < 0) continue;
        if (c == nc) {
          double sum = 0;
          for (int i = 0; i < 9; ++i) sum += W.J[i * kDJ + r] * W.oe[i];
          if (r >= 24) W.gg[4 + r - 24] = sum;
          else b[gr] += sum;
          continue;
        }
        const int gc = gcol(c);
        if (gc < 0) continue;
        double sum = 0;
        for (int j = 0; j < 9; ++j) sum += W.AtO[r * 9 + j] * W.J[j * kDJ + c];
        if (r >= 24 && c >= 24) W.gg[2 * (r - 24) + c - 24] = sum;
        else H[(size_t)gr * np + gc] += sum;
      }
    } else if (d.type == 3) {
      const int oi = off2[d.si];
      if (tl < 6 && oi >= 0) {
        H[(size_t)(oi + tl) * np + oi + tl] += W.r1 * d.info[tl];
        b[oi + tl] += W.oe[tl];
      }
    } else {
      const int oi = off2[d.si], oj = off2[d.sj];
      if (tl < 6) {
        const int k = tl;
        const double om = W.r1 * d.info[k], oe = W.oe[k];
        if (oj >= 0) {
          H[(size_t)(oj + k) * np + oj + k] += om;
          b[oj + k] += oe;
        }
        if (oi >= 0) {
          H[(size_t)(oi + k) * np + oi + k] += om;
          b[oi + k] += -oe;
          if (oj >= 0) {
            H[(size_t)(oi + k) * np + oj + k] += -om;
            H[(size_t)(oj + k) * np + oi + k] += -om;
          }
        }
      }
    }
  }

// Inertial and bias edges (one block).  Phase 0: zero H / b.  Phase 1 (one thread per edge): residual and, for IMU
// edges, the Jacobian strip.  Phase 2 (all threads): Omega e, then chi2 / Huber weight per edge, then J^T (rho' Omega)
// for every IMU edge.  Phase 3: the block walks the edges in order and adds (J^T rho' Omega) J to the mapped positions
// of H / b (consecutive edges share keyframes, so the order is fixed and serial; every sum runs in the oracle's order).
__device__ void ba_dense_block(const BaDense* __restrict__ den, int n_den, const VieoNavState* __restrict__ st,
                               const VieoImuPreint* __restrict__ pre, Vec3 gw, BaDenseWork* __restrict__ wk,
                               const int* __restrict__ off0, const int* __restrict__ off1, const int* __restrict__ off2,
                               int np, double* __restrict__ H, double* __restrict__ b, double* __restrict__ chi_dense,
                               bool eval_only = false, const BaParams* gprm = nullptr, int n_colors = 0) {
  const int T = blockDim.x;
  // the pre-integrations' fields the edges read (61 doubles each) staged in shared memory: the residual / Jacobian
  // code is one long dependent chain per thread, global-memory latency on every field would dominate it
  constexpr int kStage = 32;
  __shared__ VieoImuPreintLite s_pre[kStage];
  for (int t = threadIdx.x; t < n_den * 61; t += T) {
    const int m = t / 61, e = t % 61;
    if (m >= kStage || den[m].type != 0) continue;
    const VieoImuPreint& P = pre[den[m].pre];
    VieoImuPreintLite& Q = s_pre[m];
    if (e < 9) Q.Rij[e] = P.Rij[e];
    else if (e < 12) Q.vij[e - 9] = P.vij[e - 9];
    else if (e < 15) Q.pij[e - 12] = P.pij[e - 12];
    else if (e < 24) Q.Jgp[e - 15] = P.Jgp[e - 15];
    else if (e < 33) Q.Jap[e - 24] = P.Jap[e - 24];
    else if (e < 42) Q.Jgv[e - 33] = P.Jgv[e - 33];
    else if (e < 51) Q.Jav[e - 42] = P.Jav[e - 42];
    else if (e < 60) Q.JgR[e - 51] = P.JgR[e - 51];
    else Q.dt = P.dt;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_den; m += T) {
    const BaDense& d = den[m];
    BaDenseWork& W = wk[m];
    const NavS a = ns_load(st[d.si]), c = ns_load(st[d.sj]);
    if (d.type == 0) {
      if (m < kStage) {
        navstate_error(a, c, s_pre[m], gw, true, W.err);
        navstate_jac24(a, c, s_pre[m], gw, W.err, W.J, gprm);
      } else {
        navstate_error(a, c, pre[d.pre], gw, true, W.err);
        navstate_jac24(a, c, pre[d.pre], gw, W.err, W.J, gprm);
      }
    } else if (d.type == 3) {
      ba_prior_bias_error(d, a, W.err);
    } else {
      W.err[0] = (c.bg.x + c.dbg.x) - (a.bg.x + a.dbg.x);
      W.err[1] = (c.bg.y + c.dbg.y) - (a.bg.y + a.dbg.y);
      W.err[2] = (c.bg.z + c.dbg.z) - (a.bg.z + a.dbg.z);
      W.err[3] = (c.ba.x + c.dba.x) - (a.ba.x + a.dba.x);
      W.err[4] = (c.ba.y + c.dba.y) - (a.ba.y + a.dba.y);
      W.err[5] = (c.ba.z + c.dba.z) - (a.ba.z + a.dba.z);
    }
  }
  if (!eval_only) {
    for (int t = threadIdx.x; t < np * np; t += T) H[t] = 0;
    for (int t = threadIdx.x; t < np; t += T) b[t] = 0;
  }
  __syncthreads();
  // Omega e (row i of edge m)
  for (int t = threadIdx.x; t < n_den * 9; t += T) {
    const int m = t / 9, i = t % 9;
    const BaDense& d = den[m];
    double q = 0;
    if (d.type == 0)
      for (int j = 0; j < 9; ++j) q += d.info[i * 9 + j] * wk[m].err[j];
    else if (i < 6)
      q = d.info[i] * wk[m].err[i];
    wk[m].oe[i] = q;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_den; m += T) {
    BaDenseWork& W = wk[m];
    const int D = den[m].type == 0 ? 9 : 6;
    double chi = 0;
    for (int i = 0; i < D; ++i) chi += W.err[i] * W.oe[i];
    W.chi2 = chi;
    huber_rho(den[m].delta, chi, W.rho0, W.r1);
  }
  __syncthreads();
  // AtO[a][j] = sum_i J[i][a] (r1 Omega[i][j]) for IMU edges; oe <- -(Omega e) r1
  for (int t = threadIdx.x; t < n_den * kDJ * 9; t += T) {
    const int m = t / (kDJ * 9), e = t % (kDJ * 9), a = e / 9, j = e % 9;
    const BaDense& d = den[m];
    if (d.type != 0) continue;
    const BaDenseWork& W = wk[m];
    double q = 0;
    for (int i = 0; i < 9; ++i) q += W.J[i * kDJ + a] * (W.r1 * d.info[i * 9 + j]);
    wk[m].AtO[e] = q;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_den * 9; t += T) {
    const int m = t / 9, i = t % 9;
    wk[m].oe[i] = -wk[m].oe[i] * wk[m].r1;
  }
  __syncthreads();
  if (eval_only) return;  // global BA: accumulation by k_gba_dense_accum, one colour of edges per launch
  // local windows: edges of one colour share no keyframe (set_problem's greedy colouring, two colours for the inertial chain
  // and two for its bias-walk edges), so a colour's edges are added back to back by the whole block without a barrier between
  // them and only the colours are separated: 4 barriers instead of one per edge.  An entry shared by two edges now sums them in colour order, not
  // edge order (a last-bit difference, inside the 1e-6 chi2 tolerance the local BA is held to).
  if (n_colors > 0) {
    for (int c = 0; c < n_colors; ++c) {
      for (int m = 0; m < n_den; ++m)
        if (den[m].color == c) ba_dense_add_edge(den[m], wk[m], off0, off1, off2, np, H, b, 0, T);  // no barrier inside a colour
      __syncthreads();
    }
  } else {
    for (int m = 0; m < n_den; ++m) {
      ba_dense_add_edge(den[m], wk[m], off0, off1, off2, np, H, b, 0, T);
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int m = 0; m < n_den; ++m) tot += wk[m].rho0;
    *chi_dense = tot;
  }
}

struct BaBuf {  // device pointers of one handle (constant for its lifetime)
  BaParams* prm;
  VieoNavState *st, *st_bak;
  CamPose* cp;
  double *X, *X_bak, *chi2, *A, *Dinv, *db, *S, *bs, *bsys, *x, *partial, *scale_part, *part;
  double *W[2], *Hll[2], *bl[2], *H[2], *b[2];
  // scale vertex (global-BA handles only): per point wsp = sum of its edges' Js^T (w Omega) JX (the point's 1 x 3 block of
  // the scale row, per linearisation set), up = Dinv wsp, sred = (wsp . up, wsp . db); per edge As = [Hps 6 | Hss | bs];
  // spart: per linearize block partial sums of (Hss, bs)
  double *wsp[2], *up, *sred, *As, *spart;
  // V / Bias elimination (global-BA handles): compact system cS (cn x cn) / cbs / cx, the chain's Cholesky factors vbL
  // (diagonal blocks, 9 x 9 lower) and vbF (sub-diagonal blocks), Z = Syy^-1 [Syp | by] (9 Ky x (cn + 1)); index maps
  double *cS, *cbs, *cx, *vbL, *vbF, *Z;
  const int *pmap, *ymap, *ylo, *yhi;
  uint8_t* pt_active[2];
  const int *es, *ep, *pt_ptr, *off0, *off1, *off2, *prcol, *free_state, *free_off, *ps_ptr, *ps_edges;
  const int* ecol;  // per edge: prcol[es[edge]] (the free-keyframe column of the edge's keyframe, -1 when fixed)
  const float *obs, *w;
  const uint8_t *flags, *lvl, *sfix;
  const VieoImuPreint* pre;
  const BaDense* den;
  BaDenseWork* wk;
};

// computeActiveErrors + linearizeOplus + constructQuadraticForm of the visual edges, one warp per map point: each lane
// evaluates one reprojection edge (residual, chi2, Huber weight, Jacobians), the 3x3 Hll / bl are reduced with warp
// shuffles, W (Hpl) and A (the edge's Hpp / b part) are stored per edge.  The extra last block does the inertial edges.
// into_other: write the set that does NOT belong to the current estimate (speculative linearisation of a trial).
// kScale: the visual edges are EdgeReprojectPRS[Stereo] (src/Optimizer.cc:1132-1200; g2otypes.h:517-521): Xw = sc * X,
// J_point = (Jproj Rcw) sc, J_scale = (Jproj Rcw) X — a separate instantiation so that the local windows' kernel keeps
// its register budget.
template <bool kScale>
__global__ void __launch_bounds__(kBaWarps * 32, 2) k_ba_linearize_t(BaBuf B, int into_other) {
  __shared__ double s_chi[kBaWarps];
  __shared__ double s_ss[kBaWarps][2];
  const BaParams& prm = *B.prm;
  if (prm.done) return;
  const int set = into_other ? 1 - prm.cur : prm.cur;
  if (blockIdx.x == gridDim.x - 1) {
    if (!prm.big)
      ba_dense_block(B.den, prm.n_den, B.st, B.pre, prm.gw, B.wk, B.off0, B.off1, B.off2, prm.np, B.H[set], B.b[set],
                     &B.prm->chi_dense, false, nullptr, prm.n_colors);
    return;
  }
  if ((int)blockIdx.x >= prm.n_pblk) return;
  // a map point has ~6 observations: four points per warp, eight lanes each (a point with more edges takes more turns),
  // so the blocks past ceil(P / 32) of the (warp-per-point sized) grid have nothing to do but report a zero chi2 part
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & (kLinLanes - 1);
  if ((int)blockIdx.x * kBaWarps * kLinPts >= prm.P) {
    if (threadIdx.x == 0) {
      B.partial[blockIdx.x] = 0;
      if (kScale) B.spart[2 * blockIdx.x] = B.spart[2 * blockIdx.x + 1] = 0;
    }
    return;
  }
  const int p = (blockIdx.x * kBaWarps + warp) * kLinPts + lane / kLinLanes;
  double* Wb = B.W[set];
  constexpr int kAcc = kScale ? 12 : 9;
  double acc[kAcc], rsum = 0, hss = 0, bss = 0;
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0;
  const double sc = kScale ? prm.sc : 1.0;
  bool any = false;
  if (p < prm.P) {
    const int i0 = B.pt_ptr[p], i1 = B.pt_ptr[p + 1];
    const Vec3 Xh = ld3(B.X + 3 * (size_t)p);
    const Vec3 Xp = kScale ? Vec3{Xh.x * sc, Xh.y * sc, Xh.z * sc} : Xh;
    for (int i = i0 + sub; i < i1; i += kLinLanes) {
      double* Wi = Wb + 18 * (size_t)i;
      double* Ai = B.A + 27 * (size_t)i;
      if (B.lvl[i] & 1) {
        for (int k = 0; k < 18; ++k) Wi[k] = 0;
        for (int k = 0; k < 27; ++k) Ai[k] = 0;
        if (kScale)
          for (int k = 0; k < 8; ++k) B.As[8 * (size_t)i + k] = 0;
        continue;
      }
      any = true;
      const bool stereo = B.flags[i] & VIEO_EDGE_STEREO;
      const int DE = stereo ? 3 : 2;
      const int s = B.es[i];
      double e[3];
      reproj_error(prm.cam, B.cp[s], Xp, B.obs + 3 * (size_t)i, stereo, e);
      const double wi = (double)B.w[i];
      double c = 0;
      for (int k = 0; k < DE; ++k) c += e[k] * (wi * e[k]);
      B.chi2[i] = c;
      Mat3 Jp, Jr, JX;
      reproj_jac(prm.cam, B.cp[s], Xp, stereo, Jp, Jr, JX);
      double Js[3] = {0, 0, 0};
      if (kScale) {
        for (int k = 0; k < DE; ++k) Js[k] = JX.m[3 * k] * Xh.x + JX.m[3 * k + 1] * Xh.y + JX.m[3 * k + 2] * Xh.z;
#pragma unroll
        for (int k = 0; k < 9; ++k) JX.m[k] *= sc;
      }
      double r0, r1;
      huber_rho(edge_huber_delta(B.flags[i], B.lvl[i], prm.dm, prm.ds), c, r0, r1);
      rsum += r0;
      const double ww = r1 * wi;
      double oe[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) oe[k] = -(wi * e[k]) * r1;
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
      for (int r = 0; r < 3; ++r) {
        double sb = 0;
        for (int k = 0; k < DE; ++k) sb += JX.m[3 * k + r] * oe[k];
        acc[6 + r] += sb;
#pragma unroll
        for (int cc = r; cc < 3; ++cc) {
          double hh = 0;
          for (int k = 0; k < DE; ++k) hh += (JX.m[3 * k + r] * ww) * JX.m[3 * k + cc];
          acc[q++] += hh;
        }
      }
      if (kScale) {
        double* Asi = B.As + 8 * (size_t)i;
        double sb = 0, sh = 0;
        for (int k = 0; k < DE; ++k) {
          sb += Js[k] * oe[k];
          sh += (Js[k] * ww) * Js[k];
        }
        bss += sb;
        hss += sh;
        Asi[6] = sh;
        Asi[7] = sb;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          double hh = 0;
          for (int k = 0; k < DE; ++k) hh += (Js[k] * ww) * JX.m[3 * k + cc];
          acc[9 + cc] += hh;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double hh = 0;
          if (!B.sfix[s])
            for (int k = 0; k < DE; ++k) hh += (J[k][r] * ww) * Js[k];
          Asi[r] = hh;
        }
      }
      if (!B.sfix[s]) {
        q = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double sb = 0;
          for (int k = 0; k < DE; ++k) sb += J[k][r] * oe[k];
          Ai[21 + r] = sb;
#pragma unroll
          for (int cc = r; cc < 6; ++cc) {
            double hh = 0;
            for (int k = 0; k < DE; ++k) hh += (J[k][r] * ww) * J[k][cc];
            Ai[q++] = hh;
          }
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            double hh = 0;
            for (int k = 0; k < DE; ++k) hh += (J[k][r] * ww) * JX.m[3 * k + cc];
            Wi[3 * r + cc] = hh;
          }
        }
      } else {
        for (int k = 0; k < 18; ++k) Wi[k] = 0;
        for (int k = 0; k < 27; ++k) Ai[k] = 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kAcc; ++k)
#pragma unroll
    for (int o = kLinLanes / 2; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
  if (kScale) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      hss += __shfl_xor_sync(0xffffffffu, hss, o);
      bss += __shfl_xor_sync(0xffffffffu, bss, o);
    }
    if (lane == 0) {
      s_ss[warp][0] = hss;
      s_ss[warp][1] = bss;
    }
  }
  any = ((__ballot_sync(0xffffffffu, any) >> (lane & ~(kLinLanes - 1))) & ((1u << kLinLanes) - 1)) != 0;
  if (lane == 0) s_chi[warp] = rsum;
  if (sub == 0) {
    if (p < prm.P) {
      double* H = B.Hll[set] + 9 * (size_t)p;
      H[0] = acc[0]; H[1] = acc[1]; H[2] = acc[2];
      H[3] = acc[1]; H[4] = acc[3]; H[5] = acc[4];
      H[6] = acc[2]; H[7] = acc[4]; H[8] = acc[5];
      double* bb = B.bl[set] + 3 * (size_t)p;
      bb[0] = acc[6]; bb[1] = acc[7]; bb[2] = acc[8];
      B.pt_active[set][p] = any;
      if (kScale) {
        double* ws = B.wsp[set] + 3 * (size_t)p;
        ws[0] = acc[kAcc - 3]; ws[1] = acc[kAcc - 2]; ws[2] = acc[kAcc - 1];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < kBaWarps; ++k) t += s_chi[k];
    B.partial[blockIdx.x] = t;
    if (kScale) {
      double h = 0, b2 = 0;
      for (int k = 0; k < kBaWarps; ++k) {
        h += s_ss[k][0];
        b2 += s_ss[k][1];
      }
      B.spart[2 * blockIdx.x] = h;
      B.spart[2 * blockIdx.x + 1] = b2;
    }
  }
}

// One block (256 threads) per free keyframe: fixed-order sum of its edges' A blocks, added to the 6x6 diagonal block
// and rhs (the inertial part is already there).  Block 0 also totals the robust chi2 (dense edges first, then the
// visual partial sums in block order) and the landmark part of the gain-ratio denominator of the last solve.
__global__ void __launch_bounds__(256) k_ba_pose_reduce(BaBuf B, int into_other) {
  __shared__ double s_w[8][27];
  __shared__ double s_s[256];
  const BaParams& prm = *B.prm;
  if (prm.done) return;
  const int f = blockIdx.x;
  const bool have_kf = f < prm.nfree;  // block 0 also totals the step's scalars, free keyframes or not
  if (!have_kf && f != 0) return;
  const int set = into_other ? 1 - prm.cur : prm.cur;
  const int np = prm.np;
  double* H = B.H[set];
  double* b = B.b[set];
  const int o = have_kf ? B.off0[B.free_state[f]] : 0;
  const int e0 = have_kf ? B.ps_ptr[f] : 0, e1 = have_kf ? B.ps_ptr[f + 1] : 0;
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0;
  for (int t = e0 + threadIdx.x; t < e1; t += 256) {
    const double* Ai = B.A + 27 * (size_t)B.ps_edges[t];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] += Ai[q];
  }
#pragma unroll
  for (int q = 0; q < 27; ++q)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], s);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int q = 0; q < 27; ++q) s_w[threadIdx.x >> 5][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 27) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += s_w[w][threadIdx.x];
    s_w[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0 && have_kf) {
    int q = 0;
    for (int a = 0; a < 6; ++a) {
      b[o + a] += s_w[0][21 + a];
      for (int c = a; c < 6; ++c) {
        const double h = s_w[0][q++];
        H[(size_t)(o + a) * np + o + c] += h;
        if (c != a) H[(size_t)(o + c) * np + o + a] += h;
      }
    }
  }
  if (prm.has_scale) {  // the keyframe's 6 x 1 block of the scale column: fixed-order sum of its edges' Hps
    __syncthreads();
    double hp[6] = {0, 0, 0, 0, 0, 0};
    for (int t = e0 + threadIdx.x; t < e1; t += 256) {
      const double* Asi = B.As + 8 * (size_t)B.ps_edges[t];
#pragma unroll
      for (int q = 0; q < 6; ++q) hp[q] += Asi[q];
    }
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) hp[q] += __shfl_xor_sync(0xffffffffu, hp[q], s);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int q = 0; q < 6; ++q) s_w[threadIdx.x >> 5][q] = hp[q];
    __syncthreads();
    if (threadIdx.x < 6 && have_kf) {
      double v = 0;
      for (int w = 0; w < 8; ++w) v += s_w[w][threadIdx.x];
      H[(size_t)(o + threadIdx.x) * np + prm.off_s] += v;
      H[(size_t)prm.off_s * np + o + threadIdx.x] += v;
    }
  }
  if (f != 0) return;
  if (threadIdx.x == 0 && prm.has_scale) {  // scale diagonal / rhs: the linearize blocks' partial sums in block order
    double hs = 0, bs2 = 0;
    for (int q = 0; q < prm.n_pblk; ++q) {
      hs += B.spart[2 * q];
      bs2 += B.spart[2 * q + 1];
    }
    H[(size_t)prm.off_s * np + prm.off_s] += hs;
    b[prm.off_s] += bs2;
  }
  if (threadIdx.x == 32 && prm.has_g) {  // gravity-direction 2 x 2 block / rhs: the inertial edges' parts in edge order
    double g[6] = {0, 0, 0, 0, 0, 0};
    for (int m = 0; m < prm.n_den; ++m)
      if (B.den[m].type == 0)
        for (int q = 0; q < 6; ++q) g[q] += B.wk[m].gg[q];
    for (int r = 0; r < 2; ++r) {
      for (int c = 0; c < 2; ++c) H[(size_t)(prm.off_g + r) * np + prm.off_g + c] += g[2 * r + c];
      b[prm.off_g + r] += g[4 + r];
    }
  }
  double sc = 0;
  for (int t = threadIdx.x; t < prm.P; t += 256) sc += B.scale_part[t];
  s_s[threadIdx.x] = sc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) s_s[threadIdx.x] += s_s[threadIdx.x + st];
    __syncthreads();
  }
  const double scale_total = s_s[0];
  __syncthreads();
  // robust chi2: dense edges first, then the per-block partial sums in block order (chunks of 256 staged in smem)
  double tot = prm.chi_dense;
  for (int base = 0; base < prm.n_pblk; base += 256) {
    const int t = base + threadIdx.x;
    s_s[threadIdx.x] = t < prm.n_pblk ? B.partial[t] : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
      const int m = min(256, prm.n_pblk - base);
      for (int q = 0; q < m; ++q) tot += s_s[q];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    B.prm->pair[0] = tot;
    B.prm->pair[1] = scale_total;
    B.prm->pair[2] = prm.stop ? 1.0 : 0.0;  // summed over ranks: the abort takes effect at the same trial everywhere
  }
}

// max |diag| of Hpp and of the active Hll (computeLambdaInit), by one block of 256 threads
__device__ double ba_maxdiag(const double* __restrict__ H, int np, const double* __restrict__ Hll,
                             const uint8_t* __restrict__ pt_active, int P, double* s) {
  double mx = 0;
  for (int i = threadIdx.x; i < np; i += 256) mx = fmax(mx, fabs(H[(size_t)i * np + i]));
  for (int p = threadIdx.x; p < P; p += 256)
    if (pt_active[p])
      for (int k = 0; k < 3; ++k) mx = fmax(mx, fabs(Hll[9 * (size_t)p + 4 * k]));
  s[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
    __syncthreads();
  }
  return s[0];
}

// Sharded computeLambdaInit: the pose diagonal is a sum over ranks and the landmark maximum a maximum over ranks, but the
// exchange hook only sums.  Every rank writes [its part of diag(Hpp) | its landmark maximum in slot `rank`, zeros in the
// other slots] into the (idle) S buffer; after one all-reduce, the maximum of the buffer is the global max |diagonal|.
__global__ void __launch_bounds__(256) k_ba_diag_pack(BaBuf B) {
  __shared__ double s[256];
  const BaParams& prm = *B.prm;
  if (prm.done) return;
  const int np = prm.np, set = prm.cur;
  for (int i = threadIdx.x; i < np; i += 256) B.S[i] = B.H[set][(size_t)i * np + i];
  double mx = 0;
  for (int p = threadIdx.x; p < prm.P; p += 256)
    if (B.pt_active[set][p])
      for (int k = 0; k < 3; ++k) mx = fmax(mx, fabs(B.Hll[set][9 * (size_t)p + 4 * k]));
  s[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
    __syncthreads();
  }
  for (int r = threadIdx.x; r < prm.world; r += 256) B.S[np + r] = r == prm.rank ? s[0] : 0.0;
}

// Start of an optimize() stage without a host round trip (asynchronous LocalBA): resets the device-side state machine.
// stage > 0 keeps a raised abort flag (the reference skips the second stage once pbStopFlag is set, src/Optimizer.cc:589-593)
// and files the previous stage's iteration count.
__global__ void k_ba_begin(BaBuf B, int stage, int iterations, double lambda_init) {
  BaParams& q = *B.prm;
  if (iterations < 0) {  // graph form: the iteration plan travels in the problem header
    iterations = q.plan_it[stage];
    lambda_init = q.plan_lambda0;
  }
  if (stage > 0) q.iters_hist[stage - 1] = q.iters_done;
  else q.stop = 0;
  q.cur = 0; q.done = 0; q.ok = 1;
  q.iteration = 0; q.iters_target = iterations; q.iters_done = 0; q.qmax = 0; q.nBad = 0; q.trials = 0;
  q.user_lambda = lambda_init;
  if (q.stop) q.iters_target = 0;  // k_ba_control (mode 0) ends the stage at once
}

// Levenberg-Marquardt bookkeeping on the device (OptimizationAlgorithmLevenberg::solve, :83-166; SparseOptimizer::
// optimize loop, sparse_optimizer.cpp:376-414).  mode 0: after the initial linearisation of an optimize() call;
// mode 1: after a trial (solve + update + linearisation at the trial estimate).
__global__ void __launch_bounds__(256) k_ba_control(BaBuf B, int mode, cudaGraphConditionalHandle cond) {
  __shared__ double s[256];
  __shared__ int s_restore;
  BaParams& prm = *B.prm;
  // cond != 0: this launch is the last node of the WHILE body of the optimize() graph; the loop goes on until done
  if (prm.done) {
    if (cond && threadIdx.x == 0) cudaGraphSetConditional(cond, 0);
    return;
  }
  // the caller's abort flag (pbStopFlag): this rank's copy, or — sharded — the all-reduced count of ranks that saw it, so
  // that every rank stops enqueuing collectives at the same trial (ranks poll their host flags at different moments)
  const bool stop_now = prm.world > 1 ? prm.pair[2] > 0.0 : prm.stop != 0;
  if (mode == 0) {
    double lam = prm.user_lambda;
    if (!(lam > 0) && prm.world > 1) {  // from the all-reduced k_ba_diag_pack buffer
      double mx = 0;
      for (int i = threadIdx.x; i < prm.np + prm.world; i += 256) mx = fmax(mx, fabs(B.S[i]));
      s[threadIdx.x] = mx;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
        __syncthreads();
      }
      lam = 1e-5 * s[0];
    } else if (!(lam > 0)) lam = 1e-5 * ba_maxdiag(B.H[prm.cur], prm.np, B.Hll[prm.cur], B.pt_active[prm.cur], prm.P, s);
    if (threadIdx.x == 0) {
      prm.chi_cur = prm.ini_chi = prm.pair[0];
      prm.lambda = lam;
      prm.ni = 2;
      prm.nBad = 0;
      prm.iteration = 0;
      prm.iters_done = 0;
      prm.qmax = 0;
      if (prm.iters_target <= 0 || stop_now) prm.done = 1;
    }
    return;
  }
  if (threadIdx.x == 0) {
    double tempChi = prm.pair[0];
    if (!prm.ok) tempChi = 1.7976931348623157e308;
    double rho = prm.chi_cur - tempChi;
    double scale = prm.pscale + prm.pair[1];
    scale += 1e-3;
    rho /= scale;
    int restore = 0;
    if (rho > 0 && isfinite(tempChi)) {
      double alpha = 1. - pow((2 * rho - 1), 3);
      alpha = fmin(alpha, 2. / 3.);
      prm.lambda *= fmax(1. / 3., alpha);
      prm.ni = 2;
      prm.chi_cur = tempChi;
      prm.cur = 1 - prm.cur;  // the speculative linearisation is the current one now
    } else {
      prm.lambda *= prm.ni;
      prm.ni *= 2;
      restore = 1;
    }
    prm.qmax++;
    prm.trials++;
    const bool more_trials = rho < 0 && prm.qmax < 10 && !stop_now;
    if (!more_trials) {
      prm.iters_done++;
      bool terminate = prm.qmax == 10 || rho == 0;
      if (!terminate) {
        if ((prm.ini_chi - prm.chi_cur) * 1e3 < prm.ini_chi) prm.nBad++;
        else prm.nBad = 0;
        if (prm.nBad >= 3) terminate = true;
      }
      prm.iteration++;
      if (terminate || prm.iteration >= prm.iters_target || stop_now) prm.done = 1;
      else {
        prm.ini_chi = prm.chi_cur;
        prm.qmax = 0;
      }
    }
    s_restore = restore;
    if (cond) cudaGraphSetConditional(cond, prm.done ? 0 : 1);
  }
  __syncthreads();
  if (s_restore) {  // pop(): back to the estimate before the trial
    const int nS = prm.K * (int)(sizeof(VieoNavState) / sizeof(double)), nX = 3 * prm.P;
    const double* sb = reinterpret_cast<const double*>(B.st_bak);
    double* sd = reinterpret_cast<double*>(B.st);
    for (int t = threadIdx.x; t < nS; t += 256) sd[t] = sb[t];
    for (int t = threadIdx.x; t < nX; t += 256) B.X[t] = B.X_bak[t];
    if (threadIdx.x == 0) {
      prm.sc = prm.sc_bak;
      for (int q = 0; q < 4; ++q) prm.qwI[q] = prm.qwI_bak[q];
    }
  }
}

// Start of a solve: Dinv = (Hll + lambda I)^-1 (cofactor inverse like Eigen's fixed 3x3, block_solver.hpp:389),
// db = Dinv bl; S = H + lambda_pose I, bschur = b, sys.b = b
__global__ void k_ba_prep_solve(BaBuf B, double lambda_arg, int use_arg) {
  const BaParams& prm = *B.prm;
  if (prm.done && !use_arg) return;
  const int np = prm.np, P = prm.P, set = prm.cur;
  const double lambda = use_arg ? lambda_arg : prm.lambda;
  const double lambda_pose = prm.lambda_on_poses ? lambda : 0.0;
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < (size_t)np * np) {
    const int r = t / np, c = t % np;
    B.S[t] = B.H[set][t] + (r == c ? lambda_pose : 0.0);
  }
  if (t < (size_t)np) {
    B.bs[t] = B.b[set][t];
    B.bsys[t] = B.b[set][t];
  }
  if (t >= (size_t)P) return;
  const size_t p = t;
  double* I = B.Dinv + 9 * p;
  if (!B.pt_active[set][p]) {
    for (int k = 0; k < 9; ++k) I[k] = 0;
    B.db[3 * p] = B.db[3 * p + 1] = B.db[3 * p + 2] = 0;
    if (prm.has_scale) {
      B.up[3 * p] = B.up[3 * p + 1] = B.up[3 * p + 2] = 0;
      B.sred[2 * p] = B.sred[2 * p + 1] = 0;
    }
    return;
  }
  double D[9];
  for (int k = 0; k < 9; ++k) D[k] = B.Hll[set][9 * p + k];
  D[0] += lambda; D[4] += lambda; D[8] += lambda;
  const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
  const double det = D[0] * c00 + D[1] * c01 + D[2] * c02, id = 1.0 / det;
  I[0] = c00 * id; I[1] = (D[2] * D[7] - D[1] * D[8]) * id; I[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  I[3] = c01 * id; I[4] = (D[0] * D[8] - D[2] * D[6]) * id; I[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  I[6] = c02 * id; I[7] = (D[1] * D[6] - D[0] * D[7]) * id; I[8] = (D[0] * D[4] - D[1] * D[3]) * id;
  const double* bb = B.bl[set] + 3 * p;
  for (int a = 0; a < 3; ++a) B.db[3 * p + a] = I[3 * a] * bb[0] + I[3 * a + 1] * bb[1] + I[3 * a + 2] * bb[2];
  if (prm.has_scale) {  // the point's part of the scale row of the Schur complement: up = Dinv wsp
    const double* ws = B.wsp[set] + 3 * p;
    double u[3];
    for (int a = 0; a < 3; ++a) u[a] = I[3 * a] * ws[0] + I[3 * a + 1] * ws[1] + I[3 * a + 2] * ws[2];
    B.up[3 * p] = u[0]; B.up[3 * p + 1] = u[1]; B.up[3 * p + 2] = u[2];
    B.sred[2 * p] = ws[0] * u[0] + ws[1] * u[1] + ws[2] * u[2];
    B.sred[2 * p + 1] = ws[0] * B.db[3 * p] + ws[1] * B.db[3 * p + 1] + ws[2] * B.db[3 * p + 2];
  }
}

// Schur complement.  Block (f, s): chunk s of free keyframe f's edge list.  Each warp walks its edges in order; for
// edge a and every edge c of a's point, lane (r, slot) adds row r of (W_a Dinv) W_c^T to the warp's accumulator tile
// [6][6*nfree + 1] in shared memory (last column: W_a db).  The warps' tiles are summed in warp order into
// part[f][s]; k_ba_schur_reduce then sums the chunks in order and subtracts from S / bschur.  No atomics.
constexpr int kSchurSplit = 8;
constexpr int kSchurMaxFree = 48;  // shared-memory tile of the trial graph is sized for this many free keyframes
__global__ void __launch_bounds__(kBaWarps * 32) k_ba_schur(BaBuf B, int force) {
  extern __shared__ double s_acc[];  // [kBaWarps][6][ld]
  const BaParams& prm = *B.prm;
  if ((prm.done && !force) || (int)blockIdx.x >= prm.nfree || prm.E == 0) return;
  const int nfree = prm.nfree;
  const int ld = 6 * nfree + 1;
  const int f = blockIdx.x, sp = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* Wb = B.W[prm.cur];
  double* acc = s_acc + (size_t)warp * 6 * ld;
  for (int t = lane; t < 6 * ld; t += 32) acc[t] = 0;
  __syncwarp();
  const int t0 = B.ps_ptr[f], t1 = B.ps_ptr[f + 1];
  const int per_blk = (t1 - t0 + kSchurSplit - 1) / kSchurSplit;
  const int b0 = t0 + sp * per_blk, b1 = min(b0 + per_blk, t1);
  const int per_warp = (max(b1 - b0, 0) + kBaWarps - 1) / kBaWarps;
  const int a0 = b0 + warp * per_warp, a1 = min(a0 + per_warp, b1);
  const int r = lane % 6, slot = lane / 6;  // lanes 30, 31 idle in the tile update
  const int serial_lanes = prm.has_dup;
  for (int t = a0; t < a1; ++t) {
    const int a = B.ps_edges[t];
    const int p = B.ep[a];
    const double* Wa = Wb + 18 * (size_t)a + 3 * r;
    const double* Di = B.Dinv + 9 * (size_t)p;
    const double w0 = Wa[0], w1 = Wa[1], w2 = Wa[2];
    const double d0 = w0 * Di[0] + w1 * Di[3] + w2 * Di[6];
    const double d1 = w0 * Di[1] + w1 * Di[4] + w2 * Di[7];
    const double d2 = w0 * Di[2] + w1 * Di[5] + w2 * Di[8];
    if (lane < 6) {
      const double* d = B.db + 3 * (size_t)p;
      acc[r * ld + 6 * nfree] += w0 * d[0] + w1 * d[1] + w2 * d[2];
    }
    const int c0 = B.pt_ptr[p], c1 = B.pt_ptr[p + 1];
    for (int cb = c0; cb < c1; cb += 5) {
      const int c = cb + slot;
      const int col = (slot < 5 && c < c1) ? B.prcol[B.es[c]] : -1;
      for (int turn = 0; turn < (serial_lanes ? 5 : 1); ++turn) {
        if (col >= 0 && (!serial_lanes || turn == slot)) {
          const double* Wc = Wb + 18 * (size_t)c;
          double* dst = acc + r * ld + 6 * col;
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) dst[cc] += d0 * Wc[3 * cc] + d1 * Wc[3 * cc + 1] + d2 * Wc[3 * cc + 2];
        }
        if (serial_lanes) __syncwarp();
      }
    }
    __syncwarp();
  }
  __syncthreads();
  double* out = B.part + ((size_t)f * kSchurSplit + sp) * 6 * ld;
  for (int t = threadIdx.x; t < 6 * ld; t += kBaWarps * 32) {
    double s = 0;
    for (int w = 0; w < kBaWarps; ++w) s += s_acc[(size_t)w * 6 * ld + t];
    out[t] = s;
  }
}
__global__ void __launch_bounds__(256) k_ba_schur_reduce(BaBuf B, int force) {
  const BaParams& prm = *B.prm;
  if ((prm.done && !force) || (int)blockIdx.x >= prm.nfree || prm.E == 0) return;
  const int nfree = prm.nfree, np = prm.np;
  const int ld = 6 * nfree + 1;
  const int f = blockIdx.x, o = B.off0[B.free_state[f]];
  const double* in = B.part + (size_t)f * kSchurSplit * 6 * ld;
  for (int t = threadIdx.x; t < 6 * ld; t += 256) {
    double s = 0;
    for (int sp = 0; sp < kSchurSplit; ++sp) s += in[(size_t)sp * 6 * ld + t];
    const int r = t / ld, c = t % ld;
    if (c == 6 * nfree) B.bs[o + r] -= s;
    else B.S[(size_t)(o + r) * np + B.free_off[c / 6] + c % 6] -= s;
  }
}

// Dense Cholesky of S (n x n) + solve S x = rhs by one block.  The lower triangle is staged in shared memory when it
// fits (n <= kCholSmemN), else factorised in place in global memory.  Right-looking with panels of kCholNB columns;
// columns are kept unnormalised (s_ij = a_ij - sum l l, so a_ik -= s_ij s_kj / s_jj) and scaled by 1/sqrt(s_jj) at the
// end: one barrier per column inside a panel (cheap: one row per thread) and one rank-NB update of the trailing block
// per panel.  ok = 0 when a pivot is not positive (LinearSolverDense: LDLT::isPositive, linear_solver_dense.h:107-112).
constexpr int kCholNB = 8;
constexpr int kCholThreads = 1024;
constexpr int kCholSmemN = 165;
constexpr size_t kCholSmemBytes = sizeof(double) * ((size_t)kCholSmemN * (kCholSmemN | 1) + 3 * kCholSmemN);
__global__ void __launch_bounds__(kCholThreads) k_ba_chol(BaBuf B, int force) {
  extern __shared__ double sh[];
  __shared__ double s_inv[kCholNB];
  __shared__ int s_good;
  BaParams& prm = *B.prm;
  if (prm.done && !force) return;
  const int n = prm.vb_elim ? prm.cn : prm.np;  // the compact system when the V / Bias chain is eliminated first
  const int in_smem = n <= kCholSmemN;
  double* Ag = B.S;
  const double* rhs = B.bs;
  const int T = blockDim.x, t = threadIdx.x;
  // scratch vectors: in shared memory when the matrix is (3n doubles), else in the tail of the sys buffer's bs copy
  double* diag = in_smem ? sh : B.part;  // B.part is free between schur_reduce and the next schur
  double* y = diag + n;
  double* x = diag + 2 * n;
  double* A = in_smem ? sh + 3 * n : Ag;
  const int ld = in_smem ? (n | 1) : n;
  if (in_smem) {
    for (int i = t >> 5; i < n; i += T >> 5)
      for (int k = t & 31; k <= i; k += 32) A[i * ld + k] = Ag[(size_t)i * n + k];
  }
  for (int i = t; i < n; i += T) y[i] = rhs[i];
  if (t == 0) s_good = 1;
  __syncthreads();
  const int nw = T >> 5, wp = t >> 5, ln = t & 31;
  for (int j0 = 0; j0 < n; j0 += kCholNB) {
    const int nb = min(kCholNB, n - j0), jend = j0 + nb;
    // (1) the nb x nb diagonal block, warp 0, lane r = row j0 + r
    if (wp == 0) {
      for (int c = 0; c < nb; ++c) {
        const double sjj = A[(size_t)(j0 + c) * ld + j0 + c];
        const double inv = 1.0 / sjj;
        if (ln == 0) {
          if (!(sjj > 0) || !isfinite(sjj)) s_good = 0;
          s_inv[c] = inv;
        }
        if (ln > c && ln < nb) {
          double* Ar = A + (size_t)(j0 + ln) * ld + j0;
          const double f = Ar[c] * inv;
          for (int k = c + 1; k <= ln; ++k) Ar[k] -= f * A[(size_t)(j0 + k) * ld + j0 + c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (!s_good) break;
    // (2) rows below the panel: one thread per row eliminates its nb panel entries in registers (no barriers)
    for (int i = jend + t; i < n; i += T) {
      double* Ar = A + (size_t)i * ld + j0;
      double r[kCholNB];
#pragma unroll
      for (int c = 0; c < kCholNB; ++c) r[c] = c < nb ? Ar[c] : 0.0;
#pragma unroll
      for (int c = 0; c < kCholNB; ++c) {
        if (c < nb) {
          const double f = r[c] * s_inv[c];
#pragma unroll
          for (int k = c + 1; k < kCholNB; ++k)
            if (k < nb) r[k] -= f * A[(size_t)(j0 + k) * ld + j0 + c];
        }
      }
#pragma unroll
      for (int c = 0; c < kCholNB; ++c)
        if (c < nb) Ar[c] = r[c];
    }
    __syncthreads();
    // (3) trailing block: A[i][k] -= sum_c (s_ic / s_cc) s_kc for k >= jend, rows by warp, lanes along the row
    for (int i = jend + wp; i < n; i += nw) {
      double f[kCholNB];
#pragma unroll
      for (int c = 0; c < kCholNB; ++c) f[c] = c < nb ? A[(size_t)i * ld + j0 + c] * s_inv[c] : 0.0;
      for (int k = jend + ln; k <= i; k += 32) {
        double acc = A[(size_t)i * ld + k];
        const double* Ak = A + (size_t)k * ld + j0;
#pragma unroll
        for (int c = 0; c < kCholNB; ++c)
          if (c < nb) acc -= f[c] * Ak[c];
        A[(size_t)i * ld + k] = acc;
      }
    }
    __syncthreads();
  }
  const bool good = s_good != 0;
  if (t == 0) prm.ok = good ? 1 : 0;
  if (!good) return;
  for (int j = t; j < n; j += T) diag[j] = 1.0 / sqrt(A[(size_t)j * ld + j]);
  __syncthreads();
  // l_ij = s_ij / sqrt(s_jj)
  for (int i = wp; i < n; i += nw)
    for (int k = ln; k < i; k += 32) A[(size_t)i * ld + k] *= diag[k];
  __syncthreads();
  // forward substitution L y = b in blocks of 32 rows: warp 0 solves the 32 x 32 triangle with shuffles, then every
  // remaining row subtracts the block's contribution
  for (int b0 = 0; b0 < n; b0 += 32) {
    const int bn = min(32, n - b0);
    if (wp == 0) {
      const int i = b0 + ln;
      double yi = ln < bn ? y[i] : 0.0;
      const double rd = ln < bn ? diag[i] : 0.0;
      for (int j = 0; j < bn; ++j) {
        const double yj = __shfl_sync(0xffffffffu, yi * rd, j);
        if (ln == j) yi = yj;
        else if (ln > j && ln < bn) yi -= A[(size_t)i * ld + b0 + j] * yj;
      }
      if (ln < bn) y[i] = yi;
    }
    __syncthreads();
    for (int i = b0 + bn + t; i < n; i += T) {
      double acc = y[i];
      const double* Ar = A + (size_t)i * ld + b0;
      for (int j = 0; j < bn; ++j) acc -= Ar[j] * y[b0 + j];
      y[i] = acc;
    }
    __syncthreads();
  }
  // backward substitution L^T x = y, blocks from the bottom
  for (int b1 = n; b1 > 0; b1 -= 32) {
    const int b0 = max(b1 - 32, 0), bn = b1 - b0;
    if (wp == 0) {
      const int i = b0 + ln;
      double xi = ln < bn ? y[i] : 0.0;
      const double rd = ln < bn ? diag[i] : 0.0;
      for (int j = bn - 1; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, xi * rd, j);
        if (ln == j) xi = xj;
        else if (ln < j) xi -= A[(size_t)(b0 + j) * ld + i] * xj;
      }
      if (ln < bn) x[i] = xi;
    }
    __syncthreads();
    for (int i = t; i < b0; i += T) {
      double acc = y[i];
      for (int j = 0; j < bn; ++j) acc -= A[(size_t)(b0 + j) * ld + i] * x[b0 + j];
      y[i] = acc;
    }
    __syncthreads();
  }
  for (int i = t; i < n; i += T) B.x[i] = x[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Global-BA sized systems (Optimizer::GlobalBundleAdjustmentNavStatePRV, src/Optimizer.cc:771-1342: every keyframe of
// the map is free, np = 15 K reaches several thousand).  The reduced camera system stays DENSE in HBM (np = 6000 is
// 288 MB of 180 GB): the reference's sparse LDLT exists to save CPU memory and flops this device does not lack.
//
// Schur complement, one CTA per free keyframe row: the CTA owns the 6 x (6 nfree + 1) row tile in shared memory and
// walks the keyframe's edges in order; for edge a the whole CTA adds the (W_a Dinv) W_c^T blocks of the point's other
// observations c (one thread per entry, distinct columns, no atomics), then the tile is subtracted from S / bschur.
// ---- shared-memory pipeline primitives of k_gba_schur (mbarrier, TMA 1-D bulk copy, cp.async) --------------------------------
constexpr int kSchurRingMax = 4, kSchurChunk = 32;  // ring depth (measured: 16 slots are no faster than 4 — the walk is bound by the per-edge issue cost of producer and TMA unit, not by copy latency)
struct __align__(128) SchurSlot {
  double Wc[kSchurChunk * 18];  // the point's co-observation W blocks (bulk copy destination, 16-byte aligned)
  double Wa[18];                // the edge's own W block (bulk copy destination)
  double Di[9], db[3], up[3];   // the point's inverse Hll, Hll^-1 bl, Hll^-1 wsp
  int cols[kSchurChunk];        // free-keyframe column of every co-observation (-1: fixed keyframe)
  int n, flags;                 // co-observations in this slot; 1 = first slot of its edge, 2 = last slot of the walk
};
__host__ __device__ inline size_t schur_ring_offset(int nfree) {
  const size_t tile = sizeof(double) * (6 * (6 * (size_t)nfree + 2));
  return (tile + 127) / 128 * 128;
}
// slots behind the row tile, bounded by what is left of the CTA's 227 KB
__host__ __device__ inline int schur_ring_depth(int nfree) {
  const size_t left = 227 * 1024 - 256 - schur_ring_offset(nfree);
  const size_t per = sizeof(SchurSlot) + 16;
  const int d = (int)(left / per);
  return d < 2 ? 2 : (d > kSchurRingMax ? kSchurRingMax : d);
}
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VIEO_SW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VIEO_SD_%=;\n"
      "bra VIEO_SW_%=;\n"
      "VIEO_SD_%=:\n"
      "}\n" ::"r"(sm_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_addr(dst)),
               "l"(src), "r"(bytes), "r"(sm_addr(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sm_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sm_addr(bar)) : "memory");
}

constexpr int kGbaSchurThreads = 512;
__global__ void __launch_bounds__(kGbaSchurThreads) k_gba_schur(BaBuf B, int force) {
  extern __shared__ double s_tile[];  // [6][ld]
  const BaParams& prm = *B.prm;
  if ((prm.done && !force) || prm.E == 0) return;
  const int nfree = prm.nfree, np = prm.np, ld = 6 * nfree + 2;  // columns: keyframe blocks | W db | W up (scale column)
  const int f = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  if (f == nfree) {
    // extra block: diagonal / rhs of the scale row, S_ss -= sum_p wsp . up, bs_s -= sum_p wsp . db (fixed-order tree)
    if (!prm.has_scale) return;
    double a0 = 0, a1 = 0;
    for (int p = tid; p < prm.P; p += T) {
      a0 += B.sred[2 * (size_t)p];
      a1 += B.sred[2 * (size_t)p + 1];
    }
    s_tile[tid] = a0;
    s_tile[T + tid] = a1;
    __syncthreads();
    for (int o = T / 2; o > 0; o >>= 1) {
      if (tid < o) {
        s_tile[tid] += s_tile[tid + o];
        s_tile[T + tid] += s_tile[T + tid + o];
      }
      __syncthreads();
    }
    if (tid == 0) {
      B.S[(size_t)prm.off_s * np + prm.off_s] -= s_tile[0];
      B.bs[prm.off_s] -= s_tile[T];
    }
    return;
  }
  if (f > nfree) return;
  const bool has_scale = prm.has_scale != 0;
  const double* Wb = B.W[prm.cur];
  for (int t = tid; t < 6 * ld; t += T) s_tile[t] = 0;
  __syncthreads();
  const int t0 = B.ps_ptr[f], t1 = B.ps_ptr[f + 1];
  const int dup = prm.has_dup;
  if (!dup && t0 < t1) {
    // The walk over the keyframe's edges as a PRODUCER / CONSUMER pipeline through shared memory (two edges of a keyframe may
    // meet in a column, so the edges are consumed strictly in order, one consumer barrier per step):
    //  * warp 15 produces: it resolves the index chain of 32 edges at a time, one edge per lane (ps_edges -> ep -> pt_ptr and
    //    the point's Dinv / db / up: the five-deep dependent chain that cost the un-pipelined walk ~2300 cycles per edge is
    //    paid once per 32 edges), then fills a ring of kSchurRingMax slots, one per edge: the edge's own W block and the point's
    //    CONTIGUOUS run of co-observation W blocks (edges are sorted by point) by two bulk async copies (TMA 1-D,
    //    cp.async.bulk ... mbarrier::complete_tx), the co-observations' columns by 4-byte cp.async, the 15 doubles of the
    //    point from the owning lane's registers; a `full` mbarrier per slot collects the byte count and the lanes' arrivals;
    //  * warps 0..14 consume: wait for the slot, 12 multiply-adds per entry from shared memory into the row tile (thread e
    //    handles co-observation e / 36, row (e % 36) / 6, column e % 6), a named barrier of the 480 consumers, and one
    //    arrival on the slot's `empty` mbarrier.  A point with more than 32 co-observations takes several slots.
    // Every tile entry receives its edges in the same order with the same arithmetic as the one-barrier-per-edge form.
    SchurSlot* ring = reinterpret_cast<SchurSlot*>(reinterpret_cast<uint8_t*>(s_tile) + schur_ring_offset(nfree));
    const int kSchurRing = schur_ring_depth(nfree);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kSchurRing);
    uint64_t* empty = full + kSchurRing;
    if (tid == 0) {
      for (int k = 0; k < kSchurRing; ++k) {
        mbar_init(&full[k], 33);  // 32 lanes' cp.async arrivals + lane 0's arrive.expect_tx
        mbar_init(&empty[k], 1);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    constexpr int kConsumers = kGbaSchurThreads - 32;
    if (tid >= kConsumers) {
      // ---------------- producer warp ----------------
      const int lane = tid - kConsumers;
      int fill = 0;
      for (int tb = t0; tb < t1; tb += 32) {
        const int t = tb + lane;
        const bool valid = t < t1;
        const int a = valid ? B.ps_edges[t] : 0;
        const int p = valid ? B.ep[a] : 0;
        const int c0 = valid ? B.pt_ptr[p] : 0, c1 = valid ? B.pt_ptr[p + 1] : 0;
        double pt[15];
#pragma unroll
        for (int k = 0; k < 9; ++k) pt[k] = valid ? B.Dinv[9 * (size_t)p + k] : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          pt[9 + k] = valid ? B.db[3 * (size_t)p + k] : 0.0;
          pt[12 + k] = (valid && has_scale) ? B.up[3 * (size_t)p + k] : 0.0;
        }
        const int n_here = min(32, t1 - tb);
        for (int l = 0; l < n_here; ++l) {
          const int a_l = __shfl_sync(0xffffffffu, a, l), c0_l = __shfl_sync(0xffffffffu, c0, l), c1_l = __shfl_sync(0xffffffffu, c1, l);
          const bool last_edge = tb + l == t1 - 1;
          int cs = c0_l;
          do {  // at least one slot per edge, so the consumers always meet the walk's last slot
            const int n = max(0, min(kSchurChunk, c1_l - cs));
            const int slot = fill % kSchurRing;
            mbar_wait(&empty[slot], ((fill / kSchurRing) & 1) ^ 1);
            SchurSlot& S = ring[slot];
            if (lane == l) {
#pragma unroll
              for (int k = 0; k < 9; ++k) S.Di[k] = pt[k];
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                S.db[k] = pt[9 + k];
                S.up[k] = pt[12 + k];
              }
              S.n = n;
              S.flags = (cs == c0_l ? 1 : 0) | ((last_edge && cs + kSchurChunk >= c1_l) ? 2 : 0);  // (an empty range: one slot, both flags)
            }
            if (lane < n) cp_async4(&S.cols[lane], B.ecol + cs + lane);
            cp_async_mbar_arrive(&full[slot]);
            __syncwarp();
            if (lane == 0) {
              mbar_arrive_expect_tx(&full[slot], 144u * (unsigned)(n + 1));
              bulk_g2s(S.Wa, Wb + 18 * (size_t)a_l, 144u, &full[slot]);
              if (n > 0) bulk_g2s(S.Wc, Wb + 18 * (size_t)cs, 144u * (unsigned)n, &full[slot]);
            }
            ++fill;
            cs += kSchurChunk;
          } while (cs < c1_l);
        }
      }
    } else {
      // ---------------- consumers ----------------
      for (int k = 0;; ++k) {
        const int slot = k % kSchurRing;
        mbar_wait(&full[slot], (k / kSchurRing) & 1);
        const SchurSlot& S = ring[slot];
        const int n = S.n, fl = S.flags;
        if (fl & 1) {
          if (tid < 6) s_tile[tid * ld + 6 * nfree] += S.Wa[3 * tid] * S.db[0] + S.Wa[3 * tid + 1] * S.db[1] + S.Wa[3 * tid + 2] * S.db[2];
          else if (tid >= 32 && tid < 38 && has_scale) {
            const int r = tid - 32;
            s_tile[r * ld + 6 * nfree + 1] += S.Wa[3 * r] * S.up[0] + S.Wa[3 * r + 1] * S.up[1] + S.Wa[3 * r + 2] * S.up[2];
          }
        }
        for (int e = tid; e < 36 * n; e += kConsumers) {
          const int j = e / 36, e36 = e - 36 * j, r = e36 / 6, cc = e36 - 6 * r;
          const int col = S.cols[j];
          if (col < 0) continue;
          const double w0 = S.Wa[3 * r], w1 = S.Wa[3 * r + 1], w2 = S.Wa[3 * r + 2];
          const double d0 = w0 * S.Di[0] + w1 * S.Di[3] + w2 * S.Di[6];
          const double d1 = w0 * S.Di[1] + w1 * S.Di[4] + w2 * S.Di[7];
          const double d2 = w0 * S.Di[2] + w1 * S.Di[5] + w2 * S.Di[8];
          const double* Wc = S.Wc + 18 * j + 3 * cc;
          s_tile[r * ld + 6 * col + cc] += d0 * Wc[0] + d1 * Wc[1] + d2 * Wc[2];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
        if (tid == 0) mbar_arrive(&empty[slot]);
        if (fl & 2) break;
      }
    }
    __syncthreads();
  }
  for (int t = t0; dup && t < t1; ++t) {
    const int a = B.ps_edges[t];
    const int p = B.ep[a];
    const int c0 = B.pt_ptr[p], nc = B.pt_ptr[p + 1] - c0;
    const double* Di = B.Dinv + 9 * (size_t)p;
    const double* Wa = Wb + 18 * (size_t)a;
    if (tid < 6) {
      const double* d = B.db + 3 * (size_t)p;
      s_tile[tid * ld + 6 * nfree] += Wa[3 * tid] * d[0] + Wa[3 * tid + 1] * d[1] + Wa[3 * tid + 2] * d[2];
    } else if (tid >= 32 && tid < 38 && has_scale) {
      const int r = tid - 32;
      const double* u = B.up + 3 * (size_t)p;
      s_tile[r * ld + 6 * nfree + 1] += Wa[3 * r] * u[0] + Wa[3 * r + 1] * u[1] + Wa[3 * r + 2] * u[2];
    }
    if (!dup) {
      for (int e = tid; e < 36 * nc; e += T) {
        const int c = c0 + e / 36, r = (e % 36) / 6, cc = e % 6;
        const int col = B.prcol[B.es[c]];
        if (col < 0) continue;
        const double w0 = Wa[3 * r], w1 = Wa[3 * r + 1], w2 = Wa[3 * r + 2];
        const double d0 = w0 * Di[0] + w1 * Di[3] + w2 * Di[6];
        const double d1 = w0 * Di[1] + w1 * Di[4] + w2 * Di[7];
        const double d2 = w0 * Di[2] + w1 * Di[5] + w2 * Di[8];
        const double* Wc = Wb + 18 * (size_t)c + 3 * cc;
        s_tile[r * ld + 6 * col + cc] += d0 * Wc[0] + d1 * Wc[1] + d2 * Wc[2];
      }
    } else {
      // a point seen twice by one keyframe (multi-camera rigs): its observations may share a column, take them in turn
      for (int c = c0; c < c0 + nc; ++c) {
        const int col = B.prcol[B.es[c]];
        if (col >= 0 && tid < 36) {
          const int r = tid / 6, cc = tid % 6;
          const double w0 = Wa[3 * r], w1 = Wa[3 * r + 1], w2 = Wa[3 * r + 2];
          const double d0 = w0 * Di[0] + w1 * Di[3] + w2 * Di[6];
          const double d1 = w0 * Di[1] + w1 * Di[4] + w2 * Di[7];
          const double d2 = w0 * Di[2] + w1 * Di[5] + w2 * Di[8];
          const double* Wc = Wb + 18 * (size_t)c + 3 * cc;
          s_tile[r * ld + 6 * col + cc] += d0 * Wc[0] + d1 * Wc[1] + d2 * Wc[2];
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
  const int o = B.off0[B.free_state[f]];
  for (int t = tid; t < 6 * ld; t += T) {
    const int r = t / ld, c = t % ld;
    if (c == 6 * nfree) B.bs[o + r] -= s_tile[t];
    else if (c == 6 * nfree + 1) {
      if (has_scale) {
        B.S[(size_t)(o + r) * np + prm.off_s] -= s_tile[t];
        B.S[(size_t)prm.off_s * np + o + r] -= s_tile[t];
      }
    } else B.S[(size_t)(o + r) * np + B.free_off[c / 6] + c % 6] -= s_tile[t];
  }
}

// Blocked right-looking Cholesky of the dense reduced camera system over the whole device, panels of kGNB columns, with
// a tile-level structure map: nz[i][j] says whether the 64 x 64 tile (i, j) of the lower triangle holds anything but
// exact zeros (k_gchol_scan after the Schur complement; fill-in is propagated by k_gchol_syrk).  A tile of exact zeros
// contributes exactly nothing, so skipping it changes no bit of the result, and a map whose covisibility is mostly a
// band plus a few loop closures costs a fraction of the dense n^3 / 3.
// Per panel k: k_gchol_diag (one CTA) factorises the diagonal block in four 16-column sub-panels and advances the
// forward substitution (y_k = L_kk^-1 y_k); k_gchol_trsm (one CTA per 64 rows below, one thread per row) solves
// L_ik L_kk^T = A_ik by substitution and updates y_i -= L_ik y_k; k_gchol_syrk (one CTA per tile of the trailing lower
// triangle) subtracts L_ik L_jk^T.  Back substitution walks the panels in reverse (k_gchol_back): x_k = L_kk^-T y_k,
// then every earlier entry y_j -= L_kj^T x_k.  Products use explicit fma(): this factorisation is bound by the 1e-6
// chi2 tolerance, not by bit parity with the oracle's scalar Cholesky.  prm.ok = 0 when a pivot is not positive.
constexpr int kGNB = 64;
constexpr int kGSub = 16;
constexpr size_t kGcholSmem = sizeof(double) * 2 * kGNB * (kGNB + 4);

__global__ void __launch_bounds__(256) k_gchol_scan(BaBuf B, uint8_t* __restrict__ nz, int ldt, int force) {
  __shared__ int s_any;
  const BaParams& prm = *B.prm;
  if (prm.done && !force) return;
  const int n = prm.vb_elim ? prm.cn : prm.np, ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti || ti * kGNB >= n) return;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  int any = 0;
  for (int e = threadIdx.x; e < kGNB * kGNB; e += 256) {
    const int i = ti * kGNB + e / kGNB, j = tj * kGNB + e % kGNB;
    if (i < n && j <= i && B.S[(size_t)i * n + j] != 0.0) any = 1;
  }
  if (any) s_any = 1;
  __syncthreads();
  if (threadIdx.x == 0) nz[ti * ldt + tj] = (uint8_t)(s_any || ti == tj);
}

__global__ void __launch_bounds__(256) k_gchol_diag(BaBuf B, double* __restrict__ yv, int k0, int force) {
  extern __shared__ double s_gchol[];
  double (*A)[kGNB + 1] = reinterpret_cast<double (*)[kGNB + 1]>(s_gchol);
  __shared__ double sy[kGNB], rd[kGNB];  // rd = 1 / L_jj: fp64 division is a long instruction sequence, multiply instead
  __shared__ int s_good;
  BaParams& prm = *B.prm;
  if (prm.done && !force) return;
  const int n = prm.vb_elim ? prm.cn : prm.np;  // the compact system when the V / Bias chain is eliminated first
  if (k0 >= n) return;
  const int nb = min(kGNB, n - k0), t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (k0 == 0) {  // start of a solve: y = bschur
    for (int i = t; i < n; i += 256) yv[i] = B.bs[i];
  }
  for (int e = t; e < kGNB * kGNB; e += 256) {
    const int i = e / kGNB, j = e % kGNB;
    A[i][j] = (i < nb && j <= i) ? B.S[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
  }
  if (t == 0) s_good = 1;
  __syncthreads();
  if (t < kGNB) sy[t] = t < nb ? yv[k0 + t] : 0.0;
  for (int c0 = 0; c0 < kGNB; c0 += kGSub) {
    // (1) the 16 x 16 diagonal sub-block, by warp 0: lane i < 16 keeps row i in registers, a finished column travels by
    // shuffle; one rsqrt per pivot (L_jj = d rs, 1 / L_jj = rs), no shared-memory round trip inside the 16 dependent steps
    if (warp == 0) {
      double row[kGSub];
#pragma unroll
      for (int c = 0; c < kGSub; ++c) row[c] = lane < kGSub ? A[c0 + lane][c0 + c] : 0.0;
#pragma unroll
      for (int j = 0; j < kGSub; ++j) {
        const double d = __shfl_sync(0xffffffffu, row[j], j);
        if (lane == 0 && !(d > 0)) s_good = 0;
        const double rs = rsqrt(d);
        const double lij = lane == j ? d * rs : (lane > j ? row[j] * rs : 0.0);
        row[j] = lij;
        if (lane == j) rd[c0 + j] = rs;
#pragma unroll
        for (int c = 1; c < kGSub; ++c) {  // constant bounds: both loops unroll and row[] stays in registers
          if (c <= j) continue;
          const double lc = __shfl_sync(0xffffffffu, lij, c);
          if (lane >= c) row[c] = fma(-lij, lc, row[c]);
        }
      }
      if (lane < kGSub) {
#pragma unroll
        for (int c = 0; c < kGSub; ++c)
          if (c <= lane) A[c0 + lane][c0 + c] = row[c];
      }
    }
    __syncthreads();
    // (2) rows below the sub-block: X L^T = A by substitution, one thread per row
    const int r0 = c0 + kGSub;
    if (t < kGNB - r0) {
      const int i = r0 + t;
      for (int j = 0; j < kGSub; ++j) {
        double acc = A[i][c0 + j];
        for (int m = 0; m < j; ++m) acc = fma(-A[i][c0 + m], A[c0 + j][c0 + m], acc);
        A[i][c0 + j] = acc * rd[c0 + j];
      }
    }
    __syncthreads();
    // (3) rank-16 update of the trailing lower triangle
    const int m = kGNB - r0;
    for (int e = t; e < m * m; e += 256) {
      const int i = r0 + e / m, k = r0 + e % m;
      if (k > i) continue;
      double acc = A[i][k];
#pragma unroll
      for (int q = 0; q < kGSub; ++q) acc = fma(-A[i][c0 + q], A[k][c0 + q], acc);
      A[i][k] = acc;
    }
    __syncthreads();
  }
  // forward substitution of the panel's right-hand side by warp 0 (lane holds rows lane and lane + 32)
  if (warp == 0) {
    double y0 = sy[lane], y1 = sy[lane + 32];
    for (int j = 0; j < kGNB; ++j) {
      double yj = __shfl_sync(0xffffffffu, j < 32 ? y0 : y1, j & 31) * rd[j];
      if (lane == j) y0 = yj;
      if (lane + 32 == j) y1 = yj;
      if (lane > j) y0 = fma(-A[lane][j], yj, y0);
      if (lane + 32 > j) y1 = fma(-A[lane + 32][j], yj, y1);
    }
    if (lane < nb) yv[k0 + lane] = y0;
    if (lane + 32 < nb) yv[k0 + lane + 32] = y1;
  }
  for (int e = t; e < nb * nb; e += 256) {
    const int i = e / nb, j = e % nb;
    if (j <= i) B.S[(size_t)(k0 + i) * n + k0 + j] = A[i][j];
  }
  __syncthreads();
  if (t == 0) {
    if (k0 == 0 && !prm.vb_elim) prm.ok = s_good;  // (eliminated form: k_vb_prepare set it, k_vb_factor may have cleared it)
    else if (!s_good) prm.ok = 0;
  }
}

__global__ void __launch_bounds__(128) k_gchol_trsm(BaBuf B, const uint8_t* __restrict__ nz, int ldt, double* __restrict__ yv,
                                                    int k0, int force) {
  extern __shared__ double s_gchol[];
  double (*As)[kGNB + 1] = reinterpret_cast<double (*)[kGNB + 1]>(s_gchol);
  double (*Ls)[kGNB + 1] = reinterpret_cast<double (*)[kGNB + 1]>(s_gchol + kGNB * (kGNB + 1));
  __shared__ double sy[kGNB], rd[kGNB];
  const BaParams& prm = *B.prm;
  if (prm.done && !force) return;
  const int n = prm.vb_elim ? prm.cn : prm.np;  // the compact system when the V / Bias chain is eliminated first
  if (k0 >= n) return;
  const int nb = min(kGNB, n - k0), t = threadIdx.x, kt = k0 / kGNB;
  const int i0 = k0 + nb + kGNB * blockIdx.x;
  if (i0 >= n) return;
  if (!nz[(kt + 1 + blockIdx.x) * ldt + kt]) return;  // a tile of zeros stays zero and leaves y alone
  const int nr = min(kGNB, n - i0);
  for (int e = t; e < kGNB * kGNB; e += 128) {
    const int i = e / kGNB, j = e % kGNB;
    As[i][j] = (i < nr && j < nb) ? B.S[(size_t)(i0 + i) * n + k0 + j] : 0.0;
    Ls[i][j] = (i < nb && j <= i) ? B.S[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
  }
  if (t < kGNB) {
    sy[t] = t < nb ? yv[k0 + t] : 0.0;
    rd[t] = t < nb ? 1.0 / B.S[(size_t)(k0 + t) * n + k0 + t] : 1.0;
  }
  __syncthreads();
  // Blocked substitution over four 16-column blocks: the part of a block that depends on the finished columns to its left is a
  // small matrix product spread over the whole CTA (8 outputs per thread); only the 16 x 16 triangular solve of the block is
  // sequential per row (136 instead of 2016 dependent multiply-adds per row).  Padding rows / columns are zeros / identity.
  for (int cb = 0; cb < kGNB; cb += kGSub) {
    if (cb > 0) {
      const int r = t & 63, cg = cb + (t >> 6) * 8;
      double acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.0;
      for (int m = 0; m < cb; ++m) {
        const double a = As[r][m];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fma(-a, Ls[cg + q][m], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) As[r][cg + q] += acc[q];
      __syncthreads();
    }
    if (t < kGNB) {
#pragma unroll 4
      for (int j = cb; j < cb + kGSub; ++j) {
        double acc = As[t][j];
        for (int m = cb; m < j; ++m) acc = fma(-As[t][m], Ls[j][m], acc);
        As[t][j] = acc * rd[j];
      }
    }
    __syncthreads();
  }
  if (t < nr) {
    double dy = 0;
    for (int j = 0; j < nb; ++j) dy = fma(As[t][j], sy[j], dy);
    yv[i0 + t] -= dy;
  }
  __syncthreads();
  for (int e = t; e < kGNB * kGNB; e += 128) {
    const int i = e / kGNB, j = e % kGNB;
    if (i < nr && j < nb) B.S[(size_t)(i0 + i) * n + k0 + j] = As[i][j];
  }
}

__global__ void __launch_bounds__(256) k_gchol_syrk(BaBuf B, uint8_t* __restrict__ nz, int ldt, int k0, int force) {
  extern __shared__ double s_gchol[];
  double (*At)[kGNB + 4] = reinterpret_cast<double (*)[kGNB + 4]>(s_gchol);                          // [m][row]
  double (*Bt)[kGNB + 4] = reinterpret_cast<double (*)[kGNB + 4]>(s_gchol + kGNB * (kGNB + 4));      // [m][col]
  const BaParams& prm = *B.prm;
  if (prm.done && !force) return;
  if (blockIdx.x > blockIdx.y) return;  // lower triangle of tiles only
  const int n = prm.vb_elim ? prm.cn : prm.np;
  if (k0 >= n) return;
  const int nb = min(kGNB, n - k0), t = threadIdx.x, kt = k0 / kGNB;
  const int ti = kt + 1 + blockIdx.y, tj = kt + 1 + blockIdx.x;
  const int i0 = k0 + nb + kGNB * blockIdx.y, j0 = k0 + nb + kGNB * blockIdx.x;
  if (i0 >= n || j0 >= n) return;
  if (!nz[ti * ldt + kt] || !nz[tj * ldt + kt]) return;
  if (t == 0) nz[ti * ldt + tj] = 1;  // fill-in
  for (int e = t; e < kGNB * kGNB; e += 256) {
    const int r = e / kGNB, m = e % kGNB;
    At[m][r] = (i0 + r < n && m < nb) ? B.S[(size_t)(i0 + r) * n + k0 + m] : 0.0;
    Bt[m][r] = (j0 + r < n && m < nb) ? B.S[(size_t)(j0 + r) * n + k0 + m] : 0.0;
  }
  __syncthreads();
  // FP64 tensor cores: D(8x8) += A(8x4) B(4x8) with mma.sync.m8n8k4.f64 (SASS: DMMA).  The 64 x 64 tile is split over the 8
  // warps as 16 rows x 32 columns each = 2 x 4 accumulator blocks; per k-step of 4 a lane loads two A and four B
  // fragment entries.  Fragment layout (PTX ISA, m8n8k4 .f64): a = A[lane / 4][lane % 4], b = B[lane % 4][lane / 4],
  // c/d = C[lane / 4][2 (lane % 4) + {0, 1}].  At / Bt are k-major with a row pitch of 68 doubles: the 16 lanes of a
  // half-warp (k = 0..3, rows 0..3) hit 16 distinct 8-byte banks.
  const int warp = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
  const int r0 = 16 * (warp >> 1), c0 = 32 * (warp & 1);
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int kk = 0; kk < kGNB; kk += 4) {  // rows of At / Bt beyond nb are zero
    double af[2], bf[4];
#pragma unroll
    for (int a = 0; a < 2; ++a) af[a] = At[kk + q][r0 + 8 * a + g];
#pragma unroll
    for (int b = 0; b < 4; ++b) bf[b] = Bt[kk + q][c0 + 8 * b + g];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(acc[a][b][0]), "+d"(acc[a][b][1])
                     : "d"(af[a]), "d"(bf[b]));
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int i = i0 + r0 + 8 * a + g;
    if (i >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + c0 + 8 * b + 2 * q + e;
        if (j < n && j <= i) B.S[(size_t)i * n + j] -= acc[a][b][e];
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_gchol_back(BaBuf B, const uint8_t* __restrict__ nz, int ldt, double* __restrict__ yv,
                                                    int k0, int force) {
  extern __shared__ double s_gchol[];
  double (*Ls)[kGNB + 1] = reinterpret_cast<double (*)[kGNB + 1]>(s_gchol);
  __shared__ double sx[kGNB], rdb[kGNB];  // rdb = 1 / L_jj, taken off the 64-step dependent chain below
  const BaParams& prm = *B.prm;
  if (prm.done && !force) return;
  const int n = prm.vb_elim ? prm.cn : prm.np;  // the compact system when the V / Bias chain is eliminated first
  if (k0 >= n) return;
  const int nb = min(kGNB, n - k0), t = threadIdx.x, lane = t & 31, kt = k0 / kGNB;
  // every block needs x_k; blocks > 0 whose four column tiles are all empty have nothing to update
  if (blockIdx.x > 0) {
    const int jt0 = (blockIdx.x - 1) * 4;
    bool any = false;
    for (int q = 0; q < 4; ++q) any |= (jt0 + q < kt) && nz[kt * ldt + jt0 + q];
    if (!any) return;
  }
  for (int e = t; e < kGNB * kGNB; e += 256) {
    const int i = e / kGNB, j = e % kGNB;
    Ls[i][j] = (i < nb && j <= i) ? B.S[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
  }
  __syncthreads();
  if (t < kGNB) rdb[t] = 1.0 / Ls[t][t];
  __syncthreads();
  if (t < 32) {  // x_k = L_kk^-T y_k by back substitution (lane holds rows lane and lane + 32)
    double x0 = lane < nb ? yv[k0 + lane] : 0.0, x1 = lane + 32 < nb ? yv[k0 + lane + 32] : 0.0;
    for (int j = kGNB - 1; j >= 0; --j) {
      const double xj = __shfl_sync(0xffffffffu, j < 32 ? x0 : x1, j & 31) * rdb[j];
      if (lane == j) x0 = xj;
      if (lane + 32 == j) x1 = xj;
      if (lane < j) x0 = fma(-Ls[j][lane], xj, x0);
      if (lane + 32 < j) x1 = fma(-Ls[j][lane + 32], xj, x1);
    }
    sx[lane] = x0;
    sx[lane + 32] = x1;
    if (blockIdx.x == 0) {
      if (lane < nb) B.x[k0 + lane] = x0;
      if (lane + 32 < nb) B.x[k0 + lane + 32] = x1;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) return;
  const int j = (blockIdx.x - 1) * 256 + t;
  if (j >= k0 || !nz[kt * ldt + j / kGNB]) return;
  double acc = yv[j];
#pragma unroll 8
  for (int r = 0; r < nb; ++r) acc = fma(-B.S[(size_t)(k0 + r) * n + j], sx[r], acc);
  yv[j] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Velocity / bias elimination before the dense factorisation.  After the landmark Schur complement the reduced camera
// system couples the velocity / bias vertices y_m = (V, Bias) of keyframe m (9 dimensions) only to y_{m-1}, y_{m+1} (inertial
// + bias-walk edges between consecutive keyframes) and to a handful of PR vertices, while every reprojection edge and the
// whole landmark fill-in live in the PR block.  S = [Spp Spy; Syp Syy] with Syy block-tridiagonal: eliminating y first,
//     S' = Spp - Spy Syy^-1 Syp,  b' = bp - Spy Syy^-1 by,   then   y = Syy^-1 (by - Syp xp),
// leaves a dense system of cn = 6 nfree + border dimensions instead of 15 nfree: 15.6 x fewer factorisation flops and 2.5 x
// fewer panels at BASELINE configs[4] size (5986 -> 2395 dimensions).  The result is the same solve in a different
// elimination order (the reference's sparse LDLT picks yet another one through AMD): inside the 1e-6 chi2 budget.
//   k_vb_factor  one warp: block Cholesky of the tridiagonal chain, L_m L_m^T = D_m - F_{m-1} F_{m-1}^T, F_m = E_m L_m^-T
//   k_vb_solve   thread per right-hand side (the cn columns of Syp and by): forward / backward sweep -> Z
//   k_vb_reduce  S'(r, c) and b'(r) for the lower triangle; PR rows touch 3 chain blocks, border rows all of them
//   k_vb_back    y = Z(:, cn) - Z(:, 0:cn) xp, and the scatter of (xp, y) back into the interleaved order of B.x
constexpr int kVB = 9;

constexpr int kVbLS = 90;  // stride of a chain block in vbL: the 9 x 9 factor, then the reciprocals of its diagonal

__global__ void __launch_bounds__(32) k_vb_factor(BaBuf B, int force) {
  __shared__ double sL[kVB][kVB + 1], sF[kVB][kVB + 1], sD[kVB][kVB + 1], sR[kVB];
  BaParams& prm = *B.prm;
  if ((prm.done && !force) || !prm.vb_elim) return;
  const int lane = threadIdx.x, np = prm.np, Ky = prm.Ky;
  bool good = true;
  // the chain is sequential by definition: what can be hidden is the latency of the 2 x 81 scattered loads of a block,
  // fetched one step ahead into registers (entries e = lane, lane + 32, lane + 64 of D_m and of E_{m-1}); what can be
  // shortened is the dependent arithmetic of a step: no fp64 division or square root on the path (rsqrt of the pivot gives
  // both L_jj = d rs and 1 / L_jj = rs; the reciprocals are kept for the substitutions here and in k_vb_solve)
  double nd[3] = {0, 0, 0}, ne[3] = {0, 0, 0};
  auto fetch = [&](int m) {
    const int* ym = B.ymap + kVB * m;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = lane + 32 * q;
      if (e < kVB * kVB) {
        const int i = e / kVB, j = e % kVB;
        nd[q] = B.S[(size_t)ym[i] * np + ym[j]];
        ne[q] = m > 0 ? B.S[(size_t)ym[i] * np + B.ymap[kVB * (m - 1) + j]] : 0.0;
      }
    }
  };
  if (Ky > 0) fetch(0);
  for (int m = 0; m < Ky; ++m) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = lane + 32 * q;
      if (e < kVB * kVB) {
        sD[e / kVB][e % kVB] = nd[q];
        sF[e / kVB][e % kVB] = ne[q];
      }
    }
    if (m + 1 < Ky) fetch(m + 1);
    __syncwarp();
    if (m > 0) {
      // F = E L_{m-1}^-T: row i of F solves F(i, :) L^T = E(i, :) by forward substitution over the columns (lane i < 9)
      if (lane < kVB) {
        double f[kVB];
#pragma unroll
        for (int j = 0; j < kVB; ++j) {
          double v = sF[lane][j];
#pragma unroll
          for (int k = 0; k < j; ++k) v -= f[k] * sL[j][k];
          f[j] = v * sR[j];
        }
#pragma unroll
        for (int j = 0; j < kVB; ++j) sF[lane][j] = f[j];
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int e = lane + 32 * q;
        if (e < kVB * kVB) {
          const int i = e / kVB, j = e % kVB;
          double v = 0;
#pragma unroll
          for (int k = 0; k < kVB; ++k) v += sF[i][k] * sF[j][k];
          sD[i][j] -= v;
          B.vbF[(size_t)(m - 1) * 81 + e] = sF[i][j];
        }
      }
      __syncwarp();
    }
    // L_m = chol(D), right-looking on a register-resident row per lane (lane i < 9 owns row i of D): per column one rsqrt and
    // one broadcast of the finished column through shared memory
    {
      double row[kVB];
#pragma unroll
      for (int c = 0; c < kVB; ++c) row[c] = lane < kVB ? sD[lane][c] : 0.0;
#pragma unroll
      for (int j = 0; j < kVB; ++j) {
        const double d = __shfl_sync(0xffffffffu, row[j], j);
        if (!(d > 0) || !isfinite(d)) good = false;
        const double rs = rsqrt(d);
        const double lij = lane == j ? d * rs : (lane > j ? row[j] * rs : 0.0);
        row[j] = lij;
        if (lane < kVB) sL[lane][j] = lij;
        if (lane == j) sR[j] = rs;
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < kVB; ++c)
          if (lane >= c && lane < kVB) row[c] -= lij * sL[c][j];
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = lane + 32 * q;
      if (e < kVB * kVB) B.vbL[(size_t)m * kVbLS + e] = sL[e / kVB][e % kVB];
    }
    if (lane < kVB) B.vbL[(size_t)m * kVbLS + 81 + lane] = sR[lane];
    __syncwarp();
  }
  if (lane == 0 && !good) prm.ok = 0;  // ok was set to 1 by k_vb_prepare before this kernel
}

// marks the start of a solve for the eliminated form: prm.ok = 1 (k_gchol_diag's first panel would otherwise overwrite a
// failed chain factorisation)
__global__ void k_vb_prepare(BaBuf B, int force) {
  BaParams& prm = *B.prm;
  if ((prm.done && !force) || !prm.vb_elim) return;
  prm.ok = 1;
}

// Nine lanes per right-hand side (column c of Syp, or by), lane i = row i of the 9-dimensional chain block, three columns per
// warp: a step is a 9 x 9 matrix-vector product and a 9 x 9 triangular solve whose operands a single thread had to pull
// through 117 shared-memory loads with nothing to hide their latency behind (one warp per SM: 1.9 us per step).  Here a lane
// keeps ITS row of F and L in registers (prefetched one step ahead straight from L2 / L1: every warp reads the same factors),
// the vector travels by shuffles, and the dependent chain of a step is 9 multiply-adds + 9 (multiply, shuffle, multiply-add).
// No division (reciprocal diagonals from k_vb_factor), no load of a structural zero: column c of Syp is non-zero only in the
// chain blocks [ylo[c], yhi[c]), the forward sweep starts at the warp's first such block with z = 0.
constexpr int kVbCols = 3, kVbSolveWarps = 4, kVbSolveThreads = 32 * kVbSolveWarps;
__global__ void __launch_bounds__(kVbSolveThreads) k_vb_solve(BaBuf B, int force) {
  const BaParams& prm = *B.prm;
  if ((prm.done && !force) || !prm.vb_elim) return;
  const int cn = prm.cn, np = prm.np, Ky = prm.Ky, ld = cn + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / kVB, i = lane - kVB * g, base = kVB * g;  // lanes 27..31 idle (they shuffle along, never store)
  const int c = (blockIdx.x * kVbSolveWarps + warp) * kVbCols + g;
  const bool live = g < kVbCols && c <= cn, rhs = c == cn;
  const int pc = (live && !rhs) ? B.pmap[c] : 0;
  const int lo = !live ? Ky : (rhs ? 0 : B.ylo[c]), hi = rhs ? Ky : (live ? B.yhi[c] : 0);
  constexpr unsigned kFull = 0xffffffffu;
  const int m_start = __reduce_min_sync(kFull, lo);
  // ---- forward: z_m = L_m^-1 (r_m - F_{m-1} z_{m-1}) ----
  double Fi[kVB], Li[kVB], ri = 1.0, z = 0.0;
  int ymi = 0;
  auto fetch_fwd = [&](int m, double (&F)[kVB], double (&L)[kVB], double& r, int& ym) {
#pragma unroll
    for (int k = 0; k < kVB; ++k) {
      F[k] = m > 0 ? B.vbF[(size_t)(m - 1) * 81 + i * kVB + k] : 0.0;
      L[k] = B.vbL[(size_t)m * kVbLS + i * kVB + k];
    }
    r = B.vbL[(size_t)m * kVbLS + 81 + i];
    ym = B.ymap[kVB * m + i];
  };
  if (m_start < Ky) fetch_fwd(m_start, Fi, Li, ri, ymi);
  for (int m = m_start; m < Ky; ++m) {
    double nF[kVB], nL[kVB], nr = 1.0;
    int nym = 0;
    if (m + 1 < Ky) fetch_fwd(m + 1, nF, nL, nr, nym);
    double v = 0.0;
    if (live && m >= lo && m < hi) v = rhs ? B.bs[ymi] : __ldg(&B.S[(size_t)ymi * np + pc]);
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < kVB; ++k) a += Fi[k] * __shfl_sync(kFull, z, base + k);
    v -= a;
#pragma unroll
    for (int j = 0; j < kVB; ++j) {
      const double zj = __shfl_sync(kFull, v * ri, base + j);
      if (i == j) z = zj;
      else if (i > j) v -= Li[j] * zj;
    }
    if (live && m >= lo) B.Z[(size_t)(kVB * m + i) * ld + c] = z;
#pragma unroll
    for (int k = 0; k < kVB; ++k) {
      Fi[k] = nF[k];
      Li[k] = nL[k];
    }
    ri = nr;
    ymi = nym;
  }
  // ---- backward: w_m = L_m^-T (z_m - F_m^T w_{m+1}); blocks before lo hold z_m = 0 (the forward sweep never wrote them) ----
  double w = 0.0, nx = 0.0;
  auto fetch_bwd = [&](int m, double (&F)[kVB], double (&L)[kVB], double& r, double& zin) {
#pragma unroll
    for (int k = 0; k < kVB; ++k) {
      F[k] = m < Ky - 1 ? B.vbF[(size_t)m * 81 + k * kVB + i] : 0.0;   // column i of F_m
      L[k] = B.vbL[(size_t)m * kVbLS + k * kVB + i];                   // column i of L_m
    }
    r = B.vbL[(size_t)m * kVbLS + 81 + i];
    zin = (live && m >= lo) ? B.Z[(size_t)(kVB * m + i) * ld + c] : 0.0;
  };
  if (Ky > 0) fetch_bwd(Ky - 1, Fi, Li, ri, nx);
  for (int m = Ky - 1; m >= 0; --m) {
    double nF[kVB], nL[kVB], nr = 1.0, nz = 0.0;
    if (m > 0) fetch_bwd(m - 1, nF, nL, nr, nz);
    double v = nx, a = 0.0;
#pragma unroll
    for (int k = 0; k < kVB; ++k) a += Fi[k] * __shfl_sync(kFull, w, base + k);
    v -= a;
#pragma unroll
    for (int j = kVB - 1; j >= 0; --j) {
      const double wj = __shfl_sync(kFull, v * ri, base + j);
      if (i == j) w = wj;
      else if (i < j) v -= Li[j] * wj;
    }
    if (live) B.Z[(size_t)(kVB * m + i) * ld + c] = w;
#pragma unroll
    for (int k = 0; k < kVB; ++k) {
      Fi[k] = nF[k];
      Li[k] = nL[k];
    }
    ri = nr;
    nx = nz;
  }
}

// grid (ceil((cn + 1) / 128), cn): row r of the compact system, lower triangle (c <= r) and the rhs (c == cn)
__global__ void __launch_bounds__(128) k_vb_reduce(BaBuf B, int force) {
  __shared__ double s_row[128];
  const BaParams& prm = *B.prm;
  if ((prm.done && !force) || !prm.vb_elim) return;
  const int cn = prm.cn, np = prm.np, ld = cn + 1;
  const int r = blockIdx.y, c = blockIdx.x * 128 + threadIdx.x;
  if ((int)blockIdx.x * 128 > r && (int)blockIdx.x != (cn / 128)) return;  // tile entirely above the diagonal, no rhs in it
  const int pr = B.pmap[r];
  const bool is_rhs = c == cn, active = c <= r || is_rhs;
  double acc = 0;
  if (active && c <= cn) acc = is_rhs ? B.bs[pr] : B.S[(size_t)pr * np + B.pmap[c]];
  const int y0 = kVB * B.ylo[r], y1 = kVB * B.yhi[r];
  for (int base = y0; base < y1; base += 128) {
    const int n = min(128, y1 - base);
    __syncthreads();
    if ((int)threadIdx.x < n) s_row[threadIdx.x] = B.S[(size_t)pr * np + B.ymap[base + threadIdx.x]];
    __syncthreads();
    if (active && c <= cn)
      for (int k = 0; k < n; ++k) acc -= s_row[k] * B.Z[(size_t)(base + k) * ld + c];
  }
  if (active && c <= cn) {
    if (is_rhs) B.cbs[r] = acc;
    else B.cS[(size_t)r * cn + c] = acc;
  }
}

// one warp per chain row: y = Z(:, cn) - Z(:, 0:cn) xp; the last block scatters xp
__global__ void __launch_bounds__(256) k_vb_back(BaBuf B, int force) {
  const BaParams& prm = *B.prm;
  if ((prm.done && !force) || !prm.vb_elim) return;
  const int cn = prm.cn, ld = cn + 1, ny = kVB * prm.Ky;
  if (blockIdx.x == gridDim.x - 1) {
    for (int c = threadIdx.x; c < cn; c += 256) B.x[B.pmap[c]] = B.cx[c];
    return;
  }
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= ny) return;
  const double* Zr = B.Z + (size_t)row * ld;
  double a = 0;
  for (int c = lane; c < cn; c += 32) a += Zr[c] * B.cx[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) B.x[B.ymap[row]] = Zr[cn] - a;
}

__global__ void __launch_bounds__(256) k_gba_zero_h(BaBuf B, int into_other) {
  const BaParams& prm = *B.prm;
  if (prm.done) return;
  double2* H = reinterpret_cast<double2*>(B.H[into_other ? 1 - prm.cur : prm.cur]);
  const size_t n2 = ((size_t)prm.np * prm.np + 1) / 2;  // the buffers are allocated with an even element count
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
    H[i] = make_double2(0.0, 0.0);
  double* b = B.b[into_other ? 1 - prm.cur : prm.cur];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < prm.np; i += gridDim.x * blockDim.x) b[i] = 0;
}
// Global BA: hundreds of inertial factors.  Residuals / Jacobians / J^T rho' Omega of eight edges per CTA ...
constexpr int kGbaDenPerCta = 8;
__global__ void __launch_bounds__(256) k_gba_dense_eval(BaBuf B, int into_other) {
  const BaParams& prm = *B.prm;
  if (prm.done) return;
  const int m0 = blockIdx.x * kGbaDenPerCta;
  if (m0 >= prm.n_den) return;
  ba_dense_block(B.den + m0, min(kGbaDenPerCta, prm.n_den - m0), B.st, B.pre, ba_gravity(prm), B.wk + m0, B.off0, B.off1,
                 B.off2, prm.np, nullptr, nullptr, nullptr, true, &prm);
}
// ... then the accumulation, one colour per launch and one CTA per edge: edges of a colour share no keyframe, colours run
// one after the other, so every entry of H / b is summed in a fixed order without atomics.
__global__ void __launch_bounds__(128) k_gba_dense_accum(BaBuf B, int into_other, int color) {
  BaParams& prm = *B.prm;
  if (prm.done) return;
  const int m = blockIdx.x;
  if (m >= prm.n_den) return;
  const int set = into_other ? 1 - prm.cur : prm.cur;
  if (color == 0 && m == 0 && threadIdx.x == 0) {
    double tot = 0;
    for (int k = 0; k < prm.n_den; ++k) tot += B.wk[k].rho0;
    prm.chi_dense = tot;
  }
  if (B.den[m].color != color) return;
  ba_dense_add_edge(B.den[m], B.wk[m], B.off0, B.off1, B.off2, prm.np, B.H[set], B.b[set], 0, 128, prm.has_g ? prm.off_g : -1);
}

// One warp per point: xl = Dinv (bl - sum_a W_a^T xp), X += xl (apply != 0), landmark part of computeScale; the
// estimate before the update is kept in X_bak / st_bak (push()).  The extra last block applies the keyframe updates
// (NavState::IncSmall), refreshes the camera poses and computes the pose part of computeScale.
__global__ void __launch_bounds__(kBaWarps * 32) k_ba_backsub(BaBuf B, int apply, double lambda_arg, double* xl_out) {
  BaParams& prm = *B.prm;
  if (prm.done && apply) return;
  const int set = prm.cur;
  const double lambda = apply ? prm.lambda : lambda_arg;
  const bool ok = prm.ok != 0;
  if (blockIdx.x == gridDim.x - 1) {
    for (int k = threadIdx.x; k < prm.K; k += blockDim.x) {
      NavS s = ns_load(B.st[k]);
      if (apply) {
        B.st_bak[k] = B.st[k];
        if (ok && (B.off0[k] >= 0 || B.off1[k] >= 0 || B.off2[k] >= 0)) {
          if (B.off0[k] >= 0) ns_inc_pr(s, B.x + B.off0[k]);
          if (B.off1[k] >= 0) ns_inc_v(s, B.x + B.off1[k]);
          if (B.off2[k] >= 0) ns_inc_bias(s, B.x + B.off2[k]);
          ns_store(s, B.st[k]);
        }
      }
      B.cp[k] = cam_pose(prm.cam, s);
    }
    if (threadIdx.x == 0) {
      double s = 0;
      if (ok)
        for (int j = 0; j < prm.np; ++j) s += B.x[j] * (lambda * B.x[j] + B.bsys[j]);
      prm.pscale = s;
    }
    if (threadIdx.x == 32 && apply && (prm.has_scale || prm.has_g)) {  // push() + oplus of the two border vertices
      prm.sc_bak = prm.sc;
      for (int q = 0; q < 4; ++q) prm.qwI_bak[q] = prm.qwI[q];
      if (ok && prm.has_scale) prm.sc += B.x[prm.off_s];  // VertexScale::oplusImpl
      if (ok && prm.has_g) {  // VertexGThetaXYRwI::oplusImpl: RwI <- RwI Exp((dx, dy, 0))
        const Quat q = q_normalized(q_mul({prm.qwI[0], prm.qwI[1], prm.qwI[2], prm.qwI[3]},
                                          so3_exp_q({B.x[prm.off_g], B.x[prm.off_g + 1], 0.0})));
        prm.qwI[0] = q.w; prm.qwI[1] = q.x; prm.qwI[2] = q.y; prm.qwI[3] = q.z;
      }
    }
    return;
  }
  if ((int)blockIdx.x >= prm.n_pblk) return;
  const int p = blockIdx.x * kBaWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= prm.P) return;
  if (apply && lane < 3) B.X_bak[3 * (size_t)p + lane] = B.X[3 * (size_t)p + lane];
  if (!B.pt_active[set][p] || !ok) {
    if (lane == 0) {
      B.scale_part[p] = 0;
      if (xl_out) xl_out[3 * (size_t)p] = xl_out[3 * (size_t)p + 1] = xl_out[3 * (size_t)p + 2] = 0;
    }
    return;
  }
  const double* Wb = B.W[set];
  double c[3] = {0, 0, 0};
  for (int i = B.pt_ptr[p] + lane; i < B.pt_ptr[p + 1]; i += 32) {
    const int o = B.off0[B.es[i]];
    if (o < 0) continue;
    const double* Wi = Wb + 18 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int r = 0; r < 6; ++r) c[k] += Wi[3 * r + k] * B.x[o + r];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
  __syncwarp();  // the X_bak copy above read X before lane 0 updates it
  if (lane == 0) {
    if (prm.has_scale) {  // the point's block of the scale column times the scale increment
      const double xs = B.x[prm.off_s];
      const double* ws = B.wsp[set] + 3 * (size_t)p;
      c[0] += ws[0] * xs; c[1] += ws[1] * xs; c[2] += ws[2] * xs;
    }
    const double* bb = B.bl[set] + 3 * (size_t)p;
    const double* Di = B.Dinv + 9 * (size_t)p;
    const double cc[3] = {bb[0] - c[0], bb[1] - c[1], bb[2] - c[2]};
    double s = 0;
    for (int q = 0; q < 3; ++q) {
      const double xl = Di[3 * q] * cc[0] + Di[3 * q + 1] * cc[1] + Di[3 * q + 2] * cc[2];
      if (apply) B.X[3 * (size_t)p + q] += xl;
      if (xl_out) xl_out[3 * (size_t)p + q] = xl;
      s += xl * (lambda * xl + bb[q]);
    }
    B.scale_part[p] = s;
  }
}

// level / erase classification.  mode 0: Chi2LargeSetLevel (chi2 > rat * chi2_sig5[dim]); mode 1: the LBA gates
// (src/Optimizer.cc:597-633, 668-700): chi2 > 5.991 (x1.5 when close) / 7.815 or depth <= 0.
__global__ void k_ba_classify(CamK cam, const CamPose* __restrict__ cp, const double* __restrict__ X,
                              const int* __restrict__ es, const int* __restrict__ ep, const float* __restrict__ obs,
                              const uint8_t* __restrict__ flags, const double* __restrict__ chi2, int E, int mode, float rat,
                              int set_level, int remove_kernels, int use_close, uint8_t* __restrict__ lvl,
                              uint8_t* __restrict__ bad_out, const BaParams* __restrict__ prmq, int skip_if_stop = 0,
                              int skip_if_visual = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (E < 0) {  // graph form (capacity grid): sizes / camera / close-point gate from the device-resident problem header
    if (skip_if_visual && prmq->visual_only) return;
    E = prmq->E;
    cam = prmq->cam;
    use_close = prmq->visual_only ? 0 : 1;
  }
  if (i >= E) return;
  if (skip_if_stop && prmq->stop) return;  // aborted after the first stage: levels and kernels stay (bDoMore == false)
  const bool stereo = flags[i] & VIEO_EDGE_STEREO;
  bool bad;
  if (mode == 0) {
    const float th = rat * (stereo ? 7.815f : 5.991f);
    bad = chi2[i] > (double)th;
  } else {
    double e[3];
    const double depth = reproj_error(cam, cp[es[i]], ba_world_point(X, ep[i], prmq->sc), obs + 3 * (size_t)i, stereo, e);
    const float chi2Mono = 5.991f;
    if (stereo) bad = chi2[i] > 7.815 || !(depth > 0.);
    else if (!use_close) bad = chi2[i] > 5.991 || !(depth > 0.);  // visual LocalBundleAdjustment (src/Optimizer.cc:2198)
    else bad = chi2[i] > ((flags[i] & VIEO_EDGE_CLOSE) ? 1.5 * chi2Mono : (double)chi2Mono) || !(depth > 0.);
  }
  uint8_t l = lvl[i];
  if (set_level && bad) l |= 1;
  if (remove_kernels) l |= 2;
  lvl[i] = l;
  if (bad_out) bad_out[i] = bad;
}

}  // namespace vieo

using namespace vieo;

struct vieo_ba {
  int device = 0;
  cudaStream_t st = nullptr;
  int capK = 0, capP = 0, capE = 0, capM = 0, cap_pblk = 0, cap_free = 0;
  int K = 0, P = 0, E = 0, M = 0, np = 0, nfree = 0, n_den = 0, n_part = 0, n_pblk = 0, n_colors = 1;
  bool points_free = true, has_dup = false, visual_only = false;
  bool big = false;  // global-BA sized handle (dense multi-CTA Schur / Cholesky path, no trial graph)
  bool has_scale = false, has_g = false;  // border vertices of the current problem (global handles only)
  // asynchronous LocalBA (vieo_local_ba_prv_begin / _end): 0 idle, 1 nothing to do (no free keyframe), 2 aborted before
  // optimising, 3 enqueued, 4 ran synchronously inside begin (sharded handle / no optimize graph)
  int async_state = 0;
  bool defer_sync = false;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;  // device time of the last asynchronous LocalBA call (vieo_ba_last_ms)
  float last_ms = 0.f;
  bool upload_direct = false;  // set_problem copied from caller memory (staging too small): it must synchronise
  const volatile uint8_t* async_stop = nullptr;
  VieoBaProblem async_pb;      // shallow copy (sizes / flags) of the problem in flight
  std::vector<VieoNavState> async_states;
  std::vector<double> async_points;
  std::vector<uint8_t> sync_erase;
  std::vector<double> sync_chi2;
  VieoBaResult sync_res;
  int sync_rc = 0;
  double* d_yv = nullptr;
  uint8_t* d_nz = nullptr;  // tile structure map of the dense Cholesky
  bool vb_elim = false;     // V / Bias chain eliminated before the dense factorisation (global handles)
  int cn = 0, Ky = 0;
  int *d_pmap = nullptr, *d_ymap = nullptr, *d_ylo = nullptr, *d_yhi = nullptr;
  cudaStream_t st_aux = nullptr;  // the chain factorisation / solve runs beside the landmark Schur complement (single GPU)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int rank = 0, world = 1;
  vieo_allreduce_fn allreduce = nullptr;
  void* ar_ctx = nullptr;
  vieo_comm* comm = nullptr;  // library-owned NCCL communicator (vieo_ba_set_comm): the exchange without a host callback
  bool sharded() const { return world > 1 && (allreduce || comm); }
  CamK cam;
  Vec3 gw;
  double dm = 0, ds = 0;
  BaBuf B;                    // device pointers handed to the kernels by value
  BaParams* h_prm = nullptr;  // pinned host mirror of B.prm
  cudaGraphExec_t trial_graph = nullptr;
  // a whole optimize() as ONE launch: a WHILE conditional node whose body is the LM trial; k_ba_control keeps the loop
  // alive until the device-side state machine is done (no host round trip per trial)
  cudaGraphExec_t opt_graph = nullptr;
  // the whole LocalBA routine (everything vieo_local_ba_prv_begin enqueues after the problem upload) as ONE graph with
  // constant launch parameters: capacity grids, sizes / camera / iteration plan read from the device-resident header,
  // two WHILE nodes for the optimize() stages, the downloads as capacity-sized copies into the pinned staging buffer
  cudaGraphExec_t win_graph = nullptr;
  int win_nodes = 0;           // kernels of the graph outside the two WHILE bodies
  int plan_it[2] = {0, 0};     // set by vieo_local_ba_prv_begin before set_problem fills the header
  double plan_lambda0 = 0;
  size_t dl_off[4] = {0, 0, 0, 0};  // states | points | chi2 | erase candidates in h_stage of the call in flight
  cudaStream_t st_ctl = nullptr;  // side stream for the abort flag while the optimize graph runs
  // buffers that are not part of BaBuf
  double *d_sys = nullptr, *d_xl = nullptr, *d_ctl = nullptr;
  uint8_t *d_lvl = nullptr, *d_bad = nullptr, *d_flags = nullptr, *d_sfix = nullptr;
  int *d_es = nullptr, *d_ep = nullptr, *d_pt_ptr = nullptr, *d_off0 = nullptr, *d_off1 = nullptr, *d_off2 = nullptr,
      *d_prcol = nullptr, *d_free_state = nullptr, *d_free_off = nullptr, *d_ps_ptr = nullptr, *d_ps_edges = nullptr, *d_ecol = nullptr;
  float *d_obs = nullptr, *d_w = nullptr;
  VieoImuPreint* d_pre = nullptr;
  BaDense* d_den = nullptr;
  double* h_ctl = nullptr;  // pinned
  uint8_t* h_stage = nullptr;  // pinned staging for the problem upload / result download (no pageable copies)
  size_t stage_cap = 0, stage_used = 0;
  std::vector<int> off0, off1, off2;
  int launches = 0;
  size_t cap_np = 0;
  // sys = [bschur (cap_np) | b (cap_np) | S (np x np)]: the prefix that the sharded form all-reduces once per LM trial
  size_t sys_count() const { return 2 * cap_np + (size_t)np * np; }
};

namespace {

template <class T>
cudaError_t dalloc(T** p, size_t n) {
  return cudaMalloc((void**)p, sizeof(T) * std::max<size_t>(n, 1));
}

// SO3ex::exp -> normalised quaternion (w, x, y, z): host copy of so3_exp_q (so3.cuh) for the gravity-direction seed
void host_so3_exp_q(double wx, double wy, double wz, double q[4]) {
  const double theta = std::sqrt(wx * wx + wy * wy + wz * wz);
  double imag, real;
  if (theta < 1e-5) {
    const double t2 = theta * theta;
    imag = 0.5 - t2 / 48.;
    real = 1.0 - t2 / 8.;
  } else {
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  q[0] = real; q[1] = imag * wx; q[2] = imag * wy; q[3] = imag * wz;
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n > 0)
    for (int k = 0; k < 4; ++k) q[k] /= n;
}

bool host_inverse(const double* A, int n, double* Ai) {
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

#define BA_CK(call)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      vieo::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return VIEO_E_CUDA;                                                                \
    }                                                                                    \
  } while (0)

constexpr int kNodesPerTrial = 8;  // kernels of one LM trial (ba_enqueue_trial)
size_t schur_smem(int nfree) { return sizeof(double) * kBaWarps * 6 * (6 * (size_t)nfree + 1); }
size_t gba_schur_smem(int nfree) {  // row tile [6][6 nfree + 2]; the extra block's two reduction arrays need 2 T doubles
  const size_t tile = sizeof(double) * std::max<size_t>(6 * (6 * (size_t)nfree + 2), 2 * (size_t)kGbaSchurThreads);
  // + the producer / consumer ring behind the tile and its 2 x kSchurRing mbarriers
  const size_t R = (size_t)schur_ring_depth(nfree);
  return std::max(tile, schur_ring_offset(nfree) + R * sizeof(SchurSlot) + 2 * R * sizeof(uint64_t));
}

int ba_campose(vieo_ba* h) {
  k_ba_campose<<<(h->K + 127) / 128, 128, 0, h->st>>>(h->cam, h->B.st, h->K, h->B.cp);
  h->launches++;
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

int ba_allreduce(vieo_ba* h, double* buf, size_t n);

// The caller's abort flag as ONE decision for all ranks of a sharded handle (max over ranks through the sum all-reduce);
// a plain read on a single GPU.  Synchronises the stream when sharded.
int ba_agree_stop(vieo_ba* h, const volatile uint8_t* stop, bool* out) {
  const bool mine = stop && *stop;
  if (!h->sharded()) {
    *out = mine;
    return VIEO_OK;
  }
  h->h_ctl[8] = mine ? 1.0 : 0.0;
  BA_CK(cudaMemcpyAsync(h->d_ctl + 8, h->h_ctl + 8, 8, cudaMemcpyHostToDevice, h->st));
  int rc = ba_allreduce(h, h->d_ctl + 8, 1);
  if (rc) return rc;
  BA_CK(cudaMemcpyAsync(h->h_ctl + 8, h->d_ctl + 8, 8, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  *out = h->h_ctl[8] > 0.0;
  return VIEO_OK;
}

// stand-alone computeActiveErrors (+ every edge when all) and the robust chi2 into *d_out (used outside the LM loop)
int ba_errors(vieo_ba* h, int all, double* d_out) {
  {
    const int rc = ba_campose(h);
    if (rc) return rc;
  }
  if (h->E > 0) {
    k_ba_errors<<<h->n_part, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_w, h->d_flags,
                                              h->d_lvl, h->d_sfix, h->points_free ? 1 : 0, h->E, all, h->dm, h->ds,
                                              h->B.chi2, h->B.partial, h->B.prm);
    h->launches++;
  }
  k_ba_dense_errors<<<1, 128, 0, h->st>>>(h->d_den, h->n_den, h->B.st, h->d_pre, h->B.prm, h->B.wk, h->B.partial,
                                          h->E > 0 ? h->n_part : 0, d_out, nullptr, all == 2 ? -1 : 0, nullptr);
  h->launches++;
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

int ba_allreduce(vieo_ba* h, double* buf, size_t n) {
  if (!h->sharded()) return VIEO_OK;
  if (h->comm) return vieo::comm_allreduce_f64(h->comm, buf, n, h->st);
  if (h->allreduce(h->ar_ctx, buf, n, (void*)h->st)) {
    vieo::set_error("allreduce callback failed");
    return VIEO_E_CUDA;
  }
  return VIEO_OK;
}

// computeActiveErrors + buildSystem at the current estimate; into_other: into the set that is not the current one
int ba_enqueue_linearize(vieo_ba* h, int into_other, bool at_capacity) {
  const int pblk = at_capacity ? h->cap_pblk : h->n_pblk, nf = at_capacity ? h->cap_free : h->nfree;
  if (h->big) {  // one block cannot zero np^2 entries or walk hundreds of inertial factors: dedicated kernels
    k_gba_zero_h<<<1184, 256, 0, h->st>>>(h->B, into_other);
    if (h->has_scale) k_ba_linearize_t<true><<<pblk + 1, kBaWarps * 32, 0, h->st>>>(h->B, into_other);
    else k_ba_linearize_t<false><<<pblk + 1, kBaWarps * 32, 0, h->st>>>(h->B, into_other);
    h->launches += 2;
    if (h->n_den > 0) {
      k_gba_dense_eval<<<(h->n_den + kGbaDenPerCta - 1) / kGbaDenPerCta, 256, 0, h->st>>>(h->B, into_other);
      for (int c = 0; c < h->n_colors; ++c) k_gba_dense_accum<<<h->n_den, 128, 0, h->st>>>(h->B, into_other, c);
      h->launches += 1 + h->n_colors;
    }
    k_ba_pose_reduce<<<std::max(nf, 1), 256, 0, h->st>>>(h->B, into_other);
    h->launches++;
    return VIEO_OK;
  }
  k_ba_linearize_t<false><<<pblk + 1, kBaWarps * 32, 0, h->st>>>(h->B, into_other);
  k_ba_pose_reduce<<<std::max(nf, 1), 256, 0, h->st>>>(h->B, into_other);
  h->launches += 2;
  return VIEO_OK;
}

// (H + lambda I) x = b through the Schur complement on the current set.  Trial form (force == 0): lambda from the
// device state, estimate updated.  force == 1: explicit lambda, nothing applied (vieo_ba_debug_step).
int ba_enqueue_solve(vieo_ba* h, bool at_capacity, int force, double lambda, double* xl_out) {
  const int pblk = at_capacity ? h->cap_pblk : h->n_pblk, nf = at_capacity ? h->cap_free : h->nfree;
  const size_t P = at_capacity ? (size_t)h->capP : (size_t)h->P;
  if (h->big) {
    const int n = h->np;
    const size_t nn = std::max<size_t>(std::max<size_t>((size_t)n * n, (size_t)n), std::max<size_t>(P, 1));
    k_ba_prep_solve<<<(unsigned)((nn + 255) / 256), 256, 0, h->st>>>(h->B, lambda, force);
    // The chain blocks (V / Bias rows of S) come from the inertial edges and the damping only — the landmark Schur
    // complement never touches them — so on one GPU their factorisation and the Z solve run BESIDE k_gba_schur on a
    // second stream; sharded, the inertial part lives on rank 0 until the all-reduce, so the chain waits for it.
    const bool vb_side = h->vb_elim && !h->sharded();
    if (vb_side) {
      BA_CK(cudaEventRecord(h->ev_fork, h->st));
      BA_CK(cudaStreamWaitEvent(h->st_aux, h->ev_fork, 0));
      k_vb_prepare<<<1, 1, 0, h->st_aux>>>(h->B, force);
      k_vb_factor<<<1, 32, 0, h->st_aux>>>(h->B, force);
      k_vb_solve<<<(h->cn + 1 + kVbCols * kVbSolveWarps - 1) / (kVbCols * kVbSolveWarps), kVbSolveThreads, 0, h->st_aux>>>(h->B, force);
      BA_CK(cudaEventRecord(h->ev_join, h->st_aux));
      h->launches += 3;
    }
    k_gba_schur<<<nf + 1, kGbaSchurThreads, gba_schur_smem(h->nfree), h->st>>>(h->B, force);
    h->launches += 2;
    int rc = ba_allreduce(h, h->d_sys, h->sys_count());
    if (rc) return rc;
    if (vb_side) BA_CK(cudaStreamWaitEvent(h->st, h->ev_join, 0));
    // the dense factorisation runs on the reduced camera system itself, or — V / Bias chain eliminated first — on the
    // compact PR system (BaBuf copy whose S / bschur / x point at the compact buffers)
    BaBuf Bc = h->B;
    int nc = n;
    if (h->vb_elim) {
      nc = h->cn;
      Bc.S = h->B.cS; Bc.bs = h->B.cbs; Bc.x = h->B.cx;
      if (!vb_side) {
        k_vb_prepare<<<1, 1, 0, h->st>>>(h->B, force);
        k_vb_factor<<<1, 32, 0, h->st>>>(h->B, force);
        k_vb_solve<<<(nc + 1 + kVbCols * kVbSolveWarps - 1) / (kVbCols * kVbSolveWarps), kVbSolveThreads, 0, h->st>>>(h->B, force);
        h->launches += 3;
      }
      k_vb_reduce<<<dim3((nc + 1 + 127) / 128, nc), 128, 0, h->st>>>(h->B, force);
      h->launches++;
    }
    const int T = (nc + kGNB - 1) / kGNB;
    k_gchol_scan<<<dim3(T, T), 256, 0, h->st>>>(Bc, h->d_nz, T, force);
    h->launches++;
    for (int k0 = 0; k0 < nc; k0 += kGNB) {
      const int k1 = std::min(k0 + kGNB, nc), tiles = (nc - k1 + kGNB - 1) / kGNB;
      k_gchol_diag<<<1, 256, kGcholSmem, h->st>>>(Bc, h->d_yv, k0, force);
      h->launches++;
      if (tiles > 0) {
        k_gchol_trsm<<<tiles, 128, kGcholSmem, h->st>>>(Bc, h->d_nz, T, h->d_yv, k0, force);
        k_gchol_syrk<<<dim3(tiles, tiles), 256, kGcholSmem, h->st>>>(Bc, h->d_nz, T, k0, force);
        h->launches += 2;
      }
    }
    for (int k0 = ((nc - 1) / kGNB) * kGNB; k0 >= 0; k0 -= kGNB) {
      k_gchol_back<<<1 + (k0 + 255) / 256, 256, kGcholSmem, h->st>>>(Bc, h->d_nz, T, h->d_yv, k0, force);
      h->launches++;
    }
    if (h->vb_elim) {
      k_vb_back<<<(9 * h->Ky + 7) / 8 + 1, 256, 0, h->st>>>(h->B, force);
      h->launches++;
    }
    k_ba_backsub<<<pblk + 1, kBaWarps * 32, 0, h->st>>>(h->B, force ? 0 : 1, lambda, xl_out);
    h->launches++;
    return VIEO_OK;
  }
  const size_t npc = at_capacity ? 15 * (size_t)std::min(h->capK, kSchurMaxFree + 8) : (size_t)h->np;
  const size_t nn = std::max<size_t>(std::max<size_t>(npc * npc, npc), std::max<size_t>(P, 1));
  k_ba_prep_solve<<<(unsigned)((nn + 255) / 256), 256, 0, h->st>>>(h->B, lambda, force);
  k_ba_schur<<<dim3(std::max(nf, 1), kSchurSplit), kBaWarps * 32, schur_smem(at_capacity ? h->cap_free : h->nfree), h->st>>>(h->B, force);
  k_ba_schur_reduce<<<std::max(nf, 1), 256, 0, h->st>>>(h->B, force);
  h->launches += 3;
  if (!at_capacity) {
    int rc = ba_allreduce(h, h->d_sys, h->sys_count());
    if (rc) return rc;
  }
  k_ba_chol<<<1, kCholThreads, kCholSmemBytes, h->st>>>(h->B, force);
  k_ba_backsub<<<pblk + 1, kBaWarps * 32, 0, h->st>>>(h->B, force ? 0 : 1, lambda, xl_out);
  h->launches += 2;
  return VIEO_OK;
}

// one LM trial: solve + update + linearisation at the trial estimate + bookkeeping
int ba_enqueue_trial(vieo_ba* h, bool at_capacity, cudaGraphConditionalHandle cond = 0) {
  int rc;
  if ((rc = ba_enqueue_solve(h, at_capacity, 0, 0.0, nullptr))) return rc;
  if ((rc = ba_enqueue_linearize(h, 1, at_capacity))) return rc;
  if (!at_capacity && (rc = ba_allreduce(h, h->B.prm->pair, 3))) return rc;
  k_ba_control<<<1, 256, 0, h->st>>>(h->B, 1, cond);
  h->launches++;
  return VIEO_OK;
}

int ba_read_prm(vieo_ba* h) {
  BA_CK(cudaMemcpyAsync(h->h_prm, h->B.prm, sizeof(BaParams), cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  return VIEO_OK;
}

void ba_free(vieo_ba* h) {
  BaBuf& B = h->B;
  void* ptrs[] = {B.prm, B.st, B.st_bak, B.cp, B.X, B.X_bak, B.chi2, B.A, B.Dinv, B.db, B.x, B.partial, B.scale_part, B.part,
                  B.W[0], B.W[1], B.Hll[0], B.Hll[1], B.bl[0], B.bl[1], B.H[0], B.H[1], B.b[0], B.b[1], B.pt_active[0],
                  B.pt_active[1], B.wk, B.wsp[0], B.wsp[1], B.up, B.sred, B.As, B.spart, B.cS, B.cbs, B.cx, B.vbL, B.vbF, B.Z,
                  h->d_pmap, h->d_ymap, h->d_ylo, h->d_yhi, h->d_nz, h->d_yv, h->d_sys, h->d_xl, h->d_ctl, h->d_lvl, h->d_bad, h->d_flags, h->d_sfix, h->d_es,
                  h->d_ep, h->d_pt_ptr, h->d_off0, h->d_off1, h->d_off2, h->d_prcol, h->d_free_state, h->d_free_off,
                  h->d_ps_ptr, h->d_ps_edges, h->d_obs, h->d_w, h->d_pre, h->d_den, h->d_ecol};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->st_aux) cudaStreamDestroy(h->st_aux);
  if (h->ev_begin) cudaEventDestroy(h->ev_begin);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  if (h->trial_graph) cudaGraphExecDestroy(h->trial_graph);
  if (h->opt_graph) cudaGraphExecDestroy(h->opt_graph);
  if (h->win_graph) cudaGraphExecDestroy(h->win_graph);
  if (h->st_ctl) cudaStreamDestroy(h->st_ctl);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->h_prm) cudaFreeHost(h->h_prm);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->st) cudaStreamDestroy(h->st);
}


// download layout of a LocalBA call in the pinned staging buffer (16-byte aligned spans)
void lba_dl_layout(size_t K, size_t P, size_t E, size_t off[4], size_t* total) {
  const size_t ns = sizeof(VieoNavState) * K, nx = 24 * P, nc = 8 * E;
  off[0] = 0;
  off[1] = (ns + 15) & ~(size_t)15;
  off[2] = off[1] + ((nx + 15) & ~(size_t)15);
  off[3] = off[2] + ((nc + 15) & ~(size_t)15);
  *total = off[3] + E;
}

// The LocalBA routine of vieo_local_ba_prv_begin captured once per handle.  The two optimize() stages are WHILE nodes
// spliced into the capture (cudaStreamGetCaptureInfo -> cudaGraphAddNode -> cudaStreamUpdateCaptureDependencies); their
// bodies are captured on a second stream.  Every launch parameter is a capacity or a constant: the same executable graph
// serves every window the handle ever sees, so begin() costs the upload + ONE launch instead of ~45 driver calls
// (tools/lba_async_probe.py: 0.31 ms of host time per window, 16 windows per bench step from one thread).
bool lba_build_window_graph(vieo_ba* h) {
  cudaStream_t side = nullptr;
  if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) return false;
  const int capE = h->capE, gridE = std::max((capE + 255) / 256, 1), gridK = (h->capK + 127) / 128;
  int nodes = 0;
  bool ok = cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  bool capturing = ok;
  auto campose = [&]() {
    k_ba_campose_p<<<gridK, 128, 0, h->st>>>(h->B.prm, h->B.st, h->B.cp);
    ++nodes;
  };
  auto errors = [&](int all, double* d_out, int skipvis) {
    campose();
    k_ba_errors<<<gridE, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_w, h->d_flags, h->d_lvl,
                                          h->d_sfix, 1, -1, all, 0.0, 0.0, h->B.chi2, h->B.partial, h->B.prm, skipvis);
    k_ba_dense_errors<<<1, 128, 0, h->st>>>(h->d_den, -1, h->B.st, h->d_pre, h->B.prm, h->B.wk, h->B.partial, 0, d_out, nullptr,
                                            all == 2 ? -1 : 0, nullptr, skipvis);
    nodes += 2;
  };
  auto classify = [&](int mode, float rat, int set_level, int remove_kernels, uint8_t* bad, int skip_if_stop, int skipvis) {
    k_ba_classify<<<gridE, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_flags, h->B.chi2, -1, mode,
                                            rat, set_level, remove_kernels, -1, h->d_lvl, bad, h->B.prm, skip_if_stop, skipvis);
    ++nodes;
  };
  auto stage = [&](int st_idx) {
    k_ba_begin<<<1, 1, 0, h->st>>>(h->B, st_idx, -1, 0.0);
    ok = ok && cudaMemsetAsync(h->B.x, 0, 8 * h->cap_np, h->st) == cudaSuccess;
    campose();
    ba_enqueue_linearize(h, 0, true);
    k_ba_control<<<1, 256, 0, h->st>>>(h->B, 0, 0);
    nodes += 4;
    // WHILE node after everything captured so far
    cudaStreamCaptureStatus status;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    ok = ok && cudaStreamGetCaptureInfo_v2(h->st, &status, nullptr, &g, &deps, &ndeps) == cudaSuccess &&
         status == cudaStreamCaptureStatusActive;
    cudaGraphConditionalHandle hc = 0;
    ok = ok && cudaGraphConditionalHandleCreate(&hc, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    cudaGraphNode_t node = nullptr;
    cudaGraph_t body = nullptr, tmp = nullptr;
    if (ok) {
      cudaGraphNodeParams np_ = {};
      np_.type = cudaGraphNodeTypeConditional;
      np_.conditional.handle = hc;
      np_.conditional.type = cudaGraphCondTypeWhile;
      np_.conditional.size = 1;
      ok = cudaGraphAddNode(&node, g, deps, ndeps, &np_) == cudaSuccess;
      if (ok) body = np_.conditional.phGraph_out[0];
    }
    if (ok) {
      cudaStream_t keep = h->st;
      h->st = side;
      ok = cudaStreamBeginCaptureToGraph(side, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        ba_enqueue_trial(h, true, hc);
        ok = cudaStreamEndCapture(side, &tmp) == cudaSuccess;
      }
      h->st = keep;
    }
    ok = ok && cudaStreamUpdateCaptureDependencies(h->st, &node, 1, cudaStreamSetCaptureDependencies) == cudaSuccess;
  };
  if (ok) {
    // Chi2LargeSetLevel(100) of the PRV version only (:534-536): the three kernels return at once for a visual-only problem
    errors(1, h->d_ctl + 5, 1);
    classify(0, 100.f, 1, 0, nullptr, 0, 1);
    errors(0, h->d_ctl + 6, 0);  // err (:539)
    stage(0);
    if (ok) {
      campose();  // inlier re-classification + kernel removal (:597-633); skipped once the abort flag is up
      classify(1, 0.f, 1, 1, h->d_bad, 1, 0);
      stage(1);
    }
    if (ok) {
      errors(2, h->d_ctl + 7, 0);  // err_end over the stored errors (:652)
      campose();                   // outlier candidates (:668-700)
      classify(1, 0.f, 0, 0, h->d_bad, 0, 0);
      size_t off[4], total;
      lba_dl_layout((size_t)h->capK, (size_t)h->capP, (size_t)capE, off, &total);
      ok = total <= h->stage_cap;
      uint8_t* hs = h->h_stage;
      ok = ok && cudaMemcpyAsync(hs + off[0], h->B.st, sizeof(VieoNavState) * (size_t)h->capK, cudaMemcpyDeviceToHost, h->st) == cudaSuccess;
      if (h->capP) ok = ok && cudaMemcpyAsync(hs + off[1], h->B.X, 24 * (size_t)h->capP, cudaMemcpyDeviceToHost, h->st) == cudaSuccess;
      if (capE) {
        ok = ok && cudaMemcpyAsync(hs + off[2], h->B.chi2, 8 * (size_t)capE, cudaMemcpyDeviceToHost, h->st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(hs + off[3], h->d_bad, (size_t)capE, cudaMemcpyDeviceToHost, h->st) == cudaSuccess;
      }
      ok = ok && cudaMemcpyAsync(h->h_ctl, h->d_ctl, 8 * 16, cudaMemcpyDeviceToHost, h->st) == cudaSuccess;
      ok = ok && cudaMemcpyAsync(h->h_prm, h->B.prm, sizeof(BaParams), cudaMemcpyDeviceToHost, h->st) == cudaSuccess;
    }
  }
  cudaGraph_t g = nullptr;
  if (capturing) {
    const bool ended = cudaStreamEndCapture(h->st, &g) == cudaSuccess;
    ok = ok && ended;
  }
  if (ok) ok = cudaGraphInstantiate(&h->win_graph, g, 0) == cudaSuccess;
  if (!ok) {
    h->win_graph = nullptr;
    cudaGetLastError();
  }
  if (g) cudaGraphDestroy(g);
  cudaStreamDestroy(side);
  h->win_nodes = nodes;
  h->launches = 0;
  return ok;
}

}  // namespace

extern "C" {

// Stream priority of the engines created from now on (process-wide): high (default) when a LocalMapping thread waits for
// its window and the chain of small kernels must not queue behind the tracker's; normal for throughput runs that keep many
// windows in flight beside the front-end (bench.py), where the front-end is the critical path.
static std::atomic<bool> g_ba_high_priority{true};
void vieo_ba_stream_priority(int high) { g_ba_high_priority.store(high != 0); }

static int ba_create_impl(int max_states, int max_points, int max_edges, int max_imu, int device, bool big, vieo_ba_t** out);
int vieo_ba_create(int max_states, int max_points, int max_edges, int max_imu, int device, vieo_ba_t** out) {
  return ba_create_impl(max_states, max_points, max_edges, max_imu, device, false, out);
}
int vieo_ba_create_global(int max_states, int max_points, int max_edges, int max_imu, int device, vieo_ba_t** out) {
  return ba_create_impl(max_states, max_points, max_edges, max_imu, device, true, out);
}
static int ba_create_impl(int max_states, int max_points, int max_edges, int max_imu, int device, bool big, vieo_ba_t** out) {
  VIEO_ARG(out && max_states > 0 && max_points >= 0 && max_edges >= 0 && max_imu >= 0, "bad argument");
  int rc = use_device(device);
  if (rc) return rc;
  vieo_ba* h = new vieo_ba();
  memset(&h->B, 0, sizeof(h->B));
  h->device = device;
  h->capK = max_states; h->capP = max_points; h->capE = max_edges; h->capM = max_imu;
  h->cap_pblk = (max_points + kBaWarps - 1) / kBaWarps;
  h->big = big;
  if (h->big && max_states > 768) {
    set_error("vieo_ba_create: at most 768 keyframes (the Schur row tile of one keyframe must fit one SM's shared memory)");
    delete h;
    return VIEO_E_CAPACITY;
  }
  h->cap_free = h->big ? max_states : std::min(max_states, kSchurMaxFree);
  const size_t K = max_states, P = max_points, E = max_edges, M = max_imu;
  // free keyframes (+ a few V/Bias-only) x 15; even so that double2 stores cover H exactly
  // (global handles: + 4 for the scale and gravity-direction border, count kept even)
  const size_t NP = h->big ? 15 * (size_t)(max_states + (max_states & 1)) + 4 : 15 * (size_t)std::min(max_states, kSchurMaxFree + 8);
  BaBuf& B = h->B;
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  // the engine's kernels are tiny and latency-critical: highest stream priority, and inside the calling thread's SM
  // partition when there is one (vieo_sm_partition_bind_thread)
  step(make_stream(&h->st, g_ba_high_priority.load()));
  step(dalloc(&B.prm, 1));
  step(dalloc(&B.st, K)); step(dalloc(&B.st_bak, K)); step(dalloc(&B.cp, K));
  step(dalloc(&B.X, 3 * P)); step(dalloc(&B.X_bak, 3 * P)); step(dalloc(&B.chi2, E));
  step(dalloc(&B.A, 27 * E)); step(dalloc(&B.Dinv, 9 * P)); step(dalloc(&B.db, 3 * P));
  for (int s = 0; s < 2; ++s) {
    step(dalloc(&B.W[s], 18 * E)); step(dalloc(&B.Hll[s], 9 * P)); step(dalloc(&B.bl[s], 3 * P));
    step(dalloc(&B.H[s], NP * NP)); step(dalloc(&B.b[s], NP)); step(dalloc(&B.pt_active[s], P));
  }
  step(dalloc(&h->d_sys, NP * NP + 2 * NP + 8)); step(dalloc(&B.x, NP));
  step(dalloc(&h->d_xl, 3 * P)); step(dalloc(&B.partial, std::max((E + 255) / 256, P / kBaWarps + 1) + 2));
  step(dalloc(&B.scale_part, P));
  if (h->big) {
    for (int s = 0; s < 2; ++s) step(dalloc(&B.wsp[s], 3 * P));
    step(dalloc(&B.up, 3 * P)); step(dalloc(&B.sred, 2 * P)); step(dalloc(&B.As, 8 * E));
    step(dalloc(&B.spart, 2 * ((size_t)h->cap_pblk + 2)));
    step(make_stream(&h->st_aux, true));
    step(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    step(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    {
      const size_t CN = 6 * K + 4, NY = 9 * K;
      step(dalloc(&B.cS, CN * CN)); step(dalloc(&B.cbs, CN)); step(dalloc(&B.cx, CN));
      step(dalloc(&B.vbL, kVbLS * K)); step(dalloc(&B.vbF, 81 * K)); step(dalloc(&B.Z, NY * (CN + 1)));
      step(dalloc(&h->d_pmap, CN)); step(dalloc(&h->d_ymap, NY)); step(dalloc(&h->d_ylo, CN)); step(dalloc(&h->d_yhi, CN));
    }
    step(dalloc(&B.part, 16));
    step(dalloc(&h->d_nz, (NP / kGNB + 1) * (NP / kGNB + 1)));
    step(dalloc(&h->d_yv, NP));
  } else {
    step(dalloc(&B.part, (size_t)h->cap_free * kSchurSplit * 6 * (6 * (size_t)h->cap_free + 1)));
  }
  step(dalloc(&h->d_ctl, 16)); step(dalloc(&h->d_es, E)); step(dalloc(&h->d_ep, E)); step(dalloc(&h->d_pt_ptr, P + 1));
  step(dalloc(&h->d_off0, K)); step(dalloc(&h->d_off1, K)); step(dalloc(&h->d_off2, K)); step(dalloc(&h->d_prcol, K));
  step(dalloc(&h->d_free_state, K)); step(dalloc(&h->d_free_off, K)); step(dalloc(&h->d_ps_ptr, K + 1));
  step(dalloc(&h->d_ps_edges, E)); step(dalloc(&h->d_obs, 3 * E)); step(dalloc(&h->d_w, E));
  step(dalloc(&h->d_ecol, E));
  step(dalloc(&h->d_flags, E)); step(dalloc(&h->d_lvl, E)); step(dalloc(&h->d_sfix, K));
  step(dalloc(&h->d_bad, E)); step(dalloc(&h->d_pre, M)); step(dalloc(&h->d_den, 2 * M + 1)); step(dalloc(&B.wk, 2 * M + 1));
  step(cudaEventCreate(&h->ev_begin)); step(cudaEventCreate(&h->ev_end));
  step(cudaMallocHost((void**)&h->h_ctl, sizeof(double) * 16));
  step(cudaMallocHost((void**)&h->h_prm, sizeof(BaParams)));
  h->stage_cap = K * (sizeof(VieoNavState) + 64) + P * 32 + E * 40 + M * (sizeof(VieoImuPreint) + 2 * sizeof(BaDense)) + 4096;
  step(cudaMallocHost((void**)&h->h_stage, h->stage_cap));
  if (e == cudaSuccess) {
    h->cap_np = NP;
    B.bs = h->d_sys; B.bsys = h->d_sys + NP; B.S = h->d_sys + 2 * NP;
    B.ecol = h->d_ecol;
    B.es = h->d_es; B.ep = h->d_ep; B.pt_ptr = h->d_pt_ptr; B.off0 = h->d_off0; B.off1 = h->d_off1; B.off2 = h->d_off2;
    B.prcol = h->d_prcol; B.free_state = h->d_free_state; B.free_off = h->d_free_off; B.ps_ptr = h->d_ps_ptr;
    B.ps_edges = h->d_ps_edges; B.obs = h->d_obs; B.w = h->d_w; B.flags = h->d_flags; B.lvl = h->d_lvl; B.sfix = h->d_sfix;
    B.pre = h->d_pre; B.den = h->d_den;
    B.pmap = h->d_pmap; B.ymap = h->d_ymap; B.ylo = h->d_ylo; B.yhi = h->d_yhi;
    memset(h->h_prm, 0, sizeof(BaParams));
    h->h_prm->done = 1;
    step(cudaMemcpy(B.prm, h->h_prm, sizeof(BaParams), cudaMemcpyHostToDevice));
    if (h->big) {
      step(cudaFuncSetAttribute(k_gba_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gba_schur_smem(h->cap_free)));
      step(cudaFuncSetAttribute(k_gchol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcholSmem));
      step(cudaFuncSetAttribute(k_gchol_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcholSmem));
      step(cudaFuncSetAttribute(k_gchol_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcholSmem));
      step(cudaFuncSetAttribute(k_gchol_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcholSmem));
    } else {
      step(cudaFuncSetAttribute(k_ba_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)schur_smem(h->cap_free)));
      step(cudaFuncSetAttribute(k_ba_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
    }
  }
  if (e == cudaSuccess && !h->big) {
    // the LM trial as one CUDA graph: constant launch parameters (capacity grids, sizes read from B.prm on the device)
    cudaGraph_t g = nullptr;
    step(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
    if (e == cudaSuccess) {
      ba_enqueue_trial(h, true);
      step(cudaStreamEndCapture(h->st, &g));
      if (e == cudaSuccess) step(cudaGraphInstantiate(&h->trial_graph, g, 0));
      if (g) cudaGraphDestroy(g);
    }
    // the optimize() graph; any failure here just leaves the chunked trial-graph loop in charge
    if (e == cudaSuccess) {
      cudaGraph_t og = nullptr, body = nullptr, tmp = nullptr;
      cudaGraphConditionalHandle hc = 0;
      bool ok = cudaGraphCreate(&og, 0) == cudaSuccess &&
                cudaGraphConditionalHandleCreate(&hc, og, 1, cudaGraphCondAssignDefault) == cudaSuccess;
      if (ok) {
        cudaGraphNodeParams np_ = {};
        np_.type = cudaGraphNodeTypeConditional;
        np_.conditional.handle = hc;
        np_.conditional.type = cudaGraphCondTypeWhile;
        np_.conditional.size = 1;
        cudaGraphNode_t node;
        ok = cudaGraphAddNode(&node, og, nullptr, 0, &np_) == cudaSuccess;
        if (ok) body = np_.conditional.phGraph_out[0];
      }
      if (ok) ok = cudaStreamBeginCaptureToGraph(h->st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        ba_enqueue_trial(h, true, hc);
        ok = cudaStreamEndCapture(h->st, &tmp) == cudaSuccess;
      }
      if (ok) ok = cudaGraphInstantiate(&h->opt_graph, og, 0) == cudaSuccess;
      if (!ok) {
        h->opt_graph = nullptr;
        cudaGetLastError();
      }
      if (og) cudaGraphDestroy(og);
      if (ok) ok = make_stream(&h->st_ctl, true) == cudaSuccess;
      if (!ok && h->opt_graph) {
        cudaGraphExecDestroy(h->opt_graph);
        h->opt_graph = nullptr;
      }
      // the whole-routine graph; a failure leaves the kernel-by-kernel enqueue of vieo_local_ba_prv_begin in charge
      if (ok && !getenv("VIEO_BA_NO_WINDOW_GRAPH")) lba_build_window_graph(h);
    }
    h->launches = 0;
  }
  if (e != cudaSuccess) {
    set_error("vieo_ba_create: %s", cudaGetErrorString(e));
    ba_free(h);
    delete h;
    return VIEO_E_CUDA;
  }
  *out = h;
  return VIEO_OK;
}

void vieo_ba_destroy(vieo_ba_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  ba_free(h);
  delete h;
}

int vieo_ba_set_comm(vieo_ba_t* h, vieo_comm_t* comm) {
  VIEO_ARG(h, "null handle");
  h->comm = comm;
  h->allreduce = nullptr;
  h->ar_ctx = nullptr;
  h->rank = 0;
  h->world = 1;
  if (comm) vieo::comm_info(comm, &h->rank, &h->world);
  return VIEO_OK;
}
int vieo_ba_set_sharding(vieo_ba_t* h, int rank, int world, vieo_allreduce_fn allreduce, void* ctx) {
  VIEO_ARG(h && world >= 1 && rank >= 0 && rank < world, "bad argument");
  h->rank = rank; h->world = world; h->allreduce = allreduce; h->ar_ctx = ctx;
  h->comm = nullptr;
  return VIEO_OK;
}

void* vieo_ba_stream(vieo_ba_t* h) { return h ? (void*)h->st : nullptr; }
int vieo_ba_last_launches(const vieo_ba_t* h) { return h ? h->launches : 0; }
double vieo_ba_last_ms(const vieo_ba_t* h) { return h ? (double)h->last_ms : 0.0; }
int vieo_ba_last_trials(const vieo_ba_t* h) { return h && h->h_prm ? h->h_prm->trials : 0; }

int vieo_ba_set_problem(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam) {
  VIEO_ARG(h && pb && cam, "null argument");
  const int K = pb->n_states, P = pb->n_points, E = pb->n_edges;
  const int M = (pb->visual_only || h->rank != 0) ? 0 : pb->n_imu;
  if (K > h->capK || P > h->capP || E > h->capE || M > h->capM) {
    set_error("vieo_ba_set_problem: problem (%d states, %d points, %d edges, %d imu) exceeds the handle capacity", K, P, E, M);
    return VIEO_E_CAPACITY;
  }
  VIEO_ARG(K > 0 && pb->states && pb->state_flags, "no states");
  VIEO_ARG(E == 0 || (pb->edge_state && pb->edge_point && pb->obs && pb->inv_sigma2 && pb->edge_flags && pb->points), "null edge array");
  // the whole problem is validated BEFORE any handle state changes: a rejected problem leaves the previous one usable
  for (int i = 0; i < E; ++i) {
    VIEO_ARG(pb->edge_point[i] >= 0 && pb->edge_point[i] < P && pb->edge_state[i] >= 0 && pb->edge_state[i] < K, "edge index out of range");
    VIEO_ARG(i == 0 || pb->edge_point[i - 1] <= pb->edge_point[i], "edges must be sorted by point");
  }
  VIEO_ARG(M == 0 || (pb->imu_i && pb->imu_j && pb->preint && pb->imu_dt_kf), "null inertial array");
  for (int m = 0; m < M; ++m)
    VIEO_ARG(pb->imu_i[m] >= 0 && pb->imu_i[m] < K && pb->imu_j[m] >= 0 && pb->imu_j[m] < K, "imu state index out of range");
  VIEO_ARG(cam->model >= 0 && cam->model <= 2 && cam->num_k >= 0 && cam->num_k <= 6, "unsupported camera model");
  VIEO_ARG(!(pb->global_ba & (4 | 8 | 16)) || h->big, "the scale / gravity-direction vertices need a handle from vieo_ba_create_global");
  {
    int nfree_ = 0;
    size_t np_ = 0;
    for (int k = 0; k < K; ++k) {
      const uint8_t f = pb->state_flags[k];
      if (!(f & 1)) { ++nfree_; np_ += 6; }
      if ((f & 2) && !(f & 4) && !pb->visual_only) np_ += 9;
    }
    np_ += ((pb->global_ba & 4) ? 1 : 0) + ((pb->global_ba & 8) ? 2 : 0);
    if (nfree_ > h->cap_free || np_ > h->cap_np) {
      set_error("vieo_ba_set_problem: %d free keyframes / %zu pose dimensions exceed the engine's tile (%d / %zu)", nfree_, np_,
                h->cap_free, h->cap_np);
      return VIEO_E_CAPACITY;
    }
  }
  BA_CK(cudaSetDevice(h->device));
  h->K = K; h->P = P; h->E = E; h->M = M;
  h->launches = 0;
  h->points_free = true;
  h->n_part = (E + 255) / 256;
  h->n_pblk = (P + kBaWarps - 1) / kBaWarps;
  // camera
  cam_set(h->cam, *cam);
  VIEO_ARG(cam->model >= 0 && cam->model <= 2 && cam->num_k >= 0 && cam->num_k <= 6, "unsupported camera model");
  h->gw = {pb->gw[0], pb->gw[1], pb->gw[2]};
  const float chi2Mono = 5.991f;
  h->visual_only = pb->visual_only != 0;
  // thHuberMono: src/Optimizer.cc:361 (PRV: sqrt of the float 5.991f) vs :2069 (visual LocalBundleAdjustment: sqrt(5.991))
  h->dm = h->visual_only ? (double)(float)std::sqrt(5.991) : (double)std::sqrt(chi2Mono);
  h->ds = (double)(float)std::sqrt(7.815);       // thHuberStereo
  const bool global = pb->global_ba & 1, g_robust = pb->global_ba & 2;
  if (global) h->dm = (double)(float)std::sqrt(5.99);  // thHuber2D of the global BA (src/Optimizer.cc:1042)
  // index mapping (sparse_optimizer.cpp:166-190): states in order, PR, V, Bias
  h->off0.assign(K, -1); h->off1.assign(K, -1); h->off2.assign(K, -1);
  std::vector<int> prcol(K, -1), free_state, free_off;
  std::vector<uint8_t> sfix(K);
  int np = 0;
  for (int k = 0; k < K; ++k) {
    const uint8_t f = pb->state_flags[k];
    sfix[k] = f & 1;
    if (!(f & 1)) {
      h->off0[k] = np;
      prcol[k] = (int)free_state.size();
      free_state.push_back(k);
      free_off.push_back(np);
      np += 6;
    }
    if ((f & 2) && !(f & 4) && !pb->visual_only) {
      h->off1[k] = np; np += 3;
      h->off2[k] = np; np += 6;
    }
  }
  // border vertices after every keyframe vertex: VertexScale (id maxKFid + 1), VertexGThetaXYRwI (id maxKFid + 2)
  const bool want_scale = pb->global_ba & 4, want_g = pb->global_ba & 8, want_prior_bias = pb->global_ba & 16;
  VIEO_ARG(!(want_scale || want_g) || h->big, "the scale / gravity-direction vertices need a handle from vieo_ba_create_global");
  h->has_scale = want_scale;
  h->has_g = want_g;
  int off_s = -1, off_g = -1;
  if (want_scale) { off_s = np; np += 1; }
  if (want_g) { off_g = np; np += 2; }
  h->np = np;
  h->nfree = (int)free_state.size();
  // point ranges + per-keyframe edge lists
  std::vector<int> pt_ptr(P + 1, 0), ps_ptr(h->nfree + 1, 0), ps_edges;
  for (int i = 0; i < E; ++i) {
    const int p = pb->edge_point[i], s = pb->edge_state[i];
    VIEO_ARG(p >= 0 && p < P && s >= 0 && s < K, "edge index out of range");
    VIEO_ARG(i == 0 || pb->edge_point[i - 1] <= p, "edges must be sorted by point");
    pt_ptr[p + 1]++;
    if (prcol[s] >= 0) ps_ptr[prcol[s] + 1]++;
  }
  for (int p = 0; p < P; ++p) pt_ptr[p + 1] += pt_ptr[p];
  for (int f = 0; f < h->nfree; ++f) ps_ptr[f + 1] += ps_ptr[f];
  ps_edges.resize(std::max(ps_ptr[h->nfree], 1));
  {
    std::vector<int> fill(ps_ptr.begin(), ps_ptr.end() - 1);
    for (int i = 0; i < E; ++i) {
      const int c = prcol[pb->edge_state[i]];
      if (c >= 0) ps_edges[fill[c]++] = i;
    }
  }
  // several edges between one point and one keyframe (multi-camera rigs) make lanes collide in k_ba_schur
  h->has_dup = false;
  for (int p = 0; p < P && !h->has_dup; ++p)
    for (int a = pt_ptr[p]; a < pt_ptr[p + 1] && !h->has_dup; ++a)
      for (int c = a + 1; c < pt_ptr[p + 1]; ++c)
        if (pb->edge_state[a] == pb->edge_state[c]) {
          h->has_dup = true;
          break;
        }
  // inertial factors (src/Optimizer.cc:219-330)
  std::vector<BaDense> den;
  const float thPRV = (float)std::sqrt(16.919), thBias = (float)std::sqrt(12.592);
  for (int m = 0; m < M; ++m) {
    const int i = pb->imu_i[m], j = pb->imu_j[m];
    VIEO_ARG(i >= 0 && i < K && j >= 0 && j < K, "imu state index out of range");
    // local BA: the previous KEYFRAME is fixed (src/Optimizer.cc:262-271); global BA: its BIAS vertex is (:955-958)
    const bool bfixedkf = global ? (!(pb->state_flags[i] & 2) || (pb->state_flags[i] & 4)) : (pb->state_flags[i] & 1);
    const bool kernel = global ? g_robust : (bfixedkf || pb->rec_init);
    const VieoImuPreint& pre = pb->preint[m];
    if (pre.dt != 0) {
      BaDense d;
      memset(&d, 0, sizeof(d));
      d.type = 0; d.si = i; d.sj = j; d.pre = m;
      if (!host_inverse(pre.SigmaPRV, 9, d.info))
        for (double& v : d.info) v = std::numeric_limits<double>::quiet_NaN();
      if (bfixedkf) for (double& v : d.info) v *= 1e-2;
      if (kernel) d.delta = (double)thPRV;
      den.push_back(d);
    }
    BaDense d;
    memset(&d, 0, sizeof(d));
    d.type = 1; d.si = i; d.sj = j; d.pre = m;
    double dtij = pre.dt != 0 ? pre.dt : pb->imu_dt_kf[m];
    if (dtij <= (double)1e-6f) dtij = 15;
    for (int k = 0; k < 6; ++k) {
      const double w = (k < 3 ? pb->inv_sigma_bg2 : pb->inv_sigma_ba2) / dtij;
      d.info[k] = bfixedkf ? w * 1e-2 : w;
    }
    if (kernel) d.delta = (double)thBias;
    den.push_back(d);
  }
  // the IMU initialiser's prior-bias edge (src/Optimizer.cc:866-900, 1026-1054): EdgeNavStateBias from a fixed copy of the
  // earliest keyframe's bias (states[0]) with information invSigma / sum(dt_ij) over every keyframe pair (:940-947)
  if (want_prior_bias && M > 0 && (pb->state_flags[0] & 2)) {
    double sum_dt = 0;
    for (int m = 0; m < M; ++m) {
      double dtij = pb->preint[m].dt != 0 ? pb->preint[m].dt : pb->imu_dt_kf[m];
      if (dtij <= (double)1e-6f) dtij = 15;
      sum_dt += dtij;
    }
    BaDense d;
    memset(&d, 0, sizeof(d));
    d.type = 3; d.si = 0; d.sj = 0; d.pre = 0;
    for (int k = 0; k < 6; ++k) d.info[k] = (k < 3 ? pb->inv_sigma_bg2 : pb->inv_sigma_ba2) / sum_dt;
    const VieoNavState& s0 = pb->states[0];
    for (int k = 0; k < 3; ++k) {
      d.info[6 + k] = s0.bg[k] + s0.dbg[k];
      d.info[9 + k] = s0.ba[k] + s0.dba[k];
    }
    den.push_back(d);
  }
  // greedy colouring of the inertial edges by shared keyframes (a chain needs two colours)
  int n_colors = 1;
  {
    std::vector<std::vector<int>> used(K);
    for (BaDense& d : den) {
      int c = 0;
      for (;;) {
        bool clash = false;
        for (int k : {d.si, d.sj})
          for (int u : used[k]) clash |= u == c;
        if (!clash) break;
        ++c;
      }
      d.color = c;
      used[d.si].push_back(c);
      used[d.sj].push_back(c);
      n_colors = std::max(n_colors, c + 1);
    }
  }
  h->n_den = (int)den.size();
  VIEO_ARG(h->n_den <= 2 * h->capM + 1, "too many inertial edges");
  // V / Bias elimination (global handles): the keyframes with free V / Bias vertices form a chain in state order; usable
  // when every inertial / bias-walk edge joins neighbours of that chain (consecutive keyframes, as the reference builds
  // them from GetPrevKeyFrame()).  VIEO_GBA_NO_VB=1 keeps the 15-dimensional blocks in the dense factorisation.
  h->vb_elim = false;
  h->cn = h->Ky = 0;
  std::vector<int> pmap, ymap, ylo, yhi;
  if (h->big && !getenv("VIEO_GBA_NO_VB")) {
    std::vector<int> yblk(K, -1), prow(K, -1);
    int Ky = 0;
    for (int k = 0; k < K; ++k) {
      if (h->off0[k] >= 0) {
        prow[k] = (int)pmap.size();
        for (int t = 0; t < 6; ++t) pmap.push_back(h->off0[k] + t);
      }
      if (h->off1[k] >= 0) {
        yblk[k] = Ky++;
        for (int t = 0; t < 3; ++t) ymap.push_back(h->off1[k] + t);
        for (int t = 0; t < 6; ++t) ymap.push_back(h->off2[k] + t);
      }
    }
    const int n_kf_rows = (int)pmap.size();
    if (want_scale) pmap.push_back(off_s);
    if (want_g) { pmap.push_back(off_g); pmap.push_back(off_g + 1); }
    ylo.assign(pmap.size(), 1 << 30);
    yhi.assign(pmap.size(), 0);
    bool chain = Ky > 0;
    // The structure comes from the PROBLEM's inertial topology, not from this rank's edge list: a sharded handle keeps the
    // inertial edges on rank 0 only, but after the all-reduce every rank eliminates the same chain with the same PR <-> V /
    // Bias coupling ranges (a rank that derived them from its empty list solved a different system: states off by 1e-2)
    struct Topo { int type, si, sj; };
    std::vector<Topo> topo;
    if (!pb->visual_only && pb->imu_i && pb->imu_j && pb->preint)
      for (int m = 0; m < pb->n_imu; ++m) {
        VIEO_ARG(pb->imu_i[m] >= 0 && pb->imu_i[m] < K && pb->imu_j[m] >= 0 && pb->imu_j[m] < K, "imu state index out of range");
        if (pb->preint[m].dt != 0) topo.push_back({0, pb->imu_i[m], pb->imu_j[m]});
        topo.push_back({1, pb->imu_i[m], pb->imu_j[m]});
      }
    for (const Topo& d : topo) {
      const int a = yblk[d.si], b = yblk[d.sj];
      if (a >= 0 && b >= 0 && std::abs(a - b) != 1) chain = false;
      if (d.type != 0) continue;
      for (int s2 : {d.si, d.sj}) {
        if (prow[s2] < 0) continue;
        for (int yb : {a, b}) {
          if (yb < 0) continue;
          for (int t = 0; t < 6; ++t) {
            ylo[prow[s2] + t] = std::min(ylo[prow[s2] + t], yb);
            yhi[prow[s2] + t] = std::max(yhi[prow[s2] + t], yb + 1);
          }
        }
      }
    }
    for (size_t r = 0; r < pmap.size(); ++r) {
      if ((int)r >= n_kf_rows) { ylo[r] = 0; yhi[r] = Ky; }  // border rows (scale, gravity direction): the whole chain
      else if (ylo[r] > yhi[r]) ylo[r] = yhi[r] = 0;
    }
    if (chain) {
      h->vb_elim = true;
      h->cn = (int)pmap.size();
      h->Ky = Ky;
    }
  }
  std::vector<uint8_t> lvl(std::max(E, 1));
  for (int i = 0; i < E; ++i)
    lvl[i] = ((pb->edge_flags[i] & VIEO_EDGE_LEVEL1) ? 1 : 0) | (((pb->edge_flags[i] & VIEO_EDGE_NOKERNEL) || (global && !g_robust)) ? 2 : 0);
  // every array goes through the pinned staging buffer: asynchronous copies, no driver-side pageable staging
  h->stage_used = 0;
  h->upload_direct = false;
  auto up = [&](void* d, const void* s, size_t n) {
    if (!n) return cudaSuccess;
    const size_t o = (h->stage_used + 15) & ~(size_t)15;
    if (o + n > h->stage_cap) {
      h->upload_direct = true;
      return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, h->st);
    }
    memcpy(h->h_stage + o, s, n);
    h->stage_used = o + n;
    return cudaMemcpyAsync(d, h->h_stage + o, n, cudaMemcpyHostToDevice, h->st);
  };
  BA_CK(up(h->B.st, pb->states, sizeof(VieoNavState) * K));
  BA_CK(up(h->B.X, pb->points, 24 * (size_t)P));
  BA_CK(up(h->d_es, pb->edge_state, 4 * (size_t)E));
  BA_CK(up(h->d_ep, pb->edge_point, 4 * (size_t)E));
  BA_CK(up(h->d_obs, pb->obs, 12 * (size_t)E));
  BA_CK(up(h->d_w, pb->inv_sigma2, 4 * (size_t)E));
  BA_CK(up(h->d_flags, pb->edge_flags, (size_t)E));
  BA_CK(up(h->d_lvl, lvl.data(), (size_t)E));
  BA_CK(up(h->d_sfix, sfix.data(), (size_t)K));
  BA_CK(up(h->d_pt_ptr, pt_ptr.data(), 4 * (size_t)(P + 1)));
  BA_CK(up(h->d_off0, h->off0.data(), 4 * (size_t)K));
  BA_CK(up(h->d_off1, h->off1.data(), 4 * (size_t)K));
  BA_CK(up(h->d_off2, h->off2.data(), 4 * (size_t)K));
  BA_CK(up(h->d_prcol, prcol.data(), 4 * (size_t)K));
  {
    std::vector<int> ecol((size_t)std::max(E, 1));
    for (int i = 0; i < E; ++i) ecol[i] = prcol[pb->edge_state[i]];
    BA_CK(up(h->d_ecol, ecol.data(), 4 * (size_t)E));
  }
  BA_CK(up(h->d_free_state, free_state.data(), 4 * (size_t)h->nfree));
  BA_CK(up(h->d_free_off, free_off.data(), 4 * (size_t)h->nfree));
  BA_CK(up(h->d_ps_ptr, ps_ptr.data(), 4 * (size_t)(h->nfree + 1)));
  BA_CK(up(h->d_ps_edges, ps_edges.data(), 4 * (size_t)ps_ptr[h->nfree]));
  BA_CK(up(h->d_pre, pb->preint, sizeof(VieoImuPreint) * (size_t)M));
  BA_CK(up(h->d_den, den.data(), sizeof(BaDense) * den.size()));
  if (h->vb_elim) {
    BA_CK(up(h->d_pmap, pmap.data(), 4 * pmap.size()));
    BA_CK(up(h->d_ymap, ymap.data(), 4 * ymap.size()));
    BA_CK(up(h->d_ylo, ylo.data(), 4 * ylo.size()));
    BA_CK(up(h->d_yhi, yhi.data(), 4 * yhi.size()));
  }
  BA_CK(cudaMemsetAsync(h->B.chi2, 0, 8 * (size_t)std::max(E, 1), h->st));
  BA_CK(cudaMemsetAsync(h->d_ctl, 0, 8 * 16, h->st));
  BA_CK(cudaMemsetAsync(h->B.x, 0, 8 * (size_t)std::max(np, 1), h->st));
  BA_CK(cudaMemsetAsync(h->B.scale_part, 0, 8 * (size_t)std::max(P, 1), h->st));
  if (h->nfree > h->cap_free || (size_t)np > h->cap_np) {
    set_error("vieo_ba_set_problem: %d free keyframes / %d pose dimensions exceed the engine's tile (%d / %zu)", h->nfree, np,
              h->cap_free, h->cap_np);
    return VIEO_E_CAPACITY;
  }
  BaParams& q = *h->h_prm;
  memset(&q, 0, sizeof(q));
  q.K = K; q.P = P; q.E = E; q.np = np; q.nfree = h->nfree; q.n_den = h->n_den; q.n_pblk = h->n_pblk;
  q.has_dup = h->has_dup ? 1 : 0;
  q.big = h->big ? 1 : 0;
  q.n_colors = n_colors;
  h->n_colors = n_colors;
  q.rank = h->rank;
  q.world = h->sharded() ? h->world : 1;
  q.lambda_on_poses = h->rank == 0 ? 1 : 0;
  q.done = 1;
  q.ok = 1;
  q.ni = 2;
  q.cam = h->cam; q.gw = h->gw; q.dm = h->dm; q.ds = h->ds;
  q.visual_only = h->visual_only ? 1 : 0;
  q.plan_it[0] = h->plan_it[0]; q.plan_it[1] = h->plan_it[1]; q.plan_lambda0 = h->plan_lambda0;
  q.vb_elim = h->vb_elim ? 1 : 0; q.cn = h->cn; q.Ky = h->Ky;
  q.has_scale = want_scale ? 1 : 0; q.off_s = off_s;
  q.has_g = want_g ? 1 : 0; q.off_g = off_g;
  q.sc = q.sc_bak = (want_scale && pb->scale_init > 0) ? pb->scale_init : 1.0;
  q.qwI[0] = q.qwI_bak[0] = 1.0;
  if (want_g) {
    // VertexGThetaXYRwI::setToOriginImpl(gw) (g2otypes.h:682-690): RwI = Exp(normalized((0,0,1) x gw/|gw|) acos(gw_z/|gw|));
    // Eigen's normalized() leaves a zero vector unchanged, so gw (anti-)parallel to z gives the identity
    const double n = std::sqrt(pb->gw[0] * pb->gw[0] + pb->gw[1] * pb->gw[1] + pb->gw[2] * pb->gw[2]);
    VIEO_ARG(n > 0, "zero gravity vector");
    const double g[3] = {pb->gw[0] / n, pb->gw[1] / n, pb->gw[2] / n};
    const double a[3] = {-g[1], g[0], 0.0};
    const double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const double th = std::acos(g[2]), inv = na > 0 ? 1.0 / na : 1.0;
    host_so3_exp_q(a[0] * inv * th, a[1] * inv * th, a[2] * inv * th, q.qwI);
    for (int k = 0; k < 4; ++k) q.qwI_bak[k] = q.qwI[k];
    q.GI[0] = 0; q.GI[1] = 0; q.GI[2] = n;
  }
  BA_CK(cudaMemcpyAsync(h->B.prm, &q, sizeof(q), cudaMemcpyHostToDevice, h->st));
  // the host staging vectors die here: synchronise unless every array went through the pinned staging buffer and the
  // caller (vieo_local_ba_prv_begin) keeps the handle's host mirrors untouched until its end call
  if (h->upload_direct || !h->defer_sync) BA_CK(cudaStreamSynchronize(h->st));
  return VIEO_OK;
}

int vieo_ba_chi2_large_set_level(vieo_ba_t* h, float rat) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  int rc = ba_errors(h, 1, h->d_ctl + 5);
  if (rc) return rc;
  if (h->E > 0) {
    k_ba_classify<<<(h->E + 255) / 256, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_flags,
                                                         h->B.chi2, h->E, 0, rat, 1, 0, 1, h->d_lvl, nullptr, h->B.prm);
    h->launches++;
  }
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_ba_active_robust_chi2(vieo_ba_t* h, int recompute, double* chi2) {
  VIEO_ARG(h && chi2, "null argument");
  BA_CK(cudaSetDevice(h->device));
  if (recompute) {
    int rc = ba_errors(h, 0, h->d_ctl + 6);
    if (rc) return rc;
  } else {
    // activeRobustChi2 over the stored errors of the current active set (levels may have changed since the last
    // evaluation; after a rejected last trial the stored errors are those of the rejected estimate, as in g2o)
    int rc = ba_errors(h, 2, h->d_ctl + 6);
    if (rc) return rc;
  }
  {
    int rc2 = ba_allreduce(h, h->d_ctl + 6, 1);
    if (rc2) return rc2;
  }
  BA_CK(cudaMemcpyAsync(h->h_ctl + 6, h->d_ctl + 6, 8, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  *chi2 = h->h_ctl[6];
  return VIEO_OK;
}

int vieo_ba_optimize(vieo_ba_t* h, int iterations, double lambda_init, const volatile uint8_t* stop) {
  VIEO_ARG(h && iterations >= 0, "bad argument");
  BA_CK(cudaSetDevice(h->device));
  if (h->np == 0 || iterations == 0) return 0;
  const bool sharded = h->sharded();
  // single GPU: nothing to do once the caller asked to stop.  Sharded: a rank must not leave before its peers do — its
  // flag goes into the all-reduced trial record instead and k_ba_control ends the call on every rank at once.
  if (!sharded && stop && *stop) return 0;
  // optimize() state machine on the device: reset, then a fresh linearisation at the current estimate (levels / kernels
  // may have changed since the last call)
  BaParams& q = *h->h_prm;
  q.cur = 0; q.done = 0; q.stop = 0; q.ok = 1;
  q.iteration = 0; q.iters_target = iterations; q.iters_done = 0; q.qmax = 0; q.nBad = 0; q.trials = 0;
  q.user_lambda = lambda_init;
  if (sharded && stop && *stop) q.stop = 1;
  BA_CK(cudaMemcpyAsync(h->B.prm, &q, sizeof(q), cudaMemcpyHostToDevice, h->st));
  BA_CK(cudaMemsetAsync(h->B.x, 0, 8 * (size_t)h->np, h->st));
  int rc = ba_campose(h);
  if (rc) return rc;
  if ((rc = ba_enqueue_linearize(h, 0, false))) return rc;
  if ((rc = ba_allreduce(h, h->B.prm->pair, 3))) return rc;
  if (sharded && !(lambda_init > 0)) {  // g2o's own initial lambda needs the global maximum of the diagonal
    k_ba_diag_pack<<<1, 256, 0, h->st>>>(h->B);
    h->launches++;
    if ((rc = ba_allreduce(h, h->B.S, (size_t)h->np + h->world))) return rc;
  }
  k_ba_control<<<1, 256, 0, h->st>>>(h->B, 0, 0);
  h->launches++;
  BA_CK(cudaGetLastError());
  if (!sharded && h->opt_graph) {
    // ONE launch for the whole optimize(): the graph's WHILE node repeats the trial until k_ba_control says done.  The
    // caller's abort flag (pbStopFlag, polled once per iteration by sparse_optimizer.cpp:376) is forwarded through a side
    // stream while the graph runs.
    BA_CK(cudaGraphLaunch(h->opt_graph, h->st));
    h->launches += kNodesPerTrial;
    if (stop) {
      bool sent = false;
      while (cudaStreamQuery(h->st) == cudaErrorNotReady) {
        if (*stop && !sent) {
          static const int one = 1;
          BA_CK(cudaMemcpyAsync(&h->B.prm->stop, &one, sizeof(int), cudaMemcpyHostToDevice, h->st_ctl));
          sent = true;
        }
        std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
    }
    if ((rc = ba_read_prm(h))) return rc;
    h->launches += kNodesPerTrial * std::max(h->h_prm->trials - 1, 0);
    BA_CK(cudaGetLastError());
    return h->h_prm->iters_done;
  }
  // trials are launched in small chunks (a finished optimize() turns the remaining ones into no-ops); the abort flag is
  // polled between chunks (sparse_optimizer.cpp:376 polls it once per iteration)
  const int kChunk = 3, kNodes = kNodesPerTrial;
  for (int guard = 0; guard < 10 * iterations + 4; guard += kChunk) {
    for (int c = 0; c < kChunk; ++c) {
      if (!sharded && h->trial_graph) {
        BA_CK(cudaGraphLaunch(h->trial_graph, h->st));
        h->launches += kNodes;
      } else if ((rc = ba_enqueue_trial(h, false)))
        return rc;
    }
    if ((rc = ba_read_prm(h))) return rc;
    if (h->h_prm->done) break;
    if (stop && *stop && !h->h_prm->stop) {
      const int one = 1;
      BA_CK(cudaMemcpyAsync(&h->B.prm->stop, &one, sizeof(int), cudaMemcpyHostToDevice, h->st));
    }
  }
  BA_CK(cudaGetLastError());
  return h->h_prm->iters_done;
}

int vieo_ba_reclassify(vieo_ba_t* h, int remove_kernels, uint8_t* bad_host) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  if (h->E == 0) return VIEO_OK;
  {
    int rc = ba_campose(h);
    if (rc) return rc;
  }
  k_ba_classify<<<(h->E + 255) / 256, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_flags,
                                                       h->B.chi2, h->E, 1, 0.f, bad_host ? 0 : 1, remove_kernels, h->visual_only ? 0 : 1, h->d_lvl,
                                                       h->d_bad, h->B.prm);
  h->launches++;
  BA_CK(cudaGetLastError());
  if (bad_host) {
    const bool staged = (size_t)h->E <= h->stage_cap;
    BA_CK(cudaMemcpyAsync(staged ? (void*)h->h_stage : (void*)bad_host, h->d_bad, (size_t)h->E, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaStreamSynchronize(h->st));
    if (staged) memcpy(bad_host, h->h_stage, (size_t)h->E);
  }
  return VIEO_OK;
}

int vieo_ba_get(vieo_ba_t* h, VieoNavState* states_out, double* points_out, double* edge_chi2) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  const size_t ns = states_out ? sizeof(VieoNavState) * h->K : 0, nx = points_out ? 24 * (size_t)h->P : 0,
               nc = edge_chi2 ? 8 * (size_t)h->E : 0;
  if (ns + nx + nc + 64 <= h->stage_cap) {  // through pinned staging (the upload staging is no longer needed)
    uint8_t* ps = h->h_stage;
    uint8_t* px = ps + ((ns + 15) & ~(size_t)15);
    uint8_t* pc = px + ((nx + 15) & ~(size_t)15);
    if (ns) BA_CK(cudaMemcpyAsync(ps, h->B.st, ns, cudaMemcpyDeviceToHost, h->st));
    if (nx) BA_CK(cudaMemcpyAsync(px, h->B.X, nx, cudaMemcpyDeviceToHost, h->st));
    if (nc) BA_CK(cudaMemcpyAsync(pc, h->B.chi2, nc, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaStreamSynchronize(h->st));
    if (ns) memcpy(states_out, ps, ns);
    if (nx) memcpy(points_out, px, nx);
    if (nc) memcpy(edge_chi2, pc, nc);
    return VIEO_OK;
  }
  if (ns) BA_CK(cudaMemcpyAsync(states_out, h->B.st, ns, cudaMemcpyDeviceToHost, h->st));
  if (nx) BA_CK(cudaMemcpyAsync(points_out, h->B.X, nx, cudaMemcpyDeviceToHost, h->st));
  if (nc) BA_CK(cudaMemcpyAsync(edge_chi2, h->B.chi2, nc, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  return VIEO_OK;
}

// estimates of the border vertices: VertexScale, and RwI * GI of VertexGThetaXYRwI (gw_out nullable)
int vieo_ba_get_border(vieo_ba_t* h, double* scale_out, double* gw_out) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  int rc = ba_read_prm(h);
  if (rc) return rc;
  const BaParams& q = *h->h_prm;
  if (scale_out) *scale_out = q.has_scale ? q.sc : 1.0;
  if (gw_out) {
    if (!q.has_g) {
      gw_out[0] = q.gw.x; gw_out[1] = q.gw.y; gw_out[2] = q.gw.z;
    } else {  // q_matrix(qwI) * (0, 0, G): third column of the rotation matrix (so3.cuh q_matrix) times G
      const double w = q.qwI[0], x = q.qwI[1], y = q.qwI[2], z = q.qwI[3];
      const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
      const double twx = tx * w, twy = ty * w, txx = tx * x, txz = tz * x, tyy = ty * y, tyz = tz * y;
      gw_out[0] = (txz + twy) * q.GI[2];
      gw_out[1] = (tyz - twx) * q.GI[2];
      gw_out[2] = (1 - (txx + tyy)) * q.GI[2];
    }
  }
  return VIEO_OK;
}

int vieo_ba_debug_step(vieo_ba_t* h, double lambda, double* x_pose, double* x_points, double* H_out, double* b_out) {
  VIEO_ARG(h && x_pose, "null argument");
  BA_CK(cudaSetDevice(h->device));
  BaParams& q = *h->h_prm;
  q.cur = 0; q.done = 0; q.stop = 0; q.ok = 1;
  BA_CK(cudaMemcpyAsync(h->B.prm, &q, sizeof(q), cudaMemcpyHostToDevice, h->st));
  int rc = ba_campose(h);
  if (rc) return rc;
  if ((rc = ba_enqueue_linearize(h, 0, false))) return rc;
  if ((rc = ba_enqueue_solve(h, false, 1, lambda, h->d_xl))) return rc;
  BA_CK(cudaGetLastError());
  BA_CK(cudaMemcpyAsync(x_pose, h->B.x, 8 * (size_t)h->np, cudaMemcpyDeviceToHost, h->st));
  if (x_points && h->P) BA_CK(cudaMemcpyAsync(x_points, h->d_xl, 24 * (size_t)h->P, cudaMemcpyDeviceToHost, h->st));
  if (H_out) BA_CK(cudaMemcpyAsync(H_out, h->B.H[0], 8 * (size_t)h->np * h->np, cudaMemcpyDeviceToHost, h->st));
  if (b_out) BA_CK(cudaMemcpyAsync(b_out, h->B.b[0], 8 * (size_t)h->np, cudaMemcpyDeviceToHost, h->st));
  if ((rc = ba_read_prm(h))) return rc;
  const int ok = h->h_prm->ok;
  q.done = 1;
  BA_CK(cudaMemcpyAsync(h->B.prm, &q, sizeof(q), cudaMemcpyHostToDevice, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  if (!ok) {
    set_error("vieo_ba_debug_step: reduced camera system is not positive definite");
    return VIEO_E_ARG;
  }
  return h->np;
}

// Optimizer::GlobalBundleAdjustmentNavStatePRV, src/Optimizer.cc:771-1342, on the flattened problem: one
// optimize(nIterations) with g2o's own initial lambda, no outlier pass.  ex (nullable) selects the two variants:
//   scale_opt (bScaleOpt, System::FinalGBA src/System.cc:24-29): VertexScale seeded with 1, EdgeReprojectPRS[Stereo]; on
//     return the points are multiplied by the recovered scale (:1311-1334) and ex->scale holds the vertex estimate;
//   imu_init (pimu_initiator, src/Odom/IMUInitialization.cpp:475): VertexGThetaXYRwI seeded from ex->gw,
//     EdgeNavStatePRVG on every pair, the prior-bias edge on states[0]; ex->gw returns RwI * GI (:1262-1275).
int vieo_global_ba_prv_ex(vieo_ba_t* h, const VieoBaProblem* pb_in, const VieoCamera* cam, int n_iterations, int robust,
                          VieoGbaExtra* ex, const volatile uint8_t* stop, VieoNavState* states_out, double* points_out,
                          double* edge_chi2, VieoBaResult* res) {
  VIEO_ARG(h && pb_in && cam && res && states_out, "null argument");
  VIEO_ARG(!pb_in->visual_only, "the global BA of the visual-only system is not implemented");
  VIEO_ARG(h->big, "vieo_global_ba_prv needs a handle from vieo_ba_create_global");
  const bool scale_opt = ex && ex->scale_opt, imu_init = ex && ex->imu_init;
  memset(res, 0, sizeof(*res));
  memcpy(states_out, pb_in->states, sizeof(VieoNavState) * pb_in->n_states);
  if (points_out && pb_in->n_points) memcpy(points_out, pb_in->points, 24 * (size_t)pb_in->n_points);
  VieoBaProblem pb = *pb_in;
  pb.global_ba = 1 | (robust ? 2 : 0) | (scale_opt ? 4 : 0) | (imu_init ? 8 | 16 : 0);
  pb.large = 0; pb.rec_init = 0;
  pb.scale_init = 1.0;
  if (ex) ex->scale = 1.0;
  if (imu_init) memcpy(pb.gw, ex->gw, 24);
  int rc = vieo_ba_set_problem(h, &pb, cam);
  if (rc) return rc;
  if (h->np == 0) return 0;  // bdimPoses == false (:1249)
  double chi = 0;
  if ((rc = vieo_ba_active_robust_chi2(h, 1, &chi))) return rc;
  res->err0 = chi;
  int it = 0;
  bool stopped = false;
  if ((rc = ba_agree_stop(h, stop, &stopped))) return rc;
  if (!stopped) {
    it = vieo_ba_optimize(h, n_iterations, 0.0, stop);
    if (it < 0) return it;
  }
  res->iterations[0] = it;
  if ((rc = vieo_ba_active_robust_chi2(h, 1, &chi))) return rc;
  res->err_end = chi;
  res->lambda_final = h->h_prm->lambda;
  res->accepted = 1;
  if ((rc = vieo_ba_get(h, states_out, points_out, edge_chi2))) return rc;
  if (ex) {
    if ((rc = vieo_ba_get_border(h, &ex->scale, imu_init ? ex->gw : nullptr))) return rc;
    if (scale_opt && points_out)
      for (size_t k = 0; k < 3 * (size_t)pb.n_points; ++k) points_out[k] = ex->scale * points_out[k];
  }
  return it;
}
int vieo_global_ba_prv(vieo_ba_t* h, const VieoBaProblem* pb_in, const VieoCamera* cam, int n_iterations, int robust,
                       const volatile uint8_t* stop, VieoNavState* states_out, double* points_out, double* edge_chi2,
                       VieoBaResult* res) {
  return vieo_global_ba_prv_ex(h, pb_in, cam, n_iterations, robust, nullptr, stop, states_out, points_out, edge_chi2, res);
}

int vieo_local_ba_prv(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, const volatile uint8_t* stop,
                      VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase, VieoBaResult* res);
// ---- asynchronous form: begin enqueues the WHOLE routine on the handle's stream (problem upload, Chi2LargeSetLevel, both
// optimize() stages as device-side LM loops, the re-classification between them, the outlier pass, the downloads into
// pinned staging) and returns; end waits, forwards the abort flag while it waits, and applies the reference's accept /
// reject policy.  One host thread can keep many windows in flight (vieo_local_ba_prv_batch).
static void lba_iteration_plan(const VieoBaProblem* pb, int optit[2], double* lambda0) {
  if (pb->visual_only) { optit[0] = 5; optit[1] = 10; *lambda0 = 0; }
  else if (pb->large) { optit[0] = 2; optit[1] = 2; *lambda0 = 1e-2; }
  else { optit[0] = 4; optit[1] = 6; *lambda0 = 1e0; }
}

static int lba_enqueue_stage(vieo_ba* h, int stage, int iterations, double lambda0) {
  k_ba_begin<<<1, 1, 0, h->st>>>(h->B, stage, iterations, lambda0);
  BA_CK(cudaMemsetAsync(h->B.x, 0, 8 * (size_t)h->np, h->st));
  h->launches++;
  int rc = ba_campose(h);
  if (rc) return rc;
  if ((rc = ba_enqueue_linearize(h, 0, false))) return rc;
  k_ba_control<<<1, 256, 0, h->st>>>(h->B, 0, 0);
  h->launches++;
  BA_CK(cudaGraphLaunch(h->opt_graph, h->st));
  h->launches += kNodesPerTrial;
  return VIEO_OK;
}

int vieo_local_ba_prv_begin(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, const volatile uint8_t* stop) {
  VIEO_ARG(h && pb && cam, "null argument");
  VIEO_ARG(h->async_state == 0, "vieo_local_ba_prv_begin: the previous call on this handle was not ended");
  h->async_pb = *pb;
  h->async_stop = stop;
  h->async_states.assign(pb->states, pb->states + pb->n_states);
  h->async_points.assign(pb->points, pb->points + 3 * (size_t)pb->n_points);
  bool anyfree = false;
  for (int k = 0; k < pb->n_states; ++k) anyfree |= !(pb->state_flags[k] & 1);
  if (!anyfree) {  // if (!bdimPoses) return; (:178)
    h->async_state = 1;
    return VIEO_OK;
  }
  if (h->sharded() || !h->opt_graph || h->big) {
    // the exchange decisions of a sharded handle need host round trips: run the synchronous routine now
    h->sync_erase.assign(std::max(pb->n_edges, 1), 0);
    h->sync_chi2.assign(std::max(pb->n_edges, 1), 0.0);
    h->async_state = 0;
    h->sync_rc = vieo_local_ba_prv(h, pb, cam, stop, h->async_states.data(), h->async_points.data(), h->sync_chi2.data(),
                                   h->sync_erase.data(), &h->sync_res);
    h->async_state = 4;
    return h->sync_rc;
  }
  int optit[2];
  double lambda0;
  lba_iteration_plan(pb, optit, &lambda0);
  h->plan_it[0] = optit[0]; h->plan_it[1] = optit[1]; h->plan_lambda0 = lambda0;  // into the device-resident header
  h->defer_sync = true;
  BA_CK(cudaSetDevice(h->device));
  BA_CK(cudaEventRecord(h->ev_begin, h->st));
  int rc = vieo_ba_set_problem(h, pb, cam);
  h->defer_sync = false;
  if (rc) return rc;
  if (stop && *stop) {  // "Aborted OLBA" (:524-528)
    BA_CK(cudaStreamSynchronize(h->st));
    h->async_state = 2;
    return VIEO_OK;
  }
  const int K = h->K, P = h->P, E = h->E;
  if (h->win_graph && !h->upload_direct) {
    // ONE launch: Chi2LargeSetLevel, err, both optimize() stages (WHILE nodes), the re-classification between them, err_end,
    // the outlier pass and the (capacity-sized) downloads
    size_t total;
    lba_dl_layout((size_t)h->capK, (size_t)h->capP, (size_t)h->capE, h->dl_off, &total);
    BA_CK(cudaGraphLaunch(h->win_graph, h->st));
    h->launches += h->win_nodes + 2 * kNodesPerTrial;
    BA_CK(cudaEventRecord(h->ev_end, h->st));
    h->async_state = 3;
    return VIEO_OK;
  }
  if (!pb->visual_only && (rc = vieo_ba_chi2_large_set_level(h, 100.f))) return rc;  // PRV version only (:534-536)
  if ((rc = ba_errors(h, 0, h->d_ctl + 6))) return rc;                               // err (:539)
  if ((rc = lba_enqueue_stage(h, 0, optit[0], lambda0))) return rc;
  if (E > 0) {  // inlier re-classification + kernel removal (:597-633); skipped on the device once the abort flag is up
    if ((rc = ba_campose(h))) return rc;
    k_ba_classify<<<(E + 255) / 256, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_flags, h->B.chi2, E,
                                                     1, 0.f, 1, 1, h->visual_only ? 0 : 1, h->d_lvl, h->d_bad, h->B.prm, 1);
    h->launches++;
  }
  if ((rc = lba_enqueue_stage(h, 1, optit[1], lambda0))) return rc;
  if ((rc = ba_errors(h, 2, h->d_ctl + 7))) return rc;  // err_end over the stored errors (:652)
  if (E > 0) {  // outlier candidates (:668-700) — computed always, used only if the result is accepted
    if ((rc = ba_campose(h))) return rc;
    k_ba_classify<<<(E + 255) / 256, 256, 0, h->st>>>(h->cam, h->B.cp, h->B.X, h->d_es, h->d_ep, h->d_obs, h->d_flags, h->B.chi2, E,
                                                     1, 0.f, 0, 0, h->visual_only ? 0 : 1, h->d_lvl, h->d_bad, h->B.prm, 0);
    h->launches++;
  }
  // downloads into the pinned staging buffer (the uploads that used it are ahead of them on the stream)
  const size_t ns = sizeof(VieoNavState) * (size_t)K, nx = 24 * (size_t)P, nc = 8 * (size_t)E, nb = (size_t)E;
  size_t dl_total;
  lba_dl_layout((size_t)K, (size_t)P, (size_t)E, h->dl_off, &dl_total);
  uint8_t* ps = h->h_stage + h->dl_off[0];
  uint8_t* px = h->h_stage + h->dl_off[1];
  uint8_t* pc = h->h_stage + h->dl_off[2];
  uint8_t* pbad = h->h_stage + h->dl_off[3];
  VIEO_ARG(dl_total <= h->stage_cap, "staging buffer too small for the asynchronous download");
  BA_CK(cudaMemcpyAsync(ps, h->B.st, ns, cudaMemcpyDeviceToHost, h->st));
  if (nx) BA_CK(cudaMemcpyAsync(px, h->B.X, nx, cudaMemcpyDeviceToHost, h->st));
  if (nc) BA_CK(cudaMemcpyAsync(pc, h->B.chi2, nc, cudaMemcpyDeviceToHost, h->st));
  if (nb) BA_CK(cudaMemcpyAsync(pbad, h->d_bad, nb, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaMemcpyAsync(h->h_ctl, h->d_ctl, 8 * 16, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaMemcpyAsync(h->h_prm, h->B.prm, sizeof(BaParams), cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaEventRecord(h->ev_end, h->st));
  BA_CK(cudaGetLastError());
  h->async_state = 3;
  return VIEO_OK;
}

// 1 when vieo_local_ba_prv_end would not block (or nothing is in flight); forwards the caller's abort flag to the device
int vieo_local_ba_prv_poll(vieo_ba_t* h) {
  if (!h || h->async_state != 3) return 1;
  cudaSetDevice(h->device);
  if (cudaStreamQuery(h->st) != cudaErrorNotReady) return 1;
  if (h->async_stop && *h->async_stop) {
    static const int one = 1;
    cudaMemcpyAsync(&h->B.prm->stop, &one, sizeof(int), cudaMemcpyHostToDevice, h->st_ctl);
    h->async_stop = nullptr;  // sent once
  }
  return 0;
}

int vieo_local_ba_prv_end(vieo_ba_t* h, VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase,
                          VieoBaResult* res) {
  VIEO_ARG(h && res && states_out && erase, "null argument");
  VIEO_ARG(h->async_state != 0, "vieo_local_ba_prv_end without a begin");
  const VieoBaProblem& pb = h->async_pb;
  const int state = h->async_state;
  h->async_state = 0;
  memset(res, 0, sizeof(*res));
  memcpy(states_out, h->async_states.data(), sizeof(VieoNavState) * pb.n_states);
  if (points_out && pb.n_points) memcpy(points_out, h->async_points.data(), 24 * (size_t)pb.n_points);
  memset(erase, 0, pb.n_edges);
  if (state == 1 || state == 2) return VIEO_OK;
  if (state == 4) {
    *res = h->sync_res;
    if (edge_chi2 && pb.n_edges) memcpy(edge_chi2, h->sync_chi2.data(), 8 * (size_t)pb.n_edges);
    memcpy(erase, h->sync_erase.data(), pb.n_edges);
    return h->sync_rc;
  }
  BA_CK(cudaSetDevice(h->device));
  h->async_state = 3;
  while (!vieo_local_ba_prv_poll(h)) std::this_thread::sleep_for(std::chrono::microseconds(20));
  h->async_state = 0;
  BA_CK(cudaStreamSynchronize(h->st));
  cudaEventElapsedTime(&h->last_ms, h->ev_begin, h->ev_end);
  const int K = h->K, P = h->P, E = h->E;
  const size_t ns = sizeof(VieoNavState) * (size_t)K, nx = 24 * (size_t)P, nc = 8 * (size_t)E;
  const uint8_t* ps = h->h_stage + h->dl_off[0];
  const uint8_t* px = h->h_stage + h->dl_off[1];
  const uint8_t* pc = h->h_stage + h->dl_off[2];
  const uint8_t* pbad = h->h_stage + h->dl_off[3];
  const BaParams& q = *h->h_prm;
  h->launches += kNodesPerTrial * std::max(q.trials - 1, 0);
  const float err = (float)h->h_ctl[6], err_end = (float)h->h_ctl[7];
  res->err0 = err;
  res->err_end = err_end;
  res->iterations[0] = q.iters_hist[0];
  res->iterations[1] = q.iters_done;
  res->lambda_final = q.lambda;
  if (edge_chi2 && nc) memcpy(edge_chi2, pc, nc);
  if (!pb.visual_only && (2 * err < err_end || std::isnan(err) || std::isnan(err_end)) && !pb.large) {
    res->accepted = 0;  // "FAIL LOCAL-INERTIAL BA" (:663-666)
    return VIEO_OK;
  }
  res->accepted = 1;
  int ne = 0;
  for (int i = 0; i < E; ++i) {
    erase[i] = pbad[i];
    ne += pbad[i];
  }
  res->n_erase = ne;
  memcpy(states_out, ps, ns);
  if (points_out && nx) memcpy(points_out, px, nx);
  return VIEO_OK;
}

// n independent windows, one handle each, driven by the calling thread: everything is enqueued before anything is awaited
int vieo_local_ba_prv_batch(vieo_ba_t* const* hs, const VieoBaProblem* const* pbs, const VieoCamera* cam, int n,
                            const volatile uint8_t* stop, VieoNavState* const* states_out, double* const* points_out,
                            double* const* edge_chi2, uint8_t* const* erase, VieoBaResult* res) {
  VIEO_ARG(hs && pbs && cam && n >= 0 && states_out && erase && res, "null argument");
  int first = VIEO_OK, begun = 0;
  for (; begun < n; ++begun)
    if ((first = vieo_local_ba_prv_begin(hs[begun], pbs[begun], cam, stop))) break;
  for (int i = 0; i < begun; ++i) {
    const int rc = vieo_local_ba_prv_end(hs[i], states_out[i], points_out ? points_out[i] : nullptr,
                                         edge_chi2 ? edge_chi2[i] : nullptr, erase[i], &res[i]);
    if (rc && !first) first = rc;
  }
  return first;
}

// Optimizer::LocalBundleAdjustmentNavStatePRV, src/Optimizer.cc:133-700, on the flattened problem
int vieo_local_ba_prv(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, const volatile uint8_t* stop,
                      VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase, VieoBaResult* res) {
  VIEO_ARG(h && pb && cam && res && states_out && erase, "null argument");
  if (!h->sharded() && h->opt_graph && !h->big && h->async_state == 0) {
    // single GPU: the whole routine is enqueued at once (device-side LM loops) and awaited with ONE synchronisation
    int rc = vieo_local_ba_prv_begin(h, pb, cam, stop);
    if (rc) {
      h->async_state = 0;
      return rc;
    }
    return vieo_local_ba_prv_end(h, states_out, points_out, edge_chi2, erase, res);
  }
  memset(res, 0, sizeof(*res));
  memcpy(states_out, pb->states, sizeof(VieoNavState) * pb->n_states);
  if (points_out && pb->n_points) memcpy(points_out, pb->points, 24 * (size_t)pb->n_points);
  memset(erase, 0, pb->n_edges);
  int optit[2];
  double lambda0;
  lba_iteration_plan(pb, optit, &lambda0);
  bool anyfree = false;
  for (int k = 0; k < pb->n_states; ++k) anyfree |= !(pb->state_flags[k] & 1);
  if (!anyfree) return VIEO_OK;  // if (!bdimPoses) return; (:178)
  int rc = vieo_ba_set_problem(h, pb, cam);
  if (rc) return rc;
  bool stopped = false;
  if ((rc = ba_agree_stop(h, stop, &stopped))) return rc;
  if (stopped) return VIEO_OK;  // "Aborted OLBA" (:524-528)
  if (!pb->visual_only && (rc = vieo_ba_chi2_large_set_level(h, 100.f))) return rc;  // PRV version only (:534-536)
  double chi = 0;
  if ((rc = vieo_ba_active_robust_chi2(h, 1, &chi))) return rc;
  const float err = (float)chi;
  res->err0 = err;
  int n = vieo_ba_optimize(h, optit[0], lambda0, stop);
  if (n < 0) return n;
  res->iterations[0] = n;
  bool bDoMore = true;
  if ((rc = ba_agree_stop(h, stop, &stopped))) return rc;
  if (stopped) bDoMore = false;
  if (bDoMore) {
    if ((rc = vieo_ba_reclassify(h, 1, nullptr))) return rc;
    n = vieo_ba_optimize(h, optit[1], lambda0, stop);
    if (n < 0) return n;
    res->iterations[1] = n;
  }
  if ((rc = vieo_ba_active_robust_chi2(h, 0, &chi))) return rc;
  const float err_end = (float)chi;
  res->err_end = err_end;
  res->lambda_final = h->h_prm->lambda;
  if ((rc = vieo_ba_get(h, nullptr, nullptr, edge_chi2))) return rc;
  if (!pb->visual_only && (2 * err < err_end || std::isnan(err) || std::isnan(err_end)) && !pb->large) {
    res->accepted = 0;  // "FAIL LOCAL-INERTIAL BA" (:663-666)
    return VIEO_OK;
  }
  res->accepted = 1;
  if ((rc = vieo_ba_reclassify(h, 0, erase))) return rc;
  int ne = 0;
  for (int i = 0; i < pb->n_edges; ++i) ne += erase[i];
  res->n_erase = ne;
  return vieo_ba_get(h, states_out, points_out, nullptr);
}

}  // extern "C"

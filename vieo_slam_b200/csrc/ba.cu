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
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "ba_edges.cuh"

namespace vieo {

constexpr int kBaWarps = 8;

struct BaDense {  // one inertial (EdgeNavStatePRV) or bias random-walk (EdgeNavStateBias) factor
  int type;       // 0 IMU, 1 bias
  int si, sj, pre;
  double delta;     // Huber delta, 0 = none
  double info[81];  // IMU: 9x9 information; bias: [0..6) diagonal
};
struct BaDenseWork {
  double J[9 * 24];  // IMU: [Ji(PR) | Jj(PR) | Ji(V) | Jj(V) | Jb], 9 x 24 row-major
  double Om[81], oe[9], err[9];
  double chi2, r1;
};

__global__ void k_ba_campose(CamK cam, const VieoNavState* __restrict__ st, int K, CamPose* __restrict__ cp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) cp[k] = cam_pose(cam, ns_load(st[k]));
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
                                                   double* __restrict__ partial) {
  __shared__ double s_w[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double r0 = 0;
  if (i < E) {
    const bool active = !(lvl[i] & 1) && (points_free || !sfix[es[i]]);
    if (active || all) {
      const bool stereo = flags[i] & VIEO_EDGE_STEREO;
      double e[3];
      reproj_error(cam, cp[es[i]], ld3(X + 3 * (size_t)ep[i]), obs + 3 * (size_t)i, stereo, e);
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

// errors of the inertial / bias edges (one thread each) and the total robust chi2 (dense first, then the visual
// partial sums in block order) -> *out.  Also sums the landmark part of the gain-ratio denominator when asked.
__global__ void __launch_bounds__(128) k_ba_dense_errors(const BaDense* __restrict__ den, int n_den, const VieoNavState* __restrict__ st,
                                  const VieoImuPreint* __restrict__ pre, Vec3 gw, BaDenseWork* __restrict__ wk,
                                  const double* __restrict__ partial, int n_partial, double* __restrict__ out,
                                  const double* __restrict__ scale_part, int n_scale, double* __restrict__ scale_out) {
  for (int m = threadIdx.x; m < n_den; m += blockDim.x) {
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
    W.oe[8] = r0;  // parked until the sum below (oe is rewritten by k_ba_dense_build)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int k = 0; k < n_den; ++k) tot += wk[k].oe[8];
    for (int k = 0; k < n_partial; ++k) tot += partial[k];
    *out = tot;
    if (scale_out) {
      double s = 0;
      for (int k = 0; k < n_scale; ++k) s += scale_part[k];
      *scale_out = s;
    }
  }
}

// One warp per map point: linearise its edges, reduce Hll / bl by shuffles, store W (Hpl) and A (Hpp part) per edge.
__global__ void __launch_bounds__(kBaWarps * 32) k_ba_linearize(
    CamK cam, const CamPose* __restrict__ cp, const double* __restrict__ X, const int* __restrict__ pt_ptr, int P,
    const int* __restrict__ es, const float* __restrict__ obs, const float* __restrict__ w,
    const uint8_t* __restrict__ flags, const uint8_t* __restrict__ lvl, const uint8_t* __restrict__ sfix,
    const double* __restrict__ chi2, double dm, double ds, double* __restrict__ Wb, double* __restrict__ Ab,
    double* __restrict__ Hll, double* __restrict__ bl, uint8_t* __restrict__ pt_active) {
  const int p = blockIdx.x * kBaWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  const int i0 = pt_ptr[p], i1 = pt_ptr[p + 1];
  const Vec3 Xp = ld3(X + 3 * (size_t)p);
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0;
  bool any = false;
  for (int i = i0 + lane; i < i1; i += 32) {
    double* Wi = Wb + 18 * (size_t)i;
    double* Ai = Ab + 27 * (size_t)i;
    if (lvl[i] & 1) {
      for (int k = 0; k < 18; ++k) Wi[k] = 0;
      for (int k = 0; k < 27; ++k) Ai[k] = 0;
      continue;
    }
    any = true;
    const bool stereo = flags[i] & VIEO_EDGE_STEREO;
    const int DE = stereo ? 3 : 2;
    const int s = es[i];
    double e[3];
    reproj_error(cam, cp[s], Xp, obs + 3 * (size_t)i, stereo, e);
    Mat3 Jp, Jr, JX;
    reproj_jac(cam, cp[s], Xp, stereo, Jp, Jr, JX);
    double r0, r1;
    huber_rho(edge_huber_delta(flags[i], lvl[i], dm, ds), chi2[i], r0, r1);
    const double wi = (double)w[i], ww = r1 * wi;
    double oe[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) oe[k] = -(wi * e[k]) * r1;
    double J[3][6];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J[k][c] = Jp.m[3 * k + c];
        J[k][3 + c] = Jr.m[3 * k + c];
      }
    // point block: Hll (upper 6) and bl
    int q = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double sb = 0;
      for (int k = 0; k < DE; ++k) sb += JX.m[3 * k + a] * oe[k];
      acc[6 + a] += sb;
#pragma unroll
      for (int c = a; c < 3; ++c) {
        double h = 0;
        for (int k = 0; k < DE; ++k) h += (JX.m[3 * k + a] * ww) * JX.m[3 * k + c];
        acc[q++] += h;
      }
    }
    if (!sfix[s]) {
      q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double sb = 0;
        for (int k = 0; k < DE; ++k) sb += J[k][a] * oe[k];
        Ai[21 + a] = sb;
#pragma unroll
        for (int c = a; c < 6; ++c) {
          double h = 0;
          for (int k = 0; k < DE; ++k) h += (J[k][a] * ww) * J[k][c];
          Ai[q++] = h;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double h = 0;
          for (int k = 0; k < DE; ++k) h += (J[k][a] * ww) * JX.m[3 * k + c];
          Wi[3 * a + c] = h;
        }
      }
    } else {
      for (int k = 0; k < 18; ++k) Wi[k] = 0;
      for (int k = 0; k < 27; ++k) Ai[k] = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) {
    double* H = Hll + 9 * (size_t)p;
    H[0] = acc[0]; H[1] = acc[1]; H[2] = acc[2];
    H[3] = acc[1]; H[4] = acc[3]; H[5] = acc[4];
    H[6] = acc[2]; H[7] = acc[4]; H[8] = acc[5];
    bl[3 * (size_t)p] = acc[6]; bl[3 * (size_t)p + 1] = acc[7]; bl[3 * (size_t)p + 2] = acc[8];
    pt_active[p] = any;
  }
}

// One block (128 threads) per free keyframe: fixed-order sum of its edges' A blocks -> 6x6 diagonal block + rhs
__global__ void __launch_bounds__(128) k_ba_pose_reduce(const int* __restrict__ free_state, const int* __restrict__ off0,
                                                        const int* __restrict__ ps_ptr, const int* __restrict__ ps_edges,
                                                        const double* __restrict__ Ab, int np, double* __restrict__ H,
                                                        double* __restrict__ b) {
  __shared__ double s_w[4][27];
  const int f = blockIdx.x, k = free_state[f], o = off0[k];
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0;
  for (int t = ps_ptr[f] + threadIdx.x; t < ps_ptr[f + 1]; t += 128) {
    const double* Ai = Ab + 27 * (size_t)ps_edges[t];
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
    const double v = ((s_w[0][threadIdx.x] + s_w[1][threadIdx.x]) + s_w[2][threadIdx.x]) + s_w[3][threadIdx.x];
    s_w[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int q = 0;
    for (int a = 0; a < 6; ++a) {
      b[o + a] = s_w[0][21 + a];
      for (int c = a; c < 6; ++c) {
        const double h = s_w[0][q++];
        H[(size_t)(o + a) * np + o + c] = h;
        H[(size_t)(o + c) * np + o + a] = h;
      }
    }
  }
}

// Inertial and bias edges into Hpp / b.  Phase 1: one thread per edge builds its Jacobian strip and weighted
// information; phase 2: the block walks the edges in order and adds J^T (rho' Omega) J to the mapped positions.
__global__ void __launch_bounds__(256) k_ba_dense_build(const BaDense* __restrict__ den, int n_den,
                                                        const VieoNavState* __restrict__ st,
                                                        const VieoImuPreint* __restrict__ pre, Vec3 gw,
                                                        BaDenseWork* __restrict__ wk, const int* __restrict__ off0,
                                                        const int* __restrict__ off1, const int* __restrict__ off2, int np,
                                                        double* __restrict__ H, double* __restrict__ b) {
  for (int m = threadIdx.x; m < n_den; m += 256) {
    const BaDense& d = den[m];
    if (d.type != 0) continue;
    BaDenseWork& W = wk[m];
    const NavS a = ns_load(st[d.si]), c = ns_load(st[d.sj]);
    double Ji[81], Jj[81], Jb[54];
    navstate_jac(a, c, pre[d.pre], gw, true, W.err, Ji, Jj, Jb);
    for (int i = 0; i < 9; ++i) {
      for (int k = 0; k < 6; ++k) {
        W.J[i * 24 + k] = Ji[i * 9 + k];
        W.J[i * 24 + 6 + k] = Jj[i * 9 + k];
        W.J[i * 24 + 18 + k] = Jb[i * 6 + k];
      }
      for (int k = 0; k < 3; ++k) {
        W.J[i * 24 + 12 + k] = Ji[i * 9 + 6 + k];
        W.J[i * 24 + 15 + k] = Jj[i * 9 + 6 + k];
      }
    }
    for (int i = 0; i < 81; ++i) W.Om[i] = W.r1 * d.info[i];
    for (int i = 0; i < 9; ++i) {
      double s = 0;
      for (int j = 0; j < 9; ++j) s += d.info[i * 9 + j] * W.err[j];
      W.oe[i] = -s * W.r1;
    }
  }
  __syncthreads();
  __threadfence_block();
  for (int m = 0; m < n_den; ++m) {
    const BaDense& d = den[m];
    const BaDenseWork& W = wk[m];
    if (d.type == 0) {
      const int offs[5] = {off0[d.si], off0[d.sj], off1[d.si], off1[d.sj], off2[d.si]};
      const int base[6] = {0, 6, 12, 15, 18, 24};
      auto gcol = [&](int lc) {
        int blk = lc < 6 ? 0 : lc < 12 ? 1 : lc < 15 ? 2 : lc < 18 ? 3 : 4;
        return offs[blk] < 0 ? -1 : offs[blk] + (lc - base[blk]);
      };
      for (int t = threadIdx.x; t < 24 * 25; t += 256) {
        const int r = t / 25, c = t % 25;
        const int gr = gcol(r);
        if (gr < 0) continue;
        if (c == 24) {
          double s = 0;
          for (int i = 0; i < 9; ++i) s += W.J[i * 24 + r] * W.oe[i];
          b[gr] += s;
          continue;
        }
        const int gc = gcol(c);
        if (gc < 0) continue;
        double s = 0;
        for (int j = 0; j < 9; ++j) {
          double a = 0;
          for (int i = 0; i < 9; ++i) a += W.J[i * 24 + r] * W.Om[i * 9 + j];
          s += a * W.J[j * 24 + c];
        }
        H[(size_t)gr * np + gc] += s;
      }
    } else {
      const int oi = off2[d.si], oj = off2[d.sj];
      if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        const double om = W.r1 * d.info[k], oe = -(d.info[k] * W.err[k]) * W.r1;
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
    __syncthreads();
  }
}

// max |diag| of Hpp and of the active Hll (computeLambdaInit)
__global__ void k_ba_maxdiag(const double* __restrict__ H, int np, const double* __restrict__ Hll,
                             const uint8_t* __restrict__ pt_active, int P, double* __restrict__ out) {
  __shared__ double s[256];
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
  if (threadIdx.x == 0) *out = s[0];
}

// Dinv = (Hll + lambda I)^-1 (cofactor inverse like Eigen's fixed 3x3, block_solver.hpp:389), db = Dinv bl
__global__ void k_ba_point_inv(const double* __restrict__ Hll, const double* __restrict__ bl,
                               const uint8_t* __restrict__ pt_active, int P, double lambda, double* __restrict__ Dinv,
                               double* __restrict__ db) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double* I = Dinv + 9 * (size_t)p;
  if (!pt_active[p]) {
    for (int k = 0; k < 9; ++k) I[k] = 0;
    db[3 * (size_t)p] = db[3 * (size_t)p + 1] = db[3 * (size_t)p + 2] = 0;
    return;
  }
  double D[9];
  for (int k = 0; k < 9; ++k) D[k] = Hll[9 * (size_t)p + k];
  D[0] += lambda; D[4] += lambda; D[8] += lambda;
  const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
  const double det = D[0] * c00 + D[1] * c01 + D[2] * c02, id = 1.0 / det;
  I[0] = c00 * id; I[1] = (D[2] * D[7] - D[1] * D[8]) * id; I[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  I[3] = c01 * id; I[4] = (D[0] * D[8] - D[2] * D[6]) * id; I[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  I[6] = c02 * id; I[7] = (D[1] * D[6] - D[0] * D[7]) * id; I[8] = (D[0] * D[4] - D[1] * D[3]) * id;
  const double* bb = bl + 3 * (size_t)p;
  for (int a = 0; a < 3; ++a) db[3 * (size_t)p + a] = I[3 * a] * bb[0] + I[3 * a + 1] * bb[1] + I[3 * a + 2] * bb[2];
}

// S = H + lambda I, bschur = b
__global__ void k_ba_copy_sys(const double* __restrict__ H, const double* __restrict__ b, int np, double lambda,
                              double* __restrict__ S, double* __restrict__ bs) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < (size_t)np * np) {
    const int r = t / np, c = t % np;
    S[t] = H[t] + (r == c ? lambda : 0.0);
  }
  if (t < (size_t)np) bs[t] = b[t];
}

// Schur complement, one block per free keyframe (6 rows of S): for every edge a of the keyframe (fixed chunks per
// warp, in list order) and every edge c of a's point: acc[row][6*prcol(c) + col] += (W_a Dinv) W_c^T; the warps'
// accumulators are summed in warp order and subtracted from S.  bschur -= W_a db.
__global__ void __launch_bounds__(kBaWarps * 32) k_ba_schur(
    const int* __restrict__ free_state, const int* __restrict__ off0, const int* __restrict__ prcol,
    const int* __restrict__ free_off, int nfree, const int* __restrict__ ps_ptr, const int* __restrict__ ps_edges,
    const int* __restrict__ es, const int* __restrict__ ep, const int* __restrict__ pt_ptr, const double* __restrict__ Wb,
    const double* __restrict__ Dinv, const double* __restrict__ db, int serial_lanes, int np, double* __restrict__ S,
    double* __restrict__ bs) {
  extern __shared__ double s_acc[];  // [kBaWarps][6][6 * nfree + 1]
  const int ld = 6 * nfree + 1;      // last column: W_a db
  const int f = blockIdx.x, o = off0[free_state[f]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* acc = s_acc + (size_t)warp * 6 * ld;
  for (int t = lane; t < 6 * ld; t += 32) acc[t] = 0;
  __syncwarp();
  const int t0 = ps_ptr[f], t1 = ps_ptr[f + 1];
  const int chunk = (t1 - t0 + kBaWarps - 1) / kBaWarps;
  const int a0 = t0 + warp * chunk, a1 = min(a0 + chunk, t1);
  for (int t = a0; t < a1; ++t) {
    const int a = ps_edges[t];
    const int p = ep[a];
    const double* Wa = Wb + 18 * (size_t)a;
    const double* Di = Dinv + 9 * (size_t)p;
    double WD[18];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) WD[3 * r + c] = Wa[3 * r] * Di[c] + Wa[3 * r + 1] * Di[3 + c] + Wa[3 * r + 2] * Di[6 + c];
    if (lane < 6) {
      const double* d = db + 3 * (size_t)p;
      acc[lane * ld + 6 * nfree] += Wa[3 * lane] * d[0] + Wa[3 * lane + 1] * d[1] + Wa[3 * lane + 2] * d[2];
    }
    const int c0 = pt_ptr[p], c1 = pt_ptr[p + 1];
    for (int cb = c0; cb < c1; cb += 32) {
      const int c = cb + lane;
      const int col = c < c1 ? prcol[es[c]] : -1;
      for (int turn = 0; turn < (serial_lanes ? 32 : 1); ++turn) {
        if (col >= 0 && (!serial_lanes || turn == lane)) {
          const double* Wc = Wb + 18 * (size_t)c;
          double* dst = acc + 6 * col;
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = 0; cc < 6; ++cc)
              dst[r * ld + cc] += WD[3 * r] * Wc[3 * cc] + WD[3 * r + 1] * Wc[3 * cc + 1] + WD[3 * r + 2] * Wc[3 * cc + 2];
        }
        if (serial_lanes) __syncwarp();
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 6 * ld; t += kBaWarps * 32) {
    double s = 0;
    for (int w = 0; w < kBaWarps; ++w) s += s_acc[(size_t)w * 6 * ld + t];
    const int r = t / ld, c = t % ld;
    if (c == 6 * nfree) bs[o + r] -= s;
    else S[(size_t)(o + r) * np + free_off[c / 6] + c % 6] -= s;
  }
}

// Dense Cholesky (lower, in place) of S (n x n, row-major, global memory) and solve S x = bs by one block.
// ok = 0 when a pivot is not positive (LinearSolverDense: LDLT::isPositive, linear_solver_dense.h:107-112).
__global__ void __launch_bounds__(1024) k_ba_chol(double* __restrict__ A, const double* __restrict__ rhs, int n,
                                                  double* __restrict__ x, double* __restrict__ y, int* __restrict__ ok) {
  __shared__ double s_d;
  __shared__ int s_ok;
  const int T = blockDim.x, t = threadIdx.x;
  if (t == 0) s_ok = 1;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (t == 0) {
      const double d = A[(size_t)j * n + j];
      if (!(d > 0) || !isfinite(d)) s_ok = 0;
      s_d = sqrt(d);
      A[(size_t)j * n + j] = s_d;
    }
    __syncthreads();
    if (!s_ok) break;
    const double d = s_d;
    for (int i = j + 1 + t; i < n; i += T) A[(size_t)i * n + j] /= d;
    __syncthreads();
    // trailing update of the lower triangle: A[i][k] -= A[i][j] A[k][j], j < k <= i
    const int m = n - j - 1;
    for (int e = t; e < m * m; e += T) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) A[(size_t)i * n + k] -= A[(size_t)i * n + j] * A[(size_t)k * n + j];
    }
    __syncthreads();
  }
  if (t == 0) *ok = s_ok;
  if (!s_ok) return;
  // forward / backward substitution, one warp: lanes split the dot products
  if (t < 32) {
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int k = t; k < i; k += 32) s += A[(size_t)i * n + k] * y[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (t == 0) y[i] = (rhs[i] - s) / A[(size_t)i * n + i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = 0;
      for (int k = i + 1 + t; k < n; k += 32) s += A[(size_t)k * n + i] * x[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (t == 0) x[i] = (y[i] - s) / A[(size_t)i * n + i];
      __syncwarp();
    }
  }
}

// One warp per point: xl = Dinv (bl - sum_a W_a^T xp), X += xl (apply != 0), landmark part of computeScale
__global__ void __launch_bounds__(kBaWarps * 32) k_ba_backsub(const int* __restrict__ pt_ptr, int P,
                                                             const int* __restrict__ es, const int* __restrict__ off0,
                                                             const double* __restrict__ Wb, const double* __restrict__ Dinv,
                                                             const double* __restrict__ bl,
                                                             const uint8_t* __restrict__ pt_active,
                                                             const double* __restrict__ x, const int* __restrict__ ok,
                                                             double lambda, int apply, double* __restrict__ X,
                                                             double* __restrict__ xl_out, double* __restrict__ scale_part) {
  const int p = blockIdx.x * kBaWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  if (!pt_active[p] || !*ok) {
    if (lane == 0) {
      scale_part[p] = 0;
      if (xl_out && !*ok) xl_out[3 * (size_t)p] = xl_out[3 * (size_t)p + 1] = xl_out[3 * (size_t)p + 2] = 0;
    }
    if (!pt_active[p] && lane == 0 && xl_out) xl_out[3 * (size_t)p] = xl_out[3 * (size_t)p + 1] = xl_out[3 * (size_t)p + 2] = 0;
    return;
  }
  double c[3] = {0, 0, 0};
  for (int i = pt_ptr[p] + lane; i < pt_ptr[p + 1]; i += 32) {
    const int o = off0[es[i]];
    if (o < 0) continue;
    const double* Wi = Wb + 18 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int r = 0; r < 6; ++r) c[k] += Wi[3 * r + k] * x[o + r];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
  if (lane == 0) {
    const double* bb = bl + 3 * (size_t)p;
    const double* Di = Dinv + 9 * (size_t)p;
    const double cc[3] = {bb[0] - c[0], bb[1] - c[1], bb[2] - c[2]};
    double s = 0;
    for (int a = 0; a < 3; ++a) {
      const double xl = Di[3 * a] * cc[0] + Di[3 * a + 1] * cc[1] + Di[3 * a + 2] * cc[2];
      if (apply) X[3 * (size_t)p + a] += xl;
      if (xl_out) xl_out[3 * (size_t)p + a] = xl;
      s += xl * (lambda * xl + bb[a]);
    }
    scale_part[p] = s;
  }
}

// oplus on the keyframe vertices (NavState::IncSmall) + pose part of computeScale
__global__ void k_ba_update_states(VieoNavState* __restrict__ st, int K, const int* __restrict__ off0,
                                   const int* __restrict__ off1, const int* __restrict__ off2,
                                   const double* __restrict__ x, const double* __restrict__ b, int np, double lambda,
                                   const int* __restrict__ ok, double* __restrict__ scale_pose) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K && *ok) {
    if (off0[k] >= 0 || off1[k] >= 0 || off2[k] >= 0) {
      NavS s = ns_load(st[k]);
      if (off0[k] >= 0) ns_inc_pr(s, x + off0[k]);
      if (off1[k] >= 0) ns_inc_v(s, x + off1[k]);
      if (off2[k] >= 0) ns_inc_bias(s, x + off2[k]);
      ns_store(s, st[k]);
    }
  }
  if (k == 0) {
    double s = 0;
    if (*ok)
      for (int j = 0; j < np; ++j) s += x[j] * (lambda * x[j] + b[j]);
    *scale_pose = s;
  }
}

// level / erase classification.  mode 0: Chi2LargeSetLevel (chi2 > rat * chi2_sig5[dim]); mode 1: the LBA gates
// (src/Optimizer.cc:597-633, 668-700): chi2 > 5.991 (x1.5 when close) / 7.815 or depth <= 0.
__global__ void k_ba_classify(CamK cam, const CamPose* __restrict__ cp, const double* __restrict__ X,
                              const int* __restrict__ es, const int* __restrict__ ep, const float* __restrict__ obs,
                              const uint8_t* __restrict__ flags, const double* __restrict__ chi2, int E, int mode, float rat,
                              int set_level, int remove_kernels, uint8_t* __restrict__ lvl, uint8_t* __restrict__ bad_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  const bool stereo = flags[i] & VIEO_EDGE_STEREO;
  bool bad;
  if (mode == 0) {
    const float th = rat * (stereo ? 7.815f : 5.991f);
    bad = chi2[i] > (double)th;
  } else {
    double e[3];
    const double depth = reproj_error(cam, cp[es[i]], ld3(X + 3 * (size_t)ep[i]), obs + 3 * (size_t)i, stereo, e);
    const float chi2Mono = 5.991f;
    if (stereo) bad = chi2[i] > 7.815 || !(depth > 0.);
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
  int capK = 0, capP = 0, capE = 0, capM = 0;
  int K = 0, P = 0, E = 0, M = 0, np = 0, nfree = 0, n_den = 0, n_part = 0;
  bool points_free = true, has_dup = false;
  int rank = 0, world = 1;
  vieo_allreduce_fn allreduce = nullptr;
  void* ar_ctx = nullptr;
  CamK cam;
  Vec3 gw;
  double dm = 0, ds = 0;
  // device
  VieoNavState *d_st = nullptr, *d_st_bak = nullptr;
  CamPose* d_cp = nullptr;
  double *d_X = nullptr, *d_X_bak = nullptr, *d_chi2 = nullptr, *d_W = nullptr, *d_A = nullptr, *d_Hll = nullptr,
         *d_bl = nullptr, *d_Dinv = nullptr, *d_db = nullptr, *d_H = nullptr, *d_sys = nullptr, *d_x = nullptr,
         *d_y = nullptr, *d_xl = nullptr, *d_partial = nullptr, *d_scale_part = nullptr, *d_ctl = nullptr;
  int *d_es = nullptr, *d_ep = nullptr, *d_pt_ptr = nullptr, *d_off0 = nullptr, *d_off1 = nullptr, *d_off2 = nullptr,
      *d_prcol = nullptr, *d_free_state = nullptr, *d_free_off = nullptr, *d_ps_ptr = nullptr, *d_ps_edges = nullptr,
      *d_ok = nullptr;
  float *d_obs = nullptr, *d_w = nullptr;
  uint8_t *d_flags = nullptr, *d_lvl = nullptr, *d_sfix = nullptr, *d_pt_active = nullptr, *d_bad = nullptr;
  VieoImuPreint* d_pre = nullptr;
  BaDense* d_den = nullptr;
  BaDenseWork* d_wk = nullptr;
  double* h_ctl = nullptr;  // pinned
  // host mirrors
  std::vector<int> off0, off1, off2;
  // LM state (OptimizationAlgorithmLevenberg members)
  double lambda = 0, ni = 2;
  int nBad = 0;
  int launches = 0;
  // sys layout
  double* S() { return d_sys; }
  double* bs() { return d_sys + (size_t)np * np; }
  double* b() { return d_sys + (size_t)np * np + np; }
  double* chi_cur() { return d_sys + (size_t)np * np + 2 * np; }
  size_t sys_count() const { return (size_t)np * np + 2 * np + 1; }
};

namespace {

template <class T>
cudaError_t dalloc(T** p, size_t n) {
  return cudaMalloc((void**)p, sizeof(T) * std::max<size_t>(n, 1));
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

int ba_campose(vieo_ba* h) {
  k_ba_campose<<<(h->K + 127) / 128, 128, 0, h->st>>>(h->cam, h->d_st, h->K, h->d_cp);
  h->launches++;
  return VIEO_OK;
}

// computeActiveErrors (+ every edge when all) and the robust chi2 into *d_out; scale parts summed when asked
int ba_errors(vieo_ba* h, int all, double* d_out, bool with_scale) {
  ba_campose(h);
  if (h->E > 0) {
    k_ba_errors<<<h->n_part, 256, 0, h->st>>>(h->cam, h->d_cp, h->d_X, h->d_es, h->d_ep, h->d_obs, h->d_w, h->d_flags,
                                              h->d_lvl, h->d_sfix, h->points_free ? 1 : 0, h->E, all, h->dm, h->ds,
                                              h->d_chi2, h->d_partial);
    h->launches++;
  }
  k_ba_dense_errors<<<1, 128, 0, h->st>>>(
      h->d_den, h->n_den, h->d_st, h->d_pre, h->gw, h->d_wk, h->d_partial, h->E > 0 ? h->n_part : 0, d_out,
      h->d_scale_part, with_scale ? h->P : 0, with_scale ? h->d_ctl + 2 : nullptr);
  h->launches++;
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

// buildSystem: H (np x np), b, Hll, bl, W
int ba_build(vieo_ba* h) {
  const int np = h->np;
  BA_CK(cudaMemsetAsync(h->d_H, 0, sizeof(double) * std::max<size_t>((size_t)np * np, 1), h->st));
  BA_CK(cudaMemsetAsync(h->b(), 0, sizeof(double) * std::max(np, 1), h->st));
  if (h->P > 0) {
    k_ba_linearize<<<(h->P + kBaWarps - 1) / kBaWarps, kBaWarps * 32, 0, h->st>>>(
        h->cam, h->d_cp, h->d_X, h->d_pt_ptr, h->P, h->d_es, h->d_obs, h->d_w, h->d_flags, h->d_lvl, h->d_sfix, h->d_chi2,
        h->dm, h->ds, h->d_W, h->d_A, h->d_Hll, h->d_bl, h->d_pt_active);
    h->launches++;
  }
  if (h->nfree > 0) {
    k_ba_pose_reduce<<<h->nfree, 128, 0, h->st>>>(h->d_free_state, h->d_off0, h->d_ps_ptr, h->d_ps_edges, h->d_A, np, h->d_H,
                                                  h->b());
    h->launches++;
  }
  if (h->n_den > 0) {
    k_ba_dense_build<<<1, 256, 0, h->st>>>(h->d_den, h->n_den, h->d_st, h->d_pre, h->gw, h->d_wk, h->d_off0, h->d_off1,
                                           h->d_off2, np, h->d_H, h->b());
    h->launches++;
  }
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

// (H + lambda I) x = b through the Schur complement; apply: update the estimate
int ba_solve(vieo_ba* h, double lambda, int apply, double* xl_out) {
  const int np = h->np;
  if (h->P > 0) {
    k_ba_point_inv<<<(h->P + 127) / 128, 128, 0, h->st>>>(h->d_Hll, h->d_bl, h->d_pt_active, h->P, lambda, h->d_Dinv,
                                                          h->d_db);
    h->launches++;
  }
  const size_t nn = std::max<size_t>((size_t)np * np, np);
  // sharded: lambda is added once (rank 0); the all-reduce sums the partial systems
  k_ba_copy_sys<<<(unsigned)((nn + 255) / 256), 256, 0, h->st>>>(h->d_H, h->b(), np, h->rank == 0 ? lambda : 0.0, h->S(),
                                                                h->bs());
  h->launches++;
  if (h->nfree > 0 && h->E > 0) {
    const size_t smem = sizeof(double) * kBaWarps * 6 * (6 * (size_t)h->nfree + 1);
    k_ba_schur<<<h->nfree, kBaWarps * 32, smem, h->st>>>(h->d_free_state, h->d_off0, h->d_prcol, h->d_free_off, h->nfree,
                                                         h->d_ps_ptr, h->d_ps_edges, h->d_es, h->d_ep, h->d_pt_ptr, h->d_W,
                                                         h->d_Dinv, h->d_db, h->has_dup ? 1 : 0, np, h->S(), h->bs());
    h->launches++;
  }
  if (h->allreduce && h->world > 1) {
    int rc = h->allreduce(h->ar_ctx, h->d_sys, h->sys_count(), (void*)h->st);
    if (rc) {
      vieo::set_error("allreduce callback failed (%d)", rc);
      return VIEO_E_CUDA;
    }
  }
  k_ba_chol<<<1, 1024, 0, h->st>>>(h->S(), h->bs(), np, h->d_x, h->d_y, h->d_ok);
  h->launches++;
  if (h->P > 0) {
    k_ba_backsub<<<(h->P + kBaWarps - 1) / kBaWarps, kBaWarps * 32, 0, h->st>>>(
        h->d_pt_ptr, h->P, h->d_es, h->d_off0, h->d_W, h->d_Dinv, h->d_bl, h->d_pt_active, h->d_x, h->d_ok, lambda, apply,
        h->d_X, xl_out, h->d_scale_part);
    h->launches++;
  }
  if (apply) {
    k_ba_update_states<<<(h->K + 127) / 128, 128, 0, h->st>>>(h->d_st, h->K, h->d_off0, h->d_off1, h->d_off2, h->d_x, h->b(),
                                                              np, lambda, h->d_ok, h->d_ctl + 3);
    h->launches++;
  }
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

int ba_read_ctl(vieo_ba* h) {
  // ctl: [0] currentChi (copy of sys tail), [1] tempChi, [2] landmark scale, [3] pose scale, [4] maxdiag
  BA_CK(cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaMemcpyAsync(h->h_ctl + 8, h->chi_cur(), sizeof(double), cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaMemcpyAsync(h->h_ctl + 9, h->d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  return VIEO_OK;
}

// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-166): 0 OK, 1 Terminate, <0 error
int ba_lm_iteration(vieo_ba* h, int iteration, double user_lambda, const volatile uint8_t* stop) {
  int rc;
  if ((rc = ba_errors(h, 0, h->chi_cur(), false))) return rc;
  if ((rc = ba_build(h))) return rc;
  if (iteration == 0) {
    if (user_lambda > 0) h->lambda = user_lambda;
    else {
      k_ba_maxdiag<<<1, 256, 0, h->st>>>(h->d_H, h->np, h->d_Hll, h->d_pt_active, h->P, h->d_ctl + 4);
      h->launches++;
      if (h->allreduce && h->world > 1) {
        vieo::set_error("sharded BA needs an explicit initial lambda");
        return VIEO_E_ARG;
      }
      BA_CK(cudaMemcpyAsync(h->h_ctl + 4, h->d_ctl + 4, sizeof(double), cudaMemcpyDeviceToHost, h->st));
      BA_CK(cudaStreamSynchronize(h->st));
      h->lambda = 1e-5 * h->h_ctl[4];
    }
    h->ni = 2;
    h->nBad = 0;
  }
  double currentChi = 0, iniChi = 0, rho = 0;
  int qmax = 0;
  const bool sharded = h->allreduce && h->world > 1;
  do {
    BA_CK(cudaMemcpyAsync(h->d_st_bak, h->d_st, sizeof(VieoNavState) * h->K, cudaMemcpyDeviceToDevice, h->st));
    if (h->P) BA_CK(cudaMemcpyAsync(h->d_X_bak, h->d_X, sizeof(double) * 3 * h->P, cudaMemcpyDeviceToDevice, h->st));
    if ((rc = ba_solve(h, h->lambda, 1, nullptr))) return rc;
    if ((rc = ba_errors(h, 0, h->d_ctl + 1, true))) return rc;
    if (sharded) {  // [tempChi, landmark scale] partial sums
      if (h->allreduce(h->ar_ctx, h->d_ctl + 1, 2, (void*)h->st)) return VIEO_E_CUDA;
    }
    if ((rc = ba_read_ctl(h))) return rc;
    if (qmax == 0) currentChi = iniChi = h->h_ctl[8];
    const int ok2 = *(int*)(h->h_ctl + 9);
    double tempChi = h->h_ctl[1];
    if (!ok2) tempChi = std::numeric_limits<double>::max();
    rho = currentChi - tempChi;
    double scale = h->h_ctl[3] + h->h_ctl[2];
    scale += 1e-3;
    rho /= scale;
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = std::min(alpha, 2. / 3.);
      h->lambda *= std::max(1. / 3., alpha);
      h->ni = 2;
      currentChi = tempChi;
    } else {
      h->lambda *= h->ni;
      h->ni *= 2;
      BA_CK(cudaMemcpyAsync(h->d_st, h->d_st_bak, sizeof(VieoNavState) * h->K, cudaMemcpyDeviceToDevice, h->st));
      if (h->P) BA_CK(cudaMemcpyAsync(h->d_X, h->d_X_bak, sizeof(double) * 3 * h->P, cudaMemcpyDeviceToDevice, h->st));
    }
    qmax++;
  } while (rho < 0 && qmax < 10 && !(stop && *stop));
  if (qmax == 10 || rho == 0) return 1;
  if ((iniChi - currentChi) * 1e3 < iniChi) h->nBad++;
  else h->nBad = 0;
  if (h->nBad >= 3) return 1;
  return 0;
}

void ba_free(vieo_ba* h) {
  void* ptrs[] = {h->d_st, h->d_st_bak, h->d_cp, h->d_X, h->d_X_bak, h->d_chi2, h->d_W, h->d_A, h->d_Hll, h->d_bl, h->d_Dinv,
                  h->d_db, h->d_H, h->d_sys, h->d_x, h->d_y, h->d_xl, h->d_partial, h->d_scale_part, h->d_ctl, h->d_es,
                  h->d_ep, h->d_pt_ptr, h->d_off0, h->d_off1, h->d_off2, h->d_prcol, h->d_free_state, h->d_free_off,
                  h->d_ps_ptr, h->d_ps_edges, h->d_ok, h->d_obs, h->d_w, h->d_flags, h->d_lvl, h->d_sfix, h->d_pt_active,
                  h->d_bad, h->d_pre, h->d_den, h->d_wk};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->st) cudaStreamDestroy(h->st);
}

}  // namespace

extern "C" {

int vieo_ba_create(int max_states, int max_points, int max_edges, int max_imu, int device, vieo_ba_t** out) {
  VIEO_ARG(out && max_states > 0 && max_points >= 0 && max_edges >= 0 && max_imu >= 0, "bad argument");
  int rc = use_device(device);
  if (rc) return rc;
  vieo_ba* h = new vieo_ba();
  h->device = device;
  h->capK = max_states; h->capP = max_points; h->capE = max_edges; h->capM = max_imu;
  const size_t K = max_states, P = max_points, E = max_edges, M = max_imu, NP = 15 * K;
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  step(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  step(dalloc(&h->d_st, K)); step(dalloc(&h->d_st_bak, K)); step(dalloc(&h->d_cp, K));
  step(dalloc(&h->d_X, 3 * P)); step(dalloc(&h->d_X_bak, 3 * P)); step(dalloc(&h->d_chi2, E));
  step(dalloc(&h->d_W, 18 * E)); step(dalloc(&h->d_A, 27 * E)); step(dalloc(&h->d_Hll, 9 * P)); step(dalloc(&h->d_bl, 3 * P));
  step(dalloc(&h->d_Dinv, 9 * P)); step(dalloc(&h->d_db, 3 * P)); step(dalloc(&h->d_H, NP * NP));
  step(dalloc(&h->d_sys, NP * NP + 2 * NP + 8)); step(dalloc(&h->d_x, NP)); step(dalloc(&h->d_y, NP));
  step(dalloc(&h->d_xl, 3 * P)); step(dalloc(&h->d_partial, (E + 255) / 256 + 1)); step(dalloc(&h->d_scale_part, P));
  step(dalloc(&h->d_ctl, 16)); step(dalloc(&h->d_es, E)); step(dalloc(&h->d_ep, E)); step(dalloc(&h->d_pt_ptr, P + 1));
  step(dalloc(&h->d_off0, K)); step(dalloc(&h->d_off1, K)); step(dalloc(&h->d_off2, K)); step(dalloc(&h->d_prcol, K));
  step(dalloc(&h->d_free_state, K)); step(dalloc(&h->d_free_off, K)); step(dalloc(&h->d_ps_ptr, K + 1));
  step(dalloc(&h->d_ps_edges, E)); step(dalloc(&h->d_ok, 4)); step(dalloc(&h->d_obs, 3 * E)); step(dalloc(&h->d_w, E));
  step(dalloc(&h->d_flags, E)); step(dalloc(&h->d_lvl, E)); step(dalloc(&h->d_sfix, K)); step(dalloc(&h->d_pt_active, P));
  step(dalloc(&h->d_bad, E)); step(dalloc(&h->d_pre, M)); step(dalloc(&h->d_den, 2 * M)); step(dalloc(&h->d_wk, 2 * M));
  step(cudaMallocHost((void**)&h->h_ctl, sizeof(double) * 16));
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

int vieo_ba_set_sharding(vieo_ba_t* h, int rank, int world, vieo_allreduce_fn allreduce, void* ctx) {
  VIEO_ARG(h && world >= 1 && rank >= 0 && rank < world, "bad argument");
  h->rank = rank; h->world = world; h->allreduce = allreduce; h->ar_ctx = ctx;
  return VIEO_OK;
}

void* vieo_ba_stream(vieo_ba_t* h) { return h ? (void*)h->st : nullptr; }
int vieo_ba_last_launches(const vieo_ba_t* h) { return h ? h->launches : 0; }

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
  BA_CK(cudaSetDevice(h->device));
  h->K = K; h->P = P; h->E = E; h->M = M;
  h->launches = 0;
  h->points_free = true;
  h->n_part = (E + 255) / 256;
  // camera
  h->cam.fx = (double)cam->fx; h->cam.fy = (double)cam->fy; h->cam.cx = (double)cam->cx; h->cam.cy = (double)cam->cy;
  h->cam.bf = (double)cam->bf;
  for (int i = 0; i < 9; ++i) h->cam.Rcb.m[i] = cam->Rcb[i];
  h->cam.tcb = {cam->tcb[0], cam->tcb[1], cam->tcb[2]};
  h->gw = {pb->gw[0], pb->gw[1], pb->gw[2]};
  const float chi2Mono = 5.991f;
  h->dm = (double)std::sqrt(chi2Mono);           // thHuberMono (src/Optimizer.cc:361)
  h->ds = (double)(float)std::sqrt(7.815);       // thHuberStereo
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
    const bool bfixedkf = pb->state_flags[i] & 1;
    const VieoImuPreint& pre = pb->preint[m];
    if (pre.dt != 0) {
      BaDense d;
      memset(&d, 0, sizeof(d));
      d.type = 0; d.si = i; d.sj = j; d.pre = m;
      if (!host_inverse(pre.SigmaPRV, 9, d.info))
        for (double& v : d.info) v = std::numeric_limits<double>::quiet_NaN();
      if (bfixedkf || pb->rec_init) {
        if (bfixedkf) for (double& v : d.info) v *= 1e-2;
        d.delta = (double)thPRV;
      }
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
    if (bfixedkf || pb->rec_init) d.delta = (double)thBias;
    den.push_back(d);
  }
  h->n_den = (int)den.size();
  VIEO_ARG(h->n_den <= 1024, "too many inertial edges");
  std::vector<uint8_t> lvl(std::max(E, 1));
  for (int i = 0; i < E; ++i) lvl[i] = ((pb->edge_flags[i] & VIEO_EDGE_LEVEL1) ? 1 : 0) | ((pb->edge_flags[i] & VIEO_EDGE_NOKERNEL) ? 2 : 0);
  auto up = [&](void* d, const void* s, size_t n) { return n ? cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, h->st) : cudaSuccess; };
  BA_CK(up(h->d_st, pb->states, sizeof(VieoNavState) * K));
  BA_CK(up(h->d_X, pb->points, 24 * (size_t)P));
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
  BA_CK(up(h->d_free_state, free_state.data(), 4 * (size_t)h->nfree));
  BA_CK(up(h->d_free_off, free_off.data(), 4 * (size_t)h->nfree));
  BA_CK(up(h->d_ps_ptr, ps_ptr.data(), 4 * (size_t)(h->nfree + 1)));
  BA_CK(up(h->d_ps_edges, ps_edges.data(), 4 * (size_t)ps_ptr[h->nfree]));
  BA_CK(up(h->d_pre, pb->preint, sizeof(VieoImuPreint) * (size_t)M));
  BA_CK(up(h->d_den, den.data(), sizeof(BaDense) * den.size()));
  BA_CK(cudaMemsetAsync(h->d_chi2, 0, 8 * (size_t)std::max(E, 1), h->st));
  BA_CK(cudaMemsetAsync(h->d_ctl, 0, 8 * 16, h->st));
  BA_CK(cudaMemsetAsync(h->d_x, 0, 8 * (size_t)std::max(np, 1), h->st));
  BA_CK(cudaMemsetAsync(h->d_ok, 0, 16, h->st));
  BA_CK(cudaStreamSynchronize(h->st));  // the host staging vectors die here
  const size_t smem = sizeof(double) * kBaWarps * 6 * (6 * (size_t)h->nfree + 1);
  if (smem > 200 * 1024) {
    set_error("vieo_ba_set_problem: %d free keyframes exceed the Schur kernel's shared-memory tile", h->nfree);
    return VIEO_E_CAPACITY;
  }
  if (smem > 48 * 1024) BA_CK(cudaFuncSetAttribute(k_ba_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  h->lambda = 0; h->ni = 2; h->nBad = 0;
  return VIEO_OK;
}

int vieo_ba_chi2_large_set_level(vieo_ba_t* h, float rat) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  int rc = ba_errors(h, 1, h->d_ctl + 5, false);
  if (rc) return rc;
  if (h->E > 0) {
    k_ba_classify<<<(h->E + 255) / 256, 256, 0, h->st>>>(h->cam, h->d_cp, h->d_X, h->d_es, h->d_ep, h->d_obs, h->d_flags,
                                                         h->d_chi2, h->E, 0, rat, 1, 0, h->d_lvl, nullptr);
    h->launches++;
  }
  BA_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_ba_active_robust_chi2(vieo_ba_t* h, int recompute, double* chi2) {
  VIEO_ARG(h && chi2, "null argument");
  BA_CK(cudaSetDevice(h->device));
  if (recompute) {
    int rc = ba_errors(h, 0, h->d_ctl + 6, false);
    if (rc) return rc;
  } else {
    // activeRobustChi2 over the stored errors of the current active set (levels may have changed since)
    // -> recompute the partial sums without touching chi2: run the error kernel in "sum only" fashion is not needed:
    // the stored chi2 of active edges equals a recomputation unless the last LM trial was rejected; the reference sums
    // the stored values, so do that on the host.
    std::vector<double> c(std::max(h->E, 1));
    std::vector<uint8_t> lvl(std::max(h->E, 1)), fl(std::max(h->E, 1));
    BA_CK(cudaMemcpyAsync(c.data(), h->d_chi2, 8 * (size_t)h->E, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaMemcpyAsync(lvl.data(), h->d_lvl, (size_t)h->E, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaMemcpyAsync(fl.data(), h->d_flags, (size_t)h->E, cudaMemcpyDeviceToHost, h->st));
    std::vector<BaDenseWork> wk(std::max(h->n_den, 1));
    std::vector<BaDense> den(std::max(h->n_den, 1));
    BA_CK(cudaMemcpyAsync(wk.data(), h->d_wk, sizeof(BaDenseWork) * h->n_den, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaMemcpyAsync(den.data(), h->d_den, sizeof(BaDense) * h->n_den, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaStreamSynchronize(h->st));
    auto rho0 = [](double delta, double e) {
      const double dsqr = (double)(float)(delta * delta);
      if (delta == 0 || e <= dsqr) return e;
      return 2 * std::sqrt(e) * delta - dsqr;
    };
    double tot = 0;
    for (int m = 0; m < h->n_den; ++m) tot += rho0(den[m].delta, wk[m].chi2);
    for (int i = 0; i < h->E; ++i) {
      if (lvl[i] & 1) continue;
      const double d = (lvl[i] & 2) ? 0.0 : ((fl[i] & VIEO_EDGE_STEREO) ? h->ds : h->dm);
      tot += rho0(d, c[i]);
    }
    if (h->allreduce && h->world > 1) {  // partial sums of the ranks' own edges
      h->h_ctl[7] = tot;
      BA_CK(cudaMemcpyAsync(h->d_ctl + 7, h->h_ctl + 7, 8, cudaMemcpyHostToDevice, h->st));
      if (h->allreduce(h->ar_ctx, h->d_ctl + 7, 1, (void*)h->st)) return VIEO_E_CUDA;
      BA_CK(cudaMemcpyAsync(h->h_ctl + 7, h->d_ctl + 7, 8, cudaMemcpyDeviceToHost, h->st));
      BA_CK(cudaStreamSynchronize(h->st));
      tot = h->h_ctl[7];
    }
    *chi2 = tot;
    return VIEO_OK;
  }
  if (h->allreduce && h->world > 1)
    if (h->allreduce(h->ar_ctx, h->d_ctl + 6, 1, (void*)h->st)) return VIEO_E_CUDA;
  BA_CK(cudaMemcpyAsync(h->h_ctl + 6, h->d_ctl + 6, 8, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  *chi2 = h->h_ctl[6];
  return VIEO_OK;
}

int vieo_ba_optimize(vieo_ba_t* h, int iterations, double lambda_init, const volatile uint8_t* stop) {
  VIEO_ARG(h && iterations >= 0, "bad argument");
  BA_CK(cudaSetDevice(h->device));
  if (h->np == 0) return 0;
  BA_CK(cudaMemsetAsync(h->d_x, 0, 8 * (size_t)h->np, h->st));
  int n = 0;
  bool ok = true;
  for (int i = 0; i < iterations && !(stop && *stop) && ok; ++i) {
    const int r = ba_lm_iteration(h, i, lambda_init, stop);
    if (r < 0) return r;
    ok = r == 0;
    ++n;
  }
  return n;
}

int vieo_ba_reclassify(vieo_ba_t* h, int remove_kernels, uint8_t* bad_host) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  if (h->E == 0) return VIEO_OK;
  ba_campose(h);
  k_ba_classify<<<(h->E + 255) / 256, 256, 0, h->st>>>(h->cam, h->d_cp, h->d_X, h->d_es, h->d_ep, h->d_obs, h->d_flags,
                                                       h->d_chi2, h->E, 1, 0.f, bad_host ? 0 : 1, remove_kernels, h->d_lvl,
                                                       h->d_bad);
  h->launches++;
  BA_CK(cudaGetLastError());
  if (bad_host) {
    BA_CK(cudaMemcpyAsync(bad_host, h->d_bad, (size_t)h->E, cudaMemcpyDeviceToHost, h->st));
    BA_CK(cudaStreamSynchronize(h->st));
  }
  return VIEO_OK;
}

int vieo_ba_get(vieo_ba_t* h, VieoNavState* states_out, double* points_out, double* edge_chi2) {
  VIEO_ARG(h, "null handle");
  BA_CK(cudaSetDevice(h->device));
  if (states_out) BA_CK(cudaMemcpyAsync(states_out, h->d_st, sizeof(VieoNavState) * h->K, cudaMemcpyDeviceToHost, h->st));
  if (points_out && h->P) BA_CK(cudaMemcpyAsync(points_out, h->d_X, 24 * (size_t)h->P, cudaMemcpyDeviceToHost, h->st));
  if (edge_chi2 && h->E) BA_CK(cudaMemcpyAsync(edge_chi2, h->d_chi2, 8 * (size_t)h->E, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  return VIEO_OK;
}

int vieo_ba_debug_step(vieo_ba_t* h, double lambda, double* x_pose, double* x_points, double* H_out, double* b_out) {
  VIEO_ARG(h && x_pose, "null argument");
  BA_CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = ba_errors(h, 0, h->chi_cur(), false))) return rc;
  if ((rc = ba_build(h))) return rc;
  if ((rc = ba_solve(h, lambda, 0, h->d_xl))) return rc;
  BA_CK(cudaMemcpyAsync(x_pose, h->d_x, 8 * (size_t)h->np, cudaMemcpyDeviceToHost, h->st));
  if (x_points && h->P) BA_CK(cudaMemcpyAsync(x_points, h->d_xl, 24 * (size_t)h->P, cudaMemcpyDeviceToHost, h->st));
  if (H_out) BA_CK(cudaMemcpyAsync(H_out, h->d_H, 8 * (size_t)h->np * h->np, cudaMemcpyDeviceToHost, h->st));
  if (b_out) BA_CK(cudaMemcpyAsync(b_out, h->b(), 8 * (size_t)h->np, cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaMemcpyAsync(h->h_ctl + 9, h->d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  BA_CK(cudaStreamSynchronize(h->st));
  if (!*(int*)(h->h_ctl + 9)) {
    set_error("vieo_ba_debug_step: reduced camera system is not positive definite");
    return VIEO_E_ARG;
  }
  return h->np;
}

// Optimizer::LocalBundleAdjustmentNavStatePRV, src/Optimizer.cc:133-700, on the flattened problem
int vieo_local_ba_prv(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, const volatile uint8_t* stop,
                      VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase, VieoBaResult* res) {
  VIEO_ARG(h && pb && cam && res && states_out && erase, "null argument");
  memset(res, 0, sizeof(*res));
  memcpy(states_out, pb->states, sizeof(VieoNavState) * pb->n_states);
  if (points_out && pb->n_points) memcpy(points_out, pb->points, 24 * (size_t)pb->n_points);
  memset(erase, 0, pb->n_edges);
  int optit[2];
  double lambda0;
  if (pb->visual_only) { optit[0] = 5; optit[1] = 10; lambda0 = 0; }
  else if (pb->large) { optit[0] = 2; optit[1] = 2; lambda0 = 1e-2; }
  else { optit[0] = 4; optit[1] = 6; lambda0 = 1e0; }
  bool anyfree = false;
  for (int k = 0; k < pb->n_states; ++k) anyfree |= !(pb->state_flags[k] & 1);
  if (!anyfree) return VIEO_OK;  // if (!bdimPoses) return; (:178)
  int rc = vieo_ba_set_problem(h, pb, cam);
  if (rc) return rc;
  if (stop && *stop) return VIEO_OK;  // "Aborted OLBA" (:524-528)
  if ((rc = vieo_ba_chi2_large_set_level(h, 100.f))) return rc;
  double chi = 0;
  if ((rc = vieo_ba_active_robust_chi2(h, 1, &chi))) return rc;
  const float err = (float)chi;
  res->err0 = err;
  int n = vieo_ba_optimize(h, optit[0], lambda0, stop);
  if (n < 0) return n;
  res->iterations[0] = n;
  bool bDoMore = true;
  if (stop && *stop) bDoMore = false;
  if (bDoMore) {
    if ((rc = vieo_ba_reclassify(h, 1, nullptr))) return rc;
    n = vieo_ba_optimize(h, optit[1], lambda0, stop);
    if (n < 0) return n;
    res->iterations[1] = n;
  }
  if ((rc = vieo_ba_active_robust_chi2(h, 0, &chi))) return rc;
  const float err_end = (float)chi;
  res->err_end = err_end;
  res->lambda_final = h->lambda;
  if ((rc = vieo_ba_get(h, nullptr, nullptr, edge_chi2))) return rc;
  if ((2 * err < err_end || std::isnan(err) || std::isnan(err_end)) && !pb->large) {
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

// Motion-only bundle adjustment on the device — Optimizer::PoseOptimization.
//   mode 0  visual, one VertexNavStatePR (6)                          src/Optimizer.cc:1611-1874
//   mode 1  IMU, VertexNavStatePVR (9) + Bias (6) for the frame and, when the last frame carries a prior, for the
//           last frame too (30-dim), EdgeNavStatePVR + EdgeNavStateBias + EdgeNavStatePriorPVRBias, and the
//           kExactRobust marginal prior                                include/Optimizer.h:126-816
// One thread block per frame runs the reference's whole schedule — 4 rounds x optimize(10) of g2o's
// Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:61-189), inlier re-classification between rounds,
// rescue pass, marginalisation — without leaving the SM: the visual edges are evaluated one per thread and reduced
// with warp shuffles in a fixed order (deterministic), the <= 30-dim normal equations live in shared memory.
// Batched over frames (grid = number of frames) this is the throughput form of the tracking thread's hot loop.
#include <algorithm>

#include "ba_edges.cuh"

namespace vieo {

constexpr int kPoThreads = 256;
constexpr int kPoWarps = kPoThreads / 32;
constexpr int kPoN = 30;          // largest system: cur PVR 9 + bias 6 + last PVR 9 + bias 6
constexpr int kPoMaxEdges = 4096; // per frame (the reference has N <= nfeatures + a few per camera)

struct PoSmem {
  double H[kPoN * kPoN], S[kPoN * kPoN];
  double b[kPoN], x[kPoN], y[kPoN];
  double red[kPoWarps][32];
  double tot[32];
  double Ji[135], Jj[81], Jb[90], Om[225], AtO[15 * 15], oe[15];
  double info_imu[81], info_prior[225], info_bias[6];
  double err_imu[9], err_bias[6], err_prior[15];
  double chi2_imu, chi2_bias, chi2_prior;
  double C[225], CL[225], CCL[225];
  NavS st[2], bak[2], ini[2];
  CamPose cp;
  double lambda, ni, currentChi, tempChi;
  int nBad, ctl, total_iters, nbad_edges;
  uint8_t eflag[kPoMaxEdges];  // bit0: level 1, bit1: kernel removed, bit2: outlier
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
  double delta_mono, delta_stereo;
  Vec3 gw;
};

// block-wide sum of `nv` doubles per thread (v[0..nv)), fixed order: lanes by xor tree, warps 0..7 in sequence.
// Result in sm.tot[0..nv) (valid for all threads after the call).
template <int NV>
__device__ __forceinline__ void block_sum(PoSmem& sm, double (&v)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    v[k] = a;
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm.red[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double a = 0;
    for (int w = 0; w < kPoWarps; ++w) a += sm.red[w][threadIdx.x];
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

// errors + chi2 of the IMU / bias / prior edges (thread 0)
__device__ void dense_errors(const PoCtx& c, PoSmem& sm, const NavS& prior) {
  auto chi = [](const double* info, const double* e, int D) {
    double s = 0;
    for (int i = 0; i < D; ++i) {
      double t = 0;
      for (int j = 0; j < D; ++j) t += info[i * D + j] * e[j];
      s += e[i] * t;
    }
    return s;
  };
  if (c.has_imu) {
    navstate_error(sm.st[1], sm.st[0], c.pb->preint, c.gw, false, sm.err_imu);
    sm.chi2_imu = chi(sm.info_imu, sm.err_imu, 9);
  }
  {
    const NavS &a = sm.st[1], &d = sm.st[0];
    sm.err_bias[0] = (d.bg.x + d.dbg.x) - (a.bg.x + a.dbg.x);
    sm.err_bias[1] = (d.bg.y + d.dbg.y) - (a.bg.y + a.dbg.y);
    sm.err_bias[2] = (d.bg.z + d.dbg.z) - (a.bg.z + a.dbg.z);
    sm.err_bias[3] = (d.ba.x + d.dba.x) - (a.ba.x + a.dba.x);
    sm.err_bias[4] = (d.ba.y + d.dba.y) - (a.ba.y + a.dba.y);
    sm.err_bias[5] = (d.ba.z + d.dba.z) - (a.ba.z + a.dba.z);
    double s = 0;
    for (int i = 0; i < 6; ++i) s += sm.err_bias[i] * (sm.info_bias[i] * sm.err_bias[i]);
    sm.chi2_bias = s;
  }
  if (!c.fixed_last) {
    prior_error(sm.st[1], prior, sm.err_prior);
    sm.chi2_prior = chi(sm.info_prior, sm.err_prior, 15);
  }
}

// H[oa.., ob..] += Ja^T (r1 info) Jb, b[oa..] += Ja^T (-r1 info e)   (BaseMultiEdge::constructQuadraticForm)
struct PoBlk {
  int off, dim;
  const double* J;
  int ld, c0;
};
__device__ void add_dense(PoSmem& sm, int n, int D, const double* info, const double* err, double r1, const PoBlk* blks,
                          int nb) {
  for (int i = 0; i < D * D; ++i) sm.Om[i] = r1 * info[i];
  for (int i = 0; i < D; ++i) {
    double s = 0;
    for (int j = 0; j < D; ++j) s += info[i * D + j] * err[j];
    sm.oe[i] = -s * r1;
  }
  for (int ia = 0; ia < nb; ++ia) {
    const PoBlk& A = blks[ia];
    if (A.off < 0) continue;
    for (int a = 0; a < A.dim; ++a)
      for (int j = 0; j < D; ++j) {
        double s = 0;
        for (int i = 0; i < D; ++i) s += A.J[i * A.ld + A.c0 + a] * sm.Om[i * D + j];
        sm.AtO[a * D + j] = s;
      }
    for (int a = 0; a < A.dim; ++a) {
      double s = 0;
      for (int i = 0; i < D; ++i) s += A.J[i * A.ld + A.c0 + a] * sm.oe[i];
      sm.b[A.off + a] += s;
    }
    for (int ib = 0; ib < nb; ++ib) {
      const PoBlk& B = blks[ib];
      if (B.off < 0) continue;
      for (int a = 0; a < A.dim; ++a)
        for (int cc = 0; cc < B.dim; ++cc) {
          double s = 0;
          for (int j = 0; j < D; ++j) s += sm.AtO[a * D + j] * B.J[j * B.ld + B.c0 + cc];
          sm.H[(A.off + a) * n + B.off + cc] += s;
        }
    }
  }
}

// Cholesky solve (H + lambda I) x = b in shared memory by one thread; false when a pivot is not positive
__device__ bool chol_solve(PoSmem& sm, int n) {
  double* A = sm.S;
  for (int i = 0; i < n * n; ++i) A[i] = sm.H[i];
  for (int i = 0; i < n; ++i) A[i * n + i] += sm.lambda;
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = sm.b[i];
    for (int k = 0; k < i; ++k) s -= A[i * n + k] * sm.y[k];
    sm.y[i] = s / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = sm.y[i];
    for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * sm.x[k];
    sm.x[i] = s / A[i * n + i];
  }
  return true;
}

// computeActiveErrors + activeRobustChi2: visual edges in parallel (chi2 stored per edge), dense edges by thread 0.
// Returns the robust chi2 to every thread.
__device__ double active_errors(const PoCtx& c, PoSmem& sm, const NavS& prior) {
  if (threadIdx.x == 0) {
    sm.cp = cam_pose(c.cam, sm.st[0]);
    if (c.imu_mode) dense_errors(c, sm, prior);
  }
  __syncthreads();
  double acc[1] = {0.0};
  for (int i = threadIdx.x; i < c.E; i += kPoThreads) {
    if (sm.eflag[i] & 1) continue;
    double e[3];
    const double chi = vis_chi2(c, sm, i, e);
    c.chi2[i] = chi;
    double r0, r1;
    huber_rho(edge_delta(c, sm, i), chi, r0, r1);
    acc[0] += r0;
  }
  block_sum<1>(sm, acc);
  double total = 0;
  if (c.imu_mode) {
    double r0, r1;
    if (c.has_imu) {
      huber_rho(c.fixed_last ? sqrt(16.919) : 0.0, sm.chi2_imu, r0, r1);
      total += r0;
    }
    huber_rho(c.fixed_last ? sqrt(12.592) : 0.0, sm.chi2_bias, r0, r1);
    total += r0;
    if (!c.fixed_last) {
      huber_rho(sqrt(25.0), sm.chi2_prior, r0, r1);
      total += r0;
    }
  }
  return total + sm.tot[0];
}

// buildSystem: H, b at the current estimate with the errors of the last active_errors()
__device__ void build_system(const PoCtx& c, PoSmem& sm, const NavS& prior) {
  const int n = c.n, dv = c.dv;
  for (int i = threadIdx.x; i < n * n; i += kPoThreads) sm.H[i] = 0;
  if (threadIdx.x < n) sm.b[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0 && c.imu_mode) {
    double r0, r1;
    const int oL = c.fixed_last ? -1 : 15, oLb = c.fixed_last ? -1 : 24;
    if (c.has_imu) {
      navstate_jac(sm.st[1], sm.st[0], c.pb->preint, c.gw, false, sm.err_imu, sm.Ji, sm.Jj, sm.Jb);
      huber_rho(c.fixed_last ? sqrt(16.919) : 0.0, sm.chi2_imu, r0, r1);
      const PoBlk blks[3] = {{oL, 9, sm.Ji, 9, 0}, {0, 9, sm.Jj, 9, 0}, {oLb, 6, sm.Jb, 6, 0}};
      add_dense(sm, n, 9, sm.info_imu, sm.err_imu, r1, blks, 3);
    }
    {
      huber_rho(c.fixed_last ? sqrt(12.592) : 0.0, sm.chi2_bias, r0, r1);
      // J_i = -I (last bias), J_j = +I (frame bias), diagonal information
      for (int k = 0; k < 6; ++k) {
        const double om = r1 * sm.info_bias[k];
        const double oe = -(sm.info_bias[k] * sm.err_bias[k]) * r1;
        sm.H[(9 + k) * n + 9 + k] += om;
        sm.b[9 + k] += oe;
        if (oLb >= 0) {
          sm.H[(oLb + k) * n + oLb + k] += om;
          sm.b[oLb + k] += -oe;
          sm.H[(oLb + k) * n + 9 + k] += -om;
          sm.H[(9 + k) * n + oLb + k] += -om;
        }
      }
    }
    if (!c.fixed_last) {
      prior_jac(sm.st[1], prior, sm.err_prior, sm.Ji, sm.Jb);
      huber_rho(sqrt(25.0), sm.chi2_prior, r0, r1);
      const PoBlk blks[2] = {{15, 9, sm.Ji, 9, 0}, {24, 6, sm.Jb, 6, 0}};
      add_dense(sm, n, 15, sm.info_prior, sm.err_prior, r1, blks, 2);
    }
  }
  // visual edges: 6 non-zero pose columns (dp, dphi) -> 21 upper-triangle entries + 6 rhs
  double acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0;
  for (int i = threadIdx.x; i < c.E; i += kPoThreads) {
    if (sm.eflag[i] & 1) continue;
    const bool stereo = c.flags[i] & VIEO_EDGE_STEREO;
    const int DE = stereo ? 3 : 2;
    const Vec3 X = ld3(c.Xw + 3 * (size_t)i);
    double e[3];
    reproj_error(c.cam, sm.cp, X, c.obs + 3 * (size_t)i, stereo, e);
    Mat3 Jp, Jr, JX;
    reproj_jac(c.cam, sm.cp, X, stereo, Jp, Jr, JX);
    double r0, r1;
    huber_rho(edge_delta(c, sm, i), c.chi2[i], r0, r1);
    const double wi = (double)c.w[i], w = r1 * wi;
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
  block_sum<27>(sm, acc);
  if (threadIdx.x == 0) {
    int q = 0;
    for (int a = 0; a < 6; ++a) {
      const int ra = a < 3 ? a : dv - 6 + a;
      sm.b[ra] += sm.tot[21 + a];
      for (int cc = a; cc < 6; ++cc) {
        const int rc = cc < 3 ? cc : dv - 6 + cc;
        const double h = sm.tot[q++];
        sm.H[ra * n + rc] += h;
        if (rc != ra) sm.H[rc * n + ra] += h;
      }
    }
  }
  __syncthreads();
}

__device__ void apply_update(const PoCtx& c, PoSmem& sm) {
  if (!c.imu_mode) {
    ns_inc_pr(sm.st[0], sm.x);
    return;
  }
  ns_inc_pvr(sm.st[0], sm.x);
  ns_inc_bias(sm.st[0], sm.x + 9);
  if (!c.fixed_last) {
    ns_inc_pvr(sm.st[1], sm.x + 15);
    ns_inc_bias(sm.st[1], sm.x + 24);
  }
}

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg::solve inlined
__device__ void optimize(const PoCtx& c, PoSmem& sm, const NavS& prior, int iterations) {
  const int n = c.n;
  if (threadIdx.x < n) sm.x[threadIdx.x] = 0;
  bool ok = true;
  for (int it = 0; it < iterations && ok; ++it) {
    double currentChi = active_errors(c, sm, prior);
    const double iniChi = currentChi;
    build_system(c, sm, prior);
    if (threadIdx.x == 0) {
      if (it == 0) {  // computeLambdaInit: tau * max |diag(H)|
        double mx = 0;
        for (int i = 0; i < n; ++i) mx = fmax(fabs(sm.H[i * n + i]), mx);
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
      if (threadIdx.x == 0) {
        sm.bak[0] = sm.st[0];
        sm.bak[1] = sm.st[1];
        sm.ctl = chol_solve(sm, n) ? 1 : 0;
        apply_update(c, sm);
      }
      __syncthreads();
      double tempChi = active_errors(c, sm, prior);
      if (!sm.ctl) tempChi = 1.7976931348623157e308;
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < n; ++j) scale += sm.x[j] * (sm.lambda * sm.x[j] + sm.b[j]);
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
      } else {
        if (threadIdx.x == 0) {
          sm.lambda *= sm.ni;
          sm.ni *= 2;
          sm.st[0] = sm.bak[0];
          sm.st[1] = sm.bak[1];
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

__global__ void __launch_bounds__(kPoThreads) k_pose_opt(const VieoPoseOptProblem* __restrict__ pbs,
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
  c.gw = ld3(pb.gw);
  const NavS prior = ns_load(pb.prior);
  const int E = c.E;

  for (int i = threadIdx.x; i < E; i += kPoThreads) {
    sm.eflag[i] = 0;
    outl[i] = 0;
    c.chi2[i] = 0;
  }
  if (threadIdx.x == 0) {
    sm.st[0] = sm.ini[0] = ns_load(pb.cur);
    sm.st[1] = sm.ini[1] = ns_load(pb.last);
    sm.total_iters = 0;
    sm.lambda = 0;
    if (c.imu_mode) {
      if (c.has_imu) {  // GetProcessedInfoij = mSigmaij.inverse() (OdomPreIntegrator.h:129-138), x 1e-2 when last is fixed
        for (int i = 0; i < 81; ++i) sm.S[i] = pb.preint.SigmaPVR[i];
        if (!dense_inverse(sm.S, 9, sm.info_imu))
          for (int i = 0; i < 81; ++i) sm.info_imu[i] = nan("");
        if (c.fixed_last)
          for (int i = 0; i < 81; ++i) sm.info_imu[i] *= 1e-2;
      }
      const double dtij = pb.preint.dt != 0 ? pb.preint.dt : pb.dt_frames;
      for (int k = 0; k < 6; ++k) {
        const double w = (k < 3 ? pb.inv_sigma_bg2 : pb.inv_sigma_ba2) / dtij;
        sm.info_bias[k] = c.fixed_last ? w * 1e-2 : w;
      }
      if (!c.fixed_last)
        for (int i = 0; i < 225; ++i) sm.info_prior[i] = pb.prior_info[i];
    }
  }
  __syncthreads();
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
    optimize(c, sm, prior, 10);
    if (threadIdx.x == 0) sm.cp = cam_pose(c.cam, sm.st[0]);
    __syncthreads();
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
    block_sum<1>(sm, bad);
    nBad = (int)sm.tot[0];
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
    block_sum<1>(sm, bad);
    nBad = (int)sm.tot[0];
  }
  // activeRobustChi2 of the final active set with the stored errors
  {
    double acc[1] = {0};
    for (int i = threadIdx.x; i < E; i += kPoThreads) {
      if (sm.eflag[i] & 1) continue;
      double r0, r1;
      huber_rho(edge_delta(c, sm, i), c.chi2[i], r0, r1);
      acc[0] += r0;
    }
    block_sum<1>(sm, acc);
  }
  if (threadIdx.x == 0) {
    ns_store(sm.st[0], R.cur);
    ns_store(c.imu_mode ? sm.st[1] : sm.ini[1], R.last);
    double tot = sm.tot[0], r0, r1;
    if (c.imu_mode) {
      if (c.has_imu) { huber_rho(c.fixed_last ? sqrt(16.919) : 0.0, sm.chi2_imu, r0, r1); tot += r0; }
      huber_rho(c.fixed_last ? sqrt(12.592) : 0.0, sm.chi2_bias, r0, r1); tot += r0;
      if (!c.fixed_last) { huber_rho(sqrt(25.0), sm.chi2_prior, r0, r1); tot += r0; }
    }
    R.chi2_final = tot;
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
  __syncthreads();
  double wI = 1, wB = 1;
  if (threadIdx.x == 0) {
    sm.cp = cam_pose(c.cam, sm.st[0]);
    dense_errors(c, sm, prior);
    for (int i = 0; i < 225; ++i) sm.C[i] = 0;
    double r0;
    if (c.has_imu) {
      navstate_jac(sm.st[1], sm.st[0], pb.preint, c.gw, false, sm.err_imu, sm.Ji, sm.Jj, sm.Jb);
      huber_rho(c.fixed_last ? sqrt(16.919) : 0.0, sm.chi2_imu, r0, wI);
      jtoj(sm.Jj, 9, 0, 9, sm.info_imu, 9, wI, sm.Jj, 9, 0, 9, sm.C, 15, 0, 0, false);
    }
    huber_rho(c.fixed_last ? sqrt(12.592) : 0.0, sm.chi2_bias, r0, wB);
    for (int k = 0; k < 6; ++k) sm.C[(9 + k) * 15 + 9 + k] = wB * sm.info_bias[k];
  }
  __syncthreads();
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
    block_sum<21>(sm, acc);
  }
  if (threadIdx.x == 0) {
    int q = 0;
    for (int a = 0; a < 6; ++a) {
      const int ra = a < 3 ? a : 3 + a;
      for (int cc = a; cc < 6; ++cc) {
        const int rc = cc < 3 ? cc : 3 + cc;
        const double h = sm.tot[q++];
        sm.C[ra * 15 + rc] += h;
        if (rc != ra) sm.C[rc * 15 + ra] += h;
      }
    }
    if (!c.fixed_last) {
      double r0, wP;
      for (int i = 0; i < 225; ++i) sm.CL[i] = sm.CCL[i] = 0;
      if (c.has_imu) {
        jtoj(sm.Ji, 9, 0, 9, sm.info_imu, 9, wI, sm.Ji, 9, 0, 9, sm.CL, 15, 0, 0, false);
        jtoj(sm.Ji, 9, 0, 9, sm.info_imu, 9, wI, sm.Jb, 6, 0, 6, sm.CL, 15, 0, 9, false);
        jtoj(sm.Jb, 6, 0, 6, sm.info_imu, 9, wI, sm.Jb, 6, 0, 6, sm.CL, 15, 9, 9, false);
        for (int a = 0; a < 9; ++a)
          for (int cc = 0; cc < 6; ++cc) sm.CL[(9 + cc) * 15 + a] = sm.CL[a * 15 + 9 + cc];
        jtoj(sm.Jj, 9, 0, 9, sm.info_imu, 9, wI, sm.Ji, 9, 0, 9, sm.CCL, 15, 0, 0, false);
        jtoj(sm.Jj, 9, 0, 9, sm.info_imu, 9, wI, sm.Jb, 6, 0, 6, sm.CCL, 15, 0, 9, false);
      }
      for (int k = 0; k < 6; ++k) {
        sm.CL[(9 + k) * 15 + 9 + k] += wB * sm.info_bias[k];
        sm.CCL[(9 + k) * 15 + 9 + k] = -(wB * sm.info_bias[k]);
      }
      // prior edge blocks: Ji <- 15x9, Jb <- 15x6 (the IMU Jacobians are no longer needed)
      prior_jac(sm.st[1], prior, sm.err_prior, sm.Ji, sm.Jb);
      huber_rho(sqrt(25.0), sm.chi2_prior, r0, wP);
      jtoj(sm.Ji, 9, 0, 9, sm.info_prior, 15, wP, sm.Ji, 9, 0, 9, sm.CL, 15, 0, 0, true);
      jtoj(sm.Jb, 6, 0, 6, sm.info_prior, 15, wP, sm.Jb, 6, 0, 6, sm.CL, 15, 9, 9, true);
      jtoj(sm.Ji, 9, 0, 9, sm.info_prior, 15, wP, sm.Jb, 6, 0, 6, sm.CL, 15, 0, 9, true);
      for (int a = 0; a < 9; ++a)
        for (int cc = 0; cc < 6; ++cc) sm.CL[(9 + cc) * 15 + a] = sm.CL[a * 15 + 9 + cc];
      // cov_inv -= E C^-1 E^T (JacobiSVD inverse without clamping in the reference, :709-728)
      double* Cinv = sm.H;  // 225 <= 900
      double* T = sm.S;
      if (!dense_inverse(sm.CL, 15, Cinv))
        for (int i = 0; i < 225; ++i) Cinv[i] = nan("");
      for (int a = 0; a < 15; ++a)
        for (int cc = 0; cc < 15; ++cc) {
          double s = 0;
          for (int k = 0; k < 15; ++k) s += sm.CCL[a * 15 + k] * Cinv[k * 15 + cc];
          T[a * 15 + cc] = s;
        }
      for (int a = 0; a < 15; ++a)
        for (int cc = 0; cc < 15; ++cc) {
          double s = 0;
          for (int k = 0; k < 15; ++k) s += T[a * 15 + k] * sm.CCL[cc * 15 + k];
          sm.C[a * 15 + cc] -= s;
        }
    }
    for (int i = 0; i < 225; ++i) R.marg_cov_inv[i] = sm.C[i];
    R.prior_set = 1;
  }
}

}  // namespace vieo

using namespace vieo;

extern "C" {

int vieo_pose_opt_batch_dev(const VieoPoseOptProblem* pbs_dev, int n, const VieoCamera* cam_dev, const double* Xw_dev,
                            const float* obs_dev, const float* inv_sigma2_dev, const uint8_t* flags_dev,
                            VieoPoseOptResult* res_dev, uint8_t* outlier_dev, double* chi2_dev, void* stream) {
  VIEO_ARG(n >= 0, "bad argument");
  if (n == 0) return VIEO_OK;
  VIEO_ARG(pbs_dev && cam_dev && res_dev && outlier_dev && chi2_dev, "null argument");
  static bool attr_set = false;
  if (!attr_set) {
    VIEO_CK(cudaFuncSetAttribute(k_pose_opt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoSmem)));
    attr_set = true;
  }
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

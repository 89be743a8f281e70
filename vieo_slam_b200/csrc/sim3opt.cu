// Optimizer::OptimizeSim3 on the device (src/Optimizer.cc:2689-2920): the loop-closing thread's Sim3 refinement of a
// candidate keyframe pair.  One free VertexNavStatePR (S12 stored as a NavState: mRwb = R12^-1, mpwb = -(mRwb t12),
// :2717-2722) + VertexScale (fixed iff bFixScale) + FIXED points; per match an EdgeReprojectPRS (x1 = pi(S12 X2c), MODE 1)
// and an EdgeReprojectPRSInv (x2 = pi(S12^-1 X1c), MODE 2) of src/Odom/g2otypes.h:321-549, both Huber sqrt(th2);
// optimize(5), outlier pairs removed, optimize(5 | 10), inlier count.
// One thread block per candidate pair runs the whole schedule without leaving the SM (LoopClosing::ComputeSim3 holds
// several candidates: the batch).  The <= 7-dim system is summed in a fixed order (thread-strided partial sums, xor tree,
// warps in sequence) and solved by one thread; the Levenberg-Marquardt loop is g2o's
// (optimization_algorithm_levenberg.cpp:61-189).
#include <algorithm>

#include "ba_edges.cuh"

namespace vieo {

constexpr int kS3Threads = 128, kS3Warps = kS3Threads / 32;
constexpr int kS3Sums = 36;  // 28 upper-triangle entries of the 7 x 7 H, 7 of b, robust chi2

struct S3Smem {
  double red[kS3Warps][kS3Sums];
  double tot[kS3Sums];
  double H[49], b[7], A[49], x[7];
  Mat3 Rwb;   // of the current estimate
  Vec3 p;
  Quat q;
  Vec3 p_bak;
  Quat q_bak;
  double sc, sc_bak, lambda, ni, rho;
  int ok, stop_it;
};

struct S3Edge {
  double e[2], Jp[12], Js[2];
};

// EdgeReproject<2, 6, 3, MODE>::computeError / linearizeOplus (g2otypes.h:346-406, 439-541; USE_P_PLUS_RDP): inverse = false
// MODE 1 (PRS), true MODE 2 (PRSInv)
template <bool kJac>
__device__ __forceinline__ void sim3_edge(const CamK& c, const Mat3& Rwb_in, const Vec3& p_in, double sc, const Vec3& Xh,
                                          const float* __restrict__ obs, bool inverse, S3Edge& o) {
  Mat3 Rwb = Rwb_in;
  Vec3 twb = p_in;
  double sfac = sc;
  if (inverse) {
    sfac = 1. / sfac;
    Rwb = m3_t(Rwb);
    const Vec3 t = m3_mulv(Rwb, twb);
    twb = {-t.x, -t.y, -t.z};
  }
  const Mat3 Rcw = m3_mul(c.Rcb, m3_t(Rwb));
  Vec3 tcw = m3_mulv(Rcw, twb);
  tcw = {-tcw.x + c.tcb.x, -tcw.y + c.tcb.y, -tcw.z + c.tcb.z};
  if (inverse) tcw = {tcw.x * sfac, tcw.y * sfac, tcw.z * sfac};
  const Vec3 Xw = {Xh.x * sfac, Xh.y * sfac, Xh.z * sfac};
  const Vec3 Pc = v3_add(m3_mulv(Rcw, Xw), tcw);
  float u, v;
  double Jc[6];
  cam_project(c, Pc, u, v, kJac ? Jc : nullptr);
  o.e[0] = (double)obs[0] - (double)u;
  o.e[1] = (double)obs[1] - (double)v;
  if (!kJac) return;
  Mat3 Jproj = m3_zero();
#pragma unroll
  for (int i = 0; i < 6; ++i) Jproj.m[i] = -Jc[i];
  Mat3 JdP = m3_mul(Jproj, m3_scale(c.Rcb, -1.0));
  Mat3 JdR;
  if (!inverse) {
    const Vec3 Paux = m3_tmulv(Rwb_in, v3_sub(Xw, p_in));
    JdR = m3_mul(m3_mul(Jproj, c.Rcb), m3_hat(Paux));
  } else {
    JdR = m3_mul(m3_mul(Jproj, m3_scale(Rcw, -1.0)), m3_hat(Xw));
  }
  const Mat3 JX = m3_mul(Jproj, Rcw);
  double Js[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) Js[r] = JX.m[3 * r] * Xh.x + JX.m[3 * r + 1] * Xh.y + JX.m[3 * r + 2] * Xh.z;
  if (inverse) {
    JdP = m3_mul(JdP, m3_scale(Rwb_in, -1.0));  // J_twb_tbw (:524)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double jt = Jproj.m[3 * r] * tcw.x + Jproj.m[3 * r + 1] * tcw.y + Jproj.m[3 * r + 2] * tcw.z;
      Js[r] = (Js[r] + jt) * (-sfac * sfac);  // (:538-539)
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o.Jp[6 * r + k] = JdP.m[3 * r + k];
      o.Jp[6 * r + 3 + k] = JdR.m[3 * r + k];
    }
    o.Js[r] = Js[r];
  }
}

struct S3Ctx {
  CamK cam;
  const double *X1, *X2;
  const float *o1, *o2, *w1, *w2;
  uint8_t* keep;    // vpMatches1[i] != nullptr after the call
  uint8_t* active;  // the pair's edges are still in the graph (removed after the first stage when either chi2 > th2)
  double *c12, *c21;
  int M, n;
  double delta;
};

// computeActiveErrors (+ linearizeOplus / constructQuadraticForm when kBuild) over the active pairs; every thread calls it.
// Result: sm.tot[35] = activeRobustChi2, and with kBuild sm.H / sm.b.
template <bool kBuild>
__device__ void s3_evaluate(const S3Ctx& c, S3Smem& sm) {
  double acc[kS3Sums];
#pragma unroll
  for (int k = 0; k < kS3Sums; ++k) acc[k] = 0;
  const Mat3 Rwb = sm.Rwb;
  const Vec3 p = sm.p;
  const double sc = sm.sc;
  for (int i = threadIdx.x; i < c.M; i += kS3Threads) {
    if (!c.active[i]) continue;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const bool inv = dir == 1;
      const double* Xp = (inv ? c.X1 : c.X2) + 3 * (size_t)i;
      const float* ob = (inv ? c.o2 : c.o1) + 2 * (size_t)i;
      const double w = (double)(inv ? c.w2 : c.w1)[i];
      S3Edge o;
      sim3_edge<kBuild>(c.cam, Rwb, p, sc, {Xp[0], Xp[1], Xp[2]}, ob, inv, o);
      const double chi = o.e[0] * (w * o.e[0]) + o.e[1] * (w * o.e[1]);
      (inv ? c.c21 : c.c12)[i] = chi;
      double r0, r1;
      huber_rho(c.delta, chi, r0, r1);
      acc[35] += r0;
      if (kBuild) {
        double J[2][7];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int a = 0; a < 6; ++a) J[k][a] = o.Jp[6 * k + a];
          J[k][6] = o.Js[k];
        }
        const double ww = r1 * w;
        int q = 0;
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          double sb = 0;
#pragma unroll
          for (int k = 0; k < 2; ++k) sb += J[k][a] * (-(w * o.e[k]) * r1);
          acc[28 + a] += sb;
#pragma unroll
          for (int cc = a; cc < 7; ++cc) {
            double h = 0;
#pragma unroll
            for (int k = 0; k < 2; ++k) h += (J[k][a] * ww) * J[k][cc];
            acc[q++] += h;
          }
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kS3Sums; ++k) {
    if (!kBuild && k < 35) continue;
    double a = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) sm.red[warp][k] = a;
  }
  __syncthreads();
  if (threadIdx.x < kS3Sums && (kBuild || threadIdx.x == 35)) {
    double a = 0;
    for (int w = 0; w < kS3Warps; ++w) a += sm.red[w][threadIdx.x];
    sm.tot[threadIdx.x] = a;
  }
  __syncthreads();
  if (kBuild && threadIdx.x < 56) {  // unpack: 49 entries of H (n x n view used by the solver), 7 of b
    const int n = c.n;
    if (threadIdx.x < 49) {
      const int r = threadIdx.x / 7, cc = threadIdx.x % 7;
      const int lo = min(r, cc), hi = max(r, cc);
      if (r < n && cc < n) sm.H[r * n + cc] = sm.tot[lo * 7 - lo * (lo - 1) / 2 + (hi - lo)];
    } else {
      sm.b[threadIdx.x - 49] = sm.tot[28 + threadIdx.x - 49];
    }
  }
  if (kBuild) __syncthreads();
}

// (H + lambda I) x = b, n <= 7, by thread 0 (Cholesky; LinearSolverDense is an LDLT that must be positive)
__device__ __forceinline__ bool s3_solve(S3Smem& sm, int n) {
  double* A = sm.A;
  for (int t = 0; t < n * n; ++t) A[t] = sm.H[t] + ((t / n == t % n) ? sm.lambda : 0.0);
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0) || !isfinite(d)) return false;
    A[j * n + j] = sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / A[j * n + j];
    }
  }
  double y[7];
  for (int i = 0; i < n; ++i) {
    double s = sm.b[i];
    for (int k = 0; k < i; ++k) s -= A[i * n + k] * y[k];
    y[i] = s / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * sm.x[k];
    sm.x[i] = s / A[i * n + i];
  }
  return true;
}

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg::solve inlined; returns iterations run
__device__ int s3_optimize(const S3Ctx& c, S3Smem& sm, int iterations) {
  const int n = c.n;
  int it_run = 0, nBad = 0;
  if (threadIdx.x < 7) sm.x[threadIdx.x] = 0;
  __syncthreads();
  bool ok = true;
  for (int it = 0; it < iterations && ok; ++it) {
    s3_evaluate<true>(c, sm);
    double currentChi = sm.tot[35];
    const double iniChi = currentChi;
    if (threadIdx.x == 0 && it == 0) {
      double mx = 0;
      for (int j = 0; j < n; ++j) mx = fmax(fabs(sm.H[j * n + j]), mx);
      sm.lambda = 1e-5 * mx;
      sm.ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      __syncthreads();
      if (threadIdx.x == 0) {
        sm.p_bak = sm.p;
        sm.q_bak = sm.q;
        sm.sc_bak = sm.sc;
        sm.ok = s3_solve(sm, n) ? 1 : 0;
        if (sm.ok) {
          NavS s;
          s.p = sm.p;
          s.q = sm.q;
          ns_inc_pr(s, sm.x);
          sm.p = s.p;
          sm.q = s.q;
          sm.Rwb = q_matrix(s.q);
          if (n == 7) sm.sc += sm.x[6];
        }
      }
      __syncthreads();
      s3_evaluate<false>(c, sm);
      double tempChi = sm.tot[35];
      if (!sm.ok) tempChi = 1.7976931348623157e308;
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
          sm.p = sm.p_bak;
          sm.q = sm.q_bak;
          sm.sc = sm.sc_bak;
          sm.Rwb = q_matrix(sm.q);
        }
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    __syncthreads();
    ++it_run;
    if (qmax == 10 || rho == 0) {
      ok = false;
    } else {
      nBad = ((iniChi - currentChi) * 1e3 < iniChi) ? nBad + 1 : 0;
      if (nBad >= 3) ok = false;
    }
  }
  return it_run;
}

__global__ void __launch_bounds__(kS3Threads) k_sim3_opt(const VieoSim3Problem* __restrict__ pbs, const VieoCamera* __restrict__ camp,
                                                         const double* __restrict__ Xc1, const double* __restrict__ Xc2,
                                                         const float* __restrict__ obs1, const float* __restrict__ obs2,
                                                         const float* __restrict__ w1, const float* __restrict__ w2,
                                                         VieoSim3Result* __restrict__ res, uint8_t* __restrict__ keep,
                                                         double* __restrict__ chi2_12, double* __restrict__ chi2_21,
                                                         uint8_t* __restrict__ active_scratch) {
  __shared__ S3Smem sm;
  const VieoSim3Problem& pb = pbs[blockIdx.x];
  VieoSim3Result& R = res[blockIdx.x];
  const int b0 = pb.m_begin, M = pb.m_end - pb.m_begin;
  S3Ctx c;
  c.cam = cam_load(*camp);
  c.X1 = Xc1 + 3 * (size_t)b0; c.X2 = Xc2 + 3 * (size_t)b0;
  c.o1 = obs1 + 2 * (size_t)b0; c.o2 = obs2 + 2 * (size_t)b0;
  c.w1 = w1 + b0; c.w2 = w2 + b0;
  c.keep = keep + b0;
  c.active = active_scratch + b0;
  c.c12 = chi2_12 + b0; c.c21 = chi2_21 + b0;
  c.M = M;
  c.n = pb.fix_scale ? 6 : 7;
  c.delta = (double)sqrtf(pb.th2);  // const float deltaHuber = sqrt(th2) (:2752)
  const double th2 = (double)pb.th2;
  for (int i = threadIdx.x; i < M; i += kS3Threads) {
    c.keep[i] = 1;
    c.active[i] = 1;
    c.c12[i] = 0;
    c.c21[i] = 0;
  }
  if (threadIdx.x == 0) {
    const NavS s = ns_load(pb.ns);
    sm.p = s.p;
    sm.q = s.q;
    sm.Rwb = q_matrix(s.q);
    sm.sc = pb.scale;
    sm.lambda = 0;
    sm.ni = 2;
    sm.ok = 1;
  }
  __syncthreads();
  int iters = s3_optimize(c, sm, 5);
  // outlier pairs leave the graph (:2858-2868)
  int bad = 0;
  for (int i = threadIdx.x; i < M; i += kS3Threads)
    if (c.c12[i] > th2 || c.c21[i] > th2) {
      c.keep[i] = 0;
      c.active[i] = 0;
      ++bad;
    }
  __syncthreads();  // the masks are visible to every thread
  __shared__ int s_cnt[kS3Warps];
  {
    int v = bad;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = v;
    __syncthreads();
    bad = 0;
    for (int w = 0; w < kS3Warps; ++w) bad += s_cnt[w];
    __syncthreads();
  }
  const int nBad = bad;
  if (M - nBad < 10) {  // (:2878): return 0, the Sim3 estimate is NOT written back
    if (threadIdx.x == 0) {
      R.ns = pb.ns;
      R.scale = pb.scale;
      R.chi2_final = 0;
      R.lambda_final = sm.lambda;
      R.n_inliers = 0;
      R.n_corr = M;
      R.n_bad = nBad;
      R.iterations = iters;
    }
    return;
  }
  iters += s3_optimize(c, sm, nBad > 0 ? 10 : 5);
  int nin = 0;
  double rsum = 0;
  for (int i = threadIdx.x; i < M; i += kS3Threads) {
    if (!c.active[i]) continue;
    if (c.c12[i] > th2 || c.c21[i] > th2) c.keep[i] = 0;
    else ++nin;
    double r0, r1;
    huber_rho(c.delta, c.c12[i], r0, r1);
    rsum += r0;
    huber_rho(c.delta, c.c21[i], r0, r1);
    rsum += r0;
  }
  {
    int v = nin;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(0xffffffffu, v, o);
      rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_cnt[threadIdx.x >> 5] = v;
      sm.red[threadIdx.x >> 5][0] = rsum;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int tot = 0;
    double chi = 0;
    for (int w = 0; w < kS3Warps; ++w) {
      tot += s_cnt[w];
      chi += sm.red[w][0];
    }
    NavS s = ns_load(pb.ns);
    s.p = sm.p;
    s.q = sm.q;
    ns_store(s, R.ns);
    R.scale = sm.sc;
    R.chi2_final = chi;
    R.lambda_final = sm.lambda;
    R.n_inliers = tot;
    R.n_corr = M;
    R.n_bad = nBad;
    R.iterations = iters;
  }
}

}  // namespace vieo

using namespace vieo;

extern "C" {

int vieo_optimize_sim3_batch_dev(const VieoSim3Problem* pbs_dev, int n, const VieoCamera* cam_dev, const double* Xc1_dev,
                                 const double* Xc2_dev, const float* obs1_dev, const float* obs2_dev,
                                 const float* inv_sigma2_1_dev, const float* inv_sigma2_2_dev, VieoSim3Result* res_dev,
                                 uint8_t* keep_dev, double* chi2_12_dev, double* chi2_21_dev, uint8_t* scratch_dev, void* stream) {
  VIEO_ARG(n >= 0, "bad argument");
  if (n == 0) return VIEO_OK;
  VIEO_ARG(pbs_dev && cam_dev && res_dev && keep_dev && chi2_12_dev && chi2_21_dev && scratch_dev, "null argument");
  k_sim3_opt<<<n, kS3Threads, 0, (cudaStream_t)stream>>>(pbs_dev, cam_dev, Xc1_dev, Xc2_dev, obs1_dev, obs2_dev, inv_sigma2_1_dev,
                                                         inv_sigma2_2_dev, res_dev, keep_dev, chi2_12_dev, chi2_21_dev, scratch_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_optimize_sim3_batch(const VieoSim3Problem* pbs, int n, const VieoCamera* cam, const double* Xc1, const double* Xc2,
                             const float* obs1, const float* obs2, const float* inv_sigma2_1, const float* inv_sigma2_2,
                             int n_matches, VieoSim3Result* res, uint8_t* keep, double* chi2_12, double* chi2_21, int device) {
  VIEO_ARG(n >= 0 && n_matches >= 0, "bad argument");
  if (n == 0) return VIEO_OK;
  VIEO_ARG(pbs && cam && res && (n_matches == 0 || (Xc1 && Xc2 && obs1 && obs2 && inv_sigma2_1 && inv_sigma2_2 && keep)),
           "null argument");
  VIEO_ARG(cam->model >= 0 && cam->model <= 2 && cam->num_k >= 0 && cam->num_k <= 6, "unsupported camera model");
  for (int k = 0; k < n; ++k) {
    VIEO_ARG(pbs[k].m_begin >= 0 && pbs[k].m_end >= pbs[k].m_begin && pbs[k].m_end <= n_matches, "match range");
    VIEO_ARG(pbs[k].th2 > 0 && pbs[k].scale > 0, "bad th2 / scale");
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  const size_t nm = std::max(n_matches, 1);
  // one staging block: [problems | camera | Xc1 | Xc2 | obs1 | obs2 | w1 | w2] in, [results | chi2_12 | chi2_21 | keep] out
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  const size_t o_pb = take(sizeof(VieoSim3Problem) * n), o_cam = take(sizeof(VieoCamera)), o_x1 = take(24 * nm),
               o_x2 = take(24 * nm), o_o1 = take(8 * nm), o_o2 = take(8 * nm), o_w1 = take(4 * nm), o_w2 = take(4 * nm);
  const size_t in_bytes = off;
  const size_t o_res = take(sizeof(VieoSim3Result) * n), o_c12 = take(8 * nm), o_c21 = take(8 * nm), o_keep = take(nm);
  const size_t out_end = off;
  const size_t o_act = take(nm);
  uint8_t* d = (uint8_t*)cs->get(0, off);
  uint8_t* h = (uint8_t*)cs->get_pinned(out_end);
  if (!d || !h) return VIEO_E_CUDA;
  auto put = [&](size_t o, const void* src, size_t bytes) { if (src && bytes) memcpy(h + o, src, bytes); };
  put(o_pb, pbs, sizeof(VieoSim3Problem) * n); put(o_cam, cam, sizeof(VieoCamera));
  put(o_x1, Xc1, 24 * (size_t)n_matches); put(o_x2, Xc2, 24 * (size_t)n_matches);
  put(o_o1, obs1, 8 * (size_t)n_matches); put(o_o2, obs2, 8 * (size_t)n_matches);
  put(o_w1, inv_sigma2_1, 4 * (size_t)n_matches); put(o_w2, inv_sigma2_2, 4 * (size_t)n_matches);
  cudaStream_t st = cs->st;
  VIEO_CK(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
  rc = vieo_optimize_sim3_batch_dev((const VieoSim3Problem*)(d + o_pb), n, (const VieoCamera*)(d + o_cam), (const double*)(d + o_x1),
                                    (const double*)(d + o_x2), (const float*)(d + o_o1), (const float*)(d + o_o2),
                                    (const float*)(d + o_w1), (const float*)(d + o_w2), (VieoSim3Result*)(d + o_res), d + o_keep,
                                    (double*)(d + o_c12), (double*)(d + o_c21), d + o_act, st);
  if (rc) return rc;
  VIEO_CK(cudaMemcpyAsync(h + o_res, d + o_res, out_end - o_res, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  memcpy(res, h + o_res, sizeof(VieoSim3Result) * n);
  if (n_matches) {
    memcpy(keep, h + o_keep, (size_t)n_matches);
    if (chi2_12) memcpy(chi2_12, h + o_c12, 8 * (size_t)n_matches);
    if (chi2_21) memcpy(chi2_21, h + o_c21, 8 * (size_t)n_matches);
  }
  return VIEO_OK;
}

}  // extern "C"

// On-manifold IMU pre-integration on sm_100a — replaces IMUPreIntegratorBase::PreIntegration / update
// (src/Odom/OdomPreIntegrator.h:227-506).  One warp per interval [t_i, t_j]: the sample recurrence is
// sequential, so lane 0 carries the small state (dR, dv, dp, five 3x3 bias Jacobians) while all 32 lanes share
// the two 9x9 covariance propagations (P-R-V and P-V-R orderings) held in shared memory.  fp64 throughout;
// latency-bound by construction: batches of intervals fill the machine.
#include <vector>

#include "common.cuh"
#include "so3.cuh"

namespace vieo {

constexpr int kImuWarps = 4;

struct ImuWarpSmem {
  double S[2][81];   // covariances: [0] = P-R-V, [1] = P-V-R
  double T[2][81];
  double A[2][81];
  double Bg[2][27], Ba[2][27];
  double sg, sa;
};

// all lanes: S <- A S A^T + Bg (sg I) Bg^T + Ba (sa I) Ba^T for both orderings
__device__ __forceinline__ void imu_propagate(ImuWarpSmem& W, int lane) {
  for (int e = lane; e < 162; e += 32) {
    const int o = e / 81, ij = e - 81 * o, i = ij / 9, j = ij - 9 * i;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s += W.A[o][9 * i + k] * W.S[o][9 * k + j];
    W.T[o][ij] = s;
  }
  __syncwarp();
  for (int e = lane; e < 162; e += 32) {
    const int o = e / 81, ij = e - 81 * o, i = ij / 9, j = ij - 9 * i;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s += W.T[o][9 * i + k] * W.A[o][9 * j + k];
    double g = 0, a = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      g += (W.Bg[o][3 * i + k] * W.sg) * W.Bg[o][3 * j + k];
      a += (W.Ba[o][3 * i + k] * W.sa) * W.Ba[o][3 * j + k];
    }
    W.S[o][ij] = s + g + a;
  }
  __syncwarp();
}

struct ImuState {  // lane 0 only
  Mat3 R, Jgp, Jap, Jgv, Jav, JgR;
  Vec3 v, p;
  double dt;
};

__device__ __forceinline__ void put_block(double* A, int r, int c, const Mat3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[9 * (r + i) + c + j] = b.m[3 * i + j];
}
__device__ __forceinline__ void put_rows(double* B, int r, const Mat3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) B[3 * (r + i) + j] = b.m[3 * i + j];
}

// IMUPreIntegratorBase::update (OdomPreIntegrator.h:432-506)
__device__ void imu_update(ImuState& s, ImuWarpSmem& W, int lane, Vec3 omega, Vec3 acc, double dt,
                           const VieoImuNoise& nz) {
  for (int e = lane; e < 162; e += 32) {
    const int o = e / 81, ij = e - 81 * o;
    W.A[o][ij] = (ij % 10 == 0) ? 1.0 : 0.0;
  }
  for (int e = lane; e < 54; e += 32) {
    (&W.Bg[0][0])[e] = 0.0;
    (&W.Ba[0][0])[e] = 0.0;
  }
  __syncwarp();
  const double dt2div2 = dt * dt / 2;
  Mat3 dRt, Jrdt;
  if (lane == 0) {
    const Vec3 wdt = v3_scale(omega, dt);
    const Mat3 dR = so3_Exp(wdt), Jr = so3_Jr(wdt), skewa = m3_hat(acc);
    double sg, sa;
    if (nz.dt_cov_noise_fixed) {
      sg = nz.sigma_g;
      sa = nz.sigma_a;
    } else if (!nz.freq_ref || dt < 1.5 / nz.freq_ref) {
      sg = nz.sigma_g / dt;
      sa = nz.sigma_a / dt;
    } else {
      sg = nz.sigma_g * nz.freq_ref;
      sa = nz.sigma_a * nz.freq_ref;
    }
    W.sg = sg;
    W.sa = sa;
    const Mat3 Rsk = m3_mul(s.R, skewa), nRsk = m3_scale(Rsk, -1.0);
    const Mat3 nRsk_dt = m3_scale(nRsk, dt), nRsk_dt2 = m3_scale(nRsk, dt2div2);
    dRt = m3_t(dR);
    Jrdt = m3_scale(Jr, dt);
    const Mat3 Idt = m3_scale(m3_identity(), dt), Rdt = m3_scale(s.R, dt), Rdt2 = m3_scale(s.R, dt2div2);
    // P-R-V (:444-452)
    put_block(W.A[0], 3, 3, dRt);
    put_block(W.A[0], 6, 3, nRsk_dt);
    put_block(W.A[0], 0, 3, nRsk_dt2);
    put_block(W.A[0], 0, 6, Idt);
    put_rows(W.Bg[0], 3, Jrdt);
    put_rows(W.Ba[0], 6, Rdt);
    put_rows(W.Ba[0], 0, Rdt2);
    // P-V-R (:465-474)
    put_block(W.A[1], 6, 6, dRt);
    put_block(W.A[1], 3, 6, nRsk_dt);
    put_block(W.A[1], 0, 6, nRsk_dt2);
    put_block(W.A[1], 0, 3, Idt);
    put_rows(W.Bg[1], 6, Jrdt);
    put_rows(W.Ba[1], 3, Rdt);
    put_rows(W.Ba[1], 0, Rdt2);
    // bias Jacobians: P, then V, then R, all with the old delta-R (:488-493)
    const Mat3 RskJgR = m3_mul(Rsk, s.JgR);
    s.Jap = m3_add(s.Jap, m3_sub(m3_scale(s.Jav, dt), m3_scale(s.R, dt2div2)));
    s.Jgp = m3_add(s.Jgp, m3_sub(m3_scale(s.Jgv, dt), m3_scale(RskJgR, dt2div2)));
    s.Jav = m3_add(s.Jav, m3_scale(m3_scale(s.R, -1.0), dt));
    s.Jgv = m3_add(s.Jgv, m3_scale(m3_scale(RskJgR, -1.0), dt));
    s.JgR = m3_sub(m3_mul(dRt, s.JgR), Jrdt);
    // delta measurements (:497-503)
    const Vec3 Ra2 = m3_mulv(s.R, v3_scale(acc, dt2div2)), Ra1 = m3_mulv(s.R, v3_scale(acc, dt));
    s.p = {s.p.x + (s.v.x * dt + Ra2.x), s.p.y + (s.v.y * dt + Ra2.y), s.p.z + (s.v.z * dt + Ra2.z)};
    s.v = v3_add(s.v, Ra1);
    s.R = so3_normalize(m3_mul(s.R, dR));
    s.dt += dt;
  }
  __syncwarp();
  imu_propagate(W, lane);
}

__global__ void __launch_bounds__(kImuWarps * 32) k_imu_preint(const double* __restrict__ smp,
                                                              const int* __restrict__ seg_ptr,
                                                              const double* __restrict__ ti_tj,
                                                              const double* __restrict__ bg_ba, VieoImuNoise nz,
                                                              int n_int, VieoImuPreint* __restrict__ out) {
  __shared__ ImuWarpSmem s_w[kImuWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = blockIdx.x * kImuWarps + warp;
  if (it >= n_int) return;
  ImuWarpSmem& W = s_w[warp];
  for (int e = lane; e < 162; e += 32) (&W.S[0][0])[e] = 0.0;
  __syncwarp();
  ImuState s;
  s.R = m3_identity();
  s.Jgp = s.Jap = s.Jgv = s.Jav = s.JgR = m3_zero();
  s.v = s.p = {0, 0, 0};
  s.dt = 0;
  const double* S = smp + 7 * (size_t)seg_ptr[it];
  const int n = seg_ptr[it + 1] - seg_ptr[it];
  const double ti = ti_tj[2 * it], tj = ti_tj[2 * it + 1];
  const Vec3 bg = {bg_ba[6 * it], bg_ba[6 * it + 1], bg_ba[6 * it + 2]};
  const Vec3 ba = {bg_ba[6 * it + 3], bg_ba[6 * it + 4], bg_ba[6 * it + 5]};
  int status = 0;
  if (n > 0) {
    // sample window selection (OdomPreIntegrator.h:237-262); END = one past the last sample
    const int END = n;
    const bool back = ti > tj;
    const double tmin = back ? tj : ti, tmax = back ? ti : tj;
    int start = 0, stop = END;
    for (int j = 0; j != END && S[7 * j] <= tmin; start = j++) {
    }
    for (int j = END; j != 0;) {
      stop = j--;
      if (S[7 * j] >= tmax) continue;
      break;
    }
    if (back) {
      if (stop == END) --stop;
      const int t = start;
      start = stop;
      stop = t;
      if (S[7 * stop] > tmin) stop = END;
    }
    for (int j = start; j != stop;) {
      const int jm1 = j;
      if (back) {
        if (j == 0) j = stop; else --j;
      } else
        ++j;
      const double tj_1 = jm1 == start ? ti : S[7 * jm1];
      const double tjj = j == stop ? tj : S[7 * j];
      double dt = tjj - tj_1;
      if (dt == 0) continue;
      if (fabs(dt) > 1.5) {  // "CheckIMU!!!" (:289-293)
        s.dt = 0;
        status = -1;
        break;
      }
      Vec3 a0 = {S[7 * jm1 + 1], S[7 * jm1 + 2], S[7 * jm1 + 3]}, w0 = {S[7 * jm1 + 4], S[7 * jm1 + 5], S[7 * jm1 + 6]};
      Vec3 a1 = a0, w1 = w0;
      double t1 = S[7 * jm1];
      if (j != END) {
        a1 = {S[7 * j + 1], S[7 * j + 2], S[7 * j + 3]};
        w1 = {S[7 * j + 4], S[7 * j + 5], S[7 * j + 6]};
        t1 = S[7 * j];
        if (j == stop) {  // interpolate the last sample back to t_j (:301-311)
          const double d = S[7 * j] - tj;
          if (back ? d < 0 : d > 0) {
            const double rat = d / (S[7 * j] - S[7 * jm1]);
            w1 = v3_add(v3_scale(w0, rat), v3_scale(w1, 1 - rat));
            a1 = v3_add(v3_scale(a0, rat), v3_scale(a1, 1 - rat));
          }
        }
        if (jm1 == start) {  // interpolate the first sample forward to t_i (:312-326)
          const double d = ti - S[7 * jm1];
          if (back ? d < 0 : d > 0) {
            const double rat = d / (S[7 * j] - S[7 * jm1]);
            w0 = v3_add(v3_scale(w0, 1 - rat), v3_scale(w1, rat));
            a0 = v3_add(v3_scale(a0, 1 - rat), v3_scale(a1, rat));
          }
        }
      }
      if (jm1 == start) {  // first sample later than t_i: hold it over the gap (:403-410)
        const double dc = S[7 * jm1] - ti;
        if (back ? dc < 0 : dc > 0) {
          imu_update(s, W, lane, v3_sub(w0, bg), v3_sub(a0, ba), dc, nz);
          dt -= dc;
          if (!dt) continue;
        }
      }
      double dcs = 0;
      if (j == stop) {
        dcs = tj - t1;
        if (back ? dcs < 0 : dcs > 0) dt -= dcs;
      }
      const Vec3 wm = v3_scale(v3_add(w1, w0), 0.5), am = v3_scale(v3_add(a1, a0), 0.5);  // mid-point (:419)
      imu_update(s, W, lane, v3_sub(wm, bg), v3_sub(am, ba), dt, nz);
      if (back ? dcs < 0 : dcs > 0) imu_update(s, W, lane, v3_sub(w1, bg), v3_sub(a1, ba), dcs, nz);
    }
  }
  VieoImuPreint& o = out[it];
  for (int e = lane; e < 81; e += 32) {
    o.SigmaPRV[e] = W.S[0][e];
    o.SigmaPVR[e] = W.S[1][e];
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      o.Rij[e] = s.R.m[e];
      o.Jgp[e] = s.Jgp.m[e];
      o.Jap[e] = s.Jap.m[e];
      o.Jgv[e] = s.Jgv.m[e];
      o.Jav[e] = s.Jav.m[e];
      o.JgR[e] = s.JgR.m[e];
    }
    o.vij[0] = s.v.x; o.vij[1] = s.v.y; o.vij[2] = s.v.z;
    o.pij[0] = s.p.x; o.pij[1] = s.p.y; o.pij[2] = s.p.z;
    o.dt = s.dt;
    o.status = status;
    o.pad_ = 0;
  }
}


// ------------------------------------------------------------------------------------------------
// Optimizer::OptimizeInitialGyroBias (include/Optimizer.h:819-892) with EdgeGyrBias (src/Odom/g2otypes.h:940-973): one
// Gauss-Newton iteration on the 3-dim gyro-bias vertex seeded with zero.  One CTA: a thread per keyframe pair evaluates
// the residual Log((dRij Exp(JgR bg))^T Rwbi^T Rwbj) and its Jacobian at bg = 0, J^T W J / J^T W e are summed in a fixed
// order (thread-strided partial sums, then a shared-memory tree), thread 0 solves the 3x3 system, and every thread then
// writes the linearisation biases {bg + dbg, ba_i} of the re-integration that follows (IMUInitialization.cpp:640-648).
constexpr int kGyroThreads = 256;
struct GyroBiasOut {
  double dbg[3];
  double bg[3];
  int32_t num_equations;
  int32_t solved;
};
__global__ void __launch_bounds__(kGyroThreads) k_gyro_bias(const VieoImuPreint* __restrict__ pre,
                                                            const double* __restrict__ Rwb, int n_kf, int use_info,
                                                            double bg0x, double bg0y, double bg0z,
                                                            const double* __restrict__ bg_ba_in,
                                                            double* __restrict__ bg_ba_out, GyroBiasOut* __restrict__ out) {
  __shared__ double red[kGyroThreads][13];
  __shared__ double s_bg[3];
  const int tid = threadIdx.x;
  double acc[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) acc[k] = 0;
  for (int i = 1 + tid; i < n_kf; i += kGyroThreads) {
    const VieoImuPreint& P = pre[i];
    if (P.dt == 0) continue;
    Mat3 dR, JgR, Ri, Rj;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      dR.m[k] = P.Rij[k];
      JgR.m[k] = P.JgR[k];
      Ri.m[k] = Rwb[9 * (size_t)(i - 1) + k];
      Rj.m[k] = Rwb[9 * (size_t)i + k];
    }
    const Vec3 Jb = m3_mulv(JgR, Vec3{0, 0, 0});
    const Mat3 E = m3_mul(m3_mul(m3_t(m3_mul(dR, so3_Exp(Jb))), m3_t(Ri)), Rj);
    const Vec3 e = so3_Log(E);
    const Mat3 J = m3_scale(m3_mul(m3_mul(m3_mul(so3_JrInv(e), so3_Exp(Vec3{-e.x, -e.y, -e.z})), so3_Jr(Jb)), JgR), -1.0);
    Mat3 W = m3_identity();
    if (use_info) {
      double a[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) a[3 * r + c] = P.SigmaPRV[9 * (3 + r) + 3 + c];
      const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
      const double det = a[0] * c00 + a[1] * c01 + a[2] * c02, id = 1.0 / det;
      W = {{c00 * id, (a[2] * a[7] - a[1] * a[8]) * id, (a[1] * a[5] - a[2] * a[4]) * id,
            c01 * id, (a[0] * a[8] - a[2] * a[6]) * id, (a[2] * a[3] - a[0] * a[5]) * id,
            c02 * id, (a[1] * a[6] - a[0] * a[7]) * id, (a[0] * a[4] - a[1] * a[3]) * id}};
    }
    const Mat3 JtW = m3_mul(m3_t(J), W);
    const Mat3 JtWJ = m3_mul(JtW, J);
    const Vec3 JtWe = m3_mulv(JtW, e);
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] += JtWJ.m[k];
    acc[9] -= JtWe.x;
    acc[10] -= JtWe.y;
    acc[11] -= JtWe.z;
    acc[12] += 1.0;
  }
#pragma unroll
  for (int k = 0; k < 13; ++k) red[tid][k] = acc[k];
  __syncthreads();
  for (int s = kGyroThreads / 2; s > 0; s >>= 1) {
    if (tid < s)
#pragma unroll
      for (int k = 0; k < 13; ++k) red[tid][k] += red[tid + s][k];
    __syncthreads();
  }
  if (tid == 0) {
    const double* H = red[0];
    const double* b = red[0] + 9;
    double x[3] = {0, 0, 0};
    const int neq = (int)red[0][12];
    int solved = 0;
    if (neq >= 1) {
      // 3x3 Cholesky (the reference: LinearSolverEigen's LDLT; an indefinite H leaves the estimate at zero)
      double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      bool ok = true;
      for (int j = 0; j < 3 && ok; ++j) {
        double d = H[3 * j + j];
        for (int k = 0; k < j; ++k) d -= L[3 * j + k] * L[3 * j + k];
        if (!(d > 0)) {
          ok = false;
          break;
        }
        L[3 * j + j] = sqrt(d);
        for (int r = j + 1; r < 3; ++r) {
          double v = H[3 * r + j];
          for (int k = 0; k < j; ++k) v -= L[3 * r + k] * L[3 * j + k];
          L[3 * r + j] = v / L[3 * j + j];
        }
      }
      if (ok) {
        double y[3];
        for (int r = 0; r < 3; ++r) {
          double v = b[r];
          for (int k = 0; k < r; ++k) v -= L[3 * r + k] * y[k];
          y[r] = v / L[3 * r + r];
        }
        for (int r = 2; r >= 0; --r) {
          double v = y[r];
          for (int k = r + 1; k < 3; ++k) v -= L[3 * k + r] * x[k];
          x[r] = v / L[3 * r + r];
        }
        solved = 1;
      }
    }
    out->dbg[0] = x[0]; out->dbg[1] = x[1]; out->dbg[2] = x[2];
    s_bg[0] = bg0x + x[0]; s_bg[1] = bg0y + x[1]; s_bg[2] = bg0z + x[2];
    out->bg[0] = s_bg[0]; out->bg[1] = s_bg[1]; out->bg[2] = s_bg[2];
    out->num_equations = neq;
    out->solved = solved;
  }
  __syncthreads();
  if (bg_ba_out)
    for (int i = tid; i < n_kf; i += kGyroThreads) {
      bg_ba_out[6 * (size_t)i + 0] = s_bg[0];
      bg_ba_out[6 * (size_t)i + 1] = s_bg[1];
      bg_ba_out[6 * (size_t)i + 2] = s_bg[2];
#pragma unroll
      for (int k = 3; k < 6; ++k) bg_ba_out[6 * (size_t)i + k] = bg_ba_in ? bg_ba_in[6 * (size_t)i + k] : 0.0;
    }
}

}  // namespace vieo

using namespace vieo;

extern "C" {

void vieo_imu_set_param(VieoImuNoise* nz, const double sigma2[4], int dt_cov_noise_fixed, double freq_ref) {
  nz->sigma_g = sigma2[0];
  nz->sigma_a = sigma2[1];
  nz->sigma_bg = sigma2[2];
  nz->sigma_ba = sigma2[3];
  nz->dt_cov_noise_fixed = dt_cov_noise_fixed;
  nz->pad_ = 0;
  if (dt_cov_noise_fixed && freq_ref) {
    nz->freq_ref = 0;
    nz->sigma_g *= freq_ref;
    nz->sigma_a *= freq_ref;
  } else
    nz->freq_ref = freq_ref;
}

int vieo_imu_preint_batch_dev(const double* samples_dev, const int32_t* seg_ptr_dev, const double* ti_tj_dev,
                              const double* bg_ba_dev, const VieoImuNoise* noise, int n_intervals,
                              VieoImuPreint* out_dev, void* stream) {
  VIEO_ARG(noise && n_intervals >= 0, "bad argument");
  if (n_intervals == 0) return VIEO_OK;
  VIEO_ARG(samples_dev && seg_ptr_dev && ti_tj_dev && bg_ba_dev && out_dev, "null argument");
  k_imu_preint<<<(n_intervals + kImuWarps - 1) / kImuWarps, kImuWarps * 32, 0, (cudaStream_t)stream>>>(
      samples_dev, seg_ptr_dev, ti_tj_dev, bg_ba_dev, *noise, n_intervals, out_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_imu_preint_batch(const double* samples, const int32_t* seg_ptr, const double* ti_tj, const double* bg_ba,
                          const VieoImuNoise* noise, int n_intervals, VieoImuPreint* out, int device) {
  VIEO_ARG(noise && n_intervals >= 0 && seg_ptr, "bad argument");
  if (n_intervals == 0) return VIEO_OK;
  VIEO_ARG(ti_tj && bg_ba && out, "null argument");
  const int n_s = seg_ptr[n_intervals];
  VIEO_ARG(n_s >= 0 && (n_s == 0 || samples), "bad sample list");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  double* d_s = (double*)cs->get(0, sizeof(double) * 7 * std::max(n_s, 1));
  int* d_p = (int*)cs->get(1, sizeof(int) * (n_intervals + 1));
  double* d_t = (double*)cs->get(2, sizeof(double) * 2 * n_intervals);
  double* d_b = (double*)cs->get(3, sizeof(double) * 6 * n_intervals);
  VieoImuPreint* d_o = (VieoImuPreint*)cs->get(4, sizeof(VieoImuPreint) * n_intervals);
  if (!d_s || !d_p || !d_t || !d_b || !d_o) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  if (n_s) VIEO_CK(cudaMemcpyAsync(d_s, samples, sizeof(double) * 7 * n_s, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(d_p, seg_ptr, sizeof(int) * (n_intervals + 1), cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(d_t, ti_tj, sizeof(double) * 2 * n_intervals, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(d_b, bg_ba, sizeof(double) * 6 * n_intervals, cudaMemcpyHostToDevice, st));
  rc = vieo_imu_preint_batch_dev(d_s, d_p, d_t, d_b, noise, n_intervals, d_o, st);
  if (rc) return rc;
  VIEO_CK(cudaMemcpyAsync(out, d_o, sizeof(VieoImuPreint) * n_intervals, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

int vieo_gyro_bias_init_dev(const VieoImuPreint* pre_dev, const double* Rwb_dev, int n_kf, int use_info, const double bg[3],
                            const double* bg_ba_in_dev, double* bg_ba_out_dev, void* result_dev, void* stream) {
  VIEO_ARG(n_kf >= 0 && bg && result_dev, "bad argument");
  VIEO_ARG(n_kf == 0 || (pre_dev && Rwb_dev), "null argument");
  k_gyro_bias<<<1, kGyroThreads, 0, (cudaStream_t)stream>>>(pre_dev, Rwb_dev, n_kf, use_info, bg[0], bg[1], bg[2], bg_ba_in_dev,
                                                           bg_ba_out_dev, (GyroBiasOut*)result_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_imu_init_gyro_bias(const VieoImuPreint* pre, const double* Rwb, int n_kf, int use_info, double bg[3],
                            int* num_equations, const double* samples, const int32_t* seg_ptr, const double* ti_tj,
                            const double* ba, const VieoImuNoise* noise, VieoImuPreint* preint_out, int device) {
  VIEO_ARG(n_kf >= 0 && bg && num_equations, "bad argument");
  *num_equations = 0;
  if (n_kf == 0) return VIEO_OK;
  VIEO_ARG(pre && Rwb, "null argument");
  const bool reint = preint_out != nullptr;
  int n_s = 0;
  if (reint) {
    VIEO_ARG(seg_ptr && ti_tj && noise, "re-integration needs the sample lists");
    n_s = seg_ptr[n_kf];
    VIEO_ARG(n_s >= 0 && (n_s == 0 || samples), "bad sample list");
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  VieoImuPreint* d_pre = (VieoImuPreint*)cs->get(0, sizeof(VieoImuPreint) * n_kf);
  double* d_R = (double*)cs->get(1, sizeof(double) * 9 * n_kf);
  double* d_bi = (double*)cs->get(2, sizeof(double) * 6 * n_kf);
  double* d_bo = (double*)cs->get(3, sizeof(double) * 6 * n_kf);
  GyroBiasOut* d_res = (GyroBiasOut*)cs->get(4, sizeof(GyroBiasOut));
  double* d_s = (double*)cs->get(5, sizeof(double) * 7 * std::max(n_s, 1));
  int* d_p = (int*)cs->get(6, sizeof(int) * (n_kf + 1));
  double* d_t = (double*)cs->get(7, sizeof(double) * 2 * n_kf);
  VieoImuPreint* d_o = (VieoImuPreint*)cs->get(8, sizeof(VieoImuPreint) * n_kf);
  if (!d_pre || !d_R || !d_bi || !d_bo || !d_res || !d_s || !d_p || !d_t || !d_o) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  VIEO_CK(cudaMemcpyAsync(d_pre, pre, sizeof(VieoImuPreint) * n_kf, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(d_R, Rwb, sizeof(double) * 9 * n_kf, cudaMemcpyHostToDevice, st));
  if (reint) {
    // ba of the previous keyframe is kept (IMUInitialization.h:225-232); only bg changes
    std::vector<double> bi(6 * (size_t)n_kf, 0.0);
    if (ba)
      for (int i = 0; i < n_kf; ++i)
        for (int k = 0; k < 3; ++k) bi[6 * (size_t)i + 3 + k] = ba[3 * (size_t)i + k];
    VIEO_CK(cudaMemcpyAsync(d_bi, bi.data(), sizeof(double) * 6 * n_kf, cudaMemcpyHostToDevice, st));
    VIEO_CK(cudaStreamSynchronize(st));  // bi is a local
    if (n_s) VIEO_CK(cudaMemcpyAsync(d_s, samples, sizeof(double) * 7 * n_s, cudaMemcpyHostToDevice, st));
    VIEO_CK(cudaMemcpyAsync(d_p, seg_ptr, sizeof(int) * (n_kf + 1), cudaMemcpyHostToDevice, st));
    VIEO_CK(cudaMemcpyAsync(d_t, ti_tj, sizeof(double) * 2 * n_kf, cudaMemcpyHostToDevice, st));
  }
  rc = vieo_gyro_bias_init_dev(d_pre, d_R, n_kf, use_info, bg, reint ? d_bi : nullptr, reint ? d_bo : nullptr, d_res, st);
  if (rc) return rc;
  if (reint) {
    rc = vieo_imu_preint_batch_dev(d_s, d_p, d_t, d_bo, noise, n_kf, d_o, st);
    if (rc) return rc;
    VIEO_CK(cudaMemcpyAsync(preint_out, d_o, sizeof(VieoImuPreint) * n_kf, cudaMemcpyDeviceToHost, st));
  }
  GyroBiasOut res;
  VIEO_CK(cudaMemcpyAsync(&res, d_res, sizeof(res), cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  *num_equations = res.num_equations;
  if (res.num_equations >= 1) {  // the reference returns before bg += estimate when there is no equation
    bg[0] = res.bg[0]; bg[1] = res.bg[1]; bg[2] = res.bg[2];
  }
  return VIEO_OK;
}

}  // extern "C"

// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  CPU restatement of the on-manifold IMU
// pre-integrator: IMUPreIntegratorBase::PreIntegration / update (src/Odom/OdomPreIntegrator.h:227-506),
// the SO(3) helpers of common/so3_extra.h:121-288 and IMUDataBase::SetParam (src/Odom/OdomData.h:41-56).
// Eigen / Sophus are absent here ("parity unpinned" for their last-bit rounding): quaternion<->matrix conversions
// follow Eigen 3.3.7's published formulas.  Pinned by the reference's own PreIntegration / update() / so3_extra.h compiled unchanged
// against a stand-in for the two libraries (oracle/_ref, tests/test_oracle_ref.py: same update() calls bit for bit, recurrence to
// 1e-12) and by closed-form constant-rate answers.
#include <cmath>
#include <vector>
#include <cstring>

#include "oracle.h"

#include "so3_oracle.h"
namespace {
using namespace orc;

struct Preint {
  M3 R, Jgp, Jap, Jgv, Jav, JgR;
  double v[3], p[3], S_prv[81], S_pvr[81], dt;
};
void reset(Preint& s) {
  memset(&s, 0, sizeof(s));
  s.R = ident();
}
inline void set_block(double* A, int r, int c, const M3& b) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[9 * (r + i) + c + j] = b.m[3 * i + j];
}
// S <- A S A^T + Bg (sg I) Bg^T + Ba (sa I) Ba^T, Bg/Ba given as 9x3 row-major
void propagate(double* S, const double* A, const double* Bg, const double* Ba, double sg, double sa) {
  double T[81], N[81];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double s = 0;
      for (int k = 0; k < 9; ++k) s += A[9 * i + k] * S[9 * k + j];
      T[9 * i + j] = s;
    }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double s = 0;
      for (int k = 0; k < 9; ++k) s += T[9 * i + k] * A[9 * j + k];
      N[9 * i + j] = s;
    }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double g = 0, a = 0;
      for (int k = 0; k < 3; ++k) {
        g += (Bg[3 * i + k] * sg) * Bg[3 * j + k];
        a += (Ba[3 * i + k] * sa) * Ba[3 * j + k];
      }
      S[9 * i + j] = N[9 * i + j] + g + a;
    }
}

// IMUPreIntegratorBase::update (OdomPreIntegrator.h:432-506)
// test hook: when set, every update() call of this thread records its arguments (omega, acc, dt) — the sample selection /
// interpolation of PreIntegration is pinned against the reference's own function through this trace (tests/test_oracle_ref.py)
thread_local std::vector<double>* g_update_trace = nullptr;

void update(Preint& s, const double omega[3], const double acc[3], double dt, const OrcImuNoise& nz) {
  if (g_update_trace) {
    g_update_trace->insert(g_update_trace->end(), omega, omega + 3);
    g_update_trace->insert(g_update_trace->end(), acc, acc + 3);
    g_update_trace->push_back(dt);
  }
  const double dt2div2 = dt * dt / 2;
  const double wdt[3] = {omega[0] * dt, omega[1] * dt, omega[2] * dt};
  const M3 dR = so3_Exp(wdt), Jr = so3_Jr(wdt), skewa = hat(acc);
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
  const M3 Rsk = mul(s.R, skewa);
  const M3 nRsk_dt = scale(scale(Rsk, -1.0), dt), nRsk_dt2 = scale(scale(Rsk, -1.0), dt2div2);
  const M3 dRt = tr(dR), Idt = scale(ident(), dt), Jrdt = scale(Jr, dt), Rdt = scale(s.R, dt), Rdt2 = scale(s.R, dt2div2);
  double A[81], Bg[27], Ba[27];
  // P-R-V ordering (:444-463)
  memset(A, 0, sizeof(A)); memset(Bg, 0, sizeof(Bg)); memset(Ba, 0, sizeof(Ba));
  for (int i = 0; i < 9; ++i) A[10 * i] = 1;
  set_block(A, 3, 3, dRt);
  set_block(A, 6, 3, nRsk_dt);
  set_block(A, 0, 3, nRsk_dt2);
  set_block(A, 0, 6, Idt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Bg[3 * (3 + i) + j] = Jrdt.m[3 * i + j];
      Ba[3 * (6 + i) + j] = Rdt.m[3 * i + j];
      Ba[3 * i + j] = Rdt2.m[3 * i + j];
    }
  propagate(s.S_prv, A, Bg, Ba, sg, sa);
  // P-V-R ordering (:465-483)
  memset(A, 0, sizeof(A)); memset(Bg, 0, sizeof(Bg)); memset(Ba, 0, sizeof(Ba));
  for (int i = 0; i < 9; ++i) A[10 * i] = 1;
  set_block(A, 6, 6, dRt);
  set_block(A, 3, 6, nRsk_dt);
  set_block(A, 0, 6, nRsk_dt2);
  set_block(A, 0, 3, Idt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Bg[3 * (6 + i) + j] = Jrdt.m[3 * i + j];
      Ba[3 * (3 + i) + j] = Rdt.m[3 * i + j];
      Ba[3 * i + j] = Rdt2.m[3 * i + j];
    }
  propagate(s.S_pvr, A, Bg, Ba, sg, sa);
  // bias Jacobians, P then V then R, all with the old delta-R (:488-493)
  const M3 RskJgR = mul(Rsk, s.JgR);
  s.Jap = add(s.Jap, sub(scale(s.Jav, dt), scale(s.R, dt2div2)));
  s.Jgp = add(s.Jgp, sub(scale(s.Jgv, dt), scale(RskJgR, dt2div2)));
  s.Jav = add(s.Jav, scale(scale(s.R, -1.0), dt));
  s.Jgv = add(s.Jgv, scale(scale(RskJgR, -1.0), dt));
  s.JgR = sub(mul(dRt, s.JgR), Jrdt);
  // delta measurements (:497-503)
  double a2[3] = {acc[0] * dt2div2, acc[1] * dt2div2, acc[2] * dt2div2}, a1[3] = {acc[0] * dt, acc[1] * dt, acc[2] * dt};
  double Ra2[3], Ra1[3];
  mulv(s.R, a2, Ra2);
  mulv(s.R, a1, Ra1);
  for (int i = 0; i < 3; ++i) s.p[i] += s.v[i] * dt + Ra2[i];
  for (int i = 0; i < 3; ++i) s.v[i] += Ra1[i];
  s.R = normalize_rot(mul(s.R, dR));
  s.dt += dt;
}

}  // namespace

extern "C" {

// IMUDataBase::SetParam (src/Odom/OdomData.h:41-56): sigma2 = squared {gyro, acc, bias-gyro, bias-acc} noise
void orc_imu_set_param(OrcImuNoise* nz, const double sigma2[4], int dt_cov_noise_fixed, double freq_ref) {
  nz->sigma_g = sigma2[0];
  nz->sigma_a = sigma2[1];
  nz->sigma_bg = sigma2[2];
  nz->sigma_ba = sigma2[3];
  nz->dt_cov_noise_fixed = dt_cov_noise_fixed;
  if (dt_cov_noise_fixed && freq_ref) {
    nz->freq_ref = 0;
    nz->sigma_g *= freq_ref;
    nz->sigma_a *= freq_ref;
  } else
    nz->freq_ref = freq_ref;
}

// PreIntegration over samples[0..n) = rows {t, ax, ay, az, wx, wy, wz} (the std::list in the reference).
// Returns 0, or -1 for a gap > 1.5 s (delta-t reset to 0, OdomPreIntegrator.h:289-293).  n == 0 leaves the
// state untouched, as the reference does for an empty list.
int orc_imu_preintegrate(const double* smp, int n, double ti, double tj, const double bg[3], const double ba[3],
                         const OrcImuNoise* nz, OrcImuPreint* out) {
  Preint s;
  reset(s);
  int status = 0;
  auto T = [&](int i) { return smp[7 * i]; };
  if (n > 0) {
    const int END = n;
    const bool back = ti > tj;
    double tmin = ti, tmax = tj;
    if (back) std::swap(tmin, tmax);
    int start = 0, stop = END;
    for (int j = 0; j != END && T(j) <= tmin; start = j++) {
    }
    for (int j = END; j != 0;) {
      stop = j--;
      if (T(j) >= tmax) continue;
      break;
    }
    if (back) {
      if (stop == END) --stop;
      std::swap(start, stop);
      if (T(stop) > tmin) stop = END;  // reference asserts stop == begin here
    }
    for (int j = start; j != stop;) {
      const int jm1 = j;
      if (back) {
        if (j == 0) j = stop; else --j;
      } else
        ++j;
      const double tj_1 = jm1 == start ? ti : T(jm1);
      const double tjj = j == stop ? tj : T(j);
      double dt = tjj - tj_1;
      if (dt == 0) continue;
      if (std::fabs(dt) > 1.5) {
        s.dt = 0;
        status = -1;
        break;
      }
      double a0[3], w0[3], a1[3], w1[3], t1;  // imu (j-1) and imu_now (j)
      memcpy(a0, smp + 7 * jm1 + 1, 24);
      memcpy(w0, smp + 7 * jm1 + 4, 24);
      if (j != END) {
        memcpy(a1, smp + 7 * j + 1, 24);
        memcpy(w1, smp + 7 * j + 4, 24);
        t1 = T(j);
      } else {
        memcpy(a1, a0, 24);
        memcpy(w1, w0, 24);
        t1 = T(jm1);
      }
      if (j != END) {
        if (j == stop) {
          const double d = T(j) - tj;
          if (back ? d < 0 : d > 0) {
            const double rat = d / (T(j) - T(jm1));
            for (int k = 0; k < 3; ++k) {
              w1[k] = rat * w0[k] + (1 - rat) * w1[k];
              a1[k] = rat * a0[k] + (1 - rat) * a1[k];
            }
          }
        }
        if (jm1 == start) {
          const double d = ti - T(jm1);
          if (back ? d < 0 : d > 0) {
            const double rat = d / (T(j) - T(jm1));
            for (int k = 0; k < 3; ++k) {
              w0[k] = (1 - rat) * w0[k] + rat * w1[k];
              a0[k] = (1 - rat) * a0[k] + rat * a1[k];
            }
          }
        }
      }
      auto minus = [](const double x[3], const double b[3], double o[3]) {
        for (int k = 0; k < 3; ++k) o[k] = x[k] - b[k];
      };
      double om[3], ac[3];
      if (jm1 == start) {
        const double dc = T(jm1) - ti;
        if (back ? dc < 0 : dc > 0) {
          minus(w0, bg, om);
          minus(a0, ba, ac);
          update(s, om, ac, dc, *nz);
          dt -= dc;
          if (!dt) continue;
        }
      }
      double dcs = 0;
      if (j == stop) {
        dcs = tj - t1;
        if (back ? dcs < 0 : dcs > 0) dt -= dcs;
      }
      double wm[3], am[3];
      for (int k = 0; k < 3; ++k) {
        wm[k] = (w1[k] + w0[k]) / 2;
        am[k] = (a1[k] + a0[k]) / 2;
      }
      minus(wm, bg, om);
      minus(am, ba, ac);
      update(s, om, ac, dt, *nz);
      if (back ? dcs < 0 : dcs > 0) {
        minus(w1, bg, om);
        minus(a1, ba, ac);
        update(s, om, ac, dcs, *nz);
      }
    }
  }
  memcpy(out->Rij, s.R.m, 72);
  memcpy(out->vij, s.v, 24);
  memcpy(out->pij, s.p, 24);
  memcpy(out->SigmaPRV, s.S_prv, 648);
  memcpy(out->SigmaPVR, s.S_pvr, 648);
  memcpy(out->Jgp, s.Jgp.m, 72);
  memcpy(out->Jap, s.Jap.m, 72);
  memcpy(out->Jgv, s.Jgv.m, 72);
  memcpy(out->Jav, s.Jav.m, 72);
  memcpy(out->JgR, s.JgR.m, 72);
  out->dt = s.dt;
  out->status = status;
  return status;
}

// PreIntegration's update() calls alone: trace [cap][7] = (omega, acc, dt) per call; returns the status, *n_updates the call count
int orc_imu_preintegrate_trace(const double* smp, int n, double ti, double tj, const double bg[3], const double ba[3], double* trace,
                               int cap, int* n_updates) {
  std::vector<double> tr;
  OrcImuNoise nz;
  const double s2[4] = {1e-4, 1e-2, 1e-6, 1e-4};
  orc_imu_set_param(&nz, s2, 1, 200.0);
  OrcImuPreint out;
  g_update_trace = &tr;
  const int rc = orc_imu_preintegrate(smp, n, ti, tj, bg, ba, &nz, &out);
  g_update_trace = nullptr;
  *n_updates = (int)(tr.size() / 7);
  for (size_t k = 0; k < tr.size() && k < (size_t)cap * 7; ++k) trace[k] = tr[k];
  return rc;
}

// Optimizer::OptimizeInitialGyroBias (include/Optimizer.h:819-892) with EdgeGyrBias (src/Odom/g2otypes.h:940-973): one
// Gauss-Newton iteration (OptimizationAlgorithmGaussNewton: errors -> buildSystem -> solve -> update, no damping) on a
// 3-dim VertexGyrBias seeded with zero; one unary edge per keyframe i >= 1 whose pre-integration has dt != 0,
// information = (SigmaPRV.block<3,3>(3,3))^-1 (bInfo) or identity.  pre[i] is keyframe i's pre-integration (pre[0] is
// ignored), Rwb[i] = Rwc_i * Rcb row-major.  Output dbg = the vertex estimate (the caller adds it to bg); returns
// num_equations.  parity unpinned for Eigen's 3x3 inverse() / LDLT rounding (restated: cofactor inverse, Cholesky).
int orc_gyro_bias_init(const OrcImuPreint* pre, const double* Rwb, int n_kf, int use_info, double dbg[3]) {
  using namespace orc;
  double H[9] = {0}, b[3] = {0};
  int num_equations = 0;
  dbg[0] = dbg[1] = dbg[2] = 0;
  for (int i = 1; i < n_kf; ++i) {
    if (pre[i].dt == 0) continue;
    ++num_equations;
    M3 dR, JgR, Rwbi, Rwbj;
    memcpy(dR.m, pre[i].Rij, 72);
    memcpy(JgR.m, pre[i].JgR, 72);
    memcpy(Rwbi.m, Rwb + 9 * (i - 1), 72);
    memcpy(Rwbj.m, Rwb + 9 * i, 72);
    const double bg[3] = {0, 0, 0};
    double Jb[3];
    mulv(JgR, bg, Jb);
    // computeError: Log((deltaRij * Exp(JgRij * bg))^T * Rwbi^T * Rwbj)
    const M3 dRbg = so3_Exp(Jb);
    const M3 E = mul(mul(tr(mul(dR, dRbg)), tr(Rwbi)), Rwbj);
    double e[3];
    so3_log_q(qnormalized(mquat(E)), e);
    // linearizeOplus: -JrInv(e) * Exp(-e) * Jr(JgRij * bg) * JgRij
    const double ne[3] = {-e[0], -e[1], -e[2]};
    const M3 J = scale(mul(mul(mul(so3_JrInv(e), so3_Exp(ne)), so3_Jr(Jb)), JgR), -1.0);
    M3 W = ident();
    if (use_info) {
      // Matrix3d::inverse() of SigmaPRV(3:6, 3:6): cofactor formula
      double a[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a[3 * r + c] = pre[i].SigmaPRV[9 * (3 + r) + 3 + c];
      const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
      const double det = a[0] * c00 + a[1] * c01 + a[2] * c02, id = 1.0 / det;
      W = {{c00 * id, (a[2] * a[7] - a[1] * a[8]) * id, (a[1] * a[5] - a[2] * a[4]) * id,
            c01 * id, (a[0] * a[8] - a[2] * a[6]) * id, (a[2] * a[3] - a[0] * a[5]) * id,
            c02 * id, (a[1] * a[6] - a[0] * a[7]) * id, (a[0] * a[4] - a[1] * a[3]) * id}};
    }
    // constructQuadraticForm (base_unary_edge.hpp): H += J^T W J, b += -J^T W e
    const M3 JtW = mul(tr(J), W);
    const M3 JtWJ = mul(JtW, J);
    double JtWe[3];
    mulv(JtW, e, JtWe);
    for (int k = 0; k < 9; ++k) H[k] += JtWJ.m[k];
    for (int k = 0; k < 3; ++k) b[k] -= JtWe[k];
  }
  if (num_equations < 1) return num_equations;
  // solve H x = b (3x3 SPD): Cholesky
  double L[9] = {0};
  for (int j = 0; j < 3; ++j) {
    double d = H[3 * j + j];
    for (int k = 0; k < j; ++k) d -= L[3 * j + k] * L[3 * j + k];
    if (!(d > 0)) return num_equations;  // LDLT failure: the estimate stays zero
    L[3 * j + j] = std::sqrt(d);
    for (int r = j + 1; r < 3; ++r) {
      double v = H[3 * r + j];
      for (int k = 0; k < j; ++k) v -= L[3 * r + k] * L[3 * j + k];
      L[3 * r + j] = v / L[3 * j + j];
    }
  }
  double y[3];
  for (int r = 0; r < 3; ++r) {
    double v = b[r];
    for (int k = 0; k < r; ++k) v -= L[3 * r + k] * y[k];
    y[r] = v / L[3 * r + r];
  }
  for (int r = 2; r >= 0; --r) {
    double v = y[r];
    for (int k = r + 1; k < 3; ++k) v -= L[3 * k + r] * dbg[k];
    dbg[r] = v / L[3 * r + r];
  }
  return num_equations;
}

}  // extern "C"

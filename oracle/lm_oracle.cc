// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  g2o's Levenberg-Marquardt control flow restated over callbacks:
// SparseOptimizer::optimize (optimizer/g2o/g2o/core/sparse_optimizer.cpp:354-419) around
// OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale (optimization_algorithm_levenberg.cpp:61-189, with
// Raul Mur-Artal's stop criterion).  Pinned by the reference's own functions compiled unchanged against the same callbacks
// (oracle/ref_build/ref_lm_wrap.cc -> ref_lm_optimize; tests/test_oracle_ref.py compares every lambda, every accept / reject and
// the final state on linear, nonlinear, rank-deficient and diverging problems).
#include <algorithm>
#include <cmath>
#include <limits>

#include "ba_oracle.h"

extern "C" int orc_lm_optimize(const OrcLmCallbacks* cb, int iterations, double user_lambda_init, double* stats) {
  void* c = cb->ctx;
  const int n = cb->n;
  double lambda = -1., ni = 2;
  int nBad = 0, total_iters = 0, trials = 0;
  double chi_first = 0, chi_last = 0;
  bool ok = true;
  const double tau = 1e-5, good_upper = 2. / 3., good_lower = 1. / 3.;
  const int max_trials = 10;
  for (int it = 0; it < iterations && !(cb->terminate && cb->terminate(c)) && ok; ++it) {  // optimize() (:376)
    double currentChi = cb->errors(c);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (it == 0) chi_first = currentChi;
    cb->build(c);
    if (it == 0) {  // computeLambdaInit (:166-180)
      if (user_lambda_init > 0) lambda = user_lambda_init;
      else {
        double mx = 0;
        for (int j = 0; j < n; ++j) mx = std::max(std::fabs(cb->hessian_diag(c, j)), mx);
        lambda = tau * mx;
      }
      ni = 2;
      nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      cb->push(c);
      const bool ok2 = cb->solve(c, lambda) != 0;
      cb->update(c);  // the reference updates with whatever x holds, also after a failed solve; the pop below undoes it
      tempChi = cb->errors(c);
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;  // computeScale (:182-189)
      const double* x = cb->x(c);
      const double* b = cb->b(c);
      for (int j = 0; j < n; ++j) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, good_upper);
        lambda *= std::max(good_lower, alpha);
        ni = 2;
        currentChi = tempChi;
        cb->discard_top(c);
      } else {
        lambda *= ni;
        ni *= 2;
        cb->pop(c);
      }
      qmax++;
      trials++;
    } while (rho < 0 && qmax < max_trials && !(cb->terminate && cb->terminate(c)));
    ++total_iters;
    chi_last = currentChi;
    if (qmax == max_trials || rho == 0) { ok = false; continue; }  // Terminate
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
    else nBad = 0;
    if (nBad >= 3) ok = false;
  }
  if (stats) {
    stats[0] = chi_first; stats[1] = chi_last; stats[2] = total_iters; stats[3] = lambda; stats[4] = trials;
  }
  return total_iters;
}

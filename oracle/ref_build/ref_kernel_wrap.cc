// ORACLE — TEST INFRASTRUCTURE ONLY.
// g2o::RobustKernelHuber::setDelta / setDeltaSqr / robustify (optimizer/g2o/g2o/core/robust_kernel_impl.cpp:60-91, this fork keeps
// delta^2 in a FLOAT member, robust_kernel_impl.h:84) of the REFERENCE compiled UNCHANGED: the three definitions are cut out by
// name at build time (oracle/_ref/gen/kernel_fns.inc); this file supplies the class declaration and a three-double stand-in for
// Eigen::Vector3d (the body only indexes it).  Also g2o::GraphOperator (optimizer/optimizer_ba/g2o_graph_operator.h:10-41: the 5 % chi-square
// table and Chi2LargeSetLevel), class and function cut out by name, run over a stand-in edge that reports a given chi2.
#include <cmath>
#include <vector>
namespace Eigen {
struct Vector3d {
  double v[3];
  double& operator[](int i) { return v[i]; }
};
}  // namespace Eigen
namespace g2o {
class RobustKernel {  // core/robust_kernel.h:50-70
 public:
  virtual ~RobustKernel() {}
  double _delta = 1.;
};
class RobustKernelHuber : public RobustKernel {  // core/robust_kernel_impl.h:76-86
 public:
  virtual void setDelta(double delta);
  virtual void setDeltaSqr(const double& delta, const double& deltaSqr);
  virtual void robustify(double e2, Eigen::Vector3d& rho) const;
  float dsqr;
};
#include "kernel_fns.inc"
}  // namespace g2o

extern "C" void ref_huber(double delta, double e, double rho[3]) {
  g2o::RobustKernelHuber k;
  k.setDelta(delta);
  Eigen::Vector3d r{{0, 0, 0}};
  k.robustify(e, r);
  rho[0] = r[0]; rho[1] = r[1]; rho[2] = r[2];
}

namespace G2O_GRAPHOP {
#include "graphop_fns.inc"
struct EdgeStandin {
  double c = 0;
  int level = 0;
  void computeError() {}
  double chi2() const { return c; }
  void setLevel(int l) { level = l; }
  void linearizeOplus(int) {}
};
struct EdgeRec {
  EdgeStandin* pedge;
};
struct GraphStandin {
  void initializeOptimization() {}
  int jacobianWorkspace() { return 0; }
};
}  // namespace G2O_GRAPHOP
extern "C" int ref_chi2_large_level(double chi2, int dim_freedom, float rat_th_chi2) {
  using namespace G2O_GRAPHOP;
  EdgeStandin e;
  e.c = chi2;
  std::vector<EdgeRec> es{{&e}};
  GraphStandin g;
  GraphOperator::Chi2LargeSetLevel(g, es, dim_freedom, rat_th_chi2, false);
  return e.level;
}

// ORACLE — TEST INFRASTRUCTURE ONLY.
// g2o::RobustKernelHuber::setDelta / setDeltaSqr / robustify (optimizer/g2o/g2o/core/robust_kernel_impl.cpp:60-91, this fork keeps
// delta^2 in a FLOAT member, robust_kernel_impl.h:84) of the REFERENCE compiled UNCHANGED: the three definitions are cut out by
// name at build time (oracle/_ref/gen/kernel_fns.inc); this file supplies the class declaration and a three-double stand-in for
// Eigen::Vector3d (the body only indexes it).
#include <cmath>
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

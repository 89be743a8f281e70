// ORACLE — TEST INFRASTRUCTURE ONLY.
// IMUPreIntegratorBase<IMUDataBase>::PreIntegration (src/Odom/OdomPreIntegrator.h:227-430) of the REFERENCE compiled UNCHANGED: the
// sample selection around [t_i, t_j] (forward and reversed time), the interpolation of the first / last sample to the frame stamps,
// the mid-point rule, the dt == 0 skip and the 1.5 s gap abort.  The function definition is cut out of the header by name at
// build time (oracle/_ref/gen/imu_fns.inc); update() — the covariance / Jacobian recurrence, Eigen block algebra — is NOT compiled:
// the stand-in records its arguments, which is exactly what this function decides.  Vector3d is a three-double stand-in with the
// element-wise operators the body uses (Eigen evaluates them element by element in the same order).
// `abs(dt)` in the body is called on a double: which overload it binds to depends on the reference translation unit's includes
// (with the C++ <math.h> / <stdlib.h> wrappers or a using-directive in effect it is the double one, the intended meaning; with
// only <cmath> / <cstdlib> visible it would be ::abs(int) and the 1.5 s gate would act at 2 s).  This wrapper takes the double one.
#include <math.h>
#include <stdlib.h>
#include <cassert>
#include <cmath>
#include <iostream>
#include <list>
#include <vector>

#define listeig(T) std::list<T>

namespace VIEO_SLAM {
struct Vector3d {
  double v[3];
};
inline Vector3d operator+(const Vector3d& a, const Vector3d& b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
inline Vector3d operator-(const Vector3d& a, const Vector3d& b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
inline Vector3d operator*(double s, const Vector3d& a) { return {{s * a.v[0], s * a.v[1], s * a.v[2]}}; }
inline Vector3d operator/(const Vector3d& a, double s) { return {{a.v[0] / s, a.v[1] / s, a.v[2] / s}}; }

struct IMUDataS {  // src/Odom/OdomData.h:22-36: the members PreIntegration reads
  double mtm;
  Vector3d ma, mw;
};

template <class IMUDataBase>
class IMUPreIntegratorBase {  // declaration subset of OdomPreIntegrator.h:108-223
 public:
  typedef double Tcalc;
  double mdeltatij = 0;
  std::vector<double> trace;
  int resets = 0;
  void reset() {
    ++resets;
    mdeltatij = 0;
  }
  void update(const Vector3d& omega, const Vector3d& acc, const double& dt) {
    for (int k = 0; k < 3; ++k) trace.push_back(omega.v[k]);
    for (int k = 0; k < 3; ++k) trace.push_back(acc.v[k]);
    trace.push_back(dt);
  }
  int PreIntegration(const double& timeStampi, const double& timeStampj, const Vector3d& bgi_bar, const Vector3d& bai_bar,
                     const typename listeig(IMUDataBase)::const_iterator& iterBegin,
                     const typename listeig(IMUDataBase)::const_iterator& iterEnd, bool breset = true);
};

template <class IMUDataBase>
#include "imu_fns.inc"

}  // namespace VIEO_SLAM

// samples: rows {t, ax, ay, az, wx, wy, wz}
extern "C" int ref_imu_preintegrate_trace(const double* smp, int n, double ti, double tj, const double bg[3], const double ba[3],
                                          double* trace, int cap, int* n_updates, double* deltat) {
  using namespace VIEO_SLAM;
  std::list<IMUDataS> l;
  for (int i = 0; i < n; ++i) {
    IMUDataS d;
    d.mtm = smp[7 * i];
    for (int k = 0; k < 3; ++k) {
      d.ma.v[k] = smp[7 * i + 1 + k];
      d.mw.v[k] = smp[7 * i + 4 + k];
    }
    l.push_back(d);
  }
  IMUPreIntegratorBase<IMUDataS> p;
  p.mdeltatij = 123.0;  // the abort path resets it to 0
  const Vector3d g{{bg[0], bg[1], bg[2]}}, a{{ba[0], ba[1], ba[2]}};
  const int rc = p.PreIntegration(ti, tj, g, a, l.begin(), l.end(), true);
  *n_updates = (int)(p.trace.size() / 7);
  for (size_t k = 0; k < p.trace.size() && k < (size_t)cap * 7; ++k) trace[k] = p.trace[k];
  if (deltat) *deltat = p.mdeltatij;
  return rc;
}

// ORACLE — TEST INFRASTRUCTURE ONLY.
// g2o's Sim3 of the REFERENCE's vendored copy compiled UNCHANGED against the Eigen stand-in (eigstub/): optimizer/g2o/g2o/types/sim3.h
// and se3_ops.h / .hpp whole, from where they lie (exp as the Vector7d constructor, log, inverse, operator*, map); VertexSim3Expmap
// and EdgeSim3 (types_seven_dof_expmap.h: class definitions with oplusImpl / computeError) and the numeric-Jacobian
// BaseBinaryEdge::linearizeOplus (core/base_binary_edge.hpp:131-203), cut out by name at build time into oracle/_ref/gen/sim3_*.inc.
// Declared by hand: the vertex / edge base classes those texts name (estimate, fixed flag, a one-deep push / pop stack, oplus).
#include <math.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <iostream>
#include <set>
#include <vector>

#include "optimizer/g2o/g2o/types/sim3.h"  // the reference's, unchanged

#include "../ba_oracle.h"  // OrcSim3 layout only

namespace G2O_SIM3 {
using namespace g2o;
using namespace Eigen;

struct OptimizableGraph {
  struct Vertex {
    virtual ~Vertex() {}
  };
  typedef std::set<Vertex*> VertexSet;
};
template <int D, class T>
class BaseVertex : public OptimizableGraph::Vertex {
 public:
  static const int Dimension = D;
  const T& estimate() const { return _estimate; }
  void setEstimate(const T& e) { _estimate = e; }
  bool fixed() const { return _fixed; }
  void setFixed(bool f) { _fixed = f; }
  void push() { _backup.push_back(_estimate); }
  void pop() {
    _estimate = _backup.back();
    _backup.pop_back();
  }
  virtual void oplusImpl(const double* v) = 0;
  void oplus(const double* v) { oplusImpl(v); }

 protected:
  T _estimate;
  std::vector<T> _backup;
  bool _fixed = false, _marginalized = false;
};
class VertexSBAPointXYZ : public BaseVertex<3, Vector3d> {
 public:
  void oplusImpl(const double*) override {}
};
template <int D, class E, class VertexXiType, class VertexXjType>
class BaseBinaryEdge {
 public:
  typedef Matrix<double, D, 1> ErrorVector;
  virtual ~BaseBinaryEdge() {}
  virtual void computeError() = 0;
  virtual void linearizeOplus();
  const E& measurement() const { return _measurement; }
  void setMeasurement(const E& m) { _measurement = m; }
  OptimizableGraph::Vertex* _vertices[2] = {nullptr, nullptr};
  ErrorVector _error;
  E _measurement;
  Matrix<double, D, VertexXiType::Dimension> _jacobianOplusXi;
  Matrix<double, D, VertexXjType::Dimension> _jacobianOplusXj;
};
template <int D, typename E, typename VertexXiType, typename VertexXjType>
#include "sim3_linearize.inc"

#include "sim3_vertex.inc"
#include "sim3_edge.inc"
VertexSim3Expmap::VertexSim3Expmap() : BaseVertex<7, Sim3>() {
  _marginalized = false;
  _fix_scale = false;
}
bool VertexSim3Expmap::read(std::istream&) { return true; }
bool VertexSim3Expmap::write(std::ostream&) const { return true; }
EdgeSim3::EdgeSim3() {}
bool EdgeSim3::read(std::istream&) { return true; }
bool EdgeSim3::write(std::ostream&) const { return true; }
}  // namespace G2O_SIM3

namespace {
using namespace G2O_SIM3;
Sim3 to_s(const OrcSim3& o) { return Sim3(Quaterniond(o.q[3], o.q[0], o.q[1], o.q[2]), Vector3d(o.t), o.s); }
void from_s(const Sim3& s, OrcSim3* o) {
  for (int k = 0; k < 8; ++k) (&o->q[0])[k] = s[k];  // Sim3::operator[]: x y z w, t, s
}
}  // namespace

// op 0: exp(u[7]) -> out; 1: log(a) -> u[7]; 2: a * b -> out; 3: a^-1 -> out; 4: VertexSim3Expmap::oplusImpl(u) on a (fix_scale in b->s != 0)
extern "C" void ref_sim3(int op, const OrcSim3* a, const OrcSim3* b, double* u, OrcSim3* out) {
  static_assert(sizeof(OrcSim3) == 8 * sizeof(double), "layout");
  if (op == 0) {
    Vector7d v;
    for (int k = 0; k < 7; ++k) v[k] = u[k];
    from_s(Sim3(v), out);
  } else if (op == 1) {
    const Vector7d v = to_s(*a).log();
    for (int k = 0; k < 7; ++k) u[k] = v[k];
  } else if (op == 2) {
    from_s(to_s(*a) * to_s(*b), out);
  } else if (op == 3) {
    from_s(to_s(*a).inverse(), out);
  } else if (op == 4) {
    VertexSim3Expmap v;
    v.setEstimate(to_s(*a));
    v._fix_scale = b && b->s != 0;
    double upd[7];
    for (int k = 0; k < 7; ++k) upd[k] = u[k];
    v.oplusImpl(upd);
    from_s(v.estimate(), out);
  }
}

// EdgeSim3::computeError and the numeric Jacobians of BaseBinaryEdge::linearizeOplus: e [7], Ji / Jj [7][7] row-major (zero for a fixed vertex)
extern "C" void ref_edge_sim3_graph(const OrcSim3* meas, const OrcSim3* v0, const OrcSim3* v1, int fix0, int fix1, int fix_scale, double e[7],
                                    double* Ji, double* Jj) {
  VertexSim3Expmap a, b;
  a.setEstimate(to_s(*v0)), b.setEstimate(to_s(*v1));
  a.setFixed(fix0 != 0), b.setFixed(fix1 != 0);
  a._fix_scale = b._fix_scale = fix_scale != 0;
  EdgeSim3 ed;
  ed._vertices[0] = &a, ed._vertices[1] = &b;
  ed.setMeasurement(to_s(*meas));
  ed.computeError();
  for (int k = 0; k < 7; ++k) e[k] = ed._error[k];
  if (Ji && Jj) {
    ed.linearizeOplus();
    for (int i = 0; i < 7; ++i)
      for (int j = 0; j < 7; ++j) Ji[7 * i + j] = ed._jacobianOplusXi(i, j), Jj[7 * i + j] = ed._jacobianOplusXj(i, j);
  }
}

// ORACLE — TEST INFRASTRUCTURE ONLY.
// g2o's Levenberg-Marquardt control flow of the REFERENCE compiled UNCHANGED over the oracle's callback interface
// (OrcLmCallbacks, ../ba_oracle.h): SparseOptimizer::optimize (optimizer/g2o/g2o/core/sparse_optimizer.cpp:354-419) and
// OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale (optimization_algorithm_levenberg.cpp:61-189, with the
// stop criterion this fork adds).  The four function definitions are cut out of the sources by name at build time
// (oracle/_ref/gen/lm_fns.inc); the rest of g2o needs Eigen (absent here), so this file supplies only what the bodies touch: a
// SparseOptimizer / Solver / vertex stand-in whose every operation forwards to a callback, the Property and batch-statistics
// shells, and the constructor's constants (optimization_algorithm_levenberg.cpp:44-55).
#include <math.h>
#include <cassert>
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <vector>

#include "../ba_oracle.h"
using namespace std;

#define FIXED(s) s
namespace g2o {

inline double get_monotonic_time() { return 0.0; }
inline bool g2o_isfinite(double x) { return std::isfinite(x); }

struct G2OBatchStatistics {  // core/batch_stats.h (only the fields the two functions write)
  int iteration = 0, numVertices = 0, numEdges = 0, levenbergIterations = 0;
  double chi2 = 0, timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0, timeIteration = 0;
  static G2OBatchStatistics* globalStats() { return _globalStats; }
  static void setGlobalStats(G2OBatchStatistics* b) { _globalStats = b; }
  static G2OBatchStatistics* _globalStats;
};
G2OBatchStatistics* G2OBatchStatistics::_globalStats = nullptr;

template <class T>
struct Property {  // stuff/property.h
  T _value;
  const T& value() const { return _value; }
  void setValue(const T& v) { _value = v; }
};

struct OptimizableGraph {
  struct Vertex;  // == VertexS below (computeLambdaInit names the type)
};
struct OptimizableGraph::Vertex {  // one pseudo-vertex spanning the whole system (computeLambdaInit walks indexMapping())
  const OrcLmCallbacks* cb;
  int dimension() const { return cb->n; }
  double hessian(int i, int j) const { return i == j ? cb->hessian_diag(cb->ctx, i) : 0.0; }
};
using VertexS = OptimizableGraph::Vertex;

class OptimizationAlgorithmLevenberg;
class SparseOptimizer {
 public:
  const OrcLmCallbacks* cb = nullptr;
  VertexS vertex;
  std::vector<VertexS*> _ivMap;  // == indexMapping()
  OptimizationAlgorithmLevenberg* _algorithm = nullptr;
  std::vector<G2OBatchStatistics> _batchStatistics;
  bool _computeBatchStatistics = false;
  std::vector<int> _activeEdges, _activeVertices;
  const std::vector<VertexS*>& indexMapping() const { return _ivMap; }
  void computeActiveErrors() { last_chi = cb->errors(cb->ctx); }
  double activeRobustChi2() const { return last_chi; }
  void push() { cb->push(cb->ctx); }
  void pop() { cb->pop(cb->ctx); }
  void discardTop() { cb->discard_top(cb->ctx); }
  void update(const double*) { cb->update(cb->ctx); }
  bool terminate() { return cb->terminate ? cb->terminate(cb->ctx) != 0 : false; }
  bool verbose() const { return false; }
  void preIteration(int) {}
  void postIteration(int);  // (this wrapper's only instrumentation: sums the trials of the iteration that just ended)
  int trials = 0;
  int optimize(int iterations, bool online = false);
  double last_chi = 0;
};

class Solver {
 public:
  const OrcLmCallbacks* cb = nullptr;
  SparseOptimizer* _optimizer = nullptr;
  double lambda = 0;
  SparseOptimizer* optimizer() const { return _optimizer; }
  bool buildStructure() { return true; }
  bool buildSystem() {
    cb->build(cb->ctx);
    return true;
  }
  bool setLambda(double l, bool) {
    lambda = l;
    return true;
  }
  bool solve() { return cb->solve(cb->ctx, lambda) != 0; }
  void restoreDiagonal() {}
  const double* x() const { return cb->x(cb->ctx); }
  const double* b() const { return cb->b(cb->ctx); }
  size_t vectorSize() const { return (size_t)cb->n; }
};

class OptimizationAlgorithm {
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };  // core/optimization_algorithm.h:49
};

class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {
 public:
  OptimizationAlgorithmLevenberg(Solver* solver, SparseOptimizer* opt) : _optimizer(opt), _solver(solver) {
    // constructor body of the reference (optimization_algorithm_levenberg.cpp:44-55)
    _currentLambda = -1.;
    _tau = 1e-5;
    _goodStepUpperScale = 2. / 3.;
    _goodStepLowerScale = 1. / 3.;
    _userLambdaInit = &_pUser;
    _userLambdaInit->setValue(0.);
    _maxTrialsAfterFailure = &_pMax;
    _maxTrialsAfterFailure->setValue(10);
    _ni = 2.;
    _levenbergIterations = 0;
    _nBad = 0;
  }
  bool init(bool) { return true; }
  void printVerbose(std::ostream&) const {}
  SolverResult solve(int iteration, bool online = false);
  double computeLambdaInit() const;
  double computeScale() const;
  SparseOptimizer* _optimizer;
  Solver* _solver;
  Property<int>* _maxTrialsAfterFailure;
  Property<double>* _userLambdaInit;
  double _currentLambda, _tau, _goodStepLowerScale, _goodStepUpperScale, _ni;
  int _levenbergIterations, _nBad;
  Property<int> _pMax;
  Property<double> _pUser;
};

inline void SparseOptimizer::postIteration(int) { trials += _algorithm->_levenbergIterations; }

#include "lm_fns.inc"

}  // namespace g2o

extern "C" int ref_lm_optimize(const OrcLmCallbacks* cb, int iterations, double user_lambda_init, double* stats) {
  using namespace g2o;
  // chi2 at the first / after the last accepted step are observed from outside: the callbacks are wrapped to record them
  struct Probe {
    const OrcLmCallbacks* in;
    double first = 0, last_err = 0, cur = 0;  // cur: chi2 after the last accepted step (or at the iteration's start)
    int calls = 0;
  } probe{cb};
  OrcLmCallbacks w = *cb;
  w.ctx = &probe;
  w.errors = [](void* c) {
    Probe* p = (Probe*)c;
    const double v = p->in->errors(p->in->ctx);
    if (p->calls++ == 0) p->first = v;
    p->last_err = v;
    return v;
  };
  w.build = [](void* c) {
    Probe* p = (Probe*)c;
    p->cur = p->last_err;  // buildSystem follows the iteration's first error evaluation
    p->in->build(p->in->ctx);
  };
  w.solve = [](void* c, double l) { return ((Probe*)c)->in->solve(((Probe*)c)->in->ctx, l); };
  w.update = [](void* c) { ((Probe*)c)->in->update(((Probe*)c)->in->ctx); };
  w.push = [](void* c) { ((Probe*)c)->in->push(((Probe*)c)->in->ctx); };
  w.pop = [](void* c) { ((Probe*)c)->in->pop(((Probe*)c)->in->ctx); };
  w.discard_top = [](void* c) {
    Probe* p = (Probe*)c;
    p->cur = p->last_err;  // the step was accepted
    p->in->discard_top(p->in->ctx);
  };
  w.x = [](void* c) { return ((Probe*)c)->in->x(((Probe*)c)->in->ctx); };
  w.b = [](void* c) { return ((Probe*)c)->in->b(((Probe*)c)->in->ctx); };
  w.hessian_diag = [](void* c, int j) { return ((Probe*)c)->in->hessian_diag(((Probe*)c)->in->ctx, j); };
  w.terminate = cb->terminate ? +[](void* c) { return ((Probe*)c)->in->terminate(((Probe*)c)->in->ctx); } : nullptr;
  SparseOptimizer opt;
  opt.cb = &w;
  opt.vertex.cb = &w;
  opt._ivMap.push_back(&opt.vertex);
  Solver solver;
  solver.cb = &w;
  solver._optimizer = &opt;
  OptimizationAlgorithmLevenberg lm(&solver, &opt);
  lm._userLambdaInit->setValue(user_lambda_init);
  opt._algorithm = &lm;
  // batch statistics stay OFF as in the reference's runs (with them on, optimize() evaluates the errors once more per iteration,
  // which a caller can observe in the edges' stored chi2)
  const int its = opt.optimize(iterations);
  if (stats) {
    stats[0] = probe.first;
    stats[1] = probe.cur;
    stats[2] = its;
    stats[3] = lm._currentLambda;
    stats[4] = opt.trials;
  }
  return its;
}

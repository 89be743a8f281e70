// Development check (CPU): runs the text of vieo_slam_b200/csrc/sim3.cuh on the host against the oracle's restatement
// (oracle/posegraph_oracle.cc).  Same libm, same operation order -> expected bit-identical.  Not part of the product.
//   g++ -O2 -ffp-contract=off -std=c++17 -Ioracle tools/sim3_host_check.cc -Loracle -loracle -Wl,-rpath,$PWD/oracle -o /tmp/s3chk && /tmp/s3chk
#include <cstdio>
#include <cstring>
#include <random>

#include "../vieo_slam_b200/csrc/sim3.cuh"
#include "ba_oracle.h"

using vieo::Sim3d;
static_assert(sizeof(Sim3d) == sizeof(OrcSim3), "layout");

int main() {
  std::mt19937_64 g(5);
  std::normal_distribution<double> N(0, 1);
  int bad = 0;
  const double scales[] = {1.0, 1e-3, 1e-7, 0.0};
  for (int it = 0; it < 20000; ++it) {
    double u[7], v[7];
    const double a = scales[it % 4], b = scales[(it / 4) % 4];
    for (int k = 0; k < 3; ++k) { u[k] = N(g) * a; v[k] = N(g) * 0.8; }
    for (int k = 3; k < 6; ++k) { u[k] = N(g); v[k] = N(g) * 3; }
    u[6] = N(g) * b * 0.2; v[6] = (it % 3 == 0) ? 0.0 : N(g) * 0.1;
    OrcSim3 eo, fo;
    orc_sim3_exp(u, &eo);
    orc_sim3_exp(v, &fo);
    Sim3d e = vieo::s3_exp(u), f = vieo::s3_exp(v);
    if (memcmp(&e, &eo, sizeof(e)) || memcmp(&f, &fo, sizeof(f))) { if (bad++ < 5) printf("exp differs at %d\n", it); }
    double lo[7], l[7];
    orc_sim3_log(&fo, lo);
    vieo::s3_log(f, l);
    if (memcmp(lo, l, sizeof(l))) { if (bad++ < 5) printf("log differs at %d\n", it); }
    OrcSim3 mo, io;
    orc_sim3_mul(&eo, &fo, &mo);
    orc_sim3_inv(&fo, &io);
    Sim3d m = vieo::s3_mul(e, f), i = vieo::s3_inv(f);
    if (memcmp(&m, &mo, sizeof(m)) || memcmp(&i, &io, sizeof(i))) { if (bad++ < 5) printf("mul/inv differs at %d\n", it); }
    double eo7[7], e7[7];
    orc_edge_sim3_graph(&mo, &eo, &fo, 0, 0, 0, eo7, nullptr, nullptr);
    vieo::s3_edge_error(m, e, f, e7);
    if (memcmp(eo7, e7, sizeof(e7))) { if (bad++ < 5) printf("edge error differs at %d\n", it); }
  }
  printf("sim3 host check: %d mismatches\n", bad);
  return bad != 0;
}

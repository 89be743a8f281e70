// Optimizer::OptimizeEssentialGraph on the device (src/Optimizer.cc:2309-2688): the loop closer's Sim3 pose graph.
// VertexSim3Expmap per keyframe, EdgeSim3 per essential-graph edge (optimizer/g2o/g2o/types/types_seven_dof_expmap.h:22-129),
// numeric Jacobians as g2o takes them (core/base_binary_edge.hpp:123-187, delta 1e-9), the quadratic form of a binary edge
// without a kernel (:60-85), g2o's Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:61-189) over optimize(20).
//
// The whole optimisation is ONE cooperative kernel (one CTA per SM, device-wide barriers between phases), because every
// phase is tiny and the control flow (accept / reject, lambda, the three stop rules) depends on device results:
//   errors      one thread per edge, deterministic two-level sum (fixed thread -> edge map, CTA partials summed in order)
//   linearise   one thread per (edge, vertex side, column): two perturbed error evaluations each (E x 14 x 2 Sim3 chains)
//   assemble    diagonal blocks and b gathered per free vertex over its incident edges in edge order (no atomics, the
//               oracle's order); off-diagonal 7 x 7 blocks by atomicAdd (at most the duplicates of one keyframe pair meet)
//   solve       the 7 n_free system lives DENSE in HBM as 56 x 56 tiles (8 vertices), but only tiles of the filled pattern
//               are ever touched: the host runs a tile-level symbolic factorisation once (spanning tree + covisibility =
//               band, loop edges = a few dense tile rows) and hands the kernel per-panel work lists.  Right-looking tile
//               Cholesky: diagonal tile by one warp (Crout, y carried as an extra row = forward substitution), row tiles
//               below by one thread per row, trailing updates one CTA per tile pair; back substitution by one CTA.
//   update      Snew = exp(x) Sold per free vertex; g2o's gain ratio from the new errors.
// HBM: 2 x (7 n_free)^2 doubles (H and its damped factor): 125 MB at 400 keyframes, 3.1 GB at 2000.
#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>
#include <vector>

#include "common.cuh"
#include "sim3.cuh"

namespace cg = cooperative_groups;

namespace vieo {

static_assert(sizeof(Sim3d) == sizeof(VieoSim3), "Sim3d mirrors VieoSim3");

constexpr int kPgTS = 56;              // tile edge: 8 vertices x 7
constexpr int kPgVT = 8;               // vertices per tile
constexpr int kPgLd = kPgTS + 1;       // shared-memory row stride (doubles)
constexpr int kPgThreads = 256;
constexpr int kPgTile2 = kPgTS * kPgTS;
constexpr size_t kPgSmem = sizeof(double) * 2 * kPgTS * kPgLd;  // two tiles (the second one's first row also carries y)

struct PgParams {
  int K, E, nv, n, T, ld;
  int fix_scale, iterations, single_step;
  double lambda_init;
  Sim3d* est;
  Sim3d* bak;
  const Sim3d* meas;
  const double* info;  // [E][49] or nullptr
  const int* ei;
  const int* ej;
  const int* hidx;     // vertex -> free index or -1
  const int* fv;       // free index -> vertex
  const uint8_t* act;  // edge active
  double* err;         // [E][7]
  double* Ji;          // [E][49]
  double* Jj;
  const int* vptr;     // [nv + 1] incidence lists of the free vertices
  const int* vinc;     // (edge << 1) | side
  const int* offd;     // edges with two distinct free vertices
  int n_offd;
  double* H;
  double* L;
  double* b;
  double* x;
  double* y;
  const int* tl_i;     // tiles of the filled pattern (lower triangle incl. diagonal)
  const int* tl_j;
  const uint8_t* tl_isH;
  int n_tiles;
  const int* trsm_off;  // [T + 1] per panel: row tiles below
  const int* trsm_row;
  const int* syrk_off;  // [T + 1] per panel: tile pairs (i >= j) of the trailing update
  const int* syrk_i;
  const int* syrk_j;
  const int* row_off;   // [T + 1] per tile row k: columns j < k of the filled pattern
  const int* row_col;
  double* part;         // [3 G] CTA partial sums: errors at the iteration start | computeScale | errors of the trial
                        // (a region is rewritten only after a device-wide barrier that follows its last read)
  int* flags;           // [0] factorisation ok
  VieoPoseGraphStats* stats;
  double* Tcw;          // [K][12] or nullptr
};

__device__ __forceinline__ double pg_block_sum(double v, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // s_red free
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < kPgThreads / 32; ++w) t += s_red[w];
  return t;
}

// sum of the G CTA partials at part[0..G), same order in every CTA
__device__ __forceinline__ double pg_grid_total(const double* part, int G, double* s_bc) {
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int g = 0; g < G; ++g) t += __ldcg(part + g);
    *s_bc = t;
  }
  __syncthreads();
  return *s_bc;
}

// computeActiveErrors + activeChi2 (chi2 = e' Omega e); leaves the CTA partial in part[blockIdx.x] (caller syncs the grid)
__device__ void pg_errors(const PgParams& P, double* part, double* s_red) {
  const int gtid = blockIdx.x * kPgThreads + threadIdx.x, gsz = gridDim.x * kPgThreads;
  double local = 0;
  for (int e = gtid; e < P.E; e += gsz) {
    if (!P.act[e]) continue;
    double r[7];
    s3_edge_error(P.meas[e], P.est[P.ei[e]], P.est[P.ej[e]], r);
    double c = 0;
    if (P.info) {
      const double* O = P.info + 49 * (size_t)e;
      for (int a = 0; a < 7; ++a) {
        double s = 0;
        for (int k = 0; k < 7; ++k) s += O[7 * a + k] * r[k];
        c += r[a] * s;
      }
    } else {
      for (int a = 0; a < 7; ++a) c += r[a] * r[a];
    }
    for (int a = 0; a < 7; ++a) P.err[7 * (size_t)e + a] = r[a];
    local += c;
  }
  const double t = pg_block_sum(local, s_red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// BaseBinaryEdge::linearizeOplus: column d of the Jacobian w.r.t. one vertex by central differences
__device__ void pg_linearize(const PgParams& P) {
  const int gtid = blockIdx.x * kPgThreads + threadIdx.x, gsz = gridDim.x * kPgThreads;
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  const bool fs = P.fix_scale != 0;
  for (int w = gtid; w < P.E * 14; w += gsz) {
    const int e = w / 14, c = w % 14, side = c / 7, d = c % 7;
    if (!P.act[e]) continue;
    const int vi = P.ei[e], vj = P.ej[e];
    if (P.hidx[side ? vj : vi] < 0) continue;
    const Sim3d m = P.meas[e], v0 = P.est[vi], v1 = P.est[vj];
    double add[7] = {0, 0, 0, 0, 0, 0, 0}, e1[7], e2[7];
    add[d] = delta;
    {
      const Sim3d pv = s3_oplus(side ? v1 : v0, add, fs);
      s3_edge_error(m, side ? v0 : pv, side ? pv : v1, e1);
    }
    add[d] = -delta;
    {
      const Sim3d pv = s3_oplus(side ? v1 : v0, add, fs);
      s3_edge_error(m, side ? v0 : pv, side ? pv : v1, e2);
    }
    double* J = (side ? P.Jj : P.Ji) + 49 * (size_t)e;
    for (int r = 0; r < 7; ++r) J[7 * r + d] = scalar * (e1[r] - e2[r]);
  }
}

__device__ void pg_zero_system(const PgParams& P) {
  const int gtid = blockIdx.x * kPgThreads + threadIdx.x, gsz = gridDim.x * kPgThreads;
  for (long long idx = gtid; idx < (long long)P.n_tiles * kPgTile2; idx += gsz) {
    const int t = (int)(idx / kPgTile2), rc = (int)(idx % kPgTile2);
    if (!P.tl_isH[t]) continue;
    P.H[(size_t)(P.tl_i[t] * kPgTS + rc / kPgTS) * P.ld + P.tl_j[t] * kPgTS + rc % kPgTS] = 0.0;
  }
}

// constructQuadraticForm of every active edge (base_binary_edge.hpp:60-85, no robust kernel)
__device__ void pg_assemble(const PgParams& P) {
  const int gtid = blockIdx.x * kPgThreads + threadIdx.x, gsz = gridDim.x * kPgThreads;
  // diagonal blocks (c < 7) and b (c == 7), gathered per free vertex in edge order
  for (int w = gtid; w < P.nv * 56; w += gsz) {
    const int v = w / 56, a = (w % 56) / 8, c = w % 8;
    double acc = 0;
    for (int p = P.vptr[v]; p < P.vptr[v + 1]; ++p) {
      const int inc = P.vinc[p], e = inc >> 1;
      const double* J = ((inc & 1) ? P.Jj : P.Ji) + 49 * (size_t)e;
      const double* O = P.info ? P.info + 49 * (size_t)e : nullptr;
      double s = 0;
      if (c < 7) {
        for (int k = 0; k < 7; ++k) {
          double ao;  // (J' Omega)[a][k]
          if (O) {
            ao = 0;
            for (int m = 0; m < 7; ++m) ao += J[7 * m + a] * O[7 * m + k];
          } else ao = J[7 * k + a];
          s += ao * J[7 * k + c];
        }
      } else {
        const double* r = P.err + 7 * (size_t)e;
        for (int k = 0; k < 7; ++k) {
          double ork;  // omega_r = -Omega e
          if (O) {
            ork = 0;
            for (int m = 0; m < 7; ++m) ork += O[7 * k + m] * r[m];
            ork = -ork;
          } else ork = -r[k];
          s += J[7 * k + a] * ork;
        }
      }
      acc += s;
    }
    if (c < 7) P.H[(size_t)(7 * v + a) * P.ld + 7 * v + c] = acc;
    else P.b[7 * v + a] = acc;
  }
  // off-diagonal blocks Ji' Omega Jj into the lower triangle
  for (int w = gtid; w < P.n_offd * 49; w += gsz) {
    const int o = w / 49, a = (w % 49) / 7, c = w % 7, e = P.offd[o];
    const int hi = P.hidx[P.ei[e]], hj = P.hidx[P.ej[e]];
    const double* Ji = P.Ji + 49 * (size_t)e;
    const double* Jj = P.Jj + 49 * (size_t)e;
    const double* O = P.info ? P.info + 49 * (size_t)e : nullptr;
    double s = 0;
    for (int k = 0; k < 7; ++k) {
      double ao;
      if (O) {
        ao = 0;
        for (int m = 0; m < 7; ++m) ao += Ji[7 * m + a] * O[7 * m + k];
      } else ao = Ji[7 * k + a];
      s += ao * Jj[7 * k + c];
    }
    if (hi > hj) atomicAdd(&P.H[(size_t)(7 * hi + a) * P.ld + 7 * hj + c], s);
    else atomicAdd(&P.H[(size_t)(7 * hj + c) * P.ld + 7 * hi + a], s);
  }
}

// L = H + lambda I on the filled pattern (fill tiles start at zero, padding rows get a unit diagonal), y = b, backup
__device__ void pg_begin_trial(const PgParams& P, double lambda, bool restore) {
  const int gtid = blockIdx.x * kPgThreads + threadIdx.x, gsz = gridDim.x * kPgThreads;
  for (long long idx = gtid; idx < (long long)P.n_tiles * kPgTile2; idx += gsz) {
    const int t = (int)(idx / kPgTile2), rc = (int)(idx % kPgTile2);
    const int i = P.tl_i[t] * kPgTS + rc / kPgTS, j = P.tl_j[t] * kPgTS + rc % kPgTS;
    double v = P.tl_isH[t] ? P.H[(size_t)i * P.ld + j] : 0.0;
    if (i == j) v = i < P.n ? v + lambda : 1.0;
    P.L[(size_t)i * P.ld + j] = v;
  }
  for (int i = gtid; i < P.ld; i += gsz) P.y[i] = i < P.n ? P.b[i] : 0.0;
  for (int v = gtid; v < P.nv; v += gsz) {
    const int k = P.fv[v];
    if (restore) P.est[k] = P.bak[k];
    else P.bak[k] = P.est[k];
  }
}

// panel k, step 1: the diagonal tile by warp 0 (Cholesky-Crout, lanes own rows lane / lane + 32; row 56 carries y_k)
__device__ void pg_factor_diag(const PgParams& P, int k, double* sm) {
  double (*A)[kPgLd] = reinterpret_cast<double (*)[kPgLd]>(sm);
  const int t = threadIdx.x, lane = t & 31;
  const size_t base = (size_t)k * kPgTS;
  for (int e = t; e < kPgTile2; e += kPgThreads) {
    const int r = e / kPgTS, c = e % kPgTS;
    A[r][c] = c <= r ? P.L[(base + r) * P.ld + base + c] : 0.0;
  }
  if (t < kPgTS) A[kPgTS][t] = P.y[base + t];
  __syncthreads();
  if (t < 32) {
    const int i0 = lane, i1 = lane + 32;
    const bool has1 = i1 <= kPgTS;  // rows 32..56
    int bad = 0;
    for (int j = 0; j < kPgTS; ++j) {
      double s0 = 0, s1 = 0;
      if (i0 >= j) {
        s0 = A[i0][j];
        for (int m = 0; m < j; ++m) s0 -= A[i0][m] * A[j][m];
      }
      if (has1 && i1 >= j) {
        s1 = A[i1][j];
        for (int m = 0; m < j; ++m) s1 -= A[i1][m] * A[j][m];
      }
      const double d = __shfl_sync(0xffffffffu, j < 32 ? s0 : s1, j & 31);
      if (!(d > 0) || !isfinite(d)) bad = 1;
      const double sq = sqrt(d);
      __syncwarp();  // every lane has read row j's earlier columns before column j is written
      if (i0 == j) A[i0][j] = sq;
      else if (i0 > j) A[i0][j] = s0 / sq;
      if (has1) {
        if (i1 == j) A[i1][j] = sq;
        else if (i1 > j) A[i1][j] = s1 / sq;
      }
      __syncwarp();
    }
    if (bad && lane == 0) P.flags[0] = 0;
  }
  __syncthreads();
  for (int e = t; e < kPgTile2; e += kPgThreads) {
    const int r = e / kPgTS, c = e % kPgTS;
    if (c <= r) P.L[(base + r) * P.ld + base + c] = A[r][c];
  }
  if (t < kPgTS) P.y[base + t] = A[kPgTS][t];
  __syncthreads();
}

// panel k, step 2: L_ik = A_ik L_kk^-T for one row tile (one thread per row), y_i -= L_ik y_k
__device__ void pg_trsm_tile(const PgParams& P, int k, int i, double* sm) {
  double (*Lk)[kPgLd] = reinterpret_cast<double (*)[kPgLd]>(sm);
  double (*A)[kPgLd] = reinterpret_cast<double (*)[kPgLd]>(sm + kPgTS * kPgLd);
  __shared__ double s_yk[kPgTS];
  const int t = threadIdx.x;
  const size_t bk = (size_t)k * kPgTS, bi = (size_t)i * kPgTS;
  for (int e = t; e < kPgTile2; e += kPgThreads) {
    const int r = e / kPgTS, c = e % kPgTS;
    Lk[r][c] = P.L[(bk + r) * P.ld + bk + c];
    A[r][c] = P.L[(bi + r) * P.ld + bk + c];
  }
  if (t < kPgTS) s_yk[t] = P.y[bk + t];
  __syncthreads();
  if (t < kPgTS) {
    double dot = 0;
    for (int j = 0; j < kPgTS; ++j) {
      double s = A[t][j];
      for (int m = 0; m < j; ++m) s -= A[t][m] * Lk[j][m];
      s = s / Lk[j][j];
      A[t][j] = s;
      dot += s * s_yk[j];
    }
    P.y[bi + t] -= dot;
  }
  __syncthreads();
  for (int e = t; e < kPgTile2; e += kPgThreads) {
    const int r = e / kPgTS, c = e % kPgTS;
    P.L[(bi + r) * P.ld + bk + c] = A[r][c];
  }
  __syncthreads();
}

// panel k, step 3: A_ij -= L_ik L_jk^T (14 x 14 threads, 4 x 4 outputs each)
__device__ void pg_syrk_tile(const PgParams& P, int k, int i, int j, double* sm) {
  double (*Li)[kPgLd] = reinterpret_cast<double (*)[kPgLd]>(sm);
  double (*Lj)[kPgLd] = reinterpret_cast<double (*)[kPgLd]>(sm + kPgTS * kPgLd);
  const int t = threadIdx.x;
  const size_t bk = (size_t)k * kPgTS, bi = (size_t)i * kPgTS, bj = (size_t)j * kPgTS;
  for (int e = t; e < kPgTile2; e += kPgThreads) {
    const int r = e / kPgTS, c = e % kPgTS;
    Li[r][c] = P.L[(bi + r) * P.ld + bk + c];
    Lj[r][c] = P.L[(bj + r) * P.ld + bk + c];
  }
  __syncthreads();
  if (t < 196) {
    const int r0 = 4 * (t / 14), c0 = 4 * (t % 14);
    double acc[4][4] = {};
    for (int m = 0; m < kPgTS; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = Li[r0 + q][m];
        b[q] = Lj[c0 + q][m];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[q][p] += a[q] * b[p];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (i == j && c0 + p > r0 + q) continue;  // diagonal tile: lower triangle only
        P.L[(bi + r0 + q) * P.ld + bj + c0 + p] -= acc[q][p];
      }
  }
  __syncthreads();
}

// x = L^-T y by one CTA, tile rows in reverse
__device__ void pg_back_substitute(const PgParams& P, double* sm) {
  double (*Lk)[kPgLd] = reinterpret_cast<double (*)[kPgLd]>(sm);
  __shared__ double s_x[kPgTS + 8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int k = P.T - 1; k >= 0; --k) {
    const size_t bk = (size_t)k * kPgTS;
    for (int e = t; e < kPgTile2; e += kPgThreads) {
      const int r = e / kPgTS, c = e % kPgTS;
      Lk[r][c] = P.L[(bk + r) * P.ld + bk + c];
    }
    if (t < kPgTS) s_x[t] = P.y[bk + t];
    __syncthreads();
    if (t < 32) {
      for (int j = kPgTS - 1; j >= 0; --j) {
        const double xj = s_x[j] / Lk[j][j];
        __syncwarp();
        if (lane == 0) s_x[j] = xj;
        // y_m -= L[j][m] x_j for m < j
        if (lane < j) s_x[lane] -= Lk[j][lane] * xj;
        if (lane + 32 < j) s_x[lane + 32] -= Lk[j][lane + 32] * xj;
        __syncwarp();
      }
    }
    __syncthreads();
    if (t < kPgTS) P.x[bk + t] = s_x[t];
    // y_j -= L_kj^T x_k for the tiles left of the diagonal in row k
    for (int q = P.row_off[k] + warp; q < P.row_off[k + 1]; q += kPgThreads / 32) {
      const size_t bj = (size_t)P.row_col[q] * kPgTS;
      for (int c = lane; c < kPgTS; c += 32) {
        double s = 0;
        for (int r = 0; r < kPgTS; ++r) s += P.L[(bk + r) * P.ld + bj + c] * s_x[r];
        P.y[bj + c] -= s;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPgThreads, 1) k_posegraph(PgParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double sm[];
  __shared__ double s_red[kPgThreads / 32];
  __shared__ double s_bc;
  const int G = gridDim.x, gtid = blockIdx.x * kPgThreads + threadIdx.x, gsz = G * kPgThreads;
  const bool fs = P.fix_scale != 0;
  double lambda = 0, ni = 2, chi_first = 0, chi_last = 0;
  int nBad = 0, total_iters = 0, trials = 0, any_fail = 0;
  bool ok = true;
  const int iterations = P.single_step ? 1 : P.iterations;
  for (int it = 0; it < iterations && ok; ++it) {
    pg_errors(P, P.part, s_red);
    pg_linearize(P);  // reads est only
    pg_zero_system(P);
    grid.sync();
    double currentChi = pg_grid_total(P.part, G, &s_bc);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (it == 0) chi_first = currentChi;
    pg_assemble(P);
    grid.sync();
    if (it == 0) {
      if (P.lambda_init > 0) lambda = P.lambda_init;
      else {  // computeLambdaInit: 1e-5 * max |H_jj| (every CTA computes the same maximum)
        double mx = 0;
        for (int j = threadIdx.x; j < P.n; j += kPgThreads) mx = fmax(fabs(__ldcg(&P.H[(size_t)j * P.ld + j])), mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
        __syncthreads();
        mx = 0;
        for (int w = 0; w < kPgThreads / 32; ++w) mx = fmax(mx, s_red[w]);
        lambda = 1e-5 * mx;
      }
      ni = 2;
      nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    bool restore = false;
    do {
      if (gtid == 0) P.flags[0] = 1;
      pg_begin_trial(P, lambda, restore);
      restore = false;
      grid.sync();
      for (int k = 0; k < P.T; ++k) {
        if ((int)blockIdx.x == k % G) pg_factor_diag(P, k, sm);
        if (P.trsm_off[k + 1] > P.trsm_off[k]) {
          grid.sync();
          for (int q = P.trsm_off[k] + blockIdx.x; q < P.trsm_off[k + 1]; q += G) pg_trsm_tile(P, k, P.trsm_row[q], sm);
          grid.sync();
          for (int q = P.syrk_off[k] + blockIdx.x; q < P.syrk_off[k + 1]; q += G) pg_syrk_tile(P, k, P.syrk_i[q], P.syrk_j[q], sm);
        }
        grid.sync();
      }
      if (blockIdx.x == 0) pg_back_substitute(P, sm);
      grid.sync();
      const bool ok2 = __ldcg(P.flags) != 0;
      // update (only a successful solve moves the vertices) + computeScale partials
      double sc_local = 0;
      if (ok2) {
        for (int v = gtid; v < P.nv; v += gsz) {
          const int kf = P.fv[v];
          double u[7];
          for (int a = 0; a < 7; ++a) u[a] = __ldcg(&P.x[7 * v + a]);
          P.est[kf] = s3_oplus(P.est[kf], u, fs);
        }
        for (int j = gtid; j < P.n; j += gsz) {
          const double xj = __ldcg(&P.x[j]);
          sc_local += xj * (lambda * xj + P.b[j]);
        }
      }
      {
        const double t = pg_block_sum(sc_local, s_red);
        if (threadIdx.x == 0) P.part[G + blockIdx.x] = t;
      }
      grid.sync();
      pg_errors(P, P.part + 2 * G, s_red);
      grid.sync();
      tempChi = pg_grid_total(P.part + 2 * G, G, &s_bc);
      double scale = pg_grid_total(P.part + G, G, &s_bc);
      if (!ok2) {
        tempChi = DBL_MAX;
        any_fail = 1;
      }
      rho = currentChi - tempChi;
      scale += 1e-3;
      rho /= scale;
      if (!ok2) rho = -1;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, 2. / 3.);
        lambda *= fmax(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        restore = true;
      }
      qmax++;
      trials++;
    } while (rho < 0 && qmax < 10 && !P.single_step);
    if (restore && !P.single_step) {  // pop: the rejected estimate must not survive the loop
      for (int v = gtid; v < P.nv; v += gsz) P.est[P.fv[v]] = P.bak[P.fv[v]];
      grid.sync();
    }
    ++total_iters;
    chi_last = P.single_step ? tempChi : currentChi;
    if (P.single_step) break;
    if (qmax == 10 || rho == 0) {
      ok = false;
      break;
    }
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
    else nBad = 0;
    if (nBad >= 3) ok = false;
  }
  grid.sync();
  if (P.Tcw) {  // "SE3 Pose Recovering" (src/Optimizer.cc:2624-2642)
    for (int k = gtid; k < P.K; k += gsz) {
      double R[9];
      s3_q2R(P.est[k].q, R);
      const double f = 1. / P.est[k].s;
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) P.Tcw[12 * (size_t)k + 4 * r + c] = R[3 * r + c];
        P.Tcw[12 * (size_t)k + 4 * r + 3] = P.est[k].t[r] * f;
      }
    }
  }
  if (gtid == 0) {
    VieoPoseGraphStats& st = *P.stats;
    st.chi2_initial = chi_first;
    st.chi2_final = chi_last;
    st.lambda_final = P.single_step ? P.lambda_init : lambda;
    st.iterations = total_iters;
    st.trials = trials;
    st.n_free = P.nv;
    st.ok = !any_fail;
  }
}

__global__ void __launch_bounds__(256) k_pg_correct_points(int n, const float* __restrict__ Pw, const int* __restrict__ ref,
                                                           const Sim3d* __restrict__ before, const Sim3d* __restrict__ after,
                                                           float* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const Sim3d Srw = before[ref[i]];
  const Sim3d Swr = s3_inv(after[ref[i]]);
  const double P[3] = {(double)Pw[3 * i], (double)Pw[3 * i + 1], (double)Pw[3 * i + 2]};
  double Pr[3], Pc[3];
  s3_map(Srw, P, Pr);
  s3_map(Swr, Pr, Pc);
  for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)Pc[k];
}

namespace {

struct DevBuf {  // frees on scope exit
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
};

int pg_run(int K, const VieoSim3* Scw, const uint8_t* fixed, int fix_scale, int E, const int32_t* ei, const int32_t* ej,
           const VieoSim3* Sji, const double* info, int iterations, double lambda_init, int single_step, VieoSim3* Scw_out,
           double* Tcw_out, VieoPoseGraphStats* stats, double* H_out, double* b_out, int device) {
  VIEO_ARG(K > 0 && Scw && fixed && Scw_out && stats, "vertices / outputs missing");
  VIEO_ARG(E >= 0 && (E == 0 || (ei && ej && Sji)), "edges missing");
  VIEO_ARG(iterations >= 0, "iterations < 0");
  VIEO_ARG((long long)E * 14 < (1ll << 31), "too many edges");
  int rc = use_device(device);
  if (rc) return rc;
  // ---- host plan: active set (sparse_optimizer.cpp:199-267), index mapping, incidence lists, tile pattern ----
  std::vector<uint8_t> act(std::max(E, 1), 0), touched(K, 0);
  for (int e = 0; e < E; ++e) {
    VIEO_ARG(ei[e] >= 0 && ei[e] < K && ej[e] >= 0 && ej[e] < K, "edge vertex out of range");
    VIEO_ARG(ei[e] != ej[e], "edge connects a vertex to itself");
    act[e] = !(fixed[ei[e]] && fixed[ej[e]]);
    if (act[e]) touched[ei[e]] = touched[ej[e]] = 1;
  }
  std::vector<int> hidx(K, -1), fv;
  for (int k = 0; k < K; ++k)
    if (touched[k] && !fixed[k]) {
      hidx[k] = (int)fv.size();
      fv.push_back(k);
    }
  const int nv = (int)fv.size(), n = 7 * nv;
  memset(stats, 0, sizeof(*stats));
  stats->ok = 1;
  if (nv == 0 || iterations == 0) {  // nothing to optimise: g2o returns with the estimates untouched
    memcpy(Scw_out, Scw, sizeof(VieoSim3) * (size_t)K);
    if (Tcw_out)
      for (int k = 0; k < K; ++k) {
        double R[9];
        s3_q2R(Scw[k].q, R);
        const double f = 1. / Scw[k].s;
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) Tcw_out[12 * (size_t)k + 4 * r + c] = R[3 * r + c];
          Tcw_out[12 * (size_t)k + 4 * r + 3] = Scw[k].t[r] * f;
        }
      }
    return VIEO_OK;
  }
  const int T = (nv + kPgVT - 1) / kPgVT, ld = T * kPgTS;
  std::vector<int> vptr(nv + 1, 0), vinc, offd;
  for (int e = 0; e < E; ++e) {
    if (!act[e]) continue;
    if (hidx[ei[e]] >= 0) vptr[hidx[ei[e]] + 1]++;
    if (hidx[ej[e]] >= 0) vptr[hidx[ej[e]] + 1]++;
    if (hidx[ei[e]] >= 0 && hidx[ej[e]] >= 0) offd.push_back(e);
  }
  for (int v = 0; v < nv; ++v) vptr[v + 1] += vptr[v];
  vinc.resize(std::max(vptr[nv], 1));
  {
    std::vector<int> fill(vptr.begin(), vptr.end() - 1);
    for (int e = 0; e < E; ++e) {  // ascending edge index per vertex: the order g2o adds the edges' contributions in
      if (!act[e]) continue;
      if (hidx[ei[e]] >= 0) vinc[fill[hidx[ei[e]]]++] = (e << 1);
      if (hidx[ej[e]] >= 0) vinc[fill[hidx[ej[e]]]++] = (e << 1) | 1;
    }
  }
  std::vector<uint8_t> nzH((size_t)T * T, 0), nzL;
  for (int t = 0; t < T; ++t) nzH[(size_t)t * T + t] = 1;
  for (int e : offd) {
    const int a = hidx[ei[e]] / kPgVT, b = hidx[ej[e]] / kPgVT;
    nzH[(size_t)std::max(a, b) * T + std::min(a, b)] = 1;
  }
  nzL = nzH;
  std::vector<int> trsm_off(T + 1, 0), trsm_row, syrk_off(T + 1, 0), syrk_i, syrk_j, row_off(T + 1, 0), row_col;
  for (int k = 0; k < T; ++k) {  // tile-level symbolic factorisation + the per-panel work lists
    std::vector<int> rows;
    for (int i = k + 1; i < T; ++i)
      if (nzL[(size_t)i * T + k]) rows.push_back(i);
    for (int a : rows) {
      trsm_row.push_back(a);
      for (int b : rows)
        if (a >= b) {
          nzL[(size_t)a * T + b] = 1;
          syrk_i.push_back(a);
          syrk_j.push_back(b);
        }
    }
    trsm_off[k + 1] = (int)trsm_row.size();
    syrk_off[k + 1] = (int)syrk_i.size();
  }
  std::vector<int> tl_i, tl_j;
  std::vector<uint8_t> tl_isH;
  for (int i = 0; i < T; ++i) {
    for (int j = 0; j <= i; ++j)
      if (nzL[(size_t)i * T + j]) {
        tl_i.push_back(i);
        tl_j.push_back(j);
        tl_isH.push_back(nzH[(size_t)i * T + j]);
        if (j < i) row_col.push_back(j);
      }
    row_off[i + 1] = (int)row_col.size();
  }
  const int n_tiles = (int)tl_i.size();
  // ---- one packed upload of the integer plan ----
  std::vector<int> ints;
  auto put = [&](const std::vector<int>& v) {
    const size_t o = ints.size();
    ints.insert(ints.end(), v.begin(), v.end());
    if (v.empty()) ints.push_back(0);
    return o;
  };
  const size_t o_ei = put(std::vector<int>(ei, ei + E)), o_ej = put(std::vector<int>(ej, ej + E)), o_hidx = put(hidx), o_fv = put(fv),
               o_vptr = put(vptr), o_vinc = put(vinc), o_offd = put(offd), o_tli = put(tl_i), o_tlj = put(tl_j),
               o_trsm_off = put(trsm_off), o_trsm_row = put(trsm_row), o_syrk_off = put(syrk_off), o_syrk_i = put(syrk_i),
               o_syrk_j = put(syrk_j), o_row_off = put(row_off), o_row_col = put(row_col);
  int n_sm = 0;
  VIEO_CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  static SmemOptIn optin;
  VIEO_CK(smem_opt_in(k_posegraph, kPgSmem, optin));
  int per_sm = 0;
  VIEO_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_posegraph, kPgThreads, kPgSmem));
  VIEO_ARG(per_sm >= 1, "k_posegraph does not fit an SM");
  const int G = n_sm;
  DevBuf d_ints, d_bytes, d_sim, d_info, d_f64, d_H, d_L, d_stats, d_tcw;
  const size_t n_bytes = (size_t)std::max(E, 1) + (size_t)n_tiles;
  VIEO_CK(d_ints.alloc(sizeof(int) * (ints.size() + 4)));
  VIEO_CK(d_bytes.alloc(n_bytes));
  VIEO_CK(d_sim.alloc(sizeof(Sim3d) * ((size_t)2 * K + std::max(E, 1))));
  if (info) VIEO_CK(d_info.alloc(sizeof(double) * 49 * (size_t)E));
  const size_t f64_count = (size_t)std::max(E, 1) * (7 + 49 + 49) + 3 * (size_t)ld + (size_t)n + 3 * (size_t)G;
  VIEO_CK(d_f64.alloc(sizeof(double) * f64_count));
  VIEO_CK(d_H.alloc(sizeof(double) * (size_t)ld * ld));
  VIEO_CK(d_L.alloc(sizeof(double) * (size_t)ld * ld));
  VIEO_CK(d_stats.alloc(sizeof(VieoPoseGraphStats)));
  if (Tcw_out) VIEO_CK(d_tcw.alloc(sizeof(double) * 12 * (size_t)K));
  cudaStream_t st = call_scratch(device) ? call_scratch(device)->st : nullptr;
  VIEO_ARG(st != nullptr, "no call stream");
  int* di = (int*)d_ints.p;
  VIEO_CK(cudaMemcpyAsync(di, ints.data(), sizeof(int) * ints.size(), cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemsetAsync(di + ints.size(), 0, sizeof(int) * 4, st));
  uint8_t* db = (uint8_t*)d_bytes.p;
  VIEO_CK(cudaMemcpyAsync(db, act.data(), std::max(E, 1), cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(db + std::max(E, 1), tl_isH.data(), n_tiles, cudaMemcpyHostToDevice, st));
  Sim3d* ds = (Sim3d*)d_sim.p;
  VIEO_CK(cudaMemcpyAsync(ds, Scw, sizeof(Sim3d) * (size_t)K, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(ds + K, Scw, sizeof(Sim3d) * (size_t)K, cudaMemcpyHostToDevice, st));
  if (E) VIEO_CK(cudaMemcpyAsync(ds + 2 * (size_t)K, Sji, sizeof(Sim3d) * (size_t)E, cudaMemcpyHostToDevice, st));
  if (info) VIEO_CK(cudaMemcpyAsync(d_info.p, info, sizeof(double) * 49 * (size_t)E, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemsetAsync(d_f64.p, 0, sizeof(double) * f64_count, st));
  PgParams P{};
  P.K = K; P.E = E; P.nv = nv; P.n = n; P.T = T; P.ld = ld;
  P.fix_scale = fix_scale; P.iterations = iterations; P.single_step = single_step; P.lambda_init = lambda_init;
  P.est = ds; P.bak = ds + K; P.meas = ds + 2 * (size_t)K; P.info = (const double*)d_info.p;
  P.ei = di + o_ei; P.ej = di + o_ej; P.hidx = di + o_hidx; P.fv = di + o_fv; P.act = db;
  double* df = (double*)d_f64.p;
  P.err = df; df += (size_t)std::max(E, 1) * 7;
  P.Ji = df; df += (size_t)std::max(E, 1) * 49;
  P.Jj = df; df += (size_t)std::max(E, 1) * 49;
  P.b = df; df += n;
  P.x = df; df += ld;
  P.y = df; df += ld;
  df += ld;  // spare
  P.part = df;
  P.vptr = di + o_vptr; P.vinc = di + o_vinc; P.offd = di + o_offd; P.n_offd = (int)offd.size();
  P.H = (double*)d_H.p; P.L = (double*)d_L.p;
  P.tl_i = di + o_tli; P.tl_j = di + o_tlj; P.tl_isH = db + std::max(E, 1); P.n_tiles = n_tiles;
  P.trsm_off = di + o_trsm_off; P.trsm_row = di + o_trsm_row; P.syrk_off = di + o_syrk_off; P.syrk_i = di + o_syrk_i;
  P.syrk_j = di + o_syrk_j; P.row_off = di + o_row_off; P.row_col = di + o_row_col;
  P.flags = di + ints.size();
  P.stats = (VieoPoseGraphStats*)d_stats.p;
  P.Tcw = (double*)d_tcw.p;
  void* args[] = {&P};
  {
    // a thread bound to an SM partition (green context) owns fewer SMs than the device reports: a cooperative grid that does
    // not fit is refused, not queued — halve it until it is accepted (the kernel is written for any grid size >= 1)
    int g_try = G;
    cudaError_t le;
    for (;;) {
      le = cudaLaunchCooperativeKernel((const void*)k_posegraph, dim3(g_try), dim3(kPgThreads), args, kPgSmem, st);
      if (le != cudaErrorCooperativeLaunchTooLarge || g_try == 1) break;
      (void)cudaGetLastError();
      g_try = std::max(1, g_try / 2);
    }
    VIEO_CK(le);
  }
  VIEO_CK(cudaMemcpyAsync(Scw_out, ds, sizeof(Sim3d) * (size_t)K, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(stats, d_stats.p, sizeof(*stats), cudaMemcpyDeviceToHost, st));
  if (Tcw_out) VIEO_CK(cudaMemcpyAsync(Tcw_out, d_tcw.p, sizeof(double) * 12 * (size_t)K, cudaMemcpyDeviceToHost, st));
  std::vector<double> Hh;
  if (H_out) {
    Hh.resize((size_t)ld * ld);
    VIEO_CK(cudaMemcpyAsync(Hh.data(), d_H.p, sizeof(double) * Hh.size(), cudaMemcpyDeviceToHost, st));
  }
  if (b_out) VIEO_CK(cudaMemcpyAsync(b_out, P.b, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  if (H_out) {  // symmetric n x n from the structural tiles of the lower triangle
    std::fill(H_out, H_out + (size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        if (!nzH[(size_t)(i / kPgTS) * T + j / kPgTS]) continue;
        const double v = Hh[(size_t)i * ld + j];
        H_out[(size_t)i * n + j] = v;
        H_out[(size_t)j * n + i] = v;
      }
  }
  return VIEO_OK;
}

}  // namespace
}  // namespace vieo

extern "C" {

int vieo_essential_graph_optimize(int n_vertices, const VieoSim3* Scw, const uint8_t* fixed, int fix_scale, int n_edges,
                                  const int32_t* edge_i, const int32_t* edge_j, const VieoSim3* Sji, const double* info,
                                  int iterations, double lambda_init, VieoSim3* Scw_out, double* Tcw_out,
                                  VieoPoseGraphStats* stats, int device) {
  return vieo::pg_run(n_vertices, Scw, fixed, fix_scale, n_edges, edge_i, edge_j, Sji, info, iterations, lambda_init, 0, Scw_out,
                      Tcw_out, stats, nullptr, nullptr, device);
}

int vieo_essential_graph_debug_step(int n_vertices, const VieoSim3* Scw, const uint8_t* fixed, int fix_scale, int n_edges,
                                    const int32_t* edge_i, const int32_t* edge_j, const VieoSim3* Sji, const double* info,
                                    double lambda, VieoSim3* Scw_out, VieoPoseGraphStats* stats, double* H_out, double* b_out,
                                    int device) {
  VIEO_ARG(lambda > 0, "the debug step needs an explicit lambda");
  return vieo::pg_run(n_vertices, Scw, fixed, fix_scale, n_edges, edge_i, edge_j, Sji, info, 1, lambda, 1, Scw_out, nullptr, stats,
                      H_out, b_out, device);
}

int vieo_essential_graph_correct_points(int n_points, const float* Pw, const int32_t* ref, int n_vertices,
                                        const VieoSim3* Scw_before, const VieoSim3* Scw_after, float* Pw_out, int device) {
  using namespace vieo;
  VIEO_ARG(n_points >= 0 && n_vertices > 0 && Scw_before && Scw_after, "bad arguments");
  if (n_points == 0) return VIEO_OK;
  VIEO_ARG(Pw && ref && Pw_out, "points missing");
  for (int i = 0; i < n_points; ++i) VIEO_ARG(ref[i] >= 0 && ref[i] < n_vertices, "reference keyframe out of range");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  VIEO_ARG(cs != nullptr, "no call scratch");
  float* dP = (float*)cs->get(0, sizeof(float) * 3 * (size_t)n_points);
  float* dO = (float*)cs->get(1, sizeof(float) * 3 * (size_t)n_points);
  int* dR = (int*)cs->get(2, sizeof(int) * (size_t)n_points);
  Sim3d* dS = (Sim3d*)cs->get(3, sizeof(Sim3d) * 2 * (size_t)n_vertices);
  VIEO_ARG(dP && dO && dR && dS, "device allocation failed");
  VIEO_CK(cudaMemcpyAsync(dP, Pw, sizeof(float) * 3 * (size_t)n_points, cudaMemcpyHostToDevice, cs->st));
  VIEO_CK(cudaMemcpyAsync(dR, ref, sizeof(int) * (size_t)n_points, cudaMemcpyHostToDevice, cs->st));
  VIEO_CK(cudaMemcpyAsync(dS, Scw_before, sizeof(Sim3d) * (size_t)n_vertices, cudaMemcpyHostToDevice, cs->st));
  VIEO_CK(cudaMemcpyAsync(dS + n_vertices, Scw_after, sizeof(Sim3d) * (size_t)n_vertices, cudaMemcpyHostToDevice, cs->st));
  k_pg_correct_points<<<(n_points + 255) / 256, 256, 0, cs->st>>>(n_points, dP, dR, dS, dS + n_vertices, dO);
  VIEO_CK(cudaGetLastError());
  VIEO_CK(cudaMemcpyAsync(Pw_out, dO, sizeof(float) * 3 * (size_t)n_points, cudaMemcpyDeviceToHost, cs->st));
  VIEO_CK(cudaStreamSynchronize(cs->st));
  return VIEO_OK;
}

}  // extern "C"

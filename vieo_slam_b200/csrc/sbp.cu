// Guided searches of the tracking thread on the device: ORBmatcher::SearchByProjection(Frame&, const Frame& last, ...)
// (src/ORBmatcher.cc:1303-1467) and ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, ...) (:230-335) with the
// frame grid of src/FrameBase.cpp:95-174.
//
// The reference loop is sequential in the map points: a keypoint taken by an earlier point (with Observations() > 0) is
// skipped by every later one.  Only that claim step is order dependent, so the work is split:
//   phase A  (whole CTA; 8 list CTAs per frame, each with its own copy)  AssignFeaturesToGrid: counting sort of the keypoints into the 64 x 48 cells,
//            cell lists ordered by keypoint index like the reference's push_back order; lists live in shared memory.
//   phase B  (k_sbp_lists: 8 CTAs x 8 warps per frame, one query per warp at a time)  projection (LAST_FRAME), GetFeaturesInArea in the reference's
//            candidate order (ix-major, iy, insertion order), level band, window and stereo-ur gates, 256-bit Hamming
//            distance of every surviving candidate -> compact per-query list (dist << 16 | keypoint) in HBM scratch.
//   phase C  (k_sbp_claim: one warp per frame)  walks the queries in order: masks the candidates whose keypoint is already claimed (bitmap in
//            shared memory), arg-min / second-min by warp redux on (dist, position) keys (strict '<' of the reference ==
//            lexicographic order), acceptance tests, claim.  Lists longer than kListCap fall back to a re-enumeration.
//            LAST_FRAME then applies the rotation-histogram check.
// No atomics in anything that decides an index: results are bit-reproducible and equal the oracle's.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

using namespace vieo;

namespace {

constexpr int kCols = 64, kRows = 48, kCells = kCols * kRows;
constexpr int kMaxKp = VIEO_SBP_MAX_KEYPOINTS;
constexpr int kSbpWarps = 16;
constexpr int kListCap = 64;  // candidates kept per query in scratch (two per lane in phase C)
constexpr int TH_HIGH = 100, HISTO_LENGTH = 30;

struct SbpShared {
  uint16_t cell_start[kCells + 1];
  uint16_t cell_items[kMaxKp];
  uint32_t blocked[kMaxKp / 32];
  int cnt[kCells];  // phase A counters (reused as scan input)
  int total;
  int hist[HISTO_LENGTH];
  uint8_t oct[kMaxKp];  // claim kernel: keypoint octaves staged once (the ratio test reads two per accepted query)
};

struct Cand {  // one query's search window
  float x, y, r;
  int minlevel, maxlevel;
  float ur;  // predicted right coordinate for the stereo gate (NaN-free sentinel: use_ur == 0 switches the gate off)
  int use_ur;
};

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint8_t* __restrict__ b) {
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(b)), b1 = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Eigen::Quaternion::_transformVector (Sophus::SO3d * point): v + w (2 q x v) + q x (2 q x v)
__device__ __forceinline__ void qrot(const double* q, const double* v, double* o) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  double uv0 = y * v[2] - z * v[1], uv1 = z * v[0] - x * v[2], uv2 = x * v[1] - y * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  const double c0 = y * uv2 - z * uv1, c1 = z * uv0 - x * uv2, c2 = x * uv1 - y * uv0;
  o[0] = v[0] + w * uv0 + c0;
  o[1] = v[1] + w * uv1 + c1;
  o[2] = v[2] + w * uv2 + c2;
}

// GetFeaturesInArea (src/FrameBase.cpp:95-142), warp-cooperative.  Lanes take the window's cells in the reference's
// order (32 per pass); sink(pos, kp) is called for every candidate that passes the level / window / ur gates, pos = its
// rank in the reference's candidate order.  Returns the number of candidates (warp-uniform).
struct GridGeom {  // gridinfo_: image bounds and inverse cell sizes
  float minx, miny, grid_winv, grid_hinv;
};
// the stereo gate of the tracking searches (:1403-1408 / :292-297): |ur - uright[j]| <= r for keypoints with a right match
struct UrGate {
  const float* __restrict__ uright;
  float ur, r;
  int on;
  __device__ __forceinline__ bool operator()(int j, const VieoKeyPoint&) const {
    if (!on) return true;  // the relocalisation search has no stereo gate
    const float urj = uright[j];
    if (urj > 0) {
      const float er = fabsf(ur - urj);
      if (er > r) return false;
    }
    return true;
  }
};
template <class Gate, class Sink>
__device__ __forceinline__ int enumerate_gated(const GridGeom& F, const SbpShared& S, const VieoKeyPoint* __restrict__ kps,
                                               const Cand& c, int lane, const Gate& gate, Sink&& sink) {
  const int min_cellx = max(0, (int)floorf((c.x - F.minx - c.r) * F.grid_winv));
  if (min_cellx >= kCols) return 0;
  const int max_cellx = min(kCols - 1, (int)ceilf((c.x - F.minx + c.r) * F.grid_winv));
  if (max_cellx < 0) return 0;
  const int min_celly = max(0, (int)floorf((c.y - F.miny - c.r) * F.grid_hinv));
  if (min_celly >= kRows) return 0;
  const int max_celly = min(kRows - 1, (int)ceilf((c.y - F.miny + c.r) * F.grid_hinv));
  if (max_celly < 0) return 0;
  const bool bchecklevel = (c.minlevel > 0) || (c.maxlevel >= 0);
  const int ny = max_celly - min_celly + 1, ncell = (max_cellx - min_cellx + 1) * ny;
  if (ncell <= 0) return 0;
  auto passes = [&](int j) {
    const VieoKeyPoint kp = kps[j];
    if (bchecklevel) {
      if (kp.octave < c.minlevel) return false;
      if (c.maxlevel >= 0 && kp.octave > c.maxlevel) return false;
    }
    const float distx = kp.x - c.x, disty = kp.y - c.y;
    if (!(fabsf(distx) < c.r && fabsf(disty) < c.r)) return false;
    return gate(j, kp);
  };
  int base = 0;
  for (int c0 = 0; c0 < ncell; c0 += 32) {
    const int ci = c0 + lane;
    int s = 0, e = 0;
    if (ci < ncell) {
      const int cell = (min_cellx + ci / ny) * kRows + min_celly + ci % ny;
      s = S.cell_start[cell];
      e = S.cell_start[cell + 1];
    }
    int n = 0;
    for (int k = s; k < e; ++k) n += passes(S.cell_items[k]) ? 1 : 0;
    const int incl = warp_incl_scan(n, lane);
    int pos = base + incl - n;
    if (n > 0)
      for (int k = s; k < e; ++k) {
        const int j = S.cell_items[k];
        if (passes(j)) sink(pos++, j);
      }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  return base;
}
template <class Sink>
__device__ __forceinline__ int enumerate(const VieoSbpFrame& F, const SbpShared& S, const VieoKeyPoint* __restrict__ kps,
                                         const float* __restrict__ uright, const Cand& c, int lane, Sink&& sink) {
  return enumerate_gated(GridGeom{F.minx, F.miny, F.grid_winv, F.grid_hinv}, S, kps, c, lane, UrGate{uright, c.ur, c.r, c.use_ur},
                         static_cast<Sink&&>(sink));
}

// FrameBase::AssignFeaturesToGrid / PosInGrid (src/FrameBase.cpp:143-170) by the whole CTA: counting sort of the keypoints
// into the cells, every cell's list ordered by keypoint index (= the reference's push_back order).  S.cnt must be zero on
// entry (and the zeroing visible: __syncthreads before the call); ends with a __syncthreads.
__device__ __forceinline__ void build_grid(SbpShared& S, const GridGeom& G, const VieoKeyPoint* __restrict__ kps, int N,
                                           int tid, int T) {
  for (int i = tid; i < N; i += T) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, G.minx), G.grid_winv));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, G.miny), G.grid_hinv));
    if (px < 0 || px >= kCols || py < 0 || py >= kRows) continue;
    atomicAdd(&S.cnt[px * kRows + py], 1);
  }
  __syncthreads();
  warp0_excl_scan(S.cnt, kCells, &S.total);
  __syncthreads();
  for (int i = tid; i < kCells; i += T) S.cell_start[i] = (uint16_t)S.cnt[i];
  if (tid == 0) S.cell_start[kCells] = (uint16_t)S.total;
  __syncthreads();
  // fill (unordered inside a cell), then order every cell's few entries by keypoint index = push_back order
  for (int i = tid; i < N; i += T) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, G.minx), G.grid_winv));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, G.miny), G.grid_hinv));
    if (px < 0 || px >= kCols || py < 0 || py >= kRows) continue;
    const int slot = atomicAdd(&S.cnt[px * kRows + py], 1);
    S.cell_items[slot] = (uint16_t)i;
  }
  __syncthreads();
  for (int cidx = tid; cidx < kCells; cidx += T) {
    const int s = S.cell_start[cidx], e = S.cell_start[cidx + 1];
    for (int a = s + 1; a < e; ++a) {
      const uint16_t v = S.cell_items[a];
      int b = a - 1;
      while (b >= s && S.cell_items[b] > v) {
        S.cell_items[b + 1] = S.cell_items[b];
        --b;
      }
      S.cell_items[b + 1] = v;
    }
  }
  __syncthreads();
}

struct QueryIn {
  const double* Xw;
  const int32_t* level;
  const float* angle;
  const float* proj;
  const float* viewcos;
  const float* depth;
  const uint8_t* desc;
  const uint8_t* flags;
  // RELOC: mfMaxDistance / mfMinDistance per query, one VieoSbpReloc per frame, predicted level out (nullable)
  const float* max_dist;
  const float* min_dist;
  const VieoSbpReloc* reloc;
  int32_t* level_out;
};

// window of query q (frame-relative qi = q - F.q_begin); false: the query is skipped before the search
__device__ __forceinline__ bool make_window(int mode, const VieoSbpFrame& F, const QueryIn& Q, int q, bool fwd, bool bwd,
                                            Cand& c, const VieoSbpReloc* R = nullptr, const double* twc = nullptr) {
  c.use_ur = 1;
  if (mode == VIEO_SBP_RELOC) {
    // SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist, th_far_pts) (src/ORBmatcher.cc:1490-1536)
    const double* Xw = Q.Xw + 3 * (size_t)q;
    double x3Dcr[3];
    qrot(F.qcw, Xw, x3Dcr);
    x3Dcr[0] += F.tcw[0]; x3Dcr[1] += F.tcw[1]; x3Dcr[2] += F.tcw[2];
    if (F.th_far > 0 && x3Dcr[2] > (double)F.th_far) return false;
    const float invzc = (float)(1.0 / x3Dcr[2]);  // no "invzc < 0" skip in this routine
    const float xn = __fmul_rn((float)x3Dcr[0], invzc), yn = __fmul_rn((float)x3Dcr[1], invzc);
    const float u = __fadd_rn(__fmul_rn(F.fx, xn), F.cx), v = __fadd_rn(__fmul_rn(F.fy, yn), F.cy);
    if (!(u >= F.minx && u < F.maxx && v >= F.miny && v < F.maxy)) return false;
    const double POx = Xw[0] - twc[0], POy = Xw[1] - twc[1], POz = Xw[2] - twc[2];
    const float dist3D = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(POx, POx), __dmul_rn(POy, POy)), __dmul_rn(POz, POz)));
    const float maxD = Q.max_dist[q];
    if (dist3D < __fmul_rn(0.8f, Q.min_dist[q]) || dist3D > __fmul_rn(1.2f, maxD)) return false;
    // MapPoint::PredictScale by counting thresholds (vieo_frustum_level_table): ratio = mfMaxDistance / dist3D
    const float ratio = __fdiv_rn(maxD, dist3D);
    int lvl = 0;
    for (int k = 1; k < F.n_levels; ++k) lvl += ratio >= R->level_ratio[k] ? 1 : 0;
    if (ratio != ratio) lvl = 0;
    if (Q.level_out) Q.level_out[q] = lvl;
    c.x = u; c.y = v;
    c.r = __fmul_rn(F.th, F.scale[lvl]);
    c.minlevel = lvl - 1; c.maxlevel = lvl + 1;
    c.ur = 0.f;
    c.use_ur = 0;
    return true;
  }
  if (mode == VIEO_SBP_LAST_FRAME) {
    double x3Dr[3];
    qrot(F.qcw, Q.Xw + 3 * (size_t)q, x3Dr);
    x3Dr[0] += F.tcw[0]; x3Dr[1] += F.tcw[1]; x3Dr[2] += F.tcw[2];
    if (F.th_far > 0 && x3Dr[2] > (double)F.th_far) return false;
    const float xc = (float)x3Dr[0], yc = (float)x3Dr[1];
    const float invzc = (float)(1.0 / x3Dr[2]);
    if (invzc < 0) return false;
    const float xn = __fmul_rn(xc, invzc), yn = __fmul_rn(yc, invzc);
    const float u = __fadd_rn(__fmul_rn(F.fx, xn), F.cx), v = __fadd_rn(__fmul_rn(F.fy, yn), F.cy);
    if (!(u >= F.minx && u < F.maxx && v >= F.miny && v < F.maxy)) return false;
    const int oct = Q.level[q];
    c.x = u; c.y = v;
    c.r = __fmul_rn(F.th, F.scale[oct]);
    if (fwd) { c.minlevel = 0; c.maxlevel = oct; }
    else if (bwd) { c.minlevel = oct; c.maxlevel = -1; }
    else { c.minlevel = oct - 1; c.maxlevel = oct + 1; }
    c.ur = __fsub_rn(u, __fmul_rn(F.bf, invzc));
    return true;
  }
  const int lvl = Q.level[q];
  if (lvl < 0) return false;  // !btrack_inview_ (:244): the level k_frustum leaves for a point that is not in view
  if (F.th_far > 0 && Q.depth[q] > F.th_far) return false;
  float r = (double)Q.viewcos[q] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos (:337-342)
  if ((double)F.th != 1.0) r = __fmul_rn(r, F.th);
  c.x = Q.proj[3 * (size_t)q]; c.y = Q.proj[3 * (size_t)q + 1];
  c.r = __fmul_rn(r, F.scale[lvl]);
  c.minlevel = lvl - 1; c.maxlevel = lvl;
  c.ur = Q.proj[3 * (size_t)q + 2];
  return true;
}

// Twcr = Tcrw.inverse() (Sophus SE3::inverse): translation = Rc^-1 * (tc * -1) — the camera centre of the RELOC search (:1476-1478)
__device__ __forceinline__ void camera_centre(const VieoSbpFrame& F, double twc[3]) {
  const double qci[4] = {F.qcw[0], -F.qcw[1], -F.qcw[2], -F.qcw[3]};
  const double nt[3] = {F.tcw[0] * -1.0, F.tcw[1] * -1.0, F.tcw[2] * -1.0};
  qrot(qci, nt, twc);
}

// frame constants of the LAST_FRAME search: forward / backward motion along the optical axis (:1314-1323)
__device__ __forceinline__ void motion_flags(int mode, const VieoSbpFrame& F, bool& fwd, bool& bwd) {
  fwd = bwd = false;
  if (mode == VIEO_SBP_LAST_FRAME) {
    // Tlrcr = Tlrw * Tcrw^-1: translation = Rl (-(Rc^-1 tc)) + tl (:1314-1319)
    const double qci[4] = {F.qcw[0], -F.qcw[1], -F.qcw[2], -F.qcw[3]};
    double a[3], b3[3];
    qrot(qci, F.tcw, a);
    a[0] = -a[0]; a[1] = -a[1]; a[2] = -a[2];
    qrot(F.qlw, a, b3);
    const double tz = b3[2] + F.tlw[2];
    fwd = tz > (double)F.b && !F.mono;
    bwd = -tz > (double)F.b && !F.mono;
  }
}

// Phases A + B.  kListCtas CTAs per frame (grid = (kListCtas, frames)): each builds the frame's grid in its own shared
// memory (a few microseconds) and lists the candidates of its share of the queries, one query per warp at a time.  Round 1
// ran A, B and C in ONE 512-thread CTA per frame: 0.86 waves at 7 % of the warp slots, holding every register of 128 SMs for
// the whole search, so nothing of the other streams could run beside it (profiles/r02k_marginal.txt: the two searches cost the
// step 1.7 of their 2.0 ms).
constexpr int kListCtas = 8, kListWarps = 8;
__global__ void __launch_bounds__(kListWarps * 32) k_sbp_lists(int mode, const VieoSbpFrame* __restrict__ frames,
                                                               const VieoKeyPoint* __restrict__ kps_all,
                                                               const float* __restrict__ ur_all,
                                                               const uint8_t* __restrict__ desc_all, QueryIn Q,
                                                               uint32_t* __restrict__ lists, int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SbpShared& S = *reinterpret_cast<SbpShared*>(smem_raw);
  __shared__ VieoSbpFrame F;
  const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x;
  for (int i = tid; i < (int)(sizeof(VieoSbpFrame) / 4); i += T)
    reinterpret_cast<uint32_t*>(&F)[i] = reinterpret_cast<const uint32_t*>(frames + f)[i];
  for (int i = tid; i < kCells; i += T) S.cnt[i] = 0;
  __syncthreads();
  const int N = F.n_kp, nq = F.n_q;
  if (N > kMaxKp || N < 0 || nq < 0) return;  // reported by the claim kernel
  if ((int)blockIdx.x * kListWarps >= nq) return;
  const VieoKeyPoint* kps = kps_all + F.kp_begin;
  const float* uright = ur_all + F.kp_begin;
  const uint8_t* desc = desc_all + 32 * (size_t)F.kp_begin;
  // ---- phase A: AssignFeaturesToGrid / PosInGrid (src/FrameBase.cpp:143-170) ------------------------------------------
  build_grid(S, GridGeom{F.minx, F.miny, F.grid_winv, F.grid_hinv}, kps, N, tid, T);
  bool fwd, bwd;
  motion_flags(mode, F, fwd, bwd);
  double twc[3] = {0, 0, 0};
  const VieoSbpReloc* R = mode == VIEO_SBP_RELOC ? Q.reloc + f : nullptr;
  if (mode == VIEO_SBP_RELOC) camera_centre(F, twc);
  uint32_t* flist = lists + (size_t)F.q_begin * kListCap;
  int32_t* fcnt = counts + F.q_begin;
  // ---- phase B: candidate lists -----------------------------------------------------------------------------------------
  for (int qi = blockIdx.x * kListWarps + warp; qi < nq; qi += kListCtas * kListWarps) {
    const int q = F.q_begin + qi;
    Cand c;
    int n = 0;
    if (make_window(mode, F, Q, q, fwd, bwd, c, R, twc)) {
      const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(Q.desc + 32 * (size_t)q));
      const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(Q.desc + 32 * (size_t)q) + 1);
      uint32_t* L = flist + (size_t)qi * kListCap;
      n = enumerate(F, S, kps, uright, c, lane, [&](int pos, int j) {
        if (pos < kListCap) L[pos] = ((uint32_t)hamming256(d0, d1, desc + 32 * (size_t)j) << 16) | (uint32_t)j;
      });
    } else {
      n = -1;  // skipped before the search
    }
    if (lane == 0) fcnt[qi] = n;
  }
}

// Phase C: the sequential claim pass, ONE WARP per frame (grid = frames, 32 threads): tiny footprint, so the pass — whose
// cost is the latency of nq dependent iterations — runs beside the other streams' kernels.  The frame grid is only needed by
// the overflow path (a list longer than kListCap is re-enumerated with the claim filter): it is built lazily, by this warp,
// the first time such a query appears.
__global__ void __launch_bounds__(32, 1) k_sbp_claim(int mode, const VieoSbpFrame* __restrict__ frames,
                                                  const VieoKeyPoint* __restrict__ kps_all, const float* __restrict__ ur_all,
                                                  const uint8_t* __restrict__ desc_all, QueryIn Q,
                                                  const uint8_t* __restrict__ kp_blocked, int32_t* __restrict__ kp_match,
                                                  int32_t* __restrict__ q_match, int32_t* __restrict__ q_dist,
                                                  int32_t* __restrict__ n_matches, const uint32_t* __restrict__ lists,
                                                  int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SbpShared& S = *reinterpret_cast<SbpShared*>(smem_raw);
  __shared__ VieoSbpFrame F;
  const int f = blockIdx.x, lane = threadIdx.x;
  for (int i = lane; i < (int)(sizeof(VieoSbpFrame) / 4); i += 32)
    reinterpret_cast<uint32_t*>(&F)[i] = reinterpret_cast<const uint32_t*>(frames + f)[i];
  __syncwarp();
  const int N = F.n_kp, nq = F.n_q;
  if (N > kMaxKp || N < 0 || nq < 0) {
    if (lane == 0) n_matches[f] = -1;
    return;
  }
  const VieoKeyPoint* kps = kps_all + F.kp_begin;
  const float* uright = ur_all + F.kp_begin;
  const uint8_t* desc = desc_all + 32 * (size_t)F.kp_begin;
  int32_t* kpm = kp_match + F.kp_begin;
  for (int i = lane; i < N; i += 32) {
    kpm[i] = -1;
    S.oct[i] = (uint8_t)kps[i].octave;
  }
  for (int i = lane; i < kMaxKp / 32; i += 32) {
    uint32_t m = 0;
    if (kp_blocked) {
      const int k = 32 * i;
      for (int b = 0; b < 32 && k + b < N; ++b)
        if (kp_blocked[F.kp_begin + k + b]) m |= 1u << b;
    }
    S.blocked[i] = m;
  }
  __syncwarp();
  bool fwd, bwd;
  motion_flags(mode, F, fwd, bwd);
  double twc[3] = {0, 0, 0};
  const VieoSbpReloc* R = mode == VIEO_SBP_RELOC ? Q.reloc + f : nullptr;
  if (mode == VIEO_SBP_RELOC) camera_centre(F, twc);
  const int th_accept = mode == VIEO_SBP_RELOC ? R->orb_dist : TH_HIGH;  // bestDist <= ORBdist (:1551) / TH_HIGH
  const bool rot_check = (mode == VIEO_SBP_LAST_FRAME || mode == VIEO_SBP_RELOC) && F.check_orientation;
  bool have_grid = false;
  const uint32_t* flist = lists + (size_t)F.q_begin * kListCap;
  int32_t* fcnt = counts + F.q_begin;
  int nmatches = 0;
  const float factor = 1.0f / HISTO_LENGTH;
  // The pass is sequential by definition (a claimed keypoint changes the next query's arg-min), so its cost is the latency
  // of one iteration (~150 cycles of warp reductions and shared-memory reads) — as long as no iteration waits for HBM / L2
  // (~700 cycles): the candidate counts and flags of 32 queries are loaded at once (lane l keeps query chunk + l, broadcast
  // by shuffle), the two list entries per lane are loaded kAhead iterations ahead into a small register ring (the list
  // slots exist for every query, entries past the count are simply not used).
  constexpr int kAhead = 4;
  uint32_t ea_r[kAhead], eb_r[kAhead];
#pragma unroll
  for (int k = 0; k < kAhead; ++k) {
    const bool in = k < nq;
    ea_r[k] = in ? flist[(size_t)k * kListCap + lane] : 0u;
    eb_r[k] = in ? flist[(size_t)k * kListCap + lane + 32] : 0u;
  }
  int n_chunk = 0, fl_chunk = 0, my_take = -1, my_dist = 256;
  // kAhead iterations per trip with the ring slot a COMPILE-TIME index: a slot is consumed and refilled in place, never
  // moved, so no iteration touches a register whose load is still in flight (rotating the ring would wait for the newest
  // load every iteration: the pass then ran at one L2 round trip per query)
  for (int qb = 0; qb < nq; qb += kAhead) {
#pragma unroll
  for (int kk = 0; kk < kAhead; ++kk) {
    const int qi = qb + kk;
    if (qi >= nq) break;
    const int q = F.q_begin + qi;
    if ((qi & 31) == 0) {  // fcnt[qi' >= qi] still holds the candidate counts: the pass overwrites entries behind it only
      const int qq = qi + lane;
      n_chunk = qq < nq ? fcnt[qq] : 0;
      fl_chunk = (qq < nq && Q.flags) ? Q.flags[F.q_begin + qq] : 0;
    }
    const int n = __shfl_sync(0xffffffffu, n_chunk, qi & 31);
    const uint8_t qflag = (uint8_t)__shfl_sync(0xffffffffu, fl_chunk, qi & 31);
    const uint32_t ea = ea_r[kk], eb = eb_r[kk];
    if (qi + kAhead < nq) {
      const uint32_t* Ln = flist + (size_t)(qi + kAhead) * kListCap;
      ea_r[kk] = Ln[lane];
      eb_r[kk] = Ln[lane + 32];
    }
    uint32_t best = 0xffffffffu, second = 0xffffffffu;  // lane-local two smallest keys (dist << 16 | pos)
    int idx_a = -1, idx_b = -1;                          // keypoints of the lane's two entries
    if (n > 0 && n <= kListCap) {
      if (lane < n) {
        const uint32_t e = ea;
        idx_a = (int)(e & 0xffffu);
        if (!((S.blocked[idx_a >> 5] >> (idx_a & 31)) & 1u)) best = (e & 0xffff0000u) | (uint32_t)lane;
      }
      if (lane + 32 < n) {
        const uint32_t e = eb;
        idx_b = (int)(e & 0xffffu);
        if (!((S.blocked[idx_b >> 5] >> (idx_b & 31)) & 1u)) second = (e & 0xffff0000u) | (uint32_t)(lane + 32);
      }
      if (second < best) {
        const uint32_t t = best; best = second; second = t;
        const int ti = idx_a; idx_a = idx_b; idx_b = ti;
      }
    } else if (n > kListCap) {
      // overflow: re-enumerate with the claim filter (same order, same keys)
      if (!have_grid) {  // warp-uniform; the CTA is this one warp, so build_grid's barriers are warp barriers
        for (int i = lane; i < kCells; i += 32) S.cnt[i] = 0;
        __syncthreads();
        build_grid(S, GridGeom{F.minx, F.miny, F.grid_winv, F.grid_hinv}, kps, N, lane, 32);
        have_grid = true;
      }
      Cand c;
      make_window(mode, F, Q, q, fwd, bwd, c, R, twc);
      const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(Q.desc + 32 * (size_t)q));
      const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(Q.desc + 32 * (size_t)q) + 1);
      enumerate(F, S, kps, uright, c, lane, [&](int pos, int j) {
        if ((S.blocked[j >> 5] >> (j & 31)) & 1u) return;
        const uint32_t key = ((uint32_t)hamming256(d0, d1, desc + 32 * (size_t)j) << 16) | (uint32_t)pos;
        if (key < best) {
          second = best; idx_b = idx_a;
          best = key; idx_a = j;
        } else if (key < second) {
          second = key; idx_b = j;
        }
      });
    }
    int take = -1, take_dist = 256;
    if (n > 0) {
      const uint32_t m1 = __reduce_min_sync(0xffffffffu, best);
      if (m1 != 0xffffffffu) {
        const int bestDist = (int)(m1 >> 16);
        const int bestIdx = __reduce_max_sync(0xffffffffu, best == m1 ? idx_a : -1);
        if (bestDist <= th_accept) {
          bool ok = true;
          if (mode == VIEO_SBP_LOCAL_MAP) {
            const uint32_t c2 = best == m1 ? second : best;
            const uint32_t m2 = __reduce_min_sync(0xffffffffu, c2);
            if (m2 != 0xffffffffu) {
              const int i2 = __reduce_max_sync(0xffffffffu, c2 == m2 ? (best == m1 ? idx_b : idx_a) : -1);
              const int bestDist2 = (int)(m2 >> 16);
              if (S.oct[bestIdx] == S.oct[i2] && (float)bestDist > __fmul_rn(F.nn_ratio, (float)bestDist2)) ok = false;
            }
          }
          if (ok) {
            take = bestIdx;
            take_dist = bestDist;
          }
        }
      }
    }
    // No global store inside the pass: __syncwarp orders memory for the warp and would wait for the store's round trip to
    // L2 every iteration (measured: ~1100 cycles per query with the stores, the pass was 0.56 / 1.42 ms for 700 / 2500
    // queries).  Lane (qi & 31) keeps query qi's result; every 32 queries the lanes write theirs with coalesced stores.
    if (lane == (qi & 31)) {
      my_take = take;
      my_dist = take_dist;
    }
    // RELOC: AddMapPoint makes the keypoint's slot non-null for every later query (:1541-1543, 1553)
    if (lane == 0 && take >= 0 && ((qflag & 1) || mode == VIEO_SBP_RELOC)) S.blocked[take >> 5] |= 1u << (take & 31);
    nmatches += take >= 0;
    __syncwarp();
    if ((qi & 31) == 31 || qi == nq - 1) {
      const int qq = (qi & ~31) + lane;
      if (qq <= qi) {
        q_match[F.q_begin + qq] = my_take;
        q_dist[F.q_begin + qq] = my_dist;
        fcnt[qq] = my_take;  // reused: the keypoint an accepted query took (-1 none); the rotation bins follow after the pass
        // AddMapPoint: the LAST query that took a keypoint owns it (an earlier taker with Observations() == 0 did not
        // block it); queries are flushed out of order across lanes, the maximum is the last one
        if (my_take >= 0) atomicMax(&kpm[my_take], qq);
      }
    }
  }
  }
  __syncwarp();
  // rotation bins of the accepted queries, lanes over queries (kept out of the sequential pass: the keypoint's angle is a
  // dependent global load)
  if (rot_check) {
    for (int qi = lane; qi < nq; qi += 32) {
      const int take = fcnt[qi];
      int bin = -1;
      if (take >= 0) {
        float rot = __fsub_rn(Q.angle[F.q_begin + qi], kps[take].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == HISTO_LENGTH) bin = 0;
      }
      fcnt[qi] = bin;
    }
    __syncwarp();
  }
  // ---- rotation consistency (:1445-1464) ---------------------------------------------------------------------------------
  if (rot_check) {
    int h = 0;  // lane b < 30 counts bin b
    for (int qi = 0; qi < nq; ++qi) h += (fcnt[qi] == lane) ? 1 : 0;
    if (lane < HISTO_LENGTH) S.hist[lane] = h;
    __syncwarp();
    int ind1 = -1, ind2 = -1, ind3 = -1;
    {
      int max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = S.hist[i];
        if (s > max1) {
          max3 = max2; max2 = max1; max1 = s;
          ind3 = ind2; ind2 = ind1; ind1 = i;
        } else if (s > max2) {
          max3 = max2; max2 = s;
          ind3 = ind2; ind2 = i;
        } else if (s > max3) {
          max3 = s; ind3 = i;
        }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) {
        ind2 = -1; ind3 = -1;
      } else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) {
        ind3 = -1;
      }
    }
    int erased = 0;
    for (int qi = lane; qi < nq; qi += 32) {
      const int bin = fcnt[qi];
      if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) {
        kpm[q_match[F.q_begin + qi]] = -1;  // EraseMapPointMatch
        ++erased;
      }
    }
    erased = __reduce_add_sync(0xffffffffu, erased);
    nmatches -= erased;
  }
  if (lane == 0) n_matches[f] = nmatches;
}


// ------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227), search half: the projection of every map point into a
// keyframe, IsInImage, scale-invariance range, viewing cone, PredictScale (threshold table, see frustum.cu),
// GetFeaturesInArea(u, v, th_radius * scale[level]) with the level band [level - 1, level], the chi-square gate against the
// keypoint (with pbf) and the strict-'<' Hamming arg-min.  Used by Fuse(KF, vpMapPoints, th) (LocalMapping::
// SearchInNeighbors), Fuse(KF, Scw, ...) and the Sim3 / keyframe SearchByProjection variants.  The map-point link updates
// that follow (FuseMP / AddObservation / vpReplacePoint) never change keypoints or descriptors, so every point is
// independent: one CTA per keyframe builds the grid, one warp per map point searches; no claim pass.
struct Chi2Gate {
  const float* __restrict__ uright;
  const float* inv_level_sigma2;
  float u, v, ur;
  int use_bf;
  __device__ __forceinline__ bool operator()(int j, const VieoKeyPoint& kp) const {
    if (!use_bf) return true;
    const float kpr = uright[j];
    const float ex = __fsub_rn(u, kp.x), ey = __fsub_rn(v, kp.y);
    const float exy = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    if (kpr >= 0) {
      const float er = __fsub_rn(ur, kpr);
      const float e2 = __fadd_rn(exy, __fmul_rn(er, er));
      return !((double)__fmul_rn(e2, inv_level_sigma2[kp.octave]) > 7.8);   // chi2(0.05, 3)
    }
    return !((double)__fmul_rn(exy, inv_level_sigma2[kp.octave]) > 5.99);   // chi2(0.05, 2)
  }
};
__device__ __forceinline__ float sum3f(float a0, float a1, float a2) { return __fadd_rn(a0, __fadd_rn(a1, a2)); }

__global__ void __launch_bounds__(kSbpWarps * 32) k_proj_search(const VieoProjSearchFrame* __restrict__ frames,
                                                                const VieoKeyPoint* __restrict__ kps_all,
                                                                const float* __restrict__ ur_all,
                                                                const uint8_t* __restrict__ desc_all,
                                                                const float* __restrict__ wP, const float* __restrict__ Pn,
                                                                const float* __restrict__ max_dist,
                                                                const float* __restrict__ min_dist,
                                                                const uint8_t* __restrict__ q_desc,
                                                                const uint8_t* __restrict__ q_skip,
                                                                int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist,
                                                                int32_t* __restrict__ level_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SbpShared& S = *reinterpret_cast<SbpShared*>(smem_raw);
  __shared__ VieoProjSearchFrame F;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x;
  for (int i = tid; i < (int)(sizeof(VieoProjSearchFrame) / 4); i += T)
    reinterpret_cast<uint32_t*>(&F)[i] = reinterpret_cast<const uint32_t*>(frames + f)[i];
  for (int i = tid; i < kCells; i += T) S.cnt[i] = 0;
  __syncthreads();
  const int N = F.n_kp, nq = F.n_q;
  if (N > kMaxKp || N < 0 || nq < 0 || F.n_levels < 1 || F.n_levels > 16) {
    for (int qi = tid; qi < nq; qi += T) {  // reported per query: level -2
      best_idx[F.q_begin + qi] = -1;
      best_dist[F.q_begin + qi] = 256;
      level_out[F.q_begin + qi] = -2;
    }
    return;
  }
  const VieoKeyPoint* kps = kps_all + F.kp_begin;
  const float* uright = ur_all + F.kp_begin;
  const uint8_t* desc = desc_all + 32 * (size_t)F.kp_begin;
  const GridGeom G{F.minx, F.miny, F.grid_winv, F.grid_hinv};
  build_grid(S, G, kps, N, tid, T);
  for (int qi = warp; qi < nq; qi += kSbpWarps) {
    const size_t q = (size_t)F.q_begin + qi;
    int lvl = -1, bidx = -1, bdist = 256;
    if (!(q_skip && q_skip[q])) {
      const float X = wP[3 * q], Y = wP[3 * q + 1], Z = wP[3 * q + 2];
      float Pc[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        Pc[r] = __fadd_rn(sum3f(__fmul_rn(F.Rcw[3 * r], X), __fmul_rn(F.Rcw[3 * r + 1], Y), __fmul_rn(F.Rcw[3 * r + 2], Z)),
                          F.tcw[r]);
      if (!(Pc[2] <= 0.0f)) {
        const float invz = __fdiv_rn(1.0f, Pc[2]);
        const float xn = __fmul_rn(Pc[0], invz), yn = __fmul_rn(Pc[1], invz);
        const float u = sum3f(__fmul_rn(F.fx, xn), __fmul_rn(0.0f, yn), __fmul_rn(F.cx, 1.0f));
        const float v = sum3f(__fmul_rn(0.0f, xn), __fmul_rn(F.fy, yn), __fmul_rn(F.cy, 1.0f));
        if (u >= F.minx && u < F.maxx && v >= F.miny && v < F.maxy) {
          const float ox = __fsub_rn(X, F.Ow[0]), oy = __fsub_rn(Y, F.Ow[1]), oz = __fsub_rn(Z, F.Ow[2]);
          const float d3 = __fsqrt_rn(sum3f(__fmul_rn(ox, ox), __fmul_rn(oy, oy), __fmul_rn(oz, oz)));
          const float mx = max_dist[q];
          bool ok = !(d3 < __fmul_rn(0.8f, min_dist[q]) || d3 > __fmul_rn(1.2f, mx));
          if (ok && F.check_viewing_angle) {
            const float dot = sum3f(__fmul_rn(ox, Pn[3 * q]), __fmul_rn(oy, Pn[3 * q + 1]), __fmul_rn(oz, Pn[3 * q + 2]));
            ok = !((double)dot < 0.5 * (double)d3);
          }
          if (ok) {
            const float ratio = __fdiv_rn(mx, d3);
            lvl = 0;
            for (int k = 1; k < F.n_levels; ++k) lvl += ratio >= F.level_ratio[k] ? 1 : 0;
            Cand c;
            c.x = u; c.y = v;
            c.r = __fmul_rn(F.th_radius, F.scale[lvl]);
            c.minlevel = lvl - 1; c.maxlevel = lvl;
            c.ur = 0;
            const Chi2Gate gate{uright, F.inv_level_sigma2, u, v, __fsub_rn(u, __fmul_rn(F.bf, invz)), F.use_bf};
            const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(q_desc + 32 * q));
            const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(q_desc + 32 * q) + 1);
            uint32_t bkey = 0xffffffffu;
            int bj = -1;
            enumerate_gated(G, S, kps, c, lane, gate, [&](int pos, int j) {
              const uint32_t key = ((uint32_t)hamming256(d0, d1, desc + 32 * (size_t)j) << 16) | (uint32_t)pos;
              if (key < bkey) {
                bkey = key;
                bj = j;
              }
            });
            // strict '<' in candidate order == lexicographic minimum of (distance, position)
            const uint32_t kmin = __reduce_min_sync(0xffffffffu, bkey);
            if (kmin != 0xffffffffu) {
              const int src = __ffs(__ballot_sync(0xffffffffu, bkey == kmin)) - 1;
              bidx = __shfl_sync(0xffffffffu, bj, src);
              bdist = (int)(kmin >> 16);
            }
          }
        }
      }
    }
    if (lane == 0) {
      best_idx[q] = bidx;
      best_dist[q] = bdist;
      level_out[q] = lvl;
    }
  }
}

}  // namespace

extern "C" {

size_t vieo_sbp_scratch_bytes(int n_queries_total) {
  return (size_t)std::max(n_queries_total, 1) * (kListCap + 1) * 4;
}

static int sbp_launch(int mode, const VieoSbpFrame* frames_dev, int n_frames, const VieoKeyPoint* kps_dev,
                      const float* uright_dev, const uint8_t* desc_dev, const QueryIn& Q, const uint8_t* kp_blocked_dev,
                      int32_t* kp_match_dev, int32_t* q_match_dev, int32_t* q_dist_dev, int32_t* n_matches_dev,
                      void* scratch_dev, size_t scratch_bytes, void* stream) {
  VIEO_ARG(((uintptr_t)desc_dev | (uintptr_t)Q.desc) % 16 == 0, "descriptors must be 16-byte aligned");
  VIEO_ARG(scratch_bytes >= (size_t)(kListCap + 1) * 4, "scratch too small");
  static SmemOptIn opt_in_lists, opt_in_claim;
  VIEO_CK(smem_opt_in(k_sbp_lists, sizeof(SbpShared), opt_in_lists));
  VIEO_CK(smem_opt_in(k_sbp_claim, sizeof(SbpShared), opt_in_claim));
  // scratch = [counts (n_q_total) | lists (n_q_total x kListCap)]; the caller sized it with vieo_sbp_scratch_bytes
  const size_t nq_total = scratch_bytes / ((size_t)(kListCap + 1) * 4);
  int32_t* counts = (int32_t*)scratch_dev;
  uint32_t* lists = (uint32_t*)scratch_dev + nq_total;
  for (int f0 = 0; f0 < n_frames; f0 += 65535) {  // grid.y bound
    const int nf = std::min(n_frames - f0, 65535);
    QueryIn Qc = Q;
    if (Qc.reloc) Qc.reloc += f0;
    k_sbp_lists<<<dim3(kListCtas, nf), kListWarps * 32, sizeof(SbpShared), (cudaStream_t)stream>>>(
        mode, frames_dev + f0, kps_dev, uright_dev, desc_dev, Qc, lists, counts);
  }
  k_sbp_claim<<<n_frames, 32, sizeof(SbpShared), (cudaStream_t)stream>>>(mode, frames_dev, kps_dev, uright_dev, desc_dev, Q,
                                                                        kp_blocked_dev, kp_match_dev, q_match_dev, q_dist_dev,
                                                                        n_matches_dev, lists, counts);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_sbp_batch_dev(int mode, const VieoSbpFrame* frames_dev, int n_frames, const VieoKeyPoint* kps_dev,
                       const float* uright_dev, const uint8_t* desc_dev, const VieoSbpQueries* q, const uint8_t* kp_blocked_dev,
                       int32_t* kp_match_dev, int32_t* q_match_dev, int32_t* q_dist_dev, int32_t* n_matches_dev,
                       void* scratch_dev, size_t scratch_bytes, void* stream) {
  VIEO_ARG(mode == VIEO_SBP_LAST_FRAME || mode == VIEO_SBP_LOCAL_MAP, "bad mode");
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames_dev && kps_dev && uright_dev && desc_dev && q && kp_match_dev && q_match_dev && q_dist_dev && n_matches_dev &&
               scratch_dev, "null argument");
  VIEO_ARG(q->desc && q->flags && q->level, "null query array");
  if (mode == VIEO_SBP_LAST_FRAME) VIEO_ARG(q->Xw && q->angle, "null query array");
  else VIEO_ARG(q->proj && q->viewcos && q->depth, "null query array");
  QueryIn Q{q->Xw, q->level, q->angle, q->proj, q->viewcos, q->depth, q->desc, q->flags, nullptr, nullptr, nullptr, nullptr};
  return sbp_launch(mode, frames_dev, n_frames, kps_dev, uright_dev, desc_dev, Q, kp_blocked_dev, kp_match_dev, q_match_dev,
                    q_dist_dev, n_matches_dev, scratch_dev, scratch_bytes, stream);
}

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist, th_far_pts) (src/ORBmatcher.cc:1471-1606)
int vieo_sbp_reloc_batch_dev(const VieoSbpFrame* frames_dev, const VieoSbpReloc* reloc_dev, int n_frames,
                             const VieoKeyPoint* kps_dev, const uint8_t* desc_dev, const double* q_Xw_dev,
                             const float* q_angle_dev, const float* q_max_dist_dev, const float* q_min_dist_dev,
                             const uint8_t* q_desc_dev, const uint8_t* kp_blocked_dev, int32_t* kp_match_dev,
                             int32_t* q_match_dev, int32_t* q_dist_dev, int32_t* q_level_dev, int32_t* n_matches_dev,
                             void* scratch_dev, size_t scratch_bytes, void* stream) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames_dev && reloc_dev && kps_dev && desc_dev && q_Xw_dev && q_angle_dev && q_max_dist_dev && q_min_dist_dev &&
               q_desc_dev && kp_match_dev && q_match_dev && q_dist_dev && n_matches_dev && scratch_dev, "null argument");
  QueryIn Q{q_Xw_dev, nullptr, q_angle_dev, nullptr, nullptr, nullptr, q_desc_dev, nullptr, q_max_dist_dev, q_min_dist_dev,
            reloc_dev, q_level_dev};
  return sbp_launch(VIEO_SBP_RELOC, frames_dev, n_frames, kps_dev, nullptr, desc_dev, Q, kp_blocked_dev, kp_match_dev,
                    q_match_dev, q_dist_dev, n_matches_dev, scratch_dev, scratch_bytes, stream);
}

int vieo_sbp_reloc_batch(const VieoSbpFrame* frames, const VieoSbpReloc* reloc, int n_frames, const VieoKeyPoint* kps,
                         const uint8_t* desc, const double* q_Xw, const float* q_angle, const float* q_max_dist,
                         const float* q_min_dist, const uint8_t* q_desc, const uint8_t* kp_blocked, int32_t* kp_match,
                         int32_t* q_match, int32_t* q_dist, int32_t* q_level, int32_t* n_matches, int device) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames && reloc && n_matches, "null argument");
  size_t nk = 0, nq = 0;
  for (int f = 0; f < n_frames; ++f) {
    VIEO_ARG(frames[f].n_kp >= 0 && frames[f].n_q >= 0 && frames[f].kp_begin >= 0 && frames[f].q_begin >= 0, "bad frame range");
    if (frames[f].n_kp > kMaxKp) {
      set_error("vieo_sbp_reloc_batch: frame %d has %d keypoints (max %d)", f, frames[f].n_kp, kMaxKp);
      return VIEO_E_CAPACITY;
    }
    VIEO_ARG(frames[f].n_levels >= 1 && frames[f].n_levels <= 16, "bad pyramid depth");
    VIEO_ARG(reloc[f].orb_dist >= 0 && reloc[f].orb_dist < 256, "ORBdist must be below 256");
    nk = std::max(nk, (size_t)frames[f].kp_begin + frames[f].n_kp);
    nq = std::max(nq, (size_t)frames[f].q_begin + frames[f].n_q);
  }
  VIEO_ARG(nk == 0 || (kps && desc && kp_match), "null keypoint array");
  VIEO_ARG(nq == 0 || (q_Xw && q_angle && q_max_dist && q_min_dist && q_desc && q_match && q_dist), "null query array");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  VIEO_ARG(cs, "no call scratch");
  std::vector<VieoSbpReloc> rl(reloc, reloc + n_frames);
  for (int f = 0; f < n_frames; ++f)
    if ((rc = vieo_frustum_level_table(rl[f].log_scale_factor, frames[f].n_levels, rl[f].level_ratio))) return rc;
  const size_t nk1 = std::max<size_t>(nk, 1), nq1 = std::max<size_t>(nq, 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  const size_t o_fr = take(sizeof(VieoSbpFrame) * n_frames), o_rl = take(sizeof(VieoSbpReloc) * n_frames),
               o_kp = take(sizeof(VieoKeyPoint) * nk1), o_de = take(32 * nk1), o_bl = take(nk1), o_xw = take(24 * nq1),
               o_an = take(4 * nq1), o_mx = take(4 * nq1), o_mn = take(4 * nq1), o_qd = take(32 * nq1);
  const size_t in_bytes = off;
  const size_t o_km = take(4 * nk1), o_qm = take(4 * nq1), o_qs = take(4 * nq1), o_ql = take(4 * nq1), o_nm = take(4 * (size_t)n_frames);
  const size_t io_bytes = off;
  const size_t sc_bytes = vieo_sbp_scratch_bytes((int)nq1);
  uint8_t* dbuf = (uint8_t*)cs->get(0, io_bytes);
  void* dsc = cs->get(1, sc_bytes);
  uint8_t* hbuf = (uint8_t*)cs->get_pinned(io_bytes);
  VIEO_ARG(dbuf && dsc && hbuf, "staging allocation failed");
  auto put = [&](size_t o, const void* src, size_t bytes) { if (src && bytes) memcpy(hbuf + o, src, bytes); };
  put(o_fr, frames, sizeof(VieoSbpFrame) * n_frames); put(o_rl, rl.data(), sizeof(VieoSbpReloc) * n_frames);
  put(o_kp, kps, sizeof(VieoKeyPoint) * nk); put(o_de, desc, 32 * nk); put(o_bl, kp_blocked, nk);
  put(o_xw, q_Xw, 24 * nq); put(o_an, q_angle, 4 * nq); put(o_mx, q_max_dist, 4 * nq); put(o_mn, q_min_dist, 4 * nq);
  put(o_qd, q_desc, 32 * nq);
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf, in_bytes, cudaMemcpyHostToDevice, cs->st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_km, 0xff, o_nm - o_km, cs->st);  // outside every frame's range: -1
  if (e == cudaSuccess) {
    rc = vieo_sbp_reloc_batch_dev((const VieoSbpFrame*)(dbuf + o_fr), (const VieoSbpReloc*)(dbuf + o_rl), n_frames,
                                  (const VieoKeyPoint*)(dbuf + o_kp), dbuf + o_de, (const double*)(dbuf + o_xw),
                                  (const float*)(dbuf + o_an), (const float*)(dbuf + o_mx), (const float*)(dbuf + o_mn), dbuf + o_qd,
                                  kp_blocked ? dbuf + o_bl : nullptr, (int32_t*)(dbuf + o_km), (int32_t*)(dbuf + o_qm),
                                  (int32_t*)(dbuf + o_qs), (int32_t*)(dbuf + o_ql), (int32_t*)(dbuf + o_nm), dsc, sc_bytes, cs->st);
    if (rc == VIEO_OK) {
      e = cudaMemcpyAsync(hbuf + o_km, dbuf + o_km, io_bytes - o_km, cudaMemcpyDeviceToHost, cs->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(cs->st);
    }
  }
  if (e == cudaSuccess && rc == VIEO_OK) {
    if (nk) memcpy(kp_match, hbuf + o_km, 4 * nk);
    if (nq) {
      memcpy(q_match, hbuf + o_qm, 4 * nq);
      memcpy(q_dist, hbuf + o_qs, 4 * nq);
      if (q_level) memcpy(q_level, hbuf + o_ql, 4 * nq);
    }
    memcpy(n_matches, hbuf + o_nm, 4 * (size_t)n_frames);
  }
  if (e != cudaSuccess) {
    set_error("vieo_sbp_reloc_batch: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  return rc;
}

int vieo_sbp_batch(int mode, const VieoSbpFrame* frames, int n_frames, const VieoKeyPoint* kps, const float* uright,
                   const uint8_t* desc, const VieoSbpQueries* q, const uint8_t* kp_blocked, int32_t* kp_match, int32_t* q_match,
                   int32_t* q_dist, int32_t* n_matches, int device) {
  VIEO_ARG(mode == VIEO_SBP_LAST_FRAME || mode == VIEO_SBP_LOCAL_MAP, "bad mode");
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames && q && n_matches, "null argument");
  size_t nk = 0, nq = 0;
  for (int f = 0; f < n_frames; ++f) {
    VIEO_ARG(frames[f].n_kp >= 0 && frames[f].n_q >= 0 && frames[f].kp_begin >= 0 && frames[f].q_begin >= 0, "bad frame range");
    if (frames[f].n_kp > kMaxKp) {
      set_error("vieo_sbp_batch: frame %d has %d keypoints (max %d)", f, frames[f].n_kp, kMaxKp);
      return VIEO_E_CAPACITY;
    }
    VIEO_ARG(frames[f].n_levels >= 1 && frames[f].n_levels <= 16, "bad pyramid depth");
    nk = std::max(nk, (size_t)frames[f].kp_begin + frames[f].n_kp);
    nq = std::max(nq, (size_t)frames[f].q_begin + frames[f].n_q);
  }
  VIEO_ARG(nk == 0 || (kps && uright && desc && kp_match), "null keypoint array");
  VIEO_ARG(nq == 0 || (q->desc && q->flags && q->level && q_match && q_dist), "null query array");
  if (mode == VIEO_SBP_LAST_FRAME) VIEO_ARG(nq == 0 || (q->Xw && q->angle), "null query array");
  else VIEO_ARG(nq == 0 || (q->proj && q->viewcos && q->depth), "null query array");
  for (int f = 0; f < n_frames; ++f)  // against the frame's OWN pyramid depth: a level past it would index an unset scale[]
    for (int i = frames[f].q_begin; i < frames[f].q_begin + frames[f].n_q; ++i)
      VIEO_ARG(q->level[i] >= (mode == VIEO_SBP_LOCAL_MAP ? -1 : 0) && q->level[i] < frames[f].n_levels, "query level out of range");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  VIEO_ARG(cs, "no call scratch");
  const size_t nk1 = std::max<size_t>(nk, 1), nq1 = std::max<size_t>(nq, 1);
  // one staging buffer: every array 16-byte aligned
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  const size_t o_fr = take(sizeof(VieoSbpFrame) * n_frames), o_kp = take(sizeof(VieoKeyPoint) * nk1), o_ur = take(4 * nk1),
               o_de = take(32 * nk1), o_bl = take(nk1), o_xw = take(24 * nq1), o_lv = take(4 * nq1), o_an = take(4 * nq1),
               o_pr = take(12 * nq1), o_vc = take(4 * nq1), o_dp = take(4 * nq1), o_qd = take(32 * nq1), o_fl = take(nq1);
  const size_t in_bytes = off;
  const size_t o_km = take(4 * nk1), o_qm = take(4 * nq1), o_qs = take(4 * nq1), o_nm = take(4 * (size_t)n_frames);
  const size_t io_bytes = off;
  const size_t sc_bytes = vieo_sbp_scratch_bytes((int)nq1);
  uint8_t* dbuf = (uint8_t*)cs->get(0, io_bytes);
  void* dsc = cs->get(1, sc_bytes);
  uint8_t* hbuf = (uint8_t*)cs->get_pinned(io_bytes);
  VIEO_ARG(dbuf && dsc && hbuf, "staging allocation failed");
  auto put = [&](size_t o, const void* src, size_t bytes) { if (src && bytes) memcpy(hbuf + o, src, bytes); };
  put(o_fr, frames, sizeof(VieoSbpFrame) * n_frames);
  put(o_kp, kps, sizeof(VieoKeyPoint) * nk); put(o_ur, uright, 4 * nk); put(o_de, desc, 32 * nk);
  put(o_bl, kp_blocked, nk);
  put(o_lv, q->level, 4 * nq); put(o_qd, q->desc, 32 * nq); put(o_fl, q->flags, nq);
  if (mode == VIEO_SBP_LAST_FRAME) { put(o_xw, q->Xw, 24 * nq); put(o_an, q->angle, 4 * nq); }
  else { put(o_pr, q->proj, 12 * nq); put(o_vc, q->viewcos, 4 * nq); put(o_dp, q->depth, 4 * nq); }
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf, in_bytes, cudaMemcpyHostToDevice, cs->st);
  // keypoints / queries outside every frame's range read back as -1
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_km, 0xff, o_nm - o_km, cs->st);
  if (e == cudaSuccess) {
    VieoSbpQueries dq{};
    dq.Xw = (const double*)(dbuf + o_xw); dq.level = (const int32_t*)(dbuf + o_lv); dq.angle = (const float*)(dbuf + o_an);
    dq.proj = (const float*)(dbuf + o_pr); dq.viewcos = (const float*)(dbuf + o_vc); dq.depth = (const float*)(dbuf + o_dp);
    dq.desc = dbuf + o_qd; dq.flags = dbuf + o_fl;
    rc = vieo_sbp_batch_dev(mode, (const VieoSbpFrame*)(dbuf + o_fr), n_frames, (const VieoKeyPoint*)(dbuf + o_kp),
                            (const float*)(dbuf + o_ur), dbuf + o_de, &dq, kp_blocked ? dbuf + o_bl : nullptr,
                            (int32_t*)(dbuf + o_km), (int32_t*)(dbuf + o_qm), (int32_t*)(dbuf + o_qs), (int32_t*)(dbuf + o_nm),
                            dsc, sc_bytes, cs->st);
    if (rc == VIEO_OK) {
      e = cudaMemcpyAsync(hbuf + o_km, dbuf + o_km, io_bytes - o_km, cudaMemcpyDeviceToHost, cs->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(cs->st);
    }
  }
  if (e == cudaSuccess && rc == VIEO_OK) {
    if (nk) memcpy(kp_match, hbuf + o_km, 4 * nk);
    if (nq) { memcpy(q_match, hbuf + o_qm, 4 * nq); memcpy(q_dist, hbuf + o_qs, 4 * nq); }
    memcpy(n_matches, hbuf + o_nm, 4 * (size_t)n_frames);
  }
  if (e != cudaSuccess) {
    set_error("vieo_sbp_batch: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  return rc;
}

int vieo_proj_search_batch_dev(const VieoProjSearchFrame* frames_dev, int n_frames, const VieoKeyPoint* kps_dev,
                               const float* uright_dev, const uint8_t* desc_dev, const float* wP_dev, const float* normal_dev,
                               const float* max_dist_dev, const float* min_dist_dev, const uint8_t* q_desc_dev,
                               const uint8_t* q_skip_dev, int32_t* best_idx_dev, int32_t* best_dist_dev, int32_t* level_dev,
                               void* stream) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames_dev && kps_dev && uright_dev && desc_dev && wP_dev && normal_dev && max_dist_dev && min_dist_dev &&
               q_desc_dev && best_idx_dev && best_dist_dev && level_dev, "null argument");
  VIEO_ARG(((uintptr_t)desc_dev | (uintptr_t)q_desc_dev) % 16 == 0, "descriptors must be 16-byte aligned");
  static SmemOptIn opt_in;
  VIEO_CK(smem_opt_in(k_proj_search, sizeof(SbpShared), opt_in));
  k_proj_search<<<n_frames, kSbpWarps * 32, sizeof(SbpShared), (cudaStream_t)stream>>>(
      frames_dev, kps_dev, uright_dev, desc_dev, wP_dev, normal_dev, max_dist_dev, min_dist_dev, q_desc_dev, q_skip_dev,
      best_idx_dev, best_dist_dev, level_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_proj_search_batch(const VieoProjSearchFrame* frames, int n_frames, const VieoKeyPoint* kps, const float* uright,
                           const uint8_t* desc, const float* wP, const float* normal, const float* max_dist,
                           const float* min_dist, const uint8_t* q_desc, const uint8_t* q_skip, int32_t* best_idx,
                           int32_t* best_dist, int32_t* level, int device) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames, "null argument");
  size_t nk = 0, nq = 0;
  std::vector<VieoProjSearchFrame> fr(frames, frames + n_frames);
  for (int f = 0; f < n_frames; ++f) {
    VIEO_ARG(fr[f].n_kp >= 0 && fr[f].n_q >= 0 && fr[f].kp_begin >= 0 && fr[f].q_begin >= 0, "bad frame range");
    if (fr[f].n_kp > kMaxKp) {
      set_error("vieo_proj_search_batch: frame %d has %d keypoints (max %d)", f, fr[f].n_kp, kMaxKp);
      return VIEO_E_CAPACITY;
    }
    int rc = vieo_frustum_level_table(fr[f].log_scale_factor, fr[f].n_levels, fr[f].level_ratio);
    if (rc) return rc;
    nk = std::max(nk, (size_t)fr[f].kp_begin + fr[f].n_kp);
    nq = std::max(nq, (size_t)fr[f].q_begin + fr[f].n_q);
  }
  VIEO_ARG(nk == 0 || (kps && uright && desc), "null keypoint array");
  VIEO_ARG(nq == 0 || (wP && normal && max_dist && min_dist && q_desc && best_idx && best_dist && level), "null query array");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  VIEO_ARG(cs, "no call scratch");
  const size_t nk1 = std::max<size_t>(nk, 1), nq1 = std::max<size_t>(nq, 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  const size_t o_fr = take(sizeof(VieoProjSearchFrame) * n_frames), o_kp = take(sizeof(VieoKeyPoint) * nk1), o_ur = take(4 * nk1),
               o_de = take(32 * nk1), o_wp = take(12 * nq1), o_pn = take(12 * nq1), o_mx = take(4 * nq1), o_mn = take(4 * nq1),
               o_qd = take(32 * nq1), o_sk = take(nq1);
  const size_t in_bytes = off;
  const size_t o_bi = take(4 * nq1), o_bd = take(4 * nq1), o_lv = take(4 * nq1);
  const size_t io_bytes = off;
  uint8_t* dbuf = (uint8_t*)cs->get(0, io_bytes);
  uint8_t* hbuf = (uint8_t*)cs->get_pinned(io_bytes);
  VIEO_ARG(dbuf && hbuf, "staging allocation failed");
  auto put = [&](size_t o, const void* src, size_t bytes) { if (src && bytes) memcpy(hbuf + o, src, bytes); };
  put(o_fr, fr.data(), sizeof(VieoProjSearchFrame) * n_frames);
  put(o_kp, kps, sizeof(VieoKeyPoint) * nk); put(o_ur, uright, 4 * nk); put(o_de, desc, 32 * nk);
  put(o_wp, wP, 12 * nq); put(o_pn, normal, 12 * nq); put(o_mx, max_dist, 4 * nq); put(o_mn, min_dist, 4 * nq);
  put(o_qd, q_desc, 32 * nq); put(o_sk, q_skip, nq);
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf, in_bytes, cudaMemcpyHostToDevice, cs->st);
  // queries outside every frame's range read back as -1
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_bi, 0xff, io_bytes - o_bi, cs->st);
  if (e == cudaSuccess) {
    rc = vieo_proj_search_batch_dev((const VieoProjSearchFrame*)(dbuf + o_fr), n_frames, (const VieoKeyPoint*)(dbuf + o_kp),
                                    (const float*)(dbuf + o_ur), dbuf + o_de, (const float*)(dbuf + o_wp),
                                    (const float*)(dbuf + o_pn), (const float*)(dbuf + o_mx), (const float*)(dbuf + o_mn),
                                    dbuf + o_qd, q_skip ? dbuf + o_sk : nullptr, (int32_t*)(dbuf + o_bi),
                                    (int32_t*)(dbuf + o_bd), (int32_t*)(dbuf + o_lv), cs->st);
    if (rc == VIEO_OK) {
      e = cudaMemcpyAsync(hbuf + o_bi, dbuf + o_bi, io_bytes - o_bi, cudaMemcpyDeviceToHost, cs->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(cs->st);
    }
  }
  if (e != cudaSuccess) {
    set_error("vieo_proj_search_batch: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  if (rc == VIEO_OK && nq) {
    memcpy(best_idx, hbuf + o_bi, 4 * nq); memcpy(best_dist, hbuf + o_bd, 4 * nq); memcpy(level, hbuf + o_lv, 4 * nq);
  }
  return rc;
}

}  // extern "C"

// ORBmatcher::SearchForTriangulation on the device (src/ORBmatcher.cc:896-1150; LocalMapping::CreateNewMapPoints calls it
// for every neighbour keyframe right before the local bundle adjustment, src/LocalMapping.cc:709), single-pinhole form
// (usedistort_ == false, one camera per keyframe).  One CTA per keyframe pair:
//   A  the DBoW2::FeatureVector walk (:964-1135): both keyframes' node lists come flattened (ids ascending, CSR of keypoint
//      indices); thread a binary-searches node a of keyframe 1 in keyframe 2's ids; a block scan of the matched nodes' list
//      sizes numbers the queries (idx1 in the reference's visiting order);
//   B  one warp per query: lanes over the node's keyframe-2 list — existing map point / stereo rule, 256-bit Hamming
//      against TH_LOW, the epipole-distance rule for mono-mono pairs, the fundamental-matrix epipolar test of
//      GeometricCamera::epipolarConstrain in the reference's float / double mix — and a compact list of the candidates that
//      pass, keyed (distance, LAST position wins: the reference's `dist > best -> continue` lets an equal distance replace);
//   C  warp 0 replays the reference's sequential rule "a keyframe-2 keypoint is matched once" (mapcamidx2idxs, :1001-1002)
//      over the queries in order with warp-wide minima, then the rotation histogram + ComputeThreeMaxima (:1137-1156) and
//      the compaction of the surviving pairs in creation order.
// Integer / byte work, bit-exact against the oracle.  Bytes per pair: 56 B per keypoint of both keyframes + 4 B per
// FeatureVector entry read, 8 B per match written (HBM / L2 gathers; POPC-bound inner loop).
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace vieo {

constexpr int kSftThreads = 512, kSftWarps = kSftThreads / 32;
constexpr int kSftListCap = 32;      // candidates kept per query (one per lane in phase C)
constexpr int kSftMaxKp = 8192;      // keypoints per keyframe (13-bit fields of the candidate key)
constexpr int TH_LOW_SFT = 50, HISTO = 30;

__device__ __forceinline__ int ham256(const uint4& a0, const uint4& a1, const uint8_t* __restrict__ b) {
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(b)), e = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  return __popc(a0.x ^ c.x) + __popc(a0.y ^ c.y) + __popc(a0.z ^ c.z) + __popc(a0.w ^ c.w) + __popc(a1.x ^ e.x) +
         __popc(a1.y ^ e.y) + __popc(a1.z ^ e.z) + __popc(a1.w ^ e.w);
}

// GeometricCamera::epipolarConstrain, fundamental-matrix branch (common/camera_models/camera_base.h:360-404), Tdata = float,
// Tcalc = double; every operation an explicit _rn intrinsic so that no contraction changes a bit
__device__ __forceinline__ bool epipolar_ok(const double* F, float x1, float y1, float x2, float y2, float unc) {
  const double p1x = (double)x1, p1y = (double)y1, p2x = (double)x2, p2y = (double)y2;
  const float a = (float)__dadd_rn(__dadd_rn(__dmul_rn(p1x, F[0]), __dmul_rn(p1y, F[3])), F[6]);
  const float b = (float)__dadd_rn(__dadd_rn(__dmul_rn(p1x, F[1]), __dmul_rn(p1y, F[4])), F[7]);
  const float c = (float)__dadd_rn(__dadd_rn(__dmul_rn(p1x, F[2]), __dmul_rn(p1y, F[5])), F[8]);
  const float num = (float)__dadd_rn(__dadd_rn(__dmul_rn((double)a, p2x), __dmul_rn((double)b, p2y)), (double)c);
  const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
  if (den == 0) return false;
  const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
  return dsqr < __fmul_rn(3.84f, unc);
}

struct SftShared {
  uint32_t claimed[kSftMaxKp / 32];  // keyframe-2 keypoints already matched
  int hist[HISTO];
  int scan[kSftWarps];
  int nq, carry;
};

// candidate key: distance (6 bits) | 8191 - position in the node list (13 bits) | idx2 (13 bits): the minimum is the
// smallest distance and, among equals, the LAST visited
__device__ __forceinline__ uint32_t sft_key(int dist, int pos, int idx2) {
  return ((uint32_t)dist << 26) | ((uint32_t)(8191 - pos) << 13) | (uint32_t)idx2;
}

__global__ void __launch_bounds__(kSftThreads) k_sft(const VieoSftPair* __restrict__ pairs_in,
                                                     const VieoKeyPoint* __restrict__ kps, const float* __restrict__ uright,
                                                     const uint8_t* __restrict__ desc, const uint8_t* __restrict__ has_mp,
                                                     const int32_t* __restrict__ fv_node, const int32_t* __restrict__ fv_ptr,
                                                     const int32_t* __restrict__ fv_idx, int32_t* __restrict__ match12,
                                                     int32_t* __restrict__ out_pairs, int32_t* __restrict__ n_matches,
                                                     int32_t* __restrict__ scr_node2, int32_t* __restrict__ scr_qptr,
                                                     int32_t* __restrict__ scr_qnode, int32_t* __restrict__ scr_cnt,
                                                     uint32_t* __restrict__ scr_list, int32_t* __restrict__ scr_bin) {
  __shared__ SftShared S;
  __shared__ VieoSftPair P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (int)(sizeof(VieoSftPair) / 4); i += kSftThreads)
    reinterpret_cast<uint32_t*>(&P)[i] = reinterpret_cast<const uint32_t*>(pairs_in + blockIdx.x)[i];
  for (int i = tid; i < kSftMaxKp / 32; i += kSftThreads) S.claimed[i] = 0;
  if (tid == 0) S.carry = 0;
  __syncthreads();
  const int n1 = P.n_kp1, n2 = P.n_kp2, nn1 = P.n_nodes1, nn2 = P.n_nodes2;
  int32_t* m12 = match12 + P.out_begin;
  int32_t* opairs = out_pairs + 2 * (size_t)P.out_begin;
  if (n1 < 0 || n2 < 0 || n1 > kSftMaxKp || n2 > kSftMaxKp || nn1 < 0 || nn2 < 0) {
    if (tid == 0) n_matches[blockIdx.x] = -1;
    return;
  }
  const VieoKeyPoint* K1 = kps + P.kp1_begin;
  const VieoKeyPoint* K2 = kps + P.kp2_begin;
  const float* U1 = uright + P.kp1_begin;
  const float* U2 = uright + P.kp2_begin;
  const uint8_t* D1 = desc + 32 * (size_t)P.kp1_begin;
  const uint8_t* D2 = desc + 32 * (size_t)P.kp2_begin;
  const uint8_t* M1 = has_mp + P.kp1_begin;
  const uint8_t* M2 = has_mp + P.kp2_begin;
  const int32_t* N1 = fv_node + P.node1_begin;
  const int32_t* N2 = fv_node + P.node2_begin;
  const int32_t* T1 = fv_ptr + P.ptr1_begin;  // n_nodes + 1 entries each, values relative to idx*_begin
  const int32_t* T2 = fv_ptr + P.ptr2_begin;
  const int32_t* I1 = fv_idx + P.idx1_begin;
  const int32_t* I2 = fv_idx + P.idx2_begin;
  int32_t* node2 = scr_node2 + P.nscr_begin;    // per node of keyframe 1: matching node of keyframe 2 or -1
  int32_t* qptr = scr_qptr + P.nscr_begin;      // per node of keyframe 1: first query number
  int32_t* qnode = scr_qnode + P.out_begin;     // per query: its node of keyframe 1
  int32_t* qcnt = scr_cnt + P.out_begin;        // per query: candidates stored (-1: skipped)
  uint32_t* qlist = scr_list + (size_t)P.out_begin * kSftListCap;
  int32_t* qbin = scr_bin + P.out_begin;        // per query: rotation bin of its match (-1: none)
  for (int i = tid; i < n1; i += kSftThreads) m12[i] = -1;
  // ---- phase A: node intersection + query numbering ---------------------------------------------------------------------
  for (int base = 0; base < nn1; base += kSftThreads) {
    const int a = base + tid;
    int cnt = 0, hit = -1;
    if (a < nn1) {
      const int id = N1[a];
      int lo = 0, hi = nn2;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (N2[mid] < id) lo = mid + 1;
        else hi = mid;
      }
      if (lo < nn2 && N2[lo] == id) {
        hit = lo;
        cnt = T1[a + 1] - T1[a];
      }
      node2[a] = hit;
    }
    const int inc = warp_incl_scan(cnt, lane);
    if (lane == 31) S.scan[warp] = inc;
    __syncthreads();
    int off = S.carry;
    for (int w = 0; w < warp; ++w) off += S.scan[w];
    if (a < nn1) {
      const int q0 = off + inc - cnt;
      qptr[a] = q0;
      for (int k = 0; k < cnt; ++k) qnode[q0 + k] = a;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kSftWarps; ++w) t += S.scan[w];
      S.carry += t;
    }
    __syncthreads();
  }
  const int nq = S.carry;
  // ---- phase B: candidates of every query that pass the state-independent tests ------------------------------------------
  for (int q = warp; q < nq; q += kSftWarps) {
    const int a = qnode[q], b = node2[a];
    const int idx1 = I1[T1[a] + (q - qptr[a])];
    int n = 0;
    bool skip = M1[idx1] != 0;
    const bool st1 = U1[idx1] >= 0;
    if (P.only_stereo && !st1) skip = true;
    if (skip) {
      if (lane == 0) qcnt[q] = -1;
      continue;
    }
    const VieoKeyPoint k1 = K1[idx1];
    const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(D1 + 32 * (size_t)idx1));
    const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(D1 + 32 * (size_t)idx1) + 1);
    const int j0 = T2[b], j1 = T2[b + 1];
    for (int base = j0; base < j1; base += 32) {
      const int j = base + lane;
      bool ok = false;
      uint32_t key = 0;
      if (j < j1) {
        const int idx2 = I2[j];
        if (!M2[idx2]) {
          const bool st2 = U2[idx2] >= 0;
          if (!(P.only_stereo && !st2)) {
            const int dist = ham256(d0, d1, D2 + 32 * (size_t)idx2);
            if (dist <= TH_LOW_SFT) {
              const VieoKeyPoint k2 = K2[idx2];
              bool pass = true;
              if (!st1 && !st2) {
                const float dx = __fsub_rn(P.ex, k2.x), dy = __fsub_rn(P.ey, k2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, P.scale_factor2[k2.octave])) pass = false;
              }
              if (pass && epipolar_ok(P.F12, k1.x, k1.y, k2.x, k2.y, P.level_sigma2_2[k2.octave])) {
                ok = true;
                key = sft_key(dist, j - j0, idx2);
              }
            }
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int pos = n + __popc(bal & ((1u << lane) - 1));
        if (pos < kSftListCap) qlist[(size_t)q * kSftListCap + pos] = key;
      }
      n += __popc(bal);
    }
    if (lane == 0) qcnt[q] = n;
  }
  __syncthreads();
  if (warp != 0) return;
  // ---- phase C: sequential over the queries (a keyframe-2 keypoint is matched once) ---------------------------------------
  int nmatches = 0;
  const float factor = 1.0f / HISTO;
  for (int q = 0; q < nq; ++q) {
    const int n = qcnt[q];
    uint32_t best = 0xffffffffu;
    if (n > 0 && n <= kSftListCap) {
      if (lane < n) {
        const uint32_t e = qlist[(size_t)q * kSftListCap + lane];
        const int idx2 = (int)(e & 0x1fffu);
        if (!((S.claimed[idx2 >> 5] >> (idx2 & 31)) & 1u)) best = e;
      }
    } else if (n > kSftListCap) {  // overflow: re-run the query's tests with the claim filter
      const int a = qnode[q], b = node2[a];
      const int idx1 = I1[T1[a] + (q - qptr[a])];
      const bool st1 = U1[idx1] >= 0;
      const VieoKeyPoint k1 = K1[idx1];
      const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(D1 + 32 * (size_t)idx1));
      const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(D1 + 32 * (size_t)idx1) + 1);
      const int j0 = T2[b], j1 = T2[b + 1];
      for (int j = j0 + lane; j < j1; j += 32) {
        const int idx2 = I2[j];
        if (M2[idx2] || ((S.claimed[idx2 >> 5] >> (idx2 & 31)) & 1u)) continue;
        const bool st2 = U2[idx2] >= 0;
        if (P.only_stereo && !st2) continue;
        const int dist = ham256(d0, d1, D2 + 32 * (size_t)idx2);
        if (dist > TH_LOW_SFT) continue;
        const VieoKeyPoint k2 = K2[idx2];
        if (!st1 && !st2) {
          const float dx = __fsub_rn(P.ex, k2.x), dy = __fsub_rn(P.ey, k2.y);
          if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, P.scale_factor2[k2.octave])) continue;
        }
        if (!epipolar_ok(P.F12, k1.x, k1.y, k2.x, k2.y, P.level_sigma2_2[k2.octave])) continue;
        best = min(best, sft_key(dist, j - j0, idx2));
      }
    }
    int bin = -1;
    if (n > 0) {
      const uint32_t m = __reduce_min_sync(0xffffffffu, best);
      if (m != 0xffffffffu) {
        const int idx2 = (int)(m & 0x1fffu);
        const int a = qnode[q];
        const int idx1 = I1[T1[a] + (q - qptr[a])];
        if (lane == 0) {
          S.claimed[idx2 >> 5] |= 1u << (idx2 & 31);
          m12[idx1] = idx2;
          opairs[2 * q] = idx1;      // creation order = query order; compacted below
          opairs[2 * q + 1] = idx2;
        }
        ++nmatches;
        bin = 0;
        if (P.check_orientation) {
          float rot = __fsub_rn(K1[idx1].angle, K2[idx2].angle);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == HISTO) bin = 0;
        }
      }
    }
    if (lane == 0) qbin[q] = bin;  // -1: no match created by this query
    __syncwarp();
  }
  // ---- rotation consistency (:1137-1156) -----------------------------------------------------------------------------------
  int ind1 = -1, ind2 = -1, ind3 = -1;
  if (P.check_orientation) {
    int h = 0;
    for (int q = 0; q < nq; ++q) h += (qbin[q] == lane) ? 1 : 0;
    if (lane < HISTO) S.hist[lane] = h;
    __syncwarp();
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < HISTO; i++) {
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
  // ---- vMatchedPairs: surviving matches in creation order (:1158-1183) --------------------------------------------------------
  int n_out = 0;
  for (int base = 0; base < nq; base += 32) {
    const int q = base + lane;
    bool keep = false;
    int i1 = -1, i2 = -1;
    if (q < nq) {
      const int bin = qbin[q];
      if (bin >= 0) {
        i1 = opairs[2 * q];
        i2 = opairs[2 * q + 1];
        keep = !P.check_orientation || bin == ind1 || bin == ind2 || bin == ind3;
        if (!keep) m12[i1] = -1;  // goodmatches[...] = false
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncwarp();  // every lane has read its slot before the compacted writes below (they never pass the read position)
    if (keep) {
      const int o = n_out + __popc(bal & ((1u << lane) - 1));
      opairs[2 * o] = i1;
      opairs[2 * o + 1] = i2;
    }
    n_out += __popc(bal);
  }
  if (lane == 0) n_matches[blockIdx.x] = n_out;
}

// ------------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (src/ORBmatcher.cc:344-505), single camera: for every
// keyframe keypoint with a live map point, in FeatureVector order, the best / second-best Hamming distance over the frame
// keypoints of the same vocabulary node that have not been matched yet (:385), accepted when best <= TH_LOW and
// best < mfNNratio * second (:408-411); rotation histogram + ComputeThreeMaxima (:463-489).
// A frame keypoint lies in exactly one node and — with one camera per keyframe — a map point on exactly one keyframe
// keypoint, so the reference's sequential state (vpMapPointMatches, the (pMP, img_id) table) never couples two nodes: the
// matched nodes are processed by the CTA's warps IN PARALLEL, each warp walking its node's keyframe keypoints in order with
// lanes over the node's frame keypoints.  Histogram counts are order-free.  One CTA per (keyframe, frame) pair.
struct BowShared {
  uint32_t claimed[kSftMaxKp / 32];
  int hist[HISTO];
  int ind[3];
  int erased;
};

__global__ void __launch_bounds__(kSftThreads) k_bow(const VieoBowPair* __restrict__ pairs_in,
                                                     const VieoKeyPoint* __restrict__ kps, const uint8_t* __restrict__ desc,
                                                     const uint8_t* __restrict__ mp_ok, const int32_t* __restrict__ fv_node,
                                                     const int32_t* __restrict__ fv_ptr, const int32_t* __restrict__ fv_idx,
                                                     int32_t* __restrict__ match_f, int32_t* __restrict__ n_matches,
                                                     int8_t* __restrict__ scr_bin) {
  __shared__ BowShared S;
  __shared__ VieoBowPair P;
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (int)(sizeof(VieoBowPair) / 4); i += kSftThreads)
    reinterpret_cast<uint32_t*>(&P)[i] = reinterpret_cast<const uint32_t*>(pairs_in + blockIdx.x)[i];
  for (int i = tid; i < kSftMaxKp / 32; i += kSftThreads) S.claimed[i] = 0;
  if (tid < HISTO) S.hist[tid] = 0;
  if (tid == 0) {
    S.erased = 0;
    s_total = 0;
  }
  __syncthreads();
  const int n1 = P.n_kp1, n2 = P.n_kp2, nn1 = P.n_nodes1, nn2 = P.n_nodes2;
  if (n1 < 0 || n2 < 0 || n1 > kSftMaxKp || n2 > kSftMaxKp || nn1 < 0 || nn2 < 0) {
    if (tid == 0) n_matches[blockIdx.x] = -1;
    return;
  }
  const VieoKeyPoint* K1 = kps + P.kp1_begin;
  const VieoKeyPoint* K2 = kps + P.kp2_begin;
  const uint8_t* D1 = desc + 32 * (size_t)P.kp1_begin;
  const uint8_t* D2 = desc + 32 * (size_t)P.kp2_begin;
  const uint8_t* M1 = mp_ok + P.kp1_begin;
  const int32_t* N1 = fv_node + P.node1_begin;
  const int32_t* N2 = fv_node + P.node2_begin;
  const int32_t* T1 = fv_ptr + P.ptr1_begin;
  const int32_t* T2 = fv_ptr + P.ptr2_begin;
  const int32_t* I1 = fv_idx + P.idx1_begin;
  const int32_t* I2 = fv_idx + P.idx2_begin;
  int32_t* MF = match_f + P.out_begin;
  int8_t* BIN = scr_bin + P.out_begin;
  for (int i = tid; i < n2; i += kSftThreads) MF[i] = -1;
  __syncthreads();
  const float factor = 1.0f / HISTO;
  int mine = 0;
  for (int a = warp; a < nn1; a += kSftWarps) {
    const int id = N1[a];
    int lo = 0, hi = nn2;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (N2[mid] < id) lo = mid + 1;
      else hi = mid;
    }
    if (!(lo < nn2 && N2[lo] == id)) continue;
    const int j0 = T2[lo], j1 = T2[lo + 1];
    for (int i = T1[a]; i < T1[a + 1]; ++i) {
      const int idx1 = I1[i];
      if (!M1[idx1]) continue;
      const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(D1 + 32 * (size_t)idx1));
      const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(D1 + 32 * (size_t)idx1) + 1);
      // lane-local two smallest keys (distance << 16 | position): strict '<' of :396-402 = lexicographic minimum
      uint32_t best = 0xffffffffu, second = 0xffffffffu;
      for (int j = j0 + lane; j < j1; j += 32) {
        const int idx2 = I2[j];
        if ((S.claimed[idx2 >> 5] >> (idx2 & 31)) & 1u) continue;
        const uint32_t key = ((uint32_t)ham256(d0, d1, D2 + 32 * (size_t)idx2) << 16) | (uint32_t)(j - j0);
        if (key < best) {
          second = best;
          best = key;
        } else if (key < second) {
          second = key;
        }
      }
      const uint32_t m1 = __reduce_min_sync(0xffffffffu, best);
      if (m1 == 0xffffffffu) continue;
      const uint32_t m2 = __reduce_min_sync(0xffffffffu, best == m1 ? second : best);
      const int bestDist1 = (int)(m1 >> 16), bestDist2 = m2 == 0xffffffffu ? 256 : (int)(m2 >> 16);
      if (bestDist1 <= TH_LOW_SFT && (float)bestDist1 < __fmul_rn(P.nn_ratio, (float)bestDist2)) {
        const int idx2 = I2[j0 + (int)(m1 & 0xffffu)];
        if (lane == 0) {
          atomicOr(&S.claimed[idx2 >> 5], 1u << (idx2 & 31));  // other nodes' keypoints (other warps) share the word
          MF[idx2] = idx1;
          int bin = 0;
          if (P.check_orientation) {
            float rot = __fsub_rn(K1[idx1].angle, K2[idx2].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == HISTO) bin = 0;
            atomicAdd(&S.hist[bin], 1);
          }
          BIN[idx2] = (int8_t)bin;
        }
        ++mine;
        __syncwarp();
      }
    }
  }
  if (lane == 0 && mine) atomicAdd(&s_total, mine);
  __syncthreads();
  if (P.check_orientation) {
    if (tid == 0) {
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < HISTO; i++) {
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
      S.ind[0] = ind1; S.ind[1] = ind2; S.ind[2] = ind3;
    }
    __syncthreads();
    int er = 0;
    for (int i = tid; i < n2; i += kSftThreads) {
      if (MF[i] < 0) continue;
      const int bin = BIN[i];
      if (bin != S.ind[0] && bin != S.ind[1] && bin != S.ind[2]) {
        MF[i] = -1;
        ++er;
      }
    }
    if (er) atomicAdd(&S.erased, er);
    __syncthreads();
  }
  if (tid == 0) n_matches[blockIdx.x] = s_total - S.erased;
}

}  // namespace vieo

using namespace vieo;

extern "C" {

size_t vieo_sft_scratch_bytes(int n_out_total, int n_nodes1_total) {
  const int n_kp1_total = n_out_total;
  // node2 | qptr (per node of the first keyframes), qnode | cnt | bin (per keypoint), list (per keypoint x kSftListCap)
  return sizeof(int32_t) * (2 * (size_t)std::max(n_nodes1_total, 1) + (3 + (size_t)kSftListCap) * (size_t)std::max(n_kp1_total, 1)) + 256;
}

int vieo_search_for_triangulation_dev(const VieoSftPair* pairs_dev, int n_pairs, const VieoKeyPoint* kps_dev,
                                      const float* uright_dev, const uint8_t* desc_dev, const uint8_t* has_mp_dev,
                                      const int32_t* fv_node_dev, const int32_t* fv_ptr_dev, const int32_t* fv_idx_dev,
                                      int n_out_total, int n_nodes1_total, int32_t* match12_dev, int32_t* pairs_out_dev,
                                      int32_t* n_matches_dev, void* scratch_dev, size_t scratch_bytes, void* stream) {
  VIEO_ARG(n_pairs >= 0, "bad argument");
  if (n_pairs == 0) return VIEO_OK;
  VIEO_ARG(pairs_dev && kps_dev && uright_dev && desc_dev && has_mp_dev && fv_node_dev && fv_ptr_dev && fv_idx_dev &&
               match12_dev && pairs_out_dev && n_matches_dev && scratch_dev,
           "null argument");
  VIEO_ARG(scratch_bytes >= vieo_sft_scratch_bytes(n_out_total, n_nodes1_total), "scratch too small (vieo_sft_scratch_bytes)");
  VIEO_ARG((uintptr_t)desc_dev % 16 == 0, "descriptors must be 16-byte aligned");
  int32_t* s = (int32_t*)scratch_dev;
  const size_t nn = (size_t)std::max(n_nodes1_total, 1), nk = (size_t)std::max(n_out_total, 1);
  int32_t* node2 = s;
  int32_t* qptr = node2 + nn;
  int32_t* qnode = qptr + nn;
  int32_t* cnt = qnode + nk;
  int32_t* bin = cnt + nk;
  uint32_t* list = (uint32_t*)(bin + nk);
  k_sft<<<n_pairs, kSftThreads, 0, (cudaStream_t)stream>>>(pairs_dev, kps_dev, uright_dev, desc_dev, has_mp_dev, fv_node_dev,
                                                          fv_ptr_dev, fv_idx_dev, match12_dev, pairs_out_dev, n_matches_dev, node2,
                                                          qptr, qnode, cnt, list, bin);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

// host-buffer form: stages everything on the calling thread's scratch stream, one launch, results back
int vieo_search_for_triangulation(const VieoSftPair* pairs, int n_pairs, const VieoKeyPoint* kps, const float* uright,
                                  const uint8_t* desc, const uint8_t* has_mp, const int32_t* fv_node, const int32_t* fv_ptr,
                                  const int32_t* fv_idx, int n_kp_total, int n_node_total, int n_ptr_total, int n_idx_total,
                                  int n_out_total, int n_nodes1_total, int32_t* match12, int32_t* pairs_out, int32_t* n_matches,
                                  int device) {
  VIEO_ARG(n_pairs >= 0, "bad argument");
  if (n_pairs == 0) return VIEO_OK;
  VIEO_ARG(pairs && kps && uright && desc && has_mp && fv_node && fv_ptr && fv_idx && match12 && pairs_out && n_matches,
           "null argument");
  VIEO_ARG(n_kp_total >= 0 && n_node_total >= 0 && n_ptr_total >= 0 && n_idx_total >= 0 && n_out_total >= 0 && n_nodes1_total >= 0,
           "bad sizes");
  for (int p = 0; p < n_pairs; ++p) {  // ranges are checked here, the kernel trusts them
    const VieoSftPair& P = pairs[p];
    VIEO_ARG(P.n_kp1 >= 0 && P.n_kp2 >= 0 && P.kp1_begin >= 0 && P.kp2_begin >= 0 && P.kp1_begin + P.n_kp1 <= n_kp_total &&
                 P.kp2_begin + P.n_kp2 <= n_kp_total, "keypoint range out of bounds");
    VIEO_ARG(P.n_kp1 <= kSftMaxKp && P.n_kp2 <= kSftMaxKp, "more than 8192 keypoints in a keyframe");
    VIEO_ARG(P.n_nodes1 >= 0 && P.n_nodes2 >= 0 && P.node1_begin >= 0 && P.node2_begin >= 0 &&
                 P.node1_begin + P.n_nodes1 <= n_node_total && P.node2_begin + P.n_nodes2 <= n_node_total, "node range out of bounds");
    VIEO_ARG(P.ptr1_begin >= 0 && P.ptr2_begin >= 0 && P.ptr1_begin + P.n_nodes1 + 1 <= n_ptr_total &&
                 P.ptr2_begin + P.n_nodes2 + 1 <= n_ptr_total, "ptr range out of bounds");
    VIEO_ARG(P.idx1_begin >= 0 && P.idx2_begin >= 0 && P.idx1_begin + fv_ptr[P.ptr1_begin + P.n_nodes1] <= n_idx_total &&
                 P.idx2_begin + fv_ptr[P.ptr2_begin + P.n_nodes2] <= n_idx_total, "index range out of bounds");
    VIEO_ARG(fv_ptr[P.ptr1_begin + P.n_nodes1] <= P.n_kp1, "a keypoint of keyframe 1 appears in more than one node");
    VIEO_ARG(P.out_begin >= 0 && P.out_begin + P.n_kp1 <= n_out_total && P.nscr_begin >= 0 &&
                 P.nscr_begin + P.n_nodes1 <= n_nodes1_total, "output range out of bounds");
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  const size_t sizes[8] = {sizeof(VieoSftPair) * (size_t)n_pairs, sizeof(VieoKeyPoint) * (size_t)n_kp_total, 4 * (size_t)n_kp_total,
                           32 * (size_t)n_kp_total, (size_t)n_kp_total, 4 * (size_t)n_node_total, 4 * (size_t)n_ptr_total,
                           4 * (size_t)n_idx_total};
  const void* src[8] = {pairs, kps, uright, desc, has_mp, fv_node, fv_ptr, fv_idx};
  void* d[8];
  for (int i = 0; i < 8; ++i) {
    d[i] = cs->get(i, std::max<size_t>(sizes[i], 16));
    if (!d[i]) return VIEO_E_CUDA;
    if (sizes[i]) VIEO_CK(cudaMemcpyAsync(d[i], src[i], sizes[i], cudaMemcpyHostToDevice, st));
  }
  const size_t no = (size_t)std::max(n_out_total, 1);
  int32_t* dout = (int32_t*)cs->get(8, 4 * (3 * no + (size_t)n_pairs));
  const size_t sb = vieo_sft_scratch_bytes(n_out_total, n_nodes1_total);
  void* scr = cs->get(9, sb);
  if (!dout || !scr) return VIEO_E_CUDA;
  VIEO_CK(cudaStreamSynchronize(st));  // the sources may be pageable
  rc = vieo_search_for_triangulation_dev((const VieoSftPair*)d[0], n_pairs, (const VieoKeyPoint*)d[1], (const float*)d[2],
                                         (const uint8_t*)d[3], (const uint8_t*)d[4], (const int32_t*)d[5], (const int32_t*)d[6],
                                         (const int32_t*)d[7], n_out_total, n_nodes1_total, dout, dout + no, dout + 3 * no, scr, sb, st);
  if (rc) return rc;
  if (n_out_total) {
    VIEO_CK(cudaMemcpyAsync(match12, dout, 4 * (size_t)n_out_total, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(pairs_out, dout + no, 8 * (size_t)n_out_total, cudaMemcpyDeviceToHost, st));
  }
  VIEO_CK(cudaMemcpyAsync(n_matches, dout + 3 * no, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

int vieo_search_by_bow_dev(const VieoBowPair* pairs_dev, int n_pairs, const VieoKeyPoint* kps_dev, const uint8_t* desc_dev,
                           const uint8_t* mp_ok_dev, const int32_t* fv_node_dev, const int32_t* fv_ptr_dev,
                           const int32_t* fv_idx_dev, int n_out_total, int32_t* match_f_dev, int32_t* n_matches_dev,
                           void* scratch_dev, size_t scratch_bytes, void* stream) {
  VIEO_ARG(n_pairs >= 0, "bad argument");
  if (n_pairs == 0) return VIEO_OK;
  VIEO_ARG(pairs_dev && kps_dev && desc_dev && mp_ok_dev && fv_node_dev && fv_ptr_dev && fv_idx_dev && match_f_dev &&
               n_matches_dev && scratch_dev, "null argument");
  VIEO_ARG(scratch_bytes >= (size_t)std::max(n_out_total, 1), "scratch too small (one byte per frame keypoint)");
  VIEO_ARG((uintptr_t)desc_dev % 16 == 0, "descriptors must be 16-byte aligned");
  k_bow<<<n_pairs, kSftThreads, 0, (cudaStream_t)stream>>>(pairs_dev, kps_dev, desc_dev, mp_ok_dev, fv_node_dev, fv_ptr_dev,
                                                          fv_idx_dev, match_f_dev, n_matches_dev, (int8_t*)scratch_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_search_by_bow(const VieoBowPair* pairs, int n_pairs, const VieoKeyPoint* kps, const uint8_t* desc, const uint8_t* mp_ok,
                       const int32_t* fv_node, const int32_t* fv_ptr, const int32_t* fv_idx, int n_kp_total, int n_node_total,
                       int n_ptr_total, int n_idx_total, int n_out_total, int32_t* match_f, int32_t* n_matches, int device) {
  VIEO_ARG(n_pairs >= 0, "bad argument");
  if (n_pairs == 0) return VIEO_OK;
  VIEO_ARG(pairs && kps && desc && mp_ok && fv_node && fv_ptr && fv_idx && match_f && n_matches, "null argument");
  for (int p = 0; p < n_pairs; ++p) {
    const VieoBowPair& P = pairs[p];
    VIEO_ARG(P.n_kp1 >= 0 && P.n_kp2 >= 0 && P.kp1_begin >= 0 && P.kp2_begin >= 0 && P.kp1_begin + P.n_kp1 <= n_kp_total &&
                 P.kp2_begin + P.n_kp2 <= n_kp_total, "keypoint range out of bounds");
    VIEO_ARG(P.n_kp1 <= kSftMaxKp && P.n_kp2 <= kSftMaxKp, "more than 8192 keypoints");
    VIEO_ARG(P.n_nodes1 >= 0 && P.n_nodes2 >= 0 && P.node1_begin >= 0 && P.node2_begin >= 0 &&
                 P.node1_begin + P.n_nodes1 <= n_node_total && P.node2_begin + P.n_nodes2 <= n_node_total, "node range out of bounds");
    VIEO_ARG(P.ptr1_begin >= 0 && P.ptr2_begin >= 0 && P.ptr1_begin + P.n_nodes1 + 1 <= n_ptr_total &&
                 P.ptr2_begin + P.n_nodes2 + 1 <= n_ptr_total, "ptr range out of bounds");
    VIEO_ARG(P.idx1_begin >= 0 && P.idx2_begin >= 0 && P.idx1_begin + fv_ptr[P.ptr1_begin + P.n_nodes1] <= n_idx_total &&
                 P.idx2_begin + fv_ptr[P.ptr2_begin + P.n_nodes2] <= n_idx_total, "index range out of bounds");
    VIEO_ARG(P.out_begin >= 0 && P.out_begin + P.n_kp2 <= n_out_total, "output range out of bounds");
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  const size_t sizes[7] = {sizeof(VieoBowPair) * (size_t)n_pairs, sizeof(VieoKeyPoint) * (size_t)n_kp_total, 32 * (size_t)n_kp_total,
                           (size_t)n_kp_total, 4 * (size_t)n_node_total, 4 * (size_t)n_ptr_total, 4 * (size_t)n_idx_total};
  const void* src[7] = {pairs, kps, desc, mp_ok, fv_node, fv_ptr, fv_idx};
  void* d[7];
  for (int i = 0; i < 7; ++i) {
    d[i] = cs->get(i, std::max<size_t>(sizes[i], 16));
    if (!d[i]) return VIEO_E_CUDA;
    if (sizes[i]) VIEO_CK(cudaMemcpyAsync(d[i], src[i], sizes[i], cudaMemcpyHostToDevice, st));
  }
  const size_t no = (size_t)std::max(n_out_total, 1);
  int32_t* dout = (int32_t*)cs->get(8, 4 * (no + (size_t)n_pairs));
  void* scr = cs->get(9, no);
  if (!dout || !scr) return VIEO_E_CUDA;
  VIEO_CK(cudaStreamSynchronize(st));
  rc = vieo_search_by_bow_dev((const VieoBowPair*)d[0], n_pairs, (const VieoKeyPoint*)d[1], (const uint8_t*)d[2],
                              (const uint8_t*)d[3], (const int32_t*)d[4], (const int32_t*)d[5], (const int32_t*)d[6], n_out_total,
                              dout, dout + no, scr, no, st);
  if (rc) return rc;
  if (n_out_total) VIEO_CK(cudaMemcpyAsync(match_f, dout, 4 * (size_t)n_out_total, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(n_matches, dout + no, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

}  // extern "C"

// 256-bit Hamming matching on sm_100a: brute-force k=2 (cv::BFMatcher knnMatch semantics, src/Frame.cc:620-628)
// and candidate-list best/second-best search (src/ORBmatcher.cc:286-315, 1396-1424; src/Frame.cc:506-523).
// Integer-ALU / shared-memory bound: one query per lane held in registers, train rows staged
// in shared memory and read as 128-bit broadcasts, __popc on 8 x u32.
#include <climits>
#include <vector>

#include "common.cuh"

namespace vieo {

constexpr int kKnnWarps = 8;      // warps per CTA; each warp scans a disjoint contiguous train range
constexpr int kKnnTile = 256;     // train rows staged per step (8 KiB)

struct Top2 {
  int d0, i0, d1, i1;
};
__device__ __forceinline__ void top2_push(Top2& t, int d, int i) {  // candidates arrive in ascending index order
  if (d < t.d0) {
    t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i;
  } else if (d < t.d1) {
    t.d1 = d; t.i1 = i;
  }
}
__device__ __forceinline__ bool top2_less(int da, int ia, int db, int ib) {
  return da < db || (da == db && (unsigned)ia < (unsigned)ib);  // idx -1 (absent) sorts last
}

// grid = (ceil(max_nq / 32), n_pairs); block = 32 x kKnnWarps.  Lane = query; warp w scans train rows of
// tile chunk w; partial top-2 lists are merged in shared memory with the lowest-index tie rule.
__global__ void __launch_bounds__(32 * kKnnWarps) k_knn2(const uint8_t* __restrict__ q, size_t q_stride,
                                                        const int* __restrict__ nq_p, int max_nq,
                                                        const uint8_t* __restrict__ t, size_t t_stride,
                                                        const int* __restrict__ nt_p, int max_nt, int count_stride,
                                                        int* __restrict__ idx, int* __restrict__ dist) {
  __shared__ uint4 s_t[kKnnTile * 2];
  __shared__ Top2 s_m[kKnnWarps][32];
  const int pair = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = min(nq_p[pair * count_stride], max_nq), nt = min(nt_p[pair * count_stride], max_nt);
  const int qi = blockIdx.x * 32 + lane;
  if (blockIdx.x * 32 >= nq) return;
  const uint4* qp = reinterpret_cast<const uint4*>(q + pair * q_stride);
  const uint4* tp = reinterpret_cast<const uint4*>(t + pair * t_stride);
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (qi < nq) {
    a = __ldg(qp + 2 * qi);
    b = __ldg(qp + 2 * qi + 1);
  }
  Top2 best = {INT_MAX, -1, INT_MAX, -1};
  for (int base = 0; base < nt; base += kKnnTile) {
    const int n = min(kKnnTile, nt - base);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * n; i += 32 * kKnnWarps) s_t[i] = __ldg(tp + 2 * base + i);
    __syncthreads();
    const int per = kKnnTile / kKnnWarps;
    const int j0 = warp * per, j1 = min(j0 + per, n);
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const uint4 c = s_t[2 * j], e = s_t[2 * j + 1];
      const int d = __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) +
                    __popc(b.x ^ e.x) + __popc(b.y ^ e.y) + __popc(b.z ^ e.z) + __popc(b.w ^ e.w);
      top2_push(best, d, base + j);
    }
  }
  s_m[warp][lane] = best;
  __syncthreads();
  if (warp == 0 && qi < nq) {
    Top2 r = {INT_MAX, -1, INT_MAX, -1};
#pragma unroll
    for (int w = 0; w < kKnnWarps; ++w) {
      const Top2 m = s_m[w][lane];
      // insert (m.d0,m.i0) then (m.d1,m.i1) with (distance, index) ordering
      int cd[2] = {m.d0, m.d1}, ci[2] = {m.i0, m.i1};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (ci[e] < 0) continue;
        if (top2_less(cd[e], ci[e], r.d0, r.i0)) {
          r.d1 = r.d0; r.i1 = r.i0; r.d0 = cd[e]; r.i0 = ci[e];
        } else if (top2_less(cd[e], ci[e], r.d1, r.i1)) {
          r.d1 = cd[e]; r.i1 = ci[e];
        }
      }
    }
    const size_t o = ((size_t)pair * max_nq + qi) * 2;
    idx[o] = r.i0; idx[o + 1] = r.i1; dist[o] = r.d0; dist[o + 1] = r.d1;
  }
}

// One warp per row: lanes stride over the row's candidate list; warp-shuffle arg-min(2) keyed by
// (distance, list position) so that the first of equal distances wins, as the reference's strict '<' does.
__global__ void __launch_bounds__(128) k_csr(const uint8_t* __restrict__ q, const uint8_t* __restrict__ t,
                                            const int* __restrict__ row_ptr, const int* __restrict__ cand, int nrows,
                                            int* __restrict__ bd, int* __restrict__ bi, int* __restrict__ sd,
                                            int* __restrict__ si) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const uint4* qp = reinterpret_cast<const uint4*>(q) + 2 * (size_t)row;
  const uint4 a = __ldg(qp), b = __ldg(qp + 1);
  const int k0 = row_ptr[row], k1 = row_ptr[row + 1];
  // key = distance << 22 | list position (lists < 4M entries); UINT_MAX = absent
  unsigned m0 = 0xffffffffu, m1 = 0xffffffffu;
  for (int k = k0 + lane; k < k1; k += 32) {
    const uint4* tp = reinterpret_cast<const uint4*>(t) + 2 * (size_t)cand[k];
    const uint4 c = __ldg(tp), e = __ldg(tp + 1);
    const unsigned d = __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) +
                       __popc(b.x ^ e.x) + __popc(b.y ^ e.y) + __popc(b.z ^ e.z) + __popc(b.w ^ e.w);
    const unsigned key = (d << 22) | (unsigned)(k - k0);
    if (key < m0) {
      m1 = m0; m0 = key;
    } else if (key < m1)
      m1 = key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned o0 = __shfl_xor_sync(0xffffffffu, m0, o), o1 = __shfl_xor_sync(0xffffffffu, m1, o);
    const unsigned lo = min(m0, o0), hi = max(m0, o0);
    m1 = min(hi, min(m1, o1));
    m0 = lo;
  }
  if (lane == 0) {
    // the reference tracks the second best only by distance: a later equal distance does not replace the
    // best but does become the second best (d < bestDist2), which the (distance, position) order reproduces.
    bd[row] = m0 == 0xffffffffu ? 256 : (int)(m0 >> 22);
    bi[row] = m0 == 0xffffffffu ? -1 : cand[k0 + (int)(m0 & 0x3fffff)];
    sd[row] = m1 == 0xffffffffu ? 256 : (int)(m1 >> 22);
    si[row] = m1 == 0xffffffffu ? -1 : cand[k0 + (int)(m1 & 0x3fffff)];
  }
}


// ------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:314-378) for a batch of map points (LocalMapping calls it for
// every point it created, fused or re-observed: src/LocalMapping.cc).  One warp per map point; for
// row i the lanes hold the distances to observations lane, lane + 32, ... in registers (the first 256; a longer list —
// not seen in practice — recomputes the rest inside the search) and the sorted-row entry at index (N - 1) / 2 is found
// without sorting: bisection on the value v in [0, 256] of "count(d <= v) >= k + 1" with warp ballots.  The first row
// with the least median wins, as in the reference's strict '<' scan.
constexpr int kDdWarps = 4;
constexpr int kDdChunks = 8;
__device__ __forceinline__ int ham256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}
__global__ void __launch_bounds__(kDdWarps * 32) k_distinctive(const uint8_t* __restrict__ pool,
                                                              const int32_t* __restrict__ rows,
                                                              const int32_t* __restrict__ ptr, int n_points,
                                                              int32_t* __restrict__ best, int32_t* __restrict__ median) {
  const int lane = threadIdx.x & 31, p = blockIdx.x * kDdWarps + (threadIdx.x >> 5);
  if (p >= n_points) return;
  const int b = ptr[p], N = ptr[p + 1] - b;
  if (N <= 0) {
    if (lane == 0) {
      best[p] = -1;
      median[p] = -1;
    }
    return;
  }
  auto row_of = [&](int j) { return rows ? rows[b + j] : b + j; };
  auto load = [&](int j, uint4& d0, uint4& d1) {
    const uint4* src = reinterpret_cast<const uint4*>(pool + 32 * (size_t)row_of(j));
    d0 = __ldg(src);
    d1 = __ldg(src + 1);
  };
  const int kth = (N - 1) >> 1;  // vDists[0.5 * (N - 1)]
  const int chunks = min((N + 31) >> 5, kDdChunks);
  uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;  // this lane's observation of chunk 0
  if (lane < N) load(lane, m0, m1);
  int best_med = INT_MAX, best_idx = 0;
  for (int i = 0; i < N; ++i) {
    uint4 a0, a1;
    load(i, a0, a1);
    int d[kDdChunks];
#pragma unroll
    for (int c = 0; c < kDdChunks; ++c) {
      d[c] = 512;  // beyond any distance
      const int j = lane + 32 * c;
      if (c < chunks && j < N) {
        if (c == 0) {
          d[c] = ham256(a0, a1, m0, m1);
        } else {
          uint4 b0, b1;
          load(j, b0, b1);
          d[c] = ham256(a0, a1, b0, b1);
        }
      }
    }
    int lo = 0, hi = 256;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < kDdChunks; ++c)
        if (c < chunks) cnt += __popc(__ballot_sync(0xffffffffu, d[c] <= mid));
      if (N > 32 * kDdChunks) {
        int extra = 0;
        for (int j = 32 * kDdChunks + lane; j < N; j += 32) {
          uint4 b0, b1;
          load(j, b0, b1);
          extra += ham256(a0, a1, b0, b1) <= mid ? 1 : 0;
        }
        cnt += __reduce_add_sync(0xffffffffu, extra);
      }
      if (cnt >= kth + 1) hi = mid;
      else lo = mid + 1;
    }
    if (lo < best_med) {
      best_med = lo;
      best_idx = i;
    }
  }
  if (lane == 0) {
    best[p] = best_idx;
    median[p] = best_med;
  }
}

}  // namespace vieo

using namespace vieo;

extern "C" {

int vieo_hamming_knn2_batch_dev(const uint8_t* q_dev, size_t q_stride, const int32_t* nq_dev, int max_nq,
                                const uint8_t* t_dev, size_t t_stride, const int32_t* nt_dev, int max_nt,
                                int count_stride, int n_pairs, int32_t* idx_dev, int32_t* dist_dev, void* stream) {
  VIEO_ARG(q_dev && t_dev && nq_dev && nt_dev && idx_dev && dist_dev, "null argument");
  VIEO_ARG(max_nq >= 1 && max_nt >= 0 && n_pairs >= 1 && n_pairs <= 65535, "bad sizes");
  VIEO_ARG(((uintptr_t)q_dev | (uintptr_t)t_dev | q_stride | t_stride) % 16 == 0, "descriptors must be 16-byte aligned");
  k_knn2<<<dim3((max_nq + 31) / 32, n_pairs), 32 * kKnnWarps, 0, (cudaStream_t)stream>>>(
      q_dev, q_stride, nq_dev, max_nq, t_dev, t_stride, nt_dev, max_nt, count_stride, idx_dev, dist_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist, int device) {
  VIEO_ARG(nq >= 0 && nt >= 0 && (nq == 0 || (q && idx && dist)) && (nt == 0 || t), "bad argument");
  if (nq == 0) return VIEO_OK;
  int rc = use_device(device);
  if (rc) return rc;
  // per-thread staging (device buffers and stream reused across calls): a call costs copies + one kernel
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  uint8_t* dq = (uint8_t*)cs->get(0, (size_t)nq * 32);
  uint8_t* dt = (uint8_t*)cs->get(1, (size_t)std::max(nt, 1) * 32);
  int* dn = (int*)cs->get(2, 8);
  int* dio = (int*)cs->get(3, (size_t)nq * 16);
  if (!dq || !dt || !dn || !dio) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  const int hn[2] = {nq, nt};
  VIEO_CK(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, st));
  if (nt) VIEO_CK(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(dn, hn, 8, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaStreamSynchronize(st));  // hn is a local; q / t may be pageable
  rc = vieo_hamming_knn2_batch_dev(dq, 0, dn, nq, dt, 0, dn + 1, nt, 1, 1, dio, dio + 2 * (size_t)nq, st);
  if (rc) return rc;
  VIEO_CK(cudaMemcpyAsync(idx, dio, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(dist, dio + 2 * (size_t)nq, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

int vieo_hamming_csr(const uint8_t* q, const uint8_t* t, int nt, const int32_t* row_ptr, const int32_t* cand,
                     int nrows, int32_t* best_dist, int32_t* best_idx, int32_t* second_dist, int32_t* second_idx,
                     int device) {
  VIEO_ARG(nrows >= 0 && nt >= 0 && row_ptr, "bad argument");
  if (nrows == 0) return VIEO_OK;
  VIEO_ARG(q && best_dist && best_idx && second_dist && second_idx, "null argument");
  const int ncand = row_ptr[nrows];
  VIEO_ARG(ncand >= 0 && ncand < (1 << 22) && (ncand == 0 || (cand && t)), "candidate list too long or null");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  uint8_t* dq = (uint8_t*)cs->get(0, (size_t)nrows * 32);
  uint8_t* dt = (uint8_t*)cs->get(1, (size_t)std::max(nt, 1) * 32);
  int* drp = (int*)cs->get(2, sizeof(int) * ((size_t)nrows + 1));
  int* dc = (int*)cs->get(3, sizeof(int) * (size_t)std::max(ncand, 1));
  int* dout = (int*)cs->get(4, sizeof(int) * 4 * (size_t)nrows);
  if (!dq || !dt || !drp || !dc || !dout) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  VIEO_CK(cudaMemcpyAsync(dq, q, (size_t)nrows * 32, cudaMemcpyHostToDevice, st));
  if (nt) VIEO_CK(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(drp, row_ptr, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, st));
  if (ncand) VIEO_CK(cudaMemcpyAsync(dc, cand, sizeof(int) * (size_t)ncand, cudaMemcpyHostToDevice, st));
  k_csr<<<(nrows + 3) / 4, 128, 0, st>>>(dq, dt, drp, dc, nrows, dout, dout + nrows, dout + 2 * (size_t)nrows,
                                         dout + 3 * (size_t)nrows);
  VIEO_CK(cudaGetLastError());
  VIEO_CK(cudaMemcpyAsync(best_dist, dout, sizeof(int) * nrows, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(best_idx, dout + nrows, sizeof(int) * nrows, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(second_dist, dout + 2 * (size_t)nrows, sizeof(int) * nrows, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(second_idx, dout + 3 * (size_t)nrows, sizeof(int) * nrows, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

int vieo_distinctive_descriptors_dev(const uint8_t* desc_pool_dev, const int32_t* rows_dev, const int32_t* ptr_dev,
                                     int n_points, int32_t* best_dev, int32_t* median_dev, void* stream) {
  VIEO_ARG(n_points >= 0, "bad argument");
  if (n_points == 0) return VIEO_OK;
  VIEO_ARG(desc_pool_dev && ptr_dev && best_dev && median_dev, "null argument");
  VIEO_ARG((uintptr_t)desc_pool_dev % 16 == 0, "descriptors must be 16-byte aligned");
  k_distinctive<<<(n_points + kDdWarps - 1) / kDdWarps, kDdWarps * 32, 0, (cudaStream_t)stream>>>(
      desc_pool_dev, rows_dev, ptr_dev, n_points, best_dev, median_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_distinctive_descriptors(const uint8_t* desc_pool, int n_pool, const int32_t* rows, const int32_t* ptr,
                                 int n_points, int32_t* best, int32_t* median, int device) {
  VIEO_ARG(n_points >= 0 && n_pool >= 0, "bad argument");
  if (n_points == 0) return VIEO_OK;
  VIEO_ARG(ptr && best && median, "null argument");
  const int n_obs = ptr[n_points];
  VIEO_ARG(ptr[0] == 0 && n_obs >= 0 && (n_obs == 0 || desc_pool), "bad observation lists");
  for (int p = 0; p < n_points; ++p) VIEO_ARG(ptr[p + 1] >= ptr[p], "observation lists must be ascending");
  if (rows) {
    for (int k = 0; k < n_obs; ++k) VIEO_ARG(rows[k] >= 0 && rows[k] < n_pool, "descriptor row out of range");
  } else {
    VIEO_ARG(n_obs <= n_pool, "descriptor row out of range");
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  uint8_t* d_pool = (uint8_t*)cs->get(0, 32 * (size_t)std::max(n_pool, 1));
  int32_t* d_rows = (int32_t*)cs->get(1, 4 * (size_t)std::max(n_obs, 1));
  int32_t* d_ptr = (int32_t*)cs->get(2, 4 * (size_t)(n_points + 1));
  int32_t* d_out = (int32_t*)cs->get(3, 8 * (size_t)n_points);
  if (!d_pool || !d_rows || !d_ptr || !d_out) return VIEO_E_CUDA;
  cudaStream_t st = cs->st;
  if (n_pool) VIEO_CK(cudaMemcpyAsync(d_pool, desc_pool, 32 * (size_t)n_pool, cudaMemcpyHostToDevice, st));
  if (rows && n_obs) VIEO_CK(cudaMemcpyAsync(d_rows, rows, 4 * (size_t)n_obs, cudaMemcpyHostToDevice, st));
  VIEO_CK(cudaMemcpyAsync(d_ptr, ptr, 4 * (size_t)(n_points + 1), cudaMemcpyHostToDevice, st));
  rc = vieo_distinctive_descriptors_dev(d_pool, rows ? d_rows : nullptr, d_ptr, n_points, d_out, d_out + n_points, st);
  if (rc) return rc;
  VIEO_CK(cudaMemcpyAsync(best, d_out, 4 * (size_t)n_points, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(median, d_out + n_points, 4 * (size_t)n_points, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

}  // extern "C"

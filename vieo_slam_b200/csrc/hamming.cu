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
  uint8_t *dq = nullptr, *dt = nullptr;
  int *dn = nullptr, *di = nullptr, *dd = nullptr;
  const int hn[2] = {nq, nt};
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  step(cudaMalloc(&dq, (size_t)nq * 32));
  step(cudaMalloc(&dt, (size_t)std::max(nt, 1) * 32));
  step(cudaMalloc(&dn, 8));
  step(cudaMalloc(&di, (size_t)nq * 8));
  step(cudaMalloc(&dd, (size_t)nq * 8));
  if (e == cudaSuccess) {
    step(cudaMemcpy(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice));
    if (nt) step(cudaMemcpy(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice));
    step(cudaMemcpy(dn, hn, 8, cudaMemcpyHostToDevice));
  }
  if (e == cudaSuccess) {
    rc = vieo_hamming_knn2_batch_dev(dq, 0, dn, nq, dt, 0, dn + 1, nt, 1, 1, di, dd, nullptr);
    if (rc == VIEO_OK) {
      step(cudaMemcpy(idx, di, (size_t)nq * 8, cudaMemcpyDeviceToHost));
      step(cudaMemcpy(dist, dd, (size_t)nq * 8, cudaMemcpyDeviceToHost));
    }
  }
  cudaFree(dq); cudaFree(dt); cudaFree(dn); cudaFree(di); cudaFree(dd);
  if (e != cudaSuccess) {
    set_error("vieo_hamming_knn2: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  return rc;
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
  uint8_t *dq = nullptr, *dt = nullptr;
  int *drp = nullptr, *dc = nullptr, *dout = nullptr;
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  step(cudaMalloc(&dq, (size_t)nrows * 32));
  step(cudaMalloc(&dt, (size_t)std::max(nt, 1) * 32));
  step(cudaMalloc(&drp, sizeof(int) * (nrows + 1)));
  step(cudaMalloc(&dc, sizeof(int) * std::max(ncand, 1)));
  step(cudaMalloc(&dout, sizeof(int) * 4 * nrows));
  if (e == cudaSuccess) {
    step(cudaMemcpy(dq, q, (size_t)nrows * 32, cudaMemcpyHostToDevice));
    if (nt) step(cudaMemcpy(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice));
    step(cudaMemcpy(drp, row_ptr, sizeof(int) * (nrows + 1), cudaMemcpyHostToDevice));
    if (ncand) step(cudaMemcpy(dc, cand, sizeof(int) * ncand, cudaMemcpyHostToDevice));
  }
  if (e == cudaSuccess) {
    k_csr<<<(nrows + 3) / 4, 128>>>(dq, dt, drp, dc, nrows, dout, dout + nrows, dout + 2 * nrows, dout + 3 * nrows);
    step(cudaGetLastError());
    step(cudaMemcpy(best_dist, dout, sizeof(int) * nrows, cudaMemcpyDeviceToHost));
    step(cudaMemcpy(best_idx, dout + nrows, sizeof(int) * nrows, cudaMemcpyDeviceToHost));
    step(cudaMemcpy(second_dist, dout + 2 * nrows, sizeof(int) * nrows, cudaMemcpyDeviceToHost));
    step(cudaMemcpy(second_idx, dout + 3 * nrows, sizeof(int) * nrows, cudaMemcpyDeviceToHost));
  }
  cudaFree(dq); cudaFree(dt); cudaFree(drp); cudaFree(dc); cudaFree(dout);
  if (e != cudaSuccess) {
    set_error("vieo_hamming_csr: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  return VIEO_OK;
}

}  // extern "C"

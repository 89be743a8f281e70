// Multi-camera frames (KB8 / distorted-stereo rigs, BASELINE configs[3]: TUM-VI 4 x 512 x 512): what Frame::Frame does per
// frame on its per-camera std::threads — ORBextractor::operator() with the camera's lapping area (src/Frame.cc:259-278,
// src/ORBextractor.cc:1041-1057) — followed by the brute-force half of Frame::ComputeStereoFishEyeMatches
// (src/Frame.cc:613-663): BFMatcher(NORM_HAMMING).knnMatch(k = 2) between the in-area ("stereo") descriptors of EVERY
// camera pair (i < j), then Lowe's ratio test.  FillMatchesFromPair (camera-model triangulation, :664-680) is host policy
// and stays in the reference.
//
//   k_lapping_split   one CTA per image: stable partition of the level-ordered keypoints — in-area ones are written from
//                     the BACK of the arrays in reverse visiting order, the others from the front (monoIndex = their
//                     count) — by a block-wide exclusive scan of the in-area flags.  Bytes: 56 B read + 56 B written per
//                     keypoint (HBM streaming).
//   k_fisheye_knn     grid (query tiles, camera pairs, frames): lane = query descriptor of camera i (rows from num_mono[i]),
//                     warps scan shared-memory tiles of camera j's in-area descriptors; top-2 with the lowest-train-index
//                     tie rule of BFMatcher; epilogue = the reference's ratio test in its own arithmetic (float distances
//                     promoted to double against 0.7 / 0.9).  Integer-ALU (POPC) bound: nq * nt * 8 XOR+POPC per pair.
#include <climits>
#include <vector>

#include "common.cuh"

namespace vieo {

constexpr int kSplitThreads = 256;

__global__ void __launch_bounds__(kSplitThreads) k_lapping_split(const VieoKeyPoint* __restrict__ kin,
                                                                 const uint8_t* __restrict__ din,
                                                                 const int* __restrict__ n_kp, int cap,
                                                                 const int* __restrict__ lapping,  // [n_img][2] or null
                                                                 VieoKeyPoint* __restrict__ kout, uint8_t* __restrict__ dout,
                                                                 int* __restrict__ n_mono) {
  __shared__ int s_warp[kSplitThreads / 32];
  __shared__ int s_carry;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(n_kp[img], cap);
  const VieoKeyPoint* ki = kin + (size_t)img * cap;
  const uint4* di = reinterpret_cast<const uint4*>(din + (size_t)img * cap * 32);
  VieoKeyPoint* ko = kout + (size_t)img * cap;
  uint4* dO = reinterpret_cast<uint4*>(dout + (size_t)img * cap * 32);
  const bool has = lapping != nullptr;
  const float x0 = has ? (float)lapping[2 * img] : 0.f, x1 = has ? (float)lapping[2 * img + 1] : 0.f;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kSplitThreads) {
    const int i = base + tid;
    VieoKeyPoint k;
    bool in = false;
    if (i < n) {
      k = ki[i];
      in = has && k.x >= x0 && k.x <= x1;  // keypoint->pt.x >= (*pvLappingArea)[0] && pt.x <= (*pvLappingArea)[1]
    }
    // exclusive count of in-area keypoints before i
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    const int before_w = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_carry + before_w;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (i < n) {
      // in-area: stereoIndex counts down from n - 1; others: monoIndex counts up from 0
      const int dst = in ? n - 1 - before : i - before;
      ko[dst] = k;
      dO[2 * dst] = di[2 * i];
      dO[2 * dst + 1] = di[2 * i + 1];
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kSplitThreads / 32; ++w) t += s_warp[w];
      s_carry += t;
    }
    __syncthreads();
  }
  // pvLappingArea == nullptr: monoIndex is never advanced and operator() returns 0 (src/ORBextractor.cc:1003,1057)
  if (tid == 0) n_mono[img] = has ? n - s_carry : 0;
}

struct Top2 {
  int d0, i0, d1, i1;
};
__device__ __forceinline__ bool t2_less(int da, int ia, int db, int ib) { return da < db || (da == db && ia < ib); }
__device__ __forceinline__ void t2_push(Top2& t, int d, int i) {
  if (t2_less(d, i, t.d0, t.i0)) {
    t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i;
  } else if (t2_less(d, i, t.d1, t.i1)) {
    t.d1 = d; t.i1 = i;
  }
}

constexpr int kFkTile = 256, kFkWarps = 8;

// pair index -> (i, j), i < j, in the reference's loop order (src/Frame.cc:620-628)
__device__ __forceinline__ void pair_ij(int p, int n_cams, int& i, int& j) {
  i = 0;
  int left = p;
  while (left >= n_cams - 1 - i) {
    left -= n_cams - 1 - i;
    ++i;
  }
  j = i + 1 + left;
}

__global__ void __launch_bounds__(32 * kFkWarps) k_fisheye_knn(const uint8_t* __restrict__ desc, const int* __restrict__ n_kp,
                                                               const int* __restrict__ n_mono, int n_cams, int cap,
                                                               int frame_stride, int cam_stride, int* __restrict__ idx,
                                                               int* __restrict__ dist, uint8_t* __restrict__ good) {
  __shared__ uint4 s_t[kFkTile * 2];
  __shared__ Top2 s_m[kFkWarps][32];
  const int pair = blockIdx.y, frame = blockIdx.z, n_pairs = gridDim.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int ci, cj;
  pair_ij(pair, n_cams, ci, cj);
  const int img_i = frame * frame_stride + ci * cam_stride, img_j = frame * frame_stride + cj * cam_stride;
  const int ni = min(n_kp[img_i], cap), nj = min(n_kp[img_j], cap);
  const int mi = n_mono[img_i], mj = n_mono[img_j];
  // if (num_mono[i] >= rows_i || num_mono[j] >= rows_j) continue;  -> the pair has no matches at all
  const int nq = (mi >= ni || mj >= nj) ? 0 : ni - mi, nt = nq ? nj - mj : 0;
  const int qi = blockIdx.x * 32 + lane;
  const size_t o = (((size_t)frame * n_pairs + pair) * cap + qi);
  if (blockIdx.x * 32 >= nq) {  // rows past the query set: "no match" so that the outputs are fully defined
    if (qi < cap) {
      idx[2 * o] = idx[2 * o + 1] = -1;
      dist[2 * o] = dist[2 * o + 1] = INT_MAX;
      good[o] = 0;
    }
    return;
  }
  const uint4* qp = reinterpret_cast<const uint4*>(desc + ((size_t)img_i * cap + mi) * 32);
  const uint4* tp = reinterpret_cast<const uint4*>(desc + ((size_t)img_j * cap + mj) * 32);
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (qi < nq) {
    a = __ldg(qp + 2 * qi);
    b = __ldg(qp + 2 * qi + 1);
  }
  Top2 best = {INT_MAX, -1, INT_MAX, -1};
  for (int base = 0; base < nt; base += kFkTile) {
    const int n = min(kFkTile, nt - base);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * n; t += 32 * kFkWarps) s_t[t] = __ldg(tp + 2 * base + t);
    __syncthreads();
    const int per = kFkTile / kFkWarps;
    const int j0 = warp * per, j1 = min(j0 + per, n);
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const uint4 c = s_t[2 * j], e = s_t[2 * j + 1];
      const int d = __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) + __popc(b.x ^ e.x) +
                    __popc(b.y ^ e.y) + __popc(b.z ^ e.z) + __popc(b.w ^ e.w);
      t2_push(best, d, base + j);
    }
  }
  s_m[warp][lane] = best;
  __syncthreads();
  if (warp == 0 && qi < cap) {
    Top2 r = {INT_MAX, -1, INT_MAX, -1};
    if (qi < nq) {
#pragma unroll
      for (int w = 0; w < kFkWarps; ++w) {
        const Top2 m = s_m[w][lane];
        if (m.i0 >= 0) t2_push(r, m.d0, m.i0);
        if (m.i1 >= 0) t2_push(r, m.d1, m.i1);
      }
    }
    idx[2 * o] = r.i0; idx[2 * o + 1] = r.i1;
    dist[2 * o] = r.d0; dist[2 * o + 1] = r.d1;
    // (*it).size() >= 2 && (d0 < d1 * 0.7 || (d0 < thOrbDist && d0 < d1 * 0.9)), distances are floats, 0.7 / 0.9 doubles;
    // thOrbDist = (TH_HIGH + TH_LOW) / 2 = 75 (src/Frame.cc:659-663)
    bool g = false;
    if (r.i1 >= 0) {
      const double d0 = (double)(float)r.d0, d1 = (double)(float)r.d1;
      g = d0 < __dmul_rn(d1, 0.7) || (r.d0 < 75 && d0 < __dmul_rn(d1, 0.9));
    }
    good[o] = g ? 1 : 0;
  }
}

}  // namespace vieo

using namespace vieo;

extern "C" {

int vieo_lapping_split_dev(const VieoKeyPoint* kps_in_dev, const uint8_t* desc_in_dev, const int32_t* n_kp_dev, int n_img,
                           int cap, const int32_t* lapping_dev, VieoKeyPoint* kps_out_dev, uint8_t* desc_out_dev,
                           int32_t* n_mono_dev, void* stream) {
  VIEO_ARG(kps_in_dev && desc_in_dev && n_kp_dev && kps_out_dev && desc_out_dev && n_mono_dev, "null argument");
  VIEO_ARG(n_img >= 1 && cap >= 1, "bad sizes");
  VIEO_ARG(kps_in_dev != kps_out_dev && desc_in_dev != desc_out_dev, "the split is out of place");
  VIEO_ARG(((uintptr_t)desc_in_dev | (uintptr_t)desc_out_dev) % 16 == 0, "descriptors must be 16-byte aligned");
  k_lapping_split<<<n_img, kSplitThreads, 0, (cudaStream_t)stream>>>(kps_in_dev, desc_in_dev, n_kp_dev, cap, lapping_dev,
                                                                     kps_out_dev, desc_out_dev, n_mono_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_fisheye_knn_dev(const uint8_t* desc_dev, const int32_t* n_kp_dev, const int32_t* n_mono_dev, int n_cams,
                         int n_frames, int cap, int frame_stride, int cam_stride, int32_t* idx_dev, int32_t* dist_dev,
                         uint8_t* good_dev, void* stream) {
  VIEO_ARG(desc_dev && n_kp_dev && n_mono_dev && idx_dev && dist_dev && good_dev, "null argument");
  VIEO_ARG(n_cams >= 2 && n_cams <= 16 && n_frames >= 1 && n_frames <= 65535 && cap >= 1, "bad sizes");
  VIEO_ARG(frame_stride >= 1 && cam_stride >= 1, "bad strides");
  VIEO_ARG((uintptr_t)desc_dev % 16 == 0, "descriptors must be 16-byte aligned");
  const int n_pairs = n_cams * (n_cams - 1) / 2;
  k_fisheye_knn<<<dim3((cap + 31) / 32, n_pairs, n_frames), 32 * kFkWarps, 0, (cudaStream_t)stream>>>(
      desc_dev, n_kp_dev, n_mono_dev, n_cams, cap, frame_stride, cam_stride, idx_dev, dist_dev, good_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_multicam_frames(vieo_orb_t* h, int n_frames, int n_cams, const uint8_t* imgs, size_t img_stride, int row_stride,
                         const int32_t* lapping, VieoKeyPoint* kps, uint8_t* desc, int32_t* n_kp, int32_t* n_mono,
                         int32_t* pair_idx, int32_t* pair_dist, uint8_t* pair_good) {
  VIEO_ARG(h && imgs && kps && desc && n_kp && n_mono, "null argument");
  VIEO_ARG(n_frames >= 1 && n_cams >= 1 && n_cams <= 16, "bad sizes");
  VIEO_ARG(n_cams == 1 || (pair_idx && pair_dist && pair_good), "null pair output");
  const int n_img = n_frames * n_cams;
  int device = 0, max_batch = 0;
  orb_info(h, &device, &max_batch);
  VIEO_ARG(n_img <= max_batch, "n_frames * n_cams exceeds the handle's max_batch");
  VIEO_CK(cudaSetDevice(device));
  int rc = orb_enqueue_host(h, n_img, imgs, img_stride, row_stride);
  if (rc) return rc;
  VieoKeyPoint* dk;
  uint8_t* dd;
  int* dn;
  int cap;
  cudaStream_t st;
  orb_dev_outputs(h, &dk, &dd, &dn, &cap, &st);
  const int n_pairs = n_cams * (n_cams - 1) / 2;
  // per-thread staging for the reordered outputs and the pair results
  CallScratch* cs = call_scratch(device);
  if (!cs) return VIEO_E_CUDA;
  VieoKeyPoint* k2 = (VieoKeyPoint*)cs->get(0, sizeof(VieoKeyPoint) * (size_t)n_img * cap);
  uint8_t* d2 = (uint8_t*)cs->get(1, (size_t)32 * n_img * cap);
  int* lap = (int*)cs->get(2, sizeof(int) * 2 * (size_t)n_img);
  int* nm = (int*)cs->get(3, sizeof(int) * (size_t)n_img);
  const size_t np = (size_t)n_frames * std::max(n_pairs, 1) * cap;
  int* pi = (int*)cs->get(4, sizeof(int) * 2 * np);
  int* pd = (int*)cs->get(5, sizeof(int) * 2 * np);
  uint8_t* pg = (uint8_t*)cs->get(6, np);
  if (!k2 || !d2 || !lap || !nm || !pi || !pd || !pg) return VIEO_E_CUDA;
  if (lapping) {
    int* hl = (int*)cs->get_pinned(sizeof(int) * 2 * (size_t)n_img);
    if (!hl) return VIEO_E_CUDA;
    for (int f = 0; f < n_frames; ++f)
      for (int c = 0; c < n_cams; ++c) {
        hl[2 * (f * n_cams + c)] = lapping[2 * c];
        hl[2 * (f * n_cams + c) + 1] = lapping[2 * c + 1];
      }
    VIEO_CK(cudaMemcpyAsync(lap, hl, sizeof(int) * 2 * (size_t)n_img, cudaMemcpyHostToDevice, st));
  }
  rc = vieo_lapping_split_dev(dk, dd, dn, n_img, cap, lapping ? lap : nullptr, k2, d2, nm, st);
  if (rc) return rc;
  if (n_pairs > 0) {
    rc = vieo_fisheye_knn_dev(d2, dn, nm, n_cams, n_frames, cap, n_cams, 1, pi, pd, pg, st);
    if (rc) return rc;
  }
  VIEO_CK(cudaMemcpyAsync(kps, k2, sizeof(VieoKeyPoint) * (size_t)n_img * cap, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(desc, d2, (size_t)32 * n_img * cap, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(n_kp, dn, sizeof(int) * n_img, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaMemcpyAsync(n_mono, nm, sizeof(int) * n_img, cudaMemcpyDeviceToHost, st));
  if (n_pairs > 0) {
    VIEO_CK(cudaMemcpyAsync(pair_idx, pi, sizeof(int) * 2 * np, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(pair_dist, pd, sizeof(int) * 2 * np, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(pair_good, pg, np, cudaMemcpyDeviceToHost, st));
  }
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

}  // extern "C"

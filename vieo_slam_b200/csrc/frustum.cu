// Frame::isInFrustum (src/Frame.cc:335-416) + MapPoint::PredictScale (src/MapPoint.cc:491-509) on sm_100a, and the fused
// Tracking::SearchLocalPoints step (src/Tracking.cc:2308-2368): visibility test of every local map point followed by
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, th, th_far) (src/ORBmatcher.cc:230-335, sbp.cu) on the same
// stream — the tracking info (u, v, ur, level, viewCos, depth) goes from one kernel to the next through HBM and never
// visits the host.  Single camera (mpCameras.size() == 1: GetTcr() is the identity), usedistort_ == false.
//
// Arithmetic: the reference computes in float through Eigen 3.3 fixed-size expressions, whose 3-term reductions unroll
// to a0 + (a1 + a2) (redux_novec_unroller); every operation below is an explicit round-to-nearest float intrinsic in
// that order, so the kernel is bit-identical to the no-contraction oracle.  PredictScale's ceil(logf(ratio) / logf(s))
// is NOT evaluated with a device logf (not correctly rounded -> level flips at the boundaries): the host tabulates, with
// the same libm logf the reference calls, the smallest float ratio that reaches each level (vieo_frustum_level_table,
// a bisection over float bit patterns, valid because logf is monotone) and the kernel counts thresholds.
// Memory-bound streaming kernel: 32 B read + 29 B written per map point, one thread per point, coalesced SoA.
#include <math.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace vieo {
namespace {

constexpr int kFrThreads = 256;
constexpr int kFrBlocksPerFrame = 8;

__device__ __forceinline__ float sum3(float a0, float a1, float a2) { return __fadd_rn(a0, __fadd_rn(a1, a2)); }

__global__ void __launch_bounds__(kFrThreads) k_frustum(const VieoFrustumFrame* __restrict__ frames,
                                                        const float* __restrict__ wP, const float* __restrict__ Pn,
                                                        const float* __restrict__ max_dist,
                                                        const float* __restrict__ min_dist,
                                                        const uint8_t* __restrict__ skip, uint8_t* __restrict__ inview,
                                                        float* __restrict__ proj, int32_t* __restrict__ level,
                                                        float* __restrict__ viewcos, float* __restrict__ depth,
                                                        int32_t* __restrict__ n_inview) {
  __shared__ VieoFrustumFrame F;
  __shared__ int s_cnt;
  const int f = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(VieoFrustumFrame) / 4); i += kFrThreads)
    reinterpret_cast<uint32_t*>(&F)[i] = reinterpret_cast<const uint32_t*>(frames + f)[i];
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int qi = blockIdx.x * kFrThreads + tid; qi < F.n_q; qi += gridDim.x * kFrThreads) {
    const size_t q = (size_t)F.q_begin + qi;
    bool in = false;
    float u = 0, v = 0, ur = 0, vc = 0, d3 = 0;
    int lvl = -1;
    if (!(skip && skip[q])) {
      const float X = wP[3 * q], Y = wP[3 * q + 1], Z = wP[3 * q + 2];
      float Pc[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        Pc[r] = __fadd_rn(sum3(__fmul_rn(F.Rcw[3 * r], X), __fmul_rn(F.Rcw[3 * r + 1], Y), __fmul_rn(F.Rcw[3 * r + 2], Z)),
                          F.tcw[r]);
      const float PcZ = Pc[2];
      if (!(PcZ < 0.0f)) {
        const float invz = __fdiv_rn(1.0f, PcZ);
        const float xn = __fmul_rn(Pc[0], invz), yn = __fmul_rn(Pc[1], invz);
        // K.cast<float>() * (xn, yn, 1): the zero entries of K are multiplied too (NaN / inf propagate as in Eigen)
        u = sum3(__fmul_rn(F.fx, xn), __fmul_rn(0.0f, yn), __fmul_rn(F.cx, 1.0f));
        v = sum3(__fmul_rn(0.0f, xn), __fmul_rn(F.fy, yn), __fmul_rn(F.cy, 1.0f));
        if (!(u < F.minx || u > F.maxx) && !(v < F.miny || v > F.maxy)) {
          const float ox = __fsub_rn(X, F.Ow[0]), oy = __fsub_rn(Y, F.Ow[1]), oz = __fsub_rn(Z, F.Ow[2]);
          d3 = __fsqrt_rn(sum3(__fmul_rn(ox, ox), __fmul_rn(oy, oy), __fmul_rn(oz, oz)));
          const float mx = max_dist[q];
          const float maxD = __fmul_rn(1.2f, mx), minD = __fmul_rn(0.8f, min_dist[q]);
          if (!(d3 < minD || d3 > maxD)) {
            vc = __fdiv_rn(sum3(__fmul_rn(ox, Pn[3 * q]), __fmul_rn(oy, Pn[3 * q + 1]), __fmul_rn(oz, Pn[3 * q + 2])), d3);
            if (!(vc < F.cos_limit)) {
              const float ratio = __fdiv_rn(mx, d3);
              lvl = 0;
              for (int k = 1; k < F.n_levels; ++k) lvl += ratio >= F.level_ratio[k] ? 1 : 0;
              ur = __fsub_rn(u, __fmul_rn(F.bf, invz));
              in = true;
            }
          }
        }
      }
    }
    if (!in) u = v = ur = vc = d3 = 0.0f;
    inview[q] = in ? 1 : 0;
    level[q] = lvl;
    proj[3 * q] = u;
    proj[3 * q + 1] = v;
    proj[3 * q + 2] = ur;
    viewcos[q] = vc;
    depth[q] = d3;
    mine += in ? 1 : 0;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (tid == 0 && s_cnt) atomicAdd(n_inview + f, s_cnt);
}

// Rig form (mpCameras.size() > 1): the same thread-per-point kernel, with the per-camera steps of src/Frame.cc:351-411 in a
// loop over <= 4 cameras.  Float operations are explicit round-to-nearest intrinsics in the reference's order (Eigen's
// _transformVector: uv = q.vec x p; uv += uv; p + w uv + q.vec x uv); with usedistort_ the camera's Project runs in double
// as written in common/camera_models (library compiled with -fmad=false).  112 B written per point (4 camera slots).
__device__ __forceinline__ void cross3f(const float a[3], const float b[3], float o[3]) {
  o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
  o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
  o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}
__device__ __forceinline__ void rig_project_double(const VieoFrustumCam& c, const float Pc[3], float& u, float& v) {
  const double x = (double)Pc[0], y = (double)Pc[1];
  if (c.model == 2) {
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r = sqrt(r2);
    const float precision_r = 1e-5f;
    if (r > (double)precision_r) {
      const double z = (double)Pc[2];
      const double theta = atan2(r, z), theta2 = theta * theta;
      double thetad = (double)c.k[3] * theta2;
      thetad += (double)c.k[2];
      thetad *= theta2;
      thetad += (double)c.k[1];
      thetad *= theta2;
      thetad += (double)c.k[0];
      thetad *= theta2;
      thetad += 1;
      thetad *= theta;
      const double mx = x * thetad / r, my = y * thetad / r;
      const double invz = 1. / 1.;
      u = (float)((double)c.fx * mx * invz + (double)c.cx);
      v = (float)((double)c.fy * my * invz + (double)c.cy);
      return;
    }
  }
  const double z = (double)Pc[2], invz = 1. / z;
  u = (float)((double)c.fx * x * invz + (double)c.cx);
  v = (float)((double)c.fy * y * invz + (double)c.cy);
}

__global__ void __launch_bounds__(kFrThreads) k_frustum_rig(const VieoFrustumRigFrame* __restrict__ frames,
                                                            const float* __restrict__ wP, const float* __restrict__ Pn,
                                                            const float* __restrict__ max_dist,
                                                            const float* __restrict__ min_dist,
                                                            const uint8_t* __restrict__ skip, uint8_t* __restrict__ inview,
                                                            uint8_t* __restrict__ cam_mask, float* __restrict__ proj,
                                                            int32_t* __restrict__ level, float* __restrict__ viewcos,
                                                            float* __restrict__ depth, int32_t* __restrict__ n_inview) {
  __shared__ VieoFrustumRigFrame F;
  __shared__ int s_cnt;
  const int f = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(VieoFrustumRigFrame) / 4); i += kFrThreads)
    reinterpret_cast<uint32_t*>(&F)[i] = reinterpret_cast<const uint32_t*>(frames + f)[i];
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int qi = blockIdx.x * kFrThreads + tid; qi < F.n_q; qi += gridDim.x * kFrThreads) {
    const size_t q = (size_t)F.q_begin + qi;
    float o_proj[12], o_vc[4];
    int o_lvl[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      o_lvl[c] = -1;
      o_vc[c] = 0.0f;
      o_proj[3 * c] = o_proj[3 * c + 1] = o_proj[3 * c + 2] = 0.0f;
    }
    unsigned mask = 0;
    float sum_depth = 0.0f;
    int count = 0;
    if (!(skip && skip[q])) {
      const float X = wP[3 * q], Y = wP[3 * q + 1], Z = wP[3 * q + 2];
      const float mx = max_dist[q];
      const float maxD = __fmul_rn(1.2f, mx), minD = __fmul_rn(0.8f, min_dist[q]);
      float Pcr[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        Pcr[r] = __fadd_rn(sum3(__fmul_rn(F.Rcw[3 * r], X), __fmul_rn(F.Rcw[3 * r + 1], Y), __fmul_rn(F.Rcw[3 * r + 2], Z)),
                           F.tcw[r]);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci >= F.n_cams) continue;
        const VieoFrustumCam& C = F.cam[ci];
        const float qv[3] = {C.q_cr[0], C.q_cr[1], C.q_cr[2]};
        float uv[3], cr[3], Pc[3], twc[3];
        cross3f(qv, Pcr, uv);
#pragma unroll
        for (int k = 0; k < 3; ++k) uv[k] = __fadd_rn(uv[k], uv[k]);
        cross3f(qv, uv, cr);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          Pc[k] = __fadd_rn(__fadd_rn(__fadd_rn(Pcr[k], __fmul_rn(C.q_cr[3], uv[k])), cr[k]), C.t_cr[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          twc[k] = __fadd_rn(F.Ow[k], sum3(__fmul_rn(F.Rcw[k], C.t_rc[0]), __fmul_rn(F.Rcw[3 + k], C.t_rc[1]),
                                           __fmul_rn(F.Rcw[6 + k], C.t_rc[2])));
        const float PcZ = Pc[2];
        if (PcZ < 0.0f) continue;
        const float invz = __fdiv_rn(1.0f, PcZ);
        float u, v;
        if (C.model == 0) {
          const float xn = __fmul_rn(Pc[0], invz), yn = __fmul_rn(Pc[1], invz);
          u = sum3(__fmul_rn(C.fx, xn), __fmul_rn(0.0f, yn), __fmul_rn(C.cx, 1.0f));
          v = sum3(__fmul_rn(0.0f, xn), __fmul_rn(C.fy, yn), __fmul_rn(C.cy, 1.0f));
        } else {
          rig_project_double(C, Pc, u, v);
        }
        if (u < C.minx || u > C.maxx) continue;
        if (v < C.miny || v > C.maxy) continue;
        const float ox = __fsub_rn(X, twc[0]), oy = __fsub_rn(Y, twc[1]), oz = __fsub_rn(Z, twc[2]);
        const float d3 = __fsqrt_rn(sum3(__fmul_rn(ox, ox), __fmul_rn(oy, oy), __fmul_rn(oz, oz)));
        if (d3 < minD || d3 > maxD) continue;
        const float vc = __fdiv_rn(sum3(__fmul_rn(ox, Pn[3 * q]), __fmul_rn(oy, Pn[3 * q + 1]), __fmul_rn(oz, Pn[3 * q + 2])), d3);
        if (vc < F.cos_limit) continue;
        const float ratio = __fdiv_rn(mx, d3);
        int lvl = 0;
        for (int k = 1; k < F.n_levels; ++k) lvl += ratio >= F.level_ratio[k] ? 1 : 0;
        o_lvl[ci] = lvl;
        o_vc[ci] = vc;
        o_proj[3 * ci] = u;
        o_proj[3 * ci + 1] = v;
        o_proj[3 * ci + 2] = __fsub_rn(u, __fmul_rn(F.bf, invz));
        mask |= 1u << ci;
        sum_depth = __fadd_rn(sum_depth, d3);
        ++count;
      }
    }
    const bool in = count > 0;
    inview[q] = in ? 1 : 0;
    cam_mask[q] = (uint8_t)mask;
    depth[q] = in ? __fdiv_rn(sum_depth, (float)count) : 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      level[4 * q + c] = o_lvl[c];
      viewcos[4 * q + c] = o_vc[c];
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) proj[12 * q + k] = o_proj[k];
    mine += in ? 1 : 0;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (tid == 0 && s_cnt) atomicAdd(n_inview + f, s_cnt);
}

// ceil(logf(ratio) / lsf) >= k, evaluated as the reference does (float log, float divide, ceil)
inline bool level_reached(float ratio, float lsf, int k) {
  const float x = std::ceil(std::log(ratio) / lsf);
  return x >= (float)k;  // NaN -> false
}

}  // namespace
}  // namespace vieo

using namespace vieo;

extern "C" {

int vieo_frustum_level_table(float log_scale_factor, int n_levels, float table[16]) {
  VIEO_ARG(table && n_levels >= 1 && n_levels <= 16, "bad pyramid depth");
  VIEO_ARG(log_scale_factor > 0.0f && std::isfinite(log_scale_factor), "log scale factor must be positive");
  // a batch repeats one pyramid: the last table is kept per thread (the bisection costs ~1100 logf calls)
  thread_local float c_lsf = 0.0f, c_table[16];
  thread_local int c_levels = 0;
  if (c_levels == n_levels && c_lsf == log_scale_factor) {
    memcpy(table, c_table, sizeof(c_table));
    return VIEO_OK;
  }
  for (int k = 0; k < 16; ++k) table[k] = INFINITY;
  table[0] = 0.0f;
  for (int k = 1; k < n_levels; ++k) {
    // smallest positive float (as a bit pattern) that reaches level k; +inf always does
    uint32_t lo = 0x00000001u, hi = 0x7f800000u;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      float r;
      memcpy(&r, &mid, 4);
      if (level_reached(r, log_scale_factor, k)) hi = mid;
      else lo = mid + 1;
    }
    // logf must be monotone around the threshold for the table to be equivalent to the formula
    for (int d = -64; d <= 64; ++d) {
      const uint32_t b = lo + d;
      if (b < 1u || b > 0x7f800000u) continue;
      float r;
      memcpy(&r, &b, 4);
      if (level_reached(r, log_scale_factor, k) != (d >= 0)) {
        set_error("vieo_frustum_level_table: logf is not monotone near level %d", k);
        return VIEO_E_ARG;
      }
    }
    memcpy(&table[k], &lo, 4);
  }
  memcpy(c_table, table, sizeof(c_table));
  c_lsf = log_scale_factor;
  c_levels = n_levels;
  return VIEO_OK;
}

int vieo_frustum_batch_dev(const VieoFrustumFrame* frames_dev, int n_frames, const float* wP_dev, const float* normal_dev,
                           const float* max_dist_dev, const float* min_dist_dev, const uint8_t* skip_dev,
                           uint8_t* inview_dev, float* proj_dev, int32_t* level_dev, float* viewcos_dev, float* depth_dev,
                           int32_t* n_inview_dev, void* stream) {
  VIEO_ARG(n_frames >= 0 && n_frames <= 65535, "bad argument (at most 65535 frames per call)");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames_dev && wP_dev && normal_dev && max_dist_dev && min_dist_dev && inview_dev && proj_dev && level_dev &&
               viewcos_dev && depth_dev && n_inview_dev, "null argument");
  VIEO_CK(cudaMemsetAsync(n_inview_dev, 0, 4 * (size_t)n_frames, (cudaStream_t)stream));
  k_frustum<<<dim3(kFrBlocksPerFrame, n_frames), kFrThreads, 0, (cudaStream_t)stream>>>(
      frames_dev, wP_dev, normal_dev, max_dist_dev, min_dist_dev, skip_dev, inview_dev, proj_dev, level_dev, viewcos_dev,
      depth_dev, n_inview_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

// shared staging of the two host-buffer entry points; sbp == nullptr: visibility test only
static int frustum_host(const VieoFrustumFrame* frames, const VieoSbpFrame* sbp, int n_frames, const float* wP,
                        const float* normal, const float* max_dist, const float* min_dist, const uint8_t* skip,
                        const VieoKeyPoint* kps, const float* uright, const uint8_t* desc, const uint8_t* q_desc,
                        const uint8_t* q_flags, const uint8_t* kp_blocked, uint8_t* inview, float* proj, int32_t* level,
                        float* viewcos, float* depth, int32_t* n_inview, int32_t* kp_match, int32_t* q_match,
                        int32_t* q_dist, int32_t* n_matches, int device) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames && n_inview, "null argument");
  size_t nq = 0, nk = 0;
  std::vector<VieoFrustumFrame> fr(frames, frames + n_frames);
  for (int f = 0; f < n_frames; ++f) {
    VIEO_ARG(fr[f].q_begin >= 0 && fr[f].n_q >= 0, "bad frame range");
    int rc = vieo_frustum_level_table(fr[f].log_scale_factor, fr[f].n_levels, fr[f].level_ratio);
    if (rc) return rc;
    nq = std::max(nq, (size_t)fr[f].q_begin + fr[f].n_q);
    if (sbp) {
      VIEO_ARG(sbp[f].q_begin == fr[f].q_begin && sbp[f].n_q == fr[f].n_q, "frustum and search frames must share the query range");
      VIEO_ARG(sbp[f].n_kp >= 0 && sbp[f].kp_begin >= 0, "bad frame range");
      VIEO_ARG(sbp[f].n_levels == fr[f].n_levels, "pyramid depth mismatch");
      if (sbp[f].n_kp > VIEO_SBP_MAX_KEYPOINTS) {
        set_error("vieo_search_local_points: frame %d has %d keypoints (max %d)", f, sbp[f].n_kp, VIEO_SBP_MAX_KEYPOINTS);
        return VIEO_E_CAPACITY;
      }
      nk = std::max(nk, (size_t)sbp[f].kp_begin + sbp[f].n_kp);
    }
  }
  VIEO_ARG(nq == 0 || (wP && normal && max_dist && min_dist && inview), "null point array");
  // the tracking info is optional for the fused call (Tracking only needs btrack_inview_ and the matches)
  const bool want_info = proj || level || viewcos || depth;
  VIEO_ARG(sbp || nq == 0 || (proj && level && viewcos && depth), "null point array");
  VIEO_ARG(!want_info || nq == 0 || (proj && level && viewcos && depth), "tracking-info outputs must be given together");
  if (sbp) {
    VIEO_ARG(n_matches, "null argument");
    VIEO_ARG(nk == 0 || (kps && uright && desc && kp_match), "null keypoint array");
    VIEO_ARG(nq == 0 || (q_desc && q_flags && q_match && q_dist), "null query array");
  }
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  VIEO_ARG(cs, "no call scratch");
  const size_t nq1 = std::max<size_t>(nq, 1), nk1 = std::max<size_t>(nk, 1), nf = (size_t)n_frames;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  // inputs
  const size_t o_ff = take(sizeof(VieoFrustumFrame) * nf), o_wp = take(12 * nq1), o_pn = take(12 * nq1), o_mx = take(4 * nq1),
               o_mn = take(4 * nq1), o_sk = take(nq1);
  const size_t o_sf = take(sbp ? sizeof(VieoSbpFrame) * nf : 0), o_kp = take(sbp ? sizeof(VieoKeyPoint) * nk1 : 0),
               o_ur = take(sbp ? 4 * nk1 : 0), o_de = take(sbp ? 32 * nk1 : 0), o_bl = take(sbp ? nk1 : 0),
               o_qd = take(sbp ? 32 * nq1 : 0), o_fl = take(sbp ? nq1 : 0);
  const size_t in_bytes = off;
  // outputs: [inview | n_inview | matches] always come back, the tracking info behind them only when asked for
  const size_t o_iv = take(nq1), o_ni = take(4 * nf);
  const size_t o_km = take(sbp ? 4 * nk1 : 0), o_qm = take(sbp ? 4 * nq1 : 0), o_qs = take(sbp ? 4 * nq1 : 0),
               o_nm = take(sbp ? 4 * nf : 0);
  const size_t o_pr = take(12 * nq1), o_lv = take(4 * nq1), o_vc = take(4 * nq1), o_dp = take(4 * nq1);
  const size_t io_bytes = off, back_bytes = want_info ? io_bytes : o_pr;
  const size_t sc_bytes = sbp ? vieo_sbp_scratch_bytes((int)nq1) : 16;
  uint8_t* dbuf = (uint8_t*)cs->get(0, io_bytes);
  void* dsc = cs->get(1, sc_bytes);
  uint8_t* hbuf = (uint8_t*)cs->get_pinned(io_bytes);
  VIEO_ARG(dbuf && dsc && hbuf, "staging allocation failed");
  auto put = [&](size_t o, const void* src, size_t bytes) { if (src && bytes) memcpy(hbuf + o, src, bytes); };
  put(o_ff, fr.data(), sizeof(VieoFrustumFrame) * nf);
  put(o_wp, wP, 12 * nq); put(o_pn, normal, 12 * nq); put(o_mx, max_dist, 4 * nq); put(o_mn, min_dist, 4 * nq);
  put(o_sk, skip, nq);
  if (sbp) {
    put(o_sf, sbp, sizeof(VieoSbpFrame) * nf);
    put(o_kp, kps, sizeof(VieoKeyPoint) * nk); put(o_ur, uright, 4 * nk); put(o_de, desc, 32 * nk); put(o_bl, kp_blocked, nk);
    put(o_qd, q_desc, 32 * nq); put(o_fl, q_flags, nq);
  }
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf, in_bytes, cudaMemcpyHostToDevice, cs->st);
  // points outside every frame's range read back as "not in view" (level -1) / unmatched
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_iv, 0, o_ni - o_iv, cs->st);
  if (e == cudaSuccess && sbp) e = cudaMemsetAsync(dbuf + o_km, 0xff, o_nm - o_km, cs->st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_pr, 0, o_lv - o_pr, cs->st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_lv, 0xff, o_vc - o_lv, cs->st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_vc, 0, io_bytes - o_vc, cs->st);
  if (e == cudaSuccess) {
    rc = vieo_frustum_batch_dev((const VieoFrustumFrame*)(dbuf + o_ff), n_frames, (const float*)(dbuf + o_wp),
                                (const float*)(dbuf + o_pn), (const float*)(dbuf + o_mx), (const float*)(dbuf + o_mn),
                                skip ? dbuf + o_sk : nullptr, dbuf + o_iv, (float*)(dbuf + o_pr), (int32_t*)(dbuf + o_lv),
                                (float*)(dbuf + o_vc), (float*)(dbuf + o_dp), (int32_t*)(dbuf + o_ni), cs->st);
    if (rc == VIEO_OK && sbp) {
      VieoSbpQueries dq{};
      dq.level = (const int32_t*)(dbuf + o_lv); dq.proj = (const float*)(dbuf + o_pr);
      dq.viewcos = (const float*)(dbuf + o_vc); dq.depth = (const float*)(dbuf + o_dp);
      dq.desc = dbuf + o_qd; dq.flags = dbuf + o_fl;
      rc = vieo_sbp_batch_dev(VIEO_SBP_LOCAL_MAP, (const VieoSbpFrame*)(dbuf + o_sf), n_frames, (const VieoKeyPoint*)(dbuf + o_kp),
                              (const float*)(dbuf + o_ur), dbuf + o_de, &dq, kp_blocked ? dbuf + o_bl : nullptr,
                              (int32_t*)(dbuf + o_km), (int32_t*)(dbuf + o_qm), (int32_t*)(dbuf + o_qs),
                              (int32_t*)(dbuf + o_nm), dsc, sc_bytes, cs->st);
    }
    if (rc == VIEO_OK) {
      e = cudaMemcpyAsync(hbuf + o_iv, dbuf + o_iv, back_bytes - o_iv, cudaMemcpyDeviceToHost, cs->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(cs->st);
    }
  }
  if (e != cudaSuccess) {
    set_error("vieo_frustum: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  if (rc != VIEO_OK) return rc;
  if (nq) {
    memcpy(inview, hbuf + o_iv, nq);
    if (want_info) {
      memcpy(proj, hbuf + o_pr, 12 * nq); memcpy(level, hbuf + o_lv, 4 * nq);
      memcpy(viewcos, hbuf + o_vc, 4 * nq); memcpy(depth, hbuf + o_dp, 4 * nq);
    }
  }
  memcpy(n_inview, hbuf + o_ni, 4 * nf);
  if (sbp) {
    if (nk) memcpy(kp_match, hbuf + o_km, 4 * nk);
    if (nq) { memcpy(q_match, hbuf + o_qm, 4 * nq); memcpy(q_dist, hbuf + o_qs, 4 * nq); }
    memcpy(n_matches, hbuf + o_nm, 4 * nf);
  }
  return VIEO_OK;
}

int vieo_frustum_rig_batch_dev(const VieoFrustumRigFrame* frames_dev, int n_frames, const float* wP_dev, const float* normal_dev,
                               const float* max_dist_dev, const float* min_dist_dev, const uint8_t* skip_dev, uint8_t* inview_dev,
                               uint8_t* cam_mask_dev, float* proj_dev, int32_t* level_dev, float* viewcos_dev, float* depth_dev,
                               int32_t* n_inview_dev, void* stream) {
  VIEO_ARG(n_frames >= 0 && n_frames <= 65535, "bad argument (at most 65535 frames per call)");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames_dev && wP_dev && normal_dev && max_dist_dev && min_dist_dev && inview_dev && cam_mask_dev && proj_dev &&
               level_dev && viewcos_dev && depth_dev && n_inview_dev, "null argument");
  VIEO_CK(cudaMemsetAsync(n_inview_dev, 0, 4 * (size_t)n_frames, (cudaStream_t)stream));
  k_frustum_rig<<<dim3(kFrBlocksPerFrame, n_frames), kFrThreads, 0, (cudaStream_t)stream>>>(
      frames_dev, wP_dev, normal_dev, max_dist_dev, min_dist_dev, skip_dev, inview_dev, cam_mask_dev, proj_dev, level_dev,
      viewcos_dev, depth_dev, n_inview_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_frustum_rig_batch(const VieoFrustumRigFrame* frames, int n_frames, const float* wP, const float* normal,
                           const float* max_dist, const float* min_dist, const uint8_t* skip, uint8_t* inview, uint8_t* cam_mask,
                           float* proj, int32_t* level, float* viewcos, float* depth, int32_t* n_inview, int device) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames && n_inview, "null argument");
  size_t nq = 0;
  std::vector<VieoFrustumRigFrame> fr(frames, frames + n_frames);
  for (int f = 0; f < n_frames; ++f) {
    VIEO_ARG(fr[f].q_begin >= 0 && fr[f].n_q >= 0, "bad frame range");
    VIEO_ARG(fr[f].n_cams >= 1 && fr[f].n_cams <= 4, "a rig has 1 to 4 cameras");
    for (int c = 0; c < fr[f].n_cams; ++c) VIEO_ARG(fr[f].cam[c].model >= 0 && fr[f].cam[c].model <= 2, "camera model must be 0, 1 or 2");
    int rc = vieo_frustum_level_table(fr[f].log_scale_factor, fr[f].n_levels, fr[f].level_ratio);
    if (rc) return rc;
    nq = std::max(nq, (size_t)fr[f].q_begin + fr[f].n_q);
  }
  VIEO_ARG(nq == 0 || (wP && normal && max_dist && min_dist && inview && cam_mask && proj && level && viewcos && depth), "null point array");
  int rc = use_device(device);
  if (rc) return rc;
  CallScratch* cs = call_scratch(device);
  VIEO_ARG(cs, "no call scratch");
  const size_t nq1 = std::max<size_t>(nq, 1), nf = (size_t)n_frames;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  const size_t o_ff = take(sizeof(VieoFrustumRigFrame) * nf), o_wp = take(12 * nq1), o_pn = take(12 * nq1), o_mx = take(4 * nq1),
               o_mn = take(4 * nq1), o_sk = take(nq1);
  const size_t in_bytes = off;
  const size_t o_iv = take(nq1), o_cm = take(nq1), o_ni = take(4 * nf), o_dp = take(4 * nq1), o_vc = take(16 * nq1), o_pr = take(48 * nq1),
               o_lv = take(16 * nq1);
  const size_t io_bytes = off;
  uint8_t* dbuf = (uint8_t*)cs->get(0, io_bytes);
  uint8_t* hbuf = (uint8_t*)cs->get_pinned(io_bytes);
  VIEO_ARG(dbuf && hbuf, "staging allocation failed");
  auto put = [&](size_t o, const void* src, size_t bytes) { if (src && bytes) memcpy(hbuf + o, src, bytes); };
  put(o_ff, fr.data(), sizeof(VieoFrustumRigFrame) * nf);
  put(o_wp, wP, 12 * nq); put(o_pn, normal, 12 * nq); put(o_mx, max_dist, 4 * nq); put(o_mn, min_dist, 4 * nq);
  put(o_sk, skip, nq);
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf, in_bytes, cudaMemcpyHostToDevice, cs->st);
  // points outside every frame's range read back as "not in view" (level -1)
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_iv, 0, o_lv - o_iv, cs->st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbuf + o_lv, 0xff, io_bytes - o_lv, cs->st);
  if (e == cudaSuccess) {
    rc = vieo_frustum_rig_batch_dev((const VieoFrustumRigFrame*)(dbuf + o_ff), n_frames, (const float*)(dbuf + o_wp),
                                    (const float*)(dbuf + o_pn), (const float*)(dbuf + o_mx), (const float*)(dbuf + o_mn),
                                    skip ? dbuf + o_sk : nullptr, dbuf + o_iv, dbuf + o_cm, (float*)(dbuf + o_pr),
                                    (int32_t*)(dbuf + o_lv), (float*)(dbuf + o_vc), (float*)(dbuf + o_dp), (int32_t*)(dbuf + o_ni),
                                    cs->st);
    if (rc == VIEO_OK) {
      e = cudaMemcpyAsync(hbuf + o_iv, dbuf + o_iv, io_bytes - o_iv, cudaMemcpyDeviceToHost, cs->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(cs->st);
    }
  }
  if (e != cudaSuccess) {
    set_error("vieo_frustum_rig: %s", cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  if (rc != VIEO_OK) return rc;
  if (nq) {
    memcpy(inview, hbuf + o_iv, nq); memcpy(cam_mask, hbuf + o_cm, nq); memcpy(depth, hbuf + o_dp, 4 * nq);
    memcpy(viewcos, hbuf + o_vc, 16 * nq); memcpy(proj, hbuf + o_pr, 48 * nq); memcpy(level, hbuf + o_lv, 16 * nq);
  }
  memcpy(n_inview, hbuf + o_ni, 4 * nf);
  return VIEO_OK;
}

int vieo_frustum_batch(const VieoFrustumFrame* frames, int n_frames, const float* wP, const float* normal,
                       const float* max_dist, const float* min_dist, const uint8_t* skip, uint8_t* inview, float* proj,
                       int32_t* level, float* viewcos, float* depth, int32_t* n_inview, int device) {
  return frustum_host(frames, nullptr, n_frames, wP, normal, max_dist, min_dist, skip, nullptr, nullptr, nullptr, nullptr,
                      nullptr, nullptr, inview, proj, level, viewcos, depth, n_inview, nullptr, nullptr, nullptr, nullptr,
                      device);
}

int vieo_search_local_points(const VieoFrustumFrame* frustum, const VieoSbpFrame* frames, int n_frames, const float* wP,
                             const float* normal, const float* max_dist, const float* min_dist, const uint8_t* skip,
                             const uint8_t* q_desc, const uint8_t* q_flags, const VieoKeyPoint* kps, const float* uright,
                             const uint8_t* desc, const uint8_t* kp_blocked, uint8_t* inview, float* proj, int32_t* level,
                             float* viewcos, float* depth, int32_t* n_inview, int32_t* kp_match, int32_t* q_match,
                             int32_t* q_dist, int32_t* n_matches, int device) {
  VIEO_ARG(n_frames >= 0, "bad argument");
  if (n_frames == 0) return VIEO_OK;
  VIEO_ARG(frames, "null argument");
  return frustum_host(frustum, frames, n_frames, wP, normal, max_dist, min_dist, skip, kps, uright, desc, q_desc, q_flags,
                      kp_blocked, inview, proj, level, viewcos, depth, n_inview, kp_match, q_match, q_dist, n_matches, device);
}

}  // extern "C"

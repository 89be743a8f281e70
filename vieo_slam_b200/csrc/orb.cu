// ORB front-end on sm_100a: chained fixed-point pyramid, per-cell FAST-9/16 + NMS + threshold fallback,
// data-parallel quadtree distribution, intensity-centroid orientation and Gaussian-blurred steered
// rBRIEF — everything on the device, one fixed sequence of launches per batch of images.
//
// Reference behaviour reproduced (leavesnight/VIEO_SLAM @356e4a22): src/ORBextractor.cc:391-456 (tables),
// :1060-1081 (pyramid, cv::resize INTER_LINEAR 8U fixed-point), :723-802 (cells, cv::FAST, fallback),
// :467-721 (DistributeOctTree), :55-80 (IC_Angle, cv::fastAtan2), :1012-1024 + :83-127 (GaussianBlur 7x7
// s=2 + computeOrbDescriptor), :968-1058 (operator()).  This is a new design, not a translation: see DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "common.cuh"

namespace vieo {

constexpr int kMaxLevels = 16;
constexpr int kBorder = 16;      // EDGE_THRESHOLD - 3
constexpr int kHalfPatch = 15;
constexpr int kFastThreads = 256;
constexpr int kQtThreads = 512;
constexpr int kQtKnodeSmem = 12288;  // keys whose node index lives in shared memory (u16 each); more keys: the HBM scratch
constexpr int kDescWarps = 4;
constexpr int kPR = 21;           // raw patch radius: 18 (rotated pattern reach) + 3 (7-tap blur)
constexpr int kPW = 2 * kPR + 1;  // 43
constexpr int kPWp = 52;          // row of the staged raw patch in bytes: 13 words, the patch starts at byte xoff (0..3) of it
constexpr int kBW = 37;           // blurred window (+-18)
constexpr int kBWp = 40;          // 20 words: four outputs = two aligned word stores
constexpr int kProfRing = 1024;    // profiled calls kept between vieo_orb_profile_read()s

struct Cell {
  int16_t level, x0, y0, cw, ch;  // cell image = [x0,x0+cw) x [y0,y0+ch) in level coordinates
};

struct OrbParams {  // passed by value to kernels
  int nlevels;
  int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
  size_t img_stride[kMaxLevels];  // bytes between images of a level (levels >= 1; level 0 may be external)
  uint8_t* lvl[kMaxLevels];       // device base per level (lvl[0] = internal staging copy)
  float scale[kMaxLevels];
  float kp_size[kMaxLevels];
  int quota[kMaxLevels];
  int cell_begin[kMaxLevels + 1];
  int n_cells;
  int cell_cap;                // slots per cell
  int key_begin[kMaxLevels + 1];  // per-image offsets into the dense key arrays (cell_begin * cell_cap)
  int maxn[kMaxLevels];        // node capacity per level
  int slot_begin[kMaxLevels + 1];  // per-image offsets into lvl_kp
  int tile_w, tile_h;          // max cell dims (tile_w padded to 4)
  int tile_pitch;              // row pitch of the shared-memory cell tile (64 when the tile is TMA-staged)
  int ini_th, min_th;
  const Cell* cells;
  const int4* rx[kMaxLevels];  // per dst x: {x0, x1, a0, a1}
  const int4* ry[kMaxLevels];  // per dst y: {y0, y1, b0, b1}
  uint32_t* cand;              // [img][n_cells][cell_cap] packed x | y<<12 | response<<24
  int* cell_cnt;               // [img][n_cells]
  uint32_t* keys;              // [img][key_begin[nlevels]]
  uint16_t* knode;             // same shape
  int qt_knode_off;            // byte offset of the shared-memory node-index array in k_quadtree's dynamic region
  uint32_t* lvl_kp;            // [img][slot_begin[nlevels]] packed keypoints after the quadtree (list order)
  int* lvl_cnt;                // [img][nlevels]
};

__constant__ int c_umax[kHalfPatch + 1];
__device__ const int8_t d_pattern[1024] = {
#include "../../data/orb_pattern_31.inc"
};

// ------------------------------------------------------------------------------------------------
// Pyramid: level l from level l-1, all images of the batch in one launch.  4 pixels per thread.
// The whole pyramid of one image in ONE launch: level l is made from level l - 1 of the SAME image, so a CTA per image
// walks the levels with a block barrier in between (its own writes, through L1 / L2, are visible to its own threads after
// __syncthreads; plain loads, not the non-coherent path).  Same arithmetic as k_resize.  Seven dependent launches become
// one: under load (tracking + LocalBA streams on the device) every launch boundary of the chain costs a scheduling round.
struct PyrArgs {
  const uint8_t* img0;
  size_t img0_stride;
  int img0_pitch, nlevels;
  uint8_t* lvl[kMaxLevels];
  size_t stride[kMaxLevels];
  int pitch[kMaxLevels], w[kMaxLevels], h[kMaxLevels];
  const int4* rx[kMaxLevels];
  const int4* ry[kMaxLevels];
};
// 32 registers: two CTAs per SM, so a batch of 256 images is ONE wave on 148 SMs (1.73 waves at one CTA per SM)
__global__ void __launch_bounds__(1024, 2) k_pyramid(PyrArgs A) {
  const int img = blockIdx.x;
  for (int l = 1; l < A.nlevels; ++l) {
    const uint8_t* s = l == 1 ? A.img0 + (size_t)img * A.img0_stride : A.lvl[l - 1] + (size_t)img * A.stride[l - 1];
    const int sp = l == 1 ? A.img0_pitch : A.pitch[l - 1];
    uint8_t* d = A.lvl[l] + (size_t)img * A.stride[l];
    const int dw = A.w[l], dh = A.h[l], dp = A.pitch[l], nx4 = (dw + 3) >> 2;
    const int4* rx = A.rx[l];
    const int4* ry = A.ry[l];
    for (int i = threadIdx.x; i < nx4 * dh; i += 1024) {
      const int y = i / nx4, x4 = (i - y * nx4) << 2;
      const int4 fy = __ldg(ry + y);
      const uint8_t* r0 = s + (size_t)fy.x * sp;
      const uint8_t* r1 = s + (size_t)fy.y * sp;
      uint32_t out = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x < dw) {
          const int4 fx = __ldg(rx + x);
          const int h0 = r0[fx.x] * fx.z + r0[fx.y] * fx.w;
          const int h1 = r1[fx.x] * fx.z + r1[fx.y] * fx.w;
          const int v = (((fy.z * (h0 >> 4)) >> 16) + ((fy.w * (h1 >> 4)) >> 16) + 2) >> 2;
          out |= (uint32_t)(v & 0xff) << (8 * k);
        }
      }
      *reinterpret_cast<uint32_t*>(d + (size_t)y * dp + x4) = out;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_resize(const uint8_t* __restrict__ src, size_t src_img_stride, int src_pitch,
                                                uint8_t* __restrict__ dst, size_t dst_img_stride, int dst_pitch,
                                                int dw, int dh, const int4* __restrict__ rx,
                                                const int4* __restrict__ ry) {
  const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x4 >= dw || y >= dh) return;
  const uint8_t* s = src + blockIdx.z * src_img_stride;
  const int4 fy = __ldg(ry + y);
  const uint8_t* r0 = s + (size_t)fy.x * src_pitch;
  const uint8_t* r1 = s + (size_t)fy.y * src_pitch;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x4 + i;
    if (x < dw) {
      const int4 fx = __ldg(rx + x);
      const int h0 = r0[fx.x] * fx.z + r0[fx.y] * fx.w;
      const int h1 = r1[fx.x] * fx.z + r1[fx.y] * fx.w;
      const int v = (((fy.z * (h0 >> 4)) >> 16) + ((fy.w * (h1 >> 4)) >> 16) + 2) >> 2;
      out |= (uint32_t)(v & 0xff) << (8 * i);
    }
  }
  *reinterpret_cast<uint32_t*>(dst + blockIdx.z * dst_img_stride + (size_t)y * dst_pitch + x4) = out;
}

// ------------------------------------------------------------------------------------------------
// FAST-9/16 arc score: S = max(v - min_arcs(max9 p), max_arcs(min9 p) - v, 0) over the 16 contiguous 9-arcs of the
// radius-3 ring; equals cornerScore<16>()+1 for corners, and a pixel is a corner at threshold t iff S > t.
// Two horizontally adjacent pixels per thread, one per 16-bit lane: sm_100a has native packed min/max only for
// 16-bit lanes (VIMNMX.U16x2 and the 3-input VIMNMX3.U16x2; the u8x4 video intrinsics are emulated with ~10 LOP3/PRMT
// each).  Ring pixel k of the two centres = two adjacent bytes of the shared-memory tile, cut out of three aligned
// words per row with funnel shifts + one PRMT.  Sliding-window tree with 3-input ops: windows of 3, then 3+3+3 -> 9:
// 64 VIMNMX3 for the 16 arcs of both pixels + 16 for the final min-of-max / max-of-min.
__device__ __forceinline__ unsigned fast_score2(const uint8_t* rowc, int o8, int pitch) {
  // rowc: aligned word that holds the first centre; o8 = 8 * (byte offset of that centre in the word) (0, 8, 16 or 24)
#define VIEO_ROW3(dy)                                                          \
  {                                                                            \
    const unsigned* w_ = reinterpret_cast<const unsigned*>(rowc + (dy) * pitch); \
    const unsigned wm_ = w_[-1], w0_ = w_[0], wp_ = w_[1];                      \
    vm = __funnelshift_r(wm_, w0_, o8);                                         \
    v0 = __funnelshift_r(w0_, wp_, o8);                                         \
    vp = wp_ >> o8;                                                             \
  }
  // bytes (dx, dx + 1) relative to the first centre -> 16-bit lanes
#define VIEO_PICK(dx) \
  (__byte_perm((dx) < 0 ? vm : v0, (dx) < 0 ? v0 : vp, ((dx) < 0 ? 4 + (dx) : (dx)) | ((((dx) < 0 ? 4 + (dx) : (dx)) + 1) << 8)) & 0x00ff00ffu)
  unsigned R[16], vm, v0, vp;
  VIEO_ROW3(3);
  R[15] = VIEO_PICK(-1); R[0] = VIEO_PICK(0); R[1] = VIEO_PICK(1);
  VIEO_ROW3(2);
  R[14] = VIEO_PICK(-2); R[2] = VIEO_PICK(2);
  VIEO_ROW3(1);
  R[13] = VIEO_PICK(-3); R[3] = VIEO_PICK(3);
  VIEO_ROW3(0);
  const unsigned v = VIEO_PICK(0);
  R[12] = VIEO_PICK(-3); R[4] = VIEO_PICK(3);
  VIEO_ROW3(-1);
  R[11] = VIEO_PICK(-3); R[5] = VIEO_PICK(3);
  VIEO_ROW3(-2);
  R[10] = VIEO_PICK(-2); R[6] = VIEO_PICK(2);
  VIEO_ROW3(-3);
  R[9] = VIEO_PICK(-1); R[8] = VIEO_PICK(0); R[7] = VIEO_PICK(1);
#undef VIEO_ROW3
#undef VIEO_PICK
  unsigned lo3[16], hi3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo3[k] = __vimin3_u16x2(R[k], R[(k + 1) & 15], R[(k + 2) & 15]);
    hi3[k] = __vimax3_u16x2(R[k], R[(k + 1) & 15], R[(k + 2) & 15]);
  }
  unsigned lo9[16], hi9[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo9[k] = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
    hi9[k] = __vimax3_u16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
  }
  unsigned a = __vimin3_u16x2(hi9[0], hi9[1], hi9[2]), b = __vimax3_u16x2(lo9[0], lo9[1], lo9[2]);
#pragma unroll
  for (int k = 3; k < 15; k += 2) {
    a = __vimin3_u16x2(a, hi9[k], hi9[k + 1]);
    b = __vimax3_u16x2(b, lo9[k], lo9[k + 1]);
  }
  a = __vminu2(a, hi9[15]);
  b = __vmaxu2(b, lo9[15]);
  // max(v - a, b - v, 0) per 16-bit lane; lanes hold values <= 255, so plain 32-bit arithmetic cannot borrow across
  // lanes once the differences are clamped with max
  const unsigned va = __vmaxu2(v, a) - a, bv = __vmaxu2(b, v) - v;  // max(v,a)-a = max(v-a,0)
  return __vmaxu2(va, bv);
}

// TMA staging of the cell tile: one tensor map per pyramid level over (x, y, image) with the level's 64-byte-multiple
// pitch; a single thread issues one cp.async.bulk.tensor of a tile_pitch x tile_h box that lands in shared memory and
// signals an mbarrier, while the other threads already clear the score tile.  The unit faults ("illegal instruction") when
// the innermost coordinate is not a multiple of 16 bytes (measured on B200: tools/tma_probe2.cu), so the box starts at
// x0 - 1 rounded down to 16 and the scorer reads the cell at the residual column offset.  Out-of-image parts of the
// box are zero-filled by the unit.
struct FastTmaps {
  CUtensorMap m[kMaxLevels];
};
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_tile3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z,
                                                uint32_t bytes) {
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)map), "r"(b), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  const uint32_t b = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VIEO_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VIEO_DONE_%=;\n"
      "bra VIEO_WAIT_%=;\n"
      "VIEO_DONE_%=:\n"
      "}\n" ::"r"(b),
      "r"(phase)
      : "memory");
}

// One CTA per (cell, image): stage the cell image in shared memory, score every interior pixel (2 per thread), 3x3
// strict NMS restricted to the cell interior (each cell is an independent cv::FAST call in the reference), decide
// iniTh vs minTh from the post-NMS count, and write the survivors in raster order.
// Tile layouts: raw row pitch TP = tile_w + 8, cell column x at byte 1 + x (interior column 3 -> aligned byte 4);
// score row pitch TP, interior pixel (x, y) at row y + 1, byte 4 + x, with a zero ring around the interior.
template <bool kTma>
__global__ void __launch_bounds__(kFastThreads) k_fast_cells(OrbParams P, const uint8_t* __restrict__ img0,
                                                            size_t img0_stride, int img0_pitch,
                                                            const __grid_constant__ FastTmaps tmaps) {
  // no static shared memory in this kernel: the dynamic region then starts at offset 0 of the CTA's window, which gives
  // the TMA destination (the raw tile, first in the region) its 128-byte alignment
  extern __shared__ __align__(128) uint8_t fsmem[];
  uint8_t* smem = fsmem;
  const Cell c = P.cells[blockIdx.x];
  const int img = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t* base;
  int pitch;
  if (c.level == 0) {
    base = img0 + img * img0_stride;
    pitch = img0_pitch;
  } else {
    base = P.lvl[c.level] + img * P.img_stride[c.level];
    pitch = P.pitch[c.level];
  }
  const int TP = P.tile_pitch;
  uint8_t* raw = smem;
  const int iw = c.cw - 6, ih = c.ch - 6;
  uint8_t* sc = smem + TP * P.tile_h;  // (ih + 2) rows
  uint8_t* tail = smem + ((2 * TP * P.tile_h + 15) & ~15);  // after the raw and score tiles
  uint64_t& s_bar = *reinterpret_cast<uint64_t*>(tail);
  int* s_warp = reinterpret_cast<int*>(tail + 16);

  int* cnt_out = P.cell_cnt + (size_t)img * P.n_cells + blockIdx.x;
  if (iw <= 0 || ih <= 0) {
    if (tid == 0) *cnt_out = 0;
    return;
  }
  // tile byte (row y, column b) = level pixel (xs + b, y0 + y); cell column x sits at byte 1 + xoff + x
  const int xs = kTma ? ((c.x0 - 1) & ~15) : c.x0 - 1, xoff = c.x0 - 1 - xs;
  if (kTma) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tma_load_tile3d(raw, &tmaps.m[c.level], &s_bar, xs, c.y0, img, (uint32_t)(TP * P.tile_h));
    }
    for (int i = tid; i < (ih + 2) * (TP / 4); i += kFastThreads) reinterpret_cast<unsigned*>(sc)[i] = 0;
    __syncthreads();        // the barrier is initialised (thread 0 passed it) before anybody polls it
    mbar_wait(&s_bar, 0);
  } else {
    for (int y = wid; y < c.ch; y += kFastThreads / 32) {
      const uint8_t* src = base + (size_t)(c.y0 + y) * pitch + c.x0;
      for (int x = lane; x < c.cw; x += 32) raw[y * TP + 1 + x] = __ldg(src + x);
    }
    for (int i = tid; i < (ih + 2) * (TP / 4); i += kFastThreads) reinterpret_cast<unsigned*>(sc)[i] = 0;
    __syncthreads();
  }
  const int ng = (iw + 1) >> 1;  // pairs of pixels per row
  for (int i = tid; i < ng * ih; i += kFastThreads) {
    const int y = i / ng, g = i - y * ng;
    const int col = 4 + 2 * g;  // byte column of the first centre in the score row (and in the raw row when off == 0)
    const int rc = col + xoff;
    const unsigned s2 = fast_score2(raw + (y + 3) * TP + (rc & ~3), 8 * (rc & 3), TP);
    unsigned s0 = s2 & 0xffffu, s1 = s2 >> 16;
    s0 = s0 > (unsigned)P.min_th ? s0 : 0u;  // keep scores > minTh
    s1 = (s1 > (unsigned)P.min_th && 2 * g + 1 < iw) ? s1 : 0u;
    *reinterpret_cast<uint16_t*>(sc + (y + 1) * TP + col) = (uint16_t)(s0 | (s1 << 8));
  }
  __syncthreads();
  // contiguous raster chunk per thread -> ordered compaction
  const int npx = iw * ih;
  const int per = (npx + kFastThreads - 1) / kFastThreads;  // <= 32 (checked at create)
  const int beg = tid * per, end = min(beg + per, npx);
  uint32_t mask_a = 0, mask_b = 0;
  if (beg < end) {
    int y = beg / iw, x = beg - y * iw;
    for (int i = beg; i < end; ++i) {
      const uint8_t* q = sc + (y + 1) * TP + 4 + x;
      const int s = q[0];
      if (s != 0) {
        const int m = max(max(max(q[-TP - 1], q[-TP]), max(q[-TP + 1], q[-1])),
                          max(max(q[1], q[TP - 1]), max(q[TP], q[TP + 1])));
        if (s > m) {
          mask_b |= 1u << (i - beg);
          if (s > P.ini_th) mask_a |= 1u << (i - beg);
        }
      }
      if (++x == iw) {
        x = 0;
        ++y;
      }
    }
  }
  const int total_a = __syncthreads_count(mask_a != 0);
  const uint32_t mask = total_a > 0 ? mask_a : mask_b;
  const int cnt = __popc(mask);
  const int inc = warp_incl_scan(cnt, lane);
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int off = inc - cnt;
  int total = 0;
#pragma unroll
  for (int w2 = 0; w2 < kFastThreads / 32; ++w2) {
    if (w2 < wid) off += s_warp[w2];
    total += s_warp[w2];
  }
  uint32_t* out = P.cand + ((size_t)img * P.n_cells + blockIdx.x) * P.cell_cap;
  uint32_t m = mask;
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    const int i = beg + b;
    const int y = i / iw, x = i - y * iw;
    const int s = sc[(y + 1) * TP + 4 + x];
    out[off++] = (uint32_t)(c.x0 + 3 + x) | ((uint32_t)(c.y0 + 3 + y) << 12) | ((uint32_t)(s - 1) << 24);
  }
  if (tid == 0) *cnt_out = total;
}

// ------------------------------------------------------------------------------------------------
// Quadtree distribution, one CTA per (level, image).  The reference's std::list algorithm
// (src/ORBextractor.cc:518-721) restated as rounds of "count children per key" + prefix sums: the node
// table is kept in LIST ORDER; a round replaces the processed nodes by their non-empty children
// (pushed to the front in reverse creation order) and keeps the others behind them.  Keys never move:
// each key carries the list position of its node, and the final pick is an atomicMax of
// (response, first-in-order).  tools/quadtree_parallel_proto.py validates the formulation on CPU.
struct QtSmem {
  int16_t *x0[2], *x1[2], *y0[2], *y1[2];
  int *cnt[2], *id[2];
  uint8_t* flag[2];
  int *cc, *cpos, *prank, *keptpos, *ord, *ta, *tb;
};

__device__ __forceinline__ int qt_quad(const QtSmem& S, int cur, int p, int kx, int ky) {
  const int mx = S.x0[cur][p] + ((S.x1[cur][p] - S.x0[cur][p] + 1) >> 1);
  const int my = S.y0[cur][p] + ((S.y1[cur][p] - S.y0[cur][p] + 1) >> 1);
  return (kx >= mx ? 1 : 0) + (ky >= my ? 2 : 0);
}

__global__ void __launch_bounds__(kQtThreads) k_quadtree(OrbParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int level = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
  const int MAXN = P.maxn[level];
  const int N = P.quota[level];
  QtSmem S;
  {
    uint8_t* p = smem;
    auto take = [&](size_t bytes) {
      uint8_t* r = p;
      p += (bytes + 15) & ~size_t(15);
      return r;
    };
    for (int b = 0; b < 2; ++b) {
      S.cnt[b] = (int*)take(4 * MAXN);
      S.id[b] = (int*)take(4 * MAXN);
      S.x0[b] = (int16_t*)take(2 * MAXN);
      S.x1[b] = (int16_t*)take(2 * MAXN);
      S.y0[b] = (int16_t*)take(2 * MAXN);
      S.y1[b] = (int16_t*)take(2 * MAXN);
      S.flag[b] = take(MAXN);
    }
    S.cc = (int*)take(16 * MAXN);
    S.cpos = (int*)take(16 * MAXN);
    S.prank = (int*)take(4 * MAXN);
    S.keptpos = (int*)take(4 * MAXN);
    S.ord = (int*)take(4 * MAXN);
    S.ta = (int*)take(4 * MAXN);
    S.tb = (int*)take(4 * MAXN);
  }
  // the node index of every key is rewritten in every split round: in shared memory when the level's keys fit (they do for
  // every level of a 752 x 480 image), instead of a read-modify-write of HBM scratch per round (209 MB written per 256 images)
  uint16_t* s_knode = reinterpret_cast<uint16_t*>(smem + P.qt_knode_off);
  __shared__ int s_total, s_total2, s_nk, s_cut;
  __shared__ int s_celloff[512];  // per-level cell count <= 512 (checked at create)

  const int ncell = P.cell_begin[level + 1] - P.cell_begin[level];
  const int* ccnt = P.cell_cnt + (size_t)img * P.n_cells + P.cell_begin[level];
  uint32_t* keys = P.keys + (size_t)img * P.key_begin[P.nlevels] + P.key_begin[level];
  uint16_t* knode = P.knode + (size_t)img * P.key_begin[P.nlevels] + P.key_begin[level];  // re-pointed once nk is known
  uint32_t* out_kp = P.lvl_kp + (size_t)img * P.slot_begin[P.nlevels] + P.slot_begin[level];
  int* out_cnt = P.lvl_cnt + (size_t)img * P.nlevels + level;

  // ---- gather the cells' survivors into one dense, ordered key array
  for (int i = tid; i < ncell; i += kQtThreads) s_celloff[i] = ccnt[i];
  __syncthreads();
  warp0_excl_scan(s_celloff, ncell, &s_nk);
  __syncthreads();
  const int nk = s_nk;
  if (nk == 0) {
    if (tid == 0) *out_cnt = 0;
    return;
  }
  if (nk <= kQtKnodeSmem) knode = s_knode;
  {
    const uint32_t* cand = P.cand + ((size_t)img * P.n_cells + P.cell_begin[level]) * P.cell_cap;
    const int wid = tid >> 5, lane = tid & 31;
    for (int cidx = wid; cidx < ncell; cidx += kQtThreads / 32) {
      const int n = ccnt[cidx], o = s_celloff[cidx];
      for (int i = lane; i < n; i += 32) keys[o + i] = cand[(size_t)cidx * P.cell_cap + i];
    }
  }
  // ---- root nodes (src/ORBextractor.cc:523-566)
  const int W = P.w[level], H = P.h[level];
  const int spanX = (W - kBorder) - kBorder, spanY = (H - kBorder) - kBorder;
  const int nIni = (int)roundf((float)spanX / (float)spanY);
  const float hX = (float)spanX / (float)nIni;
  for (int i = tid; i < MAXN; i += kQtThreads) S.ta[i] = 0;
  __syncthreads();
  for (int k = tid; k < nk; k += kQtThreads) {
    const uint32_t key = keys[k];
    const int r = (int)((float)((int)(key & 0xfff) - kBorder) / hX);
    knode[k] = (uint16_t)r;
    atomicAdd(&S.ta[r], 1);
  }
  __syncthreads();
  for (int i = tid; i < nIni; i += kQtThreads) {
    S.tb[i] = S.ta[i] > 0 ? 1 : 0;
  }
  __syncthreads();
  warp0_excl_scan(S.tb, nIni, &s_total);
  __syncthreads();
  int n = s_total;
  int cur = 0;
  for (int i = tid; i < nIni; i += kQtThreads) {
    if (S.ta[i] > 0) {
      const int np = S.tb[i];
      S.x0[0][np] = (int16_t)(int)(hX * (float)i);
      S.x1[0][np] = (int16_t)(int)(hX * (float)(i + 1));
      S.y0[0][np] = 0;
      S.y1[0][np] = (int16_t)spanY;
      S.cnt[0][np] = S.ta[i];
      S.id[0][np] = i;
      S.flag[0][np] = 0;
    }
    S.keptpos[i] = S.tb[i];
  }
  __syncthreads();
  for (int k = tid; k < nk; k += kQtThreads) knode[k] = (uint16_t)S.keptpos[knode[k]];
  int next_id = nIni;
  bool phase_b = false;
  __syncthreads();

  for (int round = 0; round < 4096; ++round) {  // n grows every round or the loop ends; the cap is a safety net
    const int prev = n;
    int nproc;
    // ---- choose the nodes to split this round and their processing order
    if (!phase_b) {  // every node with >1 keys, in list order (:602-641)
      for (int p = tid; p < n; p += kQtThreads) S.ta[p] = S.cnt[cur][p] > 1 ? 1 : 0;
      __syncthreads();
      warp0_excl_scan(S.ta, n, &s_total);
      __syncthreads();
      nproc = s_total;
      for (int p = tid; p < n; p += kQtThreads) {
        if (S.cnt[cur][p] > 1) {
          S.prank[p] = S.ta[p];
          S.ord[S.ta[p]] = p;
        } else
          S.prank[p] = -1;
      }
    } else {  // the children created last round with >1 keys, largest first, ties: latest created first (:644-652)
      for (int p = tid; p < n; p += kQtThreads) S.ta[p] = S.flag[cur][p] ? 1 : 0;
      __syncthreads();
      warp0_excl_scan(S.ta, n, &s_total);
      __syncthreads();
      nproc = s_total;
      for (int p = tid; p < n; p += kQtThreads) {
        if (S.flag[cur][p]) S.tb[S.ta[p]] = p;  // compact list of expandable positions
      }
      __syncthreads();
      for (int e = tid; e < nproc; e += kQtThreads) {
        const int p = S.tb[e];
        const int c = S.cnt[cur][p], idp = S.id[cur][p];
        int r = 0;
        for (int e2 = 0; e2 < nproc; ++e2) {
          const int p2 = S.tb[e2];
          const int c2 = S.cnt[cur][p2], id2 = S.id[cur][p2];
          r += (c2 > c || (c2 == c && id2 > idp)) ? 1 : 0;
        }
        S.ord[r] = p;
      }
      __syncthreads();
      for (int p = tid; p < n; p += kQtThreads) S.prank[p] = -1;
      __syncthreads();
      for (int r = tid; r < nproc; r += kQtThreads) S.prank[S.ord[r]] = r;
    }
    for (int i = tid; i < 4 * n; i += kQtThreads) S.cc[i] = 0;
    __syncthreads();
    if (nproc == 0) break;  // nothing split: list size unchanged -> finish (:664, :716)
    // ---- children sizes
    for (int k = tid; k < nk; k += kQtThreads) {
      const int p = knode[k];
      if (S.prank[p] >= 0) {
        const uint32_t key = keys[k];
        const int q = qt_quad(S, cur, p, (int)(key & 0xfff) - kBorder, (int)((key >> 12) & 0xfff) - kBorder);
        atomicAdd(&S.cc[4 * p + q], 1);
      }
    }
    __syncthreads();
    if (phase_b) {  // stop splitting as soon as the list reaches N nodes (:713)
      for (int r = tid; r < nproc; r += kQtThreads) {
        const int p = S.ord[r];
        S.ta[r] = (S.cc[4 * p] > 0) + (S.cc[4 * p + 1] > 0) + (S.cc[4 * p + 2] > 0) + (S.cc[4 * p + 3] > 0) - 1;
      }
      if (tid == 0) s_cut = nproc;
      __syncthreads();
      warp0_excl_scan(S.ta, nproc, &s_total);
      __syncthreads();
      for (int r = tid; r < nproc; r += kQtThreads) {
        const int p = S.ord[r];
        const int d = (S.cc[4 * p] > 0) + (S.cc[4 * p + 1] > 0) + (S.cc[4 * p + 2] > 0) + (S.cc[4 * p + 3] > 0) - 1;
        const int after = n + S.ta[r] + d;
        const int before = n + S.ta[r];
        if (after >= N && before < N) s_cut = r + 1;  // unique: `after` is non-decreasing in r
      }
      __syncthreads();
      const int cut = s_cut;
      for (int r = cut + tid; r < nproc; r += kQtThreads) S.prank[S.ord[r]] = -1;
      nproc = cut;
      __syncthreads();
    }
    // ---- layout of the new list
    for (int r = tid; r < nproc; r += kQtThreads) {
      const int p = S.ord[r];
      S.ta[r] = (S.cc[4 * p] > 0) + (S.cc[4 * p + 1] > 0) + (S.cc[4 * p + 2] > 0) + (S.cc[4 * p + 3] > 0);
    }
    for (int p = tid; p < n; p += kQtThreads) S.tb[p] = S.prank[p] < 0 ? 1 : 0;
    __syncthreads();
    warp0_excl_scan(S.ta, nproc, &s_total);
    __syncthreads();
    warp0_excl_scan(S.tb, n, &s_total2);
    __syncthreads();
    const int C = s_total, K = s_total2;
    const int nxt = cur ^ 1;
    for (int p = tid; p < n; p += kQtThreads) {
      const int r = S.prank[p];
      if (r < 0) {
        const int np = C + S.tb[p];
        S.x0[nxt][np] = S.x0[cur][p];
        S.x1[nxt][np] = S.x1[cur][p];
        S.y0[nxt][np] = S.y0[cur][p];
        S.y1[nxt][np] = S.y1[cur][p];
        S.cnt[nxt][np] = S.cnt[cur][p];
        S.id[nxt][np] = S.id[cur][p];
        S.flag[nxt][np] = 0;
        S.keptpos[p] = np;
      } else {
        int c = S.ta[r];
        const int ax0 = S.x0[cur][p], ax1 = S.x1[cur][p], ay0 = S.y0[cur][p], ay1 = S.y1[cur][p];
        const int mx = ax0 + ((ax1 - ax0 + 1) >> 1), my = ay0 + ((ay1 - ay0 + 1) >> 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cq = S.cc[4 * p + q];
          if (cq == 0) continue;
          const int np = C - 1 - c;
          S.x0[nxt][np] = (int16_t)((q & 1) ? mx : ax0);
          S.x1[nxt][np] = (int16_t)((q & 1) ? ax1 : mx);
          S.y0[nxt][np] = (int16_t)((q & 2) ? my : ay0);
          S.y1[nxt][np] = (int16_t)((q & 2) ? ay1 : my);
          S.cnt[nxt][np] = cq;
          S.id[nxt][np] = next_id + c;
          S.flag[nxt][np] = cq > 1 ? 1 : 0;
          S.cpos[4 * p + q] = np;
          ++c;
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < nk; k += kQtThreads) {
      const int p = knode[k];
      if (S.prank[p] >= 0) {
        const uint32_t key = keys[k];
        const int q = qt_quad(S, cur, p, (int)(key & 0xfff) - kBorder, (int)((key >> 12) & 0xfff) - kBorder);
        knode[k] = (uint16_t)S.cpos[4 * p + q];
      } else {
        knode[k] = (uint16_t)S.keptpos[p];
      }
    }
    n = C + K;
    next_id += C;
    cur = nxt;
    __syncthreads();
    if (n >= N || n == prev) break;
    if (!phase_b) {  // would splitting every expandable child overshoot N?  then go largest-first (:669)
      if (tid == 0) s_total = 0;
      __syncthreads();
      int mine = 0;
      for (int p = tid; p < n; p += kQtThreads) mine += S.flag[cur][p];
      if (mine) atomicAdd(&s_total, mine);
      __syncthreads();
      if (n + 3 * s_total > N) phase_b = true;
      __syncthreads();
    }
  }
  // ---- per node: highest response, first in candidate order (:702-718)
  for (int p = tid; p < n; p += kQtThreads) S.ta[p] = 0;
  __syncthreads();
  for (int k = tid; k < nk; k += kQtThreads) {
    const uint32_t key = keys[k];
    atomicMax((unsigned int*)&S.ta[knode[k]], ((key >> 24) << 20) | (uint32_t)(0xFFFFF - k));
  }
  __syncthreads();
  for (int p = tid; p < n; p += kQtThreads) {
    const int k = 0xFFFFF - (int)((uint32_t)S.ta[p] & 0xFFFFF);
    out_kp[p] = keys[k];
  }
  if (tid == 0) *out_cnt = n;
}

// ------------------------------------------------------------------------------------------------
// cv::fastAtan2 scalar path, fp32 with explicit rounding (the library is built with -fmad=false).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
              p7 = -0.04432655554792128f * s;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + 2.2204460492503131e-16f);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + 2.2204460492503131e-16f);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// glibc 2.39 sinf/cosf algorithm (double polynomial after pi/2 reduction) — the reference's
// (float)cos(angle) / (float)sin(angle) at src/ORBextractor.cc:84-85 resolve to these.
__device__ __forceinline__ float sc_poly(double x, double x2, int n, bool neg_cos) {
  if ((n & 1) == 0) {
    const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
    const double x3 = x * x2, s1 = s2c + x2 * s3c, x7 = x3 * x2, s = x + x3 * s1c;
    return (float)(s + x7 * s1);
  }
  const double sg = neg_cos ? -1.0 : 1.0;
  const double c0 = sg * 0x1p0, c1c = sg * -0x1.ffffffd0c621cp-2, c2c = sg * 0x1.55553e1068f19p-5,
               c3c = sg * -0x1.6c087e89a359dp-10, c4c = sg * 0x1.99343027bf8c3p-16;
  const double x4 = x2 * x2, c2 = c3c + x2 * c4c, c1 = c0 + x2 * c1c, x6 = x4 * x2, c = c1 + x4 * c2c;
  return (float)(c + x6 * c2);
}
__device__ __forceinline__ void sincos_ref(float y, float* sn, float* cs) {
  double x = (double)y;
  const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
  if (top < ((__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ff)) {
    const double x2 = x * x;
    if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ff)) {
      *sn = y;
      *cs = 1.0f;
      return;
    }
    *sn = sc_poly(x, x2, 0, false);
    *cs = sc_poly(x, x2, 1, false);
    return;
  }
  const double r = x * 0x1.45F306DC9C883p+23;
  const int n = ((int)r + 0x800000) >> 24;
  x = x - (double)n * 0x1.921FB54442D18p0;
  const int q = n & 3;
  const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
  const bool neg = (n & 2) != 0;
  *sn = sc_poly(x * s, x * x, n, neg);
  *cs = sc_poly(x * s, x * x, n ^ 1, neg);
}

__device__ __forceinline__ int reflect101(int p, int n) {
  if (p < 0) p = -p;
  if (p >= n) p = 2 * n - 2 - p;
  return p;
}

// One warp per keypoint: stage the 43x43 raw patch (REFLECT_101 at the image edge, as the blur of the
// cloned level does), intensity-centroid angle on the raw pixels, separable 7-tap fixed-point blur
// (exact integer accumulation, one final rounding — OpenCV's 8U GaussianBlur path) evaluated only where
// the 512 steered samples land, then the 256 binary tests.  Also assembles the level-ordered output.
__global__ void __launch_bounds__(kDescWarps * 32) k_orient_desc(OrbParams P, const uint8_t* __restrict__ img0,
                                                                size_t img0_stride, int img0_pitch,
                                                                VieoKeyPoint* __restrict__ kps,
                                                                uint8_t* __restrict__ desc, int cap,
                                                                int* __restrict__ n_kp) {
  __shared__ __align__(16) uint8_t s_raw[kDescWarps][kPW * kPWp];
  __shared__ __align__(16) uint16_t s_hb[kDescWarps][kPW * kBWp];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.y;
  const int slot = blockIdx.x * kDescWarps + warp;
  const int* lc = P.lvl_cnt + (size_t)img * P.nlevels;
  int level = -1, j = 0, total = 0;
  for (int l = 0; l < P.nlevels; ++l) {
    const int c = lc[l];
    if (level < 0 && slot < total + c) {
      level = l;
      j = slot - total;
    }
    total += c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) n_kp[img] = min(total, cap);
  if (level < 0 || slot >= cap) return;
  const uint32_t key = P.lvl_kp[(size_t)img * P.slot_begin[P.nlevels] + P.slot_begin[level] + j];
  const int kx = key & 0xfff, ky = (key >> 12) & 0xfff;
  const uint8_t* base;
  int pitch;
  if (level == 0) {
    base = img0 + img * img0_stride;
    pitch = img0_pitch;
  } else {
    base = P.lvl[level] + img * P.img_stride[level];
    pitch = P.pitch[level];
  }
  const int W = P.w[level], H = P.h[level];
  uint8_t* raw = s_raw[warp];
  uint16_t* hb = s_hb[warp];
  // The kernel is bound by the shared-memory pipe (one instruction per clock per SM), so the patch moves as 32-bit words:
  // interior keypoints of a 4-byte-pitched image copy each row's 12 aligned words (the patch then starts at byte xoff of
  // the staged row); keypoints near the image edge (REFLECT_101, only there) and odd pitches take the byte path.
  int xoff = 0;
  const uint8_t* first = base + (size_t)(ky - kPR) * pitch + (kx - kPR);
  if ((pitch & 3) == 0 && kx >= kPR + 3 && kx + kPR + 3 < W && ky >= kPR && ky + kPR < H) {
    xoff = (int)((uintptr_t)first & 3);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(first - xoff);
    const int pw = pitch >> 2;
    uint32_t* raw32 = reinterpret_cast<uint32_t*>(raw);
    for (int it = lane; it < kPW * 12; it += 32) {
      const int r = it / 12, w = it - 12 * r;
      raw32[r * (kPWp / 4) + w] = __ldg(src + (size_t)r * pw + w);
    }
  } else if (kx >= kPR && kx + kPR < W && ky >= kPR && ky + kPR < H) {
#pragma unroll 4
    for (int r = 0; r < kPW; ++r) {
      raw[r * kPWp + lane] = __ldg(first + (size_t)r * pitch + lane);
      if (lane + 32 < kPW) raw[r * kPWp + lane + 32] = __ldg(first + (size_t)r * pitch + lane + 32);
    }
  } else {
    const int gx0 = reflect101(kx - kPR + lane, W), gx1 = reflect101(kx - kPR + min(lane + 32, kPW - 1), W);
    for (int r = 0; r < kPW; ++r) {
      const uint8_t* src = base + (size_t)reflect101(ky - kPR + r, H) * pitch;
      raw[r * kPWp + lane] = __ldg(src + gx0);
      if (lane + 32 < kPW) raw[r * kPWp + lane + 32] = __ldg(src + gx1);
    }
  }
  __syncwarp();
  // intensity centroid over the radius-15 disc: lane = row v in [-15, 15]
  int m10 = 0, m01 = 0;
  if (lane < 2 * kHalfPatch + 1) {
    const int v = lane - kHalfPatch;
    const int d = c_umax[v < 0 ? -v : v];
    const uint8_t* row = raw + (kPR + v) * kPWp + xoff + kPR;
    int rs = 0;
    for (int u = -d; u <= d; ++u) {
      const int val = row[u];
      m10 += u * val;
      rs += val;
    }
    m01 = v * rs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  // Horizontal pass, 8-fractional-bit kernel {18,34,48,56,48,34,18} (sum 256): exact in 16 bits.  A work item is four
  // consecutive outputs of one row: four aligned word loads, realigned by xoff with funnel shifts, then the 7 taps on PACKED
  // pairs — (b[k], b[k+2]) as two 16-bit halves of one register, so one integer multiply-add serves two outputs (each half
  // stays below 65536: no carry crosses) — and two aligned word stores.  Same integer sums as the byte form: bit-exact.
  {
    const uint32_t* raw32 = reinterpret_cast<const uint32_t*>(raw);
    uint32_t* hb32 = reinterpret_cast<uint32_t*>(hb);
    const int sh = 8 * xoff;
    for (int it = lane; it < kPW * 10; it += 32) {
      const int r = it / 10, g = it - 10 * r;
      const uint32_t* w = raw32 + r * (kPWp / 4) + g;
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
      const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
      const uint32_t E0 = v0 & 0x00ff00ffu, O0 = (v0 >> 8) & 0x00ff00ffu;  // (b0, b2), (b1, b3)
      const uint32_t E1 = v1 & 0x00ff00ffu, O1 = (v1 >> 8) & 0x00ff00ffu;  // (b4, b6), (b5, b7)
      const uint32_t E2 = v2 & 0x00ff00ffu, O2 = (v2 >> 8) & 0x00ff00ffu;  // (b8, b10), (b9, b11)
      const uint32_t P2 = __byte_perm(E0, E1, 0x5432), P3 = __byte_perm(O0, O1, 0x5432);  // (b2, b4), (b3, b5)
      const uint32_t P6 = __byte_perm(E1, E2, 0x5432), P7 = __byte_perm(O1, O2, 0x5432);  // (b6, b8), (b7, b9)
      // (o0, o2): taps P0..P6 = E0, O0, P2, P3, E1, O1, P6;  (o1, o3): taps P1..P7
      const uint32_t A = 18u * (E0 + P6) + 34u * (O0 + O1) + 48u * (P2 + E1) + 56u * P3;
      const uint32_t B = 18u * (O0 + P7) + 34u * (P2 + P6) + 48u * (P3 + O1) + 56u * E1;
      hb32[r * (kBWp / 2) + 2 * g] = __byte_perm(A, B, 0x5410);      // (o0, o1)
      hb32[r * (kBWp / 2) + 2 * g + 1] = __byte_perm(A, B, 0x7632);  // (o2, o3)
    }
  }
  __syncwarp();
  float sn, cs;
  sincos_ref(angle * (float)(3.14159265358979323846 / 180.f), &sn, &cs);
  const int8_t* pat = d_pattern + lane * 32;
  int byte = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    int v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float px = (float)pat[4 * t + 2 * e], py = (float)pat[4 * t + 2 * e + 1];
      const int ry = __float2int_rn(px * sn + py * cs);
      const int rx = __float2int_rn(px * cs - py * sn);
      const uint16_t* h = hb + (ry + 18) * kBWp + rx + 18;
      const uint32_t acc = 18u * (h[0] + h[6 * kBWp]) + 34u * (h[kBWp] + h[5 * kBWp]) +
                           48u * (h[2 * kBWp] + h[4 * kBWp]) + 56u * h[3 * kBWp];
      v[e] = (int)((acc + 32768u) >> 16);
    }
    byte |= (v[0] < v[1] ? 1 : 0) << t;
  }
  desc[((size_t)img * cap + slot) * 32 + lane] = (uint8_t)byte;
  if (lane == 0) {
    VieoKeyPoint k;
    const float sc = P.scale[level];
    k.x = level ? (float)kx * sc : (float)kx;
    k.y = level ? (float)ky * sc : (float)ky;
    k.size = P.kp_size[level];
    k.angle = angle;
    k.response = (float)(key >> 24);
    k.octave = level;
    kps[(size_t)img * cap + slot] = k;
  }
}

// ------------------------------------------------------------------------------------------------
// Rectified-stereo association, Frame::ComputeStereoMatches (src/Frame.cc:451-611), two kernels.
// k_stereo_match, grid (kStereoSplit, frames): a CTA takes one slice of the frame's left keypoints (left image 2f, right
// image 2f+1 of the batch) after staging the right keypoints' search keys (x, row band, octave: 12 B each) in shared
// memory.  A warp takes one left keypoint: lanes scan the right keypoints in index order (row-band / octave /
// disparity-range predicate, 256-bit Hamming), the lexicographic (distance, index) minimum is the reference's strict-'<'
// arg-min; matches under (TH_HIGH+TH_LOW)/2 are refined with the 11x11 L1 block search over +-5 px in the keypoint's
// pyramid level (integer arithmetic, lanes over pixels) and the parabola fit.
// k_stereo_filter, one CTA per frame — the only per-frame step: the 1.5 * 1.4 * median(SAD) filter with a rank selection
// over the accepted matches.
constexpr int kStereoSplit = 8;
struct StereoLevels {
  const uint8_t* base[kMaxLevels];
  size_t img_stride[kMaxLevels];
  int pitch[kMaxLevels], w[kMaxLevels];
  float scale[kMaxLevels], inv_scale[kMaxLevels];
};
__global__ void __launch_bounds__(256) k_stereo_match(StereoLevels Lv, const VieoKeyPoint* __restrict__ kps,
                                                      const uint8_t* __restrict__ desc, const int* __restrict__ nkp, int cap,
                                                      int n_rows, float bf, float minZ, float* __restrict__ uright,
                                                      float* __restrict__ depth, int* __restrict__ sad_out) {
  extern __shared__ float s_rx[];                      // [cap] right keypoint x
  int* s_band = reinterpret_cast<int*>(s_rx + cap);    // [cap] row band (minr << 16) | (maxr & 0xffff), both int16
  int8_t* s_oct = reinterpret_cast<int8_t*>(s_band + cap);  // [cap]
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iml = 2 * f, imr = 2 * f + 1;
  const int nl = min(nkp[iml], cap), nr = min(nkp[imr], cap);
  const VieoKeyPoint* KL = kps + (size_t)iml * cap;
  const VieoKeyPoint* KR = kps + (size_t)imr * cap;
  const uint8_t* DL = desc + (size_t)iml * cap * 32;
  const uint8_t* DR = desc + (size_t)imr * cap * 32;
  float* UR = uright + (size_t)f * cap;
  float* DP = depth + (size_t)f * cap;
  int* SD = sad_out + (size_t)f * cap;
  const float minD = 0, maxD = bf / minZ;
  // this CTA's slice of the left keypoints; the last slice also clears the unused tail of the output rows
  const int per = (nl + (int)gridDim.x - 1) / (int)gridDim.x;
  const int l0 = min((int)blockIdx.x * per, nl), l1 = min(l0 + per, nl);
  const int c1 = blockIdx.x == gridDim.x - 1 ? cap : l1;
  for (int i = l0 + threadIdx.x; i < c1; i += 256) {
    UR[i] = -1.0f;
    DP[i] = -1.0f;
    SD[i] = -1;
  }
  for (int i = threadIdx.x; i < nr; i += 256) {
    const VieoKeyPoint k = KR[i];
    const float r = 2.0f * Lv.scale[k.octave];
    const int maxr = (int)ceilf(k.y + r), minr = (int)floorf(k.y - r);
    s_rx[i] = k.x;
    s_band[i] = (minr << 16) | (maxr & 0xffff);
    s_oct[i] = (int8_t)k.octave;
  }
  __syncthreads();
  for (int iL = l0 + warp; iL < l1; iL += 8) {
    const VieoKeyPoint kpL = KL[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    const float minU = uL - maxD, maxU = uL - minD;
    if (row < 0 || row >= n_rows || maxU < 0) continue;
    uint32_t dl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dl[k] = reinterpret_cast<const uint32_t*>(DL + 32 * (size_t)iL)[k];
    int best = 100, bestIdx = 0x7fffffff;  // TH_HIGH
    for (int iR = lane; iR < nr; iR += 32) {
      const int oct = s_oct[iR];
      if (oct < levelL - 1 || oct > levelL + 1) continue;
      const int band = s_band[iR];
      const int minr = band >> 16, maxr = (int)(short)(band & 0xffff);
      if (row < minr || row > maxr) continue;
      const float xr = s_rx[iR];
      if (!(xr >= minU && xr <= maxU)) continue;
      const uint4* dr = reinterpret_cast<const uint4*>(DR + 32 * (size_t)iR);
      const uint4 r0 = __ldg(dr), r1 = __ldg(dr + 1);
      const int d = __popc(dl[0] ^ r0.x) + __popc(dl[1] ^ r0.y) + __popc(dl[2] ^ r0.z) + __popc(dl[3] ^ r0.w) +
                    __popc(dl[4] ^ r1.x) + __popc(dl[5] ^ r1.y) + __popc(dl[6] ^ r1.z) + __popc(dl[7] ^ r1.w);
      if (d < best) {  // lanes see ascending indices: strict '<' keeps the first
        best = d;
        bestIdx = iR;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
      if (ob < best || (ob == best && oi < bestIdx)) {
        best = ob;
        bestIdx = oi;
      }
    }
    if (!(best < 75)) continue;  // thOrbDist = (TH_HIGH + TH_LOW) / 2
    const float uR0 = s_rx[bestIdx];
    const float sf = Lv.inv_scale[levelL];
    const float scaleduL = roundf(kpL.x * sf), scaledvL = roundf(kpL.y * sf), scaleduR0 = roundf(uR0 * sf);
    const int w = 5, L = 5;
    const int W = Lv.w[levelL];
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= W) continue;
    const int pitch = Lv.pitch[levelL];
    const uint8_t* IL = Lv.base[levelL] + (size_t)iml * Lv.img_stride[levelL];
    const uint8_t* IR = Lv.base[levelL] + (size_t)imr * Lv.img_stride[levelL];
    const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
    const int cL = IL[(size_t)cv * pitch + cu];
    // the lane's pixels of the 11x11 left patch, centre subtracted
    int a[4], py[4], px[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = lane + 32 * q;
      py[q] = t / 11 - w;
      px[q] = t % 11 - w;
      a[q] = t < 121 ? (int)IL[(size_t)(cv + py[q]) * pitch + cu + px[q]] - cL : 0;
    }
    int bestSad = 0x7fffffff, bestinc = 0, dprev = 0, d1 = 0, d2 = 0, d3 = 0;
    for (int inc = -L; inc <= L; ++inc) {
      const int cR = IR[(size_t)cv * pitch + cr + inc];
      int sum = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < 121) sum += abs(a[q] - ((int)IR[(size_t)(cv + py[q]) * pitch + cr + inc + px[q]] - cR));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (sum < bestSad) {
        bestSad = sum;
        bestinc = inc;
        d1 = dprev;
        d2 = sum;
        d3 = -1;  // filled by the next shift
      } else if (inc == bestinc + 1)
        d3 = sum;
      dprev = sum;
    }
    if (bestinc == -L || bestinc == L) continue;
    if (lane == 0) {
      const float dist1 = (float)d1, dist2 = (float)d2, dist3 = (float)d3;
      const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
      if (!(deltaR < -1 || deltaR > 1)) {
        float bestuR = Lv.scale[levelL] * ((float)scaleduR0 + (float)bestinc + deltaR);
        float disparity = uL - bestuR;
        if (disparity >= minD && disparity < maxD) {
          if (disparity <= 0) {
            disparity = 0.01f;
            bestuR = uL - 0.01f;
          }
          DP[iL] = bf / disparity;
          UR[iL] = bestuR;
          SD[iL] = bestSad;  // >= 0 exactly for the accepted matches (k_stereo_filter's input)
        }
      }
    }
  }
}

// median of the accepted SADs = element n/2 of the sorted (sad, index) list: the value v with
// #(sad < v) <= n/2 < #(sad <= v); matches at or above 1.5 * 1.4 * median lose uright / depth (src/Frame.cc:599-610)
__global__ void __launch_bounds__(256) k_stereo_filter(const int* __restrict__ nkp, int cap, float* __restrict__ uright,
                                                       float* __restrict__ depth, const int* __restrict__ sad_out) {
  extern __shared__ int s_sad[];  // [cap]
  __shared__ int s_med;
  __shared__ int s_cnt[8];
  const int f = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nl = min(nkp[2 * f], cap);
  float* UR = uright + (size_t)f * cap;
  float* DP = depth + (size_t)f * cap;
  const int* SD = sad_out + (size_t)f * cap;
  int n_acc = 0;
  for (int i = threadIdx.x; i < nl; i += 256) {
    const int v = SD[i];
    s_sad[i] = v;
    n_acc += v >= 0;
  }
  if (threadIdx.x == 0) s_med = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
  if (lane == 0) s_cnt[warp] = n_acc;
  __syncthreads();
  int total = 0;
  for (int k = 0; k < 8; ++k) total += s_cnt[k];
  if (total == 0) return;
  const int rank = total / 2;
  for (int i = threadIdx.x; i < nl; i += 256) {
    const int v = s_sad[i];
    if (v < 0) continue;
    int less = 0, leq = 0;
    for (int j = 0; j < nl; ++j) {
      const int u = s_sad[j];
      if (u < 0) continue;
      less += u < v;
      leq += u <= v;
    }
    if (less <= rank && rank < leq) s_med = v;  // every thread that qualifies writes the same value
  }
  __syncthreads();
  const float thDist = 1.5f * 1.4f * (float)s_med;
  for (int i = threadIdx.x; i < nl; i += 256) {
    const int v = s_sad[i];
    if (v >= 0 && !((float)v < thDist)) {
      UR[i] = -1.0f;
      DP[i] = -1.0f;
    }
  }
}

}  // namespace vieo

// =================================================================================================
using namespace vieo;

struct vieo_orb {
  VieoOrbConfig cfg;
  int device;
  OrbParams P;
  cudaStream_t stream;
  float inv_scale[kMaxLevels], sigma2[kMaxLevels], inv_sigma2[kMaxLevels];
  int cap_total;  // sum of maxn
  size_t fast_smem, qt_smem;
  // TMA staging of the FAST cell tiles: tensor maps of the internal levels (encoded once) and of the level-0 source of
  // the current call (re-encoded when it changes)
  bool tma_ok;
  FastTmaps tmaps;
  const uint8_t* tm0_ptr;
  size_t tm0_stride;
  int tm0_pitch, tm0_n;
  std::vector<void*> allocs;
  // staging for the host API
  VieoKeyPoint* d_kps;
  uint8_t* d_desc;
  int* d_nkp;
  void* h_pinned;  // pinned bounce for small D2H (counts)
  int last_launches;
  // optional per-stage CUDA-event timing (bench.py's roofline): ring of 5 events per call
  bool prof_on;
  int prof_calls;
  std::vector<cudaEvent_t> prof_ev;
  int last_n_img;
  const uint8_t* last_img0;  // level-0 source of the last call (debug)
  size_t last_img0_stride;
  int last_img0_pitch;
};

namespace {

inline int cv_roundf(float v) { return (int)nearbyintf(v); }

template <class T>
int dev_alloc(vieo_orb* h, T** p, size_t count) {
  void* q = nullptr;
  VIEO_CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = (T*)q;
  return VIEO_OK;
}

template <class T>
int dev_upload(vieo_orb* h, const T** p, const std::vector<T>& v) {
  T* q;
  int rc = dev_alloc(h, &q, v.size());
  if (rc) return rc;
  VIEO_CK(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *p = q;
  return VIEO_OK;
}


// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency): u8 tensor
// (x: pitch bytes, y: rows, z: images), box tile_pitch x tile_h x 1, no swizzle, zero fill outside
bool encode_level_map(CUtensorMap* map, const uint8_t* base, int pitch, int rows, size_t img_stride, int n_img, int box_w,
                      int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeFn)f;
  }();
  if (!fn || pitch % 16 || img_stride % 16 || (uintptr_t)base % 16) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)std::max(n_img, 1)};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)img_stride};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int orb_run(vieo_orb* h, int n_img, const uint8_t* img0, size_t img0_stride, int img0_pitch, VieoKeyPoint* kps,
            uint8_t* desc, int cap, int* n_kp, cudaStream_t st) {
  const OrbParams& P = h->P;
  int launches = 0;
  cudaEvent_t* ev = nullptr;
  if (h->prof_on && h->prof_calls < kProfRing) {
    if (h->prof_ev.empty()) {
      h->prof_ev.resize((size_t)kProfRing * 5);
      for (auto& e : h->prof_ev) VIEO_CK(cudaEventCreate(&e));
    }
    ev = h->prof_ev.data() + (size_t)5 * h->prof_calls++;
    VIEO_CK(cudaEventRecord(ev[0], st));
  }
  static const bool pyr_multi = getenv("VIEO_ORB_PYR_MULTI") != nullptr;  // the seven-launch form, for comparison
  if (!pyr_multi && P.nlevels > 1) {
    PyrArgs A;
    A.img0 = img0; A.img0_stride = img0_stride; A.img0_pitch = img0_pitch; A.nlevels = P.nlevels;
    for (int l = 0; l < P.nlevels; ++l) {
      A.lvl[l] = P.lvl[l]; A.stride[l] = P.img_stride[l]; A.pitch[l] = P.pitch[l]; A.w[l] = P.w[l]; A.h[l] = P.h[l];
      A.rx[l] = P.rx[l]; A.ry[l] = P.ry[l];
    }
    k_pyramid<<<n_img, 1024, 0, st>>>(A);
    ++launches;
  } else
  for (int l = 1; l < P.nlevels; ++l) {
    const uint8_t* src = l == 1 ? img0 : P.lvl[l - 1];
    const size_t sstride = l == 1 ? img0_stride : P.img_stride[l - 1];
    const int spitch = l == 1 ? img0_pitch : P.pitch[l - 1];
    dim3 grid((P.w[l] + 127) / 128, (P.h[l] + 7) / 8, n_img), block(32, 8);
    k_resize<<<grid, block, 0, st>>>(src, sstride, spitch, P.lvl[l], P.img_stride[l], P.pitch[l], P.w[l], P.h[l],
                                     P.rx[l], P.ry[l]);
    ++launches;
  }
  if (ev) VIEO_CK(cudaEventRecord(ev[1], st));
  bool tma = h->tma_ok && ((uintptr_t)img0 % 16 == 0) && img0_pitch % 16 == 0 && img0_stride % 16 == 0;
  if (tma && (h->tm0_ptr != img0 || h->tm0_stride != img0_stride || h->tm0_pitch != img0_pitch || h->tm0_n < n_img)) {
    tma = encode_level_map(&h->tmaps.m[0], img0, img0_pitch, P.h[0], img0_stride, n_img, P.tile_pitch, P.tile_h);
    h->tm0_ptr = tma ? img0 : nullptr; h->tm0_stride = img0_stride; h->tm0_pitch = img0_pitch; h->tm0_n = n_img;
  }
  if (tma)
    k_fast_cells<true><<<dim3(P.n_cells, n_img), kFastThreads, h->fast_smem, st>>>(P, img0, img0_stride, img0_pitch, h->tmaps);
  else
    k_fast_cells<false><<<dim3(P.n_cells, n_img), kFastThreads, h->fast_smem, st>>>(P, img0, img0_stride, img0_pitch, h->tmaps);
  if (ev) VIEO_CK(cudaEventRecord(ev[2], st));
  k_quadtree<<<dim3(P.nlevels, n_img), kQtThreads, h->qt_smem, st>>>(P);
  if (ev) VIEO_CK(cudaEventRecord(ev[3], st));
  const int slots = std::min(cap, h->cap_total);
  k_orient_desc<<<dim3((slots + kDescWarps - 1) / kDescWarps, n_img), kDescWarps * 32, 0, st>>>(
      P, img0, img0_stride, img0_pitch, kps, desc, cap, n_kp);
  if (ev) VIEO_CK(cudaEventRecord(ev[4], st));
  launches += 3;
  VIEO_CK(cudaGetLastError());
  h->last_launches = launches;
  h->last_n_img = n_img;
  h->last_img0 = img0;
  h->last_img0_stride = img0_stride;
  h->last_img0_pitch = img0_pitch;
  return VIEO_OK;
}

}  // namespace

// ---- internal hooks used by frontend.cu (same library) ----
namespace vieo {
int orb_enqueue_host(vieo_orb* h, int n_img, const uint8_t* imgs, size_t img_stride, int row_stride) {
  const OrbParams& P = h->P;
  cudaStream_t st = h->stream;
  // Keep the caller's layout on the device whenever it fits the staging buffer: ONE contiguous H2D copy runs at PCIe
  // speed, a 2-D copy of 752-byte rows into a padded pitch does not (measured 6x slower on B200 hosts).
  if (row_stride >= h->cfg.width && row_stride <= P.pitch[0] && img_stride >= (size_t)row_stride * h->cfg.height &&
      img_stride <= P.img_stride[0]) {
    VIEO_CK(cudaMemcpyAsync(P.lvl[0], imgs, img_stride * (n_img - 1) + (size_t)row_stride * h->cfg.height,
                            cudaMemcpyHostToDevice, st));
    return orb_run(h, n_img, P.lvl[0], img_stride, row_stride, h->d_kps, h->d_desc, h->cap_total, h->d_nkp, st);
  }
  for (int i = 0; i < n_img; ++i)
    VIEO_CK(cudaMemcpy2DAsync(P.lvl[0] + i * P.img_stride[0], P.pitch[0], imgs + i * img_stride, row_stride,
                              h->cfg.width, h->cfg.height, cudaMemcpyHostToDevice, st));
  return orb_run(h, n_img, P.lvl[0], P.img_stride[0], P.pitch[0], h->d_kps, h->d_desc, h->cap_total, h->d_nkp, st);
}
void orb_info(vieo_orb* h, int* device, int* max_batch) {
  *device = h->device;
  *max_batch = h->cfg.max_batch;
}
void orb_dev_outputs(vieo_orb* h, VieoKeyPoint** kps, uint8_t** desc, int** nkp, int* cap, cudaStream_t* st) {
  *kps = h->d_kps;
  *desc = h->d_desc;
  *nkp = h->d_nkp;
  *cap = h->cap_total;
  *st = h->stream;
}
}  // namespace vieo

extern "C" {

int vieo_orb_create(const VieoOrbConfig* cfg, int device, vieo_orb_t** out) {
  VIEO_ARG(cfg && out, "null argument");
  VIEO_ARG(cfg->nlevels >= 1 && cfg->nlevels <= kMaxLevels, "nlevels must be in [1,16]");
  VIEO_ARG(cfg->width >= 64 && cfg->height >= 64 && cfg->width < 4096 && cfg->height < 4096,
           "image size must be in [64,4096)");
  VIEO_ARG(cfg->nfeatures > 0 && cfg->scale_factor > 1.0f && cfg->max_batch >= 1, "bad nfeatures/scale/max_batch");
  VIEO_ARG(cfg->min_th_fast >= 1 && cfg->ini_th_fast >= cfg->min_th_fast && cfg->ini_th_fast < 255,
           "need 1 <= minThFAST <= iniThFAST < 255");
  int rc = use_device(device);
  if (rc) return rc;
  vieo_orb* h = new vieo_orb();
  h->cfg = *cfg;
  h->device = device;
  OrbParams& P = h->P;
  memset(&P, 0, sizeof(P));
  const int L = cfg->nlevels;
  P.nlevels = L;
  P.ini_th = cfg->ini_th_fast;
  P.min_th = cfg->min_th_fast;
  // scale tables and quotas: same arithmetic as src/ORBextractor.cc:397-431 (scaleFactor is a double member)
  const double sf = (double)cfg->scale_factor;
  P.scale[0] = 1.f;
  h->sigma2[0] = 1.f;
  for (int i = 1; i < L; ++i) {
    P.scale[i] = (float)(P.scale[i - 1] * sf);
    h->sigma2[i] = P.scale[i] * P.scale[i];
  }
  for (int i = 0; i < L; ++i) {
    h->inv_scale[i] = 1.0f / P.scale[i];
    h->inv_sigma2[i] = 1.0f / h->sigma2[i];
    P.kp_size[i] = (float)(int)(31 * P.scale[i]);
  }
  {
    const float factor = (float)(1.0f / sf);
    float per = (float)(cfg->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L)));
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {
      P.quota[l] = cv_roundf(per);
      sum += P.quota[l];
      per *= factor;
    }
    P.quota[L - 1] = std::max(cfg->nfeatures - sum, 0);
  }
  int umax[kHalfPatch + 1];
  {
    const int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) umax[v] = (int)nearbyint(sqrt((double)kHalfPatch * kHalfPatch - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }
  // levels, cells
  std::vector<Cell> cells;
  int max_cw = 0, max_ch = 0, max_npx = 0;
  for (int l = 0; l < L; ++l) {
    P.w[l] = cv_roundf((float)cfg->width * h->inv_scale[l]);
    P.h[l] = cv_roundf((float)cfg->height * h->inv_scale[l]);
    P.pitch[l] = (P.w[l] + 63) & ~63;
    P.img_stride[l] = (size_t)P.pitch[l] * P.h[l];
    P.cell_begin[l] = (int)cells.size();
    const int maxBX = P.w[l] - kBorder, maxBY = P.h[l] - kBorder;
    const float width = (float)(maxBX - kBorder), height = (float)(maxBY - kBorder);
    const int nCols = (int)(width / 35.f), nRows = (int)(height / 35.f);
    if (P.w[l] < 1 || P.h[l] < 1) {
      set_error("level %d is empty (%dx%d)", l, P.w[l], P.h[l]);
      delete h;
      return VIEO_E_ARG;
    }
    if (nCols <= 0 || nRows <= 0) {  // level smaller than one 35-px cell: the reference's cell loops do not run
      P.maxn[l] = 1;
      continue;
    }
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    for (int i = 0; i < nRows; ++i) {
      const float iniY = (float)(kBorder + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; ++j) {
        const float iniX = (float)(kBorder + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        Cell c;
        c.level = (int16_t)l;
        c.x0 = (int16_t)iniX;
        c.y0 = (int16_t)iniY;
        c.cw = (int16_t)((int)maxX - (int)iniX);
        c.ch = (int16_t)((int)maxY - (int)iniY);
        cells.push_back(c);
        max_cw = std::max<int>(max_cw, c.cw);
        max_ch = std::max<int>(max_ch, c.ch);
        max_npx = std::max(max_npx, std::max(0, c.cw - 6) * std::max(0, c.ch - 6));
      }
    }
    const int spanX = maxBX - kBorder, spanY = maxBY - kBorder;
    const int nIni = (int)roundf((float)spanX / (float)spanY);
    if (nIni < 1) {
      set_error("level %d: aspect ratio gives zero quadtree roots (the reference divides by zero here)", l);
      delete h;
      return VIEO_E_ARG;
    }
    P.maxn[l] = std::max(P.quota[l] + 3, 4 * nIni) + 1;
  }
  P.cell_begin[L] = (int)cells.size();
  P.n_cells = (int)cells.size();
  P.tile_w = (max_cw + 3) & ~3;
  P.tile_h = max_ch;
  P.cell_cap = ((max_cw - 6 + 1) / 2) * ((max_ch - 6 + 1) / 2);
  bool ok = (max_npx + kFastThreads - 1) / kFastThreads <= 32;
  for (int l = 0; l < L && ok; ++l) {
    ok = (P.cell_begin[l + 1] - P.cell_begin[l]) <= 512 &&
         (size_t)(P.cell_begin[l + 1] - P.cell_begin[l]) * P.cell_cap < (1u << 20) && P.maxn[l] < 65536;
  }
  if (!ok) {
    set_error("unsupported geometry (cell grid / quota limits exceeded)");
    delete h;
    return VIEO_E_ARG;
  }
  P.slot_begin[0] = 0;
  int max_maxn = 0;
  for (int l = 0; l < L; ++l) {
    P.key_begin[l] = P.cell_begin[l] * P.cell_cap;
    P.slot_begin[l + 1] = P.slot_begin[l] + P.maxn[l];
    max_maxn = std::max(max_maxn, P.maxn[l]);
  }
  P.key_begin[L] = P.n_cells * P.cell_cap;
  h->cap_total = P.slot_begin[L];
  // TMA wants box rows of a multiple of 16 bytes that start on a 16-byte boundary of the level row: tile rows hold the
  // cell, the scorer's 8 bytes of slack and up to 15 bytes of alignment residue
  h->tma_ok = P.tile_w + 24 <= 256 && P.tile_h <= 256;
  P.tile_pitch = h->tma_ok ? ((P.tile_w + 24 + 15) & ~15) : P.tile_w + 8;
  h->tm0_ptr = nullptr; h->tm0_stride = 0; h->tm0_pitch = 0; h->tm0_n = 0;
  h->fast_smem = (((size_t)2 * P.tile_pitch * P.tile_h + 15) & ~(size_t)15) + 16 + 4 * (kFastThreads / 32);
  h->qt_smem = (size_t)max_maxn * (2 * (4 + 4 + 2 * 4 + 1) + 16 + 16 + 5 * 4) + 16 * 32;
  P.qt_knode_off = (int)((h->qt_smem + 15) & ~(size_t)15);
  h->qt_smem = (size_t)P.qt_knode_off + 2 * (size_t)kQtKnodeSmem;
  if (h->qt_smem > 200 * 1024) {
    set_error("per-level feature quota %d needs %zu B of shared memory for the quadtree (limit 200 KiB)", max_maxn,
              h->qt_smem);
    delete h;
    return VIEO_E_ARG;
  }

#define ORB_TRY(expr)          \
  do {                         \
    int rc_ = (expr);          \
    if (rc_) {                 \
      vieo_orb_destroy(h);     \
      return rc_;              \
    }                          \
  } while (0)
#define ORB_TRY_CUDA(call)                                                                         \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));              \
      vieo_orb_destroy(h);                                                                         \
      return VIEO_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

  ORB_TRY_CUDA(make_stream(&h->stream, false));
  ORB_TRY_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(umax)));
  const size_t B = cfg->max_batch;
  for (int l = 0; l < L; ++l) ORB_TRY(dev_alloc(h, &P.lvl[l], B * P.img_stride[l] + 64));
  memset(&h->tmaps, 0, sizeof(h->tmaps));
  for (int l = 1; l < L && h->tma_ok; ++l)
    h->tma_ok = encode_level_map(&h->tmaps.m[l], P.lvl[l], P.pitch[l], P.h[l], P.img_stride[l], (int)B, P.tile_pitch, P.tile_h);
  ORB_TRY(dev_upload(h, &P.cells, cells));
  // resize tables (cv::resize INTER_LINEAR 8U: 11-bit coefficients from float fx, see oracle + goldens)
  for (int l = 1; l < L; ++l) {
    const int sw = P.w[l - 1], sh = P.h[l - 1], dw = P.w[l], dh = P.h[l];
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    std::vector<int4> tx(dw), ty(dh);
    for (int dx = 0; dx < dw; ++dx) {
      float fx = (float)((dx + 0.5) * scale_x - 0.5);
      int sx = (int)floorf(fx);
      fx -= sx;
      if (sx < 0) fx = 0, sx = 0;
      if (sx >= sw - 1) fx = 0, sx = sw - 1;
      tx[dx] = make_int4(sx, std::min(sx + 1, sw - 1), (short)cv_roundf((1.f - fx) * 2048.f),
                         (short)cv_roundf(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
      float fy = (float)((dy + 0.5) * scale_y - 0.5);
      int sy = (int)floorf(fy);
      fy -= sy;
      ty[dy] = make_int4(std::min(std::max(sy, 0), sh - 1), std::min(std::max(sy + 1, 0), sh - 1),
                         (short)cv_roundf((1.f - fy) * 2048.f), (short)cv_roundf(fy * 2048.f));
    }
    ORB_TRY(dev_upload(h, &P.rx[l], tx));
    ORB_TRY(dev_upload(h, &P.ry[l], ty));
  }
  ORB_TRY(dev_alloc(h, &P.cand, B * P.n_cells * P.cell_cap));
  ORB_TRY(dev_alloc(h, &P.cell_cnt, B * P.n_cells));
  ORB_TRY(dev_alloc(h, &P.keys, B * P.key_begin[L]));
  ORB_TRY(dev_alloc(h, &P.knode, B * P.key_begin[L]));
  ORB_TRY(dev_alloc(h, &P.lvl_kp, B * P.slot_begin[L]));
  ORB_TRY(dev_alloc(h, &P.lvl_cnt, B * L));
  ORB_TRY(dev_alloc(h, &h->d_kps, B * h->cap_total));
  ORB_TRY(dev_alloc(h, &h->d_desc, B * h->cap_total * 32));
  ORB_TRY(dev_alloc(h, &h->d_nkp, B));
  ORB_TRY_CUDA(cudaMallocHost(&h->h_pinned, sizeof(int) * B));
  ORB_TRY_CUDA(cudaFuncSetAttribute(k_fast_cells<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fast_smem));
  ORB_TRY_CUDA(cudaFuncSetAttribute(k_fast_cells<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fast_smem));
  ORB_TRY_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->qt_smem));
  *out = h;
  return VIEO_OK;
}

void vieo_orb_destroy(vieo_orb_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  for (void* p : h->allocs) cudaFree(p);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int vieo_orb_max_keypoints(const vieo_orb_t* h) { return h ? h->cap_total : VIEO_E_ARG; }

int vieo_orb_get_tables(const vieo_orb_t* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                        int32_t* quota, int32_t* level_w, int32_t* level_h) {
  VIEO_ARG(h, "null handle");
  for (int l = 0; l < h->P.nlevels; ++l) {
    if (scale) scale[l] = h->P.scale[l];
    if (inv_scale) inv_scale[l] = h->inv_scale[l];
    if (sigma2) sigma2[l] = h->sigma2[l];
    if (inv_sigma2) inv_sigma2[l] = h->inv_sigma2[l];
    if (quota) quota[l] = h->P.quota[l];
    if (level_w) level_w[l] = h->P.w[l];
    if (level_h) level_h[l] = h->P.h[l];
  }
  return VIEO_OK;
}

int vieo_orb_extract_batch_dev(vieo_orb_t* h, int n_img, const uint8_t* imgs_dev, size_t img_stride, int row_stride,
                               VieoKeyPoint* kps_dev, uint8_t* desc_dev, int cap, int32_t* n_kp_dev, void* stream) {
  VIEO_ARG(h && imgs_dev && kps_dev && desc_dev && n_kp_dev, "null argument");
  VIEO_ARG(n_img >= 1 && n_img <= h->cfg.max_batch, "n_img exceeds max_batch");
  VIEO_ARG(row_stride >= h->cfg.width && cap >= 1, "bad stride/cap");
  VIEO_CK(cudaSetDevice(h->device));
  return orb_run(h, n_img, imgs_dev, img_stride, row_stride, kps_dev, desc_dev, cap, n_kp_dev, (cudaStream_t)stream);
}

int vieo_orb_extract_batch(vieo_orb_t* h, int n_img, const uint8_t* imgs, size_t img_stride, int row_stride,
                           VieoKeyPoint* kps, uint8_t* desc, int cap, int32_t* n_kp) {
  VIEO_ARG(h && imgs && kps && desc && n_kp, "null argument");
  VIEO_ARG(n_img >= 1 && n_img <= h->cfg.max_batch, "n_img exceeds max_batch");
  VIEO_ARG(row_stride >= h->cfg.width && cap >= 1, "bad stride/cap");
  VIEO_CK(cudaSetDevice(h->device));
  const OrbParams& P = h->P;
  const int dcap = std::min(cap, h->cap_total);
  cudaStream_t st = h->stream;
  size_t dev_stride = P.img_stride[0];
  int dev_pitch = P.pitch[0];
  if (row_stride <= P.pitch[0] && img_stride >= (size_t)row_stride * h->cfg.height && img_stride <= P.img_stride[0]) {
    // the caller's layout kept on the device: one contiguous copy at PCIe speed (see orb_enqueue_host)
    VIEO_CK(cudaMemcpyAsync(P.lvl[0], imgs, img_stride * (n_img - 1) + (size_t)row_stride * h->cfg.height,
                            cudaMemcpyHostToDevice, st));
    dev_stride = img_stride;
    dev_pitch = row_stride;
  } else {
    for (int i = 0; i < n_img; ++i)
      VIEO_CK(cudaMemcpy2DAsync(P.lvl[0] + i * P.img_stride[0], P.pitch[0], imgs + i * img_stride, row_stride,
                                h->cfg.width, h->cfg.height, cudaMemcpyHostToDevice, st));
  }
  int rc = orb_run(h, n_img, P.lvl[0], dev_stride, dev_pitch, h->d_kps, h->d_desc, dcap, h->d_nkp, st);
  if (rc) return rc;
  if (dcap == cap) {
    VIEO_CK(cudaMemcpyAsync(kps, h->d_kps, sizeof(VieoKeyPoint) * (size_t)n_img * cap, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(desc, h->d_desc, (size_t)32 * n_img * cap, cudaMemcpyDeviceToHost, st));
  } else {
    VIEO_CK(cudaMemcpy2DAsync(kps, sizeof(VieoKeyPoint) * (size_t)cap, h->d_kps, sizeof(VieoKeyPoint) * (size_t)dcap,
                              sizeof(VieoKeyPoint) * (size_t)dcap, n_img, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpy2DAsync(desc, (size_t)32 * cap, h->d_desc, (size_t)32 * dcap, (size_t)32 * dcap, n_img,
                              cudaMemcpyDeviceToHost, st));
  }
  VIEO_CK(cudaMemcpyAsync(n_kp, h->d_nkp, sizeof(int) * n_img, cudaMemcpyDeviceToHost, st));
  VIEO_CK(cudaStreamSynchronize(st));
  return VIEO_OK;
}

int vieo_orb_extract(vieo_orb_t* h, const uint8_t* img, int stride, const int32_t* lapping, VieoKeyPoint* kps,
                     uint8_t* desc, int cap, int32_t* n_mono, uint8_t* const* pyr_host) {
  VIEO_ARG(h, "null handle");
  if (n_mono) *n_mono = 0;
  if (!img || stride <= 0) return VIEO_E_EMPTY;
  VIEO_ARG(kps && desc && cap >= 1, "null output");
  const int W = h->cfg.width;
  VIEO_ARG(stride >= W, "stride < width");
  int n = 0;
  const int full = h->cap_total;
  std::vector<VieoKeyPoint> tk;
  std::vector<uint8_t> td;
  VieoKeyPoint* ok = kps;
  uint8_t* od = desc;
  const bool bounce = lapping != nullptr || cap < full;
  if (bounce) {
    tk.resize(full);
    td.resize((size_t)full * 32);
    ok = tk.data();
    od = td.data();
  }
  int rc = vieo_orb_extract_batch(h, 1, img, (size_t)stride * h->cfg.height, stride, ok, od, bounce ? full : cap, &n);
  if (rc) return rc;
  if (bounce) {
    if (n > cap) {
      set_error("caller capacity %d < %d keypoints", cap, n);
      return VIEO_E_CAPACITY;
    }
    if (lapping) {  // in-area keypoints fill from the back, the others from the front (src/ORBextractor.cc:1041-1052)
      int mono = 0, stereo = n - 1;
      for (int i = 0; i < n; ++i) {
        const bool in = tk[i].x >= (float)lapping[0] && tk[i].x <= (float)lapping[1];
        const int s = in ? stereo-- : mono++;
        kps[s] = tk[i];
        memcpy(desc + (size_t)s * 32, td.data() + (size_t)i * 32, 32);
      }
      if (n_mono) *n_mono = mono;
    } else {
      memcpy(kps, tk.data(), sizeof(VieoKeyPoint) * n);
      memcpy(desc, td.data(), (size_t)32 * n);
    }
  }
  if (pyr_host) {
    for (int l = 0; l < h->P.nlevels; ++l)
      if (pyr_host[l]) {
        rc = vieo_orb_debug_level(h, 0, l, pyr_host[l]);
        if (rc) return rc;
      }
  }
  return n;
}

int vieo_orb_debug_level(vieo_orb_t* h, int img_index, int level, uint8_t* out) {
  VIEO_ARG(h && out && level >= 0 && level < h->P.nlevels && img_index >= 0 && img_index < h->last_n_img,
           "bad argument");
  VIEO_CK(cudaSetDevice(h->device));
  const OrbParams& P = h->P;
  const uint8_t* src = level == 0 ? h->last_img0 + img_index * h->last_img0_stride
                                  : P.lvl[level] + img_index * P.img_stride[level];
  const int pitch = level == 0 ? h->last_img0_pitch : P.pitch[level];
  VIEO_CK(cudaDeviceSynchronize());
  VIEO_CK(cudaMemcpy2D(out, P.w[level], src, pitch, P.w[level], P.h[level], cudaMemcpyDeviceToHost));
  return VIEO_OK;
}

int vieo_orb_debug_candidates(vieo_orb_t* h, int img_index, int level, int32_t* xyr, int cap) {
  VIEO_ARG(h && level >= 0 && level < h->P.nlevels && img_index >= 0 && img_index < h->last_n_img, "bad argument");
  VIEO_CK(cudaSetDevice(h->device));
  const OrbParams& P = h->P;
  const int nc = P.cell_begin[level + 1] - P.cell_begin[level];
  std::vector<int> cnt(nc);
  std::vector<uint32_t> cand((size_t)nc * P.cell_cap);
  VIEO_CK(cudaDeviceSynchronize());
  VIEO_CK(cudaMemcpy(cnt.data(), P.cell_cnt + (size_t)img_index * P.n_cells + P.cell_begin[level], sizeof(int) * nc,
                     cudaMemcpyDeviceToHost));
  VIEO_CK(cudaMemcpy(cand.data(), P.cand + ((size_t)img_index * P.n_cells + P.cell_begin[level]) * P.cell_cap,
                     sizeof(uint32_t) * cand.size(), cudaMemcpyDeviceToHost));
  int n = 0;
  for (int c = 0; c < nc; ++c)
    for (int i = 0; i < cnt[c]; ++i, ++n) {
      if (n < cap && xyr) {
        const uint32_t k = cand[(size_t)c * P.cell_cap + i];
        xyr[3 * n] = k & 0xfff;
        xyr[3 * n + 1] = (k >> 12) & 0xfff;
        xyr[3 * n + 2] = k >> 24;
      }
    }
  return n;
}

// Frame::ComputeStereoMatches over the device-resident results of the last extract call of `h` (its pyramid levels
// are still in place): images 2f / 2f+1 are the left / right view of frame f.
int vieo_orb_stereo_match_dev(vieo_orb_t* h, int n_frames, const VieoKeyPoint* kps_dev, const uint8_t* desc_dev,
                              const int32_t* n_kp_dev, int cap, float bf, float min_z, float* uright_dev, float* depth_dev,
                              int32_t* sad_dev, void* stream) {
  VIEO_ARG(h && kps_dev && desc_dev && n_kp_dev && uright_dev && depth_dev && sad_dev, "null argument");
  VIEO_ARG(n_frames >= 1 && 2 * n_frames <= h->last_n_img && cap >= 1 && bf > 0 && min_z > 0, "bad argument");
  VIEO_CK(cudaSetDevice(h->device));
  const OrbParams& P = h->P;
  StereoLevels Lv;
  for (int l = 0; l < P.nlevels; ++l) {
    const bool ext = l == 0 && h->last_img0 != nullptr;
    Lv.base[l] = ext ? h->last_img0 : P.lvl[l];
    Lv.img_stride[l] = ext ? h->last_img0_stride : P.img_stride[l];
    Lv.pitch[l] = ext ? h->last_img0_pitch : P.pitch[l];
    Lv.w[l] = P.w[l];
    Lv.scale[l] = P.scale[l];
    Lv.inv_scale[l] = h->inv_scale[l];
  }
  // shared memory: 9 B per right keypoint (match) / 4 B per left keypoint (filter); stays below the 48 KB default
  const size_t smem = ((size_t)9 * cap + 15) & ~(size_t)15, smem_f = sizeof(int) * (size_t)cap;
  VIEO_ARG(smem <= 48 * 1024, "cap too large for the stereo kernels (at most 5461 keypoints per image)");
  VIEO_ARG(n_frames <= 65535, "too many frames per call");
  k_stereo_match<<<dim3(kStereoSplit, n_frames), 256, smem, (cudaStream_t)stream>>>(Lv, kps_dev, desc_dev, n_kp_dev, cap, P.h[0],
                                                                                    bf, min_z, uright_dev, depth_dev, sad_dev);
  k_stereo_filter<<<n_frames, 256, smem_f, (cudaStream_t)stream>>>(n_kp_dev, cap, uright_dev, depth_dev, sad_dev);
  VIEO_CK(cudaGetLastError());
  return VIEO_OK;
}

int vieo_orb_last_launches(const vieo_orb_t* h) { return h ? h->last_launches : VIEO_E_ARG; }

int vieo_orb_profile(vieo_orb_t* h, int enable) {
  VIEO_ARG(h, "null handle");
  h->prof_on = enable != 0;
  h->prof_calls = 0;
  return VIEO_OK;
}

int vieo_orb_profile_read(vieo_orb_t* h, float stage_ms[4], int32_t* n_calls) {
  VIEO_ARG(h && stage_ms, "null argument");
  VIEO_CK(cudaSetDevice(h->device));
  for (int i = 0; i < 4; ++i) stage_ms[i] = 0.f;
  for (int c = 0; c < h->prof_calls; ++c) {
    cudaEvent_t* ev = h->prof_ev.data() + (size_t)5 * c;
    VIEO_CK(cudaEventSynchronize(ev[4]));
    for (int i = 0; i < 4; ++i) {
      float ms = 0.f;
      VIEO_CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      stage_ms[i] += ms;
    }
  }
  if (n_calls) *n_calls = h->prof_calls;
  h->prof_calls = 0;
  return VIEO_OK;
}

}  // extern "C"

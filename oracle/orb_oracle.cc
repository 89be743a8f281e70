// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's ORB front-end.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product path (vieo_slam_b200/) never does.
//
// Parity status: the reference (leavesnight/VIEO_SLAM @356e4a22) cannot be compiled here (needs
// OpenCV C++/Eigen/Sophus, all absent) and ships no tests, so this restatement is pinned against
// python cv2 4.13 for every OpenCV primitive the path calls (tests/golden/gen_cv2_goldens.py):
// resize(INTER_LINEAR, 8U), FAST(9/16, nms), GaussianBlur(7x7, sigma 2, REFLECT_101), fastAtan2;
// and against glibc 2.39 cosf/sinf for the descriptor steering.  The quadtree and the glue follow
// the reference source directly (citations on each function, relative to /root/reference).
//
// Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -std=c++17 -shared -fPIC (oracle/Makefile).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

const int8_t kPattern[1024] = {
#include "../data/orb_pattern_31.inc"
};

constexpr int kHalfPatch = 15;   // src/ORBextractor.cc:52
constexpr int kEdge = 19;        // src/ORBextractor.cc:53
constexpr int kBorder = kEdge - 3;  // 16, src/ORBextractor.cc:729

inline int cv_round(double v) { return (int)std::nearbyint(v); }  // round-half-even (cvRound)
inline int cv_roundf(float v) { return (int)std::nearbyintf(v); }

// ---------------------------------------------------------------------------------------------
// glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h algorithm), restated so that the CUDA
// kernel can evaluate the identical double-precision polynomial.  Verified bit-identical to this
// container's libm over every float in [0, 2*pi] (tests/test_oracle_orb.py samples it).
struct SinCosTab {
  double sign[4], hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};
const SinCosTab kSC[2] = {
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2,
     0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2,
     -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
inline float sc_poly(double x, double x2, const SinCosTab& p, int n) {
  if ((n & 1) == 0) {
    double x3 = x * x2, s1 = p.s2 + x2 * p.s3, x7 = x3 * x2, s = x + x3 * p.s1;
    return (float)(s + x7 * s1);
  }
  double x4 = x2 * x2, c2 = p.c3 + x2 * p.c4, c1 = p.c0 + x2 * p.c1, x6 = x4 * x2, c = c1 + x4 * p.c2;
  return (float)(c + x6 * c2);
}
inline uint32_t top12(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return (u >> 20) & 0x7ff;
}
// valid for |y| < 120 (the extractor only feeds [0, 2*pi])
void sincos_glibc(float y, float* sn, float* cs) {
  double x = y;
  if (top12(y) < top12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (top12(y) < top12(0x1p-12f)) {
      *sn = y;
      *cs = 1.0f;
      return;
    }
    *sn = sc_poly(x, x2, kSC[0], 0);
    *cs = sc_poly(x, x2, kSC[0], 1);
    return;
  }
  double r = x * kSC[0].hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  x = x - n * kSC[0].hpi;
  double s = kSC[0].sign[n & 3];
  const SinCosTab& p = kSC[(n & 2) ? 1 : 0];
  *sn = sc_poly(x * s, x * x, p, n);
  *cs = sc_poly(x * s, x * x, p, n ^ 1);
}

}  // namespace

extern "C" {

void orc_sincosf(float x, float* s, float* c) { sincos_glibc(x, s, c); }

// cv::fastAtan2 (scalar path of OpenCV core/mathfuncs_core: 7th-order odd polynomial in fp32,
// degrees).  Called at src/ORBextractor.cc:79.  Pinned against cv2.fastAtan2.
float orc_fast_atan2(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
  const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
  const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
  const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 (OpenCV imgproc/resize.cpp fixed-point
// path: 11-bit coefficients, HResizeLinear -> int, VResizeLinear with the (>>4, >>16, +2, >>2)
// rounding).  Called per level at src/ORBextractor.cc:1070.  Pinned against cv2.resize.
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                          int dstride) {
  const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
  std::vector<int> xofs(dw), xofs1(dw);
  std::vector<short> a0(dw), a1(dw);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) fx = 0, sx = 0;
    if (sx >= sw - 1) fx = 0, sx = sw - 1;
    xofs[dx] = sx;
    xofs1[dx] = std::min(sx + 1, sw - 1);
    a0[dx] = (short)cv_roundf((1.f - fx) * 2048.f);
    a1[dx] = (short)cv_roundf(fx * 2048.f);
  }
  std::vector<int> row0(dw), row1(dw);
  int cached0 = -1, cached1 = -1;
  auto hrow = [&](int sy, std::vector<int>& out) {
    const uint8_t* S = src + (size_t)sy * sstride;
    for (int dx = 0; dx < dw; ++dx) out[dx] = S[xofs[dx]] * a0[dx] + S[xofs1[dx]] * a1[dx];
  };
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    short b0 = (short)cv_roundf((1.f - fy) * 2048.f), b1 = (short)cv_roundf(fy * 2048.f);
    int y0 = std::min(std::max(sy, 0), sh - 1), y1 = std::min(std::max(sy + 1, 0), sh - 1);
    if (y0 != cached0) {
      if (y0 == cached1) {
        row0.swap(row1);
        std::swap(cached0, cached1);
      } else {
        hrow(y0, row0);
        cached0 = y0;
      }
    }
    if (y1 != cached1) {
      if (y1 == cached0) {
        row1 = row0;
      } else {
        hrow(y1, row1);
      }
      cached1 = y1;
    }
    uint8_t* D = dst + (size_t)dy * dstride;
    for (int dx = 0; dx < dw; ++dx)
      D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

// FAST-9/16 arc score S(x,y) = max over the 16 contiguous 9-arcs of min(v - p_k), and of min(p_k - v)
// (OpenCV features2d/fast_score.cpp cornerScore<16> returns max(th, S) - 1; a pixel is a corner at
// threshold th iff S > th).  Output clamped to [0,255]; pixels closer than 3 to the border get 0.
// Written on raw u8 values, S = max(v -sat min_arcs(max9 p), max_arcs(min9 p) -sat v), so that the x loop
// auto-vectorises (vpminub/vpmaxub): this is what makes the CPU baseline a fair one.
static inline uint8_t u8min(uint8_t a, uint8_t b) { return a < b ? a : b; }
static inline uint8_t u8max(uint8_t a, uint8_t b) { return a > b ? a : b; }
static void score_row(const uint8_t* __restrict c, int stride, int x0, int x1, uint8_t* __restrict out) {
  const uint8_t* r[16] = {c + 3 * stride,     c + 3 * stride + 1, c + 2 * stride + 2, c + stride + 3,
                          c + 3,              c - stride + 3,     c - 2 * stride + 2, c - 3 * stride + 1,
                          c - 3 * stride,     c - 3 * stride - 1, c - 2 * stride - 2, c - stride - 3,
                          c - 3,              c + stride - 3,     c + 2 * stride - 2, c + 3 * stride - 1};
  for (int x = x0; x < x1; ++x) {
    uint8_t d[16], lo2[16], hi2[16], lo4[16], hi4[16];
    for (int k = 0; k < 16; ++k) d[k] = r[k][x];
    for (int k = 0; k < 16; ++k) {
      lo2[k] = u8min(d[k], d[(k + 1) & 15]);
      hi2[k] = u8max(d[k], d[(k + 1) & 15]);
    }
    for (int k = 0; k < 16; ++k) {
      lo4[k] = u8min(lo2[k], lo2[(k + 2) & 15]);
      hi4[k] = u8max(hi2[k], hi2[(k + 2) & 15]);
    }
    uint8_t A = 255, B = 0;
    for (int k = 0; k < 16; ++k) {
      A = u8min(A, u8max(u8max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]));
      B = u8max(B, u8min(u8min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]));
    }
    const uint8_t v = c[x];
    const uint8_t s1 = v > A ? (uint8_t)(v - A) : (uint8_t)0, s2 = B > v ? (uint8_t)(B - v) : (uint8_t)0;
    out[x] = u8max(s1, s2);
  }
}
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ostride) {
  for (int y = 0; y < h; ++y) {
    uint8_t* o = out + (size_t)y * ostride;
    if (y < 3 || y >= h - 3 || w < 7) {
      memset(o, 0, w);
      continue;
    }
    o[0] = o[1] = o[2] = o[w - 1] = o[w - 2] = o[w - 3] = 0;
    score_row(img + (size_t)y * stride, stride, 3, w - 3, o);
  }
}

// cv::FAST(img, kps, th, nonmaxSuppression=true) on one (cell) image, OpenCV features2d/fast.cpp FAST_t<16>
// structure: opposite-pair quick rejection, 9-contiguous test, cornerScore only for corners, then strict 3x3
// NMS on the score rows (scores of non-corners are 0, the 3-px frame is never a corner).
// Raster order; response = cornerScore = S-1.  Returns count (xs/ys/resp may be null).  Pinned against cv2.
static inline int arc_score(const uint8_t* p, const int* off) {
  int v = p[0], d[25];
  for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = 0;
  for (int k = 0; k < 16; ++k) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; ++j) {
      mn = std::min(mn, d[k + j]);
      mx = std::max(mx, d[k + j]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}
static inline bool has9(uint32_t m) {  // 9 consecutive set bits in a circular 16-bit mask
  uint32_t r = m | (m << 16);
  r &= r >> 1;
  r &= r >> 2;
  r &= r >> 4;
  r &= (m | (m << 16)) >> 8;
  return r != 0;
}
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int th, int* xs, int* ys, int* resp, int cap) {
  if (w < 7 || h < 7) return 0;
  static const int ox[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int oy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  int off[16];
  for (int k = 0; k < 16; ++k) off[k] = oy[k] * stride + ox[k];
  std::vector<uint8_t> S((size_t)w * h, 0);  // corner scores (S-1 < 255), 0 elsewhere
  for (int y = 3; y < h - 3; ++y) {
    const uint8_t* row = img + (size_t)y * stride;
    uint8_t* srow = S.data() + (size_t)y * w;
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = row + x;
      const int v = p[0], lo = v - th, hi = v + th;
      // a 9-arc contains one pixel of every opposite pair
      int a = p[off[0]], b = p[off[8]];
      if (a >= lo && a <= hi && b >= lo && b <= hi) continue;
      a = p[off[4]], b = p[off[12]];
      if (a >= lo && a <= hi && b >= lo && b <= hi) continue;
      uint32_t mb = 0, md = 0;
      for (int k = 0; k < 16; ++k) {
        const int q = p[off[k]];
        mb |= (uint32_t)(q > hi) << k;
        md |= (uint32_t)(q < lo) << k;
      }
      if (!has9(mb) && !has9(md)) continue;
      srow[x] = (uint8_t)std::min(arc_score(p, off) - 1, 255);
    }
  }
  int n = 0;
  for (int y = 3; y < h - 3; ++y) {
    const uint8_t* s1 = S.data() + (size_t)y * w;
    const uint8_t *s0 = s1 - w, *s2 = s1 + w;
    for (int x = 3; x < w - 3; ++x) {
      const int s = s1[x];
      if (!s) continue;
      if (s > s0[x - 1] && s > s0[x] && s > s0[x + 1] && s > s1[x - 1] && s > s1[x + 1] && s > s2[x - 1] &&
          s > s2[x] && s > s2[x + 1]) {
        if (n < cap && xs) {
          xs[n] = x;
          ys[n] = y;
          resp[n] = s;
        }
        ++n;
      }
    }
  }
  return n;
}

// cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (OpenCV
// imgproc/smooth fixed-point path: 8-fractional-bit kernel from getGaussianKernel(7,2) with error
// diffusion, exact integer separable accumulation, one final round (+2^15)>>16).
// Called at src/ORBextractor.cc:1013 on a clone of the level ROI.  Pinned against cv2.GaussianBlur.
void orc_gauss7_kernel(int k[7]) {
  double g[7], sum = 0;
  for (int i = 0; i < 7; ++i) {
    double x = i - 3;
    g[i] = std::exp(-0.5 * x * x / 4.0);
    sum += g[i];
  }
  double err = 0;
  int s = 0;
  for (int i = 0; i < 3; ++i) {
    double adj = g[i] / sum * 256.0 + err;
    int v = cv_round(adj);
    err = adj - v;
    k[i] = k[6 - i] = v;
    s += v;
  }
  k[3] = 256 - 2 * s;
}
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}
void orc_gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  int k[7];
  orc_gauss7_kernel(k);
  std::vector<uint16_t> tmp((size_t)w * h);
  std::vector<uint8_t> line(w + 6);
  for (int y = 0; y < h; ++y) {
    for (int x = -3; x < w + 3; ++x) line[x + 3] = src[(size_t)y * sstride + reflect101(x, w)];
    uint16_t* t = &tmp[(size_t)y * w];
    const uint8_t* l = line.data();
    for (int x = 0; x < w; ++x)
      t[x] = (uint16_t)(k[0] * (l[x] + l[x + 6]) + k[1] * (l[x + 1] + l[x + 5]) + k[2] * (l[x + 2] + l[x + 4]) +
                        k[3] * l[x + 3]);
  }
  for (int y = 0; y < h; ++y) {
    const uint16_t* r[7];
    for (int i = 0; i < 7; ++i) r[i] = &tmp[(size_t)reflect101(y + i - 3, h) * w];
    uint8_t* d = dst + (size_t)y * dstride;
    for (int x = 0; x < w; ++x) {
      uint32_t acc = (uint32_t)k[0] * (r[0][x] + r[6][x]) + (uint32_t)k[1] * (r[1][x] + r[5][x]) +
                     (uint32_t)k[2] * (r[2][x] + r[4][x]) + (uint32_t)k[3] * r[3][x];
      d[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

}  // extern "C"

// =============================================================================================
// The extractor proper.
namespace {

struct Cand {
  float x, y;  // level coordinates relative to (minBorder, minBorder)
  float response;
};

struct OrbOracle {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // the reference stores the float argument in a double member (ORBextractor.h:64)
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> quota;
  int umax[kHalfPatch + 1];
  // last run
  int w0 = 0, h0 = 0;
  std::vector<int> lw, lh;
  std::vector<std::vector<uint8_t>> pyr;
  std::vector<std::vector<Cand>> cands;  // per level, pre-quadtree, cell-major raster order
  std::vector<uint8_t> score;
};

// src/ORBextractor.cc:391-456
void orb_init(OrbOracle& o, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  o.nfeatures = nfeatures;
  o.nlevels = nlevels;
  o.iniTh = iniTh;
  o.minTh = minTh;
  o.scaleFactor = scaleFactor;
  o.scale.assign(nlevels, 1.f);
  o.sigma2.assign(nlevels, 1.f);
  for (int i = 1; i < nlevels; ++i) {
    o.scale[i] = (float)(o.scale[i - 1] * o.scaleFactor);
    o.sigma2[i] = o.scale[i] * o.scale[i];
  }
  o.invScale.resize(nlevels);
  o.invSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; ++i) {
    o.invScale[i] = 1.0f / o.scale[i];
    o.invSigma2[i] = 1.0f / o.sigma2[i];
  }
  o.quota.resize(nlevels);
  float factor = (float)(1.0f / o.scaleFactor);
  float per = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; ++l) {
    o.quota[l] = cv_roundf(per);
    sum += o.quota[l];
    per *= factor;
  }
  o.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
  // circular patch row extents (src/ORBextractor.cc:439-455)
  int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
  int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
  const double hp2 = kHalfPatch * kHalfPatch;
  for (int v = 0; v <= vmax; ++v) o.umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (o.umax[v0] == o.umax[v0 + 1]) ++v0;
    o.umax[v] = v0;
    ++v0;
  }
}

// src/ORBextractor.cc:1060-1081 (the 19-px REFLECT_101 frame is never read by this path, so the
// oracle keeps bare ROIs)
void build_pyramid(OrbOracle& o, const uint8_t* img, int w, int h, int stride) {
  o.w0 = w;
  o.h0 = h;
  o.lw.resize(o.nlevels);
  o.lh.resize(o.nlevels);
  o.pyr.resize(o.nlevels);
  for (int l = 0; l < o.nlevels; ++l) {
    o.lw[l] = cv_roundf((float)w * o.invScale[l]);
    o.lh[l] = cv_roundf((float)h * o.invScale[l]);
    o.pyr[l].resize((size_t)o.lw[l] * o.lh[l]);
    if (l == 0) {
      for (int y = 0; y < h; ++y) memcpy(&o.pyr[0][(size_t)y * w], img + (size_t)y * stride, w);
    } else {
      orc_resize_linear_u8(o.pyr[l - 1].data(), o.lw[l - 1], o.lh[l - 1], o.lw[l - 1], o.pyr[l].data(), o.lw[l],
                           o.lh[l], o.lw[l]);
    }
  }
}

// per-cell FAST with ini->min threshold fallback (src/ORBextractor.cc:738-779).  Each cell is an independent
// cv::FAST call: corners only inside the cell's 3-px-inset interior and NMS blind to neighbouring cells.  S is a
// function of the pixel neighbourhood only, so it is computed once per level and each cell applies
// "S > th and S > S_n for the 8 neighbours inside my interior" — identical to running orc_fast_detect on
// the cell image (tests compare both against per-cell cv2.FAST goldens).
void detect_level(const OrbOracle& o, int level, std::vector<Cand>& out, std::vector<uint8_t>& S) {
  out.clear();
  const int W = o.lw[level], H = o.lh[level];
  const uint8_t* img = o.pyr[level].data();
  const int minBX = kBorder, minBY = kBorder, maxBX = W - kEdge + 3, maxBY = H - kEdge + 3;
  const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
  const int nCols = (int)(width / 35.f), nRows = (int)(height / 35.f);
  if (nCols <= 0 || nRows <= 0) return;
  const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
  S.resize((size_t)W * H);
  orc_fast_score_map(img, W, H, W, S.data(), W);
  for (int i = 0; i < nRows; ++i) {
    const float iniY = (float)(minBY + i * hCell);
    float maxY = iniY + hCell + 6;
    if (iniY >= maxBY - 3) continue;
    if (maxY > maxBY) maxY = (float)maxBY;
    for (int j = 0; j < nCols; ++j) {
      const float iniX = (float)(minBX + j * wCell);
      float maxX = iniX + wCell + 6;
      if (iniX >= maxBX - 6) continue;
      if (maxX > maxBX) maxX = (float)maxBX;
      const int xa = (int)iniX + 3, xb = (int)maxX - 3, ya = (int)iniY + 3, yb = (int)maxY - 3;  // interior
      const size_t first = out.size();
      for (int pass = 0; pass < 2 && out.size() == first; ++pass) {
        const int th = pass == 0 ? o.iniTh : o.minTh;
        for (int y = ya; y < yb; ++y) {
          const uint8_t* s1 = S.data() + (size_t)y * W;
          for (int x = xa; x < xb; ++x) {
            const int s = s1[x];
            if (s <= th) continue;
            bool keep = true;
            for (int dy = -1; dy <= 1 && keep; ++dy) {
              const int yy = y + dy;
              if (yy < ya || yy >= yb) continue;
              for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if ((!dx && !dy) || xx < xa || xx >= xb) continue;
                if (S[(size_t)yy * W + xx] >= s) {
                  keep = false;
                  break;
                }
              }
            }
            if (keep) out.push_back({(float)(x - minBX), (float)(y - minBY), (float)(s - 1)});
          }
        }
      }
    }
  }
}

// Quadtree distribution (src/ORBextractor.cc:467-721).  Own formulation: index-linked list over a
// node pool.  Deterministic tie-break for the (size, pointer) sort of :647-651 — the reference
// compares heap addresses there, which is not reproducible — is creation order.
struct QNode {
  int x0, x1, y0, y1;
  std::vector<int> keys;
  int prev = -1, next = -1;
  bool single = false;
};
struct QList {
  std::vector<QNode> pool;
  int head = -1, tail = -1, count = 0;
  int push_front(QNode&& n) {
    int id = (int)pool.size();
    pool.push_back(std::move(n));
    pool[id].prev = -1;
    pool[id].next = head;
    if (head >= 0) pool[head].prev = id;
    head = id;
    if (tail < 0) tail = id;
    ++count;
    return id;
  }
  int push_back(QNode&& n) {
    int id = (int)pool.size();
    pool.push_back(std::move(n));
    pool[id].next = -1;
    pool[id].prev = tail;
    if (tail >= 0) pool[tail].next = id;
    tail = id;
    if (head < 0) head = id;
    ++count;
    return id;
  }
  int erase(int id) {  // returns next
    int p = pool[id].prev, n = pool[id].next;
    if (p >= 0) pool[p].next = n; else head = n;
    if (n >= 0) pool[n].prev = p; else tail = p;
    --count;
    std::vector<int>().swap(pool[id].keys);
    return n;
  }
};

void split_node(const QNode& n, const std::vector<Cand>& c, QNode ch[4]) {
  const int halfX = (int)std::ceil((float)(n.x1 - n.x0) / 2), halfY = (int)std::ceil((float)(n.y1 - n.y0) / 2);
  const int mx = n.x0 + halfX, my = n.y0 + halfY;
  ch[0] = {n.x0, mx, n.y0, my};
  ch[1] = {mx, n.x1, n.y0, my};
  ch[2] = {n.x0, mx, my, n.y1};
  ch[3] = {mx, n.x1, my, n.y1};
  for (int k : n.keys) {
    int q = (c[k].x < (float)mx ? 0 : 1) + (c[k].y < (float)my ? 0 : 2);
    ch[q].keys.push_back(k);
  }
  for (int q = 0; q < 4; ++q) ch[q].single = ch[q].keys.size() == 1;
}

void distribute_quadtree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N,
                         std::vector<int>& picked) {
  picked.clear();
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  const float hX = (float)(maxX - minX) / nIni;
  QList L;
  L.pool.reserve(4 * c.size() + 16);
  std::vector<int> ini(nIni);
  for (int i = 0; i < nIni; ++i) {
    QNode n{(int)(hX * (float)i), (int)(hX * (float)(i + 1)), 0, maxY - minY};
    ini[i] = L.push_back(std::move(n));
  }
  for (int k = 0; k < (int)c.size(); ++k) L.pool[ini[(int)(c[k].x / hX)]].keys.push_back(k);
  for (int it = L.head; it >= 0;) {
    QNode& n = L.pool[it];
    if (n.keys.size() == 1) {
      n.single = true;
      it = n.next;
    } else if (n.keys.empty())
      it = L.erase(it);
    else
      it = n.next;
  }
  bool finish = false;
  std::vector<std::pair<int, int>> expandable;  // (size, node id == creation order)
  auto add_children = [&](QNode ch[4], int& nToExpand) {
    for (int q = 0; q < 4; ++q) {
      if (ch[q].keys.empty()) continue;
      int sz = (int)ch[q].keys.size();
      int id = L.push_front(std::move(ch[q]));
      if (sz > 1) {
        ++nToExpand;
        expandable.emplace_back(sz, id);
      }
    }
  };
  while (!finish) {
    const int prevSize = L.count;
    int nToExpand = 0;
    expandable.clear();
    for (int it = L.head; it >= 0;) {
      if (L.pool[it].single) {
        it = L.pool[it].next;
        continue;
      }
      QNode ch[4];
      split_node(L.pool[it], c, ch);
      add_children(ch, nToExpand);
      it = L.erase(it);
    }
    if (L.count >= N || L.count == prevSize) {
      finish = true;
    } else if (L.count + nToExpand * 3 > N) {
      while (!finish) {
        const int prev2 = L.count;
        std::vector<std::pair<int, int>> prevExp = expandable;
        expandable.clear();
        std::sort(prevExp.begin(), prevExp.end());
        for (int j = (int)prevExp.size() - 1; j >= 0; --j) {
          QNode ch[4];
          int dummy = 0;
          split_node(L.pool[prevExp[j].second], c, ch);
          add_children(ch, dummy);
          L.erase(prevExp[j].second);
          if (L.count >= N) break;
        }
        if (L.count >= N || L.count == prev2) finish = true;
      }
    }
  }
  for (int it = L.head; it >= 0; it = L.pool[it].next) {
    const std::vector<int>& ks = L.pool[it].keys;
    int best = ks[0];
    for (size_t k = 1; k < ks.size(); ++k)
      if (c[ks[k]].response > c[best].response) best = ks[k];
    picked.push_back(best);
  }
}

// src/ORBextractor.cc:55-80
float ic_angle(const OrbOracle& o, const uint8_t* img, int stride, int cx, int cy) {
  int m01 = 0, m10 = 0;
  const uint8_t* center = img + (size_t)cy * stride + cx;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vsum = 0, d = o.umax[v];
    for (int u = -d; u <= d; ++u) {
      int vp = center[u + v * stride], vm = center[u - v * stride];
      vsum += vp - vm;
      m10 += u * (vp + vm);
    }
    m01 += v * vsum;
  }
  return orc_fast_atan2((float)m01, (float)m10);
}

// src/ORBextractor.cc:82-127
void brief_descriptor(const uint8_t* blurred, int stride, int cx, int cy, float angleDeg, uint8_t* desc) {
  const float factorPI = (float)(M_PI / 180.f);
  float angle = angleDeg * factorPI, a, b;
  sincos_glibc(angle, &b, &a);
  const uint8_t* center = blurred + (size_t)cy * stride + cx;
  const int8_t* p = kPattern;
  for (int i = 0; i < 32; ++i) {
    int val = 0;
    for (int t = 0; t < 8; ++t, p += 4) {
      int t0 = center[cv_roundf(p[0] * b + p[1] * a) * stride + cv_roundf(p[0] * a - p[1] * b)];
      int t1 = center[cv_roundf(p[2] * b + p[3] * a) * stride + cv_roundf(p[2] * a - p[3] * b)];
      val |= (t0 < t1) << t;
    }
    desc[i] = (uint8_t)val;
  }
}

}  // namespace

extern "C" {

void* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  OrbOracle* o = new OrbOracle;
  orb_init(*o, nfeatures, scaleFactor, nlevels, iniTh, minTh);
  return o;
}
void orc_orb_destroy(void* h) { delete (OrbOracle*)h; }
void orc_orb_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* quota, int* umax) {
  OrbOracle& o = *(OrbOracle*)h;
  for (int i = 0; i < o.nlevels; ++i) {
    scale[i] = o.scale[i];
    invScale[i] = o.invScale[i];
    sigma2[i] = o.sigma2[i];
    invSigma2[i] = o.invSigma2[i];
    quota[i] = o.quota[i];
  }
  for (int i = 0; i <= kHalfPatch; ++i) umax[i] = o.umax[i];
}

// ORBextractor::operator() (src/ORBextractor.cc:968-1058).  lapping==NULL -> level-ordered output.
// Returns the number of keypoints (<= cap written), *n_mono = the reference's return value; -1 for
// an empty image.
int orc_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, const int* lapping, OrcKeyPoint* kps,
                    uint8_t* desc, int cap, int* n_mono) {
  OrbOracle& o = *(OrbOracle*)h;
  if (!img || w <= 0 || hgt <= 0) return -1;
  build_pyramid(o, img, w, hgt, stride);
  o.cands.resize(o.nlevels);
  std::vector<std::vector<OrcKeyPoint>> all(o.nlevels);
  std::vector<int> picked;
  for (int l = 0; l < o.nlevels; ++l) {
    detect_level(o, l, o.cands[l], o.score);
    const int W = o.lw[l], H = o.lh[l];
    distribute_quadtree(o.cands[l], kBorder, W - kEdge + 3, kBorder, H - kEdge + 3, o.quota[l], picked);
    const int patch = (int)(31 * o.scale[l]);
    for (int id : picked) {
      const Cand& c = o.cands[l][id];
      OrcKeyPoint k;
      k.x = c.x + kBorder;
      k.y = c.y + kBorder;
      k.size = (float)patch;
      k.response = c.response;
      k.octave = l;
      k.angle = ic_angle(o, o.pyr[l].data(), W, cv_roundf(k.x), cv_roundf(k.y));
      all[l].push_back(k);
    }
  }
  int total = 0;
  for (auto& v : all) total += (int)v.size();
  int mono = 0, stereo = total - 1, offset = 0;
  std::vector<uint8_t> blurred, d(32);
  for (int l = 0; l < o.nlevels; ++l) {
    if (all[l].empty()) continue;
    const int W = o.lw[l], H = o.lh[l];
    blurred.resize((size_t)W * H);
    orc_gaussian_blur7_u8(o.pyr[l].data(), W, H, W, blurred.data(), W);
    for (OrcKeyPoint k : all[l]) {
      brief_descriptor(blurred.data(), W, cv_roundf(k.x), cv_roundf(k.y), k.angle, d.data());
      if (l != 0) {
        k.x *= o.scale[l];
        k.y *= o.scale[l];
      }
      int slot;
      if (lapping) {
        slot = (k.x >= lapping[0] && k.x <= lapping[1]) ? stereo-- : mono++;
      } else
        slot = offset++;
      if (slot < cap) {
        kps[slot] = k;
        memcpy(desc + (size_t)slot * 32, d.data(), 32);
      }
    }
  }
  if (n_mono) *n_mono = mono;
  return total;
}

// introspection for stage-by-stage parity tests
int orc_orb_level_size(void* h, int level, int* w, int* hgt) {
  OrbOracle& o = *(OrbOracle*)h;
  if (level >= (int)o.lw.size()) return -1;
  *w = o.lw[level];
  *hgt = o.lh[level];
  return 0;
}
void orc_orb_get_level(void* h, int level, uint8_t* out) {
  OrbOracle& o = *(OrbOracle*)h;
  memcpy(out, o.pyr[level].data(), o.pyr[level].size());
}
// candidates (pre-quadtree) of a level as packed (x, y, response) ints in level coordinates
int orc_orb_get_candidates(void* h, int level, int* xyr, int cap) {
  OrbOracle& o = *(OrbOracle*)h;
  int n = (int)o.cands[level].size();
  for (int i = 0; i < n && i < cap; ++i) {
    xyr[3 * i] = (int)o.cands[level][i].x + kBorder;
    xyr[3 * i + 1] = (int)o.cands[level][i].y + kBorder;
    xyr[3 * i + 2] = (int)o.cands[level][i].response;
  }
  return n;
}
// quadtree alone: cands = (x,y,response) in level coords (absolute), returns picked indices
int orc_quadtree(const int* xyr, int n, int W, int H, int N, int* picked, int cap) {
  std::vector<Cand> c(n);
  for (int i = 0; i < n; ++i) c[i] = {(float)(xyr[3 * i] - kBorder), (float)(xyr[3 * i + 1] - kBorder), (float)xyr[3 * i + 2]};
  std::vector<int> p;
  distribute_quadtree(c, kBorder, W - kEdge + 3, kBorder, H - kEdge + 3, N, p);
  for (int i = 0; i < (int)p.size() && i < cap; ++i) picked[i] = p[i];
  return (int)p.size();
}

}  // extern "C"

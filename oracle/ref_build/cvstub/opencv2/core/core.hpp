// ORACLE — TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for the OpenCV headers, just wide enough that the REFERENCE's own translation unit
// /root/reference/src/ORBextractor.cc compiles UNCHANGED in this image (which has no OpenCV C++).  The container /
// header types below carry no algorithm; the six OpenCV *primitives* the reference calls (cv::FAST, cv::resize
// INTER_LINEAR, cv::GaussianBlur 7x7, cv::copyMakeBorder, cv::fastAtan2, cvRound/cvFloor/cvCeil) forward to the oracle's
// restatements, each of which is pinned bit-exactly against python cv2 4.13 by tests/golden/cv2_orb_goldens.npz.
// Behaviour that matters for the reference's control flow and is reproduced here on purpose:
//   * Mat is a ref-counted header over a shared buffer; ROI headers (rowRange / colRange / operator()(Rect)) alias it;
//   * Mat::create() on a header that already has the requested size and type is a no-op (so cv::resize into an ROI,
//     and `descriptors = Mat::zeros(...)` on a rowRange header, write THROUGH to the parent buffer — the reference's
//     computeDescriptors relies on exactly that, src/ORBextractor.cc:960-966,1019-1024);
//   * cvRound is round-half-to-even (SSE cvtss2si / cvtsd2si under the default rounding mode).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <list>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

typedef unsigned char uchar;

extern "C" {  // the cv2-pinned restatements (oracle/orb_oracle.cc)
float orc_fast_atan2(float y, float x);
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int th, int* xs, int* ys, int* resp, int cap);
void orc_gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
}

inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline int cvRound(float v) { return (int)std::nearbyintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvFloor(float v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline int cvCeil(float v) { return (int)std::ceil(v); }

namespace cv {
using ::cvRound;
using ::cvFloor;
using ::cvCeil;

template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <class U>
  Point_(U x_, U y_) : x((T)x_), y((T)y_) {}  // Point2i(float, float) truncates like saturate_cast does not: see note
  Point_& operator*=(float s) {
    x = (T)(x * s);
    y = (T)(y * s);
    return *this;
  }
};
// NOTE: cv::Point2i(float, float) does not exist in OpenCV either — the reference's `cv::Point2i(hX * float(i), 0)`
// (src/ORBextractor.cc:536-537) converts the float argument to int implicitly at the call (truncation), which is what
// the converting constructor above does.
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Range {
  int start, end;
  Range(int s, int e) : start(s), end(e) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class Mat;
struct MatZeros {  // what Mat::zeros returns (MatExpr in OpenCV): assigned with create()-if-different + fill
  int rows, cols, type;
};

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size sz, int type) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int /*type*/, void* ext, size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : (size_t)c), data((uchar*)ext) {}
  Mat(const Mat& m) : rows(m.rows), cols(m.cols), step(m.step), data(m.data), buf_(m.buf_) { if (buf_) ++*buf_; }
  Mat& operator=(const Mat& m) {
    if (m.buf_) ++*m.buf_;
    unref();
    rows = m.rows; cols = m.cols; step = m.step; data = m.data; buf_ = m.buf_;
    return *this;
  }
  ~Mat() { unref(); }

  void create(int r, int c, int type) {
    assert(type == CV_8UC1);
    (void)type;
    if (data && rows == r && cols == c) return;  // OpenCV: same size and type -> keep the buffer (also for ROI headers)
    unref();
    // pixel buffers come from malloc, never from operator new: libref.so replaces operator new by a monotonic arena
    // (mono_alloc.cc) and the pyramid levels outlive the call that allocated them
    buf_ = (int*)std::malloc(64 + (size_t)r * c + 64);
    *buf_ = 1;
    rows = r; cols = c; step = (size_t)c; data = (uchar*)buf_ + 64;
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { unref(); rows = cols = 0; step = 0; data = nullptr; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  size_t step1() const { return step; }
  Mat getMat() const { return *this; }

  template <class T> T& at(int r, int c) { return *(T*)(data + step * r + c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *(const T*)(data + step * r + c * sizeof(T)); }
  uchar* ptr(int r = 0) { return data + step * r; }
  const uchar* ptr(int r = 0) const { return data + step * r; }
  template <class T> T* ptr(int r = 0) { return (T*)(data + step * r); }
  template <class T> const T* ptr(int r = 0) const { return (const T*)(data + step * r); }

  Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + step * a; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
  Mat row(int r) const { return rowRange(r, r + 1); }
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols);
    return m;
  }
  inline void copyTo(OutputArray dst) const;
  static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
  Mat& operator=(const MatZeros& z) {
    create(z.rows, z.cols, z.type);
    for (int r = 0; r < rows; r++) std::memset(ptr(r), 0, (size_t)cols);
    return *this;
  }

 private:
  void unref() {
    if (buf_ && --*buf_ == 0) std::free(buf_);
    buf_ = nullptr;
  }
  int* buf_ = nullptr;  // malloc'ed block: reference count in the first 64 bytes, pixels after (one thread per extractor)
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_->empty(); }
  Mat getMat() const { return *m_; }
 protected:
  const Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray(Mat& m) : _InputArray(m) {}
  _OutputArray(const Mat& m) : _InputArray(m) {}  // fixed-size destination (e.g. the temporary of descriptors.row(i))
  void create(int r, int c, int type) const { mut().create(r, c, type); }
  void create(Size sz, int type) const { mut().create(sz, type); }
  void release() const { mut().release(); }
  Mat& mut() const { return *const_cast<Mat*>(m_); }
};
inline void Mat::copyTo(OutputArray dst) const {
  dst.create(rows, cols, CV_8UC1);
  Mat& d = dst.mut();
  for (int r = 0; r < rows; r++) std::memmove(d.ptr(r), ptr(r), (size_t)cols);
}

// only referenced by the reference's unused ComputeKeyPointsOld (src/ORBextractor.cc:804-958); never executed here
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& k, int n) {
    if ((int)k.size() <= n) return;
    std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    k.resize(n);
  }
};

enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::FAST(image, keypoints, threshold, nonmaxSuppression) — TYPE_9_16; keypoints in raster order, size 7, response = score
inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
  assert(nonmaxSuppression);
  (void)nonmaxSuppression;
  Mat m = image.getMat();
  keypoints.clear();
  if (m.rows < 7 || m.cols < 7) return;
  int cap = m.rows * m.cols;
  std::vector<int> xs(cap), ys(cap), rs(cap);
  int n = orc_fast_detect(m.data, m.cols, m.rows, (int)m.step, threshold, xs.data(), ys.data(), rs.data(), cap);
  keypoints.reserve(n);
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1.f, (float)rs[i]));
}

inline void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
  assert(fx == 0 && fy == 0 && interpolation == INTER_LINEAR);
  (void)fx; (void)fy; (void)interpolation;
  Mat s = src.getMat();
  dst.create(dsize, CV_8UC1);
  Mat& d = dst.mut();
  orc_resize_linear_u8(s.data, s.cols, s.rows, (int)s.step, d.data, d.cols, d.rows, (int)d.step);
}

inline void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sx, double sy = 0, int borderType = BORDER_DEFAULT) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
  (void)ksize; (void)sx; (void)sy; (void)borderType;
  Mat s = src.getMat();
  Mat tmp = s.clone();  // in-place call in the reference
  dst.create(s.rows, s.cols, CV_8UC1);
  Mat& d = dst.mut();
  orc_gaussian_blur7_u8(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, d.data, (int)d.step);
}

inline int borderInterpolate101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - p - 2;
  return p;
}
// copyMakeBorder, BORDER_REFLECT_101 (ISOLATED or not: the sources the reference passes are either a whole image or an
// ROI it wants treated as isolated).  Handles src being the central ROI of dst (the reference's level >= 1 call).
inline void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  (void)borderType;
  Mat s = src.getMat();
  dst.create(s.rows + top + bottom, s.cols + left + right, CV_8UC1);
  Mat& d = dst.mut();
  Mat sc = s.clone();
  for (int y = 0; y < d.rows; y++) {
    int sy = borderInterpolate101(y - top, sc.rows);
    for (int x = 0; x < d.cols; x++) d.ptr(y)[x] = sc.ptr(sy)[borderInterpolate101(x - left, sc.cols)];
  }
}
}  // namespace cv

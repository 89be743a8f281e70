// ORACLE — TEST INFRASTRUCTURE ONLY.
// Frame::ComputeStereoMatches (src/Frame.cc:451-611) of the REFERENCE compiled UNCHANGED: the row-band table, the Hamming arg-min
// with the octave band and the [uL - maxD, uL] gate, the 11 x 11 L1 search over +-5 px in the pyramid level, the parabola fit, the
// disparity tests and the 1.5 * 1.4 * median filter.  The member-function definition is cut out of the source by name at build
// time (oracle/_ref/gen/stereo_fns.inc) together with ORBmatcher::DescriptorDistance (orbmatcher_fns.inc); this file supplies the
// members of Frame the body reads and a TU-local image type (namespace cvst_st, `#define cv cvst_st` around the cut text: the cvstub
// header of the extractor build is u8-only) with the handful of cv::Mat operations the body uses: ROI headers, convertTo(CV_32F),
// at<float>, Mat::ones, scalar * Mat, Mat - Mat, norm(NORM_L1).  All float values in the correlation are small integers, so every
// one of those operations is exact in any evaluation order.
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>
using namespace std;

#define CV_8U 0
#define CV_32F 5
namespace cvst_st {
struct Point2f {
  float x, y;
};
struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
enum { NORM_L1 = 2 };
class Mat {
 public:
  int rows = 0, cols = 0, type_ = CV_8U;
  size_t step = 0;  // bytes per row
  uint8_t* data = nullptr;
  std::shared_ptr<std::vector<uint8_t>> buf;
  Mat() {}
  Mat(int r, int c, int type, void* ext, size_t step_) : rows(r), cols(c), type_(type), step(step_), data((uint8_t*)ext) {}
  static Mat create(int r, int c, int type) {
    Mat m;
    m.rows = r; m.cols = c; m.type_ = type;
    m.step = (size_t)c * (type == CV_32F ? 4 : 1);
    m.buf = std::make_shared<std::vector<uint8_t>>(m.step * (size_t)r);
    m.data = m.buf->data();
    return m;
  }
  size_t esz() const { return type_ == CV_32F ? 4 : 1; }
  Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + step * (size_t)a; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.data = data + esz() * (size_t)a; m.cols = b - a; return m; }
  Mat row(int r) const { return rowRange(r, r + 1); }
  template <class T> T* ptr(int r = 0) { return (T*)(data + step * (size_t)r); }
  template <class T> const T* ptr(int r = 0) const { return (const T*)(data + step * (size_t)r); }
  template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  void convertTo(Mat& dst, int type) const {  // u8 -> f32 (the only conversion the body asks for); dst may alias *this
    Mat out = create(rows, cols, type);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) out.at<float>(r, c) = type_ == CV_32F ? at<float>(r, c) : (float)at<uint8_t>(r, c);
    dst = out;
  }
  static Mat ones(int r, int c, int type) {
    Mat m = create(r, c, type);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.0f;
    return m;
  }
};
inline Mat operator*(float s, const Mat& a) {
  Mat o = Mat::create(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) o.at<float>(r, c) = s * a.at<float>(r, c);
  return o;
}
inline Mat operator-(const Mat& a, const Mat& b) {
  Mat o = Mat::create(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) o.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c);
  return o;
}
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {
  double s = 0;
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
  return s;
}
}  // namespace cvst_st

#define cv cvst_st
namespace VIEO_SLAM_STEREO {  // (a namespace of its own: ORBmatcher::ComputeThreeMaxima is also compiled in ref_match_wrap.cc)
class ORBmatcher {  // declaration subset of include/ORBmatcher.h
 public:
  static const int TH_LOW, TH_HIGH;
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
const int ORBmatcher::TH_HIGH = 100;  // src/ORBmatcher.cc:20-21
const int ORBmatcher::TH_LOW = 50;
#include "orbmatcher_fns.inc"

struct ORBextractor {
  std::vector<cv::Mat> mvImagePyramid;
};
class Frame {  // the members ComputeStereoMatches touches (include/Frame.h, include/FrameBase.h)
 public:
  struct {
    vector<float> vuright_, vdepth_;
    float baseline_bf_[2];
  } stereoinfo_;
  struct {
    vector<float> vscalefactor_;
  } scalepyrinfo_;
  int N = 0;
  vector<ORBextractor*> mpORBextractors;
  vector<vector<cv::KeyPoint>> vvkeys_;
  vector<float> mvInvScaleFactors;
  cv::Mat mDescriptors;
  vector<cv::Mat> vdescriptors_;
  void ComputeStereoMatches();
};
#include "stereo_fns.inc"
}  // namespace VIEO_SLAM_STEREO
#undef cv

struct RefKp {  // == OrcKeyPoint
  float x, y, size, angle, response;
  int32_t octave;
};
extern "C" int ref_stereo_matches(const RefKp* kl, const uint8_t* dl, int nl, const RefKp* kr, const uint8_t* dr, int nr,
                                  const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int n_levels,
                                  const float* scale, const float* inv_scale, float bf, float minZ, float* uright, float* depth) {
  using namespace VIEO_SLAM_STEREO;
  Frame F;
  ORBextractor eL, eR;
  for (int l = 0; l < n_levels; ++l) {
    eL.mvImagePyramid.push_back(cvst_st::Mat(lh[l], lw[l], CV_8U, (void*)pyrL[l], (size_t)lw[l]));
    eR.mvImagePyramid.push_back(cvst_st::Mat(lh[l], lw[l], CV_8U, (void*)pyrR[l], (size_t)lw[l]));
  }
  F.mpORBextractors = {&eL, &eR};
  F.vvkeys_.resize(2);
  auto fill = [](const RefKp* k, int n, std::vector<cvst_st::KeyPoint>& out) {
    out.resize(n);
    for (int i = 0; i < n; ++i) {
      out[i].pt.x = k[i].x; out[i].pt.y = k[i].y; out[i].size = k[i].size; out[i].angle = k[i].angle;
      out[i].response = k[i].response; out[i].octave = k[i].octave; out[i].class_id = -1;
    }
  };
  fill(kl, nl, F.vvkeys_[0]);
  fill(kr, nr, F.vvkeys_[1]);
  F.N = nl;
  F.mDescriptors = cvst_st::Mat(nl, 32, CV_8U, (void*)dl, 32);
  F.vdescriptors_.resize(2);
  F.vdescriptors_[0] = F.mDescriptors;
  F.vdescriptors_[1] = cvst_st::Mat(nr, 32, CV_8U, (void*)dr, 32);
  F.scalepyrinfo_.vscalefactor_.assign(scale, scale + n_levels);
  F.mvInvScaleFactors.assign(inv_scale, inv_scale + n_levels);
  F.stereoinfo_.baseline_bf_[0] = minZ;  // the body reads [0] as minZ (:476) and [1] as bf
  F.stereoinfo_.baseline_bf_[1] = bf;
  F.ComputeStereoMatches();
  int kept = 0;
  for (int i = 0; i < nl; ++i) {
    uright[i] = F.stereoinfo_.vuright_[i];
    depth[i] = F.stereoinfo_.vdepth_[i];
    kept += uright[i] >= 0;
  }
  return kept;
}

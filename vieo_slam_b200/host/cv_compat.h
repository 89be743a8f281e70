// Minimal stand-ins for the OpenCV types that appear in the reference's ORBextractor / ORBmatcher signatures, used
// only when the real OpenCV headers are absent (this image).  With OpenCV present define VIEO_HAVE_OPENCV and the
// shims compile against <opencv2/core.hpp> instead.
#pragma once
#ifdef VIEO_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <vector>
namespace cv {
struct Point2f {
  float x = 0, y = 0;
};
struct KeyPoint {  // field order of cv::KeyPoint
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};
// CV_8UC1 matrix: either a view on caller memory or an owning buffer (enough for images and descriptor rows)
struct Mat {
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t* data = nullptr;
  std::vector<uint8_t> own;
  Mat() = default;
  Mat(int r, int c, uint8_t* d, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data(d) {}
  void create(int r, int c) {
    rows = r; cols = c; step = (size_t)c;
    own.assign((size_t)r * c, 0);
    data = own.data();
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  uint8_t* ptr(int r) { return data + step * r; }
  const uint8_t* ptr(int r) const { return data + step * r; }
  Mat row(int r) const { return Mat(1, cols, data + step * r, step); }
};
using InputArray = const Mat&;
using OutputArray = Mat&;
}  // namespace cv
#endif

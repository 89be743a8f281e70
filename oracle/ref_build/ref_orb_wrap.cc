// ORACLE — TEST INFRASTRUCTURE ONLY.
// C interface over the REFERENCE's own ORB extractor, compiled UNCHANGED from /root/reference/src/ORBextractor.cc
// (textually included below from where it lies; nothing of it is copied into this repository) against the minimal
// OpenCV stand-in of cvstub/.  What runs here is the reference's control flow — constructor tables (:391-456),
// ComputePyramid (:1060-1081), ComputeKeyPointsOctTree (:723-802), DistributeOctTree / DivideNode (:467-721), IC_Angle
// (:55-80), computeOrbDescriptor (:83-127), operator() incl. the lapping area (:968-1058) — with the OpenCV primitives
// supplied by the cv2-pinned restatements.  Used only by tests/ to pin oracle/orb_oracle.cc ("oracle == _ref").
#include "ORBextractor.cc"  // found through -I/root/reference/src

#include <cstdint>
#include <cstring>

namespace {
struct RefOrb : VIEO_SLAM::ORBextractor {  // derived only to reach the protected members
  using ORBextractor::ORBextractor;
  using ORBextractor::DistributeOctTree;
  using ORBextractor::mnFeaturesPerLevel;
  using ORBextractor::umax;
};
struct RefKeyPoint {  // same six fields as OrcKeyPoint (oracle.h)
  float x, y, size, angle, response;
  int32_t octave;
};
}  // namespace

extern "C" void ref_mono_begin();
extern "C" void ref_mono_end();
namespace {
struct MonoScope {  // address order == creation order inside the call (mono_alloc.cc)
  MonoScope() { ref_mono_begin(); }
  ~MonoScope() { ref_mono_end(); }
};
}  // namespace

extern "C" {
void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new RefOrb(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_orb_destroy(void* h) { delete (RefOrb*)h; }

void ref_orb_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* quota, int* umax) {
  RefOrb* o = (RefOrb*)h;
  int n = o->GetLevels();
  std::vector<float> a = o->GetScaleFactors(), b = o->GetInverseScaleFactors(), c = o->GetScaleSigmaSquares(),
                     d = o->GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; i++) {
    scale[i] = a[i]; invScale[i] = b[i]; sigma2[i] = c[i]; invSigma2[i] = d[i];
    quota[i] = o->mnFeaturesPerLevel[i];
  }
  for (int i = 0; i < 16; i++) umax[i] = o->umax[i];
}

// ORBextractor::operator(); returns its return value (monoIndex / -1), *n = number of keypoints
int ref_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, const int* lapping, RefKeyPoint* kps,
                    uint8_t* desc, int cap, int* n) {
  RefOrb* o = (RefOrb*)h;
  MonoScope mono;  // constructed first, destroyed last: every local below dies inside it
  cv::Mat image = img ? cv::Mat(hgt, w, CV_8UC1, (void*)img, (size_t)stride) : cv::Mat();
  cv::Mat mask, descriptors;
  std::vector<cv::KeyPoint> keypoints;
  std::vector<int> lap;
  if (lapping) lap.assign(lapping, lapping + 2);
  int ret = (*o)(image, mask, keypoints, descriptors, lapping ? &lap : nullptr);
  *n = ret < 0 && image.empty() ? 0 : (int)keypoints.size();
  if (*n > cap) return -2;
  for (int i = 0; i < *n; i++) {
    const cv::KeyPoint& k = keypoints[i];
    kps[i] = RefKeyPoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave};
    std::memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
  }
  return ret;
}

int ref_orb_level_size(void* h, int level, int* w, int* hgt) {
  RefOrb* o = (RefOrb*)h;
  if (level < 0 || level >= o->GetLevels() || o->mvImagePyramid[level].empty()) return -1;
  *w = o->mvImagePyramid[level].cols;
  *hgt = o->mvImagePyramid[level].rows;
  return 0;
}
void ref_orb_get_level(void* h, int level, uint8_t* out) {  // the public mvImagePyramid[level] ROI, tightly packed
  const cv::Mat& m = ((RefOrb*)h)->mvImagePyramid[level];
  for (int r = 0; r < m.rows; r++) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
}

// ORBextractor::DistributeOctTree on (x, y, response) candidates; returns the picked keypoints in list order
int ref_quadtree(void* h, const int* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int level, int* out_xyr,
                 int cap) {
  MonoScope mono;
  std::vector<cv::KeyPoint> in(n);
  for (int i = 0; i < n; i++) in[i] = cv::KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2]);
  std::vector<cv::KeyPoint> res = ((RefOrb*)h)->DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
  if ((int)res.size() > cap) return -1;
  for (size_t i = 0; i < res.size(); i++) {
    out_xyr[3 * i] = (int)res[i].pt.x; out_xyr[3 * i + 1] = (int)res[i].pt.y; out_xyr[3 * i + 2] = (int)res[i].response;
  }
  return (int)res.size();
}

// IC_Angle (static in the reference TU) on a tightly described u8 image
float ref_ic_angle(void* h, const uint8_t* img, int w, int hgt, int stride, float x, float y) {
  cv::Mat m(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
  return VIEO_SLAM::IC_Angle(m, cv::Point2f(x, y), ((RefOrb*)h)->umax);
}
// computeOrbDescriptor (static in the reference TU) on an already blurred image
void ref_orb_descriptor(const uint8_t* blurred, int w, int hgt, int stride, float x, float y, float angle_deg, uint8_t* desc) {
  cv::Mat m(hgt, w, CV_8UC1, (void*)blurred, (size_t)stride);
  cv::KeyPoint k(x, y, 31.f, angle_deg);
  VIEO_SLAM::computeOrbDescriptor(k, m, (const cv::Point*)VIEO_SLAM::bit_pattern_31_, desc);
}
}

// ORACLE — TEST INFRASTRUCTURE ONLY.
// Frame::ComputeStereoFishEyeMatches (src/Frame.cc:613-779) of the REFERENCE compiled UNCHANGED (cut out by name at build time): the
// camera-pair loop with its rowRange(num_mono) in-area slices and the "no in-area row" skip, cv::BFMatcher::knnMatch(k = 2), Lowe's
// ratio test in its float / double mix (`size >= 2 && (d0 < d1 * 0.7 || (d0 < 75 && d0 < d1 * 0.9))`), the num_mono offsets added
// back to the indices handed on, the second pass with the looser disparity threshold when fewer than 30 matches were accepted, and the
// concatenation of the per-camera keypoints / descriptors (mvKeys, mDescriptors, mapn2in_, mapin2n_, N).
// Stand-ins, stated for what they are: cv::BFMatcher::knnMatch forwards to the cv2-pinned restatement (../match_oracle.cc,
// orc_hamming_knn2, pinned against cv2.BFMatcher(NORM_HAMMING).knnMatch by tests/golden/cv2_orb_goldens.npz); the camera's
// FillMatchesFromPair is a RECORDING stand-in that accepts every pair (the reference's own is compiled in ref_sft_wrap.cc; here its
// triangulation test would need KB8 UnProject + Eigen's SVD) — what is compared is the sequence of (camera, keypoint, camera, keypoint,
// distance) tuples the function hands to it, i.e. exactly the output of the device's / oracle's brute-force half.
#include <math.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <memory>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "mini_eigen.h"
#include "sophus/se3.hpp"
#include "common/unordered_hash.h"  // the reference's, unchanged

using namespace std;

#define PRINT_DEBUG_FILE_MUTEX(...)
#define PRINT_DEBUG_FILE(...)
#define USE_STRATEGY_MIN_DIST  // common/config.h:10-13

extern "C" void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);  // ../match_oracle.cc

namespace Eigen {
template <class T>
using aligned_vector = std::vector<T>;
}

namespace cvst_fe {
struct Point2f {
  float x, y;
};
struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
};
class Mat {  // CV_8U descriptor rows (32 bytes each); a shared buffer when it owns one (clone / vconcat)
 public:
  const uint8_t* data = nullptr;
  int rows = 0;
  std::shared_ptr<std::vector<uint8_t>> own;
  Mat() {}
  Mat(const uint8_t* d, int r) : data(d), rows(r) {}
  bool empty() const { return rows == 0; }
  Mat rowRange(int a, int b) const { return Mat(data + 32 * (size_t)a, b - a); }
  Mat clone() const {
    Mat o;
    o.own = std::make_shared<std::vector<uint8_t>>(data, data + 32 * (size_t)rows);
    o.data = o.own->data();
    o.rows = rows;
    return o;
  }
};
inline void vconcat(const Mat& a, const Mat& b, Mat& out) {
  auto buf = std::make_shared<std::vector<uint8_t>>();
  buf->insert(buf->end(), a.data, a.data + 32 * (size_t)a.rows);
  buf->insert(buf->end(), b.data, b.data + 32 * (size_t)b.rows);
  Mat o;
  o.own = buf;
  o.data = buf->data();
  o.rows = a.rows + b.rows;
  out = o;
}
enum { NORM_HAMMING = 6 };
class BFMatcher {
 public:
  BFMatcher(int = NORM_HAMMING) {}
  // cv::DescriptorMatcher::knnMatch: one vector per query row with min(k, train rows) matches, nearest first (ties: lower train index)
  void knnMatch(const Mat& q, const Mat& t, vector<vector<DMatch>>& matches, int k) const {
    assert(k == 2);
    vector<int32_t> idx(2 * (size_t)q.rows), dist(2 * (size_t)q.rows);
    orc_hamming_knn2(q.data, q.rows, t.data, t.rows, idx.data(), dist.data());
    matches.assign(q.rows, vector<DMatch>());
    for (int r = 0; r < q.rows; ++r)
      for (int c = 0; c < 2; ++c)
        if (idx[2 * r + c] >= 0) matches[r].push_back(DMatch{r, idx[2 * r + c], 0, (float)dist[2 * r + c]});
  }
};
}  // namespace cvst_fe
#define cv cvst_fe

namespace VIEO_SLAM_FE {
using Eigen::Vector2f;
using Eigen::Vector3d;
template <class T>
using aligned_vector = Eigen::aligned_vector<T>;

struct RecordedPair {
  int32_t cami, idxi, camj, idxj;
  float dist;
};

namespace camm {
class Camera {
 public:
  using Ptr = std::shared_ptr<Camera>;
  using Mat3data = Eigen::Matrix<float, 3, 3>;
  using MapCamIdx2Idx = std::unordered_map<std::pair<size_t, size_t>, size_t, PairHash>;
  float fx = 190.f, fy = 190.f, cx = 256.f, cy = 256.f;
  Sophus::SE3<float> Tcr_;
  std::vector<RecordedPair>* rec = nullptr;
  const Sophus::SE3<float>& GetTcr() const { return Tcr_; }
  Mat3data toK() const {
    Mat3data K;
    K << fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f;
    return K;
  }
  // recording stand-in: accepts every pair, keeps no bookkeeping
  bool FillMatchesFromPair(const vector<const Camera*>& pcams, size_t n_cams_tot, const vector<pair<size_t, size_t>>& vcamidx, float dist,
                           vector<vector<size_t>>&, vector<bool>&, MapCamIdx2Idx&, const float, aligned_vector<Vector3d>*,
                           aligned_vector<Vector2f>* pkpts, vector<float>* psigmas, vector<vector<float>>*, int* pcount_descmatch) const {
    assert(pcams.size() == 1 && vcamidx.size() == 2 && pkpts && pkpts->size() == 2 && psigmas && psigmas->size() == 2);
    rec->push_back(RecordedPair{(int32_t)vcamidx[0].first, (int32_t)vcamidx[0].second, (int32_t)vcamidx[1].first, (int32_t)vcamidx[1].second, dist});
    if (pcount_descmatch) ++*pcount_descmatch;
    return true;
  }
  vector<float> TriangulateMatches(const vector<const Camera*>&, const aligned_vector<Vector2f>&, const vector<float>&, Vector3d*, double) const {
    abort();  // mvidxsMatches stays empty with the recording stand-in: the n_cams > 2 re-triangulation loop has nothing to visit
  }
};
}  // namespace camm

struct ORBmatcher {
  static const int TH_LOW = 50, TH_HIGH = 100;  // src/ORBmatcher.cc:20-21
};

class Frame {  // the members ComputeStereoFishEyeMatches touches (include/Frame.h:128-139, include/FrameBase.h:160-176)
 public:
  vector<size_t> num_mono;
  vector<vector<cv::KeyPoint>> vvkeys_;
  vector<cv::Mat> vdescriptors_;
  static cv::BFMatcher BFmatcher;
  vector<vector<size_t>> mvidxsMatches;
  vector<size_t> mapidxs2n_;
  vector<vector<size_t>> mapin2n_;
  vector<camm::Camera::Ptr> mpCameras;
  struct {
    vector<float> vdepth_, vuright_;
    aligned_vector<Vector3d> v3dpoints_;
    vector<bool> goodmatches_;
    camm::Camera::MapCamIdx2Idx mapcamidx2idxs_;
    float baseline_bf_[2] = {15.f / 250, 15.f};
  } stereoinfo_;
  struct {
    vector<float> vlevelsigma2_;
  } scalepyrinfo_;
  cv::Mat mDescriptors;
  vector<cv::KeyPoint> mvKeys;
  vector<pair<size_t, size_t>> mapn2in_;
  int N = 0;
  void ComputeStereoFishEyeMatches(const float th_far_pts = 0);
};
cv::BFMatcher Frame::BFmatcher = cv::BFMatcher(cv::NORM_HAMMING);  // src/Frame.cc: static member
#include "fisheye_fns.inc"
}  // namespace VIEO_SLAM_FE
#undef cv

// desc: n_cams blocks of `cap` rows (like orc_fisheye_matches); octave [n_cams][cap] (read for the sigma handed on).
// rec [rec_cap][5] = (cami, idxi, camj, idxj, dist) in call order; returns the number of recorded calls.  n_out = N;
// map_cam / map_idx [sum n_kp] = mapn2in_; desc_out [sum n_kp][32] = mDescriptors.
extern "C" int ref_fisheye_matches(const uint8_t* desc, const int32_t* octave, const int32_t* n_kp, const int32_t* n_mono, int n_cams, int cap,
                                   float th_far_pts, int32_t* rec, int rec_cap, int32_t* n_out, int32_t* map_cam, int32_t* map_idx,
                                   uint8_t* desc_out) {
  using namespace VIEO_SLAM_FE;
  Frame F;
  std::vector<RecordedPair> recorded;
  F.num_mono.assign(n_mono, n_mono + n_cams);
  F.vvkeys_.resize(n_cams);
  F.vdescriptors_.resize(n_cams);
  F.scalepyrinfo_.vlevelsigma2_.resize(8);
  for (int l = 0; l < 8; ++l) F.scalepyrinfo_.vlevelsigma2_[l] = std::pow(1.2f, 2.f * l);
  for (int c = 0; c < n_cams; ++c) {
    auto cam = std::make_shared<camm::Camera>();
    cam->rec = &recorded;
    F.mpCameras.push_back(cam);
    F.vvkeys_[c].resize(n_kp[c]);
    for (int k = 0; k < n_kp[c]; ++k) {
      F.vvkeys_[c][k].pt.x = (float)k;
      F.vvkeys_[c][k].pt.y = (float)c;
      F.vvkeys_[c][k].octave = octave[(size_t)c * cap + k];
    }
    F.vdescriptors_[c] = cvst_fe::Mat(desc + (size_t)c * cap * 32, n_kp[c]);
  }
  F.ComputeStereoFishEyeMatches(th_far_pts);
  const int n = (int)recorded.size();
  for (int i = 0; i < n && i < rec_cap; ++i) {
    rec[5 * i] = recorded[i].cami; rec[5 * i + 1] = recorded[i].idxi; rec[5 * i + 2] = recorded[i].camj; rec[5 * i + 3] = recorded[i].idxj;
    rec[5 * i + 4] = (int32_t)recorded[i].dist;
  }
  *n_out = F.N;
  for (int k = 0; k < F.N; ++k) {
    map_cam[k] = (int32_t)F.mapn2in_[k].first;
    map_idx[k] = (int32_t)F.mapn2in_[k].second;
    assert(F.mvKeys[k].pt.x == (float)map_idx[k] && F.mvKeys[k].pt.y == (float)map_cam[k]);
    assert(F.mapin2n_[map_cam[k]][map_idx[k]] == (size_t)k);
  }
  if (F.N) memcpy(desc_out, F.mDescriptors.data, 32 * (size_t)F.N);
  return n;
}

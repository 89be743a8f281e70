// C++ host shims with the reference's class surfaces over the C ABI of libvieo_b200.so (include/vieo_b200.h).
// A VIEO_SLAM build replaces the bodies of src/ORBextractor.cc, ORBmatcher::DescriptorDistance, the knnMatch call of
// Frame::ComputeStereoFishEyeMatches, IMUPreIntegratorBase::PreIntegration and the optimizer.optimize() sections of
// src/Optimizer.cc / include/Optimizer.h with these (INTEGRATION.md shows the exact edits).  No CPU fallback: every
// call throws std::runtime_error with vieo_last_error() when the library reports a failure.
#pragma once
#include <algorithm>
#include <cstring>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vieo_b200.h"
#include "cv_compat.h"

namespace VIEO_SLAM_B200 {

inline void vieo_check(int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + vieo_last_error());
}

// ORBextractor (include/ORBextractor.h:27-80).  One instance per camera, driven by one thread at a time
// (src/Frame.cc:259-278); the image size is fixed at the first call (cameras do not change size).
class ORBextractor {
 public:
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
      : cfg_{0, 0, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, 1}, device_(device) {}
  ~ORBextractor() { vieo_orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // Returns monoIndex, or -1 for an empty image (src/ORBextractor.cc:968-1058).  `mask` is ignored like upstream.
  int operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                 cv::OutputArray descriptors, const std::vector<int>* pvLappingArea = nullptr) {
    if (image.empty()) return -1;
    ensure(image.cols, image.rows);
    std::vector<uint8_t*> pyr(cfg_.nlevels);
    for (int l = 0; l < cfg_.nlevels; ++l) {
      mvImagePyramid[l].create(level_h_[l], level_w_[l]);
      pyr[l] = mvImagePyramid[l].data;
    }
    const int32_t* lap = (pvLappingArea && pvLappingArea->size() >= 2) ? pvLappingArea->data() : nullptr;
    int32_t mono = 0;
    kp_.resize(cap_);
    desc_.resize((size_t)cap_ * 32);
    const int n = vieo_orb_extract(h_, image.data, (int)image.step, lap, kp_.data(), desc_.data(), cap_, &mono, pyr.data());
    if (n == VIEO_E_EMPTY) return -1;
    vieo_check(n, "vieo_orb_extract");
    keypoints.resize(n);
    for (int i = 0; i < n; ++i) {
      cv::KeyPoint& k = keypoints[i];
      k.pt.x = kp_[i].x; k.pt.y = kp_[i].y; k.size = kp_[i].size; k.angle = kp_[i].angle;
      k.response = kp_[i].response; k.octave = kp_[i].octave; k.class_id = -1;
    }
    descriptors.create(n, 32);
    if (n) std::memcpy(descriptors.data, desc_.data(), (size_t)n * 32);
    return mono;
  }
  int GetLevels() const { return cfg_.nlevels; }
  float GetScaleFactor() const { return cfg_.scale_factor; }
  const std::vector<float>& GetScaleFactors() { tables(); return mvScaleFactor; }
  const std::vector<float>& GetInverseScaleFactors() { tables(); return mvInvScaleFactor; }
  const std::vector<float>& GetScaleSigmaSquares() { tables(); return mvLevelSigma2; }
  const std::vector<float>& GetInverseScaleSigmaSquares() { tables(); return mvInvLevelSigma2; }
  std::vector<cv::Mat> mvImagePyramid;  // public in the reference; read by Frame::ComputeStereoMatches

 private:
  void ensure(int w, int h) {
    if (h_ && (w != cfg_.width || h != cfg_.height)) {
      vieo_orb_destroy(h_);
      h_ = nullptr;
    }
    if (h_) return;
    cfg_.width = w; cfg_.height = h;
    vieo_check(vieo_orb_create(&cfg_, device_, &h_), "vieo_orb_create");
    cap_ = vieo_orb_max_keypoints(h_);
    tables();
    mvImagePyramid.resize(cfg_.nlevels);
  }
  void tables() {
    // scale tables do not depend on the image size; a 64x64 probe handle serves the getters before the first frame
    const int n = cfg_.nlevels;
    if ((int)mvScaleFactor.size() == n && (h_ == nullptr || !level_w_.empty())) return;
    vieo_orb_t* t = h_;
    VieoOrbConfig c = cfg_;
    if (!t) {
      c.width = c.height = 256;
      vieo_check(vieo_orb_create(&c, device_, &t), "vieo_orb_create");
    }
    mvScaleFactor.resize(n); mvInvScaleFactor.resize(n); mvLevelSigma2.resize(n); mvInvLevelSigma2.resize(n);
    std::vector<int32_t> quota(n), lw(n), lh(n);
    vieo_check(vieo_orb_get_tables(t, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                   mvInvLevelSigma2.data(), quota.data(), lw.data(), lh.data()), "vieo_orb_get_tables");
    if (t == h_) { level_w_ = lw; level_h_ = lh; } else vieo_orb_destroy(t);
  }
  VieoOrbConfig cfg_;
  int device_, cap_ = 0;
  vieo_orb_t* h_ = nullptr;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<int32_t> level_w_, level_h_;
  std::vector<VieoKeyPoint> kp_;
  std::vector<uint8_t> desc_;
};

// The Hamming kernels behind ORBmatcher (include/ORBmatcher.h:18-113) and the BFMatcher call of src/Frame.cc:620-628
struct DMatch2 {
  int32_t idx[2], dist[2];
};
class ORBmatcher {
 public:
  static constexpr int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;  // src/ORBmatcher.cc:20-22
  explicit ORBmatcher(float nnratio = 0.6f, bool checkOri = true, int device = 0)
      : mfNNratio(nnratio), mbCheckOrientation(checkOri), device_(device) {}
  // cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k = 2)
  std::vector<DMatch2> knnMatch2(const cv::Mat& q, const cv::Mat& t) const {
    std::vector<int32_t> idx((size_t)2 * q.rows), dist((size_t)2 * q.rows);
    vieo_check(vieo_hamming_knn2(q.data, q.rows, t.data, t.rows, idx.data(), dist.data(), device_), "vieo_hamming_knn2");
    std::vector<DMatch2> out(q.rows);
    for (int i = 0; i < q.rows; ++i) out[i] = {{idx[2 * i], idx[2 * i + 1]}, {dist[2 * i], dist[2 * i + 1]}};
    return out;
  }
  // best / second-best over candidate lists (the inner loops of SearchByProjection & co.), CSR rows
  void SearchCandidates(const cv::Mat& q, const cv::Mat& t, const std::vector<int32_t>& row_ptr,
                        const std::vector<int32_t>& cand, std::vector<int32_t>& best, std::vector<int32_t>& best_idx,
                        std::vector<int32_t>& second, std::vector<int32_t>& second_idx) const {
    const int n = (int)row_ptr.size() - 1;
    best.resize(n); best_idx.resize(n); second.resize(n); second_idx.resize(n);
    vieo_check(vieo_hamming_csr(q.data, t.data, t.rows, row_ptr.data(), cand.data(), n, best.data(), best_idx.data(),
                                second.data(), second_idx.data(), device_), "vieo_hamming_csr");
  }
  // SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist, th_far_pts) (src/ORBmatcher.cc:
  // 1471-1606), Tracking::Relocalization's guided search, one record per (candidate keyframe, current-frame copy) pair.
  // Queries = pKF's map points that are good and not in sAlreadyFound, in keypoint order.  th goes into
  // frames[f].th, ORBdist into reloc[f].orb_dist.  Returns nmatches summed over the records.
  int SearchByProjectionReloc(std::vector<VieoSbpFrame>& frames, const std::vector<VieoSbpReloc>& reloc, const VieoKeyPoint* keys_un,
                              const uint8_t* descriptors, const double* q_Xw, const float* q_angle, const float* q_max_dist,
                              const float* q_min_dist, const uint8_t* q_desc, const uint8_t* kp_blocked,
                              std::vector<int32_t>& kp_match, std::vector<int32_t>& q_match, std::vector<int32_t>& q_dist,
                              std::vector<int32_t>& n_matches) const {
    size_t nk = 0, nq = 0;
    for (VieoSbpFrame& f : frames) {
      f.check_orientation = mbCheckOrientation ? 1 : 0;
      nk = std::max(nk, (size_t)f.kp_begin + f.n_kp);
      nq = std::max(nq, (size_t)f.q_begin + f.n_q);
    }
    kp_match.assign(nk, -1); q_match.assign(nq, -1); q_dist.assign(nq, -1); n_matches.assign(frames.size(), 0);
    vieo_check(vieo_sbp_reloc_batch(frames.data(), reloc.data(), (int)frames.size(), keys_un, descriptors, q_Xw, q_angle, q_max_dist,
                                    q_min_dist, q_desc, kp_blocked, kp_match.data(), q_match.data(), q_dist.data(), nullptr,
                                    n_matches.data(), device_), "vieo_sbp_reloc_batch");
    int total = 0;
    for (int v : n_matches) total += v;
    return total;
  }
  // The tracking thread's SearchByProjection overloads on the flattened frame(s) (src/ORBmatcher.cc:230-335,
  // 1303-1467; INTEGRATION.md 2b shows how Frame / MapPoint fields map onto the arrays).  mode = VIEO_SBP_LAST_FRAME or
  // VIEO_SBP_LOCAL_MAP; the matcher's mfNNratio / mbCheckOrientation are written into every frame record.
  // Returns the reference's nmatches summed over the frames; n_matches gets the per-frame values.
  int SearchByProjection(int mode, std::vector<VieoSbpFrame>& frames, const VieoKeyPoint* keys_un, const float* vuright,
                         const uint8_t* descriptors, const VieoSbpQueries& queries, const uint8_t* kp_blocked,
                         std::vector<int32_t>& kp_match, std::vector<int32_t>& q_match, std::vector<int32_t>& q_dist,
                         std::vector<int32_t>& n_matches) const {
    size_t nk = 0, nq = 0;
    for (VieoSbpFrame& f : frames) {
      f.nn_ratio = mfNNratio;
      f.check_orientation = mbCheckOrientation ? 1 : 0;
      nk = std::max(nk, (size_t)f.kp_begin + f.n_kp);
      nq = std::max(nq, (size_t)f.q_begin + f.n_q);
    }
    kp_match.assign(nk, -1); q_match.assign(nq, -1); q_dist.assign(nq, -1); n_matches.assign(frames.size(), 0);
    vieo_check(vieo_sbp_batch(mode, frames.data(), (int)frames.size(), keys_un, vuright, descriptors, &queries, kp_blocked,
                              kp_match.data(), q_match.data(), q_dist.data(), n_matches.data(), device_), "vieo_sbp_batch");
    int total = 0;
    for (int32_t n : n_matches) total += n;
    return total;
  }
  // Tracking::SearchLocalPoints (src/Tracking.cc:2308-2368): Frame::isInFrustum(pMP, 0.5) for every candidate local map
  // point, then SearchByProjection(F, vpMapPoints, th, th_far) over those in view; one call, one stream.  The caller
  // copies inview / proj / level / viewcos / depth back into MapPoint::trackinfo_ only if it needs them (IncreaseVisible
  // uses `inview`).  Returns nmatches summed over the frames.
  int SearchLocalPoints(std::vector<VieoFrustumFrame>& frustum, std::vector<VieoSbpFrame>& frames, const float* wP,
                        const float* normal, const float* max_dist, const float* min_dist, const uint8_t* skip,
                        const uint8_t* q_desc, const uint8_t* q_flags, const VieoKeyPoint* keys_un, const float* vuright,
                        const uint8_t* descriptors, const uint8_t* kp_blocked, std::vector<uint8_t>& inview,
                        std::vector<float>& proj, std::vector<int32_t>& level, std::vector<float>& viewcos,
                        std::vector<float>& depth, std::vector<int32_t>& n_inview, std::vector<int32_t>& kp_match,
                        std::vector<int32_t>& q_match, std::vector<int32_t>& q_dist, std::vector<int32_t>& n_matches) const {
    size_t nk = 0, nq = 0;
    for (VieoSbpFrame& f : frames) {
      f.nn_ratio = mfNNratio;
      f.check_orientation = mbCheckOrientation ? 1 : 0;
      nk = std::max(nk, (size_t)f.kp_begin + f.n_kp);
      nq = std::max(nq, (size_t)f.q_begin + f.n_q);
    }
    inview.assign(nq, 0); proj.assign(3 * nq, 0.f); level.assign(nq, -1); viewcos.assign(nq, 0.f); depth.assign(nq, 0.f);
    n_inview.assign(frames.size(), 0);
    kp_match.assign(nk, -1); q_match.assign(nq, -1); q_dist.assign(nq, -1); n_matches.assign(frames.size(), 0);
    vieo_check(vieo_search_local_points(frustum.data(), frames.data(), (int)frames.size(), wP, normal, max_dist, min_dist,
                                        skip, q_desc, q_flags, keys_un, vuright, descriptors, kp_blocked, inview.data(),
                                        proj.data(), level.data(), viewcos.data(), depth.data(), n_inview.data(),
                                        kp_match.data(), q_match.data(), q_dist.data(), n_matches.data(), device_),
               "vieo_search_local_points");
    int total = 0;
    for (int32_t n : n_matches) total += n;
    return total;
  }
  // The search half of SearchByProjectionBase (src/ORBmatcher.cc:26-227) for a batch of keyframes — what Fuse(pKF,
  // vpMapPoints, th) runs before FuseMP: best_idx[i] >= 0 && best_dist[i] <= TH_LOW are the points to fuse / add.
  void SearchByProjectionBase(const std::vector<VieoProjSearchFrame>& kfs, const VieoKeyPoint* keys_un, const float* vuright,
                              const uint8_t* descriptors, const float* wP, const float* normal, const float* max_dist,
                              const float* min_dist, const uint8_t* q_desc, const uint8_t* q_skip,
                              std::vector<int32_t>& best_idx, std::vector<int32_t>& best_dist,
                              std::vector<int32_t>& level) const {
    size_t nq = 0;
    for (const VieoProjSearchFrame& f : kfs) nq = std::max(nq, (size_t)f.q_begin + f.n_q);
    best_idx.assign(nq, -1); best_dist.assign(nq, 256); level.assign(nq, -1);
    vieo_check(vieo_proj_search_batch(kfs.data(), (int)kfs.size(), keys_un, vuright, descriptors, wP, normal, max_dist,
                                      min_dist, q_desc, q_skip, best_idx.data(), best_dist.data(), level.data(), device_),
               "vieo_proj_search_batch");
  }
  // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:314-378) for every map point LocalMapping touched in one
  // call: ptr / rows = CSR lists of the observed descriptor rows in desc_pool; best[p] indexes point p's list.
  void ComputeDistinctiveDescriptors(const uint8_t* desc_pool, int n_pool, const std::vector<int32_t>& rows,
                                     const std::vector<int32_t>& ptr, std::vector<int32_t>& best,
                                     std::vector<int32_t>& median) const {
    const int n = (int)ptr.size() - 1;
    best.assign(std::max(n, 0), -1); median.assign(std::max(n, 0), -1);
    if (n <= 0) return;
    vieo_check(vieo_distinctive_descriptors(desc_pool, n_pool, rows.empty() ? nullptr : rows.data(), ptr.data(), n,
                                            best.data(), median.data(), device_), "vieo_distinctive_descriptors");
  }
  // ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:896-1150) for a batch
  // of keyframe pairs (LocalMapping::CreateNewMapPoints loops over the neighbours, src/LocalMapping.cc:680-709); the
  // caller flattens mvKeysUn / vuright_ / mDescriptors / GetMapPoint() != nullptr and the two DBoW2::FeatureVectors per
  // pair and forms F12 and the epipole.  pairs_out holds vMatchedPairs of pair p at
  // [pairs[p].out_begin, + n_matches[p]) as (idx1, idx2); returns the sum of the return values.
  int SearchForTriangulation(std::vector<VieoSftPair>& pairs, const std::vector<VieoKeyPoint>& keys_un,
                             const std::vector<float>& vuright, const std::vector<uint8_t>& descriptors,
                             const std::vector<uint8_t>& has_mp, const std::vector<int32_t>& fv_node,
                             const std::vector<int32_t>& fv_ptr, const std::vector<int32_t>& fv_idx,
                             std::vector<int32_t>& match12, std::vector<int32_t>& pairs_out, std::vector<int32_t>& n_matches,
                             bool bOnlyStereo = false) const {
    int n_out = 0, n_nodes1 = 0;
    for (VieoSftPair& p : pairs) {
      p.out_begin = n_out; p.nscr_begin = n_nodes1;
      n_out += p.n_kp1; n_nodes1 += p.n_nodes1;
      p.only_stereo = bOnlyStereo ? 1 : 0;
      p.check_orientation = mbCheckOrientation ? 1 : 0;
    }
    match12.assign(std::max(n_out, 1), -1); pairs_out.assign((size_t)2 * std::max(n_out, 1), -1); n_matches.assign(pairs.size(), 0);
    vieo_check(vieo_search_for_triangulation(pairs.data(), (int)pairs.size(), keys_un.data(), vuright.data(), descriptors.data(),
                                             has_mp.data(), fv_node.data(), fv_ptr.data(), fv_idx.data(), (int)keys_un.size(),
                                             (int)fv_node.size(), (int)fv_ptr.size(), (int)fv_idx.size(), n_out, n_nodes1,
                                             match12.data(), pairs_out.data(), n_matches.data(), device_),
               "vieo_search_for_triangulation");
    int total = 0;
    for (int32_t n : n_matches) total += n;
    return total;
  }
  // ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:344-505) for a batch of (keyframe, frame) pairs:
  // match_f[pairs[p].out_begin + i] = keyframe keypoint whose map point frame keypoint i received (-1: none).
  int SearchByBoW(std::vector<VieoBowPair>& pairs, const std::vector<VieoKeyPoint>& keys, const std::vector<uint8_t>& descriptors,
                  const std::vector<uint8_t>& mp_ok, const std::vector<int32_t>& fv_node, const std::vector<int32_t>& fv_ptr,
                  const std::vector<int32_t>& fv_idx, std::vector<int32_t>& match_f, std::vector<int32_t>& n_matches) const {
    int n_out = 0;
    for (VieoBowPair& p : pairs) {
      p.out_begin = n_out;
      n_out += p.n_kp2;
      p.nn_ratio = mfNNratio;
      p.check_orientation = mbCheckOrientation ? 1 : 0;
    }
    match_f.assign(std::max(n_out, 1), -1); n_matches.assign(pairs.size(), 0);
    vieo_check(vieo_search_by_bow(pairs.data(), (int)pairs.size(), keys.data(), descriptors.data(), mp_ok.data(), fv_node.data(),
                                  fv_ptr.data(), fv_idx.data(), (int)keys.size(), (int)fv_node.size(), (int)fv_ptr.size(),
                                  (int)fv_idx.size(), n_out, match_f.data(), n_matches.data(), device_), "vieo_search_by_bow");
    int total = 0;
    for (int32_t n : n_matches) total += n;
    return total;
  }
  float mfNNratio;
  bool mbCheckOrientation;

 private:
  int device_;
};

// Frame::Frame's per-camera extraction with the lapping area + the brute-force half of
// Frame::ComputeStereoFishEyeMatches (src/Frame.cc:259-278, 613-663) for one multi-camera frame: n_cams images of the
// extractor's size, lapping = n_cams x {x0, x1} (nullptr: none).  vvkeys / vdescriptors / num_mono as the reference's
// members; allmatches[p] = knnMatch(k = 2) rows of camera pair p (i < j in the reference's order) with `good` = passed the
// ratio test (:659-663).
struct FisheyePairMatches {
  std::vector<DMatch2> knn;
  std::vector<uint8_t> good;
};
inline void ExtractAndMatchMultiCam(vieo_orb_t* extractor, int n_cams, const uint8_t* imgs, size_t img_stride, int row_stride,
                                    const int32_t* lapping, std::vector<std::vector<VieoKeyPoint>>& vvkeys,
                                    std::vector<std::vector<uint8_t>>& vdescriptors, std::vector<int>& num_mono,
                                    std::vector<FisheyePairMatches>& allmatches) {
  const int cap = vieo_orb_max_keypoints(extractor), n_pairs = n_cams * (n_cams - 1) / 2;
  std::vector<VieoKeyPoint> k((size_t)n_cams * cap);
  std::vector<uint8_t> d((size_t)n_cams * cap * 32), pg((size_t)std::max(n_pairs, 1) * cap);
  std::vector<int32_t> nk(n_cams), nm(n_cams), pi((size_t)std::max(n_pairs, 1) * cap * 2), pd(pi.size());
  vieo_check(vieo_multicam_frames(extractor, 1, n_cams, imgs, img_stride, row_stride, lapping, k.data(), d.data(), nk.data(),
                                  nm.data(), pi.data(), pd.data(), pg.data()), "vieo_multicam_frames");
  vvkeys.resize(n_cams); vdescriptors.resize(n_cams); num_mono.assign(nm.begin(), nm.end());
  for (int c = 0; c < n_cams; ++c) {
    vvkeys[c].assign(k.begin() + (size_t)c * cap, k.begin() + (size_t)c * cap + nk[c]);
    vdescriptors[c].assign(d.begin() + (size_t)c * cap * 32, d.begin() + ((size_t)c * cap + nk[c]) * 32);
  }
  allmatches.assign(n_pairs, FisheyePairMatches());
  int p = 0;
  for (int i = 0; i < n_cams - 1; ++i)
    for (int j = i + 1; j < n_cams; ++j, ++p) {
      if (nm[i] >= nk[i] || nm[j] >= nk[j]) continue;  // the pair is skipped (:623)
      const int nq = nk[i] - nm[i];
      allmatches[p].knn.resize(nq); allmatches[p].good.resize(nq);
      for (int r = 0; r < nq; ++r) {
        const size_t o = ((size_t)p * cap + r);
        allmatches[p].knn[r] = {{pi[2 * o], pi[2 * o + 1]}, {pd[2 * o], pd[2 * o + 1]}};
        allmatches[p].good[r] = pg[o];
      }
    }
}

// IMUPreIntegratorBase<IMUDataBase> (src/Odom/OdomPreIntegrator.h:108-223): public members kept, PreIntegration on
// the device.  Sample = {t, a, w}.
struct IMUSample {
  double t, a[3], w[3];
};
class IMUPreintegrator {
 public:
  double mRij[9], mvij[3], mpij[3], mSigmaijPRV[81], mSigmaij[81], mJgpij[9], mJapij[9], mJgvij[9], mJavij[9], mJgRij[9];
  double mdeltatij = 0;
  static VieoImuNoise& Noise() {
    static VieoImuNoise nz = {};
    return nz;
  }
  // IMUDataBase::SetParam (src/Odom/OdomData.h:41-56)
  static void SetParam(const double sigma2[4], int dt_cov_noise_fixed, double freq_ref) {
    vieo_imu_set_param(&Noise(), sigma2, dt_cov_noise_fixed, freq_ref);
  }
  template <class It>
  int PreIntegration(double ti, double tj, const double bg[3], const double ba[3], It iterBegin, It iterEnd, int device = 0) {
    std::vector<double> smp;
    for (It it = iterBegin; it != iterEnd; ++it) {
      smp.push_back(it->t);
      smp.insert(smp.end(), it->a, it->a + 3);
      smp.insert(smp.end(), it->w, it->w + 3);
    }
    const int32_t seg[2] = {0, (int32_t)(smp.size() / 7)};
    const double tt[2] = {ti, tj}, bb[6] = {bg[0], bg[1], bg[2], ba[0], ba[1], ba[2]};
    VieoImuPreint o;
    vieo_check(vieo_imu_preint_batch(smp.data(), seg, tt, bb, &Noise(), 1, &o, device), "vieo_imu_preint_batch");
    std::memcpy(mRij, o.Rij, 72); std::memcpy(mvij, o.vij, 24); std::memcpy(mpij, o.pij, 24);
    std::memcpy(mSigmaijPRV, o.SigmaPRV, 648); std::memcpy(mSigmaij, o.SigmaPVR, 648);
    std::memcpy(mJgpij, o.Jgp, 72); std::memcpy(mJapij, o.Jap, 72); std::memcpy(mJgvij, o.Jgv, 72);
    std::memcpy(mJavij, o.Jav, 72); std::memcpy(mJgRij, o.JgR, 72);
    mdeltatij = o.dt;
    return o.status;
  }
};

// Optimizer (include/Optimizer.h:46-121): the flattened forms the reference's static methods call after collecting
// the graph (INTEGRATION.md lists the Frame / KeyFrame / MapPoint fields each array comes from).
struct Optimizer {
  // PoseOptimization(Frame*, Frame*) / PoseOptimization<KF>(Frame*, KF*, gw, bComputeMarg, bNoMPs): one frame
  static int PoseOptimization(VieoPoseOptProblem& pb, const VieoCamera& cam, const std::vector<double>& Xw,
                              const std::vector<float>& obs, const std::vector<float>& inv_sigma2,
                              const std::vector<uint8_t>& flags, VieoPoseOptResult& res, std::vector<uint8_t>& mvbOutlier,
                              int device = 0) {
    const int E = (int)flags.size();
    pb.edge_begin = 0;
    pb.edge_end = E;
    mvbOutlier.assign(E, 0);
    std::vector<double> chi2(E);
    vieo_check(vieo_pose_opt_batch(&pb, 1, &cam, Xw.data(), obs.data(), inv_sigma2.data(), flags.data(), E, &res,
                                   mvbOutlier.data(), chi2.data(), device), "vieo_pose_opt_batch");
    return res.n_inliers;
  }
  // The g2o part of OptimizeEssentialGraph(pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, bFixScale)
  // (src/Optimizer.cc:2309-2688) on a collected graph (vieo_flatten.hpp: CollectEssentialGraph): setUserLambdaInit(1e-16),
  // optimize(20).  Scw_out[k] / Tcw_out[12 k] per vertex; returns the number of LM iterations run.
  static int OptimizeEssentialGraph(const std::vector<VieoSim3>& vScw, const std::vector<uint8_t>& fixed, bool bFixScale,
                                    const std::vector<int32_t>& edge_i, const std::vector<int32_t>& edge_j,
                                    const std::vector<VieoSim3>& Sji, const std::vector<double>& info /* empty: identity */,
                                    std::vector<VieoSim3>& Scw_out, std::vector<double>& Tcw_out, VieoPoseGraphStats& stats,
                                    int device = 0) {
    const int K = (int)vScw.size(), E = (int)edge_i.size();
    Scw_out.assign(K, VieoSim3{});
    Tcw_out.assign(12 * (size_t)K, 0.0);
    vieo_check(vieo_essential_graph_optimize(K, vScw.data(), fixed.data(), bFixScale ? 1 : 0, E, edge_i.data(), edge_j.data(), Sji.data(),
                                             info.empty() ? nullptr : info.data(), 20, 1e-16, Scw_out.data(), Tcw_out.data(), &stats,
                                             device), "vieo_essential_graph_optimize");
    return stats.iterations;
  }
  // OptimizeSim3(pKF1, pKF2, vpMatches1, g2oS12, th2, bFixScale) (src/Optimizer.cc:2689-2920) on flattened matches: pb.ns /
  // pb.scale carry g2oS12 in and res.ns / res.scale carry it out; keep[i] == 0 means vpMatches1[idx_of_match[i]] = nullptr
  static int OptimizeSim3(VieoSim3Problem& pb, const VieoCamera& cam, const std::vector<double>& Xc1,
                          const std::vector<double>& Xc2, const std::vector<float>& obs1, const std::vector<float>& obs2,
                          const std::vector<float>& inv_sigma2_1, const std::vector<float>& inv_sigma2_2, VieoSim3Result& res,
                          std::vector<uint8_t>& keep, int device = 0) {
    const int M = (int)inv_sigma2_1.size();
    pb.m_begin = 0;
    pb.m_end = M;
    keep.assign(M, 0);
    vieo_check(vieo_optimize_sim3_batch(&pb, 1, &cam, Xc1.data(), Xc2.data(), obs1.data(), obs2.data(), inv_sigma2_1.data(),
                                        inv_sigma2_2.data(), M, &res, keep.data(), nullptr, nullptr, device),
               "vieo_optimize_sim3_batch");
    return res.n_inliers;
  }
};

// Optimizer::OptimizeInitialGyroBias(vpKFInit, bg, bInfo) (include/Optimizer.h:819-892) followed by the re-integration
// of every keyframe interval with the new bias (src/Odom/IMUInitialization.cpp:640-648).  pre[i] = vpKFInit[i]->
// GetIMUPreInt() (entry 0 ignored), Rwb[i] = Rwc_i * Rcb row-major; with sample lists given, reint receives
// ComputePreInt()'s results.  Returns num_equations.
inline int OptimizeInitialGyroBias(const std::vector<VieoImuPreint>& pre, const std::vector<double>& Rwb, double bg[3],
                                   bool bInfo = true, const std::vector<double>* samples = nullptr,
                                   const std::vector<int32_t>* seg_ptr = nullptr, const std::vector<double>* ti_tj = nullptr,
                                   const std::vector<double>* ba = nullptr, std::vector<VieoImuPreint>* reint = nullptr,
                                   int device = 0) {
  int neq = 0;
  const int n = (int)pre.size();
  if (reint) reint->resize(n);
  vieo_check(vieo_imu_init_gyro_bias(pre.data(), Rwb.data(), n, bInfo ? 1 : 0, bg, &neq, samples ? samples->data() : nullptr,
                                     seg_ptr ? seg_ptr->data() : nullptr, ti_tj ? ti_tj->data() : nullptr,
                                     ba ? ba->data() : nullptr, &IMUPreintegrator::Noise(), reint ? reint->data() : nullptr,
                                     device), "vieo_imu_init_gyro_bias");
  return neq;
}

// LocalBundleAdjustmentNavStatePRV engine: long-lived (LocalMapping thread), reused across windows
class LocalBA {
 public:
  explicit LocalBA(int max_states = 256, int max_points = 16384, int max_edges = 131072, int max_imu = 64, int device = 0) {
    vieo_check(vieo_ba_create(max_states, max_points, max_edges, max_imu, device, &h_), "vieo_ba_create");
  }
  ~LocalBA() { vieo_ba_destroy(h_); }
  LocalBA(const LocalBA&) = delete;
  LocalBA& operator=(const LocalBA&) = delete;
  // pbStopFlag is the reference's bool* (mbAbortBA); outputs sized by the caller
  int Run(const VieoBaProblem& pb, const VieoCamera& cam, const bool* pbStopFlag, VieoNavState* states_out,
          double* points_out, double* edge_chi2, uint8_t* erase, VieoBaResult& res) {
    static_assert(sizeof(bool) == 1, "bool* is polled as a byte");
    int rc = vieo_local_ba_prv(h_, &pb, &cam, reinterpret_cast<const volatile uint8_t*>(pbStopFlag), states_out, points_out,
                               edge_chi2, erase, &res);
    vieo_check(rc, "vieo_local_ba_prv");
    return rc;
  }
  // asynchronous form: Begin enqueues the whole routine on the engine's stream and returns (the LocalMapping thread can
  // go on collecting the next window), End waits and fills the outputs
  void Begin(const VieoBaProblem& pb, const VieoCamera& cam, const bool* pbStopFlag) {
    vieo_check(vieo_local_ba_prv_begin(h_, &pb, &cam, reinterpret_cast<const volatile uint8_t*>(pbStopFlag)), "vieo_local_ba_prv_begin");
  }
  bool Ready() { return vieo_local_ba_prv_poll(h_) != 0; }
  void End(VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase, VieoBaResult& res) {
    vieo_check(vieo_local_ba_prv_end(h_, states_out, points_out, edge_chi2, erase, &res), "vieo_local_ba_prv_end");
  }

 private:
  vieo_ba_t* h_ = nullptr;
};

// GlobalBundleAdjustmentNavStatePRV engine (src/Optimizer.cc:771-1342): a handle sized for the whole map
class GlobalBA {
 public:
  explicit GlobalBA(int max_keyframes = 768, int max_points = 1 << 17, int max_edges = 1 << 21, int device = 0) {
    vieo_check(vieo_ba_create_global(max_keyframes, max_points, max_edges, max_keyframes, device, &h_), "vieo_ba_create_global");
  }
  ~GlobalBA() { vieo_ba_destroy(h_); }
  GlobalBA(const GlobalBA&) = delete;
  GlobalBA& operator=(const GlobalBA&) = delete;
  vieo_ba_t* handle() { return h_; }
  // returns the LM iterations run; pbStopFlag is the reference's bool* (mbStopGBA)
  int Run(const VieoBaProblem& pb, const VieoCamera& cam, int nIterations, bool bRobust, const bool* pbStopFlag,
          VieoNavState* states_out, double* points_out, VieoBaResult& res) {
    int rc = vieo_global_ba_prv(h_, &pb, &cam, nIterations, bRobust ? 1 : 0, reinterpret_cast<const volatile uint8_t*>(pbStopFlag),
                                states_out, points_out, nullptr, &res);
    vieo_check(rc, "vieo_global_ba_prv");
    return rc;
  }
  // bScaleOpt (System::FinalGBA): *scale receives the VertexScale estimate, points_out are already multiplied by it;
  // gw_init (the IMU initialiser's call, pimu_initiator != nullptr): in = its gravity estimate, out = RwI * GI
  int RunEx(const VieoBaProblem& pb, const VieoCamera& cam, int nIterations, bool bRobust, bool bScaleOpt, double* scale,
            double* gw_init, const bool* pbStopFlag, VieoNavState* states_out, double* points_out, VieoBaResult& res) {
    VieoGbaExtra ex{};
    ex.scale_opt = bScaleOpt ? 1 : 0;
    ex.imu_init = gw_init ? 1 : 0;
    if (gw_init) std::memcpy(ex.gw, gw_init, 24);
    int rc = vieo_global_ba_prv_ex(h_, &pb, &cam, nIterations, bRobust ? 1 : 0, &ex,
                                   reinterpret_cast<const volatile uint8_t*>(pbStopFlag), states_out, points_out, nullptr, &res);
    vieo_check(rc, "vieo_global_ba_prv_ex");
    if (scale) *scale = ex.scale;
    if (gw_init) std::memcpy(gw_init, ex.gw, 24);
    return rc;
  }

 private:
  vieo_ba_t* h_ = nullptr;
};

// SM partition of one device between the front-end / tracking streams and the BA streams (CUDA green contexts); a
// thread binds itself before it creates its handles (include/vieo_b200.h)
class SmPartition {
 public:
  explicit SmPartition(int ba_sms, int device = 0) { vieo_check(vieo_sm_partition_create(device, ba_sms, &p_), "vieo_sm_partition_create"); }
  ~SmPartition() { vieo_sm_partition_destroy(p_); }
  SmPartition(const SmPartition&) = delete;
  SmPartition& operator=(const SmPartition&) = delete;
  void BindThisThread(int which) { vieo_check(vieo_sm_partition_bind_thread(p_, which), "vieo_sm_partition_bind_thread"); }
  int SMs(int which) const { return vieo_sm_partition_sms(p_, which); }

 private:
  vieo_sm_partition_t* p_ = nullptr;
};

}  // namespace VIEO_SLAM_B200

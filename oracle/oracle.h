// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  C interface of the CPU restatement.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcKeyPoint {  // cv::KeyPoint fields the reference fills (src/ORBextractor.cc:784-800)
  float x, y, size, angle, response;
  int32_t octave;
} OrcKeyPoint;

void orc_sincosf(float x, float* s, float* c);
float orc_fast_atan2(float y, float x);
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ostride);
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int th, int* xs, int* ys, int* resp, int cap);
void orc_gauss7_kernel(int k[7]);
void orc_gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);

void* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
void orc_orb_destroy(void* h);
void orc_orb_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* quota, int* umax);
int orc_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, const int* lapping, OrcKeyPoint* kps,
                    uint8_t* desc, int cap, int* n_mono);
int orc_orb_level_size(void* h, int level, int* w, int* hgt);
void orc_orb_get_level(void* h, int level, uint8_t* out);
int orc_orb_get_candidates(void* h, int level, int* xyr, int cap);
int orc_quadtree(const int* xyr, int n, int W, int H, int N, int* picked, int cap);

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
/* MapPoint::ComputeDistinctiveDescriptors over a CSR batch of map points (match_oracle.cc) */
void orc_distinctive_descriptors(const uint8_t* desc_pool, const int32_t* rows, const int32_t* ptr, int n_points,
                                 int32_t* best, int32_t* median);
void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);
void orc_hamming_csr(const uint8_t* q, const uint8_t* t, const int32_t* row_ptr, const int32_t* cand, int nrows,
                     int32_t* best_dist, int32_t* best_idx, int32_t* second_dist, int32_t* second_idx);
int orc_stereo_matches(const OrcKeyPoint* kl, const uint8_t* dl, int nl, const OrcKeyPoint* kr, const uint8_t* dr, int nr,
                       const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh,
                       const float* scale, const float* inv_scale, float bf, float minZ, float* uright, float* depth,
                       int32_t* sad);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* IMU pre-integration (imu_oracle.cc).  Matrices row-major. */
typedef struct OrcImuNoise {
  double sigma_g, sigma_a, sigma_bg, sigma_ba; /* diagonal entries of mSigmag/mSigmaa/mSigmabg/mSigmaba after SetParam */
  double freq_ref;
  int32_t dt_cov_noise_fixed;
  int32_t pad_;
} OrcImuNoise;
typedef struct OrcImuPreint { /* public members of IMUPreIntegratorBase (OdomPreIntegrator.h:108-150) */
  double Rij[9], vij[3], pij[3];
  double SigmaPRV[81], SigmaPVR[81];
  double Jgp[9], Jap[9], Jgv[9], Jav[9], JgR[9];
  double dt;
  int32_t status;
  int32_t pad_;
} OrcImuPreint;
void orc_imu_set_param(OrcImuNoise* nz, const double sigma2[4], int dt_cov_noise_fixed, double freq_ref);
int orc_imu_preintegrate(const double* smp, int n, double ti, double tj, const double bg[3], const double ba[3],
                         const OrcImuNoise* nz, OrcImuPreint* out);
/* the update() calls of the same routine alone: trace [cap][7] = (omega, acc, dt) per call (test hook) */
int orc_imu_preintegrate_trace(const double* smp, int n, double ti, double tj, const double bg[3], const double ba[3], double* trace,
                               int cap, int* n_updates);
/* Optimizer::OptimizeInitialGyroBias: pre [n_kf] (entry 0 ignored), Rwb [n_kf][9] -> dbg, returns num_equations */
int orc_gyro_bias_init(const OrcImuPreint* pre, const double* Rwb, int n_kf, int use_info, double dbg[3]);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* Guided searches of ORBmatcher (sbp_oracle.cc).  One current frame per call. */
typedef struct OrcSbpFrame {
  int32_t kp_begin, n_kp;       /* this frame's keypoints in the keypoint arrays */
  int32_t q_begin, n_q;         /* this frame's queries (map points in the reference's loop order) */
  float minx, maxx, miny, maxy; /* FrameBase::gridinfo_.minmax_xy_ */
  float grid_winv, grid_hinv;   /* fgrids_widthinv_ / fgrids_heightinv_ */
  float bf, b;                  /* stereoinfo_.baseline_bf_[1] / [0] */
  float fx, fy, cx, cy;         /* mpCameras[0]->toK() cast to float */
  float th, th_far;             /* window factor; th_far_pts (<= 0: off) */
  float nn_ratio;               /* mfNNratio */
  int32_t mono, check_orientation, n_levels;
  float scale[16];              /* scalepyrinfo_.vscalefactor_ */
  double qcw[4], tcw[3];        /* CurrentFrame.GetTcwCst(): unit quaternion (w, x, y, z) and translation */
  double qlw[4], tlw[3];        /* LastFrame.GetTcwCst() */
} OrcSbpFrame;
int orc_sbp_last_frame(const OrcSbpFrame* f, const OrcKeyPoint* kps, const float* uright, const uint8_t* desc,
                       const double* q_Xw, const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                       const uint8_t* q_flags, const uint8_t* kp_blocked, int32_t* kp_match, int32_t* q_match,
                       int32_t* q_dist);
int orc_sbp_local_map(const OrcSbpFrame* f, const OrcKeyPoint* kps, const float* uright, const uint8_t* desc,
                      const float* q_proj, const int32_t* q_level, const float* q_viewcos, const float* q_depth,
                      const uint8_t* q_desc, const uint8_t* q_flags, const uint8_t* kp_blocked, int32_t* kp_match,
                      int32_t* q_match, int32_t* q_dist);
/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist, th_far_pts)
 * (src/ORBmatcher.cc:1471-1606), the relocalisation search.  Queries = pKF's map points that are good and not in
 * sAlreadyFound, in keypoint order; q_angle = pKF->mvKeys[i].angle; q_max_dist / q_min_dist = mfMaxDistance / mfMinDistance
 * (the 1.2 / 0.8 factors are applied inside); kp_blocked: keypoints of the current frame that hold a map point. */
int orc_sbp_reloc(const OrcSbpFrame* f, int orb_dist, float log_scale_factor, const OrcKeyPoint* kps, const uint8_t* desc,
                  const double* q_Xw, const float* q_angle, const float* q_max_dist, const float* q_min_dist,
                  const uint8_t* q_desc, const uint8_t* kp_blocked, int32_t* kp_match, int32_t* q_match, int32_t* q_dist,
                  int32_t* q_level);
/* Frame::isInFrustum + MapPoint::PredictScale (sbp_oracle.cc), single camera, usedistort_ == false. */
typedef struct OrcFrustumFrame {
  int32_t q_begin, n_q;         /* this frame's candidate map points in the point arrays */
  float Rcw[9], tcw[3], Ow[3];  /* Tcw_ rotation (row-major), mtcw, mOw cast to float */
  float fx, fy, cx, cy;         /* mpCameras[0]->toK() cast to float */
  float minx, maxx, miny, maxy; /* gridinfo_.minmax_xy_ */
  float bf;                     /* stereoinfo_.baseline_bf_[1] */
  float cos_limit;              /* viewingCosLimit (0.5 in SearchLocalPoints) */
  float log_scale_factor;       /* scalepyrinfo_.flogscalefactor_ */
  int32_t n_levels;             /* scalepyrinfo_.vscalefactor_.size() */
} OrcFrustumFrame;
/* ORBmatcher::SearchByProjectionBase search half (sbp_oracle.cc).  One keyframe per call. */
typedef struct OrcProjSearchFrame {
  int32_t kp_begin, n_kp, q_begin, n_q;
  float Rcw[9], tcw[3], Ow[3];  /* Rcrw, tcrw, pKF->GetCameraCenter() cast to float */
  float fx, fy, cx, cy;
  float minx, maxx, miny, maxy;
  float grid_winv, grid_hinv;
  float bf;                     /* *pbf */
  int32_t use_bf;               /* pbf != nullptr: chi-square gate on */
  int32_t check_viewing_angle;  /* bCheckViewingAngle */
  float th_radius;
  int32_t n_levels;
  float log_scale_factor;
  float scale[16];              /* vscalefactor_ */
  float inv_level_sigma2[16];   /* vinvlevelsigma2_ */
  float level_ratio[16];        /* device only (vieo_frustum_level_table); ignored by the oracle */
} OrcProjSearchFrame;
void orc_sbp_base(const OrcProjSearchFrame* f, const OrcKeyPoint* kps, const float* uright, const uint8_t* desc,
                  const float* wP, const float* Pn, const float* max_dist, const float* min_dist, const uint8_t* q_desc,
                  const uint8_t* q_skip, int32_t* best_idx, int32_t* best_dist, int32_t* level);
/* Frame::isInFrustum with a camera rig (mpCameras.size() > 1, src/Frame.cc:335-416): per camera Pc = GetTcr() * Pcr
 * (Sophus::SE3f: Eigen's float _transformVector + translation), twc = mOw + Rcrw^T GetTrc().translation(), projection by
 * K (usedistort_ == false, model 0) or by the camera's own Project in double (usedistort_: 1 pinhole, 2 KB8,
 * common/camera_models/camera_{pinhole,kb8}.h), per-camera image bounds. */
typedef struct OrcFrustumCam {
  float q_cr[4];                /* GetTcr().unit_quaternion() (x, y, z, w) */
  float t_cr[3];                /* GetTcr().translation() */
  float t_rc[3];                /* GetTrc().translation() */
  float fx, fy, cx, cy;         /* parameters_[0..3] */
  float k[4];                   /* KB8 k1..k4 */
  float minx, maxx, miny, maxy; /* gridinfo_.minmax_xy_[cami] */
  int32_t model;                /* 0: K * p_normalize in float; 1: PinholeCamera::Project; 2: KB8Camera::Project */
  int32_t pad_;
} OrcFrustumCam;
typedef struct OrcFrustumRigFrame {
  int32_t q_begin, n_q;
  float Rcw[9], tcw[3], Ow[3];
  float bf, cos_limit, log_scale_factor;
  int32_t n_levels, n_cams;     /* n_cams <= 4 */
  float level_ratio[16];        /* device only */
  OrcFrustumCam cam[4];
} OrcFrustumRigFrame;
/* Outputs per point: inview = btrack_inview_, cam_mask bit c = camera c pushed an entry (vtrack_cami_), per camera slot
 * proj [n][4][3] = (u, v, ur), level [n][4] (-1), viewcos [n][4]; depth [n] = track_depth_ (mean over the cameras in view). */
int orc_is_in_frustum_rig(const OrcFrustumRigFrame* f, int n, const float* wP, const float* Pn, const float* max_dist,
                          const float* min_dist, uint8_t* inview, uint8_t* cam_mask, float* proj, int32_t* level, float* viewcos,
                          float* depth);
/* the frame grid alone (test hook): candidate lists of n_q windows (x, y, r | minlevel, maxlevel) in the reference's walk order */
int orc_features_in_area(const OrcKeyPoint* kps, int n_kp, float minx, float maxx, float miny, float maxy, float winv, float hinv,
                         const float* q_xyr, const int32_t* q_levels, int n_q, int32_t* out_ptr, int32_t* out_idx, int cap,
                         uint8_t* in_image);
int orc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);
int orc_is_in_frustum(const OrcFrustumFrame* f, int n, const float* wP, const float* Pn, const float* max_dist,
                      const float* min_dist, uint8_t* inview, float* proj, int32_t* level, float* viewcos, float* depth);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* Frame::ComputeStereoFishEyeMatches brute-force half for one multi-camera frame (match_oracle.cc) */
void orc_fisheye_matches(const uint8_t* desc, const int32_t* n_kp, const int32_t* n_mono, int n_cams, int cap, int32_t* idx,
                         int32_t* dist, uint8_t* good);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* ORBmatcher::SearchForTriangulation, single pinhole camera (sft_oracle.cc) */
int orc_search_for_triangulation(const OrcKeyPoint* kp1, const float* ur1, const uint8_t* desc1, const uint8_t* has_mp1,
                                 const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int n_nodes1,
                                 const OrcKeyPoint* kp2, const float* ur2, const uint8_t* desc2, const uint8_t* has_mp2,
                                 const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int n_nodes2,
                                 const double F12[9], float ex, float ey, const float* scale_factor2,
                                 const float* level_sigma2_2, int only_stereo, int check_orientation, int32_t* pairs, int cap);
/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), single camera (sft_oracle.cc) */
int orc_search_by_bow(const OrcKeyPoint* kp_kf, const uint8_t* desc_kf, const int32_t* mp_id, const int32_t* fv1_node,
                      const int32_t* fv1_ptr, const int32_t* fv1_idx, int n_nodes1, const OrcKeyPoint* kp_f,
                      const uint8_t* desc_f, int n_f, const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx,
                      int n_nodes2, float nn_ratio, int check_orientation, int32_t* match_f);
#ifdef __cplusplus
}
#endif

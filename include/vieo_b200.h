/* vieo_b200.h — C ABI of the B200-native VIEO_SLAM hot path (libvieo_b200.so).
 *
 * Plain C: opaque handles, plain pointers and sizes, int status (0 = ok, <0 = VIEO_E_*), no
 * exceptions, no torch / OpenCV / Eigen types.  Buffers are HOST pointers unless the function name
 * ends in _dev, in which case every pointer is a device pointer on the handle's GPU and the call is
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream).
 * A handle owns its device scratch and one CUDA stream and must not be used from two threads at
 * once (the reference drives one ORBextractor per camera thread, src/Frame.cc:259-278).
 * There is NO CPU fallback: every entry point fails with VIEO_E_CUDA if no sm_100 device is usable.
 *
 * Each entry point cites the reference interface (leavesnight/VIEO_SLAM @356e4a22) it replaces.
 */
#ifndef VIEO_B200_H
#define VIEO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIEO_OK 0
#define VIEO_E_ARG (-1)      /* bad argument / unsupported configuration */
#define VIEO_E_CUDA (-2)     /* CUDA runtime error or no device; see vieo_last_error() */
#define VIEO_E_CAPACITY (-3) /* caller buffer or handle capacity too small */
#define VIEO_E_EMPTY (-4)    /* empty image: ORBextractor::operator() returns -1 (src/ORBextractor.cc:970) */

const char* vieo_last_error(void); /* thread-local text of the last failure */
int vieo_device_count(void);
const char* vieo_version(void);
/* How host threads wait for the device in this process (cudaSetDeviceFlags on `device`): 0 the driver's heuristic (spins
 * while there are more logical CPUs than contexts), 1 spin, 2 yield between polls, 3 block on an interrupt.  The reference
 * runs Tracking / LocalMapping / LoopClosing threads beside the per-camera extractor threads: several processes of that
 * shape on one host (one per GPU) oversubscribe the cores with spinning waiters, mode 2 keeps the launching threads running. */
int vieo_set_host_sync(int device, int mode);

/* ------------------------------------------------------------------------------------------------
 * ORB extractor — replaces ORBextractor (include/ORBextractor.h:27-80, src/ORBextractor.cc:391-1081).
 * Fields of cv::KeyPoint the reference fills (src/ORBextractor.cc:784-800); class_id is unused. */
typedef struct VieoKeyPoint {
  float x, y;     /* level-0 pixel coordinates (pt *= mvScaleFactor[octave], :1036) */
  float size;     /* (int)(31 * scale[octave]) */
  float angle;    /* degrees, cv::fastAtan2 of the intensity centroid (:55-80) */
  float response; /* FAST-9/16 score (cornerScore<16>) */
  int32_t octave;
} VieoKeyPoint;

typedef struct VieoOrbConfig {
  int32_t width, height;  /* fixed per handle (the reference re-derives sizes per call; cameras are fixed) */
  int32_t nfeatures;      /* ORBextractor.nFeatures */
  float scale_factor;     /* ORBextractor.scaleFactor */
  int32_t nlevels;        /* ORBextractor.nLevels (<= 16) */
  int32_t ini_th_fast;    /* ORBextractor.iniThFAST */
  int32_t min_th_fast;    /* ORBextractor.minThFAST */
  int32_t max_batch;      /* images in flight per call (cameras x frames); device scratch is sized for it */
} VieoOrbConfig;

typedef struct vieo_orb vieo_orb_t;

/* ORBextractor::ORBextractor (src/ORBextractor.cc:391-456). */
int vieo_orb_create(const VieoOrbConfig* cfg, int device, vieo_orb_t** out);
void vieo_orb_destroy(vieo_orb_t* h);
/* per-image keypoint capacity the handle can produce (sum over levels of quota+3) */
int vieo_orb_max_keypoints(const vieo_orb_t* h);
/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:47-57) plus the per-level feature quota and level sizes. Arrays of nlevels. */
int vieo_orb_get_tables(const vieo_orb_t* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                        int32_t* quota, int32_t* level_w, int32_t* level_h);

/* ORBextractor::operator() (src/ORBextractor.cc:968-1058) for one image.
 *   img/stride  CV_8UC1 rows;  lapping  NULL or {x0,x1} (KB8 "lapping area", :1041-1057)
 *   kps/desc    caller arrays of `cap` entries (desc = cap x 32 bytes)
 *   n_mono      the reference's return value (monoIndex), may be NULL
 *   pyr_host    NULL, or nlevels host pointers receiving tightly packed copies of the pyramid levels
 *               (public mvImagePyramid contract, read by Frame::ComputeStereoMatches src/Frame.cc:536-557)
 * Returns the number of keypoints (>= 0), VIEO_E_EMPTY for a NULL/0-sized image, or an error. */
int vieo_orb_extract(vieo_orb_t* h, const uint8_t* img, int stride, const int32_t* lapping, VieoKeyPoint* kps,
                     uint8_t* desc, int cap, int32_t* n_mono, uint8_t* const* pyr_host);

/* Batched form: the per-camera std::threads of Frame::Frame (src/Frame.cc:259-278) collapsed into one
 * call; also used to keep several frames in flight.  imgs = n_img images, `img_stride` bytes apart, rows
 * `row_stride` bytes apart.  kps = [n_img][cap], desc = [n_img][cap][32], n_kp = [n_img].  Level order
 * output (no lapping area). */
int vieo_orb_extract_batch(vieo_orb_t* h, int n_img, const uint8_t* imgs, size_t img_stride, int row_stride,
                           VieoKeyPoint* kps, uint8_t* desc, int cap, int32_t* n_kp);
/* Same with device-resident input and output; asynchronous on `stream`. */
int vieo_orb_extract_batch_dev(vieo_orb_t* h, int n_img, const uint8_t* imgs_dev, size_t img_stride, int row_stride,
                               VieoKeyPoint* kps_dev, uint8_t* desc_dev, int cap, int32_t* n_kp_dev, void* stream);
/* Stage introspection for parity tests (device scratch of the last call -> host).
 * level pixels (tightly packed w*h), FAST candidates before the quadtree as (x, y, response) int32
 * triples in the reference's visiting order (cell-major, raster inside a cell). */
int vieo_orb_debug_level(vieo_orb_t* h, int img_index, int level, uint8_t* out);
int vieo_orb_debug_candidates(vieo_orb_t* h, int img_index, int level, int32_t* xyr, int cap);
/* Rectified-stereo association — Frame::ComputeStereoMatches (src/Frame.cc:451-611) on the device-resident results of
 * the LAST extract call of `h` (the pyramid levels it needs are still in the handle): images 2f / 2f+1 of that batch are
 * the left / right view of frame f.  kps/desc/n_kp: that call's device outputs with `cap` slots per image.
 *   bf = stereoinfo_.baseline_bf_[1], min_z = baseline_bf_[0] (= bf / fx)
 * Outputs [n_frames][cap] indexed by left keypoint: uright (vuright_), depth (vdepth_) (-1: no match) and the block-search
 * SAD of every match that passed the disparity checks (-1: none; matches dropped by the 2.1 x median filter keep it). */
int vieo_orb_stereo_match_dev(vieo_orb_t* h, int n_frames, const VieoKeyPoint* kps_dev, const uint8_t* desc_dev,
                              const int32_t* n_kp_dev, int cap, float bf, float min_z, float* uright_dev, float* depth_dev,
                              int32_t* sad_dev, void* stream);
/* number of kernel launches issued by the last extract call (bench.py's gpu_launches) */
int vieo_orb_last_launches(const vieo_orb_t* h);
/* Per-stage device timing with CUDA events on the launching stream (the reference's mlog::Timer stamps,
 * common/mlog/log.h:109-155, moved to the device).  vieo_orb_profile(h,1) arms it; every following extract
 * call records 5 events; vieo_orb_profile_read sums {pyramid, fast_cells, quadtree, orient_desc} ms over the
 * calls since the last read (at most 1024) and rearms. */
int vieo_orb_profile(vieo_orb_t* h, int enable);
int vieo_orb_profile_read(vieo_orb_t* h, float stage_ms[4], int32_t* n_calls);

/* ------------------------------------------------------------------------------------------------
 * 256-bit Hamming matching — replaces ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1645-1667),
 * cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) in Frame::ComputeStereoFishEyeMatches (src/Frame.cc:620-628)
 * and the best/second-best candidate loops of the guided searches (src/ORBmatcher.cc:286-315,
 * 1396-1424; src/Frame.cc:506-523).  Descriptors are rows of 32 bytes. */

/* knnMatch k=2: idx/dist are [nq][2], ascending distance, ties -> lowest train index; missing
 * neighbours are idx -1 / dist INT32_MAX. */
int vieo_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist, int device);
/* n_pairs independent (query set, train set) pairs, device resident.  Pair p: q + p*q_stride (bytes),
 * nq_dev[p*count_stride] rows (<= max_nq); t likewise; idx/dist = [n_pairs][max_nq][2]. */
int vieo_hamming_knn2_batch_dev(const uint8_t* q_dev, size_t q_stride, const int32_t* nq_dev, int max_nq,
                                const uint8_t* t_dev, size_t t_stride, const int32_t* nt_dev, int max_nt,
                                int count_stride, int n_pairs, int32_t* idx_dev, int32_t* dist_dev, void* stream);
/* ---- multi-camera frames (KB8 / distorted-stereo rigs; BASELINE configs[3]) ---------------------------------
 * ORBextractor::operator()'s lapping-area ordering (src/ORBextractor.cc:1041-1057) on device-resident level-ordered
 * outputs of vieo_orb_extract_batch_dev: per image, keypoints with lapping[0] <= pt.x <= lapping[1] are written from the
 * BACK of the arrays in reverse visiting order, the others from the front; n_mono = monoIndex (the reference's return
 * value).  lapping_dev = [n_img][2] int32 on the device, or NULL = pvLappingArea == nullptr: order kept and n_mono = 0
 * (monoIndex is never advanced, :1003,1057 — every keypoint then takes part in the pair matching).  Out of place. */
int vieo_lapping_split_dev(const VieoKeyPoint* kps_in_dev, const uint8_t* desc_in_dev, const int32_t* n_kp_dev, int n_img,
                           int cap, const int32_t* lapping_dev, VieoKeyPoint* kps_out_dev, uint8_t* desc_out_dev,
                           int32_t* n_mono_dev, void* stream);
/* Frame::ComputeStereoFishEyeMatches, brute-force half (src/Frame.cc:613-663) for n_frames frames of n_cams cameras
 * (image of (frame f, camera c) = f * frame_stride + c * cam_stride in the `cap`-slot arrays — [frame][camera] order is
 * (n_cams, 1), the per-camera shards gathered from the GPUs of a rig [camera][frame] are (1, frames per shard) —
 * lapping-ordered as above): for every camera pair i < j in the
 * reference's order, BFMatcher(NORM_HAMMING).knnMatch(desc_i[num_mono_i:], desc_j[num_mono_j:], k = 2) — skipped (no
 * matches) when either side has no in-area keypoint (:623) — and Lowe's ratio test `size >= 2 && (d0 < 0.7 d1 || (d0 < 75 &&
 * d0 < 0.9 d1))` (:659-663).  Outputs [n_frames][n_pairs][cap]: idx / dist [..][2] (queryIdx row r <-> keypoint r +
 * num_mono_i; trainIdx + num_mono_j is the keypoint of camera j; -1 / INT32_MAX = none), good = passed the ratio test
 * (the matches FillMatchesFromPair is called on). */
int vieo_fisheye_knn_dev(const uint8_t* desc_dev, const int32_t* n_kp_dev, const int32_t* n_mono_dev, int n_cams,
                         int n_frames, int cap, int frame_stride, int cam_stride, int32_t* idx_dev, int32_t* dist_dev,
                         uint8_t* good_dev, void* stream);
/* Host-buffer form of both for a batch of frames: per-camera extraction (Frame::Frame, src/Frame.cc:259-278) with the
 * camera's lapping area (lapping = [n_cams][2], NULL = none) + the pair matching above.  imgs: [n_frames][n_cams] images
 * `img_stride` bytes apart; cap = vieo_orb_max_keypoints(h); n_frames * n_cams <= max_batch.
 * kps / desc [n_img][cap], n_kp / n_mono [n_img], pair_* as above (may be NULL when n_cams == 1). */
int vieo_multicam_frames(vieo_orb_t* h, int n_frames, int n_cams, const uint8_t* imgs, size_t img_stride, int row_stride,
                         const int32_t* lapping, VieoKeyPoint* kps, uint8_t* desc, int32_t* n_kp, int32_t* n_mono,
                         int32_t* pair_idx, int32_t* pair_dist, uint8_t* pair_good);
/* Candidate-list search: row r compares q row r with t rows cand[row_ptr[r] .. row_ptr[r+1]) in list
 * order; strict '<' keeps the first of equal distances (the reference's loops).  Outputs per row: best
 * and second-best distance (256 when absent) and their train indices (-1 when absent). */
int vieo_hamming_csr(const uint8_t* q, const uint8_t* t, int nt, const int32_t* row_ptr, const int32_t* cand,
                     int nrows, int32_t* best_dist, int32_t* best_idx, int32_t* second_dist, int32_t* second_idx,
                     int device);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:314-378) for a batch of map points: map
 * point p owns the observed descriptors desc_pool[rows[k]], k in [ptr[p], ptr[p+1]) (rows nullable: desc_pool[k]) in the
 * reference's vDescriptors order; all-pairs Hamming distances, per row the sorted entry at index int(0.5 * (N - 1)), first
 * row with the least median wins.  best[p] = index into the point's list (-1: no observation, mDescriptor untouched),
 * median[p] = that median.  desc_pool 16-byte aligned in the _dev form. */
int vieo_distinctive_descriptors(const uint8_t* desc_pool, int n_pool, const int32_t* rows, const int32_t* ptr,
                                 int n_points, int32_t* best, int32_t* median, int device);
int vieo_distinctive_descriptors_dev(const uint8_t* desc_pool_dev, const int32_t* rows_dev, const int32_t* ptr_dev,
                                     int n_points, int32_t* best_dev, int32_t* median_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-manifold IMU pre-integration — replaces IMUPreIntegratorBase<IMUDataBase>::PreIntegration / update
 * (src/Odom/OdomPreIntegrator.h:227-506) for a batch of intervals (one per frame pair / keyframe pair; the
 * batched re-integration of IMUInitialization.cpp:648,1148,1422 is the same call).  Matrices are row-major. */
typedef struct VieoImuNoise { /* IMUDataBase statics after SetParam (src/Odom/OdomData.h:41-56) */
  double sigma_g, sigma_a, sigma_bg, sigma_ba; /* diagonal of mSigmag / mSigmaa / mSigmabg / mSigmaba */
  double freq_ref;                             /* mFreqRef */
  int32_t dt_cov_noise_fixed;                  /* mdt_cov_noise_fixed */
  int32_t pad_;
} VieoImuNoise;
typedef struct VieoImuPreint { /* public members of IMUPreIntegratorBase (OdomPreIntegrator.h:108-150) */
  double Rij[9], vij[3], pij[3];     /* mRij, mvij, mpij */
  double SigmaPRV[81], SigmaPVR[81]; /* mSigmaijPRV (P,R,V order), mSigmaij (P,V,R order) */
  double Jgp[9], Jap[9], Jgv[9], Jav[9], JgR[9]; /* mJgpij, mJapij, mJgvij, mJavij, mJgRij */
  double dt;                         /* mdeltatij (0 = not pre-integrated) */
  int32_t status;                    /* PreIntegration's return value: 0, or -1 for a sample gap > 1.5 s */
  int32_t pad_;
} VieoImuPreint;
/* IMUDataBase::SetParam: sigma2 = squared {gyro, acc, gyro-bias, acc-bias} noise (src/Tracking.cc:744-748). */
void vieo_imu_set_param(VieoImuNoise* nz, const double sigma2[4], int dt_cov_noise_fixed, double freq_ref);
/* samples: rows {t, ax, ay, az, wx, wy, wz}; interval k integrates samples[seg_ptr[k] .. seg_ptr[k+1]) (the
 * iterBegin/iterEnd list of the reference) over [ti_tj[2k], ti_tj[2k+1]] (reversed time allowed) with the
 * linearisation biases bg_ba[6k..6k+6) = {bg, ba}.  An empty list leaves the identity state (dt = 0). */
int vieo_imu_preint_batch(const double* samples, const int32_t* seg_ptr, const double* ti_tj, const double* bg_ba,
                          const VieoImuNoise* noise, int n_intervals, VieoImuPreint* out, int device);
int vieo_imu_preint_batch_dev(const double* samples_dev, const int32_t* seg_ptr_dev, const double* ti_tj_dev,
                              const double* bg_ba_dev, const VieoImuNoise* noise, int n_intervals,
                              VieoImuPreint* out_dev, void* stream);

/* Step 1 of the IMU initialiser — replaces Optimizer::OptimizeInitialGyroBias
 * (include/Optimizer.h:819-892: EdgeGyrBias per consecutive keyframe pair, src/Odom/g2otypes.h:940-973, one Gauss-Newton
 * iteration from a zero seed) and the re-integration of every keyframe interval with the new bias that follows it
 * (src/Odom/IMUInitialization.cpp:606-648, IMUKeyFrameInit::ComputePreInt IMUInitialization.h:225-232): gyro-bias
 * kernel -> pre-integration kernel on one stream, the bias never leaves the device in between.
 *   pre  [n_kf]     keyframe i's pre-integration from keyframe i-1 (entry 0 and entries with dt == 0 are ignored)
 *   Rwb  [n_kf][9]  Rwc_i * Rcb, row-major
 *   use_info        bInfo: information = (SigmaPRV.block<3,3>(3,3))^-1, else identity
 *   bg   [3]        in: the caller's bias; out: bias + estimate (untouched when there is no equation)
 *   preint_out [n_kf] nullable: when given, interval i (samples[seg_ptr[i] .. seg_ptr[i+1]) over ti_tj[2i..2i+1]) is
 *                   integrated again with {bg out, ba[3i..3i+3)} (ba nullable = zero); interval 0 is normally empty
 * *num_equations = the reference's return value. */
int vieo_imu_init_gyro_bias(const VieoImuPreint* pre, const double* Rwb, int n_kf, int use_info, double bg[3],
                            int* num_equations, const double* samples, const int32_t* seg_ptr, const double* ti_tj,
                            const double* ba, const VieoImuNoise* noise, VieoImuPreint* preint_out, int device);
/* Device-pointer form of the gyro-bias step alone.  result_dev: 56 bytes {double dbg[3]; double bg[3]; int32
 * num_equations; int32 solved}; bg_ba_out_dev (nullable) [n_kf][6] receives {bg + dbg, ba of bg_ba_in_dev (nullable: 0)}
 * ready for vieo_imu_preint_batch_dev on the same stream. */
int vieo_gyro_bias_init_dev(const VieoImuPreint* pre_dev, const double* Rwb_dev, int n_kf, int use_info, const double bg[3],
                            const double* bg_ba_in_dev, double* bg_ba_out_dev, void* result_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bundle adjustment — replaces the g2o graphs behind Optimizer::PoseOptimization (src/Optimizer.cc:1611-1874
 * visual / PR vertex; include/Optimizer.h:208-816 IMU / PVR vertex incl. the marginal prior) and
 * Optimizer::LocalBundleAdjustmentNavStatePRV / LocalBundleAdjustment (src/Optimizer.cc:21-769, 1876-2307):
 * per-edge residual + Jacobian evaluation (src/Odom/g2otypes.h:321-541, 725-884; g2otypes.cpp:14-124), Huber
 * weighting, the Levenberg-Marquardt loop and Schur complement of the vendored g2o
 * (optimizer/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189, block_solver.hpp:353-604). */
typedef struct VieoNavState { /* NavState (src/Odom/NavState.h:17-36) */
  double p[3];           /* mpwb */
  double q[4];           /* mRwb unit quaternion (w, x, y, z) */
  double v[3];           /* mvwb */
  double bg[3], ba[3];   /* mbg, mba: frozen linearisation point */
  double dbg[3], dba[3]; /* mdbg, mdba: optimised deltas */
} VieoNavState;
#define VIEO_CAM_PINHOLE 0 /* common/camera_models/camera_pinhole.h:70-106 */
#define VIEO_CAM_RADTAN 1  /* camera_radtan.h:61-129: dist = {k1..k_num_k, p1, p2} */
#define VIEO_CAM_KB8 2     /* camera_kb8.h:68-157: dist = {k1, k2, k3, k4} */
typedef struct VieoCamera { /* camm::Camera parameters (all float in the reference) + Frame::meigRcb / meigtcb */
  float fx, fy, cx, cy, bf;
  int32_t model;  /* VIEO_CAM_* */
  int32_t num_k;  /* radtan: number of radial coefficients (2 or 3) */
  float pad_;
  float dist[8];
  double Rcb[9], tcb[3];
} VieoCamera;
/* visual edge flags */
#define VIEO_EDGE_STEREO 1   /* EdgeReproject*Stereo (ul, vl, ur), else the 2-dim mono edge */
#define VIEO_EDGE_CLOSE 2    /* track_depth_ < max(10, mThDepth): 1.5x chi2 gate (include/Optimizer.h:554,568) */
#define VIEO_EDGE_LEVEL1 4   /* edge starts at level 1 (far-point guard, src/Optimizer.cc:514-518) */
#define VIEO_EDGE_NOKERNEL 8 /* no Huber kernel */

typedef struct VieoPoseOptProblem {
  VieoNavState cur, last, prior; /* pFrame->mNavState, pLastKF->GetNavState(), pLastKF->mNavStatePrior */
  VieoImuPreint preint;          /* pFrame->GetIMUPreInt(); dt == 0: no IMU edge */
  double prior_info[225];        /* pLastKF->mMargCovInv, row-major, order P V R bg ba */
  double gw[3];                  /* gravity in the world frame */
  double inv_sigma_bg2, inv_sigma_ba2; /* IMUDataBase::mInvSigmabg2 / mInvSigmaba2 */
  double dt_frames;              /* pFrame->ftimestamp_ - pLastKF->ftimestamp_ */
  int32_t mode;                  /* 0: PoseOptimization(Frame*, Frame*) visual; 1: the IMU template (PVR vertex) */
  int32_t last_has_prior;        /* pLastKF->mbPrior: last frame's vertices are free, 15-dim prior edge added */
  int32_t compute_marg;          /* bComputeMarg */
  int32_t no_mps;                /* bNoMPs */
  int32_t edge_begin, edge_end;  /* this frame's range in the shared edge arrays */
} VieoPoseOptProblem;
typedef struct VieoPoseOptResult {
  VieoNavState cur;          /* optimised pFrame->mNavState */
  VieoNavState last;         /* optimised last-frame state when it was free */
  double marg_cov_inv[225];  /* pFrame->mMargCovInv (compute_marg) */
  double chi2_final;         /* activeRobustChi2 after the last optimize() */
  double lambda_final;
  int32_t n_inliers;         /* the reference's return value */
  int32_t n_initial;         /* nInitialCorrespondences */
  int32_t iterations;        /* LM iterations run in total */
  int32_t prior_set;         /* pFrame->mbPrior after the call */
} VieoPoseOptResult;
/* A batch of independent frames, one thread block each, the whole 4 x optimize(10) schedule on the device.
 *   Xw [E][3] f64 map points, obs [E][3] f32 (ul, vl, ur), inv_sigma2 [E] f32, flags [E] u8 (STEREO | CLOSE)
 *   outputs: res [n], outlier [E] u8 (pFrame->mvbOutlier), chi2 [E] f64 (last e->chi2()) */
int vieo_pose_opt_batch(const VieoPoseOptProblem* pbs, int n, const VieoCamera* cam, const double* Xw, const float* obs,
                        const float* inv_sigma2, const uint8_t* flags, int n_edges, VieoPoseOptResult* res,
                        uint8_t* outlier, double* chi2, int device);
int vieo_pose_opt_batch_dev(const VieoPoseOptProblem* pbs_dev, int n, const VieoCamera* cam_dev, const double* Xw_dev,
                            const float* obs_dev, const float* inv_sigma2_dev, const uint8_t* flags_dev,
                            VieoPoseOptResult* res_dev, uint8_t* outlier_dev, double* chi2_dev, void* stream);

/* ---- Optimizer::OptimizeSim3 (src/Optimizer.cc:2689-2920): the loop closer's Sim3 refinement of candidate keyframe pairs ----
 * One free VertexNavStatePR (S12 as a NavState: mRwb = R12^-1, mpwb = -(mRwb t12), :2717-2722) + VertexScale (fixed iff
 * bFixScale) + fixed points; per match i an EdgeReprojectPRS (kpUn1 <- S12 P3D2c) and an EdgeReprojectPRSInv
 * (kpUn2 <- S12^-1 P3D1c) (src/Odom/g2otypes.h:321-549), Huber sqrt(th2), information invSigma2(octave) I; optimize(5), pairs
 * with a chi2 > th2 removed, optimize(10 if any was removed else 5).  Replaces the g2o graph of the routine; the caller
 * (LoopClosing::ComputeSim3, src/LoopClosing.cc) keeps the candidate policy and converts S12 <-> NavState. */
typedef struct VieoSim3Problem {
  VieoNavState ns;        /* only p, q are read */
  double scale;           /* g2oS12.scale() */
  float th2;              /* chi2 gate (10 in LoopClosing); Huber delta = sqrtf(th2) */
  int32_t fix_scale;      /* bFixScale (stereo / RGB-D / VIO with known scale) */
  int32_t m_begin, m_end; /* this candidate's range in the shared match arrays */
} VieoSim3Problem;
typedef struct VieoSim3Result {
  VieoNavState ns;    /* optimised vertex (the input when n_inliers == 0: the reference returns before the write-back) */
  double scale;
  double chi2_final;  /* robust chi2 of the pairs still in the graph after the last optimize() */
  double lambda_final;
  int32_t n_inliers;  /* the reference's return value nIn */
  int32_t n_corr;     /* nCorrespondences */
  int32_t n_bad;      /* pairs removed after the first stage */
  int32_t iterations; /* LM iterations run in total */
} VieoSim3Result;
/* A batch of candidates, one thread block each.  Per match of the shared arrays: Xc1 / Xc2 [M][3] f64 = P3D1c / P3D2c (the
 * float products of :2771-2783, cast to double), obs1 / obs2 [M][2] f32 (kpUn.pt), inv_sigma2_1 / _2 [M] f32.  Outputs: res
 * [n], keep [M] u8 (0: vpMatches1[i] = nullptr), chi2_12 / chi2_21 [M] f64 (last e->chi2() of the pair; may be NULL in the
 * host-buffer form).  _dev: scratch_dev [M] bytes. */
int vieo_optimize_sim3_batch(const VieoSim3Problem* pbs, int n, const VieoCamera* cam, const double* Xc1, const double* Xc2,
                             const float* obs1, const float* obs2, const float* inv_sigma2_1, const float* inv_sigma2_2,
                             int n_matches, VieoSim3Result* res, uint8_t* keep, double* chi2_12, double* chi2_21, int device);
int vieo_optimize_sim3_batch_dev(const VieoSim3Problem* pbs_dev, int n, const VieoCamera* cam_dev, const double* Xc1_dev,
                                 const double* Xc2_dev, const float* obs1_dev, const float* obs2_dev,
                                 const float* inv_sigma2_1_dev, const float* inv_sigma2_2_dev, VieoSim3Result* res_dev,
                                 uint8_t* keep_dev, double* chi2_12_dev, double* chi2_21_dev, uint8_t* scratch_dev, void* stream);

/* ---- Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:2309-2688): the loop closer's pose-graph optimisation ----------
 * Replaces the g2o graph of the routine (BlockSolver_7_3 + LinearSolverEigen + Levenberg, setUserLambdaInit(1e-16),
 * optimize(20), :2316-2334, :2618-2620): VertexSim3Expmap per keyframe (types_seven_dof_expmap.h:22-66, estimate = vScw,
 * _fix_scale = bFixScale, the loop keyframe fixed), EdgeSim3 per collected edge (error log(Sji Siw Sjw^-1), :101-129; its
 * Jacobians are g2o's central differences with delta 1e-9, core/base_binary_edge.hpp:123-187), 7x7 information per edge
 * (identity except the reduced pure-odometry edges, :2507-2541).  The caller keeps the edge-collection policy (loop
 * connections, spanning tree, earlier loop edges, covisibility >= 100: :2396-2615; template
 * VIEO_SLAM_B200::flatten::OptimizeEssentialGraph in vieo_flatten.hpp) and the write-back under mMutexMapUpdate.
 * g2o::Sim3 (types/sim3.h:40-57): rotation in Eigen's coefficient order (x, y, z, w), translation, scale. */
typedef struct VieoSim3 {
  double q[4];
  double t[3];
  double s;
} VieoSim3;
typedef struct VieoPoseGraphStats {
  double chi2_initial; /* activeChi2 before the first iteration */
  double chi2_final;   /* after the last accepted step */
  double lambda_final;
  int32_t iterations;  /* LM iterations run (<= the argument: g2o stops after 3 iterations gaining < 0.1 %) */
  int32_t trials;      /* damped solves in total */
  int32_t n_free;      /* free vertices with at least one active edge */
  int32_t ok;          /* 0: a factorisation met a non-positive pivot (the trial was rejected, like LDLT failing) */
} VieoPoseGraphStats;
/* Scw [n_vertices]: initial estimates; fixed [n_vertices] u8; edge_i / edge_j [n_edges]: vertex 0 / vertex 1 of the EdgeSim3;
 * Sji [n_edges]: measurements; info: [n_edges][49] row-major or NULL (identity for every edge); lambda_init <= 0: g2o's
 * 1e-5 * max diag(H).  Outputs: Scw_out [n_vertices] (vertices without an active edge and fixed ones are copied), Tcw_out
 * [n_vertices][12] row-major 3x4 [R | t / s] (the "SE3 Pose Recovering" of :2624-2642; may be NULL), stats.
 * The whole Levenberg-Marquardt loop is ONE cooperative kernel launch (device-wide barriers between the phases). */
int vieo_essential_graph_optimize(int n_vertices, const VieoSim3* Scw, const uint8_t* fixed, int fix_scale, int n_edges,
                                  const int32_t* edge_i, const int32_t* edge_j, const VieoSim3* Sji, const double* info,
                                  int iterations, double lambda_init, VieoSim3* Scw_out, double* Tcw_out,
                                  VieoPoseGraphStats* stats, int device);
/* Test hook: builds the system at the initial estimate, applies ONE damped step with `lambda` and returns; H_out [n][n] /
 * b_out [n] (n = 7 * stats->n_free, free vertices in ascending index; may be NULL) receive the undamped system. */
int vieo_essential_graph_debug_step(int n_vertices, const VieoSim3* Scw, const uint8_t* fixed, int fix_scale, int n_edges,
                                    const int32_t* edge_i, const int32_t* edge_j, const VieoSim3* Sji, const double* info,
                                    double lambda, VieoSim3* Scw_out, VieoPoseGraphStats* stats, double* H_out,
                                    double* b_out, int device);
/* "Correct points" (:2645-2676): Pw_out[i] = (Scw_after[ref[i]]^-1).map(Scw_before[ref[i]].map(Pw[i])), positions are float
 * (MapPoint::Tdata), arithmetic in double.  ref[i] = mnCorrectedReference or the reference keyframe's index (:2655-2664). */
int vieo_essential_graph_correct_points(int n_points, const float* Pw, const int32_t* ref, int n_vertices,
                                        const VieoSim3* Scw_before, const VieoSim3* Scw_after, float* Pw_out, int device);

/* ---- local bundle adjustment: PR-V-Bias vertices per keyframe, marginalised map points --------------------
 * The flattened graph of Optimizer::LocalBundleAdjustmentNavStatePRV (src/Optimizer.cc:133-520).  Keyframes
 * ("states") come local-first in ascending id (the reference's vertex ids 3k, 3k+1, 3k+2), then the fixed ones.
 * Visual edges must be sorted by point index (they are created point by point, :367-520). */
typedef struct VieoBaProblem {
  int32_t n_states, n_points, n_edges, n_imu;
  const VieoNavState* states;
  const uint8_t* state_flags; /* bit0: PR vertex fixed; bit1: has V and Bias vertices; bit2: V / Bias fixed */
  const double* points;       /* [P][3] MapPoint::GetWorldPos cast to double */
  const int32_t* edge_state;  /* [E] keyframe of the observation */
  const int32_t* edge_point;  /* [E] ascending */
  const float* obs;           /* [E][3] (ul, vl, ur) kpUn.pt / vuright_ */
  const float* inv_sigma2;    /* [E] vinvlevelsigma2_[octave] */
  const uint8_t* edge_flags;  /* [E] VIEO_EDGE_* */
  const int32_t* imu_i;       /* [n_imu] state of pKF0 (previous keyframe) */
  const int32_t* imu_j;       /* [n_imu] state of pKF1 */
  const VieoImuPreint* preint;/* [n_imu] pKF1->GetIMUPreInt(); dt == 0: bias edge only */
  const double* imu_dt_kf;    /* [n_imu] pKF1->ftimestamp_ - pKF0->ftimestamp_ */
  double gw[3];
  double inv_sigma_bg2, inv_sigma_ba2;
  int32_t large;       /* bLarge */
  int32_t rec_init;    /* bRecInit */
  int32_t visual_only; /* Optimizer::LocalBundleAdjustment: PR vertices only, no inertial edges */
  int32_t global_ba;   /* 0; vieo_global_ba_prv[_ex] sets bit 0 (graph of GlobalBundleAdjustmentNavStatePRV), bit 1 = bRobust,
                          bit 2 = VertexScale + EdgeReprojectPRS[Stereo] (bScaleOpt), bit 3 = VertexGThetaXYRwI +
                          EdgeNavStatePRVG, bit 4 = the prior-bias edge on states[0] (pimu_initiator sets 3 and 4); bits
                          2 - 4 need a handle from vieo_ba_create_global */
  double scale_init;   /* estimate VertexScale starts from (bit 2); <= 0 means 1 (setEstimate(1.), src/Optimizer.cc:845) */
} VieoBaProblem;
typedef struct VieoBaResult {
  double err0, err_end; /* activeRobustChi2 before / after (src/Optimizer.cc:539, 652), rounded to float like the reference */
  double lambda_final;
  int32_t iterations[2]; /* LM iterations of the two optimize() stages */
  int32_t accepted;      /* 0: the "FAIL LOCAL-INERTIAL BA" guard (:663-666) rejected the result, nothing to write back */
  int32_t n_erase;
} VieoBaResult;
typedef struct vieo_ba vieo_ba_t;
int vieo_ba_create(int max_states, int max_points, int max_edges, int max_imu, int device, vieo_ba_t** out);
/* Handle for map-sized problems (vieo_global_ba_prv): every one of up to max_states (<= 768) keyframes may be free; the
 * reduced camera system is kept dense in HBM and factorised by multi-CTA kernels.  vieo_ba_create handles keep at most 48
 * free keyframes (local windows) and run an LM trial as one CUDA graph. */
int vieo_ba_create_global(int max_states, int max_points, int max_edges, int max_imu, int device, vieo_ba_t** out);
void vieo_ba_destroy(vieo_ba_t* h);
/* Stream priority of the engines created AFTER the call (process-wide): 1 = highest (default: one LocalMapping thread
 * waits for its window, its chain of small kernels must not queue behind the tracker's), 0 = normal (many windows kept in
 * flight beside the front-end: the front-end is the critical path). */
void vieo_ba_stream_priority(int high);
/* Sharding over GPUs: each rank holds a subset of the points with all their edges, the keyframe states
 * replicated, the inertial edges on rank 0.  `allreduce` sums `count` doubles at device pointer `buf` in place over
 * all ranks, ordered on `stream` (the host wraps ncclAllReduce / torch.distributed.all_reduce).  NULL: single GPU. */
typedef int (*vieo_allreduce_fn)(void* ctx, double* buf_dev, size_t count, void* stream);
int vieo_ba_set_sharding(vieo_ba_t* h, int rank, int world, vieo_allreduce_fn allreduce, void* ctx);
void* vieo_ba_stream(vieo_ba_t* h);
/* The same exchange issued by the library itself: an NCCL communicator created from a 128-byte unique id (rank 0 makes it
 * with vieo_comm_unique_id, the launcher hands it to the other ranks — torch.distributed broadcast, MPI, a file).  With
 * vieo_ba_set_comm the handle calls ncclAllReduce(sum, fp64) on its own stream: no host callback in the data path, and
 * the abort flag (pbStopFlag) is folded into the all-reduced trial record so that every rank stops at the same LM trial.
 * libnccl.so.2 is resolved at run time (dlopen); single-GPU users never load it. */
#define VIEO_COMM_ID_BYTES 128
typedef struct vieo_comm vieo_comm_t;
int vieo_comm_unique_id(uint8_t id[VIEO_COMM_ID_BYTES]);
int vieo_comm_create(const uint8_t id[VIEO_COMM_ID_BYTES], int rank, int world, int device, vieo_comm_t** out);
void vieo_comm_destroy(vieo_comm_t* c);
int vieo_comm_allreduce_f64(vieo_comm_t* c, double* buf_dev, size_t count, void* stream);
int vieo_ba_set_comm(vieo_ba_t* h, vieo_comm_t* comm /* NULL: single GPU */);
/* The whole reference routine from "Setup optimizer" to the err/err_end guard on the flattened problem: Chi2LargeSetLevel,
 * optimize(optit[0]), inlier re-classification + kernel removal, optimize(optit[1]), outlier list.
 *   stop      pbStopFlag (mbAbortBA), polled before optimising and every LM iteration; may be NULL
 *   outputs   states_out [n_states], points_out [P][3], edge_chi2 [E], erase [E] (1 = ErasePairObs candidate) */
int vieo_local_ba_prv(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, const volatile uint8_t* stop,
                      VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase, VieoBaResult* res);
/* Asynchronous form of the same routine: _begin uploads the problem and enqueues EVERYTHING on the handle's stream — both
 * optimize() stages run as device-side Levenberg-Marquardt loops (CUDA graph with a WHILE node), the inlier
 * re-classification between them and the outlier pass are kernels, the results land in pinned staging — and returns without
 * waiting; _end waits (forwarding the abort flag to the device while it does), applies the "FAIL LOCAL-INERTIAL BA" guard and
 * fills the outputs; _poll returns 1 once _end would not block.  pb's arrays may be released after _begin.  One host thread
 * can keep many windows in flight (LocalMapping + several sessions): _batch begins all n, then ends all n.
 * Sharded handles run synchronously inside _begin (their exchange decisions need the host). */
int vieo_local_ba_prv_begin(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, const volatile uint8_t* stop);
int vieo_local_ba_prv_poll(vieo_ba_t* h);
int vieo_local_ba_prv_end(vieo_ba_t* h, VieoNavState* states_out, double* points_out, double* edge_chi2, uint8_t* erase,
                          VieoBaResult* res);
int vieo_local_ba_prv_batch(vieo_ba_t* const* hs, const VieoBaProblem* const* pbs, const VieoCamera* cam, int n,
                            const volatile uint8_t* stop, VieoNavState* const* states_out, double* const* points_out,
                            double* const* edge_chi2, uint8_t* const* erase, VieoBaResult* res);
/* Optimizer::GlobalBundleAdjustmentNavStatePRV (src/Optimizer.cc:771-1342; bScaleOpt = false, no IMU initiator): every
 * keyframe carries PR / V / Bias vertices (keyframe 0: state_flags 1|2|4, the others 2), IMU + bias-walk edges between
 * consecutive keyframes (information x 1e-2 where the previous bias vertex is fixed, :955-958, :979-983), reprojection
 * edges of every map point, Huber kernels (sqrt(16.919), sqrt(12.592), sqrt(5.99) / sqrt(7.815), :903-904, :1042-1043)
 * only when `robust`; one optimize(n_iterations) from g2o's own initial lambda; no outlier pass.  The handle must come
 * from vieo_ba_create_global.  Returns the number of LM iterations run (>= 0) or an error; res->err0 / err_end =
 * activeRobustChi2 before / after.  edge_chi2 may be NULL. */
int vieo_global_ba_prv(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, int n_iterations, int robust,
                       const volatile uint8_t* stop, VieoNavState* states_out, double* points_out, double* edge_chi2,
                       VieoBaResult* res);
/* The two other callers of GlobalBundleAdjustmentNavStatePRV (src/Optimizer.cc:771-1342):
 *   scale_opt  bScaleOpt = true — System::FinalGBA (src/System.cc:24-29): VertexScale (g2otypes.h:294-311, :843-851)
 *              seeded with 1 as one more row / column of the reduced camera system, every reprojection edge an
 *              EdgeReprojectPRS[Stereo] (:1132-1200; Xw = s X, J_scale = Jproj Rcw X, g2otypes.h:517-521); points_out are
 *              multiplied by the recovered scale (:1311-1334), `scale` returns the vertex estimate;
 *   imu_init   pimu_initiator != nullptr — IMUInitialization::Run (src/Odom/IMUInitialization.cpp:475; the caller passes
 *              robust = 0 and state_flags 1|2 for keyframe 0: PR fixed, V / Bias free, :825-831): VertexGThetaXYRwI
 *              (g2otypes.h:674-698, :852-865) seeded from `gw`, EdgeNavStatePRVG on every keyframe pair (:955-957), one
 *              EdgeNavStateBias from a fixed copy of states[0]'s bias with information invSigma / sum(dt) (:866-900,
 *              :1026-1054; states[0] must be the earliest keyframe); `gw` returns RwI * GI (:1262-1275).
 * ex == NULL is vieo_global_ba_prv. */
typedef struct VieoGbaExtra {
  int32_t scale_opt, imu_init;
  double scale; /* out */
  double gw[3]; /* in (imu_init) / out */
} VieoGbaExtra;
int vieo_global_ba_prv_ex(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam, int n_iterations, int robust,
                          VieoGbaExtra* ex, const volatile uint8_t* stop, VieoNavState* states_out, double* points_out,
                          double* edge_chi2, VieoBaResult* res);
/* Estimates of the border vertices of the current problem: VertexScale (1 without it) and RwI * GI (gw without it). */
int vieo_ba_get_border(vieo_ba_t* h, double* scale_out, double* gw_out /* nullable [3] */);
/* Building blocks (what `optimizer.optimize(n)` and friends do), for callers that keep the policy on their side. */
int vieo_ba_set_problem(vieo_ba_t* h, const VieoBaProblem* pb, const VieoCamera* cam);
int vieo_ba_chi2_large_set_level(vieo_ba_t* h, float rat);           /* GraphOperator::Chi2LargeSetLevel */
int vieo_ba_active_robust_chi2(vieo_ba_t* h, int recompute, double* chi2); /* [computeActiveErrors +] activeRobustChi2 */
int vieo_ba_optimize(vieo_ba_t* h, int iterations, double lambda_init, const volatile uint8_t* stop); /* -> iterations run */
int vieo_ba_reclassify(vieo_ba_t* h, int remove_kernels, uint8_t* bad_host /* nullable [E] */); /* chi2 / depth gates -> level 1 */
int vieo_ba_get(vieo_ba_t* h, VieoNavState* states_out, double* points_out, double* edge_chi2);
/* One damped Gauss-Newton step at the current estimate without applying it: x_pose [np] (Hessian index order),
 * x_points [P][3]; returns np.  H_out (np x np, nullable) / b_out receive the pose block of the normal equations
 * (vieo_ba_get_hessian_blocks of SURVEY.md 8b). */
int vieo_ba_debug_step(vieo_ba_t* h, double lambda, double* x_pose, double* x_points, double* H_out, double* b_out);
int vieo_ba_last_launches(const vieo_ba_t* h);
/* Device time (CUDA events on the handle's stream, upload to last download) of the last vieo_local_ba_prv[_begin/_end]
 * call that took the asynchronous path, and the LM trials of its last optimize() stage — for bench.py's roofline. */
double vieo_ba_last_ms(const vieo_ba_t* h);
int vieo_ba_last_trials(const vieo_ba_t* h);

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:896-1150; called by
 * LocalMapping::CreateNewMapPoints for every neighbour keyframe, src/LocalMapping.cc:709) for single-pinhole keyframes
 * (usedistort_ == false, one camera each): the DBoW2::FeatureVector walk over common vocabulary nodes (:964-1135), per
 * unmatched keypoint of KF1 the Hamming arg-min (<= TH_LOW, equal distances: the later candidate wins) over the node's
 * unmatched keypoints of KF2 that pass the mono-mono epipole-distance rule (:1031-1035) and
 * GeometricCamera::epipolarConstrain (common/camera_models/camera_base.h:360-404), each KF2 keypoint matched at most once
 * in the reference's visiting order (:1001-1002), the rotation histogram with ComputeThreeMaxima (:1137-1156).
 * A batch of keyframe pairs, device resident, one CTA per pair.  Shared arrays: kps (mvKeysUn), uright (vuright_), desc
 * (mDescriptors), has_mp (GetMapPoint(idx) != nullptr) per keypoint; the FeatureVectors flattened: fv_node = node ids in
 * std::map order (ascending), fv_ptr = n_nodes + 1 offsets per vector (relative to its idx*_begin), fv_idx = keypoint
 * indices in push order.  The shim forms F12 = K1^-T [t12]x R12 K2^-1 in double (Eigen inverse(), host side) and the
 * epipole (ex, ey) of KF1's centre in KF2 (:908-927) per pair. */
typedef struct VieoSftPair {
  int32_t kp1_begin, n_kp1, kp2_begin, n_kp2;           /* keypoint ranges of the two keyframes (<= 8192 keypoints each) */
  int32_t node1_begin, n_nodes1, node2_begin, n_nodes2; /* ranges in fv_node */
  int32_t ptr1_begin, ptr2_begin;                       /* first of the n_nodes + 1 entries in fv_ptr */
  int32_t idx1_begin, idx2_begin;                       /* ranges in fv_idx */
  int32_t out_begin;                                    /* this pair's n_kp1 slots in match12 / pairs_out (prefix sum of n_kp1) */
  int32_t nscr_begin;                                   /* prefix sum of n_nodes1 (scratch) */
  int32_t only_stereo, check_orientation;               /* bOnlyStereo, mbCheckOrientation */
  float ex, ey;                                         /* epipole in KF2 */
  float scale_factor2[16];                              /* pKF2->scalepyrinfo_.vscalefactor_ */
  float level_sigma2_2[16];                             /* pKF2->scalepyrinfo_.vlevelsigma2_ */
  double F12[9];                                        /* row-major */
} VieoSftPair;
size_t vieo_sft_scratch_bytes(int n_out_total, int n_nodes1_total);
/* match12 [n_out_total]: per KF1 keypoint of each pair the matched KF2 keypoint or -1 (after the rotation filter);
 * pairs_out [n_out_total][2]: vMatchedPairs as (idx1, idx2) in the reference's creation order, n_matches[p] of them for
 * pair p (its return value; -1: a size limit was exceeded). */
int vieo_search_for_triangulation_dev(const VieoSftPair* pairs_dev, int n_pairs, const VieoKeyPoint* kps_dev,
                                      const float* uright_dev, const uint8_t* desc_dev, const uint8_t* has_mp_dev,
                                      const int32_t* fv_node_dev, const int32_t* fv_ptr_dev, const int32_t* fv_idx_dev,
                                      int n_out_total, int n_nodes1_total, int32_t* match12_dev, int32_t* pairs_out_dev,
                                      int32_t* n_matches_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
/* Host-buffer form (every range of every pair is validated against the array sizes given). */
int vieo_search_for_triangulation(const VieoSftPair* pairs, int n_pairs, const VieoKeyPoint* kps, const float* uright,
                                  const uint8_t* desc, const uint8_t* has_mp, const int32_t* fv_node, const int32_t* fv_ptr,
                                  const int32_t* fv_idx, int n_kp_total, int n_node_total, int n_ptr_total, int n_idx_total,
                                  int n_out_total, int n_nodes1_total, int32_t* match12, int32_t* pairs_out, int32_t* n_matches,
                                  int device);

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:344-505;
 * Tracking::TrackReferenceKeyFrame, Relocalization, LoopClosing), single camera: FeatureVector walk, per keyframe keypoint
 * with a live map point the best / second-best Hamming distance over the not-yet-matched frame keypoints of the same node,
 * accepted when best <= TH_LOW and best < mfNNratio * second, rotation histogram.  Same flattened FeatureVector layout as
 * VieoSftPair; keyframe = side 1, frame = side 2; mp_ok [keypoint] = GetMapPoint(idx) && !isBad() (keyframe side).
 * One camera per keyframe means a map point sits on one keyframe keypoint (the (pMP, img_id) table of :363 never fires);
 * the host-buffer form does not check that.  match_f [n_out_total]: per frame keypoint of each pair (out_begin = prefix sum
 * of n_kp2) the keyframe keypoint whose map point it received, -1 none; n_matches [n_pairs] = the return value. */
typedef struct VieoBowPair {
  int32_t kp1_begin, n_kp1, kp2_begin, n_kp2;
  int32_t node1_begin, n_nodes1, node2_begin, n_nodes2;
  int32_t ptr1_begin, ptr2_begin, idx1_begin, idx2_begin;
  int32_t out_begin;
  int32_t check_orientation; /* mbCheckOrientation */
  float nn_ratio;            /* mfNNratio */
  int32_t pad_;
} VieoBowPair;
int vieo_search_by_bow_dev(const VieoBowPair* pairs_dev, int n_pairs, const VieoKeyPoint* kps_dev, const uint8_t* desc_dev,
                           const uint8_t* mp_ok_dev, const int32_t* fv_node_dev, const int32_t* fv_ptr_dev,
                           const int32_t* fv_idx_dev, int n_out_total, int32_t* match_f_dev, int32_t* n_matches_dev,
                           void* scratch_dev /* >= n_out_total bytes */, size_t scratch_bytes, void* stream);
int vieo_search_by_bow(const VieoBowPair* pairs, int n_pairs, const VieoKeyPoint* kps, const uint8_t* desc, const uint8_t* mp_ok,
                       const int32_t* fv_node, const int32_t* fv_ptr, const int32_t* fv_idx, int n_kp_total, int n_node_total,
                       int n_ptr_total, int n_idx_total, int n_out_total, int32_t* match_f, int32_t* n_matches, int device);

/* ------------------------------------------------------------------------------------------------
 * SM partitions (CUDA green contexts).  The reference runs Tracking, LocalMapping and LoopClosing on separate CPU
 * threads (src/System.cc); here they share one device, and the BA engines' chains of tiny ordered kernels starve beside
 * the front-end's saturating ones unless they own a few SMs.  A partition splits the device into `ba_sms` SMs for the
 * BA streams (VIEO_SM_BA) and the rest for the front-end / tracking streams (VIEO_SM_FRONTEND).  Handles
 * (vieo_orb_create, vieo_frontend_create, vieo_ba_create) and the per-thread staging streams of the host-buffer calls are
 * created inside the partition the CALLING THREAD is bound to at that moment; unbound threads get ordinary streams. */
#define VIEO_SM_FRONTEND 0
#define VIEO_SM_BA 1
typedef struct vieo_sm_partition vieo_sm_partition_t;
int vieo_sm_partition_create(int device, int ba_sms, vieo_sm_partition_t** out);
void vieo_sm_partition_destroy(vieo_sm_partition_t* p);
int vieo_sm_partition_sms(const vieo_sm_partition_t* p, int which);
int vieo_sm_partition_bind_thread(vieo_sm_partition_t* p, int which); /* p == NULL: unbind */
void* vieo_sm_partition_stream(vieo_sm_partition_t* p, int which, int high_priority); /* a new stream for *_dev callers */

/* ------------------------------------------------------------------------------------------------
 * Guided searches of the tracking thread: ORBmatcher::SearchByProjection(Frame&, const Frame& last, th, bMono, th_far)
 * (src/ORBmatcher.cc:1303-1467, VIEO_SBP_LAST_FRAME) and ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, th,
 * th_far) (:230-335, VIEO_SBP_LOCAL_MAP), including FrameBase::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid /
 * IsInImage (src/FrameBase.cpp:95-174), the stereo ur gate, the Hamming arg-min (best / second best with the same-level
 * ratio test), the greedy one-keypoint-one-point rule (a keypoint taken by a map point with Observations() > 0 is
 * skipped by later points, :1400-1401 / :289-291) and the rotation-histogram check (:1445-1464, ComputeThreeMaxima
 * :1608-1641).  Single pinhole camera (usedistort_ == false).  A batch holds n_frames independent current frames. */
#define VIEO_SBP_LAST_FRAME 0
#define VIEO_SBP_LOCAL_MAP 1
#define VIEO_SBP_MAX_KEYPOINTS 4096
typedef struct VieoSbpFrame {
  int32_t kp_begin, n_kp;       /* this frame's keypoints in the keypoint arrays (n_kp <= VIEO_SBP_MAX_KEYPOINTS) */
  int32_t q_begin, n_q;         /* this frame's queries = map points in the reference's loop order */
  float minx, maxx, miny, maxy; /* FrameBase::gridinfo_.minmax_xy_ */
  float grid_winv, grid_hinv;   /* gridinfo_.fgrids_widthinv_ / fgrids_heightinv_ (64 x 48 cells) */
  float bf, b;                  /* stereoinfo_.baseline_bf_[1] / [0] */
  float fx, fy, cx, cy;         /* mpCameras[0]->toK() cast to float */
  float th, th_far;             /* window factor; th_far_pts (<= 0: off) */
  float nn_ratio;               /* mfNNratio (LOCAL_MAP) */
  int32_t mono, check_orientation, n_levels; /* bMono, mbCheckOrientation (LAST_FRAME) */
  float scale[16];              /* scalepyrinfo_.vscalefactor_ */
  double qcw[4], tcw[3];        /* CurrentFrame.GetTcwCst(): unit quaternion (w, x, y, z), translation (LAST_FRAME) */
  double qlw[4], tlw[3];        /* LastFrame.GetTcwCst() (LAST_FRAME) */
} VieoSbpFrame;
typedef struct VieoSbpQueries { /* arrays over all queries of the batch; unused ones may be null */
  const double* Xw;       /* [n][3] LAST_FRAME: MapPoint::GetWorldPos() cast to double */
  const int32_t* level;   /* LAST_FRAME: LastFrame.mvKeys[i].octave; LOCAL_MAP: vtrack_scalelevel_, -1 = !btrack_inview_
                             (the query is skipped, src/ORBmatcher.cc:244) */
  const float* angle;     /* LAST_FRAME: LastFrame.mvKeys[i].angle */
  const float* proj;      /* [n][3] LOCAL_MAP: vtrack_proj_ (u, v, ur) left by Frame::isInFrustum */
  const float* viewcos;   /* LOCAL_MAP: vtrack_viewcos_ */
  const float* depth;     /* LOCAL_MAP: track_depth_ */
  const uint8_t* desc;    /* [n][32] MapPoint::GetDescriptor(), 16-byte aligned */
  const uint8_t* flags;   /* bit 0: Observations() > 0 */
} VieoSbpQueries;
/* Outputs: kp_match[keypoint] = frame-relative query that owns the keypoint at the end (AddMapPoint), -1 = untouched or
 * erased by the rotation check; q_match / q_dist [query] = frame-relative keypoint the query took and its distance
 * (-1 / 256) before the rotation check; n_matches[frame] = the reference's return value.  kp_blocked (nullable):
 * keypoints that already hold a map point with Observations() > 0 on entry. */
size_t vieo_sbp_scratch_bytes(int n_queries_total);
int vieo_sbp_batch_dev(int mode, const VieoSbpFrame* frames_dev, int n_frames, const VieoKeyPoint* kps_dev,
                       const float* uright_dev, const uint8_t* desc_dev, const VieoSbpQueries* q_of_dev_pointers,
                       const uint8_t* kp_blocked_dev, int32_t* kp_match_dev, int32_t* q_match_dev, int32_t* q_dist_dev,
                       int32_t* n_matches_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
int vieo_sbp_batch(int mode, const VieoSbpFrame* frames, int n_frames, const VieoKeyPoint* kps, const float* uright,
                   const uint8_t* desc, const VieoSbpQueries* q, const uint8_t* kp_blocked, int32_t* kp_match,
                   int32_t* q_match, int32_t* q_dist, int32_t* n_matches, int device);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist,
 * th_far_pts) (src/ORBmatcher.cc:1471-1606): Tracking::Relocalization's guided search after the PnP pose, on the same two
 * kernels (third window rule: level = MapPoint::PredictScale of the camera-centre distance inside the 0.8 / 1.2 invariance
 * range, band [level - 1, level + 1], no stereo gate, every accepted match claims its keypoint, acceptance bestDist <=
 * ORBdist, rotation histogram with pKF->mvKeys[i].angle).  frames[f]: keypoint / query ranges, image bounds, grid, fx..cy,
 * th, th_far, check_orientation, n_levels, scale[], qcw / tcw = CurrentFrame.GetTcwCst(); the other fields are unused.
 * Queries = pKF's map points that are good and not in sAlreadyFound (the host filters, :1487-1489), in keypoint order:
 * q_Xw [n][3] f64, q_angle = pKF->mvKeys[i].angle, q_max_dist / q_min_dist = mfMaxDistance / mfMinDistance, q_desc [n][32].
 * kp_blocked: keypoints of CurrentFrame that hold a map point on entry.  Outputs as vieo_sbp_batch, plus q_level
 * (nullable) = nPredictedLevel (-1: the point failed a geometric test).  Single camera, usedistort_ == false. */
#define VIEO_SBP_RELOC 2
typedef struct VieoSbpReloc {
  int32_t orb_dist;       /* ORBdist (< 256) */
  float log_scale_factor; /* CurrentFrame.scalepyrinfo_.flogscalefactor_ */
  float level_ratio[16];  /* vieo_frustum_level_table(log_scale_factor, n_levels): filled by the host-buffer call, by the
                             caller for _dev */
} VieoSbpReloc;
int vieo_sbp_reloc_batch(const VieoSbpFrame* frames, const VieoSbpReloc* reloc, int n_frames, const VieoKeyPoint* kps,
                         const uint8_t* desc, const double* q_Xw, const float* q_angle, const float* q_max_dist,
                         const float* q_min_dist, const uint8_t* q_desc, const uint8_t* kp_blocked, int32_t* kp_match,
                         int32_t* q_match, int32_t* q_dist, int32_t* q_level, int32_t* n_matches, int device);
int vieo_sbp_reloc_batch_dev(const VieoSbpFrame* frames_dev, const VieoSbpReloc* reloc_dev, int n_frames,
                             const VieoKeyPoint* kps_dev, const uint8_t* desc_dev, const double* q_Xw_dev,
                             const float* q_angle_dev, const float* q_max_dist_dev, const float* q_min_dist_dev,
                             const uint8_t* q_desc_dev, const uint8_t* kp_blocked_dev, int32_t* kp_match_dev,
                             int32_t* q_match_dev, int32_t* q_dist_dev, int32_t* q_level_dev, int32_t* n_matches_dev,
                             void* scratch_dev, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Frame::isInFrustum (src/Frame.cc:335-416) with MapPoint::PredictScale (src/MapPoint.cc:491-509) for every candidate map
 * point of a batch of frames, and Tracking::SearchLocalPoints' pair "visibility test -> SearchByProjection(F,
 * vpMapPoints, th, th_far)" (src/Tracking.cc:2308-2368) as one call.  Single camera
 * (mpCameras.size() == 1), usedistort_ == false; float arithmetic in Eigen's evaluation order. */
typedef struct VieoFrustumFrame {
  int32_t q_begin, n_q;         /* this frame's candidate map points in the point arrays */
  float Rcw[9], tcw[3], Ow[3];  /* Tcw_ rotation (row-major), mtcw, mOw cast to float */
  float fx, fy, cx, cy;         /* mpCameras[0]->toK() cast to float */
  float minx, maxx, miny, maxy; /* gridinfo_.minmax_xy_ */
  float bf;                     /* stereoinfo_.baseline_bf_[1] */
  float cos_limit;              /* viewingCosLimit (0.5 in SearchLocalPoints) */
  float log_scale_factor;       /* scalepyrinfo_.flogscalefactor_ */
  int32_t n_levels;             /* scalepyrinfo_.vscalefactor_.size() (<= 16) */
  float level_ratio[16];        /* vieo_frustum_level_table(log_scale_factor, n_levels): the host-buffer calls fill it,
                                   *_dev callers fill it themselves */
} VieoFrustumFrame;
/* table[k] = the smallest float ratio for which ceil(logf(ratio) / log_scale_factor) >= k, found with the host libm's
 * logf (the function the reference calls), so that counting thresholds on the device reproduces PredictScale exactly. */
int vieo_frustum_level_table(float log_scale_factor, int n_levels, float table[16]);
/* Per point: wP / normal [n][3] (GetWorldPos / GetNormal), max_dist / min_dist = mfMaxDistance / mfMinDistance (the 1.2 /
 * 0.8 factors of Get*DistanceInvariance are applied here), skip (nullable) != 0: already matched in the frame or bad.
 * Outputs per point: inview = btrack_inview_, proj [n][3] = (u, v, ur), level = vtrack_scalelevel_ (-1 when not in view),
 * viewcos, depth = track_depth_; n_inview [n_frames] = nToMatch. */
int vieo_frustum_batch_dev(const VieoFrustumFrame* frames_dev, int n_frames, const float* wP_dev, const float* normal_dev,
                           const float* max_dist_dev, const float* min_dist_dev, const uint8_t* skip_dev,
                           uint8_t* inview_dev, float* proj_dev, int32_t* level_dev, float* viewcos_dev, float* depth_dev,
                           int32_t* n_inview_dev, void* stream);
int vieo_frustum_batch(const VieoFrustumFrame* frames, int n_frames, const float* wP, const float* normal,
                       const float* max_dist, const float* min_dist, const uint8_t* skip, uint8_t* inview, float* proj,
                       int32_t* level, float* viewcos, float* depth, int32_t* n_inview, int device);
/* ---- Frame::isInFrustum with a camera rig (mpCameras.size() > 1, KB8 / multi-camera configs; src/Frame.cc:351-411) ----
 * Per camera cami: Pc = mpCameras[cami]->GetTcr() * Pcr (Sophus::SE3f: the float quaternion rotation of
 * common/so3_extra.h:102-104 + translation), twc = mOw + Rcrw^T GetTrc().translation(), the pixel by K * p_normalize in float
 * (usedistort_ == false: model 0) or by the camera's own Project evaluated in double and rounded to float (usedistort_: model 1
 * PinholeCamera, 2 KB8Camera, common/camera_models/camera_{pinhole,kb8}.h), per-camera bounds gridinfo_.minmax_xy_[cami];
 * distance / viewing-angle / PredictScale tests as in the single-camera form.  Every float operation in the reference's
 * order; the KB8 angle uses the device's fp64 atan2 (<= 2 ulp from the host's): the float pixel can differ in its last bit
 * when the double lands within 1e-16 relative of a float rounding boundary (about once in 1e8 projections). */
typedef struct VieoFrustumCam {
  float q_cr[4];                /* GetTcr().unit_quaternion() coefficients (x, y, z, w) */
  float t_cr[3];                /* GetTcr().translation() */
  float t_rc[3];                /* GetTrc().translation() */
  float fx, fy, cx, cy;
  float k[4];                   /* KB8 k1..k4 (model 2) */
  float minx, maxx, miny, maxy; /* gridinfo_.minmax_xy_[cami] */
  int32_t model;                /* 0, 1 or 2 (see above) */
  int32_t pad_;
} VieoFrustumCam;
typedef struct VieoFrustumRigFrame {
  int32_t q_begin, n_q;         /* this frame's candidate map points in the point arrays */
  float Rcw[9], tcw[3], Ow[3];  /* Tcw_ rotation (row-major), mtcw, mOw cast to float (the rig's reference frame) */
  float bf, cos_limit, log_scale_factor;
  int32_t n_levels, n_cams;     /* n_cams <= 4 */
  float level_ratio[16];        /* vieo_frustum_level_table: filled by the host-buffer call, by the caller for _dev */
  VieoFrustumCam cam[4];
} VieoFrustumRigFrame;
/* Outputs per point: inview = btrack_inview_; cam_mask bit c set = camera c pushed an entry into vtrack_* (the shim appends the
 * set slots in camera order: vtrack_cami_); per camera slot proj [n][4][3] = (u, v, ur), level [n][4] (-1 when not set),
 * viewcos [n][4]; depth [n] = track_depth_ (mean dist3D over the set cameras); n_inview [n_frames]. */
int vieo_frustum_rig_batch_dev(const VieoFrustumRigFrame* frames_dev, int n_frames, const float* wP_dev, const float* normal_dev,
                               const float* max_dist_dev, const float* min_dist_dev, const uint8_t* skip_dev, uint8_t* inview_dev,
                               uint8_t* cam_mask_dev, float* proj_dev, int32_t* level_dev, float* viewcos_dev, float* depth_dev,
                               int32_t* n_inview_dev, void* stream);
int vieo_frustum_rig_batch(const VieoFrustumRigFrame* frames, int n_frames, const float* wP, const float* normal,
                           const float* max_dist, const float* min_dist, const uint8_t* skip, uint8_t* inview, uint8_t* cam_mask,
                           float* proj, int32_t* level, float* viewcos, float* depth, int32_t* n_inview, int device);
/* Visibility test + local-map guided search on one stream; frames[f] and frustum[f] share q_begin / n_q, the search
 * reads the tracking info the first kernel left in HBM.  Outputs of both halves as documented above / at
 * vieo_sbp_batch; proj / level / viewcos / depth may ALL be null (Tracking::SearchLocalPoints only needs inview and the
 * matches): the tracking info then never leaves the device. */
int vieo_search_local_points(const VieoFrustumFrame* frustum, const VieoSbpFrame* frames, int n_frames, const float* wP,
                             const float* normal, const float* max_dist, const float* min_dist, const uint8_t* skip,
                             const uint8_t* q_desc, const uint8_t* q_flags, const VieoKeyPoint* kps, const float* uright,
                             const uint8_t* desc, const uint8_t* kp_blocked, uint8_t* inview, float* proj, int32_t* level,
                             float* viewcos, float* depth, int32_t* n_inview, int32_t* kp_match, int32_t* q_match,
                             int32_t* q_dist, int32_t* n_matches, int device);

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227), the search half behind ORBmatcher::Fuse(KeyFrame*,
 * vector<MapPoint*>, th) (:1152-1165, LocalMapping::SearchInNeighbors), Fuse(KeyFrame*, Scw, ...) (:1167-1220) and the
 * Sim3 / keyframe projection searches: per map point the projection into the keyframe, IsInImage,
 * the scale-invariance range, the 60-degree viewing cone (bCheckViewingAngle), PredictScale, GetFeaturesInArea(u, v,
 * th_radius * scale[level]), level band [level - 1, level], the chi-square gate against the keypoint (stereo 7.8 / mono
 * 5.99, only with pbf) and the strict-'<' Hamming arg-min.  What follows in the reference (bestDist <= th_bestdist,
 * FuseMP / AddObservation / vpReplacePoint, the IsInKeyFrame skip) only edits map-point links and stays on the host.
 * Single camera, usedistort_ == false.  One record per keyframe of the batch. */
typedef struct VieoProjSearchFrame {
  int32_t kp_begin, n_kp;       /* the keyframe's keypoints (mvKeysUn) in the keypoint arrays (<= VIEO_SBP_MAX_KEYPOINTS) */
  int32_t q_begin, n_q;         /* the map points projected into it */
  float Rcw[9], tcw[3], Ow[3];  /* Rcrw, tcrw, pKF->GetCameraCenter() cast to float */
  float fx, fy, cx, cy;         /* mpCameras[0]->toK() cast to float */
  float minx, maxx, miny, maxy; /* gridinfo_.minmax_xy_ */
  float grid_winv, grid_hinv;   /* gridinfo_.fgrids_widthinv_ / fgrids_heightinv_ */
  float bf;                     /* *pbf */
  int32_t use_bf;               /* pbf != nullptr: chi-square gate on */
  int32_t check_viewing_angle;  /* bCheckViewingAngle */
  float th_radius;
  int32_t n_levels;
  float log_scale_factor;       /* scalepyrinfo_.flogscalefactor_ */
  float scale[16];              /* scalepyrinfo_.vscalefactor_ */
  float inv_level_sigma2[16];   /* scalepyrinfo_.vinvlevelsigma2_ */
  float level_ratio[16];        /* vieo_frustum_level_table; filled by the host-buffer call, by the caller for _dev */
} VieoProjSearchFrame;
/* Per map point: wP / normal [n][3], max_dist / min_dist = mfMaxDistance / mfMinDistance, q_desc [n][32] =
 * GetDescriptor(), q_skip (nullable) != 0: null / bad / already in the keyframe.  Outputs: best_idx = keyframe-relative
 * keypoint (-1: none), best_dist (256: none), level = nPredictedLevel (-1: the point failed a geometric test). */
int vieo_proj_search_batch(const VieoProjSearchFrame* frames, int n_frames, const VieoKeyPoint* kps, const float* uright,
                           const uint8_t* desc, const float* wP, const float* normal, const float* max_dist,
                           const float* min_dist, const uint8_t* q_desc, const uint8_t* q_skip, int32_t* best_idx,
                           int32_t* best_dist, int32_t* level, int device);
int vieo_proj_search_batch_dev(const VieoProjSearchFrame* frames_dev, int n_frames, const VieoKeyPoint* kps_dev,
                               const float* uright_dev, const uint8_t* desc_dev, const float* wP_dev, const float* normal_dev,
                               const float* max_dist_dev, const float* min_dist_dev, const uint8_t* q_desc_dev,
                               const uint8_t* q_skip_dev, int32_t* best_idx_dev, int32_t* best_dist_dev, int32_t* level_dev,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stereo front-end over HOST buffers — the hot work of the Frame::Frame stereo constructor
 * (src/Frame.cc:218-316): ORBextractor::operator() for both cameras (:259-278) and the brute-force
 * left->right knnMatch(k=2) of ComputeStereoFishEyeMatches (:620-628), for a batch of frames, with the
 * host<->device copies pipelined against the kernels on independent streams.
 *   imgs        [n_frames][2][height] rows of row_stride bytes (left, right), ideally pinned
 *   kps/desc    [2*n_frames][cap] / [2*n_frames][cap][32], cap = vieo_frontend_max_keypoints()
 *   n_kp        [2*n_frames]
 *   match_idx / match_dist   [n_frames][cap][2]: for left keypoint i its two nearest right descriptors */
typedef struct vieo_frontend vieo_frontend_t;
int vieo_frontend_create(const VieoOrbConfig* cfg, int max_frames, int device, vieo_frontend_t** out);
void vieo_frontend_destroy(vieo_frontend_t* f);
int vieo_frontend_max_keypoints(const vieo_frontend_t* f);
int vieo_frontend_last_launches(const vieo_frontend_t* f);
int vieo_frontend_process(vieo_frontend_t* f, int n_frames, const uint8_t* imgs, int row_stride, VieoKeyPoint* kps,
                          uint8_t* desc, int32_t* n_kp, int32_t* match_idx, int32_t* match_dist);
/* Frame::ComputeStereoMatches (rectified configs, src/Frame.cc:451-611) for the frames of the last
 * vieo_frontend_process call (their keypoints, descriptors and pyramids are still on the device).
 * uright / depth / sad: [n_frames][cap] host arrays indexed by left keypoint. */
int vieo_frontend_stereo_rectified(vieo_frontend_t* f, int n_frames, float bf, float min_z, float* uright, float* depth,
                                   int32_t* sad);

#ifdef __cplusplus
}
#endif
#endif /* VIEO_B200_H */

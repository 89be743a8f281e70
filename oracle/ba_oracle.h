// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  C interface of the CPU restatement of the
// bundle-adjustment path: edges of src/Odom/g2otypes.h / g2otypes.cpp, the g2o Levenberg-Marquardt + Schur solver
// semantics (optimizer/g2o/g2o/core) and the drivers Optimizer::PoseOptimization (visual and IMU/PVR) and
// Optimizer::LocalBundleAdjustmentNavStatePRV / GlobalBundleAdjustmentNavStatePRV.  Layouts are identical to
// include/vieo_b200.h so the parity tests feed both sides the same bytes.
#pragma once
#include <stdint.h>

#include "oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcNavState { /* NavState (src/Odom/NavState.h:17-36) */
  double p[3];           /* mpwb */
  double q[4];           /* mRwb unit quaternion (w, x, y, z) */
  double v[3];           /* mvwb */
  double bg[3], ba[3];   /* mbg, mba (frozen linearisation point) */
  double dbg[3], dba[3]; /* mdbg, mdba (optimised) */
} OrcNavState;

typedef struct OrcCamera { /* intrinsics are float in the reference (camera_pinhole.h:70-83) */
  float fx, fy, cx, cy, bf;
  int32_t model; /* 0 pinhole, 1 radtan (camera_radtan.h), 2 KB8 (camera_kb8.h) */
  int32_t num_k; /* radtan: number of radial coefficients */
  float pad_;
  float dist[8]; /* radtan: k1..k_num_k, p1, p2; KB8: k1..k4 */
  double Rcb[9], tcb[3]; /* Frame::meigRcb / meigtcb */
} OrcCamera;

/* Flags of a visual edge */
#define ORC_EDGE_STEREO 1 /* 3-dim EdgeReproject*Stereo, else 2-dim */
#define ORC_EDGE_CLOSE 2  /* track_depth_ < max(10, mThDepth): 1.5x chi2 gate */
#define ORC_EDGE_LEVEL1 4 /* level 1 (outlier) */
#define ORC_EDGE_NOKERNEL 8

typedef struct OrcPoseOptProblem {
  OrcNavState cur, last, prior; /* pFrame->mNavState, pLastKF->GetNavState(), pLastKF->mNavStatePrior */
  OrcImuPreint preint;          /* pFrame->GetIMUPreInt() (dt == 0: no IMU edge) */
  double prior_info[225];       /* pLastKF->mMargCovInv (row-major, order P V R bg ba) */
  double gw[3];                 /* gravity in world */
  double inv_sigma_bg2, inv_sigma_ba2; /* IMUDataBase::mInvSigmabg2 / mInvSigmaba2 */
  double dt_frames;             /* pFrame->ftimestamp_ - pLastKF->ftimestamp_ */
  int32_t mode;                 /* 0: visual, PR vertex (src/Optimizer.cc:1611); 1: IMU, PVR vertex (include/Optimizer.h:208) */
  int32_t last_has_prior;       /* pLastKF->mbPrior -> last frame's vertices are free */
  int32_t compute_marg;         /* bComputeMarg */
  int32_t no_mps;               /* bNoMPs */
  int32_t edge_begin, edge_end; /* this frame's range in the edge arrays */
} OrcPoseOptProblem;

typedef struct OrcPoseOptResult {
  OrcNavState cur;          /* optimised pFrame->mNavState */
  OrcNavState last;         /* optimised last-frame state (only meaningful when it was free) */
  double marg_cov_inv[225]; /* pFrame->mMargCovInv when compute_marg */
  double chi2_final;        /* activeRobustChi2 of the last optimize() */
  double lambda_final;
  int32_t n_inliers;        /* return value */
  int32_t n_initial;
  int32_t iterations;       /* total LM iterations run */
  int32_t prior_set;        /* mbPrior after the call */
} OrcPoseOptResult;

/* One frame.  Edge arrays (global; the problem names its range):
 *   Xw [E][3] map point positions (MapPoint::GetWorldPos cast to double), obs [E][3] (ul, vl, ur) as float,
 *   inv_sigma2 [E], flags [E] (STEREO, CLOSE).  Outputs: outlier [E] (mvbOutlier), chi2 [E] last e->chi2(). */
int orc_pose_optimization(const OrcPoseOptProblem* pb, const OrcCamera* cam, const double* Xw, const float* obs,
                          const float* inv_sigma2, const uint8_t* flags, OrcPoseOptResult* res, uint8_t* outlier,
                          double* chi2);

/* Test hook: when set, the next orc_pose_optimization with compute_marg also records what Optimizer::FillCovInv
 * (include/Optimizer.h:126-206) sees and yields — the information matrices and Huber deltas (< 0: no kernel) of the inertial /
 * bias / prior edges and of every visual edge at that moment, the visual edges' levels, and the three blocks it assembles
 * (schur_bec 0 / 2 / 1) BEFORE the Schur complement — so that tests can hand the same edges to the reference's compiled function. */
typedef struct OrcMargDump {
  int32_t filled, has_imu, fixed_last, n_vis, cap, pad_;
  double info_imu[81], info_bias[36], info_prior[225];
  double delta_imu, delta_bias, delta_prior;
  double C[225], CL[225], CCL[225];
  int32_t* level; /* [cap] */
  double* delta;  /* [cap] */
} OrcMargDump;
void orc_set_marg_dump(OrcMargDump* d);

/* ---- local / global BA (PR-V-Bias vertices per keyframe, marginalised points) ------------------------------- */
typedef struct OrcBaProblem {
  int32_t n_states;  /* keyframes: local first (ascending id), then fixed */
  int32_t n_points, n_edges, n_imu;
  const OrcNavState* states;
  const uint8_t* state_flags; /* bit0: PR fixed, bit1: has V+Bias vertices, bit2: V/Bias fixed */
  const double* points;       /* [P][3] */
  const int32_t* edge_state;  /* [E] */
  const int32_t* edge_point;  /* [E]  edges sorted by point */
  const float* obs;           /* [E][3] */
  const float* inv_sigma2;    /* [E] */
  const uint8_t* edge_flags;  /* [E] STEREO | CLOSE | LEVEL1 */
  /* inertial factors between consecutive keyframes */
  const int32_t* imu_i;       /* [n_imu] state index of KF0 */
  const int32_t* imu_j;       /* [n_imu] state index of KF1 */
  const OrcImuPreint* preint; /* [n_imu] (dt == 0: only the bias edge) */
  const double* imu_dt_kf;    /* [n_imu] pKF1->ftimestamp_ - pKF0->ftimestamp_ */
  double gw[3];
  double inv_sigma_bg2, inv_sigma_ba2;
  int32_t large;    /* bLarge */
  int32_t rec_init; /* bRecInit */
  int32_t visual_only; /* LocalBundleAdjustment (PR vertices only, src/Optimizer.cc:1876) */
  int32_t global_ba;   /* set by orc_global_ba_prv: bit 0 = GlobalBundleAdjustmentNavStatePRV graph, bit 1 = bRobust */
} OrcBaProblem;

typedef struct OrcBaResult {
  double err0, err_end; /* activeRobustChi2 before / after (src/Optimizer.cc:539, 652) */
  double lambda_final;
  int32_t iterations[2];
  int32_t accepted; /* 0 when the "FAIL LOCAL-INERTIAL BA" guard rejected the result */
  int32_t n_erase;
} OrcBaResult;

/* Optimizer::LocalBundleAdjustmentNavStatePRV from "Setup optimizer" to before the write-back.
 * Outputs: states_out [n_states], points_out [P][3], edge_chi2 [E], erase [E] (1 = vToErase entry). */
int orc_local_ba_prv(const OrcBaProblem* pb, const OrcCamera* cam, OrcNavState* states_out, double* points_out,
                     double* edge_chi2, uint8_t* erase, OrcBaResult* res);

/* Optimizer::GlobalBundleAdjustmentNavStatePRV (src/Optimizer.cc:771-1342; bScaleOpt = false, no IMU initiator) from the
 * vertex set-up to before the write-back.  Returns the LM iterations run. */
int orc_global_ba_prv(const OrcBaProblem* pb, const OrcCamera* cam, int n_iterations, int robust, OrcNavState* states_out,
                      double* points_out, double* edge_chi2, OrcBaResult* res);
/* bScaleOpt = true (System::FinalGBA): VertexScale + EdgeReprojectPRS[Stereo]; points_out = scale * points. */
int orc_global_ba_prv_scale(const OrcBaProblem* pb, const OrcCamera* cam, int n_iterations, int robust,
                            OrcNavState* states_out, double* points_out, double* edge_chi2, OrcBaResult* res,
                            double* scale_out);
/* the IMU initialiser's call: gravity-direction vertex + EdgeNavStatePRVG + one prior-bias edge; gw_io in / out */
int orc_global_ba_prv_init(const OrcBaProblem* pb, const OrcCamera* cam, int n_iterations, double gw_io[3],
                           OrcNavState* states_out, double* points_out, double* edge_chi2, OrcBaResult* res);
int orc_ba_debug_step_gdir(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, const double gw[3], double* x_pose,
                           double* x_points, double* chi2);
int orc_ba_debug_step_scale(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, double scale0, double* x_pose,
                            double* x_points, double* chi2);

/* One damped Gauss-Newton step (build + Schur solve at the given lambda) at the input estimate, for the tests'
 * cross-check against a dense solve of the full normal equations.  Returns the pose dimension np or < 0. */
int orc_ba_debug_step(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, double* x_pose, double* x_points,
                      double* chi2);

/* Reduced camera system of a (partial) problem for the sharding tests: S [np*np], bs [np], b [np]. */
int orc_ba_debug_system(const OrcBaProblem* pb, const OrcCamera* cam, double lambda, int lambda_on_poses, double* S,
                        double* bs, double* b, double* chi2);

/* ---- building blocks exposed for the tests -------------------------------------------------------------------- */
/* EdgeReproject<DE,DV,2> error and Jacobians (g2otypes.h:400-541): e[3], J_pose[3][6] (dp, dphi), J_point[3][3] */
void orc_edge_reproject(const OrcCamera* cam, const OrcNavState* ns, const double Xw[3], const float obs[3], int stereo,
                        double e[3], double* J_pose, double* J_point, double* depth);
/* EdgeNavStateI<NV> (g2otypes.h:725-884): order 0 = PVR (NV=3), 1 = PRV (NV=5).  e[9], Ji[9][9], Jj[9][9], Jb[9][6]
 * with the 9 state columns in the same order as the residual. */
void orc_edge_navstate(const OrcNavState* nsi, const OrcNavState* nsj, const OrcImuPreint* pre, const double gw[3],
                       int order, double e[9], double* Ji, double* Jj, double* Jb);
/* NavState::IncSmall variants: kind 0 = PR(6), 1 = PVR(9), 2 = V(3), 3 = Bias(6) */
void orc_navstate_oplus(OrcNavState* ns, int kind, const double* dx);
/* Edges of the scale / gravity-direction variants of the global BA (restated ahead of their device side):
 * EdgeReprojectPRS[Stereo] with a VertexScale, VertexGThetaXYRwI and EdgeNavStatePRVG.  J_scale [3], JG [9][2]. */
void orc_edge_reproject_scale(const OrcCamera* cam, const OrcNavState* ns, const double Xh[3], double scale_est,
                              const float obs[3], int stereo, double e[3], double* J_pose, double* J_point, double* J_scale);
void orc_gdir_init(const double gw[3], double q_wI[4]);
void orc_gdir_oplus(double q_wI[4], const double d[2]);
void orc_edge_navstate_g(const OrcNavState* nsi, const OrcNavState* nsj, const OrcImuPreint* pre, const double q_wI[4],
                         const double GI[3], double e[9], double* Ji, double* Jj, double* Jb, double* JG);
void orc_edge_prior_pvr(const OrcNavState* ns, const OrcNavState* prior, double e[15], double* Jpvr /*15x9*/);

/* ---- Optimizer::OptimizeSim3 (src/Optimizer.cc:2689-2920) ---------------------------------------------------------
 * One free VertexNavStatePR (S12 as a NavState: mRwb = R12^-1, mpwb = -(mRwb t12), :2717-2722) + VertexScale (fixed iff
 * bFixScale) + FIXED points; per match an EdgeReprojectPRS (x1 = pi(S12 X2c), g2otypes.h MODE 1) and an
 * EdgeReprojectPRSInv (x2 = pi(S12^-1 X1c), MODE 2), Huber sqrt(th2); optimize(5), outlier pairs removed, optimize(5 | 10). */
typedef struct OrcSim3Problem {
  OrcNavState ns;   /* only p, q are read */
  double scale;     /* g2oS12.scale() */
  float th2;        /* chi2 gate; Huber delta = sqrtf(th2) */
  int32_t fix_scale;
  int32_t m_begin, m_end; /* this candidate's range in the shared match arrays */
} OrcSim3Problem;
typedef struct OrcSim3Result {
  OrcNavState ns;   /* optimised (input when the call returned 0 before the second stage) */
  double scale;
  double chi2_final, lambda_final;
  int32_t n_inliers;  /* return value nIn (0: fewer than 10 inlier pairs after the first stage) */
  int32_t n_corr;     /* nCorrespondences */
  int32_t n_bad;      /* pairs removed after the first stage */
  int32_t iterations; /* LM iterations run in total */
} OrcSim3Result;
/* Xc1 / Xc2 [M][3]: P3D1c / P3D2c (float arithmetic of :2771-2783 done by the caller, cast to double); obs1 / obs2 [M][2];
 * inv_sigma2_1 / _2 [M].  keep [M]: 1 = vpMatches1 entry kept.  chi2_12 / chi2_21 [M]: last e->chi2() of the pair. */
int orc_optimize_sim3(const OrcSim3Problem* pb, const OrcCamera* cam, const double* Xc1, const double* Xc2, const float* obs1,
                      const float* obs2, const float* inv_sigma2_1, const float* inv_sigma2_2, OrcSim3Result* res,
                      uint8_t* keep, double* chi2_12, double* chi2_21);
/* the two edges alone: inverse = 0 EdgeReprojectPRS, 1 EdgeReprojectPRSInv.  e[2], J_pose[2][6] (dp, dphi), J_scale[2] */
void orc_edge_sim3(const OrcCamera* cam, const OrcNavState* ns, double scale, const double Xh[3], const float obs[2],
                   int inverse, double e[2], double* J_pose, double* J_scale);
int orc_inverse(const double* A, int n, double* Ainv);
/* ---- g2o's Levenberg-Marquardt control flow over callbacks (lm_oracle.cc) ---------------------------------------------------
 * SparseOptimizer::optimize (sparse_optimizer.cpp:354-419) around OptimizationAlgorithmLevenberg::solve
 * (optimization_algorithm_levenberg.cpp:61-189).  The problem is whatever the callbacks say; the oracle's essential-graph
 * optimisation runs through this driver, and oracle/_ref compiles the REFERENCE's own two functions against the same callbacks
 * (ref_lm_optimize), so the control flow — lambda schedule, accept / reject, the three stop rules — is pinned by the reference's
 * text on identical arithmetic. */
typedef struct OrcLmCallbacks {
  void* ctx;
  int n;                                   /* dimension of the system (solver->vectorSize()) */
  double (*errors)(void* ctx);             /* computeActiveErrors() then activeRobustChi2() */
  void (*build)(void* ctx);                /* solver->buildSystem() */
  int (*solve)(void* ctx, double lambda);  /* setLambda(lambda, true); solve(); restoreDiagonal() -> 1 when the solve succeeded */
  void (*update)(void* ctx);               /* optimizer->update(solver->x()) */
  void (*push)(void* ctx);
  void (*pop)(void* ctx);
  void (*discard_top)(void* ctx);
  const double* (*x)(void* ctx);           /* solver->x() */
  const double* (*b)(void* ctx);           /* solver->b() */
  double (*hessian_diag)(void* ctx, int j); /* H(j, j) of the last buildSystem (computeLambdaInit) */
  int (*terminate)(void* ctx);             /* the force-stop flag; may be NULL */
} OrcLmCallbacks;
/* stats = {chi2 at the first iteration's start, chi2 after the last accepted step, iterations run, final lambda, trials} */
typedef int (*OrcLmDriver)(const OrcLmCallbacks* cb, int iterations, double user_lambda_init, double* stats);
int orc_lm_optimize(const OrcLmCallbacks* cb, int iterations, double user_lambda_init, double* stats);
/* g2o::RobustKernelHuber::robustify with setDelta(delta): rho[0] = rho(e), rho[1] = rho'(e) */
void orc_huber(double delta, double e, double rho[2]);
/* GraphOperator::Chi2LargeSetLevel's per-edge decision: 1 = setLevel(1) */
int orc_chi2_large_level(double chi2, int dim_freedom, float rat_th_chi2);
/* SO3ex helpers (common/so3_extra.h) as restated in so3_oracle.h: op 0 exp -> quaternion, 1 Exp -> R, 2 log(q), 3 Log(R),
 * 4 JacobianR, 5 JacobianRInv, 6 normalizeRotationM; matrices row-major */
void orc_so3(int op, const double* in, double* out);
/* test hook: the LM driver behind every BA driver of ba_oracle.cc (NULL restores orc_lm_optimize) */
void orc_set_lm_driver(OrcLmDriver d);
/* camm::{Pinhole,Radtan,KB8}Camera::Project: float pixel + d(img)/d(p3d) (2x3 row-major, may be NULL) */
void orc_cam_project(const OrcCamera* cam, const double P[3], float uv[2], double* J);

/* ---- Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:2309-2688), posegraph_oracle.cc --------------------------------
 * g2o::Sim3 (types/sim3.h): r as Eigen coefficient order (x, y, z, w), t, s. */
typedef struct OrcSim3 {
  double q[4];
  double t[3];
  double s;
} OrcSim3;
void orc_sim3_exp(const double u[7], OrcSim3* out);
void orc_sim3_log(const OrcSim3* S, double out[7]);
void orc_sim3_mul(const OrcSim3* a, const OrcSim3* b, OrcSim3* out);
void orc_sim3_inv(const OrcSim3* a, OrcSim3* out);
void orc_sim3_from_Rt(const double R[9], const double t[3], double s, OrcSim3* out);
void orc_edge_sim3_graph(const OrcSim3* meas, const OrcSim3* v0, const OrcSim3* v1, int fix0, int fix1, int fix_scale, double e[7],
                         double Ji[49], double Jj[49]);
int orc_essential_graph(int K, const OrcSim3* Scw, const uint8_t* fixed, int fix_scale, int E, const int32_t* ei, const int32_t* ej,
                        const OrcSim3* meas, const double* info, int iterations, double lambda_init, int single_step, OrcSim3* out,
                        double* stats, double* H_out, double* b_out);
/* orc_essential_graph with an explicit LM driver (NULL: orc_lm_optimize); tests pass oracle/_ref's ref_lm_optimize */
int orc_essential_graph_lm(int K, const OrcSim3* Scw, const uint8_t* fixed, int fix_scale, int E, const int32_t* ei, const int32_t* ej,
                           const OrcSim3* meas, const double* info, int iterations, double lambda_init, OrcLmDriver lm, OrcSim3* out,
                           double* stats);
void orc_essential_graph_recover_se3(int K, const OrcSim3* S, double* Tcw);
void orc_essential_graph_correct_points(int n, const float* Pw, const int32_t* ref, const OrcSim3* Scw_before, const OrcSim3* Scw_after,
                                        float* out);
#ifdef __cplusplus
}
#endif

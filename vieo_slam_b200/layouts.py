"""numpy dtypes of the C-ABI structs in include/vieo_b200.h (the oracle uses byte-identical layouts)."""
import numpy as np

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])

PREINT_DTYPE = np.dtype([("Rij", "f8", (3, 3)), ("vij", "f8", 3), ("pij", "f8", 3), ("SigmaPRV", "f8", (9, 9)),
                         ("SigmaPVR", "f8", (9, 9)), ("Jgp", "f8", (3, 3)), ("Jap", "f8", (3, 3)), ("Jgv", "f8", (3, 3)),
                         ("Jav", "f8", (3, 3)), ("JgR", "f8", (3, 3)), ("dt", "f8"), ("status", "i4"), ("pad_", "i4")])

# NavState (src/Odom/NavState.h:17-36): q = (w, x, y, z)
NAVSTATE_DTYPE = np.dtype([("p", "f8", 3), ("q", "f8", 4), ("v", "f8", 3), ("bg", "f8", 3), ("ba", "f8", 3),
                           ("dbg", "f8", 3), ("dba", "f8", 3)])

# model: 0 pinhole, 1 radtan (dist = k1..k_numk, p1, p2), 2 KB8 (dist = k1..k4)
CAMERA_DTYPE = np.dtype([("fx", "f4"), ("fy", "f4"), ("cx", "f4"), ("cy", "f4"), ("bf", "f4"), ("model", "i4"),
                         ("num_k", "i4"), ("pad_", "f4"), ("dist", "f4", 8), ("Rcb", "f8", (3, 3)), ("tcb", "f8", 3)])
CAM_PINHOLE, CAM_RADTAN, CAM_KB8 = 0, 1, 2

EDGE_STEREO, EDGE_CLOSE, EDGE_LEVEL1, EDGE_NOKERNEL = 1, 2, 4, 8

POSEOPT_PROBLEM_DTYPE = np.dtype([("cur", NAVSTATE_DTYPE), ("last", NAVSTATE_DTYPE), ("prior", NAVSTATE_DTYPE),
                                  ("preint", PREINT_DTYPE), ("prior_info", "f8", (15, 15)), ("gw", "f8", 3),
                                  ("inv_sigma_bg2", "f8"), ("inv_sigma_ba2", "f8"), ("dt_frames", "f8"),
                                  ("mode", "i4"), ("last_has_prior", "i4"), ("compute_marg", "i4"), ("no_mps", "i4"),
                                  ("edge_begin", "i4"), ("edge_end", "i4")])

POSEOPT_RESULT_DTYPE = np.dtype([("cur", NAVSTATE_DTYPE), ("last", NAVSTATE_DTYPE), ("marg_cov_inv", "f8", (15, 15)),
                                 ("chi2_final", "f8"), ("lambda_final", "f8"), ("n_inliers", "i4"), ("n_initial", "i4"),
                                 ("iterations", "i4"), ("prior_set", "i4")])

# VieoSim3Problem / VieoSim3Result (Optimizer::OptimizeSim3); the oracle's OrcSim3* are byte-identical
SIM3_PROBLEM_DTYPE = np.dtype([("ns", NAVSTATE_DTYPE), ("scale", "f8"), ("th2", "f4"), ("fix_scale", "i4"), ("m_begin", "i4"),
                               ("m_end", "i4")])
SIM3_RESULT_DTYPE = np.dtype([("ns", NAVSTATE_DTYPE), ("scale", "f8"), ("chi2_final", "f8"), ("lambda_final", "f8"),
                              ("n_inliers", "i4"), ("n_corr", "i4"), ("n_bad", "i4"), ("iterations", "i4")])
assert SIM3_PROBLEM_DTYPE.itemsize == 200 and SIM3_RESULT_DTYPE.itemsize == 216

BA_RESULT_DTYPE = np.dtype([("err0", "f8"), ("err_end", "f8"), ("lambda_final", "f8"), ("iterations", "i4", 2),
                            ("accepted", "i4"), ("n_erase", "i4")])

# VieoSbpFrame (include/vieo_b200.h): one current frame of a guided-search batch
SBP_FRAME_DTYPE = np.dtype([("kp_begin", "i4"), ("n_kp", "i4"), ("q_begin", "i4"), ("n_q", "i4"), ("minx", "f4"), ("maxx", "f4"),
                            ("miny", "f4"), ("maxy", "f4"), ("grid_winv", "f4"), ("grid_hinv", "f4"), ("bf", "f4"), ("b", "f4"),
                            ("fx", "f4"), ("fy", "f4"), ("cx", "f4"), ("cy", "f4"), ("th", "f4"), ("th_far", "f4"),
                            ("nn_ratio", "f4"), ("mono", "i4"), ("check_orientation", "i4"), ("n_levels", "i4"),
                            ("scale", "f4", 16), ("qcw", "f8", 4), ("tcw", "f8", 3), ("qlw", "f8", 4), ("tlw", "f8", 3)])
SBP_LAST_FRAME, SBP_LOCAL_MAP, SBP_RELOC = 0, 1, 2
# VieoSbpReloc: per-frame record of the relocalisation search (ORBmatcher::SearchByProjection(Frame&, KeyFrame*, ...))
SBP_RELOC_DTYPE = np.dtype([("orb_dist", "i4"), ("log_scale_factor", "f4"), ("level_ratio", "f4", 16)])
assert SBP_RELOC_DTYPE.itemsize == 72

# VieoFrustumFrame (include/vieo_b200.h): one current frame of a Frame::isInFrustum batch; the oracle's OrcFrustumFrame is
# the same record without the trailing level_ratio table
_FRUSTUM_FIELDS = [("q_begin", "i4"), ("n_q", "i4"), ("Rcw", "f4", 9), ("tcw", "f4", 3), ("Ow", "f4", 3), ("fx", "f4"),
                   ("fy", "f4"), ("cx", "f4"), ("cy", "f4"), ("minx", "f4"), ("maxx", "f4"), ("miny", "f4"), ("maxy", "f4"),
                   ("bf", "f4"), ("cos_limit", "f4"), ("log_scale_factor", "f4"), ("n_levels", "i4")]
FRUSTUM_FRAME_DTYPE = np.dtype(_FRUSTUM_FIELDS + [("level_ratio", "f4", 16)])
ORC_FRUSTUM_FRAME_DTYPE = np.dtype(_FRUSTUM_FIELDS)
assert FRUSTUM_FRAME_DTYPE.itemsize == 180 and ORC_FRUSTUM_FRAME_DTYPE.itemsize == 116

# VieoProjSearchFrame (include/vieo_b200.h) == OrcProjSearchFrame: one keyframe of a SearchByProjectionBase batch
PROJ_SEARCH_FRAME_DTYPE = np.dtype([("kp_begin", "i4"), ("n_kp", "i4"), ("q_begin", "i4"), ("n_q", "i4"), ("Rcw", "f4", 9),
                                    ("tcw", "f4", 3), ("Ow", "f4", 3), ("fx", "f4"), ("fy", "f4"), ("cx", "f4"), ("cy", "f4"),
                                    ("minx", "f4"), ("maxx", "f4"), ("miny", "f4"), ("maxy", "f4"), ("grid_winv", "f4"),
                                    ("grid_hinv", "f4"), ("bf", "f4"), ("use_bf", "i4"), ("check_viewing_angle", "i4"),
                                    ("th_radius", "f4"), ("n_levels", "i4"), ("log_scale_factor", "f4"), ("scale", "f4", 16),
                                    ("inv_level_sigma2", "f4", 16), ("level_ratio", "f4", 16)])
assert PROJ_SEARCH_FRAME_DTYPE.itemsize == 332

assert SBP_FRAME_DTYPE.itemsize == 88 + 64 + 112
assert NAVSTATE_DTYPE.itemsize == 22 * 8 and CAMERA_DTYPE.itemsize == 64 + 96

# VieoSftPair (ORBmatcher::SearchForTriangulation, one keyframe pair)
SFT_PAIR_DTYPE = np.dtype([("kp1_begin", "i4"), ("n_kp1", "i4"), ("kp2_begin", "i4"), ("n_kp2", "i4"), ("node1_begin", "i4"),
                           ("n_nodes1", "i4"), ("node2_begin", "i4"), ("n_nodes2", "i4"), ("ptr1_begin", "i4"), ("ptr2_begin", "i4"),
                           ("idx1_begin", "i4"), ("idx2_begin", "i4"), ("out_begin", "i4"), ("nscr_begin", "i4"),
                           ("only_stereo", "i4"), ("check_orientation", "i4"), ("ex", "f4"), ("ey", "f4"),
                           ("scale_factor2", "f4", 16), ("level_sigma2_2", "f4", 16), ("F12", "f8", 9)])
assert SFT_PAIR_DTYPE.itemsize == 72 + 128 + 72

# VieoBowPair (ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), one (keyframe, frame) pair)
BOW_PAIR_DTYPE = np.dtype([("kp1_begin", "i4"), ("n_kp1", "i4"), ("kp2_begin", "i4"), ("n_kp2", "i4"), ("node1_begin", "i4"),
                           ("n_nodes1", "i4"), ("node2_begin", "i4"), ("n_nodes2", "i4"), ("ptr1_begin", "i4"), ("ptr2_begin", "i4"),
                           ("idx1_begin", "i4"), ("idx2_begin", "i4"), ("out_begin", "i4"), ("check_orientation", "i4"),
                           ("nn_ratio", "f4"), ("pad_", "i4")])
assert BOW_PAIR_DTYPE.itemsize == 64

# VieoSim3 (g2o::Sim3, optimizer/g2o/g2o/types/sim3.h): r in Eigen coefficient order (x, y, z, w), t, s; OrcSim3 is identical
SIM3_DTYPE = np.dtype([("q", "f8", 4), ("t", "f8", 3), ("s", "f8")])
# VieoPoseGraphStats (Optimizer::OptimizeEssentialGraph)
POSEGRAPH_STATS_DTYPE = np.dtype([("chi2_initial", "f8"), ("chi2_final", "f8"), ("lambda_final", "f8"), ("iterations", "i4"),
                                  ("trials", "i4"), ("n_free", "i4"), ("ok", "i4")])
assert SIM3_DTYPE.itemsize == 64 and POSEGRAPH_STATS_DTYPE.itemsize == 40

# VieoFrustumCam / VieoFrustumRigFrame (Frame::isInFrustum with a camera rig); OrcFrustumCam / OrcFrustumRigFrame are identical
FRUSTUM_CAM_DTYPE = np.dtype([("q_cr", "f4", 4), ("t_cr", "f4", 3), ("t_rc", "f4", 3), ("fx", "f4"), ("fy", "f4"), ("cx", "f4"),
                              ("cy", "f4"), ("k", "f4", 4), ("minx", "f4"), ("maxx", "f4"), ("miny", "f4"), ("maxy", "f4"),
                              ("model", "i4"), ("pad_", "i4")])
FRUSTUM_RIG_FRAME_DTYPE = np.dtype([("q_begin", "i4"), ("n_q", "i4"), ("Rcw", "f4", 9), ("tcw", "f4", 3), ("Ow", "f4", 3), ("bf", "f4"),
                                    ("cos_limit", "f4"), ("log_scale_factor", "f4"), ("n_levels", "i4"), ("n_cams", "i4"),
                                    ("level_ratio", "f4", 16), ("cam", FRUSTUM_CAM_DTYPE, 4)])
assert FRUSTUM_CAM_DTYPE.itemsize == 96 and FRUSTUM_RIG_FRAME_DTYPE.itemsize == 8 + 60 + 20 + 64 + 4 * 96

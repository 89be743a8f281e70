#!/usr/bin/env python3
"""Regression fixtures of the ORACLE itself (not an independent reference): outputs of the CPU restatements on seeded
synthetic inputs, committed so that a later edit of oracle/*.cc that changes a result is caught by the CPU suite
(tests/test_oracle_goldens.py).  Integer outputs are compared exactly, floating-point ones to 1e-9 relative.
  python tests/golden/gen_oracle_goldens.py      (rewrites tests/golden/oracle_goldens.npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from vieo_slam_b200 import synth  # noqa: E402


def compute():
    out = {}
    # guided searches / visibility / fuse search
    pb = synth.make_sbp_problem(77, n_frames=2, mode=synth.SBP_LAST_FRAME, th=15.0)
    for k, v in zip(("kp_match", "q_match", "q_dist", "n"), O.search_by_projection(pb)):
        out["sbp_last_" + k] = v
    pb = synth.make_frustum_problem(43, n_frames=2, n_q=1200, blocked_frac=0.2)
    pb["frames"]["nn_ratio"] = 0.8
    fo, kpm, qm, qd, nm = O.search_local_points(pb)
    for k in ("inview", "level", "proj", "viewcos", "depth", "n_inview"):
        out["slp_" + k] = fo[k]
    out["slp_kp_match"], out["slp_q_match"], out["slp_q_dist"], out["slp_n"] = kpm, qm, qd, nm
    pb = synth.make_fuse_problem(81, n_frames=2, n_q=1200)
    for k, v in zip(("best", "dist", "level"), O.proj_search(pb)):
        out["fuse_" + k] = v
    d = synth.make_distinctive_problem(51, n_points=400, max_obs=30, long_lists=(70, 300))
    out["dd_best"], out["dd_median"] = O.distinctive_descriptors(d["pool"], d["ptr"], d["rows"])
    # IMU: pre-integration, initial gyro bias
    g = synth.make_gyro_bias_problem(61, n_kf=20, kf_gap=(1, 12))
    nz = O.imu_noise()
    pre = O.imu_preintegrate_frames(g["seq"], g["kf_idx"], nz)
    out["imu_Rij"], out["imu_SigmaPRV"], out["imu_dt"] = pre["Rij"], pre["SigmaPRV"], pre["dt"]
    out["gyro_neq"], out["gyro_dbg"] = O.gyro_bias_init(pre, g["Rwb"], True)
    # BA: pose optimisation, local BA, global BA (+ scale / gravity variants)
    seq = synth.vio_sequence(5, 60)
    cam = synth.euroc_camera()
    pre_all = O.imu_preintegrate_frames(seq, list(range(60)), nz)
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, pre_all, cam, n_points=300, seed=1)
    res, outl, chi2 = O.pose_optimization(pbs[:4], cam, X, obs, w, fl)
    out["po_p"], out["po_inliers"], out["po_outlier"], out["po_marg"] = res["cur"]["p"], res["n_inliers"], outl[:1200], res["marg_cov_inv"]
    kf = list(range(0, 60, 3))
    pre_kf = O.imu_preintegrate_frames(seq, kf, nz)
    dl = synth.make_lba_problem(seq, pre_kf, kf, cam, n_local=8, n_fixed=6, n_points=300, seed=4)
    lba = O.local_ba_prv(dl, cam)
    out["lba_p"], out["lba_points"], out["lba_erase"], out["lba_err"] = lba["states"]["p"], lba["points"], lba["erase"], np.array(
        [lba["res"]["err0"], lba["res"]["err_end"]])
    gb = synth.make_gba_problem(seq, pre_kf, kf, cam, n_points=400, seed=3)
    gba = O.global_ba_prv(gb, cam, n_iterations=6, robust=False)
    out["gba_p"], out["gba_err"] = gba["states"]["p"], np.array([gba["res"]["err0"], gba["res"]["err_end"]])
    gs = O.global_ba_prv_scale(gb, cam, n_iterations=6, robust=False)
    out["gbas_p"], out["gbas_scale"], out["gbas_err"] = gs["states"]["p"], np.array([gs["scale"]]), np.array([gs["res"]["err_end"]])
    flags = gb["state_flags"].copy(); flags[0] = 3
    gi = dict(gb); gi["state_flags"] = flags
    tilt = synth.so3_exp(np.array([0.03, -0.02, 0.0])) @ synth.GRAVITY_W
    go = O.global_ba_prv_init(gi, cam, 8, tilt)
    out["gbai_gw"], out["gbai_err"] = go["gw"], np.array([go["res"]["err_end"]])
    # loop closing: Sim3 refinement, essential graph (single damped step: no accept / reject decisions on numeric-Jacobian noise),
    # point correction; the rig form of the visibility test
    eg = synth.make_essential_graph(K=30, seed=5, fix_scale=False, odom_info_every=4, n_points=200)
    so, st, H, b = O.essential_graph(eg, lambda_init=1e-3, single_step=True, want_system=True)
    out["eg_t"], out["eg_s"], out["eg_chi"], out["eg_Hdiag"], out["eg_b"] = so["t"], so["s"], np.array([st["chi2_initial"], st["chi2_final"]]), np.diag(H).copy(), b
    full, stf = O.essential_graph(eg)
    out["eg_full_chi"] = np.array([stf["chi2_final"]])
    out["eg_points"] = O.essential_graph_correct_points(eg["Pw"], eg["ref"], eg["Scw"], so)
    rp = synth.make_frustum_rig_problem(19, n_frames=2, n_q=800, n_cams=4, model=2)
    ro = O.is_in_frustum_rig(rp)
    for k in ("inview", "cam_mask", "level", "proj", "viewcos", "depth", "n_inview"):
        out["rig_" + k] = ro[k]
    return {k: np.asarray(v) for k, v in out.items()}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_goldens.npz"), **compute())
    print("written", os.path.getsize(os.path.join(HERE, "oracle_goldens.npz")), "bytes")

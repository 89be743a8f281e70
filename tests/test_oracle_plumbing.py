"""CPU: BASELINE.json configs[0] — "EuRoC V1_01 stereo (no IMU), 1200 feats, CPU reference run (plumbing)" — as SURVEY.md
8(d) restates it: the CPU oracle alone, ORBextractor on both views -> Frame::ComputeStereoMatches -> back-projection of
the stereo keypoints (Frame::UnprojectStereo) -> SearchByProjection against the next frame -> visual PoseOptimization,
chained end to end on a synthetic rectified stereo stream.  The scene is a fronto-parallel textured plane panned by whole
pixels, so the camera motion between two frames is a known pure translation."""
import numpy as np

import oracle_lib as O
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import EDGE_STEREO, KP_DTYPE, POSEOPT_PROBLEM_DTYPE, SBP_FRAME_DTYPE, SBP_LAST_FRAME
from vieo_slam_b200.synth import EUROC, stereo_stream


def _frame(img_l, img_r, bf, minz):
    oL, oR = O.OrbOracle(1200, 1.2, 8, 20, 7), O.OrbOracle(1200, 1.2, 8, 20, 7)
    nl, kl, dl, _ = oL.extract(img_l)
    nr, kr, dr, _ = oR.extract(img_r)
    ur, depth, sad, kept = O.stereo_matches(oL, kl, dl, oR, kr, dr, bf, minz)
    return dict(kps=kl, desc=dl, ur=ur, depth=depth, kept=kept, scale=oL.tables()["scale"])


def test_stereo_plumbing_extract_match_track_optimise():
    fx, fy, cx, cy, bf = (np.float32(EUROC[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    minz = np.float32(bf / fx)
    imgs = stereo_stream(2, 101).reshape(2, 2, 480, 752)
    f0 = _frame(imgs[0, 0], imgs[0, 1], bf, minz)
    f1 = _frame(imgs[1, 0], imgs[1, 1], bf, minz)
    assert f0["kept"] > 400 and f1["kept"] > 400
    # the plane is fronto-parallel: every stereo keypoint of a frame has the same disparity (sub-pixel fit included)
    d0 = (f0["kps"]["x"] - f0["ur"])[f0["ur"] >= 0]
    # (to the sub-pixel accuracy of the SAD refinement, which works on the keypoint's pyramid level)
    assert np.percentile(np.abs(d0 - np.median(d0)), 90) < 1.0 and np.abs(d0 - np.median(d0)).max() < 5.0
    # Frame::UnprojectStereo with the body frame = camera frame and frame 0 at the origin
    cam = synth.euroc_camera().copy()
    cam["Rcb"] = np.eye(3); cam["tcb"] = 0
    ok = f0["depth"] > 0
    z = f0["depth"][ok].astype(np.float64)
    X = np.stack([(f0["kps"]["x"][ok] - cx) * z / fx, (f0["kps"]["y"][ok] - cy) * z / fy, z], 1).astype(np.float32).astype(np.float64)
    # TrackWithMotionModel: SearchByProjection(CurrentFrame = frame 1, LastFrame = frame 0) with the identity as the guess
    F = np.zeros(1, SBP_FRAME_DTYPE)
    n1 = len(f1["kps"])
    F["n_kp"], F["n_q"] = n1, len(X)
    F["maxx"], F["maxy"] = 752, 480
    F["grid_winv"], F["grid_hinv"] = np.float32(64) / np.float32(752), np.float32(48) / np.float32(480)
    F["bf"], F["b"] = bf, bf / fx
    F["fx"], F["fy"], F["cx"], F["cy"] = fx, fy, cx, cy
    F["th"], F["nn_ratio"], F["mono"], F["check_orientation"], F["n_levels"] = 40.0, 0.9, 0, 1, 8
    F["scale"][0, :8] = f0["scale"]
    F["qcw"] = F["qlw"] = [1, 0, 0, 0]
    pb = dict(frames=F, mode=SBP_LAST_FRAME, kps=np.ascontiguousarray(f1["kps"], KP_DTYPE), uright=f1["ur"], desc=f1["desc"],
              kp_blocked=np.zeros(n1, np.uint8), q_Xw=X, q_level=f0["kps"]["octave"][ok].astype(np.int32),
              q_angle=f0["kps"]["angle"][ok].astype(np.float32), q_desc=np.ascontiguousarray(f0["desc"][ok]),
              q_flags=np.ones(len(X), np.uint8))
    kp_match, q_match, q_dist, nm = O.search_by_projection(pb)
    assert nm[0] > 150, nm
    # visual PoseOptimization of frame 1 on the matches (stereo edges where frame 1 has a right coordinate)
    k1 = np.nonzero(kp_match >= 0)[0]
    q = kp_match[k1]
    obs = np.stack([f1["kps"]["x"][k1], f1["kps"]["y"][k1], f1["ur"][k1]], 1).astype(np.float32)
    flags = np.where(f1["ur"][k1] >= 0, EDGE_STEREO, 0).astype(np.uint8)
    inv_s2 = (1.0 / (f0["scale"][f1["kps"]["octave"][k1]] ** 2)).astype(np.float32)
    pbs = np.zeros(1, POSEOPT_PROBLEM_DTYPE)
    pbs["cur"]["q"] = pbs["last"]["q"] = pbs["prior"]["q"] = [1, 0, 0, 0]
    pbs["mode"], pbs["edge_begin"], pbs["edge_end"] = 0, 0, len(k1)
    res, outl, chi2 = O.pose_optimization(pbs, cam, X[q], obs, inv_s2, flags)
    assert res[0]["n_inliers"] > 0.7 * len(k1)
    # the plane moved by whole pixels (ox, oy) between the frames: t = -shift * z / f, rotation stays the identity
    inl = outl[:len(k1)] == 0
    du = np.median(obs[inl, 0] - (fx * X[q][inl, 0] / X[q][inl, 2] + cx))
    dv = np.median(obs[inl, 1] - (fy * X[q][inl, 1] / X[q][inl, 2] + cy))
    zbar = np.median(X[q][inl, 2])
    p = res[0]["cur"]["p"]           # camera centre in the frame-0 world (Twc); the scene shifts by -p
    assert abs(-p[0] * fx / zbar - du) < 0.5 and abs(-p[1] * fy / zbar - dv) < 0.5, (p, du, dv, zbar)
    R = synth.R_from_quat(res[0]["cur"]["q"])
    assert np.abs(R - np.eye(3)).max() < 5e-3

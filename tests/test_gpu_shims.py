"""GPU: the C++ boundary.  tools/host_shim_gpu_test.cc drives the reference-facing classes of
vieo_slam_b200/host/vieo_shims.hpp (ORBextractor::operator(), IMUPreintegrator::PreIntegration, LocalBA::Run / Begin / End)
and the flatten / write-back templates of vieo_flatten.hpp (PoseOptimizationVisual on a stand-in Frame with the reference's
member names) with real data; its outputs are compared BYTE FOR BYTE with the ctypes path on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = tmp_path / "host_shim_gpu_test"
    libdir = os.path.join(ROOT, "vieo_slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tools", "host_shim_gpu_test.cc"), "-o",
                           str(exe), "-L", libdir, "-lvieo_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_flatten_templates_on_cpu(tmp_path):
    """FlattenLocalWindow / WriteBackLocalWindow bookkeeping (keyframe order, edge order by point, fixed keyframes that carry
    V / Bias, ErasePairObs, SetNavState, SetWorldPos + UpdateNormalAndDepth) — no device needed."""
    out = subprocess.run([str(_build(tmp_path)), "--templates-only"], capture_output=True, text=True)
    assert out.returncode == 0 and "HOST_TEMPLATES_OK" in out.stdout, out.stderr


@pytest.mark.gpu
def test_cpp_shims_equal_ctypes_on_gpu(tmp_path):
    import vieo_slam_b200.api as api
    d = tmp_path / "io"
    d.mkdir()
    w = lambda name, a: np.ascontiguousarray(a).tofile(str(d / name))
    # 1. extractor
    img = synth.texture(480, 752, 17)
    w("orb_meta.i32", np.array([752, 480, 1200, 8, 100, 400], np.int32))
    w("orb_img.u8", img)
    # 2. IMU
    seq = synth.vio_sequence(9, 6)
    t = seq["times"]
    imu = seq["imu"]
    lo = max(np.searchsorted(imu[:, 0], t[1], "right") - 1, 0); hi = min(np.searchsorted(imu[:, 0], t[2], "left") + 1, len(imu))
    smp = np.ascontiguousarray(imu[lo:hi])
    bg, ba = seq["truth"][1]["bg"], seq["truth"][1]["ba"]
    s2 = np.array([s * s for s in O.EUROC_IMU_SIGMA])
    w("imu_samples.f64", smp)
    w("imu_par.f64", np.r_[t[1], t[2], bg, ba, s2])
    # 3. visual PoseOptimization
    seq2 = synth.vio_sequence(31, 4, speed=1.0, rot=0.5)
    pre = O.imu_preintegrate_frames(seq2, list(range(4)), O.imu_noise())
    cam = synth.euroc_camera()
    pbs, Xw, obs, ww, fl = synth.make_pose_problems(seq2, pre, cam, n_points=300, seed=3, mode=0)[:5]
    e0, e1 = int(pbs["edge_begin"][1]), int(pbs["edge_end"][1])
    obs = obs.copy()
    obs[(fl & 1) == 0, 2] = 0  # a monocular keypoint has no right coordinate (vuright_ < 0)
    w("po_cam.bin", np.asarray(cam).reshape(1)); w("po_state.bin", pbs["cur"][1:2])
    w("po_Xw.f64", Xw[e0:e1]); w("po_obs.f32", obs[e0:e1]); w("po_w.f32", ww[e0:e1]); w("po_flags.u8", fl[e0:e1])
    # 4. LocalBA window
    seq3 = synth.vio_sequence(77, 100, speed=1.5, rot=1.0)
    kf = list(range(0, 100, 4))
    pre3 = O.imu_preintegrate_frames(seq3, kf, O.imu_noise())
    lba = synth.make_lba_problem(seq3, pre3, kf, cam, n_local=8, n_fixed=10, n_points=600, seed=5)
    w("ba_cam.bin", np.asarray(cam).reshape(1)); w("ba_states.bin", lba["states"]); w("ba_state_flags.u8", lba["state_flags"])
    w("ba_points.f64", lba["points"]); w("ba_edge_state.i32", lba["edge_state"]); w("ba_edge_point.i32", lba["edge_point"])
    w("ba_obs.f32", lba["obs"]); w("ba_w.f32", lba["inv_sigma2"]); w("ba_edge_flags.u8", lba["edge_flags"])
    w("ba_imu_i.i32", lba["imu_i"]); w("ba_imu_j.i32", lba["imu_j"]); w("ba_preint.bin", lba["preint"]); w("ba_dt.f64", lba["imu_dt_kf"])
    w("ba_par.f64", np.r_[lba["gw"], lba["inv_sigma_bg2"], lba["inv_sigma_ba2"]])
    # 5. OptimizeSim3
    from vieo_slam_b200.layouts import SIM3_PROBLEM_DTYPE
    cam3 = synth.euroc_camera()
    cam3["Rcb"] = np.eye(3); cam3["tcb"] = 0
    s3 = synth.make_sim3_problems(cam3, n_candidates=1, n_matches=90, seed=13, fix_scale=True)
    spb, sX1, sX2, so1, so2, sw1, sw2, _ = s3
    inv_s2, _scl = synth.inv_level_sigma2()
    oct_of = lambda wv: np.array([int(np.argmin(np.abs(inv_s2 - v))) for v in wv], np.int32)
    q_ns = spb["ns"]["q"][0]; p_ns = spb["ns"]["p"][0]
    q12 = np.array([q_ns[0], -q_ns[1], -q_ns[2], -q_ns[3]])

    def rot(q):  # the template's quaternion -> matrix formula, same operation order
        w_, x, y, z = (float(v) for v in q)
        return [1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y), 2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z),
                2 * (y * z - w_ * x), 2 * (x * z - w_ * y), 2 * (y * z + w_ * x), 1 - 2 * (x * x + y * y)]
    R12 = rot(q12)
    t12 = np.array([-(R12[3 * r] * float(p_ns[0]) + R12[3 * r + 1] * float(p_ns[1]) + R12[3 * r + 2] * float(p_ns[2])) for r in range(3)])
    w("s3_cam.bin", np.asarray(cam3).reshape(1)); w("s3_X1.f64", sX1); w("s3_X2.f64", sX2); w("s3_obs1.f32", so1); w("s3_obs2.f32", so2)
    w("s3_oct1.i32", oct_of(sw1)); w("s3_oct2.i32", oct_of(sw2)); w("s3_invsigma2.f32", inv_s2)
    w("s3_par.f64", np.r_[q12, t12, float(spb["scale"][0]), 10.0, 1.0])
    out = subprocess.run([str(_build(tmp_path)), str(d)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "HOST_SHIM_GPU_OK" in out.stdout, out.stdout + out.stderr
    r = lambda name, dt: np.fromfile(str(d / name), dt)
    # 1
    orb = api.ORBextractor(1200, 1.2, 8, 20, 7, 752, 480)
    mono, kps, desc = orb(img, pvLappingArea=[100, 400], want_pyramid=True)
    info = r("orb_info.out", np.int32)
    assert info[0] == mono and info[1] == len(kps) and info[2] == 8
    assert r("orb_kps.out", np.uint8).tobytes() == kps.tobytes()
    assert r("orb_desc.out", np.uint8).tobytes() == desc.tobytes()
    assert r("orb_level3.out", np.uint8).tobytes() == orb.mvImagePyramid[3].tobytes()
    # 2
    pg = api.IMUPreintegrator()
    p = pg.preintegrate_batch(smp, np.array([0, len(smp)], np.int32), np.array([[t[1], t[2]]]), np.r_[bg, ba][None])[0]
    got = r("imu.out", np.float64)
    want = np.r_[p["Rij"].ravel(), p["vij"], p["pij"], p["SigmaPRV"].ravel(), p["SigmaPVR"].ravel(), p["Jgp"].ravel(), p["Jap"].ravel(),
                 p["Jgv"].ravel(), p["Jav"].ravel(), p["JgR"].ravel(), p["dt"], float(p["status"])]
    assert got.tobytes() == want.tobytes()
    # 3
    one = np.zeros(1, api.POSEOPT_PROBLEM_DTYPE)  # what PoseOptimizationVisual fills: the frame's state and the edges
    one["cur"] = pbs["cur"][1]; one["edge_begin"] = 0; one["edge_end"] = e1 - e0; one["mode"] = 0
    res, outl, _ = api.Optimizer.PoseOptimizationBatch(one, cam, Xw[e0:e1], obs[e0:e1], ww[e0:e1], fl[e0:e1])
    assert r("po_inliers.out", np.int32)[0] == res["n_inliers"][0]
    assert r("po_outlier.out", np.uint8).tobytes() == outl.tobytes()
    assert r("po_state.out", np.uint8).tobytes() == res["cur"][0].tobytes()
    # 4
    ba = api.BundleAdjuster(max_states=64, max_points=4096, max_edges=32768, max_imu=32)
    ref = ba.LocalBundleAdjustmentNavStatePRV(lba, cam)
    assert r("ba_states.out", np.uint8).tobytes() == ref["states"].tobytes()
    assert r("ba_points.out", np.uint8).tobytes() == ref["points"].tobytes()
    assert r("ba_erase.out", np.uint8).tobytes() == ref["erase"].tobytes()
    assert r("ba_res.out", np.uint8).tobytes() == ref["res"].tobytes()
    # 5: the template rebuilds the vertex from (q12, t12) with the same formulas
    Rq = rot(q_ns)
    one3 = np.zeros(1, SIM3_PROBLEM_DTYPE)
    one3["ns"]["q"] = q_ns
    one3["ns"]["p"] = [-(Rq[3 * r_] * float(t12[0]) + Rq[3 * r_ + 1] * float(t12[1]) + Rq[3 * r_ + 2] * float(t12[2])) for r_ in range(3)]
    one3["scale"] = spb["scale"][0]; one3["th2"] = 10.0; one3["fix_scale"] = 1; one3["m_end"] = 90
    r3, keep3, _, _ = api.Optimizer.OptimizeSim3Batch(one3, cam3, sX1.astype(np.float32).astype(np.float64),
                                                      sX2.astype(np.float32).astype(np.float64), so1, so2, sw1, sw2)
    got3 = r("s3.out", np.float64)
    assert int(got3[0]) == int(r3["n_inliers"][0]) > 50
    assert r("s3_keep.out", np.uint8).tobytes() == keep3.tobytes()
    qo = r3["ns"]["q"][0]; po = r3["ns"]["p"][0]
    q12o = np.array([qo[0], -qo[1], -qo[2], -qo[3]]); Ro = rot(q12o)
    t12o = [-(Ro[3 * r_] * float(po[0]) + Ro[3 * r_ + 1] * float(po[1]) + Ro[3 * r_ + 2] * float(po[2])) for r_ in range(3)]
    assert got3[1:5].tobytes() == q12o.tobytes() and got3[5:8].tobytes() == np.array(t12o).tobytes() and got3[8] == r3["scale"][0]
    # 6: the essential graph collected by the template, optimised and written back in C++, against the ctypes path on the same arrays
    from vieo_slam_b200.layouts import SIM3_DTYPE, POSEGRAPH_STATS_DTYPE
    pg = dict(Scw=r("pg_Scw.bin", SIM3_DTYPE), fixed=r("pg_fixed.u8", np.uint8), fix_scale=1, ei=r("pg_ei.i32", np.int32),
              ej=r("pg_ej.i32", np.int32), meas=r("pg_meas.bin", SIM3_DTYPE), info=r("pg_info.f64", np.float64).reshape(-1, 49))
    assert len(pg["Scw"]) == 24 and len(pg["ei"]) > 40
    out6, T6, st6 = api.Optimizer.OptimizeEssentialGraph(pg)
    assert r("pg_Scw.out", np.uint8).tobytes() == out6.tobytes()
    assert r("pg_stats.out", np.uint8).tobytes() == st6.tobytes() and st6["chi2_final"] < st6["chi2_initial"]
    Tc = r("pg_Tcw.out", np.float64).reshape(24, 3, 4)
    keep = np.arange(24) != 9  # the bad keyframe's pose is not written
    assert Tc[keep].tobytes() == T6[keep].tobytes()
    P6 = api.Optimizer.essential_graph_correct_points(r("pg_Pw.f32", np.float32), r("pg_ref.i32", np.int32), pg["Scw"], out6)
    assert r("pg_Pw.out", np.uint8).tobytes() == P6.tobytes()

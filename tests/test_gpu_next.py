"""GPU parity for the SURVEY.md 8(f) rows, through the C ABI, against the oracle:
 * Frame::isInFrustum + MapPoint::PredictScale (k_frustum): bit-exact floats and levels;
 * Tracking::SearchLocalPoints = visibility test + local-map SearchByProjection on one stream: bit-exact matches;
 * MapPoint::ComputeDistinctiveDescriptors (k_distinctive): bit-exact index and median;
 * Optimizer::OptimizeInitialGyroBias + re-integration (k_gyro_bias -> k_imu_preint): 1e-10 relative on the bias
   (fp64; device sin/cos/atan differ from glibc's by an ulp), 1e-12 on the re-integrated states given the same bias."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth

pytestmark = pytest.mark.gpu

FR_KEYS = ("inview", "level", "proj", "viewcos", "depth", "n_inview")


def _same_frustum(got, ref):
    """Bit-exact, except that a NaN only has to be a NaN on both sides (x86 and the GPU produce different payloads)."""
    for k in FR_KEYS:
        a, b = np.asarray(got[k]).ravel(), np.asarray(ref[k]).ravel()
        if a.dtype.kind == "f":
            na, nb = np.isnan(a), np.isnan(b)
            assert np.array_equal(na, nb), (k, "NaN pattern")
            a, b = a[~na], b[~nb]
        assert a.tobytes() == b.tobytes(), (k, np.nonzero(a != b)[0][:10])


@pytest.mark.parametrize("kw", [dict(n_frames=4, n_q=1500), dict(n_frames=1, n_q=20000, skip_frac=0.3),
                                dict(n_frames=9, n_q=37, skip_frac=0.0)])
def test_is_in_frustum_matches_oracle(kw):
    import vieo_slam_b200.api as api
    pb = synth.make_frustum_problem(41, **kw)
    got = api.isInFrustum(pb)
    ref = O.is_in_frustum(pb)
    _same_frustum(got, ref)
    assert got["n_inview"].sum() > 0.15 * len(pb["p_max_dist"])


def test_is_in_frustum_edge_cases():
    import vieo_slam_b200.api as api
    pb = synth.make_frustum_problem(42, n_frames=3, n_q=400, skip_frac=0.0)
    pb["p_skip"] = None
    # degenerate points: on the camera plane (PcZ == 0 -> inf / NaN projection), at the camera centre (dist 0), zero
    # invariance range, NaN coordinates
    G = pb["frustum"][0]
    b = int(G["q_begin"])
    pb["p_wP"][b] = G["Ow"]
    pb["p_min_dist"][b + 1] = 0; pb["p_max_dist"][b + 1] = 0
    pb["p_wP"][b + 2] = np.nan
    pb["p_max_dist"][b + 3] = np.inf
    # an empty frame in the middle of the batch
    pb["frustum"][1]["n_q"] = 0
    got = api.isInFrustum(pb)
    ref = O.is_in_frustum(pb)
    _same_frustum(got, ref)
    b1 = int(pb["frustum"][1]["q_begin"])
    assert not got["inview"][b1:b1 + 400].any() and (got["level"][b1:b1 + 400] == -1).all()
    # other pyramids (VR config: scale 2, 4 levels)
    for f in pb["frustum"]:
        f["log_scale_factor"] = np.log(np.float32(2.0)); f["n_levels"] = 4
    pb["frames"]["n_levels"] = 4
    _same_frustum(api.isInFrustum(pb), O.is_in_frustum(pb))
    pb["frustum"][0]["log_scale_factor"] = -1.0
    with pytest.raises(api.VieoError):
        api.isInFrustum(pb)


@pytest.mark.parametrize("kw", [dict(th=1.0, n_q=2500, blocked_frac=0.2), dict(th=3.0, th_far=8.0),
                                dict(th=6.0, cluster=True, n_kp=1500, n_q=600)])
def test_search_local_points_matches_oracle(kw):
    import vieo_slam_b200.api as api
    pb = synth.make_frustum_problem(43, n_frames=5, **kw)
    pb["frames"]["nn_ratio"] = 0.8
    fo, kp_match, q_match, q_dist, nm = api.ORBmatcher(0.8).SearchLocalPoints(pb)
    rfo, rkp, rqm, rqd, rnm = O.search_local_points(pb)
    _same_frustum(fo, rfo)
    for a, b, name in ((kp_match, rkp, "kp_match"), (q_match, rqm, "q_match"), (q_dist, rqd, "q_dist"), (nm, rnm, "n")):
        assert np.array_equal(a, b), (name, np.nonzero(a != b)[0][:10])
    assert nm.min() > 20
    assert (q_match[fo["inview"] == 0] == -1).all()
    # without the tracking info: same matches, same visibility flags
    fo2, kp2, qm2, qd2, nm2 = api.ORBmatcher(0.8).SearchLocalPoints(pb, want_tracking_info=False)
    assert np.array_equal(fo2["inview"], fo["inview"]) and np.array_equal(fo2["n_inview"], fo["n_inview"])
    assert np.array_equal(kp2, kp_match) and np.array_equal(qm2, q_match) and np.array_equal(qd2, q_dist) and np.array_equal(nm2, nm)
    # the stand-alone search fed with the frustum outputs (level -1 = skipped query) gives the same answer
    q = dict(pb); q["mode"] = 1
    q["q_proj"], q["q_level"], q["q_viewcos"], q["q_depth"] = fo["proj"], fo["level"], fo["viewcos"], fo["depth"]
    out = api.ORBmatcher(0.8).SearchByProjection(q)
    assert np.array_equal(out[1], q_match) and np.array_equal(out[3], nm)


def test_distinctive_descriptors_match_oracle():
    import vieo_slam_b200.api as api
    d = synth.make_distinctive_problem(51, n_points=3000, max_obs=40, long_lists=(64, 65, 255, 256, 257, 300, 700))
    best, med = api.ORBmatcher().ComputeDistinctiveDescriptors(d["pool"], d["ptr"], d["rows"])
    rb, rm = O.distinctive_descriptors(d["pool"], d["ptr"], d["rows"])
    assert np.array_equal(best, rb) and np.array_equal(med, rm), np.nonzero((best != rb) | (med != rm))[0][:10]
    assert (best == -1).sum() > 0
    # without a row table
    pool2 = np.ascontiguousarray(d["pool"][d["rows"]])
    b2, m2 = api.ORBmatcher().ComputeDistinctiveDescriptors(pool2, d["ptr"])
    assert np.array_equal(b2, rb) and np.array_equal(m2, rm)
    # errors are reported: row outside the pool, descending lists
    bad = d["rows"].copy(); bad[3] = len(d["pool"])
    with pytest.raises(api.VieoError):
        api.ORBmatcher().ComputeDistinctiveDescriptors(d["pool"], d["ptr"], bad)
    with pytest.raises(api.VieoError):
        api.ORBmatcher().ComputeDistinctiveDescriptors(d["pool"], [0, 5, 3], d["rows"])


@pytest.mark.parametrize("use_info", [True, False])
def test_initial_gyro_bias_and_reintegration(use_info):
    import vieo_slam_b200.api as api
    g = synth.make_gyro_bias_problem(61, n_kf=40, kf_gap=(1, 12))
    nz = O.imu_noise()
    imu = api.IMUPreintegrator()
    n_kf = len(g["kf_idx"])
    bias0 = np.zeros((n_kf, 6)); bias0[:, 3:] = g["seq"]["ba"]
    pre = imu.preintegrate_batch(g["samples"], g["seg_ptr"], g["ti_tj"], bias0)
    rng = np.random.default_rng(4)
    Rwb = np.stack([R @ synth.so3_exp(rng.normal(0, 2e-3, 3)) for R in g["Rwb"]])
    bg_in = np.array([1e-3, -2e-3, 5e-4])
    neq, bg, re = imu.OptimizeInitialGyroBias(pre, Rwb, bg_in, bInfo=use_info, samples=g["samples"], seg_ptr=g["seg_ptr"],
                                              ti_tj=g["ti_tj"], ba=bias0[:, 3:])
    rn, dbg = O.gyro_bias_init(pre, Rwb, use_info)
    assert neq == rn == n_kf - 1
    assert np.abs(bg - (bg_in + dbg)).max() < 1e-10 * np.abs(dbg).max(), (bg - bg_in, dbg)
    # every interval re-integrated with {bg, ba}: compare with the oracle at the SAME bias the device used
    for k in range(1, n_kf):
        ref = O.imu_preintegrate(g["samples"][g["seg_ptr"][k]:g["seg_ptr"][k + 1]], g["ti_tj"][k][0], g["ti_tj"][k][1], bg,
                                 bias0[k, 3:], nz)
        for f in ("Rij", "vij", "pij", "SigmaPRV", "JgR", "dt"):
            a, b = np.asarray(re[k][f]), np.asarray(ref[f])
            assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-300), (k, f)
    assert re[0]["dt"] == 0
    # without the sample lists only the bias is estimated; a chain without equations leaves bg untouched
    neq2, bg2, none = imu.OptimizeInitialGyroBias(pre, Rwb, bg_in, bInfo=use_info)
    assert neq2 == neq and none is None and np.array_equal(bg2, bg)
    neq3, bg3, _ = imu.OptimizeInitialGyroBias(pre[:1], Rwb[:1], bg_in)
    assert neq3 == 0 and np.array_equal(bg3, bg_in)
    # with exact rotations the step recovers the bias the gyro samples carry (first order)
    _, bg4, _ = imu.OptimizeInitialGyroBias(pre, g["Rwb"], np.zeros(3), bInfo=use_info)
    assert np.abs(bg4 - g["bg_true"]).max() < 3e-4


@pytest.mark.parametrize("kw", [dict(n_frames=6, n_q=2000), dict(use_bf=False, check_viewing_angle=False, th_radius=4.0),
                                dict(cluster=True, th_radius=6.0, n_kp=1500, n_q=800), dict(n_frames=1, n_q=9000, th_radius=2.5)])
def test_search_by_projection_base_matches_oracle(kw):
    """ORBmatcher::SearchByProjectionBase search half (Fuse / Sim3 / keyframe projection searches): bit-exact keypoint,
    distance and predicted level per map point."""
    import vieo_slam_b200.api as api
    pb = synth.make_fuse_problem(81, **kw)
    got = api.ORBmatcher().SearchByProjectionBase(pb)
    ref = O.proj_search(pb)
    for a, b, name in zip(got, ref, ("best_idx", "best_dist", "level")):
        assert np.array_equal(a, b), (name, np.nonzero(a != b)[0][:10])
    hit, nfused = api.ORBmatcher().Fuse(pb)
    assert nfused.min() > 100 and (hit >= 0).sum() == nfused.sum()
    assert ((ref[2] >= 0) & (ref[0] < 0)).sum() > 20


def test_search_by_projection_base_edge_cases():
    import vieo_slam_b200.api as api
    pb = synth.make_fuse_problem(82, n_frames=3, n_q=300)
    pb["frames"][1]["n_kp"] = 0      # keyframe without keypoints
    pb["frames"][2]["n_q"] = 0       # nothing to project
    pb["p_skip"] = None
    got = api.ORBmatcher().SearchByProjectionBase(pb)
    ref = O.proj_search(pb)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    b1 = int(pb["frames"][1]["q_begin"])
    assert (got[0][b1:b1 + 300] == -1).all()
    big = synth.make_fuse_problem(83, n_frames=1, n_kp=4200, n_q=50)
    with pytest.raises(api.VieoError):
        api.ORBmatcher().SearchByProjectionBase(big)


# ---------------------------------------------------------------- ORBmatcher::SearchForTriangulation
@pytest.mark.parametrize("seed,n_kp,n_nodes,share", [(1, 1200, 90, True), (2, 300, 6, False), (3, 2000, 400, True)])
def test_search_for_triangulation_matches_oracle(seed, n_kp, n_nodes, share):
    """Keyframe pairs of LocalMapping::CreateNewMapPoints: matches, their creation order, the per-keypoint map and the
    return value bit-exact against the oracle; few nodes -> long candidate lists (> 32 passing candidates: overflow path),
    shared / separate first keyframes, bOnlyStereo and mbCheckOrientation on and off."""
    import vieo_slam_b200.api as api
    pb = synth.make_sft_problem(seed, n_pairs=5, n_kp=n_kp, n_nodes=n_nodes, share_kf1=share)
    m12, po, nm = api.search_for_triangulation(pb)
    tot = 0
    for p in range(5):
        P = pb["pairs"][p]
        ref, n = O.search_for_triangulation(pb, p)
        assert nm[p] == n, (p, nm[p], n)
        ob = int(P["out_begin"])
        assert np.array_equal(po[ob:ob + n], ref), p
        want = np.full(int(P["n_kp1"]), -1, np.int32)
        want[ref[:, 0]] = ref[:, 1]
        assert np.array_equal(m12[ob:ob + int(P["n_kp1"])], want)
        tot += n
    assert tot > 50


def test_search_for_triangulation_matches_compiled_reference():
    """The CUDA path against the REFERENCE's own ORBmatcher::SearchForTriangulation (+ GeometricCamera::epipolarConstrain /
    FillMatchesFromPair) compiled unchanged (oracle/_ref/libref.so, built in the container and shipped prebuilt): the reference
    starts from keyframe poses, the C ABI from the F12 / epipole the caller forms from them (ref_sft_geometry)."""
    import ref_lib as R
    if not R.available():
        pytest.skip("oracle/_ref/libref.so not present")
    import vieo_slam_b200.api as api
    from vieo_slam_b200.synth import EUROC
    K4 = np.array([EUROC[k] for k in ("fx", "fy", "cx", "cy")], np.float32)
    pb = synth.make_sft_problem(41, n_pairs=6, n_kp=1200, n_nodes=80, share_kf1=False)
    poses = []
    for p in range(6):
        q1, t1, q2, t2 = R.sft_poses(pb, p, 41)
        ex, ey, F = R.sft_geometry(K4, q1, t1, q2, t2)
        pb["pairs"]["F12"][p] = F; pb["pairs"]["ex"][p] = ex; pb["pairs"]["ey"][p] = ey
        pb["pairs"]["only_stereo"][p] = p % 2; pb["pairs"]["check_orientation"][p] = (p // 2) % 2
        poses.append((q1, t1, q2, t2))
    m12, po, nm = api.search_for_triangulation(pb)
    tot = 0
    for p in range(6):
        ref, n = R.search_for_triangulation(pb, p, K4, *poses[p])
        ob = int(pb["pairs"][p]["out_begin"])
        assert nm[p] == n and np.array_equal(po[ob:ob + n], ref), (p, nm[p], n)
        tot += n
    assert tot > 300


def test_search_for_triangulation_edge_cases():
    import vieo_slam_b200.api as api
    pb = synth.make_sft_problem(7, n_pairs=2, n_kp=200, n_nodes=10)
    full = dict(pb); full["has_mp"] = np.ones_like(pb["has_mp"])
    assert (api.search_for_triangulation(full)[2] == 0).all()
    bad = dict(pb); bad["pairs"] = pb["pairs"].copy(); bad["pairs"]["n_kp2"][1] = 10 ** 6
    with pytest.raises(api.VieoError):
        api.search_for_triangulation(bad)
    none = dict(pb); none["pairs"] = pb["pairs"][:0]
    assert len(api.search_for_triangulation(none)[2]) == 0


# ---------------------------------------------------------------- ORBmatcher::SearchByBoW(KeyFrame, Frame)
@pytest.mark.parametrize("seed,n_kp,n_nodes", [(1, 1200, 90), (2, 400, 5), (3, 2500, 700)])
def test_search_by_bow_matches_oracle(seed, n_kp, n_nodes):
    """vpMapPointMatches (which keyframe keypoint's map point every frame keypoint received) and the return value bit-exact
    against the oracle; few nodes -> long lists with many equal distances, ratios 0.7 / 0.9, orientation check on / off."""
    import vieo_slam_b200.api as api
    pb = synth.make_bow_problem(seed, n_pairs=6, n_kp=n_kp, n_nodes=n_nodes)
    mf, nm = api.search_by_bow(pb)
    tot = 0
    for p in range(6):
        P = pb["pairs"][p]
        ref, n = O.search_by_bow(pb, p)
        ob = int(P["out_begin"])
        assert nm[p] == n, (p, nm[p], n)
        assert np.array_equal(mf[ob:ob + int(P["n_kp2"])], ref), p
        tot += n
    assert tot > 100
    none = dict(pb); none["mp_ok"] = np.zeros_like(pb["mp_ok"])
    assert (api.search_by_bow(none)[1] == 0).all()


# ---- Frame::isInFrustum with a camera rig (vieo_frustum_rig_batch, csrc/frustum.cu k_frustum_rig) -------------------------------
@pytest.mark.parametrize("model,n_cams,n_frames,n_q", [(2, 4, 3, 2500), (0, 2, 2, 900), (1, 3, 2, 900), (2, 1, 1, 300)])
def test_is_in_frustum_rig_matches_oracle(model, n_cams, n_frames, n_q):
    """Bit-exact floats, levels, camera masks and mean depths for every camera model.  The KB8 form evaluates atan2 in double on
    the device (<= 2 ulp from glibc's): a float pixel could differ in its last bit once in ~1e8 projections, so the few thousand of
    this test are compared exactly."""
    import vieo_slam_b200.api as api
    pb = synth.make_frustum_rig_problem(40 + model + n_cams, n_frames=n_frames, n_q=n_q, n_cams=n_cams, model=model)
    ref = O.is_in_frustum_rig(pb)
    got = api.frustum_rig_batch(pb)
    for k in ("inview", "cam_mask", "level", "n_inview"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("proj", "viewcos", "depth"):
        assert got[k].tobytes() == ref[k].tobytes(), k
    assert ref["n_inview"].sum() > 0.1 * n_q * n_frames / 2
    if n_cams >= 3:
        assert (np.bitwise_count(ref["cam_mask"]) > 1).sum() > 10


def test_is_in_frustum_rig_edge_cases():
    import vieo_slam_b200.api as api
    pb = synth.make_frustum_rig_problem(77, n_frames=2, n_q=400, n_cams=4, model=2, skip_frac=0.5)
    # a point on the optical axis of camera 0 (r <= 1e-5: the pinhole branch of KB8Camera::Project), NaN / inf positions, an empty frame
    G = pb["rig"][0]
    R = G["Rcw"].reshape(3, 3).astype(np.float64)
    from vieo_slam_b200.synth import R_from_quat
    C = G["cam"][0]
    Rcr = R_from_quat(np.array([C["q_cr"][3], C["q_cr"][0], C["q_cr"][1], C["q_cr"][2]], np.float64))
    Pc = np.array([0.0, 0.0, 3.0])
    Pcr = Rcr.T @ (Pc - C["t_cr"].astype(np.float64))
    Pw = R.T @ (Pcr - G["tcw"].astype(np.float64))
    b = int(G["q_begin"])
    pb["p_wP"][b] = Pw.astype(np.float32); pb["p_skip"][b] = 0
    pb["p_max_dist"][b] = 30.0; pb["p_min_dist"][b] = 0.1
    pb["p_normal"][b] = (Pw - G["Ow"]) / np.linalg.norm(Pw - G["Ow"])
    pb["p_wP"][b + 1] = [np.nan, 0, 1]; pb["p_skip"][b + 1] = 0
    pb["p_wP"][b + 2] = [np.inf, 1, 1]; pb["p_skip"][b + 2] = 0
    pb["rig"][1]["n_q"] = 0
    ref = O.is_in_frustum_rig(pb)
    got = api.frustum_rig_batch(pb)
    assert ref["inview"][b] == 1 and ref["cam_mask"][b] & 1
    for k in ("inview", "cam_mask", "level", "n_inview"):
        assert np.array_equal(got[k], ref[k]), k
    # a NaN position passes every comparison of the reference (all false) and is "in view" with NaN pixels on both sides; the
    # NaN payload bits are the only thing that differs between the host's and the device's arithmetic
    for k in ("proj", "viewcos", "depth"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
        fin = np.isfinite(ref[k])
        assert got[k][fin].tobytes() == ref[k][fin].tobytes(), k
    assert ref["inview"][b + 1] == 1 and np.isnan(ref["proj"][b + 1]).any()
    assert got["n_inview"][1] == 0
    bad = dict(pb); bad["rig"] = pb["rig"].copy(); bad["rig"]["n_cams"] = 5
    with pytest.raises(Exception):
        api.frustum_rig_batch(bad)

"""CPU tests pinning oracle/ba_oracle.cc (the reference's own tests hold no vectors for this path):
analytic Jacobians against g2o's central-difference recipe (base_multi_edge.hpp:67-107), the Schur solve against a
numpy dense solve of the full normal equations, closed-form / convergence properties of the drivers."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import EDGE_STEREO, NAVSTATE_DTYPE


@pytest.fixture(scope="module")
def seq():
    s = synth.vio_sequence(5, 60)
    s["pre"] = O.imu_preintegrate_frames(s, list(range(60)), O.imu_noise())
    return s


def test_reproject_jacobians_numeric():
    cam = synth.euroc_camera()
    r = np.random.default_rng(1)
    seq = synth.vio_sequence(3, 4)
    ns = seq["truth"][2]
    X = synth.landmarks_in_view(cam, ns, 8, r)
    for i in range(8):
        for stereo in (0, 1):
            obs = np.array([100, 200, 90], np.float32)
            e, Jp, JX, depth = O.edge_reproject(cam, ns, X[i], obs, stereo)
            assert depth > 0
            # the projection is rounded to float (camera_pinhole.h:81-82): numeric differences need a coarse step
            d = 2e-3
            Jn = np.zeros((3, 6)); JXn = np.zeros((3, 3))
            for c in range(6):
                dx = np.zeros(6); dx[c] = d
                ep = O.edge_reproject(cam, O.navstate_oplus(ns, 0, dx), X[i], obs, stereo)[0]
                em = O.edge_reproject(cam, O.navstate_oplus(ns, 0, -dx), X[i], obs, stereo)[0]
                Jn[:, c] = (ep - em) / (2 * d)
            for c in range(3):
                dx = np.zeros(3); dx[c] = d
                JXn[:, c] = (O.edge_reproject(cam, ns, X[i] + dx, obs, stereo)[0] - O.edge_reproject(cam, ns, X[i] - dx, obs, stereo)[0]) / (2 * d)
            rows = 3 if stereo else 2
            assert np.allclose(Jp[:rows], Jn[:rows], rtol=2e-3, atol=0.15), (Jp, Jn)
            assert np.allclose(JX[:rows], JXn[:rows], rtol=2e-3, atol=0.15)
            if not stereo:
                assert e[2] == 0 and np.all(Jp[2] == 0)


def test_reproject_float_rounding_and_stereo_row():
    cam = synth.euroc_camera()
    ns = np.zeros(1, NAVSTATE_DTYPE)[0]; ns["q"] = [1, 0, 0, 0]
    # body at the origin: Pc = Rcb X + tcb
    X = np.array([0.3, -0.2, 4.0])
    Pc = cam["Rcb"] @ X + cam["tcb"]
    u = np.float32(np.float64(cam["fx"]) * Pc[0] / Pc[2] + np.float64(cam["cx"]))
    obs = np.array([10.25, 20.5, 5.75], np.float32)
    e = O.edge_reproject(cam, ns, X, obs, 1)[0]
    assert abs(e[0] - (10.25 - float(u))) < 1e-12          # exact float pixel, not the double projection
    assert abs(e[2] - (5.75 - (float(u) - float(cam["bf"]) / Pc[2]))) < 1e-9


@pytest.mark.parametrize("order", [0, 1])
def test_navstate_edge_jacobians_numeric(seq, order):
    r = np.random.default_rng(2)
    i, j = 10, 11
    nsi = synth.perturb_state(seq["truth"][i], r); nsj = synth.perturb_state(seq["truth"][j], r)
    nsi["dbg"] = r.normal(0, 1e-3, 3); nsi["dba"] = r.normal(0, 1e-2, 3)
    pre = seq["pre"][j]
    gw = synth.GRAVITY_W
    e, Ji, Jj, Jb = O.edge_navstate(nsi, nsj, pre, gw, order)
    kind = 1 if order == 0 else None
    d = 1e-6

    def err(a, b):
        return O.edge_navstate(a, b, pre, gw, order)[0]

    def plus(ns, col, s):
        # column order of Ji/Jj == residual order: PVR (P,V,R) for order 0, PRV (P,R,V) for order 1
        dx9 = np.zeros(9); dx9[col] = s
        if order == 0:
            return O.navstate_oplus(ns, 1, dx9)
        out = O.navstate_oplus(ns, 0, dx9[:6])
        return O.navstate_oplus(out, 2, dx9[6:])
    for c in range(9):
        ni = (err(plus(nsi, c, d), nsj) - err(plus(nsi, c, -d), nsj)) / (2 * d)
        nj = (err(nsi, plus(nsj, c, d)) - err(nsi, plus(nsj, c, -d))) / (2 * d)
        assert np.allclose(Ji[:, c], ni, rtol=1e-5, atol=2e-5), (c, Ji[:, c], ni)
        assert np.allclose(Jj[:, c], nj, rtol=1e-5, atol=2e-5), (c, Jj[:, c], nj)
    for c in range(6):
        dx = np.zeros(6); dx[c] = d
        nb = (err(O.navstate_oplus(nsi, 3, dx), nsj) - err(O.navstate_oplus(nsi, 3, -dx), nsj)) / (2 * d)
        # the bias columns are the first-order model of the reference (Forster eq. 44): exact for p, v; approximate for R
        assert np.allclose(Jb[:, c], nb, rtol=1e-3, atol=1e-4), (c, Jb[:, c], nb)


def test_navstate_residual_near_zero_at_truth():
    s = synth.vio_sequence(9, 30, noisy_imu=False)
    pre = O.imu_preintegrate_frames(s, list(range(30)), O.imu_noise())
    for j in (5, 17, 29):
        for order in (0, 1):
            e = O.edge_navstate(s["truth"][j - 1], s["truth"][j], pre[j], synth.GRAVITY_W, order)[0]
            assert np.abs(e).max() < 2e-4, e  # mid-point integration error only


def test_prior_edge(seq):
    r = np.random.default_rng(3)
    pr = seq["truth"][7]
    ns = synth.perturb_state(pr, r, dbg=0, dba=0)
    ns["dbg"] = [1e-3, 0, 0]
    e, J = O.edge_prior_pvr(ns, pr)
    R0 = synth.R_from_quat(pr["q"]); R1 = synth.R_from_quat(ns["q"])
    assert np.allclose(e[:3], R0.T @ (ns["p"] - pr["p"]), atol=1e-12)
    assert np.allclose(e[3:6], ns["v"] - pr["v"], atol=1e-15)
    assert np.allclose(e[6:9], synth.so3_log(R0.T @ R1), atol=1e-9)
    assert np.allclose(e[9:12], [1e-3, 0, 0]) and np.allclose(e[12:], 0)
    d = 1e-6
    for c in range(9):
        dx = np.zeros(9); dx[c] = d
        n = (O.edge_prior_pvr(O.navstate_oplus(ns, 1, dx), pr)[0] - O.edge_prior_pvr(O.navstate_oplus(ns, 1, -dx), pr)[0]) / (2 * d)
        assert np.allclose(J[:, c], n, atol=1e-6)


def test_oplus_conventions():
    ns = np.zeros(1, NAVSTATE_DTYPE)[0]
    ns["q"] = synth.quat_from_R(synth.so3_exp(np.array([0.3, -0.2, 0.5]))); ns["p"] = [1, 2, 3]
    R = synth.R_from_quat(ns["q"])
    dx = np.array([0.1, -0.2, 0.05, 0.01, 0.02, -0.03])
    o = O.navstate_oplus(ns, 0, dx)
    assert np.allclose(o["p"], ns["p"] + R @ dx[:3], atol=1e-14)          # p <- p + R dp (USE_P_PLUS_RDP)
    assert np.allclose(synth.R_from_quat(o["q"]), R @ synth.so3_exp(dx[3:]), atol=1e-12)  # R <- R Exp(dphi)
    o = O.navstate_oplus(ns, 1, np.r_[dx[:3], [1, 2, 3], dx[3:]])
    assert np.allclose(o["v"], [1, 2, 3]) and np.allclose(synth.R_from_quat(o["q"]), R @ synth.so3_exp(dx[3:]), atol=1e-12)
    o = O.navstate_oplus(ns, 3, np.arange(6.0))
    assert np.allclose(o["dbg"], [0, 1, 2]) and np.allclose(o["dba"], [3, 4, 5]) and np.all(o["bg"] == 0)


def test_pose_optimization_imu_converges(seq):
    cam = synth.euroc_camera()
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=400, seed=1)
    pbs = pbs[:10]
    res, outl, chi2 = O.pose_optimization(pbs, cam, X, obs, w, fl)
    for k in range(len(pbs)):
        tru = seq["truth"][k + 1]
        assert np.linalg.norm(res[k]["cur"]["p"] - tru["p"]) < 0.012
        assert 0.7 * 400 < res[k]["n_inliers"] < 0.92 * 400
        assert res[k]["n_inliers"] == 400 - outl[k * 400:(k + 1) * 400].sum()
        M = res[k]["marg_cov_inv"]
        assert res[k]["prior_set"] == 1 and np.allclose(M, M.T, rtol=1e-9, atol=1e-6 * np.abs(M).max())
        assert np.linalg.eigvalsh((M + M.T) / 2).min() > -1e-6 * np.abs(M).max()
        assert np.all(M[:9, 9:] == 0)  # fixed last frame: PVR and bias blocks decouple (include/Optimizer.h:135-138)


def test_pose_optimization_visual_mode_and_few_points(seq):
    cam = synth.euroc_camera()
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=300, seed=2, mode=0, compute_marg=False)
    res, outl, _ = O.pose_optimization(pbs[:4], cam, X, obs, w, fl)
    for k in range(4):
        assert np.linalg.norm(res[k]["cur"]["p"] - seq["truth"][k + 1]["p"]) < 0.015
        assert np.array_equal(res[k]["cur"]["v"], pbs[k]["cur"]["v"])  # PR vertex: velocity untouched
    pb = pbs[:1].copy(); pb["edge_end"] = pb["edge_begin"] + 2    # < 3 correspondences -> return 0 (src/Optimizer.cc:1787)
    res, _, _ = O.pose_optimization(pb, cam, X, obs, w, fl)
    assert res[0]["n_inliers"] == 0 and res[0]["iterations"] == 0


def test_pose_optimization_free_last_frame_prior(seq):
    cam = synth.euroc_camera()
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=350, seed=3, chain_prior=True)
    res, _, _ = O.pose_optimization(pbs[1:2], cam, X, obs, w, fl)  # k = 2: last frame free with a 15-dim prior
    assert pbs[1]["last_has_prior"] == 1
    M = res[0]["marg_cov_inv"]
    assert np.isfinite(M).all() and np.abs(M[:9, 9:]).max() > 0  # Schur over the last frame couples PVR and bias
    assert np.linalg.norm(res[0]["cur"]["p"] - seq["truth"][2]["p"]) < 0.012
    assert not np.array_equal(res[0]["last"]["p"], pbs[1]["last"]["p"])


def _lba(seq, **kw):
    cam = synth.euroc_camera()
    kf = list(range(0, 60, 3))
    pre = O.imu_preintegrate_frames(seq, kf, O.imu_noise())
    return cam, synth.make_lba_problem(seq, pre, kf, cam, n_local=8, n_fixed=6, n_points=300, seed=4, **kw)


def _dense_normal_equations(d, cam, scale=None, gdir=None):
    """Full (poses [+ scale] + points) normal equations of the LBA graph assembled from the per-edge oracle functions,
    independently of the oracle's Schur code.  Layout: [pose vertices (n) | scale (1, if given) | points (3P)]."""
    K, P = len(d["states"]), len(d["points"])
    off = {}; n = 0
    for k in range(K):
        f = d["state_flags"][k]
        if not f & 1:
            off[(k, 0)] = n; n += 6
        if f & 2 and not f & 4:
            off[(k, 1)] = n; n += 3
            off[(k, 2)] = n; n += 6
    ns = n + (1 if scale is not None else 0) + (2 if gdir is not None else 0)   # gdir = (q_wI, GI): two more columns
    N = ns + 3 * P
    H = np.zeros((N, N)); b = np.zeros(N)

    def huber(c, delta):
        d2 = float(np.float32(delta * delta))
        return 1.0 if c <= d2 else delta / np.sqrt(c)
    dm, ds = float(np.float32(np.sqrt(np.float32(5.991)))), float(np.float32(np.sqrt(7.815)))
    for i in range(len(d["edge_state"])):
        s, p = d["edge_state"][i], d["edge_point"][i]
        st = bool(d["edge_flags"][i] & EDGE_STEREO)
        if scale is None:
            e, Jp, JX, _ = O.edge_reproject(cam, d["states"][s], d["points"][p], d["obs"][i], st)
        else:
            e, Jp, JX, Js = O.edge_reproject_scale(cam, d["states"][s], d["points"][p], scale, d["obs"][i], st)
        rows = 3 if st else 2
        w = float(d["inv_sigma2"][i]); c = w * (e[:rows] ** 2).sum()
        w *= huber(c, ds if st else dm)
        blocks = [(ns + 3 * p, JX[:rows])]
        if (s, 0) in off:
            blocks.append((off[(s, 0)], Jp[:rows]))
        if scale is not None:
            blocks.append((n, Js[:rows].reshape(rows, 1)))
        for oa, Ja in blocks:
            b[oa:oa + Ja.shape[1]] -= Ja.T @ (w * e[:rows])
            for ob, Jb_ in blocks:
                H[oa:oa + Ja.shape[1], ob:ob + Jb_.shape[1]] += Ja.T @ (w * Jb_)
    for m in range(len(d["imu_i"])):
        i, j = d["imu_i"][m], d["imu_j"][m]
        pre = d["preint"][m]
        fixed = bool(d["state_flags"][i] & 1)
        if gdir is None:
            e, Ji, Jj, Jb = O.edge_navstate(d["states"][i], d["states"][j], pre, d["gw"], 1)
        else:
            e, Ji, Jj, Jb, JG = O.edge_navstate_g(d["states"][i], d["states"][j], pre, gdir[0], gdir[1])
        info = np.linalg.inv(pre["SigmaPRV"]) * (1e-2 if fixed else 1.0)
        c = e @ info @ e
        wI = huber(c, float(np.float32(np.sqrt(16.919)))) if fixed else 1.0
        blocks = [((i, 0), Ji[:, :6]), ((j, 0), Jj[:, :6]), ((i, 1), Ji[:, 6:]), ((j, 1), Jj[:, 6:]), ((i, 2), Jb)]
        blocks = [(off[k], J) for k, J in blocks if k in off]
        if gdir is not None:
            blocks.append((ns - 2, JG))
        for oa, Ja in blocks:
            b[oa:oa + Ja.shape[1]] -= Ja.T @ (wI * info @ e)
            for ob, Jb_ in blocks:
                H[oa:oa + Ja.shape[1], ob:ob + Jb_.shape[1]] += Ja.T @ (wI * info) @ Jb_
        si, sj = d["states"][i], d["states"][j]
        eb = np.r_[sj["bg"] + sj["dbg"] - si["bg"] - si["dbg"], sj["ba"] + sj["dba"] - si["ba"] - si["dba"]]
        ib = np.diag([d["inv_sigma_bg2"]] * 3 + [d["inv_sigma_ba2"]] * 3) / pre["dt"] * (1e-2 if fixed else 1.0)
        wb = huber(eb @ ib @ eb, float(np.float32(np.sqrt(12.592)))) if fixed else 1.0
        blocks = [(off[k], J) for k, J in (((i, 2), -np.eye(6)), ((j, 2), np.eye(6))) if k in off]
        for oa, Ja in blocks:
            b[oa:oa + 6] -= Ja.T @ (wb * ib @ eb)
            for ob, Jb_ in blocks:
                H[oa:oa + 6, ob:ob + 6] += Ja.T @ (wb * ib) @ Jb_
    return H, b, ns


def test_schur_step_matches_dense_normal_equations(seq):
    """F3: the oracle's Schur-complement solve equals a numpy solve of the full (poses + points) system assembled
    independently from the per-edge Jacobians."""
    cam, d = _lba(seq)
    lam = 1.0
    xp, xl, chi2 = O.ba_debug_step(d, cam, lam)
    H, b, n = _dense_normal_equations(d, cam)
    assert n == len(xp)
    x = np.linalg.solve(H + lam * np.eye(len(b)), b)
    assert np.allclose(xp, x[:n], rtol=1e-7, atol=1e-10)
    assert np.allclose(xl.ravel(), x[n:], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("scale0", [1.0, 0.95])
def test_schur_step_with_scale_vertex_matches_dense_normal_equations(seq, scale0):
    """The scale row / column of the reduced camera system (VertexScale + EdgeReprojectPRS, bScaleOpt) against the dense
    solve: pose-scale, scale-scale and scale-landmark blocks all enter the Schur complement."""
    cam, d = _lba(seq)
    d = dict(d)
    d["points"] = d["points"] / scale0          # unscaled landmarks: the world points stay where they were
    lam = 0.7
    xp, xl, chi2 = O.ba_debug_step_scale(d, cam, lam, scale0)
    H, b, n = _dense_normal_equations(d, cam, scale=scale0)
    assert n == len(xp)
    x = np.linalg.solve(H + lam * np.eye(len(b)), b)
    assert np.allclose(xp, x[:n], rtol=1e-6, atol=1e-9), np.abs(xp - x[:n]).max()
    assert np.allclose(xl.ravel(), x[n:], rtol=1e-6, atol=1e-9)
    assert abs(xp[-1]) > 1e-9                   # the scale does move


def test_local_ba_prv_converges(seq):
    cam, d = _lba(seq)
    out = O.local_ba_prv(d, cam)
    res = out["res"]
    assert res["accepted"] == 1 and res["err_end"] < res["err0"]
    assert tuple(res["iterations"]) == (4, 6) or res["iterations"][0] <= 4
    # fixed keyframes untouched, local ones moved towards the truth
    kf = list(range(0, 60, 3))
    nk = len(kf); local = list(range(nk - 8, nk))
    err_in = [np.linalg.norm(d["states"][s]["p"] - seq["truth"][kf[k]]["p"]) for s, k in enumerate(local)]
    err_out = [np.linalg.norm(out["states"][s]["p"] - seq["truth"][kf[k]]["p"]) for s, k in enumerate(local)]
    assert np.mean(err_out) < 0.5 * np.mean(err_in)
    for s in range(8, len(d["states"])):
        assert out["states"][s].tobytes() == d["states"][s].tobytes()
    assert 0.01 < out["erase"].mean() < 0.15  # ~5 % outlier observations flagged for ErasePairObs


def test_local_ba_large_and_visual_only(seq):
    cam, d = _lba(seq)
    a = O.local_ba_prv(d, cam, large=True)["res"]
    assert a["accepted"] == 1 and max(a["iterations"]) <= 2
    v = O.local_ba_prv(d, cam, visual_only=True)
    assert v["res"]["err_end"] < v["res"]["err0"]
    assert np.array_equal(v["states"]["v"], d["states"]["v"])  # PR-only vertices: velocities and biases untouched


def test_all_fixed_returns_untouched(seq):
    cam, d = _lba(seq)
    d = dict(d); d["state_flags"] = d["state_flags"] | 1
    out = O.local_ba_prv(d, cam)
    assert out["states"].tobytes() == d["states"].tobytes() and out["res"]["iterations"][0] == 0


# ---- camera models: radtan and KB8 next to pinhole ----------------------------------------------
def _cam_points(cam, n, seed):
    r = np.random.default_rng(seed)
    ns = np.zeros(1, NAVSTATE_DTYPE)[0]; ns["q"] = [1, 0, 0, 0]
    return ns, synth.landmarks_in_view(cam, ns, n, r, zmin=0.5, zmax=8.0)


@pytest.mark.parametrize("model", ["radtan", "radtan3", "kb8"])
def test_camera_models_projection_matches_cv2(model):
    """Pins Project() of camera_radtan.h:61-129 / camera_kb8.h:68-157 against OpenCV's projectPoints (the reference
    calibrations are OpenCV/Kalibr ones): the float pixel of the oracle equals cv2's double pixel rounded to float up
    to one ulp-of-pixel."""
    cv2 = pytest.importorskip("cv2")
    cam = {"radtan": synth.radtan_camera(), "radtan3": synth.radtan_camera(k=(-0.28, 0.074, -0.01)),
           "kb8": synth.kb8_camera(k=(-0.0135, 0.021, -0.03, 0.012))}[model]
    ns, X = _cam_points(cam, 64, 3)
    Pc = X @ cam["Rcb"].T + cam["tcb"]
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float64)
    if model == "kb8":
        ref = cv2.fisheye.projectPoints(Pc.reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, cam["dist"][:4].astype(np.float64))[0]
    else:
        nk = int(cam["num_k"])
        d = cam["dist"].astype(np.float64)
        dc = np.array([d[0], d[1], d[nk], d[nk + 1], d[2] if nk > 2 else 0.0])
        ref = cv2.projectPoints(Pc.reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, dc)[0]
    ref = ref.reshape(-1, 2)
    for i in range(len(X)):
        obs = np.zeros(3, np.float32)
        e = O.edge_reproject(cam, ns, X[i], obs, 0)[0]
        assert np.allclose(-e[:2], ref[i], rtol=0, atol=1e-4), (model, i, -e[:2], ref[i])
    su, sv, _, _ = synth.project(cam, ns, X)
    assert np.allclose(np.stack([su, sv], 1), ref, atol=1e-9)  # the generator's lens model is the same one


@pytest.mark.parametrize("model", ["radtan_tangential", "kb8"])
def test_camera_models_jacobians_numeric(model):
    # radtan: the reference's d(img)/d(p3d) carries the radial derivative only from k3 on (camera_radtan.h:85-99), so it
    # is exact for tangential-only distortion; KB8's is exact everywhere
    cam = synth.radtan_camera(k=(0.0, 0.0), p=(2e-3, -1.5e-3)) if model != "kb8" else synth.kb8_camera(k=(-0.0135, 0.021, -0.03, 0.012))
    ns, X = _cam_points(cam, 8, 5)
    obs = np.array([100, 200, 90], np.float32)
    d = 2e-3
    for i in range(8):
        for stereo in (0, 1):
            e, Jp, JX, depth = O.edge_reproject(cam, ns, X[i], obs, stereo)
            JXn = np.zeros((3, 3)); Jn = np.zeros((3, 6))
            for c in range(3):
                dx = np.zeros(3); dx[c] = d
                JXn[:, c] = (O.edge_reproject(cam, ns, X[i] + dx, obs, stereo)[0] - O.edge_reproject(cam, ns, X[i] - dx, obs, stereo)[0]) / (2 * d)
            for c in range(6):
                dx = np.zeros(6); dx[c] = d
                Jn[:, c] = (O.edge_reproject(cam, O.navstate_oplus(ns, 0, dx), X[i], obs, stereo)[0]
                            - O.edge_reproject(cam, O.navstate_oplus(ns, 0, -dx), X[i], obs, stereo)[0]) / (2 * d)
            rows = 3 if stereo else 2
            assert np.allclose(JX[:rows], JXn[:rows], rtol=2e-3, atol=0.15), (model, JX, JXn)
            assert np.allclose(Jp[:rows], Jn[:rows], rtol=2e-3, atol=0.15)


def test_radtan_jacobian_is_the_references_formula():
    """With radial coefficients the reference's Jacobian is NOT the exact derivative; the oracle must reproduce the
    formula as written (camera_radtan.h:85-99), restated independently here in numpy."""
    cam = synth.radtan_camera(k=(-0.28, 0.074, -0.01))
    ns, X = _cam_points(cam, 6, 9)
    k = cam["dist"][:3].astype(np.float64); p = cam["dist"][3:5].astype(np.float64)
    fx, fy = float(cam["fx"]), float(cam["fy"])
    for i in range(6):
        Pc = cam["Rcb"] @ X[i] + cam["tcb"]
        invz = 1 / Pc[2]; x = Pc[0] * invz; y = Pc[1] * invz; r2 = x * x + y * y
        fd = 1 + k[0] * r2 + k[1] * r2 * r2 + k[2] * r2 ** 3
        fd2 = 2 * k[2]
        du_dx = fx * invz * (fd + fd2 * x * x + 2 * (p[0] * y + 3 * p[1] * x))
        du_dy = fx * invz * (fd2 * x * y + 2 * (p[0] * x + p[1] * y))
        dv_dx = du_dy * fy / fx
        dv_dy = fy * invz * (fd + fd2 * y * y + 2 * (p[1] * x + 3 * p[0] * y))
        J = np.array([[du_dx, du_dy, -(x * du_dx + y * du_dy)], [dv_dx, dv_dy, -(x * dv_dx + y * dv_dy)]])
        JX = O.edge_reproject(cam, ns, X[i], np.zeros(3, np.float32), 0)[2]
        assert np.allclose(JX[:2], -J @ cam["Rcb"], rtol=1e-9, atol=1e-9), np.abs(JX[:2] + J @ cam["Rcb"]).max()  # J_X = Jproj Rcw, Jproj = -d(img)/d(p3d)


@pytest.mark.parametrize("model", ["radtan", "kb8"])
def test_drivers_converge_with_lens_models(seq, model):
    cam = synth.radtan_camera() if model == "radtan" else synth.kb8_camera()
    pbs, Xw, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=300, seed=2)
    pbs = pbs[:4]
    res, outl, chi2 = O.pose_optimization(pbs, cam, Xw, obs, w, fl)
    for k in range(len(pbs)):
        assert np.linalg.norm(res[k]["cur"]["p"] - seq["truth"][k + 1]["p"]) < 0.02
        assert 0.7 * 300 < res[k]["n_inliers"] < 0.93 * 300
    # the same observations through the wrong (pinhole) model must fit visibly worse
    res0, _, _ = O.pose_optimization(pbs, synth.euroc_camera(), Xw, obs, w, fl)
    assert sum(r["n_inliers"] for r in res0) < sum(r["n_inliers"] for r in res)


# ---- global BA ------------------------------------------------------------------------------------
def _gba(n_kf=12, n_points=300, seed=3, **kw):
    s = synth.vio_sequence(40 + seed, 4 * n_kf + 1, speed=1.0, rot=0.6)
    kf = list(range(0, 4 * n_kf, 4))
    pre = O.imu_preintegrate_frames(s, kf, O.imu_noise())
    cam = synth.euroc_camera()
    return s, kf, cam, synth.make_gba_problem(s, pre, kf, cam, n_points=n_points, seed=seed, **kw)


def test_global_ba_converges_and_keeps_keyframe0():
    s, kf, cam, d = _gba()
    assert d["state_flags"][0] == 7 and np.all(d["state_flags"][1:] == 2)
    out = O.global_ba_prv(d, cam, n_iterations=10, robust=False)
    assert out["iterations"] >= 3
    assert out["res"]["err_end"] < 0.05 * out["res"]["err0"]
    assert out["states"][0].tobytes() == d["states"][0].tobytes()       # keyframe 0: PR, V and Bias fixed
    e0 = max(np.linalg.norm(d["states"][k]["p"] - s["truth"][kf[k]]["p"]) for k in range(1, len(kf)))
    e1 = max(np.linalg.norm(out["states"][k]["p"] - s["truth"][kf[k]]["p"]) for k in range(1, len(kf)))
    assert e1 < 0.5 * e0 and e1 < 0.01


def test_global_ba_robust_flag_changes_kernels():
    _, _, cam, d = _gba(outlier_frac=0.05)
    a = O.global_ba_prv(d, cam, n_iterations=6, robust=True)
    b = O.global_ba_prv(d, cam, n_iterations=6, robust=False)
    # with outliers, the Huber cost is far below the plain quadratic one and the estimates differ
    assert a["res"]["err0"] < 0.7 * b["res"]["err0"]
    assert np.abs(a["states"]["p"] - b["states"]["p"]).max() > 1e-5
    # global BA uses sqrt(5.99), not the local BA's sqrt(5.991f): a mono edge with chi2 between the two deltas^2 tells
    c = O.local_ba_prv(d, cam)
    assert c["res"]["err0"] != a["res"]["err0"]


def test_scale_edge_jacobians_numeric():
    """EdgeReprojectPRS[Stereo] (g2otypes.h:321-541, scale vertex of GlobalBundleAdjustmentNavStatePRV with bScaleOpt):
    error at Xw = s * Xh, point block scaled by s, scale block = (Jproj Rcw) Xh — against central differences."""
    cam = synth.euroc_camera()
    r = np.random.default_rng(4)
    ns = synth.vio_sequence(3, 4)["truth"][2]
    X = synth.landmarks_in_view(cam, ns, 6, r)
    obs = np.array([100, 200, 90], np.float32)
    for i in range(6):
        for stereo in (0, 1):
            for s in (1.0, 0.93, 1.2):
                Xh = X[i] / s                     # the same world point through a different scale estimate
                e, Jp, JX, Js = O.edge_reproject_scale(cam, ns, Xh, s, obs, stereo)
                e0, Jp0, JX0, _ = O.edge_reproject(cam, ns, X[i], obs, stereo)
                assert np.allclose(e, e0, atol=1e-4) and np.allclose(Jp, Jp0, rtol=1e-6, atol=1e-6)   # float-rounded projection
                assert np.allclose(JX, JX0 * s, rtol=1e-9, atol=1e-9)
                d = 2e-3
                Jsn = (O.edge_reproject_scale(cam, ns, Xh, s + d, obs, stereo)[0] -
                       O.edge_reproject_scale(cam, ns, Xh, s - d, obs, stereo)[0]) / (2 * d)
                JXn = np.zeros((3, 3))
                for c in range(3):
                    dx = np.zeros(3); dx[c] = d
                    JXn[:, c] = (O.edge_reproject_scale(cam, ns, Xh + dx, s, obs, stereo)[0] -
                                 O.edge_reproject_scale(cam, ns, Xh - dx, s, obs, stereo)[0]) / (2 * d)
                rows = 3 if stereo else 2
                assert np.allclose(Js[:rows], Jsn[:rows], rtol=3e-3, atol=0.3), (Js, Jsn)
                assert np.allclose(JX[:rows], JXn[:rows], rtol=3e-3, atol=0.3)


def test_gravity_direction_vertex_and_edge(seq):
    """VertexGThetaXYRwI + EdgeNavStatePRVG (g2otypes.h:674-698, 725-884): RwI * GI reproduces gw, the edge equals the PRV
    edge at that gravity, the 9x2 block matches central differences of the 2-dim right update, rows R are zero."""
    r = np.random.default_rng(6)
    G = 9.81
    GI = np.array([0, 0, G])
    # exactly (anti-)parallel to z: a_wI is the zero vector, Eigen's normalized() leaves it, RwI = I (reference behaviour)
    assert np.allclose(O.gdir_init(np.array([0, 0, -G])), [1, 0, 0, 0]) and np.allclose(O.gdir_init(GI), [1, 0, 0, 0])
    for gw in (np.array([0.02, -0.01, -9.81]), np.array([0.3, -0.2, -9.7]), np.array([1.0, 2.0, 9.5])):
        gw = gw / np.linalg.norm(gw) * G
        q = O.gdir_init(gw)
        assert np.allclose(synth.R_from_quat(q) @ GI, gw, atol=1e-9)
        i, j = 20, 21
        nsi = synth.perturb_state(seq["truth"][i], r); nsj = synth.perturb_state(seq["truth"][j], r)
        pre = seq["pre"][j]
        e, Ji, Jj, Jb, JG = O.edge_navstate_g(nsi, nsj, pre, q, GI)
        e0, Ji0, Jj0, Jb0 = O.edge_navstate(nsi, nsj, pre, gw, 1)
        assert np.allclose(e, e0, atol=1e-12) and np.allclose(Ji, Ji0) and np.allclose(Jj, Jj0) and np.allclose(Jb, Jb0)
        d = 1e-6
        for c in range(2):
            dx = np.zeros(2); dx[c] = d
            num = (O.edge_navstate_g(nsi, nsj, pre, O.gdir_oplus(q, dx), GI)[0] -
                   O.edge_navstate_g(nsi, nsj, pre, O.gdir_oplus(q, -dx), GI)[0]) / (2 * d)
            assert np.allclose(JG[:, c], num, rtol=1e-5, atol=1e-7), (c, JG[:, c], num)
        assert not JG[3:6].any() and np.abs(JG[6:9]).max() > 0.1 * pre["dt"] * G
        # a rotation about the gravity axis itself (third component) is not a degree of freedom: the update ignores it
        q2 = O.gdir_oplus(q, [0.0, 0.0])
        assert np.allclose(q2, q, atol=1e-15)


def test_global_ba_with_scale_vertex(seq):
    """GlobalBundleAdjustmentNavStatePRV with bScaleOpt (System::FinalGBA): same cost as the plain global BA on the same
    map (the scale is a gauge direction of the free landmarks), landmarks stored at a wrong common scale are absorbed,
    keyframe 0 stays fixed and the returned points are the scaled ones."""
    cam = synth.euroc_camera()
    kf = list(range(0, 60, 3))
    pre = O.imu_preintegrate_frames(seq, kf, O.imu_noise())
    g = synth.make_gba_problem(seq, pre, kf, cam, n_points=400, seed=3)
    plain = O.global_ba_prv(g, cam, n_iterations=10, robust=False)
    sc = O.global_ba_prv_scale(g, cam, n_iterations=10, robust=False)
    assert sc["res"]["err_end"] < 1.05 * plain["res"]["err_end"] and sc["res"]["err_end"] < 0.2 * sc["res"]["err0"]
    assert sc["states"][0]["p"].tobytes() == g["states"][0]["p"].tobytes()
    assert abs(sc["scale"] - 1.0) < 0.05
    assert np.median(np.linalg.norm(sc["points"] - plain["points"], axis=1)) < 0.02
    # landmarks 8 % too small: with the metric scale pinned by the inertial edges the estimate s * X returns to the truth
    g2 = dict(g); g2["points"] = g["points"] / 1.08
    sc2 = O.global_ba_prv_scale(g2, cam, n_iterations=15, robust=False)
    assert sc2["res"]["err_end"] < 1.2 * plain["res"]["err_end"]
    assert np.median(np.linalg.norm(sc2["points"] - plain["points"], axis=1)) < 0.03


def test_schur_step_with_gravity_direction_vertex_matches_dense_normal_equations(seq):
    """The two gravity-direction columns (VertexGThetaXYRwI + EdgeNavStatePRVG) of the reduced camera system against the
    dense solve of the full normal equations."""
    cam, d = _lba(seq)
    gw = np.array([0.35, -0.25, -9.79]); gw *= 9.81 / np.linalg.norm(gw)     # a tilted estimate: the residuals see it
    lam = 0.5
    xp, xl, chi2 = O.ba_debug_step_gdir(d, cam, lam, gw)
    H, b, n = _dense_normal_equations(d, cam, gdir=(O.gdir_init(gw), np.array([0, 0, 9.81])))
    assert n == len(xp)
    x = np.linalg.solve(H + lam * np.eye(len(b)), b)
    assert np.allclose(xp, x[:n], rtol=1e-6, atol=1e-9), np.abs(xp - x[:n]).max()
    assert np.allclose(xl.ravel(), x[n:], rtol=1e-6, atol=1e-9)
    assert np.abs(xp[-2:]).max() > 1e-6


def test_global_ba_of_the_imu_initialiser_refines_gravity(seq):
    """GlobalBundleAdjustmentNavStatePRV as IMUInitialization calls it (gravity-direction vertex, keyframe 0 with free V /
    Bias, prior-bias edge): a gravity estimate 2 degrees off is pulled back to the true direction, its norm is kept, the
    cost falls and keyframe 0's pose stays fixed."""
    cam = synth.euroc_camera()
    kf = list(range(0, 60, 3))
    pre = O.imu_preintegrate_frames(seq, kf, O.imu_noise())
    g = dict(synth.make_gba_problem(seq, pre, kf, cam, n_points=400, seed=3))
    flags = g["state_flags"].copy(); flags[0] = 1 | 2          # keyframe 0: PR fixed, V / Bias free (:825-831)
    g["state_flags"] = flags
    tilt = synth.so3_exp(np.array([np.deg2rad(2.0), np.deg2rad(-1.0), 0.0]))
    gw0 = tilt @ synth.GRAVITY_W
    out = O.global_ba_prv_init(g, cam, 15, gw0)

    def ang(a, b):
        return np.degrees(np.arccos(np.clip(a @ b / np.linalg.norm(a) / np.linalg.norm(b), -1, 1)))
    assert ang(gw0, synth.GRAVITY_W) > 2.0
    assert ang(out["gw"], synth.GRAVITY_W) < 0.3, ang(out["gw"], synth.GRAVITY_W)
    assert abs(np.linalg.norm(out["gw"]) - np.linalg.norm(gw0)) < 1e-9
    assert out["res"]["err_end"] < 0.2 * out["res"]["err0"]
    assert out["states"][0]["p"].tobytes() == g["states"][0]["p"].tobytes()
    assert out["states"][0]["q"].tobytes() == g["states"][0]["q"].tobytes()
    assert not np.array_equal(out["states"][0]["v"], g["states"][0]["v"])      # free now
    b0 = g["states"][0]["bg"] + g["states"][0]["dbg"]; b1 = out["states"][0]["bg"] + out["states"][0]["dbg"]
    assert np.abs(b1 - b0).max() < 5e-3                                        # held by the prior-bias edge

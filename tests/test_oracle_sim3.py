"""Oracle of Optimizer::OptimizeSim3 (src/Optimizer.cc:2689-2920; edges src/Odom/g2otypes.h:321-549, MODE 1 / 2) pinned:
the two edges by numeric differentiation through the vertices' own oplus (NavState::IncSmallPR, VertexScale += d) and by the
closed form of the similarity they encode; the driver by recovering a known Sim3 and by its exit rules."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import NAVSTATE_DTYPE


def pinhole_cam():
    cam = synth.euroc_camera()
    cam["Rcb"] = np.eye(3)  # OptimizeSim3: SetParams(CamInst) leaves Rcb = I, tcb = 0 for a single pinhole camera
    cam["tcb"] = 0
    return cam


def _ns(R12, t12):
    ns = np.zeros(1, NAVSTATE_DTYPE)[0]
    ns["q"] = synth.quat_from_R(R12.T)
    ns["p"] = -(R12.T @ t12)
    return ns


@pytest.mark.parametrize("body_cam", [False, True])
@pytest.mark.parametrize("inverse,s", [(0, 1.13), (0, 1.0), (1, 1.0), (1, 1.13)])
def test_edge_is_the_similarity_and_its_jacobian(inverse, s, body_cam):
    r = np.random.default_rng(5 + inverse)
    cam = synth.euroc_camera() if body_cam else pinhole_cam()
    R12 = synth.so3_exp(r.normal(0, 0.2, 3)); t12 = r.normal(0, 0.3, 3)
    ns = _ns(R12, t12)
    Xh = np.array([0.4, -0.3, 4.0])
    obs = np.array([300.0, 200.0], np.float32)
    e, Jp, Js = O.edge_sim3(cam, ns, s, Xh, obs, inverse)
    if not body_cam:  # closed form: x1 = s R12 X2 + t12, resp. x2 = R12^T (X1 - t12) / s
        Pc = s * R12 @ Xh + t12 if not inverse else R12.T @ (Xh - t12) / s
        u = np.float32(cam["fx"] * Pc[0] / Pc[2] + cam["cx"]); v = np.float32(cam["fy"] * Pc[1] / Pc[2] + cam["cy"])
        assert np.allclose(e, [obs[0] - u, obs[1] - v], atol=1e-4)
    # Numeric Jacobians through the vertices' oplus (the projection is rounded to float pixels, so the steps are large).
    # MODE 1 is exact for every s.  MODE 2 as the reference writes it (g2otypes.h:456-459, 524, 538-539) is exact at s = 1
    # (bFixScale: every stereo / VIO configuration); for s != 1 its position block lacks the 1 / s of "tcw *= scale" (:371)
    # and its scale column uses the scaled tcw where the comment says unscaled — Gauss-Newton approximations the oracle
    # restates as they are: there only the rotation block and the 1 / s relation of the position block are checked.
    exact = (not inverse) or s == 1.0
    h = 1e-3
    for a in range(6):
        d = np.zeros(6); d[a] = h
        ep, _, _ = O.edge_sim3(cam, O.navstate_oplus(ns, 0, d), s, Xh, obs, inverse)
        em, _, _ = O.edge_sim3(cam, O.navstate_oplus(ns, 0, -d), s, Xh, obs, inverse)
        num = (ep - em) / (2 * h)
        ref = Jp[:, a] if (exact or a >= 3) else Jp[:, a] / s
        assert np.allclose(num, ref, rtol=2e-2, atol=0.3), (a, num, Jp[:, a])
    if exact:
        ep, _, _ = O.edge_sim3(cam, ns, s + h, Xh, obs, inverse)
        em, _, _ = O.edge_sim3(cam, ns, s - h, Xh, obs, inverse)
        assert np.allclose((ep - em) / (2 * h), Js, rtol=2e-2, atol=0.3)


@pytest.mark.parametrize("fix_scale", [False, True])
def test_recovers_the_similarity(fix_scale):
    cam = pinhole_cam()
    pbs, X1, X2, o1, o2, w1, w2, truth = synth.make_sim3_problems(cam, n_candidates=3, n_matches=150, seed=3, fix_scale=fix_scale)
    res, keep, c12, c21 = O.optimize_sim3(pbs, cam, X1, X2, o1, o2, w1, w2)
    for k, (R12, t12, s) in enumerate(truth):
        Rwb = synth.R_from_quat(res[k]["ns"]["q"])
        Re = Rwb.T
        te = -(Re @ res[k]["ns"]["p"])
        assert np.linalg.norm(synth.so3_log(Re.T @ R12)) < np.deg2rad(0.3)
        assert np.linalg.norm(te - t12) < 0.03
        assert abs(res[k]["scale"] - s) < (1e-12 if fix_scale else 0.01)
        m = slice(pbs[k]["m_begin"], pbs[k]["m_end"])
        assert res[k]["n_inliers"] == int(keep[m].sum()) > 90
        assert res[k]["n_corr"] == 150 and res[k]["n_bad"] > 5
        # kept pairs pass both gates, dropped ones of the second stage fail one
        assert np.all((c12[m][keep[m] == 1] <= 10.0) & (c21[m][keep[m] == 1] <= 10.0))


def test_too_few_inliers_returns_zero_and_keeps_the_input():
    cam = pinhole_cam()
    pbs, X1, X2, o1, o2, w1, w2, _ = synth.make_sim3_problems(cam, n_candidates=2, n_matches=80, seed=9, few_matches_every=2)
    res, keep, _, _ = O.optimize_sim3(pbs, cam, X1, X2, o1, o2, w1, w2)
    assert res[0]["n_inliers"] > 40
    assert res[1]["n_inliers"] == 0 and res[1]["n_corr"] - res[1]["n_bad"] < 10
    assert res[1]["ns"].tobytes() == pbs[1]["ns"].tobytes() and res[1]["scale"] == pbs[1]["scale"]


def test_naive_python_restatement_of_the_first_lm_step():
    """One Gauss-Newton / LM step restated in numpy from the edge outputs: H = sum J^T (rho' w) J over both edge types,
    lambda = 1e-5 max diag; the oracle's first accepted step must reduce the robust chi2 the same way."""
    cam = pinhole_cam()
    pbs, X1, X2, o1, o2, w1, w2, _ = synth.make_sim3_problems(cam, n_candidates=1, n_matches=60, seed=11, outlier_frac=0.0)
    ns, s = pbs[0]["ns"], float(pbs[0]["scale"])
    delta = float(np.sqrt(np.float32(10.0))); dsqr = float(np.float32(delta * delta))
    H = np.zeros((7, 7)); b = np.zeros(7); chi0 = 0.0
    for i in range(60):
        for inv, X, o, w in ((0, X2[i], o1[i], w1[i]), (1, X1[i], o2[i], w2[i])):
            e, Jp, Js = O.edge_sim3(cam, ns, s, X, o, inv)
            J = np.concatenate([Jp, Js[:, None]], 1)
            chi = float(w) * float(e @ e)
            r1 = 1.0 if chi <= dsqr else delta / np.sqrt(chi)
            chi0 += chi if chi <= dsqr else 2 * np.sqrt(chi) * delta - dsqr
            H += r1 * float(w) * J.T @ J
            b += -r1 * float(w) * J.T @ e
    lam = 1e-5 * np.max(np.abs(np.diag(H)))
    x = np.linalg.solve(H + lam * np.eye(7), b)
    ns1 = O.navstate_oplus(ns, 0, x[:6]); s1 = s + x[6]
    chi1 = 0.0
    for i in range(60):
        for inv, X, o, w in ((0, X2[i], o1[i], w1[i]), (1, X1[i], o2[i], w2[i])):
            e, _, _ = O.edge_sim3(cam, ns1, s1, X, o, inv)
            chi = float(w) * float(e @ e)
            chi1 += chi if chi <= dsqr else 2 * np.sqrt(chi) * delta - dsqr
    assert chi1 < 0.5 * chi0
    res, keep, c12, c21 = O.optimize_sim3(pbs, cam, X1, X2, o1, o2, w1, w2)
    assert res[0]["chi2_final"] <= chi1 * 1.0001 and res[0]["n_inliers"] >= 55

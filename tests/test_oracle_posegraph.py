"""Pins of the essential-graph oracle (oracle/posegraph_oracle.cc, Optimizer::OptimizeEssentialGraph src/Optimizer.cc:2309-2688) against
independent computations: scipy's matrix exponential of the similarity generator, numpy matrix algebra, an independent central
difference, numpy's dense normal equations, and drift recovery on a consistent loop."""
import numpy as np
import pytest
import scipy.linalg

import oracle_lib as O
from vieo_slam_b200 import synth


def _mat(S):
    x, y, z, w = S["q"]
    R = synth.R_from_quat(np.array([w, x, y, z]))
    M = np.eye(4)
    M[:3, :3] = S["s"] * R
    M[:3, 3] = S["t"]
    return M


def _gen(u):
    G = np.zeros((4, 4))
    G[:3, :3] = synth.hat(u[:3]) + u[6] * np.eye(3)
    G[:3, 3] = u[3:6]
    return G


@pytest.mark.parametrize("seed", range(6))
def test_exp_log_vs_expm(seed):
    r = np.random.default_rng(seed)
    for scale_w, scale_s in [(1.0, 0.3), (1e-7, 0.3), (1.0, 0.0), (1e-8, 1e-8), (2.5, 0.05)]:
        u = np.concatenate([r.normal(0, 1, 3) * scale_w, r.normal(0, 2, 3), [r.normal() * scale_s]])
        S = O.sim3_exp(u)
        # below 1e-5 the reference takes C = 1 / A = 1/2 / B = 1/6 (sim3.h:80-88): an O(|sigma| |upsilon|) approximation of its own
        approx = 0 < abs(u[6]) < 1e-5
        np.testing.assert_allclose(_mat(S), scipy.linalg.expm(_gen(u)), rtol=0, atol=(abs(u[6]) * 10 if approx else 0) + 2e-12)
        if np.linalg.norm(u[:3]) < 3.0:
            np.testing.assert_allclose(O.sim3_log(S), u, rtol=0, atol=1e-9)


def test_mul_inv_map_vs_matrices():
    r = np.random.default_rng(3)
    for _ in range(20):
        a = O.sim3_exp(r.normal(0, 0.7, 7)); b = O.sim3_exp(r.normal(0, 0.7, 7))
        np.testing.assert_allclose(_mat(O.sim3_mul(a, b)), _mat(a) @ _mat(b), atol=1e-12)
        np.testing.assert_allclose(_mat(O.sim3_inv(a)), np.linalg.inv(_mat(a)), atol=1e-12)


def test_edge_error_and_numeric_jacobian():
    r = np.random.default_rng(5)
    for fix_scale in (False, True):
        v0 = O.sim3_exp(r.normal(0, 0.5, 7)); v1 = O.sim3_exp(r.normal(0, 0.5, 7))
        meas = O.sim3_mul(O.sim3_mul(O.sim3_exp(r.normal(0, 0.02, 7)), v1), O.sim3_inv(v0))
        e, Ji, Jj = O.edge_sim3_graph(meas, v0, v1, fix_scale=fix_scale)
        # error: log of the matrix product, through scipy's logm
        E = scipy.linalg.logm(_mat(meas) @ _mat(v0) @ np.linalg.inv(_mat(v1))).real
        ref = np.array([E[2, 1], E[0, 2], E[1, 0], E[0, 3], E[1, 3], E[2, 3], E[0, 0]])
        np.testing.assert_allclose(e, ref, atol=1e-10)
        # Jacobians: independent central difference with a larger step (the edge is smooth; g2o's 1e-9 step is noisier)
        h = 1e-6
        for side, J in ((0, Ji), (1, Jj)):
            for d in range(7):
                up = np.zeros(7); up[d] = h
                if fix_scale and d == 6:
                    assert np.all(J[:, 6] == 0)
                    continue
                vp = O.sim3_mul(O.sim3_exp(up), v0 if side == 0 else v1); vm = O.sim3_mul(O.sim3_exp(-up), v0 if side == 0 else v1)
                ep = O.edge_sim3_graph(meas, vp if side == 0 else v0, v1 if side == 0 else vp, jac=False)[0]
                em = O.edge_sim3_graph(meas, vm if side == 0 else v0, v1 if side == 0 else vm, jac=False)[0]
                np.testing.assert_allclose(J[:, d], (ep - em) / (2 * h), atol=5e-6)
        # a fixed vertex gets no Jacobian
        _, Ji0, Jj0 = O.edge_sim3_graph(meas, v0, v1, fix0=True)
        assert not Ji0.any() and Jj0.any()


@pytest.mark.parametrize("fix_scale,odom", [(True, 0), (False, 3)])
def test_single_damped_step_vs_numpy(fix_scale, odom):
    pb = synth.make_essential_graph(K=24, seed=2, fix_scale=fix_scale, odom_info_every=odom)
    lam = 1e-3
    out, st, H, b = O.essential_graph(pb, lambda_init=lam, single_step=True, want_system=True)
    n = st["n"]
    assert n == 7 * (24 - 1) and np.allclose(H, H.T)
    # independent assembly from the per-edge errors / Jacobians
    free = [k for k in range(24) if not pb["fixed"][k]]
    col = {k: 7 * i for i, k in enumerate(free)}
    Hn = np.zeros((n, n)); bn = np.zeros(n); chi = 0.0
    for e in range(len(pb["ei"])):
        i, j = int(pb["ei"][e]), int(pb["ej"][e])
        err, Ji, Jj = O.edge_sim3_graph(pb["meas"][e], pb["Scw"][i], pb["Scw"][j], fix0=bool(pb["fixed"][i]), fix1=bool(pb["fixed"][j]),
                                        fix_scale=fix_scale)
        Om = np.eye(7) if pb["info"] is None else pb["info"][e].reshape(7, 7)
        chi += err @ Om @ err
        J = np.zeros((7, n))
        if i in col: J[:, col[i]:col[i] + 7] = Ji
        if j in col: J[:, col[j]:col[j] + 7] = Jj
        Hn += J.T @ Om @ J; bn -= J.T @ Om @ err
    assert abs(st["chi2_initial"] - chi) <= 1e-9 * chi
    np.testing.assert_allclose(H, Hn, rtol=0, atol=1e-9 * np.abs(Hn).max())
    np.testing.assert_allclose(b, bn, rtol=0, atol=1e-9 * np.abs(bn).max())
    x = np.linalg.solve(Hn + lam * np.eye(n), bn)
    for k in free:
        up = x[col[k]:col[k] + 7].copy()
        if fix_scale: up[6] = 0
        want = O.sim3_mul(O.sim3_exp(up), pb["Scw"][k])
        for f in ("q", "t", "s"):
            np.testing.assert_allclose(out[k][f], want[f], atol=1e-8)


@pytest.mark.parametrize("fix_scale", [True, False])
def test_loop_closure_distributes_the_drift(fix_scale):
    pb = synth.make_essential_graph(K=80, seed=4, fix_scale=fix_scale, n_points=50)
    out, st = O.essential_graph(pb)
    assert st["chi2_final"] < 0.05 * st["chi2_initial"] and 2 <= st["iterations"] <= 20
    assert np.array_equal(out[pb["loop"]].tobytes(), pb["Scw"][pb["loop"]].tobytes())  # the fixed vertex is untouched
    # the worst edge inconsistency after the optimisation is far below the loop's initial jump
    def worst(S):
        return max(np.abs(O.edge_sim3_graph(pb["meas"][e], S[pb["ei"][e]], S[pb["ej"][e]], jac=False)[0]).max() for e in range(len(pb["ei"])))
    assert worst(out) < 0.2 * worst(pb["Scw"])
    if fix_scale:
        assert np.allclose(out["s"], pb["Scw"]["s"], atol=1e-12)
    # SE3 recovery and point correction: a point attached to its reference keyframe keeps its camera coordinates up to scale
    T = O.essential_graph_recover_se3(out)
    Pc = O.essential_graph_correct_points(pb["Pw"], pb["ref"], pb["Scw"], out)
    for i in range(len(Pc)):
        k = pb["ref"][i]
        before = _mat(pb["Scw"][k]) @ np.append(pb["Pw"][i].astype(float), 1)
        after = _mat(out[k]) @ np.append(Pc[i].astype(float), 1)
        np.testing.assert_allclose(after[:3], before[:3], atol=2e-5 * max(1, np.abs(before).max()))
        Rk = T[k][:, :3]
        np.testing.assert_allclose(Rk @ Rk.T, np.eye(3), atol=1e-9)
        np.testing.assert_allclose(T[k][:, 3] * out[k]["s"], out[k]["t"], atol=1e-12)


# ---- property tests (hypothesis) of the Sim3 restatement ---------------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_vec7 = st.tuples(*([st.floats(-1.2, 1.2)] * 3 + [st.floats(-5, 5)] * 3 + [st.floats(-0.4, 0.4)]))


@settings(max_examples=150, deadline=None)
@given(_vec7, _vec7, _vec7)
def test_sim3_group_properties(u, v, w):
    a, b, c = O.sim3_exp(np.array(u)), O.sim3_exp(np.array(v)), O.sim3_exp(np.array(w))
    # associativity, inverse, log(exp) on the principal branch, exp(0) = identity
    np.testing.assert_allclose(_mat(O.sim3_mul(O.sim3_mul(a, b), c)), _mat(O.sim3_mul(a, O.sim3_mul(b, c))), atol=1e-9)
    np.testing.assert_allclose(_mat(O.sim3_mul(a, O.sim3_inv(a))), np.eye(4), atol=1e-10)
    uu = np.array(u)
    if 1e-4 < np.linalg.norm(uu[:3]) < 3.0 and abs(uu[6]) > 1e-4:   # away from the reference's own small-angle approximations
        np.testing.assert_allclose(O.sim3_log(a), uu, atol=1e-8)
    ident = O.sim3_exp(np.zeros(7))
    assert ident["s"] == 1.0 and np.array_equal(ident["t"], np.zeros(3)) and np.array_equal(ident["q"], [0, 0, 0, 1])


@settings(max_examples=60, deadline=None)
@given(_vec7, _vec7)
def test_edge_error_vanishes_on_its_own_measurement(u, v):
    v0, v1 = O.sim3_exp(np.array(u)), O.sim3_exp(np.array(v))
    meas = O.sim3_mul(v1, O.sim3_inv(v0))      # Sji = Sjw Siw^-1
    e, _, _ = O.edge_sim3_graph(meas, v0, v1, jac=False)
    assert np.abs(e).max() < 1e-8

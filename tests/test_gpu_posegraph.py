"""Optimizer::OptimizeEssentialGraph on the device (vieo_essential_graph_*, csrc/posegraph.cu) against the oracle restatement of
src/Optimizer.cc:2309-2688 (oracle/posegraph_oracle.cc) through the C ABI with host buffers.  The edge's Jacobians are g2o's central
differences with delta 1e-9: a last-bit difference of an error (CUDA's sin / cos / acos / log vs glibc's) is a 5e-8 difference of a
Jacobian entry, so the system is compared to 1e-6 of its largest entry, the damped step to 1e-6, the converged chi2 to 1e-6 relative
(the north-star tolerance) and the vertices to 1e-6; the fixed vertex and untouched vertices bit for bit."""
import numpy as np
import pytest

import oracle_lib as O
import vieo_slam_b200.api as api
from vieo_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _close_sim3(a, b, atol):
    # q and -q are the same rotation; both sides start from the same bytes and move smoothly, so no sign handling is needed
    assert np.allclose(a["q"], b["q"], atol=atol) and np.allclose(a["t"], b["t"], atol=atol) and np.allclose(a["s"], b["s"], atol=atol)


@pytest.mark.parametrize("fix_scale,odom", [(True, 0), (False, 3), (True, 4)])
def test_single_damped_step(fix_scale, odom):
    pb = synth.make_essential_graph(K=40, seed=6, fix_scale=fix_scale, odom_info_every=odom)
    lam = 1e-3
    out_r, st_r, H_r, b_r = O.essential_graph(pb, lambda_init=lam, single_step=True, want_system=True)
    out_g, st_g, H_g, b_g = api.Optimizer.essential_graph_debug_step(pb, lam)
    assert 7 * int(st_g["n_free"]) == st_r["n"] == H_g.shape[0]
    assert abs(st_g["chi2_initial"] - st_r["chi2_initial"]) <= 1e-9 * st_r["chi2_initial"]
    assert np.allclose(H_g, H_r, rtol=0, atol=1e-6 * np.abs(H_r).max())
    assert np.allclose(b_g, b_r, rtol=0, atol=1e-6 * np.abs(b_r).max())
    _close_sim3(out_g, out_r, 1e-6)
    assert abs(st_g["chi2_final"] - st_r["chi2_final"]) <= 1e-5 * st_r["chi2_initial"]
    assert out_g[pb["loop"]].tobytes() == pb["Scw"][pb["loop"]].tobytes()


@pytest.mark.parametrize("K,fix_scale,odom", [(60, True, 0), (120, False, 5), (400, True, 7), (9, True, 0)])
def test_optimize_matches_oracle(K, fix_scale, odom):
    pb = synth.make_essential_graph(K=K, seed=K, fix_scale=fix_scale, odom_info_every=odom, n_neighbors=min(5, K - 5))
    out_r, st_r = O.essential_graph(pb)
    out_g, T_g, st_g = api.Optimizer.OptimizeEssentialGraph(pb)
    assert st_g["ok"] == 1
    assert abs(st_g["chi2_initial"] - st_r["chi2_initial"]) <= 1e-9 * st_r["chi2_initial"]
    assert abs(st_g["chi2_final"] - st_r["chi2_final"]) <= 1e-6 * st_r["chi2_final"] + 1e-12
    # near the optimum the accept / reject decisions ride on last-bit noise of the numeric Jacobians: the count may differ by a step or two
    assert abs(int(st_g["iterations"]) - st_r["iterations"]) <= 2
    _close_sim3(out_g, out_r, 1e-6)
    assert out_g[pb["loop"]].tobytes() == pb["Scw"][pb["loop"]].tobytes()
    if fix_scale:
        assert np.allclose(out_g["s"], pb["Scw"]["s"], atol=1e-12)
    # SE3 recovery (src/Optimizer.cc:2624-2642) of the device estimate, against the oracle's formula on the same estimate
    assert np.allclose(T_g, O.essential_graph_recover_se3(out_g), atol=1e-12)


def test_g2o_default_lambda_and_fixed_only_edges():
    pb = synth.make_essential_graph(K=30, seed=12, fix_scale=True)
    # a second fixed vertex next to the loop keyframe: the edge between the two is not in the active set (sparse_optimizer.cpp:226)
    pb["fixed"][pb["loop"] + 1] = 1
    out_r, st_r = O.essential_graph(pb, lambda_init=0.0)
    out_g, _, st_g = api.Optimizer.OptimizeEssentialGraph(pb, lambda_init=0.0)
    assert abs(st_g["chi2_initial"] - st_r["chi2_initial"]) <= 1e-9 * st_r["chi2_initial"]
    assert abs(st_g["chi2_final"] - st_r["chi2_final"]) <= 1e-6 * st_r["chi2_final"] + 1e-12
    _close_sim3(out_g, out_r, 1e-6)
    for k in (pb["loop"], pb["loop"] + 1):
        assert out_g[k].tobytes() == pb["Scw"][k].tobytes()


def test_isolated_vertices_and_empty_graph():
    pb = synth.make_essential_graph(K=20, seed=3, fix_scale=True)
    # two extra keyframes without any edge (bad keyframes keep their slot): copied through
    from vieo_slam_b200.layouts import SIM3_DTYPE
    extra = np.zeros(2, SIM3_DTYPE); extra["q"][:, 3] = 1; extra["s"] = 1; extra["t"] = [[1, 2, 3], [4, 5, 6]]
    pb["Scw"] = np.concatenate([pb["Scw"], extra]); pb["fixed"] = np.concatenate([pb["fixed"], [0, 0]]).astype(np.uint8)
    out_r, st_r = O.essential_graph(pb)
    out_g, _, st_g = api.Optimizer.OptimizeEssentialGraph(pb)
    assert int(st_g["n_free"]) == 19 and out_g[20:].tobytes() == extra.tobytes()
    _close_sim3(out_g, out_r, 1e-6)
    # no edges at all: nothing moves
    pb2 = dict(pb, ei=np.zeros(0, np.int32), ej=np.zeros(0, np.int32), meas=np.zeros(0, SIM3_DTYPE), info=None)
    out2, _, st2 = api.Optimizer.OptimizeEssentialGraph(pb2)
    assert out2.tobytes() == pb["Scw"].tobytes() and int(st2["iterations"]) == 0


def test_correct_points_bit_exact():
    pb = synth.make_essential_graph(K=50, seed=9, fix_scale=False, n_points=3000)
    out_g, _, _ = api.Optimizer.OptimizeEssentialGraph(pb)
    got = api.Optimizer.essential_graph_correct_points(pb["Pw"], pb["ref"], pb["Scw"], out_g)
    want = O.essential_graph_correct_points(pb["Pw"], pb["ref"], pb["Scw"], out_g)
    assert got.tobytes() == want.tobytes()  # +, -, *, / in double then one float rounding: no libm call on the path


def test_bad_arguments():
    pb = synth.make_essential_graph(K=12, seed=1, n_neighbors=3)
    bad = dict(pb, ei=pb["ei"].copy()); bad["ei"][0] = 99
    with pytest.raises(Exception):
        api.Optimizer.OptimizeEssentialGraph(bad)
    bad = dict(pb, ej=pb["ei"].copy())
    with pytest.raises(Exception):
        api.Optimizer.OptimizeEssentialGraph(bad)

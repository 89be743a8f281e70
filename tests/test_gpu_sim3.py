"""Optimizer::OptimizeSim3 on the device (vieo_optimize_sim3_batch, csrc/sim3opt.cu) against the oracle restatement of
src/Optimizer.cc:2689-2920: identical kept-match sets and inlier counts, chi2 within 1e-6 relative, the Sim3 vertex within
1e-8 — through the C ABI with host buffers."""
import numpy as np
import pytest

import oracle_lib as O
import vieo_slam_b200.api as api
from vieo_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _cam(kind):
    if kind == "pinhole":  # SetParams(CamInst): Rcb = I, tcb = 0 (src/Optimizer.cc:2802-2806)
        cam = synth.euroc_camera()
        cam["Rcb"] = np.eye(3); cam["tcb"] = 0
        return cam
    if kind == "body":  # a camera with extrinsics Tcr != I (multi-camera rigs, :2807-2809)
        return synth.euroc_camera()
    cam = synth.kb8_camera()
    cam["Rcb"] = np.eye(3); cam["tcb"] = 0
    return cam


def _compare(pbs, cam, arrs):
    got = api.Optimizer.OptimizeSim3Batch(pbs, cam, *arrs)
    ref = O.optimize_sim3(pbs, cam, *arrs)
    (rg, kg, c12g, c21g), (rr, kr, c12r, c21r) = got, ref
    assert np.array_equal(kg, kr), "kept matches differ"
    for name in ("n_inliers", "n_corr", "n_bad", "iterations"):
        assert np.array_equal(rg[name], rr[name]), name
    assert np.allclose(rg["chi2_final"], rr["chi2_final"], rtol=1e-6, atol=1e-9)
    assert np.allclose(c12g, c12r, rtol=1e-6, atol=1e-8) and np.allclose(c21g, c21r, rtol=1e-6, atol=1e-8)
    assert np.allclose(rg["ns"]["p"], rr["ns"]["p"], atol=1e-8) and np.allclose(rg["ns"]["q"], rr["ns"]["q"], atol=1e-8)
    assert np.allclose(rg["scale"], rr["scale"], rtol=1e-9)
    assert np.allclose(rg["lambda_final"], rr["lambda_final"], rtol=1e-5)
    return rg


@pytest.mark.parametrize("kind", ["pinhole", "body", "kb8"])
@pytest.mark.parametrize("fix_scale", [True, False])
def test_optimize_sim3_matches_oracle(kind, fix_scale):
    cam = _cam(kind)
    pbs, X1, X2, o1, o2, w1, w2, _ = synth.make_sim3_problems(cam, n_candidates=6, n_matches=140, seed=21, fix_scale=fix_scale)
    res = _compare(pbs, cam, (X1, X2, o1, o2, w1, w2))
    if kind == "pinhole":
        assert (res["n_inliers"] > 80).all()


def test_few_inliers_exit_and_ragged_batch():
    cam = _cam("pinhole")
    pbs, X1, X2, o1, o2, w1, w2, _ = synth.make_sim3_problems(cam, n_candidates=5, n_matches=60, seed=4, few_matches_every=2)
    res = _compare(pbs, cam, (X1, X2, o1, o2, w1, w2))
    assert res["n_inliers"][1] == 0 and res["ns"][1].tobytes() == pbs["ns"][1].tobytes()


def test_no_outliers_takes_five_more_iterations():
    cam = _cam("pinhole")
    pbs, X1, X2, o1, o2, w1, w2, _ = synth.make_sim3_problems(cam, n_candidates=2, n_matches=50, seed=8, outlier_frac=0.0, th2=60.0)
    res = _compare(pbs, cam, (X1, X2, o1, o2, w1, w2))
    assert (res["n_bad"] == 0).all() and (res["iterations"] <= 10).all()


def test_empty_batch_and_bad_arguments():
    cam = _cam("pinhole")
    from vieo_slam_b200.layouts import SIM3_PROBLEM_DTYPE
    res, keep, _, _ = api.Optimizer.OptimizeSim3Batch(np.zeros(0, SIM3_PROBLEM_DTYPE), cam, np.zeros((0, 3)), np.zeros((0, 3)),
                                                      np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), np.zeros(0))
    assert len(res) == 0
    pbs = np.zeros(1, SIM3_PROBLEM_DTYPE)
    pbs["m_end"] = 5; pbs["scale"] = 1; pbs["th2"] = 10
    with pytest.raises(api.VieoError):
        api.Optimizer.OptimizeSim3Batch(pbs, cam, np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 2)), np.zeros((2, 2)),
                                        np.zeros(2), np.zeros(2))

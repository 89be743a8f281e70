"""GPU parity: bundle-adjustment kernels vs the oracle on the same seeded problems.  Tolerances (north_star): final
chi2 within 1e-6 relative, inlier/outlier sets identical, states within 1e-8 (well under 1 mm)."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seq():
    s = synth.vio_sequence(5, 40)
    s["pre"] = O.imu_preintegrate_frames(s, list(range(40)), O.imu_noise())
    return s


def _cmp_pose(res, outl, chi2, ores, ooutl, ochi2, pbs):
    assert np.array_equal(outl, ooutl), f"outlier sets differ at {np.nonzero(outl != ooutl)[0][:10]}"
    for k in range(len(res)):
        assert res[k]["n_inliers"] == ores[k]["n_inliers"] and res[k]["n_initial"] == ores[k]["n_initial"]
        for f in ("p", "q", "v", "dbg", "dba"):
            assert np.abs(res[k]["cur"][f] - ores[k]["cur"][f]).max() < 1e-8, (k, f)
            assert np.abs(res[k]["last"][f] - ores[k]["last"][f]).max() < 1e-8, (k, "last", f)
        assert abs(res[k]["chi2_final"] - ores[k]["chi2_final"]) <= 1e-6 * max(1.0, abs(ores[k]["chi2_final"]))
        # lambda_final is not compared: once converged, accept/reject is decided by rounding noise in chi2 differences
        # of ~1e-13 and every rejected trial multiplies lambda by 2, 4, 8, ... (optimization_algorithm_levenberg.cpp:143-146)
        M, Mo = res[k]["marg_cov_inv"], ores[k]["marg_cov_inv"]
        assert res[k]["prior_set"] == ores[k]["prior_set"]
        assert np.abs(M - Mo).max() <= 1e-6 * max(np.abs(Mo).max(), 1e-300), (k, np.abs(M - Mo).max(), np.abs(Mo).max())
    sc = np.maximum(np.abs(ochi2), 1.0)
    assert (np.abs(chi2 - ochi2) / sc).max() < 1e-6
    # the LM trajectory (iteration count) may differ by a trial where termination is decided by rounding noise of a
    # converged chi2 (Raul's < 0.1 % criterion, rho == 0): allowed on a small minority of frames only
    assert (res["iterations"] != ores["iterations"]).mean() <= 0.1


@pytest.mark.parametrize("mode,chain,npts", [(1, False, 450), (1, True, 300), (0, False, 400), (1, False, 20)])
def test_pose_optimization_matches_oracle(seq, mode, chain, npts):
    import vieo_slam_b200.api as api
    cam = synth.euroc_camera()
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=npts, seed=10 + mode, mode=mode,
                                                  compute_marg=(mode == 1), chain_prior=chain)
    res, outl, chi2 = api.Optimizer.PoseOptimizationBatch(pbs, cam, X, obs, w, fl)
    ores, ooutl, ochi2 = O.pose_optimization(pbs, cam, X, obs, w, fl)
    _cmp_pose(res, outl, chi2, ores, ooutl, ochi2, pbs)


def test_pose_optimization_edge_cases(seq):
    import vieo_slam_b200.api as api
    cam = synth.euroc_camera()
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=40, seed=21)
    pbs = pbs[:6].copy()
    pbs[0]["edge_end"] = pbs[0]["edge_begin"] + 2      # < 3 correspondences: returns 0 untouched
    pbs[1]["edge_end"] = pbs[1]["edge_begin"] + 5      # < 10 edges in total: a single round
    pbs[2]["preint"]["dt"] = 0                          # no IMU edge: estimate reset every round
    pbs[3]["edge_end"] = pbs[3]["edge_begin"] + 2
    pbs[3]["no_mps"] = 1                                # odom-only tracking keeps going with < 3 points
    pbs[4]["edge_end"] = pbs[4]["edge_begin"] + 25      # < 30 inliers: rescue pass
    res, outl, chi2 = api.Optimizer.PoseOptimizationBatch(pbs, cam, X, obs, w, fl)
    ores, ooutl, ochi2 = O.pose_optimization(pbs, cam, X, obs, w, fl)
    _cmp_pose(res, outl, chi2, ores, ooutl, ochi2, pbs)
    assert res[0]["n_inliers"] == 0 and res[0]["iterations"] == 0
    assert res[0]["cur"].tobytes() == pbs[0]["cur"].tobytes()


def test_pose_optimization_deterministic(seq):
    import vieo_slam_b200.api as api
    cam = synth.euroc_camera()
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=500, seed=33)
    a = api.Optimizer.PoseOptimizationBatch(pbs, cam, X, obs, w, fl)
    b = api.Optimizer.PoseOptimizationBatch(pbs, cam, X, obs, w, fl)
    assert a[0].tobytes() == b[0].tobytes() and a[2].tobytes() == b[2].tobytes()

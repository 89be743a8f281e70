"""CPU: host-side logic of the SURVEY 8(f) entry points that runs before any device call — argument and range checks,
capacity errors, the empty-batch early returns — and the loud failure (no CPU fallback) when no B200 is present."""
import numpy as np
import pytest

from vieo_slam_b200 import synth


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


def test_empty_batches_return_without_a_device():
    import vieo_slam_b200.api as api
    L = api.lib()
    # n_frames == 0 / n_points == 0 / n_kf == 0: VIEO_OK before use_device()
    assert L.vieo_frustum_batch(None, 0, *([None] * 11), 0) == 0
    assert L.vieo_search_local_points(None, None, 0, *([None] * 21), 0) == 0
    assert L.vieo_proj_search_batch(None, 0, *([None] * 12), 0) == 0
    assert L.vieo_distinctive_descriptors(None, 0, None, None, 0, None, None, 0) == 0
    bg = np.zeros(3)
    neq = np.zeros(1, np.int32)
    assert L.vieo_imu_init_gyro_bias(None, None, 0, 1, bg.ctypes.data, neq.ctypes.data, None, None, None, None, None, None, 0) == 0
    assert neq[0] == 0


def test_argument_errors_are_reported_before_the_device_is_touched():
    import vieo_slam_b200.api as api
    m = api.ORBmatcher()
    # frustum and search frames must describe the same query range
    pb = synth.make_frustum_problem(5, n_frames=2, n_kp=300, n_q=200)
    pb["frustum"][1]["n_q"] -= 1
    with pytest.raises(api.VieoError, match="share the query range"):
        m.SearchLocalPoints(pb)
    # pyramid depth mismatch between the two records
    pb = synth.make_frustum_problem(5, n_frames=1, n_kp=300, n_q=200)
    pb["frames"]["n_levels"] = 4
    with pytest.raises(api.VieoError, match="pyramid depth"):
        m.SearchLocalPoints(pb)
    # keypoint capacity of the shared-memory grid
    big = synth.make_fuse_problem(6, n_frames=1, n_kp=4200, n_q=20)
    with pytest.raises(api.VieoError, match="max 4096"):
        m.SearchByProjectionBase(big)
    # a non-positive log scale factor cannot define PredictScale
    pb = synth.make_frustum_problem(5, n_frames=1, n_kp=300, n_q=200)
    pb["frustum"]["log_scale_factor"] = 0.0
    with pytest.raises(api.VieoError, match="log scale factor"):
        api.isInFrustum(pb)
    # distinctive descriptors: rows outside the pool, lists that do not start at 0 or descend
    d = synth.make_distinctive_problem(7, n_points=20, max_obs=5)
    bad = d["rows"].copy()
    bad[0] = len(d["pool"])
    with pytest.raises(api.VieoError, match="out of range"):
        m.ComputeDistinctiveDescriptors(d["pool"], d["ptr"], bad)
    with pytest.raises(api.VieoError, match="ascending|observation lists"):
        m.ComputeDistinctiveDescriptors(d["pool"], [0, 4, 2], d["rows"])
    with pytest.raises(api.VieoError, match="observation lists"):
        m.ComputeDistinctiveDescriptors(d["pool"], [1, 4], d["rows"])


def test_no_cpu_fallback_in_the_new_entry_points():
    if not _no_gpu():
        pytest.skip("GPU present")
    import vieo_slam_b200.api as api
    m = api.ORBmatcher()
    with pytest.raises(api.VieoError, match="CUDA|device"):
        api.isInFrustum(synth.make_frustum_problem(8, n_frames=1, n_kp=200, n_q=100))
    with pytest.raises(api.VieoError, match="CUDA|device"):
        m.SearchLocalPoints(synth.make_frustum_problem(8, n_frames=1, n_kp=200, n_q=100))
    with pytest.raises(api.VieoError, match="CUDA|device"):
        m.SearchByProjectionBase(synth.make_fuse_problem(8, n_frames=1, n_kp=200, n_q=100))
    d = synth.make_distinctive_problem(9, n_points=10, max_obs=4)
    with pytest.raises(api.VieoError, match="CUDA|device"):
        m.ComputeDistinctiveDescriptors(d["pool"], d["ptr"], d["rows"])
    pre = np.zeros(3, api.PREINT_DTYPE)
    with pytest.raises(api.VieoError, match="CUDA|device"):
        api.IMUPreintegrator().OptimizeInitialGyroBias(pre, np.tile(np.eye(3), (3, 1, 1)), np.zeros(3))

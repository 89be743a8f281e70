"""CPU: pin the IMU pre-integration oracle with closed-form answers, finite differences and invariants
(the reference has no tests for this path and Eigen/Sophus are absent: SURVEY.md §8c)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_lib as O


def _stream(n, dt, a, w, t0=0.0):
    s = np.zeros((n, 7))
    s[:, 0] = t0 + dt * np.arange(n)
    s[:, 1:4] = a
    s[:, 4:7] = w
    return s


def test_constant_acceleration_no_rotation_closed_form():
    nz = O.imu_noise()
    a = np.array([0.3, -1.2, 9.7])
    s = _stream(41, 0.005, a, 0.0)
    T = 0.2
    p = O.imu_preintegrate(s, 0.0, T, np.zeros(3), np.zeros(3), nz)
    assert p["status"] == 0 and abs(p["dt"] - T) < 1e-15
    assert np.allclose(p["Rij"], np.eye(3), atol=1e-15)
    assert np.allclose(p["vij"], a * T, rtol=1e-13)
    assert np.allclose(p["pij"], a * T * T / 2, rtol=1e-13)
    assert np.allclose(p["Jav"], -T * np.eye(3), rtol=1e-13) and np.allclose(p["Jap"], -T * T / 2 * np.eye(3), rtol=1e-12)
    assert np.allclose(p["JgR"], -T * np.eye(3), rtol=1e-12)
    # bias subtraction: integrating (a + ba) with bias ba gives the same result
    q = O.imu_preintegrate(_stream(41, 0.005, a + 0.1, 0.02), 0.0, T, np.full(3, 0.02), np.full(3, 0.1), nz)
    assert np.allclose(q["vij"], p["vij"], rtol=1e-12) and np.allclose(q["Rij"], np.eye(3), atol=1e-14)


def test_constant_rotation_matches_exponential():
    nz = O.imu_noise()
    w = np.array([0.4, -0.7, 1.1])
    s = _stream(201, 0.005, 0.0, w)
    p = O.imu_preintegrate(s, 0.0, 1.0, np.zeros(3), np.zeros(3), nz)
    assert np.allclose(p["Rij"], Rotation.from_rotvec(w * 1.0).as_matrix(), atol=1e-12)
    assert np.allclose(p["Rij"] @ p["Rij"].T, np.eye(3), atol=1e-14)


def test_covariance_orderings_are_permutations_and_psd():
    nz = O.imu_noise()
    rng = np.random.default_rng(2)
    s = _stream(30, 0.005, 0.0, 0.0)
    s[:, 1:4] = rng.normal(0, 2, (30, 3)) + [0, 0, 9.8]
    s[:, 4:7] = rng.normal(0, 0.5, (30, 3))
    p = O.imu_preintegrate(s, 0.002, 0.141, [0.01, -0.02, 0.005], [0.1, 0.05, -0.2], nz)
    perm = np.r_[0:3, 6:9, 3:6]  # P,R,V -> P,V,R
    assert np.allclose(p["SigmaPVR"], p["SigmaPRV"][np.ix_(perm, perm)], rtol=1e-10, atol=1e-22)
    assert np.allclose(p["SigmaPRV"], p["SigmaPRV"].T, rtol=1e-9, atol=1e-22)
    assert np.linalg.eigvalsh((p["SigmaPRV"] + p["SigmaPRV"].T) / 2).min() > 0
    # fixed-noise model (EuRoC: sigma^2 * 200 Hz per step): rotation block after n steps of dt ~ n * sg * dt^2 * I
    n_steps = 28.0
    assert np.allclose(np.diag(p["SigmaPRV"])[3:6], (1.6968e-4 ** 2 * 200) * 0.005 ** 2 * n_steps, rtol=0.1)


def test_bias_jacobians_by_finite_differences():
    nz = O.imu_noise()
    rng = np.random.default_rng(4)
    s = _stream(25, 0.005, 0.0, 0.0)
    s[:, 1:4] = rng.normal(0, 1.5, (25, 3)) + [0, 0, 9.8]
    s[:, 4:7] = rng.normal(0, 0.4, (25, 3))
    bg, ba = np.array([0.01, 0.02, -0.01]), np.array([0.05, -0.1, 0.2])
    p0 = O.imu_preintegrate(s, 0.0, 0.12, bg, ba, nz)
    eps = 1e-6
    for k in range(3):
        d = np.zeros(3); d[k] = eps
        pg = O.imu_preintegrate(s, 0.0, 0.12, bg + d, ba, nz)
        pa = O.imu_preintegrate(s, 0.0, 0.12, bg, ba + d, nz)
        assert np.allclose((pg["pij"] - p0["pij"]) / eps, p0["Jgp"][:, k], atol=2e-6)
        assert np.allclose((pg["vij"] - p0["vij"]) / eps, p0["Jgv"][:, k], atol=2e-5)
        assert np.allclose((pa["pij"] - p0["pij"]) / eps, p0["Jap"][:, k], atol=1e-8)
        assert np.allclose((pa["vij"] - p0["vij"]) / eps, p0["Jav"][:, k], atol=1e-8)
        dphi = Rotation.from_matrix(p0["Rij"].T @ pg["Rij"]).as_rotvec() / eps
        assert np.allclose(dphi, p0["JgR"][:, k], atol=2e-6)


def test_boundary_handling():
    nz = O.imu_noise()
    a, w = np.array([0.0, 0.0, 1.0]), np.array([0.0, 0.0, 0.0])
    s = _stream(21, 0.005, a, w, t0=1.0)  # samples at 1.000 .. 1.100
    for ti, tj in ((1.0, 1.1), (1.0012, 1.0987), (0.9991, 1.1), (1.0, 1.1034), (1.02, 1.02 + 1e-9)):
        p = O.imu_preintegrate(s, ti, tj, np.zeros(3), np.zeros(3), nz)
        assert p["status"] == 0
        assert abs(p["dt"] - (tj - ti)) < 1e-12, (ti, tj, p["dt"])
        assert np.allclose(p["vij"], a * (tj - ti), rtol=1e-9)
    # linear-in-time acceleration: mid-point rule with end interpolation is exact for v
    s2 = s.copy(); s2[:, 3] = 10 * (s2[:, 0] - 1.0)
    p = O.imu_preintegrate(s2, 1.0012, 1.0987, np.zeros(3), np.zeros(3), nz)
    assert abs(p["vij"][2] - 5 * (0.0987 ** 2 - 0.0012 ** 2)) < 1e-12
    # duplicated timestamp (dt == 0 is skipped), gap > 1.5 s fails with delta-t 0, empty list is a no-op
    s3 = np.vstack([s[:5], s[4:]])
    assert abs(O.imu_preintegrate(s3, 1.0, 1.1, np.zeros(3), np.zeros(3), nz)["dt"] - 0.1) < 1e-12
    s4 = s.copy(); s4[10:, 0] += 2.0
    bad = O.imu_preintegrate(s4, 1.0, 3.1, np.zeros(3), np.zeros(3), nz)
    assert bad["status"] == -1 and bad["dt"] == 0
    e = O.imu_preintegrate(np.zeros((0, 7)), 1.0, 1.1, np.zeros(3), np.zeros(3), nz)
    assert e["status"] == 0 and e["dt"] == 0 and np.array_equal(e["Rij"], np.eye(3))
    # reversed time (map reuse, OdomPreIntegrator.h:240-262): negative delta-t, velocity integrates backwards
    r = O.imu_preintegrate(s, 1.1, 1.0, np.zeros(3), np.zeros(3), nz)
    assert r["status"] == 0 and abs(r["dt"] + 0.1) < 1e-12 and np.allclose(r["vij"], -a * 0.1, rtol=1e-9)


def test_noise_models():
    s = _stream(21, 0.005, [0, 0, 9.8], [0.1, 0.0, 0.0])
    fixed = O.imu_preintegrate(s, 0.0, 0.1, np.zeros(3), np.zeros(3), O.imu_noise(dt_cov_noise_fixed=1, freq_ref=200.0))
    cont = O.imu_preintegrate(s, 0.0, 0.1, np.zeros(3), np.zeros(3), O.imu_noise(dt_cov_noise_fixed=0, freq_ref=0.0))
    # sigma^2/dt with dt = 1/200 equals sigma^2 * 200
    assert np.allclose(fixed["SigmaPRV"], cont["SigmaPRV"], rtol=1e-9, atol=1e-24)

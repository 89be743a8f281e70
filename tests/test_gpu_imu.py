"""GPU parity: batched IMU pre-integration kernel vs the oracle (fp64; tolerance 1e-12 relative — only the
device sin/cos/atan differ from glibc's by an ulp)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

FIELDS = ("Rij", "vij", "pij", "SigmaPRV", "SigmaPVR", "Jgp", "Jap", "Jgv", "Jav", "JgR", "dt")


def _close(a, b, name):
    # The reference divides the noise by dt in the non-fixed branch and calls update() with dt == 0 when tj lies
    # beyond the last sample (OdomPreIntegrator.h:403-422, 449-451): 0*inf = NaN there is reference behaviour, so
    # the NaN pattern must match and the finite entries are compared.
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), (name, "NaN pattern differs")
    fin = ~np.isnan(b)
    if not fin.any():
        return
    scale = max(np.abs(b[fin]).max(), 1e-300)
    assert np.abs(a[fin] - b[fin]).max() <= 1e-12 * scale + 1e-300, (name, np.abs(a[fin] - b[fin]).max(), scale)


def _imu_stream(rng, n, rate=200.0, t0=0.0):
    s = np.zeros((n, 7))
    s[:, 0] = t0 + np.arange(n) / rate + rng.normal(0, 1e-5, n)
    s[:, 0].sort()
    s[:, 1:4] = rng.normal(0, 2.0, (n, 3)) + [0, 0, 9.81]
    s[:, 4:7] = rng.normal(0, 0.6, (n, 3))
    return s


@pytest.mark.parametrize("fixed,freq", [(1, 200.0), (0, 0.0), (0, 200.0)])
def test_batch_matches_oracle(fixed, freq):
    import vieo_slam_b200.api as api
    rng = np.random.default_rng(10 + fixed)
    pre = api.IMUPreintegrator(dt_cov_noise_fixed=fixed, freq_ref=freq)
    nz = O.imu_noise(dt_cov_noise_fixed=fixed, freq_ref=freq)
    samples, seg, tt, bb = [], [0], [], []
    cases = [10, 11, 12, 600, 1, 2, 0, 40, 40, 40, 40, 7]
    for k, n in enumerate(cases):
        s = _imu_stream(rng, n, t0=5.0 * k)
        if n >= 2:
            lo, hi = s[0, 0], s[-1, 0]
            ti, tj = [(lo, hi), (lo + 0.0013, hi - 0.0021), (lo - 0.002, hi + 0.003), (lo + 0.001, hi), (lo, hi)][k % 5]
            if k == 8:
                ti, tj = tj, ti  # reversed time
            if k == 9:
                s[20:, 0] += 2.0  # gap > 1.5 s -> status -1
                tj = s[-1, 0]
            if k == 10:
                s[15] = s[14]  # duplicated sample: dt == 0 skipped
        else:
            ti, tj = 5.0 * k, 5.0 * k + 0.05
        samples.append(s); seg.append(seg[-1] + n); tt.append((ti, tj))
        bb.append(np.r_[rng.normal(0, 0.02, 3), rng.normal(0, 0.1, 3)])
    smp = np.vstack(samples)
    got = pre.preintegrate_batch(smp, seg, tt, bb)
    for k in range(len(cases)):
        ref = O.imu_preintegrate(smp[seg[k]:seg[k + 1]], tt[k][0], tt[k][1], bb[k][:3], bb[k][3:], nz)
        assert got[k]["status"] == ref["status"], k
        for f in FIELDS:
            _close(got[k][f], ref[f], (k, f))
    assert got[9]["status"] == -1 and got[9]["dt"] == 0
    assert got[6]["dt"] == 0 and np.array_equal(got[6]["Rij"], np.eye(3))  # empty list: untouched identity state
    assert got[8]["dt"] < 0


def test_single_interval_surface():
    import vieo_slam_b200.api as api
    pre = api.IMUPreintegrator()
    s = np.zeros((41, 7)); s[:, 0] = 0.005 * np.arange(41); s[:, 1:4] = [0.3, -1.2, 9.7]
    p = pre.PreIntegration(s, 0.0, 0.2, np.zeros(3), np.zeros(3))
    assert p["status"] == 0 and np.allclose(p["vij"], np.array([0.3, -1.2, 9.7]) * 0.2, rtol=1e-13)

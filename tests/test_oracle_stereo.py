"""CPU: Frame::ComputeStereoMatches restatement (oracle/match_oracle.cc) on synthetic rectified stereo with a known
integer disparity per frame (the right view is the left one shifted, synth.stereo_stream)."""
import numpy as np

import oracle_lib as O
from vieo_slam_b200.synth import EUROC, stereo_stream


def test_stereo_matches_recover_the_disparity():
    imgs = stereo_stream(3, 77).reshape(3, 2, 480, 752)
    bf = np.float32(EUROC["bf"]); minZ = np.float32(bf / np.float32(EUROC["fx"]))
    for f in range(3):
        oL, oR = O.OrbOracle(1200, 1.2, 8, 20, 7), O.OrbOracle(1200, 1.2, 8, 20, 7)
        nl, kl, dl, _ = oL.extract(imgs[f, 0]); nr, kr, dr, _ = oR.extract(imgs[f, 1])
        ur, dp, sad, kept = O.stereo_matches(oL, kl, dl, oR, kr, dr, bf, minZ)
        ok = ur >= 0
        assert kept == ok.sum() and kept > 0.3 * nl
        disp = kl["x"][ok] - ur[ok]
        d_true = np.median(disp)
        assert abs(d_true - round(float(d_true))) < 0.05 and d_true > 5      # the generator's integer shift
        assert (np.abs(disp - d_true) < 1.5).mean() > 0.95                    # sub-pixel refinement lands on it
        assert np.allclose(dp[ok], bf / disp, rtol=1e-6)
        assert np.all(dp[~ok] == -1) and np.all(ur[~ok] == -1)
        # the median filter: every kept match has SAD < 2.1 * median of the accepted ones
        acc = sad >= 0
        med = np.sort(sad[acc])[acc.sum() // 2]
        assert np.all(sad[ok] < np.float32(1.5) * np.float32(1.4) * np.float32(med))
        assert np.all(sad[acc & ~ok] >= np.float32(1.5) * np.float32(1.4) * np.float32(med))

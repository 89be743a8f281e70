"""GPU parity: the guided-search kernel (ORBmatcher::SearchByProjection, both tracking overloads) against the oracle,
through the C ABI.  Integer / index work: bit-exact (keypoint owner per keypoint, keypoint and distance per query,
match count per frame)."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import SBP_LAST_FRAME, SBP_LOCAL_MAP

pytestmark = pytest.mark.gpu

CASES = [dict(mode=SBP_LAST_FRAME, th=15.0, motion="still"), dict(mode=SBP_LAST_FRAME, th=7.0, motion="forward"),
         dict(mode=SBP_LAST_FRAME, th=30.0, motion="backward", th_far=9.0), dict(mode=SBP_LAST_FRAME, th=15.0, mono=True),
         dict(mode=SBP_LOCAL_MAP, th=1.0, blocked_frac=0.3, n_q=2500), dict(mode=SBP_LOCAL_MAP, th=3.0, blocked_frac=0.1, th_far=8.0),
         # > 64 candidates per query: the claim pass re-enumerates instead of reading the stored list
         dict(mode=SBP_LAST_FRAME, th=30.0, cluster=True, n_kp=1500, n_q=600),
         dict(mode=SBP_LOCAL_MAP, th=6.0, cluster=True, n_kp=1500, n_q=600)]


def _run(pb, nnratio=0.8, check_ori=True):
    import vieo_slam_b200.api as api
    pb["frames"]["nn_ratio"] = nnratio
    pb["frames"]["check_orientation"] = int(check_ori)
    out = api.ORBmatcher(nnratio, check_ori).SearchByProjection(pb)
    ref = O.search_by_projection(pb)
    for a, b, name in zip(out, ref, ("kp_match", "q_match", "q_dist", "n_matches")):
        assert np.array_equal(a, b), (name, np.nonzero(a != b)[0][:10])
    return out


@pytest.mark.parametrize("kw", CASES)
def test_search_by_projection_matches_oracle(kw):
    pb = synth.make_sbp_problem(77, n_frames=6, **kw)
    kp_match, q_match, q_dist, nm = _run(pb)
    assert nm.min() > 50
    if kw.get("cluster"):
        assert (q_match >= 0).sum() > 100


def test_search_by_projection_edge_cases():
    import vieo_slam_b200.api as api
    pb = synth.make_sbp_problem(5, n_frames=3, mode=SBP_LAST_FRAME)
    _run(pb, check_ori=False)
    # an empty frame in the middle of a batch, and a frame without queries
    fr = pb["frames"]
    fr[1]["n_kp"] = 0
    fr[2]["n_q"] = 0
    out = _run(pb)
    assert out[3][1] == 0 and out[3][2] == 0
    # keypoint capacity and bad mode are reported, not silently truncated
    big = synth.make_sbp_problem(6, n_frames=1, n_kp=4200, n_q=50)
    with pytest.raises(api.VieoError):
        api.ORBmatcher().SearchByProjection(big)
    pb["mode"] = 7
    with pytest.raises(api.VieoError):
        api.ORBmatcher().SearchByProjection(pb)


def test_search_by_projection_deterministic():
    pb = synth.make_sbp_problem(9, n_frames=8, mode=SBP_LOCAL_MAP, th=3.0, n_q=2000, blocked_frac=0.2)
    a = _run(pb)
    b = _run(pb)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


# ---- the relocalisation search: ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist, th_far_pts) ----
RELOC_CASES = [dict(th=10.0, orb_dist=100), dict(th=3.0, orb_dist=64, blocked_frac=0.3), dict(th=15.0, orb_dist=100, th_far=7.0),
               dict(th=30.0, orb_dist=100, cluster=True, n_kp=1500, n_q=500)]   # > 64 candidates: the re-enumeration path


@pytest.mark.parametrize("kw", RELOC_CASES)
@pytest.mark.parametrize("check_ori", [True, False])
def test_reloc_search_matches_oracle(kw, check_ori):
    import vieo_slam_b200.api as api
    pb = synth.make_reloc_problem(31, n_frames=5, **kw)
    pb["frames"]["check_orientation"] = int(check_ori)
    out = api.ORBmatcher(0.9, check_ori).SearchByProjectionReloc(pb)
    ref = O.sbp_reloc(pb)
    for a, b, name in zip(out, ref, ("kp_match", "q_match", "q_dist", "q_level", "n_matches")):
        assert np.array_equal(a, b), (name, np.nonzero(a != b)[0][:10])
    assert out[4].min() > 30 and (out[3] == -1).sum() > 10


def test_reloc_search_edge_cases():
    import vieo_slam_b200.api as api
    pb = synth.make_reloc_problem(32, n_frames=3)
    pb["frames"][1]["n_q"] = 0
    pb["frames"][2]["n_kp"] = 0
    out = api.ORBmatcher(0.9, True).SearchByProjectionReloc(pb)
    ref = O.sbp_reloc(pb)
    assert all(np.array_equal(a, b) for a, b in zip(out, ref)) and out[4][1] == 0 and out[4][2] == 0
    pb["reloc"]["orb_dist"] = 256
    with pytest.raises(api.VieoError):
        api.ORBmatcher().SearchByProjectionReloc(pb)

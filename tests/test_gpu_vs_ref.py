"""GPU parity against the REFERENCE ITSELF: the CUDA path (through the C ABI) compared directly with the reference's own functions
compiled unchanged in oracle/_ref/libref.so (built in the container from /root/reference by oracle/ref_build/Makefile and shipped
prebuilt; nothing here reads /root/reference).  The oracle is not in the loop: these are the same comparisons as
tests/test_oracle_ref.py with the device in the oracle's place.  Integer / index work and float outputs: bit-exact."""
import numpy as np
import pytest

import ref_lib as R
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import SBP_LAST_FRAME, SBP_LOCAL_MAP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    if not R.available():
        pytest.skip("oracle/_ref/libref.so not present")
    import vieo_slam_b200.api as api
    return api


@pytest.mark.parametrize("kw", [dict(seed=5, th=15.0), dict(seed=6, th=7.0, mono=True), dict(seed=7, th=15.0, motion="forward"),
                                dict(seed=8, th=15.0, motion="backward"), dict(seed=9, th=30.0, cluster=True, blocked_frac=0.3),
                                dict(seed=10, th=15.0, th_far=6.0)])
def test_search_by_projection_last_frame(api, kw):
    """ORBmatcher::SearchByProjection(Frame&, const Frame& last, ...) (src/ORBmatcher.cc:1303-1467): keypoint -> map point and counts"""
    import inspect
    args = {k: v for k, v in kw.items() if k in inspect.signature(synth.make_sbp_problem).parameters and k != "seed"}
    pb = synth.make_sbp_problem(kw["seed"], 3, mode=SBP_LAST_FRAME, **args)
    pb["q_Xw"] = np.ascontiguousarray(pb["q_Xw"].astype(np.float32).astype(np.float64))   # MapPoint positions are float in the reference
    for chk in (1, 0):
        pb["frames"]["check_orientation"] = chk
        kp_d, _, _, n_d = api.ORBmatcher(float(pb["frames"]["nn_ratio"][0]), bool(chk)).SearchByProjection(pb)
        kp_r, n_r = R.search_by_projection_last_frame(pb)
        assert np.array_equal(n_d, n_r), (n_d, n_r)
        assert np.array_equal(kp_d, kp_r)
        assert n_d.sum() > 100


@pytest.mark.parametrize("kw", [dict(seed=15, th=1.0), dict(seed=17, th=1.0, cluster=True, blocked_frac=0.3), dict(seed=18, th=5.0, th_far=6.0),
                                dict(seed=19, th=1.0, mono=True)])
def test_search_by_projection_local_map(api, kw):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) + RadiusByViewingCos (src/ORBmatcher.cc:230-342)"""
    pb = synth.make_sbp_problem(kw["seed"], 3, mode=SBP_LOCAL_MAP, n_q=1800, **{k: v for k, v in kw.items() if k != "seed"})
    for ratio in (0.8, 0.6):
        pb["frames"]["nn_ratio"] = ratio
        kp_d, _, _, n_d = api.ORBmatcher(ratio, bool(pb["frames"]["check_orientation"][0])).SearchByProjection(pb)
        kp_r, n_r = R.search_by_projection_local_map(pb)
        assert np.array_equal(n_d, n_r), (n_d, n_r)
        assert np.array_equal(kp_d, kp_r)
        assert n_d.sum() > 100


@pytest.mark.parametrize("kw", [dict(seed=51), dict(seed=52, th=15.0, orb_dist=64), dict(seed=53, th_far=8.0, cluster=True),
                                dict(seed=54, blocked_frac=0.4, orb_dist=80)])
def test_search_by_projection_reloc(api, kw):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist, th_far_pts) (src/ORBmatcher.cc:1471-1606)"""
    pb = synth.make_reloc_problem(kw["seed"], n_frames=3, **{k: v for k, v in kw.items() if k != "seed"})
    for chk in (0, 1):
        pb["frames"]["check_orientation"] = chk
        out = api.ORBmatcher(0.9, bool(chk)).SearchByProjectionReloc(pb)
        kp_r, n_r = R.sbp_reloc(pb)
        assert np.array_equal(out[4], n_r), (out[4], n_r)
        assert np.array_equal(out[0], kp_r)
        assert n_r.sum() > 150


@pytest.mark.parametrize("kw", [dict(), dict(use_bf=False), dict(check_viewing_angle=False, th_radius=4.0), dict(cluster=True, skip_frac=0.3)])
def test_search_by_projection_base(api, kw):
    """ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227), the search half of Fuse / SearchBySim3: keypoint and distance"""
    pb = synth.make_fuse_problem(61, **kw)
    bd, dd, _ = api.ORBmatcher().SearchByProjectionBase(pb)
    br, dr = R.proj_search(pb)
    assert (bd >= 0).sum() > 1500
    assert np.array_equal(bd, br)
    assert np.array_equal(dd[bd >= 0], dr[bd >= 0])


@pytest.mark.parametrize("seed", [3, 4])
def test_search_by_bow(api, seed):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:344-505)"""
    for ratio in (0.7, 0.9):
        pb = synth.make_bow_problem(seed, n_pairs=4, n_kp=1000, n_nodes=70, nn_ratio=ratio)
        for chk in (0, 1):
            pb["pairs"]["check_orientation"][:] = chk
            mf, nm = api.search_by_bow(pb)
            for p in range(4):
                mr, nr = R.search_by_bow(pb, p)
                ob = int(pb["pairs"][p]["out_begin"])
                assert nm[p] == nr and np.array_equal(mf[ob:ob + int(pb["pairs"][p]["n_kp2"])], mr), (seed, ratio, chk, p, nm[p], nr)
        assert nm.sum() > 100


@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("n_cams", [1, 2, 4])
def test_is_in_frustum_rig(api, model, n_cams):
    """Frame::isInFrustum (src/Frame.cc:335-416) for single cameras and rigs, K multiply / pinhole / KB8 projection: every output"""
    pb = synth.make_frustum_rig_problem(40 + model, n_frames=3, n_q=1500, n_cams=n_cams, model=model)
    got, ref = api.frustum_rig_batch(pb), R.is_in_frustum_rig(pb)
    assert ref["n_inview"].sum() > 1500 and np.array_equal(got["n_inview"], ref["n_inview"])
    for k in ("inview", "cam_mask", "level", "proj", "viewcos", "depth"):
        assert np.asarray(got[k]).tobytes() == np.asarray(ref[k]).tobytes(), (k, int((got[k] != ref[k]).sum()))


# ---- added after the round's GPU budget was spent: run so far with the oracle standing in for the device search (kept last) ----
@pytest.mark.parametrize("kw", [dict(), dict(cluster=True, skip_frac=0.3), dict(th_radius=4.0)])
def test_fuse(api, kw):
    """ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:1152-1165): the keypoint each map point is fused into
    (what KeyFrame::FuseMP receives) and nFused"""
    pb = synth.make_fuse_problem(71, **kw)
    hit, nf = api.ORBmatcher().Fuse(pb)
    hr, nr = R.fuse(pb)
    assert np.array_equal(hit, hr) and np.array_equal(nf, nr)
    assert nr.sum() > 500


@pytest.mark.parametrize("seed,th", [(3, 7.5), (5, 4.0)])
def test_search_by_sim3(api, seed, th):
    """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1222-1302): vpMatches12 and nFound (one two-frame device search + the host agreement)"""
    import sim3_search_data as D
    s1, s2, sim3, prior = D.make(seed, th=th)
    m12, n, p21, p12 = R.search_by_sim3(s1, s2, sim3, th, prior)
    got, nf = api.ORBmatcher().SearchBySim3(D.flat_problem(s1, s2, p21, p12, prior, th))
    assert nf == n and np.array_equal(got, m12)
    assert n > 200

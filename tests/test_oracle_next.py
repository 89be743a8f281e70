"""CPU: pins for the oracle restatements of the SURVEY.md 8(f) rows — Frame::isInFrustum + MapPoint::PredictScale,
MapPoint::ComputeDistinctiveDescriptors and Optimizer::OptimizeInitialGyroBias — by independent, deliberately naive
restatements written from the reference lines, known answers, and (for the level table the CUDA kernel uses instead of a
device logf) equivalence of vieo_frustum_level_table with the formula over float neighbourhoods of every boundary.  The
reference's own tests hold no vectors for these functions ("parity unpinned" for last-bit float rounding)."""
import numpy as np

import oracle_lib as O
from vieo_slam_b200 import synth

f32 = np.float32


def _sum3(a0, a1, a2):  # Eigen 3.3 redux_novec_unroller<.., 0, 3>: a0 + (a1 + a2)
    return f32(a0 + f32(a1 + a2))


def _naive_frustum(G, wP, Pn, mx, mn):
    """src/Frame.cc:335-416 line by line in numpy float32 scalars (one camera, no distortion)."""
    R = G["Rcw"].reshape(3, 3)
    Pc = [f32(_sum3(f32(R[r, 0] * wP[0]), f32(R[r, 1] * wP[1]), f32(R[r, 2] * wP[2])) + G["tcw"][r]) for r in range(3)]
    if Pc[2] < f32(0):
        return None
    with np.errstate(all="ignore"):
        invz = f32(f32(1) / Pc[2])
        xn, yn = f32(Pc[0] * invz), f32(Pc[1] * invz)
        u = _sum3(f32(G["fx"] * xn), f32(f32(0) * yn), G["cx"])
        v = _sum3(f32(f32(0) * xn), f32(G["fy"] * yn), G["cy"])
    if u < G["minx"] or u > G["maxx"] or v < G["miny"] or v > G["maxy"]:
        return None
    PO = [f32(wP[k] - G["Ow"][k]) for k in range(3)]
    dist = f32(np.sqrt(_sum3(f32(PO[0] * PO[0]), f32(PO[1] * PO[1]), f32(PO[2] * PO[2]))))
    if dist < f32(f32(0.8) * mn) or dist > f32(f32(1.2) * mx):
        return None
    vc = f32(_sum3(f32(PO[0] * Pn[0]), f32(PO[1] * Pn[1]), f32(PO[2] * Pn[2])) / dist)
    if vc < G["cos_limit"]:
        return None
    lvl = O.predict_scale(float(mx), float(dist), float(G["log_scale_factor"]), int(G["n_levels"]))
    return u, v, f32(u - f32(G["bf"] * invz)), lvl, vc, dist


def test_frustum_oracle_matches_naive_restatement():
    pb = synth.make_frustum_problem(11, n_frames=2, n_q=900, skip_frac=0.0)
    out = O.is_in_frustum(pb)
    n_in = 0
    reasons = dict(behind=0, kept=0)
    for f, G in enumerate(pb["frustum"]):
        for q in range(int(G["q_begin"]), int(G["q_begin"]) + int(G["n_q"])):
            ref = _naive_frustum(G, pb["p_wP"][q], pb["p_normal"][q], pb["p_max_dist"][q], pb["p_min_dist"][q])
            assert bool(out["inview"][q]) == (ref is not None), q
            if ref is None:
                assert out["level"][q] == -1 and not out["proj"][q].any()
                continue
            n_in += 1
            u, v, ur, lvl, vc, dist = ref
            assert out["proj"][q].tobytes() == np.array([u, v, ur], f32).tobytes(), q
            assert out["level"][q] == lvl and out["viewcos"][q] == vc and out["depth"][q] == dist, q
    assert n_in == out["n_inview"].sum()
    # the generator exercises every exit of the function
    assert 0.2 < n_in / len(pb["p_max_dist"]) < 0.9


def test_frustum_skip_mask_and_search_chain():
    pb = synth.make_frustum_problem(12, n_frames=2)
    fo, kp_match, q_match, q_dist, nm = O.search_local_points(pb)
    assert not fo["inview"][pb["p_skip"] != 0].any()
    assert (q_match[fo["inview"] == 0] == -1).all()         # !btrack_inview_ points are never matched
    assert nm.min() > 50


def test_predict_scale_known_answers():
    lsf = float(np.log(f32(1.2)))
    # ratio = 1.2^k * (1 -+ 1e-3): k resp. k + 1 (ceil), clamped to [0, n_levels - 1]
    for k in range(-3, 12):
        for eps, exp in ((-1e-3, k), (1e-3, k + 1)):
            ratio = 1.2 ** k * (1 + eps)
            assert O.predict_scale(ratio, 1.0, lsf, 8) == min(max(exp, 0), 7), (k, eps)
    assert O.predict_scale(1.0, 0.0, lsf, 8) == 7          # ratio = +inf (UB in the reference, defined as the top level)
    assert O.predict_scale(0.0, 1.0, lsf, 8) == 0          # log(0) = -inf
    assert O.predict_scale(float("nan"), 1.0, lsf, 8) == 0


def test_level_table_is_equivalent_to_predict_scale():
    """The kernel counts thresholds instead of calling logf: the table must reproduce ceil(logf(r) / logf(s)) for every
    float, checked on random ratios and on the +-40 ulp neighbourhood of every boundary (host-only entry point)."""
    import vieo_slam_b200.api as api
    rng = np.random.default_rng(5)
    for sf, nl in ((1.2, 8), (2.0, 4), (1.1, 16), (1.2, 1)):
        lsf = float(np.log(f32(sf)))
        tab = api.frustum_level_table(lsf, nl)
        assert tab[0] == 0 and np.all(np.diff(tab[:nl]) > 0) and np.all(np.isinf(tab[nl:]))
        ratios = np.exp(rng.uniform(np.log(0.2), np.log(sf ** (nl + 1)), 4000)).astype(f32)
        for k in range(1, nl):
            bits = tab[k:k + 1].view(np.uint32)[0]
            ratios = np.r_[ratios, (bits + np.arange(-40, 41)).astype(np.uint32).view(f32)]
        for r in ratios:
            lvl = int(sum(r >= tab[k] for k in range(1, nl)))
            assert lvl == O.predict_scale(float(r), 1.0, lsf, nl), (sf, nl, float(r))
    import pytest
    with pytest.raises(api.VieoError):
        api.frustum_level_table(0.0, 8)
    with pytest.raises(api.VieoError):
        api.frustum_level_table(0.18, 17)


def _naive_distinctive(descs):
    """src/MapPoint.cc:346-372 with python ints."""
    N = len(descs)
    if N == 0:
        return -1, -1
    ints = [int.from_bytes(d.tobytes(), "little") for d in descs]
    best_med, best = 2 ** 31 - 1, 0
    for i in range(N):
        row = sorted((ints[i] ^ ints[j]).bit_count() for j in range(N))
        med = row[int(0.5 * (N - 1))]
        if med < best_med:
            best_med, best = med, i
    return best, best_med


def test_distinctive_descriptors_oracle_matches_naive():
    d = synth.make_distinctive_problem(21, n_points=300, max_obs=24, long_lists=(70,))
    best, med = O.distinctive_descriptors(d["pool"], d["ptr"], d["rows"])
    ties = 0
    for p in range(len(best)):
        rows = d["rows"][d["ptr"][p]:d["ptr"][p + 1]]
        b, m = _naive_distinctive(d["pool"][rows])
        assert (best[p], med[p]) == (b, m), p
        if len(rows) > 2:
            ints = [int.from_bytes(x.tobytes(), "little") for x in d["pool"][rows]]
            meds = [sorted((a ^ c).bit_count() for c in ints)[int(0.5 * (len(ints) - 1))] for a in ints]
            ties += meds.count(m) > 1
    assert ties > 5                       # the first-row-wins rule is exercised
    assert (best == -1).sum() > 0         # points without observations
    # identity row table == no row table
    pool2 = d["pool"][d["rows"]]
    b2, m2 = O.distinctive_descriptors(pool2, d["ptr"], None)
    assert np.array_equal(b2, best) and np.array_equal(m2, med)
    # known answers: one observation -> itself with median 0; two -> the first (int(0.5 * 1) = 0 -> median 0 for both)
    one = np.arange(64, dtype=np.uint8).reshape(2, 32)
    b, m = O.distinctive_descriptors(one, [0, 1, 2, 2], None)
    assert list(b) == [0, 0, -1] and list(m) == [0, 0, -1]
    b, m = O.distinctive_descriptors(one, [0, 2], None)
    assert (b[0], m[0]) == (0, 0)


def _gn_step_numeric(pre, Rwb, use_info):
    """One Gauss-Newton step of OptimizeInitialGyroBias with g2o's central-difference Jacobian (delta 1e-9,
    base_unary_edge.hpp) and scipy rotations: shares no code with the oracle."""
    from scipy.spatial.transform import Rotation as Rot

    def err(k, bg):
        dR = pre[k]["Rij"] @ Rot.from_rotvec(pre[k]["JgR"] @ bg).as_matrix()
        return Rot.from_matrix(dR.T @ Rwb[k - 1].T @ Rwb[k]).as_rotvec()
    H = np.zeros((3, 3)); b = np.zeros(3); n = 0
    for k in range(1, len(pre)):
        if pre[k]["dt"] == 0:
            continue
        n += 1
        e = err(k, np.zeros(3))
        J = np.zeros((3, 3))
        for c in range(3):
            d = np.zeros(3); d[c] = 1e-6
            J[:, c] = (err(k, d) - err(k, -d)) / 2e-6
        W = np.linalg.inv(pre[k]["SigmaPRV"][3:6, 3:6]) if use_info else np.eye(3)
        H += J.T @ W @ J
        b -= J.T @ W @ e
    return n, np.linalg.solve(H, b)


def test_gyro_bias_oracle():
    g = synth.make_gyro_bias_problem(31, n_kf=16, kf_gap=(1, 12))
    nz = O.imu_noise()
    pre = O.imu_preintegrate_frames(g["seq"], g["kf_idx"], nz)
    # visual rotations are noisy in practice: perturb Rwb so that the information weighting matters
    rng = np.random.default_rng(3)
    Rwb = np.stack([R @ synth.so3_exp(rng.normal(0, 2e-3, 3)) for R in g["Rwb"]])
    for use_info in (True, False):
        n, dbg = O.gyro_bias_init(pre, Rwb, use_info)
        n2, ref = _gn_step_numeric(pre, Rwb, use_info)
        assert n == n2 == 15
        assert np.abs(dbg - ref).max() < 1e-6 * np.abs(ref).max() + 1e-9, (dbg, ref)
    a, b = O.gyro_bias_init(pre, Rwb, True)[1], O.gyro_bias_init(pre, Rwb, False)[1]
    assert np.abs(a - b).max() > 1e-6                              # bInfo changes the estimate
    # exact rotations: the single GN step recovers the bias the samples carry to first order
    n, dbg = O.gyro_bias_init(pre, g["Rwb"], True)
    assert np.abs(dbg - g["bg_true"]).max() < 2e-4
    # dt == 0 entries and keyframe 0 are ignored; no equation -> 0 and a zero estimate
    pre2 = pre.copy(); pre2[5]["dt"] = 0
    assert O.gyro_bias_init(pre2, g["Rwb"], True)[0] == 14
    n, dbg = O.gyro_bias_init(pre[:1], g["Rwb"][:1], True)
    assert n == 0 and not dbg.any()


def _naive_proj_search(pb, f, qs):
    """src/ORBmatcher.cc:26-227 (single camera, default MatchMultiCam mode, search half) in numpy float32 scalars, with the
    brute-force GetFeaturesInArea of tests/test_oracle_sbp.py."""
    from test_oracle_sbp import hamming, naive_candidates
    F = pb["frames"][f]
    kb, n, qb = int(F["kp_begin"]), int(F["n_kp"]), int(F["q_begin"])
    kps, ur, desc = pb["kps"][kb:kb + n], pb["uright"][kb:kb + n], pb["desc"][kb:kb + n]
    R = F["Rcw"].reshape(3, 3)
    out = {}
    for qi in qs:
        q = qb + qi
        out[qi] = (-1, 256, -1)
        if pb["p_skip"][q]:
            continue
        P, Pn = pb["p_wP"][q], pb["p_normal"][q]
        Pc = [f32(_sum3(f32(R[r, 0] * P[0]), f32(R[r, 1] * P[1]), f32(R[r, 2] * P[2])) + F["tcw"][r]) for r in range(3)]
        if Pc[2] <= 0:
            continue
        invz = f32(f32(1) / Pc[2])
        u = _sum3(f32(F["fx"] * f32(Pc[0] * invz)), f32(f32(0) * f32(Pc[1] * invz)), F["cx"])
        v = _sum3(f32(f32(0) * f32(Pc[0] * invz)), f32(F["fy"] * f32(Pc[1] * invz)), F["cy"])
        if not (F["minx"] <= u < F["maxx"] and F["miny"] <= v < F["maxy"]):
            continue
        PO = [f32(P[k] - F["Ow"][k]) for k in range(3)]
        dist = f32(np.sqrt(_sum3(f32(PO[0] * PO[0]), f32(PO[1] * PO[1]), f32(PO[2] * PO[2]))))
        if dist < f32(f32(0.8) * pb["p_min_dist"][q]) or dist > f32(f32(1.2) * pb["p_max_dist"][q]):
            continue
        if F["check_viewing_angle"] and float(_sum3(f32(PO[0] * Pn[0]), f32(PO[1] * Pn[1]), f32(PO[2] * Pn[2]))) < 0.5 * float(dist):
            continue
        lvl = O.predict_scale(float(pb["p_max_dist"][q]), float(dist), float(F["log_scale_factor"]), int(F["n_levels"]))
        radius = f32(F["th_radius"] * F["scale"][lvl])
        best, bidx = 2 ** 31 - 1, -1
        for j in naive_candidates(F, kps, u, v, radius, -1, -1):
            kl = int(kps["octave"][j])
            if kl < lvl - 1 or kl > lvl:
                continue
            if F["use_bf"]:
                ex, ey = f32(u - kps["x"][j]), f32(v - kps["y"][j])
                if ur[j] >= 0:
                    er = f32(f32(u - f32(F["bf"] * invz)) - ur[j])
                    e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                    if float(f32(e2 * F["inv_level_sigma2"][kl])) > 7.8:
                        continue
                else:
                    e2 = f32(f32(ex * ex) + f32(ey * ey))
                    if float(f32(e2 * F["inv_level_sigma2"][kl])) > 5.99:
                        continue
            d = hamming(pb["q_desc"][q], desc[j])
            if d < best:
                best, bidx = d, j
        out[qi] = (bidx, best if bidx >= 0 else 256, lvl)
    return out


def test_proj_search_oracle_matches_naive_restatement():
    for seed, kw in ((71, dict()), (72, dict(use_bf=False, check_viewing_angle=False, th_radius=4.0)),
                     (73, dict(cluster=True, th_radius=6.0, n_kp=700))):
        pb = synth.make_fuse_problem(seed, n_frames=2, n_q=260, **kw)
        best, dist, lvl = O.proj_search(pb)
        found = gated = 0
        for f in range(2):
            qb = int(pb["frames"][f]["q_begin"])
            ref = _naive_proj_search(pb, f, range(int(pb["frames"][f]["n_q"])))
            for qi, (b, d, l) in ref.items():
                assert (best[qb + qi], dist[qb + qi], lvl[qb + qi]) == (b, d, l), (seed, f, qi)
                found += b >= 0
                gated += l >= 0 and b < 0
        assert found > 60 and gated > 20     # both outcomes occur


def test_properties_of_the_restatements():
    """Size-independent properties (hypothesis): PredictScale is monotone in the ratio and clamped; the winner of
    ComputeDistinctiveDescriptors does not depend on how often the other observations' order is rotated behind it, and its
    median is the minimum over all rows."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.floats(1e-3, 1e3), st.floats(1.0001, 1.5), st.floats(1.05, 2.5), st.integers(1, 16))
    def monotone(ratio, step, sf, nl):
        lsf = float(np.log(f32(sf)))
        a = O.predict_scale(ratio, 1.0, lsf, nl)
        b = O.predict_scale(ratio * step, 1.0, lsf, nl)
        assert 0 <= a <= b <= nl - 1
    monotone()

    rng = np.random.default_rng(8)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
    def distinct(n, seed):
        r = np.random.default_rng(seed)
        proto = r.integers(0, 256, 32, dtype=np.uint8)
        d = np.stack([np.bitwise_xor(proto, (r.random(32) < 0.08).astype(np.uint8) * r.integers(1, 256, 32, dtype=np.uint8))
                      for _ in range(n)])
        best, med = O.distinctive_descriptors(d, [0, n])
        ints = [int.from_bytes(x.tobytes(), "little") for x in d]
        meds = [sorted((a ^ c).bit_count() for c in ints)[int(0.5 * (n - 1))] for a in ints]
        assert med[0] == min(meds) and best[0] == meds.index(min(meds))
    distinct()
    del rng


# ---- Frame::isInFrustum with a camera rig (src/Frame.cc:351-411, mpCameras.size() > 1) ------------------------------------------
def _cross32(a, b):
    return [f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])), f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))]


def _naive_rig_point(G, wP, Pn, mx, mn):
    """src/Frame.cc:351-411 line by line in numpy float32 scalars; the distorted projection through python floats (doubles) and
    math.atan2 — the libm the reference calls — written straight from camera_kb8.h:68-106 / camera_pinhole.h:70-83."""
    import math
    R = G["Rcw"].reshape(3, 3)
    Pcr = [f32(_sum3(f32(R[r, 0] * wP[0]), f32(R[r, 1] * wP[1]), f32(R[r, 2] * wP[2])) + G["tcw"][r]) for r in range(3)]
    res, sum_depth = [], f32(0)
    for ci in range(int(G["n_cams"])):
        C = G["cam"][ci]
        qv, w = C["q_cr"][:3], C["q_cr"][3]
        uv = _cross32(qv, Pcr)
        uv = [f32(x + x) for x in uv]
        cr = _cross32(qv, uv)
        Pc = [f32(f32(f32(Pcr[k] + f32(w * uv[k])) + cr[k]) + C["t_cr"][k]) for k in range(3)]
        twc = [f32(G["Ow"][k] + _sum3(f32(R[0, k] * C["t_rc"][0]), f32(R[1, k] * C["t_rc"][1]), f32(R[2, k] * C["t_rc"][2]))) for k in range(3)]
        if Pc[2] < f32(0):
            continue
        with np.errstate(all="ignore"):
            invz = f32(f32(1) / Pc[2])
        if int(C["model"]) == 0:
            xn, yn = f32(Pc[0] * invz), f32(Pc[1] * invz)
            u = _sum3(f32(C["fx"] * xn), f32(f32(0) * yn), C["cx"])
            v = _sum3(f32(f32(0) * xn), f32(C["fy"] * yn), C["cy"])
        else:
            x, y, z = float(Pc[0]), float(Pc[1]), float(Pc[2])
            r = math.sqrt(x * x + y * y)
            if int(C["model"]) == 2 and r > float(f32(1e-5)):
                th = math.atan2(r, z); th2 = th * th
                td = float(C["k"][3]) * th2
                td += float(C["k"][2]); td *= th2
                td += float(C["k"][1]); td *= th2
                td += float(C["k"][0]); td *= th2
                td += 1; td *= th
                u = f32(float(C["fx"]) * (x * td / r) * 1.0 + float(C["cx"]))
                v = f32(float(C["fy"]) * (y * td / r) * 1.0 + float(C["cy"]))
            else:
                iz = 1.0 / z if z != 0 else math.inf
                u = f32(float(C["fx"]) * x * iz + float(C["cx"]))
                v = f32(float(C["fy"]) * y * iz + float(C["cy"]))
        if u < C["minx"] or u > C["maxx"] or v < C["miny"] or v > C["maxy"]:
            continue
        PO = [f32(wP[k] - twc[k]) for k in range(3)]
        dist = f32(np.sqrt(_sum3(f32(PO[0] * PO[0]), f32(PO[1] * PO[1]), f32(PO[2] * PO[2]))))
        if dist < f32(f32(0.8) * mn) or dist > f32(f32(1.2) * mx):
            continue
        vc = f32(_sum3(f32(PO[0] * Pn[0]), f32(PO[1] * Pn[1]), f32(PO[2] * Pn[2])) / dist)
        if vc < G["cos_limit"]:
            continue
        lvl = O.predict_scale(float(mx), float(dist), float(G["log_scale_factor"]), int(G["n_levels"]))
        res.append((ci, u, v, f32(u - f32(G["bf"] * invz)), lvl, vc))
        sum_depth = f32(sum_depth + dist)
    return res, (f32(sum_depth / f32(len(res))) if res else f32(0))


import pytest  # noqa: E402


@pytest.mark.parametrize("model,n_cams", [(2, 4), (0, 2), (1, 3)])
def test_frustum_rig_oracle_matches_naive_restatement(model, n_cams):
    pb = synth.make_frustum_rig_problem(21 + model, n_frames=2, n_q=500, n_cams=n_cams, model=model, skip_frac=0.0)
    out = O.is_in_frustum_rig(pb)
    n_in, n_multi = 0, 0
    for f, G in enumerate(pb["rig"]):
        for q in range(int(G["q_begin"]), int(G["q_begin"]) + int(G["n_q"])):
            res, depth = _naive_rig_point(G, pb["p_wP"][q], pb["p_normal"][q], pb["p_max_dist"][q], pb["p_min_dist"][q])
            assert bool(out["inview"][q]) == bool(res), q
            assert int(out["cam_mask"][q]) == sum(1 << r[0] for r in res), q
            assert out["depth"][q] == depth, q
            for ci, u, v, ur, lvl, vc in res:
                assert out["proj"][q, ci].tobytes() == np.array([u, v, ur], f32).tobytes(), (q, ci)
                assert out["level"][q, ci] == lvl and out["viewcos"][q, ci] == vc
            for ci in range(4):
                if not (int(out["cam_mask"][q]) >> ci) & 1:
                    assert out["level"][q, ci] == -1 and not out["proj"][q, ci].any()
            n_in += bool(res); n_multi += len(res) > 1
    assert n_in == out["n_inview"].sum() and n_in > 50
    if n_cams >= 3:
        assert n_multi > 10   # points seen by several cameras of the rig: track_depth_ is a mean

"""Oracle of the relocalisation search ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist,
th_far_pts) (src/ORBmatcher.cc:1471-1606), pinned by a brute-force numpy restatement (no grid: every keypoint is tested
against the window) of the same sequential rules."""
import numpy as np

import oracle_lib as O
from vieo_slam_b200 import synth


def _qrot(q, v):  # Eigen::Quaternion::_transformVector
    w, x, y, z = q
    u = np.array([x, y, z])
    uv = 2 * np.cross(u, v)
    return v + w * uv + np.cross(u, uv)


def naive_reloc(pb, f):
    F = pb["frames"][f]
    kb, n = int(F["kp_begin"]), int(F["n_kp"]); qb, nq = int(F["q_begin"]), int(F["n_q"])
    kps = pb["kps"][kb:kb + n]; desc = pb["desc"][kb:kb + n]
    blocked = pb["kp_blocked"][kb:kb + n].astype(bool).copy()
    orb_dist = int(pb["reloc"][f]["orb_dist"]); lsf = float(pb["reloc"][f]["log_scale_factor"])
    qc = F["qcw"]; tc = F["tcw"]
    twc = _qrot(np.array([qc[0], -qc[1], -qc[2], -qc[3]]), -tc)
    kp_match = np.full(n, -1, np.int32); q_match = np.full(nq, -1, np.int32); q_dist = np.full(nq, 256, np.int32)
    hist = [[] for _ in range(30)]
    nm = 0
    f32 = np.float32
    for i in range(nq):
        X = pb["q_Xw"][qb + i]
        Pc = _qrot(qc, X) + tc
        if F["th_far"] > 0 and Pc[2] > float(F["th_far"]):
            continue
        invz = f32(1.0 / Pc[2])
        u = f32(f32(F["fx"] * f32(f32(Pc[0]) * invz)) + F["cx"]); v = f32(f32(F["fy"] * f32(f32(Pc[1]) * invz)) + F["cy"])
        if not (u >= F["minx"] and u < F["maxx"] and v >= F["miny"] and v < F["maxy"]):
            continue
        d3 = f32(np.sqrt(np.sum((X - twc) ** 2)))
        mx, mn = pb["q_max_dist"][qb + i], pb["q_min_dist"][qb + i]
        if d3 < f32(0.8) * mn or d3 > f32(1.2) * mx:
            continue
        lvl = O.predict_scale(mx, d3, lsf, int(F["n_levels"]))
        r = f32(F["th"] * F["scale"][lvl])
        # GetFeaturesInArea order: cells ix-major, iy, insertion (= keypoint) order inside a cell
        cx_ = np.round((kps["x"] - F["minx"]) * F["grid_winv"]).astype(int); cy_ = np.round((kps["y"] - F["miny"]) * F["grid_hinv"]).astype(int)
        ok = (cx_ >= 0) & (cx_ < 64) & (cy_ >= 0) & (cy_ < 48)
        x0 = max(0, int(np.floor((u - F["minx"] - r) * F["grid_winv"]))); x1 = min(63, int(np.ceil((u - F["minx"] + r) * F["grid_winv"])))
        y0 = max(0, int(np.floor((v - F["miny"] - r) * F["grid_hinv"]))); y1 = min(47, int(np.ceil((v - F["miny"] + r) * F["grid_hinv"])))
        sel = ok & (cx_ >= x0) & (cx_ <= x1) & (cy_ >= y0) & (cy_ <= y1)
        sel &= (kps["octave"] >= lvl - 1) & (kps["octave"] <= lvl + 1)
        sel &= (np.abs(kps["x"] - u) < r) & (np.abs(kps["y"] - v) < r)
        idx = np.nonzero(sel)[0]
        idx = idx[np.lexsort((idx, cy_[idx], cx_[idx]))]
        best, bi = 256, -1
        for j in idx:
            if blocked[j]:
                continue
            d = O.descriptor_distance(pb["q_desc"][qb + i], desc[j])
            if d < best:
                best, bi = d, j
        if best <= orb_dist:
            kp_match[bi] = i; blocked[bi] = True; q_match[i] = bi; q_dist[i] = best; nm += 1
            rot = f32(pb["q_angle"][qb + i] - kps["angle"][bi])
            if rot < 0:
                rot = f32(rot + f32(360))
            b = int(np.round(f32(rot * f32(1.0 / 30))))   # round half away from zero == numpy for the positive halves hit here
            hist[0 if b == 30 else b].append(bi)
    cnt = [len(h) for h in hist]
    order = sorted(range(30), key=lambda b: (-cnt[b], b))
    m1, m2, m3 = (cnt[order[0]], cnt[order[1]], cnt[order[2]])
    keep = {order[0]}
    if not (np.float32(m2) < np.float32(0.1) * np.float32(m1)):
        keep.add(order[1])
        if not (np.float32(m3) < np.float32(0.1) * np.float32(m1)):
            keep.add(order[2])
    for b in range(30):
        if b not in keep:
            for k in hist[b]:
                kp_match[k] = -1; nm -= 1
    return kp_match, q_match, q_dist, nm


def test_reloc_oracle_equals_naive_restatement():
    pb = synth.make_reloc_problem(5, n_frames=2, n_kp=500, n_q=260, th=10.0, orb_dist=100)
    kpm, qm, qd, ql, nm = O.sbp_reloc(pb)
    for f in range(2):
        F = pb["frames"][f]
        kb, n = int(F["kp_begin"]), int(F["n_kp"]); qb, nq = int(F["q_begin"]), int(F["n_q"])
        k2, q2, d2, n2 = naive_reloc(pb, f)
        assert np.array_equal(qm[qb:qb + nq], q2) and np.array_equal(qd[qb:qb + nq], d2)
        assert np.array_equal(kpm[kb:kb + n], k2) and nm[f] == n2
    assert nm.min() > 60 and (ql == -1).sum() > 20 and (ql >= 0).sum() > 300


def test_reloc_rules():
    pb = synth.make_reloc_problem(8, n_frames=1, n_kp=600, n_q=300, th=15.0, orb_dist=60)
    kpm, qm, qd, ql, nm = O.sbp_reloc(pb)
    taken = qm[qm >= 0]
    assert len(np.unique(taken)) == len(taken), "every accepted match claims its keypoint"
    assert (qd[qm >= 0] <= 60).all()
    assert not pb["kp_blocked"][taken].any()
    # a looser ORBdist can only add matches up to the claim rule; the far cut removes queries
    pb2 = dict(pb); pb2["reloc"] = pb["reloc"].copy(); pb2["reloc"]["orb_dist"] = 100
    assert O.sbp_reloc(pb2)[4][0] >= nm[0]
    pb3 = dict(pb); pb3["frames"] = pb["frames"].copy(); pb3["frames"]["th_far"] = 4.0
    assert (O.sbp_reloc(pb3)[1] >= 0).sum() < (qm >= 0).sum()

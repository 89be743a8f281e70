"""CPU tests pinning oracle/sbp_oracle.cc (ORBmatcher::SearchByProjection, both tracking overloads).  The reference holds
no vectors for this path, so the oracle is checked against an independent, deliberately naive numpy restatement written
straight from src/ORBmatcher.cc:230-335, 1303-1467 and src/FrameBase.cpp:95-174 (no shared code with the oracle), and
against hand-built cases whose answer is known."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import KP_DTYPE, SBP_LAST_FRAME, SBP_LOCAL_MAP

f32 = np.float32
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming(a, b):
    return int(POP[np.bitwise_xor(a, b)].sum())


def qrot(q, v):
    w, x, y, z = q
    qv = np.array([x, y, z])
    uv = 2 * np.cross(qv, v)
    return v + w * uv + np.cross(qv, uv)


def naive_candidates(F, kps, x, y, r, minlevel, maxlevel):
    """GetFeaturesInArea by brute force over all keypoints, then ordered like the grid walk (cell x, cell y, index)."""
    x, y, r = f32(x), f32(y), f32(r)
    lo_x = max(0, int(np.floor(f32(f32(x - F["minx"]) - r) * F["grid_winv"])))
    hi_x = min(63, int(np.ceil(f32(f32(x - F["minx"]) + r) * F["grid_winv"])))
    lo_y = max(0, int(np.floor(f32(f32(y - F["miny"]) - r) * F["grid_hinv"])))
    hi_y = min(47, int(np.ceil(f32(f32(y - F["miny"]) + r) * F["grid_hinv"])))
    if lo_x >= 64 or hi_x < 0 or lo_y >= 48 or hi_y < 0:
        return []
    out = []
    check = minlevel > 0 or maxlevel >= 0
    for j in range(len(kps)):
        px = int(np.floor(f32(f32(kps["x"][j] - F["minx"]) * F["grid_winv"]) + f32(0.5)))   # round() of a non-negative float
        py = int(np.floor(f32(f32(kps["y"][j] - F["miny"]) * F["grid_hinv"]) + f32(0.5)))
        if not (0 <= px < 64 and 0 <= py < 48) or not (lo_x <= px <= hi_x and lo_y <= py <= hi_y):
            continue
        if check and (kps["octave"][j] < minlevel or (maxlevel >= 0 and kps["octave"][j] > maxlevel)):
            continue
        if abs(f32(kps["x"][j] - x)) < r and abs(f32(kps["y"][j] - y)) < r:
            out.append((px, py, j))
    return [j for _, _, j in sorted(out)]


def naive_sbp(pb, f):
    F = pb["frames"][f]
    kb, n, qb, nq = int(F["kp_begin"]), int(F["n_kp"]), int(F["q_begin"]), int(F["n_q"])
    kps, ur, desc = pb["kps"][kb:kb + n], pb["uright"][kb:kb + n], pb["desc"][kb:kb + n]
    blocked = pb["kp_blocked"][kb:kb + n].astype(bool).copy()
    kp_match = np.full(n, -1, np.int32); q_match = np.full(nq, -1, np.int32); q_dist = np.full(nq, 256, np.int32)
    nmatches = 0
    hist = [[] for _ in range(30)]
    if pb["mode"] == SBP_LAST_FRAME:
        qci = F["qcw"] * np.array([1, -1, -1, -1])
        tz = (qrot(F["qlw"], -qrot(qci, F["tcw"])) + F["tlw"])[2]
        fwd = tz > float(F["b"]) and not F["mono"]
        bwd = -tz > float(F["b"]) and not F["mono"]
    for i in range(nq):
        q = qb + i
        if pb["mode"] == SBP_LAST_FRAME:
            P = qrot(F["qcw"], pb["q_Xw"][q]) + F["tcw"]
            if F["th_far"] > 0 and P[2] > float(F["th_far"]):
                continue
            invz = f32(1.0 / P[2])
            if invz < 0:
                continue
            u = f32(f32(F["fx"] * f32(f32(P[0]) * invz)) + F["cx"]); v = f32(f32(F["fy"] * f32(f32(P[1]) * invz)) + F["cy"])
            if not (F["minx"] <= u < F["maxx"] and F["miny"] <= v < F["maxy"]):
                continue
            o = int(pb["q_level"][q])
            r = f32(F["th"] * F["scale"][o])
            lv = (0, o) if fwd else (o, -1) if bwd else (o - 1, o + 1)
            urp = f32(u - f32(F["bf"] * invz))
        else:
            if F["th_far"] > 0 and pb["q_depth"][q] > F["th_far"]:
                continue
            o = int(pb["q_level"][q])
            r = f32(2.5) if float(pb["q_viewcos"][q]) > 0.998 else f32(4.0)
            if float(F["th"]) != 1.0:
                r = f32(r * F["th"])
            r = f32(r * F["scale"][o])
            u, v, urp = pb["q_proj"][q]
            lv = (o - 1, o)
        scored = []
        for j in naive_candidates(F, kps, u, v, r, *lv):
            if blocked[j]:
                continue
            if ur[j] > 0 and abs(f32(urp - ur[j])) > r:
                continue
            scored.append((hamming(pb["q_desc"][q], desc[j]), len(scored), j))
        if not scored:
            continue
        scored.sort()
        d1, _, j1 = scored[0]
        if d1 > 100:
            continue
        if pb["mode"] == SBP_LOCAL_MAP and len(scored) > 1:
            d2, _, j2 = scored[1]
            if kps["octave"][j1] == kps["octave"][j2] and f32(d1) > f32(F["nn_ratio"] * f32(d2)):
                continue
        kp_match[j1] = i; q_match[i] = j1; q_dist[i] = d1
        if pb["q_flags"][q] & 1:
            blocked[j1] = True
        nmatches += 1
        if pb["mode"] == SBP_LAST_FRAME and F["check_orientation"]:
            rot = f32(pb["q_angle"][q] - kps["angle"][j1])
            if rot < 0:
                rot = f32(rot + f32(360))
            b = int(np.floor(f32(rot * f32(1.0 / 30)) + f32(0.5)))
            hist[0 if b == 30 else b].append(j1)
    if pb["mode"] == SBP_LAST_FRAME and F["check_orientation"]:
        cnt = [len(h) for h in hist]
        order = sorted(range(30), key=lambda b: (-cnt[b], b))   # three largest, earliest bin first on ties
        top = [order[0] if cnt[order[0]] > 0 else -1]
        for k in (1, 2):
            ok = cnt[order[k]] > 0 and not (f32(cnt[order[k]]) < f32(0.1) * f32(cnt[order[0]])) and (k == 1 or top[1] != -1)
            top.append(order[k] if ok else -1)
        for b in range(30):
            if b not in top:
                for j in hist[b]:
                    kp_match[j] = -1
                    nmatches -= 1
    return kp_match, q_match, q_dist, nmatches


CASES = [dict(mode=SBP_LAST_FRAME, th=15.0, motion="still"), dict(mode=SBP_LAST_FRAME, th=7.0, motion="forward"),
         dict(mode=SBP_LAST_FRAME, th=30.0, motion="backward", th_far=9.0), dict(mode=SBP_LAST_FRAME, th=15.0, mono=True, motion="forward"),
         dict(mode=SBP_LOCAL_MAP, th=1.0, blocked_frac=0.3), dict(mode=SBP_LOCAL_MAP, th=3.0, blocked_frac=0.1, th_far=8.0),
         dict(mode=SBP_LAST_FRAME, th=15.0, cluster=True, n_kp=700, n_q=300)]


@pytest.mark.parametrize("kw", CASES)
def test_oracle_matches_naive_restatement(kw):
    kw = dict(kw)
    kw.setdefault("n_kp", 500); kw.setdefault("n_q", 300)
    pb = synth.make_sbp_problem(31, n_frames=2, **kw)
    if kw["mode"] == SBP_LOCAL_MAP:
        pb["frames"]["nn_ratio"] = 0.8
    kp_match, q_match, q_dist, nm = O.search_by_projection(pb)
    for f in range(2):
        F = pb["frames"][f]
        ks = slice(int(F["kp_begin"]), int(F["kp_begin"] + F["n_kp"])); qs = slice(int(F["q_begin"]), int(F["q_begin"] + F["n_q"]))
        nk, nq_, nd, nn = naive_sbp(pb, f)
        assert np.array_equal(q_match[qs], nq_) and np.array_equal(q_dist[qs], nd)
        assert np.array_equal(kp_match[ks], nk) and nm[f] == nn
    assert nm.min() > 20          # the generator does produce matches
    assert (q_match >= 0).sum() > nm.sum() - 1 or kw["mode"] == SBP_LOCAL_MAP


def _one_frame(kps_xy, octaves, descs, mode=SBP_LAST_FRAME):
    pb = synth.make_sbp_problem(1, n_frames=1, mode=mode, n_kp=60, n_q=4)
    n = len(kps_xy)
    kp = np.zeros(n, KP_DTYPE); kp["x"], kp["y"] = np.array(kps_xy, f32).T; kp["octave"] = octaves
    F = pb["frames"][0]
    F["n_kp"] = n
    pb["kps"], pb["uright"], pb["desc"] = kp, np.full(n, -1, f32), np.array(descs, np.uint8)
    pb["kp_blocked"] = np.zeros(n, np.uint8)
    F["qcw"], F["tcw"], F["qlw"], F["tlw"] = [1, 0, 0, 0], 0, [1, 0, 0, 0], 0
    return pb, F


def test_known_answers_claim_order_and_ties():
    """Two map points that prefer the same keypoint: the first (with observations) keeps it, the second falls back to its
    runner-up; equal distances keep the first candidate in grid order (column-major cells)."""
    z = np.zeros(32, np.uint8)
    a = z.copy(); a[0] = 0x0f          # 4 bits from z
    pb, F = _one_frame([(100, 100), (104, 100), (100.5, 90)], [0, 0, 0], [z, a, a])
    fx, fy, cx, cy = (float(F[k]) for k in ("fx", "fy", "cx", "cy"))
    P = np.array([[(101 - cx) / fx * 2, (100 - cy) / fy * 2, 2.0]] * 2)
    F["n_q"] = 2; F["check_orientation"] = 0
    pb["q_Xw"] = P; pb["q_level"] = np.zeros(2, np.int32); pb["q_angle"] = np.zeros(2, f32)
    pb["q_desc"] = np.stack([z, z]); pb["q_flags"] = np.array([1, 1], np.uint8)
    kp_match, q_match, q_dist, nm = O.search_by_projection(pb)
    # keypoints 1 and 2 tie at distance 4; cell x of kp 2 (100.5 -> 9) == kp 1's (104 -> 9), cell y smaller -> kp 2 first
    assert list(q_match) == [0, 2] and list(q_dist) == [0, 4] and nm[0] == 2 and list(kp_match) == [0, -1, 1]
    pb["q_flags"] = np.array([0, 1], np.uint8)   # a temporal point (no observations) does not block: the second overwrites it
    kp_match, q_match, q_dist, nm = O.search_by_projection(pb)
    assert list(q_match) == [0, 0] and list(kp_match) == [1, -1, -1] and nm[0] == 2


def test_known_answers_ratio_and_rotation():
    z = np.zeros(32, np.uint8)
    a = z.copy(); a[0] = 0xff; a[1] = 0x03     # 10 bits
    b = z.copy(); b[0] = 0xff; b[1] = 0x0f     # 12 bits
    pb, F = _one_frame([(200, 200), (203, 201)], [1, 1], [a, b], mode=SBP_LOCAL_MAP)
    F["n_q"] = 1; F["th"] = 1.0
    pb["q_proj"] = np.array([[201, 200, -1]], f32); pb["q_level"] = np.array([1], np.int32)
    pb["q_viewcos"] = np.array([0.5], f32); pb["q_depth"] = np.array([3], f32)
    pb["q_desc"] = z[None].copy(); pb["q_flags"] = np.array([1], np.uint8)
    F["nn_ratio"] = 0.8      # 10 > 0.8 * 12 = 9.6 on the same level: rejected
    assert O.search_by_projection(pb)[3][0] == 0
    F["nn_ratio"] = 0.9      # 10 <= 10.8: accepted
    assert O.search_by_projection(pb)[1][0] == 0
    pb["kps"]["octave"][1] = 0   # second best on another level: the ratio test does not apply
    F["nn_ratio"] = 0.8
    assert O.search_by_projection(pb)[1][0] == 0

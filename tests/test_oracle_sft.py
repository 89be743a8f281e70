"""CPU: the oracle's ORBmatcher::SearchForTriangulation (oracle/sft_oracle.cc) against a deliberately naive python
restatement written straight from src/ORBmatcher.cc:896-1150 (dict-based FeatureVector walk, python-int Hamming, numpy
float32 / float64 scalars for the epipolar test in the reference's types) and known answers."""
import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import synth

f32, f64 = np.float32, np.float64


def _naive(pb, p):
    P = pb["pairs"][p]

    def kf(kb, n, nb, nn, pb_, ib):
        ptr = pb["fv_ptr"][pb_:pb_ + nn + 1]
        fv = {int(pb["fv_node"][nb + a]): [int(v) for v in pb["fv_idx"][ib + ptr[a]:ib + ptr[a + 1]]] for a in range(nn)}
        return pb["kps"][kb:kb + n], pb["uright"][kb:kb + n], pb["desc"][kb:kb + n], pb["has_mp"][kb:kb + n], fv
    k1, u1, d1, m1, fv1 = kf(P["kp1_begin"], P["n_kp1"], P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"])
    k2, u2, d2, m2, fv2 = kf(P["kp2_begin"], P["n_kp2"], P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"])
    F = P["F12"].reshape(3, 3).astype(f64)
    matched2, entries, hist = set(), [], [[] for _ in range(30)]
    for node in sorted(set(fv1) & set(fv2)):  # the lower_bound walk visits exactly the common ids in ascending order
        for idx1 in fv1[node]:
            if m1[idx1] or (P["only_stereo"] and not u1[idx1] >= 0):
                continue
            best, bestj = 50, -1
            for idx2 in fv2[node]:
                if m2[idx2] or idx2 in matched2 or (P["only_stereo"] and not u2[idx2] >= 0):
                    continue
                dist = int(np.unpackbits(d1[idx1] ^ d2[idx2]).sum())
                if dist > best:
                    continue
                if not u1[idx1] >= 0 and not u2[idx2] >= 0:
                    dx, dy = f32(P["ex"]) - k2["x"][idx2], f32(P["ey"]) - k2["y"][idx2]
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * P["scale_factor2"][k2["octave"][idx2]]):
                        continue
                x1, y1, x2, y2 = (f64(v) for v in (k1["x"][idx1], k1["y"][idx1], k2["x"][idx2], k2["y"][idx2]))
                a = f32(x1 * F[0, 0] + y1 * F[1, 0] + F[2, 0]); b = f32(x1 * F[0, 1] + y1 * F[1, 1] + F[2, 1])
                c = f32(x1 * F[0, 2] + y1 * F[1, 2] + F[2, 2])
                num = f32(f64(a) * x2 + f64(b) * y2 + f64(c))
                den = f32(f32(a * a) + f32(b * b))
                if den == 0 or not (f32(f32(num * num) / den) < f32(f32(3.84) * P["level_sigma2_2"][k2["octave"][idx2]])):
                    continue
                best, bestj = dist, idx2
            if bestj >= 0:
                matched2.add(bestj)
                entries.append([idx1, bestj, True])
                if P["check_orientation"]:
                    rot = f32(k1["angle"][idx1] - k2["angle"][bestj])
                    if rot < 0:
                        rot = f32(rot + f32(360))
                    b_ = int(np.floor(f64(f32(rot * f32(1.0 / 30))) + 0.5))  # C round(): half away from zero, rot >= 0
                    hist[0 if b_ == 30 else b_].append(idx1)
    n = len(entries)
    if P["check_orientation"]:
        sizes = [len(h) for h in hist]
        order = sorted(range(30), key=lambda i: (-sizes[i], i))
        m1_, m2_, m3_ = (sizes[order[k]] for k in range(3))
        keep = {order[0]} if m1_ > 0 else set()
        if m2_ > 0 and not m2_ < f32(0.1) * f32(m1_):
            keep.add(order[1])
            if m3_ > 0 and not m3_ < f32(0.1) * f32(m1_):
                keep.add(order[2])
        drop = {i for b_ in range(30) if b_ not in keep for i in hist[b_]}
        for e in entries:
            if e[0] in drop:
                e[2] = False
                n -= 1
    return np.array([[e[0], e[1]] for e in entries if e[2]], np.int32).reshape(-1, 2), n


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_matches_naive_restatement(seed):
    pb = synth.make_sft_problem(seed, n_pairs=4, n_kp=300, n_nodes=25)
    total = 0
    for p in range(4):
        got, n = O.search_for_triangulation(pb, p)
        ref, rn = _naive(pb, p)
        # ComputeThreeMaxima keeps the FIRST of equally large bins; the naive sort above does too (stable on the index)
        assert n == rn and np.array_equal(got, ref), p
        total += n
    assert total > 100


def test_known_answers():
    pb = synth.make_sft_problem(5, n_pairs=1, n_kp=200, n_nodes=10)
    P = pb["pairs"][0]
    pairs, n = O.search_for_triangulation(pb, 0)
    assert n == len(pairs) > 20
    assert len(set(pairs[:, 1].tolist())) == n, "a keyframe-2 keypoint is matched once"
    k1 = pb["has_mp"][P["kp1_begin"]:P["kp1_begin"] + P["n_kp1"]]; k2 = pb["has_mp"][P["kp2_begin"]:P["kp2_begin"] + P["n_kp2"]]
    assert not k1[pairs[:, 0]].any() and not k2[pairs[:, 1]].any(), "keypoints with a map point never match"
    d1 = pb["desc"][P["kp1_begin"]:][pairs[:, 0]]; d2 = pb["desc"][P["kp2_begin"]:][pairs[:, 1]]
    assert (np.unpackbits(d1 ^ d2, axis=1).sum(1) <= 50).all()
    # everything has a map point -> nothing to match
    pb2 = dict(pb); pb2["has_mp"] = np.ones_like(pb["has_mp"])
    assert O.search_for_triangulation(pb2, 0)[1] == 0
    # no common vocabulary node -> nothing
    pb3 = dict(pb); pb3["fv_node"] = pb["fv_node"].copy()
    pb3["fv_node"][P["node2_begin"]:P["node2_begin"] + P["n_nodes2"]] += 1
    assert O.search_for_triangulation(pb3, 0)[1] == 0


def _naive_bow(pb, p):
    P = pb["pairs"][p]
    k1 = slice(int(P["kp1_begin"]), int(P["kp1_begin"] + P["n_kp1"])); k2 = slice(int(P["kp2_begin"]), int(P["kp2_begin"] + P["n_kp2"]))
    kp1, d1, ok1 = pb["kps"][k1], pb["desc"][k1], pb["mp_ok"][k1]
    kp2, d2 = pb["kps"][k2], pb["desc"][k2]

    def fv(nb, nn, pb_, ib):
        ptr = pb["fv_ptr"][pb_:pb_ + nn + 1]
        return {int(pb["fv_node"][nb + a]): [int(v) for v in pb["fv_idx"][ib + ptr[a]:ib + ptr[a + 1]]] for a in range(nn)}
    f1 = fv(P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"]); f2 = fv(P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"])
    match = np.full(int(P["n_kp2"]), -1, np.int32)
    hist = [[] for _ in range(30)]
    for node in sorted(set(f1) & set(f2)):
        for i1 in f1[node]:
            if not ok1[i1]:
                continue
            ds = [(int(np.unpackbits(d1[i1] ^ d2[i2]).sum()), k, i2) for k, i2 in enumerate(f2[node]) if match[i2] < 0]
            if not ds:
                continue
            ds.sort(key=lambda t: (t[0], t[1]))
            b1 = ds[0][0]; b2 = ds[1][0] if len(ds) > 1 else 256
            if b1 <= 50 and f32(b1) < f32(P["nn_ratio"]) * f32(b2):
                i2 = ds[0][2]
                match[i2] = i1
                if P["check_orientation"]:
                    rot = f32(kp1["angle"][i1] - kp2["angle"][i2])
                    if rot < 0:
                        rot = f32(rot + f32(360))
                    b_ = int(np.floor(f64(f32(rot * f32(1.0 / 30))) + 0.5))
                    hist[0 if b_ == 30 else b_].append(i2)
    if P["check_orientation"]:
        sizes = [len(h) for h in hist]
        order = sorted(range(30), key=lambda i: (-sizes[i], i))
        keep = {order[0]} if sizes[order[0]] > 0 else set()
        if sizes[order[1]] > 0 and not sizes[order[1]] < f32(0.1) * f32(sizes[order[0]]):
            keep.add(order[1])
            if sizes[order[2]] > 0 and not sizes[order[2]] < f32(0.1) * f32(sizes[order[0]]):
                keep.add(order[2])
        for b_ in range(30):
            if b_ not in keep:
                for i2 in hist[b_]:
                    match[i2] = -1
    return match, int((match >= 0).sum())


@pytest.mark.parametrize("seed", [3, 4])
def test_search_by_bow_oracle_matches_naive(seed):
    pb = synth.make_bow_problem(seed, n_pairs=4, n_kp=300, n_nodes=25)
    tot = 0
    for p in range(4):
        got, n = O.search_by_bow(pb, p)
        ref, rn = _naive_bow(pb, p)
        assert n == rn and np.array_equal(got, ref), p
        tot += n
    assert tot > 80

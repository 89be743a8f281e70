"""CPU: pin the oracle (oracle/*.cc) against the committed cv2 goldens and known-answer checks."""
import math
import struct

import numpy as np
import pytest

import oracle_lib as O


def _levels(name, g):
    n = 0
    while f"{name}_cand{n}" in g:
        n += 1
    return n


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_pyramid_fast_blur_match_cv2(goldens, name):
    g = goldens
    nlev = _levels(name, g)
    orb = O.OrbOracle(1200 if name == "A" else 300, float(g[f"{name}_scale"]), nlev, 20, 7)
    img = g[f"{name}_img"]
    n, kps, desc, mono = orb.extract(img)
    assert n > 0 and mono == 0
    for l in range(nlev):
        lv = orb.level(l)
        crc = np.array([int(lv.astype(np.uint64).sum()),
                        int((lv.astype(np.uint64) * (np.arange(lv.size, dtype=np.uint64).reshape(lv.shape) % 251)).sum())],
                       np.uint64)
        assert (crc == g[f"{name}_crc{l}"]).all(), f"level {l} pixels differ from cv2.resize chain"
        if f"{name}_L{l}" in g:
            assert np.array_equal(lv, g[f"{name}_L{l}"])
        cand = orb.candidates(l)
        assert np.array_equal(cand, g[f"{name}_cand{l}"].astype(np.int32)), f"level {l} FAST candidates differ"
        if f"{name}_blur{l}" in g:
            assert np.array_equal(O.gaussian_blur7(lv), g[f"{name}_blur{l}"])


def test_min_threshold_cells_present(goldens):
    # image B has a low-contrast half: some cells must have fired only at minThFAST (response < 20)
    c = goldens["B_cand0"].astype(np.int32)
    assert (c[:, 2] < 20).any() and (c[:, 2] >= 20).any()


def test_fast_atan2_matches_cv2(goldens):
    yx = goldens["atan_in"]
    got = np.array([O.fast_atan2(float(y), float(x)) for y, x in yx], np.float32)
    assert np.array_equal(got.view(np.uint32), goldens["atan_out"].view(np.uint32))


def test_sincos_matches_libm_sample():
    # restated glibc sincosf vs this box's libm (exhaustive version: tools/check_sincosf.c)
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float; libm.cosf.argtypes = [ctypes.c_float]
    libm.sinf.restype = ctypes.c_float; libm.sinf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(0, 2 * math.pi, 20000), [0.0, 1e-5, math.pi / 4, math.pi / 2, math.pi, 6.2831855]])
    for x in xs.astype(np.float32):
        s, c = O.sincosf(float(x))
        assert struct.pack("f", s) == struct.pack("f", libm.sinf(float(x)))
        assert struct.pack("f", c) == struct.pack("f", libm.cosf(float(x)))


def test_tables_euroc():
    t = O.OrbOracle(1200, 1.2, 8, 20, 7).tables()
    assert list(t["quota"]) == [261, 217, 181, 151, 126, 105, 87, 72]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    t = O.OrbOracle(1000, 1.2, 8, 20, 7).tables()
    assert list(t["quota"]) == [217, 181, 151, 126, 105, 87, 73, 60]


def test_extract_structure(goldens):
    img = goldens["A_img"]
    orb = O.OrbOracle(1200, 1.2, 8, 20, 7)
    n, kps, desc, mono = orb.extract(img)
    assert 1100 <= n <= 1300
    assert (np.diff(kps["octave"]) >= 0).all()  # level-ordered
    sc = orb.tables()["scale"]
    for l in range(8):
        k = kps[kps["octave"] == l]
        cand = orb.candidates(l)
        cs = {(int(x), int(y)): r for x, y, r in cand}
        # every keypoint is one of the level's candidates (after undoing the scale) with its response
        for p in k:
            x, y = (p["x"], p["y"]) if l == 0 else (p["x"] / sc[l], p["y"] / sc[l])
            key = (int(round(float(x))), int(round(float(y))))
            assert key in cs and cs[key] == p["response"]
        assert (k["size"] == np.float32(int(31 * sc[l]))).all()
    assert (kps["angle"] >= 0).all() and (kps["angle"] < 360.0001).all()
    # same input twice -> identical output (determinism incl. the quadtree tie-break)
    n2, kps2, desc2, _ = orb.extract(img)
    assert n2 == n and np.array_equal(desc, desc2) and kps.tobytes() == kps2.tobytes()
    assert orb.extract(np.zeros((0, 0), np.uint8))[0] == -1


def test_quadtree_properties():
    rng = np.random.default_rng(9)
    W, H = 752, 480
    pts = set()
    while len(pts) < 5000:
        pts.add((int(rng.integers(19, W - 19)), int(rng.integers(19, H - 19))))
    xyr = np.array([(x, y, int(rng.integers(7, 120))) for x, y in sorted(pts, key=lambda p: (p[1], p[0]))], np.int32)
    for N in (50, 261, 1000):
        pick = O.quadtree(xyr, W, H, N)
        assert len(set(pick.tolist())) == len(pick)
        assert N <= len(pick) <= N + 3
    few = xyr[:40]
    assert sorted(O.quadtree(few, W, H, 261).tolist()) == list(range(40))  # fewer candidates than quota: all kept
    assert len(O.quadtree(xyr[:0], W, H, 261)) == 0


def test_lapping_area_partition(goldens):
    img = goldens["B_img"]
    orb = O.OrbOracle(300, 1.2, 4, 20, 7)
    n, kps, desc, _ = orb.extract(img)
    n2, kps2, desc2, mono = orb.extract(img, lapping=[100, 200])
    assert n2 == n and 0 < mono < n
    inside = (kps2["x"] >= 100) & (kps2["x"] <= 200)
    assert not inside[:mono].any() and inside[mono:].all()
    a = sorted(zip(kps["x"].tolist(), kps["y"].tolist(), kps["octave"].tolist(), map(bytes, desc)))
    b = sorted(zip(kps2["x"].tolist(), kps2["y"].tolist(), kps2["octave"].tolist(), map(bytes, desc2)))
    assert a == b


def test_hamming_oracle(goldens):
    q, t = goldens["bf_q"], goldens["bf_t"]
    idx, dist = O.hamming_knn2(q, t)
    assert np.array_equal(dist, goldens["bf_dist"])
    assert np.array_equal(idx, goldens["bf_idx"])
    # known answer: python popcount
    for i in (0, 7, 100):
        for j in (0, 5, 9, 200):
            ref = sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(q[i], t[j]))
            assert O.descriptor_distance(q[i], t[j]) == ref
    # CSR search == brute force when every row lists all train rows in order
    rp = np.arange(0, (len(q) + 1) * len(t), len(t), dtype=np.int32)
    cand = np.tile(np.arange(len(t), dtype=np.int32), len(q))
    bd, bi, sd, si = O.hamming_csr(q, t, rp, cand)
    assert np.array_equal(bd, dist[:, 0]) and np.array_equal(bi, idx[:, 0]) and np.array_equal(sd, dist[:, 1])
    # empty rows
    bd, bi, sd, si = O.hamming_csr(q[:3], t, np.zeros(4, np.int32), np.zeros(0, np.int32))
    assert (bd == 256).all() and (bi == -1).all()


def test_fisheye_matches_against_cv2_bfmatcher():
    """Frame::ComputeStereoFishEyeMatches brute-force half: the oracle against cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2)
    on the in-area slices of every camera pair + the ratio test evaluated in numpy float32 / float64."""
    cv2 = pytest.importorskip("cv2")
    r = np.random.default_rng(17)
    n_cams, cap = 4, 96
    desc = r.integers(0, 256, (n_cams, cap, 32), dtype=np.uint8)
    desc[1, 30:60] = desc[0, 10:40] ^ (r.integers(0, 256, (30, 32)) < 20).astype(np.uint8)  # near copies
    nk = np.array([90, 96, 50, 7], np.int32); nm = np.array([5, 20, 50, 0], np.int32)
    idx, dist, good = O.fisheye_matches(desc, nk, nm)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    p = 0
    for i in range(n_cams - 1):
        for j in range(i + 1, n_cams):
            if nm[i] >= nk[i] or nm[j] >= nk[j]:
                assert (idx[p] == -1).all() and not good[p].any()
                p += 1
                continue
            m = bf.knnMatch(desc[i, nm[i]:nk[i]], desc[j, nm[j]:nk[j]], k=2)
            for q, mm in enumerate(m):
                assert [x.trainIdx for x in mm] == [v for v in idx[p, q] if v >= 0]
                assert [int(x.distance) for x in mm] == [int(v) for v, k in zip(dist[p, q], idx[p, q]) if k >= 0]
                g = len(mm) >= 2 and (np.float64(np.float32(mm[0].distance)) < np.float64(np.float32(mm[1].distance)) * 0.7 or
                                      (mm[0].distance < 75 and
                                       np.float64(np.float32(mm[0].distance)) < np.float64(np.float32(mm[1].distance)) * 0.9))
                assert bool(good[p, q]) == bool(g)
            assert not good[p, len(m):].any()
            p += 1
    assert good.sum() > 10

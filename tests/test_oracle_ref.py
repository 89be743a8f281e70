"""CPU: oracle == _ref.  The oracle restatement (oracle/orb_oracle.cc, match_oracle.cc, sbp_oracle.cc) against the
REFERENCE's own code compiled unchanged here (oracle/_ref/libref.so, recipe oracle/ref_build/Makefile):
src/ORBextractor.cc as a whole — constructor tables, ComputePyramid, ComputeKeyPointsOctTree, DistributeOctTree /
DivideNode, IC_Angle, computeOrbDescriptor, operator() with the lapping area — and ORBmatcher::DescriptorDistance /
ComputeThreeMaxima.  Rows A1, A4, A5, A6, A7, B1 of SURVEY 8(a) are therefore pinned by the reference itself, not by a
second restatement; the OpenCV primitives underneath are pinned separately against cv2 (test_oracle_orb.py)."""
import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R
from vieo_slam_b200.synth import texture

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")

CONFIGS = {  # name -> (ctor args, image)
    "euroc": ((1200, 1.2, 8, 20, 7), lambda: texture(480, 752, 7)),
    "euroc_dark": ((1200, 1.2, 8, 20, 7), lambda: texture(480, 752, 11, gain=0.35)),
    "euroc_sparse": ((1200, 1.2, 8, 20, 7), lambda: _sparse(480, 752)),
    "tumvi512": ((1000, 1.2, 8, 20, 7), lambda: texture(512, 512, 21)),
    "vr_scale2": ((187, 2.0, 4, 20, 7), lambda: texture(480, 640, 31)),
    "few_feats": ((375, 1.2, 4, 20, 7), lambda: texture(480, 752, 41)),
    "odd_size": ((500, 1.2, 8, 20, 7), lambda: texture(241, 323, 51)),
}


def _sparse(h, w):
    img = np.full((h, w), 90, np.uint8)
    r = np.random.default_rng(5)
    for _ in range(40):  # a few isolated blobs: most cells empty, minTh pass exercised, nodes with one point
        y, x = int(r.integers(30, h - 30)), int(r.integers(30, w - 30))
        img[y - 2:y + 3, x - 2:x + 3] = int(r.integers(130, 255))
    return img


@pytest.mark.parametrize("name", list(CONFIGS))
def test_ctor_tables_equal_reference(name):
    args, _ = CONFIGS[name]
    a, b = O.OrbOracle(*args).tables(), R.RefOrb(*args).tables()
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("lapping", [None, (100, 200), (0, 10000)])
def test_extract_equal_reference(name, lapping):
    args, mk = CONFIGS[name]
    img = mk()
    o, r = O.OrbOracle(*args), R.RefOrb(*args)
    n, kps, desc, mono = o.extract(img, lapping)
    rn, rkps, rdesc, rret = r.extract(img, lapping)
    for l in range(args[2]):
        assert np.array_equal(o.level(l), r.level(l)), f"pyramid level {l}"
    assert n == rn and mono == rret
    assert kps.tobytes() == rkps.tobytes(), "keypoints (order, position, size, angle, response, octave)"
    assert np.array_equal(desc, rdesc), "descriptors"


def test_empty_image_returns_minus_one():
    assert R.RefOrb().extract(None)[3] == -1
    assert O.OrbOracle().extract(np.zeros((0, 0), np.uint8))[0] == -1


def test_flat_image_no_keypoints():
    img = np.full((480, 752), 128, np.uint8)
    n, _, _, ret = R.RefOrb().extract(img)
    on, _, _, _ = O.OrbOracle().extract(img)
    assert n == 0 and on == 0 and ret == 0


@pytest.mark.parametrize("seed", range(6))
def test_quadtree_equal_reference_random_candidates(seed):
    # DistributeOctTree alone on synthetic candidate sets (dense, clustered, fewer than N, duplicates of position)
    r = np.random.default_rng(seed)
    W, H = (720, 448) if seed % 2 == 0 else (480, 480)
    n = [6000, 300, 40, 2500, 1, 900][seed]
    N = [261, 217, 181, 60, 10, 1000][seed]
    if seed == 3:  # clustered
        c = r.normal([W * 0.3, H * 0.6], 25, (n, 2))
        xy = np.clip(c, 0, [W - 1, H - 1]).astype(np.int32)
    else:
        xy = np.stack([r.integers(0, W, n), r.integers(0, H, n)], 1).astype(np.int32)
    xy = np.unique(xy, axis=0)
    # cell-major visiting order is irrelevant for the function itself; responses distinct so the per-node maximum is unique
    resp = r.permutation(len(xy)).astype(np.int32) + 7
    xyr = np.concatenate([xy, resp[:, None]], 1).astype(np.int32)
    # the reference takes region coordinates + the region bounds; the oracle entry point level coordinates + the level size
    ref = R.RefOrb(1200, 1.2, 8, 20, 7).quadtree(xyr, 16, 16 + W, 16, 16 + H, N)
    picked = O.quadtree(xyr + np.array([16, 16, 0], np.int32), W + 32, H + 32, N)
    # bit-equal INCLUDING the order: libref's monotonic allocator makes the reference's (size, node pointer) sort order
    # equal-size nodes by creation, which is the oracle's (and the kernel's) definition of that corner
    assert np.array_equal(xyr[picked], ref)
    assert len(ref) > 0


def test_ic_angle_and_descriptor_equal_reference():
    img = texture(200, 260, 77)
    blurred = O.gaussian_blur7(img)
    o = O.OrbOracle()
    r = R.RefOrb()
    # through the oracle's extractor on the same image: every keypoint's angle/descriptor must come out of the
    # reference's static functions given the same level image
    n, kps, desc, _ = O.OrbOracle(400, 1.2, 1, 20, 7).extract(img)
    assert n > 50
    for k, d in zip(kps, desc):
        a = r.ic_angle(img, k["x"], k["y"])
        assert np.float32(a).tobytes() == np.float32(k["angle"]).tobytes()
        assert np.array_equal(R.orb_descriptor(blurred, k["x"], k["y"], k["angle"]), d)


def test_descriptor_distance_equal_reference():
    r = np.random.default_rng(1)
    d = r.integers(0, 256, (400, 32), dtype=np.uint8)
    d[0] = 0
    d[1] = 255
    for i in range(0, 400, 2):
        a, b = d[i], d[i + 1]
        ref = R.descriptor_distance(a, b)
        assert ref == O.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
    assert R.descriptor_distance(d[0], d[1]) == 256


def test_three_maxima_equal_reference():
    # the rotation-histogram pick inside the oracle's SearchByProjection restatements follows ComputeThreeMaxima
    # (src/ORBmatcher.cc:1608-1641); pin the same function of the reference by known answers + a python restatement
    def py(s):
        m1 = m2 = m3 = 0
        i1 = i2 = i3 = -1
        for i, v in enumerate(s):
            if v > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, v, i2, i1, i
            elif v > m2:
                m3, m2, i3, i2 = m2, v, i2, i
            elif v > m3:
                m3, i3 = v, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        return i1, i2, i3

    assert R.three_maxima([0] * 30) == (-1, -1, -1)
    assert R.three_maxima([5] + [0] * 29) == (0, -1, -1)
    assert R.three_maxima([100, 9, 10] + [0] * 27) == (0, 2, -1)
    r = np.random.default_rng(2)
    for _ in range(300):
        s = r.integers(0, r.integers(1, 60), 30).astype(np.int32)
        assert R.three_maxima(s) == py(s)

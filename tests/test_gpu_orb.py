"""GPU parity: the CUDA ORB front-end and Hamming kernels against the oracle and the cv2 goldens, through
the C ABI.  Bit-exact (integer / index / byte work; fp32 angle and coordinates compared bitwise too)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import vieo_slam_b200.api as a
    return a


def _same_extract(api, img, nfeat, scale, nlev, lapping=None, check_stages=True):
    h, w = img.shape
    orb = api.ORBextractor(nfeat, scale, nlev, 20, 7, w, h, max_batch=2)
    ora = O.OrbOracle(nfeat, scale, nlev, 20, 7)
    ret, kps, desc = orb(img, pvLappingArea=lapping, want_pyramid=True)
    n, okps, odesc, omono = ora.extract(img, lapping=lapping)
    if check_stages:
        for l in range(nlev):
            assert np.array_equal(orb.mvImagePyramid[l], ora.level(l)), f"pyramid level {l}"
            assert np.array_equal(orb.debug_candidates(0, l), ora.candidates(l)), f"FAST candidates level {l}"
    assert len(kps) == n and ret == omono
    assert kps.tobytes() == okps.tobytes(), "keypoints (x,y,size,angle,response,octave) differ"
    assert np.array_equal(desc, odesc)
    orb.close()
    return n


def test_extract_matches_goldens_and_oracle(api, goldens):
    g = goldens
    n = _same_extract(api, g["A_img"], 1200, 1.2, 8)
    assert 1100 <= n <= 1300
    # the goldens pin the device pyramid / candidates to real cv2 output directly as well
    orb = api.ORBextractor(1200, 1.2, 8, 20, 7, 752, 480)
    orb(g["A_img"], want_pyramid=True)
    for l in (1, 4, 7):
        assert np.array_equal(orb.mvImagePyramid[l], g[f"A_L{l}"])
    for l in range(8):
        assert np.array_equal(orb.debug_candidates(0, l), g[f"A_cand{l}"].astype(np.int32))


def test_extract_matches_compiled_reference(api):
    """The CUDA extractor against the REFERENCE's own src/ORBextractor.cc compiled unchanged (oracle/_ref/libref.so, built
    in the container from /root/reference and shipped prebuilt): keypoints, order, angles, descriptors, lapping split."""
    import ref_lib as R
    if not R.available():
        pytest.skip("oracle/_ref/libref.so not present")
    from vieo_slam_b200.synth import texture
    for (nf, sc, nl), img, lap in [((1200, 1.2, 8), texture(480, 752, 7), None),
                                   ((1200, 1.2, 8), texture(480, 752, 11, gain=0.35), None),
                                   ((1000, 1.2, 8), texture(512, 512, 21), (0, 10000)),
                                   ((1000, 1.2, 8), texture(512, 512, 22), (100, 200)),
                                   ((187, 2.0, 4), texture(480, 640, 31), None)]:
        h, w = img.shape
        orb = api.ORBextractor(nf, sc, nl, 20, 7, w, h, max_batch=2)
        ret, kps, desc = orb(img, pvLappingArea=None if lap is None else list(lap), want_pyramid=True)
        ref = R.RefOrb(nf, sc, nl, 20, 7)
        rn, rkps, rdesc, rret = ref.extract(img, lap)
        assert len(kps) == rn and ret == rret
        assert kps.tobytes() == rkps.tobytes() and np.array_equal(desc, rdesc)
        for l in range(nl):
            assert np.array_equal(orb.mvImagePyramid[l], ref.level(l))
        orb.close()


def test_extract_small_dark_and_scale2(api, goldens):
    g = goldens
    _same_extract(api, g["B_img"], 300, 1.2, 4)
    _same_extract(api, g["B_img"], 300, 1.2, 4, lapping=[100, 200])
    _same_extract(api, g["C_img"], 375, 2.0, 3)


def test_extract_edge_inputs(api):
    from vieo_slam_b200.synth import texture
    flat = np.full((480, 752), 100, np.uint8)
    assert _same_extract(api, flat, 1200, 1.2, 8) == 0  # no corners at all
    sparse = flat.copy()
    sparse[100:104, 200:204] = 255
    sparse[300:303, 600:603] = 0
    _same_extract(api, sparse, 1200, 1.2, 8)  # fewer candidates than quota: every node single
    _same_extract(api, texture(480, 752, 3, gain=0.35), 1200, 1.2, 8)  # dark: minThFAST cells
    _same_extract(api, texture(512, 512, 4), 1000, 1.2, 8)  # TUM-VI shape, one quadtree root
    orb = api.ORBextractor(1200, 1.2, 8, 20, 7, 752, 480)
    assert orb(None)[0] == -1
    with pytest.raises(api.VieoError):
        api.ORBextractor(1200, 1.2, 8, 20, 7, 32, 32)  # below the supported minimum size


def test_batch_equals_single(api):
    from vieo_slam_b200.synth import stereo_stream
    imgs = stereo_stream(3, 21, dark_every=3)
    orb = api.ORBextractor(1200, 1.2, 8, 20, 7, 752, 480, max_batch=6)
    kps, desc, nk = orb.extract_batch(imgs)
    ora = O.OrbOracle(1200, 1.2, 8, 20, 7)
    for i in range(6):
        n, okps, odesc, _ = ora.extract(imgs[i])
        assert nk[i] == n
        assert kps[i, :n].tobytes() == okps.tobytes() and np.array_equal(desc[i, :n], odesc)


def test_hamming_knn2_and_csr(api, goldens):
    m = api.ORBmatcher()
    idx, dist = m.knnMatch2(goldens["bf_q"], goldens["bf_t"])
    assert np.array_equal(idx, goldens["bf_idx"]) and np.array_equal(dist, goldens["bf_dist"])
    rng = np.random.default_rng(1)
    for nq, nt in ((1200, 1200), (1, 1), (33, 2), (5, 1), (700, 1301)):
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        t[rng.integers(0, nt, nt // 3)] = q[rng.integers(0, nq)]  # many exact ties
        idx, dist = m.knnMatch2(q, t)
        oi, od = O.hamming_knn2(q, t)
        assert np.array_equal(dist, od) and np.array_equal(idx, oi), (nq, nt)
    # candidate lists: ragged, empty rows, duplicates
    q = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (900, 32), dtype=np.uint8) & 0x0F  # low-entropy -> frequent equal distances
    lens = rng.integers(0, 70, 500)
    lens[::7] = 0
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, 900, rp[-1]).astype(np.int32)
    got = m.search_candidates(q, t, rp, cand)
    ref = O.hamming_csr(q, t, rp, cand)
    for a, b, name in zip(got, ref, ("best_dist", "best_idx", "second_dist", "second_idx")):
        assert np.array_equal(a, b), name
    assert m.DescriptorDistance(q[0], t[0]) == O.descriptor_distance(q[0], t[0])


def test_stereo_frontend_host_api(api):
    """vieo_frontend_process (chunked, multi-stream, host buffers) == oracle per image and per pair."""
    from vieo_slam_b200.synth import stereo_stream
    F = 9  # 4 chunks of 3,3,3 frames (last chunk empty) -> exercises ragged chunking
    imgs = stereo_stream(F, 77, dark_every=4).reshape(F, 2, 480, 752)
    fe = api.StereoFrontend(1200, 1.2, 8, 20, 7, 752, 480, max_frames=F)
    outs = fe.alloc_outputs(F)
    kps, desc, nkp, midx, mdist = fe.process(imgs, outs)
    ora = O.OrbOracle(1200, 1.2, 8, 20, 7)
    for f in range(F):
        d = []
        for c in range(2):
            n, okps, odesc, _ = ora.extract(imgs[f, c])
            i = 2 * f + c
            assert nkp[i] == n
            assert kps[i, :n].tobytes() == okps.tobytes() and np.array_equal(desc[i, :n], odesc)
            d.append(odesc)
        oi, od = O.hamming_knn2(d[0], d[1])
        assert np.array_equal(midx[f, :len(oi)], oi) and np.array_equal(mdist[f, :len(oi)], od)
    assert fe.last_launches() == 3 * 5  # per chunk: pyramid, FAST, quadtree, orientation + descriptors, knnMatch


def test_rectified_stereo_matches_bit_exact(api):
    """vieo_frontend_stereo_rectified (Frame::ComputeStereoMatches on the device) == oracle: uright / depth bit-exact,
    SADs identical, same matches dropped by the median filter."""
    from vieo_slam_b200.synth import EUROC, stereo_stream
    F = 5
    imgs = stereo_stream(F, 91, dark_every=3).reshape(F, 2, 480, 752)
    fe = api.StereoFrontend(1200, 1.2, 8, 20, 7, 752, 480, max_frames=F)
    outs = fe.alloc_outputs(F)
    kps, desc, nkp, _, _ = fe.process(imgs, outs)
    bf = np.float32(EUROC["bf"]); minZ = np.float32(bf / np.float32(EUROC["fx"]))
    ur, dp, sad = fe.stereo_rectified(F, bf, minZ)
    for f in range(F):
        oL, oR = O.OrbOracle(1200, 1.2, 8, 20, 7), O.OrbOracle(1200, 1.2, 8, 20, 7)
        nl, kl, dl, _ = oL.extract(imgs[f, 0]); nr, kr, dr, _ = oR.extract(imgs[f, 1])
        our, odp, osad, kept = O.stereo_matches(oL, kl, dl, oR, kr, dr, bf, minZ)
        assert kept > 100
        assert np.array_equal(sad[f, :nl], osad)
        assert ur[f, :nl].tobytes() == our.tobytes() and dp[f, :nl].tobytes() == odp.tobytes()
        assert np.all(ur[f, nl:] == -1)

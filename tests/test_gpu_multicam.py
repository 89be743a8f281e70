"""GPU parity for BASELINE configs[3] (TUM-VI 4 x 512 x 512 KB8 rig): per-camera extraction with the lapping area and the
brute-force half of Frame::ComputeStereoFishEyeMatches for all 6 camera pairs — bit-exact against the oracle (whose
extractor is pinned to the compiled reference and whose knnMatch to cv2.BFMatcher)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _rig_images(n_frames, n_cams, seed, size=512, dark_cam=None):
    """n_cams overlapping views of one large texture per frame (neighbouring cameras share ~60 % of their field)."""
    from vieo_slam_b200.synth import texture
    big = texture(size + 64, size * 3, seed)
    out = np.empty((n_frames * n_cams, size, size), np.uint8)
    for f in range(n_frames):
        for c in range(n_cams):
            x = 40 + c * int(size * 0.4) + 7 * f
            v = big[10 + 3 * f:10 + 3 * f + size, x:x + size]
            if dark_cam == c:
                v = (v.astype(np.float32) * 0.35).astype(np.uint8)
            out[f * n_cams + c] = v
    return out


def _oracle_frames(imgs, n_cams, nfeat, lapping, cap):
    ora = O.OrbOracle(nfeat, 1.2, 8, 20, 7)
    n_img = len(imgs)
    kps = np.zeros((n_img, cap), O.KP_DTYPE); desc = np.zeros((n_img, cap, 32), np.uint8)
    nk = np.zeros(n_img, np.int32); nm = np.zeros(n_img, np.int32)
    for i in range(n_img):
        lap = None if lapping is None else lapping[i % n_cams]
        n, k, d, mono = ora.extract(imgs[i], lapping=lap)
        kps[i, :n] = k; desc[i, :n] = d; nk[i] = n; nm[i] = mono
    pairs = [O.fisheye_matches(desc[f * n_cams:(f + 1) * n_cams], nk[f * n_cams:(f + 1) * n_cams],
                               nm[f * n_cams:(f + 1) * n_cams]) for f in range(n_img // n_cams)]
    return kps, desc, nk, nm, pairs


@pytest.mark.parametrize("lapping", [[(0, 10000)] * 4, [(100, 400), (50, 300), (0, 10000), (600, 700)], None])
def test_four_camera_frames_match_oracle(lapping):
    import vieo_slam_b200.api as api
    n_cams, n_frames = 4, 2
    imgs = _rig_images(n_frames, n_cams, 411, dark_cam=2)
    orb = api.ORBextractor(1000, 1.2, 8, 20, 7, 512, 512, max_batch=n_cams * n_frames)
    out = orb.multicam_frames(imgs, n_cams, lapping)
    kps, desc, nk, nm, pairs = _oracle_frames(imgs, n_cams, 1000, lapping, orb.cap)
    assert np.array_equal(out["n_kp"], nk) and np.array_equal(out["n_mono"], nm)
    for i in range(len(imgs)):
        n = nk[i]
        assert out["kps"][i, :n].tobytes() == kps[i, :n].tobytes(), f"image {i}: keypoint order / values"
        assert np.array_equal(out["desc"][i, :n], desc[i, :n])
    n_good = 0
    for f in range(n_frames):
        idx, dist, good = pairs[f]
        assert np.array_equal(out["pair_idx"][f], idx)
        assert np.array_equal(out["pair_dist"][f], dist)
        assert np.array_equal(out["pair_good"][f], good)
        n_good += int(good.sum())
    assert n_good > 50  # overlapping views do produce accepted matches
    if lapping is not None and lapping[3] == (600, 700):
        assert (out["n_mono"][3::4] == out["n_kp"][3::4]).all()  # camera 3: empty lapping area -> its pairs are skipped
        assert (out["pair_idx"][:, [2, 4, 5]] == -1).all()


def test_two_camera_ties_and_small_sets():
    """Pair matching alone on crafted descriptors: ties -> lowest train index, one-row train set (size < 2: never good),
    the 0.7 / 0.9 ratio boundaries in the reference's float-vs-double arithmetic."""
    import ctypes as C
    import torch
    import vieo_slam_b200.api as api
    r = np.random.default_rng(9)
    cap = 64
    desc = r.integers(0, 256, (3, 3, cap, 32), dtype=np.uint8)  # 3 frames x 3 cameras
    nk = np.array([[40, 40, 1], [10, 64, 64], [0, 5, 5]], np.int32)
    nm = np.array([[0, 8, 0], [3, 0, 60], [0, 0, 5]], np.int32)
    desc[0, 1, 8:40] = desc[0, 0, :32]          # exact matches, distance 0
    desc[0, 1, 20] = desc[0, 1, 21]             # tie between two train rows
    for k in range(10):                         # d0 / d1 ratios around 0.7 and 0.9 with d0 below / above 75
        desc[1, 1, k] = 0
        desc[1, 1, k + 10] = 0
    dd = torch.from_numpy(desc.copy()).cuda()
    dnk = torch.from_numpy(nk.reshape(-1).copy()).cuda(); dnm = torch.from_numpy(nm.reshape(-1).copy()).cuda()
    idx = torch.empty((3, 3, cap, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
    good = torch.empty((3, 3, cap), dtype=torch.uint8, device="cuda")
    api._check(api.lib().vieo_fisheye_knn_dev(dd.data_ptr(), dnk.data_ptr(), dnm.data_ptr(), 3, 3, cap, 3, 1, idx.data_ptr(),
                                              dist.data_ptr(), good.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for f in range(3):
        oi, od, og = O.fisheye_matches(desc[f], nk[f], nm[f])
        assert np.array_equal(idx[f].cpu().numpy(), oi) and np.array_equal(dist[f].cpu().numpy(), od)
        assert np.array_equal(good[f].cpu().numpy(), og)
    assert good[0, 1].sum().item() == 0  # camera 2 of frame 0 has ONE row: knnMatch returns a single neighbour


def test_camera_major_layout_equals_frame_major():
    """[camera][frame] shards (what an all-gather over the per-camera GPUs of a rig yields) give the same pair results."""
    import torch
    import vieo_slam_b200.api as api
    r = np.random.default_rng(10)
    cap, F, Cn = 96, 5, 4
    desc = r.integers(0, 256, (F, Cn, cap, 32), dtype=np.uint8)
    nk = r.integers(20, cap + 1, (F, Cn)).astype(np.int32); nm = r.integers(0, 20, (F, Cn)).astype(np.int32)
    outs = []
    for layout in ("fc", "cf"):
        d = desc if layout == "fc" else np.ascontiguousarray(desc.transpose(1, 0, 2, 3))
        k = nk if layout == "fc" else np.ascontiguousarray(nk.T)
        m = nm if layout == "fc" else np.ascontiguousarray(nm.T)
        dd, dk, dm = (torch.from_numpy(x.copy()).cuda() for x in (d, k.reshape(-1), m.reshape(-1)))
        idx = torch.empty((F, 6, cap, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
        good = torch.empty((F, 6, cap), dtype=torch.uint8, device="cuda")
        fs, cs = (Cn, 1) if layout == "fc" else (1, F)
        api._check(api.lib().vieo_fisheye_knn_dev(dd.data_ptr(), dk.data_ptr(), dm.data_ptr(), Cn, F, cap, fs, cs, idx.data_ptr(),
                                                  dist.data_ptr(), good.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        outs.append((idx.cpu().numpy(), dist.cpu().numpy(), good.cpu().numpy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    oi, od, og = O.fisheye_matches(desc[2], nk[2], nm[2])
    assert np.array_equal(outs[0][0][2], oi) and np.array_equal(outs[0][2][2], og)

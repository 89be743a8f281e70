#!/usr/bin/env python3
"""Generate the committed OpenCV goldens (tests/golden/cv2_orb_goldens.npz).

Runs in the BUILD container only (python cv2 4.13.0, the one OpenCV available offline).  Every OpenCV
primitive on the reference's ORB path is executed here through cv2 and stored with its input, so that
the oracle (oracle/orb_oracle.cc) is pinned to real OpenCV output without cv2 at test time:
  * cv::resize INTER_LINEAR chain     (src/ORBextractor.cc:1070)
  * per-cell cv::FAST ini/min fallback (src/ORBextractor.cc:738-779), composed exactly like the reference loop
  * cv::GaussianBlur 7x7 s=2 REFLECT_101 (src/ORBextractor.cc:1013)
  * cv::fastAtan2                      (src/ORBextractor.cc:79)
  * cv::BFMatcher(NORM_HAMMING).knnMatch k=2 (src/Frame.cc:625)
"""
import os, sys
import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vieo_slam_b200.synth import texture

cv2.setNumThreads(1)
HERE = os.path.dirname(os.path.abspath(__file__))


def level_sizes(w, h, scale, nlevels):
    sc = np.float32(1.0)
    out = [(w, h)]
    for _ in range(1, nlevels):
        sc = np.float32(np.float64(sc) * np.float64(np.float32(scale)))
        inv = np.float32(1.0) / sc
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def cell_fast(img, ini_th, min_th):
    """The reference's per-cell loop with real cv2.FAST.  Returns (n,3) int32 (x, y, response) in level coords."""
    H, W = img.shape
    minB, maxBX, maxBY = 16, W - 16, H - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
    if nCols <= 0 or nRows <= 0:
        return np.zeros((0, 3), np.int32)
    wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
    fi = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fm = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    out = []
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cell = img[iniY:maxY, iniX:maxX]
            k = fi.detect(cell)
            if len(k) == 0:
                k = fm.detect(cell)
            for p in k:
                out.append((int(p.pt[0]) + j * wCell + minB, int(p.pt[1]) + i * hCell + minB, int(p.response)))
    return np.array(out, np.int32).reshape(-1, 3)


def main():
    g = {}
    # image A: full EuRoC size, normal contrast.  image B: small, with a dark half (min-threshold cells)
    A = texture(480, 752, 11)
    B = texture(190, 260, 12)
    B[:, 130:] = (B[:, 130:].astype(np.float32) * 0.25 + 40).astype(np.uint8)
    B[60:120, 20:100] = 90  # flat area -> empty cells at both thresholds
    for name, img, scale, nlev in (("A", A, 1.2, 8), ("B", B, 1.2, 4), ("C", A[:240, :376].copy(), 2.0, 3)):
        g[f"{name}_img"] = img
        g[f"{name}_scale"] = np.float32(scale)
        cur = img
        for l, (w, h) in enumerate(level_sizes(img.shape[1], img.shape[0], scale, nlev)):
            if l:
                cur = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
                if name != "A" or l in (1, 4, 7):
                    g[f"{name}_L{l}"] = cur
            g[f"{name}_crc{l}"] = np.array([int(cur.astype(np.uint64).sum()), int((cur.astype(np.uint64) * (np.arange(cur.size, dtype=np.uint64).reshape(cur.shape) % 251)).sum())], np.uint64)
            g[f"{name}_cand{l}"] = cell_fast(cur, 20, 7).astype(np.int16)
            if name != "A" or l in (0, 5):
                g[f"{name}_blur{l}"] = cv2.GaussianBlur(cur, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    r = np.random.default_rng(5)
    yx = r.integers(-40000, 40000, (4000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [3, -3]]
    g["atan_in"] = yx
    g["atan_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    # brute-force Hamming k=2
    q = r.integers(0, 256, (300, 32), dtype=np.uint8)
    t = r.integers(0, 256, (257, 32), dtype=np.uint8)
    t[5] = q[7]; t[9] = q[7]  # exact duplicates -> tie on distance 0
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    g["bf_q"], g["bf_t"] = q, t
    g["bf_idx"] = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    g["bf_dist"] = np.array([[a.distance, b.distance] for a, b in m], np.int32)
    out = os.path.join(HERE, "cv2_orb_goldens.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB; cv2", cv2.__version__)


if __name__ == "__main__":
    main()

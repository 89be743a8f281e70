#!/usr/bin/env python3
"""BASELINE.md §2 cross-check: the OpenCV primitives the reference's ORB path calls (cv::resize, per-cell cv::FAST,
cv::GaussianBlur, BFMatcher::knnMatch), timed single-threaded through python cv2 next to the oracle's restatements of
the same primitives on the same 752x480 input — evidence that the CPU baseline (`cpu_baseline.kind = "port"`) is not an
unfairly slow stand-in for the reference's OpenCV calls.  CPU only; results are identical by the golden tests, only the
time is compared.   python tools/cpu_primitives_crosscheck.py [out.md]"""
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from vieo_slam_b200.synth import texture  # noqa: E402


def best_of(fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return 1e3 * min(t)


def level_sizes():
    """EuRoC 752x480, scale 1.2, 8 levels with the reference's float arithmetic."""
    return [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231), (302, 193), (252, 161), (210, 134)]


def cells(w, h):
    """ComputeKeyPointsOctTree's cell rectangles (src/ORBextractor.cc:723-760)."""
    minx, miny, maxx, maxy = 16 - 3, 16 - 3, w - 16 + 3, h - 16 + 3
    W, H = maxx - minx, maxy - miny
    nc, nr = W // 35, H // 35
    wc, hc = int(np.ceil(W / nc)), int(np.ceil(H / nr))
    out = []
    for i in range(nr):
        y0 = miny + i * hc
        y1 = min(y0 + hc + 6, maxy)
        if y0 >= maxy - 3:
            continue
        for j in range(nc):
            x0 = minx + j * wc
            x1 = min(x0 + wc + 6, maxx)
            if x0 >= maxx - 6:
                continue
            out.append((x0, y0, x1, y1))
    return out


def main(out_path=None):
    cv2.setNumThreads(1)
    img = texture(480, 752, 7)
    sizes = level_sizes()
    pyr = [img]
    for (w, h) in sizes[1:]:
        pyr.append(cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR))
    rows = []

    def pyr_cv():
        p = img
        for (w, h) in sizes[1:]:
            p = cv2.resize(p, (w, h), interpolation=cv2.INTER_LINEAR)

    def pyr_or():
        p = img
        for (w, h) in sizes[1:]:
            p = O.resize_linear(p, w, h)
    rows.append(("pyramid: 7 x resize INTER_LINEAR", best_of(pyr_cv), best_of(pyr_or)))

    det = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det7 = cv2.FastFeatureDetector_create(7, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    rects = [[(x0, y0, x1, y1) for (x0, y0, x1, y1) in cells(l.shape[1], l.shape[0])] for l in pyr]
    ncell = sum(len(r) for r in rects)

    def fast_cv():
        n = 0
        for l, rs in zip(pyr, rects):
            for (x0, y0, x1, y1) in rs:
                c = l[y0:y1, x0:x1]
                k = det.detect(c)
                if not k:
                    k = det7.detect(c)
                n += len(k)
        return n

    def fast_or():
        n = 0
        for l, rs in zip(pyr, rects):
            for (x0, y0, x1, y1) in rs:
                c = l[y0:y1, x0:x1]
                k = O.fast_detect(c, 20)
                if len(k) == 0:
                    k = O.fast_detect(c, 7)
                n += len(k)
        return n
    assert fast_cv() == fast_or()
    rows.append((f"per-cell FAST-9/16 + NMS, ini 20 -> min 7 ({ncell} cells, 8 levels; {ncell} python calls on either side, the "
                 "oracle's ctypes wrapper allocating its outputs per call)", best_of(fast_cv, 3), best_of(fast_or, 3)))

    def fastl_cv():
        return sum(len(det.detect(l)) for l in pyr)

    def fastl_or():
        return sum(len(O.fast_detect(l, 20)) for l in pyr)
    assert fastl_cv() == fastl_or()
    rows.append(("FAST-9/16 + NMS th 20 on whole levels (8 calls)", best_of(fastl_cv, 3), best_of(fastl_or, 3)))

    def blur_cv():
        for l in pyr:
            cv2.GaussianBlur(l, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)

    def blur_or():
        for l in pyr:
            O.gaussian_blur7(l)
    rows.append(("GaussianBlur 7x7 sigma 2, 8 levels", best_of(blur_cv), best_of(blur_or)))

    rng = np.random.default_rng(1)
    q = rng.integers(0, 256, (1200, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1200, 32), dtype=np.uint8)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    rows.append(("BFMatcher(NORM_HAMMING).knnMatch k=2, 1200 x 1200", best_of(lambda: bf.knnMatch(q, t, 2), 3),
                 best_of(lambda: O.hamming_knn2(q, t), 3)))

    orb = O.OrbOracle(1200, 1.2, 8, 20, 7)
    full = best_of(lambda: orb.extract(img), 3)
    lines = ["CPU cross-check of the oracle's OpenCV restatements against python cv2 %s (cv2.setNumThreads(1)), one 752x480 "
             "image, best of 3-5, this container's CPU (%d logical cores); results are bit-identical (tests/test_oracle_orb.py), "
             "only time is compared\n" % (cv2.__version__, os.cpu_count()),
             "| primitive | cv2 ms | oracle ms | oracle / cv2 |", "|---|---|---|---|"]
    for name, a, b in rows:
        lines.append(f"| {name} | {a:.2f} | {b:.2f} | {b / a:.2f} |")
    cv_sum = rows[0][1] + rows[1][1] + rows[3][1]
    lines.append(f"\ncv2 pyramid + per-cell FAST + blur alone: {cv_sum:.1f} ms per image (no quadtree, orientation or descriptors). "
                 "The stand-alone `orc_fast_detect` timed above is the simple per-pixel pin of cv::FAST used by the golden tests; the "
                 "oracle's extractor scores each level once; its whole pipeline costs about what the three cv2 primitives cost alone:")
    lines.append(f"ORBextractor::operator() restated by the oracle (pyramid + FAST + quadtree + orientation + blur + descriptors), "
                 f"one image: {full:.1f} ms — the reference reports 35-43 ms per stereo FRAME for its whole front-end "
                 f"(README.md:60, i9-14900HX, one thread per camera).")
    txt = "\n".join(lines) + "\n"
    print(txt)
    if out_path:
        open(out_path, "w").write(txt)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)

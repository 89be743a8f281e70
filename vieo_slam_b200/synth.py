"""Synthetic EuRoC-shaped inputs: 752x480 u8 stereo frames, 200 Hz IMU, PoseOptimization
problems.  numpy only; seeds fixed by the caller.  Used by tests and bench.py (no dataset is available
offline)."""
import numpy as np

EUROC = dict(w=752, h=480, fx=435.2047, fy=435.2047, cx=367.4517, cy=252.2009, bf=47.9064,
             nfeatures=1200, scale=1.2, nlevels=8, ini_th=20, min_th=7)


def texture(h, w, seed, sigma=40.0, mean=110.0, gain=1.0, alpha=1.0):
    """Band-limited (1/f^alpha) noise texture, u8.  FAST@20 yields >=5k level-0 candidates at 752x480."""
    r = np.random.default_rng(seed)
    f = np.fft.rfft2(r.standard_normal((h, w)))
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    k = np.sqrt(fx * fx + fy * fy)
    k[0, 0] = 1.0
    im = np.fft.irfft2(f / k ** alpha, (h, w))
    im = (im - im.mean()) / im.std() * sigma + mean
    return np.clip(im * gain, 0, 255).astype(np.uint8)


def stereo_stream(n_frames, seed, w=752, h=480, max_disp=40, dark_every=0):
    """n_frames stereo pairs (2*n_frames, h, w) u8: a large texture panned along a smooth path; the right
    view is the left shifted by a per-frame disparity (fronto-parallel plane).  Every `dark_every`-th
    frame is low-contrast (gain 0.4) to exercise the minThFAST fallback."""
    big = texture(h + 256, w + 512 + max_disp, seed)
    out = np.empty((2 * n_frames, h, w), np.uint8)
    r = np.random.default_rng(seed + 1)
    ph = r.uniform(0, 2 * np.pi, 3)
    for f in range(n_frames):
        ox = int(128 + 120 * np.sin(0.05 * f + ph[0])) + max_disp
        oy = int(128 + 100 * np.sin(0.037 * f + ph[1]))
        d = int(10 + (max_disp - 12) * (0.5 + 0.5 * np.sin(0.02 * f + ph[2])))
        L = big[oy:oy + h, ox:ox + w]
        R = big[oy:oy + h, ox + d:ox + d + w]  # scene point at uL appears at uR = uL - d
        if dark_every and f % dark_every == dark_every - 1:
            L = (L.astype(np.float32) * 0.4).astype(np.uint8)
            R = (R.astype(np.float32) * 0.4).astype(np.uint8)
        out[2 * f] = L
        out[2 * f + 1] = R
    return out

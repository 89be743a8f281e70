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


# ------------------------------------------------------------------------------------------------------------
# Synthetic visual-inertial sequences and BA problems: smooth 6-DoF trajectory,
# 200 Hz IMU with EuRoC noise (Examples/Stereo/EuRoC/EuRoC_VIO.yaml:13-18), pinhole stereo observations
# quantised to float like cv::KeyPoint.  All numpy, seeded.
from .layouts import (CAM_KB8, CAM_RADTAN, CAMERA_DTYPE, EDGE_CLOSE, EDGE_STEREO, NAVSTATE_DTYPE,  # noqa: E402
                      POSEOPT_PROBLEM_DTYPE)

EUROC_IMU_SIGMA = (1.6968e-4, 2.0e-3, 1.9393e-5, 3.0e-3)
# Camera.Tbc of EuRoC_VIO.yaml:24-28 (body <- camera)
EUROC_TBC = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
                      [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
                      [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
                      [0, 0, 0, 1.0]])
GRAVITY_W = np.array([0.0, 0.0, -9.81])


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def so3_exp(w):
    th = np.linalg.norm(w)
    K = hat(w)
    if th < 1e-9:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def so3_log(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return w if th < 1e-9 else w * th / np.sin(th)


def quat_from_R(R):
    """(w, x, y, z), w >= 0"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q /= np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def R_from_quat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def euroc_camera():
    cam = np.zeros(1, CAMERA_DTYPE)[0]
    cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"] = EUROC["fx"], EUROC["fy"], EUROC["cx"], EUROC["cy"], EUROC["bf"]
    Tbc = EUROC_TBC.astype(np.float32).astype(np.float64)  # Camera.Tbc is read into CV_32F (src/Tracking.cc:704-732)
    Rcb = Tbc[:3, :3].T
    cam["Rcb"] = Rcb
    cam["tcb"] = -Rcb @ Tbc[:3, 3]
    return cam


def radtan_camera(k=(-0.28340811, 0.07395907), p=(0.00019359, 1.76187114e-05)):
    """EuRoC cam0 with its raw radial-tangential distortion (Examples/Stereo/EuRoC/EuRoC_dist*.yaml style)."""
    cam = euroc_camera()
    cam["model"], cam["num_k"] = CAM_RADTAN, len(k)
    cam["dist"][:len(k)] = k
    cam["dist"][len(k):len(k) + 2] = p
    return cam


def kb8_camera(k=(0.0034823894, 0.0007150348, -0.0020532361, 0.00020293673)):
    """Kannala-Brandt fisheye with TUM-VI-like coefficients (Examples/Stereo/TUM_VI/TUM_VI_512_VIO.yaml)."""
    cam = euroc_camera()
    cam["model"] = CAM_KB8
    cam["dist"][:4] = k
    return cam


class Trajectory:
    """Smooth body trajectory in a ~6x6x3 m room: p(t) sums of sinusoids, R(t) = Exp(theta(t))."""

    def __init__(self, seed, speed=1.0, rot=0.5):
        r = np.random.default_rng(seed)
        self.A = np.array([2.0, 2.0, 0.6]) * r.uniform(0.6, 1.0, 3)
        self.w = speed * r.uniform(0.25, 0.5, 3)
        self.ph = r.uniform(0, 2 * np.pi, 3)
        self.B = rot * r.uniform(0.3, 0.6, 3)
        self.u = r.uniform(0.2, 0.45, 3)
        self.qh = r.uniform(0, 2 * np.pi, 3)

    def p(self, t):
        return self.A * np.sin(self.w * t + self.ph)

    def v(self, t):
        return self.A * self.w * np.cos(self.w * t + self.ph)

    def acc(self, t):
        return -self.A * self.w ** 2 * np.sin(self.w * t + self.ph)

    def R(self, t):
        return so3_exp(self.B * np.sin(self.u * t + self.qh))

    def omega(self, t, h=1e-5):
        return so3_log(self.R(t - h).T @ self.R(t + h)) / (2 * h)


def vio_sequence(seed, n_frames, frame_rate=20.0, imu_rate=200.0, speed=1.0, rot=0.5, noisy_imu=True):
    """Returns dict(times[n], truth NAVSTATE[n], imu (m,7) rows {t, a, w}, bg, ba, traj)."""
    traj = Trajectory(seed, speed, rot)
    r = np.random.default_rng(seed + 7)
    times = np.arange(n_frames) / frame_rate
    m = int(np.ceil(times[-1] * imu_rate)) + 3
    ti = np.arange(m) / imu_rate
    bg = r.normal(0, 2e-3, 3)
    ba = r.normal(0, 2e-2, 3)
    imu = np.zeros((m, 7))
    imu[:, 0] = ti
    sg, sa = EUROC_IMU_SIGMA[0] * np.sqrt(imu_rate), EUROC_IMU_SIGMA[1] * np.sqrt(imu_rate)
    for k, t in enumerate(ti):
        R = traj.R(t)
        imu[k, 1:4] = R.T @ (traj.acc(t) - GRAVITY_W) + ba
        imu[k, 4:7] = traj.omega(t) + bg
    if noisy_imu:
        imu[:, 1:4] += r.normal(0, sa, (m, 3))
        imu[:, 4:7] += r.normal(0, sg, (m, 3))
    truth = np.zeros(n_frames, NAVSTATE_DTYPE)
    for k, t in enumerate(times):
        truth[k]["p"] = traj.p(t)
        truth[k]["q"] = quat_from_R(traj.R(t))
        truth[k]["v"] = traj.v(t)
        truth[k]["bg"] = bg
        truth[k]["ba"] = ba
    return dict(times=times, truth=truth, imu=imu, bg=bg, ba=ba, traj=traj, seed=seed)


def perturb_state(ns, r, dp=0.01, drot=np.deg2rad(0.3), dv=0.02, dbg=1e-3, dba=1e-2):
    out = ns.copy()
    out["p"] = ns["p"] + r.normal(0, dp, 3)
    out["q"] = quat_from_R(R_from_quat(ns["q"]) @ so3_exp(r.normal(0, drot, 3)))
    out["v"] = ns["v"] + r.normal(0, dv, 3)
    out["bg"] = ns["bg"] + r.normal(0, dbg, 3)
    out["ba"] = ns["ba"] + r.normal(0, dba, 3)
    return out


def distort(cam, x, y):
    """Normalised image-plane coordinates of camera-frame directions (x, y, 1)·z through the camera's lens model
    (pinhole: identity; radtan camera_radtan.h:61-129; KB8 camera_kb8.h:68-157); x, y are Pc.x/Pc.z, Pc.y/Pc.z."""
    model = int(cam["model"])
    if model == CAM_RADTAN:
        nk = int(cam["num_k"])
        k = cam["dist"][:nk].astype(np.float64); p = cam["dist"][nk:nk + 2].astype(np.float64)
        r2 = x * x + y * y
        fd = 1 + sum(k[i] * r2 ** (i + 1) for i in range(nk))
        return (x * fd + 2 * p[0] * x * y + p[1] * (r2 + 2 * x * x), y * fd + 2 * p[1] * x * y + p[0] * (r2 + 2 * y * y))
    if model == CAM_KB8:
        k1, k2, k3, k4 = cam["dist"][:4].astype(np.float64)
        r = np.sqrt(x * x + y * y)
        th = np.arctan(r)
        t2 = th * th
        thd = th * (1 + t2 * (k1 + t2 * (k2 + t2 * (k3 + t2 * k4))))
        s = np.where(r > 1e-5, thd / np.maximum(r, 1e-300), 1.0)
        return x * s, y * s
    return x, y


def project(cam, ns, X):
    """Stereo projection of world points X (n,3) through body state ns and the camera's lens model: (u, v, ur, z)."""
    Rwb = R_from_quat(ns["q"])
    Rcw = cam["Rcb"] @ Rwb.T
    tcw = -Rcw @ ns["p"] + cam["tcb"]
    Pc = X @ Rcw.T + tcw
    z = Pc[:, 2]
    xd, yd = distort(cam, Pc[:, 0] / z, Pc[:, 1] / z)
    u = cam["fx"] * xd + cam["cx"]
    v = cam["fy"] * yd + cam["cy"]
    return u, v, u - cam["bf"] / z, z


def landmarks_in_view(cam, ns, n, r, zmin=0.8, zmax=9.0):
    """n world points visible from state ns (uniform in the image, depth uniform in [zmin, zmax])."""
    u = r.uniform(20, EUROC["w"] - 20, n)
    v = r.uniform(20, EUROC["h"] - 20, n)
    z = r.uniform(zmin, zmax, n)
    Pc = np.stack([(u - cam["cx"]) / cam["fx"] * z, (v - cam["cy"]) / cam["fy"] * z, z], 1)
    Rwb = R_from_quat(ns["q"])
    Rcw = cam["Rcb"] @ Rwb.T
    tcw = -Rcw @ ns["p"] + cam["tcb"]
    return (Pc - tcw) @ Rcw  # Rcw^T (Pc - tcw)


def inv_level_sigma2(nlevels=8, scale=1.2):
    s = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        s[i] = s[i - 1] * np.float32(scale)
    return (np.float32(1.0) / (s * s)).astype(np.float32), s


def make_observations(cam, ns, X, r, outlier_frac=0.15, stereo_frac=0.7, th_depth=35.0):
    """Noisy float keypoint observations of X from ns: obs (n,3) f32 (ur = -1 for mono), inv_sigma2 f32, flags u8."""
    n = len(X)
    inv_s2, scl = inv_level_sigma2()
    u, v, ur, z = project(cam, ns, X)
    octave = r.integers(0, 8, n)
    sig = scl[octave].astype(np.float64)
    obs = np.stack([u + r.normal(0, 1, n) * sig, v + r.normal(0, 1, n) * sig, ur + r.normal(0, 1, n) * sig], 1)
    out = r.random(n) < outlier_frac
    obs[out, :2] += r.uniform(-20, 20, (int(out.sum()), 2))
    obs[out, 2] = obs[out, 0] - (u - ur)[out] + r.uniform(-10, 10, int(out.sum()))
    stereo = r.random(n) < stereo_frac
    obs[~stereo, 2] = -1.0
    flags = np.where(stereo, EDGE_STEREO, 0).astype(np.uint8)
    thd = max(10.0, cam["bf"] / cam["fx"] * th_depth)
    flags |= np.where(z < thd, EDGE_CLOSE, 0).astype(np.uint8)
    return obs.astype(np.float32), inv_s2[octave], flags, out


def make_pose_problems(seq, preints, cam, n_points=450, seed=0, mode=1, compute_marg=True, chain_prior=False,
                       outlier_frac=0.15):
    """One PoseOptimization problem per consecutive frame pair (k-1 -> k).  preints[k] = pre-integration over
    (t_{k-1}, t_k] (PREINT record; dt == 0 -> no IMU edge).  Returns (problems, Xw, obs, inv_sigma2, flags)."""
    r = np.random.default_rng(seed + 1000)
    n = len(seq["times"])
    pbs = np.zeros(n - 1, POSEOPT_PROBLEM_DTYPE)
    Xs, Os, Ws, Fs = [], [], [], []
    e0 = 0
    sig = EUROC_IMU_SIGMA
    for k in range(1, n):
        pb = pbs[k - 1]
        tru, last = seq["truth"][k], seq["truth"][k - 1]
        pb["cur"] = perturb_state(tru, r, dbg=0, dba=0)
        pb["last"] = perturb_state(last, r, dp=0.002, drot=np.deg2rad(0.05), dv=0.005, dbg=0, dba=0)
        pb["prior"] = pb["last"]
        pb["preint"] = preints[k]
        A = r.normal(0, 1, (15, 15))
        pb["prior_info"] = A @ A.T * 10 + np.diag([1e4] * 3 + [1e3] * 3 + [1e5] * 3 + [1e6] * 3 + [1e4] * 3)
        pb["gw"] = GRAVITY_W
        pb["inv_sigma_bg2"] = 1.0 / sig[2] ** 2
        pb["inv_sigma_ba2"] = 1.0 / sig[3] ** 2
        pb["dt_frames"] = seq["times"][k] - seq["times"][k - 1]
        pb["mode"] = mode
        pb["last_has_prior"] = int(chain_prior and k % 2 == 0)
        pb["compute_marg"] = int(compute_marg)
        X = landmarks_in_view(cam, tru, n_points, r)
        obs, w, fl, _ = make_observations(cam, tru, X, r, outlier_frac)
        Xf = X.astype(np.float32).astype(np.float64)  # MapPoint positions are float (include/MapPoint.h:52)
        pb["edge_begin"], pb["edge_end"] = e0, e0 + n_points
        e0 += n_points
        Xs.append(Xf); Os.append(obs); Ws.append(w); Fs.append(fl)
    return pbs, np.concatenate(Xs), np.concatenate(Os), np.concatenate(Ws), np.concatenate(Fs)


def make_lba_problem(seq, preints_kf, kf_idx, cam, n_local=10, n_fixed=20, n_points=1500, obs_per_point=(4, 8), seed=0,
                     outlier_frac=0.05):
    """A LocalBundleAdjustmentNavStatePRV window over keyframes kf_idx (frame indices, ascending): the last n_local
    are free, the one before is the fixed 'previous' KF (with V/Bias vertices), `n_fixed` earlier ones are
    covisible fixed KFs (PR only).  preints_kf[i] = pre-integration from kf_idx[i-1] to kf_idx[i].
    Returns dict of arrays laid out as the C ABI wants (edges sorted by point)."""
    r = np.random.default_rng(seed + 2000)
    kf_idx = np.asarray(kf_idx)
    nk = len(kf_idx)
    n_local = min(n_local, nk)
    local = list(range(nk - n_local, nk))
    prev = [nk - n_local - 1] if nk - n_local - 1 >= 0 else []
    fixed = list(range(max(0, nk - n_local - 1 - n_fixed), nk - n_local - 1))
    order = local + prev + fixed  # states: local first (ascending id), then fixed (reference adds lFixedCameras after)
    states = np.zeros(len(order), NAVSTATE_DTYPE)
    sflags = np.zeros(len(order), np.uint8)
    for s, k in enumerate(order):
        tru = seq["truth"][kf_idx[k]]
        if k in local:
            states[s] = perturb_state(tru, r)
            states[s]["dbg"] = 0
            sflags[s] = 2
        else:
            states[s] = perturb_state(tru, r, dp=0.002, drot=np.deg2rad(0.05), dv=0.005, dbg=1e-4, dba=1e-3)
            sflags[s] = 1 | (2 | 4 if k in prev else 0)
    pos_of = {k: s for s, k in enumerate(order)}
    imu_i, imu_j, pre, dtk = [], [], [], []
    for k in local:
        if k - 1 in pos_of and k - 1 >= 0 and (k - 1 in local or k - 1 in prev):
            imu_i.append(pos_of[k - 1]); imu_j.append(pos_of[k]); pre.append(preints_kf[k])
            dtk.append(seq["times"][kf_idx[k]] - seq["times"][kf_idx[k - 1]])
    # points: seen from a random local KF, then observed by a run of neighbouring KFs
    es, ep, eo, ew, ef = [], [], [], [], []
    X_all = []
    p = 0
    while p < n_points:
        k0 = int(r.choice(local))
        X = landmarks_in_view(cam, seq["truth"][kf_idx[k0]], 1, r)
        nobs = int(r.integers(obs_per_point[0], obs_per_point[1] + 1))
        cands = [k for k in range(k0 - nobs, k0 + nobs + 1) if k in pos_of]
        seen = []
        for k in cands:
            u, v, ur, z = project(cam, seq["truth"][kf_idx[k]], X)
            if z[0] > 0.3 and 0 < u[0] < EUROC["w"] and 0 < v[0] < EUROC["h"]:
                seen.append(k)
        if len(seen) < 2:
            continue
        seen = sorted(seen, key=lambda k: abs(k - k0))[:nobs]
        for k in sorted(seen, key=lambda k: pos_of[k]):
            obs, w, fl, _ = make_observations(cam, seq["truth"][kf_idx[k]], X, r, outlier_frac)
            es.append(pos_of[k]); ep.append(p); eo.append(obs[0]); ew.append(w[0]); ef.append(fl[0])
        X_all.append(X[0] + r.normal(0, 0.02, 3))
        p += 1
    return dict(states=states, state_flags=sflags, points=np.asarray(X_all, np.float32).astype(np.float64),
                edge_state=np.asarray(es, np.int32), edge_point=np.asarray(ep, np.int32),
                obs=np.asarray(eo, np.float32), inv_sigma2=np.asarray(ew, np.float32),
                edge_flags=np.asarray(ef, np.uint8), imu_i=np.asarray(imu_i, np.int32), imu_j=np.asarray(imu_j, np.int32),
                preint=np.asarray(pre), imu_dt_kf=np.asarray(dtk, np.float64), gw=GRAVITY_W.copy(),
                inv_sigma_bg2=1.0 / EUROC_IMU_SIGMA[2] ** 2, inv_sigma_ba2=1.0 / EUROC_IMU_SIGMA[3] ** 2)


def make_gba_problem(seq, preints_kf, kf_idx, cam, n_points=3000, obs_window=5, seed=0, loop_frac=0.1, outlier_frac=0.0):
    """A GlobalBundleAdjustmentNavStatePRV problem (src/Optimizer.cc:771-1342) over every keyframe kf_idx of a sequence:
    keyframe 0 fixed (PR, V, Bias), all others free with PR / V / Bias vertices; IMU + bias-walk factors between
    consecutive keyframes; every point is observed by the keyframes within `obs_window` of the one that created it, and
    a `loop_frac` share of the points additionally by far-away keyframes that happen to see them (loop closures: they
    fill the reduced camera system off its band).  Vectorised: BASELINE configs[4] sizes (400 keyframes, 25k points,
    250k observations) build in seconds.  Same dict layout as make_lba_problem."""
    r = np.random.default_rng(seed + 4000)
    kf_idx = np.asarray(kf_idx)
    nk = len(kf_idx)
    truth = [seq["truth"][i] for i in kf_idx]
    states = np.zeros(nk, NAVSTATE_DTYPE)
    sflags = np.full(nk, 2, np.uint8)
    for k in range(nk):
        states[k] = perturb_state(truth[k], r) if k else truth[k]
        states[k]["dbg"] = 0
    sflags[0] = 1 | 2 | 4
    imu_i = np.arange(0, nk - 1, dtype=np.int32); imu_j = imu_i + 1
    pre = np.asarray([preints_kf[k] for k in range(1, nk)])
    dtk = np.asarray([seq["times"][kf_idx[k]] - seq["times"][kf_idx[k - 1]] for k in range(1, nk)], np.float64)
    k0 = r.integers(0, nk, n_points)
    X = np.zeros((n_points, 3))
    for k in range(nk):
        m = np.nonzero(k0 == k)[0]
        if len(m):
            X[m] = landmarks_in_view(cam, truth[k], len(m), r)
    loop = r.random(n_points) < loop_frac
    es, ep, eo, ew, ef = [], [], [], [], []
    for k in range(nk):
        u, v, ur, z = project(cam, truth[k], X)
        vis = (z > 0.3) & (u > 0) & (u < EUROC["w"]) & (v > 0) & (v < EUROC["h"])
        near = np.abs(k0 - k) <= obs_window
        far = loop & (np.abs(k0 - k) > 4 * obs_window) & ((k0 + k) % 7 == 0)
        m = np.nonzero(vis & (near | far))[0]
        if not len(m):
            continue
        obs, w, fl, _ = make_observations(cam, truth[k], X[m], r, outlier_frac)
        es.append(np.full(len(m), k, np.int32)); ep.append(m.astype(np.int32)); eo.append(obs); ew.append(w); ef.append(fl)
    es, ep, eo, ew, ef = (np.concatenate(a) for a in (es, ep, eo, ew, ef))
    cnt = np.bincount(ep, minlength=n_points)
    keep_pt = cnt >= 2
    remap = np.cumsum(keep_pt) - 1
    ke = keep_pt[ep]
    es, ep, eo, ew, ef = es[ke], remap[ep[ke]].astype(np.int32), eo[ke], ew[ke], ef[ke]
    order = np.argsort(ep, kind="stable")   # by point, ascending keyframe inside a point
    Xp = (X[keep_pt] + r.normal(0, 0.02, (int(keep_pt.sum()), 3))).astype(np.float32).astype(np.float64)
    return dict(states=states, state_flags=sflags, points=Xp, edge_state=es[order], edge_point=ep[order],
                obs=np.ascontiguousarray(eo[order]), inv_sigma2=ew[order], edge_flags=ef[order], imu_i=imu_i, imu_j=imu_j,
                preint=pre, imu_dt_kf=dtk, gw=GRAVITY_W.copy(),
                inv_sigma_bg2=1.0 / EUROC_IMU_SIGMA[2] ** 2, inv_sigma_ba2=1.0 / EUROC_IMU_SIGMA[3] ** 2)


# ------------------------------------------------------------------------------------------------------------
# Guided-search problems: current frames with ~1200 keypoints on the 64 x 48 frame grid and the
# map points of the last frame / the local map that project onto them.  Descriptors are random 256-bit strings; a true
# match is its keypoint's descriptor with a few flipped bits, so Hamming arg-mins, ties and the one-keypoint-one-point
# rule are all exercised.
from .layouts import KP_DTYPE, SBP_FRAME_DTYPE, SBP_LAST_FRAME, SBP_LOCAL_MAP  # noqa: E402


def _flip_bits(d, nflip, r):
    out = d.copy()
    for i in range(len(out)):
        bits = r.choice(256, int(nflip[i]), replace=False)
        for b in bits:
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def make_sbp_problem(seed, n_frames=2, mode=SBP_LAST_FRAME, n_kp=1200, n_q=700, th=15.0, motion="still", cluster=False,
                     th_far=0.0, mono=False, blocked_frac=0.0):
    """Batch of guided-search problems.  Returns dict(frames[SBP_FRAME_DTYPE], kps, uright, desc, q_* arrays, kp_blocked)."""
    r = np.random.default_rng(seed)
    cam = euroc_camera()
    fx, fy, cx, cy, bf = (np.float32(EUROC[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    W, H = EUROC["w"], EUROC["h"]
    _, scl = inv_level_sigma2()
    quota = np.array([261, 217, 181, 151, 126, 105, 87, 72], np.float64)
    seq = vio_sequence(seed + 1, n_frames + 1, noisy_imu=False)
    frames = np.zeros(n_frames, SBP_FRAME_DTYPE)
    kps, urs, descs, blocked = [], [], [], []
    qs = dict(Xw=[], level=[], angle=[], proj=[], viewcos=[], depth=[], desc=[], flags=[])
    kb = qb = 0
    for f in range(n_frames):
        ns = seq["truth"][f + 1]
        Rwb = R_from_quat(ns["q"])
        Rcw = cam["Rcb"] @ Rwb.T
        tcw = -Rcw @ ns["p"] + cam["tcb"]
        nsl = seq["truth"][f]
        Rlw = cam["Rcb"] @ R_from_quat(nsl["q"]).T
        tlw = -Rlw @ nsl["p"] + cam["tcb"]
        if motion == "forward":    # Tlrcr translation z = +0.5 m > baseline
            tlw = Rlw @ Rcw.T @ tcw + np.array([0, 0, 0.5])
        elif motion == "backward":
            tlw = Rlw @ Rcw.T @ tcw - np.array([0, 0, 0.5])
        else:
            tlw = Rlw @ Rcw.T @ tcw   # same camera centre: neither forward nor backward
        n = n_kp + int(r.integers(-40, 40))
        kp = np.zeros(n, KP_DTYPE)
        if cluster:
            kp["x"] = r.uniform(300, 380, n).astype(np.float32); kp["y"] = r.uniform(200, 260, n).astype(np.float32)
        else:
            kp["x"] = r.uniform(16, W - 16, n).astype(np.float32); kp["y"] = r.uniform(16, H - 16, n).astype(np.float32)
        kp["octave"] = r.choice(8, n, p=quota / quota.sum())
        kp["angle"] = r.uniform(0, 360, n).astype(np.float32)
        z = r.uniform(0.8, 12.0, n)
        ur = (kp["x"] - bf / z.astype(np.float32)).astype(np.float32)
        ur[r.random(n) > 0.7] = -1.0
        d = r.integers(0, 256, (n, 32), dtype=np.uint8)
        # queries: true matches, duplicates of true matches, strays
        m_true = min(int(0.8 * n_q), n)
        src = np.concatenate([r.choice(n, m_true, replace=False), r.choice(n, n_q - m_true, replace=True)])
        r.shuffle(src)
        stray = r.random(n_q) < 0.12
        u = kp["x"][src].astype(np.float64) + r.normal(0, 1.5, n_q)
        v = kp["y"][src].astype(np.float64) + r.normal(0, 1.5, n_q)
        u[stray] = r.uniform(-40, W + 40, int(stray.sum())); v[stray] = r.uniform(-40, H + 40, int(stray.sum()))
        zq = z[src] * r.uniform(0.97, 1.03, n_q)
        zq[r.random(n_q) < 0.02] *= -1.0    # behind the camera
        Pc = np.stack([(u - cx) / fx * zq, (v - cy) / fy * zq, zq], 1)
        Xw = ((Pc - tcw) @ Rcw).astype(np.float32).astype(np.float64)   # MapPoint positions are float
        qd = _flip_bits(d[src], np.where(stray, 128, r.integers(0, 70, n_q)), r)
        lvl = np.clip(kp["octave"][src] + r.integers(-1, 2, n_q), 0, 7).astype(np.int32)
        ang = (kp["angle"][src] + r.normal(0, 4, n_q) + np.where(r.random(n_q) < 0.15, r.uniform(0, 360, n_q), 0)) % 360
        qs["Xw"].append(Xw); qs["level"].append(lvl); qs["angle"].append(ang.astype(np.float32))
        uq = (u + r.normal(0, 0.5, n_q)).astype(np.float32); vq = (v + r.normal(0, 0.5, n_q)).astype(np.float32)
        qs["proj"].append(np.stack([uq, vq, (uq - bf / np.abs(zq).astype(np.float32)).astype(np.float32)], 1))
        qs["viewcos"].append(r.uniform(0.9, 1.0, n_q).astype(np.float32))
        qs["depth"].append(np.abs(zq).astype(np.float32))
        qs["desc"].append(qd); qs["flags"].append((r.random(n_q) < 0.9).astype(np.uint8))
        kps.append(kp); urs.append(ur); descs.append(d)
        blocked.append((r.random(n) < blocked_frac).astype(np.uint8))
        F = frames[f]
        F["kp_begin"], F["n_kp"], F["q_begin"], F["n_q"] = kb, n, qb, n_q
        F["minx"], F["maxx"], F["miny"], F["maxy"] = 0.0, W, 0.0, H
        F["grid_winv"] = np.float32(64) / (np.float32(W) - np.float32(0)); F["grid_hinv"] = np.float32(48) / (np.float32(H) - np.float32(0))
        F["bf"], F["b"] = bf, bf / fx
        F["fx"], F["fy"], F["cx"], F["cy"] = fx, fy, cx, cy
        F["th"], F["th_far"], F["nn_ratio"] = th, th_far, 0.8
        F["mono"], F["check_orientation"], F["n_levels"] = int(mono), 1, 8
        F["scale"][:8] = scl
        F["qcw"], F["tcw"] = quat_from_R(Rcw), tcw
        F["qlw"], F["tlw"] = quat_from_R(Rlw), tlw
        kb += n; qb += n_q
    out = dict(frames=frames, kps=np.concatenate(kps), uright=np.concatenate(urs), desc=np.concatenate(descs),
               kp_blocked=np.concatenate(blocked), mode=mode)
    for k, v_ in qs.items():
        out["q_" + k] = np.ascontiguousarray(np.concatenate(v_))
    return out


def make_reloc_problem(seed, n_frames=3, n_kp=1200, n_q=600, th=10.0, orb_dist=100, th_far=0.0, blocked_frac=0.1, cluster=False):
    """Tracking::Relocalization's guided search (ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist,
    th_far_pts), src/ORBmatcher.cc:1471-1606): the frames / keypoints / queries of make_sbp_problem plus, per query, the map
    point's scale-invariance range chosen so that MapPoint::PredictScale lands near the source keypoint's octave (some
    queries fall outside the 0.8 / 1.2 gate), and the per-frame VieoSbpReloc record."""
    from .layouts import SBP_RELOC_DTYPE
    pb = make_sbp_problem(seed, n_frames=n_frames, mode=SBP_LAST_FRAME, n_kp=n_kp, n_q=n_q, th=th, th_far=th_far,
                          blocked_frac=blocked_frac, cluster=cluster)
    r = np.random.default_rng(seed + 4242)
    _, scl = inv_level_sigma2()
    fr = pb["frames"]
    mx = np.zeros(len(pb["q_level"]), np.float32); mn = np.zeros(len(pb["q_level"]), np.float32)
    for f in range(len(fr)):
        Rcw = R_from_quat(fr[f]["qcw"]); tcw = fr[f]["tcw"]
        Ow = -Rcw.T @ tcw
        q0, nq = int(fr[f]["q_begin"]), int(fr[f]["n_q"])
        d = np.linalg.norm(pb["q_Xw"][q0:q0 + nq] - Ow, axis=1)
        lvl = pb["q_level"][q0:q0 + nq]
        mxd = d * scl[lvl] * r.uniform(0.93, 1.07, nq)          # PredictScale = ceil(log(max / d) / log 1.2) ~ lvl
        mnd = mxd / scl[7]
        far = r.random(nq) < 0.06
        mxd[far] *= 0.5                                          # beyond 1.2 x max: dropped
        near = r.random(nq) < 0.04
        mnd[near] = d[near] * 1.5                                # closer than 0.8 x min: dropped
        mx[q0:q0 + nq] = mxd; mn[q0:q0 + nq] = mnd
    rl = np.zeros(len(fr), SBP_RELOC_DTYPE)
    rl["orb_dist"] = orb_dist
    rl["log_scale_factor"] = np.float32(np.log(np.float32(1.2)))
    pb["q_max_dist"], pb["q_min_dist"], pb["reloc"] = mx, mn, rl
    pb["mode"] = 2
    return pb


# ---------------------------------------------------------------- Frame::isInFrustum / SearchLocalPoints problems
from .layouts import FRUSTUM_FRAME_DTYPE  # noqa: E402


def make_frustum_problem(seed, n_frames=2, n_kp=1200, n_q=1500, th=1.0, th_far=0.0, skip_frac=0.1, blocked_frac=0.0,
                         cluster=False):
    """A make_sbp_problem(LOCAL_MAP) batch plus what Frame::isInFrustum reads: per frame the float pose / intrinsics
    record (`frustum`, FRUSTUM_FRAME_DTYPE), per candidate map point p_wP, p_normal, p_max_dist (mfMaxDistance),
    p_min_dist (mfMinDistance), p_skip.  Points fall behind the camera, outside the image, outside the scale-invariance
    range and beyond the 60-degree viewing cone in realistic proportions."""
    pb = make_sbp_problem(seed, n_frames, mode=SBP_LOCAL_MAP, n_kp=n_kp, n_q=n_q, th=th, th_far=th_far,
                          blocked_frac=blocked_frac, cluster=cluster)
    r = np.random.default_rng(seed + 1000)
    fr = pb["frames"]
    ff = np.zeros(n_frames, FRUSTUM_FRAME_DTYPE)
    nq = len(pb["q_level"])
    wP = pb["q_Xw"].astype(np.float32)
    Pn = np.zeros((nq, 3), np.float32); mx = np.zeros(nq, np.float32); mn = np.zeros(nq, np.float32)
    sf = np.float32(1.2)
    scale = sf ** np.arange(8, dtype=np.float32)
    for f in range(n_frames):
        F = fr[f]
        b, m = int(F["q_begin"]), int(F["n_q"])
        Rcw = R_from_quat(F["qcw"]); tcw = np.array(F["tcw"])
        Ow = -Rcw.T @ tcw
        G = ff[f]
        G["q_begin"], G["n_q"] = b, m
        G["Rcw"] = Rcw.astype(np.float32).ravel(); G["tcw"] = tcw.astype(np.float32); G["Ow"] = Ow.astype(np.float32)
        for k in ("fx", "fy", "cx", "cy", "minx", "maxx", "miny", "maxy", "bf"):
            G[k] = F[k]
        G["cos_limit"] = 0.5
        G["log_scale_factor"] = np.log(sf)     # float log of the float scale factor (src/FrameBase.cpp:288)
        G["n_levels"] = 8
        PO = wP[b:b + m].astype(np.float64) - Ow
        d = np.linalg.norm(PO, axis=1)
        # mean viewing direction: the ray perturbed by up to ~75 degrees, so a share of the points leaves the cone
        ang = np.abs(r.normal(0, 0.6, m))
        ax = np.cross(PO, r.normal(0, 1, (m, 3))); ax /= np.linalg.norm(ax, axis=1, keepdims=True) + 1e-300
        dirn = PO / (d[:, None] + 1e-300)
        nrm = dirn * np.cos(ang)[:, None] + ax * np.sin(ang)[:, None]
        Pn[b:b + m] = (nrm * r.uniform(0.9, 1.0, (m, 1))).astype(np.float32)   # mNormalVector is a mean of unit vectors
        lvl = r.integers(0, 8, m)
        dref = d * r.uniform(0.5, 1.9, m)                                      # reference distance of the observation
        mx[b:b + m] = (dref * scale[lvl]).astype(np.float32)
        mn[b:b + m] = mx[b:b + m] / scale[7]
    pb["frustum"] = ff
    pb["p_wP"], pb["p_normal"], pb["p_max_dist"], pb["p_min_dist"] = wP, Pn, mx, mn
    pb["p_skip"] = (r.random(nq) < skip_frac).astype(np.uint8)
    return pb


def make_distinctive_problem(seed, n_points=500, max_obs=30, long_lists=()):
    """CSR batch for MapPoint::ComputeDistinctiveDescriptors: every map point owns n observations (0 .. max_obs, plus the
    explicit lengths of `long_lists`) whose descriptors are noisy copies of a per-point prototype, addressed through a
    shuffled row table into one descriptor pool.  -> dict(pool, rows, ptr)"""
    r = np.random.default_rng(seed)
    n_obs = np.concatenate([r.integers(0, max_obs + 1, n_points), np.asarray(long_lists, np.int64)]).astype(np.int64)
    r.shuffle(n_obs)
    ptr = np.zeros(len(n_obs) + 1, np.int32)
    ptr[1:] = np.cumsum(n_obs)
    total = int(ptr[-1])
    proto = r.integers(0, 256, (len(n_obs), 32), dtype=np.uint8)
    owner = np.repeat(np.arange(len(n_obs)), n_obs)
    d = _flip_bits(proto[owner], r.integers(0, 60, total), r)
    # coarse distances create many equal medians: the first-row-wins rule is exercised
    perm = r.permutation(total + 17)[:total].astype(np.int32)
    pool = r.integers(0, 256, (total + 17, 32), dtype=np.uint8)
    pool[perm] = d
    return dict(pool=pool, rows=perm, ptr=ptr)


def make_gyro_bias_problem(seed, n_kf=20, kf_gap=4, bg_true=(0.02, -0.015, 0.01), noisy_imu=True):
    """Keyframe chain for Optimizer::OptimizeInitialGyroBias: IMU samples carry the gyro bias bg_true, the keyframe
    pre-integrations are linearised at zero bias.  -> dict(seq, kf_idx, Rwb (n_kf,3,3), samples, seg_ptr, ti_tj)"""
    # kf_gap = (lo, hi): keyframe spacing drawn per pair, so the rotation covariances (the bInfo weights) differ
    if np.ndim(kf_gap) == 0:
        idx = np.arange(n_kf) * int(kf_gap)
    else:
        gaps = np.random.default_rng(seed + 5).integers(kf_gap[0], kf_gap[1] + 1, n_kf - 1)
        idx = np.r_[0, np.cumsum(gaps)]
    seq = vio_sequence(seed, int(idx[-1]) + 1, noisy_imu=noisy_imu)
    seq["imu"][:, 4:7] += np.asarray(bg_true) - seq["bg"]
    seq["truth"]["bg"] = 0.0
    seq["truth"]["ba"] = seq["ba"]
    Rwb = np.stack([R_from_quat(seq["truth"][i]["q"]) for i in idx])
    imu, t = seq["imu"], seq["times"]
    seg, tt, chunks = [0, 0], [(0.0, 0.0)], []
    for k in range(1, n_kf):
        ti, tj = t[idx[k - 1]], t[idx[k]]
        lo = max(np.searchsorted(imu[:, 0], ti, "right") - 1, 0)
        hi = min(np.searchsorted(imu[:, 0], tj, "left") + 1, len(imu))
        chunks.append(imu[lo:hi]); seg.append(seg[-1] + hi - lo); tt.append((ti, tj))
    return dict(seq=seq, kf_idx=idx, Rwb=Rwb, samples=np.vstack(chunks), seg_ptr=np.array(seg, np.int32),
                ti_tj=np.array(tt), bg_true=np.asarray(bg_true, np.float64))


from .layouts import PROJ_SEARCH_FRAME_DTYPE  # noqa: E402


def make_fuse_problem(seed, n_frames=3, n_kp=1200, n_q=1500, th_radius=3.0, use_bf=True, check_viewing_angle=True,
                      skip_frac=0.1, cluster=False):
    """Batch for ORBmatcher::SearchByProjectionBase / Fuse: keyframes with keypoints + map points to project into them
    (make_frustum_problem geometry), predicted levels consistent with the keypoint octaves for the true matches so that the
    level band and the chi-square gate both pass and reject in realistic proportions."""
    pb = make_frustum_problem(seed, n_frames=n_frames, n_kp=n_kp, n_q=n_q, th=1.0, skip_frac=skip_frac, cluster=cluster)
    r = np.random.default_rng(seed + 2000)
    fr = np.zeros(n_frames, PROJ_SEARCH_FRAME_DTYPE)
    inv_s2, scl = inv_level_sigma2()
    for f in range(n_frames):
        F, G, H = pb["frames"][f], pb["frustum"][f], fr[f]
        for k in ("kp_begin", "n_kp", "q_begin", "n_q", "minx", "maxx", "miny", "maxy", "grid_winv", "grid_hinv", "bf", "fx",
                  "fy", "cx", "cy"):
            H[k] = F[k]
        for k in ("Rcw", "tcw", "Ow", "log_scale_factor", "n_levels"):
            H[k] = G[k]
        H["use_bf"], H["check_viewing_angle"], H["th_radius"] = int(use_bf), int(check_viewing_angle), th_radius
        H["scale"][:8] = scl; H["inv_level_sigma2"][:8] = inv_s2
        # scale-invariance range such that PredictScale lands on (or next to) the query's level: mfMaxDistance =
        # dist * 1.2^level * jitter
        b, m = int(F["q_begin"]), int(F["n_q"])
        d = np.linalg.norm(pb["p_wP"][b:b + m].astype(np.float64) - np.array(G["Ow"], np.float64), axis=1)
        lv = pb["q_level"][b:b + m]
        mx = d * 1.2 ** (lv - 0.5) * r.uniform(0.9, 1.1, m)
        far = r.random(m) < 0.08                       # outside the invariance range
        mx[far] *= r.choice([0.2, 6.0], int(far.sum()))
        pb["p_max_dist"][b:b + m] = mx.astype(np.float32)
        pb["p_min_dist"][b:b + m] = pb["p_max_dist"][b:b + m] / np.float32(scl[7])
    pb["frames_sbp"] = pb["frames"]
    pb["frames"] = fr
    return pb


# ------------------------------------------------------------------------------------------------------------
# SearchForTriangulation problems: keyframe pairs related by a known relative pose, keypoints of
# KF2 = projections of KF1's back-projected keypoints (so the epipolar constraint holds for true matches) + distractors,
# descriptors = near copies, DBoW2-like FeatureVectors (keypoints hashed into vocabulary nodes, true matches mostly in
# the same node), existing map points, stereo / mono mix, rotation outliers.
from .layouts import SFT_PAIR_DTYPE  # noqa: E402


def _fundamental(K1, K2, R12, t12):
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    return np.linalg.inv(K1.T) @ tx @ R12 @ np.linalg.inv(K2)


def make_sft_problem(seed, n_pairs=4, n_kp=1200, n_nodes=90, share_kf1=True):
    """-> dict of flat arrays in the C-ABI layout of vieo_search_for_triangulation_dev (+ 'pairs' SFT_PAIR_DTYPE)."""
    r = np.random.default_rng(seed)
    fx, fy, cx, cy = (np.float32(EUROC[k]) for k in ("fx", "fy", "cx", "cy"))
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    scale = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    sigma2 = (scale * scale).astype(np.float32)
    kps, ur, desc, has_mp, fv_node, fv_ptr, fv_idx = [], [], [], [], [], [], []
    kf = []  # per keyframe: (kp_begin, n_kp, node_begin, n_nodes, ptr_begin, idx_begin)

    def add_kf(k, u, d, m, node_of_kp):
        kb = sum(len(x) for x in kps)
        ids = np.unique(node_of_kp)
        lists = [np.nonzero(node_of_kp == i)[0] for i in ids]
        for L in lists:
            r.shuffle(L)  # push order inside a node is the keypoint loop order in DBoW2; any fixed order serves
        ptr = np.concatenate([[0], np.cumsum([len(L) for L in lists])]).astype(np.int32)
        rec = (kb, len(k), sum(len(x) for x in fv_node), len(ids), sum(len(x) for x in fv_ptr), sum(len(x) for x in fv_idx))
        kps.append(k); ur.append(u); desc.append(d); has_mp.append(m)
        fv_node.append(ids.astype(np.int32)); fv_ptr.append(ptr); fv_idx.append(np.concatenate(lists).astype(np.int32))
        kf.append(rec)
        return len(kf) - 1

    def rand_kf(n):
        k = np.zeros(n, KP_DTYPE)
        k["x"] = r.uniform(20, EUROC["w"] - 20, n).astype(np.float32)
        k["y"] = r.uniform(20, EUROC["h"] - 20, n).astype(np.float32)
        k["octave"] = r.integers(0, 8, n)
        k["angle"] = r.uniform(0, 360, n).astype(np.float32)
        k["size"] = 31
        return k

    pairs = np.zeros(n_pairs, SFT_PAIR_DTYPE)
    rel_R12, rel_t12 = [], []  # the relative poses the geometry was made with (for callers that start from poses, not from F12)
    k1 = rand_kf(n_kp)
    d1 = r.integers(0, 256, (n_kp, 32), dtype=np.uint8)
    node1 = r.integers(0, n_nodes, n_kp) * 7 + 3  # sparse node ids
    u1 = np.where(r.random(n_kp) < 0.5, k1["x"] - r.uniform(2, 40, n_kp), -1).astype(np.float32)
    m1 = (r.random(n_kp) < 0.35).astype(np.uint8)
    i_kf1 = add_kf(k1, u1, d1, m1, node1)
    out_b = nscr_b = 0
    for p in range(n_pairs):
        if not share_kf1 and p > 0:
            k1 = rand_kf(n_kp); d1 = r.integers(0, 256, (n_kp, 32), dtype=np.uint8); node1 = r.integers(0, n_nodes, n_kp) * 7 + 3
            u1 = np.where(r.random(n_kp) < 0.5, k1["x"] - r.uniform(2, 40, n_kp), -1).astype(np.float32)
            m1 = (r.random(n_kp) < 0.35).astype(np.uint8)
            i_kf1 = add_kf(k1, u1, d1, m1, node1)
        # relative pose 1 <- 2 and the second keyframe
        a = r.normal(0, 0.05, 3); th = np.linalg.norm(a); ax = a / th
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R12 = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t12 = r.normal(0, 0.25, 3)
        n2 = n_kp + int(r.integers(-100, 100))
        k2 = rand_kf(n2)
        d2 = r.integers(0, 256, (n2, 32), dtype=np.uint8)
        node2 = r.integers(0, n_nodes + 10, n2) * 7 + 3
        # true correspondences for ~60 % of KF1's keypoints: X1 = depth * K^-1 x1; x2 = K R21 (X1 - t12)
        nm = int(0.6 * min(n_kp, n2))
        src = r.permutation(n_kp)[:nm]; dst = r.permutation(n2)[:nm]
        z = r.uniform(2, 15, nm)
        X1 = (np.linalg.inv(K) @ np.stack([k1["x"][src], k1["y"][src], np.ones(nm)])) * z
        X2 = R12.T @ (X1 - t12[:, None])
        x2 = K @ (X2 / X2[2])
        ok = (X2[2] > 0.5) & (x2[0] > 5) & (x2[0] < EUROC["w"] - 5) & (x2[1] > 5) & (x2[1] < EUROC["h"] - 5)
        src, dst, x2 = src[ok], dst[ok], x2[:, ok]
        noise = r.normal(0, 0.6, (2, len(src))) * scale[k1["octave"][src]]
        k2["x"][dst] = (x2[0] + noise[0]).astype(np.float32); k2["y"][dst] = (x2[1] + noise[1]).astype(np.float32)
        k2["octave"][dst] = np.clip(k1["octave"][src] + r.integers(-1, 2, len(src)), 0, 7)
        rot_common = r.uniform(-20, 20)
        k2["angle"][dst] = ((k1["angle"][src] - rot_common + r.normal(0, 3, len(src))) % 360).astype(np.float32)
        bad_rot = r.random(len(src)) < 0.08
        k2["angle"][dst[bad_rot]] = r.uniform(0, 360, int(bad_rot.sum())).astype(np.float32)
        d2[dst] = d1[src]
        nflip = r.integers(0, 60, len(src))  # some beyond TH_LOW
        for j, f in zip(dst, nflip):
            bits = r.choice(256, int(f), replace=False)
            np.bitwise_xor.at(d2[j], bits // 8, (1 << (bits % 8)).astype(np.uint8))
        node2[dst] = node1[src]
        stray = r.random(len(src)) < 0.1
        node2[dst[stray]] = r.integers(0, n_nodes, int(stray.sum())) * 7 + 3
        # decoys: duplicates of some matched descriptors in the same node (ties / competing claims)
        nd = len(src) // 10
        dec = r.permutation(n2)[:nd]; which = r.integers(0, len(src), nd)
        d2[dec] = d2[dst[which]]; node2[dec] = node2[dst[which]]
        k2["x"][dec] = k2["x"][dst[which]] + r.normal(0, 1.0, nd).astype(np.float32)
        k2["y"][dec] = k2["y"][dst[which]] + r.normal(0, 1.0, nd).astype(np.float32)
        k2["octave"][dec] = k2["octave"][dst[which]]
        u2 = np.where(r.random(n2) < 0.5, k2["x"] - r.uniform(2, 40, n2), -1).astype(np.float32)
        m2 = (r.random(n2) < 0.35).astype(np.uint8)
        i_kf2 = add_kf(k2, u2, d2, m2, node2)
        A, B = kf[i_kf1], kf[i_kf2]
        P = pairs[p]
        P["kp1_begin"], P["n_kp1"], P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"] = A
        P["kp2_begin"], P["n_kp2"], P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"] = B
        P["out_begin"], P["nscr_begin"] = out_b, nscr_b
        out_b += A[1]; nscr_b += A[3]
        P["only_stereo"] = int(p % 4 == 3)
        P["check_orientation"] = int(p % 3 != 2)
        C2 = R12.T @ (-t12)  # centre of camera 1 in camera 2
        e = K.astype(np.float32) @ np.array([C2[0] / C2[2], C2[1] / C2[2], 1], np.float32)
        P["ex"], P["ey"] = e[0], e[1]
        P["scale_factor2"][:8] = scale; P["level_sigma2_2"][:8] = sigma2
        P["F12"] = _fundamental(K, K, R12, t12).reshape(-1)
        rel_R12.append(R12); rel_t12.append(t12)
    return dict(rel_R12=np.array(rel_R12), rel_t12=np.array(rel_t12), pairs=pairs, kps=np.concatenate(kps), uright=np.concatenate(ur), desc=np.concatenate(desc),
                has_mp=np.concatenate(has_mp), fv_node=np.concatenate(fv_node), fv_ptr=np.concatenate(fv_ptr),
                fv_idx=np.concatenate(fv_idx), n_out_total=out_b, n_nodes1_total=nscr_b)


def make_bow_problem(seed, n_pairs=4, n_kp=1200, n_nodes=90, nn_ratio=0.7):
    """SearchByBoW(KeyFrame, Frame) problems on the keyframe pairs of make_sft_problem: side 1 = keyframe (its keypoints WITH
    a map point are the queries), side 2 = frame.  -> the same flat arrays + 'pairs' (BOW_PAIR_DTYPE), 'mp_ok'."""
    from .layouts import BOW_PAIR_DTYPE
    pb = make_sft_problem(seed, n_pairs=n_pairs, n_kp=n_kp, n_nodes=n_nodes, share_kf1=False)
    bp = np.zeros(n_pairs, BOW_PAIR_DTYPE)
    ob = 0
    for p in range(n_pairs):
        for k in ("kp1_begin", "n_kp1", "kp2_begin", "n_kp2", "node1_begin", "n_nodes1", "node2_begin", "n_nodes2", "ptr1_begin",
                  "ptr2_begin", "idx1_begin", "idx2_begin"):
            bp[k][p] = pb["pairs"][k][p]
        bp["out_begin"][p] = ob
        ob += int(bp["n_kp2"][p])
        bp["check_orientation"][p] = int(p % 2 == 0)
        bp["nn_ratio"][p] = nn_ratio if p % 3 else 0.9
    r = np.random.default_rng(seed + 99)
    out = dict(pb)
    out["pairs"] = bp
    out["mp_ok"] = (r.random(len(pb["has_mp"])) < 0.6).astype(np.uint8)
    out["n_out_total"] = ob
    return out


def make_sim3_problems(cam, n_candidates=4, n_matches=120, seed=0, fix_scale=False, outlier_frac=0.15, th2=10.0,
                       few_matches_every=0):
    """Loop-closing candidates for Optimizer::OptimizeSim3 (src/Optimizer.cc:2689-2920): two keyframes looking at the same
    landmarks whose camera frames are related by a similarity S12 (x1 = s R12 x2 + t12, the drift the loop closer corrects).
    Per match: the landmark in both camera frames (float products like :2771-2783), its keypoints in both images with
    octave-dependent noise, some outliers.  The initial estimate is the true S12 perturbed.  few_matches_every = k: every
    k-th candidate has 12 matches most of which are outliers (the "fewer than 10 inliers" exit).
    -> (problems SIM3_PROBLEM_DTYPE, Xc1, Xc2, obs1, obs2, w1, w2, truth list of (R12, t12, s))"""
    from .layouts import SIM3_PROBLEM_DTYPE
    r = np.random.default_rng(seed + 7000)
    inv_s2, scl = inv_level_sigma2()
    pbs = np.zeros(n_candidates, SIM3_PROBLEM_DTYPE)
    X1s, X2s, O1, O2, W1, W2, truth = [], [], [], [], [], [], []
    m0 = 0
    fx, fy, cx, cy = (float(cam[k]) for k in ("fx", "fy", "cx", "cy"))
    for c in range(n_candidates):
        few = few_matches_every and (c % few_matches_every == few_matches_every - 1)
        M = 12 if few else n_matches
        s = 1.0 if fix_scale else float(r.uniform(0.85, 1.2))
        R12 = so3_exp(r.normal(0, 0.15, 3))
        t12 = r.normal(0, 0.4, 3)
        # landmarks in camera-2 coordinates, in front of both cameras
        u2 = r.uniform(40, EUROC["w"] - 40, M); v2 = r.uniform(40, EUROC["h"] - 40, M); z2 = r.uniform(2.0, 9.0, M)
        X2 = np.stack([(u2 - cx) / fx * z2, (v2 - cy) / fy * z2, z2], 1)
        X1 = s * X2 @ R12.T + t12
        X1f = X1.astype(np.float32).astype(np.float64)
        X2f = X2.astype(np.float32).astype(np.float64)
        oct1 = r.integers(0, 8, M); oct2 = r.integers(0, 8, M)
        o1 = np.stack([fx * X1[:, 0] / X1[:, 2] + cx, fy * X1[:, 1] / X1[:, 2] + cy], 1) + r.normal(0, 1, (M, 2)) * scl[oct1][:, None]
        o2 = np.stack([u2, v2], 1) + r.normal(0, 1, (M, 2)) * scl[oct2][:, None]
        out = r.random(M) < (0.7 if few else outlier_frac)
        o1[out] += r.uniform(-25, 25, (int(out.sum()), 2))
        pb = pbs[c]
        # the vertex: mRwb = R12^-1, mpwb = -(mRwb t12) (:2717-2722), started off the truth
        Rp = R12 @ so3_exp(r.normal(0, np.deg2rad(1.0), 3))
        tp = t12 + r.normal(0, 0.03, 3)
        pb["ns"]["q"] = quat_from_R(Rp.T)
        pb["ns"]["p"] = -(Rp.T @ tp)
        pb["scale"] = 1.0 if fix_scale else s * float(r.uniform(0.97, 1.03))
        pb["th2"] = th2
        pb["fix_scale"] = int(fix_scale)
        pb["m_begin"], pb["m_end"] = m0, m0 + M
        m0 += M
        X1s.append(X1f); X2s.append(X2f); O1.append(o1.astype(np.float32)); O2.append(o2.astype(np.float32))
        W1.append(inv_s2[oct1]); W2.append(inv_s2[oct2]); truth.append((R12, t12, s))
    return (pbs, np.concatenate(X1s), np.concatenate(X2s), np.concatenate(O1), np.concatenate(O2), np.concatenate(W1),
            np.concatenate(W2), truth)


# ---- Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:2309-2688) ----------------------------------------------------------
def _s3(R, t, s=1.0):
    return (np.asarray(R, float), np.asarray(t, float), float(s))


def _s3_mul(a, b):
    return (a[0] @ b[0], a[2] * (a[0] @ b[1]) + a[1], a[2] * b[2])


def _s3_inv(a):
    return (a[0].T, -(a[0].T @ a[1]) / a[2], 1.0 / a[2])


def _s3_rec(S):
    from .layouts import SIM3_DTYPE
    o = np.zeros((), SIM3_DTYPE)
    q = quat_from_R(S[0])  # (w, x, y, z)
    o["q"] = [q[1], q[2], q[3], q[0]]
    o["t"] = S[1]
    o["s"] = S[2]
    return o


def make_essential_graph(K=60, seed=0, fix_scale=True, n_neighbors=5, odom_info_every=0, n_points=0, extra_loops=2,
                         drift_rot_deg=0.15, drift_trans=0.01, drift_scale=0.004):
    """The graph Optimizer::OptimizeEssentialGraph hands to g2o, as LoopClosing::CorrectLoop produces it: K keyframes on a closed
    circuit whose map poses carry accumulated odometry drift (rotation, translation and, when !fix_scale, scale); the last keyframe
    re-observes keyframe `loop` (2) and it and its n_neighbors predecessors carry a corrected Sim3 (CorrectedSim3, propagated through the
    non-corrected relative poses, src/LoopClosing.cc CorrectLoop).  Vertices start from vScw (corrected where available, :2357-2371), the
    loop keyframe is fixed (:2373).  Edges (vertex 0 = i, vertex 1 = j, measurement Sji):
      * new loop connections (:2396-2430): corrected-set keyframe i -> keyframes j around the loop keyframe, Sji = vScw[j] * vScw[i]^-1;
      * spanning tree (:2490-2549): parent i-1 (every 7th keyframe: i-2), Sji from the NON-corrected poses; with odom_info_every = k every
        k-th of them carries the reduced odometry information matrix diag(a I3, b I3, 1) of :2507-2541;
      * earlier loop edges (:2551-2577) and covisibility >= 100 edges (:2579-2615), both with j < i, non-corrected.
    n_points > 0 adds map points (float positions, reference keyframe index) for the correction step (:2645-2676).
    -> dict(Scw, fixed, fix_scale, ei, ej, meas, info (None or [E][49]), truth (Scw of the drift-free circuit), Pw, ref)"""
    from .layouts import SIM3_DTYPE
    r = np.random.default_rng(seed + 9100)
    # drift-free circuit: camera centres on a circle of radius 6 m with some height variation, looking along the tangent
    ang = np.linspace(0, 2 * np.pi, K, endpoint=False) * (K - 1) / K * 0.98
    Twc = []
    for a in ang:
        Rwc = so3_exp(np.array([0.0, 0.0, a])) @ so3_exp(np.array([0.03 * np.sin(3 * a), 0.02 * np.cos(2 * a), 0.0]))
        pw = np.array([6 * np.cos(a), 6 * np.sin(a), 0.4 * np.sin(2 * a)])
        Twc.append((Rwc, pw))
    true_cw = [_s3(Rwc.T, -(Rwc.T @ pw)) for Rwc, pw in Twc]
    # the map: relative motions integrated with drift
    map_cw = [true_cw[0]]
    sc = 1.0
    for k in range(1, K):
        rel = _s3_mul(true_cw[k], _s3_inv(true_cw[k - 1]))  # S_k,k-1
        dR = so3_exp(r.normal(0, np.deg2rad(drift_rot_deg), 3))
        if not fix_scale:
            sc *= 1.0 + r.normal(0, drift_scale)
        rel_d = _s3(dR @ rel[0], rel[1] * sc + r.normal(0, drift_trans, 3))
        map_cw.append(_s3_mul(rel_d, map_cw[-1]))
    loop, cur = 2, K - 1
    # corrected Sim3 of the current keyframe: measured S_cur,loop (near the truth, with the scale the drift accumulated) * S_loop,w
    rel_cl = _s3_mul(true_cw[cur], _s3_inv(true_cw[loop]))
    s_cl = 1.0 if fix_scale else 1.0 / sc * (1.0 + r.normal(0, 0.002))
    meas_cl = _s3(so3_exp(r.normal(0, np.deg2rad(0.05), 3)) @ rel_cl[0], rel_cl[1] + r.normal(0, 0.005, 3), s_cl)
    corr_cur = _s3_mul(meas_cl, map_cw[loop])
    corrected = {}
    non_corrected = {}
    for i in range(cur - n_neighbors, cur + 1):
        Sic = _s3_mul(map_cw[i], _s3_inv(map_cw[cur]))
        corrected[i] = _s3_mul(Sic, corr_cur)
        non_corrected[i] = map_cw[i]
    vScw = [corrected.get(k, map_cw[k]) for k in range(K)]
    nc = lambda k: non_corrected.get(k, vScw[k])  # noqa: E731  (:2470-2476)
    ei, ej, meas, info = [], [], [], []
    I7 = np.eye(7)

    def add(i, j, Sji, om=I7):
        ei.append(i); ej.append(j); meas.append(_s3_rec(Sji)); info.append(om.reshape(-1))
    inserted = set()
    for i in sorted(corrected):  # new loop connections
        for j in range(max(loop - 2, 0), loop + 3):
            if r.random() < 0.75 or (i == cur and j == loop):
                add(i, j, _s3_mul(vScw[j], _s3_inv(vScw[i])))
                inserted.add((min(i, j), max(i, j)))
    prev_loops = set()
    for _ in range(extra_loops if K >= 24 else 0):
        a = int(r.integers(K // 2, K - n_neighbors - 2)); b = int(r.integers(4, K // 3))
        prev_loops.add((a, b))
    for i in range(K):
        Swi = _s3_inv(nc(i))
        par = None
        if i > 0:
            par = i - 2 if (i % 7 == 0 and i >= 2) else i - 1
            om = I7
            if odom_info_every and i % odom_info_every == 0:
                om = np.diag([*(3 * [float(r.uniform(0.05, 0.9))]), *(3 * [float(r.uniform(0.05, 0.9))]), 1.0])
            add(i, par, _s3_mul(nc(par), Swi), om)
        for (a, b) in sorted(prev_loops):
            if a == i:
                add(i, b, _s3_mul(nc(b), Swi))
        for d in (2, 3, 4):
            j = i - d
            if j < 0 or j == par or (min(i, j), max(i, j)) in inserted or (i, j) in prev_loops:
                continue
            if r.random() < (0.8 if d == 2 else 0.4):
                add(i, j, _s3_mul(nc(j), Swi))
    fixed = np.zeros(K, np.uint8)
    fixed[loop] = 1
    out = dict(Scw=np.array([_s3_rec(S) for S in vScw], SIM3_DTYPE), fixed=fixed, fix_scale=int(fix_scale),
               ei=np.array(ei, np.int32), ej=np.array(ej, np.int32), meas=np.array(meas, SIM3_DTYPE),
               info=np.array(info) if odom_info_every else None, truth=np.array([_s3_rec(S) for S in true_cw], SIM3_DTYPE),
               loop=loop, cur=cur)
    if n_points:
        ref = r.integers(0, K, n_points).astype(np.int32)
        Pw = np.zeros((n_points, 3), np.float32)
        for i in range(n_points):
            S = _s3_inv(vScw[ref[i]])
            Pc = np.array([r.uniform(-2, 2), r.uniform(-1.5, 1.5), r.uniform(1.5, 9.0)])
            Pw[i] = S[2] * (S[0] @ Pc) + S[1]
        out["Pw"], out["ref"] = Pw, ref
    return out


def make_frustum_rig_problem(seed, n_frames=2, n_q=1200, n_cams=4, model=2, skip_frac=0.05):
    """Frame::isInFrustum with a camera rig (src/Frame.cc:351-411, mpCameras.size() > 1): the map points, normals and distance
    ranges of make_frustum_problem, and per frame a rig of n_cams cameras looking left / right / up-left / up-right of the
    reference frame (TUM-VI-like 512 x 512 KB8 cameras when model == 2; model 0 = undistorted K multiply, 1 = PinholeCamera::Project
    in double), each with its own extrinsics Trc, intrinsics and image bounds.  -> dict(rig FRUSTUM_RIG_FRAME_DTYPE[n_frames],
    p_wP, p_normal, p_max_dist, p_min_dist, p_skip)"""
    from .layouts import FRUSTUM_RIG_FRAME_DTYPE
    pb = make_frustum_problem(seed, n_frames=n_frames, n_q=n_q, skip_frac=skip_frac)
    r = np.random.default_rng(seed + 4100)
    rig = np.zeros(n_frames, FRUSTUM_RIG_FRAME_DTYPE)
    yaw = [0.35, -0.35, 0.9, -0.9]
    for f in range(n_frames):
        G, S = rig[f], pb["frustum"][f]
        for k in ("q_begin", "n_q", "Rcw", "tcw", "Ow", "bf", "cos_limit", "log_scale_factor", "n_levels"):
            G[k] = S[k]
        G["n_cams"] = n_cams
        for c in range(n_cams):
            C = G["cam"][c]
            Rrc = so3_exp(np.array([0.05 * r.normal(), yaw[c] + 0.02 * r.normal(), 0.03 * r.normal()]))  # camera c in the reference frame
            trc = np.array([0.05 * (c - 1.5), 0.01 * r.normal(), 0.02 * r.normal()])
            Rcr = Rrc.T
            tcr = -Rcr @ trc
            q = quat_from_R(Rcr)  # (w, x, y, z)
            C["q_cr"] = np.array([q[1], q[2], q[3], q[0]], np.float32)
            C["t_cr"] = tcr.astype(np.float32)
            C["t_rc"] = trc.astype(np.float32)
            if model == 2:
                C["fx"], C["fy"], C["cx"], C["cy"] = 190.98 + c, 190.97 - c, 254.93 + 2 * c, 256.9 - c
                C["k"] = np.array([0.0034823894, 0.0007150348, -0.0020532361, 0.00020293673]) * (1 + 0.05 * c)
                C["minx"], C["maxx"], C["miny"], C["maxy"] = 0.0, 512.0, 0.0, 512.0
            else:
                C["fx"], C["fy"], C["cx"], C["cy"] = S["fx"] + c, S["fy"] - c, S["cx"] + 2 * c, S["cy"] - c
                C["minx"], C["maxx"], C["miny"], C["maxy"] = S["minx"], S["maxx"], S["miny"], S["maxy"]
            C["model"] = model
    pb["rig"] = rig
    return pb

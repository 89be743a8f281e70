"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (parity checker / CPU baseline)."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


class OrcKeyPoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32)]


KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float)
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_sincosf.argtypes = [C.c_float, f32p, f32p]
        L.orc_orb_create.restype = C.c_void_p
        L.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_orb_destroy.argtypes = [C.c_void_p]
        L.orc_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, i32p]
        L.orc_orb_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.orc_orb_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_orb_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        for name in ("orc_resize_linear_u8", "orc_fast_score_map", "orc_fast_detect", "orc_gaussian_blur7_u8",
                     "orc_quadtree", "orc_gauss7_kernel", "orc_descriptor_distance", "orc_hamming_knn2",
                     "orc_hamming_csr"):
            getattr(L, name).argtypes = None
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(src.strides[0]), _p(dst),
                               C.c_int(dw), C.c_int(dh), C.c_int(dw))
    return dst


def fast_score_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_fast_score_map(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]), _p(out),
                             C.c_int(img.shape[1]))
    return out


def fast_detect(img, th):
    """cv::FAST(img, th, nms=True) -> (n,3) int array of (x, y, response), raster order."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    cap = max(16, img.shape[0] * img.shape[1] // 4)
    xs, ys, rs = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orc_fast_detect(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]),
                              C.c_int(th), _p(xs), _p(ys), _p(rs), C.c_int(cap))
    return np.stack([xs[:n], ys[:n], rs[:n]], 1)


def gaussian_blur7(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_gaussian_blur7_u8(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(img.strides[0]), _p(out),
                                C.c_int(img.shape[1]))
    return out


def gauss7_kernel():
    k = np.empty(7, np.int32)
    lib().orc_gauss7_kernel(_p(k))
    return k


def fast_atan2(y, x):
    return lib().orc_fast_atan2(C.c_float(y), C.c_float(x))


def sincosf(x):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincosf(C.c_float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def quadtree(xyr, W, H, N):
    xyr = np.ascontiguousarray(xyr, np.int32)
    out = np.empty(max(len(xyr), 1), np.int32)
    n = lib().orc_quadtree(_p(xyr), C.c_int(len(xyr)), C.c_int(W), C.c_int(H), C.c_int(N), _p(out), C.c_int(len(out)))
    return out[:n]


class OrbOracle:
    """Mirror of ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)."""

    def __init__(self, nfeatures=1200, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.h = lib().orc_orb_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_orb_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        q, um = np.empty(n, np.int32), np.empty(16, np.int32)
        lib().orc_orb_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(q), _p(um))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, quota=q, umax=um)

    def extract(self, img, lapping=None, cap=None):
        img = np.ascontiguousarray(img, np.uint8)
        cap = cap or 8192
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        mono = C.c_int32(0)
        lap = None if lapping is None else _p(np.asarray(lapping, np.int32))
        self._keep = lapping
        n = lib().orc_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], lap, _p(kps), _p(desc),
                                  cap, C.byref(mono))
        if n < 0:
            return n, None, None, 0
        return n, kps[:n].copy(), desc[:n].copy(), mono.value

    def level(self, l):
        w, h = C.c_int32(), C.c_int32()
        lib().orc_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        lib().orc_orb_get_level(self.h, l, _p(out))
        return out

    def candidates(self, l):
        cap = 1 << 16
        out = np.empty((cap, 3), np.int32)
        n = lib().orc_orb_get_candidates(self.h, l, _p(out), cap)
        return out[:n].copy()


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def hamming_knn2(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), 2), np.int32); dist = np.empty((len(q), 2), np.int32)
    lib().orc_hamming_knn2(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), _p(idx), _p(dist))
    return idx, dist


def fisheye_matches(desc, n_kp, n_mono):
    """Frame::ComputeStereoFishEyeMatches brute-force half for one frame: desc [n_cams][cap][32] -> idx, dist
    [n_pairs][cap][2], good [n_pairs][cap]."""
    desc = np.ascontiguousarray(desc, np.uint8)
    n_cams, cap = desc.shape[0], desc.shape[1]
    n_pairs = n_cams * (n_cams - 1) // 2
    n_kp = np.ascontiguousarray(n_kp, np.int32); n_mono = np.ascontiguousarray(n_mono, np.int32)
    idx = np.empty((n_pairs, cap, 2), np.int32); dist = np.empty((n_pairs, cap, 2), np.int32)
    good = np.empty((n_pairs, cap), np.uint8)
    L = lib()
    L.orc_fisheye_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_fisheye_matches.restype = None
    L.orc_fisheye_matches(_p(desc), _p(n_kp), _p(n_mono), n_cams, cap, _p(idx), _p(dist), _p(good))
    return idx, dist, good


def hamming_csr(q, t, row_ptr, cand):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    row_ptr = np.ascontiguousarray(row_ptr, np.int32); cand = np.ascontiguousarray(cand, np.int32)
    n = len(row_ptr) - 1
    o = [np.empty(n, np.int32) for _ in range(4)]
    lib().orc_hamming_csr(_p(q), _p(t), _p(row_ptr), _p(cand), C.c_int(n), *[_p(x) for x in o])
    return o  # best_dist, best_idx, second_dist, second_idx


# ---------------------------------------------------------------- IMU pre-integration
class OrcImuNoise(C.Structure):
    _fields_ = [("sigma_g", C.c_double), ("sigma_a", C.c_double), ("sigma_bg", C.c_double), ("sigma_ba", C.c_double),
                ("freq_ref", C.c_double), ("dt_cov_noise_fixed", C.c_int32), ("pad_", C.c_int32)]


PREINT_DTYPE = np.dtype([("Rij", "f8", (3, 3)), ("vij", "f8", 3), ("pij", "f8", 3), ("SigmaPRV", "f8", (9, 9)),
                         ("SigmaPVR", "f8", (9, 9)), ("Jgp", "f8", (3, 3)), ("Jap", "f8", (3, 3)), ("Jgv", "f8", (3, 3)),
                         ("Jav", "f8", (3, 3)), ("JgR", "f8", (3, 3)), ("dt", "f8"), ("status", "i4"), ("pad_", "i4")])

EUROC_IMU_SIGMA = (1.6968e-4, 2.0e-3, 1.9393e-5, 3.0e-3)  # Examples/Stereo/EuRoC/EuRoC_VIO.yaml:13-18


def imu_noise(sigma=EUROC_IMU_SIGMA, dt_cov_noise_fixed=1, freq_ref=200.0):
    """IMUDataBase::SetParam with the caller-side squaring of src/Tracking.cc:744-745."""
    nz = OrcImuNoise()
    s2 = (C.c_double * 4)(*[s * s for s in sigma])
    lib().orc_imu_set_param(C.byref(nz), s2, dt_cov_noise_fixed, C.c_double(freq_ref))
    return nz


def imu_preintegrate(samples, ti, tj, bg, ba, nz):
    """samples: (n,7) rows {t, ax, ay, az, wx, wy, wz} -> PREINT_DTYPE record"""
    L = lib()
    L.orc_imu_preintegrate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    samples = np.ascontiguousarray(samples, np.float64).reshape(-1, 7)
    bg = np.ascontiguousarray(bg, np.float64); ba = np.ascontiguousarray(ba, np.float64)
    out = np.zeros(1, PREINT_DTYPE)
    L.orc_imu_preintegrate(_p(samples), len(samples), ti, tj, _p(bg), _p(ba), C.byref(nz), _p(out))
    return out[0]


# ---------------------------------------------------------------- bundle adjustment (oracle/ba_oracle.cc)
import sys as _sys  # noqa: E402
_sys.path.insert(0, _ROOT)
from vieo_slam_b200.layouts import (BA_RESULT_DTYPE, CAMERA_DTYPE, NAVSTATE_DTYPE, POSEOPT_PROBLEM_DTYPE,  # noqa: E402,F401
                                    POSEOPT_RESULT_DTYPE)


class OrcBaProblem(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_imu", C.c_int32),
                ("states", C.c_void_p), ("state_flags", C.c_void_p), ("points", C.c_void_p), ("edge_state", C.c_void_p),
                ("edge_point", C.c_void_p), ("obs", C.c_void_p), ("inv_sigma2", C.c_void_p), ("edge_flags", C.c_void_p),
                ("imu_i", C.c_void_p), ("imu_j", C.c_void_p), ("preint", C.c_void_p), ("imu_dt_kf", C.c_void_p),
                ("gw", C.c_double * 3), ("inv_sigma_bg2", C.c_double), ("inv_sigma_ba2", C.c_double),
                ("large", C.c_int32), ("rec_init", C.c_int32), ("visual_only", C.c_int32), ("global_ba", C.c_int32)]


def ba_problem_struct(d, large=False, rec_init=False, visual_only=False, cls=OrcBaProblem):
    """dict from synth.make_lba_problem -> (ctypes struct, keepalive list)."""
    keep = {k: np.ascontiguousarray(d[k]) for k in ("states", "state_flags", "points", "edge_state", "edge_point", "obs",
                                                    "inv_sigma2", "edge_flags", "imu_i", "imu_j", "preint", "imu_dt_kf")}
    pb = cls()
    pb.n_states, pb.n_points, pb.n_edges, pb.n_imu = len(keep["states"]), len(keep["points"]), len(keep["edge_state"]), len(keep["imu_i"])
    for k, a in keep.items():
        setattr(pb, k, a.ctypes.data)
    pb.gw = (C.c_double * 3)(*d["gw"])
    pb.inv_sigma_bg2, pb.inv_sigma_ba2 = d["inv_sigma_bg2"], d["inv_sigma_ba2"]
    pb.large, pb.rec_init, pb.visual_only = int(large), int(rec_init), int(visual_only)
    return pb, keep


def pose_optimization(pbs, cam, Xw, obs, inv_sigma2, flags):
    """Run orc_pose_optimization on every problem of `pbs` -> (results, outlier u8 [E], chi2 f64 [E])."""
    L = lib()
    L.orc_pose_optimization.argtypes = [C.c_void_p] * 9
    pbs = np.ascontiguousarray(pbs); cam = np.ascontiguousarray(cam)
    Xw = np.ascontiguousarray(Xw, np.float64); obs = np.ascontiguousarray(obs, np.float32)
    inv_sigma2 = np.ascontiguousarray(inv_sigma2, np.float32); flags = np.ascontiguousarray(flags, np.uint8)
    res = np.zeros(len(pbs), POSEOPT_RESULT_DTYPE)
    outlier = np.zeros(len(flags), np.uint8); chi2 = np.zeros(len(flags), np.float64)
    for k in range(len(pbs)):
        L.orc_pose_optimization(pbs[k:k + 1].ctypes.data, cam.ctypes.data, _p(Xw), _p(obs), _p(inv_sigma2), _p(flags),
                                res[k:k + 1].ctypes.data, _p(outlier), _p(chi2))
    return res, outlier, chi2


def local_ba_prv(d, cam, **kw):
    L = lib()
    L.orc_local_ba_prv.argtypes = [C.c_void_p] * 7
    pb, keep = ba_problem_struct(d, **kw)
    cam = np.ascontiguousarray(cam)
    st = np.zeros(pb.n_states, NAVSTATE_DTYPE); pts = np.zeros((pb.n_points, 3)); chi2 = np.zeros(pb.n_edges)
    erase = np.zeros(pb.n_edges, np.uint8); res = np.zeros(1, BA_RESULT_DTYPE)
    L.orc_local_ba_prv(C.byref(pb), cam.ctypes.data, _p(st), _p(pts), _p(chi2), _p(erase), _p(res))
    return dict(states=st, points=pts, edge_chi2=chi2, erase=erase, res=res[0])


def global_ba_prv(d, cam, n_iterations=10, robust=False):
    pb, keep = ba_problem_struct(d)
    L = lib()
    cam = np.ascontiguousarray(cam)
    st = np.zeros(len(d["states"]), NAVSTATE_DTYPE); pts = np.zeros((len(d["points"]), 3))
    chi2 = np.zeros(len(d["edge_state"])); res = np.zeros(1, BA_RESULT_DTYPE)
    L.orc_global_ba_prv.argtypes = None
    it = L.orc_global_ba_prv(C.byref(pb), C.c_void_p(cam.ctypes.data), C.c_int(n_iterations), C.c_int(int(robust)), _p(st), _p(pts),
                             _p(chi2), _p(res))
    return dict(states=st, points=pts, edge_chi2=chi2, res=res[0], iterations=it)


def ba_debug_step(d, cam, lam, **kw):
    L = lib()
    L.orc_ba_debug_step.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    pb, keep = ba_problem_struct(d, **kw)
    cam = np.ascontiguousarray(cam)
    xp = np.zeros(15 * pb.n_states); xl = np.zeros((pb.n_points, 3)); chi2 = np.zeros(1)
    n = L.orc_ba_debug_step(C.byref(pb), cam.ctypes.data, lam, _p(xp), _p(xl), _p(chi2))
    assert n >= 0, n
    return xp[:n], xl, chi2[0]


def edge_reproject(cam, ns, Xw, obs, stereo):
    L = lib()
    L.orc_edge_reproject.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    cam = np.ascontiguousarray(cam); ns = np.ascontiguousarray(ns)
    Xw = np.ascontiguousarray(Xw, np.float64); obs = np.ascontiguousarray(obs, np.float32)
    e = np.zeros(3); Jp = np.zeros((3, 6)); JX = np.zeros((3, 3)); d = np.zeros(1)
    L.orc_edge_reproject(cam.ctypes.data, ns.ctypes.data, _p(Xw), _p(obs), int(stereo), _p(e), _p(Jp), _p(JX), _p(d))
    return e, Jp, JX, d[0]


def edge_navstate(nsi, nsj, pre, gw, order):
    L = lib()
    L.orc_edge_navstate.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4
    nsi = np.ascontiguousarray(nsi); nsj = np.ascontiguousarray(nsj); pre = np.ascontiguousarray(pre)
    gw = np.ascontiguousarray(gw, np.float64)
    e = np.zeros(9); Ji = np.zeros((9, 9)); Jj = np.zeros((9, 9)); Jb = np.zeros((9, 6))
    L.orc_edge_navstate(nsi.ctypes.data, nsj.ctypes.data, pre.ctypes.data, _p(gw), order, _p(e), _p(Ji), _p(Jj), _p(Jb))
    return e, Ji, Jj, Jb


def navstate_oplus(ns, kind, dx):
    L = lib()
    L.orc_navstate_oplus.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    out = np.array(ns, dtype=NAVSTATE_DTYPE).reshape(1).copy()
    dx = np.ascontiguousarray(dx, np.float64)
    L.orc_navstate_oplus(out.ctypes.data, kind, _p(dx))
    return out[0]


def edge_prior_pvr(ns, prior):
    L = lib()
    L.orc_edge_prior_pvr.argtypes = [C.c_void_p] * 4
    ns = np.ascontiguousarray(ns); prior = np.ascontiguousarray(prior)
    e = np.zeros(15); J = np.zeros((15, 9))
    L.orc_edge_prior_pvr(ns.ctypes.data, prior.ctypes.data, _p(e), _p(J))
    return e, J


def imu_preintegrate_frames(seq, idx, nz):
    """PREINT records for consecutive entries of frame-index list idx (record 0 is the identity state)."""
    out = np.zeros(len(idx), PREINT_DTYPE)
    out["Rij"] = np.eye(3); out["JgR"] = 0
    imu, t = seq["imu"], seq["times"]
    for k in range(1, len(idx)):
        ti, tj = t[idx[k - 1]], t[idx[k]]
        lo = max(np.searchsorted(imu[:, 0], ti, "right") - 1, 0)
        hi = min(np.searchsorted(imu[:, 0], tj, "left") + 1, len(imu))
        out[k] = imu_preintegrate(imu[lo:hi], ti, tj, seq["truth"][idx[k - 1]]["bg"], seq["truth"][idx[k - 1]]["ba"], nz)
    return out


def ba_debug_system(d, cam, lam, lambda_on_poses=True, **kw):
    L = lib()
    L.orc_ba_debug_system.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int] + [C.c_void_p] * 4
    pb, keep = ba_problem_struct(d, **kw)
    cam = np.ascontiguousarray(cam)
    n = 15 * pb.n_states
    S = np.zeros(n * n); bs = np.zeros(n); b = np.zeros(n); chi2 = np.zeros(1)
    m = L.orc_ba_debug_system(C.byref(pb), cam.ctypes.data, lam, int(lambda_on_poses), _p(S), _p(bs), _p(b), _p(chi2))
    assert m >= 0, m
    return S[:m * m].reshape(m, m), bs[:m], b[:m], chi2[0]


def stereo_matches(orbL, kl, dl, orbR, kr, dr, bf, minZ):
    """Frame::ComputeStereoMatches on two OrbOracle instances that have just extracted the left / right image
    (their pyramids are read back).  -> (uright f32[nl], depth f32[nl], sad i32[nl], kept)"""
    L = lib()
    L.orc_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    n = orbL.nlevels
    lvL = [orbL.level(l) for l in range(n)]
    lvR = [orbR.level(l) for l in range(n)]
    pl = (C.c_void_p * n)(*[a.ctypes.data for a in lvL]); pr = (C.c_void_p * n)(*[a.ctypes.data for a in lvR])
    lw = np.array([a.shape[1] for a in lvL], np.int32); lh = np.array([a.shape[0] for a in lvL], np.int32)
    tb = orbL.tables()
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32); sad = np.empty(len(kl), np.int32)
    kept = L.orc_stereo_matches(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), pl, pr, _p(lw), _p(lh), _p(tb["scale"]),
                                _p(tb["inv_scale"]), C.c_float(bf), C.c_float(minZ), _p(ur), _p(dp), _p(sad))
    return ur, dp, sad, kept


# ---- guided searches (oracle/sbp_oracle.cc) ----------------------------------------------------------------------------
def search_by_projection(pb, frames=None):
    """Runs the oracle over every frame of a synth.make_sbp_problem batch -> (kp_match, q_match, q_dist, n_matches)."""
    L = lib()
    for name in ("orc_sbp_last_frame", "orc_sbp_local_map"):
        getattr(L, name).argtypes = None
        getattr(L, name).restype = C.c_int
    fr = pb["frames"]
    kp_match = np.full(len(pb["kps"]), -1, np.int32)   # entries outside every frame's range stay -1
    q_match = np.full(len(pb["q_level"]), -1, np.int32); q_dist = np.full(len(pb["q_level"]), -1, np.int32)
    nm = np.zeros(len(fr), np.int32)
    vp = C.c_void_p

    def at(a, i):
        a = pb[a]
        return vp(a.ctypes.data + i * a.strides[0])
    for f in (range(len(fr)) if frames is None else frames):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        blk = at("kp_blocked", kb) if pb.get("kp_blocked") is not None else None
        outs = (vp(kp_match.ctypes.data + 4 * kb), vp(q_match.ctypes.data + 4 * qb), vp(q_dist.ctypes.data + 4 * qb))
        fp = vp(fr.ctypes.data + f * fr.strides[0])
        if pb["mode"] == 0:
            nm[f] = L.orc_sbp_last_frame(fp, at("kps", kb), at("uright", kb), at("desc", kb), at("q_Xw", qb), at("q_level", qb),
                                         at("q_angle", qb), at("q_desc", qb), at("q_flags", qb), blk, *outs)
        else:
            nm[f] = L.orc_sbp_local_map(fp, at("kps", kb), at("uright", kb), at("desc", kb), at("q_proj", qb), at("q_level", qb),
                                        at("q_viewcos", qb), at("q_depth", qb), at("q_desc", qb), at("q_flags", qb), blk, *outs)
    return kp_match, q_match, q_dist, nm


# ---- Frame::isInFrustum / MapPoint::PredictScale (oracle/sbp_oracle.cc) --------------------------------------------------
def predict_scale(max_distance, current_dist, log_scale_factor, n_levels):
    L = lib()
    L.orc_predict_scale.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    return L.orc_predict_scale(max_distance, current_dist, log_scale_factor, n_levels)


def is_in_frustum(pb, frames=None):
    """Oracle over every frame of a synth.make_frustum_problem dict -> dict(inview, proj, level, viewcos, depth, n_inview);
    points with p_skip set are left "not in view" (Tracking::SearchLocalPoints never tests them)."""
    from vieo_slam_b200.layouts import ORC_FRUSTUM_FRAME_DTYPE
    L = lib()
    L.orc_is_in_frustum.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
    ff = pb["frustum"]
    n = len(pb["p_max_dist"])
    out = dict(inview=np.zeros(n, np.uint8), proj=np.zeros((n, 3), np.float32), level=np.full(n, -1, np.int32),
               viewcos=np.zeros(n, np.float32), depth=np.zeros(n, np.float32), n_inview=np.zeros(len(ff), np.int32))
    skip = pb.get("p_skip")
    for f in (range(len(ff)) if frames is None else frames):
        of = np.zeros(1, ORC_FRUSTUM_FRAME_DTYPE)
        for name in ORC_FRUSTUM_FRAME_DTYPE.names:
            of[0][name] = ff[f][name]
        b, m = int(ff[f]["q_begin"]), int(ff[f]["n_q"])
        idx = np.arange(b, b + m)
        if skip is not None:
            idx = idx[skip[b:b + m] == 0]
        if len(idx) == 0:
            continue
        a = [np.ascontiguousarray(pb[k][idx], np.float32) for k in ("p_wP", "p_normal", "p_max_dist", "p_min_dist")]
        o = [np.zeros(len(idx), np.uint8), np.zeros((len(idx), 3), np.float32), np.zeros(len(idx), np.int32),
             np.zeros(len(idx), np.float32), np.zeros(len(idx), np.float32)]
        out["n_inview"][f] = L.orc_is_in_frustum(_p(of), len(idx), *[_p(x) for x in a], *[_p(x) for x in o])
        for key, arr in zip(("inview", "proj", "level", "viewcos", "depth"), o):
            out[key][idx] = arr
    return out


def search_local_points(pb, frames=None):
    """isInFrustum then the local-map SearchByProjection of the oracle, the latter fed with what the former left."""
    fo = is_in_frustum(pb, frames)
    q = dict(pb)
    q["mode"] = 1
    q["q_proj"], q["q_level"], q["q_viewcos"], q["q_depth"] = fo["proj"], fo["level"], fo["viewcos"], fo["depth"]
    return (fo,) + search_by_projection(q, frames)


# ---- MapPoint::ComputeDistinctiveDescriptors (oracle/match_oracle.cc) ------------------------------------------------------
def distinctive_descriptors(desc_pool, ptr, rows=None):
    L = lib()
    L.orc_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_distinctive_descriptors.restype = None
    pool = np.ascontiguousarray(desc_pool, np.uint8).reshape(-1, 32)
    ptr = np.ascontiguousarray(ptr, np.int32)
    rows = None if rows is None else np.ascontiguousarray(rows, np.int32)
    n = len(ptr) - 1
    best = np.empty(n, np.int32); med = np.empty(n, np.int32)
    L.orc_distinctive_descriptors(_p(pool), None if rows is None else _p(rows), _p(ptr), n, _p(best), _p(med))
    return best, med


# ---- Optimizer::OptimizeInitialGyroBias (oracle/imu_oracle.cc) -----------------------------------------------------------
def gyro_bias_init(pre, Rwb, use_info=True):
    """-> (num_equations, dbg)"""
    L = lib()
    L.orc_gyro_bias_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    pre = np.ascontiguousarray(pre, PREINT_DTYPE)
    Rwb = np.ascontiguousarray(Rwb, np.float64).reshape(-1, 9)
    dbg = np.zeros(3)
    n = L.orc_gyro_bias_init(_p(pre), _p(Rwb), len(pre), int(use_info), _p(dbg))
    return n, dbg


# ---- ORBmatcher::SearchByProjectionBase, search half (oracle/sbp_oracle.cc) -------------------------------------------------
def proj_search(pb):
    """Oracle over every keyframe of a synth.make_fuse_problem dict -> (best_idx, best_dist, level) per map point."""
    L = lib()
    L.orc_sbp_base.argtypes = [C.c_void_p] * 13
    L.orc_sbp_base.restype = None
    fr = pb["frames"]
    nq = len(pb["p_max_dist"])
    best = np.full(nq, -1, np.int32); dist = np.full(nq, -1, np.int32); lvl = np.full(nq, -1, np.int32)
    vp = C.c_void_p

    def at(a, i):
        a = pb[a] if isinstance(a, str) else a
        return vp(a.ctypes.data + i * a.strides[0])
    for f in range(len(fr)):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        skip = at("p_skip", qb) if pb.get("p_skip") is not None else None
        L.orc_sbp_base(vp(fr.ctypes.data + f * fr.strides[0]), at("kps", kb), at("uright", kb), at("desc", kb), at("p_wP", qb),
                       at("p_normal", qb), at("p_max_dist", qb), at("p_min_dist", qb), at("q_desc", qb), skip, at(best, qb),
                       at(dist, qb), at(lvl, qb))
    return best, dist, lvl


# ---- scale / gravity-direction variants of the global-BA edges (oracle/ba_oracle.cc; restated ahead of the device side)
def edge_reproject_scale(cam, ns, Xh, scale, obs, stereo):
    """EdgeReprojectPRS[Stereo] -> (e, J_pose (3,6), J_point (3,3), J_scale (3,))"""
    L = lib()
    L.orc_edge_reproject_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.orc_edge_reproject_scale.restype = None
    cam = np.ascontiguousarray(cam); ns = np.ascontiguousarray(ns)
    Xh = np.ascontiguousarray(Xh, np.float64); obs = np.ascontiguousarray(obs, np.float32)
    e = np.zeros(3); Jp = np.zeros((3, 6)); JX = np.zeros((3, 3)); Js = np.zeros(3)
    L.orc_edge_reproject_scale(cam.ctypes.data, ns.ctypes.data, _p(Xh), float(scale), _p(obs), int(stereo), _p(e), _p(Jp),
                               _p(JX), _p(Js))
    return e, Jp, JX, Js


def gdir_init(gw):
    """VertexGThetaXYRwI::setToOriginImpl(gw) -> unit quaternion (w, x, y, z) of RwI"""
    L = lib()
    L.orc_gdir_init.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_gdir_init.restype = None
    gw = np.ascontiguousarray(gw, np.float64); q = np.zeros(4)
    L.orc_gdir_init(_p(gw), _p(q))
    return q


def gdir_oplus(q, d):
    L = lib()
    L.orc_gdir_oplus.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_gdir_oplus.restype = None
    q = np.array(q, np.float64).copy(); d = np.ascontiguousarray(d, np.float64)
    L.orc_gdir_oplus(_p(q), _p(d))
    return q


def edge_navstate_g(nsi, nsj, pre, q_wI, GI):
    """EdgeNavStatePRVG -> (e, Ji, Jj, Jb, JG (9,2)); residual / column order P, R, V"""
    L = lib()
    L.orc_edge_navstate_g.argtypes = [C.c_void_p] * 10
    L.orc_edge_navstate_g.restype = None
    nsi = np.ascontiguousarray(nsi); nsj = np.ascontiguousarray(nsj); pre = np.ascontiguousarray(pre)
    q_wI = np.ascontiguousarray(q_wI, np.float64); GI = np.ascontiguousarray(GI, np.float64)
    e = np.zeros(9); Ji = np.zeros((9, 9)); Jj = np.zeros((9, 9)); Jb = np.zeros((9, 6)); JG = np.zeros((9, 2))
    L.orc_edge_navstate_g(nsi.ctypes.data, nsj.ctypes.data, pre.ctypes.data, _p(q_wI), _p(GI), _p(e), _p(Ji), _p(Jj), _p(Jb),
                          _p(JG))
    return e, Ji, Jj, Jb, JG


def global_ba_prv_scale(d, cam, n_iterations=10, robust=False):
    """GlobalBundleAdjustmentNavStatePRV with bScaleOpt = true -> dict incl. `scale`; points are returned scaled."""
    pb, keep = ba_problem_struct(d)
    L = lib()
    cam = np.ascontiguousarray(cam)
    st = np.zeros(len(d["states"]), NAVSTATE_DTYPE); pts = np.zeros((len(d["points"]), 3))
    chi2 = np.zeros(len(d["edge_state"])); res = np.zeros(1, BA_RESULT_DTYPE); sc = np.zeros(1)
    L.orc_global_ba_prv_scale.argtypes = None
    it = L.orc_global_ba_prv_scale(C.byref(pb), C.c_void_p(cam.ctypes.data), C.c_int(n_iterations), C.c_int(int(robust)), _p(st),
                                   _p(pts), _p(chi2), _p(res), _p(sc))
    return dict(states=st, points=pts, edge_chi2=chi2, res=res[0], iterations=it, scale=float(sc[0]))


def ba_debug_step_scale(d, cam, lam, scale0, **kw):
    L = lib()
    L.orc_ba_debug_step_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    pb, keep = ba_problem_struct(d, **kw)
    cam = np.ascontiguousarray(cam)
    xp = np.zeros(15 * pb.n_states + 1); xl = np.zeros((pb.n_points, 3)); chi2 = np.zeros(1)
    n = L.orc_ba_debug_step_scale(C.byref(pb), cam.ctypes.data, lam, scale0, _p(xp), _p(xl), _p(chi2))
    assert n >= 0, n
    return xp[:n], xl, chi2[0]


def global_ba_prv_init(d, cam, n_iterations, gw):
    """GlobalBundleAdjustmentNavStatePRV with an IMU initiator -> dict incl. the refined gravity `gw`."""
    pb, keep = ba_problem_struct(d)
    L = lib()
    cam = np.ascontiguousarray(cam)
    st = np.zeros(len(d["states"]), NAVSTATE_DTYPE); pts = np.zeros((len(d["points"]), 3))
    chi2 = np.zeros(len(d["edge_state"])); res = np.zeros(1, BA_RESULT_DTYPE); g = np.array(gw, np.float64).copy()
    L.orc_global_ba_prv_init.argtypes = None
    it = L.orc_global_ba_prv_init(C.byref(pb), C.c_void_p(cam.ctypes.data), C.c_int(n_iterations), _p(g), _p(st), _p(pts), _p(chi2),
                                  _p(res))
    return dict(states=st, points=pts, edge_chi2=chi2, res=res[0], iterations=it, gw=g)


def ba_debug_step_gdir(d, cam, lam, gw, **kw):
    L = lib()
    L.orc_ba_debug_step_gdir.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    pb, keep = ba_problem_struct(d, **kw)
    cam = np.ascontiguousarray(cam); gw = np.ascontiguousarray(gw, np.float64)
    xp = np.zeros(15 * pb.n_states + 2); xl = np.zeros((pb.n_points, 3)); chi2 = np.zeros(1)
    n = L.orc_ba_debug_step_gdir(C.byref(pb), cam.ctypes.data, lam, _p(gw), _p(xp), _p(xl), _p(chi2))
    assert n >= 0, n
    return xp[:n], xl, chi2[0]


def search_for_triangulation(pb, p):
    """ORBmatcher::SearchForTriangulation for pair p of a synth.make_sft_problem dict -> (pairs [n, 2], nmatches)."""
    P = pb["pairs"][p]
    L = lib()
    L.orc_search_for_triangulation.restype = C.c_int
    L.orc_search_for_triangulation.argtypes = ([C.c_void_p] * 7 + [C.c_int]) * 2 + [C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                                                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]

    def kf(kb, n, nb, nn, pb_, ib):
        ptr = np.ascontiguousarray(pb["fv_ptr"][pb_:pb_ + nn + 1], np.int32)
        return [np.ascontiguousarray(pb["kps"][kb:kb + n]), np.ascontiguousarray(pb["uright"][kb:kb + n], np.float32),
                np.ascontiguousarray(pb["desc"][kb:kb + n], np.uint8), np.ascontiguousarray(pb["has_mp"][kb:kb + n], np.uint8),
                np.ascontiguousarray(pb["fv_node"][nb:nb + nn], np.int32), ptr,
                np.ascontiguousarray(pb["fv_idx"][ib:ib + ptr[-1]], np.int32)]
    a = kf(P["kp1_begin"], P["n_kp1"], P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"])
    b = kf(P["kp2_begin"], P["n_kp2"], P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"])
    F = np.ascontiguousarray(P["F12"], np.float64); sf = np.ascontiguousarray(P["scale_factor2"], np.float32)
    s2 = np.ascontiguousarray(P["level_sigma2_2"], np.float32)
    cap = int(P["n_kp1"]) + 1
    out = np.full((cap, 2), -1, np.int32)
    n = L.orc_search_for_triangulation(*[_p(x) for x in a], int(P["n_nodes1"]), *[_p(x) for x in b], int(P["n_nodes2"]), _p(F),
                                       float(P["ex"]), float(P["ey"]), _p(sf), _p(s2), int(P["only_stereo"]),
                                       int(P["check_orientation"]), _p(out), cap)
    return out[:n], n


def search_by_bow(pb, p, mp_id=None):
    """ORBmatcher::SearchByBoW(KeyFrame, Frame) for pair p of a synth.make_bow_problem dict -> (match_f [n_kp2], nmatches)."""
    P = pb["pairs"][p]
    L = lib()
    L.orc_search_by_bow.restype = C.c_int
    L.orc_search_by_bow.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_float, C.c_int, C.c_void_p]

    def fv(nb, nn, pb_, ib):
        ptr = np.ascontiguousarray(pb["fv_ptr"][pb_:pb_ + nn + 1], np.int32)
        return [np.ascontiguousarray(pb["fv_node"][nb:nb + nn], np.int32), ptr, np.ascontiguousarray(pb["fv_idx"][ib:ib + ptr[-1]], np.int32)]
    k1 = slice(int(P["kp1_begin"]), int(P["kp1_begin"] + P["n_kp1"])); k2 = slice(int(P["kp2_begin"]), int(P["kp2_begin"] + P["n_kp2"]))
    mp_id_default = np.where(pb["mp_ok"][k1] != 0, np.arange(int(P["n_kp1"])), -1).astype(np.int32)  # one camera: a map point per keypoint
    mp_id = mp_id_default if mp_id is None else np.ascontiguousarray(mp_id, np.int32)
    a = fv(P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"]); b = fv(P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"])
    kk1 = np.ascontiguousarray(pb["kps"][k1]); dd1 = np.ascontiguousarray(pb["desc"][k1], np.uint8)
    kk2 = np.ascontiguousarray(pb["kps"][k2]); dd2 = np.ascontiguousarray(pb["desc"][k2], np.uint8)
    mf = np.empty(max(int(P["n_kp2"]), 1), np.int32)
    n = L.orc_search_by_bow(_p(kk1), _p(dd1), _p(mp_id), *[_p(x) for x in a], int(P["n_nodes1"]), _p(kk2), _p(dd2), int(P["n_kp2"]),
                            *[_p(x) for x in b], int(P["n_nodes2"]), float(P["nn_ratio"]), int(P["check_orientation"]), _p(mf))
    return mf[:int(P["n_kp2"])], n


def edge_sim3(cam, ns, scale, Xh, obs, inverse):
    """orc_edge_sim3: EdgeReprojectPRS (inverse = 0) / EdgeReprojectPRSInv (1) -> (e[2], J_pose[2][6], J_scale[2])"""
    L = lib()
    L.orc_edge_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_edge_sim3.restype = None
    cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1); ns = np.ascontiguousarray(ns, NAVSTATE_DTYPE).reshape(1)
    Xh = np.ascontiguousarray(Xh, np.float64); obs = np.ascontiguousarray(obs, np.float32)
    e = np.zeros(2); Jp = np.zeros((2, 6)); Js = np.zeros(2)
    L.orc_edge_sim3(_p(cam), _p(ns), float(scale), _p(Xh), _p(obs), int(inverse), _p(e), _p(Jp), _p(Js))
    return e, Jp, Js


def optimize_sim3(pbs, cam, Xc1, Xc2, obs1, obs2, w1, w2):
    """orc_optimize_sim3 on every problem -> (results, keep u8[M], chi2_12[M], chi2_21[M])"""
    from vieo_slam_b200.layouts import SIM3_PROBLEM_DTYPE, SIM3_RESULT_DTYPE
    L = lib()
    L.orc_optimize_sim3.argtypes = [C.c_void_p] * 12
    pbs = np.ascontiguousarray(pbs, SIM3_PROBLEM_DTYPE); cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
    Xc1 = np.ascontiguousarray(Xc1, np.float64); Xc2 = np.ascontiguousarray(Xc2, np.float64)
    obs1 = np.ascontiguousarray(obs1, np.float32); obs2 = np.ascontiguousarray(obs2, np.float32)
    w1 = np.ascontiguousarray(w1, np.float32); w2 = np.ascontiguousarray(w2, np.float32)
    M = len(Xc1)
    res = np.zeros(len(pbs), SIM3_RESULT_DTYPE); keep = np.zeros(M, np.uint8); c12 = np.zeros(M); c21 = np.zeros(M)
    for k in range(len(pbs)):
        L.orc_optimize_sim3(pbs[k:k + 1].ctypes.data, _p(cam), _p(Xc1), _p(Xc2), _p(obs1), _p(obs2), _p(w1), _p(w2),
                            res[k:k + 1].ctypes.data, _p(keep), _p(c12), _p(c21))
    return res, keep, c12, c21


def sbp_reloc(pb, frames=None):
    """orc_sbp_reloc over every frame of a synth.make_reloc_problem batch -> (kp_match, q_match, q_dist, q_level, n_matches)"""
    L = lib()
    L.orc_sbp_reloc.restype = C.c_int
    L.orc_sbp_reloc.argtypes = [C.c_void_p, C.c_int, C.c_float] + [C.c_void_p] * 12
    fr = pb["frames"]
    nq = len(pb["q_angle"])
    kp_match = np.full(len(pb["kps"]), -1, np.int32)
    q_match = np.full(nq, -1, np.int32); q_dist = np.full(nq, -1, np.int32); q_level = np.full(nq, -1, np.int32)
    nm = np.zeros(len(fr), np.int32)

    def at(a, i):
        a = pb[a] if isinstance(a, str) else a
        return a.ctypes.data + i * a.strides[0]
    for f in (range(len(fr)) if frames is None else frames):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        blk = at("kp_blocked", kb) if pb.get("kp_blocked") is not None else None
        nm[f] = L.orc_sbp_reloc(fr.ctypes.data + f * fr.strides[0], int(pb["reloc"][f]["orb_dist"]),
                                float(pb["reloc"][f]["log_scale_factor"]), at("kps", kb), at("desc", kb), at("q_Xw", qb),
                                at("q_angle", qb), at("q_max_dist", qb), at("q_min_dist", qb), at("q_desc", qb), blk,
                                at(kp_match, kb), at(q_match, qb), at(q_dist, qb), at(q_level, qb))
    return kp_match, q_match, q_dist, q_level, nm


# ---- Optimizer::OptimizeEssentialGraph (oracle/posegraph_oracle.cc) ---------------------------------------------------------
def _pg():
    L = lib()
    if not getattr(L, "_pg_ready", False):
        for f, n in (("orc_sim3_exp", 2), ("orc_sim3_log", 2), ("orc_sim3_mul", 3), ("orc_sim3_inv", 2)):
            getattr(L, f).argtypes = [C.c_void_p] * n
            getattr(L, f).restype = None
        L.orc_essential_graph_recover_se3.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_essential_graph_recover_se3.restype = None
        L.orc_essential_graph_correct_points.argtypes = [C.c_int] + [C.c_void_p] * 5
        L.orc_essential_graph_correct_points.restype = None
        L._pg_ready = True
    return L


def _sim3(a):
    from vieo_slam_b200.layouts import SIM3_DTYPE
    return np.ascontiguousarray(a, SIM3_DTYPE)


def sim3_exp(u):
    from vieo_slam_b200.layouts import SIM3_DTYPE
    out = np.zeros(1, SIM3_DTYPE); u = np.ascontiguousarray(u, np.float64)
    _pg().orc_sim3_exp(_p(u), _p(out))
    return out[0]


def sim3_log(S):
    S = _sim3(S).reshape(1); out = np.zeros(7)
    _pg().orc_sim3_log(_p(S), _p(out))
    return out


def sim3_mul(a, b):
    from vieo_slam_b200.layouts import SIM3_DTYPE
    a = _sim3(a).reshape(1); b = _sim3(b).reshape(1); out = np.zeros(1, SIM3_DTYPE)
    _pg().orc_sim3_mul(_p(a), _p(b), _p(out))
    return out[0]


def sim3_inv(a):
    from vieo_slam_b200.layouts import SIM3_DTYPE
    a = _sim3(a).reshape(1); out = np.zeros(1, SIM3_DTYPE)
    _pg().orc_sim3_inv(_p(a), _p(out))
    return out[0]


def sim3_from_Rt(R, t, s=1.0):
    from vieo_slam_b200.layouts import SIM3_DTYPE
    R = np.ascontiguousarray(R, np.float64); t = np.ascontiguousarray(t, np.float64); out = np.zeros(1, SIM3_DTYPE)
    L = lib()
    L.orc_sim3_from_Rt.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.orc_sim3_from_Rt(_p(R), _p(t), float(s), _p(out))
    return out[0]


def edge_sim3_graph(meas, v0, v1, fix0=False, fix1=False, fix_scale=False, jac=True):
    """EdgeSim3::computeError (+ the numeric Jacobians of BaseBinaryEdge::linearizeOplus) -> (e[7], Ji[7][7], Jj[7][7])"""
    L = lib()
    L.orc_edge_sim3_graph.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 3
    L.orc_edge_sim3_graph.restype = None
    meas = _sim3(meas).reshape(1); v0 = _sim3(v0).reshape(1); v1 = _sim3(v1).reshape(1)
    e = np.zeros(7); Ji = np.zeros((7, 7)); Jj = np.zeros((7, 7))
    L.orc_edge_sim3_graph(_p(meas), _p(v0), _p(v1), int(fix0), int(fix1), int(fix_scale), _p(e), _p(Ji) if jac else None,
                          _p(Jj) if jac else None)
    return e, Ji, Jj


def essential_graph(pb, iterations=20, lambda_init=1e-16, single_step=False, want_system=False):
    """orc_essential_graph on a synth.make_essential_graph problem -> (Scw_out SIM3_DTYPE[K], stats dict[, H, b])"""
    from vieo_slam_b200.layouts import SIM3_DTYPE
    L = lib()
    L.orc_essential_graph.restype = C.c_int
    L.orc_essential_graph.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    S = _sim3(pb["Scw"]); fixed = np.ascontiguousarray(pb["fixed"], np.uint8)
    ei = np.ascontiguousarray(pb["ei"], np.int32); ej = np.ascontiguousarray(pb["ej"], np.int32); meas = _sim3(pb["meas"])
    info = None if pb.get("info") is None else np.ascontiguousarray(pb["info"], np.float64)
    out = np.zeros(len(S), SIM3_DTYPE); stats = np.zeros(5)
    K = len(S)
    nfree = 7 * K
    H = np.zeros((nfree, nfree)) if want_system else None
    b = np.zeros(nfree) if want_system else None
    n = L.orc_essential_graph(K, _p(S), _p(fixed), int(pb["fix_scale"]), len(ei), _p(ei), _p(ej), _p(meas), None if info is None else _p(info),
                              int(iterations), float(lambda_init), int(single_step), _p(out), _p(stats), None if H is None else _p(H),
                              None if b is None else _p(b))
    assert n >= 0, n
    st = dict(chi2_initial=stats[0], chi2_final=stats[1], iterations=int(stats[2]), lambda_final=stats[3], trials=int(stats[4]), n=n)
    if want_system:
        Hn = H.reshape(-1)[:n * n].reshape(n, n).copy()
        return out, st, Hn, b[:n].copy()
    return out, st


def essential_graph_recover_se3(S):
    S = _sim3(S); T = np.zeros((len(S), 3, 4))
    _pg().orc_essential_graph_recover_se3(len(S), _p(S), _p(T))
    return T


def essential_graph_correct_points(Pw, ref, S_before, S_after):
    Pw = np.ascontiguousarray(Pw, np.float32); ref = np.ascontiguousarray(ref, np.int32)
    out = np.zeros_like(Pw)
    _pg().orc_essential_graph_correct_points(len(Pw), _p(Pw), _p(ref), _p(_sim3(S_before)),
                                             _p(_sim3(S_after)), _p(out))
    return out


def is_in_frustum_rig(pb):
    """orc_is_in_frustum_rig over every frame of a synth.make_frustum_rig_problem dict -> dict(inview, cam_mask, proj [n][4][3],
    level [n][4], viewcos [n][4], depth, n_inview); points with p_skip set stay "not in view"."""
    L = lib()
    L.orc_is_in_frustum_rig.restype = C.c_int
    L.orc_is_in_frustum_rig.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
    rig = pb["rig"]
    n = len(pb["p_max_dist"])
    out = dict(inview=np.zeros(n, np.uint8), cam_mask=np.zeros(n, np.uint8), proj=np.zeros((n, 4, 3), np.float32),
               level=np.full((n, 4), -1, np.int32), viewcos=np.zeros((n, 4), np.float32), depth=np.zeros(n, np.float32),
               n_inview=np.zeros(len(rig), np.int32))
    skip = pb.get("p_skip")
    for f in range(len(rig)):
        b, m = int(rig[f]["q_begin"]), int(rig[f]["n_q"])
        idx = np.arange(b, b + m)
        if skip is not None:
            idx = idx[skip[b:b + m] == 0]
        if len(idx) == 0:
            continue
        a = [np.ascontiguousarray(pb[k][idx], np.float32) for k in ("p_wP", "p_normal", "p_max_dist", "p_min_dist")]
        k = len(idx)
        o = [np.zeros(k, np.uint8), np.zeros(k, np.uint8), np.zeros((k, 4, 3), np.float32), np.zeros((k, 4), np.int32),
             np.zeros((k, 4), np.float32), np.zeros(k, np.float32)]
        one = np.ascontiguousarray(rig[f:f + 1])
        out["n_inview"][f] = L.orc_is_in_frustum_rig(_p(one), k, *[_p(x) for x in a], *[_p(x) for x in o])
        for key, arr in zip(("inview", "cam_mask", "proj", "level", "viewcos", "depth"), o):
            out[key][idx] = arr
    return out


def cam_project(cam, P, want_jac=True):
    """orc_cam_project: the oracle's camera Project() -> (uv f32[2], J f64[2][3] | None)"""
    L = lib()
    L.orc_cam_project.argtypes = [C.c_void_p] * 4
    L.orc_cam_project.restype = None
    cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1); P = np.ascontiguousarray(P, np.float64)
    uv = np.zeros(2, np.float32); J = np.zeros((2, 3)) if want_jac else None
    L.orc_cam_project(_p(cam), _p(P), _p(uv), None if J is None else _p(J))
    return uv, J


# ---- g2o's Levenberg-Marquardt control flow over callbacks (oracle/lm_oracle.cc; oracle/_ref compiles the reference's own) -----
class OrcLmCallbacks(C.Structure):
    _F = {"d_v": C.CFUNCTYPE(C.c_double, C.c_void_p), "v_v": C.CFUNCTYPE(None, C.c_void_p), "i_vd": C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_double),
          "p_v": C.CFUNCTYPE(C.c_void_p, C.c_void_p), "d_vi": C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int), "i_v": C.CFUNCTYPE(C.c_int, C.c_void_p)}
    _fields_ = [("ctx", C.c_void_p), ("n", C.c_int), ("errors", _F["d_v"]), ("build", _F["v_v"]), ("solve", _F["i_vd"]), ("update", _F["v_v"]),
                ("push", _F["v_v"]), ("pop", _F["v_v"]), ("discard_top", _F["v_v"]), ("x", _F["p_v"]), ("b", _F["p_v"]),
                ("hessian_diag", _F["d_vi"]), ("terminate", _F["i_v"])]


def lm_callbacks(problem):
    """Wrap a python problem object (methods errors / build / solve / update / push / pop / discard_top / hessian_diag / terminate,
    persistent float64 arrays x and b, attribute n) into an OrcLmCallbacks struct; keep the returned object alive during the call."""
    F = OrcLmCallbacks._F
    cb = OrcLmCallbacks()
    cb.ctx = None
    cb.n = int(problem.n)
    cb.errors = F["d_v"](lambda c: float(problem.errors()))
    cb.build = F["v_v"](lambda c: problem.build())
    cb.solve = F["i_vd"](lambda c, lam: int(bool(problem.solve(lam))))
    cb.update = F["v_v"](lambda c: problem.update())
    cb.push = F["v_v"](lambda c: problem.push())
    cb.pop = F["v_v"](lambda c: problem.pop())
    cb.discard_top = F["v_v"](lambda c: problem.discard_top())
    cb.x = F["p_v"](lambda c: problem.x.ctypes.data)
    cb.b = F["p_v"](lambda c: problem.b.ctypes.data)
    cb.hessian_diag = F["d_vi"](lambda c, j: float(problem.hessian_diag(j)))
    cb.terminate = F["i_v"](lambda c: int(bool(problem.terminate())))
    return cb


def lm_optimize(problem, iterations, user_lambda_init=0.0, driver=None):
    """orc_lm_optimize (or another driver with the same signature, e.g. ref_lib.lib().ref_lm_optimize) -> stats[5]"""
    cb = lm_callbacks(problem)
    stats = np.zeros(5)
    fn = driver
    if fn is None:
        fn = lib().orc_lm_optimize
        fn.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        fn.restype = C.c_int
    fn(C.byref(cb), int(iterations), float(user_lambda_init), _p(stats))
    return stats


def essential_graph_lm(pb, driver_ptr, iterations=20, lambda_init=1e-16):
    """orc_essential_graph_lm with an explicit LM driver (a C function pointer as an int / c_void_p; None: the oracle's own)"""
    from vieo_slam_b200.layouts import SIM3_DTYPE
    L = lib()
    L.orc_essential_graph_lm.restype = C.c_int
    L.orc_essential_graph_lm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    S = _sim3(pb["Scw"]); fixed = np.ascontiguousarray(pb["fixed"], np.uint8)
    ei = np.ascontiguousarray(pb["ei"], np.int32); ej = np.ascontiguousarray(pb["ej"], np.int32); meas = _sim3(pb["meas"])
    info = None if pb.get("info") is None else np.ascontiguousarray(pb["info"], np.float64)
    out = np.zeros(len(S), SIM3_DTYPE); stats = np.zeros(5)
    n = L.orc_essential_graph_lm(len(S), _p(S), _p(fixed), int(pb["fix_scale"]), len(ei), _p(ei), _p(ej), _p(meas),
                                 None if info is None else _p(info), int(iterations), float(lambda_init), driver_ptr, _p(out), _p(stats))
    assert n >= 0
    return out, stats


def imu_preintegrate_trace(samples, ti, tj, bg, ba, cap=4096):
    """orc_imu_preintegrate_trace: the update() calls of the oracle's PreIntegration -> (status, trace [n][7])"""
    L = lib()
    L.orc_imu_preintegrate_trace.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p]
    L.orc_imu_preintegrate_trace.restype = C.c_int
    smp = np.ascontiguousarray(samples, np.float64).reshape(-1, 7)
    bg = np.ascontiguousarray(bg, np.float64); ba = np.ascontiguousarray(ba, np.float64)
    tr = np.zeros((cap, 7)); n = C.c_int(0)
    rc = L.orc_imu_preintegrate_trace(_p(smp), len(smp), float(ti), float(tj), _p(bg), _p(ba), _p(tr), cap, C.byref(n))
    return rc, tr[:n.value].copy()


def so3(op, x):
    """orc_so3: op 0 exp -> q (w,x,y,z), 1 Exp -> R, 2 log(q), 3 Log(R), 4 JacobianR, 5 JacobianRInv, 6 normalizeRotationM"""
    L = lib()
    L.orc_so3.argtypes = [C.c_int, C.c_void_p, C.c_void_p]; L.orc_so3.restype = None
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros({0: 4, 2: 3, 3: 3}.get(op, 9))
    L.orc_so3(op, _p(x), _p(out))
    return out.reshape(3, 3) if out.size == 9 else out


class MargDump(C.Structure):
    """OrcMargDump (oracle/ba_oracle.h): what Optimizer::FillCovInv sees and yields inside orc_pose_optimization"""
    _fields_ = [("filled", C.c_int32), ("has_imu", C.c_int32), ("fixed_last", C.c_int32), ("n_vis", C.c_int32), ("cap", C.c_int32),
                ("pad_", C.c_int32), ("info_imu", C.c_double * 81), ("info_bias", C.c_double * 36), ("info_prior", C.c_double * 225),
                ("delta_imu", C.c_double), ("delta_bias", C.c_double), ("delta_prior", C.c_double),
                ("C", C.c_double * 225), ("CL", C.c_double * 225), ("CCL", C.c_double * 225),
                ("level", C.c_void_p), ("delta", C.c_void_p)]


def pose_optimization_marg_dump(pb, cam, Xw, obs, inv_sigma2, flags):
    """orc_pose_optimization on ONE problem with the marginal dump hook set -> (result, dump dict)"""
    L = lib()
    L.orc_set_marg_dump.argtypes = [C.c_void_p]
    L.orc_set_marg_dump.restype = None
    E = int(pb["edge_end"][0] - pb["edge_begin"][0])
    level = np.zeros(max(E, 1), np.int32); delta = np.zeros(max(E, 1), np.float64)
    d = MargDump()
    d.cap = E; d.level = level.ctypes.data; d.delta = delta.ctypes.data
    L.orc_set_marg_dump(C.addressof(d))
    try:
        res, outl, chi2 = pose_optimization(pb, cam, Xw, obs, inv_sigma2, flags)
    finally:
        L.orc_set_marg_dump(None)
    out = dict(filled=int(d.filled), has_imu=int(d.has_imu), fixed_last=int(d.fixed_last), n_vis=int(d.n_vis),
               info_imu=np.array(d.info_imu).reshape(9, 9), info_bias=np.array(d.info_bias).reshape(6, 6),
               info_prior=np.array(d.info_prior).reshape(15, 15), delta_imu=float(d.delta_imu), delta_bias=float(d.delta_bias),
               delta_prior=float(d.delta_prior), C=np.array(d.C).reshape(15, 15), CL=np.array(d.CL).reshape(15, 15),
               CCL=np.array(d.CCL).reshape(15, 15), level=level[:E].copy(), delta=delta[:E].copy())
    return res[0], out

"""ctypes binding of oracle/_ref/libref.so — TEST INFRASTRUCTURE ONLY.

libref.so is the REFERENCE compiled here where it compiles, unchanged from where it lies, by oracle/ref_build/Makefile (needs
/root/reference; the prebuilt .so travels to the GPU box): src/ORBextractor.cc whole, and — cut out by name at build time — the
matchers (every SearchByProjection overload, SearchByProjectionBase, Fuse, SearchByBoW, SearchForTriangulation with the cameras'
epipolarConstrain / FillMatchesFromPair), the frame grid, ComputeStereoMatches, ComputeStereoFishEyeMatches, isInFrustum, the camera
models' Project(), the IMU pre-integrator, NavState / so3_extra, every vertex / edge class of g2otypes with the fork's getHessian*
members, Optimizer::FillCovInv, g2o's Levenberg-Marquardt control flow, Huber kernel and Sim3 (DESIGN.md section 2 lists what each
wrapper pins and which stand-ins it declares).  It exists to pin the oracle restatement (tests/test_oracle_ref.py asserts
oracle == _ref) and to be the checker of tests/test_gpu_vs_ref.py.  Nothing under vieo_slam_b200/ may load it."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_lib import KP_DTYPE, _p

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_ref", "libref.so")
REF_SRC = "/root/reference"


def available():
    """True when libref.so exists or can be built (the reference sources are present)."""
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REF_SRC, "src"))


def build(force=False):
    if os.path.isdir(os.path.join(REF_SRC, "src")):  # make decides whether anything is stale
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle", "ref_build")] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libref.so missing and /root/reference absent: cannot build the compiled reference")
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = _lib = C.CDLL(build())
        i32p = C.POINTER(C.c_int32)
        L.ref_orb_create.restype = C.c_void_p
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [C.c_void_p]
        L.ref_orb_tables.argtypes = [C.c_void_p] * 7
        L.ref_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, i32p]
        L.ref_orb_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.ref_orb_get_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_quadtree.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_int]
        L.ref_ic_angle.restype = C.c_float
        L.ref_ic_angle.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        L.ref_orb_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_void_p]
        L.ref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_cam_project.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_cam_project.restype = None
        L.ref_predict_scale.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
        L.ref_predict_scale.restype = C.c_int
        L.ref_lm_optimize.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.ref_lm_optimize.restype = C.c_int
        L.ref_imu_preintegrate_trace.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_imu_preintegrate_trace.restype = C.c_int
    return _lib


class RefOrb:
    """The reference's ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST), compiled here."""

    def __init__(self, nfeatures=1200, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.h = lib().ref_orb_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_orb_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        q, um = np.empty(n, np.int32), np.empty(16, np.int32)
        lib().ref_orb_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(q), _p(um))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, quota=q, umax=um)

    def extract(self, img, lapping=None, cap=8192):
        """operator(): returns (n, kps, desc, ret) like OrbOracle.extract (ret = monoIndex or -1)."""
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int32(0)
        lap = None if lapping is None else np.asarray(lapping, np.int32)
        if img is None or img.size == 0:
            ret = lib().ref_orb_extract(self.h, None, 0, 0, 0, None, _p(kps), _p(desc), cap, C.byref(n))
            return 0, kps[:0], desc[:0], ret
        img = np.ascontiguousarray(img, np.uint8)
        ret = lib().ref_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                    None if lap is None else _p(lap), _p(kps), _p(desc), cap, C.byref(n))
        assert ret != -2, "capacity"
        return n.value, kps[:n.value], desc[:n.value], ret

    def level(self, l):
        w, h = C.c_int32(), C.c_int32()
        assert lib().ref_orb_level_size(self.h, l, C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.uint8)
        lib().ref_orb_get_level(self.h, l, _p(out))
        return out

    def quadtree(self, xyr, minX, maxX, minY, maxY, N, level=0):
        xyr = np.ascontiguousarray(xyr, np.int32)
        out = np.empty((max(len(xyr), 1), 3), np.int32)
        n = lib().ref_quadtree(self.h, _p(xyr), len(xyr), minX, maxX, minY, maxY, N, level, _p(out), len(out))
        assert n >= 0
        return out[:n]

    def ic_angle(self, img, x, y):
        img = np.ascontiguousarray(img, np.uint8)
        return lib().ref_ic_angle(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], float(x), float(y))


def orb_descriptor(blurred, x, y, angle_deg):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    d = np.empty(32, np.uint8)
    lib().ref_orb_descriptor(_p(blurred), blurred.shape[1], blurred.shape[0], blurred.strides[0], float(x), float(y),
                             float(angle_deg), _p(d))
    return d


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a), _p(b))


def three_maxima(sizes):
    sizes = np.ascontiguousarray(sizes, np.int32)
    out = np.empty(3, np.int32)
    lib().ref_three_maxima(_p(sizes), len(sizes), _p(out))
    return tuple(int(v) for v in out)


def cam_project(model, params, P, want_jac=True):
    """camm::{Pinhole (0), Radtan (1), KB8 (2)}Camera::Project of the reference, compiled unchanged -> (uv f32[2], J f64[2][3] | None)"""
    params = np.ascontiguousarray(params, np.float32); P = np.ascontiguousarray(P, np.float64)
    uv = np.zeros(2, np.float32); J = np.zeros((2, 3)) if want_jac else None
    lib().ref_cam_project(int(model), _p(params), len(params), _p(P), _p(uv), None if J is None else _p(J), None)
    return uv, J


def predict_scale(max_distance, current_dist, log_scale_factor, n_levels):
    """MapPoint::PredictScale of the reference, compiled unchanged"""
    return int(lib().ref_predict_scale(float(max_distance), float(current_dist), float(log_scale_factor), int(n_levels)))


def imu_preintegrate_trace(samples, ti, tj, bg, ba, cap=4096):
    """IMUPreIntegratorBase::PreIntegration of the reference, compiled unchanged, with a recording update() ->
    (status, trace [n][7] = (omega, acc, dt) per update() call, mdeltatij after the call started from 123)"""
    smp = np.ascontiguousarray(samples, np.float64).reshape(-1, 7)
    bg = np.ascontiguousarray(bg, np.float64); ba = np.ascontiguousarray(ba, np.float64)
    tr = np.zeros((cap, 7)); n = C.c_int(0); dtj = C.c_double(0)
    rc = lib().ref_imu_preintegrate_trace(_p(smp), len(smp), float(ti), float(tj), _p(bg), _p(ba), _p(tr), cap, C.byref(n), C.byref(dtj))
    return rc, tr[:n.value].copy(), dtj.value


def stereo_matches(orbL, kl, dl, orbR, kr, dr, bf, minZ):
    """Frame::ComputeStereoMatches of the reference, compiled unchanged, on the pyramids of two ORACLE extractor instances (the same
    inputs as oracle_lib.stereo_matches) -> (uright f32[nl], depth f32[nl], kept)"""
    L = lib()
    L.ref_stereo_matches.restype = C.c_int
    L.ref_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    n = orbL.nlevels
    lvL = [np.ascontiguousarray(orbL.level(l)) for l in range(n)]
    lvR = [np.ascontiguousarray(orbR.level(l)) for l in range(n)]
    pl = (C.c_void_p * n)(*[a.ctypes.data for a in lvL]); pr = (C.c_void_p * n)(*[a.ctypes.data for a in lvR])
    lw = np.array([a.shape[1] for a in lvL], np.int32); lh = np.array([a.shape[0] for a in lvL], np.int32)
    tb = orbL.tables()
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
    kept = L.ref_stereo_matches(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), pl, pr, _p(lw), _p(lh), n, _p(tb["scale"]),
                                _p(tb["inv_scale"]), float(bf), float(minZ), _p(ur), _p(dp))
    return ur, dp, kept


def search_by_projection_last_frame(pb):
    """ORBmatcher::SearchByProjection(Frame&, const Frame& last, ...) of the reference, compiled unchanged, over every frame of a
    synth.make_sbp_problem(mode = SBP_LAST_FRAME) batch -> (kp_match, n_matches)"""
    L = lib()
    L.ref_sbp_last_frame.argtypes = None
    L.ref_sbp_last_frame.restype = C.c_int
    fr = pb["frames"]
    kp_match = np.full(len(pb["kps"]), -1, np.int32)
    nm = np.zeros(len(fr), np.int32)
    vp = C.c_void_p

    def at(a, i):
        a = pb[a]
        return vp(a.ctypes.data + i * a.strides[0])
    for f in range(len(fr)):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        blk = at("kp_blocked", kb) if pb.get("kp_blocked") is not None else None
        nm[f] = L.ref_sbp_last_frame(vp(fr.ctypes.data + f * fr.strides[0]), at("kps", kb), at("uright", kb), at("desc", kb), at("q_Xw", qb),
                                     at("q_level", qb), at("q_angle", qb), at("q_desc", qb), at("q_flags", qb), blk,
                                     vp(kp_match.ctypes.data + 4 * kb))
    return kp_match, nm


def search_by_projection_local_map(pb):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, th_far_pts) of the reference, compiled unchanged, over
    every frame of a synth.make_sbp_problem(mode = SBP_LOCAL_MAP) batch -> (kp_match, n_matches)"""
    L = lib()
    L.ref_sbp_local_map.argtypes = None
    L.ref_sbp_local_map.restype = C.c_int
    fr = pb["frames"]
    kp_match = np.full(len(pb["kps"]), -1, np.int32)
    nm = np.zeros(len(fr), np.int32)
    vp = C.c_void_p

    def at(a, i):
        a = pb[a]
        return vp(a.ctypes.data + i * a.strides[0])
    for f in range(len(fr)):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        blk = at("kp_blocked", kb) if pb.get("kp_blocked") is not None else None
        nm[f] = L.ref_sbp_local_map(vp(fr.ctypes.data + f * fr.strides[0]), at("kps", kb), at("uright", kb), at("desc", kb), at("q_proj", qb),
                                    at("q_level", qb), at("q_viewcos", qb), at("q_depth", qb), at("q_desc", qb), at("q_flags", qb), blk,
                                    vp(kp_match.ctypes.data + 4 * kb))
    return kp_match, nm


# ---- the inertial arithmetic (oracle/ref_build/ref_inertial_wrap.cc): so3_extra.h, NavState.h, IMUPreIntegratorBase::update and the
# inertial vertices / edges of g2otypes compiled unchanged against the Eigen / Sophus stand-in
def so3(op, x):
    """op 0 exp -> q (w,x,y,z), 1 Exp -> R, 2 SO3ex(q).log(), 3 Log(R), 4 JacobianR, 5 JacobianRInv, 6 normalizeRotationM,
    7 Sophus' base-class log() of q"""
    L = lib()
    L.ref_so3.argtypes = [C.c_int, C.c_void_p, C.c_void_p]; L.ref_so3.restype = None
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros({0: 4, 2: 3, 3: 3, 7: 3}.get(op, 9))
    L.ref_so3(op, x.ctypes.data, out.ctypes.data)
    return out.reshape(3, 3) if out.size == 9 else out


def imu_update_sequence(nz, trace, preint_dtype):
    """reset state + IMUPreIntegratorBase::update per row of trace [n][7] = (omega, acc, dt) -> PREINT record"""
    L = lib()
    L.ref_imu_update_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]; L.ref_imu_update_sequence.restype = None
    tr = np.ascontiguousarray(trace, np.float64).reshape(-1, 7)
    out = np.zeros(1, preint_dtype)
    L.ref_imu_update_sequence(C.byref(nz), tr.ctypes.data, len(tr), out.ctypes.data)
    return out[0]


def edge_navstate(nsi, nsj, pre, gw, order, q_wI=None):
    """EdgeNavStateI<3 | 5 | 6>: -> (e, Ji, Jj, Jb[, JG]); with q_wI, gw is GI and the edge is EdgeNavStatePRVG"""
    L = lib()
    L.ref_edge_navstate.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 6; L.ref_edge_navstate.restype = None
    nsi = np.ascontiguousarray(nsi); nsj = np.ascontiguousarray(nsj); pre = np.ascontiguousarray(pre)
    gw = np.ascontiguousarray(gw, np.float64)
    e = np.zeros(9); Ji = np.zeros((9, 9)); Jj = np.zeros((9, 9)); Jb = np.zeros((9, 6)); JG = np.zeros((9, 2))
    q = None if q_wI is None else np.ascontiguousarray(q_wI, np.float64)
    L.ref_edge_navstate(nsi.ctypes.data, nsj.ctypes.data, pre.ctypes.data, gw.ctypes.data, int(order), None if q is None else q.ctypes.data,
                        e.ctypes.data, Ji.ctypes.data, Jj.ctypes.data, Jb.ctypes.data, JG.ctypes.data)
    return (e, Ji, Jj, Jb) if q is None else (e, Ji, Jj, Jb, JG)


def navstate_oplus(ns, kind, dx):
    L = lib()
    L.ref_navstate_oplus.argtypes = [C.c_void_p, C.c_int, C.c_void_p]; L.ref_navstate_oplus.restype = None
    out = np.array(ns).reshape(1).copy()
    dx = np.ascontiguousarray(dx, np.float64)
    L.ref_navstate_oplus(out.ctypes.data, int(kind), dx.ctypes.data)
    return out[0]


def gdir_init(gw):
    L = lib()
    L.ref_gdir.argtypes = [C.c_int, C.c_void_p, C.c_void_p]; L.ref_gdir.restype = None
    gw = np.ascontiguousarray(gw, np.float64); q = np.zeros(4)
    L.ref_gdir(0, gw.ctypes.data, q.ctypes.data)
    return q


def gdir_oplus(q, d):
    L = lib()
    L.ref_gdir.argtypes = [C.c_int, C.c_void_p, C.c_void_p]; L.ref_gdir.restype = None
    q = np.array(q, np.float64).copy(); d = np.ascontiguousarray(d, np.float64)
    L.ref_gdir(1, d.ctypes.data, q.ctypes.data)
    return q


def edge_prior(form, ns, prior):
    """form 0 EdgeNavStatePriorPVRBias, 1 EdgeNavStatePriorPRVBias -> (e [15], J_state [15][9], J_bias [15][6])"""
    L = lib()
    L.ref_edge_prior.argtypes = [C.c_int] + [C.c_void_p] * 5; L.ref_edge_prior.restype = None
    ns = np.ascontiguousarray(ns); prior = np.ascontiguousarray(prior)
    e = np.zeros(15); J = np.zeros((15, 9)); Jb = np.zeros((15, 6))
    L.ref_edge_prior(int(form), ns.ctypes.data, prior.ctypes.data, e.ctypes.data, J.ctypes.data, Jb.ctypes.data)
    return e, J, Jb


def edge_bias(nsi, nsj):
    L = lib()
    L.ref_edge_bias.argtypes = [C.c_void_p] * 5; L.ref_edge_bias.restype = None
    nsi = np.ascontiguousarray(nsi); nsj = np.ascontiguousarray(nsj)
    e = np.zeros(6); Ji = np.zeros((6, 6)); Jj = np.zeros((6, 6))
    L.ref_edge_bias(nsi.ctypes.data, nsj.ctypes.data, e.ctypes.data, Ji.ctypes.data, Jj.ctypes.data)
    return e, Ji, Jj


def edge_gyr_bias(dRij, JgRij, Rwbi, Rwbj, bg):
    L = lib()
    L.ref_edge_gyr_bias.argtypes = [C.c_void_p] * 7; L.ref_edge_gyr_bias.restype = None
    a = [np.ascontiguousarray(x, np.float64) for x in (dRij, JgRij, Rwbi, Rwbj, bg)]
    e = np.zeros(3); J = np.zeros((3, 3))
    L.ref_edge_gyr_bias(*[x.ctypes.data for x in a], e.ctypes.data, J.ctypes.data)
    return e, J


def edge_reproject(form, cam, ns, X, obs, scale=1.0):
    """EdgeReproject<DE, DV, NV, MODE> of the reference compiled unchanged over its own Project(): form 0 PR, 1 PRStereo, 2 PVR, 3 PVRStereo,
    4 PRS, 5 PRSStereo, 6 PRSInv -> (e [DE], J_pose [DE][DV], J_point [DE][3], J_scale [DE], depth)"""
    L = lib()
    L.ref_edge_reproject.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 5; L.ref_edge_reproject.restype = None
    DE = 3 if form in (1, 3, 5) else 2
    DV = 9 if form in (2, 3) else 6
    cam = np.ascontiguousarray(cam).reshape(1); ns = np.ascontiguousarray(ns).reshape(1)
    X = np.ascontiguousarray(X, np.float64); obs = np.ascontiguousarray(obs, np.float32)
    e = np.zeros(DE); Jp = np.zeros((DE, DV)); JX = np.zeros((DE, 3)); Js = np.zeros(DE); d = np.zeros(1)
    L.ref_edge_reproject(int(form), cam.ctypes.data, ns.ctypes.data, X.ctypes.data, obs.ctypes.data, float(scale), e.ctypes.data,
                         Jp.ctypes.data, JX.ctypes.data, Js.ctypes.data, d.ctypes.data)
    return e, Jp, JX, Js, d[0]


# ---- g2o's Sim3 / VertexSim3Expmap / EdgeSim3 + the numeric-Jacobian linearizeOplus (oracle/ref_build/ref_sim3_wrap.cc)
def _sim3_rec(a):
    from vieo_slam_b200.layouts import SIM3_DTYPE
    return np.ascontiguousarray(a, SIM3_DTYPE).reshape(1)


def sim3(op, a=None, b=None, u=None):
    """op 0 exp(u) -> S, 1 log(a) -> u, 2 a * b, 3 a^-1, 4 VertexSim3Expmap::oplusImpl(u) on a (b truthy: _fix_scale)"""
    from vieo_slam_b200.layouts import SIM3_DTYPE
    L = lib()
    L.ref_sim3.argtypes = [C.c_int] + [C.c_void_p] * 4; L.ref_sim3.restype = None
    out = np.zeros(1, SIM3_DTYPE)
    uu = np.zeros(7) if u is None else np.array(u, np.float64).copy()
    ar = None if a is None else _sim3_rec(a)
    if op == 4:
        br = np.zeros(1, SIM3_DTYPE); br["s"] = 1.0 if b else 0.0
    else:
        br = None if b is None else _sim3_rec(b)
    L.ref_sim3(int(op), None if ar is None else ar.ctypes.data, None if br is None else br.ctypes.data, uu.ctypes.data, out.ctypes.data)
    return uu if op == 1 else out[0]


def edge_sim3_graph(meas, v0, v1, fix0=False, fix1=False, fix_scale=False):
    L = lib()
    L.ref_edge_sim3_graph.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 3; L.ref_edge_sim3_graph.restype = None
    m, a, b = _sim3_rec(meas), _sim3_rec(v0), _sim3_rec(v1)
    e = np.zeros(7); Ji = np.zeros((7, 7)); Jj = np.zeros((7, 7))
    L.ref_edge_sim3_graph(m.ctypes.data, a.ctypes.data, b.ctypes.data, int(fix0), int(fix1), int(fix_scale), e.ctypes.data, Ji.ctypes.data,
                          Jj.ctypes.data)
    return e, Ji, Jj


def is_in_frustum_rig(pb):
    """Frame::isInFrustum of the reference, compiled unchanged, over every frame of a synth.make_frustum_rig_problem dict; same outputs
    as oracle_lib.is_in_frustum_rig"""
    L = lib()
    L.ref_is_in_frustum_rig.restype = C.c_int
    L.ref_is_in_frustum_rig.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
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
        out["n_inview"][f] = L.ref_is_in_frustum_rig(_p(one), k, *[_p(x) for x in a], *[_p(x) for x in o])
        for key, arr in zip(("inview", "cam_mask", "proj", "level", "viewcos", "depth"), o):
            out[key][idx] = arr
    return out


def search_by_bow(pb, p, mp_id=None):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) of the reference, compiled unchanged, for pair p of a synth.make_bow_problem dict
    -> (match_f [n_kp2], nmatches); same inputs as oracle_lib.search_by_bow"""
    P = pb["pairs"][p]
    L = lib()
    L.ref_search_by_bow.restype = C.c_int
    L.ref_search_by_bow.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_float, C.c_int, C.c_void_p]

    def fv(nb, nn, pb_, ib):
        ptr = np.ascontiguousarray(pb["fv_ptr"][pb_:pb_ + nn + 1], np.int32)
        return [np.ascontiguousarray(pb["fv_node"][nb:nb + nn], np.int32), ptr, np.ascontiguousarray(pb["fv_idx"][ib:ib + ptr[-1]], np.int32)]
    k1 = slice(int(P["kp1_begin"]), int(P["kp1_begin"] + P["n_kp1"])); k2 = slice(int(P["kp2_begin"]), int(P["kp2_begin"] + P["n_kp2"]))
    mp_id_default = np.where(pb["mp_ok"][k1] != 0, np.arange(int(P["n_kp1"])), -1).astype(np.int32)
    mp_id = mp_id_default if mp_id is None else np.ascontiguousarray(mp_id, np.int32)
    a = fv(P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"]); b = fv(P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"])
    kk1 = np.ascontiguousarray(pb["kps"][k1]); dd1 = np.ascontiguousarray(pb["desc"][k1], np.uint8)
    kk2 = np.ascontiguousarray(pb["kps"][k2]); dd2 = np.ascontiguousarray(pb["desc"][k2], np.uint8)
    mf = np.empty(max(int(P["n_kp2"]), 1), np.int32)
    n = L.ref_search_by_bow(_p(kk1), _p(dd1), _p(mp_id), *[_p(x) for x in a], int(P["n_nodes1"]), _p(kk2), _p(dd2), int(P["n_kp2"]),
                            *[_p(x) for x in b], int(P["n_nodes2"]), float(P["nn_ratio"]), int(P["check_orientation"]), _p(mf))
    return mf[:int(P["n_kp2"])], n


def sbp_reloc(pb):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist, th_far_pts) of the reference, compiled unchanged,
    over every frame of a synth.make_reloc_problem batch -> (kp_match, n_matches)"""
    L = lib()
    L.ref_sbp_reloc.restype = C.c_int
    L.ref_sbp_reloc.argtypes = [C.c_void_p, C.c_int, C.c_float] + [C.c_void_p] * 9
    fr = pb["frames"]
    kp_match = np.full(len(pb["kps"]), -1, np.int32)
    nm = np.zeros(len(fr), np.int32)

    def at(a, i):
        a = pb[a] if isinstance(a, str) else a
        return a.ctypes.data + i * a.strides[0]
    for f in range(len(fr)):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        blk = at("kp_blocked", kb) if pb.get("kp_blocked") is not None else None
        nm[f] = L.ref_sbp_reloc(fr.ctypes.data + f * fr.strides[0], int(pb["reloc"][f]["orb_dist"]), float(pb["reloc"][f]["log_scale_factor"]),
                                at("kps", kb), at("desc", kb), at("q_Xw", qb), at("q_angle", qb), at("q_max_dist", qb), at("q_min_dist", qb),
                                at("q_desc", qb), blk, at(kp_match, kb))
    return kp_match, nm


def proj_search(pb):
    """ORBmatcher::SearchByProjectionBase of the reference, compiled unchanged, over every keyframe of a synth.make_fuse_problem dict
    -> (best_idx, best_dist) per map point (same inputs as oracle_lib.proj_search)"""
    L = lib()
    L.ref_sbp_base.argtypes = [C.c_void_p] * 12
    L.ref_sbp_base.restype = None
    fr = pb["frames"]
    nq = len(pb["p_max_dist"])
    best = np.full(nq, -1, np.int32); dist = np.full(nq, -1, np.int32)
    vp = C.c_void_p

    def at(a, i):
        a = pb[a] if isinstance(a, str) else a
        return vp(a.ctypes.data + i * a.strides[0])
    for f in range(len(fr)):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        skip = at("p_skip", qb) if pb.get("p_skip") is not None else None
        L.ref_sbp_base(vp(fr.ctypes.data + f * fr.strides[0]), at("kps", kb), at("uright", kb), at("desc", kb), at("p_wP", qb),
                       at("p_normal", qb), at("p_max_dist", qb), at("p_min_dist", qb), at("q_desc", qb), skip, at(best, qb), at(dist, qb))
    return best, dist


def distinctive_descriptors(desc_pool, ptr, rows=None):
    """MapPoint::ComputeDistinctiveDescriptors of the reference, compiled unchanged, per CSR point -> best (position in the point's list)"""
    L = lib()
    L.ref_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_distinctive_descriptors.restype = None
    pool = np.ascontiguousarray(desc_pool, np.uint8).reshape(-1, 32)
    ptr = np.ascontiguousarray(ptr, np.int32)
    rows = None if rows is None else np.ascontiguousarray(rows, np.int32)
    best = np.empty(len(ptr) - 1, np.int32)
    L.ref_distinctive_descriptors(_p(pool), None if rows is None else _p(rows), _p(ptr), len(ptr) - 1, _p(best))
    return best


def _quat_wxyz(R):
    """unit quaternion (w, x, y, z) of a rotation matrix"""
    from scipy.spatial.transform import Rotation
    x, y, z, w = Rotation.from_matrix(R).as_quat()
    return np.array([w, x, y, z], np.float64)


def sft_poses(pb, p, seed=0):
    """World poses (q_cw wxyz, t_cw) of the two keyframes of pair p such that Tc1w * Twc2 = (rel_R12, rel_t12): keyframe 1 gets a
    random pose, keyframe 2 follows."""
    from scipy.spatial.transform import Rotation
    r = np.random.default_rng(1000 + seed)
    R1 = Rotation.from_rotvec(r.normal(0, 0.4, 3)).as_matrix(); t1 = r.normal(0, 2.0, 3)
    R12, t12 = pb["rel_R12"][p], pb["rel_t12"][p]
    R2 = R12.T @ R1; t2 = R12.T @ (t1 - t12)          # Tc2w = T21 * Tc1w
    return _quat_wxyz(R1), t1.astype(np.float64), _quat_wxyz(R2), t2.astype(np.float64)


def sft_geometry(K4, q1, t1, q2, t2):
    """(ex, ey, F12 [9]) formed from the poses with the stand-in's operations in the reference's order (ref_sft_wrap.cc)"""
    L = lib()
    L.ref_sft_geometry.restype = None
    L.ref_sft_geometry.argtypes = [C.c_void_p] * 9
    K4 = np.ascontiguousarray(K4, np.float32); ex = np.zeros(1, np.float32); ey = np.zeros(1, np.float32); F = np.zeros(9)
    L.ref_sft_geometry(_p(K4), _p(q1), _p(t1), _p(K4), _p(q2), _p(t2), _p(ex), _p(ey), _p(F))
    return float(ex[0]), float(ey[0]), F


def search_for_triangulation(pb, p, K4, q1, t1, q2, t2):
    """ORBmatcher::SearchForTriangulation of the reference, compiled unchanged (with GeometricCamera::epipolarConstrain and
    FillMatchesFromPair), for pair p of a synth.make_sft_problem dict and keyframe poses -> (pairs [n, 2], nmatches)"""
    P = pb["pairs"][p]
    L = lib()
    L.ref_search_for_triangulation.restype = C.c_int
    side = [C.c_void_p] * 7 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int]
    L.ref_search_for_triangulation.argtypes = side * 2 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]

    def kf(kb, n, nb, nn, pb_, ib):
        ptr = np.ascontiguousarray(pb["fv_ptr"][pb_:pb_ + nn + 1], np.int32)
        return ([np.ascontiguousarray(pb["kps"][kb:kb + n]), np.ascontiguousarray(pb["uright"][kb:kb + n], np.float32),
                 np.ascontiguousarray(pb["desc"][kb:kb + n], np.uint8), np.ascontiguousarray(pb["has_mp"][kb:kb + n], np.uint8)], int(n),
                [np.ascontiguousarray(pb["fv_node"][nb:nb + nn], np.int32), ptr, np.ascontiguousarray(pb["fv_idx"][ib:ib + ptr[-1]], np.int32)], int(nn))
    a, na, fa, nna = kf(P["kp1_begin"], P["n_kp1"], P["node1_begin"], P["n_nodes1"], P["ptr1_begin"], P["idx1_begin"])
    b, nb_, fb, nnb = kf(P["kp2_begin"], P["n_kp2"], P["node2_begin"], P["n_nodes2"], P["ptr2_begin"], P["idx2_begin"])
    K4 = np.ascontiguousarray(K4, np.float32)
    sf = np.ascontiguousarray(P["scale_factor2"][:8], np.float32); s2 = np.ascontiguousarray(P["level_sigma2_2"][:8], np.float32)
    cap = int(P["n_kp1"]) + 1
    out = np.full((cap, 2), -1, np.int32); npairs = np.zeros(1, np.int32)
    n = L.ref_search_for_triangulation(_p(K4), _p(q1), _p(t1), *[_p(x) for x in a], na, *[_p(x) for x in fa], nna,
                                       _p(K4), _p(q2), _p(t2), *[_p(x) for x in b], nb_, *[_p(x) for x in fb], nnb,
                                       _p(sf), _p(s2), 8, int(P["only_stereo"]), int(P["check_orientation"]), _p(out), cap, _p(npairs))
    return out[:int(npairs[0])], n


def fisheye_matches(desc, n_kp, n_mono, octave=None, th_far_pts=0.0):
    """Frame::ComputeStereoFishEyeMatches of the reference, compiled unchanged, for one frame: desc [n_cams][cap][32] ->
    (rec [n, 5] = (cami, idxi, camj, idxj, dist) handed to FillMatchesFromPair in call order, N, mapn2in_ [N, 2], mDescriptors [N, 32])"""
    desc = np.ascontiguousarray(desc, np.uint8)
    n_cams, cap = desc.shape[0], desc.shape[1]
    n_kp = np.ascontiguousarray(n_kp, np.int32); n_mono = np.ascontiguousarray(n_mono, np.int32)
    octave = np.zeros((n_cams, cap), np.int32) if octave is None else np.ascontiguousarray(octave, np.int32)
    rec_cap = 2 * n_cams * n_cams * cap + 1
    rec = np.zeros((rec_cap, 5), np.int32); n_out = np.zeros(1, np.int32)
    tot = int(n_kp.sum())
    mc = np.zeros(max(tot, 1), np.int32); mi = np.zeros(max(tot, 1), np.int32); dout = np.zeros((max(tot, 1), 32), np.uint8)
    L = lib()
    L.ref_fisheye_matches.restype = C.c_int
    L.ref_fisheye_matches.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    n = L.ref_fisheye_matches(_p(desc), _p(octave), _p(n_kp), _p(n_mono), n_cams, cap, float(th_far_pts), _p(rec), rec_cap, _p(n_out),
                              _p(mc), _p(mi), _p(dout))
    assert n <= rec_cap
    N = int(n_out[0])
    return rec[:n], N, np.stack([mc[:N], mi[:N]], 1), dout[:N]


def fill_cov_inv(cam, cur, last, pre, gw, info_imu, delta_imu, info_bias, delta_bias, prior, info_prior, delta_prior, X, obs, stereo, w,
                 level, delta):
    """Optimizer::FillCovInv of the reference (include/Optimizer.h:126-206), compiled unchanged over the compiled edge classes and the
    fork's getHessian() members -> (C, CL, CCL) [15, 15] for schur_bec 0 / 2 / 1 (CL, CCL zero when prior is None)"""
    L = lib()
    L.ref_fill_cov_inv.restype = None
    L.ref_fill_cov_inv.argtypes = ([C.c_void_p] * 6 + [C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
                                   + [C.c_void_p] * 9)
    cam = np.ascontiguousarray(cam).reshape(1); cur = np.ascontiguousarray(cur).reshape(1); last = np.ascontiguousarray(last).reshape(1)
    pre_a = None if pre is None else np.ascontiguousarray(pre).reshape(1)
    prior_a = None if prior is None else np.ascontiguousarray(prior).reshape(1)
    gw = np.ascontiguousarray(gw, np.float64)
    ii = np.ascontiguousarray(info_imu, np.float64); ib = np.ascontiguousarray(info_bias, np.float64); ip = np.ascontiguousarray(info_prior, np.float64)
    X = np.ascontiguousarray(X, np.float64); obs = np.ascontiguousarray(obs, np.float32); stereo = np.ascontiguousarray(stereo, np.uint8)
    w = np.ascontiguousarray(w, np.float64); level = np.ascontiguousarray(level, np.int32); delta = np.ascontiguousarray(delta, np.float64)
    Cm = np.zeros((15, 15)); CL = np.zeros((15, 15)); CCL = np.zeros((15, 15))
    L.ref_fill_cov_inv(_p(cam), _p(cur), _p(last), None if pre_a is None else _p(pre_a), _p(gw), _p(ii), float(delta_imu), _p(ib),
                       float(delta_bias), None if prior_a is None else _p(prior_a), _p(ip), float(delta_prior), len(X), _p(X), _p(obs),
                       _p(stereo), _p(w), _p(level), _p(delta), _p(Cm), _p(CL), _p(CCL))
    return Cm, CL, CCL


def fuse(pb):
    """ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) of the reference, compiled unchanged over the compiled
    SearchByProjectionBase, for every keyframe of a synth.make_fuse_problem dict -> (keypoint FuseMP received per map point or -1,
    nFused per keyframe)"""
    L = lib()
    L.ref_fuse.argtypes = [C.c_void_p] * 11
    L.ref_fuse.restype = C.c_int
    fr = pb["frames"]
    nq = len(pb["p_max_dist"])
    hit = np.full(nq, -1, np.int32); nf = np.zeros(len(fr), np.int32)
    vp = C.c_void_p

    def at(a, i):
        a = pb[a] if isinstance(a, str) else a
        return vp(a.ctypes.data + i * a.strides[0])
    for f in range(len(fr)):
        kb, qb = int(fr[f]["kp_begin"]), int(fr[f]["q_begin"])
        skip = at("p_skip", qb) if pb.get("p_skip") is not None else None
        nf[f] = L.ref_fuse(vp(fr.ctypes.data + f * fr.strides[0]), at("kps", kb), at("uright", kb), at("desc", kb), at("p_wP", qb),
                           at("p_normal", qb), at("p_max_dist", qb), at("p_min_dist", qb), at("q_desc", qb), skip, at(hit, qb))
    return hit, nf


def search_by_sim3(side1, side2, sim3, th, prior12):
    """ORBmatcher::SearchBySim3 of the reference, compiled unchanged over the compiled SearchByProjectionBase (sides / sim3 / prior12 as
    tests/sim3_search_data.make returns them) -> (match12 [n_kp1], nFound, pose21 [12], pose12 [12])"""
    L = lib()
    L.ref_search_by_sim3.restype = C.c_int
    L.ref_search_by_sim3.argtypes = [C.c_void_p] * 20 + [C.c_float, C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 4
    keep = []

    def side(S):
        a = [np.ascontiguousarray(S["frame"]).reshape(1), np.ascontiguousarray(S["kps"]), np.ascontiguousarray(S["uright"], np.float32),
             np.ascontiguousarray(S["desc"], np.uint8), np.ascontiguousarray(S["wP"], np.float32), np.ascontiguousarray(S["Pn"], np.float32),
             np.ascontiguousarray(S["maxd"], np.float32), np.ascontiguousarray(S["mind"], np.float32),
             np.ascontiguousarray(S["qdesc"], np.uint8), np.ascontiguousarray(S["skip"], np.uint8)]
        keep.extend(a)
        return [_p(x) for x in a]
    s12, R12, t12 = sim3
    R12 = np.ascontiguousarray(R12, np.float32); t12 = np.ascontiguousarray(t12, np.float32)
    prior12 = np.ascontiguousarray(prior12, np.int32)
    n1 = len(side1["kps"])
    m12 = np.full(n1, -1, np.int32); p21 = np.zeros(15, np.float32); p12 = np.zeros(15, np.float32)
    n = L.ref_search_by_sim3(*side(side1), *side(side2), float(s12), _p(R12), _p(t12), float(th), _p(prior12), _p(m12), _p(p21), _p(p12))
    return m12, n, p21[:12].copy(), p12[:12].copy()

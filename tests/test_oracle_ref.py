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


# ---- camera models: camm::{Pinhole,Radtan,KB8}Camera::Project compiled unchanged -----------------------------
def _cam_params(cam):
    m = int(cam["model"])
    p = [cam["fx"], cam["fy"], cam["cx"], cam["cy"]]
    if m == 1:
        p += list(cam["dist"][:int(cam["num_k"]) + 2])
    elif m == 2:
        p += list(cam["dist"][:4])
    return m, np.array(p, np.float32)


@pytest.mark.parametrize("model", ["pinhole", "radtan", "radtan3", "kb8"])
def test_camera_project_equal_reference(model):
    """The oracle's Project() (float pixel and d(img)/d(p3d)) against the reference's own function bodies, bit for bit: 2000 random
    camera-frame points incl. points behind the principal plane of the fisheye, on the optical axis (KB8's r <= 1e-5 branch) and
    far off-axis."""
    from vieo_slam_b200 import synth
    cam = {"pinhole": synth.euroc_camera(), "radtan": synth.radtan_camera(), "radtan3": synth.radtan_camera(k=(-0.28, 0.074, -0.01)),
           "kb8": synth.kb8_camera(k=(-0.0135, 0.021, -0.03, 0.012))}[model]
    m, params = _cam_params(cam)
    r = np.random.default_rng(17)
    P = np.concatenate([r.normal(0, 1.5, (1990, 3)) + [0, 0, 3.0],
                        [[0, 0, 2.0], [1e-7, -2e-7, 1.0], [3e-6, 0, 5.0], [5.0, -4.0, 0.5], [0.3, 0.2, -1.0], [1e-3, 1e-3, 1e-3],
                         [2.0, 0.0, 1e-9], [-1.0, 2.0, 8.0], [0.0, 1e-5, 1.0], [7e-6, 7e-6, 1.0]]])
    if model != "kb8":
        P = P[np.abs(P[:, 2]) > 1e-6]      # 1 / z of the plane models
    n_axis = 0
    for p in P:
        uo, Jo = O.cam_project(cam, p)
        ur, Jr = R.cam_project(m, params, p)
        assert uo.tobytes() == ur.tobytes(), (model, p, uo, ur)
        assert np.array_equal(Jo, Jr, equal_nan=True) and (Jo[np.isfinite(Jr)].tobytes() == Jr[np.isfinite(Jr)].tobytes()), (model, p)
        n_axis += np.hypot(p[0], p[1]) <= 1e-5
    assert n_axis >= 3


def test_rig_frustum_projection_equal_reference():
    """The rig visibility test's own double-precision projection (oracle/sbp_oracle.cc project_double) against the compiled reference:
    a rig whose reference frame and camera extrinsics are identities sees the map point at Pc = wP exactly."""
    from vieo_slam_b200 import synth
    for model in (1, 2):
        pb = synth.make_frustum_rig_problem(5, n_frames=1, n_q=400, n_cams=1, model=model, skip_frac=0.0)
        G = pb["rig"][0]
        G["Rcw"] = np.eye(3, dtype=np.float32).ravel(); G["tcw"] = 0; G["Ow"] = 0
        C = G["cam"][0]
        C["q_cr"] = [0, 0, 0, 1]; C["t_cr"] = 0; C["t_rc"] = 0
        C["minx"], C["maxx"], C["miny"], C["maxy"] = -1e9, 1e9, -1e9, 1e9
        G["cos_limit"] = -2.0
        r = np.random.default_rng(3)
        pb["p_wP"][:] = (r.normal(0, 1.0, (400, 3)) + [0, 0, 2.5]).astype(np.float32)
        pb["p_max_dist"][:] = 1e6; pb["p_min_dist"][:] = 0.0
        out = O.is_in_frustum_rig(pb)
        params = np.array([C["fx"], C["fy"], C["cx"], C["cy"], *C["k"]], np.float32)[:8 if model == 2 else 4]
        seen = 0
        for q in range(400):
            if not out["inview"][q]:
                continue
            uv, _ = R.cam_project(2 if model == 2 else 0, params, pb["p_wP"][q].astype(np.float64), want_jac=False)
            assert out["proj"][q, 0, :2].tobytes() == uv.tobytes(), (model, q)
            seen += 1
        assert seen > 300


def test_predict_scale_equal_reference():
    """MapPoint::PredictScale (src/MapPoint.cc:491-509) compiled unchanged against the oracle's restatement: random ratios, the float
    neighbourhood of every level boundary 1.2^k, both pyramid shapes of the configs."""
    r = np.random.default_rng(9)
    for sf, nl in ((1.2, 8), (2.0, 4)):
        lsf = float(np.log(np.float32(sf)))
        cases = [(float(np.float32(mx)), float(np.float32(d))) for mx, d in zip(r.uniform(0.5, 60, 3000), r.uniform(0.2, 40, 3000))]
        for k in range(-2, nl + 2):
            b = np.float32(sf) ** np.float32(k)
            for ulp in range(-6, 7):
                x = b
                for _ in range(abs(ulp)):
                    x = np.nextafter(x, np.float32(np.inf if ulp > 0 else -np.inf), dtype=np.float32)
                cases.append((float(x), 1.0))
        for mx, d in cases:
            assert O.predict_scale(mx, d, lsf, nl) == R.predict_scale(mx, d, lsf, nl), (mx, d, sf)


# ---- g2o's Levenberg-Marquardt control flow: the reference's own solve() / optimize() compiled unchanged ------
class _Lsq:
    """A nonlinear least-squares problem behind the LM callbacks: residual r(theta), Jacobian J(theta); records every lambda the
    driver asks a solve for and every push / pop / discard, so two drivers can be compared event by event."""

    def __init__(self, res, jac, theta0, fail_below=None, stop_after_solves=None):
        self.res, self.jac = res, jac
        self.theta = np.array(theta0, np.float64)
        self.n = len(self.theta)
        self.x = np.zeros(self.n); self.b = np.zeros(self.n); self.H = np.zeros((self.n, self.n))
        self.stack, self.log = [], []
        self.fail_below, self.stop_after = fail_below, stop_after_solves
        self.n_solves = 0

    def errors(self):
        r = self.res(self.theta)
        v = float(r @ r)
        self.log.append(("chi", v))
        return v

    def build(self):
        J, r = self.jac(self.theta), self.res(self.theta)
        self.H[:] = J.T @ J
        self.b[:] = -(J.T @ r)

    def solve(self, lam):
        self.log.append(("lambda", lam))
        self.n_solves += 1
        if self.fail_below is not None and lam < self.fail_below:
            return False          # like LDLT meeting a non-positive pivot: x keeps its previous content
        try:
            c = np.linalg.cholesky(self.H + lam * np.eye(self.n))
        except np.linalg.LinAlgError:
            return False
        self.x[:] = np.linalg.solve(c.T, np.linalg.solve(c, self.b))
        return True

    def update(self):
        self.theta = self.theta + self.x

    def push(self):
        self.stack.append(self.theta.copy()); self.log.append(("push",))

    def pop(self):
        self.theta = self.stack.pop(); self.log.append(("pop",))

    def discard_top(self):
        self.stack.pop(); self.log.append(("discard",))

    def hessian_diag(self, j):
        return self.H[j, j]

    def terminate(self):
        return self.stop_after is not None and self.n_solves >= self.stop_after


def _lm_cases():
    r = np.random.default_rng(4)
    A = r.normal(0, 1, (12, 4)); y = r.normal(0, 1, 12)
    t = np.linspace(0, 2, 25); yexp = 2.0 * np.exp(-1.3 * t) + 0.5 + 0.01 * r.normal(0, 1, 25)
    rosen = (lambda th: np.array([10 * (th[1] - th[0] ** 2), 1 - th[0]]), lambda th: np.array([[-20 * th[0], 10.0], [-1.0, 0.0]]))
    expfit = (lambda th: th[0] * np.exp(th[1] * t) + th[2] - yexp,
              lambda th: np.stack([np.exp(th[1] * t), th[0] * t * np.exp(th[1] * t), np.ones_like(t)], 1))
    # rank-deficient: the second parameter never enters the residual -> H singular, solvable only through the damping
    deficient = (lambda th: np.array([th[0] - 1.0, 2 * th[0] - 2.0]), lambda th: np.array([[1.0, 0.0], [2.0, 0.0]]))
    # a residual that grows when the step is taken (tan blows up): rejections, lambda *= ni, ni *= 2
    wild = (lambda th: np.array([np.tan(th[0]) - 0.3, 0.1 * th[0]]), lambda th: np.array([[1 / np.cos(th[0]) ** 2], [0.1]]))
    return {
        "linear": (dict(res=lambda th: A @ th - y, jac=lambda th: A, theta0=np.zeros(4)), 10, 0.0),
        "linear_user_lambda": (dict(res=lambda th: A @ th - y, jac=lambda th: A, theta0=np.ones(4)), 10, 1e-16),
        "rosenbrock": (dict(res=rosen[0], jac=rosen[1], theta0=[-1.2, 1.0]), 20, 0.0),
        "expfit": (dict(res=expfit[0], jac=expfit[1], theta0=[1.0, -0.5, 0.0]), 15, 0.0),
        "rank_deficient": (dict(res=deficient[0], jac=deficient[1], theta0=[5.0, 3.0]), 8, 1e-300),
        "solver_fails_until_damped": (dict(res=rosen[0], jac=rosen[1], theta0=[-1.2, 1.0], fail_below=5e-2), 12, 1e-6),
        "ten_failures_terminate": (dict(res=rosen[0], jac=rosen[1], theta0=[-1.2, 1.0], fail_below=1e300), 5, 1e-3),
        "wild": (dict(res=wild[0], jac=wild[1], theta0=[1.4]), 12, 1e-9),
        "terminate_flag": (dict(res=expfit[0], jac=expfit[1], theta0=[1.0, -0.5, 0.0], stop_after_solves=3), 15, 0.0),
        "zero_iterations": (dict(res=rosen[0], jac=rosen[1], theta0=[-1.2, 1.0]), 0, 0.0),
    }


@pytest.mark.parametrize("name", list(_lm_cases()))
def test_lm_control_flow_equal_reference(name):
    """The oracle's LM driver against OptimizationAlgorithmLevenberg::solve + SparseOptimizer::optimize of the reference, compiled
    unchanged, on the same callbacks: every lambda handed to the solver, every push / pop / discard, every chi2 evaluation, the
    iteration and trial counts, the final lambda and the final state are identical (the arithmetic is the callbacks', the control
    flow — gain ratio, lambda schedule, the three stop rules, the failed-solve path — is what is compared)."""
    kw, iters, lam0 = _lm_cases()[name]
    a, b = _Lsq(**kw), _Lsq(**kw)
    with np.errstate(all="ignore"):
        sa = O.lm_optimize(a, iters, lam0)
        sb = O.lm_optimize(b, iters, lam0, driver=R.lib().ref_lm_optimize)
    assert a.log == b.log, name   # every chi2 evaluation, lambda, push / pop / discard in the same order with the same values
    assert a.theta.tobytes() == b.theta.tobytes(), name
    assert sa[2] == sb[2] and sa[4] == sb[4], (name, sa, sb)            # iterations, trials
    assert sa[0] == sb[0] and sa[3] == sb[3], (name, sa, sb)            # first chi2, final lambda
    assert sa[1] == sb[1] or (np.isnan(sa[1]) and np.isnan(sb[1])), (name, sa, sb)   # chi2 after the last accepted step
    if name == "ten_failures_terminate":
        assert sa[4] == 10 and sa[2] == 1
    if name == "terminate_flag":
        assert a.n_solves == 3


def test_essential_graph_through_the_reference_lm():
    """The whole essential-graph optimisation of the oracle driven by the reference's compiled LM functions: bit-identical vertices and
    statistics to the oracle's own driver (graphs with accepted-only steps and with rejections at the optimum)."""
    import ctypes as C
    from vieo_slam_b200 import synth
    drv = C.cast(R.lib().ref_lm_optimize, C.c_void_p)
    for K, seed, fs, odom in ((12, 12, True, 0), (40, 3, False, 4), (80, 7, True, 6)):
        pb = synth.make_essential_graph(K=K, seed=seed, fix_scale=fs, odom_info_every=odom, n_neighbors=min(5, K - 5))
        oa, sa = O.essential_graph_lm(pb, None)
        ob, sb = O.essential_graph_lm(pb, drv)
        assert oa.tobytes() == ob.tobytes(), K
        assert sa[0] == sb[0] and sa[2] == sb[2] and sa[3] == sb[3] and sa[4] == sb[4], (K, sa, sb)
        assert sa[1] == sb[1]


def test_ba_drivers_through_the_reference_lm():
    """Every bundle-adjustment driver of the oracle — PoseOptimization (visual and IMU, 4 x optimize(10) with re-classification),
    LocalBundleAdjustmentNavStatePRV (two stages, outlier erasure), GlobalBundleAdjustmentNavStatePRV (plain and with the scale vertex),
    OptimizeSim3 — run with its Levenberg-Marquardt control flow supplied by the REFERENCE's compiled solve() / optimize() (orc_set_lm_driver):
    states, points, outlier / erase sets, chi2 and iteration counts bit-identical to the oracle's own driver."""
    import ctypes as C
    from vieo_slam_b200 import synth
    L = O.lib()
    L.orc_set_lm_driver.argtypes = [C.c_void_p]
    L.orc_set_lm_driver.restype = None
    nz = O.imu_noise()
    seq = synth.vio_sequence(15, 48)
    cam = synth.euroc_camera()
    pre_all = O.imu_preintegrate_frames(seq, list(range(48)), nz)
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, pre_all, cam, n_points=250, seed=2)
    pbs_v = synth.make_pose_problems(seq, pre_all, cam, n_points=250, seed=2, mode=0)
    kf = list(range(0, 48, 3))
    pre_kf = O.imu_preintegrate_frames(seq, kf, nz)
    dl = synth.make_lba_problem(seq, pre_kf, kf, cam, n_local=6, n_fixed=5, n_points=250, seed=6)
    gb = synth.make_gba_problem(seq, pre_kf, kf, cam, n_points=300, seed=5, outlier_frac=0.05)
    cam3 = synth.euroc_camera(); cam3["Rcb"] = np.eye(3); cam3["tcb"] = 0
    s3p = synth.make_sim3_problems(cam3, n_candidates=4, n_matches=100, seed=8, fix_scale=False, few_matches_every=3)
    s3 = (s3p[0], cam3) + tuple(s3p[1:7])

    def run_all():
        a = O.pose_optimization(pbs[:5], cam, X, obs, w, fl)
        b = O.pose_optimization(pbs_v[0][:3], cam, *pbs_v[1:5])
        c = O.local_ba_prv(dl, cam)
        d = O.global_ba_prv(gb, cam, n_iterations=7, robust=True)
        e = O.global_ba_prv_scale(gb, cam, n_iterations=6, robust=False)
        f = O.optimize_sim3(*s3)
        return [f[0].tobytes(), f[1].tobytes(), f[2].tobytes(), f[3].tobytes(), a[0].tobytes(), a[1].tobytes(), a[2].tobytes(), b[0].tobytes(), b[1].tobytes(),
                c["states"].tobytes(), c["points"].tobytes(), c["erase"].tobytes(), c["res"].tobytes(),
                d["states"].tobytes(), d["points"].tobytes(), d["res"].tobytes(),
                e["states"].tobytes(), e["points"].tobytes(), np.float64(e["scale"]).tobytes(), e["res"].tobytes()]
    own = run_all()
    L.orc_set_lm_driver(C.cast(R.lib().ref_lm_optimize, C.c_void_p))
    try:
        ref = run_all()
    finally:
        L.orc_set_lm_driver(None)
    assert [x == y for x, y in zip(own, ref)] == [True] * len(own)
    assert own == run_all()   # the hook is off again


def test_imu_preintegration_sample_selection_equal_reference():
    """IMUPreIntegratorBase::PreIntegration (src/Odom/OdomPreIntegrator.h:227-430) compiled unchanged with a recording update(): the
    oracle issues the same update(omega, acc, dt) calls, bit for bit — forward and reversed time, frame stamps between / on / before /
    after the samples, duplicated samples (dt == 0), a single sample, a > 1.5 s gap (status -1), an empty list."""
    r = np.random.default_rng(12)
    bg, ba = r.normal(0, 0.01, 3), r.normal(0, 0.1, 3)
    n_cases = n_nonempty = n_abort = n_back = 0
    for trial in range(400):
        n = int(r.integers(1, 40))
        t = np.cumsum(r.uniform(0.002, 0.012, n)) + 10.0
        if trial % 7 == 0 and n > 3:
            t[n // 2] = t[n // 2 - 1]                      # a duplicated stamp
        if trial % 11 == 0 and n > 4:
            t[n // 3:] += r.choice([1.2, 1.8, 2.5])         # a gap around the 1.5 s gate
        smp = np.concatenate([t[:, None], r.normal(0, 1, (n, 3)) + [0, 0, 9.8], r.normal(0, 0.3, (n, 3))], 1)
        choices = np.concatenate([t, t + 0.001, t - 0.001, [t[0] - 0.05, t[-1] + 0.05, t[0] - 3.0, t[-1] + 3.0]])
        for _ in range(6):
            ti, tj = (float(x) for x in r.choice(choices, 2))
            so, to = O.imu_preintegrate_trace(smp, ti, tj, bg, ba)
            sr, tr, dtj = R.imu_preintegrate_trace(smp, ti, tj, bg, ba)
            assert so == sr, (trial, ti, tj)
            assert to.shape == tr.shape and to.tobytes() == tr.tobytes(), (trial, ti, tj, to, tr)
            n_cases += 1; n_nonempty += len(to) > 0; n_abort += sr == -1; n_back += ti > tj
            if sr == -1:
                assert dtj == 0.0
    assert n_cases == 2400 and n_nonempty > 1500 and n_abort > 20 and n_back > 800
    # the empty list leaves everything untouched
    so, to = O.imu_preintegrate_trace(np.zeros((0, 7)), 1.0, 2.0, bg, ba)
    sr, tr, dtj = R.imu_preintegrate_trace(np.zeros((0, 7)), 1.0, 2.0, bg, ba)
    assert so == sr == 0 and len(to) == len(tr) == 0 and dtj == 123.0


def test_huber_kernel_equal_reference():
    """g2o::RobustKernelHuber (setDelta + robustify, delta^2 held in a float member in this fork) compiled unchanged against the
    oracle's kernel: rho and rho' bit for bit for the deltas the drivers use and random ones, incl. the e == float(delta^2) boundary."""
    import ctypes as C
    Lo, Lr = O.lib(), R.lib()
    Lo.orc_huber.argtypes = [C.c_double, C.c_double, C.c_void_p]; Lo.orc_huber.restype = None
    Lr.ref_huber.argtypes = [C.c_double, C.c_double, C.c_void_p]; Lr.ref_huber.restype = None
    r = np.random.default_rng(2)
    deltas = [np.sqrt(5.991), np.sqrt(7.815), np.sqrt(16.919), np.sqrt(12.592), 5.0, float(np.sqrt(np.float32(10.0)))] + list(r.uniform(0.1, 9, 40))
    n_out = 0
    for d in deltas:
        ds = float(np.float32(d * d))
        es = list(r.uniform(0, 4 * d * d, 60)) + [ds, np.nextafter(ds, 0), np.nextafter(ds, 1e9), 0.0, 1e-300, 1e12]
        for e in es:
            a = np.zeros(2); b = np.zeros(3)
            Lo.orc_huber(float(d), float(e), a.ctypes.data)
            Lr.ref_huber(float(d), float(e), b.ctypes.data)
            assert a.tobytes() == b[:2].tobytes(), (d, e, a, b)
            n_out += b[1] != 1.0
    assert n_out > 500


def test_compute_stereo_matches_equal_reference():
    """Frame::ComputeStereoMatches (src/Frame.cc:451-611) compiled unchanged against the oracle's restatement on rectified stereo pairs
    (bright, dark and 512 x 512 frames, different disparities): uright and depth of every left keypoint bit for bit."""
    from vieo_slam_b200.synth import EUROC, stereo_stream
    bf = np.float32(EUROC["bf"]); minZ = np.float32(bf / np.float32(EUROC["fx"]))
    imgs = stereo_stream(4, 31, dark_every=2).reshape(4, 2, 480, 752)
    total = 0
    for f in range(4):
        oL, oR = O.OrbOracle(1200, 1.2, 8, 20, 7), O.OrbOracle(1200, 1.2, 8, 20, 7)
        nl, kl, dl, _ = oL.extract(imgs[f, 0]); nr, kr, dr, _ = oR.extract(imgs[f, 1])
        ur, dp, sad, kept = O.stereo_matches(oL, kl, dl, oR, kr, dr, bf, minZ)
        ur2, dp2, kept2 = R.stereo_matches(oL, kl, dl, oR, kr, dr, bf, minZ)
        assert kept == kept2 and ur.tobytes() == ur2.tobytes() and dp.tobytes() == dp2.tobytes(), f
        total += kept
    assert total > 1000


def test_frame_grid_equal_reference():
    """FrameBase::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea / IsInImage (src/FrameBase.cpp:95-174) compiled unchanged: the
    oracle returns the same candidate lists IN THE SAME ORDER for random windows — windows that leave the image on every side, radii
    from sub-pixel to half the image, every level-band form the searches use (forward, backward, +-1, [l - 1, l], none), keypoints
    outside the undistorted bounds, clustered keypoints (long cell lists)."""
    import ctypes as C
    from oracle_lib import KP_DTYPE
    Lo, Lr = O.lib(), R.lib()
    sig = [C.c_void_p, C.c_int] + [C.c_float] * 6 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    for fn in (Lo.orc_features_in_area, Lr.ref_features_in_area):
        fn.argtypes = sig; fn.restype = C.c_int
    r = np.random.default_rng(23)
    n_total = 0
    for case in range(12):
        w, h = (752, 480) if case % 2 == 0 else (512, 512)
        minx, maxx, miny, maxy = (np.float32(v) for v in ((-3.7, w + 4.2, -2.1, h + 1.6) if case % 3 else (0, w, 0, h)))
        winv = np.float32(64) / np.float32(maxx - minx); hinv = np.float32(48) / np.float32(maxy - miny)
        n_kp = int(r.integers(50, 2000))
        kps = np.zeros(n_kp, KP_DTYPE)
        if case % 4 == 3:   # clustered
            c = r.uniform([50, 50], [w - 50, h - 50], (8, 2))
            p = c[r.integers(0, 8, n_kp)] + r.normal(0, 6, (n_kp, 2))
        else:
            p = r.uniform([minx - 5, miny - 5], [maxx + 5, maxy + 5], (n_kp, 2))
        kps["x"], kps["y"] = p[:, 0].astype(np.float32), p[:, 1].astype(np.float32)
        kps["octave"] = r.integers(0, 8, n_kp)
        n_q = 600
        q = np.zeros((n_q, 3), np.float32)
        q[:, 0] = r.uniform(minx - 40, maxx + 40, n_q); q[:, 1] = r.uniform(miny - 40, maxy + 40, n_q)
        q[:, 2] = r.choice([0.4, 3.0, 7.0, 15.0, 40.0, 250.0], n_q) * r.uniform(0.5, 1.5, n_q)
        q[:50, :2] = p[r.integers(0, n_kp, 50)].astype(np.float32)      # windows centred on keypoints
        lv = r.integers(0, 8, n_q)
        form = r.integers(0, 5, n_q)
        ql = np.stack([np.where(form == 0, 0, np.where(form == 1, lv, np.where(form == 2, lv - 1, np.where(form == 3, lv - 1, -1)))),
                       np.where(form == 0, lv, np.where(form == 1, -1, np.where(form == 2, lv + 1, np.where(form == 3, lv, -1))))], 1).astype(np.int32)
        outs = []
        for fn in (Lo.orc_features_in_area, Lr.ref_features_in_area):
            ptr = np.zeros(n_q + 1, np.int32); idx = np.zeros(400000, np.int32); inim = np.zeros(n_q, np.uint8)
            tot = fn(kps.ctypes.data, n_kp, float(minx), float(maxx), float(miny), float(maxy), float(winv), float(hinv), q.ctypes.data,
                     ql.ctypes.data, n_q, ptr.ctypes.data, idx.ctypes.data, len(idx), inim.ctypes.data)
            assert tot <= len(idx)
            outs.append((tot, ptr.copy(), idx[:tot].copy(), inim.copy()))
        assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2]), case
        assert np.array_equal(outs[0][3], outs[1][3])
        n_total += outs[0][0]
    assert n_total > 50000


@pytest.mark.parametrize("kw", [dict(seed=5, th=15.0), dict(seed=6, th=7.0, mono=True), dict(seed=7, th=15.0, motion="forward"),
                                dict(seed=8, th=15.0, motion="backward"), dict(seed=9, th=30.0, cluster=True, blocked_frac=0.3),
                                dict(seed=10, th=15.0, th_far=6.0)])
def test_search_by_projection_last_frame_equal_reference(kw):
    """ORBmatcher::SearchByProjection(Frame&, const Frame& last, th, bMono, th_far_pts) (src/ORBmatcher.cc:1303-1467) compiled
    unchanged over the reference's own grid functions: the oracle assigns the same map point to the same keypoint and returns the
    same match count — level bands for forward / backward / uncertain motion, monocular frames, the th_far gate, blocked keypoints,
    clustered keypoints with many candidates per window, the rotation histogram."""
    from vieo_slam_b200 import synth
    import inspect
    args = {k: v for k, v in kw.items() if k in inspect.signature(synth.make_sbp_problem).parameters}
    pb = synth.make_sbp_problem(kw["seed"], 3, mode=synth.SBP_LAST_FRAME, **{k: v for k, v in args.items() if k != "seed"})
    pb["q_Xw"] = np.ascontiguousarray(pb["q_Xw"].astype(np.float32).astype(np.float64))   # MapPoint positions are float in the reference
    kp_o, q_o, d_o, n_o = O.search_by_projection(pb)
    kp_r, n_r = R.search_by_projection_last_frame(pb)
    assert np.array_equal(n_o, n_r), (n_o, n_r)
    assert np.array_equal(kp_o, kp_r)
    assert n_o.sum() > 100


@pytest.mark.parametrize("kw", [dict(seed=15, th=1.0), dict(seed=16, th=3.0), dict(seed=17, th=1.0, cluster=True, blocked_frac=0.3),
                                dict(seed=18, th=5.0, th_far=6.0), dict(seed=19, th=1.0, mono=True)])
def test_search_by_projection_local_map_equal_reference(kw):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, th_far_pts) + RadiusByViewingCos (src/ORBmatcher.cc:230-342)
    compiled unchanged over the reference's own grid functions: same keypoint -> map point assignment and match count as the oracle
    — the 2.5 / 4.0 window by viewing cosine, the [l - 1, l] level band, the ur gate, best / second-best with the same-level ratio
    test, blocked keypoints, the th_far gate on track_depth_."""
    from vieo_slam_b200 import synth
    pb = synth.make_sbp_problem(kw["seed"], 3, mode=synth.SBP_LOCAL_MAP, n_q=1800, **{k: v for k, v in kw.items() if k != "seed"})
    for ratio in (0.8, 0.6):
        pb["frames"]["nn_ratio"] = ratio
        kp_o, q_o, d_o, n_o = O.search_by_projection(pb)
        kp_r, n_r = R.search_by_projection_local_map(pb)
        assert np.array_equal(n_o, n_r), (n_o, n_r)
        assert np.array_equal(kp_o, kp_r)
        assert n_o.sum() > 100


# ---- the inertial arithmetic, compiled unchanged against the Eigen / Sophus stand-in (oracle/ref_build/eigstub) ------------------------
# These comparisons carry a relative tolerance instead of bit equality: the stand-in evaluates Eigen's expressions eagerly, and the
# order in which a dot product's terms are added is Eigen's own business (it depends on its version, vectorisation and -march).
def _close(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= tol * (1.0 + np.maximum(np.abs(a), np.abs(b)))))


def test_so3_helpers_equal_reference():
    """common/so3_extra.h compiled whole: SO3ex::exp / Exp / log / Log / JacobianR / JacobianRInv / normalizeRotationM against
    so3_oracle.h over the small-angle branch (|w| < 1e-5), ordinary angles and angles next to pi; Sophus' own log() (what a product of
    two SO3ex yields in the USE_SOPHUS_NEWEST build, used by EdgeNavStateI's rotation residual) equals SO3ex::log away from pi."""
    r = np.random.default_rng(1)
    for s in [0.0, 1e-12, 1e-8, 9.9e-6, 1.01e-5, 1e-4, 1e-2, 0.3, 1.0, 2.0, 3.0, np.pi - 1e-3, np.pi - 1e-7]:
        for _ in range(20):
            w = r.normal(0, 1, 3); w *= s / np.linalg.norm(w)
            for op in (0, 1, 4, 5):
                assert _close(O.so3(op, w), R.so3(op, w), 1e-14), (op, s, w)
            q = R.so3(0, w)
            assert _close(O.so3(2, q), R.so3(2, q), 1e-14) and _close(O.so3(2, -q), R.so3(2, -q), 1e-14)
            Rm = R.so3(1, w)
            assert _close(O.so3(3, Rm), R.so3(3, Rm), 1e-13), (s, w)
            noisy = Rm + r.normal(0, 1e-9, (3, 3))
            assert _close(O.so3(6, noisy), R.so3(6, noisy), 1e-14)
            if s < 3.1:
                assert _close(R.so3(7, q), R.so3(2, q), 1e-12), (s, w)
            # round trip and the Jacobians' defining property Jr * Jr^-1 = I
            if 0 < s < 3.1:
                assert np.allclose(R.so3(2, q), w, rtol=1e-9, atol=1e-15)
                assert np.allclose(R.so3(4, w) @ R.so3(5, w), np.eye(3), atol=1e-9)


def test_imu_update_recurrence_equal_reference():
    """IMUPreIntegratorBase::update (src/Odom/OdomPreIntegrator.h:431-506) compiled unchanged: delta R / v / p, the five bias Jacobians
    and both 9x9 covariances (PRV and PVR order) after feeding it the update() calls the sample selection issues — which the recording
    test above shows to be the reference's — equal the oracle's whole pre-integration; all three noise forms (discrete sigma, sigma / dt,
    sigma * freq_ref for long steps)."""
    s = synth_mod().vio_sequence(21, 40)
    imu, t = s["imu"], s["times"]
    n_checked = 0
    for fixed, freq in ((1, 200.0), (0, 0.0), (0, 200.0), (0, 1000.0)):
        nz = O.imu_noise(dt_cov_noise_fixed=fixed, freq_ref=freq)
        for k in range(1, 40, 3):
            for (ti, tj) in ((t[k - 1], t[k]), (t[k], t[k - 1]), (t[max(k - 3, 0)], t[k])):
                lo = max(np.searchsorted(imu[:, 0], min(ti, tj), "right") - 1, 0)
                hi = min(np.searchsorted(imu[:, 0], max(ti, tj), "left") + 1, len(imu))
                bg, ba = s["truth"][k]["bg"], s["truth"][k]["ba"]
                want = O.imu_preintegrate(imu[lo:hi], ti, tj, bg, ba, nz)
                rc, tr = O.imu_preintegrate_trace(imu[lo:hi], ti, tj, bg, ba)
                assert rc == 0 and len(tr) > 0
                got = R.imu_update_sequence(nz, tr, O.PREINT_DTYPE)
                for f in ("Rij", "vij", "pij", "Jgp", "Jap", "Jgv", "Jav", "JgR", "dt"):
                    assert _close(want[f], got[f], 1e-12), (f, fixed, freq, k)
                for f in ("SigmaPRV", "SigmaPVR"):
                    sc = np.abs(want[f]).max()
                    assert sc > 0 and np.abs(want[f] - got[f]).max() <= 1e-11 * sc, (f, fixed, freq, k)
                n_checked += 1
    assert n_checked == 4 * 13 * 3


def synth_mod():
    from vieo_slam_b200 import synth
    return synth


@pytest.fixture(scope="module")
def inertial_seq():
    s = synth_mod().vio_sequence(5, 40)
    s["pre"] = O.imu_preintegrate_frames(s, list(range(40)), O.imu_noise())
    return s


def test_inertial_edges_equal_reference(inertial_seq):
    """EdgeNavStateI<3> (PVR), <5> (PRV) and <6> (PRVG, with VertexGThetaXYRwI) of src/Odom/g2otypes.h:725-884 compiled unchanged — class
    definitions, computeError and linearizeOplus — over the reference's own NavState / SO3ex: residuals and every Jacobian block equal
    the oracle's, incl. large bias corrections and rotation residuals of tens of degrees."""
    synth = synth_mod()
    seq = inertial_seq
    r = np.random.default_rng(4)
    G = 9.81
    for trial in range(60):
        j = int(r.integers(1, 40)); i = j - 1
        big = trial % 3 == 2
        nsi = synth.perturb_state(seq["truth"][i], r, drot=np.deg2rad(20 if big else 0.3), dp=0.3 if big else 0.01)
        nsj = synth.perturb_state(seq["truth"][j], r, drot=np.deg2rad(20 if big else 0.3))
        nsi["dbg"] = r.normal(0, 2e-2 if big else 1e-3, 3); nsi["dba"] = r.normal(0, 1e-1 if big else 1e-2, 3)
        pre = seq["pre"][j]
        gw = synth.GRAVITY_W if trial % 2 else r.normal(0, 1, 3) * 5
        for order in (0, 1):
            a = O.edge_navstate(nsi, nsj, pre, gw, order)
            b = R.edge_navstate(nsi, nsj, pre, gw, order)
            for x, y, name in zip(a, b, ("e", "Ji", "Jj", "Jb")):
                assert _close(x, y, 1e-11), (trial, order, name, np.abs(x - y).max())
        gdir = gw / np.linalg.norm(gw) * G
        qo, qr = O.gdir_init(gdir), R.gdir_init(gdir)
        assert _close(qo, qr, 1e-14)
        d2 = r.normal(0, 0.05, 2)
        assert _close(O.gdir_oplus(qo, d2), R.gdir_oplus(qo, d2), 1e-14)
        GI = np.array([0.0, 0.0, G])
        a = O.edge_navstate_g(nsi, nsj, pre, qo, GI)
        b = R.edge_navstate(nsi, nsj, pre, GI, 1, q_wI=qo)
        for x, y, name in zip(a, b, ("e", "Ji", "Jj", "Jb", "JG")):
            assert _close(x, y, 1e-11), (trial, name, np.abs(x - y).max())
    # gravity (anti-)parallel to the z axis: normalized() of the zero cross product, RwI = I on both sides
    for gz in ([0, 0, G], [0, 0, -G]):
        assert _close(O.gdir_init(np.array(gz, float)), R.gdir_init(np.array(gz, float)), 1e-15)


def test_vertex_updates_and_prior_edges_equal_reference(inertial_seq):
    """NavState::IncSmall / IncSmallBias through VertexNavState<6 | 9 | 3>::oplusImpl and VertexNavStateBias::oplusImpl (src/Odom/NavState.h
    whole, g2otypes.h:270-285, 553-567), EdgeNavStatePriorPVRBias (g2otypes.cpp:84-124) and EdgeNavStateBias (:14-35), compiled unchanged."""
    synth = synth_mod()
    seq = inertial_seq
    r = np.random.default_rng(8)
    for trial in range(80):
        ns = synth.perturb_state(seq["truth"][int(r.integers(0, 40))], r, drot=np.deg2rad(30))
        ns["dbg"] = r.normal(0, 1e-3, 3); ns["dba"] = r.normal(0, 1e-2, 3)
        scale = [1e-9, 1e-3, 0.3][trial % 3]
        for kind, n in ((0, 6), (1, 9), (2, 3), (3, 6)):
            dx = r.normal(0, scale, n)
            a, b = O.navstate_oplus(ns, kind, dx), R.navstate_oplus(ns, kind, dx)
            for f in ("p", "q", "v", "bg", "ba", "dbg", "dba"):
                assert _close(a[f], b[f], 1e-14), (trial, kind, f)
        prior = synth.perturb_state(ns, r, drot=np.deg2rad([0.3, 15][trial % 2]), dp=0.2)
        prior["dbg"] = r.normal(0, 1e-3, 3); prior["dba"] = r.normal(0, 1e-2, 3)
        eo, Jo = O.edge_prior_pvr(ns, prior)
        er, Jr, Jbr = R.edge_prior(0, ns, prior)
        assert _close(eo, er, 1e-12) and _close(Jo, Jr, 1e-12), trial
        assert np.array_equal(Jbr[9:], np.eye(6)) and not Jbr[:9].any()
        # the PRV form holds the same quantities in P R V order (the LocalBA / GlobalBA engines use the PVR form's values re-ordered)
        e1, J1, Jb1 = R.edge_prior(1, ns, prior)
        perm = [0, 1, 2, 6, 7, 8, 3, 4, 5]
        assert _close(e1[:9], er[perm], 1e-12) and _close(e1[9:], er[9:], 1e-15)
        assert _close(J1[:9][:, :9], Jr[perm][:, perm], 1e-12)
        eb, Ji, Jj = R.edge_bias(ns, prior)
        want = np.r_[(prior["bg"] + prior["dbg"]) - (ns["bg"] + ns["dbg"]), (prior["ba"] + prior["dba"]) - (ns["ba"] + ns["dba"])]
        assert np.array_equal(eb, want) and np.array_equal(Ji, -np.eye(6)) and np.array_equal(Jj, np.eye(6))


def test_initial_gyro_bias_edge_equal_reference():
    """EdgeGyrBias (src/Odom/g2otypes.h:940-973, the edge of Optimizer::OptimizeInitialGyroBias) compiled unchanged: one Gauss-Newton
    step assembled from the REFERENCE's residuals and Jacobians (H = sum J^T W J, b = -sum J^T W e at bg = 0, W the inverse rotation
    block of the PRV covariance) gives the oracle's estimate; the edge also agrees at a non-zero bias."""
    synth = synth_mod()
    g = synth.make_gyro_bias_problem(31, n_kf=16, kf_gap=(1, 12))
    pre = O.imu_preintegrate_frames(g["seq"], g["kf_idx"], O.imu_noise())
    rng = np.random.default_rng(3)
    Rwb = np.stack([Rm @ synth.so3_exp(rng.normal(0, 2e-3, 3)) for Rm in g["Rwb"]])
    for use_info in (True, False):
        H = np.zeros((3, 3)); b = np.zeros(3)
        for i in range(1, len(pre)):
            e, J = R.edge_gyr_bias(pre[i]["Rij"], pre[i]["JgR"], Rwb[i - 1], Rwb[i], np.zeros(3))
            W = np.linalg.inv(pre[i]["SigmaPRV"][3:6, 3:6]) if use_info else np.eye(3)
            H += J.T @ W @ J; b -= J.T @ W @ e
        n, dbg = O.gyro_bias_init(pre, Rwb, use_info)
        want = np.linalg.solve(H, b)
        assert n == 15 and np.abs(dbg - want).max() <= 1e-9 * np.abs(want).max(), (dbg, want)
    # the residual at a non-zero bias is the rotation part of the PRV inertial edge with that bias correction
    i = 4
    bg = np.array([2e-3, -1e-3, 3e-3])
    e, J = R.edge_gyr_bias(pre[i]["Rij"], pre[i]["JgR"], Rwb[i - 1], Rwb[i], bg)
    nsi = np.zeros(1, O.NAVSTATE_DTYPE)[0]; nsj = nsi.copy()
    nsi["q"] = synth.quat_from_R(Rwb[i - 1]); nsj["q"] = synth.quat_from_R(Rwb[i]); nsi["dbg"] = bg
    e9, Ji, Jj, Jb = O.edge_navstate(nsi, nsj, pre[i], np.zeros(3), 1)
    assert np.allclose(e, e9[3:6], rtol=0, atol=1e-12) and np.allclose(J, Jb[3:6, :3], rtol=0, atol=1e-10)


@pytest.mark.parametrize("model", ["pinhole", "radtan", "kb8"])
def test_visual_edges_equal_reference(model, inertial_seq):
    """EdgeReproject<DE, DV, NV, MODE_OPT_VAR> (src/Odom/g2otypes.h:321-541) compiled unchanged — GetTcw_wX, cam_project, computeError,
    GetDepth, linearizeOplus — over the reference's own compiled Project() of the three camera models: the monocular / stereo PR edge of
    PoseOptimization / LocalBA / GlobalBA, the scale-vertex forms PRS / PRSStereo of the final GlobalBA and the PRS / PRSInv pair of
    OptimizeSim3 equal the oracle's residuals, Jacobian blocks and depths."""
    synth = synth_mod()
    cam = {"pinhole": synth.euroc_camera, "radtan": synth.radtan_camera, "kb8": synth.kb8_camera}[model]()
    r = np.random.default_rng(17)
    n = 0
    for trial in range(60):
        ns = synth.perturb_state(inertial_seq["truth"][int(r.integers(0, 40))], r, drot=np.deg2rad(10), dp=0.2)
        Rwb = synth.R_from_quat(ns["q"])
        Rcb = np.asarray(cam["Rcb"]).reshape(3, 3); tcb = np.asarray(cam["tcb"])
        Pc = np.array([r.uniform(-1.5, 1.5), r.uniform(-1.0, 1.0), 1.0]) * r.uniform(0.8, 25.0)
        Xw = Rwb @ (Rcb.T @ (Pc - tcb)) + ns["p"]
        obs = np.array([r.uniform(0, 752), r.uniform(0, 480), r.uniform(0, 752)], np.float32)
        for stereo in (0, 1):
            eo, Jpo, JXo, do = O.edge_reproject(cam, ns, Xw, obs, stereo)
            rows = 3 if stereo else 2
            er, Jpr, JXr, _, dr = R.edge_reproject(stereo, cam, ns, Xw, obs)
            assert _close(eo[:rows], er, 1e-12) and _close(Jpo[:rows], Jpr, 1e-11) and _close(JXo[:rows], JXr, 1e-11), (trial, stereo)
            assert _close(do, dr, 1e-13)
            # the PVR forms hold the same blocks with zero velocity columns in between
            e9, J9, JX9, _, _ = R.edge_reproject(2 + stereo, cam, ns, Xw, obs)
            assert np.array_equal(e9, er) and np.array_equal(J9[:, :3], Jpr[:, :3]) and np.array_equal(J9[:, 6:], Jpr[:, 3:])
            assert not J9[:, 3:6].any() and np.array_equal(JX9, JXr)
            # scale vertex (final GlobalBA): X = s * Xh
            s = float(r.uniform(0.7, 1.4))
            eo, Jpo, JXo, Jso = O.edge_reproject_scale(cam, ns, Xw / s, s, obs, stereo)
            er, Jpr, JXr, Jsr, _ = R.edge_reproject(4 + stereo, cam, ns, Xw / s, obs, scale=s)
            assert _close(eo[:rows], er, 1e-11) and _close(Jpo[:rows], Jpr, 1e-10) and _close(JXo[:rows], JXr, 1e-10), (trial, stereo)
            assert _close(Jso[:rows], Jsr, 1e-10)
            n += 1
        # OptimizeSim3's pair: S12 as a NavState, points in the other camera's frame, Rcb = I, tcb = 0
        cam_i = cam.copy(); cam_i["Rcb"] = np.eye(3).reshape(cam["Rcb"].shape); cam_i["tcb"] = 0
        s12 = np.zeros(1, O.NAVSTATE_DTYPE)[0]
        s12["q"] = synth.quat_from_R(synth.so3_exp(r.normal(0, 0.2, 3))); s12["p"] = r.normal(0, 0.3, 3)
        s = float(r.uniform(0.8, 1.25))
        for inverse in (0, 1):
            eo, Jpo, Jso = O.edge_sim3(cam_i, s12, s, Pc, obs[:2], inverse)
            er, Jpr, _, Jsr, _ = R.edge_reproject(6 if inverse else 4, cam_i, s12, Pc, obs[:2], scale=s)
            assert _close(eo, er, 1e-11) and _close(Jpo, Jpr, 1e-10) and _close(Jso, Jsr, 1e-10), (trial, inverse)
    assert n == 120


def test_sim3_equal_reference():
    """optimizer/g2o/g2o/types/sim3.h compiled whole against the Eigen stand-in: the Vector7d constructor (exp), log(), inverse(),
    operator* and VertexSim3Expmap::oplusImpl equal the oracle's restatement in every branch (|sigma| and theta below / above 1e-5,
    rotations up to ~pi)."""
    r = np.random.default_rng(23)
    n = 0
    for th in (0.0, 1e-7, 9e-6, 1.1e-5, 1e-3, 0.5, 2.0, 3.0):
        for sg in (0.0, 1e-7, 9e-6, 1.1e-5, 1e-2, 0.4, -0.7):
            for _ in range(6):
                w = r.normal(0, 1, 3); w *= th / np.linalg.norm(w)
                u = np.r_[w, r.normal(0, 2, 3), sg]
                So, Sr = O.sim3_exp(u), R.sim3(0, u=u)
                for f in ("q", "t", "s"):
                    assert _close(So[f], Sr[f], 1e-13), (th, sg, f, So[f], Sr[f])
                lo, lr = O.sim3_log(So), R.sim3(1, a=So)
                assert _close(lo, lr, 1e-9 if th < 1e-4 else 1e-11), (th, sg, lo, lr)
                io, ir = O.sim3_inv(So), R.sim3(3, a=So)
                T = O.sim3_exp(np.r_[r.normal(0, 0.3, 3), r.normal(0, 1, 3), r.normal(0, 0.2)])
                mo, mr = O.sim3_mul(So, T), R.sim3(2, a=So, b=T)
                for f in ("q", "t", "s"):
                    assert _close(io[f], ir[f], 1e-13) and _close(mo[f], mr[f], 1e-13)
                for fix in (False, True):                      # oplus: S <- exp(update) * S, sigma forced to 0 when the scale is fixed
                    upd = np.r_[r.normal(0, 1e-2, 6), 0.03]
                    want = O.sim3_mul(O.sim3_exp(np.r_[upd[:6], 0.0 if fix else upd[6]]), So)
                    got = R.sim3(4, a=So, b=fix, u=upd)
                    for f in ("q", "t", "s"):
                        assert _close(want[f], got[f], 1e-13)
                n += 1
    assert n == 8 * 7 * 6


def test_essential_graph_edge_equal_reference():
    """EdgeSim3::computeError (types_seven_dof_expmap.h) and g2o's numeric-Jacobian BaseBinaryEdge::linearizeOplus
    (core/base_binary_edge.hpp:131-203: delta 1e-9, central differences through push / oplus / pop, fixed vertices skipped) compiled
    unchanged: residual to 1e-11; the Jacobians are differences of 1e-9 steps, so they agree to the noise of that quotient."""
    r = np.random.default_rng(29)
    for trial in range(40):
        def rnd(scale_sigma=0.1):
            return O.sim3_exp(np.r_[r.normal(0, 0.5, 3), r.normal(0, 2, 3), r.normal(0, scale_sigma)])
        v0, v1 = rnd(), rnd()
        meas = O.sim3_mul(O.sim3_mul(v1, O.sim3_inv(v0)), O.sim3_exp(np.r_[r.normal(0, 0.02, 3), r.normal(0, 0.05, 3), r.normal(0, 0.01)]))
        fix0, fix1, fs = trial % 5 == 0, trial % 7 == 0, trial % 2 == 0
        eo, Jio, Jjo = O.edge_sim3_graph(meas, v0, v1, fix0, fix1, fs)
        er, Jir, Jjr = R.edge_sim3_graph(meas, v0, v1, fix0, fix1, fs)
        assert _close(eo, er, 1e-11), (trial, eo, er)
        if fix0 and fix1:
            continue
        sc = max(np.abs(Jir).max(), np.abs(Jjr).max(), 1.0)
        assert np.abs(Jio - Jir).max() <= 2e-6 * sc and np.abs(Jjo - Jjr).max() <= 2e-6 * sc, (trial, np.abs(Jio - Jir).max())
        assert (not Jir.any()) == bool(fix0) and (not Jjr.any()) == bool(fix1)
        if fs:
            assert not Jir[:, 6].any() and not Jjr[:, 6].any()


@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("n_cams", [1, 2, 4])
def test_is_in_frustum_equal_reference(model, n_cams):
    """Frame::isInFrustum (src/Frame.cc:335-416) compiled unchanged — per camera of the rig: GetTcr() * Pcr, the depth sign test, K *
    p_normalize in float or the camera's own Project() (usedistort_), the per-camera image bounds, the scale-invariance distance band
    (MapPoint::GetMin / MaxDistanceInvariance cut out with it), the viewing cosine, PredictScale, the tracking-info lists and the mean
    depth — against the oracle: every output bit for bit (float reductions in the order of Eigen's unrolled redux, which the stand-in
    restates: t0 + (t1 + t2))."""
    synth = synth_mod()
    pb = synth.make_frustum_rig_problem(40 + model, n_frames=3, n_q=1500, n_cams=n_cams, model=model)
    # some points exactly on the decision boundaries: behind / on the camera plane, at the distance band's ends
    G = pb["rig"][0]
    Rcw = np.asarray(G["Rcw"], np.float32).reshape(3, 3); Ow = np.asarray(G["Ow"], np.float32)
    for k in range(8):
        pb["p_wP"][k] = Ow + Rcw.T @ np.array([0.1 * k, -0.05 * k, [0.0, -1.0, 1e-6, 2.0][k % 4]], np.float32)
    d = np.linalg.norm(pb["p_wP"][8:16] - Ow, axis=1).astype(np.float32)
    pb["p_max_dist"][8:12] = d[:4] / np.float32(1.2); pb["p_min_dist"][12:16] = d[4:] / np.float32(0.8)
    a, b = O.is_in_frustum_rig(pb), R.is_in_frustum_rig(pb)
    assert a["n_inview"].sum() > 1500 and np.array_equal(a["n_inview"], b["n_inview"])
    for k in ("inview", "cam_mask", "level", "proj", "viewcos", "depth"):
        assert a[k].tobytes() == b[k].tobytes(), (k, int((a[k] != b[k]).sum()))
    if n_cams > 1:
        assert len(np.unique(a["cam_mask"])) > 3      # points seen by different camera subsets


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_search_by_bow_equal_reference(seed):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:344-505) compiled unchanged over a std::map
    FeatureVector (what DBoW2's is): the node merge with lower_bound jumps, best / second-best per keyframe map point over the frame
    keypoints of the node not yet taken, TH_LOW and the ratio test, the rotation histogram with its three maxima — same assignment of
    keyframe map points to frame keypoints and the same count as the oracle, with and without the orientation check."""
    synth = synth_mod()
    for ratio in (0.7, 0.9):
        pb = synth.make_bow_problem(seed, n_pairs=4, n_kp=1000, n_nodes=70, nn_ratio=ratio)
        tot = 0
        for p in range(4):
            for chk in (0, 1):
                pb["pairs"]["check_orientation"][p] = chk
                mo, no = O.search_by_bow(pb, p)
                mr, nr = R.search_by_bow(pb, p)
                assert no == nr and np.array_equal(mo, mr), (seed, p, chk, no, nr)
                assert no == int((mo >= 0).sum())
                tot += no
        assert tot > 300


def test_search_by_bow_shared_map_points_equal_reference():
    """The same search when several keyframe keypoints hold the SAME map point (rigs: one point seen by more than one camera): the
    second, closer match replaces the first one, frees its frame keypoint and removes its rotation-histogram entry
    (src/ORBmatcher.cc:424-441, 485-492)."""
    synth = synth_mod()
    r = np.random.default_rng(31)
    replaced = 0
    for seed in (6, 7, 8):
        pb = synth.make_bow_problem(seed, n_pairs=3, n_kp=900, n_nodes=40, nn_ratio=0.9)
        for p in range(3):
            n1 = int(pb["pairs"]["n_kp1"][p])
            k1 = slice(int(pb["pairs"]["kp1_begin"][p]), int(pb["pairs"]["kp1_begin"][p]) + n1)
            mp = np.where(pb["mp_ok"][k1] != 0, r.integers(0, n1 // 3, n1), -1).astype(np.int32)   # ~3 keypoints per map point
            for chk in (0, 1):
                pb["pairs"]["check_orientation"][p] = chk
                mo, no = O.search_by_bow(pb, p, mp_id=mp)
                mr, nr = R.search_by_bow(pb, p, mp_id=mp)
                ids_o = np.where(mo >= 0, mp[np.maximum(mo, 0)], -1)      # the oracle reports the keyframe keypoint, the reference its map point
                assert no == nr and np.array_equal(ids_o, mr), (seed, p, chk, no, nr)
                m1, n1_ = O.search_by_bow(pb, p)
                replaced += int(no < n1_)
    assert replaced > 6


@pytest.mark.parametrize("kw", [dict(seed=51), dict(seed=52, th=15.0, orb_dist=64), dict(seed=53, th_far=8.0, cluster=True),
                                dict(seed=54, blocked_frac=0.4, orb_dist=80)])
def test_search_by_projection_reloc_equal_reference(kw):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist, th_far_pts) (src/ORBmatcher.cc:1471-1606,
    Tracking::Relocalization's guided search) compiled unchanged over the compiled grid functions and PredictScale: the th_far gate,
    the projection, the 0.8 / 1.2 scale-invariance gate on the camera-centre distance, the [l - 1, l + 1] band, claimed keypoints,
    the caller's ORBdist, the rotation histogram — same keypoint assignments and counts as the oracle."""
    synth = synth_mod()
    pb = synth.make_reloc_problem(kw["seed"], n_frames=3, **{k: v for k, v in kw.items() if k != "seed"})
    for chk in (0, 1):
        pb["frames"]["check_orientation"] = chk
        kp_o, q_o, d_o, l_o, n_o = O.sbp_reloc(pb)
        kp_r, n_r = R.sbp_reloc(pb)
        assert np.array_equal(n_o, n_r), (n_o, n_r)
        assert np.array_equal(kp_o, kp_r)
        assert n_o.sum() > 150


@pytest.mark.parametrize("kw", [dict(), dict(use_bf=False), dict(check_viewing_angle=False, th_radius=4.0), dict(cluster=True, skip_frac=0.3)])
def test_search_by_projection_base_equal_reference(kw):
    """ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227), the search half behind Fuse and SearchBySim3, compiled unchanged
    over the compiled grid functions and PredictScale: projection, image / scale-invariance / viewing-cone tests, the window
    th_radius * scale[level], the [l - 1, l] band, the stereo 7.8 / mono 5.99 chi-square gate, the strict-'<' Hamming arg-min — the
    keypoint found for every map point (read from pvnMatch1 in the FuseLater mode with no distance threshold) and its distance equal
    the oracle's."""
    synth = synth_mod()
    pb = synth.make_fuse_problem(61, **kw)
    bo, do, lo = O.proj_search(pb)
    br, dr = R.proj_search(pb)
    assert (bo >= 0).sum() > 1500
    assert np.array_equal(bo, br)
    assert np.array_equal(do[bo >= 0], dr[bo >= 0])


def test_distinctive_descriptors_equal_reference():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:314-378) compiled unchanged: all-pairs Hamming distances, the sorted
    row's entry at int(0.5 (N - 1)), the first row with the least median — the same row as the oracle for 1 ... 40 observations,
    near-duplicate descriptors (many equal medians) and points without observations."""
    r = np.random.default_rng(37)
    sizes = np.r_[0, 1, 2, 3, r.integers(1, 41, 300), 0]
    ptr = np.r_[0, np.cumsum(sizes)].astype(np.int32)
    pool = r.integers(0, 256, (int(ptr[-1]), 32), dtype=np.uint8)
    for p in range(0, len(sizes), 3):                       # clusters of near-identical rows: ties between medians
        b, e = ptr[p], ptr[p + 1]
        if e - b > 2:
            pool[b:e] = pool[b]
            flips = r.integers(0, 256, e - b)
            for i in range(b + 1, e):
                pool[i, flips[i - b] // 8] ^= 1 << (flips[i - b] % 8)
    bo, mo = O.distinctive_descriptors(pool, ptr)
    br = R.distinctive_descriptors(pool, ptr)
    assert np.array_equal(bo, br)
    assert (bo == -1).sum() == 2 and (bo > 0).sum() > 100
    rows = r.permutation(int(ptr[-1])).astype(np.int32)     # through a row list
    bo, _ = O.distinctive_descriptors(pool, ptr, rows)
    assert np.array_equal(bo, R.distinctive_descriptors(pool, ptr, rows))


def test_chi2_large_set_level_equal_reference():
    """g2o::GraphOperator (optimizer/optimizer_ba/g2o_graph_operator.h): the 5 % chi-square table and Chi2LargeSetLevel's decision
    `chi2 > rat_th_chi2 * chi2_sig5_[dim]` — a float product against the edge's double chi2 — compiled unchanged against the oracle's,
    on the float neighbourhood of every threshold the local BA uses (rat 100, dims 2 and 3) and random ones."""
    import ctypes as C
    Lo, Lr = O.lib(), R.lib()
    for L, f in ((Lo, "orc_chi2_large_level"), (Lr, "ref_chi2_large_level")):
        getattr(L, f).argtypes = [C.c_double, C.c_int, C.c_float]; getattr(L, f).restype = C.c_int
    table = np.array([0, 3.841, 5.991, 7.815, 9.488, 11.070, 12.592, 14.067, 15.507, 16.919, 18.307, 19.675, 21.026, 22.362, 23.685, 24.996], np.float32)
    r = np.random.default_rng(41)
    n1 = 0
    for dim in range(16):
        for rat in [np.float32(100.0), np.float32(1.0)] + list(r.uniform(0.5, 200, 4).astype(np.float32)):
            th = np.float32(rat * table[dim])
            for c in [float(th), float(np.nextafter(th, np.float32(np.inf))), float(np.nextafter(th, np.float32(-np.inf))),
                      float(np.nextafter(float(th), np.inf)), float(np.nextafter(float(th), -np.inf)), float(rat) * float(table[dim])] + \
                     list(r.uniform(0, 2.5 * max(float(th), 1.0), 10)):
                a, b = Lo.orc_chi2_large_level(c, dim, rat), Lr.ref_chi2_large_level(c, dim, rat)
                assert a == b, (dim, rat, c)
                n1 += a
    assert n1 > 300


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_search_for_triangulation_equal_reference(seed):
    """ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:896-1150) compiled unchanged, with GeometricCamera::epipolarConstrain
    (camera_base.h:287-406, the fundamental-matrix branch) and FillMatchesFromPair (:408-585, USE_STRATEGY_MIN_DIST) cut out of the
    camera header: the FeatureVector walk, the map-point / only-stereo / injection skips, the epipole gate of monocular pairs, the
    float / double mix of the epipolar distance, the match bookkeeping, the rotation histogram — the same (idx1, idx2) list in the
    same order and the same count as the oracle.  The reference starts from keyframe POSES (T12 rounded to float), the oracle and
    the C ABI from the caller's F12 / epipole: both are formed from the poses by ref_sft_geometry with the stand-in's operations."""
    synth = synth_mod()
    from vieo_slam_b200.synth import EUROC
    K4 = np.array([EUROC[k] for k in ("fx", "fy", "cx", "cy")], np.float32)
    pb = synth.make_sft_problem(seed, n_pairs=4, n_kp=900, n_nodes=60, share_kf1=(seed % 2 == 0))
    tot = rejected = 0
    for p in range(4):
        q1, t1, q2, t2 = R.sft_poses(pb, p, seed)
        ex, ey, F = R.sft_geometry(K4, q1, t1, q2, t2)
        # the float-rounded T12 gives the generator's double F12 up to scale and ~1e-6
        F0 = np.asarray(pb["pairs"][p]["F12"], np.float64)
        assert np.allclose(F / np.linalg.norm(F), F0 / np.linalg.norm(F0), atol=2e-5)
        assert abs(ex - float(pb["pairs"][p]["ex"])) < 0.05 * (1 + abs(ex)) and abs(ey - float(pb["pairs"][p]["ey"])) < 0.05 * (1 + abs(ey))
        pb["pairs"]["F12"][p] = F; pb["pairs"]["ex"][p] = ex; pb["pairs"]["ey"][p] = ey
        for only_stereo in (0, 1):
            n_chk = []
            for chk in (0, 1):
                pb["pairs"]["only_stereo"][p] = only_stereo; pb["pairs"]["check_orientation"][p] = chk
                po, no = O.search_for_triangulation(pb, p)
                pr, nr = R.search_for_triangulation(pb, p, K4, q1, t1, q2, t2)
                assert no == nr and np.array_equal(po, pr), (seed, p, only_stereo, chk, no, nr)
                assert no == len(po)
                tot += no; n_chk.append(no)
            rejected += int(n_chk[1] < n_chk[0])                                        # the histogram dropped outliers
    assert tot > 400 and rejected > 0


def test_search_for_triangulation_epipole_gate_equal_reference():
    """Monocular pairs near the epipole (src/ORBmatcher.cc:1031-1035): keyframe 2 moved straight ahead, so the epipole lies inside the
    image and the 100 * scale gate decides; no stereo coordinates at all."""
    synth = synth_mod()
    from vieo_slam_b200.synth import EUROC
    K4 = np.array([EUROC[k] for k in ("fx", "fy", "cx", "cy")], np.float32)
    pb = synth.make_sft_problem(21, n_pairs=2, n_kp=700, n_nodes=12)
    pb["uright"][:] = -1
    for p in range(2):
        pb["rel_R12"][p] = np.eye(3); pb["rel_t12"][p] = np.array([0.01, -0.02, 0.6])
        P = pb["pairs"][p]
        k2 = slice(int(P["kp2_begin"]), int(P["kp2_begin"] + P["n_kp2"]))
        q1, t1, q2, t2 = R.sft_poses(pb, p, 5)
        ex, ey, F = R.sft_geometry(K4, q1, t1, q2, t2)
        assert 0 < ex < EUROC["w"] and 0 < ey < EUROC["h"]
        r = np.random.default_rng(p)
        near = r.random(int(P["n_kp2"])) < 0.5                                      # half of keyframe 2's keypoints around the epipole
        pb["kps"]["x"][k2] = np.where(near, ex + r.normal(0, 12, near.size), pb["kps"]["x"][k2]).astype(np.float32)
        pb["kps"]["y"][k2] = np.where(near, ey + r.normal(0, 12, near.size), pb["kps"]["y"][k2]).astype(np.float32)
        pb["pairs"]["F12"][p] = F; pb["pairs"]["ex"][p] = ex; pb["pairs"]["ey"][p] = ey
        pb["pairs"]["only_stereo"][p] = 0
        for chk in (0, 1):
            pb["pairs"]["check_orientation"][p] = chk
            po, no = O.search_for_triangulation(pb, p)
            pr, nr = R.search_for_triangulation(pb, p, K4, q1, t1, q2, t2)
            assert no == nr and np.array_equal(po, pr), (p, chk, no, nr)


def _fisheye_frame(r, n_cams, cap, n_kp, n_mono, dup_frac=0.5, max_flip=40):
    """random descriptors; in-area rows of camera j > i partly copied from camera i's in-area rows with 0 ... max_flip bits flipped"""
    desc = r.integers(0, 256, (n_cams, cap, 32), dtype=np.uint8)
    for i in range(n_cams - 1):
        for j in range(i + 1, n_cams):
            ni, nj = n_kp[i] - n_mono[i], n_kp[j] - n_mono[j]
            if ni <= 0 or nj <= 0:
                continue
            m = int(dup_frac * min(ni, nj) / (n_cams - 1))
            src = n_mono[i] + r.permutation(ni)[:m]; dst = n_mono[j] + r.permutation(nj)[:m]
            desc[j, dst] = desc[i, src]
            for d in dst:
                bits = r.choice(256, int(r.integers(0, max_flip + 1)), replace=False)
                np.bitwise_xor.at(desc[j, d], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    return desc


def _expected_fisheye_calls(desc, n_kp, n_mono):
    oi, od, og = O.fisheye_matches(desc, n_kp, n_mono)
    n_cams = desc.shape[0]
    rows, pair = [], 0
    for i in range(n_cams - 1):
        for j in range(i + 1, n_cams):
            for q in np.nonzero(og[pair])[0]:
                rows.append((i, q + n_mono[i], j, oi[pair, q, 0] + n_mono[j], od[pair, q, 0]))
            pair += 1
    return np.array(rows, np.int32).reshape(-1, 5)


@pytest.mark.parametrize("n_cams", [2, 3, 4])
def test_fisheye_matches_equal_reference(n_cams):
    """Frame::ComputeStereoFishEyeMatches (src/Frame.cc:613-779) compiled unchanged (knnMatch forwarding to the cv2-pinned restatement, a
    recording FillMatchesFromPair): the (camera, keypoint, camera, keypoint, distance) tuples it hands on — pair order, in-area slices,
    the skipped pairs, the ratio test in its float / double mix, the num_mono offsets — equal the accepted matches of the oracle's
    brute-force half, in order; the concatenated frame (N, mapn2in_, mDescriptors) is camera-major."""
    r = np.random.default_rng(70 + n_cams)
    cap, tot, passes2 = 400, 0, 0
    for case in range(8):
        n_kp = r.integers(150, cap + 1, n_cams).astype(np.int32)
        n_mono = (n_kp * r.uniform(0.2, 0.7, n_cams)).astype(np.int32)
        if case == 1:
            n_mono[0] = n_kp[0]                      # no in-area row: every pair with camera 0 is skipped (:623)
        if case == 2:
            n_mono[-1] = n_kp[-1] - 1                # a train set of ONE row: knnMatch returns one neighbour, size() < 2
        if case == 3:
            n_kp[:] = r.integers(8, 14, n_cams); n_mono[:] = n_kp // 2     # fewer than 30 accepted matches: the second pass
        desc = _fisheye_frame(r, n_cams, cap, n_kp, n_mono, max_flip=40 if case != 4 else 90)
        rec, N, m2in, dall = R.fisheye_matches(desc, n_kp, n_mono, octave=r.integers(0, 8, (n_cams, cap)))
        want = _expected_fisheye_calls(desc, n_kp, n_mono)
        if len(want) < 30:                           # thresh_cosdisparity[0] != [1] and nMatches < 30: the pair loop runs twice (:648-690)
            want = np.concatenate([want, want]); passes2 += 1
        assert np.array_equal(rec, want), (n_cams, case, len(rec), len(want))
        tot += len(want)
        assert N == int(n_kp.sum())
        assert np.array_equal(m2in, np.concatenate([np.stack([np.full(n, c), np.arange(n)], 1) for c, n in enumerate(n_kp)]))
        assert np.array_equal(dall, np.concatenate([desc[c, :n] for c, n in enumerate(n_kp)]))
    assert tot > 200 and passes2 >= 1


def test_fisheye_ratio_boundaries_equal_reference():
    """The ratio test on its boundaries: d0 == 0.7 * d1 and 0.9 * d1 in exact arithmetic (d1 = 10, 20, ... -> d0 = 7, 14 / 9, 18, ...),
    d0 around thOrbDist = 75 — float distances against double constants (src/Frame.cc:659-663)."""
    cases = [(7, 10), (14, 20), (63, 90), (9, 10), (18, 20), (72, 80), (74, 83), (75, 84), (76, 85), (70, 100), (69, 100), (0, 0), (0, 1),
             (27, 30), (26, 30), (21, 30), (20, 30), (81, 90), (74, 82), (75, 83)]
    n = len(cases)
    desc = np.zeros((2, 2 * n + 2, 32), np.uint8)
    # query q of camera 0 = block q of a 256-bit code space: train rows 2q / 2q + 1 differ from it in d0 / d1 bits, others are far
    r = np.random.default_rng(5)
    base = r.integers(0, 256, (n, 32), dtype=np.uint8)
    for q, (d0, d1) in enumerate(cases):
        desc[0, q] = base[q]
        for k, d in enumerate((d0, d1)):
            row = base[q].copy()
            bits = r.choice(256, d, replace=False)
            np.bitwise_xor.at(row, bits // 8, (1 << (bits % 8)).astype(np.uint8))
            desc[1, 2 * q + k] = row
    n_kp = np.array([n, 2 * n], np.int32); n_mono = np.zeros(2, np.int32)
    oi, od, og = O.fisheye_matches(desc, n_kp, n_mono)
    assert all(od[0, q, 0] == d0 and od[0, q, 1] == d1 for q, (d0, d1) in enumerate(cases))      # the planted neighbours are the two nearest
    rec, N, _, _ = R.fisheye_matches(desc, n_kp, n_mono)
    want = _expected_fisheye_calls(desc, n_kp, n_mono)
    if len(want) < 30:
        want = np.concatenate([want, want])
    assert np.array_equal(rec, want)
    good = og[0, :n].astype(bool)
    assert good.sum() >= 6 and (~good).sum() >= 6


@pytest.mark.parametrize("chain_prior", [False, True])
def test_fill_cov_inv_equal_reference(inertial_seq, chain_prior):
    """Optimizer::FillCovInv (include/Optimizer.h:126-206) compiled unchanged, over the fork's getRho / getHessian / getHessianij /
    getHessianXi / Xj / Xij / Xji members (src/Odom/g2otypes.h:36-254, the three *EdgeEx classes cut out by name), the compiled edge
    classes and the compiled Huber kernel: the explicit J^T (rho' Omega) J assembly of PoseOptimization's marginal — which blocks of
    which edges go where for schur_bec 0 / 2 / 1, level-0 visual edges only, the robust weights — equals the three blocks the oracle
    forms before its Schur complement (row D6).  Tolerance, not bit equality: the reference adds the mono edges, then the stereo ones."""
    synth = synth_mod()
    seq = inertial_seq
    cam = synth.euroc_camera()
    n_pts = 350
    pbs, X, obs, w, fl = synth.make_pose_problems(seq, seq["pre"], cam, n_points=n_pts, seed=3, chain_prior=chain_prior)
    checked = 0
    for k in ([1, 2] if chain_prior else [0, 3, 5]):
        pb = pbs[k:k + 1]
        res, d = O.pose_optimization_marg_dump(pb, cam, X, obs, w, fl)
        assert d["filled"] == 1 and d["n_vis"] == n_pts and d["has_imu"] == 1
        assert d["fixed_last"] == (0 if pb["last_has_prior"][0] else 1)
        b, e = int(pb["edge_begin"][0]), int(pb["edge_end"][0])
        free_last = not d["fixed_last"]
        Cm, CL, CCL = R.fill_cov_inv(cam, res["cur"], res["last"], pb["preint"][0], pb["gw"][0], d["info_imu"], d["delta_imu"], d["info_bias"],
                                     d["delta_bias"], pb["prior"][0] if free_last else None, d["info_prior"], d["delta_prior"], X[b:e], obs[b:e],
                                     (fl[b:e] & 1).astype(np.uint8), w[b:e].astype(np.float64), d["level"], d["delta"])
        sc = np.abs(d["C"]).max()
        assert np.abs(Cm - d["C"]).max() <= 1e-13 * sc, (k, np.abs(Cm - d["C"]).max() / sc)
        assert (d["level"] != 0).sum() > 10 and (d["level"] == 0).sum() > 100      # outliers were left out, inliers summed
        if free_last:
            for a, bb, nm in ((CL, d["CL"], "CL"), (CCL, d["CCL"], "CCL")):
                s2 = np.abs(bb).max()
                assert s2 > 0 and np.abs(a - bb).max() <= 1e-13 * s2, (k, nm, np.abs(a - bb).max() / s2)
            assert np.abs(d["CL"][:9, 9:]).max() > 0 and d["delta_prior"] > 0
        else:
            assert np.all(Cm[:9, 9:] == 0) and np.all(Cm[9:, :9] == 0)
        checked += 1
    assert checked >= 2


@pytest.mark.parametrize("kw", [dict(), dict(cluster=True, skip_frac=0.3), dict(th_radius=4.0)])
def test_fuse_equal_reference(kw):
    """ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:1152-1165) compiled unchanged over the compiled
    SearchByProjectionBase, KeyFrame::FuseMP recording: viewing-cone test on, the keyframe's bf for the stereo gate, `bestDist <=
    TH_LOW` — the keypoint every map point is fused into and nFused equal the rule the host mirror applies to the search output
    (vieo_slam_b200/api.py ORBmatcher.Fuse: best >= 0 and dist <= 50)."""
    synth = synth_mod()
    pb = synth.make_fuse_problem(71, **kw)          # defaults: use_bf, check_viewing_angle — what Fuse itself passes
    bo, do, _ = O.proj_search(pb)
    want = np.where((bo >= 0) & (do <= 50), bo, -1).astype(np.int32)
    hit, nf = R.fuse(pb)
    assert np.array_equal(hit, want)
    fr = pb["frames"]
    assert np.array_equal(nf, [int((want[int(f["q_begin"]):int(f["q_begin"]) + int(f["n_q"])] >= 0).sum()) for f in fr])
    assert (want >= 0).sum() > 500 and ((bo >= 0) & (do > 50)).sum() > 20      # some found keypoints are too far in Hamming distance


@pytest.mark.parametrize("seed,th", [(3, 7.5), (4, 7.5), (5, 4.0)])
def test_search_by_sim3_equal_reference(seed, th):
    """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1222-1302) compiled unchanged over the compiled SearchByProjectionBase: the similarity
    algebra that forms the two search poses, the already-matched masks, the two searches in the one-match mode (the arg-min keypoint
    must hold a map point, bestDist <= TH_HIGH), the agreement check — vpMatches12 and nFound equal the oracle's two-frame search
    followed by the host mirror's agreement rule (vieo_slam_b200/api.py ORBmatcher.sim3_agreement, plain numpy)."""
    import sim3_search_data as D
    import vieo_slam_b200.api as api
    s1, s2, sim3, prior = D.make(seed, th=th)
    m12, n, p21, p12 = R.search_by_sim3(s1, s2, sim3, th, prior)
    assert np.abs(p21[:9] - s2["frame"]["Rcw"]).max() < 1e-5 and np.abs(p12[9:] - s1["frame"]["tcw"]).max() < 1e-4   # S12 is consistent
    pb = D.flat_problem(s1, s2, p21, p12, prior, th)
    best, dist, _ = O.proj_search(pb)
    N1, N2 = len(s1["kps"]), len(s2["kps"])
    want, nf = api.ORBmatcher.sim3_agreement(best, dist, pb["has_mp"], N2, N1, prior)
    assert nf == n and np.array_equal(want, m12), (seed, nf, n)
    assert n > 200 and (prior >= 0).sum() > 10 and np.array_equal(m12[prior >= 0], prior[prior >= 0])
    bA, dA = best[:N1], dist[:N1]
    assert ((bA >= 0) & (dA <= 100) & ~pb["has_mp"][:N2][np.maximum(bA, 0)]).sum() > 10     # arg-min keypoints without a map point: no match
    one_way = (bA >= 0) & (dA <= 100) & pb["has_mp"][:N2][np.maximum(bA, 0)] & (m12 < 0)
    assert one_way.sum() > 0                                                                   # found one way only: rejected by the agreement

#!/usr/bin/env python3
"""bench.py — stereo-VIO front-end frames/s on synthetic EuRoC-MH05-shaped input (BASELINE.json configs[1]).

A "step" = one batch of F stereo frames (2F 752x480 u8 images) through the hot path:
  ORB extraction x2 cameras (pyramid, per-cell FAST, quadtree, orientation, rBRIEF) + 256-bit brute-force
  left->right knnMatch(k=2) [+ IMU pre-integration + 2x PoseOptimization once those stages are enabled].
`value`  : whole-job frames/s with inputs resident in HBM (device API, CUDA events, max over ranks).
`e2e`    : the same through the host-buffer C-ABI call (vieo_frontend_process): pinned host images in,
           keypoints/descriptors/matches out, copies inside the timed region.
`--impl reference` : the CPU oracle (the reference cannot be built offline: DESIGN.md) on the host cores.
Multi-GPU: frames are independent units -> sharded across ranks, no data-path collective ("weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one stream per handle (extractors, front-end chunks, BA engines, staging): more hardware queues than the default 8, or
# unrelated streams serialise on a shared queue; must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vieo_slam_b200.synth import EUROC, stereo_stream  # noqa: E402

METRIC = "stereo-VIO frames/sec (ORB+PoseOpt+LocalBA) on EuRoC MH05 at 1 GPU vs CPU ref"
UNIT = "frames/s"
W, H = EUROC["w"], EUROC["h"]
LEVEL_PX = 1117367            # sum of the 8 level sizes
ORB_BYTES_PER_IMAGE = 1189367  # SURVEY.md §8d algorithmic bytes for the whole extractor
FAST_BYTES_PER_IMAGE = LEVEL_PX + 700 * 4  # k_fast_cells: every level pixel once + one count per cell


def workload_name(frames):
    return (f"EuRoC MH05 stereo-VIO 1200 feats (configs[1]), synthetic 752x480 stereo + 200 Hz IMU, batches of {frames} frames: "
            "ORBextractor x2 + ComputeStereoMatches + IMU pre-integration + SearchByProjection (last frame) + isInFrustum + SearchByProjection (local map) "
            "+ 2x PoseOptimization (PVR) per frame, "
            "LocalBundleAdjustmentNavStatePRV every 8th frame")

def workload_config(frames, pool, with_lba):
    """The `config` object of the JSON line: the workload only, so that both arms print the same bytes."""
    return {"workload": workload_name(frames), "frames_per_step_per_gpu": frames, "image": f"{W}x{H}", "nfeatures": 1200, "levels": 8,
            "cache": f"inputs rotate over {pool} batches ({pool * frames * 2 * H * W / 1e6:.0f} MB > 126 MB L2)",
            "stages": ["orb_extract_x2", "stereo_matches_rectified", "imu_preint", "search_by_projection_last_frame",
                       "is_in_frustum+search_by_projection_local_map", "pose_opt_x2"] + (["local_ba_prv"] if with_lba else []),
            "pose_opt_points": list(POSE_POINTS), "sbp_queries": list(SBP_QUERIES), "lba_every": LBA_EVERY if with_lba else 0,
            "lba_window": LBA_SHAPE}



def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if mx and s > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# Workload: BASELINE.json configs[1] as SURVEY.md 8d restates it.  Per stereo frame: ORBextractor x2 + ComputeStereoMatches,
# one IMU pre-integration (10 samples at 200 Hz / 20 fps), two PoseOptimization calls (TrackWithMotionModel with ~350
# matches, TrackLocalMap with ~550; IMU/PVR vertex, 15 % outliers, 70 % stereo); every LBA_EVERY-th frame is a keyframe
# and triggers one LocalBundleAdjustmentNavStatePRV window (N_local = 10, 20 fixed keyframes, 1500 points).
LBA_EVERY = 8
POSE_POINTS = (350, 550)
SBP_QUERIES = (700, 2500)   # map points per frame: last frame's / local-map candidates tested by isInFrustum (~60 % in view)
LBA_SHAPE = dict(n_local=10, n_fixed=20, n_points=1500)


def make_tracking_inputs(seed, F, preint_fn):
    """F frames of tracking-thread inputs: IMU segments + 2F PoseOptimization problems (numpy, C-ABI layouts)."""
    from vieo_slam_b200 import synth
    seq = synth.vio_sequence(seed, F + 1, speed=1.5, rot=1.0)
    imu, t = seq["imu"], seq["times"]
    seg, smp, tt, bb = [0], [], [], []
    for k in range(1, F + 1):
        lo = max(np.searchsorted(imu[:, 0], t[k - 1], "right") - 1, 0)
        hi = min(np.searchsorted(imu[:, 0], t[k], "left") + 1, len(imu))
        smp.append(imu[lo:hi]); seg.append(seg[-1] + hi - lo); tt.append((t[k - 1], t[k]))
        bb.append(np.r_[seq["truth"][k - 1]["bg"], seq["truth"][k - 1]["ba"]])
    imu_in = (np.ascontiguousarray(np.vstack(smp)), np.asarray(seg, np.int32), np.asarray(tt), np.asarray(bb))
    pre = preint_fn(*imu_in)
    pre = np.concatenate([pre[:1], pre])  # make_pose_problems indexes preints by frame
    cam = synth.euroc_camera()
    sets = [synth.make_pose_problems(seq, pre, cam, n_points=n, seed=seed + i, compute_marg=True, chain_prior=(i == 0))
            for i, n in enumerate(POSE_POINTS)]
    # concatenate the two PoseOptimization calls of every frame into one batch of 2F problems
    pbs = np.concatenate([sets[0][0], sets[1][0]])
    off = len(sets[0][1])
    pbs["edge_begin"][F:] += off
    pbs["edge_end"][F:] += off
    arrays = [np.concatenate([sets[0][i], sets[1][i]]) for i in range(1, 5)]
    # the two guided searches of every frame: TrackWithMotionModel (last frame's map points, th = 7 for stereo,
    # src/Tracking.cc:292-297) and TrackLocalMap (local map points in view, th = 1, :2367)
    sbp = [synth.make_sbp_problem(seed + 7, n_frames=F, mode=synth.SBP_LAST_FRAME, n_kp=1200, n_q=SBP_QUERIES[0], th=7.0),
           # SearchLocalPoints: Frame::isInFrustum over the candidates, then the search over those in view (:2308-2368)
           synth.make_frustum_problem(seed + 8, n_frames=F, n_kp=1200, n_q=SBP_QUERIES[1], th=1.0, blocked_frac=0.3,
                                      skip_frac=0.1)]
    return dict(imu=imu_in, pbs=pbs, cam=cam, Xw=arrays[0], obs=arrays[1], w=arrays[2], flags=arrays[3], seq=seq, sbp=sbp)


def single_frame_latency(api, imgs1, trk, pre_gpu, matcher, device, reps=20):
    """Median / minimum wall time of ONE stereo frame through the host-buffer API (copies included): front-end (two
    extractions + rectified-stereo association), IMU pre-integration, the two guided searches with the visibility test,
    two PoseOptimization calls.  imgs1: (1, 2, H, W) host images."""
    from vieo_slam_b200 import synth
    fe1 = api.StereoFrontend(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H,
                             max_frames=1, device=device)
    outs1 = fe1.alloc_outputs(1, pinned=True)
    smp, seg, tt, bb = trk["imu"]
    imu1 = (np.ascontiguousarray(smp[seg[0]:seg[1]]), np.array([0, seg[1] - seg[0]], np.int32), tt[:1], bb[:1])
    sbp1 = synth.make_sbp_problem(91, n_frames=1, mode=synth.SBP_LAST_FRAME, n_kp=1200, n_q=SBP_QUERIES[0], th=7.0)
    slp1 = synth.make_frustum_problem(92, n_frames=1, n_kp=1200, n_q=SBP_QUERIES[1], th=1.0, blocked_frac=0.3, skip_frac=0.1)
    F = len(trk["pbs"]) // 2
    pbs1 = np.ascontiguousarray(trk["pbs"][[0, F]])
    bf = float(np.float32(EUROC["bf"])); minz = float(np.float32(EUROC["bf"]) / np.float32(EUROC["fx"]))
    stages = {"front_end": [], "tracking": []}

    def one():
        t0 = time.perf_counter()
        fe1.process(imgs1, outs1)
        fe1.stereo_rectified(1, bf, minz)
        t1 = time.perf_counter()
        pre_gpu.preintegrate_batch(*imu1)
        matcher.SearchByProjection(sbp1)
        matcher.SearchLocalPoints(slp1, want_tracking_info=False)
        api.Optimizer.PoseOptimizationBatch(pbs1, trk["cam"], trk["Xw"], trk["obs"], trk["w"], trk["flags"], device=device)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1
    for _ in range(3):
        one()
    for _ in range(reps):
        a, b = one()
        stages["front_end"].append(1e3 * a); stages["tracking"].append(1e3 * b)
    fe1.close()
    tot = np.array(stages["front_end"]) + np.array(stages["tracking"])
    return {"median": float(np.median(tot)), "min": float(tot.min()), "front_end_median": float(np.median(stages["front_end"])),
            "tracking_median": float(np.median(stages["tracking"])), "reps": reps,
            "what": "one stereo frame, host buffers in and out, tracking stages in sequence on one thread"}


def make_lba_windows(seed, n, preint_fn):
    from vieo_slam_b200 import synth
    out = []
    for i in range(n):
        seq = synth.vio_sequence(seed + 31 * i, 130, speed=1.5, rot=1.0)
        kf = list(range(0, 130, 4))
        imu, t = seq["imu"], seq["times"]
        seg, smp, tt, bb = [0], [], [], []
        for k in range(1, len(kf)):
            lo = max(np.searchsorted(imu[:, 0], t[kf[k - 1]], "right") - 1, 0)
            hi = min(np.searchsorted(imu[:, 0], t[kf[k]], "left") + 1, len(imu))
            smp.append(imu[lo:hi]); seg.append(seg[-1] + hi - lo); tt.append((t[kf[k - 1]], t[kf[k]]))
            bb.append(np.r_[seq["truth"][kf[k - 1]]["bg"], seq["truth"][kf[k - 1]]["ba"]])
        pre = preint_fn(np.vstack(smp), np.asarray(seg, np.int32), np.asarray(tt), np.asarray(bb))
        pre = np.concatenate([pre[:1], pre])
        out.append(synth.make_lba_problem(seq, pre, kf, synth.euroc_camera(), seed=seed + i, **LBA_SHAPE))
    return out


# ------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    return O


def oracle_preint_fn():
    O = _oracle()
    nz = O.imu_noise()

    def fn(samples, seg, tt, bb):
        out = np.zeros(len(seg) - 1, O.PREINT_DTYPE)
        for k in range(len(seg) - 1):
            out[k] = O.imu_preintegrate(samples[seg[k]:seg[k + 1]], tt[k][0], tt[k][1], bb[k][:3], bb[k][3:], nz)
        return out
    return fn


def cpu_pipeline(images, trk, lbas, threads_like_reference=True, workers=None):
    """The reference's CPU path (oracle port) over `images` (n,2,H,W): per frame extraction on one thread per camera,
    stereo knn + IMU pre-integration + 2x PoseOptimization on the tracking thread; LocalBA windows on the LocalMapping
    thread (src/Frame.cc:259-278, src/Tracking.cc, src/LocalMapping.cc).  With threads_like_reference=False frames are
    instead spread over `workers` threads (all host cores).  Returns (frames/s, seconds)."""
    O = _oracle()
    from concurrent.futures import ThreadPoolExecutor
    n = images.shape[0]
    nz = O.imu_noise()
    cam = trk["cam"]
    F = len(trk["pbs"]) // 2
    tl = threading.local()

    def orb():
        if not hasattr(tl, "orb"):
            tl.orb = O.OrbOracle(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"])
        return tl.orb

    def orb2():  # the right camera's extractor (its pyramid is read by the stereo matcher)
        if not hasattr(tl, "orb2"):
            tl.orb2 = O.OrbOracle(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"])
        return tl.orb2

    bf = np.float32(EUROC["bf"]); minz = np.float32(bf / np.float32(EUROC["fx"]))

    def tracking(f, L, R):
        # L / R = (oracle extractor that just ran on the view, keypoints, descriptors)
        O.stereo_matches(L[0], L[1], L[2], R[0], R[1], R[2], bf, minz)
        k = f % F
        smp, seg, tt, bb = trk["imu"]
        O.imu_preintegrate(smp[seg[k]:seg[k + 1]], tt[k][0], tt[k][1], bb[k][:3], bb[k][3:], nz)
        O.search_by_projection(trk["sbp"][0], frames=[k])
        O.search_local_points(trk["sbp"][1], frames=[k])
        for j in (k, F + k):
            O.pose_optimization(trk["pbs"][j:j + 1], cam, trk["Xw"], trk["obs"], trk["w"], trk["flags"])

    def lba_thread(count):
        for i in range(count):
            O.local_ba_prv(lbas[i % len(lbas)], cam)

    n_lba = n // LBA_EVERY
    if threads_like_reference:
        with ThreadPoolExecutor(2) as cams, ThreadPoolExecutor(1) as lm:
            orbs = [O.OrbOracle(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"])
                    for _ in range(2)]
            a = cams.submit(lambda: orbs[0].extract(images[0, 0])); b = cams.submit(lambda: orbs[1].extract(images[0, 1]))
            a.result(); b.result()
            t0 = time.perf_counter()
            fut = lm.submit(lba_thread, n_lba)
            for f in range(n):
                a = cams.submit(lambda f=f: orbs[0].extract(images[f, 0]))
                b = cams.submit(lambda f=f: orbs[1].extract(images[f, 1]))
                ra, rb = a.result(), b.result()
                tracking(f, (orbs[0], ra[1], ra[2]), (orbs[1], rb[1], rb[2]))
            fut.result()
            dt = time.perf_counter() - t0
        return n / dt, dt

    def work(f):
        oL, oR = orb(), orb2()
        _, kl, dl, _ = oL.extract(images[f, 0])
        _, kr, dr, _ = oR.extract(images[f, 1])
        tracking(f, (oL, kl, dl), (oR, kr, dr))
        if f % LBA_EVERY == LBA_EVERY - 1:
            O.local_ba_prv(lbas[(f // LBA_EVERY) % len(lbas)], cam)

    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(lambda f: orb().extract(images[f % n, 0]), range(workers)))  # warm: per-thread oracles
        t0 = time.perf_counter()
        list(ex.map(work, range(n)))
        dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = args.frames  # the same batch per step as the GPU arm (same `config`)
    frames -= frames % LBA_EVERY
    imgs = stereo_stream(frames, 505, dark_every=16).reshape(frames, 2, H, W)
    pre = oracle_preint_fn()
    trk = make_tracking_inputs(505, frames, pre)
    lbas = make_lba_windows(203, 2, pre)
    for _ in range(min(args.warmup, 1)):
        cpu_pipeline(imgs[:cores], trk, lbas, False, cores)
    tot_t, tot_f = 0.0, 0
    for _ in range(args.steps):
        _, dt = cpu_pipeline(imgs, trk, lbas, False, cores)
        tot_t += dt
        tot_f += frames
    v = tot_f / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
            "config": workload_config(frames, args.pool, True),
            "run_info": {"note": "CPU oracle port of the reference path (the BA / matcher / IMU units need Eigen/Sophus/g2o: unbuildable "
                                 "offline; the extractor restatement is pinned to the compiled reference, oracle/_ref)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{frames} synthetic stereo frames per step (+{frames // LBA_EVERY} LocalBA windows), "
                                       "frames spread over one worker thread per core"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import vieo_slam_b200.api as api
    from concurrent.futures import ThreadPoolExecutor

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # Each rank runs ~10 host threads (front-end lanes, tracking stages, the LocalBA driver) that mostly wait for the device;
    # the driver's default makes every waiter spin.  When the ranks of this node together outnumber the cores this process
    # may use, waiters yield between polls instead, so the threads that enqueue work are not starved (round 1: 0.53 scaling
    # efficiency at 8 ranks on a 32-core mask with nothing but the host as the limiter).
    try:
        cores_allowed = len(os.sched_getaffinity(0))
    except AttributeError:
        cores_allowed = os.cpu_count() or 1
    host_sync = "auto"
    if 10 * world > cores_allowed:
        api.set_host_sync(2, device=local_rank)
        host_sync = "yield"
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    F = args.frames
    pool = args.pool
    n_img = 2 * F
    n_lba = F // LBA_EVERY if args.lba else 0
    # SM partition: the LocalBA streams own `ba_sms` SMs, front-end + tracking the rest (include/vieo_b200.h)
    part = api.SmPartition(args.ba_sms, device=local_rank) if (args.ba_sms > 0 and n_lba) else None
    if part:
        part.bind(api.SM_FRONTEND)   # every handle / staging stream this thread creates from here on
    orb = api.ORBextractor(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H,
                           max_batch=n_img, device=local_rank)
    cap = orb.cap
    pre_gpu = api.IMUPreintegrator(device=local_rank)
    # synthetic MH05-shaped stream: `pool` distinct batches so the per-step input (pool*F*722 KB > L2) is not cache resident
    host = stereo_stream(F * pool, 505 + rank, dark_every=16).reshape(pool, F, 2, H, W)
    trk = make_tracking_inputs(505 + rank, F, pre_gpu.preintegrate_batch)
    lbas = make_lba_windows(203 + rank, max(1, min(args.lba_windows, max(n_lba, 1))), pre_gpu.preintegrate_batch) if n_lba else []
    host_t = torch.from_numpy(host).pin_memory()
    dev_imgs = host_t.to(dev, non_blocking=True)
    kps = torch.empty((n_img, cap, 6), dtype=torch.float32, device=dev)
    desc = torch.empty((n_img, cap, 32), dtype=torch.uint8, device=dev)
    nkp = torch.empty((n_img,), dtype=torch.int32, device=dev)
    midx = torch.empty((F, cap, 2), dtype=torch.int32, device=dev)
    mdist = torch.empty((F, cap, 2), dtype=torch.int32, device=dev)
    s_ur = torch.empty((F, cap), dtype=torch.float32, device=dev)
    s_dp = torch.empty((F, cap), dtype=torch.float32, device=dev)
    s_sad = torch.empty((F, cap), dtype=torch.int32, device=dev)
    BF = float(np.float32(EUROC["bf"])); MINZ = float(np.float32(EUROC["bf"]) / np.float32(EUROC["fx"]))

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    smp, seg, tt, bb = trk["imu"]
    d_smp, d_seg, d_tt, d_bb = to_dev(smp), to_dev(seg), to_dev(tt), to_dev(bb)
    d_pre = torch.empty(F * api.PREINT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_pbs, d_cam = to_dev(trk["pbs"]), to_dev(np.asarray(trk["cam"]).reshape(1))
    d_Xw, d_obs, d_w, d_fl = to_dev(trk["Xw"]), to_dev(trk["obs"]), to_dev(trk["w"]), to_dev(trk["flags"])
    n_pb, n_edges = len(trk["pbs"]), len(trk["flags"])
    d_res = torch.empty(n_pb * api.POSEOPT_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_outl = torch.empty(n_edges, dtype=torch.uint8, device=dev)
    d_chi = torch.empty(n_edges, dtype=torch.float64, device=dev)
    sbp_dev = []
    for pb in trk["sbp"]:
        d = {k: to_dev(v) for k, v in pb.items() if isinstance(v, np.ndarray)}
        q = api.VieoSbpQueries()
        for name in ("Xw", "level", "angle", "proj", "viewcos", "depth", "desc", "flags"):
            setattr(q, name, d["q_" + name].data_ptr())
        nk, nq = len(pb["kps"]), len(pb["q_level"])
        out = [torch.empty(n, dtype=torch.int32, device=dev) for n in (nk, nq, nq, F)]
        scr = torch.empty(api.lib().vieo_sbp_scratch_bytes(nq), dtype=torch.uint8, device=dev)
        fr = None
        if "frustum" in pb:
            # the local-map search reads the tracking info k_frustum leaves in HBM
            ff = np.ascontiguousarray(pb["frustum"], api.FRUSTUM_FRAME_DTYPE).copy()
            for G in ff:
                G["level_ratio"] = api.frustum_level_table(float(G["log_scale_factor"]), int(G["n_levels"]))
            fr = dict(frames=to_dev(ff), inview=torch.empty(nq, dtype=torch.uint8, device=dev),
                      proj=torch.empty(3 * nq, dtype=torch.float32, device=dev), level=torch.empty(nq, dtype=torch.int32, device=dev),
                      viewcos=torch.empty(nq, dtype=torch.float32, device=dev), depth=torch.empty(nq, dtype=torch.float32, device=dev),
                      n_inview=torch.empty(F, dtype=torch.int32, device=dev))
            for name in ("proj", "level", "viewcos", "depth"):
                setattr(q, name, fr[name].data_ptr())
        sbp_dev.append((int(pb["mode"]), d, q, out, scr, fr))

    def sbp_enqueue(stream, which=None):
        for mode, d, q, out, scr, fr in (sbp_dev if which is None else sbp_dev[which:which + 1]):
            if fr is not None:
                api.frustum_batch_dev(fr["frames"].data_ptr(), F, d["p_wP"].data_ptr(), d["p_normal"].data_ptr(),
                                      d["p_max_dist"].data_ptr(), d["p_min_dist"].data_ptr(), d["p_skip"].data_ptr(),
                                      fr["inview"].data_ptr(), fr["proj"].data_ptr(), fr["level"].data_ptr(),
                                      fr["viewcos"].data_ptr(), fr["depth"].data_ptr(), fr["n_inview"].data_ptr(), stream)
            api.sbp_batch_dev(mode, d["frames"].data_ptr(), F, d["kps"].data_ptr(), d["uright"].data_ptr(), d["desc"].data_ptr(),
                              q, d["kp_blocked"].data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                              out[3].data_ptr(), scr.data_ptr(), scr.numel(), stream)
    n_workers = max(1, min(args.lba_workers, n_lba)) if n_lba else 0
    if part:
        part.bind(api.SM_BA)
    # one engine (handle + stream) per LocalBA window of a step; each host worker thread drives its share of them through
    # the asynchronous C ABI: every window is enqueued (device-side LM loops) before any is awaited
    if args.prio:
        api.lib().vieo_ba_stream_priority(0)
    bas = [api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16, device=local_rank)
           for _ in range(n_lba)]
    lba_pool = ThreadPoolExecutor(n_workers) if n_workers else None
    torch.cuda.synchronize()
    if part:
        part.bind(api.SM_FRONTEND)
        main = torch.cuda.ExternalStream(part.stream(api.SM_FRONTEND), device=dev)
        side, side_b, side_c = (torch.cuda.ExternalStream(part.stream(api.SM_FRONTEND), device=dev) for _ in range(3))
        torch.cuda.set_stream(main)
    else:
        # the extractor + stereo chain is the step's critical path: it gets the high-priority stream, the tracking stages and
        # the LocalBA engines (pipelined over a whole step) normal priority
        main = torch.cuda.Stream(priority=-1) if args.prio else torch.cuda.current_stream()
        torch.cuda.set_stream(main)
        side, side_b, side_c = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    # the batch holds 128 independent frames at different stages of Tracking::Track(): the extractor + stereo association
    # (main), IMU + TrackWithMotionModel's search (side), SearchLocalPoints (side_b) and the PoseOptimization calls (side_c)
    # are enqueued on four streams, the LocalBA windows on their engines' streams
    sides = (side, side_b, side_c)

    # LocalMapping runs beside Tracking in the reference (its own thread, src/LocalMapping.cc): a window begun in step i is
    # awaited when its engine is needed again in step i + 1, so its latency overlaps the next frames' tracking; every
    # window begun inside the timed region is also ended inside it (lba_drain before the closing event).
    inflight = [False] * max(n_lba, 1)

    def lba_end(i, on):
        out = bas[i].end()
        inflight[i] = False
        assert out["res"]["accepted"] == 1
        if on:
            lba_ms_acc.append((bas[i].last_ms(), int(out["res"]["iterations"].sum())))
        return bas[i].last_launches()

    def lba_job(wk, i0, on=False):
        n = 0
        for i in range(i0, n_lba, n_workers):
            if inflight[i]:
                n += lba_end(i, on)
            bas[i].begin(lbas[i % len(lbas)], trk["cam"])
            inflight[i] = True
        return n

    def lba_drain(on=False):
        return sum(lba_end(i, on) for i in range(n_lba) if inflight[i])

    # CUDA-event brackets around every kernel group of the step, recorded on the stream the group is launched on (the
    # extractor's four stages are bracketed inside the library, vieo_orb_profile; a LocalBA window on its engine's stream,
    # vieo_ba_last_ms).  Read only after the timed region: no synchronisation is added to it.
    TIMED = ("stereo_match", "imu_preint", "search_by_projection_last_frame", "is_in_frustum+search_by_projection_local_map",
             "pose_opt_x2")
    evs = {k: [] for k in TIMED}
    lba_ms_acc = []

    def timed(name, stream, fn, on):
        if not on:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        evs[name].append((a, b))

    # DIAGNOSTIC ONLY (never set by the driver): VIEO_BENCH_SKIP=orb,stereo,imu,sbp0,sbp1,po leaves kernel groups out of the
    # step to read each group's marginal cost; the JSON line then carries "diagnostic_skip" and is not a bench value
    skip = set(filter(None, os.environ.get("VIEO_BENCH_SKIP", "").split(",")))

    def step(i, on=False):
        futs = [lba_pool.submit(lba_job, wk, wk, on) for wk in range(n_workers)]
        imgs = dev_imgs[i % pool]
        s = main.cuda_stream
        if "orb" not in skip:
            orb.extract_batch_dev(imgs.data_ptr(), n_img, H * W, W, kps.data_ptr(), desc.data_ptr(), cap, nkp.data_ptr(), s)
        if "stereo" not in skip:
            timed("stereo_match", main, lambda: orb.stereo_match_dev(F, kps.data_ptr(), desc.data_ptr(), nkp.data_ptr(), cap, BF, MINZ,
                                                                     s_ur.data_ptr(), s_dp.data_ptr(), s_sad.data_ptr(), s), on)
        s2, s3, s4 = side.cuda_stream, side_b.cuda_stream, side_c.cuda_stream
        if "imu" not in skip:
            timed("imu_preint", side, lambda: pre_gpu.preintegrate_batch_dev(d_smp.data_ptr(), d_seg.data_ptr(), d_tt.data_ptr(),
                                                                             d_bb.data_ptr(), F, d_pre.data_ptr(), s2), on)
        if "sbp0" not in skip:
            timed("search_by_projection_last_frame", side, lambda: sbp_enqueue(s2, 0), on)
        if "sbp1" not in skip:
            timed("is_in_frustum+search_by_projection_local_map", side_b, lambda: sbp_enqueue(s3, 1), on)
        if "po" not in skip:
            timed("pose_opt_x2", side_c, lambda: api.Optimizer.pose_opt_batch_dev(
                d_pbs.data_ptr(), n_pb, d_cam.data_ptr(), d_Xw.data_ptr(), d_obs.data_ptr(), d_w.data_ptr(), d_fl.data_ptr(),
                d_res.data_ptr(), d_outl.data_ptr(), d_chi.data_ptr(), s4), on)
        for sd in sides:
            main.wait_stream(sd)
        return sum(f.result() for f in futs)

    ba_launches = 0
    for i in range(args.warmup):
        for sd in sides:
            sd.wait_stream(main)
        ba_launches = max(ba_launches, step(i))
    launches_per_step = orb.last_launches() + 1 + 2 + 2 + ba_launches  # extractor + stereo match + (imu, pose opt) + 2 guided searches + LocalBA
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    orb.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for sd in sides:
        sd.wait_stream(main)
    for i in range(args.steps):
        step(args.warmup + i, True)
    lba_drain(True)   # every LocalBA window of the timed region has finished before the closing event
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    stage_ms, ncalls = orb.profile_read()
    orb.profile(False)
    if dist_on:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * F * args.steps / (ms_max / 1e3)

    # stage times in isolation (CUDA events, device-resident), for the report
    def time_it(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    s = main.cuda_stream
    iso = {
        "imu_preint": time_it(lambda: pre_gpu.preintegrate_batch_dev(d_smp.data_ptr(), d_seg.data_ptr(), d_tt.data_ptr(),
                                                                     d_bb.data_ptr(), F, d_pre.data_ptr(), s)),
        "pose_opt_x2": time_it(lambda: api.Optimizer.pose_opt_batch_dev(d_pbs.data_ptr(), n_pb, d_cam.data_ptr(),
                                                                        d_Xw.data_ptr(), d_obs.data_ptr(), d_w.data_ptr(),
                                                                        d_fl.data_ptr(), d_res.data_ptr(), d_outl.data_ptr(),
                                                                        d_chi.data_ptr(), s)),
    }
    iso["search_by_projection_x2"] = time_it(lambda: sbp_enqueue(s))
    if n_lba:
        t0 = time.perf_counter()
        for _ in range(3):
            bas[0].LocalBundleAdjustmentNavStatePRV(lbas[0], trk["cam"])
        iso["local_ba_window_ms"] = (time.perf_counter() - t0) / 3 * 1e3

    # ---- e2e: host buffers through the host-buffer C-ABI calls, copies inside the timed region; threads as in the
    # reference: front-end, tracking (IMU + PoseOptimization), LocalMapping workers
    # Two batches are in flight (E2E_DEPTH): batch i + 1 is uploaded and extracted while batch i finishes its tracking
    # stages and downloads, like consecutive frames of a live run overlap across the reference's threads.  Every batch still
    # pays its own host->device and device->host copies inside the timed region; each lane owns its front-end handle and
    # pinned output buffers, the tracking stages run on 3 host threads per lane (thread-local staging in the library).
    E2E_DEPTH = 2
    fes = [api.StereoFrontend(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H,
                              max_frames=F, device=local_rank) for _ in range(E2E_DEPTH)]
    outs_l = [fe_.alloc_outputs(F, pinned=True) for fe_ in fes]
    outs = outs_l[0]
    host_np = host_t.numpy()
    trk_pool = ThreadPoolExecutor(3 * E2E_DEPTH, initializer=(lambda: part.bind(api.SM_FRONTEND)) if part else None)
    step_pool = ThreadPoolExecutor(E2E_DEPTH, initializer=(lambda: part.bind(api.SM_FRONTEND)) if part else None)
    matcher = api.ORBmatcher(0.8, True, device=local_rank)

    def trk_motion_model():
        pre_gpu.preintegrate_batch(*trk["imu"])
        return int(matcher.SearchByProjection(trk["sbp"][0])[3][0])

    def trk_local_map():
        return int(matcher.SearchLocalPoints(trk["sbp"][1], want_tracking_info=False)[4][0])

    def trk_pose_opt():
        r = api.Optimizer.PoseOptimizationBatch(trk["pbs"], trk["cam"], trk["Xw"], trk["obs"], trk["w"], trk["flags"],
                                                device=local_rank)
        return int(r[0]["n_inliers"][0])

    lba_drain()

    def e2e_step(i, futs):
        lane = i % E2E_DEPTH
        fe_ = fes[lane]
        fts = [trk_pool.submit(fn) for fn in (trk_motion_model, trk_local_map, trk_pose_opt)]
        res = fe_.process(host_np[i % pool], outs_l[lane])
        st = fe_.stereo_rectified(F, BF, MINZ)
        _ = int(res[2][0]) + sum(ft.result() for ft in fts) + int(st[2][0, 0])
        for f in futs:
            f.result()

    def e2e_run(i0, n):
        """n batches, E2E_DEPTH in flight; the LocalBA jobs are submitted in batch order (one driver thread, so the windows of
        batch i are ended before batch i + 1 begins new ones on the same engines)."""
        inflight_steps = []
        for i in range(i0, i0 + n):
            futs = [lba_pool.submit(lba_job, wk, wk) for wk in range(n_workers)]
            inflight_steps.append(step_pool.submit(e2e_step, i, futs))
            if len(inflight_steps) >= E2E_DEPTH:
                inflight_steps.pop(0).result()
        for f in inflight_steps:
            f.result()

    e2e_run(0, max(8, args.warmup))  # every pool thread has run every stage once (thread-local staging is allocated on first use)
    # ... which a fixed count does not guarantee: the pools hand a stage to whichever thread is free, and a thread that meets a
    # larger stage for the first time grows its pinned staging (cudaHostAlloc, slow and device-synchronising; slower still when
    # 8 ranks do it at once).  Untimed chunks of K steps are repeated until two consecutive ones agree within 10 % (at most 6).
    prev_chunk = None
    for _ in range(6):
        lba_drain()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(0, args.steps)
        lba_drain()
        torch.cuda.synchronize()
        chunk = time.perf_counter() - t0
        flag = torch.tensor([1.0 if (prev_chunk is not None and abs(chunk - prev_chunk) <= 0.1 * prev_chunk) else 0.0], device=dev)
        if dist_on:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # every rank leaves the warm-up in the same round
        prev_chunk = chunk
        if float(flag.item()) > 0:
            break
    # The host-side figure is wall clock over K steps of ~10 ms: one scheduling hiccup of the shared host moves it by tens
    # of percent, so the K steps are timed three times back to back (max over ranks each) and the MEDIAN is reported; all
    # three are in the line.  The collector is paused inside the timed regions.
    import gc
    e2e_runs = []
    for rep_ in range(3):
        lba_drain()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()
        gc.disable()
        t0 = time.perf_counter()
        e2e_run(args.warmup, args.steps)
        lba_drain()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        gc.enable()
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_runs.append(world * F * args.steps / float(t.item()))
    e2e_value = float(np.median(e2e_runs))
    lba_bytes = sum(int(np.asarray(v).nbytes) for v in lbas[0].values() if hasattr(v, "nbytes")) if lbas else 0
    trk_in = sum(int(np.asarray(a).nbytes) for a in trk["imu"]) + sum(int(trk[k].nbytes) for k in ("pbs", "Xw", "obs", "w", "flags"))
    # bytes the two host-buffer search calls stage: the last-frame search copies every array of its problem; the fused
    # SearchLocalPoints copies the map-point arrays isInFrustum reads, not the projections (they are made on the device)
    lm_keys = ("frames", "frustum", "p_wP", "p_normal", "p_max_dist", "p_min_dist", "p_skip", "q_desc", "q_flags", "kps",
               "uright", "desc", "kp_blocked")
    sbp_in = (sum(int(v.nbytes) for v in trk["sbp"][0].values() if isinstance(v, np.ndarray)) +
              sum(int(trk["sbp"][1][k].nbytes) for k in lm_keys))
    sbp_out = sum(4 * (len(pb["kps"]) + 2 * len(pb["q_level"]) + F) for pb in trk["sbp"]) + len(trk["sbp"][1]["q_level"]) + 4 * F   # + inview, n_inview
    h2d = F * 2 * H * W + trk_in + sbp_in + n_lba * lba_bytes
    d2h = (sum(int(o.nbytes) for o in outs) + 3 * F * cap * 4 + sbp_out + F * api.PREINT_DTYPE.itemsize + n_pb * api.POSEOPT_RESULT_DTYPE.itemsize
           + 9 * n_edges + n_lba * (64 * 176 + 2048 * 24 + 16384 * 9))

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    # ---- roofline: every kernel group of the step with its ALGORITHMIC bytes per launch, timed live over the
    # timed region; the dominant group = largest device time per step.  frac = achieved / measured HBM peak.
    peak, peak_src = peaks()
    nc = max(ncalls, 1)
    orb_ms = sum(stage_ms.values()) / nc
    groups = {}

    def add(name, ms_per_step, alg_bytes, launches, note):
        groups[name] = {"ms_per_step": ms_per_step, "launches_per_step": launches, "alg_bytes_per_step": int(alg_bytes),
                        "GBps": alg_bytes / (ms_per_step / 1e3) / 1e9 if ms_per_step > 0 else 0.0,
                        "frac": (alg_bytes / (ms_per_step / 1e3) / 1e9 / peak) if ms_per_step > 0 else 0.0, "bytes": note}
    add("orb_extract_x2", orb_ms, ORB_BYTES_PER_IMAGE * n_img, orb.last_launches(),
        "SURVEY 8(d): 1,189,367 B per 752x480 1200-feature image (input + levels >= 1 + 60 B per keypoint) x images per step, "
        "over the summed stage times of k_resize x7, k_fast_cells, k_quadtree, k_orient_desc")

    def ev_ms(name):
        v = [a.elapsed_time(b) for a, b in evs[name]]
        return float(np.mean(v)) if v else 0.0
    nL = EUROC["nfeatures"]
    add("stereo_match", ev_ms("stereo_match"), F * ((24 + 32) * 2 * nL + (121 + 231) * nL + 12 * nL), 1,
        "per stereo frame: keypoints + descriptors of both views (56 B x 2 x 1200), 11x11 left patch + 11x21 right window per left "
        "keypoint, 12 B out per left keypoint")
    add("imu_preint", ev_ms("imu_preint"), 56 * len(trk["imu"][0]) + 1792 * F, 1, "SURVEY 8(d): 56 B per sample + 1792 B out per interval")
    for nm, pb in zip(("search_by_projection_last_frame", "is_in_frustum+search_by_projection_local_map"), trk["sbp"]):
        nk_, nq_ = len(pb["kps"]), len(pb["q_level"])
        add(nm, ev_ms(nm), 56 * nk_ + 77 * nq_ + 4 * nk_ + 8 * nq_ + (62 * nq_ if "frustum" in pb else 0), 2 if "frustum" in pb else 1,
            "56 B per keypoint + 77 B per query read, 4 B per keypoint + 8 B per query written" +
            (" (+ 33 B read + 29 B written per map point of the visibility test)" if "frustum" in pb else ""))
    n_e = trk["pbs"]["edge_end"].astype(np.int64) - trk["pbs"]["edge_begin"].astype(np.int64)
    add("pose_opt_x2", ev_ms("pose_opt_x2"), int((6376 + 50 * n_e).sum()), 1,
        "per frame problem: 4192 B problem + 41 B per edge read once, 2184 B result + 9 B per edge written once (the 4 x 10 LM "
        "iterations run out of shared memory)")
    if lba_ms_acc:
        d0 = lbas[0]
        E_, P_, K_ = len(d0["edge_state"]), len(d0["points"]), len(d0["states"])
        b_lin = 52 * E_ + 48 * P_ + 176 * K_ + 1600 * (len(d0["imu_i"]))
        n_lin = float(np.mean([it for _, it in lba_ms_acc])) + 2   # one linearisation per LM iteration at least + one per stage
        # Every other group lives on ONE stream, so its event time is that stream's busy time per step.  The n_lba windows of a
        # step run CONCURRENTLY, one engine / stream each: the group's per-step device time is the busy time of one of those
        # streams (the mean window time), not the sum over streams, which would count the same wall interval n_lba times
        # (it was reported that way before: 105 ms "per step" of a 9.8 ms step).  The sum stays in the group for transparency.
        win_ms = float(np.mean([m for m, _ in lba_ms_acc]))
        add("local_ba_prv_windows", win_ms, n_lba * n_lin * b_lin, ba_launches,
            "SURVEY 8(d): B_lin = 52 E + 48 P + 176 K + 1.6k (K - 1) per linearisation x (LM iterations + 2) per window x windows "
            "per step; ms = device time of ONE window (the step's windows run concurrently on their own engines / streams)")
        groups["local_ba_prv_windows"]["summed_window_ms_per_step"] = win_ms * n_lba
        groups["local_ba_prv_windows"]["concurrent_windows"] = n_lba
    dom = max(groups, key=lambda k: groups[k]["ms_per_step"])
    tot_ms = sum(g["ms_per_step"] for g in groups.values())
    for g in groups.values():
        g["share_of_summed_group_time"] = g["ms_per_step"] / tot_ms if tot_ms > 0 else 0.0
    traffic, traffic_over_alg = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = {"orb_extract_x2": "orb_total"}.get(dom, dom)
        per_unit = tj["per_image"].get(key) if key == "orb_total" else tj.get("per_step_256_problems", {}).get(key)
        if per_unit is not None:
            traffic = int(per_unit) * (n_img if key == "orb_total" else 1)   # per launch, like `achieved`
            traffic_over_alg = traffic / groups[dom]["alg_bytes_per_step"]
    except Exception:
        pass
    # ---- single-frame latency (SURVEY.md 8d asks for both figures): ONE stereo frame through the same host-buffer calls,
    # nothing batched; reported beside the throughput, never part of `value`.  Guarded: a failure here leaves null.
    latency = None
    try:
        latency = single_frame_latency(api, host_np[0][:1], trk, pre_gpu, matcher, local_rank)
    except Exception as ex:   # noqa: BLE001
        print(f"single-frame latency not measured: {ex}", file=sys.stderr)
    # ---- CPU baseline: the oracle threaded like the reference, bounded sample
    cpu_frames = max(LBA_EVERY, args.cpu_frames - args.cpu_frames % LBA_EVERY)   # whole LocalBA periods, at least one
    cpu_lbas = make_lba_windows(203, 1, oracle_preint_fn()) if not lbas else lbas
    cpu_v, cpu_dt = cpu_pipeline(host[0, :cpu_frames], trk, cpu_lbas, True)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8+f64", "data": "synthetic",
        # `config` describes the WORKLOAD and is byte-identical in both arms (same frames per step, stages, sizes); what is
        # specific to this run of this arm (threads, host, isolated stage timings) sits in `run_info`
        "config": workload_config(F, pool, bool(n_lba)),
        "run_info": {"lba_workers": n_workers, "host_sync": host_sync, "host_cores_allowed": cores_allowed,
                     "sm_partition": {"ba": part.sms(api.SM_BA), "frontend_tracking": part.sms(api.SM_FRONTEND)} if part else None,
                     "isolated_stage_ms": iso},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "runs": e2e_runs, "batches_in_flight": 2,
                "how": "median of three back-to-back timings of the K steps (wall clock, max over ranks); two batches in flight, "
                       "each with its own host->device / device->host copies inside the timed region"},
        "single_frame_latency_ms": latency,
        "gpu_launches": launches_per_step * args.steps,
        **({"diagnostic_skip": sorted(skip), "invalid": "diagnostic run with kernel groups left out"} if skip else {}),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": groups[dom]["GBps"], "peak": peak, "unit": "GB/s",
                     "frac": groups[dom]["frac"], "traffic": traffic, "traffic_over_alg": traffic_over_alg, "peak_source": peak_src,
                     "alg_bytes_per_launch": groups[dom]["alg_bytes_per_step"], "launch_ms": groups[dom]["ms_per_step"],
                     "method": "SURVEY 8(d) algorithmic bytes of the group per step / its device time per step (CUDA events on its "
                               "own stream inside the timed region; the LocalBA group: busy time of one of its concurrent streams); dominant = largest "
                               "device time per step over ALL groups",
                     "orb_stage_ms_per_step": {k: v / nc for k, v in stage_ms.items()},
                     "all_groups": groups,
                     "whole_step_GBps": sum(g["alg_bytes_per_step"] for g in groups.values()) / (ms_max / args.steps / 1e3) / 1e9,
                     "whole_step_frac": sum(g["alg_bytes_per_step"] for g in groups.values()) / (ms_max / args.steps / 1e3) / 1e9 / peak},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 3, "kind": "port",
                         "sample": f"{cpu_frames} stereo frames of the same stream + {cpu_frames // LBA_EVERY} LocalBA windows; "
                                   f"one thread per camera (src/Frame.cc:259-278), one tracking thread, one LocalMapping thread "
                                   f"({os.cpu_count()} host cores available)"},
    }
    print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=128, help="stereo frames per step per GPU")
    ap.add_argument("--pool", type=int, default=4, help="distinct input batches rotated through")
    ap.add_argument("--cpu-frames", type=int, default=24)
    ap.add_argument("--lba", type=int, default=1, help="0: leave LocalBA out of the step")
    ap.add_argument("--lba-workers", type=int, default=1,
                    help="host threads driving the LocalBA windows of a step (one engine per window, enqueued asynchronously)")
    ap.add_argument("--lba-windows", type=int, default=3, help="distinct LocalBA problems generated")
    ap.add_argument("--ba-sms", type=int, default=0, help="SMs reserved for the LocalBA streams (CUDA green context); 0: no partition")
    ap.add_argument("--prio", type=int, default=0, help="1: front-end stream at high priority, LocalBA engines at normal priority")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 4],
                    help="BASELINE.json configs index: 1 = the metric's workload (default, what the driver runs); 3 = 4-camera KB8 rig, "
                         "per-camera ORB shard + pair matching (tools/bench_multicam.py); 4 = final GlobalBA with the scale vertex, "
                         "landmark-sharded (tools/bench_gba.py)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config != 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        mod = __import__("bench_multicam" if args.config == 3 else "bench_gba")
        if args.impl == "reference":
            mod.run_reference(args, rank, world)
        else:
            mod.run_gpu(args, rank, world, local_rank, ClockSampler, peaks)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:  # timing rule: at least 3 warm-up steps; the JSON line reports the count actually run
            print(f"bench.py: --warmup {args.warmup} raised to 3 (minimum for a valid measurement)", file=sys.stderr)
            args.warmup = 3
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

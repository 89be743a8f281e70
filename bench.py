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
    return (f"EuRoC MH05 stereo-VIO 1200 feats (configs[1]), synthetic 752x480 stereo, batches of {frames} frames: "
            "ORBextractor x2 + stereo knnMatch")


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
def cpu_frames_per_s(images, threads, log=None):
    """Oracle front-end over `images` (n_frames,2,H,W): ORB per camera + stereo knn2, `threads` workers."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    n = images.shape[0]
    tl = threading.local()

    def work(f):
        if not hasattr(tl, "orb"):
            tl.orb = O.OrbOracle(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"])
        _, _, dl, _ = tl.orb.extract(images[f, 0])
        _, _, dr, _ = tl.orb.extract(images[f, 1])
        O.hamming_knn2(dl, dr)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(min(threads, n))))  # warm: build per-thread oracles
        t0 = time.perf_counter()
        list(ex.map(work, range(n)))
        dt = time.perf_counter() - t0
    return n / dt, dt


def cpu_reference_threads_like_reference(images):
    """The reference's own threading for one frame: one thread per camera for extraction (src/Frame.cc:259-278),
    matching on the tracking thread; frames strictly sequential."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    orbs = [O.OrbOracle(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"])
            for _ in range(2)]
    with ThreadPoolExecutor(2) as ex:
        def frame(f):
            a = ex.submit(lambda: orbs[0].extract(images[f, 0]))
            b = ex.submit(lambda: orbs[1].extract(images[f, 1]))
            O.hamming_knn2(a.result()[2], b.result()[2])
        frame(0)
        t0 = time.perf_counter()
        for f in range(images.shape[0]):
            frame(f)
        dt = time.perf_counter() - t0
    return images.shape[0] / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = max(2 * cores, 8)
    imgs = stereo_stream(frames, 505, dark_every=16).reshape(frames, 2, H, W)
    for _ in range(args.warmup):
        cpu_frames_per_s(imgs[: max(cores, 2)], cores)
    tot_t, tot_f = 0.0, 0
    for _ in range(args.steps):
        _, dt = cpu_frames_per_s(imgs, cores)
        tot_t += dt
        tot_f += frames
    v = tot_f / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(frames), "frames_per_step": frames,
                       "note": "CPU oracle port of the reference path (reference needs OpenCV/Eigen: unbuildable offline)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{frames} synthetic stereo frames per step, one worker thread per core"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import vieo_slam_b200.api as api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    F = args.frames
    pool = args.pool
    n_img = 2 * F
    orb = api.ORBextractor(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H,
                           max_batch=n_img, device=local_rank)
    cap = orb.cap
    # synthetic MH05-shaped stream: `pool` distinct batches so the per-step input (pool*F*722 KB > L2) is not cache resident
    host = stereo_stream(F * pool, 505 + rank, dark_every=16).reshape(pool, F, 2, H, W)
    host_t = torch.from_numpy(host).pin_memory()
    dev_imgs = host_t.to(dev, non_blocking=True)
    kps = torch.empty((n_img, cap, 6), dtype=torch.float32, device=dev)
    desc = torch.empty((n_img, cap, 32), dtype=torch.uint8, device=dev)
    nkp = torch.empty((n_img,), dtype=torch.int32, device=dev)
    midx = torch.empty((F, cap, 2), dtype=torch.int32, device=dev)
    mdist = torch.empty((F, cap, 2), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        imgs = dev_imgs[i % pool]
        orb.extract_batch_dev(imgs.data_ptr(), n_img, H * W, W, kps.data_ptr(), desc.data_ptr(), cap, nkp.data_ptr(), stream)
        api.hamming_knn2_batch_dev(desc.data_ptr(), 2 * cap * 32, nkp.data_ptr(), cap, desc.data_ptr() + cap * 32,
                                   2 * cap * 32, nkp.data_ptr() + 4, cap, 2, F, midx.data_ptr(), mdist.data_ptr(), stream)

    launches_per_step = None
    for i in range(args.warmup):
        step(i)
    launches_per_step = orb.last_launches() + 1
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
    for i in range(args.steps):
        step(args.warmup + i)
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

    # ---- e2e: host buffers through the C-ABI front-end call, copies inside the timed region
    fe = api.StereoFrontend(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H,
                            max_frames=F, device=local_rank)
    outs = fe.alloc_outputs(F, pinned=True)
    host_np = host_t.numpy()
    for i in range(max(3, args.warmup)):
        fe.process(host_np[i % pool], outs)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = fe.process(host_np[(args.warmup + i) % pool], outs)
        _ = int(res[2][0])  # read a result on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * F * args.steps / float(t.item())
    h2d = F * 2 * H * W
    d2h = sum(int(o.nbytes) for o in outs)

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (live CUDA-event stage times over the timed region)
    peak, peak_src = peaks()
    dom = max(stage_ms, key=stage_ms.get)
    stage_bytes = {"fast_cells": FAST_BYTES_PER_IMAGE, "pyramid": LEVEL_PX, "orient_desc": 1200 * (43 * 43 + 56),
                   "quadtree": 25000 * 6}
    dom_ms = stage_ms[dom] / max(ncalls, 1)
    alg_bytes = stage_bytes[dom] * n_img
    achieved = alg_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    # ---- CPU baseline: the oracle threaded like the reference (1 thread per camera), bounded sample
    cpu_frames = args.cpu_frames
    cpu_v, cpu_dt = cpu_reference_threads_like_reference(host[0, :cpu_frames])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(F), "frames_per_step_per_gpu": F, "image": f"{W}x{H}", "nfeatures": 1200,
                   "levels": 8, "cache": f"inputs rotate over {pool} batches ({pool * F * 2 * H * W / 1e6:.0f} MB > 126 MB L2)",
                   "stages": ["orb_extract_x2", "stereo_knn2"]},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "alg_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
                     "stage_ms_per_step": {k: v / max(ncalls, 1) for k, v in stage_ms.items()},
                     "orb_pipeline_GBps": ORB_BYTES_PER_IMAGE * n_img * ncalls / (sum(stage_ms.values()) / 1e3) / 1e9},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 2, "kind": "port",
                         "sample": f"{cpu_frames} stereo frames of the same stream, one thread per camera like "
                                   f"src/Frame.cc:259-278 ({os.cpu_count()} host cores available)"},
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
    ap.add_argument("--frames", type=int, default=64, help="stereo frames per step per GPU")
    ap.add_argument("--pool", type=int, default=4, help="distinct input batches rotated through")
    ap.add_argument("--cpu-frames", type=int, default=24)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

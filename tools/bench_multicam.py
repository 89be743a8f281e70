"""bench.py --config 3: BASELINE configs[3] — TUM-VI corridor1-shaped 4-camera KB8 rig (4 x 512 x 512 u8, 1000 features,
lapping area = full width), per-camera ORB shard across the GPUs + the 6 camera-pair knnMatch + ratio test of
Frame::ComputeStereoFishEyeMatches (src/Frame.cc:613-663).

Sharding: unit = (camera, frame).  A "rig group" is 4 cameras.  With N = 1 the GPU holds all 4 cameras of
its frames; N = 2 two cameras each; N = 4 ONE camera per GPU; N = 8 two rig groups.  Per-GPU work is fixed (IMGS images per
step) as N grows -> "weak".  The pair matching needs both cameras' descriptors: ONE all_gather of the lapping-ordered
descriptors + counts inside the rig group per step (NCCL; 2.3 MB per rank), then every rank matches its share of the
group's frames (all 6 pairs of a frame on one GPU).
`value`: rig frames/s with the images resident in HBM; `e2e`: pinned host images in (H2D inside the timed region),
keypoints / descriptors / pair matches of the rank's frames out (D2H inside)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CAMS, SIZE, NFEAT = 4, 512, 1000
TUMVI_BYTES_PER_IMAGE = 871960  # 262,144 + 549,816 + 1000 * 60
METRIC = "4-camera rig frames/sec (ORBextractor x4 with lapping area + 6-pair knnMatch/ratio) on TUM-VI corridor1-shaped input"


def workload(imgs_per_gpu):
    return (f"TUM-VI corridor1 dist-stereo 4-cam (configs[3]): synthetic 4 x 512x512 u8 KB8-rig frames, 1000 feats, 8 levels, "
            f"lapping area = full width; {imgs_per_gpu} images per step per GPU; ORBextractor::operator() per camera + "
            f"ComputeStereoFishEyeMatches brute-force half (6 pairs knnMatch k=2 + ratio test)")


def rig_images(n_frames, seed):
    sys.path.insert(0, ROOT)
    from vieo_slam_b200.synth import texture
    big = texture(SIZE + 160, SIZE * 3 + 200, seed)
    out = np.empty((N_CAMS, n_frames, SIZE, SIZE), np.uint8)  # [camera][frame]
    for f in range(n_frames):
        for c in range(N_CAMS):
            x = 20 + c * int(SIZE * 0.4) + (5 * f) % 150
            y = (3 * f) % 150
            out[c, f] = big[y:y + SIZE, x:x + SIZE]
    return out


def cpu_rig(images_cf, n_frames):
    """The reference's CPU path for this config (oracle port): one extraction thread per camera (src/Frame.cc:259-278),
    then the 6 pair matches on the calling thread.  Returns (frames/s, seconds, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    orbs = [O.OrbOracle(NFEAT, 1.2, 8, 20, 7) for _ in range(N_CAMS)]
    lap = (0, 10000)
    cap = 1400
    with ThreadPoolExecutor(N_CAMS) as ex:
        list(ex.map(lambda c: orbs[c].extract(images_cf[c, 0], lapping=lap), range(N_CAMS)))
        t0 = time.perf_counter()
        for f in range(n_frames):
            res = list(ex.map(lambda c: orbs[c].extract(images_cf[c, f], lapping=lap), range(N_CAMS)))
            desc = np.zeros((N_CAMS, cap, 32), np.uint8)
            nk = np.zeros(N_CAMS, np.int32); nm = np.zeros(N_CAMS, np.int32)
            for c, (n, k, d, mono) in enumerate(res):
                desc[c, :n] = d; nk[c] = n; nm[c] = mono
            O.fisheye_matches(desc, nk, nm)
        dt = time.perf_counter() - t0
    return n_frames / dt, dt, N_CAMS


def run_reference(args, rank, world):
    if rank != 0:
        return
    frames = 8
    imgs = rig_images(frames, 411)
    tot, n = 0.0, 0
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_rig(imgs, 2)
    for _ in range(args.steps):
        _, dt, cores = cpu_rig(imgs, frames)
        tot += dt; n += frames
    v = n / tot
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": {"workload": workload(frames * N_CAMS), "frames_per_step": frames},
                      "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                       "sample": f"{frames} rig frames per step, one extraction thread per camera as the reference"},
                      "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_gpu(args, rank, world, local_rank, ClockSampler, peaks):
    import torch
    import vieo_slam_b200.api as api
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert world in (1, 2, 4) or world % 4 == 0, "--config 3 shards one 4-camera rig group over 1, 2 or 4 GPUs"
    gsize = min(world, N_CAMS)                 # ranks per rig group
    grank = rank % gsize
    cams_here = N_CAMS // gsize                # cameras this rank extracts
    IMGS = args.frames                         # images per step per GPU (fixed: weak scaling)
    Fg = IMGS // cams_here                     # frames of the rig group per step
    assert Fg * cams_here == IMGS and Fg % gsize == 0, "--frames must be a multiple of 16 here"
    group = None
    if world > 1:
        groups = [dist.new_group(list(range(g * gsize, (g + 1) * gsize))) for g in range(world // gsize)]
        group = groups[rank // gsize]
    pool = args.pool
    orb = api.ORBextractor(NFEAT, 1.2, 8, 20, 7, SIZE, SIZE, max_batch=IMGS, device=local_rank)
    cap = orb.cap
    # this rank's cameras of `pool` distinct batches, [pool][cam_local][frame]
    host = np.stack([rig_images(Fg, 411 + 10 * p + rank // gsize)[grank * cams_here:(grank + 1) * cams_here] for p in range(pool)])
    host_t = torch.from_numpy(host).pin_memory()
    dev_imgs = host_t.to(dev)
    lap = torch.tensor([[0, 10000]] * IMGS, dtype=torch.int32, device=dev)
    k0 = torch.empty((IMGS, cap, 6), dtype=torch.float32, device=dev); d0 = torch.empty((IMGS, cap, 32), dtype=torch.uint8, device=dev)
    k1 = torch.empty_like(k0)
    # gathered [rank][cam_local][frame] == [camera][frame]: descriptors, then counts
    d_all = torch.empty((N_CAMS * Fg, cap, 32), dtype=torch.uint8, device=dev)
    cnt = torch.empty((2, IMGS), dtype=torch.int32, device=dev)           # n_kp | n_mono of this rank
    cnt_all = torch.empty((gsize, 2, IMGS), dtype=torch.int32, device=dev)
    nk_all = torch.empty((N_CAMS * Fg,), dtype=torch.int32, device=dev); nm_all = torch.empty_like(nk_all)
    my_frames = Fg // gsize
    f0 = grank * my_frames
    pidx = torch.empty((my_frames, 6, cap, 2), dtype=torch.int32, device=dev); pdist = torch.empty_like(pidx)
    pgood = torch.empty((my_frames, 6, cap), dtype=torch.uint8, device=dev)
    d_mine = d_all[grank * IMGS:(grank + 1) * IMGS]
    L = api.lib()

    def step(imgs):
        s = torch.cuda.current_stream().cuda_stream
        orb.extract_batch_dev(imgs.data_ptr(), IMGS, SIZE * SIZE, SIZE, k0.data_ptr(), d0.data_ptr(), cap, cnt[0].data_ptr(), s)
        api._check(L.vieo_lapping_split_dev(k0.data_ptr(), d0.data_ptr(), cnt[0].data_ptr(), IMGS, cap, lap.data_ptr(), k1.data_ptr(),
                                            d_mine.data_ptr(), cnt[1].data_ptr(), s))
        if gsize > 1:  # the one exchange step of this path: lapping-ordered descriptors + counts of the group's cameras
            dist.all_gather_into_tensor(d_all, d_mine, group=group)
            dist.all_gather_into_tensor(cnt_all, cnt, group=group)
            nk_all.copy_(cnt_all[:, 0].reshape(-1)); nm_all.copy_(cnt_all[:, 1].reshape(-1))
        else:
            nk_all.copy_(cnt[0]); nm_all.copy_(cnt[1])
        # [camera][frame] layout: frame stride 1, camera stride Fg; this rank's frames [f0, f0 + my_frames)
        api._check(L.vieo_fisheye_knn_dev(d_all.data_ptr() + f0 * cap * 32, nk_all.data_ptr() + 4 * f0, nm_all.data_ptr() + 4 * f0,
                                          N_CAMS, my_frames, cap, 1, Fg, pidx.data_ptr(), pdist.data_ptr(), pgood.data_ptr(), s))

    for i in range(args.warmup):
        step(dev_imgs[i % pool])
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    orb.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(dev_imgs[(args.warmup + i) % pool])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    stage_ms, ncalls = orb.profile_read()
    orb.profile(False)
    if dist:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    frames_total = (world // gsize) * Fg   # rig frames per step over the whole job
    value = frames_total * args.steps / (ms_max / 1e3)
    n_good = int(pgood.sum().item())

    # e2e: pinned host images in, this rank's results out
    stage = torch.empty_like(dev_imgs[0])
    h_k = torch.empty(k1.shape, dtype=k1.dtype).pin_memory(); h_d = torch.empty((IMGS, cap, 32), dtype=torch.uint8).pin_memory()
    h_cnt = torch.empty(cnt.shape, dtype=cnt.dtype).pin_memory()
    h_pi = torch.empty(pidx.shape, dtype=pidx.dtype).pin_memory(); h_pd = torch.empty(pdist.shape, dtype=pdist.dtype).pin_memory()
    h_pg = torch.empty(pgood.shape, dtype=pgood.dtype).pin_memory()

    def e2e_step(i):
        stage.copy_(host_t[i % pool], non_blocking=True)
        step(stage)
        h_k.copy_(k1, non_blocking=True); h_d.copy_(d_mine, non_blocking=True); h_cnt.copy_(cnt, non_blocking=True)
        h_pi.copy_(pidx, non_blocking=True); h_pd.copy_(pdist, non_blocking=True); h_pg.copy_(pgood, non_blocking=True)
        torch.cuda.synchronize()
        return int(h_cnt[0, 0]) + int(h_pg[0, 0, 0])
    for i in range(3):
        e2e_step(i)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(3 + i)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = frames_total * args.steps / float(t.item())
    h2d = IMGS * SIZE * SIZE
    d2h = sum(x.numel() * x.element_size() for x in (h_k, h_d, h_cnt, h_pi, h_pd, h_pg))
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    orb_ms = sum(stage_ms.values()) / max(ncalls, 1)
    alg = TUMVI_BYTES_PER_IMAGE * IMGS
    achieved = alg / (orb_ms / 1e3) / 1e9 if orb_ms > 0 else 0.0
    cpu_frames = 6
    cpu_v, cpu_dt, cores = cpu_rig(rig_images(cpu_frames, 411), cpu_frames)
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload(IMGS), "images_per_step_per_gpu": IMGS, "rig_frames_per_step": frames_total,
                       "cameras_per_gpu": cams_here, "ranks_per_rig_group": gsize,
                       "exchange": "none (all 4 cameras on the GPU)" if gsize == 1 else
                       f"one NCCL all_gather of {IMGS * cap * 32 / 1e6:.1f} MB descriptors + counts per rank per step inside the rig group",
                       "cache": f"inputs rotate over {pool} batches", "accepted_pair_matches_last_step_rank0": n_good},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": (orb.last_launches() + 2) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "ORB extractor (k_resize x7 + k_fast_cells + k_quadtree + k_orient_desc)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "alg_bytes_per_launch": alg, "launch_ms": orb_ms,
                         "stage_ms_per_step": {k: v / max(ncalls, 1) for k, v in stage_ms.items()},
                         "note": "SURVEY 8(d): 871,960 algorithmic bytes per 512x512 1000-feature image over the summed extractor stage time"},
            "cpu_baseline": {"value": cpu_v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_frames} rig frames, one extraction thread per camera + the 6 pair matches"}}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()

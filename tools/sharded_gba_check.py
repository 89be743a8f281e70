"""torchrun --nproc-per-node N tools/sharded_gba_check.py [n_kf n_points iters] — BASELINE configs[4]: the final global BA
with its landmarks (Schur blocks) partitioned over N B200s, ONE all-reduce of the reduced camera system per LM trial over
NCCL / NVLink, against the single-GPU run (states must agree far inside the 1e-6 chi2 budget)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vieo_slam_b200.api as api  # noqa: E402
from vieo_slam_b200 import sharding, synth  # noqa: E402


def problem(n_kf, n_pt, seed, pre_gpu):
    s = synth.vio_sequence(40 + seed, 4 * n_kf + 1, speed=1.0, rot=0.6)
    kf = list(range(0, 4 * n_kf, 4))
    imu, t = s["imu"], s["times"]
    seg, smp, tt, bb = [0], [], [], []
    for k in range(1, len(kf)):
        lo = max(np.searchsorted(imu[:, 0], t[kf[k - 1]], "right") - 1, 0)
        hi = min(np.searchsorted(imu[:, 0], t[kf[k]], "left") + 1, len(imu))
        smp.append(imu[lo:hi]); seg.append(seg[-1] + hi - lo); tt.append((t[kf[k - 1]], t[kf[k]]))
        bb.append(np.r_[s["truth"][kf[k - 1]]["bg"], s["truth"][kf[k - 1]]["ba"]])
    pre = pre_gpu.preintegrate_batch(np.vstack(smp), np.asarray(seg, np.int32), np.asarray(tt), np.asarray(bb))
    pre = np.concatenate([pre[:1], pre])
    cam = synth.euroc_camera()
    return cam, synth.make_gba_problem(s, pre, kf, cam, n_points=n_pt, seed=seed)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    n_pt = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cam, d = problem(n_kf, n_pt, 8, api.IMUPreintegrator(device=local))
    part = sharding.shard_lba_problem(d, rank, world)
    caps = dict(max_states=n_kf + 8, max_points=len(d["points"]) + 8, max_edges=len(d["edge_state"]) + 8, max_imu=n_kf + 8)
    ba = api.BundleAdjuster(device=local, global_ba=True, **caps)
    comm = sharding.make_comm(rank, world, local)
    if os.environ.get("VIEO_SHARD_CALLBACK") == "1":
        sharding.install_allreduce(ba, rank, world)
    else:
        sharding.install_comm(ba, comm)
    ba.GlobalBundleAdjustmentNavStatePRV(part, cam, nIterations=2, bRobust=False)  # warm-up (NCCL channels, allocations)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    out = ba.GlobalBundleAdjustmentNavStatePRV(part, cam, nIterations=iters, bRobust=False)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        single = api.BundleAdjuster(device=local, global_ba=True, **caps)
        single.GlobalBundleAdjustmentNavStatePRV(d, cam, nIterations=2, bRobust=False)
        t1 = time.perf_counter()
        ref = single.GlobalBundleAdjustmentNavStatePRV(d, cam, nIterations=iters, bRobust=False)
        t1 = time.perf_counter() - t1
        dp = max(np.abs(out["states"][f] - ref["states"][f]).max() for f in ("p", "q", "v", "dbg", "dba"))
        rel = abs(out["res"]["err_end"] - ref["res"]["err_end"]) / ref["res"]["err_end"]
        ok = dp < 1e-7 and rel < 1e-6 and out["iterations"] == ref["iterations"]
        print(f"GBA {n_kf} KFs / {len(d['points'])} points / {len(d['edge_state'])} obs, {out['iterations']} LM iterations: "
              f"world={world} {dt.item() * 1e3:.1f} ms (max over ranks) vs single GPU {t1 * 1e3:.1f} ms; max state diff {dp:.3e}, "
              f"chi2 {out['res']['err_end']:.8g} vs {ref['res']['err_end']:.8g} (rel {rel:.2e}); "
              f"edges per rank ~{len(part['edge_state'])}, all-reduce {8 * (len(d['states']) * 15) ** 2 / 1e6:.0f} MB per trial")
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1.0:
        print("SHARDED_GBA_OK")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

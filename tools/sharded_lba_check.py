"""torchrun --nproc-per-node N tools/sharded_lba_check.py — landmark-sharded LocalBA over NCCL vs the single-GPU run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vieo_slam_b200.api as api  # noqa: E402
from vieo_slam_b200 import sharding, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seq = synth.vio_sequence(203, 130, speed=1.5, rot=1.0)
    kf = list(range(0, 130, 4))
    pre_gpu = api.IMUPreintegrator(device=local)
    imu, t = seq["imu"], seq["times"]
    seg, smp, tt, bb = [0], [], [], []
    for k in range(1, len(kf)):
        lo = max(np.searchsorted(imu[:, 0], t[kf[k - 1]], "right") - 1, 0)
        hi = min(np.searchsorted(imu[:, 0], t[kf[k]], "left") + 1, len(imu))
        smp.append(imu[lo:hi]); seg.append(seg[-1] + hi - lo); tt.append((t[kf[k - 1]], t[kf[k]]))
        bb.append(np.r_[seq["truth"][kf[k - 1]]["bg"], seq["truth"][kf[k - 1]]["ba"]])
    pre = pre_gpu.preintegrate_batch(np.vstack(smp), np.asarray(seg, np.int32), np.asarray(tt), np.asarray(bb))
    pre = np.concatenate([pre[:1], pre])
    cam = synth.euroc_camera()
    d = synth.make_lba_problem(seq, pre, kf, cam, n_local=10, n_fixed=20, n_points=800, seed=7)
    part = sharding.shard_lba_problem(d, rank, world)
    ba = api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16, device=local)
    sharding.install_allreduce(ba, rank, world)
    out = ba.LocalBundleAdjustmentNavStatePRV(part, cam)
    # gather the points back in global numbering
    pts = torch.zeros((len(d["points"]), 3), dtype=torch.float64, device="cuda")
    pts[torch.from_numpy(part["point_ids"]).cuda()] = torch.from_numpy(out["points"]).cuda()
    dist.all_reduce(pts)
    ok = True
    if rank == 0:
        single = api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16, device=local)
        ref = single.LocalBundleAdjustmentNavStatePRV(d, cam)
        dp = max(np.abs(out["states"][f] - ref["states"][f]).max() for f in ("p", "q", "v", "dbg", "dba"))
        dx = np.abs(pts.cpu().numpy() - ref["points"]).max()
        ok = dp < 1e-7 and dx < 1e-6 and out["res"]["accepted"] == ref["res"]["accepted"]
        print(f"world={world} max state diff {dp:.3e}, max point diff {dx:.3e}, iterations {out['res']['iterations']} vs "
              f"{ref['res']['iterations']}")
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1.0:
        print("SHARDED_LBA_OK")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

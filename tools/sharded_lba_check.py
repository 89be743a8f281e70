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
    comm = sharding.make_comm(rank, world, local)
    if os.environ.get("VIEO_SHARD_CALLBACK") == "1":
        sharding.install_allreduce(ba, rank, world)   # the host-callback form of the same exchange
    else:
        sharding.install_comm(ba, comm)               # ncclAllReduce issued by the library on the handle's stream
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
    # abort with skewed timing (mbAbortBA is set per process): rank 1 sees the flag from the start, rank 0 never does; the
    # flag travels in the all-reduced trial record, so both ranks must leave the call together (no hang) having run the
    # same number of LM iterations
    stop = np.array([1 if rank == world - 1 else 0], np.uint8)
    out2 = ba.LocalBundleAdjustmentNavStatePRV(part, cam, stop=stop)
    its = torch.tensor([float(out2["res"]["iterations"][0]), float(out2["res"]["iterations"][1])], device="cuda")
    lo, hi = its.clone(), its.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all().item()) and float(hi.sum().item()) == 0.0
    # the same inside optimize(): the flag seen by ONE rank rides in the all-reduced trial record (k_ba_control)
    ba.set_problem(part, cam)
    n_it = torch.tensor([float(ba.optimize(4, 1.0, stop=stop))], device="cuda")
    lo2, hi2 = n_it.clone(), n_it.clone()
    dist.all_reduce(lo2, op=dist.ReduceOp.MIN); dist.all_reduce(hi2, op=dist.ReduceOp.MAX)
    same = same and lo2.item() == hi2.item() == 0.0
    if rank == 0:
        print(f"skewed abort: iterations {its.tolist()} on every rank: {same}")
    ok = ok and same
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1.0:
        print("SHARDED_LBA_OK")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

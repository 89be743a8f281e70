"""Multi-GPU host logic on CPU (world_size 2, gloo): landmark partition of a LocalBA window and the identity the
sharded engine relies on — the sum over ranks of the partial reduced camera systems [S | bschur | b | chi2] equals the
single-rank system (lambda added to the pose diagonal by rank 0 only, inertial edges on rank 0 only)."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from vieo_slam_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    seq = synth.vio_sequence(5, 60)
    kf = list(range(0, 60, 3))
    pre = O.imu_preintegrate_frames(seq, kf, O.imu_noise())
    cam = synth.euroc_camera()
    return cam, synth.make_lba_problem(seq, pre, kf, cam, n_local=8, n_fixed=6, n_points=200, seed=4)


def test_shard_range_and_partition():
    for n, w in ((10, 3), (64, 8), (5, 8), (0, 2)):
        parts = [sharding.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    cam, d = _problem()
    owner = sharding.partition_landmarks(d["edge_point"], len(d["points"]), 4)
    k = np.bincount(d["edge_point"], minlength=len(d["points"])).astype(np.int64)
    loads = np.array([(k[owner == r] ** 2).sum() for r in range(4)])
    assert owner.min() == 0 and owner.max() == 3 and loads.max() <= 1.1 * loads.mean()
    parts = [sharding.shard_lba_problem(d, r, 4) for r in range(4)]
    assert sum(len(p["points"]) for p in parts) == len(d["points"])
    assert sorted(np.concatenate([p["edge_ids"] for p in parts]).tolist()) == list(range(len(d["edge_state"])))
    for r, p in enumerate(parts):
        # the inertial topology travels to every rank (the library evaluates those edges on rank 0 only, but every rank
        # eliminates the same V / Bias chain after the all-reduce of the global BA)
        assert np.all(np.diff(p["edge_point"]) >= 0) and np.array_equal(p["imu_i"], d["imu_i"]) and np.array_equal(p["imu_j"], d["imu_j"])
        assert np.array_equal(p["points"], d["points"][p["point_ids"]])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cam, d = _problem()
        part = sharding.shard_lba_problem(d, rank, world)
        lam = 1.0
        S, bs, b, chi2 = O.ba_debug_system(sharding.evaluated_edges(part, rank), cam, lam, lambda_on_poses=(rank == 0))
        buf = torch.from_numpy(np.concatenate([S.ravel(), bs, b, [chi2]]))
        dist.all_reduce(buf)  # the ONE collective per LM trial
        if rank == 0:
            Sf, bsf, bf, chif = O.ba_debug_system(d, cam, lam)
            n = len(bsf)
            got = buf.numpy()
            ok = (np.allclose(got[:n * n].reshape(n, n), Sf, rtol=1e-9, atol=1e-9 * np.abs(Sf).max())
                  and np.allclose(got[n * n:n * n + n], bsf, rtol=1e-9, atol=1e-9 * np.abs(bsf).max())
                  and np.allclose(got[n * n + n:n * n + 2 * n], bf, rtol=1e-9, atol=1e-9 * np.abs(bf).max())
                  and abs(got[-1] - chif) <= 1e-9 * chif)
            # solving the summed system gives the single-rank pose step
            xp, _, _ = O.ba_debug_step(d, cam, lam)
            x = np.linalg.solve(got[:n * n].reshape(n, n), got[n * n:n * n + n])
            ok = ok and np.allclose(x, xp, rtol=1e-7, atol=1e-10)
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_partial_reduced_systems_sum_to_the_full_one_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.gpu
def test_sharded_local_ba_matches_single_gpu():
    """2 ranks / 2 GPUs over NCCL: the landmark-sharded engine reproduces the single-GPU result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(ROOT, "tools", "sharded_lba_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARDED_LBA_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_sharded_global_ba_matches_single_gpu():
    """2 ranks / 2 GPUs over NCCL: the landmark-sharded global BA (V / Bias chain eliminated on every rank after the
    all-reduce) reproduces the single-GPU states to 1e-7 and chi2 to 1e-6."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633",
                          os.path.join(ROOT, "tools", "sharded_gba_check.py"), "60", "3000", "6"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARDED_GBA_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]

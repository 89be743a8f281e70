"""Host-side sharding helpers.  Units that shard without any exchange (cameras, frames,
PoseOptimization problems) are split by `shard_range`; LocalBA is landmark-partitioned: every rank keeps all keyframe
states, a subset of the map points with ALL their edges, and only rank 0 evaluates the inertial edges; the reduced camera
system is summed with one all-reduce per LM trial through the callback installed by `install_allreduce`."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) of n independent units for `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_landmarks(edge_point, n_points, world):
    """Greedy balance of the Schur work: point l costs k_l^2 (k_l = its number of edges).  Returns owner[P] (int32).
    Deterministic (ties -> lowest rank), every point gets exactly one owner."""
    k = np.bincount(np.asarray(edge_point), minlength=n_points).astype(np.int64)
    cost = k * k
    owner = np.empty(n_points, np.int32)
    load = np.zeros(world, np.int64)
    for p in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += cost[p]
    return owner


def shard_lba_problem(d, rank, world):
    """The part of LocalBA problem `d` (synth.make_lba_problem layout) that `rank` works on.  Points are renumbered
    densely; `point_ids` maps local -> global point index for the write-back."""
    owner = partition_landmarks(d["edge_point"], len(d["points"]), world)
    mine = np.nonzero(owner == rank)[0]
    remap = -np.ones(len(d["points"]), np.int64)
    remap[mine] = np.arange(len(mine))
    sel = owner[d["edge_point"]] == rank
    out = dict(d)
    out["points"] = d["points"][mine]
    for k in ("edge_state", "obs", "inv_sigma2", "edge_flags"):
        out[k] = d[k][sel]
    out["edge_point"] = remap[d["edge_point"][sel]].astype(np.int32)  # stays sorted: remap is monotone on `mine`
    # The inertial / bias-walk edges are EVALUATED on rank 0 only (the library drops them on the other ranks of a sharded
    # handle), but their arrays travel to every rank: the global BA derives the V / Bias chain it eliminates before the dense
    # factorisation from this topology, and every rank must eliminate the same chain after the all-reduce.
    out["point_ids"] = mine
    out["edge_ids"] = np.nonzero(sel)[0]
    return out


def evaluated_edges(part, rank):
    """The edges a rank actually evaluates: its visual edges, plus the inertial / bias-walk edges on rank 0 only — what the
    library does with a sharded handle, for code (the CPU oracle in the tests) that has no notion of ranks."""
    if rank == 0:
        return part
    out = dict(part)
    for k in ("imu_i", "imu_j", "preint", "imu_dt_kf"):
        out[k] = part[k][:0]
    return out


def install_allreduce(ba, rank, world, group=None):
    """Wire a BundleAdjuster to torch.distributed: the library calls back with (device pointer, count, stream) and the
    sum over ranks runs on that stream (NCCL over NVLink).  The tensor wrapping the library's buffer is created once
    per (pointer, count)."""
    import torch
    import torch.distributed as dist
    cache = {}

    def allreduce(ptr, count, stream):
        key = (ptr, count)
        if key not in cache:
            cache[key] = _tensor_from_ptr(ptr, count)
        t = cache[key]
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    ba.set_sharding(rank, world, allreduce)
    return allreduce


def make_comm(rank, world, device, group=None):
    """Library-owned NCCL communicator for a torch.distributed job: rank 0 creates the unique id, one broadcast hands it
    to the others (the only use of torch.distributed: rendez-vous, never the data path)."""
    import torch
    import torch.distributed as dist
    from . import api
    uid = torch.from_numpy(api.comm_unique_id() if rank == 0 else np.zeros(api.COMM_ID_BYTES, np.uint8))
    if dist.get_backend(group) == "nccl":
        uid = uid.cuda(device)
    dist.broadcast(uid, src=0, group=group)
    return api.Comm(uid.cpu().numpy(), rank, world, device)


def install_comm(ba, comm):
    """Sharded BundleAdjuster with the all-reduce issued by the library (ncclAllReduce on the handle's stream)."""
    ba.set_comm(comm)
    return comm


def _tensor_from_ptr(ptr, count):
    """Zero-copy float64 CUDA tensor over memory owned by the C library (via __cuda_array_interface__)."""
    import torch

    class _Buf:
        pass
    b = _Buf()
    b.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(b, device="cuda")

"""Time the batched PoseOptimization kernel (device-resident, CUDA events)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import vieo_slam_b200.api as api

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pre = api.IMUPreintegrator()
trk = bench.make_tracking_inputs(505, F, pre.preintegrate_batch)
dev = torch.device("cuda")
to = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
d = [to(trk[k]) for k in ("pbs",)] + [to(np.asarray(trk["cam"]).reshape(1))] + [to(trk[k]) for k in ("Xw", "obs", "w", "flags")]
n, E = len(trk["pbs"]), len(trk["flags"])
res = torch.empty(n * api.POSEOPT_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
outl = torch.empty(E, dtype=torch.uint8, device=dev); chi = torch.empty(E, dtype=torch.float64, device=dev)
s = torch.cuda.current_stream().cuda_stream
run = lambda: api.Optimizer.pose_opt_batch_dev(d[0].data_ptr(), n, d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                               d[5].data_ptr(), res.data_ptr(), outl.data_ptr(), chi.data_ptr(), s)
run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
r = np.frombuffer(res.cpu().numpy().tobytes(), api.POSEOPT_RESULT_DTYPE)
print(f"pose_opt: {n} problems, {a.elapsed_time(b)/5:.3f} ms per batch, mean LM iterations {r['iterations'].mean():.1f}, inliers {r['n_inliers'].mean():.0f}")
import ctypes as C
L = api.lib()
if hasattr(L, "vieo_debug_po_prof"):
    L.vieo_debug_po_prof(None, 1)
    run(); torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    L.vieo_debug_po_prof(buf, 0)
    v = list(buf)
    calls = max(v[7], 1)
    names = ["visual(thread0)", "imu(w7l0)", "bias+prior(w6l0)", "wait+reduce", "Oe/chi2/AtO", "H entries", "evaluate total", "calls",
             "chol+solves", "oplus+campose", "prologue", "reclassify", "marginalise", "kernel total"]
    print("block 0, cycles per evaluate call (calls = %d):" % calls)
    for k, nm in enumerate(names):
        if k != 7: print(f"  {nm:20s} {v[k] / calls:10.0f}   total {v[k]:12d}")

"""Wall-clock breakdown of the e2e step: each host thread group alone and together."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
import bench
from bench import EUROC, W, H, LBA_EVERY
import vieo_slam_b200.api as api
from vieo_slam_b200.synth import stereo_stream

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pre = api.IMUPreintegrator()
host = stereo_stream(F * 2, 505, dark_every=16).reshape(2, F, 2, H, W)
host_t = torch.from_numpy(host).pin_memory(); host_np = host_t.numpy()
trk = bench.make_tracking_inputs(505, F, pre.preintegrate_batch)
lbas = bench.make_lba_windows(203, 2, pre.preintegrate_batch)
fe = api.StereoFrontend(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H, max_frames=F)
outs = fe.alloc_outputs(F, pinned=True)
n_lba = F // LBA_EVERY
bas = [api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16) for _ in range(NW)]
pool = ThreadPoolExecutor(NW); tpool = ThreadPoolExecutor(1)

def lba_job(wk):
    for i in range(wk, n_lba, NW):
        bas[wk].LocalBundleAdjustmentNavStatePRV(lbas[i % 2], trk["cam"])
def tracking():
    pre.preintegrate_batch(*trk["imu"])
    api.Optimizer.PoseOptimizationBatch(trk["pbs"], trk["cam"], trk["Xw"], trk["obs"], trk["w"], trk["flags"])
def frontend(i):
    fe.process(host_np[i % 2], outs)
def timeit(fn, reps=6):
    fn(0); torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(reps): fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3
print("frontend alone ms", timeit(frontend))
print("tracking alone ms", timeit(lambda i: tracking()))
print("lba alone ms", timeit(lambda i: [f.result() for f in [pool.submit(lba_job, w) for w in range(NW)]]))
def both(i):
    ft = tpool.submit(tracking); frontend(i); ft.result()
print("frontend+tracking ms", timeit(both))
def all3(i):
    fs = [pool.submit(lba_job, w) for w in range(NW)]; ft = tpool.submit(tracking); frontend(i); ft.result(); [f.result() for f in fs]
print("all ms", timeit(all3), "->", F / timeit(all3) * 1e3, "frames/s")

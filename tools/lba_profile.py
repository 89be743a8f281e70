"""Time one LocalBA window (host wall clock) — run plain or under ncu for the launch list."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import vieo_slam_b200.api as api
from vieo_slam_b200 import synth

pre = api.IMUPreintegrator()
lbas = bench.make_lba_windows(203, 1, pre.preintegrate_batch)
cam = synth.euroc_camera()
ba = api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16)
ba.LocalBundleAdjustmentNavStatePRV(lbas[0], cam)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t = time.perf_counter()
for _ in range(reps):
    out = ba.LocalBundleAdjustmentNavStatePRV(lbas[0], cam)
print("LBA ms", (time.perf_counter() - t) / reps * 1e3, out["res"], "launches", ba.last_launches())
t = time.perf_counter()
for _ in range(reps):
    ba.set_problem(lbas[0], cam)
print("set_problem ms", (time.perf_counter() - t) / reps * 1e3)
t = time.perf_counter()
for _ in range(reps):
    pb, keep = api.ba_problem(lbas[0])
print("python marshalling ms", (time.perf_counter() - t) / reps * 1e3)

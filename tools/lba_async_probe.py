"""Host cost of the asynchronous LocalBA path: time per vieo_local_ba_prv_begin (host enqueue) and the wall / device time of a
batch of 16 windows driven by one thread, alone on the device."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import vieo_slam_b200.api as api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pre = api.IMUPreintegrator()
lbas = bench.make_lba_windows(203, 3, pre.preintegrate_batch)
cam = __import__("vieo_slam_b200.synth", fromlist=["x"]).euroc_camera()
bas = [api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16) for _ in range(n)]
for rep in range(3):
    tb = []
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        bas[i].begin(lbas[i % 3], cam)
        tb.append(time.perf_counter() - a)
    t1 = time.perf_counter()
    outs = [bas[i].end() for i in range(n)]
    t2 = time.perf_counter()
    print(f"rep {rep}: begin {1e3 * np.mean(tb):.3f} ms each (max {1e3 * max(tb):.3f}), all begins {1e3 * (t1 - t0):.2f} ms, "
          f"ends {1e3 * (t2 - t1):.2f} ms, total {1e3 * (t2 - t0):.2f} ms; device ms per window {np.mean([b.last_ms() for b in bas]):.2f}, "
          f"launches {bas[0].last_launches()}, iterations {outs[0]['res']['iterations']}")
# python-side share of begin: problem struct packing only
t0 = time.perf_counter()
for i in range(200):
    api.ba_problem(lbas[i % 3])
print(f"ba_problem() packing: {1e3 * (time.perf_counter() - t0) / 200:.3f} ms")
t0 = time.perf_counter()
out = bas[0].LocalBundleAdjustmentNavStatePRV(lbas[0], cam)
print(f"one synchronous window alone: {1e3 * (time.perf_counter() - t0):.2f} ms wall, device {bas[0].last_ms():.2f} ms")

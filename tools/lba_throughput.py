"""LocalBA windows/s with W concurrent host threads (one handle + stream each)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ThreadPoolExecutor
import bench
import vieo_slam_b200.api as api
from vieo_slam_b200 import synth
pre = api.IMUPreintegrator()
lbas = bench.make_lba_windows(203, 2, pre.preintegrate_batch)
cam = synth.euroc_camera()
for W in (1, 2, 4, 8, 16):
    bas = [api.BundleAdjuster(max_states=64, max_points=2048, max_edges=16384, max_imu=16) for _ in range(W)]
    def job(w):
        for i in range(6):
            bas[w].LocalBundleAdjustmentNavStatePRV(lbas[i % 2], cam)
    with ThreadPoolExecutor(W) as ex:
        list(ex.map(job, range(W)))
        t = time.perf_counter()
        list(ex.map(job, range(W)))
        dt = time.perf_counter() - t
    print(f"workers {W:2d}: {6 * W / dt:7.1f} windows/s ({dt / 6 * 1e3:.2f} ms per window per worker)")

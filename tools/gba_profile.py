"""GlobalBundleAdjustmentNavStatePRV at BASELINE configs[4] size on one B200: wall time, LM iterations, launches; and the
CPU oracle beside it at the largest size its scalar dense Cholesky finishes in seconds.
  python tools/gba_profile.py [n_kf] [n_points] [iters]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import vieo_slam_b200.api as api  # noqa: E402
from vieo_slam_b200 import synth  # noqa: E402

n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n_pt = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
with_oracle = os.environ.get("GBA_ORACLE", "1") == "1"
pre_gpu = api.IMUPreintegrator()


def problem(n_kf, n_pt, seed):
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


cam, d = problem(n_kf, n_pt, 8)
ba = api.BundleAdjuster(max_states=max(64, n_kf + 8), max_points=len(d["points"]) + 8, max_edges=len(d["edge_state"]) + 8,
                        max_imu=n_kf + 8, global_ba=True)
ba.GlobalBundleAdjustmentNavStatePRV(d, cam, nIterations=2, bRobust=False)   # warm-up
for robust in (False, True):
    t0 = time.perf_counter()
    out = ba.GlobalBundleAdjustmentNavStatePRV(d, cam, nIterations=iters, bRobust=robust)
    dt = time.perf_counter() - t0
    print(f"GPU  {n_kf} KFs / {len(d['points'])} points / {len(d['edge_state'])} obs, np = {15 * n_kf - 15}, robust={int(robust)}: "
          f"{dt * 1e3:.1f} ms, {out['iterations']} LM iterations, {ba.last_launches()} launches, "
          f"chi2 {out['res']['err0']:.4g} -> {out['res']['err_end']:.6g}", flush=True)
if with_oracle:
    import oracle_lib as O
    n2 = 100
    cam2, d2 = problem(n2, 6000, 9)
    t0 = time.perf_counter()
    g = ba.GlobalBundleAdjustmentNavStatePRV(d2, cam2, nIterations=10, bRobust=False)
    tg = time.perf_counter() - t0
    t0 = time.perf_counter()
    o = O.global_ba_prv(d2, cam2, n_iterations=10, robust=False)
    to = time.perf_counter() - t0
    print(f"{n2} KFs / {len(d2['points'])} points / {len(d2['edge_state'])} obs: GPU {tg * 1e3:.1f} ms ({g['iterations']} it, chi2 "
          f"{g['res']['err_end']:.8g}), CPU oracle {to * 1e3:.0f} ms ({o['iterations']} it, chi2 {o['res']['err_end']:.8g}), "
          f"rel diff {abs(g['res']['err_end'] - o['res']['err_end']) / o['res']['err_end']:.2e}")

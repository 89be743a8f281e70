"""Essential-graph optimisation (Optimizer::OptimizeEssentialGraph) timing: device engine through the host-buffer C ABI against the
CPU oracle on the same graphs.  Development tool (imports the oracle as the CPU comparison)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import vieo_slam_b200.api as api  # noqa: E402
from vieo_slam_b200 import synth  # noqa: E402

for K in (100, 400, 1000):
    pb = synth.make_essential_graph(K=K, seed=K, fix_scale=True, odom_info_every=7)
    api.Optimizer.OptimizeEssentialGraph(pb)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); out, T, st = api.Optimizer.OptimizeEssentialGraph(pb); ts.append(time.perf_counter() - t)
    t = time.perf_counter(); ro, rs = O.essential_graph(pb); tc = time.perf_counter() - t
    print(f"K={K} edges={len(pb['ei'])} gpu_ms={min(ts) * 1e3:.2f} cpu_oracle_ms={tc * 1e3:.1f} iterations={int(st['iterations'])}/{rs['iterations']} "
          f"trials={int(st['trials'])}/{rs['trials']} chi2={st['chi2_final']:.6e}/{rs['chi2_final']:.6e} "
          f"max|dt|={np.abs(out['t'] - ro['t']).max():.2e}", flush=True)

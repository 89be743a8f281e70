import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import EUROC, W, H
import vieo_slam_b200.api as api
from vieo_slam_b200.synth import stereo_stream
host = torch.from_numpy(stereo_stream(64, 505, dark_every=16).reshape(64, 2, H, W)).pin_memory()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for ch in (1, 2, 4, 8):
    os.environ["VIEO_FE_CHUNKS"] = str(ch)
    fe = api.StereoFrontend(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H, max_frames=64)
    outs = fe.alloc_outputs(64, pinned=True)
    h = host.numpy()
    print(ch, "chunks: process(64 frames) %.3f ms" % t(lambda: fe.process(h, outs)))

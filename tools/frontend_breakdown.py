"""Where the host-buffer front-end call spends its time: copies vs kernels (64 stereo frames)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import EUROC, W, H
import vieo_slam_b200.api as api
from vieo_slam_b200.synth import stereo_stream
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
host = torch.from_numpy(stereo_stream(F, 505, dark_every=16).reshape(F, 2, H, W)).pin_memory()
dev = torch.empty_like(host, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D %.1f MB: %.3f ms" % (host.numel() / 1e6, t(lambda: dev.copy_(host, non_blocking=True))))
fe = api.StereoFrontend(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H, max_frames=F)
outs = fe.alloc_outputs(F, pinned=True)
print("frontend process (host buffers): %.3f ms" % t(lambda: fe.process(host.numpy(), outs)))
n_out = sum(o.nbytes for o in outs)
big = torch.empty(n_out, dtype=torch.uint8, device="cuda"); hb = torch.empty(n_out, dtype=torch.uint8).pin_memory()
print("D2H %.1f MB: %.3f ms" % (n_out / 1e6, t(lambda: hb.copy_(big, non_blocking=True))))
orb = api.ORBextractor(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H, max_batch=2 * F)
cap = orb.cap
kps = torch.empty((2 * F, cap, 6), dtype=torch.float32, device="cuda"); desc = torch.empty((2 * F, cap, 32), dtype=torch.uint8, device="cuda")
nkp = torch.empty((2 * F,), dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
print("extract kernels (device-resident, one batch of %d images): %.3f ms" % (2 * F, t(lambda: orb.extract_batch_dev(dev.data_ptr(), 2 * F, H * W, W, kps.data_ptr(), desc.data_ptr(), cap, nkp.data_ptr(), s))))
for nb in (8, 16, 32):
    o2 = api.ORBextractor(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H, max_batch=nb)
    print("  batch of %d images: %.3f ms" % (nb, t(lambda: o2.extract_batch_dev(dev.data_ptr(), nb, H * W, W, kps.data_ptr(), desc.data_ptr(), cap, nkp.data_ptr(), s))))
# per-stage CUDA-event times of the extractor alone on the device (no other stream active)
orb.profile(True)
for _ in range(10):
    orb.extract_batch_dev(dev.data_ptr(), 2 * F, H * W, W, kps.data_ptr(), desc.data_ptr(), cap, nkp.data_ptr(), s)
torch.cuda.synchronize()
stage_ms, ncalls = orb.profile_read()
orb.profile(False)
print("isolated stages, ms per batch of %d images:" % (2 * F), {k: round(v / max(ncalls, 1), 3) for k, v in stage_ms.items()},
      "-> us per image:", {k: round(1e3 * v / max(ncalls, 1) / (2 * F), 2) for k, v in stage_ms.items()})

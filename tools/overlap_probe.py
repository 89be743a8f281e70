"""Does k_pose_opt (one CTA per frame, ~254 registers x 256 threads) keep other kernels off the SMs?
Times, device-resident with CUDA events: (a) the ORB extractor on 2F images alone, (b) PoseOptimization on 2F problems
alone, (c) both enqueued back to back on two streams.  If (c) ~ (a) + (b) the kernels serialise; if
(c) ~ max(a, b) they overlap.  One short gpurun call:  python tools/overlap_probe.py [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import vieo_slam_b200.api as api  # noqa: E402
from vieo_slam_b200.synth import EUROC, stereo_stream  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
W, H = EUROC["w"], EUROC["h"]
n_img = 2 * F
imgs = torch.from_numpy(np.ascontiguousarray(stereo_stream(min(F, 16), 505)).reshape(-1, H, W)).to(dev)
imgs = imgs.repeat((n_img + imgs.shape[0] - 1) // imgs.shape[0], 1, 1)[:n_img].contiguous()
orb = api.ORBextractor(EUROC["nfeatures"], EUROC["scale"], EUROC["nlevels"], EUROC["ini_th"], EUROC["min_th"], W, H,
                       max_batch=n_img)
cap = orb.cap
kps = torch.empty((n_img, cap, 6), dtype=torch.float32, device=dev)
desc = torch.empty((n_img, cap, 32), dtype=torch.uint8, device=dev)
nkp = torch.empty((n_img,), dtype=torch.int32, device=dev)

pre = api.IMUPreintegrator()
trk = bench.make_tracking_inputs(505, F, pre.preintegrate_batch)


def to(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)


d = [to(trk["pbs"]), to(np.asarray(trk["cam"]).reshape(1))] + [to(trk[k]) for k in ("Xw", "obs", "w", "flags")]
n, E = len(trk["pbs"]), len(trk["flags"])
res = torch.empty(n * api.POSEOPT_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
outl = torch.empty(E, dtype=torch.uint8, device=dev)
chi = torch.empty(E, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def fe(s):
    orb.extract_batch_dev(imgs.data_ptr(), n_img, H * W, W, kps.data_ptr(), desc.data_ptr(), cap, nkp.data_ptr(), s.cuda_stream)


def po(s):
    api.Optimizer.pose_opt_batch_dev(d[0].data_ptr(), n, d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                     d[5].data_ptr(), res.data_ptr(), outl.data_ptr(), chi.data_ptr(), s.cuda_stream)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
        for s in (s1, s2):  # the default stream waits for both work streams each repetition
            torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def both():
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    po(s2)
    fe(s1)


def only(fn, s):
    s.wait_stream(torch.cuda.current_stream())
    fn(s)


t_fe = timed(lambda: only(fe, s1))
t_po = timed(lambda: only(po, s2))
t_both = timed(both)
print(f"{F} frames: extractor {t_fe:.3f} ms, pose_opt {t_po:.3f} ms, both on two streams {t_both:.3f} ms "
      f"(sum {t_fe + t_po:.3f}, max {max(t_fe, t_po):.3f}) -> overlap {100 * (t_fe + t_po - t_both) / min(t_fe, t_po):.0f} % of the shorter one")

"""HBM roofline of the geometry kernels (BASELINE configs[3]: 640x480 maps): z-buffer
reprojection (9 B/pixel algorithmic) and dense depth->point-cloud (17 B/pixel).

    python tools/bench_geometry.py [--maps 512] [--iters 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import geometry, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--maps", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--only", default="", help="reproject | depth2pc")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, H, W = a.maps, a.h, a.w
d = synthetic.synthetic_depth_batch(0, 8, H, W)
d = (d * 10).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().to(dev)       # metres
K = torch.tensor(synthetic.synthetic_intrinsics(B, None if (H, W) == (480, 640) else W)).to(dev)
P = torch.tensor(synthetic.synthetic_poses(B)).to(dev)
peak = 6458.4
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        peak = json.load(f)["hbm_gbs"]
except Exception:
    pass


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


px = B * H * W
if a.only != "depth2pc":
    ms = timed(lambda: geometry.reproject_tensor(d, K, P))
    print("reproject_tensor %d maps %dx%d: %.3f ms  %.0f GB/s algorithmic (9 B/px)  %.2f of measured HBM peak %.0f; %.0f maps/s"
          % (B, H, W, ms, px * 9 / ms / 1e6, px * 9 / ms / 1e6 / peak, peak, B / ms * 1e3))
if a.only != "reproject":
    ms = timed(lambda: geometry.depth2pc_tensor(d, K, clip=[0, 10]))
    print("depth2pc_tensor  %d maps %dx%d: %.3f ms  %.0f GB/s algorithmic (17 B/px)  %.2f of measured HBM peak; %.0f maps/s"
          % (B, H, W, ms, px * 17 / ms / 1e6, px * 17 / ms / 1e6 / peak, B / ms * 1e3))

"""Per-layer device time of one network evaluation (CUDA events on the launching stream,
sampled through prg_profile_set / prg_profile_ops).

    python tools/layer_table.py [--batch 32] [--size 256] [--net unet|mask] [--iters 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import _ffi, nets

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--net", default="unet", choices=["unet", "mask"])
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda", 0)
if a.net == "unet":
    net = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).to(dev)
    x = torch.randn(a.batch, 1, a.size, a.size, device=dev)
    t = torch.full((a.batch,), 500, device=dev, dtype=torch.long)
    pc = torch.tensor([[303.9, 304.2, 128.5, 128.0]], device=dev).repeat(a.batch, 1)
    run = lambda: net(x, t, pc)
else:
    net = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).to(dev)
    x = torch.rand(a.batch, 1, a.size, a.size, device=dev)
    run = lambda: net(x)
for _ in range(3):
    run()
torch.cuda.synchronize()
# whole-forward time without per-op events
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
print("forward (no per-op events): %.3f ms / eval at batch %d" % (e0.elapsed_time(e1) / a.iters, a.batch))
_ffi.profile_set(1)
_ffi.profile_ops(reset=True)
_ffi.profile_read(reset=True)
for _ in range(a.iters):
    run()
torch.cuda.synchronize()
rows = _ffi.profile_ops(reset=True)
fam = _ffi.profile_read(reset=True)
_ffi.profile_set(0)
tot = 0.0
print("%-100s %9s %9s" % ("op", "us", "TFLOP/s"))
for lab, n, ms, fl in rows:
    us = ms / n * 1e3
    tot += us
    tf = fl * a.batch / (us * 1e-6) / 1e12 if fl > 0 else 0
    print("%-100s %9.1f %9.0f" % (lab[:100], us, tf))
print("sum of ops: %.3f ms" % (tot / 1e3))
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-18s %8.3f ms  (%d launches / eval)" % (k, v["ms"] / v["forwards"], v["launches"] // v["forwards"]))

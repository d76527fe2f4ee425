"""One warm-up + one profiled U-Net evaluation (for ncu launch lists / --set full captures).

    ncu --profile-from-start off ... python tools/profile_forward.py --batch 32
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import nets

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--net", default="unet", choices=["unet", "mask"])
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
run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

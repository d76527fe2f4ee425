"""Relative L2 error of one U-Net evaluation (256x256, B = 1) against the fp32 oracle, for the
build / environment switches in effect (PRG_NO_XF, PRG_CONV_FLAGS=128 = exact SiLU in the transform warps).

    python tools/unet_error.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import torch_ref as R
from pointreggpt_b200 import nets

torch.manual_seed(0)
net = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
net = net.cuda()
g = torch.Generator().manual_seed(6)
pc = torch.tensor([[303.88547, 304.18253, 128.5, 128.0]])
errs = []
for t in (999, 500, 20):
    x = torch.randn(1, 1, 256, 256, generator=g)
    tt = torch.tensor([t])
    ref = R.unet_forward(sd, x, tt, pc)
    got = net(x.cuda(), tt.cuda(), pc.cuda()).cpu()
    errs.append(((got - ref).norm() / ref.norm()).item())
print("unet 256 rel-l2 at t=999/500/20: %s  [PRG_NO_XF=%s PRG_CONV_FLAGS=%s]"
      % (" ".join("%.3e" % e for e in errs), os.environ.get("PRG_NO_XF"), os.environ.get("PRG_CONV_FLAGS")))

"""Timeline of CTA 0 of the fused q/to_out kernel (PRG_QOUT_TRACE): one U-Net evaluation, the
first C=64 LinearAttention prints its per-tile clock stamps to stderr.  The stamps are compiled in only when
the library was built with PRG_BUILD_DEFINES="-DPRG_QOUT_TRACE_BUILD" python -m pointreggpt_b200.build --force."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import nets

torch.manual_seed(0)
dev = torch.device("cuda", 0)
B = 32
net = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).to(dev)
x = torch.randn(B, 1, 256, 256, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.long)
pc = torch.tensor([[303.9, 304.2, 128.5, 128.0]], device=dev).repeat(B, 1)
net(x, t, pc)
torch.cuda.synchronize()
os.environ["PRG_QOUT_TRACE"] = "1"
net(x, t, pc)
torch.cuda.synchronize()

"""One warm-up + one profiled pass of the geometry kernels on 640x480 maps (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import geometry, synthetic

dev = torch.device("cuda", 0)
B, H, W = 128, 480, 640
d = (synthetic.synthetic_depth_batch(0, 8, H, W) * 10).repeat(B // 8, 1, 1, 1).contiguous().to(dev)
K = torch.tensor(synthetic.synthetic_intrinsics(B, None)).to(dev)
P = torch.tensor(synthetic.synthetic_poses(B)).to(dev)
geometry.reproject_tensor(d, K, P)
geometry.depth2pc_tensor(d, K, clip=[0, 10])
torch.cuda.synchronize()
torch.cuda.profiler.start()
geometry.reproject_tensor(d, K, P)
geometry.depth2pc_tensor(d, K, clip=[0, 10])
torch.cuda.synchronize()
torch.cuda.profiler.stop()

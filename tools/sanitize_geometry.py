"""Small reprojection / unprojection calls for compute-sanitizer (memcheck): several map sizes, ring reuse."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import geometry, synthetic

dev = torch.device("cuda", 0)
for B, H, W in ((48, 256, 256), (6, 480, 640), (5, 33, 47), (3, 121, 127)):
    d = (synthetic.synthetic_depth_batch(0, 8, H, W) * 10).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().to(dev)
    K = torch.tensor(synthetic.synthetic_intrinsics(B, None if (H, W) == (480, 640) else W)).to(dev)
    P = torch.tensor(synthetic.synthetic_poses(B)).to(dev)
    for _ in range(2):
        r, m = geometry.reproject_tensor(d, K, P)
        pc, v = geometry.depth2pc_tensor(d, K, clip=[0, 10])
    torch.cuda.synchronize()
    print(B, H, W, int(m.sum()), int(v.sum()))

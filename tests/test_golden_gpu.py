"""CUDA path vs the golden vectors minted from the unmodified reference."""
import hashlib
import os

import numpy as np
import pytest
import torch

from pointreggpt_b200 import geometry as pg, nets
from pointreggpt_b200 import synthetic as S
from pointreggpt_b200.diffusion import GaussianDiffusion

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["256", "640"])
def test_geometry_kernels_bit_exact_vs_reference(tag):
    geo = np.load(os.path.join(GOLD, "geometry.npz"))
    B, H, W = {"256": (2, 256, 256), "640": (1, 480, 640)}[tag]
    d01 = S.synthetic_depth_batch(40, B, H, W)
    K = torch.tensor(geo["K_" + tag]).cuda()
    P = torch.tensor(geo["P_" + tag]).cuda()
    dm = (d01 * 10).cuda()
    rd, rm = pg.reproject_tensor(dm, K, P)
    assert sha(rd.cpu().numpy()) == str(geo["reproject_depth_sha_" + tag])
    assert sha(rm.cpu().numpy()) == str(geo["reproject_mask_sha_" + tag])
    pc, valid = pg.depth2pc_tensor(dm, K)
    assert sha(pc.cpu().numpy()) == str(geo["depth2pc_pc_sha_" + tag])
    assert sha(valid.cpu().numpy()) == str(geo["depth2pc_valid_sha_" + tag])
    pc64, counts = pg.point_cloud_batch(d01.cuda(), K)
    back64, _ = pg.point_cloud_batch(d01.cuda(), K, pose=P)
    for b in range(B):
        n = int(counts[b])
        assert sha(pc64[b, :n].cpu().numpy()) == str(geo["point_cloud_sha_" + tag][b])
        assert sha(back64[b, :n].cpu().numpy()) == str(geo["point_cloud_back_sha_" + tag][b])
    # Generator.generate's path: float32 cloud -> pose -> z-buffer
    offs = torch.arange(B + 1) * (H * W)
    idx = torch.arange(H * W, device="cuda")[None]
    d, m = pg.pc2depth_ragged(pc64.float().reshape(-1, 3), offs, K, image_size=[H, W],
                              valid=idx < counts[:, None], pose=P)
    assert sha(d.cpu().numpy()) == str(geo["generate_pc2depth_depth_sha_" + tag])
    assert sha(m.cpu().numpy()) == str(geo["generate_pc2depth_mask_sha_" + tag])


def test_networks_and_samplers_vs_reference_outputs():
    net = np.load(os.path.join(GOLD, "networks.npz"))
    torch.manual_seed(0)
    u = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).cuda()
    torch.manual_seed(0)
    m = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).cuda()
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(1, 1, 128, 128, generator=gen)
    x2 = torch.randn(1, 1, 256, 256, generator=gen)
    draws = torch.stack([torch.randn(1, 1, 128, 128, generator=gen) for _ in range(6)])
    pc = torch.tensor([[303.88547, 304.18253, 128.5, 128.0]]).cuda()

    def rel(a, b):
        b = torch.tensor(b)
        return ((a.cpu() - b).norm() / b.norm()).item()

    assert rel(u(x.cuda(), torch.tensor([417]).cuda(), pc), net["unet_128_out"]) <= 1e-3
    assert rel(u(x2.cuda(), torch.tensor([999]).cuda(), pc), net["unet_256_out"]) <= 1e-3
    assert rel(m(S.synthetic_depth_batch(3, 1, 128, 128).cuda()), net["mask_128_out"]) <= 1e-3
    dcond = S.synthetic_depth_batch(9, 1, 128, 128)
    ic = (torch.cat([dcond, (dcond > 0).float()], 1) * 2 - 1).cuda()
    d = GaussianDiffusion(u, image_size=128, timesteps=3, objective="pred_x0", beta_schedule="sigmoid").cuda()
    out = d.sample(param_cond=pc, img_cond=ic, has_refine_step=True, noise=draws[:3].cuda())
    assert rel(out, net["p_sample_out"]) <= 2e-3
    d2 = GaussianDiffusion(u, image_size=128, timesteps=12, sampling_timesteps=3, objective="pred_x0",
                           beta_schedule="sigmoid", ddim_sampling_eta=1.0).cuda()
    out = d2.sample(param_cond=pc, img_cond=ic, has_refine_step=True, noise=draws[:3].cuda())
    assert rel(out, net["ddim_out"]) <= 2e-3

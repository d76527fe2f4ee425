"""Pins the oracle restatements (oracle/) to the golden vectors minted from the unmodified
reference (oracle/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from oracle import torch_ref as R
from pointreggpt_b200 import nets
from pointreggpt_b200 import synthetic as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def geo():
    return np.load(os.path.join(GOLD, "geometry.npz"))


@pytest.fixture(scope="module")
def net():
    return np.load(os.path.join(GOLD, "networks.npz"))


def _geo_inputs(geo, tag):
    B, H, W = {"256": (2, 256, 256), "640": (1, 480, 640)}[tag]
    d01 = S.synthetic_depth_batch(40, B, H, W)
    assert sha(d01.numpy()) == str(geo["depth_sha_" + tag]), "synthetic inputs drifted"
    return d01, geo["K_" + tag], geo["P_" + tag], (B, H, W)


@pytest.mark.parametrize("tag", ["256", "640"])
def test_c_oracle_reproject_and_depth2pc(geo, tag):
    d01, K, P, (B, H, W) = _geo_inputs(geo, tag)
    dm = (d01 * 10).numpy()
    od, om = G.reproject(dm, K, P)
    assert sha(od) == str(geo["reproject_depth_sha_" + tag])
    assert sha(om) == str(geo["reproject_mask_sha_" + tag])
    if tag == "256":
        assert np.array_equal(od, geo["reproject_depth_256"])
        assert np.array_equal(np.packbits(om), geo["reproject_mask_256"])
    pc, valid = G.depth2pc(dm, K)
    assert sha(pc) == str(geo["depth2pc_pc_sha_" + tag])
    assert sha(valid) == str(geo["depth2pc_valid_sha_" + tag])


@pytest.mark.parametrize("tag", ["pose", "fwd"])
def test_c_oracle_occlusion_filter(geo, tag):
    occ = np.load(os.path.join(GOLD, "occlusion.npz"))
    d01, K, P, _ = _geo_inputs(geo, "256")
    if tag == "fwd":
        P = occ["P_fwd"]
    rd, rm = G.reproject((d01 * 10).numpy(), K, P)
    assert sha(rd) == str(occ["in_depth_sha_" + tag])
    fd, fm = G.occlusion_filter(rd, rm)
    assert sha(fd) == str(occ["out_depth_sha_" + tag])
    assert sha(fm) == str(occ["out_mask_sha_" + tag])
    assert int((fd != rd).sum()) == int(occ["changed_" + tag]) > 0
    assert np.array_equal(fd[:, :, 96:160, 96:160], occ["out_crop_" + tag])


@pytest.mark.parametrize("tag", ["256", "640"])
def test_c_oracle_generate_path(geo, tag):
    d01, K, P, (B, H, W) = _geo_inputs(geo, tag)
    pcs64 = G.depth2pc_compact(d01.numpy(), K, None)
    assert [sha(p) for p in pcs64] == list(geo["point_cloud_sha_" + tag])
    back = G.depth2pc_compact(d01.numpy(), K, P)
    assert [sha(p) for p in back] == list(geo["point_cloud_back_sha_" + tag])
    pcs = [p.astype(np.float32) for p in pcs64]
    offs = np.cumsum([0] + [p.shape[0] for p in pcs])
    od, om = G.pc2depth(np.concatenate(pcs), None, offs, K, (H, W), pose=P)
    assert sha(od) == str(geo["generate_pc2depth_depth_sha_" + tag])
    assert sha(om) == str(geo["generate_pc2depth_mask_sha_" + tag])


def test_host_helpers_match_reference(geo):
    from pointreggpt_b200 import geometry as pg
    out = pg.intrinsic_transform(geo["intrinsic_in"], resize=256, centercrop=256)
    assert np.array_equal(out, geo["intrinsic_out"])
    np.random.seed(21)
    assert np.array_equal(pg.random_sample_pose(3), geo["pose_seed21"])
    np.random.seed(22)
    assert np.array_equal(pg.random_sample_intrinsic(5), geo["intrinsic_seed22"])


@pytest.fixture(scope="module")
def weights(net):
    torch.manual_seed(0)
    u = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    torch.manual_seed(0)
    m = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
    usd = {k: v.detach() for k, v in u.state_dict().items()}
    msd = {k: v.detach() for k, v in m.state_dict().items()}
    fp = sum(float(usd[k].double().abs().sum()) for k in sorted(usd))
    assert abs(fp - float(net["unet_fingerprint"])) < 1e-6 * fp, "seeded init differs from the reference's"
    assert np.array_equal(usd["downs.2.1.block2.proj.weight"][3, 5].numpy(), net["unet_probe"])
    fpm = sum(float(msd[k].double().abs().sum()) for k in sorted(msd))
    assert abs(fpm - float(net["mask_fingerprint"])) < 1e-6 * fpm
    return usd, msd


def _net_inputs():
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(1, 1, 128, 128, generator=gen)
    x2 = torch.randn(1, 1, 256, 256, generator=gen)
    draws = [torch.randn(1, 1, 128, 128, generator=gen) for _ in range(6)]
    pc = torch.tensor([[303.88547, 304.18253, 128.5, 128.0]])
    dcond = S.synthetic_depth_batch(9, 1, 128, 128)
    ic = torch.cat([dcond, (dcond > 0).float()], 1) * 2 - 1
    return x, x2, draws, pc, ic


def test_torch_oracle_unet_and_mask(net, weights):
    usd, msd = weights
    x, x2, draws, pc, ic = _net_inputs()
    torch.set_num_threads(8)
    y = R.unet_forward(usd, x, torch.tensor([417]), pc)
    assert np.allclose(y.numpy(), net["unet_128_out"], atol=2e-5, rtol=0)
    m = R.maskunet_forward(msd, S.synthetic_depth_batch(3, 1, 128, 128))
    assert np.allclose(m.numpy(), net["mask_128_out"], atol=2e-6, rtol=0)


def test_torch_oracle_samplers(net, weights):
    usd, _ = weights
    x, x2, draws, pc, ic = _net_inputs()
    out = R.p_sample_loop(usd, R.make_schedule(3), pc, ic, draws, has_refine_step=True)
    assert np.allclose(out.numpy(), net["p_sample_out"], atol=5e-5, rtol=0)
    out = R.ddim_sample(usd, R.make_schedule(12), pc, ic, draws, 3, 1.0, has_refine_step=True)
    assert np.allclose(out.numpy(), net["ddim_out"], atol=5e-5, rtol=0)


def test_schedule_buffers(net):
    sch = R.make_schedule(1000, "sigmoid")
    for k, v in sch.items():
        assert np.array_equal(v.numpy(), net["sched_" + k]), k
    assert [t for t, _ in R.ddim_times(1000, 250)] + [-1] == list(net["ddim_times_1000_250"])


def test_c_oracle_overlap_and_voxel_against_independent_formulations():
    """open3d is not installed, so these two restatements have no reference pin; they are checked
    against independent implementations instead: scipy's KD-tree and the torch-op voxel grid."""
    from scipy.spatial import cKDTree
    from pointreggpt_b200 import cloud
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (6000, 3))
    b = rng.uniform(-0.6, 1.4, (5000, 3))
    for r in (0.0375, 0.1):
        want = sum(1 for x in cKDTree(b).query_ball_point(a, r) if len(x) > 0)
        assert G.overlap_count(a, b, r) == want
    for v in (0.025, 0.2):
        c, k = G.voxel_down_sample(a, v)
        t = cloud.voxel_down_sample_torch(torch.tensor(a), v).numpy()   # the torch-op cross-check (CPU)
        assert c.shape == t.shape and np.abs(c - t).max() < 1e-12 and np.all(np.diff(k) > 0)
    r1, r2 = G.compute_overlap_ratio(a, b)
    assert 0 < r1 < 1 and 0 < r2 < 1

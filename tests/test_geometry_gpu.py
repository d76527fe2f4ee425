"""GPU parity (through the C ABI) of the geometry kernels against the C oracle: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from pointreggpt_b200 import geometry as pg
from pointreggpt_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _inputs(B, H, W, seed):
    d01 = S.synthetic_depth_batch(100 + seed, B, H, W)
    K = S.synthetic_intrinsics(B, 256 if H == 256 else None, seed=seed)
    if (H, W) not in ((256, 256), (480, 640)):      # tiny maps: principal point at the centre
        K = K.copy()
        K[:, 0, 0] = K[:, 1, 1] = 1.2 * W
        K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    P = S.synthetic_poses(B, seed=seed + 1)
    return d01, K, P


@pytest.mark.parametrize("shape", [(4, 256, 256), (3, 480, 640), (2, 33, 47), (1, 8, 8)])
def test_reproject_bit_exact(shape):
    B, H, W = shape
    d01, K, P = _inputs(B, H, W, 3)
    dm = d01 * 10
    od, om = G.reproject(dm.numpy(), K, P)
    gd, gm = pg.reproject_tensor(dm.cuda(), torch.tensor(K).cuda(), torch.tensor(P).cuda())
    assert gd.dtype == torch.float32 and gm.dtype == torch.bool
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gm.cpu().numpy(), om)
    assert om.any()


def test_reproject_identity_and_empty():
    d01, K, _ = _inputs(2, 256, 256, 5)
    dm = (d01 * 10).cuda()
    eye = torch.eye(4)[None].repeat(2, 1, 1).cuda()
    gd, gm = pg.reproject_tensor(dm, torch.tensor(K).cuda(), eye)
    od, om = G.reproject(dm.cpu().numpy(), K, eye.cpu().numpy())
    assert np.array_equal(gd.cpu().numpy(), od) and np.array_equal(gm.cpu().numpy(), om)
    # all-invalid input -> empty output
    z = torch.zeros_like(dm)
    gd, gm = pg.reproject_tensor(z, torch.tensor(K).cuda(), eye)
    assert not gm.any() and (gd == 0).all()
    # B = 0
    gd, gm = pg.reproject_tensor(dm[:0], torch.tensor(K[:0]).cuda(), eye[:0])
    assert gd.shape == (0, 1, 256, 256)


@pytest.mark.parametrize("shape", [(3, 256, 256), (2, 480, 640), (1, 5, 7)])
def test_depth2pc_bit_exact(shape):
    B, H, W = shape
    d01, K, _ = _inputs(B, H, W, 7)
    dm = d01 * 10
    for clip, inv in [([0, 10], None), ([0.5, 10], 0.0), (None, None)]:
        opc, ov = G.depth2pc(dm.numpy(), K, clip=clip, invalid=float("nan") if inv is None else inv)
        gpc, gv = pg.depth2pc_tensor(dm.cuda(), torch.tensor(K).cuda(), clip=clip, invalid_num=inv)
        assert np.array_equal(gpc.cpu().numpy().view(np.uint32), opc.view(np.uint32))
        assert np.array_equal(gv.cpu().numpy(), ov)


def _same_bits_or_nan(a, b):
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def test_extreme_depths_bit_exact():
    """Signed zeros, negatives, denormals, huge values, inf and NaN: the fast exact division in the
    kernels must hand these to the IEEE division and agree with the oracle bit for bit."""
    B, H, W = 2, 64, 96
    d01, K, P = _inputs(B, H, W, 13)
    dm = (d01 * 10).numpy().copy()
    special = np.array([0.0, -0.0, -1.5, 1e-42, 1e-30, 1e-12, 1e12, 1e30, 3e38, np.inf, -np.inf, np.nan,
                        1.0, 2.5e-7, 7.7e19, 65504.0], np.float32)
    flat = dm.reshape(B, -1)
    flat[:, ::3] = special[np.arange(flat[:, ::3].shape[1]) % special.size]
    dm_t = torch.tensor(dm)
    with np.errstate(all="ignore"):
        opc, ov = G.depth2pc(dm, K, clip=None, invalid=float("nan"))
        od, om = G.reproject(dm, K, P)
    gpc, gv = pg.depth2pc_tensor(dm_t.cuda(), torch.tensor(K).cuda(), clip=None, invalid_num=None)
    assert _same_bits_or_nan(gpc.cpu().numpy(), opc)
    assert np.array_equal(gv.cpu().numpy(), ov)
    gd, gm = pg.reproject_tensor(dm_t.cuda(), torch.tensor(K).cuda(), torch.tensor(P).cuda())
    assert _same_bits_or_nan(gd.cpu().numpy(), od)
    assert np.array_equal(gm.cpu().numpy(), om)


def test_pc2depth_ragged_bit_exact():
    B, H, W = 4, 256, 256
    d01, K, P = _inputs(B, H, W, 11)
    pcs = G.depth2pc_compact(d01.numpy(), K, None)
    pcs = [p.astype(np.float32) for p in pcs]
    pcs[2] = pcs[2][:0]                       # an empty cloud in the middle
    offs = np.cumsum([0] + [p.shape[0] for p in pcs])
    allpc = np.concatenate(pcs)
    for pose in (None, P):
        od, om = G.pc2depth(allpc, None, offs, K, (H, W), pose=pose)
        gd, gm = pg.pc2depth_ragged(torch.tensor(allpc).cuda(), torch.tensor(offs), torch.tensor(K),
                                    image_size=[H, W],
                                    pose=None if pose is None else torch.tensor(pose))
        assert np.array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32))
        assert np.array_equal(gm.cpu().numpy(), om)
        assert not om[2].any()


def test_pc2depth_dense_with_valid_mask():
    B, H, W = 2, 64, 80
    d01, K, _ = _inputs(B, H, W, 13)
    dm = d01 * 10
    pc, valid = G.depth2pc(dm.numpy(), K)
    offs = np.arange(B + 1) * (H * W)
    od, om = G.pc2depth(pc.reshape(-1, 3), valid.reshape(-1), offs, K, (H, W))
    gd, gm = pg.pc2depth_tensor(torch.tensor(pc).cuda(), torch.tensor(valid).cuda(),
                                torch.tensor(K).cuda(), image_size=[H, W])
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gm.cpu().numpy(), om)
    # identity reprojection reproduces the valid input pixels exactly
    assert np.array_equal(om[:, 0].reshape(B, -1), valid)


@pytest.mark.parametrize("shape", [(3, 256, 256), (2, 480, 640)])
def test_point_cloud_compact_f64(shape):
    B, H, W = shape
    d01, K, P = _inputs(B, H, W, 17)
    for pose in (None, P):
        ref = G.depth2pc_compact(d01.numpy(), K, pose)
        pc, counts = pg.point_cloud_batch(d01.cuda(), torch.tensor(K).cuda(),
                                          pose=None if pose is None else torch.tensor(pose).cuda())
        counts = counts.cpu().numpy()
        for b in range(B):
            assert counts[b] == ref[b].shape[0]
            got = pc[b, :counts[b]].cpu().numpy()
            assert np.array_equal(got.view(np.uint64), ref[b].view(np.uint64))


def test_reproject_full_size_properties():
    """BASELINE config-4 size (640x480 maps), size-independent checks: output is a z-buffer
    (every filled pixel holds a positive depth, empty pixels are exactly 0 <=> mask 0), the
    result is deterministic (order-independent atomic min), idempotent under re-running."""
    B, H, W = 64, 480, 640
    base, K, P = _inputs(4, H, W, 19)
    dm = (base * 10).repeat(16, 1, 1, 1).cuda()
    K = torch.tensor(K).repeat(16, 1, 1).cuda()
    P = torch.tensor(P).repeat(16, 1, 1).cuda()
    d1, m1 = pg.reproject_tensor(dm, K, P)
    d2, m2 = pg.reproject_tensor(dm, K, P)
    assert torch.equal(d1, d2) and torch.equal(m1, m2)
    assert torch.equal(d1 > 0, m1)
    assert torch.equal(d1[:4], d1[4:8])
    od, om = G.reproject(dm[:4].cpu().numpy(), K[:4].cpu().numpy(), P[:4].cpu().numpy())
    assert np.array_equal(d1[:4].cpu().numpy(), od)

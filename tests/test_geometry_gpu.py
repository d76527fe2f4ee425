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


def test_reproject_ring_reuse_many_maps_and_streams():
    """The fused reprojection keeps its z-buffers in a ring of scratch slots that a call reuses many
    times (300 maps of 256x256 through 64 slots; 70 maps of 640x480 through 20) and hands back empty:
    every map bit-exact against the oracle, repeated calls identical, and a call issued on another
    stream waits for the ring."""
    for (B, H, W, nref) in ((300, 256, 256, 24), (70, 480, 640, 6)):
        base, K, P = _inputs(nref, H, W, 29)
        rep = (B + nref - 1) // nref
        dm = (base * 10).repeat(rep, 1, 1, 1)[:B].cuda()
        Kt = torch.tensor(K).repeat(rep, 1, 1)[:B].cuda()
        Pt = torch.tensor(P).repeat(rep, 1, 1)[:B].cuda()
        d1, m1 = pg.reproject_tensor(dm, Kt, Pt)
        od, om = G.reproject((base * 10).numpy(), K, P)
        for s in range(0, B, nref):
            n = min(nref, B - s)
            assert np.array_equal(d1[s:s + n].cpu().numpy().view(np.uint32), od[:n].view(np.uint32)), (B, s)
            assert np.array_equal(m1[s:s + n].cpu().numpy(), om[:n]), (B, s)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            d2, m2 = pg.reproject_tensor(dm, Kt, Pt)          # other stream: ordered behind the first call
        d3, m3 = pg.reproject_tensor(dm[:1], Kt[:1], Pt[:1])  # B = 1 right behind it, on the first stream
        torch.cuda.synchronize()
        assert torch.equal(d1, d2) and torch.equal(m1, m2)
        assert torch.equal(d3, d1[:1]) and torch.equal(m3, m1[:1])


def test_reproject_unusual_maps_take_the_scalar_path():
    """Maps whose intrinsics / pose / clip fall outside the validated fast range (huge pose entries, a
    clip that admits negative or enormous depths) mixed with ordinary ones in one batch."""
    B, H, W = 6, 96, 128
    d01, K, P = _inputs(B, H, W, 31)
    dm = (d01 * 10).numpy()
    P = P.copy()
    P[1, 0, 3] = 2.5e4            # |t| beyond the fast-path bound
    P[2, :3, :3] *= 3e4           # absurd scale
    P[3, 2, 3] = -2.9             # pushes part of the map behind the camera (z <= 0)
    K = K.copy()
    K[4, 0, 2] = 3e6              # principal point outside the validated range
    dm[5, 0, ::7, ::5] = 1e-12    # tiny depths inside a clip that admits them
    for clip in ([0, 10], [-1.0, 3e38]):
        with np.errstate(all="ignore"):
            od, om = G.reproject(dm, K, P, clip=tuple(clip))
        gd, gm = pg.reproject_tensor(torch.tensor(dm).cuda(), torch.tensor(K).cuda(), torch.tensor(P).cuda(), clip=clip)
        a, b = gd.cpu().numpy(), od
        assert bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))), clip
        assert np.array_equal(gm.cpu().numpy(), om), clip


def test_range_restricted_reciprocal_is_the_ieee_one():
    """frcp_rn_normal (MUFU.RCP + one Newton step, no range checks) == __frcp_rn for EVERY float with
    2^-100 <= |z| <= 2^100 -- a superset of the (1e-6, 1e12) the reprojection kernel feeds it."""
    from pointreggpt_b200 import _ffi
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    _ffi.check(_ffi.lib().prg_test_frcp_exhaustive(2.0 ** -100, 2.0 ** 100, _ffi.ptr(bad), _ffi.stream(bad)))
    assert int(bad.item()) == 0

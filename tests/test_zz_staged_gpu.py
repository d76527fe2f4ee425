"""GPU parity tests of the section-8(f) kernels and drivers: occlusion filter, image condition,
voxel-grid down-sampling, overlap counting, Tester.sample, keep-mask dropout / denoise samplers.

These were staged as non-strict xfail at the end of round 1 (written after the GPU budget was spent);
the round-1 hardware run showed 24 passes and 3 failures, all three caused by `tensor / python_scalar`
being evaluated as a multiplication by the reciprocal on CUDA (cloud.voxel_down_sample's torch
formulation and geometry.image_condition), fixed since.  They are plain, strict tests now.
"""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from pointreggpt_b200 import cloud
from pointreggpt_b200 import geometry as pg
from pointreggpt_b200 import synthetic as S

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _reprojected(B, H, W, seed):
    d01 = S.synthetic_depth_batch(300 + seed, B, H, W)
    K = S.synthetic_intrinsics(B, 256 if H == 256 else None, seed=seed)
    if (H, W) not in ((256, 256), (480, 640)):
        K = K.copy()
        K[:, 0, 0] = K[:, 1, 1] = 1.2 * W
        K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    P = S.synthetic_poses(B, seed=seed + 1)
    rd, rm = G.reproject((d01 * 10).numpy(), K, P)
    return d01, K, P, rd, rm


@pytest.mark.parametrize("shape", [(2, 256, 256), (1, 480, 640), (2, 33, 47), (1, 5, 7), (1, 40, 260)])
def test_occlusion_filter_bit_exact(shape):
    B, H, W = shape
    _, _, _, rd, rm = _reprojected(B, H, W, 17)
    # threshold neighbours: exactly at, just below and just above 0.0375f behind the neighbour
    thr = np.float32(0.0375)
    if W >= 16 and H >= 4:
        for i, t in enumerate([thr, np.nextafter(thr, np.float32(0)), np.nextafter(thr, np.float32(1))]):
            rd[0, 0, 2, 4 * i + 1] = np.float32(1.25)
            rd[0, 0, 2, 4 * i + 2] = np.float32(1.25) + t
            rm[0, 0, 2, 4 * i + 1: 4 * i + 3] = True
    od, om = G.occlusion_filter(rd, rm)
    gd, gm = pg.occlusion_filter(torch.tensor(rd).cuda(), torch.tensor(rm).cuda())
    assert gd.dtype == torch.float32 and gm.dtype == torch.bool
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gm.cpu().numpy(), om)


def test_occlusion_filter_golden_and_empty():
    occ = np.load(os.path.join(GOLD, "occlusion.npz"))
    geo = np.load(os.path.join(GOLD, "geometry.npz"))
    d01 = S.synthetic_depth_batch(40, 2, 256, 256)
    K = torch.tensor(geo["K_256"]).cuda()
    for tag, P in (("pose", geo["P_256"]), ("fwd", occ["P_fwd"])):
        rd, rm = pg.reproject_tensor((d01 * 10).cuda(), K, torch.tensor(P).cuda())
        fd, fm = pg.occlusion_filter(rd, rm)
        assert np.array_equal(fd.cpu().numpy()[:, :, 96:160, 96:160], occ["out_crop_" + tag])
        assert int((fd != rd).sum()) == int(occ["changed_" + tag])
        assert torch.equal(fm, rm)
    # nothing valid: output equals input (zeros); B = 0 is a no-op
    z = torch.zeros(1, 1, 64, 64, device="cuda")
    fd, _ = pg.occlusion_filter(z, torch.zeros_like(z, dtype=torch.bool))
    assert (fd == 0).all()
    fd, _ = pg.occlusion_filter(z[:0], torch.zeros_like(z[:0], dtype=torch.bool))
    assert fd.shape == (0, 1, 64, 64)


def test_image_condition_matches_composition():
    B, H, W = 2, 256, 256
    d01, K, P, _, _ = _reprojected(B, H, W, 23)
    Kt, Pt = torch.tensor(K).cuda(), torch.tensor(P).cuda()
    for occl in (False, True):
        ic = pg.image_condition(d01.cuda(), Kt, Pt, use_occlusion_filter=occl)
        rd, rm = G.reproject((d01 * 10).numpy(), K, P)
        if occl:
            rd, rm = G.occlusion_filter(rd, rm)
        want = np.concatenate([rd / np.float32(10), rm.astype(np.float32)], 1) * np.float32(2) - np.float32(1)
        assert ic.shape == (B, 2, H, W)
        assert np.array_equal(ic.cpu().numpy(), want)


@pytest.mark.parametrize("n,voxel", [(200000, 0.025), (200000, 0.002), (5000, 0.1), (3, 0.5), (1, 0.1)])
def test_voxel_down_sample_native(n, voxel):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-1.0, 1.0, (n, 3)) + np.array([0.3, -1.1, 2.5])
    pts[: n // 5] = np.round(pts[: n // 5] / 0.025) * 0.025        # points on voxel faces
    pts[n // 5: n // 4] = pts[0]                                    # duplicates
    want_c, _ = G.voxel_down_sample(pts, voxel)
    t = torch.tensor(pts).cuda()
    got = cloud.voxel_down_sample(t, voxel)
    assert got.dtype == torch.float64 and got.shape == want_c.shape
    assert np.abs(got.cpu().numpy() - want_c).max() < 1e-10         # fixed-point sums: < 2e-11 m
    again = cloud.voxel_down_sample(t.flip(0), voxel)               # order of arrival does not matter
    assert torch.equal(got, again)
    lib = cloud.voxel_down_sample_torch(t, voxel)                   # independent torch-op formulation
    assert lib.shape == want_c.shape and np.abs(lib.cpu().numpy() - want_c).max() < 1e-10


def test_voxel_down_sample_native_rejects_bad_points():
    from pointreggpt_b200._ffi import PrgError
    pts = torch.rand(100, 3, dtype=torch.float64).cuda()
    assert cloud.voxel_down_sample_native(pts[:0], 0.1).shape == (0, 3)
    pts[7, 2] = float("nan")
    with pytest.raises(PrgError):
        cloud.voxel_down_sample_native(pts, 0.1)


@pytest.mark.parametrize("nq,nt,radius", [(60000, 50000, 0.0375), (5000, 70000, 0.1), (300, 1, 0.5)])
def test_overlap_count_exact(nq, nt, radius):
    from scipy.spatial import cKDTree
    from pointreggpt_b200 import overlap
    rng = np.random.default_rng(nq)
    q = rng.uniform(-1, 1, (nq, 3))
    t = rng.uniform(-0.5, 1.5, (nt, 3))
    want = sum(1 for x in cKDTree(t).query_ball_point(q, radius) if len(x) > 0)
    got = overlap.overlap_count(torch.tensor(q).cuda(), torch.tensor(t).cuda(), radius)
    assert got == want


def test_compute_overlap_ratio_matches_oracle():
    from pointreggpt_b200 import overlap
    d01 = S.synthetic_depth_batch(77, 2, 256, 256)
    K = S.synthetic_intrinsics(2, 256, seed=3)
    P = S.synthetic_poses(2, seed=4)
    a = G.depth2pc_compact(d01[:1].numpy(), K[:1], None)[0]
    b = G.depth2pc_compact(d01[:1].numpy(), K[:1], P[:1])[0]      # the same surface seen from a moved camera
    want = G.compute_overlap_ratio(a, b)
    got = overlap.compute_overlap_ratio(torch.tensor(a).cuda(), torch.tensor(b).cuda())
    # the voxel centroids come from torch index_add here (last-bit differences in the means are
    # possible), so a point exactly at the search radius may flip: allow a handful of points
    assert got == pytest.approx(want, abs=1e-3) and 0.05 < got[0] < 1.0


def test_tester_sample_successive_views(tmp_path):
    """Tester.sample (SDD:1961-2065): unconditional view, then views conditioned on the previous one
    reprojected 0.5 m forward through the occlusion filter; files as the reference names them."""
    from pointreggpt_b200 import nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    from pointreggpt_b200.tester import Tester
    torch.manual_seed(0)
    np.random.seed(0)
    unet = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    diff = GaussianDiffusion(unet, image_size=128, timesteps=8, sampling_timesteps=2, objective="pred_x0",
                             beta_schedule="sigmoid", ddim_sampling_eta=1.0)
    t = Tester(diff, batch_size=2, results_folder=str(tmp_path / "res"), samples_folder=str(tmp_path / "out"))
    strip = t.sample(num_scenes=3, num_samples=3)
    assert strip.shape == (3, 1, 128, 3 * 128) and torch.isfinite(strip).all()
    assert 0.0 <= float(strip.min()) and float(strip.max()) <= 1.0
    for scene in range(3):
        assert (tmp_path / "out" / f"scene-{scene}-camera-intrinsics.txt").is_file()
        for k in range(3):
            assert (tmp_path / "out" / f"scene-{scene}-sample-{k}.png").is_file()
            pts = cloud.read_ply(str(tmp_path / "out" / f"scene-{scene}-sample-{k}.ply"))
            assert pts.shape[1] == 3 and np.isfinite(pts).all()
    assert (tmp_path / "out" / "overview.png").is_file()


def test_generate_batch_against_composition_oracle():
    """pipeline.generate_batch (the per-batch body of Generator.generate, SDD:2479-2628) end to end
    against oracle/pipeline_ref.py, which is pinned bit-exactly to the reference's own call sequence
    (tests/test_oracle_vs_reference.py).  Geometry stages must agree exactly wherever the keep masks
    agree; the keep masks (sigmoid > 0.99 of a network output) may differ on the few pixels whose
    logit sits within the fp16 tolerance of the threshold."""
    from oracle import pipeline_ref
    from pointreggpt_b200 import nets, pipeline
    from pointreggpt_b200.diffusion import GaussianDiffusion
    B, SZ, T = 2, 128, 3
    torch.manual_seed(0)
    unet = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    mask = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
    with torch.no_grad():
        mask.final_conv[0].bias.fill_(4.6)           # logits straddle the 0.99 threshold (4.595)
    usd = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    msd = {k: v.detach().clone() for k, v in mask.state_dict().items()}
    diff = GaussianDiffusion(unet, image_size=SZ, timesteps=T, objective="pred_x0",
                             beta_schedule="sigmoid", is_ddnm_sampling=True).cuda()
    mask = mask.cuda()
    d01 = S.synthetic_depth_batch(55, B, SZ, SZ)
    K = S.synthetic_intrinsics(B, None, seed=3).copy()
    K[:, 0, 0] = K[:, 1, 1] = 1.2 * SZ
    K[:, 0, 2], K[:, 1, 2] = SZ / 2, SZ / 2
    P = S.synthetic_poses(B, seed=4)
    g = torch.Generator().manual_seed(6)
    n = diff.num_noise_draws(True)
    noise = torch.randn(n, B, 1, SZ, SZ, generator=g)
    want = pipeline_ref.generate_batch(usd, msd, d01, K, P, list(noise), timesteps=T, has_refine_step=True)
    pc, counts, images, mid = pipeline.generate_batch(
        diff, mask, d01.cuda(), torch.tensor(K).cuda(), torch.tensor(P).cuda(), has_refine_step=True,
        noise=noise.cuda(), return_intermediates=True)
    # reprojection + first keep mask
    m_got, m_want = mid["mask_rpj"].cpu(), want["mask_rpj"]
    agree = (m_got == m_want)
    assert agree.float().mean().item() > 0.995
    both = (m_got & m_want)
    assert torch.equal(mid["images_rpj"].cpu()[both], want["images_rpj"][both])     # z-buffer values: exact
    # sampled images where both pipelines kept the pixel and were conditioned alike
    img_got, img_want = images.cpu(), want["images"]
    kept = (img_got > 0) & (img_want > 0) & agree     # a flipped condition pixel changes its value outright
    assert kept.float().mean().item() > 0.3
    err = ((img_got - img_want)[kept].norm() / img_want[kept].norm()).item()
    assert err < 1e-2, err                           # T = 3 chained evaluations + refine, a few flipped inputs
    # clouds: same number of points up to the mask disagreements
    for b in range(B):
        nb = int(counts[b])
        assert abs(nb - want["clouds"][b].shape[0]) <= 0.02 * SZ * SZ
        assert np.isfinite(pc[b, :nb].cpu().numpy()).all()


@pytest.mark.parametrize("mode", ["ddnm_none", "ddnm_linear_ddim", "denoise"])
def test_keep_mask_dropout_sampling(mode):
    """DDNM keep-mask dropout and denoise() (SDD:1213-1225) through the step-wise host loop, with the
    Gaussian and uniform draws injected, against the oracle (pinned to the reference with the same
    injection in tests/test_oracle_vs_reference.py)."""
    from oracle import torch_ref as R
    from pointreggpt_b200 import nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    torch.manual_seed(0)
    net = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    SZ, B = 128, 2
    cfg = dict(ddnm_none=dict(timesteps=4, ddnm_sampling_dropout=0.3),
               ddnm_linear_ddim=dict(timesteps=8, sampling_timesteps=3, ddnm_sampling_dropout=0.5,
                                     ddnm_dropout_schedule='linear', ddim_sampling_eta=1.0),
               denoise=dict(timesteps=4, is_ddnm_sampling=False))[mode]
    diff = GaussianDiffusion(net, image_size=SZ, objective='pred_x0', beta_schedule='sigmoid', **cfg).cuda()
    T = cfg["timesteps"]
    g = torch.Generator().manual_seed(10)
    noises = torch.stack([torch.randn(B, 1, SZ, SZ, generator=g) for _ in range(9)])
    uniforms = torch.stack([torch.rand(B, 1, SZ, SZ, generator=g) for _ in range(9)])
    d = torch.rand(B, 1, SZ, SZ, generator=g) * 0.3
    d[torch.rand(B, 1, SZ, SZ, generator=g) < 0.4] = 0
    ic = torch.cat([d, (d > 0).float()], 1) * 2 - 1
    pc = torch.tensor([[303.9, 304.2, 64.5, 64.], [290., 291., 64.5, 64.]])
    ddnm, denoise = R.dropout_tables(T, cfg.get("ddnm_sampling_dropout", 0.), cfg.get("ddnm_dropout_schedule", "none"))
    keep = R.KeepMask("denoise", denoise, list(uniforms)) if mode == "denoise" else R.KeepMask("ddnm", ddnm, list(uniforms))
    sch = R.make_schedule(T)
    if "sampling_timesteps" in cfg:
        want = R.ddim_sample(sd, sch, pc, ic, list(noises), cfg["sampling_timesteps"], 1.0, has_refine_step=True, keep=keep)
    else:
        want = R.p_sample_loop(sd, sch, pc, ic, list(noises), has_refine_step=True, keep=keep)
    fn = diff.denoise if mode == "denoise" else diff.sample
    got = fn(param_cond=pc.cuda(), img_cond=ic.cuda(), has_refine_step=True, noise=noises.cuda(),
             keep_uniform=uniforms.cuda()).cpu()
    err = ((got - want).norm() / want.norm()).item()
    assert err <= 3e-3, err
    # without injection: runs, seeded, and differs from the no-dropout result
    a = fn(param_cond=pc.cuda(), img_cond=ic.cuda(), seed=5)
    b_ = fn(param_cond=pc.cuda(), img_cond=ic.cuda(), seed=5)
    assert torch.equal(a, b_) and torch.isfinite(a).all()


@pytest.mark.parametrize("shape", [(1, 5, 7), (2, 9, 3), (1, 44, 1), (1, 1, 1), (2, 3, 15)])
def test_reproject_tiny_and_narrow_maps(shape):
    """Maps of at most 44 pixels (ATen's scalar bmm rounding, a separate kernel instantiation) and
    maps narrower than four pixels (several row wraps per thread)."""
    B, H, W = shape
    d01 = S.synthetic_depth_batch(500, B, H, W)
    K = S.synthetic_intrinsics(B, None, seed=2).copy()
    K[:, 0, 0] = K[:, 1, 1] = 1.2 * max(W, 4)
    K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    P = S.synthetic_poses(B, seed=3)
    dm = d01 * 10
    od, om = G.reproject(dm.numpy(), K, P)
    gd, gm = pg.reproject_tensor(dm.cuda(), torch.tensor(K).cuda(), torch.tensor(P).cuda())
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gm.cpu().numpy(), om)
    opc, ov = G.depth2pc(dm.numpy(), K, clip=[0, 10])
    gpc, gv = pg.depth2pc_tensor(dm.cuda(), torch.tensor(K).cuda(), clip=[0, 10])
    a, b = gpc.cpu().numpy(), opc
    assert np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))
    assert np.array_equal(gv.cpu().numpy(), ov)


def test_tester_generate_scene_fusion(tmp_path):
    """Tester.generate (SDD:2099-2247): every further view is conditioned on the z-buffer of the
    voxel-merged cloud of the previous views under a random in-view rotation; files as the reference
    names them; the fused 25 mm cloud of a scene stays inside the 0.5-3.5 m clip it was built from."""
    from pointreggpt_b200 import nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    from pointreggpt_b200.tester import Tester
    torch.manual_seed(0)
    np.random.seed(0)
    unet = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    diff = GaussianDiffusion(unet, image_size=128, timesteps=8, sampling_timesteps=2, objective="pred_x0",
                             beta_schedule="sigmoid", ddim_sampling_eta=1.0)
    t = Tester(diff, batch_size=2, results_folder=str(tmp_path / "res"), samples_folder=str(tmp_path / "out"))
    strip = t.generate(num_scenes=3, num_samples=3)
    assert strip.shape == (3, 1, 128, 3 * 128) and torch.isfinite(strip).all()
    assert strip.min() >= 0 and strip.max() <= 1
    for scene in range(3):
        for k in range(3):
            assert (tmp_path / "out" / ("scene-%d-sample-%d.png" % (scene, k))).is_file()
        pts = cloud.read_ply(str(tmp_path / "out" / ("scene-%d.ply" % scene)))
        assert pts.shape[1] == 3 and np.isfinite(pts).all()
        if pts.shape[0]:
            assert np.linalg.norm(pts, axis=1).max() < 3.5 * 2.0     # clip 3.5 m depth, 128-px field of view
    assert (tmp_path / "out" / "overview.png").is_file()

"""Live cross-check of the oracle against the unmodified reference modules.  Only runs where
/root/reference exists (the build container); skipped on the GPU box."""
import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from oracle import torch_ref as R
from oracle.ref_import import load_reference, reference_available
from pointreggpt_b200 import synthetic as S

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def test_geometry_bit_exact(ref):
    sdd, _ = ref
    B, H, W = 3, 96, 128
    d01 = S.synthetic_depth_batch(7, B, H, W)
    K = S.synthetic_intrinsics(B, None, seed=1).copy()
    K[:, 0, 0] = K[:, 1, 1] = 150.0
    K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    P = S.synthetic_poses(B, seed=2)
    dm = d01 * 10
    rd, rm = sdd.reproject_tensor(dm, torch.tensor(K), torch.tensor(P))
    od, om = G.reproject(dm.numpy(), K, P)
    assert np.array_equal(rd.numpy(), od) and np.array_equal(rm.numpy(), om) and om.any()
    rpc, rv = sdd.depth2pc_tensor(dm, torch.tensor(K), clip=[0.5, 10], invalid_num=0)
    opc, ov = G.depth2pc(dm.numpy(), K, clip=(0.5, 10), invalid=0.0)
    assert np.array_equal(rpc.numpy(), opc) and np.array_equal(rv.numpy(), ov)
    rp = sdd.point_cloud(d01[0, 0].numpy() * 10, K[0], clip=[0.5, 10])
    op = G.depth2pc_compact(d01[:1].numpy(), K[:1], None)[0]
    assert rp.dtype == np.float64 and np.array_equal(rp, op)


def test_unet_and_sampler_bit_exact(ref):
    sdd, dc = ref
    torch.manual_seed(3)
    net = sdd.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).eval()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 1, 64, 64, generator=g)
    t = torch.tensor([11, 800])
    pc = torch.tensor([[304., 304., 128.5, 128.], [290., 291., 128.5, 128.]])
    with torch.no_grad():
        assert torch.equal(net(x, t, pc), R.unet_forward(sd, x, t, pc))
    torch.manual_seed(3)
    m = dc.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).eval()
    msd = {k: v.detach() for k, v in m.state_dict().items()}
    d = torch.rand(2, 1, 64, 64, generator=g)
    d[d < 0.3] = 0
    with torch.no_grad():
        assert torch.equal(m(d), R.maskunet_forward(msd, d))
    diff = sdd.GaussianDiffusion(net, image_size=64, timesteps=4, objective='pred_x0',
                                 beta_schedule='sigmoid', is_ddnm_sampling=True)
    noises = [torch.randn(2, 1, 64, 64, generator=g) for _ in range(5)]
    ic = torch.cat([d, (d > 0).float()], 1) * 2 - 1
    it = iter(noises)
    o1, o2 = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    try:
        want = diff.sample(param_cond=pc, img_cond=ic, disable_tqdm=True, has_refine_step=True)
    finally:
        torch.randn, torch.randn_like = o1, o2
    got = R.p_sample_loop(sd, R.make_schedule(4), pc, ic, noises, has_refine_step=True)
    assert torch.equal(want, got)


def test_occlusion_filter_bit_exact(ref):
    sdd, _ = ref
    B, H, W = 3, 96, 128
    d01 = S.synthetic_depth_batch(21, B, H, W)
    K = S.synthetic_intrinsics(B, None, seed=1).copy()
    K[:, 0, 0] = K[:, 1, 1] = 150.0
    K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    P = S.synthetic_poses(B, seed=5)
    rd, rm = sdd.reproject_tensor(d01 * 10, torch.tensor(K), torch.tensor(P))
    # differences exactly at, one ulp below and one ulp above the fp32 threshold, next to each other
    thr = np.float32(0.0375)
    base = np.float32(1.25)
    edge = rd.clone()
    for i, t in enumerate([thr, np.nextafter(thr, np.float32(0)), np.nextafter(thr, np.float32(1))]):
        edge[0, 0, 10, 4 * i + 8] = float(base)
        edge[0, 0, 10, 4 * i + 9] = float(np.float32(base + t))
        rm[0, 0, 10, 4 * i + 8: 4 * i + 10] = True
    edge[1, 0, 40:50, 30:50] = 0.0          # an empty region: all-invalid windows
    rm[1, 0, 40:50, 30:50] = False
    for dep in (rd, edge):
        want_d, want_m = sdd.occlusion_filter(dep.clone(), rm.clone())
        got_d, got_m = G.occlusion_filter(dep.numpy(), rm.numpy())
        assert np.array_equal(want_d.numpy().view(np.uint32), got_d.view(np.uint32))
        assert np.array_equal(want_m.numpy(), got_m)
    assert (want_d != edge).any()            # the filter did replace something


def test_generate_batch_composition_bit_exact(ref):
    """oracle/pipeline_ref.generate_batch against the same sequence of reference calls as
    Generator.generate's per-batch body (SDD:2479-2628; open3d's crop restated as an inclusive box)."""
    from oracle import pipeline_ref
    sdd, dc = ref
    B, SZ, T = 2, 64, 3
    torch.manual_seed(5)
    net = sdd.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).eval()
    mnet = dc.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).eval()
    with torch.no_grad():
        mnet.final_conv[0].bias.fill_(4.6)          # logits straddle the 0.99 threshold (4.595)
    diff = sdd.GaussianDiffusion(net, image_size=SZ, timesteps=T, objective='pred_x0',
                                 beta_schedule='sigmoid', is_ddnm_sampling=True)
    d01 = S.synthetic_depth_batch(55, B, SZ, SZ)
    K = S.synthetic_intrinsics(B, None, seed=3).copy()
    K[:, 0, 0] = K[:, 1, 1] = 1.2 * SZ
    K[:, 0, 2], K[:, 1, 2] = SZ / 2, SZ / 2
    P = S.synthetic_poses(B, seed=4)
    g = torch.Generator().manual_seed(6)
    noises = [torch.randn(B, 1, SZ, SZ, generator=g) for _ in range(T + 1)]
    # ---- the reference's own sequence
    lo, hi = np.array([-1.5, -1.5, 0.5]), np.array([1.5, 1.5, 3.5])
    rpj, msk = [], []
    for b in range(B):
        pc = sdd.point_cloud(d01[b, 0].numpy() * 10, K[b], clip=[0.5, 10]).astype(np.float32)
        pc = pc[np.all((pc >= lo) & (pc <= hi), axis=1)]
        moved = pc @ P[b, :3, :3].T + P[b, :3, 3]
        d, m = sdd.pc2depth_tensor(torch.tensor(moved[None]), torch.ones((1, moved.shape[0]), dtype=torch.bool),
                                   torch.tensor(K[b][None]), image_size=[SZ, SZ])
        rpj.append(d)
        msk.append(m)
    with torch.no_grad():
        images_rpj = torch.cat(rpj) * 0.1
        mask_rpj = torch.cat(msk)
        mask_crt = mnet(images_rpj) > 0.99
        images_rpj[~mask_crt] = 0
        mask_rpj = mask_rpj & mask_crt
        img_cond = sdd.normalize_to_neg_one_to_one(torch.cat([images_rpj, mask_rpj], dim=1))
        it = iter(noises)
        o1, o2 = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: next(it).clone()
        torch.randn_like = lambda *a, **k: next(it).clone()
        try:
            images = diff.sample(param_cond=sdd.param_vector(torch.tensor(K)), img_cond=img_cond,
                                 disable_tqdm=True, has_refine_step=True)
        finally:
            torch.randn, torch.randn_like = o1, o2
        mask_crt2 = mnet(images) > 0.99
        images[~mask_crt2] = 0
    clouds = []
    for b in range(B):
        pc = sdd.point_cloud(images[b, 0].numpy() * 10, K[b], clip=[0.5, 10])
        clouds.append((pc - P[b, :3, 3]) @ P[b, :3, :3])
    # ---- the oracle composition
    got = pipeline_ref.generate_batch({k: v.detach() for k, v in net.state_dict().items()},
                                      {k: v.detach() for k, v in mnet.state_dict().items()},
                                      d01, K, P, noises, timesteps=T, has_refine_step=True)
    assert torch.equal(got["images_rpj"], images_rpj) and torch.equal(got["mask_rpj"], mask_rpj)
    assert torch.equal(got["img_cond"], img_cond)
    assert torch.equal(got["images"], images)
    assert 0 < int(mask_crt.sum()) < mask_crt.numel()
    for a, b_ in zip(got["clouds"], clouds):
        assert np.array_equal(a, b_)



@pytest.mark.parametrize("mode", ["ddnm_none", "ddnm_linear_ddim", "denoise"])
def test_keep_mask_dropout_bit_exact(ref, mode):
    """The Bernoulli keep-mask of model_predictions (SDD:1213-1225) with injected uniform draws."""
    sdd, _ = ref
    torch.manual_seed(9)
    net = sdd.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).eval()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    SZ, B = 32, 2
    cfg = dict(ddnm_none=dict(timesteps=4, ddnm_sampling_dropout=0.3),
               ddnm_linear_ddim=dict(timesteps=8, sampling_timesteps=3, ddnm_sampling_dropout=0.5,
                                     ddnm_dropout_schedule='linear', ddim_sampling_eta=1.0),
               denoise=dict(timesteps=4, is_ddnm_sampling=False))[mode]
    diff = sdd.GaussianDiffusion(net, image_size=SZ, objective='pred_x0', beta_schedule='sigmoid', **cfg)
    g = torch.Generator().manual_seed(10)
    noises = [torch.randn(B, 1, SZ, SZ, generator=g) for _ in range(9)]
    uniforms = [torch.rand(B, 1, SZ, SZ, generator=g) for _ in range(9)]
    d = torch.rand(B, 1, SZ, SZ, generator=g)
    d[d < 0.3] = 0
    ic = torch.cat([d, (d > 0).float()], 1) * 2 - 1
    pc = torch.tensor([[304., 304., 16.5, 16.], [290., 291., 16.5, 16.]])
    it, iu = iter(noises), iter(uniforms)
    o1, o2, o3 = torch.randn, torch.randn_like, torch.Tensor.uniform_
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    torch.Tensor.uniform_ = lambda self, a=0, b=1: self.copy_(next(iu))
    try:
        fn = diff.denoise if mode == "denoise" else diff.sample
        want = fn(param_cond=pc, img_cond=ic, disable_tqdm=True, has_refine_step=True)
    finally:
        torch.randn, torch.randn_like, torch.Tensor.uniform_ = o1, o2, o3
    T = cfg["timesteps"]
    ddnm, denoise = R.dropout_tables(T, cfg.get("ddnm_sampling_dropout", 0.), cfg.get("ddnm_dropout_schedule", "none"))
    keep = R.KeepMask("denoise", denoise, uniforms) if mode == "denoise" else R.KeepMask("ddnm", ddnm, uniforms)
    sch = R.make_schedule(T)
    if "sampling_timesteps" in cfg:
        got = R.ddim_sample(sd, sch, pc, ic, noises, cfg["sampling_timesteps"], 1.0, has_refine_step=True, keep=keep)
    else:
        got = R.p_sample_loop(sd, sch, pc, ic, noises, has_refine_step=True, keep=keep)
    assert torch.equal(want, got)
    plain = R.p_sample_loop(sd, sch, pc, ic, noises, has_refine_step=True) if "sampling_timesteps" not in cfg else None
    if plain is not None:
        assert not torch.equal(plain, got)          # the dropout did change the result


def test_occlusion_filter_fuzz_bit_exact(ref):
    """Random shapes (down to 1x1), quantised depths whose differences land exactly on the 0.0375
    threshold, infinities: oracle == reference bit for bit.  (NaN depths are outside the contract:
    the inputs are z-buffer outputs, and torch's max_pool2d would propagate NaN where the oracle
    skips it.)"""
    from test_kernel_emulation import _occlusion_case, _same_bits_or_nan     # tests/ is on sys.path (rootdir conftest)
    sdd, _ = ref
    rng = np.random.default_rng(1)
    for trial in range(100):
        d, m = _occlusion_case(rng)
        want, _ = sdd.occlusion_filter(torch.tensor(d), torch.tensor(m))
        got, _ = G.occlusion_filter(d, m)
        assert _same_bits_or_nan(got, want.numpy()), (trial, d.shape)


def test_geometry_fuzz_bit_exact(ref):
    """Differential fuzz of the C oracle against the reference: small random maps (down to one pixel,
    widths below four, at most 44 pixels where ATen's bmm switches to its scalar loop), intrinsics
    and depths from lists of awkward values, random clips and poses.  Single-map batches: for some
    small batched shapes (seen: B = 2 with N = 100 or 200) the BLAS batch kernel rounds the last
    N mod 32 rows differently from the rest -- a property of the library build, absent at the sizes of
    the path (checked up to B = 32, N = 65 536 and B = 8, N = 307 200) and not restated."""
    sdd, _ = ref
    rng = np.random.default_rng(0)
    FX = [0.5, 0.999, 1.0, 1.5, 300.0, 585.0, 1e6, 2e6, 1e-3, 1e9, -300.0, 511.99997]
    CX = [0.0, -0.0, 1e-4, 1e-3, 128.5, 320.0, 1e6, -5.0, 1e-30, 5.0000005]
    Z = np.array([0.0, -0.0, -1.5, 1e-42, 1e-30, 1e-12, 1e-9, 1e9, 1e12, 1e30, 3e38, np.inf, -np.inf, np.nan, 1.0,
                  2.5e-7, 65504.0], np.float32)

    def same(x, y):
        return bool(np.all((x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))))

    for trial in range(150):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        dm = (rng.random((1, 1, H, W)) * 10).astype(np.float32)
        m = rng.random((1, 1, H, W)) < 0.3
        dm[m] = Z[rng.integers(0, Z.size, int(m.sum()))]
        K = np.zeros((1, 3, 3), np.float32)
        K[:, 2, 2] = 1
        if rng.random() < 0.5:
            K[0, 0, 0], K[0, 1, 1], K[0, 0, 2], K[0, 1, 2] = rng.choice(FX), rng.choice(FX), rng.choice(CX), rng.choice(CX)
        else:
            K[0, 0, 0], K[0, 1, 1] = 1.2 * W * rng.uniform(0.5, 2), 1.2 * W * rng.uniform(0.5, 2)
            K[0, 0, 2], K[0, 1, 2] = W / 2 + rng.uniform(-1, 1), H / 2
        P = np.eye(4, dtype=np.float32)[None].copy()
        ang = rng.uniform(-0.3, 0.3)
        P[0, 0, 0], P[0, 0, 2], P[0, 2, 0], P[0, 2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
        P[0, :3, 3] = rng.normal(0, 0.3, 3)
        clip = [[0.0, 10.0], [0.5, 10.0], None, [-1.0, 1e30]][int(rng.integers(0, 4))]
        inv = [None, 0.0][int(rng.integers(0, 2))]
        rclip = [[0.0, 10.0], [0.5, 3.5], [-1.0, 3e38]][int(rng.integers(0, 3))]
        with np.errstate(all="ignore"):
            rpc, rv = sdd.depth2pc_tensor(torch.tensor(dm), torch.tensor(K), clip=clip, invalid_num=inv)
            opc, ov = G.depth2pc(dm, K, clip=clip, invalid=float("nan") if inv is None else inv)
            rd, rm = sdd.reproject_tensor(torch.tensor(dm), torch.tensor(K), torch.tensor(P), clip=rclip)
            od, om = G.reproject(dm, K, P, clip=rclip)
        ctx = (trial, H, W, K[0, [0, 1, 0, 1], [0, 1, 2, 2]].tolist(), clip, rclip)
        assert same(rpc.numpy(), opc) and np.array_equal(rv.numpy(), ov), ctx
        assert same(rd.numpy(), od) and np.array_equal(rm.numpy(), om), ctx


@pytest.mark.parametrize("dim,mults,size,batch", [(32, (1, 2, 4, 8), 32, 3), (48, (1, 2, 4), 48, 2), (64, (1, 2), 32, 1)])
def test_network_oracle_other_architectures_bit_exact(ref, dim, mults, size, batch):
    """The functional restatement is generic in width, depth, image size and batch: bit-exact against
    the reference modules for configurations other than the shipped one."""
    sdd, dc = ref
    torch.manual_seed(dim + size)
    net = sdd.Unet(dim=dim, param_cond_dim=4, dim_mults=mults, channels=1).eval()
    m = dc.MaskUnet(dim=dim, dim_mults=mults).eval()
    x = torch.randn(batch, 1, size, size)
    t = torch.randint(0, 1000, (batch,))
    pc = torch.rand(batch, 4) * 300
    d = torch.rand(batch, 1, size, size)
    d[d < 0.3] = 0
    with torch.no_grad():
        assert torch.equal(net(x, t, pc), R.unet_forward({k: v.detach() for k, v in net.state_dict().items()}, x, t, pc))
        assert torch.equal(m(d), R.maskunet_forward({k: v.detach() for k, v in m.state_dict().items()}, d))


def test_host_samplers_match_the_reference_rng_stream(ref):
    """random_sample_pose / random_sample_intrinsic / random_sample_transform / intrinsic_transform of
    the product package against the reference's functions under the same numpy seed (same draws in
    the same order -> identical matrices)."""
    sdd, _ = ref
    from pointreggpt_b200 import geometry as pg
    for seed in (0, 5, 99):
        np.random.seed(seed)
        Kr = sdd.random_sample_intrinsic(7)
        Pr = sdd.random_sample_pose(7)
        Ktr = sdd.intrinsic_transform(Kr, resize=256, centercrop=256).astype(np.float32)
        Tr = sdd.random_sample_transform(Ktr, image_size=256)
        np.random.seed(seed)
        Ko = pg.random_sample_intrinsic(7)
        Po = pg.random_sample_pose(7)
        Kto = pg.intrinsic_transform(Ko, resize=256, centercrop=256).astype(np.float32)
        To = pg.random_sample_transform(Kto, image_size=256)
        assert np.array_equal(Kr, Ko) and np.array_equal(Pr, Po) and np.array_equal(Ktr, Kto)
        assert np.array_equal(Tr, To) and To.dtype == np.float32
        assert np.all(To[:, :3, 3] == 0)

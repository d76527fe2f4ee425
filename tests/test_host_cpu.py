"""CPU-side checks of the product package: the C ABI loads and exports every symbol of
include/prg.h, compute entries fail loudly without a GPU (no CPU fallback), packing, sampler
step tables, state-dict compatibility, rank sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import torch_ref as R
from pointreggpt_b200 import _ffi, dist as pdist, geometry as pg, nets, packing
from pointreggpt_b200.diffusion import GaussianDiffusion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "prg.h")).read()
    declared = set(re.findall(r"\b(prg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libprg.so does not export " + name
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert _ffi.lib().prg_abi_version() == 2


@pytest.mark.skipif(not NO_GPU, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    x = torch.zeros(1, 1, 8, 8)
    K = torch.eye(3)[None]
    with pytest.raises(_ffi.PrgError):
        pg.reproject_tensor(x, K, torch.eye(4)[None])
    torch.manual_seed(0)
    net = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
    with pytest.raises(_ffi.PrgError):
        net(torch.zeros(1, 1, 128, 128))
    # creating a handle without a device must fail with an error code, not crash
    h = ctypes.c_void_p()
    blob = net._pack()
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = _ffi.lib().prg_net_create(ctypes.byref(h), packing.KIND_MASKUNET, buf, len(blob), 1, 128, 0)
    assert rc != 0 and _ffi.lib().prg_last_error()


def test_packing_layouts():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(5, 3, 3, 3, generator=g)
    k = packing.conv_weight_kmajor(w)
    assert k.shape == (5, 27)
    assert torch.equal(k[2, (1 * 3 + 2) * 3 + 1], w[2, 1, 1, 2])
    # folded nearest-x2 upsample + conv3x3 == four 2x2 parity convs
    x = torch.randn(1, 3, 6, 6, generator=g)
    ref = torch.nn.functional.conv2d(torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest"),
                                     w, padding=1)
    f = packing.upsample_fold_weight(w).reshape(5, 4, 4, 3)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in (0, 1):
        for px in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    dy, dx = ty - (1 - py), tx - (1 - px)
                    patch = xp[:, :, 1 + dy:1 + dy + 6, 1 + dx:1 + dx + 6]
                    out[:, :, py::2, px::2] += torch.einsum("oc,bchw->bohw", f[:, py * 2 + px, ty * 2 + tx], patch)
    assert torch.allclose(out, ref, atol=1e-5)
    # the same once more as the class-bound row-streaming kernel sees it (conv2_tc.cu, rows3): per parity class a
    # plain 3x3 conv (pad 1) whose ky = 2 / kx = 2 taps are zero, on the input shifted by (py, px)
    r3 = packing.upsample_rows3_weight(w).reshape(5, 4, 3, 3, 3)          # [co][class][ky][kx][ci]
    assert torch.count_nonzero(r3[:, :, 2]) == 0 and torch.count_nonzero(r3[:, :, :, 2]) == 0
    out3 = torch.zeros_like(ref)
    for py in (0, 1):
        for px in (0, 1):
            # out[y][x] = sum w[ky][kx] in[y + ky - 1 + py][x + kx - 1 + px], zero outside the image (TMA fill)
            xq = torch.nn.functional.pad(x, (1, 2, 1, 2))
            window = xq[:, :, py:py + 8, px:px + 8]
            wk = r3[:, py * 2 + px].permute(0, 3, 1, 2).contiguous()      # (co, ci, ky, kx)
            out3[:, :, py::2, px::2] = torch.nn.functional.conv2d(window, wk)
    assert torch.allclose(out3, ref, atol=1e-5)
    ws = packing.standardize(w)
    assert torch.allclose(ws, R.standardize_weight(w))


def test_blob_roundtrip_header():
    torch.manual_seed(0)
    net = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
    blob = net._pack()
    assert blob[:4] == b"PRGW"
    assert len(blob) % 256 == 0 and len(blob) > 60e6


def test_state_dict_layout_and_loading():
    torch.manual_seed(0)
    u = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    d = GaussianDiffusion(u, image_size=256, timesteps=1000, sampling_timesteps=250,
                          objective="pred_x0", beta_schedule="sigmoid")
    sd = d.state_dict()
    assert len(sd) == 293 and len(u.state_dict()) == 280
    assert list(sd)[:3] == ["betas", "alphas_cumprod", "alphas_cumprod_prev"]
    assert "model.downs.0.2.fn.fn.to_out.1.g" in sd and "model.ups.0.3.1.weight" in sd
    assert len(nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).state_dict()) == 234
    d.load_state_dict(sd)     # a reference checkpoint's 'model' entry has exactly this layout
    assert d.is_ddim_sampling and d.num_timesteps == 1000


def test_sampling_step_tables_follow_reference_arithmetic():
    torch.manual_seed(0)
    u = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    d = GaussianDiffusion(u, image_size=256, timesteps=50, objective="pred_x0", beta_schedule="sigmoid")
    sch = R.make_schedule(50)
    steps = d.sampling_steps(has_refine_step=True)
    assert len(steps) == 51 and d.num_noise_draws(True) == 50
    assert [s.t for s in steps[:50]] == list(reversed(range(50)))
    for s in steps[:50]:
        assert s.kind == _ffi.STEP_P_SAMPLE and s.add_noise == int(s.t > 0)
        assert s.c0 == float(sch["posterior_mean_coef1"][s.t])
        assert s.c2 == float((0.5 * sch["posterior_log_variance_clipped"][s.t]).exp())
    assert steps[-1].kind == _ffi.STEP_REFINE_P and steps[-1].unnormalize == 1
    d2 = GaussianDiffusion(u, image_size=256, timesteps=1000, sampling_timesteps=250,
                           objective="pred_x0", beta_schedule="sigmoid", ddim_sampling_eta=1.0)
    st = d2.sampling_steps()
    assert len(st) == 250 and st[0].t == 999 and st[-1].kind == _ffi.STEP_DDIM_LAST
    ac = R.make_schedule(1000)["alphas_cumprod"]
    t, tn = st[0].t, st[1].t
    sigma = 1.0 * ((1 - ac[t] / ac[tn]) * (1 - ac[tn]) / (1 - ac[t])).sqrt()
    assert st[0].c4 == float(sigma) and st[0].c2 == float(ac[tn].sqrt())


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        nets.Unet(dim=64, param_cond_dim=4, learned_variance=True)
    torch.manual_seed(0)
    u = nets.Unet(dim=64, param_cond_dim=4)
    with pytest.raises(NotImplementedError):
        GaussianDiffusion(u, image_size=256, objective="pred_noise")
    with pytest.raises(ValueError):
        GaussianDiffusion(u, image_size=256, objective="pred_x0", ddnm_dropout_schedule="cosine")
    d = GaussianDiffusion(u, image_size=256, objective="pred_x0", timesteps=10, ddnm_sampling_dropout=0.1,
                          ddnm_dropout_schedule="linear")
    assert d.ddnm_dropouts.dtype == torch.float64 and float(d.ddnm_dropouts[0]) == 0.1 and float(d.ddnm_dropouts[-1]) == 0.0
    assert float(d.denoise_dropouts[0]) == 1.0 and float(d.denoise_dropouts[-1]) == 0.0
    assert "ddnm_dropouts" not in d.state_dict()            # plain tensors, not buffers (SDD:1076-1094)


def test_shard_range_covers_everything_once():
    for n, w in [(10, 4), (3, 8), (256, 8), (0, 2), (7, 1)]:
        seen = []
        for r in range(w):
            lo, hi = pdist.shard_range(100, 100 + n, r, w)
            assert 100 <= lo <= hi <= 100 + n
            seen += list(range(lo, hi))
        assert seen == list(range(100, 100 + n))
    assert pg.num_to_groups(10, 4) == [4, 4, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)            # different init per rank ...
    net = nets.MaskUnet(dim=64, dim_mults=(1,))
    nbytes = pdist.broadcast_weights([net], src=0)   # ... identical after the broadcast
    fp = sum(float(p.double().abs().sum()) for p in net.parameters())
    lo, hi = pdist.shard_range(0, 5, rank, world)
    tot = pdist.sum_counters([hi - lo, 1.0], torch.device("cpu"))
    ret[rank] = (fp, nbytes, lo, hi, tot)
    dist.destroy_process_group()


def test_two_rank_weight_broadcast_and_sharding_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    (fp0, nb0, lo0, hi0, tot0), (fp1, nb1, lo1, hi1, tot1) = ret[0], ret[1]
    assert fp0 == fp1 and nb0 == nb1 > 0
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 5)
    assert tot0 == tot1 == [5.0, 2.0]


def test_generate_gt_file_logic(tmp_path, monkeypatch):
    """generate_gt.py:105-199: pair enumeration, the < 1000 points and < 0.1 / < 0.1 filters, the TSV
    format, skip-if-exists and gather.  The GPU ratio is replaced by the oracle's (CPU)."""
    import numpy as np
    import torch
    from oracle import geometry_ref as G
    from pointreggpt_b200 import cloud, overlap
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(overlap, "compute_overlap_ratio",
                        lambda a, b: G.compute_overlap_ratio(a.cpu().numpy(), b.cpu().numpy()))
    rng = np.random.default_rng(0)
    base = rng.uniform(0, 1, (3000, 3))
    clouds = {0: base, 1: base + [0.3, 0, 0], 2: base + [5.0, 0, 0], 3: base[:500]}
    sdir = tmp_path / "ds" / "data" / "scene-000004"
    sdir.mkdir(parents=True)
    for i, c in clouds.items():
        cloud.write_ply(str(sdir / "sample-{:0>6d}.cloud.ply".format(i)), torch.tensor(c))
    (tmp_path / "ds" / "data" / "scene-000005").mkdir()
    assert overlap.generate_gt("ds", 4, 6, 4, device="cpu") == 2
    lines = (sdir / "gt.log").read_text().splitlines()
    assert len(lines) == 1                                  # only (0, 1): 2 is far away, 3 is too small
    name, s, t, o1, o2 = lines[0].split("\t")
    r1, r2 = G.compute_overlap_ratio(clouds[0], clouds[1])
    assert (name, s, t) == ("scene-000004", "0", "1") and o1 == "%.4f" % r1 and o2 == "%.4f" % r2
    assert (tmp_path / "ds" / "data" / "scene-000005" / "gt.log").read_text() == ""
    assert overlap.generate_gt("ds", 4, 6, 4, device="cpu") == 0       # both exist now
    final = overlap.gather_gt("ds", 4, 6)
    assert open(final).read().splitlines() == lines


def test_tester_sample_host_logic(tmp_path, monkeypatch):
    """Tester.sample's orchestration (SDD:1961-2065) with the device operations stubbed: views are
    chained, the occlusion filter only runs once the camera has moved, files follow the reference's
    names."""
    import numpy as np
    import torch
    from pointreggpt_b200 import geometry, tester
    S = 32
    calls = {"occ": 0, "cond": []}

    class FakeModel:
        channels, image_size = 1, S

        def to(self, d):
            return self

        def sample(self, *, param_cond, img_cond=None, disable_tqdm=False):
            calls["cond"].append(None if img_cond is None else tuple(img_cond.shape))
            return torch.rand(param_cond.shape[0], 1, S, S)

    def occ(d, m):
        calls["occ"] += 1
        return d, m

    monkeypatch.setattr(geometry, "reproject_tensor", lambda d, K, P, **k: (d.clone(), d > 3))
    monkeypatch.setattr(geometry, "occlusion_filter", occ)
    monkeypatch.setattr(geometry, "point_cloud_batch",
                        lambda d, K, pose=None, scale=10.0, clip=(0.5, 10):
                        (torch.rand(d.shape[0], S * S, 3, dtype=torch.float64), torch.full((d.shape[0],), 17)))
    np.random.seed(0)
    t = tester.Tester(FakeModel(), batch_size=2, results_folder=str(tmp_path / "r"),
                      samples_folder=str(tmp_path / "o"), device="cpu")
    out = t.sample(3, 3)
    assert out.shape == (3, 1, S, 3 * S)
    assert calls["cond"] == [None, (2, 2, S, S), (2, 2, S, S), None, (1, 2, S, S), (1, 2, S, S)]
    assert calls["occ"] == 4                                  # every conditional view: the pose is never zero
    names = sorted(p.name for p in (tmp_path / "o").iterdir())
    assert len(names) == 3 * (1 + 2 * 3) + 1 and "overview.png" in names and "scene-2-sample-2.ply" in names
    assert t.sample_uncondition(4).shape == (4, 1, S, S)


@pytest.mark.parametrize("mode", ["ddnm_none", "ddnm_linear_ddim", "denoise", "no_ddnm_refine_only"])
def test_stepwise_sampler_host_logic(monkeypatch, mode):
    """Keep-mask dropout / denoise / refine-only conditioning drive the sampler step by step from the
    host (diffusion._sample_stepwise).  With the single device step replaced by the oracle's arithmetic
    for that step, the host loop must reproduce the oracle's full loop bit for bit: order of the
    Gaussian and uniform draws, per-step masks, which steps replace."""
    import torch
    from oracle import torch_ref as R
    from pointreggpt_b200 import _ffi, nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    torch.manual_seed(9)
    net = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    SZ, B = 32, 2
    cfg = dict(ddnm_none=dict(timesteps=4, ddnm_sampling_dropout=0.3),
               ddnm_linear_ddim=dict(timesteps=8, sampling_timesteps=3, ddnm_sampling_dropout=0.5,
                                     ddnm_dropout_schedule='linear', ddim_sampling_eta=1.0),
               denoise=dict(timesteps=4, is_ddnm_sampling=False),
               no_ddnm_refine_only=dict(timesteps=3, is_ddnm_sampling=False))[mode]
    diff = GaussianDiffusion(net, image_size=SZ, objective='pred_x0', beta_schedule='sigmoid', **cfg)
    T = cfg["timesteps"]
    sch = R.make_schedule(T)
    g = torch.Generator().manual_seed(10)
    noises = torch.stack([torch.randn(B, 1, SZ, SZ, generator=g) for _ in range(9)])
    uniforms = torch.stack([torch.rand(B, 1, SZ, SZ, generator=g) for _ in range(9)])
    d = torch.rand(B, 1, SZ, SZ, generator=g)
    d[d < 0.3] = 0
    ic = torch.cat([d, (d > 0).float()], 1) * 2 - 1
    pc = torch.tensor([[304., 304., 16.5, 16.], [290., 291., 16.5, 16.]])

    def fake_step(h, cap, st, p, ic_step, buf):
        """The arithmetic of one prg_step, from the oracle's pieces (SDD:1234-1281, 1343-1389)."""
        x, t = buf[0], st.t
        ddim = st.kind in (_ffi.STEP_DDIM, _ffi.STEP_DDIM_LAST, _ffi.STEP_REFINE_DDIM)
        refine = st.kind in (_ffi.STEP_REFINE_P, _ffi.STEP_REFINE_DDIM)
        pred_noise, x0 = R._predictions(sd, sch, x, t, p, ic_step, ddim, refine, None)
        mask = None if ic_step is None else ((ic_step[:, 1:2] + 1) * 0.5) > 0.5
        if st.kind in (_ffi.STEP_P_SAMPLE, _ffi.STEP_REFINE_P):
            x0 = x0.clamp(-1., 1.)
            mean = sch["posterior_mean_coef1"][t] * x0 + sch["posterior_mean_coef2"][t] * x
            sig = (0.5 * sch["posterior_log_variance_clipped"][t]).exp()
            out = mean + sig * (buf[1] if st.add_noise else 0.)
            if refine:
                out = torch.where(mask, out, x)
        elif st.kind == _ffi.STEP_DDIM:
            out = x0 * st.c2 + st.c3 * pred_noise + st.c4 * buf[1]
        elif st.kind == _ffi.STEP_DDIM_LAST:
            out = x0
        else:
            out = torch.where(mask, x0, x)
        return (out + 1) * 0.5 if st.unnormalize else out

    monkeypatch.setattr(_ffi, "require_cuda", lambda *a: None)
    monkeypatch.setattr(GaussianDiffusion, "_run_single_step", lambda self, *a: fake_step(*a))
    monkeypatch.setattr(type(net), "native_handle", lambda self, b, s, dev: (None, 1 << 30), raising=False)
    fn = diff.denoise if mode == "denoise" else diff.sample
    got = fn(param_cond=pc, img_cond=ic, has_refine_step=True, noise=noises, keep_uniform=uniforms)
    ddnm, denoise = R.dropout_tables(T, cfg.get("ddnm_sampling_dropout", 0.), cfg.get("ddnm_dropout_schedule", "none"))
    keep = {"denoise": R.KeepMask("denoise", denoise, list(uniforms)),
            "no_ddnm_refine_only": None}.get(mode, R.KeepMask("ddnm", ddnm, list(uniforms)))
    if mode == "no_ddnm_refine_only":
        # no replacement in the loop; the refine step still uses the mask (SDD:1307-1314)
        img = R.p_sample_loop(sd, sch, pc, None, list(noises))          # unnormalised at the end
        x = img * 2 - 1
        _, x0 = R._predictions(sd, sch, x, 0, pc, ic, False, True, None)
        x0 = x0.clamp(-1., 1.)
        mean = sch["posterior_mean_coef1"][0] * x0 + sch["posterior_mean_coef2"][0] * x
        want = (torch.where(((ic[:, 1:2] + 1) * 0.5) > 0.5, mean, x) + 1) * 0.5
        assert torch.allclose(got, want, atol=1e-6)
        return
    if "sampling_timesteps" in cfg:
        want = R.ddim_sample(sd, sch, pc, ic, list(noises), cfg["sampling_timesteps"], 1.0, has_refine_step=True, keep=keep)
    else:
        want = R.p_sample_loop(sd, sch, pc, ic, list(noises), has_refine_step=True, keep=keep)
    assert torch.allclose(got, want, atol=1e-6)
    assert torch.equal(got, want) or mode == "ddnm_linear_ddim"     # ddim coefficients pass through float()


def test_default_sampling_stays_on_the_fused_device_loop(monkeypatch):
    """The shipped configuration (DDNM, dropout 0) must issue ONE prg_sampler_run covering every
    step; only keep-mask dropout / denoise / refine-only conditioning may take the step-wise loop."""
    import torch
    from pointreggpt_b200 import _ffi, nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    calls = []

    class FakeLib:
        def prg_sampler_run(self, h, arr, nsteps, p, ic, noise, seed, out, b, stream):
            calls.append((nsteps, ic is not None, noise is not None, b))
            return 0

    monkeypatch.setattr(_ffi, "lib", lambda: FakeLib())
    monkeypatch.setattr(_ffi, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_ffi, "stream", lambda ref=None: None)
    torch.manual_seed(0)
    net = nets.Unet(dim=64, param_cond_dim=4)
    monkeypatch.setattr(type(net), "native_handle", lambda self, b, s, dev: (None, 1 << 30), raising=False)
    monkeypatch.setattr(GaussianDiffusion, "_sample_stepwise",
                        lambda self, *a, **k: (_ for _ in ()).throw(AssertionError("step-wise loop used")))
    pc = torch.zeros(3, 4)
    ic = torch.zeros(3, 2, 32, 32)
    for kw, nsteps in ((dict(timesteps=20), 20), (dict(timesteps=20, sampling_timesteps=5), 5)):
        d = GaussianDiffusion(net, image_size=32, objective="pred_x0", beta_schedule="sigmoid", **kw)
        calls.clear()
        d.sample(param_cond=pc, img_cond=ic)
        d.sample(param_cond=pc, img_cond=ic, has_refine_step=True)
        d.sample(param_cond=pc)
        assert calls == [(nsteps, True, False, 3), (nsteps + 1, True, False, 3), (nsteps, False, False, 3)]
    # a model built without DDNM ignores the condition in the loop: still the fused path
    d = GaussianDiffusion(net, image_size=32, objective="pred_x0", timesteps=6, is_ddnm_sampling=False)
    calls.clear()
    d.sample(param_cond=pc, img_cond=ic)
    assert calls == [(6, False, False, 3)]


def test_res1x1_gn_fragment_permutation_is_a_gemm():
    """k_res1x1_gn (elementwise.cu) feeds mma.sync.m16n8k16 straight from global memory by permuting the k
    index of the MMA and the n index of an n-tile so that every lane touches 16 contiguous bytes of a pixel
    row.  This restates the kernel's lane -> element assignment against the PTX fragment layout and checks
    that the accumulators the lanes end up with are the plain GEMM y[px][co] = sum_ci x[px][ci] w[co][ci],
    with lane (g, t) holding channels 8t .. 8t+7 of rows g and g + 8 of each 32-channel group."""
    rng = np.random.default_rng(0)
    cin, cout = 128, 64
    x = rng.integers(-3, 4, (16, cin)).astype(np.float64)        # 16 pixels of one warp tile
    w = rng.integers(-3, 4, (cout, cin)).astype(np.float64)
    want = x @ w.T

    def mma(a_frag, b_frag):
        """m16n8k16 from per-lane fragments: a_frag[lane] = 8 values (a0.lo, a0.hi, a1.., a2.., a3..),
        b_frag[lane] = 4 values (b0.lo, b0.hi, b1.lo, b1.hi) -> c[lane] = (c0, c1, c2, c3)."""
        A = np.zeros((16, 16)); B = np.zeros((16, 8))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for e in range(2):
                A[g, 2 * t + e] = a_frag[lane][0 + e]            # a0: row g,     k 2t, 2t+1
                A[g + 8, 2 * t + e] = a_frag[lane][2 + e]        # a1: row g + 8, k 2t, 2t+1
                A[g, 2 * t + 8 + e] = a_frag[lane][4 + e]        # a2: row g,     k 2t+8, 2t+9
                A[g + 8, 2 * t + 8 + e] = a_frag[lane][6 + e]    # a3: row g + 8, k 2t+8, 2t+9
                B[2 * t + e, g] = b_frag[lane][0 + e]            # b0: k 2t, 2t+1,   n g
                B[2 * t + 8 + e, g] = b_frag[lane][2 + e]        # b1: k 2t+8, 2t+9, n g
        C = A @ B
        return [(C[l >> 2, 2 * (l & 3)], C[l >> 2, 2 * (l & 3) + 1], C[(l >> 2) + 8, 2 * (l & 3)],
                 C[(l >> 2) + 8, 2 * (l & 3) + 1]) for l in range(32)]

    got = np.zeros((16, cout))
    for ng in range(cout // 32):                                  # 32-channel output groups
        for jj in range(4):                                       # n-tiles of the group
            acc = [np.zeros(4) for _ in range(32)]
            for kg in range(cin // 32):                           # 32-channel input groups
                for s in range(2):                                # the two k-steps one LDG.128 / LDS.128 feeds
                    a_frag, b_frag = [], []
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        xa = x[g, 32 * kg + 8 * t:32 * kg + 8 * t + 8]          # the lane's 16 bytes of row g
                        xb = x[g + 8, 32 * kg + 8 * t:32 * kg + 8 * t + 8]      # ... and of row g + 8
                        # f0 = {xa.x, xb.x, xa.y, xb.y}, f1 = {xa.z, xb.z, xa.w, xb.w} (32-bit words = channel pairs)
                        lo = 4 * s
                        a_frag.append([xa[lo], xa[lo + 1], xb[lo], xb[lo + 1], xa[lo + 2], xa[lo + 3], xb[lo + 2], xb[lo + 3]])
                        row = 32 * ng + 8 * (g >> 1) + 2 * jj + (g & 1)          # weight row of this lane's B fragment
                        wv = w[row, 32 * kg + 8 * t:32 * kg + 8 * t + 8]
                        b_frag.append([wv[lo], wv[lo + 1], wv[lo + 2], wv[lo + 3]])
                    for lane, c in enumerate(mma(a_frag, b_frag)):
                        acc[lane] += np.array(c)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                ch = 32 * ng + 8 * t + 2 * jj                     # the epilogue's channel of acc[..][0] (and + 1)
                got[g, ch], got[g, ch + 1], got[g + 8, ch], got[g + 8, ch + 1] = acc[lane]
    assert np.array_equal(got, want)

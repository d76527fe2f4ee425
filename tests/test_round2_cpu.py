"""Host-side logic added in round 2: per-scene seeds, checkpoint loading, PLY parsing, handle copies,
the CLI's argument surface."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from pointreggpt_b200 import cloud, geometry, nets, rng
from pointreggpt_b200.diffusion import GaussianDiffusion
from pointreggpt_b200.generator import Generator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scene_seeds_do_not_depend_on_sharding():
    """Scene k's key is a function of (base, k, sample) only: any split of the range over ranks or
    batches hands the same keys to the same scenes."""
    from pointreggpt_b200 import dist as pdist
    base = 7
    want = {k: rng.scene_seed(base, k, 0) for k in range(10, 31)}
    assert len(set(want.values())) == len(want)
    for world in (1, 2, 3, 8):
        got = {}
        for r in range(world):
            lo, hi = pdist.shard_range(10, 31, r, world)
            for bs in (1, 4):
                for b_idx, batch in enumerate(geometry.num_to_groups(hi - lo, bs)):
                    first = lo + b_idx * bs
                    for s in range(batch):
                        got[first + s] = rng.scene_seed(base, first + s, 0)
        assert got == want
    # different base / sample index / scene -> different keys; keys use all 64 bits
    assert rng.scene_seed(8, 10, 0) != want[10] and rng.scene_seed(7, 10, 1) != want[10]
    assert max(want.values()) > 2 ** 60
    # poses: same generator -> same pose, independent of the global numpy RNG state
    np.random.seed(1)
    a = geometry.random_sample_pose(1, rng=rng.scene_rng(base, 12, 0))
    np.random.seed(2)
    b = geometry.random_sample_pose(1, rng=rng.scene_rng(base, 12, 0))
    c = geometry.random_sample_pose(1, rng=rng.scene_rng(base, 13, 0))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.dtype == np.float32 and a.shape == (1, 4, 4)
    R = a[0, :3, :3].astype(np.float64)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)


def test_image_seeds_single_integer_is_position_keyed():
    s5 = rng.image_seeds(123, 5)
    assert s5[:3] == rng.image_seeds(123, 3) and len(set(s5)) == 5
    assert GaussianDiffusion._image_seeds([4, 5, 6], 3) == [4, 5, 6]
    assert GaussianDiffusion._image_seeds(torch.tensor([4, 5]), 2) == [4, 5]
    with pytest.raises(AssertionError):
        GaussianDiffusion._image_seeds([1, 2], 3)


def _small_diffusion():
    torch.manual_seed(0)
    unet = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2), channels=1)
    return GaussianDiffusion(unet, image_size=32, timesteps=10, sampling_timesteps=5, objective="pred_x0",
                             beta_schedule="sigmoid", ddim_sampling_eta=1.0)


def test_generator_load_reference_checkpoint_layout(tmp_path):
    """SDD:2307-2324 / 1681-1717: {'step','model','opt','ema','scaler','version'}; the 'ema' entry is
    ema_pytorch's state dict (online_model.*, ema_model.*, initted, step).  Generation samples from the
    EMA copy (SDD:2572)."""
    diff = _small_diffusion()
    online = {k: v.clone() for k, v in diff.state_dict().items()}
    ema = {k: (v + 0.25 if v.dtype.is_floating_point and k.startswith("model.") else v.clone())
           for k, v in online.items()}
    ckpt = {"step": 7, "model": online, "opt": {}, "scaler": None, "version": "1.5.4",
            "ema": dict({"online_model." + k: v for k, v in online.items()},
                        **{"ema_model." + k: v for k, v in ema.items()},
                        initted=torch.tensor(True), step=torch.tensor(7))}
    res = tmp_path / "res"
    res.mkdir()
    torch.save(ckpt, str(res / "model-official.pt"))
    fresh = _small_diffusion()
    with torch.no_grad():
        for p in fresh.parameters():
            p.add_(1.0)                      # make sure load() really overwrites
    gen = Generator(fresh, "synthetic", batch_size=2, results_folder=str(res),
                    samples_folder=str(tmp_path / "out"), device="cpu")
    gen.load("official")
    got_online = gen.model.state_dict()
    got_ema = gen.ema.ema_model.state_dict()
    assert gen.ema.ema_model is not gen.model
    for k in online:
        assert torch.equal(got_online[k], online[k]), k
        assert torch.equal(got_ema[k], ema[k]), k
    assert len(got_ema) == len(online)
    # a checkpoint without ema_model.* keys is rejected, not silently ignored
    bad = dict(ckpt, ema={"initted": torch.tensor(True)})
    torch.save(bad, str(res / "model-bad.pt"))
    with pytest.raises(KeyError):
        gen.load("bad")
    # missing file: the reference fails at torch.load as well
    with pytest.raises(FileNotFoundError):
        gen.load("nope")


def test_generator_rejects_missing_data_root(tmp_path):
    gen = Generator(_small_diffusion(), str(tmp_path / "no_such_tree"), batch_size=2,
                    results_folder=str(tmp_path / "r"), samples_folder=str(tmp_path / "o"), device="cpu")
    # generate() selects its CUDA device first; the missing-tree check fires before any device work
    import unittest.mock as mock
    with mock.patch("torch.cuda.set_device"), pytest.raises(FileNotFoundError):
        gen.generate(0, 1, 1)


def test_native_net_deepcopy_and_pickle_drop_the_handle():
    import ctypes
    import pickle
    torch.manual_seed(0)
    net = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2), channels=1)
    net._handle = ctypes.c_void_p(0)          # a ctypes pointer: deepcopy/pickle of it would raise
    net._handle_key = (1, 2, 32, 0)
    net._blob, net._blob_sig = b"x", ("sig",)
    dup = copy.deepcopy(net)
    assert dup._handle is None and dup._blob is None and dup._handle_key is None
    for (k, a), (_, b) in zip(net.state_dict().items(), dup.state_dict().items()):
        assert torch.equal(a, b) and a.data_ptr() != b.data_ptr(), k
    diff = GaussianDiffusion(net, image_size=32, timesteps=4, objective="pred_x0")
    assert copy.deepcopy(diff).model._handle is None
    back = pickle.loads(pickle.dumps(net))
    assert back._handle is None and torch.equal(back.init_conv.weight, net.init_conv.weight)
    net._handle = None                         # nothing real to destroy
    net.invalidate()
    assert net._blob is None


def test_broadcast_invalidates_packed_networks(monkeypatch):
    """a network packed before the broadcast must re-pack afterwards."""
    import torch.distributed as dist
    from pointreggpt_b200 import dist as pdist
    torch.manual_seed(0)
    net = nets.MaskUnet(dim=64, dim_mults=(1, 2))
    net._blob, net._blob_sig = b"stale", net._signature()
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda: 2)
    monkeypatch.setattr(dist, "broadcast", lambda t, src=0: t.mul_(0).add_(3.0))
    n = pdist.broadcast_weights([net], src=0)
    assert n > 0 and net._blob is None
    assert all(bool((p == 3).all()) for p in net.parameters())
    assert all(p._version > 0 for p in net.parameters())


def test_read_ply_parses_the_header(tmp_path):
    pts = np.random.default_rng(0).uniform(-1, 1, (17, 3))
    p0 = str(tmp_path / "a.ply")
    cloud.write_ply(p0, pts)
    assert np.array_equal(cloud.read_ply(p0), pts)
    # float positions + normals + colours in another order (what other writers produce)
    dt = np.dtype([("nx", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("ny", "<f4")])
    rec = np.zeros(17, dt)
    for i, k in enumerate("xyz"):
        rec[k] = pts[:, i].astype(np.float32)
    head = ("ply\nformat binary_little_endian 1.0\ncomment x\nelement vertex 17\nproperty float nx\n"
            "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty float ny\n"
            "element face 0\nproperty list uchar int vertex_indices\nend_header\n")
    p1 = str(tmp_path / "b.ply")
    with open(p1, "wb") as f:
        f.write(head.encode() + rec.tobytes())
    assert np.array_equal(cloud.read_ply(p1), pts.astype(np.float32).astype(np.float64))
    # ascii
    p2 = str(tmp_path / "c.ply")
    with open(p2, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 3\nproperty double x\nproperty double y\nproperty double z\nend_header\n")
        f.write("1 2 3\n4 5 6\n7 8 9.5\n")
    assert np.array_equal(cloud.read_ply(p2), [[1, 2, 3], [4, 5, 6], [7, 8, 9.5]])
    # refused instead of misread
    for bad in ("ply\nformat binary_big_endian 1.0\nelement vertex 1\nproperty double x\nproperty double y\nproperty double z\nend_header\n",
                "ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty double x\nproperty double y\nend_header\n",
                "ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty double x\nproperty double y\nproperty double z\nend_header\n",
                "nope\n"):
        p3 = str(tmp_path / "bad.ply")
        with open(p3, "wb") as f:
            f.write(bad.encode() + b"\0" * 24)
        with pytest.raises(ValueError):
            cloud.read_ply(p3)


def test_generate_dataset_cli_surface(tmp_path):
    """GD:6-30: --resume is required; -start/-stop/--num_samples/--dataset_name keep their names and
    defaults; a data root that does not exist is an error, not a silent switch to synthetic frames."""
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    run = lambda *a: subprocess.run([sys.executable, os.path.join(ROOT, "generate_dataset.py"), *a],
                                    cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    r = run()
    assert r.returncode == 2 and "--resume" in r.stderr
    r = run("--help")
    assert r.returncode == 0
    for flag in ("--resume", "--dataset_name", "--start_scene_index", "-start", "--stop_scene_index", "-stop",
                 "--num_samples", "--synthetic", "--seed", "--batch_size", "--device_batch"):
        assert flag in r.stdout, flag
    r = run("--resume", "official", "--data_root", str(tmp_path / "missing"))
    assert r.returncode != 0 and "does not exist" in r.stderr


def test_reprojection_ticket_protocol_is_deadlock_free_and_slot_safe():
    """Model of k_reproject_fused's scheduling (geometry.cu, RpItem): tickets are claimed in increasing order by
    whichever CTA is free; ticket p = (round k, item j) splats item j of map k (k < B) and finalises item j of map
    k - D (k >= D); it may start once done[k - D] == items; a finished item bumps done[k].  With R = 2 D ring slots:
    whatever the interleaving, every item runs, a slot is never splatted before its previous map is fully finalised
    and never finalised before its own map is fully splatted."""
    import random
    for trial in range(200):
        rnd = random.Random(trial)
        B, items, D = rnd.randint(1, 9), rnd.randint(1, 5), rnd.randint(1, 4)
        D = min(D, B)
        R = 2 * D
        total = (B + D) * items
        ncta = rnd.randint(1, 7)
        done = [0] * (B + D)
        splat_left = {m: items for m in range(B)}      # items of map m still to splat
        fin_left = {m: items for m in range(B)}        # items of map m still to finalise
        next_ticket = 0
        held = [None] * ncta                            # the ticket a CTA has claimed and not finished
        finished = 0
        steps = 0
        while finished < total:
            steps += 1
            assert steps < 100000, "no progress: deadlock"
            c = rnd.randrange(ncta)
            if held[c] is None:
                if next_ticket < total:
                    held[c] = next_ticket
                    next_ticket += 1
                continue
            k, j = divmod(held[c], items)
            if k >= D and done[k - D] < items:
                # blocked -- but then some smaller ticket is unfinished and runnable (checked globally below)
                smaller = [h for h in held if h is not None and h < held[c]]
                assert smaller or any(d < items for d in done[:k - D + 1])
                continue
            if k < B:                                   # splat half: slot k % R must be free of map k - R
                assert k - R < 0 or fin_left[k - R] == 0, "slot reused before its map was finalised"
                splat_left[k] -= 1
            if k >= D:                                  # finalise half: map k - D fully splatted
                assert splat_left[k - D] == 0, "finalised before every pixel was splatted"
                fin_left[k - D] -= 1
            done[k] += 1
            held[c] = None
            finished += 1
        assert all(v == 0 for v in splat_left.values()) and all(v == 0 for v in fin_left.values())
        assert all(d == items for d in done)

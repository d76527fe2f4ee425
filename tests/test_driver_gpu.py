"""Generator / generate_dataset drop-in on the GPU with synthetic source frames."""
import numpy as np
import pytest
import torch

from pointreggpt_b200 import cloud, nets
from pointreggpt_b200.diffusion import GaussianDiffusion
from pointreggpt_b200.generator import Generator

pytestmark = pytest.mark.gpu


def test_generator_writes_reference_layout_and_resumes(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    unet = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    diff = GaussianDiffusion(unet, image_size=128, timesteps=8, sampling_timesteps=2, objective="pred_x0",
                             beta_schedule="sigmoid", ddim_sampling_eta=1.0)
    gen = Generator(diff, "synthetic", batch_size=2, results_folder=str(tmp_path / "res"),
                    samples_folder=str(tmp_path / "ds" / "data"))
    mask = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
    with torch.no_grad():
        mask.final_conv[0].bias.fill_(8.0)
    np.random.seed(0)
    n = gen.generate(start_scene_index=5, stop_scene_index=8, num_samples=1, has_refine_step=False,
                     depth_correction=mask)
    assert n == 3
    for idx in (5, 6, 7):
        d = tmp_path / "ds" / "data" / ("scene-%06d" % idx)
        for f in ("camera-intrinsics.txt", "sample-000000.image.png", "sample-000000.cloud.ply",
                  "reprojected.image.png", "corrected.image.png", "sample-000001.pose.txt",
                  "sample-000001.image.png", "sample-000001.depth.png", "sample-000001.cloud.ply"):
            assert (d / f).is_file(), f
        src = cloud.read_ply(str(d / "sample-000000.cloud.ply"))
        assert src.shape[0] > 100 and np.isfinite(src).all()
        assert src[:, 2].min() >= 0.5 - 1e-6 and src[:, 2].max() <= 3.5 + 1e-6
        pose = np.loadtxt(str(d / "sample-000001.pose.txt"))
        assert pose.shape == (4, 4)
    # resume: finished batches are skipped (SDD:2371-2381)
    stamp = (tmp_path / "ds" / "data" / "scene-000005" / "sample-000001.cloud.ply").stat().st_mtime
    gen.generate(start_scene_index=5, stop_scene_index=8, num_samples=1, has_refine_step=False,
                 depth_correction=mask)
    assert (tmp_path / "ds" / "data" / "scene-000005" / "sample-000001.cloud.ply").stat().st_mtime == stamp


def test_voxel_down_sample_centroids():
    pts = torch.tensor([[0.0, 0, 0], [0.01, 0, 0], [1.0, 1, 1]], dtype=torch.float64).cuda()
    out = cloud.voxel_down_sample(pts, 0.5).cpu().numpy()
    assert out.shape == (2, 3)
    assert np.allclose(sorted(out[:, 0]), [0.005, 1.0])


def test_generate_dataset_cli_end_to_end(tmp_path):
    """The drop-in script itself (GD:1-63): reference flags, 256x256, T = 1000 with a DDIM schedule (cut
    to 3 steps here), files under ./{dataset_name}/data/scene-%06d/; a second invocation over a shifted
    range (the reference's way to parallelise) reproduces the overlapping scene byte for byte."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    base = [sys.executable, os.path.join(root, "generate_dataset.py"), "--resume", "none", "--random_init",
            "--synthetic", "--sampling_timesteps", "3"]
    r = subprocess.run(base + ["--dataset_name", "ds_a", "-start", "3", "-stop", "5"], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "generated 2 scenes" in r.stdout
    for idx in (3, 4):
        d = tmp_path / "ds_a" / "data" / ("scene-%06d" % idx)
        for f in ("camera-intrinsics.txt", "sample-000000.image.png", "sample-000000.cloud.ply",
                  "reprojected.image.png", "corrected.image.png", "sample-000001.pose.txt",
                  "sample-000001.image.png", "sample-000001.depth.png", "sample-000001.cloud.ply"):
            assert (d / f).is_file(), f
    r = subprocess.run(base + ["--dataset_name", "ds_b", "-start", "4", "-stop", "5", "--batch_size", "1"],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for f in ("sample-000001.cloud.ply", "sample-000001.depth.png", "sample-000001.pose.txt"):
        a = (tmp_path / "ds_a" / "data" / "scene-000004" / f).read_bytes()
        b = (tmp_path / "ds_b" / "data" / "scene-000004" / f).read_bytes()
        assert a == b, f

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

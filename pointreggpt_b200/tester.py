"""`Tester` under the reference's name (SDD:1829-2247), the parts on the generation path:
`load`, `sample_uncondition`, `sample` -- successive conditional generation: an unconditional first
view, then each next view conditioned on the previous one reprojected 0.5 m forward, with the
occlusion filter (SDD:1961-2065) -- and `generate` -- scene fusion: every next view is conditioned on
the z-buffer of the voxel-merged cloud of all previous views seen from a random rotation
(SDD:2099-2247).  FID / Inception scoring is not part of the path.

Differences forced by the offline environment: accelerate / ema_pytorch replaced as in
`Generator`; PNGs are written with PIL (8-bit grey) instead of matplotlib; PLYs by `cloud.write_ply`.
"""
import math
import os
import shutil
from pathlib import Path

import numpy as np
import torch

from . import cloud, geometry
from .generator import _EmaHolder

STEP_FORWARD = (0.0, 0.0, 0.5)       # SDD:2024: each new view is 0.5 m further along the optical axis
CLOUD_CLIP = (0.5, 3.5)              # SDD:2003-2005


def _save_grey(path, img01):
    from PIL import Image
    a = np.clip(np.asarray(img01, dtype=np.float32), 0.0, 1.0)
    Image.fromarray(np.rint(a * 255.0).astype(np.uint8), mode="L").save(path)


class Tester(object):
    def __init__(self, diffusion_model, folder=None, *, batch_size=16, ema_update_every=10,
                 ema_decay=0.995, results_folder='./results', samples_folder='./samples',
                 amp=False, fp16=False, split_batches=True, device=None, **unused):
        super().__init__()
        if amp or fp16:
            raise NotImplementedError("the native path has its own fixed fp16-operand numerics")
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        self._device = torch.device(device)
        if self._device.type == "cuda":
            torch.cuda.set_device(self._device)
        self.model = diffusion_model.to(self._device)
        self.channels = diffusion_model.channels
        self.batch_size = batch_size
        self.image_size = diffusion_model.image_size
        self.ema = _EmaHolder(self.model)
        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(parents=True, exist_ok=True)
        self.samples_folder = Path(samples_folder)
        if self.samples_folder.exists():                  # SDD:1884-1886: a fresh folder per run
            shutil.rmtree(str(self.samples_folder))
        self.samples_folder.mkdir(parents=True, exist_ok=True)

    @property
    def device(self):
        return self._device

    def load(self, milestone):
        data = torch.load(str(self.results_folder / f'model-{milestone}.pt'), map_location="cpu",
                          weights_only=False)
        self.model.load_state_dict(data['model'])
        self.ema.load_state_dict(data['ema'])
        if 'version' in data:
            print(f"loading from version {data['version']}")

    def _intrinsics(self, batch):
        return geometry.intrinsic_transform(geometry.random_sample_intrinsic(batch_size=batch),
                                            resize=self.image_size,
                                            centercrop=self.image_size).astype(np.float32)

    @torch.no_grad()
    def sample_uncondition(self, num_samples=25):
        """SDD:1935-1959: unconditional depth maps, saved as one grid image."""
        from torchvision import utils
        out = []
        for batch in geometry.num_to_groups(num_samples, self.batch_size):
            K = self._intrinsics(batch)
            out.append(self.ema.ema_model.sample(
                param_cond=geometry.param_vector(torch.tensor(K).to(self.device)), disable_tqdm=True))
        images = torch.cat(out, dim=0)
        utils.save_image(images, str(self.samples_folder / 'unconditional.png'),
                         nrow=int(math.sqrt(num_samples)))
        return images

    def _save_view(self, scene, sample_idx, last, rpj, image, K, absolute_pose):
        stem = self.samples_folder / f'scene-{scene}-sample-{sample_idx}'
        _save_grey(str(stem) + '.png', np.concatenate([last, rpj, image], axis=-1))
        # point_cloud(image * 10, K, clip) and, for the later views, (pc - t) @ R  (SDD:2003, 2072-2074)
        pc, counts = geometry.point_cloud_batch(torch.as_tensor(image)[None, None].to(self.device),
                                                torch.tensor(K[None]).to(self.device),
                                                pose=None if absolute_pose is None else
                                                torch.tensor(absolute_pose[None]).to(self.device),
                                                scale=10.0, clip=CLOUD_CLIP)
        cloud.write_ply(str(stem) + '.ply', pc[0, :int(counts[0])])

    @torch.no_grad()
    def sample(self, num_scenes, num_samples):
        """SDD:1961-2065.  Returns the (num_scenes, 1, S, num_samples * S) strip of all views."""
        model = self.ema.ema_model
        dev = self.device
        strips = []
        for b_idx, batch in enumerate(geometry.num_to_groups(num_scenes, self.batch_size)):
            K = self._intrinsics(batch)
            Kt = torch.tensor(K).to(dev)
            param_cond = geometry.param_vector(Kt)
            absolute_pose = np.stack([np.eye(4) for _ in range(batch)]).astype(np.float32)
            images = model.sample(param_cond=param_cond, disable_tqdm=True)
            views = [images]
            zero = np.zeros((self.image_size, self.image_size), np.float32)
            for s, image in enumerate(images):
                scene = b_idx * self.batch_size + s
                self._save_view(scene, 0, zero, zero, image[0].cpu().numpy(), K[s], None)
                np.savetxt(str(self.samples_folder / f'scene-{scene}-camera-intrinsics.txt'), K[s])
            for sample_idx in range(1, num_samples):
                relative_pose = np.stack([np.eye(4) for _ in range(batch)])
                relative_pose[..., :3, 3] = np.array(STEP_FORWARD)
                relative_pose = relative_pose.astype(np.float32)
                absolute_pose = relative_pose @ absolute_pose
                images_rpj, mask_rpj = geometry.reproject_tensor(images * 10, Kt,
                                                                 torch.tensor(relative_pose).to(dev))
                if np.sum(absolute_pose[..., :3, 3] ** 2) != 0:                    # SDD:2035-2037
                    images_rpj, mask_rpj = geometry.occlusion_filter(images_rpj, mask_rpj)
                images_rpj = images_rpj * 0.1
                img_cond = geometry.normalize_to_neg_one_to_one(
                    torch.cat([images_rpj, mask_rpj.to(images_rpj.dtype)], dim=1))
                images_last = images
                images = model.sample(param_cond=param_cond, img_cond=img_cond, disable_tqdm=True)
                views.append(images)
                for s, image in enumerate(images):
                    self._save_view(b_idx * self.batch_size + s, sample_idx,
                                    images_last[s, 0].cpu().numpy(), images_rpj[s, 0].cpu().numpy(),
                                    image[0].cpu().numpy(), K[s], absolute_pose[s])
            strips.append(torch.cat(views, dim=-1))
        all_images = torch.cat(strips, dim=0)
        _save_grey(str(self.samples_folder / 'overview.png'),
                   torch.cat(list(all_images[:, 0]), dim=0).cpu().numpy())
        return all_images


    @torch.no_grad()
    def generate(self, num_scenes, num_samples, voxel_size=0.005):
        """SDD:2099-2247: scene fusion.  View 0 is unconditional; its cloud (clip 0.5-3.5 m, voxel
        `voxel_size`) seeds the scene memory.  For every further view a random in-view rotation is
        composed onto the camera pose, the scene memory is z-buffered into that camera
        (`pc2depth_tensor`, pose applied in the kernel), the model samples with that condition, and the
        new view's cloud -- moved back to the first camera, `(pc - t) @ R` -- is merged into the memory
        by voxel down-sampling.  Per scene a 25 mm cloud `scene-{i}.ply` is written at the end.
        Returns the (num_scenes, 1, S, num_samples * S) strip of all views."""
        model = self.ema.ema_model
        dev = self.device
        S = self.image_size
        strips = []
        for b_idx, batch in enumerate(geometry.num_to_groups(num_scenes, self.batch_size)):
            K = self._intrinsics(batch)
            Kt = torch.tensor(K).to(dev)
            param_cond = geometry.param_vector(Kt)
            absolute_pose = np.stack([np.eye(4) for _ in range(batch)]).astype(np.float32)
            images = model.sample(param_cond=param_cond, disable_tqdm=True)
            views = [images]
            zero = np.zeros((S, S), np.float32)
            pc0, counts = geometry.point_cloud_batch(images, Kt, scale=10.0, clip=CLOUD_CLIP)
            counts = counts.cpu().tolist()
            memory = []
            for s in range(batch):
                _save_grey(str(self.samples_folder / f'scene-{b_idx * self.batch_size + s}-sample-0.png'),
                           np.concatenate([zero, zero, images[s, 0].cpu().numpy()], axis=-1))
                memory.append(cloud.voxel_down_sample(pc0[s, :counts[s]], voxel_size).to(torch.float32))
            for sample_idx in range(1, num_samples):
                relative_pose = geometry.random_sample_transform(K, image_size=S)          # SDD:2156-2158
                absolute_pose = relative_pose @ absolute_pose
                pose_t = torch.tensor(absolute_pose).to(dev)
                offsets = torch.tensor(np.cumsum([0] + [m.shape[0] for m in memory]))
                images_rpj, mask_rpj = geometry.pc2depth_ragged(torch.cat(memory), offsets, Kt,
                                                                image_size=[S, S], pose=pose_t)
                images_rpj = images_rpj * 0.1
                img_cond = geometry.normalize_to_neg_one_to_one(
                    torch.cat([images_rpj, mask_rpj.to(images_rpj.dtype)], dim=1))
                images_last = images
                images = model.sample(param_cond=param_cond, img_cond=img_cond, disable_tqdm=True)
                views.append(images)
                pc_new, counts = geometry.point_cloud_batch(images, Kt, pose=pose_t, scale=10.0,
                                                            clip=CLOUD_CLIP)
                counts = counts.cpu().tolist()
                for s in range(batch):
                    _save_grey(str(self.samples_folder /
                                   f'scene-{b_idx * self.batch_size + s}-sample-{sample_idx}.png'),
                               np.concatenate([images_last[s, 0].cpu().numpy(), images_rpj[s, 0].cpu().numpy(),
                                               images[s, 0].cpu().numpy()], axis=-1))
                    merged = torch.cat([memory[s].to(torch.float64), pc_new[s, :counts[s]]], dim=0)
                    memory[s] = cloud.voxel_down_sample(merged, voxel_size).to(torch.float32)
            for s in range(batch):
                cloud.write_ply(str(self.samples_folder / f'scene-{b_idx * self.batch_size + s}.ply'),
                                cloud.voxel_down_sample(memory[s], 0.025))
            strips.append(torch.cat(views, dim=-1))
        all_images = torch.cat(strips, dim=0)
        _save_grey(str(self.samples_folder / 'overview.png'),
                   torch.cat(list(all_images[:, 0]), dim=0).cpu().numpy())
        return all_images

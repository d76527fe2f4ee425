"""`Generator` under the reference's name and signature (SDD:2250-2694): the driver of
`generate_dataset.py`.  Per batch of scenes: source frame -> cloud -> random pose -> z-buffer
reprojection -> depth correction -> DDNM sampling -> depth correction -> cloud -> fuse /
voxel-downsample -> files, with the reference's file layout and skip-if-done resume.

Differences, all forced by the offline environment or by the B200 design:
  * accelerate / ema_pytorch / open3d are replaced by a few lines each (device pick, EMA
    weight holder, pointreggpt_b200.cloud);
  * the scene range is sharded across ranks (pointreggpt_b200.dist) instead of every rank
    generating the same range;
  * source frames come from the 3DMatch tree when `folder` exists, otherwise from the seeded
    synthetic generator (`folder="synthetic"`), since no dataset can be downloaded here.
"""
import copy
import os
import pickle
import shutil
from pathlib import Path

import numpy as np
import torch

from . import cloud, dist as pdist, geometry, pipeline, synthetic


class Generator(object):
    def __init__(self, diffusion_model, folder, *, batch_size=16, ema_update_every=10,
                 ema_decay=0.995, results_folder='./results', samples_folder='./samples',
                 amp=False, fp16=False, split_batches=True, device=None):
        super().__init__()
        if amp or fp16:
            raise NotImplementedError("the reference runs generation with amp=False (GD:54); the "
                                      "native path has its own fixed fp16-operand numerics")
        self.rank, self.world_size = pdist.world()
        if device is None:
            local = int(os.environ.get("LOCAL_RANK", "0"))
            device = torch.device("cuda", local)
        self._device = torch.device(device)
        self.folder = folder
        self.model = diffusion_model.to(self._device)
        self.channels = diffusion_model.channels
        self.batch_size = batch_size
        self.image_size = diffusion_model.image_size
        # ema_pytorch.EMA keeps an averaged copy as `ema_model`; generation samples from it
        # (SDD:2572).  Only the holder is needed here.
        self.ema = _EmaHolder(self.model)
        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(parents=True, exist_ok=True)
        self.samples_folder = Path(samples_folder)
        self.samples_folder.mkdir(parents=True, exist_ok=True)
        self.depth_correction = None

    @property
    def device(self):
        return self._device

    def load(self, milestone):
        """SDD:2307-2324: 'model' -> live model, 'ema' -> the EMA copy."""
        path = str(self.results_folder / f'model-{milestone}.pt')
        data = torch.load(path, map_location="cpu", weights_only=False)
        self.model.load_state_dict(data['model'])
        self.ema.load_state_dict(data['ema'])
        if 'version' in data:
            print(f"loading from version {data['version']}")

    # ------------------------------------------------------------------ source frames
    def _source_frame(self, abs_scene_idx, info_train):
        """(depth01 (1,S,S) f32 in units of 10 m, K (3,3) f32) of the scene's source frame."""
        S = self.image_size
        if info_train is None:
            d = synthetic.synthetic_depth(abs_scene_idx, S, S)[None]
            K = synthetic.synthetic_intrinsics(1, S, seed=abs_scene_idx)[0]
            return d, K
        from PIL import Image
        from torchvision import transforms as T
        n = len(info_train['src'])
        key = 'src' if (abs_scene_idx // n) % 2 == 0 else 'tgt'          # SDD:2397-2410
        src_path = os.path.join("./dataset/indoor/data", info_train[key][abs_scene_idx % n])
        with open(src_path.replace(".pth", ".info.txt"), "r") as f:
            scene_name, seq_name, frame_start_idx, _ = f.readline().removesuffix("\n").split()
        scene_path = os.path.join(self.folder, scene_name)
        K = geometry.intrinsic_transform(np.loadtxt(os.path.join(scene_path, "camera-intrinsics.txt")),
                                         resize=S, centercrop=S).astype(np.float32)
        tf = T.Compose([T.Resize(S, interpolation=T.InterpolationMode.NEAREST), T.CenterCrop(S),
                        T.ToTensor()])
        frame = os.path.join(scene_path, seq_name, "frame-{:0>6d}.depth.png".format(int(frame_start_idx)))
        d = tf(Image.open(frame)) * 1e-4                                   # SDD:2458-2459
        d[d > 1] = 0
        return d.to(torch.float32), K

    # ------------------------------------------------------------------ generate
    @torch.no_grad()
    def generate(self, start_scene_index, stop_scene_index, num_samples, memory_voxel_size=0.002,
                 save_voxel_size=0.025, has_refine_step=True, depth_correction=None,
                 write_images=True):
        from torchvision import utils
        dev = self.device
        model = self.ema.ema_model
        if depth_correction is not None:
            self.depth_correction = depth_correction.to(dev)
            ckpt_path = './depth_correction_results/model-best.pt'
            if os.path.isfile(ckpt_path):
                ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
                self.depth_correction.load_state_dict(ckpt['model'])           # SDD:2343-2345
            else:
                print("depth-correction checkpoint %s not found: using the module's current "
                      "weights" % ckpt_path)
        info_train = None
        if self.folder != "synthetic" and os.path.isdir(str(self.folder)):
            with open("./dataset/indoor/metadata/train_info.pkl", 'rb') as f:    # SDD:2352-2354
                info_train = pickle.load(f)
        lo, hi = pdist.shard_range(start_scene_index, stop_scene_index, self.rank, self.world_size)
        done = 0
        for b_idx, batch in enumerate(geometry.num_to_groups(hi - lo, self.batch_size)):
            first = lo + b_idx * self.batch_size
            last_ply = self.samples_folder / 'scene-{:0>6d}/sample-{:0>6d}.cloud.ply'.format(
                first + batch - 1, num_samples // 2)
            if os.path.isfile(last_ply):                                          # SDD:2371-2381
                print("Skip completed scene {:0>6d} - {:0>6d}.".format(first, first + batch - 1))
                done += batch
                continue
            depths, Ks, scene_pcs = [], [], []
            for s in range(batch):
                idx = first + s
                sdir = self.samples_folder / 'scene-{:0>6d}'.format(idx)
                if sdir.exists():
                    shutil.rmtree(str(sdir), ignore_errors=True)
                sdir.mkdir(parents=True, exist_ok=True)
                d, K = self._source_frame(idx, info_train)
                np.savetxt(str(sdir / 'camera-intrinsics.txt'), K)
                if write_images:
                    utils.save_image(d, str(sdir / 'sample-{:0>6d}.image.png'.format(0)))
                depths.append(d)
                Ks.append(K)
            depth01 = torch.stack(depths).to(dev)
            K = torch.tensor(np.stack(Ks)).to(dev)
            # source clouds (SDD:2479-2500): f32 cloud cropped to the room box, 25 mm copy on disk
            pc_src, keep_src = pipeline.source_clouds(depth01, K)
            for s in range(batch):
                pts = pc_src[s][keep_src[s]]
                scene_pcs.append(pts)
                cloud.write_ply(str(self.samples_folder / 'scene-{:0>6d}/sample-{:0>6d}.cloud.ply'.format(
                    first + s, 0)), cloud.voxel_down_sample(pts, save_voxel_size))
            fragments = [None] * batch
            frag_pose = [None] * batch
            for sample_idx in range(num_samples):
                pose_np = geometry.random_sample_pose(batch).astype(np.float32)    # SDD:2526
                pose = torch.tensor(pose_np).to(dev)
                offsets = torch.tensor(np.cumsum([0] + [p.shape[0] for p in scene_pcs]))
                allpc = torch.cat(scene_pcs).to(torch.float32)
                images_rpj, mask_rpj = geometry.pc2depth_ragged(
                    allpc, offsets, K, image_size=[self.image_size, self.image_size], pose=pose)
                images_rpj = images_rpj * 0.1                                       # SDD:2552
                if write_images:
                    for s in range(batch):
                        utils.save_image(images_rpj[s], str(self.samples_folder /
                                         'scene-{:0>6d}/reprojected.image.png'.format(first + s)))
                if self.depth_correction is not None:                               # SDD:2564-2567
                    m = self.depth_correction.keep_mask(images_rpj, pipeline.KEEP_THRESHOLD)
                    images_rpj = torch.where(m, images_rpj, torch.zeros_like(images_rpj))
                    mask_rpj = mask_rpj & m
                img_cond = torch.cat([images_rpj, mask_rpj.to(images_rpj.dtype)], dim=1) * 2 - 1
                images = model.sample(param_cond=geometry.param_vector(K), img_cond=img_cond,
                                      disable_tqdm=True, has_refine_step=has_refine_step)
                if self.depth_correction is not None:                               # SDD:2579-2581
                    m = self.depth_correction.keep_mask(images, pipeline.KEEP_THRESHOLD)
                    images = torch.where(m, images, torch.zeros_like(images))
                pc_new, counts = geometry.point_cloud_batch(images, K, pose=pose, scale=10.0,
                                                            clip=pipeline.DEPTH_CLIP)
                counts = counts.cpu().tolist()
                for s in range(batch):
                    sdir = self.samples_folder / 'scene-{:0>6d}'.format(first + s)
                    np.savetxt(str(sdir / 'sample-{:0>6d}.pose.txt'.format(sample_idx + 1)),
                               np.linalg.inv(pose_np[s]))                            # SDD:2593-2594
                    if write_images:
                        import cv2
                        utils.save_image(images_rpj[s], str(sdir / 'corrected.image.png'))
                        utils.save_image(images[s], str(sdir / 'sample-{:0>6d}.image.png'.format(sample_idx + 1)))
                        dep = (images[s, 0].cpu().numpy() * 1e4).astype(np.uint16)   # SDD:2618-2620
                        cv2.imwrite(str(sdir / 'sample-{:0>6d}.depth.png'.format(sample_idx + 1)), dep)
                    pc = pc_new[s, :counts[s]]
                    if sample_idx == 0:
                        fragments[s], frag_pose[s] = pc, pose_np[s]
                    else:
                        fragments[s] = torch.cat([fragments[s], pc], dim=0)
                    if sample_idx == num_samples - 1:                                # SDD:2640-2658
                        p = cloud.transform(fragments[s], frag_pose[s])
                        p = cloud.crop(p, pipeline.BBOX_MIN, pipeline.BBOX_MAX)
                        p = cloud.voxel_down_sample(p, save_voxel_size)
                        p = cloud.transform(p, np.linalg.inv(frag_pose[s]))
                        cloud.write_ply(str(sdir / 'sample-{:0>6d}.cloud.ply'.format(1)), p)
                    # scene memory (SDD:2661-2680)
                    merged = torch.cat([scene_pcs[s].to(torch.float64), pc], dim=0)
                    scene_pcs[s] = cloud.voxel_down_sample(merged, memory_voxel_size).to(torch.float32)
            done += batch
        totals = pdist.sum_counters([done], dev)
        return int(totals[0])


class _EmaHolder:
    """Stand-in for ema_pytorch.EMA at inference time: owns `ema_model` and accepts the EMA
    entry of a reference checkpoint (keys 'ema_model.*', 'online_model.*', 'initted', 'step')."""

    def __init__(self, model):
        self.online_model = model
        self.ema_model = model            # until a checkpoint provides separate EMA weights
        self._own_copy = False

    def load_state_dict(self, sd):
        ema = {k[len("ema_model."):]: v for k, v in sd.items() if k.startswith("ema_model.")}
        if not ema:
            raise KeyError("checkpoint 'ema' entry has no ema_model.* keys")
        if not self._own_copy:
            self.ema_model = copy.deepcopy(self.online_model)
            self._own_copy = True
        self.ema_model.load_state_dict(ema)

    def to(self, device):
        self.ema_model.to(device)
        return self

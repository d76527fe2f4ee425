"""`Generator` under the reference's name and signature (SDD:2250-2694): the driver of
`generate_dataset.py`.  Per batch of scenes: source frame -> cloud -> random pose -> z-buffer
reprojection -> depth correction -> DDNM sampling -> depth correction -> cloud -> fuse /
voxel-downsample -> files, with the reference's file layout and skip-if-done resume.

Differences, all forced by the offline environment or by the B200 design:
  * accelerate / ema_pytorch / open3d are replaced by a few lines each (device pick, EMA
    weight holder, pointreggpt_b200.cloud);
  * the scene range is sharded across ranks (pointreggpt_b200.dist) instead of every rank
    generating the same range;
  * source frames come from the 3DMatch tree `folder`, or from the seeded synthetic generator when
    `folder == "synthetic"` (no dataset can be downloaded here); a missing tree is an error;
  * poses and sampler noise are keyed by (base_seed, scene, sample), see pointreggpt_b200.rng;
  * `batch_size` keeps the reference's meaning for the skip-if-done bookkeeping, but the scenes of
    several such batches run through the device together (`device_batch`, default 32): a scene's files
    do not depend on the batch it travels in, and a B200 is at 60 % of its batch-32 throughput at batch 4;
  * files are written by a worker thread while the next batch runs on the GPU.
"""
import copy
import os
import pickle
import shutil
from pathlib import Path

import numpy as np
import torch

from . import cloud, dist as pdist, geometry, pipeline, rng, synthetic


class Generator(object):
    def __init__(self, diffusion_model, folder, *, batch_size=16, ema_update_every=10,
                 ema_decay=0.995, results_folder='./results', samples_folder='./samples',
                 amp=False, fp16=False, split_batches=True, device=None, device_batch=None):
        super().__init__()
        if amp or fp16:
            raise NotImplementedError("the reference runs generation with amp=False (GD:54); the "
                                      "native path has its own fixed fp16-operand numerics")
        self.rank, self.world_size = pdist.world()
        if device is None:
            local = int(os.environ.get("LOCAL_RANK", "0"))
            device = torch.device("cuda", local)
        self._device = torch.device(device)
        if self._device.type == "cuda":
            torch.cuda.set_device(self._device)
        self.folder = folder
        self.model = diffusion_model.to(self._device)
        self.channels = diffusion_model.channels
        self.batch_size = batch_size
        # scenes per pass through the device (>= batch_size); PRG_DEVICE_BATCH overrides the default
        if device_batch is None:
            device_batch = int(os.environ.get("PRG_DEVICE_BATCH", "32"))
        self.device_batch = max(int(device_batch), int(batch_size), 1)
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
                 write_images=True, base_seed=0):
        """SDD:2327-2694.  `base_seed`: every (scene, sample) draws its pose and its sampler noise from
        `rng.scene_seed(base_seed, absolute scene index, sample index)`, so a scene's files do not
        depend on batch size, rank, world size or the -start/-stop split (the reference draws from
        unseeded global generators).  The device work of a batch is queued without host
        synchronisation; a writer thread waits for it, copies the results to the host once and writes
        the files while the main thread queues the next batch."""
        dev = self.device
        torch.cuda.set_device(dev)
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
        if self.folder != "synthetic":
            if not os.path.isdir(str(self.folder)):
                raise FileNotFoundError(
                    "3DMatch RGB-D tree %r does not exist (pass folder='synthetic' for the seeded "
                    "synthetic source frames)" % (self.folder,))
            with open("./dataset/indoor/metadata/train_info.pkl", 'rb') as f:    # SDD:2352-2354
                info_train = pickle.load(f)
        lo, hi = pdist.shard_range(start_scene_index, stop_scene_index, self.rank, self.world_size)
        S = self.image_size
        done = 0
        writer = _Writer(dev)
        try:
            # the reference's batches decide what is skipped (SDD:2371-2381: a batch is complete when the last
            # sample of its last scene exists); the scenes of the remaining ones are coalesced
            todo = []
            for b_idx, batch in enumerate(geometry.num_to_groups(hi - lo, self.batch_size)):
                first = lo + b_idx * self.batch_size
                last_ply = self.samples_folder / 'scene-{:0>6d}/sample-{:0>6d}.cloud.ply'.format(
                    first + batch - 1, num_samples // 2)
                if os.path.isfile(last_ply):
                    print("Skip completed scene {:0>6d} - {:0>6d}.".format(first, first + batch - 1))
                    done += batch
                    continue
                todo.extend(range(first, first + batch))
            for g0 in range(0, len(todo), self.device_batch):
                scenes = todo[g0:g0 + self.device_batch]
                batch = len(scenes)
                sdirs = []
                depths, Ks = [], []
                for sc in scenes:
                    sdir = self.samples_folder / 'scene-{:0>6d}'.format(sc)
                    if sdir.exists():
                        shutil.rmtree(str(sdir), ignore_errors=True)
                    sdir.mkdir(parents=True, exist_ok=True)
                    sdirs.append(sdir)
                    d, K = self._source_frame(sc, info_train)
                    depths.append(d)
                    Ks.append(K)
                K_np = np.stack(Ks)
                depth01 = torch.stack(depths).pin_memory().to(dev, non_blocking=True)
                K = torch.tensor(K_np).pin_memory().to(dev, non_blocking=True)
                # source clouds (SDD:2479-2500): f32 cloud cropped to the room box; dense slabs + keep mask
                pc_src, keep_src = pipeline.source_clouds(depth01, K)
                writer.submit(self._write_sources, sdirs, K_np, depth01, pc_src, keep_src,
                              save_voxel_size, write_images)
                memory = None                  # num_samples > 1: per-scene accumulated clouds (list)
                fragments = [None] * batch
                frag_pose = [None] * batch
                for sample_idx in range(num_samples):
                    seeds = [rng.scene_seed(base_seed, sc, sample_idx) for sc in scenes]
                    pose_np = np.concatenate([
                        geometry.random_sample_pose(1, rng=rng.scene_rng(base_seed, sc, sample_idx))
                        for sc in scenes]).astype(np.float32)                         # SDD:2526
                    pose = torch.tensor(pose_np).pin_memory().to(dev, non_blocking=True)
                    if memory is None:
                        offsets = torch.arange(batch + 1, device=dev, dtype=torch.int64) * (S * S)
                        images_rpj, mask_rpj = geometry.pc2depth_ragged(
                            pc_src.reshape(-1, 3), offsets, K, image_size=[S, S], valid=keep_src, pose=pose)
                    else:
                        offsets = torch.tensor(np.cumsum([0] + [p.shape[0] for p in memory]))
                        images_rpj, mask_rpj = geometry.pc2depth_ragged(
                            torch.cat(memory), offsets, K, image_size=[S, S], pose=pose)
                    images_rpj = images_rpj * 0.1                                       # SDD:2552
                    rpj_raw = images_rpj
                    if self.depth_correction is not None:                               # SDD:2564-2567
                        m = self.depth_correction.keep_mask(images_rpj, pipeline.KEEP_THRESHOLD)
                        images_rpj = torch.where(m, images_rpj, torch.zeros_like(images_rpj))
                        mask_rpj = mask_rpj & m
                    img_cond = torch.cat([images_rpj, mask_rpj.to(images_rpj.dtype)], dim=1) * 2 - 1
                    images = model.sample(param_cond=geometry.param_vector(K), img_cond=img_cond,
                                          disable_tqdm=True, has_refine_step=has_refine_step, seed=seeds)
                    if self.depth_correction is not None:                               # SDD:2579-2581
                        m = self.depth_correction.keep_mask(images, pipeline.KEEP_THRESHOLD)
                        images = torch.where(m, images, torch.zeros_like(images))
                    pc_new, counts = geometry.point_cloud_batch(images, K, pose=pose, scale=10.0,
                                                                clip=pipeline.DEPTH_CLIP)
                    last = sample_idx == num_samples - 1
                    if num_samples == 1:
                        # the common case (GD:59-63): nothing on this thread needs the results
                        writer.submit(self._write_sample, sdirs, sample_idx, pose_np, rpj_raw if write_images else None,
                                      images_rpj if write_images else None, images, pc_new, counts,
                                      [None] * batch, list(pose_np), True, save_voxel_size, write_images)
                        continue
                    counts_h = counts.cpu().tolist()
                    prev = list(fragments)
                    for s in range(batch):
                        pc = pc_new[s, :counts_h[s]]
                        if sample_idx == 0:
                            fragments[s], frag_pose[s] = pc, pose_np[s]
                        else:
                            fragments[s] = torch.cat([fragments[s], pc], dim=0)
                    writer.submit(self._write_sample, sdirs, sample_idx, pose_np, rpj_raw if write_images else None,
                                  images_rpj if write_images else None, images, pc_new, counts, prev,
                                  list(frag_pose), last, save_voxel_size, write_images)
                    if not last:
                        # scene memory (SDD:2661-2680) feeds the next sample's reprojection
                        if memory is None:
                            memory = [pc_src[s][keep_src[s]] for s in range(batch)]
                        for s in range(batch):
                            merged = torch.cat([memory[s].to(torch.float64), pc_new[s, :counts_h[s]]], dim=0)
                            memory[s] = cloud.voxel_down_sample(merged, memory_voxel_size).to(torch.float32)
                done += batch
        finally:
            writer.close()
        totals = pdist.sum_counters([done], dev)
        return int(totals[0])

    # ------------------------------------------------------------------ writer-thread jobs
    def _write_sources(self, sdirs, K_np, depth01, pc_src, keep_src, save_voxel_size, write_images):
        from torchvision import utils
        for s, sdir in enumerate(sdirs):
            np.savetxt(str(sdir / 'camera-intrinsics.txt'), K_np[s])
            if write_images:
                utils.save_image(depth01[s], str(sdir / 'sample-{:0>6d}.image.png'.format(0)))
            pts = pc_src[s][keep_src[s]]
            cloud.write_ply(str(sdir / 'sample-{:0>6d}.cloud.ply'.format(0)),
                            cloud.voxel_down_sample(pts, save_voxel_size))

    def _write_sample(self, sdirs, sample_idx, pose_np, rpj_raw, rpj_crt, images, pc_new, counts,
                      prev_fragments, frag_pose, last, save_voxel_size, write_images):
        from torchvision import utils
        counts_h = counts.cpu().tolist()
        img_h = images.cpu()
        for s, sdir in enumerate(sdirs):
            np.savetxt(str(sdir / 'sample-{:0>6d}.pose.txt'.format(sample_idx + 1)),
                       np.linalg.inv(pose_np[s]))                                    # SDD:2593-2594
            if write_images:
                import cv2
                utils.save_image(rpj_raw[s], str(sdir / 'reprojected.image.png'))         # SDD:2555-2561
                utils.save_image(rpj_crt[s], str(sdir / 'corrected.image.png'))
                utils.save_image(img_h[s], str(sdir / 'sample-{:0>6d}.image.png'.format(sample_idx + 1)))
                dep = (img_h[s, 0].numpy() * 1e4).astype(np.uint16)                  # SDD:2618-2620
                cv2.imwrite(str(sdir / 'sample-{:0>6d}.depth.png'.format(sample_idx + 1)), dep)
            if last:                                                                 # SDD:2640-2658
                pc = pc_new[s, :counts_h[s]]
                frag = pc if prev_fragments[s] is None else torch.cat([prev_fragments[s], pc], dim=0)
                p = cloud.transform(frag, frag_pose[s])
                p = cloud.crop(p, pipeline.BBOX_MIN, pipeline.BBOX_MAX)
                p = cloud.voxel_down_sample(p, save_voxel_size)
                p = cloud.transform(p, np.linalg.inv(frag_pose[s]))
                cloud.write_ply(str(sdir / 'sample-{:0>6d}.cloud.ply'.format(1)), p)


class _Writer:
    """One worker thread with its own CUDA stream.  `submit(fn, *args)` records an event on the
    caller's stream; the worker makes its stream wait for that event, runs `fn` (small device ops,
    device-to-host copies, file writes) and the main thread never blocks on it.  Exceptions surface
    at the next submit / close."""

    def __init__(self, device, depth=2):
        import queue
        import threading
        self.device = device
        self.q = queue.Queue(maxsize=depth)     # bounds the results kept alive on the device
        self.err = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        torch.cuda.set_device(self.device)
        stream = torch.cuda.Stream(self.device)
        while True:
            job = self.q.get()
            if job is None:
                return
            ev, fn, args = job
            if self.err is not None:
                continue
            try:
                with torch.cuda.stream(stream), torch.no_grad():
                    stream.wait_event(ev)
                    fn(*args)
                stream.synchronize()
            except BaseException as e:   # noqa: BLE001 - re-raised on the main thread
                self.err = e

    def _check(self):
        if self.err is not None:
            err, self.err = self.err, None
            raise err

    def submit(self, fn, *args):
        self._check()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.q.put((ev, fn, args))

    def close(self):
        self.q.put(None)
        self.t.join()
        self._check()


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

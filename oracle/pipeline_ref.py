"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-batch body of `Generator.generate`
(SDD:2479-2628): source cloud -> bbox crop -> random-pose z-buffer reprojection -> depth correction
-> DDNM-conditioned sampling -> depth correction -> cloud back in the source frame.

Built from the pinned pieces in oracle/geometry_ref.py and oracle/torch_ref.py; the composition itself
is pinned against the same sequence of *reference* calls in tests/test_oracle_vs_reference.py.  It is
the checker for pointreggpt_b200.pipeline.generate_batch; the product never imports it.
"""
import numpy as np
import torch

from . import geometry_ref as G
from . import torch_ref as R

BBOX_MIN = np.array([-1.5, -1.5, 0.5], np.float32)      # SDD:2348-2349
BBOX_MAX = np.array([1.5, 1.5, 3.5], np.float32)


@torch.no_grad()
def generate_batch(unet_sd, mask_sd, d01, K, P, noises, *, timesteps, sampling_timesteps=None,
                   eta=1.0, has_refine_step=False, keep_threshold=0.99):
    """d01 (B,1,S,S) torch f32, K (B,3,3), P (B,4,4) numpy f32, noises: list of (B,1,S,S) draws.
    Returns dict(images, images_rpj, mask_rpj, img_cond, clouds=[(N_b,3) f64])."""
    B, _, S, _ = d01.shape
    K = np.asarray(K, np.float32)
    P = np.asarray(P, np.float32)
    rpj, msk = [], []
    for b in range(B):
        # SDD:2479-2490: float64 point_cloud -> float32 -> crop to the box (boundary kept)
        pc = G.depth2pc_compact(d01[b:b + 1].numpy(), K[b:b + 1], None)[0].astype(np.float32)
        pc = pc[np.all((pc >= BBOX_MIN) & (pc <= BBOX_MAX), axis=1)]
        # SDD:2531-2545: pose applied (float32), then the z-buffer with every point valid
        offs = np.array([0, pc.shape[0]], np.int64)
        d, m = G.pc2depth(pc, None, offs, K[b:b + 1], (S, S), pose=P[b:b + 1])
        rpj.append(d)
        msk.append(m)
    images_rpj = torch.tensor(np.concatenate(rpj)) * 0.1                         # SDD:2552
    mask_rpj = torch.tensor(np.concatenate(msk))
    mask_crt = R.maskunet_forward(mask_sd, images_rpj) > keep_threshold          # SDD:2564-2565
    images_rpj = torch.where(mask_crt, images_rpj, torch.zeros_like(images_rpj))  # SDD:2566
    mask_rpj = mask_rpj & mask_crt                                               # SDD:2567
    img_cond = torch.cat([images_rpj, mask_rpj.to(images_rpj.dtype)], dim=1) * 2 - 1
    param_cond = torch.tensor(np.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], axis=1))
    sch = R.make_schedule(timesteps)
    if sampling_timesteps is None or sampling_timesteps >= timesteps:
        images = R.p_sample_loop(unet_sd, sch, param_cond, img_cond, noises, has_refine_step)
    else:
        images = R.ddim_sample(unet_sd, sch, param_cond, img_cond, noises, sampling_timesteps, eta,
                               has_refine_step)
    mask_crt2 = R.maskunet_forward(mask_sd, images) > keep_threshold             # SDD:2579-2580
    images = torch.where(mask_crt2, images, torch.zeros_like(images))            # SDD:2581
    clouds = G.depth2pc_compact(images.numpy(), K, P)                            # SDD:2623-2628
    return dict(images=images, images_rpj=images_rpj, mask_rpj=mask_rpj, img_cond=img_cond,
                clouds=clouds)

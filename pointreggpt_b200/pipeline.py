"""The per-batch data-generation path of `Generator.generate` (SDD:2479-2628) as one
device-resident function: source cloud -> random-pose z-buffer reprojection -> depth
correction -> DDNM-conditioned sampling -> depth correction -> point cloud in the source frame.

Every numerical stage runs in libprg.so; torch only glues tiny elementwise steps (mask and,
concat) and owns the device buffers.
"""
import torch

from . import geometry

BBOX_MIN = (-1.5, -1.5, 0.5)   # SDD:2348-2349
BBOX_MAX = (1.5, 1.5, 3.5)
DEPTH_CLIP = (0.5, 10)         # SDD:2482, 2625
KEEP_THRESHOLD = 0.99          # SDD:2565, 2580


def source_clouds(depth01, K):
    """`point_cloud(depth*10, K, clip=[0.5,10]).astype(float32)` + bbox crop (SDD:2479-2490).
    Returns dense slabs pc (B, HW, 3) f32 and the per-point keep mask (B, HW) bool."""
    pc64, counts = geometry.point_cloud_batch(depth01, K, scale=10.0, clip=DEPTH_CLIP)
    pc = pc64.to(torch.float32)
    B, HW, _ = pc.shape
    idx = torch.arange(HW, device=pc.device)[None, :]
    valid = idx < counts[:, None]
    lo = torch.tensor(BBOX_MIN, device=pc.device, dtype=torch.float32)
    hi = torch.tensor(BBOX_MAX, device=pc.device, dtype=torch.float32)
    inside = ((pc >= lo) & (pc <= hi)).all(dim=-1)      # open3d crop keeps the boundary
    return pc, valid & inside


def generate_batch(diffusion, depth_correction, depth01, K, pose, *, has_refine_step=False,
                   noise=None, seed=None, return_intermediates=False):
    """One batch of pairs.

    depth01 (B,1,S,S) f32 source frames in units of 10 m, K (B,3,3), pose (B,4,4): CUDA tensors.
    Returns (pc (B, S*S, 3) f64 slabs in the source frame, counts (B) i64, images (B,1,S,S)).
    """
    B, _, S, _ = depth01.shape
    pc, keep = source_clouds(depth01, K)
    offsets = torch.arange(B + 1, device=pc.device, dtype=torch.int64) * (S * S)
    # z-buffer reprojection into the new view (SDD:2531-2552); the pose is applied in-kernel
    images_rpj, mask_rpj = geometry.pc2depth_ragged(pc.reshape(-1, 3), offsets, K,
                                                    image_size=[S, S], valid=keep, pose=pose)
    images_rpj = images_rpj * 0.1
    if depth_correction is not None:                     # SDD:2564-2567
        mask_crt = depth_correction.keep_mask(images_rpj, KEEP_THRESHOLD)
        images_rpj = torch.where(mask_crt, images_rpj, torch.zeros_like(images_rpj))
        mask_rpj = mask_rpj & mask_crt
    img_cond = torch.cat([images_rpj, mask_rpj.to(images_rpj.dtype)], dim=1) * 2 - 1   # SDD:2569-2570
    param_cond = geometry.param_vector(K)
    images = diffusion.sample(param_cond=param_cond, img_cond=img_cond, disable_tqdm=True,
                              has_refine_step=has_refine_step, noise=noise, seed=seed)
    if depth_correction is not None:                     # SDD:2579-2581
        mask_crt = depth_correction.keep_mask(images, KEEP_THRESHOLD)
        images = torch.where(mask_crt, images, torch.zeros_like(images))
    # depth -> cloud, expressed back in the source frame (SDD:2623-2628)
    pc_out, counts = geometry.point_cloud_batch(images, K, pose=pose, scale=10.0, clip=DEPTH_CLIP)
    if return_intermediates:
        return pc_out, counts, images, dict(images_rpj=images_rpj, mask_rpj=mask_rpj,
                                            img_cond=img_cond)
    return pc_out, counts, images

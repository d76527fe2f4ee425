"""Geometry layer of the hot path, under the reference's own function names.

Mirrors the free functions of the reference's
``denoising_diffusion_pytorch/successive_ddnm_diffusion.py`` (SDD): same names, argument
meaning and return types; the tensor functions run hand-written CUDA kernels through
libprg.so (bit-exact z-buffer / indices) and accept CUDA tensors only.
"""
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _ffi


# ------------------------------------------------------------------ host-side helpers
def intrinsic_transform(intrinsic: np.ndarray,
                        resize: Optional[Union[int, Sequence[int]]] = None,
                        centercrop: Optional[Union[int, Sequence[int]]] = None) -> np.ndarray:
    """Rescale a pinhole K for torchvision Resize(resize) + CenterCrop(centercrop).

    Follows SDD:47-119 including its conventions: the original image size is taken as
    (2*cx, 2*cy) truncated to int32, an int ``resize`` scales the shorter side, the crop
    origin is ``round((new - crop) / 2)`` with numpy's half-to-even rounding.
    """
    K = np.asarray(intrinsic)
    fx, fy = K[..., 0, 0], K[..., 1, 1]
    cx, cy = K[..., 0, 2], K[..., 1, 2]
    size_x, size_y = np.int32(cx * 2), np.int32(cy * 2)
    # SDD:64-67 (no-resize defaults; note the reference seeds new_cy with old_cx)
    nfx, nfy, ncx, ncy = fx, fy, cx, cx
    nsx, nsy = size_x, size_y
    if resize is not None:
        if type(resize) == int:
            if (size_x < size_y).all():
                nsx = int(resize)
                nsy = np.int32(np.floor(resize * size_y / size_x))
            else:
                nsx = np.int32(np.floor(resize * size_x / size_y))
                nsy = np.int32(resize)
        elif type(resize) == tuple:
            nsx, nsy = np.int32(resize[1]), np.int32(resize[0])
        nfx = np.float32(fx * nsx / size_x)
        nfy = np.float32(fy * nsy / size_y)
        ncx = np.float32(nsx / 2)
        ncy = np.float32(nsy / 2)
    if centercrop is not None:
        if type(centercrop) == int:
            cw = ch = centercrop
        elif type(centercrop) == tuple:
            cw, ch = centercrop[1], centercrop[0]
        ncx = ncx - np.int32(np.round((nsx - cw) / 2.0))
        ncy = ncy - np.int32(np.round((nsy - ch) / 2.0))
    out = np.zeros_like(K)
    out[..., 0, 0] = nfx
    out[..., 1, 1] = nfy
    out[..., 0, 2] = ncx
    out[..., 1, 2] = ncy
    out[..., 2, 2] = 1.0
    return out


def param_vector(intrinsic):
    """[fx, fy, cx, cy] per camera (SDD:343-351)."""
    return torch.stack([intrinsic[..., 0, 0], intrinsic[..., 1, 1],
                        intrinsic[..., 0, 2], intrinsic[..., 1, 2]], dim=-1)


_3DMATCH_FOCALS = (585.0, 572.0, 583.0, 540.021232, 570.342205, 533.069214)
_3DMATCH_WEIGHTS = (7, 8, 18, 5, 47, 5)


def random_sample_intrinsic(batch_size) -> np.ndarray:
    """Draw 3DMatch intrinsics with their dataset frequencies (SDD:354-374).
    Consumes the numpy global RNG exactly like the reference (one `choice` call)."""
    cand = np.zeros((len(_3DMATCH_FOCALS), 3, 3), dtype=np.float32)
    for i, f in enumerate(_3DMATCH_FOCALS):
        cand[i] = [[f, 0.0, 320.0], [0.0, f, 240.0], [0.0, 0.0, 1.0]]
    prob = np.array(_3DMATCH_WEIGHTS)
    prob = prob / np.sum(prob)
    idx = np.random.choice(len(cand), batch_size, replace=True, p=prob)
    return cand[idx]


def random_sample_pose(batch_size, center=(0, 0, 3), rng=None):
    """Random camera motion about a pivot 3 m ahead (SDD:417-443): pitch in +-pi/24, yaw in
    +-pi/12, no roll, x/y translation jitter N(0, 1/9).  Same numpy RNG call order as the
    reference (rand, rand, randn) on numpy's global RNG; with `rng` (a numpy Generator, see
    pointreggpt_b200.rng.scene_rng) the same three draws come from that generator instead."""
    from scipy.spatial.transform import Rotation
    tmin, tmax = -np.pi / 24, np.pi / 24
    pmin, pmax = -np.pi / 12, np.pi / 12
    rand = np.random.rand if rng is None else (lambda n: rng.random(n))
    randn = np.random.randn if rng is None else (lambda n, m: rng.standard_normal((n, m)))
    theta = rand(batch_size) * (tmax - tmin) + tmin
    phi = rand(batch_size) * (pmax - pmin) + pmin
    euler = np.stack((theta, phi, np.zeros(batch_size)), axis=-1)
    rot = Rotation.from_euler("XYZ", euler, degrees=False).as_matrix()
    c = np.array(center)
    jitter = randn(batch_size, 3) / 3
    jitter[:, -1] = 0
    trans = c - rot @ c + jitter
    T = np.stack([np.eye(4) for _ in range(batch_size)])
    T[..., :3, :3] = rot
    T[..., :3, 3] = trans
    return T.astype(np.float32)


def random_sample_transform(intrinsic, image_size=256):
    """Pure random rotation that keeps the optical axis inside the old view (SDD:377-414): pitch / yaw
    bounded by the field of view, free roll, zero translation.  Same numpy RNG call order as the
    reference (rand, rand, rand, randn)."""
    from scipy.spatial.transform import Rotation
    batch_size = intrinsic.shape[0]
    h, w = image_size, image_size
    fx, fy = intrinsic[..., 0, 0], intrinsic[..., 1, 1]
    cx, cy = intrinsic[..., 0, 2], intrinsic[..., 1, 2]
    theta_min, theta_max = -np.arctan((h - cy) / fy), np.arctan(cy / fy)      # about the x axis
    phi_min, phi_max = -np.arctan(cx / fx), np.arctan((w - cx) / fx)          # about the y axis
    theta = np.random.rand(batch_size) * (theta_max - theta_min) + theta_min
    phi = np.random.rand(batch_size) * (phi_max - phi_min) + phi_min
    psi = np.random.rand(batch_size) * 2 * np.pi - np.pi
    rot = Rotation.from_euler("XYZ", np.stack((theta, phi, psi), axis=-1), degrees=False).as_matrix()
    trans = np.random.randn(batch_size, 3) / 3 * 0
    T = np.stack([np.eye(4) for _ in range(batch_size)])
    T[..., :3, :3] = rot
    T[..., :3, 3] = trans
    return T.astype(np.float32)


def num_to_groups(num, divisor):
    """SDD:538-544."""
    arr = [divisor] * (num // divisor)
    if num % divisor > 0:
        arr.append(num % divisor)
    return arr


def normalize_to_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_to_zero_to_one(t):
    return (t + 1) * 0.5


def get_mask_from_img_cond(img_cond):
    """SDD:507-508."""
    return unnormalize_to_zero_to_one(img_cond[:, 1, None, ...]) > 0.5


def null_image_condition(batch_size, image_size, dtype=None, device=None):
    """SDD:499-504."""
    return -torch.ones((batch_size, 2, image_size, image_size)).to(dtype=dtype, device=device)


# ------------------------------------------------------------------ CUDA-backed tensor ops
def _f32c(t, device=None):
    t = t.to(dtype=torch.float32)
    if device is not None:
        t = t.to(device)
    return t.contiguous()


def depth2pc_tensor(depth, intrinsic, *, clip=[0, 10], invalid_num=None):
    """depth (b,1,h,w) -> pc (b, h*w, 3), valid (b, h*w) bool (SDD:176-209)."""
    _ffi.require_cuda(depth)
    b, c, h, w = depth.shape
    assert c == 1, "depth2pc_tensor expects a single-channel depth map"
    d = _f32c(depth)
    K = _f32c(intrinsic, d.device)
    inv = float("nan") if invalid_num is None else float(invalid_num)
    pc = torch.empty((b, h * w, 3), dtype=torch.float32, device=d.device)
    valid = torch.empty((b, h * w), dtype=torch.uint8, device=d.device)
    use_clip = clip is not None
    lo, hi = (float(clip[0]), float(clip[1])) if use_clip else (0.0, 0.0)
    _ffi.check(_ffi.lib().prg_depth2pc_f32(_ffi.ptr(d), _ffi.ptr(K), lo, hi, int(use_clip), inv,
                                           _ffi.ptr(pc), _ffi.ptr(valid), b, h, w, _ffi.stream(d)))
    return pc, valid.view(torch.bool)


def pc2depth_ragged(pc, offsets, intrinsic, *, image_size, valid=None, pose=None):
    """Z-buffer a ragged batch of point clouds: pc (sumN,3), offsets (B+1) int64 CSR.
    Optional per-image rigid `pose` (B,4,4) applied first (p' = R p + t)."""
    _ffi.require_cuda(pc)
    rows, cols = image_size
    pcs = _f32c(pc).reshape(-1, 3)
    dev = pcs.device
    off = offsets.to(device=dev, dtype=torch.int64).contiguous()
    B = off.numel() - 1
    K = _f32c(intrinsic, dev)
    v = None if valid is None else valid.to(device=dev).reshape(-1).to(torch.uint8).contiguous()
    P = None if pose is None else _f32c(pose, dev)
    depth = torch.empty((B, 1, rows, cols), dtype=torch.float32, device=dev)
    mask = torch.empty((B, 1, rows, cols), dtype=torch.uint8, device=dev)
    _ffi.check(_ffi.lib().prg_pc2depth_f32(_ffi.ptr(pcs), _ffi.ptr(v), _ffi.ptr(off),
                                           pcs.shape[0], _ffi.ptr(K), _ffi.ptr(P), _ffi.ptr(depth),
                                           _ffi.ptr(mask), B, rows, cols, _ffi.stream(pcs)))
    return depth, mask.view(torch.bool)


def pc2depth_tensor(pc, valid, intrinsic, *, image_size=[480, 640]):
    """pc (b,n,3), valid (b,n) -> depth (b,1,h,w) (0 where empty), mask bool (SDD:212-265)."""
    _ffi.require_cuda(pc)
    b, n, _ = pc.shape
    offsets = torch.arange(b + 1, dtype=torch.int64, device=pc.device) * n
    return pc2depth_ragged(pc.reshape(-1, 3), offsets, intrinsic, image_size=image_size,
                           valid=valid)


def reproject_tensor(depth, intrinsic, relative_pose, *, clip=[0, 10], invalid_num=None):
    """Reproject depth maps into the camera at `relative_pose` (SDD:268-286); fused
    unproject -> SE(3) -> project -> z-buffer kernel."""
    _ffi.require_cuda(depth)
    b, c, h, w = depth.shape
    assert c == 1
    d = _f32c(depth)
    K = _f32c(intrinsic, d.device)
    P = _f32c(relative_pose, d.device)
    out = torch.empty((b, 1, h, w), dtype=torch.float32, device=d.device)
    mask = torch.empty((b, 1, h, w), dtype=torch.uint8, device=d.device)
    _ffi.check(_ffi.lib().prg_reproject_f32(_ffi.ptr(d), _ffi.ptr(K), _ffi.ptr(P), float(clip[0]),
                                            float(clip[1]), _ffi.ptr(out), _ffi.ptr(mask), b, h, w,
                                            _ffi.stream(d)))
    return out, mask.view(torch.bool)


def occlusion_filter(depth_rpj, mask_rpj):
    """Replace pixels that lie more than 0.0375 behind the nearest valid depth of their 3x3
    neighbourhood by that depth (SDD:446-463).  Returns (depth, mask_rpj); the mask is passed
    through unchanged, as in the reference."""
    _ffi.require_cuda(depth_rpj, mask_rpj)
    b, c, h, w = depth_rpj.shape
    assert c == 1 and mask_rpj.shape == depth_rpj.shape
    d = _f32c(depth_rpj)
    m = mask_rpj.to(torch.bool).contiguous().view(torch.uint8)
    out = torch.empty_like(d)
    _ffi.check(_ffi.lib().prg_occlusion_filter_f32(_ffi.ptr(d), _ffi.ptr(m), _ffi.ptr(out), b, h, w,
                                                   _ffi.stream(d)))
    return out, mask_rpj


def image_condition(depth, intrinsic, relative_pose, depth_unit=10, depth_clip=[0, 10],
                    use_occlusion_filter=False):
    """depth (b,1,h,w) in [0,1] -> img_cond (b,2,h,w) in [-1,1]: reprojected depth / depth_unit and
    the z-buffer mask (SDD:466-504)."""
    depth_rpj, mask_rpj = reproject_tensor(depth * depth_unit, intrinsic, relative_pose,
                                           clip=depth_clip)
    if use_occlusion_filter:
        depth_rpj, mask_rpj = occlusion_filter(depth_rpj, mask_rpj)
    # a TENSOR divisor: CUDA evaluates `tensor / python_scalar` as tensor * (1 / scalar), which is not
    # the correctly rounded quotient the reference's CPU path produces (SDD:495)
    unit = torch.tensor(depth_unit, dtype=depth_rpj.dtype, device=depth_rpj.device)
    img_cond = torch.cat([depth_rpj / unit, mask_rpj.to(depth_rpj.dtype)], dim=1)
    return normalize_to_neg_one_to_one(img_cond)


def point_cloud_batch(depth01, intrinsic, *, pose=None, scale=10.0, clip=(0.5, 10)):
    """Batched `point_cloud(depth01 * scale, K, clip)` (+ optional `(pc - t) @ R`,
    SDD:2623-2628) on the device.  Returns (pc (B, H*W, 3) float64 slabs, counts (B) int64)."""
    _ffi.require_cuda(depth01)
    d = _f32c(depth01)
    if d.dim() == 4:
        d = d[:, 0].contiguous()
    B, H, W = d.shape
    K = _f32c(intrinsic, d.device)
    P = None if pose is None else _f32c(pose, d.device)
    pc = torch.empty((B, H * W, 3), dtype=torch.float64, device=d.device)
    counts = torch.empty((B,), dtype=torch.int64, device=d.device)
    nblk = (H * W + 1023) // 1024
    scratch = torch.empty((B * (nblk + 1),), dtype=torch.int64, device=d.device)
    _ffi.check(_ffi.lib().prg_depth2pc_compact_f64(
        _ffi.ptr(d), _ffi.ptr(K), _ffi.ptr(P), float(scale), float(clip[0]), float(clip[1]),
        _ffi.ptr(pc), _ffi.ptr(counts), _ffi.ptr(scratch), B, H, W, _ffi.stream(d)))
    return pc, counts


def point_cloud(depth, intrinsic, clip=[0, 10], device="cuda"):
    """numpy-facing `point_cloud` (SDD:122-143): depth (H,W) -> (N,3) float64."""
    d = torch.as_tensor(np.ascontiguousarray(depth, dtype=np.float32), device=device)[None]
    K = torch.as_tensor(np.asarray(intrinsic, dtype=np.float32), device=device)[None]
    pc, counts = point_cloud_batch(d, K, scale=1.0, clip=clip)
    n = int(counts[0].item())
    return pc[0, :n].cpu().numpy()

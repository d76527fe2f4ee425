"""Seeded synthetic inputs for the hot path (no 3DMatch data is available offline).

Depth maps follow SURVEY.md section 8(d): a room-like tilted plane with three nearer
axis-aligned boxes and 10 % holes, quantised to uint16 millimetres like a 3DMatch depth
PNG and scaled by 1e-4 exactly as the reference loader does (SDD:1553-1554, 2458-2459):
values are in [0, 1] in units of 10 m, 0 = invalid.
"""
import numpy as np
import torch

from . import geometry


def synthetic_depth(index, height=256, width=256):
    """One depth map (H, W) float32 in units of 10 m, deterministic in `index`."""
    g = torch.Generator().manual_seed(1000 + int(index))
    u = torch.linspace(0, 1, width)[None, :].expand(height, width)
    v = torch.linspace(0, 1, height)[:, None].expand(height, width)
    r = torch.rand(6, generator=g)
    z0 = 1.5 + 1.5 * r[0]
    a = -0.4 + 0.8 * r[1]
    b = -0.4 + 0.8 * r[2]
    z = z0 + a * u + b * v
    for k in range(3):
        q = torch.rand(5, generator=g)
        x0, y0 = int(q[0] * width * 0.7), int(q[1] * height * 0.7)
        w = int(width * (0.1 + 0.2 * q[2]))
        h = int(height * (0.1 + 0.2 * q[3]))
        z[y0:y0 + h, x0:x0 + w] = 0.8 + 0.7 * q[4]
    holes = torch.rand(height, width, generator=g) < 0.10
    z = torch.where(holes, torch.zeros_like(z), z)
    png = torch.round(z * 1000.0).to(torch.int32).clamp_(0, 65535)   # uint16 millimetres
    d = png.to(torch.float32) * 1e-4
    d[d > 1] = 0
    return d


def synthetic_depth_batch(start, count, height=256, width=256):
    return torch.stack([synthetic_depth(start + i, height, width) for i in range(count)])[:, None]


def synthetic_intrinsics(count, image_size=256, seed=0):
    """3DMatch intrinsics rescaled for Resize+CenterCrop(image_size) (float32, (B,3,3))."""
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        K = geometry.random_sample_intrinsic(count)
    finally:
        np.random.set_state(state)
    if image_size is None:
        return K
    return geometry.intrinsic_transform(K, resize=image_size, centercrop=image_size).astype(np.float32)


def synthetic_poses(count, seed=1):
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        return geometry.random_sample_pose(count)
    finally:
        np.random.set_state(state)

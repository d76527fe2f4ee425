"""TEST INFRASTRUCTURE ONLY -- numpy-facing wrapper of oracle/geometry_ref.c.

Checker for the CUDA geometry kernels; never imported by the product package.
The shared object is built by ``oracle/build.py`` (also called from
``__graft_entry__.build()``); if it is missing and gcc is available it is
built on first use.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libprg_oracle.so")
_SRC = os.path.join(_HERE, "geometry_ref.c")
_lib = None


def build(force=False):
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def depth2pc(depth, K, clip=(0, 10), invalid=float("nan")):
    """depth (B,1,H,W) or (B,H,W) -> pc (B,H*W,3) f32, valid (B,H*W) bool."""
    depth = _f32(depth)
    if depth.ndim == 4:
        depth = depth[:, 0]
    depth = np.ascontiguousarray(depth)
    B, H, W = depth.shape
    K = _f32(K)
    pc = np.empty((B, H * W, 3), np.float32)
    valid = np.empty((B, H * W), np.uint8)
    use_clip = clip is not None
    lo, hi = (clip if use_clip else (0.0, 0.0))
    _load().prg_ref_depth2pc_f32(
        _p(depth, ctypes.c_float), _p(K, ctypes.c_float), ctypes.c_float(lo), ctypes.c_float(hi),
        ctypes.c_int(int(use_clip)), ctypes.c_float(invalid), _p(pc, ctypes.c_float),
        _p(valid, ctypes.c_uint8), B, H, W)
    return pc, valid.astype(bool)


def pc2depth(pc, valid, offsets, K, image_size, pose=None):
    """Ragged z-buffer: pc (sumN,3), valid (sumN) or None, offsets (B+1)."""
    pc = _f32(pc).reshape(-1, 3)
    offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    B = offsets.shape[0] - 1
    H, W = image_size
    K = _f32(K)
    v = None if valid is None else np.ascontiguousarray(np.asarray(valid).reshape(-1).astype(np.uint8))
    P = None if pose is None else _f32(pose)
    depth = np.empty((B, 1, H, W), np.float32)
    mask = np.empty((B, 1, H, W), np.uint8)
    _load().prg_ref_pc2depth_f32(
        _p(pc, ctypes.c_float), _p(v, ctypes.c_uint8), _p(offsets, ctypes.c_int64),
        _p(K, ctypes.c_float), _p(P, ctypes.c_float), _p(depth, ctypes.c_float),
        _p(mask, ctypes.c_uint8), B, H, W)
    return depth, mask.astype(bool)


def reproject(depth, K, pose, clip=(0, 10)):
    """depth (B,1,H,W) metres -> (depth (B,1,H,W), mask (B,1,H,W) bool)."""
    depth = _f32(depth)
    if depth.ndim == 4:
        depth = np.ascontiguousarray(depth[:, 0])
    B, H, W = depth.shape
    K, P = _f32(K), _f32(pose)
    out = np.empty((B, 1, H, W), np.float32)
    mask = np.empty((B, 1, H, W), np.uint8)
    _load().prg_ref_reproject_f32(
        _p(depth, ctypes.c_float), _p(K, ctypes.c_float), _p(P, ctypes.c_float),
        ctypes.c_float(clip[0]), ctypes.c_float(clip[1]), _p(out, ctypes.c_float),
        _p(mask, ctypes.c_uint8), B, H, W)
    return out, mask.astype(bool)


def depth2pc_compact(depth01, K, pose=None, scale=10.0, clip=(0.5, 10)):
    """point_cloud(depth01*scale, K, clip) [+ (pc - t) @ R]; list of (N_b,3) f64."""
    depth01 = _f32(depth01)
    if depth01.ndim == 4:
        depth01 = np.ascontiguousarray(depth01[:, 0])
    B, H, W = depth01.shape
    K = _f32(K)
    P = None if pose is None else _f32(pose)
    pc = np.empty((B, H * W, 3), np.float64)
    counts = np.zeros((B,), np.int64)
    _load().prg_ref_depth2pc_compact_f64(
        _p(depth01, ctypes.c_float), _p(K, ctypes.c_float), _p(P, ctypes.c_float),
        ctypes.c_float(scale), ctypes.c_float(clip[0]), ctypes.c_float(clip[1]),
        _p(pc, ctypes.c_double), _p(counts, ctypes.c_int64), B, H, W)
    return [pc[b, :counts[b]].copy() for b in range(B)]


def occlusion_filter(depth_rpj, mask_rpj):
    """occlusion_filter (SDD:446-463): depth (B,1,H,W) f32 + mask (B,1,H,W) bool -> (depth, mask)."""
    depth = _f32(depth_rpj)
    shape = depth.shape
    H, W = shape[-2:]
    depth = np.ascontiguousarray(depth.reshape(-1, H, W))
    m = np.ascontiguousarray(np.asarray(mask_rpj).reshape(-1, H, W).astype(np.uint8))
    out = np.empty_like(depth)
    _load().prg_ref_occlusion_filter_f32(_p(depth, ctypes.c_float), _p(m, ctypes.c_uint8),
                                         _p(out, ctypes.c_float), depth.shape[0], H, W)
    return out.reshape(shape), np.asarray(mask_rpj).astype(bool)


def voxel_down_sample(points, voxel_size):
    """Open3D-style voxel-grid centroids of points (N,3) f64, ordered by packed voxel index.
    Returns (centroids (M,3) f64, keys (M,) i64)."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
    n = pts.shape[0]
    out = np.empty((max(n, 1), 3), np.float64)
    keys = np.empty((max(n, 1),), np.int64)
    fn = _load().prg_ref_voxel_downsample_f64
    fn.restype = ctypes.c_int64
    m = fn(_p(pts, ctypes.c_double), ctypes.c_int64(n), ctypes.c_double(voxel_size),
           _p(out, ctypes.c_double), _p(keys, ctypes.c_int64))
    if m < 0:
        raise ValueError("voxel index does not fit 21 bits per axis")
    return out[:m].copy(), keys[:m].copy()


def overlap_count(query, target, radius):
    """Number of query points (Nq,3) f64 with a target point (Nt,3) closer than radius."""
    q = np.ascontiguousarray(np.asarray(query, dtype=np.float64).reshape(-1, 3))
    t = np.ascontiguousarray(np.asarray(target, dtype=np.float64).reshape(-1, 3))
    fn = _load().prg_ref_overlap_count_f64
    fn.restype = ctypes.c_int64
    return int(fn(_p(q, ctypes.c_double), ctypes.c_int64(q.shape[0]), _p(t, ctypes.c_double),
                  ctypes.c_int64(t.shape[0]), ctypes.c_double(radius)))


def compute_overlap_ratio(pc1, pc2, voxel_size=0.025, overlap_factor=1.5, is_down_sample=True):
    """generate_gt.py:68-102 on (N,3) arrays."""
    r = voxel_size * overlap_factor
    if is_down_sample:
        pc1, _ = voxel_down_sample(pc1, voxel_size)
        pc2, _ = voxel_down_sample(pc2, voxel_size)
    return overlap_count(pc1, pc2, r) / pc1.shape[0], overlap_count(pc2, pc1, r) / pc2.shape[0]

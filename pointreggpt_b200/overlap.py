"""Ground-truth overlap ratios of generated cloud pairs -- the step after generation
(generate_gt.py:68-199): voxel-down-sample both clouds at 25 mm, count the points of
each that have a point of the other within 37.5 mm, write `gt.log` per scene and gather them.

The radius search runs on the GPU (csrc/cloud.cu, prg_overlap_count_f64: hash grid of cell size
`radius`, exact counts) instead of open3d's KD-tree loop in Python.

STAGED: written after the round's GPU budget was spent; the parity tests are in
tests/test_zz_staged_gpu.py and have not run on hardware yet.
"""
import os
from itertools import combinations
from pathlib import Path

import numpy as np
import torch

from . import _ffi, cloud


def overlap_count(query, target, radius):
    """Number of points of `query` (Nq,3) with a point of `target` (Nt,3) closer than `radius`."""
    _ffi.require_cuda(query, target)
    q = query.to(torch.float64).contiguous()
    t = target.to(torch.float64).contiguous()
    nq, nt = q.shape[0], t.shape[0]
    if nq == 0 or nt == 0:
        return 0
    ws_bytes = int(_ffi.lib().prg_overlap_workspace_bytes(nt))
    ws = torch.empty((ws_bytes + 15) // 16 * 2, dtype=torch.int64, device=q.device)
    ce = torch.empty((2,), dtype=torch.int32, device=q.device)
    _ffi.check(_ffi.lib().prg_overlap_count_f64(_ffi.ptr(q), nq, _ffi.ptr(t), nt, float(radius),
                                                _ffi.ptr(ce), _ffi.ptr(ws), ws_bytes, _ffi.stream()))
    hits, err = ce.tolist()
    if err:
        raise _ffi.PrgError("overlap_count: non-finite point or coordinates beyond 2^20 search cells")
    return hits


def compute_overlap_ratio(pc1, pc2, voxel_size=0.025, overlap_factor=1.5, is_down_sample=True):
    """generate_gt.py:68-102 on (N,3) CUDA tensors: (ratio of pc1 covered by pc2, ratio of pc2
    covered by pc1).  An empty cloud gives nan, as the reference's 0/0 does."""
    search_voxel_size = voxel_size * overlap_factor
    if is_down_sample:
        pc1 = cloud.voxel_down_sample(pc1, voxel_size)
        pc2 = cloud.voxel_down_sample(pc2, voxel_size)
    n1, n2 = pc1.shape[0], pc2.shape[0]
    r1 = overlap_count(pc1, pc2, search_voxel_size) / n1 if n1 else float("nan")
    r2 = overlap_count(pc2, pc1, search_voxel_size) / n2 if n2 else float("nan")
    return r1, r2


def generate_gt(dataset_name, start_scene_index, stop_scene_index, num_samples, device="cuda"):
    """generate_gt.py:105-176: one `gt.log` per scene with a TSV line per kept pair."""
    root_path = Path("./{}/data".format(dataset_name))
    written = 0
    for scene_idx in range(start_scene_index, stop_scene_index):
        scene_name = "scene-{:0>6d}".format(scene_idx)
        scene_path = root_path.joinpath(scene_name)
        gt_path = scene_path.joinpath("gt.log")
        if gt_path.exists():
            print("scene gt log has existed, skip over it")
            continue
        lines = []
        for src_idx, tgt_idx in combinations(range(num_samples), 2):
            src_path = scene_path.joinpath("sample-{:0>6d}.cloud.ply".format(src_idx))
            tgt_path = scene_path.joinpath("sample-{:0>6d}.cloud.ply".format(tgt_idx))
            if (not src_path.exists()) or (not tgt_path.exists()):
                continue
            src = cloud.read_ply(str(src_path))
            tgt = cloud.read_ply(str(tgt_path))
            if src.shape[0] < 1000 or tgt.shape[0] < 1000:
                continue
            overlap_src, overlap_tgt = compute_overlap_ratio(
                torch.tensor(src, device=device), torch.tensor(tgt, device=device))
            if np.isnan(overlap_src) or np.isnan(overlap_tgt):
                continue
            if overlap_src < 0.1 and overlap_tgt < 0.1:
                continue
            lines.append("{}\t{}\t{}\t{:.4f}\t{:.4f}\n".format(scene_name, src_idx, tgt_idx,
                                                             overlap_src, overlap_tgt))
        gt_path.parent.mkdir(parents=True, exist_ok=True)
        with open(gt_path, "w") as f:
            f.writelines(lines)
        written += 1
    return written


def gather_gt(dataset_name, start_index, stop_index):
    """generate_gt.py:178-190: concatenate the per-scene logs into metadata/gt.log."""
    final_gt_path = Path("./{}/metadata/gt.log".format(dataset_name))
    final_gt_path.parent.mkdir(parents=True, exist_ok=True)
    if final_gt_path.exists():
        print("gt log exists, delete it")
        os.remove(str(final_gt_path))
    with open(final_gt_path, "ab") as out:
        for scene_idx in range(start_index, stop_index):
            scene_gt_path = "./{}/data/scene-{:0>6d}/gt.log".format(dataset_name, scene_idx)
            if os.path.isfile(scene_gt_path):
                with open(scene_gt_path, "rb") as f:
                    out.write(f.read())
    return str(final_gt_path)

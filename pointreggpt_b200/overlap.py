"""Ground-truth overlap ratios of generated cloud pairs -- the step after generation
(generate_gt.py:68-199): voxel-down-sample both clouds at 25 mm, count the points of
each that have a point of the other within 37.5 mm, write `gt.log` per scene and gather them.

The radius search runs on the GPU (csrc/cloud.cu, prg_overlap_count_f64: hash grid of cell size
`radius`, exact counts) instead of open3d's KD-tree loop in Python.

STAGED: written after the round's GPU budget was spent; the parity tests are in
tests/test_zz_staged_gpu.py and have not run on hardware yet.
"""
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
                                                _ffi.ptr(ce), _ffi.ptr(ws), ws_bytes, _ffi.stream(q)))
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


MIN_POINTS = 1000        # generate_gt.py:143-145: pairs with a smaller cloud are not rated
MIN_OVERLAP = 0.1        # generate_gt.py:154: dropped when both ratios are below


def _scene_dir(dataset_name, scene_idx):
    return Path(".") / dataset_name / "data" / "scene-{:0>6d}".format(scene_idx)


def _rate_pair(scene_dir, i, j, device):
    """TSV line for clouds i, j of a scene, or None when the pair is skipped (generate_gt.py:131-163)."""
    files = [scene_dir / "sample-{:0>6d}.cloud.ply".format(k) for k in (i, j)]
    if not all(f.exists() for f in files):
        return None
    pts = [cloud.read_ply(str(f)) for f in files]
    if min(p.shape[0] for p in pts) < MIN_POINTS:
        return None
    ra, rb = compute_overlap_ratio(*(torch.tensor(p, device=device) for p in pts))
    if np.isnan(ra) or np.isnan(rb) or (ra < MIN_OVERLAP and rb < MIN_OVERLAP):
        return None
    return "\t".join([scene_dir.name, str(i), str(j), "%.4f" % ra, "%.4f" % rb]) + "\n"


def generate_gt(dataset_name, start_scene_index, stop_scene_index, num_samples, device="cuda"):
    """One `gt.log` per scene, a TSV line per kept pair (generate_gt.py:105-176); scenes that already
    have one are left alone.  Returns the number of logs written."""
    written = 0
    for scene_idx in range(start_scene_index, stop_scene_index):
        sdir = _scene_dir(dataset_name, scene_idx)
        log = sdir / "gt.log"
        if log.exists():
            print("scene gt log has existed, skip over it")
            continue
        rated = (_rate_pair(sdir, i, j, device) for i, j in combinations(range(num_samples), 2))
        sdir.mkdir(parents=True, exist_ok=True)
        log.write_text("".join(line for line in rated if line is not None))
        written += 1
    return written


def gather_gt(dataset_name, start_index, stop_index):
    """Concatenate the per-scene logs into metadata/gt.log, replacing an older one
    (generate_gt.py:178-190, which shells out to `cat`)."""
    target = Path(".") / dataset_name / "metadata" / "gt.log"
    target.parent.mkdir(parents=True, exist_ok=True)
    parts = []
    for scene_idx in range(start_index, stop_index):
        log = _scene_dir(dataset_name, scene_idx) / "gt.log"
        if log.is_file():
            parts.append(log.read_bytes())
    target.write_bytes(b"".join(parts))
    return str(target)

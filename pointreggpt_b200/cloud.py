"""Point-cloud post-processing the reference delegates to open3d (not available offline):
axis-aligned crop, voxel-grid centroid down-sampling and a binary PLY writer.

Semantics follow Open3D 0.17 as used at SDD:2486-2500, 2640-2680: `crop` keeps points on the
box boundary; `voxel_down_sample` averages the points of each voxel of a grid anchored at
`min_bound - voxel/2`; the order of the output points is unspecified there (hash-map order),
here it is the sorted voxel index.  Runs on whatever device the input tensor lives on.
"""
import numpy as np
import torch


def crop(points, min_bound, max_bound):
    lo = torch.as_tensor(min_bound, dtype=points.dtype, device=points.device)
    hi = torch.as_tensor(max_bound, dtype=points.dtype, device=points.device)
    keep = ((points >= lo) & (points <= hi)).all(dim=-1)
    return points[keep]


def voxel_down_sample(points, voxel_size):
    if points.shape[0] == 0:
        return points
    p = points.to(torch.float64)
    origin = p.min(dim=0).values - voxel_size * 0.5
    idx = torch.floor((p - origin) / voxel_size).to(torch.int64)
    dims = idx.max(dim=0).values + 1
    key = (idx[:, 0] * dims[1] + idx[:, 1]) * dims[2] + idx[:, 2]
    uniq, inv = torch.unique(key, return_inverse=True)
    out = torch.zeros((uniq.shape[0], 3), dtype=torch.float64, device=p.device)
    out.index_add_(0, inv, p)
    cnt = torch.zeros((uniq.shape[0],), dtype=torch.float64, device=p.device)
    cnt.index_add_(0, inv, torch.ones_like(key, dtype=torch.float64))
    return out / cnt[:, None]


def transform(points, T):
    T = torch.as_tensor(T, dtype=torch.float64, device=points.device)
    return points.to(torch.float64) @ T[:3, :3].T + T[:3, 3]


def write_ply(path, points):
    """Binary little-endian PLY with double-precision vertices (what open3d writes for a
    float64 PointCloud)."""
    pts = np.ascontiguousarray(points.detach().cpu().numpy() if torch.is_tensor(points) else points,
                               dtype="<f8")
    header = ("ply\nformat binary_little_endian 1.0\ncomment pointreggpt_b200\n"
              "element vertex %d\nproperty double x\nproperty double y\nproperty double z\n"
              "end_header\n" % pts.shape[0])
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(pts.tobytes())


def read_ply(path):
    with open(path, "rb") as f:
        n = 0
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        return np.frombuffer(f.read(n * 24), dtype="<f8").reshape(n, 3)

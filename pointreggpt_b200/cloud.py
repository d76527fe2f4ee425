"""Point-cloud post-processing the reference delegates to open3d (not available offline):
axis-aligned crop, voxel-grid centroid down-sampling and a binary PLY writer.

Semantics follow Open3D 0.17 as used at SDD:2486-2500, 2640-2680: `crop` keeps points on the
box boundary; `voxel_down_sample` averages the points of each voxel of a grid anchored at
`min_bound - voxel/2`; the order of the output points is unspecified there (hash-map order),
here it is the sorted voxel index.
"""
import numpy as np
import torch


def crop(points, min_bound, max_bound):
    lo = torch.as_tensor(min_bound, dtype=points.dtype, device=points.device)
    hi = torch.as_tensor(max_bound, dtype=points.dtype, device=points.device)
    keep = ((points >= lo) & (points <= hi)).all(dim=-1)
    return points[keep]


def voxel_down_sample(points, voxel_size):
    """open3d `voxel_down_sample` (SDD:2492, 2652, 2676; generate_gt.py:75-76) through the hand-written
    kernels (csrc/cloud.cu, prg_voxel_downsample_f64): one hash-grid pass with order-independent
    fixed-point sums; output in sorted-voxel-index order.  Centroids agree with the float64 mean to
    2e-11 m and do not depend on the order of the input points.  CUDA tensors only."""
    from . import _ffi
    _ffi.require_cuda(points)
    n = points.shape[0]
    if n == 0:
        return points.to(torch.float64)
    p = points.to(torch.float64).contiguous()
    dev = p.device
    with torch.cuda.device(dev):
        ws_bytes = int(_ffi.lib().prg_voxel_downsample_workspace_bytes(n))
        ws = torch.empty((ws_bytes + 15) // 16 * 2, dtype=torch.int64, device=dev)
        cent = torch.empty((n, 3), dtype=torch.float64, device=dev)
        keys = torch.empty((n,), dtype=torch.int64, device=dev)
        ce = torch.empty((2,), dtype=torch.int32, device=dev)
        _ffi.check(_ffi.lib().prg_voxel_downsample_f64(_ffi.ptr(p), n, float(voxel_size), _ffi.ptr(cent),
                                                       _ffi.ptr(keys), _ffi.ptr(ce), _ffi.ptr(ws), ws_bytes,
                                                       _ffi.stream(dev)))
    m, err = ce.tolist()
    if err:
        raise _ffi.PrgError("voxel_down_sample: non-finite point or more than 2^21 voxels along an axis")
    order = torch.argsort(keys[:m])
    return cent[:m][order]


voxel_down_sample_native = voxel_down_sample     # round-1 name


def voxel_down_sample_torch(points, voxel_size):
    """The same operation written with torch ops (unique + index_add), kept as an independent
    cross-check for the tests; not used by the product path.  The voxel index divides by a TENSOR:
    on CUDA `tensor / python_float` is evaluated as `tensor * (1 / float)`, which moves points that
    sit on voxel faces into the neighbouring voxel (the round-1 hardware failures)."""
    if points.shape[0] == 0:
        return points
    p = points.to(torch.float64)
    vs = torch.tensor(float(voxel_size), dtype=torch.float64, device=p.device)
    origin = p.min(dim=0).values - vs * 0.5
    idx = torch.floor((p - origin) / vs).to(torch.int64)
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


_PLY_TYPES = {"double": "<f8", "float64": "<f8", "float": "<f4", "float32": "<f4",
              "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1", "ushort": "<u2", "uint16": "<u2",
              "short": "<i2", "int16": "<i2", "uint": "<u4", "uint32": "<u4", "int": "<i4", "int32": "<i4"}


def read_ply(path):
    """(N,3) float64 vertex positions of a binary little-endian or ascii PLY.  The header is parsed:
    vertex properties of any scalar type and order (float positions, normals, colours as open3d or
    the reference may write them) are handled; list properties on the vertex element, big-endian
    files and a missing x/y/z raise ValueError instead of being misread."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt, n, props, elem, before = None, None, [], None, 0
        while True:
            raw = f.readline()
            if not raw:
                raise ValueError("%s: truncated PLY header" % path)
            tok = raw.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elem = tok[1]
                if elem == "vertex":
                    n = int(tok[2])
                elif n is None and int(tok[2]) > 0:
                    before += 1
            elif tok[0] == "property" and elem == "vertex":
                if tok[1] == "list" or tok[1] not in _PLY_TYPES:
                    raise ValueError("%s: unsupported vertex property %r" % (path, " ".join(tok[1:])))
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if n is None or before or not all(k in names for k in "xyz"):
            raise ValueError("%s: the vertex element must come first and have x, y, z" % path)
        if fmt == "binary_little_endian":
            dt = np.dtype(props)
            buf = f.read(n * dt.itemsize)
            if len(buf) != n * dt.itemsize:
                raise ValueError("%s: truncated vertex data" % path)
            rec = np.frombuffer(buf, dtype=dt)
            return np.stack([rec[k].astype(np.float64) for k in "xyz"], axis=1).reshape(n, 3)
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=n, ndmin=2) if n else np.zeros((0, len(names)))
            if rows.shape != (n, len(names)):
                raise ValueError("%s: truncated vertex data" % path)
            return np.stack([rows[:, names.index(k)] for k in "xyz"], axis=1).astype(np.float64)
        raise ValueError("%s: unsupported PLY format %r" % (path, fmt))

"""Multi-GPU plumbing for the generation path: pairs are independent, so the scene range is
split across ranks (one process per GPU) and the only collective on the path is one weight
broadcast at start-up (NCCL over NVLink; gloo in the CPU tests).

The reference's `Generator.generate` does not shard by rank: multi-GPU generation there means
launching the CLI N times with disjoint -start/-stop.
"""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size) from torch.distributed or the torchrun environment."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(start, stop, rank, world_size):
    """Contiguous block split of [start, stop): rank r gets ceil(n/W) scenes (the last ranks
    may get fewer or none).  Scene indices stay globally unique, so output paths
    `scene-%06d` and the skip-if-exists resume logic (SDD:2371-2381) work unchanged."""
    n = max(0, stop - start)
    per = (n + world_size - 1) // world_size
    lo = min(stop, start + rank * per)
    hi = min(stop, lo + per)
    return lo, hi


def broadcast_weights(modules, src=0):
    """Broadcast every parameter and buffer of `modules` from rank `src` as ONE flat message
    per dtype; every native network inside `modules` drops its packed blob and device handle, so the
    next call packs the received parameters."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = []
    for m in modules:
        tensors += list(m.parameters()) + list(m.buffers())
    total = 0
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for (dtype, device), ts in by_dtype.items():
        with torch.no_grad():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            dist.broadcast(flat, src=src)
            off = 0
            for t in ts:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))      # in-place on the parameter: bumps _version
                off += n
        total += flat.numel() * flat.element_size()
    # a network that was already packed (any forward / sample before the broadcast) must re-pack
    for m in modules:
        for sub in m.modules():
            if hasattr(sub, "invalidate"):
                sub.invalidate()
    return total


def sum_counters(values, device):
    """All-reduce a few Python numbers (pairs done, seconds) for the throughput report."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.tolist()

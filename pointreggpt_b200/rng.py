"""Seeds that depend on WHAT is generated, not on where or in which batch it is generated.

The reference never seeds: `torch.randn` / `numpy.random` draws depend on process history
(SDD:2526, 1293, 1279).  Here every scene sample gets its own Philox key
`scene_seed(base_seed, absolute scene index, sample index)`, used for the sampler's Gaussian draws on
the device (csrc/elementwise.cu k_fill_normal / k_net_tail) and for the random camera pose on the
host, so scene k's outputs are identical at any world size, batch size or -start/-stop split
(SURVEY.md section 8e).
"""
import numpy as np

_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def mix(*values):
    """64-bit hash of a tuple of integers (splitmix64 chained)."""
    h = 0x243F6A8885A308D3
    for v in values:
        h = _splitmix64(h ^ (int(v) & _M64))
    return h


def scene_seed(base_seed, scene_index, sample_index=0):
    return mix(base_seed, scene_index, sample_index)


def image_seeds(seed, count):
    """Per-image keys of one `sample()` call given a single integer: image i gets mix(seed, i)."""
    return [mix(seed, i) for i in range(count)]


def scene_rng(base_seed, scene_index, sample_index=0):
    """numpy Generator for the host-side draws (pose) of one scene sample."""
    return np.random.Generator(np.random.Philox(key=scene_seed(base_seed, scene_index, sample_index)))


def fill_normal(count, per_image, seeds, offset=0, device="cuda"):
    """(count, per_image) float32 N(0,1) draws from the sampler's device generator
    (prg_fill_normal_f32): row b = Philox stream keyed by seeds[b], counters offset ... ."""
    import torch
    from . import _ffi
    out = torch.empty((count, per_image), dtype=torch.float32, device=device)
    assert len(seeds) == count
    _ffi.require_cuda(out)
    _ffi.check(_ffi.lib().prg_fill_normal_f32(_ffi.ptr(out), count, per_image, _ffi.seed_array(seeds),
                                              _ffi.c_uint64(int(offset)), _ffi.stream(out)))
    return out

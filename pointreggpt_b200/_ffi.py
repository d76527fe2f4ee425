"""ctypes binding of libprg.so (the C ABI declared in include/prg.h).

PyTorch is used here only for device memory and streams: every call forwards raw
``data_ptr()`` values and the current CUDA stream.  There is no CPU fallback -- if the
library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRG_LIB_PATH: load another build of the library (A/B measurements of two builds on one box)
LIB_PATH = os.environ.get("PRG_LIB_PATH") or os.path.join(_HERE, "libprg.so")

c_void_p, c_int, c_float, c_int64, c_size_t, c_uint64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64, ctypes.c_size_t,
    ctypes.c_uint64)


class PrgError(RuntimeError):
    pass


class Step(ctypes.Structure):
    """struct prg_step (include/prg.h)."""
    _fields_ = [("t", c_int), ("kind", c_int), ("add_noise", c_int), ("unnormalize", c_int),
                ("c0", c_float), ("c1", c_float), ("c2", c_float), ("c3", c_float),
                ("c4", c_float)]


class Profile(ctypes.Structure):
    """struct prg_profile (include/prg.h)."""
    _fields_ = [("name", ctypes.c_char * 32), ("launches", c_uint64), ("ms", ctypes.c_double),
                ("forwards", c_uint64)]


STEP_P_SAMPLE, STEP_DDIM, STEP_DDIM_LAST, STEP_REFINE_P, STEP_REFINE_DDIM = range(5)


# name -> (restype, argtypes); must list every symbol include/prg.h declares.
SIGNATURES = {
    "prg_last_error": (ctypes.c_char_p, []),
    "prg_abi_version": (c_int, []),
    "prg_launch_count": (c_uint64, []),
    "prg_reproject_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_void_p]),
    "prg_pc2depth_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "prg_depth2pc_f32": (c_int, [c_void_p, c_void_p, c_float, c_float, c_int, c_float, c_void_p,
                                 c_void_p, c_int, c_int, c_int, c_void_p]),
    "prg_depth2pc_compact_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                         c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p]),
    "prg_net_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_size_t, c_int, c_int,
                               c_int]),
    "prg_net_destroy": (None, [c_void_p]),
    "prg_net_device_bytes": (c_size_t, [c_void_p]),
    "prg_unet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p]),
    "prg_maskunet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int,
                                     c_void_p]),
    "prg_sampler_run": (c_int, [c_void_p, ctypes.POINTER(Step), c_int, c_void_p, c_void_p,
                                c_void_p, ctypes.POINTER(c_uint64), c_void_p, c_int, c_void_p]),
    "prg_fill_normal_f32": (c_int, [c_void_p, c_int, c_int64, ctypes.POINTER(c_uint64), c_uint64,
                                    c_void_p]),
    "prg_occlusion_filter_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "prg_voxel_downsample_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64]),
    "prg_voxel_downsample_f64": (c_int, [c_void_p, ctypes.c_int64, ctypes.c_double, c_void_p, c_void_p,
                                         c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "prg_overlap_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64]),
    "prg_overlap_count_f64": (c_int, [c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64, ctypes.c_double,
                                      c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "prg_profile_set": (c_int, [c_int]),
    "prg_profile_read": (c_int, [ctypes.POINTER(Profile), c_int, c_int]),
    "prg_test_frcp_exhaustive": (c_int, [c_float, c_float, c_void_p, c_void_p]),
    "prg_test_mma_rate": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "prg_profile_ops": (c_int, [ctypes.c_char_p, c_int, c_int]),
    "prg_test_conv_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises PrgError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise PrgError(
                "libprg.so is missing (%s); build it with `python -m pointreggpt_b200.build`. "
                "There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        if l.prg_abi_version() != 2:
            raise PrgError("libprg.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise PrgError("libprg error %d: %s" % (rc, lib().prg_last_error().decode()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream(ref=None):
    """The current CUDA stream of the device `ref` (a tensor or a device) lives on -- not of whatever
    device happens to be current; the library's entry points select the device of their buffers
    themselves (csrc/common.cuh PtrDeviceGuard / net.cu DeviceGuard)."""
    dev = ref.device if torch.is_tensor(ref) else ref
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PrgError("pointreggpt_b200 runs on CUDA tensors only (got a %s tensor); "
                           "there is no CPU fallback" % t.device)


def seed_array(seeds):
    """Host uint64 array for the per-image Philox keys."""
    vals = [int(v) & 0xFFFFFFFFFFFFFFFF for v in seeds]
    return (c_uint64 * len(vals))(*vals)


def launch_count():
    return int(lib().prg_launch_count())


def profile_set(every):
    check(lib().prg_profile_set(int(every)))


def profile_read(reset=True):
    """{family: dict(launches, ms, forwards)} of the sampled launches since the last reset."""
    arr = (Profile * 16)()
    n = lib().prg_profile_read(arr, 16, int(reset))
    return {arr[i].name.decode(): dict(launches=int(arr[i].launches), ms=float(arr[i].ms),
                                       forwards=int(arr[i].forwards)) for i in range(n)}


def profile_ops(reset=True):
    """[(label, launches, ms, flops_per_image)] per layer of the sampled network evaluations."""
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib().prg_profile_ops(buf, len(buf), int(reset))
    rows = []
    for ln in buf.raw[:n].decode().splitlines():
        lab, launches, ms, fl = ln.split("\t")
        rows.append((lab, int(launches), float(ms), float(fl)))
    return rows

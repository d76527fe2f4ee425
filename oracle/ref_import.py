"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Only usable where ``/root/reference`` exists (the build container).  It is used
by ``oracle/make_golden.py`` to mint the fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent) to
pin the restatements in ``oracle/``.  Nothing in the product package, the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file.

The reference imports seven third-party packages that are not installed
offline.  None of them is touched by ``Unet``,
``GaussianDiffusion``, ``MaskUnet`` or the geometry helpers, so empty stand-in
modules are registered before the import.  The reference tree is never
modified or copied.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PRG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "denoising_diffusion_pytorch", "successive_ddnm_diffusion.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns (sdd, dc): the reference diffusion and depth-correction modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    class _Missing:  # placeholder for classes the hot path never instantiates
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")

    _stub("ema_pytorch", EMA=_Missing)
    _stub("accelerate", Accelerator=_Missing)
    _stub("open3d")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    cm = _stub("matplotlib.cm")
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = plt
    if not hasattr(mpl, "cm"):
        mpl.cm = cm
    _stub("pytorch_fid")
    _stub("pytorch_fid.inception", InceptionV3=_Missing)
    _stub("pytorch_fid.fid_score", calculate_frechet_distance=None)
    _stub("imageio")
    _stub("coloredlogs")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sdd = importlib.import_module(
        "denoising_diffusion_pytorch.successive_ddnm_diffusion")
    dc = importlib.import_module("depth_correction_pytorch.depth_correction")
    return sdd, dc

"""Host emulation of barrier-free CUDA kernels: the kernel text is cut out of the .cu file, compiled
with g++ against tests/emu/cuda_shim.h and every thread of the launch grid is run in turn.  This
checks the thread decomposition, border handling and rounding order against the oracle on CPU --
useful for kernels written when no GPU is at hand.  The GPU parity tests remain the real gate."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import geometry_ref as G
from pointreggpt_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CU = os.path.join(ROOT, "pointreggpt_b200", "csrc", "geometry.cu")

DRIVER = '''
extern "C" void emu_occlusion(const float* d, const uint8_t* m, float* o, int B, int H, int W) {
  const unsigned gx = (W + 255) / 256, gy = (H + 4 * kOccRows - 1) / (4 * kOccRows);   // as the ABI entry
  for (unsigned z = 0; z < (unsigned)B; ++z)
    for (unsigned y = 0; y < gy; ++y)
      for (unsigned x = 0; x < gx; ++x)
        for (unsigned t = 0; t < 256; ++t) {
          blockIdx = {x, y, z};
          threadIdx = {t, 0, 0};
          k_occlusion_filter(d, m, o, H, W);
        }
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = open(CU).read()
    a = src.index("constexpr int kOccRows")
    b = src.index("// ------------------------------------------------------------------ point_cloud")
    d = tmp_path_factory.mktemp("emu")
    cpp = d / "occ.cpp"
    cpp.write_text('#include "cuda_shim.h"\n' + src[a:b] + DRIVER)
    so = d / "occ.so"
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "emu"), "-o", str(so), str(cpp)])
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("shape", [(2, 256, 256), (1, 480, 640), (2, 33, 47), (1, 5, 7), (1, 40, 260),
                                   (1, 1, 1), (1, 70, 4)])
def test_occlusion_filter_kernel_logic(emu, shape):
    B, H, W = shape
    d01 = S.synthetic_depth_batch(317, B, H, W)
    K = S.synthetic_intrinsics(B, 256 if H == 256 else None, seed=17).copy()
    if (H, W) not in ((256, 256), (480, 640)):
        K[:, 0, 0] = K[:, 1, 1] = 1.2 * W
        K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    rd, rm = G.reproject((d01 * 10).numpy(), K, S.synthetic_poses(B, seed=18))
    want, _ = G.occlusion_filter(rd, rm)
    d = np.ascontiguousarray(rd.reshape(B, H, W))
    m = np.ascontiguousarray(rm.reshape(B, H, W).astype(np.uint8))
    got = np.full_like(d, -777.0)
    vp = ctypes.c_void_p
    emu.emu_occlusion(d.ctypes.data_as(vp), m.ctypes.data_as(vp), got.ctypes.data_as(vp), B, H, W)
    assert np.array_equal(got.reshape(rd.shape).view(np.uint32), want.view(np.uint32))

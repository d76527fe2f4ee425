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


GEOM_DRIVER = '''
static void run_block(unsigned x, unsigned y, unsigned first, unsigned last, void (*fn)(void*), void* arg) {
  for (unsigned t = first; t < last; ++t) {
    blockIdx = {x, y, 0};
    threadIdx = {t, 0, 0};
    fn(arg);
  }
}
// prg_reproject_f32: the work items of the fused persistent kernel, run one after another in the order
// the kernel deals them (every dependency of an item is an earlier item); ring of R slots, lag D as
// the ABI entry plans them, but with a ring small enough that slots are reused even in tiny batches.
extern "C" void emu_reproject(const float* depth, const float* K, const float* pose, float lo, float hi,
                              float* out, uint8_t* mask, unsigned* scratch, int B, int H, int W, int R, int item_chunks) {
  RpPlan pl;
  pl.B = B; pl.H = H; pl.W = W; pl.HW = H * W;
  pl.st = rp_steps(W);
  pl.item_px = item_chunks * kRpChunkPx;
  pl.items = (pl.HW + pl.item_px - 1) / pl.item_px;
  pl.D = R / 2 < 1 ? 1 : R / 2;
  if (pl.D > B) pl.D = B;
  pl.R = 2 * pl.D;
  pl.total = (long long)(B + pl.D) * pl.items;
  memset(scratch, 0xFF, (size_t)pl.R * pl.HW * 4);
  blockDim = {(unsigned)kRpThreads, 1, 1};
  gridDim = {1, 1, 1};
  blockIdx = {0, 0, 0};
  for (long long p = 0; p < pl.total; ++p)
    for (unsigned t = 0; t < (unsigned)kRpThreads; ++t) {
      threadIdx = {t, 0, 0};
      if ((long long)pl.HW * 9 < 400) rp_run_item<true>(p, depth, K, pose, lo, hi, scratch, out, mask, pl);
      else rp_run_item<false>(p, depth, K, pose, lo, hi, scratch, out, mask, pl);
    }
  // the call must hand the ring back empty
  for (size_t i = 0; i < (size_t)pl.R * pl.HW; ++i)
    if (scratch[i] != 0xFFFFFFFFu) { out[0] = -12345.f; break; }
}
struct PArgs { const float* pc; const uint8_t* valid; const int64_t* off; int64_t total; const float *K, *pose; unsigned* z; int B, H, W; };
static void call_pc2d(void* p) {
  PArgs& a = *(PArgs*)p;
  k_pc2depth_splat(a.pc, a.valid, a.off, a.total, a.K, a.pose, a.z, a.B, a.H, a.W);
}
// prg_pc2depth_f32: fill 0xFF, splat the ragged clouds (offsets staged in shared memory), finalise
extern "C" void emu_pc2depth(const float* pc, const uint8_t* valid, const int64_t* off, long long total,
                             const float* K, const float* pose, float* out, uint8_t* mask, unsigned* scratch,
                             int B, int H, int W, int gx) {
  const size_t n = (size_t)B * H * W;
  memset(out, 0xFF, n * 4);
  blockDim = {256, 1, 1};
  gridDim = {(unsigned)gx, 1, 1};
  for (unsigned x = 0; x < (unsigned)gx; ++x) {
    PArgs a{pc, valid, off, total, K, pose, scratch, B, H, W};
    run_block(x, 0, 0, 256, call_pc2d, &a);        // fills the staged offsets
    a.z = (unsigned*)out;
    run_block(x, 0, 0, 256, call_pc2d, &a);
  }
  gridDim = {(unsigned)((n + 1023) / 1024), 1, 1};
  for (unsigned x = 0; x < gridDim.x; ++x)
    for (unsigned t = 0; t < 256; ++t) {
      blockIdx = {x, 0, 0};
      threadIdx = {t, 0, 0};
      k_zbuf_finalize((unsigned*)out, mask, n);
    }
}
struct DArgs { const float *depth, *K; float lo, hi; int use_clip; float invalid; float* pc; uint8_t* valid; int HW, W; };
static void call_d2pc(void* p) {
  DArgs& a = *(DArgs*)p;
  if ((a.W & 3) == 0) k_depth2pc_vec(a.depth, a.K, a.lo, a.hi, a.use_clip, a.invalid, a.pc, a.valid, a.HW, a.W);   // as prg_depth2pc_f32
  else k_depth2pc(a.depth, a.K, a.lo, a.hi, a.use_clip, a.invalid, a.pc, a.valid, a.HW, a.W);
}
extern "C" void emu_depth2pc(const float* depth, const float* K, float lo, float hi, int use_clip, float invalid,
                             float* pc, uint8_t* valid, int B, int H, int W, int gx) {
  const int HW = H * W;
  blockDim = {256, 1, 1};
  gridDim = {(unsigned)gx, (unsigned)B, 1};
  DArgs a{depth, K, lo, hi, use_clip, invalid, pc, valid, HW, W};
  for (unsigned y = 0; y < (unsigned)B; ++y)
    for (unsigned x = 0; x < (unsigned)gx; ++x) {
      run_block(x, y, 0, 256, call_d2pc, &a);      // fills the per-warp transpose buffers
      run_block(x, y, 0, 256, call_d2pc, &a);      // reads them: the real output
    }
}
'''


VOX_DRIVER = '''
template <class F> static void run_grid(unsigned blocks, F f) {
  blockDim = {256, 1, 1};
  gridDim = {blocks, 1, 1};
  for (unsigned x = 0; x < blocks; ++x)
    for (unsigned t = 0; t < 256; ++t) {
      blockIdx = {x, 0, 0};
      threadIdx = {t, 0, 0};
      f();
    }
}
// the body of prg_voxel_downsample_f64 on host memory; thread order reversed on request to show that
// the result does not depend on the order of arrival
extern "C" long long emu_voxel(const double* pts, long long n, double voxel, double* cent, long long* keys_out,
                               int* count_err, unsigned char* ws, int blocks) {
  const unsigned long long cap = vox_capacity(n);
  unsigned long long* minb = (unsigned long long*)ws;
  unsigned long long* keys = minb + 4;
  unsigned long long* sums = keys + cap;
  int* counts = (int*)(sums + 3 * cap);
  memset(minb, 0xFF, 32 + cap * 8);
  memset(sums, 0, cap * 28);
  count_err[0] = count_err[1] = 0;
  run_grid(blocks, [&] { k_vox_bounds(pts, n, minb); });
  run_grid(blocks, [&] { k_vox_insert(pts, n, voxel, minb, keys, sums, counts, cap - 1, count_err); });
  run_grid((unsigned)((cap + 1023) / 1024), [&] { k_vox_emit(minb, keys, sums, counts, cap, voxel, cent, keys_out, count_err); });
  return (long long)cap;
}
// the body of prg_overlap_count_f64
extern "C" void emu_overlap(const double* q, long long nq, const double* t, long long nt, double radius,
                            int* count_err, unsigned char* ws, int blocks) {
  const unsigned long long cap = vox_capacity(nt);
  unsigned long long* keys = (unsigned long long*)ws;
  int* head = (int*)(keys + cap);
  int* next = head + cap;
  memset(keys, 0xFF, cap * 12);
  count_err[0] = count_err[1] = 0;
  run_grid(blocks, [&] { k_ovl_build(t, nt, radius, keys, head, next, cap - 1, count_err); });
  run_grid(blocks, [&] { k_ovl_query(q, nq, t, radius, keys, head, next, cap - 1, count_err); });
}
'''


def _compile(tmp, name, text):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cpp = tmp / (name + ".cpp")
    cpp.write_text('#include "cuda_shim.h"\n' + text)
    so = tmp / (name + ".so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "emu"), "-o", str(so), str(cpp)])
    return ctypes.CDLL(str(so))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    src = open(CU).read()
    a = src.index("constexpr int kOccRows")
    b = src.index("// ------------------------------------------------------------------ point_cloud")
    return _compile(tmp_path_factory.mktemp("emu"), "occ", src[a:b] + DRIVER)


@pytest.fixture(scope="module")
def vox(tmp_path_factory):
    src = open(os.path.join(ROOT, "pointreggpt_b200", "csrc", "cloud.cu")).read()
    a = src.index("constexpr unsigned long long kVoxEmpty")
    b = src.index("}  // namespace prg")
    return _compile(tmp_path_factory.mktemp("emu"), "vox", src[a:b] + VOX_DRIVER)


@pytest.fixture(scope="module")
def geom(tmp_path_factory):
    """Device helpers (division, unproject, rigid, splat) + the work items of k_reproject_fused,
    k_pc2depth_splat / k_zbuf_finalize and k_depth2pc, exactly as they stand in geometry.cu."""
    src = open(CU).read()
    a = src.index("constexpr unsigned kEmpty")
    b = src.index("// ------------------------------------------------------------------ pc2depth (ragged)")
    c = src.index("// ------------------------------------------------------------------ depth2pc (dense)")
    d = src.index("// ------------------------------------------------------------------ occlusion_filter")
    text = src[a:c] + src[c:d]
    # dynamic shared memory has no host counterpart: a fixed array of the ABI's maximum (B <= 4096)
    assert "extern __shared__ int64_t s_off[];" in text
    text = text.replace("extern __shared__ int64_t s_off[];", "static int64_t s_off[4097];")
    return _compile(tmp_path_factory.mktemp("emu"), "geom", text + GEOM_DRIVER)


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


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _geo_inputs(B, H, W, seed, extreme=False):
    d01 = S.synthetic_depth_batch(400 + seed, B, H, W)
    K = S.synthetic_intrinsics(B, 256 if H == 256 else None, seed=seed).copy()
    if (H, W) not in ((256, 256), (480, 640)):
        K[:, 0, 0] = K[:, 1, 1] = 1.2 * W
        K[:, 0, 2], K[:, 1, 2] = W / 2, H / 2
    dm = (d01 * 10).numpy().reshape(B, H, W).copy()
    if extreme:
        special = np.array([0.0, -0.0, -1.5, 1e-42, 1e-30, 1e-12, 1e12, 1e30, 3e38, np.inf, -np.inf,
                            np.nan, 1.0, 2.5e-7, 7.7e19, 65504.0], np.float32)
        flat = dm.reshape(B, -1)
        flat[:, ::3] = special[np.arange(flat[:, ::3].shape[1]) % special.size]
    return dm, np.ascontiguousarray(K, np.float32), np.ascontiguousarray(S.synthetic_poses(B, seed=seed + 1), np.float32)


def _same_bits_or_nan(a, b):
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("shape,extreme", [((2, 256, 256), False), ((1, 480, 640), False),
                                           ((2, 33, 47), False), ((2, 64, 96), True)])
def test_reproject_kernel_numerics(geom, shape, extreme):
    """The fast exact division, the rigid transform and the z-buffer of the CUDA source, compiled for
    the host, against the oracle: bit-exact, including signed zeros / denormals / inf / NaN depths."""
    B, H, W = shape
    dm, K, P = _geo_inputs(B, H, W, 3, extreme)
    with np.errstate(all="ignore"):
        want_d, want_m = G.reproject(dm, K, P)
    out = np.empty((B, H, W), np.float32)
    mask = np.empty((B, H, W), np.uint8)
    scratch = np.full((2, H, W), 0, np.uint32)            # ring of two slots: reused within the batch
    geom.emu_reproject(_vp(dm), _vp(K), _vp(P), ctypes.c_float(0.0), ctypes.c_float(10.0), _vp(out), _vp(mask),
                       _vp(scratch), B, H, W, 2, 1 + B % 2)
    assert _same_bits_or_nan(out.reshape(want_d.shape), want_d)
    assert np.array_equal(mask.reshape(want_m.shape).astype(bool), want_m) and want_m.any()


@pytest.mark.parametrize("shape,extreme", [((2, 256, 256), False), ((1, 480, 640), False),
                                           ((1, 5, 7), False), ((2, 64, 96), True)])
def test_depth2pc_kernel_numerics(geom, shape, extreme):
    B, H, W = shape
    dm, K, _ = _geo_inputs(B, H, W, 7, extreme)
    gx = (H * W + 1023) // 1024
    for clip, inv in [((0.0, 10.0), float("nan")), ((0.5, 10.0), 0.0), (None, float("nan"))]:
        with np.errstate(all="ignore"):
            want_pc, want_v = G.depth2pc(dm, K, clip=clip, invalid=inv)
        pc = np.full((B, H * W, 3), -777.0, np.float32)
        valid = np.full((B, H * W), 7, np.uint8)
        lo, hi = clip if clip is not None else (0.0, 0.0)
        geom.emu_depth2pc(_vp(dm), _vp(K), ctypes.c_float(lo), ctypes.c_float(hi), int(clip is not None),
                          ctypes.c_float(inv), _vp(pc), _vp(valid), B, H, W, gx)
        assert _same_bits_or_nan(pc, want_pc)
        assert np.array_equal(valid.astype(bool), want_v)


def _cloud(n, seed, extent=1.0):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-extent, extent, (n, 3)) + np.array([0.3, -1.1, 2.5])
    pts[: n // 5] = np.round(pts[: n // 5] / 0.025) * 0.025        # points on voxel faces
    pts[n // 5: n // 4] = pts[0]                                    # duplicates
    return np.ascontiguousarray(pts)


@pytest.mark.parametrize("n,voxel,blocks", [(20000, 0.025, 20), (20000, 0.1, 7), (5000, 0.002, 5), (3, 0.5, 1),
                                            (1, 0.1, 1), (4097, 0.05, 64)])
def test_voxel_downsample_kernel_logic(vox, n, voxel, blocks):
    pts = _cloud(n, n)
    want_c, want_k = G.voxel_down_sample(pts, voxel)
    cap = 1024
    while cap < 2 * n:
        cap *= 2
    ws = np.zeros(32 + cap * 36, np.uint8)
    cent = np.zeros((n, 3), np.float64)
    keys = np.zeros((n,), np.int64)
    ce = np.zeros((2,), np.int32)
    got_cap = vox.emu_voxel(_vp(pts), ctypes.c_longlong(n), ctypes.c_double(voxel), _vp(cent), _vp(keys), _vp(ce),
                            _vp(ws), blocks)
    assert got_cap == cap
    m, err = int(ce[0]), int(ce[1])
    assert err == 0 and m == want_k.shape[0]
    if n > 100:
        assert m < n                                               # several points per voxel somewhere
    order = np.argsort(keys[:m])
    assert np.array_equal(keys[:m][order], want_k)
    assert np.abs(cent[:m][order] - want_c).max() < 1e-10          # fixed-point sums: < 2e-11 m


def test_voxel_downsample_kernel_flags_bad_points(vox):
    pts = _cloud(100, 1)
    pts[17, 1] = np.nan
    ws = np.zeros(32 + 1024 * 36, np.uint8)
    cent, keys, ce = np.zeros((100, 3)), np.zeros((100,), np.int64), np.zeros((2,), np.int32)
    vox.emu_voxel(_vp(pts), ctypes.c_longlong(100), ctypes.c_double(0.05), _vp(cent), _vp(keys), _vp(ce), _vp(ws), 1)
    assert ce[1] == 1
    pts = _cloud(100, 2)
    pts[3, 0] += 1e6                                                  # 1e6 / 1e-3 voxels > 2^21
    vox.emu_voxel(_vp(pts), ctypes.c_longlong(100), ctypes.c_double(1e-3), _vp(cent), _vp(keys), _vp(ce), _vp(ws), 1)
    assert ce[1] == 1


@pytest.mark.parametrize("nq,nt,radius,blocks", [(4000, 3000, 0.075, 9), (3000, 4000, 0.0375, 3), (500, 1, 0.5, 1),
                                                 (1, 500, 0.2, 2), (2000, 2000, 1.5, 4)])
def test_overlap_kernel_logic(vox, nq, nt, radius, blocks):
    rng = np.random.default_rng(nq + nt)
    q = np.ascontiguousarray(rng.uniform(-1, 1, (nq, 3)))
    t = np.ascontiguousarray(rng.uniform(-0.5, 1.5, (nt, 3)))
    if nt > 10 and nq > 10:
        t[:5] = q[:5] + np.array([radius, 0, 0])          # exactly at the radius: not a neighbour (strict <)
        t[5:10] = q[5:10]                                 # coincident points
    want = G.overlap_count(q, t, radius)
    cap = 1024
    while cap < 2 * nt:
        cap *= 2
    ws = np.zeros(cap * 12 + nt * 4, np.uint8)
    ce = np.zeros((2,), np.int32)
    vox.emu_overlap(_vp(q), ctypes.c_longlong(nq), _vp(t), ctypes.c_longlong(nt), ctypes.c_double(radius), _vp(ce),
                    _vp(ws), blocks)
    assert ce[1] == 0 and int(ce[0]) == want
    assert 0 < want or nt == 1
    q[0, 0] = np.inf
    vox.emu_overlap(_vp(q), ctypes.c_longlong(nq), _vp(t), ctypes.c_longlong(nt), ctypes.c_double(radius), _vp(ce),
                    _vp(ws), blocks)
    assert ce[1] == 1


_FX = [0.5, 0.999, 1.0, 1.5, 300.0, 585.0, 1e6, 1.0000001e6, 2e6, 1e-3, 1e9, 3e38, 1e-30, -300.0, 0.0, np.inf, np.nan,
       511.99997]
_CX = [0.0, -0.0, 1e-4, 1e-3, 0.999e-3, 128.5, 320.0, 1e6, 1.1e6, -5.0, 1e-30, -1e-3, np.nan, np.inf, 5.0000005]
_Z = np.array([0.0, -0.0, -1.5, 1e-42, 1e-30, 1e-12, 0.9e-9, 1.1e-9, 1e-9, 1e9, 0.9e9, 1.1e9, 1e12, 1e30, 3e38, np.inf,
               -np.inf, np.nan, 1.0, 2.5e-7, 7.7e19, 65504.0, 1e-20, 1e-6, 0.99e-6], np.float32)


@pytest.mark.parametrize("seed", [0, 1])
def test_geometry_kernels_differential_fuzz(geom, seed):
    """Random small images (including widths below four, where four consecutive pixels span several
    rows), intrinsics on and beyond the borders of the fast-division range, depths from the list of
    special values, random clips and poses: the emulated kernels must match the oracle bit for bit.
    (This is the test that found the W < 4 row-wrap bug.)"""
    rng = np.random.default_rng(seed)
    for trial in range(120):
        B, H, W = int(rng.integers(1, 3)), int(rng.integers(1, 40)), int(rng.integers(1, 70))
        if rng.random() < 0.3:
            W = (W // 4 + 1) * 4
        dm = (rng.random((B, H, W)) * 10).astype(np.float32)
        m = rng.random((B, H, W)) < 0.3
        dm[m] = _Z[rng.integers(0, _Z.size, int(m.sum()))]
        K = np.zeros((B, 3, 3), np.float32)
        K[:, 2, 2] = 1
        for b in range(B):
            if rng.random() < 0.5:
                K[b, 0, 0], K[b, 1, 1] = rng.choice(_FX), rng.choice(_FX)
                K[b, 0, 2], K[b, 1, 2] = rng.choice(_CX), rng.choice(_CX)
            else:
                K[b, 0, 0], K[b, 1, 1] = 1.2 * W * rng.uniform(0.5, 2), 1.2 * W * rng.uniform(0.5, 2)
                K[b, 0, 2], K[b, 1, 2] = W / 2 + rng.uniform(-1, 1), H / 2
        P = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
        ang = rng.uniform(-0.3, 0.3, B)
        P[:, 0, 0], P[:, 0, 2], P[:, 2, 0], P[:, 2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
        P[:, :3, 3] = rng.normal(0, 0.3, (B, 3))
        if rng.random() < 0.1:
            P[0, 0, 3] = np.float32(rng.choice([np.inf, np.nan, 1e30]))
        gx = (H * W + 1023) // 1024
        clip = [(0.0, 10.0), (0.5, 10.0), None, (-1.0, 1e30)][int(rng.integers(0, 4))]
        inv = [float("nan"), 0.0][int(rng.integers(0, 2))]
        rclip = [(0.0, 10.0), (0.5, 3.5), (-1.0, 3e38)][int(rng.integers(0, 3))]
        with np.errstate(all="ignore"):
            want_pc, want_v = G.depth2pc(dm, K, clip=clip, invalid=inv)
            want_d, want_m = G.reproject(dm, K, P, clip=rclip)
        pc = np.full((B, H * W, 3), -777.0, np.float32)
        valid = np.full((B, H * W), 7, np.uint8)
        lo, hi = clip if clip is not None else (0.0, 0.0)
        geom.emu_depth2pc(_vp(dm), _vp(K), ctypes.c_float(lo), ctypes.c_float(hi), int(clip is not None),
                          ctypes.c_float(inv), _vp(pc), _vp(valid), B, H, W, gx)
        ctx = "trial %d shape %s K %s clip %s %s" % (trial, (B, H, W), K[:, [0, 1, 0, 1], [0, 1, 2, 2]].tolist(), clip, rclip)
        assert _same_bits_or_nan(pc, want_pc) and np.array_equal(valid.astype(bool), want_v), ctx
        out = np.empty((B, H, W), np.float32)
        mask = np.empty((B, H, W), np.uint8)
        ring = 2 + trial % 3
        scratch = np.full((ring, H, W), 0, np.uint32)
        geom.emu_reproject(_vp(dm), _vp(K), _vp(P), ctypes.c_float(rclip[0]), ctypes.c_float(rclip[1]), _vp(out),
                           _vp(mask), _vp(scratch), B, H, W, ring, 1 + trial % 3)
        assert _same_bits_or_nan(out.reshape(want_d.shape), want_d), ctx
        assert np.array_equal(mask.reshape(want_m.shape).astype(bool), want_m), ctx


def _occlusion_case(rng):
    B, H, W = int(rng.integers(1, 3)), int(rng.integers(1, 70)), int(rng.integers(1, 300))
    if rng.random() < 0.4:
        W = (W // 4 + 1) * 4
    d = (rng.random((B, 1, H, W)) * 3).astype(np.float32)
    d = np.round(d / 0.0125).astype(np.float32) * np.float32(0.0125)     # plateaus: differences hit 0.0375 exactly
    m = rng.random((B, 1, H, W)) < 0.7
    d[~m] = 0
    if rng.random() < 0.2:
        d[rng.random(d.shape) < 0.02] = np.float32(np.inf)
    return d, m


def test_occlusion_filter_kernel_fuzz(emu):
    rng = np.random.default_rng(0)
    for trial in range(150):
        d, m = _occlusion_case(rng)
        B, _, H, W = d.shape
        want, _ = G.occlusion_filter(d, m)
        got = np.full((B, H, W), -7.0, np.float32)
        dd = np.ascontiguousarray(d.reshape(B, H, W))
        mm = np.ascontiguousarray(m.reshape(B, H, W).astype(np.uint8))
        emu.emu_occlusion(_vp(dd), _vp(mm), _vp(got), B, H, W)
        assert _same_bits_or_nan(got.reshape(want.shape), want), (trial, d.shape)


def test_pc2depth_kernel_fuzz(geom):
    """Ragged z-buffer (the Generator path: pose applied in the kernel) against the oracle: random
    clouds with awkward coordinates, validity masks, empty clouds in the batch."""
    rng = np.random.default_rng(3)
    V = np.array([0.0, -0.0, -1.5, 1e-42, 1e-30, 1e-12, 1e9, 1e12, 1e30, 3e38, np.inf, -np.inf, np.nan, 1.0, 0.5, 2.5],
                 np.float32)
    for trial in range(120):
        B, H, W = int(rng.integers(1, 4)), int(rng.integers(1, 40)), int(rng.integers(1, 70))
        sizes = rng.integers(0, 300, B)
        if trial % 7 == 0:
            sizes[int(rng.integers(0, B))] = 0
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        N = int(offs[-1])
        pc = (rng.normal(0, 1, (max(N, 1), 3)) * np.array([1, 1, 0.5]) + np.array([0, 0, 2.0])).astype(np.float32)
        m = rng.random(pc.shape) < 0.1
        pc[m] = V[rng.integers(0, V.size, int(m.sum()))]
        valid = (rng.random(max(N, 1)) < 0.8) if trial % 3 else None
        K = np.zeros((B, 3, 3), np.float32)
        K[:, 2, 2] = 1
        K[:, 0, 0], K[:, 1, 1] = 1.2 * W * rng.uniform(0.5, 2, B), 1.2 * W * rng.uniform(0.5, 2, B)
        K[:, 0, 2], K[:, 1, 2] = W / 2 + rng.uniform(-1, 1, B), H / 2
        if trial % 5 == 0:
            K[0, 0, 0], K[0, 0, 2] = rng.choice(_FX), rng.choice(_CX)
        P = None
        if trial % 2:
            P = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
            ang = rng.uniform(-0.3, 0.3, B)
            P[:, 0, 0], P[:, 0, 2], P[:, 2, 0], P[:, 2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
            P[:, :3, 3] = rng.normal(0, 0.3, (B, 3))
        with np.errstate(all="ignore"):
            want_d, want_m = G.pc2depth(pc[:N], None if valid is None else valid[:N], offs, K, (H, W), pose=P)
        out = np.empty((B, H, W), np.float32)
        mask = np.empty((B, H, W), np.uint8)
        scratch = np.full((B, H, W), 0xFFFFFFFF, np.uint32)
        v8 = None if valid is None else np.ascontiguousarray(valid.astype(np.uint8))
        gx = max(1, (N + 255) // 256)
        geom.emu_pc2depth(_vp(pc), None if v8 is None else _vp(v8), _vp(offs), ctypes.c_longlong(N), _vp(K),
                          None if P is None else _vp(np.ascontiguousarray(P)), _vp(out), _vp(mask), _vp(scratch), B, H, W, gx)
        assert _same_bits_or_nan(out.reshape(want_d.shape), want_d), (trial, B, H, W, N)
        assert np.array_equal(mask.reshape(want_m.shape).astype(bool), want_m), (trial, B, H, W, N)


TAIL_DRIVER_HEAD = '''
struct TailParams {
  const float* x_t; const float* img_cond; const float* noise; float* out;
  int HW, clip_x_start, use_ddnm, sampler, add_noise, unnormalize;
  float c0, c1, c2, c3, c4;
  const unsigned long long* seeds; unsigned long long noise_offset;
};
static float philox_normal(unsigned long long, unsigned long long) { return 0.f; }   // noise is always injected here
static void tail_pixel(const TailParams& t, float net, int b, long long p, size_t o) {
'''
TAIL_DRIVER_TAIL = '''
}
extern "C" void emu_tail(const float* x_t, const float* img_cond, const float* noise, const float* net, float* out,
                         int B, int HW, int kind, int add_noise, int unnormalize, int has_cond,
                         float c0, float c1, float c2, float c3, float c4) {
  // the host side of prg_sampler_run for one step (net.cu): which kinds clamp before pred_noise, which replace
  TailParams t{x_t, has_cond ? img_cond : nullptr, noise, out, HW, 0, 0, kind, add_noise, unnormalize, c0, c1, c2, c3, c4, nullptr, 0};
  t.clip_x_start = (kind == 1 || kind == 2 || kind == 4);
  t.use_ddnm = has_cond && (kind == 0 || kind == 1 || kind == 2);
  for (int b = 0; b < B; ++b)
    for (long long p = 0; p < HW; ++p) {
      const size_t o = (size_t)b * HW + p;
      tail_pixel(t, net[o], b, p, o);
    }
}
'''


@pytest.fixture(scope="module")
def tail(tmp_path_factory):
    """The sampler-step arithmetic of k_net_tail (mode 2), cut out of elementwise.cu verbatim."""
    src = open(os.path.join(ROOT, "pointreggpt_b200", "csrc", "elementwise.cu")).read()
    a = src.index("  const float xt = t.x_t[o];")
    b = src.index("  t.out[o] = xn;") + len("  t.out[o] = xn;")
    return _compile(tmp_path_factory.mktemp("emu"), "tail", TAIL_DRIVER_HEAD + src[a:b] + TAIL_DRIVER_TAIL)


@pytest.mark.parametrize("mode", ["p_sample", "p_sample_refine", "ddim", "ddim_refine", "ddim_eta0", "uncond"])
def test_sampler_step_arithmetic_bit_exact(tail, monkeypatch, mode):
    """Host step tables (diffusion.sampling_steps) + the device's per-pixel sampler arithmetic, run on
    the host with a stand-in network, must reproduce the oracle's sampling loops BIT FOR BIT (the
    oracle is pinned bit-exactly to the reference): given the same network output, the DDNM
    replacement, clamps, posterior / DDIM updates, refine step and unnormalisation are identical."""
    import torch
    from oracle import torch_ref as R
    from pointreggpt_b200 import _ffi, nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    B, SZ = 2, 24
    cfg = dict(p_sample=dict(timesteps=12), p_sample_refine=dict(timesteps=12),
               ddim=dict(timesteps=50, sampling_timesteps=7, ddim_sampling_eta=1.0),
               ddim_refine=dict(timesteps=50, sampling_timesteps=7, ddim_sampling_eta=1.0),
               ddim_eta0=dict(timesteps=40, sampling_timesteps=5, ddim_sampling_eta=0.0),
               uncond=dict(timesteps=9))[mode]
    refine = mode.endswith("refine")
    T = cfg["timesteps"]

    def fake_net(x, t):          # exceeds [-1, 1] so that every clamp is exercised
        return torch.tanh(x * 0.7 + 0.013 * float(t)) * 1.3 - 0.05

    monkeypatch.setattr(R, "unet_forward", lambda sd, x, tt, pc, emu=None: fake_net(x, int(tt[0])))
    g = torch.Generator().manual_seed(4)
    noises = [torch.randn(B, 1, SZ, SZ, generator=g) for _ in range(T + 2)]
    d = torch.rand(B, 1, SZ, SZ, generator=g)
    d[d < 0.4] = 0
    ic = None if mode == "uncond" else torch.cat([d, (d > 0).float()], 1) * 2 - 1
    pc = torch.zeros(B, 4)
    sch = R.make_schedule(T)
    if "sampling_timesteps" in cfg:
        want = R.ddim_sample(None, sch, pc, ic, noises, cfg["sampling_timesteps"], cfg["ddim_sampling_eta"],
                             has_refine_step=refine)
    else:
        want = R.p_sample_loop(None, sch, pc, ic, noises, has_refine_step=refine)
    torch.manual_seed(0)
    diff = GaussianDiffusion(nets.Unet(dim=64, param_cond_dim=4), image_size=SZ, objective="pred_x0",
                             beta_schedule="sigmoid", **cfg)
    x = noises[0].clone().numpy()
    k = 1
    icn = None if ic is None else np.ascontiguousarray(ic.numpy())
    for st in diff.sampling_steps(refine):
        net = np.ascontiguousarray(fake_net(torch.tensor(x), st.t).numpy())
        nz = None
        if st.add_noise:
            nz = np.ascontiguousarray(noises[k].numpy())
            k += 1
        out = np.empty_like(x)
        f = ctypes.c_float
        tail.emu_tail(_vp(x), None if icn is None else _vp(icn), None if nz is None else _vp(nz), _vp(net), _vp(out),
                      B, SZ * SZ, st.kind, st.add_noise, st.unnormalize, int(icn is not None),
                      f(st.c0), f(st.c1), f(st.c2), f(st.c3), f(st.c4))
        x = out
    assert np.array_equal(x.view(np.uint32), want.numpy().view(np.uint32))

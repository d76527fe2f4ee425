// Geometry kernels of the data-generation hot path (HBM-bound, bit-exact).
//
// Reference semantics (SDD = denoising_diffusion_pytorch/successive_ddnm_diffusion.py):
//   depth2pc_tensor  SDD:176-209   x = ((c - cx) * z) / fx      (each op rounded, fp32)
//   pc2depth_tensor  SDD:212-265   c = round_half_even((x * fx) / z + cx); scatter amin
//   reproject_tensor SDD:268-286   matmul(pc, R^T) + t  ==  fma(z,r2, fma(y,r1, x*r0)) + t
//   point_cloud      SDD:122-143   float64 unprojection of the valid pixels, row-major
//
// All fp32 arithmetic uses the _rn intrinsics so nvcc can neither contract nor
// reorder it; the z-buffer is an order-independent atomicMin on the bit pattern of
// the (strictly positive) depth, so the result is deterministic and bit-exact.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace prg {

constexpr unsigned kEmpty = 0xFFFFFFFFu;

struct Intr {
  float fx, fy, cx, cy;
  float rfx, rfy;   // correctly rounded reciprocals of fx, fy
  bool fast;        // intrinsics inside the range the three-instruction division is validated for
};

// Correctly rounded a / b in three instructions, given y = RN(1 / b) (Markstein's correction step):
//   q0 = RN(a y);  r = a - q0 b (exact, one FMA);  q = RN(q0 + r y).
// tools/check_division.py: 0 mismatches against the IEEE quotient on 16 M emulated trials over the
// operand ranges of these kernels; the parity tests compare every pixel with the C oracle, which
// divides.  __fdiv_rn costs ~12 instructions and made the kernels issue-bound (profiles/
// r1_geometry_summary.txt).  Only valid while nothing under- or overflows: callers check the operand
// ranges (Intr::fast, depth_in_range) and take the IEEE division otherwise.  A zero numerator
// yields +0 whatever its sign; div_exact_signed() restores the sign the division gives (b > 0).
__device__ __forceinline__ float div_exact(float a, float b, float y) {
  const float q0 = __fmul_rn(a, y);
  return __fmaf_rn(__fmaf_rn(-q0, b, a), y, q0);
}
__device__ __forceinline__ float div_exact_signed(float a, float b, float y) {
  const unsigned q = __float_as_uint(div_exact(a, b, y));
  return __uint_as_float((q & 0x7fffffffu) | (__float_as_uint(a) & 0x80000000u));
}

// |z| in [1e-9, 1e9] or exactly zero: with Intr::fast the numerators (c - cx) z are zero or in
// [1e-19, 1e16] and the quotients zero or in [1e-25, 1e16] -- no underflow in q0 or in the residual
__device__ __forceinline__ bool depth_in_range(float z) {
  const float az = fabsf(z);
  return (az > 1e-9f && az < 1e9f) || z == 0.f;
}

__device__ __forceinline__ Intr load_intr(const float* __restrict__ K, int b) {
  const float* k = K + b * 9;
  Intr i;
  i.fx = __ldg(k + 0);
  i.fy = __ldg(k + 4);
  i.cx = __ldg(k + 2);
  i.cy = __ldg(k + 5);
  i.rfx = __frcp_rn(i.fx);
  i.rfy = __frcp_rn(i.fy);
  // focal lengths in [1, 1e6]; principal point zero or in [1e-3, 1e6] in magnitude, so that c - cx is
  // zero or at least 1e-10 in magnitude
  const float acx = fabsf(i.cx), acy = fabsf(i.cy);
  i.fast = i.fx >= 1.f && i.fx <= 1e6f && i.fy >= 1.f && i.fy <= 1e6f &&
           (i.cx == 0.f || (acx >= 1e-3f && acx <= 1e6f)) &&
           (i.cy == 0.f || (acy >= 1e-3f && acy <= 1e6f));
  return i;
}

__device__ __forceinline__ void row_col(int i, int W, float inv_w, int& r, int& c) {
  r = __float2int_rz(__fmul_rn((float)i, inv_w));
  c = i - r * W;
  if (c < 0) { c += W; --r; }
  if (c >= W) { c -= W; ++r; }
}

__device__ __forceinline__ void unproject_ieee(int r, int c, float z, const Intr& k, float& x, float& y) {
  x = __fdiv_rn(__fmul_rn(__fsub_rn((float)c, k.cx), z), k.fx);
  y = __fdiv_rn(__fmul_rn(__fsub_rn((float)r, k.cy), z), k.fy);
}

// requires k.fast && depth_in_range(z)
__device__ __forceinline__ void unproject_fast(int r, int c, float z, const Intr& k, float& x, float& y) {
  x = div_exact_signed(__fmul_rn(__fsub_rn((float)c, k.cx), z), k.fx, k.rfx);
  y = div_exact_signed(__fmul_rn(__fsub_rn((float)r, k.cy), z), k.fy, k.rfy);
}

__device__ __forceinline__ void unproject(int r, int c, float z, const Intr& k, float& x, float& y) {
  if (k.fast && depth_in_range(z)) unproject_fast(r, c, z, k, x, y);
  else unproject_ieee(r, c, z, k, x, y);
}

// torch.matmul(pc, R^T) + t (SDD:279).  For every realistic size ATen's bmm accumulates like
// fma(z, r2, fma(y, r1, x * r0)); for maps of at most 44 pixels (N * 9 < 400) it runs a scalar loop
// that rounds every product and sum on its own.  kScalarBmm selects that form so that even tiny test
// maps match the reference bit for bit; it is a separate kernel instantiation, the real one is untouched.
template <bool kScalarBmm = false>
__device__ __forceinline__ void rigid(const float* __restrict__ P, float& x, float& y, float& z) {
  float xn, yn, zn;
  if (kScalarBmm) {
    xn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, P[0]), __fmul_rn(y, P[1])), __fmul_rn(z, P[2])), P[3]);
    yn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, P[4]), __fmul_rn(y, P[5])), __fmul_rn(z, P[6])), P[7]);
    zn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, P[8]), __fmul_rn(y, P[9])), __fmul_rn(z, P[10])), P[11]);
  } else {
    xn = __fadd_rn(__fmaf_rn(z, P[2], __fmaf_rn(y, P[1], __fmul_rn(x, P[0]))), P[3]);
    yn = __fadd_rn(__fmaf_rn(z, P[6], __fmaf_rn(y, P[5], __fmul_rn(x, P[4]))), P[7]);
    zn = __fadd_rn(__fmaf_rn(z, P[10], __fmaf_rn(y, P[9], __fmul_rn(x, P[8]))), P[11]);
  }
  x = xn;
  y = yn;
  z = zn;
}

__device__ __forceinline__ void splat(float x, float y, float z, const Intr& k, int H, int W,
                                      unsigned* __restrict__ zimg) {
  float qx, qy;
  if (k.fast && z > 1e-6f && z < 1e12f) {
    // one IEEE reciprocal shared by both divisions by z.  Quotients too small for the residual to be
    // exact vanish in the sum with cx / cy (zero or >= 1e-3); an overflowing one gives NaN instead of
    // +-inf and both fail the bounds test below.
    const float rz = __frcp_rn(z);
    qx = div_exact(__fmul_rn(x, k.fx), z, rz);
    qy = div_exact(__fmul_rn(y, k.fy), z, rz);
  } else {                                 // non-positive / extreme depths: plain divisions
    qx = __fdiv_rn(__fmul_rn(x, k.fx), z);
    qy = __fdiv_rn(__fmul_rn(y, k.fy), z);
  }
  float cf = rintf(__fadd_rn(qx, k.cx));
  float rf = rintf(__fadd_rn(qy, k.cy));
  bool ok = (cf >= 0.f) && (cf < (float)W) && (rf >= 0.f) && (rf < (float)H) && (z > 0.f);
  if (ok) atomicMin(zimg + (int)rf * W + (int)cf, __float_as_uint(z));
}

// ------------------------------------------------------------------ reproject
// One persistent kernel, 9 bytes of DRAM traffic per pixel (4 read, 4 + 1 written):
//   * the z-buffer of a map lives in a small ring of scratch slots (R maps, <= 48 MB) that stays in
//     the 126 MB L2 for the whole call: atomicMin (RED) goes to L2, depth_out / mask_out are written
//     exactly once by the finalisation, which also hands the slot back filled with 0xFFFFFFFF -- no
//     fill pass, no read-modify-write of the output;
//   * a work item SPLATS a pixel range of map k and FINALISES the same range of map k - D (R = 2 D):
//     the two halves are independent, so the finalisation's slot loads are issued before the splat
//     arithmetic and consumed after it; ONE counter per round k carries both dependencies (see RpItem).
//     Items are claimed from a ticket counter; every wait targets a smaller ticket, so with all CTAs
//     resident the smallest unfinished item can always run: no deadlock, no grid-wide barrier, no
//     launch gaps between maps;
//   * lanes own CONSECUTIVE pixels: neighbouring pixels land on neighbouring targets, so a warp's 32
//     RED operations fall into a few 32-byte sectors (measured 3.6 RED/clk/SM against 1.4 when a lane
//     owns four consecutive pixels -- tools/microbench/zbuf_atomics.cu);
//   * the arithmetic of two pixels runs on the packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2 round each
//     lane to nearest like the scalar instructions: bit-identical), rounding + bounds test is one
//     F2I.RN (round-half-even, saturating) and one unsigned compare per coordinate; the RED itself is
//     unconditional (see rp_red_min), so a sub-block's arithmetic is branch-free.
// Launch shape (tuning builds with other values: tools/build_variant.sh).  Measured on 500 maps of
// 640x480 (profiles/r2_reproject_ncu.txt): the kernel is bound by per-warp latency chains, not by
// occupancy -- 15 fat worker warps per SM (122 registers: both load batches and the finalisation's loads
// stay in registers) beat 27 thin ones (72 registers, spills).
#ifndef PRG_RP_THREADS
#define PRG_RP_THREADS 480           // worker threads per CTA (+ one helper warp = 512)
#endif
#ifndef PRG_RP_CTAS_PER_SM
#define PRG_RP_CTAS_PER_SM 1
#endif
#ifndef PRG_RP_FENCE
#define PRG_RP_FENCE 1               // scheduling fence between the pixel pairs of a sub-block (see rp_splat_sub)
#endif
constexpr int kRpThreads = PRG_RP_THREADS;
constexpr int kRpPer = 8;                       // pixels per thread and sub-block (one batch of loads in flight)
constexpr int kRpSubPx = kRpPer * kRpThreads;   // an item is walked in sub-blocks; the loads of the next one are
                                                // issued before the arithmetic of the current one
constexpr int kRpItemPx = 4 * kRpSubPx;         // pixels per work item: the resident CTAs together hold grid x item
                                                // pixels in flight, which must stay well below the D maps that
                                                // separate dependent items
constexpr float kPoseMax = 1e4f;

struct RpMap {
  Intr k;
  float P[12];
  bool fast;      // k.fast, |pose entries| <= 1e4 and the clip inside [0, 1e9]: the fast path cannot overflow
};

__device__ __forceinline__ RpMap rp_load_map(const float* __restrict__ K, const float* __restrict__ pose, int b,
                                             float lo, float hi) {
  RpMap m;
  m.k = load_intr(K, b);
  bool ok = m.k.fast && lo >= 0.f && hi <= 1e9f;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    m.P[i] = __ldg(pose + b * 16 + i);
    ok = ok && (fabsf(m.P[i]) <= kPoseMax);     // false for NaN / inf
  }
  m.fast = ok;
  return m;
}

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

// two correctly rounded quotients a / b given y = RN(1 / b) and nb = -b (see div_exact)
__device__ __forceinline__ float2 div_exact2(float2 a, float2 nb, float2 y) {
  const float2 q0 = __fmul2_rn(a, y);
  return __ffma2_rn(__ffma2_rn(q0, nb, a), y, q0);
}
// RN(1 / z) for a normal z with a normal reciprocal (no zero / denormal / inf / NaN handling): the fast
// path of __frcp_rn without its range checks -- MUFU.RCP and one Newton step on the FMA pipe.  Checked
// against __frcp_rn for every float in [2^-100, 2^100] by prg_test_frcp_exhaustive (tests/test_geometry_gpu.py).
__device__ __forceinline__ float frcp_rn_normal(float z) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  return __fmaf_rn(r, __fmaf_rn(-z, r, 1.f), r);
#else
  return __frcp_rn(z);
#endif
}

// One pixel through the validated scalar helpers (any intrinsics / pose / depth).
template <bool kScalarBmm>
__device__ __forceinline__ void rp_pixel_generic(float z0, int r, int c, const RpMap& m, int H, int W,
                                                 unsigned* __restrict__ zslot) {
  float x, y, z = z0;
  unproject(r, c, z, m.k, x, y);
  rigid<kScalarBmm>(m.P, x, y, z);
  splat(x, y, z, m.k, H, W, zslot);
}

// the same, out of line: the rare pixels of a fast map that leave the validated ranges
__device__ __noinline__ void rp_pixel_slow(float z0, int r, int c, const RpMap& m, int H, int W,
                                           unsigned* __restrict__ zslot) {
  rp_pixel_generic<false>(z0, r, c, m, H, W, zslot);
}

// z-buffer update of one pixel WITHOUT a branch or a predicate: a pixel that must not be drawn sends
// the identity of min (0xFFFFFFFF) to its own source position (in range, lane-consecutive) instead.  The
// unrolled arithmetic of a sub-block is then ONE basic block, which lets ptxas interleave the dependency
// chains of its pixel pairs (with a branch around every RED the chains ran one after another: 2.5 of 12
// cycles per issued instruction were fixed-latency waits, 1.4 branch resolution).  RED.MIN on the GLOBAL
// window: the generic-address form the compiler picks when it loses track of the address space resolves
// the window per lane first.
__device__ __forceinline__ void rp_red_min(unsigned* __restrict__ zslot, unsigned long long zglobal, int ri, int ci,
                                           int H, int W, bool ok, unsigned zbits, int own_index) {
  ok = ok && (unsigned)ci < (unsigned)W && (unsigned)ri < (unsigned)H;
#ifdef __CUDA_ARCH__
  (void)zslot;
  const unsigned idx = ok ? (unsigned)(ri * W + ci) : (unsigned)own_index;
  const unsigned val = ok ? zbits : kEmpty;
  asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, 4, %0;\n\tred.relaxed.gpu.global.min.u32 [a], %2;\n\t}\n"
               ::"l"(zglobal), "r"(idx), "r"(val));
#else
  (void)own_index; (void)zglobal;
  if (ok) atomicMin(zslot + (unsigned)(ri * W + ci), zbits);
#endif
}

// per-thread state of the fast path: map constants broadcast into both lanes of packed registers, and
// the (row, column) of the thread's current pixel PAIR (A = i, B = i + kRpThreads) carried as floats
// (exact below 2^24).  2 * kRpThreads pixels further = dr2 rows and dc2 < W columns, so ONE conditional
// wrap per step is enough for any width.
struct RpFast {
  float2 ncx, ncy, rfx, rfy, nfx, nfy, fx2, fy2, cx2, cy2;
  float2 nW2, dc2, dr2;
  float Wf;
  float2 c2, r2;
};

struct RpSteps {          // kRpThreads and 2 * kRpThreads pixels as (rows, columns < W) of the map
  int d1r, d1c, d2r, d2c;
};
__host__ __device__ inline RpSteps rp_steps(int W) {
  RpSteps st;
  st.d1r = kRpThreads / W; st.d1c = kRpThreads - st.d1r * W;
  st.d2r = (2 * kRpThreads) / W; st.d2c = 2 * kRpThreads - st.d2r * W;
  return st;
}

__device__ __forceinline__ void rp_fast_init(RpFast& f, const Intr& k, int W, const RpSteps& st, int i0) {
  f.ncx = f2(-k.cx); f.ncy = f2(-k.cy); f.rfx = f2(k.rfx); f.rfy = f2(k.rfy); f.nfx = f2(-k.fx); f.nfy = f2(-k.fy);
  f.fx2 = f2(k.fx); f.fy2 = f2(k.fy); f.cx2 = f2(k.cx); f.cy2 = f2(k.cy);
  f.Wf = (float)W;
  f.nW2 = f2(-f.Wf); f.dc2 = f2((float)st.d2c); f.dr2 = f2((float)st.d2r);
  const int r0 = i0 / W, c0 = i0 - r0 * W;
  int cb = c0 + st.d1c, rb = r0 + st.d1r;
  if (cb >= W) { cb -= W; ++rb; }
  f.c2 = make_float2((float)c0, (float)cb);
  f.r2 = make_float2((float)r0, (float)rb);
}

__device__ __forceinline__ void rp_fast_step(RpFast& f) {      // next pair: 2 * kRpThreads pixels further
  f.c2 = __fadd2_rn(f.c2, f.dc2);
  f.r2 = __fadd2_rn(f.r2, f.dr2);
  const float2 wrap = make_float2(f.c2.x >= f.Wf ? 1.f : 0.f, f.c2.y >= f.Wf ? 1.f : 0.f);
  f.c2 = __ffma2_rn(wrap, f.nW2, f.c2);
  f.r2 = __fadd2_rn(f.r2, wrap);
}

// One full sub-block (pixels base + q * kRpThreads, q < kRpPer, all inside the map) of the fast path.
__device__ __forceinline__ void rp_splat_sub(const float (&dv)[kRpPer], const float* __restrict__ dimg,
                                             unsigned* __restrict__ zslot, unsigned long long zglobal, int base,
                                             int H, int W, float lo, float hi, const RpMap& m, RpFast& f) {
#ifdef __CUDA_ARCH__
  {  // nothing inside the clip in the whole warp (background): only advance the position
    bool any = false;
#pragma unroll
    for (int q = 0; q < kRpPer; ++q) any = any || (dv[q] > lo && dv[q] < hi);
    if (!__any_sync(__activemask(), any)) {
#pragma unroll
      for (int j = 0; j < kRpPer / 2; ++j) rp_fast_step(f);
      return;
    }
  }
#endif
  bool slow = false;
#pragma unroll
  for (int j = 0; j < kRpPer / 2; ++j) {
#if defined(__CUDA_ARCH__) && PRG_RP_FENCE
    // a scheduling fence every two pairs: with the whole sub-block as one region ptxas interleaved all
    // eight pixels, ran out of predicate registers and kept them as bits of a general register (a third
    // of the instructions were that bookkeeping); two pairs are four independent chains already
    if (j > 0 && (j & 1) == 0) __syncwarp();
#endif
    const float da = dv[2 * j], db = dv[2 * j + 1];
    // valid <=> inside the clip; the clip is inside [0, 1e9], so d > 1e-9 completes depth_in_range
    const bool va = da > lo && da < hi, vb = db > lo && db < hi;
    const float2 d = make_float2(da, db);
    // x = ((c - cx) * z) / fx, y = ((r - cy) * z) / fy   (SDD:196-197).  No sign fix-up as in
    // div_exact_signed: with d > 1e-9 a numerator is zero only as the exact +0 of c - cx.
    const float2 x = div_exact2(__fmul2_rn(__fadd2_rn(f.c2, f.ncx), d), f.nfx, f.rfx);
    const float2 y = div_exact2(__fmul2_rn(__fadd2_rn(f.r2, f.ncy), d), f.nfy, f.rfy);
    // matmul(pc, R^T) + t  (SDD:279): fma(z, r2, fma(y, r1, x * r0)) + t
    const float2 X = __fadd2_rn(__ffma2_rn(d, f2(m.P[2]), __ffma2_rn(y, f2(m.P[1]), __fmul2_rn(x, f2(m.P[0])))), f2(m.P[3]));
    const float2 Y = __fadd2_rn(__ffma2_rn(d, f2(m.P[6]), __ffma2_rn(y, f2(m.P[5]), __fmul2_rn(x, f2(m.P[4])))), f2(m.P[7]));
    const float2 Z = __fadd2_rn(__ffma2_rn(d, f2(m.P[10]), __ffma2_rn(y, f2(m.P[9]), __fmul2_rn(x, f2(m.P[8])))), f2(m.P[11]));
    // c' = round(((x * fx) / z) + cx)  (SDD:225-226).  For z in (1e-6, 1e12) nothing over- or
    // underflows (|x * fx / z| < 1e33 with the bounds of RpMap::fast): no NaN can reach the F2I, whose
    // round-half-even + saturation then IS torch.round + the bounds test on the integer index.
    const bool sa = va && da > 1e-9f && Z.x > 1e-6f && Z.x < 1e12f;
    const bool sb = vb && db > 1e-9f && Z.y > 1e-6f && Z.y < 1e12f;
    // (the reciprocal of a depth outside that range may be anything: such a pixel draws nothing below)
    const float2 rz = make_float2(frcp_rn_normal(Z.x), frcp_rn_normal(Z.y));
    const float2 nZ = make_float2(-Z.x, -Z.y);
    const float2 u = __fadd2_rn(div_exact2(__fmul2_rn(X, f.fx2), nZ, rz), f.cx2);
    const float2 v = __fadd2_rn(div_exact2(__fmul2_rn(Y, f.fy2), nZ, rz), f.cy2);
    rp_red_min(zslot, zglobal, __float2int_rn(v.x), __float2int_rn(u.x), H, W, sa, __float_as_uint(Z.x), base + 2 * j * kRpThreads);
    rp_red_min(zslot, zglobal, __float2int_rn(v.y), __float2int_rn(u.y), H, W, sb, __float_as_uint(Z.y), base + (2 * j + 1) * kRpThreads);
    // valid pixels outside the validated ranges (tiny depths, extreme depth after the transform) are
    // only noted here (rare)
    slow = slow || (va && !sa) || (vb && !sb);
    rp_fast_step(f);
  }
  if (slow) {
    // ... and then ALL valid pixels of the thread's sub-block take the scalar path: it is bit-identical
    // to the fast one where both apply, and min is idempotent, so drawing a pixel twice changes nothing
    for (int q = 0; q < kRpPer; ++q) {
      const int i = base + q * kRpThreads;
      const float z0 = __ldcs(dimg + i);
      if (z0 > lo && z0 < hi) rp_pixel_slow(z0, i / W, i % W, m, H, W, zslot);
    }
  }
}

__device__ __forceinline__ void rp_load_sub(float (&dv)[kRpPer], const float* __restrict__ dimg, int base) {
#pragma unroll
  for (int q = 0; q < kRpPer; ++q) dv[q] = __ldcs(dimg + base + q * kRpThreads);
}

// Pixels [c0, e) of one map, one at a time through the scalar helpers (any intrinsics / pose, tiny maps,
// the partial chunk at the end of a map whose size is not a multiple of kRpChunkPx).  Thread t takes
// pixels c0 + t + kRpThreads j.
template <bool kScalarBmm>
__device__ __forceinline__ void rp_splat_range(const float* __restrict__ dimg, unsigned* __restrict__ zslot,
                                               int c0, int e, int H, int W, float lo, float hi, const RpMap& m) {
  int i = c0 + (int)threadIdx.x;
  if (i >= e) return;
  int r = i / W, c = i - r * W;
  for (; i < e; i += kRpThreads) {
    const float z0 = __ldcs(dimg + i);
    if (z0 > lo && z0 < hi) rp_pixel_generic<kScalarBmm>(z0, r, c, m, H, W, zslot);
    c += kRpThreads;
    while (c >= W) { c -= W; ++r; }
  }
}

// the same out of line, for the rest of a fast map's last chunk
__device__ __noinline__ void rp_splat_rest(const float* __restrict__ dimg, unsigned* __restrict__ zslot, int c0, int e,
                                           int H, int W, float lo, float hi, const RpMap& m) {
  rp_splat_range<false>(dimg, zslot, c0, e, H, W, lo, hi, m);
}

// Finalisation of pixels [c0, min(c0 + kRpChunkPx, end)) of one map, in two halves: the loads of the
// z-buffer slot (L2) are issued BEFORE the splat arithmetic of the same work item and consumed after it,
// so their latency is covered by the thread itself rather than by occupancy.  Depth (0 where empty) and
// mask out, slot reset to empty.
constexpr int kRpChunkPx = 2 * kRpSubPx;                     // 16 pixels per thread
constexpr int kRpFinIt = kRpChunkPx / (kRpThreads * 4);      // 128-bit accesses per thread and chunk

__device__ __forceinline__ void rp_finalize_load(uint4 (&vv)[kRpFinIt], const unsigned* __restrict__ zslot, int c0,
                                                 int end, int HW) {
  if ((HW & 3) != 0) return;
#pragma unroll
  for (int q = 0; q < kRpFinIt; ++q) {
    const int i = c0 + ((int)threadIdx.x + q * kRpThreads) * 4;
    if (i < end) vv[q] = __ldcg(reinterpret_cast<const uint4*>(zslot + i));
  }
}

__device__ __forceinline__ void rp_finalize_store(const uint4 (&vv)[kRpFinIt], unsigned* __restrict__ zslot,
                                                  float* __restrict__ dout, uint8_t* __restrict__ mout, int c0,
                                                  int end, int HW) {
  if ((HW & 3) == 0) {
#pragma unroll
    for (int q = 0; q < kRpFinIt; ++q) {
      const int i = c0 + ((int)threadIdx.x + q * kRpThreads) * 4;
      if (i >= end) break;
      uint4 v = vv[q];
      uchar4 mk;
      mk.x = v.x != kEmpty; mk.y = v.y != kEmpty; mk.z = v.z != kEmpty; mk.w = v.w != kEmpty;
      v.x = mk.x ? v.x : 0u; v.y = mk.y ? v.y : 0u; v.z = mk.z ? v.z : 0u; v.w = mk.w ? v.w : 0u;
      __stcs(reinterpret_cast<uint4*>(dout + i), v);
      *reinterpret_cast<uchar4*>(mout + i) = mk;
      __stcg(reinterpret_cast<uint4*>(zslot + i), make_uint4(kEmpty, kEmpty, kEmpty, kEmpty));
    }
  } else {
    const int e = min(c0 + kRpChunkPx, end);
    for (int i = c0 + (int)threadIdx.x; i < e; i += kRpThreads) {
      const unsigned v = __ldcg(zslot + i);
      mout[i] = v != kEmpty;
      reinterpret_cast<unsigned*>(dout)[i] = v != kEmpty ? v : 0u;
      zslot[i] = kEmpty;
    }
  }
}

struct RpPlan {
  int B, H, W, HW;
  RpSteps st;       // rp_steps(W)
  int item_px;      // pixels per work item (a multiple of kRpChunkPx)
  int items;        // work items per round
  int R, D;         // ring slots and finalisation lag, R = 2 D
  long long total;  // work items of the call = (B + D) * items
};

// ---- item order and synchronisation ------------------------------------------------------
// Round k (k = 0 .. B + D - 1) consists of `items` FUSED work items: item j splats pixels
// [j * item_px, (j + 1) * item_px) of map k (if k < B) and finalises the same pixel range of map k - D
// (if k >= D).  Map m lives in ring slot m % R with R = 2 D, so ONE counter per round carries both
// dependencies: done[k - D] == items says that every pixel of map k - D has been splatted (its
// finalisation may start) and that map k - 2 D has been finalised (its slot, which is map k's, is empty).
// The two halves of an item touch different slots and are independent, which is what the interleaving
// of rp_finalize_load / splat / rp_finalize_store uses.
struct RpItem {
  int round, j;
  int map_s, map_f;      // map to splat / to finalise, -1 = none
};

__device__ __forceinline__ RpItem rp_decode(long long p, const RpPlan& pl) {
  RpItem it;
  it.round = (int)(p / pl.items);
  it.j = (int)(p - (long long)it.round * pl.items);
  it.map_s = it.round < pl.B ? it.round : -1;
  it.map_f = it.round >= pl.D ? it.round - pl.D : -1;
  return it;
}

__device__ __forceinline__ int rp_peek(const int* cnt) {
  int v = 0x7fffffff;
#ifdef __CUDA_ARCH__
  if (cnt != nullptr) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
#endif
  return v;
}

// The pixel work of a fused item for the calling thread (worker threads 0 .. kRpThreads - 1); `m` is
// only read when there is a map to splat.  Thread t takes pixels px0 + t + kRpThreads j of the splat
// half and four consecutive pixels per 128-bit access of the finalisation half.
//
// Fast maps run a software pipeline over the item's chunks (a chunk = two sub-blocks of eight pixels
// per thread): the depth loads of the NEXT sub-block and the slot loads of the chunk's finalisation are
// in flight during the arithmetic of the current sub-block, so a warp covers its own memory latency
// (15 fat warps per SM do better than 27 thin ones, see the launch shape above).
template <bool kScalarBmm>
__device__ __forceinline__ void rp_work(int map_s, int map_f, int j, const RpMap& m,
                                        const float* __restrict__ depth, float lo, float hi,
                                        unsigned* __restrict__ scratch, float* __restrict__ depth_out,
                                        uint8_t* __restrict__ mask_out, const RpPlan& pl) {
  const int t = threadIdx.x;
  const int px0 = j * pl.item_px, end = min(px0 + pl.item_px, pl.HW);
  const int H = pl.H, W = pl.W, HW = pl.HW;
  const float* dimg = depth + (size_t)max(map_s, 0) * HW;
  unsigned* zs = scratch + (size_t)(max(map_s, 0) % pl.R) * HW;
  unsigned* zf = scratch + (size_t)(max(map_f, 0) % pl.R) * HW;
  float* dout = depth_out + (size_t)max(map_f, 0) * HW;
  uint8_t* mout = mask_out + (size_t)max(map_f, 0) * HW;
  int c0 = px0;
  if (!kScalarBmm && map_s >= 0 && m.fast && c0 + kRpSubPx <= end) {
    RpFast f;
    rp_fast_init(f, m.k, W, pl.st, c0 + t);
    unsigned long long zglobal = 0;
#ifdef __CUDA_ARCH__
    zglobal = __cvta_generic_to_global(zs);
    asm volatile("" : "+l"(zglobal));     // keep the slot base in one register pair (else it is re-derived per pixel)
#endif
    float dva[kRpPer], dvb[kRpPer];
    rp_load_sub(dva, dimg, c0 + t);
    for (; c0 + kRpSubPx <= end; c0 += kRpChunkPx) {
      const bool two = c0 + kRpChunkPx <= end;    // false only in the last chunk of a map: one full sub-block + a rest
      uint4 fv[kRpFinIt];
      if (map_f >= 0) rp_finalize_load(fv, zf, c0, end, HW);
      if (two) rp_load_sub(dvb, dimg, c0 + kRpSubPx + t);
      rp_splat_sub(dva, dimg, zs, zglobal, c0 + t, H, W, lo, hi, m, f);
      if (c0 + kRpChunkPx + kRpSubPx <= end) rp_load_sub(dva, dimg, c0 + kRpChunkPx + t);
      if (two) rp_splat_sub(dvb, dimg, zs, zglobal, c0 + kRpSubPx + t, H, W, lo, hi, m, f);
      else rp_splat_rest(dimg, zs, c0 + kRpSubPx, end, H, W, lo, hi, m);
      if (map_f >= 0) rp_finalize_store(fv, zf, dout, mout, c0, end, HW);
    }
  }
  for (; c0 < end; c0 += kRpChunkPx) {      // everything else (see rp_splat_range)
    uint4 fv[kRpFinIt];
    if (map_f >= 0) rp_finalize_load(fv, zf, c0, end, HW);
    if (map_s >= 0) rp_splat_range<kScalarBmm>(dimg, zs, c0, min(c0 + kRpChunkPx, end), H, W, lo, hi, m);
    if (map_f >= 0) rp_finalize_store(fv, zf, dout, mout, c0, end, HW);
  }
}

// Host emulation of the tests: the items one after another (every dependency is then already satisfied).
template <bool kScalarBmm>
__device__ __forceinline__ void rp_run_item(long long p, const float* __restrict__ depth, const float* __restrict__ K,
                                            const float* __restrict__ pose, float lo, float hi,
                                            unsigned* __restrict__ scratch, float* __restrict__ depth_out,
                                            uint8_t* __restrict__ mask_out, const RpPlan& pl) {
  const RpItem it = rp_decode(p, pl);
  RpMap m;
  if (it.map_s >= 0) m = rp_load_map(K, pose, it.map_s, lo, hi);
  rp_work<kScalarBmm>(it.map_s, it.map_f, it.j, m, depth, lo, hi, scratch, depth_out, mask_out, pl);
}

// The kernel: kRpThreads / 32 worker warps + TWO HELPER THREADS per CTA (lanes 0 and 1 of one more warp).
//  * Publishing an item ("all its REDs / stores are visible device-wide") needs a fence that waits for
//    the CTA's outstanding memory operations (~1.5 us).  With thread 0 of the workers doing it, the whole
//    CTA waited for that fence at the next barrier (4 of 5 stalled issue slots in the ncu capture of that
//    version).  Here each worker warp arrives on a shared-memory mbarrier when it is done with an item and
//    goes on; the helper sees the phase complete, fences, bumps the round's counter.
//  * The other helper lane CLAIMS the CTA's next item from a global ticket counter (one item ahead of the
//    workers), decodes it (64-bit division), loads the map's intrinsics / pose, polls the item's
//    dependency and publishes all of that through a shared-memory ring (`s_slot`, `s_ready`): the worker
//    threads find their item decoded (with every worker decoding and loading for itself, half of the
//    executed instructions of a 16-pixel-per-thread item were not pixel work), and ONE thread per CTA
//    polls.  (Polling from every warp put 8 x 443 spinning readers on a handful of counter lines in L2
//    and slowed the atomics that would have released them: 1.06-1.6 ms instead of 0.86.)
//  * Tickets are handed out in increasing order and every wait targets an item with a smaller ticket,
//    so the smallest unfinished item can always run (all CTAs are resident).
//  Four mbarriers rotate; a worker starts its k-th item only when the signaller is through the (k - 4)-th
//  (so a barrier never collects arrivals of two items), which bounds how far the workers run ahead.
constexpr int kRpCtaThreads = kRpThreads + 32;
constexpr int kRpClaimAhead = 1;       // the poller claims the CTA's (k + 1)-th item when the workers start the k-th
constexpr int kRpItemRing = 8;         // published items in shared memory (> kRing + kRpClaimAhead + 1)

struct RpSlot {
  int round;        // -1: no more work
  int j, map_s, map_f;
  RpMap m;
};

#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned rp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool rp_mbar_test(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(rp_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
#endif

template <bool kScalarBmm>
__global__ void __launch_bounds__(kRpCtaThreads, PRG_RP_CTAS_PER_SM)
k_reproject_fused(const float* __restrict__ depth, const float* __restrict__ K, const float* __restrict__ pose,
                  float lo, float hi, unsigned* __restrict__ scratch, float* __restrict__ depth_out,
                  uint8_t* __restrict__ mask_out, int* __restrict__ done, unsigned long long* __restrict__ ticket,
                  const int flags, const int claim_ahead, const RpPlan pl) {
#ifdef __CUDA_ARCH__
  constexpr int kRing = 4;                      // mbarriers in rotation = items a worker may run ahead of the signaller
  __shared__ unsigned long long s_done[kRing];  // mbarriers: the worker warps are through their k-th item (k % kRing)
  __shared__ RpSlot s_slot[kRpItemRing];        // the CTA's k-th item (k % kRpItemRing), decoded
  __shared__ volatile int s_signalled;          // items of this CTA the signaller has published
  __shared__ volatile int s_ready;              // items of this CTA that are claimed, decoded and free to run
  __shared__ volatile int s_started;            // items of this CTA the workers have started
  if (threadIdx.x == 0) {
    s_signalled = 0;
    s_ready = 0;
    s_started = 0;
    for (int b = 0; b < kRing; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rp_smem_u32(&s_done[b])), "r"(kRpThreads / 32) : "memory");
  }
  __syncthreads();
  if (threadIdx.x >= kRpThreads) {
    // ---- helper warp: lane 0 publishes finished items, lane 1 claims / decodes / polls (the two lanes
    // diverge for the whole kernel; independent thread scheduling interleaves them, so the poller's L2
    // round trips and the signaller's fences do not wait for each other)
    if (threadIdx.x == kRpThreads) {
      for (int ks = 0;; ++ks) {
        while (s_ready < ks + 1) __nanosleep(40);
        __threadfence_block();
        const int round = s_slot[ks % kRpItemRing].round;
        if (round < 0) break;
        while (!rp_mbar_test(&s_done[ks % kRing], (unsigned)(ks / kRing) & 1u)) __nanosleep(40);
        if (!(flags & 4)) __threadfence();        // the workers' REDs / stores before the counter (cumulative)
        atomicAdd(done + round, 1);
        s_signalled = ks + 1;
      }
    } else if (threadIdx.x == kRpThreads + 1) {
      for (int kp = 0;; ++kp) {
        while (s_started + (claim_ahead - 1) < kp) __nanosleep(40);   // a claimed ticket blocks its dependents: stay close
        long long p = (flags & 1) ? (long long)blockIdx.x + (long long)kp * gridDim.x
                                  : (long long)atomicAdd(ticket, 1ull);
        RpSlot& sl = s_slot[kp % kRpItemRing];
        if (p >= pl.total) {
          sl.round = -1;
        } else {
          const RpItem ip = rp_decode(p, pl);
          sl.round = ip.round; sl.j = ip.j; sl.map_s = ip.map_s; sl.map_f = ip.map_f;
          if (ip.map_s >= 0) sl.m = rp_load_map(K, pose, ip.map_s, lo, hi);
          if (ip.round >= pl.D && !(flags & 2))   // (bit 1: timing experiments only -- results are wrong)
            while (rp_peek(done + (ip.round - pl.D)) < pl.items) __nanosleep(40);
        }
        __threadfence_block();
        s_ready = kp + 1;
        if (p >= pl.total) break;
      }
    }
    return;
  }
  // ---- worker warps
  const int lane = threadIdx.x & 31;
  for (int k = 0;; ++k) {
    // published by the helper and this item's barrier free (signaller through item k - kRing)
    while (s_ready < k + 1 || s_signalled < k - (kRing - 1)) __nanosleep(20);
    __threadfence_block();
    const RpSlot& sl = s_slot[k % kRpItemRing];
    if (sl.round < 0) break;
    if (threadIdx.x == 0) s_started = k + 1;
    rp_work<kScalarBmm>(sl.map_s, sl.map_f, sl.j, sl.m, depth, lo, hi, scratch, depth_out, mask_out, pl);
    __syncwarp();
    if (lane == 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rp_smem_u32(&s_done[k % kRing])) : "memory");
  }
#endif
}

__global__ void __launch_bounds__(256)
k_zbuf_finalize(unsigned* __restrict__ zbuf, uint8_t* __restrict__ mask, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      uint4 v = __ldcg(reinterpret_cast<const uint4*>(zbuf + i));
      uchar4 m;
      m.x = v.x != kEmpty; m.y = v.y != kEmpty; m.z = v.z != kEmpty; m.w = v.w != kEmpty;
      v.x = m.x ? v.x : 0u; v.y = m.y ? v.y : 0u; v.z = m.z ? v.z : 0u; v.w = m.w ? v.w : 0u;
      *reinterpret_cast<uint4*>(zbuf + i) = v;
      *reinterpret_cast<uchar4*>(mask + i) = m;
    } else {
      for (size_t j = i; j < n; ++j) {
        unsigned v = __ldcg(zbuf + j);
        mask[j] = v != kEmpty;
        zbuf[j] = v != kEmpty ? v : 0u;
      }
    }
  }
}

// ------------------------------------------------------------------ pc2depth (ragged)
__global__ void __launch_bounds__(256)
k_pc2depth_splat(const float* __restrict__ pc, const uint8_t* __restrict__ valid,
                 const int64_t* __restrict__ offsets, int64_t total, const float* __restrict__ K,
                 const float* __restrict__ pose, unsigned* __restrict__ zbuf, int B, int H, int W) {
  extern __shared__ int64_t s_off[];
  for (int i = threadIdx.x; i <= B; i += blockDim.x) s_off[i] = offsets[i];
  __syncthreads();
  const int HW = H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (valid != nullptr && !valid[i]) continue;
    if (i < s_off[0] || i >= s_off[B]) continue;
    int lo = 0, hi = B;  // find b with off[b] <= i < off[b+1]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (s_off[mid] <= i) lo = mid; else hi = mid;
    }
    const int b = lo;
    float x = __ldcs(pc + i * 3 + 0), y = __ldcs(pc + i * 3 + 1), z = __ldcs(pc + i * 3 + 2);
    if (pose != nullptr) rigid(pose + b * 16, x, y, z);
    const Intr k = load_intr(K, b);
    splat(x, y, z, k, H, W, zbuf + (size_t)b * HW);
  }
}

// ------------------------------------------------------------------ depth2pc (dense)
// (q & ~sign) | (a & sign) in one LOP3: the sign the division gives a zero quotient (b > 0)
__device__ __forceinline__ float sign_of(float q, float a) {
#ifdef __CUDA_ARCH__
  unsigned r;
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=r"(r) : "r"(__float_as_uint(q)), "r"(__float_as_uint(a)));
  return __uint_as_float(r);
#else
  return __uint_as_float((__float_as_uint(q) & 0x7fffffffu) | (__float_as_uint(a) & 0x80000000u));
#endif
}

// Vector path (W % 4 == 0, so the four pixels of a thread share a row and every access is 16-byte
// aligned): the arithmetic of two pixels per instruction on the packed fp32x2 pipe, no row wraps.
__global__ void __launch_bounds__(256)
k_depth2pc_vec(const float* __restrict__ depth, const float* __restrict__ K, float lo, float hi,
               int use_clip, float invalid, float* __restrict__ pc, uint8_t* __restrict__ valid,
               int HW, int W) {
  const int b = blockIdx.y;
  const Intr k = load_intr(K, b);
  const float* dimg = depth + (size_t)b * HW;
  float* pimg = pc + (size_t)b * HW * 3;
  uint8_t* vimg = valid + (size_t)b * HW;
  __shared__ float4 sT[8][96];     // per-warp transpose: every store instruction writes 512 contiguous bytes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_w = 1.f / (float)W;
  const bool clip_bounds = use_clip && lo >= 0.f && hi <= 1e9f;
  const float2 ncx = make_float2(-k.cx, -k.cx), nfx = make_float2(-k.fx, -k.fx), nfy = make_float2(-k.fy, -k.fy);
  const float2 rfx = make_float2(k.rfx, k.rfx), rfy = make_float2(k.rfy, k.rfy);
  for (int base4 = blockIdx.x * blockDim.x; base4 * 4 < HW; base4 += gridDim.x * blockDim.x) {
    const int i = (base4 + (int)threadIdx.x) * 4;
    const bool in = i < HW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) v = __ldcs(reinterpret_cast<const float4*>(dimg + i));
    int r, c0;
    row_col(in ? i : 0, W, inv_w, r, c0);
    const float d[4] = {v.x, v.y, v.z, v.w};
    bool ok[4];
    float z[4];
    bool safe = k.fast;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ok[j] = use_clip ? (d[j] > lo && d[j] < hi) : true;
      z[j] = ok[j] ? d[j] : invalid;
      safe = safe && (!ok[j] || (clip_bounds ? z[j] > 1e-9f : depth_in_range(z[j])));
    }
    float o[12];
    if (safe) {
      const float c0f = (float)c0;
      const float ry = __fsub_rn((float)r, k.cy);
      const float2 ry2 = make_float2(ry, ry);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 z2 = make_float2(z[2 * h], z[2 * h + 1]);
        const float2 xn = __fmul2_rn(__fadd2_rn(make_float2(c0f + (float)(2 * h), c0f + (float)(2 * h + 1)), ncx), z2);
        const float2 yn = __fmul2_rn(ry2, z2);
        const float2 x = div_exact2(xn, nfx, rfx), y = div_exact2(yn, nfy, rfy);
        o[(2 * h) * 3 + 0] = sign_of(x.x, xn.x); o[(2 * h) * 3 + 1] = sign_of(y.x, yn.x);
        o[(2 * h + 1) * 3 + 0] = sign_of(x.y, xn.y); o[(2 * h + 1) * 3 + 1] = sign_of(y.y, yn.y);
      }
    } else {         // extreme depths or intrinsics somewhere in these four pixels: IEEE divisions
#pragma unroll
      for (int j = 0; j < 4; ++j) unproject_ieee(r, c0 + j, z[j], k, o[j * 3 + 0], o[j * 3 + 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j * 3 + 0] = ok[j] ? o[j * 3 + 0] : invalid;
      o[j * 3 + 1] = ok[j] ? o[j * 3 + 1] : invalid;
      o[j * 3 + 2] = z[j];
    }
    const int w0 = (base4 + warp * 32) * 4;           // first pixel of this warp
    if (w0 + 127 < HW) {                              // the whole warp inside the image: coalesced path
      sT[warp][lane * 3 + 0] = make_float4(o[0], o[1], o[2], o[3]);
      sT[warp][lane * 3 + 1] = make_float4(o[4], o[5], o[6], o[7]);
      sT[warp][lane * 3 + 2] = make_float4(o[8], o[9], o[10], o[11]);
      __syncwarp();
      float4* dst = reinterpret_cast<float4*>(pimg + (size_t)w0 * 3);
#pragma unroll
      for (int q = 0; q < 3; ++q) __stcs(dst + q * 32 + lane, sT[warp][q * 32 + lane]);
      __syncwarp();
      *reinterpret_cast<uchar4*>(vimg + i) = make_uchar4(ok[0], ok[1], ok[2], ok[3]);
    } else if (in) {
      float4* dst = reinterpret_cast<float4*>(pimg + (size_t)i * 3);
      __stcs(dst + 0, make_float4(o[0], o[1], o[2], o[3]));
      __stcs(dst + 1, make_float4(o[4], o[5], o[6], o[7]));
      __stcs(dst + 2, make_float4(o[8], o[9], o[10], o[11]));
      *reinterpret_cast<uchar4*>(vimg + i) = make_uchar4(ok[0], ok[1], ok[2], ok[3]);
    }
  }
}

// Generic path (any width; a thread's four pixels may span rows).
__global__ void __launch_bounds__(256)
k_depth2pc(const float* __restrict__ depth, const float* __restrict__ K, float lo, float hi,
           int use_clip, float invalid, float* __restrict__ pc, uint8_t* __restrict__ valid,
           int HW, int W) {
  const int b = blockIdx.y;
  const Intr k = load_intr(K, b);
  const float* dimg = depth + (size_t)b * HW;
  float* pimg = pc + (size_t)b * HW * 3;
  uint8_t* vimg = valid + (size_t)b * HW;
  // per-warp transpose buffer: a lane produces 48 contiguous bytes (4 points); staged so that
  // every store instruction of the warp writes 512 contiguous bytes
  __shared__ float4 sT[8][96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_w = 1.f / (float)W;
  const bool clip_bounds = use_clip && lo >= 0.f && hi <= 1e9f;
  for (int base4 = blockIdx.x * blockDim.x; base4 * 4 < HW; base4 += gridDim.x * blockDim.x) {
    const int i4 = base4 + threadIdx.x;
    const int i = i4 * 4;
    float d[4];
    const bool full = (i + 3 < HW) && (HW & 3) == 0;  // image bases stay 16-byte aligned
    if (full) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(dimg + i));
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = (i + j < HW) ? dimg[i + j] : 0.f;
    }
    float o[12];
    uint8_t ok[4];
    int r0, c0;
    row_col(i < HW ? i : 0, W, inv_w, r0, c0);
    bool safe = k.fast;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = r0, c = c0 + j;
      if (c >= W) { c -= W; ++r; while (c >= W) { c -= W; ++r; } }     // a second wrap only when W < 4
      ok[j] = use_clip ? (d[j] > lo && d[j] < hi) : 1;
      const float z = ok[j] ? d[j] : invalid;
      // a clip inside [0, 1e9] already bounds the valid depths from above and keeps them positive:
      // one comparison is left of depth_in_range (clip_bounds is uniform over the launch)
      safe = safe && (!ok[j] || (clip_bounds ? z > 1e-9f : depth_in_range(z)));
      float x, y;
      unproject_fast(r, c, z, k, x, y);
      o[j * 3 + 0] = ok[j] ? x : invalid;
      o[j * 3 + 1] = ok[j] ? y : invalid;
      o[j * 3 + 2] = z;
    }
    if (!safe) {   // extreme depths or intrinsics somewhere in these four pixels: IEEE divisions
      for (int j = 0; j < 4; ++j) {
        int r = r0, c = c0 + j;
        if (c >= W) { c -= W; ++r; while (c >= W) { c -= W; ++r; } }   // a second wrap only when W < 4
        if (ok[j]) unproject_ieee(r, c, o[j * 3 + 2], k, o[j * 3 + 0], o[j * 3 + 1]);
      }
    }
    // the whole warp inside the image and 16-byte aligned: coalesced path
    const int w0 = (base4 + warp * 32) * 4;           // first pixel of this warp
    const bool warp_full = (w0 + 127 < HW) && (HW & 3) == 0;
    if (warp_full) {
      sT[warp][lane * 3 + 0] = make_float4(o[0], o[1], o[2], o[3]);
      sT[warp][lane * 3 + 1] = make_float4(o[4], o[5], o[6], o[7]);
      sT[warp][lane * 3 + 2] = make_float4(o[8], o[9], o[10], o[11]);
      __syncwarp();
      float4* dst = reinterpret_cast<float4*>(pimg + (size_t)w0 * 3);
#pragma unroll
      for (int q = 0; q < 3; ++q) __stcs(dst + q * 32 + lane, sT[warp][q * 32 + lane]);
      __syncwarp();
      *reinterpret_cast<uchar4*>(vimg + i) = make_uchar4(ok[0], ok[1], ok[2], ok[3]);
    } else if (full) {
      float4* dst = reinterpret_cast<float4*>(pimg + (size_t)i * 3);
      __stcs(dst + 0, make_float4(o[0], o[1], o[2], o[3]));
      __stcs(dst + 1, make_float4(o[4], o[5], o[6], o[7]));
      __stcs(dst + 2, make_float4(o[8], o[9], o[10], o[11]));
      *reinterpret_cast<uchar4*>(vimg + i) = make_uchar4(ok[0], ok[1], ok[2], ok[3]);
    } else {
      for (int j = 0; j < 4 && i + j < HW; ++j) {
        pimg[(size_t)(i + j) * 3 + 0] = o[j * 3 + 0];
        pimg[(size_t)(i + j) * 3 + 1] = o[j * 3 + 1];
        pimg[(size_t)(i + j) * 3 + 2] = o[j * 3 + 2];
        vimg[i + j] = ok[j];
      }
    }
  }
}

// ------------------------------------------------------------------ occlusion_filter (SDD:446-463)
// out = (depth - minN < 0.0375f) ? depth : minN, minN = 3x3 minimum over the valid pixels (window
// clipped at the border).  A thread owns 4 adjacent columns and walks kOccRows rows keeping the
// horizontal 3-minima of the previous / current / next row in registers, so a row is read once per
// thread plus two edge scalars that hit L1.  9 B/pixel: 4 B depth + 1 B mask in, 4 B out.
constexpr int kOccRows = 8;
constexpr float kOccThreshold = 0.0375f;

struct OccRow {
  float z[4];   // raw depth of the thread's 4 pixels
  float h[4];   // min over columns c-1, c, c+1 of (mask ? depth : +inf)
};

__device__ __forceinline__ OccRow occ_load_row(const float* __restrict__ d, const uint8_t* __restrict__ m,
                                               int r, int c0, int H, int W, bool vec) {
  OccRow o;
  const float inf = __int_as_float(0x7f800000);
  if (r < 0 || r >= H) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { o.z[j] = 0.f; o.h[j] = inf; }
    return o;
  }
  const float* dr = d + (size_t)r * W;
  const uint8_t* mr = m + (size_t)r * W;
  float p[6];
  if (vec) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(dr + c0));
    const uchar4 k = __ldg(reinterpret_cast<const uchar4*>(mr + c0));
    o.z[0] = v.x; o.z[1] = v.y; o.z[2] = v.z; o.z[3] = v.w;
    p[1] = k.x ? v.x : inf; p[2] = k.y ? v.y : inf; p[3] = k.z ? v.z : inf; p[4] = k.w ? v.w : inf;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = c0 + j < W;
      o.z[j] = in ? __ldg(dr + c0 + j) : 0.f;
      p[1 + j] = (in && __ldg(mr + c0 + j)) ? o.z[j] : inf;
    }
  }
  p[0] = (c0 > 0 && __ldg(mr + c0 - 1)) ? __ldg(dr + c0 - 1) : inf;
  p[5] = (c0 + 4 < W && __ldg(mr + c0 + 4)) ? __ldg(dr + c0 + 4) : inf;
#pragma unroll
  for (int j = 0; j < 4; ++j) o.h[j] = fminf(fminf(p[j], p[j + 1]), p[j + 2]);
  return o;
}

__global__ void __launch_bounds__(256)
k_occlusion_filter(const float* __restrict__ depth, const uint8_t* __restrict__ mask,
                   float* __restrict__ out, int H, int W) {
  const int b = blockIdx.z;
  const int c0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int r0 = (blockIdx.y * 4 + (threadIdx.x >> 6)) * kOccRows;
  if (c0 >= W || r0 >= H) return;
  const float* d = depth + (size_t)b * H * W;
  const uint8_t* m = mask + (size_t)b * H * W;
  float* o = out + (size_t)b * H * W;
  const bool vec = (W & 3) == 0;          // then c0 + 3 < W and every row start is 16-byte aligned
  OccRow prev = occ_load_row(d, m, r0 - 1, c0, H, W, vec);
  OccRow cur = occ_load_row(d, m, r0, c0, H, W, vec);
#pragma unroll 1
  for (int r = r0; r < r0 + kOccRows && r < H; ++r) {
    const OccRow next = occ_load_row(d, m, r + 1, c0, H, W, vec);
    float res[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mn = fminf(fminf(prev.h[j], cur.h[j]), next.h[j]);
      res[j] = (__fsub_rn(cur.z[j], mn) < kOccThreshold) ? cur.z[j] : mn;
    }
    if (vec) {
      *reinterpret_cast<float4*>(o + (size_t)r * W + c0) = make_float4(res[0], res[1], res[2], res[3]);
    } else {
      for (int j = 0; j < 4 && c0 + j < W; ++j) o[(size_t)r * W + c0 + j] = res[j];
    }
    prev = cur;
    cur = next;
  }
}

// ------------------------------------------------------------------ point_cloud (f64, compacted)
constexpr int kCompactBlock = 1024;

__device__ __forceinline__ bool pc_valid(float d01, float scale, float lo, float hi, float& zf) {
  zf = __fmul_rn(d01, scale);
  return zf > lo && zf < hi;
}

__global__ void __launch_bounds__(kCompactBlock)
k_compact_count(const float* __restrict__ depth01, float scale, float lo, float hi,
                int64_t* __restrict__ scratch, int HW, int nblk) {
  const int b = blockIdx.y, i = blockIdx.x * kCompactBlock + threadIdx.x;
  float zf;
  const bool ok = (i < HW) && pc_valid(depth01[(size_t)b * HW + i], scale, lo, hi, zf);
  const int n = __syncthreads_count(ok);
  if (threadIdx.x == 0) scratch[(size_t)b * (nblk + 1) + blockIdx.x] = n;
}

__global__ void __launch_bounds__(1024)
k_compact_scan(int64_t* __restrict__ scratch, int64_t* __restrict__ counts, int nblk) {
  // one CTA per image: exclusive scan of the per-block counts (in place).
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  int64_t* row = scratch + (size_t)blockIdx.x * (nblk + 1);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = (i < nblk) ? row[i] : 0;
    int64_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) s_warp[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t carry = s_carry;
    const int64_t incl = s + (warp > 0 ? s_warp[warp - 1] : 0) + carry;
    if (i < nblk) row[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    row[nblk] = s_carry;
    counts[blockIdx.x] = s_carry;
  }
}

__global__ void __launch_bounds__(kCompactBlock)
k_compact_write(const float* __restrict__ depth01, const float* __restrict__ K,
                const float* __restrict__ pose, float scale, float lo, float hi,
                const int64_t* __restrict__ scratch, double* __restrict__ pc_out, int HW, int W,
                int nblk) {
  __shared__ int s_warp[32];
  const int b = blockIdx.y, i = blockIdx.x * kCompactBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float zf = 0.f;
  const bool ok = (i < HW) && pc_valid(depth01[(size_t)b * HW + i], scale, lo, hi, zf);
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane], s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    s_warp[lane] = s - w;  // exclusive
  }
  __syncthreads();
  if (!ok) return;
  const int64_t rank = scratch[(size_t)b * (nblk + 1) + blockIdx.x] + s_warp[warp] + in_warp;
  const float* k = K + b * 9;
  const double fx = (double)k[0], fy = (double)k[4], cx = (double)k[2], cy = (double)k[5];
  const int r = i / W, c = i - r * W;
  double z = (double)zf;
  double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)c, cx), z), fx);
  double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)r, cy), z), fy);
  if (pose != nullptr) {
    const float* P = pose + b * 16;
    const double dx = __dsub_rn(x, (double)P[3]), dy = __dsub_rn(y, (double)P[7]),
                 dz = __dsub_rn(z, (double)P[11]);
    // numpy's float64 (N,3)@(3,3) accumulates with fused multiply-adds in k order
    const double nx = __fma_rn(dz, (double)P[8], __fma_rn(dy, (double)P[4], __dmul_rn(dx, (double)P[0])));
    const double ny = __fma_rn(dz, (double)P[9], __fma_rn(dy, (double)P[5], __dmul_rn(dx, (double)P[1])));
    const double nz = __fma_rn(dz, (double)P[10], __fma_rn(dy, (double)P[6], __dmul_rn(dx, (double)P[2])));
    x = nx;
    y = ny;
    z = nz;
  }
  double* o = pc_out + ((size_t)b * HW + rank) * 3;
  o[0] = x;
  o[1] = y;
  o[2] = z;
}

// test hook: counts the floats in [lo_bits, hi_bits] (bit patterns, both signs) whose frcp_rn_normal differs from __frcp_rn
__global__ void __launch_bounds__(256)
k_frcp_check(unsigned lo_bits, unsigned hi_bits, unsigned long long* __restrict__ mismatches) {
  unsigned long long bad = 0;
  for (unsigned long long b = lo_bits + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= hi_bits;
       b += (unsigned long long)gridDim.x * blockDim.x) {
    const float z = __uint_as_float((unsigned)b);
    bad += __float_as_uint(frcp_rn_normal(z)) != __float_as_uint(__frcp_rn(z));
    bad += __float_as_uint(frcp_rn_normal(-z)) != __float_as_uint(__frcp_rn(-z));
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace prg

using namespace prg;

extern "C" __attribute__((visibility("default"))) int prg_test_frcp_exhaustive(float lo, float hi, uint64_t* mismatches_dev,
                                                                               prg_stream_t stream) {
  PRG_CHECK_ARG(mismatches_dev && lo > 0.f && hi >= lo, "bad range");
  PtrDeviceGuard guard(mismatches_dev);
  unsigned lb, hb;
  memcpy(&lb, &lo, 4);
  memcpy(&hb, &hi, 4);
  PRG_CUDA_OK(cudaMemsetAsync(mismatches_dev, 0, sizeof(uint64_t), (cudaStream_t)stream));
  k_frcp_check<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(lb, hb, reinterpret_cast<unsigned long long*>(mismatches_dev));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

static int grid_for(int64_t work_items, int threads, int per_thread) {
  int64_t blocks = (work_items + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// Library-owned z-buffer ring + item counters of prg_reproject_f32, one per device.  Calls are
// serialised on the ring: a call on another stream waits (device side) for the previous one.
namespace {
struct ZScratch {
  std::mutex mu;
  unsigned* ring = nullptr;
  size_t ring_words = 0;
  int* counters = nullptr;
  size_t ncounters = 0;
  cudaEvent_t done = nullptr;
  cudaStream_t last = nullptr;
  bool used = false;
  int grid = 0;
};
ZScratch g_zscratch[64];
}  // namespace

extern "C" __attribute__((visibility("default"))) int prg_reproject_f32(const float* depth, const float* K, const float* pose,
                                 float clip_lo, float clip_hi, float* depth_out,
                                 uint8_t* mask_out, int B, int H, int W, prg_stream_t stream) {
  if (B == 0) return PRG_OK;
  PRG_CHECK_ARG(depth && K && pose && depth_out && mask_out, "null pointer");
  PtrDeviceGuard guard(depth_out);
  PRG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0;
  PRG_CUDA_OK(cudaGetDevice(&dev));
  PRG_CHECK_ARG(dev >= 0 && dev < 64, "device index");
  ZScratch& z = g_zscratch[dev];
  std::lock_guard<std::mutex> lock(z.mu);

  RpPlan pl;
  pl.B = B; pl.H = H; pl.W = W; pl.HW = H * W;
  pl.st = rp_steps(W);
  static const int item_px = getenv("PRG_RP_ITEM_PX") ? std::max(1, atoi(getenv("PRG_RP_ITEM_PX")) / kRpChunkPx) * kRpChunkPx : kRpItemPx;   // tuning
  pl.item_px = item_px;
  pl.items = (pl.HW + pl.item_px - 1) / pl.item_px;
  // ring: as many maps as fit 48 MB (L2-resident next to the streaming traffic), an even number >= 2;
  // a map is finalised D = R / 2 rounds after it was splatted
  {
    const size_t per_map = (size_t)pl.HW * sizeof(unsigned);
    static const long long ring_mb = getenv("PRG_RP_RING_MB") ? std::max(1, atoi(getenv("PRG_RP_RING_MB"))) : 48;   // tuning
    long long r = (long long)((size_t)(ring_mb << 20) / per_map);
    r = std::max(2ll, std::min(192ll, r));
    r = std::min<long long>(r, 2ll * B);
    pl.D = (int)std::max(1ll, r / 2);
    pl.R = 2 * pl.D;
  }
  pl.total = (long long)(B + pl.D) * pl.items;
  const size_t need_words = (size_t)pl.R * pl.HW;
  const size_t need_cnt = (size_t)(B + pl.D) + 2;      // [ticket (64 bit)] [one counter per round]
  if (z.ring_words < need_words || z.ncounters < need_cnt) {
    PRG_CUDA_OK(cudaDeviceSynchronize());          // nobody is using the old buffers
    if (z.ring_words < need_words) {
      if (z.ring) cudaFree(z.ring);
      z.ring = nullptr; z.ring_words = 0;
      PRG_CUDA_OK(cudaMalloc(&z.ring, need_words * sizeof(unsigned)));
      PRG_CUDA_OK(cudaMemset(z.ring, 0xFF, need_words * sizeof(unsigned)));   // empty; every call leaves it so
      z.ring_words = need_words;
    }
    if (z.ncounters < need_cnt) {
      if (z.counters) cudaFree(z.counters);
      z.counters = nullptr; z.ncounters = 0;
      PRG_CUDA_OK(cudaMalloc(&z.counters, need_cnt * sizeof(int)));
      z.ncounters = need_cnt;
    }
  }
  if (z.done == nullptr) PRG_CUDA_OK(cudaEventCreateWithFlags(&z.done, cudaEventDisableTiming));
  if (z.grid == 0) {
    int occ = 0;
    PRG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_reproject_fused<false>, kRpCtaThreads, 0));
    int occ2 = 0;
    PRG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_reproject_fused<true>, kRpCtaThreads, 0));
    z.grid = num_sms() * std::max(1, std::min(occ, occ2));   // every CTA resident: the waits rely on it
  }
  if (z.used && z.last != s) PRG_CUDA_OK(cudaStreamWaitEvent(s, z.done, 0));
  PRG_CUDA_OK(cudaMemsetAsync(z.counters, 0, need_cnt * sizeof(int), s));
  const int grid = (int)std::min<long long>(z.grid, pl.total);
  // bit 0: round-robin deal instead of tickets (A/B); bits 1, 2: timing experiments that BREAK the result
  static const int flags = getenv("PRG_RP_FLAGS") ? atoi(getenv("PRG_RP_FLAGS")) : 0;
  static const int claim_ahead = getenv("PRG_RP_AHEAD") ? std::max(1, std::min(3, atoi(getenv("PRG_RP_AHEAD")))) : kRpClaimAhead;
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(z.counters);
  int* done = z.counters + 2;
  if ((long long)pl.HW * 9 < 400)     // maps of at most 44 pixels: ATen's scalar bmm rounding (see rigid())
    k_reproject_fused<true><<<grid, kRpCtaThreads, 0, s>>>(depth, K, pose, clip_lo, clip_hi, z.ring, depth_out, mask_out,
                                                       done, ticket, flags, claim_ahead, pl);
  else
    k_reproject_fused<false><<<grid, kRpCtaThreads, 0, s>>>(depth, K, pose, clip_lo, clip_hi, z.ring, depth_out, mask_out,
                                                        done, ticket, flags, claim_ahead, pl);
  PRG_LAUNCH_CHECK();
  PRG_CUDA_OK(cudaEventRecord(z.done, s));
  z.last = s;
  z.used = true;
  return PRG_OK;
}

extern "C" __attribute__((visibility("default"))) int prg_pc2depth_f32(const float* pc, const uint8_t* valid, const int64_t* offsets,
                                int64_t total_points, const float* K, const float* pose,
                                float* depth_out, uint8_t* mask_out, int B, int H, int W,
                                prg_stream_t stream) {
  if (B == 0) return PRG_OK;
  PRG_CHECK_ARG(offsets && K && depth_out && mask_out, "null pointer");
  PtrDeviceGuard guard(depth_out);
  PRG_CHECK_ARG(pc || total_points == 0, "null point cloud");
  PRG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && total_points >= 0 && B <= 4096, "bad shape");
  if (B == 0) return PRG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)B * H * W;
  PRG_CUDA_OK(cudaMemsetAsync(depth_out, 0xFF, n * sizeof(float), s));
  if (total_points > 0) {
    k_pc2depth_splat<<<grid_for(total_points, 256, 1), 256, (B + 1) * sizeof(int64_t), s>>>(
        pc, valid, offsets, total_points, K, pose, (unsigned*)depth_out, B, H, W);
    PRG_LAUNCH_CHECK();
  }
  k_zbuf_finalize<<<grid_for((int64_t)n, 256, 4), 256, 0, s>>>((unsigned*)depth_out, mask_out, n);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

extern "C" __attribute__((visibility("default"))) int prg_depth2pc_f32(const float* depth, const float* K, float clip_lo, float clip_hi,
                                int use_clip, float invalid, float* pc, uint8_t* valid, int B,
                                int H, int W, prg_stream_t stream) {
  if (B == 0) return PRG_OK;
  PRG_CHECK_ARG(depth && K && pc && valid, "null pointer");
  PtrDeviceGuard guard(pc);
  PRG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && B <= 65535, "bad shape");
  if (B == 0) return PRG_OK;
  const int HW = H * W;
  dim3 g(grid_for(HW, 256, 4), B);
  if ((W & 3) == 0)
    k_depth2pc_vec<<<g, 256, 0, (cudaStream_t)stream>>>(depth, K, clip_lo, clip_hi, use_clip, invalid, pc,
                                                       valid, HW, W);
  else
    k_depth2pc<<<g, 256, 0, (cudaStream_t)stream>>>(depth, K, clip_lo, clip_hi, use_clip, invalid, pc,
                                                   valid, HW, W);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

extern "C" __attribute__((visibility("default"))) int prg_occlusion_filter_f32(const float* depth, const uint8_t* mask, float* depth_out,
                                        int B, int H, int W, prg_stream_t stream) {
  if (B == 0) return PRG_OK;
  PRG_CHECK_ARG(depth && mask && depth_out, "null pointer");
  PtrDeviceGuard guard(depth_out);
  PRG_CHECK_ARG(depth != depth_out, "occlusion filter cannot run in place");
  PRG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && B <= 65535, "bad shape");
  dim3 g((W + 255) / 256, (H + 4 * kOccRows - 1) / (4 * kOccRows), B);
  PRG_CHECK_ARG(g.y <= 65535, "image too tall");
  k_occlusion_filter<<<g, 256, 0, (cudaStream_t)stream>>>(depth, mask, depth_out, H, W);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

extern "C" __attribute__((visibility("default"))) int prg_depth2pc_compact_f64(const float* depth01, const float* K, const float* pose,
                                        float scale, float clip_lo, float clip_hi, double* pc_out,
                                        int64_t* counts, int64_t* scratch, int B, int H, int W,
                                        prg_stream_t stream) {
  if (B == 0) return PRG_OK;
  PRG_CHECK_ARG(depth01 && K && pc_out && counts && scratch, "null pointer");
  PtrDeviceGuard guard(pc_out);
  PRG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && B <= 65535, "bad shape");
  if (B == 0) return PRG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = H * W;
  const int nblk = (HW + kCompactBlock - 1) / kCompactBlock;
  dim3 g(nblk, B);
  k_compact_count<<<g, kCompactBlock, 0, s>>>(depth01, scale, clip_lo, clip_hi, scratch, HW, nblk);
  PRG_LAUNCH_CHECK();
  k_compact_scan<<<B, 1024, 0, s>>>(scratch, counts, nblk);
  PRG_LAUNCH_CHECK();
  k_compact_write<<<g, kCompactBlock, 0, s>>>(depth01, K, pose, scale, clip_lo, clip_hi, scratch,
                                             pc_out, HW, W, nblk);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

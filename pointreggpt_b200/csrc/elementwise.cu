// Non-GEMM kernels of the U-Net forward / sampler step.  See elementwise.cuh for semantics.
#include <algorithm>

#include "common.cuh"
#include "elementwise.cuh"
#include "conv_tc.cuh"

namespace prg {

// x * sigmoid(x) with the hardware reciprocal (<= 2 ulp; an IEEE division costs ~10 instructions and
// made the GroupNorm-apply pass issue-bound instead of HBM-bound)
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// Two SiLUs at once on the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2 issue one instruction per
// element PAIR; the 3-register scalar forms issue at half rate on this part, which made the
// GroupNorm-apply pass co-limited by instruction issue).  Same approximations as silu().
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 silu2(float2 t) {
  const float2 u = __fmul2_rn(t, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 d = __fadd2_rn(make_float2(ex2_approx(u.x), ex2_approx(u.y)), make_float2(1.f, 1.f));
  return __fmul2_rn(t, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

constexpr int kStemT = 16;                 // stem patch: 16x16 output pixels per CTA iteration
constexpr int kStemHalo = kStemT + 6;

// ------------------------------------------------------------------------------------------
// stems on the tensor pipe (mma.sync m16n8k16, fp16 operands, fp32 accumulate).
// K = 49 * CIN is thin, so precision is kept by operand splitting instead of wider types:
//   x = xh + xl, w = wh + wl (fp16 halves);  x.w ~= xh.wh + xl.wh + xh.wl   (error ~2^-22)
// i.e. one GEMM with K' = 3 * 49 * CIN columns [xh | xl | xh] against rows [wh | wh | wl].
// Persistent CTAs (8 warps) walk 16x16-pixel patches; a warp owns two 16-pixel rows (two
// M = 16 tiles) x all 64 output channels.  The direct CUDA-core version was LDS-bound
// (17 shared-memory loads per 64 FMAs, 0.50 ms per U-Net evaluation at batch 32).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void hmma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CIN, bool AUG>
__global__ void __launch_bounds__(256)
k_stem_tc(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
          __half* __restrict__ y, int S, int B) {
  pdl_trigger();
  pdl_wait();
  constexpr int KT = 49 * CIN;                 // taps
  constexpr int KR = 3 * KT;                   // real K of the split GEMM
  constexpr int KS = (KR + 15) / 16;           // k-steps
  constexpr int PW = kStemHalo;                // patch width (22)
  constexpr int PSZ = PW * PW;                 // 484 values per channel
  constexpr int SEG = CIN * PSZ;               // one segment (xh or xl)
  extern __shared__ __align__(16) uint8_t stem_smem[];
  // layout: B fragments [KS][8 n-tiles][32 lanes] uint2 | lut [KS*16] u16 | vals: xh[SEG] xl[SEG] zero[PSZ] halves
  //         | fp32 patch scratch (AUG: raw (PW+2)^2) | per-warp staging 8 x 2 KB
  uint2* sB = reinterpret_cast<uint2*>(stem_smem);
  uint16_t* lut = reinterpret_cast<uint16_t*>(sB + KS * 8 * 32);
  __half* vals = reinterpret_cast<__half*>(lut + KS * 16);
  float* sraw = reinterpret_cast<float*>(vals + 2 * SEG + PSZ + 8);
  constexpr int RAWN = AUG ? (PW + 2) * (PW + 2) : 0;
  uint8_t* stage = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(sraw + ((RAWN + 3) & ~3)) + 127) & ~static_cast<uintptr_t>(127));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- once per CTA: weight fragments and the k -> patch-offset table
  for (int i = tid; i < KS * 8 * 32; i += 256) {
    const int ln = i & 31, nt = (i >> 5) & 7, ks = i >> 8;
    const int n = nt * 8 + (ln >> 2), tt = ln & 3;
    __half hv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = ks * 16 + 2 * tt + (j & 1) + (j >> 1) * 8;
      float v = 0.f;
      if (k < KR) {
        const int seg = k / KT, tap = k - seg * KT;     // tap = cin*49 + ky*7 + kx
        const float wf = __ldg(w + (size_t)n * KT + tap);
        const __half wh = __float2half_rn(wf);
        v = (seg < 2) ? __half2float(wh) : (wf - __half2float(wh));
      }
      hv[j] = __float2half_rn(v);
    }
    uint2 o;
    o.x = (uint32_t)__half_as_ushort(hv[0]) | ((uint32_t)__half_as_ushort(hv[1]) << 16);
    o.y = (uint32_t)__half_as_ushort(hv[2]) | ((uint32_t)__half_as_ushort(hv[3]) << 16);
    sB[i] = o;
  }
  for (int k = tid; k < KS * 16; k += 256) {
    int off = 2 * SEG;                                   // padding columns read the zero block
    if (k < KR) {
      const int seg = k / KT, tap = k - seg * KT;
      const int ci = tap / 49, r = tap - ci * 49;
      off = ((seg == 1) ? SEG : 0) + ci * PSZ + (r / 7) * PW + (r % 7);
    }
    lut[k] = (uint16_t)off;
  }
  for (int i = tid; i < PSZ + 8; i += 256) vals[2 * SEG + i] = __float2half_rn(0.f);

  const int tiles_x = (S + kStemT - 1) / kStemT;
  const int patches = tiles_x * tiles_x * B;
  for (int p = blockIdx.x; p < patches; p += gridDim.x) {
    const int b = p / (tiles_x * tiles_x), r = p - b * tiles_x * tiles_x;
    const int y0 = (r / tiles_x) * kStemT, x0 = (r % tiles_x) * kStemT;
    const float* img = x + (size_t)b * S * S;
    __syncthreads();   // previous patch fully consumed (and the tables are written)
    if (!AUG) {
      for (int i = tid; i < PSZ; i += 256) {
        const int pr = i / PW, pc = i - pr * PW;
        const int gy = y0 + pr - 3, gx = x0 + pc - 3;
        const float v = (gy >= 0 && gy < S && gx >= 0 && gx < S) ? __ldg(img + (size_t)gy * S + gx) : 0.f;
        const __half h = __float2half_rn(v);
        vals[i] = h;
        vals[SEG + i] = __float2half_rn(v - __half2float(h));
      }
    } else {
      // DepthAugment (DC:582-604), see the comment in the previous revision of this kernel: ch0 =
      // depth, ch1 = 3x3 min over valid (!= 0) neighbours (out-of-image ignored), falling back to
      // the plain 3x3 min when no neighbour is valid; ch2 = ch1 - ch0; zero outside the image.
      constexpr int R = PW + 2;
      for (int i = tid; i < R * R; i += 256) {
        const int pr = i / R, pc = i - pr * R;
        const int gy = y0 + pr - 4, gx = x0 + pc - 4;
        sraw[i] = (gy >= 0 && gy < S && gx >= 0 && gx < S) ? __ldg(img + (size_t)gy * S + gx)
                                                           : __int_as_float(0x7fc00000);  // NaN = outside
      }
      __syncthreads();
      for (int i = tid; i < PSZ; i += 256) {
        const int pr = i / PW, pc = i - pr * PW;
        const float d = sraw[(pr + 1) * R + pc + 1];
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (d == d) {  // inside the image
          float mn_valid = INFINITY, mn_all = INFINITY;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float v = sraw[(pr + dy) * R + pc + dx];
              if (v == v) {
                mn_all = fminf(mn_all, v);
                if (v != 0.f) mn_valid = fminf(mn_valid, v);
              }
            }
          const float mn = isinf(mn_valid) ? mn_all : mn_valid;
          c0 = d;
          c1 = mn;
          c2 = mn - d;
        }
        const float cc[3] = {c0, c1, c2};
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const __half h = __float2half_rn(cc[ci]);
          vals[ci * PSZ + i] = h;
          vals[SEG + ci * PSZ + i] = __float2half_rn(cc[ci] - __half2float(h));
        }
      }
    }
    __syncthreads();
    // ---- two M tiles per warp: patch rows 2*warp and 2*warp + 1; A row = pixel x = g / g + 8
    float acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b0 = __ldg(bias + j * 8 + 2 * t), b1 = __ldg(bias + j * 8 + 2 * t + 1);
        acc[m][j][0] = b0; acc[m][j][1] = b1; acc[m][j][2] = b0; acc[m][j][3] = b1;
      }
    const uint16_t* vs = reinterpret_cast<const uint16_t*>(vals);
    const int base0 = (2 * warp) * PW + g;           // pixel (row 2*warp, x = g) patch origin offset
#pragma unroll 2
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t l01 = *reinterpret_cast<const uint32_t*>(lut + ks * 16 + 2 * t);       // k = 2t, 2t+1
      const uint32_t l23 = *reinterpret_cast<const uint32_t*>(lut + ks * 16 + 2 * t + 8);   // k = 2t+8, +9
      const int o0 = l01 & 0xffff, o1 = l01 >> 16, o2 = l23 & 0xffff, o3 = l23 >> 16;
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int pb = base0 + m * PW;
        a[m][0] = (uint32_t)vs[o0 + pb] | ((uint32_t)vs[o1 + pb] << 16);            // row g,   k 2t..
        a[m][1] = (uint32_t)vs[o0 + pb + 8] | ((uint32_t)vs[o1 + pb + 8] << 16);    // row g+8, k 2t..
        a[m][2] = (uint32_t)vs[o2 + pb] | ((uint32_t)vs[o3 + pb] << 16);            // row g,   k 2t+8..
        a[m][3] = (uint32_t)vs[o2 + pb + 8] | ((uint32_t)vs[o3 + pb + 8] << 16);    // row g+8, k 2t+8..
      }
      const uint2* bp = sB + (ks * 8) * 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint2 bb = bp[j * 32];
        hmma16816(acc[0][j], a[0], bb.x, bb.y);
        hmma16816(acc[1][j], a[1], bb.x, bb.y);
      }
    }
    // ---- fp16, through a per-warp staging tile so that each lane stores 16 contiguous bytes
    uint8_t* st = stage + warp * 2048;                // 16 pixels x 128 B
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half2 lo = __floats2half2_rn(acc[m][j][0], acc[m][j][1]);
        const __half2 hi = __floats2half2_rn(acc[m][j][2], acc[m][j][3]);
        // 16-byte chunk j of a pixel row, XOR-swizzled by the pixel index (conflict-free both ways)
        *reinterpret_cast<__half2*>(st + g * 128 + ((j ^ (g & 7)) << 4) + t * 4) = lo;
        *reinterpret_cast<__half2*>(st + (g + 8) * 128 + ((j ^ (g & 7)) << 4) + t * 4) = hi;
      }
      __syncwarp();
      const int gy = y0 + 2 * warp + m;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = q * 32 + lane, px = idx >> 3, ch = idx & 7;
        const int gx = x0 + px;
        if (gy < S && gx < S) {
          const uint4 v = *reinterpret_cast<const uint4*>(st + px * 128 + ((ch ^ (px & 7)) << 4));
          *reinterpret_cast<uint4*>(y + (((size_t)b * S + gy) * S + gx) * 64 + ch * 8) = v;
        }
      }
    }
  }
}

template <int CIN, bool AUG>
static int stem_tc_launch(const float* x, const float* w, const float* bias, __half* y, int B, int S,
                          cudaStream_t s) {
  constexpr int KS = (3 * 49 * CIN + 15) / 16;
  constexpr int PSZ = kStemHalo * kStemHalo;
  constexpr int RAWN = AUG ? (kStemHalo + 2) * (kStemHalo + 2) : 0;
  const size_t smem = (size_t)KS * 8 * 32 * 8 + (size_t)KS * 16 * 2 + (size_t)(2 * CIN * PSZ + PSZ + 8) * 2 +
                      (size_t)((RAWN + 3) & ~3) * 4 + 8 * 2048 + 256;
  static bool configured = false;
  if (!configured) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_stem_tc<CIN, AUG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    configured = true;
  }
  const int tiles_x = (S + kStemT - 1) / kStemT;
  const int patches = tiles_x * tiles_x * B;
  int grid = num_sms() * 2;
  if (grid > patches) grid = patches;
  PRG_CUDA_OK(launch_pdl(k_stem_tc<CIN, AUG>, grid, dim3(256), smem, s, x, w, bias, y, S, B));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

int stem_unet(const float* x, const float* w, const float* bias, __half* y, int B, int S,
              cudaStream_t s) {
  return stem_tc_launch<1, false>(x, w, bias, y, B, S, s);
}

int stem_mask(const float* depth01, const float* w, const float* bias, __half* y, int B, int S,
              cudaStream_t s) {
  return stem_tc_launch<3, true>(depth01, w, bias, y, B, S, s);
}

// ------------------------------------------------------------------------------------------
// conditioning MLPs
// ------------------------------------------------------------------------------------------
// One CTA per image: sinusoidal embedding -> Linear -> GELU -> Linear, param MLP likewise,
// SiLU of both written to cond_act (every block MLP starts with SiLU, SDD:709-710).
__global__ void __launch_bounds__(512)
k_cond_embed(CondWeights w, const int64_t* __restrict__ time, int time_scalar,
             const float* __restrict__ pcond, float* __restrict__ cond_act) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int dim = w.dim, hid = 4 * w.dim;
  float* emb = sm;            // [dim]
  float* h1 = emb + dim;      // [hid]
  float* pin = h1 + hid;      // [pdim]
  float* h2 = pin + w.pdim;   // [hid]
  const int b = blockIdx.x;
  const float t = (time != nullptr) ? (float)time[b] : (float)time_scalar;
  const int half = dim / 2;
  // SDD:650-657: freq_i = exp(i * -(ln 1e4 / (half-1)))
  const float step = logf(10000.f) / (float)(half - 1);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf((float)i * -step);
    const float a = t * f;
    emb[i] = sinf(a);
    emb[i + half] = cosf(a);
  }
  for (int i = threadIdx.x; i < w.pdim; i += blockDim.x) pin[i] = pcond[b * w.pdim + i];
  __syncthreads();
  // warp-per-row GEMVs (lanes stride K -> coalesced weight reads, fixed-order shuffle reduction)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < 2 * hid; j += nwarps) {
    const bool is_p = j >= hid;
    const int r = is_p ? j - hid : j;
    const int K = is_p ? w.pdim : dim;
    const float* wr = (is_p ? w.p1w : w.t1w) + (size_t)r * K;
    const float* in = is_p ? pin : emb;
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a = fmaf(__ldg(wr + k), in[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) (is_p ? h2 : h1)[r] = gelu_erf(a + (is_p ? w.p1b : w.t1b)[r]);
  }
  __syncthreads();
  for (int j = warp; j < 2 * hid; j += nwarps) {
    const bool is_p = j >= hid;
    const int r = is_p ? j - hid : j;
    const float* wr = (is_p ? w.p2w : w.t2w) + (size_t)r * hid;
    const float* in = is_p ? h2 : h1;
    float a = 0.f;
    for (int k = lane; k < hid; k += 32) a = fmaf(__ldg(wr + k), in[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) cond_act[(size_t)b * 2 * hid + j] = silu(a + (is_p ? w.p2b : w.t2b)[r]);
  }
}

int cond_embed(const CondWeights& w, const int64_t* time, int time_scalar, const float* pcond,
               float* cond_act, int B, cudaStream_t s) {
  const size_t smem = sizeof(float) * (w.dim + 8 * w.dim + w.pdim);
  PRG_CUDA_OK(launch_pdl(k_cond_embed, dim3(B), dim3(512), smem, s, w, time, time_scalar, pcond, cond_act));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// ---- sampler fast path: every image of the batch shares the timestep and param_cond is constant
// over the steps, and Linear(SiLU(cat(t, p))) = W_t SiLU(t) + W_p SiLU(p) + b is separable:
//   once per sample():  act_t[i] = SiLU(time_mlp(t_i)) for ALL steps i,  ss_p[b] = W_p SiLU(param_mlp(p_b))
//   per step:           ss[b] = W_t act_t[i] + bias + ss_p[b]            (one small launch)
__global__ void __launch_bounds__(256)
k_cond_time_all(CondWeights w, const int* __restrict__ ts, float* __restrict__ act_t) {
  extern __shared__ float sm[];
  const int dim = w.dim, hid = 4 * w.dim;
  float* emb = sm;        // [dim]
  float* h1 = emb + dim;  // [hid]
  const float t = (float)ts[blockIdx.x];
  const int half = dim / 2;
  const float step = logf(10000.f) / (float)(half - 1);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf((float)i * -step);
    const float a = t * f;
    emb[i] = sinf(a);
    emb[i + half] = cosf(a);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < hid; r += nwarps) {
    const float* wr = w.t1w + (size_t)r * dim;
    float a = 0.f;
    for (int k = lane; k < dim; k += 32) a = fmaf(__ldg(wr + k), emb[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) h1[r] = gelu_erf(a + w.t1b[r]);
  }
  __syncthreads();
  for (int r = warp; r < hid; r += nwarps) {
    const float* wr = w.t2w + (size_t)r * hid;
    float a = 0.f;
    for (int k = lane; k < hid; k += 32) a = fmaf(__ldg(wr + k), h1[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) act_t[(size_t)blockIdx.x * hid + r] = silu(a + w.t2b[r]);
  }
}

// ss_p[b][r] = W[r][k0 : k0 + K] . act[b][k0 : k0 + K]   (no bias): the param_cond half, once per sample()
__global__ void __launch_bounds__(256)
k_cond_mlp_part(const float* __restrict__ W, const float* __restrict__ act, float* __restrict__ out, int rows,
                int Ktot, int k0, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= rows) return;
  const float* wr = W + (size_t)warp * Ktot + k0;
  const float* c = act + (size_t)b * Ktot + k0;
  float a = 0.f;
  for (int k = lane; k < K; k += 32) a = fmaf(__ldg(wr + k), c[k], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[(size_t)b * rows + warp] = a;
}

// per step: one warp per output row computes the time half once and adds every image's param half
__global__ void __launch_bounds__(256)
k_cond_mlp_step(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ act_t,
                const float* __restrict__ ss_p, float* __restrict__ ss, int rows, int Ktot, int K, int B,
                const int* __restrict__ step_idx, int act_stride) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  if (step_idx != nullptr) act_t += (size_t)(*step_idx) * act_stride;   // this step's time embedding
  const float* wr = W + (size_t)warp * Ktot;
  float a = 0.f;
  for (int k = lane; k < K; k += 32) a = fmaf(__ldg(wr + k), act_t[k], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  a += __ldg(bias + warp);
  for (int b = lane; b < B; b += 32) ss[(size_t)b * rows + warp] = a + ss_p[(size_t)b * rows + warp];
}

int cond_time_all(const CondWeights& w, const int* ts_dev, int nsteps, float* act_t, cudaStream_t s) {
  k_cond_time_all<<<nsteps, 256, sizeof(float) * 5 * w.dim, s>>>(w, ts_dev, act_t);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}
int cond_mlp_param(const float* W, const float* cond_act, float* ss_p, int rows, int Ktot, int B,
                   cudaStream_t s) {
  dim3 g((rows * 32 + 255) / 256, B);
  k_cond_mlp_part<<<g, 256, 0, s>>>(W, cond_act, ss_p, rows, Ktot, Ktot / 2, Ktot / 2);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}
int cond_mlp_step(const float* W, const float* bias, const float* act_t, const float* ss_p, float* ss,
                  int rows, int Ktot, int B, cudaStream_t s, const int* step_idx, int act_stride) {
  PRG_CUDA_OK(launch_pdl(k_cond_mlp_step, dim3((rows * 32 + 255) / 256), dim3(256), 0, s, W, bias, act_t, ss_p, ss, rows,
                         Ktot, Ktot / 2, B, step_idx, act_stride));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// one warp per (image, output row)
__global__ void __launch_bounds__(256)
k_cond_mlp(const float* __restrict__ W, const float* __restrict__ bias,
           const float* __restrict__ cond_act, float* __restrict__ ss, int rows, int K) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= rows) return;
  const float* wr = W + (size_t)warp * K;
  const float* c = cond_act + (size_t)b * K;
  float a = 0.f;
  for (int k = lane; k < K; k += 32) a = fmaf(__ldg(wr + k), c[k], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) ss[(size_t)b * rows + warp] = a + __ldg(bias + warp);
}

int cond_mlp(const float* W, const float* bias, const float* cond_act, float* ss, int rows, int K,
             int B, cudaStream_t s) {
  dim3 g((rows * 32 + 255) / 256, B);
  PRG_CUDA_OK(launch_pdl(k_cond_mlp, g, dim3(256), 0, s, W, bias, cond_act, ss, rows, K));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// ------------------------------------------------------------------------------------------
// GroupNorm apply (+ scale/shift, SiLU, optional residual).  One thread = 8 channels of a pixel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_coeffs(const long long* stats, const float* gamma, const float* beta,
                                          const float* ss_row, int C, int HW, int b, float* sA,
                                          float* sB) {
  // the eight group statistics need the fp64 path (exact integer sums -> mean / variance without
  // cancellation); everything per channel is fp32.  Must be called by the whole CTA (it syncs).
  __shared__ float s_mean[8], s_rstd[8];
  const int gs = C >> 3;
  if (threadIdx.x < 8) {
    const int g = threadIdx.x;
    const double inv_n = 1.0 / ((double)gs * (double)HW);
    const double sum = (double)stats[((size_t)b * 8 + g) * 2 + 0] * (double)kStatUnscale;
    const double sq = (double)stats[((size_t)b * 8 + g) * 2 + 1] * (double)kStatUnscale;
    const double meand = sum * inv_n;
    const float var = fmaxf((float)(sq * inv_n - meand * meand), 0.f);
    s_mean[g] = (float)meand;
    s_rstd[g] = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / gs;
    const float mean = s_mean[g], rstd = s_rstd[g];
    float a = rstd * gamma[c], d = beta[c] - mean * rstd * gamma[c];
    if (ss_row != nullptr) {
      const float sc = ss_row[c] + 1.f, sh = ss_row[C + c];
      a *= sc;
      d = d * sc + sh;
    }
    sA[c] = a;
    sB[c] = d;
  }
}

// (A, B) per (image, channel) of y = SiLU(A * raw + B), for epilogues that apply the GroupNorm
// on the fly (conv engine EPI_GNRES).  One CTA per image.
__global__ void __launch_bounds__(256)
k_gn_coef(const long long* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
          const float* __restrict__ ss, int ss_stride, int ss_off, int C, int HW, float2* __restrict__ coef) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* sA = sm;
  float* sB = sm + C;
  const int b = blockIdx.x;
  gn_coeffs(stats, gamma, beta, ss ? ss + (size_t)b * ss_stride + ss_off : nullptr, C, HW, b, sA, sB);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) coef[(size_t)b * C + c] = make_float2(sA[c], sB[c]);
}

int gn_coef(const GnApply& a, float2* coef, int B, cudaStream_t s) {
  PRG_CUDA_OK(launch_pdl(k_gn_coef, dim3(B), dim3(256), 2 * a.C * sizeof(float), s, a.stats, a.gamma, a.beta, a.ss,
                         a.ss_stride, a.ss_off, a.C, a.HW, coef));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

template <bool LN>
__global__ void __launch_bounds__(256, LN ? 3 : 4)
k_gn_apply(GnApply a, int total_blocks, int nblk) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* sA = sm;
  float* sB = sm + a.C;
  float* sG = sm + 2 * a.C;
  // Persistent: the (image, 16 KiB block) pairs are dealt to the CTAs in contiguous ranges, so
  // the per-image coefficient prologue (fp64 mean / variance from the integer statistics) runs
  // once or twice per CTA instead of once per 16 KiB of data.
  const int g_begin = (int)(((long long)blockIdx.x * total_blocks) / gridDim.x);
  const int g_end = (int)(((long long)(blockIdx.x + 1) * total_blocks) / gridDim.x);
  if (LN)
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) sG[c] = a.ln_g[c];
  const int cvec = a.C >> 3;                       // 16-byte vectors per pixel (power of two)
  const int nvec = a.HW * cvec;                    // per image: fits 32 bits (HW * C / 8)
  constexpr int U = 4;                             // independent 16-byte loads in flight per thread
  const float inv_c = 1.f / (float)a.C;
  int cvec_log2 = 0;
  while ((1 << cvec_log2) < cvec) ++cvec_log2;
  // 256 is a multiple of the vectors per pixel, so a thread meets the same eight channels in
  // every vector it touches: their coefficients live in registers
  const int cv = threadIdx.x & (cvec - 1);
  const int c0 = cv * 8;
  float cA[8], cB[8], cG[8];
  int b = -1;
  const uint4* src = nullptr;
  uint4* dst = nullptr;
  uint4* ldst = nullptr;
  const __half* resb = nullptr;
  for (int g = g_begin; g < g_end; ++g) {
    const int gb = g / nblk, blk = g - gb * nblk;
    if (gb != b) {
      b = gb;
      __syncthreads();   // everyone is done with the previous image's coefficients
      gn_coeffs(a.stats, a.gamma, a.beta, a.ss ? a.ss + (size_t)b * a.ss_stride + a.ss_off : nullptr,
                a.C, a.HW, b, sA, sB);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cA[j] = sA[c0 + j];
        cB[j] = sB[c0 + j];
        cG[j] = LN ? sG[c0 + j] : 0.f;
      }
      src = reinterpret_cast<const uint4*>(a.raw + (size_t)b * a.HW * a.C);
      dst = reinterpret_cast<uint4*>(a.y + (size_t)b * a.HW * a.C);
      ldst = LN ? reinterpret_cast<uint4*>(a.ln_out + (size_t)b * a.HW * a.C) : nullptr;
      resb = a.res ? a.res + (size_t)b * a.HW * a.res_pix_stride + c0 : nullptr;
    }
    const int i0 = blk * (256 * U) + (int)threadIdx.x;
    uint4 v[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 256;
      if (i < nvec) v[u] = __ldcs(src + i);
    }
    if (a.res != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        if (i < nvec)
          rv[u] = __ldg(reinterpret_cast<const uint4*>(resb + (size_t)(i >> cvec_log2) * a.res_pix_stride));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 256;
      // nvec is a multiple of the warp size (HW % 128 == 0), so a warp is all-in or all-out
      if (i >= nvec) continue;
      const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 y = silu2(__ffma2_rn(__half22float2(h[j]), make_float2(cA[2 * j], cA[2 * j + 1]),
                                    make_float2(cB[2 * j], cB[2 * j + 1])));
        if (a.res != nullptr)
          y = __fadd2_rn(y, __half22float2(reinterpret_cast<const __half2*>(&rv[u])[j]));
        f[2 * j] = y.x;
        f[2 * j + 1] = y.y;
      }
      uint4 o;
      __half2 o0 = __floats2half2_rn(f[0], f[1]), o1 = __floats2half2_rn(f[2], f[3]),
              o2 = __floats2half2_rn(f[4], f[5]), o3 = __floats2half2_rn(f[6], f[7]);
      o.x = *reinterpret_cast<uint32_t*>(&o0);
      o.y = *reinterpret_cast<uint32_t*>(&o1);
      o.z = *reinterpret_cast<uint32_t*>(&o2);
      o.w = *reinterpret_cast<uint32_t*>(&o3);
      dst[i] = o;
      if (LN) {
        // LayerNorm of the fp16-rounded values the attention would otherwise re-read
        float g8[8];
        {
          const float2 t0 = __half22float2(o0), t1 = __half22float2(o1), t2 = __half22float2(o2),
                       t3 = __half22float2(o3);
          g8[0] = t0.x; g8[1] = t0.y; g8[2] = t1.x; g8[3] = t1.y;
          g8[4] = t2.x; g8[5] = t2.y; g8[6] = t3.x; g8[7] = t3.y;
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += g8[j];
        for (int o2s = cvec >> 1; o2s > 0; o2s >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2s);
        const float mean = s * inv_c;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = g8[j] - mean;
          q = fmaf(d, d, q);
        }
        for (int o2s = cvec >> 1; o2s > 0; o2s >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o2s);
        const float rstd = rsqrtf(q * inv_c + 1e-5f);
        uint4 lo;
        __half2 l0 = __floats2half2_rn((g8[0] - mean) * rstd * cG[0], (g8[1] - mean) * rstd * cG[1]);
        __half2 l1 = __floats2half2_rn((g8[2] - mean) * rstd * cG[2], (g8[3] - mean) * rstd * cG[3]);
        __half2 l2 = __floats2half2_rn((g8[4] - mean) * rstd * cG[4], (g8[5] - mean) * rstd * cG[5]);
        __half2 l3 = __floats2half2_rn((g8[6] - mean) * rstd * cG[6], (g8[7] - mean) * rstd * cG[7]);
        lo.x = *reinterpret_cast<uint32_t*>(&l0);
        lo.y = *reinterpret_cast<uint32_t*>(&l1);
        lo.z = *reinterpret_cast<uint32_t*>(&l2);
        lo.w = *reinterpret_cast<uint32_t*>(&l3);
        ldst[i] = lo;
      }
    }
  }
}

int gn_apply(const GnApply& a, int B, cudaStream_t s) {
  const int nvec = a.HW * (a.C >> 3);
  const int nblk = (nvec + 256 * 4 - 1) / (256 * 4);     // 16 KiB blocks per image
  const int total = nblk * B;
  const bool ln = a.ln_out != nullptr;
  if (ln && (a.C > 256)) {
    set_error("gn_apply: fused LayerNorm needs C <= 256 (got %d)", a.C);
    return PRG_ERR_ARG;
  }
  int grid = num_sms() * (ln ? 3 : 4);                    // one resident wave
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  if (ln)
    PRG_CUDA_OK(launch_pdl(k_gn_apply<true>, dim3(grid), dim3(256), 3 * a.C * sizeof(float), s, a, total, nblk));
  else
    PRG_CUDA_OK(launch_pdl(k_gn_apply<false>, dim3(grid), dim3(256), 3 * a.C * sizeof(float), s, a, total, nblk));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// ------------------------------------------------------------------------------------------
// channel LayerNorm * g  (one warp per pixel, exact two-pass mean / variance in registers)
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256)
k_ln_apply(const __half* __restrict__ x, const float* __restrict__ g,
           const __half* __restrict__ res, __half* __restrict__ y, int64_t npix) {
  pdl_trigger();
  pdl_wait();
  constexpr int PER = C / 32;  // channels per lane (2, 4, 8 or 16), contiguous
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float gl[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) gl[j] = __ldg(g + lane * PER + j);
  for (int64_t p = warp0; p < npix; p += nwarps) {
    const __half* src = x + p * C + lane * PER;
    float f[PER];
    if (PER == 2) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(src));
      f[0] = t.x; f[1] = t.y;
    } else if (PER == 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(src);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      const float2 t0 = __half22float2(h[0]), t1 = __half22float2(h[1]);
      f[0] = t0.x; f[1] = t0.y; f[2 % PER] = t1.x; f[3 % PER] = t1.y;
    } else {
#pragma unroll
      for (int q = 0; q < PER / 8; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + q * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(h[j]);
          f[(q * 8 + 2 * j) % PER] = t.x;
          f[(q * 8 + 2 * j + 1) % PER] = t.y;
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) s += f[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const float d = f[j] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + 1e-5f);
    __half* dst = y + p * C + lane * PER;
    if (PER >= 8) {
      // 16-byte residual loads / stores (with 4-byte ones every warp store touched 32 sectors for 128 useful bytes)
#pragma unroll
      for (int q = 0; q < PER / 8; ++q) {
        uint4 rv = make_uint4(0u, 0u, 0u, 0u);
        if (res != nullptr) rv = __ldg(reinterpret_cast<const uint4*>(res + p * C + lane * PER + q * 8));
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = (q * 8 + 2 * j) % PER;
          float o0 = (f[c] - mean) * rstd * gl[c], o1 = (f[(c + 1) % PER] - mean) * rstd * gl[(c + 1) % PER];
          if (res != nullptr) {
            const float2 r2 = __half22float2(rh[j]);
            o0 += r2.x;
            o1 += r2.y;
          }
          const __half2 h = __floats2half2_rn(o0, o1);
          o[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + q * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < PER; j += 2) {
        float o0 = (f[j] - mean) * rstd * gl[j], o1 = (f[j + 1] - mean) * rstd * gl[j + 1];
        if (res != nullptr) {
          const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(res + p * C + lane * PER + j));
          o0 += r2.x;
          o1 += r2.y;
        }
        *reinterpret_cast<__half2*>(dst + j) = __floats2half2_rn(o0, o1);
      }
    }
  }
}

int ln_apply(const __half* x, const float* g, const __half* res, __half* y, int64_t npix, int C,
             cudaStream_t s) {
  int64_t blocks = (npix * 32 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  switch (C) {
    case 64: PRG_CUDA_OK(launch_pdl(k_ln_apply<64>, dim3((int)blocks), dim3(256), 0, s, x, g, res, y, npix)); break;
    case 128: PRG_CUDA_OK(launch_pdl(k_ln_apply<128>, dim3((int)blocks), dim3(256), 0, s, x, g, res, y, npix)); break;
    case 256: PRG_CUDA_OK(launch_pdl(k_ln_apply<256>, dim3((int)blocks), dim3(256), 0, s, x, g, res, y, npix)); break;
    case 512: PRG_CUDA_OK(launch_pdl(k_ln_apply<512>, dim3((int)blocks), dim3(256), 0, s, x, g, res, y, npix)); break;
    default:
      set_error("ln_apply: unsupported channel count %d", C);
      return PRG_ERR_ARG;
  }
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (device noise for throughput runs; parity runs inject noise)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long idx) {
  const uint4 r = philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u),
                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u1 = ((float)r.x + 1.f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = (float)r.y * 2.3283064365386963e-10f;
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// One Philox stream per image: key = seeds[b], counter = offset + index inside the image.  The draws of
// an image therefore depend only on its own seed -- not on the batch it travels in, its position in
// the batch, the rank or the world size.
__global__ void __launch_bounds__(256)
k_fill_normal(float* __restrict__ x, int64_t per_image, const unsigned long long* __restrict__ seeds,
              unsigned long long offset) {
  const int b = blockIdx.y;
  const unsigned long long seed = seeds[b];
  float* xi = x + (size_t)b * per_image;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image;
       i += (int64_t)gridDim.x * blockDim.x)
    xi[i] = philox_normal(seed, offset + (unsigned long long)i);
}

int fill_normal(float* x, int B, int64_t per_image, const unsigned long long* seeds_dev,
                unsigned long long offset, cudaStream_t s) {
  int blocks = (int)((per_image + 255) / 256);
  const int cap = std::max(1, num_sms() * 8 / std::max(1, B));
  if (blocks > cap) blocks = cap;
  k_fill_normal<<<dim3(blocks, B), 256, 0, s>>>(x, per_image, seeds_dev, offset);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// ------------------------------------------------------------------------------------------
// network tail: GN+SiLU(block2) + res, final 1x1 (64 -> 1), then forward / sigmoid / sampler step
// 8 lanes per pixel (8 channels each), 4 pixels per warp -> 512 contiguous bytes per warp load.
// ------------------------------------------------------------------------------------------
__global__ void k_step_advance(int* step_idx) {
  pdl_trigger();
  pdl_wait();
  *step_idx += 1;
}

int step_advance(int* step_idx, cudaStream_t s) {
  PRG_CUDA_OK(launch_pdl(k_step_advance, dim3(1), dim3(1), 0, s, step_idx));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// What one pixel does with the network output `net`: forward / sigmoid (+ keep mask) / one sampler step.
__device__ __forceinline__ void tail_pixel(const TailParams& t, float net, int b, int64_t p, size_t o) {
  if (t.mode == 0) {
    t.out[o] = net;
    return;
  }
  if (t.mode == 1) {
    const float pr = 1.f / (1.f + expf(-net));
    if (t.out != nullptr) t.out[o] = pr;
    if (t.keep != nullptr) t.keep[o] = pr > t.thresh;
    return;
  }
  const float xt = t.x_t[o];
  float x0 = net;
  if (t.clip_x_start) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  float pred_noise = 0.f;
  if (t.sampler == 1 || t.sampler == 2 || t.sampler == 4)
    pred_noise = __fdiv_rn(__fsub_rn(__fmul_rn(t.c0, xt), x0), t.c1);      // SDD:1158-1162
  bool m = false;
  if (t.img_cond != nullptr) {
    const float mk = t.img_cond[((size_t)b * 2 + 1) * t.HW + p];
    m = __fmul_rn(__fadd_rn(mk, 1.f), 0.5f) > 0.5f;                         // SDD:507-508
    if (t.use_ddnm && m) x0 = t.img_cond[((size_t)b * 2 + 0) * t.HW + p];  // SDD:1218
  }
  float nz = 0.f;
  if (t.add_noise)
    nz = (t.noise != nullptr) ? t.noise[o] : philox_normal(t.seeds[b], t.noise_offset + (unsigned long long)p);
  float xn;
  if (t.sampler == 0 || t.sampler == 3) {
    x0 = fminf(fmaxf(x0, -1.f), 1.f);                                       // SDD:1250-1251
    const float mean = __fadd_rn(__fmul_rn(t.c0, x0), __fmul_rn(t.c1, xt));  // SDD:1174-1176
    xn = __fadd_rn(mean, __fmul_rn(t.c2, nz));                              // SDD:1280
    if (t.sampler == 3) xn = m ? xn : xt;                                   // SDD:1313-1314
  } else if (t.sampler == 1) {
    xn = __fadd_rn(__fadd_rn(__fmul_rn(x0, t.c2), __fmul_rn(t.c3, pred_noise)),
                   __fmul_rn(t.c4, nz));                                   // SDD:1371-1373
  } else if (t.sampler == 2) {
    xn = x0;                                                                // SDD:1358-1360
  } else {
    xn = m ? x0 : xt;                                                       // SDD:1388-1389
  }
  if (t.unnormalize) xn = __fmul_rn(__fadd_rn(xn, 1.f), 0.5f);              // SDD:560-561
  t.out[o] = xn;
}

// sampler loop: this step's parameters come from the device-resident list (same launch for every step)
__device__ __forceinline__ void tail_resolve_step(TailParams& t) {
  if (t.mode != 2 || t.steps == nullptr) return;
  const StepDev st = t.steps[*t.step_idx];
  const SamplerCtx cx = *t.ctx;
  t.sampler = st.kind;
  t.add_noise = st.add_noise;
  t.unnormalize = st.unnormalize;
  t.c0 = st.c0; t.c1 = st.c1; t.c2 = st.c2; t.c3 = st.c3; t.c4 = st.c4;
  t.img_cond = cx.img_cond;
  t.noise = (cx.noise != nullptr && st.add_noise) ? cx.noise + (size_t)st.noise_slab * t.slab_stride : nullptr;
  t.noise_offset = (unsigned long long)st.noise_slab * (unsigned long long)t.HW;
  t.clip_x_start = (st.kind == PRG_STEP_DDIM || st.kind == PRG_STEP_DDIM_LAST || st.kind == PRG_STEP_REFINE_DDIM);
  t.use_ddnm = (cx.img_cond != nullptr) &&
               (st.kind == PRG_STEP_P_SAMPLE || st.kind == PRG_STEP_DDIM || st.kind == PRG_STEP_DDIM_LAST);
}

// A warp walks blocks of 32 consecutive pixels.  Pass k (of eight) reduces pixels 4k .. 4k + 3 (eight lanes
// per pixel, eight channels per lane: 512 contiguous bytes per warp load) and hands each result to lane
// 4k + g; after the eight passes EVERY lane owns one pixel and runs the per-pixel stage -- Philox +
// Box-Muller, DDNM, posterior / DDIM update -- so that part runs 32 lanes wide instead of on one lane
// in eight (it was more than a third of the kernel inside the sampler loop).
__global__ void __launch_bounds__(256)
k_net_tail(TailParams t) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sA[64], sB[64], sW[64];
  const int b = blockIdx.y;
  tail_resolve_step(t);
  gn_coeffs(t.stats, t.gamma, t.beta, nullptr, 64, t.HW, b, sA, sB);
  if (threadIdx.x < 64) sW[threadIdx.x] = t.fw[threadIdx.x];
  __syncthreads();
  const float fb = __ldg(t.fb);
  const int lane = threadIdx.x & 31;
  const int c0 = (lane & 7) * 8;          // which 8-channel slice
  float ca[8], cb[8], cw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ca[j] = sA[c0 + j]; cb[j] = sB[c0 + j]; cw[j] = sW[c0 + j]; }
  const int64_t nblk = ((int64_t)t.HW + 31) >> 5;
  for (int64_t blk = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); blk < nblk; blk += (int64_t)gridDim.x * 8) {
    const int64_t p0 = blk * 32;
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t p = p0 + 4 * k + (lane >> 3);
      float dot = 0.f;
      if (p < t.HW) {
        const size_t e = ((size_t)b * t.HW + p) * 64 + c0;
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(t.raw + e));
        const uint4 r = __ldcs(reinterpret_cast<const uint4*>(t.res + e));
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
        const __half2* hr = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __half22float2(hv[j]), q = __half22float2(hr[j]);
          const float y0 = silu(fmaf(a.x, ca[2 * j], cb[2 * j])) + q.x;
          const float y1 = silu(fmaf(a.y, ca[2 * j + 1], cb[2 * j + 1])) + q.y;
          dot = fmaf(y0, cw[2 * j], dot);
          dot = fmaf(y1, cw[2 * j + 1], dot);
        }
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      const float got = __shfl_sync(0xffffffffu, dot, (lane & 3) * 8);   // pixel 4k + (lane & 3) of the block
      if ((lane >> 2) == k) mine = got;
    }
    const int64_t p = p0 + lane;
    if (p < t.HW) tail_pixel(t, mine + fb, b, p, (size_t)b * t.HW + p);
  }
}

int net_tail(const TailParams& t, int B, cudaStream_t s) {
  // a CTA = eight warps, a warp takes 32-pixel blocks; about eight resident CTAs per SM over the whole batch
  int gx = (int)(((int64_t)t.HW + 255) / 256);
  const int cap = std::max(1, (num_sms() * 8 + B - 1) / std::max(1, B));
  if (gx > cap) gx = cap;
  dim3 g(gx, B);
  PRG_CUDA_OK(launch_pdl(k_net_tail, g, dim3(256), 0, s, t));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

// ------------------------------------------------------------------------------------------
// ResnetBlock output in ONE streaming pass (SDD:731-734):  y = SiLU(GN(raw2)) + res_conv(cat(x0, x1))
// [+ the PreNorm LayerNorm of y for the attention that follows].
//
// The separate form -- 1x1 conv on the tcgen05 engine (reads x, writes the shortcut tensor) followed by
// k_gn_apply (reads raw2 and the shortcut, writes y) -- is two HBM-bound passes; the fused tcgen05
// epilogue (EPI_GNRES) removed the shortcut tensor but put all the SiLU work on eight epilogue warps
// and was no faster.  Here the 1x1 conv (16 K MACs per pixel at most) rides on mma.sync inside an
// elementwise-style kernel with many resident warps: x, raw2 in, y out, nothing else.
//
// No shared-memory staging of activations: the k index of an MMA (and the n index of an n-tile) may be
// permuted freely as long as both operands agree, so the fragments are chosen such that every lane
// reads / writes 16 contiguous bytes of a pixel row:
//   A (m16 x k16, k-step s of a 32-channel group): lane (g, t) supplies channels 8t + 4s + {0, 1} as
//     k slots {2t, 2t+1} and 8t + 4s + {2, 3} as k slots {2t+8, 2t+9} of rows g and g + 8 -- one
//     LDG.128 of channels 8t .. 8t+7 per row feeds both k-steps;
//   B: W[ch][8t .. 8t+7] from shared memory (one LDS.128 feeds both k-steps);
//   C (m16 x n8, n-tile jj of a 32-channel group): column c of the tile is channel 8 (c >> 1) + 2 jj +
//     (c & 1), so lane (g, t) ends up with channels 8t .. 8t+7 of rows g and g + 8: raw2 in and y out are
//     LDG.128 / STG.128.
// ------------------------------------------------------------------------------------------
//
// TAIL = 1 (the network's final block, COUT = 64): y is not written at all -- it goes straight into the
// final 1x1 conv (64 -> 1, a dot product over the four lanes that hold a pixel's channels) and the
// per-pixel stage of the network tail (forward / sigmoid / sampler step, see tail_pixel).
template <int COUT, int CIN, bool LN, bool TAIL>
__global__ void __launch_bounds__(256, 2)
k_res1x1_gn(ResGn a, TailParams tp, int total_blocks, int nblk) {
  pdl_trigger();
  constexpr int NH = COUT / 64;            // the output channels are produced in halves of 64 (two 32-channel groups):
                                           // 32 accumulator registers at a time, so that 16 warps per SM fit at COUT = 128
  constexpr int KG = CIN / 32;             // input channel groups
  constexpr int kPitch = CIN * 2 + 64;     // bytes per weight row in shared memory (conflict-free LDS.128)
  extern __shared__ __align__(16) uint8_t rsm[];
  uint8_t* sW = rsm;
  float* sCA = reinterpret_cast<float*>(rsm + COUT * kPitch);     // y = SiLU(sCA[c] * raw + sCB[c]) + ...
  float* sCB = sCA + COUT;
  float* sBias = sCB + COUT;
  float* sG = sBias + COUT;
  // weights / bias / gain: not written by any kernel of the evaluation, so before the dependency wait
  for (int i = threadIdx.x; i < COUT * (CIN / 8); i += 256) {
    const int r = i / (CIN / 8), c = i - r * (CIN / 8);
    *reinterpret_cast<uint4*>(sW + r * kPitch + c * 16) = __ldg(reinterpret_cast<const uint4*>(a.w + (size_t)r * CIN) + c);
  }
  for (int i = threadIdx.x; i < COUT; i += 256) {
    sBias[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;
    sG[i] = LN ? __ldg(a.ln_g + i) : (TAIL ? __ldg(tp.fw + i) : 0.f);     // TAIL: weight of the final 1x1 conv
  }
  const float fb = TAIL ? __ldg(tp.fb) : 0.f;
  pdl_wait();
  if (TAIL) tail_resolve_step(tp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int g_begin = (int)(((long long)blockIdx.x * total_blocks) / gridDim.x);
  const int g_end = (int)(((long long)(blockIdx.x + 1) * total_blocks) / gridDim.x);
  // weight row of n-tile jj of 32-channel group ng for this lane's B fragment: channel 32 ng + 8 (g >> 1) + 2 jj + (g & 1)
  const uint8_t* wrow = sW + (8 * (g >> 1) + (g & 1)) * kPitch + t * 16;
  int b = -1;
  for (int blk = g_begin; blk < g_end; ++blk) {
    const int gb = blk / nblk, pb = blk - gb * nblk;
    if (gb != b) {
      b = gb;
      __syncthreads();                     // everyone is done with the previous image's coefficients
      gn_coeffs(a.stats, a.gamma, a.beta, nullptr, COUT, a.HW, b, sCA, sCB);   // block2's GroupNorm as (A, B) per channel
      __syncthreads();
    }
    const size_t row0 = (size_t)b * a.HW + (size_t)pb * 128 + warp * 16 + g;    // pixel rows row0 and row0 + 8
    float sum0 = 0.f, sum1 = 0.f;          // LN: channel sums of the two rows; TAIL: their dot products with the final weight
    uint32_t pk0[LN ? NH * 8 : 1], pk1[LN ? NH * 8 : 1];     // LN: the fp16-rounded y of this lane (packed pairs)
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      // ---- every load of the half first (the x loads of the second half hit L1 / L2)
      uint4 xa[KG], xb[KG], ra[2], rb[2];
#pragma unroll
      for (int kg = 0; kg < KG; ++kg) {
        const int c = kg * 32;
        const __half* src = (c < a.c0) ? a.x0 + row0 * a.c0 + c : a.x1 + row0 * a.c1 + (c - a.c0);
        const size_t step = (c < a.c0) ? (size_t)8 * a.c0 : (size_t)8 * a.c1;
        if (NH == 1) {
          xa[kg] = __ldcs(reinterpret_cast<const uint4*>(src) + t);
          xb[kg] = __ldcs(reinterpret_cast<const uint4*>(src + step) + t);
        } else {
          xa[kg] = __ldg(reinterpret_cast<const uint4*>(src) + t);
          xb[kg] = __ldg(reinterpret_cast<const uint4*>(src + step) + t);
        }
      }
#pragma unroll
      for (int ng = 0; ng < 2; ++ng) {
        const __half* src = a.raw + row0 * COUT + hf * 64 + ng * 32;
        ra[ng] = __ldcs(reinterpret_cast<const uint4*>(src) + t);
        rb[ng] = __ldcs(reinterpret_cast<const uint4*>(src + 8 * COUT) + t);
      }
      // ---- shortcut 1x1 conv, 64 output channels
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int kg = 0; kg < KG; ++kg) {
        const uint32_t f0[4] = {xa[kg].x, xb[kg].x, xa[kg].y, xb[kg].y};
        const uint32_t f1[4] = {xa[kg].z, xb[kg].z, xa[kg].w, xb[kg].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 wv = *reinterpret_cast<const uint4*>(wrow + (hf * 64 + (j >> 2) * 32 + (j & 3) * 2) * kPitch + kg * 64);
          hmma16816(acc[j], f0, wv.x, wv.y);
          hmma16816(acc[j], f1, wv.z, wv.w);
        }
      }
      // ---- y = SiLU(A raw + B) + shortcut + bias, rounded to fp16
#pragma unroll
      for (int ng = 0; ng < 2; ++ng) {
        const int ch = hf * 64 + ng * 32 + t * 8;
        const __half2* h0 = reinterpret_cast<const __half2*>(&ra[ng]);
        const __half2* h1 = reinterpret_cast<const __half2*>(&rb[ng]);
        uint32_t o0[4], o1[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float2 cA = *reinterpret_cast<const float2*>(sCA + ch + 2 * jj);       // two channels
          const float2 cB = *reinterpret_cast<const float2*>(sCB + ch + 2 * jj);
          const float2 bi = *reinterpret_cast<const float2*>(sBias + ch + 2 * jj);
          const float* c = acc[ng * 4 + jj];
          float2 y0 = silu2(__ffma2_rn(__half22float2(h0[jj]), cA, cB));
          float2 y1 = silu2(__ffma2_rn(__half22float2(h1[jj]), cA, cB));
          y0 = __fadd2_rn(y0, __fadd2_rn(make_float2(c[0], c[1]), bi));
          y1 = __fadd2_rn(y1, __fadd2_rn(make_float2(c[2], c[3]), bi));
          if (TAIL) {                       // final 1x1 conv: this lane's share of the two pixels' dot products
            const float2 fw = *reinterpret_cast<const float2*>(sG + ch + 2 * jj);
            sum0 = fmaf(y0.y, fw.y, fmaf(y0.x, fw.x, sum0));
            sum1 = fmaf(y1.y, fw.y, fmaf(y1.x, fw.x, sum1));
            continue;
          }
          const __half2 q0 = __floats2half2_rn(y0.x, y0.y), q1 = __floats2half2_rn(y1.x, y1.y);
          o0[jj] = *reinterpret_cast<const uint32_t*>(&q0);
          o1[jj] = *reinterpret_cast<const uint32_t*>(&q1);
          if (LN) {
            const float2 r0 = __half22float2(q0), r1 = __half22float2(q1);
            pk0[(hf * 2 + ng) * 4 + jj] = o0[jj];
            pk1[(hf * 2 + ng) * 4 + jj] = o1[jj];
            sum0 += r0.x + r0.y;
            sum1 += r1.x + r1.y;
          }
        }
        if (TAIL) continue;
        __half* dst = a.y + row0 * COUT + ch;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
        *reinterpret_cast<uint4*>(dst + 8 * COUT) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
      }
    }
    if (TAIL) {
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      // lane t = 0 finishes the pixel of row g, lane t = 1 the pixel of row g + 8
      if (t < 2) {
        const int64_t p = (int64_t)pb * 128 + warp * 16 + g + 8 * t;
        tail_pixel(tp, (t == 0 ? sum0 : sum1) + fb, b, p, (size_t)b * a.HW + p);
      }
      continue;
    }
    if (LN) {
      // channel LayerNorm of the fp16-rounded y (what the attention's PreNorm would read back), two-pass
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      const float mean0 = sum0 * (1.f / COUT), mean1 = sum1 * (1.f / COUT);
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int i = 0; i < NH * 8; ++i) {
        const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&pk0[i]));
        const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&pk1[i]));
        const float d0 = r0.x - mean0, d1 = r0.y - mean0, d2 = r1.x - mean1, d3 = r1.y - mean1;
        q0 = fmaf(d0, d0, q0); q0 = fmaf(d1, d1, q0);
        q1 = fmaf(d2, d2, q1); q1 = fmaf(d3, d3, q1);
      }
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      const float rs0 = rsqrtf(q0 * (1.f / COUT) + 1e-5f), rs1 = rsqrtf(q1 * (1.f / COUT) + 1e-5f);
#pragma unroll
      for (int gi = 0; gi < NH * 2; ++gi) {
        const int ch = gi * 32 + t * 8;
        uint32_t o0[4], o1[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float2 gn = *reinterpret_cast<const float2*>(sG + ch + 2 * jj);
          const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&pk0[gi * 4 + jj]));
          const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&pk1[gi * 4 + jj]));
          const __half2 l0 = __floats2half2_rn((r0.x - mean0) * rs0 * gn.x, (r0.y - mean0) * rs0 * gn.y);
          const __half2 l1 = __floats2half2_rn((r1.x - mean1) * rs1 * gn.x, (r1.y - mean1) * rs1 * gn.y);
          o0[jj] = *reinterpret_cast<const uint32_t*>(&l0);
          o1[jj] = *reinterpret_cast<const uint32_t*>(&l1);
        }
        __half* dst = a.ln_out + row0 * COUT + ch;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
        *reinterpret_cast<uint4*>(dst + 8 * COUT) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
      }
    }
  }
}

bool res1x1_gn_supported(int cout, int c0, int c1, int HW) {
  return HW % 128 == 0 && c0 % 32 == 0 && c1 % 32 == 0 &&
         ((cout == 64 && c0 + c1 == 128) || (cout == 128 && c0 + c1 == 192));
}

template <int COUT, int CIN>
static int res1x1_gn_launch(const ResGn& a, const TailParams* tail, int B, cudaStream_t s) {
  const int nblk = a.HW / 128, total = nblk * B;
  const size_t smem = (size_t)COUT * (CIN * 2 + 64) + COUT * 4 * sizeof(float);
  int grid = num_sms() * 2;                             // one resident wave
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  const TailParams tp = tail ? *tail : TailParams{};
  auto go = [&](auto kern) -> int {
    PRG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
    PRG_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(256), smem, s, a, tp, total, nblk));
    PRG_LAUNCH_CHECK();
    return PRG_OK;
  };
  if (tail != nullptr) {
    if (COUT != 64) { set_error("res1x1_gn: the fused tail needs 64 channels"); return PRG_ERR_ARG; }
    return go(k_res1x1_gn<64, 128, false, true>);
  }
  if (a.ln_out != nullptr) return go(k_res1x1_gn<COUT, CIN, true, false>);
  return go(k_res1x1_gn<COUT, CIN, false, false>);
}

int res1x1_gn(const ResGn& a, int B, cudaStream_t s) {
  if (!res1x1_gn_supported(a.Cout, a.c0, a.c1, a.HW)) {
    set_error("res1x1_gn: unsupported shape (Cout %d, Cin %d + %d, HW %d)", a.Cout, a.c0, a.c1, a.HW);
    return PRG_ERR_ARG;
  }
  if (a.Cout == 64) return res1x1_gn_launch<64, 128>(a, nullptr, B, s);
  return res1x1_gn_launch<128, 192>(a, nullptr, B, s);
}

int net_tail_fused(const TailParams& t, const ResGn& a, int B, cudaStream_t s) {
  if (a.Cout != 64 || !res1x1_gn_supported(a.Cout, a.c0, a.c1, a.HW) || t.HW != a.HW) {
    set_error("net_tail_fused: unsupported shape (Cout %d, Cin %d + %d, HW %d)", a.Cout, a.c0, a.c1, a.HW);
    return PRG_ERR_ARG;
  }
  return res1x1_gn_launch<64, 128>(a, &t, B, s);
}

}  // namespace prg

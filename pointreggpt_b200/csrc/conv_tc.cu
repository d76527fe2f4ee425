// tcgen05 implicit-GEMM convolution / GEMM engine (sm_100a).
//
// Replaces, for the U-Net denoiser and the depth-correction U-Net, the cuDNN/cuBLAS calls
// behind F.conv2d / nn.Conv2d at SDD:594-598, 615, 717, 743-745, 779-780 (same lines in DC).
//
// Warp roles in a 192-thread CTA (one 128 x BN output tile per CTA, 2 CTAs co-resident per SM
// so one CTA's epilogue overlaps the other's main loop):
//   warp 0   : TMA producer (one elected lane): per K block (tap, 64 input channels) one
//              4-D/5-D box load of the shifted activation patch + one box of the weight matrix,
//              both into 128B-swizzled K-major stages, completion on an mbarrier (expect_tx).
//   warp 1   : allocates TMEM, then one lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per
//              stage and tcgen05.commit's the stage back to the producer.
//   warps 2-5: epilogue: tcgen05.ld the fp32 accumulators (32 lanes x 32 columns per load),
//              fused bias / GroupNorm statistics / softmax / LayerNorm / residual, fp16 stores.
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <limits.h>
#include <string.h>

namespace prg {

using namespace ptx;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // fp16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kThreads = 192;

template <int BN>
struct Cfg {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 64) ? 4 : (BN == 128) ? 3 : 2;  // ~96 KiB -> 2 CTAs/SM
  static constexpr int kSmem = kStages * kStageBytes + 1024 /*align*/ + 1024 /*barriers+scratch*/;
};

struct alignas(8) SmemCtl {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t tmem_full;
  uint32_t tmem_addr;
  uint32_t pad;
  float stats[64];      // EPI_GN: [epilogue warp][up to 8 groups][sum, sumsq]
  int colmax[128];      // EPI_QKV (k tile): per-column max across the 4 epilogue warps
};
static_assert(sizeof(SmemCtl) <= 1024, "control block too large");

constexpr double kStatScale = 1048576.0;  // 2^20 fixed point, see conv_tc.cuh

__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 2)
k_conv_tc(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + C::kStages * C::kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int img = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - img * tiles_per_img;
  const int tyi = t_in / p.tiles_x, txi = t_in - tyi * p.tiles_x;
  const int tile_w = 1 << p.tile_w_log2;
  const int x0 = txi << p.tile_w_log2;
  const int y0 = tyi * (kBlockM >> p.tile_w_log2);
  const int n0 = blockIdx.y * BN;
  const int cls = blockIdx.z;
  const int cpy = cls >> 1, cpx = cls & 1;

  const int chunks = p.chunks0 + p.chunks1;
  const int ntaps = (p.mode == 1) ? 16 : p.kh * p.kw;
  const int num_kb = ntaps * chunks;

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (p.chunks1 > 0) prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    mbar_init(&ctl->tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_addr, BN);
    tmem_relinquish();
  }
  if (EPI == EPI_GN && threadIdx.x < 64) ctl->stats[threadIdx.x] = 0.f;
  if (EPI == EPI_QKV && threadIdx.x < 128) ctl->colmax[threadIdx.x] = INT_MIN;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = ctl->tmem_addr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int pad_y = p.pad, pad_x = p.pad;
      if (p.classes == 4) { pad_y = 1 - cpy; pad_x = 1 - cpx; }
      const int wz = p.w_batched ? img : 0;
      const int wk0 = cls * num_kb * kBlockK;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / chunks, cc = kb - tap * chunks;
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        uint8_t* sA = smem + stage * C::kStageBytes;
        uint8_t* sB = sA + kABytes;
        mbar_arrive_expect_tx(&ctl->full[stage], C::kStageBytes);
        if (p.mode == 0) {
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          const int dy = ky - pad_y, dx = kx - pad_x;
          if (cc < p.chunks0)
            tma_load_4d(&tmA0, &ctl->full[stage], sA, cc * kBlockK, x0 + dx, y0 + dy, img);
          else
            tma_load_4d(&tmA1, &ctl->full[stage], sA, (cc - p.chunks0) * kBlockK, x0 + dx, y0 + dy,
                        img);
        } else {
          // input row 2*oy + ky - 1 = 2*(oy + qy) + ry with ry in {0,1}
          const int ey = (tap >> 2) - 1, ex = (tap & 3) - 1;
          const int qy = ey >> 1, ry = ey & 1, qx = ex >> 1, rx = ex & 1;
          tma_load_5d(&tmA0, &ctl->full[stage], sA, rx * p.cin0 + cc * kBlockK, x0 + qx, ry, y0 + qy,
                      img);
        }
        tma_load_3d(&tmB, &ctl->full[stage], sB, wk0 + kb * kBlockK, n0, wz);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    constexpr uint32_t idesc = idesc_f16(kBlockM, BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&ctl->full[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + stage * C::kStageBytes);
        const uint64_t da = smem_desc_sw128(a_addr);
        const uint64_t db = smem_desc_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the (>>4) address
          umma_f16(taddr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                   (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&ctl->empty[stage]);
        if (kb == num_kb - 1) umma_commit(&ctl->tmem_full);
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // =============================== epilogue ===================================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;          // row of the 128 x BN tile
    const int tyr = row >> p.tile_w_log2, txr = row & (tile_w - 1);
    const int oy = (y0 + tyr) * p.out_scale + cpy, ox = (x0 + txr) * p.out_scale + cpx;
    const long long off = (long long)img * p.out_img_stride + (long long)oy * p.out_row_stride +
                          (long long)ox * p.out_pix_stride + n0;
    __half* orow = p.out + off;
    const uint32_t trow = taddr + ((uint32_t)(quarter * 32) << 16);

    mbar_wait(&ctl->tmem_full, 0);
    tc_fence_after();

    float ln_mean = 0.f, ln_rstd = 0.f;
    if (EPI == EPI_LN_RES) {
      // channel LayerNorm over the whole row (BN == Cout): exact two-pass mean / variance
      float s = 0.f;
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) s += __uint_as_float(v[j]) + __ldg(p.bias + n0 + c + j);
      }
      ln_mean = s * (1.f / BN);
      float ss = 0.f;
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(v[j]) + __ldg(p.bias + n0 + c + j) - ln_mean;
          ss += d * d;
        }
      }
      ln_rstd = rsqrtf(ss * (1.f / BN) + 1e-5f);
    }
    const int qkv_part = (EPI == EPI_QKV) ? (n0 >> 7) : 0;  // 0 = q, 1 = k, 2 = v

#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(trow + c, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + n0 + c + j);
      }

      if (EPI == EPI_GN) {
        float s4[4], q4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = f[g * 8 + j];
            s += x;
            q = fmaf(x, x, q);
          }
          s4[g] = s;
          q4[g] = q;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            s4[g] += __shfl_xor_sync(0xffffffffu, s4[g], o);
            q4[g] += __shfl_xor_sync(0xffffffffu, q4[g], o);
          }
        }
        if (lane == 0) {
          // this warp's private slots, accumulated in program order (deterministic)
          float* ws = ctl->stats + quarter * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int grp = (c + g * 8) >> p.gs_log2;  // group index inside this tile
            ws[grp * 2 + 0] += s4[g];
            ws[grp * 2 + 1] += q4[g];
          }
        }
      } else if (EPI == EPI_QKV) {
        if (qkv_part == 0) {
          if (p.q_softmax) {
            float m = f[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) m = fmaxf(m, f[j]);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              f[j] = fast_exp(f[j] - m);
              s += f[j];
            }
            const float inv = p.q_scale / s;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= inv;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= p.q_scale;
          }
        } else if (qkv_part == 1 && p.colmax != nullptr) {
          // column max over the tile rows of the stored (fp16-rounded) k: that is the value
          // the context kernel exponentiates.  lane j ends up owning column c + j.
          int mine = INT_MIN;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float r = __half2float(__float2half_rn(f[j]));
            const int m = __reduce_max_sync(0xffffffffu, float_to_ordered(r));
            if (lane == j) mine = m;
          }
          atomicMax(&ctl->colmax[c + lane], mine);
        }
      } else if (EPI == EPI_LN_RES) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          f[j] = (f[j] - ln_mean) * ln_rstd * __ldg(p.ln_g + n0 + c + j);
      }

      if (EPI == EPI_RES || EPI == EPI_LN_RES) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + off + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 rv = __ldg(rp + q);
          const __half2* h = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 r2 = __half22float2(h[j]);
            f[q * 8 + j * 2 + 0] += r2.x;
            f[q * 8 + j * 2 + 1] += r2.y;
          }
        }
      }

      uint4* op = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        __half2 h0 = __floats2half2_rn(f[q * 8 + 0], f[q * 8 + 1]);
        __half2 h1 = __floats2half2_rn(f[q * 8 + 2], f[q * 8 + 3]);
        __half2 h2 = __floats2half2_rn(f[q * 8 + 4], f[q * 8 + 5]);
        __half2 h3 = __floats2half2_rn(f[q * 8 + 6], f[q * 8 + 7]);
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2);
        o.w = *reinterpret_cast<uint32_t*>(&h3);
        op[q] = o;
      }
    }

    if (EPI == EPI_GN) {
      // 4 epilogue warps -> one global atomic per (group, moment) of this tile.  The cross-CTA
      // sum is a 64-bit fixed-point integer add: order-independent, hence bit-reproducible
      // (an fp32 atomic's rounding would depend on arrival order, and the fp16 re-rounding of
      // the activations amplifies even 1e-7 differences to the 5e-4 level a few layers on).
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - 64;
      const int ngrp = BN >> p.gs_log2;
      if (e < ngrp * 2) {
        const int g0 = n0 >> p.gs_log2;
        const float v = (ctl->stats[e] + ctl->stats[16 + e]) + (ctl->stats[32 + e] + ctl->stats[48 + e]);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + ((size_t)img * 8 + g0) * 2 + e,
                  (unsigned long long)__double2ll_rn((double)v * kStatScale));
      }
    }
    if (EPI == EPI_QKV) {
      if (qkv_part == 1 && p.colmax != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int e = threadIdx.x - 64;
        atomicMax(&p.colmax[img * 128 + e], ctl->colmax[e]);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(taddr, BN);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return PRG_ERR_CUDA;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s,
                  b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0));
    return PRG_ERR_CUDA;
  }
  return PRG_OK;
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

int conv_plan(ConvLaunch* L, int epi, int B, const ActSrc& s0, const ActSrc* s1, int mode, int ksize,
              int classes, const __half* w, int w_batched, int Cout) {
  memset(L, 0, sizeof(*L));
  ConvParams& p = L->p;
  const int Ho = (mode == 1) ? s0.H / 2 : s0.H, Wo = (mode == 1) ? s0.W / 2 : s0.W;
  if (s0.C % 64 != 0 || (s1 && s1->C % 64 != 0) || Cout % 64 != 0) {
    set_error("conv_plan: channel counts must be multiples of 64 (%d,%d->%d)", s0.C,
              s1 ? s1->C : 0, Cout);
    return PRG_ERR_ARG;
  }
  if ((Ho * Wo) % kBlockM != 0) {
    set_error("conv_plan: %dx%d output is not a multiple of the 128-pixel tile", Ho, Wo);
    return PRG_ERR_ARG;
  }
  if (mode == 1 && (s1 != nullptr || ksize != 4 || classes != 1)) {
    set_error("conv_plan: stride-2 mode takes one source, 4x4 taps");
    return PRG_ERR_ARG;
  }
  int tile_w = Wo < kBlockM ? Wo : kBlockM;
  if ((tile_w & (tile_w - 1)) != 0 || tile_w < 8 || (kBlockM / tile_w) > Ho ||
      Ho % (kBlockM / tile_w) != 0 || Wo % tile_w != 0) {
    set_error("conv_plan: unsupported spatial size %dx%d", Ho, Wo);
    return PRG_ERR_ARG;
  }
  const int tile_h = kBlockM / tile_w;
  p.B = B; p.Ho = Ho; p.Wo = Wo;
  p.tile_w_log2 = ilog2(tile_w);
  p.tiles_x = Wo / tile_w; p.tiles_y = Ho / tile_h;
  p.mode = mode;
  p.kh = p.kw = (classes == 4) ? 2 : ksize;
  p.pad = (ksize == 3) ? 1 : 0;
  p.chunks0 = s0.C / 64; p.chunks1 = s1 ? s1->C / 64 : 0;
  p.cin0 = s0.C;
  p.classes = classes;
  p.w_batched = w_batched;
  p.out_scale = (classes == 4) ? 2 : 1;
  L->epi = epi;
  L->cout = Cout;

  int bn = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
  if (epi == EPI_QKV) bn = 128;
  if (epi == EPI_LN_RES) bn = Cout;
  if (bn != 64 && bn != 128 && bn != 256) {
    set_error("conv_plan: unsupported N tile %d", bn);
    return PRG_ERR_ARG;
  }
  L->bn = bn;
  L->grid = dim3((unsigned)(p.tiles_x * p.tiles_y * B), (unsigned)(Cout / bn), (unsigned)classes);

  // activation maps
  for (int si = 0; si < 2; ++si) {
    const ActSrc* s = si == 0 ? &s0 : s1;
    CUtensorMap* tm = si == 0 ? &L->tmA0 : &L->tmA1;
    if (s == nullptr) { *tm = L->tmA0; continue; }
    const uint64_t ps = (uint64_t)s->pix_stride * 2;
    if (mode == 0) {
      uint64_t dims[4] = {(uint64_t)s->C, (uint64_t)s->W, (uint64_t)s->H, (uint64_t)B};
      uint64_t str[3] = {ps, ps * s->W, ps * s->W * s->H};
      uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, 1};
      int rc = encode(tm, s->ptr, 4, dims, str, box);
      if (rc) return rc;
    } else {
      // (2C [x parity, channel], W/2, 2 [y parity], H/2, B)
      if (s->pix_stride != s->C) {
        set_error("conv_plan: stride-2 source must be dense");
        return PRG_ERR_ARG;
      }
      uint64_t dims[5] = {(uint64_t)2 * s->C, (uint64_t)s->W / 2, 2, (uint64_t)s->H / 2, (uint64_t)B};
      uint64_t str[4] = {2 * ps, ps * s->W, 2 * ps * s->W, ps * s->W * s->H};
      uint32_t box[5] = {64, (uint32_t)tile_w, 1, (uint32_t)tile_h, 1};
      int rc = encode(tm, s->ptr, 5, dims, str, box);
      if (rc) return rc;
    }
  }
  // weight map: (K, Cout, nb*classes folded into K for classes; nb as 3rd dim)
  {
    const int cin = s0.C + (s1 ? s1->C : 0);
    const int ntaps = (mode == 1) ? 16 : p.kh * p.kw;
    const uint64_t ktot = (uint64_t)classes * ntaps * cin;
    // layout in memory: [nb][Cout][classes*ntaps*cin]
    uint64_t dims[3] = {ktot, (uint64_t)Cout, (uint64_t)(w_batched ? B : 1)};
    uint64_t str[2] = {ktot * 2, ktot * 2 * Cout};
    uint32_t box[3] = {64, (uint32_t)bn, 1};
    int rc = encode(&L->tmB, w, 3, dims, str, box);
    if (rc) return rc;
  }
  return PRG_OK;
}

template <int BN, int EPI>
static int launch_one(const ConvLaunch& L, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv_tc<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<BN>::kSmem));
    configured = true;
  }
  k_conv_tc<BN, EPI><<<L.grid, kThreads, Cfg<BN>::kSmem, stream>>>(L.tmA0, L.tmA1, L.tmB, L.p);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

int conv_run(const ConvLaunch& L, cudaStream_t stream) {
#define PRG_CASE(BN_, EPI_) \
  if (L.bn == BN_ && L.epi == EPI_) return launch_one<BN_, EPI_>(L, stream);
  PRG_CASE(64, EPI_BIAS) PRG_CASE(128, EPI_BIAS) PRG_CASE(256, EPI_BIAS)
  PRG_CASE(64, EPI_GN) PRG_CASE(128, EPI_GN) PRG_CASE(256, EPI_GN)
  PRG_CASE(128, EPI_QKV)
  PRG_CASE(64, EPI_LN_RES) PRG_CASE(128, EPI_LN_RES) PRG_CASE(256, EPI_LN_RES)
  PRG_CASE(64, EPI_RES) PRG_CASE(128, EPI_RES) PRG_CASE(256, EPI_RES)
#undef PRG_CASE
  set_error("conv_run: no kernel for N tile %d / epilogue %d", L.bn, L.epi);
  return PRG_ERR_ARG;
}

}  // namespace prg

// LinearAttention, fused k/v projection + context (tcgen05, sm_100a).
//
// Replaces, for  k = softmax_n(W_k xn),  v = W_v xn / n,  ctx = k v^T  (SDD:750-761), the k and v
// thirds of the to_qkv GEMM *and* the context reduction: k and v never reach HBM.
//
// Per 128-pixel tile of xn (NHWC fp16, TMA):
//   MMA1   D_k[d][px] = W_k[d][:] . xn[px][:]   and   D_v[e][px] = W_v[e][:] . xn[px][:]
//          (weights are the M operand, pixels the N operand: TMEM lane = channel, column = pixel,
//          so everything a softmax over pixels needs is local to one thread)
//   k-warps (one thread per channel d): tile max m, p = exp(k - m) as fp16 -> P[d][px] in
//          shared memory (K-major, 128B swizzle), z = sum p
//   v-warps (one thread per channel e): v as fp16 -> Vt[e][px] in shared memory
//   MMA2   D2[d][e] = sum_px P[d][px] Vt[e][px]   (M = N = K = 128; the four 32x32 diagonal
//          blocks are the per-head contexts)
//   k-warps: flash-style running (m, z, ctx[32]) per thread, rescaled when the running max moves.
// A CTA owns a contiguous range of tiles; per image it touches it writes one partial
// (m[128], z[128], ctx[128][32]); k_linattn_fold combines the partials of an image in a fixed
// order (bit-reproducible, no atomics) and folds the normalised context into the to_out weight:
//   W_eff[b][c][h*32+d] = sum_e W_out[c][h*32+e] ctx_b[h][d][e] / (z_b[h*32+d] n)      (SDD:763-769)
#include <algorithm>
#include <limits.h>
#include <string.h>

#include "attention.cuh"
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace prg {

using namespace ptx;

namespace {

constexpr int kThreads = 320;           // TMA warp, MMA warp, 4 k-warps, 4 v-warps
constexpr int kXBytes = 128 * 128;      // xn K block: 128 pixels x 64 ch
constexpr int kWBytes = 256 * 128;      // W K block: (k rows, v rows) x 64 ch
constexpr int kStageBytes = kXBytes + kWBytes;
constexpr int kPvBytes = 2 * 128 * 128; // P or Vt: 128 rows x 128 pixels fp16 (two 64-pixel K blocks)
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 227 * 1024;
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(8) Ctl {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t k_full[2], v_full[2], k_empty[2], v_empty[2], d2_full[2];
  uint64_t pv_ready;
  uint32_t tmem_addr, pad;
};

struct Params {
  int tile_w_log2, tiles_x, tpi;   // tile geometry, tiles per image
  int total_tiles, num_kb, stages, max_slots;
  float* partials;                 // [B][max_slots][kPartialFloats]
};

__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {   // packed fp16 exp2
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float ex2_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// first tile of CTA c when `total` tiles are dealt in contiguous ranges to `grid` CTAs
__host__ __device__ inline int range_begin(int c, int total, int grid) {
  return (int)(((long long)c * total) / grid);
}
// the CTA whose range contains tile t
__host__ __device__ inline int range_owner(int t, int total, int grid) {
  int c = (int)(((long long)t * grid) / total);
  while (c + 1 < grid && range_begin(c + 1, total, grid) <= t) ++c;
  while (c > 0 && range_begin(c, total, grid) > t) --c;
  return c;
}

__global__ void __launch_bounds__(kThreads, 1)
k_kvctx(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
        const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sStage = smem;
  uint8_t* sP = sStage + (size_t)P.stages * kStageBytes;
  uint8_t* sVt = sP + kPvBytes;
  Ctl* ctl = reinterpret_cast<Ctl*>(sVt + kPvBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_h = 128 >> P.tile_w_log2;
  const int t_begin = range_begin(blockIdx.x, P.total_tiles, gridDim.x);
  const int t_end = range_begin(blockIdx.x + 1, P.total_tiles, gridDim.x);
  const int ntiles = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctl->k_full[b], 1);
      mbar_init(&ctl->v_full[b], 1);
      mbar_init(&ctl->d2_full[b], 1);
      mbar_init(&ctl->k_empty[b], 128);
      mbar_init(&ctl->v_empty[b], 128);
    }
    mbar_init(&ctl->pv_ready, 256);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = ctl->tmem_addr;
  // TMEM columns: [K0 | V0 | K1 | V1], 128 each; D2 of tile i reuses V(i & 1)
  auto k_cols = [&](int b) { return taddr + (uint32_t)(b * 256); };
  auto v_cols = [&](int b) { return taddr + (uint32_t)(b * 256 + 128); };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < ntiles; ++i) {
        const int t = t_begin + i;
        const int img = t / P.tpi, r = t - img * P.tpi;
        const int tyi = r / P.tiles_x, txi = r - tyi * P.tiles_x;
        for (int kb = 0; kb < P.num_kb; ++kb) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          uint8_t* dst = sStage + (size_t)stage * kStageBytes;
          mbar_arrive_expect_tx(&ctl->full[stage], kStageBytes);
          tma_load_4d(&tmX, &ctl->full[stage], dst, kb * 64, txi << P.tile_w_log2, tyi * tile_h, img);
          tma_load_3d(&tmW, &ctl->full[stage], dst + kXBytes, kb * 64, 128, 0);   // rows 128..383 = k, v
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    constexpr uint32_t idesc = idesc_f16(128, 128);
    const uint32_t taddr_u = __reduce_or_sync(0xffffffffu, taddr);
    const uint32_t sStage_u = __reduce_or_sync(0xffffffffu, smem_u32(sStage));
    const uint32_t sP_u = __reduce_or_sync(0xffffffffu, smem_u32(sP));
    const uint32_t sVt_u = __reduce_or_sync(0xffffffffu, smem_u32(sVt));
    const uint64_t desc_hi = smem_desc_sw128(0);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i <= ntiles; ++i) {
      if (i < ntiles) {
        const int b = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        mbar_wait(&ctl->k_empty[b], par ^ 1u);
        mbar_wait(&ctl->v_empty[b], par ^ 1u);
        tc_fence_after();
        const uint32_t dk = taddr_u + (uint32_t)(b * 256), dv = dk + 128u;
        for (int kb = 0; kb < P.num_kb; ++kb) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t x_lo = (sStage_u + (uint32_t)stage * kStageBytes) >> 4;
          if (elect_one()) {
            const uint64_t dx = desc_hi | (uint64_t)x_lo;                       // N operand: pixels
            const uint64_t dwk = desc_hi | (uint64_t)(x_lo + (kXBytes >> 4));   // M operand: k rows
            const uint64_t dwv = dwk + (uint64_t)((128 * 128) >> 4);            //            v rows
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dk, dwk + (uint64_t)(2 * k), dx + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dv, dwv + (uint64_t)(2 * k), dx + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&ctl->empty[stage]);
            if (kb == P.num_kb - 1) {
              umma_commit(&ctl->k_full[b]);
              umma_commit(&ctl->v_full[b]);
            }
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (i >= 1) {
        const int j = i - 1, bj = j & 1;
        mbar_wait(&ctl->pv_ready, (uint32_t)j & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d2 = taddr_u + (uint32_t)(bj * 256 + 128);
          const uint64_t dp = desc_hi | (uint64_t)(sP_u >> 4), dvt = desc_hi | (uint64_t)(sVt_u >> 4);
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const uint64_t o = (uint64_t)((s >> 2) * (16384 >> 4) + (s & 3) * 2);
            umma_f16(d2, dp + o, dvt + o, idesc, s > 0 ? 1u : 0u);
          }
          umma_commit(&ctl->d2_full[bj]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // =============================== k-warps ====================================
    const int quarter = warp & 3;            // TMEM lane quarter = head
    const int d = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m_run = -INFINITY, z = 0.f;
    float acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0.f;
    float m_prev = 0.f, z_prev = 0.f;        // tile-local max / sum of the tile whose D2 is pending

    auto epi2 = [&](int j) {                 // fold tile j's D2 into the running state
      const int bj = j & 1;
      mbar_wait(&ctl->d2_full[bj], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(v_cols(bj) + lane_off + (uint32_t)(quarter * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&ctl->v_empty[bj]);
      const float m_new = fmaxf(m_run, m_prev);
      const float so = ex2_f((m_run - m_new) * kLog2e), sn = ex2_f((m_prev - m_new) * kLog2e);
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = fmaf(acc[e], so, __uint_as_float(v[e]) * sn);
      z = fmaf(z, so, z_prev * sn);
      m_run = m_new;
    };
    auto flush = [&](int img) {
      const int first = range_owner(img * P.tpi, P.total_tiles, gridDim.x);
      float* dst = P.partials + ((size_t)img * P.max_slots + (blockIdx.x - first)) * kPartialFloats;
      dst[d] = m_run;
      dst[128 + d] = z;
      float4* c4 = reinterpret_cast<float4*>(dst + 256 + d * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) c4[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
      m_run = -INFINITY;
      z = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = 0.f;
    };

    for (int i = 0; i < ntiles; ++i) {
      const int b = i & 1;
      mbar_wait(&ctl->k_full[b], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const uint32_t kaddr = k_cols(b) + lane_off;
      // pass 1: this channel's maximum over the tile's 128 pixels
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(kaddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2)
          m_tile = fmaxf(m_tile, fmaxf(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1])));
      }
      // tile i-1: its P / Vt have been consumed once D2 is complete
      if (i > 0) {
        epi2(i - 1);
        const int img_prev = (t_begin + i - 1) / P.tpi, img_cur = (t_begin + i) / P.tpi;
        if (img_prev != img_cur) flush(img_prev);
      }
      // pass 2: p = exp(k - m_tile) -> fp16 -> P[d][px]
      const float mb = m_tile * kLog2e;
      float zt = 0.f;
      uint8_t* prow = sP + (size_t)d * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(kaddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        uint32_t h[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const __half2 x = __floats2half2_rn(fmaf(__uint_as_float(v[2 * jj]), kLog2e, -mb),
                                              fmaf(__uint_as_float(v[2 * jj + 1]), kLog2e, -mb));
          h[jj] = ex2_h2(*reinterpret_cast<const uint32_t*>(&x));
          const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&h[jj]));
          zt += pf.x + pf.y;
        }
        uint8_t* blk = prow + (size_t)(c >> 1) * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(blk + ((((c & 1) * 4 + q) ^ (d & 7)) << 4)) =
              make_uint4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
      }
      tc_fence_before();
      mbar_arrive(&ctl->k_empty[b]);
      fence_proxy_async();
      mbar_arrive(&ctl->pv_ready);
      m_prev = m_tile;
      z_prev = zt;
    }
    if (ntiles > 0) {
      epi2(ntiles - 1);
      flush((t_end - 1) / P.tpi);
    }
    tc_fence_before();
  } else {
    // =============================== v-warps ====================================
    const int quarter = warp & 3;
    const int e = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint8_t* vrow = sVt + (size_t)e * 128;
    for (int i = 0; i < ntiles; ++i) {
      const int b = i & 1;
      mbar_wait(&ctl->v_full[b], (uint32_t)(i >> 1) & 1u);
      if (i > 0) mbar_wait(&ctl->d2_full[(i - 1) & 1], (uint32_t)((i - 1) >> 1) & 1u);   // Vt free
      tc_fence_after();
      const uint32_t vaddr = v_cols(b) + lane_off;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(vaddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        uint32_t h[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const __half2 x = __floats2half2_rn(__uint_as_float(v[2 * jj]), __uint_as_float(v[2 * jj + 1]));
          h[jj] = *reinterpret_cast<const uint32_t*>(&x);
        }
        uint8_t* blk = vrow + (size_t)(c >> 1) * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(blk + ((((c & 1) * 4 + q) ^ (e & 7)) << 4)) =
              make_uint4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&ctl->pv_ready);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(taddr, 512);
  }
}

// Combine the per-CTA partials of one image (fixed order) and fold the normalised context into
// the to_out weight.  grid (B, 4): blockIdx.y takes a quarter of the output channels.
__global__ void __launch_bounds__(256)
k_linattn_fold(const float* __restrict__ partials, const float* __restrict__ wout,
               __half* __restrict__ weff, int C, int tpi, int total_tiles, int grid_kv, int max_slots,
               float inv_n) {
  __shared__ float sM[128];
  __shared__ float sZ[128];
  __shared__ float sC[128 * 33];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int first = range_owner(b * tpi, total_tiles, grid_kv);
  int nslots = 0;
  for (int k = 0; k < max_slots; ++k) {
    const int c = first + k;
    if (c < grid_kv && range_begin(c, total_tiles, grid_kv) < (b + 1) * tpi) nslots = k + 1;
  }
  const float* base = partials + (size_t)b * max_slots * kPartialFloats;
  if (tid < 128) {
    float m = -INFINITY;
    for (int k = 0; k < nslots; ++k) m = fmaxf(m, base[(size_t)k * kPartialFloats + tid]);
    float zz = 0.f;
    for (int k = 0; k < nslots; ++k)
      zz = fmaf(base[(size_t)k * kPartialFloats + 128 + tid],
                exp2f((base[(size_t)k * kPartialFloats + tid] - m) * kLog2e), zz);
    sM[tid] = m;
    sZ[tid] = zz;
  }
  __syncthreads();
  for (int i = tid; i < 4096; i += 256) {
    const int dd = i >> 5;
    float a = 0.f;
    for (int k = 0; k < nslots; ++k)
      a = fmaf(base[(size_t)k * kPartialFloats + 256 + i],
               exp2f((base[(size_t)k * kPartialFloats + dd] - sM[dd]) * kLog2e), a);
    sC[dd * 33 + (i & 31)] = a;
  }
  __syncthreads();
  const int hd = tid & 127, h = hd >> 5;
  const float norm = inv_n / sZ[hd];
  const int cq = C >> 2;
  for (int c = blockIdx.y * cq + (tid >> 7); c < (blockIdx.y + 1) * cq; c += 2) {
    const float* w = wout + (size_t)c * 128 + h * 32;
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) a = fmaf(__ldg(w + e), sC[hd * 33 + e], a);
    weff[((size_t)b * C + c) * 128 + hd] = __float2half_rn(a * norm);
  }
}

struct KvCtxLaunch {
  CUtensorMap tmX, tmW;
  Params P;
  int smem, maxB, tpi;
};

}  // namespace

KvCtxOp::KvCtxOp() : impl(nullptr) {}
KvCtxOp::~KvCtxOp() { delete reinterpret_cast<KvCtxLaunch*>(impl); }
KvCtxOp::KvCtxOp(const KvCtxOp& o) : impl(nullptr) {
  if (o.impl) impl = new KvCtxLaunch(*reinterpret_cast<KvCtxLaunch*>(o.impl));
}
KvCtxOp& KvCtxOp::operator=(const KvCtxOp& o) {
  if (this != &o) {
    delete reinterpret_cast<KvCtxLaunch*>(impl);
    impl = o.impl ? new KvCtxLaunch(*reinterpret_cast<KvCtxLaunch*>(o.impl)) : nullptr;
  }
  return *this;
}

int kvctx_max_slots(int maxB) {
  // an image's tiles are spread over at most ceil(grid / B) + 1 contiguous CTA ranges
  return (num_sms() + maxB - 1) / maxB + 2;
}

int kvctx_plan(KvCtxOp* op, int maxB, const __half* xn, int H, int W, int C, int pix_stride,
               const __half* wqkv, float* partials) {
  if (C % 64 != 0 || (H * W) % 128 != 0) {
    set_error("kvctx_plan: unsupported shape %dx%d C=%d", H, W, C);
    return PRG_ERR_ARG;
  }
  const int tile_w = W < 128 ? W : 128;
  if ((tile_w & (tile_w - 1)) != 0 || tile_w < 8 || W % tile_w != 0 || H % (128 / tile_w) != 0) {
    set_error("kvctx_plan: unsupported spatial size %dx%d", H, W);
    return PRG_ERR_ARG;
  }
  KvCtxLaunch* L = new KvCtxLaunch();
  memset(L, 0, sizeof(*L));
  const int tile_h = 128 / tile_w;
  int l2 = 0;
  while ((1 << l2) < tile_w) ++l2;
  Params& P = L->P;
  P.tile_w_log2 = l2;
  P.tiles_x = W / tile_w;
  P.tpi = (H * W) / 128;
  P.num_kb = C / 64;
  P.max_slots = kvctx_max_slots(maxB);
  P.partials = partials;
  int stages = (kSmemBudget - 2 * kPvBytes - 2048) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = stages;
  L->smem = stages * kStageBytes + 2 * kPvBytes + 1024 + (int)sizeof(Ctl) + 64;
  L->maxB = maxB;
  L->tpi = P.tpi;
  int rc;
  {
    const uint64_t ps = (uint64_t)pix_stride * 2;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)maxB};
    uint64_t str[3] = {ps, ps * W, ps * W * H};
    uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, 1};
    rc = tmap_encode_f16(&L->tmX, xn, 4, dims, str, box);
  }
  if (rc == PRG_OK) {
    uint64_t dims[3] = {(uint64_t)C, 384, 1};
    uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)C * 2 * 384};
    uint32_t box[3] = {64, 256, 1};
    rc = tmap_encode_f16(&L->tmW, wqkv, 3, dims, str, box);
  }
  if (rc != PRG_OK) {
    delete L;
    return rc;
  }
  delete reinterpret_cast<KvCtxLaunch*>(op->impl);
  op->impl = L;
  return PRG_OK;
}

// k/v projection + context partials for the first B images, then W_eff.
int kvctx_run(KvCtxOp& op, int B, const float* wout, __half* weff, int C, cudaStream_t s) {
  KvCtxLaunch L = *reinterpret_cast<KvCtxLaunch*>(op.impl);
  Params& P = L.P;
  P.total_tiles = B * L.tpi;
  // The partial slots were sized for maxB images over all SMs; with fewer images use fewer CTAs,
  // so that an image never spans more contiguous CTA ranges than it has slots.
  const int grid = std::max(1, std::min(std::min(P.total_tiles, num_sms()), (P.max_slots - 2) * B));
  static int configured = 0;
  if (!configured) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_kvctx, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = 1;
  }
  k_kvctx<<<grid, kThreads, L.smem, s>>>(L.tmX, L.tmW, P);
  PRG_LAUNCH_CHECK();
  dim3 g(B, 4);
  k_linattn_fold<<<g, 256, 0, s>>>(P.partials, wout, weff, C, L.tpi, P.total_tiles, grid, P.max_slots,
                                   1.f / (float)(L.tpi * 128));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

}  // namespace prg

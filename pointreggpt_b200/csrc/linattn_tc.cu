// LinearAttention, fused k/v projection + context (tcgen05, sm_100a).
//
// Replaces, for  k = softmax_n(W_k xn),  v = W_v xn / n,  ctx = k v^T  (SDD:750-761), the k and v
// thirds of the to_qkv GEMM *and* the context reduction: k and v never reach HBM.
//
// Per 128-pixel tile of xn (NHWC fp16, TMA):
//   MMA1   D_k[d][px] = W_k[d][:] . xn[px][:]   and   D_v[e][px] = W_v[e][:] . xn[px][:]
//          (weights are the M operand, pixels the N operand: TMEM lane = channel, column = pixel,
//          so everything a softmax over pixels needs is local to one thread)
//   eight epilogue warps = 4 TMEM lane quarters (32 channels = one head) x 2 pixel halves; the
//   thread of (channel c, half hf) handles the 64 pixels of its half for BOTH k row c and v row c:
//          max m over its 64 pixels, p = exp(k - m) as fp16 -> P[c][px] (shared memory, K-major,
//          128B swizzle), z = sum p;  v as fp16 -> Vt[c][px]
//   MMA2   D2_hf[d][e] = sum_{px in half hf} P[d][px] Vt[e][px]   (M = N = 128, K = 64 per half;
//          the four 32x32 diagonal blocks are the per-head contexts)
//   each thread keeps a flash-style running (m, z, ctx[32]) for its (channel, half) pixel stream,
//   rescaled when the running max moves.
// An image is cut into fixed chunks of tiles (a function of the image size only); a CTA owns a
// contiguous range of chunks and writes two partials (one per pixel half; m[128], z[128],
// ctx[128][32]) per chunk, so
// the numbers do not depend on the batch size or the grid.  k_linattn_fold combines the partials
// of an image in a fixed order (bit-reproducible, no atomics) and folds the normalised context
// into the to_out weight:
//   W_eff[b][c][h*32+d] = sum_e W_out[c][h*32+e] ctx_b[h][d][e] / (z_b[h*32+d] n)      (SDD:763-769)
#include <algorithm>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "attention.cuh"
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace prg {

using namespace ptx;

namespace {

constexpr int kThreads = 320;           // TMA warp, MMA warp, 8 epilogue warps
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
  int tpc, total_chunks;           // tiles per chunk, chunks of this launch (B * tpi / tpc)
  int num_kb, stages;
  float* partials;                 // [total_chunks][2 halves][kPartialFloats]
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

__device__ __forceinline__ void tma_store_4d_(const CUtensorMap* m, const void* src, int c0, int c1, int c2,
                                              int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0_() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0_() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// first tile of CTA c when `total` tiles are dealt in contiguous ranges to `grid` CTAs
__host__ __device__ inline int range_begin(int c, int total, int grid) {
  return (int)(((long long)c * total) / grid);
}
__global__ void __launch_bounds__(kThreads, 1)
k_kvctx(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
        const Params P) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sStage = smem;
  uint8_t* sP = sStage + (size_t)P.stages * kStageBytes;
  uint8_t* sVt = sP + kPvBytes;
  Ctl* ctl = reinterpret_cast<Ctl*>(sVt + kPvBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_h = 128 >> P.tile_w_log2;
  const int t_begin = range_begin(blockIdx.x, P.total_chunks, gridDim.x) * P.tpc;
  const int t_end = range_begin(blockIdx.x + 1, P.total_chunks, gridDim.x) * P.tpc;
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
  pdl_wait();                 // everything above touched no data of a preceding kernel
  const uint32_t taddr = ctl->tmem_addr;
  // TMEM columns: [K0 | V0 | K1 | V1], 128 each; D2 of tile i, half 0 / 1 reuses K(i & 1) / V(i & 1)
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
          const uint32_t d2 = taddr_u + (uint32_t)(bj * 256);   // half 0 -> K(bj), half 1 -> V(bj)
          const uint64_t dp = desc_hi | (uint64_t)(sP_u >> 4), dvt = desc_hi | (uint64_t)(sVt_u >> 4);
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const uint64_t o = (uint64_t)((s >> 2) * (16384 >> 4) + (s & 3) * 2);
            umma_f16(d2 + (uint32_t)((s >> 2) * 128), dp + o, dvt + o, idesc, (s & 3) ? 1u : 0u);
          }
          umma_commit(&ctl->d2_full[bj]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue warps ==============================
    const int quarter = warp & 3;            // TMEM lane quarter = head
    const int hf = (warp - 2) >> 2;          // pixel half of the tile this warp owns
    const int d = quarter * 32 + lane;       // k channel and v channel of this thread
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m_run = -INFINITY, z = 0.f;
    float acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0.f;
    float m_prev = 0.f, z_prev = 0.f;        // max / sum of the tile whose D2 is pending
    uint8_t* prow = sP + (size_t)hf * 16384 + (size_t)d * 128;    // K block hf, row d
    uint8_t* vrow = sVt + (size_t)hf * 16384 + (size_t)d * 128;

    auto epi2 = [&](int j) {                 // fold tile j's D2 (this half) into the running state
      const int bj = j & 1;
      mbar_wait(&ctl->d2_full[bj], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32((hf ? v_cols(bj) : k_cols(bj)) + lane_off + (uint32_t)(quarter * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(hf ? &ctl->v_empty[bj] : &ctl->k_empty[bj]);
      const float m_new = fmaxf(m_run, m_prev);
      const float so = ex2_f((m_run - m_new) * kLog2e), sn = ex2_f((m_prev - m_new) * kLog2e);
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = fmaf(acc[e], so, __uint_as_float(v[e]) * sn);
      z = fmaf(z, so, z_prev * sn);
      m_run = m_new;
    };
    auto flush = [&](int chunk) {
      float* dst = P.partials + ((size_t)chunk * 2 + hf) * kPartialFloats;
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
      const uint32_t kaddr = k_cols(b) + lane_off + (uint32_t)(hf * 64);
      // pass 1: this channel's maximum over the 64 pixels of this half
      float m_tile = -INFINITY;
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(kaddr, v0);
        tmem_ld32(kaddr + 32, v1);
        tmem_ld_wait();
        float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
          mp[(jj >> 1) & 1] = fmaxf(mp[(jj >> 1) & 1], fmaxf(__uint_as_float(v0[jj]), __uint_as_float(v0[jj + 1])));
          mp[2 + ((jj >> 1) & 1)] = fmaxf(mp[2 + ((jj >> 1) & 1)], fmaxf(__uint_as_float(v1[jj]), __uint_as_float(v1[jj + 1])));
        }
        m_tile = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
      }
      // tile i-1: its P / Vt have been consumed once D2 is complete
      if (i > 0) {
        epi2(i - 1);
        if (i % P.tpc == 0) flush((t_begin + i - 1) / P.tpc);   // chunk boundary
      }
      // pass 2: p = exp(k - m_tile) -> fp16 -> P[d][px]
      const float mb = m_tile * kLog2e;
      float zt;
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(kaddr, v0);
        tmem_ld32(kaddr + 32, v1);
        tmem_ld_wait();
        uint32_t h0[16], h1[16];
        float zp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const __half2 x0 = __floats2half2_rn(fmaf(__uint_as_float(v0[2 * jj]), kLog2e, -mb),
                                               fmaf(__uint_as_float(v0[2 * jj + 1]), kLog2e, -mb));
          const __half2 x1 = __floats2half2_rn(fmaf(__uint_as_float(v1[2 * jj]), kLog2e, -mb),
                                               fmaf(__uint_as_float(v1[2 * jj + 1]), kLog2e, -mb));
          h0[jj] = ex2_h2(*reinterpret_cast<const uint32_t*>(&x0));
          h1[jj] = ex2_h2(*reinterpret_cast<const uint32_t*>(&x1));
          const float2 pf0 = __half22float2(*reinterpret_cast<const __half2*>(&h0[jj]));
          const float2 pf1 = __half22float2(*reinterpret_cast<const __half2*>(&h1[jj]));
          zp[jj & 1] += pf0.x + pf0.y;
          zp[2 + (jj & 1)] += pf1.x + pf1.y;
        }
        zt = (zp[0] + zp[1]) + (zp[2] + zp[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(prow + ((q ^ (d & 7)) << 4)) =
              make_uint4(h0[q * 4], h0[q * 4 + 1], h0[q * 4 + 2], h0[q * 4 + 3]);
          *reinterpret_cast<uint4*>(prow + (((4 + q) ^ (d & 7)) << 4)) =
              make_uint4(h1[q * 4], h1[q * 4 + 1], h1[q * 4 + 2], h1[q * 4 + 3]);
        }
      }
      // v row of the same channel / half -> Vt
      mbar_wait(&ctl->v_full[b], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const uint32_t vaddr = v_cols(b) + lane_off + (uint32_t)(hf * 64);
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(vaddr, v0);
        tmem_ld32(vaddr + 32, v1);
        tmem_ld_wait();
        uint32_t h0[16], h1[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const __half2 x0 = __floats2half2_rn(__uint_as_float(v0[2 * jj]), __uint_as_float(v0[2 * jj + 1]));
          const __half2 x1 = __floats2half2_rn(__uint_as_float(v1[2 * jj]), __uint_as_float(v1[2 * jj + 1]));
          h0[jj] = *reinterpret_cast<const uint32_t*>(&x0);
          h1[jj] = *reinterpret_cast<const uint32_t*>(&x1);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(vrow + ((q ^ (d & 7)) << 4)) =
              make_uint4(h0[q * 4], h0[q * 4 + 1], h0[q * 4 + 2], h0[q * 4 + 3]);
          *reinterpret_cast<uint4*>(vrow + (((4 + q) ^ (d & 7)) << 4)) =
              make_uint4(h1[q * 4], h1[q * 4 + 1], h1[q * 4 + 2], h1[q * 4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&ctl->pv_ready);
      m_prev = m_tile;
      z_prev = zt;
    }
    if (ntiles > 0) {
      epi2(ntiles - 1);
      flush((t_end - 1) / P.tpc);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(taddr, 512);
  }
}

// Combine the chunk partials of one image (fixed order) and fold the normalised context into
// the to_out weight.  grid (B, 4 heads): one CTA per (image, head); up to 32 chunks.
constexpr int kMaxChunks = 64;      // partial slots per image (two pixel halves per chunk)
__global__ void __launch_bounds__(256)
k_linattn_fold(const float* __restrict__ partials, const float* __restrict__ wout,
               __half* __restrict__ weff, int C, int cpi, float inv_n) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sS[kMaxChunks][32];   // exp(m_k - m) per chunk and channel of this head
  __shared__ float sN[32];               // 1 / (z n)
  __shared__ float sC[32 * 33];          // combined context [d][e]
  __shared__ float sW[256 * 33];         // W_out[:, h*32 .. h*32+31] (C <= 256 rows; larger C loops)
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const float* base = partials + (size_t)b * cpi * kPartialFloats;
  if (tid < 32) {
    const int hd = h * 32 + tid;
    float mk[kMaxChunks];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
      mk[k] = (k < cpi) ? base[(size_t)k * kPartialFloats + hd] : -INFINITY;
      m = fmaxf(m, mk[k]);
    }
    float zz = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
      const float sc = exp2f((mk[k] - m) * kLog2e);   // 0 for the unused slots
      sS[k][tid] = sc;
      if (k < cpi) zz = fmaf(base[(size_t)k * kPartialFloats + 128 + hd], sc, zz);
    }
    sN[tid] = inv_n / zz;
  }
  __syncthreads();
  for (int i = tid; i < 1024; i += 256) {
    const int dd = i >> 5;
    const float* src = base + 256 + (h * 32) * 32 + i;
    float v[kMaxChunks];
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) v[k] = (k < cpi) ? src[(size_t)k * kPartialFloats] : 0.f;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) a = fmaf(v[k], sS[k][dd], a);
    sC[dd * 33 + (i & 31)] = a;
  }
  __syncthreads();
  // this head's slice of the to_out weight, staged (coalesced) in blocks of 256 output channels
  // instead of being re-read from global memory by every warp in a latency-bound loop.  The blocks of
  // 64 output channels are spread over blockIdx.z (each CTA re-derives the small combined context): at
  // batch 4 the kernel otherwise runs on 16 CTAs and its latency is a sixth of the attention's time.
  const int dd = tid & 31;
  const float norm = sN[dd];
  const int c_lo = blockIdx.z * 64, c_hi = min(C, c_lo + 64);
  for (int cb = c_lo; cb < c_hi; cb += 256) {
    const int nc = min(256, c_hi - cb);
    __syncthreads();
    for (int i = tid; i < nc * 32; i += 256) {
      const int c = i >> 5, e = i & 31;
      sW[c * 33 + e] = __ldg(wout + (size_t)(cb + c) * 128 + h * 32 + e);
    }
    __syncthreads();
    for (int c = tid >> 5; c < nc; c += 8) {
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) a = fmaf(sW[c * 33 + e], sC[dd * 33 + e], a);
      weff[((size_t)b * C + cb + c) * 128 + h * 32 + dd] = __float2half_rn(a * norm);
    }
  }
}

// ------------------------------------------------------------------------------------------
// LinearAttention, fused q projection + softmax_d + (context . q) + to_out + LayerNorm + residual
// (SDD:750-756, 763-769, Residual SDD:583-589): q never reaches HBM.
//
// Per 128-pixel tile:  MMA1  D_q[px][128] = xn[px][:] . W_q^T          (TMEM, double-buffered)
//   q-warps (thread = pixel): softmax over each 32-channel head * scale -> fp16 Q tile in
//          shared memory (K-major, 128B swizzle) = the M operand of
//   MMA2   D_o[px][C] = Q[px][:] . W_eff[b]^T    (W_eff = to_out weight folded with the image's
//          normalised context, one [C][128] fp16 matrix per image, resident in shared memory)
//   o-warps (thread = pixel): + bias, channel LayerNorm * g, + residual x -> fp16 -> per-warp
//          staging slab -> TMA store.
// ------------------------------------------------------------------------------------------
struct alignas(16) QCtl {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t dq_full[2], dq_empty[2];
  uint64_t q_ready[2], q_free[2], do_full[2], do_empty[2];
  uint64_t weff_full, weff_free;
  uint32_t tmem_addr, pad;
};

struct QParams {
  int tile_w_log2, tiles_x, tpi, total_tiles, num_kb, stages, H, W;
  const float *bias, *gain;
  const __half* res;          // residual x, NHWC, C channels dense
  float q_scale;
  long long* trace;           // debug: clock64 stamps of CTA 0 (nullptr = off), 16 per tile
};

constexpr int kQStage = 128 * 128 + 128 * 128;   // xn K block + W_q K block

// C <= 128: the Q tile and the D_o accumulator are double-buffered (shared memory and TMEM allow
// it), so softmax (tile i+1), MMA2 (tile i) and the LayerNorm/store (tile i-1) overlap.
template <int C>
__global__ void __launch_bounds__(kThreads, 1)
k_qout(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
       const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmO,
       const QParams P) {
  // per-tile clock64 timeline of CTA 0 (tools/trace_qout.py): compiled in only with -DPRG_QOUT_TRACE_BUILD
#ifdef PRG_QOUT_TRACE_BUILD
  long long* const p_trace = P.trace;
#else
  constexpr long long* p_trace = nullptr;
#endif
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int NB = (C <= 128) ? 2 : 1;         // Q tiles / D_o accumulators
  constexpr int kWeffBytes = C * 256;            // two K blocks of [C rows][64]
  constexpr int kSlabBytes = 32 * C * 2;         // one o-warp: 32 pixels x C channels
  uint8_t* sStage = smem;
  uint8_t* sQ = sStage + (size_t)P.stages * kQStage;   // NB x 2 K blocks x [128 px][64 ch]
  uint8_t* sE = sQ + NB * 32768;
  uint8_t* sO = sE + kWeffBytes;
  QCtl* ctl = reinterpret_cast<QCtl*>(sO + 4 * kSlabBytes);
  // bias / gain: shared-memory copies when they fit (C <= 128), else read through L1
  float* sxch = reinterpret_cast<float*>(ctl + 1);    // LayerNorm partial sums: [4 quarters][192]
  float* scoef = sxch + 768;
  const float* sbias = (C <= 128) ? scoef : P.bias;
  const float* sgain = (C <= 128) ? scoef + C : P.gain;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_w = 1 << P.tile_w_log2, tile_h = 128 >> P.tile_w_log2;
  if (C <= 128)
    for (int i = threadIdx.x; i < C; i += kThreads) {
      scoef[i] = __ldg(P.bias + i);
      scoef[C + i] = __ldg(P.gain + i);
    }
  const int t_begin = range_begin(blockIdx.x, P.total_tiles, gridDim.x);
  const int t_end = range_begin(blockIdx.x + 1, P.total_tiles, gridDim.x);
  const int ntiles = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmE);
    prefetch_tmap(&tmO);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctl->dq_full[b], 1);
      mbar_init(&ctl->dq_empty[b], 256);
      mbar_init(&ctl->q_ready[b], 256);
      mbar_init(&ctl->q_free[b], 1);
      mbar_init(&ctl->do_full[b], 1);
      mbar_init(&ctl->do_empty[b], 256);
    }
    mbar_init(&ctl->weff_full, 1);
    mbar_init(&ctl->weff_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                 // everything above touched no data of a preceding kernel (bias / gain are weights)
  const uint32_t taddr = ctl->tmem_addr;     // columns: [Dq0 | Dq1 | Do0 (C) | Do1 (C)]

  auto tile_xy = [&](int t, int& img, int& x0, int& y0) {
    img = t / P.tpi;
    const int r = t - img * P.tpi;
    const int tyi = r / P.tiles_x, txi = r - tyi * P.tiles_x;
    x0 = txi << P.tile_w_log2;
    y0 = tyi * tile_h;
  };
  // buffer index / wait parity of the n-th use of an NB-deep ring
  auto ring_b = [](int n) { return NB == 2 ? (n & 1) : 0; };
  auto ring_par = [](int n) { return (uint32_t)(NB == 2 ? (n >> 1) : n) & 1u; };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < ntiles; ++i) {
        int img, x0, y0;
        tile_xy(t_begin + i, img, x0, y0);
        for (int kb = 0; kb < P.num_kb; ++kb) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          uint8_t* dst = sStage + (size_t)stage * kQStage;
          mbar_arrive_expect_tx(&ctl->full[stage], kQStage);
          tma_load_4d(&tmX, &ctl->full[stage], dst, kb * 64, x0, y0, img);
          tma_load_3d(&tmW, &ctl->full[stage], dst + 16384, kb * 64, 0, 0);   // rows 0..127 = q
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    constexpr uint32_t idesc1 = idesc_f16(128, 128);
    constexpr uint32_t idesc2 = idesc_f16(128, C);
    const uint32_t taddr_u = __reduce_or_sync(0xffffffffu, taddr);
    const uint32_t sStage_u = __reduce_or_sync(0xffffffffu, smem_u32(sStage));
    const uint32_t sQ_u = __reduce_or_sync(0xffffffffu, smem_u32(sQ));
    const uint32_t sE_u = __reduce_or_sync(0xffffffffu, smem_u32(sE));
    const uint64_t desc_hi = smem_desc_sw128(0);
    int stage = 0;
    uint32_t phase = 0;
    int cur_img = -1;
    uint32_t n_weff = 0;      // W_eff loads so far (phase of weff_full / weff_free)
    for (int i = 0; i <= ntiles; ++i) {
      if (i < ntiles) {
        const int b = i & 1;
        mbar_wait(&ctl->dq_empty[b], ((uint32_t)(i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (p_trace != nullptr && blockIdx.x == 0 && lane == 0 && i < 48) p_trace[i * 16 + 5] = clock64();
        const uint32_t dq = taddr_u + (uint32_t)(b * 128);
        for (int kb = 0; kb < P.num_kb; ++kb) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t x_lo = (sStage_u + (uint32_t)stage * kQStage) >> 4;
          if (elect_one()) {
            const uint64_t dx = desc_hi | (uint64_t)x_lo;
            const uint64_t dw = desc_hi | (uint64_t)(x_lo + (16384 >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(dq, dx + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&ctl->empty[stage]);
            if (kb == P.num_kb - 1) umma_commit(&ctl->dq_full[b]);
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (i >= 1) {
        const int j = i - 1;
        const int img = (t_begin + j) / P.tpi;
        if (img != cur_img) {
          // new image: its W_eff replaces the resident one once every MMA2 that read it is done
          if (cur_img >= 0) {
            if (elect_one()) umma_commit(&ctl->weff_free);
            __syncwarp();
            mbar_wait(&ctl->weff_free, (n_weff - 1u) & 1u);
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&ctl->weff_full, kWeffBytes);
            tma_load_3d(&tmE, &ctl->weff_full, sE, 0, 0, img);
            tma_load_3d(&tmE, &ctl->weff_full, sE + C * 128, 64, 0, img);
          }
          __syncwarp();
          mbar_wait(&ctl->weff_full, n_weff & 1u);
          ++n_weff;
          cur_img = img;
        }
        const int bj = ring_b(j);
        if (p_trace != nullptr && blockIdx.x == 0 && lane == 0 && j < 48) p_trace[j * 16 + 6] = clock64();
        mbar_wait(&ctl->q_ready[bj], ring_par(j));
        mbar_wait(&ctl->do_empty[bj], ring_par(j) ^ 1u);
        tc_fence_after();
        if (p_trace != nullptr && blockIdx.x == 0 && lane == 0 && j < 48) p_trace[j * 16 + 7] = clock64();
        if (elect_one()) {
          const uint32_t dout = taddr_u + 256u + (uint32_t)(bj * C);
          const uint64_t dq_ = desc_hi | (uint64_t)((sQ_u + (uint32_t)bj * 32768u) >> 4);
          const uint64_t de = desc_hi | (uint64_t)(sE_u >> 4);
#pragma unroll
          for (int s = 0; s < 8; ++s)
            umma_f16(dout, dq_ + (uint64_t)((s >> 2) * (16384 >> 4) + (s & 3) * 2),
                     de + (uint64_t)((s >> 2) * ((C * 128) >> 4) + (s & 3) * 2), idesc2, s > 0 ? 1u : 0u);
          umma_commit(&ctl->q_free[bj]);
          umma_commit(&ctl->do_full[bj]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue warps ==============================
    // Symmetric roles: warp (quarter, hh) owns the 32 pixels of its TMEM lane quarter and, of those,
    //   S(i):   the softmax of heads 2hh, 2hh+1 (D_q columns [64 hh, 64 hh + 64)) -> K block hh of Q
    //   O(i-1): the LayerNorm / residual / store of output channels [hh C/2, (hh+1) C/2)
    // so every scheduler holds two warps whose MUFU- and FMA-bound phases overlap.  The two warps of
    // a quarter exchange their LayerNorm partial sums through shared memory (named barrier 1+quarter).
    const int quarter = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int tyr = row >> P.tile_w_log2, txr = row & (tile_w - 1);
    constexpr int CH = C / 2;                      // output channels per warp
    uint8_t* slab = sO + (size_t)quarter * kSlabBytes;
    float* xch = sxch + quarter * 192;             // [2 hh][sum, centred squares, shift][32 lanes]
    const float qs = P.q_scale;
    auto pair_bar = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };

    auto softmax_tile = [&](int i) {
      const int b = i & 1;
      const bool tr = p_trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 48;
      if (tr) p_trace[i * 16 + 0] = clock64();
      mbar_wait(&ctl->dq_full[b], (uint32_t)(i >> 1) & 1u);
      if (tr) p_trace[i * 16 + 1] = clock64();
      const int bq = ring_b(i);
      if (i >= NB) mbar_wait(&ctl->q_free[bq], ring_par(i - NB));   // MMA2(i - NB) has read this Q tile
      tc_fence_after();
      if (tr) p_trace[i * 16 + 2] = clock64();
      uint8_t* qrow = sQ + (size_t)bq * 32768 + (size_t)hh * 16384 + (size_t)row * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {                // one head per 32-column chunk
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(b * 128 + hh * 64 + c * 32) + lane_off, v);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          mbar_arrive(&ctl->dq_empty[b]);
        }
        float m4[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) m4[r] = __uint_as_float(v[r]);
#pragma unroll
        for (int jj = 4; jj < 32; ++jj) m4[jj & 3] = fmaxf(m4[jj & 3], __uint_as_float(v[jj]));
        const float mb = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * kLog2e;
        // packed fp32x2 math (FFMA2 / FADD2 / FMUL2): one instruction per element pair
        const float2 l2 = make_float2(kLog2e, kLog2e), nmb = make_float2(-mb, -mb);
        float2 f2[16];
        float2 s2a = make_float2(0.f, 0.f), s2b = make_float2(0.f, 0.f);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(v[2 * jj]), __uint_as_float(v[2 * jj + 1])), l2, nmb);
          f2[jj] = make_float2(ex2_f(a.x), ex2_f(a.y));
          if (jj & 1) s2b = __fadd2_rn(s2b, f2[jj]);
          else s2a = __fadd2_rn(s2a, f2[jj]);
        }
        const float2 st = __fadd2_rn(s2a, s2b);
        const float inv = __fdividef(qs, st.x + st.y);
        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t o[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 y = __fmul2_rn(f2[q * 4 + jj], inv2);
            const __half2 x = __floats2half2_rn(y.x, y.y);
            o[jj] = *reinterpret_cast<const uint32_t*>(&x);
          }
          *reinterpret_cast<uint4*>(qrow + (((c * 4 + q) ^ (row & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      if (tr) p_trace[i * 16 + 3] = clock64();
      fence_proxy_async();
      mbar_arrive(&ctl->q_ready[bq]);
      if (tr) p_trace[i * 16 + 4] = clock64();
    };

    // residual row of tile i (this warp's channel half): fetched one softmax phase ahead of its use
    uint4 rv[4];
    int o_img = 0, o_x0 = 0, o_y0 = 0;         // coordinates of the tile whose output is pending
    const __half* rp = nullptr;
    auto prefetch_res = [&](int i) {
      tile_xy(t_begin + i, o_img, o_x0, o_y0);
      rp = P.res + (((size_t)o_img * P.H + (o_y0 + tyr)) * P.W + (o_x0 + txr)) * C + hh * CH;
      const uint4* r4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
      for (int q = 0; q < 4; ++q) rv[q] = __ldg(r4 + q);
    };

    auto output_tile = [&](int i) {
      const int img = o_img, x0 = o_x0, y0 = o_y0;
      const int bo = ring_b(i);
      const uint32_t dout = taddr + 256u + (uint32_t)(bo * C + hh * CH) + lane_off;
      const float* cb = sbias + hh * CH;
      const float* cg = sgain + hh * CH;
      const bool tr = p_trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 48;
      if (tr) p_trace[i * 16 + 8] = clock64();
      mbar_wait(&ctl->do_full[bo], ring_par(i));
      tc_fence_after();
      if (tr) p_trace[i * 16 + 9] = clock64();
      // one sweep: S = sum x and Q = sum (x - a)^2 over this warp's channels (x = acc + bias, a = its
      // first value), exchanged with the partner warp and combined exactly:
      //   sum_h (x - mean)^2 = Q_h - 2 (mean - a_h) (S_h - n a_h) + n (mean - a_h)^2
      float sh = 0.f;
      float2 pa = make_float2(0.f, 0.f), pb = pa, qa = pa, qb = pa;
#pragma unroll 1
      for (int c = 0; c < CH; c += 32) {
        uint32_t v[32];
        tmem_ld32(dout + (uint32_t)c, v);
        tmem_ld_wait();
        if (c == 0) sh = __uint_as_float(v[0]) + cb[0];
        const float2 nsh = make_float2(-sh, -sh);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(cb + c + q * 4);
          const float2 xa = __fadd2_rn(make_float2(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1])),
                                       make_float2(b4.x, b4.y));
          const float2 xb = __fadd2_rn(make_float2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])),
                                       make_float2(b4.z, b4.w));
          pa = __fadd2_rn(pa, xa);
          pb = __fadd2_rn(pb, xb);
          const float2 da = __fadd2_rn(xa, nsh), db = __fadd2_rn(xb, nsh);
          qa = __ffma2_rn(da, da, qa);
          qb = __ffma2_rn(db, db, qb);
        }
      }
      const float p4s = (pa.x + pa.y) + (pb.x + pb.y), q4s = (qa.x + qa.y) + (qb.x + qb.y);
      {
        float* mine = xch + hh * 96;
        mine[lane] = p4s;
        mine[32 + lane] = q4s;
        mine[64 + lane] = sh;
      }
      if (hh == 0 && lane == 0) bulk_wait_read0_();   // the slab's previous TMA store has read it
      pair_bar();
      if (tr) p_trace[i * 16 + 10] = clock64();
      float mean, rstd;
      {
        const float s0 = xch[lane], q0 = xch[32 + lane], a0 = xch[64 + lane];
        const float s1 = xch[96 + lane], q1 = xch[128 + lane], a1 = xch[160 + lane];
        mean = (s0 + s1) * (1.f / C);
        const float e0 = mean - a0, e1 = mean - a1;
        const float m2 = (q0 - 2.f * e0 * (s0 - CH * a0) + CH * e0 * e0) +
                         (q1 - 2.f * e1 * (s1 - CH * a1) + CH * e1 * e1);
        rstd = rsqrtf(fmaxf(m2, 0.f) * (1.f / C) + 1e-5f);
      }
      if (tr) p_trace[i * 16 + 11] = clock64();
      // normalise + gain + residual -> slab (this warp's channel half of each pixel row)
#pragma unroll 1
      for (int c = 0; c < CH; c += 32) {
        uint32_t v[32];
        tmem_ld32(dout + (uint32_t)c, v);
        tmem_ld_wait();
        if (c + 32 >= CH) {                  // last read of this accumulator
          tc_fence_before();
          mbar_arrive(&ctl->do_empty[bo]);
        }
        const int cglob = hh * CH + c;       // first output channel of this chunk
        uint8_t* blk = slab + (size_t)(cglob >> 6) * (32 * 128) + (size_t)lane * 128;
        const int ch0 = (cglob & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __half2* rh = reinterpret_cast<const __half2*>(&rv[q]);
          const float4 b0 = *reinterpret_cast<const float4*>(cb + c + q * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(cb + c + q * 8 + 4);
          const float4 g0 = *reinterpret_cast<const float4*>(cg + c + q * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(cg + c + q * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          uint32_t o[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 r2 = __half22float2(rh[jj]);
            const float2 bm = __fadd2_rn(make_float2(bb[jj * 2], bb[jj * 2 + 1]), make_float2(-mean, -mean));
            const float2 x2 = __fadd2_rn(make_float2(__uint_as_float(v[q * 8 + jj * 2]), __uint_as_float(v[q * 8 + jj * 2 + 1])), bm);
            const float2 u2 = __fmul2_rn(x2, make_float2(rstd, rstd));
            const float2 y2 = __ffma2_rn(u2, make_float2(gg[jj * 2], gg[jj * 2 + 1]), r2);
            const __half2 hv = __floats2half2_rn(y2.x, y2.y);
            o[jj] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          *reinterpret_cast<uint4*>(blk + (((ch0 + q) ^ (lane & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (c + 32 < CH) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rv[q] = __ldg(reinterpret_cast<const uint4*>(rp + c + 32) + q);
        }
      }
      if (tr) p_trace[i * 16 + 12] = clock64();
      fence_proxy_async();
      pair_bar();
      if (tr) p_trace[i * 16 + 13] = clock64();
      if (hh == 0 && lane == 0) {
        const int px = quarter * 32;
#pragma unroll
        for (int cbx = 0; cbx < C / 64; ++cbx)
          tma_store_4d_(&tmO, slab + (size_t)cbx * (32 * 128), cbx * 64, x0 + (px & (tile_w - 1)),
                        y0 + (px >> P.tile_w_log2), img);
        bulk_commit_();
      }
    };

    for (int i = 0; i < ntiles; ++i) {
      if (i > 0) prefetch_res(i - 1);
      softmax_tile(i);
      if (i > 0) output_tile(i - 1);
    }
    if (ntiles > 0) {
      prefetch_res(ntiles - 1);
      output_tile(ntiles - 1);
    }
    if (lane == 0) bulk_wait0_();
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(taddr, 512);
  }
}

struct QOutLaunch {
  CUtensorMap tmX, tmW, tmE, tmO;
  QParams P;
  int smem, C;
};

struct KvCtxLaunch {
  CUtensorMap tmX, tmW;
  Params P;
  int smem, maxB, cpi;
};

// tiles per chunk: a function of the image size only (batch-invariant results); about 32 chunks
// per image, at most 16 tiles each
int tiles_per_chunk(int tpi) {
  int t = tpi / 32;
  if (t < 1) t = 1;
  if (t > 16) t = 16;
  while (tpi % t != 0) --t;
  return t;
}

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

size_t kvctx_partial_floats(int maxB, int H, int W) {
  const int tpi = (H * W) / 128;
  return (size_t)maxB * (tpi / tiles_per_chunk(tpi)) * 2 * kPartialFloats;
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
  P.tpc = tiles_per_chunk(P.tpi);
  P.partials = partials;
  int stages = (kSmemBudget - 2 * kPvBytes - 2048) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = stages;
  L->smem = stages * kStageBytes + 2 * kPvBytes + 1024 + (int)sizeof(Ctl) + 64;
  L->maxB = maxB;
  L->cpi = P.tpi / P.tpc;
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
  P.total_chunks = B * L.cpi;
  const int grid = std::min(P.total_chunks, num_sms());
  static int configured = 0;
  if (!configured) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_kvctx, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = 1;
  }
  PRG_CUDA_OK(launch_pdl(k_kvctx, dim3(grid), dim3(kThreads), L.smem, s, L.tmX, L.tmW, P));
  PRG_LAUNCH_CHECK();
  dim3 g(B, 4, (C + 63) / 64);
  PRG_CUDA_OK(launch_pdl(k_linattn_fold, g, dim3(256), 0, s, P.partials, wout, weff, C, 2 * L.cpi, 1.f / (float)(P.tpi * 128)));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

QOutOp::QOutOp() : impl(nullptr) {}
QOutOp::~QOutOp() { delete reinterpret_cast<QOutLaunch*>(impl); }
QOutOp::QOutOp(const QOutOp& o) : impl(nullptr) {
  if (o.impl) impl = new QOutLaunch(*reinterpret_cast<QOutLaunch*>(o.impl));
}
QOutOp& QOutOp::operator=(const QOutOp& o) {
  if (this != &o) {
    delete reinterpret_cast<QOutLaunch*>(impl);
    impl = o.impl ? new QOutLaunch(*reinterpret_cast<QOutLaunch*>(o.impl)) : nullptr;
  }
  return *this;
}

int qout_plan(QOutOp* op, int maxB, const __half* xn, int H, int W, int C, const __half* wqkv,
              const __half* weff, const float* bias, const float* gain, const __half* res, __half* out) {
  if ((C != 64 && C != 128 && C != 256) || (H * W) % 128 != 0) {
    set_error("qout_plan: unsupported shape %dx%d C=%d", H, W, C);
    return PRG_ERR_ARG;
  }
  const int tile_w = W < 128 ? W : 128;
  if ((tile_w & (tile_w - 1)) != 0 || tile_w < 8 || W % tile_w != 0 || H % (128 / tile_w) != 0) {
    set_error("qout_plan: unsupported spatial size %dx%d", H, W);
    return PRG_ERR_ARG;
  }
  QOutLaunch* L = new QOutLaunch();
  memset(L, 0, sizeof(*L));
  const int tile_h = 128 / tile_w;
  int l2 = 0;
  while ((1 << l2) < tile_w) ++l2;
  QParams& P = L->P;
  P.tile_w_log2 = l2;
  P.tiles_x = W / tile_w;
  P.tpi = (H * W) / 128;
  P.num_kb = C / 64;
  P.H = H;
  P.W = W;
  P.bias = bias;
  P.gain = gain;
  P.res = res;
  P.q_scale = 0.17677669529663687f;   // 32^-0.5 (SDD:742)
  const int fixed = (C <= 128 ? 2 : 1) * 32768 + 2 * C * 256 + 1024 + (int)sizeof(QCtl) + 3072 + (C <= 128 ? 8 * C : 0) + 64;
  int stages = (kSmemBudget - fixed) / kQStage;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) {   // C = 256 gets a single load stage (two small layers of the network)
    delete L;
    set_error("qout_plan: shared memory budget too small");
    return PRG_ERR_ARG;
  }
  P.stages = stages;
  L->smem = stages * kQStage + fixed;
  L->C = C;
  int rc;
  const uint64_t ps = (uint64_t)C * 2;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)maxB};
    uint64_t str[3] = {ps, ps * W, ps * W * H};
    uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, 1};
    rc = tmap_encode_f16(&L->tmX, xn, 4, dims, str, box);
    if (rc == PRG_OK) {
      uint32_t sbox[4] = {64, (uint32_t)std::min(tile_w, 32), 0, 1};
      sbox[2] = 32u / sbox[1];
      rc = tmap_encode_f16(&L->tmO, out, 4, dims, str, sbox);
    }
  }
  if (rc == PRG_OK) {
    uint64_t dims[3] = {(uint64_t)C, 384, 1};
    uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)C * 2 * 384};
    uint32_t box[3] = {64, 128, 1};
    rc = tmap_encode_f16(&L->tmW, wqkv, 3, dims, str, box);
  }
  if (rc == PRG_OK) {
    uint64_t dims[3] = {128, (uint64_t)C, (uint64_t)maxB};
    uint64_t str[2] = {256, (uint64_t)C * 256};
    uint32_t box[3] = {64, (uint32_t)C, 1};
    rc = tmap_encode_f16(&L->tmE, weff, 3, dims, str, box);
  }
  if (rc != PRG_OK) {
    delete L;
    return rc;
  }
  delete reinterpret_cast<QOutLaunch*>(op->impl);
  op->impl = L;
  return PRG_OK;
}

template <int C>
static int qout_launch(const QOutLaunch& L, int grid, cudaStream_t s) {
  static int configured = 0;
  if (!configured) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_qout<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = 1;
  }
  PRG_CUDA_OK(launch_pdl(k_qout<C>, dim3(grid), dim3(kThreads), L.smem, s, L.tmX, L.tmW, L.tmE, L.tmO, L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

int qout_run(QOutOp& op, int B, cudaStream_t s) {
  QOutLaunch L = *reinterpret_cast<QOutLaunch*>(op.impl);
  L.P.total_tiles = B * L.P.tpi;
  const int grid = std::min(L.P.total_tiles, num_sms());
  if (getenv("PRG_QOUT_TRACE") != nullptr && L.C == 64) {
    // debug timeline of CTA 0, warp 2 (synchronous): S: 0 top | 1 Dq ready | 2 Q tile free |
    // 3 softmax + Q written | 4 fenced + arrived;  MMA warp: 5 MMA1 go | 6 MMA2 loop top | 7 MMA2 go;
    // O: 8 top | 9 Do ready | 10 mean exchanged | 11 var exchanged | 12 slab written | 13 pair ready
    long long* d = nullptr;
    PRG_CUDA_OK(cudaMalloc(&d, 48 * 16 * sizeof(long long)));
    PRG_CUDA_OK(cudaMemset(d, 0, 48 * 16 * sizeof(long long)));
    L.P.trace = d;
    int rc = qout_launch<64>(L, grid, s);
    long long h[48 * 16];
    PRG_CUDA_OK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    const long long t0 = h[1];
    for (int t = 0; t < 32; ++t) {
      fprintf(stderr, "tile %2d:", t);
      for (int k = 0; k < 14; ++k) fprintf(stderr, " %6lld", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
    return rc;
  }
  if (L.C == 64) return qout_launch<64>(L, grid, s);
  if (L.C == 128) return qout_launch<128>(L, grid, s);
  return qout_launch<256>(L, grid, s);
}

}  // namespace prg

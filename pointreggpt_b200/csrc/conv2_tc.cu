// tcgen05 implicit-GEMM convolution / GEMM engine, persistent version (sm_100a).
//
// Replaces, for the U-Net denoiser and the depth-correction U-Net, the cuDNN/cuBLAS calls
// behind F.conv2d / nn.Conv2d at SDD:594-598, 615, 717, 743-745, 779-780 (same lines in DC).
//
// One CTA per SM walks a static list of work items.  Warp roles (192 threads):
//   warp 0   : TMA producer.  A operand, two modes:
//                per-tap : one 4-D/5-D box per (tap, 64-channel chunk) -- shifted window of the
//                          NHWC activation, zero padding = TMA out-of-bounds fill;
//                halo    : (3x3, stride 1, Cin = Cout = 64, row tiles) input rows stream through a
//                          ring, each loaded ONCE (130 pixels x 64 ch, 128B-swizzled).  One input
//                          row feeds the THREE output rows it touches in a single N = 192 MMA per
//                          (dx, k16): the weights sit in shared memory stacked [dy=2; dy=1; dy=0]
//                          per dx and the three output rows own neighbouring 64-column TMEM
//                          accumulators (a ring of eight).  dx = shifted view (start address +
//                          dx*128 B).  12-13 MMAs of N = 192 per output row instead of 36 of
//                          N = 64: the A operand is read from shared memory 3x less often
//                          (N = 64 MMAs are shared-memory-bandwidth bound: 6 KB per 32 cycles).
//              B operand (weights): resident in shared memory for the whole kernel when it
//              fits (loaded once per CTA), otherwise streamed next to A.
//   warp 1   : one lane issues tcgen05.mma (M=128, N=BN, K=16), accumulating in one of TWO
//              TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warps 2-9: epilogue (two warps per TMEM lane quarter, each draining half the columns): tcgen05.ld -> fused bias / GroupNorm statistics / softmax /
//              LayerNorm / residual -> fp16 -> 128B-swizzled staging tile -> TMA store.
#include <limits.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace prg {

using namespace ptx;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;     // 16 KiB per-tap A stage
constexpr int kHaloPix = kBlockM + 2;              // 130 pixels per ring row
constexpr int kHaloBytes = kHaloPix * 128;         // 16640 B written by TMA
constexpr int kHaloSlot = 17 * 1024;               // slot pitch (1024-aligned)
constexpr int kThreads = 320;                     // TMA warp + MMA warp + 8 epilogue warps
constexpr int kXfThreads = 192;                   // XF: + 6 transform warps (GroupNorm apply on the ring slot): 512 threads, 128 registers each
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 12;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kCtlBytes = 8192;
constexpr double kStatScale = 1048576.0;           // 2^20 fixed point (see conv_tc.cuh)

struct alignas(8) Ctl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t wfull;
  uint64_t xfull[kMaxStages];   // XF: ring slot transformed in place, ready for the MMA warp
  uint64_t tmem_full[8];
  uint64_t tmem_empty[8];
  uint32_t tmem_addr;
  uint32_t pad;
  float stats[4 * 32 * 2];  // EPI_GN: [TMEM quarter][8-column sub-block][sum, sumsq], written once per tile
  int colmax[128];          // EPI_QKV k tile
  float rowsum[2][128];     // EPI_LN_RES: per-row partial sums of the two column halves
  float bias[512];          // bias of every output channel (0 when absent)
  float gain[256];          // EPI_LN_RES: LayerNorm gain
};
static_assert(sizeof(Ctl) <= kCtlBytes, "control block too large");

__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read2() {
  asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct Item {
  int img, x0, y0, n_tile, cls;
};

}  // namespace

struct Conv2Params {
  ConvParams c;
  int halo, wres, n_tiles, stages;
  int pair;             // per-tap mode, BN = 128: an item is TWO adjacent M tiles sharing every B stage
  int dxs;              // pair mode, 3x3, rows of 128 pixels: the three dx taps of a (dy, K block) share ONE
                        // 130-pixel halo row per tile (shifted views, as in the row-streaming mode)
  int nsrc;             // halo mode: 1, or 2 = cat(x0, x1) of two 64-channel sources streamed row by row
  int cls_bind;         // halo mode, folded x2 upsample ("rows3"): CTA i works on parity class i & 3 only -- its
                        // class's weights (K columns cls * num_kb * 64 ...) stay resident, its input window is
                        // shifted by (py, px) and only the dx = 0, 1 taps are issued (see conv2_plan)
  int direct_store;     // BN = 64 epilogue writes global memory itself (no staging slabs / TMA store)
  int total_items;      // per-tap: m_tiles * n_tiles * classes ; halo: number of row segments
  int m_tiles;          // per-tap: tiles_x * tiles_y * B
  int rseg, segs_per_strip;
  int num_kb;           // K blocks per tile
  int a_slot;           // bytes per A stage / ring slot
  int w_bytes;          // resident weight bytes
  long long* trace;     // debug: per-tile clock64 stamps of CTA 0 (nullptr = off)
  int dbg_flags;        // debug: bit2 = skip the output TMA store, bit3 = skip GN flush
};

// CG = 2 (per-tap mode, BN = 256): the two CTAs of a cluster work as a tcgen05 CTA pair on M = 256
// (two adjacent M tiles): each CTA loads its own A tile and HALF of every B stage, the leader CTA
// issues tcgen05.mma.cta_group::2, and each CTA drains its own 128 accumulator rows.  Per SM a
// K block costs 32 KB of loads instead of 48 KB -- these layers sit at the ~64 B/clk/SM the TMA
// path sustains, not at the tensor-pipe limit.
// XF = 1 (row-streaming mode, one source): the input of the convolution is SiLU(A[b][c] * raw + B[b][c])
// -- the GroupNorm apply (+ scale / shift) of the producing Block (SDD:690-696) -- and is never
// materialised: four extra warps apply it IN PLACE to every ring slot between the TMA load and the MMAs
// (same arithmetic as k_gn_apply: fp32 affine on the packed pipe, ex2 + rcp SiLU, round to fp16), leave
// the zero padding (halo pixels and rows outside the image) untouched, then `fence.proxy.async` and
// hand the slot to the MMA warp through its own barrier.  A thread keeps one 8-channel group of
// coefficients in registers (the slot is 128B-swizzled: logical chunk g of pixel p sits at g ^ (p & 7)).
__device__ __forceinline__ float xf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float xf_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 xf_silu2(float2 t) {
  const float2 u = __fmul2_rn(t, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 d = __fadd2_rn(make_float2(xf_ex2(u.x), xf_ex2(u.y)), make_float2(1.f, 1.f));
  return __fmul2_rn(t, make_float2(xf_rcp(d.x), xf_rcp(d.y)));
}

// SiLU(t) = h + h tanh(h), h = t / 2: one MUFU operation per element instead of two (ex2 + rcp)
__device__ __forceinline__ float silu_tanh(float t) {
  const float h = 0.5f * t;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

template <int BN, int EPI, int CG = 1, int XF = 0, int R3 = 0>
__global__ void __launch_bounds__(kThreads + XF * kXfThreads, 1)
k_conv2(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
        const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO0,
        const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2,
        const __grid_constant__ CUtensorMap tmO3, const Conv2Params P) {
  pdl_trigger();
  const ConvParams& p = P.c;
  // per-tile clock64 timeline of CTA 0 (tools/trace_conv.py): compiled in only with -DPRG_CONV_TRACE_BUILD, so the
  // production kernels carry neither its tests nor its registers
#ifdef PRG_CONV_TRACE_BUILD
  long long* const p_trace = p_trace;
#else
  constexpr long long* p_trace = nullptr;
#endif
  // the input-transform kernel is always the single-source row-streaming form with resident weights and staged
  // stores: as compile-time constants these remove every other mode from that instantiation (it sits at its
  // register cap)
  // (R3 = 2: the two-source row-streaming conv with the launch shape fixed in the same way, rows3 code excluded)
  // (R3 = 3: the single-source row-streaming conv, likewise)
  const int p_halo = (XF || R3) ? 1 : P.halo, p_nsrc = (XF || R3 == 3) ? 1 : R3 ? 2 : P.nsrc, p_wres = (XF || R3) ? 1 : P.wres;
  const int p_pair = (XF || R3) ? 0 : P.pair, p_direct = (XF || R3 == 3) ? 0 : R3 ? 1 : P.direct_store;
  constexpr int kBBytes = (BN / CG) * kBlockK * 2;   // B rows this CTA stages per K block
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int item0 = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // per-tap item walk
  const int istep = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // row-streaming segment walk; with cls_bind the grid is four interleaved groups, one per parity class
  // (compile-time gates: the code of the two special modes exists only in the instantiations that can run them --
  // the input-transform kernel sits at its register cap and lost 15 % when unrelated paths grew)
  constexpr bool kRows3 = (R3 == 1);      // its own instantiation: step counts / weight strides stay compile-time
                                          // constants, so the MMA descriptors are base + constant (no dependent math)
  constexpr bool kDxs = (BN == 128 && CG == 1 && XF == 0);
  constexpr bool cls_bind = kRows3;
  // (R3 also fixes the launch shape: two-source row streaming, resident weights, direct stores)
  const int cls_b = cls_bind ? (int)(blockIdx.x & 3u) : 0;
  const int seg0 = cls_bind ? (int)(blockIdx.x >> 2) : (int)blockIdx.x;
  const int sstep = cls_bind ? (int)(gridDim.x >> 2) : (int)gridDim.x;
  // fp16 staging.  BN = 64: eight per-warp 4 KB slabs.  BN >= 128: ONE 64-channel box (16 KB) that
  // the epilogue fills and stores BN/64 times per tile -- the shared memory this frees buys a
  // fourth load stage, and these layers are bound by the bytes in flight, not by the epilogue.
  constexpr int kStageOut = (BN == 64) ? (EPI == EPI_GNRES ? 4 * kABytes : 2 * kABytes) : kABytes;
  constexpr int kOutBufs = 1;
  // TMEM: BN = 64 owns all 512 columns (eight accumulators, halo mode walks them as a ring),
  // the wider tiles two accumulators.
  constexpr int kTmemCols = 512;
  const uint32_t acc_mask = p_halo ? 7u : 1u;      // accumulator slots - 1
  const int acc_log2 = p_halo ? 3 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;                               // resident weights (may be empty)
  uint8_t* sA = sW + P.w_bytes;                     // A ring
  uint8_t* sB = sA + P.stages * P.a_slot;           // streamed B ring (when !wres)
  uint8_t* sO = sB + (p_wres ? 0 : P.stages * kBBytes);
  Ctl* ctl = reinterpret_cast<Ctl*>(sO + (p_direct ? 0 : kOutBufs * kStageOut));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_w = 1 << p.tile_w_log2;
  const int tile_h = kBlockM >> p.tile_w_log2;
  const int chunks = p.chunks0 + p.chunks1;
  const int num_kb = P.num_kb;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (p.chunks1 > 0) prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmO0);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    mbar_init(&ctl->wfull, 1);
    if (XF)
      for (int s = 0; s < P.stages; ++s) mbar_init(&ctl->xfull[s], 32);   // the one warp that transforms the row
    for (int a = 0; a < 8; ++a) {
      mbar_init(&ctl->tmem_full[a], 1);
      mbar_init(&ctl->tmem_empty[a], (BN == 64 ? kEpiThreads / 2 : kEpiThreads) * CG);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_2cta(&ctl->tmem_addr, kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(&ctl->tmem_addr, kTmemCols);
      tmem_relinquish();
    }
  }
  if (EPI == EPI_QKV && threadIdx.x < 128) ctl->colmax[threadIdx.x] = INT_MIN;
  for (int i = threadIdx.x; i < BN * P.n_tiles; i += blockDim.x)
    ctl->bias[i] = (p.bias != nullptr) ? __ldg(p.bias + i) : 0.f;
  if (EPI == EPI_LN_RES || (EPI == EPI_GNRES && p.has_ln_out))
    for (int i = threadIdx.x; i < BN; i += blockDim.x) ctl->gain[i] = __ldg(p.ln_g + i);
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  pdl_wait();                        // everything above touched weights only, no data of a preceding kernel
  const uint32_t taddr = ctl->tmem_addr;

  // item -> coordinates (per-tap mode)
  auto decode = [&](int item, int sub = 0) {
    Item it;
    it.n_tile = item % P.n_tiles;
    int m = item / P.n_tiles;
    const int msh = (CG == 2) ? 1 : p_pair;       // pair mode / CTA pair: item m covers tiles 2m, 2m + 1
    const int m_items = P.m_tiles >> msh;
    it.cls = m / m_items;
    m = ((m - it.cls * m_items) << msh) + sub;
    it.img = m / tiles_per_img;
    const int t_in = m - it.img * tiles_per_img;
    const int tyi = t_in / p.tiles_x, txi = t_in - tyi * p.tiles_x;
    it.x0 = txi << p.tile_w_log2;
    it.y0 = tyi * tile_h;
    return it;
  };
  // halo mode: every 128-pixel strip is cut into row segments that are dealt round-robin to
  // the CTAs (concurrently running CTAs work inside one small address window).
  auto decode_seg = [&](int seg, int& img, int& x0, int& y0, int& nr) {
    const int strips = p.tiles_x;
    const int per_img = strips * P.segs_per_strip;
    img = seg / per_img;
    const int r = seg - img * per_img;
    const int strip = r / P.segs_per_strip, sidx = r - strip * P.segs_per_strip;
    x0 = strip * kBlockM;
    y0 = sidx * P.rseg;
    nr = min(P.rseg, p.Ho - y0);
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      if (p_wres) {
        mbar_arrive_expect_tx(&ctl->wfull, (uint32_t)P.w_bytes);
        for (int nt = 0; nt < P.n_tiles; ++nt)
          for (int kb = 0; kb < num_kb; ++kb) {
            // halo mode: tap (dy, dx) goes to block dx*3 + (2 - dy), so that for one dx the taps
            // dy = 2, 1, 0 are one contiguous N = 192 B operand
            const int wtap = kb / chunks, wcc = kb - wtap * chunks;   // halo: per source, per dx, dy = 2, 1, 0
            // rows3: only the ky, kx in {0, 1} taps are non-zero and ever read; they are kept compactly
            // ([source][dx = 0, 1][dy = 1, 0]: 64 KB instead of 147 KB, which buys nine ring slots instead of four)
            if (cls_bind && (wtap % 3 == 2 || wtap / 3 == 2)) continue;
            const int blk = cls_bind ? wcc * 4 + (wtap % 3) * 2 + (1 - wtap / 3)
                          : p_halo ? wcc * 9 + (wtap % 3) * 3 + (2 - wtap / 3) : nt * num_kb + kb;
            tma_load_3d(&tmB, &ctl->wfull, sW + (size_t)blk * kBBytes, (cls_b * num_kb + kb) * kBlockK, nt * BN, 0);
          }
      }
      int stage = 0;
      uint32_t phase = 0;
      int dbg_row = 0;
      if (p_halo) {
        for (int seg = seg0; seg < P.total_items; seg += sstep) {
          int img, x0, y0, nr;
          decode_seg(seg, img, x0, y0, nr);
          for (int r = 0; r < nr + 2; ++r) {
            for (int src = 0; src < p_nsrc; ++src) {
              mbar_wait(&ctl->empty[stage], phase ^ 1);
              if (p_trace != nullptr && blockIdx.x == 0 && dbg_row < 96) p_trace[512 + 2 * dbg_row] = clock64();
              ++dbg_row;
              mbar_arrive_expect_tx(&ctl->full[stage], kHaloBytes);
              tma_load_4d(src == 0 ? &tmA0 : &tmA1, &ctl->full[stage], sA + (size_t)stage * P.a_slot, 0,
                          x0 - 1 + (cls_b & 1), y0 - 1 + r + (cls_b >> 1), img);
              if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else {
        const uint32_t full0 = (CG == 2) ? mapa_u32(smem_u32(&ctl->full[0]), 0) : 0u;   // leader's barriers
        for (int item = item0; item < P.total_items; item += istep) {
          const Item it = decode(item, (int)cta_rank);
          int pad_y = p.pad, pad_x = p.pad;
          if (p.classes == 4) { pad_y = 1 - (it.cls >> 1); pad_x = 1 - (it.cls & 1); }
          const int wz = p.w_batched ? it.img : 0;
          const int wk0 = it.cls * num_kb * kBlockK;
          const Item i2 = p_pair ? decode(item, 1) : it;
          if (kDxs && P.dxs) {
            // stages in the order (ky, K block, kx); the kx = 0 stage carries the two tiles' halo rows
            // (130 pixels from x0 - 1), all three carry their tap's B block
            for (int g = 0; g < 3 * chunks; ++g) {
              const int ky = g / chunks, cc = g - ky * chunks;
              const CUtensorMap* tm = cc < p.chunks0 ? &tmA0 : &tmA1;
              const int ch = (cc < p.chunks0 ? cc : cc - p.chunks0) * kBlockK;
              for (int kx = 0; kx < 3; ++kx) {
                mbar_wait(&ctl->empty[stage], phase ^ 1);
                uint8_t* a_dst = sA + (size_t)stage * P.a_slot;
                mbar_arrive_expect_tx(&ctl->full[stage], (kx == 0 ? 2 * kHaloBytes : 0) + kBBytes);
                if (kx == 0) {
                  tma_load_4d(tm, &ctl->full[stage], a_dst, ch, it.x0 - 1, it.y0 + ky - 1, it.img);
                  tma_load_4d(tm, &ctl->full[stage], a_dst + kHaloSlot, ch, i2.x0 - 1, i2.y0 + ky - 1, i2.img);
                }
                tma_load_3d(&tmB, &ctl->full[stage], sB + (size_t)stage * kBBytes,
                            wk0 + ((ky * 3 + kx) * chunks + cc) * kBlockK, it.n_tile * BN, wz);
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
              }
            }
            continue;
          }
          for (int kb = 0; kb < num_kb; ++kb) {
            const int tap = kb / chunks, cc = kb - tap * chunks;
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            uint8_t* a_dst = sA + (size_t)stage * P.a_slot;
            if (CG == 2) {
              // both CTAs' bytes are counted on the leader's barrier
              const uint32_t fb = full0 + (uint32_t)stage * 8u;
              if (cta_rank == 0) mbar_arrive_expect_tx(&ctl->full[stage], 2 * (kABytes + kBBytes));
              if (p.mode == 0) {
                const int ky = tap / p.kw, kx = tap - ky * p.kw;
                const int dy = ky - pad_y, dx = kx - pad_x;
                if (cc < p.chunks0)
                  tma_load_4d_2cta(&tmA0, fb, a_dst, cc * kBlockK, it.x0 + dx, it.y0 + dy, it.img);
                else
                  tma_load_4d_2cta(&tmA1, fb, a_dst, (cc - p.chunks0) * kBlockK, it.x0 + dx, it.y0 + dy, it.img);
              } else {
                const int ey = (tap >> 2) - 1, ex = (tap & 3) - 1;
                const int qy = ey >> 1, ry = ey & 1, qx = ex >> 1, rx = ex & 1;
                tma_load_5d_2cta(&tmA0, fb, a_dst, rx * p.cin0 + cc * kBlockK, it.x0 + qx, ry, it.y0 + qy, it.img);
              }
              tma_load_3d_2cta(&tmB, fb, sB + (size_t)stage * kBBytes, wk0 + kb * kBlockK,
                               it.n_tile * BN + (int)cta_rank * (BN / 2), wz);
              if (++stage == P.stages) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_arrive_expect_tx(&ctl->full[stage], (kABytes << p_pair) + (p_wres ? 0 : kBBytes));
            if (p.mode == 0) {
              const int ky = tap / p.kw, kx = tap - ky * p.kw;
              const int dy = ky - pad_y, dx = kx - pad_x;
              if (cc < p.chunks0)
                tma_load_4d(&tmA0, &ctl->full[stage], a_dst, cc * kBlockK, it.x0 + dx, it.y0 + dy,
                            it.img);
              else
                tma_load_4d(&tmA1, &ctl->full[stage], a_dst, (cc - p.chunks0) * kBlockK, it.x0 + dx,
                            it.y0 + dy, it.img);
              if (p_pair) {   // second M tile of the item, same tap / channel chunk
                if (cc < p.chunks0)
                  tma_load_4d(&tmA0, &ctl->full[stage], a_dst + kABytes, cc * kBlockK, i2.x0 + dx,
                              i2.y0 + dy, i2.img);
                else
                  tma_load_4d(&tmA1, &ctl->full[stage], a_dst + kABytes, (cc - p.chunks0) * kBlockK,
                              i2.x0 + dx, i2.y0 + dy, i2.img);
              }
            } else {
              const int ey = (tap >> 2) - 1, ex = (tap & 3) - 1;
              const int qy = ey >> 1, ry = ey & 1, qx = ex >> 1, rx = ex & 1;
              tma_load_5d(&tmA0, &ctl->full[stage], a_dst, rx * p.cin0 + cc * kBlockK, it.x0 + qx,
                          ry, it.y0 + qy, it.img);
            }
            if (!p_wres)
              tma_load_3d(&tmB, &ctl->full[stage], sB + (size_t)stage * kBBytes, wk0 + kb * kBlockK,
                          it.n_tile * BN, wz);
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    constexpr uint32_t idesc = idesc_f16(kBlockM, BN);
    if (p_wres) mbar_wait(&ctl->wfull, 0);
    // warp-reductions return provably uniform values: lets the compiler keep the TMEM address
    // and the descriptors in uniform registers instead of re-broadcasting them per MMA
    const uint32_t taddr_u = __reduce_or_sync(0xffffffffu, taddr);
    const uint32_t sA_u = __reduce_or_sync(0xffffffffu, smem_u32(sA));
    const uint32_t sB_u = __reduce_or_sync(0xffffffffu, smem_u32(sB));
    const uint32_t sW_u = __reduce_or_sync(0xffffffffu, smem_u32(sW));
    const uint64_t desc_hi = smem_desc_sw128(0);     // everything but the start-address field
    auto mkdesc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFFu); };
    int stage = 0;
    uint32_t phase = 0;
    uint32_t tcount = 0;  // tiles issued by this CTA
    if (p_halo) {
      // Row-streaming 3x3: input row ri (image row y0 - 1 + ri) of a segment contributes to the
      // output rows j = ri - dy (dy = 0, 1, 2).  Their accumulators are neighbouring 64-column
      // TMEM slots ((tiles so far + j) & 7), and the weights of one dx are stacked [dy=2; dy=1;
      // dy=0], so the whole contribution is ONE MMA with N = 64 * (number of targets) per
      // (dx, k16) -- split in two only where the slot ring wraps, and on the very first
      // (dx, k16) = (0, 0) step, where the new target (dy = 0) must overwrite, not accumulate.
      constexpr uint32_t idesc0 = idesc_f16(kBlockM, 0);   // + (N >> 3) << 17
      for (int seg = seg0; seg < P.total_items; seg += sstep) {
        int img, x0, y0, nr;
        decode_seg(seg, img, x0, y0, nr);
        for (int ri = 0; ri < nr + 2; ++ri) {
         for (int src = 0; src < p_nsrc; ++src) {
          mbar_wait(XF ? &ctl->xfull[stage] : &ctl->full[stage], phase);
          if (ri < nr && src == 0) {   // slot of the new target must have been drained
            const uint32_t tn = tcount + (uint32_t)ri;
            mbar_wait(&ctl->tmem_empty[tn & 7u], ((tn >> 3) & 1u) ^ 1u);
          }
          tc_fence_after();
          const bool tr = p_trace != nullptr && blockIdx.x == 0 && lane == 0 && ri >= 2 && src == 0 &&
                          tcount + (uint32_t)ri - 2u < 64u;
          if (tr) p_trace[(tcount + ri - 2) * 8 + 0] = clock64();   // last input row + slot ready
          // rows3: the dy = 2 taps are zero as well, so row ri only feeds the output rows ri - 1 and ri
          // (N = 128 instead of 192; the last row of a segment feeds nothing)
          const int j_lo = max(ri - (cls_bind ? 1 : 2), 0), j_hi = min(ri, nr - 1);
          const uint32_t row_lo = (sA_u + (uint32_t)stage * P.a_slot) >> 4;
          const uint32_t w_lo = (sW_u >> 4) + (uint32_t)(src * (cls_bind ? 4 : 9)) * (kBBytes >> 4);   // this source's taps
          const bool last_src = (src == p_nsrc - 1);
          if (elect_one()) {
            // Descriptors of step s = dx*4 + k: A = row + 2*s (dx*128 B + k*32 B, in 16-byte
            // units), B = first target's block + dx*3 blocks + 2*k.  All twelve are the row's base
            // descriptor plus a compile-time constant: independent 64-bit adds, no chains.
            constexpr uint32_t kBlk = kBBytes >> 4;                 // one 64x64 weight block
            const uint64_t da0 = desc_hi | (uint64_t)row_lo;
            const uint32_t sa = (tcount + (uint32_t)j_lo) & 7u;      // TMEM slot of the first target
            const int cnt = j_hi - j_lo + 1;
            const int c1 = min(cnt, 8 - (int)sa);                    // targets before the slot ring wraps
            const uint32_t dxs_blk = cls_bind ? 2u : 3u;             // weight blocks per dx (rows3: dy = 1, 0 only)
            const uint64_t db0 = desc_hi | (uint64_t)(w_lo + (uint32_t)((cls_bind ? 1 : 2) - ri + j_lo) * kBlk);
            const uint32_t d0 = taddr_u + sa * 64u;
            // step 0: targets that already hold a partial sum accumulate, the new one (j = ri, first
            // source only) overwrites its slot
            const int n_old = (ri < nr && src == 0) ? cnt - 1 : cnt;  // old targets come first
            if (!kRows3 || cnt > 0) {
              const int o1 = min(n_old, c1);                         // old targets before the wrap
              if (o1 > 0) umma_f16(d0, da0, db0, idesc0 | ((uint32_t)(o1 * 8) << 17), 1u);
              if (n_old > o1)
                umma_f16(taddr_u, da0, db0 + (uint64_t)(o1 * kBlk),
                         idesc0 | ((uint32_t)((n_old - o1) * 8) << 17), 1u);
              if (ri < nr && src == 0) {
                const uint32_t sn = (tcount + (uint32_t)ri) & 7u;
                umma_f16(taddr_u + sn * 64u, da0, db0 + (uint64_t)(n_old * kBlk), idesc0 | (8u << 17), 0u);
              }
            }
            const int smax = cls_bind ? 8 : 12;                     // rows3: the dx = 2 taps are zero, skip them
            if (kRows3 && cnt <= 0) {
              // nothing to issue (rows3, last row of the segment)
            } else if (c1 == cnt) {
              const uint32_t id = idesc0 | ((uint32_t)(cnt * 8) << 17);
#pragma unroll
              for (int s = 1; s < 12; ++s)
                if (s < smax)
                  umma_f16(d0, da0 + (uint64_t)(2 * s), db0 + (uint64_t)((s >> 2) * dxs_blk * kBlk + (s & 3) * 2),
                           id, 1u);
            } else {
              const uint32_t id1 = idesc0 | ((uint32_t)(c1 * 8) << 17);
              const uint32_t id2 = idesc0 | ((uint32_t)((cnt - c1) * 8) << 17);
              const uint64_t db1 = db0 + (uint64_t)(c1 * kBlk);
#pragma unroll
              for (int s = 1; s < 12; ++s) {
                if (s >= smax) break;
                const uint64_t bo = (uint64_t)((s >> 2) * dxs_blk * kBlk + (s & 3) * 2);
                umma_f16(d0, da0 + (uint64_t)(2 * s), db0 + bo, id1, 1u);
                umma_f16(taddr_u, da0 + (uint64_t)(2 * s), db1 + bo, id2, 1u);
              }
            }
            umma_commit(&ctl->empty[stage]);                       // input row consumed
            if (ri >= 2 && last_src) umma_commit(&ctl->tmem_full[(tcount + (uint32_t)(ri - 2)) & 7u]);
          }
          __syncwarp();
          if (tr) p_trace[(tcount + ri - 2) * 8 + 2] = clock64();   // row's MMAs issued
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
         }
        }
        tcount += (uint32_t)nr;
      }
    } else {
      for (int item = item0; item < P.total_items; item += istep) {
        if (CG == 2 && cta_rank != 0) break;    // only the leader CTA of a pair issues MMAs
        const Item it = decode(item);
        const uint32_t acc = tcount & 1;
        mbar_wait(&ctl->tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_addr = taddr_u + ((acc * BN) << p_pair);
        const uint32_t w_base = sW_u + (uint32_t)(it.n_tile * num_kb) * kBBytes;
        if (kDxs && P.dxs) {
          for (int g = 0; g < 3 * chunks; ++g) {
            uint32_t a_grp = 0;
            int st[3];
            for (int kx = 0; kx < 3; ++kx) {
              mbar_wait(&ctl->full[stage], phase);
              tc_fence_after();
              st[kx] = stage;
              if (kx == 0) a_grp = sA_u + (uint32_t)stage * P.a_slot;
              const uint32_t b_addr = sB_u + (uint32_t)stage * kBBytes;
              if (elect_one()) {
                const uint64_t da = mkdesc(a_grp + (uint32_t)kx * 128u);          // view shifted by kx pixels
                const uint64_t da2 = mkdesc(a_grp + kHaloSlot + (uint32_t)kx * 128u);
                const uint64_t db = mkdesc(b_addr);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                           (g > 0 || kx > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16(d_addr + BN, da2 + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                           (g > 0 || kx > 0 || k > 0) ? 1u : 0u);
                if (kx == 2) {
                  // the halo rows of the kx = 0 stage were read by all three taps: release the group together
                  umma_commit(&ctl->empty[st[0]]);
                  umma_commit(&ctl->empty[st[1]]);
                  umma_commit(&ctl->empty[st[2]]);
                  if (g == 3 * chunks - 1) umma_commit(&ctl->tmem_full[acc]);
                }
              }
              __syncwarp();
              if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
          }
          ++tcount;
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = sA_u + (uint32_t)stage * P.a_slot;
          const uint32_t b_addr = p_wres ? w_base + (uint32_t)kb * kBBytes
                                         : sB_u + (uint32_t)stage * kBBytes;
          if (elect_one()) {
            const uint64_t da = mkdesc(a_addr);
            const uint64_t db = mkdesc(b_addr);
            if (CG == 2) {
              constexpr uint32_t idesc2 = idesc_f16(2 * kBlockM, BN);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16_2cta(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2,
                              (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit_2cta(&ctl->empty[stage]);
              if (kb == num_kb - 1) umma_commit_2cta(&ctl->tmem_full[acc]);
            } else {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
              if (p_pair) {   // the second M tile reuses the B stage
                const uint64_t da2 = mkdesc(a_addr + kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16(d_addr + BN, da2 + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(&ctl->empty[stage]);
              if (kb == num_kb - 1) umma_commit(&ctl->tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        ++tcount;
      }
    }
  } else if (XF && warp >= kThreads / 32) {
    // =============================== input transform (XF) ========================
    // Row-parallel: transform warp w owns the ring rows g = w, w + 6, w + 12, ... (g counts the rows the
    // producer loads, in order) and transforms each of them alone, so six rows are in flight at
    // once and the latency chain of a row (barrier wait -> LDS -> FMA -> MUFU -> FMA -> STS -> proxy
    // fence -> arrive) overlaps with five others instead of being paid once per row by everybody.
    constexpr int kXfWarps = kXfThreads / 32;
    const int xw = warp - kThreads / 32;             // 0..5
    const int grp = lane & 7;                        // logical 16-byte chunk = channels 8 grp .. 8 grp + 7
    const int pq = lane >> 3;                        // pixel of the slot this lane touches first (then + 4)
    constexpr int kCells = (kHaloPix + 3) / 4;       // 33 cells per lane and row
    constexpr int kBatch = 6;                        // cells in flight per lane
    const bool exact = (P.dbg_flags & 128) != 0;     // A/B switch: ex2 + rcp SiLU as in k_gn_apply
    long long g = 0;                                 // ring row counter (all rows of all segments of this CTA)
    for (int seg = seg0; seg < P.total_items; seg += sstep) {
      int img, x0, y0, nr;
      decode_seg(seg, img, x0, y0, nr);
      // (A, B) of this lane's eight channels for this image (k_gn_coef wrote them; L2 hits), halved:
      // SiLU(t) = h + h tanh(h) with h = t / 2 = x (A / 2) + B / 2  (scaling by 1/2 is exact)
      float2 cA[4], cB[4];
      {
        const float4* cf = reinterpret_cast<const float4*>(p.gn_coef + (size_t)img * 64 + grp * 8);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ab = __ldg(cf + q);             // (A0, B0, A1, B1)
          cA[q] = make_float2(0.5f * ab.x, 0.5f * ab.z);
          cB[q] = make_float2(0.5f * ab.y, 0.5f * ab.w);
        }
      }
      // columns of the slot that lie inside the image: pixel pp <-> x = x0 - 1 + pp
      const int pp_lo = (x0 == 0) ? 1 : 0;
      const int pp_hi = min(kHaloPix, p.Wo - x0 + 1);   // exclusive
      for (int r = 0; r < nr + 2; ++r, ++g) {
        if ((int)(g % kXfWarps) != xw) continue;
        const int stage = (int)(g % P.stages);
        const uint32_t phase = (uint32_t)((g / P.stages) & 1);
        mbar_wait(&ctl->full[stage], phase);
        const int y = y0 - 1 + r;
        if (y >= 0 && y < p.Ho) {
          uint8_t* slot = sA + (size_t)stage * P.a_slot;
#pragma unroll 1
          for (int c0 = 0; c0 < kCells; c0 += kBatch) {
            uint4 v[kBatch];
            bool on[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const int pp = pq + 4 * (c0 + u);
              on[u] = (c0 + u < kCells) && pp >= pp_lo && pp < pp_hi;
              if (on[u]) v[u] = *reinterpret_cast<const uint4*>(slot + pp * 128 + ((grp ^ (pp & 7)) << 4));
            }
            if (exact) {
#pragma unroll
              for (int u = 0; u < kBatch; ++u) {
                __half2* h = reinterpret_cast<__half2*>(&v[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 hv = __ffma2_rn(__half22float2(h[e]), cA[e], cB[e]);
                  const float2 yv = xf_silu2(__fadd2_rn(hv, hv));
                  h[e] = __floats2half2_rn(yv.x, yv.y);
                }
              }
            } else {
              // three stages over the batch, so that the MUFU results are consumed long after issue
              float2 hv[kBatch][4];
              float tx[kBatch][4], ty[kBatch][4];
#pragma unroll
              for (int u = 0; u < kBatch; ++u) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e) hv[u][e] = __ffma2_rn(__half22float2(h[e]), cA[e], cB[e]);
              }
#pragma unroll
              for (int u = 0; u < kBatch; ++u)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  asm volatile("tanh.approx.f32 %0, %1;" : "=f"(tx[u][e]) : "f"(hv[u][e].x));
                  asm volatile("tanh.approx.f32 %0, %1;" : "=f"(ty[u][e]) : "f"(hv[u][e].y));
                }
#pragma unroll
              for (int u = 0; u < kBatch; ++u) {
                __half2* h = reinterpret_cast<__half2*>(&v[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 yv = __ffma2_rn(hv[u][e], make_float2(tx[u][e], ty[u][e]), hv[u][e]);
                  h[e] = __floats2half2_rn(yv.x, yv.y);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const int pp = pq + 4 * (c0 + u);
              if (on[u]) *reinterpret_cast<uint4*>(slot + pp * 128 + ((grp ^ (pp & 7)) << 4)) = v[u];
            }
          }
        }
        fence_proxy_async();                 // generic-proxy writes -> visible to tcgen05.mma
        mbar_arrive(&ctl->xfull[stage]);
      }
    }
  } else {
    // =============================== epilogue ===================================
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;               // which half of the columns this warp drains
    const int row = quarter * 32 + lane;
    const int e = threadIdx.x - 64;                 // 0..255
    const int tyr = row >> p.tile_w_log2, txr = row & (tile_w - 1);
    constexpr int kHalfCols = BN / 2;
    uint32_t tcount = 0;

    uint32_t scount = 0;   // tiles staged so far (tcount counts items: one or, in pair mode, two tiles)
    auto do_tile = [&](int img, int x0, int y0, int n_tile, int cls, int sub = 0, bool last = true) {
      const int n0 = n_tile * BN;
      const int cpy = cls >> 1, cpx = cls & 1;
      const uint32_t acc = tcount & acc_mask;
      // (1) the TMA store that last used this staging buffer must have finished reading it
      uint8_t* sOut = sO + (scount % kOutBufs) * kStageOut;
      if (e == 0) {
        if (kOutBufs == 3) bulk_wait_read2();
        else if (kOutBufs == 2) bulk_wait_read1();
        else bulk_wait_read0();
      }
      const bool tr = p_trace != nullptr && blockIdx.x == 0 && e == 0 && tcount < 64;
      epi_bar();
      if (tr) p_trace[tcount * 8 + 4] = clock64();       // all epilogue warps arrived
      mbar_wait(&ctl->tmem_full[acc], (tcount >> acc_log2) & 1);
      tc_fence_after();
      if (tr) p_trace[tcount * 8 + 5] = clock64();       // accumulator complete
      const uint32_t trow = taddr + ((acc * BN) << p_pair) + (uint32_t)(sub * BN) +
                            ((uint32_t)(quarter * 32) << 16);
      const float* sbias = ctl->bias + n0;

      const int oy = (y0 + tyr) * p.out_scale + cpy, ox = (x0 + txr) * p.out_scale + cpx;
      const long long off = (long long)img * p.out_img_stride + (long long)oy * p.out_row_stride +
                            (long long)ox * p.out_pix_stride + n0;

      float ln_mean = 0.f, ln_rstd = 0.f;
      if (EPI == EPI_LN_RES) {
        // channel LayerNorm over the whole row (BN == Cout); the two threads that share a row
        // exchange their half-row partial sums through shared memory (exact two-pass variance)
        float s = 0.f;
        for (int c = half * 32; c < BN; c += 64) {
          uint32_t v[32];
          tmem_ld32(trow + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) s += __uint_as_float(v[j]) + sbias[c + j];
        }
        ctl->rowsum[half][row] = s;
        epi_bar();
        ln_mean = (ctl->rowsum[0][row] + ctl->rowsum[1][row]) * (1.f / BN);
        epi_bar();
        float ss = 0.f;
        for (int c = half * 32; c < BN; c += 64) {
          uint32_t v[32];
          tmem_ld32(trow + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) + sbias[c + j] - ln_mean;
            ss += d * d;
          }
        }
        ctl->rowsum[half][row] = ss;
        epi_bar();
        ln_rstd = rsqrtf((ctl->rowsum[0][row] + ctl->rowsum[1][row]) * (1.f / BN) + 1e-5f);
      }
      const int qkv_part = (EPI == EPI_QKV) ? (n0 >> 7) : 0;

      // one 64-column box per pass; inside it this warp owns columns [half * 32, half * 32 + 32)
#pragma unroll 1
      for (int c = half * 32; c < BN; c += 64) {
        if (c >= 64) {   // the previous box's TMA store must have read the staging tile
          if (e == 0) bulk_wait_read0();
          epi_bar();
        }
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
        if (c + 64 >= BN) {   // last read of this accumulator (pair mode: of the item's last tile)
          tc_fence_before();
          if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&ctl->tmem_empty[acc]), 0));   // leader's barrier
          else if (last) mbar_arrive(&ctl->tmem_empty[acc]);
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + sbias[c + j];

        if (EPI == EPI_GN && p.res != nullptr) {
          // second half of a source-split convolution: add the partial sum of the first half
          // (so the GroupNorm statistics see the complete convolution output)
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
        if (EPI == EPI_GN) {
          // per 8-column sub-block (sum, sumsq) of this row, then a halving butterfly over the
          // warp: 8 values x 32 lanes -> 9 shuffles, fixed order (bit-reproducible)
          float w8[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float sm = 0.f, sq = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = f[g * 8 + j];
              sm += x;
              sq = fmaf(x, x, sq);
            }
            w8[g * 2] = sm;
            w8[g * 2 + 1] = sq;
          }
          const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
          float w4[4], w2[2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = h16 ? w8[i] : w8[i + 4];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            w4[i] = (h16 ? w8[i + 4] : w8[i]) + recv;
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = h8 ? w4[i] : w4[i + 2];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
            w2[i] = (h8 ? w4[i + 2] : w4[i]) + recv;
          }
          const float send = h4 ? w2[0] : w2[1];
          float t = (h4 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, send, 4);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          if ((lane & 3) == 0) {
            const int idx = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);   // = sub-block*2 + moment
            ctl->stats[(quarter * 32 + (c >> 3)) * 2 + idx] = t;
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
          for (int j = 0; j < 32; ++j) f[j] = (f[j] - ln_mean) * ln_rstd * ctl->gain[c + j];
        }
        if (EPI == EPI_GNRES) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + off + c);
          const float4* cf = reinterpret_cast<const float4*>(p.gn_coef + (size_t)img * (BN * P.n_tiles) + n0 + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = __ldg(rp + q);
            const __half2* h = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 ab = __ldg(cf + q * 4 + j);
              const float2 r2 = __half22float2(h[j]);
              const float t0 = fmaf(r2.x, ab.x, ab.y), t1 = fmaf(r2.y, ab.z, ab.w);
              f[q * 8 + j * 2 + 0] += silu_tanh(t0);
              f[q * 8 + j * 2 + 1] += silu_tanh(t1);
            }
          }
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
        // fp16 -> staging tile: box (c / 64), row `row`, 16-byte chunk index XOR (row & 7)
        uint8_t* box = sOut + (size_t)row * 128;
        const int ch0 = (c & 63) >> 3;
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
          *reinterpret_cast<uint4*>(box + (((ch0 + q) ^ (row & 7)) << 4)) = o;
        }
        // box complete and visible to the async proxy -> store it
        fence_proxy_async();
        epi_bar();
        if (e == 0 && !(P.dbg_flags & 4)) {
          const CUtensorMap* tmo = cls == 0 ? &tmO0 : cls == 1 ? &tmO1 : cls == 2 ? &tmO2 : &tmO3;
          tma_store_4d(tmo, sOut, n0 + (c & ~63), x0, y0, img);
          bulk_commit();
        }
      }
      if (tr) p_trace[tcount * 8 + 6] = clock64();       // accumulator drained
      if (EPI == EPI_GN) {
        // one 64-bit fixed-point atomic per (group, moment) of this tile: order-independent.
        // Slots were each written exactly once above; summed here in a fixed order.
        const int ngrp = BN >> p.gs_log2;
        if (e < ngrp * 2) {
          const int grp = e >> 1, m = e & 1;
          const int sb0 = grp << (p.gs_log2 - 3), nsb = 1 << (p.gs_log2 - 3);
          float v = 0.f;
          for (int sb = sb0; sb < sb0 + nsb; ++sb) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v += ctl->stats[(q * 32 + sb) * 2 + m];
          }
          const int g0 = n0 >> p.gs_log2;
          atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + ((size_t)img * 8 + g0) * 2 + e,
                    (unsigned long long)__float2ll_rn(v * (float)kStatScale));
        }
      }
      if (tr) p_trace[tcount * 8 + 7] = clock64();       // tile done
      if (EPI == EPI_QKV) {
        if (qkv_part == 1 && p.colmax != nullptr && e < 128) {
          atomicMax(&p.colmax[img * 128 + e], ctl->colmax[e]);
          ctl->colmax[e] = INT_MIN;   // slot e is touched again only after the next epi_bar
        }
      }
      ++scount;
      if (last) ++tcount;
    };


    // ---- BN = 64: warp-independent epilogue.  The eight warps form two groups that take tiles
    // alternately (tile t -> group t & 1); inside a group each warp owns its TMEM lane quarter =
    // 32 pixels x all 64 channels, stages them in its own 4 KB slab and stores the slab with its
    // own TMA box (32 pixels x 64 ch).  No CTA-level barrier: the latency chains of consecutive
    // tiles overlap instead of adding up (the eight-slot accumulator ring gives the slack).
    auto do_tile64 = [&](int img, int x0, int y0, int cls) {
      const int grp = (warp - 2) >> 2;
      if ((int)(tcount & 1u) != grp) { ++tcount; return; }
      const uint32_t acc = tcount & acc_mask;
      const int cpy = cls >> 1, cpx = cls & 1;
      const int oy = (y0 + tyr) * p.out_scale + cpy, ox = (x0 + txr) * p.out_scale + cpx;
      const long long off = (long long)img * p.out_img_stride + (long long)oy * p.out_row_stride +
                            (long long)ox * p.out_pix_stride;
      const bool has_res = (EPI == EPI_RES || EPI == EPI_LN_RES || EPI == EPI_GNRES ||
                            (EPI == EPI_GN && p.res != nullptr));
      uint4 rv[8];
      if (has_res) {   // this pixel's 64 residual channels = one 128-byte line; issued before the wait
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + off);
#pragma unroll
        for (int q = 0; q < 8; ++q) rv[q] = __ldg(rp + q);
      }
      const bool tr = p_trace != nullptr && blockIdx.x == 0 && warp == 2 + 4 * grp && lane == 0 && tcount < 64;
      if (tr) p_trace[tcount * 8 + 4] = clock64();
      mbar_wait(&ctl->tmem_full[acc], (tcount >> acc_log2) & 1);
      tc_fence_after();
      if (tr) p_trace[tcount * 8 + 5] = clock64();       // accumulator complete
      const uint32_t trow = taddr + acc * BN + ((uint32_t)(quarter * 32) << 16);
      uint32_t v0[32], v1[32];
      tmem_ld32(trow, v0);
      tmem_ld32(trow + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&ctl->tmem_empty[acc]);                // accumulator back to the MMA warp
      if (tr) p_trace[tcount * 8 + 6] = clock64();       // accumulator drained
      float f[64];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 ba = *reinterpret_cast<const float4*>(ctl->bias + q * 8);
        const float4 bb = *reinterpret_cast<const float4*>(ctl->bias + q * 8 + 4);
        const uint32_t* vv = (q < 4) ? (v0 + q * 8) : (v1 + (q - 4) * 8);
        f[q * 8 + 0] = __uint_as_float(vv[0]) + ba.x; f[q * 8 + 1] = __uint_as_float(vv[1]) + ba.y;
        f[q * 8 + 2] = __uint_as_float(vv[2]) + ba.z; f[q * 8 + 3] = __uint_as_float(vv[3]) + ba.w;
        f[q * 8 + 4] = __uint_as_float(vv[4]) + bb.x; f[q * 8 + 5] = __uint_as_float(vv[5]) + bb.y;
        f[q * 8 + 6] = __uint_as_float(vv[6]) + bb.z; f[q * 8 + 7] = __uint_as_float(vv[7]) + bb.w;
      }
      if (EPI == EPI_GN && p.res != nullptr) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __half2* h = reinterpret_cast<const __half2*>(&rv[q]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 r2 = __half22float2(h[j]);
            f[q * 8 + j * 2 + 0] += r2.x;
            f[q * 8 + j * 2 + 1] += r2.y;
          }
        }
      }
      if (EPI == EPI_GN) {
        // (sum, sumsq) of the eight 8-channel sub-blocks of this row, then a halving butterfly over
        // the 32 rows of the warp: 16 values x 32 lanes -> 16 shuffles, fixed order.  Lanes 0, 2,
        // .., 30 end up with one (sub-block, moment) total each and add it to the image's
        // fixed-point statistics (64-bit integer atomics: order-independent, bit-reproducible).
        float w16[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float sm = 0.f, sq = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = f[g * 8 + j];
            sm += x;
            sq = fmaf(x, x, sq);
          }
          w16[g * 2] = sm;
          w16[g * 2 + 1] = sq;
        }
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
        float w8[8], w4[4], w2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float send = h16 ? w16[i] : w16[i + 8];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
          w8[i] = (h16 ? w16[i + 8] : w16[i]) + recv;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = h8 ? w8[i] : w8[i + 4];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
          w4[i] = (h8 ? w8[i + 4] : w8[i]) + recv;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = h4 ? w4[i] : w4[i + 2];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
          w2[i] = (h4 ? w4[i + 2] : w4[i]) + recv;
        }
        const float send = h2 ? w2[0] : w2[1];
        float t = (h2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, send, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        if ((lane & 1) == 0 && !(P.dbg_flags & 8)) {
          const int idx = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);  // sub-block*2 + moment
          const int grp_c = (idx >> 1) >> (p.gs_log2 - 3);
          atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + ((size_t)img * 8 + grp_c) * 2 + (idx & 1),
                    (unsigned long long)__float2ll_rn(t * (float)kStatScale));
        }
      } else if (EPI == EPI_LN_RES) {
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) sm += f[j];
        const float mean = sm * (1.f / 64.f);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float d = f[j] - mean;
          ss = fmaf(d, d, ss);
        }
        const float rstd = rsqrtf(ss * (1.f / 64.f) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 64; ++j) f[j] = (f[j] - mean) * rstd * ctl->gain[j];
      }
      if (EPI == EPI_RES || EPI == EPI_LN_RES) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __half2* h = reinterpret_cast<const __half2*>(&rv[q]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 r2 = __half22float2(h[j]);
            f[q * 8 + j * 2 + 0] += r2.x;
            f[q * 8 + j * 2 + 1] += r2.y;
          }
        }
      }
      float ln_m = 0.f, ln_r = 0.f;
      if (EPI == EPI_GNRES) {
        // y = res_conv(x) + SiLU(GroupNorm(raw) [* (scale + 1) + shift]): rv holds this pixel's raw row
        const float4* cf = reinterpret_cast<const float4*>(p.gn_coef + (size_t)img * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __half2* h = reinterpret_cast<const __half2*>(&rv[q]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 ab = __ldg(cf + q * 4 + j);        // (A, B) of two consecutive channels
            const float2 r2 = __half22float2(h[j]);
            const float t0 = fmaf(r2.x, ab.x, ab.y), t1 = fmaf(r2.y, ab.z, ab.w);
            f[q * 8 + j * 2 + 0] += silu_tanh(t0);
            f[q * 8 + j * 2 + 1] += silu_tanh(t1);
          }
        }
        if (p.has_ln_out) {
          // LayerNorm of the fp16-rounded y (what the attention would otherwise re-read)
          float sm4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            f[j] = __half2float(__float2half_rn(f[j]));
            sm4[j & 3] += f[j];
          }
          ln_m = ((sm4[0] + sm4[1]) + (sm4[2] + sm4[3])) * (1.f / 64.f);
          float ss4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float d = f[j] - ln_m;
            ss4[j & 3] = fmaf(d, d, ss4[j & 3]);
          }
          ln_r = rsqrtf(((ss4[0] + ss4[1]) + (ss4[2] + ss4[3])) * (1.f / 64.f) + 1e-5f);
        }
      }
      if (p_direct) {
        // no staging: this thread writes its pixel's 64 channels (one 128-byte line) itself -- the
        // shared memory goes to the 147 KB of resident weights of the two-source conv instead
        uint4* gdst = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 o;
          __half2 h0 = __floats2half2_rn(f[q * 8 + 0], f[q * 8 + 1]);
          __half2 h1 = __floats2half2_rn(f[q * 8 + 2], f[q * 8 + 3]);
          __half2 h2 = __floats2half2_rn(f[q * 8 + 4], f[q * 8 + 5]);
          __half2 h3 = __floats2half2_rn(f[q * 8 + 6], f[q * 8 + 7]);
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = *reinterpret_cast<uint32_t*>(&h3);
          gdst[q] = o;
        }
        if (tr) p_trace[tcount * 8 + 7] = clock64();
        ++tcount;
        return;
      }
      // the slab's previous TMA store must have finished reading it
      uint8_t* slab = sO + (size_t)(grp * 4 + quarter) * 4096;
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      uint8_t* srow = slab + (size_t)lane * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 o;
        __half2 h0 = __floats2half2_rn(f[q * 8 + 0], f[q * 8 + 1]);
        __half2 h1 = __floats2half2_rn(f[q * 8 + 2], f[q * 8 + 3]);
        __half2 h2 = __floats2half2_rn(f[q * 8 + 4], f[q * 8 + 5]);
        __half2 h3 = __floats2half2_rn(f[q * 8 + 6], f[q * 8 + 7]);
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2);
        o.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(srow + ((q ^ (lane & 7)) << 4)) = o;
      }
      if (EPI == EPI_GNRES && p.has_ln_out) {
        // second slab (after the eight y slabs): LayerNorm_c(y) * g
        uint8_t* lrow = slab + 8 * 4096 + (size_t)lane * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 o;
          __half2 h0 = __floats2half2_rn((f[q * 8 + 0] - ln_m) * ln_r * ctl->gain[q * 8 + 0], (f[q * 8 + 1] - ln_m) * ln_r * ctl->gain[q * 8 + 1]);
          __half2 h1 = __floats2half2_rn((f[q * 8 + 2] - ln_m) * ln_r * ctl->gain[q * 8 + 2], (f[q * 8 + 3] - ln_m) * ln_r * ctl->gain[q * 8 + 3]);
          __half2 h2 = __floats2half2_rn((f[q * 8 + 4] - ln_m) * ln_r * ctl->gain[q * 8 + 4], (f[q * 8 + 5] - ln_m) * ln_r * ctl->gain[q * 8 + 5]);
          __half2 h3 = __floats2half2_rn((f[q * 8 + 6] - ln_m) * ln_r * ctl->gain[q * 8 + 6], (f[q * 8 + 7] - ln_m) * ln_r * ctl->gain[q * 8 + 7]);
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(lrow + ((q ^ (lane & 7)) << 4)) = o;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !(P.dbg_flags & 4)) {
        const CUtensorMap* tmo = cls == 0 ? &tmO0 : cls == 1 ? &tmO1 : cls == 2 ? &tmO2 : &tmO3;
        const int px = quarter * 32;
        tma_store_4d(tmo, slab, 0, x0 + (px & (tile_w - 1)), y0 + (px >> p.tile_w_log2), img);
        if (EPI == EPI_GNRES && p.has_ln_out)
          tma_store_4d(&tmO1, slab + 8 * 4096, 0, x0 + (px & (tile_w - 1)), y0 + (px >> p.tile_w_log2), img);
        bulk_commit();
      }
      if (tr) p_trace[tcount * 8 + 7] = clock64();       // tile done
      ++tcount;
    };

    if constexpr (BN == 64) {
      if (p_halo) {
        for (int seg = seg0; seg < P.total_items; seg += sstep) {
          int img, x0, y0, nr;
          decode_seg(seg, img, x0, y0, nr);
          for (int j = 0; j < nr; ++j) do_tile64(img, x0, y0 + j, cls_b);
        }
      } else {
        for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
          const Item it = decode(item);
          do_tile64(it.img, it.x0, it.y0, it.cls);
        }
      }
      if (lane == 0) bulk_wait0();
    } else {
      if (p_halo) {
        for (int seg = seg0; seg < P.total_items; seg += sstep) {
          int img, x0, y0, nr;
          decode_seg(seg, img, x0, y0, nr);
          for (int j = 0; j < nr; ++j) do_tile(img, x0, y0 + j, 0, 0);
        }
      } else {
        for (int item = item0; item < P.total_items; item += istep) {
          if (CG == 2) {
            const Item it = decode(item, (int)cta_rank);
            do_tile(it.img, it.x0, it.y0, it.n_tile, it.cls);
            continue;
          }
          for (int sub = 0; sub <= p_pair; ++sub) {
            const Item it = decode(item, sub);
            do_tile(it.img, it.x0, it.y0, it.n_tile, it.cls, sub, sub == p_pair);
          }
        }
      }
      if (e == 0) bulk_wait0();
    }
    tc_fence_before();
  }

  if (CG == 2) cluster_sync_all();   // neither CTA leaves while the other may still signal it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta(taddr, kTmemCols);
    else tmem_dealloc(taddr, kTmemCols);
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

int tmap_encode_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
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
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box "
              "%u,%u,%u)", (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return PRG_ERR_CUDA;
  }
  return PRG_OK;
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

struct Conv2Launch {
  CUtensorMap tmA0, tmA1, tmB, tmO[4];
  Conv2Params P;
  int bn, epi, smem;
  int grid;
  int cg;   // 2 = CTA-pair kernel (cluster of two, tcgen05 cta_group::2)
  int xf;   // 1 = input transform warps (GroupNorm apply of the producing Block on the ring slot)
};

// PRG_CONV_FLAGS (A/B measurements; read whenever a layer is planned): bit 0 no row-streaming mode, 1 no
// per-tap pair mode, 4 no CTA-pair kernel, 5 no two-source row-streaming, 6 no input transform, 7 ex2 + rcp
// SiLU in the transform warps, 8 no shared halo rows in the pair-mode 3x3 convs, 10 streamed weights for the
// 64 -> 64 stride-2 conv
static int conv_flags() {
  const char* e = getenv("PRG_CONV_FLAGS");
  return e ? atoi(e) : 0;
}

// Segment length for the halo mode: MMA time scales with the rows a CTA owns; the two extra halo
// rows per segment cost a TMA load and twelve N = 64 MMAs each (about half an output row); pick
// the length that minimises the busiest CTA.
static void halo_segments(Conv2Params* P, int B, int sms) {
  const int Ho = P->c.Ho, strips = P->c.tiles_x;
  int best_r = 8;
  double best_cost = 1e30;
  for (int r = 4; r <= 32 && r <= Ho; ++r) {
    const int segs = (Ho + r - 1) / r;
    const long long items = (long long)B * strips * segs;
    const long long waves = (items + sms - 1) / sms;
    const double cost = (double)waves * (r + 1.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_r = r; }
  }
  if (const char* e = getenv("PRG_HALO_RSEG")) best_r = std::max(1, atoi(e));  // tuning override
  P->rseg = best_r;
  P->segs_per_strip = (Ho + best_r - 1) / best_r;
  P->total_items = B * strips * P->segs_per_strip;
}

// rows3 = 1: the folded nearest-x2 upsample + 3x3 conv, 128 -> 64 channels, as FOUR class-bound row-streaming
// convolutions in one launch.  Parity class (py, px) of the output is a 2x2-tap conv on the low-res input;
// written as a 3x3 conv whose ky = 2 / kx = 2 taps are zero on the input shifted by (py, px), it runs on the
// two-source row-streaming kernel (the 128 input channels = two 64-channel sources of the same tensor):
// every input row is loaded once per class instead of once per tap, the class's weights stay resident, and
// the zero kx = 2 taps are not issued.  `w` = [64][4 classes][3][3][128] (packing.upsample_rows3_weight).
// s0 is the 128-channel input; the per-tap form of this layer was TMA-bound at a third of the tensor rate.
static int conv2_plan(Conv2Launch* L, int epi, int B, const ActSrc& s0_in, const ActSrc* s1_in, int mode,
                      int ksize, int classes, const __half* w, int w_batched, int Cout,
                      const ActSrc& out, int rows3 = 0) {
  ActSrc s0 = s0_in, s1v;
  const ActSrc* s1 = s1_in;
  if (rows3) {
    if (s0_in.C != 128 || s1_in != nullptr || Cout != 64 || mode != 0 || ksize != 3 || classes != 4 || w_batched ||
        epi != EPI_BIAS || s0_in.W % kBlockM != 0 || num_sms() < 4) {
      set_error("conv_plan: the class-bound row-streaming upsample needs 128 -> 64 channels, rows of 128 pixels");
      return PRG_ERR_ARG;
    }
    s0.C = 64;
    s1v = s0;
    s1v.ptr = s0.ptr + 64;
    s1 = &s1v;
    classes = 1;            // per CTA: one class, planned like a 3x3 two-source conv
  }
  memset(L, 0, sizeof(*L));
  Conv2Params& P = L->P;
  ConvParams& p = P.c;
  const int Ho = (mode == 1) ? s0.H / 2 : s0.H, Wo = (mode == 1) ? s0.W / 2 : s0.W;
  if (s0.C % 64 != 0 || (s1 && s1->C % 64 != 0) || Cout % 64 != 0) {
    set_error("conv_plan: channel counts must be multiples of 64 (%d,%d->%d)", s0.C, s1 ? s1->C : 0,
              Cout);
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
  const int tile_w = Wo < kBlockM ? Wo : kBlockM;
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
  p.out_scale = (classes == 4 || rows3) ? 2 : 1;
  L->epi = epi;

  int bn = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : 64;
  if (epi == EPI_QKV) bn = 128;
  if (epi == EPI_LN_RES) bn = Cout;
  if (bn != 64 && bn != 128 && bn != 256) {
    set_error("conv_plan: unsupported N tile %d", bn);
    return PRG_ERR_ARG;
  }
  L->bn = bn;
  P.n_tiles = Cout / bn;
  const int cin = s0.C + (s1 ? s1->C : 0);
  const int ntaps = (mode == 1) ? 16 : p.kh * p.kw;
  P.num_kb = ntaps * (cin / 64);
  P.m_tiles = p.tiles_x * p.tiles_y * B;
  const int b_bytes = bn * 128;
  // BN = 64: eight 4 KB slabs (+ eight more when EPI_GNRES also stores the LayerNorm output)
  const int stage_out = (bn == 64) ? (epi == EPI_GNRES ? 4 * kABytes : 2 * kABytes) : kABytes;
  const int fixed = stage_out + kCtlBytes + 1024;

  // ---- mode selection
  // all weights of one class (rows3: the 2 x 2 non-zero taps of the two sources only)
  const long long w_all = rows3 ? 8ll * b_bytes : (long long)P.n_tiles * P.num_kb * b_bytes;
  P.halo = 0;
  P.wres = 0;
  const bool halo_ok = !w_batched && classes == 1 && mode == 0 && ksize == 3 && s1 == nullptr &&
                       cin == 64 && bn == 64 && tile_w == kBlockM && P.n_tiles == 1 &&
                       !(conv_flags() & 1) &&
                       w_all + fixed + 5 * kHaloSlot <= kSmemBudget;
  // cat(x0, x1) of two 64-channel sources -> 64: both weight halves (147 KB) stay resident, the
  // rows of the two sources alternate through a four-slot ring, and the epilogue stores to global
  // memory directly (no room for staging slabs)
  const bool halo2_ok = !w_batched && classes == 1 && mode == 0 && ksize == 3 && s1 != nullptr &&
                        s0.C == 64 && s1->C == 64 && bn == 64 && tile_w == kBlockM && P.n_tiles == 1 &&
                        (epi == EPI_BIAS || epi == EPI_GN) && !(conv_flags() & 33) &&
                        w_all + kCtlBytes + 1024 + 4 * kHaloSlot <= kSmemBudget;
  P.nsrc = 1;
  P.direct_store = 0;
  P.cls_bind = 0;
  if (rows3 && !halo2_ok) {
    set_error("conv_plan: class-bound row-streaming upsample not plannable here");
    return PRG_ERR_ARG;
  }
  if (halo2_ok) {
    P.halo = 1;
    P.wres = 1;
    P.nsrc = 2;
    P.direct_store = 1;
  } else if (halo_ok) {
    P.halo = 1;
    P.wres = 1;
  } else if (!w_batched && classes == 1 && mode == 1 && bn == 64 && P.n_tiles == 1 && epi == EPI_BIAS &&
             !(conv_flags() & 1024) && w_all + kCtlBytes + 1024 + 5 * kABytes <= kSmemBudget) {
    // 4x4 stride-2, 64 -> 64: all sixteen taps resident (128 KB), the epilogue stores to global memory itself
    // (no room for staging slabs): 166 -> 155 us at 128x128 -- the layer is bound by the strided 5-D gathers of
    // its A tiles, not by the bytes
    P.wres = 1;
    P.direct_store = 1;
  } else if (!w_batched && classes == 1 && w_all + fixed + 8 * kABytes <= kSmemBudget) {
    // resident weights only when enough A stages remain to keep ~128 KiB of loads in flight
    P.wres = 1;
  }
  P.dbg_flags = conv_flags();
  P.w_bytes = P.wres ? (int)w_all : 0;
  // pair mode: with N = 128 one B stage (16 KB) only feeds 256 MMA cycles, and the loads in
  // flight (not L2 bandwidth) bound the kernel; two M tiles per B stage need a third less
  P.pair = (bn == 128 && !P.halo && !P.wres && mode == 0 && epi != EPI_QKV && !(conv_flags() & 2) &&
            (p.tiles_x * p.tiles_y) % 2 == 0) ? 1 : 0;
  // 3x3 pair-mode layers on rows of 128 pixels: the three dx taps share one halo row per tile and (dy, K block)
  // -- 82 KB through TMA per 1536 MMA cycles instead of 144 KB (these layers were TMA-bound at ~970 TFLOP/s)
  P.dxs = (P.pair && mode == 0 && ksize == 3 && classes == 1 && tile_w == kBlockM && !(conv_flags() & 256)) ? 1 : 0;
  P.a_slot = P.halo ? kHaloSlot : P.dxs ? 2 * kHaloSlot : (kABytes << P.pair);
  // CTA pair: N = 256 per-tap layers are bound by the bytes each SM pulls through TMA; as a pair
  // each CTA stages only half of B
  L->cg = (bn == 256 && !P.halo && !P.wres && !P.pair && !w_batched && (epi == EPI_BIAS || epi == EPI_GN) &&
           (p.tiles_x * p.tiles_y) % 2 == 0 && !(conv_flags() & 16)) ? 2 : 1;
  int per_stage = P.a_slot + (P.wres ? 0 : b_bytes / L->cg);
  const int fixed_used = P.direct_store ? kCtlBytes + 1024 : fixed;
  int stages = (kSmemBudget - fixed_used - P.w_bytes) / per_stage;
  if (P.dxs && stages < 4) {      // a (dy, K block) group holds three stages at once and the next one must be loadable
    P.dxs = 0;
    P.a_slot = kABytes << P.pair;
    per_stage = P.a_slot + (P.wres ? 0 : b_bytes / L->cg);
    stages = (kSmemBudget - fixed_used - P.w_bytes) / per_stage;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    set_error("conv_plan: shared memory budget too small (%d stages)", stages);
    return PRG_ERR_ARG;
  }
  P.stages = stages;
  L->smem = P.w_bytes + stages * per_stage + fixed_used;

  const int sms = num_sms();
  if (rows3) {
    P.cls_bind = 1;
    halo_segments(&P, B, sms / 4);                 // per class: a quarter of the CTAs
    L->grid = 4 * std::min(P.total_items, sms / 4);
  } else if (P.halo) {
    halo_segments(&P, B, sms);
  } else {
    P.total_items = (P.m_tiles >> (P.pair | (L->cg == 2))) * P.n_tiles * classes;
  }
  if (!rows3) L->grid = (L->cg == 2) ? 2 * std::min(P.total_items, sms / 2) : std::min(P.total_items, sms);

  // ---- tensor maps: activations
  for (int si = 0; si < 2; ++si) {
    const ActSrc* s = si == 0 ? &s0 : s1;
    CUtensorMap* tm = si == 0 ? &L->tmA0 : &L->tmA1;
    if (s == nullptr) { *tm = L->tmA0; continue; }
    const uint64_t ps = (uint64_t)s->pix_stride * 2;
    if (mode == 0) {
      uint64_t dims[4] = {(uint64_t)s->C, (uint64_t)s->W, (uint64_t)s->H, (uint64_t)B};
      uint64_t str[3] = {ps, ps * s->W, ps * s->W * s->H};
      uint32_t box[4] = {64, (uint32_t)((P.halo || P.dxs) ? kHaloPix : tile_w), (uint32_t)(P.halo ? 1 : tile_h), 1};
      int rc = tmap_encode_f16(tm, s->ptr, 4, dims, str, box);
      if (rc) return rc;
    } else {
      if (s->pix_stride != s->C) {
        set_error("conv_plan: stride-2 source must be dense");
        return PRG_ERR_ARG;
      }
      uint64_t dims[5] = {(uint64_t)2 * s->C, (uint64_t)s->W / 2, 2, (uint64_t)s->H / 2, (uint64_t)B};
      uint64_t str[4] = {2 * ps, ps * s->W, 2 * ps * s->W, ps * s->W * s->H};
      uint32_t box[5] = {64, (uint32_t)tile_w, 1, (uint32_t)tile_h, 1};
      int rc = tmap_encode_f16(tm, s->ptr, 5, dims, str, box);
      if (rc) return rc;
    }
  }
  // ---- weights: [nb][Cout][classes*ntaps*cin]
  {
    const uint64_t ktot = (uint64_t)(rows3 ? 4 : classes) * ntaps * cin;
    uint64_t dims[3] = {ktot, (uint64_t)Cout, (uint64_t)(w_batched ? B : 1)};
    uint64_t str[2] = {ktot * 2, ktot * 2 * Cout};
    uint32_t box[3] = {64, (uint32_t)(bn / L->cg), 1};
    int rc = tmap_encode_f16(&L->tmB, w, 3, dims, str, box);
    if (rc) return rc;
  }
  // ---- output maps (one per upsample parity class)
  {
    const int sc = p.out_scale;
    const uint64_t ps = (uint64_t)out.pix_stride * 2;
    for (int cls = 0; cls < 4; ++cls) {
      if (cls >= classes) { L->tmO[cls] = L->tmO[0]; continue; }
      const int cpy = cls >> 1, cpx = cls & 1;
      const __half* base = out.ptr + ((size_t)cpy * out.W + cpx) * out.pix_stride;
      uint64_t dims[4] = {(uint64_t)out.C, (uint64_t)(out.W / sc), (uint64_t)(out.H / sc), (uint64_t)B};
      uint64_t str[3] = {ps * sc, ps * out.W * sc, ps * out.W * out.H};
      uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, 1};
      if (bn == 64) {   // warp-independent epilogue: one box per TMEM lane quarter (32 pixels)
        box[1] = (uint32_t)std::min(tile_w, 32);
        box[2] = 32u / box[1];
      }
      int rc = tmap_encode_f16(&L->tmO[cls], base, 4, dims, str, box);
      if (rc) return rc;
    }
  }
  p.out = const_cast<__half*>(out.ptr);
  p.out_pix_stride = out.pix_stride;
  p.out_row_stride = out.W * out.pix_stride;
  p.out_img_stride = (long long)out.H * out.W * out.pix_stride;
  return PRG_OK;
}

template <int BN, int EPI>
static int launch2(const Conv2Launch& L, cudaStream_t stream) {
  static int configured = 0;
  if (configured < L.smem) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv2<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBudget));
    configured = kSmemBudget;
  }
  PRG_CUDA_OK(launch_pdl(k_conv2<BN, EPI>, dim3(L.grid), dim3(kThreads), L.smem, stream, L.tmA0, L.tmA1, L.tmB, L.tmO[0],
                         L.tmO[1], L.tmO[2], L.tmO[3], L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

template <int EPI>
static int launch2_one_source(const Conv2Launch& L, cudaStream_t stream) {
  static int configured = 0;
  if (configured < L.smem) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv2<64, EPI, 1, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = kSmemBudget;
  }
  PRG_CUDA_OK(launch_pdl(k_conv2<64, EPI, 1, 0, 3>, dim3(L.grid), dim3(kThreads), L.smem, stream, L.tmA0, L.tmA1,
                         L.tmB, L.tmO[0], L.tmO[1], L.tmO[2], L.tmO[3], L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

template <int EPI>
static int launch2_two_source(const Conv2Launch& L, cudaStream_t stream) {
  static int configured = 0;
  if (configured < L.smem) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv2<64, EPI, 1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = kSmemBudget;
  }
  PRG_CUDA_OK(launch_pdl(k_conv2<64, EPI, 1, 0, 2>, dim3(L.grid), dim3(kThreads), L.smem, stream, L.tmA0, L.tmA1,
                         L.tmB, L.tmO[0], L.tmO[1], L.tmO[2], L.tmO[3], L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

static int launch2_rows3(const Conv2Launch& L, cudaStream_t stream) {
  static int configured = 0;
  if (configured < L.smem) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv2<64, EPI_BIAS, 1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBudget));
    configured = kSmemBudget;
  }
  PRG_CUDA_OK(launch_pdl(k_conv2<64, EPI_BIAS, 1, 0, 1>, dim3(L.grid), dim3(kThreads), L.smem, stream, L.tmA0, L.tmA1,
                         L.tmB, L.tmO[0], L.tmO[1], L.tmO[2], L.tmO[3], L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

template <int BN, int EPI>
static int launch2_xf(const Conv2Launch& L, cudaStream_t stream) {
  static int configured = 0;
  if (configured < L.smem) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv2<BN, EPI, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBudget));
    configured = kSmemBudget;
  }
  PRG_CUDA_OK(launch_pdl(k_conv2<BN, EPI, 1, 1>, dim3(L.grid), dim3(kThreads + kXfThreads), L.smem, stream, L.tmA0, L.tmA1,
                         L.tmB, L.tmO[0], L.tmO[1], L.tmO[2], L.tmO[3], L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

template <int BN, int EPI>
static int launch2_pair(const Conv2Launch& L, cudaStream_t stream) {
  static int configured = 0;
  if (configured < L.smem) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_conv2<BN, EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBudget));
    configured = kSmemBudget;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 2 : 1;
  PRG_CUDA_OK(cudaLaunchKernelEx(&cfg, k_conv2<BN, EPI, 2>, L.tmA0, L.tmA1, L.tmB, L.tmO[0], L.tmO[1], L.tmO[2],
                                 L.tmO[3], L.P));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

static int conv2_run(const Conv2Launch& L, cudaStream_t stream) {
  if (L.P.cls_bind) {
    if (L.bn == 64 && L.epi == EPI_BIAS && L.P.halo && L.P.nsrc == 2 && L.P.direct_store) return launch2_rows3(L, stream);
    set_error("conv_run: class-bound plan without its kernel");
    return PRG_ERR_ARG;
  }
  if (L.bn == 64 && L.P.halo && L.P.nsrc == 2 && L.P.direct_store && L.P.wres && !L.xf) {
    if (L.epi == EPI_GN) return launch2_two_source<EPI_GN>(L, stream);
    if (L.epi == EPI_BIAS) return launch2_two_source<EPI_BIAS>(L, stream);
  }
  if (L.bn == 64 && L.P.halo && L.P.nsrc == 1 && !L.P.direct_store && L.P.wres && !L.P.pair && !L.xf && L.cg == 1) {
    if (L.epi == EPI_GN) return launch2_one_source<EPI_GN>(L, stream);
    if (L.epi == EPI_BIAS) return launch2_one_source<EPI_BIAS>(L, stream);
  }
  if (L.xf) {
    if (L.bn == 64 && L.epi == EPI_GN && L.P.halo && L.P.nsrc == 1) return launch2_xf<64, EPI_GN>(L, stream);
    set_error("conv_run: the input transform exists for the row-streaming N = 64 EPI_GN kernel only");
    return PRG_ERR_ARG;
  }
  if (L.cg == 2) {
    if (L.bn == 256 && L.epi == EPI_BIAS) return launch2_pair<256, EPI_BIAS>(L, stream);
    if (L.bn == 256 && L.epi == EPI_GN) return launch2_pair<256, EPI_GN>(L, stream);
    set_error("conv_run: no CTA-pair kernel for N tile %d / epilogue %d", L.bn, L.epi);
    return PRG_ERR_ARG;
  }
#define PRG_CASE(BN_, EPI_) \
  if (L.bn == BN_ && L.epi == EPI_) return launch2<BN_, EPI_>(L, stream);
  PRG_CASE(64, EPI_BIAS) PRG_CASE(128, EPI_BIAS) PRG_CASE(256, EPI_BIAS)
  PRG_CASE(64, EPI_GN) PRG_CASE(128, EPI_GN) PRG_CASE(256, EPI_GN)
  PRG_CASE(128, EPI_QKV)
  PRG_CASE(64, EPI_LN_RES) PRG_CASE(128, EPI_LN_RES) PRG_CASE(256, EPI_LN_RES)
  PRG_CASE(64, EPI_RES) PRG_CASE(128, EPI_RES) PRG_CASE(256, EPI_RES)
  PRG_CASE(64, EPI_GNRES) PRG_CASE(128, EPI_GNRES) PRG_CASE(256, EPI_GNRES)
#undef PRG_CASE
  set_error("conv_run: no kernel for N tile %d / epilogue %d", L.bn, L.epi);
  return PRG_ERR_ARG;
}

// ---- public wrappers (conv_tc.cuh) ---------------------------------------------------------
ConvOp::ConvOp() : impl(nullptr) {}
ConvOp::~ConvOp() { delete reinterpret_cast<Conv2Launch*>(impl); }
ConvOp::ConvOp(const ConvOp& o) : impl(nullptr) {
  if (o.impl) impl = new Conv2Launch(*reinterpret_cast<Conv2Launch*>(o.impl));
}
ConvOp& ConvOp::operator=(const ConvOp& o) {
  if (this != &o) {
    delete reinterpret_cast<Conv2Launch*>(impl);
    impl = o.impl ? new Conv2Launch(*reinterpret_cast<Conv2Launch*>(o.impl)) : nullptr;
  }
  return *this;
}
ConvParams& ConvOp::params() { return reinterpret_cast<Conv2Launch*>(impl)->P.c; }
void conv_op_set_trace(ConvOp& op, long long* buf) { reinterpret_cast<Conv2Launch*>(op.impl)->P.trace = buf; }

int conv_op_plan(ConvOp* op, int epi, int B, const ActSrc& s0, const ActSrc* s1, int mode, int ksize,
                 int classes, const __half* w, int w_batched, int Cout, const ActSrc& out) {
  Conv2Launch* L = new Conv2Launch();
  int rc = conv2_plan(L, epi, B, s0, s1, mode, ksize, classes, w, w_batched, Cout, out);
  if (rc) {
    delete L;
    return rc;
  }
  delete reinterpret_cast<Conv2Launch*>(op->impl);
  op->impl = L;
  return PRG_OK;
}

int conv_op_plan_upsample_rows3(ConvOp* op, int B, const ActSrc& s0, const __half* w_rows3, const ActSrc& out) {
  Conv2Launch* L = new Conv2Launch();
  int rc = conv2_plan(L, EPI_BIAS, B, s0, nullptr, 0, 3, 4, w_rows3, 0, 64, out, 1);
  if (rc) {
    delete L;
    return rc;
  }
  delete reinterpret_cast<Conv2Launch*>(op->impl);
  op->impl = L;
  return PRG_OK;
}

bool conv_op_can_transform_input(const ConvOp& op) {
  const Conv2Launch* L = reinterpret_cast<const Conv2Launch*>(op.impl);
  return L->P.halo && L->P.nsrc == 1 && L->bn == 64 && L->epi == EPI_GN && L->cg == 1 && !(conv_flags() & 64);
}

int conv_op_set_input_transform(ConvOp& op, const float2* coef) {
  Conv2Launch* L = reinterpret_cast<Conv2Launch*>(op.impl);
  if (!conv_op_can_transform_input(op) || coef == nullptr) {
    set_error("conv_op_set_input_transform: not a single-source row-streaming N = 64 EPI_GN plan");
    return PRG_ERR_ARG;
  }
  L->xf = 1;
  L->P.c.gn_coef = coef;
  return PRG_OK;
}

int conv_op_set_ln_out(ConvOp& op, const ActSrc& ln_out) {
  Conv2Launch* L = reinterpret_cast<Conv2Launch*>(op.impl);
  if (L->bn != 64 || L->epi != EPI_GNRES || L->P.c.classes != 1) {
    set_error("conv_op_set_ln_out: only the N = 64 EPI_GNRES epilogue has a LayerNorm output");
    return PRG_ERR_ARG;
  }
  const ConvParams& p = L->P.c;
  const int tile_w = 1 << p.tile_w_log2;
  const uint64_t ps = (uint64_t)ln_out.pix_stride * 2;
  uint64_t dims[4] = {(uint64_t)ln_out.C, (uint64_t)ln_out.W, (uint64_t)ln_out.H, (uint64_t)p.B};
  uint64_t str[3] = {ps, ps * ln_out.W, ps * ln_out.W * ln_out.H};
  uint32_t box[4] = {64, (uint32_t)std::min(tile_w, 32), 0, 1};
  box[2] = 32u / box[1];
  int rc = tmap_encode_f16(&L->tmO[1], ln_out.ptr, 4, dims, str, box);
  if (rc) return rc;
  L->P.c.has_ln_out = 1;
  return PRG_OK;
}

// Runs the planned conv on the first `B` images (B <= the planned batch).
int conv_op_run(ConvOp& op, int B, cudaStream_t stream) {
  Conv2Launch L = *reinterpret_cast<Conv2Launch*>(op.impl);
  Conv2Params& P = L.P;
  P.c.B = B;
  P.m_tiles = P.c.tiles_x * P.c.tiles_y * B;
  if (P.cls_bind) {
    halo_segments(&P, B, num_sms() / 4);
    L.grid = 4 * std::min(P.total_items, num_sms() / 4);
    return conv2_run(L, stream);
  }
  if (P.halo)
    halo_segments(&P, B, num_sms());
  else
    P.total_items = (P.m_tiles >> (P.pair | (L.cg == 2))) * P.n_tiles * P.c.classes;
  L.grid = (L.cg == 2) ? 2 * std::min(P.total_items, num_sms() / 2) : std::min(P.total_items, num_sms());
  return conv2_run(L, stream);
}

const char* conv_op_describe(const ConvOp& op, char* buf, int n) {
  const Conv2Launch* L = reinterpret_cast<const Conv2Launch*>(op.impl);
  snprintf(buf, n, "bn=%d epi=%d halo=%d wres=%d pair=%d dxs=%d cg=%d xf=%d rows3=%d stages=%d smem=%d kb=%d rseg=%d", L->bn, L->epi,
           L->P.halo, L->P.wres, L->P.pair, L->P.dxs, L->cg, L->xf, L->P.cls_bind, L->P.stages, L->smem, L->P.num_kb, L->P.rseg);
  return buf;
}

}  // namespace prg

// ------------------------------------------------------------------------------------------
// Tensor-pipe micro-benchmark (test hook): one thread per CTA issues `iters` back-to-back
// tcgen05.mma (M = 128, N = n, K = 16) on whatever shared memory holds; reports SM cycles per MMA.
// a_shift: byte displacement of the A start address (0 / 128 / 256 = the dx views of the halo
// mode), used to check that displaced 128B-swizzled views run at full rate.
// ------------------------------------------------------------------------------------------
namespace prg {
__global__ void __launch_bounds__(128, 1)
k_mma_rate(int n, int iters, int a_shift, int same_ab, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_addr;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&tmem_addr, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t taddr = tmem_addr;
  if (threadIdx.x < 32) {
    const uint32_t idesc = ptx::idesc_f16(128, 0) | ((uint32_t)(n >> 3) << 17);
    const uint32_t a0 = ptx::smem_u32(smem) + a_shift, b0 = ptx::smem_u32(smem) + 32 * 1024;
    const uint64_t hi = ptx::smem_desc_sw128(0);
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // walk A over 16 KB and B over 48 KB like a real K loop (same_ab: hammer one address)
          const uint32_t ao = same_ab ? 0u : (uint32_t)(((i >> 2) & 0) * 0 + k * 32);
          const uint32_t bo = same_ab ? 0u : (uint32_t)((((i >> 2) % 3) * 8192 * 0) + k * 32);
          ptx::umma_f16(taddr + ((i >> 2) & 1) * 256u, hi | (uint64_t)(((a0 + ao) >> 4) & 0x3FFFu),
                        hi | (uint64_t)(((b0 + bo) >> 4) & 0x3FFFu), idesc, 1u);
        }
      }
      ptx::umma_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    t1 = clock64();
    t0 = __shfl_sync(0xffffffffu, t0, 0);   // elected lane is lane 0 on current hardware
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(taddr, 512);
  }
}

int mma_rate_probe(int grid, int n, int iters, int a_shift, int same_ab, long long* out_dev,
                   cudaStream_t s) {
  static bool cfg = false;
  if (!cfg) {
    PRG_CUDA_OK(cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    cfg = true;
  }
  k_mma_rate<<<grid, 128, 100 * 1024, s>>>(n, iters, a_shift, same_ab, out_dev);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}
}  // namespace prg

// Micro-benchmark behind the z-buffer design: how many min-scatter operations per
// clock per SM the B200 sustains through each path a z-buffer could take.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/bin/zbuf_atomics tools/microbench/zbuf_atomics.cu
// Patterns: 0 = lane-consecutive targets (pixel i -> target i + shift), 1 = the round-1 kernel's pattern (a lane owns
// 4 consecutive pixels, so one instruction touches stride-4 targets), 2 = consecutive with +-32 pseudo-random jitter.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ int target_of(int i, int pattern, int n) {
  int t = i + 37;
  if (pattern == 2) t += (int)(hash32((unsigned)i) & 63) - 32;
  if (t < 0) t = 0;
  if (t >= n) t -= n;
  return t;
}

// ---- global memory: red.min / plain store / load+store, region of n words (L2 resident when n*4 <= ~64 MB)
template <int MODE>   // 0 = atomicMin (RED), 1 = st.cg, 2 = ld.cg + conditional st.cg
__global__ void __launch_bounds__(256) k_global(unsigned* __restrict__ z, int n, int pattern, int reps) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    if (pattern == 1) {
      for (int i4 = tid; i4 * 4 < n; i4 += nth) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t = target_of(i4 * 4 + j, 0, n);
          const unsigned v = hash32(i4 * 4 + j + r) | 0x3f000000u;
          if (MODE == 0) atomicMin(z + t, v);
          else if (MODE == 1) __stcg(z + t, v);
          else { if (v < __ldcg(z + t)) __stcg(z + t, v); }
        }
      }
    } else {
      for (int i = tid; i < n; i += nth) {
        const int t = target_of(i, pattern, n);
        const unsigned v = hash32(i + r) | 0x3f000000u;
        if (MODE == 0) atomicMin(z + t, v);
        else if (MODE == 1) __stcg(z + t, v);
        else { if (v < __ldcg(z + t)) __stcg(z + t, v); }
      }
    }
  }
}

// ---- shared memory of one CTA: atomicMin / plain load+store on a `words`-word tile
template <int MODE>   // 0 = atomicMin (ATOMS), 1 = ld + conditional st (no atomics)
__global__ void __launch_bounds__(1024) k_shared(unsigned* __restrict__ out, int words, int pattern, int reps) {
  extern __shared__ unsigned zs[];
  for (int i = threadIdx.x; i < words; i += blockDim.x) zs[i] = 0xFFFFFFFFu;
  __syncthreads();
  for (int r = 0; r < reps; ++r)
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      const int t = target_of(i, pattern == 1 ? 0 : pattern, words);
      const unsigned v = hash32(i + r + blockIdx.x) | 0x3f000000u;
      if (MODE == 0) atomicMin(zs + t, v);
      else { if (v < zs[t]) zs[t] = v; }
    }
  __syncthreads();
  unsigned acc = 0;
  for (int i = threadIdx.x; i < words; i += blockDim.x) acc ^= zs[i];
  if (acc == 0x12345) out[blockIdx.x] = acc;
}

// ---- distributed shared memory: a cluster of CS CTAs owns CS * words words; every CTA scatters over the whole tile
template <int MODE>   // 0 = atomicMin on shared::cluster, 1 = ld + conditional st on shared::cluster
__global__ void __launch_bounds__(1024) k_dsmem(unsigned* __restrict__ out, int words, int pattern, int reps, int local_only) {
  extern __shared__ unsigned zs[];
  cg::cluster_group cl = cg::this_cluster();
  const int cs = cl.num_blocks(), rank = cl.block_rank();
  for (int i = threadIdx.x; i < words; i += blockDim.x) zs[i] = 0xFFFFFFFFu;
  cl.sync();
  const int total = words * cs;
  for (int r = 0; r < reps; ++r)
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      // this CTA's pixels map half a tile further (=> mostly a neighbour CTA's slice) unless local_only
      int t = target_of(rank * words + i + (local_only ? 0 : words / 2), pattern == 1 ? 0 : pattern, total);
      if (local_only) t = rank * words + (t % words);
      unsigned* remote = cl.map_shared_rank(zs, t / words) + (t % words);
      const unsigned v = hash32(i + r + blockIdx.x) | 0x3f000000u;
      if (MODE == 0) atomicMin(remote, v);
      else { if (v < *remote) *remote = v; }
    }
  cl.sync();
  unsigned acc = 0;
  for (int i = threadIdx.x; i < words; i += blockDim.x) acc ^= zs[i];
  if (acc == 0x12345) out[blockIdx.x] = acc;
}

static float run(void (*launch)(cudaStream_t), int iters = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch(0); launch(0);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    cudaEventRecord(a); launch(0); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
  return best;
}

static unsigned* g_z; static unsigned* g_out; static int g_n, g_pattern, g_reps, g_words, g_sms, g_local;
template <int M> static void l_global(cudaStream_t s) { k_global<M><<<g_sms * 8, 256, 0, s>>>(g_z, g_n, g_pattern, g_reps); }
template <int M> static void l_shared(cudaStream_t s) { k_shared<M><<<g_sms, 1024, g_words * 4, s>>>(g_out, g_words, g_pattern, g_reps); }
template <int M, int CS> static void l_dsmem(cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((g_sms / CS) * CS); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = g_words * 4; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k_dsmem<M>, g_out, g_words, g_pattern, g_reps, g_local);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  g_sms = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;     // nominal max SM clock; ops/clk/SM below use it (an under-estimate if the part clocks lower)
  printf("%s, %d SMs, %.0f MHz nominal\n", p.name, g_sms, clk / 1e6);
  cudaMalloc(&g_z, 256u << 20); cudaMalloc(&g_out, 4096);
  cudaMemset(g_z, 0xFF, 256u << 20);
  const char* pn[3] = {"consecutive", "stride-4/lane", "jitter+-32"};
  const char* gm[3] = {"RED.MIN global", "st.cg", "ld.cg+st.cg"};
  printf("\n-- global / L2 (ops per clock per SM; region MB)\n");
  for (int mb : {8, 32, 128}) for (int pat = 0; pat < 3; ++pat) {
    g_n = mb << 18; g_pattern = pat; g_reps = (mb <= 32) ? 8 : 2;
    float t[3] = {run(l_global<0>), run(l_global<1>), run(l_global<2>)};
    for (int m = 0; m < 3; ++m)
      printf("  %-16s %3d MB %-14s %8.3f ms  %6.2f ops/clk/SM  %7.1f Gop/s\n", gm[m], mb, pn[pat], t[m],
             (double)g_n * g_reps / (t[m] * 1e-3) / clk / g_sms, (double)g_n * g_reps / (t[m] * 1e-3) / 1e9);
  }
  printf("\n-- shared memory, one CTA per SM, 1024 threads, 150 KB tile\n");
  g_words = 150 * 256; g_reps = 64;
  cudaFuncSetAttribute(k_shared<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_words * 4);
  cudaFuncSetAttribute(k_shared<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_words * 4);
  for (int pat : {0, 2}) {
    g_pattern = pat;
    float t0 = run(l_shared<0>), t1 = run(l_shared<1>);
    printf("  ATOMS.MIN        %-14s %8.3f ms  %6.2f ops/clk/SM\n", pn[pat], t0, (double)g_words * g_reps / (t0 * 1e-3) / clk);
    printf("  LDS+cond STS     %-14s %8.3f ms  %6.2f ops/clk/SM\n", pn[pat], t1, (double)g_words * g_reps / (t1 * 1e-3) / clk);
  }
  printf("\n-- distributed shared memory (cluster), 1024 threads per CTA, 150 KB per CTA\n");
  cudaFuncSetAttribute(k_dsmem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_words * 4);
  cudaFuncSetAttribute(k_dsmem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_words * 4);
  cudaFuncSetAttribute(k_dsmem<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int local = 0; local < 2; ++local) for (int pat : {0, 2}) {
    g_pattern = pat; g_local = local; g_reps = 16;
    float a2 = run(l_dsmem<0, 2>), a8 = run(l_dsmem<0, 8>), p2 = run(l_dsmem<1, 2>), p8 = run(l_dsmem<1, 8>);
    const double w = (double)g_words * g_reps / clk;
    printf("  %s %-12s  atomicMin cs2 %6.2f  cs8 %6.2f | ld+st cs2 %6.2f  cs8 %6.2f  ops/clk/SM\n",
           local ? "own slice   " : "remote slice", pn[pat], w / (a2 * 1e-3), w / (a8 * 1e-3), w / (p2 * 1e-3), w / (p8 * 1e-3));
  }
  return 0;
}

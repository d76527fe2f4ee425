// Attention cores of the U-Net (everything between to_qkv and to_out).
//
//  LinearAttention (SDD:748-769) lives in linattn_tc.cu (fused tcgen05 kernels).
//  Attention (mid block, SDD:782-796): flash-style softmax(q k^T) v over n = (S/8)^2 keys.
//
// These reductions are < 1.5 % of the forward FLOPs; the dense
// contractions around them (to_qkv, to_out) run on the tcgen05 engine (conv_tc.cu).
#include "attention.cuh"
#include "common.cuh"
#include "conv_tc.cuh"

namespace prg {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------
// mid-block attention: softmax(q k^T) v, head dim 32, q pre-scaled by the to_qkv epilogue
// CTA = 4 warps x 16 queries of one (image, head); keys/values streamed in tiles of 64.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_attn_mid(const __half* __restrict__ qkv, __half* __restrict__ out, int n) {
  pdl_trigger();
  pdl_wait();
  // [64 rows][4 chunks of 16 B], chunk ^= (row >> 1) & 3
  __shared__ __align__(16) __half sK[64 * 32];
  __shared__ __align__(16) __half sV[64 * 32];
  const int b = blockIdx.z, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 64 + warp * 16;
  const __half* base = qkv + (size_t)b * n * 384;

  // Q fragments straight from global: rows q0 + lane/4 (+8), d = kk*16 + (lane%4)*2 (+8)
  uint32_t qa[2][4];
  {
    const __half* qr0 = base + (size_t)(q0 + (lane >> 2)) * 384 + h * 32 + (lane & 3) * 2;
    const __half* qr1 = qr0 + (size_t)8 * 384;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(qr0 + kk * 16);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(qr1 + kk * 16);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(qr0 + kk * 16 + 8);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(qr1 + kk * 16 + 8);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

  for (int j0 = 0; j0 < n; j0 += 64) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = i * 128 + tid;
      const int row = id >> 2, ch = id & 3;
      const __half* src = base + (size_t)(j0 + row) * 384 + h * 32 + ch * 8;
      const int sw = (row * 4 + (ch ^ ((row >> 1) & 3))) * 8;
      *reinterpret_cast<uint4*>(sK + sw) = __ldg(reinterpret_cast<const uint4*>(src + 128));
      *reinterpret_cast<uint4*>(sV + sw) = __ldg(reinterpret_cast<const uint4*>(src + 256));
    }
    __syncthreads();
    // S = Q K^T : 16 x 64 per warp
    float sacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key blocks
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        // non-transposed: m0 = keys jb..+7 x d kk*16..+7, m1 = same keys x d +8,
        //                 m2 = keys jb+8..+15 x d kk*16..+7, m3 = keys jb+8.. x d +8
        const int row = jp * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int ch = kk * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_addr(sK + (row * 4 + (ch ^ ((row >> 1) & 3))) * 8), b0, b1, b2, b3);
        mma16816(sacc[jp * 2 + 0], qa[kk], b0, b1);
        mma16816(sacc[jp * 2 + 1], qa[kk], b2, b3);
      }
    }
    // online softmax (rows lane/4 and lane/4 + 8)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx0 = fmaxf(mx0, fmaxf(sacc[i][0], sacc[i][1]));
      mx1 = fmaxf(mx1, fmaxf(sacc[i][2], sacc[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float corr0 = __expf(m0 - mx0), corr1 = __expf(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // round to fp16 first so the normaliser matches the operand fed to the MMA
      const __half2 p01 = __floats2half2_rn(__expf(sacc[i][0] - mx0), __expf(sacc[i][1] - mx0));
      const __half2 p23 = __floats2half2_rn(__expf(sacc[i][2] - mx1), __expf(sacc[i][3] - mx1));
      const float2 f01 = __half22float2(p01), f23 = __half22float2(p23);
      rs0 += f01.x + f01.y;
      rs1 += f23.x + f23.y;
      pa[i >> 1][(i & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&p01);
      pa[i >> 1][(i & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&p23);
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i][0] *= corr0; o[i][1] *= corr0;
      o[i][2] *= corr1; o[i][3] *= corr1;
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {     // 16 keys per step
#pragma unroll
      for (int ep = 0; ep < 2; ++ep) {   // pairs of 8-wide e blocks
        const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int ch = ep * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_addr(sV + (row * 4 + (ch ^ ((row >> 1) & 3))) * 8), b0, b1, b2, b3);
        mma16816(o[ep * 2 + 0], pa[kk], b0, b1);
        mma16816(o[ep * 2 + 1], pa[kk], b2, b3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __half* o0 = out + ((size_t)b * n + q0 + (lane >> 2)) * 128 + h * 32 + (lane & 3) * 2;
  __half* o1 = o0 + (size_t)8 * 128;
#pragma unroll
  for (int eb = 0; eb < 4; ++eb) {
    *reinterpret_cast<__half2*>(o0 + eb * 8) = __floats2half2_rn(o[eb][0] * i0, o[eb][1] * i0);
    *reinterpret_cast<__half2*>(o1 + eb * 8) = __floats2half2_rn(o[eb][2] * i1, o[eb][3] * i1);
  }
}

int attn_mid(const __half* qkv, __half* out, int B, int n, cudaStream_t s) {
  if (n % 64 != 0) {
    set_error("attn_mid: n=%d is not a multiple of 64", n);
    return PRG_ERR_ARG;
  }
  dim3 g(n / 64, 4, B);
  PRG_CUDA_OK(launch_pdl(k_attn_mid, g, dim3(128), 0, s, qkv, out, n));
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

}  // namespace prg

// Attention cores (see attention.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace prg {

// qkv (B, n, 384) fp16: [0,128) q', [128,256) k, [256,384) v.  colmax (B,128): max_n k as
// ordered ints.  ctx (B,4,32,32) and zsum (B,128) are 2^-24 fixed-point int64 sums (order-independent,
// bit-reproducible), zero on entry.
int linattn_context(const __half* qkv, const int* colmax, long long* ctx, long long* zsum, int B, int n,
                    cudaStream_t s);
// weff (B, C, 128) fp16 = W_out (C,128) fp32 folded with the normalised context.
int linattn_weff(const float* wout, const long long* ctx, const long long* zsum, __half* weff, int B, int C,
                 int n, cudaStream_t s);
// ---- fused k/v projection + context on the tcgen05 engine (linattn_tc.cu)
constexpr int kPartialFloats = 256 + 4096;   // per (image, CTA range): m[128], z[128], ctx[128][32]
struct KvCtxOp {
  void* impl;
  KvCtxOp();
  ~KvCtxOp();
  KvCtxOp(const KvCtxOp&);
  KvCtxOp& operator=(const KvCtxOp&);
};
int kvctx_max_slots(int maxB);
// xn (maxB, H, W, C) NHWC fp16 (pix_stride elements between pixels); wqkv (384, C) fp16 K-major
// (rows 128..383 = k, v); partials: maxB * kvctx_max_slots(maxB) * kPartialFloats floats.
int kvctx_plan(KvCtxOp* op, int maxB, const __half* xn, int H, int W, int C, int pix_stride,
               const __half* wqkv, float* partials);
// weff (B, C, 128) fp16 = W_out (C,128) fp32 folded with the normalised context of xn[0..B).
int kvctx_run(KvCtxOp& op, int B, const float* wout, __half* weff, int C, cudaStream_t s);

// out (B, n, 128) fp16 = softmax(q k^T) v per head (q already scaled).
int attn_mid(const __half* qkv, __half* out, int B, int n, cudaStream_t s);

}  // namespace prg

// Attention cores (see attention.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace prg {

// ---- fused k/v projection + context on the tcgen05 engine (linattn_tc.cu)
constexpr int kPartialFloats = 256 + 4096;   // per chunk of an image: m[128], z[128], ctx[128][32]
struct KvCtxOp {
  void* impl;
  KvCtxOp();
  ~KvCtxOp();
  KvCtxOp(const KvCtxOp&);
  KvCtxOp& operator=(const KvCtxOp&);
};
size_t kvctx_partial_floats(int maxB, int H, int W);
// xn (maxB, H, W, C) NHWC fp16 (pix_stride elements between pixels); wqkv (384, C) fp16 K-major
// (rows 128..383 = k, v); partials: kvctx_partial_floats(maxB, H, W) floats of scratch.
int kvctx_plan(KvCtxOp* op, int maxB, const __half* xn, int H, int W, int C, int pix_stride,
               const __half* wqkv, float* partials);
// weff (B, C, 128) fp16 = W_out (C,128) fp32 folded with the normalised context of xn[0..B).
int kvctx_run(KvCtxOp& op, int B, const float* wout, __half* weff, int C, cudaStream_t s);

// ---- fused q projection + softmax_d + to_out (W_eff) + LayerNorm + residual (linattn_tc.cu)
struct QOutOp {
  void* impl;
  QOutOp();
  ~QOutOp();
  QOutOp(const QOutOp&);
  QOutOp& operator=(const QOutOp&);
};
// C in {64, 128, 256}.  out = LN_c(W_eff[b] softmax_d(W_q xn) * scale + bias) * gain + res.
int qout_plan(QOutOp* op, int maxB, const __half* xn, int H, int W, int C, const __half* wqkv,
              const __half* weff, const float* bias, const float* gain, const __half* res, __half* out);
int qout_run(QOutOp& op, int B, cudaStream_t s);

// out (B, n, 128) fp16 = softmax(q k^T) v per head (q already scaled).
int attn_mid(const __half* qkv, __half* out, int B, int n, cudaStream_t s);

}  // namespace prg

// tcgen05 implicit-GEMM convolution / GEMM engine: interface (implementation: conv2_tc.cu).
//
// One launch computes  D[m, n] = sum_k A[m, k] * W[n, k]  with
//   m = output pixel (128-pixel tile = tile_w x tile_h patch of one image, NHWC fp16),
//   n = output channel, k = (tap, input channel).
// A tiles are fetched per tap by TMA (tiled mode, 4-D/5-D tensor map over the NHWC
// activation, out-of-bounds = zero padding) straight into the 128B-swizzled K-major layout
// tcgen05.mma consumes; accumulators live in TMEM; four epilogue warps drain them.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace prg {

enum ConvEpi : int {
  EPI_BIAS = 0,    // y = acc + bias                                   -> fp16
  EPI_GN = 1,      // y = acc + bias, + per-(image, group) sum / sumsq -> fp16 raw + stats
  EPI_QKV = 2,     // to_qkv: q softmax_d * scale | k (+ column max) | v -> fp16
  EPI_LN_RES = 3,  // y = LayerNorm_c(acc + bias) * g + residual       -> fp16
  EPI_RES = 4,     // y = acc + bias + residual                        -> fp16
  EPI_GNRES = 5,   // y = acc + bias + SiLU(A[b][c] * raw + B[b][c])    -> fp16 (+ LayerNorm_c(y) * g, N = 64)
                   // = res_conv fused with the block's second GroupNorm apply (SDD:734)
};

struct ConvParams {
  int B, Ho, Wo;         // output grid the M tiles walk (low-res grid when classes == 4)
  int tile_w_log2;       // tile = (1 << tile_w_log2) x (128 >> tile_w_log2) pixels
  int tiles_x, tiles_y;
  int mode;              // 0: stride-1 taps (kh x kw, pad), 1: 4x4 stride-2 pad-1 (5-D parity view)
  int kh, kw, pad;
  int chunks0, chunks1;  // 64-channel K chunks taken from source 0 / source 1 (concat)
  int cin0;              // channels of source 0 (mode 1: parity offset inside the 5-D view)
  int classes;           // 1, or 4 = nearest-x2 upsample folded into four 2x2-tap parity classes
  int w_batched;         // weight matrix differs per image (3rd coordinate of tmB = image)
  // output addressing (elements)
  __half* out;
  long long out_img_stride;
  int out_row_stride, out_pix_stride;
  int out_scale;         // 1, or 2 when classes == 4
  const float* bias;     // [Cout] or nullptr
  const __half* res;     // residual, same addressing as out (EPI_RES / EPI_LN_RES)
  long long* stats;      // EPI_GN: [B][8][2] (sum, sumsq) as 2^-20 fixed point, zeroed by the caller
  int gs_log2;           // EPI_GN: log2(channels per group)
  const float* ln_g;     // EPI_LN_RES: [Cout]
  int* colmax;           // EPI_QKV: [B][128] order-preserving int encoding of max_n k, or nullptr
  int q_softmax;         // EPI_QKV: 1 = softmax over each 32-channel head then * q_scale
  float q_scale;
  const float2* gn_coef; // EPI_GNRES: [B][Cout] (A, B) of the GroupNorm (+ scale/shift) affine, see gn_coef();
                         // input transform (XF): [B][64] (A, B) applied to the INPUT
  int has_ln_out;        // EPI_GNRES, N = 64: also store LayerNorm_c(y) * ln_g through the second output map
};

// Description of one NHWC fp16 activation tensor (source or destination).
struct ActSrc {
  const __half* ptr;   // first element of channel 0 of pixel (0,0) of image 0
  int H, W, C;         // C = channels used (multiple of 64)
  int pix_stride;      // elements between consecutive pixels (>= C)
};

// ---- persistent engine (conv2_tc.cu): plan once for the maximum batch, run for any B <= that.
struct ConvOp {
  void* impl;
  ConvOp();
  ~ConvOp();
  ConvOp(const ConvOp&);
  ConvOp& operator=(const ConvOp&);
  ConvParams& params();   // epilogue extras (bias, stats, res, ...) are filled in by the caller
};
// out: destination tensor (H, W = full output size, C = Cout, pix_stride).
int conv_op_plan(ConvOp* op, int epi, int B, const ActSrc& s0, const ActSrc* s1, int mode, int ksize,
                 int classes, const __half* w, int w_batched, int Cout, const ActSrc& out);
// The folded nearest-x2 upsample + 3x3 conv for 128 -> 64 channels on rows of 128 (or 256, ...) input pixels as
// four class-bound row-streaming convolutions in one launch (see conv2_plan).  s0 = the 128-channel low-res
// input, w_rows3 = [64][4][3][3][128] fp16 (packing.upsample_rows3_weight), out = the full-size output.
int conv_op_plan_upsample_rows3(ConvOp* op, int B, const ActSrc& s0, const __half* w_rows3, const ActSrc& out);
int conv_op_run(ConvOp& op, int B, cudaStream_t stream);
// Row-streaming N = 64 EPI_GN plans with one source can apply y = SiLU(A[b][c] * x + B[b][c]) to their
// INPUT on the fly (the GroupNorm apply of the producing Block; coef = [B][64] (A, B) from gn_coef()),
// so that the activated tensor is never written to or read from HBM.
bool conv_op_can_transform_input(const ConvOp& op);
int conv_op_set_input_transform(ConvOp& op, const float2* coef);
// EPI_GNRES with N = 64: second destination (same shape as the output) for LayerNorm_c(y) * ln_g.
int conv_op_set_ln_out(ConvOp& op, const ActSrc& ln_out);
const char* conv_op_describe(const ConvOp& op, char* buf, int n);
void conv_op_set_trace(ConvOp& op, long long* buf);  // debug: clock64 stamps of CTA 0 (64 tiles x 8)

// fp16 tiled tensor map, 128B swizzle, zero out-of-bounds fill (dims innermost first; strides in
// bytes for dims 1..rank-1).
int tmap_encode_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

constexpr float kStatUnscale = 1.0f / 1048576.0f;  // fixed-point statistics -> float

// order-preserving float <-> int (for atomicMax on floats)
__host__ __device__ inline int float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  int i = __float_as_int(f);
#else
  int i;
  memcpy(&i, &f, 4);
#endif
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ inline float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

}  // namespace prg

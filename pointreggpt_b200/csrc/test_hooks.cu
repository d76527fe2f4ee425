// Test hooks exported through the C ABI (single-layer entry points for parity tests).
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"

using namespace prg;

extern "C" __attribute__((visibility("default"))) int prg_test_conv_f16(
    const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
    int mode, prg_stream_t stream) {
  PRG_CHECK_ARG(x && w && y, "null pointer");
  PRG_CHECK_ARG(mode >= 0 && mode <= 3, "mode");
  ActSrc s{(const __half*)x, H, W, Cin, Cin};
  int Ho = H, Wo = W;
  int cmode = 0, ksize = 1, classes = 1;
  if (mode == 1) { ksize = 3; }
  else if (mode == 2) { cmode = 1; ksize = 4; Ho = H / 2; Wo = W / 2; }
  else if (mode == 3) { ksize = 3; classes = 4; Ho = H * 2; Wo = W * 2; }
  ActSrc o{(const __half*)y, Ho, Wo, Cout, Cout};
  ConvOp op;
  int rc = conv_op_plan(&op, EPI_BIAS, B, s, nullptr, cmode, ksize, classes, (const __half*)w, 0, Cout, o);
  if (rc) return rc;
  op.params().bias = bias;
  if (getenv("PRG_CONV_TRACE") != nullptr) {
    // debug timeline of CTA 0: stamps per tile = rows ready | acc free | MMAs issued | staging
    // free | epilogue warps arrived | acc complete | acc drained | tile done
    long long* d = nullptr;
    PRG_CUDA_OK(cudaMalloc(&d, (64 * 8 + 192) * sizeof(long long)));
    PRG_CUDA_OK(cudaMemset(d, 0, (64 * 8 + 192) * sizeof(long long)));
    conv_op_set_trace(op, d);
    rc = conv_op_run(op, B, (cudaStream_t)stream);
    long long h[64 * 8 + 192];
    PRG_CUDA_OK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    const long long t0 = h[0];
    for (int t = 0; t < 24; ++t) {
      fprintf(stderr, "tile %2d:", t);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", h[t * 8 + k] ? h[t * 8 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
    for (int r = 0; r < 40; ++r)
      fprintf(stderr, "row %2d: issued %7lld  seen %7lld\n", r, h[512 + 2 * r] - t0, h[512 + 2 * r + 1] - t0);
    return rc;
  }
  return conv_op_run(op, B, (cudaStream_t)stream);
}

namespace prg {
int mma_rate_probe(int grid, int n, int iters, int a_shift, int same_ab, long long* out_dev,
                   cudaStream_t s);
}
// Test hook: tensor-pipe rate probe.  out (grid) i64 device: SM cycles for `iters` MMAs per CTA.
extern "C" __attribute__((visibility("default"))) int prg_test_mma_rate(
    int grid, int n, int iters, int a_shift, int same_ab, int64_t* out, prg_stream_t stream) {
  PRG_CHECK_ARG(out && grid >= 1 && n >= 16 && n <= 256 && n % 16 == 0 && iters % 4 == 0, "args");
  return mma_rate_probe(grid, n, iters, a_shift, same_ab, reinterpret_cast<long long*>(out),
                        (cudaStream_t)stream);
}

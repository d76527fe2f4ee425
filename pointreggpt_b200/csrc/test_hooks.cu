// Test hooks exported through the C ABI (single-layer entry points for parity tests).
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
  return conv_op_run(op, B, (cudaStream_t)stream);
}

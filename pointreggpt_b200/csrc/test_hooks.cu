// Test hooks exported through the C ABI (single-layer entry points for parity tests).
#include "common.cuh"
#include "conv_tc.cuh"

using namespace prg;

extern "C" __attribute__((visibility("default"))) int prg_test_conv_f16(
    const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
    int mode, prg_stream_t stream) {
  PRG_CHECK_ARG(x && w && y, "null pointer");
  PRG_CHECK_ARG(mode >= 0 && mode <= 3, "mode");
  ConvLaunch L;
  ActSrc s{(const __half*)x, H, W, Cin, Cin};
  int rc;
  int Ho = H, Wo = W;
  if (mode == 0) {
    rc = conv_plan(&L, EPI_BIAS, B, s, nullptr, 0, 1, 1, (const __half*)w, 0, Cout);
  } else if (mode == 1) {
    rc = conv_plan(&L, EPI_BIAS, B, s, nullptr, 0, 3, 1, (const __half*)w, 0, Cout);
  } else if (mode == 2) {
    rc = conv_plan(&L, EPI_BIAS, B, s, nullptr, 1, 4, 1, (const __half*)w, 0, Cout);
    Ho = H / 2; Wo = W / 2;
  } else {  // nearest x2 upsample folded into four 2x2 parity classes (weights pre-combined)
    rc = conv_plan(&L, EPI_BIAS, B, s, nullptr, 0, 3, 4, (const __half*)w, 0, Cout);
    Ho = H * 2; Wo = W * 2;
  }
  if (rc) return rc;
  L.p.out = (__half*)y;
  L.p.out_pix_stride = Cout;
  L.p.out_row_stride = Wo * Cout;
  L.p.out_img_stride = (long long)Ho * Wo * Cout;
  L.p.bias = bias;
  return conv_run(L, (cudaStream_t)stream);
}

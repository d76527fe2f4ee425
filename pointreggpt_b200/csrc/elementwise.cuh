// Non-GEMM kernels of the U-Net forward / sampler step (HBM- or latency-bound).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace prg {

// ---- stems -------------------------------------------------------------------------------
// Unet.init_conv (SDD:824, 929): 7x7, 1 -> 64, pad 3.  x (B,S,S) f32 -> y (B,S,S,64) f16.
int stem_unet(const float* x, const float* w /*[64][49]*/, const float* bias, __half* y, int B,
              int S, cudaStream_t s);
// MaskUnet: DepthAugment (DC:582-604) + init_conv 7x7, 3 -> 64 (DC:822, 873-874).
int stem_mask(const float* depth01, const float* w /*[64][3][49]*/, const float* bias, __half* y,
              int B, int S, cudaStream_t s);

// ---- conditioning (SDD:845-856, 709-713, 925, 932) ------------------------------------------
struct CondWeights {
  const float *t1w, *t1b, *t2w, *t2b;  // time_mlp: Linear(dim,4dim), Linear(4dim,4dim)
  const float *p1w, *p1b, *p2w, *p2b;  // param_mlp: Linear(pdim,4dim), Linear(4dim,4dim)
  int dim, pdim;
};
// cond_act[b][0:4dim] = SiLU(time_mlp(t_b)), [4dim:8dim] = SiLU(param_mlp(p_b)).
// time comes either from `time` (int64 per image) or, when time == nullptr, from
// `time_scalar` broadcast (sampler).
int cond_embed(const CondWeights& w, const int64_t* time, int time_scalar, const float* pcond,
               float* cond_act, int B, cudaStream_t s);
// sampler fast path (shared timestep, constant param_cond): see elementwise.cu
int cond_time_all(const CondWeights& w, const int* ts_dev, int nsteps, float* act_t, cudaStream_t s);
int cond_mlp_param(const float* W, const float* cond_act, float* ss_p, int rows, int Ktot, int B,
                   cudaStream_t s);
int cond_mlp_step(const float* W, const float* bias, const float* act_t, const float* ss_p, float* ss,
                  int rows, int Ktot, int B, cudaStream_t s, const int* step_idx = nullptr, int act_stride = 0);
// ss[b][r] = W[r] . cond_act[b] + bias[r] for the concatenated rows of every block MLP.
int cond_mlp(const float* W, const float* bias, const float* cond_act, float* ss, int rows, int K,
             int B, cudaStream_t s);

// ---- GroupNorm apply (SDD:690-696) -----------------------------------------------------------
// y = SiLU(((raw - mean_g) * rstd_g * gamma + beta) * (scale + 1) + shift) [+ res]
// stats [B][8][2] = (sum, sumsq) over the fp32 conv outputs; ss: per-image (scale | shift)
// rows of this block inside the cond_mlp output (nullptr = no conditioning).
struct GnApply {
  const __half* raw;
  const long long* stats;  // fixed point (2^-20), see conv_tc.cuh
  const float *gamma, *beta;
  const float* ss;
  int ss_stride;     // floats per image in ss
  int ss_off;        // offset of this block's scale row; shift row is ss_off + C
  const __half* res; // optional residual (same layout as y)
  int res_pix_stride;
  __half* y;
  int HW, C;
  const float* ln_g;  // optional fused channel LayerNorm of y (C <= 256): gain ...
  __half* ln_out;     // ... and destination
};
int gn_apply(const GnApply& a, int B, cudaStream_t s);
// coef[b][c] = (A, B) with y = SiLU(A * raw + B) (uses stats / gamma / beta / ss / HW / C of `a`).
int gn_coef(const GnApply& a, float2* coef, int B, cudaStream_t s);

// ---- ResnetBlock output in one pass (SDD:731-734): y = SiLU(GN(raw)) + res_conv(cat(x0, x1)) + bias
// [+ ln_out = LN_c(y) * ln_g].  The GroupNorm affine per (image, channel) is derived inside the kernel from
// the integer statistics block2's conv epilogue left behind.
struct ResGn {
  const __half *x0, *x1;      // cat(x0, x1) along channels; contiguous NHWC (pixel stride = channel count)
  int c0, c1;                 // channel counts (multiples of 32; c1 = 0 and x1 = nullptr without a skip input)
  const __half* w;            // res_conv weight [Cout][c0 + c1] fp16
  const float* bias;          // [Cout] or nullptr
  const __half* raw;          // block2's raw conv output (B, HW, Cout)
  const long long* stats;     // its GroupNorm statistics [B][8][2] (fixed point, see conv_tc.cuh) ...
  const float *gamma, *beta;  // ... and affine
  __half* y;                  // (B, HW, Cout)
  const float* ln_g;          // optional fused channel LayerNorm of y: gain ...
  __half* ln_out;             // ... and destination
  int HW, Cout;
};
bool res1x1_gn_supported(int cout, int c0, int c1, int HW);
int res1x1_gn(const ResGn& a, int B, cudaStream_t s);

// ---- channel LayerNorm with gain (SDD:619-628): y = LN_c(x) * g [+ res] -----------------------
int ln_apply(const __half* x, const float* g, const __half* res, __half* y, int64_t npix, int C,
             cudaStream_t s);

// ---- network tail ----------------------------------------------------------------------------
// GroupNorm+SiLU of final_res_block.block2 + residual, final 1x1 conv (64 -> 1), then either
//   mode 0: out = conv                                        (Unet.forward, SDD:964)
//   mode 1: out = sigmoid(conv), keep = out > thresh         (MaskUnet tail, DC:868-869)
//   mode 2: one sampler update x_t -> x_{t-1}                 (SDD:1199-1218, 1250-1251,
//                                                              1173-1180, 1279-1280 / 1358-1373)
// One sampler step as the device sees it (prg_sampler_run uploads the whole list once per call; the
// kernels of a step index it with a device-resident counter, so a step's launches are identical for
// every step and can be replayed as ONE CUDA graph).
struct StepDev {
  int kind, add_noise, unnormalize, noise_slab;   // noise_slab: index of this step's Gaussian draw (slab 0 = x_T)
  float c0, c1, c2, c3, c4;
  int pad[3];
};
// per-call pointers of the sampler, device resident for the same reason
struct SamplerCtx {
  const float* img_cond;   // (B,2,HW) or nullptr
  const float* noise;      // injected draws (slabs of B*HW) or nullptr (=> Philox)
};

struct TailParams {
  const __half* raw;
  const long long* stats;
  const float *gamma, *beta;
  const __half* res;      // res_conv output (B,S,S,64)
  const float* fw;        // final conv weight [64]
  const float* fb;        // final conv bias [1]
  int HW;
  int mode;
  float* out;             // mode 0/1: (B,HW) f32 ; mode 2: x_{t-1} (may alias x_t)
  uint8_t* keep;          // mode 1 (optional)
  float thresh;
  // mode 2
  const float* x_t;       // current sample (B,HW)
  const float* img_cond;  // (B,2,HW) or nullptr
  const float* noise;     // (B,HW) or nullptr (=> Philox)
  const unsigned long long* seeds;  // per-image Philox keys (device, B entries)
  unsigned long long noise_offset;  // counter offset of this draw inside an image's stream
  int clip_x_start;       // ddim: clamp the network output before pred_noise
  int use_ddnm;
  int sampler;            // 0 = p_sample, 1 = ddim, 2 = ddim last step (x = x0), 3 = refine
  float c0, c1, c2, c3;   // p_sample: coef1, coef2, sigma, - ; ddim: sqrt_recip, sqrt_recipm1,
                          // sqrt(alpha_next), c ; c4 = sigma
  float c4;
  int add_noise;
  int unnormalize;        // write (x+1)/2 (last step)
  // mode 2, device-resident step list (sampler loop): when `steps` is set the fields above from
  // img_cond to unnormalize are filled inside the kernel from steps[*step_idx] and *ctx
  const StepDev* steps;
  const int* step_idx;
  const SamplerCtx* ctx;
  size_t slab_stride;     // B * HW: distance between two injected-noise slabs
};
// step_idx += 1 (last node of the per-step graph)
int step_advance(int* step_idx, cudaStream_t s);
int net_tail(const TailParams& t, int B, cudaStream_t s);
// the same with the final block's shortcut 1x1 conv + GroupNorm apply computed inside (t.raw / t.stats /
// t.gamma / t.beta / t.res are not used: `a` carries x0, x1, raw, the weights and the gn_coef output)
int net_tail_fused(const TailParams& t, const ResGn& a, int B, cudaStream_t s);

// N(0,1) draws from one Philox stream per image (key = seeds_dev[b], counter = offset + index)
int fill_normal(float* x, int B, int64_t per_image, const unsigned long long* seeds_dev,
                unsigned long long offset, cudaStream_t s);

}  // namespace prg

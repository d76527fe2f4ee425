/* libprg.so -- C ABI of the B200-native PointRegGPT data-generation hot path.
 *
 * The reference (Chen-Suyi/PointRegGPT) has no FFI: its boundary for this path
 * is the Python API of
 *   SDD = denoising_diffusion_pytorch/successive_ddnm_diffusion.py
 *   DC  = depth_correction_pytorch/depth_correction.py
 * Every entry point below cites the reference function it replaces; the Python
 * package `pointreggpt_b200` binds them with ctypes under the reference's own
 * names and signatures.
 *
 * Conventions
 *  - plain C, raw DEVICE pointers unless a parameter says "host", explicit
 *    stream (a cudaStream_t passed as void*; NULL = legacy default stream);
 *  - the caller owns every input/output buffer; the library owns only the
 *    opaque network handles (packed weights, workspace, tensor maps);
 *  - return 0 on success, <0 on error; prg_last_error() (thread-local) has the
 *    message; no exceptions cross the boundary, no hidden host syncs on the
 *    data path (create/destroy do synchronise);
 *  - sm_100a only.  There is no CPU fallback: without a B200 every compute
 *    entry returns PRG_ERR_CUDA.
 */
#ifndef PRG_H_
#define PRG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRG_ABI_VERSION 2

#define PRG_OK 0
#define PRG_ERR_ARG (-1)
#define PRG_ERR_CUDA (-2)
#define PRG_ERR_BLOB (-3)
#define PRG_ERR_STATE (-4)

typedef void* prg_stream_t; /* cudaStream_t */

const char* prg_last_error(void);
int prg_abi_version(void);
/* Number of kernels this library has launched in this process (all streams).
 * bench.py reports the delta over the timed region as "gpu_launches". */
uint64_t prg_launch_count(void);

/* ------------------------------------------------------------------ geometry
 * HBM-bound kernels; results are bit-exact against the reference. */

/* reproject_tensor (SDD:268-286) = depth2pc_tensor -> R p + t -> pc2depth_tensor.
 * depth (B,H,W) f32 metres; K (B,3,3); pose (B,4,4); valid iff clip_lo<d<clip_hi.
 * depth_out (B,H,W) f32 (0 where empty), mask_out (B,H,W) u8 {0,1}. */
int prg_reproject_f32(const float* depth, const float* K, const float* pose,
                      float clip_lo, float clip_hi, float* depth_out, uint8_t* mask_out,
                      int B, int H, int W, prg_stream_t stream);

/* pc2depth_tensor (SDD:212-265) for a ragged batch: pc (sumN,3) f32, valid
 * (sumN) u8 or NULL, offsets (B+1) i64 CSR row starts (device), K (B,3,3).
 * pose (B,4,4) or NULL: optional rigid transform applied first (replaces the
 * host-side `pc @ R.T + t` of Generator.generate, SDD:2533-2535). */
int prg_pc2depth_f32(const float* pc, const uint8_t* valid, const int64_t* offsets,
                     int64_t total_points, const float* K, const float* pose,
                     float* depth_out, uint8_t* mask_out, int B, int H, int W,
                     prg_stream_t stream);

/* depth2pc_tensor (SDD:176-209).  use_clip=0 <=> clip=None.  invalid = the
 * value written for invalid pixels (NaN by default in the reference).
 * pc (B,H*W,3) f32, valid (B,H*W) u8. */
int prg_depth2pc_f32(const float* depth, const float* K, float clip_lo, float clip_hi,
                     int use_clip, float invalid, float* pc, uint8_t* valid,
                     int B, int H, int W, prg_stream_t stream);

/* occlusion_filter (SDD:446-463), used by image_condition(use_occlusion_filter=True)
 * (SDD:492-493) and Tester.sample (SDD:2035-2037).  depth (B,H,W) f32 metres, mask (B,H,W) u8
 * (the z-buffer's outputs); depth_out (B,H,W) must not alias depth.  A pixel farther than
 * 0.0375 behind the nearest valid depth of its 3x3 neighbourhood takes that depth.  The mask
 * is returned unchanged by the reference, so there is no mask output. */
int prg_occlusion_filter_f32(const float* depth, const uint8_t* mask, float* depth_out,
                             int B, int H, int W, prg_stream_t stream);

/* Voxel-grid centroid down-sampling: what the reference asks of open3d's
 * PointCloud.voxel_down_sample at SDD:2486-2500, 2640-2680.  points (N,3) f64 on the device; the grid
 * is anchored at min_bound - voxel_size/2; centroids (up to N,3) f64 and keys_out (up to N) i64 (the
 * packed voxel index ix << 42 | iy << 21 | iz) receive one entry per occupied voxel in unspecified
 * order -- sort by key for a canonical order.  count_err (2 x i32, device): [0] = number of voxels,
 * [1] = 1 if a point was skipped (non-finite coordinate or more than 2^21 voxels along an axis).
 * workspace: prg_voxel_downsample_workspace_bytes(N) bytes, 16-byte aligned.  Sums are order-
 * independent 2^-36 m fixed point: results are reproducible and within 2e-11 m of a float64 mean. */
size_t prg_voxel_downsample_workspace_bytes(int64_t n_points);
int prg_voxel_downsample_f64(const double* points, int64_t n_points, double voxel_size,
                             double* centroids, int64_t* keys_out, int32_t* count_err,
                             void* workspace, size_t workspace_bytes, prg_stream_t stream);

/* Inner loop of compute_overlap_ratio (generate_gt.py:84-101): count_err[0] receives the
 * number of query points (Nq,3 f64) that have a target point (Nt,3 f64) at squared distance
 * < radius^2; count_err[1] = 1 if a point was skipped (non-finite, or beyond +-2^20 cells of size
 * radius).  workspace: prg_overlap_workspace_bytes(Nt) bytes, 16-byte aligned. */
size_t prg_overlap_workspace_bytes(int64_t n_target);
int prg_overlap_count_f64(const double* query, int64_t n_query, const double* target,
                          int64_t n_target, double radius, int32_t* count_err, void* workspace,
                          size_t workspace_bytes, prg_stream_t stream);

/* point_cloud (SDD:122-143) applied to depth01*scale, then optionally the
 * back-transform (pc - t) @ R of SDD:2627-2628 (pose NULL = skip).  Valid
 * pixels are compacted in row-major order.  pc_out: B slabs of H*W*3 f64,
 * counts (B) i64 = points written per slab.  scratch: >= B*(ceil(H*W/1024)+1)
 * i64 of device memory. */
int prg_depth2pc_compact_f64(const float* depth01, const float* K, const float* pose,
                             float scale, float clip_lo, float clip_hi, double* pc_out,
                             int64_t* counts, int64_t* scratch, int B, int H, int W,
                             prg_stream_t stream);

/* ------------------------------------------------------------------ networks
 * tcgen05 tensor-core kernels.  A handle owns packed weights (fp16 K-major
 * tap-major conv matrices, weight standardisation folded in at pack time),
 * the activation workspace for up to max_batch images of size x size, and
 * the TMA descriptors of every layer.
 *
 * The packed blob is produced by pointreggpt_b200.packing.pack_unet /
 * pack_maskunet from a reference state-dict (280 / 234 entries; SDD:802-918,
 * DC:807-869).  It is a host buffer; create copies it to the device. */
typedef struct prg_net prg_net;

#define PRG_NET_UNET 1      /* Unet, SDD:802-964 (time + intrinsics conditioning) */
#define PRG_NET_MASKUNET 2  /* MaskUnet, DC:807-906 (DepthAugment stem, sigmoid tail) */

int prg_net_create(prg_net** out, int kind, const void* blob_host, size_t nbytes,
                   int max_batch, int size, int device);
void prg_net_destroy(prg_net* net);
/* Bytes of device memory held by the handle (weights + workspace). */
size_t prg_net_device_bytes(const prg_net* net);

/* Unet.forward(x, time, param_cond) (SDD:920-964).  x (B,1,S,S) f32, time (B)
 * i64, param_cond (B,4) f32 -> out (B,1,S,S) f32. */
int prg_unet_forward(prg_net* net, const float* x, const int64_t* time,
                     const float* param_cond, float* out, int B, prg_stream_t stream);

/* MaskUnet.forward(x) (DC:871-906) plus the caller's `> thresh` (SDD:2565,
 * 2580).  depth01 (B,1,S,S) f32 -> prob (B,1,S,S) f32 and/or keep (B,1,S,S) u8
 * (either may be NULL). */
int prg_maskunet_forward(prg_net* net, const float* depth01, float* prob, uint8_t* keep,
                         float thresh, int B, prg_stream_t stream);

/* One sampler step.  The host (pointreggpt_b200.diffusion) derives the coefficients from the
 * GaussianDiffusion schedule buffers (SDD:1098-1134) with the same fp32 tensor arithmetic the
 * reference uses, so the device only applies them.
 *   P_SAMPLE     (SDD:1258-1281): c0 = posterior_mean_coef1[t], c1 = posterior_mean_coef2[t],
 *                c2 = exp(0.5 * posterior_log_variance_clipped[t]); add_noise = (t > 0)
 *   DDIM         (SDD:1343-1373): c0 = sqrt_recip_alphas_cumprod[t], c1 = sqrt_recipm1_...[t],
 *                c2 = sqrt(alpha_next), c3 = c, c4 = sigma; add_noise = 1
 *   DDIM_LAST    (SDD:1358-1360): x = x0 (c0, c1 as DDIM)
 *   REFINE_P     (SDD:1307-1314) / REFINE_DDIM (SDD:1375-1389): the optional refine step
 * A step with `unnormalize` set writes (x + 1) / 2 (SDD:1316, 1391); it must be the last. */
typedef struct prg_step {
  int t;
  int kind;
  int add_noise;
  int unnormalize;
  float c0, c1, c2, c3, c4;
} prg_step;

#define PRG_STEP_P_SAMPLE 0
#define PRG_STEP_DDIM 1
#define PRG_STEP_DDIM_LAST 2
#define PRG_STEP_REFINE_P 3
#define PRG_STEP_REFINE_DDIM 4

/* GaussianDiffusion.sample (SDD:1394-1409) for objective pred_x0 with the DDNM null-space
 * replacement (SDD:1210-1218) when img_cond != NULL.
 *  steps: host array.  noise: NULL => device Philox, one stream per image keyed by
 *  philox_seeds[b] (HOST array of B entries; the draws of an image depend on its seed only, not
 *  on batch composition, rank or world size); else (1 + #noisy steps, B,1,S,S) f32: slab 0 =
 *  x_T, slab 1+i = the i-th randn_like draw (parity runs inject the reference's draws; seeds may
 *  then be NULL).  out (B,1,S,S) f32 (in [0,1] when the last step unnormalizes).  B <= max_batch. */
int prg_sampler_run(prg_net* unet, const prg_step* steps, int nsteps, const float* param_cond,
                    const float* img_cond, const float* noise, const uint64_t* philox_seeds,
                    float* out, int B, prg_stream_t stream);

/* The sampler's Gaussian generator on its own (replaces torch.randn / randn_like, SDD:1279, 1293,
 * 1339, 1369): out (B, per_image) f32, image b filled from Philox4x32-10 keyed by seeds[b] (HOST
 * array) at counters offset .. offset + per_image - 1, Box-Muller on the first two words. */
int prg_fill_normal_f32(float* out, int B, int64_t per_image, const uint64_t* seeds,
                        uint64_t offset, prg_stream_t stream);

/* Sampled kernel timing for bench.py's roofline: every n-th network evaluation (0 = off) each
 * launch is bracketed by CUDA events on the launching stream.  prg_profile_read synchronises
 * on the recorded events and returns one entry per kernel family: total device ms, launches,
 * and the number of profiled network evaluations. */
typedef struct prg_profile {
  char name[32];
  uint64_t launches;
  double ms;
  uint64_t forwards;
} prg_profile;
int prg_profile_set(int every_n_forwards);
int prg_profile_read(prg_profile* out, int max_entries, int reset);
/* Per-layer view of the same samples: writes one text line per op,
 * "<U|M><op index>:<description>\t<launches>\t<total ms>\t<algorithmic FLOP per image>\n",
 * into buf (host, cap bytes); returns the bytes written. */
int prg_profile_ops(char* buf, int cap, int reset);

/* Test hook: one implicit-GEMM convolution through the tcgen05 engine.
 * x (B,H,W,Cin) f16 NHWC, w (Cout, taps*Cin) f16 K-major tap-major, bias (Cout)
 * f32 or NULL -> y (B,Ho,Wo,Cout) f16.  mode: 0 = 1x1, 1 = 3x3 p1, 2 = 4x4 s2 p1. */
int prg_test_conv_f16(const void* x, const void* w, const float* bias, void* y, int B,
                      int H, int W, int Cin, int Cout, int mode, prg_stream_t stream);

/* Test hook: exhaustive check of the reprojection kernel's range-restricted reciprocal against the
 * IEEE one: *mismatches_dev (device u64) = number of floats z with lo <= |z| <= hi (both signs) whose
 * results differ. */
int prg_test_frcp_exhaustive(float lo, float hi, uint64_t* mismatches_dev, prg_stream_t stream);

/* Test hook: tensor-pipe rate probe (tools/mma_rate.py).  `grid` CTAs each issue `iters`
 * tcgen05.mma (M=128, N=n, K=16) back to back; out (grid) i64 device = SM cycles taken. */
int prg_test_mma_rate(int grid, int n, int iters, int a_shift, int same_ab, int64_t* out,
                      prg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PRG_H_ */

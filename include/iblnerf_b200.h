/*
 * iblnerf_b200.h -- C ABI of the B200-native IBL-NeRF per-ray hot path (libiblnerf_b200.so).
 *
 * The reference (changwoonchoi/IBL-NeRF) has no FFI of its own: its boundary is the Python module
 * surface of src/nerf_models/*.  Every entry point below replaces the block of eager-PyTorch
 * launches cited next to it (paths relative to the reference's src/).  The Python host package
 * (ibl_nerf_b200/) binds these with ctypes and mirrors the reference API on top of them.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer to contiguous row-major data unless
 *     the name ends in _host; float = IEEE fp32; "nullable" arguments may be NULL.
 *   - every function returns 0 on success or a cudaError_t value (> 0); IBLN_EINVAL (-1) for bad
 *     arguments.  Nothing throws, nothing allocates device memory: the caller owns all buffers.
 *   - last two arguments are always the CUDA device ordinal and the cudaStream_t to launch on;
 *     all entry points are re-entrant (PyTorch's autograd engine calls backward from its own thread).
 */
#ifndef IBLNERF_B200_H
#define IBLNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBLN_EINVAL (-1)
#define IBLN_ABI_VERSION 4   /* 2: ibln_mlp_bwd gained freeze_mode; training-tail / export entry points. 3: ibln_depth_to_normal; word-major relu masks in the stash.
                              * 4: diagnostics (probes, debug switches) left the product ABI (iblnerf_b200_diag.h); training-tail additions */

/* packed per-ray output of the compositing kernels: one row of IBLN_MAPS_STRIDE floats per ray */
#define IBLN_MAPS_STRIDE 24
#define IBLN_MAP_DEPTH 0      /* sum w z                      ibl_nerf_renderer.py:249 */
#define IBLN_MAP_ACC 1        /* sum w                        :259 */
#define IBLN_MAP_DISP 2       /* 1/max(1e-10, depth/acc)      :258 */
#define IBLN_MAP_TEND 3       /* prod(1-alpha+1e-10)          :140-141 (visibility) */
#define IBLN_MAP_ROUGH 4      /* sum wd sigmoid(raw4)         :284-285 */
#define IBLN_MAP_IRR 5        /* sum wd sigmoid(raw5)         :287-288 */
#define IBLN_MAP_ALBEDO 6     /* 3 floats                     :281-282 */
#define IBLN_MAP_RAD 9        /* 3 floats, live weights       :305-306 */
#define IBLN_MAP_COARSE 12    /* 3*n_coarse floats (<= 9)     :311-318 */

/* packed per-ray output of the shading kernel */
#define IBLN_SHADE_STRIDE 16
#define IBLN_SH_NDV 0
#define IBLN_SH_SPEC 1
#define IBLN_SH_DIFF 4
#define IBLN_SH_PRE 7
#define IBLN_SH_COLOR 10

int ibln_abi_version(void);
const char* ibln_error_string(int code);

/* ---- (1) sampling ------------------------------------------------------------------------- */

/* Stratified depths. Replaces ibl_nerf_renderer.py:670-692 (linspace, lerp near/far, mids, jitter).
 * near, far: [N]; t_rand: [N,S] uniforms or NULL (perturb == 0); z_out: [N,S]. */
int ibln_stratified_z(const float* near, const float* far, const float* t_rand, int n_rays, int n_samples,
                      int lindisp, float* z_out, int device, void* stream);

/* Inverse-CDF sampling, full path.  Replaces nerf_renderer_helper.py:91-134 (sample_pdf) given the
 * uniforms.  bins: row r at bins + r*bins_stride, nbins entries; weights: row r at weights +
 * r*w_stride, nbins-1 entries (strides in floats, so the caller can pass weights[:,1:-1] views);
 * u, samples: [N,nsamp]. */
int ibln_sample_pdf(const float* bins, int64_t bins_stride, const float* weights, int64_t w_stride,
                    const float* u, int n_rays, int nbins, int nsamp, float* samples, int device, void* stream);

/* Inverse-CDF given an explicit CDF (the bit-exact entry: inds == torch.searchsorted(cdf,u,right=True),
 * nerf_renderer_helper.py:117-132).  cdf, bins: [N,nbins]; inds_out: [N,nsamp] int64 (nullable). */
int ibln_inverse_cdf(const float* cdf, const float* bins, const float* u, int n_rays, int nbins, int nsamp,
                     int64_t* inds_out, float* samples, int device, void* stream);

/* Fused hierarchical step of render_rays: z mids -> sample_pdf(weights[:,1:-1]) -> sort(cat(z, samples)).
 * Replaces ibl_nerf_renderer.py:702-707.  z: [N,S0], weights: [N,S0], u: [N,S1];
 * z_samples: [N,S1]; z_merged: [N,S0+S1] ascending. */
int ibln_hierarchical_sample(const float* z, const float* weights, const float* u, int n_rays, int s0, int s1,
                             float* z_samples, float* z_merged, int device, void* stream);

/* sort(cat(za, zb)) along the last axis; ibl_nerf_renderer.py:707.  za [N,sa], zb [N,sb], out [N,sa+sb]. */
int ibln_merge_sort_z(const float* za, const float* zb, int n_rays, int sa, int sb, float* z_out,
                      int device, void* stream);

/* ---- (4) alpha compositing ------------------------------------------------------------------ */

/* raw2outputs compositing core, forward.  Replaces ibl_nerf_renderer.py:204-206,241-259,281-318.
 * raw [N,S,C] (C >= 9+3*n_coarse), z [N,S], rays_d [N,3], noise [N,S] nullable.
 * Outputs: weights [N,S]; maps [N,24] linear; maps_srgb [N,24] nullable = pow(x+1e-12,1/2.2) of the
 * colour-like columns (rough/depth/acc/disp/tend copied unchanged; :485-510).
 * radiance_sigmoid: 1 = sigmoid (kitchen), 0 = relu radiance/irradiance (use_radiance_linear). */
int ibln_composite_fwd(const float* raw, const float* z, const float* rays_d, const float* noise,
                       int n_rays, int n_samples, int n_ch, int n_coarse, int radiance_sigmoid,
                       float* weights, float* maps, float* maps_srgb, int device, void* stream);

/* Backward of the above: g_raw [N,S,C] (fully written).  g_weights [N,S], g_maps [N,24],
 * g_maps_srgb [N,24] are each nullable.  Albedo/roughness/irradiance/coarse radiance use DETACHED
 * weights (no gradient to sigma), radiance/depth/acc/disp/weights use live weights (:246,:282-315). */
int ibln_composite_bwd(const float* raw, const float* z, const float* rays_d, const float* noise,
                       const float* g_weights, const float* g_maps, const float* g_maps_srgb,
                       int n_rays, int n_samples, int n_ch, int n_coarse, int radiance_sigmoid,
                       float* g_raw, int device, void* stream);

/* raw2outputs_simple (reflected ray, no grad): ibl_nerf_renderer.py:38-68.
 * pre_out [N,1+n_coarse,3] = radiance, coarse radiance 1..n_coarse. */
int ibln_composite_simple_fwd(const float* raw, const float* z, const float* dirs, int n_rays, int n_samples,
                              int n_ch, int n_coarse, int radiance_sigmoid, float* pre_out, float* pre_srgb /* nullable:
                              pow(x+1e-12, 1/2.2) of pre_out, :485-496 */, int device, void* stream);

/* Depth-only compositing: raw2outputs_depth (:121-150) and raw2depth (normal_from_depth.py:164-169).
 * sigma [reps*N, S] (row m uses z / rays_d of ray m % N), depth [reps*N]; weights [reps*N,S] and
 * visibility [reps*N] nullable. */
int ibln_depth_fwd(const float* sigma, const float* z, const float* rays_d, int reps, int n_rays, int n_samples,
                   float* depth, float* weights, float* visibility, int device, void* stream);

/* ---- (5) normals + split-sum shading --------------------------------------------------------- */

/* Shifted sample points of the epsilon normal estimator: normal_from_depth.py:143-156.
 * pts_out [4,N,S,3] = +eps*right, -eps*right, +eps*up, -eps*up. */
int ibln_normal_eps_points(const float* rays_o, const float* rays_d, const float* z, int n_rays, int n_samples,
                           float eps, float* pts_out, int device, void* stream);

/* Tail of the estimator: normal_from_depth.py:177-183 plus the reflection of :439.
 * depths4 [4,N] (right,left,up,down); normal [N,3]; refl [N,3] nullable.  Optionally also the surface point of
 * ibl_nerf_renderer.py:262: x_surface [N,3] = rays_o + rays_d * depth[r * depth_ld]  (x_surface nullable; rays_o / depth
 * are only read when it is given; depth_ld = 1 for a dense vector, IBLN_MAPS_STRIDE for column 0 of the packed maps). */
int ibln_normal_eps_finish(const float* rays_d, const float* depths4, int n_rays, float eps,
                           float* normal, float* refl, const float* rays_o, const float* depth, int depth_ld,
                           float* x_surface, int device, void* stream);

/* Split-sum shading forward: ibl_nerf_renderer.py:412-438,455-474 + microfacet.py:8-12.
 * rays_d,normal,albedo [N,3]; rough,irr,mip_rough,depth,near,far [N]; prefiltered [N,n_pref,3];
 * lut [lut_c,lut_h,lut_w] (channel 0 = scale, 1 = bias); lut_coef 0 = 'F', 1 = 'F0';
 * out [N,16] linear; out_srgb [N,16] nullable (n.v column copied unchanged). */
int ibln_shade_fwd(const float* rays_d, const float* normal, const float* albedo, const float* rough,
                   const float* irr, const float* mip_rough, const float* depth, const float* near, const float* far,
                   const float* prefiltered, int n_pref, const float* lut, int lut_h, int lut_w,
                   int lut_coef, int correct_depth, int n_rays, float* out, float* out_srgb, int device, void* stream);

/* Backward: gradients reach albedo, rough (also through the LUT row coordinate), irr and mip_rough.
 * g_out, g_out_srgb [N,16] nullable; g_albedo [N,3], g_rough, g_irr, g_mip_rough [N] all written. */
int ibln_shade_bwd(const float* rays_d, const float* normal, const float* albedo, const float* rough,
                   const float* irr, const float* mip_rough, const float* depth, const float* near, const float* far,
                   const float* prefiltered, int n_pref, const float* lut, int lut_h, int lut_w,
                   int lut_coef, int correct_depth, int n_rays, const float* g_out, const float* g_out_srgb,
                   float* g_albedo, float* g_rough, float* g_irr, float* g_mip_rough, int device, void* stream);

/* The same two kernels reading albedo / roughness / irradiance / depth straight from the packed compositing output
 * `maps` [N,24] (no column copies; mip_rough = roughness_map, depth = depth_map as in the shipped configuration), and
 * the backward writing one full row of d loss / d maps per ray: g_maps [N,24] = 0 except roughness (LUT / Fresnel / mip
 * level terms summed), irradiance and albedo columns. */
int ibln_shade_fwd_maps(const float* rays_d, const float* normal, const float* maps, const float* near, const float* far,
                        const float* prefiltered, int n_pref, const float* lut, int lut_h, int lut_w, int lut_coef,
                        int correct_depth, int n_rays, float* out, float* out_srgb, int device, void* stream);
int ibln_shade_bwd_maps(const float* rays_d, const float* normal, const float* maps, const float* near, const float* far,
                        const float* prefiltered, int n_pref, const float* lut, int lut_h, int lut_w, int lut_coef,
                        int correct_depth, int n_rays, const float* g_out, const float* g_out_srgb, float* g_maps,
                        int device, void* stream);

/* ---- (2)+(3) positional encoding and the intrinsic-component MLP ----------------------------- */

/* Point generators understood by the MLP kernels (so sample points never round-trip through HBM):
 *   mode 0  explicit   : pts [P,3] (+ dirs [P/S,3] per ray)                  ibl_nerf.py:236-252
 *   mode 1  ray march  : pts = o + d*z                                       ibl_nerf_renderer.py:200
 *   mode 2  eps normal : 4 shifted copies of the ray march, sigma only       normal_from_depth.py:149-158
 * Encoding: [x, sin(2^k x), cos(2^k x)] k<10 for points (63), k<4 for UN-normalised dirs (27),
 *           positional_embedder.py:9-34. */

/* fp32 exact path (SIMT): explicit encoding + one generic GEMM; used for stage-wise 1e-4 parity and
 * as the high-precision mode of the normal estimator. */
int ibln_encode(const float* x, int64_t n_pts, int n_freqs, float* out, int64_t ld_out, int device, void* stream);
/* expand per-ray rows to per-sample rows while encoding: x [n_rays,3] -> out rows r*S+s */
int ibln_encode_dirs(const float* dirs, int64_t n_rays, int n_samples, int n_freqs, float* out, int64_t ld_out,
                     int device, void* stream);
/* C[M,N] (ldc) = act(A[M,K] (lda) * op(B) + bias) (+ C if accumulate).  trans_b = 1: B is [N,K] (ldb)
 * (a torch Linear weight, forward); trans_b = 0: B is [K,N] (ldb) (dgrad).  relu_mask nullable [M,N]
 * (ld_mask): output is zeroed where mask <= 0 (dgrad through relu).  act 0 none, 1 relu. */
int ibln_sgemm(const float* a, int64_t lda, const float* b, int64_t ldb, int trans_b, const float* bias,
               float* c, int64_t ldc, int64_t m, int n, int k, int act, int accumulate,
               const float* relu_mask, int64_t ld_mask, int device, void* stream);
/* dW[N,K] (ldw) += dY[M,N]^T (ldy) * X[M,K] (ldx); db[N] += colsum(dY) (nullable).  Deterministic
 * two-stage reduction; workspace >= ibln_wgrad_workspace_bytes(n, k) bytes. */
int64_t ibln_wgrad_workspace_bytes(int n, int k);
int ibln_sgemm_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t m, int n, int k,
                     float* dw, int64_t ldw, float* db, int accumulate, void* workspace, int device, void* stream);

/* bf16 tensor-core path (tcgen05 / TMEM / bulk-TMA weight streaming). */
/* Size in bytes of the packed bf16 weight image of one IBLNeRF (kitchen architecture). */
int64_t ibln_mlp_packed_bytes(void);
/* Pack the 46 fp32 state-dict tensors into the kernel's pre-swizzled bf16 chunk stream + fp32 biases.
 * params: HOST array of 46 DEVICE pointers in state-dict order (ibl_nerf.py:44-72):
 * positions_linears.{0..7}.{weight,bias}, views_linears.0.*, feature_linear.*, sigma_linear.*,
 * albedo_feature_linear.*, albedo_linear.*, roughness_linear.*, irradiance_feature_linear.*,
 * irradiance_linear.*, radiance_linear.*, additional_radiance_feature_linear.{0,1,2}.*,
 * additional_radiance_linear.{0,1,2}.*.  Call after every optimizer.step(). */
int ibln_mlp_pack_weights(const float* const* params_host, void* packed, int device, void* stream);

/* Fused encode + MLP forward.  mode per the table above; o,d [n_rays,3]; z [n_rays,S];
 * pts nullable (mode 0: [n_rays*S,3]).  sigma_only = 1: out [P] (P = n_rays*S, or 4*n_rays*S in mode 2),
 * else out [P,18] fp32 straight from the fp32 accumulators; sigma_only = 2: [P,18] WITHOUT the albedo / irradiance
 * feature layer (channels 1..3 and 5 hold the head biases): the reflected-ray march of ibl_nerf_renderer.py:440-448 reads
 * only sigma and the radiance heads (raw2outputs_simple, :38-68).  saved (nullable): activation stash for
 * the backward pass, >= ibln_mlp_saved_bytes(P) bytes. */
int64_t ibln_mlp_saved_bytes(int64_t n_pts);
int ibln_mlp_fwd(const void* packed, int mode, const float* pts, const float* rays_o, const float* rays_d,
                 const float* z, int64_t n_rays, int n_samples, float eps, int sigma_only,
                 float* out, void* saved, int device, void* stream);
/* Backward (dgrad chain kernel + one persistent wgrad kernel): g_out [P,18]; ACCUMULATES (red.global.add) into the
 * flat fp32 gradient image flat_grad (798 994 floats, state-dict order, each tensor row-major: weight then
 * bias per Linear).  `saved` is the stash written by ibln_mlp_fwd for the same points;
 * workspace >= ibln_mlp_bwd_workspace_bytes(P) (per-layer dY tiles).
 * freeze_mode mirrors IBLNeRF.forward_freezed (ibl_nerf.py:88-152, train.py:275-283): 0 = everything trains;
 * 1 = freeze_radiance (only albedo/irradiance feature layers + heads and the roughness head get gradients);
 * 2 = freeze_radiance + freeze_roughness (roughness head frozen too).  Frozen entries of flat_grad are untouched. */
int64_t ibln_mlp_bwd_workspace_bytes(int64_t n_pts);
int ibln_mlp_bwd(const void* packed, const void* saved, const float* g_out, int64_t n_pts,
                 float* flat_grad, void* workspace, int freeze_mode, int device, void* stream);

/* ---- training-step tail ------------------------------------------------------ */

/* Phase-gated image losses of src/train.py:299-441 for ONE pass (fine or coarse) on the packed gamma-corrected outputs
 * of ibln_composite_fwd / ibln_shade_fwd, forward and backward in one launch.  mse = mean over all elements (img2mse,
 * nerf_renderer_helper.py:8):
 *   *loss += scale * [ w_radiance (mse(radiance_map, rgb) + sum_k mse(radiance_map_k, rgb_k))      train.py:326-334, 420-423
 *                    + w_color mse(color_map, rgb)                                                 :323, 437-438
 *                    + w_prior_albedo mse(albedo_map, prior_albedo)                                :401-403, 444-446
 *                    + w_irradiance_reg mse(irradiance_map, irradiance_target) ]                   :410-412, 447
 * maps_srgb [N,24] (irradiance col 5, albedo 6..8, radiance 9..11, coarse radiance k 12+3k..), shade_srgb [N,16] nullable
 * (colour cols 10..12; null = radiance-only phase), rgb_k / prior_albedo [N,3] nullable (term skipped).  g_maps [N,24],
 * g_shade [N,16] receive d loss / d input (all columns written; 16-byte aligned).  *loss must be zeroed by the caller. */
int ibln_image_losses(const float* maps_srgb, const float* shade_srgb, const float* rgb, const float* rgb_1,
                      const float* rgb_2, const float* rgb_3, const float* prior_albedo, int n, float w_radiance,
                      float w_color, float w_prior_albedo, float w_irradiance_reg, float irradiance_target,
                      float scale, float* loss, float* g_maps, float* g_shade, int device, void* stream);

/* torch.optim.Adam step (amsgrad off, no weight decay; src/train.py:479-498) over one flat fp32 buffer:
 * step = 1-based iteration count (bias correction), grad_scale multiplies the gradient first (e.g. 1/world).
 * All four buffers 16-byte aligned, n elements. */
int ibln_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, int step, float grad_scale, int device, void* stream);

/* The same update over a flat buffer that holds n_nets (<= 4) networks back to back (798 994 floats each, state-dict
 * order: training.FlatParameters), followed by the bf16 re-pack (ibln_mlp_pack_weights) of every network straight from that
 * buffer into packed_host[i] (HOST array of device pointers): 2 launches per optimisation step. */
int ibln_adam_step_pack(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int n_nets, float lr,
                        float beta1, float beta2, float eps, int step, float grad_scale, void* const* packed_host,
                        int device, void* stream);
/* Data-parallel training: gradient all-reduce FUSED into the Adam kernel over NVLink / NVSwitch peer memory.
 * Every rank keeps its flat gradient buffer in symmetric memory (torch.distributed._symmetric_memory); this ONE kernel reads
 * the sum over all ranks -- grad_multicast: the buffer's NVSwitch multicast address, read with multimem.ld_reduce.add.v4.f32
 * (the switch adds the ranks' values in flight); or, when grad_multicast is NULL, peer_grads_host: HOST array of `world`
 * peer-mapped device pointers (own buffer included) summed with P2P loads -- scales it by grad_scale (1/world) and applies
 * Adam to this rank's replica.  n % 4 == 0, world <= 8.  The caller brackets it with the symmetric-memory barrier
 * (all gradients written before / all ranks done reading after).  Replaces ncclAllReduce + ibln_adam_step. */
int ibln_adam_allreduce_step(float* param, const float* grad_multicast, const float* const* peer_grads_host, int world,
                             float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                             int step, float grad_scale, int device, void* stream);
/* ... followed by the bf16 re-pack of the n_nets networks that live back to back in `param` (see ibln_adam_step_pack). */
int ibln_adam_allreduce_step_pack(float* param, const float* grad_multicast, const float* const* peer_grads_host, int world,
                                  float* exp_avg, float* exp_avg_sq, int n_nets, float lr, float beta1, float beta2,
                                  float eps, int step, float grad_scale, void* const* packed_host, int device, void* stream);
/* cudaMemsetAsync(buf, 0, bytes) on the given stream (gradient / loss accumulators of the fused training step). */
int ibln_zero(void* buf, int64_t bytes, int device, void* stream);

/* Ray generation + target gather for one training batch: get_rays_few
 * (nerf_renderer_helper.py:14-23) for pixels (u[i], v[i]) of a camera (intrinsics fx, fy, cx, cy; c2w [3,4]
 * row-major, device) -> rays_o, rays_d [N,3], and NerfDataset.get_info pixel gathers
 * (dataset_interface.py:178-197): outputs[k][i,:] = images[k][v[i], u[i], :] for n_images (<= 12) device images
 * [H,W,channels[k]].  images / outputs / channels are HOST arrays (of device pointers / ints). */
int ibln_sample_rays(const int* u, const int* v, int n, int height, int width, float fx, float fy, float cx, float cy,
                     const float* c2w, float* rays_o, float* rays_d, const float* const* images,
                     float* const* outputs, const int* channels, int n_images, int device, void* stream);

/* Test-render export: to8b (nerf_renderer_helper.py:10) of n_maps (<= 32) fp32 device maps into one
 * packed uint8 atlas (map k at offset sum(sizes[:k])).  transforms[k]: 0 identity, 1 (x+1)/2 (normal maps),
 * 2 1/max(1e-10, x/scales[k]) (depth maps), as in ibl_nerf_renderer.py:840-848.  maps/sizes/transforms/scales are
 * HOST arrays. */
int ibln_pack_u8(const float* const* maps, const int64_t* sizes, const int* transforms, const float* scales, int n_maps,
                 uint8_t* out, int device, void* stream);

/* Normal map from a rendered depth image for the test-render export:
 * utils/depth_to_normal_utils.py:9-46 (depth_to_position + depth_to_normal_image_space, called at
 * ibl_nerf_renderer.py:903-906).  depth [height,width] device fp32; c2w HOST [3,4] row-major; intrinsics fx, fy, cx,
 * cy; normal [height,width,3] device fp32.  Edge pixels replicate their neighbours (np.pad 'edge'). */
int ibln_depth_to_normal(const float* depth, int height, int width, float fx, float fy, float cx, float cy,
                         const float* c2w_host, float* normal, int device, void* stream);

/* Self-test of the tcgen05 building block: D[128,N] = A[128,K] * B[N,K]^T with bf16 inputs staged
 * through the same swizzled shared-memory layout the MLP kernels use. a,b fp32 (rounded to bf16
 * inside), d fp32.  variant selects descriptor hypotheses (0 = production). */
int ibln_umma_selftest(const float* a, const float* b, float* d, int n, int k, int variant, int device, void* stream);

/* Self-test of the MN-major operand path used by the wgrad kernel: D[128,N] = X^T Y with X [128 points,128],
 * Y [128 points,N] (N in {64,128,192,256}) staged as swizzled operand tiles and read as MN-major. */
int ibln_umma_mn_selftest(const float* x, const float* y, float* d, int n, int device, void* stream);
/* cta_group::2 operand/commit self-test: D[256,256] = A[256,K] * B[256,K]^T on one CTA pair (K in {64..256}). */
int ibln_umma_pair_selftest(const float* a, const float* b, float* d, int k, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IBLNERF_B200_H */

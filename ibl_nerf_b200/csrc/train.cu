// Training-step tail kernels: the phase-gated image losses of src/train.py and the Adam update of
// both networks, each as ONE launch over flat buffers instead of ~60 small ATen launches + ~100 per-tensor updates;
// ray generation / target gather, export packing, and the ABI bookkeeping entry points.
#include "common.cuh"

namespace ibln {

// The phase-gated image losses of train.py:299-441 for ONE pass (fine or coarse) on the packed, gamma-corrected kernel
// outputs, forward and backward in one launch.  With mse = mean over all elements (img2mse, nerf_renderer_helper.py:8):
//   *loss += scale * [ w_rad   * ( mse(radiance_map, rgb) + sum_k mse(radiance_map_k, rgb_k) )     train.py:326-334, 420-423
//                    + w_color * mse(color_map, rgb)                                               :323, 437-438 (full-IBL phase)
//                    + w_alb   * mse(albedo_map, prior_albedo)                                     :401-403, 444-446 (prior phase)
//                    + w_irr   * mse(irradiance_map, irr_target) ]                                 :410-412, 447 (fine pass only)
// maps cols: irradiance 5, albedo 6..8, radiance 9..11, coarse radiance k 12+3k..; shade cols: colour 10..12.
struct LossWeights { float rad, color, alb, irr, irr_target, scale; };
constexpr int LOSS_THREADS = 256;
__global__ void __launch_bounds__(LOSS_THREADS)
image_losses_kernel(const float* __restrict__ maps, const float* __restrict__ shade, const float* __restrict__ rgb,
                    const float* __restrict__ rgb1, const float* __restrict__ rgb2, const float* __restrict__ rgb3,
                    const float* __restrict__ prior_albedo, int n, LossWeights w, float* __restrict__ loss,
                    float* __restrict__ g_maps, float* __restrict__ g_shade) {
  const int r = blockIdx.x * LOSS_THREADS + threadIdx.x;
  float acc = 0.f;
  if (r < n) {
    const float inv3 = w.scale / (3.0f * (float)n), inv1 = w.scale / (float)n;
    const float* m = maps + (size_t)r * 24;
    float g[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) g[j] = 0.f;
    const float* tg[4] = {rgb, rgb1, rgb2, rgb3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (tg[k] == nullptr || w.rad == 0.f) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = m[9 + 3 * k + c] - tg[k][(size_t)r * 3 + c];
        acc += w.rad * inv3 * d * d;
        g[9 + 3 * k + c] = 2.0f * w.rad * inv3 * d;
      }
    }
    if (prior_albedo != nullptr && w.alb != 0.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = m[6 + c] - prior_albedo[(size_t)r * 3 + c];
        acc += w.alb * inv3 * d * d;
        g[6 + c] = 2.0f * w.alb * inv3 * d;
      }
    }
    if (w.irr != 0.f) {
      const float d = m[5] - w.irr_target;
      acc += w.irr * inv1 * d * d;
      g[5] = 2.0f * w.irr * inv1 * d;
    }
    float4* gm = reinterpret_cast<float4*>(g_maps + (size_t)r * 24);
#pragma unroll
    for (int q = 0; q < 6; ++q) gm[q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
    if (shade != nullptr) {
      const float* s = shade + (size_t)r * 16;
      float gs[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) gs[j] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = s[10 + c] - rgb[(size_t)r * 3 + c];
        acc += w.color * inv3 * d * d;
        gs[10 + c] = 2.0f * w.color * inv3 * d;
      }
      float4* gp = reinterpret_cast<float4*>(g_shade + (size_t)r * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) gp[q] = make_float4(gs[4 * q], gs[4 * q + 1], gs[4 * q + 2], gs[4 * q + 3]);
    }
  }
  __shared__ float part[LOSS_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < LOSS_THREADS / 32 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v);
  }
}

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False) on flat buffers, train.py:479-498:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float step_size, float b1, float b2, float eps, float inv_sqrt_bc2, float grad_scale) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4), gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * grad_scale;
      M[k] = __fmaf_rn(b1, M[k], (1.0f - b1) * gr);           // lerp form of torch: m + (g - m)(1 - b1)
      V[k] = __fmaf_rn(b2, V[k], (1.0f - b2) * gr * gr);
      const float denom = sqrtf(V[k]) * inv_sqrt_bc2 + eps;
      P[k] -= step_size * (M[k] / denom);
    }
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float gr = g[i] * grad_scale;
      const float mi = __fmaf_rn(b1, m[i], (1.0f - b1) * gr);
      const float vi = __fmaf_rn(b2, v[i], (1.0f - b2) * gr * gr);
      m[i] = mi; v[i] = vi;
      p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
  }
}

// Gradient all-reduce FUSED into the optimizer (data-parallel training, one process per GPU): the flat gradient buffer
// of every rank lives in symmetric memory (torch.distributed._symmetric_memory: the same virtual layout on all ranks,
// mapped into every peer and bound to one NVSwitch multicast object).  Each rank then runs this ONE kernel:
//   MC = true : g = multimem.ld_reduce.add.v4.f32 [multicast address]  -- the switch sums the ranks' values in flight
//               (NVLS), the GPU receives the reduced gradient once, 16 bytes per load
//   MC = false: g = sum over the ranks' peer-mapped pointers (plain P2P loads over NVLink), for fabrics without multicast
// and applies Adam to its replica -- no NCCL launch, no reduced-gradient round trip through HBM.  The caller brackets it
// with the symmetric-memory barrier (all gradients written / all ranks done reading).
struct PeerGrads { const float* p[8]; int world; };
template <bool MC>
__global__ void __launch_bounds__(256)
adam_allreduce_kernel(float* __restrict__ p, const float* g_mc, PeerGrads peers, float* __restrict__ m, float* __restrict__ v,
                      int64_t n, float step_size, float b1, float b2, float eps, float inv_sqrt_bc2, float grad_scale) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float4 gg;
  if (MC) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(gg.x), "=f"(gg.y), "=f"(gg.z), "=f"(gg.w) : "l"(g_mc + i4) : "memory");
  } else {
    gg = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < peers.world; ++r) {
      float4 t;
      asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(peers.p[r] + i4) : "memory");
      gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
    }
  }
  float4 pp = *reinterpret_cast<float4*>(p + i4);
  float4 mm = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
  float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gr = G[k] * grad_scale;
    M[k] = __fmaf_rn(b1, M[k], (1.0f - b1) * gr);
    V[k] = __fmaf_rn(b2, V[k], (1.0f - b2) * gr * gr);
    const float denom = sqrtf(V[k]) * inv_sqrt_bc2 + eps;
    P[k] -= step_size * (M[k] / denom);
  }
  *reinterpret_cast<float4*>(p + i4) = pp;
  *reinterpret_cast<float4*>(m + i4) = mm;
  *reinterpret_cast<float4*>(v + i4) = vv;
}

}  // namespace ibln

using namespace ibln;

extern "C" int ibln_adam_allreduce_step(float* param, const float* grad_multicast, const float* const* peer_grads_host, int world,
                                        float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                                        float eps, int step, float grad_scale, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || (n & 3) || !param || !exp_avg || !exp_avg_sq || step < 1 || world < 1 || world > 8) return IBLN_EINVAL;
  if (!grad_multicast && !peer_grads_host) return IBLN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad_multicast) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) return IBLN_EINVAL;
  DeviceGuard g(device);
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const unsigned blocks = (unsigned)((n / 4 + 255) / 256);
  PeerGrads peers;
  peers.world = world;
  for (int r = 0; r < 8; ++r) {
    peers.p[r] = (peer_grads_host && r < world) ? peer_grads_host[r] : nullptr;
    if (!grad_multicast && r < world && (!peers.p[r] || (reinterpret_cast<uintptr_t>(peers.p[r]) & 15))) return IBLN_EINVAL;
  }
  if (grad_multicast)
    adam_allreduce_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad_multicast, peers, exp_avg, exp_avg_sq, n, step_size,
                                                                          beta1, beta2, eps, inv_sqrt_bc2, grad_scale);
  else
    adam_allreduce_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(param, nullptr, peers, exp_avg, exp_avg_sq, n, step_size,
                                                                           beta1, beta2, eps, inv_sqrt_bc2, grad_scale);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_image_losses(const float* maps_srgb, const float* shade_srgb, const float* rgb, const float* rgb_1,
                                 const float* rgb_2, const float* rgb_3, const float* prior_albedo, int n, float w_radiance,
                                 float w_color, float w_prior_albedo, float w_irradiance_reg, float irradiance_target,
                                 float scale, float* loss, float* g_maps, float* g_shade, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !maps_srgb || !rgb || !loss || !g_maps || (shade_srgb && !g_shade)) return IBLN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(g_maps) | reinterpret_cast<uintptr_t>(g_shade)) & 15) return IBLN_EINVAL;
  DeviceGuard g(device);
  LossWeights w = {w_radiance, w_color, w_prior_albedo, w_irradiance_reg, irradiance_target, scale};
  image_losses_kernel<<<(n + LOSS_THREADS - 1) / LOSS_THREADS, LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      maps_srgb, shade_srgb, rgb, rgb_1, rgb_2, rgb_3, prior_albedo, n, w, loss, g_maps, g_shade);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_zero(void* buf, int64_t bytes, int device, void* stream) {
  if (bytes == 0) return 0;
  if (!buf || bytes < 0) return IBLN_EINVAL;
  DeviceGuard g(device);
  IBLN_CUDA(cudaMemsetAsync(buf, 0, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}

extern "C" int ibln_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, int step, float grad_scale, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !param || !grad || !exp_avg || !exp_avg_sq || step < 1) return IBLN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) return IBLN_EINVAL;
  DeviceGuard g(device);
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int64_t threads = (n + 3) / 4;
  adam_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, step_size,
                                                                                  beta1, beta2, eps, inv_sqrt_bc2, grad_scale);
  IBLN_RETURN_LAST();
}

// ---------------------------------------------------------------- ray generation + target gather
// One launch per training batch instead of ~15 small ATen launches: get_rays_few (nerf_renderer_helper.py:14-23)
// for N random pixels of one image + NerfDataset.get_info pixel gathers (dataset_interface.py:178-197) of the
// target image and its prefiltered copies.
namespace ibln {
struct GatherArgs { const float* img[12]; float* out[12]; int ch[12]; int n_img; };
__global__ void __launch_bounds__(256)
sample_rays_kernel(const int* __restrict__ u, const int* __restrict__ v, int n, int H, int W, float fx, float fy, float cx,
                   float cy, const float* __restrict__ c2w /* [3,4] row-major */, float* __restrict__ rays_o,
                   float* __restrict__ rays_d, GatherArgs ga) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int ui = u[r], vi = v[r];
  // dirs = ((i - cx)/fx, -(j - cy)/fy, -1);  rays_d = sum_k dirs[k] * c2w[:, k]  (un-contracted, summed k = 0,1,2)
  const float d0 = __fdiv_rn(__fsub_rn((float)ui, cx), fx);
  const float d1 = -__fdiv_rn(__fsub_rn((float)vi, cy), fy);
  const float d2 = -1.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(d0, c2w[4 * a]), __fmul_rn(d1, c2w[4 * a + 1])), __fmul_rn(d2, c2w[4 * a + 2]));
    rays_d[3 * r + a] = s;
    rays_o[3 * r + a] = c2w[4 * a + 3];
  }
  const bool inside = ui >= 0 && ui < W && vi >= 0 && vi < H;
  for (int k = 0; k < ga.n_img; ++k) {
    const int ch = ga.ch[k];
    const float* src = ga.img[k] + ((size_t)vi * W + ui) * ch;      // height is first (image[v, u, :])
    for (int c = 0; c < ch; ++c) ga.out[k][(size_t)r * ch + c] = inside ? src[c] : 0.f;
  }
}
}  // namespace ibln

extern "C" int ibln_sample_rays(const int* u, const int* v, int n, int height, int width, float fx, float fy, float cx, float cy,
                                const float* c2w, float* rays_o, float* rays_d, const float* const* images,
                                float* const* outputs, const int* channels, int n_images, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !u || !v || !c2w || !rays_o || !rays_d || n_images < 0 || n_images > 12 || height < 1 || width < 1) return IBLN_EINVAL;
  if (n_images > 0 && (!images || !outputs || !channels)) return IBLN_EINVAL;
  ibln::GatherArgs ga;
  ga.n_img = n_images;
  for (int k = 0; k < n_images; ++k) {
    if (!images[k] || !outputs[k] || channels[k] < 1) return IBLN_EINVAL;
    ga.img[k] = images[k]; ga.out[k] = outputs[k]; ga.ch[k] = channels[k];
  }
  DeviceGuard g(device);
  ibln::sample_rays_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(u, v, n, height, width, fx, fy, cx, cy, c2w, rays_o,
                                                                             rays_d, ga);
  IBLN_RETURN_LAST();
}

// ---------------------------------------------------------------- test-render export
// to8b (nerf_renderer_helper.py:10) of all output maps of an image into ONE packed uint8 atlas: one launch and one
// D2H copy instead of ~30 .cpu().numpy() + host conversions per image (ibl_nerf_renderer.py:840-900).
// transform 0: identity; 1: (x + 1) / 2 (normal / tangent maps); 2: 1 / max(1e-10, x / scale) (depth maps).
namespace ibln {
struct PackArgsU8 { const float* src[32]; long long off[33]; int transform[32]; float scale[32]; int n; };
__global__ void __launch_bounds__(256) pack_u8_kernel(PackArgsU8 a, uint8_t* __restrict__ out) {
  const long long total = a.off[a.n];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= a.off[k + 1]) ++k;
    float x = a.src[k][i - a.off[k]];
    if (a.transform[k] == 1) x = __fmul_rn(__fadd_rn(x, 1.0f), 0.5f);
    else if (a.transform[k] == 2) x = __fdiv_rn(1.0f, fmaxf(1e-10f, __fdiv_rn(x, a.scale[k])));
    const float c = fminf(fmaxf(x, 0.0f), 1.0f);            // np.clip(x, 0, 1); NaN -> 0 after the cast, like numpy on x86
    out[i] = (uint8_t)__fmul_rn(255.0f, c);                 // .astype(np.uint8) truncates
  }
}
}  // namespace ibln

extern "C" int ibln_pack_u8(const float* const* maps, const int64_t* sizes, const int* transforms, const float* scales, int n_maps,
                            uint8_t* out, int device, void* stream) {
  if (n_maps == 0) return 0;
  if (n_maps < 0 || n_maps > 32 || !maps || !sizes || !transforms || !scales || !out) return IBLN_EINVAL;
  ibln::PackArgsU8 a;
  a.n = n_maps;
  a.off[0] = 0;
  for (int k = 0; k < n_maps; ++k) {
    if (!maps[k] || sizes[k] < 0 || transforms[k] < 0 || transforms[k] > 2) return IBLN_EINVAL;
    a.src[k] = maps[k]; a.off[k + 1] = a.off[k] + sizes[k]; a.transform[k] = transforms[k]; a.scale[k] = scales[k];
  }
  if (a.off[n_maps] == 0) return 0;
  DeviceGuard g(device);
  long long blocks = (a.off[n_maps] + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ibln::pack_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, out);
  IBLN_RETURN_LAST();
}

// ---------------------------------------------------------------- normal map from a rendered depth image
// utils/depth_to_normal_utils.py:9-46 (depth_to_position + depth_to_normal_image_space), which the export path calls
// once per test image (ibl_nerf_renderer.py:903-906) on the host in numpy: unit camera rays * depth -> world
// positions, edge-replicated central differences along x and y, normalise, cross(vb, va), normalise.
// One thread per pixel; the four neighbour positions are recomputed from their depths (4 loads instead of a
// position image round trip).  Degenerate differences give NaN exactly like the numpy 0/0.
namespace ibln {
struct Cam { float fx, fy, cx, cy; float r[9]; float t[3]; };
__device__ __forceinline__ void pixel_position(const Cam& c, const float* __restrict__ depth, int width, int x, int y, float (&p)[3]) {
  float d0 = __fdiv_rn((float)x - c.cx, c.fx), d1 = -__fdiv_rn((float)y - c.cy, c.fy), d2 = -1.0f;
  const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), 1.0f)), 1e-12f);   // F.normalize
  d0 = __fdiv_rn(d0, nrm); d1 = __fdiv_rn(d1, nrm); d2 = __fdiv_rn(d2, nrm);
  const float z = depth[(long long)y * width + x];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float rd = __fadd_rn(__fadd_rn(__fmul_rn(d0, c.r[3 * i]), __fmul_rn(d1, c.r[3 * i + 1])), __fmul_rn(d2, c.r[3 * i + 2]));
    p[i] = __fadd_rn(c.t[i], __fmul_rn(rd, z));
  }
}
__device__ __forceinline__ void unit3(float (&v)[3]) {
  const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = __fdiv_rn(v[i], n);
}
__global__ void __launch_bounds__(256) depth_to_normal_kernel(Cam c, const float* __restrict__ depth, int height, int width,
                                                              float* __restrict__ normal) {
  const long long total = (long long)height * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / width), x = (int)(i - (long long)y * width);
    float l[3], r[3], u[3], b[3];
    pixel_position(c, depth, width, max(x - 1, 0), y, l);
    pixel_position(c, depth, width, min(x + 1, width - 1), y, r);
    pixel_position(c, depth, width, x, max(y - 1, 0), u);
    pixel_position(c, depth, width, x, min(y + 1, height - 1), b);
    float va[3], vb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { va[k] = __fsub_rn(r[k], l[k]); vb[k] = __fsub_rn(b[k], u[k]); }
    unit3(va); unit3(vb);
    float vc[3] = {__fsub_rn(__fmul_rn(vb[1], va[2]), __fmul_rn(vb[2], va[1])),
                   __fsub_rn(__fmul_rn(vb[2], va[0]), __fmul_rn(vb[0], va[2])),
                   __fsub_rn(__fmul_rn(vb[0], va[1]), __fmul_rn(vb[1], va[0]))};
    unit3(vc);
    normal[3 * i] = vc[0]; normal[3 * i + 1] = vc[1]; normal[3 * i + 2] = vc[2];
  }
}
}  // namespace ibln

extern "C" int ibln_depth_to_normal(const float* depth, int height, int width, float fx, float fy, float cx, float cy,
                                    const float* c2w_host, float* normal, int device, void* stream) {
  if (height < 0 || width < 0) return IBLN_EINVAL;
  if (height == 0 || width == 0) return 0;
  if (!depth || !c2w_host || !normal) return IBLN_EINVAL;
  ibln::Cam c;
  c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c.r[3 * i + j] = c2w_host[4 * i + j];
    c.t[i] = c2w_host[4 * i + 3];
  }
  DeviceGuard g(device);
  long long blocks = ((long long)height * width + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ibln::depth_to_normal_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(c, depth, height, width, normal);
  IBLN_RETURN_LAST();
}

// ---------------------------------------------------------------- ABI bookkeeping
extern "C" int ibln_abi_version(void) { return IBLN_ABI_VERSION; }

extern "C" const char* ibln_error_string(int code) {
  if (code == 0) return "success";
  if (code == IBLN_EINVAL) return "iblnerf_b200: invalid argument";
  return cudaGetErrorString((cudaError_t)code);
}

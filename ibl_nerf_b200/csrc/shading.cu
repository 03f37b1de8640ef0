// Normal-from-depth (epsilon) helpers and split-sum image-based-lighting shading, forward and
// backward (north-star subsystem 5).  Per-ray streaming work: one thread per ray, the 2 MB BRDF LUT
// stays L2 resident; bilinear filtering is done by hand in fp32 (hardware texture filtering has 8-bit
// weights and would miss the 1e-4 criterion).
#include "common.cuh"

namespace ibln {

// right = d x (0,1,0), up = right x d  (normal_from_depth.py:143-147, not normalised)
__device__ __forceinline__ void eps_frame(const float d[3], float right[3], float up[3]) {
  right[0] = -d[2]; right[1] = 0.f; right[2] = d[0];
  up[0] = __fsub_rn(__fmul_rn(right[1], d[2]), __fmul_rn(right[2], d[1]));
  up[1] = __fsub_rn(__fmul_rn(right[2], d[0]), __fmul_rn(right[0], d[2]));
  up[2] = __fsub_rn(__fmul_rn(right[0], d[1]), __fmul_rn(right[1], d[0]));
}

__global__ void normal_eps_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                         const float* __restrict__ z, int n, int S, float eps, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)n * S;
  if (idx >= tot) return;
  int r = (int)(idx / S);
  float dd[3] = {d[3 * r], d[3 * r + 1], d[3 * r + 2]}, right[3], up[3];
  eps_frame(dd, right, up);
  float zi = z[idx];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p = __fadd_rn(o[3 * r + c], __fmul_rn(dd[c], zi));
    float er = __fmul_rn(eps, right[c]), eu = __fmul_rn(eps, up[c]);
    out[(0 * tot + idx) * 3 + c] = __fadd_rn(p, er);
    out[(1 * tot + idx) * 3 + c] = __fsub_rn(p, er);
    out[(2 * tot + idx) * 3 + c] = __fadd_rn(p, eu);
    out[(3 * tot + idx) * 3 + c] = __fsub_rn(p, eu);
  }
}

__global__ void normal_eps_finish_kernel(const float* __restrict__ d, const float* __restrict__ depths4, int n, float eps,
                                         float* __restrict__ normal, float* __restrict__ refl,
                                         const float* __restrict__ o, const float* __restrict__ depth, int depth_ld,
                                         float* __restrict__ x_surface) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float dd[3] = {d[3 * r], d[3 * r + 1], d[3 * r + 2]}, right[3], up[3];
  if (x_surface != nullptr) {      // x_surface = rays_o + rays_d * depth   (ibl_nerf_renderer.py:262)
    const float t = depth[(size_t)r * depth_ld];
#pragma unroll
    for (int c = 0; c < 3; ++c) x_surface[3 * r + c] = __fadd_rn(o[3 * r + c], __fmul_rn(dd[c], t));
  }
  eps_frame(dd, right, up);
  float ddx = depths4[r] - depths4[n + r];
  float ddy = depths4[2 * n + r] - depths4[3 * n + r];
  float dx[3], dy[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dx[c] = 2.f * eps * right[c] + ddx * dd[c];
    dy[c] = 2.f * eps * up[c] + ddy * dd[c];
  }
  float nx = dx[1] * dy[2] - dx[2] * dy[1];
  float ny = dx[2] * dy[0] - dx[0] * dy[2];
  float nz = dx[0] * dy[1] - dx[1] * dy[0];
  float len = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);   // F.normalize eps
  nx /= len; ny /= len; nz /= len;
  normal[3 * r] = nx; normal[3 * r + 1] = ny; normal[3 * r + 2] = nz;
  if (refl != nullptr) {
    float nd = nx * dd[0] + ny * dd[1] + nz * dd[2];
    refl[3 * r] = dd[0] - 2.f * nd * nx;
    refl[3 * r + 1] = dd[1] - 2.f * nd * ny;
    refl[3 * r + 2] = dd[2] - 2.f * nd * nz;
  }
}

struct LutFetch { float a, b, da_dy, db_dy; };

// F.grid_sample(bilinear, zeros, align_corners=True) at (x = n.v, y = roughness) + d/dy in pixels
__device__ __forceinline__ LutFetch lut_fetch(const float* __restrict__ lut, int H, int W, float ndv, float rough) {
  float ix = ((2.f * ndv - 1.f + 1.f) / 2.f) * (float)(W - 1);
  float iy = ((2.f * rough - 1.f + 1.f) / 2.f) * (float)(H - 1);
  float x0 = floorf(ix), y0 = floorf(iy);
  float wx1 = ix - x0, wx0 = (x0 + 1.f) - ix, wy1 = iy - y0, wy0 = (y0 + 1.f) - iy;
  int xi = (int)x0, yi = (int)y0;
  LutFetch f = {0.f, 0.f, 0.f, 0.f};
  const float* ch0 = lut;
  const float* ch1 = lut + (size_t)H * W;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      int xx = xi + dx, yy = yi + dy;
      if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
        float wx = dx ? wx1 : wx0, wy = dy ? wy1 : wy0;
        float v0 = __ldg(ch0 + (size_t)yy * W + xx), v1 = __ldg(ch1 + (size_t)yy * W + xx);
        f.a += v0 * wx * wy; f.b += v1 * wx * wy;
        float s = dy ? 1.f : -1.f;
        f.da_dy += s * v0 * wx; f.db_dy += s * v1 * wx;
      }
    }
  return f;
}

struct ShadeIn {
  float d[3], n[3], alb[3], rough, irr, mip_rough, depth, nearv, farv;
};

struct ShadeMid {
  float ndv, p5, F0[3], Fr[3], s0[3], pre[3], diff[3], spec[3], color[3], rem, lvl_raw;
  int i1, i2, above[3];
  LutFetch lf;
};

// ld3 / ld1: row stride (floats) of albedo and of the per-ray scalars rough / irr / mip_rough / depth: 3 / 1 for
// separate tensors, IBLN_MAPS_STRIDE for columns of the packed compositing output.
__device__ __forceinline__ ShadeIn load_in(const float* rays_d, const float* normal, const float* albedo, const float* rough,
                                           const float* irr, const float* mip_rough, const float* depth,
                                           const float* nearp, const float* farp, int r, int ld3, int ld1) {
  ShadeIn s;
#pragma unroll
  for (int c = 0; c < 3; ++c) { s.d[c] = rays_d[3 * r + c]; s.n[c] = normal[3 * r + c]; s.alb[c] = albedo[(size_t)ld3 * r + c]; }
  s.rough = rough[(size_t)ld1 * r]; s.irr = irr[(size_t)ld1 * r]; s.mip_rough = mip_rough[(size_t)ld1 * r];
  s.depth = depth[(size_t)ld1 * r]; s.nearv = nearp[r]; s.farv = farp[r];
  return s;
}

__device__ __forceinline__ ShadeMid shade_core(const ShadeIn& in, const float* __restrict__ pref, int n_pref,
                                               const float* __restrict__ lut, int H, int W, int lut_coef, int correct_depth) {
  ShadeMid m;
  float ndv = -(in.d[0] * in.n[0]) - (in.d[1] * in.n[1]) - (in.d[2] * in.n[2]);
  m.ndv = fminf(fmaxf(ndv, 0.f), 1.f);
  m.lf = lut_fetch(lut, H, W, m.ndv, in.rough);
  float metallic = 1.f - in.rough;
  float omc = fminf(fmaxf(1.f - m.ndv, 0.f), 1.f);
  m.p5 = omc * omc * omc * omc * omc;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    m.F0[c] = 0.04f * (1.f - metallic) + in.alb[c] * metallic;
    float mx = fmaxf(1.f - in.rough, m.F0[c]);
    m.above[c] = (1.f - in.rough > m.F0[c]) ? 1 : ((1.f - in.rough < m.F0[c]) ? -1 : 0);
    m.Fr[c] = m.F0[c] + (mx - m.F0[c]) * m.p5;
    m.s0[c] = (lut_coef == 0 ? m.Fr[c] : m.F0[c]) * m.lf.a + m.lf.b;
  }
  float lvl = in.mip_rough;
  m.lvl_raw = lvl;
  if (correct_depth) {
    m.lvl_raw = in.mip_rough * in.depth / ((in.farv + in.nearv) * 0.5f);
    lvl = fminf(fmaxf(m.lvl_raw, 0.f), 1.f);
  }
  float t = lvl * (float)(n_pref - 1);
  long long i1 = (long long)t;     // .long() truncation
  i1 = i1 < 0 ? 0 : (i1 > n_pref - 1 ? n_pref - 1 : i1);
  m.i1 = (int)i1;
  m.i2 = m.i1 + 1 > n_pref - 1 ? n_pref - 1 : m.i1 + 1;
  m.rem = t - (float)m.i1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    m.pre[c] = (1.f - m.rem) * pref[m.i1 * 3 + c] + m.rem * pref[m.i2 * 3 + c];
    m.diff[c] = (1.f - m.Fr[c]) * (1.f - metallic) * in.alb[c] * in.irr;
    m.spec[c] = m.s0[c] * m.pre[c];
    m.color[c] = m.diff[c] + m.spec[c];
  }
  return m;
}

__global__ void shade_fwd_kernel(const float* __restrict__ rays_d, const float* __restrict__ normal,
                                 const float* __restrict__ albedo, const float* __restrict__ rough,
                                 const float* __restrict__ irr, const float* __restrict__ mip_rough,
                                 const float* __restrict__ depth, const float* __restrict__ nearp,
                                 const float* __restrict__ farp, const float* __restrict__ pref, int n_pref,
                                 const float* __restrict__ lut, int H, int W, int lut_coef, int correct_depth, int n,
                                 float* __restrict__ out, float* __restrict__ out_srgb, int ld3, int ld1) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  ShadeIn in = load_in(rays_d, normal, albedo, rough, irr, mip_rough, depth, nearp, farp, r, ld3, ld1);
  ShadeMid m = shade_core(in, pref + (size_t)r * n_pref * 3, n_pref, lut, H, W, lut_coef, correct_depth);
  float o[IBLN_SHADE_STRIDE];
  o[IBLN_SH_NDV] = m.ndv;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[IBLN_SH_SPEC + c] = m.spec[c]; o[IBLN_SH_DIFF + c] = m.diff[c]; o[IBLN_SH_PRE + c] = m.pre[c];
    o[IBLN_SH_COLOR + c] = m.color[c];
  }
  o[13] = o[14] = o[15] = 0.f;
  float4* dst = reinterpret_cast<float4*>(out + (size_t)r * IBLN_SHADE_STRIDE);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  if (out_srgb != nullptr) {
#pragma unroll
    for (int c = 1; c < 13; ++c) o[c] = srgbf(o[c]);
    float4* ds = reinterpret_cast<float4*>(out_srgb + (size_t)r * IBLN_SHADE_STRIDE);
#pragma unroll
    for (int q = 0; q < 4; ++q) ds[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
}

__global__ void shade_bwd_kernel(const float* __restrict__ rays_d, const float* __restrict__ normal,
                                 const float* __restrict__ albedo, const float* __restrict__ rough,
                                 const float* __restrict__ irr, const float* __restrict__ mip_rough,
                                 const float* __restrict__ depth, const float* __restrict__ nearp,
                                 const float* __restrict__ farp, const float* __restrict__ pref, int n_pref,
                                 const float* __restrict__ lut, int H, int W, int lut_coef, int correct_depth, int n,
                                 const float* __restrict__ g_out, const float* __restrict__ g_srgb,
                                 float* __restrict__ g_albedo, float* __restrict__ g_rough, float* __restrict__ g_irr,
                                 float* __restrict__ g_mip, int ld3, int ld1, float* __restrict__ g_maps) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  ShadeIn in = load_in(rays_d, normal, albedo, rough, irr, mip_rough, depth, nearp, farp, r, ld3, ld1);
  const float* P = pref + (size_t)r * n_pref * 3;
  ShadeMid m = shade_core(in, P, n_pref, lut, H, W, lut_coef, correct_depth);
  float G[IBLN_SHADE_STRIDE];
#pragma unroll
  for (int c = 0; c < IBLN_SHADE_STRIDE; ++c) G[c] = g_out ? g_out[(size_t)r * IBLN_SHADE_STRIDE + c] : 0.f;
  if (g_srgb != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      G[IBLN_SH_SPEC + c] += g_srgb[(size_t)r * IBLN_SHADE_STRIDE + IBLN_SH_SPEC + c] * dsrgbf(m.spec[c]);
      G[IBLN_SH_DIFF + c] += g_srgb[(size_t)r * IBLN_SHADE_STRIDE + IBLN_SH_DIFF + c] * dsrgbf(m.diff[c]);
      G[IBLN_SH_PRE + c] += g_srgb[(size_t)r * IBLN_SHADE_STRIDE + IBLN_SH_PRE + c] * dsrgbf(m.pre[c]);
      G[IBLN_SH_COLOR + c] += g_srgb[(size_t)r * IBLN_SHADE_STRIDE + IBLN_SH_COLOR + c] * dsrgbf(m.color[c]);
    }
  }
  float rho = in.rough, I = in.irr;
  float gA = 0.f, gB = 0.f, gRho = 0.f, gI = 0.f, gRem = 0.f;
  float gAlb[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float Gspec = G[IBLN_SH_SPEC + c] + G[IBLN_SH_COLOR + c];
    float Gdiff = G[IBLN_SH_DIFF + c] + G[IBLN_SH_COLOR + c];
    float Gpre = G[IBLN_SH_PRE + c] + Gspec * m.s0[c];
    float Gs0 = Gspec * m.pre[c];
    float X = lut_coef == 0 ? m.Fr[c] : m.F0[c];
    gA += Gs0 * X; gB += Gs0;
    float GFr = Gdiff * (-rho * in.alb[c] * I) + (lut_coef == 0 ? Gs0 * m.lf.a : 0.f);
    float GF0 = (lut_coef == 0 ? 0.f : Gs0 * m.lf.a);
    // Fr = F0 + (max(1-rho, F0) - F0) p5
    float dFr_dF0, dFr_drho;
    if (m.above[c] > 0) { dFr_dF0 = 1.f - m.p5; dFr_drho = -m.p5; }
    else if (m.above[c] < 0) { dFr_dF0 = 1.f; dFr_drho = 0.f; }
    else { dFr_dF0 = 1.f - 0.5f * m.p5; dFr_drho = -0.5f * m.p5; }
    GF0 += GFr * dFr_dF0;
    gRho += GFr * dFr_drho;
    // F0 = 0.04 rho + alb (1 - rho)
    gRho += GF0 * (0.04f - in.alb[c]);
    gAlb[c] = GF0 * (1.f - rho) + Gdiff * (1.f - m.Fr[c]) * rho * I;
    gRho += Gdiff * (1.f - m.Fr[c]) * in.alb[c] * I;
    gI += Gdiff * (1.f - m.Fr[c]) * rho * in.alb[c];
    gRem += Gpre * (P[m.i2 * 3 + c] - P[m.i1 * 3 + c]);
  }
  gRho += (gA * m.lf.da_dy + gB * m.lf.db_dy) * (float)(H - 1);
  float gLvl = gRem * (float)(n_pref - 1);
  float gMip;
  if (correct_depth) {
    bool inside = (m.lvl_raw >= 0.f) && (m.lvl_raw <= 1.f);
    gMip = inside ? gLvl * in.depth / ((in.farv + in.nearv) * 0.5f) : 0.f;
  } else {
    gMip = gLvl;
  }
  if (g_maps != nullptr) {
    // packed form: one row of d loss / d maps (linear compositing outputs); roughness_map feeds both the LUT / Fresnel
    // terms and the mip level (ibl_nerf_renderer.py:324, 459), the depth is detached (:458)
    float row[IBLN_MAPS_STRIDE];
#pragma unroll
    for (int c = 0; c < IBLN_MAPS_STRIDE; ++c) row[c] = 0.f;
    row[IBLN_MAP_ROUGH] = gRho + gMip; row[IBLN_MAP_IRR] = gI;
#pragma unroll
    for (int c = 0; c < 3; ++c) row[IBLN_MAP_ALBEDO + c] = gAlb[c];
    float4* dst = reinterpret_cast<float4*>(g_maps + (size_t)r * IBLN_MAPS_STRIDE);
#pragma unroll
    for (int q = 0; q < IBLN_MAPS_STRIDE / 4; ++q) dst[q] = make_float4(row[4 * q], row[4 * q + 1], row[4 * q + 2], row[4 * q + 3]);
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) g_albedo[3 * r + c] = gAlb[c];
  g_rough[r] = gRho; g_irr[r] = gI; g_mip[r] = gMip;
}

}  // namespace ibln

using namespace ibln;

extern "C" int ibln_normal_eps_points(const float* rays_o, const float* rays_d, const float* z, int n, int S, float eps,
                                      float* pts_out, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || S < 1 || !rays_o || !rays_d || !z || !pts_out) return IBLN_EINVAL;
  DeviceGuard g(device);
  int64_t tot = (int64_t)n * S;
  normal_eps_points_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, n, S, eps, pts_out);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_normal_eps_finish(const float* rays_d, const float* depths4, int n, float eps, float* normal,
                                      float* refl, const float* rays_o, const float* depth, int depth_ld,
                                      float* x_surface, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !rays_d || !depths4 || !normal || (x_surface && (!rays_o || !depth || depth_ld < 1))) return IBLN_EINVAL;
  DeviceGuard g(device);
  normal_eps_finish_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rays_d, depths4, n, eps, normal, refl, rays_o,
                                                                             depth, depth_ld, x_surface);
  IBLN_RETURN_LAST();
}

#define SHADE_ARGS_OK (n >= 0 && n_pref >= 1 && lut_h >= 2 && lut_w >= 2 && rays_d && normal && albedo && rough && irr && \
                       mip_rough && depth && nearp && farp && prefiltered && lut && (lut_coef == 0 || lut_coef == 1))

extern "C" int ibln_shade_fwd(const float* rays_d, const float* normal, const float* albedo, const float* rough,
                              const float* irr, const float* mip_rough, const float* depth, const float* nearp,
                              const float* farp, const float* prefiltered, int n_pref, const float* lut, int lut_h,
                              int lut_w, int lut_coef, int correct_depth, int n, float* out, float* out_srgb, int device,
                              void* stream) {
  if (n == 0) return 0;
  if (!SHADE_ARGS_OK || !out) return IBLN_EINVAL;
  DeviceGuard g(device);
  shade_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rays_d, normal, albedo, rough, irr, mip_rough, depth,
                                                                     nearp, farp, prefiltered, n_pref, lut, lut_h, lut_w,
                                                                     lut_coef, correct_depth, n, out, out_srgb, 3, 1);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_shade_fwd_maps(const float* rays_d, const float* normal, const float* maps, const float* nearp,
                                   const float* farp, const float* prefiltered, int n_pref, const float* lut, int lut_h,
                                   int lut_w, int lut_coef, int correct_depth, int n, float* out, float* out_srgb, int device,
                                   void* stream) {
  if (n == 0) return 0;
  if (n < 0 || n_pref < 1 || lut_h < 2 || lut_w < 2 || !rays_d || !normal || !maps || !nearp || !farp || !prefiltered || !lut ||
      !(lut_coef == 0 || lut_coef == 1) || !out) return IBLN_EINVAL;
  DeviceGuard g(device);
  shade_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      rays_d, normal, maps + IBLN_MAP_ALBEDO, maps + IBLN_MAP_ROUGH, maps + IBLN_MAP_IRR, maps + IBLN_MAP_ROUGH,
      maps + IBLN_MAP_DEPTH, nearp, farp, prefiltered, n_pref, lut, lut_h, lut_w, lut_coef, correct_depth, n, out, out_srgb,
      IBLN_MAPS_STRIDE, IBLN_MAPS_STRIDE);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_shade_bwd_maps(const float* rays_d, const float* normal, const float* maps, const float* nearp,
                                   const float* farp, const float* prefiltered, int n_pref, const float* lut, int lut_h,
                                   int lut_w, int lut_coef, int correct_depth, int n, const float* g_out,
                                   const float* g_out_srgb, float* g_maps, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || n_pref < 1 || lut_h < 2 || lut_w < 2 || !rays_d || !normal || !maps || !nearp || !farp || !prefiltered || !lut ||
      !(lut_coef == 0 || lut_coef == 1) || !g_maps) return IBLN_EINVAL;
  DeviceGuard g(device);
  shade_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      rays_d, normal, maps + IBLN_MAP_ALBEDO, maps + IBLN_MAP_ROUGH, maps + IBLN_MAP_IRR, maps + IBLN_MAP_ROUGH,
      maps + IBLN_MAP_DEPTH, nearp, farp, prefiltered, n_pref, lut, lut_h, lut_w, lut_coef, correct_depth, n, g_out, g_out_srgb,
      nullptr, nullptr, nullptr, nullptr, IBLN_MAPS_STRIDE, IBLN_MAPS_STRIDE, g_maps);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_shade_bwd(const float* rays_d, const float* normal, const float* albedo, const float* rough,
                              const float* irr, const float* mip_rough, const float* depth, const float* nearp,
                              const float* farp, const float* prefiltered, int n_pref, const float* lut, int lut_h,
                              int lut_w, int lut_coef, int correct_depth, int n, const float* g_out,
                              const float* g_out_srgb, float* g_albedo, float* g_rough, float* g_irr, float* g_mip_rough,
                              int device, void* stream) {
  if (n == 0) return 0;
  if (!SHADE_ARGS_OK || !g_albedo || !g_rough || !g_irr || !g_mip_rough) return IBLN_EINVAL;
  DeviceGuard g(device);
  shade_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rays_d, normal, albedo, rough, irr, mip_rough, depth,
                                                                     nearp, farp, prefiltered, n_pref, lut, lut_h, lut_w,
                                                                     lut_coef, correct_depth, n, g_out, g_out_srgb,
                                                                     g_albedo, g_rough, g_irr, g_mip_rough, 3, 1, nullptr);
  IBLN_RETURN_LAST();
}

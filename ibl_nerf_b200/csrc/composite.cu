// raw2outputs alpha compositing, forward and backward (north-star subsystem 4).
// One warp per ray.  The ray's raw[S,C] block is streamed through shared memory in rows of 32 samples with
// 16-byte cp.async (fully coalesced, double buffered per warp); the transmittance scan is a warp shuffle
// product scan with a carry across rows; per-sample state lives in registers / shared memory only.
#include "common.cuh"

namespace ibln {

constexpr int CP_WARPS = 4;   // warps per CTA (fewer when one ray tile is large)
constexpr int MAXCH = 32;

__device__ __forceinline__ float head_act(float x, int sigm) { return sigm ? sigmoidf_fast(x) : fmaxf(x, 0.f); }
__device__ __forceinline__ float head_dact(float x, float y, int sigm) { return sigm ? y * (1.f - y) : (x > 0.f ? 1.f : 0.f); }

// Issue the async copy of one ray's tile (n_float floats) into smem. Falls back to scalar loads
// when the tile is not 16-byte tileable.
__device__ __forceinline__ void tile_load_async(float* dst, const float* __restrict__ src, int n_float, int lane, bool vec_ok) {
  if (vec_ok) {
    int n4 = n_float >> 2;
    for (int i = lane; i < n4; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
  } else {
    for (int i = lane; i < n_float; i += 32) dst[i] = src[i];
  }
  cp_async_commit();
}

struct RayAlpha {   // per-sample quantities of one 32-sample row
  float alpha, om, T, w;
};

// alpha / transmittance / weight for sample i of the row, given the running carry (product of om
// over all previous rows).  Updates carry.
__device__ __forceinline__ RayAlpha row_alpha(float sig, float dist, bool valid, float& carry, int lane) {
  RayAlpha r;
  r.alpha = valid ? 1.0f - expf(-fmaxf(sig, 0.f) * dist) : 0.f;
  r.om = valid ? (1.0f - r.alpha) + 1e-10f : 1.0f;
  float p = r.om;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(FULL, p, o);
    if (lane >= o) p *= v;
  }
  float excl = __shfl_up_sync(FULL, p, 1);
  r.T = carry * (lane == 0 ? 1.0f : excl);
  r.w = r.alpha * r.T;
  carry *= __shfl_sync(FULL, p, 31);
  return r;
}

// Row streaming: a ray's raw[S,C] block is consumed in rows of 32 samples (32*C contiguous floats).  Each warp
// double-buffers rows with 16-byte cp.async and treats (ray, row) as one flat stream, so the next row -- of this
// ray or of the warp's next ray -- is always in flight while the current one is processed.  ~4.7 KB of shared
// memory per warp keeps 32+ warps resident per SM, which is what hides the MUFU / shuffle latency.
constexpr int ROW = 32;

__device__ __forceinline__ void row_load_async(float* dst, const float* __restrict__ raw, int64_t r, int k, int S, int C,
                                               int lane, bool vec_ok) {
  const int n_float = min(ROW, S - k * ROW) * C;
  const float* src = raw + ((int64_t)r * S + (int64_t)k * ROW) * C;
  if (vec_ok) {
    for (int i = lane; i < (n_float >> 2); i += 32) cp_async16(dst + 4 * i, src + 4 * i);
  } else {
    for (int i = lane; i < n_float; i += 32) dst[i] = src[i];
  }
  cp_async_commit();
}

// The 18 channels the kernels read of one staged sample.  FIXED layout (C = 18): 9 float2 loads -- with the 72-byte
// sample stride the 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs, so each access is conflict-free, while
// scalar accesses at stride 18 are 2-way bank-conflicted (ncu: 6.7 shared-memory wavefronts per sample in the backward).
template <bool FIXED>
__device__ __forceinline__ void load_sample(float (&x)[18], const float* px, int C) {
  if (FIXED) {
    const float2* p2 = reinterpret_cast<const float2*>(px);
#pragma unroll
    for (int q = 0; q < 9; ++q) { const float2 v = p2[q]; x[2 * q] = v.x; x[2 * q + 1] = v.y; }
  } else {
#pragma unroll
    for (int c = 0; c < 18; ++c) x[c] = c < C ? px[c] : 0.f;
  }
}

// FIXED: the reference's channel layout (C = 18, 3 coarse heads, sigmoid radiance) at compile time; FULL: S % 32 == 0
// (complete rows: no validity predicates, unrolled row copies) -- see composite_bwd_kernel.
template <bool SIMPLE, bool FIXED, bool FULL>
__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     const float* __restrict__ noise, int n, int S, int C_rt, int nc_rt, int sigm_rt,
                     float* __restrict__ weights, float* __restrict__ maps, float* __restrict__ maps_srgb,
                     float* __restrict__ pre_out) {
  const int C = FIXED ? 18 : C_rt, nc = FIXED ? 3 : nc_rt, sigm = FIXED ? 1 : sigm_rt;
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rowf = ROW * C;
  float* buf = sm + (size_t)warp * (2 * rowf + 32);
  float* s_out = buf + 2 * rowf;
  // rows must start 16-byte aligned: C*32*4 bytes per row always is; the ray base needs S*C*4 % 16 == 0
  const bool vec_ok = (((int64_t)S * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(raw) & 15) == 0);
  const int nwarps = blockDim.x >> 5;
  const int stride = gridDim.x * nwarps;
  const int nrows = (S + ROW - 1) / ROW;
  int r = blockIdx.x * nwarps + warp;
  if (r >= n) return;
  auto load_row = [&](float* dst, int64_t rr, int kk) {
    if (FIXED && FULL) {          // 144 float4 per row: 4.5 per lane, fully unrolled
      const float* src = raw + ((int64_t)rr * S + (int64_t)kk * ROW) * 18;
      if (vec_ok) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int e = lane + 32 * j;
          if (j < 4 || e < 144) cp_async16(dst + 4 * e, src + 4 * e);
        }
        cp_async_commit();
        return;
      }
    }
    row_load_async(dst, raw, rr, kk, S, C, lane, vec_ok);
  };
  load_row(buf, r, 0);
  int it = 0;
  for (; r < n; r += stride) {
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (int64_t)r * S;
    float carry = 1.0f;
    float a_depth = 0.f, a_acc = 0.f, a_rough = 0.f, a_irr = 0.f;
    float a_col[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) a_col[c] = 0.f;   // albedo 0..2, radiance 3..5, coarse 6..14

    for (int k = 0; k < nrows; ++k, ++it) {
      float* cur = buf + (it & 1) * rowf;
      float* nxt = buf + ((it + 1) & 1) * rowf;
      // prefetch the next row of the stream
      const bool more_rows = k + 1 < nrows;
      const int rn = more_rows ? r : r + stride;
      if (rn < n) { load_row(nxt, rn, more_rows ? k + 1 : 0); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncwarp();

      const int i = k * ROW + lane;
      const bool valid = FULL ? true : (i < S);
      const float zi = valid ? zr[i] : 0.f;
      float dist = (valid && i < S - 1) ? (zr[i + 1] - zi) : 1e10f;
      dist *= dnorm;
      // (scalar channel reads here: preloading the sample as float2 pairs, which pays off in the backward kernels, made
      // the forward 15-20 % slower -- it is bound by load latency, not by the shared-memory pipe)
      const float* px = cur + (size_t)(valid ? lane : 0) * C;
      float sig = px[0];
      if (noise != nullptr && valid) sig += noise[(int64_t)r * S + i];
      RayAlpha ra = row_alpha(sig, dist, valid, carry, lane);
      if (valid) {
        if (!SIMPLE) {
          weights[(int64_t)r * S + i] = ra.w;
          a_depth += ra.w * zi;
          a_acc += ra.w;
          a_rough += ra.w * sigmoidf_fast(px[4]);
          a_irr += ra.w * head_act(px[5], sigm);
#pragma unroll
          for (int c = 0; c < 3; ++c) a_col[c] += ra.w * sigmoidf_fast(px[1 + c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) a_col[3 + c] += ra.w * head_act(px[6 + c], sigm);
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
          if (kk < nc) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a_col[6 + 3 * kk + c] += ra.w * head_act(px[9 + 3 * kk + c], sigm);
          }
      }
      __syncwarp();      // everyone is done with `cur` before it becomes the prefetch target of the next iteration
    }
    // reductions
    if (!SIMPLE) {
      a_depth = warp_sum(a_depth); a_acc = warp_sum(a_acc); a_rough = warp_sum(a_rough); a_irr = warp_sum(a_irr);
    }
#pragma unroll
    for (int c = 0; c < 15; ++c) a_col[c] = warp_sum(a_col[c]);
    if (lane == 0) {
      if (!SIMPLE) {
        float q = a_depth / a_acc;
        float m = (q != q) ? q : fmaxf(1e-10f, q);    // torch.max propagates NaN (acc == 0)
        s_out[IBLN_MAP_DEPTH] = a_depth; s_out[IBLN_MAP_ACC] = a_acc; s_out[IBLN_MAP_DISP] = 1.0f / m;
        s_out[IBLN_MAP_TEND] = carry; s_out[IBLN_MAP_ROUGH] = a_rough; s_out[IBLN_MAP_IRR] = a_irr;
#pragma unroll
        for (int c = 0; c < 15; ++c) s_out[IBLN_MAP_ALBEDO + c] = (c < 6 + 3 * nc) ? a_col[c] : 0.f;
        for (int c = 21; c < 24; ++c) s_out[c] = 0.f;
      } else {
#pragma unroll
        for (int c = 0; c < 12; ++c) s_out[c] = a_col[3 + c];
      }
    }
    __syncwarp();
    if (!SIMPLE) {
      if (lane < IBLN_MAPS_STRIDE) {
        float v = s_out[lane];
        maps[(int64_t)r * IBLN_MAPS_STRIDE + lane] = v;
        if (maps_srgb != nullptr) {
          bool colour = (lane == IBLN_MAP_IRR) || (lane >= IBLN_MAP_ALBEDO && lane < IBLN_MAP_COARSE + 9);
          maps_srgb[(int64_t)r * IBLN_MAPS_STRIDE + lane] = colour ? srgbf(v) : v;
        }
      }
    } else {
      if (lane < 3 + 3 * nc) {
        pre_out[(int64_t)r * (3 + 3 * nc) + lane] = s_out[lane];
        if (maps_srgb != nullptr) maps_srgb[(int64_t)r * (3 + 3 * nc) + lane] = srgbf(s_out[lane]);   // :485-496 gamma
      }
    }
    __syncwarp();
  }
}

// Backward.  Pass 1 streams the ray's rows forward, recomputing alpha / T (kept in shared memory, 3 floats per
// sample) and the maps the non-linear outputs need; pass 2 streams the rows again in reverse (L2 hits) for the
// suffix sum  sum_{k>i} gw_k w_k, builds each row of g_raw in place in the row buffer and writes it out with
// coalesced 16-byte stores.
// FIXED = true: the reference's layout (C = 18 channels, 3 coarse-radiance heads, sigmoid radiance) as compile-time
// constants -- the kernel is instruction-issue bound and the run-time channel / activation switches cost ~15 %.
// FULL = true: S is a multiple of 32 (every shipped / benchmarked sample count), so every row is complete: the
// per-sample validity predicates and the ragged row copies fold away.
template <bool FIXED, bool FULL>
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     const float* __restrict__ noise, const float* __restrict__ g_weights,
                     const float* __restrict__ g_maps, const float* __restrict__ g_srgb, int n, int S, int C_rt, int nc_rt,
                     int sigm_rt, float* __restrict__ g_raw) {
  const int C = FIXED ? 18 : C_rt, nc = FIXED ? 3 : nc_rt, sigm = FIXED ? 1 : sigm_rt;
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rowf = ROW * C;
  const int nrows = (S + ROW - 1) / ROW;
  const int Sp = nrows * ROW;
  float* buf = sm + (size_t)warp * (2 * rowf + 4 * Sp + 32);
  float* s_alpha = buf + 2 * rowf;
  float* s_T = s_alpha + Sp;
  float* s_dist = s_T + Sp;
  float* s_z = s_dist + Sp;     // the ray's depths: pass 2 re-reads them from here, not from global memory
  float* s_g = s_z + Sp;        // 24 combined per-ray gradients
  const bool vec_ok = (((int64_t)S * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(raw) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(g_raw) & 15) == 0);
  const int nwarps = blockDim.x >> 5;
  const int stride = gridDim.x * nwarps;
  int r = blockIdx.x * nwarps + warp;
  if (r >= n) return;
  // flat task stream per ray: t in [0, 2*nrows): row(t) = t < nrows ? t : 2*nrows-1-t
  // L2 policy: a row read by pass 1 is read once more by pass 2 (keep it: evict_last), after which it is dead
  // (evict_first).  Without the hints the reverse pass of long rays misses L2 (ncu at S = 512: DRAM reads 1.64 x raw).
  const uint64_t pol_keep = l2_evict_last_policy(), pol_drop = l2_evict_first_policy();
  auto load_row = [&](float* dst, int64_t rr, int kk, bool second) {
    if (FIXED && FULL) {          // 32 samples x 18 channels = 144 float4 per row: 4.5 per lane, fully unrolled
      const float* src = raw + ((int64_t)rr * S + (int64_t)kk * ROW) * 18;
      if (vec_ok) {
        const uint64_t pol = second ? pol_drop : pol_keep;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int e = lane + 32 * j;
          if (j < 4 || e < 144) cp_async16_hint(dst + 4 * e, src + 4 * e, pol);
        }
        cp_async_commit();
        return;
      }
    }
    row_load_async(dst, raw, rr, kk, S, C, lane, vec_ok);
  };
  load_row(buf, r, 0, false);
  int it = 0;
  for (; r < n; r += stride) {
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (int64_t)r * S;
    float carry = 1.0f;
    float a_depth = 0.f, a_acc = 0.f, a_irr = 0.f;
    float a_col[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) a_col[c] = 0.f;
    float gdepth = 0.f, gacc = 0.f, suffix = 0.f;

    for (int t = 0; t < 2 * nrows; ++t, ++it) {
      float* cur = buf + (it & 1) * rowf;
      float* nxt = buf + ((it + 1) & 1) * rowf;
      const bool more = t + 1 < 2 * nrows;
      const int rn = more ? r : r + stride;
      const int tn = more ? t + 1 : 0;
      if (rn < n) { load_row(nxt, rn, tn < nrows ? tn : 2 * nrows - 1 - tn, tn >= nrows); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncwarp();
      const int k = t < nrows ? t : 2 * nrows - 1 - t;
      const int i = k * ROW + lane;
      const bool valid = FULL ? true : (i < S);
      const float zi = t < nrows ? (valid ? zr[i] : 0.f) : s_z[i];
      float* pxp = cur + (size_t)(valid ? lane : 0) * C;
      float px[18];
      load_sample<FIXED>(px, pxp, C);

      if (t < nrows) {
        // ---- pass 1
        s_z[i] = zi;
        float dist = (valid && i < S - 1) ? (zr[i + 1] - zi) : 1e10f;
        dist *= dnorm;
        float sig = px[0];
        if (noise != nullptr && valid) sig += noise[(int64_t)r * S + i];
        RayAlpha ra = row_alpha(sig, dist, valid, carry, lane);
        s_alpha[i] = ra.alpha; s_T[i] = ra.T; s_dist[i] = (sig > 0.f) ? dist : 0.f;
        if (valid) {
          a_depth += ra.w * zi; a_acc += ra.w;
          if (g_srgb != nullptr) {
            a_irr += ra.w * head_act(px[5], sigm);
#pragma unroll
            for (int c = 0; c < 3; ++c) a_col[c] += ra.w * sigmoidf_fast(px[1 + c]);
#pragma unroll
            for (int c = 0; c < 3; ++c) a_col[3 + c] += ra.w * head_act(px[6 + c], sigm);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
              if (kk < nc) {
#pragma unroll
                for (int c = 0; c < 3; ++c) a_col[6 + 3 * kk + c] += ra.w * head_act(px[9 + 3 * kk + c], sigm);
              }
          }
        }
        if (t == nrows - 1) {
          // ---- end of pass 1: combined per-ray gradients wrt the LINEAR maps
          a_depth = warp_sum(a_depth); a_acc = warp_sum(a_acc);
          if (g_srgb != nullptr) {
            a_irr = warp_sum(a_irr);
#pragma unroll
            for (int c = 0; c < 15; ++c) a_col[c] = warp_sum(a_col[c]);
          }
          if (lane < IBLN_MAPS_STRIDE) {
            float g = g_maps ? g_maps[(int64_t)r * IBLN_MAPS_STRIDE + lane] : 0.f;
            if (g_srgb != nullptr) {
              float gs = g_srgb[(int64_t)r * IBLN_MAPS_STRIDE + lane];
              bool colour = (lane == IBLN_MAP_IRR) || (lane >= IBLN_MAP_ALBEDO && lane < IBLN_MAP_COARSE + 9);
              if (colour) {
                float lin = (lane == IBLN_MAP_IRR) ? a_irr : 0.f;
#pragma unroll
                for (int c = 0; c < 15; ++c) if (lane == IBLN_MAP_ALBEDO + c) lin = a_col[c];
                g += gs * dsrgbf(lin);
              } else {
                g += gs;
              }
            }
            s_g[lane] = g;
          }
          __syncwarp();
          gdepth = s_g[IBLN_MAP_DEPTH]; gacc = s_g[IBLN_MAP_ACC];
          const float gdisp = s_g[IBLN_MAP_DISP];
          if (gdisp != 0.f) {
            float q = a_depth / a_acc;
            if (q > 1e-10f) {     // disp = 1/q : d/d depth = -1/(q^2 acc), d/d acc = depth/(q^2 acc^2)
              float iq2 = 1.0f / (q * q);
              gdepth += gdisp * (-iq2 / a_acc);
              gacc += gdisp * (iq2 * a_depth / (a_acc * a_acc));
            }
          }
          // d T_end / d alpha_i = -T_end / om_i : enters the suffix term as an extra "sample" past the end
          suffix = s_g[IBLN_MAP_TEND] * carry;
        }
      } else {
        // ---- pass 2 (reverse)
        const float alpha = s_alpha[i], T = s_T[i], w = alpha * T;
        float gw = 0.f;
        float go[18];
#pragma unroll
        for (int c = 0; c < 18; ++c) go[c] = 0.f;
        if (valid) {
          gw = (g_weights ? g_weights[(int64_t)r * S + i] : 0.f) + gdepth * zi + gacc;
#pragma unroll
          for (int c = 0; c < 3; ++c) {   // radiance (live weights)
            float x = px[6 + c], y = head_act(x, sigm), gm = s_g[IBLN_MAP_RAD + c];
            gw += gm * y;
            go[6 + c] = w * gm * head_dact(x, y, sigm);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) { float x = px[1 + c], y = sigmoidf_fast(x); go[1 + c] = w * s_g[IBLN_MAP_ALBEDO + c] * y * (1.f - y); }
          { float x = px[4], y = sigmoidf_fast(x); go[4] = w * s_g[IBLN_MAP_ROUGH] * y * (1.f - y); }
          { float x = px[5], y = head_act(x, sigm); go[5] = w * s_g[IBLN_MAP_IRR] * head_dact(x, y, sigm); }
#pragma unroll
          for (int kk = 0; kk < 3; ++kk)
            if (kk < nc) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float x = px[9 + 3 * kk + c], y = head_act(x, sigm);
                go[9 + 3 * kk + c] = w * s_g[IBLN_MAP_COARSE + 3 * kk + c] * head_dact(x, y, sigm);
              }
            }
        }
        // inclusive suffix scan of gw*w over the row (towards higher lanes)
        const float tv = valid ? gw * w : 0.f;
        float p = tv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float v = __shfl_down_sync(FULL, p, o);
          if (lane + o < 32) p += v;
        }
        const float excl = p - tv + suffix;          // sum over k > i
        suffix += __shfl_sync(FULL, p, 0);
        if (valid) {
          const float om = (1.0f - alpha) + 1e-10f;
          const float galpha = gw * T - excl / om;
          go[0] = galpha * s_dist[i] * (1.0f - alpha);   // d alpha/d sigma = dist * exp(-sigma dist), 0 where sigma <= 0
          if (FIXED) {
            float2* p2 = reinterpret_cast<float2*>(pxp);
#pragma unroll
            for (int q = 0; q < 9; ++q) p2[q] = make_float2(go[2 * q], go[2 * q + 1]);
          } else {
#pragma unroll
            for (int c = 0; c < 18; ++c) if (c < C) pxp[c] = go[c];
            for (int c = 18; c < C; ++c) pxp[c] = 0.f;
          }
        }
        __syncwarp();
        const int n_float = FULL ? ROW * C : min(ROW, S - k * ROW) * C;
        float* dst = g_raw + ((int64_t)r * S + (int64_t)k * ROW) * C;
        if (FIXED && FULL && vec_ok) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const int e = lane + 32 * j;
            if (j < 4 || e < 144) __stcs(reinterpret_cast<float4*>(dst) + e, reinterpret_cast<float4*>(cur)[e]);   // streaming: keep the rows
          }                                                                                                     // pass 2 re-reads in L2
        } else if (vec_ok) {
          for (int e = lane; e < (n_float >> 2); e += 32) __stcs(reinterpret_cast<float4*>(dst) + e, reinterpret_cast<float4*>(cur)[e]);
        } else {
          for (int e = lane; e < n_float; e += 32) dst[e] = cur[e];
        }
      }
      __syncwarp();
    }
  }
}

// Backward, RAY-RESIDENT variant for the reference layout (C = 18, 3 coarse heads, sigmoid heads) and S % 32 == 0,
// S <= RESIDENT_MAX_S (the coarse pass's 64 samples; at 192 samples the 16 KB tile per warp leaves 12 warps per SM and
// the streaming kernel is faster: 0.76 vs 0.54 of the HBM copy peak, measured).
// The streaming kernel above re-reads every row from L2 /
// DRAM in its reverse pass and evaluates every head sigmoid twice; ncu's instruction mix puts it at the MUFU pipe
// (68 ex2 / rcp per sample: 2 x 17 sigmoids + the alpha exponential) rather than at HBM.  Here the ray's whole
// [S,18] tile stays in shared memory: pass 1 overwrites each head channel with its ACTIVATION y = sigmoid(x), pass 2
// (reverse) needs only y and y (1 - y) -- no second global read, no second MUFU pass -- and builds g_raw in place.
// While pass 2 walks the rows backwards, every row it has streamed out is immediately refilled with the same row of
// the warp's NEXT ray (cp.async), so the loads of ray r+1 overlap the arithmetic of ray r.
constexpr int RESIDENT_MAX_S = 128;
__global__ void __launch_bounds__(192, 6)
composite_bwd_resident_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                              const float* __restrict__ noise, const float* __restrict__ g_weights,
                              const float* __restrict__ g_maps, const float* __restrict__ g_srgb, int n, int S,
                              float* __restrict__ g_raw) {
  constexpr int C = 18, rowf = ROW * C;           // 576 floats = 144 float4 per row
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = S / ROW;
  float* tile = sm + (size_t)warp * ((size_t)S * C + 4 * S + 32);
  float* s_alpha = tile + (size_t)S * C;
  float* s_T = s_alpha + S;
  float* s_dist = s_T + S;
  float* s_z = s_dist + S;
  float* s_g = s_z + S;
  const int nwarps = blockDim.x >> 5;
  const int stride = gridDim.x * nwarps;
  int r = blockIdx.x * nwarps + warp;
  if (r >= n) return;
  auto load_row = [&](int64_t rr, int k) {       // row k of ray rr -> tile row k (one commit group)
    const float* src = raw + ((int64_t)rr * S + (int64_t)k * ROW) * C;
    float* dst = tile + (size_t)k * rowf;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int e = lane + 32 * j;
      if (j < 4 || e < 144) cp_async16(dst + 4 * e, src + 4 * e);
    }
    cp_async_commit();
  };
  for (int k = 0; k < nrows; ++k) load_row(r, k);
  for (; r < n; r += stride) {
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (int64_t)r * S;
    cp_async_wait<0>();                            // (rows were issued during the previous ray's reverse pass)
    __syncwarp();
    float carry = 1.0f, a_depth = 0.f, a_acc = 0.f, a_irr = 0.f;
    float a_col[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) a_col[c] = 0.f;
    // ---- pass 1: alpha / T per sample, head activations written back in place, per-ray sums
    float z_cur = zr[lane];
    for (int k = 0; k < nrows; ++k) {
      const int i = k * ROW + lane;
      const float zi = z_cur;
      if (k + 1 < nrows) z_cur = zr[i + ROW];      // next row's depth: one global load per row, issued a row ahead
      float z_up = __shfl_down_sync(FULL, zi, 1);  // z[i + 1]: the neighbour lane's, or the next row's first
      const float z_first_next = __shfl_sync(FULL, z_cur, 0);
      if (lane == 31) z_up = z_first_next;
      s_z[i] = zi;
      float dist = (i < S - 1) ? (z_up - zi) : 1e10f;
      dist *= dnorm;
      // a sample's 18 channels as 9 float2: with the 72-byte sample stride the 16 lanes of a half-warp hit 16 distinct
      // 8-byte bank pairs, so each access is conflict-free (scalar accesses at stride 18 are 2-way conflicted)
      float2* px2 = reinterpret_cast<float2*>(tile + (size_t)i * C);
      float x[18];
#pragma unroll
      for (int q = 0; q < 9; ++q) { const float2 v = px2[q]; x[2 * q] = v.x; x[2 * q + 1] = v.y; }
      float sig = x[0];
      if (noise != nullptr) sig += noise[(int64_t)r * S + i];
      RayAlpha ra = row_alpha(sig, dist, true, carry, lane);
      s_alpha[i] = ra.alpha; s_T[i] = ra.T; s_dist[i] = (sig > 0.f) ? dist : 0.f;
      a_depth += ra.w * zi; a_acc += ra.w;
      float y[17];
#pragma unroll
      for (int c = 0; c < 17; ++c) y[c] = sigmoidf_fast(x[1 + c]);
      px2[0] = make_float2(x[0], y[0]);
#pragma unroll
      for (int q = 1; q < 9; ++q) px2[q] = make_float2(y[2 * q - 1], y[2 * q]);
      a_irr += ra.w * y[4];
#pragma unroll
      for (int c = 0; c < 3; ++c) a_col[c] += ra.w * y[c];
#pragma unroll
      for (int c = 0; c < 12; ++c) a_col[3 + c] += ra.w * y[5 + c];
    }
    // ---- combined per-ray gradients wrt the LINEAR maps (same expressions as the streaming kernel)
    a_depth = warp_sum(a_depth); a_acc = warp_sum(a_acc);
    if (g_srgb != nullptr) {
      a_irr = warp_sum(a_irr);
#pragma unroll
      for (int c = 0; c < 15; ++c) a_col[c] = warp_sum(a_col[c]);
    }
    if (lane < IBLN_MAPS_STRIDE) {
      float g = g_maps ? g_maps[(int64_t)r * IBLN_MAPS_STRIDE + lane] : 0.f;
      if (g_srgb != nullptr) {
        const float gs = g_srgb[(int64_t)r * IBLN_MAPS_STRIDE + lane];
        const bool colour = (lane == IBLN_MAP_IRR) || (lane >= IBLN_MAP_ALBEDO && lane < IBLN_MAP_COARSE + 9);
        if (colour) {
          float lin = (lane == IBLN_MAP_IRR) ? a_irr : 0.f;
#pragma unroll
          for (int c = 0; c < 15; ++c) if (lane == IBLN_MAP_ALBEDO + c) lin = a_col[c];
          g += gs * dsrgbf(lin);
        } else {
          g += gs;
        }
      }
      s_g[lane] = g;
    }
    __syncwarp();
    float gdepth = s_g[IBLN_MAP_DEPTH], gacc = s_g[IBLN_MAP_ACC];
    const float gdisp = s_g[IBLN_MAP_DISP];
    if (gdisp != 0.f) {
      const float q = a_depth / a_acc;
      if (q > 1e-10f) {
        const float iq2 = 1.0f / (q * q);
        gdepth += gdisp * (-iq2 / a_acc);
        gacc += gdisp * (iq2 * a_depth / (a_acc * a_acc));
      }
    }
    float suffix = s_g[IBLN_MAP_TEND] * carry;
    float gm[17];                                  // d loss / d (linear map) of the 17 head channels, raw-channel order
#pragma unroll
    for (int c = 0; c < 3; ++c) gm[c] = s_g[IBLN_MAP_ALBEDO + c];
    gm[3] = s_g[IBLN_MAP_ROUGH]; gm[4] = s_g[IBLN_MAP_IRR];
#pragma unroll
    for (int c = 0; c < 12; ++c) gm[5 + c] = s_g[IBLN_MAP_RAD + c];
    const int rn = r + stride;
    // ---- pass 2 (reverse): g_raw rows built in place, streamed out, row refilled with the next ray
    for (int k = nrows - 1; k >= 0; --k) {
      const int i = k * ROW + lane;
      const float zi = s_z[i];
      float2* px2 = reinterpret_cast<float2*>(tile + (size_t)i * C);
      const float alpha = s_alpha[i], T = s_T[i], w = alpha * T;
      float gw = (g_weights ? g_weights[(int64_t)r * S + i] : 0.f) + gdepth * zi + gacc;
      float yy[18], go[18];
#pragma unroll
      for (int q = 0; q < 9; ++q) { const float2 v = px2[q]; yy[2 * q] = v.x; yy[2 * q + 1] = v.y; }
#pragma unroll
      for (int c = 0; c < 17; ++c) {
        const float y = yy[1 + c];
        if (c >= 5 && c < 8) gw += gm[c] * y;     // radiance composites with LIVE weights (:305-306)
        go[1 + c] = w * gm[c] * y * (1.f - y);
      }
      const float tv = gw * w;
      float p = tv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_down_sync(FULL, p, o);
        if (lane + o < 32) p += v;
      }
      const float excl = p - tv + suffix;
      suffix += __shfl_sync(FULL, p, 0);
      const float om = (1.0f - alpha) + 1e-10f;
      const float galpha = gw * T - excl / om;
      go[0] = galpha * s_dist[i] * (1.0f - alpha);
#pragma unroll
      for (int q = 0; q < 9; ++q) px2[q] = make_float2(go[2 * q], go[2 * q + 1]);
      __syncwarp();
      float4* dst = reinterpret_cast<float4*>(g_raw + ((int64_t)r * S + (int64_t)k * ROW) * C);
      const float4* srow = reinterpret_cast<const float4*>(tile + (size_t)k * rowf);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int e = lane + 32 * j;
        if (j < 4 || e < 144) __stcs(dst + e, srow[e]);
      }
      __syncwarp();                                // the row has been read out: refill it
      if (rn < n) load_row(rn, k);
    }
  }
}

// Backward, CTA-PER-RAY variant for long rays (reference layout, S % 32 == 0, S >= 256: the sweep's 256 / 512).  A warp cannot keep such a ray resident (36 KB at S = 512 would leave 4 warps per SM),
// and the streaming kernel re-reads every row in its reverse pass -- from DRAM once the rays in flight exceed L2 (ncu at
// S = 512: DRAM reads 1.64 x raw, 0.61 of the HBM copy peak).  Here ONE CTA of W warps owns the ray: its [S,18] tile is
// resident in shared memory, warp w owns rows w, w + W, ... in both passes (so the tile rows stay warp-private and a
// streamed-out row is refilled with the same row of the CTA's next ray right away), and the two scans are split into a
// row-local shuffle scan + a carry over the row totals exchanged through shared memory:
//   T_i      = (prod_{rows before} rowprod) * (exclusive product inside the row)        -- same order as the warp kernels
//   sum_{k>i} gw_k w_k = (exclusive suffix inside the row) + sum_{rows after} rowsum + g_Tend T_end
// Like the ray-resident warp kernel it evaluates every head sigmoid once and never reads raw twice.
template <int W>
__global__ void __launch_bounds__(W * 32, 4)       // 32 warps per SM at W = 8 (64 registers)
composite_bwd_cta_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                         const float* __restrict__ noise, const float* __restrict__ g_weights,
                         const float* __restrict__ g_maps, const float* __restrict__ g_srgb, int n, int S,
                         float* __restrict__ g_raw) {
  constexpr int C = 18, rowf = ROW * C;
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = S / ROW;
  float* tile = sm;                               // [S][18]
  float* s_alpha = tile + (size_t)S * C;          // [S] each
  float* s_T = s_alpha + S;                       // pass 1a: exclusive product inside the row; pass 1b on: T
  float* s_dist = s_T + S;
  float* s_z = s_dist + S;
  float* s_suf = s_z + S;                         // exclusive suffix of gw * w inside the row
  float* s_gw = s_suf + S;
  float* s_rowprod = s_gw + S;                    // [nrows]
  float* s_rowsum = s_rowprod + nrows;            // [nrows]
  float* s_red = s_rowsum + nrows;                // [W][20] per-warp partial sums
  float* s_g = s_red + W * 20;                    // [32] combined per-ray gradients
  auto load_row = [&](int64_t rr, int k) {
    const float* src = raw + ((int64_t)rr * S + (int64_t)k * ROW) * C;
    float* dst = tile + (size_t)k * rowf;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int e = lane + 32 * j;
      if (j < 4 || e < 144) cp_async16(dst + 4 * e, src + 4 * e);
    }
  };
  int r = blockIdx.x;
  if (r >= n) return;
  for (int k = warp; k < nrows; k += W) load_row(r, k);
  cp_async_commit();
  for (; r < n; r += gridDim.x) {
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (int64_t)r * S;
    cp_async_wait<0>();
    __syncwarp();                                  // my rows have landed (rows are warp-private)
    // ---- pass 1a: alpha, row-local exclusive transmittance product, activations in place
    for (int k = warp; k < nrows; k += W) {
      const int i = k * ROW + lane;
      const float zi = zr[i];
      float dist = (i < S - 1) ? (zr[i + 1] - zi) : 1e10f;
      dist *= dnorm;
      float2* px2 = reinterpret_cast<float2*>(tile + (size_t)i * C);
      float x[18];
#pragma unroll
      for (int q = 0; q < 9; ++q) { const float2 v = px2[q]; x[2 * q] = v.x; x[2 * q + 1] = v.y; }
      float sig = x[0];
      if (noise != nullptr) sig += noise[(int64_t)r * S + i];
      const float alpha = 1.0f - expf(-fmaxf(sig, 0.f) * dist);
      const float om = (1.0f - alpha) + 1e-10f;
      float pr = om;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(FULL, pr, o);
        if (lane >= o) pr *= v;
      }
      const float excl = __shfl_up_sync(FULL, pr, 1);
      s_alpha[i] = alpha; s_T[i] = lane == 0 ? 1.0f : excl; s_dist[i] = (sig > 0.f) ? dist : 0.f; s_z[i] = zi;
      if (lane == 31) s_rowprod[k] = pr;
      float y[17];
#pragma unroll
      for (int c = 0; c < 17; ++c) y[c] = sigmoidf_fast(x[1 + c]);
      px2[0] = make_float2(x[0], y[0]);
#pragma unroll
      for (int q = 1; q < 9; ++q) px2[q] = make_float2(y[2 * q - 1], y[2 * q]);
    }
    __syncthreads();
    // ---- pass 1b: carries across rows -> T, weights, per-ray sums
    float a_depth = 0.f, a_acc = 0.f, a_irr = 0.f;
    float a_col[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) a_col[c] = 0.f;
    {
      float carry = 1.0f;
      int kk = 0;
      for (int k = warp; k < nrows; k += W) {
        for (; kk < k; ++kk) carry *= s_rowprod[kk];         // same multiplication order as the sequential warp kernels
        const int i = k * ROW + lane;
        const float T = carry * s_T[i];
        s_T[i] = T;
        const float w = s_alpha[i] * T;
        a_depth += w * s_z[i]; a_acc += w;
        if (g_srgb != nullptr) {
          const float2* px2 = reinterpret_cast<const float2*>(tile + (size_t)i * C);
          float yy[18];
#pragma unroll
          for (int q = 0; q < 9; ++q) { const float2 v = px2[q]; yy[2 * q] = v.x; yy[2 * q + 1] = v.y; }
          a_irr += w * yy[5];
#pragma unroll
          for (int c = 0; c < 3; ++c) a_col[c] += w * yy[1 + c];
#pragma unroll
          for (int c = 0; c < 12; ++c) a_col[3 + c] += w * yy[6 + c];
        }
      }
    }
    a_depth = warp_sum(a_depth); a_acc = warp_sum(a_acc);
    if (g_srgb != nullptr) {
      a_irr = warp_sum(a_irr);
#pragma unroll
      for (int c = 0; c < 15; ++c) a_col[c] = warp_sum(a_col[c]);
    }
    if (lane == 0) {
      float* rd = s_red + warp * 20;
      rd[0] = a_depth; rd[1] = a_acc; rd[2] = a_irr;
#pragma unroll
      for (int c = 0; c < 15; ++c) rd[3 + c] = a_col[c];
    }
    __syncthreads();
    if (warp == 0) {
      // totals over the warps (lane j < 18 owns sum j), then the combined gradients wrt the LINEAR maps
      float tot = 0.f;
      if (lane < 18)
        for (int ww = 0; ww < W; ++ww) tot += s_red[ww * 20 + lane];
      const float t_depth = __shfl_sync(FULL, tot, 0), t_acc = __shfl_sync(FULL, tot, 1), t_irr = __shfl_sync(FULL, tot, 2);
      float lin = 0.f;                                        // linear map value of maps column `lane` (colour columns only)
      if (lane == IBLN_MAP_IRR) lin = t_irr;
#pragma unroll
      for (int c = 0; c < 15; ++c) { const float v = __shfl_sync(FULL, tot, 3 + c); if (lane == IBLN_MAP_ALBEDO + c) lin = v; }
      float g = 0.f;
      if (lane < IBLN_MAPS_STRIDE) {
        g = g_maps ? g_maps[(int64_t)r * IBLN_MAPS_STRIDE + lane] : 0.f;
        if (g_srgb != nullptr) {
          const float gs = g_srgb[(int64_t)r * IBLN_MAPS_STRIDE + lane];
          const bool colour = (lane == IBLN_MAP_IRR) || (lane >= IBLN_MAP_ALBEDO && lane < IBLN_MAP_COARSE + 9);
          g += colour ? gs * dsrgbf(lin) : gs;
        }
      }
      float gdepth = __shfl_sync(FULL, g, IBLN_MAP_DEPTH), gacc = __shfl_sync(FULL, g, IBLN_MAP_ACC);
      const float gdisp = __shfl_sync(FULL, g, IBLN_MAP_DISP);
      if (gdisp != 0.f) {
        const float q = t_depth / t_acc;
        if (q > 1e-10f) {
          const float iq2 = 1.0f / (q * q);
          gdepth += gdisp * (-iq2 / t_acc);
          gacc += gdisp * (iq2 * t_depth / (t_acc * t_acc));
        }
      }
      if (lane < IBLN_MAPS_STRIDE) s_g[lane] = g;
      if (lane == 0) { s_g[24] = gdepth; s_g[25] = gacc; }
    }
    __syncthreads();
    const float gdepth = s_g[24], gacc = s_g[25];
    float gm[17];
#pragma unroll
    for (int c = 0; c < 3; ++c) gm[c] = s_g[IBLN_MAP_ALBEDO + c];
    gm[3] = s_g[IBLN_MAP_ROUGH]; gm[4] = s_g[IBLN_MAP_IRR];
#pragma unroll
    for (int c = 0; c < 12; ++c) gm[5 + c] = s_g[IBLN_MAP_RAD + c];
    // ---- pass 2a: gw, row-local exclusive suffix of gw * w, row totals
    for (int k = warp; k < nrows; k += W) {
      const int i = k * ROW + lane;
      const float2* px2 = reinterpret_cast<const float2*>(tile + (size_t)i * C);
      const float2 r3 = px2[3], r4 = px2[4];       // channels 6..9: radiance activations = channels 6, 7, 8
      float gw = (g_weights ? g_weights[(int64_t)r * S + i] : 0.f) + gdepth * s_z[i] + gacc;
      gw += gm[5] * r3.x + gm[6] * r3.y + gm[7] * r4.x;
      const float tv = gw * (s_alpha[i] * s_T[i]);
      float ps = tv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_down_sync(FULL, ps, o);
        if (lane + o < 32) ps += v;
      }
      s_gw[i] = gw; s_suf[i] = ps - tv;
      if (lane == 0) s_rowsum[k] = ps;
    }
    __syncthreads();
    // T_end = product of all row products (sequential order)
    float t_end = 1.0f;
    for (int kk = 0; kk < nrows; ++kk) t_end *= s_rowprod[kk];
    const float tend_term = s_g[IBLN_MAP_TEND] * t_end;
    const int rn = r + gridDim.x;
    // ---- pass 2b: g_raw rows in place, streamed out, refilled with the next ray's rows
    for (int k = warp; k < nrows; k += W) {
      float carry = tend_term;
      for (int kk = nrows - 1; kk > k; --kk) carry += s_rowsum[kk];
      const int i = k * ROW + lane;
      float2* px2 = reinterpret_cast<float2*>(tile + (size_t)i * C);
      const float alpha = s_alpha[i], T = s_T[i], w = alpha * T, gw = s_gw[i];
      float yy[18], go[18];
#pragma unroll
      for (int q = 0; q < 9; ++q) { const float2 v = px2[q]; yy[2 * q] = v.x; yy[2 * q + 1] = v.y; }
#pragma unroll
      for (int c = 0; c < 17; ++c) { const float y = yy[1 + c]; go[1 + c] = w * gm[c] * y * (1.f - y); }
      const float excl = s_suf[i] + carry;
      const float om = (1.0f - alpha) + 1e-10f;
      const float galpha = gw * T - excl / om;
      go[0] = galpha * s_dist[i] * (1.0f - alpha);
#pragma unroll
      for (int q = 0; q < 9; ++q) px2[q] = make_float2(go[2 * q], go[2 * q + 1]);
      __syncwarp();
      float4* dst = reinterpret_cast<float4*>(g_raw + ((int64_t)r * S + (int64_t)k * ROW) * C);
      const float4* srow = reinterpret_cast<const float4*>(tile + (size_t)k * rowf);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int e = lane + 32 * j;
        if (j < 4 || e < 144) __stcs(dst + e, srow[e]);
      }
      __syncwarp();
      if (rn < n) load_row(rn, k);
    }
    cp_async_commit();
    __syncthreads();                               // s_rowprod / s_rowsum / s_g / s_red are rewritten by the next ray
  }
}

// sigma-only depth compositing (normal estimator / raw2outputs_depth): no tile staging needed.
__global__ void __launch_bounds__(CP_WARPS * 32)
depth_fwd_kernel(const float* __restrict__ sigma, const float* __restrict__ z, const float* __restrict__ rays_d,
                 int reps, int n, int S, float* __restrict__ depth, float* __restrict__ weights,
                 float* __restrict__ visibility) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t rows = (int64_t)reps * n;
  for (int64_t m = (int64_t)blockIdx.x * CP_WARPS + warp; m < rows; m += (int64_t)gridDim.x * CP_WARPS) {
    int r = (int)(m % n);
    float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* zr = z + (int64_t)r * S;
    const float* sr = sigma + m * S;
    float carry = 1.0f, a_depth = 0.f;
    for (int base = 0; base < S; base += 32) {
      int i = base + lane;
      bool valid = i < S;
      float zi = valid ? zr[i] : 0.f;
      float dist = (valid && i < S - 1) ? (zr[i + 1] - zi) : 1e10f;
      dist *= dnorm;
      RayAlpha ra = row_alpha(valid ? sr[i] : 0.f, dist, valid, carry, lane);
      if (valid) {
        a_depth += ra.w * zi;
        if (weights != nullptr) weights[m * S + i] = ra.w;
      }
    }
    a_depth = warp_sum(a_depth);
    if (lane == 0) {
      depth[m] = a_depth;
      if (visibility != nullptr) visibility[m] = carry;
    }
  }
}

static int comp_grid(int64_t n, int device, int ctas_per_sm, int warps = CP_WARPS) {
  int64_t need = (n + warps - 1) / warps;
  int64_t cap = (int64_t)num_sms(device) * ctas_per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace ibln

using namespace ibln;

template <bool SIMPLE>
static int launch_fwd(const float* raw, const float* z, const float* d, const float* noise, int n, int S, int C, int nc,
                      int sigm, float* weights, float* maps, float* maps_srgb, float* pre, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || S < 1 || C < 9 + 3 * nc || C > MAXCH || nc < 0 || nc > 3 || !raw || !z || !d) return IBLN_EINVAL;
  DeviceGuard g(device);
  const int warps = 8;
  size_t smem = (size_t)warps * (2 * (size_t)ROW * C + 32) * sizeof(float);
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<comp_grid(n, device, per_sm, warps), warps * 32, smem, (cudaStream_t)stream>>>(raw, z, d, noise, n, S, C, nc, sigm,
                                                                                     weights, maps, maps_srgb, pre);
    return 0;
  };
  const bool fixed = C == 18 && nc == 3 && sigm == 1, full = S % 32 == 0;
  int rc = fixed ? (full ? launch(composite_fwd_kernel<SIMPLE, true, true>) : launch(composite_fwd_kernel<SIMPLE, true, false>))
                 : (full ? launch(composite_fwd_kernel<SIMPLE, false, true>) : launch(composite_fwd_kernel<SIMPLE, false, false>));
  if (rc != 0) return rc;
  IBLN_RETURN_LAST();
}

extern "C" int ibln_composite_fwd(const float* raw, const float* z, const float* rays_d, const float* noise, int n,
                                  int S, int C, int nc, int sigm, float* weights, float* maps, float* maps_srgb,
                                  int device, void* stream) {
  if (n == 0) return 0;
  if (!weights || !maps) return IBLN_EINVAL;
  return launch_fwd<false>(raw, z, rays_d, noise, n, S, C, nc, sigm, weights, maps, maps_srgb, nullptr, device, stream);
}

extern "C" int ibln_composite_simple_fwd(const float* raw, const float* z, const float* dirs, int n, int S, int C, int nc,
                                         int sigm, float* pre_out, float* pre_srgb, int device, void* stream) {
  if (n == 0) return 0;
  if (!pre_out) return IBLN_EINVAL;
  return launch_fwd<true>(raw, z, dirs, nullptr, n, S, C, nc, sigm, nullptr, nullptr, pre_srgb, pre_out, device, stream);
}

extern "C" int ibln_composite_bwd(const float* raw, const float* z, const float* rays_d, const float* noise,
                                  const float* g_weights, const float* g_maps, const float* g_maps_srgb, int n, int S,
                                  int C, int nc, int sigm, float* g_raw, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || S < 1 || C < 9 + 3 * nc || C > MAXCH || nc < 0 || nc > 3 || !raw || !z || !rays_d || !g_raw) return IBLN_EINVAL;
  DeviceGuard g(device);
  if (C == 18 && nc == 3 && sigm == 1 && S % 32 == 0 && S <= RESIDENT_MAX_S && (reinterpret_cast<uintptr_t>(raw) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(g_raw) & 15) == 0) {
    // ray-resident kernel: [S,18] tile + 4 S floats per warp (5.6 KB at S = 64)
    const size_t pw = ((size_t)S * 18 + 4 * (size_t)S + 32) * sizeof(float);
    int warps = 6;                                   // 6 CTAs of 6 warps per SM at S = 64 (36 warps, 56 registers)
    while (warps > 2 && warps * pw > 100 * 1024) warps >>= 1;
    const size_t smem = warps * pw;
    int per_sm = (int)((224 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    IBLN_CUDA(cudaFuncSetAttribute(composite_bwd_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    composite_bwd_resident_kernel<<<comp_grid(n, device, per_sm, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        raw, z, rays_d, noise, g_weights, g_maps, g_maps_srgb, n, S, g_raw);
    IBLN_RETURN_LAST();
  }
  if (C == 18 && nc == 3 && sigm == 1 && S % 32 == 0 && S >= 256 && S <= 1024 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(g_raw) & 15) == 0) {
    // CTA-per-ray kernel: [S,18] tile + 6 S floats + row tables per CTA (50 KB at S = 512).  Measured against the streaming
    // kernel: S = 512 0.67 vs 0.61 of the HBM copy peak; at S = 192 (one row per warp: five CTA barriers per 6 rows) 0.46 vs
    // 0.78, so shorter rays stay on the streaming kernel.
    const int nrows = S / 32;
    auto go = [&](auto kern, int w) -> int {
      const size_t smem = ((size_t)S * 18 + 6 * (size_t)S + 2 * nrows + w * 20 + 32) * sizeof(float);
      if (smem > 200 * 1024) return IBLN_EINVAL;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      int per_sm = (int)((226 * 1024) / (smem + 1024));
      if (per_sm > 48 / w) per_sm = 48 / w;           // <= 48 warps per SM
      if (per_sm < 1) per_sm = 1;
      int64_t grid = (int64_t)num_sms(device) * per_sm;
      if (grid > n) grid = n;
      kern<<<(unsigned)grid, w * 32, smem, (cudaStream_t)stream>>>(raw, z, rays_d, noise, g_weights, g_maps, g_maps_srgb, n, S, g_raw);
      return 0;
    };
    int rc = go(composite_bwd_cta_kernel<8>, 8);
    if (rc != 0) return rc;
    IBLN_RETURN_LAST();
  }
  int Sp = (S + 31) & ~31;
  size_t per_warp = (2 * (size_t)ROW * C + 4 * (size_t)Sp + 32) * sizeof(float);
  // 8-warp CTAs (halved while one CTA does not fit); as many CTAs as 226 KB of shared memory hold (1 KB reserved per CTA).
  // (Picking the CTA shape that maximises resident warps -- e.g. 5 CTAs of 5 warps at S = 192 -- measured slower: 0.57 vs 0.78.)
  int warps = 8;
  while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
  size_t smem = warps * per_warp;
  if (smem > 220 * 1024) return IBLN_EINVAL;
  int per_sm = (int)((226 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<comp_grid(n, device, per_sm, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        raw, z, rays_d, noise, g_weights, g_maps, g_maps_srgb, n, S, C, nc, sigm, g_raw);
    return 0;
  };
  const bool fixed = C == 18 && nc == 3 && sigm == 1, full = S % 32 == 0;
  int rc = fixed ? (full ? launch(composite_bwd_kernel<true, true>) : launch(composite_bwd_kernel<true, false>))
                 : (full ? launch(composite_bwd_kernel<false, true>) : launch(composite_bwd_kernel<false, false>));
  if (rc != 0) return rc;
  IBLN_RETURN_LAST();
}

extern "C" int ibln_depth_fwd(const float* sigma, const float* z, const float* rays_d, int reps, int n, int S,
                              float* depth, float* weights, float* visibility, int device, void* stream) {
  if (n == 0) return 0;
  if (reps < 1 || n < 0 || S < 1 || !sigma || !z || !rays_d || !depth) return IBLN_EINVAL;
  DeviceGuard g(device);
  depth_fwd_kernel<<<comp_grid((int64_t)reps * n, device, 16), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
      sigma, z, rays_d, reps, n, S, depth, weights, visibility);
  IBLN_RETURN_LAST();
}

// Stratified and hierarchical sampling kernels (north-star subsystem 1).
// One warp per ray: warp-level prefix scan for the CDF, binary search in shared memory for the
// inverse, warp bitonic sort for the coarse+fine merge.
#include "common.cuh"

namespace ibln {

// ---------------------------------------------------------------- stratified z
// torch.linspace(0,1,S) is evaluated symmetrically: start+step*i in the lower half and
// end-step*(S-1-i) as ONE fused multiply-add in the upper half (both its CPU and CUDA kernels contract
// it); reproduce that and keep every other op un-contracted so z is bit-identical to
// ibl_nerf_renderer.py:670-692.
__device__ __forceinline__ float linspace01(int i, int s) {
  float step = __fdiv_rn(1.0f, (float)(s - 1));
  return (i < s / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(s - 1 - i), 1.0f);
}
__device__ __forceinline__ float z_lin(float nr, float fr, int i, int s, int lindisp) {
  float t = linspace01(i, s);
  if (!lindisp) return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
  float a = __fmul_rn(__fdiv_rn(1.0f, nr), __fsub_rn(1.0f, t));
  float b = __fmul_rn(__fdiv_rn(1.0f, fr), t);
  return __fdiv_rn(1.0f, __fadd_rn(a, b));
}

__global__ void stratified_z_kernel(const float* __restrict__ nearp, const float* __restrict__ farp,
                                    const float* __restrict__ t_rand, int n, int s, int lindisp,
                                    float* __restrict__ z_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * s) return;
  int r = (int)(idx / s), i = (int)(idx % s);
  float nr = nearp[r], fr = farp[r];
  float zi = z_lin(nr, fr, i, s, lindisp);
  if (t_rand != nullptr && s > 1) {
    float lo = zi, hi = zi;
    if (i > 0) lo = __fmul_rn(0.5f, __fadd_rn(zi, z_lin(nr, fr, i - 1, s, lindisp)));
    if (i < s - 1) hi = __fmul_rn(0.5f, __fadd_rn(z_lin(nr, fr, i + 1, s, lindisp), zi));
    zi = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[idx]));
  }
  z_out[idx] = zi;
}

// ---------------------------------------------------------------- inverse CDF
// searchsorted(cdf, u, right=True) = number of entries <= u.  A per-ray binary search costs log2(n)+1 dependent
// shared-memory probes with random (bank-conflicting) addresses per sample and made the kernel LSU-bound at a
// quarter of the HBM roofline.  Instead every ray gets a 256-cell GUIDE TABLE over the value range [0,1]:
//   G[m] = #{k : floor(256 cdf[k]) <= m}
// (a shared-memory histogram of the nbins entries + one warp prefix scan).  For a sample u in cell m every entry
// counted in G[m-1] is <= u and every entry beyond G[m] is > u, so the answer lies in [G[m-1], G[m]] -- for the
// ~3/4 of the cells that hold no entry it IS G[m], with no probe at all; otherwise a short exact binary search
// over the entries of that one cell.  One 16-byte record {cdf[i-1], cdf[i], bins[i-1], bins[i]} then feeds the
// interpolation.  Indices are bit-identical to torch.searchsorted (the cell tests are exact in fp32).
constexpr int GUIDE_CELLS = 256;

struct RaySmem {            // per-warp carve-up (floats)
  float* cdf;               // [nbins]
  float4* rec;              // [nbins + 1]
  uint32_t* guide;          // [GUIDE_CELLS / 2 + 64] packed u16 inclusive counts (or byte histogram + pair table)
  float* bins;              // [nbins] staging copy (register-prefetching kernel only)
};
__host__ __device__ inline int ray_smem_floats(int nbins) { return 2 * ((nbins + 3) & ~3) + 4 * (nbins + 1) + GUIDE_CELLS / 2 + 64; }
__device__ __forceinline__ RaySmem carve(float* base, int nbins) {
  RaySmem r;
  r.rec = reinterpret_cast<float4*>(base);
  r.cdf = base + 4 * (nbins + 1);
  r.guide = reinterpret_cast<uint32_t*>(r.cdf + ((nbins + 3) & ~3));
  r.bins = reinterpret_cast<float*>(r.guide + GUIDE_CELLS / 2 + 64);
  return r;
}

__device__ __forceinline__ int guide_cell(float v) {
  const int c = (int)__fmul_rn(v, (float)GUIDE_CELLS);       // exact power-of-two scaling, truncation
  return min(GUIDE_CELLS - 1, max(0, c));
}

// cdf[0..nbins) is in shared memory (written by this warp, visible after __syncwarp): build guide + records
__device__ __forceinline__ void warp_build_guide(const RaySmem& sm, const float* __restrict__ bins, int nbins, int lane) {
#pragma unroll
  for (int j = 0; j < GUIDE_CELLS / 64; ++j) sm.guide[lane + 32 * j] = 0u;
  __syncwarp();
  for (int i = lane; i <= nbins; i += 32) {
    const int below = max(i - 1, 0), above = min(i, nbins - 1);
    sm.rec[i] = make_float4(sm.cdf[below], sm.cdf[above], bins[below], bins[above]);
    if (i < nbins) {
      const int c = guide_cell(sm.cdf[i]);
      atomicAdd(&sm.guide[c >> 1], 1u << (16 * (c & 1)));
    }
  }
  __syncwarp();
  // inclusive prefix over 256 packed u16 cells: lane owns words 4*lane .. 4*lane+3 (counts <= 65535)
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { w[j] = sm.guide[4 * lane + j]; w[j] += w[j] << 16; }
#pragma unroll
  for (int j = 1; j < 4; ++j) w[j] += (w[j - 1] >> 16) * 0x10001u;
  uint32_t tot = w[3] >> 16, inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += v;
  }
  const uint32_t base = (inc - tot) * 0x10001u;
#pragma unroll
  for (int j = 0; j < 4; ++j) sm.guide[4 * lane + j] = w[j] + base;
  __syncwarp();
}

// EXACT = true : IEEE division, every op un-contracted -> samples bit-identical to the reference given (cdf, u)
// EXACT = false: one MUFU reciprocal-multiply (<= 2 ulp); used by the fused sample_pdf / hierarchical paths
template <bool EXACT>
__device__ __forceinline__ float invert_one(const RaySmem& sm, float u, int* ind_out) {
  const uint16_t* g16 = reinterpret_cast<const uint16_t*>(sm.guide);
  const int m = guide_cell(u);
  int lo = m ? (int)g16[m - 1] : 0, hi = (int)g16[m];
  while (lo < hi) {                       // only for samples whose cell holds cdf entries
    const int mid = (lo + hi) >> 1;
    if (sm.cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  *ind_out = lo;
  const float4 r = sm.rec[lo];
  float den = __fsub_rn(r.y, r.x);
  if (den < 1e-5f) den = 1.0f;
  const float t = EXACT ? __fdiv_rn(__fsub_rn(u, r.x), den) : __fdividef(__fsub_rn(u, r.x), den);
  return __fadd_rn(r.z, __fmul_rn(t, __fsub_rn(r.w, r.z)));
}

// nbins <= 64 variant (every shipped config): counts fit a byte, so the table is 256 packed (G[m] << 8 | G[m-1])
// 16-bit entries built from two 32-bit histogram words per lane, and ONE 2-byte load gives a sample both ends of
// its search range.  Uses guide[0..63] as the byte histogram and guide[64..191] for the pair table.
__device__ __forceinline__ void warp_build_guide_small(const RaySmem& sm, const float* __restrict__ bins, int nbins, int lane) {
  sm.guide[lane] = 0u;
  sm.guide[lane + 32] = 0u;
  __syncwarp();
  for (int i = lane; i <= nbins; i += 32) {
    const int below = max(i - 1, 0), above = min(i, nbins - 1);
    sm.rec[i] = make_float4(sm.cdf[below], sm.cdf[above], bins[below], bins[above]);
    if (i < nbins) {
      const int c = guide_cell(sm.cdf[i]);
      atomicAdd(&sm.guide[c >> 2], 1u << (8 * (c & 3)));
    }
  }
  __syncwarp();
  uint32_t w0 = sm.guide[2 * lane], w1 = sm.guide[2 * lane + 1];      // cells 8*lane .. 8*lane+7, one byte each
  w0 += w0 << 8; w0 += w0 << 16;
  w1 += w1 << 8; w1 += w1 << 16;
  w1 += (w0 >> 24) * 0x01010101u;
  uint32_t tot = w1 >> 24, inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += v;
  }
  const uint32_t base = inc - tot;            // entries in the cells before this lane's = G[8*lane - 1]
  w0 += base * 0x01010101u;
  w1 += base * 0x01010101u;
  // pair table: entry m = G[m] << 8 | G[m-1]
  const uint32_t prev0 = (w0 << 8) | base, prev1 = (w1 << 8) | (w0 >> 24);     // G[m-1] for the 4 cells of each word
  uint32_t* pair = sm.guide + 64 + 4 * lane;
  pair[0] = __byte_perm(prev0, w0, 0x5140);   // cells 0,1: bytes {prev0.0, w0.0, prev0.1, w0.1}
  pair[1] = __byte_perm(prev0, w0, 0x7362);   // cells 2,3
  pair[2] = __byte_perm(prev1, w1, 0x5140);
  pair[3] = __byte_perm(prev1, w1, 0x7362);
  __syncwarp();
}
template <bool EXACT>
__device__ __forceinline__ float invert_one_small(const RaySmem& sm, float u, int* ind_out) {
  const uint16_t* g16 = reinterpret_cast<const uint16_t*>(sm.guide + 64);
  const uint32_t e = g16[guide_cell(u)];
  int lo = (int)(e & 0xffu), hi = (int)(e >> 8);
  while (lo < hi) {                       // only for samples whose cell holds cdf entries
    const int mid = (lo + hi) >> 1;
    if (sm.cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  *ind_out = lo;
  const float4 r = sm.rec[lo];
  float den = __fsub_rn(r.y, r.x);
  if (den < 1e-5f) den = 1.0f;
  const float t = EXACT ? __fdiv_rn(__fsub_rn(u, r.x), den) : __fdividef(__fsub_rn(u, r.x), den);
  return __fadd_rn(r.z, __fmul_rn(t, __fsub_rn(r.w, r.z)));
}

// Build cdf[0..nbins) in shared memory from nbins-1 weights (warp-cooperative).
// pdf = (w+1e-5)/sum; cdf = [0, cumsum(pdf)]  (nerf_renderer_helper.py:93-96)
__device__ __forceinline__ void warp_build_cdf(const float* __restrict__ w, int nw, float* cdf, int lane) {
  float part = 0.f;
  for (int i = lane; i < nw; i += 32) part += w[i] + 1e-5f;
  const float inv_total = __fdividef(1.0f, warp_sum(part));
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < nw; base += 32) {
    const int i = base + lane;
    float p = (i < nw) ? (w[i] + 1e-5f) * inv_total : 0.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float v = __shfl_up_sync(FULL, p, o);
      if (lane >= o) p += v;
    }
    if (i < nw) cdf[i + 1] = carry + p;
    carry += __shfl_sync(FULL, p, 31);
  }
  __syncwarp();
}

constexpr int SP_WARPS = 4;

template <bool EXACT>
__global__ void __launch_bounds__(SP_WARPS * 32)
sample_pdf_kernel(const float* __restrict__ bins, int64_t bins_stride, const float* __restrict__ weights,
                  int64_t w_stride, const float* __restrict__ cdf_in, const float* __restrict__ u, int n, int nbins,
                  int nsamp, int64_t* __restrict__ inds_out, float* __restrict__ samples) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RaySmem rs = carve(sm + (size_t)warp * ray_smem_floats(nbins), nbins);
  for (int r = blockIdx.x * SP_WARPS + warp; r < n; r += gridDim.x * SP_WARPS) {
    const float* brow = bins + (int64_t)r * bins_stride;
    const float* urow = u + (int64_t)r * nsamp;
    float* orow = samples + (int64_t)r * nsamp;
    if (cdf_in != nullptr) {
      for (int i = lane; i < nbins; i += 32) rs.cdf[i] = cdf_in[(int64_t)r * nbins + i];
      __syncwarp();
    } else {
      warp_build_cdf(weights + (int64_t)r * w_stride, nbins - 1, rs.cdf, lane);
    }
    warp_build_guide(rs, brow, nbins, lane);
    int j = lane;
    for (; j + 96 < nsamp; j += 128) {          // 4 independent lookups in flight per lane
      float uu[4], sv[4];
      int ind[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) uu[q] = urow[j + 32 * q];
#pragma unroll
      for (int q = 0; q < 4; ++q) sv[q] = invert_one<EXACT>(rs, uu[q], &ind[q]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        orow[j + 32 * q] = sv[q];
        if (inds_out != nullptr) inds_out[(int64_t)r * nsamp + j + 32 * q] = ind[q];
      }
    }
    for (; j < nsamp; j += 32) {
      int ind;
      orow[j] = invert_one<EXACT>(rs, urow[j], &ind);
      if (inds_out != nullptr) inds_out[(int64_t)r * nsamp + j] = ind;
    }
    __syncwarp();
  }
}

// Larger rays (64 < nbins <= 32 K, the sweep's 191 / 384 and 511 / 1024 shapes) through the weights -> samples path:
// ncu on sample_pdf_kernel put `long_scoreboard` (exposed global-load latency: weights twice, bins in the guide build, one
// round trip per 128-sample batch of uniforms) at 7 of 13 stalled warps per issue.  Same arithmetic (bit-identical CDF and
// samples), but every global load is issued ahead of its use: the NEXT ray's weights / bins travel in K register slots
// per lane while the current ray is sampled, the next batch of uniforms while the current batch is inverted.
// (Q = 8 lookups per lane per batch measured slower than 4: 0.35 / 0.23 vs 0.38 / 0.26 of the HBM copy peak at 191 / 511 bins.)
template <int K, int Q>     // K register slots per lane for weights / bins, Q lookups per lane in flight per batch
__global__ void __launch_bounds__(SP_WARPS * 32)
sample_pdf_prefetch_kernel(const float* __restrict__ bins, int64_t bins_stride, const float* __restrict__ weights,
                           int64_t w_stride, const float* __restrict__ u, int n, int nbins, int nsamp,
                           float* __restrict__ samples) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RaySmem rs = carve(sm + (size_t)warp * ray_smem_floats(nbins), nbins);
  const int stride = gridDim.x * SP_WARPS;
  const int nw = nbins - 1;
  const int nbatch = (nsamp + 32 * Q - 1) / (32 * Q);
  auto fetch_wb = [&](int r, float (&w)[K], float (&b)[K]) {
    const float* wrow = weights + (int64_t)r * w_stride;
    const float* brow = bins + (int64_t)r * bins_stride;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int i = lane + 32 * k;
      w[k] = (r < n && i < nw) ? wrow[i] : 0.f;
      b[k] = (r < n && i < nbins) ? brow[i] : 0.f;
    }
  };
  auto fetch_u = [&](int r, int batch, float (&uu)[Q]) {
    const float* urow = u + (int64_t)r * nsamp;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int j = batch * 32 * Q + 32 * q + lane;
      uu[q] = j < nsamp ? urow[j] : 0.f;
    }
  };
  int r = blockIdx.x * SP_WARPS + warp;
  float wc[K], bc[K], wn[K], bn[K];
  fetch_wb(r, wc, bc);
  for (; r < n; r += stride) {
    float uc[Q], un[Q];
    fetch_u(r, 0, uc);
    // ---- CDF from the register-held weights: the summation order of warp_build_cdf
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (lane + 32 * k < nw) part += wc[k] + 1e-5f;
    const float inv_total = __fdividef(1.0f, warp_sum(part));
    float carry = 0.f;
    if (lane == 0) rs.cdf[0] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (32 * k < nw) {
        const int i = lane + 32 * k;
        float p = (i < nw) ? (wc[k] + 1e-5f) * inv_total : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float v = __shfl_up_sync(FULL, p, o);
          if (lane >= o) p += v;
        }
        if (i < nw) rs.cdf[i + 1] = carry + p;
        carry += __shfl_sync(FULL, p, 31);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (lane + 32 * k < nbins) rs.bins[lane + 32 * k] = bc[k];
    __syncwarp();
    fetch_wb(r + stride, wn, bn);                   // in flight during the guide build and the sampling of this ray
    warp_build_guide(rs, rs.bins, nbins, lane);
    float* orow = samples + (int64_t)r * nsamp;
    for (int b = 0; b < nbatch; ++b) {
      if (b + 1 < nbatch) fetch_u(r, b + 1, un);
      float sv[Q];
      int ind[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) sv[q] = invert_one<false>(rs, uc[q], &ind[q]);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int j = b * 32 * Q + 32 * q + lane;
        if (j < nsamp) orow[j] = sv[q];
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) uc[q] = un[q];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; ++k) { wc[k] = wn[k]; bc[k] = bn[k]; }
  }
}

// The shipped shape (nbins <= 64, nsamp <= 128: 63 bins / 128 samples): every input of a ray fits in 8 registers
// per lane, so the NEXT ray's weights / bins / uniforms are fetched while the current ray is processed -- the
// generic kernel exposes three dependent global-load round trips per ray (ncu: 53 % long-scoreboard stalls).
struct RayRegs { float w0, w1, b0, b1, u[4]; };
// NB / NS: compile-time nbins / nsamp (0 = run-time); the shipped 63 / 128 instantiation folds every bounds test.
// EXACT instantiations are the explicit-CDF entry (ibln_inverse_cdf: cdf_in given, indices optionally returned);
// the others build the CDF from the weights and return samples only.
template <bool EXACT, int NB, int NS>
__global__ void __launch_bounds__(SP_WARPS * 32)
sample_pdf_small_kernel(const float* __restrict__ bins, int64_t bins_stride, const float* __restrict__ weights,
                        int64_t w_stride, const float* __restrict__ cdf_in, const float* __restrict__ u, int n, int nbins_rt,
                        int nsamp_rt, int64_t* __restrict__ inds_out, float* __restrict__ samples) {
  extern __shared__ __align__(16) float sm[];
  const int nbins = NB ? NB : nbins_rt, nsamp = NS ? NS : nsamp_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RaySmem rs = carve(sm + (size_t)warp * ray_smem_floats(nbins), nbins);
  const int stride = gridDim.x * SP_WARPS;
  const int nw = nbins - 1;
  auto fetch = [&](int r) {
    RayRegs q;
    q.w0 = q.w1 = q.b0 = q.b1 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) q.u[k] = 0.f;
    if (r < n) {
      const float* brow = bins + (int64_t)r * bins_stride;
      const float* urow = u + (int64_t)r * nsamp;
      if (EXACT) {                      // explicit CDF entry (always EXACT): w0/w1 carry cdf[lane], cdf[lane + 32]
        const float* crow = cdf_in + (int64_t)r * nbins;
        if (lane < nbins) q.w0 = crow[lane];
        if (lane + 32 < nbins) q.w1 = crow[lane + 32];
      } else {
        const float* wrow = weights + (int64_t)r * w_stride;
        if (lane < nw) q.w0 = wrow[lane];
        if (lane + 32 < nw) q.w1 = wrow[lane + 32];
      }
      if (lane < nbins) q.b0 = brow[lane];
      if (lane + 32 < nbins) q.b1 = brow[lane + 32];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lane + 32 * k < nsamp) q.u[k] = urow[lane + 32 * k];
    }
    return q;
  };
  int r = blockIdx.x * SP_WARPS + warp;
  RayRegs cur = fetch(r), nx1 = fetch(r + stride);     // two rays in flight: one ray (~400 warp instructions) does not
  for (; r < n; r += stride) {                         // cover the loaded DRAM latency (ncu: 19 % of samples on the first use)
    const RayRegs nx2 = fetch(r + 2 * stride);
    if (EXACT) {
      if (lane < nbins) rs.cdf[lane] = cur.w0;
      if (lane + 32 < nbins) rs.cdf[lane + 32] = cur.w1;
    } else {              // same arithmetic / summation order as warp_build_cdf
      float part = 0.f;
      if (lane < nw) part += cur.w0 + 1e-5f;
      if (lane + 32 < nw) part += cur.w1 + 1e-5f;
      const float inv_total = __fdividef(1.0f, warp_sum(part));
      float p0 = (lane < nw) ? (cur.w0 + 1e-5f) * inv_total : 0.f;
      float p1 = (lane + 32 < nw) ? (cur.w1 + 1e-5f) * inv_total : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v0 = __shfl_up_sync(FULL, p0, o), v1 = __shfl_up_sync(FULL, p1, o);
        if (lane >= o) { p0 += v0; p1 += v1; }
      }
      const float carry = __shfl_sync(FULL, p0, 31);
      if (lane == 0) rs.cdf[0] = 0.f;
      if (lane < nw) rs.cdf[lane + 1] = 0.f + p0;
      if (lane + 32 < nw) rs.cdf[lane + 33] = carry + p1;
    }
    if (lane < nbins) rs.bins[lane] = cur.b0;
    if (lane + 32 < nbins) rs.bins[lane + 32] = cur.b1;
    __syncwarp();
    warp_build_guide_small(rs, rs.bins, nbins, lane);
    float* orow = samples + (int64_t)r * nsamp;
    float sv[4];
    int ind[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sv[k] = invert_one_small<EXACT>(rs, cur.u[k], &ind[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < nsamp) {
        orow[lane + 32 * k] = sv[k];
        if (EXACT && inds_out != nullptr) inds_out[(int64_t)r * nsamp + lane + 32 * k] = ind[k];
      }
    __syncwarp();
    cur = nx1;
    nx1 = nx2;
  }
}

// ---------------------------------------------------------------- merge / sort
__device__ __forceinline__ void warp_bitonic_sort(float* s, int npad, int lane) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (npad >> 1); t += 32) {
        int i = ((t / j) * 2 * j) + (t % j);
        int l = i + j;
        bool up = ((i & k) == 0);
        float a = s[i], b = s[l];
        if ((a > b) == up) { s[i] = b; s[l] = a; }
      }
      __syncwarp();
    }
  }
}

__host__ __device__ inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

__global__ void __launch_bounds__(SP_WARPS * 32)
merge_sort_kernel(const float* __restrict__ za, const float* __restrict__ zb, int n, int sa, int sb, int npad,
                  float* __restrict__ out) {
  extern __shared__ float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s = sm + (size_t)warp * npad;
  int tot = sa + sb;
  for (int r = blockIdx.x * SP_WARPS + warp; r < n; r += gridDim.x * SP_WARPS) {
    for (int i = lane; i < npad; i += 32)
      s[i] = i < sa ? za[(int64_t)r * sa + i] : (i < tot ? zb[(int64_t)r * sb + (i - sa)] : __int_as_float(0x7f800000));
    __syncwarp();
    warp_bitonic_sort(s, npad, lane);
    for (int i = lane; i < tot; i += 32) out[(int64_t)r * tot + i] = s[i];
    __syncwarp();
  }
}

// z mids -> cdf(weights[1:-1]) -> samples -> sort(cat(z, samples)); ibl_nerf_renderer.py:702-707
__host__ __device__ inline int hier_smem_floats(int s0, int npad) { return ray_smem_floats(s0 - 1) + ((s0 + 3) & ~3) + npad; }
__global__ void __launch_bounds__(SP_WARPS * 32)
hierarchical_kernel(const float* __restrict__ z, const float* __restrict__ weights, const float* __restrict__ u, int n,
                    int s0, int s1, int npad, float* __restrict__ z_samples, float* __restrict__ z_merged) {
  extern __shared__ __align__(16) float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int nbins = s0 - 1;
  float* base = sm + (size_t)warp * hier_smem_floats(s0, npad);
  const RaySmem rs = carve(base, nbins);
  float* s_bins = base + ray_smem_floats(nbins);
  float* s_sort = s_bins + ((s0 + 3) & ~3);
  int tot = s0 + s1;
  for (int r = blockIdx.x * SP_WARPS + warp; r < n; r += gridDim.x * SP_WARPS) {
    const float* zr = z + (int64_t)r * s0;
    for (int i = lane; i < s0; i += 32) {
      float zi = zr[i];
      s_sort[i] = zi;
      if (i < nbins) s_bins[i] = __fmul_rn(0.5f, __fadd_rn(zr[i + 1], zi));
    }
    warp_build_cdf(weights + (int64_t)r * s0 + 1, nbins - 1, rs.cdf, lane);
    warp_build_guide(rs, s_bins, nbins, lane);
    for (int j = lane; j < s1; j += 32) {
      int ind;
      float sv = invert_one<false>(rs, u[(int64_t)r * s1 + j], &ind);
      z_samples[(int64_t)r * s1 + j] = sv;
      s_sort[s0 + j] = sv;
    }
    for (int i = tot + lane; i < npad; i += 32) s_sort[i] = __int_as_float(0x7f800000);
    __syncwarp();
    warp_bitonic_sort(s_sort, npad, lane);
    for (int i = lane; i < tot; i += 32) z_merged[(int64_t)r * tot + i] = s_sort[i];
    __syncwarp();
  }
}

static int ray_grid(int n, int device, int warps, int ctas_per_sm) {
  int need = (n + warps - 1) / warps;
  int cap = num_sms(device) * ctas_per_sm;
  return need < cap ? (need > 0 ? need : 1) : cap;
}

}  // namespace ibln

using namespace ibln;

extern "C" int ibln_stratified_z(const float* nearp, const float* farp, const float* t_rand, int n, int s,
                                 int lindisp, float* z_out, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || s < 1 || !nearp || !farp || !z_out) return IBLN_EINVAL;
  DeviceGuard g(device);
  int64_t tot = (int64_t)n * s;
  stratified_z_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nearp, farp, t_rand, n, s, lindisp, z_out);
  IBLN_RETURN_LAST();
}

static int launch_sample_pdf(const float* bins, int64_t bs, const float* w, int64_t ws, const float* cdf, const float* u,
                             int n, int nbins, int nsamp, int64_t* inds, float* samples, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || nbins < 2 || nsamp < 1 || !bins || !u || !samples) return IBLN_EINVAL;
  DeviceGuard g(device);
  if (nbins > 65535) return IBLN_EINVAL;
  size_t smem = (size_t)SP_WARPS * ray_smem_floats(nbins) * sizeof(float);
  if (smem > 200 * 1024) return IBLN_EINVAL;
  int grid = ray_grid(n, device, SP_WARPS, 16);
  if (nbins <= 64 && nsamp <= 128) {     // register-prefetching kernel (N_samples = 64, N_importance = 128 of every shipped config)
    if (cdf != nullptr)
      sample_pdf_small_kernel<true, 0, 0><<<grid, SP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, bs, w, ws, cdf, u, n, nbins, nsamp, inds, samples);
    else if (nbins == 63 && nsamp == 128)
      sample_pdf_small_kernel<false, 63, 128><<<grid, SP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, bs, w, ws, cdf, u, n, nbins, nsamp, inds, samples);
    else
      sample_pdf_small_kernel<false, 0, 0><<<grid, SP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, bs, w, ws, cdf, u, n, nbins, nsamp, inds, samples);
    IBLN_RETURN_LAST();
  }
  if (cdf == nullptr && nbins <= 512) {   // weights -> samples for longer rays: register-prefetching kernel
    auto go = [&](auto kern) -> int {
      if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return (int)e; }
      kern<<<grid, SP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, bs, w, ws, u, n, nbins, nsamp, samples);
      return 0;
    };
    int rc = nbins <= 192 ? go(sample_pdf_prefetch_kernel<6, 4>) : go(sample_pdf_prefetch_kernel<16, 4>);
    if (rc != 0) return rc;
    IBLN_RETURN_LAST();
  }
  if (cdf != nullptr) {   // explicit-CDF entry: bit-exact arithmetic
    if (smem > 48 * 1024) IBLN_CUDA(cudaFuncSetAttribute(sample_pdf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_pdf_kernel<true><<<grid, SP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, bs, w, ws, cdf, u, n, nbins, nsamp, inds, samples);
  } else {
    if (smem > 48 * 1024) IBLN_CUDA(cudaFuncSetAttribute(sample_pdf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_pdf_kernel<false><<<grid, SP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, bs, w, ws, cdf, u, n, nbins, nsamp, inds, samples);
  }
  IBLN_RETURN_LAST();
}

extern "C" int ibln_sample_pdf(const float* bins, int64_t bins_stride, const float* weights, int64_t w_stride,
                               const float* u, int n, int nbins, int nsamp, float* samples, int device, void* stream) {
  if (n == 0) return 0;
  if (!weights) return IBLN_EINVAL;
  return launch_sample_pdf(bins, bins_stride, weights, w_stride, nullptr, u, n, nbins, nsamp, nullptr, samples, device, stream);
}

extern "C" int ibln_inverse_cdf(const float* cdf, const float* bins, const float* u, int n, int nbins, int nsamp,
                                int64_t* inds_out, float* samples, int device, void* stream) {
  if (n == 0) return 0;
  if (!cdf) return IBLN_EINVAL;
  return launch_sample_pdf(bins, nbins, nullptr, 0, cdf, u, n, nbins, nsamp, inds_out, samples, device, stream);
}

extern "C" int ibln_merge_sort_z(const float* za, const float* zb, int n, int sa, int sb, float* z_out, int device,
                                 void* stream) {
  if (n == 0) return 0;
  if (n < 0 || sa < 0 || sb < 0 || sa + sb < 1 || !z_out) return IBLN_EINVAL;
  DeviceGuard g(device);
  int npad = next_pow2(sa + sb);
  if (npad < 2) npad = 2;
  size_t smem = (size_t)SP_WARPS * npad * sizeof(float);
  if (smem > 200 * 1024) return IBLN_EINVAL;
  if (smem > 48 * 1024) IBLN_CUDA(cudaFuncSetAttribute(merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_sort_kernel<<<ray_grid(n, device, SP_WARPS, 16), SP_WARPS * 32, smem, (cudaStream_t)stream>>>(za, zb, n, sa, sb, npad, z_out);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_hierarchical_sample(const float* z, const float* weights, const float* u, int n, int s0, int s1,
                                        float* z_samples, float* z_merged, int device, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || s0 < 4 || s1 < 1 || !z || !weights || !u || !z_samples || !z_merged) return IBLN_EINVAL;
  DeviceGuard g(device);
  int npad = next_pow2(s0 + s1);
  size_t smem = (size_t)SP_WARPS * hier_smem_floats(s0, npad) * sizeof(float);
  if (smem > 200 * 1024) return IBLN_EINVAL;
  if (smem > 48 * 1024) IBLN_CUDA(cudaFuncSetAttribute(hierarchical_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hierarchical_kernel<<<ray_grid(n, device, SP_WARPS, 16), SP_WARPS * 32, smem, (cudaStream_t)stream>>>(z, weights, u, n, s0, s1, npad, z_samples, z_merged);
  IBLN_RETURN_LAST();
}

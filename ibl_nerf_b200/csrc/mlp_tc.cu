// Fused positional-encoding + IBLNeRF MLP on the 5th-generation tensor cores (north-star
// subsystems 2 and 3).
//
// One persistent CTA per SM, 10 warps:
//   warp 0      weight producer : streams the pre-swizzled bf16 weight chunks (16 KB each, packed by
//               ibln_mlp_pack_weights in consumption order) from L2 into a 4-stage shared-memory
//               ring with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx)
//   warp 1      MMA issuer      : one elected thread issues tcgen05.mma (M=128, N=128, K=16, bf16 ->
//               fp32 in TMEM) for two tile slots in ping-pong order; tcgen05.commit frees ring stages
//               and publishes finished accumulators
//   warps 2-5   slot-0 epilogue : generate sample points + positional encoding straight into the
//   warps 6-9   slot-1 epilogue   swizzled A-operand tile, drain TMEM (tcgen05.ld), bias + relu + bf16
//               pack back into the A tile of the next layer, and the small heads (sigma, roughness,
//               albedo, irradiance, radiance x4) as fp32 dot products on the un-rounded accumulators
// Activations never leave the SM; a tile of 128 points flows through all layers while the other
// slot's epilogue overlaps its MMAs.
#include "mlp_tc.cuh"

namespace ibln {
namespace mlp {

// ---------------------------------------------------------------- the fused forward kernel
struct FwdParams {
  const uint8_t* packed;     // chunk stream + const section
  PointGen gen;
  int sigma_only;
  float* out;                // [P] or [P,18]
  uint8_t* saved;            // activation stash (SV_BYTES per tile) or nullptr
  long long n_tiles;
};

__device__ __forceinline__ int n_chunks_of(const Step& st) { return (st.aux_first + st.kb_act + st.aux_last) * (st.n / 128); }

__global__ void __launch_bounds__(N_THREADS, 1) mlp_fwd_kernel(FwdParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* w_full = bars;                 // [N_STAGES]
  uint64_t* w_empty = bars + N_STAGES;     // [N_STAGES]
  uint64_t* act_ready = bars + 2 * N_STAGES;      // [2] epilogue -> MMA
  uint64_t* acc_ready = bars + 2 * N_STAGES + 2;  // [2] MMA -> epilogue
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * N_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = prm.sigma_only ? N_STEPS_SIGMA : N_STEPS_FULL;
  // tiles of this CTA: t = blockIdx.x + k * gridDim.x, k-th tile goes to slot k & 1
  const long long my_tiles = (prm.n_tiles > (long long)blockIdx.x) ? (prm.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&act_ready[i], 4); mbar_init(&acc_ready[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long rounds = (my_tiles + 1) / 2;
      for (long long r = 0; r < rounds; ++r) {
        for (int s = 0; s < n_steps; ++s) {
          const Step st = step_at(s);
          const int nch = n_chunks_of(st);
          for (int slot = 0; slot < 2; ++slot) {
            if (2 * r + slot >= my_tiles) continue;
            for (int c = 0; c < nch; ++c) {
              mbar_wait(&w_empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&w_full[stage], KB_BYTES);
              bulk_g2s(smem + SMEM_RING + stage * KB_BYTES, prm.packed + (size_t)(st.chunk_base + c) * KB_BYTES, KB_BYTES,
                       &w_full[stage]);
              if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t act_phase[2] = {0, 0};
      long long rounds = (my_tiles + 1) / 2;
      for (long long r = 0; r < rounds; ++r) {
        for (int s = 0; s < n_steps; ++s) {
          const Step st = step_at(s);
          const int nh_count = st.n / 128;
          const int nkb = st.aux_first + st.kb_act + st.aux_last;
          for (int slot = 0; slot < 2; ++slot) {
            if (2 * r + slot >= my_tiles) continue;
            mbar_wait(&act_ready[slot], act_phase[slot]);
            act_phase[slot] ^= 1;
            tc_fence_after();
            const uint32_t act_addr = smem_u32(smem + SMEM_ACT + slot * ACT_BYTES);
            const uint32_t aux_addr = smem_u32(smem + SMEM_AUX + slot * AUX_BYTES);
            const uint32_t d_tmem = tmem_base + slot * 256;
            for (int kbi = 0; kbi < nkb; ++kbi) {
              const bool from_aux = (st.aux_first && kbi == 0) || (st.aux_last && kbi == nkb - 1);
              const uint32_t a_addr = from_aux ? aux_addr : act_addr + (kbi - st.aux_first) * KB_BYTES;
              const int ksteps = (st.aux_last && kbi == nkb - 1) ? 2 : 4;
              for (int nh = 0; nh < nh_count; ++nh) {
                mbar_wait(&w_full[stage], phase);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(smem + SMEM_RING + stage * KB_BYTES);
                for (int ks = 0; ks < ksteps; ++ks) {
                  umma_bf16(d_tmem + nh * 128, make_desc_kmajor_sw128(a_addr + ks * 32), make_desc_kmajor_sw128(b_addr + ks * 32),
                            IDESC, (kbi > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&w_empty[stage]);
                if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
              }
            }
            umma_commit(&acc_ready[slot]);
          }
        }
      }
    }
  } else {
    // ===================== epilogue groups =====================
    const int slot = (warp - 2) >> 2;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    uint8_t* act = smem + SMEM_ACT + slot * ACT_BYTES;
    uint8_t* aux = smem + SMEM_AUX + slot * AUX_BYTES;
    const float* cst = reinterpret_cast<const float*>(prm.packed + PACKED_CONST_OFF);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 256;
    uint32_t acc_phase = 0;
    for (long long k = slot; k < my_tiles; k += 2) {
      const long long tile = blockIdx.x + k * gridDim.x;
      const long long p = tile * TILE_M + row;
      const bool valid = p < prm.gen.P;
      float x[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
      if (valid) gen_point(prm.gen, p, x, dir);
      uint8_t* rec = prm.saved ? prm.saved + (size_t)tile * SV_BYTES : nullptr;
      write_encoding<10, 8>(aux, row, x, rec ? rec + (size_t)SV_PE * KB_BYTES : nullptr);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&act_ready[slot]);

      float o_sigma = 0.f, o_rough = 0.f, o_alb[3] = {0.f, 0.f, 0.f}, o_irr = 0.f;
      float o_rad[4][3];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) o_rad[a][c] = 0.f;

      for (int s = 0; s < n_steps; ++s) {
        const Step st = step_at(s);
        mbar_wait(&acc_ready[slot], acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const float* bias = cst + C_BIAS + s * 256;
        const bool write_act = (st.epi == EPI_RELU_ACT) || (st.epi == EPI_FEATURE) || (st.epi == EPI_VIEW) ||
                               (st.epi == EPI_L7 && !prm.sigma_only);
        // stash destination of this step's output tile and its relu-mask slot
        const int sv_blk = s <= 7 ? SV_H(s) : s == 8 ? SV_AF : s == 9 ? SV_FEAT : s == 10 ? SV_HV : s == 11 ? SV_ADDF : SV_ADDF + 4;
        const int sv_mask = s <= 7 ? s : s == 8 ? 8 : s == 10 ? 9 : s == 11 ? 10 : s == 12 ? 11 : -1;
        uint32_t* mask_row = (rec && sv_mask >= 0)
                                 ? reinterpret_cast<uint32_t*>(rec + (size_t)SV_MASK * KB_BYTES + sv_mask * 4096 + row * 32) : nullptr;
        for (int cc = 0; cc < st.n / 32; ++cc) {
          uint32_t v[32];
          tmem_ld32(t_lane + cc * 32, v);
          tmem_wait_ld();
          float h[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 b = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + 4 * j4));
            h[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + b.x;
            h[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b.y;
            h[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b.z;
            h[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b.w;
          }
          if (mask_row != nullptr) {        // bit j = (pre-activation >= 0), gathered from the sign bits
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) m = (m >> 1) | (__float_as_uint(h[j]) & 0x80000000u);
            mask_row[cc] = ~m;
          }
          if (st.epi != EPI_FEATURE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
          }
          if (write_act || rec != nullptr) {
            const int kb = cc >> 1;                    // 64 columns per K-block
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c16 = (cc & 1) * 4 + q;
              uint4 pk = make_uint4(pack_bf16x2(h[8 * q], h[8 * q + 1]), pack_bf16x2(h[8 * q + 2], h[8 * q + 3]),
                                    pack_bf16x2(h[8 * q + 4], h[8 * q + 5]), pack_bf16x2(h[8 * q + 6], h[8 * q + 7]));
              if (write_act) *reinterpret_cast<uint4*>(act + kb * KB_BYTES + swz_offset(row, c16)) = pk;
              if (rec != nullptr) *reinterpret_cast<uint4*>(rec + (size_t)(sv_blk + kb) * KB_BYTES + swz_offset(row, c16)) = pk;
            }
          }
          if (st.epi == EPI_L7) {
            const float2* w = reinterpret_cast<const float2*>(cst + C_SR) + cc * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) { float2 ww = __ldg(w + j); o_sigma = fmaf(h[j], ww.x, o_sigma); o_rough = fmaf(h[j], ww.y, o_rough); }
          } else if (st.epi == EPI_AF) {
            const float4* w = reinterpret_cast<const float4*>(cst + C_AF) + cc * 32;
            if (cc < 4) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { float4 ww = __ldg(w + j); o_alb[0] = fmaf(h[j], ww.x, o_alb[0]); o_alb[1] = fmaf(h[j], ww.y, o_alb[1]); o_alb[2] = fmaf(h[j], ww.z, o_alb[2]); }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) { float4 ww = __ldg(w + j); o_irr = fmaf(h[j], ww.x, o_irr); }
            }
          } else if (st.epi == EPI_VIEW) {
            const float4* w = reinterpret_cast<const float4*>(cst + C_RAD) + cc * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) { float4 ww = __ldg(w + j); o_rad[0][0] = fmaf(h[j], ww.x, o_rad[0][0]); o_rad[0][1] = fmaf(h[j], ww.y, o_rad[0][1]); o_rad[0][2] = fmaf(h[j], ww.z, o_rad[0][2]); }
          } else if (st.epi == EPI_ADD01 || st.epi == EPI_ADD2) {
            const int head = (st.epi == EPI_ADD2) ? 2 : (cc >> 2);       // 128 columns per head
            const float4* w = reinterpret_cast<const float4*>(cst + C_ADD) + head * 128 + (cc & 3) * 32;
            float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { float4 ww = __ldg(w + j); r0 = fmaf(h[j], ww.x, r0); r1 = fmaf(h[j], ww.y, r1); r2 = fmaf(h[j], ww.z, r2); }
            if (head == 0) { o_rad[1][0] += r0; o_rad[1][1] += r1; o_rad[1][2] += r2; }
            else if (head == 1) { o_rad[2][0] += r0; o_rad[2][1] += r1; o_rad[2][2] += r2; }
            else { o_rad[3][0] += r0; o_rad[3][1] += r1; o_rad[3][2] += r2; }
          }
        }
        if (st.epi == EPI_FEATURE) write_encoding<4, 4>(aux, row, dir, rec ? rec + (size_t)SV_DE * KB_BYTES : nullptr);   // view encoding for the next step
        if (s + 1 < n_steps) {
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&act_ready[slot]);
        }
      }
      // ---- outputs
      if (valid) {
        o_sigma += __ldg(cst + C_SR + 512);
        if (prm.sigma_only) {
          prm.out[p] = o_sigma;
        } else {
          float r[18];
          r[0] = o_sigma;
          r[1] = o_alb[0] + __ldg(cst + C_AF + 1024); r[2] = o_alb[1] + __ldg(cst + C_AF + 1025); r[3] = o_alb[2] + __ldg(cst + C_AF + 1026);
          r[4] = o_rough + __ldg(cst + C_SR + 513);
          r[5] = o_irr + __ldg(cst + C_AF + 1027);
#pragma unroll
          for (int c = 0; c < 3; ++c) r[6 + c] = o_rad[0][c] + __ldg(cst + C_RAD + 1024 + c);
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) r[9 + 3 * a + c] = o_rad[1 + a][c] + __ldg(cst + C_ADD + 1536 + 4 * a + c);
          float2* dst = reinterpret_cast<float2*>(prm.out + p * 18);
#pragma unroll
          for (int j = 0; j < 9; ++j) dst[j] = make_float2(r[2 * j], r[2 * j + 1]);
        }
      }
      tc_fence_before();   // order this tile's TMEM reads before the next tile's act_ready arrival
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------- tcgen05 self-test
// D[128,N] = A[128,K] * B[N,K]^T through exactly the operand layout / descriptors used above.
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nkb = K / 64;
  uint8_t* sA = smem;                                 // nkb x [128][64]
  uint8_t* sB = smem + (size_t)nkb * KB_BYTES;        // nkb x (N/128) x [128][64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)nkb * (N / 128) * KB_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  // stage operands (fp32 -> bf16, swizzled)
  for (int e = threadIdx.x; e < 128 * (K / 8); e += blockDim.x) {
    int row = e / (K / 8), cg = e % (K / 8);
    int kb = cg / 8, c16 = cg % 8;
    const float* src = A + (size_t)row * K + cg * 8;
    uint4 v = make_uint4(pack_bf16x2(src[0], src[1]), pack_bf16x2(src[2], src[3]), pack_bf16x2(src[4], src[5]), pack_bf16x2(src[6], src[7]));
    *reinterpret_cast<uint4*>(sA + (size_t)kb * KB_BYTES + swz_offset(row, c16)) = v;
  }
  for (int e = threadIdx.x; e < N * (K / 8); e += blockDim.x) {
    int n = e / (K / 8), cg = e % (K / 8);
    int kb = cg / 8, c16 = cg % 8;
    int nh = n / 128, row = n % 128;
    const float* src = B + (size_t)n * K + cg * 8;
    uint4 v = make_uint4(pack_bf16x2(src[0], src[1]), pack_bf16x2(src[2], src[3]), pack_bf16x2(src[4], src[5]), pack_bf16x2(src[6], src[7]));
    *reinterpret_cast<uint4*>(sB + ((size_t)kb * (N / 128) + nh) * KB_BYTES + swz_offset(row, c16)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    for (int kb = 0; kb < nkb; ++kb)
      for (int nh = 0; nh < N / 128; ++nh)
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a_addr = smem_u32(sA + (size_t)kb * KB_BYTES) + ks * 32;
          uint32_t b_addr = smem_u32(sB + ((size_t)kb * (N / 128) + nh) * KB_BYTES) + ks * 32;
          uint64_t da = make_desc_kmajor_sw128(a_addr), db = make_desc_kmajor_sw128(b_addr);
          if (variant == 1) { da &= ~((uint64_t)0x3FFF << 16); db &= ~((uint64_t)0x3FFF << 16); }   // LBO = 0
          umma_bf16(tmem_base + nh * 128, da, db, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int row = (warp & 3) * 32 + lane;
  for (int cc = 0; cc < N / 32; ++cc) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + cc * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace mlp
}  // namespace ibln

using namespace ibln;
using namespace ibln::mlp;

extern "C" int64_t ibln_mlp_packed_bytes(void) { return PACKED_BYTES; }

extern "C" int ibln_mlp_pack_weights(const float* const* params_host, void* packed, int device, void* stream) {
  if (!params_host || !packed) return IBLN_EINVAL;
  DeviceGuard g(device);
  PackArgs a;
  for (int i = 0; i < 46; ++i) { if (!params_host[i]) return IBLN_EINVAL; a.p[i] = params_host[i]; }
  pack_chunks_kernel<<<N_CHUNKS, 256, 0, (cudaStream_t)stream>>>(a, (uint8_t*)packed);
  pack_consts_kernel<<<(C_TOTAL + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, (float*)((uint8_t*)packed + PACKED_CONST_OFF));
  int rc = launch_pack_bwd(a, (uint8_t*)packed, (cudaStream_t)stream);
  if (rc != 0) return rc;
  IBLN_RETURN_LAST();
}

extern "C" int64_t ibln_mlp_saved_bytes(int64_t n_pts) { return ((n_pts + TILE_M - 1) / TILE_M) * SV_BYTES; }

extern "C" int ibln_mlp_fwd(const void* packed, int mode, const float* pts, const float* rays_o, const float* rays_d,
                            const float* z, int64_t n_rays, int n_samples, float eps, int sigma_only, float* out,
                            void* saved, int device, void* stream) {
  if (!packed || !out || !rays_d || n_rays < 0 || n_samples < 1 || mode < 0 || mode > 2) return IBLN_EINVAL;
  if (mode == 0 && !pts) return IBLN_EINVAL;
  if (mode != 0 && (!rays_o || !z)) return IBLN_EINVAL;
  if (mode == 2 && !sigma_only) return IBLN_EINVAL;
  if (saved && sigma_only) return IBLN_EINVAL;
  if (saved && (reinterpret_cast<uintptr_t>(saved) & 15) != 0) return IBLN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return IBLN_EINVAL;
  if (n_rays == 0) return 0;
  DeviceGuard g(device);
  FwdParams prm;
  prm.packed = (const uint8_t*)packed;
  prm.gen.pts = pts; prm.gen.o = rays_o; prm.gen.d = rays_d; prm.gen.z = z;
  prm.gen.n_rays = n_rays; prm.gen.S = n_samples; prm.gen.eps = eps; prm.gen.mode = mode;
  prm.gen.P = n_rays * n_samples * (mode == 2 ? 4 : 1);
  prm.sigma_only = sigma_only;
  prm.out = out;
  prm.saved = (uint8_t*)saved;
  prm.n_tiles = (prm.gen.P + TILE_M - 1) / TILE_M;
  IBLN_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_REQUEST));
  long long grid = prm.n_tiles < (long long)num_sms(device) ? prm.n_tiles : (long long)num_sms(device);
  mlp_fwd_kernel<<<(unsigned)grid, N_THREADS, SMEM_REQUEST, (cudaStream_t)stream>>>(prm);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_umma_selftest(const float* a, const float* b, float* d, int n, int k, int variant, int device, void* stream) {
  if (!a || !b || !d || (n != 128 && n != 256) || k < 64 || k % 64 != 0 || k > 256) return IBLN_EINVAL;
  DeviceGuard g(device);
  int smem = (k / 64) * KB_BYTES * (1 + n / 128) + 64 + 1024;
  IBLN_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, n, k, variant);
  IBLN_RETURN_LAST();
}

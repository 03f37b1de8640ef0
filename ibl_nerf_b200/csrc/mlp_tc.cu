// Fused positional-encoding + IBLNeRF MLP on the 5th-generation tensor cores (north-star
// subsystems 2 and 3).
//
// Persistent CTA PAIRS (cluster of 2 = the two SMs of a TPC), one CTA per SM, 10 warps each:
//   warp 0      weight producer : streams its half of the pre-swizzled bf16 weight K-blocks (packed by
//               ibln_mlp_pack_weights in consumption order) from L2 into a 4-stage shared-memory ring with
//               tensor-map TMA (cp.async.bulk.tensor.2d.cta_group::2, both CTAs' bytes credited to the
//               leader's mbarrier); a step's K-blocks are streamed once per round and read by both tile slots
//   warp 1      MMA issuer      : (leader CTA) one elected thread issues tcgen05.mma.cta_group::2 (M=256 over
//               the pair's two 128-point tiles, N=256 or 128, K=16, bf16 -> fp32 in TMEM) for two tile slots in
//               ping-pong order; tcgen05.commit (multicast) frees ring stages and publishes accumulators
//   warps 2-5   slot-0 epilogue : generate sample points + positional encoding straight into the
//   warps 6-9   slot-1 epilogue   swizzled A-operand tile, drain TMEM (tcgen05.ld), bias + relu + bf16
//               pack back into the A tile of the next layer, and the small heads (sigma, roughness,
//               albedo, irradiance, radiance x4) as fp32 dot products on the un-rounded accumulators
// Activations never leave the SM; a tile of 128 points flows through all layers while the other
// slot's epilogue overlaps its MMAs.  STASH instantiation (training): every layer's activation tile, the relu
// bit masks and the encodings are also copied to the per-tile stash record for dgrad / wgrad.
#include "mlp_tc.cuh"

namespace ibln {
namespace mlp {

#ifdef IBLN_DIAGNOSTICS
int g_dbg_host = 0;
void* g_timeline = nullptr;
#endif

int make_chunk_stream_map(CUtensorMap* map, const void* base, int n_chunks) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = []() -> EncodeFn {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (encode == nullptr) return IBLN_EINVAL;
  const cuuint64_t dims[2] = {64, (cuuint64_t)n_chunks * 128};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : IBLN_EINVAL;
}

// ---------------------------------------------------------------- the fused forward kernel
struct FwdParams {
  CUtensorMap wmap;          // forward chunk stream as a 2-D tensor (rows of 64 bf16)
  const uint8_t* packed;     // chunk stream + const section
  PointGen gen;
  float* out;                // [P] or [P,18]
  uint8_t* saved;            // activation stash (SV_BYTES per tile) or nullptr
  long long n_tiles;
  unsigned long long* tl;    // optional timeline buffer (diagnostics)
};

__device__ __forceinline__ int n_chunks_of(const Step& st) { return (st.aux_first + st.kb_act + st.aux_last) * (st.n / 128); }

// per-thread epilogue state
struct Heads {     // fp32x2 accumulators (even / odd columns); summed at the end
  float2 sigma, rough, irr, alb[3], rad[4][3];
};
// acc += sum over the chunk's 32 columns of h[col] * row[col], as 16 packed FMAs
__device__ __forceinline__ void dot32(float2& acc, const float (&h)[32], const float* row) {
  const float4* w = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 ww = w[j4];
    acc = ffma2(make_float2(h[4 * j4], h[4 * j4 + 1]), make_float2(ww.x, ww.y), acc);
    acc = ffma2(make_float2(h[4 * j4 + 2], h[4 * j4 + 3]), make_float2(ww.z, ww.w), acc);
  }
}
struct EpiCtx {
  uint8_t* act;            // activation tile of the slot (A operand of the next step)
  const float* bias_s;     // smem: bias row of the current step
  const float* heads_s;    // smem (the slot's idle encoding tile): small-head weight table of the current step
  const float* cst;        // global const section
  uint8_t* rec;            // stash record of the tile (STASH only)
  uint32_t off[8];         // swizzled byte offset of logical 16-byte chunk c in this thread's row
  uint32_t t_lane;         // TMEM address of this thread's lane, column 0 of the slot accumulator
  int row;
};

enum Kind : int { K_RELU_ACT, K_L7_SIGMA, K_L7_FULL, K_AF, K_FEATURE, K_VIEW, K_ADD01, K_ADD2 };

// One 32-column chunk of the accumulator: bias, activation, pack, store, small-head dot products.
// `cc` is a run-time value (the drain loop is NOT unrolled, see drain()); what must be static is: ODD = cc & 1
// (selects the register-held swizzle offsets) and HALF = cc >= 4 (selects the head group of K_AF / K_ADD01).
template <int KIND, bool STASH, bool ODD, int HALF>
__device__ __forceinline__ void process_chunk(const EpiCtx& c, Heads& hd, const uint32_t (&v)[32], int cc, int sv_blk,
                                              uint32_t& mask_word) {
  float h[32];
  const float4* b4 = reinterpret_cast<const float4*>(c.bias_s + cc * 32);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b = b4[j4];
    const float2 r0 = fadd2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), make_float2(b.x, b.y));
    const float2 r1 = fadd2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), make_float2(b.z, b.w));
    h[4 * j4] = r0.x; h[4 * j4 + 1] = r0.y; h[4 * j4 + 2] = r1.x; h[4 * j4 + 3] = r1.y;
  }
  if (STASH && KIND != K_FEATURE) {      // relu bit mask from the sign bits of the pre-activation
    // one funnel shift per column (bit j = sign of h[j]) as four independent 8-long chains: a single 32-long
    // dependent chain costs ~160 cycles of latency per chunk with only two epilogue warps per scheduler
    uint32_t pm[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 7; i >= 0; --i)
#pragma unroll
      for (int q = 0; q < 4; ++q) pm[q] = __funnelshift_l(__float_as_uint(h[8 * q + i]), pm[q], 1);
    const uint32_t m = (pm[0] & 0xffu) | ((pm[1] & 0xffu) << 8) | ((pm[2] & 0xffu) << 16) | (pm[3] << 24);
    mask_word = ~m;
  }
  // (stash mode: the last step's features also go through the activation tile -- its A operand is dead by then)
  constexpr bool kWriteAct = KIND == K_RELU_ACT || KIND == K_L7_FULL || KIND == K_FEATURE || KIND == K_VIEW || (STASH && KIND == K_ADD2);
  constexpr bool kNeedF32Relu = KIND != K_RELU_ACT && KIND != K_FEATURE;     // heads consume fp32 relu(h)
  if (kNeedF32Relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
  }
  if (kWriteAct || (STASH && KIND != K_L7_SIGMA)) {
    const int kb = cc >> 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 pk;
      if (KIND == K_RELU_ACT)
        pk = make_uint4(pack_bf16x2_relu(h[8 * q], h[8 * q + 1]), pack_bf16x2_relu(h[8 * q + 2], h[8 * q + 3]),
                        pack_bf16x2_relu(h[8 * q + 4], h[8 * q + 5]), pack_bf16x2_relu(h[8 * q + 6], h[8 * q + 7]));
      else
        pk = make_uint4(pack_bf16x2(h[8 * q], h[8 * q + 1]), pack_bf16x2(h[8 * q + 2], h[8 * q + 3]),
                        pack_bf16x2(h[8 * q + 4], h[8 * q + 5]), pack_bf16x2(h[8 * q + 6], h[8 * q + 7]));
      const uint32_t o = c.off[(ODD ? 4 : 0) + q];
      if (kWriteAct) *reinterpret_cast<uint4*>(c.act + kb * KB_BYTES + o) = pk;
      else if (c.rec != nullptr)      // STASH only (AF / ADD01): no smem copy exists -> slice-interleaved no-swizzle image, coalesced per warp
        __stcs(reinterpret_cast<uint4*>(c.rec + (size_t)(sv_blk + kb) * KB_BYTES + (c.row >> 5) * 4096 + ((ODD ? 4 : 0) + q) * 512 + (c.row & 31) * 16), pk);
    }
  }
  const float* T = c.heads_s;     // this step's head rows (staged in the idle encoding tile)
  if (KIND == K_L7_SIGMA) {
    dot32(hd.sigma, h, T + cc * 32);
  } else if (KIND == K_L7_FULL) {
    dot32(hd.sigma, h, T + cc * 32);
    dot32(hd.rough, h, T + 256 + cc * 32);
  } else if (KIND == K_AF) {
    if (HALF == 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q) dot32(hd.alb[q], h, T + q * 128 + cc * 32);
    } else {
      dot32(hd.irr, h, T + 384 + (cc - 4) * 32);
    }
  } else if (KIND == K_VIEW) {
#pragma unroll
    for (int q = 0; q < 3; ++q) dot32(hd.rad[0][q], h, T + q * 256 + cc * 32);
  } else if (KIND == K_ADD01) {
    if (HALF == 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q) dot32(hd.rad[1][q], h, T + q * 128 + cc * 32);
    } else {
#pragma unroll
      for (int q = 0; q < 3; ++q) dot32(hd.rad[2][q], h, T + 384 + q * 128 + (cc - 4) * 32);
    }
  } else if (KIND == K_ADD2) {
#pragma unroll
    for (int q = 0; q < 3; ++q) dot32(hd.rad[3][q], h, T + q * 128 + cc * 32);
  }
}

// Drain chunks [CC0, CC0 + NCHUNK) of the accumulator (32 columns each) with the TMEM loads software-pipelined one
// chunk ahead.  In the stash instantiation the chunk-pair loop of the head steps is only unrolled by 2: fully
// unrolled, those six steps were ~120 KB of straight-line code executed once per tile and 30 % of the epilogue
// warps' stall samples there were instruction-fetch misses (stall_no_inst).  Same-box A/B of the stash forward
// (786 k points): fully unrolled 1.63 ms, rolled 1.45 ms, unrolled by 2 1.41 ms.  Everywhere else full unrolling
// wins or ties (inference forward 1.00 ms unrolled or by 2, 1.05 heads rolled, 1.22 all rolled).
constexpr int kDrainUnroll = 2;
template <int KIND, bool STASH> constexpr bool drain_rolled() { return STASH && KIND != K_RELU_ACT; }
template <int KIND, bool STASH, int CC0, int NCHUNK>
__device__ __forceinline__ void drain(const EpiCtx& c, Heads& hd, int sv_blk, int sv_mask) {
  constexpr int HALF = CC0 >= 4 ? 1 : 0;
  uint32_t* mdst = nullptr;
  if (STASH && KIND != K_FEATURE && c.rec != nullptr)       // [mask slot][word 0..7][row]: a warp's store of one word = 4 full sectors
    mdst = reinterpret_cast<uint32_t*>(c.rec + (size_t)SV_MASK * KB_BYTES + sv_mask * 4096) + c.row;
  uint32_t va[32], vb[32];
  tmem_ld32(c.t_lane + CC0 * 32, va);
  auto pair = [&](int cc) {
    uint32_t m0 = 0, m1 = 0;
    tmem_wait_ld();
    tmem_ld32(c.t_lane + (cc + 1) * 32, vb);
    process_chunk<KIND, STASH, false, HALF>(c, hd, va, cc, sv_blk, m0);
    tmem_wait_ld();
    if (cc + 2 < CC0 + NCHUNK) tmem_ld32(c.t_lane + (cc + 2) * 32, va);
    process_chunk<KIND, STASH, true, HALF>(c, hd, vb, cc + 1, sv_blk, m1);
    if (mdst != nullptr) { __stcs(mdst + cc * 128, m0); __stcs(mdst + (cc + 1) * 128, m1); }
  };
  if (drain_rolled<KIND, STASH>()) {
#pragma unroll kDrainUnroll
    for (int cc = CC0; cc < CC0 + NCHUNK; cc += 2) pair(cc);
  } else {
#pragma unroll
    for (int cc = CC0; cc < CC0 + NCHUNK; cc += 2) pair(cc);
  }
}

// CTA pair (cluster of 2 = the two SMs of a TPC): every GEMM step is ONE M=256 tcgen05.mma.cta_group::2 chain over
// the two CTAs' 128-point tiles.  Each CTA streams only ITS half of the weight rows (N/2) into its ring, so the
// per-SM L2->shared weight traffic and the shared-memory B-operand reads are halved -- with all 148 SMs
// re-streaming the full 1.5 MB image per 128 points the chip-wide L2 bandwidth capped the MMA rate at ~2/3.
// Rank 0 issues the MMAs; tcgen05.commit multicasts stage-free / accumulator-ready to both CTAs; rank 1's
// otherwise idle warp 1 relays "my half of the stage has landed" to the leader; the epilogue warps of both CTAs
// arrive on the leader's act_ready barrier.
// NO_AF: the albedo | irradiance feature step (step 8) is skipped -- the reflected-ray march (raw2outputs_simple,
// ibl_nerf_renderer.py:38-68) reads only sigma and the four radiance heads, so 65 536 of the 795 776 MACs per point are
// dead there; output channels 1..3 and 5 then hold the head biases.
template <bool SIGMA_ONLY, bool STASH, bool NO_AF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(N_THREADS, 1) mlp_fwd_kernel(const __grid_constant__ FwdParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* w_full = bars;                 // [N_STAGES] (leader) both CTAs' halves of the stage have landed
  uint64_t* w_empty = bars + N_STAGES;     // [N_STAGES] stage consumed (multicast commit)
  uint64_t* act_ready = bars + 2 * N_STAGES;      // [2] (leader) epilogues of both CTAs -> MMA
  uint64_t* acc_ready = bars + 2 * N_STAGES + 2;  // [2] MMA -> epilogue (multicast commit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * N_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr int n_steps = SIGMA_ONLY ? N_STEPS_SIGMA : N_STEPS_FULL;
  // tiles of this CTA: t = blockIdx.x + k * gridDim.x, k-th tile goes to slot k & 1.  The pair runs in lock step
  // for pair_tiles rounds (= the leader's count); a tile the second CTA does not have is a phantom (no loads/stores).
  const long long lead = (long long)(blockIdx.x & ~1u);
  const long long pair_tiles = (prm.n_tiles > lead) ? (prm.n_tiles - lead + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < N_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&act_ready[i], 8); mbar_init(&acc_ready[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const long long rounds = (pair_tiles + 1) / 2;

  if (warp == 0) {
    // ===================== weight producer (both CTAs: own half of every K-block) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long r = 0; r < rounds; ++r) {
        for (int s = 0; s < n_steps; ++s) {
          if (NO_AF && s == 8) continue;
          const Step st = step_at(s);
          const int nkb = st.aux_first + st.kb_act + st.aux_last;
          const int halves = st.n == 256 ? 2 : 1;      // my N/2 weight rows of a K-block = 128 or 64 rows of 128 B
          // A step whose K-blocks fit the ring is streamed ONCE per round: both tile slots' MMAs read the same stages
          // (the second pass releases them).  Halves the L2 -> shared weight traffic, which in the stash launches
          // competes with the stash stores for the chip-wide L2 throughput.
          const int passes = nkb <= N_STAGES ? 1 : 2;
          for (int slot = 0; slot < passes; ++slot) {
            if (2 * r + slot >= pair_tiles) continue;
            for (int kbi = 0; kbi < nkb; ++kbi) {
              const int row0 = st.n == 256 ? (st.chunk_base + 2 * kbi + (int)rank) * 128 : (st.chunk_base + kbi) * 128 + (int)rank * 64;
              mbar_wait(&w_empty[stage], phase ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&w_full[stage], 2u * halves * (KB_BYTES / 2));   // both CTAs' bytes
              for (int h = 0; h < halves; ++h)
                tma_load_2d_pair(smem + SMEM_RING + stage * KB_BYTES + h * (KB_BYTES / 2), &prm.wmap, 0, row0 + 64 * h, &w_full[stage]);
              if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (rank == 0) {
        // ===================== MMA issuer (leader) =====================
        constexpr uint32_t IDESC256 = make_idesc_bf16(256, 256, 0, 0);
        constexpr uint32_t IDESC128 = make_idesc_bf16(256, 128, 0, 0);
        uint32_t act_phase[2] = {0, 0};
        int tl_n = 0;
        for (long long r = 0; r < rounds; ++r) {
          for (int s = 0; s < n_steps; ++s) {
            if (NO_AF && s == 8) continue;
            const Step st = step_at(s);
            const int nkb = st.aux_first + st.kb_act + st.aux_last;
            const uint32_t idesc = st.n == 256 ? IDESC256 : IDESC128;
            const bool shared = nkb <= N_STAGES;            // see the producer: one weight pass serves both slots
            const bool has1 = 2 * r + 1 < pair_tiles;
            for (int slot = 0; slot < 2; ++slot) {
              if (2 * r + slot >= pair_tiles) continue;
              mbar_wait(&act_ready[slot], act_phase[slot]);
              act_phase[slot] ^= 1;
              tc_fence_after();
              tl_mark(prm.tl, 2048, tl_n, 100 + 2 * s + slot);
              const uint32_t act_addr = smem_u32(smem + SMEM_ACT + slot * ACT_BYTES);
              const uint32_t aux_addr = smem_u32(smem + SMEM_AUX + slot * AUX_BYTES);
              const uint32_t d_tmem = tmem_base + slot * 256;
              const bool release = !shared || slot == 1 || !has1;   // last reader of these stages
              int st_i = stage;
              uint32_t ph_i = phase;
              for (int kbi = 0; kbi < nkb; ++kbi) {
                const bool from_aux = (st.aux_first && kbi == 0) || (st.aux_last && kbi == nkb - 1);
                const uint32_t a_addr = from_aux ? aux_addr : act_addr + (kbi - st.aux_first) * KB_BYTES;
                const int ksteps = (st.aux_last && kbi == nkb - 1) ? 2 : 4;
                mbar_wait(&w_full[st_i], ph_i);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(smem + SMEM_RING + st_i * KB_BYTES);
                for (int ks = 0; ks < ksteps; ++ks)
                  umma_bf16_pair(d_tmem, make_desc_kmajor_sw128(a_addr + ks * 32), make_desc_kmajor_sw128(b_addr + ks * 32),
                                 idesc, (kbi > 0 || ks > 0) ? 1u : 0u);
                if (release) umma_commit_pair(&w_empty[st_i]);      // covers the first pass's reads of the stage too
                if (++st_i == N_STAGES) { st_i = 0; ph_i ^= 1; }
              }
              umma_commit_pair(&acc_ready[slot]);
              if (release) { stage = st_i; phase = ph_i; }
              tl_mark(prm.tl, 2048, tl_n, 200 + 2 * s + slot);
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue groups =====================
    const int slot = (warp - 2) >> 2;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int gtid = threadIdx.x - 64 - slot * 128; // 0..127 inside the group
    uint8_t* aux = smem + SMEM_AUX + slot * AUX_BYTES;
    float* bias_s = reinterpret_cast<float*>(smem + SMEM_BIAS + slot * 1024);
    EpiCtx c;
    c.act = smem + SMEM_ACT + slot * ACT_BYTES;
    c.bias_s = bias_s;
    c.heads_s = reinterpret_cast<const float*>(aux);
    c.cst = reinterpret_cast<const float*>(prm.packed + PACKED_CONST_OFF);
    c.rec = nullptr;
    c.row = row;
    c.t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 256;
#pragma unroll
    for (int q = 0; q < 8; ++q) c.off[q] = swz_offset(row, q);
    uint32_t acc_phase = 0;
    int tl_n = 0;
    unsigned long long* tl = (quarter == 0 && lane == 0) ? prm.tl : nullptr;
    const int tl_base = slot * 1024;
    // bias row of step `s` -> smem (each thread 2 floats); consumed after the group barrier of that step
    // The bias row / head table of step s+1 are FETCHED (global -> registers) before step s is drained and only
    // STORED to shared memory after the group barrier that ends the drain: the L2 latency of these loads is off
    // the publish critical path.
    auto bias_src = [&](int s) { return reinterpret_cast<const float2*>(c.cst + C_BIAS + s * 256) + gtid; };
    auto load_bias = [&](int s) { reinterpret_cast<float2*>(bias_s)[gtid] = __ldg(bias_src(s)); };
    // small-head weight table of a head step -> the slot's encoding tile, which is idle during every head epilogue
    // (positional encoding is dead after L5, the view encoding after the view GEMM)
    auto heads_src = [&](int s) {
      const int off = s == 7 ? C_SR : s == 8 ? C_AF : s == 10 ? C_RAD : s == 11 ? C_ADD : C_ADD + 768;
      return reinterpret_cast<const float4*>(c.cst + off);
    };
    auto heads_n4 = [&](int s) { return SIGMA_ONLY ? 64 : (s == 7 ? 128 : s == 8 ? 128 : s == 12 ? 96 : 192); };
    auto has_heads = [&](int s) { return s == 7 || (!SIGMA_ONLY && (s == 8 || s == 11 || s == 12)); };
    auto load_heads = [&](int s) {
      const float4* src = heads_src(s);
      float4* dst = reinterpret_cast<float4*>(aux);
      for (int i = gtid; i < heads_n4(s); i += 128) dst[i] = __ldg(src + i);
    };
    auto publish = [&]() {
      if (!STASH) fence_proxy_async();       // stash mode: the writers fenced before the group barrier already
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&act_ready[slot], 0);     // the leader's barrier counts both CTAs' warps
    };
    // STASH: after the group barrier a finished shared-memory tile is copied verbatim to the stash record by ONE
    // bulk (TMA) store with an evict-first L2 policy (the 1.5 MB weight image must stay L2-resident); the
    // epilogue warps are the critical path of the CTA-pair pipeline and spend no issue slots on the copy.  The
    // issuing thread waits for the store's shared-memory reads before the tile is overwritten (next drain).
    auto stash_tile = [&](const uint8_t* tile_smem, int blk, int nblk) {
      const uint4* src = reinterpret_cast<const uint4*>(tile_smem);
      uint4* dst = reinterpret_cast<uint4*>(c.rec + (size_t)blk * KB_BYTES);
      if (c.rec == nullptr) return;
      if (gtid == 0) { bulk_s2g_hint(dst, src, (uint32_t)nblk * KB_BYTES, l2_policy_evict_first()); bulk_commit(); }
    };
    // ---- tile pipeline.  The point data and the layer-0 bias row of the NEXT tile are fetched a whole tile ahead,
    // the small-head biases are folded into the accumulators' initial values, and at a tile boundary the next
    // tile's encoding is published BEFORE the finished tile's outputs are stored: the boundary costs one barrier
    // plus the encoding arithmetic instead of three exposed L2 round trips.
    struct Pt { float x[3], dir[3]; bool real, valid; };
    auto fetch_point = [&](long long kk) {
      Pt q;
      const long long tile = blockIdx.x + kk * gridDim.x;
      q.real = kk < pair_tiles && tile < prm.n_tiles;       // phantom tile: keep the pair in lock step, touch no memory
      const long long p = tile * TILE_M + row;
      q.valid = q.real && p < prm.gen.P;
#pragma unroll
      for (int i = 0; i < 3; ++i) { q.x[i] = 0.f; q.dir[i] = 0.f; }
      if (q.valid) gen_point(prm.gen, p, q.x, q.dir);
      return q;
    };
    auto begin_tile = [&](const Pt& q, long long kk, float2 bias0) {
      if (STASH) c.rec = q.real ? prm.saved + (size_t)(blockIdx.x + kk * gridDim.x) * SV_BYTES : nullptr;
      reinterpret_cast<float2*>(bias_s)[gtid] = bias0;
      write_encoding<10, 8>(aux, row, q.x);
      if (STASH) { fence_proxy_async(); named_bar_sync(1 + slot, 128); }
      publish();
      if (STASH) stash_tile(aux, SV_PE, 1);      // off the critical path: overlaps the layer-0 GEMM
    };
    Pt cur = fetch_point(slot);
    if (slot < pair_tiles) begin_tile(cur, slot, __ldg(bias_src(0)));
    for (long long k = slot; k < pair_tiles; k += 2) {
      const long long p = (blockIdx.x + k * gridDim.x) * TILE_M + row;
      const bool valid = cur.valid;
      float dir[3] = {cur.dir[0], cur.dir[1], cur.dir[2]};
      const Pt nxt = fetch_point(k + 2);           // in flight during the whole tile
      const float2 bias0 = __ldg(bias_src(0));

      Heads hd;                                    // accumulators start at the head biases (column 0 of each pair)
      {
        const float4 b_sr = __ldg(reinterpret_cast<const float4*>(c.cst + C_SR + 512));
        hd.sigma = make_float2(b_sr.x, 0.f);
        hd.rough = make_float2(b_sr.y, 0.f);
        if (!SIGMA_ONLY) {
          const float4 b_af = __ldg(reinterpret_cast<const float4*>(c.cst + C_AF + 512));
          const float4 b_rad = __ldg(reinterpret_cast<const float4*>(c.cst + C_RAD + 768));
          hd.alb[0] = make_float2(b_af.x, 0.f); hd.alb[1] = make_float2(b_af.y, 0.f); hd.alb[2] = make_float2(b_af.z, 0.f);
          hd.irr = make_float2(b_af.w, 0.f);
          hd.rad[0][0] = make_float2(b_rad.x, 0.f); hd.rad[0][1] = make_float2(b_rad.y, 0.f); hd.rad[0][2] = make_float2(b_rad.z, 0.f);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float4 b_add = __ldg(reinterpret_cast<const float4*>(c.cst + C_ADD + 1152 + 4 * a));
            hd.rad[1 + a][0] = make_float2(b_add.x, 0.f); hd.rad[1 + a][1] = make_float2(b_add.y, 0.f); hd.rad[1 + a][2] = make_float2(b_add.z, 0.f);
          }
        } else {
          const float2 zero2 = make_float2(0.f, 0.f);
          hd.irr = zero2;
#pragma unroll
          for (int a = 0; a < 3; ++a) hd.alb[a] = zero2;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 3; ++q) hd.rad[a][q] = zero2;
        }
      }

#pragma unroll 1
      for (int s = 0; s < n_steps; ++s) {
        if (NO_AF && s == 8) continue;
        const int sn = (NO_AF && s == 7) ? 9 : s + 1;      // the step that follows this one
        mbar_wait(&acc_ready[slot], acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        tl_mark(tl, tl_base, tl_n, 10 + 2 * s);
        if (STASH && gtid == 0) bulk_wait_read0();   // the previous tile image has left shared memory
        if (!SIGMA_ONLY && s == 10) { if (STASH) named_bar_sync(1 + slot, 128); load_heads(10); }   // view encoding consumed (and stashed): reuse its tile
        // prefetch the next step's constants into registers (consumed after the drain)
        float2 nb = make_float2(0.f, 0.f);
        float4 nh0 = make_float4(0.f, 0.f, 0.f, 0.f), nh1 = nh0;
        if (sn < n_steps) {
          nb = __ldg(bias_src(sn));
          if (has_heads(sn)) {
            const float4* src = heads_src(sn);
            const int n4 = heads_n4(sn);
            if (gtid < n4) nh0 = __ldg(src + gtid);
            if (gtid + 128 < n4) nh1 = __ldg(src + gtid + 128);
          }
        }
        named_bar_sync(1 + slot, 128);       // bias row (+ head table) of this step are in smem
        if (s == 3 || s == 12) tl_mark(tl, tl_base, tl_n, 50 + s);
        switch (s) {
          case 7:
            if (SIGMA_ONLY) drain<K_L7_SIGMA, false, 0, 8>(c, hd, 0, 0);
            else drain<K_L7_FULL, STASH, 0, 8>(c, hd, SV_H(7), 7);
            break;
          case 8: drain<K_AF, STASH, 0, 4>(c, hd, SV_AF, 8); drain<K_AF, STASH, 4, 4>(c, hd, SV_AF, 8); break;
          case 9:
            drain<K_FEATURE, STASH, 0, 8>(c, hd, SV_FEAT, 0);
            write_encoding<4, 4>(aux, row, dir);   // view encoding for step 10
            break;
          case 10: drain<K_VIEW, STASH, 0, 8>(c, hd, SV_HV, 9); break;
          case 11: drain<K_ADD01, STASH, 0, 4>(c, hd, SV_ADDF, 10); drain<K_ADD01, STASH, 4, 4>(c, hd, SV_ADDF, 10); break;
          case 12: drain<K_ADD2, STASH, 0, 4>(c, hd, SV_ADDF + 4, 11); break;
          default: drain<K_RELU_ACT, STASH, 0, 8>(c, hd, SV_H(s), s); break;
        }
        if (sn < n_steps) {
          if (STASH) fence_proxy_async();
          named_bar_sync(1 + slot, 128);     // everyone is done reading this step's bias row / head table; tile complete
          reinterpret_cast<float2*>(bias_s)[gtid] = nb;
          if (has_heads(sn)) {
            const int n4 = heads_n4(sn);
            if (gtid < n4) reinterpret_cast<float4*>(aux)[gtid] = nh0;
            if (gtid + 128 < n4) reinterpret_cast<float4*>(aux)[gtid + 128] = nh1;
          }
          publish();
          tl_mark(tl, tl_base, tl_n, 11 + 2 * s);
          if (STASH) {   // copy the finished tile out while the next GEMM reads it (both only read)
            if (s <= 7) stash_tile(c.act, SV_H(s), 4);
            else if (s == 9) { stash_tile(c.act, SV_FEAT, 4); stash_tile(aux, SV_DE, 1); }
            else if (s == 10) stash_tile(c.act, SV_HV, 4);
          }
        }
      }
      // ---- tile boundary: free the slot's shared tables, start the next tile, then store this tile's outputs
      tl_mark(tl, tl_base, tl_n, 40);
      if (STASH) fence_proxy_async();
      named_bar_sync(1 + slot, 128);         // everyone is done with the last step's bias row / head table
      if (STASH) stash_tile(c.act, SV_ADDF + 4, 2);   // coarse-radiance feature 2 (written into the dead hv tile)
      tc_fence_before();                     // order this tile's TMEM reads before the next tile's act_ready arrival
      tl_mark(tl, tl_base, tl_n, 41);
      if (k + 2 < pair_tiles) begin_tile(nxt, k + 2, bias0);
      tl_mark(tl, tl_base, tl_n, 42);
      if (valid) {
        auto sum2 = [](float2 v) { return v.x + v.y; };
        if (SIGMA_ONLY) {
          prm.out[p] = sum2(hd.sigma);
        } else {
          float r[18];
          r[0] = sum2(hd.sigma);
#pragma unroll
          for (int q = 0; q < 3; ++q) r[1 + q] = sum2(hd.alb[q]);
          r[4] = sum2(hd.rough);
          r[5] = sum2(hd.irr);
#pragma unroll
          for (int q = 0; q < 3; ++q) r[6 + q] = sum2(hd.rad[0][q]);
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int q = 0; q < 3; ++q) r[9 + 3 * a + q] = sum2(hd.rad[1 + a][q]);
          float2* dst = reinterpret_cast<float2*>(prm.out + p * 18);
#pragma unroll
          for (int j = 0; j < 9; ++j) dst[j] = make_float2(r[2 * j], r[2 * j + 1]);
        }
      }
      tl_mark(tl, tl_base, tl_n, 43);
      cur = nxt;
    }
    if (STASH && gtid == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer may still be reading my shared memory / arriving on my barriers
  if (warp == 1) { __syncwarp(); tmem_dealloc_pair(tmem_base, 512); }
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

// ---------------------------------------------------------------- cta_group::2 self-test
// D[256,256] = A[256,K] * B[256,K]^T on a CTA pair: CTA r stages rows 128r.. of A and of B, the leader issues
// the M=256 MMAs, each CTA drains its 128 rows of D.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_pair_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nkb = K / 64;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sA = smem;                                 // nkb x [128][64]
  uint8_t* sB = smem + (size_t)nkb * KB_BYTES;        // nkb x [128][64]
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(sB + (size_t)nkb * KB_BYTES);
  uint64_t* bar_peer = bar_done + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_done + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar_done, 1); mbar_init(bar_peer, 1); fence_barrier_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc_pair(tmem_ptr, 256);
  for (int e = threadIdx.x; e < 128 * (K / 8); e += blockDim.x) {
    int row = e / (K / 8), cg = e % (K / 8);
    int kb = cg / 8, c16 = cg % 8;
    const float* sa = A + (size_t)(rank * 128 + row) * K + cg * 8;
    const float* sb = B + (size_t)(rank * 128 + row) * K + cg * 8;
    *reinterpret_cast<uint4*>(sA + (size_t)kb * KB_BYTES + swz_offset(row, c16)) =
        make_uint4(pack_bf16x2(sa[0], sa[1]), pack_bf16x2(sa[2], sa[3]), pack_bf16x2(sa[4], sa[5]), pack_bf16x2(sa[6], sa[7]));
    *reinterpret_cast<uint4*>(sB + (size_t)kb * KB_BYTES + swz_offset(row, c16)) =
        make_uint4(pack_bf16x2(sb[0], sb[1]), pack_bf16x2(sb[2], sb[3]), pack_bf16x2(sb[4], sb[5]), pack_bf16x2(sb[6], sb[7]));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // barriers initialised and TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (rank == 1 && threadIdx.x == 0) mbar_arrive_cluster(bar_peer, 0);     // my operand halves are staged
  if (rank == 0 && warp == 1 && elect_one()) {
    mbar_wait(bar_peer, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(256, 256, 0, 0);
    for (int kb = 0; kb < nkb; ++kb)
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16_pair(tmem_base, make_desc_kmajor_sw128(smem_u32(sA + (size_t)kb * KB_BYTES) + ks * 32),
                       make_desc_kmajor_sw128(smem_u32(sB + (size_t)kb * KB_BYTES) + ks * 32), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
    umma_commit_pair(bar_done);
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const int row = (warp & 3) * 32 + lane;
  for (int cc = 0; cc < 8; ++cc) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + row) * 256 + cc * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { __syncwarp(); tmem_dealloc_pair(tmem_base, 256); }
}

}  // namespace mlp
}  // namespace ibln

using namespace ibln;
using namespace ibln::mlp;

#ifdef IBLN_DIAGNOSTICS
extern "C" int ibln_debug_timeline(void* device_buf) { ibln::mlp::g_timeline = device_buf; return 0; }
extern "C" int ibln_debug_set(int flags) { ibln::mlp::g_dbg_host = flags; return 0; }
#endif

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
  if (n_rays == 0) return 0;
  if (!packed || !out || !rays_d || n_rays < 0 || n_samples < 1 || mode < 0 || mode > 2) return IBLN_EINVAL;
  if (mode == 0 && !pts) return IBLN_EINVAL;
  if (mode != 0 && (!rays_o || !z)) return IBLN_EINVAL;
  if (sigma_only < 0 || sigma_only > 2) return IBLN_EINVAL;
  if (mode == 2 && sigma_only != 1) return IBLN_EINVAL;
  if (saved && sigma_only) return IBLN_EINVAL;
  if (saved && (reinterpret_cast<uintptr_t>(saved) & 15) != 0) return IBLN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return IBLN_EINVAL;
  DeviceGuard g(device);
  FwdParams prm;
  prm.packed = (const uint8_t*)packed;
  prm.gen.pts = pts; prm.gen.o = rays_o; prm.gen.d = rays_d; prm.gen.z = z;
  prm.gen.n_rays = n_rays; prm.gen.S = n_samples; prm.gen.eps = eps; prm.gen.mode = mode;
  prm.gen.P = n_rays * n_samples * (mode == 2 ? 4 : 1);
  prm.out = out;
  prm.saved = (uint8_t*)saved;
  prm.n_tiles = (prm.gen.P + TILE_M - 1) / TILE_M;
  prm.tl = (unsigned long long*)IBLN_DBG_TIMELINE;
  { int rc = make_chunk_stream_map(&prm.wmap, packed, N_CHUNKS); if (rc != 0) return rc; }
  long long grid = (long long)(num_sms(device) & ~1);            // CTA pairs
  if (((prm.n_tiles + 1) & ~1LL) < grid) grid = (prm.n_tiles + 1) & ~1LL;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_REQUEST);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)grid, N_THREADS, SMEM_REQUEST, (cudaStream_t)stream>>>(prm);
    return (int)cudaGetLastError();
  };
  if (sigma_only == 1) return launch(mlp_fwd_kernel<true, false>);
  if (sigma_only == 2) return launch(mlp_fwd_kernel<false, false, true>);
  if (saved) return launch(mlp_fwd_kernel<false, true>);
  return launch(mlp_fwd_kernel<false, false>);
}

extern "C" int ibln_umma_pair_selftest(const float* a, const float* b, float* d, int k, int device, void* stream) {
  if (!a || !b || !d || k < 64 || k % 64 != 0 || k > 256) return IBLN_EINVAL;
  DeviceGuard g(device);
  int smem = 2 * (k / 64) * KB_BYTES + 64 + 1024;
  IBLN_CUDA(cudaFuncSetAttribute(umma_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_pair_selftest_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(a, b, d, k);
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

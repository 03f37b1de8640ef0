// Tensor-core backward of the IBLNeRF MLP (north-star subsystem 3, gradients).
//
//  dgrad kernel : same warp-specialised CTA-pair structure as the forward kernel (tensor-map TMA weight
//                 producer streaming each step's K-blocks once per round for both tile slots, single-thread
//                 tcgen05.mma.cta_group::2 issuer, two ping-pong tile slots with 4 epilogue warps
//                 each).  Walks the layers in reverse: dX = dY * W as M=256 (pair) x N=256 GEMMs against
//                 the TRANSPOSED packed weight stream, applies the relu bit masks stashed by the
//                 forward pass, keeps the running gradient tile in shared memory as the next A
//                 operand, and writes every dY tile (bf16, operand layout) for the wgrad kernel.
//                 The 18-channel head gradients enter as fp32 rank-<=3 updates on the CUDA cores.
//  wgrad kernel : dW[out,in] += sum_points dY[pt,out] * X[pt,in] as a split-K tcgen05 GEMM whose
//                 operands are the stashed tiles read as MN-major UMMA operands (no transposes);
//                 fp32 accumulators live in TMEM for the whole K loop (all points of the CTA) and
//                 are flushed once with red.global.add.f32 into the flat gradient image.
#include <cstring>
#include "mlp_tc.cuh"

namespace ibln {
namespace mlp {

// ---------------------------------------------------------------- dgrad step program
constexpr int N_STEPS_BWD = 12;
struct BStep { int nkb; int accumulate; int chunk_base; };
__host__ __device__ constexpr BStep bstep_at(int t) {
  return t == 0 ? BStep{4, 0, 0} : t == 1 ? BStep{2, 1, 8} : t == 4 ? BStep{4, 1, 28} : BStep{4, 0, 12 + 8 * (t - 2)};
}
static_assert(bstep_at(11).chunk_base + 8 == N_CHUNKS_BWD, "bwd chunk table");

// chunk element (n, k) = W_src[row0 + k][col0 + n]
struct BChunkSrc { int param; int ld; int row0; int col0; };
__host__ __device__ inline BChunkSrc bchunk_src(int chunk) {
  for (int t = 0; t < N_STEPS_BWD; ++t) {
    BStep st = bstep_at(t);
    int local = chunk - st.chunk_base;
    if (local < 0 || local >= st.nkb * 2) continue;
    int j = local >> 1, nh = local & 1;
    BChunkSrc c;
    c.col0 = nh * 128;
    c.ld = 256;
    c.row0 = 64 * j;
    if (t == 0) { c.param = 34 + 2 * (j >> 1); c.row0 = 64 * (j & 1); }
    else if (t == 1) c.param = 38;
    else if (t == 2) { c.param = 16; c.ld = 283; }
    else if (t == 3) c.param = 18;
    else if (t == 4) { c.param = j < 2 ? 22 : 28; c.row0 = 64 * (j & 1); }
    else {
      int l = 12 - t;
      c.param = 2 * l;
      if (l == 5) { c.ld = 319; c.col0 += 63; }
    }
    return c;
  }
  return BChunkSrc{0, 0, 0, 0};
}

__device__ __forceinline__ void pack_bwd_chunk_body(const PackArgs& a, uint8_t* __restrict__ out, int chunk) {
  BChunkSrc c = bchunk_src(chunk);
  const float* W = a.p[c.param];
  for (int e = threadIdx.x; e < 128 * 8; e += blockDim.x) {
    int n = e & 127, c16 = e >> 7;        // consecutive threads -> consecutive n (coalesced reads of W rows)
    uint32_t w[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int k = c16 * 8 + 2 * jj;
      float lo = W[(int64_t)(c.row0 + k) * c.ld + c.col0 + n];
      float hi = W[(int64_t)(c.row0 + k + 1) * c.ld + c.col0 + n];
      w[jj] = pack_bf16x2(lo, hi);
    }
    *reinterpret_cast<uint4*>(out + (size_t)chunk * KB_BYTES + swz_offset(n, c16)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

static __global__ void pack_bwd_chunks_kernel(PackArgs a, uint8_t* __restrict__ out) { pack_bwd_chunk_body(a, out, blockIdx.x); }

int launch_pack_bwd(const PackArgs& a, uint8_t* packed, cudaStream_t stream) {
  pack_bwd_chunks_kernel<<<N_CHUNKS_BWD, 256, 0, stream>>>(a, packed + PACKED_BWD_OFF);
  return (int)cudaGetLastError();
}

// flat gradient image: state-dict order, weight [out,in] row-major then bias [out]
struct FlatOff { int w[23]; int b[23]; };
__host__ __device__ inline FlatOff flat_offsets() {
  const int outs[23] = {256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 1, 128, 3, 1, 128, 1, 3, 128, 128, 128, 3, 3, 3};
  const int ins[23] = {63, 256, 256, 256, 256, 319, 256, 256, 283, 256, 256, 256, 128, 256, 256, 128, 256, 256, 256, 256, 128, 128, 128};
  FlatOff f;
  int off = 0;
  for (int i = 0; i < 23; ++i) { f.w[i] = off; off += outs[i] * ins[i]; f.b[i] = off; off += outs[i]; }
  return f;
}
// indices into the 23 Linear layers: 0..7 positions, 8 views, 9 feature, 10 sigma, 11 albedo_f, 12 albedo,
// 13 rough, 14 irr_f, 15 irr, 16 rad, 17..19 add_f, 20..22 add
constexpr int FLAT_TOTAL = 798994;

// Re-pack straight from the optimizer's flat parameter buffer (training.FlatParameters): blockIdx.y = network,
// blockIdx.x walks the forward chunks, then the constant section, then the transposed (dgrad) chunks.
constexpr int PACK_CONST_BLOCKS = (C_TOTAL + 255) / 256;
static __global__ void __launch_bounds__(256) pack_flat_kernel(PackFlat pf) {
  const float* base = pf.flat[blockIdx.y];
  uint8_t* out = pf.packed[blockIdx.y];
  PackArgs a;
  const FlatOff fo = flat_offsets();
#pragma unroll
  for (int i = 0; i < 23; ++i) { a.p[2 * i] = base + fo.w[i]; a.p[2 * i + 1] = base + fo.b[i]; }
  int b = blockIdx.x;
  if (b < N_CHUNKS) { pack_chunk_body(a, out, b); return; }
  b -= N_CHUNKS;
  if (b < PACK_CONST_BLOCKS) { pack_consts_body(a, reinterpret_cast<float*>(out + PACKED_CONST_OFF), b * 256 + (int)threadIdx.x); return; }
  pack_bwd_chunk_body(a, out + PACKED_BWD_OFF, b - PACK_CONST_BLOCKS);
}

int launch_pack_flat(const PackFlat& pf, int n_nets, cudaStream_t stream) {
  pack_flat_kernel<<<dim3(N_CHUNKS + PACK_CONST_BLOCKS + N_CHUNKS_BWD, n_nets), 256, 0, stream>>>(pf);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- dgrad kernel
struct DgradParams {
  CUtensorMap wmap;          // transposed (dgrad) chunk stream as a 2-D tensor
  const uint8_t* packed;
  const uint8_t* saved;      // forward stash (masks)
  const float* g_out;        // [P,18]
  uint8_t* dy;               // DY_BYTES per tile
  float* flat_grad;          // head-bias gradients are accumulated here directly
  long long P;
  long long n_tiles;
  int dbg;
  unsigned long long* tl;    // optional timeline buffer (diagnostics): block 0 appends (tag << 48 | clock64)
};

// v[0..31] += a * row[0..31] as 16 packed fp32x2 FMAs
__device__ __forceinline__ void axpy32(float (&v)[32], float a, const float* row) {
  const float4* w = reinterpret_cast<const float4*>(row);
  const float2 aa = make_float2(a, a);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 ww = w[j4];
    const float2 r0 = ffma2(aa, make_float2(ww.x, ww.y), make_float2(v[4 * j4], v[4 * j4 + 1]));
    const float2 r1 = ffma2(aa, make_float2(ww.z, ww.w), make_float2(v[4 * j4 + 2], v[4 * j4 + 3]));
    v[4 * j4] = r0.x; v[4 * j4 + 1] = r0.y; v[4 * j4 + 2] = r1.x; v[4 * j4 + 3] = r1.y;
  }
}

__device__ __forceinline__ void store_tile4(uint8_t* tile, int kb, uint32_t swz, uint4 pk) {
  *reinterpret_cast<uint4*>(tile + (size_t)kb * KB_BYTES + swz) = pk;
}

// CTA pair like the forward kernel (see mlp_tc.cu): M=256 cta_group::2 MMAs over the two CTAs' tiles, each CTA
// streams half of the transposed weight rows, the leader issues, commits are multicast.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(N_THREADS, 1) mlp_dgrad_kernel(const __grid_constant__ DgradParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + N_STAGES;
  uint64_t* act_ready = bars + 2 * N_STAGES;
  uint64_t* acc_ready = bars + 2 * N_STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * N_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long lead = (long long)(blockIdx.x & ~1u);
  const long long pair_tiles = (prm.n_tiles > lead) ? (prm.n_tiles - lead + gridDim.x - 1) / gridDim.x : 0;
  const long long rounds = (pair_tiles + 1) / 2;

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

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long r = 0; r < rounds; ++r)
        for (int t = 0; t < N_STEPS_BWD; ++t) {
          const BStep st = bstep_at(t);
          static_assert(N_STAGES >= 4, "every dgrad step (<= 4 K-blocks) must fit the ring");
          // every step's K-blocks fit the ring: streamed ONCE per round, both tile slots' MMAs read the same stages
          // (mlp_tc.cu, forward producer)
          {
            for (int kbi = 0; kbi < st.nkb; ++kbi) {
              const int row0 = (st.chunk_base + 2 * kbi + (int)rank) * 128;     // my 128 of the 256 output rows
              mbar_wait(&w_empty[stage], phase ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&w_full[stage], 2u * KB_BYTES);   // both CTAs' halves
              for (int h = 0; h < 2; ++h)
                tma_load_2d_pair(smem + SMEM_RING + stage * KB_BYTES + h * (KB_BYTES / 2), &prm.wmap, 0, row0 + 64 * h, &w_full[stage]);
              if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t IDESC256 = make_idesc_bf16(256, 256, 0, 0);
      int stage = 0;
      int tl_n = 0;
      uint32_t phase = 0;
      uint32_t act_phase[2] = {0, 0};
      for (long long r = 0; r < rounds; ++r)
        for (int t = 0; t < N_STEPS_BWD; ++t) {
          const BStep st = bstep_at(t);
          for (int slot = 0; slot < 2; ++slot) {
            if (2 * r + slot >= pair_tiles) continue;
            mbar_wait(&act_ready[slot], act_phase[slot]);
            act_phase[slot] ^= 1;
            tc_fence_after();
            tl_mark(prm.tl, 2048, tl_n, 100 + t * 2 + slot);
            const uint32_t act_addr = smem_u32(smem + SMEM_ACT + slot * ACT_BYTES);
            const uint32_t d_tmem = tmem_base + slot * 256;
            const bool release = slot == 1 || 2 * r + 1 >= pair_tiles;     // last reader of these stages
            int st_i = stage;
            uint32_t ph_i = phase;
            for (int kbi = 0; kbi < st.nkb; ++kbi) {
              mbar_wait(&w_full[st_i], ph_i);
              tc_fence_after();
              const uint32_t b_addr = smem_u32(smem + SMEM_RING + st_i * KB_BYTES);
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16_pair(d_tmem, make_desc_kmajor_sw128(act_addr + kbi * KB_BYTES + ks * 32),
                               make_desc_kmajor_sw128(b_addr + ks * 32), IDESC256, (st.accumulate || kbi > 0 || ks > 0) ? 1u : 0u);
              if (release) umma_commit_pair(&w_empty[st_i]);
              if (++st_i == N_STAGES) { st_i = 0; ph_i ^= 1; }
            }
            if (release) { stage = st_i; phase = ph_i; }
            umma_commit_pair(&acc_ready[slot]);
            tl_mark(prm.tl, 2048, tl_n, 200 + t * 2 + slot);
          }
        }
    }
  } else {
    const int slot = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint8_t* act = smem + SMEM_ACT + slot * ACT_BYTES;
    const float* cst = reinterpret_cast<const float*>(prm.packed + PACKED_CONST_OFF);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 256;
    const FlatOff fo = flat_offsets();
    uint32_t acc_phase = 0;
    // small-head weight tables -> the slot's (otherwise unused) encoding tile, once per kernel:
    // channel-major rows: [coarse radiance k: 3x128 each, 1152][albedo 3x128 | irradiance 128: 512][radiance 3x256: 768][sigma, rough: 512]
    const float* tab = reinterpret_cast<const float*>(smem + SMEM_AUX + slot * AUX_BYTES);
    const float* T_ADD = tab; const float* T_AF = tab + 1152; const float* T_RAD = tab + 1664; const float* T_SR = tab + 2432;
    const int gtid = threadIdx.x - 64 - slot * 128;
    {
      float4* dst = reinterpret_cast<float4*>(smem + SMEM_AUX + slot * AUX_BYTES);
      for (int i = gtid; i < 736; i += 128) {
        const float* src = i < 288 ? cst + C_ADD + 4 * i : i < 416 ? cst + C_AF + 4 * (i - 288)
                         : i < 608 ? cst + C_RAD + 4 * (i - 416) : cst + C_SR + 4 * (i - 608);
        dst[i] = __ldg(reinterpret_cast<const float4*>(src));
      }
      named_bar_sync(1 + slot, 128);
    }
    uint32_t off[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) off[q] = swz_offset(row, q);
    int tl_n = 0;
    unsigned long long* tl = (quarter == 0 && lane == 0) ? prm.tl : nullptr;
    const int tl_base = slot * 1024;
    for (long long k = slot; k < pair_tiles; k += 2) {
      const long long tile = blockIdx.x + k * gridDim.x;
      const bool real = tile < prm.n_tiles;      // phantom tile (second CTA of the last pair round): no memory traffic
      const long long p = tile * TILE_M + row;
      const bool valid = real && p < prm.P;
      uint8_t* dy = prm.dy + (size_t)(real ? tile : 0) * DY_BYTES;
      const uint8_t* rec = prm.saved + (size_t)(real ? tile : 0) * SV_BYTES;
      const uint32_t* masks = reinterpret_cast<const uint32_t*>(rec + (size_t)SV_MASK * KB_BYTES) + row;   // + m * 1024 + word * 128
      float g[18];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float2 v = valid ? __ldg(reinterpret_cast<const float2*>(prm.g_out + p * 18) + j) : make_float2(0.f, 0.f);
        g[2 * j] = v.x; g[2 * j + 1] = v.y;
      }
      // ---- head-bias gradients: column sums of g over the warp's 32 rows
      if (real) {
        float s[18];
#pragma unroll
        for (int j = 0; j < 18; ++j) s[j] = warp_sum(g[j]);
        if (lane == 0) {
          atomicAdd(prm.flat_grad + fo.b[10], s[0]);
#pragma unroll
          for (int c = 0; c < 3; ++c) atomicAdd(prm.flat_grad + fo.b[12] + c, s[1 + c]);
          atomicAdd(prm.flat_grad + fo.b[13], s[4]);
          atomicAdd(prm.flat_grad + fo.b[15], s[5]);
#pragma unroll
          for (int c = 0; c < 3; ++c) atomicAdd(prm.flat_grad + fo.b[16] + c, s[6 + c]);
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(prm.flat_grad + fo.b[20 + a] + c, s[9 + 3 * a + c]);
        }
      }
      // ---- G tile: g_raw as bf16 [128][64] (columns 18.. zero) for the small-head wgrads
      if (real) {
        float e[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) e[j] = j < 18 ? g[j] : 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          store_tile4(dy + (size_t)DY_G * KB_BYTES, 0, swz_offset(row, ch),
                      make_uint4(pack_bf16x2(e[8 * ch], e[8 * ch + 1]), pack_bf16x2(e[8 * ch + 2], e[8 * ch + 3]),
                                 pack_bf16x2(e[8 * ch + 4], e[8 * ch + 5]), pack_bf16x2(e[8 * ch + 6], e[8 * ch + 7])));
      }
      // ---- phase writers for the CUDA-core produced gradient tiles
      // heads: 0 = coarse radiance 0|1 (256 cols), 1 = coarse radiance 2 (128 cols), 2 = albedo|irradiance features
      // relu bit masks of the phase that writes the NEXT tile are fetched one phase ahead (the 32-byte row load is an
      // exposed DRAM round trip otherwise).  Phase t >= 0 is step t; slot of the mask record it applies:
      auto mask_slot = [](int t) { return t == 0 ? 11 : t == 1 ? 9 : t == 2 ? -1 : t == 3 ? 8 : t == 4 ? 7 : 11 - t; };
      auto fetch_masks = [&](int mslot, uint4& a, uint4& b) {
        a = b = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (mslot >= 0 && real) {
          const uint32_t* m = masks + mslot * 1024;      // word-major: a warp's load of one word is 128 contiguous bytes
          a = make_uint4(__ldg(m), __ldg(m + 128), __ldg(m + 256), __ldg(m + 384));
          b = make_uint4(__ldg(m + 512), __ldg(m + 640), __ldg(m + 768), __ldg(m + 896));
        }
      };
      uint4 cm0, cm1, nm0, nm1;               // masks of the current / next phase
      fetch_masks(10, cm0, cm1);
      fetch_masks(mask_slot(0), nm0, nm1);
      auto write_head_tile = [&](int which) {
        const int ncc = which == 1 ? 4 : 8;
        const uint32_t mw[8] = {cm0.x, cm0.y, cm0.z, cm0.w, cm1.x, cm1.y, cm1.z, cm1.w};
        for (int cc = 0; cc < ncc; ++cc) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          if (which == 2) {
            if (cc < 4) {
#pragma unroll
              for (int q = 0; q < 3; ++q) axpy32(v, g[1 + q], T_AF + q * 128 + cc * 32);
            } else {
              axpy32(v, g[5], T_AF + 384 + (cc - 4) * 32);
            }
          } else {
            const int head = which == 1 ? 2 : (cc >> 2);
#pragma unroll
            for (int q = 0; q < 3; ++q) axpy32(v, g[9 + 3 * head + q], T_ADD + head * 384 + q * 128 + (cc & 3) * 32);
          }
          const uint32_t m = mw[cc];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_tile4(act, cc >> 1, off[(cc & 1) * 4 + q],
                        make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
        }
      };
      // finished gradient tile in `act`: hand it to the MMA issuer (one arrival per warp on the LEADER's barrier) and
      // copy it verbatim to the dY record with one bulk (TMA) store (evict-first: the weight image stays in L2)
      auto publish = [&](int dblk, int nblk, bool arrive) {
        fence_proxy_async();
        named_bar_sync(1 + slot, 128);
        if (arrive) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&act_ready[slot], 0);
        }
        if (real && gtid == 0) {
          bulk_s2g_hint(dy + (size_t)dblk * KB_BYTES, act, (uint32_t)nblk * KB_BYTES, l2_policy_evict_first());
          bulk_commit();
        }
      };
      // before overwriting `act`: the bulk store of the previous tile has finished reading it
      auto pre_write = [&]() {
        if (gtid == 0) bulk_wait_read0();
        named_bar_sync(1 + slot, 128);
      };
      pre_write();
      tl_mark(tl, tl_base, tl_n, 1);
      write_head_tile(0);
      publish(DY_ADDF01, 4, true);
      tl_mark(tl, tl_base, tl_n, 2);
      for (int t = 0; t < N_STEPS_BWD; ++t) {
        mbar_wait(&acc_ready[slot], acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        tl_mark(tl, tl_base, tl_n, 10 + 2 * t);
        cm0 = nm0; cm1 = nm1;
        if (t + 1 < N_STEPS_BWD) fetch_masks(mask_slot(t + 1), nm0, nm1);
        pre_write();
        if (t == 0) { write_head_tile(1); publish(DY_ADDF2, 2, true); tl_mark(tl, tl_base, tl_n, 11 + 2 * t); continue; }
        if (t == 3) { write_head_tile(2); publish(DY_AF, 4, true); tl_mark(tl, tl_base, tl_n, 11 + 2 * t); continue; }
        // drain: t=1 -> dY_view (mask HV, + radiance term); t=2 -> dY_feat (no mask); t=4 -> dY_7 (+ sigma/rough terms);
        // t>=5 -> dY_{11-t} (mask h_{11-t})
        const int dblk = t == 1 ? DY_VIEW : t == 2 ? DY_FEAT : t == 4 ? DY_H(7) : DY_H(11 - t);
        const uint32_t mw[8] = {cm0.x, cm0.y, cm0.z, cm0.w, cm1.x, cm1.y, cm1.z, cm1.w};
        const bool last = (t == N_STEPS_BWD - 1);
        auto chunk = [&](const uint32_t (&raw)[32], int cc) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (t == 1) {
#pragma unroll
            for (int q = 0; q < 3; ++q) axpy32(v, g[6 + q], T_RAD + q * 256 + cc * 32);
          } else if (t == 4) {
            axpy32(v, g[0], T_SR + cc * 32);
            axpy32(v, g[4], T_SR + 256 + cc * 32);
          }
          const uint32_t m = mw[cc];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            store_tile4(act, cc >> 1, off[(cc & 1) * 4 + q],
                        make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
        };
        uint32_t ra[32], rb[32];
        tmem_ld32(t_lane, ra);
#pragma unroll
        for (int cc = 0; cc < 8; cc += 2) {
          tmem_wait_ld();
          tmem_ld32(t_lane + (cc + 1) * 32, rb);
          chunk(ra, cc);
          tmem_wait_ld();
          if (cc + 2 < 8) tmem_ld32(t_lane + (cc + 2) * 32, ra);
          chunk(rb, cc + 1);
        }
        publish(dblk, 4, !last);
        tl_mark(tl, tl_base, tl_n, 11 + 2 * t);
      }
      tc_fence_before();
    }
    if (gtid == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) { __syncwarp(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---------------------------------------------------------------- freeze-mode head gradients
// forward_freezed (ibl_nerf.py:88-152, iterations >= N_iter_freeze of the shipped schedule): only the albedo /
// irradiance feature layers and heads (and the roughness head unless freeze_roughness) receive gradients, so the
// dgrad chain collapses to the albedo|irradiance feature tile: dY_af = relu'(af) * (g_albedo W_alb | g_irr W_irr)
// on CUDA cores, written with the G tile for the three wgrad jobs that remain.  One tile per 128-thread CTA pass.
struct FreezeParams {
  const uint8_t* packed; const uint8_t* saved; const float* g_out; uint8_t* dy; float* flat_grad;
  long long P, n_tiles;
  int train_roughness;
};
__global__ void __launch_bounds__(128, 2) mlp_dgrad_freeze_kernel(FreezeParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                                        // [4 K-blocks] dY_af tile (operand layout)
  float* T_AF = reinterpret_cast<float*>(smem + ACT_BYTES);   // albedo [3][128] | irradiance [128]
  const float* cst = reinterpret_cast<const float*>(prm.packed + PACKED_CONST_OFF);
  for (int i = threadIdx.x; i < 128; i += 128)
    for (int j = 0; j < 4; ++j) T_AF[4 * (i + 0) + j] = __ldg(cst + C_AF + 4 * i + j);
  __syncthreads();
  const int row = threadIdx.x, lane = threadIdx.x & 31;
  const FlatOff fo = flat_offsets();
  uint32_t off[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) off[q] = swz_offset(row, q);
  for (long long tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
    const long long p = tile * TILE_M + row;
    const bool valid = p < prm.P;
    uint8_t* dy = prm.dy + (size_t)tile * DY_BYTES;
    const uint32_t* masks = reinterpret_cast<const uint32_t*>(prm.saved + (size_t)tile * SV_BYTES + (size_t)SV_MASK * KB_BYTES) + row;
    float g[18];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float2 v = valid ? __ldg(reinterpret_cast<const float2*>(prm.g_out + p * 18) + j) : make_float2(0.f, 0.f);
      g[2 * j] = v.x; g[2 * j + 1] = v.y;
    }
    {   // head-bias gradients (albedo, irradiance, roughness)
      float s1 = warp_sum(g[1]), s2 = warp_sum(g[2]), s3 = warp_sum(g[3]), s4 = warp_sum(g[4]), s5 = warp_sum(g[5]);
      if (lane == 0) {
        atomicAdd(prm.flat_grad + fo.b[12], s1); atomicAdd(prm.flat_grad + fo.b[12] + 1, s2); atomicAdd(prm.flat_grad + fo.b[12] + 2, s3);
        atomicAdd(prm.flat_grad + fo.b[15], s5);
        if (prm.train_roughness) atomicAdd(prm.flat_grad + fo.b[13], s4);
      }
    }
    {   // G tile (channels 1..5 are the only ones the remaining jobs read; write all 18 like the full kernel)
      float e[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) e[j] = j < 18 ? g[j] : 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        store_tile4(dy + (size_t)DY_G * KB_BYTES, 0, swz_offset(row, ch),
                    make_uint4(pack_bf16x2(e[8 * ch], e[8 * ch + 1]), pack_bf16x2(e[8 * ch + 2], e[8 * ch + 3]),
                               pack_bf16x2(e[8 * ch + 4], e[8 * ch + 5]), pack_bf16x2(e[8 * ch + 6], e[8 * ch + 7])));
    }
    uint32_t mw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mw[j] = __ldg(masks + 8 * 1024 + j * 128);
    for (int cc = 0; cc < 8; ++cc) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
      if (cc < 4) {
#pragma unroll
        for (int q = 0; q < 3; ++q) axpy32(v, g[1 + q], T_AF + q * 128 + cc * 32);
      } else {
        axpy32(v, g[5], T_AF + 384 + (cc - 4) * 32);
      }
      const uint32_t m = mw[cc];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        store_tile4(act, cc >> 1, off[(cc & 1) * 4 + q],
                    make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                               pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(act);
    uint4* dst = reinterpret_cast<uint4*>(dy + (size_t)DY_AF * KB_BYTES);
#pragma unroll 8
    for (int i = threadIdx.x; i < ACT_BYTES / 16; i += 128) __stcs(dst + i, src[i]);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- wgrad kernel
// D[m][n] += sum_pt A[pt][m] * B[pt][n].  A = 2 adjacent 16 KB blocks of a record (128 "m" columns), B = nb
// adjacent blocks (64*nb "n" columns), both read as MN-major SWIZZLE_128B operands straight from the stash /
// dY records.  Optional extra N=64 MMA against a constant ones tile gives the column sums of A (bias
// gradient) in accumulator column 256.
//
// ONE persistent launch per network covers all 21 GEMMs ("jobs").  The (job, tile) space is linearised with a
// per-job cost (bytes streamed per tile) and cut into equal contiguous ranges, one per CTA PAIR; the two
// CTAs of a pair take the two 128-row halves of the same tiles at the same time (the second read of the
// shared B operand hits L2) or, for single-half jobs, split the tile range.  A CTA therefore flushes its
// TMEM accumulator (red.global.add.f32) only once per (job, range) segment: ~200 flushes per network
// instead of one per CTA per job, and no launch gaps between the jobs.
struct WOut { int off; int m_lo, m_hi, n_lo, n_hi, stride_m, stride_n; };
struct WHead { int off, ch_lo, ch_hi; };      // rows [ch_lo, ch_hi) of the G-head accumulator -> flat[off + (ch - ch_lo) * 256 + feature]
struct WJob {
  int a_sv, a_blk;       // A operand: record (1 = forward stash, 0 = dY record) and first block of m-half 0
  int b_sv, b_blk;
  int m_halves;          // 1 or 2: CTA class c handles blocks a_blk + 2c, a_blk + 2c + 1
  int nb;                // B blocks (1..4)
  int n_total;           // N of the main MMA (multiple of 16, <= 256)
  int with_ones;
  int a_noswz;           // A blocks are in the slice-interleaved no-swizzle layout (SV_AF, SV_ADDF 0..3)
  int xb, x_sv, x_blk;   // one extra 64-column B block sharing A (accumulator columns 320..383): the encoding part of a
                         // layer whose input is [activation | encoding]
  int gh;                // G head: a second GEMM sharing B, D2[g channel][B column] = G^T B (M = 64; class c takes B
                         // columns 128c..128c+127, accumulator columns 320..447): the small heads that read this job's B
  int cost;              // relative streaming cost of one tile for a CTA pair
  int n_out, n_head;
  WOut out[4];
  WHead head[2];
};
constexpr int WG_MAX_JOBS = 17;
struct WgradParams {
  const uint8_t* sv; const uint8_t* dy; float* flat;
  long long n_tiles;
  int n_jobs;
  WJob job[WG_MAX_JOBS];
};
static_assert(sizeof(WgradParams) <= 4096, "kernel parameter space");

// Operand ring: a stage holds WG_STAGE_PTS points of every block of the job (the rows of a 16 KB block are contiguous,
// so a row range is one bulk copy per block).  Small stages keep more bytes in flight per SM (Little's law: the
// kernel is a pure HBM/L2 stream, ~0.09 FLOP/B short of the tensor roofline) than whole-tile double buffering.
// The ring is re-cut per job: a job with k blocks per stage gets min(16, 52 / k) stages of k * 4 KB (3 blocks: 16
// stages, 6: 8, 7: 7), after the stages of the previous geometry have drained.
constexpr int WG_STAGE_PTS = 32;
constexpr int WG_SUB = TILE_M / WG_STAGE_PTS;              // stages per tile
constexpr int WG_BLK_BYTES = WG_STAGE_PTS * 128;           // bytes of one block's row range
constexpr int WG_RING_BLOCKS = 52;
constexpr int WG_MAX_STAGES = 16;
constexpr int WG_SMEM_ONES = WG_RING_BLOCKS * WG_BLK_BYTES;  // [WG_STAGE_PTS rows][64] ones tile
constexpr int WG_SMEM_BAR = WG_SMEM_ONES + WG_BLK_BYTES;
constexpr int WG_SMEM_REQUEST = WG_SMEM_BAR + 512 + 1024;
constexpr int WG_COL_ONES = 256, WG_COL_EXTRA = 320;       // accumulator columns of the bias sums / the extra GEMM
__host__ __device__ inline int wg_stage_blocks(const WJob& jb) { return 2 + jb.nb + jb.xb + jb.gh; }
__device__ __forceinline__ int wg_stages(int nblk) { const int n = WG_RING_BLOCKS / nblk; return n > WG_MAX_STAGES ? WG_MAX_STAGES : n; }
// warps 0..3 producers (stage i of the running stage counter belongs to producer i % 4: one thread issuing the
// 3-6 bulk copies of every 32-point slice was the pacing item -- ncu: the lone producer busy 85 % of the time, the
// MMA thread waiting for data 55 %), warp 4 MMA, warps 5-8 epilogue
constexpr int WG_PRODUCERS = 4;
constexpr int WG_THREADS = (WG_PRODUCERS + 5) * 32;

// tile range [t0, t1) and m-half of job j for CTA (pair, r); pairs = number of CTA pairs in the grid
struct WSeg { long long t0, t1; int cls; };
__device__ __forceinline__ WSeg wgrad_segment(const WgradParams& prm, int j, long long job_start, long long w_total, int pair,
                                               int pairs, int r) {
  const long long lo = w_total * pair / pairs, hi = w_total * (pair + 1) / pairs;
  const long long c = prm.job[j].cost;
  auto f = [&](long long x) {
    long long t = (x - job_start) / c;
    if (x <= job_start) t = 0;
    return t > prm.n_tiles ? prm.n_tiles : t;
  };
  WSeg s;
  s.t0 = f(lo); s.t1 = f(hi); s.cls = r;
  if (prm.job[j].m_halves == 1) {
    const long long mid = s.t0 + (s.t1 - s.t0 + 1) / 2;
    if (r == 0) s.t1 = mid; else s.t0 = mid;
    s.cls = 0;
  }
  return s;
}

// MMA issue loop of one (job, tile range) segment: n stages of 32 points = 2 K-steps each.
// EXTRA: 0 none, 1 extra B block (N = 64, shares A), 2 G head (A = G block with M = 64, B = this CTA's half of B)
struct WgMma {
  uint32_t smem, st_bytes, e_off, h_off, ones_addr, tmem, idesc;
  int nst;
  uint64_t* full; uint64_t* empty;
};
template <bool ONES, bool NOSWZ, int EXTRA>
__device__ __forceinline__ void wg_mma_segment(const WgMma& mm, long long n, int& slot, uint32_t& par) {
  constexpr uint32_t idesc1 = make_idesc_bf16(128, 64, 1, 1);   // whole 64-wide swizzle atom (ones tile: only column 0 is non-zero)
  constexpr uint32_t idesc_head = make_idesc_bf16(64, 128, 1, 1);
  for (long long it = 0; it < n; ++it) {
    const int stage = slot;
    const uint32_t ph = (par >> stage) & 1u;
    par ^= 1u << stage;
    if (++slot == mm.nst) slot = 0;
    mbar_wait(&mm.full[stage], ph);
    tc_fence_after();
    const uint32_t a_addr = mm.smem + stage * mm.st_bytes;
    const uint32_t b_addr = a_addr + 2 * WG_BLK_BYTES;
#pragma unroll
    for (int ks = 0; ks < WG_STAGE_PTS / 16; ++ks) {       // 16 points per MMA
      const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
      const uint64_t da = NOSWZ ? make_desc_mnmajor_noswz(a_addr + ks * 256, 128, 512)
                                : make_desc_mnmajor_sw128(a_addr + ks * 2048, WG_BLK_BYTES);
      umma_bf16(mm.tmem, da, make_desc_mnmajor_sw128(b_addr + ks * 2048, WG_BLK_BYTES), mm.idesc, acc);
      if (ONES) umma_bf16(mm.tmem + WG_COL_ONES, da, make_desc_mnmajor_sw128(mm.ones_addr + ks * 2048, WG_BLK_BYTES), idesc1, acc);
      if (EXTRA == 1)
        umma_bf16(mm.tmem + WG_COL_EXTRA, da, make_desc_mnmajor_sw128(a_addr + mm.e_off + ks * 2048, WG_BLK_BYTES), idesc1, acc);
      if (EXTRA == 2)
        umma_bf16(mm.tmem + WG_COL_EXTRA, make_desc_mnmajor_sw128(a_addr + mm.e_off + ks * 2048, WG_BLK_BYTES),
                  make_desc_mnmajor_sw128(a_addr + mm.h_off + ks * 2048, WG_BLK_BYTES), idesc_head, acc);
    }
    umma_commit(&mm.empty[stage]);
  }
}

__global__ void __launch_bounds__(WG_THREADS, 1) mlp_wgrad_kernel(const __grid_constant__ WgradParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_SMEM_BAR);
  uint64_t* full = bars;                       // [WG_MAX_STAGES]
  uint64_t* empty = bars + WG_MAX_STAGES;      // [WG_MAX_STAGES]
  uint64_t* done = bars + 2 * WG_MAX_STAGES;   // MMA -> epilogue: the segment's accumulator is complete
  uint64_t* acc_free = bars + 2 * WG_MAX_STAGES + 1;   // epilogue -> MMA: the accumulator has been drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * WG_MAX_STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x >> 1, r = blockIdx.x & 1, pairs = gridDim.x >> 1;
  long long w_total = 0;
  for (int j = 0; j < prm.n_jobs; ++j) w_total += prm.n_tiles * prm.job[j].cost;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_MAX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    mbar_init(acc_free, 4);
    fence_barrier_init();
  }
  if (warp == WG_PRODUCERS) tmem_alloc(tmem_ptr, 512);
  // ones tile: column 0 of every row = 1.0 (bf16), everything else 0
  for (int e = threadIdx.x; e < WG_STAGE_PTS * 8; e += blockDim.x) {
    int rr = e >> 3, c16 = e & 7;
    *reinterpret_cast<uint4*>(smem + WG_SMEM_ONES + swz_offset(rr, c16)) = make_uint4(c16 == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Every role walks the same (job, tile, slice) sequence; `par` holds the phase parity of each ring slot's next use.
  if (warp < WG_PRODUCERS) {
    if (elect_one()) {
      long long i = 0, job_start = 0;           // i: running stage counter of this CTA (producer round robin)
      uint32_t par = 0;
      int nblk = 0, nst = 0, slot = 0;
      for (int j = 0; j < prm.n_jobs; ++j) {
        const WJob& jb = prm.job[j];
        const WSeg sg = wgrad_segment(prm, j, job_start, w_total, pair, pairs, r);
        job_start += prm.n_tiles * jb.cost;
        if (sg.t1 <= sg.t0) continue;
        if (wg_stage_blocks(jb) != nblk) {       // new ring geometry: every stage of the old one has been consumed
          for (int s2 = 0; s2 < nst; ++s2) mbar_wait(&empty[s2], ((par >> s2) & 1u) ^ 1u);
          nblk = wg_stage_blocks(jb); nst = wg_stages(nblk); slot = 0;
        }
        const uint8_t* a_base = (jb.a_sv ? prm.sv : prm.dy) + (size_t)(jb.a_blk + 2 * sg.cls) * KB_BYTES;
        const uint8_t* b_base = (jb.b_sv ? prm.sv : prm.dy) + (size_t)jb.b_blk * KB_BYTES;
        const uint8_t* x_base = (jb.x_sv ? prm.sv : prm.dy) + (size_t)jb.x_blk * KB_BYTES;
        const uint8_t* g_base = prm.dy + (size_t)DY_G * KB_BYTES;
        const long long a_stride = jb.a_sv ? SV_BYTES : DY_BYTES, b_stride = jb.b_sv ? SV_BYTES : DY_BYTES;
        const long long x_stride = jb.x_sv ? SV_BYTES : DY_BYTES;
        const uint32_t st_bytes = (uint32_t)nblk * WG_BLK_BYTES;
        const int nb = jb.nb;                      // job fields in registers: the asm memory clobbers would re-read them
        const bool xb = jb.xb, gh = jb.gh;
        for (long long t = sg.t0; t < sg.t1; ++t) {
          const uint8_t* a_src = a_base + (size_t)t * a_stride;
          const uint8_t* b_src = b_base + (size_t)t * b_stride;
          for (int q = 0; q < WG_SUB; ++q, ++i) {
            const int stage = slot;
            const uint32_t ph = (par >> stage) & 1u;
            par ^= 1u << stage;
            if (++slot == nst) slot = 0;
            if ((int)(i % WG_PRODUCERS) != warp) continue;
            mbar_wait(&empty[stage], ph ^ 1);
            mbar_arrive_expect_tx(&full[stage], st_bytes);
            uint8_t* dst = smem + (size_t)stage * st_bytes;
            for (int b = 0; b < 2; ++b)
              bulk_g2s(dst + b * WG_BLK_BYTES, a_src + (size_t)b * KB_BYTES + q * WG_BLK_BYTES, WG_BLK_BYTES, &full[stage]);
            for (int b = 0; b < nb; ++b)
              bulk_g2s(dst + (2 + b) * WG_BLK_BYTES, b_src + (size_t)b * KB_BYTES + q * WG_BLK_BYTES, WG_BLK_BYTES, &full[stage]);
            if (xb)
              bulk_g2s(dst + (2 + nb) * WG_BLK_BYTES, x_base + (size_t)t * x_stride + q * WG_BLK_BYTES, WG_BLK_BYTES, &full[stage]);
            if (gh)
              bulk_g2s(dst + (2 + nb) * WG_BLK_BYTES, g_base + (size_t)t * DY_BYTES + q * WG_BLK_BYTES, WG_BLK_BYTES, &full[stage]);
          }
        }
      }
    }
  } else if (warp == WG_PRODUCERS) {
    if (elect_one()) {
      const uint32_t ones_addr = smem_u32(smem + WG_SMEM_ONES);
      long long job_start = 0;
      uint32_t nseg = 0, par = 0;
      int nblk = 0, nst = 0, slot = 0;
      for (int j = 0; j < prm.n_jobs; ++j) {
        const WJob& jb = prm.job[j];
        const WSeg sg = wgrad_segment(prm, j, job_start, w_total, pair, pairs, r);
        job_start += prm.n_tiles * jb.cost;
        if (sg.t1 <= sg.t0) continue;
        if (wg_stage_blocks(jb) != nblk) { nblk = wg_stage_blocks(jb); nst = wg_stages(nblk); slot = 0; }
        if (nseg > 0) { mbar_wait(acc_free, (nseg - 1) & 1); tc_fence_after(); }
        ++nseg;
        WgMma mm;
        mm.smem = smem_u32(smem); mm.st_bytes = (uint32_t)nblk * WG_BLK_BYTES; mm.e_off = (uint32_t)(2 + jb.nb) * WG_BLK_BYTES;
        mm.h_off = (uint32_t)(2 + 2 * sg.cls) * WG_BLK_BYTES; mm.ones_addr = ones_addr; mm.tmem = tmem_base;
        mm.idesc = make_idesc_bf16(128, (uint32_t)jb.n_total, 1, 1); mm.nst = nst; mm.full = full; mm.empty = empty;
        const long long n = (sg.t1 - sg.t0) * WG_SUB;
        // the issuing thread paces the 12 KB-stage jobs (~560 clk per stage): one loop per job shape, nothing predicated
        const int shape = (jb.a_noswz ? 1 : 0) | (jb.with_ones ? 2 : 0) | (jb.xb ? 4 : 0) | (jb.gh ? 8 : 0);
        switch (shape) {
          case 0: wg_mma_segment<false, false, 0>(mm, n, slot, par); break;
          case 1: wg_mma_segment<false, true, 0>(mm, n, slot, par); break;
          case 2: wg_mma_segment<true, false, 0>(mm, n, slot, par); break;
          case 6: wg_mma_segment<true, false, 1>(mm, n, slot, par); break;
          case 10: wg_mma_segment<true, false, 2>(mm, n, slot, par); break;
          default: __trap();
        }
        umma_commit(done);
      }
    }
  } else {
    const int quarter = warp & 3;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    long long job_start = 0;
    uint32_t nseg = 0;
    for (int j = 0; j < prm.n_jobs; ++j) {
      const WJob& jb = prm.job[j];
      const WSeg sg = wgrad_segment(prm, j, job_start, w_total, pair, pairs, r);
      job_start += prm.n_tiles * jb.cost;
      if (sg.t1 <= sg.t0) continue;
      mbar_wait(done, nseg & 1);
      ++nseg;
      tc_fence_after();
      const int m = sg.cls * 128 + quarter * 32 + lane;
      const int nmain = (jb.n_total + 31) & ~31;
      const int ncols = jb.xb ? WG_COL_EXTRA + 64 : jb.with_ones ? WG_COL_ONES + 32 : nmain;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        if (c0 >= nmain && c0 < WG_COL_ONES) continue;
        if (c0 >= WG_COL_ONES + 32 && c0 < WG_COL_EXTRA) continue;
        if (c0 == WG_COL_ONES && !jb.with_ones) continue;
        uint32_t raw[32];
        tmem_ld32(t_lane + c0, raw);
        tmem_wait_ld();
        for (int o = 0; o < jb.n_out; ++o) {
          const WOut sp = jb.out[o];
          if (m < sp.m_lo || m >= sp.m_hi) continue;
          float* dst = prm.flat + sp.off + (long long)(m - sp.m_lo) * sp.stride_m;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int n = c0 + jj;
            if (n >= sp.n_lo && n < sp.n_hi) atomicAdd(dst + (long long)(n - sp.n_lo) * sp.stride_n, __uint_as_float(raw[jj]));
          }
        }
      }
      if (jb.gh && quarter == 0) {     // G head: accumulator row = g channel (rows 0..15 live in lanes 0..15), column = B column
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(t_lane + WG_COL_EXTRA + c0, raw);
          tmem_wait_ld();
          for (int o = 0; o < jb.n_head; ++o) {
            const WHead hd = jb.head[o];
            if (lane < hd.ch_lo || lane >= hd.ch_hi) continue;
            float* dst = prm.flat + hd.off + (lane - hd.ch_lo) * 256 + sg.cls * 128 + c0;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) atomicAdd(dst + jj, __uint_as_float(raw[jj]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_PRODUCERS) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// the 17 jobs of one network's weight gradient (freeze = 0; 21 GEMMs, four of them riding on a job that streams
// their operand anyway), or the 2 that remain in the freeze modes
// (1: freeze_radiance, 2: freeze_radiance + freeze_roughness)
static WgradParams make_wgrad_jobs(int freeze) {
  WgradParams P;
  memset(&P, 0, sizeof(P));
  const FlatOff fo = flat_offsets();
  auto job = [&](int a_sv, int a_blk, int m_halves, int b_sv, int b_blk, int nb, int n_total, int with_ones) -> WJob& {
    WJob& j = P.job[P.n_jobs++];
    j.a_sv = a_sv; j.a_blk = a_blk; j.m_halves = m_halves; j.b_sv = b_sv; j.b_blk = b_blk; j.nb = nb; j.n_total = n_total;
    j.with_ones = with_ones; j.n_out = 0; j.a_noswz = 0;
    j.cost = (2 + nb) * (m_halves == 2 ? 2 : 1);
    return j;
  };
  // the encoding block that completes this job's layer input rides along as a fifth B block
  auto add_extra = [&](WJob& j, int x_sv, int x_blk) { j.xb = 1; j.x_sv = x_sv; j.x_blk = x_blk; j.cost = wg_stage_blocks(j) * 2; };
  // a small head that reads this job's B operand: rows [ch_lo, ch_hi) of G^T B
  auto add_head = [&](WJob& j, int off, int ch_lo, int ch_hi) {
    j.gh = 1; j.cost = wg_stage_blocks(j) * 2;
    WHead& h = j.head[j.n_head++];
    h.off = off; h.ch_lo = ch_lo; h.ch_hi = ch_hi;
  };
  auto add_out = [&](WJob& j, int off, int m_lo, int m_hi, int n_lo, int n_hi, int stride_m, int stride_n) {
    WOut& o = j.out[j.n_out++];
    o.off = off; o.m_lo = m_lo; o.m_hi = m_hi; o.n_lo = n_lo; o.n_hi = n_hi; o.stride_m = stride_m; o.stride_n = stride_n;
  };
  constexpr int DYR = 0, SVR = 1;
  if (freeze) {
    {   // albedo / irradiance feature linears: dY_af x h7
      WJob& p = job(DYR, DY_AF, 2, SVR, SV_H(7), 4, 256, 1);
      add_out(p, fo.w[11], 0, 128, 0, 256, 256, 1);
      add_out(p, fo.w[14], 128, 256, 0, 256, 256, 1);
      add_out(p, fo.b[11], 0, 128, 256, 257, 1, 0);
      add_out(p, fo.b[14], 128, 256, 256, 257, 1, 0);
      if (freeze == 1) add_head(p, fo.w[13], 4, 5);   // roughness head from h7
    }
    {   // albedo / irradiance heads from AF
      WJob& p = job(SVR, SV_AF, 2, DYR, DY_G, 1, 64, 0);
      p.a_noswz = 1;
      add_out(p, fo.w[12], 0, 128, 1, 4, 1, 128);
      add_out(p, fo.w[15], 128, 256, 5, 6, 1, 0);
    }
    return P;
  }
  // trunk layers: dW_l = dY_l^T X_l (+ bias through the ones column)
  for (int l = 0; l < 8; ++l) {
    const int ld = l == 0 ? 63 : (l == 5 ? 319 : 256);
    if (l == 0) {   // positional encoding (63 valid of 64 columns)
      WJob& p = job(DYR, DY_H(l), 2, SVR, SV_PE, 1, 64, 1);
      add_out(p, fo.w[l], 0, 256, 0, 63, ld, 1);
      add_out(p, fo.b[l], 0, 256, 256, 257, 1, 0);
    } else {
      WJob& p = job(DYR, DY_H(l), 2, SVR, SV_H(l - 1), 4, 256, 1);
      add_out(p, fo.w[l] + (l == 5 ? 63 : 0), 0, 256, 0, 256, ld, 1);
      add_out(p, fo.b[l], 0, 256, 256, 257, 1, 0);
      if (l == 5) {   // skip layer: input = [encoding | h4]
        add_extra(p, SVR, SV_PE);
        add_out(p, fo.w[l], 0, 256, WG_COL_EXTRA, WG_COL_EXTRA + 63, ld, 1);
      }
    }
  }
  {   // feature_linear: dY_feat x h7
    WJob& p = job(DYR, DY_FEAT, 2, SVR, SV_H(7), 4, 256, 1);
    add_out(p, fo.w[9], 0, 256, 0, 256, 256, 1);
    add_out(p, fo.b[9], 0, 256, 256, 257, 1, 0);
    add_head(p, fo.w[10], 0, 1);      // sigma / roughness heads from h7
    add_head(p, fo.w[13], 4, 5);
  }
  {   // albedo / irradiance feature linears: dY_af x h7 (rows 0..127 albedo_f, 128..255 irradiance_f)
    WJob& p = job(DYR, DY_AF, 2, SVR, SV_H(7), 4, 256, 1);
    add_out(p, fo.w[11], 0, 128, 0, 256, 256, 1);
    add_out(p, fo.w[14], 128, 256, 0, 256, 256, 1);
    add_out(p, fo.b[11], 0, 128, 256, 257, 1, 0);
    add_out(p, fo.b[14], 128, 256, 256, 257, 1, 0);
  }
  {   // views_linears.0: dY_view x [feature | view encoding]
    WJob& p = job(DYR, DY_VIEW, 2, SVR, SV_FEAT, 4, 256, 1);
    add_out(p, fo.w[8], 0, 256, 0, 256, 283, 1);
    add_out(p, fo.b[8], 0, 256, 256, 257, 1, 0);
    add_extra(p, SVR, SV_DE);                                  // columns 27.. of the encoding tile are unused
    add_out(p, fo.w[8] + 256, 0, 256, WG_COL_EXTRA, WG_COL_EXTRA + 27, 283, 1);
  }
  {   // coarse-radiance feature linears: dY_addf x hv
    WJob& p = job(DYR, DY_ADDF01, 2, SVR, SV_HV, 4, 256, 1);
    add_out(p, fo.w[17], 0, 128, 0, 256, 256, 1);
    add_out(p, fo.w[18], 128, 256, 0, 256, 256, 1);
    add_out(p, fo.b[17], 0, 128, 256, 257, 1, 0);
    add_out(p, fo.b[18], 128, 256, 256, 257, 1, 0);
    add_head(p, fo.w[16], 6, 9);      // radiance head from hv
    WJob& q = job(DYR, DY_ADDF2, 1, SVR, SV_HV, 4, 256, 1);
    add_out(q, fo.w[19], 0, 128, 0, 256, 256, 1);
    add_out(q, fo.b[19], 0, 128, 256, 257, 1, 0);
  }
  // small heads: D[feature col][g channel] = X^T G
  {   // albedo (cols 0..127 x channels 1..3) / irradiance (cols 128..255 x channel 5) from AF
    WJob& p = job(SVR, SV_AF, 2, DYR, DY_G, 1, 64, 0);
    p.a_noswz = 1;
    add_out(p, fo.w[12], 0, 128, 1, 4, 1, 128);
    add_out(p, fo.w[15], 128, 256, 5, 6, 1, 0);
  }
  for (int k = 0; k < 3; ++k) {   // coarse radiance heads from ADDF block pair k
    WJob& p = job(SVR, SV_ADDF + 2 * k, 1, DYR, DY_G, 1, 64, 0);
    p.a_noswz = k < 2;      // feature 2 leaves the forward kernel through the activation tile (swizzled image)
    add_out(p, fo.w[20 + k], 0, 128, 9 + 3 * k, 12 + 3 * k, 1, 128);
  }
  return P;
}

}  // namespace mlp
}  // namespace ibln

using namespace ibln;
using namespace ibln::mlp;

extern "C" int64_t ibln_mlp_bwd_workspace_bytes(int64_t n_pts) { return ((n_pts + TILE_M - 1) / TILE_M) * DY_BYTES; }

extern "C" int ibln_mlp_bwd(const void* packed, const void* saved, const float* g_out, int64_t n_pts, float* flat_grad,
                            void* workspace, int freeze_mode, int device, void* stream_) {
  if (n_pts == 0) return 0;
  if (!packed || !saved || !g_out || !flat_grad || !workspace || n_pts < 0 || freeze_mode < 0 || freeze_mode > 2) return IBLN_EINVAL;
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long n_tiles = (n_pts + TILE_M - 1) / TILE_M;
  const int sms = num_sms(device);
  if (freeze_mode == 0) {
    // ---- dgrad chain
    DgradParams dp;
    dp.packed = (const uint8_t*)packed; dp.saved = (const uint8_t*)saved; dp.g_out = g_out; dp.dy = (uint8_t*)workspace;
    dp.flat_grad = flat_grad; dp.P = n_pts; dp.n_tiles = n_tiles; dp.dbg = IBLN_DBG_FLAGS; dp.tl = (unsigned long long*)IBLN_DBG_TIMELINE;
    IBLN_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_REQUEST));
    long long grid = (long long)(sms & ~1);                      // CTA pairs
    if (((n_tiles + 1) & ~1LL) < grid) grid = (n_tiles + 1) & ~1LL;
    { int rc = make_chunk_stream_map(&dp.wmap, (const uint8_t*)packed + PACKED_BWD_OFF, N_CHUNKS_BWD); if (rc != 0) return rc; }
    if (!(IBLN_DBG_FLAGS & 16)) mlp_dgrad_kernel<<<(unsigned)grid, N_THREADS, SMEM_REQUEST, stream>>>(dp);
    IBLN_CUDA(cudaGetLastError());
    if (IBLN_DBG_FLAGS & 32) return 0;
  } else {
    FreezeParams fp;
    fp.packed = (const uint8_t*)packed; fp.saved = (const uint8_t*)saved; fp.g_out = g_out; fp.dy = (uint8_t*)workspace;
    fp.flat_grad = flat_grad; fp.P = n_pts; fp.n_tiles = n_tiles; fp.train_roughness = freeze_mode == 1;
    const int smem = ACT_BYTES + 2048 + 1024;
    IBLN_CUDA(cudaFuncSetAttribute(mlp_dgrad_freeze_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long grid = n_tiles < 2LL * sms ? n_tiles : 2LL * sms;
    mlp_dgrad_freeze_kernel<<<(unsigned)grid, 128, smem, stream>>>(fp);
    IBLN_CUDA(cudaGetLastError());
  }
  // ---- all weight-gradient GEMMs in one persistent launch
  static const WgradParams job_tables[3] = {make_wgrad_jobs(0), make_wgrad_jobs(1), make_wgrad_jobs(2)};
  WgradParams wp = job_tables[freeze_mode];
  wp.sv = (const uint8_t*)saved; wp.dy = (const uint8_t*)workspace; wp.flat = flat_grad; wp.n_tiles = n_tiles;
  IBLN_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_REQUEST));
  long long pairs = sms / 2;
  if (pairs > n_tiles * 4) pairs = n_tiles * 4;      // tiny inputs: do not launch CTAs that would only idle
  if (pairs < 1) pairs = 1;
  mlp_wgrad_kernel<<<(unsigned)(2 * pairs), WG_THREADS, WG_SMEM_REQUEST, stream>>>(wp);
  IBLN_RETURN_LAST();
}

// MN-major operand self-test: D[128,N] = X^T Y with X [128 pts][128], Y [128 pts][N] staged as operand tiles.
namespace ibln { namespace mlp {
__global__ void __launch_bounds__(128, 1)
umma_mn_selftest_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D, int N) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // 2 blocks
  uint8_t* sB = smem + 2 * KB_BYTES;        // N/64 blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * KB_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  auto stage = [&](uint8_t* dst, const float* src, int cols) {
    for (int e = threadIdx.x; e < 128 * (cols / 8); e += blockDim.x) {
      int row = e / (cols / 8), cg = e % (cols / 8);
      const float* s = src + (size_t)row * cols + cg * 8;
      *reinterpret_cast<uint4*>(dst + (size_t)(cg / 8) * KB_BYTES + swz_offset(row, cg % 8)) =
          make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
    }
  };
  stage(sA, X, 128);
  stage(sB, Y, N);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)N, 1, 1);
    for (int ks = 0; ks < 8; ++ks)
      umma_bf16(tmem_base, make_desc_mnmajor_sw128(smem_u32(sA) + ks * 2048, KB_BYTES),
                make_desc_mnmajor_sw128(smem_u32(sB) + ks * 2048, KB_BYTES), idesc, ks > 0 ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int row = (warp & 3) * 32 + lane;
  for (int cc = 0; cc < (N + 31) / 32; ++cc) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) if (cc * 32 + j < N) D[(size_t)row * N + cc * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}
}}  // namespace ibln::mlp

extern "C" int ibln_umma_mn_selftest(const float* x, const float* y, float* d, int n, int device, void* stream) {
  if (!x || !y || !d || n < 64 || n > 256 || n % 64 != 0) return IBLN_EINVAL;
  DeviceGuard g(device);
  int smem = 6 * KB_BYTES + 64 + 1024;
  IBLN_CUDA(cudaFuncSetAttribute(umma_mn_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_mn_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(x, y, d, n);
  IBLN_RETURN_LAST();
}

// Adam over the flat parameter buffer of n_nets networks + the bf16 re-pack of all of them: 2 launches per step
// instead of 1 + 3 per network (ibln_adam_step + ibln_mlp_pack_weights).
extern "C" int ibln_adam_step_pack(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int n_nets, float lr,
                                   float beta1, float beta2, float eps, int step, float grad_scale,
                                   void* const* packed_host, int device, void* stream) {
  if (n_nets < 1 || n_nets > 4 || !packed_host) return IBLN_EINVAL;
  int rc = ibln_adam_step(param, grad, exp_avg, exp_avg_sq, (int64_t)n_nets * FLAT_TOTAL, lr, beta1, beta2, eps, step, grad_scale,
                          device, stream);
  if (rc != 0) return rc;
  DeviceGuard guard(device);
  PackFlat pf;
  for (int i = 0; i < 4; ++i) {
    pf.flat[i] = i < n_nets ? param + (size_t)i * FLAT_TOTAL : nullptr;
    pf.packed[i] = i < n_nets ? (uint8_t*)packed_host[i] : nullptr;
    if (i < n_nets && (!pf.packed[i] || (reinterpret_cast<uintptr_t>(pf.packed[i]) & 15))) return IBLN_EINVAL;
  }
  return launch_pack_flat(pf, n_nets, (cudaStream_t)stream);
}

// The data-parallel form of the above: the gradient all-reduce is fused into the Adam kernel (ibln_adam_allreduce_step:
// NVSwitch multimem.ld_reduce on the symmetric gradient buffer, or P2P loads), then the same single re-pack launch.
extern "C" int ibln_adam_allreduce_step_pack(float* param, const float* grad_multicast, const float* const* peer_grads_host,
                                             int world, float* exp_avg, float* exp_avg_sq, int n_nets, float lr, float beta1,
                                             float beta2, float eps, int step, float grad_scale, void* const* packed_host,
                                             int device, void* stream) {
  if (n_nets < 1 || n_nets > 4 || !packed_host) return IBLN_EINVAL;
  int rc = ibln_adam_allreduce_step(param, grad_multicast, peer_grads_host, world, exp_avg, exp_avg_sq, (int64_t)n_nets * FLAT_TOTAL,
                                    lr, beta1, beta2, eps, step, grad_scale, device, stream);
  if (rc != 0) return rc;
  DeviceGuard guard(device);
  PackFlat pf;
  for (int i = 0; i < 4; ++i) {
    pf.flat[i] = i < n_nets ? param + (size_t)i * FLAT_TOTAL : nullptr;
    pf.packed[i] = i < n_nets ? (uint8_t*)packed_host[i] : nullptr;
    if (i < n_nets && (!pf.packed[i] || (reinterpret_cast<uintptr_t>(pf.packed[i]) & 15))) return IBLN_EINVAL;
  }
  return launch_pack_flat(pf, n_nets, (cudaStream_t)stream);
}

// Bandwidth probes: compiled only into the tuning build (-DIBLN_DIAGNOSTICS, python -m ibl_nerf_b200.build --diag);
// the product library contains none of this.
#ifdef IBLN_DIAGNOSTICS
#include "../../include/iblnerf_b200_diag.h"
#include "tc_common.cuh"

namespace ibln {
using namespace tc;

// mode 0: one 64 KB bulk store in flight per CTA; mode 1: up to 4 in flight (16 KB each from 4 buffers);
// mode 2: coalesced st.global.v4 from registers; mode 3: bulk store with evict-first style L2 hint
__global__ void __launch_bounds__(256, 1) store_probe_kernel(uint8_t* __restrict__ out, long long bytes_per_cta, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, i, i, i);
  fence_proxy_async();
  __syncthreads();
  uint8_t* dst = out + (size_t)blockIdx.x * bytes_per_cta;
  const long long n64 = bytes_per_cta / 65536;
  if (mode == 0) {
    if (threadIdx.x == 0) {
      for (long long i = 0; i < n64; ++i) { bulk_s2g(dst + i * 65536, smem, 65536); bulk_commit(); bulk_wait_read0(); }
      bulk_wait0();
    }
  } else if (mode == 1) {
    if (threadIdx.x == 0) {
      for (long long i = 0; i < n64 * 4; ++i) {
        bulk_s2g(dst + i * 16384, smem + (i & 3) * 16384, 16384);
        bulk_commit();
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      }
      bulk_wait0();
    }
  } else {
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    for (long long i = 0; i < n64; ++i) {
      uint4* p = reinterpret_cast<uint4*>(dst + i * 65536);
#pragma unroll 4
      for (int j = threadIdx.x; j < 4096; j += 256) p[j] = v;
    }
  }
}
}  // namespace ibln

extern "C" int ibln_store_probe(void* out, int64_t total_bytes, int mode, int ctas, int device, void* stream) {
  using namespace ibln;
  if (!out || ctas < 1) return IBLN_EINVAL;
  DeviceGuard g(device);
  long long per = (total_bytes / ctas) / 65536 * 65536;
  IBLN_CUDA(cudaFuncSetAttribute(store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  store_probe_kernel<<<ctas, 256, 65536, (cudaStream_t)stream>>>((uint8_t*)out, per, mode);
  IBLN_RETURN_LAST();
}

// TMEM read-bandwidth probe: `warps` warps (warp w -> lane quarter w % 4) each issue `iters` tcgen05.ld
// 32x32b.x32 (4 KB per warp-instruction); out[0] = elapsed clocks of warp 0, out[1] = checksum.
namespace ibln {
__global__ void __launch_bounds__(512, 1) tmem_probe_kernel(long long* __restrict__ out, int iters, int depth) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32], w[32];
    tmem_ld32(base + ((i * 2) & 15) * 32, v);
    if (depth > 1) tmem_ld32(base + ((i * 2 + 1) & 15) * 32, w);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
    if (depth > 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= w[j];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; }
  if (acc == 0x12345678u) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_ptr, 512); }
}
}  // namespace ibln

extern "C" int ibln_tmem_probe(long long* out, int warps, int iters, int depth, int device, void* stream) {
  using namespace ibln;
  if (!out || warps < 1 || warps > 16) return IBLN_EINVAL;
  DeviceGuard g(device);
  tmem_probe_kernel<<<1, warps * 32, 0, (cudaStream_t)stream>>>(out, iters, depth);
  IBLN_RETURN_LAST();
}
#endif  // IBLN_DIAGNOSTICS

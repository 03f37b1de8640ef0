// Shared helpers for the sm_100a kernels of libiblnerf_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/iblnerf_b200.h"

namespace ibln {

struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (dev >= 0 && dev != prev) { cudaSetDevice(dev); changed = true; }
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

inline int num_sms(int device) {
  static int cached[64] = {0};
  int d = device < 0 ? 0 : device;
  if (d < 64 && cached[d]) return cached[d];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
  if (d < 64) cached[d] = n;
  return n;
}

#define IBLN_RETURN_LAST()                          \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    return e__ == cudaSuccess ? 0 : (int)e__;       \
  } while (0)

#define IBLN_CUDA(x)                                \
  do {                                              \
    cudaError_t e__ = (x);                          \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// sigmoid with MUFU exp2 / MUFU reciprocal (rel. error ~4e-7, far inside the 1e-4 parity budget).
// NB: __frcp_rn is the correctly-rounded (multi-instruction) reciprocal; __fdividef maps to one MUFU.RCP + FMUL.
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// pow(x + 1e-12, 1/2.2): ibl_nerf_renderer.py:26-27
__device__ __forceinline__ float srgbf(float x) { return powf(x + 1e-12f, 1.0f / 2.2f); }
__device__ __forceinline__ float dsrgbf(float x) { return (1.0f / 2.2f) * powf(x + 1e-12f, 1.0f / 2.2f - 1.0f); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// the same copy with an L2 eviction-priority policy (createpolicy.fractional.L2::evict_last / evict_first)
__device__ __forceinline__ void cp_async16_hint(void* smem, const void* gmem, uint64_t policy) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "l"(policy));
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}  // namespace ibln

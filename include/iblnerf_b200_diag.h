/* Tuning-build diagnostics of libiblnerf_b200 -- NOT part of the product ABI.
 * These entry points exist only in libiblnerf_b200_diag.so (python -m ibl_nerf_b200.build --diag, which compiles the
 * same sources with -DIBLN_DIAGNOSTICS); tools/*probe*.py and tools/timeline_*.py load that library via IBLN_LIB. */
#ifndef IBLNERF_B200_DIAG_H
#define IBLNERF_B200_DIAG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics: write `total_bytes` to `out` from `ctas` CTAs (mode 0/1: bulk TMA stores from shared memory,
 * 1 / 4 in flight; mode 2: coalesced st.global.v4) -- used to measure the achievable stash write bandwidth. */
int ibln_store_probe(void* out, int64_t total_bytes, int mode, int ctas, int device, void* stream);
/* Diagnostics: TMEM read bandwidth -- `warps` warps issue `iters` x `depth` tcgen05.ld 32x32b.x32 each;
 * out[0] = elapsed clocks of warp 0. */
int ibln_tmem_probe(long long* out, int warps, int iters, int depth, int device, void* stream);
/* Diagnostics (process-global, not thread-safe; tuning tools only): host-side switches of ibln_mlp_bwd
 * (bit 4 skip the dgrad launch, bit 5 skip the wgrad launch), and an optional device buffer (>= 8192 uint64) into
 * which CTAs 0/1 of the MLP kernels append (tag << 48 | clock64) marks of their pipeline phases (NULL = off). */
int ibln_debug_set(int flags);
int ibln_debug_timeline(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* IBLNERF_B200_DIAG_H */

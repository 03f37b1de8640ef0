// ABI bookkeeping entry points.
#include "common.cuh"

extern "C" int ibln_abi_version(void) { return IBLN_ABI_VERSION; }

extern "C" const char* ibln_error_string(int code) {
  if (code == 0) return "success";
  if (code == IBLN_EINVAL) return "iblnerf_b200: invalid argument";
  return cudaGetErrorString((cudaError_t)code);
}

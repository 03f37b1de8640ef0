// fp32 "exact" MLP path: positional encoding + generic SIMT GEMMs (forward, dgrad, wgrad).
// This is the high-precision mode used for the stage-wise 1e-4 parity tests and available for the
// normal estimator; the throughput path is the bf16 tcgen05 kernel in mlp_tc.cu.
#include "common.cuh"

namespace ibln {

// [x, sin(2^k x), cos(2^k x)]_k : positional_embedder.py:9-34 (frequency-major, sin then cos, xyz innermost)
__global__ void encode_kernel(const float* __restrict__ x, int64_t n_pts, int rep, int n_freqs, float* __restrict__ out,
                              int64_t ld) {
  int od = 3 + 6 * n_freqs;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pts * od) return;
  int64_t p = idx / od;
  int j = (int)(idx % od);
  const float* xp = x + (p / rep) * 3;
  float v;
  if (j < 3) {
    v = xp[j];
  } else {
    int k = (j - 3) / 6, rem = (j - 3) % 6, comp = rem % 3;
    float a = xp[comp] * exp2f((float)k);
    v = rem < 3 ? sinf(a) : cosf(a);
  }
  out[p * ld + j] = v;
}

constexpr int TM = 64, TN = 64, TK = 16;

// C = act(A * op(B) + bias) (+C); optional relu mask on the output (dgrad).
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int trans_b,
             const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int64_t M, int N, int K, int act,
             int accumulate, const float* __restrict__ mask, int64_t ld_mask) {
  __shared__ float sA[TK][TM + 1];
  __shared__ float sB[TK][TN + 1];
  int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  int64_t m0 = (int64_t)blockIdx.x * TM;
  int n0 = blockIdx.y * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      int mm = e / TK, kk = e % TK;
      int64_t gm = m0 + mm;
      int gk = k0 + kk;
      sA[kk][mm] = (gm < M && gk < K) ? A[gm * lda + gk] : 0.f;
    }
    for (int e = threadIdx.x; e < TN * TK; e += 256) {
      int nn, kk;
      float v = 0.f;
      if (trans_b) { nn = e / TK; kk = e % TK; if (n0 + nn < N && k0 + kk < K) v = B[(int64_t)(n0 + nn) * ldb + k0 + kk]; }
      else { kk = e / TN; nn = e % TN; if (n0 + nn < N && k0 + kk < K) v = B[(int64_t)(k0 + kk) * ldb + n0 + nn]; }
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? bias[gn] : 0.f);
      if (accumulate) v += C[gm * ldc + gn];
      if (act == 1) v = fmaxf(v, 0.f);
      if (mask != nullptr && !(mask[gm * ld_mask + gn] > 0.f)) v = 0.f;
      C[gm * ldc + gn] = v;
    }
  }
}

constexpr int WG_SPLITS = 64;

// partial[s][n][k] = sum over the s-th slice of rows of dY[m][n] * X[m][k]
__global__ void __launch_bounds__(256)
wgrad_partial_kernel(const float* __restrict__ dY, int64_t ldy, const float* __restrict__ X, int64_t ldx, int64_t M,
                     int N, int K, float* __restrict__ partial, float* __restrict__ partial_b) {
  __shared__ float sY[TK][TN + 1];
  __shared__ float sX[TK][TM + 1];
  int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  int n0 = blockIdx.x * TN, k0 = blockIdx.y * TM, s = blockIdx.z;
  int64_t per = (M + WG_SPLITS - 1) / WG_SPLITS;
  int64_t mb = s * per, me = mb + per < M ? mb + per : M;
  float acc[4][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t m0 = mb; m0 < me; m0 += TK) {
    for (int e = threadIdx.x; e < TK * TN; e += 256) {
      int mm = e / TN, nn = e % TN;
      sY[mm][nn] = (m0 + mm < me && n0 + nn < N) ? dY[(m0 + mm) * ldy + n0 + nn] : 0.f;
    }
    for (int e = threadIdx.x; e < TK * TM; e += 256) {
      int mm = e / TM, kk = e % TM;
      sX[mm][kk] = (m0 + mm < me && k0 + kk < K) ? X[(m0 + mm) * ldx + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < TK; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sY[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sX[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        accb[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gn = n0 + ty * 4 + i;
    if (gn >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gk = k0 + tx * 4 + j;
      if (gk < K) partial[((int64_t)s * N + gn) * K + gk] = acc[i][j];
    }
    if (partial_b != nullptr && blockIdx.y == 0 && tx == 0) partial_b[(int64_t)s * N + gn] = accb[i];
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b, int N, int K,
                                    float* __restrict__ dW, int64_t ldw, float* __restrict__ db, int accumulate) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nk = (int64_t)N * K;
  if (idx < nk) {
    float s = 0.f;
    for (int p = 0; p < WG_SPLITS; ++p) s += partial[p * nk + idx];
    int gn = (int)(idx / K), gk = (int)(idx % K);
    float* dst = dW + (int64_t)gn * ldw + gk;
    *dst = accumulate ? *dst + s : s;
  } else if (db != nullptr && idx < nk + N) {
    int gn = (int)(idx - nk);
    float s = 0.f;
    for (int p = 0; p < WG_SPLITS; ++p) s += partial_b[(int64_t)p * N + gn];
    db[gn] = accumulate ? db[gn] + s : s;
  }
}

}  // namespace ibln

using namespace ibln;

extern "C" int ibln_encode(const float* x, int64_t n_pts, int n_freqs, float* out, int64_t ld_out, int device, void* stream) {
  if (n_pts == 0) return 0;
  if (n_pts < 0 || n_freqs < 0 || !x || !out || ld_out < 3 + 6 * n_freqs) return IBLN_EINVAL;
  DeviceGuard g(device);
  int64_t tot = n_pts * (3 + 6 * n_freqs);
  encode_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n_pts, 1, n_freqs, out, ld_out);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_encode_dirs(const float* dirs, int64_t n_rays, int n_samples, int n_freqs, float* out, int64_t ld_out,
                                int device, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0 || n_samples < 1 || n_freqs < 0 || !dirs || !out || ld_out < 3 + 6 * n_freqs) return IBLN_EINVAL;
  DeviceGuard g(device);
  int64_t tot = n_rays * n_samples * (3 + 6 * n_freqs);
  encode_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dirs, n_rays * n_samples, n_samples, n_freqs, out, ld_out);
  IBLN_RETURN_LAST();
}

extern "C" int ibln_sgemm(const float* a, int64_t lda, const float* b, int64_t ldb, int trans_b, const float* bias, float* c,
                          int64_t ldc, int64_t m, int n, int k, int act, int accumulate, const float* relu_mask,
                          int64_t ld_mask, int device, void* stream) {
  if (m == 0) return 0;
  if (m < 0 || n < 1 || k < 1 || !a || !b || !c) return IBLN_EINVAL;
  DeviceGuard g(device);
  dim3 grid((unsigned)((m + TM - 1) / TM), (unsigned)((n + TN - 1) / TN));
  sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, trans_b, bias, c, ldc, m, n, k, act, accumulate,
                                                     relu_mask, ld_mask);
  IBLN_RETURN_LAST();
}

extern "C" int64_t ibln_wgrad_workspace_bytes(int n, int k) {
  return (int64_t)WG_SPLITS * ((int64_t)n * k + n) * (int64_t)sizeof(float);
}

extern "C" int ibln_sgemm_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t m, int n, int k,
                                float* dw, int64_t ldw, float* db, int accumulate, void* workspace, int device,
                                void* stream) {
  if (m < 0 || n < 1 || k < 1 || !dy || !x || !dw || !workspace) return IBLN_EINVAL;
  DeviceGuard g(device);
  float* partial = (float*)workspace;
  float* partial_b = partial + (int64_t)WG_SPLITS * n * k;
  dim3 grid((n + TN - 1) / TN, (k + TM - 1) / TM, WG_SPLITS);
  wgrad_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, ldy, x, ldx, m, n, k, partial, db ? partial_b : nullptr);
  int64_t tot = (int64_t)n * k + n;
  wgrad_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, partial_b, n, k, dw, ldw, db, accumulate);
  IBLN_RETURN_LAST();
}

// Dense Gram (Y^T Y), pairwise prediction and squared-error reduction.
//   gram      : Spark computeYtY (implicit mode), behind src/als_model.py:62
//   predict   : ALSModel.transform's fp32 dot, src/als_model.py:75
//   sse       : RMSE harness for the parity statement (not a reference function)
#include <cstdlib>

#include "common.cuh"

namespace hals {

constexpr int kGramThreads = 256;
constexpr int kGramRows = 32;        // rows staged per step
constexpr int kGramMaxBlocks = 592;  // 4 x 148 partial sums at most

// Partial Gram of a contiguous slab of rows, (KP/16)^2 register tile per thread.
template <int KP>
__global__ void __launch_bounds__(kGramThreads)
gram_partial_kernel(const float* __restrict__ src, int64_t n, int k, float* __restrict__ partial) {
  constexpr int TM = KP / 16;
  __shared__ __align__(16) float G[kGramRows * KP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < n ? r0 + per : n;
  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
  for (int64_t base = r0; base < r1; base += kGramRows) {
    __syncthreads();
    for (int e = tid; e < kGramRows * KP; e += kGramThreads) {
      const int t = e / KP, f = e - t * KP;
      const int64_t r = base + t;
      G[e] = (r < r1 && f < k) ? src[r * k + f] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < kGramRows; ++t) {
      float a[TM], b[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        a[i] = G[t * KP + ty + 16 * i];
        b[i] = G[t * KP + tx + 16 * i];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* P = partial + (size_t)blockIdx.x * KP * KP;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) P[(ty + 16 * i) * KP + tx + 16 * j] = acc[i][j];
}

// Fixed-order sum of the partials (deterministic), double accumulation, fp32 out [k,k].
__global__ void gram_reduce_kernel(const float* __restrict__ partial, int nblocks, int KP, int k,
                                   float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * k) return;
  const int i = e / k, j = e - i * k;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)partial[(size_t)b * KP * KP + i * KP + j];
  out[e] = (float)s;
}

int gram_tc(const float* src, int64_t n, int k, float* out, float* partial, cudaStream_t st);   // gram_tc.cu

static int gram_padded(int k) { return k <= 16 ? 16 : k <= 32 ? 32 : k <= 64 ? 64 : 128; }
static int gram_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > kGramMaxBlocks) b = kGramMaxBlocks;
  return (int)b;
}

__global__ void predict_kernel(const float* __restrict__ X, const float* __restrict__ Y, int k,
                               const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                               int64_t n, const uint8_t* __restrict__ up, const uint8_t* __restrict__ ip,
                               float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int u = users[p], i = items[p];
  const float* a = X + (int64_t)u * k;
  const float* b = Y + (int64_t)i * k;
  float s = 0.f;
  for (int f = 0; f < k; ++f) s = fmaf(a[f], b[f], s);
  if ((up && !up[u]) || (ip && !ip[i])) s = __int_as_float(0x7fc00000);
  out[p] = s;
}

// Per-block partial SSE in double, then a single-block fixed-order reduction.
__global__ void sse_partial_kernel(const float* __restrict__ X, const float* __restrict__ Y, int k,
                                   const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                                   const float* __restrict__ ratings, int64_t n, double* __restrict__ part) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const float* a = X + (int64_t)users[p] * k;
    const float* b = Y + (int64_t)items[p] * k;
    float d = 0.f;
    for (int f = 0; f < k; ++f) d = fmaf(a[f], b[f], d);
    const double e = (double)d - (double)ratings[p];
    s += e * e;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void sse_final_kernel(const double* __restrict__ part, int nb, int64_t n, double* sse, int64_t* count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += part[b];
    *sse = s;
    *count = n;
  }
}

}  // namespace hals

using namespace hals;

extern "C" size_t hals_gram_workspace_bytes(int k) {
  const int KP = gram_padded(k);
  return (size_t)kGramMaxBlocks * KP * KP * sizeof(float);
}

extern "C" int hals_gram(const float* src, int64_t n, int k, float* out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  HALS_REQUIRE(src && out && workspace, "null pointer");
  HALS_REQUIRE(k >= 1 && k <= 128 && n >= 0, "rank must be in [1,128]");
  if (workspace_bytes < hals_gram_workspace_bytes(k))
    return fail(HALS_ERR_WORKSPACE, "%s: workspace too small%s", __func__);
  cudaStream_t st = (cudaStream_t)stream;
  const int KP = gram_padded(k);
  const int nb = gram_blocks(n);
  float* part = (float*)workspace;
  // ranks 64 / 128: tcgen05 path (gram_tc.cu); HALS_FORCE_SIMT=1 keeps the CUDA-core kernel (A/B runs, small inputs)
  static const bool force_simt = [] { const char* e = getenv("HALS_FORCE_SIMT"); return e && e[0] == '1'; }();
  if ((k == 64 || k == 128) && n >= 2048 && !force_simt) return gram_tc(src, n, k, out, part, st);
  switch (KP) {
    case 16: gram_partial_kernel<16><<<nb, kGramThreads, 0, st>>>(src, n, k, part); break;
    case 32: gram_partial_kernel<32><<<nb, kGramThreads, 0, st>>>(src, n, k, part); break;
    case 64: gram_partial_kernel<64><<<nb, kGramThreads, 0, st>>>(src, n, k, part); break;
    default: gram_partial_kernel<128><<<nb, kGramThreads, 0, st>>>(src, n, k, part); break;
  }
  HALS_LAUNCH_CHECK();
  gram_reduce_kernel<<<(k * k + 255) / 256, 256, 0, st>>>(part, nb, KP, k, out);
  HALS_LAUNCH_CHECK();
  return 0;
}

extern "C" int hals_als_predict(const float* X, const float* Y, int k, const int32_t* users,
                                const int32_t* items, int64_t n, const uint8_t* user_present,
                                const uint8_t* item_present, float* out, void* stream) {
  HALS_REQUIRE(X && Y && users && items && out, "null pointer");
  HALS_REQUIRE(k >= 1, "rank must be positive");
  if (n == 0) return 0;
  predict_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      X, Y, k, users, items, n, user_present, item_present, out);
  HALS_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t hals_als_sse_workspace_bytes(void) { return 1024 * sizeof(double); }

extern "C" int hals_als_sse(const float* X, const float* Y, int k, const int32_t* users,
                            const int32_t* items, const float* ratings, int64_t n, double* sse,
                            int64_t* count, void* workspace, size_t workspace_bytes, void* stream) {
  HALS_REQUIRE(X && Y && users && items && ratings && sse && count, "null pointer");
  HALS_REQUIRE(workspace != nullptr, "null workspace");
  if (workspace_bytes < hals_als_sse_workspace_bytes()) return fail(HALS_ERR_WORKSPACE, "%s: workspace too small%s", __func__);
  HALS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "workspace must be 8-byte aligned");
  double* part = reinterpret_cast<double*>(workspace);   // per-block partial sums (caller-owned: nothing allocates here)
  int nb = (int)((n + 255) / 256);
  if (nb > 1024) nb = 1024;
  if (nb < 1) nb = 1;
  sse_partial_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(X, Y, k, users, items, ratings, n, part);
  HALS_LAUNCH_CHECK();
  sse_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(part, nb, n, sse, count);
  HALS_LAUNCH_CHECK();
  return 0;
}

// Shared host/device helpers for the hals_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>

#include "../../include/hals_b200.h"

namespace hals {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launch_count;

inline int fail(hals_status st, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
  return (int)st;
}

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define HALS_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) return ::hals::fail(HALS_ERR_INVALID, "%s: %s", __func__, msg); \
  } while (0)

#define HALS_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::hals::fail(HALS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));      \
  } while (0)

// Checked after every launch (cheap: no sync).
#define HALS_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    ::hals::count_launch();                                                              \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ::hals::fail(HALS_ERR_CUDA, "%s launch: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// `valid == false` zero-fills the destination (src-size 0: nothing is read from global)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid = true) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(valid ? 16 : 0));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool valid = true) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(valid ? 4 : 0));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace hals

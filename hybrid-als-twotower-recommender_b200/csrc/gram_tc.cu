// Dense Gram Y^T Y on the tensor cores (ranks 64 and 128): Spark's computeYtY (implicit feedback), reached from
// src/als_model.py:62.  The same error-compensated bf16 scheme as the normal-equation build (als_ws64.cu, als_tc128.cu):
// every factor row is split on the fly into h = bf16(y), l = bf16(y - h); per 16 rows ONE tcgen05.mma yields
//   rank 64 : A = [h ; l] (M = 128), B = h        (N = 64)   ->  D[0:64] = sum h h^T, D[64:128] = sum l h^T
//   rank 128: A = h       (M = 128), B = [h | l]  (N = 256)  ->  D[:, 0:128] = sum h h^T, D[:, 128:256] = sum h l^T
// with fp32 accumulation in TMEM over the CTA's whole slab of rows; G = D_hh + D_x + D_x^T (drops l l^T ~ 2^-18).
// The kernel is HBM-bound by construction (k/2 flop per byte, far below the tensor ridge): one pass over the fp32
// factors, coalesced 16-byte loads, conversion in registers, 8-byte swizzled stores into the MN-major operand
// stage (the layout tests/test_gpu_umma.py pins), double-buffered against the asynchronous MMAs.
// Persistent grid of one CTA per SM, contiguous slabs of rows; the per-CTA partials are summed in CTA order by
// gram_tc_finish_kernel (deterministic, fp64 accumulation like the CUDA-core path).
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace hals {

constexpr int kGtThreads = 256;
constexpr int kGtRows = 32;                 // rows per stage (two K = 16 steps)

template <int K>
struct GtCfg {
  static constexpr int kAtoms = K / 64;                       // 64-wide MN atoms per h (or l) block row
  static constexpr int kBlk = kGtRows * 128;                  // one [32][64] bf16 block
  static constexpr int kStage = 2 * kAtoms * kBlk;            // H atoms | L atoms
  static constexpr int kN = K == 64 ? 64 : 256;               // MMA N
  static constexpr int kCols = K == 64 ? 64 : 256;            // TMEM columns
};

template <int K>
__global__ void __launch_bounds__(kGtThreads, 1)
gram_tc_kernel(const float* __restrict__ src, int64_t n, float* __restrict__ partial) {
  using C = GtCfg<K>;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t st_free[2];
  __shared__ uint64_t acc_done;
  __shared__ uint32_t tmem_slot;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, C::kCols < 32 ? 32 : C::kCols);
  if (tid == 32) {
    umma::mbar_init(&st_free[0], 1);
    umma::mbar_init(&st_free[1], 1);
    umma::mbar_init(&acc_done, 1);
    umma::mbar_fence_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = umma::smem_u32(base);
  const int64_t per = ((n + gridDim.x - 1) / gridDim.x + kGtRows - 1) / kGtRows * kGtRows;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < n ? r0 + per : n;
  constexpr uint32_t idesc = umma::make_instr_desc(umma::kFmtBF16, true, true, 128, C::kN);
  constexpr int kVecPerRow = K / 4;                            // float4 per factor row
  constexpr int kVecPerThread = kGtRows * kVecPerRow / kGtThreads;
  uint32_t it = 0;
  for (int64_t rb = r0; rb < r1; rb += kGtRows, ++it) {
    const uint32_t s = it & 1, u = it >> 1;
    if (u > 0) umma::mbar_wait(&st_free[s], (u - 1) & 1);      // the MMAs that read this stage have completed
    uint8_t* st = base + s * C::kStage;
#pragma unroll
    for (int v = 0; v < kVecPerThread; ++v) {
      const int e = tid + v * kGtThreads;
      const int t = e / kVecPerRow, f = e % kVecPerRow;        // row of the chunk, float4 of the row (components 4f..4f+3)
      const int64_t r = rb + t;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < r1) y = *reinterpret_cast<const float4*>(src + r * K + 4 * f);
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(y.x - __low2float(h0), y.y - __high2float(h0));
      const __nv_bfloat162 l1 = __floats2bfloat162_rn(y.z - __low2float(h1), y.w - __high2float(h1));
      // component m = 4f lives in atom m / 64, 16-byte chunk (m % 64) / 8 (swizzled by the row), byte (m % 8) * 2
      const int atom = (4 * f) >> 6, c = ((4 * f) & 63) >> 3, o = ((4 * f) & 7) * 2;
      const uint32_t off = (uint32_t)atom * C::kBlk + (uint32_t)t * 128 + (uint32_t)((c ^ (t & 7)) << 4) + (uint32_t)o;
      uint2 hv, lv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
      lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
      *reinterpret_cast<uint2*>(st + off) = hv;
      *reinterpret_cast<uint2*>(st + C::kAtoms * C::kBlk + off) = lv;
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t sa = sbase + s * C::kStage;
#pragma unroll
      for (int ks = 0; ks < kGtRows / 16; ++ks) {
        // rank 64 : A walks H | L (M = 128), B = H (N = 64).  rank 128: A = H0 H1 (M = 128), B walks H0 H1 L0 L1 (N = 256)
        const uint64_t d = umma::make_smem_desc(sa + ks * 2048, C::kBlk, 1024, umma::kSwizzle128B);
        umma::mma_bf16(tmem, d, d, idesc, (it | (uint32_t)ks) != 0);
      }
      umma::commit(&st_free[s]);
    }
  }
  if (tid == 0) umma::commit(&acc_done);
  float* P = partial + (size_t)blockIdx.x * 128 * C::kCols;
  if (r0 < r1) {
    umma::mbar_wait(&acc_done, 0);
    umma::fence_after_sync();
    if (warp < 4) {                                             // TMEM lane = tid
      const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < C::kCols; c0 += 32) {
        float a[32];
        umma::tmem_ld32(ta + c0, a);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(P + (size_t)tid * C::kCols + c0 + j) = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
      }
    }
  } else {
    for (int e = tid; e < 128 * C::kCols; e += kGtThreads) P[e] = 0.f;   // a CTA without rows
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, C::kCols < 32 ? 32 : C::kCols);
}

// G[i][j] = sum over CTAs (fixed order, fp64) of  hh[i][j] + x[i][j] + x[j][i]
//   rank 64 : hh = D[i][j], x = D[64 + i][j]  (l h^T)       rank 128: hh = D[i][j], x = D[i][128 + j]  (h l^T)
template <int K>
__global__ void gram_tc_finish_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  constexpr int CO = GtCfg<K>::kCols;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * K) return;
  const int i = e / K, j = e - i * K;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    const float* D = partial + (size_t)b * 128 * CO;
    if (K == 64) s += (double)D[i * CO + j] + (double)D[(64 + i) * CO + j] + (double)D[(64 + j) * CO + i];
    else s += (double)D[i * CO + j] + (double)D[i * CO + 128 + j] + (double)D[j * CO + 128 + i];
  }
  out[e] = (float)s;
}

template <int K>
static int gram_tc_launch(const float* src, int64_t n, float* out, float* partial, cudaStream_t st) {
  using C = GtCfg<K>;
  const size_t smem = 2 * (size_t)C::kStage + 1024;
  HALS_CUDA(cudaFuncSetAttribute(gram_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = sm_count();
  const int64_t chunks = (n + kGtRows - 1) / kGtRows;
  if (grid > chunks) grid = chunks;
  if (grid > 148) grid = 148;                                   // the workspace is sized for 148 partials
  gram_tc_kernel<K><<<(unsigned)grid, kGtThreads, smem, st>>>(src, n, partial);
  HALS_LAUNCH_CHECK();
  gram_tc_finish_kernel<K><<<(K * K + 255) / 256, 256, 0, st>>>(partial, (int)grid, out);
  HALS_LAUNCH_CHECK();
  return 0;
}

int gram_tc(const float* src, int64_t n, int k, float* out, float* partial, cudaStream_t st) {
  return k == 64 ? gram_tc_launch<64>(src, n, out, partial, st) : gram_tc_launch<128>(src, n, out, partial, st);
}

}  // namespace hals

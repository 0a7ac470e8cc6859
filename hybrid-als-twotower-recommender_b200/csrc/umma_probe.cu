// Self-test entry for the tcgen05 path: runs `n_mma` UMMA instructions over a caller-supplied
// shared-memory image with caller-supplied descriptor fields and returns the TMEM accumulator.
// tests/test_gpu_umma.py uses it to pin the operand layouts (MN-major, 128B swizzle) that
// als_tc.cu and score_tc.cu rely on against a numpy product -- layout mistakes produce silent
// garbage, so they are checked in isolation.
#include "common.cuh"
#include "umma.cuh"

namespace hals {

__global__ void __launch_bounds__(128)
umma_probe_kernel(const uint8_t* __restrict__ img, int img_bytes, uint32_t a_off, uint32_t b_off,
                  uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t swizzle,
                  uint32_t idesc, int n_mma, uint32_t a_step, uint32_t b_step, int kind_tf32, int ncols,
                  float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  // dynamic smem base is not guaranteed 1024-aligned: align by hand
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid * 16; e < img_bytes; e += 128 * 16)
    *reinterpret_cast<uint4*>(base + e) = *reinterpret_cast<const uint4*>(img + e);
  umma::fence_proxy_async();
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 256);
  if (tid == 0) { umma::mbar_init(&mbar, 1); umma::mbar_fence_init(); }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    const uint32_t sb = umma::smem_u32(base);
    for (int s = 0; s < n_mma; ++s) {
      const uint64_t ad = umma::make_smem_desc(sb + a_off + s * a_step, a_lbo, a_sbo, swizzle);
      const uint64_t bd = umma::make_smem_desc(sb + b_off + s * b_step, b_lbo, b_sbo, swizzle);
      if (kind_tf32) umma::mma_tf32(tmem, ad, bd, idesc, s > 0);
      else umma::mma_bf16(tmem, ad, bd, idesc, s > 0);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  const int row = tid;  // warp w reads TMEM lanes [32w, 32w+32)
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < ncols) out[row * ncols + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace hals

using namespace hals;

extern "C" int hals_debug_umma_probe(const void* img, int img_bytes, uint32_t a_off, uint32_t b_off, uint32_t a_lbo,
                                     uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t swizzle, uint32_t idesc,
                                     int n_mma, uint32_t a_step, uint32_t b_step, int kind_tf32, int ncols,
                                     float* out, void* stream) {
  HALS_REQUIRE(img && out, "null pointer");
  HALS_REQUIRE(img_bytes > 0 && img_bytes % 16 == 0 && img_bytes <= 160 * 1024, "image must be 16B-multiple, <= 160 KiB");
  HALS_REQUIRE(ncols >= 8 && ncols <= 256 && n_mma >= 1, "bad shape");
  const size_t smem = (size_t)img_bytes + 1024;
  HALS_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)img, img_bytes, a_off, b_off, a_lbo, a_sbo,
                                                            b_lbo, b_sbo, swizzle, idesc, n_mma, a_step, b_step,
                                                            kind_tf32, ncols, out);
  HALS_LAUNCH_CHECK();
  return 0;
}

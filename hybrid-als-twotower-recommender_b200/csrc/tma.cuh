// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map construction through the driver entry
// point (no link-time dependency on libcuda) and the device-side load / mbarrier transaction ops.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "umma.cuh"

namespace hals {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D bf16 row-major matrix [rows][row_elems]; box = 64 elements (128 bytes) x box_rows rows, 128B swizzle.
// Out-of-bounds rows are zero-filled.  Returns false when the driver entry point is unavailable.
inline bool make_bf16_rowmajor_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t row_elems,
                                   uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {row_elems, rows};
  cuuint64_t strides[1] = {row_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(umma::smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(mbar)) : "memory");
}
// global (tensor map, coordinates {x = element, y = row}) -> shared, completion on the mbarrier
__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* mbar, int32_t x, int32_t y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      ::"r"(umma::smem_u32(smem_dst)), "l"(map), "r"(umma::smem_u32(mbar)), "r"(x), "r"(y) : "memory");
}

}  // namespace tma
}  // namespace hals

// ALS half-step, tensor-core path (explicit feedback, rank 64): the normal-equation build
// runs on tcgen05.mma with TMEM accumulators; replaces Spark's per-rating dspr/daxpy
// (NormalEquation.add) and CholeskySolver.solve reached from src/als_model.py:62.
//
// Precision: Spark accumulates in fp64; an fp32 tolerance needs better than one bf16 pass.
// Every source factor y is split once per half-step into two bf16 parts, h = bf16(y) and
// l = bf16(y - h) (|y - h - l| <= 2^-18 |y|), stored side by side ([n][h(64) | l(64)]).
//   A-operand  (M = 128, MN-major): rows 0..63 = h, rows 64..127 = l      (K = ratings)
//   B-operand  (N = 80,  MN-major): cols 0..63 = h, col 64 = bf16(r), col 65 = bf16(r - bf16(r))
//   D[0:64 , 0:64] = sum h h^T      D[64:128, 0:64] = sum l h^T
//   D[0:64 , 64:66] = sum h r_hi, sum h r_lo        D[64:128, 64] = sum l r_hi
//   A = D_hh + D_lh + D_lh^T  (drops only l l^T ~ 2^-18 relative),  b = the three r-columns.
// So ONE M=128 x N=80 x K=16 instruction per 16 ratings yields both the fp32-accurate Gram
// update and the right-hand side: 2.5 tensor cycles per rating per SM.
//
// Data movement: the h|l rows (256 B, contiguous) are gathered global->shared with cp.async
// straight into the 128B-swizzled MN-major layout the UMMA descriptors expect (layout pinned by
// tests/test_gpu_umma.py); 3-stage ring, stages recycled through tcgen05.commit -> mbarrier.
// Persistent CTAs (4 per SM: 128 TMEM columns each), static round-robin over work items, so
// while one CTA factorises a row (CUDA cores) the others keep the gather/tensor pipes busy.
#include <cuda_bf16.h>

#include "als_common.cuh"
#include "umma.cuh"

namespace hals {

constexpr int kTcThreads = 128;
constexpr int kTcKC = 32;          // ratings per stage
constexpr int kTcStages = 3;
constexpr int kTcAhead = kTcStages - 1;
constexpr int kTcK = 64;
constexpr int kTcRowBytes = 128;   // one 64-wide bf16 MN atom row
constexpr int kTcBlk = kTcKC * kTcRowBytes;          // 4096: one [KC][64] block
constexpr int kTcStageBytes = 3 * kTcBlk;            // H | L | R
constexpr int kTcLD = kTcK + 1;
constexpr int kTcS1Floats = ((kTcK + 1) * kTcLD + 3) / 4 * 4;
constexpr int kTcTmemCols = 128;
constexpr int kTcN = 80;

__global__ void split_bf16_kernel(const float* __restrict__ src, int64_t n_rows, int k,
                                  __nv_bfloat16* __restrict__ out) {
  // one thread per 8 consecutive factors: writes a 16-byte h chunk and a 16-byte l chunk
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = k >> 3;
  if (gid >= n_rows * per_row) return;
  const int64_t row = gid / per_row;
  const int c = (int)(gid - row * per_row);
  const float4 a = *reinterpret_cast<const float4*>(src + row * k + c * 8);
  const float4 b = *reinterpret_cast<const float4*>(src + row * k + c * 8 + 4);
  const float y[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2bfloat16_rn(y[i]);
    l[i] = __float2bfloat16_rn(y[i] - __bfloat162float(h[i]));
  }
  __nv_bfloat16* o = out + row * (2 * k);
  *reinterpret_cast<uint4*>(o + c * 8) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(o + k + c * 8) = *reinterpret_cast<const uint4*>(l);
}

__global__ void __launch_bounds__(kTcThreads, 4)
als_tc64_kernel(const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                const __nv_bfloat16* __restrict__ src_hl, float* __restrict__ dst, float reg,
                const int32_t* __restrict__ item_row, const int64_t* __restrict__ item_begin,
                const int32_t* __restrict__ item_len, const int32_t* __restrict__ item_slot,
                int64_t n_items, float* __restrict__ workspace) {
  constexpr int K = kTcK, KC = kTcKC, LD = kTcLD;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t mbar_free[kTcStages];
  __shared__ uint64_t mbar_acc;
  __shared__ uint32_t tmem_slot;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  float* S1 = reinterpret_cast<float*>(base + kTcStages * kTcStageBytes);  // normal matrix + rhs row
  float* S2 = reinterpret_cast<float*>(base);                              // aliases the stage ring (epilogue only)
  float* S2b = S2 + K * LD;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_sub = tid >> 4, piece = tid & 15;      // gather role: rating row within 8, 16-byte piece of the 256 B row
  if (warp == 0) umma::tmem_alloc(&tmem_slot, kTcTmemCols);
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) umma::mbar_init(&mbar_free[s], 1);
    umma::mbar_init(&mbar_acc, 1);
    umma::mbar_fence_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = umma::smem_u32(base);
  constexpr uint32_t idesc = umma::make_instr_desc(umma::kFmtBF16, true, true, 128, kTcN);

  uint32_t g = 0;         // chunks produced so far by this CTA (stage = g % stages)
  uint32_t acc_phase = 0;

  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int row = item_row[item];
    const int64_t begin = item_begin[item];
    const int len = item_len[item];
    const int slot = item_slot[item];
    const int nc = (len + KC - 1) / KC;

    auto produce = [&](int c) {
      const uint32_t gi = g + c, s = gi % kTcStages, u = gi / kTcStages;
      if (u > 0) umma::mbar_wait(&mbar_free[s], (u - 1) & 1);
      uint8_t* st = base + s * kTcStageBytes;
      // the 16 threads sharing t read one contiguous 256-byte h|l row
      const int blk_off = (piece < 8 ? 0 : kTcBlk);
      const int chunk = piece & 7;
#pragma unroll
      for (int i = 0; i < KC / 8; ++i) {
        const int t = t_sub + 8 * i;
        const int q = c * KC + t;
        const bool ok = q < len;
        const int ci = ok ? __ldg(colidx + begin + q) : 0;
        cp_async16(st + blk_off + t * kTcRowBytes + ((chunk ^ (t & 7)) << 4),
                   reinterpret_cast<const uint8_t*>(src_hl) + (size_t)ci * (4 * K) + piece * 16, ok);
      }
      if (warp == 0) {      // rating column(s) of the B operand: element 0 = bf16(r), element 1 = bf16(r - bf16(r))
        const int q = c * KC + lane;
        const float r = q < len ? __ldg(vals + begin + q) : 0.f;
        const __nv_bfloat16 rh = __float2bfloat16_rn(r);
        const __nv_bfloat16 rl = __float2bfloat16_rn(r - __bfloat162float(rh));
        const uint32_t packed = (uint32_t)__bfloat16_as_ushort(rh) | ((uint32_t)__bfloat16_as_ushort(rl) << 16);
        uint4 v = make_uint4(packed, 0u, 0u, 0u);
        // chunks 0 and 1 of the row (N columns 64..79), swizzled
        *reinterpret_cast<uint4*>(st + 2 * kTcBlk + lane * kTcRowBytes + ((0 ^ (lane & 7)) << 4)) = v;
        *reinterpret_cast<uint4*>(st + 2 * kTcBlk + lane * kTcRowBytes + ((1 ^ (lane & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
      cp_async_commit();
    };

    for (int c = 0; c < kTcAhead; ++c) {
      if (c < nc) produce(c); else cp_async_commit();
    }
    for (int c = 0; c < nc; ++c) {
      if (c + kTcAhead < nc) produce(c + kTcAhead); else cp_async_commit();
      cp_async_wait<kTcAhead>();
      umma::fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        umma::fence_after_sync();
        const uint32_t s = (g + c) % kTcStages;
        const uint32_t sa = sbase + s * kTcStageBytes;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          const uint64_t ad = umma::make_smem_desc(sa + ks * 2048, kTcBlk, 1024, umma::kSwizzle128B);
          const uint64_t bd = umma::make_smem_desc(sa + ks * 2048, 2 * kTcBlk, 1024, umma::kSwizzle128B);
          umma::mma_bf16(tmem, ad, bd, idesc, (c | ks) != 0);
        }
        umma::commit(&mbar_free[s]);
        if (c == nc - 1) umma::commit(&mbar_acc);
      }
    }
    g += nc;

    // ---- epilogue: TMEM -> registers -> combine through shared memory ------------------------------
    umma::mbar_wait(&mbar_acc, acc_phase);
    acc_phase ^= 1;
    umma::fence_after_sync();
    {
      const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
      float v0[32], v1[32], e[16];
      umma::tmem_ld32(ta, v0);
      umma::tmem_ld32(ta + 32, v1);
      umma::tmem_ld16(ta + 64, e);
      umma::fence_before_sync();
      if (tid >= 64) {  // l h^T rows
        float* r2 = S2 + (tid - 64) * LD;
#pragma unroll
        for (int n = 0; n < 32; ++n) { r2[n] = v0[n]; r2[32 + n] = v1[n]; }
        S2b[tid - 64] = e[0];
      }
      __syncthreads();
      if (tid < 64) {
        const int m = tid;
        float* r1 = S1 + m * LD;
#pragma unroll
        for (int n = 0; n < 32; ++n) {
          r1[n] = v0[n] + S2[m * LD + n] + S2[n * LD + m];
          r1[32 + n] = v1[n] + S2[m * LD + 32 + n] + S2[(32 + n) * LD + m];
        }
        S1[K * LD + m] = e[0] + e[1] + S2b[m];
      }
      if (tid == 0) S1[K * LD + K] = (float)len;
      __syncthreads();
    }
    if (slot >= 0) {
      // slice of a long row: park (A, b, n) in the slot (same layout as the SIMT path)
      float* W = workspace + (size_t)slot * ((size_t)K * K + K + 4);
      for (int e2 = tid; e2 < K * K; e2 += kTcThreads) W[e2] = S1[(e2 >> 6) * LD + (e2 & 63)];
      if (tid <= K) W[K * K + tid] = S1[K * LD + tid];
      __syncthreads();
    } else {
      if (tid < K) S1[tid * LD + tid] += reg * (float)len;
      __syncthreads();
      cholesky_solve_smem<K>(S1, K);
      if (tid < K) dst[(int64_t)row * K + tid] = S1[K * LD + tid];
      __syncthreads();
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTcTmemCols);
}

int als_half_step_tc64(const int32_t* colidx, const float* vals, const float* src, int64_t n_src, float* dst,
                       float reg, const hals_als_plan* plan, float* slots, void* split_buf, cudaStream_t st) {
  __nv_bfloat16* hl = reinterpret_cast<__nv_bfloat16*>(split_buf);
  const int64_t nthreads = n_src * (kTcK / 8);
  split_bf16_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(src, n_src, kTcK, hl);
  HALS_LAUNCH_CHECK();
  const size_t smem = (size_t)kTcStages * kTcStageBytes + kTcS1Floats * sizeof(float) + 1024;
  HALS_CUDA(cudaFuncSetAttribute(als_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = 4 * (int64_t)sm_count();
  if (grid > plan->n_items) grid = plan->n_items;
  als_tc64_kernel<<<(unsigned)grid, kTcThreads, smem, st>>>(colidx, vals, hl, dst, reg, plan->item_row,
                                                             plan->item_begin, plan->item_len, plan->item_slot,
                                                             plan->n_items, slots);
  HALS_LAUNCH_CHECK();
  return als_launch_reduce_solve(slots, dst, kTcK, reg, nullptr, plan, st);
}

}  // namespace hals

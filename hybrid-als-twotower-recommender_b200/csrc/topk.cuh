// Per-row streaming top-k shared by the scoring kernels and the cross-shard merge.
//
// A row's candidates are 64-bit keys: (order-preserving bits of the fp32 score) << 32 |
// (0xFFFFFFFF - item index), so "larger key" == "higher score, then lower item index" --
// the order of the reference's stable sorted(..., reverse=True)[:top_k] over candidates
// listed in ascending item order (src/hybrid_system.py:108).
//
// One warp owns a row.  It appends survivors (score > row threshold) to a buffer of CAP
// keys; when fewer than one tile's worth of free slots remain the warp sorts the buffer
// (bitonic, registers + shuffles), keeps the best k and raises the threshold to the k-th.
#pragma once
#include <stdint.h>

namespace hals {

__device__ __forceinline__ uint32_t f32_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t topk_key(float score, int32_t idx) {
  return ((uint64_t)f32_orderable(score) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)idx);
}
__device__ __forceinline__ float topk_key_score(uint64_t key) { return f32_from_orderable((uint32_t)(key >> 32)); }
__device__ __forceinline__ int32_t topk_key_index(uint64_t key) { return (int32_t)(0xFFFFFFFFu - (uint32_t)key); }

constexpr uint64_t kTopkEmpty = 0ull;  // below every real key (orderable(-inf) = 0x007fffff.. > 0)

__host__ __device__ inline int topk_capacity(int topk) { return topk <= 64 ? 128 : topk <= 128 ? 256 : 512; }

// Sort R*32 keys held as v[r] = element (r*32 + lane), descending.
template <int R>
__device__ __forceinline__ void warp_bitonic_desc(uint64_t (&v)[R], int lane) {
  constexpr int N = R * 32;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        constexpr int dummy = 0;
        (void)dummy;
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int pr = r ^ jr;
          if (pr > r) {
            const int e = r * 32 + lane;
            const bool desc = ((e & k) == 0);
            const uint64_t a = v[r], b = v[pr];
            const bool sw = desc ? (a < b) : (a > b);
            if (sw) { v[r] = b; v[pr] = a; }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = r * 32 + lane;
          const uint64_t other = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool desc = ((e & k) == 0);
          const bool lower = ((lane & j) == 0);
          // the lower index of the pair keeps the larger key when the run is descending
          const bool keep_max = (desc == lower);
          const uint64_t mx = v[r] > other ? v[r] : other;
          const uint64_t mn = v[r] > other ? other : v[r];
          v[r] = keep_max ? mx : mn;
        }
      }
    }
  }
}

// Compacts a row buffer (count valid keys of CAP) to its best `topk`; returns new count and
// writes the threshold key (k-th best, or kTopkEmpty while the row has < topk candidates).
template <int CAP>
__device__ __forceinline__ int topk_compact(uint64_t* buf, int count, int topk, int lane, uint64_t* thr_out) {
  constexpr int R = CAP / 32;
  uint64_t v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    v[r] = e < count ? buf[e] : kTopkEmpty;
  }
  warp_bitonic_desc<R>(v, lane);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    if (e < topk) buf[e] = v[r];
  }
  // k-th best lives at element topk-1
  uint64_t kth = kTopkEmpty;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t cand = __shfl_sync(0xffffffffu, v[r], (topk - 1) & 31);
    if (r == ((topk - 1) >> 5)) kth = cand;
  }
  __syncwarp();
  *thr_out = (count >= topk) ? kth : kTopkEmpty;
  return count < topk ? count : topk;
}

// Selection-only compaction for the streaming filter (no sort): finds the keep-th largest score of the
// row buffer by a 32-step bitwise search on the order-preserving score bits (one warp-wide count per
// step), then keeps every key whose score is >= that threshold, in place.  ~10x cheaper than the
// bitonic sort; the final ordering is established once, by topk_compact, when the row is finished.
// Returns the new count (min(count, keep): ties at the threshold are broken by item index) and the threshold score.
template <int CAP>
__device__ __forceinline__ int topk_select_compact(uint64_t* buf, int count, int keep, int lane, float* thr_out) {
  constexpr int R = CAP / 32;
  uint32_t hi[R], lo[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    const uint64_t k = e < count ? buf[e] : 0ull;
    hi[r] = (uint32_t)(k >> 32);
    lo[r] = (uint32_t)k;
  }
  uint32_t T = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) c += (hi[r] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= keep) T = cand;
  }
  // Exactly `keep` survive.  More than `keep` keys at or above T means score ties AT T: of those only the
  // keep - #(score > T) with the largest low words (= lowest item index, the reference's stable order) stay, found by
  // the same bitwise search on the low word.  Without this a row of mass ties stays above CAP - W after a compaction
  // and the next group of columns writes past the end of its buffer.
  int ge = 0, gt = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) { ge += (hi[r] >= T && hi[r] != 0u) ? 1 : 0; gt += (hi[r] > T) ? 1 : 0; }
  ge = __reduce_add_sync(0xffffffffu, ge);
  gt = __reduce_add_sync(0xffffffffu, gt);
  uint32_t Lmin = 0;
  if (ge > keep) {
    const int want = keep - gt;                           // >= 1: T is the keep-th largest score
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = Lmin | (1u << bit);
      int c = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) c += (hi[r] == T && lo[r] >= cand) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= want) Lmin = cand;
    }
  }
  int mine = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) mine += (hi[r] != 0u && (hi[r] > T || (hi[r] == T && lo[r] >= Lmin))) ? 1 : 0;
  int off = mine;                                        // inclusive warp scan
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, off, o);
    if (lane >= o) off += v;
  }
  const int total = __shfl_sync(0xffffffffu, off, 31);
  off -= mine;
  __syncwarp();                                          // every lane holds its keys in registers: safe to overwrite
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (hi[r] != 0u && (hi[r] > T || (hi[r] == T && lo[r] >= Lmin))) buf[off++] = ((uint64_t)hi[r] << 32) | lo[r];
  __syncwarp();
  *thr_out = f32_from_orderable(T);
  return total;
}

}  // namespace hals

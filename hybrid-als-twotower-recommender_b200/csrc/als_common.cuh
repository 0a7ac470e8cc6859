// Pieces shared by the SIMT and tensor-core ALS paths: tile geometry, workspace slot layout and
// the in-shared-memory Cholesky solve (Spark's CholeskySolver.solve -> LAPACK dppsv, behind
// src/als_model.py:62).
#pragma once
#include "common.cuh"

namespace hals {

constexpr int kAlsThreads = 256;
constexpr int kAlsChunk = 32;

template <int KP>
struct AlsTile {
  static constexpr int TM = KP / 16;            // tile edge per thread
  static constexpr int V = TM >= 4 ? 4 : TM;    // vector width of a shared read
  static constexpr int NG = TM / V;             // vector groups
  static constexpr int LD = KP + 1;             // padded leading dimension of A in smem
  // logical row/column owned by (group g, lane-in-vector v) of thread coordinate t
  __device__ static __forceinline__ int idx(int g, int v, int t) { return g * (16 * V) + t * V + v; }
};

template <int KP>
struct AlsSmem {
  float A[((KP + 1) * AlsTile<KP>::LD + 3) / 4 * 4];  // k x k normal matrix + rhs row (row KP), 16B multiple
  alignas(16) float G[2][kAlsChunk * KP];                   // gathered source rows, double buffered
  int idx[3][kAlsChunk];
  float wa[3][kAlsChunk];                       // weight of y y^T
  float wb[3][kAlsChunk];                       // weight of y in b
};

// Slot layout in the workspace: KP*KP (A, full square) + KP (b) + 1 (n) floats, padded to 4.
__host__ __device__ inline size_t als_slot_floats(int KP) { return (size_t)KP * KP + KP + 4; }

__host__ __device__ inline int als_padded_rank(int k) {
  return k <= 16 ? 16 : k <= 32 ? 32 : k <= 64 ? 64 : 128;
}

// In-place Cholesky of the leading k x k block of S (lower triangle), rhs in row KP,
// followed by the back substitution; result x is left in S[KP*LD + 0..k).
template <int KP>
__device__ void cholesky_solve_smem(float* S, int k) {
  constexpr int LD = AlsTile<KP>::LD;
  const int tid = threadIdx.x;
  // Cholesky-Crout, one column per iteration, one row per thread (rows j..k-1 and rhs row).
  for (int j = 0; j < k; ++j) {
    const int r = j + tid;                        // candidate row (r == k: the rhs row)
    const int row = (r == k) ? KP : r;
    const bool active = r <= k;
    float s = 0.f, d = 1.f;
    if (active) {
      const float* Lr = S + row * LD;
      const float* Lj = S + j * LD;
      float s0 = Lr[j], s1 = 0.f, d0 = Lj[j], d1 = 0.f;
      int p = 0;
      for (; p + 1 < j; p += 2) {
        const float a0 = Lj[p], a1 = Lj[p + 1];
        s0 = fmaf(-Lr[p], a0, s0);
        s1 = fmaf(-Lr[p + 1], a1, s1);
        d0 = fmaf(-a0, a0, d0);
        d1 = fmaf(-a1, a1, d1);
      }
      if (p < j) {
        const float a0 = Lj[p];
        s0 = fmaf(-Lr[p], a0, s0);
        d0 = fmaf(-a0, a0, d0);
      }
      // every thread derives the pivot itself: no barrier between pivot and column scale
      d = sqrtf(d0 + d1);
      s = s0 + s1;
    }
    __syncthreads();                              // all reads of A[j][j] done before it is overwritten
    if (active) S[row * LD + j] = (r == j) ? d : s / d;
    __syncthreads();                              // column j of L visible
  }
  // Back substitution L^T x = z by warp 0 (z = rhs row after the factorisation).
  if (tid < 32) {
    float* z = S + KP * LD;
    for (int j = k - 1; j >= 0; --j) {
      const float xj = z[j] / S[j * LD + j];
      __syncwarp();
      if (tid == 0) z[j] = xj;
      const float* Lj = S + j * LD;
      for (int i = tid; i < j; i += 32) z[i] = fmaf(-Lj[i], xj, z[i]);
      __syncwarp();
    }
  }
  __syncthreads();
}


// Long rows: sums the per-slice partials in slot order, adds Gram / ridge, solves (als_simt.cu).
int als_launch_reduce_solve(const float* workspace, float* dst, int k, float reg, const float* gram,
                            const hals_als_plan* plan, cudaStream_t st);

}  // namespace hals

// FP64 "NT" GEMM main loop on DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
//   acc[i][j] += sum_{k in [k_begin,k_end)} A[i][k] * B[j][k]
//
// Both operands are K-major (row-major with the contraction index contiguous), which is
// the natural layout of every dense step of the GP path once L^-1 is kept both as a lower
// matrix and as its transpose:
//   Cholesky panel solve / trailing update, TRTRI levels, LAUUM (K^-1 = L^-T L^-1) and the
//   posterior-variance panel V = K* L^-T all contract along contiguous rows.
// tcgen05 has no f64 kind, so the FP64 path is DMMA; the TF32 scoring path uses tcgen05.
//
// Tiles are staged with a 3-stage cp.async (LDGSTS) pipeline; fragments are read with
// conflict-free LDS.64 (row stride BK+4 doubles).
#pragma once
#include "common.cuh"

namespace bbo {

template <int BM_, int BN_, int WM_, int WN_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_;
  static constexpr int BK = 16;
  static constexpr int LDS = BK + 4;  // doubles per smem row: 20 -> half-warp LDS.64 conflict free
  static constexpr int STAGES = 3;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MF = WM / 8, NF = WN / 8;
  static constexpr int STAGE_DOUBLES = (BM + BN) * LDS;
  static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_DOUBLES * sizeof(double);
  static_assert(BM % WM == 0 && BN % WN == 0 && WM % 8 == 0 && WN % 8 == 0, "tile shape");
};

using Cfg128x64 = GemmCfg<128, 64, 32, 32>;  // 8 warps, 64 accumulator registers / thread
using Cfg64x64 = GemmCfg<64, 64, 32, 32>;    // 4 warps; every block boundary is 64-aligned
using Cfg64x32 = GemmCfg<64, 32, 32, 16>;    // 4 warps; short-list re-scoring (many small CTAs)
using Cfg64x64w8 = GemmCfg<64, 64, 16, 32>;  // 8 warps on one 64x64 tile: the single-launch small-problem fit

template <class Cfg>
struct Acc {
  double v[Cfg::MF][Cfg::NF][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NF; ++j) v[i][j][0] = v[i][j][1] = 0.0;
  }
};

template <class Cfg>
__device__ __forceinline__ void gemm_load_stage(double* st, const double* __restrict__ A, int64_t lda,
                                                const double* __restrict__ B, int64_t ldb, int k0) {
  // BK=16 doubles = 8 x 16-byte chunks per row
  constexpr int A_CHUNKS = Cfg::BM * 8, B_CHUNKS = Cfg::BN * 8;
  double* as = st;
  double* bs = st + Cfg::BM * Cfg::LDS;
#pragma unroll
  for (int c = threadIdx.x; c < A_CHUNKS; c += Cfg::THREADS) {
    int r = c >> 3, ch = c & 7;
    cp_async16(as + r * Cfg::LDS + ch * 2, A + int64_t(r) * lda + k0 + ch * 2);
  }
#pragma unroll
  for (int c = threadIdx.x; c < B_CHUNKS; c += Cfg::THREADS) {
    int r = c >> 3, ch = c & 7;
    cp_async16(bs + r * Cfg::LDS + ch * 2, B + int64_t(r) * ldb + k0 + ch * 2);
  }
}

// A: first row of the BM-row operand tile, column 0 of the contraction space.
// B: first row of the BN-row operand tile.  k_begin/k_end are multiples of 16.
// All threads of the CTA must call this; ends with a __syncthreads() so smem can be reused.
template <class Cfg>
__device__ __forceinline__ void gemm_nt_mainloop(const double* __restrict__ A, int64_t lda,
                                                 const double* __restrict__ B, int64_t ldb,
                                                 int k_begin, int k_end, Acc<Cfg>& acc, double* smem) {
  const int nk = (k_end - k_begin) / Cfg::BK;
  if (nk <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;

#pragma unroll
  for (int s = 0; s < Cfg::STAGES - 1; ++s) {
    if (s < nk) gemm_load_stage<Cfg>(smem + s * Cfg::STAGE_DOUBLES, A, lda, B, ldb, k_begin + s * Cfg::BK);
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<Cfg::STAGES - 2>();
    __syncthreads();
    {
      int nx = it + Cfg::STAGES - 1;
      if (nx < nk)
        gemm_load_stage<Cfg>(smem + (nx % Cfg::STAGES) * Cfg::STAGE_DOUBLES, A, lda, B, ldb,
                             k_begin + nx * Cfg::BK);
      cp_async_commit();
    }
    const double* as = smem + (it % Cfg::STAGES) * Cfg::STAGE_DOUBLES + (wm * Cfg::WM + g) * Cfg::LDS + t;
    const double* bs = smem + (it % Cfg::STAGES) * Cfg::STAGE_DOUBLES + Cfg::BM * Cfg::LDS +
                       (wn * Cfg::WN + g) * Cfg::LDS + t;
#pragma unroll
    for (int kk = 0; kk < Cfg::BK / 4; ++kk) {
      double a[Cfg::MF], b[Cfg::NF];
#pragma unroll
      for (int i = 0; i < Cfg::MF; ++i) a[i] = as[i * 8 * Cfg::LDS + kk * 4];
#pragma unroll
      for (int j = 0; j < Cfg::NF; ++j) b[j] = bs[j * 8 * Cfg::LDS + kk * 4];
#pragma unroll
      for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NF; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// One BK-wide step of the product from a staged tile pair (the body of the main loop above), for
// kernels that run their own (persistent, cross-tile) pipeline.
template <class Cfg>
__device__ __forceinline__ void gemm_compute_stage(const double* __restrict__ st, Acc<Cfg>& acc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  const double* as = st + (wm * Cfg::WM + g) * Cfg::LDS + t;
  const double* bs = st + Cfg::BM * Cfg::LDS + (wn * Cfg::WN + g) * Cfg::LDS + t;
#pragma unroll
  for (int kk = 0; kk < Cfg::BK / 4; ++kk) {
    double a[Cfg::MF], b[Cfg::NF];
#pragma unroll
    for (int i = 0; i < Cfg::MF; ++i) a[i] = as[i * 8 * Cfg::LDS + kk * 4];
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) b[j] = bs[j * 8 * Cfg::LDS + kk * 4];
#pragma unroll
    for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NF; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
  }
}

// Calls f(row, col, v0, v1) for every accumulator pair: (row, col) and (row, col+1) are tile-
// relative coordinates of v0 and v1.
template <class Cfg, class F>
__device__ __forceinline__ void acc_foreach(const Acc<Cfg>& acc, F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j)
      f(wm * Cfg::WM + i * 8 + g, wn * Cfg::WN + j * 8 + 2 * t, acc.v[i][j][0], acc.v[i][j][1]);
}

}  // namespace bbo

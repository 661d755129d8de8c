// GP posterior + acquisition scoring + top-k selection over a candidate pool.
//
// Replaces the body of BODesigner.score_fn (bbo/algorithms/designers/bo/bo.py:110-115):
//   GP.predict                bbo/algorithms/surrogates/gp/gp.py:75-83
//   EI / UCB                  bbo/algorithms/designers/bo/acquisitions.py:18-31
//   get_best_trials           bbo/shared/trial.py:278-294
//
// Per chunk of qc candidates (FP64 mode):
//   k_kstar     Ks[c,j] = s2 k(x_c, X_j)  (exact-difference form), mu_c = const + Ks[c,:] alpha
//   k_vnorm     |L^-1 k*_c|^2 = sum_i (sum_{j<=i} Ks[c,j] Linv[i,j])^2 : DMMA panel GEMM whose
//               epilogue squares and row-sums the accumulators, so V is never written
//   k_acq       var = k(x,x) - |v|^2 + noise, sd = sqrt(var), EI/UCB/PI
//   k_topk_seg  per-segment top-k (descending score, ties -> lowest index), applied in levels
// The TF32 mode swaps k_kstar/k_vnorm for the tcgen05 pipeline in score_tf32.cu.
#include "score.cuh"

#include "fit.cuh"
#include "gemm_f64.cuh"

namespace bbo {

constexpr int kDC = 32;
constexpr int kXS = kDC + 1;
constexpr int kSeg = 2048;  // top-k segment length

// same tile routine as fit.cu (kept local so each TU is self-contained)
template <typename T1>
__device__ __forceinline__ void dist_tile_q(const T1* __restrict__ X1, int64_t n1, int64_t i0,
                                            const double* __restrict__ X2, int64_t n2, int64_t j0, int d,
                                            const double* __restrict__ invl, double eps, double (&s)[4][4],
                                            double* x1s, double* x2s, double* ils) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) s[a][b] = 0.0;
  for (int kc = 0; kc < d; kc += kDC) {
    const int kw = min(kDC, d - kc);
    __syncthreads();
    // only the kw live columns are loaded (all lanes busy, coalesced when kw == d: 64 x d contiguous doubles);
    // the loop over all kDC columns left 2/3 of the lanes idle at d = 10
    for (int idx = tid; idx < 64 * kw; idx += 256) {
      const int r = idx / kw, k = idx - r * kw;
      double v1 = 0.0, v2 = 0.0;
      if (i0 + r < n1) v1 = static_cast<double>(X1[(i0 + r) * d + kc + k]);
      if (j0 + r < n2) v2 = X2[(j0 + r) * d + kc + k];
      x1s[r * kXS + k] = v1;
      x2s[r * kXS + k] = v2;
    }
    if (tid < kDC) ils[tid] = tid < kw ? invl[kc + tid] : 0.0;
    __syncthreads();
    for (int k = 0; k < kw; ++k) {
      const double il = ils[k];
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = x1s[(ty * 4 + a) * kXS + k];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = x2s[(tx + 16 * b) * kXS + k];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          double de = (bv[b] - av[a]) * il + eps;
          s[a][b] = fma(de, de, s[a][b]);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Ks (qc_pad, np) and mu.  grid (qc_pad/64), 256 threads.
// ------------------------------------------------------------------------------------------
template <typename T1>
__global__ void __launch_bounds__(256)
k_kstar(bbo_gp_hyper h, const T1* __restrict__ Xq, int64_t qc, const double* __restrict__ X, int64_t n,
        int64_t np, const double* __restrict__ hyp, const double* __restrict__ alpha,
        double* __restrict__ Ks, double* __restrict__ mu, int64_t qp) {
  __shared__ double x1s[64 * kXS], x2s[64 * kXS], ils[kDC];
  {   // batched launch (gridDim.z problems of one shape): every operand advances by one problem
    const int64_t b = blockIdx.z;
    Xq += b * qc * h.d;
    X += b * n * h.d;
    hyp += b * (2 * h.d + 2);
    alpha += b * np;
    Ks += b * qp * np;
    mu += b * qp;
  }
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t c0 = int64_t(blockIdx.x) * 64;
  const double eps = h.kind == BBO_KERNEL_MATERN52 ? h.pairwise_eps : 0.0;
  const double s2 = hyp[2 * h.d];
  double macc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j0 = 0; j0 < np; j0 += 64) {
    double s[4][4];
    dist_tile_q<T1>(Xq, qc, c0, X, n, j0, h.d, hyp + h.d, eps, s, x1s, x2s, ils);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t j = j0 + tx + 16 * b;
        double v = (j < n) ? s2 * kernel_from_s(h.kind, s[a][b]) : 0.0;
        Ks[(c0 + ty * 4 + a) * np + j] = v;
        macc[a] = fma(v, alpha[j], macc[a]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double v = macc[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tx == 0) mu[c0 + ty * 4 + a] = hyp[2 * h.d + 1] + v;
  }
}

// ------------------------------------------------------------------------------------------
// vpart[split, c] = sum over this split's row tiles i of (sum_{j<=i} Ks[c,j] Linv[i,j])^2
// grid (qc_pad/128, nsplit), Cfg128x64.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(Cfg128x64::THREADS)
k_vnorm(const double* __restrict__ Ks, const double* __restrict__ Linv, int64_t np, int64_t qc_pad,
        double* __restrict__ vpart) {
  using Cfg = Cfg128x64;
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[Cfg::BM][Cfg::WARPS_N];
  Ks += int64_t(blockIdx.z) * qc_pad * np;          // batched launch: one problem per blockIdx.z
  Linv += int64_t(blockIdx.z) * np * np;
  vpart += int64_t(blockIdx.z) * kMaxSplit * qc_pad;
  const int64_t c0 = int64_t(blockIdx.x) * Cfg::BM;
  const double* A = Ks + c0 * np;
  double rs[Cfg::MF];
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i) rs[i] = 0.0;
  const int ntile = int(np / 64);
  for (int it = blockIdx.y; it < ntile; it += gridDim.y) {
    Acc<Cfg> acc;
    acc.zero();
    gemm_nt_mainloop<Cfg>(A, np, Linv + int64_t(it) * 64 * np, np, 0, (it + 1) * 64, acc, smem);
#pragma unroll
    for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NF; ++j) {
        rs[i] = fma(acc.v[i][j][0], acc.v[i][j][0], rs[i]);
        rs[i] = fma(acc.v[i][j][1], acc.v[i][j][1], rs[i]);
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i) {
    double v = rs[i];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) red[wm * Cfg::WM + i * 8 + g][wn] = v;
  }
  __syncthreads();
  if (threadIdx.x < Cfg::BM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < Cfg::WARPS_N; ++w) v += red[threadIdx.x][w];
    vpart[int64_t(blockIdx.y) * qc_pad + c0 + threadIdx.x] = v;
  }
}

// V[c,i] = sum_{j<=i} Ks[c,j] Linv[i,j]  (only for the full-covariance predict).
// grid (qc_pad/64, np/64), Cfg64x64.
__global__ void __launch_bounds__(Cfg64x64::THREADS)
k_vstore(const double* __restrict__ Ks, const double* __restrict__ Linv, int64_t np, double* __restrict__ V) {
  using Cfg = Cfg64x64;
  extern __shared__ __align__(16) double smem[];
  const int64_t c0 = int64_t(blockIdx.x) * 64, i0 = int64_t(blockIdx.y) * 64;
  Acc<Cfg> acc;
  acc.zero();
  gemm_nt_mainloop<Cfg>(Ks + c0 * np, np, Linv + i0 * np, np, 0, int(i0) + 64, acc, smem);
  double* C = V + c0 * np + i0;
  acc_foreach<Cfg>(acc, [&](int r, int c, double v0, double v1) {
    *reinterpret_cast<double2*>(C + int64_t(r) * np + c) = make_double2(v0, v1);
  });
}

// C[i,j] -= sum_k V[i,k] V[j,k] for i,j < q (C is (q,q), ld = q).  grid (qp/64, qp/64), Cfg64x64.
__global__ void __launch_bounds__(Cfg64x64::THREADS)
k_cov_sub(const double* __restrict__ V, int64_t np, double* __restrict__ C, int64_t q) {
  using Cfg = Cfg64x64;
  extern __shared__ __align__(16) double smem[];
  const int64_t i0 = int64_t(blockIdx.x) * 64, j0 = int64_t(blockIdx.y) * 64;
  Acc<Cfg> acc;
  acc.zero();
  gemm_nt_mainloop<Cfg>(V + i0 * np, np, V + j0 * np, np, 0, int(np), acc, smem);
  acc_foreach<Cfg>(acc, [&](int r, int c, double v0, double v1) {
    int64_t i = i0 + r, j = j0 + c;
    if (i < q) {
      if (j < q) C[i * q + j] -= v0;
      if (j + 1 < q) C[i * q + j + 1] -= v1;
    }
  });
}

// ------------------------------------------------------------------------------------------
// acquisition (acquisitions.py:18-31) from mu and sum of |v|^2 partials.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T acq_value(const bbo_acq& a, T mu, T sd) {
  if (a.kind == BBO_ACQ_UCB) return mu + T(a.beta) * sd;
  const T imp = mu - T(a.best_f) - T(a.eps);
  const T z = imp / sd;
  const T Phi = T(0.5) * (T(1) + erf(z * T(0.70710678118654752440)));  // Normal(0,1).cdf
  if (a.kind == BBO_ACQ_PI) return Phi;
  const T phi = exp(T(-0.5) * z * z - T(0.91893853320467274178));      // exp(log_prob)
  return imp * Phi + sd * phi;
}

// |v|^2 of candidate c from its partial sums.  nsplit >= 0: that many slots, added in order.  nsplit < 0: one slot per
// 256-row block (-nsplit of them, persistent TF32 panel), added with the association of the four-slot form -- slot s
// of that form is the in-order sum of the blocks b = s, s+4, ... -- so both forms give the same bits.
__device__ __forceinline__ double sum_vpart(const double* __restrict__ vpart, int nsplit, int64_t qc_pad, int64_t c) {
  double vn = 0.0;
  if (nsplit >= 0) {
    for (int s = 0; s < nsplit; ++s) vn += vpart[int64_t(s) * qc_pad + c];
  } else {
    const int nb = -nsplit;
    for (int s = 0; s < 4; ++s) {
      double t = 0.0;
      for (int b = s; b < nb; b += 4) t += vpart[int64_t(b) * qc_pad + c];
      vn += t;
    }
  }
  return vn;
}

__global__ void __launch_bounds__(256)
k_acq(bbo_gp_hyper h, bbo_acq a, int has_acq, int clamp_var, const double* __restrict__ hyp,
      const double* __restrict__ mu, const double* __restrict__ vpart, int nsplit, int64_t qc_pad, int64_t qc,
      double* __restrict__ score_ws, double* __restrict__ mu_out, double* __restrict__ var_out,
      double* __restrict__ score_out, int32_t* __restrict__ nan_count) {
  const int64_t c = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (c >= qc) return;
  const double eps = h.kind == BBO_KERNEL_MATERN52 ? h.pairwise_eps : 0.0;
  const double kxx = hyp[2 * h.d] * kernel_from_s(h.kind, double(h.d) * eps * eps);
  const double vn = sum_vpart(vpart, nsplit, qc_pad, c);
  double var = kxx - vn + h.noise;
  if (clamp_var) var = fmax(var, 0.0);
  const double m = mu[c];
  if (mu_out) mu_out[c] = m;
  if (var_out) var_out[c] = var;
  if (has_acq) {
    const double sc = acq_value<double>(a, m, sqrt(var));
    score_ws[c] = sc;
    if (score_out) score_out[c] = sc;
    if (sc != sc && nan_count) atomicAdd(nan_count, 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_acq_elem(bbo_acq a, const T* __restrict__ mu, const T* __restrict__ sd, T* __restrict__ out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[i] = acq_value<T>(a, mu[i], sd[i]);
}

// k selection rounds over the kSeg (score, index) pairs staged in shared memory: per round a warp-shuffle
// arg-max with the (score desc, index asc) key, then thread 0 combines the 8 warps, emits the winner and
// retires it.  Must be called by all 256 threads after the staging stores (it starts with a barrier).
template <int N = kSeg>
__device__ __forceinline__ void select_rounds(double* ss, int64_t* si, double* ws, int64_t* wi, int* wp, int k,
                                              double* __restrict__ out_s, int64_t* __restrict__ out_i) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    double bs = -INFINITY;
    int64_t bi = -1;
    int bp = -1;
#pragma unroll
    for (int u = 0; u < N / 256; ++u) {
      const int p = tid + u * 256;
      if (better(ss[p], si[p], bs, bi)) {
        bs = ss[p];
        bi = si[p];
        bp = p;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double os = __shfl_xor_sync(0xffffffffu, bs, o);
      int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (better(os, oi, bs, bi)) {
        bs = os;
        bi = oi;
        bp = op;
      }
    }
    if (lane == 0) {
      ws[warp] = bs;
      wi[warp] = bi;
      wp[warp] = bp;
    }
    __syncthreads();
    if (tid == 0) {
      double fs = ws[0];
      int64_t fi = wi[0];
      int fp = wp[0];
      for (int w = 1; w < 8; ++w)
        if (better(ws[w], wi[w], fs, fi)) {
          fs = ws[w];
          fi = wi[w];
          fp = wp[w];
        }
      out_s[r] = fi >= 0 ? fs : -INFINITY;
      out_i[r] = fi;
      if (fp >= 0) si[fp] = -1;  // taken
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Segment top-k: block b scans entries [b*kSeg, (b+1)*kSeg) of (score, index) and emits its k
// best in order.  idx_in == nullptr -> index = base + position.  NaN scores and index < 0 are
// skipped; missing slots are (-inf, -1).  256 threads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_topk_seg(const double* __restrict__ score, const int64_t* __restrict__ idx_in, int64_t base, int64_t m,
           int k, double* __restrict__ out_s, int64_t* __restrict__ out_i) {
  __shared__ double ss[kSeg];
  __shared__ int64_t si[kSeg];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s0 = int64_t(blockIdx.x) * kSeg;
  for (int p = tid; p < kSeg; p += 256) {
    const int64_t g = s0 + p;
    double sc = -INFINITY;
    int64_t ix = -1;
    if (g < m) {
      sc = score[g];
      ix = idx_in ? idx_in[g] : base + g;
      if (sc != sc) ix = -1;
    }
    ss[p] = sc;
    si[p] = ix;
  }
  select_rounds(ss, si, ws, wi, wp, k, out_s + int64_t(blockIdx.x) * k, out_i + int64_t(blockIdx.x) * k);
}

// ------------------------------------------------------------------------------------------
// Fused acquisition epilogue + segment top-k (the default selection path): block b turns candidates
// [b*kSeg, (b+1)*kSeg) of a finished chunk into (variance, EI/UCB/PI score) in registers, stages the scores
// in shared memory and emits its k best -- the scores never go through global memory unless the caller asked
// for them.  Same arithmetic as k_acq; same ordering as k_topk_seg.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_acq_topk(bbo_gp_hyper h, bbo_acq a, int clamp_var, const double* __restrict__ hyp,
           const double* __restrict__ mu, const double* __restrict__ vpart, int nsplit, int64_t qc_pad, int64_t qc,
           int64_t base, int k, double* __restrict__ mu_out, double* __restrict__ var_out,
           double* __restrict__ score_out, int32_t* __restrict__ nan_count, double* __restrict__ out_s,
           int64_t* __restrict__ out_i, const bbo_acq* __restrict__ acq_batch) {
  __shared__ double ss[kSeg];
  __shared__ int64_t si[kSeg];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  const int tid = threadIdx.x;
  if (acq_batch) {   // batched launch: problem blockIdx.z has its own incumbent / parameters and slices
    const int64_t b = blockIdx.z;
    a = acq_batch[b];
    hyp += b * (2 * h.d + 2);
    mu += b * qc_pad;
    vpart += b * kMaxSplit * qc_pad;
    out_s += b * int64_t(gridDim.x) * k;
    out_i += b * int64_t(gridDim.x) * k;
  }
  const int64_t s0 = int64_t(blockIdx.x) * kSeg;
  const double eps = h.kind == BBO_KERNEL_MATERN52 ? h.pairwise_eps : 0.0;
  const double kxx = hyp[2 * h.d] * kernel_from_s(h.kind, double(h.d) * eps * eps);
  int nans = 0;
  for (int p = tid; p < kSeg; p += 256) {
    const int64_t c = s0 + p;
    double sc = -INFINITY;
    int64_t ix = -1;
    if (c < qc) {
      const double vn = sum_vpart(vpart, nsplit, qc_pad, c);
      double var = kxx - vn + h.noise;
      if (clamp_var) var = fmax(var, 0.0);
      const double m = mu[c];
      sc = acq_value<double>(a, m, sqrt(var));
      if (mu_out) mu_out[c] = m;
      if (var_out) var_out[c] = var;
      if (score_out) score_out[c] = sc;
      ix = base + c;
      if (sc != sc) {
        ix = -1;
        ++nans;
      }
    }
    ss[p] = sc;
    si[p] = ix;
  }
  if (nans && nan_count) atomicAdd(nan_count, nans);
  select_rounds(ss, si, ws, wi, wp, k, out_s + int64_t(blockIdx.x) * k, out_i + int64_t(blockIdx.x) * k);
}

// Merge of the per-segment lists (ma pairs) with the running top-k (k pairs, read before it is overwritten):
// one block, ma + k <= kSeg.
__global__ void __launch_bounds__(256)
k_topk_merge_running(const double* __restrict__ as, const int64_t* __restrict__ ai, int ma, int k,
                     double* __restrict__ run_s, int64_t* __restrict__ run_i) {
  as += int64_t(blockIdx.z) * ma;                  // batched launch: one list set per blockIdx.z
  ai += int64_t(blockIdx.z) * ma;
  run_s += int64_t(blockIdx.z) * k;
  run_i += int64_t(blockIdx.z) * k;
  __shared__ double ss[kSeg];
  __shared__ int64_t si[kSeg];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  for (int p = threadIdx.x; p < kSeg; p += 256) {
    double sc = -INFINITY;
    int64_t ix = -1;
    if (p < k) {
      sc = run_s[p];
      ix = run_i[p];
    } else if (p < k + ma) {
      sc = as[p - k];
      ix = ai[p - k];
    }
    if (sc != sc) ix = -1;
    ss[p] = sc;
    si[p] = ix;
  }
  select_rounds(ss, si, ws, wi, wp, k, run_s, run_i);
}

// ------------------------------------------------------------------------------------------
// Threshold-filtered selection (default).  After the first chunks of a pool almost no candidate can enter the
// running list any more, yet the round-based kernels above pay k (or, for the tf32+refine short-list, 2k + 64)
// block-wide arg-max rounds per chunk (51 + 59 us per chunk in round 2's launch list).  Here a candidate is compared
// with the CURRENT last entry of the running list (read at kernel start: stream-ordered behind the previous chunk's
// merge) and only candidates that beat it are appended, through a counter, to a survivor buffer; a block in which
// more than 64 candidates pass (first chunks, or a pool sorted by score) falls back to the arg-max rounds locally
// and appends its own K best.  k_merge_survivors then merges the few survivors into the running list by RANKING
// (every pair compared once, ~1 us) or, for many survivors, by arg-max rounds in batches.
// The result is the same set as before: the K best by (score desc, index asc) of a set does not depend on the
// order in which its members arrive, so atomics may order the survivor buffer freely.
// ------------------------------------------------------------------------------------------
// Two kernels: k_acq_eval (128 threads, two candidates each, no staging arrays) is light enough to run NEXT TO the
// persistent tensor-core kernels that hold every SM -- it used to be a 19-block, 33 KB-smem kernel that waited for
// them and then delayed the next panel by its own 43 us.  A block (256 candidates) with more than kEvalLocal passing
// candidates parks its scores in `score_tmp` and raises its flag; k_filter_heavy (one block per kSeg candidates,
// exits at once when no flag is up) reduces the flagged blocks of its segment to their K best.
constexpr int kEvalThreads = 128, kEvalPer = 2, kEvalBlock = kEvalThreads * kEvalPer;
constexpr int kEvalLocal = 32;     // more passing candidates than this in one block -> k_filter_heavy
constexpr int kFilterSlack = (kSeg / kEvalBlock - 1) * kEvalLocal;   // survivors of a segment: <= K + this

// counters[0] = survivors, counters[1] = flagged blocks
__global__ void __launch_bounds__(kEvalThreads)
k_acq_eval(bbo_gp_hyper h, bbo_acq a, int clamp_var, const double* __restrict__ hyp, const double* __restrict__ mu,
           const double* __restrict__ vpart, int nsplit, int64_t qc_pad, int64_t qc, int64_t base, int K,
           const double* __restrict__ run_s, const int64_t* __restrict__ run_i, double* __restrict__ mu_out,
           double* __restrict__ var_out, double* __restrict__ score_out, int32_t* __restrict__ nan_count,
           double* __restrict__ surv_s, int64_t* __restrict__ surv_i, int* __restrict__ counters,
           double* __restrict__ score_tmp, int* __restrict__ heavy) {
  __shared__ int wp[kEvalThreads / 32];
  const int tid = threadIdx.x;
  const int64_t s0 = int64_t(blockIdx.x) * kEvalBlock;
  const double eps = h.kind == BBO_KERNEL_MATERN52 ? h.pairwise_eps : 0.0;
  const double kxx = hyp[2 * h.d] * kernel_from_s(h.kind, double(h.d) * eps * eps);
  const double thr_s = run_s[K - 1];
  const int64_t thr_i = run_i[K - 1];
  double sc[kEvalPer];
  bool pass[kEvalPer];
  int nans = 0, npass = 0;
#pragma unroll
  for (int u = 0; u < kEvalPer; ++u) {
    const int64_t c = s0 + tid + kEvalThreads * u;
    sc[u] = nan("");
    pass[u] = false;
    if (c < qc) {
      const double vn = sum_vpart(vpart, nsplit, qc_pad, c);
      double var = kxx - vn + h.noise;
      if (clamp_var) var = fmax(var, 0.0);
      const double m = mu[c];
      const double v = acq_value<double>(a, m, sqrt(var));
      if (mu_out) mu_out[c] = m;
      if (var_out) var_out[c] = var;
      if (score_out) score_out[c] = v;
      if (v != v) ++nans;
      else if (better(v, base + c, thr_s, thr_i)) {
        sc[u] = v;
        pass[u] = true;
        ++npass;
      }
    }
  }
  if (nans && nan_count) atomicAdd(nan_count, nans);
  if (__syncthreads_count(npass > 0) == 0) return;   // nothing in this block beats the running list (the usual case)
  int cnt = npass;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) wp[tid >> 5] = cnt;
  __syncthreads();
  int block_pass = 0;
#pragma unroll
  for (int w = 0; w < kEvalThreads / 32; ++w) block_pass += wp[w];
  if (block_pass <= kEvalLocal) {
#pragma unroll
    for (int u = 0; u < kEvalPer; ++u)
      if (pass[u]) {
        const int p = atomicAdd(counters, 1);
        surv_s[p] = sc[u];
        surv_i[p] = base + s0 + tid + kEvalThreads * u;
      }
    return;
  }
  // many candidates pass (first chunks, or a pool sorted by score): park the scores (NaN = did not pass)
#pragma unroll
  for (int u = 0; u < kEvalPer; ++u) {
    const int64_t c = s0 + tid + kEvalThreads * u;
    if (c < qc) score_tmp[c] = sc[u];
  }
  if (tid == 0) {
    heavy[blockIdx.x] = 1;
    atomicAdd(counters + 1, 1);
  }
}

// grid = segments of kSeg candidates; a segment's flagged blocks -> their K best by arg-max rounds -> survivor list
__global__ void __launch_bounds__(256)
k_filter_heavy(int64_t qc, int64_t base, int K, const double* __restrict__ score_tmp, int* __restrict__ heavy,
               double* __restrict__ surv_s, int64_t* __restrict__ surv_i, int* __restrict__ counters) {
  if (counters[1] == 0) return;                 // no flag anywhere in this chunk (the usual case)
  __shared__ double ss[kSeg];
  __shared__ int64_t si[kSeg];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  __shared__ int s_base;
  const int tid = threadIdx.x;
  const int64_t s0 = int64_t(blockIdx.x) * kSeg;
  constexpr int kSub = kSeg / kEvalBlock;       // blocks of k_acq_eval per segment (thread tid + 256 u is in block u)
  const int64_t nsub = (qc + kEvalBlock - 1) / kEvalBlock;
  int total = 0;
#pragma unroll
  for (int u = 0; u < kSub; ++u) {
    const int64_t sub = int64_t(blockIdx.x) * kSub + u, c = s0 + tid + 256 * u;
    double v = -INFINITY;
    int64_t ix = -1;
    if (sub < nsub && heavy[sub] != 0 && c < qc) {
      const double t = score_tmp[c];
      if (t == t) {
        v = t;
        ix = base + c;
      }
    }
    ss[tid + 256 * u] = v;
    si[tid + 256 * u] = ix;
    total += __syncthreads_count(ix >= 0);
  }
  if (total == 0) return;
  if (tid < kSub && int64_t(blockIdx.x) * kSub + tid < nsub) heavy[int64_t(blockIdx.x) * kSub + tid] = 0;
  const int kk = total < K ? total : K;
  if (tid == 0) s_base = atomicAdd(counters, kk);
  __syncthreads();
  select_rounds(ss, si, ws, wi, wp, kk, surv_s + s_base, surv_i + s_base);
}

constexpr int kMergeCap = 1024;    // staged entries: 16 KB, so that a merge block fits beside a persistent panel CTA

// Many survivors (first chunk of a pool: every segment sends its K best): block b reduces survivors
// [b kMergeCap, (b+1) kMergeCap) to their K best in parallel with the others, into `lvl`; k_merge_survivors then
// merges ceil(c / kMergeCap) K entries instead of c.  Does nothing when the survivors fit one staged batch.
__global__ void __launch_bounds__(256)
k_merge_level(const double* __restrict__ surv_s, const int64_t* __restrict__ surv_i, const int* __restrict__ surv_count,
              int K, double* __restrict__ lvl_s, int64_t* __restrict__ lvl_i) {
  const int c = *surv_count;
  if (c + K <= kMergeCap) return;
  const int b0 = int(blockIdx.x) * kMergeCap;
  if (b0 >= c) return;
  __shared__ double ss[kMergeCap];
  __shared__ int64_t si[kMergeCap];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  for (int p = threadIdx.x; p < kMergeCap; p += 256) {
    const bool in = b0 + p < c;
    ss[p] = in ? surv_s[b0 + p] : -INFINITY;
    si[p] = in ? surv_i[b0 + p] : -1;
  }
  select_rounds<kMergeCap>(ss, si, ws, wi, wp, K, lvl_s + int64_t(blockIdx.x) * K, lvl_i + int64_t(blockIdx.x) * K);
}

// Merge the survivors of one chunk (or their per-block K best, see k_merge_level) into the running list (K entries,
// sorted) and reset the counters.  One block.
__global__ void __launch_bounds__(256)
k_merge_survivors(const double* __restrict__ surv_s, const int64_t* __restrict__ surv_i, int* __restrict__ surv_count,
                  int K, double* __restrict__ run_s, int64_t* __restrict__ run_i, const double* __restrict__ lvl_s,
                  const int64_t* __restrict__ lvl_i) {
  constexpr int kCap = kMergeCap;
  __shared__ double ss[kCap];
  __shared__ int64_t si[kCap];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  const int tid = threadIdx.x;
  int c = *surv_count;
  if (c + K > kCap) {            // k_merge_level ran: its lists are the input
    c = (c + kCap - 1) / kCap * K;
    surv_s = lvl_s;
    surv_i = lvl_i;
  }
  __syncthreads();
  if (tid == 0) {
    surv_count[0] = 0;
    surv_count[1] = 0;
  }
  if (c == 0) return;
  if (K + c <= 512) {
    // ranking merge: position of an element = number of elements that beat it (keys are distinct: indices are)
    const int nu = K + c;
    for (int p = tid; p < nu; p += 256) {
      ss[p] = p < K ? run_s[p] : surv_s[p - K];
      si[p] = p < K ? run_i[p] : surv_i[p - K];
    }
    __syncthreads();
    int rank[2] = {-1, -1};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + 256 * u;
      if (e < nu && si[e] >= 0) {
        const double es = ss[e];
        const int64_t ei = si[e];
        int r = 0;
        for (int x = 0; x < nu; ++x) r += better(ss[x], si[x], es, ei) ? 1 : 0;
        rank[u] = r;
      }
    }
    __syncthreads();
    for (int p = tid; p < K; p += 256) {       // fewer than K valid entries leave (-inf, -1) behind them
      run_s[p] = -INFINITY;
      run_i[p] = -1;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + 256 * u;
      if (rank[u] >= 0 && rank[u] < K) {
        run_s[rank[u]] = ss[e];
        run_i[rank[u]] = si[e];
      }
    }
    return;
  }
  // many survivors (first chunks): arg-max rounds over [running list | batch of survivors], batch after batch
  for (int b0 = 0; b0 < c; b0 += kCap - K) {
    const int nb = (c - b0 < kCap - K) ? c - b0 : kCap - K;
    for (int p = tid; p < kCap; p += 256) {
      double sc = -INFINITY;
      int64_t ix = -1;
      if (p < K) {
        sc = run_s[p];
        ix = run_i[p];
      } else if (p < K + nb) {
        sc = surv_s[b0 + p - K];
        ix = surv_i[b0 + p - K];
      }
      ss[p] = sc;
      si[p] = ix;
    }
    select_rounds<kCap>(ss, si, ws, wi, wp, K, run_s, run_i);
    __syncthreads();
  }
}

__global__ void k_topk_init(double* s, int64_t* i, int k) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < k) {
    s[t] = -INFINITY;
    i[t] = -1;
  }
}

// reduce m (score,index) pairs in buf A to k pairs in (out_s,out_i), ping-ponging through buf B
static int topk_levels(double* as, int64_t* ai, int64_t m, double* bs, int64_t* bi, int k, double* out_s,
                       int64_t* out_i, cudaStream_t st) {
  while (true) {
    const int64_t nb = ceil_div(m, kSeg);
    if (nb == 1) {
      k_topk_seg<<<1, 256, 0, st>>>(as, ai, 0, m, k, out_s, out_i);
      BBO_LAUNCH_CHECK();
      return BBO_OK;
    }
    k_topk_seg<<<unsigned(nb), 256, 0, st>>>(as, ai, 0, m, k, bs, bi);
    BBO_LAUNCH_CHECK();
    double* ts = as;
    as = bs;
    bs = ts;
    int64_t* ti = ai;
    ai = bi;
    bi = ti;
    m = nb * k;
  }
}

// (score, index) pairs <-> one interleaved int64 buffer [bits(score_0), index_0, bits(score_1), ...]: the message
// of the single all-gather of the multi-GPU path, packed and merged without any framework-side reshuffling.
__global__ void k_topk_pack(const double* __restrict__ s, const int64_t* __restrict__ i, int k,
                            int64_t* __restrict__ packed) {
  const int t = threadIdx.x;
  if (t < k) {
    packed[2 * t] = __double_as_longlong(s[t]);
    packed[2 * t + 1] = i[t];
  }
}
__global__ void __launch_bounds__(256)
k_topk_merge_packed(const int64_t* __restrict__ packed, int m, int k, double* __restrict__ out_s,
                    int64_t* __restrict__ out_i) {
  __shared__ double ss[kSeg];
  __shared__ int64_t si[kSeg];
  __shared__ double ws[8];
  __shared__ int64_t wi[8];
  __shared__ int wp[8];
  for (int p = threadIdx.x; p < kSeg; p += 256) {
    double sc = -INFINITY;
    int64_t ix = -1;
    if (p < m) {
      sc = __longlong_as_double(packed[2 * p]);
      ix = packed[2 * p + 1];
      if (sc != sc) ix = -1;
    }
    ss[p] = sc;
    si[p] = ix;
  }
  select_rounds(ss, si, ws, wi, wp, k, out_s, out_i);
}
int topk_pack(const double* score, const int64_t* index, int k, int64_t* packed, cudaStream_t st) {
  if (k < 1 || k > BBO_TOPK_MAX) return BBO_ERR_BAD_ARG;
  k_topk_pack<<<1, BBO_TOPK_MAX, 0, st>>>(score, index, k, packed);
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}
int topk_merge_packed(const int64_t* packed, int64_t m, int k, double* out_s, int64_t* out_i, cudaStream_t st) {
  if (k < 1 || k > BBO_TOPK_MAX || m < 0) return BBO_ERR_BAD_ARG;
  if (m > kSeg) return BBO_ERR_BAD_SHAPE;
  k_topk_merge_packed<<<1, 256, 0, st>>>(packed, int(m), k, out_s, out_i);
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

int topk_merge(const double* score, const int64_t* index, int64_t m, int k, double* out_s, int64_t* out_i,
               cudaStream_t st) {
  if (k < 1 || k > BBO_TOPK_MAX || m < 0) return BBO_ERR_BAD_ARG;
  if (m > kSeg) return BBO_ERR_BAD_SHAPE;  // gathered lists are tiny: G * k <= 8 * 128
  k_topk_seg<<<1, 256, 0, st>>>(score, index, 0, m, k, out_s, out_i);
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

// ==========================================================================================
// workspace + drivers
// ==========================================================================================
// chunks are padded to 256 candidates: the tensor-core panel works on PAIRS of 128-candidate tiles (cta_group::2)
static int64_t round_chunk(int64_t q_chunk) { return ceil_div(q_chunk < 256 ? 256 : q_chunk, 256) * 256; }

// Number of row-tile splits of the variance panel.  It must depend on n ONLY: the partial sums of
// a candidate are then associated identically for every chunking / sharding of the pool, which is
// what makes the multi-GPU selection bit-identical to the single-GPU one.
static int pick_nsplit(int64_t /*qc_pad*/, int64_t np) {
  int64_t want = np / 64;
  if (want > 8) want = 8;
  if (want < 1) want = 1;
  return int(want);
}

static size_t refine_bytes(int64_t np, int64_t d);
size_t ScoreWorkspace::bytes(int64_t n, int64_t d, int64_t q_chunk, int precision, bool want_full) {
  const int64_t np = padded(n), qp = round_chunk(q_chunk);
  const int64_t nblk = ceil_div(qp, kSeg);
  const bool tf32 = precision != BBO_PREC_FP64;
  const size_t ks = (tf32 && !want_full) ? 0 : size_t(qp) * np;   // TF32 keeps K* in float32
  size_t doubles = ks * (want_full ? 2 : 1)  // Ks (+V)
                   + size_t(kMaxSplit) * qp + 2 * size_t(qp)  // vpart, mu, score
                   + 2 * (size_t(512) * (nblk + 1) + 256) * 2     // two (score,index) lists
                   + 32 + size_t(qp) / 512 + 8;                   // survivor counters + one flag per 256 candidates
  size_t extra = 0;
  if (tf32) extra = score_tf32_extra_bytes(n, d, qp);
  if (precision == BBO_PREC_TF32_REFINE) extra += refine_bytes(np, d);
  return doubles * sizeof(double) + extra + 512;
}

int ScoreWorkspace::bind(void* ws, size_t ws_bytes, int64_t n, int64_t d, int64_t q_chunk, int precision,
                         bool want_full) {
  np = padded(n);
  qp = round_chunk(q_chunk);
  if (ws == nullptr || ws_bytes < bytes(n, d, q_chunk, precision, want_full)) return BBO_ERR_WORKSPACE;
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  double* q = reinterpret_cast<double*>(p);
  Ks = q;
  const bool tf32m = precision != BBO_PREC_FP64;
  if (!(tf32m && !want_full)) q += size_t(qp) * np;
  V = want_full ? q : nullptr;
  if (want_full) q += size_t(qp) * np;
  vpart = q;
  q += size_t(kMaxSplit) * qp;
  mu = q;
  q += qp;
  score = q;
  q += qp;
  const int64_t nblk = ceil_div(qp, kSeg);
  list_len = int64_t(512) * (nblk + 1) + 256;
  la_s = q;
  q += list_len;
  lb_s = q;
  q += list_len;
  la_i = reinterpret_cast<int64_t*>(q);
  q += list_len;
  lb_i = reinterpret_cast<int64_t*>(q);
  q += list_len;
  surv_count = reinterpret_cast<int*>(q);       // [0] survivors, [1] flagged blocks, [64 ..] the flags
  q += 32 + size_t(qp) / 512 + 8;
  tf32 = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(q) + 255) & ~uintptr_t(255));
  refine = nullptr;
  if (precision == BBO_PREC_TF32_REFINE)
    refine = static_cast<char*>(tf32) + ((score_tf32_extra_bytes(n, d, qp) + 255) & ~size_t(255));
  return BBO_OK;
}


// ==========================================================================================
// BBO_PREC_TF32_REFINE: short-list by TF32 score, float64 re-score of the short-list
// ==========================================================================================
constexpr int kRefineRows = 256;   // capacity of the short-list (refine_shortlist_len(k) <= this)
int refine_shortlist_len(int k) {
  int m = 8 * k;
  if (m < 64) m = 64;
  if (m > kRefineRows) m = kRefineRows;
  return m;
}
static int refine_kseg(int k) { return 2 * k; }   // entries kept per 2048-candidate segment

struct RefineScratch {
  double* sh_s;        // (256) short-list TF32 scores, descending
  int64_t* sh_i;       // (256) their global indices (-1 = empty slot)
  double* bound;       // (1)   max TF32 score dropped at the segment level
  double* info;        // (4)   see bbo_gp_refine_info
  double* Xg[2];       // (256, d) rows of the short-listed candidates, ping-pong across continued calls
  int64_t* gi[2];      // (256) indices the rows of Xg[p] belong to
  double* tf_s;        // (256) TF32 scores in the order of the gathered rows
  double* Ks;          // (256, np) float64 cross-covariance of the short-list
  double* mu;          // (np/64, 256) partial mean sums, one slot per 64-observation tile
  double* part;        // (np/32, 256) |v|^2 partial sums, one slot per 32-row block of L^-1
  double* la_s;        // (512) selection list: running top-k, then the re-scored short-list
  int64_t* la_i;
};
static size_t refine_bytes(int64_t np, int64_t d) {
  const size_t R = kRefineRows;
  size_t doubles = R + R + 8 + 8 + 2 * R * size_t(d) + 2 * R + R + R * size_t(np) + size_t(np / 64) * R +
                   size_t(np / 32) * R + 4 * R;
  return doubles * sizeof(double) + 1024;
}
static RefineScratch carve_refine(void* base, int64_t np, int64_t d) {
  RefineScratch r;
  const size_t R = kRefineRows;
  double* q = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
  r.sh_s = q;  q += R;
  r.sh_i = reinterpret_cast<int64_t*>(q);  q += R;
  r.bound = q;  q += 8;
  r.info = q;  q += 8;
  for (int p = 0; p < 2; ++p) { r.Xg[p] = q;  q += R * size_t(d); }
  for (int p = 0; p < 2; ++p) { r.gi[p] = reinterpret_cast<int64_t*>(q);  q += R; }
  r.tf_s = q;  q += R;
  r.Ks = q;  q += R * size_t(np);
  r.mu = q;  q += size_t(np / 64) * R;
  r.part = q;  q += size_t(np / 32) * R;
  r.la_s = q;  q += 2 * R;
  r.la_i = reinterpret_cast<int64_t*>(q);  q += 2 * R;
  return r;
}

__global__ void k_refine_init(double* sh_s, int64_t* sh_i, double* bound, int64_t* gi0, int64_t* gi1, int m) {
  const int t = threadIdx.x;
  if (t < m) {
    sh_s[t] = -INFINITY;
    sh_i[t] = -1;
    gi0[t] = -1;      // row-buffer indices: slots past the short-list length must never match a carried index
    gi1[t] = -1;
  }
  if (t == 0) *bound = -INFINITY;
}

// bound = max(bound, the kseg-th (last kept) TF32 score of every segment list): everything a segment dropped
// scored no more than that.  Lists with fewer than kseg valid entries end in (-inf, -1): nothing was dropped.
__global__ void __launch_bounds__(256)
k_refine_bound(const double* __restrict__ seg_s, const int64_t* __restrict__ seg_i, int nb, int kseg,
               double* __restrict__ bound) {
  __shared__ double red[256];
  double v = -INFINITY;
  for (int b = threadIdx.x; b < nb; b += 256) {
    const int64_t p = int64_t(b) * kseg + kseg - 1;
    if (seg_i[p] >= 0) v = fmax(v, seg_s[p]);
  }
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *bound = fmax(*bound, red[0]);
}

// Rows of the short-listed candidates -> Xg_new (float64), in short-list order.  A candidate of this call's pool
// [lo, lo+q) is read from Xq; one carried over from an earlier (continued) call is found by index in the previous
// row buffer.  One block of 256 threads, one warp per row; rows [0, rows) are written (rows = m rounded up to 64).
template <typename T1>
__global__ void __launch_bounds__(256)
k_refine_gather(const T1* __restrict__ Xq, int64_t lo, int64_t q, int d, const double* __restrict__ sh_s,
                const int64_t* __restrict__ sh_i, int m, int rows, const double* __restrict__ Xg_prev,
                const int64_t* __restrict__ gi_prev, double* __restrict__ Xg_new, int64_t* __restrict__ gi_new,
                double* __restrict__ tf_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < rows; j += 8) {
    const int64_t ix = j < m ? sh_i[j] : -1;
    const bool local = ix >= lo && ix < lo + q;
    int sp = -1;
    if (ix >= 0 && !local && gi_prev != nullptr) {
      for (int u = 0; u < kRefineRows / 32 && sp < 0; ++u) {
        const unsigned hit = __ballot_sync(0xffffffffu, gi_prev[lane + 32 * u] == ix);
        if (hit) sp = 32 * u + __ffs(int(hit)) - 1;
      }
    }
    for (int c = lane; c < d; c += 32) {
      double v = 0.0;
      if (ix >= 0 && local) v = static_cast<double>(Xq[(ix - lo) * d + c]);
      else if (ix >= 0 && sp >= 0) v = Xg_prev[int64_t(sp) * d + c];
      Xg_new[int64_t(j) * d + c] = v;
    }
    if (lane == 0) {
      // (every carried index is in gi_prev; if it were not, the row is dropped rather than scored as zeros)
      gi_new[j] = (ix >= 0 && !local && sp < 0) ? -1 : ix;
      tf_s[j] = j < m ? sh_s[j] : -INFINITY;
    }
  }
}

// Float64 cross-covariance of the short-list against ONE 64-observation tile per CTA (k_kstar walks all tiles in one
// CTA: fine for a pool, 0.69 ms for a 64-row short-list at n = 4096).  Ks (rows, np) and, per observation tile, the
// partial mean sums mupart[jt, c] = sum_{j in tile jt} Ks[c,j] alpha[j], added in tile order by k_refine_acq.
// grid (np/64, rows/64), 256 threads.
__global__ void __launch_bounds__(256)
k_refine_kstar(bbo_gp_hyper h, const double* __restrict__ Xg, int64_t rows, const double* __restrict__ X, int64_t n,
               int64_t np, const double* __restrict__ hyp, const double* __restrict__ alpha,
               double* __restrict__ Ks, double* __restrict__ mupart) {
  __shared__ double x1s[64 * kXS], x2s[64 * kXS], ils[kDC];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t j0 = int64_t(blockIdx.x) * 64, c0 = int64_t(blockIdx.y) * 64;
  const double eps = h.kind == BBO_KERNEL_MATERN52 ? h.pairwise_eps : 0.0;
  const double s2 = hyp[2 * h.d];
  double s[4][4];
  dist_tile_q<double>(Xg, rows, c0, X, n, j0, h.d, hyp + h.d, eps, s, x1s, x2s, ils);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double macc = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = j0 + tx + 16 * b;
      const double v = (j < n) ? s2 * kernel_from_s(h.kind, s[a][b]) : 0.0;
      Ks[(c0 + ty * 4 + a) * np + j] = v;
      macc = fma(v, alpha[j], macc);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) macc += __shfl_xor_sync(0xffffffffu, macc, o);
    if (tx == 0) mupart[int64_t(blockIdx.x) * kRefineRows + c0 + ty * 4 + a] = macc;
  }
}

// part[it, c] = sum over the 32 rows i of block `it` of (sum_{j <= i} Ks[c,j] Linv[i,j])^2.
// grid (np/32, rows/64), Cfg64x32: many small CTAs so that a 64..256-candidate short-list still uses the chip.
__global__ void __launch_bounds__(Cfg64x32::THREADS)
k_refine_vnorm(const double* __restrict__ Ks, const double* __restrict__ Linv, int64_t np, double* __restrict__ part) {
  using Cfg = Cfg64x32;
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[Cfg::BM][Cfg::WARPS_N];
  const int it = blockIdx.x;
  const int64_t c0 = int64_t(blockIdx.y) * Cfg::BM;
  Acc<Cfg> acc;
  acc.zero();
  gemm_nt_mainloop<Cfg>(Ks + c0 * np, np, Linv + int64_t(it) * Cfg::BN * np, np, 0, (it + 1) * Cfg::BN, acc, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) {
      v = fma(acc.v[i][j][0], acc.v[i][j][0], v);
      v = fma(acc.v[i][j][1], acc.v[i][j][1], v);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) red[wm * Cfg::WM + i * 8 + g][wn] = v;
  }
  __syncthreads();
  if (threadIdx.x < Cfg::BM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < Cfg::WARPS_N; ++w) v += red[threadIdx.x][w];
    part[int64_t(it) * kRefineRows + c0 + threadIdx.x] = v;
  }
}

// float64 acquisition of the re-scored short-list + the selection list [running top-k | short-list] + evidence.
__global__ void __launch_bounds__(kRefineRows)
k_refine_acq(bbo_gp_hyper h, bbo_acq a, const double* __restrict__ hyp, const double* __restrict__ mupart, int nmu,
             const double* __restrict__ part, int nblk, const int64_t* __restrict__ gi,
             const double* __restrict__ tf_s, int m, int k, int keep_running, const double* __restrict__ run_s,
             const int64_t* __restrict__ run_i, double* __restrict__ la_s, int64_t* __restrict__ la_i,
             double* __restrict__ info, int32_t* __restrict__ nan_count) {
  __shared__ double red[kRefineRows];
  __shared__ int cnt[kRefineRows];
  const int c = threadIdx.x;
  const double eps = h.kind == BBO_KERNEL_MATERN52 ? h.pairwise_eps : 0.0;
  const double kxx = hyp[2 * h.d] * kernel_from_s(h.kind, double(h.d) * eps * eps);
  double err = 0.0;
  int valid = 0;
  if (c < m) {
    double vn = 0.0;
    for (int b = 0; b < nblk; ++b) vn += part[int64_t(b) * kRefineRows + c];
    const double var = kxx - vn + h.noise;
    double mu = 0.0;
    for (int b = 0; b < nmu; ++b) mu += mupart[int64_t(b) * kRefineRows + c];
    mu += hyp[2 * h.d + 1];
    double sc = acq_value<double>(a, mu, sqrt(var));
    int64_t ix = gi[c];
    if (ix >= 0 && sc != sc) {
      ix = -1;
      if (nan_count) atomicAdd(nan_count, 1);
    }
    if (ix < 0) sc = -INFINITY;
    else {
      valid = 1;
      const double e = fabs(sc - tf_s[c]);
      if (e == e && e < INFINITY) err = e;
    }
    la_s[k + c] = sc;
    la_i[k + c] = ix;
  }
  if (c < k) {
    la_s[c] = keep_running ? run_s[c] : -INFINITY;
    la_i[c] = keep_running ? run_i[c] : -1;
  }
  red[c] = err;
  cnt[c] = valid;
  __syncthreads();
  for (int o = kRefineRows / 2; o > 0; o >>= 1) {
    if (c < o) {
      red[c] = fmax(red[c], red[c + o]);
      cnt[c] += cnt[c + o];
    }
    __syncthreads();
  }
  if (c == 0) info[3] = double(cnt[0]);     // (the error entries are k_refine_info's)
}

// Evidence (see bbo_gp_refine_info).  The observed TF32 score error is taken over the short-listed candidates that
// did NOT make the returned k -- the ones nearest the cut: for EI / PI a score error scales with the score, and the
// largest scores of the list (orders of magnitude above the k-th one in a large pool) say nothing about candidates
// at the cut.  info[4] keeps the maximum over the whole list.
__global__ void __launch_bounds__(kRefineRows)
k_refine_info(const double* __restrict__ topk_s, int k, const double* __restrict__ sh_s,
              const int64_t* __restrict__ sh_i, int m, const double* __restrict__ bound,
              const double* __restrict__ f64_s, const int64_t* __restrict__ f64_i, const double* __restrict__ tf_s,
              double* __restrict__ info) {
  __shared__ double red_all[kRefineRows], red_cut[kRefineRows];
  __shared__ int cnt_cut[kRefineRows];
  const int c = threadIdx.x;
  const double kth = topk_s[k - 1];
  double e_all = 0.0, e_cut = 0.0;
  int n_cut = 0;
  if (c < m && f64_i[c] >= 0) {
    const double e = fabs(f64_s[c] - tf_s[c]);
    if (e == e && e < INFINITY) {
      e_all = e;
      if (f64_s[c] < kth) {
        e_cut = e;
        n_cut = 1;
      }
    }
  }
  red_all[c] = e_all;
  red_cut[c] = e_cut;
  cnt_cut[c] = n_cut;
  __syncthreads();
  for (int o = kRefineRows / 2; o > 0; o >>= 1) {
    if (c < o) {
      red_all[c] = fmax(red_all[c], red_all[c + o]);
      red_cut[c] = fmax(red_cut[c], red_cut[c + o]);
      cnt_cut[c] += cnt_cut[c + o];
    }
    __syncthreads();
  }
  if (c == 0) {
    info[0] = kth;
    double b = *bound;
    if (sh_i[m - 1] >= 0) b = fmax(b, sh_s[m - 1]);   // full short-list: whatever it displaced scored <= its last entry
    info[1] = b;
    info[2] = cnt_cut[0] > 0 ? red_cut[0] : red_all[0];
    info[4] = red_all[0];
    info[5] = double(cnt_cut[0]);
  }
}

static DeviceOnce g_attr_once;
static int set_smem_attrs() {
  if (!g_attr_once.first()) return BBO_OK;
  BBO_CUDA_CHECK(cudaFuncSetAttribute(k_vnorm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(Cfg128x64::SMEM_BYTES)));
  BBO_CUDA_CHECK(cudaFuncSetAttribute(k_vstore, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(Cfg64x64::SMEM_BYTES)));
  BBO_CUDA_CHECK(cudaFuncSetAttribute(k_cov_sub, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(Cfg64x64::SMEM_BYTES)));
  return BBO_OK;
}

// mu and |v|^2 partials for one chunk (FP64 or TF32)
static int posterior_chunk(const bbo_gp_state& s, const ScoreWorkspace& w, int precision, const void* Xq,
                           int xq_is_f32, int64_t qc, int* nsplit_out, cudaStream_t st) {
  const int64_t qc_pad = ceil_div(qc, 128) * 128;
  if (precision != BBO_PREC_FP64) return BBO_ERR_BAD_ARG;
  const int nsplit = pick_nsplit(qc_pad, w.np);
  *nsplit_out = nsplit;
  prof_begin(kProfKstar, st);
  if (xq_is_f32)
    k_kstar<float><<<unsigned(qc_pad / 64), 256, 0, st>>>(s.h, static_cast<const float*>(Xq), qc, s.X, s.n, w.np,
                                                         s.hyp, s.alpha, w.Ks, w.mu, w.qp);
  else
    k_kstar<double><<<unsigned(qc_pad / 64), 256, 0, st>>>(s.h, static_cast<const double*>(Xq), qc, s.X, s.n,
                                                          w.np, s.hyp, s.alpha, w.Ks, w.mu, w.qp);
  BBO_LAUNCH_CHECK();
  prof_end(kProfKstar, st);
  prof_begin(kProfVnorm, st);
  k_vnorm<<<dim3(unsigned(qc_pad / 128), unsigned(nsplit)), Cfg128x64::THREADS, Cfg128x64::SMEM_BYTES, st>>>(
      w.Ks, s.linv, w.np, w.qp, w.vpart);
  BBO_LAUNCH_CHECK();
  prof_end(kProfVnorm, st);
  return BBO_OK;
}

int gp_predict(const bbo_gp_state& s, const double* Xq, int64_t q, double* mu, double* var, double* cov_full,
               void* ws, size_t ws_bytes, int64_t q_chunk, cudaStream_t st) {
  if (q < 0 || s.n < 1 || s.h.d < 1) return BBO_ERR_BAD_SHAPE;
  if (q == 0) return BBO_OK;
  int rc = set_smem_attrs();
  if (rc != BBO_OK) return rc;
  const bool full = cov_full != nullptr;
  ScoreWorkspace w;
  rc = w.bind(ws, ws_bytes, s.n, s.h.d, q_chunk, BBO_PREC_FP64, full);
  if (rc != BBO_OK) return rc;
  if (full && q > w.qp) return BBO_ERR_BAD_SHAPE;
  bbo_acq none{};
  for (int64_t c0 = 0; c0 < q; c0 += w.qp) {
    const int64_t qc = (q - c0 < w.qp) ? q - c0 : w.qp;
    int nsplit = 1;
    rc = posterior_chunk(s, w, BBO_PREC_FP64, Xq + c0 * s.h.d, 0, qc, &nsplit, st);
    if (rc != BBO_OK) return rc;
    k_acq<<<unsigned(ceil_div(qc, 256)), 256, 0, st>>>(s.h, none, 0, 0, s.hyp, w.mu, w.vpart, nsplit, w.qp, qc,
                                                       w.score, mu ? mu + c0 : nullptr, var ? var + c0 : nullptr,
                                                       nullptr, nullptr);
    BBO_LAUNCH_CHECK();
    if (full) {
      // gp.py:81-82: cov(Xq,Xq) - v^T v + noise I
      const int64_t qc_pad = ceil_div(qc, 128) * 128;
      k_vstore<<<dim3(unsigned(qc_pad / 64), unsigned(w.np / 64)), Cfg64x64::THREADS, Cfg64x64::SMEM_BYTES, st>>>(
          w.Ks, s.linv, w.np, w.V);
      BBO_LAUNCH_CHECK();
      rc = cov_matrix(s.h, s.hyp, Xq, q, Xq, q, cov_full, q, s.h.noise, st);
      if (rc != BBO_OK) return rc;
      k_cov_sub<<<dim3(unsigned(ceil_div(q, 64)), unsigned(ceil_div(q, 64))), Cfg64x64::THREADS,
                  Cfg64x64::SMEM_BYTES, st>>>(w.V, w.np, cov_full, q);
      BBO_LAUNCH_CHECK();
    }
  }
  return BBO_OK;
}

// float64 re-score of the gathered short-list rows (r.Xg[par], r.gi[par], r.tf_s) and the final selection.
static int refine_rescore(const bbo_gp_state& s, const bbo_acq& acq, const ScoreWorkspace& w, const RefineScratch& r,
                          int par, int m, int k, double* topk_s, int64_t* topk_i, bool keep_running,
                          int32_t* nan_count, cudaStream_t st) {
  const int64_t np = w.np;
  const int rows = int(ceil_div(m, 64) * 64);
  prof_begin(kProfKstar, st);
  k_refine_kstar<<<dim3(unsigned(np / 64), unsigned(rows / 64)), 256, 0, st>>>(s.h, r.Xg[par], rows, s.X, s.n, np,
                                                                              s.hyp, s.alpha, r.Ks, r.mu);
  BBO_LAUNCH_CHECK();
  prof_end(kProfKstar, st);
  prof_begin(kProfVnorm, st);
  k_refine_vnorm<<<dim3(unsigned(np / 32), unsigned(rows / 64)), Cfg64x32::THREADS, Cfg64x32::SMEM_BYTES, st>>>(
      r.Ks, s.linv, np, r.part);
  BBO_LAUNCH_CHECK();
  prof_end(kProfVnorm, st);
  prof_begin(kProfSelect, st);
  k_refine_acq<<<1, kRefineRows, 0, st>>>(s.h, acq, s.hyp, r.mu, int(np / 64), r.part, int(np / 32), r.gi[par], r.tf_s, m, k,
                                          keep_running ? 1 : 0, topk_s, topk_i, r.la_s, r.la_i, r.info, nan_count);
  BBO_LAUNCH_CHECK();
  k_topk_seg<<<1, 256, 0, st>>>(r.la_s, r.la_i, 0, int64_t(k) + m, k, topk_s, topk_i);
  BBO_LAUNCH_CHECK();
  k_refine_info<<<1, kRefineRows, 0, st>>>(topk_s, k, r.sh_s, r.sh_i, m, r.bound, r.la_s + k, r.la_i + k, r.tf_s,
                                           r.info);
  BBO_LAUNCH_CHECK();
  prof_end(kProfSelect, st);
  return BBO_OK;
}

int refine_info(const void* ws, size_t ws_bytes, int64_t n, int64_t d, int64_t q_chunk, double* info_host,
                cudaStream_t st) {
  ScoreWorkspace w;
  int rc = w.bind(const_cast<void*>(ws), ws_bytes, n, d, q_chunk, BBO_PREC_TF32_REFINE, false);
  if (rc != BBO_OK) return rc;
  const RefineScratch r = carve_refine(w.refine, w.np, d);
  BBO_CUDA_CHECK(cudaMemcpyAsync(info_host, r.info, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  BBO_CUDA_CHECK(cudaStreamSynchronize(st));
  return BBO_OK;
}

int gp_score_pool(const bbo_gp_state& s, const bbo_acq& acq, int precision, const void* Xq, int xq_is_f32,
                  int64_t q, int64_t index_offset, int k, double* topk_s, int64_t* topk_i, double* mu_out,
                  double* var_out, double* score_out, int32_t* nan_count, void* ws, size_t ws_bytes,
                  int64_t q_chunk, bool keep_running, cudaStream_t st, int refine_flags) {
  if (q < 0 || s.n < 1 || s.h.d < 1) return BBO_ERR_BAD_SHAPE;
  if (k < 1 || k > BBO_TOPK_MAX) return BBO_ERR_BAD_ARG;
  if (precision != BBO_PREC_FP64 && precision != BBO_PREC_TF32 && precision != BBO_PREC_TF32_REFINE)
    return BBO_ERR_BAD_ARG;
  const bool tf32 = precision != BBO_PREC_FP64;
  const bool refine = precision == BBO_PREC_TF32_REFINE;
  if (tf32 && s.linv_tf32 == nullptr) return BBO_ERR_BAD_ARG;
  if (refine && k > BBO_REFINE_KMAX) return BBO_ERR_BAD_ARG;
  int rc = set_smem_attrs();
  if (rc != BBO_OK) return rc;
  ScoreWorkspace w;
  rc = w.bind(ws, ws_bytes, s.n, s.h.d, q_chunk, precision, false);
  if (rc != BBO_OK) return rc;
  BBO_CUDA_CHECK(cudaMemsetAsync(w.surv_count, 0, (64 + size_t(w.qp) / 256 + 1) * sizeof(int), st));
  RefineScratch r{};
  const int m = refine ? refine_shortlist_len(k) : 0, kseg = refine ? refine_kseg(k) : 0;
  if (refine) {
    r = carve_refine(w.refine, w.np, s.h.d);
    if (!(refine_flags & kRefineContinue)) {
      k_refine_init<<<1, kRefineRows, 0, st>>>(r.sh_s, r.sh_i, r.bound, r.gi[0], r.gi[1], kRefineRows);
      BBO_LAUNCH_CHECK();
    }
  } else if (!keep_running) {
    k_topk_init<<<1, BBO_TOPK_MAX, 0, st>>>(topk_s, topk_i, k);
    BBO_LAUNCH_CHECK();
  }
  // acquisition + running top-k of one finished chunk (mu and |v|^2 partials are ready on st)
  auto consume = [&](int64_t c0, int64_t qc, const double* mu_chunk, int nsplit, const double* vpart,
                     cudaStream_t sc) -> int {
    prof_begin(kProfSelect, sc);
    const int64_t nb = ceil_div(qc, kSeg);
    static const bool unfused = getenv("BBO_SELECT_UNFUSED") != nullptr;
    static const bool rounds_only = getenv("BBO_SELECT_ROUNDS") != nullptr;   // round-based selection everywhere
    // running list of this mode: the k best (score, index) pairs, or the tf32+refine short-list (m entries)
    const int K = refine ? m : k;
    double* run_s = refine ? r.sh_s : topk_s;
    int64_t* run_i = refine ? r.sh_i : topk_i;
    if (!unfused && !rounds_only && nb * int64_t(K + kFilterSlack) <= w.list_len) {
      // threshold filter (default): only candidates that beat the current last entry of the running list survive
      k_acq_eval<<<unsigned(ceil_div(qc, kEvalBlock)), kEvalThreads, 0, sc>>>(
          s.h, acq, tf32, s.hyp, mu_chunk, vpart, nsplit, w.qp, qc, index_offset + c0, K, run_s, run_i,
          mu_out ? mu_out + c0 : nullptr, var_out ? var_out + c0 : nullptr, score_out ? score_out + c0 : nullptr,
          nan_count, w.la_s, w.la_i, w.surv_count, w.score, w.surv_count + 64);
      BBO_LAUNCH_CHECK();
      k_filter_heavy<<<unsigned(nb), 256, 0, sc>>>(qc, index_offset + c0, K, w.score, w.surv_count + 64, w.la_s,
                                                  w.la_i, w.surv_count);
      BBO_LAUNCH_CHECK();
      k_merge_level<<<unsigned(ceil_div(nb * int64_t(K + kFilterSlack), kMergeCap)), 256, 0, sc>>>(
          w.la_s, w.la_i, w.surv_count, K, w.lb_s, w.lb_i);
      BBO_LAUNCH_CHECK();
      k_merge_survivors<<<1, 256, 0, sc>>>(w.la_s, w.la_i, w.surv_count, K, run_s, run_i, w.lb_s, w.lb_i);
      BBO_LAUNCH_CHECK();
    } else if (refine) {
      // round-based fallback (BBO_SELECT_ROUNDS=1 or a chunk too large for the survivor buffer): TF32 scores ->
      // per-segment lists of kseg = 2k entries -> running short-list of m entries (by TF32 score); what a segment
      // dropped is bounded by its kseg-th entry (k_refine_bound)
      if (nb * kseg + m > w.list_len) return BBO_ERR_WORKSPACE;
      k_acq_topk<<<unsigned(nb), 256, 0, sc>>>(
          s.h, acq, 1, s.hyp, mu_chunk, vpart, nsplit, w.qp, qc, index_offset + c0, kseg,
          mu_out ? mu_out + c0 : nullptr, var_out ? var_out + c0 : nullptr, score_out ? score_out + c0 : nullptr,
          nan_count, w.la_s + m, w.la_i + m, nullptr);
      BBO_LAUNCH_CHECK();
      k_refine_bound<<<1, 256, 0, sc>>>(w.la_s + m, w.la_i + m, int(nb), kseg, r.bound);
      BBO_LAUNCH_CHECK();
      if (nb * kseg + m <= kSeg) {
        k_topk_merge_running<<<1, 256, 0, sc>>>(w.la_s + m, w.la_i + m, int(nb * kseg), m, r.sh_s, r.sh_i);
        BBO_LAUNCH_CHECK();
      } else {   // large chunks: the running short-list goes in front of the segment lists, reduced in levels
        BBO_CUDA_CHECK(cudaMemcpyAsync(w.la_s, r.sh_s, size_t(m) * sizeof(double), cudaMemcpyDeviceToDevice, sc));
        BBO_CUDA_CHECK(cudaMemcpyAsync(w.la_i, r.sh_i, size_t(m) * sizeof(int64_t), cudaMemcpyDeviceToDevice, sc));
        int rc2 = topk_levels(w.la_s, w.la_i, nb * kseg + m, w.lb_s, w.lb_i, m, r.sh_s, r.sh_i, sc);
        if (rc2 != BBO_OK) return rc2;
      }
    } else if (!unfused && (nb + 1) * k <= kSeg) {
      // fused acquisition + per-segment top-k, then one merge with the running list (2 launches per chunk)
      k_acq_topk<<<unsigned(nb), 256, 0, sc>>>(
          s.h, acq, tf32, s.hyp, mu_chunk, vpart, nsplit, w.qp, qc, index_offset + c0, k,
          mu_out ? mu_out + c0 : nullptr, var_out ? var_out + c0 : nullptr, score_out ? score_out + c0 : nullptr,
          nan_count, w.la_s, w.la_i, nullptr);
      BBO_LAUNCH_CHECK();
      k_topk_merge_running<<<1, 256, 0, sc>>>(w.la_s, w.la_i, int(nb * k), k, topk_s, topk_i);
      BBO_LAUNCH_CHECK();
    } else {
      k_acq<<<unsigned(ceil_div(qc, 256)), 256, 0, sc>>>(
          s.h, acq, 1, tf32, s.hyp, mu_chunk, vpart, nsplit, w.qp, qc, w.score,
          mu_out ? mu_out + c0 : nullptr, var_out ? var_out + c0 : nullptr, score_out ? score_out + c0 : nullptr,
          nan_count);
      BBO_LAUNCH_CHECK();
      // level 1: segments of this chunk -> list A (after the k running entries)
      BBO_CUDA_CHECK(cudaMemcpyAsync(w.la_s, topk_s, size_t(k) * sizeof(double), cudaMemcpyDeviceToDevice, sc));
      BBO_CUDA_CHECK(cudaMemcpyAsync(w.la_i, topk_i, size_t(k) * sizeof(int64_t), cudaMemcpyDeviceToDevice, sc));
      k_topk_seg<<<unsigned(nb), 256, 0, sc>>>(w.score, nullptr, index_offset + c0, qc, k, w.la_s + k, w.la_i + k);
      BBO_LAUNCH_CHECK();
      int rc2 = topk_levels(w.la_s, w.la_i, (nb + 1) * k, w.lb_s, w.lb_i, k, topk_s, topk_i, sc);
      if (rc2 != BBO_OK) return rc2;
    }
    prof_end(kProfSelect, sc);
    return BBO_OK;
  };
  if (tf32) {
    if (q > 0) {
      // the selection starts from an empty list unless this call continues a running list / short-list
      const bool fresh = refine ? (refine_flags & kRefineContinue) == 0 : !keep_running;
      rc = score_tf32_pipeline(s, w, Xq, xq_is_f32, q, consume, st, /*short_first=*/fresh);
      if (rc != BBO_OK) return rc;
    }
    if (!refine) return BBO_OK;
    // rows of the short-list -> the row buffer of this call's parity (the other one holds the previous call's)
    const int par = (refine_flags & kRefineParity) ? 1 : 0;
    const bool cont = (refine_flags & kRefineContinue) != 0;
    const int grows = int(ceil_div(m, 64) * 64);
    if (xq_is_f32)
      k_refine_gather<float><<<1, 256, 0, st>>>(static_cast<const float*>(Xq), index_offset, q, s.h.d, r.sh_s, r.sh_i,
                                                m, grows, cont ? r.Xg[par ^ 1] : nullptr,
                                                cont ? r.gi[par ^ 1] : nullptr, r.Xg[par], r.gi[par], r.tf_s);
    else
      k_refine_gather<double><<<1, 256, 0, st>>>(static_cast<const double*>(Xq), index_offset, q, s.h.d, r.sh_s,
                                                 r.sh_i, m, grows, cont ? r.Xg[par ^ 1] : nullptr,
                                                 cont ? r.gi[par ^ 1] : nullptr, r.Xg[par], r.gi[par], r.tf_s);
    BBO_LAUNCH_CHECK();
    if (refine_flags & kRefineDefer) return BBO_OK;
    return refine_rescore(s, acq, w, r, par, m, k, topk_s, topk_i, keep_running, nan_count, st);
  }
  const size_t esz = xq_is_f32 ? sizeof(float) : sizeof(double);
  for (int64_t c0 = 0; c0 < q; c0 += w.qp) {
    const int64_t qc = (q - c0 < w.qp) ? q - c0 : w.qp;
    const void* xq = static_cast<const char*>(Xq) + size_t(c0) * s.h.d * esz;
    int nsplit = 1;
    rc = posterior_chunk(s, w, precision, xq, xq_is_f32, qc, &nsplit, st);
    if (rc != BBO_OK) return rc;
    rc = consume(c0, qc, w.mu, nsplit, w.vpart, st);
    if (rc != BBO_OK) return rc;
  }
  return BBO_OK;
}

// ------------------------------------------------------------------------------------------
// Batched FP64 scoring: B independent problems of one shape (the state of a batched bbo_gp_factorize), one
// pool of q candidates per problem, per-problem acquisition parameters.  Same kernels as the single-problem
// path with blockIdx.z = problem; problems are processed in groups that keep K* under ~512 MB.
// ------------------------------------------------------------------------------------------
static int64_t batched_group(int64_t n, int64_t B, int64_t q) {
  const int64_t np = padded(n), qp = round_chunk(q);
  int64_t g = (int64_t(512) << 20) / (qp * np * 8);
  if (g < 1) g = 1;
  if (g > B) g = B;
  if (g > 65535) g = 65535;
  return g;
}
size_t score_batched_workspace_bytes(int64_t n, int64_t d, int64_t B, int64_t q) {
  const int64_t np = padded(n), qp = round_chunk(q), g = batched_group(n, B, q);
  const int64_t nb = ceil_div(qp, kSeg);
  return size_t(g) * (size_t(qp) * np + size_t(kMaxSplit) * qp + qp + 2 * size_t(nb) * BBO_TOPK_MAX) * 8 + 1024;
}
int gp_score_pool_batched(const bbo_gp_hyper& h, int64_t n, int64_t B, const double* X, const double* hyp,
                          const double* linv, const double* alpha, const bbo_acq* acq_dev, const double* Xq,
                          int64_t q, int k, double* topk_s, int64_t* topk_i, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
  if (n < 1 || B < 1 || q < 1 || h.d < 1) return BBO_ERR_BAD_SHAPE;
  if (k < 1 || k > BBO_TOPK_MAX) return BBO_ERR_BAD_ARG;
  if (ws_bytes < score_batched_workspace_bytes(n, h.d, B, q)) return BBO_ERR_WORKSPACE;
  int rc = set_smem_attrs();
  if (rc != BBO_OK) return rc;
  const int64_t np = padded(n), qp = round_chunk(q), g = batched_group(n, B, q);
  const int64_t nb = ceil_div(q, kSeg);
  if ((nb + 1) * k > kSeg) return BBO_ERR_BAD_ARG;       // pool too large for the single-merge selection
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  double* Ks = reinterpret_cast<double*>(p);
  double* vpart = Ks + size_t(g) * qp * np;
  double* mu = vpart + size_t(g) * kMaxSplit * qp;
  double* la_s = mu + size_t(g) * qp;
  int64_t* la_i = reinterpret_cast<int64_t*>(la_s + size_t(g) * ceil_div(qp, kSeg) * BBO_TOPK_MAX);
  const int nsplit = 1;                                   // B problems already fill the machine
  k_topk_init<<<unsigned(ceil_div(B * k, 256)), 256, 0, st>>>(topk_s, topk_i, int(B * k));
  BBO_LAUNCH_CHECK();
  for (int64_t b0 = 0; b0 < B; b0 += g) {
    const unsigned gb = unsigned((B - b0 < g) ? B - b0 : g);
    k_kstar<double><<<dim3(unsigned(qp / 64), 1, gb), 256, 0, st>>>(
        h, Xq + b0 * q * h.d, q, X + b0 * n * h.d, n, np, hyp + b0 * (2 * h.d + 2), alpha + b0 * np, Ks, mu, qp);
    BBO_LAUNCH_CHECK();
    k_vnorm<<<dim3(unsigned(qp / 128), unsigned(nsplit), gb), Cfg128x64::THREADS, Cfg128x64::SMEM_BYTES, st>>>(
        Ks, linv + b0 * np * np, np, qp, vpart);
    BBO_LAUNCH_CHECK();
    k_acq_topk<<<dim3(unsigned(nb), 1, gb), 256, 0, st>>>(h, bbo_acq{}, 0, hyp + b0 * (2 * h.d + 2), mu, vpart,
                                                         nsplit, qp, q, 0, k, nullptr, nullptr, nullptr, nullptr,
                                                         la_s, la_i, acq_dev + b0);
    BBO_LAUNCH_CHECK();
    k_topk_merge_running<<<dim3(1, 1, gb), 256, 0, st>>>(la_s, la_i, int(nb * k), k, topk_s + b0 * k,
                                                        topk_i + b0 * k);
    BBO_LAUNCH_CHECK();
  }
  return BBO_OK;
}

int acq_eval(const bbo_acq& a, const void* mu, const void* sd, void* out, int64_t n, int is_f32,
             cudaStream_t st) {
  if (n <= 0) return BBO_OK;
  if (is_f32)
    k_acq_elem<float><<<unsigned(ceil_div(n, 256)), 256, 0, st>>>(a, static_cast<const float*>(mu),
                                                                 static_cast<const float*>(sd),
                                                                 static_cast<float*>(out), n);
  else
    k_acq_elem<double><<<unsigned(ceil_div(n, 256)), 256, 0, st>>>(a, static_cast<const double*>(mu),
                                                                  static_cast<const double*>(sd),
                                                                  static_cast<double*>(out), n);
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

}  // namespace bbo

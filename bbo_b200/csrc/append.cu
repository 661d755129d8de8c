// Incremental update of a fitted GP state when observations are appended.
//
// The reference refits from scratch on every suggestion (designers/bo/bo.py:101-106: a new surrogate,
// 100 Adam epochs of O(n^3)).  With the hyper-parameters kept, appending m observations only needs
//
//   K' = [K  .]      L' = [L   0 ]     L'^-1 = [ L^-1              0     ]
//        [B  C]           [S   Lc]             [-Lc^-1 S L^-1     Lc^-1  ]
//
//   B  = k(X_new, X_old)            (m x n)      cov_matrix (k_cov_rect)
//   S  = B L^-T                     (m x n)      k_rows_dot     one pass over L^-1
//   Lc = chol(C + noise I - S S^T)  (m x m)      k_append_gram + k_append_schur (one CTA)
//   R  = -Lc^-1 (S L^-1)            (m x n)      k_cols_dot_partial / k_cols_finish: one pass over L^-1
//   z  = L'^-1 (y - c),  alpha = L'^-T z         the same two kernels with one vector
//
// i.e. four passes over the n^2/2 doubles of L^-1: HBM bound, O(m n^2) flops instead of O(n^3).
// Only the lower triangle of K' is used (rows of the new points against everything before them), the
// same convention as the full factorisation (the eps-Matern K is not symmetric).
// All kernels are deterministic (fixed reduction order).
#include "fit.cuh"

namespace bbo {

constexpr int kAppMax = BBO_APPEND_MAX;   // new rows per call
constexpr int kAppMB = 8;                 // vectors per register pass
constexpr int kAppRC = 512;               // rows of L^-1 per partial-sum chunk

// out[i][r] = sum_{k <= min(r, kcap-1)} V[i][k] * W[r][k]  for r in [0, nrows).
// One warp per FOUR consecutive rows (W is zero above its diagonal, so all four contract k <= r+3): every
// V value loaded serves four rows, which keeps the V traffic (m vectors, L2 resident) at 2 m/4 of the W bytes.
constexpr int kAppRW = 4;
__global__ void __launch_bounds__(256)
k_rows_dot(const double* __restrict__ W, int64_t ld, int64_t nrows, int64_t kcap, const double* __restrict__ V,
           int64_t ldv, int m, double* __restrict__ out, int64_t ldo) {
  const int64_t r = (int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5)) * kAppRW;
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const int64_t rl = (r + kAppRW <= nrows) ? r + kAppRW : nrows;      // rows [r, rl)
  const int64_t kend = (rl < kcap) ? rl : kcap;
  const double* w = W + r * ld;
  for (int i0 = 0; i0 < m; i0 += kAppMB) {
    double acc[kAppRW][kAppMB];
#pragma unroll
    for (int u = 0; u < kAppRW; ++u)
#pragma unroll
      for (int i = 0; i < kAppMB; ++i) acc[u][i] = 0.0;
#pragma unroll 2
    for (int64_t k = lane; k < kend; k += 32) {
      double wv[kAppRW], vv[kAppMB];
#pragma unroll
      for (int u = 0; u < kAppRW; ++u) wv[u] = (r + u < rl) ? w[u * ld + k] : 0.0;
#pragma unroll
      for (int i = 0; i < kAppMB; ++i) vv[i] = (i0 + i < m) ? V[int64_t(i0 + i) * ldv + k] : 0.0;
#pragma unroll
      for (int u = 0; u < kAppRW; ++u)
#pragma unroll
        for (int i = 0; i < kAppMB; ++i) acc[u][i] = fma(vv[i], wv[u], acc[u][i]);
    }
#pragma unroll
    for (int u = 0; u < kAppRW; ++u)
#pragma unroll
      for (int i = 0; i < kAppMB; ++i) {
        double a = acc[u][i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && i0 + i < m && r + u < rl) out[int64_t(i0 + i) * ldo + r + u] = a;
      }
  }
}

// P[rc][i][j] = sum_{r in chunk rc, r >= j, r < nrows} Sv[i][r] * W[r][j].
// grid (ceil(ncols/128), ceil(nrows/kAppRC)), 512 threads = 4 row groups x 128 columns: thread (g, c) walks
// rows r0 + g, r0 + g + 4, ... of column jb + c with eight independent loads in flight; the four groups
// are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(512)
k_cols_dot_partial(const double* __restrict__ W, int64_t ld, int64_t nrows, int64_t ncols,
                   const double* __restrict__ Sv, int64_t lds, int m, double* __restrict__ P, int64_t ldp) {
  __shared__ double red[3][kAppMB][128];
  const int c = threadIdx.x & 127, g = threadIdx.x >> 7;
  const int64_t jb = int64_t(blockIdx.x) * 128, j = jb + c;
  const int64_t r0 = int64_t(blockIdx.y) * kAppRC;
  const int64_t r1 = (r0 + kAppRC < nrows) ? r0 + kAppRC : nrows;
  const int64_t rs = (r0 > jb ? r0 : jb) + g;      // rows above the CTA's first column hold only zeros
  const bool live = j < ncols;
  for (int i0 = 0; i0 < m; i0 += kAppMB) {
    double acc[kAppMB];
#pragma unroll
    for (int i = 0; i < kAppMB; ++i) acc[i] = 0.0;
    if (live) {
      int64_t r = rs;
      for (; r + 28 < r1; r += 32) {
        double wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = W[(r + 4 * u) * ld + j];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int i = 0; i < kAppMB; ++i)
            if (i0 + i < m) acc[i] = fma(Sv[int64_t(i0 + i) * lds + r + 4 * u], wv[u], acc[i]);
      }
      for (; r < r1; r += 4) {
        const double wv = W[r * ld + j];
#pragma unroll
        for (int i = 0; i < kAppMB; ++i)
          if (i0 + i < m) acc[i] = fma(Sv[int64_t(i0 + i) * lds + r], wv, acc[i]);
      }
    }
    __syncthreads();
    if (g > 0) {
#pragma unroll
      for (int i = 0; i < kAppMB; ++i) red[g - 1][i][c] = acc[i];
    }
    __syncthreads();
    if (g == 0 && live) {
#pragma unroll
      for (int i = 0; i < kAppMB; ++i)
        if (i0 + i < m)
          P[(int64_t(blockIdx.y) * m + i0 + i) * ldp + j] = ((acc[i] + red[0][i][c]) + red[1][i][c]) + red[2][i][c];
    }
  }
}

// T[i][j] = sum_rc P[rc][i][j];  mode 0: out[(row0+i)*ldo + j] = -sum_{i' <= i} Lci[i][i'] T[i'][j]
//                                mode 1: out[j] = T[0][j]
__global__ void __launch_bounds__(128)
k_cols_finish(const double* __restrict__ P, int64_t ldp, int nchunk, int m, int64_t ncols,
              const double* __restrict__ Lci, int mode, double* __restrict__ out, int64_t ldo, int64_t row0) {
  __shared__ double lc[kAppMax * kAppMax];
  if (mode == 0)
    for (int e = threadIdx.x; e < m * m; e += 128) lc[e] = Lci[e];
  __syncthreads();
  const int64_t j = int64_t(blockIdx.x) * 128 + threadIdx.x;
  if (j >= ncols) return;
  double t[kAppMax];
#pragma unroll
  for (int i = 0; i < kAppMax; ++i) {
    double a = 0.0;
    if (i < m)
      for (int rc = 0; rc < nchunk; ++rc) a += P[(int64_t(rc) * m + i) * ldp + j];
    t[i] = a;
  }
  if (mode == 1) {
    out[j] = t[0];
    return;
  }
#pragma unroll
  for (int i = 0; i < kAppMax; ++i) {
    if (i < m) {
      double a = 0.0;
#pragma unroll
      for (int ip = 0; ip < kAppMax; ++ip)
        if (ip <= i) a = fma(lc[i * m + ip], t[ip], a);
      out[(row0 + i) * ldo + j] = -a;
    }
  }
}

// Gram[i][j] = sum_r S[i][r] S[j][r] (j <= i): one CTA per pair, fixed-order block reduction.
__global__ void __launch_bounds__(256)
k_append_gram(const double* __restrict__ S, int64_t lds, int64_t n, int m, double* __restrict__ Gram) {
  __shared__ double sh[8];
  const int i = blockIdx.x / m, j = blockIdx.x % m;
  if (j > i) return;
  double a0 = 0.0, a1 = 0.0;
  int64_t r = threadIdx.x;
  for (; r + 256 < n; r += 512) {
    a0 = fma(S[int64_t(i) * lds + r], S[int64_t(j) * lds + r], a0);
    a1 = fma(S[int64_t(i) * lds + r + 256], S[int64_t(j) * lds + r + 256], a1);
  }
  if (r < n) a0 = fma(S[int64_t(i) * lds + r], S[int64_t(j) * lds + r], a0);
  double a = a0 + a1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    Gram[blockIdx.x] = t;
  }
}

// One CTA: Cs = C + noise I - S S^T (lower), Lc = chol(Cs), Lci = Lc^-1 -> Lci (m x m, row-major, zeros
// above) and the diagonal block of linv; logdet += sum log Lc_ii; info = n + j + 1 on a non-positive pivot.
__global__ void __launch_bounds__(256)
k_append_schur(const double* __restrict__ Kx, int64_t ldk, const double* __restrict__ Gram, int64_t n,
               int m, double noise, double* __restrict__ Lci, double* __restrict__ linv, int64_t ld,
               double* __restrict__ logdet, int32_t* __restrict__ info) {
  __shared__ double A[kAppMax][kAppMax + 1];
  __shared__ double Iv[kAppMax][kAppMax + 1];
  __shared__ int fail;
  const int tid = threadIdx.x;
  if (tid == 0) fail = 0;
  for (int e = tid; e < m * m; e += 256) {
    const int i = e / m, j = e % m;
    if (j <= i) A[i][j] = Kx[int64_t(i) * ldk + n + j] + (i == j ? noise : 0.0) - Gram[e];
  }
  __syncthreads();
  // unblocked right-looking Cholesky (m <= 32): thread (i) owns row i
  for (int j = 0; j < m; ++j) {
    if (tid == 0) {
      const double d = A[j][j];
      if (!(d > 0.0) && fail == 0) fail = j + 1;
      A[j][j] = sqrt(d);
    }
    __syncthreads();
    if (tid > j && tid < m) A[tid][j] /= A[j][j];
    __syncthreads();
    for (int e = tid; e < m * m; e += 256) {
      const int i = e / m, c = e % m;
      if (c > j && c <= i) A[i][c] = fma(-A[i][j], A[c][j], A[i][c]);
    }
    __syncthreads();
  }
  // inverse by forward substitution: column c of Lc^-1 by thread c
  if (tid < m) {
    const int c = tid;
    for (int i = 0; i < m; ++i) {
      double a = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) a = fma(-A[i][k], Iv[k][c], a);
      Iv[i][c] = (i >= c) ? a / A[i][i] : 0.0;
    }
  }
  __syncthreads();
  for (int e = tid; e < m * m; e += 256) {
    const int i = e / m, c = e % m;
    Lci[e] = Iv[i][c];
    linv[(n + i) * ld + n + c] = Iv[i][c];
  }
  if (tid == 0) {
    double lg = 0.0;
    for (int j = 0; j < m; ++j) lg += log(A[j][j]);
    if (logdet) *logdet += lg;
    if (fail != 0 && *info == 0) *info = int32_t(n) + fail;
  }
}

__global__ void k_resid(const double* __restrict__ y, const double* __restrict__ hyp, int d, int64_t n,
                        double* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] - hyp[2 * d + 1];
}

struct AppendWs {
  double *Kx, *S, *Lci, *Gram, *resid, *z, *P;
  int64_t ldp;
  int nchunk;
};
static size_t a256(size_t b) { return (b + 255) / 256 * 256; }
size_t append_workspace_bytes(int64_t n_old, int64_t m) {
  const int64_t nt = n_old + m, np = padded(nt);
  const int64_t nchunk = ceil_div(np, kAppRC);
  return a256(size_t(m) * nt * 8) + a256(size_t(m) * np * 8) + 2 * a256(size_t(m) * m * 8) + 2 * a256(size_t(np) * 8) +
         a256(size_t(nchunk) * m * np * 8) + 256;
}

int gp_append(const bbo_gp_hyper& h, int64_t n, int64_t m, const double* X, const double* y, const double* hyp,
              double* linv, double* alpha, double* logdet, int32_t* info, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  if (m < 1 || m > kAppMax || n < 1) return BBO_ERR_BAD_SHAPE;
  if (ws_bytes < append_workspace_bytes(n, m)) return BBO_ERR_WORKSPACE;
  const int64_t nt = n + m, np = padded(nt), d = h.d;
  AppendWs w;
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  w.Kx = reinterpret_cast<double*>(p);
  p += a256(size_t(m) * nt * 8);
  w.S = reinterpret_cast<double*>(p);
  p += a256(size_t(m) * np * 8);
  w.Lci = reinterpret_cast<double*>(p);
  p += a256(size_t(m) * m * 8);
  w.Gram = reinterpret_cast<double*>(p);
  p += a256(size_t(m) * m * 8);
  w.resid = reinterpret_cast<double*>(p);
  p += a256(size_t(np) * 8);
  w.z = reinterpret_cast<double*>(p);
  p += a256(size_t(np) * 8);
  w.P = reinterpret_cast<double*>(p);
  w.ldp = np;
  w.nchunk = int(ceil_div(np, kAppRC));

  // B and C: rows of the new points against all points (lower triangle of K')
  int rc = cov_matrix(h, hyp, X + n * d, m, X, nt, w.Kx, nt, 0.0, st);
  if (rc != BBO_OK) return rc;
  // S = B L^-T
  k_rows_dot<<<unsigned(ceil_div(n, 8 * kAppRW)), 256, 0, st>>>(linv, np, n, n, w.Kx, nt, int(m), w.S, np);
  BBO_LAUNCH_CHECK();
  k_append_gram<<<unsigned(m * m), 256, 0, st>>>(w.S, np, n, int(m), w.Gram);
  BBO_LAUNCH_CHECK();
  k_append_schur<<<1, 256, 0, st>>>(w.Kx, nt, w.Gram, n, int(m), h.noise, w.Lci, linv, np, logdet, info);
  BBO_LAUNCH_CHECK();
  // R = -Lc^-1 (S L^-1) -> rows [n, n+m) of linv, columns [0, n)
  {
    const int nch = int(ceil_div(n, kAppRC));
    k_cols_dot_partial<<<dim3(unsigned(ceil_div(n, 128)), unsigned(nch)), 512, 0, st>>>(linv, np, n, n, w.S, np,
                                                                                     int(m), w.P, w.ldp);
    BBO_LAUNCH_CHECK();
    k_cols_finish<<<unsigned(ceil_div(n, 128)), 128, 0, st>>>(w.P, w.ldp, nch, int(m), n, w.Lci, 0, linv, np, n);
    BBO_LAUNCH_CHECK();
  }
  // alpha = L'^-T L'^-1 (y - c)
  k_resid<<<unsigned(ceil_div(nt, 256)), 256, 0, st>>>(y, hyp, int(d), nt, w.resid);
  BBO_LAUNCH_CHECK();
  k_rows_dot<<<unsigned(ceil_div(nt, 8 * kAppRW)), 256, 0, st>>>(linv, np, nt, nt, w.resid, nt, 1, w.z, np);
  BBO_LAUNCH_CHECK();
  {
    const int nch = int(ceil_div(nt, kAppRC));
    k_cols_dot_partial<<<dim3(unsigned(ceil_div(nt, 128)), unsigned(nch)), 512, 0, st>>>(linv, np, nt, nt, w.z, np,
                                                                                      1, w.P, w.ldp);
    BBO_LAUNCH_CHECK();
    k_cols_finish<<<unsigned(ceil_div(nt, 128)), 128, 0, st>>>(w.P, w.ldp, nch, 1, nt, w.Lci, 1, alpha, 0, 0);
    BBO_LAUNCH_CHECK();
  }
  if (np > nt) BBO_CUDA_CHECK(cudaMemsetAsync(alpha + nt, 0, size_t(np - nt) * 8, st));
  return BBO_OK;
}

}  // namespace bbo

// Gradient of the GP objective w.r.t. the INPUTS and the TARGETS.
//
// The reference lets the covariance see warped inputs (WarpKernel: KumarWarp / MLPWarp,
// gp/kernels.py:59-68, gp/warpers.py:31-65) and lets the mean be a network (MLPMean, gp/means.py:28-35);
// autograd then differentiates objectives.py:34-45 through K(warp(X), warp(X)) and through y - mean(X).
// The O(n^3) part of that chain is d value / d Xw (n x d) and d value / d resid (n):
//
//   d value / d resid   = -alpha
//   d value / d Xw_ak   = (1 / l_k^2) [ Xw_ak * sum_j M_aj  -  sum_j M_aj Xw_jk ]
//   M_aj = W_aj s2 (h(r_aj) + h(r_ja)),  W = 1/2 (alpha alpha^T - K^-1)
//   h    = -exp(-|u|^2 / 2) (RBF)   |   -(5/3)(1 + r) e^-r (Matern-5/2; both orientations of the
//          eps-shifted distance; the O(eps^2) term eps/l * sum_j (h_ja - h_aj) is dropped)
//
// With K^-1 and alpha left in the fit workspace by fit_value_grad this is one more kernel: CTA = 32
// rows a, loop over all 64-wide column tiles j: build the M tile in shared memory, then accumulate
// M [Xw | 1] into a (32, d+1) shared accumulator (fixed order: deterministic).  FP64 CUDA-core work,
// ~8 n^2 d flops; everything else of the chain (the warp / mean modules themselves) is tiny and is
// left to autograd on top of this library (bbo_b200/gp.py).
#include "fit.cuh"

namespace bbo {

__global__ void __launch_bounds__(256)
k_grad_x(bbo_gp_hyper h, int64_t n, int64_t np, const double* __restrict__ X, const double* __restrict__ hyp,
         const double* __restrict__ alpha, const double* __restrict__ Kinv, double* __restrict__ gX) {
  extern __shared__ __align__(16) double gx_sm[];
  const int d = h.d;
  const int dp = d | 1;                     // odd row stride: conflict-free column walks
  double* xa = gx_sm;                       // 32 x dp
  double* xj = xa + 32 * dp;                // 64 x dp
  double* Ms = xj + 64 * dp;                // 32 x 65
  double* G = Ms + 32 * 65;                 // 32 x (d+1)
  double* il = G + 32 * (d + 1);            // d
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  const int64_t a0 = int64_t(blockIdx.x) * 32;
  const bool matern = h.kind == BBO_KERNEL_MATERN52;
  const double eps = matern ? h.pairwise_eps : 0.0;
  const double s2 = hyp[2 * d];
  for (int e = tid; e < 32 * d; e += 256) {
    const int r = e / d, k = e % d;
    xa[r * dp + k] = (a0 + r < n) ? X[(a0 + r) * d + k] : 0.0;
  }
  for (int e = tid; e < 32 * (d + 1); e += 256) G[e] = 0.0;
  for (int k = tid; k < d; k += 256) il[k] = hyp[d + k];
  for (int64_t j0 = 0; j0 < n; j0 += 64) {
    __syncthreads();                        // previous tile fully consumed (and the prologue stores done)
    for (int e = tid; e < 64 * d; e += 256) {
      const int r = e / d, k = e % d;
      xj[r * dp + k] = (j0 + r < n) ? X[(j0 + r) * d + k] : 0.0;
    }
    __syncthreads();
    double sf[4][2], sb[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) sf[a][0] = sf[a][1] = sb[a][0] = sb[a][1] = 0.0;
    for (int k = 0; k < d; ++k) {
      const double ik = il[k];
      const double x0 = xj[lane * dp + k], x1 = xj[(lane + 32) * dp + k];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double xv = xa[(rg * 4 + a) * dp + k];
        const double u0 = (x0 - xv) * ik, u1 = (x1 - xv) * ik;
        sf[a][0] = fma(u0 + eps, u0 + eps, sf[a][0]);
        sf[a][1] = fma(u1 + eps, u1 + eps, sf[a][1]);
        if (matern) {
          sb[a][0] = fma(eps - u0, eps - u0, sb[a][0]);
          sb[a][1] = fma(eps - u1, eps - u1, sb[a][1]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int row = rg * 4 + a, col = lane + 32 * c;
        const int64_t i = a0 + row, j = j0 + col;
        double m = 0.0;
        if (i < n && j < n) {
          const double kin = (i >= j) ? Kinv[i * np + j] : Kinv[j * np + i];
          const double W = 0.5 * (alpha[i] * alpha[j] - kin);
          double hsum;
          if (matern) {
            const double rf = kSqrt5 * sqrt(sf[a][c]), rb = kSqrt5 * sqrt(sb[a][c]);
            hsum = -(5.0 / 3.0) * ((1.0 + rf) * exp(-rf) + (1.0 + rb) * exp(-rb));
          } else {
            hsum = -2.0 * exp(-0.5 * sf[a][c]);
          }
          m = W * s2 * hsum;
        }
        Ms[row * 65 + col] = m;
      }
    __syncthreads();
    for (int e = tid; e < 32 * (d + 1); e += 256) {
      const int row = e / (d + 1), kk = e % (d + 1);
      const double* mr = Ms + row * 65;
      double a0s = 0.0, a1s = 0.0;
      if (kk < d) {
#pragma unroll 8
        for (int c = 0; c < 64; c += 2) {
          a0s = fma(mr[c], xj[c * dp + kk], a0s);
          a1s = fma(mr[c + 1], xj[(c + 1) * dp + kk], a1s);
        }
      } else {
#pragma unroll 8
        for (int c = 0; c < 64; c += 2) {
          a0s += mr[c];
          a1s += mr[c + 1];
        }
      }
      G[e] += a0s + a1s;
    }
  }
  __syncthreads();
  for (int e = tid; e < 32 * d; e += 256) {
    const int r = e / d, k = e % d;
    if (a0 + r < n) gX[(a0 + r) * d + k] = il[k] * il[k] * (xa[r * dp + k] * G[r * (d + 1) + d] - G[r * (d + 1) + k]);
  }
}

__global__ void k_neg_copy(const double* __restrict__ a, int64_t n, double* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = -a[i];
}

static size_t grad_x_smem(int d) {
  const int dp = d | 1;
  return sizeof(double) * (size_t(96) * dp + 32 * 65 + size_t(32) * (d + 1) + d);
}

// Needs the K^-1 (w.T) and alpha left by fit_value_grad on the same stream.
int fit_grad_inputs(const bbo_gp_hyper& h, const FitWorkspace& w, const double* X, double* gradX, double* grady,
                    cudaStream_t st) {
  if (w.batch != 1) return BBO_ERR_BAD_SHAPE;
  const size_t sm = grad_x_smem(h.d);
  if (sm > 200 * 1024) return BBO_ERR_BAD_SHAPE;   // d <= ~250
  BBO_CUDA_CHECK(cudaFuncSetAttribute(k_grad_x, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
  k_grad_x<<<unsigned(ceil_div(w.n, 32)), 256, sm, st>>>(h, w.n, w.np, X, w.hyp, w.alpha, w.T, gradX);
  BBO_LAUNCH_CHECK();
  k_neg_copy<<<unsigned(ceil_div(w.n, 256)), 256, 0, st>>>(w.alpha, w.n, grady);
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

}  // namespace bbo

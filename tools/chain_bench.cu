// Micro-benchmark of the 64x64 "factor + inverse in one sweep" kernel variants used by k_chol_chain
// (bbo_b200/csrc/fit.cu).  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
//   -o build/chain_bench tools/chain_bench.cu ; run on the GPU box: ./build/chain_bench
// Prints, per variant: us per launch (CUDA events over 200 back-to-back launches), cycles per column of
// the sweep loop (clock64 of thread 0), max |L - L_ref| and max |Linv L - I|.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

__device__ __forceinline__ double rcp3(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}

__device__ __forceinline__ void finish64(const double* dv, double* rsv, double* sqv, int tid) {
  if (tid < 64) {
    const double d = dv[tid];
    double r = double(rsqrtf(float(d)));
    r = r * fma(-0.5 * d * r, r, 1.5);
    r = r * fma(-0.5 * d * r, r, 1.5);
    double q = d * r;
    q = fma(0.5 * r, fma(-q, q, d), q);
    rsv[tid] = r;
    sqv[tid] = q;
  }
}

// ---------------------------------------------------------------- variant A: as in fit.cu (VAR 1)
__global__ void __launch_bounds__(256) sweepA(const double* __restrict__ A, double* __restrict__ Lo,
                                              double* __restrict__ Wo, long long* __restrict__ cyc) {
  __shared__ double colb[128], wrb[128], dv[64], rsv[64], sqv[64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double a[4][4], w[4][4];
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int r = ty + 16 * ia, c = tx + 16 * ib;
      a[ia][ib] = (c <= r) ? A[r * 64 + c] : 0.0;
      w[ia][ib] = (ia == ib && ty == tx) ? 1.0 : 0.0;
    }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll
  for (int jb = 0; jb < 4; ++jb) {
    for (int jt = 0; jt < 16; ++jt) {
      const int j = jb * 16 + jt;
      double* col = colb + (j & 1) * 64;
      double* wr = wrb + (j & 1) * 64;
      if (tx == jt) {
#pragma unroll
        for (int ia = jb; ia < 4; ++ia) col[ty + 16 * ia] = a[ia][jb];
      }
      if (ty == jt) {
#pragma unroll
        for (int ib = 0; ib <= jb; ++ib) wr[tx + 16 * ib] = w[jb][ib];
      }
      __syncthreads();
      const double d = col[j];
      const double rinv = rcp3(d);
      if (tid == 0) dv[j] = d;
      double m[4], cv[4], wv[4];
#pragma unroll
      for (int ia = jb; ia < 4; ++ia) m[ia] = col[ty + 16 * ia] * rinv;
#pragma unroll
      for (int ib = jb; ib < 4; ++ib) cv[ib] = col[tx + 16 * ib];
#pragma unroll
      for (int ib = 0; ib <= jb; ++ib) wv[ib] = wr[tx + 16 * ib];
#pragma unroll
      for (int ia = jb; ia < 4; ++ia)
#pragma unroll
        for (int ib = jb; ib <= ia; ++ib)
          if (ib > jb || tx > jt) a[ia][ib] = fma(-m[ia], cv[ib], a[ia][ib]);
#pragma unroll
      for (int ia = jb; ia < 4; ++ia)
#pragma unroll
        for (int ib = 0; ib <= jb; ++ib)
          if (ia > jb || ty > jt) w[ia][ib] = fma(-m[ia], wv[ib], w[ia][ib]);
      if (tx == jt) {
#pragma unroll
        for (int ia = jb; ia < 4; ++ia)
          if (ia > jb || ty > jt) a[ia][jb] = m[ia];
      }
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  __syncthreads();
  finish64(dv, rsv, sqv, tid);
  __syncthreads();
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int r = ty + 16 * ia, c = tx + 16 * ib;
      double l = 0.0, inv = 0.0;
      if (ia >= ib) {
        if (c < r) l = a[ia][ib] * sqv[c];
        else if (c == r) l = sqv[c];
        if (c <= r) inv = w[ia][ib] * rsv[r];
      }
      Lo[r * 64 + c] = l;
      Wo[r * 64 + c] = inv;
    }
}

// ---------------------------------------------------------------- variant B: no per-element predicates
// Columns keep their un-normalised values (L_ij = a_ij / sqrt(d_j) at the end), the only masks are on
// one multiplier and one column value per step.
__global__ void __launch_bounds__(256) sweepB(const double* __restrict__ A, double* __restrict__ Lo,
                                              double* __restrict__ Wo, long long* __restrict__ cyc) {
  __shared__ double colb[128], wrb[128], dv[64], rsv[64], sqv[64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double a[4][4], w[4][4];
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int r = ty + 16 * ia, c = tx + 16 * ib;
      a[ia][ib] = (c <= r) ? A[r * 64 + c] : 0.0;
      w[ia][ib] = (ia == ib && ty == tx) ? 1.0 : 0.0;
    }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll
  for (int jb = 0; jb < 4; ++jb) {
    for (int jt = 0; jt < 16; ++jt) {
      const int j = jb * 16 + jt;
      double* col = colb + (j & 1) * 64;
      double* wr = wrb + (j & 1) * 64;
      if (tx == jt) {
#pragma unroll
        for (int ia = jb; ia < 4; ++ia) col[ty + 16 * ia] = a[ia][jb];
      }
      if (ty == jt) {
#pragma unroll
        for (int ib = 0; ib <= jb; ++ib) wr[tx + 16 * ib] = w[jb][ib];
      }
      __syncthreads();
      const double d = col[j];
      const double rinv = rcp3(d);
      if (tid == 0) dv[j] = d;
      double m[4], cv[4], wv[4];
#pragma unroll
      for (int ia = jb; ia < 4; ++ia) m[ia] = col[ty + 16 * ia] * rinv;
#pragma unroll
      for (int ib = jb; ib < 4; ++ib) cv[ib] = col[tx + 16 * ib];
#pragma unroll
      for (int ib = 0; ib <= jb; ++ib) wv[ib] = wr[tx + 16 * ib];
      cv[jb] = (tx > jt) ? cv[jb] : 0.0;            // columns <= j of the pivot block are final
      const double mw = (ty > jt) ? m[jb] : 0.0;    // rows <= j of the pivot block are final (W)
#pragma unroll
      for (int ia = jb; ia < 4; ++ia)
#pragma unroll
        for (int ib = jb; ib <= ia; ++ib) a[ia][ib] = fma(-m[ia], cv[ib], a[ia][ib]);
#pragma unroll
      for (int ib = 0; ib <= jb; ++ib) w[jb][ib] = fma(-mw, wv[ib], w[jb][ib]);
#pragma unroll
      for (int ia = jb + 1; ia < 4; ++ia)
#pragma unroll
        for (int ib = 0; ib <= jb; ++ib) w[ia][ib] = fma(-m[ia], wv[ib], w[ia][ib]);
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  __syncthreads();
  finish64(dv, rsv, sqv, tid);
  __syncthreads();
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int r = ty + 16 * ia, c = tx + 16 * ib;
      double l = 0.0, inv = 0.0;
      if (ia >= ib) {
        if (c < r) l = a[ia][ib] * rsv[c];
        else if (c == r) l = sqv[c];
        if (c <= r) inv = w[ia][ib] * rsv[r];
      }
      Lo[r * 64 + c] = l;
      Wo[r * 64 + c] = inv;
    }
}

// ---------------------------------------------------------------- variant C: B with 128 threads
// thread (ty < 8, tx < 16): rows ty + 8*ia (ia < 8), cols tx + 16*ib (ib < 4)
__global__ void __launch_bounds__(128) sweepC(const double* __restrict__ A, double* __restrict__ Lo,
                                              double* __restrict__ Wo, long long* __restrict__ cyc) {
  __shared__ double colb[128], wrb[128], dv[64], rsv[64], sqv[64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double a[8][4], w[8][4];
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int r = ty + 8 * ia, c = tx + 16 * ib;
      a[ia][ib] = (c <= r) ? A[r * 64 + c] : 0.0;
      w[ia][ib] = (r == c) ? 1.0 : 0.0;
    }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll
  for (int jb = 0; jb < 4; ++jb) {
    for (int jt = 0; jt < 16; ++jt) {
      const int j = jb * 16 + jt;
      double* col = colb + (j & 1) * 64;
      double* wr = wrb + (j & 1) * 64;
      if (tx == jt) {
#pragma unroll
        for (int ia = 2 * jb; ia < 8; ++ia) col[ty + 8 * ia] = a[ia][jb];
      }
      // row j of W: row block ia = j / 8 (2jb or 2jb+1), ty == j % 8
      if (ty == (jt & 7)) {
        if (jt < 8) {
#pragma unroll
          for (int ib = 0; ib <= jb; ++ib) wr[tx + 16 * ib] = w[2 * jb][ib];
        } else {
#pragma unroll
          for (int ib = 0; ib <= jb; ++ib) wr[tx + 16 * ib] = w[2 * jb + 1][ib];
        }
      }
      __syncthreads();
      const double d = col[j];
      const double rinv = rcp3(d);
      if (tid == 0) dv[j] = d;
      double m[8], cv[4], wv[4];
#pragma unroll
      for (int ia = 2 * jb; ia < 8; ++ia) m[ia] = col[ty + 8 * ia] * rinv;
#pragma unroll
      for (int ib = jb; ib < 4; ++ib) cv[ib] = col[tx + 16 * ib];
#pragma unroll
      for (int ib = 0; ib <= jb; ++ib) wv[ib] = wr[tx + 16 * ib];
      cv[jb] = (tx > jt) ? cv[jb] : 0.0;
      // rows <= j are final for W: row blocks 2jb and 2jb+1 need masks
      const double mw0 = (ty + 16 * jb > j) ? m[2 * jb] : 0.0;
      const double mw1 = (ty + 16 * jb + 8 > j) ? m[2 * jb + 1] : 0.0;
#pragma unroll
      for (int ia = 2 * jb; ia < 8; ++ia)
#pragma unroll
        for (int ib = jb; ib < 4; ++ib)
          if (8 * ia + 7 >= 16 * ib) a[ia][ib] = fma(-m[ia], cv[ib], a[ia][ib]);
#pragma unroll
      for (int ib = 0; ib <= jb; ++ib) {
        w[2 * jb][ib] = fma(-mw0, wv[ib], w[2 * jb][ib]);
        w[2 * jb + 1][ib] = fma(-mw1, wv[ib], w[2 * jb + 1][ib]);
      }
#pragma unroll
      for (int ia = 2 * jb + 2; ia < 8; ++ia)
#pragma unroll
        for (int ib = 0; ib <= jb; ++ib) w[ia][ib] = fma(-m[ia], wv[ib], w[ia][ib]);
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  __syncthreads();
  finish64(dv, rsv, sqv, tid);
  __syncthreads();
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int r = ty + 8 * ia, c = tx + 16 * ib;
      double l = 0.0, inv = 0.0;
      if (c < r) l = a[ia][ib] * rsv[c];
      else if (c == r) l = sqv[c];
      if (c <= r) inv = w[ia][ib] * rsv[r];
      Lo[r * 64 + c] = l;
      Wo[r * 64 + c] = inv;
    }
}

// ---------------------------------------------------------------- variant D: pure latency probes
// dependent DFMA chain / MUFU.RCP64H + 3 DFMA chain / barrier + smem round trip, one CTA of 256 threads
__global__ void probes(double* out, long long* cyc) {
  __shared__ double sh[64];
  const int tid = threadIdx.x;
  double x = 1.0 + tid * 1e-9, y = 0.999999;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x = fma(x, y, 1e-12);
  }
  long long t1 = clock64();
  double r = 1.0 + tid * 1e-9;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) r = rcp3(r) + 0.5;
  long long t2 = clock64();
  double z = x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    if (tid == (i & 255)) sh[i & 63] = z;
    __syncthreads();
    z += sh[i & 63];
  }
  long long t3 = clock64();
  if (tid == 0) {
    cyc[1] = (t1 - t0) / 512;   // per dependent DFMA
    cyc[2] = (t2 - t1) / 256;   // per rcp3 + DADD
    cyc[3] = (t3 - t2) / 256;   // per STS + BAR + LDS + DADD round
  }
  out[tid] = x + r + z;
}

template <class K>
static void run(const char* name, K kern, int threads, const double* dA, double* dL, double* dW, long long* dc,
                const std::vector<double>& Lref) {
  for (int i = 0; i < 5; ++i) kern<<<1, threads>>>(dA, dL, dW, dc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int reps = 200;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) kern<<<1, threads>>>(dA, dL, dW, dc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<double> L(4096), W(4096);
  long long cyc = 0;
  CK(cudaMemcpy(L.data(), dL, 4096 * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(W.data(), dW, 4096 * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  double eL = 0, eI = 0;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j <= i; ++j) eL = fmax(eL, fabs(L[i * 64 + j] - Lref[i * 64 + j]));
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += W[i * 64 + k] * Lref[k * 64 + j];
      eI = fmax(eI, fabs(s - (i == j ? 1.0 : 0.0)));
    }
  printf("%s: %.2f us/launch, sweep %lld cycles (%.0f per column), max|L-Lref| %.2e, max|W L - I| %.2e\n", name,
         ms * 1e3 / reps, cyc, cyc / 64.0, eL, eI);
}

int main() {
  std::vector<double> B(4096), A(4096, 0.0), Lref(4096, 0.0);
  srand(1);
  for (auto& v : B) v = rand() / double(RAND_MAX) - 0.5;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {
      double s = (i == j) ? 4.0 : 0.0;
      for (int k = 0; k < 64; ++k) s += B[i * 64 + k] * B[j * 64 + k];
      A[i * 64 + j] = s;
    }
  for (int j = 0; j < 64; ++j) {
    double s = A[j * 64 + j];
    for (int k = 0; k < j; ++k) s -= Lref[j * 64 + k] * Lref[j * 64 + k];
    Lref[j * 64 + j] = sqrt(s);
    for (int i = j + 1; i < 64; ++i) {
      double t = A[i * 64 + j];
      for (int k = 0; k < j; ++k) t -= Lref[i * 64 + k] * Lref[j * 64 + k];
      Lref[i * 64 + j] = t / Lref[j * 64 + j];
    }
  }
  double *dA, *dL, *dW;
  long long* dc;
  CK(cudaMalloc(&dA, 4096 * 8));
  CK(cudaMalloc(&dL, 4096 * 8));
  CK(cudaMalloc(&dW, 4096 * 8));
  CK(cudaMalloc(&dc, 64));
  CK(cudaMemcpy(dA, A.data(), 4096 * 8, cudaMemcpyHostToDevice));
  run("A (fit.cu VAR 1, 256 thr)", sweepA, 256, dA, dL, dW, dc, Lref);
  run("B (mask-only, 256 thr)   ", sweepB, 256, dA, dL, dW, dc, Lref);
  run("C (mask-only, 128 thr)   ", sweepC, 128, dA, dL, dW, dc, Lref);
  probes<<<1, 256>>>(dL, dc);
  CK(cudaDeviceSynchronize());
  long long c[4];
  CK(cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost));
  printf("probes: dependent DFMA %lld cyc, rcp3+DADD %lld cyc, STS+BAR+LDS+DADD round (256 thr) %lld cyc\n", c[1],
         c[2], c[3]);
  return 0;
}

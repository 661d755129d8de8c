// Correctness + timing of block_factor64 (bbo_b200/csrc/block_factor.cuh) on one 64x64 SPD block.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/factor_bench tools/factor_bench.cu
//   gpurun -- ./build/factor_bench
// Prints cycles per call (clock64 of thread 0), max |L - L_ref| / max|L_ref| and max |W L - I| for 128 and 256 threads.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define BBO_FACTOR_PROF 1
#include "block_factor.cuh"

namespace bbo {
int cuda_fail(cudaError_t, const char*) { return BBO_ERR_CUDA; }
void count_launch() {}
}  // namespace bbo
using namespace bbo;

template <int NT>
__global__ void __launch_bounds__(NT) k_test(const double* __restrict__ A, double* __restrict__ Lo,
                                             double* __restrict__ Wo, double* __restrict__ misc,
                                             long long* __restrict__ cyc) {
  extern __shared__ __align__(16) double dyn[];
  double* Ds = dyn;
  double* Ws = Ds + 64 * kLT;
  BlockFactorScratch& sc = *reinterpret_cast<BlockFactorScratch*>(Ws + 64 * kLT);
  for (int idx = threadIdx.x; idx < 4096; idx += NT) Ds[(idx >> 6) * kLT + (idx & 63)] = A[idx];
  __syncthreads();
  const long long t0 = clock64();
  block_factor64<NT>(Ds, Ws, &sc);
  const long long t1 = clock64();
  for (int idx = threadIdx.x; idx < 4096; idx += NT) {
    Lo[idx] = Ds[(idx >> 6) * kLT + (idx & 63)];
    Wo[idx] = Ws[(idx >> 6) * kLT + (idx & 63)];
  }
  if (threadIdx.x < 64) cyc[8 + threadIdx.x] = sc.prof[threadIdx.x];
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0;
    misc[0] = sc.logsum;
    misc[1] = double(sc.fail);
  }
}

template <int NT>
static void run(const std::vector<double>& A, const std::vector<double>& Lref, double logref) {
  double *dA, *dL, *dW, *dm;
  long long* dc;
  cudaMalloc(&dA, 4096 * 8);
  cudaMalloc(&dL, 4096 * 8);
  cudaMalloc(&dW, 4096 * 8);
  cudaMalloc(&dm, 64);
  cudaMalloc(&dc, 1024);
  cudaMemcpy(dA, A.data(), 4096 * 8, cudaMemcpyHostToDevice);
  constexpr size_t kSm = 2 * 64 * kLT * sizeof(double) + sizeof(BlockFactorScratch);
  cudaFuncSetAttribute(k_test<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSm));
  for (int rep = 0; rep < 3; ++rep) k_test<NT><<<1, NT, kSm>>>(dA, dL, dW, dm, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("NT=%d CUDA error %s\n", NT, cudaGetErrorString(e));
    exit(1);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int rep = 0; rep < 200; ++rep) k_test<NT><<<1, NT, kSm>>>(dA, dL, dW, dm, dc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> L(4096), W(4096);
  double misc[2];
  long long cyc;
  cudaMemcpy(L.data(), dL, 4096 * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(W.data(), dW, 4096 * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(misc, dm, 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  double el = 0, lmax = 0, ew = 0, up = 0;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {
      lmax = fmax(lmax, fabs(Lref[i * 64 + j]));
      el = fmax(el, fabs(L[i * 64 + j] - Lref[i * 64 + j]));
      if (j > i) up = fmax(up, fmax(fabs(L[i * 64 + j]), fabs(W[i * 64 + j])));
      double s = 0;
      for (int k = 0; k < 64; ++k) s += W[i * 64 + k] * Lref[k * 64 + j];
      ew = fmax(ew, fabs(s - (i == j ? 1.0 : 0.0)));
    }
  long long pr[64];
  cudaMemcpy(pr, dc + 8, sizeof(pr), cudaMemcpyDeviceToHost);
  printf("NT=%d phases (cycles): ", NT);
  for (int kb = 0; kb < 4; ++kb)
    printf("[kb%d factor %lld panel %lld trail %lld] ", kb, pr[2 + 4 * kb] - pr[1 + 4 * kb],
           kb < 3 ? pr[3 + 4 * kb] - pr[2 + 4 * kb] : 0LL, kb < 3 ? pr[4 + 4 * kb] - pr[3 + 4 * kb] : 0LL);
  printf("inverse %lld %lld %lld\n", pr[21] - pr[20], pr[22] - pr[21], pr[23] - pr[22]);
  printf("  kb0: load + 16 steps %lld, scale+store %lld, barrier after %lld\n", pr[40] - pr[1], pr[41] - pr[40], pr[2] - pr[41]);
  printf("NT=%d: %lld cycles per block (%.2f us per launch incl. launch), |L-Lref|/max %.2e, |W L - I| %.2e, "
         "upper %.1e, logsum err %.2e, fail %d\n", NT, cyc, ms / 200 * 1e3, el / lmax, ew, up, fabs(misc[0] - logref),
         int(misc[1]));
}

__global__ void probes(double* out, long long* cyc) {
  double x = out[0] + 1.000001, y = 1.0000001;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x = fma(x, y, 1e-9);
  long long t1 = clock64();
  double r = x;
#pragma unroll
  for (int i = 0; i < 16; ++i) r = fast_rcp<true>(r) + 1.5;
  long long t2 = clock64();
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = x + i;
  long long t3 = clock64();
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], y, r);
  long long t4 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[1] = s + r;
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0) / 64;
    cyc[1] = (t2 - t1) / 16;
    cyc[2] = (t4 - t3) / 8;
  }
}

// step-loop experiments (one warp, 64 steps = 4 panels), cycles per step
//   MODE 0 full step; 1 no smem traffic (stale column); 2 no reciprocal (rinv = 1); 3 single buffer + two __syncwarp;
//   4 shuffle broadcast instead of smem
template <int MODE>
__global__ void k_steps(double* out, long long* cyc) {
  __shared__ __align__(16) double colb[2][16];
  const int tid = threadIdx.x, c = tid & 15, h = (tid >> 4) & 1;
  // (the other warps of the CTA wait at the barrier at the end, as in block_factor64)
  if (tid < 32) {
  double v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 4.0 + 0.01 * (i + c) + out[0];
  if (tid < 16) colb[0][tid] = colb[1][tid] = 4.0 + 0.01 * tid;
  __syncwarp();
  const long long t0 = clock64();
#pragma unroll 1
  for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      double* col = MODE == 3 ? colb[0] : colb[j & 1];
      volatile double* colv = col;
      double cj[16];
      if (MODE == 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) cj[i] = __shfl_sync(0xffffffffu, v[i], j);
      } else {
        if (MODE != 1) {
          if (MODE == 3) __syncwarp();
          if (tid == j) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) *reinterpret_cast<double2*>(col + i) = make_double2(v[i], v[i + 1]);
          }
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const double2 t = *reinterpret_cast<const double2*>(col + i);
          cj[i] = t.x;
          cj[i + 1] = t.y;
        }
      }
      const double d = cj[j];
      const double rinv = MODE == 2 ? 1.0 : fast_rcp<true>(d);
      const bool act = h == 0 ? (c > j) : (c <= j);
      const double x = h == 0 ? (MODE == 4 ? __shfl_sync(0xffffffffu, cj[0], 0) : colv[c > j ? c : j]) : v[j];
      const double f = act ? rinv * x * 1e-3 : 0.0;
#pragma unroll
      for (int i = j + 1; i < 16; ++i) v[i] = fma(-cj[i], f, v[i]);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[1 + tid] = s;
  if (tid == 0) cyc[0] = (t1 - t0) / 64;
  }
  __syncthreads();
}

template <int MODE>
static void run_steps(const char* what) {
  double* d;
  long long* c;
  cudaMalloc(&d, 1024);
  cudaMalloc(&c, 64);
  cudaMemset(d, 0, 1024);
  for (int nt : {32, 128, 256}) {
    k_steps<MODE><<<1, nt>>>(d, c);
    k_steps<MODE><<<1, nt>>>(d, c);
    long long h = 0;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("step experiment %d (%s), %d threads in the CTA: %lld cycles per step\n", MODE, what, nt, h);
  }
}

int main() {
  run_steps<0>("full");
  run_steps<1>("no smem publish");
  run_steps<2>("no reciprocal");
  run_steps<3>("single buffer, two syncwarps");

  std::vector<double> B(4096), A(4096, 0.0), Lref(4096, 0.0);
  srand(1);
  for (auto& v : B) v = rand() / double(RAND_MAX) - 0.5;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {
      double s = (i == j) ? 4.0 : 0.0;
      for (int k = 0; k < 64; ++k) s += B[i * 64 + k] * B[j * 64 + k];
      A[i * 64 + j] = s;
    }
  double logref = 0;
  for (int j = 0; j < 64; ++j) {
    double s = A[j * 64 + j];
    for (int k = 0; k < j; ++k) s -= Lref[j * 64 + k] * Lref[j * 64 + k];
    Lref[j * 64 + j] = sqrt(s);
    logref += log(Lref[j * 64 + j]);
    for (int i = j + 1; i < 64; ++i) {
      double t = A[i * 64 + j];
      for (int k = 0; k < j; ++k) t -= Lref[i * 64 + k] * Lref[j * 64 + k];
      Lref[i * 64 + j] = t / Lref[j * 64 + j];
    }
  }
  // garbage in the strict upper triangle of the input: must be ignored
  for (int i = 0; i < 64; ++i)
    for (int j = i + 1; j < 64; ++j) A[i * 64 + j] = 1e30;
  {
    double* d;
    long long* c;
    cudaMalloc(&d, 64);
    cudaMalloc(&c, 64);
    cudaMemset(d, 0, 64);
    probes<<<1, 32>>>(d, c);
    long long h[3];
    cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
    printf("probes (one warp): dependent DFMA %lld cycles, fast_rcp<true>+DADD %lld cycles, 16 independent DFMA %lld cycles\n",
           h[0], h[1], h[2]);
  }
  run<128>(A, Lref, logref);
  run<256>(A, Lref, logref);
  return 0;
}

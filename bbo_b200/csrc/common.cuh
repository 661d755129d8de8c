// Shared device helpers for the bbo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/bbo_b200.h"

namespace bbo {

constexpr int kNB = 64;            // Cholesky / TRTRI block size
constexpr int kPad = BBO_ROW_ALIGN;  // n is padded to a multiple of this (128)
constexpr double kSqrt5 = 2.23606797749978969640917366873128;
constexpr double kHalfLog2Pi = 0.91893853320467274178032973640562;

inline int64_t padded(int64_t n) { return (n + kPad - 1) / kPad * kPad; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Records the last CUDA error text for bbo_last_cuda_error().
int cuda_fail(cudaError_t e, const char* what);
#define BBO_CUDA_CHECK(expr)                                    \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::bbo::cuda_fail(_e, #expr);  \
  } while (0)
// every kernel launch goes through this: counts launches (bbo_launch_count) and checks errors
void count_launch();
#define BBO_LAUNCH_CHECK()               \
  do {                                   \
    ::bbo::count_launch();               \
    BBO_CUDA_CHECK(cudaGetLastError());  \
  } while (0)

// optional per-region device timing with CUDA events on the launching stream (bbo_profile_*)
enum ProfTag {
  kProfVnorm = 0,   // FP64 posterior-variance panel (k_vnorm)
  kProfKstar = 1,   // cross-covariance + mean (k_kstar)
  kProfSelect = 2,  // acquisition epilogue + top-k
  kProfChol = 3,    // covariance build + blocked Cholesky
  kProfTrtri = 4,   // triangular inverse + solves
  kProfGrad = 5,    // LAUUM + gradient reduction (+ Adam)
  kProfTf32 = 6,    // TF32 scoring panel (tcgen05)
  kProfTags = 8
};
bool prof_enabled();
long long launch_count();
void add_launches(long long n);
// destroys an executable graph once everything enqueued on `st` so far has completed
void retire_graph(cudaGraphExec_t exec, cudaGraph_t graph, cudaStream_t st);
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);

// cudaFuncSetAttribute (opt-in shared memory) applies to the CURRENT device only: one flag per device.
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // D(8x8) += A(8x4, row) * B(4x8, col); SASS: DMMA.8x8x4
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ---------------------------------------------------------------- small math
__device__ __forceinline__ double softplus_d(double x) {
  // torch.nn.functional.softplus, beta=1, threshold=20 (kernels.py:35,52)
  return x > 20.0 ? x : log1p(exp(x));
}
__device__ __forceinline__ double softplus_grad_d(double x) {
  if (x > 20.0) return 1.0;
  double z = exp(x);
  return z / (z + 1.0);
}

// base kernel value from the accumulated squared scaled distance `s`
//   RBF      s = sum ((x1-x2)/l)^2            -> exp(-s/2)          kernel_impl.py:29-31
//   Matern   s = sum ((x2-x1)/l + eps)^2, r = sqrt(5 s)
//            -> (1 + r + r^2/3) exp(-r)                             kernel_impl.py:33-34
__device__ __forceinline__ double kernel_from_s(int kind, double s) {
  if (kind == BBO_KERNEL_RBF) return exp(-0.5 * s);
  double r = kSqrt5 * sqrt(s);
  return (1.0 + r + r * r * (1.0 / 3.0)) * exp(-r);
}

// Ordering of (score, index) pairs for selection: descending score, ties -> lower index
// (the stable sort of shared/trial.py:289-294).  NaN never wins.
__device__ __forceinline__ bool better(double sa, int64_t ia, double sb, int64_t ib) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  return (sa > sb) || (sa == sb && ia < ib);
}

}  // namespace bbo

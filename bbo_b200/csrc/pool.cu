// Counter-based uniform candidate generator (Philox4x32-10).  Replaces, for the acquisition
// pool, RandomDesigner.suggest (bbo/algorithms/designers/random.py:41-57): values depend only
// on (seed, global row, column) so any sharding of the pool yields identical candidates.
#include "score.cuh"

namespace bbo {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

// one Philox block (4 x 32 bit) per (row, group of 2 columns): 53-bit doubles / 24-bit floats
template <typename T>
__global__ void __launch_bounds__(256)
k_pool_uniform(uint64_t seed, int64_t row_offset, int64_t q, int64_t d, T* __restrict__ out) {
  const int64_t pairs = (d + 1) / 2;
  const int64_t t = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (t >= q * pairs) return;
  const int64_t i = t / pairs, p = t % pairs;
  const uint64_t row = uint64_t(row_offset + i);
  uint32_t c[4] = {uint32_t(row), uint32_t(row >> 32), uint32_t(p), 0u};
  philox4x32_10(c, uint32_t(seed), uint32_t(seed >> 32));
  const uint64_t a = (uint64_t(c[0]) << 32) | c[1], b = (uint64_t(c[2]) << 32) | c[3];
  const int64_t col = 2 * p;
  if (sizeof(T) == 8) {
    out[i * d + col] = T(double(a >> 11) * (1.0 / 9007199254740992.0));
    if (col + 1 < d) out[i * d + col + 1] = T(double(b >> 11) * (1.0 / 9007199254740992.0));
  } else {
    out[i * d + col] = T(float(a >> 40) * (1.0f / 16777216.0f));
    if (col + 1 < d) out[i * d + col + 1] = T(float(b >> 40) * (1.0f / 16777216.0f));
  }
}

int pool_uniform(uint64_t seed, int64_t row_offset, int64_t q, int64_t d, void* out, int is_f32, cudaStream_t st) {
  if (q <= 0 || d <= 0) return BBO_OK;
  const int64_t total = q * ((d + 1) / 2);
  if (is_f32)
    k_pool_uniform<float><<<unsigned(ceil_div(total, 256)), 256, 0, st>>>(seed, row_offset, q, d, static_cast<float*>(out));
  else
    k_pool_uniform<double><<<unsigned(ceil_div(total, 256)), 256, 0, st>>>(seed, row_offset, q, d, static_cast<double*>(out));
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

// Kumaraswamy input warp of a candidate pool (gp/warpers.py:52-65), the prologue of scoring when the
// covariance is a WarpKernel(KumarWarp): out = 1 - (1 - clip(x, eps, 1-eps)^a)^b, computed in float64,
// one pass over the pool (read once, written once).
template <typename T>
__global__ void __launch_bounds__(256)
k_warp_kumar(const T* __restrict__ x, int64_t count, double a, double b, double eps, T* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= count) return;
  double v = double(x[i]);
  v = fmin(fmax(v, eps), 1.0 - eps);      // NaN stays NaN (torch.clip propagates it too)
  out[i] = T(1.0 - pow(1.0 - pow(v, a), b));
}

int warp_kumar(const void* x, int64_t count, double a, double b, double eps, void* out, int is_f32,
               cudaStream_t st) {
  if (count <= 0) return BBO_OK;
  if (is_f32)
    k_warp_kumar<float><<<unsigned(ceil_div(count, 256)), 256, 0, st>>>(static_cast<const float*>(x), count, a, b, eps,
                                                                       static_cast<float*>(out));
  else
    k_warp_kumar<double><<<unsigned(ceil_div(count, 256)), 256, 0, st>>>(static_cast<const double*>(x), count, a, b,
                                                                        eps, static_cast<double*>(out));
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

}  // namespace bbo

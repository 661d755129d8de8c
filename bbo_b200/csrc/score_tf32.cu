// TF32 scoring mode (tcgen05 panel) -- placeholder until the tcgen05 pipeline lands.
#include "score.cuh"

namespace bbo {

size_t score_tf32_extra_bytes(int64_t, int64_t, int64_t) { return 0; }

int score_tf32_chunk(const bbo_gp_state&, const ScoreWorkspace&, const void*, int, int64_t, int64_t,
                     cudaStream_t) {
  return BBO_ERR_BAD_ARG;
}

__global__ void k_make_tf32(int64_t total, const double* __restrict__ in, float* __restrict__ out) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) {
    float f = static_cast<float>(in[i]);
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
    out[i] = __uint_as_float(u);
  }
}

int make_tf32(int64_t n, const double* linv, float* out, cudaStream_t st) {
  const int64_t np = padded(n), total = np * np;
  k_make_tf32<<<unsigned(ceil_div(total, 256)), 256, 0, st>>>(total, linv, out);
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

}  // namespace bbo

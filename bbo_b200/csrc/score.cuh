// Host-side declarations of the scoring path (score.cu, score_tf32.cu).
#pragma once
#include <functional>

#include "common.cuh"

namespace bbo {

constexpr int kMaxSplit = 32;  // max row-tile splits of the variance panel per candidate tile

struct ScoreWorkspace {
  int64_t np = 0, qp = 0;      // padded n, padded chunk length
  double* Ks = nullptr;        // (qp, np)  cross-covariance of the current chunk
  double* V = nullptr;         // (qp, np)  L^-1 k*, only for the full-covariance predict
  double* vpart = nullptr;     // (kMaxSplit, qp) partial |v|^2
  double* mu = nullptr;        // (qp)
  double* score = nullptr;     // (qp)
  int64_t list_len = 0;
  double *la_s = nullptr, *lb_s = nullptr;   // top-k ping-pong lists
  int64_t *la_i = nullptr, *lb_i = nullptr;
  int* surv_count = nullptr;   // survivor counter of the threshold-filtered selection
  void* tf32 = nullptr;        // TF32-mode scratch (score_tf32.cu)
  void* refine = nullptr;      // BBO_PREC_TF32_REFINE scratch (short-list, gathered rows, small FP64 buffers)

  static size_t bytes(int64_t n, int64_t d, int64_t q_chunk, int precision, bool want_full);
  int bind(void* ws, size_t ws_bytes, int64_t n, int64_t d, int64_t q_chunk, int precision, bool want_full);
};

int gp_predict(const bbo_gp_state& s, const double* Xq, int64_t q, double* mu, double* var, double* cov_full,
               void* ws, size_t ws_bytes, int64_t q_chunk, cudaStream_t st);
// refine_flags (BBO_PREC_TF32_REFINE only; 0 for a self-contained call): the host-pool path scores its staging
// buffers with kRefineDefer (the short-list and its rows stay in the workspace, nothing is re-scored) and
// kRefineContinue (the short-list of the previous call, kept in the other row buffer, is merged), and re-scores once
// on its last call.
constexpr int kRefineContinue = 1, kRefineDefer = 2, kRefineParity = 4;
int gp_score_pool(const bbo_gp_state& s, const bbo_acq& acq, int precision, const void* Xq, int xq_is_f32,
                  int64_t q, int64_t index_offset, int k, double* topk_s, int64_t* topk_i, double* mu_out,
                  double* var_out, double* score_out, int32_t* nan_count, void* ws, size_t ws_bytes,
                  int64_t q_chunk, bool keep_running, cudaStream_t st, int refine_flags = 0);
int refine_shortlist_len(int k);
int refine_info(const void* ws, size_t ws_bytes, int64_t n, int64_t d, int64_t q_chunk, double* info_host,
                cudaStream_t st);
size_t score_batched_workspace_bytes(int64_t n, int64_t d, int64_t B, int64_t q);
int gp_score_pool_batched(const bbo_gp_hyper& h, int64_t n, int64_t B, const double* X, const double* hyp,
                          const double* linv, const double* alpha, const bbo_acq* acq_dev, const double* Xq,
                          int64_t q, int k, double* topk_s, int64_t* topk_i, void* ws, size_t ws_bytes,
                          cudaStream_t st);
int acq_eval(const bbo_acq& a, const void* mu, const void* sd, void* out, int64_t n, int is_f32, cudaStream_t st);
int topk_merge(const double* score, const int64_t* index, int64_t m, int k, double* out_s, int64_t* out_i,
               cudaStream_t st);
int topk_pack(const double* score, const int64_t* index, int k, int64_t* packed, cudaStream_t st);
int topk_merge_packed(const int64_t* packed, int64_t m, int k, double* out_s, int64_t* out_i, cudaStream_t st);

// TF32 mode (score_tf32.cu): pipelined K* (aux stream) / tcgen05 panel (st) over all chunks of a pool;
// consume(chunk_start, qc, mu, nsplit, vpart, stream) enqueues the per-chunk acquisition / selection work.
using ConsumeFn = std::function<int(int64_t, int64_t, const double*, int, const double*, cudaStream_t)>;
size_t score_tf32_extra_bytes(int64_t n, int64_t d, int64_t qp);
int score_tf32_pipeline(const bbo_gp_state& s, const ScoreWorkspace& w, const void* Xq, int xq_is_f32, int64_t q,
                        const ConsumeFn& consume, cudaStream_t st, bool short_first = false);
int make_tf32(int64_t n, const double* linv, float* out, cudaStream_t st);
int64_t tf32_rows(int64_t n);

int warp_kumar(const void* x, int64_t count, double a, double b, double eps, void* out, int is_f32,
               cudaStream_t st);
int pool_uniform(uint64_t seed, int64_t row_offset, int64_t q, int64_t d, void* out, int is_f32, cudaStream_t st);

}  // namespace bbo

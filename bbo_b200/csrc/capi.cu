// extern "C" surface of libbbo_b200.so (see include/bbo_b200.h for the contract).
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "fit.cuh"
#include "score.cuh"

namespace bbo {
static thread_local std::string g_last_error;
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + what;
  return BBO_ERR_CUDA;
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }
void add_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Executable graphs retired by fit_adam: destroyed lazily, once the event recorded after their last
// launch has completed (checked on the next retire call) -- never while launches are pending.
struct Retired {
  cudaGraphExec_t exec;
  cudaGraph_t graph;
  cudaEvent_t done;
};
static std::vector<Retired> g_retired;
static std::mutex g_state_mutex;   // guards the library-global caches below (g_retired, g_prof, HostAux)
void retire_graph(cudaGraphExec_t exec, cudaGraph_t graph, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_state_mutex);
  for (size_t i = 0; i < g_retired.size();) {
    if (cudaEventQuery(g_retired[i].done) == cudaSuccess) {
      cudaGraphExecDestroy(g_retired[i].exec);
      cudaGraphDestroy(g_retired[i].graph);
      cudaEventDestroy(g_retired[i].done);
      g_retired[i] = g_retired.back();
      g_retired.pop_back();
    } else {
      ++i;
    }
  }
  cudaGetLastError();   // cudaEventQuery returns cudaErrorNotReady for pending work
  Retired r{exec, graph, nullptr};
  if (cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming) == cudaSuccess) {
    cudaEventRecord(r.done, st);
    g_retired.push_back(r);
  }
}

struct ProfState {
  bool enabled = false;
  std::vector<cudaEvent_t> ev[kProfTags];  // begin/end pairs
  size_t used[kProfTags] = {0};
};
static ProfState g_prof;
bool prof_enabled() { return g_prof.enabled; }
constexpr size_t kProfMaxEvents = 1 << 14;
static cudaEvent_t prof_event(int tag) {
  auto& v = g_prof.ev[tag];
  if (g_prof.used[tag] >= kProfMaxEvents) return nullptr;
  if (g_prof.used[tag] >= v.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    v.push_back(e);
  }
  return v[g_prof.used[tag]++];
}
void prof_begin(int tag, cudaStream_t st) {
  if (!g_prof.enabled) return;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  if (g_prof.used[tag] + 2 > kProfMaxEvents) return;
  cudaEvent_t e = prof_event(tag);
  if (e) cudaEventRecord(e, st);
}
void prof_end(int tag, cudaStream_t st) {
  if (!g_prof.enabled) return;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  if ((g_prof.used[tag] & 1) == 0) return;  // no matching begin
  cudaEvent_t e = prof_event(tag);
  if (e) cudaEventRecord(e, st);
}
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
static bool hyper_ok(const bbo_gp_hyper* h) {
  return h != nullptr && h->d >= 1 && (h->kind == BBO_KERNEL_RBF || h->kind == BBO_KERNEL_MATERN52) &&
         h->noise >= 0.0;
}
}  // namespace bbo

using namespace bbo;

extern "C" {

int bbo_version(void) { return 100; }
int64_t bbo_launch_count(void) { return g_launches.load(); }
void bbo_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lock(g_state_mutex);
  g_prof.enabled = on != 0;
  for (int t = 0; t < kProfTags; ++t) g_prof.used[t] = 0;
}
int bbo_profile_read(int32_t tag, int64_t* regions, double* total_ms) {
  if (tag < 0 || tag >= kProfTags || !regions || !total_ms) return BBO_ERR_BAD_ARG;
  BBO_CUDA_CHECK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_state_mutex);
  double tot = 0.0;
  int64_t n = 0;
  for (size_t i = 0; i + 1 < g_prof.used[tag]; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.ev[tag][i], g_prof.ev[tag][i + 1]) == cudaSuccess) {
      tot += ms;
      ++n;
    }
  }
  g_prof.used[tag] = 0;
  *regions = n;
  *total_ms = tot;
  return BBO_OK;
}
const char* bbo_last_cuda_error(void) { return g_last_error.c_str(); }
int64_t bbo_padded_n(int64_t n) { return padded(n); }
int64_t bbo_tf32_rows(int64_t n) { return tf32_rows(n); }

size_t bbo_gp_fit_workspace_bytes(int64_t n, int64_t d, int64_t batch) {
  if (n < 1 || d < 1 || batch < 1) return 0;
  return FitWorkspace::bytes(n, d, batch);
}
size_t bbo_gp_score_workspace_bytes(int64_t n, int64_t d, int64_t q_chunk, int32_t precision) {
  if (n < 1 || d < 1 || q_chunk < 1) return 0;
  return ScoreWorkspace::bytes(n, d, q_chunk, precision, false);
}
size_t bbo_gp_predict_workspace_bytes(int64_t n, int64_t d, int64_t q_chunk, int32_t want_full) {
  if (n < 1 || d < 1 || q_chunk < 1) return 0;
  return ScoreWorkspace::bytes(n, d, q_chunk, BBO_PREC_FP64, want_full != 0);
}

int bbo_cov_matrix(const bbo_gp_hyper* h, const double* hyp_dev, const double* X1_dev, int64_t q,
                   const double* X2_dev, int64_t n, double* out_dev, int64_t ldo, void* stream) {
  if (!hyper_ok(h) || q < 0 || n < 0 || ldo < n) return BBO_ERR_BAD_SHAPE;
  if (!hyp_dev || !X1_dev || !X2_dev || !out_dev) return BBO_ERR_BAD_ARG;
  return cov_matrix(*h, hyp_dev, X1_dev, q, X2_dev, n, out_dev, ldo, 0.0, S(stream));
}

size_t bbo_gp_append_workspace_bytes(int64_t n_old, int64_t m) {
  if (n_old < 1 || m < 1 || m > BBO_APPEND_MAX) return 0;
  return append_workspace_bytes(n_old, m);
}
int bbo_gp_append(const bbo_gp_hyper* h, int64_t n_old, int64_t m, const double* X_dev, const double* y_dev,
                  const double* hyp_dev, double* linv_dev, double* alpha_dev, double* logdet_dev,
                  int32_t* info_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!hyper_ok(h) || n_old < 1 || m < 1 || m > BBO_APPEND_MAX) return BBO_ERR_BAD_SHAPE;
  if (!X_dev || !y_dev || !hyp_dev || !linv_dev || !alpha_dev || !info_dev || !workspace_dev) return BBO_ERR_BAD_ARG;
  return gp_append(*h, n_old, m, X_dev, y_dev, hyp_dev, linv_dev, alpha_dev, logdet_dev, info_dev, workspace_dev,
                   workspace_bytes, S(stream));
}

int bbo_gp_fit_value_grad(const bbo_gp_hyper* h, int64_t n, int64_t batch, const double* X_dev,
                          const double* y_dev, const double* hyp_dev, double* value_dev, double* grad_dev,
                          int32_t* info_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!hyper_ok(h) || n < 1 || batch < 1) return BBO_ERR_BAD_SHAPE;
  if (!X_dev || !y_dev || !hyp_dev || !value_dev || !grad_dev || !info_dev) return BBO_ERR_BAD_ARG;
  FitWorkspace w;
  int rc = w.bind(workspace_dev, workspace_bytes, n, h->d, batch);
  if (rc != BBO_OK) return rc;
  return fit_value_grad(*h, w, X_dev, y_dev, hyp_dev, value_dev, grad_dev, info_dev, S(stream));
}

int bbo_gp_fit_value_grad_inputs(const bbo_gp_hyper* h, int64_t n, const double* X_dev, const double* y_dev,
                                 const double* hyp_dev, double* value_dev, double* grad_dev, double* gradX_dev,
                                 double* grady_dev, int32_t* info_dev, void* workspace_dev,
                                 size_t workspace_bytes, void* stream) {
  if (!hyper_ok(h) || n < 1) return BBO_ERR_BAD_SHAPE;
  if (!X_dev || !y_dev || !hyp_dev || !value_dev || !grad_dev || !gradX_dev || !grady_dev || !info_dev)
    return BBO_ERR_BAD_ARG;
  FitWorkspace w;
  int rc = w.bind(workspace_dev, workspace_bytes, n, h->d, 1);
  if (rc != BBO_OK) return rc;
  rc = fit_value_grad(*h, w, X_dev, y_dev, hyp_dev, value_dev, grad_dev, info_dev, S(stream));
  if (rc != BBO_OK) return rc;
  return fit_grad_inputs(*h, w, X_dev, gradX_dev, grady_dev, S(stream));
}

int bbo_gp_fit_adam(const bbo_gp_hyper* h, int64_t n, int64_t batch, int32_t ard, int32_t has_scale,
                    const double* X_dev, const double* y_dev, double* theta_dev, double* adam_m_dev,
                    double* adam_v_dev, int64_t step0, int64_t epochs, double lr, double sign,
                    double* values_dev, int32_t* info_dev, void* workspace_dev, size_t workspace_bytes,
                    void* stream) {
  if (!hyper_ok(h) || n < 1 || batch < 1 || epochs < 0 || step0 < 0) return BBO_ERR_BAD_SHAPE;
  if (!X_dev || !y_dev || !theta_dev || !adam_m_dev || !adam_v_dev || !info_dev) return BBO_ERR_BAD_ARG;
  FitWorkspace w;
  int rc = w.bind(workspace_dev, workspace_bytes, n, h->d, batch);
  if (rc != BBO_OK) return rc;
  return fit_adam(*h, w, ard != 0, has_scale != 0, X_dev, y_dev, theta_dev, adam_m_dev, adam_v_dev, step0,
                  epochs, lr, sign, values_dev, info_dev, S(stream));
}

int bbo_gp_theta_to_hyp(int64_t d, int64_t batch, int32_t ard, int32_t has_scale, const double* theta_dev,
                        double* hyp_dev, void* stream) {
  if (d < 1 || batch < 1) return BBO_ERR_BAD_SHAPE;
  if (!theta_dev || !hyp_dev) return BBO_ERR_BAD_ARG;
  return theta_to_hyp(d, batch, ard != 0, has_scale != 0, theta_dev, hyp_dev, S(stream));
}

int bbo_gp_factorize(const bbo_gp_hyper* h, int64_t n, int64_t batch, const double* X_dev, const double* y_dev,
                     const double* hyp_dev, double* linv_dev, double* alpha_dev, double* chol_dev,
                     double* logdet_dev, int32_t* info_dev, void* workspace_dev, size_t workspace_bytes,
                     void* stream) {
  if (!hyper_ok(h) || n < 1 || batch < 1) return BBO_ERR_BAD_SHAPE;
  if (!X_dev || !y_dev || !hyp_dev || !linv_dev || !alpha_dev || !info_dev) return BBO_ERR_BAD_ARG;
  FitWorkspace w;
  int rc = w.bind(workspace_dev, workspace_bytes, n, h->d, batch);
  if (rc != BBO_OK) return rc;
  cudaStream_t st = S(stream);
  const size_t hb = size_t(batch) * (2 * h->d + 2) * sizeof(double);
  BBO_CUDA_CHECK(cudaMemcpyAsync(w.hyp, hyp_dev, hb, cudaMemcpyDeviceToDevice, st));
  rc = fit_factor(*h, w, X_dev, y_dev, w.hyp, info_dev, st);
  if (rc != BBO_OK) return rc;
  const size_t mb = size_t(batch) * w.np * w.np * sizeof(double);
  BBO_CUDA_CHECK(cudaMemcpyAsync(linv_dev, w.Wlo, mb, cudaMemcpyDeviceToDevice, st));
  BBO_CUDA_CHECK(cudaMemcpyAsync(alpha_dev, w.alpha, size_t(batch) * w.np * sizeof(double),
                                 cudaMemcpyDeviceToDevice, st));
  if (chol_dev) BBO_CUDA_CHECK(cudaMemcpyAsync(chol_dev, w.L, mb, cudaMemcpyDeviceToDevice, st));
  if (logdet_dev)
    BBO_CUDA_CHECK(cudaMemcpyAsync(logdet_dev, w.logdet, size_t(batch) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return BBO_OK;
}

int bbo_gp_make_tf32(int64_t n, const double* linv_dev, float* linv_tf32_dev, void* stream) {
  if (n < 1) return BBO_ERR_BAD_SHAPE;
  if (!linv_dev || !linv_tf32_dev) return BBO_ERR_BAD_ARG;
  return make_tf32(n, linv_dev, linv_tf32_dev, S(stream));
}

static bool state_ok(const bbo_gp_state* st) {
  return st && hyper_ok(&st->h) && st->n >= 1 && st->X && st->hyp && st->linv && st->alpha;
}

int bbo_gp_predict(const bbo_gp_state* st, const double* Xq_dev, int64_t q, double* mu_dev, double* var_dev,
                   double* cov_full_dev, void* workspace_dev, size_t workspace_bytes, int64_t q_chunk,
                   void* stream) {
  if (!state_ok(st) || (q > 0 && !Xq_dev) || q_chunk < 1) return BBO_ERR_BAD_ARG;
  return gp_predict(*st, Xq_dev, q, mu_dev, var_dev, cov_full_dev, workspace_dev, workspace_bytes, q_chunk,
                    S(stream));
}

int bbo_acq_eval(const bbo_acq* a, const void* mu_dev, const void* sigma_dev, void* out_dev, int64_t count,
                 int32_t is_f32, void* stream) {
  if (!a || a->kind < BBO_ACQ_EI || a->kind > BBO_ACQ_PI || count < 0) return BBO_ERR_BAD_ARG;
  if (count > 0 && (!mu_dev || !sigma_dev || !out_dev)) return BBO_ERR_BAD_ARG;
  return acq_eval(*a, mu_dev, sigma_dev, out_dev, count, is_f32, S(stream));
}

int bbo_gp_score_pool(const bbo_gp_state* st, const bbo_acq* acq, int32_t precision, const void* Xq_dev,
                      int32_t xq_is_f32, int64_t q, int64_t index_offset, int32_t k, double* topk_score_dev,
                      int64_t* topk_index_dev, double* mu_dev, double* var_dev, double* score_dev,
                      int32_t* nan_count_dev, void* workspace_dev, size_t workspace_bytes, int64_t q_chunk,
                      int32_t accumulate, void* stream) {
  if (!state_ok(st) || !acq || acq->kind < BBO_ACQ_EI || acq->kind > BBO_ACQ_PI) return BBO_ERR_BAD_ARG;
  if ((q > 0 && !Xq_dev) || !topk_score_dev || !topk_index_dev || q_chunk < 1) return BBO_ERR_BAD_ARG;
  return gp_score_pool(*st, *acq, precision, Xq_dev, xq_is_f32, q, index_offset, k, topk_score_dev,
                       topk_index_dev, mu_dev, var_dev, score_dev, nan_count_dev, workspace_dev, workspace_bytes,
                       q_chunk, accumulate != 0, S(stream));
}

int bbo_gp_score_pool_host(const bbo_gp_state* st, const bbo_acq* acq, int32_t precision, const void* Xq_host,
                           int32_t xq_is_f32, int64_t q, int64_t index_offset, int32_t k,
                           double* topk_score_host, int64_t* topk_index_host, void* staging_dev,
                           double* topk_score_dev, int64_t* topk_index_dev, void* workspace_dev,
                           size_t workspace_bytes, int64_t q_chunk, void* stream) {
  if (!state_ok(st) || !acq || acq->kind < BBO_ACQ_EI || acq->kind > BBO_ACQ_PI) return BBO_ERR_BAD_ARG;
  if ((q > 0 && !Xq_host) || !topk_score_host || !topk_index_host || !staging_dev || !topk_score_dev ||
      !topk_index_dev || q_chunk < 1 || q < 0)
    return BBO_ERR_BAD_ARG;
  cudaStream_t cs = S(stream);
  // copy stream and hand-over events: created once per device (creating and destroying them per call costs a few
  // tens of microseconds and a driver lock per suggestion)
  struct HostAux {
    cudaStream_t copy = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  };
  static HostAux g_host_aux[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return BBO_ERR_BAD_ARG;
  HostAux& ha = g_host_aux[dev];
  // one host-pool call per device at a time: the staging hand-over events are per device
  static std::mutex g_host_mutex[64];
  std::lock_guard<std::mutex> host_lock(g_host_mutex[dev]);
  // one staging buffer = BBO_HOST_STAGE_CHUNKS scoring chunks, so that the K*/panel pipeline inside
  // gp_score_pool has several chunks to overlap
  const int64_t qp = BBO_HOST_STAGE_CHUNKS * (q_chunk < 128 ? 128 : (q_chunk + 127) / 128 * 128);
  const size_t esz = xq_is_f32 ? sizeof(float) : sizeof(double);
  const size_t row = size_t(st->h.d) * esz;
  int rc = BBO_OK;
#define BBO_HOST_CHECK(expr)                          \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) {                          \
      rc = cuda_fail(_e, #expr);                      \
      return rc;                                      \
    }                                                 \
  } while (0)
  if (!ha.copy) {
    BBO_HOST_CHECK(cudaStreamCreateWithFlags(&ha.copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      BBO_HOST_CHECK(cudaEventCreateWithFlags(&ha.ready[i], cudaEventDisableTiming));
      BBO_HOST_CHECK(cudaEventCreateWithFlags(&ha.freed[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t copy = ha.copy;
  cudaEvent_t* ready = ha.ready;
  cudaEvent_t* freed = ha.freed;
  // the staging buffers of an earlier call on another stream may still be in use
  BBO_HOST_CHECK(cudaStreamSynchronize(copy));
  if (q == 0) {
    rc = gp_score_pool(*st, *acq, precision, staging_dev, xq_is_f32, 0, index_offset, k, topk_score_dev,
                       topk_index_dev, nullptr, nullptr, nullptr, nullptr, workspace_dev, workspace_bytes, q_chunk,
                       false, cs);
    if (rc != BBO_OK) return rc;
  }
  // the first transfer is a single chunk, so that the kernels start after 1/BBO_HOST_STAGE_CHUNKS of a staging
  // buffer has crossed PCIe; every later transfer is hidden behind the previous buffer's kernels
  const int64_t q_first = qp / BBO_HOST_STAGE_CHUNKS;
  int64_t ci = 0;
  for (int64_t c0 = 0; c0 < q; ++ci) {
    const int buf = int(ci & 1);
    const int64_t want = ci == 0 ? q_first : qp;
    const int64_t qc = (q - c0 < want) ? q - c0 : want;
    char* dst = static_cast<char*>(staging_dev) + size_t(buf) * size_t(qp) * row;
    if (ci >= 2) BBO_HOST_CHECK(cudaStreamWaitEvent(copy, freed[buf], 0));
    BBO_HOST_CHECK(cudaMemcpyAsync(dst, static_cast<const char*>(Xq_host) + size_t(c0) * row, size_t(qc) * row,
                                   cudaMemcpyHostToDevice, copy));
    BBO_HOST_CHECK(cudaEventRecord(ready[buf], copy));
    BBO_HOST_CHECK(cudaStreamWaitEvent(cs, ready[buf], 0));
    // BBO_PREC_TF32_REFINE: every buffer only extends the short-list (its rows are kept in the workspace);
    // the float64 re-score runs once, after the last buffer
    const int rflags = kRefineDefer | (ci > 0 ? kRefineContinue : 0) | ((ci & 1) ? kRefineParity : 0);
    rc = gp_score_pool(*st, *acq, precision, dst, xq_is_f32, qc, index_offset + c0, k, topk_score_dev,
                       topk_index_dev, nullptr, nullptr, nullptr, nullptr, workspace_dev, workspace_bytes, q_chunk,
                       ci > 0, cs, rflags);
    if (rc != BBO_OK) { cudaStreamSynchronize(copy); return rc; }
    BBO_HOST_CHECK(cudaEventRecord(freed[buf], cs));
    c0 += qc;
  }
  if (precision == BBO_PREC_TF32_REFINE && q > 0) {
    rc = gp_score_pool(*st, *acq, precision, staging_dev, xq_is_f32, 0, index_offset + q, k, topk_score_dev,
                       topk_index_dev, nullptr, nullptr, nullptr, nullptr, workspace_dev, workspace_bytes, q_chunk,
                       false, cs, kRefineContinue | ((ci & 1) ? kRefineParity : 0));
    if (rc != BBO_OK) { cudaStreamSynchronize(copy); return rc; }
  }
  BBO_HOST_CHECK(cudaMemcpyAsync(topk_score_host, topk_score_dev, size_t(k) * sizeof(double),
                                 cudaMemcpyDeviceToHost, cs));
  BBO_HOST_CHECK(cudaMemcpyAsync(topk_index_host, topk_index_dev, size_t(k) * sizeof(int64_t),
                                 cudaMemcpyDeviceToHost, cs));
  BBO_HOST_CHECK(cudaStreamSynchronize(cs));
  BBO_HOST_CHECK(cudaStreamSynchronize(copy));
#undef BBO_HOST_CHECK
  return BBO_OK;
}

int32_t bbo_refine_shortlist(int32_t k) { return refine_shortlist_len(k); }
int bbo_gp_refine_info(const void* workspace_dev, size_t workspace_bytes, int64_t n, int64_t d, int64_t q_chunk,
                       double* info_host, void* stream) {
  if (!workspace_dev || !info_host || n < 1 || d < 1 || q_chunk < 1) return BBO_ERR_BAD_ARG;
  return refine_info(workspace_dev, workspace_bytes, n, d, q_chunk, info_host, S(stream));
}

size_t bbo_gp_score_batched_workspace_bytes(int64_t n, int64_t d, int64_t batch, int64_t q) {
  if (n < 1 || d < 1 || batch < 1 || q < 1) return 0;
  return score_batched_workspace_bytes(n, d, batch, q);
}
int bbo_gp_score_pool_batched(const bbo_gp_hyper* h, int64_t n, int64_t batch, const double* X_dev,
                              const double* hyp_dev, const double* linv_dev, const double* alpha_dev,
                              const bbo_acq* acq_dev, const double* Xq_dev, int64_t q, int32_t k,
                              double* topk_score_dev, int64_t* topk_index_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream) {
  if (!hyper_ok(h)) return BBO_ERR_BAD_SHAPE;
  if (!X_dev || !hyp_dev || !linv_dev || !alpha_dev || !acq_dev || !Xq_dev || !topk_score_dev || !topk_index_dev ||
      !workspace_dev)
    return BBO_ERR_BAD_ARG;
  return gp_score_pool_batched(*h, n, batch, X_dev, hyp_dev, linv_dev, alpha_dev, acq_dev, Xq_dev, q, k,
                               topk_score_dev, topk_index_dev, workspace_dev, workspace_bytes, S(stream));
}

int bbo_topk_merge(const double* score_dev, const int64_t* index_dev, int64_t m, int32_t k,
                   double* out_score_dev, int64_t* out_index_dev, void* stream) {
  if (!score_dev || !index_dev || !out_score_dev || !out_index_dev) return BBO_ERR_BAD_ARG;
  return topk_merge(score_dev, index_dev, m, k, out_score_dev, out_index_dev, S(stream));
}

int bbo_topk_pack(const double* score_dev, const int64_t* index_dev, int32_t k, int64_t* packed_dev, void* stream) {
  if (!score_dev || !index_dev || !packed_dev) return BBO_ERR_BAD_ARG;
  return topk_pack(score_dev, index_dev, k, packed_dev, S(stream));
}
int bbo_topk_merge_packed(const int64_t* packed_dev, int64_t m, int32_t k, double* out_score_dev,
                          int64_t* out_index_dev, void* stream) {
  if (!packed_dev || !out_score_dev || !out_index_dev) return BBO_ERR_BAD_ARG;
  return topk_merge_packed(packed_dev, m, k, out_score_dev, out_index_dev, S(stream));
}

int bbo_warp_kumaraswamy(const void* x_dev, int64_t count, double a, double b, double eps, void* out_dev,
                         int32_t is_f32, void* stream) {
  if (count < 0 || !(a > 0.0) || !(b > 0.0)) return BBO_ERR_BAD_ARG;
  if (count > 0 && (!x_dev || !out_dev)) return BBO_ERR_BAD_ARG;
  return warp_kumar(x_dev, count, a, b, eps, out_dev, is_f32, S(stream));
}

int bbo_pool_uniform(uint64_t seed, int64_t row_offset, int64_t q, int64_t d, void* out_dev, int32_t is_f32,
                     void* stream) {
  if (q < 0 || d < 1 || (q > 0 && !out_dev)) return BBO_ERR_BAD_ARG;
  return pool_uniform(seed, row_offset, q, d, out_dev, is_f32, S(stream));
}

}  // extern "C"

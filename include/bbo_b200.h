/*
 * bbo_b200.h -- C ABI of the B200-native GP-BO hot path (libbbo_b200.so).
 *
 * Drop-in boundary for songlei00/bbo's Gaussian-process surrogate + acquisition
 * scoring.  Every entry point replaces a specific piece of reference Python
 * (paths relative to the reference repo root):
 *
 *   bbo_gp_fit_value_grad   bbo/algorithms/surrogates/gp/objectives.py:34-45 (value)
 *                           + loss.backward() at gp/gp.py:62 (gradient)
 *   bbo_gp_fit_value_grad_inputs  the same + autograd's gradient w.r.t. cov inputs / targets
 *                           (WarpKernel gp/kernels.py:59-68, MLPMean gp/means.py:28-35)
 *   bbo_gp_fit_adam         bbo/algorithms/surrogates/gp/gp.py:54-63 (GP.train loop)
 *   bbo_gp_factorize        bbo/algorithms/surrogates/gp/objectives.py:20-31
 *                           (solve_gp_linear_system, cached at gp/gp.py:66-74)
 *   bbo_gp_append           designers/bo/bo.py:101-106 (the refit after update(): rebuilt state
 *                           for the enlarged data set, hyper-parameters kept)
 *   bbo_gp_predict          bbo/algorithms/surrogates/gp/gp.py:75-83 (GP.predict)
 *   bbo_gp_score_pool       gp/gp.py:75-83 + designers/bo/acquisitions.py:18-31
 *                           + shared/trial.py:278-294 (get_best_trials), i.e. the
 *                           body of BODesigner.score_fn (designers/bo/bo.py:110-115)
 *                           and DesignerAsOptimizer.optimize's selection
 *                           (optimizers/designer_optimizer.py:29-42)
 *   bbo_gp_score_pool_host  same, candidates in HOST memory (H2D overlapped)
 *   bbo_gp_score_pool_batched  same for B independent small problems at once (BASELINE config 4)
 *   bbo_cov_matrix          bbo/algorithms/surrogates/gp/kernel_impl.py:40-128
 *                           (KernelProtocol: rbf/matern52, vmap and cdist forms)
 *   bbo_acq_eval            bbo/algorithms/designers/bo/acquisitions.py:18-31
 *   bbo_topk_merge          bbo/shared/trial.py:285-294 over gathered per-GPU lists
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer marked "dev" is CUDA device
 *     memory owned by the caller; "host" pointers are host memory.  All data buffers and
 *     workspaces are the caller's.  The library itself keeps, per device and created on
 *     first use, only scheduling aids: auxiliary streams and events (fit pipeline, TF32
 *     scoring pipeline, host-pool copies), cached CUDA graphs of the fit iteration and 256 KB of
 *     device memory for the cached work lists of the persistent TF32 panel.  These caches are
 *     guarded by mutexes, so calls from several host threads are safe; calls that target the SAME
 *     device serialise where they share a cache (host-pool scoring, the cached fit graph).
 *   - all matrices row-major, contiguous; float64 unless stated.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it
 *     unless the description says they synchronise.
 *   - return value: BBO_OK or a bbo_status; numerical failures (non-PD matrix,
 *     NaN scores) are reported through device-side `info` words so that nothing
 *     synchronises implicitly:
 *         info[b] == 0            ok
 *         info[b] == j+1  (>0)    leading minor j of K is not positive definite
 *                                 (what makes torch.linalg.cholesky raise at
 *                                 objectives.py:29)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns BBO_ERR_CUDA.
 */
#ifndef BBO_B200_H_
#define BBO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bbo_status {
  BBO_OK = 0,
  BBO_ERR_BAD_SHAPE = 1,
  BBO_ERR_BAD_ARG = 2,
  BBO_ERR_WORKSPACE = 3,   /* workspace too small */
  BBO_ERR_CUDA = 4,        /* CUDA runtime error (see bbo_last_cuda_error) */
  BBO_ERR_NOT_PD = 5,      /* only from the synchronising *_host helpers */
  BBO_ERR_NAN = 6
} bbo_status;

/* kernel_impl.py:29-34 */
typedef enum bbo_kernel_kind { BBO_KERNEL_RBF = 0, BBO_KERNEL_MATERN52 = 1 } bbo_kernel_kind;
/* acquisitions.py:13-31 (+PI, absent from the reference) */
typedef enum bbo_acq_kind { BBO_ACQ_EI = 0, BBO_ACQ_UCB = 1, BBO_ACQ_PI = 2 } bbo_acq_kind;
/* arithmetic of the n^2 panel of the scoring pass.
 *   BBO_PREC_FP64         every candidate in float64 (DMMA)
 *   BBO_PREC_TF32         every candidate on the tcgen05 panel with 10-bit-mantissa operands (TF32's mantissa;
 *                         stored in binary16 containers: correlations (K* divided by s2) and a power-of-two multiple of L^-1)
 *                         and float32 accumulation; scores are TF32-accurate (1e-3 class)
 *   BBO_PREC_TF32_REFINE  the TF32 pass only builds a short-list of bbo_refine_shortlist(k) candidates by TF32
 *                         score; the short-list is re-scored by the float64 kernels and the k best BY FLOAT64
 *                         SCORE are returned, so the returned scores are float64-exact and the selection equals
 *                         BBO_PREC_FP64's whenever the true top-k are inside the short-list (bbo_gp_refine_info
 *                         returns the evidence).  k <= BBO_REFINE_KMAX.                                        */
typedef enum bbo_precision { BBO_PREC_FP64 = 0, BBO_PREC_TF32 = 1, BBO_PREC_TF32_REFINE = 2 } bbo_precision;

#define BBO_ROW_ALIGN 128   /* n is padded to a multiple of this in all n x n buffers */
#define BBO_TOPK_MAX 128
#define BBO_APPEND_MAX 32   /* observations appended per bbo_gp_append call */
#define BBO_REFINE_KMAX 32   /* largest k of BBO_PREC_TF32_REFINE */
#define BBO_HOST_STAGE_CHUNKS 8   /* bbo_gp_score_pool_host copies this many chunks per H2D transfer */

int bbo_version(void);
/* text of the last CUDA error seen by this thread ("" if none) */
const char* bbo_last_cuda_error(void);
/* number of CUDA kernels this library has launched in this process (cumulative) */
int64_t bbo_launch_count(void);
/* Optional device-side timing of library regions with CUDA events recorded on the launching
 * stream.  enable(1) clears the counters; read() synchronises the device, returns how many
 * regions of `tag` ran and their summed duration, and clears that tag.
 * tags: 0 FP64 variance panel (k_vnorm), 1 cross-covariance+mean (k_kstar), 2 acquisition+top-k,
 *       3 covariance build + Cholesky, 4 triangular inverse + solves, 5 LAUUM + gradient (+Adam),
 *       6 TF32 scoring panel.                                                              */
void bbo_profile_enable(int32_t on);
int bbo_profile_read(int32_t tag, int64_t* regions, double* total_ms);
/* n rounded up to BBO_ROW_ALIGN */
int64_t bbo_padded_n(int64_t n);

/* ---- hyper-parameters (constrained values, i.e. after softplus, kernels.py:35,52) */
typedef struct bbo_gp_hyper {
  int32_t kind;           /* bbo_kernel_kind */
  int32_t d;              /* feature dimension */
  double pairwise_eps;    /* 1e-6 for matern52_vmap (torch.nn.PairwiseDistance eps,
                             kernel_impl.py:60,94), 0 for the cdist forms and RBF */
  double noise;           /* fixed noise variance (gp.py:41) */
} bbo_gp_hyper;

/* Device layout of one problem's hyper-parameter block `hyp` (float64[2*d+2]):
 *   [0,d)      lengthscale l_k (isotropic kernels replicate the scalar)
 *   [d,2d)     1 / l_k
 *   [2d]       output variance s^2 (1.0 when there is no ScaleKernel)
 *   [2d+1]     mean constant c (means.py:22)                                     */
static inline int64_t bbo_hyp_len(int64_t d) { return 2 * d + 2; }

/* ---- workspace sizing --------------------------------------------------------- */
/* bytes of device workspace needed by bbo_gp_fit_value_grad / bbo_gp_fit_adam /
 * bbo_gp_factorize for `batch` problems of n observations in d dims.            */
size_t bbo_gp_fit_workspace_bytes(int64_t n, int64_t d, int64_t batch);
/* bytes needed by bbo_gp_predict / bbo_gp_score_pool for chunks of `q_chunk`
 * candidates (q_chunk is rounded up to 128 internally).                          */
size_t bbo_gp_score_workspace_bytes(int64_t n, int64_t d, int64_t q_chunk, int32_t precision);
/* bytes needed by bbo_gp_predict; want_full != 0 when cov_full_dev will be requested.       */
size_t bbo_gp_predict_workspace_bytes(int64_t n, int64_t d, int64_t q_chunk, int32_t want_full);

/* ---- covariance: KernelProtocol(X1, X2, lengthscale) -> (q, n) ------------------
 * out[i*ldo + j] = s2 * k(X1_i, X2_j);  X1 (q,d), X2 (n,d) dev; hyp dev (layout above) */
int bbo_cov_matrix(const bbo_gp_hyper* h, const double* hyp_dev, const double* X1_dev, int64_t q,
                   const double* X2_dev, int64_t n, double* out_dev, int64_t ldo, void* stream);

/* ---- fit: value and gradient of objectives.py:34-45 ---------------------------------
 * For b in [0,batch): X (batch,n,d), y (batch,n), hyp (batch, 2d+2).
 * Outputs (dev): value[b] = -1/2 dY^T K^-1 dY - sum log L_ii - n/2 log 2pi  (this IS
 * +log p(y); the reference's name is misleading),
 * grad[b, 0:d] = d value / d l_k, grad[b,d] = d value / d s2, grad[b,d+1] = d value / d c,
 * info[b] as described above.  Gradients are w.r.t. the constrained values; the raw
 * (pre-softplus) chain rule is applied by bbo_gp_fit_adam or by the caller.           */
int bbo_gp_fit_value_grad(const bbo_gp_hyper* h, int64_t n, int64_t batch, const double* X_dev,
                          const double* y_dev, const double* hyp_dev, double* value_dev,
                          double* grad_dev, int32_t* info_dev, void* workspace_dev,
                          size_t workspace_bytes, void* stream);

/* Same for ONE problem, plus the gradients w.r.t. the inputs and the targets:
 *   gradX (n,d) = d value / d X,  grady (n) = d value / d y = -K^-1 (y - c).
 * These are what autograd needs to train an input warp (WarpKernel with KumarWarp / MLPWarp,
 * gp/kernels.py:59-68, gp/warpers.py:31-65) or a learned mean (MLPMean, gp/means.py:28-35) on top
 * of this library: X is then warp(X) and y is Y - mean(X) (pass c = 0 in hyp).                   */
int bbo_gp_fit_value_grad_inputs(const bbo_gp_hyper* h, int64_t n, const double* X_dev, const double* y_dev,
                                 const double* hyp_dev, double* value_dev, double* grad_dev,
                                 double* gradX_dev, double* grady_dev, int32_t* info_dev,
                                 void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- train: GP.train (gp.py:54-63) entirely on the device ----------------------------
 * theta (batch, P) raw parameters in the reference's Adam order (gp.py:55):
 *   [0] mean constant, [1 .. 1+n_ls) raw lengthscale (n_ls = d for ARD, 1 isotropic),
 *   [1+n_ls] raw output variance (present iff has_scale).  P = 1 + n_ls + has_scale.
 * adam_m, adam_v (batch, P) moment buffers (zero them before the first call);
 * step0 = number of Adam steps already taken (bias correction continues from it).
 * Runs `epochs` iterations of: value/grad -> chain rule through softplus -> Adam
 * (lr, betas 0.9/0.999, eps 1e-8: torch.optim.Adam defaults) minimising
 * sign * value  (sign=+1 reproduces the reference, which minimises +log p; sign=-1
 * minimises the true negative log-likelihood).  values_dev (epochs, batch) receives
 * the value seen at each epoch (may be NULL).  No host synchronisation.
 * A non-PD epoch sets info[b] and freezes that problem's parameters.                   */
int bbo_gp_fit_adam(const bbo_gp_hyper* h, int64_t n, int64_t batch, int32_t ard, int32_t has_scale,
                    const double* X_dev, const double* y_dev, double* theta_dev, double* adam_m_dev,
                    double* adam_v_dev, int64_t step0, int64_t epochs, double lr, double sign,
                    double* values_dev, int32_t* info_dev, void* workspace_dev,
                    size_t workspace_bytes, void* stream);

/* theta (raw) -> hyp (constrained) on the device: softplus (kernels.py:35,52).          */
int bbo_gp_theta_to_hyp(int64_t d, int64_t batch, int32_t ard, int32_t has_scale,
                        const double* theta_dev, double* hyp_dev, void* stream);

/* ---- factorise: solve_gp_linear_system (objectives.py:20-31) ---------------------------
 * For each problem computes, with np = bbo_padded_n(n):
 *   linv  (batch, np, np)  L^-1, lower triangular, zero above the diagonal, identity
 *                          on the padding (L = chol(K + noise I), lower)
 *   alpha (batch, np)      K^-1 (y - c), zero on the padding      (`kinvy`)
 *   chol  (batch, np, np)  L itself (lower triangle valid); may be NULL
 *   logdet(batch)          sum_i log L_ii; may be NULL                                   */
int bbo_gp_factorize(const bbo_gp_hyper* h, int64_t n, int64_t batch, const double* X_dev,
                     const double* y_dev, const double* hyp_dev, double* linv_dev,
                     double* alpha_dev, double* chol_dev, double* logdet_dev, int32_t* info_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- incremental update: append observations to a factorised problem (hyper-parameters kept) ----
 * Replaces the from-scratch refit of designers/bo/bo.py:101-106 between two hyper-parameter fits:
 * O(m n^2) (four passes over L^-1) instead of O(n^3).  Opt-in: the reference has no counterpart.
 *   X (n_old+m, d), y (n_old+m): ALL observations, the m new ones last; 1 <= m <= BBO_APPEND_MAX.
 *   linv (np', np'), np' = bbo_padded_n(n_old+m): on entry L^-1 of the first n_old observations in its
 *        leading block, zero elsewhere except a unit diagonal on rows >= n_old (exactly what
 *        bbo_gp_factorize leaves when np' == bbo_padded_n(n_old); otherwise the caller re-lays it out);
 *        on exit L^-1 of all n_old+m observations (same layout).
 *   alpha (np') is recomputed; logdet (may be NULL) is incremented by the new pivots' logs.
 *   info[0] = j+1 if the leading minor j (j >= n_old) is not positive definite.                      */
size_t bbo_gp_append_workspace_bytes(int64_t n_old, int64_t m);
int bbo_gp_append(const bbo_gp_hyper* h, int64_t n_old, int64_t m, const double* X_dev, const double* y_dev,
                  const double* hyp_dev, double* linv_dev, double* alpha_dev, double* logdet_dev,
                  int32_t* info_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- fitted state used by predict / score (all dev pointers, one problem) ------------- */
typedef struct bbo_gp_state {
  bbo_gp_hyper h;
  int64_t n;              /* observations */
  const double* X;        /* (n, d) */
  const double* hyp;      /* (2d+2) */
  const double* linv;     /* (np, np) from bbo_gp_factorize */
  const double* alpha;    /* (np) */
  const float* linv_tf32; /* bbo_tf32_rows(n) * np floats of storage filled by bbo_gp_make_tf32 (opaque: the
                             10-bit-mantissa copy of linv the tensor-core panel reads); only read when
                             precision != BBO_PREC_FP64 */
} bbo_gp_state;

/* Copy of linv for the TF32 scoring mode, rounded to nearest at TF32's 10-bit mantissa.  The caller provides
 * bbo_tf32_rows(n) * np floats; the library lays out, in the first half, (bbo_tf32_rows(n), np) binary16 values
 * 2^e linv (np rows of data followed by zero rows up to a multiple of 256, so that every TMA box of the
 * tensor-core panel is in bounds; e puts the largest |entry| in [2^13, 2^14)) and right behind them the float
 * words 2^-2e, 2^e.  The layout is private to the library; only the size is part of the contract.           */
int64_t bbo_tf32_rows(int64_t n);
int bbo_gp_make_tf32(int64_t n, const double* linv_dev, float* linv_tf32_dev, void* stream);

/* ---- posterior: GP.predict (gp.py:75-83) ------------------------------------------------
 * Xq (q,d) dev -> mu[q], var[q] = k(x,x) - ||L^-1 k*||^2 + noise  (the diagonal of the
 * reference's (q,q) result == its (q,1,d)-batched result).  If cov_full_dev != NULL the
 * full (q,q) posterior covariance of gp.py:81-82 is written as well (q <= q_chunk).       */
int bbo_gp_predict(const bbo_gp_state* st, const double* Xq_dev, int64_t q, double* mu_dev,
                   double* var_dev, double* cov_full_dev, void* workspace_dev,
                   size_t workspace_bytes, int64_t q_chunk, void* stream);

/* ---- acquisition parameters (acquisitions.py:13-31) -------------------------------------- */
typedef struct bbo_acq {
  int32_t kind;    /* bbo_acq_kind */
  int32_t _pad;
  double best_f;   /* EI / PI incumbent (EI._best_f) */
  double eps;      /* EI._eps, default 1e-6 */
  double beta;     /* UCB.beta, default 1.8 */
} bbo_acq;

/* elementwise acquisition: out[i] = acq(mu[i], sigma[i]); is_f32 selects float/double.    */
int bbo_acq_eval(const bbo_acq* a, const void* mu_dev, const void* sigma_dev, void* out_dev,
                 int64_t count, int32_t is_f32, void* stream);

/* ---- pool scoring + selection ----------------------------------------------------------------
 * Scores q candidates and returns the k best (descending score, ties -> lowest index, the
 * stable order of trial.py:289-294).  Xq: (q,d) dev, float64 when xq_is_f32 == 0 else float32.
 * index_offset is added to every returned index (global index of this shard's first row).
 * topk_score_dev[k] (float64) / topk_index_dev[k] (int64) receive the result; slots beyond
 * the number of non-NaN candidates get score -inf and index -1.
 * Optional full outputs (NULL to skip): mu_dev[q], var_dev[q], score_dev[q] (float64).
 * nan_count_dev (int32, may be NULL) is incremented by the number of NaN scores (the
 * reference raises ValueError on NaN, acquisitions.py:21 via Normal's arg validation).
 * accumulate != 0: topk_* already hold a running top-k from earlier calls (other chunks of
 * the same pool) and are merged with this call's candidates instead of being reset.        */
int bbo_gp_score_pool(const bbo_gp_state* st, const bbo_acq* acq, int32_t precision,
                      const void* Xq_dev, int32_t xq_is_f32, int64_t q, int64_t index_offset,
                      int32_t k, double* topk_score_dev, int64_t* topk_index_dev, double* mu_dev,
                      double* var_dev, double* score_dev, int32_t* nan_count_dev,
                      void* workspace_dev, size_t workspace_bytes, int64_t q_chunk,
                      int32_t accumulate, void* stream);

/* BBO_PREC_TF32_REFINE: size of the short-list kept by the TF32 pass (max(64, 8k), at most 256).        */
int32_t bbo_refine_shortlist(int32_t k);
/* Evidence of the last BBO_PREC_TF32_REFINE call that used this workspace (same n, d, q_chunk as that call);
 * synchronises `stream`.  info_host[6]:
 *   [0] k-th best float64 score returned
 *   [1] upper bound of the TF32 score of every candidate that was NOT re-scored (-inf if all were)
 *   [2] observed TF32 score error AT THE CUT: max |TF32 score - float64 score| over the short-listed candidates
 *       that did not make the returned k (for EI / PI a score error scales with the score, so the top of the list
 *       says nothing about candidates at the cut); [4] when no such candidate exists
 *   [3] number of short-listed candidates
 *   [4] max |TF32 score - float64 score| over the whole short-list
 *   [5] number of candidates behind [2]
 * The selection equals the float64 mode's if no unscored candidate's float64 score can reach [0], i.e. if
 * [0] - [1] exceeds the TF32 score error of the candidates at the cut; callers compare it with a multiple of [2]. */
int bbo_gp_refine_info(const void* workspace_dev, size_t workspace_bytes, int64_t n, int64_t d, int64_t q_chunk,
                       double* info_host, void* stream);

/* Same with the candidate pool in HOST memory (pinned for full overlap).  The pool is
 * streamed to the device in chunks of q_chunk rows on an internal copy stream, double
 * buffered against the scoring kernels; staging_dev must hold
 * 2*BBO_HOST_STAGE_CHUNKS*q_chunk*d elements of the candidate dtype (q_chunk rounded up to 128).  topk_*_host receive the result; the call synchronises `stream`.       */
int bbo_gp_score_pool_host(const bbo_gp_state* st, const bbo_acq* acq, int32_t precision,
                           const void* Xq_host, int32_t xq_is_f32, int64_t q, int64_t index_offset,
                           int32_t k, double* topk_score_host, int64_t* topk_index_host,
                           void* staging_dev, double* topk_score_dev, int64_t* topk_index_dev,
                           void* workspace_dev, size_t workspace_bytes, int64_t q_chunk,
                           void* stream);

/* Batched FP64 pool scoring: `batch` independent problems of one shape (n, d) -- the layout a batched
 * bbo_gp_factorize leaves: X (batch,n,d), hyp (batch,2d+2), linv (batch,np,np), alpha (batch,np) -- each with its
 * own pool Xq (batch,q,d) float64 and its own acquisition parameters acq_dev[batch] (DEVICE array of bbo_acq:
 * the incumbent best_f differs per problem).  One set of launches serves all problems (blockIdx.z = problem).
 * Outputs topk_score (batch,k), topk_index (batch,k), indices local to each pool.  (q/2048 + 1) * k <= 2048.  */
size_t bbo_gp_score_batched_workspace_bytes(int64_t n, int64_t d, int64_t batch, int64_t q);
int bbo_gp_score_pool_batched(const bbo_gp_hyper* h, int64_t n, int64_t batch, const double* X_dev,
                              const double* hyp_dev, const double* linv_dev, const double* alpha_dev,
                              const bbo_acq* acq_dev, const double* Xq_dev, int64_t q, int32_t k,
                              double* topk_score_dev, int64_t* topk_index_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream);

/* Merge m (score,index) pairs (e.g. the all-gathered per-GPU top-k lists) into the k best,
 * same ordering rule.  In-place safe when out != in.                                        */
int bbo_topk_merge(const double* score_dev, const int64_t* index_dev, int64_t m, int32_t k,
                   double* out_score_dev, int64_t* out_index_dev, void* stream);

/* The multi-GPU step: each rank packs its k (score, index) pairs into ONE interleaved
 * int64 buffer packed[2k] = {bits(score_0), index_0, ...}; the caller all-gathers the packed buffers of all ranks
 * with a single ncclAllGather (torch.distributed.all_gather_into_tensor in bbo_b200/distributed.py -- the library
 * does not link NCCL, so that the communicator stays the caller's); bbo_topk_merge_packed reduces the gathered
 * m = world * k pairs to the k best with the ordering rule above.  Two launches around one collective.          */
int bbo_topk_pack(const double* score_dev, const int64_t* index_dev, int32_t k, int64_t* packed_dev, void* stream);
int bbo_topk_merge_packed(const int64_t* packed_dev, int64_t m, int32_t k, double* out_score_dev,
                          int64_t* out_index_dev, void* stream);

/* On-device uniform candidate generator (counter based, Philox4x32-10): fills out (q,d) with
 * U[0,1) values that depend only on (seed, global row = row_offset + i, column), so any
 * sharding of the pool produces the same candidates.  is_f32 selects float/double output.  */
int bbo_pool_uniform(uint64_t seed, int64_t row_offset, int64_t q, int64_t d, void* out_dev,
                     int32_t is_f32, void* stream);

/* Kumaraswamy input warp (gp/warpers.py:52-65) of `count` values, the scoring prologue for a
 * WarpKernel(KumarWarp) covariance: out = 1 - (1 - clip(x, eps, 1-eps)^a)^b with a, b already
 * softplus-transformed; in place allowed (out_dev == x_dev).                                       */
int bbo_warp_kumaraswamy(const void* x_dev, int64_t count, double a, double b, double eps, void* out_dev,
                         int32_t is_f32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BBO_B200_H_ */

// Host-side declarations of the fit path (fit.cu).
#pragma once
#include "common.cuh"

namespace bbo {

// Carves the caller-provided device workspace into the fit buffers (all float64).
struct FitWorkspace {
  int64_t n = 0, d = 0, batch = 0, np = 0;
  int ntile = 0, ntp = 0;     // 64-row tiles of the n real rows; lower tile pairs
  double* L = nullptr;        // (batch, np, np)  K, then its Cholesky factor (lower)
  double* Wlo = nullptr;      // (batch, np, np)  L^-1, lower, zeros above
  double* Wup = nullptr;      // (batch, np, np)  L^-T, upper, zeros below
  double* T = nullptr;        // (batch, np, np)  TRTRI scratch, then K^-1 (lower)
  double* z = nullptr;        // (batch, np)      L^-1 (y - c)
  double* alpha = nullptr;    // (batch, np)      K^-1 (y - c)
  double* logdet = nullptr;   // (batch)          sum log L_ii
  double* value = nullptr;    // (batch)
  double* sum_alpha = nullptr;  // (batch)
  double* gpart = nullptr;    // (batch, ntp, d+1) gradient tile partials
  double* hyp = nullptr;      // (batch, 2d+2)
  int64_t* step_dev = nullptr;  // (batch) Adam step counters (device resident: graph replay)
  double* side = nullptr;     // (batch, 2, 64, 64) un-solved copies of the tile below the diagonal block

  static size_t bytes(int64_t n, int64_t d, int64_t batch);
  int bind(void* ws, size_t ws_bytes, int64_t n, int64_t d, int64_t batch);
};

int fit_factor(const bbo_gp_hyper& h, const FitWorkspace& w, const double* X, const double* y,
               const double* hyp, int32_t* info, cudaStream_t st);
int fit_grad_partials(const bbo_gp_hyper& h, const FitWorkspace& w, const double* X, const double* hyp,
                      cudaStream_t st);
int fit_value_grad(const bbo_gp_hyper& h, const FitWorkspace& w, const double* X, const double* y,
                   const double* hyp_in, double* value, double* grad, int32_t* info, cudaStream_t st);
int fit_adam(const bbo_gp_hyper& h, const FitWorkspace& w, int ard, int has_scale, const double* X,
             const double* y, double* theta, double* am, double* av, int64_t step0, int64_t epochs,
             double lr, double sign, double* values, int32_t* info, cudaStream_t st);
int theta_to_hyp(int64_t d, int64_t batch, int ard, int has_scale, const double* theta, double* hyp,
                 cudaStream_t st);
int fit_grad_inputs(const bbo_gp_hyper& h, const FitWorkspace& w, const double* X, double* gradX, double* grady,
                    cudaStream_t st);
size_t append_workspace_bytes(int64_t n_old, int64_t m);
int gp_append(const bbo_gp_hyper& h, int64_t n, int64_t m, const double* X, const double* y, const double* hyp,
              double* linv, double* alpha, double* logdet, int32_t* info, void* ws, size_t ws_bytes,
              cudaStream_t st);
int cov_matrix(const bbo_gp_hyper& h, const double* hyp, const double* X1, int64_t q, const double* X2,
               int64_t n, double* out, int64_t ldo, double diag_add, cudaStream_t st);

}  // namespace bbo

// host_api.cpp - C entry points of the host-side library (libngsfhmm_host.so):
// the pieces of iter_EM (EM.cpp:139-289) that stay on the CPU - optimiser
// bookkeeping and iteration order - driving the CUDA C ABI.
#include <cmath>
#include <cstring>
#include <vector>

#include "bfgs_driver.hpp"
#include "ngsfhmm_host.h"

using namespace nfh_host;

extern "C" {

double nfh_host_minimize(int n, double *x, nfh_objective fun, const void *data, const double *lower,
                         const double *upper, int *n_evals) {
  if (n < 1 || n > GradientPlan::kMaxDim) return NAN;
  return minimize_with_numeric_gradient(n, x, fun, data, lower, upper, n_evals);
}

int nfh_host_bfgs_update(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, int F_fixed, int alpha_fixed,
                         uint64_t stats_out[3]) {
  BfgsStats st;
  int rc = bfgs_update_lockstep(ctx, n_ind, indF, alpha, F_fixed != 0, alpha_fixed != 0, &st);
  if (stats_out) {
    stats_out[0] = st.rounds; stats_out[1] = st.evaluations; stats_out[2] = st.max_rounds_one_individual;
  }
  return rc;
}

int nfh_host_estep_bfgs_update(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, int F_fixed,
                               int alpha_fixed, double *ind_lkl_out, uint64_t stats_out[3]) {
  std::vector<double> lkl_tmp;
  if (!ind_lkl_out) { lkl_tmp.resize(n_ind); ind_lkl_out = lkl_tmp.data(); }
  BfgsStats st;
  int rc = bfgs_update_lockstep(ctx, n_ind, indF, alpha, F_fixed != 0, alpha_fixed != 0, &st, ind_lkl_out);
  if (stats_out) {
    stats_out[0] = st.rounds; stats_out[1] = st.evaluations; stats_out[2] = st.max_rounds_one_individual;
  }
  return rc;
}

int nfh_host_estep_bfgs_update_hook(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, int F_fixed,
                                    int alpha_fixed, double *ind_lkl_out, uint64_t stats_out[3],
                                    nfh_stage_hook posterior_ready, void *user) {
  std::vector<double> lkl_tmp;
  if (!ind_lkl_out) { lkl_tmp.resize(n_ind); ind_lkl_out = lkl_tmp.data(); }
  BfgsStats st;
  int rc = bfgs_update_lockstep(ctx, n_ind, indF, alpha, F_fixed != 0, alpha_fixed != 0, &st, ind_lkl_out,
                                posterior_ready, user);
  if (stats_out) {
    stats_out[0] = st.rounds; stats_out[1] = st.evaluations; stats_out[2] = st.max_rounds_one_individual;
  }
  return rc;
}

int nfh_host_em_iteration(nfh_ctx *ctx, double *indF, double *alpha, int F_fixed, int alpha_fixed, int freq_est,
                          double *ind_lkl_out, double *freq_out, uint64_t stats_out[3]) {
  const uint64_t n = nfh_n_ind_owned(ctx);
  int rc = nfh_set_ind_params(ctx, indF, alpha);                 // parameters of the previous iteration
  if (rc != NFH_OK) return rc;
  // E-step (EM.cpp:151-185) and F / alpha update (EM.cpp:188-205, old emissions): the optimiser starts at the
  // E-step's parameters, so its first batched round and the E-step share one forward pass
  rc = nfh_host_estep_bfgs_update(ctx, n, indF, alpha, F_fixed, alpha_fixed, ind_lkl_out, stats_out);
  if (rc != NFH_OK) return rc;
  if (freq_est != 0) rc = nfh_freq_update(ctx, 1, 0, freq_out);  // EM.cpp:224-271 (new posterior)
  return rc;
}

}  // extern "C"

/*
 * ngsfhmm_host.h - host-side C entry points (libngsfhmm_host.so) that sit
 * above the CUDA C ABI (ngsfhmm_b200.h): the optimiser bookkeeping and the
 * iteration order of iter_EM, which stay on the CPU as in the reference.
 */
#ifndef NGSFHMM_HOST_H
#define NGSFHMM_HOST_H

#include <stdint.h>
#include "ngsfhmm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* same shape as the reference's objective callback (shared/bfgs.h:58-61) */
typedef double (*nfh_objective)(const double *x, const void *data);

/* Replaces findmax_bfgs (shared/bfgs.cpp:83-138) with numerical gradient
 * (getgradient/Yanggradient :22-65), nbd = 2 for every coordinate, n <= 2.
 * x is updated in place; returns -f like the reference. */
double nfh_host_minimize(int n, double *x, nfh_objective fun, const void *data, const double *lower,
                         const double *upper, int *n_evals);

/* Replaces the type-4 tasks of iter_EM (EM.cpp:198-201, 423-441): all owned
 * individuals' (F, alpha) optimised in lockstep around nfh_lkl_batch.
 * stats_out = {batched rounds, objective evaluations, max rounds of one individual}. */
int nfh_host_bfgs_update(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, int F_fixed, int alpha_fixed,
                         uint64_t stats_out[3]);

/* E-step + F / alpha update of one EM iteration (EM.cpp:151-205) for the parameters last set with
 * nfh_set_ind_params: the optimiser's first batched round carries the E-step (nfh_estep_with_batch).
 * ind_lkl_out[n_ind] = the E-step's log-likelihoods; indF / alpha are updated in place. */
int nfh_host_estep_bfgs_update(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, int F_fixed,
                               int alpha_fixed, double *ind_lkl_out, uint64_t stats_out[3]);

/* Replaces iter_EM (EM.cpp:139-289) on one rank: E-step with the given
 * parameters, F/alpha update against the old emissions, frequency update +
 * emission refresh with the new posteriors.  indF/alpha in/out [n_ind_owned];
 * ind_lkl_out [n_ind_owned]; freq_out [sites_owned] or NULL. */
int nfh_host_em_iteration(nfh_ctx *ctx, double *indF, double *alpha, int F_fixed, int alpha_fixed, int freq_est,
                          double *ind_lkl_out, double *freq_out, uint64_t stats_out[3]);

#ifdef __cplusplus
}
#endif
#endif

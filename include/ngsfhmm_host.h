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

/* The same with a hook for multi-rank callers: posterior_ready(user) is called once, from the calling thread, as soon
 * as the posteriors of this E-step are complete in NFH_WIN_POST_SEND - after the first optimiser round; the rounds
 * that follow only read the emission window - so that their exchange to the frequency side (EM.cpp:224-271 needs
 * every individual's posterior at a site) can run behind the rest of the optimisation. */
typedef void (*nfh_stage_hook)(void *user);
int nfh_host_estep_bfgs_update_hook(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, int F_fixed,
                                    int alpha_fixed, double *ind_lkl_out, uint64_t stats_out[3],
                                    nfh_stage_hook posterior_ready, void *user);

/* Replaces iter_EM (EM.cpp:139-289) on one rank: E-step with the given
 * parameters, F/alpha update against the old emissions, frequency update +
 * emission refresh with the new posteriors.  indF/alpha in/out [n_ind_owned];
 * ind_lkl_out [n_ind_owned]; freq_out [sites_owned] or NULL. */
int nfh_host_em_iteration(nfh_ctx *ctx, double *indF, double *alpha, int F_fixed, int alpha_fixed, int freq_est,
                          double *ind_lkl_out, double *freq_out, uint64_t stats_out[3]);

/* ---------------------------------------------------------------------------
 * One process, several GPUs.  Replaces main() -> EM() -> iter_EM() (ngsF-HMM.cpp:27-171, EM.cpp:27-135,
 * EM.cpp:139-289) when the individuals are sharded over the GPUs of one box: a group owns one context per
 * rank (devices[r]; the same ordinal may repeat, which runs the multi-rank geometry on one GPU), one host
 * thread per rank drives its context, the stages are ordered by host joins.  All host arrays are GLOBAL:
 * indF / alpha / ind_lkl [n_ind_total], freq [n_sites], path and posterior [n_ind_total][n_sites],
 * log_gl site-major [n][n_ind_total][3] as in the input file.
 * fused_exchange != 0: the E-step and frequency kernels store straight into the windows of the rank that owns
 * the data next (peer memory, NVLink between devices); 0: block copies between the windows.
 * ------------------------------------------------------------------------- */
typedef struct nfh_group nfh_group;
int nfh_group_create(nfh_group **out, int n_ranks, const int *devices, uint64_t n_ind_total, uint64_t n_sites,
                     int fused_exchange);
void nfh_group_destroy(nfh_group *g);
const char *nfh_group_last_error(const nfh_group *g);
int nfh_group_size(const nfh_group *g);
nfh_ctx *nfh_group_ctx(nfh_group *g, int rank);
int nfh_group_upload_gl(nfh_group *g, const double *log_gl, uint64_t first_site, uint64_t n);
int nfh_group_upload_pos_dist(nfh_group *g, const double *dist_mb);
int nfh_group_set_freq(nfh_group *g, const double *freq);
int nfh_group_set_ind_params(nfh_group *g, const double *indF, const double *alpha);
/* calc_emission over everything (+ e0 for Viterbi), exchanged and reduced across the ranks */
int nfh_group_refresh_emissions(nfh_group *g, int with_e0);
/* "--freq e": est_maf with F = 0 for everyone (parse_args.cpp:316-318) + emission refresh */
int nfh_group_freq_init(nfh_group *g, double *freq_out);
/* iter_EM (EM.cpp:139-289) over all ranks; stats_out = {max rounds, total evaluations, max rounds of one individual} */
int nfh_group_em_iteration(nfh_group *g, double *indF, double *alpha, int F_fixed, int alpha_fixed, int freq_est,
                           double *ind_lkl_out, double *freq_out, uint64_t stats_out[3]);
int nfh_group_estep(nfh_group *g, double *ind_lkl_out);
int nfh_group_viterbi(nfh_group *g, char *path_out);
int nfh_group_get_posterior(nfh_group *g, double *marg1_out);
int nfh_group_get_freq(nfh_group *g, double *freq_out);
int nfh_group_geno_posterior(nfh_group *g, const char *path_all, double *geno_out);

#ifdef __cplusplus
}
#endif
#endif

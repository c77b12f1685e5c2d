/*
 * fake_group_oracle.c - TEST INFRASTRUCTURE ONLY (see fake_device_oracle.c).
 *
 * The nfh_group_* entry points (include/ngsfhmm_host.h) on ONE fake context whose arithmetic is the oracle's.
 * tests/test_cli_on_oracle.py links the PRODUCT's command-line sources (host/cli/ *.cpp) and host library
 * sources (lbfgsb, bfgs_driver, host_api) against this file and fake_device_oracle.c: a CPU-only build of the
 * drop-in binary in which only the device arithmetic is replaced - by the reference's own.  Its output files must
 * then equal the reference binary's BYTE FOR BYTE for every input mode and flag combination, which pins the whole
 * host side of the binary (readers, start values, iteration control, optimiser, output writer) without a GPU.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ngsfhmm_host.h"
#include "ngsfhmm_oracle.h"

/* the fake context of fake_device_oracle.c */
struct nfh_ctx {
  uint64_t N, S;
  double *gl, *dist, *freq, *e_prob, *marg1, *indF, *alpha;
  uint64_t n_estep, n_batch, n_freq;
};

struct nfh_group { nfh_ctx *c; };

int nfh_device_count(void) { return 1; }
int nfh_host_register(nfh_ctx *ctx, void *ptr, uint64_t bytes) { (void) ctx; (void) ptr; (void) bytes; return NFH_ERR_ARG; }
int nfh_host_unregister(nfh_ctx *ctx, void *ptr) { (void) ctx; (void) ptr; return NFH_OK; }

int nfh_group_create(nfh_group **out, int n_ranks, const int *devices, uint64_t N, uint64_t S, int fused) {
  (void) n_ranks; (void) devices; (void) fused;
  nfh_group *g = (nfh_group *) calloc(1, sizeof *g);
  nfh_ctx *c = g->c = (nfh_ctx *) calloc(1, sizeof *c);
  c->N = N; c->S = S;
  c->gl = (double *) calloc(N * S * 3, sizeof(double));
  c->dist = (double *) calloc(S, sizeof(double));
  c->freq = (double *) calloc(S, sizeof(double));
  c->e_prob = (double *) calloc(N * S * 2, sizeof(double));
  c->marg1 = (double *) calloc(N * S, sizeof(double));
  c->indF = (double *) calloc(N, sizeof(double));
  c->alpha = (double *) calloc(N, sizeof(double));
  *out = g;
  return NFH_OK;
}

void nfh_group_destroy(nfh_group *g) {
  if (!g) return;
  nfh_ctx *c = g->c;
  free(c->gl); free(c->dist); free(c->freq); free(c->e_prob); free(c->marg1); free(c->indF); free(c->alpha);
  free(c); free(g);
}

const char *nfh_group_last_error(const nfh_group *g) { (void) g; return "fake device"; }
int nfh_group_size(const nfh_group *g) { (void) g; return 1; }
nfh_ctx *nfh_group_ctx(nfh_group *g, int rank) { (void) rank; return g->c; }

/* log_gl: the reader's output, already normalised as read_geno + main do: taken as it is */
int nfh_group_upload_gl(nfh_group *g, const double *log_gl, uint64_t first_site, uint64_t n) {
  nfh_ctx *c = g->c;
  for (uint64_t s = 0; s < n; s++)
    for (uint64_t i = 0; i < c->N; i++)
      memcpy(c->gl + (i * c->S + first_site + s) * 3, log_gl + (s * c->N + i) * 3, 3 * sizeof(double));
  return NFH_OK;
}
int nfh_group_upload_pos_dist(nfh_group *g, const double *d) { memcpy(g->c->dist, d, g->c->S * sizeof(double)); return NFH_OK; }
int nfh_group_set_freq(nfh_group *g, const double *f) { memcpy(g->c->freq, f, g->c->S * sizeof(double)); return NFH_OK; }
int nfh_group_get_freq(nfh_group *g, double *f) { memcpy(f, g->c->freq, g->c->S * sizeof(double)); return NFH_OK; }
int nfh_group_set_ind_params(nfh_group *g, const double *F, const double *a) { return nfh_set_ind_params(g->c, F, a); }

int nfh_group_refresh_emissions(nfh_group *g, int with_e0) {
  (void) with_e0;
  nfh_ctx *c = g->c;
  orc_freq_emission(c->N, c->S, c->gl, NULL, 0, c->freq, c->e_prob);
  return NFH_OK;
}

int nfh_group_freq_init(nfh_group *g, double *freq_out) { return nfh_freq_update(g->c, 1, 1, freq_out); }

int nfh_group_em_iteration(nfh_group *g, double *indF, double *alpha, int F_fixed, int alpha_fixed, int freq_est,
                           double *ind_lkl_out, double *freq_out, uint64_t stats_out[3]) {
  /* the product's own one-rank iteration (host_api.cpp) */
  return nfh_host_em_iteration(g->c, indF, alpha, F_fixed, alpha_fixed, freq_est, ind_lkl_out, freq_out, stats_out);
}

int nfh_group_estep(nfh_group *g, double *ind_lkl_out) { return nfh_estep(g->c, ind_lkl_out); }

int nfh_group_viterbi(nfh_group *g, char *path_out) {
  nfh_ctx *c = g->c;
  for (uint64_t i = 0; i < c->N; i++)
    orc_viterbi(c->S, c->e_prob + i * c->S * 2, c->dist, c->indF[i], c->alpha[i], path_out + i * c->S);
  return NFH_OK;
}

int nfh_group_get_posterior(nfh_group *g, double *marg1_out) {
  memcpy(marg1_out, g->c->marg1, g->c->N * g->c->S * sizeof(double));
  return NFH_OK;
}

/* EM.cpp:369-376: HWE prior with F = Viterbi state, posterior in log space, exp */
int nfh_group_geno_posterior(nfh_group *g, const char *path_all, double *geno_out) {
  nfh_ctx *c = g->c;
  for (uint64_t s = 0; s < c->S; s++)
    for (uint64_t i = 0; i < c->N; i++) {
      double prior[3], pp[3];
      orc_calc_HWE(prior, c->freq[s], (double) path_all[i * c->S + s], 1);
      orc_post_prob(pp, c->gl + (i * c->S + s) * 3, prior);
      for (int k = 0; k < 3; k++) geno_out[(s * c->N + i) * 3 + k] = exp(pp[k]);
    }
  return NFH_OK;
}

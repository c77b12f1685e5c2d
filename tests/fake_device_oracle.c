/*
 * fake_device_oracle.c - TEST INFRASTRUCTURE ONLY.
 *
 * A stand-in for the part of the CUDA C ABI (include/ngsfhmm_b200.h) that the host library's EM iteration
 * calls, implemented on the CPU with the oracle (oracle/ngsfhmm_oracle.h, bit-identical to the reference).
 * tests/test_host_iteration_cpu.py links the PRODUCT's host sources (host/lbfgsb.cpp, bfgs_driver.cpp,
 * host_api.cpp, compiled as they are) against this file instead of libngsfhmm_b200.so and runs
 * nfh_host_em_iteration: with the device arithmetic replaced by the reference's own, everything the host side
 * contributes - the lockstep optimiser, the order of E-step, F/alpha update and frequency update, what is
 * evaluated at which parameters - must reproduce the reference's iter_EM (EM.cpp:139-289) to the last bit.
 * Never linked into anything that ships.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ngsfhmm_b200.h"
#include "ngsfhmm_oracle.h"

struct nfh_ctx {
  uint64_t N, S;
  double *gl;      /* N x S x 3, individual-major, normalised log GL */
  double *dist;    /* S */
  double *freq;    /* S */
  double *e_prob;  /* N x S x 2 */
  double *marg1;   /* N x S */
  double *indF, *alpha;
  uint64_t n_estep, n_batch, n_freq;   /* calls seen, for the test */
};

/* gl_site_major: S x N x 3 as nfh_upload_gl takes it */
nfh_ctx *fake_ctx_create(uint64_t N, uint64_t S, const double *gl_site_major, const double *dist, const double *freq) {
  nfh_ctx *c = (nfh_ctx *) calloc(1, sizeof *c);
  c->N = N; c->S = S;
  c->gl = (double *) malloc(N * S * 3 * sizeof(double));
  for (uint64_t s = 0; s < S; s++)
    for (uint64_t i = 0; i < N; i++)
      memcpy(c->gl + (i * S + s) * 3, gl_site_major + (s * N + i) * 3, 3 * sizeof(double));
  orc_normalize_gl(N * S, c->gl);
  c->dist = (double *) malloc(S * sizeof(double)); memcpy(c->dist, dist, S * sizeof(double));
  c->freq = (double *) malloc(S * sizeof(double)); memcpy(c->freq, freq, S * sizeof(double));
  c->e_prob = (double *) malloc(N * S * 2 * sizeof(double));
  c->marg1 = (double *) calloc(N * S, sizeof(double));
  c->indF = (double *) calloc(N, sizeof(double));
  c->alpha = (double *) calloc(N, sizeof(double));
  orc_freq_emission(N, S, c->gl, NULL, 0, c->freq, c->e_prob);          /* nfh_emission_refresh */
  return c;
}

void fake_ctx_destroy(nfh_ctx *c) {
  if (!c) return;
  free(c->gl); free(c->dist); free(c->freq); free(c->e_prob); free(c->marg1); free(c->indF); free(c->alpha);
  free(c);
}

void fake_ctx_counts(const nfh_ctx *c, uint64_t out[3]) { out[0] = c->n_estep; out[1] = c->n_batch; out[2] = c->n_freq; }
void fake_ctx_get(const nfh_ctx *c, double *marg1, double *e_prob) {
  if (marg1) memcpy(marg1, c->marg1, c->N * c->S * sizeof(double));
  if (e_prob) memcpy(e_prob, c->e_prob, c->N * c->S * 2 * sizeof(double));
}

const char *nfh_last_error(const nfh_ctx *ctx) { (void) ctx; return "fake device"; }
const char *nfh_strerror(int status) { (void) status; return "fake device"; }
uint64_t nfh_n_ind_owned(const nfh_ctx *ctx) { return ctx->N; }

int nfh_set_ind_params(nfh_ctx *c, const double *indF, const double *alpha) {
  memcpy(c->indF, indF, c->N * sizeof(double));
  memcpy(c->alpha, alpha, c->N * sizeof(double));
  return NFH_OK;
}

int nfh_estep(nfh_ctx *c, double *ind_lkl_out) {
  double *lk = (double *) malloc(c->N * sizeof(double));
  const int st = orc_estep(c->N, c->S, c->e_prob, c->dist, c->indF, c->alpha, c->marg1, lk);
  if (ind_lkl_out) memcpy(ind_lkl_out, lk, c->N * sizeof(double));
  free(lk);
  c->n_estep++;
  return st == 0 ? NFH_OK : st == 1 ? NFH_ERR_FWBW : NFH_ERR_NAN;
}

int nfh_lkl_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                  double *neg_lkl_out) {
  for (uint64_t q = 0; q < n_req; q++)
    neg_lkl_out[q] = orc_lkl(c->S, c->e_prob + (uint64_t) ind[q] * c->S * 2, c->dist, F[q], alpha[q]);
  c->n_batch++;
  return NFH_OK;
}

int nfh_estep_with_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                         double *neg_lkl_out, double *ind_lkl_out) {
  /* the contract of the real entry point: every individual's first request is its current (F, alpha) */
  uint64_t q = 0;
  for (uint64_t i = 0; i < c->N; i++) {
    if (q >= n_req || (uint64_t) ind[q] != i || F[q] != c->indF[i] || alpha[q] != c->alpha[i]) return NFH_ERR_ARG;
    while (q < n_req && (uint64_t) ind[q] == i) q++;
  }
  if (q != n_req) return NFH_ERR_ARG;
  int rc = nfh_lkl_batch(c, n_req, ind, F, alpha, neg_lkl_out);
  return rc != NFH_OK ? rc : nfh_estep(c, ind_lkl_out);
}

int nfh_freq_update(nfh_ctx *c, int method, int posterior_is_zero, double *freq_out) {
  double *post = c->marg1, *zero = NULL;
  if (posterior_is_zero) post = zero = (double *) calloc(c->N * c->S, sizeof(double));
  orc_freq_emission(c->N, c->S, c->gl, post, method != 0, c->freq, c->e_prob);
  free(zero);
  if (freq_out) memcpy(freq_out, c->freq, c->S * sizeof(double));
  c->n_freq++;
  return NFH_OK;
}

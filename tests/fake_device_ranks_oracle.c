/*
 * fake_device_ranks_oracle.c - TEST INFRASTRUCTURE ONLY.
 *
 * The context-level C ABI (include/ngsfhmm_b200.h) with its multi-rank geometry - individuals sharded for the
 * recursions, sites sharded for the frequency update, exchange windows [rank][local individual][site in block],
 * peer-direct stores - implemented on the CPU with the oracle's arithmetic (oracle/ngsfhmm_oracle.h, bit-identical
 * to the reference).  tests/test_cli_on_oracle.py and tests/test_group_cpu.py link the PRODUCT's host sources
 * (host/group.cpp, host_api.cpp, bfgs_driver.cpp, lbfgsb.cpp and the command line) against this file instead of
 * libngsfhmm_b200.so.  est_maf sums over the individuals in index order at the owner of the site and every
 * recursion is per individual, so a run over R ranks must equal the reference TO THE LAST BIT whatever R and
 * whichever exchange mode - which pins the host-side multi-rank orchestration without a GPU.
 *
 * Differences from the device library that do not matter to the host: emission windows hold the reference's
 * (log e0, log e1) pairs instead of the ratio (16 instead of 8 bytes per individual-site, so NFH_WIN_LOGE0_SUM
 * stays zero and the E0 windows are empty), site blocks are not rounded to tiles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ngsfhmm_b200.h"
#include "ngsfhmm_oracle.h"

#define MAXR 8

struct nfh_ctx {
  int R, r, direct, post_in_peers;
  uint64_t N, S, n_loc, ind_begin, n_owned, Sb, site_begin, sites_owned;
  double *gl;        /* frequency side: [N][sites_owned][3] */
  double *freq;      /* [sites_owned] */
  double *dist;      /* [S] */
  double *indF, *alpha;                       /* [n_loc] */
  double *post_send, *post_recv;              /* [R][n_loc][Sb] */
  double *emis_send, *emis_recv;              /* [R][n_loc][Sb][2] */
  double *loge0;                              /* [R * n_loc], stays zero */
  nfh_ctx *peer_post[MAXR], *peer_emis[MAXR];
};

static uint64_t cdiv(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static uint64_t umin(uint64_t a, uint64_t b) { return a < b ? a : b; }

const char *nfh_last_error(const nfh_ctx *ctx) { (void) ctx; return "fake device"; }
const char *nfh_strerror(int status) { (void) status; return "fake device"; }
int nfh_device_count(void) { return MAXR; }
int nfh_host_register(nfh_ctx *c, void *p, uint64_t b) { (void) c; (void) p; (void) b; return NFH_ERR_ARG; }
int nfh_host_unregister(nfh_ctx *c, void *p) { (void) c; (void) p; return NFH_OK; }
int nfh_sync(nfh_ctx *c) { (void) c; return NFH_OK; }

int nfh_ctx_create(nfh_ctx **out, int device, uint64_t N, uint64_t S, int R, int r) {
  (void) device;
  if (R < 1 || R > MAXR || r < 0 || r >= R || N > 65535) return NFH_ERR_ARG;
  nfh_ctx *c = (nfh_ctx *) calloc(1, sizeof *c);
  c->R = R; c->r = r; c->N = N; c->S = S;
  c->n_loc = cdiv(N, (uint64_t) R);
  c->ind_begin = umin((uint64_t) r * c->n_loc, N);
  c->n_owned = umin(c->n_loc, N - c->ind_begin);
  c->Sb = cdiv(S, (uint64_t) R);
  c->site_begin = umin((uint64_t) r * c->Sb, S);
  c->sites_owned = umin(c->Sb, S - c->site_begin);
  const size_t win = (size_t) R * c->n_loc * c->Sb;
  c->gl = (double *) calloc(N * (c->sites_owned ? c->sites_owned : 1) * 3, sizeof(double));
  c->freq = (double *) calloc(c->sites_owned ? c->sites_owned : 1, sizeof(double));
  c->dist = (double *) calloc(S, sizeof(double));
  c->indF = (double *) calloc(c->n_loc, sizeof(double));
  c->alpha = (double *) calloc(c->n_loc, sizeof(double));
  c->post_send = (double *) calloc(win, sizeof(double));
  c->emis_send = (double *) calloc(2 * win, sizeof(double));
  /* one rank: the send and the receive window are the same buffer, as in the device library */
  c->post_recv = R == 1 ? c->post_send : (double *) calloc(win, sizeof(double));
  c->emis_recv = R == 1 ? c->emis_send : (double *) calloc(2 * win, sizeof(double));
  c->loge0 = (double *) calloc((size_t) R * c->n_loc, sizeof(double));
  *out = c;
  return NFH_OK;
}

void nfh_ctx_destroy(nfh_ctx *c) {
  if (!c) return;
  free(c->gl); free(c->freq); free(c->dist); free(c->indF); free(c->alpha);
  if (c->R > 1) { free(c->post_recv); free(c->emis_recv); }
  free(c->post_send); free(c->emis_send); free(c->loge0);
  free(c);
}

uint64_t nfh_n_ind_local(const nfh_ctx *c) { return c->n_loc; }
uint64_t nfh_n_ind_owned(const nfh_ctx *c) { return c->n_owned; }
uint64_t nfh_ind_begin(const nfh_ctx *c) { return c->ind_begin; }
uint64_t nfh_site_block(const nfh_ctx *c) { return c->Sb; }
uint64_t nfh_site_begin(const nfh_ctx *c) { return c->site_begin; }
uint64_t nfh_sites_owned(const nfh_ctx *c) { return c->sites_owned; }

/* element (block b, local individual j, site w of the block) of a [R][n_loc][Sb] window */
static size_t at(const nfh_ctx *c, uint64_t b, uint64_t j, uint64_t w) { return ((size_t) b * c->n_loc + j) * c->Sb + w; }

int nfh_upload_gl(nfh_ctx *c, const double *log_gl, uint64_t first_site, uint64_t n) {
  if (first_site < c->site_begin || first_site + n > c->site_begin + c->sites_owned) return NFH_ERR_ARG;
  for (uint64_t s = 0; s < n; s++)
    for (uint64_t i = 0; i < c->N; i++)
      memcpy(c->gl + (i * c->sites_owned + (first_site - c->site_begin) + s) * 3, log_gl + (s * c->N + i) * 3,
             3 * sizeof(double));
  return NFH_OK;
}
int nfh_upload_pos_dist(nfh_ctx *c, const double *d) { memcpy(c->dist, d, c->S * sizeof(double)); return NFH_OK; }
int nfh_set_freq(nfh_ctx *c, const double *f) { memcpy(c->freq, f, c->sites_owned * sizeof(double)); return NFH_OK; }
int nfh_get_freq(nfh_ctx *c, double *f) { memcpy(f, c->freq, c->sites_owned * sizeof(double)); return NFH_OK; }
int nfh_set_ind_params(nfh_ctx *c, const double *F, const double *a) {
  memcpy(c->indF, F, c->n_owned * sizeof(double));
  memcpy(c->alpha, a, c->n_owned * sizeof(double));
  return NFH_OK;
}

int nfh_peer_set(nfh_ctx *c, int window, int peer_rank, nfh_ctx *peer) {
  if (peer_rank < 0 || peer_rank >= c->R) return NFH_ERR_ARG;
  if (window == NFH_WIN_POST_RECV) c->peer_post[peer_rank] = peer;
  else if (window == NFH_WIN_EMIS_RECV) c->peer_emis[peer_rank] = peer;
  else return NFH_ERR_ARG;
  return NFH_OK;
}
int nfh_peer_direct(nfh_ctx *c, int enable) { c->direct = enable; return NFH_OK; }

/* frequency side -> the owners of the individuals: e [N][sites_owned][2] */
static void scatter_emissions(nfh_ctx *c, const double *e) {
  for (uint64_t i = 0; i < c->N; i++) {
    const uint64_t q = i / c->n_loc, j = i % c->n_loc;
    double *dst = c->direct ? c->peer_emis[q]->emis_recv + 2 * at(c, (uint64_t) c->r, j, 0) : c->emis_send + 2 * at(c, q, j, 0);
    memcpy(dst, e + i * c->sites_owned * 2, c->sites_owned * 2 * sizeof(double));
  }
}

int nfh_emission_refresh(nfh_ctx *c, int with_e0) {
  (void) with_e0;
  if (!c->sites_owned) return NFH_OK;
  double *e = (double *) malloc(c->N * c->sites_owned * 2 * sizeof(double));
  orc_freq_emission(c->N, c->sites_owned, c->gl, NULL, 0, c->freq, e);
  scatter_emissions(c, e);
  free(e);
  return NFH_OK;
}

int nfh_freq_update(nfh_ctx *c, int method, int posterior_is_zero, double *freq_out) {
  if (!c->sites_owned) return NFH_OK;
  double *marg = (double *) calloc(c->N * c->sites_owned, sizeof(double));
  if (!posterior_is_zero)
    for (uint64_t i = 0; i < c->N; i++)
      memcpy(marg + i * c->sites_owned, c->post_recv + at(c, i / c->n_loc, i % c->n_loc, 0), c->sites_owned * sizeof(double));
  double *e = (double *) malloc(c->N * c->sites_owned * 2 * sizeof(double));
  orc_freq_emission(c->N, c->sites_owned, c->gl, marg, method != 0, c->freq, e);
  scatter_emissions(c, e);
  free(e); free(marg);
  if (freq_out) memcpy(freq_out, c->freq, c->sites_owned * sizeof(double));
  return NFH_OK;
}

/* recursion side: the S x 2 emissions of local individual j, gathered from the blocks of the receive window */
static double *gather_emissions(const nfh_ctx *c, uint64_t j) {
  double *e = (double *) malloc(c->S * 2 * sizeof(double));
  for (uint64_t b = 0; b < (uint64_t) c->R; b++) {
    const uint64_t s0 = b * c->Sb;
    if (s0 >= c->S) break;
    const uint64_t w = umin(c->Sb, c->S - s0);
    memcpy(e + s0 * 2, c->emis_recv + 2 * at(c, b, j, 0), w * 2 * sizeof(double));
  }
  return e;
}

int nfh_estep(nfh_ctx *c, double *ind_lkl_out) {
  int status = NFH_OK;
  c->post_in_peers = c->direct == 1;
  double *marg = (double *) malloc(c->S * sizeof(double));
  for (uint64_t j = 0; j < c->n_owned; j++) {
    double *e = gather_emissions(c, j), lk = 0.0;
    const int st = orc_estep(1, c->S, e, c->dist, c->indF + j, c->alpha + j, marg, &lk);
    free(e);
    if (st == 1) status = NFH_ERR_FWBW;
    else if (st != 0) status = NFH_ERR_NAN;
    if (ind_lkl_out) ind_lkl_out[j] = lk;
    for (uint64_t b = 0; b < (uint64_t) c->R; b++) {
      const uint64_t s0 = b * c->Sb;
      if (s0 >= c->S) break;
      const uint64_t w = umin(c->Sb, c->S - s0);
      double *dst = c->post_in_peers ? c->peer_post[b]->post_recv + at(c, (uint64_t) c->r, j, 0) : c->post_send + at(c, b, j, 0);
      memcpy(dst, marg + s0, w * sizeof(double));
    }
  }
  free(marg);
  return status;
}

int nfh_lkl_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                  double *neg_lkl_out) {
  double *e = NULL;
  int32_t have = -1;
  for (uint64_t q = 0; q < n_req; q++) {
    if (ind[q] != have) { free(e); e = gather_emissions(c, (uint64_t) ind[q]); have = ind[q]; }
    neg_lkl_out[q] = orc_lkl(c->S, e, c->dist, F[q], alpha[q]);
  }
  free(e);
  return NFH_OK;
}

int nfh_estep_with_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                         double *neg_lkl_out, double *ind_lkl_out) {
  if (c->n_owned == 0) return NFH_OK;
  uint64_t q = 0;   /* the contract of the real entry point: every individual's first request is its current point */
  for (uint64_t i = 0; i < c->n_owned; i++) {
    if (q >= n_req || (uint64_t) ind[q] != i || F[q] != c->indF[i] || alpha[q] != c->alpha[i]) return NFH_ERR_ARG;
    while (q < n_req && (uint64_t) ind[q] == i) q++;
  }
  if (q != n_req) return NFH_ERR_ARG;
  const int rc = nfh_lkl_batch(c, n_req, ind, F, alpha, neg_lkl_out);
  return rc != NFH_OK ? rc : nfh_estep(c, ind_lkl_out);
}

int nfh_viterbi(nfh_ctx *c, char *path_out) {
  for (uint64_t j = 0; j < c->n_owned; j++) {
    double *e = gather_emissions(c, j);
    if (path_out) orc_viterbi(c->S, e, c->dist, c->indF[j], c->alpha[j], path_out + j * c->S);
    free(e);
  }
  return NFH_OK;
}

int nfh_get_posterior(nfh_ctx *c, double *out) {
  for (uint64_t j = 0; j < c->n_owned; j++)
    for (uint64_t b = 0; b < (uint64_t) c->R; b++) {
      const uint64_t s0 = b * c->Sb;
      if (s0 >= c->S) break;
      const uint64_t w = umin(c->Sb, c->S - s0);
      const double *src = c->post_in_peers ? c->peer_post[b]->post_recv + at(c, (uint64_t) c->r, j, 0) : c->post_send + at(c, b, j, 0);
      memcpy(out + j * c->S + s0, src, w * sizeof(double));
    }
  return NFH_OK;
}

/* EM.cpp:369-376: HWE prior with F = Viterbi state, posterior in log space, exp */
int nfh_geno_posterior(nfh_ctx *c, const char *path_all, double *geno_out) {
  for (uint64_t s = 0; s < c->sites_owned; s++)
    for (uint64_t i = 0; i < c->N; i++) {
      double prior[3], pp[3];
      orc_calc_HWE(prior, c->freq[s], (double) path_all[i * c->sites_owned + s], 1);
      orc_post_prob(pp, c->gl + (i * c->sites_owned + s) * 3, prior);
      for (int k = 0; k < 3; k++) geno_out[(s * c->N + i) * 3 + k] = exp(pp[k]);
    }
  return NFH_OK;
}

static double *window(nfh_ctx *c, int w, size_t *elems_per_block) {
  const size_t blk = (size_t) c->n_loc * c->Sb;
  switch (w) {
    case NFH_WIN_POST_SEND: *elems_per_block = blk; return c->post_send;
    case NFH_WIN_POST_RECV: *elems_per_block = blk; return c->post_recv;
    case NFH_WIN_EMIS_SEND: *elems_per_block = 2 * blk; return c->emis_send;
    case NFH_WIN_EMIS_RECV: *elems_per_block = 2 * blk; return c->emis_recv;
    case NFH_WIN_LOGE0_SUM: *elems_per_block = c->n_loc; return c->loge0;
    default: *elems_per_block = 0; return NULL;           /* E0 windows: nothing to move */
  }
}

int nfh_exchange_window(nfh_ctx *c, int w, void **dev_ptr, uint64_t *bytes, uint64_t *bytes_per_peer) {
  size_t per = 0;
  double *p = window(c, w, &per);
  if (dev_ptr) *dev_ptr = p;
  if (bytes) *bytes = (uint64_t) per * c->R * sizeof(double);
  if (bytes_per_peer) *bytes_per_peer = (uint64_t) per * sizeof(double);
  return NFH_OK;
}

int nfh_window_copy_block(nfh_ctx *dst, int dst_window, int dst_block, nfh_ctx *src, int src_window, int src_block) {
  size_t per_d = 0, per_s = 0;
  double *d = window(dst, dst_window, &per_d), *s = window(src, src_window, &per_s);
  if (!d && !s) return NFH_OK;
  if (!d || !s || per_d != per_s) return NFH_ERR_ARG;
  memcpy(d + (size_t) dst_block * per_d, s + (size_t) src_block * per_s, per_s * sizeof(double));
  return NFH_OK;
}

int nfh_window_read(nfh_ctx *c, int w, uint64_t offset, uint64_t bytes, void *host_dst) {
  size_t per = 0;
  double *p = window(c, w, &per);
  if (!p) return NFH_ERR_ARG;
  memcpy(host_dst, (char *) p + offset, bytes);
  return NFH_OK;
}
int nfh_window_write(nfh_ctx *c, int w, uint64_t offset, uint64_t bytes, const void *host_src) {
  size_t per = 0;
  double *p = window(c, w, &per);
  if (!p) return NFH_ERR_ARG;
  memcpy((char *) p + offset, host_src, bytes);
  return NFH_OK;
}

// fake_device_arith.cpp - TEST INFRASTRUCTURE ONLY.
//
// The part of the CUDA C ABI (include/ngsfhmm_b200.h) that the host library's EM iteration calls, implemented on the
// CPU with the KERNELS' OWN ARITHMETIC: tests/device_arith_host.cpp, i.e. the per-thread bodies of the product's
// kernels (nfh_estep_math.cuh, nfh_freq_math.cuh, nfh_viterbi_math.cuh, ...) compiled for the host.  Linked with
// the product's host sources (host/lbfgsb.cpp, bfgs_driver.cpp, host_api.cpp, compiled as they are) it runs whole EM
// iterations and whole EM runs without a GPU: tests/test_em_on_device_arith_cpu.py holds the results against the
// golden fixtures generated from the unmodified reference, at the north star's tolerances.
//
// The sibling tests/fake_device_oracle.c does the same with the REFERENCE's arithmetic (and must then match the
// reference bit for bit); between the two, the CPU suite separates "the host logic is the reference's" from "the
// kernels' arithmetic computes the reference's functions".  Never linked into anything that ships.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ngsfhmm_b200.h"

extern "C" {
int da_estep(uint64_t S, const double *ratio, const double *dist, double F, double alpha, double loge0_sum,
             double *post_out, double *lkl, int *tiers_used);
double da_neg_lkl(uint64_t S, const double *ratio, const double *dist, double F, double alpha, double loge0_sum);
int da_freq_site(int shape, uint64_t n_ind, const double *L0, const double *L1, const double *L2, const double *post,
                 int update_freq, double *freq_io, double *ratio_out, double *e0_out);
void da_viterbi(uint64_t S, const double *ratio, const double *e0_lin, const double *dist, double F, double al,
                unsigned char *path_out);
}

struct nfh_ctx {
  uint64_t N = 0, S = 0;
  int shape = 0;
  std::vector<double> L[3];        // linear GL, site-major [S][N] (what a lane group of the frequency kernel reads)
  std::vector<double> dist, freq;
  std::vector<double> ratio, e0;   // [N][S] e1/e0 and e0
  std::vector<double> loge0;       // [N] sum over sites of log e0
  std::vector<double> post;        // [N][S]
  std::vector<double> indF, alpha;
  uint64_t freq_passes = 0;
};

static void refresh(nfh_ctx *c, bool update) {
  const uint64_t N = c->N, S = c->S;
  std::vector<double> p(N), r(N), e(N);
  for (uint64_t i = 0; i < N; i++) c->loge0[i] = 0.0;
  for (uint64_t s = 0; s < S; s++) {
    for (uint64_t i = 0; i < N; i++) p[i] = c->post[i * S + s];
    const int n = da_freq_site(c->shape, N, &c->L[0][s * N], &c->L[1][s * N], &c->L[2][s * N], p.data(), update ? 1 : 0,
                               &c->freq[s], r.data(), e.data());
    if (update) c->freq_passes += (uint64_t) n;
    for (uint64_t i = 0; i < N; i++) {
      c->ratio[i * S + s] = r[i];
      c->e0[i * S + s] = e[i];
      c->loge0[i] += std::log(e[i]);        // the kernels keep a running product with the exponent split off
    }
  }
}

extern "C" {

// gl_norm_site_major: S x N x 3 normalised log GL, as nfh_upload_gl takes it
nfh_ctx *fake_ctx_create(uint64_t N, uint64_t S, const double *gl_norm_site_major, const double *dist,
                         const double *freq) {
  if (N > 512) return nullptr;
  nfh_ctx *c = new nfh_ctx;
  c->N = N; c->S = S;
  c->shape = N <= 16 ? 4 : N <= 64 ? 1 : N <= 104 ? 0 : N <= 128 ? 2 : 3;
  for (int k = 0; k < 3; k++) {
    c->L[k].resize(S * N);
    for (uint64_t s = 0; s < S; s++)
      for (uint64_t i = 0; i < N; i++) c->L[k][s * N + i] = std::exp(gl_norm_site_major[(s * N + i) * 3 + k]);   // gl_ingest
  }
  c->dist.assign(dist, dist + S);
  c->freq.assign(freq, freq + S);
  c->ratio.assign(N * S, 1.0); c->e0.assign(N * S, 1.0); c->post.assign(N * S, 0.0);
  c->loge0.assign(N, 0.0); c->indF.assign(N, 0.0); c->alpha.assign(N, 0.0);
  refresh(c, false);                                                        // nfh_emission_refresh
  return c;
}
void fake_ctx_destroy(nfh_ctx *c) { delete c; }
void fake_ctx_get(const nfh_ctx *c, double *post, double *freq) {
  if (post) std::memcpy(post, c->post.data(), c->N * c->S * sizeof(double));
  if (freq) std::memcpy(freq, c->freq.data(), c->S * sizeof(double));
}
void fake_ctx_set_freq(nfh_ctx *c, const double *freq) {                    // nfh_set_freq + nfh_emission_refresh
  c->freq.assign(freq, freq + c->S);
  refresh(c, false);
}
uint64_t fake_ctx_freq_passes(const nfh_ctx *c) { return c->freq_passes; }
void fake_ctx_viterbi(const nfh_ctx *c, unsigned char *path) {              // nfh_viterbi: [N][S]
  for (uint64_t i = 0; i < c->N; i++)
    da_viterbi(c->S, &c->ratio[i * c->S], &c->e0[i * c->S], c->dist.data(), c->indF[i], c->alpha[i], path + i * c->S);
}

const char *nfh_last_error(const nfh_ctx *) { return "fake device (kernel arithmetic on the host)"; }
const char *nfh_strerror(int) { return "fake device (kernel arithmetic on the host)"; }
uint64_t nfh_n_ind_owned(const nfh_ctx *ctx) { return ctx->N; }

int nfh_set_ind_params(nfh_ctx *c, const double *indF, const double *alpha) {
  c->indF.assign(indF, indF + c->N);
  c->alpha.assign(alpha, alpha + c->N);
  return NFH_OK;
}

int nfh_estep(nfh_ctx *c, double *ind_lkl_out) {
  int flags = 0;
  for (uint64_t i = 0; i < c->N; i++) {
    double lk[2];
    flags |= da_estep(c->S, &c->ratio[i * c->S], c->dist.data(), c->indF[i], c->alpha[i], c->loge0[i],
                      &c->post[i * c->S], lk, nullptr);
    if (ind_lkl_out) ind_lkl_out[i] = lk[0];
  }
  return (flags & 1) ? NFH_ERR_NAN : (flags & 2) ? NFH_ERR_FWBW : NFH_OK;
}

int nfh_lkl_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                  double *neg_lkl_out) {
  for (uint64_t q = 0; q < n_req; q++) {
    const uint64_t i = (uint64_t) ind[q];
    neg_lkl_out[q] = da_neg_lkl(c->S, &c->ratio[i * c->S], c->dist.data(), F[q], alpha[q], c->loge0[i]);
  }
  return NFH_OK;
}

int nfh_estep_with_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                         double *neg_lkl_out, double *ind_lkl_out) {
  uint64_t q = 0;                          // the real entry point's contract: first request = current (F, alpha)
  for (uint64_t i = 0; i < c->N; i++) {
    if (q >= n_req || (uint64_t) ind[q] != i || F[q] != c->indF[i] || alpha[q] != c->alpha[i]) return NFH_ERR_ARG;
    while (q < n_req && (uint64_t) ind[q] == i) q++;
  }
  if (q != n_req) return NFH_ERR_ARG;
  const int rc = nfh_lkl_batch(c, n_req, ind, F, alpha, neg_lkl_out);
  return rc != NFH_OK ? rc : nfh_estep(c, ind_lkl_out);
}

int nfh_freq_update(nfh_ctx *c, int method, int posterior_is_zero, double *freq_out) {
  std::vector<double> keep;
  if (posterior_is_zero) { keep.swap(c->post); c->post.assign(c->N * c->S, 0.0); }
  refresh(c, method != 0);
  if (posterior_is_zero) c->post.swap(keep);
  if (freq_out) std::memcpy(freq_out, c->freq.data(), c->S * sizeof(double));
  return NFH_OK;
}

}  // extern "C"

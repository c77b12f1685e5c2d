// fake_device_simt.cpp - TEST INFRASTRUCTURE ONLY.
//
// The part of the CUDA C ABI (include/ngsfhmm_b200.h) that the host library's EM iteration calls, implemented with
// the product's KERNELS running under the SIMT emulator (tests/simt/simt.h, tests/simt_kernels_host.cpp): E-step,
// objective batches, the E-step riding on the first batch, Viterbi - and, when asked, the frequency kernels too
// (minutes per thousand sites under an emulator; otherwise the frequency EM runs on the kernels' arithmetic compiled
// for the host, tests/device_arith_host.cpp, which is what tests/fake_device_arith.cpp uses throughout).
// Linked with the product's host sources compiled as they are, it runs tests/test_gpu_golden.py's cases without a GPU:
// the closest the CPU suite gets to the round-end GPU run - same kernels, same launch geometry, same host logic; only
// the hardware is missing.  Never linked into anything that ships.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ngsfhmm_b200.h"

struct SimtCtx;
struct SimtFreq;
extern "C" {
SimtCtx *simt_create(uint64_t N, uint64_t S, const double *ratio, const double *e0, const double *dist, const double *loge0_sum);
void simt_destroy(SimtCtx *c);
void simt_update_emissions(SimtCtx *c, const double *ratio, const double *e0, const double *loge0_sum);
void simt_set_params(SimtCtx *c, const double *F, const double *alpha);
int simt_estep(SimtCtx *c, double *ind_lkl_out, double *post_out);
int simt_lkl_batch(SimtCtx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha, double *neg_lkl_out,
                   int with_estep, double *ind_lkl_out, double *post_out);
void simt_viterbi(SimtCtx *c, unsigned char *path_out);
SimtFreq *simt_freq_create(uint64_t N, uint64_t S, const double *gl_norm_site_major, int sm_count);
void simt_freq_destroy(SimtFreq *c);
void simt_freq_set(SimtFreq *c, const double *post, const double *freq);
int simt_freq_run(SimtFreq *c, int update, int zero_post, int with_e0, int want_maps, int *used_maps, double *freq_out,
                  double *ratio_out, double *e0_out, double *loge0_out, unsigned long long *passes_out);
int da_freq_site(int shape, uint64_t n_ind, const double *L0, const double *L1, const double *L2, const double *post,
                 int update_freq, double *freq_io, double *ratio_out, double *e0_out);
}

struct nfh_ctx {
  uint64_t N = 0, S = 0;
  bool freq_kernels = false;
  SimtCtx *rec = nullptr;
  SimtFreq *frq = nullptr;
  int shape = 0;
  std::vector<double> L[3];                 // linear GL [S][N] for the host-arithmetic frequency path
  std::vector<double> dist, freq, ratio, e0, loge0, post, indF, alpha;
};

static void refresh(nfh_ctx *c, bool update, bool zero_post) {
  const uint64_t N = c->N, S = c->S;
  if (c->freq_kernels) {
    simt_freq_set(c->frq, zero_post ? nullptr : c->post.data(), c->freq.data());
    unsigned long long passes = 0;
    int used = 0;
    simt_freq_run(c->frq, update ? 1 : 0, zero_post ? 1 : 0, 1, 1, &used, c->freq.data(), c->ratio.data(), c->e0.data(),
                  c->loge0.data(), &passes);
    return;
  }
  std::vector<double> p(N, 0.0), r(N), e(N);
  for (uint64_t i = 0; i < N; i++) c->loge0[i] = 0.0;
  for (uint64_t s = 0; s < S; s++) {
    if (!zero_post)
      for (uint64_t i = 0; i < N; i++) p[i] = c->post[i * S + s];
    da_freq_site(c->shape, N, &c->L[0][s * N], &c->L[1][s * N], &c->L[2][s * N], zero_post ? nullptr : p.data(),
                 update ? 1 : 0, &c->freq[s], r.data(), e.data());
    for (uint64_t i = 0; i < N; i++) {
      c->ratio[i * S + s] = r[i];
      c->e0[i * S + s] = e[i];
      c->loge0[i] += std::log(e[i]);
    }
  }
}

extern "C" {

// gl_norm_site_major: S x N x 3 normalised log GL, as nfh_upload_gl takes it; freq_kernels != 0: the frequency EM and
// the emission refresh run on the emulated kernels as well
nfh_ctx *fake_ctx_create(uint64_t N, uint64_t S, const double *gl_norm_site_major, const double *dist, const double *freq,
                         int freq_kernels) {
  if (N > 512) return nullptr;
  nfh_ctx *c = new nfh_ctx;
  c->N = N; c->S = S; c->freq_kernels = freq_kernels != 0;
  c->shape = N <= 16 ? 4 : N <= 64 ? 1 : N <= 104 ? 0 : N <= 128 ? 2 : 3;
  if (c->freq_kernels) {
    c->frq = simt_freq_create(N, S, gl_norm_site_major, 2);
  } else {
    for (int k = 0; k < 3; k++) {
      c->L[k].resize(S * N);
      for (uint64_t s = 0; s < S; s++)
        for (uint64_t i = 0; i < N; i++) c->L[k][s * N + i] = std::exp(gl_norm_site_major[(s * N + i) * 3 + k]);
    }
  }
  c->dist.assign(dist, dist + S);
  c->freq.assign(freq, freq + S);
  c->ratio.assign(N * S, 1.0); c->e0.assign(N * S, 1.0); c->post.assign(N * S, 0.0);
  c->loge0.assign(N, 0.0); c->indF.assign(N, 0.0); c->alpha.assign(N, 0.0);
  refresh(c, false, true);                                                  // nfh_emission_refresh
  c->rec = simt_create(N, S, c->ratio.data(), c->e0.data(), c->dist.data(), c->loge0.data());
  return c;
}
void fake_ctx_destroy(nfh_ctx *c) {
  if (!c) return;
  if (c->rec) simt_destroy(c->rec);
  if (c->frq) simt_freq_destroy(c->frq);
  delete c;
}
void fake_ctx_get(const nfh_ctx *c, double *post, double *freq) {
  if (post) std::memcpy(post, c->post.data(), c->N * c->S * sizeof(double));
  if (freq) std::memcpy(freq, c->freq.data(), c->S * sizeof(double));
}
void fake_ctx_set_freq(nfh_ctx *c, const double *freq) {                    // nfh_set_freq + nfh_emission_refresh
  c->freq.assign(freq, freq + c->S);
  refresh(c, false, true);
  simt_update_emissions(c->rec, c->ratio.data(), c->e0.data(), c->loge0.data());
}
void fake_ctx_viterbi(nfh_ctx *c, unsigned char *path) { simt_viterbi(c->rec, path); }   // after nfh_set_ind_params

const char *nfh_last_error(const nfh_ctx *) { return "fake device (kernels under the SIMT emulator)"; }
const char *nfh_strerror(int) { return "fake device (kernels under the SIMT emulator)"; }
uint64_t nfh_n_ind_owned(const nfh_ctx *ctx) { return ctx->N; }

int nfh_set_ind_params(nfh_ctx *c, const double *indF, const double *alpha) {
  c->indF.assign(indF, indF + c->N);
  c->alpha.assign(alpha, alpha + c->N);
  simt_set_params(c->rec, indF, alpha);
  return NFH_OK;
}

static int status_of(int flags) { return (flags & 1) ? NFH_ERR_NAN : (flags & 2) ? NFH_ERR_FWBW : NFH_OK; }

int nfh_estep(nfh_ctx *c, double *ind_lkl_out) { return status_of(simt_estep(c->rec, ind_lkl_out, c->post.data())); }

int nfh_lkl_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                  double *neg_lkl_out) {
  return simt_lkl_batch(c->rec, n_req, ind, F, alpha, neg_lkl_out, 0, nullptr, nullptr) < 0 ? NFH_ERR_ARG : NFH_OK;
}

int nfh_estep_with_batch(nfh_ctx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                         double *neg_lkl_out, double *ind_lkl_out) {
  uint64_t q = 0;                          // the real entry point's contract: first request = current (F, alpha)
  for (uint64_t i = 0; i < c->N; i++) {
    if (q >= n_req || (uint64_t) ind[q] != i || F[q] != c->indF[i] || alpha[q] != c->alpha[i]) return NFH_ERR_ARG;
    while (q < n_req && (uint64_t) ind[q] == i) q++;
  }
  if (q != n_req) return NFH_ERR_ARG;
  const int flags = simt_lkl_batch(c->rec, n_req, ind, F, alpha, neg_lkl_out, 1, ind_lkl_out, c->post.data());
  return flags < 0 ? NFH_ERR_ARG : status_of(flags);
}

int nfh_freq_update(nfh_ctx *c, int method, int posterior_is_zero, double *freq_out) {
  refresh(c, method != 0, posterior_is_zero != 0);
  simt_update_emissions(c->rec, c->ratio.data(), c->e0.data(), c->loge0.data());
  if (freq_out) std::memcpy(freq_out, c->freq.data(), c->S * sizeof(double));
  return NFH_OK;
}

}  // extern "C"

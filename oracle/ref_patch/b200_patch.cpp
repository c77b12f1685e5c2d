// b200_patch.cpp - TEST INFRASTRUCTURE: INTEGRATION.md section B, compiled against the unmodified reference.
//
// The reference's objects are linked as they are, with two of their symbols made weak by objcopy
// (oracle/Makefile, target `patched`): iter_EM(params*) (EM.cpp:139-289) and viterbi(...) (HMM.cpp:98-125).
// This file supplies the strong definitions, which route both through the C ABI of the B200 hot path; main(),
// argument parsing, input readers, EM()'s iteration control and print_iter() remain the reference's own code.
// Built only where /root/reference exists; the binary (oracle/_ref/ngsF-HMM_b200patch) travels to the GPU box
// and tests/test_gpu_cli.py compares its output files with those of the unmodified reference binary.
#include <cstring>
#include <mutex>
#include <vector>

#include "ngsF-HMM.hpp"
#include "ngsfhmm_b200.h"
#include "ngsfhmm_host.h"

namespace {

nfh_ctx *g_ctx = nullptr;
params *g_pars = nullptr;
std::once_flag g_decoded;
std::vector<char> g_paths;

void fail_if(int rc, const char *where) {
  if (rc == NFH_ERR_NAN) error("forward", "invalid Lkl found!");                  // HMM.cpp:18-21
  if (rc == NFH_ERR_FWBW) error(where, "Fw and Bw lkl do not match!");            // EM.cpp:166-170
  if (rc) error(where, nfh_last_error(g_ctx));
}

// context, GL (1-based [i][s][g] -> site-major), distances, start frequencies, emissions
void attach(params *pars) {
  g_pars = pars;
  if (nfh_ctx_create(&g_ctx, 0, pars->n_ind, pars->n_sites, 1, 0)) error(__FUNCTION__, nfh_last_error(NULL));
  std::vector<double> gl(pars->n_sites * pars->n_ind * 3);
  for (uint64_t s = 1; s <= pars->n_sites; s++)
    for (uint64_t i = 0; i < pars->n_ind; i++)
      memcpy(&gl[((s - 1) * pars->n_ind + i) * 3], pars->geno_lkl[i][s], 3 * sizeof(double));
  fail_if(nfh_upload_gl(g_ctx, gl.data(), 0, pars->n_sites), "nfh_upload_gl");
  fail_if(nfh_upload_pos_dist(g_ctx, pars->pos_dist + 1), "nfh_upload_pos_dist");
  fail_if(nfh_set_freq(g_ctx, pars->freq + 1), "nfh_set_freq");
  fail_if(nfh_emission_refresh(g_ctx, 0), "nfh_emission_refresh");
}

}  // namespace

// the body of iter_EM is one call; print_iter may run after any iteration (--log), so the
// posteriors it prints are fetched every time
void iter_EM(params *pars) {
  if (!g_ctx) attach(pars);
  uint64_t stats[3];
  fail_if(nfh_host_em_iteration(g_ctx, pars->indF, pars->alpha, pars->indF_fixed, pars->alpha_fixed, pars->freq_est,
                                pars->ind_lkl, pars->freq + 1, stats),
          __FUNCTION__);
  std::vector<double> marg1(pars->n_ind * pars->n_sites);
  fail_if(nfh_get_posterior(g_ctx, marg1.data()), "nfh_get_posterior");
  for (uint64_t i = 0; i < pars->n_ind; i++)
    for (uint64_t s = 1; s <= pars->n_sites; s++) {
      pars->marg_prob[i][s][1] = marg1[i * pars->n_sites + s - 1];
      pars->marg_prob[i][s][0] = 1.0 - pars->marg_prob[i][s][1];
    }
}

// EM() queues one viterbi task per individual (EM.cpp:110-116).  The first task to run decodes
// every individual on the device with the final parameters; each task then copies its own row (the individual is
// recognised by its path array).
double viterbi(double **, double *, double, double **, char *path, double *, uint64_t length, int) {
  params *pars = g_pars;
  std::call_once(g_decoded, [&]() {
    g_paths.resize(pars->n_ind * pars->n_sites);
    fail_if(nfh_set_ind_params(g_ctx, pars->indF, pars->alpha), "nfh_set_ind_params");
    fail_if(nfh_emission_refresh(g_ctx, 1), "nfh_emission_refresh");
    fail_if(nfh_viterbi(g_ctx, g_paths.data()), "nfh_viterbi");
  });
  for (uint64_t i = 0; i < pars->n_ind; i++)
    if (pars->path[i] == path) {
      for (uint64_t s = 1; s <= length; s++) path[s] = g_paths[i * pars->n_sites + s - 1];
      path[0] = path[1];
      return 0.0;
    }
  error(__FUNCTION__, "unknown path array");
  return 0.0;
}

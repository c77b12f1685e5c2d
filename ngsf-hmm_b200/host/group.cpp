// group.cpp - one host process driving several device contexts: the multi-GPU form of the reference's
// main()/EM()/iter_EM() call sequence (ngsF-HMM.cpp:27-171, EM.cpp:27-135, EM.cpp:139-289).
//
// Rank r of the group owns a contiguous block of individuals for the recursions and site block r for the
// allele-frequency update (geometry of include/ngsfhmm_b200.h).  Every stage runs on all ranks at once, one
// host thread per rank; the stages are separated by host joins + nfh_sync, which is all the cross-rank
// ordering the kernels' peer stores need.  The only reduction across ranks is the per-individual sum of
// log e0 (N doubles), done on the host in rank order.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ngsfhmm_host.h"

struct nfh_group {
  int n_ranks = 0;
  bool direct = true;
  uint64_t n_ind = 0, n_sites = 0;
  std::vector<nfh_ctx *> ctx;
  std::string err;
};

namespace {

// run fn(rank) on every rank concurrently; returns the first non-zero status
template <class Fn>
int on_all_ranks(nfh_group *g, Fn fn) {
  std::vector<int> rc(g->n_ranks, NFH_OK);
  if (g->n_ranks == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(g->n_ranks);
    for (int r = 0; r < g->n_ranks; r++) th.emplace_back([&, r] { rc[r] = fn(r); });
    for (auto &t : th) t.join();
  }
  for (int r = 0; r < g->n_ranks; r++)
    if (rc[r] != NFH_OK) {
      g->err = std::string("rank ") + std::to_string(r) + ": " + nfh_last_error(g->ctx[r]);
      return rc[r];
    }
  return NFH_OK;
}

// the all-to-all of one window pair without a communicator: block q of rank r's send window becomes
// block r of rank q's receive window
int exchange(nfh_group *g, int send_window, int recv_window) {
  for (int r = 0; r < g->n_ranks; r++)
    for (int q = 0; q < g->n_ranks; q++) {
      int rc = nfh_window_copy_block(g->ctx[q], recv_window, r, g->ctx[r], send_window, q);
      if (rc != NFH_OK) { g->err = nfh_last_error(g->ctx[r]); return rc; }
    }
  return on_all_ranks(g, [&](int r) { return nfh_sync(g->ctx[r]); });
}

// all-reduce (sum) of NFH_WIN_LOGE0_SUM: every rank holds the partial sums over its own sites
int reduce_loge0(nfh_group *g) {
  if (g->n_ranks == 1) return NFH_OK;
  uint64_t bytes = 0;
  int rc = nfh_exchange_window(g->ctx[0], NFH_WIN_LOGE0_SUM, nullptr, &bytes, nullptr);
  if (rc != NFH_OK) return rc;
  const size_t n = bytes / sizeof(double);
  std::vector<double> sum(n, 0.0), part(n);
  for (int r = 0; r < g->n_ranks; r++) {
    rc = nfh_window_read(g->ctx[r], NFH_WIN_LOGE0_SUM, 0, bytes, part.data());
    if (rc != NFH_OK) { g->err = nfh_last_error(g->ctx[r]); return rc; }
    for (size_t i = 0; i < n; i++) sum[i] += part[i];
  }
  for (int r = 0; r < g->n_ranks; r++) {
    rc = nfh_window_write(g->ctx[r], NFH_WIN_LOGE0_SUM, 0, bytes, sum.data());
    if (rc != NFH_OK) { g->err = nfh_last_error(g->ctx[r]); return rc; }
  }
  return NFH_OK;
}

}  // namespace

extern "C" {

int nfh_group_create(nfh_group **out, int n_ranks, const int *devices, uint64_t n_ind_total, uint64_t n_sites,
                     int fused_exchange) {
  if (!out || n_ranks < 1 || n_ranks > 8 || !devices) return NFH_ERR_ARG;
  nfh_group *g = new nfh_group;
  g->n_ranks = n_ranks; g->n_ind = n_ind_total; g->n_sites = n_sites;
  g->direct = fused_exchange != 0 && n_ranks > 1;
  g->ctx.assign(n_ranks, nullptr);
  *out = g;
  for (int r = 0; r < n_ranks; r++) {
    int rc = nfh_ctx_create(&g->ctx[r], devices[r], n_ind_total, n_sites, n_ranks, r);
    if (rc != NFH_OK) { g->err = nfh_last_error(nullptr); return rc; }
  }
  if (g->direct) {
    for (int r = 0; r < n_ranks; r++) {
      for (int q = 0; q < n_ranks; q++) {
        int rc = nfh_peer_set(g->ctx[r], NFH_WIN_POST_RECV, q, g->ctx[q]);
        if (rc == NFH_OK) rc = nfh_peer_set(g->ctx[r], NFH_WIN_EMIS_RECV, q, g->ctx[q]);
        if (rc != NFH_OK) { g->err = nfh_last_error(g->ctx[r]); return rc; }
      }
      int rc = nfh_peer_direct(g->ctx[r], 1);
      if (rc != NFH_OK) { g->err = nfh_last_error(g->ctx[r]); return rc; }
    }
  }
  return NFH_OK;
}

void nfh_group_destroy(nfh_group *g) {
  if (!g) return;
  for (nfh_ctx *c : g->ctx) if (c) nfh_sync(c);      // no kernel may still store into a window that is about to go
  for (nfh_ctx *c : g->ctx) if (c) nfh_ctx_destroy(c);
  delete g;
}

const char *nfh_group_last_error(const nfh_group *g) { return g ? g->err.c_str() : ""; }
int nfh_group_size(const nfh_group *g) { return g->n_ranks; }
nfh_ctx *nfh_group_ctx(nfh_group *g, int rank) { return rank >= 0 && rank < g->n_ranks ? g->ctx[rank] : nullptr; }

int nfh_group_upload_gl(nfh_group *g, const double *log_gl, uint64_t first_site, uint64_t n) {
  // every rank takes the part of [first_site, first_site + n) that lies in its site block
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    const uint64_t lo = std::max(first_site, nfh_site_begin(c));
    const uint64_t hi = std::min(first_site + n, nfh_site_begin(c) + nfh_sites_owned(c));
    if (hi <= lo) return (int) NFH_OK;
    return nfh_upload_gl(c, log_gl + (lo - first_site) * g->n_ind * 3, lo, hi - lo);
  });
}

int nfh_group_upload_pos_dist(nfh_group *g, const double *dist_mb) {
  return on_all_ranks(g, [&](int r) { return nfh_upload_pos_dist(g->ctx[r], dist_mb); });
}

int nfh_group_set_freq(nfh_group *g, const double *freq) {
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    return nfh_sites_owned(c) ? nfh_set_freq(c, freq + nfh_site_begin(c)) : (int) NFH_OK;
  });
}

int nfh_group_set_ind_params(nfh_group *g, const double *indF, const double *alpha) {
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    return nfh_set_ind_params(c, indF + nfh_ind_begin(c), alpha + nfh_ind_begin(c));
  });
}

int nfh_group_refresh_emissions(nfh_group *g, int with_e0) {
  int rc = on_all_ranks(g, [&](int r) {
    int s = nfh_emission_refresh(g->ctx[r], with_e0);
    return s != NFH_OK ? s : nfh_sync(g->ctx[r]);
  });
  if (rc != NFH_OK) return rc;
  if (g->n_ranks > 1) {
    if (!g->direct) rc = exchange(g, NFH_WIN_EMIS_SEND, NFH_WIN_EMIS_RECV);
    if (rc == NFH_OK && with_e0) rc = exchange(g, NFH_WIN_E0_SEND, NFH_WIN_E0_RECV);   // e0 only travels for Viterbi
    if (rc == NFH_OK) rc = reduce_loge0(g);
  }
  return rc;
}

int nfh_group_freq_init(nfh_group *g, double *freq_out) {
  int rc = on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    int s = nfh_freq_update(c, 1, 1, freq_out ? freq_out + nfh_site_begin(c) : nullptr);
    return s != NFH_OK ? s : nfh_sync(c);
  });
  if (rc != NFH_OK) return rc;
  if (g->n_ranks > 1) {
    if (!g->direct) rc = exchange(g, NFH_WIN_EMIS_SEND, NFH_WIN_EMIS_RECV);
    if (rc == NFH_OK) rc = reduce_loge0(g);
  }
  return rc;
}

int nfh_group_em_iteration(nfh_group *g, double *indF, double *alpha, int F_fixed, int alpha_fixed, int freq_est,
                           double *ind_lkl_out, double *freq_out, uint64_t stats_out[3]) {
  std::vector<uint64_t> st(3 * (size_t) g->n_ranks, 0);
  // fixed frequencies: nobody on the frequency side reads the posteriors, they stay with their individuals
  if (g->direct)
    for (nfh_ctx *c : g->ctx) nfh_peer_direct(c, freq_est ? 1 : 2);
  // stage 1: E-step + F / alpha update on the owners of the individuals (EM.cpp:151-205)
  int rc = on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    const uint64_t b = nfh_ind_begin(c), n = nfh_n_ind_owned(c);
    int s = nfh_set_ind_params(c, indF + b, alpha + b);
    if (s == NFH_OK && n)
      s = nfh_host_estep_bfgs_update(c, n, indF + b, alpha + b, F_fixed, alpha_fixed, ind_lkl_out ? ind_lkl_out + b : nullptr,
                                     &st[3 * (size_t) r]);
    return s != NFH_OK ? s : nfh_sync(c);
  });
  if (rc != NFH_OK) return rc;
  if (stats_out) {
    stats_out[0] = stats_out[1] = stats_out[2] = 0;
    for (int r = 0; r < g->n_ranks; r++) {
      stats_out[0] = std::max(stats_out[0], st[3 * (size_t) r]);
      stats_out[1] += st[3 * (size_t) r + 1];
      stats_out[2] = std::max(stats_out[2], st[3 * (size_t) r + 2]);
    }
  }
  if (!freq_est) return NFH_OK;
  // posteriors to the owners of the site blocks (already there when the kernels store into peer windows)
  if (g->n_ranks > 1 && !g->direct) {
    rc = exchange(g, NFH_WIN_POST_SEND, NFH_WIN_POST_RECV);
    if (rc != NFH_OK) return rc;
  }
  // stage 2: per-site frequency EM + emission refresh on the owners of the sites (EM.cpp:224-271)
  rc = on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    int s = nfh_freq_update(c, 1, 0, freq_out ? freq_out + nfh_site_begin(c) : nullptr);
    return s != NFH_OK ? s : nfh_sync(c);
  });
  if (rc != NFH_OK) return rc;
  if (g->n_ranks > 1) {
    if (!g->direct) rc = exchange(g, NFH_WIN_EMIS_SEND, NFH_WIN_EMIS_RECV);
    if (rc == NFH_OK) rc = reduce_loge0(g);
  }
  return rc;
}

int nfh_group_estep(nfh_group *g, double *ind_lkl_out) {
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    if (!nfh_n_ind_owned(c)) return (int) NFH_OK;
    return nfh_estep(c, ind_lkl_out + nfh_ind_begin(c));
  });
}

int nfh_group_viterbi(nfh_group *g, char *path_out) {
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    return nfh_viterbi(c, path_out ? path_out + nfh_ind_begin(c) * g->n_sites : nullptr);
  });
}

int nfh_group_get_posterior(nfh_group *g, double *marg1_out) {
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    return nfh_get_posterior(c, marg1_out + nfh_ind_begin(c) * g->n_sites);
  });
}

int nfh_group_get_freq(nfh_group *g, double *freq_out) {
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    return nfh_sites_owned(c) ? nfh_get_freq(c, freq_out + nfh_site_begin(c)) : (int) NFH_OK;
  });
}

int nfh_group_geno_posterior(nfh_group *g, const char *path_all, double *geno_out) {
  // path_all [n_ind][n_sites]; every rank needs the columns of its site block as [n_ind][sites_owned]
  return on_all_ranks(g, [&](int r) {
    nfh_ctx *c = g->ctx[r];
    const uint64_t s0 = nfh_site_begin(c), w = nfh_sites_owned(c);
    if (!w) return (int) NFH_OK;
    std::vector<char> cols((size_t) g->n_ind * w);
    for (uint64_t i = 0; i < g->n_ind; i++) memcpy(&cols[(size_t) i * w], path_all + i * g->n_sites + s0, w);
    return nfh_geno_posterior(c, cols.data(), geno_out + s0 * g->n_ind * 3);
  });
}

}  // extern "C"

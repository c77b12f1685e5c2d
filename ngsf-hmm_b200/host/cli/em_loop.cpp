// em_loop.cpp - iteration control of the drop-in binary (EM(), EM.cpp:27-135):
// stop rule, per-iteration report, --log dumps, signal handling, final Viterbi.
// Every iteration body is one call into the host library
// (nfh_group_em_iteration = iter_EM on the device(s) of the group).
#include <signal.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>

#include "ngsfhmm_host.h"
#include "run_state.hpp"

namespace nfh_cli {

namespace {

volatile sig_atomic_t g_keep_going = 1;
volatile sig_atomic_t g_patience = 3;

void on_signal(int s) {
  if (g_keep_going)
    fprintf(stderr, "\n\"%s\" signal caught! Will try to exit nicely (results of the current iteration are written).\n",
            strsignal(s));
  g_patience = g_patience - 1;
  if (g_patience > 0)
    fprintf(stderr, "\t-> If you really want to force an unclean exit Ctr+C %d more times\n", (int) g_patience);
  fflush(stderr);
  if (!g_patience) _exit(0);
  g_keep_going = 0;
}

void install_handlers() {   // catch_SIG, gen_func.cpp:40-52
  struct sigaction sa;
  sigemptyset(&sa.sa_mask);
  sa.sa_flags = 0;
  sa.sa_handler = on_signal;
  sigaction(SIGTERM, &sa, nullptr);
  sigaction(SIGQUIT, &sa, nullptr);
  sigaction(SIGPIPE, &sa, nullptr);
  sigaction(SIGINT, &sa, nullptr);
}

// array_max_pos, gen_func.cpp:73-84: first element strictly larger than everything before it
size_t first_max(const std::vector<double> &v) {
  size_t best = 0;
  double top = -INFINITY;
  for (size_t i = 0; i < v.size(); i++)
    if (v[i] > top) { best = i; top = v[i]; }
  return best;
}

}  // namespace

void run_em(RunState &st, bool write_final) {
  Options &o = st.opt;
  const uint64_t N = o.n_ind, S = o.n_sites;
  g_keep_going = 1;
  install_handlers();

  // the frequencies come back every iteration (8 bytes per site): page-lock their host array
  nfh_ctx *ctx0 = nfh_group_ctx(st.grp, 0);
  const bool pinned = nfh_host_register(ctx0, st.freq.data(), S * sizeof(double)) == NFH_OK;

  uint64_t iter = 0;
  double max_eps = -INFINITY;
  std::vector<double> prev_lkl(N, -INFINITY), eps(N, -INFINITY);

  if (o.verbose >= 5) {
    printf("==> Initial parameters:\n");
    for (uint64_t i = 0; i < N; i++) printf("\t%.10f\t%f\n", st.indF[i], st.alpha[i]);
    for (uint64_t s = 0; s < S; s++) printf("\t%f", st.freq[s]);
    printf("\n");
  }

  while ((st.prev_tot_lkl - st.tot_lkl > o.min_epsilon || max_eps > o.min_epsilon || iter < o.min_iters) &&
         iter < o.max_iters && g_keep_going) {
    if (o.log && (iter == 1 || iter % o.log == 0)) {
      if (o.verbose >= 1) printf("==> Printing current iteration parameters\n");
      check(st, nfh_group_get_posterior(st.grp, st.marg1.data()), "nfh_get_posterior");
      write_outputs(st);
    }
    const time_t t0 = time(nullptr);
    iter++;
    if (o.verbose >= 1) printf("\nIteration %lu:\n", (unsigned long) iter);
    if (o.verbose >= 1)
      printf("==> Forward Recursion\n==> Backward Recursion\n==> Marginal probabilities\n%s%s",
             (o.indF_fixed && o.alpha_fixed) ? "==> Inbreeding and transition parameter not estimated!\n"
                                             : "==> Update inbreeding and transition parameter\n",
             o.freq_est == 0 ? "==> Alelle frequencies not estimated!\n"
                             : "==> Estimating allele frequencies and calculating emission probabilities\n");

    uint64_t stats[3] = {0, 0, 0};
    check(st, nfh_group_em_iteration(st.grp, st.indF.data(), st.alpha.data(), o.indF_fixed, o.alpha_fixed, o.freq_est,
                                     st.ind_lkl.data(), st.freq.data(), stats),
          "iter_EM");
    if (o.verbose >= 4 && !(o.indF_fixed && o.alpha_fixed))
      for (uint64_t i = 0; i < N; i++) printf("\t%.10f\t%f\n", st.indF[i], st.alpha[i]);

    st.prev_tot_lkl = st.tot_lkl;
    st.tot_lkl = 0;
    for (uint64_t i = 0; i < N; i++) {
      st.tot_lkl += st.ind_lkl[i];
      eps[i] = (st.ind_lkl[i] - prev_lkl[i]) / fabs(prev_lkl[i]);
    }
    const size_t who = first_max(eps);
    max_eps = eps[who];
    prev_lkl = st.ind_lkl;

    const time_t t1 = time(nullptr);
    if (o.verbose >= 1)
      printf("\tLogLkl: %.15f\t max lkl epsilon: %.15f\ttime: %.0f (s)\n", st.tot_lkl, max_eps, difftime(t1, t0));
    if (o.verbose >= 3) {
      for (uint64_t i = 0; i < N; i++)
        printf("\tInd %lu: %.15f\t lkl epsilon: %.15f%s\n", (unsigned long) (i + 1), st.ind_lkl[i], eps[i],
               i == who ? " (max)" : "");
      printf("\tBFGS: %lu batched rounds, %lu objective evaluations\n", (unsigned long) stats[0],
             (unsigned long) stats[1]);
    }
    fflush(stdout);
  }
  if (iter >= o.max_iters) printf("WARN: Maximum number of iterations reached! Check if analysis converged... \n");

  if (pinned) nfh_host_unregister(ctx0, st.freq.data());

  if (o.verbose >= 1) printf("\n==> Decoding most probable path (Viterbi)\n");
  check(st, nfh_group_set_ind_params(st.grp, st.indF.data(), st.alpha.data()), "nfh_set_ind_params");
  check(st, nfh_group_refresh_emissions(st.grp, 1), "nfh_emission_refresh");
  check(st, nfh_group_viterbi(st.grp, st.path.data()), "nfh_viterbi");

  if (o.verbose >= 1) printf("Final logLkl: %f\n", st.tot_lkl);
  check(st, nfh_group_get_posterior(st.grp, st.marg1.data()), "nfh_get_posterior");
  if (write_final) {
    if (o.verbose >= 1) printf("Printing final results\n");
    write_outputs(st);
  }
}

}  // namespace nfh_cli

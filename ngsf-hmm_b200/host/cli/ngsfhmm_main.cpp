// ngsfhmm_main.cpp - drop-in replacement for the reference's command-line
// program (ngsF-HMM.cpp:27-171): same flags, same input formats, same output
// files; the EM iteration body and the Viterbi decoding run on the GPU.
#include <cstdio>
#include <vector>

#include "run_state.hpp"

using namespace nfh_cli;

int main(int argc, char **argv) {
  RunState st;
  parse_options(st.opt, argc, argv);
  if (st.opt.n_threads > st.opt.n_ind) {
    warn("main", "adjusting threads (--n_threads) to match number of individuals!");
    st.opt.n_threads = (unsigned) st.opt.n_ind;
  }
  inspect_geno_file(st);
  if (st.opt.verbose >= 1) printf("==> Reading data\n> Sites coordinates\n");
  read_positions(st);
  read_genotypes(st);
  if (st.opt.verbose >= 6) printf("> Init output\n");
  create_device_state(st);
  st.log_gl.reset();                       // GL now lives on the device
  if (st.opt.n_rep == 1) {
    init_start_values(st, st.opt.seed);
    run_em(st, true);
  } else {
    // Replicates (ngsF-HMM.sh:83-116 runs the whole program once per replicate and keeps the files of the
    // highest final logLkl): here the parsed input and the device copy of GL are shared, only the start
    // values change (seed, seed + 1, ...); the final state of the best replicate is kept on the host.
    struct Best { bool have = false; unsigned rep = 0; std::vector<double> freq, indF, alpha, ind_lkl, marg1;
                  std::vector<char> path; double tot = 0, prev = 0; } best;
    for (unsigned rep = 0; rep < st.opt.n_rep; rep++) {
      if (st.opt.verbose >= 1) printf("\n==> Replicate %u of %u (seed %u)\n", rep + 1, st.opt.n_rep, st.opt.seed + rep);
      init_start_values(st, st.opt.seed + rep);
      run_em(st, false);
      if (!best.have || st.tot_lkl > best.tot) {
        best.have = true; best.rep = rep; best.tot = st.tot_lkl; best.prev = st.prev_tot_lkl;
        best.freq = st.freq; best.indF = st.indF; best.alpha = st.alpha; best.ind_lkl = st.ind_lkl;
        best.marg1.swap(st.marg1); best.path.swap(st.path);
      }
    }
    if (st.opt.verbose >= 1)
      printf("\n==> Best replicate: %u (seed %u), logLkl %f\nPrinting final results\n", best.rep + 1,
             st.opt.seed + best.rep, best.tot);
    st.freq.swap(best.freq); st.indF.swap(best.indF); st.alpha.swap(best.alpha); st.ind_lkl.swap(best.ind_lkl);
    st.marg1.swap(best.marg1); st.path.swap(best.path);
    st.tot_lkl = best.tot; st.prev_tot_lkl = best.prev;
    write_outputs(st);
  }
  if (st.opt.verbose >= 1) printf("Freeing memory...\n");
  nfh_group_destroy(st.grp);
  if (st.opt.verbose >= 1) printf("Done!\n");
  return 0;
}

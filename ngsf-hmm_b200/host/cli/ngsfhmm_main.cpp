// ngsfhmm_main.cpp - drop-in replacement for the reference's command-line
// program (ngsF-HMM.cpp:27-171): same flags, same input formats, same output
// files; the EM iteration body and the Viterbi decoding run on the GPU.
#include <cstdio>

#include "run_state.hpp"

using namespace nfh_cli;

int main(int argc, char **argv) {
  RunState st;
  parse_options(st.opt, argc, argv);
  if (st.opt.n_threads > st.opt.n_ind) {
    warn("main", "adjusting threads (--n_threads) to match number of individuals!");
    st.opt.n_threads = (unsigned) st.opt.n_ind;
  }
  if (st.opt.verbose >= 1) printf("==> Reading data\n> Sites coordinates\n");
  read_positions(st);
  read_genotypes(st);
  if (st.opt.verbose >= 6) printf("> Init output\n");
  init_start_values(st);
  std::vector<double>().swap(st.log_gl);   // GL now lives on the device
  run_em(st);
  if (st.opt.verbose >= 1) printf("Freeing memory...\n");
  nfh_ctx_destroy(st.ctx);
  if (st.opt.verbose >= 1) printf("Done!\n");
  return 0;
}

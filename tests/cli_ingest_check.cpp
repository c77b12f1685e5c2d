// cli_ingest_check.cpp - CPU test helper: runs the drop-in binary's option parser and input readers
// (ngsf-hmm_b200/host/cli/options.cpp, ingest.cpp - the product sources, compiled as they are) and dumps what
// they would hand to the device: <out>.dist (n_sites doubles, Mb) and <out>.gl (site-major n_sites x n_ind x 3
// normalised natural-log GL).  tests/test_cli_ingest.py compares both, bit for bit, with the reference's own
// readers (read_data.cpp through oracle/ref_harness.cpp::ref_read_inputs).  No device call is made.
#include <cstdio>

#include "run_state.hpp"

using namespace nfh_cli;

int main(int argc, char **argv) {
  RunState st;
  parse_options(st.opt, argc, argv);
  inspect_geno_file(st);
  read_positions(st);
  read_genotypes(st);
  const std::string base = st.opt.out;
  FILE *f = fopen((base + ".dist").c_str(), "wb");
  if (!f) return 2;
  fwrite(st.dist_mb.data(), sizeof(double), st.dist_mb.size(), f);
  fclose(f);
  f = fopen((base + ".gl").c_str(), "wb");
  if (!f) return 2;
  fwrite(st.log_gl.get(), sizeof(double), st.opt.n_sites * st.opt.n_ind * 3, f);
  fclose(f);
  return 0;
}

// cli_report_check.cpp - CPU test helper: the drop-in binary's output writer (host/cli/report.cpp, the product
// source compiled as it is) on a state read from raw files; the device call behind <out>.geno is stubbed (zeros).
//   cli_report_check <in prefix> <out prefix> <n_ind> <n_sites> <host threads>
// reads <in>.tot (1 double) .indF .alpha .ind_lkl (n_ind) .freq (n_sites) .marg (n_ind x n_sites doubles)
// .path (n_ind x n_sites bytes); tests/test_cli_report.py compares <out>.indF / <out>.ibd byte for byte with the
// reference's print_iter (EM.cpp:293-380 through oracle/ref_harness.cpp::ref_print_iter).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "run_state.hpp"

extern "C" {
const char *nfh_last_error(const nfh_ctx *) { return ""; }
const char *nfh_strerror(int) { return "stub"; }
const char *nfh_group_last_error(const nfh_group *) { return ""; }
int nfh_group_set_freq(nfh_group *, const double *) { return 0; }
int nfh_group_geno_posterior(nfh_group *, const char *, double *out) { (void) out; return 0; }
}

namespace nfh_cli {
void check(RunState &, int rc, const char *) { if (rc) exit(3); }
}

using namespace nfh_cli;

template <class T>
static void slurp(const std::string &path, std::vector<T> &v, size_t n) {
  v.resize(n);
  FILE *f = fopen(path.c_str(), "rb");
  if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc != 6) return 2;
  const std::string in = argv[1];
  RunState st;
  st.opt.out = argv[2];
  const uint64_t N = st.opt.n_ind = (uint64_t) atoll(argv[3]);
  const uint64_t S = st.opt.n_sites = (uint64_t) atoll(argv[4]);
  st.opt.host_threads = (unsigned) atoi(argv[5]);
  std::vector<double> tot;
  slurp(in + ".tot", tot, 1);
  st.tot_lkl = tot[0];
  slurp(in + ".indF", st.indF, N);
  slurp(in + ".alpha", st.alpha, N);
  slurp(in + ".ind_lkl", st.ind_lkl, N);
  slurp(in + ".freq", st.freq, S);
  slurp(in + ".marg", st.marg1, N * S);
  slurp(in + ".path", st.path, N * S);
  write_outputs(st);
  return 0;
}

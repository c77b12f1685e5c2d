// cli_startvalues_check.cpp - CPU test helper: the drop-in binary's start values (host/cli/startvalues.cpp, the
// product source compiled as it is) with the device calls stubbed out.  Usage:
//   cli_startvalues_check <--indF argument> <--freq argument> <seed> <n_ind> <n_sites> <freq_est> <out prefix>
// writes <out>.indF / .alpha / .freq as raw doubles; tests/test_cli_startvalues.py compares them with the
// reference's init_output (parse_args.cpp:229-419, through oracle/ref_harness.cpp::ref_init_start_values).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "run_state.hpp"

// ---- stubs for the device side: start values that need the GPU (--freq e) are covered by the GPU tests
extern "C" {
int nfh_device_count(void) { return 0; }
const char *nfh_last_error(const nfh_ctx *) { return ""; }
const char *nfh_strerror(int) { return "stub"; }
const char *nfh_group_last_error(const nfh_group *) { return ""; }
int nfh_group_create(nfh_group **, int, const int *, uint64_t, uint64_t, int) { return 0; }
int nfh_group_upload_gl(nfh_group *, const double *, uint64_t, uint64_t) { return 0; }
int nfh_group_upload_pos_dist(nfh_group *, const double *) { return 0; }
int nfh_group_set_freq(nfh_group *, const double *) { return 0; }
int nfh_group_refresh_emissions(nfh_group *, int) { return 0; }
int nfh_group_freq_init(nfh_group *, double *) { return 0; }
}

using namespace nfh_cli;

static void dump(const std::string &path, const std::vector<double> &v) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) exit(2);
  fwrite(v.data(), sizeof(double), v.size(), f);
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc != 8) return 2;
  RunState st;
  st.opt.indF_arg = argv[1];
  st.opt.freq_arg = argv[2];
  const unsigned seed = (unsigned) atoi(argv[3]);
  st.opt.n_ind = (uint64_t) atoll(argv[4]);
  st.opt.n_sites = (uint64_t) atoll(argv[5]);
  st.opt.freq_est = atoi(argv[6]);
  st.opt.verbose = 0;
  init_start_values(st, seed);
  const std::string out = argv[7];
  dump(out + ".indF", st.indF);
  dump(out + ".alpha", st.alpha);
  dump(out + ".freq", st.freq);
  return 0;
}

// run_state.hpp - host-side state of one ngsF-HMM run (drop-in CLI).
//
// Mirrors what the reference keeps in its `params` struct (ngsF-HMM.hpp:13-52)
// minus the big per-individual-site arrays, which live on the device behind
// the C ABI (include/ngsfhmm_b200.h).
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "ngsfhmm_host.h"

namespace nfh_cli {

struct Options {   // one field per command-line flag of the reference (parse_args.cpp:43-68)
  std::string geno, pos, freq_arg, indF_arg, out;
  bool lkl = false, loglkl = false, call_geno = false, indF_fixed = false, alpha_fixed = false, log_bin = false;
  bool in_bin = false;
  uint64_t n_ind = 0, n_sites = 0;
  int freq_est = 1, e_prob = 1;
  unsigned log = 0, min_iters = 10, max_iters = 100, n_threads = 1, verbose = 1, seed = 0;
  bool n_threads_given = false;
  unsigned host_threads = 1;   // threads for parsing / formatting: --n_threads when given, else the host's cores (<= 16)
  double min_epsilon = 1e-5;
  bool have_geno = false, have_pos = false, have_out = false;
  int device = 0;          // --device (extension; not a reference flag): first CUDA ordinal used
  int n_gpus = 1;          // --n_gpus (extension): individuals sharded over devices device .. device + n_gpus - 1 of this
                           // box, one context per GPU driven by this one process (nfh_group_*, ngsfhmm_host.h)
  unsigned n_rep = 1;      // --n_rep (extension): replicates of the whole EM from different --seed values on one
                           // ingest and one GL upload; the best final logLkl is written (what ngsF-HMM.sh does
                           // with one process, one parse and one upload per replicate)
};

struct RunState {
  Options opt;
  nfh_group *grp = nullptr;           // one context per GPU (a group of one is the single-GPU run)
  std::vector<double> dist_mb;        // n_sites
  std::unique_ptr<double[]> log_gl;   // site-major n_sites x n_ind x 3, normalised natural-log GL (2.4 GB at
                                      // configs[1]: allocated without a fill pass)
  std::vector<double> freq, indF, alpha, ind_lkl;
  std::vector<char> path;             // n_ind x n_sites
  std::vector<double> marg1;          // n_ind x n_sites (fetched only for output)
  double prev_tot_lkl = 0.0, tot_lkl = 0.0;
};

// Reference-style fatal error (gen_func.cpp:12-18): message to stderr, perror, exit(-1).
[[noreturn]] void fatal(const char *where, const char *msg);
void warn(const char *where, const char *msg);
void check(RunState &st, int rc, const char *where);   // maps C-ABI status to fatal()

// options.cpp
void parse_options(Options &o, int argc, char **argv);
// ingest.cpp
void inspect_geno_file(RunState &st);   // binary / gz text by name, size check; before anything is read
void read_positions(RunState &st);
void read_genotypes(RunState &st);
// startvalues.cpp
void create_device_state(RunState &st);                   // context + GL / distance upload (once per process)
void init_start_values(RunState &st, unsigned seed);      // start values + initial emissions (once per replicate)
// em_loop.cpp
void run_em(RunState &st, bool write_final);              // EM loop, Viterbi, posterior; outputs if write_final
// report.cpp
void write_outputs(RunState &st);

extern const char *kVersion;

}  // namespace nfh_cli

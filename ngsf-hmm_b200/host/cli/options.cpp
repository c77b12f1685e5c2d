// options.cpp - command line of the drop-in binary.  Same flags, defaults,
// echo and checks as the reference (parse_args.cpp:5-225); --n_threads is
// accepted and ignored (the device path has no thread pool), --device is new.
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <thread>

#include "run_state.hpp"

namespace nfh_cli {

const char *kVersion = "1.1.0-b200";

[[noreturn]] void fatal(const char *where, const char *msg) {
  fflush(stdout);
  fprintf(stderr, "\n=====\nERROR: [%s] %s\n=====\n\n", where, msg);
  perror("\t");
  fflush(stderr);
  exit(-1);
}

void warn(const char *where, const char *msg) {
  fflush(stdout);
  fprintf(stderr, "\n=======\nWARNING: [%s] %s\n=======\n\n", where, msg);
  fflush(stderr);
}

void parse_options(Options &o, int argc, char **argv) {
  o.seed = rand() % 1000;
  static struct option table[] = {
      {"geno", required_argument, nullptr, 'g'},      {"pos", required_argument, nullptr, 'Z'},
      {"lkl", no_argument, nullptr, 'l'},             {"loglkl", no_argument, nullptr, 'L'},
      {"n_ind", required_argument, nullptr, 'n'},     {"n_sites", required_argument, nullptr, 's'},
      {"call_geno", no_argument, nullptr, 'G'},       {"freq", required_argument, nullptr, 'f'},
      {"freq_est", required_argument, nullptr, 'F'},  {"e_prob", required_argument, nullptr, 'e'},
      {"indF", required_argument, nullptr, 'i'},      {"indF_fixed", no_argument, nullptr, 'I'},
      {"alpha_fixed", no_argument, nullptr, 'A'},     {"out", required_argument, nullptr, 'o'},
      {"log", required_argument, nullptr, 'X'},       {"log_bin", required_argument, nullptr, 'b'},
      {"min_iters", required_argument, nullptr, 'm'}, {"max_iters", required_argument, nullptr, 'M'},
      {"min_epsilon", required_argument, nullptr, 'E'}, {"n_threads", required_argument, nullptr, 'x'},
      {"verbose", required_argument, nullptr, 'V'},   {"seed", required_argument, nullptr, 'S'},
      {"device", required_argument, nullptr, 'D'},    {"n_rep", required_argument, nullptr, 'R'},
      {"n_gpus", required_argument, nullptr, 'U'},
      {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long_only(argc, argv, "g:Z:lLn:s:Gf:F:e:i:IAo:X:b:m:M:E:x:V:S:D:R:U:", table, nullptr)) != -1) {
    switch (c) {
      case 'g': o.geno = optarg; o.have_geno = true; break;
      case 'Z': o.pos = optarg; o.have_pos = true; break;
      case 'l': o.lkl = true; break;
      case 'L': o.lkl = true; o.loglkl = true; break;
      case 'n': o.n_ind = (uint64_t) atoi(optarg); break;
      case 's': o.n_sites = (uint64_t) atoi(optarg); break;
      case 'G': o.call_geno = true; break;
      case 'f': o.freq_arg = optarg; break;
      case 'F': o.freq_est = atoi(optarg); break;
      case 'e': o.e_prob = atoi(optarg); break;
      case 'i': o.indF_arg = optarg; break;
      case 'I': o.indF_fixed = true; break;
      case 'A': o.alpha_fixed = true; break;
      case 'o': o.out = optarg; o.have_out = true; break;
      case 'X': o.log = (unsigned) atoi(optarg); break;
      case 'b': o.log = (unsigned) atoi(optarg); o.log_bin = true; break;
      case 'm': o.min_iters = (unsigned) atoi(optarg); break;
      case 'M': o.max_iters = (unsigned) atoi(optarg); break;
      case 'E': o.min_epsilon = atof(optarg); break;
      case 'x': o.n_threads = (unsigned) atoi(optarg); o.n_threads_given = true; break;
      case 'V': o.verbose = (unsigned) atoi(optarg); break;
      case 'S': o.seed = (unsigned) atoi(optarg); break;
      case 'D': o.device = atoi(optarg); break;
      case 'R': o.n_rep = (unsigned) atoi(optarg); break;
      case 'U': o.n_gpus = atoi(optarg); break;
      default: exit(-1);
    }
  }
  if (o.freq_arg.empty()) o.freq_arg = "r";                 // parse_args.cpp:150-153
  if (o.indF_arg.empty()) o.indF_arg = "0.01-0.001";        // parse_args.cpp:154-157

  if (o.verbose >= 1) {
    auto tf = [](bool b) { return b ? "true" : "false"; };
    printf("==> Input Arguments:\n");
    printf("\tgeno: %s\n\tpos: %s\n\tlkl: %s\n\tloglkl: %s\n\tn_ind: %lu\n\tn_sites: %lu\n\tcall_geno: %s\n\tfreq: %s\n"
           "\tfreq_est: %d\n\te_prob: %d\n\tindF: %s\n\tindF_fixed: %s\n\talpha_fixed: %s\n\tout: %s\n\tlog: %u\n"
           "\tlog_bin: %s\n\tmin_iters: %d\n\tmax_iters: %d\n\tmin_epsilon: %.10f\n\tn_threads: %d\n\tverbose: %d\n"
           "\tseed: %d\n\tversion: %s (%s @ %s)\n\n",
           o.have_geno ? o.geno.c_str() : "(null)", o.have_pos ? o.pos.c_str() : "(null)", tf(o.lkl), tf(o.loglkl),
           (unsigned long) o.n_ind, (unsigned long) o.n_sites, tf(o.call_geno), o.freq_arg.c_str(), o.freq_est, o.e_prob,
           o.indF_arg.c_str(), tf(o.indF_fixed), tf(o.alpha_fixed), o.have_out ? o.out.c_str() : "(null)", o.log,
           tf(o.log_bin), (int) o.min_iters, (int) o.max_iters, o.min_epsilon, (int) o.n_threads, (int) o.verbose,
           (int) o.seed, kVersion, __DATE__, __TIME__);
  }
  if (o.verbose >= 4)
    printf("==> Verbose values greater than 4 for debugging purpose only. Expect large amounts of info on screen\n");

  const char *fn = "parse_cmd_args";
  if (!o.have_geno) fatal(fn, "genotype input file (--geno) missing!");
  if (!o.have_pos) fatal(fn, "positions input file (--pos) missing!");
  if (o.n_ind == 0) fatal(fn, "number of individuals (--n_ind) missing!");
  if (o.n_sites == 0) fatal(fn, "number of sites (--n_sites) missing!");
  if (o.call_geno && !o.lkl) fatal(fn, "can only call genotypes from likelihoods!");
  if (o.freq_est < 0 || o.freq_est > 2) fatal(fn, "invalid MAF estimation method!");
  if (o.e_prob < 0 || o.e_prob > 2) fatal(fn, "invalid emission probability calculation method!");
  if (o.e_prob > 1) warn(fn, "calculation of emission probabilities accounting for LD is still under development!");
  if (!o.have_out) fatal(fn, "output prefix (--out) missing!");
  if (o.min_iters < 1 || o.max_iters < 1 || o.min_iters >= o.max_iters) fatal(fn, "invalid number of iterations!");
  if (o.n_threads < 1) fatal(fn, "invalid number of threads!");
  if (o.n_rep < 1) fatal(fn, "invalid number of replicates!");
  if (o.n_gpus < 1 || o.n_gpus > 8) fatal(fn, "invalid number of GPUs (--n_gpus 1..8)!");
  {
    const unsigned hw = std::thread::hardware_concurrency();
    o.host_threads = o.n_threads_given ? o.n_threads : std::max(1u, std::min(hw ? hw : 1u, 16u));
  }
  // The haplotype-frequency paths abort in the reference itself (freq[0] = -1 reaches haplo_freq,
  // gen_func.cpp:1030-1031); keep the same message instead of inventing behaviour.
  if (o.freq_est == 2 || o.e_prob == 2) fatal("haplo_freq", "invalid allele frequencies");
}

}  // namespace nfh_cli

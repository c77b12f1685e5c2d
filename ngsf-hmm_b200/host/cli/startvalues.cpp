// startvalues.cpp - initial F / alpha / allele frequencies and device upload
// (the parts of init_output that touch the hot-path state, parse_args.cpp:229-419).
//   --indF  "r" | file (2 columns) | "F,alpha" / "F-alpha"     clamp [1e-6, 1-1e-6]
//   --freq  "r" | "e" | file (1 column) | number                clamp [0.01, 0.49]
// "r" draws from a combined Tausworthe generator seeded with --seed (the
// reference uses GSL's gsl_rng_taus; GSL is not available to verify the
// stream bit for bit).  "e" runs the frequency EM with F = 0 on the device.
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "run_state.hpp"

namespace nfh_cli {

namespace {

struct Taus {   // three-component combined Tausworthe generator (L'Ecuyer 1996)
  uint32_t a, b, c;
  static uint32_t step(uint32_t s, int p, int q, uint32_t mask, int sh) { return ((s & mask) << sh) ^ (((s << p) ^ s) >> q); }
  uint32_t next() {
    a = step(a, 13, 19, 4294967294u, 12);
    b = step(b, 2, 25, 4294967288u, 4);
    c = step(c, 3, 11, 4294967280u, 17);
    return a ^ b ^ c;
  }
  explicit Taus(uint32_t seed) {
    if (seed == 0) seed = 1;
    a = 69069u * seed; b = 69069u * a; c = 69069u * b;
    for (int w = 0; w < 6; w++) next();
  }
  double uniform() { return next() / 4294967296.0; }
};

inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// numbers of a line split on any of `seps`; tokens that are not numbers are dropped
int split_numbers(const char *text, const char *seps, double *out, int cap) {
  int n = 0;
  const char *p = text;
  while (*p) {
    size_t len = strcspn(p, seps);
    if (len) {
      char tok[256];
      size_t m = len < sizeof tok - 1 ? len : sizeof tok - 1;
      memcpy(tok, p, m);
      tok[m] = '\0';
      char *end = nullptr;
      double v = strtod(tok, &end);
      if (end != tok && *end == '\0' && n < cap) out[n++] = v;
    }
    p += len;
    if (*p) p++;
  }
  return n;
}

void chomp(char *s) {   // ONE trailing '\n' or '\r', as gen_func.cpp:192-199 (a CRLF file keeps its '\r' and fails the field count)
  const size_t n = strlen(s);
  if (n && (s[n - 1] == '\n' || s[n - 1] == '\r')) s[n - 1] = '\0';
}

}  // namespace

void check(RunState &st, int rc, const char *where) {
  if (rc == NFH_OK) return;
  const char *msg = st.grp ? nfh_group_last_error(st.grp) : nfh_last_error(nullptr);
  if (!msg || !*msg) msg = nfh_last_error(nullptr);
  if (!msg || !*msg) msg = nfh_strerror(rc);
  switch (rc) {
    case NFH_ERR_NAN: fatal("forward", "invalid Lkl found!");          // HMM.cpp:18-21
    case NFH_ERR_FWBW: fatal("iter_EM", "Fw and Bw lkl do not match!"); // EM.cpp:166-170
    default: fatal(where, msg);
  }
}

void create_device_state(RunState &st) {
  Options &o = st.opt;
  // rank r of the group runs on device o.device + r; the kernels exchange posteriors and emission ratios
  // through peer memory (NVLink), see ngsfhmm_host.h
  std::vector<int> devices(o.n_gpus);
  const int visible = nfh_device_count();
  // NFH_SHARE_DEVICES=1 (testing aid): ranks wrap around the visible devices, so the sharded run can be
  // exercised on a box with fewer GPUs than ranks
  const char *share = getenv("NFH_SHARE_DEVICES");
  for (int r = 0; r < o.n_gpus; r++)
    devices[r] = (share && *share == '1' && visible > 0) ? (o.device + r) % visible : o.device + r;
  check(st, nfh_group_create(&st.grp, o.n_gpus, devices.data(), o.n_ind, o.n_sites, 1), "nfh_group_create");
  check(st, nfh_group_upload_pos_dist(st.grp, st.dist_mb.data()), "nfh_upload_pos_dist");
  check(st, nfh_group_upload_gl(st.grp, st.log_gl.get(), 0, o.n_sites), "nfh_upload_gl");
}

void init_start_values(RunState &st, unsigned seed) {
  Options &o = st.opt;
  const char *fn = "init_output";
  const uint64_t N = o.n_ind, S = o.n_sites;
  Taus rng(seed);
  const double lo = 0.000001, hi = 1 - lo;
  st.indF.assign(N, 0.0);
  st.alpha.assign(N, 0.0);
  std::vector<char> buf(500000);

  gzFile fh;
  if (o.indF_arg == "r") {
    if (o.verbose >= 1) printf("==> Using random initial inbreeding values.\n");
    for (uint64_t i = 0; i < N; i++) {
      st.indF[i] = lo + rng.uniform() * (hi - lo);
      st.alpha[i] = lo + rng.uniform() * (hi - lo);
    }
  } else if ((fh = gzopen(o.indF_arg.c_str(), "r")) != nullptr) {
    if (o.verbose >= 1) printf("==> Reading initial inbreeding values from file \"%s\".\n", o.indF_arg.c_str());
    uint64_t i = 0;
    while (gzgets(fh, buf.data(), (int) buf.size()) != nullptr) {
      chomp(buf.data());
      if (buf[0] == '\0') continue;
      double t[4];
      if (i >= N || split_numbers(buf.data(), " ,-\t", t, 4) != 2) fatal(fn, "wrong INDF file format!");
      st.indF[i] = clampd(t[0], lo, hi);
      st.alpha[i] = clampd(t[1], lo, hi);
      i++;
    }
    gzclose(fh);
  } else {
    if (o.verbose >= 1) printf("==> Setting initial inbreeding values to: %s\n", o.indF_arg.c_str());
    double t[4];
    if (split_numbers(o.indF_arg.c_str(), ",-", t, 4) != 2) fatal(fn, "wrong INDF parameters format!");
    for (uint64_t i = 0; i < N; i++) {
      st.indF[i] = clampd(t[0], lo, hi);
      st.alpha[i] = clampd(t[1], lo, hi);
    }
  }

  const double flo = 0.01, fhi = 0.5 - flo;
  st.freq.assign(S, flo);
  bool estimate = false;
  if (o.freq_arg == "r") {
    if (o.verbose >= 1) printf("==> Using random initial frequency values.\n");
    for (uint64_t s = 0; s < S; s++) st.freq[s] = flo + rng.uniform() * (fhi - flo);
  } else if (o.freq_arg == "e") {
    if (o.verbose >= 1) printf("==> Estimating initial frequency values assuming HWE.\n");
    estimate = true;
  } else if ((fh = gzopen(o.freq_arg.c_str(), "r")) != nullptr) {
    if (o.verbose >= 1) printf("==> Reading initial frequency values from file \"%s\".\n", o.freq_arg.c_str());
    uint64_t s = 0;
    while (gzgets(fh, buf.data(), (int) buf.size()) != nullptr) {
      chomp(buf.data());
      if (buf[0] == '\0') continue;
      double t[4];
      int n = split_numbers(buf.data(), " ,-\t", t, 4);
      if (n == 0) { printf("> Header found! Skipping line...\n"); continue; }
      if (s >= S || n != 1) fatal(fn, "wrong FREQ file format!");
      st.freq[s++] = clampd(t[0], flo, fhi);
    }
    gzclose(fh);
  } else {
    if (o.verbose >= 1) printf("==> Setting initial frequency values to: %s\n", o.freq_arg.c_str());
    const double v = clampd(atof(o.freq_arg.c_str()), flo, fhi);
    for (uint64_t s = 0; s < S; s++) st.freq[s] = v;
  }

  if (o.verbose >= 1) printf("==> Calculating initial emission probabilities\n");
  if (estimate) {
    check(st, nfh_group_freq_init(st.grp, st.freq.data()), "nfh_freq_update");   // est_maf with F = 0
    if (o.freq_est != 1 && S > 1) {
      // parse_args.cpp:316-318 estimates a site only `if(freq_est == 1 || s == 1)`: with --freq_est 0 the first
      // site gets its estimate and every other site stays at the lower limit of the range (0.01) for the run
      std::fill(st.freq.begin() + 1, st.freq.end(), flo);
      check(st, nfh_group_set_freq(st.grp, st.freq.data()), "nfh_set_freq");
      check(st, nfh_group_refresh_emissions(st.grp, 0), "nfh_emission_refresh");
    }
  } else {
    check(st, nfh_group_set_freq(st.grp, st.freq.data()), "nfh_set_freq");
    check(st, nfh_group_refresh_emissions(st.grp, 0), "nfh_emission_refresh");
  }
  st.ind_lkl.assign(N, -INFINITY);
  st.path.assign(N * S, 0);
  st.marg1.assign(N * S, 0.0);
  st.prev_tot_lkl = 0.0;
  st.tot_lkl = 0.0;
}

}  // namespace nfh_cli

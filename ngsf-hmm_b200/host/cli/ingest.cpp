// ingest.cpp - input files of the drop-in binary, in the reference's formats
// (shared/read_data.cpp:13-218, ngsF-HMM.cpp:47-117):
//   --pos   text (optionally gz): chrom <TAB> position; distance to the previous
//           site in bp -> Mb, +inf where the chromosome changes
//   --geno  *.gz => text: called genotypes (1 column per individual) or, with
//           --lkl/--loglkl, three likelihoods per individual; the LAST
//           n_ind*n_geno numeric fields of a line are used and non-numeric
//           tokens are dropped, so BEAGLE files work as they are
//           otherwise => raw doubles, site-major [site][individual][3]
// Every genotype triple is normalised in log space exactly as the reference
// does (twice: reader + main), with optional genotype calling in between.
#include <sys/stat.h>
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>
#include <atomic>

#include "parallel.hpp"
#include "run_state.hpp"

namespace nfh_cli {

namespace {

constexpr double kLogZero = -1e15;        // INF, gen_func.hpp:15
constexpr size_t kLineMax = 500000;       // BUFF_LEN, gen_func.hpp:17

inline double ref_max(double a, double b) { return a >= b ? a : b; }

// logsum of three values as gen_func.cpp:135-151
inline double logsum3(const double *a) {
  double top = ref_max(a[2], ref_max(a[1], a[0]));
  if (top == -INFINITY) return -INFINITY;
  double acc = 0;
  for (int g = 0; g < 3; g++) acc += exp(a[g] - top);
  return log(acc) + top;
}

inline void normalise(double *g) {        // post_prob(g, g, NULL, 3), gen_func.cpp:920-932
  const double norm = logsum3(g);
  for (int k = 0; k < 3; k++) g[k] -= norm;
}

inline void log_or_floor(double *g) {     // conv_space(g, 3, log), gen_func.cpp:123-130
  for (int k = 0; k < 3; k++) {
    g[k] = log(g[k]);
    if (g[k] == -INFINITY) g[k] = kLogZero;
  }
}

// call_geno with main()'s defaults (gen_func.cpp:886-914, ngsF-HMM.cpp:103)
inline void call_genotype(double *g) {
  int hi = 0, lo = 0;
  double top = -INFINITY, bot = INFINITY;
  for (int k = 0; k < 3; k++) {
    if (g[k] > top) { top = g[k]; hi = k; }
    if (g[k] < bot) { bot = g[k]; lo = k; }
  }
  if (g[lo] == g[hi]) {
    for (int k = 0; k < 3; k++) g[k] = log((double) 1 / 3);
  } else {
    for (int k = 0; k < 3; k++) g[k] = kLogZero;
    g[hi] = log(1);
  }
}

// chomp (gen_func.cpp:192-199) removes ONE trailing character, '\n' or '\r': a CRLF line keeps its '\r', which
// then hides the last numeric field of a GENO line from split() - reproduced, not repaired
void chomp(char *s) {
  const size_t n = strlen(s);
  if (n && (s[n - 1] == '\n' || s[n - 1] == '\r')) s[n - 1] = '\0';
}

// numeric fields of a whitespace-separated line; tokens that are not entirely a number are dropped
// (split(char*, " \t", double**), gen_func.cpp:390-417)
void numeric_fields(char *line, std::vector<double> &out) {
  out.clear();
  char *p = line;
  while (*p) {
    while (*p == ' ' || *p == '\t') p++;
    if (!*p) break;
    char *tok = p;
    while (*p && *p != ' ' && *p != '\t') p++;
    const char saved = *p;
    *p = '\0';
    char *end = nullptr;
    const double v = strtod(tok, &end);
    if (end != tok && *end == '\0') out.push_back(v);
    *p = saved;
  }
}

}  // namespace

void read_positions(RunState &st) {
  const Options &o = st.opt;
  const char *fn = "read_dist";
  gzFile fh = gzopen(o.pos.c_str(), "r");
  if (!fh) fatal("read_file", "cannot open file!");
  // read_file (gen_func.cpp:236-275): every line that is neither empty nor a '#' comment; the end-of-file test
  // comes AFTER the read, so a last line without a newline is lost (and the line count below then fails)
  std::vector<char> buf(kLineMax);
  std::string lines;                       // kept lines, each NUL-terminated
  std::vector<size_t> start;
  for (;;) {
    buf[0] = '\0';
    gzgets(fh, buf.data(), (int) kLineMax);
    if (gzeof(fh)) break;
    chomp(buf.data());
    if (buf[0] == '\0' || buf[0] == '#') continue;
    start.push_back(lines.size());
    lines.append(buf.data(), strlen(buf.data()) + 1);
  }
  gzclose(fh);
  // read_split (read_data.cpp:129-153): tab-separated, the same number of fields on every line
  if (start.empty()) fatal("read_split", "cannot open file!");
  size_t n_fields = 0;
  for (size_t r = 0; r < start.size(); r++) {
    size_t n = 1;
    for (const char *p = lines.data() + start[r]; *p; p++) n += *p == '\t';
    if (n_fields == 0) n_fields = n;
    if (n != n_fields) fatal("read_split", "invalid number of fields in file!");
  }
  if (start.size() != o.n_sites) fatal(fn, "wrong number of lines in POS file!");
  if (n_fields < 2) fatal(fn, "wrong POS file format!");

  st.dist_mb.assign(o.n_sites, INFINITY);
  std::string prev_chr;
  unsigned long prev_pos = 0;
  for (uint64_t s = 0; s < o.n_sites; s++) {
    char *chr = &lines[start[s]];
    char *tab = strchr(chr, '\t');
    *tab = '\0';
    char *pos_txt = tab + 1;
    char *tab2 = strchr(pos_txt, '\t');
    if (tab2) *tab2 = '\0';
    const double pos = strtod(pos_txt, nullptr);
    // the reference treats a zero position as a header and then never leaves its loop (read_data.cpp:188-196);
    // the one place where this reader stops with a message of its own
    if (pos == 0) fatal(fn, "header found in POS file (prefix header lines with #)");
    if (prev_chr.empty()) prev_chr = chr;
    if (prev_chr == chr) {
      st.dist_mb[s] = pos - (double) prev_pos;
      if (st.dist_mb[s] < 1) fatal(fn, "invalid distance between adjacent sites!");
    } else {
      st.dist_mb[s] = INFINITY;
      prev_chr = chr;
    }
    prev_pos = strtoul(pos_txt, nullptr, 0);
  }
  for (uint64_t i = 0; i < o.n_sites; i++) st.dist_mb[i] /= 1e6;   // bp -> Mb, ngsF-HMM.cpp:85-86
  if (o.verbose >= 7)
    for (uint64_t i = 0; i < o.n_sites && i < 10; i++) printf("%f\n", st.dist_mb[i]);
}

// main()'s look at the GENO file before anything is read (ngsF-HMM.cpp:47-66): text or binary by the file name,
// and a binary file must hold exactly n_sites x n_ind x 3 doubles
void inspect_geno_file(RunState &st) {
  Options &o = st.opt;
  struct stat sb;
  if (stat(o.geno.c_str(), &sb) != 0) fatal("main", "cannot check GENO file size!");
  const char *dot = strrchr(o.geno.c_str(), '.');
  if (dot && strcmp(dot, ".gz") == 0) {
    if (o.verbose >= 1) printf("==> GZIP input file (not BINARY)\n");
    o.in_bin = false;
  } else {
    if (o.verbose >= 1) printf("==> BINARY input file (always lkl)\n");
    o.in_bin = true;
    o.lkl = true;
    if (o.n_sites != (uint64_t) sb.st_size / sizeof(double) / o.n_ind / 3) fatal("main", "invalid/corrupt genotype input file!");
  }
}

void read_genotypes(RunState &st) {
  Options &o = st.opt;
  const char *fn = "read_geno";
  const uint64_t N = o.n_ind, S = o.n_sites;
  if (o.verbose >= 1) printf("> GENO data\n");

  st.log_gl.reset(new double[S * N * 3]);
  if (!o.in_bin) std::fill(st.log_gl.get(), st.log_gl.get() + S * N * 3, kLogZero);   // skipped lines keep this value
  gzFile fh = gzopen(o.geno.c_str(), o.in_bin ? "rb" : "r");
  if (!fh) fatal(fn, "cannot open GENO file!");
  gzbuffer(fh, 1 << 20);
  const uint64_t n_geno = o.lkl ? 3 : 1;

  if (o.in_bin) {
    // blocks of whole sites: one large read, then the reader's normalisation on all host threads
    const size_t row = N * 3 * sizeof(double);
    const uint64_t block_sites = std::max<uint64_t>(1, ((size_t) 128 << 20) / row);
    std::atomic<bool> saw_nan(false);
    for (uint64_t s0 = 0; s0 < S; s0 += block_sites) {
      const uint64_t ns = std::min(block_sites, S - s0);
      double *block = st.log_gl.get() + s0 * N * 3;
      if ((size_t) gzread(fh, block, (unsigned) (ns * row)) != ns * row) {
        if (gzeof(fh)) fatal(fn, "GENO file at premature EOF. Check GENO file and number of sites!");
        fatal(fn, "cannot read binary GENO file. Check GENO file and number of sites!");
      }
      const bool take_log = !o.loglkl;
      parallel_for(ns * N, o.host_threads, [&](uint64_t lo, uint64_t hi, unsigned) {
        for (uint64_t j = lo; j < hi; j++) {
          double *g = block + 3 * j;
          if (take_log) log_or_floor(g);
          normalise(g);
          if (std::isnan(g[0]) || std::isnan(g[1]) || std::isnan(g[2])) saw_nan = true;
        }
      });
      if (saw_nan) fatal(fn, "NaN found! Is the file format correct?");
    }
  } else {
    std::vector<char> buf(kLineMax);
    std::vector<double> fields;
    for (uint64_t s = 0; s < S; s++) {
      if (gzgets(fh, buf.data(), (int) kLineMax) == nullptr) {
        if (gzeof(fh)) fatal(fn, "GENO file at premature EOF. Check GENO file and number of sites!");
        fatal(fn, "cannot read GZip GENO file. Check GENO file and number of sites!");
      }
      chomp(buf.data());
      if (buf[0] == '\0') continue;    // an empty line leaves the site at its initial value (read_data.cpp:56-57)
      numeric_fields(buf.data(), fields);
      if (fields.empty() || (s == 0 && fields.size() < N * n_geno)) {
        fprintf(stderr, "> Header found! Skipping line...\n");
        if (s != 0) warn(fn, " header found but not on first line. Is this an error?");
        s--;                            // wraps to UINT64_MAX at s == 0 and back to 0 by the loop increment
        continue;
      }
      if (fields.size() < N * n_geno) fatal(fn, "wrong GENO file format. Less fields than expected!");
      const double *last = fields.data() + (fields.size() - N * n_geno);
      double *site = st.log_gl.get() + s * N * 3;
      for (uint64_t i = 0; i < N; i++) {
        double *g = site + 3 * i;
        if (o.lkl) {
          for (int k = 0; k < 3; k++) g[k] = o.loglkl ? last[i * 3 + k] : log(last[i * 3 + k]);
        } else {
          const int call = (int) last[i];
          if (call >= 0) {
            if (call > 2) fatal(fn, "wrong GENO file format. Genotypes must be coded as {-1,0,1,2} !");
            g[call] = log(1);
          } else {
            g[0] = g[1] = g[2] = log((double) 1 / 3);
          }
        }
        normalise(g);
      }
    }
  }
  char one;
  gzread(fh, &one, 1);
  if (!gzeof(fh)) fatal(fn, "GENO file not at EOF. Check GENO file and number of sites!");
  gzclose(fh);
  o.loglkl = true;

  // main(): optional genotype calling, then a second normalisation (ngsF-HMM.cpp:99-117)
  const bool call = o.call_geno;
  double *all = st.log_gl.get();
  parallel_for(S * N, o.host_threads, [&](uint64_t lo, uint64_t hi, unsigned) {
    for (uint64_t j = lo; j < hi; j++) {
      double *g = all + 3 * j;
      if (call) call_genotype(g);
      normalise(g);
    }
  });
}

}  // namespace nfh_cli

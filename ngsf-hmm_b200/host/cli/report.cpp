// report.cpp - output files of the drop-in binary, byte-compatible with the
// reference's print_iter (EM.cpp:293-380).  The reference opens its outputs in
// zlib's transparent mode ("wT"), i.e. they are plain text / raw binary.
//   <out>.indF : total logLkl (%.10f); per individual "F<TAB>alpha" (%.5f, %f; alpha is NA when F is
//                within 1e-5 of 0 or 1); then one allele frequency per site (%f)
//   <out>.ibd  : "//" + per-individual logLkl (%.10f, tab separated); n_ind lines of 0/1 (Viterbi);
//                n_ind lines of tab-separated IBD posteriors (%f)
//   <out>.geno : raw doubles, site-major, 3 per individual: genotype posterior under HWE with F = Viterbi state
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cmath>
#include <memory>

#include "format.hpp"
#include "parallel.hpp"
#include "run_state.hpp"

namespace nfh_cli {

namespace {

FILE *open_out(const std::string &name, const char *what) {
  FILE *fh = fopen(name.c_str(), "wb");
  if (!fh) fatal("print_iter", what);
  setvbuf(fh, nullptr, _IOFBF, 1 << 22);
  return fh;
}

}  // namespace

void write_outputs(RunState &st) {
  const Options &o = st.opt;
  const uint64_t N = o.n_ind, S = o.n_sites;
  const double eps = 1e-5;   // EPSILON

  FILE *fh = open_out(o.out + ".indF", "cannot open INDF output file!");
  fprintf(fh, "%.10f\n", st.tot_lkl);
  for (uint64_t i = 0; i < N; i++) {
    if (st.indF[i] < eps) fprintf(fh, "%.5f\tNA\n", (double) 0);
    else if (st.indF[i] > 1 - eps) fprintf(fh, "%.5f\tNA\n", (double) 1);
    else fprintf(fh, "%.5f\t%f\n", st.indF[i], st.alpha[i]);
  }
  for (uint64_t s = 0; s < S; s++) {
    char num[40];
    const int len = format_unit_f(st.freq[s], num);
    num[len] = '\n';
    fwrite(num, 1, (size_t) len + 1, fh);
  }
  fclose(fh);

  fh = open_out(o.out + ".ibd", "cannot open IBD output file!");
  fputs("//\t", fh);
  for (uint64_t i = 0; i < N; i++) fprintf(fh, i ? "\t%.10f" : "%.10f", st.ind_lkl[i]);
  fputc('\n', fh);
  std::vector<char> line(S + 1);
  for (uint64_t i = 0; i < N; i++) {
    const char *p = st.path.data() + i * S;
    for (uint64_t s = 0; s < S; s++) line[s] = (char) (p[s] + 48);
    line[S] = '\n';
    if (fwrite(line.data(), 1, S + 1, fh) != S + 1) fatal("print_iter", "cannot write PATH info to file!");
  }
  // posterior lines: formatted by all host threads, a batch of rows at a time, written in order
  {
    const unsigned T = std::max(1u, o.host_threads);
    std::vector<std::vector<char>> rows(T);
    std::vector<size_t> used(T, 0);
    for (uint64_t i0 = 0; i0 < N; i0 += T) {
      const uint64_t nb = std::min<uint64_t>(T, N - i0);
      parallel_for(nb, T, [&](uint64_t lo, uint64_t hi, unsigned) {
        for (uint64_t b = lo; b < hi; b++) {
          std::vector<char> &buf = rows[b];
          buf.resize(S * 33 + 2);                       // worst case: every value through snprintf
          const double *m = st.marg1.data() + (i0 + b) * S;
          char *w = buf.data();
          for (uint64_t s = 0; s < S; s++) {
            if (s) *w++ = '\t';
            w += format_unit_f(m[s], w);
          }
          *w++ = '\n';
          used[b] = (size_t) (w - buf.data());
        }
      });
      for (uint64_t b = 0; b < nb; b++)
        if (fwrite(rows[b].data(), 1, used[b], fh) != used[b]) fatal("print_iter", "cannot write IBD output file!");
    }
  }
  fclose(fh);

  // genotype posteriors come from the device (GL lives there): HWE prior with F = Viterbi state
  fh = open_out(o.out + ".geno", "cannot open GENO output file!");
  const size_t n_geno = S * N * 3;
  std::unique_ptr<double[]> geno(new double[n_geno]);      // no fill pass: the device writes every value
  check(st, nfh_group_set_freq(st.grp, st.freq.data()), "nfh_set_freq");
  check(st, nfh_group_geno_posterior(st.grp, st.path.data(), geno.get()), "nfh_geno_posterior");
  if (fwrite(geno.get(), sizeof(double), n_geno, fh) != n_geno)
    fatal("print_iter", "cannot write GENO output file!");
  fclose(fh);
}

}  // namespace nfh_cli

/*
 * ref_harness.cpp - C-callable window onto the UNMODIFIED reference.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the
 * reference's own translation units, taken where they lie under
 * /root/reference (never copied into this repo), into
 * oracle/_ref/libngsfhmm_ref.so by oracle/Makefile.  It lets tests/ and the
 * golden-vector generator call the reference's hot-path functions on flat
 * arrays:
 *
 *   forward / backward / viterbi / calc_emission   shared/HMM.cpp:6-154
 *   est_maf / calc_HWE / post_prob                 shared/gen_func.cpp:920-1009
 *   lkl (BFGS objective)                           EM.cpp:449-464
 *   findmax_bfgs (L-BFGS-B driver)                 shared/bfgs.cpp:83-138
 *   iter_EM / EM                                   EM.cpp:27-289
 *
 * Nothing here re-implements reference arithmetic; it only converts between
 * flat 0-based arrays and the reference's ragged 1-based double*** layout
 * (ngsF-HMM.hpp:38-49) and calls through.
 */
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <unistd.h>
#include <fcntl.h>

#include "ngsF-HMM.hpp"   /* from /root/reference (defines abs/min/max macros) */

char const *version = "oracle-harness";

/* Prototype with external linkage in EM.cpp:22 / EM.cpp:449. */
double lkl(const double *, const void *);

/* Mirror of the task descriptor the reference's objective expects as its
 * opaque `data` argument (declared file-locally at EM.cpp:6-17, so it cannot
 * be included).  Field order and types must match for lkl() to read it. */
struct ref_task {
  int type;
  double **ptr;
  double *F;
  bool F_fixed;
  double *alpha;
  bool alpha_fixed;
  double **e_prob;
  char *path;
  double *pos_dist;
  uint64_t length;
};

namespace {

/* (S x 2) flat, 0-based  ->  (S+1) rows of 2, 1-based; row 0 zeroed. */
double **ragged2(const double *flat, uint64_t S, int width) {
  double **r = new double *[S + 1];
  for (uint64_t s = 0; s <= S; s++) {
    r[s] = new double[width];
    for (int k = 0; k < width; k++) r[s][k] = (s == 0) ? 0.0 : flat[(s - 1) * width + k];
  }
  return r;
}

void free2(double **r, uint64_t S) {
  for (uint64_t s = 0; s <= S; s++) delete[] r[s];
  delete[] r;
}

/* dist (S, Mb, 0-based) -> pos_dist (S+1, 1-based, [0] = +inf as main() leaves it) */
double *shifted_dist(const double *dist, uint64_t S) {
  double *d = new double[S + 1];
  d[0] = INFINITY;
  for (uint64_t s = 0; s < S; s++) d[s + 1] = dist[s];
  return d;
}

struct quiet_stdout {
  int saved;
  explicit quiet_stdout(bool on) : saved(-1) {
    if (!on) return;
    fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~quiet_stdout() {
    if (saved < 0) return;
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
  }
};

}  // namespace

extern "C" {

/* ---- a1/a2: forward / backward (shared/HMM.cpp:6-60) ---- */
double ref_forward(uint64_t S, const double *e_prob, const double *dist, double F, double alpha,
                   double *Fw_out /* (S+1)*2 or NULL */) {
  double **e = ragged2(e_prob, S, 2);
  double **Fw = ragged2(e_prob, S, 2);
  double *d = shifted_dist(dist, S);
  double q[2] = {1 - F, F};
  double v = forward(Fw, q, alpha, e, d, S, 2);
  if (Fw_out)
    for (uint64_t s = 0; s <= S; s++) { Fw_out[2 * s] = Fw[s][0]; Fw_out[2 * s + 1] = Fw[s][1]; }
  free2(e, S); free2(Fw, S); delete[] d;
  return v;
}

double ref_backward(uint64_t S, const double *e_prob, const double *dist, double F, double alpha,
                    double *Bw_out /* (S+1)*2 or NULL */) {
  double **e = ragged2(e_prob, S, 2);
  double **Bw = ragged2(e_prob, S, 2);
  double *d = shifted_dist(dist, S);
  double q[2] = {1 - F, F};
  double v = backward(Bw, q, alpha, e, d, S, 2);
  if (Bw_out)
    for (uint64_t s = 0; s <= S; s++) { Bw_out[2 * s] = Bw[s][0]; Bw_out[2 * s + 1] = Bw[s][1]; }
  free2(e, S); free2(Bw, S); delete[] d;
  return v;
}

/* ---- a4: viterbi (shared/HMM.cpp:98-125); path_out has S entries (sites 1..S) ---- */
double ref_viterbi(uint64_t S, const double *e_prob, const double *dist, double F, double alpha,
                   char *path_out) {
  double **e = ragged2(e_prob, S, 2);
  double **Vi = ragged2(e_prob, S, 2);
  double *d = shifted_dist(dist, S);
  char *path = new char[S + 1];
  memset(path, 0, S + 1);
  double q[2] = {1 - F, F};
  double v = viterbi(Vi, q, alpha, e, path, d, S, 2);
  memcpy(path_out, path + 1, S);
  free2(e, S); free2(Vi, S); delete[] d; delete[] path;
  return v;
}

/* ---- a11: emission (shared/HMM.cpp:144-154, uint64_t overload) ---- */
double ref_calc_emission(const double *gl3, double maf, uint64_t k) {
  double g[3] = {gl3[0], gl3[1], gl3[2]};
  return calc_emission(g, maf, k);
}

/* ---- a9/a10 ---- */
void ref_calc_HWE(double *out3, double maf, double F, int log_scale) { calc_HWE(out3, maf, F, log_scale != 0); }

void ref_post_prob(double *pp3, const double *lkl3, const double *prior3_or_null) {
  double l[3] = {lkl3[0], lkl3[1], lkl3[2]};
  double p[3];
  if (prior3_or_null) { p[0] = prior3_or_null[0]; p[1] = prior3_or_null[1]; p[2] = prior3_or_null[2]; }
  post_prob(pp3, l, prior3_or_null ? p : NULL, 3);
}

/* ---- a8: est_maf (shared/gen_func.cpp:974-1009); gl is n_ind x 3 log-normalised ---- */
double ref_est_maf(uint64_t n_ind, const double *gl, const double *indF) {
  double **pdg = new double *[n_ind];
  for (uint64_t i = 0; i < n_ind; i++) pdg[i] = const_cast<double *>(gl + 3 * i);
  std::vector<double> F(indF, indF + n_ind);
  double f = est_maf(n_ind, pdg, F.data(), false);
  delete[] pdg;
  return f;
}

/* ---- a6: BFGS objective (EM.cpp:449-464); returns -logLkl ---- */
double ref_lkl(uint64_t S, const double *e_prob, const double *dist, double F, double alpha) {
  double **e = ragged2(e_prob, S, 2);
  double *d = shifted_dist(dist, S);
  ref_task t;
  memset(&t, 0, sizeof t);
  t.e_prob = e; t.pos_dist = d; t.length = S;
  double x[2] = {F, alpha};
  double v = lkl(x, &t);
  free2(e, S); delete[] d;
  return v;
}

/* ---- a7: the per-individual BFGS task, exactly the call made at EM.cpp:423-441 ---- */
static uint64_t g_lkl_calls = 0;
static double counted_lkl(const double *x, const void *data) { g_lkl_calls++; return lkl(x, data); }

void ref_bfgs_individual(uint64_t S, const double *e_prob, const double *dist, double *F, double *alpha,
                         int F_fixed, int alpha_fixed, uint64_t *n_eval_out) {
  double **e = ragged2(e_prob, S, 2);
  double *d = shifted_dist(dist, S);
  ref_task t;
  memset(&t, 0, sizeof t);
  t.type = 4; t.e_prob = e; t.pos_dist = d; t.length = S;
  double val[2] = {*F, *alpha};
  double l_bound[2] = {1 / INF, 1 / INF};
  double u_bound[2] = {1 - l_bound[0], 10};
  int lims[2] = {2, 2};
  if (F_fixed) { l_bound[0] = *F; u_bound[0] = *F; }
  if (alpha_fixed) { l_bound[1] = *alpha; u_bound[1] = *alpha; }
  g_lkl_calls = 0;
  findmax_bfgs(2, val, &t, &counted_lkl, NULL, l_bound, u_bound, lims, -1);
  *F = val[0]; *alpha = val[1];
  if (n_eval_out) *n_eval_out = g_lkl_calls;
  free2(e, S); delete[] d;
}

/* ---- generic optimiser window: findmax_bfgs on a caller-supplied objective ---- */
typedef double (*ref_objective)(const double *, const void *);
double ref_findmax_bfgs(int n, double *x, ref_objective fun, const void *data, double *lb, double *ub) {
  std::vector<int> nbd(n, 2);
  return findmax_bfgs(n, x, data, fun, NULL, lb, ub, nbd.data(), -1);
}

/* ---- whole-state window: build params as main() does (ngsF-HMM.cpp:75-135) ---- */
struct ref_state {
  params *p;
};

/*
 * gl: site-major S x N x 3 natural-log GL (the binary input layout,
 * read_data.cpp:28-31); normalised here with post_prob exactly as
 * read_geno + main do.  dist: S distances in Mb (+inf allowed).
 */
void *ref_state_create(uint64_t N, uint64_t S, const double *gl, const double *dist, const double *freq,
                       const double *indF, const double *alpha, int freq_est, int indF_fixed,
                       int alpha_fixed, int call_genotypes, unsigned n_threads, const char *out_prefix) {
  params *p = new params;
  init_pars(p);
  p->n_ind = N; p->n_sites = S;
  p->freq_est = freq_est; p->indF_fixed = indF_fixed; p->alpha_fixed = alpha_fixed;
  p->verbose = 0; p->n_threads = n_threads < 1 ? 1 : n_threads;
  p->out_prefix = strdup(out_prefix ? out_prefix : "/tmp/ngsfhmm_ref_harness");
  p->pos_dist = init_ptr(S + 1, (double) INFINITY);
  for (uint64_t s = 1; s <= S; s++) p->pos_dist[s] = dist[s - 1];
  p->geno_lkl = init_ptr(N, S + 1, N_GENO, -INF);
  for (uint64_t s = 1; s <= S; s++)
    for (uint64_t i = 0; i < N; i++) {
      for (int g = 0; g < 3; g++) p->geno_lkl[i][s][g] = gl[((s - 1) * N + i) * 3 + g];
      post_prob(p->geno_lkl[i][s], p->geno_lkl[i][s], NULL, N_GENO);      /* read_data.cpp:40 */
      if (call_genotypes) call_geno(p->geno_lkl[i][s], N_GENO);            /* ngsF-HMM.cpp:103 */
      post_prob(p->geno_lkl[i][s], p->geno_lkl[i][s], NULL, N_GENO);      /* ngsF-HMM.cpp:116 */
    }
  p->geno_lkl_s = transp_matrix(p->geno_lkl, N, S + 1);
  /* what init_output does for explicit start values (parse_args.cpp:229-419) */
  p->indF = init_ptr(N, 0.0); p->alpha = init_ptr(N, 0.0);
  for (uint64_t i = 0; i < N; i++) { p->indF[i] = indF[i]; p->alpha[i] = alpha[i]; }
  p->freq = init_ptr(S + 1, 0.01);
  p->freq[0] = -1;
  for (uint64_t s = 1; s <= S; s++) p->freq[s] = freq[s - 1];
  p->e_prob = init_ptr(N, S + 1, N_STATES, 0.0);
  for (uint64_t s = 1; s <= S; s++)
    for (uint64_t i = 0; i < N; i++)
      for (uint64_t k = 0; k < N_STATES; k++)
        p->e_prob[i][s][k] = calc_emission(p->geno_lkl[i][s], p->freq[s], k);
  p->path = init_ptr(N, S + 1, (const char *) '\0');
  p->marg_prob = init_ptr(N, S + 1, N_STATES, 0.0);
  for (uint64_t i = 0; i < N; i++) p->marg_prob[i][0][0] = p->marg_prob[i][0][1] = -1;
  p->ind_lkl = init_ptr(N, (double) -INFINITY);
  p->thread_pool = threadpool_create(p->n_threads, 2 * N, 0);
  ref_state *st = new ref_state;
  st->p = p;
  return st;
}

/* ---- f-2: input ingest, exactly what main() does between parse_cmd_args and init_output
 * (ngsF-HMM.cpp:47-117): binary/gz detection by file name, read_dist (read_data.cpp:165-218) scaled to Mb,
 * read_geno (read_data.cpp:13-116), optional call_geno and the second normalisation.
 * Outputs: gl_out site-major S x N x 3 (the layout of the binary input file), dist_out S (Mb).
 * Errors exit(-1) as in the reference - call with valid files only (error texts are compared through the
 * two binaries instead). */
void ref_read_inputs(const char *geno_file, const char *pos_file, uint64_t N, uint64_t S, int in_lkl, int in_loglkl,
                     int call_genotypes, double *gl_out, double *dist_out) {
  char *geno = strdup(geno_file), *pos = strdup(pos_file);
  bool lkl_flag = in_lkl != 0, loglkl_flag = in_loglkl != 0, in_bin;
  const char *dot = strrchr(geno, '.');
  if (dot && strcmp(dot, ".gz") == 0) in_bin = false;
  else { in_bin = true; lkl_flag = true; }
  quiet_stdout q(true);
  double *pd = read_dist(pos, 0, S);
  for (uint64_t s = 0; s < S; s++) dist_out[s] = pd[s] / 1e6;
  free_ptr((void *) pd);
  double ***g = read_geno(geno, in_bin, lkl_flag, &loglkl_flag, N, S);
  for (uint64_t i = 0; i < N; i++)
    for (uint64_t s = 1; s <= S; s++) {
      if (call_genotypes) call_geno(g[i][s], N_GENO);
      post_prob(g[i][s], g[i][s], NULL, N_GENO);
      for (int k = 0; k < 3; k++) gl_out[((s - 1) * N + i) * 3 + k] = g[i][s][k];
    }
  free_ptr((void ***) g, N, S + 1);
  free(geno); free(pos);
}

/* ---- f-3: start values, exactly init_output (parse_args.cpp:229-419) on a params with uniform GL
 * (the GL only matter for --freq e, which the GPU tests cover).  Strings as given on the command line. */
void ref_init_start_values(const char *indF_arg, const char *freq_arg, unsigned seed, uint64_t N, uint64_t S,
                           int freq_est, double *indF_out, double *alpha_out, double *freq_out) {
  params *p = new params;
  init_pars(p);
  p->n_ind = N; p->n_sites = S; p->freq_est = freq_est; p->verbose = 0; p->seed = seed;
  p->in_indF = strdup(indF_arg); p->in_freq = strdup(freq_arg);
  p->geno_lkl = init_ptr(N, S + 1, N_GENO, log((double) 1 / 3));
  p->geno_lkl_s = transp_matrix(p->geno_lkl, N, S + 1);
  {
    quiet_stdout q(true);
    init_output(p);
  }
  for (uint64_t i = 0; i < N; i++) { indF_out[i] = p->indF[i]; alpha_out[i] = p->alpha[i]; }
  for (uint64_t s = 0; s < S; s++) freq_out[s] = p->freq[s + 1];
  free_ptr((void ***) p->geno_lkl, N, S + 1);
  free_ptr((void **) p->geno_lkl_s, S + 1);
  free_ptr((void *) p->freq); free_ptr((void *) p->indF); free_ptr((void *) p->alpha);
  free_ptr((void **) p->path, N);
  free_ptr((void ***) p->marg_prob, N, S + 1);
  free_ptr((void ***) p->e_prob, N, S + 1);
  free_ptr((void *) p->ind_lkl);
  free(p->in_indF); free(p->in_freq);
  delete p;
}

/* ---- f-1: output files, exactly print_iter (EM.cpp:293-380) on a params filled from flat arrays.
 * gl: site-major S x N x 3 normalised log GL (used by the .geno posterior). */
void ref_print_iter(const char *out_prefix, uint64_t N, uint64_t S, double tot_lkl, const double *indF,
                    const double *alpha, const double *freq, const double *ind_lkl, const char *path,
                    const double *marg1, const double *gl) {
  params *p = new params;
  init_pars(p);
  p->n_ind = N; p->n_sites = S; p->tot_lkl = tot_lkl;
  p->indF = init_ptr(N, 0.0); p->alpha = init_ptr(N, 0.0); p->ind_lkl = init_ptr(N, 0.0);
  for (uint64_t i = 0; i < N; i++) { p->indF[i] = indF[i]; p->alpha[i] = alpha[i]; p->ind_lkl[i] = ind_lkl[i]; }
  p->freq = init_ptr(S + 1, 0.0);
  p->freq[0] = -1;
  for (uint64_t s = 0; s < S; s++) p->freq[s + 1] = freq[s];
  p->path = init_ptr(N, S + 1, (const char *) '\0');
  p->marg_prob = init_ptr(N, S + 1, N_STATES, 0.0);
  p->geno_lkl = init_ptr(N, S + 1, N_GENO, 0.0);
  for (uint64_t i = 0; i < N; i++)
    for (uint64_t s = 0; s < S; s++) {
      p->path[i][s + 1] = path[i * S + s];
      p->marg_prob[i][s + 1][1] = marg1[i * S + s];
      for (int g = 0; g < 3; g++) p->geno_lkl[i][s + 1][g] = gl[(s * N + i) * 3 + g];
    }
  char *prefix = strdup(out_prefix);
  print_iter(prefix, p);
  free(prefix);
  free_ptr((void ***) p->geno_lkl, N, S + 1);
  free_ptr((void ***) p->marg_prob, N, S + 1);
  free_ptr((void **) p->path, N);
  free_ptr((void *) p->freq); free_ptr((void *) p->indF); free_ptr((void *) p->alpha); free_ptr((void *) p->ind_lkl);
  delete p;
}

void ref_state_iter_EM(void *h) {
  ref_state *st = (ref_state *) h;
  quiet_stdout q(true);
  iter_EM(st->p);
}

/* Full EM() incl. final Viterbi and print_iter to <out_prefix>.{indF,ibd,geno}. */
void ref_state_run_EM(void *h, unsigned min_iters, unsigned max_iters, double min_epsilon) {
  ref_state *st = (ref_state *) h;
  st->p->min_iters = min_iters; st->p->max_iters = max_iters; st->p->min_epsilon = min_epsilon;
  quiet_stdout q(true);
  EM(st->p);
}

/* Only the final Viterbi, as dispatched at EM.cpp:110-116. */
void ref_state_viterbi(void *h) {
  params *p = ((ref_state *) h)->p;
  for (uint64_t i = 0; i < p->n_ind; i++) {
    double **Vi = init_ptr(p->n_sites + 1, N_STATES, 0.0);
    double q[2] = {1 - p->indF[i], p->indF[i]};
    viterbi(Vi, q, p->alpha[i], p->e_prob[i], p->path[i], p->pos_dist, p->n_sites, 2);
    free_ptr((void **) Vi, p->n_sites + 1);
  }
}

/* Any pointer may be NULL.  Layouts: e_prob N x S x 2, marg1 N x S, path N x S,
 * gl_norm N x S x 3 (the normalised log GL the reference computes with). */
void ref_state_get(void *h, double *freq, double *indF, double *alpha, double *ind_lkl, double *e_prob,
                   double *marg1, char *path, double *gl_norm, double *tot_lkl) {
  params *p = ((ref_state *) h)->p;
  uint64_t N = p->n_ind, S = p->n_sites;
  if (freq) for (uint64_t s = 0; s < S; s++) freq[s] = p->freq[s + 1];
  if (indF) for (uint64_t i = 0; i < N; i++) indF[i] = p->indF[i];
  if (alpha) for (uint64_t i = 0; i < N; i++) alpha[i] = p->alpha[i];
  if (ind_lkl) for (uint64_t i = 0; i < N; i++) ind_lkl[i] = p->ind_lkl[i];
  for (uint64_t i = 0; i < N; i++)
    for (uint64_t s = 0; s < S; s++) {
      if (e_prob) { e_prob[(i * S + s) * 2] = p->e_prob[i][s + 1][0]; e_prob[(i * S + s) * 2 + 1] = p->e_prob[i][s + 1][1]; }
      if (marg1) marg1[i * S + s] = p->marg_prob[i][s + 1][1];
      if (path) path[i * S + s] = p->path[i][s + 1];
      if (gl_norm) for (int g = 0; g < 3; g++) gl_norm[(i * S + s) * 3 + g] = p->geno_lkl[i][s + 1][g];
    }
  if (tot_lkl) *tot_lkl = p->tot_lkl;
}

void ref_state_set(void *h, const double *freq, const double *indF, const double *alpha, const double *e_prob) {
  params *p = ((ref_state *) h)->p;
  uint64_t N = p->n_ind, S = p->n_sites;
  if (freq) for (uint64_t s = 0; s < S; s++) p->freq[s + 1] = freq[s];
  if (indF) for (uint64_t i = 0; i < N; i++) p->indF[i] = indF[i];
  if (alpha) for (uint64_t i = 0; i < N; i++) p->alpha[i] = alpha[i];
  if (e_prob)
    for (uint64_t i = 0; i < N; i++)
      for (uint64_t s = 0; s < S; s++) {
        p->e_prob[i][s + 1][0] = e_prob[(i * S + s) * 2];
        p->e_prob[i][s + 1][1] = e_prob[(i * S + s) * 2 + 1];
      }
}

void ref_state_destroy(void *h) {
  ref_state *st = (ref_state *) h;
  params *p = st->p;
  threadpool_wait(p->thread_pool);
  threadpool_destroy(p->thread_pool, threadpool_graceful);
  free_ptr((void ***) p->geno_lkl, p->n_ind, p->n_sites + 1);
  free_ptr((void **) p->geno_lkl_s, p->n_sites + 1);
  free_ptr((void *) p->pos_dist);
  free_ptr((void *) p->freq);
  free_ptr((void **) p->path, p->n_ind);
  free_ptr((void ***) p->marg_prob, p->n_ind, p->n_sites + 1);
  free_ptr((void ***) p->e_prob, p->n_ind, p->n_sites + 1);
  free_ptr((void *) p->indF);
  free_ptr((void *) p->alpha);
  free_ptr((void *) p->ind_lkl);
  free(p->out_prefix);
  delete p;
  delete st;
}

}  /* extern "C" */

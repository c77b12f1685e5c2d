/*
 * ngsfhmm_oracle.c - CPU restatement of the ngsF-HMM EM hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see ngsfhmm_oracle.h).  Each function names the
 * reference lines it restates.  The arithmetic is kept in the reference's
 * log space and operation order on purpose: on the same libm this file is
 * bit-identical to the reference (checked by tests/test_oracle_vs_reference.py),
 * which is what lets it stand in for the reference on the GPU box.
 *
 * Parity status: PINNED against oracle/_ref (the compiled reference) and the
 * fixtures under tests/golden/.
 */
#include "ngsfhmm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* gen_func.hpp:21-23 are macros with these exact comparison directions. */
#define REF_MAX(a, b) ((a) >= (b) ? (a) : (b))
#define REF_ABS(x) ((x) >= 0 ? (x) : -(x))

/* gen_func.cpp:135-151 - max-shifted log-sum-exp; all -inf gives -inf. */
double orc_logsum(const double *a, uint64_t n) {
  double top = a[0];
  for (uint64_t i = 1; i < n; i++) top = REF_MAX(a[i], top);
  if (top == -INFINITY) return -INFINITY;
  double acc = 0;
  for (uint64_t i = 0; i < n; i++) acc += exp(a[i] - top);
  return log(acc) + top;
}

/* gen_func.cpp:55-70 */
double orc_check_interv(double v, int *nan_flag) {
  if (isnan(v)) { if (nan_flag) *nan_flag = 1; return v; }
  if (v < ORC_EPSILON) return 0;
  if (v > 1 - ORC_EPSILON) return 1;
  return v;
}

/* HMM.cpp:130-139 */
double orc_calc_trans(int k, int l, double q_l, double alpha, double d) {
  double keep = exp(-alpha * d);
  double t = (1 - keep) * q_l;
  if (k == l) t += keep;
  return log(t);
}

/* gen_func.cpp:123-130 with func = log */
static void log_in_place(double *g) {
  for (int j = 0; j < 3; j++) {
    g[j] = log(g[j]);
    if (g[j] == -INFINITY) g[j] = -ORC_INF;
  }
}

/* gen_func.cpp:123-130 with func = exp */
static void exp_in_place(double *g) {
  for (int j = 0; j < 3; j++) {
    g[j] = exp(g[j]);
    if (g[j] == -INFINITY) g[j] = -ORC_INF;
  }
}

/* gen_func.cpp:938-957 */
void orc_calc_HWE(double out[3], double maf, double F, int log_scale) {
  out[0] = pow(1 - maf, 2) + (1 - maf) * maf * F;
  out[1] = 2 * (1 - maf) * maf - 2 * (1 - maf) * maf * F;
  out[2] = pow(maf, 2) + (1 - maf) * maf * F;
  if (log_scale) log_in_place(out);
  if (F == 1) out[1] = log_scale ? -ORC_INF : 1 / ORC_INF;
}

/* gen_func.cpp:920-932 */
void orc_post_prob(double pp[3], const double lkl[3], const double *prior) {
  for (int g = 0; g < 3; g++) {
    pp[g] = lkl[g];
    if (prior) pp[g] += prior[g];
  }
  double norm = orc_logsum(pp, 3);
  for (int g = 0; g < 3; g++) pp[g] -= norm;
}

/* HMM.cpp:144-154: the state (0/1) plays the role of F in the HWE prior. */
double orc_calc_emission(const double gl[3], double maf, int k) {
  double prior[3], term[3];
  orc_calc_HWE(prior, maf, (double) k, 1);
  for (int g = 0; g < 3; g++) term[g] = gl[g] + prior[g];
  return orc_logsum(term, 3);
}

/* gen_func.cpp:974-1009.  Note num/den are NOT reset between passes and the
 * start value is always 0.01 (SURVEY.md finding 2). */
double orc_est_maf(uint64_t n_ind, const double *gl, const double *indF) {
  return orc_est_maf_counted(n_ind, gl, indF, NULL);
}

/* Same loop; *n_passes = number of times the body ran (1..101). */
double orc_est_maf_counted(uint64_t n_ind, const double *gl, const double *indF, int *n_passes) {
  int passes = 0, ran = 0;
  double num = 0, den = 0, freq = 0.01, before;
  double prior[3], pp[3];
  do {
    before = freq;
    ran++;
    for (uint64_t i = 0; i < n_ind; i++) {
      double F = indF[i];
      orc_calc_HWE(prior, freq, F, 1);
      orc_post_prob(pp, gl + 3 * i, prior);
      exp_in_place(pp);
      num += pp[1] + pp[2] * (2 - F);
      den += 2 * pp[1] + (pp[0] + pp[2]) * (2 - F);
    }
    freq = num / den;
  } while (REF_ABS(before - freq) > ORC_EPSILON && passes++ < 100);
  if (n_passes) *n_passes = ran;
  return freq;
}

/* HMM.cpp:6-28 */
double orc_forward(uint64_t S, const double *e_prob, const double *dist, double F, double alpha, double *Fw) {
  double q[2] = {1 - F, F};
  double prev[2], cur[2], tmp[2];
  for (int k = 0; k < 2; k++) prev[k] = log(q[k]);
  if (Fw) { Fw[0] = prev[0]; Fw[1] = prev[1]; }
  for (uint64_t s = 0; s < S; s++) {
    for (int l = 0; l < 2; l++) {
      for (int k = 0; k < 2; k++) {
        tmp[k] = prev[k] + orc_calc_trans(k, l, q[l], alpha, dist[s]);
        if (isnan(tmp[k])) return NAN;   /* reference aborts here (HMM.cpp:18-21) */
      }
      cur[l] = orc_logsum(tmp, 2) + e_prob[2 * s + l];
    }
    prev[0] = cur[0]; prev[1] = cur[1];
    if (Fw) { Fw[2 * (s + 1)] = cur[0]; Fw[2 * (s + 1) + 1] = cur[1]; }
  }
  return orc_logsum(prev, 2);
}

/* HMM.cpp:33-60 */
double orc_backward(uint64_t S, const double *e_prob, const double *dist, double F, double alpha, double *Bw) {
  double q[2] = {1 - F, F};
  double nxt[2] = {log(1), log(1)}, cur[2], tmp[2];
  if (Bw) { Bw[2 * S] = nxt[0]; Bw[2 * S + 1] = nxt[1]; }
  for (uint64_t s = S; s > 0; s--) {
    for (int k = 0; k < 2; k++) {
      for (int l = 0; l < 2; l++) {
        tmp[l] = orc_calc_trans(k, l, q[l], alpha, dist[s - 1]) + e_prob[2 * (s - 1) + l] + nxt[l];
        if (isnan(tmp[l])) return NAN;   /* HMM.cpp:45-48 */
      }
      cur[k] = orc_logsum(tmp, 2);
    }
    nxt[0] = cur[0]; nxt[1] = cur[1];
    if (Bw) { Bw[2 * (s - 1)] = cur[0]; Bw[2 * (s - 1) + 1] = cur[1]; }
  }
  for (int k = 0; k < 2; k++) nxt[k] += log(q[k]);
  if (Bw) { Bw[0] = nxt[0]; Bw[1] = nxt[1]; }
  return orc_logsum(nxt, 2);
}

/* HMM.cpp:98-125.  The running score of state 0 is overwritten before state 1
 * of the same site is evaluated (SURVEY.md finding 3); ties keep k = 0. */
double orc_viterbi(uint64_t S, const double *e_prob, const double *dist, double F, double alpha, char *path) {
  double q[2] = {1 - F, F};
  double score[2];
  unsigned char *from = (unsigned char *) malloc(2 * (S ? S : 1));
  for (int k = 0; k < 2; k++) score[k] = log(q[k]);
  for (uint64_t s = 0; s < S; s++)
    for (int l = 0; l < 2; l++) {
      double best = -ORC_INF;
      int arg = 0;
      for (int k = 0; k < 2; k++) {
        double cand = score[k] + orc_calc_trans(k, l, q[l], alpha, dist[s]);
        if (best < cand) { best = cand; arg = k; }
      }
      from[2 * s + l] = (unsigned char) arg;
      score[l] = best + e_prob[2 * s + l];
    }
  /* array_max_pos (gen_func.cpp:73-84): first strictly larger wins */
  int state = 0;
  double top = -INFINITY;
  for (int k = 0; k < 2; k++) if (score[k] > top) { state = k; top = score[k]; }
  double ret = score[state];
  /* reference: path[S] = state; path[s-1] = Vi[s][path[s]] (1-based). */
  for (uint64_t s = S; s > 0; s--) {
    path[s - 1] = (char) state;
    state = from[2 * (s - 1) + state];
  }
  free(from);
  return ret;
}

/* EM.cpp:449-464 */
double orc_lkl(uint64_t S, const double *e_prob, const double *dist, double F, double alpha) {
  double v;
  if (isnan(F) || isinf(F) || isnan(alpha) || isinf(alpha)) v = ORC_INF;
  else v = orc_forward(S, e_prob, dist, F, alpha, NULL);
  return -v;
}

/* EM.cpp:151-185 */
int orc_estep(uint64_t N, uint64_t S, const double *e_prob, const double *dist, const double *F,
              const double *alpha, double *marg1, double *ind_lkl) {
  int status = 0;
  double *Fw = (double *) malloc(sizeof(double) * 2 * (S + 1));
  double *Bw = (double *) malloc(sizeof(double) * 2 * (S + 1));
  for (uint64_t i = 0; i < N; i++) {
    const double *e = e_prob + 2 * S * i;
    double lf = orc_forward(S, e, dist, F[i], alpha[i], Fw);
    double lb = orc_backward(S, e, dist, F[i], alpha[i], Bw);
    if (isnan(lf) || isnan(lb)) { status = 2; continue; }
    if (REF_ABS(lf - lb) > 0.001 && status == 0) status = 1;
    ind_lkl[i] = lf;
    for (uint64_t s = 1; s <= S; s++) {
      int bad = 0;
      marg1[i * S + (s - 1)] = orc_check_interv(exp(Bw[2 * s + 1] + Fw[2 * s + 1] - lf), &bad);
      if (bad) status = 2;
    }
  }
  free(Fw); free(Bw);
  return status;
}

/* EM.cpp:224-271 (freq_est == 1, e_prob_calc == 1) */
void orc_freq_emission(uint64_t N, uint64_t S, const double *gl, const double *marg1, int update_freq,
                       double *freq, double *e_prob) {
  double *site_gl = (double *) malloc(sizeof(double) * 3 * N);
  double *site_F = (double *) malloc(sizeof(double) * N);
  for (uint64_t s = 0; s < S; s++) {
    if (update_freq) {
      for (uint64_t i = 0; i < N; i++) {
        memcpy(site_gl + 3 * i, gl + (i * S + s) * 3, 3 * sizeof(double));
        site_F[i] = marg1[i * S + s];
      }
      freq[s] = orc_est_maf(N, site_gl, site_F);
    }
    for (uint64_t i = 0; i < N; i++)
      for (int k = 0; k < 2; k++)
        e_prob[(i * S + s) * 2 + k] = orc_calc_emission(gl + (i * S + s) * 3, freq[s], k);
  }
  free(site_gl); free(site_F);
}

void orc_normalize_gl(uint64_t n, double *gl) {
  for (uint64_t j = 0; j < n; j++) {
    orc_post_prob(gl + 3 * j, gl + 3 * j, NULL);
    orc_post_prob(gl + 3 * j, gl + 3 * j, NULL);
  }
}

/* gen_func.cpp:886-914 with the defaults main() uses (ngsF-HMM.cpp:103: log scale,
 * both thresholds 0, miss_data 0): all-equal GLs stay missing (1/3 each), anything
 * else becomes a hard call on the first maximum. */
void orc_call_geno(uint64_t n, double *gl) {
  for (uint64_t j = 0; j < n; j++) {
    double *g = gl + 3 * j;
    int hi = 0, lo = 0;
    double top = -INFINITY, bot = INFINITY;
    for (int k = 0; k < 3; k++) {
      if (g[k] > top) { top = g[k]; hi = k; }
      if (g[k] < bot) { bot = g[k]; lo = k; }
    }
    if (g[lo] == g[hi]) {
      for (int k = 0; k < 3; k++) g[k] = log((double) 1 / 3);
    } else {
      for (int k = 0; k < 3; k++) g[k] = -ORC_INF;
      g[hi] = log(1);
    }
  }
}

/* Not in the reference: scaled linear-space forward-backward in long double. */
void orc_estep_extended(uint64_t S, const double *e_prob, const double *dist, double F, double alpha,
                        double *marg1_unclamped, double *lkl) {
  long double q[2] = {1.0L - (long double) F, (long double) F};
  long double *fw = (long double *) malloc(sizeof(long double) * 2 * (S + 1));
  long double *scale = (long double *) malloc(sizeof(long double) * (S + 1));
  long double loglik = 0;
  fw[0] = q[0]; fw[1] = q[1];
  for (uint64_t s = 0; s < S; s++) {
    long double keep = expl(-(long double) alpha * (long double) dist[s]);
    long double tot = fw[2 * s] + fw[2 * s + 1];
    long double a0 = ((1 - keep) * q[0] * tot + keep * fw[2 * s]) * expl((long double) e_prob[2 * s]);
    long double a1 = ((1 - keep) * q[1] * tot + keep * fw[2 * s + 1]) * expl((long double) e_prob[2 * s + 1]);
    long double z = a0 + a1;
    scale[s + 1] = z;
    loglik += logl(z);
    fw[2 * (s + 1)] = a0 / z;
    fw[2 * (s + 1) + 1] = a1 / z;
  }
  *lkl = (double) loglik;
  long double b0 = 1, b1 = 1;
  for (uint64_t s = S; s > 0; s--) {
    long double post1 = fw[2 * s + 1] * b1, post0 = fw[2 * s] * b0;
    marg1_unclamped[s - 1] = (double) (post1 / (post0 + post1));
    long double keep = expl(-(long double) alpha * (long double) dist[s - 1]);
    long double w0 = expl((long double) e_prob[2 * (s - 1)]) * b0;
    long double w1 = expl((long double) e_prob[2 * (s - 1) + 1]) * b1;
    long double mix = (1 - keep) * (q[0] * w0 + q[1] * w1);
    b0 = (mix + keep * w0) / scale[s];
    b1 = (mix + keep * w1) / scale[s];
  }
  free(fw); free(scale);
}

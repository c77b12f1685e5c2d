/*
 * ngsfhmm_oracle.h - CPU restatement of the ngsF-HMM EM hot path.
 *
 * TEST INFRASTRUCTURE ONLY: may be used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the CHECKER.  Product code never includes,
 * links or calls anything declared here.
 *
 * Parity status: PINNED.  Every function is checked against the unmodified
 * reference compiled into oracle/_ref/ (tests/test_oracle_vs_reference.py,
 * run where /root/reference exists) and against the committed fixtures in
 * tests/golden/ generated from that reference (tests/golden/make_golden.py).
 *
 * All arrays are flat, 0-based over sites (site index s here is the
 * reference's s+1).  Log-space, natural logarithms, same operation order as
 * the reference so results agree to the last bit on the same libm.
 */
#ifndef NGSFHMM_ORACLE_H
#define NGSFHMM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INF      1e15   /* gen_func.hpp:15 */
#define ORC_EPSILON  1e-5   /* gen_func.hpp:16 */

/* gen_func.cpp:135-151 */
double orc_logsum(const double *a, uint64_t n);
/* gen_func.cpp:55-70 (value clamp only; NaN reported through *nan_flag) */
double orc_check_interv(double v, int *nan_flag);
/* HMM.cpp:130-139 */
double orc_calc_trans(int k, int l, double q_l, double alpha, double d);
/* gen_func.cpp:938-957 */
void orc_calc_HWE(double out[3], double maf, double F, int log_scale);
/* gen_func.cpp:920-932; prior may be NULL; pp may alias lkl */
void orc_post_prob(double pp[3], const double lkl[3], const double *prior);
/* HMM.cpp:144-154 (uint64_t overload), k in {0,1} */
double orc_calc_emission(const double gl[3], double maf, int k);
/* gen_func.cpp:974-1009; gl is n_ind x 3 (log, normalised) */
double orc_est_maf(uint64_t n_ind, const double *gl, const double *indF);
double orc_est_maf_counted(uint64_t n_ind, const double *gl, const double *indF, int *n_passes);

/* HMM.cpp:6-28.  e_prob: S x 2 log emissions; dist: S (Mb, +inf allowed).
 * Fw: (S+1) x 2 or NULL.  Returns logsum(Fw[S]); NaN if a NaN term appears
 * (where the reference aborts). */
double orc_forward(uint64_t S, const double *e_prob, const double *dist, double F, double alpha, double *Fw);
/* HMM.cpp:33-60 */
double orc_backward(uint64_t S, const double *e_prob, const double *dist, double F, double alpha, double *Bw);
/* HMM.cpp:98-125 incl. the in-place Vi_prob update; path: S bytes (0/1) */
double orc_viterbi(uint64_t S, const double *e_prob, const double *dist, double F, double alpha, char *path);
/* EM.cpp:449-464: returns -logLkl, or -1e15 for NaN/Inf parameters */
double orc_lkl(uint64_t S, const double *e_prob, const double *dist, double F, double alpha);

/* EM.cpp:151-185 for all individuals: forward, backward, consistency check,
 * ind_lkl, clamped posterior of state 1.
 * e_prob: N x S x 2; marg1: N x S; ind_lkl: N.
 * Returns 0, 1 if |lklFw - lklBw| > 1e-3 (EM.cpp:166-170), 2 on NaN. */
int orc_estep(uint64_t N, uint64_t S, const double *e_prob, const double *dist, const double *F,
              const double *alpha, double *marg1, double *ind_lkl);

/* EM.cpp:224-271 with freq_est == 1, e_prob_calc == 1.
 * gl: N x S x 3 (individual-major, log, normalised); marg1: N x S.
 * If update_freq == 0 the given freq is kept and only emissions are refreshed
 * (what init_output does, parse_args.cpp:381-386).  e_prob: N x S x 2. */
void orc_freq_emission(uint64_t N, uint64_t S, const double *gl, const double *marg1, int update_freq,
                       double *freq, double *e_prob);

/* read_data.cpp:37-40 + ngsF-HMM.cpp:116: normalise n x 3 log GL in place (applied twice, as the reference does) */
void orc_normalize_gl(uint64_t n, double *gl);

/* gen_func.cpp:886-914 with main()'s defaults (ngsF-HMM.cpp:103); n x 3 log GL in place */
void orc_call_geno(uint64_t n, double *gl);

/* Extended-precision adjudicator (not in the reference): the same E-step in
 * scaled linear space with long double accumulation.  Used only to decide
 * which side is noisier when log-space double noise exceeds the tolerance at
 * S >= 1e5 (SURVEY.md finding 5).  Outputs UNclamped posterior. */
void orc_estep_extended(uint64_t S, const double *e_prob, const double *dist, double F, double alpha,
                        double *marg1_unclamped, double *lkl);

#ifdef __cplusplus
}
#endif
#endif

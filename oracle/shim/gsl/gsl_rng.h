/*
 * Minimal stand-in for <gsl/gsl_rng.h>, used ONLY to compile the unmodified
 * reference (fgvieira/ngsF-HMM) as a CPU oracle under oracle/_ref/.
 *
 * GSL is not installed in this image.  The reference touches GSL in exactly
 * one place: random start values in init_output() (parse_args.cpp:232-233,
 * 252-253, 310) through gsl_rng_alloc(gsl_rng_taus) / gsl_rng_set /
 * gsl_rng_uniform / gsl_rng_free.  Parity runs never use the "r" start
 * modes, so only the interface has to exist.  The generator below is a
 * three-component combined Tausworthe generator written from its published
 * description (L'Ecuyer 1996); it is NOT verified bit-for-bit against GSL.
 *
 * Test infrastructure - never included by product code.
 */
#ifndef NGSFHMM_ORACLE_GSL_RNG_SHIM_H
#define NGSFHMM_ORACLE_GSL_RNG_SHIM_H

#include <stdint.h>
#include <stdlib.h>

typedef struct { int id; } gsl_rng_type;
typedef struct { uint32_t a, b, c; } gsl_rng;

static const gsl_rng_type gsl_rng_taus_instance = { 1 };
static const gsl_rng_type *const gsl_rng_taus = &gsl_rng_taus_instance;

static inline uint32_t shim_taus_step(uint32_t s, int p, int q, uint32_t mask, int sh) {
  return ((s & mask) << sh) ^ (((s << p) ^ s) >> q);
}

static inline uint32_t shim_taus_next(gsl_rng *g) {
  g->a = shim_taus_step(g->a, 13, 19, 4294967294u, 12);
  g->b = shim_taus_step(g->b, 2, 25, 4294967288u, 4);
  g->c = shim_taus_step(g->c, 3, 11, 4294967280u, 17);
  return g->a ^ g->b ^ g->c;
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) {
  (void) t;
  return (gsl_rng *) calloc(1, sizeof(gsl_rng));
}

static inline void gsl_rng_set(gsl_rng *g, unsigned long seed) {
  uint32_t s = (uint32_t) seed;
  if (s == 0) s = 1;
  g->a = 69069u * s;
  g->b = 69069u * g->a;
  g->c = 69069u * g->b;
  for (int w = 0; w < 6; w++) (void) shim_taus_next(g);
}

static inline double gsl_rng_uniform(gsl_rng *g) {
  return shim_taus_next(g) / 4294967296.0;
}

static inline void gsl_rng_free(gsl_rng *g) { free(g); }

#endif

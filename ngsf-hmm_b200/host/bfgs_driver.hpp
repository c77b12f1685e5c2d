// bfgs_driver.hpp - the reference's numerical-gradient driver around the
// optimiser (findmax_bfgs / getgradient / Yanggradient, shared/bfgs.cpp:22-138),
// restated so that many individuals advance in lockstep around ONE batched
// objective launch per round (nfh_lkl_batch) instead of one forward() call
// per evaluation per thread (EM.cpp:423-464).
#pragma once

#include <cstdint>
#include <vector>

#include "lbfgsb.hpp"
#include "ngsfhmm_b200.h"

namespace nfh_host {

constexpr int kLbfgsMemory = 10;      // MVAL   bfgs.h:22
constexpr double kLbfgsFactr = 1e6;   // FACTR  bfgs.h:23
constexpr double kLbfgsPgtol = 1e-3;  // PGTOL  bfgs.h:24

// The evaluation points one function+gradient request needs: the point itself
// and, per coordinate, a central pair or a one-sided point when a bound is in
// the way (Yanggradient, bfgs.cpp:30-41).
struct GradientPlan {
  static constexpr int kMaxDim = 2;
  int n = 0;
  double x[kMaxDim];
  double eh[kMaxDim];
  int kind[kMaxDim];            // 0 central, 1 forward one-sided, 2 backward one-sided, 3 not evaluated
  int n_points = 0;             // including the centre (slot 0)
  double pts[1 + 2 * kMaxDim][kMaxDim];
  int hi_slot[kMaxDim], lo_slot[kMaxDim];

  void build(int n, const double *x, const double *lower, const double *upper);
  // values[slot] = objective at pts[slot]; writes the gradient with the
  // bound projection of getgradient (bfgs.cpp:58-63)
  void gradient(const double *values, const double *lower, const double *upper, double *g) const;
};

struct BfgsStats {
  uint64_t rounds = 0;          // batched launches
  uint64_t evaluations = 0;     // objective evaluations requested from the device
  uint64_t max_rounds_one_individual = 0;
};

// Generic objective for tests: same signature as the reference's callback (bfgs.h:58-61).
typedef double (*objective_fn)(const double *x, const void *data);

// findmax_bfgs equivalent on a caller-supplied objective (n <= 2). Returns -f at the end like the reference.
double minimize_with_numeric_gradient(int n, double *x, objective_fn fun, const void *data, const double *lower,
                                      const double *upper, int *n_evals);

// The F / alpha update of iter_EM (EM.cpp:188-205, task type 4 EM.cpp:423-441)
// for all individuals this context owns.  indF / alpha are updated in place.
// Bounds: F in [1e-15, 1-1e-15], alpha in [1e-15, 10]; a fixed parameter
// collapses its bounds to the current value.
// With estep_lkl_out != nullptr the E-step of the same iteration (EM.cpp:151-185) is run here as well: the
// first batched round goes through nfh_estep_with_batch (its centre points are the E-step's parameters),
// or, when nothing is optimised, through nfh_estep; estep_lkl_out[n_ind] receives ind_lkl.
// posterior_ready (optional) is called once, as soon as the E-step's posteriors are complete in their window - after
// the first round, while the remaining rounds only read the emission window - so that a multi-rank caller can start
// moving them to the frequency side behind the rest of the optimisation.
typedef void (*stage_hook)(void *user);
int bfgs_update_lockstep(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, bool F_fixed, bool alpha_fixed,
                         BfgsStats *stats, double *estep_lkl_out = nullptr, stage_hook posterior_ready = nullptr,
                         void *hook_user = nullptr);

}  // namespace nfh_host

// bfgs_driver.cpp - see bfgs_driver.hpp.
#include "bfgs_driver.hpp"

#include <cmath>
#include <memory>

namespace nfh_host {

void GradientPlan::build(int n_, const double *x_, const double *lower, const double *upper) {
  n = n_;
  n_points = 1;
  for (int j = 0; j < n; j++) { x[j] = x_[j]; pts[0][j] = x_[j]; }
  for (int i = 0; i < n; i++) {
    hi_slot[i] = lo_slot[i] = -1;
    eh[i] = std::pow(1e-8 * (std::fabs(x[i]) + 1.0), 0.67);       // bfgs.cpp:33
    double lo = x[i] - eh[i], hi = x[i] + eh[i];
    if (lower[i] >= upper[i]) {
      // Fixed coordinate: whatever the one-sided difference gives, getgradient's
      // bound projection sets it to zero (x <= lower and x >= upper both hold),
      // so the evaluation is not spent.
      kind[i] = 3;
      continue;
    }
    if (lo < lower[i]) {
      hi += eh[i];
      kind[i] = 1;
      hi_slot[i] = n_points;
      for (int j = 0; j < n; j++) pts[n_points][j] = x[j];
      pts[n_points++][i] = hi;
    } else if (hi > upper[i]) {
      lo -= eh[i];
      kind[i] = 2;
      lo_slot[i] = n_points;
      for (int j = 0; j < n; j++) pts[n_points][j] = x[j];
      pts[n_points++][i] = lo;
    } else {
      kind[i] = 0;
      lo_slot[i] = n_points;
      for (int j = 0; j < n; j++) pts[n_points][j] = x[j];
      pts[n_points++][i] = lo;
      hi_slot[i] = n_points;
      for (int j = 0; j < n; j++) pts[n_points][j] = x[j];
      pts[n_points++][i] = hi;
    }
  }
}

void GradientPlan::gradient(const double *values, const double *lower, const double *upper, double *g) const {
  const double f0 = values[0];
  for (int i = 0; i < n; i++) {
    switch (kind[i]) {
      case 0: g[i] = (values[hi_slot[i]] - values[lo_slot[i]]) / (eh[i] * 2.0); break;
      case 1: g[i] = (values[hi_slot[i]] - f0) / (eh[i] * 2.0); break;
      case 2: g[i] = (f0 - values[lo_slot[i]]) / (eh[i] * 2.0); break;
      default: g[i] = 0.0; break;
    }
    if (x[i] <= lower[i] && g[i] > 0.0) g[i] = 0.0;
    if (x[i] >= upper[i] && g[i] < 0.0) g[i] = 0.0;
  }
}

double minimize_with_numeric_gradient(int n, double *x, objective_fn fun, const void *data, const double *lower,
                                      const double *upper, int *n_evals) {
  int nbd[GradientPlan::kMaxDim] = {2, 2};
  BoxLbfgs opt(n, kLbfgsMemory, x, lower, upper, nbd, kLbfgsFactr, kLbfgsPgtol);
  int evals = 0;
  double f = 0.0;
  BoxLbfgs::Request req = opt.start();
  while (req == BoxLbfgs::Request::Evaluate) {
    GradientPlan plan;
    plan.build(n, opt.x(), lower, upper);
    double vals[1 + 2 * GradientPlan::kMaxDim], g[GradientPlan::kMaxDim];
    for (int s = 0; s < plan.n_points; s++) { vals[s] = fun(plan.pts[s], data); evals++; }
    plan.gradient(vals, lower, upper, g);
    f = vals[0];
    req = opt.advance(f, g);
  }
  for (int i = 0; i < n; i++) x[i] = opt.x()[i];
  if (n_evals) *n_evals = evals;
  return -opt.f();
}

int bfgs_update_lockstep(nfh_ctx *ctx, uint64_t n_ind, double *indF, double *alpha, bool F_fixed, bool alpha_fixed,
                         BfgsStats *stats, double *estep_lkl_out, stage_hook posterior_ready, void *hook_user) {
  // The hook fires exactly once on every successful path - also for a rank that owns no individual: a multi-rank
  // caller starts a collective in it, and a rank that skipped it would leave the others waiting.
  bool hook_due = estep_lkl_out != nullptr && posterior_ready != nullptr;
  auto fire_hook = [&]() { if (hook_due) { hook_due = false; posterior_ready(hook_user); } };
  if (F_fixed && alpha_fixed) {
    const int rc = estep_lkl_out ? nfh_estep(ctx, estep_lkl_out) : NFH_OK;
    if (rc == NFH_OK) fire_hook();
    return rc;
  }
  const double inf_inv = 1.0 / 1e15;                 // 1/INF, EM.cpp:425
  struct Slot {
    std::unique_ptr<BoxLbfgs> opt;
    GradientPlan plan;
    double lower[2], upper[2];
    bool active;
    uint64_t rounds;
    size_t first_request;
  };
  std::vector<Slot> slots(n_ind);
  const int nbd[2] = {2, 2};
  for (uint64_t i = 0; i < n_ind; i++) {
    Slot &s = slots[i];
    s.lower[0] = inf_inv; s.lower[1] = inf_inv;
    s.upper[0] = 1.0 - s.lower[0]; s.upper[1] = 10.0;
    if (F_fixed) s.lower[0] = s.upper[0] = indF[i];
    if (alpha_fixed) s.lower[1] = s.upper[1] = alpha[i];
    const double x0[2] = {indF[i], alpha[i]};
    s.opt.reset(new BoxLbfgs(2, kLbfgsMemory, x0, s.lower, s.upper, nbd, kLbfgsFactr, kLbfgsPgtol));
    s.active = s.opt->start() == BoxLbfgs::Request::Evaluate;
    s.rounds = 0;
  }
  std::vector<int32_t> req_ind;
  std::vector<double> req_F, req_a, req_out;
  BfgsStats local;
  // the E-step rides on the first round when every individual takes part in it (always, unless an optimiser
  // stops before its first evaluation)
  bool estep_pending = estep_lkl_out != nullptr;
  if (estep_pending) {
    // ... and when the optimiser's first centre point IS the parameter point the context holds: start()
    // projects x0 into the box (F = 0, F = 1 or alpha > 10 move), while the reference runs forward/backward
    // at the unprojected parameters (EM.cpp:151-185) and only lets setulb_ project for the optimiser
    bool all_active = true;
    for (uint64_t i = 0; i < n_ind; i++)
      all_active = all_active && slots[i].active && slots[i].opt->x()[0] == indF[i] && slots[i].opt->x()[1] == alpha[i];
    if (!all_active) {
      int rc = nfh_estep(ctx, estep_lkl_out);
      if (rc != NFH_OK) return rc;
      estep_pending = false;
      fire_hook();
    }
  }
  for (;;) {
    req_ind.clear(); req_F.clear(); req_a.clear();
    for (uint64_t i = 0; i < n_ind; i++) {
      Slot &s = slots[i];
      if (!s.active) continue;
      s.plan.build(2, s.opt->x(), s.lower, s.upper);
      s.first_request = req_ind.size();
      for (int p = 0; p < s.plan.n_points; p++) {
        req_ind.push_back((int32_t) i);
        req_F.push_back(s.plan.pts[p][0]);
        req_a.push_back(s.plan.pts[p][1]);
      }
    }
    if (req_ind.empty()) break;
    req_out.resize(req_ind.size());
    int rc;
    if (estep_pending) {
      rc = nfh_estep_with_batch(ctx, req_ind.size(), req_ind.data(), req_F.data(), req_a.data(), req_out.data(),
                                estep_lkl_out);
      estep_pending = false;
      if (rc == NFH_OK) fire_hook();             // the call returned: the posteriors are in
    } else {
      rc = nfh_lkl_batch(ctx, req_ind.size(), req_ind.data(), req_F.data(), req_a.data(), req_out.data());
    }
    if (rc != NFH_OK) return rc;
    local.rounds++;
    local.evaluations += req_ind.size();
    for (uint64_t i = 0; i < n_ind; i++) {
      Slot &s = slots[i];
      if (!s.active) continue;
      double g[2];
      const double *vals = req_out.data() + s.first_request;
      s.plan.gradient(vals, s.lower, s.upper, g);
      s.rounds++;
      s.active = s.opt->advance(vals[0], g) == BoxLbfgs::Request::Evaluate;
      if (!s.active && s.rounds > local.max_rounds_one_individual) local.max_rounds_one_individual = s.rounds;
    }
  }
  for (uint64_t i = 0; i < n_ind; i++) {
    indF[i] = slots[i].opt->x()[0];
    alpha[i] = slots[i].opt->x()[1];
  }
  fire_hook();                                   // no individual here (or none needed an evaluation)
  if (stats) *stats = local;
  return NFH_OK;
}

}  // namespace nfh_host

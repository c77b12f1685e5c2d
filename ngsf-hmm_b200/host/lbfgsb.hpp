// lbfgsb.hpp - bound-constrained limited-memory BFGS as a reverse-communication
// state machine (host-side bookkeeping of the F / alpha update).
//
// The reference optimises each individual's (F, alpha) with a vendored f2c
// translation of L-BFGS-B 2.1 driven by findmax_bfgs (shared/bfgs.cpp:83-138,
// setulb_ at :173).  This is an independent implementation of the same
// published algorithm (Byrd, Lu, Nocedal & Zhu, SIAM J. Sci. Comput. 16, 1995;
// Zhu, Byrd, Lu & Nocedal, ACM TOMS 23, 1997; line search of More' & Thuente,
// ACM TOMS 20, 1994) with the same constants the reference passes
// (m = 10, factr = 1e6, pgtol = 1e-3, bfgs.h:22-24; ftol = 1e-3, gtol = 0.9,
// xtol = 0.1, at most 20 line-search evaluations), so that, fed the same
// objective values, it requests the same sequence of evaluation points.
// Checked iterate-by-iterate against the reference's own optimiser in
// tests/test_lbfgsb.py.
//
// Reverse communication lets n_ind optimisers advance in lockstep around one
// batched objective launch per round (SURVEY.md finding 8).
#pragma once

#include <vector>

namespace nfh_host {

class BoxLbfgs {
 public:
  enum class Request {
    Evaluate,    // caller must supply f and g at x() through advance()
    Converged,   // projected gradient or relative reduction test met
    Abnormal,    // line search failed with no memory to discard
    Error        // invalid input
  };

  // nbd[i]: 0 unbounded, 1 lower only, 2 both, 3 upper only
  BoxLbfgs(int n, int m, const double *x0, const double *lower, const double *upper, const int *nbd,
           double factr, double pgtol);

  // First call: returns Evaluate (f, g wanted at the projected start point).
  Request start();
  // Supply f, g at x(); runs until the next evaluation is needed or the run ends.
  Request advance(double f, const double *g);

  const double *x() const { return x_.data(); }
  double f() const { return f_; }
  int iterations() const { return iter_; }
  int evaluations() const { return nfgv_; }
  const char *why() const { return why_; }

 private:
  enum class Stage { Fresh, AwaitStartFG, AwaitLineFG, Done };

  // algorithm pieces (names follow the papers' terminology)
  void classify_bounds();
  double projected_gradient_norm() const;
  bool cauchy_point();                     // generalised Cauchy point; false => singular system
  void pick_free_variables();
  bool form_reduced_system();              // LEL^T factorisation of the indefinite K matrix
  bool reduced_gradient();
  bool subspace_minimise();
  bool middle_times(const double *v, double *p) const;   // p = M v with the compact middle matrix
  bool form_t_factor();
  void store_correction();
  void forget_memory();
  // line search
  bool line_search_step();                 // returns true when another evaluation is needed
  void more_thuente(double f, double g);
  static void trial_step(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
                         double fp, double dp, bool &brackt, double stpmin, double stpmax);
  Request iterate();                       // the main loop between evaluations
  Request finish(Request r, const char *why);

  int n_, m_;
  std::vector<double> x_, l_, u_, g_;
  std::vector<int> nbd_;
  double factr_, pgtol_, f_ = 0.0;
  Stage stage_ = Stage::Fresh;
  const char *why_ = "";

  // limited-memory matrices (column j of ws/wy = j-th stored correction, circular from head_)
  std::vector<double> ws_, wy_, sy_, ss_, wt_, wn_;
  // inner products behind the reduced system, kept incrementally across iterations (see form_reduced_system)
  std::vector<double> wn1_;
  // work vectors
  std::vector<double> z_, r_, d_, t_, xcp_c_, p_, c_, wbp_, v_, wv_;
  std::vector<int> index_, iwhere_, indx2_, iorder_;
  std::vector<double> brk_;

  int col_ = 0, head_ = 0, itail_ = 0, iupdat_ = 0, iter_ = 0, nfgv_ = 0, nfree_ = 0, nenter_ = 0, ileave_ = 0;
  int ifun_ = 0, iback_ = 0, info_ = 0;
  bool updatd_ = false, prjctd_ = false, cnstnd_ = false, boxed_ = false, wrk_ = false;
  double theta_ = 1.0, fold_ = 0.0, tol_ = 0.0, dnorm_ = 0.0, epsmch_ = 0.0, gd_ = 0.0, gdold_ = 0.0, stp_ = 0.0,
         stpmx_ = 0.0, sbgnrm_ = 0.0, dtd_ = 0.0;

  // More'-Thuente state
  enum class LsTask { Start, Fg, Converged, Warning, Error } ls_task_ = LsTask::Start;
  bool brackt_ = false;
  int ls_stage_ = 1;
  double ginit_ = 0, gtest_ = 0, gx_ = 0, gy_ = 0, finit_ = 0, fx_ = 0, fy_ = 0, stx_ = 0, sty_ = 0, stmin_ = 0,
         stmax_ = 0, width_ = 0, width1_ = 0;
};

}  // namespace nfh_host

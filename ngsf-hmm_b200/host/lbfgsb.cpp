// lbfgsb.cpp - see lbfgsb.hpp.  Independent implementation of L-BFGS-B with the
// control flow of the published code (v2.1) so that it asks for the same
// evaluation points as the optimiser behind the reference's findmax_bfgs
// (shared/bfgs.cpp:83-138) when given the same objective values.
//
// Conventions: 0-based indices; matrices are column-major.
//   ws_/wy_ : n x m, column j holds correction pair j (circular buffer, oldest at head_)
//   sy_, ss_, wt_ : m x m;  wn_ : 2m x 2m
#include "lbfgsb.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace nfh_host {

namespace {

inline double dot(int n, const double *a, const double *b) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

// Cholesky factor of the leading k x k block of a (ld rows), upper triangle: a = R'R.
// Returns false if the block is not positive definite.
bool cholesky_upper(double *a, int ld, int k) {
  for (int j = 0; j < k; j++) {
    double s = 0.0;
    for (int i = 0; i < j; i++) {
      double t = a[i + j * ld] - dot(i, a + i * ld, a + j * ld);
      t /= a[i + i * ld];
      a[i + j * ld] = t;
      s += t * t;
    }
    s = a[j + j * ld] - s;
    if (s <= 0.0) return false;
    a[j + j * ld] = std::sqrt(s);
  }
  return true;
}

// Solve with the upper-triangular R stored in t (ld rows, order k).
// transposed == true : R' x = b ;  false : R x = b.  False on a zero pivot.
bool solve_upper(const double *t, int ld, int k, double *b, bool transposed) {
  for (int j = 0; j < k; j++)
    if (t[j + j * ld] == 0.0) return false;
  if (transposed) {
    b[0] /= t[0];
    for (int j = 1; j < k; j++) {
      b[j] -= dot(j, t + j * ld, b);
      b[j] /= t[j + j * ld];
    }
  } else {
    b[k - 1] /= t[(k - 1) + (k - 1) * ld];
    for (int j = k - 2; j >= 0; j--) {
      const double tmp = -b[j + 1];
      for (int i = 0; i <= j; i++) b[i] += tmp * t[i + (j + 1) * ld];
      b[j] /= t[j + j * ld];
    }
  }
  return true;
}

// One extraction step of the breakpoint heap: after the call the smallest of
// t[0..n) sits at t[n-1] and t[0..n-1) is a heap again.  first == true builds
// the heap from scratch.
void heap_pop_min(int n, double *t, int *order, bool first) {
  if (first) {
    for (int k = 1; k < n; k++) {
      double key = t[k];
      int tag = order[k];
      int i = k;
      while (i > 0) {
        int parent = (i - 1) / 2;
        if (key < t[parent]) {
          t[i] = t[parent];
          order[i] = order[parent];
          i = parent;
        } else {
          break;
        }
      }
      t[i] = key;
      order[i] = tag;
    }
  }
  if (n > 1) {
    const double out = t[0];
    const int out_tag = order[0];
    const double key = t[n - 1];
    const int tag = order[n - 1];
    int i = 0;
    for (;;) {
      int child = 2 * i + 1;
      if (child < n - 1) {
        if (child + 1 < n - 1 && t[child + 1] < t[child]) child++;
        if (t[child] < key) {
          t[i] = t[child];
          order[i] = order[child];
          i = child;
          continue;
        }
      }
      break;
    }
    t[i] = key;
    order[i] = tag;
    t[n - 1] = out;
    order[n - 1] = out_tag;
  }
}

}  // namespace

BoxLbfgs::BoxLbfgs(int n, int m, const double *x0, const double *lower, const double *upper, const int *nbd,
                   double factr, double pgtol)
    : n_(n), m_(m), x_(x0, x0 + n), l_(lower, lower + n), u_(upper, upper + n), g_(n, 0.0), nbd_(nbd, nbd + n),
      factr_(factr), pgtol_(pgtol) {
  ws_.assign((size_t) n * m, 0.0);
  wy_.assign((size_t) n * m, 0.0);
  sy_.assign((size_t) m * m, 0.0);
  ss_.assign((size_t) m * m, 0.0);
  wt_.assign((size_t) m * m, 0.0);
  wn_.assign((size_t) 4 * m * m, 0.0);
  wn1_.assign((size_t) 4 * m * m, 0.0);
  z_.assign(n, 0.0); r_.assign(n, 0.0); d_.assign(n, 0.0); t_.assign(n, 0.0); brk_.assign(n, 0.0);
  p_.assign(2 * m, 0.0); c_.assign(2 * m, 0.0); wbp_.assign(2 * m, 0.0); v_.assign(2 * m, 0.0);
  wv_.assign(2 * m, 0.0);
  index_.assign(n, 0); iwhere_.assign(n, 0); indx2_.assign(n, 0); iorder_.assign(n, 0);
}

BoxLbfgs::Request BoxLbfgs::finish(Request r, const char *why) {
  stage_ = Stage::Done;
  why_ = why;
  return r;
}

void BoxLbfgs::forget_memory() {
  info_ = 0;
  col_ = 0;
  head_ = 0;
  theta_ = 1.0;
  iupdat_ = 0;
  updatd_ = false;
}

// Project the start point into the box and classify every variable.
void BoxLbfgs::classify_bounds() {
  prjctd_ = false; cnstnd_ = false; boxed_ = true;
  for (int i = 0; i < n_; i++) {
    if (nbd_[i] > 0) {
      if (nbd_[i] <= 2 && x_[i] <= l_[i]) {
        if (x_[i] < l_[i]) { prjctd_ = true; x_[i] = l_[i]; }
      } else if (nbd_[i] >= 2 && x_[i] >= u_[i]) {
        if (x_[i] > u_[i]) { prjctd_ = true; x_[i] = u_[i]; }
      }
    }
  }
  for (int i = 0; i < n_; i++) {
    if (nbd_[i] != 2) boxed_ = false;
    if (nbd_[i] == 0) {
      iwhere_[i] = -1;                      // never bounded
    } else {
      cnstnd_ = true;
      iwhere_[i] = (nbd_[i] == 2 && u_[i] - l_[i] <= 0.0) ? 3 : 0;   // 3: always fixed
    }
  }
}

double BoxLbfgs::projected_gradient_norm() const {
  double norm = 0.0;
  for (int i = 0; i < n_; i++) {
    double gi = g_[i];
    if (nbd_[i] != 0) {
      if (gi < 0.0) {
        if (nbd_[i] >= 2) gi = std::max(x_[i] - u_[i], gi);
      } else {
        if (nbd_[i] <= 2) gi = std::min(x_[i] - l_[i], gi);
      }
    }
    norm = std::max(norm, std::fabs(gi));
  }
  return norm;
}

// p = M v, M the 2col x 2col middle matrix of the compact representation,
// through the triangular factor kept in wt_ and the SY matrix.
bool BoxLbfgs::middle_times(const double *v, double *p) const {
  const int col = col_, m = m_;
  if (col == 0) return true;
  p[col] = v[col];
  for (int i = 1; i < col; i++) {
    double sum = 0.0;
    for (int k = 0; k < i; k++) sum += sy_[i + k * m] * v[k] / sy_[k + k * m];
    p[col + i] = v[col + i] + sum;
  }
  if (!solve_upper(wt_.data(), m, col, p + col, true)) return false;
  for (int i = 0; i < col; i++) p[i] = v[i] / std::sqrt(sy_[i + i * m]);
  if (!solve_upper(wt_.data(), m, col, p + col, false)) return false;
  for (int i = 0; i < col; i++) p[i] = -p[i] / std::sqrt(sy_[i + i * m]);
  for (int i = 0; i < col; i++) {
    double sum = 0.0;
    for (int k = i + 1; k < col; k++) sum += sy_[k + i * m] * p[col + k] / sy_[i + i * m];
    p[i] += sum;
  }
  return true;
}

// Generalised Cauchy point along the projected steepest-descent path.
// Leaves the point in z_, the vector c = W'(xcp - x) in c_, the breakpoint
// bookkeeping in iwhere_.  Returns false if the middle matrix is singular.
bool BoxLbfgs::cauchy_point() {
  const int n = n_, m = m_, col = col_, col2 = 2 * col_;
  double *xcp = z_.data();
  if (sbgnrm_ <= 0.0) {
    std::copy(x_.begin(), x_.end(), xcp);
    return true;
  }
  bool bnded = true;
  int nfree = n;          // free-at-infinity variables are stacked from the end of iorder_
  int nbreak = 0, ibkmin = 0;
  double bkmin = 0.0, f1 = 0.0;
  double tl = 0.0, tu = 0.0;
  for (int i = 0; i < col2; i++) p_[i] = 0.0;

  for (int i = 0; i < n; i++) {
    const double neggi = -g_[i];
    if (iwhere_[i] != 3 && iwhere_[i] != -1) {
      if (nbd_[i] <= 2) tl = x_[i] - l_[i];
      if (nbd_[i] >= 2) tu = u_[i] - x_[i];
      const bool xlower = nbd_[i] <= 2 && tl <= 0.0;
      const bool xupper = nbd_[i] >= 2 && tu <= 0.0;
      iwhere_[i] = 0;
      if (xlower) {
        if (neggi <= 0.0) iwhere_[i] = 1;
      } else if (xupper) {
        if (neggi >= 0.0) iwhere_[i] = 2;
      } else {
        if (std::fabs(neggi) <= 0.0) iwhere_[i] = -3;
      }
    }
    int ptr = head_;
    if (iwhere_[i] != 0 && iwhere_[i] != -1) {
      d_[i] = 0.0;
    } else {
      d_[i] = neggi;
      f1 -= neggi * neggi;
      for (int j = 0; j < col; j++) {
        p_[j] += wy_[i + ptr * n] * neggi;
        p_[col + j] += ws_[i + ptr * n] * neggi;
        ptr = (ptr + 1) % m;
      }
      if (nbd_[i] <= 2 && nbd_[i] != 0 && neggi < 0.0) {
        iorder_[nbreak] = i;
        brk_[nbreak] = tl / (-neggi);
        if (nbreak == 0 || brk_[nbreak] < bkmin) { bkmin = brk_[nbreak]; ibkmin = nbreak; }
        nbreak++;
      } else if (nbd_[i] >= 2 && neggi > 0.0) {
        iorder_[nbreak] = i;
        brk_[nbreak] = tu / neggi;
        if (nbreak == 0 || brk_[nbreak] < bkmin) { bkmin = brk_[nbreak]; ibkmin = nbreak; }
        nbreak++;
      } else {
        nfree--;
        iorder_[nfree] = i;
        if (std::fabs(neggi) > 0.0) bnded = false;
      }
    }
  }

  if (theta_ != 1.0)
    for (int j = 0; j < col; j++) p_[col + j] *= theta_;
  std::copy(x_.begin(), x_.end(), xcp);
  if (nbreak == 0 && nfree == n) return true;      // d is zero: xcp = x

  for (int j = 0; j < col2; j++) c_[j] = 0.0;
  double f2 = -theta_ * f1;
  if (col > 0) {
    if (!middle_times(p_.data(), v_.data())) return false;
    f2 -= dot(col2, v_.data(), p_.data());
  }
  double dtm = -f1 / f2;
  double tsum = 0.0;
  bool skip_final_axpy = false;

  if (nbreak > 0) {
    int nleft = nbreak;
    int seg = 1;
    double tj = 0.0;
    for (;;) {
      const double tj0 = tj;
      int ibp;
      if (seg == 1) {
        tj = bkmin;
        ibp = iorder_[ibkmin];
      } else {
        if (seg == 2 && ibkmin != nbreak - 1) {
          // replace the already used breakpoint with the last one
          brk_[ibkmin] = brk_[nbreak - 1];
          iorder_[ibkmin] = iorder_[nbreak - 1];
        }
        heap_pop_min(nleft, brk_.data(), iorder_.data(), seg == 2);
        tj = brk_[nleft - 1];
        ibp = iorder_[nleft - 1];
      }
      const double dt = tj - tj0;
      if (dtm < dt) break;                         // minimiser lies inside this segment
      // otherwise fix variable ibp at its bound and move on
      tsum += dt;
      nleft--;
      seg++;
      const double dibp = d_[ibp];
      d_[ibp] = 0.0;
      double zibp;
      if (dibp > 0.0) {
        zibp = u_[ibp] - x_[ibp];
        xcp[ibp] = u_[ibp];
        iwhere_[ibp] = 2;
      } else {
        zibp = l_[ibp] - x_[ibp];
        xcp[ibp] = l_[ibp];
        iwhere_[ibp] = 1;
      }
      if (nleft == 0 && nbreak == n) {             // every variable is now fixed
        dtm = dt;
        skip_final_axpy = true;
        break;
      }
      const double dibp2 = dibp * dibp;
      f1 = f1 + dt * f2 + dibp2 - theta_ * dibp * zibp;
      f2 -= theta_ * dibp2;
      if (col > 0) {
        for (int j = 0; j < col2; j++) c_[j] += dt * p_[j];
        int ptr = head_;
        for (int j = 0; j < col; j++) {
          wbp_[j] = wy_[ibp + ptr * n];
          wbp_[col + j] = theta_ * ws_[ibp + ptr * n];
          ptr = (ptr + 1) % m;
        }
        if (!middle_times(wbp_.data(), v_.data())) return false;
        const double wmc = dot(col2, c_.data(), v_.data());
        const double wmp = dot(col2, p_.data(), v_.data());
        const double wmw = dot(col2, wbp_.data(), v_.data());
        for (int j = 0; j < col2; j++) p_[j] += -dibp * wbp_[j];
        f1 += dibp * wmc;
        f2 = f2 + dibp * 2.0 * wmp - dibp2 * wmw;
      }
      if (nleft > 0) {
        dtm = -f1 / f2;
        continue;
      } else if (bnded) {
        f1 = 0.0; f2 = 0.0; dtm = 0.0;
      } else {
        dtm = -f1 / f2;
      }
      break;
    }
  }

  if (!skip_final_axpy) {
    if (dtm <= 0.0) dtm = 0.0;
    tsum += dtm;
    for (int i = 0; i < n; i++) xcp[i] += tsum * d_[i];
  }
  if (col > 0)
    for (int j = 0; j < col2; j++) c_[j] += dtm * p_[j];
  return true;
}

// Count entering/leaving variables at the Cauchy point and index the free set.
void BoxLbfgs::pick_free_variables() {
  const int n = n_;
  nenter_ = 0;
  ileave_ = n;                              // leaving variables are stacked at indx2_[ileave_..n)
  if (iter_ > 0 && cnstnd_) {
    for (int i = 0; i < nfree_; i++) {
      const int k = index_[i];
      if (iwhere_[k] > 0) indx2_[--ileave_] = k;
    }
    for (int i = nfree_; i < n; i++) {
      const int k = index_[i];
      if (iwhere_[k] <= 0) indx2_[nenter_++] = k;
    }
  }
  wrk_ = (ileave_ < n) || (nenter_ > 0) || updatd_;
  nfree_ = 0;
  int iact = n;
  for (int i = 0; i < n; i++) {
    if (iwhere_[i] <= 0) index_[nfree_++] = i;
    else index_[--iact] = i;
  }
}

// Build and factorise the 2col x 2col matrix
//   K = [ D + Y'ZZ'Y/theta     -L_a' + R_z'  ]
//       [ -L_a + R_z           theta S'AA'S  ]
// (Z: free variables, A: active ones) as L E L'.
//
// The inner products behind K live in wn1_ and are maintained INCREMENTALLY, as the published code does
// (formk): a new pair adds its row / column computed over the current free and active sets, the older entries
// are corrected by the variables that entered or left the free set since the previous iteration, and nothing
// else ever touches them.  This is deliberately not a recomputation from the stored pairs: an iteration
// that skips the subspace step (no free variable at the Cauchy point, or an empty memory) also skips this
// bookkeeping in the reference (bfgs.cpp:986-1008), so a pair stored in such an iteration never gets its row
// and a change of the free set in such an iteration is never applied - from then on the reference minimises
// over a K that differs from the true one, and takes different steps.  Iterate-for-iterate parity needs the
// same history dependence (found by tests/test_lbfgsb.py::test_random_objectives_same_iterates: 10 of 200
// random boxed problems diverged with the recomputed K).  wn1_ also survives forget_memory(), as there.
//   wn1_ (2m x 2m, lower triangle used):  [ Y'ZZ'Y            ]      rows/cols 0..m-1   : y index
//                                         [ L_a + R_z  S'AA'S ]      rows/cols m..2m-1 : s index
bool BoxLbfgs::form_reduced_system() {
  const int n = n_, m = m_, col = col_, ld = 2 * m_;
  auto slot = [&](int j) { return (head_ + j) % m; };            // storage column of the j-th oldest pair
  auto yy = [&](int i, int j) -> double & { return wn1_[i + j * ld]; };
  auto ss = [&](int i, int j) -> double & { return wn1_[(m + i) + (m + j) * ld]; };
  auto sy = [&](int i, int j) -> double & { return wn1_[(m + i) + j * ld]; };     // s index i, y index j
  int n_old = col;                          // leading pairs that only see the change of the free set
  if (updatd_) {
    if (iupdat_ > m) {                      // the oldest pair was dropped: everything moves up and left by one
      for (int j = 0; j < m - 1; j++) {
        for (int i = j; i < m - 1; i++) {
          yy(i, j) = yy(i + 1, j + 1);
          ss(i, j) = ss(i + 1, j + 1);
        }
        for (int i = 0; i < m - 1; i++) sy(i, j) = sy(i + 1, j + 1);
      }
    }
    const int last = col - 1, pl = slot(last);
    for (int j = 0; j < col; j++) {         // row of the newest pair
      const int pj = slot(j);
      double t1 = 0.0, t2 = 0.0, t3 = 0.0;
      for (int k = 0; k < nfree_; k++) { const int v = index_[k]; t1 += wy_[v + pl * n] * wy_[v + pj * n]; }
      for (int k = nfree_; k < n; k++) {
        const int v = index_[k];
        t2 += ws_[v + pl * n] * ws_[v + pj * n];
        t3 += ws_[v + pl * n] * wy_[v + pj * n];
      }
      yy(last, j) = t1;
      ss(last, j) = t2;
      sy(last, j) = t3;                     // L_a part (active variables); the diagonal is overwritten below
    }
    for (int i = 0; i < col; i++) {         // column of the newest pair: R_z part (free variables)
      const int pi = slot(i);
      double t3 = 0.0;
      for (int k = 0; k < nfree_; k++) { const int v = index_[k]; t3 += ws_[v + pi * n] * wy_[v + pl * n]; }
      sy(i, last) = t3;
    }
    n_old = col - 1;
  }
  // older entries: add the entering variables to the free-set sums and take the leaving ones out (and the
  // other way round for the active-set sums), in the order the sets were collected
  for (int i = 0; i < n_old; i++) {
    const int pi = slot(i);
    for (int j = 0; j <= i; j++) {
      const int pj = slot(j);
      double e_yy = 0.0, e_ss = 0.0, l_yy = 0.0, l_ss = 0.0;
      for (int k = 0; k < nenter_; k++) {
        const int v = indx2_[k];
        e_yy += wy_[v + pi * n] * wy_[v + pj * n];
        e_ss += ws_[v + pi * n] * ws_[v + pj * n];
      }
      for (int k = ileave_; k < n; k++) {
        const int v = indx2_[k];
        l_yy += wy_[v + pi * n] * wy_[v + pj * n];
        l_ss += ws_[v + pi * n] * ws_[v + pj * n];
      }
      yy(i, j) = yy(i, j) + e_yy - l_yy;
      ss(i, j) = ss(i, j) - e_ss + l_ss;
    }
  }
  for (int i = 0; i < n_old; i++) {
    const int pi = slot(i);
    for (int j = 0; j < n_old; j++) {
      const int pj = slot(j);
      double e = 0.0, l = 0.0;
      for (int k = 0; k < nenter_; k++) { const int v = indx2_[k]; e += ws_[v + pi * n] * wy_[v + pj * n]; }
      for (int k = ileave_; k < n; k++) { const int v = indx2_[k]; l += ws_[v + pi * n] * wy_[v + pj * n]; }
      if (i <= j) sy(i, j) = sy(i, j) + e - l;      // R_z: over the free variables
      else sy(i, j) = sy(i, j) - e + l;             // L_a: over the active variables
    }
  }
  // upper triangle of K from the inner products
  for (int iy = 0; iy < col; iy++) {
    const int is = col + iy;
    for (int jy = 0; jy <= iy; jy++) {
      wn_[jy + iy * ld] = yy(iy, jy) / theta_;
      wn_[(col + jy) + is * ld] = ss(iy, jy) * theta_;
    }
    for (int jy = 0; jy < iy; jy++) wn_[jy + is * ld] = -sy(iy, jy);
    for (int jy = iy; jy < col; jy++) wn_[jy + is * ld] = sy(iy, jy);
    wn_[iy + iy * ld] += sy_[iy + iy * m];
  }
  // Cholesky of the (1,1) block, L^-1 applied to the (1,2) block
  if (!cholesky_upper(wn_.data(), ld, col)) { info_ = -1; return false; }
  for (int js = col; js < 2 * col; js++)
    if (!solve_upper(wn_.data(), ld, col, wn_.data() + (size_t) js * ld, true)) { info_ = -1; return false; }
  // (2,2) block += (L^-1 B)'(L^-1 B), then its Cholesky factor
  for (int is = col; is < 2 * col; is++)
    for (int js = is; js < 2 * col; js++)
      wn_[is + js * ld] += dot(col, wn_.data() + (size_t) is * ld, wn_.data() + (size_t) js * ld);
  if (!cholesky_upper(wn_.data() + col + (size_t) col * ld, ld, col)) { info_ = -2; return false; }
  return true;
}

// r = -Z'(B(xcp - x) + g) restricted to the free variables.
bool BoxLbfgs::reduced_gradient() {
  const int n = n_, m = m_, col = col_;
  if (!cnstnd_ && col > 0) {
    for (int i = 0; i < n; i++) r_[i] = -g_[i];
    return true;
  }
  for (int i = 0; i < nfree_; i++) {
    const int k = index_[i];
    r_[i] = -theta_ * (z_[k] - x_[k]) - g_[k];
  }
  if (!middle_times(c_.data(), p_.data())) { info_ = -8; return false; }
  int ptr = head_;
  for (int j = 0; j < col; j++) {
    const double a1 = p_[j];
    const double a2 = theta_ * p_[col + j];
    for (int i = 0; i < nfree_; i++) {
      const int k = index_[i];
      r_[i] = r_[i] + wy_[k + ptr * n] * a1 + ws_[k + ptr * n] * a2;
    }
    ptr = (ptr + 1) % m;
  }
  return true;
}

// Direct primal subspace minimisation over the free variables, then
// backtrack into the box along the step (v2.1 behaviour).
bool BoxLbfgs::subspace_minimise() {
  const int n = n_, m = m_, col = col_, ld = 2 * m_, nsub = nfree_;
  if (nsub <= 0) return true;
  double *d = r_.data();     // on entry the reduced gradient, on exit the subspace step
  int ptr = head_;
  for (int i = 0; i < col; i++) {
    double t1 = 0.0, t2 = 0.0;
    for (int j = 0; j < nsub; j++) {
      const int k = index_[j];
      t1 += wy_[k + ptr * n] * d[j];
      t2 += ws_[k + ptr * n] * d[j];
    }
    wv_[i] = t1;
    wv_[col + i] = theta_ * t2;
    ptr = (ptr + 1) % m;
  }
  if (!solve_upper(wn_.data(), ld, 2 * col, wv_.data(), true)) { info_ = 1; return false; }
  for (int i = 0; i < col; i++) wv_[i] = -wv_[i];
  if (!solve_upper(wn_.data(), ld, 2 * col, wv_.data(), false)) { info_ = 1; return false; }
  ptr = head_;
  for (int jy = 0; jy < col; jy++) {
    const int js = col + jy;
    for (int i = 0; i < nsub; i++) {
      const int k = index_[i];
      d[i] = d[i] + wy_[k + ptr * n] * wv_[jy] / theta_ + ws_[k + ptr * n] * wv_[js];
    }
    ptr = (ptr + 1) % m;
  }
  for (int i = 0; i < nsub; i++) d[i] /= theta_;

  double alpha = 1.0, temp1 = alpha;
  int ibd = 0;
  for (int i = 0; i < nsub; i++) {
    const int k = index_[i];
    const double dk = d[i];
    if (nbd_[k] != 0) {
      if (dk < 0.0 && nbd_[k] <= 2) {
        const double gap = l_[k] - z_[k];
        if (gap >= 0.0) temp1 = 0.0;
        else if (dk * alpha < gap) temp1 = gap / dk;
      } else if (dk > 0.0 && nbd_[k] >= 2) {
        const double gap = u_[k] - z_[k];
        if (gap <= 0.0) temp1 = 0.0;
        else if (dk * alpha > gap) temp1 = gap / dk;
      }
      if (temp1 < alpha) { alpha = temp1; ibd = i; }
    }
  }
  if (alpha < 1.0) {
    const double dk = d[ibd];
    const int k = index_[ibd];
    if (dk > 0.0) { z_[k] = u_[k]; d[ibd] = 0.0; }
    else if (dk < 0.0) { z_[k] = l_[k]; d[ibd] = 0.0; }
  }
  for (int i = 0; i < nsub; i++) z_[index_[i]] += alpha * d[i];
  return true;
}

// Upper triangle of T = theta S'S + L D^-1 L', then its Cholesky factor in wt_.
bool BoxLbfgs::form_t_factor() {
  const int m = m_, col = col_;
  for (int j = 0; j < col; j++) wt_[0 + j * m] = theta_ * ss_[0 + j * m];
  for (int i = 1; i < col; i++)
    for (int j = i; j < col; j++) {
      const int k1 = std::min(i, j);
      double acc = 0.0;
      for (int k = 0; k < k1; k++) acc += sy_[i + k * m] * sy_[j + k * m] / sy_[k + k * m];
      wt_[i + j * m] = acc + theta_ * ss_[i + j * m];
    }
  if (!cholesky_upper(wt_.data(), m, col)) { info_ = -3; return false; }
  return true;
}

// Store the newest pair (s = d_, y = r_) and update S'S and the lower part of S'Y.
void BoxLbfgs::store_correction() {
  const int n = n_, m = m_;
  const double rr = dot(n, r_.data(), r_.data());
  double dr;
  if (stp_ == 1.0) {
    dr = gd_ - gdold_;
  } else {
    dr = (gd_ - gdold_) * stp_;
    for (int i = 0; i < n; i++) d_[i] *= stp_;
  }
  // (caller has already decided the pair is accepted)
  if (iupdat_ <= m) {
    col_ = iupdat_;
    itail_ = (head_ + iupdat_ - 1) % m;
  } else {
    itail_ = (itail_ + 1) % m;
    head_ = (head_ + 1) % m;
  }
  for (int i = 0; i < n; i++) { ws_[i + itail_ * n] = d_[i]; wy_[i + itail_ * n] = r_[i]; }
  theta_ = rr / dr;
  const int col = col_;
  if (iupdat_ > m) {
    // drop the oldest pair: move the trailing blocks up and left by one
    for (int j = 0; j < col - 1; j++) {
      for (int i = 0; i <= j; i++) ss_[i + j * m] = ss_[(i + 1) + (j + 1) * m];
      for (int i = j; i < col - 1; i++) sy_[i + j * m] = sy_[(i + 1) + (j + 1) * m];
    }
  }
  int ptr = head_;
  for (int j = 0; j < col - 1; j++) {
    sy_[(col - 1) + j * m] = dot(n, d_.data(), wy_.data() + (size_t) ptr * n);
    ss_[j + (col - 1) * m] = dot(n, ws_.data() + (size_t) ptr * n, d_.data());
    ptr = (ptr + 1) % m;
  }
  ss_[(col - 1) + (col - 1) * m] = (stp_ == 1.0) ? dtd_ : stp_ * stp_ * dtd_;
  sy_[(col - 1) + (col - 1) * m] = dr;
}

// Safeguarded cubic/quadratic trial step of More' & Thuente.
void BoxLbfgs::trial_step(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
                          double fp, double dp, bool &brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / std::fabs(dx));
  double stpf;
  if (fp > fx) {
    // higher function value: the minimum is bracketed
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = std::max(std::max(std::fabs(theta), std::fabs(dx)), std::fabs(dp));
    const double ts = theta / s;
    double gamma = s * std::sqrt(ts * ts - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta;
    const double q = ((gamma - dx) + gamma) + dp;
    const double r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (std::fabs(stpc - stx) < std::fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = true;
  } else if (sgnd < 0.0) {
    // lower function value, derivatives of opposite sign: bracketed
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = std::max(std::max(std::fabs(theta), std::fabs(dx)), std::fabs(dp));
    const double ts = theta / s;
    double gamma = s * std::sqrt(ts * ts - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta;
    const double q = ((gamma - dp) + gamma) + dx;
    const double r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (std::fabs(stpc - stp) > std::fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = true;
  } else if (std::fabs(dp) < std::fabs(dx)) {
    // lower value, same sign, derivative magnitude decreases
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = std::max(std::max(std::fabs(theta), std::fabs(dx)), std::fabs(dp));
    const double ts = theta / s;
    double gamma = s * std::sqrt(std::max(0.0, ts * ts - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta;
    const double q = (gamma + (dx - dp)) + gamma;
    const double r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (std::fabs(stpc - stp) < std::fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = std::min(stp + 0.66 * (sty - stp), stpf);
      else stpf = std::max(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
      stpf = std::min(stpmax, stpf);
      stpf = std::max(stpmin, stpf);
    }
  } else {
    // lower value, same sign, derivative magnitude does not decrease
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = std::max(std::max(std::fabs(theta), std::fabs(dy)), std::fabs(dp));
      const double ts = theta / s;
      double gamma = s * std::sqrt(ts * ts - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta;
      const double q = ((gamma - dp) + gamma) + dy;
      const double r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// One call of the More'-Thuente search with (ftol, gtol, xtol) = (1e-3, 0.9, 0.1),
// stpmin = 0, stpmax = stpmx_.  Updates stp_ and ls_task_.
void BoxLbfgs::more_thuente(double f, double g) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = stpmx_;
  if (ls_task_ == LsTask::Start) {
    if (stp_ < stpmin || stp_ > stpmax || g >= 0.0 || stpmax < stpmin) { ls_task_ = LsTask::Error; return; }
    brackt_ = false;
    ls_stage_ = 1;
    finit_ = f; ginit_ = g; gtest_ = ftol * ginit_;
    width_ = stpmax - stpmin;
    width1_ = width_ / 0.5;
    stx_ = 0.0; fx_ = finit_; gx_ = ginit_;
    sty_ = 0.0; fy_ = finit_; gy_ = ginit_;
    stmin_ = 0.0;
    stmax_ = stp_ + stp_ * 4.0;
    ls_task_ = LsTask::Fg;
    return;
  }
  const double ftest = finit_ + stp_ * gtest_;
  if (ls_stage_ == 1 && f <= ftest && g >= 0.0) ls_stage_ = 2;
  LsTask verdict = LsTask::Fg;
  if (brackt_ && (stp_ <= stmin_ || stp_ >= stmax_)) verdict = LsTask::Warning;       // rounding errors
  if (brackt_ && stmax_ - stmin_ <= xtol * stmax_) verdict = LsTask::Warning;          // xtol test
  if (stp_ == stpmax && f <= ftest && g <= gtest_) verdict = LsTask::Warning;          // at stpmax
  if (stp_ == stpmin && (f > ftest || g >= gtest_)) verdict = LsTask::Warning;         // at stpmin
  if (f <= ftest && std::fabs(g) <= gtol * (-ginit_)) verdict = LsTask::Converged;
  if (verdict != LsTask::Fg) { ls_task_ = verdict; return; }

  if (ls_stage_ == 1 && f <= fx_ && f > ftest) {
    // modified function psi(stp) = f(stp) - f(0) - ftol stp f'(0)
    double fm = f - stp_ * gtest_, fxm = fx_ - stx_ * gtest_, fym = fy_ - sty_ * gtest_;
    double gm = g - gtest_, gxm = gx_ - gtest_, gym = gy_ - gtest_;
    trial_step(stx_, fxm, gxm, sty_, fym, gym, stp_, fm, gm, brackt_, stmin_, stmax_);
    fx_ = fxm + stx_ * gtest_;
    fy_ = fym + sty_ * gtest_;
    gx_ = gxm + gtest_;
    gy_ = gym + gtest_;
  } else {
    trial_step(stx_, fx_, gx_, sty_, fy_, gy_, stp_, f, g, brackt_, stmin_, stmax_);
  }
  if (brackt_) {
    if (std::fabs(sty_ - stx_) >= 0.66 * width1_) stp_ = stx_ + 0.5 * (sty_ - stx_);
    width1_ = width_;
    width_ = std::fabs(sty_ - stx_);
  }
  if (brackt_) {
    stmin_ = std::min(stx_, sty_);
    stmax_ = std::max(stx_, sty_);
  } else {
    stmin_ = stp_ + 1.1 * (stp_ - stx_);
    stmax_ = stp_ + 4.0 * (stp_ - stx_);
  }
  stp_ = std::max(stp_, stpmin);
  stp_ = std::min(stp_, stpmax);
  if ((brackt_ && (stp_ <= stmin_ || stp_ >= stmax_)) || (brackt_ && stmax_ - stmin_ <= xtol * stmax_)) stp_ = stx_;
  ls_task_ = LsTask::Fg;
}

// Line search along d_ from t_ (the point where the iteration started).
// Returns true if the objective is needed at the new x_; false when the
// search ended (accepted point in x_, or info_ != 0 on failure).
bool BoxLbfgs::line_search_step() {
  const int n = n_;
  if (stage_ != Stage::AwaitLineFG) {
    dtd_ = dot(n, d_.data(), d_.data());
    dnorm_ = std::sqrt(dtd_);
    stpmx_ = 1e10;
    if (cnstnd_) {
      if (iter_ == 0) {
        stpmx_ = 1.0;
      } else {
        for (int i = 0; i < n; i++) {
          const double a1 = d_[i];
          if (nbd_[i] != 0) {
            if (a1 < 0.0 && nbd_[i] <= 2) {
              const double a2 = l_[i] - x_[i];
              if (a2 >= 0.0) stpmx_ = 0.0;
              else if (a1 * stpmx_ < a2) stpmx_ = a2 / a1;
            } else if (a1 > 0.0 && nbd_[i] >= 2) {
              const double a2 = u_[i] - x_[i];
              if (a2 <= 0.0) stpmx_ = 0.0;
              else if (a1 * stpmx_ > a2) stpmx_ = a2 / a1;
            }
          }
        }
      }
    }
    if (iter_ == 0 && !boxed_) stp_ = std::min(1.0 / dnorm_, stpmx_);
    else stp_ = 1.0;
    t_ = x_;
    r_ = g_;
    fold_ = f_;
    ifun_ = 0;
    iback_ = 0;
    ls_task_ = LsTask::Start;
  }
  gd_ = dot(n, g_.data(), d_.data());
  if (ifun_ == 0) {
    gdold_ = gd_;
    if (gd_ >= 0.0) { info_ = -4; return false; }   // not a descent direction
  }
  more_thuente(f_, gd_);
  if (ls_task_ != LsTask::Converged && ls_task_ != LsTask::Warning) {
    ifun_++;
    nfgv_++;
    iback_ = ifun_ - 1;
    if (stp_ == 1.0) {
      x_ = z_;
    } else {
      for (int i = 0; i < n; i++) x_[i] = stp_ * d_[i] + t_[i];
    }
    return true;
  }
  return false;
}

BoxLbfgs::Request BoxLbfgs::start() {
  epsmch_ = DBL_EPSILON;
  forget_memory();
  iter_ = 0; nfgv_ = 0; nfree_ = n_;
  tol_ = factr_ * epsmch_;
  if (n_ <= 0 || m_ <= 0 || factr_ < 0.0) return finish(Request::Error, "ERROR: invalid dimensions or factr");
  for (int i = 0; i < n_; i++) {
    if (nbd_[i] < 0 || nbd_[i] > 3) return finish(Request::Error, "ERROR: INVALID NBD");
    if (nbd_[i] == 2 && l_[i] > u_[i]) return finish(Request::Error, "ERROR: NO FEASIBLE SOLUTION");
  }
  classify_bounds();
  stage_ = Stage::AwaitStartFG;
  return Request::Evaluate;
}

BoxLbfgs::Request BoxLbfgs::advance(double f, const double *g) {
  if (stage_ == Stage::Done || stage_ == Stage::Fresh) return Request::Error;
  f_ = f;
  std::copy(g, g + n_, g_.begin());
  if (stage_ == Stage::AwaitStartFG) {
    nfgv_ = 1;
    sbgnrm_ = projected_gradient_norm();
    if (sbgnrm_ <= pgtol_) return finish(Request::Converged, "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL");
    stage_ = Stage::Fresh;   // marks "not inside a line search" for iterate()
  }
  return iterate();
}

BoxLbfgs::Request BoxLbfgs::iterate() {
  const int n = n_;
  bool resume_line_search = (stage_ == Stage::AwaitLineFG);
  for (;;) {
    if (!resume_line_search) {
      // ---- new iteration: Cauchy point, free variables, subspace minimisation
      bool have_direction = false;
      while (!have_direction) {
        if (!cnstnd_ && col_ > 0) {
          z_ = x_;
          wrk_ = updatd_;
        } else {
          if (!cauchy_point()) { forget_memory(); continue; }
          pick_free_variables();
        }
        if (nfree_ != 0 && col_ != 0) {
          if (wrk_ && !form_reduced_system()) { forget_memory(); continue; }
          if (!reduced_gradient() || !subspace_minimise()) { forget_memory(); continue; }
        }
        have_direction = true;
      }
      for (int i = 0; i < n; i++) d_[i] = z_[i] - x_[i];
      stage_ = Stage::Fresh;
    }
    resume_line_search = false;

    // ---- line search
    const bool need_eval = line_search_step();
    if (info_ != 0 || iback_ >= 20) {
      // restore the previous iterate
      x_ = t_;
      g_ = r_;
      f_ = fold_;
      if (col_ == 0) {
        if (info_ == 0) { info_ = -9; nfgv_--; ifun_--; iback_--; }
        iter_++;
        return finish(Request::Abnormal, "ABNORMAL_TERMINATION_IN_LNSRCH");
      }
      if (info_ == 0) nfgv_--;
      forget_memory();
      stage_ = Stage::Fresh;
      continue;                                   // restart the iteration without memory
    }
    if (need_eval) {
      stage_ = Stage::AwaitLineFG;
      return Request::Evaluate;
    }

    // ---- the line search accepted x_: new iterate
    stage_ = Stage::Fresh;
    iter_++;
    sbgnrm_ = projected_gradient_norm();
    if (sbgnrm_ <= pgtol_) return finish(Request::Converged, "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL");
    const double ddum = std::max(std::max(std::fabs(fold_), std::fabs(f_)), 1.0);
    if (fold_ - f_ <= tol_ * ddum) {
      if (iback_ >= 10) info_ = -5;
      return finish(Request::Converged, "CONVERGENCE: REL_REDUCTION_OF_F <= FACTR*EPSMCH");
    }
    // y = g_new - g_old
    for (int i = 0; i < n; i++) r_[i] = g_[i] - r_[i];
    double dr, curv_floor;
    if (stp_ == 1.0) { dr = gd_ - gdold_; curv_floor = -gdold_; }
    else { dr = (gd_ - gdold_) * stp_; curv_floor = -gdold_ * stp_; }
    if (dr <= epsmch_ * curv_floor) {
      updatd_ = false;                            // skip the update, keep the memory
      // the published code scales d by stp before testing; d is recomputed next iteration
      continue;
    }
    updatd_ = true;
    iupdat_++;
    store_correction();
    if (!form_t_factor()) forget_memory();
  }
}

}  // namespace nfh_host

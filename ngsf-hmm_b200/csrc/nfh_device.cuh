// nfh_device.cuh - device-side building blocks shared by the hot-path kernels.
//
// The two-state HMM of ngsF-HMM (shared/HMM.cpp) has, at site s, the transition
//   T_s[k][l] = (1 - c_s) * q_l + [k == l] * c_s,   c_s = exp(-alpha * d_s)   (HMM.cpp:130-139)
// with q = (1 - F, F).  The reference runs forward/backward/Viterbi in log
// space, site by site.  Here the recursions run in SCALED LINEAR space as
// products of 2x2 matrices  M_s = T_s * diag(1, r_s)  where r_s = e1/e0 is the
// ratio of the two state emissions (the common factor e0 only adds
// sum_s log e0 to the log-likelihood and cancels in the posterior), so that a
// whole individual becomes a chunked associative scan.
#pragma once

#include <stdint.h>

#include "nfh_math.cuh"   // NFH_DEV; <cuda_runtime.h> under nvcc
#if defined(__CUDACC__)
#include "nfh_tma.cuh"
#endif

namespace nfh {

// Site tiling of the scan kernels.  Every thread owns kChunk CONSECUTIVE sites;
// kChunk is odd so that thread t reading element t*kChunk + j of a tile staged
// densely in shared memory is bank-conflict free (stride 33 doubles).
constexpr int kChunk = 33;                               // sites per thread
constexpr int kSub = 11;                                 // sub-block of a chunk held in registers (kChunk = 3 kSub)
constexpr int kScanThreads = 128;                        // threads per CTA in the scan kernels
constexpr int kTile = kChunk * kScanThreads;             // 4224 sites per CTA tile
constexpr uint32_t kTileBytes = kTile * sizeof(double);  // 33792, a multiple of 16 (TMA bulk copy)
constexpr double kLn2 = 0.693147180559945309417232121458;
constexpr double kEps = 1e-5;                            // EPSILON, gen_func.hpp:16
constexpr unsigned kFull = 0xffffffffu;

enum : int { kFlagNaN = 1, kFlagFwBw = 2 };

struct M2 {  // row-major [[a b] [c d]], acts on row vectors from the left: v' = v * M
  double a, b, c, d;
};

NFH_DEV M2 matmul(const M2 &x, const M2 &y) {
  M2 r;
  r.a = fma(x.a, y.a, x.b * y.c);
  r.b = fma(x.a, y.b, x.b * y.d);
  r.c = fma(x.c, y.a, x.d * y.c);
  r.d = fma(x.c, y.b, x.d * y.d);
  return r;
}

// 2^e as a double, e clamped to the normal range.
NFH_DEV double pow2i(int e) {
  e = max(-1000, min(1000, e));
  return __hiloint2double((1023 + e) << 20, 0);
}

// Exponent (floor(log2)) of a non-negative double from its bit pattern;
// 0 for zero/subnormal/inf/NaN so that those are left untouched.
NFH_DEV int exponent_of(double m) {
  int be = (__double2hiint(m) >> 20) & 0x7ff;
  return (be == 0 || be == 0x7ff) ? 0 : max(-1000, min(1000, be - 1023));
}

// Scale a non-negative matrix so its largest entry lies in [1, 2); returns the
// removed power of two (exact: multiplication by 2^-e does not round).
NFH_DEV int renorm(M2 &m) {
  int e = exponent_of(fmax(fmax(m.a, m.b), fmax(m.c, m.d)));
  double s = pow2i(-e);
  m.a *= s; m.b *= s; m.c *= s; m.d *= s;
  return e;
}

NFH_DEV int renorm2(double &x, double &y) {
  int e = exponent_of(fmax(x, y));
  double s = pow2i(-e);
  x *= s; y *= s;
  return e;
}

// The same two renormalisations with the largest entry found on the INTEGER pipe: for non-negative
// doubles the high words order like the values, so an IMNMX on them replaces the DSETP + selects of
// fmax() (the FP64 pipe is the bottleneck of every recursion kernel).  NaN entries have the largest
// high word and leave the scale untouched, as in renorm().
NFH_DEV int exponent_of_hi(int hi) {
  const int be = (hi >> 20) & 0x7ff;
  return (be == 0 || be == 0x7ff) ? 0 : max(-1000, min(1000, be - 1023));
}
NFH_DEV int renorm_i(M2 &m) {
  const int e = exponent_of_hi(max(max(__double2hiint(m.a), __double2hiint(m.b)),
                                   max(__double2hiint(m.c), __double2hiint(m.d))));
  const double s = pow2i(-e);
  m.a *= s; m.b *= s; m.c *= s; m.d *= s;
  return e;
}
NFH_DEV int renorm2_i(double &x, double &y) {
  const int e = exponent_of_hi(max(__double2hiint(x), __double2hiint(y)));
  const double s = pow2i(-e);
  x *= s; y *= s;
  return e;
}

// ---------------------------------------------------------------------------
// Factored site matrix.  With c = exp(-alpha d) and kappa = (1-c)/c = e^{alpha d} - 1
//   T_s = c I + (1-c) 1 q'  =  c [ I + kappa 1 q' ]
// so  M_s = T_s diag(1, r) = c * N_s,   N_s = [ I + kappa 1 q' ] diag(1, r).
// The scalar c only shifts the log-likelihood by -alpha d and cancels in the
// posterior and in every arg-max, so the recursions multiply by N_s (4 FMA +
// 2 MUL + 2 ADD per 2x2 update instead of 15 operations) and the scalars are
// summed separately (the log_scale argument of site_kappa).
//
// Chromosome starts have d = +inf (c = 0, T_s = 1 q').  They, and any site with
// alpha d > 110 ln 2, are evaluated at x = 110 ln 2 (kappa = 2^110 - 1, scalar 2^-110): one
// clamp, no selects.  The neglected term I/kappa must vanish against kappa q_l for the
// smallest q_l the optimiser can reach (1e-15 ~ 2^-50, EM.cpp:425-426): 2^-110 / 2^-50 = 2^-60,
// below half an ulp.  Growth is bounded by renormalising at least every 6 sites (2^660).
// ---------------------------------------------------------------------------
constexpr double kBigX = 76.24618986159398;           // 110 ln 2: largest exponent of the factored form

// Three evaluation tiers of kappa, chosen per (individual, tile) from alpha * (largest distance of the
// tile), so the choice is uniform over the CTA:
//   kTierFast  x < 0.0054 (< ln2/128) everywhere: expm1_pos() reduces to its polynomial (k = 0, table
//              entry 1), so expm1_small() returns the SAME BITS with 5 instead of 11 FP64 instructions;
//   kTierMid   x <= 1 everywhere: expm1_pos() without the clamp;
//   kTierSlow  anything else (chromosome starts d = +inf, NaN, large alpha d): clamp at kBigX.
// In the first two tiers nothing is clamped, so the scalar part of the tile's transition product,
// sum_s -alpha d_s, is -alpha * (sum of the tile's distances), a per-tile constant precomputed at upload.
enum : int { kTierFast = 0, kTierMid = 1, kTierSlow = 2 };
constexpr double kFastX = 0.0054;
constexpr double kMidX = 1.0;
NFH_DEV int kappa_tier(double alpha_max, double tile_dmax) {
  const double x = alpha_max * tile_dmax;               // NaN compares false twice -> slow tier
  return x < kFastX ? kTierFast : (x <= kMidX ? kTierMid : kTierSlow);
}
// sites between two renormalisations: growth per site is at most (1 + kappa) * max(1, r), r = e1/e0 <= 2^50
template <int TIER> struct TierTraits { static constexpr int kWindow = TIER == kTierSlow ? 4 : 8; };

template <int TIER>
NFH_DEV double tier_kappa(double x, const double *__restrict__ tab, double &log_scale) {
  if (TIER == kTierFast) return expm1_small(x);
  if (TIER == kTierMid) return expm1_pos(x, tab);
  const double xc = fmin(x, kBigX);
  log_scale -= xc;
  return expm1_pos(xc, tab);
}

// kappa for x = alpha * d; adds log c = -x (natural log) to log_scale.
NFH_DEV double site_kappa(double x, const double *__restrict__ tab, double &log_scale) {
  const double xc = fmin(x, kBigX);                   // fmin also maps NaN to the clamp
  log_scale -= xc;
  return expm1_pos(xc, tab);
}
NFH_DEV double site_kappa(double x, const double *__restrict__ tab) {
  return expm1_pos(fmin(x, kBigX), tab);
}

// M <- M * N_s with k0 = kappa q0, k1 = kappa q1:
//   row (x0, x1) -> ( x0 + (x0+x1) k0 ,  (x1 + (x0+x1) k1) r )
NFH_DEV void apply_site(M2 &m, double k0, double k1, double r) {
  const double s0 = m.a + m.b, s1 = m.c + m.d;
  m.a = fma(s0, k0, m.a);
  m.b = fma(s0, k1, m.b) * r;
  m.c = fma(s1, k0, m.c);
  m.d = fma(s1, k1, m.d) * r;
}

// row vector (a0, a1) <- (a0, a1) * N_s
NFH_DEV void forward_site(double &a0, double &a1, double k0, double k1, double r) {
  const double s = a0 + a1;
  a0 = fma(s, k0, a0);
  a1 = fma(s, k1, a1) * r;
}

// column vector (b0, b1) <- N_s * (b0, b1)
NFH_DEV void backward_site(double &b0, double &b1, double k0, double k1, double r) {
  const double w1 = r * b1;
  const double mix = fma(k0, b0, k1 * w1);
  b0 = b0 + mix;
  b1 = w1 + mix;
}

NFH_DEV M2 identity2() { M2 m; m.a = 1; m.b = 0; m.c = 0; m.d = 1; return m; }

#if defined(__CUDACC__)   // warp-level pieces: device only
__device__ __forceinline__ M2 shfl_down_m(const M2 &m, int off) {
  M2 r;
  r.a = __shfl_down_sync(kFull, m.a, off); r.b = __shfl_down_sync(kFull, m.b, off);
  r.c = __shfl_down_sync(kFull, m.c, off); r.d = __shfl_down_sync(kFull, m.d, off);
  return r;
}
__device__ __forceinline__ M2 shfl_up_m(const M2 &m, int off) {
  M2 r;
  r.a = __shfl_up_sync(kFull, m.a, off); r.b = __shfl_up_sync(kFull, m.b, off);
  r.c = __shfl_up_sync(kFull, m.c, off); r.d = __shfl_up_sync(kFull, m.d, off);
  return r;
}

// Ordered product over the warp: lane 0 ends with P_0 P_1 ... P_31 (and the sum
// of exponents).  Other lanes hold partial results.
__device__ __forceinline__ void warp_ordered_product(M2 &m, int &e) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    M2 o = shfl_down_m(m, off);
    int oe = __shfl_down_sync(kFull, e, off);
    if ((lane & (2 * off - 1)) == 0) {
      m = matmul(m, o);
      e += oe + renorm(m);
    }
  }
}
#endif

// Address of (individual row, site) in the site-blocked layout
// [n_ranks][n_ind_local][site_block]; a tile never straddles a block because
// site_block is a multiple of kTile.
NFH_DEV size_t blocked_index(uint64_t row, uint64_t site, uint64_t n_rows, uint64_t site_block) {
  uint64_t blk = site / site_block;
  uint64_t off = site - blk * site_block;
  return (size_t) ((blk * n_rows + row) * site_block + off);
}

}  // namespace nfh

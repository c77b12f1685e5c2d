// nfh_estep_math.cuh - the per-thread bodies of the E-step kernels (nfh_estep.cu): one chunk of kChunk consecutive
// sites of one individual as pure arithmetic on pointers and registers - no thread indices, no shuffles, no
// barriers.  estep_chunk_products calls products_chunk(), estep_chunk_apply calls apply_chunk().
//
// They live in a header of their own so that tests/device_arith_host.cpp can compile exactly these functions with
// g++ (NFH_DEV, see nfh_math.cuh) and the CPU test suite can check the scaled linear-space arithmetic - the factored
// transition, the renormalisation bookkeeping, the one-reciprocal posterior - against the reference's log-space code
// (forward/backward, HMM.cpp:6-60; posterior and clamp, EM.cpp:178-185, gen_func.cpp:55-70) without a GPU.
#pragma once

#include "nfh_device.cuh"

namespace nfh {

// estep_chunk_apply keeps 33 values per thread between its two phases; the kStash oldest of them live in
// shared memory ([value][thread], conflict free) so that the rest fits the 168 registers of 3 CTAs per SM.
constexpr int kStash = 8;

// M <- M * N_s, kappa not yet multiplied into q: 2 ADD-free form
//   row (x0, x1) -> ( x0 + (x0+x1) kappa q0 ,  (x1 + (x0+x1) kappa q1) r )
NFH_DEV void apply_site_k(M2 &m, double kap, double q0, double q1, double r) {
  const double t0 = (m.a + m.b) * kap, t1 = (m.c + m.d) * kap;
  m.a = fma(t0, q0, m.a);
  m.b = fma(t0, q1, m.b) * r;
  m.c = fma(t1, q0, m.c);
  m.d = fma(t1, q1, m.d) * r;
}
NFH_DEV void forward_site_k(double &a0, double &a1, double kap, double q0, double q1, double r) {
  const double t = (a0 + a1) * kap;
  a0 = fma(t, q0, a0);
  a1 = fma(t, q1, a1) * r;
}
NFH_DEV void backward_site_k(double &b0, double &b1, double kap, double q0, double q1, double r) {
  const double w1 = r * b1;
  const double mix = fma(q0, b0, q1 * w1) * kap;
  b0 = b0 + mix;
  b1 = w1 + mix;
}

// One thread's chunk: two independent chains (sites 0..15 and 16..32) whose products are multiplied at
// the end, so that four FMA chains and two kappa evaluations are in flight per thread.
template <int TIER>
NFH_DEV void products_chunk(const double *__restrict__ r, const double *__restrict__ d,
                                               const double *__restrict__ tab, double al, double q0, double q1,
                                               M2 &m, int &e, double &ls) {
  constexpr int H = kChunk / 2;                     // 16
  constexpr int W = TierTraits<TIER>::kWindow;
  static_assert(H % W == 0, "half chunk is a whole number of renormalisation windows");
  M2 lo = identity2(), hi = identity2();
  e = 0;
  ls = 0.0;
#pragma unroll 1
  for (int w0 = 0; w0 < H; w0 += W) {
#pragma unroll
    for (int i = 0; i < W; i++) {
      const int j = w0 + i;
      apply_site_k(lo, tier_kappa<TIER>(al * d[j], tab, ls), q0, q1, r[j]);
      apply_site_k(hi, tier_kappa<TIER>(al * d[H + j], tab, ls), q0, q1, r[H + j]);
    }
    e += renorm_i(lo) + renorm_i(hi);
  }
  apply_site_k(hi, tier_kappa<TIER>(al * d[2 * H], tab, ls), q0, q1, r[2 * H]);
  m = matmul(lo, hi);
  e += renorm_i(m);
}

// check_interv (gen_func.cpp:55-70) on the integer pipe: p is non-negative, so its bit pattern orders
// like its value.  NaN raises the flag (reference: error("value is NaN!")).
NFH_DEV double clamp_posterior(double p, bool &bad) {
  const long long bits = __double_as_longlong(p);
  constexpr long long kLo = 0x3ee4f8b588e368f1ll;    // bits of 1e-5 (EPSILON, gen_func.hpp:16)
  constexpr long long kHi = 0x3fefffeb074a771dll;    // bits of 1.0 - 1e-5
  bad |= (unsigned long long) bits > 0x7ff0000000000000ull;
  p = bits < kLo ? 0.0 : p;
  p = bits > kHi ? 1.0 : p;
  return p;
}

// One thread's 33 sites, given the forward vector entering the chunk (a) and the backward vector
// leaving it (b).  Sites 0..15 are H1, sites 16..32 are H2.
//   phase 1: the forward chain walks H1 left to right and keeps f_j(1); independently the backward chain
//            walks H2 right to left and keeps b_j(1).  kappa_j replaces d_j in shared memory.
//   middle : b_15 = N_16 b_16, L = f_15 . b_15, one reciprocal.
//   phase 2: the forward chain walks H2 and emits p_j = f_j(1) * kept b_j(1) / L; the backward chain
//            walks H1 and emits p_j = kept f_j(1) * b_j(1) / L.
// Every renormalisation is an exact power of two 2^-e; a value kept before it is too large by 2^e
// relative to L, a chain renormalised after L was taken is too small by 2^e: both corrections are
// folded into the running factors fF / fB (exact multiplications).
template <int TIER>
NFH_DEV bool apply_chunk(double *__restrict__ r, double *__restrict__ d,
                                            const double *__restrict__ tab, double al, double q0, double q1,
                                            double a0, double a1, double b0, double b1,
                                            double *__restrict__ stash) {
  constexpr int H = kChunk / 2;                     // 16
  constexpr int W = TierTraits<TIER>::kWindow;
  constexpr int NW = H / W;
  constexpr int KS = kStash / 2;                    // y[0..KS) and w(H-KS..H] live in shared memory
  static_assert(H % W == 0 && kChunk == 2 * H + 1, "chunk layout");
  double y[H];          // f_j(1), j = 0..15
  double w[H + 1];      // b_j(1), j = 16..32
  int ef[NW], eb[NW];
  double unused = 0.0;
  bool bad = false;

  stash[KS * kScanThreads] = b1;                    // w[H]
#pragma unroll
  for (int i = 0; i < H; i++) {
    {
      const double kap = tier_kappa<TIER>(al * d[i], tab, unused);
      d[i] = kap;
      forward_site_k(a0, a1, kap, q0, q1, r[i]);                    // f_i
    }
    {
      const int j = 2 * H - i;
      const double kap = tier_kappa<TIER>(al * d[j], tab, unused);
      d[j] = kap;
      backward_site_k(b0, b1, kap, q0, q1, r[j]);                   // b_{j-1}
    }
    if ((i + 1) % W == 0) { ef[i / W] = renorm2_i(a0, a1); eb[i / W] = renorm2_i(b0, b1); }
    if (i < KS) stash[i * kScanThreads] = a1; else y[i] = a1;
    if (i < KS - 1) stash[(KS + 1 + i) * kScanThreads] = b1; else w[H - 1 - i] = b1;
  }

  // (a0, a1) = f_15, (b0, b1) = b_16
  const double kap_mid = tier_kappa<TIER>(al * d[H], tab, unused);
  const double r_mid = r[H];
  double g0 = b0, g1 = b1;                                          // backward chain continues on a copy
  backward_site_k(g0, g1, kap_mid, q0, q1, r_mid);                  // b_15
  const double inv = rcp_pos<true>(fma(a0, g0, a1 * g1));           // 1 / L
  double fF = inv, fB = inv;

  // phase 2, fully unrolled: step k of the forward chain handles site 16 + k, step k of the backward
  // chain handles site 15 - k.
#pragma unroll
  for (int k = 0; k <= H; k++) {
    {                                                               // forward chain, site j = 16 + k
      const int j = H + k;
      if (k >= 1 && (k - 1) % W == 0 && (k - 1) / W < NW) fF *= pow2i(-eb[NW - 1 - (k - 1) / W]);
      const double kap = k == 0 ? kap_mid : d[j];
      const double rj = k == 0 ? r_mid : r[j];
      forward_site_k(a0, a1, kap, q0, q1, rj);
      if ((k + 1) % W == 0) fF *= pow2i(renorm2_i(a0, a1));
      const double wk = k > H - KS ? stash[(KS + H - k) * kScanThreads] : w[k];
      r[j] = clamp_posterior((a1 * wk) * fF, bad);
    }
    if (k < H) {                                                    // backward chain, site i = 15 - k, holds b_i
      const int i = H - 1 - k;
      if (k >= 1 && (k - 1) % W == 0 && (k - 1) / W < NW) fB *= pow2i(-ef[NW - 1 - (k - 1) / W]);
      const double ri = r[i];
      const double yi = i < KS ? stash[i * kScanThreads] : y[i];
      r[i] = clamp_posterior((yi * g1) * fB, bad);
      if (i > 0) {
        backward_site_k(g0, g1, d[i], q0, q1, ri);                  // b_{i-1}
        if ((k + 2) % W == 0) fB *= pow2i(renorm2_i(g0, g1));       // N_16 was step 0 of this chain
      }
    }
  }
  return bad;
}

}  // namespace nfh

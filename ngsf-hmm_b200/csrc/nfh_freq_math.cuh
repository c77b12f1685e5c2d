// nfh_freq_math.cuh - the arithmetic of the per-site allele-frequency EM and of the emission refresh
// (kernels: nfh_freq.cu) as pure functions on registers: an individual's pass-invariant coefficients, one pass over
// a lane's individuals at allele odds t, the state emissions.  No thread indices, no shuffles, no memory.  (The scalar
// update that follows the lane reduction - six statements - stays inline in the kernels: as a function it changed
// the register allocation of the widest instantiations.)
//
//   est_maf()        shared/gen_func.cpp:974-1009  (with calc_HWE :938-957, post_prob :920-932)
//   calc_emission()  shared/HMM.cpp:144-154
//
// A header of its own so that tests/device_arith_host.cpp can compile exactly these functions with g++ (NFH_DEV, see
// nfh_math.cuh): the CPU test suite then checks this restatement - linear space, allele odds, running
// (denominator - numerator), four-way reciprocals - against the reference's log-space est_maf without a GPU.
#pragma once

#include "nfh_device.cuh"

namespace nfh {

struct IndCoef {   // pass-invariant coefficients of one individual at one site
  double a0, a2, h;      // S = a0 u + a2 v + h a          (sum of the three genotype weights)
  double na, nv, da;     // numerator weights: (na a + nv v) / S ; F-weighted het term: da a / S
  double g;              // 2 - F (only summed once per site, not used per pass)
};

// IEEE-754 binary64 exponent of a positive normal x (0 for zero, subnormal, inf and NaN) and x scaled into [1, 2)
NFH_DEV double split_exponent(double x, int &e) {
  const int hi = __double2hiint(x);
  const int field = (hi >> 20) & 0x7ff;
  e = (field == 0 || field == 0x7ff) ? 0 : field - 1023;
  return __hiloint2double(hi - (e << 20), __double2loint(x));
}

// With u = (1-f)^2, v = f^2, a = f(1-f), GL (L0,L1,L2), IBD posterior F, g = 2 - F:
//   w0 = L0 (u + a F)   w1 = 2 L1 (1-F) a   w2 = L2 (v + a F)           (HWE prior x GL)
//   S  = w0 + w1 + w2 = L0 u + L2 v + (L0 F + 2 L1 (1-F) + L2 F) a
//   num += (w1 + g w2) / S      = [ (2 L1 (1-F) + g L2 F) a + g L2 v ] / S
//   den += (2 w1 + (w0+w2) g)/S = g + F w1 / S  = g + [ 2 L1 (1-F) F a ] / S
// so per pass and individual only S, 1/S and three FMAs into
//   A1 = sum na/S,  A2 = sum nv/S,  A3 = sum da/S
// are needed; num = a A1 + v A2 and den = sum g + a A3 are formed once per pass.
// The register kernels divide everything by u and work with the allele odds
// t = f/(1-f) = num/(den-num):  S/u = a0 + h t + a2 t^2 is two Horner FMAs, so an
// individual costs 8 FP64 instructions per pass (2 + 3 for its share of a
// four-way reciprocal + 3), and num = t (A1 + t A2), den = sum g + t A3.
NFH_DEV IndCoef make_coef(double L0, double L1, double L2, double F) {
  IndCoef k;
  double c1 = 2.0 * L1 * (1.0 - F);
  // A heterozygote call (L0 = L2 = 0) with an IBD posterior of exactly 1 has zero weight for every genotype.
  // INTENTIONAL DIVERGENCE, unreachable through the EM: a hard het call makes the emission ratio e1/e0 zero,
  // which forces the posterior to 0, so only a caller that writes its own posterior window can get here.  The
  // reference's log-space code (-1e15 standing for log 0, quantised at 0.125 at that magnitude) then yields
  // genotype weights roughly proportional to (1-f, 1, f) (num += (1+f)/2, den += 1.5); here the individual is
  // taken as certainly heterozygous (num += 1, den += 2): keep a vanishing het weight so the ratio is 1, not 0/0.
  if (L0 == 0.0 && L2 == 0.0 && F == 1.0) c1 = 1e-30;   // products of four S must stay normal
  k.g = 2.0 - F;
  k.a0 = L0; k.a2 = L2;
  k.h = L0 * F + c1 + L2 * F;
  k.na = c1 + k.g * (L2 * F);
  k.nv = k.g * L2;
  k.da = c1 * F;
  return k;
}

NFH_DEV IndCoef null_coef() {   // padding slot: contributes exactly 0
  IndCoef k;
  k.a0 = 0.5; k.a2 = 0.5; k.h = 0.0; k.na = 0.0; k.nv = 0.0; k.da = 0.0; k.g = 0.0;
  return k;
}

// u/v/a form used by the streaming path: 9 FP64 instructions + 1 MUFU, no branches.
NFH_DEV void accumulate(const IndCoef &k, double u, double v, double a, double &A1, double &A2,
                                           double &A3) {
  const double S = fma(k.a0, u, fma(k.a2, v, k.h * a));
  const double rinv = rcp_pos(S);
  A1 = fma(k.na, rinv, A1);
  A2 = fma(k.nv, rinv, A2);
  A3 = fma(k.da, rinv, A3);
}

// Reciprocals of S[0..K) with ONE hardware seed per group of four: 1/S_i is
// recovered from 1/(S_0 S_1 S_2 S_3) by multiplications.  Same FP64 instruction
// count as four separate Newton sequences (12 per group) but a quarter of the
// MUFU traffic: MUFU and SHFL share the SM's MIO queue, and with one MUFU per
// individual the shuffle reduction at the end of every pass waited behind the
// other warp's seeds (measured: removing the shuffles cut the kernel time by
// 45 %, removing the division or the vote changed nothing).
template <int K>
NFH_DEV void reciprocals(const double (&S)[K], double (&inv)[K]) {
#pragma unroll
  for (int k = 0; k + 3 < K; k += 4) {
    const double p01 = S[k] * S[k + 1], p23 = S[k + 2] * S[k + 3];
    const double r = rcp_pos(p01 * p23);
    const double r01 = r * p23, r23 = r * p01;
    inv[k] = r01 * S[k + 1]; inv[k + 1] = r01 * S[k];
    inv[k + 2] = r23 * S[k + 3]; inv[k + 3] = r23 * S[k + 2];
  }
  constexpr int rem = K % 4, k = K - rem;
  if (rem == 1) {
    inv[k] = rcp_pos(S[k]);
  } else if (rem == 2) {
    const double r = rcp_pos(S[k] * S[k + 1]);
    inv[k] = r * S[k + 1]; inv[k + 1] = r * S[k];
  } else if (rem == 3) {
    const double p01 = S[k] * S[k + 1];
    const double r = rcp_pos(p01 * S[k + 2]);
    const double r01 = r * S[k + 2];
    inv[k] = r01 * S[k + 1]; inv[k + 1] = r01 * S[k]; inv[k + 2] = r * p01;
  }
}

// One pass over a lane's K individuals at allele odds t = f/(1-f).  With x = 1/(S/u):
//   X = sum (na + t nv) x      -> this pass adds t X to the running numerator,
//   Z = sum (dz - t nv) x      -> and g + t Z to the running (denominator - numerator),  dz = da - na.
template <int K>
NFH_DEV void pass_denominators(const double (&a0)[K], const double (&a2)[K], const double (&hh)[K],
                                                  double t, double (&S)[K]) {
#pragma unroll
  for (int k = 0; k < K; k++) S[k] = fma(fma(a2[k], t, hh[k]), t, a0[k]);
}

template <int K>
NFH_DEV void pass_sums(const double (&S)[K], const double (&na)[K], const double (&nv)[K],
                                          const double (&dz)[K], double t, double &X, double &Z) {
  double inv[K];
  reciprocals<K>(S, inv);
  double A1 = 0.0, A2 = 0.0, A3 = 0.0, B1 = 0.0, B2 = 0.0, B3 = 0.0;   // two interleaved accumulator sets
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (k & 1) { B1 = fma(na[k], inv[k], B1); B2 = fma(nv[k], inv[k], B2); B3 = fma(dz[k], inv[k], B3); }
    else       { A1 = fma(na[k], inv[k], A1); A2 = fma(nv[k], inv[k], A2); A3 = fma(dz[k], inv[k], A3); }
  }
  const double s2 = A2 + B2;
  X = fma(t, s2, A1 + B1);
  Z = fma(-t, s2, A3 + B3);
}

// est_maf starts every site at f = 0.01 (gen_func.cpp:980)
constexpr double kStartFreq = 0.01;
constexpr double kStartOdds = 0.01 / 0.99;
// A site fixed for the minor allele drives den - num to 0 (or to a rounding residue of either sign):
// the odds are capped at 1e35, i.e. f = 1 to the last bit, and S/u ~ 1e70 keeps four-way products finite.
constexpr double kMinOddsInv = 1e-35;

// state emissions from linear GL at frequency f (calc_emission with F = 0 / 1)
NFH_DEV void emissions(double L0, double L1, double L2, double f, double &e0, double &e1) {
  double omf = 1.0 - f;
  double a = omf * f;
  double u = omf * omf, v = f * f;
  e0 = fma(L0, u, fma(L1, 2.0 * a, L2 * v));
  e1 = fma(L0, u + a, L2 * (v + a));       // het prior is exp(-1e15) = 0 when F == 1
}

}  // namespace nfh

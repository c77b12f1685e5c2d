// nfh_viterbi_math.cuh - the arithmetic of the Viterbi kernels (nfh_viterbi.cu) as pure functions: the (max, x) site
// map with the reference's in-place quirk (HMM.cpp:98-125), its product, composition of back-pointer maps, the
// traceback through one chunk.  No thread indices, no shuffles, no barriers.  In a header of its own for the same
// reason as nfh_estep_math.cuh: tests/device_arith_host.cpp compiles these functions with g++ and the CPU test suite
// decodes paths with them against the reference's own arithmetic.  (The two per-site loops of viterbi_chunk_products and
// viterbi_chunk_pointers stay inline in the kernels - as functions they compiled to different code; the test harness
// restates them and a test pins the restatement to the kernel text.)
#pragma once

#include "nfh_device.cuh"

namespace nfh {

// (max, x) product and helpers; entries are non-negative
NFH_DEV M2 tropmul(const M2 &x, const M2 &y) {
  M2 r;
  r.a = fmax(x.a * y.a, x.b * y.c);
  r.b = fmax(x.a * y.b, x.b * y.d);
  r.c = fmax(x.c * y.a, x.d * y.c);
  r.d = fmax(x.c * y.b, x.d * y.d);
  return r;
}

struct SiteQ { double q00, q01, q10, q11; };

// true 0->1 transition probability (1-c) q1 = kappa q1 / (1 + kappa); = q1 at chromosome starts
NFH_DEV double trans01(double kap, double q1) { return kap * q1 * rcp_pos(1.0 + kap); }

NFH_DEV SiteQ site_q(double kap, double q0, double q1, double e0, double r) {
  const double k0 = kap * q0, k1 = kap * q1, e1 = e0 * r;
  const double t01 = trans01(kap, q1);
  SiteQ s;
  s.q00 = (1.0 + k0) * e0;
  s.q10 = k0 * e0;
  s.q01 = s.q00 * t01 * e1;
  s.q11 = fmax(s.q10 * t01, 1.0 + k1) * e1;
  return s;
}

NFH_DEV void trop_apply(M2 &m, const SiteQ &s) {
  const double a = fmax(m.a * s.q00, m.b * s.q10), b = fmax(m.a * s.q01, m.b * s.q11);
  const double c = fmax(m.c * s.q00, m.d * s.q10), d = fmax(m.c * s.q01, m.d * s.q11);
  m.a = a; m.b = b; m.c = c; m.d = d;
}

// maps {0,1} -> {0,1} as 2 bits: bit x = image of x.  compose(f, g)(x) = f(g(x)).
NFH_DEV unsigned map_compose(unsigned f, unsigned g) {
  return ((f >> (g & 1u)) & 1u) | (((f >> ((g >> 1) & 1u)) & 1u) << 1);
}

// Traceback through a thread's kChunk back-pointer pairs from the state at its last site; b[] becomes the path
// (viterbi_chunk_trace)
NFH_DEV void vit_chunk_trace(unsigned char *b, int n_valid, unsigned state) {
#pragma unroll 3
  for (int j = kChunk - 1; j >= 0; j--) {
    const unsigned bits = b[j];
    b[j] = (unsigned char) (j < n_valid ? state : 0u);
    state = (bits >> state) & 1u;
  }
}

}  // namespace nfh

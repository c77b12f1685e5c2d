// nfh_math.cuh - branch-free FP64 primitives for the hot loops.
//
// The library exp()/division carry slow-path branches that split the unrolled
// per-site / per-individual bodies into basic blocks and stop the compiler
// from interleaving independent dependency chains (ncu, profiles/r01a: FP64
// pipe 45 % busy, warps stalled on fixed-latency dependencies).  The hot
// loops only ever need
//   * kappa = e^x - 1 for x = alpha * d in [0, 138]      (transition odds, see nfh_device.cuh)
//   * 1 / x for normal positive x
// so both are written without branches and with fewer FP64 instructions.
#pragma once

// NFH_DEV marks the pure-arithmetic device functions.  Under nvcc - the only way the product is built - it is
// `__device__ __forceinline__` and nothing else changes.  tests/device_arith_host.cpp defines it (plus the four bit-cast
// intrinsics and a stand-in for the hardware reciprocal seed) BEFORE including these headers and compiles the same
// functions with g++, so that the CPU test suite checks the kernels' arithmetic against the reference's own arithmetic without a GPU.
// That build is test infrastructure: no product code path reaches it.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define NFH_DEV __device__ __forceinline__
#define NFH_DEV_TABLE __device__ const
#elif !defined(NFH_DEV)
#error "device header: build with nvcc (only tests/device_arith_host.cpp may predefine NFH_DEV)"
#endif

namespace nfh {

// 2^(j/64), j = 0..63, round-to-nearest doubles; copied to shared memory by
// every CTA that calls expm1_pos().  Kept in global memory: the copy is one
// coalesced load per thread, whereas per-lane indices into the constant bank
// would serialise in the constant cache (and so would the lookups themselves).
NFH_DEV_TABLE double kExp2Table[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138, 1.0556451783605572,
    1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841, 1.1143867425958924,
    1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, 1.241857812073484, 1.255380757024691,
    1.2690509571917332, 1.2828700160787783, 1.2968395546510096, 1.3109612115247644, 1.3252366431597413,
    1.339667524053303, 1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647, 1.4768261459394993,
    1.4929077282912648, 1.5091644275934228, 1.5255981507445384, 1.5422108254079407, 1.559004400237837,
    1.5759808451078865, 1.593142151342267, 1.6104903319492543, 1.6280274218573478, 1.645755478153965,
    1.6636765803267364, 1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989, 1.8340080864093424,
    1.8539791250833855, 1.8741676341103, 1.8945759815869656, 1.9152065613971474, 1.9360617934922943,
    1.9571441241754002, 1.978456026387951};

#if defined(__CUDACC__)
__device__ __forceinline__ void load_exp_table(double *smem_tab) {
  if (threadIdx.x < 64) smem_tab[threadIdx.x] = kExp2Table[threadIdx.x];
}
#endif

// kappa = e^x - 1 for 0 <= x <= 138 (caller guarantees the range), ~1 ulp of e^x.
// x = k ln2/64 + r, e^x = 2^(k>>6) T[k&63] (1 + p(r)), |r| <= ln2/128.
// 11 FP64 instructions + a shared-memory load; exact (= p) for x < ln2/128.
NFH_DEV double expm1_pos(double x, const double *__restrict__ tab) {
  const double kInv = 92.33248261689366;             // 64 / ln 2
  const double kMagic = 6755399441055744.0;              // 1.5 * 2^52
  const double kLn2_64_hi = 0.010830424696223417;  // ln2/64, top bits
  const double kLn2_64_lo = 2.572804622327669e-14;   // ln2/64 - hi (hi has 16 trailing zero bits: k*hi is exact)
  double kd = fma(x, kInv, kMagic);
  const int ki = __double2loint(kd);
  kd -= kMagic;
  double r = fma(kd, -kLn2_64_hi, x);
  r = fma(kd, -kLn2_64_lo, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = p * r;
  p = fma(p, r, r);                                      // e^r - 1
  const double t = tab[ki & 63];
  const double ts = __hiloint2double(__double2hiint(t) + ((ki >> 6) << 20), __double2loint(t));   // T 2^(k>>6)
  return fma(ts, p, ts - 1.0);
}

// kappa = e^x - 1 for 0 <= x < 0.0054: the polynomial of expm1_pos() alone.  For such x expm1_pos() has
// k = 0, r = x, table entry 1.0 and returns fma(1, p, 0) = p, so the two functions agree bit for bit.
NFH_DEV double expm1_small(double x) {
  double p = fma(x, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, x, 1.0 / 6.0);
  p = fma(p, x, 0.5);
  p = p * x;
  return fma(p, x, x);
}

// 1/x for normal positive x: hardware seed (~2^-23) + one cubic Newton step
// (y (1 + e + e^2), e = 1 - x y) -> error ~2^-69 before rounding, i.e. ~1 ulp.
// With refine = true one more linear step is added (the full library sequence).
// hardware reciprocal seed alone (relative error ~2^-23)
NFH_DEV double rcp_seed(double x) {
#if defined(__CUDACC__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
#else
  return nfh_host_rcp_seed(x);
#endif
}

template <bool refine = false>
NFH_DEV double rcp_pos(double x) {
#if defined(__CUDACC__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  double y = nfh_host_rcp_seed(x);
#endif
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  if (refine) {
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
}

}  // namespace nfh

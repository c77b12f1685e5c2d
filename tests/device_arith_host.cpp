// device_arith_host.cpp - TEST INFRASTRUCTURE: the arithmetic of the product's CUDA kernels, compiled for the host.
//
// The per-thread bodies of the hot kernels are pure functions in headers of ngsf-hmm_b200/csrc
//   nfh_math.cuh        expm1_pos / expm1_small / rcp_pos
//   nfh_device.cuh      factored site matrix, kappa tiers, renormalisation by exact powers of two
//   nfh_estep_math.cuh  products_chunk<TIER>() and apply_chunk<TIER>(): what one thread of estep_chunk_products /
//                       estep_chunk_apply does with its 33 sites
//   nfh_freq_math.cuh   make_coef(), pass_denominators(), pass_sums(), reciprocals(), emissions()
//   nfh_viterbi_math.cuh  site_q() / trop_apply() / tropmul() / map_compose() / vit_chunk_trace()
// marked NFH_DEV.  This file defines NFH_DEV as `static inline`, supplies the four bit-cast intrinsics and a stand-in
// for the hardware reciprocal seed, includes those headers unchanged and strings the bodies together the way the
// kernels' warps and CTAs do (sequentially instead of by shuffles).  tests/test_device_arith_cpu.py compares the
// results with the oracle, so the CPU suite checks the kernels' restatement of the reference's log-space
// arithmetic - without a GPU, and without being a code path of the product (nothing under ngsf-hmm_b200/ builds or
// loads this file; the product still fails without a CUDA device).
//
// What is NOT the product's code here, and therefore not what these tests pin: the combination of chunk products
// across a warp / CTA / individual (a plain left-to-right loop here, scans on the device - associativity makes them
// equal up to rounding), the six statements of the frequency kernels' scalar update after the lane reduction and the two
// per-site loops of the Viterbi kernels, which are restated below (tests pin them against the kernel source text).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define NFH_DEV static inline
#define NFH_DEV_TABLE static const

static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int) (uint32_t) (b >> 32); }
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int) (uint32_t) b; }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t b = ((uint64_t) (uint32_t) hi << 32) | (uint64_t) (uint32_t) lo;
  double x; std::memcpy(&x, &b, 8); return x;
}
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
using std::max;
using std::min;

// Stand-in for MUFU.RCP64H (rcp.approx.ftz.f64): the hardware looks at the high word of x only and returns a
// reciprocal good to about 2^-23 with a zero low word.  Here: 1/x of x truncated to its high word, truncated to its
// high word - two truncations of 2^-20 each, i.e. a seed several times WORSE than the hardware's, so the Newton
// steps of rcp_pos() are tested from a pessimistic start.
static inline double nfh_host_rcp_seed(double x) {
  const double xt = __hiloint2double(__double2hiint(x), 0);
  const double y = 1.0 / xt;
  return __hiloint2double(__double2hiint(y), 0);
}

#include "nfh_device.cuh"
#include "nfh_estep_math.cuh"
#include "nfh_freq_math.cuh"
#include "nfh_viterbi_math.cuh"

using namespace nfh;

extern "C" {

// ---------------------------------------------------------------------------------------------------------------
// nfh_math.cuh
// ---------------------------------------------------------------------------------------------------------------
double da_expm1_pos(double x) { return expm1_pos(x, kExp2Table); }
double da_expm1_small(double x) { return expm1_small(x); }
double da_rcp_pos(double x, int refine) { return refine ? rcp_pos<true>(x) : rcp_pos<false>(x); }
double da_rcp_seed(double x) { return rcp_seed(x); }
void da_constants(double *out) {
  out[0] = kBigX; out[1] = kFastX; out[2] = kMidX; out[3] = kEps; out[4] = kStartFreq; out[5] = kStartOdds;
  out[6] = kMinOddsInv; out[7] = (double) kChunk; out[8] = (double) kTile; out[9] = (double) kStash;
}

// ---------------------------------------------------------------------------------------------------------------
// E-step of one individual (HMM.cpp:6-60, EM.cpp:166-185) from the kernels' chunk bodies.
//   ratio[s] = e1/e0, dist[s] in Mb (+inf at chromosome starts), s = 0..S-1; loge0_sum = sum_s log e0
//   post_out[S]: clamped IBD posterior; lkl[0] forwards, lkl[1] backwards (EM.cpp:166 compares them)
// returns the status flags the kernels raise (1 NaN, 2 forward/backward mismatch)
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

template <int TIER>
static void chunk_product_tier(const double *r, const double *d, double al, double q0, double q1, M2 &m, int &e,
                               double &ls) {
  products_chunk<TIER>(r, d, kExp2Table, al, q0, q1, m, e, ls);
}
template <int TIER>
static bool chunk_apply_tier(double *r, double *d, double al, double q0, double q1, double a0, double a1, double b0,
                             double b1, double *stash) {
  return apply_chunk<TIER>(r, d, kExp2Table, al, q0, q1, a0, a1, b0, b1, stash);
}

extern "C" {

int da_estep(uint64_t S, const double *ratio, const double *dist, double F, double alpha, double loge0_sum,
             double *post_out, double *lkl, int *tiers_used) {
  const uint64_t n_tiles = (S + kTile - 1) / kTile, n_pad = n_tiles * kTile, n_chunks = n_pad / kChunk;
  // the context keeps r = 1, d = 0 at the padding sites, which makes them the identity (nfh_ctx.cu)
  std::vector<double> r(n_pad, 1.0), d(n_pad, 0.0);
  std::copy(ratio, ratio + S, r.begin());
  std::copy(dist, dist + S, d.begin());
  // per tile: largest distance (NaN poisons, +inf stays) and sum - nfh_upload_pos_dist
  std::vector<double> tmax(n_tiles), tsum(n_tiles);
  for (uint64_t t = 0; t < n_tiles; t++) {
    double mx = 0.0, sum = 0.0;
    bool nan = false;
    for (uint64_t s = t * kTile; s < std::min<uint64_t>(S, (t + 1) * kTile); s++) {
      nan |= d[s] != d[s];
      mx = d[s] > mx ? d[s] : mx;
      sum += d[s];
    }
    tmax[t] = nan ? std::nan("") : mx;
    tsum[t] = sum;
  }
  const double q0 = 1.0 - F, q1 = F;
  std::vector<M2> prod(n_chunks);
  std::vector<int> pexp(n_chunks);
  double lsum = 0.0;
  for (uint64_t t = 0; t < n_tiles; t++) {
    const int tier = kappa_tier(alpha, tmax[t]);
    if (tiers_used) tiers_used[tier]++;
    double tile_ls = 0.0;
    for (uint64_t c = t * kScanThreads; c < (t + 1) * kScanThreads; c++) {
      double ls = 0.0;
      const double *rc = r.data() + c * kChunk, *dc = d.data() + c * kChunk;
      if (tier == kTierFast) chunk_product_tier<kTierFast>(rc, dc, alpha, q0, q1, prod[c], pexp[c], ls);
      else if (tier == kTierMid) chunk_product_tier<kTierMid>(rc, dc, alpha, q0, q1, prod[c], pexp[c], ls);
      else chunk_product_tier<kTierSlow>(rc, dc, alpha, q0, q1, prod[c], pexp[c], ls);
      tile_ls += ls;
    }
    lsum += tier == kTierSlow ? tile_ls : -(alpha * tsum[t]);      // estep_chunk_products, thread 0
  }
  // carries (estep_tile_carries + chunk_carries, sequentially): forward vector entering, backward vector leaving
  std::vector<double> fa0(n_chunks), fa1(n_chunks), bb0(n_chunks), bb1(n_chunks);
  double x0 = q0, x1 = q1;
  long long ex = 0;
  for (uint64_t c = 0; c < n_chunks; c++) {
    fa0[c] = x0; fa1[c] = x1;
    const M2 &p = prod[c];
    const double y0 = fma(x0, p.a, x1 * p.c), y1 = fma(x0, p.b, x1 * p.d);
    x0 = y0; x1 = y1;
    ex += pexp[c] + renorm2_i(x0, x1);
  }
  const double base = lsum + loge0_sum;
  const double lf = std::log(x0 + x1) + (double) ex * kLn2 + base;
  double b0 = 1.0, b1 = 1.0;
  long long eb = 0;
  for (uint64_t c = n_chunks; c-- > 0;) {
    bb0[c] = b0; bb1[c] = b1;
    const M2 &p = prod[c];
    const double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
    b0 = y0; b1 = y1;
    eb += pexp[c] + renorm2_i(b0, b1);
  }
  const double lb = std::log(fma(q0, b0, q1 * b1)) + (double) eb * kLn2 + base;
  lkl[0] = lf; lkl[1] = lb;
  int status = 0;
  if (lf != lf || lb != lb) status |= kFlagNaN;
  else if (std::fabs(lf - lb) > 1e-3) status |= kFlagFwBw;

  // posterior sweep: apply_chunk overwrites its 33 ratios with the posteriors and its distances with kappa
  std::vector<double> stash((size_t) kStash * kScanThreads);
  bool bad = false;
  for (uint64_t c = 0; c < n_chunks; c++) {
    const int tier = kappa_tier(alpha, tmax[c / kScanThreads]);
    double *rc = r.data() + c * kChunk, *dc = d.data() + c * kChunk;
    double *st = stash.data() + (c % kScanThreads);               // [value][thread] as in ApplySmem
    if (tier == kTierFast) bad |= chunk_apply_tier<kTierFast>(rc, dc, alpha, q0, q1, fa0[c], fa1[c], bb0[c], bb1[c], st);
    else if (tier == kTierMid) bad |= chunk_apply_tier<kTierMid>(rc, dc, alpha, q0, q1, fa0[c], fa1[c], bb0[c], bb1[c], st);
    else bad |= chunk_apply_tier<kTierSlow>(rc, dc, alpha, q0, q1, fa0[c], fa1[c], bb0[c], bb1[c], st);
  }
  if (bad) status |= kFlagNaN;
  std::copy(r.begin(), r.begin() + S, post_out);
  return status;
}

// ---------------------------------------------------------------------------------------------------------------
// lkl() (EM.cpp:449-464) the way lkl_chunk_run (nfh_lkl.cu) walks a chunk for one point: apply_site with kappa q,
// renormalised every kBody sites, chunk products chained in order.  Returns -log-likelihood.
// ---------------------------------------------------------------------------------------------------------------
double da_neg_lkl(uint64_t S, const double *ratio, const double *dist, double F, double alpha, double loge0_sum) {
  if (F != F || alpha != alpha || std::isinf(F) || std::isinf(alpha)) return -1e15;   // EM.cpp:454-456, as nfh_ctx.cu
  const uint64_t n_chunks = (S + kChunk - 1) / kChunk;
  const double q0 = 1.0 - F, q1 = F;
  double x0 = q0, x1 = q1, ls = 0.0;
  long long ex = 0;
  constexpr int kBody = 6;                                        // the slow tier's window; the others use 8
  for (uint64_t c = 0; c < n_chunks; c++) {
    M2 m = identity2();
    int e = 0;
    for (int j = 0; j < kChunk; j++) {
      const uint64_t s = c * kChunk + j;
      const double rj = s < S ? ratio[s] : 1.0, dj = s < S ? dist[s] : 0.0;
      const double k = tier_kappa<kTierSlow>(alpha * dj, kExp2Table, ls);
      apply_site(m, k * q0, k * q1, rj);
      if (j % kBody == kBody - 1) e += renorm_i(m);
    }
    e += renorm_i(m);
    const double y0 = fma(x0, m.a, x1 * m.c), y1 = fma(x0, m.b, x1 * m.d);
    x0 = y0; x1 = y1;
    ex += e + renorm2_i(x0, x1);
  }
  return -(std::log(x0 + x1) + (double) ex * kLn2 + ls + loge0_sum);
}

// ---------------------------------------------------------------------------------------------------------------
// One site of the frequency EM + emission refresh (gen_func.cpp:974-1009, HMM.cpp:144-154) the way
// freq_emission_warp<G, K> runs it: G lanes share the site, lane g keeps individuals g, g + G, g + 2G, ... (K of
// them) as coefficient arrays; a pass = pass_sums() per lane, butterfly sum over the lanes, the scalar update.
//   L0/L1/L2[n_ind] linear GL, post[n_ind] IBD posterior (NULL: F = 0, parse_args.cpp:316-318)
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

template <int G, int K>
static int freq_site(uint64_t n_ind, const double *L0, const double *L1, const double *L2, const double *post,
                     int update_freq, double *freq_io, double *ratio_out, double *e0_out) {
  double a0[G][K], a2[G][K], hh[G][K], na[G][K], nv[G][K], dz[G][K], S[G][K];
  double g_sum[G];
  for (int g = 0; g < G; g++) {
    g_sum[g] = 0.0;
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) g + (uint64_t) G * k;
      const IndCoef c = i < n_ind ? make_coef(L0[i], L1[i], L2[i], post ? post[i] : 0.0) : null_coef();
      a0[g][k] = c.a0; a2[g][k] = c.a2; hh[g][k] = c.h; na[g][k] = c.na; nv[g][k] = c.nv; dz[g][k] = c.da - c.na;
      g_sum[g] += c.g;
    }
  }
  auto butterfly = [](double (&v)[G]) {               // __shfl_xor_sync levels m = 1, 2, 4, ...
    for (int m = 1; m < G; m <<= 1) {
      double t[G];
      for (int g = 0; g < G; g++) t[g] = v[g] + v[g ^ m];
      for (int g = 0; g < G; g++) v[g] = t[g];
    }
  };
  double freq = update_freq ? kStartFreq : *freq_io;
  int site_passes = 0;
  if (update_freq) {
    butterfly(g_sum);
    const double gs = g_sum[0];
    // ---- the kernels' scalar state and update, restated (pinned against nfh_freq.cu by the test suite)
    double num = 0.0, dmn_next = gs;
    double odds = kStartOdds, prev = kStartFreq;
    bool active = true;
    int passes = 0;
    for (int g = 0; g < G; g++) pass_denominators<K>(a0[g], a2[g], hh[g], odds, S[g]);
    do {
      double X[G], Z[G];
      for (int g = 0; g < G; g++) pass_sums<K>(S[g], na[g], nv[g], dz[g], odds, X[g], Z[g]);
      butterfly(X);
      butterfly(Z);
      num = fma(odds, X[0], num);
      const double dmn = fmax(fma(odds, Z[0], dmn_next), num * kMinOddsInv);
      odds = num * rcp_pos(dmn);
      for (int g = 0; g < G; g++) pass_denominators<K>(a0[g], a2[g], hh[g], odds, S[g]);
      dmn_next = dmn + gs;
      const double now = num * rcp_pos<true>(num + dmn);
      passes++;
      freq = active ? now : freq;
      site_passes = active ? passes : site_passes;
      active = active && (fabs(prev - now) > kEps) && (passes <= 100);
      prev = now;
    } while (active);
    *freq_io = freq;
  }
  for (uint64_t i = 0; i < n_ind; i++) {
    const int g = (int) (i % G), k = (int) (i / G);
    double e0, e1;
    emissions(a0[g][k], L1[i], a2[g][k], freq, e0, e1);            // a0 = L0, a2 = L2 (make_coef)
    ratio_out[i] = e1 * rcp_pos<true>(e0);
    e0_out[i] = e0;
  }
  return site_passes;
}

extern "C" {

// shape: 0 = <8,13> (configs[1], 100 individuals), 1 = <4,16>, 2 = <16,8> (125 individuals), 3 = <32,16>, 4 = <1,16>
int da_freq_site(int shape, uint64_t n_ind, const double *L0, const double *L1, const double *L2, const double *post,
                 int update_freq, double *freq_io, double *ratio_out, double *e0_out) {
  switch (shape) {
    case 0: return n_ind <= 8 * 13 ? freq_site<8, 13>(n_ind, L0, L1, L2, post, update_freq, freq_io, ratio_out, e0_out) : -1;
    case 1: return n_ind <= 4 * 16 ? freq_site<4, 16>(n_ind, L0, L1, L2, post, update_freq, freq_io, ratio_out, e0_out) : -1;
    case 2: return n_ind <= 16 * 8 ? freq_site<16, 8>(n_ind, L0, L1, L2, post, update_freq, freq_io, ratio_out, e0_out) : -1;
    case 3: return n_ind <= 32 * 16 ? freq_site<32, 16>(n_ind, L0, L1, L2, post, update_freq, freq_io, ratio_out, e0_out) : -1;
    case 4: return n_ind <= 16 ? freq_site<1, 16>(n_ind, L0, L1, L2, post, update_freq, freq_io, ratio_out, e0_out) : -1;
    default: return -1;
  }
}

// The streaming kernel's form of a pass (freq_emission_stream: u / v / a, accumulate(), num / den, individuals in
// index order - the reference's own summation order).
int da_freq_site_stream(uint64_t n_ind, const double *L0, const double *L1, const double *L2, const double *post,
                        double *freq_out) {
  double freq = 0.01, num = 0.0, den = 0.0, before;
  int passes = 0;
  do {
    before = freq;
    const double omf = 1.0 - freq;
    const double u = omf * omf, v = freq * freq, a = omf * freq;
    double A1 = 0.0, A2 = 0.0, A3 = 0.0, gs = 0.0;
    for (uint64_t i = 0; i < n_ind; i++) {
      IndCoef k = make_coef(L0[i], L1[i], L2[i], post ? post[i] : 0.0);
      accumulate(k, u, v, a, A1, A2, A3);
      gs += k.g;
    }
    num += fma(a, A1, v * A2); den += fma(a, A3, gs);
    freq = num / den;
  } while (fabs(before - freq) > kEps && passes++ < 100);
  *freq_out = freq;
  return min(passes + 1, 101);
}

// ---------------------------------------------------------------------------------------------------------------
// viterbi() (HMM.cpp:98-125) of one individual the way nfh_viterbi.cu decodes it: (max, x) chunk products -> scores
// entering every chunk -> back-pointer pairs and composed maps per chunk -> state at every chunk end -> traceback.
//   ratio = e1/e0, e0 linear, dist in Mb; path_out[S] in {0, 1}
// The two per-site loops are the kernels' text (viterbi_chunk_products, viterbi_chunk_pointers), restated.
// ---------------------------------------------------------------------------------------------------------------
void da_viterbi(uint64_t S, const double *ratio, const double *e0_lin, const double *dist, double F, double al,
                unsigned char *path_out) {
  const uint64_t n_chunks = (S + kChunk - 1) / kChunk, n_pad = n_chunks * kChunk;
  std::vector<double> rr(n_pad, 1.0), ee(n_pad, 1.0), dd(n_pad, 0.0);
  std::copy(ratio, ratio + S, rr.begin());
  std::copy(e0_lin, e0_lin + S, ee.begin());
  std::copy(dist, dist + S, dd.begin());
  const double *tab = kExp2Table;
  const double q0 = 1.0 - F, q1 = F;
  std::vector<M2> prod(n_chunks);
  for (uint64_t c = 0; c < n_chunks; c++) {
    const double *r = rr.data() + c * kChunk, *e0 = ee.data() + c * kChunk, *d = dd.data() + c * kChunk;
    const int n_valid = (int) std::min<uint64_t>(kChunk, S - c * kChunk);
    // ---- viterbi_chunk_products
    M2 m = identity2();
    for (int j = 0; j < kChunk; j++) {
      if (j < n_valid) {
        const double kap = site_kappa(al * d[j], tab);
        trop_apply(m, site_q(kap, q0, q1, e0[j], r[j]));
      }
      if (j % 6 == 5) renorm(m);
    }
    renorm(m);
    prod[c] = m;
  }
  std::vector<unsigned char> bytes(n_pad), maps(n_chunks);
  double s0 = q0, s1 = q1;                                          // Vi_prob = q, linear (viterbi_tile_scores)
  for (uint64_t c = 0; c < n_chunks; c++) {
    const double *r = rr.data() + c * kChunk, *e0 = ee.data() + c * kChunk, *d = dd.data() + c * kChunk;
    const int n_valid = (int) std::min<uint64_t>(kChunk, S - c * kChunk);
    unsigned char *bp = bytes.data() + c * kChunk;
    double v0 = s0, v1 = s1;
    // ---- viterbi_chunk_pointers
    unsigned chunk_map = 2u;
    for (int j = 0; j < kChunk; j++) {
      unsigned bits = 2u;
      if (j < n_valid) {
        const double kap = site_kappa(al * d[j], tab);
        const double k0 = kap * q0, k1 = kap * q1, e1 = e0[j] * r[j];
        double from0 = v0 * (1.0 + k0), from1 = v1 * k0;
        const unsigned bp0 = from1 > from0;
        const double n0 = (bp0 ? from1 : from0) * e0[j];
        from0 = n0 * trans01(kap, q1); from1 = v1 * (1.0 + k1);
        const unsigned bp1 = from1 > from0;
        const double n1 = (bp1 ? from1 : from0) * e1;
        v0 = n0; v1 = n1;
        if (j % 4 == 3) renorm2(v0, v1);
        bits = bp0 | (bp1 << 1);
      }
      bp[j] = (unsigned char) bits;
      chunk_map = map_compose(chunk_map, bits);
    }
    maps[c] = (unsigned char) chunk_map;
    // scores entering the next chunk come from the PRODUCTS, as on the device (not from v0 / v1 above)
    const M2 &p = prod[c];
    const double y0 = fmax(s0 * p.a, s1 * p.c), y1 = fmax(s0 * p.b, s1 * p.d);
    s0 = y0; s1 = y1;
    renorm2(s0, s1);
  }
  unsigned state = s1 > s0 ? 1 : 0;                                 // array_max_pos: first maximum
  for (uint64_t c = n_chunks; c-- > 0;) {
    const int n_valid = (int) std::min<uint64_t>(kChunk, S - c * kChunk);
    const unsigned end = state;                                     // state at the last site of chunk c
    state = (maps[c] >> state) & 1u;
    vit_chunk_trace(bytes.data() + c * kChunk, n_valid, end);
  }
  std::copy(bytes.begin(), bytes.begin() + S, path_out);
}

}  // extern "C"

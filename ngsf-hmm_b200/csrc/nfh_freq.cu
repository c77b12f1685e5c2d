// nfh_freq.cu - per-site allele-frequency EM fused with the emission refresh.
//
// Replaces the serial site loop of iter_EM (EM.cpp:224-271):
//   est_maf()        shared/gen_func.cpp:974-1009  (with calc_HWE :938-957, post_prob :920-932)
//   calc_emission()  shared/HMM.cpp:144-154
//
// est_maf is a fixed point over ALL individuals of one site: up to 101 passes,
// each pass adds every individual's expected allele counts to running sums
// that are never reset, and the frequency is the ratio of the running sums
// (SURVEY.md finding 2).  It is the dominant cost of the reference.  Here one
// site is shared by a group of lanes; every lane keeps the pass-invariant
// coefficients of its individuals IN REGISTERS for all passes, so GL and
// posterior are read from HBM once per EM iteration and each pass is pure
// FP64 arithmetic plus a shuffle reduction.  The refreshed emission ratio
// e1/e0 is produced from the same registers when the site has converged.
//
// Linear-space form of one individual's contribution at frequency f, with
// u = (1-f)^2, v = f^2, a = f(1-f), GL (L0,L1,L2) and IBD posterior F:
//   w0 = L0 (u + a F)   w1 = L1 2a(1-F)   w2 = L2 (v + a F)     (HWE prior x GL)
//   num += (w1 + w2 (2-F)) / (w0+w1+w2)
//   den += (2 w1 + (w0+w2)(2-F)) / (w0+w1+w2)
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "nfh_device.cuh"
#include "nfh_freq_math.cuh"
#include "nfh_kernels.h"

namespace nfh {

constexpr int kFreqThreads = 128;
constexpr int kMaxK = 16;             // most individuals one lane keeps in registers

// Resident CTAs per SM the register budget of K individuals per lane allows
// (12 K registers of coefficients + 4 K of per-pass temporaries + ~40).
__host__ __device__ constexpr int freq_occupancy(int K) { return K <= 4 ? 4 : K <= 8 ? 3 : 2; }

// Where the refreshed emission of (individual i, local site) goes: the local window, or - in direct
// mode - the recursion-side window of the rank that owns individual i, source block = this rank.
__device__ __forceinline__ double *emis_slot(const FreqArgs &A, uint64_t i, uint64_t site) {
  if (!A.emis_peers.direct) return A.emis + (size_t) i * A.site_block + site;
  const uint64_t owner = i / A.emis_peers.n_loc, il = i - owner * A.emis_peers.n_loc;
  return A.emis_peers.base[owner] + ((uint64_t) A.emis_peers.rank * A.emis_peers.n_loc + il) * A.site_block + site;
}

// Shared-memory image of one site tile of freq_emission_warp.  Each of the four planes (GL0, GL1, GL2,
// posterior; [n_ind_pad][site_block] in HBM) is fetched as 2-D tensor boxes of (<= 256 individuals) x
// (<= 16 sites) by cp.async.bulk.tensor - a handful of copies per tile issued by one thread, instead of
// one bulk copy per individual row (ncu r01f: issuing 400 row copies per tile cost 8 % of the kernel).
// The boxes are swizzled (32/64/128-byte pattern = the row length) so that the lanes of a half-warp,
// which read different individuals at 1-4 neighbouring sites, spread over the banks.
template <int G> struct FreqTile {
  static constexpr int kSitesPerWarp = 32 / G;
  static constexpr int kSitesPerCta = kSitesPerWarp * (kFreqThreads / 32);
  static constexpr int kInner = kSitesPerCta < 16 ? kSitesPerCta : 16;   // sites per box row
  static constexpr int kInnerBytes = kInner * 8;                         // 32, 64 or 128 = swizzle span
  static constexpr int kBoxesX = kSitesPerCta / kInner;
  static constexpr int kMaxBoxRows = 256;
  static constexpr unsigned kSwizzleMask = kInnerBytes / 16 - 1;
  static constexpr unsigned kAlign = 8 * kInnerBytes;                    // period of the swizzle pattern in bytes
  __host__ __device__ static unsigned boxes_y(uint64_t n_ind) { return n_ind > kMaxBoxRows ? 2u : 1u; }   // warp shapes hold n_ind <= 512
  __host__ __device__ static unsigned box_rows(uint64_t n_ind) { return (unsigned) ((n_ind + boxes_y(n_ind) - 1) / boxes_y(n_ind)); }
  __host__ __device__ static size_t region_bytes(uint64_t n_ind) { return ((size_t) box_rows(n_ind) * kInnerBytes + kAlign - 1) / kAlign * kAlign; }
  __host__ __device__ static size_t tile_bytes(uint64_t n_ind) { return 4 * kBoxesX * boxes_y(n_ind) * region_bytes(n_ind); }
  // byte offset of (plane, individual i, site s of the CTA tile) inside a tile buffer (kAlign-aligned);
  // rows = box_rows, by = boxes_y, region = region_bytes of the tile (hoisted by the caller)
  __device__ static unsigned offset(unsigned plane, unsigned i, unsigned s, unsigned rows, unsigned by, unsigned region) {
    const unsigned y = i >= rows ? 1u : 0u;          // n_ind <= 512: at most two boxes of rows
    unsigned off = (i - y * rows) * kInnerBytes + (s % kInner) * 8;
    off ^= ((off >> 7) & kSwizzleMask) << 4;
    return ((plane * kBoxesX + s / kInner) * by + y) * region + off;
  }
  // per warp and individual: running product of e0 (mantissa in [1,2) as double + exponent as int)
  __host__ __device__ static size_t acc_doubles(uint64_t n_ind_pad) { return (((size_t) (kFreqThreads / 32) * n_ind_pad * 3 / 2 + 15) / 16) * 16; }
};

// G lanes share one site (G divides 32); each lane keeps K individuals.
//
// PREFETCH: the CTA's next site tile (GL x3 + posterior rows of all individuals, 16-32 sites wide)
// is fetched into shared memory by TMA tensor copies while the current tile runs its ~101 passes, so
// the set-up of a tile reads shared memory instead of waiting on HBM (ncu r01c: long-scoreboard
// stalls were 18 % of the kernel).  Two buffers, one mbarrier each.
//
// The pass loop is one dependency chain per site (sums -> lane reduction -> division -> next odds),
// and a scheduler holds only two such warps, so the chain is kept as short as it can be: no
// predication on it (converged sites simply keep iterating, their frequency was latched when they
// stopped), the running (den - num) is accumulated directly so the next odds need one FMA and one
// reciprocal after the reduction, and the stop test + vote hang off the side of the chain.
// MODE 0: no prefetch; 1: prefetch, accumulators in shared memory; 2: prefetch, accumulators in global scratch
template <int G, int K, int MODE, int OCC = freq_occupancy(K)>
__global__ void __launch_bounds__(kFreqThreads, OCC)
freq_emission_warp(const __grid_constant__ FreqArgs A, unsigned n_site_tiles) {
  constexpr bool PREFETCH = MODE != 0, ACC_GLOBAL = MODE == 2;
  using Tile = FreqTile<G>;
  constexpr int kSitesPerWarp = Tile::kSitesPerWarp;
  constexpr int kSitesPerCta = Tile::kSitesPerCta;
  constexpr int kWarps = kFreqThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane & (G - 1), sub = lane / G;
  extern __shared__ __align__(128) double freq_smem[];
  // accumulators [warps][n_ind_pad]: shared memory, or - when the tile buffers need all of it - this CTA's
  // slice of a global scratch array (touched once per tile, latency hidden by the next tile's passes)
  double *mant_acc = ACC_GLOBAL ? A.acc_scratch + (size_t) blockIdx.x * Tile::acc_doubles(A.n_ind_pad) : freq_smem;
  int *expo_acc = reinterpret_cast<int *>(mant_acc + (size_t) kWarps * A.n_ind_pad);
  for (unsigned i = threadIdx.x; i < kWarps * A.n_ind_pad; i += kFreqThreads) { mant_acc[i] = 1.0; expo_acc[i] = 0; }
  double *my_mant = mant_acc + (size_t) warp * A.n_ind_pad;
  int *my_expo = expo_acc + (size_t) warp * A.n_ind_pad;

  // ---- prefetch machinery: two tile buffers (aligned to the swizzle period), one mbarrier each
  const unsigned n_planes = A.post ? 4u : 3u;
  const size_t buf_bytes = Tile::tile_bytes(A.n_ind);
  const unsigned box_rows = Tile::box_rows(A.n_ind), boxes_y = Tile::boxes_y(A.n_ind);
  const unsigned region_bytes = (unsigned) Tile::region_bytes(A.n_ind);
  unsigned char *bufs = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(freq_smem + (ACC_GLOBAL ? 0 : Tile::acc_doubles(A.n_ind_pad))) + Tile::kAlign - 1) &
      ~(uintptr_t) (Tile::kAlign - 1));
  __shared__ alignas(8) uint64_t bars[2];
  auto issue_tile = [&](unsigned t, int b) {      // thread 0 only
    mbar_arrive_expect_tx(&bars[b], n_planes * Tile::kBoxesX * boxes_y * box_rows * Tile::kInnerBytes);
    const int first_site = (int) (t * kSitesPerCta);
    for (unsigned p = 0; p < n_planes; p++)
      for (unsigned bx = 0; bx < (unsigned) Tile::kBoxesX; bx++)
        for (unsigned y = 0; y < boxes_y; y++)
          tma_load_2d(bufs + (size_t) b * buf_bytes + ((p * Tile::kBoxesX + bx) * boxes_y + y) * region_bytes,
                      &A.maps[p], first_site + (int) bx * Tile::kInner, (int) (y * box_rows), &bars[b]);
  };
  if (PREFETCH) {
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  }
  __syncthreads();
  if (PREFETCH && threadIdx.x == 0 && blockIdx.x < n_site_tiles) issue_tile(blockIdx.x, 0);

  unsigned round = 0;
  unsigned my_passes = 0;                       // passes of the sites this lane reports (grp == 0)
  for (unsigned tile = blockIdx.x; tile < n_site_tiles; tile += gridDim.x, round++) {
    const int site_in_cta = warp * kSitesPerWarp + sub;
    const uint64_t site = (uint64_t) tile * kSitesPerCta + site_in_cta;
    const bool site_ok = site < A.sites_owned;
    const uint64_t sl = site_ok ? site : 0;
    const unsigned char *tile_buf = bufs + (size_t) (round & 1) * buf_bytes;
    auto staged = [&](unsigned plane, uint64_t i) {
      return *reinterpret_cast<const double *>(
          tile_buf + Tile::offset(plane, (unsigned) i, (unsigned) site_in_cta, box_rows, boxes_y, region_bytes));
    };
    if (PREFETCH) {
      // the other buffer was last read before the __syncthreads that ended the previous tile
      if (threadIdx.x == 0 && tile + gridDim.x < n_site_tiles) issue_tile(tile + gridDim.x, (round & 1) ^ 1);
      mbar_wait(&bars[round & 1], (round >> 1) & 1);
    }

    // all loads of the tile are issued before the first coefficient is formed; the coefficient
    // arrays double as staging for L0, L2, L1, F
    double a0[K], a2[K], hh[K], na[K], nv[K], dz[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      const uint64_t il = i < A.n_ind ? i : 0;
      if (PREFETCH) {
        a0[k] = staged(0, il); hh[k] = staged(1, il); a2[k] = staged(2, il);
        na[k] = A.post ? staged(3, il) : 0.0;
      } else {
        const size_t at = (size_t) il * A.site_block + sl;
        a0[k] = A.gl0[at]; hh[k] = A.gl1[at]; a2[k] = A.gl2[at];
        na[k] = A.post ? A.post[at] : 0.0;
      }
    }
    double g_sum = 0.0;                       // sum over individuals of (2 - F): constant part of den per pass
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      double L0 = a0[k], L1 = hh[k], L2 = a2[k], F = na[k];
      if (PREFETCH && !site_ok) { L0 = 1.0 / 3; L1 = 1.0 / 3; L2 = 1.0 / 3; F = 0.0; }   // padding sites: harmless values
      const IndCoef c = i < A.n_ind ? make_coef(L0, L1, L2, F) : null_coef();
      a0[k] = c.a0; a2[k] = c.a2; hh[k] = c.h; na[k] = c.na; nv[k] = c.nv; dz[k] = c.da - c.na;
      g_sum += c.g;
    }

    double freq = A.update_freq ? kStartFreq : A.freq[sl];
    if (A.update_freq) {
#pragma unroll
      for (int m = 1; m < G; m <<= 1) g_sum += __shfl_xor_sync(kFull, g_sum, m);
      double num = 0.0, dmn_next = g_sum;     // running numerator; running (den - num) + this pass's sum of g
      double odds = kStartOdds, prev = kStartFreq;
      bool active = site_ok;
      int passes = 0, site_passes = 0;
      double S[K];
      pass_denominators<K>(a0, a2, hh, odds, S);
      do {
        double X, Z;
        pass_sums<K>(S, na, nv, dz, odds, X, Z);
#pragma unroll
        for (int m = 1; m < G; m <<= 1) {
          X += __shfl_xor_sync(kFull, X, m);
          Z += __shfl_xor_sync(kFull, Z, m);
        }
        num = fma(odds, X, num);
        const double dmn = fmax(fma(odds, Z, dmn_next), num * kMinOddsInv);   // f -> 1: odds stay finite
        odds = num * rcp_pos(dmn);
        // the next pass starts here, inside the same basic block, so that the compiler schedules the
        // chain num -> odds -> S ahead of the stop test (one wasted set of denominators at the exit)
        pass_denominators<K>(a0, a2, hh, odds, S);
        dmn_next = dmn + g_sum;
        const double now = num * rcp_pos<true>(num + dmn);   // frequency after this pass: off the chain
        passes++;
        freq = active ? now : freq;
        site_passes = active ? passes : site_passes;
        // do { ... } while (|before - freq| > EPSILON && iters++ < 100)   gen_func.cpp:1006
        active = active && (fabs(prev - now) > kEps) && (passes <= 100);
        prev = now;
      } while (__any_sync(kFull, active));
      if (site_ok && grp == 0) { A.freq[site] = freq; my_passes += site_passes; }
    }

    // emission refresh from the same site (L1 is re-read: the coefficients fold it with F)
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      double pe0 = 1.0;
      if (i < A.n_ind && site_ok) {
        const size_t at = (size_t) i * A.site_block + site;
        const double L1 = PREFETCH ? staged(1, i) : A.gl1[at];
        double e0, e1;
        emissions(a0[k], L1, a2[k], freq, e0, e1);
        *emis_slot(A, i, site) = e1 * rcp_pos<true>(e0);
        if (A.e0) A.e0[at] = e0;
        pe0 = e0;
      }
      // sum of log e0 over sites = log of the running product: multiply the warp's sites (lanes with
      // equal grp, fixed order), then fold into the warp's (mantissa, exponent) accumulator
#pragma unroll
      for (int m = G; m < 32; m <<= 1) pe0 *= __shfl_xor_sync(kFull, pe0, m);
      if (sub == 0 && i < A.n_ind_pad) {
        int e;
        my_mant[i] = split_exponent(my_mant[i] * pe0, e);
        my_expo[i] += e;
      }
    }
    if (PREFETCH) __syncthreads();            // the buffer is refilled two tiles from now
  }
  __syncthreads();
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) my_passes += __shfl_xor_sync(kFull, my_passes, m);
  if (lane == 0 && my_passes) atomicAdd(A.pass_total, (unsigned long long) my_passes);
  for (unsigned i = threadIdx.x; i < A.n_ind_pad; i += kFreqThreads) {
    double m = 1.0;
    int e = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) { m *= mant_acc[(size_t) w * A.n_ind_pad + i]; e += expo_acc[(size_t) w * A.n_ind_pad + i]; }
    A.loge0_part[(size_t) blockIdx.x * A.n_ind_pad + i] = fma((double) e, 0.6931471805599453, log(m));
  }
}

// ---------------------------------------------------------------------------
// Hybrid variant (used for 512 < n_ind <= 832): G lanes share a site and every lane keeps
// K = kHybridRegK + KS individuals - the first kHybridRegK in registers, the other KS in
// shared memory (its own column, read once per pass, conflict-free).  Keeping the lane
// group small is what matters at large n_ind: the per-pass tail (lane reduction, division,
// stop test) is paid per warp, so a warp should carry as many individual-passes as it can;
// spreading the individuals over more lanes (G = 32 registers-only, or teams of warps)
// leaves the FP64 pipe half idle behind 5 shuffle levels and a barrier per pass.  The price
// is 48 bytes of shared-memory reads per stored individual and pass, about what an SM can
// deliver beside the FP64 work.
// ---------------------------------------------------------------------------
constexpr int kHybridRegK = 12;      // 9 gave the same time: the compiler fills 255 registers either way
constexpr int kHybridMaxKS = 14;     // with kHybridRegK register individuals: n_ind <= 832
// Beyond that (the frequency side of 8 ranks x 125 individuals sees 1,000): 13 or 14 register individuals and up to
// 18 in shared memory (110,592 bytes per CTA, still two CTAs per SM) reach n_ind = 1,024.  The team kernel that took
// these sizes before gives every CTA ONE site, so that a CTA uses 8 of the 32 bytes of every sector it fetches and
// slows down with the row length (1.71 ps per individual-pass at 50,000 sites per row, 2.21 at 1.25M, 2.76 with peer
// stores on 8 GPUs); here the four warps of a CTA work on four neighbouring sites.
constexpr int kHybridMaxKSWide = 18;

// partial sums of the KS shared-memory individuals of this lane (same algebra as pass_sums)
template <int KS>
__device__ __forceinline__ void hybrid_sums(const double *__restrict__ coef, double t, double &A1, double &A2, double &A3) {
  constexpr int kStride = kFreqThreads;          // doubles between consecutive coefficient planes
#pragma unroll
  for (int j0 = 0; j0 < KS; j0 += 4) {
    constexpr int kFull4 = 4;
    const int n = KS - j0 < kFull4 ? KS - j0 : kFull4;
    double S[4], inv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j < n) {
        const double *c = coef + (size_t) (j0 + j) * 6 * kStride;
        S[j] = fma(fma(c[kStride], t, c[2 * kStride]), t, c[0]);      // planes: a0, a2, h, na, nv, dz
      } else {
        S[j] = 1.0;
      }
    }
    reciprocals<4>(S, inv);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j < n) {
        const double *c = coef + (size_t) (j0 + j) * 6 * kStride;
        A1 = fma(c[3 * kStride], inv[j], A1);
        A2 = fma(c[4 * kStride], inv[j], A2);
        A3 = fma(c[5 * kStride], inv[j], A3);
      }
    }
  }
}

template <int G, int KS, int KR>
__global__ void __launch_bounds__(kFreqThreads, 2)
freq_emission_hybrid(FreqArgs A, unsigned n_site_tiles) {
  constexpr int K = KR + KS;
  constexpr int kSitesPerWarp = 32 / G;
  constexpr int kSitesPerCta = kSitesPerWarp * (kFreqThreads / 32);
  constexpr int kWarps = kFreqThreads / 32;
  constexpr int kStride = kFreqThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane & (G - 1), sub = lane / G;
  extern __shared__ __align__(16) double hybrid_smem[];        // [KS * 6][kFreqThreads]
  double *coef = hybrid_smem + threadIdx.x;
  // log e0 accumulators of this CTA in global scratch (see freq_emission_warp, MODE 2)
  const size_t acc_doubles = (((size_t) kWarps * A.n_ind_pad * 3 / 2 + 15) / 16) * 16;
  double *mant_acc = A.acc_scratch + (size_t) blockIdx.x * acc_doubles;
  int *expo_acc = reinterpret_cast<int *>(mant_acc + (size_t) kWarps * A.n_ind_pad);
  for (unsigned i = threadIdx.x; i < kWarps * A.n_ind_pad; i += kFreqThreads) { mant_acc[i] = 1.0; expo_acc[i] = 0; }
  __syncthreads();
  double *my_mant = mant_acc + (size_t) warp * A.n_ind_pad;
  int *my_expo = expo_acc + (size_t) warp * A.n_ind_pad;
  unsigned my_passes = 0;

  for (unsigned tile = blockIdx.x; tile < n_site_tiles; tile += gridDim.x) {
    const uint64_t site = (uint64_t) tile * kSitesPerCta + warp * kSitesPerWarp + sub;
    const bool site_ok = site < A.sites_owned;
    const uint64_t sl = site_ok ? site : 0;

    // ---- set-up: loads first, coefficients second; register part, then the shared-memory part
    double a0[KR], a2[KR], hh[KR], na[KR], nv[KR], dz[KR];
#pragma unroll
    for (int k = 0; k < KR; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      const size_t at = (size_t) (i < A.n_ind ? i : 0) * A.site_block + sl;
      a0[k] = A.gl0[at]; hh[k] = A.gl1[at]; a2[k] = A.gl2[at];
      na[k] = A.post ? A.post[at] : 0.0;
    }
    double g_sum = 0.0;
#pragma unroll
    for (int k = 0; k < KR; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      const IndCoef c = i < A.n_ind ? make_coef(a0[k], hh[k], a2[k], na[k]) : null_coef();
      a0[k] = c.a0; a2[k] = c.a2; hh[k] = c.h; na[k] = c.na; nv[k] = c.nv; dz[k] = c.da - c.na;
      g_sum += c.g;
    }
#pragma unroll
    for (int j = 0; j < KS; j++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * (KR + j);
      IndCoef c = null_coef();
      if (i < A.n_ind) {
        const size_t at = (size_t) i * A.site_block + sl;
        c = make_coef(A.gl0[at], A.gl1[at], A.gl2[at], A.post ? A.post[at] : 0.0);
      }
      double *dst = coef + (size_t) j * 6 * kStride;
      dst[0] = c.a0; dst[kStride] = c.a2; dst[2 * kStride] = c.h;
      dst[3 * kStride] = c.na; dst[4 * kStride] = c.nv; dst[5 * kStride] = c.da - c.na;
      g_sum += c.g;
    }

    double freq = A.update_freq ? kStartFreq : A.freq[sl];
    if (A.update_freq) {
#pragma unroll
      for (int m = 1; m < G; m <<= 1) g_sum += __shfl_xor_sync(kFull, g_sum, m);
      double num = 0.0, dmn_next = g_sum;
      double odds = kStartOdds, prev = kStartFreq;
      bool active = site_ok;
      int passes = 0, site_passes = 0;
      double S[KR];
      pass_denominators<KR>(a0, a2, hh, odds, S);
      do {
        // register individuals
        double inv[KR];
        reciprocals<KR>(S, inv);
        double A1 = 0.0, A2 = 0.0, A3 = 0.0, B1 = 0.0, B2 = 0.0, B3 = 0.0;
#pragma unroll
        for (int k = 0; k < KR; k++) {
          if (k & 1) { B1 = fma(na[k], inv[k], B1); B2 = fma(nv[k], inv[k], B2); B3 = fma(dz[k], inv[k], B3); }
          else       { A1 = fma(na[k], inv[k], A1); A2 = fma(nv[k], inv[k], A2); A3 = fma(dz[k], inv[k], A3); }
        }
        // shared-memory individuals
        hybrid_sums<KS>(coef, odds, B1, B2, B3);
        const double s2 = A2 + B2;
        double X = fma(odds, s2, A1 + B1), Z = fma(-odds, s2, A3 + B3);
#pragma unroll
        for (int m = 1; m < G; m <<= 1) {
          X += __shfl_xor_sync(kFull, X, m);
          Z += __shfl_xor_sync(kFull, Z, m);
        }
        num = fma(odds, X, num);
        const double dmn = fmax(fma(odds, Z, dmn_next), num * kMinOddsInv);
        odds = num * rcp_pos(dmn);
        pass_denominators<KR>(a0, a2, hh, odds, S);      // next pass, ahead of the stop test
        dmn_next = dmn + g_sum;
        const double now = num * rcp_pos<true>(num + dmn);
        passes++;
        freq = active ? now : freq;
        site_passes = active ? passes : site_passes;
        active = active && (fabs(prev - now) > kEps) && (passes <= 100);   // gen_func.cpp:1006
        prev = now;
      } while (__any_sync(kFull, active));
      if (site_ok && grp == 0) { A.freq[site] = freq; my_passes += site_passes; }
    }

    // ---- emission refresh (L1 is re-read: the coefficients fold it with F)
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      double pe0 = 1.0;
      if (i < A.n_ind && site_ok) {
        const size_t at = (size_t) i * A.site_block + site;
        const double L0 = k < KR ? a0[k < KR ? k : 0] : coef[(size_t) (k - KR) * 6 * kStride];
        const double L2 = k < KR ? a2[k < KR ? k : 0] : coef[(size_t) ((k - KR) * 6 + 1) * kStride];
        double e0, e1;
        emissions(L0, A.gl1[at], L2, freq, e0, e1);
        *emis_slot(A, i, site) = e1 * rcp_pos<true>(e0);
        if (A.e0) A.e0[at] = e0;
        pe0 = e0;
      }
#pragma unroll
      for (int m = G; m < 32; m <<= 1) pe0 *= __shfl_xor_sync(kFull, pe0, m);
      if (sub == 0 && i < A.n_ind_pad) {
        int e;
        my_mant[i] = split_exponent(my_mant[i] * pe0, e);
        my_expo[i] += e;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) my_passes += __shfl_xor_sync(kFull, my_passes, m);
  if (lane == 0 && my_passes) atomicAdd(A.pass_total, (unsigned long long) my_passes);
  for (unsigned i = threadIdx.x; i < A.n_ind_pad; i += kFreqThreads) {
    double m = 1.0;
    int e = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) { m *= mant_acc[(size_t) w * A.n_ind_pad + i]; e += expo_acc[(size_t) w * A.n_ind_pad + i]; }
    A.loge0_part[(size_t) blockIdx.x * A.n_ind_pad + i] = fma((double) e, 0.6931471805599453, log(m));
  }
}

// Teams of W warps share one site (32 W lanes, K individuals per lane): the
// register-resident scheme for 512 < n_ind <= 4096, e.g. the frequency side of
// a multi-rank run, which always sees ALL individuals.  Per pass the lanes of a
// warp reduce by shuffles and the W warps of a team through shared memory with
// a NAMED barrier of the team only (partials double-buffered by pass parity):
// teams work on different sites and never wait for each other, and the CTA is
// kept small (128 threads for W <= 4) so that two CTAs share an SM.
__device__ __forceinline__ void team_barrier(int team, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(n_threads) : "memory");
}

template <int W, int K, int THREADS>
__global__ void __launch_bounds__(THREADS)
freq_emission_team(FreqArgs A, unsigned n_site_tiles) {
  constexpr int kWarps = THREADS / 32;
  constexpr int kTeams = kWarps / W;
  constexpr int G = 32 * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int team = warp / W, wt = warp % W;
  const int grp = wt * 32 + lane;
  // per (team, individual): running product of e0 as mantissa in [1,2) + exponent (see freq_emission_warp)
  extern __shared__ __align__(16) double team_smem[];
  double *mant_acc = team_smem;                                                        // [kTeams][n_ind_pad]
  int *expo_acc = reinterpret_cast<int *>(team_smem + (size_t) kTeams * A.n_ind_pad);   // [kTeams][n_ind_pad]
  __shared__ double2 part[2][kTeams][W];
  __shared__ double gpart[kTeams][W];
  for (unsigned i = threadIdx.x; i < kTeams * A.n_ind_pad; i += THREADS) { mant_acc[i] = 1.0; expo_acc[i] = 0; }
  __syncthreads();
  double *my_mant = mant_acc + (size_t) team * A.n_ind_pad;
  int *my_expo = expo_acc + (size_t) team * A.n_ind_pad;
  unsigned my_passes = 0;

  for (unsigned tile = blockIdx.x; tile < n_site_tiles; tile += gridDim.x) {
    const uint64_t site = (uint64_t) tile * kTeams + team;
    const bool site_ok = site < A.sites_owned;
    const uint64_t sl = site_ok ? site : 0;

    // all loads of the tile are issued before the first coefficient is formed (one memory latency
    // per tile instead of K): the coefficient arrays double as staging for L0, L2, L1, F
    double a0[K], a2[K], hh[K], na[K], nv[K], dz[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      const size_t at = (size_t) (i < A.n_ind ? i : 0) * A.site_block + sl;
      a0[k] = A.gl0[at]; hh[k] = A.gl1[at]; a2[k] = A.gl2[at];
      na[k] = A.post ? A.post[at] : 0.0;
    }
    double g_sum = 0.0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      const IndCoef c = i < A.n_ind ? make_coef(a0[k], hh[k], a2[k], na[k]) : null_coef();
      a0[k] = c.a0; a2[k] = c.a2; hh[k] = c.h; na[k] = c.na; nv[k] = c.nv; dz[k] = c.da - c.na;
      g_sum += c.g;
    }

    double freq = A.update_freq ? kStartFreq : A.freq[sl];
    if (A.update_freq) {
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) g_sum += __shfl_xor_sync(kFull, g_sum, m);
      if (lane == 0) gpart[team][wt] = g_sum;
      team_barrier(team, G);
      g_sum = 0.0;
#pragma unroll
      for (int w = 0; w < W; w++) g_sum += gpart[team][w];

      double num = 0.0, dmn_next = g_sum, odds = kStartOdds, prev = kStartFreq;
      bool active = site_ok;                            // identical in every thread of the team
      int passes = 0;
      double S[K];
      pass_denominators<K>(a0, a2, hh, odds, S);
      while (active) {
        double X, Z;
        pass_sums<K>(S, na, nv, dz, odds, X, Z);
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
          X += __shfl_xor_sync(kFull, X, m);
          Z += __shfl_xor_sync(kFull, Z, m);
        }
        const int buf = passes & 1;
        if (lane == 0) part[buf][team][wt] = make_double2(X, Z);
        team_barrier(team, G);
        X = 0.0; Z = 0.0;
#pragma unroll
        for (int w = 0; w < W; w++) { const double2 q = part[buf][team][w]; X += q.x; Z += q.y; }
        passes++;
        num = fma(odds, X, num);
        const double dmn = fmax(fma(odds, Z, dmn_next), num * kMinOddsInv);   // f -> 1: odds stay finite
        odds = num * rcp_pos(dmn);
        pass_denominators<K>(a0, a2, hh, odds, S);      // next pass, ahead of the stop test (see the warp kernel)
        dmn_next = dmn + g_sum;
        freq = num * rcp_pos<true>(num + dmn);
        active = (fabs(prev - freq) > kEps) && (passes <= 100);   // gen_func.cpp:1006
        prev = freq;
      }
      if (site_ok && grp == 0) { A.freq[site] = freq; my_passes += passes; }
      team_barrier(team, G);                            // gpart / part are reused by the next tile
    }

#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) G * k;
      if (i < A.n_ind && site_ok) {
        const size_t at = (size_t) i * A.site_block + site;
        double e0, e1;
        emissions(a0[k], A.gl1[at], a2[k], freq, e0, e1);
        *emis_slot(A, i, site) = e1 * rcp_pos<true>(e0);
        if (A.e0) A.e0[at] = e0;
        int e;                                          // each (team, individual) slot has one writer
        my_mant[i] = split_exponent(my_mant[i] * e0, e);
        my_expo[i] += e;
      }
    }
  }
  if (my_passes) atomicAdd(A.pass_total, (unsigned long long) my_passes);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < A.n_ind_pad; i += THREADS) {
    double m = 1.0;
    int e = 0;
#pragma unroll
    for (int t = 0; t < kTeams; t++) { m *= mant_acc[(size_t) t * A.n_ind_pad + i]; e += expo_acc[(size_t) t * A.n_ind_pad + i]; }
    A.loge0_part[(size_t) blockIdx.x * A.n_ind_pad + i] = fma((double) e, 0.6931471805599453, log(m));
  }
}

// Any number of individuals: one thread per site walks the individuals in
// index order every pass (the reference's own summation order), re-reading GL
// and posterior through L2.  Slow path for n_ind beyond the register variants.
__global__ void __launch_bounds__(kFreqThreads)
freq_emission_stream(FreqArgs A) {
  for (uint64_t site = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; site < A.sites_owned;
       site += (uint64_t) gridDim.x * blockDim.x) {
    double freq = A.update_freq ? 0.01 : A.freq[site];
    if (A.update_freq) {
      double num = 0.0, den = 0.0, before;
      int passes = 0;
      do {
        before = freq;
        const double omf = 1.0 - freq;
        const double u = omf * omf, v = freq * freq, a = omf * freq;
        double A1 = 0.0, A2 = 0.0, A3 = 0.0, gs = 0.0;
        for (uint64_t i = 0; i < A.n_ind; i++) {
          const size_t at = (size_t) i * A.site_block + site;
          const double F = A.post ? A.post[at] : 0.0;
          IndCoef k = make_coef(A.gl0[at], A.gl1[at], A.gl2[at], F);
          accumulate(k, u, v, a, A1, A2, A3);
          gs += k.g;
        }
        num += fma(a, A1, v * A2); den += fma(a, A3, gs);
        freq = num / den;
      } while (fabs(before - freq) > kEps && passes++ < 100);
      A.freq[site] = freq;
      atomicAdd(A.pass_total, (unsigned long long) min(passes + 1, 101));
    }
    for (uint64_t i = 0; i < A.n_ind; i++) {
      const size_t at = (size_t) i * A.site_block + site;
      double e0, e1;
      emissions(A.gl0[at], A.gl1[at], A.gl2[at], freq, e0, e1);
      *emis_slot(A, i, site) = e1 / e0;
      if (A.e0) A.e0[at] = e0;
    }
  }
}

// Deterministic per-individual sum of log e0 over this rank's sites, recomputed
// from GL and freq (used after the streaming variant).
__global__ void __launch_bounds__(256)
loge0_rowsum(FreqArgs A, unsigned chunks) {
  const uint64_t i = blockIdx.y;
  const uint64_t per = (A.sites_owned + chunks - 1) / chunks;
  const uint64_t lo = (uint64_t) blockIdx.x * per, hi = min(A.sites_owned, lo + per);
  double s = 0.0;
  for (uint64_t site = lo + threadIdx.x; site < hi; site += blockDim.x) {
    const size_t at = (size_t) i * A.site_block + site;
    double e0, e1;
    emissions(A.gl0[at], A.gl1[at], A.gl2[at], A.freq[site], e0, e1);
    s += log(e0);
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if ((int) threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) A.loge0_part[(size_t) blockIdx.x * A.n_ind_pad + i] = red[0];
}

__global__ void reduce_loge0(const double *__restrict__ part, unsigned n_part, uint64_t n_ind_pad,
                             double *__restrict__ out) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ind_pad) return;
  // pairwise-ish: four interleaved accumulators, fixed order
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  unsigned p = 0;
  for (; p + 3 < n_part; p += 4) {
    s0 += part[(size_t) p * n_ind_pad + i];
    s1 += part[(size_t) (p + 1) * n_ind_pad + i];
    s2 += part[(size_t) (p + 2) * n_ind_pad + i];
    s3 += part[(size_t) (p + 3) * n_ind_pad + i];
  }
  for (; p < n_part; p++) s0 += part[(size_t) p * n_ind_pad + i];
  out[i] = (s0 + s1) + (s2 + s3);
}

// Staged chunk of the input file layout [site][individual][3] (natural-log,
// normalised GL) -> three linear-space planes [individual][site_block].
__global__ void gl_ingest(const double *__restrict__ staged, uint64_t n_chunk_sites, uint64_t n_ind,
                          uint64_t first_local_site, uint64_t site_block, double *__restrict__ gl0,
                          double *__restrict__ gl1, double *__restrict__ gl2) {
  const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i = blockIdx.y;
  if (s >= n_chunk_sites) return;
  const double *src = staged + (s * n_ind + i) * 3;
  const size_t at = (size_t) i * site_block + first_local_site + s;
  gl0[at] = exp(src[0]);
  gl1[at] = exp(src[1]);
  gl2[at] = exp(src[2]);
}

// .geno posterior (EM.cpp:369-376): exp(post_prob(GL, HWE(freq, F = path))).
__global__ void geno_posterior(const double *__restrict__ gl0, const double *__restrict__ gl1,
                               const double *__restrict__ gl2, const double *__restrict__ freq,
                               const char *__restrict__ path, uint64_t n_ind, uint64_t site_block,
                               uint64_t path_stride, uint64_t n_chunk, double *__restrict__ out) {
  // all site-indexed pointers are already offset to the first site of the chunk
  const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i = blockIdx.y;
  if (s >= n_chunk) return;
  const size_t at = (size_t) i * site_block + s;
  const double f = freq[s], omf = 1.0 - f, a = omf * f;
  const bool ibd = path[(size_t) i * path_stride + s] != 0;
  const double L0 = gl0[at], L1 = gl1[at], L2 = gl2[at];
  double w0 = L0 * (omf * omf + (ibd ? a : 0.0));
  double w1 = ibd ? 0.0 : L1 * 2.0 * a;
  double w2 = L2 * (f * f + (ibd ? a : 0.0));
  double tot = w0 + w1 + w2;
  if (tot == 0.0) { w0 = L0; w1 = L1; w2 = L2; tot = L0 + L1 + L2; }   // see make_coef()
  double *dst = out + (s * n_ind + i) * 3;
  dst[0] = w0 / tot; dst[1] = w1 / tot; dst[2] = w2 / tot;
}

// FP64 pipe probe: independent DFMA chains, no memory traffic.
__global__ void __launch_bounds__(256) fp64_probe(double *sink, int iters) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  const double m = 0.999999999, c = 1e-12;
  for (int it = 0; it < iters; it++) {
    x0 = fma(x0, m, c); x1 = fma(x1, m, c); x2 = fma(x2, m, c); x3 = fma(x3, m, c);
    x4 = fma(x4, m, c); x5 = fma(x5, m, c); x6 = fma(x6, m, c); x7 = fma(x7, m, c);
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) sink[0] = s;
}

// ---------------------------------------------------------------------------

// Choose lanes-per-site G and individuals-per-lane K <= 16 by a measured cost model (profiles/r02/freq_shapes.md,
// B200, picoseconds per individual SLOT and pass): 0.92 for the slot itself + a per-pass tail that is paid per lane
// group and grows with the group (shuffle levels) and is shared by the K individuals of a lane + a spill penalty
// beyond K = 13.  Padded slots cost like real ones.  Examples: 100 -> (8,13), 125 -> (16,8), 200 -> (16,13),
// 400 -> (32,13); the round-1 rule (fewest padded slots, widest group on ties) took (32,4) for 125: 1.54 vs 1.09 ps.
static double shape_cost(int g, int k) {
  const double tail = g == 4 ? 0.2 : g == 8 ? 0.364 : g == 16 ? 1.144 : 2.18;
  return (double) g * k * (0.922 + tail / k + (k > 13 ? 0.06 * (k - 13) : 0.0));
}

static bool pick_shape(uint64_t n_ind, int &G, int &K) {
  if (const char *force = getenv("NFH_FREQ_G")) {   // tuning override: lanes per site
    const int g = atoi(force);
    const int k = (int) ((n_ind + g - 1) / g);
    if ((g == 4 || g == 8 || g == 16 || g == 32) && k >= 1 && k <= kMaxK) { G = g; K = k; return true; }
  }
  int best_g = 0, best_k = 0;
  double best = 0.0;
  for (int g = 4; g <= 32; g <<= 1) {
    const int k = (int) ((n_ind + g - 1) / g);
    if (k < 1 || k > kMaxK) continue;
    const double c = shape_cost(g, k);
    if (!best_g || c < best) { best = c; best_g = g; best_k = k; }
  }
  G = best_g; K = best_k;
  return best_g != 0;
}

// Hybrid variant (G = 32): 512 < n_ind <= 1,024; up to 832 it beats two-warp teams
// (13.9 vs 15.1 ms at 800 individuals x 125,000 sites).  With fewer individuals the register-only shapes with
// their tile prefetch are faster (200: 10.4 vs 12.6 ms, 400: 11.4 vs 13.4 ms).
static bool pick_hybrid_shape(uint64_t n_ind, int &G, int &KS, int &KR) {
  if (getenv("NFH_FREQ_NO_HYBRID")) return false;
  const int k = (int) ((n_ind + 31) / 32);
  if (n_ind <= 512 || k > 14 + kHybridMaxKSWide) return false;
  G = 32;
  KR = k <= kHybridRegK + kHybridMaxKS ? kHybridRegK : (k <= 13 + kHybridMaxKSWide ? 13 : 14);
  KS = k - KR;
  return true;
}

// Team variant: fewest warps per site W in {2,4,8} with K = ceil(n / 32W) <= 13 (else <= 16).
static bool pick_team_shape(uint64_t n_ind, int &W, int &K) {
  for (int cap : {13, kMaxK})
    for (int w = 2; w <= 8; w <<= 1) {
      const int k = (int) ((n_ind + 32 * w - 1) / (32 * w));
      if (k >= 1 && k <= cap) { W = w; K = k; return true; }
    }
  return false;
}

// Global scratch for the log e0 accumulators of every CTA the grid can have (needed when the tile buffers
// leave no shared memory for them): [grid][warps or teams <= 8][n_ind_pad] x (double + int), rounded up.
size_t freq_acc_scratch_bytes(uint64_t n_ind, uint64_t n_ind_pad, int sm_count) {
  if (n_ind <= 128) return 0;
  return (size_t) sm_count * 4 * ((8 * n_ind_pad * 3 / 2 + 15) / 16 * 16) * sizeof(double);
}

unsigned freq_grid_size(const FreqArgs &a, int sm_count) {
  int G, K, W;
  unsigned per_cta = 0;
  unsigned ctas_per_sm = 4;
  int KS, KR;
  if (a.acc_scratch && pick_hybrid_shape(a.n_ind, G, KS, KR)) { per_cta = (32 / G) * (kFreqThreads / 32); ctas_per_sm = 2; }
  else if (pick_shape(a.n_ind, G, K)) { per_cta = (32 / G) * (kFreqThreads / 32); ctas_per_sm = freq_occupancy(K); }
  else if (pick_team_shape(a.n_ind, W, K)) per_cta = ((W <= 4 ? 128 : 256) / 32) / W;
  if (per_cta) {
    unsigned tiles = (unsigned) ((a.sites_owned + per_cta - 1) / per_cta);
    unsigned cap = (unsigned) sm_count * ctas_per_sm;   // persistent CTAs: one resident wave
    return tiles < cap ? (tiles ? tiles : 1u) : cap;
  }
  return 64;   // chunks of loge0_rowsum
}

// The dynamic shared-memory limit belongs to the (kernel, device) pair: set it once for each, not per launch.
static void smem_limit_once(const void *kernel, int bytes) {
  static std::unordered_map<const void *, unsigned long long> done;   // kernel -> bit per device
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  unsigned long long &mask = done[kernel];
  if (dev >= 0 && dev < 64 && (mask >> dev & 1ull)) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (dev >= 0 && dev < 64) mask |= 1ull << dev;
}

template <int G, int K>
static void launch_warp_variant(const FreqArgs &a, unsigned grid, cudaStream_t st) {
  const unsigned per_cta = FreqTile<G>::kSitesPerCta;
  const unsigned tiles = (unsigned) ((a.sites_owned + per_cta - 1) / per_cta);
  const size_t acc = FreqTile<G>::acc_doubles(a.n_ind_pad) * sizeof(double);
  // all resident CTAs of an SM must fit their double buffers in its 228 KB of shared memory (1 KB reserved each)
  const size_t smem_cap = (size_t) 228 * 1024 / freq_occupancy(K) - 1024 - 256;
  const size_t bufs = 2 * FreqTile<G>::tile_bytes(a.n_ind) + FreqTile<G>::kAlign;   // + alignment slack
  const bool want = a.use_maps && getenv("NFH_FREQ_NO_PREFETCH") == nullptr;
  smem_limit_once((const void *) freq_emission_warp<G, K, 1>, (int) smem_cap);
  smem_limit_once((const void *) freq_emission_warp<G, K, 2>, (int) smem_cap);
  if (want && acc + bufs <= smem_cap) freq_emission_warp<G, K, 1><<<grid, kFreqThreads, acc + bufs, st>>>(a, tiles);
  else if (want && a.acc_scratch && bufs <= smem_cap) freq_emission_warp<G, K, 2><<<grid, kFreqThreads, bufs, st>>>(a, tiles);
  else freq_emission_warp<G, K, 0><<<grid, kFreqThreads, acc, st>>>(a, tiles);
}

// Tensor maps of the four planes for the lane-group shape n_ind selects: [n_ind_pad][site_block] FP64,
// box = (<= 256 rows) x (<= 16 sites), swizzle = box row length.  cuTensorMapEncodeTiled is taken from
// the driver at run time so that the library does not link libcuda.
template <int G>
static bool encode_maps(FreqArgs &a, const double *post_plane, void *encode) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  using Tile = FreqTile<G>;
  const double *planes[4] = {a.gl0, a.gl1, a.gl2, post_plane};
  const cuuint64_t dims[2] = {a.site_block, a.n_ind_pad};
  const cuuint64_t strides[1] = {a.site_block * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t) Tile::kInner, Tile::box_rows(a.n_ind)};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle swz = Tile::kInnerBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : Tile::kInnerBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int p = 0; p < 4; p++) {
    if (!planes[p]) return false;
    if (((EncodeFn) encode)(&a.maps[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *) planes[p], dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

bool freq_tensor_maps(FreqArgs &a, const double *post_plane) {
  a.use_maps = 0;
  static void *encode = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &encode, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      encode = nullptr;
  }
  int G, K;
  if (!encode || !pick_shape(a.n_ind, G, K)) return false;
  bool ok = false;
  switch (G) {
    case 4: ok = encode_maps<4>(a, post_plane, encode); break;
    case 8: ok = encode_maps<8>(a, post_plane, encode); break;
    case 16: ok = encode_maps<16>(a, post_plane, encode); break;
    case 32: ok = encode_maps<32>(a, post_plane, encode); break;
  }
  a.use_maps = ok ? 1 : 0;
  return ok;
}

template <int W, int K>
static void launch_team_variant(const FreqArgs &a, unsigned grid, cudaStream_t st) {
  constexpr int kThreads = W <= 4 ? 128 : 256;
  constexpr int kTeams = (kThreads / 32) / W;
  const unsigned tiles = (unsigned) ((a.sites_owned + kTeams - 1) / kTeams);
  size_t smem = (size_t) kTeams * a.n_ind_pad * (sizeof(double) + sizeof(int));
  smem_limit_once((const void *) freq_emission_team<W, K, kThreads>, 160 * 1024);
  freq_emission_team<W, K, kThreads><<<grid, kThreads, smem, st>>>(a, tiles);
}

#define NFH_K_CASES(MACRO)                                                                          \
  MACRO(1) MACRO(2) MACRO(3) MACRO(4) MACRO(5) MACRO(6) MACRO(7) MACRO(8) MACRO(9) MACRO(10) MACRO(11) \
  MACRO(12) MACRO(13) MACRO(14) MACRO(15) MACRO(16)

template <int G>
static bool dispatch_k(int K, const FreqArgs &a, unsigned grid, cudaStream_t st) {
  switch (K) {
#define NFH_CASE(k) case k: launch_warp_variant<G, k>(a, grid, st); return true;
    NFH_K_CASES(NFH_CASE)
#undef NFH_CASE
    default: return false;
  }
}

template <int W>
static bool dispatch_team_k(int K, const FreqArgs &a, unsigned grid, cudaStream_t st) {
  switch (K) {
#define NFH_CASE(k) case k: launch_team_variant<W, k>(a, grid, st); return true;
    NFH_K_CASES(NFH_CASE)
#undef NFH_CASE
    default: return false;
  }
}

template <int G, int KS, int KR>
static void launch_hybrid_variant(const FreqArgs &a, unsigned grid, cudaStream_t st) {
  const unsigned per_cta = (32 / G) * (kFreqThreads / 32);
  const unsigned tiles = (unsigned) ((a.sites_owned + per_cta - 1) / per_cta);
  const size_t smem = (size_t) KS * 6 * kFreqThreads * sizeof(double);
  smem_limit_once((const void *) freq_emission_hybrid<G, KS, KR>, (int) smem);
  freq_emission_hybrid<G, KS, KR><<<grid, kFreqThreads, smem, st>>>(a, tiles);
}

template <int G>
static bool dispatch_hybrid(int KS, int KR, const FreqArgs &a, unsigned grid, cudaStream_t st) {
#define NFH_CASE(ks, kr) if (KS == ks && KR == kr) { launch_hybrid_variant<G, ks, kr>(a, grid, st); return true; }
  NFH_CASE(5, 12) NFH_CASE(6, 12) NFH_CASE(7, 12) NFH_CASE(8, 12) NFH_CASE(9, 12) NFH_CASE(10, 12)
  NFH_CASE(11, 12) NFH_CASE(12, 12) NFH_CASE(13, 12) NFH_CASE(14, 12)
  NFH_CASE(14, 13) NFH_CASE(15, 13) NFH_CASE(16, 13) NFH_CASE(17, 13) NFH_CASE(18, 13) NFH_CASE(18, 14)
#undef NFH_CASE
  return false;
}

int launch_freq_emission(const FreqArgs &a, unsigned grid, cudaStream_t st) {
  int G, K, W, KS, KR;
  if (a.acc_scratch && pick_hybrid_shape(a.n_ind, G, KS, KR)) {
    bool ok = false;
    if (G == 32) ok = dispatch_hybrid<32>(KS, KR, a, grid, st);
    if (ok) return 1;
  }
  if (pick_shape(a.n_ind, G, K)) {
    bool ok = false;
    switch (G) {
      case 4: ok = dispatch_k<4>(K, a, grid, st); break;
      case 8: ok = dispatch_k<8>(K, a, grid, st); break;
      case 16: ok = dispatch_k<16>(K, a, grid, st); break;
      case 32: ok = dispatch_k<32>(K, a, grid, st); break;
    }
    if (ok) return 1;
  } else if (pick_team_shape(a.n_ind, W, K)) {
    bool ok = false;
    switch (W) {
      case 2: ok = dispatch_team_k<2>(K, a, grid, st); break;
      case 4: ok = dispatch_team_k<4>(K, a, grid, st); break;
      case 8: ok = dispatch_team_k<8>(K, a, grid, st); break;
    }
    if (ok) return 1;
  }
  unsigned blocks = (unsigned) ((a.sites_owned + kFreqThreads - 1) / kFreqThreads);
  freq_emission_stream<<<blocks ? blocks : 1, kFreqThreads, 0, st>>>(a);
  dim3 g2(grid, (unsigned) a.n_ind);
  loge0_rowsum<<<g2, 256, 0, st>>>(a, grid);
  return 2;
}

__global__ void fill_double(double *__restrict__ dst, double value, size_t n) {
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) dst[i] = value;
}

void launch_fill(double *dst, double value, size_t n, cudaStream_t st) {
  fill_double<<<1184, 256, 0, st>>>(dst, value, n);
}

void launch_reduce_loge0(const double *part, unsigned n_part, uint64_t n_ind_pad, double *out, cudaStream_t st) {
  reduce_loge0<<<(unsigned) ((n_ind_pad + 127) / 128), 128, 0, st>>>(part, n_part, n_ind_pad, out);
}

void launch_gl_ingest(const double *staged, uint64_t n_chunk_sites, uint64_t n_ind, uint64_t first_local_site,
                      uint64_t site_block, double *gl0, double *gl1, double *gl2, cudaStream_t st) {
  dim3 grid((unsigned) ((n_chunk_sites + 255) / 256), (unsigned) n_ind);
  gl_ingest<<<grid, 256, 0, st>>>(staged, n_chunk_sites, n_ind, first_local_site, site_block, gl0, gl1, gl2);
}

void launch_geno_posterior(const double *gl0, const double *gl1, const double *gl2, const double *freq,
                           const char *path, uint64_t n_ind, uint64_t site_block, uint64_t path_stride,
                           uint64_t n_chunk, double *out, cudaStream_t st) {
  dim3 grid((unsigned) ((n_chunk + 255) / 256), (unsigned) n_ind);
  geno_posterior<<<grid, 256, 0, st>>>(gl0, gl1, gl2, freq, path, n_ind, site_block, path_stride, n_chunk, out);
}

double launch_fp64_probe(cudaStream_t st, int sm_count) {
  double *sink = nullptr;
  cudaMalloc(&sink, sizeof(double));
  const int iters = 1 << 14, threads = 256;
  const unsigned grid = (unsigned) sm_count * 8u;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fp64_probe<<<grid, threads, 0, st>>>(sink, iters);   // warm-up
  cudaEventRecord(e0, st);
  fp64_probe<<<grid, threads, 0, st>>>(sink, iters);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  const double fmas = (double) grid * threads * (double) iters * 8.0;
  return 2.0 * fmas / (ms * 1e-3);
}

}  // namespace nfh

// nfh_estep.cu - fused forward-backward E-step and the batched forward-only
// objective, as chunked scans of scaled 2x2 products.
//
// Replaces, for all individuals at once:
//   forward()  shared/HMM.cpp:6-28      backward() shared/HMM.cpp:33-60
//   ind_lkl + clamped posterior          EM.cpp:178-185, check_interv gen_func.cpp:55-70
//   Fw/Bw consistency check              EM.cpp:166-170
//   lkl() (BFGS objective)               EM.cpp:449-464
//
// E-step = three launches:
//   estep_chunk_products : every thread reduces its 33 consecutive sites to one
//                          scaled 2x2 product (tile staged in shared memory by TMA).
//                          The CTA also reduces its 128 chunk products to one tile product.
//   estep_tile_carries   : per individual, chain of the tile products: forward
//                          carry into and backward carry out of every tile, and
//                          the log-likelihood computed both ways.
//   estep_chunk_apply    : the CTA turns the tile carries plus its 128 chunk products
//                          into per-chunk carries (warp scans), then every thread
//                          re-reads its 33 sites (TMA-staged), runs the forward and
//                          the backward vector recursion and writes the clamped IBD
//                          posterior (TMA store).
// HBM traffic per individual-site: 8 B + 8 B of emission ratio read, 8 B of
// posterior written, ~2.4 B of chunk products/carries; site distances come
// from L2.
#include "nfh_device.cuh"
#include "nfh_kernels.h"

namespace nfh {
namespace v1 {

struct TileSmem {
  alignas(128) double r[kTile];     // emission ratio; overwritten with the posterior by estep_chunk_apply
  alignas(128) double d[kTile];     // distance (Mb); overwritten with kappa
  alignas(8) uint64_t bar;
  double tab[64];
  double2 ck[kScanThreads * (kChunk / kSub)];   // forward checkpoints of estep_chunk_apply
};

// Stage one tile of the emission plane and of the distance vector.
__device__ __forceinline__ void stage_tile(TileSmem &sm, const double *__restrict__ emis_tile,
                                           const double *__restrict__ dist_tile) {
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
  }
  load_exp_table(sm.tab);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&sm.bar, 2 * kTileBytes);
    tma_load_1d(sm.r, emis_tile, kTileBytes, &sm.bar);
    tma_load_1d(sm.d, dist_tile, kTileBytes, &sm.bar);
  }
  mbar_wait(&sm.bar, 0);
}

// Number of real sites among this thread's kChunk (the last tile is padded).
__device__ __forceinline__ int valid_sites(uint64_t first_site, uint64_t n_sites) {
  return first_site >= n_sites ? 0 : (int) min((uint64_t) kChunk, n_sites - first_site);
}

__global__ void __launch_bounds__(kScanThreads)
estep_chunk_products(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
                     const double *__restrict__ alpha, double4 *__restrict__ chunk_prod,
                     TileProd *__restrict__ tile_prod, uint64_t n_rows, uint64_t n_sites, uint64_t site_block,
                     uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  stage_tile(sm, emis + blocked_index(row, tile_first, n_rows, site_block), dist + tile_first);

  const double F = indF[row], al = alpha[row];
  const double q0 = 1.0 - F, q1 = F;
  const double *r = sm.r + threadIdx.x * kChunk;
  const double *d = sm.d + threadIdx.x * kChunk;

  // Straight-line bodies of kBody sites, no branches and no selects: padding sites are the identity
  // because the context keeps r = 1 and d = 0 there (kappa = 0, scale term 0).
  M2 m = identity2();
  int e = 0;
  double ls = 0.0;
  constexpr int kBody = 6;
#pragma unroll 1
  for (int j0 = 0; j0 + kBody <= kChunk; j0 += kBody) {
#pragma unroll
    for (int i = 0; i < kBody; i++) {
      const int j = j0 + i;
      const double kap = site_kappa(al * d[j], sm.tab, ls);
      apply_site(m, kap * q0, kap * q1, r[j]);
    }
    e += renorm(m);
  }
#pragma unroll
  for (int j = (kChunk / kBody) * kBody; j < kChunk; j++) {
    const double kap = site_kappa(al * d[j], sm.tab, ls);
    apply_site(m, kap * q0, kap * q1, r[j]);
  }
  e += renorm(m);
  // per-chunk product (direction only: the apply kernel is scale free)
  chunk_prod[((size_t) row * n_tiles + tile) * kScanThreads + threadIdx.x] = make_double4(m.a, m.b, m.c, m.d);

  // per-tile product with its scale, for the carry kernel and the log-likelihood
  __shared__ M2 sm_m[kScanThreads / 32];
  __shared__ int sm_e[kScanThreads / 32];
  __shared__ double sm_l[kScanThreads / 32];
  warp_ordered_product(m, e);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ls += __shfl_down_sync(kFull, ls, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm_m[warp] = m; sm_e[warp] = e; sm_l[warp] = ls; }
  __syncthreads();
  if (threadIdx.x == 0) {
    M2 acc = sm_m[0];
    int ae = sm_e[0];
    double al_sum = sm_l[0];
#pragma unroll
    for (int w = 1; w < kScanThreads / 32; w++) {
      acc = matmul(acc, sm_m[w]);
      ae += sm_e[w] + renorm(acc);
      al_sum += sm_l[w];
    }
    TileProd out;
    out.a = acc.a; out.b = acc.b; out.c = acc.c; out.d = acc.d; out.e = (double) ae; out.l = al_sum;
    tile_prod[(size_t) row * n_tiles + tile] = out;
  }
}

// One warp per individual.  Tiles are taken 32 at a time: a coalesced load, an
// inclusive warp scan of the 2x2 products (with exponents), and the running
// vector gives the carry of each of the 32 tiles; forwards from q, then
// backwards from 1.  Also the log-likelihood both ways (EM.cpp:166).
__global__ void __launch_bounds__(128)
estep_tile_carries(const TileProd *__restrict__ tile_prod, const double *__restrict__ indF,
                   const double *__restrict__ loge0_sum, double2 *__restrict__ fwd_carry,
                   double2 *__restrict__ bwd_carry, double *__restrict__ ind_lkl, int *__restrict__ status,
                   uint32_t n_rows_valid, uint32_t n_tiles) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows_valid) return;                       // warp-uniform
  const double F = indF[row];
  const double q0 = 1.0 - F, q1 = F;
  const TileProd *tp = tile_prod + (size_t) row * n_tiles;
  double2 *fc = fwd_carry + (size_t) row * n_tiles;
  double2 *bc = bwd_carry + (size_t) row * n_tiles;
  const uint32_t n_groups = (n_tiles + 31) / 32;

  // ---- forwards
  double x0 = q0, x1 = q1, lsum = 0.0;
  long long ex = 0;
  for (uint32_t g = 0; g < n_groups; g++) {
    const uint32_t t = g * 32 + lane;
    M2 m = identity2();
    int e = 0;
    double l = 0.0;
    if (t < n_tiles) {
      const TileProd p = tp[t];
      m.a = p.a; m.b = p.b; m.c = p.c; m.d = p.d; e = (int) p.e; l = p.l;
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const M2 o = shfl_up_m(m, off);
      const int oe = __shfl_up_sync(kFull, e, off);
      if (lane >= off) { m = matmul(o, m); e += oe + renorm(m); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(kFull, l, off);
    lsum += l;
    const M2 before = shfl_up_m(m, 1);
    double c0 = x0, c1 = x1;
    if (lane > 0) { c0 = fma(x0, before.a, x1 * before.c); c1 = fma(x0, before.b, x1 * before.d); renorm2(c0, c1); }
    if (t < n_tiles) fc[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 31); tot.b = __shfl_sync(kFull, m.b, 31);
    tot.c = __shfl_sync(kFull, m.c, 31); tot.d = __shfl_sync(kFull, m.d, 31);
    const int te = __shfl_sync(kFull, e, 31);
    const double y0 = fma(x0, tot.a, x1 * tot.c), y1 = fma(x0, tot.b, x1 * tot.d);
    x0 = y0; x1 = y1;
    ex += te + renorm2(x0, x1);
  }
  const double base = lsum + loge0_sum[row];
  const double lf = log(x0 + x1) + (double) ex * kLn2 + base;

  // ---- backwards
  double b0 = 1.0, b1 = 1.0;
  long long eb = 0;
  for (uint32_t g = n_groups; g-- > 0;) {
    const uint32_t t = g * 32 + lane;
    M2 m = identity2();
    int e = 0;
    if (t < n_tiles) {
      const TileProd p = tp[t];
      m.a = p.a; m.b = p.b; m.c = p.c; m.d = p.d; e = (int) p.e;
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {             // inclusive suffix: P_t ... P_{last of group}
      const M2 o = shfl_down_m(m, off);
      const int oe = __shfl_down_sync(kFull, e, off);
      if (lane + off < 32) { m = matmul(m, o); e += oe + renorm(m); }
    }
    const M2 after = shfl_down_m(m, 1);
    double c0 = b0, c1 = b1;
    if (lane < 31) { c0 = fma(after.a, b0, after.b * b1); c1 = fma(after.c, b0, after.d * b1); renorm2(c0, c1); }
    if (t < n_tiles) bc[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 0); tot.b = __shfl_sync(kFull, m.b, 0);
    tot.c = __shfl_sync(kFull, m.c, 0); tot.d = __shfl_sync(kFull, m.d, 0);
    const int te = __shfl_sync(kFull, e, 0);
    const double y0 = fma(tot.a, b0, tot.b * b1), y1 = fma(tot.c, b0, tot.d * b1);
    b0 = y0; b1 = y1;
    eb += te + renorm2(b0, b1);
  }
  const double lb = log(fma(q0, b0, q1 * b1)) + (double) eb * kLn2 + base;

  if (lane == 0) {
    ind_lkl[row] = lf;
    if (lf != lf || lb != lb) atomicOr(status, kFlagNaN);
    else if (fabs(lf - lb) > 1e-3) atomicOr(status, kFlagFwBw);   // EM.cpp:166
  }
}

// Carries of this thread's chunk from the tile carries and the tile's 128
// chunk products: warp prefix/suffix scans (direction only) plus a 4-way
// combine through shared memory.
__device__ __forceinline__ void chunk_carries(const double4 *__restrict__ chunk_prod_tile, double2 tile_fwd,
                                              double2 tile_bwd, double &a0, double &a1, double &b0, double &b1) {
  constexpr int kWarps = kScanThreads / 32;
  __shared__ M2 warp_tot[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double4 mine4 = chunk_prod_tile[threadIdx.x];
  M2 mine; mine.a = mine4.x; mine.b = mine4.y; mine.c = mine4.z; mine.d = mine4.w;
  M2 pre = mine, suf = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    M2 o = shfl_up_m(pre, off);
    if (lane >= off) { pre = matmul(o, pre); renorm(pre); }
    M2 u = shfl_down_m(suf, off);
    if (lane + off < 32) { suf = matmul(suf, u); renorm(suf); }
  }
  if (lane == 31) warp_tot[warp] = pre;
  __syncthreads();
  a0 = tile_fwd.x; a1 = tile_fwd.y;
  for (int w = 0; w < warp; w++) {                 // warp-uniform trip count
    const M2 p = warp_tot[w];
    const double y0 = fma(a0, p.a, a1 * p.c), y1 = fma(a0, p.b, a1 * p.d);
    a0 = y0; a1 = y1;
    renorm2(a0, a1);
  }
  b0 = tile_bwd.x; b1 = tile_bwd.y;
  for (int w = kWarps - 1; w > warp; w--) {
    const M2 p = warp_tot[w];
    const double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
    b0 = y0; b1 = y1;
    renorm2(b0, b1);
  }
  const M2 before = shfl_up_m(pre, 1);     // product of lanes < me (valid for lane > 0)
  const M2 after = shfl_down_m(suf, 1);    // product of lanes > me (valid for lane < 31)
  if (lane > 0) {
    const double y0 = fma(a0, before.a, a1 * before.c), y1 = fma(a0, before.b, a1 * before.d);
    a0 = y0; a1 = y1;
  }
  if (lane < 31) {
    const double y0 = fma(after.a, b0, after.b * b1), y1 = fma(after.c, b0, after.d * b1);
    b0 = y0; b1 = y1;
  }
}

__global__ void __launch_bounds__(kScanThreads)
estep_chunk_apply(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
                  const double *__restrict__ alpha, const double4 *__restrict__ chunk_prod,
                  const double2 *__restrict__ fwd_carry, const double2 *__restrict__ bwd_carry,
                  double *__restrict__ post, PeerWindows peers, int *__restrict__ status, uint64_t n_rows,
                  uint64_t n_sites, uint64_t site_block, uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  const size_t tile_at = blocked_index(row, tile_first, n_rows, site_block);
  stage_tile(sm, emis + tile_at, dist + tile_first);

  const int n_valid = valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, n_sites);
  const double F = indF[row], al = alpha[row];
  const double q0 = 1.0 - F, q1 = F;
  double *r = sm.r + threadIdx.x * kChunk;     // becomes the posterior
  double *d = sm.d + threadIdx.x * kChunk;     // becomes kappa
  double cf0, cf1, cb0, cb1;
  chunk_carries(chunk_prod + ((size_t) row * n_tiles + tile) * kScanThreads, fwd_carry[(size_t) row * n_tiles + tile],
                bwd_carry[(size_t) row * n_tiles + tile], cf0, cf1, cb0, cb1);

  // sweep 1: kappa for every site (kept in shared memory) and the forward
  // vector at the start of each sub-block (checkpoints, also in shared memory)
  double a0 = cf0, a1 = cf1;
  double2 *ck = sm.ck + threadIdx.x * (kChunk / kSub);
#pragma unroll 1
  for (int sb = 0; sb < kChunk / kSub; sb++) {
    renorm2(a0, a1);
    ck[sb] = make_double2(a0, a1);
#pragma unroll
    for (int i = 0; i < kSub; i++) {
      const int j = sb * kSub + i;
      const double kap = site_kappa(al * d[j], sm.tab);     // padding: d = 0 -> kappa = 0
      d[j] = kap;
      forward_site(a0, a1, kap * q0, kap * q1, r[j]);
      if (i == kSub / 2) renorm2(a0, a1);
    }
  }

  // sweep 2, sub-blocks from the right: rebuild the forward vectors of the
  // sub-block in registers, then run the backward vector through it.
  double b0 = cb0, b1 = cb1;
  bool bad = false;
#pragma unroll 1
  for (int sb = kChunk / kSub - 1; sb >= 0; sb--) {
    double rr[kSub], k0[kSub], k1[kSub], f0[kSub], f1[kSub];
    const double2 c = ck[sb];
    a0 = c.x; a1 = c.y;
#pragma unroll
    for (int i = 0; i < kSub; i++) {
      const int j = sb * kSub + i;
      rr[i] = r[j];
      const double kap = d[j];
      k0[i] = kap * q0; k1[i] = kap * q1;
      forward_site(a0, a1, k0[i], k1[i], rr[i]);
      if (i == kSub / 2) renorm2(a0, a1);
      f0[i] = a0; f1[i] = a1;
    }
#pragma unroll
    for (int i = kSub - 1; i >= 0; i--) {
      const int j = sb * kSub + i;
      const double num = f1[i] * b1;
      const double den = fma(f0[i], b0, num);
      double p = num * rcp_pos(den);
      bad |= (p != p);
      p = (p < kEps) ? 0.0 : p;                // check_interv, gen_func.cpp:59-66
      p = (p > 1.0 - kEps) ? 1.0 : p;
      backward_site(b0, b1, k0[i], k1[i], rr[i]);   // identity on padding sites (kappa = 0, r = 1)
      if (i == kSub / 2) renorm2(b0, b1);
      r[j] = j < n_valid ? p : 0.0;
    }
    renorm2(b0, b1);
  }
  if (bad) atomicOr(status, kFlagNaN);

  fence_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_1d(post + tile_at, sm.r, kTileBytes);       // individual-major copy (output, nfh_get_posterior)
    if (peers.direct) {
      // this tile belongs to site block b: also store it into rank b's frequency-side window over
      // NVLink, source block = me - the posterior "all-to-all" happens inside this kernel
      const uint64_t b = tile_first / site_block;
      double *dst = peers.base[b] + ((uint64_t) peers.rank * n_rows + row) * site_block + (tile_first - b * site_block);
      tma_store_1d(dst, sm.r, kTileBytes);
    }
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------
// Batched forward-only objective: up to kMaxPoints (F, alpha) points of one
// individual share one staged read of its emissions.
// ---------------------------------------------------------------------------

constexpr int kLklThreads = 2 * kScanThreads;   // two halves share one staged tile, each takes part of the points

struct LklSmem {
  alignas(128) double r[kTile];     // emission ratio
  alignas(128) double d[kTile];     // distance (Mb)
  alignas(8) uint64_t bar;
  double tab[64];
  M2 m[kMaxPoints][kScanThreads / 32];
  int e[kMaxPoints][kScanThreads / 32];
  double l[kMaxPoints][kScanThreads / 32];
};

// FP64 instructions per site of a set of points: 13 per distinct alpha (kappa) + 10 per point (2x2 update)
__host__ __device__ constexpr int lkl_cost(int n_same, int n_other) {
  return 13 * ((n_same > 0 ? 1 : 0) + n_other) + 10 * (n_same + n_other);
}
// The points of a group are ordered [NS sharing alpha[0]] [NA with their own alpha].  The first k go to
// half 0 of the CTA, the rest to half 1; k balances the two instruction counts.
__host__ __device__ constexpr int lkl_split(int NS, int NA) {
  int best = 1, best_cost = 1 << 30;
  for (int k = 1; k <= NS + NA; k++) {
    const int sa = k < NS ? k : NS, oa = k - sa;
    const int ca = lkl_cost(sa, oa), cb = lkl_cost(NS - sa, NA - oa);
    const int c = ca > cb ? ca : cb;
    if (c < best_cost) { best_cost = c; best = k; }
  }
  return best;
}

// One thread's chunk of kChunk sites for points [first, first + NS + NA) of the group: the first NS
// share alpha[first] (one kappa per site for all of them), the next NA each have their own.  The
// layout is fixed per group, so the whole body is straight-line code.  Lane 0 of every warp leaves
// the warp's ordered product in shared memory.
template <int NS, int NA>
__device__ __forceinline__ void lkl_chunk_run(const LklGroup &g, int first, LklSmem &sm, int t,
                                              double4 *__restrict__ emit_chunks) {
  constexpr int NP = NS + NA;
  constexpr int kBody = 6;
  const double *r = sm.r + t * kChunk;
  const double *d = sm.d + t * kChunk;
  M2 m[NP];
  int e[NP];
  double ls[1 + NA];          // scale sums: one for the shared alpha, one per extra alpha
  double q1[NP], q0[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) { m[p] = identity2(); e[p] = 0; q1[p] = g.F[first + p]; q0[p] = 1.0 - g.F[first + p]; }
#pragma unroll
  for (int a = 0; a <= NA; a++) ls[a] = 0.0;

  auto site = [&](int j) {
    const double dj = d[j];
    const double rj = r[j];                          // padding: r = 1, d = 0 -> identity
    if (NS > 0) {
      const double ks = site_kappa(g.alpha[first] * dj, sm.tab, ls[0]);
#pragma unroll
      for (int p = 0; p < NS; p++) apply_site(m[p], ks * q0[p], ks * q1[p], rj);
    }
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const double ka = site_kappa(g.alpha[first + NS + a] * dj, sm.tab, ls[1 + a]);
      apply_site(m[NS + a], ka * q0[NS + a], ka * q1[NS + a], rj);
    }
  };
#pragma unroll 1
  for (int j0 = 0; j0 + kBody <= kChunk; j0 += kBody) {
#pragma unroll
    for (int i = 0; i < kBody; i++) site(j0 + i);
#pragma unroll
    for (int p = 0; p < NP; p++) e[p] += renorm(m[p]);
  }
#pragma unroll
  for (int j = (kChunk / kBody) * kBody; j < kChunk; j++) site(j);

  const int warp = (t >> 5), lane = t & 31;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    e[p] += renorm(m[p]);
    // the group's first point doubles as the E-step's forward product of this chunk (direction only)
    if (p == 0 && first == 0 && emit_chunks) emit_chunks[t] = make_double4(m[0].a, m[0].b, m[0].c, m[0].d);
    warp_ordered_product(m[p], e[p]);
    double l = ls[p < NS ? 0 : 1 + (p - NS)];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l += __shfl_down_sync(kFull, l, off);
    if (lane == 0) { sm.m[first + p][warp] = m[p]; sm.e[first + p][warp] = e[p]; sm.l[first + p][warp] = l; }
  }
}

template <int NS, int NA>
__device__ __forceinline__ void lkl_tile_halves(const LklGroup &g, LklSmem &sm, int half, int t,
                                                double4 *__restrict__ emit_chunks) {
  constexpr int k = lkl_split(NS, NA);
  constexpr int NSa = k < NS ? k : NS, NAa = k - NSa, NSb = NS - NSa, NAb = NA - NAa;
  if (half == 0) {
    lkl_chunk_run<NSa, NAa>(g, 0, sm, t, emit_chunks);
  } else {
    if constexpr (NSb + NAb > 0) lkl_chunk_run<NSb, NAb>(g, k, sm, t, nullptr);
  }
}

// One CTA per (tile, group): the tile of emission ratios and distances is staged once by TMA and read by
// both halves of the CTA (2 x 128 threads, each thread kChunk sites), which doubles the warps an SM can
// hold for the same shared memory (the tile, not registers, limits occupancy: 3 CTAs per SM).
__global__ void __launch_bounds__(kLklThreads)
lkl_tile_products(const double *__restrict__ emis, const double *__restrict__ dist,
                  const LklGroup *__restrict__ groups, TileProd *__restrict__ tile_prod, uint64_t n_rows,
                  uint64_t n_sites, uint64_t site_block, uint32_t n_tiles, double4 *__restrict__ emit_chunk_prod,
                  TileProd *__restrict__ emit_tile_prod) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LklSmem &sm = *reinterpret_cast<LklSmem *>(smem_raw);
  const uint32_t tile = blockIdx.x, grp = blockIdx.y;
  const LklGroup g = groups[grp];
  const uint64_t tile_first = (uint64_t) tile * kTile;
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
  }
  load_exp_table(sm.tab);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&sm.bar, 2 * kTileBytes);
    tma_load_1d(sm.r, emis + blocked_index((uint64_t) g.ind, tile_first, n_rows, site_block), kTileBytes, &sm.bar);
    tma_load_1d(sm.d, dist + tile_first, kTileBytes, &sm.bar);
  }
  mbar_wait(&sm.bar, 0);
  const int half = threadIdx.x / kScanThreads, t = threadIdx.x % kScanThreads;

  double4 *emit_chunks =
      emit_chunk_prod ? emit_chunk_prod + ((size_t) g.ind * n_tiles + tile) * kScanThreads : nullptr;
#define NFH_LKL(ns, na) case (ns) * 8 + (na): lkl_tile_halves<ns, na>(g, sm, half, t, emit_chunks); break;
  switch (g.n_same * 8 + (g.npts - g.n_same)) {
    NFH_LKL(1, 0) NFH_LKL(1, 1) NFH_LKL(1, 2) NFH_LKL(1, 3) NFH_LKL(1, 4)
    NFH_LKL(2, 0) NFH_LKL(2, 1) NFH_LKL(2, 2) NFH_LKL(2, 3)
    NFH_LKL(3, 0) NFH_LKL(3, 1) NFH_LKL(3, 2)
    NFH_LKL(4, 0) NFH_LKL(4, 1)
    NFH_LKL(5, 0)
    default: break;
  }
#undef NFH_LKL
  __syncthreads();
  if ((int) threadIdx.x < g.npts) {
    const int p = threadIdx.x;
    M2 acc = sm.m[p][0];
    int ae = sm.e[p][0];
    double al_sum = sm.l[p][0];
    for (int w = 1; w < kScanThreads / 32; w++) {
      acc = matmul(acc, sm.m[p][w]);
      ae += sm.e[p][w] + renorm(acc);
      al_sum += sm.l[p][w];
    }
    TileProd out;
    out.a = acc.a; out.b = acc.b; out.c = acc.c; out.d = acc.d; out.e = (double) ae; out.l = al_sum;
    tile_prod[((size_t) grp * kMaxPoints + p) * n_tiles + tile] = out;
    if (p == 0 && emit_tile_prod) emit_tile_prod[(size_t) g.ind * n_tiles + tile] = out;
  }
}

// One warp per (group, point): lanes take contiguous runs of tile products,
// an ordered warp product combines them; lane 0 emits -logLkl.
__global__ void __launch_bounds__(128)
lkl_finish(const TileProd *__restrict__ tile_prod, const LklGroup *__restrict__ groups,
           const double *__restrict__ loge0_sum, double *__restrict__ neg_lkl, uint32_t n_groups, uint32_t n_tiles) {
  const uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t grp = idx / kMaxPoints, p = idx % kMaxPoints;
  if (grp >= n_groups) return;
  const LklGroup &g = groups[grp];
  if ((int) p >= g.npts) return;       // warp-uniform
  const TileProd *tp = tile_prod + ((size_t) grp * kMaxPoints + p) * n_tiles;
  const uint32_t per = (n_tiles + 31) / 32;
  const uint32_t lo = min(n_tiles, lane * per), hi = min(n_tiles, lo + per);
  M2 m = identity2();
  int e = 0;
  double l = 0.0;
  for (uint32_t t = lo; t < hi; t++) {
    const TileProd q = tp[t];
    M2 o; o.a = q.a; o.b = q.b; o.c = q.c; o.d = q.d;
    m = matmul(m, o);
    e += (int) q.e + renorm(m);
    l += q.l;
  }
  warp_ordered_product(m, e);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) l += __shfl_down_sync(kFull, l, off);
  if (lane == 0) {
    const double x0 = 1.0 - g.F[p], x1 = g.F[p];
    const double y0 = fma(x0, m.a, x1 * m.c), y1 = fma(x0, m.b, x1 * m.d);
    neg_lkl[g.out[p]] = -(log(y0 + y1) + (double) e * kLn2 + l + loge0_sum[g.ind]);
  }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------

// per launch: the attribute belongs to the current device, and a process may drive several
static void set_smem_attrs() {
  cudaFuncSetAttribute(estep_chunk_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(TileSmem));
  cudaFuncSetAttribute(estep_chunk_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(TileSmem));
  cudaFuncSetAttribute(lkl_tile_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LklSmem));
}

void launch_estep_v1(const EstepArgs &a, cudaStream_t st) {
  set_smem_attrs();
  dim3 grid(a.n_tiles, (unsigned) a.n_rows_valid);
  estep_chunk_products<<<grid, kScanThreads, sizeof(TileSmem), st>>>(a.emis, a.dist, a.indF, a.alpha, a.chunk_prod,
                                                                     a.tile_prod, a.n_rows, a.n_sites, a.site_block,
                                                                     a.n_tiles);
  estep_tile_carries<<<(unsigned) ((a.n_rows_valid + 3) / 4), 128, 0, st>>>(
      a.tile_prod, a.indF, a.loge0_sum, a.fwd_carry, a.bwd_carry, a.ind_lkl, a.status, (uint32_t) a.n_rows_valid,
      a.n_tiles);
  estep_chunk_apply<<<grid, kScanThreads, sizeof(TileSmem), st>>>(a.emis, a.dist, a.indF, a.alpha, a.chunk_prod,
                                                                  a.fwd_carry, a.bwd_carry, a.post, a.post_peers, a.status,
                                                                  a.n_rows, a.n_sites, a.site_block, a.n_tiles);
}

void launch_estep_tail_v1(const EstepArgs &a, cudaStream_t st) {
  set_smem_attrs();
  dim3 grid(a.n_tiles, (unsigned) a.n_rows_valid);
  estep_tile_carries<<<(unsigned) ((a.n_rows_valid + 3) / 4), 128, 0, st>>>(
      a.tile_prod, a.indF, a.loge0_sum, a.fwd_carry, a.bwd_carry, a.ind_lkl, a.status, (uint32_t) a.n_rows_valid,
      a.n_tiles);
  estep_chunk_apply<<<grid, kScanThreads, sizeof(TileSmem), st>>>(a.emis, a.dist, a.indF, a.alpha, a.chunk_prod,
                                                                  a.fwd_carry, a.bwd_carry, a.post, a.post_peers, a.status,
                                                                  a.n_rows, a.n_sites, a.site_block, a.n_tiles);
}

void launch_lkl_batch_v1(const LklArgs &a, cudaStream_t st) {
  set_smem_attrs();
  dim3 grid(a.n_tiles, a.n_groups);
  lkl_tile_products<<<grid, kLklThreads, sizeof(LklSmem), st>>>(a.emis, a.dist, a.groups, a.tile_prod, a.n_rows,
                                                                 a.n_sites, a.site_block, a.n_tiles, a.emit_chunk_prod,
                                                                 a.emit_tile_prod);
  const unsigned warps = a.n_groups * kMaxPoints;
  lkl_finish<<<(warps + 3) / 4, 128, 0, st>>>(a.tile_prod, a.groups, a.loge0_sum, a.neg_lkl, a.n_groups, a.n_tiles);
}

}  // namespace v1
}  // namespace nfh

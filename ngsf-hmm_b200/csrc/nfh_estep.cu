// nfh_estep.cu - fused forward-backward E-step as a chunked scan of scaled 2x2 products.
//
// Replaces, for all individuals at once:
//   forward()  shared/HMM.cpp:6-28      backward() shared/HMM.cpp:33-60
//   ind_lkl + clamped posterior          EM.cpp:178-185, check_interv gen_func.cpp:55-70
//   Fw/Bw consistency check              EM.cpp:166-170
//
// E-step = three launches:
//   estep_chunk_products : every thread reduces its 33 consecutive sites to one scaled 2x2 product
//                          (two independent half-chunk chains), the CTA reduces its 128 chunk
//                          products to one tile product.
//   estep_tile_carries   : per individual, chain of the tile products: forward carry into and
//                          backward carry out of every tile, log-likelihood computed both ways.
//   estep_chunk_apply    : the CTA turns the tile carries plus its 128 chunk products into per-chunk
//                          carries (warp scans, done while the tile is still in flight), then every
//                          thread runs ONE forward and ONE backward vector pass over its 33 sites -
//                          as two independent chains that meet in the middle - and writes the clamped
//                          IBD posterior (TMA store).
//
// The posterior needs no division per site: f_s . b_s = sum_k f_s(k) b_s(k) is the same number L at
// every site (it is the likelihood), so  p_s = f_s(1) b_s(1) / L  with ONE reciprocal per chunk and the
// exact powers of two of the renormalisations carried along (see apply_chunk).
//
// HBM traffic per individual-site: 8 B + 8 B of emission ratio read, 8 B of posterior written,
// ~2 B of chunk products; site distances come from L2 (shared by all individuals).
// Persistent variants of both sweeps (TMA ring, distances from a transposed copy in L2, two threads per
// chunk) were built and measured slower ("tried and not kept"): with the distances out of
// shared memory their L2 latency lands on every renormalisation window.
#include <algorithm>

#include "nfh_device.cuh"
#include "nfh_estep_math.cuh"
#include "nfh_kernels.h"
#include "nfh_schedule.h"

namespace nfh {

struct TileSmem {
  alignas(128) double r[kTile];     // emission ratio; overwritten with the posterior by estep_chunk_apply
  alignas(128) double d[kTile];     // distance (Mb); overwritten with kappa by estep_chunk_apply
  alignas(8) uint64_t bar;
  double tab[64];
};

struct ApplySmem {                  // stash: the kStash values per thread apply_chunk() keeps in shared memory
  TileSmem tile;
  double stash[kStash][kScanThreads];
};

// Start the bulk copies of one tile of the emission plane and of the distance vector; the caller
// overlaps independent work and then waits on sm.bar (parity 0).
__device__ __forceinline__ void stage_tile_begin(TileSmem &sm, const double *__restrict__ emis_tile,
                                                 const double *__restrict__ dist_tile) {
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&sm.bar, 2 * kTileBytes);
    tma_load_1d(sm.r, emis_tile, kTileBytes, &sm.bar);
    tma_load_1d(sm.d, dist_tile, kTileBytes, &sm.bar);
  }
  load_exp_table(sm.tab);
  __syncthreads();                  // barrier initialised and table visible before anyone waits / reads
}

// Number of real sites among this thread's kChunk (the last tile is padded).
__device__ __forceinline__ int valid_sites(uint64_t first_site, uint64_t n_sites) {
  return first_site >= n_sites ? 0 : (int) min((uint64_t) kChunk, n_sites - first_site);
}

// Ordered product over the warp with the exponents; renormalised after the third and the last level
// only (entries stay below 2^63 in between).
__device__ __forceinline__ void warp_ordered_product_i(M2 &m, int &e) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const M2 o = shfl_down_m(m, off);
    const int oe = __shfl_down_sync(kFull, e, off);
    if ((lane & (2 * off - 1)) == 0) {
      m = matmul(m, o);
      e += oe;
      if (off == 4 || off == 16) e += renorm_i(m);
    }
  }
}

__global__ void __launch_bounds__(kScanThreads, 3)
estep_chunk_products(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
                     const double *__restrict__ alpha, const double *__restrict__ tile_dmax,
                     const double *__restrict__ tile_dsum, double4 *__restrict__ chunk_prod,
                     TileProd *__restrict__ tile_prod, uint64_t n_rows, uint32_t n_rows_valid, uint64_t n_sites,
                     uint64_t site_block, uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
  // rows vary fastest over the grid: the CTAs resident at any time share a few distance tiles in L2
  const uint32_t row = blockIdx.x % n_rows_valid, tile = blockIdx.x / n_rows_valid;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  const double F = indF[row], al = alpha[row];
  const int tier = kappa_tier(al, tile_dmax[tile]);
  stage_tile_begin(sm, emis + blocked_index(row, tile_first, n_rows, site_block), dist + tile_first);

  const double q0 = 1.0 - F, q1 = F;
  const double *r = sm.r + threadIdx.x * kChunk;
  const double *d = sm.d + threadIdx.x * kChunk;
  mbar_wait(&sm.bar, 0);

  // Padding sites are the identity because the context keeps r = 1 and d = 0 there.
  M2 m;
  int e;
  double ls;
  if (tier == kTierFast) products_chunk<kTierFast>(r, d, sm.tab, al, q0, q1, m, e, ls);
  else if (tier == kTierMid) products_chunk<kTierMid>(r, d, sm.tab, al, q0, q1, m, e, ls);
  else products_chunk<kTierSlow>(r, d, sm.tab, al, q0, q1, m, e, ls);
  // per-chunk product (direction only: the apply kernel is scale free)
  chunk_prod[((size_t) row * n_tiles + tile) * kScanThreads + threadIdx.x] = make_double4(m.a, m.b, m.c, m.d);

  // per-tile product with its scale, for the carry kernel and the log-likelihood
  __shared__ M2 sm_m[kScanThreads / 32];
  __shared__ int sm_e[kScanThreads / 32];
  __shared__ double sm_l[kScanThreads / 32];
  warp_ordered_product_i(m, e);
  if (tier == kTierSlow) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ls += __shfl_down_sync(kFull, ls, off);
  }
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
      ae += sm_e[w] + renorm_i(acc);
      al_sum += sm_l[w];
    }
    if (tier != kTierSlow) al_sum = -(al * tile_dsum[tile]);
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
      if (lane >= off) { m = matmul(o, m); e += oe + renorm_i(m); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(kFull, l, off);
    lsum += l;
    const M2 before = shfl_up_m(m, 1);
    double c0 = x0, c1 = x1;
    if (lane > 0) { c0 = fma(x0, before.a, x1 * before.c); c1 = fma(x0, before.b, x1 * before.d); renorm2_i(c0, c1); }
    if (t < n_tiles) fc[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 31); tot.b = __shfl_sync(kFull, m.b, 31);
    tot.c = __shfl_sync(kFull, m.c, 31); tot.d = __shfl_sync(kFull, m.d, 31);
    const int te = __shfl_sync(kFull, e, 31);
    const double y0 = fma(x0, tot.a, x1 * tot.c), y1 = fma(x0, tot.b, x1 * tot.d);
    x0 = y0; x1 = y1;
    ex += te + renorm2_i(x0, x1);
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
      if (lane + off < 32) { m = matmul(m, o); e += oe + renorm_i(m); }
    }
    const M2 after = shfl_down_m(m, 1);
    double c0 = b0, c1 = b1;
    if (lane < 31) { c0 = fma(after.a, b0, after.b * b1); c1 = fma(after.c, b0, after.d * b1); renorm2_i(c0, c1); }
    if (t < n_tiles) bc[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 0); tot.b = __shfl_sync(kFull, m.b, 0);
    tot.c = __shfl_sync(kFull, m.c, 0); tot.d = __shfl_sync(kFull, m.d, 0);
    const int te = __shfl_sync(kFull, e, 0);
    const double y0 = fma(tot.a, b0, tot.b * b1), y1 = fma(tot.c, b0, tot.d * b1);
    b0 = y0; b1 = y1;
    eb += te + renorm2_i(b0, b1);
  }
  const double lb = log(fma(q0, b0, q1 * b1)) + (double) eb * kLn2 + base;

  if (lane == 0) {
    ind_lkl[row] = lf;
    if (lf != lf || lb != lb) atomicOr(status, kFlagNaN);
    else if (fabs(lf - lb) > 1e-3) atomicOr(status, kFlagFwBw);   // EM.cpp:166
  }
}

// Carries of this thread's chunk from the tile carries and the tile's 128 chunk products: warp
// prefix/suffix scans (direction only, renormalised after the third and the last level: entries stay
// below 2^63 in between) plus a 4-way combine through shared memory.  Pure register / shuffle work: it
// runs while the tile's bulk copies are in flight.
__device__ __forceinline__ void chunk_carries(const double4 mine4, double2 tile_fwd, double2 tile_bwd, double &a0,
                                              double &a1, double &b0, double &b1, M2 *__restrict__ warp_tot) {
  constexpr int kWarps = kScanThreads / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  M2 mine; mine.a = mine4.x; mine.b = mine4.y; mine.c = mine4.z; mine.d = mine4.w;
  M2 pre = mine, suf = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const M2 o = shfl_up_m(pre, off);
    if (lane >= off) pre = matmul(o, pre);
    const M2 u = shfl_down_m(suf, off);
    if (lane + off < 32) suf = matmul(suf, u);
    if (off == 4 || off == 16) { renorm_i(pre); renorm_i(suf); }
  }
  if (lane == 31) warp_tot[warp] = pre;
  __syncthreads();
  a0 = tile_fwd.x; a1 = tile_fwd.y;
  for (int w = 0; w < warp; w++) {                 // warp-uniform trip count
    const M2 p = warp_tot[w];
    const double y0 = fma(a0, p.a, a1 * p.c), y1 = fma(a0, p.b, a1 * p.d);
    a0 = y0; a1 = y1;
    renorm2_i(a0, a1);
  }
  b0 = tile_bwd.x; b1 = tile_bwd.y;
  for (int w = kWarps - 1; w > warp; w--) {
    const M2 p = warp_tot[w];
    const double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
    b0 = y0; b1 = y1;
    renorm2_i(b0, b1);
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
  renorm2_i(a0, a1);
  renorm2_i(b0, b1);
}

__global__ void __launch_bounds__(kScanThreads, 3)
estep_chunk_apply(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
                  const double *__restrict__ alpha, const double *__restrict__ tile_dmax,
                  const double4 *__restrict__ chunk_prod, const double2 *__restrict__ fwd_carry,
                  const double2 *__restrict__ bwd_carry, double *__restrict__ post, PeerWindows peers,
                  int *__restrict__ status, uint64_t n_rows, uint32_t n_rows_valid, uint64_t n_sites,
                  uint64_t site_block, uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ApplySmem &as = *reinterpret_cast<ApplySmem *>(smem_raw);
  TileSmem &sm = as.tile;
  double *stash = &as.stash[0][threadIdx.x];
  // rows vary fastest over the grid: the CTAs resident at any time share a few distance tiles in L2
  const uint32_t row = blockIdx.x % n_rows_valid, tile = blockIdx.x / n_rows_valid;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  const size_t tile_at = blocked_index(row, tile_first, n_rows, site_block);
  // the loads the carries depend on go out first
  const size_t t_at = (size_t) row * n_tiles + tile;
  const double4 mine = chunk_prod[t_at * kScanThreads + threadIdx.x];
  const double2 tf = fwd_carry[t_at], tb = bwd_carry[t_at];
  const double F = indF[row], al = alpha[row];
  const double dmax = tile_dmax[tile];
  stage_tile_begin(sm, emis + tile_at, dist + tile_first);

  // while the tile is in flight: carries of this thread's chunk
  const double q0 = 1.0 - F, q1 = F;
  const int tier = kappa_tier(al, dmax);
  double cf0, cf1, cb0, cb1;
  __shared__ M2 warp_tot[kScanThreads / 32];
  chunk_carries(mine, tf, tb, cf0, cf1, cb0, cb1, warp_tot);

  double *r = sm.r + threadIdx.x * kChunk;     // becomes the posterior
  double *d = sm.d + threadIdx.x * kChunk;     // becomes kappa
  mbar_wait(&sm.bar, 0);
  bool bad;
  if (tier == kTierFast) bad = apply_chunk<kTierFast>(r, d, sm.tab, al, q0, q1, cf0, cf1, cb0, cb1, stash);
  else if (tier == kTierMid) bad = apply_chunk<kTierMid>(r, d, sm.tab, al, q0, q1, cf0, cf1, cb0, cb1, stash);
  else bad = apply_chunk<kTierSlow>(r, d, sm.tab, al, q0, q1, cf0, cf1, cb0, cb1, stash);
  if (bad) atomicOr(status, kFlagNaN);

  // padding sites of the last tile carry no posterior
  const int n_valid = valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, n_sites);
  for (int j = n_valid; j < kChunk; j++) r[j] = 0.0;

  fence_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (peers.direct) {
      // this tile belongs to site block b: store it straight into rank b's frequency-side window over
      // NVLink, source block = me - the posterior "all-to-all" happens inside this kernel
      // (nfh_get_posterior gathers from the same windows)
      const uint64_t b = tile_first / site_block;
      double *dst = peers.base[b] + ((uint64_t) peers.rank * n_rows + row) * site_block + (tile_first - b * site_block);
      tma_store_1d(dst, sm.r, kTileBytes);
    } else {
      tma_store_1d(post + tile_at, sm.r, kTileBytes);     // individual-major plane (nfh_get_posterior, frequency EM)
    }
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------
// Single-launch E-step.  Persistent CTAs take (phase, individual, tile) items from one ticket counter in
// the order of nfh_schedule.h: a P item is what estep_chunk_products does for one tile, an A item what
// estep_chunk_apply does, and the carries of an individual (estep_tile_carries) are computed by the CTA
// that completes the individual's last P item.  Every CTA keeps the single tile buffer of the two-launch
// kernels (three CTAs per SM are in different phases of their items, which is what overlaps copies and
// arithmetic); the next item's tile is requested as soon as the buffer is free, before the bookkeeping of
// the current item.
// ---------------------------------------------------------------------------
// An item as the CTA's thread 0 hands it to the other threads: decoded once, with the individual's
// parameters and the tile's largest distance already fetched.
struct ItemSlot {
  uint32_t ticket, apply, row, tile;
  double F, al, dmax;
  uint32_t ready, pad_;
};
constexpr uint32_t kTicketBatch = 4;

struct FusedSmem {                      // <= 76,800 bytes: three CTAs per SM
  alignas(128) double r[kTile];         // emission ratio -> posterior
  alignas(128) double d[kTile];         // distance -> kappa
  double stash[kStash][kScanThreads];   // see ApplySmem; scratch of row_carries before the item's posteriors start
  double tab[64];
  // per-warp results of an item, double-buffered by item parity: the thread that finishes a P item reads them
  // while the other warps may already be in the next item (P: warp products; A: warp totals of chunk_carries)
  M2 wm[2][kScanThreads / 32];
  int we[2][kScanThreads / 32];
  double wl[2][kScanThreads / 32];
  ItemSlot item[2];
  alignas(8) uint64_t bar;
  int last;
};
static_assert(sizeof(FusedSmem) <= 76800, "three CTAs of estep_fused must fit one SM (228 KB - 3 x 1 KB reserved)");

struct CarryScratch {                   // row_carries, aliased onto FusedSmem::stash
  M2 wm[kScanThreads / 32];
  long long we[kScanThreads / 32];
  double wl[kScanThreads / 32];
  double res[2];
};

__device__ __forceinline__ void load_tile_prod(const TileProd *p, M2 &m, int &e, double &l) {
  // written by other SMs during this launch: read through L2
  const double2 *q = reinterpret_cast<const double2 *>(p);
  const double2 ab = __ldcg(q), cd = __ldcg(q + 1), el = __ldcg(q + 2);
  m.a = ab.x; m.b = ab.y; m.c = cd.x; m.d = cd.y; e = (int) el.x; l = el.y;
}

// Carries of one individual by a whole CTA (the job of estep_tile_carries, which gives one warp per
// individual): the tiles are cut into one segment per warp; pass 1 multiplies each segment up, pass 2 scans
// each segment from the vector that enters it.  Log-likelihood both ways as in EM.cpp:166.
__device__ __noinline__ void row_carries(const EstepArgs &a, uint32_t row, CarryScratch &fs) {
  constexpr int kWarps = kScanThreads / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_tiles = a.n_tiles;
  const uint32_t n_groups = (n_tiles + 31) / 32;
  const uint32_t gper = (n_groups + kWarps - 1) / kWarps;
  const uint32_t g0 = min((uint32_t) warp * gper, n_groups), g1 = min(g0 + gper, n_groups);
  const TileProd *tp = a.tile_prod + (size_t) row * n_tiles;
  double2 *fc = a.fwd_carry + (size_t) row * n_tiles;
  double2 *bc = a.bwd_carry + (size_t) row * n_tiles;
  const double F = a.indF[row];
  const double q0 = 1.0 - F, q1 = F;

  // ---- pass 1: product, exponent and log-scale of this warp's segment
  {
    M2 acc = identity2();
    long long ae = 0;
    double al = 0.0;
    for (uint32_t g = g0; g < g1; g++) {
      const uint32_t t = g * 32 + lane;
      M2 m = identity2();
      int e = 0;
      double l = 0.0;
      if (t < n_tiles) load_tile_prod(tp + t, m, e, l);
      warp_ordered_product_i(m, e);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(kFull, l, off);
      al += l;
      if (lane == 0) { acc = matmul(acc, m); ae += e + renorm_i(acc); }
    }
    if (lane == 0) { fs.wm[warp] = acc; fs.we[warp] = ae; fs.wl[warp] = al; }
  }
  __syncthreads();

  // ---- vectors entering this warp's segment from the left and from the right
  double x0 = q0, x1 = q1, b0 = 1.0, b1 = 1.0, lsum = 0.0;
  long long ex = 0, eb = 0;
  for (int w = 0; w < kWarps; w++) lsum += fs.wl[w];
  for (int w = 0; w < warp; w++) {
    const M2 p = fs.wm[w];
    const double y0 = fma(x0, p.a, x1 * p.c), y1 = fma(x0, p.b, x1 * p.d);
    x0 = y0; x1 = y1;
    ex += fs.we[w] + renorm2_i(x0, x1);
  }
  for (int w = kWarps - 1; w > warp; w--) {
    const M2 p = fs.wm[w];
    const double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
    b0 = y0; b1 = y1;
    eb += fs.we[w] + renorm2_i(b0, b1);
  }

  // ---- pass 2, forwards: carry into every tile of the segment
  for (uint32_t g = g0; g < g1; g++) {
    const uint32_t t = g * 32 + lane;
    M2 m = identity2();
    int e = 0;
    double l;
    if (t < n_tiles) load_tile_prod(tp + t, m, e, l);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const M2 o = shfl_up_m(m, off);
      const int oe = __shfl_up_sync(kFull, e, off);
      if (lane >= off) { m = matmul(o, m); e += oe + renorm_i(m); }
    }
    const M2 before = shfl_up_m(m, 1);
    double c0 = x0, c1 = x1;
    if (lane > 0) { c0 = fma(x0, before.a, x1 * before.c); c1 = fma(x0, before.b, x1 * before.d); renorm2_i(c0, c1); }
    if (t < n_tiles) fc[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 31); tot.b = __shfl_sync(kFull, m.b, 31);
    tot.c = __shfl_sync(kFull, m.c, 31); tot.d = __shfl_sync(kFull, m.d, 31);
    const int te = __shfl_sync(kFull, e, 31);
    const double y0 = fma(x0, tot.a, x1 * tot.c), y1 = fma(x0, tot.b, x1 * tot.d);
    x0 = y0; x1 = y1;
    ex += te + renorm2_i(x0, x1);
  }
  // ---- pass 2, backwards: carry out of every tile of the segment
  for (uint32_t g = g1; g-- > g0;) {
    const uint32_t t = g * 32 + lane;
    M2 m = identity2();
    int e = 0;
    double l;
    if (t < n_tiles) load_tile_prod(tp + t, m, e, l);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const M2 o = shfl_down_m(m, off);
      const int oe = __shfl_down_sync(kFull, e, off);
      if (lane + off < 32) { m = matmul(m, o); e += oe + renorm_i(m); }
    }
    const M2 after = shfl_down_m(m, 1);
    double c0 = b0, c1 = b1;
    if (lane < 31) { c0 = fma(after.a, b0, after.b * b1); c1 = fma(after.c, b0, after.d * b1); renorm2_i(c0, c1); }
    if (t < n_tiles) bc[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 0); tot.b = __shfl_sync(kFull, m.b, 0);
    tot.c = __shfl_sync(kFull, m.c, 0); tot.d = __shfl_sync(kFull, m.d, 0);
    const int te = __shfl_sync(kFull, e, 0);
    const double y0 = fma(tot.a, b0, tot.b * b1), y1 = fma(tot.c, b0, tot.d * b1);
    b0 = y0; b1 = y1;
    eb += te + renorm2_i(b0, b1);
  }
  // the last warp now holds the whole forward vector, warp 0 the whole backward vector
  if (warp == kWarps - 1 && lane == 0) fs.res[0] = log(x0 + x1) + (double) ex * kLn2;
  if (warp == 0 && lane == 0) fs.res[1] = log(fma(q0, b0, q1 * b1)) + (double) eb * kLn2;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double base = lsum + a.loge0_sum[row];
    const double lf = fs.res[0] + base, lb = fs.res[1] + base;
    a.ind_lkl[row] = lf;
    if (lf != lf || lb != lb) atomicOr(a.status, kFlagNaN);
    else if (fabs(lf - lb) > 1e-3) atomicOr(a.status, kFlagFwBw);   // EM.cpp:166
  }
}

struct FusedArgs {
  EstepArgs a;
  EstepSchedule s;
  unsigned long long *ticket;        // one counter, never reset: this launch owns [ticket_base, ticket_base + total + grid)
  unsigned long long ticket_base;
  unsigned long long *row_done;      // per individual: P items completed, over all launches
  unsigned long long done_target;    // value of row_done[] once this launch's products are in = launch number * n_tiles
  unsigned *row_claim;               // per individual: launch number whose carries somebody has taken on
  unsigned *row_ready;               // per individual: launch number whose carries are in place
  unsigned epoch;                    // launch number (from 1)
  int l2_hints;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// count one more finished item; everything this thread (and, through a preceding bar.sync, its CTA) wrote
// before is visible to whoever observes the count.  No return value: the thread does not wait for the
// round trip.
__device__ __forceinline__ void red_release_inc(unsigned long long *p) {
  asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

__global__ void __launch_bounds__(kScanThreads, 3)
estep_fused(const __grid_constant__ FusedArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedSmem &fs = *reinterpret_cast<FusedSmem *>(smem_raw);
  FusedSmem &sm = fs;
  double *stash = &fs.stash[0][threadIdx.x];
  const EstepArgs &a = A.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t total = A.s.total;

  // ---- thread 0 only: L2 policies, ticket fetch, decoding, tile requests
  // Tickets are taken kTicketBatch at a time and the next batch is asked for while the current one is in use,
  // so that the round trip of the atomic is never waited for.  A CTA stops asking once it is handed a batch
  // beyond the end: every launch advances the counter by kTicketBatch * (ceil(total / kTicketBatch) + gridDim.x).
  uint64_t pol_keep = 0, pol_once = 0;
  uint32_t cur = total, cur_end = total, ahead = total;   // current batch [cur, cur_end), first ticket of the next one
  uint32_t nxt = 0;                            // the ticket after the current item
  ItemSlot nxt_item;                           // ... decoded (valid when nxt < total)
  auto fetch_batch = [&]() -> uint32_t {
    const unsigned long long t = atomicAdd(A.ticket, (unsigned long long) kTicketBatch) - A.ticket_base;
    return t < (unsigned long long) total ? (uint32_t) t : total;
  };
  auto fetch = [&]() -> uint32_t {
    if (cur == cur_end) {
      cur = ahead;
      cur_end = min(cur + kTicketBatch, total);
      if (cur < total) ahead = fetch_batch();
    }
    return cur < cur_end ? cur++ : total;
  };
  auto describe = [&](uint32_t t, ItemSlot &s) {
    s.ticket = t;
    if (t >= total) return;
    const EstepItem it = decode_ticket(A.s, t);
    s.apply = it.apply; s.row = it.row; s.tile = it.tile;
    s.F = a.indF[it.row]; s.al = a.alpha[it.row]; s.dmax = a.tile_dmax[it.tile];
    // posteriors: are the individual's carries in place already?  (asked early; the item itself waits if not)
    s.ready = it.apply ? (ld_acquire(A.row_ready + it.row) == A.epoch) : 1u;
  };
  const uint32_t tiles_per_block = (uint32_t) (a.site_block / kTile);     // a tile never straddles a site block
  auto tile_offset = [&](uint32_t row, uint32_t tile) -> size_t {         // blocked_index() with 32-bit divisions
    const uint32_t blk = tile / tiles_per_block;
    return ((size_t) blk * a.n_rows + row) * a.site_block + (size_t) (tile - blk * tiles_per_block) * kTile;
  };
  auto request_tile = [&](const ItemSlot &s) {
    const uint64_t tile_first = (uint64_t) s.tile * kTile;
    const double *src = a.emis + tile_offset(s.row, s.tile);
    mbar_arrive_expect_tx(&sm.bar, 2 * kTileBytes);
    if (A.l2_hints) tma_load_1d_hint(sm.r, src, kTileBytes, &sm.bar, s.apply ? pol_once : pol_keep);
    else tma_load_1d(sm.r, src, kTileBytes, &sm.bar);
    tma_load_1d(sm.d, a.dist + tile_first, kTileBytes, &sm.bar);
  };
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
    pol_keep = l2_policy_evict_last();
    pol_once = l2_policy_evict_first();
    ahead = fetch_batch();
    ItemSlot first;
    describe(fetch(), first);
    fs.item[0] = first;
    nxt = total;
    if (first.ticket < total) { request_tile(first); nxt = fetch(); }
  }
  load_exp_table(sm.tab);
  __syncthreads();

  for (unsigned n = 0;; n++) {
    const ItemSlot item = fs.item[n & 1];
    if (item.ticket >= total) break;
    const uint32_t row = item.row, tile = item.tile;
    const uint64_t tile_first = (uint64_t) tile * kTile;
    const size_t t_at = (size_t) row * a.n_tiles + tile;
    const double F = item.F, al = item.al;
    const double q0 = 1.0 - F, q1 = F;
    const int tier = kappa_tier(al, item.dmax);
    double *r = sm.r + threadIdx.x * kChunk;
    double *d = sm.d + threadIdx.x * kChunk;
    // the next item: described while this one computes (the loads are in flight behind the arithmetic)
    if (threadIdx.x == 0) describe(nxt, nxt_item);

    if (!item.apply) {
      // ------------------------------------------------------------------ P item
      mbar_wait(&sm.bar, n & 1);
      M2 m;
      int e;
      double ls;
      if (tier == kTierFast) products_chunk<kTierFast>(r, d, sm.tab, al, q0, q1, m, e, ls);
      else if (tier == kTierMid) products_chunk<kTierMid>(r, d, sm.tab, al, q0, q1, m, e, ls);
      else products_chunk<kTierSlow>(r, d, sm.tab, al, q0, q1, m, e, ls);
      a.chunk_prod[t_at * kScanThreads + threadIdx.x] = make_double4(m.a, m.b, m.c, m.d);
      warp_ordered_product_i(m, e);
      if (tier == kTierSlow) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ls += __shfl_down_sync(kFull, ls, off);
      }
      if (lane == 0) { fs.wm[n & 1][warp] = m; fs.we[n & 1][warp] = e; fs.wl[n & 1][warp] = ls; }
      if (threadIdx.x == 0) fs.item[(n + 1) & 1] = nxt_item;
      __syncthreads();                       // the tile buffer is free; warp products and chunk products are out
      if (threadIdx.x == 0 && nxt < total) { request_tile(nxt_item); nxt = fetch(); }
      if (threadIdx.x == 32) {
        // tile product and completion count by a thread of another warp than the one that feeds the CTA; the
        // other warps may be in their next item meanwhile (it uses the other half of wm / we / wl; this half is
        // written again two items on, behind a barrier this warp has to pass first)
        M2 acc = fs.wm[n & 1][0];
        long long ae = fs.we[n & 1][0];
        double al_sum = fs.wl[n & 1][0];
#pragma unroll
        for (int w = 1; w < kScanThreads / 32; w++) {
          acc = matmul(acc, fs.wm[n & 1][w]);
          ae += fs.we[n & 1][w] + renorm_i(acc);
          al_sum += fs.wl[n & 1][w];
        }
        if (tier != kTierSlow) al_sum = -(al * a.tile_dsum[tile]);
        TileProd out;
        out.a = acc.a; out.b = acc.b; out.c = acc.c; out.d = acc.d; out.e = (double) ae; out.l = al_sum;
        a.tile_prod[t_at] = out;
        red_release_inc(A.row_done + row);   // orders the CTA's chunk products and this tile product before the count
      }
    } else {
      // ------------------------------------------------------------------ A item
      if (!item.ready) {                     // uniform over the CTA
        if (threadIdx.x == 0) {
          int scan = 0;
          if (ld_acquire(A.row_ready + row) != A.epoch) {
            while (ld_acquire(A.row_done + row) < A.done_target) __nanosleep(40);   // products of all tiles are in
            // whoever gets here first computes the individual's carries with its whole CTA
            if (atomicCAS(A.row_claim + row, A.epoch - 1, A.epoch) == A.epoch - 1) scan = 1;
            else while (ld_acquire(A.row_ready + row) != A.epoch) __nanosleep(40);
          }
          fs.last = scan;
        }
        __syncthreads();
        if (fs.last) {
          row_carries(a, row, *reinterpret_cast<CarryScratch *>(&fs.stash[0][0]));
          __syncthreads();                   // all carries of the individual written
          if (threadIdx.x == 0) { __threadfence(); st_release(A.row_ready + row, A.epoch); }
        }
      }
      const double2 *cp2 = reinterpret_cast<const double2 *>(a.chunk_prod + t_at * kScanThreads + threadIdx.x);
      const double2 m_ab = __ldcg(cp2), m_cd = __ldcg(cp2 + 1);
      const double4 mine = make_double4(m_ab.x, m_ab.y, m_cd.x, m_cd.y);
      const double2 tf = __ldcg(a.fwd_carry + t_at), tb = __ldcg(a.bwd_carry + t_at);
      double cf0, cf1, cb0, cb1;
      chunk_carries(mine, tf, tb, cf0, cf1, cb0, cb1, fs.wm[n & 1]);
      mbar_wait(&sm.bar, n & 1);
      bool bad;
      if (tier == kTierFast) bad = apply_chunk<kTierFast>(r, d, sm.tab, al, q0, q1, cf0, cf1, cb0, cb1, stash);
      else if (tier == kTierMid) bad = apply_chunk<kTierMid>(r, d, sm.tab, al, q0, q1, cf0, cf1, cb0, cb1, stash);
      else bad = apply_chunk<kTierSlow>(r, d, sm.tab, al, q0, q1, cf0, cf1, cb0, cb1, stash);
      if (bad) atomicOr(a.status, kFlagNaN);
      const int n_valid = valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, a.n_sites);
      for (int j = n_valid; j < kChunk; j++) r[j] = 0.0;
      if (threadIdx.x == 0) fs.item[(n + 1) & 1] = nxt_item;
      fence_async_shared();
      __syncthreads();
      if (threadIdx.x == 0) {
        double *dst;
        if (a.post_peers.direct) {
          const uint32_t b = tile / tiles_per_block;
          dst = a.post_peers.base[b] + ((uint64_t) a.post_peers.rank * a.n_rows + row) * a.site_block +
                (size_t) (tile - b * tiles_per_block) * kTile;
        } else {
          dst = a.post + tile_offset(row, tile);
        }
        if (A.l2_hints) tma_store_1d_hint(dst, sm.r, kTileBytes, pol_once);
        else tma_store_1d(dst, sm.r, kTileBytes);
        tma_store_wait_read();               // the tile buffer is free again
        if (nxt < total) { request_tile(nxt_item); nxt = fetch(); }
      }
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
// The dynamic shared-memory attribute belongs to the (function, device) pair; set it once per device.
static void set_smem_attrs() {
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && done[dev]) return;
  cudaFuncSetAttribute(estep_chunk_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(TileSmem));
  cudaFuncSetAttribute(estep_chunk_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(ApplySmem));
  if (dev >= 0 && dev < 64) done[dev] = true;
}

void launch_estep_tail(const EstepArgs &a, cudaStream_t st) {
  set_smem_attrs();
  estep_tile_carries<<<(unsigned) ((a.n_rows_valid + 3) / 4), 128, 0, st>>>(
      a.tile_prod, a.indF, a.loge0_sum, a.fwd_carry, a.bwd_carry, a.ind_lkl, a.status, (uint32_t) a.n_rows_valid,
      a.n_tiles);
  const unsigned grid = (unsigned) (a.n_rows_valid * a.n_tiles);
  estep_chunk_apply<<<grid, kScanThreads, sizeof(ApplySmem), st>>>(
      a.emis, a.dist, a.indF, a.alpha, a.tile_dmax, a.chunk_prod, a.fwd_carry, a.bwd_carry, a.post, a.post_peers,
      a.status, a.n_rows, (uint32_t) a.n_rows_valid, a.n_sites, a.site_block, a.n_tiles);
}

static long env_long(const char *name, long fallback) {
  const char *v = getenv(name);
  return v && *v ? atol(v) : fallback;
}

// Wave policy of the single-launch E-step (nfh_schedule.h).  Short sequences: as many individuals per wave as
// fit NFH_ESTEP_WAVE_MB (default 40) of ratio plane, so that the second sweep of a wave reads L2; the bulk
// copies then carry eviction hints.  Long sequences (two individuals do not fit): 16 individuals per wave, which
// still shares every distance tile between 16 CTAs and alternates product and posterior items on every SM.
static bool launch_estep_fused(const EstepArgs &a, cudaStream_t st) {
  EstepFusedState *f = a.fused;
  // opt-in: measured slower than the three launches ("single-launch E-step")
  if (!f || !f->d_ticket || env_long("NFH_ESTEP_FUSED", 0) == 0 || a.n_rows_valid == 0) return false;
  static int ctas_per_sm[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return false;
  if (!ctas_per_sm[dev]) {
    if (cudaFuncSetAttribute(estep_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(FusedSmem)) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[dev], estep_fused, kScanThreads, sizeof(FusedSmem)) != cudaSuccess ||
        ctas_per_sm[dev] < 1) {
      cudaGetLastError();
      ctas_per_sm[dev] = -1;
    }
  }
  if (ctas_per_sm[dev] < 1) return false;
  const double row_mb = (double) a.n_tiles * kTileBytes / 1048576.0;
  const long budget_mb = env_long("NFH_ESTEP_WAVE_MB", 40);
  long wave_rows = (long) ((double) budget_mb / row_mb);
  const bool l2_mode = wave_rows >= 2;
  if (!l2_mode) wave_rows = 16;
  wave_rows = env_long("NFH_ESTEP_WAVE_ROWS", wave_rows);
  FusedArgs A;
  A.a = a;
  const unsigned long long total = 2ull * a.n_rows_valid * a.n_tiles;
  const unsigned grid = (unsigned) std::min<unsigned long long>(total, (unsigned long long) a.sm_count * ctas_per_sm[dev]);
  A.s = make_schedule((uint32_t) a.n_rows_valid, a.n_tiles, (uint32_t) std::max(1l, wave_rows),
                      (uint32_t) env_long("NFH_ESTEP_LOOKAHEAD", 2l * grid));
  if (A.s.total == 0) return false;            // more than 2^32 items: three-launch path
  f->epoch++;
  A.ticket = f->d_ticket; A.ticket_base = f->ticket_base;
  A.row_done = f->d_row_done; A.done_target = (unsigned long long) f->epoch * a.n_tiles;
  A.row_claim = f->d_row_claim; A.row_ready = f->d_row_ready; A.epoch = f->epoch;
  A.l2_hints = (int) env_long("NFH_ESTEP_HINTS", l2_mode ? 1 : 0);
  f->ticket_base += (unsigned long long) kTicketBatch * ((total + kTicketBatch - 1) / kTicketBatch + grid);
  estep_fused<<<grid, kScanThreads, sizeof(FusedSmem), st>>>(A);
  return true;
}

void launch_estep(const EstepArgs &a, cudaStream_t st) {
  if (launch_estep_fused(a, st)) return;
  set_smem_attrs();
  const unsigned grid = (unsigned) (a.n_rows_valid * a.n_tiles);
  estep_chunk_products<<<grid, kScanThreads, sizeof(TileSmem), st>>>(
      a.emis, a.dist, a.indF, a.alpha, a.tile_dmax, a.tile_dsum, a.chunk_prod, a.tile_prod, a.n_rows,
      (uint32_t) a.n_rows_valid, a.n_sites, a.site_block, a.n_tiles);
  launch_estep_tail(a, st);
}

}  // namespace nfh

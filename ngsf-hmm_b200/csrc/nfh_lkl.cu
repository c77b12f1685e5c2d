// nfh_lkl.cu - the batched forward-only objective lkl() (EM.cpp:449-464) that findmax_bfgs asks for
// (bfgs.cpp:108-121): up to kMaxPoints (F, alpha) points of one individual - the centre and the
// finite-difference points of Yanggradient (bfgs.cpp:22-43) - share one staged read of its emissions.
#include "nfh_device.cuh"
#include "nfh_kernels.h"

namespace nfh {

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

// FP64 instructions per site of a set of points: per distinct alpha one kappa (6 / 12 / 14 by tier, 12 used
// for the split), per point the 2x2 update (10)
__host__ __device__ constexpr int lkl_cost(int n_same, int n_other) {
  return 12 * ((n_same > 0 ? 1 : 0) + n_other) + 10 * (n_same + n_other);
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
template <int TIER, int NS, int NA>
__device__ __forceinline__ void lkl_chunk_run(const LklGroup &g, int first, LklSmem &sm, int t,
                                              double4 *__restrict__ emit_chunks) {
  constexpr int NP = NS + NA;
  constexpr int kBody = TIER == kTierSlow ? 6 : 8;
  const double *r = sm.r + t * kChunk;
  const double *d = sm.d + t * kChunk;
  M2 m[NP];
  int e[NP];
  double ls[1 + NA];          // scale sums (slow tier only): one for the shared alpha, one per extra alpha
  double q1[NP], q0[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) { m[p] = identity2(); e[p] = 0; q1[p] = g.F[first + p]; q0[p] = 1.0 - g.F[first + p]; }
#pragma unroll
  for (int a = 0; a <= NA; a++) ls[a] = 0.0;

  auto site = [&](int j) {
    const double dj = d[j];
    const double rj = r[j];                          // padding: r = 1, d = 0 -> identity
    if (NS > 0) {
      const double ks = tier_kappa<TIER>(g.alpha[first] * dj, sm.tab, ls[0]);
#pragma unroll
      for (int p = 0; p < NS; p++) apply_site(m[p], ks * q0[p], ks * q1[p], rj);
    }
#pragma unroll
    for (int a = 0; a < NA; a++) {
      const double ka = tier_kappa<TIER>(g.alpha[first + NS + a] * dj, sm.tab, ls[1 + a]);
      apply_site(m[NS + a], ka * q0[NS + a], ka * q1[NS + a], rj);
    }
  };
#pragma unroll 1
  for (int j0 = 0; j0 + kBody <= kChunk; j0 += kBody) {
#pragma unroll
    for (int i = 0; i < kBody; i++) site(j0 + i);
#pragma unroll
    for (int p = 0; p < NP; p++) e[p] += renorm_i(m[p]);
  }
#pragma unroll
  for (int j = (kChunk / kBody) * kBody; j < kChunk; j++) site(j);

  const int warp = (t >> 5), lane = t & 31;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    e[p] += renorm_i(m[p]);
    // the group's first point doubles as the E-step's forward product of this chunk (direction only)
    if (p == 0 && first == 0 && emit_chunks) emit_chunks[t] = make_double4(m[0].a, m[0].b, m[0].c, m[0].d);
    warp_ordered_product(m[p], e[p]);
    double l = 0.0;
    if (TIER == kTierSlow) {
      l = ls[p < NS ? 0 : 1 + (p - NS)];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) l += __shfl_down_sync(kFull, l, off);
    }
    if (lane == 0) { sm.m[first + p][warp] = m[p]; sm.e[first + p][warp] = e[p]; sm.l[first + p][warp] = l; }
  }
}

template <int TIER, int NS, int NA>
__device__ __forceinline__ void lkl_tile_halves(const LklGroup &g, LklSmem &sm, int half, int t,
                                                double4 *__restrict__ emit_chunks) {
  constexpr int k = lkl_split(NS, NA);
  constexpr int NSa = k < NS ? k : NS, NAa = k - NSa, NSb = NS - NSa, NAb = NA - NAa;
  if (half == 0) {
    lkl_chunk_run<TIER, NSa, NAa>(g, 0, sm, t, emit_chunks);
  } else {
    if constexpr (NSb + NAb > 0) lkl_chunk_run<TIER, NSb, NAb>(g, k, sm, t, nullptr);
  }
}

template <int TIER>
__device__ __forceinline__ void lkl_dispatch(const LklGroup &g, LklSmem &sm, int half, int t,
                                             double4 *__restrict__ emit_chunks) {
#define NFH_LKL(ns, na) case (ns) * 8 + (na): lkl_tile_halves<TIER, ns, na>(g, sm, half, t, emit_chunks); break;
  switch (g.n_same * 8 + (g.npts - g.n_same)) {
    NFH_LKL(1, 0) NFH_LKL(1, 1) NFH_LKL(1, 2) NFH_LKL(1, 3) NFH_LKL(1, 4)
    NFH_LKL(2, 0) NFH_LKL(2, 1) NFH_LKL(2, 2) NFH_LKL(2, 3)
    NFH_LKL(3, 0) NFH_LKL(3, 1) NFH_LKL(3, 2)
    NFH_LKL(4, 0) NFH_LKL(4, 1)
    NFH_LKL(5, 0)
    default: break;
  }
#undef NFH_LKL
}

// One CTA per (tile, group): the tile of emission ratios and distances is staged once by TMA and read by
// both halves of the CTA (2 x 128 threads, each thread kChunk sites), which doubles the warps an SM can
// hold for the same shared memory (the tile, not registers, limits occupancy: 3 CTAs per SM).
__global__ void __launch_bounds__(kLklThreads)
lkl_tile_products(const double *__restrict__ emis, const double *__restrict__ dist,
                  const double *__restrict__ tile_dmax, const double *__restrict__ tile_dsum,
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
    mbar_arrive_expect_tx(&sm.bar, 2 * kTileBytes);
    tma_load_1d(sm.r, emis + blocked_index((uint64_t) g.ind, tile_first, n_rows, site_block), kTileBytes, &sm.bar);
    tma_load_1d(sm.d, dist + tile_first, kTileBytes, &sm.bar);
  }
  load_exp_table(sm.tab);
  __syncthreads();
  double amax = g.alpha[0];
  for (int p = 1; p < g.npts; p++) amax = fmax(amax, g.alpha[p]);
  const int tier = kappa_tier(amax, tile_dmax[tile]);
  const int half = threadIdx.x / kScanThreads, t = threadIdx.x % kScanThreads;
  double4 *emit_chunks =
      emit_chunk_prod ? emit_chunk_prod + ((size_t) g.ind * n_tiles + tile) * kScanThreads : nullptr;
  mbar_wait(&sm.bar, 0);

  if (tier == kTierFast) lkl_dispatch<kTierFast>(g, sm, half, t, emit_chunks);
  else if (tier == kTierMid) lkl_dispatch<kTierMid>(g, sm, half, t, emit_chunks);
  else lkl_dispatch<kTierSlow>(g, sm, half, t, emit_chunks);
  __syncthreads();
  if ((int) threadIdx.x < g.npts) {
    const int p = threadIdx.x;
    M2 acc = sm.m[p][0];
    int ae = sm.e[p][0];
    double al_sum = sm.l[p][0];
    for (int w = 1; w < kScanThreads / 32; w++) {
      acc = matmul(acc, sm.m[p][w]);
      ae += sm.e[p][w] + renorm_i(acc);
      al_sum += sm.l[p][w];
    }
    if (tier != kTierSlow) al_sum = -(g.alpha[p] * tile_dsum[tile]);
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
    e += (int) q.e + renorm_i(m);
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

void launch_lkl_batch(const LklArgs &a, cudaStream_t st) {
  static bool done[64] = {false};     // the attribute belongs to the (function, device) pair
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !done[dev]) {
    cudaFuncSetAttribute(lkl_tile_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LklSmem));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  dim3 grid(a.n_tiles, a.n_groups);
  lkl_tile_products<<<grid, kLklThreads, sizeof(LklSmem), st>>>(a.emis, a.dist, a.tile_dmax, a.tile_dsum, a.groups,
                                                                 a.tile_prod, a.n_rows, a.n_sites, a.site_block,
                                                                 a.n_tiles, a.emit_chunk_prod, a.emit_tile_prod);
  const unsigned warps = a.n_groups * kMaxPoints;
  lkl_finish<<<(warps + 3) / 4, 128, 0, st>>>(a.tile_prod, a.groups, a.loge0_sum, a.neg_lkl, a.n_groups, a.n_tiles);
}

}  // namespace nfh

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
//   estep_chunk_scan     : per individual, scan over the chunk products: forward
//                          carry into and backward carry out of every chunk, and
//                          the log-likelihood computed both ways.
//   estep_chunk_apply    : every thread re-reads its 33 sites (TMA-staged), runs
//                          the forward and the backward vector recursion from its
//                          carries and writes the clamped IBD posterior (TMA store).
// HBM traffic per individual-site: 8 B + 8 B of emission ratio read, 8 B of
// posterior written, ~2.4 B of chunk products/carries; site distances come
// from L2.
#include "nfh_device.cuh"
#include "nfh_kernels.h"

namespace nfh {

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
                     const double *__restrict__ alpha, ChunkProd *__restrict__ chunk_prod, uint64_t n_rows,
                     uint64_t n_sites, uint64_t site_block, uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  stage_tile(sm, emis + blocked_index(row, tile_first, n_rows, site_block), dist + tile_first);

  const int n_valid = valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, n_sites);
  const double F = indF[row], al = alpha[row];
  const double q0 = 1.0 - F, q1 = F;
  const double *r = sm.r + threadIdx.x * kChunk;
  const double *d = sm.d + threadIdx.x * kChunk;

  M2 m = identity2();
  int e = 0;
  double ls = 0.0;
#pragma unroll
  for (int j = 0; j < kChunk; j++) {
    if (j < n_valid) {
      const double kap = site_kappa(al * d[j], sm.tab, ls);
      apply_site(m, kap * q0, kap * q1, r[j]);
    }
    if (j % 6 == 5) e += renorm(m);
  }
  e += renorm(m);
  ChunkProd out;
  out.a = m.a; out.b = m.b; out.c = m.c; out.d = m.d; out.e = (double) e; out.l = ls;
  chunk_prod[((size_t) row * n_tiles + tile) * kScanThreads + threadIdx.x] = out;
}

// vector-matrix and matrix-vector steps with renormalisation
__device__ __forceinline__ int step_fwd(double &x0, double &x1, const ChunkProd &p) {
  const double y0 = fma(x0, p.a, x1 * p.c), y1 = fma(x0, p.b, x1 * p.d);
  x0 = y0; x1 = y1;
  return renorm2(x0, x1);
}
__device__ __forceinline__ int step_bwd(double &b0, double &b1, const ChunkProd &p) {
  const double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
  b0 = y0; b1 = y1;
  return renorm2(b0, b1);
}

// One CTA per individual.  Thread t owns a contiguous run of chunk products.
//   pass 1: product of the run (as a 2x2 with exponent) -> shared memory
//   middle: two threads chain the 256 run products, forwards and backwards
//   pass 2: every thread walks its run again, writing the carry of each chunk
constexpr int kCarryThreads = 256;

__global__ void __launch_bounds__(kCarryThreads)
estep_chunk_scan(const ChunkProd *__restrict__ chunk_prod, const double *__restrict__ indF,
                 const double *__restrict__ loge0_sum, double2 *__restrict__ fwd_carry,
                 double2 *__restrict__ bwd_carry, double *__restrict__ ind_lkl, int *__restrict__ status,
                 uint32_t n_chunks) {
  const uint32_t row = blockIdx.x;
  const double F = indF[row];
  const double q0 = 1.0 - F, q1 = F;
  const ChunkProd *cp = chunk_prod + (size_t) row * n_chunks;
  double2 *fc = fwd_carry + (size_t) row * n_chunks;
  double2 *bc = bwd_carry + (size_t) row * n_chunks;

  const uint32_t per = (n_chunks + kCarryThreads - 1) / kCarryThreads;
  const uint32_t lo = min(n_chunks, threadIdx.x * per), hi = min(n_chunks, lo + per);

  __shared__ ChunkProd run[kCarryThreads];
  __shared__ double2 run_in[kCarryThreads], run_out[kCarryThreads];
  __shared__ double lkl_both[2];

  {
    M2 m = identity2();
    long long e = 0;
    double l = 0.0;
    for (uint32_t c = lo; c < hi; c++) {
      const ChunkProd p = cp[c];
      M2 o; o.a = p.a; o.b = p.b; o.c = p.c; o.d = p.d;
      m = matmul(m, o);
      e += (long long) p.e + renorm(m);
      l += p.l;
    }
    ChunkProd t;
    t.a = m.a; t.b = m.b; t.c = m.c; t.d = m.d; t.e = (double) e; t.l = l;
    run[threadIdx.x] = t;
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    double x0 = q0, x1 = q1, e = 0.0, l = 0.0;
    for (int t = 0; t < kCarryThreads; t++) {
      run_in[t] = make_double2(x0, x1);
      const ChunkProd p = run[t];
      e += p.e + step_fwd(x0, x1, p);
      l += p.l;
    }
    const double lf = log(x0 + x1) + e * kLn2 + l + loge0_sum[row];
    ind_lkl[row] = lf;
    lkl_both[0] = lf;
  } else if (threadIdx.x == 32) {
    double b0 = 1.0, b1 = 1.0, e = 0.0, l = 0.0;
    for (int t = kCarryThreads - 1; t >= 0; t--) {
      run_out[t] = make_double2(b0, b1);
      const ChunkProd p = run[t];
      e += p.e + step_bwd(b0, b1, p);
      l += p.l;
    }
    lkl_both[1] = log(fma(q0, b0, q1 * b1)) + e * kLn2 + l + loge0_sum[row];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double lf = lkl_both[0], lb = lkl_both[1];
    if (lf != lf || lb != lb) atomicOr(status, kFlagNaN);
    else if (fabs(lf - lb) > 1e-3) atomicOr(status, kFlagFwBw);   // EM.cpp:166
  }

  {
    double x0 = run_in[threadIdx.x].x, x1 = run_in[threadIdx.x].y;
    for (uint32_t c = lo; c < hi; c++) {
      fc[c] = make_double2(x0, x1);
      step_fwd(x0, x1, cp[c]);
    }
    double b0 = run_out[threadIdx.x].x, b1 = run_out[threadIdx.x].y;
    for (uint32_t c = hi; c-- > lo;) {
      bc[c] = make_double2(b0, b1);
      step_bwd(b0, b1, cp[c]);
    }
  }
}

__global__ void __launch_bounds__(kScanThreads)
estep_chunk_apply(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
                  const double *__restrict__ alpha, const double2 *__restrict__ fwd_carry,
                  const double2 *__restrict__ bwd_carry, double *__restrict__ post, int *__restrict__ status,
                  uint64_t n_rows, uint64_t n_sites, uint64_t site_block, uint32_t n_tiles) {
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
  const size_t chunk = ((size_t) row * n_tiles + tile) * kScanThreads + threadIdx.x;
  const double2 cf = fwd_carry[chunk], cb = bwd_carry[chunk];

  // sweep 1: kappa for every site (kept in shared memory) and the forward
  // vector at the start of each sub-block (checkpoints, also in shared memory)
  double a0 = cf.x, a1 = cf.y;
  double2 *ck = sm.ck + threadIdx.x * (kChunk / kSub);
#pragma unroll 1
  for (int sb = 0; sb < kChunk / kSub; sb++) {
    renorm2(a0, a1);
    ck[sb] = make_double2(a0, a1);
#pragma unroll
    for (int i = 0; i < kSub; i++) {
      const int j = sb * kSub + i;
      const double kap = site_kappa(al * d[j], sm.tab);
      d[j] = kap;
      if (j < n_valid) forward_site(a0, a1, kap * q0, kap * q1, r[j]);
      if (i == kSub / 2) renorm2(a0, a1);
    }
  }

  // sweep 2, sub-blocks from the right: rebuild the forward vectors of the
  // sub-block in registers, then run the backward vector through it.
  double b0 = cb.x, b1 = cb.y;
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
      if (j < n_valid) forward_site(a0, a1, k0[i], k1[i], rr[i]);
      if (i == kSub / 2) renorm2(a0, a1);
      f0[i] = a0; f1[i] = a1;
    }
#pragma unroll
    for (int i = kSub - 1; i >= 0; i--) {
      const int j = sb * kSub + i;
      const double num = f1[i] * b1;
      const double den = fma(f0[i], b0, num);
      double p = num * rcp_pos(den);
      if (j < n_valid) {
        bad |= (p != p);
        p = (p < kEps) ? 0.0 : p;              // check_interv, gen_func.cpp:59-66
        p = (p > 1.0 - kEps) ? 1.0 : p;
        backward_site(b0, b1, k0[i], k1[i], rr[i]);
        if (i == kSub / 2) renorm2(b0, b1);
      } else {
        p = 0.0;
      }
      r[j] = p;
    }
    renorm2(b0, b1);
  }
  if (bad) atomicOr(status, kFlagNaN);

  fence_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_1d(post + tile_at, sm.r, kTileBytes);
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------
// Batched forward-only objective: up to kMaxPoints (F, alpha) points of one
// individual share one staged read of its emissions.
// ---------------------------------------------------------------------------

struct LklSmem {
  TileSmem t;
  M2 m[kMaxPoints][kScanThreads / 32];
  int e[kMaxPoints][kScanThreads / 32];
  double l[kMaxPoints][kScanThreads / 32];
};

template <int NP>
__device__ __forceinline__ void lkl_tile_body(const LklGroup &g, LklSmem &sm, int n_valid,
                                              TileProd *__restrict__ out_row, uint32_t n_tiles, uint32_t tile) {
  const double *r = sm.t.r + threadIdx.x * kChunk;
  const double *d = sm.t.d + threadIdx.x * kChunk;
  M2 m[NP];
  int e[NP];
  double ls[NP];
  bool fresh[NP];   // does point p need its own kappa, or does it share alpha with p-1
#pragma unroll
  for (int p = 0; p < NP; p++) {
    m[p] = identity2(); e[p] = 0; ls[p] = 0.0;
    fresh[p] = (p == 0) || (g.alpha[p] != g.alpha[p - 1]);
  }
#pragma unroll 3
  for (int j = 0; j < kChunk; j++) {
    const double dj = d[j], rj = r[j];
    const bool live = j < n_valid;
    double kap = 0.0, l = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      if (fresh[p]) { l = 0.0; kap = site_kappa(g.alpha[p] * dj, sm.t.tab, l); }
      if (live) {
        apply_site(m[p], kap * (1.0 - g.F[p]), kap * g.F[p], rj);
        ls[p] += l;
      }
    }
    if (j % 6 == 5) {
#pragma unroll
      for (int p = 0; p < NP; p++) e[p] += renorm(m[p]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    e[p] += renorm(m[p]);
    warp_ordered_product(m[p], e[p]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ls[p] += __shfl_down_sync(kFull, ls[p], off);
    if (lane == 0) { sm.m[p][warp] = m[p]; sm.e[p][warp] = e[p]; sm.l[p][warp] = ls[p]; }
  }
  __syncthreads();
  if ((int) threadIdx.x < NP) {
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
    out_row[(size_t) p * n_tiles + tile] = out;
  }
}

__global__ void __launch_bounds__(kScanThreads)
lkl_tile_products(const double *__restrict__ emis, const double *__restrict__ dist,
                  const LklGroup *__restrict__ groups, TileProd *__restrict__ tile_prod, uint64_t n_rows,
                  uint64_t n_sites, uint64_t site_block, uint32_t n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LklSmem &sm = *reinterpret_cast<LklSmem *>(smem_raw);
  const uint32_t tile = blockIdx.x, grp = blockIdx.y;
  const LklGroup g = groups[grp];
  const uint64_t tile_first = (uint64_t) tile * kTile;
  stage_tile(sm.t, emis + blocked_index((uint64_t) g.ind, tile_first, n_rows, site_block), dist + tile_first);
  const int n_valid = valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, n_sites);
  TileProd *out_row = tile_prod + (size_t) grp * kMaxPoints * n_tiles;
  switch (g.npts) {   // uniform per CTA
    case 1: lkl_tile_body<1>(g, sm, n_valid, out_row, n_tiles, tile); break;
    case 2: lkl_tile_body<2>(g, sm, n_valid, out_row, n_tiles, tile); break;
    case 3: lkl_tile_body<3>(g, sm, n_valid, out_row, n_tiles, tile); break;
    case 4: lkl_tile_body<4>(g, sm, n_valid, out_row, n_tiles, tile); break;
    default: lkl_tile_body<5>(g, sm, n_valid, out_row, n_tiles, tile); break;
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

static bool g_attr_done = false;
static void set_smem_attrs() {
  if (g_attr_done) return;
  cudaFuncSetAttribute(estep_chunk_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(TileSmem));
  cudaFuncSetAttribute(estep_chunk_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(TileSmem));
  cudaFuncSetAttribute(lkl_tile_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LklSmem));
  g_attr_done = true;
}

void launch_estep(const EstepArgs &a, cudaStream_t st) {
  set_smem_attrs();
  dim3 grid(a.n_tiles, (unsigned) a.n_rows_valid);
  estep_chunk_products<<<grid, kScanThreads, sizeof(TileSmem), st>>>(a.emis, a.dist, a.indF, a.alpha, a.chunk_prod,
                                                                     a.n_rows, a.n_sites, a.site_block, a.n_tiles);
  estep_chunk_scan<<<(unsigned) a.n_rows_valid, kCarryThreads, 0, st>>>(
      a.chunk_prod, a.indF, a.loge0_sum, a.fwd_carry, a.bwd_carry, a.ind_lkl, a.status, a.n_tiles * kScanThreads);
  estep_chunk_apply<<<grid, kScanThreads, sizeof(TileSmem), st>>>(a.emis, a.dist, a.indF, a.alpha, a.fwd_carry,
                                                                  a.bwd_carry, a.post, a.status, a.n_rows, a.n_sites,
                                                                  a.site_block, a.n_tiles);
}

void launch_lkl_batch(const LklArgs &a, cudaStream_t st) {
  set_smem_attrs();
  dim3 grid(a.n_tiles, a.n_groups);
  lkl_tile_products<<<grid, kScanThreads, sizeof(LklSmem), st>>>(a.emis, a.dist, a.groups, a.tile_prod, a.n_rows,
                                                                 a.n_sites, a.site_block, a.n_tiles);
  const unsigned warps = a.n_groups * kMaxPoints;
  lkl_finish<<<(warps + 3) / 4, 128, 0, st>>>(a.tile_prod, a.groups, a.loge0_sum, a.neg_lkl, a.n_groups, a.n_tiles);
}

}  // namespace nfh

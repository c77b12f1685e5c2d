// nfh_estep.cu - fused forward-backward E-step and the batched forward-only
// objective, as chunked scans of scaled 2x2 products.
//
// Replaces, for all individuals at once:
//   forward()  shared/HMM.cpp:6-28      backward() shared/HMM.cpp:33-60
//   ind_lkl + clamped posterior          EM.cpp:178-185, check_interv gen_func.cpp:55-70
//   Fw/Bw consistency check              EM.cpp:166-170
//   lkl() (BFGS objective)               EM.cpp:449-464
//
// Three launches per E-step:
//   estep_tile_products : every CTA reduces one 2048-site tile of one
//                         individual to a single scaled 2x2 product.
//   estep_carries       : per individual, running products over tiles give the
//                         forward carry into and the backward carry out of
//                         every tile, plus the log-likelihood (both ways).
//   estep_apply         : every CTA re-reads its tile, rebuilds the per-thread
//                         products, scans them inside the CTA, then runs the
//                         forward and backward vector recursions per site and
//                         writes the clamped posterior of the IBD state.
#include "nfh_device.cuh"
#include "nfh_kernels.h"

namespace nfh {

// Load this thread's kSitesPerThread consecutive doubles (16-byte vector loads).
__device__ __forceinline__ void load_chunk(const double *__restrict__ base, double (&v)[kSitesPerThread]) {
  const double2 *p = reinterpret_cast<const double2 *>(base);
#pragma unroll
  for (int j = 0; j < kSitesPerThread / 2; j++) {
    double2 t = __ldg(p + j);
    v[2 * j] = t.x; v[2 * j + 1] = t.y;
  }
}

__device__ __forceinline__ void store_chunk(double *base, const double (&v)[kSitesPerThread]) {
  double2 *p = reinterpret_cast<double2 *>(base);
#pragma unroll
  for (int j = 0; j < kSitesPerThread / 2; j++) p[j] = make_double2(v[2 * j], v[2 * j + 1]);
}

// Number of real sites among this thread's kSitesPerThread (the last tile is padded).
__device__ __forceinline__ int valid_sites(uint64_t first_site, uint64_t n_sites) {
  return first_site >= n_sites ? 0 : (int) min((uint64_t) kSitesPerThread, n_sites - first_site);
}

// Product of this thread's factored site matrices N_s, left to right; padding
// sites act as the identity.  On return r[] is untouched and kap[] holds kappa_s.
__device__ __forceinline__ M2 chunk_product(const double (&r)[kSitesPerThread], const double (&kap)[kSitesPerThread],
                                            double q0, double q1, int n_valid, int &e) {
  M2 m = identity2();
  e = 0;
#pragma unroll
  for (int j = 0; j < kSitesPerThread; j++) {
    if (j < n_valid) apply_site(m, kap[j] * q0, kap[j] * q1, r[j]);
    if (j == kSitesPerThread / 2 - 1) e += renorm(m);
  }
  e += renorm(m);
  return m;
}

__global__ void __launch_bounds__(kScanThreads)
estep_tile_products(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
                    const double *__restrict__ alpha, TileProd *__restrict__ tile_prod, uint64_t n_rows,
                    uint64_t n_sites, uint64_t site_block, uint32_t n_tiles) {
  __shared__ double tab[64];
  __shared__ M2 sm[kScanThreads / 32];
  __shared__ int se[kScanThreads / 32];
  __shared__ double sl[kScanThreads / 32];
  load_exp_table(tab);
  __syncthreads();

  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t first = (uint64_t) tile * kTile + (uint64_t) threadIdx.x * kSitesPerThread;
  const int n_valid = valid_sites(first, n_sites);
  const double F = indF[row], al = alpha[row];
  const double q0 = 1.0 - F, q1 = F;

  double r[kSitesPerThread], kap[kSitesPerThread];
  load_chunk(emis + blocked_index(row, first, n_rows, site_block), r);
  load_chunk(dist + first, kap);
  double ls = 0.0;
#pragma unroll
  for (int j = 0; j < kSitesPerThread; j++) {
    double l = 0.0;
    kap[j] = site_kappa(al * kap[j], tab, l);
    if (j < n_valid) ls += l;
  }

  int e;
  M2 m = chunk_product(r, kap, q0, q1, n_valid, e);
  warp_ordered_product(m, e);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ls += __shfl_down_sync(kFull, ls, off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm[warp] = m; se[warp] = e; sl[warp] = ls; }
  __syncthreads();
  if (threadIdx.x == 0) {
    M2 acc = sm[0];
    int ae = se[0];
    double al_sum = sl[0];
#pragma unroll
    for (int w = 1; w < kScanThreads / 32; w++) {
      acc = matmul(acc, sm[w]);
      ae += se[w] + renorm(acc);
      al_sum += sl[w];
    }
    TileProd out;
    out.a = acc.a; out.b = acc.b; out.c = acc.c; out.d = acc.d; out.e = (double) ae; out.l = al_sum;
    tile_prod[(size_t) row * n_tiles + tile] = out;
  }
}

// One thread per individual: sequential pass over its tile products.
__global__ void estep_carries(const TileProd *__restrict__ tile_prod, const double *__restrict__ indF,
                              const double *__restrict__ loge0_sum, double2 *__restrict__ fwd_carry,
                              double2 *__restrict__ bwd_carry, double *__restrict__ ind_lkl, int *__restrict__ status,
                              uint32_t n_rows_valid, uint32_t n_tiles) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows_valid) return;
  const double F = indF[row];
  const double q0 = 1.0 - F, q1 = F;
  const TileProd *tp = tile_prod + (size_t) row * n_tiles;

  double x0 = q0, x1 = q1, lsum = 0.0;
  long long ex = 0;
  for (uint32_t t = 0; t < n_tiles; t++) {
    fwd_carry[(size_t) row * n_tiles + t] = make_double2(x0, x1);
    TileProd p = tp[t];
    double y0 = fma(x0, p.a, x1 * p.c), y1 = fma(x0, p.b, x1 * p.d);
    x0 = y0; x1 = y1;
    ex += (long long) p.e + renorm2(x0, x1);
    lsum += p.l;
  }
  const double base = lsum + loge0_sum[row];
  const double lf = log(x0 + x1) + (double) ex * kLn2 + base;

  double b0 = 1.0, b1 = 1.0;
  long long eb = 0;
  for (uint32_t t = n_tiles; t-- > 0;) {
    bwd_carry[(size_t) row * n_tiles + t] = make_double2(b0, b1);
    TileProd p = tp[t];
    double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
    b0 = y0; b1 = y1;
    eb += (long long) p.e + renorm2(b0, b1);
  }
  const double lb = log(fma(q0, b0, q1 * b1)) + (double) eb * kLn2 + base;

  ind_lkl[row] = lf;
  if (lf != lf || lb != lb) atomicOr(status, kFlagNaN);
  else if (fabs(lf - lb) > 1e-3) atomicOr(status, kFlagFwBw);   // EM.cpp:166
}

__global__ void __launch_bounds__(kScanThreads)
estep_apply(const double *__restrict__ emis, const double *__restrict__ dist, const double *__restrict__ indF,
            const double *__restrict__ alpha, const double2 *__restrict__ fwd_carry,
            const double2 *__restrict__ bwd_carry, double *__restrict__ post, int *__restrict__ status,
            uint64_t n_rows, uint64_t n_sites, uint64_t site_block, uint32_t n_tiles) {
  constexpr int kWarps = kScanThreads / 32;
  __shared__ double tab[64];
  __shared__ M2 warp_tot[kWarps];
  __shared__ double2 warp_in[kWarps], warp_out[kWarps];
  load_exp_table(tab);
  __syncthreads();

  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t first = (uint64_t) tile * kTile + (uint64_t) threadIdx.x * kSitesPerThread;
  const int n_valid = valid_sites(first, n_sites);
  const double F = indF[row], al = alpha[row];
  const double q0 = 1.0 - F, q1 = F;
  const size_t base = blocked_index(row, first, n_rows, site_block);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  double r[kSitesPerThread], kap[kSitesPerThread];
  load_chunk(emis + base, r);
  load_chunk(dist + first, kap);
#pragma unroll
  for (int j = 0; j < kSitesPerThread; j++) kap[j] = site_kappa(al * kap[j], tab);

  int e_unused;
  const M2 mine = chunk_product(r, kap, q0, q1, n_valid, e_unused);

  // Inclusive prefix (lanes <= me) and suffix (lanes >= me) products inside the
  // warp.  Only directions matter from here on (the posterior is scale free),
  // so products are renormalised without tracking exponents.
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
  if (threadIdx.x < kWarps) {
    // forward carry into warp w: tile carry times totals of warps < w
    double2 cf = fwd_carry[(size_t) row * n_tiles + tile];
    double x0 = cf.x, x1 = cf.y;
    for (int w = 0; w < (int) threadIdx.x; w++) {
      M2 p = warp_tot[w];
      double y0 = fma(x0, p.a, x1 * p.c), y1 = fma(x0, p.b, x1 * p.d);
      x0 = y0; x1 = y1;
      renorm2(x0, x1);
    }
    warp_in[threadIdx.x] = make_double2(x0, x1);
    // backward carry out of warp w: totals of warps > w times tile carry
    double2 cb = bwd_carry[(size_t) row * n_tiles + tile];
    double b0 = cb.x, b1 = cb.y;
    for (int w = kWarps - 1; w > (int) threadIdx.x; w--) {
      M2 p = warp_tot[w];
      double y0 = fma(p.a, b0, p.b * b1), y1 = fma(p.c, b0, p.d * b1);
      b0 = y0; b1 = y1;
      renorm2(b0, b1);
    }
    warp_out[threadIdx.x] = make_double2(b0, b1);
  }
  __syncthreads();

  // exclusive products of the neighbouring lanes
  M2 before = shfl_up_m(pre, 1);     // product of lanes < me (valid for lane > 0)
  M2 after = shfl_down_m(suf, 1);    // product of lanes > me (valid for lane < 31)
  double a0 = warp_in[warp].x, a1 = warp_in[warp].y;
  if (lane > 0) {
    double y0 = fma(a0, before.a, a1 * before.c), y1 = fma(a0, before.b, a1 * before.d);
    a0 = y0; a1 = y1;
    renorm2(a0, a1);
  }
  double b0 = warp_out[warp].x, b1 = warp_out[warp].y;
  if (lane < 31) {
    double y0 = fma(after.a, b0, after.b * b1), y1 = fma(after.c, b0, after.d * b1);
    b0 = y0; b1 = y1;
    renorm2(b0, b1);
  }

  // forward vectors at each of my sites (after absorbing the site)
  double f0[kSitesPerThread], f1[kSitesPerThread];
#pragma unroll
  for (int j = 0; j < kSitesPerThread; j++) {
    if (j < n_valid) {
      forward_site(a0, a1, kap[j] * q0, kap[j] * q1, r[j]);
      if (j == kSitesPerThread / 2 - 1) renorm2(a0, a1);
    }
    f0[j] = a0; f1[j] = a1;
  }

  // backward sweep: posterior at site j uses f[j] and the backward vector
  // *after* site j; then the site is absorbed into the backward vector.
  double out[kSitesPerThread];
  bool bad = false;
#pragma unroll
  for (int j = kSitesPerThread - 1; j >= 0; j--) {
    const double num = f1[j] * b1;
    const double den = fma(f0[j], b0, num);
    double p = num * rcp_pos(den);
    if (j < n_valid) {
      bad |= (p != p);
      p = (p < kEps) ? 0.0 : p;              // check_interv, gen_func.cpp:59-66
      p = (p > 1.0 - kEps) ? 1.0 : p;
      out[j] = p;
      backward_site(b0, b1, kap[j] * q0, kap[j] * q1, r[j]);
      if (j == kSitesPerThread / 2) renorm2(b0, b1);
    } else {
      out[j] = 0.0;
    }
  }
  store_chunk(post + base, out);
  if (bad) atomicOr(status, kFlagNaN);
}

// ---------------------------------------------------------------------------
// Batched forward-only objective: up to kMaxPoints (F, alpha) points of one
// individual share one read of its emissions.
// ---------------------------------------------------------------------------

template <int NP>
__device__ __forceinline__ void lkl_tile_body(const LklGroup &g, const double (&r)[kSitesPerThread],
                                              const double (&d)[kSitesPerThread], int n_valid,
                                              const double *__restrict__ tab, M2 (*sm)[kScanThreads / 32],
                                              int (*se)[kScanThreads / 32], double (*sl)[kScanThreads / 32],
                                              TileProd *__restrict__ out_row, uint32_t n_tiles, uint32_t tile) {
  M2 m[NP];
  int e[NP];
  double ls[NP];
  bool fresh[NP];   // does point p need its own exp(), or does it share alpha with p-1
#pragma unroll
  for (int p = 0; p < NP; p++) {
    m[p] = identity2(); e[p] = 0; ls[p] = 0.0;
    fresh[p] = (p == 0) || (g.alpha[p] != g.alpha[p - 1]);
  }
#pragma unroll
  for (int j = 0; j < kSitesPerThread; j++) {
    double kap = 0.0, l = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      if (fresh[p]) { l = 0.0; kap = site_kappa(g.alpha[p] * d[j], tab, l); }
      if (j < n_valid) {
        apply_site(m[p], kap * (1.0 - g.F[p]), kap * g.F[p], r[j]);
        ls[p] += l;
      }
    }
    if (j == kSitesPerThread / 2 - 1) {
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
    if (lane == 0) { sm[p][warp] = m[p]; se[p][warp] = e[p]; sl[p][warp] = ls[p]; }
  }
  __syncthreads();
  if ((int) threadIdx.x < NP) {
    const int p = threadIdx.x;
    M2 acc = sm[p][0];
    int ae = se[p][0];
    double al_sum = sl[p][0];
    for (int w = 1; w < kScanThreads / 32; w++) {
      acc = matmul(acc, sm[p][w]);
      ae += se[p][w] + renorm(acc);
      al_sum += sl[p][w];
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
  __shared__ double tab[64];
  __shared__ M2 sm[kMaxPoints][kScanThreads / 32];
  __shared__ int se[kMaxPoints][kScanThreads / 32];
  __shared__ double sl[kMaxPoints][kScanThreads / 32];
  load_exp_table(tab);
  __syncthreads();

  const uint32_t tile = blockIdx.x, grp = blockIdx.y;
  const LklGroup g = groups[grp];
  const uint64_t first = (uint64_t) tile * kTile + (uint64_t) threadIdx.x * kSitesPerThread;
  const int n_valid = valid_sites(first, n_sites);

  double r[kSitesPerThread], d[kSitesPerThread];
  load_chunk(emis + blocked_index((uint64_t) g.ind, first, n_rows, site_block), r);
  load_chunk(dist + first, d);
  TileProd *out_row = tile_prod + (size_t) grp * kMaxPoints * n_tiles;
  switch (g.npts) {   // uniform per CTA
    case 1: lkl_tile_body<1>(g, r, d, n_valid, tab, sm, se, sl, out_row, n_tiles, tile); break;
    case 2: lkl_tile_body<2>(g, r, d, n_valid, tab, sm, se, sl, out_row, n_tiles, tile); break;
    case 3: lkl_tile_body<3>(g, r, d, n_valid, tab, sm, se, sl, out_row, n_tiles, tile); break;
    case 4: lkl_tile_body<4>(g, r, d, n_valid, tab, sm, se, sl, out_row, n_tiles, tile); break;
    default: lkl_tile_body<5>(g, r, d, n_valid, tab, sm, se, sl, out_row, n_tiles, tile); break;
  }
}

// One thread per (group, point): chain the tile products, emit -logLkl.
__global__ void lkl_finish(const TileProd *__restrict__ tile_prod, const LklGroup *__restrict__ groups,
                           const double *__restrict__ loge0_sum, double *__restrict__ neg_lkl, uint32_t n_groups,
                           uint32_t n_tiles) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t grp = idx / kMaxPoints, p = idx % kMaxPoints;
  if (grp >= n_groups) return;
  const LklGroup &g = groups[grp];
  if ((int) p >= g.npts) return;
  const TileProd *tp = tile_prod + ((size_t) grp * kMaxPoints + p) * n_tiles;
  double x0 = 1.0 - g.F[p], x1 = g.F[p], lsum = 0.0;
  long long ex = 0;
  for (uint32_t t = 0; t < n_tiles; t++) {
    TileProd q = tp[t];
    double y0 = fma(x0, q.a, x1 * q.c), y1 = fma(x0, q.b, x1 * q.d);
    x0 = y0; x1 = y1;
    ex += (long long) q.e + renorm2(x0, x1);
    lsum += q.l;
  }
  neg_lkl[g.out[p]] = -(log(x0 + x1) + (double) ex * kLn2 + lsum + loge0_sum[g.ind]);
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------

void launch_estep(const EstepArgs &a, cudaStream_t st) {
  dim3 grid(a.n_tiles, (unsigned) a.n_rows_valid);
  estep_tile_products<<<grid, kScanThreads, 0, st>>>(a.emis, a.dist, a.indF, a.alpha, a.tile_prod, a.n_rows,
                                                      a.n_sites, a.site_block, a.n_tiles);
  estep_carries<<<(unsigned) ((a.n_rows_valid + 63) / 64), 64, 0, st>>>(
      a.tile_prod, a.indF, a.loge0_sum, a.fwd_carry, a.bwd_carry, a.ind_lkl, a.status, (uint32_t) a.n_rows_valid,
      a.n_tiles);
  estep_apply<<<grid, kScanThreads, 0, st>>>(a.emis, a.dist, a.indF, a.alpha, a.fwd_carry, a.bwd_carry, a.post,
                                              a.status, a.n_rows, a.n_sites, a.site_block, a.n_tiles);
}

void launch_lkl_batch(const LklArgs &a, cudaStream_t st) {
  dim3 grid(a.n_tiles, a.n_groups);
  lkl_tile_products<<<grid, kScanThreads, 0, st>>>(a.emis, a.dist, a.groups, a.tile_prod, a.n_rows, a.n_sites,
                                                    a.site_block, a.n_tiles);
  const unsigned n = a.n_groups * kMaxPoints;
  lkl_finish<<<(n + 63) / 64, 64, 0, st>>>(a.tile_prod, a.groups, a.loge0_sum, a.neg_lkl, a.n_groups, a.n_tiles);
}

}  // namespace nfh

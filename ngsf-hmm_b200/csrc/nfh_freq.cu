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
#include "nfh_device.cuh"
#include "nfh_kernels.h"

namespace nfh {

constexpr int kGroupLanes = 8;        // lanes sharing one site (individual groups) in the warp variant
constexpr int kFreqThreads = 128;
constexpr int kSitesPerWarp = 32 / kGroupLanes;
constexpr int kSitesPerCta = kSitesPerWarp * (kFreqThreads / 32);
constexpr int kMaxK = 16;             // individuals per lane -> n_ind <= 128 for the warp variant

struct IndCoef {   // pass-invariant coefficients of one individual at one site
  double a0, a2, b2, c1, h, g;
};

// With u = (1-f)^2, v = f^2, a = f(1-f):
//   w1 = c1 a            (= L1 * het prior)         c1 = 2 L1 (1-F)
//   w2 = a2 v + b2 a     (= L2 * hom-alt prior)     a2 = L2, b2 = L2 F
//   S  = a0 u + a2 v + h a  (= w0 + w1 + w2)        a0 = L0, h = L0 F + c1 + b2
//   num += (w1 + g w2) / S                          g = 2 - F
//   den += g + F w1 / S      [2 w1 + (w0 + w2) g = g S + F w1]
__device__ __forceinline__ IndCoef make_coef(double L0, double L1, double L2, double F) {
  IndCoef k;
  k.a0 = L0; k.a2 = L2; k.b2 = L2 * F;
  k.c1 = 2.0 * L1 * (1.0 - F);
  // A heterozygote call (L0 = L2 = 0) at a site whose IBD posterior was
  // clamped to exactly 1 has zero weight for every genotype; the reference's
  // log-space arithmetic (-1e15 stands for log 0) resolves this to "certainly
  // heterozygous".  Keep a vanishing het weight so the ratio is 1, not 0/0.
  if (L0 == 0.0 && L2 == 0.0 && F == 1.0) k.c1 = 1e-280;
  k.h = L0 * F + k.c1 + k.b2;
  k.g = 2.0 - F;
  return k;
}

__device__ __forceinline__ IndCoef null_coef() {   // padding slot: contributes exactly 0
  IndCoef k;
  k.a0 = 0.5; k.a2 = 0.5; k.b2 = 0.0; k.c1 = 0.0; k.h = 0.0; k.g = 0.0;
  return k;
}

// 14 FP64 instructions, no branches: the K independent chains of a lane interleave.
// den_f accumulates sum F w1 / S only; the constant sum of g is added per pass by the caller.
__device__ __forceinline__ void accumulate(const IndCoef &k, double u, double v, double a, double &num_a,
                                           double &num_b, double &den_f) {
  const double x2 = k.a2 * v;
  const double w2 = fma(k.b2, a, x2);
  const double w1 = k.c1 * a;
  const double S = fma(k.a0, u, fma(k.h, a, x2));
  const double rinv = rcp_pos(S);
  const double t1 = w1 * rinv;
  num_a += t1;
  num_b = fma(k.g * w2, rinv, num_b);
  den_f = fma(2.0 - k.g, t1, den_f);
}

// state emissions from linear GL at frequency f (calc_emission with F = 0 / 1)
__device__ __forceinline__ void emissions(double L0, double L1, double L2, double f, double &e0, double &e1) {
  double omf = 1.0 - f;
  double a = omf * f;
  double u = omf * omf, v = f * f;
  e0 = fma(L0, u, fma(L1, 2.0 * a, L2 * v));
  e1 = fma(L0, u + a, L2 * (v + a));       // het prior is exp(-1e15) = 0 when F == 1
}

template <int K>
__global__ void __launch_bounds__(kFreqThreads)
freq_emission_warp(FreqArgs A, unsigned n_site_tiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane & (kGroupLanes - 1), sub = lane / kGroupLanes;
  extern __shared__ double loge0_acc[];     // [warps][n_ind_pad]
  for (unsigned i = threadIdx.x; i < (kFreqThreads / 32) * A.n_ind_pad; i += kFreqThreads) loge0_acc[i] = 0.0;
  __syncthreads();
  double *my_acc = loge0_acc + (size_t) warp * A.n_ind_pad;

  for (unsigned tile = blockIdx.x; tile < n_site_tiles; tile += gridDim.x) {
    const uint64_t site = (uint64_t) tile * kSitesPerCta + warp * kSitesPerWarp + sub;
    const bool site_ok = site < A.sites_owned;
    const uint64_t sl = site_ok ? site : 0;

    IndCoef coef[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) kGroupLanes * k;
      if (i < A.n_ind) {
        const size_t at = (size_t) i * A.site_block + sl;
        const double F = A.post ? A.post[at] : 0.0;
        coef[k] = make_coef(A.gl0[at], A.gl1[at], A.gl2[at], F);
      } else {
        coef[k] = null_coef();
      }
    }

    double freq = A.update_freq ? 0.01 : A.freq[sl];
    if (A.update_freq) {
      double g_sum = 0.0;                       // sum over individuals of (2 - F): constant part of den per pass
#pragma unroll
      for (int k = 0; k < K; k++) g_sum += coef[k].g;
#pragma unroll
      for (int m = 1; m < kGroupLanes; m <<= 1) g_sum += __shfl_xor_sync(kFull, g_sum, m);
      double num = 0.0, den = 0.0;
      bool active = site_ok;
      int passes = 0;
      while (__any_sync(kFull, active)) {
        const double omf = 1.0 - freq;
        const double u = omf * omf, v = freq * freq, a = omf * freq;
        double pa = 0.0, pb = 0.0, pd = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) accumulate(coef[k], u, v, a, pa, pb, pd);
        double pn = pa + pb;
#pragma unroll
        for (int m = 1; m < kGroupLanes; m <<= 1) {
          pn += __shfl_xor_sync(kFull, pn, m);
          pd += __shfl_xor_sync(kFull, pd, m);
        }
        pd += g_sum;
        passes++;
        if (active) {
          num += pn; den += pd;
          const double before = freq;
          freq = num / den;
          // do { ... } while (|before - freq| > EPSILON && iters++ < 100)   gen_func.cpp:1006
          active = (fabs(before - freq) > kEps) && (passes <= 100);
        }
      }
      if (site_ok && grp == 0) A.freq[site] = freq;
    }

    // emission refresh from the same site (L1 is re-read: the coefficients
    // hold 2 L1 (1-F), which loses L1 when F == 1)
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint64_t i = (uint64_t) grp + (uint64_t) kGroupLanes * k;
      double le0 = 0.0;
      if (i < A.n_ind && site_ok) {
        const size_t at = (size_t) i * A.site_block + site;
        double e0, e1;
        emissions(coef[k].a0, A.gl1[at], coef[k].a2, freq, e0, e1);
        A.emis[at] = e1 / e0;
        if (A.e0) A.e0[at] = e0;
        le0 = log(e0);
      }
      // sum over the warp's sites (lanes with equal grp), fixed order
#pragma unroll
      for (int m = kGroupLanes; m < 32; m <<= 1) le0 += __shfl_xor_sync(kFull, le0, m);
      if (sub == 0 && i < A.n_ind_pad) my_acc[i] += le0;
    }
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < A.n_ind_pad; i += kFreqThreads) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kFreqThreads / 32; w++) s += loge0_acc[(size_t) w * A.n_ind_pad + i];
    A.loge0_part[(size_t) blockIdx.x * A.n_ind_pad + i] = s;
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
        double pa = 0.0, pb = 0.0, pd = 0.0;
        for (uint64_t i = 0; i < A.n_ind; i++) {
          const size_t at = (size_t) i * A.site_block + site;
          const double F = A.post ? A.post[at] : 0.0;
          IndCoef k = make_coef(A.gl0[at], A.gl1[at], A.gl2[at], F);
          accumulate(k, u, v, a, pa, pb, pd);
          pd += k.g;
        }
        num += pa + pb; den += pd;
        freq = num / den;
      } while (fabs(before - freq) > kEps && passes++ < 100);
      A.freq[site] = freq;
    }
    for (uint64_t i = 0; i < A.n_ind; i++) {
      const size_t at = (size_t) i * A.site_block + site;
      double e0, e1;
      emissions(A.gl0[at], A.gl1[at], A.gl2[at], freq, e0, e1);
      A.emis[at] = e1 / e0;
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

unsigned freq_grid_size(const FreqArgs &a, int sm_count) {
  if (a.n_ind <= (uint64_t) kGroupLanes * kMaxK) {
    unsigned tiles = (unsigned) ((a.sites_owned + kSitesPerCta - 1) / kSitesPerCta);
    unsigned cap = (unsigned) sm_count * 4u;
    return tiles < cap ? (tiles ? tiles : 1u) : cap;
  }
  return 64;   // chunks of loge0_rowsum
}

template <int K>
static void launch_warp_variant(const FreqArgs &a, unsigned grid, unsigned tiles, cudaStream_t st) {
  size_t smem = (size_t) (kFreqThreads / 32) * a.n_ind_pad * sizeof(double);
  freq_emission_warp<K><<<grid, kFreqThreads, smem, st>>>(a, tiles);
}

int launch_freq_emission(const FreqArgs &a, unsigned grid, cudaStream_t st) {
  if (a.n_ind <= (uint64_t) kGroupLanes * kMaxK) {
    const unsigned tiles = (unsigned) ((a.sites_owned + kSitesPerCta - 1) / kSitesPerCta);
    const int K = (int) ((a.n_ind + kGroupLanes - 1) / kGroupLanes);
    switch (K) {
#define NFH_CASE(k) case k: launch_warp_variant<k>(a, grid, tiles, st); break;
      NFH_CASE(1) NFH_CASE(2) NFH_CASE(3) NFH_CASE(4) NFH_CASE(5) NFH_CASE(6) NFH_CASE(7) NFH_CASE(8)
      NFH_CASE(9) NFH_CASE(10) NFH_CASE(11) NFH_CASE(12) NFH_CASE(13) NFH_CASE(14) NFH_CASE(15) NFH_CASE(16)
#undef NFH_CASE
      default: return 1;
    }
    return 1;   // launches
  }
  unsigned blocks = (unsigned) ((a.sites_owned + kFreqThreads - 1) / kFreqThreads);
  freq_emission_stream<<<blocks ? blocks : 1, kFreqThreads, 0, st>>>(a);
  dim3 g2(grid, (unsigned) a.n_ind);
  loge0_rowsum<<<g2, 256, 0, st>>>(a, grid);
  return 2;
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

// simt_kernels_host.cpp - TEST INFRASTRUCTURE ONLY: the product's CUDA kernels and their launchers, compiled for the
// SIMT emulator of tests/simt/simt.h from mechanically rewritten copies of ngsf-hmm_b200/csrc/*.cu (see
// tests/test_simt_kernels_cpu.py for the rewrites), behind a few C entry points that fill EstepArgs / LklArgs /
// ViterbiArgs the way nfh_ctx.cu does for one rank.
#include <cstdint>
#include <cstring>
#include <vector>

#include "simt.h"

// the rewritten kernel sources (scratch directory, first on the include path)
#include "nfh_estep.cu"
#include "nfh_lkl.cu"
#include "nfh_viterbi.cu"
// the frequency kernels and their entry points: tests/simt_freq_host.cpp, compiled side by side (most of the compile time)

using namespace nfh;

struct SimtCtx {
  uint64_t N = 0, S = 0, S_pad = 0;
  uint32_t n_tiles = 0;
  std::vector<double> emis, e0, dist, post, tile_dmax, tile_dsum, indF, alpha, loge0, ind_lkl, neg_lkl;
  std::vector<double4> chunk_prod, vit_tile_prod;
  std::vector<TileProd> tile_prod, lkl_tile_prod;
  std::vector<double2> fwd, bwd;
  std::vector<LklGroup> groups;
  std::vector<unsigned char> vit_work, vit_maps;
  std::vector<int> vit_final;
  int status = 0;
};

static EstepArgs estep_args(SimtCtx *c) {                       // nfh_ctx.cu: estep_args(), one rank
  EstepArgs a;
  std::memset(&a, 0, sizeof a);
  a.emis = c->emis.data(); a.dist = c->dist.data(); a.indF = c->indF.data(); a.alpha = c->alpha.data();
  a.tile_dmax = c->tile_dmax.data(); a.tile_dsum = c->tile_dsum.data();
  a.loge0_sum = c->loge0.data();
  a.chunk_prod = c->chunk_prod.data(); a.tile_prod = c->tile_prod.data();
  a.fwd_carry = c->fwd.data(); a.bwd_carry = c->bwd.data();
  a.post = c->post.data(); a.ind_lkl = c->ind_lkl.data(); a.status = &c->status;
  a.post_peers.direct = 0; a.post_peers.rank = 0; a.post_peers.n_loc = c->N;
  a.n_rows = c->N; a.n_rows_valid = c->N; a.n_sites = c->S; a.site_block = c->S_pad; a.n_tiles = c->n_tiles;
  a.sm_count = 148;
  a.fused = nullptr;
  return a;
}

extern "C" {

// ratio / e0: [N][S] (e1/e0 and e0, linear), dist [S] Mb, loge0_sum [N]
SimtCtx *simt_create(uint64_t N, uint64_t S, const double *ratio, const double *e0, const double *dist,
                     const double *loge0_sum) {
  SimtCtx *c = new SimtCtx;
  c->N = N; c->S = S;
  c->n_tiles = (uint32_t) ((S + kTile - 1) / kTile);
  c->S_pad = (uint64_t) c->n_tiles * kTile;
  // the context keeps r = 1, e0 = 1, d = 0 at the padding sites (nfh_ctx.cu)
  c->emis.assign(N * c->S_pad, 1.0); c->e0.assign(N * c->S_pad, 1.0); c->post.assign(N * c->S_pad, 0.0);
  c->dist.assign(c->S_pad, 0.0);
  for (uint64_t i = 0; i < N; i++) {
    std::memcpy(&c->emis[i * c->S_pad], ratio + i * S, S * sizeof(double));
    if (e0) std::memcpy(&c->e0[i * c->S_pad], e0 + i * S, S * sizeof(double));
  }
  std::memcpy(c->dist.data(), dist, S * sizeof(double));
  c->tile_dmax.resize(c->n_tiles); c->tile_dsum.resize(c->n_tiles);
  for (uint32_t t = 0; t < c->n_tiles; t++) {                    // nfh_upload_pos_dist
    const uint64_t lo = (uint64_t) t * kTile, hi = std::min<uint64_t>(S, lo + kTile);
    double mx = 0.0, sum = 0.0;
    bool nan = false;
    for (uint64_t s = lo; s < hi; s++) {
      const double d = dist[s];
      nan |= d != d;
      mx = d > mx ? d : mx;
      sum += d;
    }
    c->tile_dmax[t] = nan ? std::nan("") : mx;
    c->tile_dsum[t] = sum;
  }
  c->indF.assign(N, 0.0); c->alpha.assign(N, 0.0); c->ind_lkl.assign(N, 0.0);
  c->loge0.assign(loge0_sum, loge0_sum + N);
  const size_t n_chunks = (size_t) c->n_tiles * kScanThreads;
  c->chunk_prod.resize(N * n_chunks); c->tile_prod.resize(N * c->n_tiles);
  c->lkl_tile_prod.resize(N * kMaxPoints * c->n_tiles);
  c->fwd.resize(N * c->n_tiles); c->bwd.resize(N * c->n_tiles);
  c->groups.resize(N); c->neg_lkl.resize(N * kMaxPoints);
  c->vit_work.assign(N * c->S_pad, 0);
  c->vit_maps.resize(N * (n_chunks + 2 * (size_t) c->n_tiles));
  c->vit_tile_prod.resize(N * c->n_tiles); c->vit_final.resize(N);
  return c;
}
void simt_destroy(SimtCtx *c) { delete c; }
// refreshed emissions (after a frequency update): ratio / e0 [N][S], loge0_sum [N]
void simt_update_emissions(SimtCtx *c, const double *ratio, const double *e0, const double *loge0_sum) {
  for (uint64_t i = 0; i < c->N; i++) {
    std::memcpy(&c->emis[i * c->S_pad], ratio + i * c->S, c->S * sizeof(double));
    if (e0) std::memcpy(&c->e0[i * c->S_pad], e0 + i * c->S, c->S * sizeof(double));
  }
  c->loge0.assign(loge0_sum, loge0_sum + c->N);
}
void simt_set_params(SimtCtx *c, const double *F, const double *alpha) {
  c->indF.assign(F, F + c->N);
  c->alpha.assign(alpha, alpha + c->N);
}
void simt_counters(unsigned long long *out) { out[0] = simt::launches(); out[1] = simt::switches(); }

static void copy_out(SimtCtx *c, double *ind_lkl_out, double *post_out) {
  if (ind_lkl_out) std::memcpy(ind_lkl_out, c->ind_lkl.data(), c->N * sizeof(double));
  if (post_out)
    for (uint64_t i = 0; i < c->N; i++) std::memcpy(post_out + i * c->S, &c->post[i * c->S_pad], c->S * sizeof(double));
}

// nfh_estep: returns the status word the kernels raised (kFlagNaN | kFlagFwBw)
int simt_estep(SimtCtx *c, double *ind_lkl_out, double *post_out) {
  c->status = 0;
  const EstepArgs a = estep_args(c);
  launch_estep(a, nullptr);
  copy_out(c, ind_lkl_out, post_out);
  return c->status;
}

// nfh_lkl_batch / nfh_estep_with_batch: the grouping of nfh_ctx.cu's lkl_batch_impl, then the launchers
int simt_lkl_batch(SimtCtx *c, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                   double *neg_lkl_out, int with_estep, double *ind_lkl_out, double *post_out) {
  if (n_req > c->N * kMaxPoints) return -1;
  uint32_t n_groups = 0;
  LklGroup *g = nullptr;
  for (uint64_t q = 0; q < n_req; q++) {
    if (ind[q] < 0 || (uint64_t) ind[q] >= c->N) return -1;
    if (std::isnan(F[q]) || std::isinf(F[q]) || std::isnan(alpha[q]) || std::isinf(alpha[q])) {
      if (with_estep) return -1;
      neg_lkl_out[q] = -1e15;
      continue;
    }
    if (!g || g->ind != ind[q] || g->npts == kMaxPoints) {
      if (n_groups >= c->N) return -1;
      g = &c->groups[n_groups++];
      g->ind = ind[q];
      g->npts = 0;
    }
    g->F[g->npts] = F[q]; g->alpha[g->npts] = alpha[q]; g->out[g->npts] = (int) q;
    g->npts++;
  }
  for (uint32_t k = 0; k < n_groups; k++) {
    LklGroup &gg = c->groups[k];
    gg.n_same = 1;
    while (gg.n_same < gg.npts && gg.alpha[gg.n_same] == gg.alpha[0]) gg.n_same++;
    gg.pad_ = 0;
  }
  if (n_groups == 0) return 0;
  LklArgs a;
  std::memset(&a, 0, sizeof a);
  a.emis = c->emis.data(); a.dist = c->dist.data(); a.loge0_sum = c->loge0.data();
  a.tile_dmax = c->tile_dmax.data(); a.tile_dsum = c->tile_dsum.data();
  a.groups = c->groups.data(); a.tile_prod = c->lkl_tile_prod.data(); a.neg_lkl = c->neg_lkl.data();
  a.n_rows = c->N; a.n_sites = c->S; a.site_block = c->S_pad; a.n_tiles = c->n_tiles; a.n_groups = n_groups;
  a.emit_chunk_prod = with_estep ? c->chunk_prod.data() : nullptr;
  a.emit_tile_prod = with_estep ? c->tile_prod.data() : nullptr;
  launch_lkl_batch(a, nullptr);
  c->status = 0;
  if (with_estep) {
    const EstepArgs ea = estep_args(c);
    launch_estep_tail(ea, nullptr);
    copy_out(c, ind_lkl_out, post_out);
  }
  for (uint32_t k = 0; k < n_groups; k++)
    for (int p = 0; p < c->groups[k].npts; p++) neg_lkl_out[c->groups[k].out[p]] = c->neg_lkl[c->groups[k].out[p]];
  return c->status;
}

// nfh_viterbi: path_out [N][S]
void simt_viterbi(SimtCtx *c, unsigned char *path_out) {
  const size_t n_chunks = (size_t) c->n_tiles * kScanThreads;
  ViterbiArgs a;
  std::memset(&a, 0, sizeof a);
  a.emis = c->emis.data(); a.e0 = c->e0.data(); a.dist = c->dist.data(); a.indF = c->indF.data(); a.alpha = c->alpha.data();
  a.work = c->vit_work.data(); a.work_stride = c->S_pad;
  a.chunk_prod = c->chunk_prod.data(); a.tile_prod = c->vit_tile_prod.data();
  a.tile_score = c->fwd.data();
  a.chunk_map = c->vit_maps.data();
  a.tile_map = c->vit_maps.data() + c->N * n_chunks;
  a.tile_state = a.tile_map + c->N * (size_t) c->n_tiles;
  a.final_state = c->vit_final.data();
  a.n_rows = c->N; a.n_rows_valid = c->N; a.n_sites = c->S; a.site_block = c->S_pad; a.n_tiles = c->n_tiles;
  launch_viterbi(a, nullptr);
  for (uint64_t i = 0; i < c->N; i++) std::memcpy(path_out + i * c->S, &c->vit_work[i * c->S_pad], c->S);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Self-tests of the emulator: deliberately broken kernels, so that the checks above are known not to be vacuous.
//   mode 0: correct - thread 0 issues the bulk copy, every thread waits on the mbarrier, then reads
//   mode 1: reads the tile BEFORE waiting                       -> poison (NaN) in the output
//   mode 2: no __syncthreads between a shared-memory write of the LAST thread and the read by thread 0
//           -> the result depends on the order threads run in (SIMT_ORDER)
//   mode 3: overwrites the source of a bulk store before tma_store_wait_read -> the wrong bytes arrive
// ---------------------------------------------------------------------------------------------------------------
namespace nfh {
static void selftest_kernel(int mode, const double *src, double *out) {
  alignas(128) __shared__ double tile[128];
  alignas(8) __shared__ uint64_t bar;
  __shared__ double mailbox;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); mailbox = -1.0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, sizeof tile);
    tma_load_1d(tile, src, sizeof tile, &bar);
  }
  double early = 0.0;
  if (mode == 1) early = tile[threadIdx.x];                 // BUG: the copy has not been waited for
  mbar_wait(&bar, 0);
  double v = mode == 1 ? early : tile[threadIdx.x];
  __syncthreads();
  if (mode == 2) {
    if (threadIdx.x == blockDim.x - 1) mailbox = 42.0;
    /* BUG: no __syncthreads() here */
    if (threadIdx.x == 0) v += mailbox;
  }
  __syncthreads();
  tile[threadIdx.x] = 2.0 * v;
  fence_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_1d(out, tile, sizeof tile);
    if (mode == 3) tile[5] = -7.0;                          // BUG: the store may not have read its source yet
    tma_store_wait_read();
  }
}
}  // namespace nfh

extern "C" void simt_selftest(int mode, const double *src /* [128], 16-byte aligned */, double *out /* [128] */) {
  simt::launch(dim3(1), dim3(128), 0, [&]() { nfh::selftest_kernel(mode, src, out); });
}

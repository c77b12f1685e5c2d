// nfh_kernels.h - host-visible launch interfaces of the CUDA kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfh {

constexpr int kMaxPoints = 5;   // objective points per individual per round: x, x -/+ eh_F, x -/+ eh_alpha (bfgs.cpp:22-43)

constexpr int kMaxRanks = 8;

// Destination windows of the other ranks (CUDA IPC mappings over NVLink).  When `direct` is set the
// producing kernel stores straight into the window of the rank that owns the data next, instead of a
// local send window that an all-to-all would move afterwards.
struct PeerWindows {
  double *base[kMaxRanks];
  uint64_t n_loc;      // individuals per rank (rows per source block)
  int rank;            // this rank = source block index in the destination window
  int direct;
};

struct TileProd {   // product of one tile's site matrices: [[a b][c d]] * 2^e * exp(l)
  double a, b, c, d, e, l;
};

struct LklGroup {   // objective requests of one individual sharing one read of its emissions
  int ind;
  int npts;
  int n_same;    // leading points whose alpha equals alpha[0] (they share one kappa per site)
  int pad_;
  double F[kMaxPoints];
  double alpha[kMaxPoints];
  int out[kMaxPoints];
};

// Device counters of the single-launch E-step (estep_fused) plus the host's record of how far they have run;
// owned by the context, zeroed at creation and never reset.
struct EstepFusedState {
  unsigned long long *d_ticket = nullptr;     // [1]
  unsigned long long *d_row_done = nullptr;   // [n_rows]
  unsigned *d_row_claim = nullptr;            // [n_rows]
  unsigned *d_row_ready = nullptr;            // [n_rows]
  unsigned long long ticket_base = 0;
  unsigned epoch = 0;
};

struct EstepArgs {
  const double *emis;        // emission ratio, blocked [n_ranks][n_rows][site_block]
  const double *dist;        // [n_ranks * site_block] Mb
  const double *tile_dmax, *tile_dsum;   // per tile: largest distance and sum of distances (kappa tiers, nfh_device.cuh)
  const double *indF, *alpha;
  const double *loge0_sum;   // [n_rows] sum over sites of log e0
  double4 *chunk_prod;       // [n_rows][n_tiles * 128] direction-only chunk products
  TileProd *tile_prod;       // [n_rows][n_tiles]
  double2 *fwd_carry, *bwd_carry;   // per tile
  double *post;              // blocked like emis
  PeerWindows post_peers;    // destination of posterior tile b = rank b's window
  double *ind_lkl;
  int *status;
  uint64_t n_rows, n_rows_valid, n_sites, site_block;
  uint32_t n_tiles;
  int sm_count;
  EstepFusedState *fused;    // NULL: three-launch E-step
};

struct LklArgs {
  const double *emis, *dist, *loge0_sum;
  const double *tile_dmax, *tile_dsum;
  const LklGroup *groups;
  TileProd *tile_prod;       // [n_groups][kMaxPoints][n_tiles]
  double *neg_lkl;
  // When set, point 0 of every group is also the E-step's forward product: its chunk products and tile
  // products go to the E-step buffers (EstepArgs::chunk_prod / tile_prod, row = the group's individual),
  // so that estep_chunk_products need not run (launch_estep_tail follows).
  double4 *emit_chunk_prod;
  TileProd *emit_tile_prod;
  uint64_t n_rows, n_sites, site_block;
  uint32_t n_tiles, n_groups;
};

struct FreqArgs {
  const double *gl0, *gl1, *gl2;  // linear GL planes [n_ind_pad][site_block]
  const double *post;             // [n_ind_pad][site_block] posterior of IBD state (or NULL: F_i = 0)
  double *freq;                   // [site_block]
  double *emis;                   // out: e1/e0  [n_ind_pad][site_block]
  double *e0;                     // out: e0 or NULL
  PeerWindows emis_peers;         // destination of row i = owner of individual i
  double *loge0_part;             // out: [gridDim.x][n_ind_pad] partial sums of log e0
  unsigned long long *pass_total; // += sum over sites of the est_maf passes each site needed (bench.py's flop count)
  uint64_t n_ind, n_ind_pad, site_block, sites_owned;
  int update_freq;                // 1: run est_maf; 0: keep freq
  double *acc_scratch;            // global home of the per-(CTA, warp|team, individual) log e0 accumulators, or NULL
  int use_maps;                   // tensor maps below are valid (freq_emission_warp prefetches through them)
  alignas(64) CUtensorMap maps[4];   // GL0, GL1, GL2, posterior planes as [n_ind_pad][site_block] tensors, see nfh_freq.cu
};

struct ViterbiArgs {
  const double *emis, *e0, *dist, *indF, *alpha;
  unsigned char *work;            // [n_rows][work_stride] back-pointer bytes, overwritten with the path
  double4 *chunk_prod, *tile_prod;   // (max,x) products: [n_rows][n_tiles*128], [n_rows][n_tiles]
  double2 *tile_score;            // [n_rows][n_tiles] scores entering each tile
  unsigned char *chunk_map, *tile_map, *tile_state;   // [n_rows][n_tiles*128], [n_rows][n_tiles] x2
  int *final_state;               // [n_rows]
  uint64_t n_rows, n_rows_valid, n_sites, site_block, work_stride;
  uint32_t n_tiles;
};

void launch_estep(const EstepArgs &a, cudaStream_t st);
void launch_estep_tail(const EstepArgs &a, cudaStream_t st);   // carries + apply, products already in place
void launch_lkl_batch(const LklArgs &a, cudaStream_t st);
// returns the grid size used (rows of loge0_part)
void launch_fill(double *dst, double value, size_t n, cudaStream_t st);
size_t freq_acc_scratch_bytes(uint64_t n_ind, uint64_t n_ind_pad, int sm_count);   // 0: never needed
bool freq_tensor_maps(FreqArgs &a, const double *post_plane);   // fills a.maps / a.use_maps for the shape of a.n_ind
unsigned freq_grid_size(const FreqArgs &a, int sm_count);
int launch_freq_emission(const FreqArgs &a, unsigned grid, cudaStream_t st);
void launch_reduce_loge0(const double *part, unsigned n_part, uint64_t n_ind_pad, double *out, cudaStream_t st);
void launch_gl_ingest(const double *staged_log_gl, uint64_t n_chunk_sites, uint64_t n_ind, uint64_t first_local_site,
                      uint64_t site_block, double *gl0, double *gl1, double *gl2, cudaStream_t st);
void launch_viterbi(const ViterbiArgs &a, cudaStream_t st);
void launch_geno_posterior(const double *gl0, const double *gl1, const double *gl2, const double *freq,
                           const char *path, uint64_t n_ind, uint64_t site_block, uint64_t path_stride,
                           uint64_t n_chunk, double *out, cudaStream_t st);
double launch_fp64_probe(cudaStream_t st, int sm_count);

}  // namespace nfh

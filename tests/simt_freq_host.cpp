// simt_freq_host.cpp - TEST INFRASTRUCTURE ONLY: second translation unit of the kernel harness (see
// simt_kernels_host.cpp): the frequency-EM kernels of nfh_freq.cu with their launchers under the SIMT emulator, behind C
// entry points that fill FreqArgs the way run_freq_family() of nfh_ctx.cu does for one rank.
#include <cstdint>
#include <cstring>
#include <vector>

#include "simt.h"

#include "nfh_freq.cu"

using namespace nfh;

// ---------------------------------------------------------------------------------------------------------------
// frequency side of one rank: run_freq_family() of nfh_ctx.cu - gl_ingest, freq_tensor_maps, freq_grid_size,
// launch_freq_emission (lane-group / hybrid / team / streaming kernels), launch_reduce_loge0
// ---------------------------------------------------------------------------------------------------------------
struct SimtFreq {
  uint64_t N = 0, S = 0, site_block = 0;
  std::vector<double> gl[3], post, freq, emis, e0, loge0_part, loge0_sum, acc, staged;
  unsigned long long passes = 0;
  int sm_count = 4;
};

extern "C" {

// gl_norm_site_major: [S][N][3] normalised log GL, the layout nfh_upload_gl takes
SimtFreq *simt_freq_create(uint64_t N, uint64_t S, const double *gl_norm_site_major, int sm_count) {
  SimtFreq *c = new SimtFreq;
  c->N = N; c->S = S; c->sm_count = sm_count;
  c->site_block = (S + kTile - 1) / kTile * kTile;
  const size_t plane = (size_t) N * c->site_block;
  for (int g = 0; g < 3; g++) c->gl[g].assign(plane, 0.0);
  c->post.assign(plane, 0.0); c->emis.assign(plane, 1.0); c->e0.assign(plane, 1.0);
  c->freq.assign(c->site_block, 0.0);
  c->loge0_part.assign((size_t) std::max(sm_count * 4, 64) * N, 0.0);   // the streaming path uses 64 row chunks
  c->loge0_sum.assign(N, 0.0);
  if (const size_t bytes = freq_acc_scratch_bytes(N, N, sm_count)) c->acc.assign(bytes / sizeof(double), 0.0);
  c->staged.assign(gl_norm_site_major, gl_norm_site_major + S * N * 3);
  launch_gl_ingest(c->staged.data(), S, N, 0, c->site_block, c->gl[0].data(), c->gl[1].data(), c->gl[2].data(), nullptr);
  return c;
}
void simt_freq_destroy(SimtFreq *c) { delete c; }
void simt_freq_set(SimtFreq *c, const double *post /* [N][S] or NULL */, const double *freq /* [S] or NULL */) {
  if (post)
    for (uint64_t i = 0; i < c->N; i++) std::memcpy(&c->post[i * c->site_block], post + i * c->S, c->S * sizeof(double));
  if (freq) std::memcpy(c->freq.data(), freq, c->S * sizeof(double));
}
void simt_freq_get_gl(SimtFreq *c, double *out /* [3][N][S] linear */) {
  for (int g = 0; g < 3; g++)
    for (uint64_t i = 0; i < c->N; i++)
      std::memcpy(out + ((size_t) g * c->N + i) * c->S, &c->gl[g][i * c->site_block], c->S * sizeof(double));
}

// returns launch_freq_emission's code (1 register / hybrid / team kernels, 2 streaming path); *used_maps: tile prefetch
int simt_freq_run(SimtFreq *c, int update, int zero_post, int with_e0, int want_maps, int *used_maps, double *freq_out,
                  double *ratio_out, double *e0_out, double *loge0_out, unsigned long long *passes_out) {
  FreqArgs a;
  std::memset(&a, 0, sizeof a);
  a.gl0 = c->gl[0].data(); a.gl1 = c->gl[1].data(); a.gl2 = c->gl[2].data();
  a.post = zero_post ? nullptr : c->post.data();
  a.freq = c->freq.data(); a.emis = c->emis.data(); a.e0 = with_e0 ? c->e0.data() : nullptr;
  a.loge0_part = c->loge0_part.data();
  c->passes = 0;
  a.pass_total = &c->passes;
  a.acc_scratch = c->acc.empty() ? nullptr : c->acc.data();
  a.emis_peers.direct = 0; a.emis_peers.rank = 0; a.emis_peers.n_loc = c->N;
  a.n_ind = c->N; a.n_ind_pad = c->N; a.site_block = c->site_block; a.sites_owned = c->S;
  a.update_freq = update;
  if (want_maps) freq_tensor_maps(a, c->post.data());
  if (used_maps) *used_maps = a.use_maps;
  const unsigned grid = freq_grid_size(a, c->sm_count);
  if ((size_t) grid * c->N > c->loge0_part.size()) return -1;
  const int n = launch_freq_emission(a, grid, nullptr);
  launch_reduce_loge0(c->loge0_part.data(), grid, c->N, c->loge0_sum.data(), nullptr);
  if (freq_out) std::memcpy(freq_out, c->freq.data(), c->S * sizeof(double));
  for (uint64_t i = 0; i < c->N; i++) {
    if (ratio_out) std::memcpy(ratio_out + i * c->S, &c->emis[i * c->site_block], c->S * sizeof(double));
    if (e0_out && with_e0) std::memcpy(e0_out + i * c->S, &c->e0[i * c->site_block], c->S * sizeof(double));
  }
  if (loge0_out) std::memcpy(loge0_out, c->loge0_sum.data(), c->N * sizeof(double));
  if (passes_out) *passes_out = c->passes;
  return n;
}

// The streaming path of launch_freq_emission() (no register shape fits: more than 4,096 individuals), launched
// directly so that a test can run it at a size an emulator can afford: freq_emission_stream + loge0_rowsum with
// `chunks` row chunks, then reduce_loge0.
int simt_freq_run_stream(SimtFreq *c, int update, unsigned chunks, double *freq_out, double *ratio_out, double *loge0_out,
                         unsigned long long *passes_out) {
  FreqArgs a;
  std::memset(&a, 0, sizeof a);
  a.gl0 = c->gl[0].data(); a.gl1 = c->gl[1].data(); a.gl2 = c->gl[2].data();
  a.post = c->post.data(); a.freq = c->freq.data(); a.emis = c->emis.data(); a.e0 = nullptr;
  a.loge0_part = c->loge0_part.data();
  c->passes = 0;
  a.pass_total = &c->passes;
  a.emis_peers.direct = 0; a.emis_peers.n_loc = c->N;
  a.n_ind = c->N; a.n_ind_pad = c->N; a.site_block = c->site_block; a.sites_owned = c->S;
  a.update_freq = update;
  if ((size_t) chunks * c->N > c->loge0_part.size()) return -1;
  const unsigned blocks = (unsigned) ((a.sites_owned + kFreqThreads - 1) / kFreqThreads);
  simt::launch(dim3(blocks ? blocks : 1), dim3(kFreqThreads), 0, [&]() { freq_emission_stream(a); });
  simt::launch(dim3(chunks, (unsigned) a.n_ind), dim3(256), 0, [&]() { loge0_rowsum(a, chunks); });
  launch_reduce_loge0(c->loge0_part.data(), chunks, c->N, c->loge0_sum.data(), nullptr);
  std::memcpy(freq_out, c->freq.data(), c->S * sizeof(double));
  for (uint64_t i = 0; i < c->N; i++) std::memcpy(ratio_out + i * c->S, &c->emis[i * c->site_block], c->S * sizeof(double));
  std::memcpy(loge0_out, c->loge0_sum.data(), c->N * sizeof(double));
  *passes_out = c->passes;
  return 2;
}

// nfh_geno_posterior: path [N][S] (0/1) -> out [S][N][3]
void simt_geno_posterior(SimtFreq *c, const char *path, double *out) {
  launch_geno_posterior(c->gl[0].data(), c->gl[1].data(), c->gl[2].data(), c->freq.data(), path, c->N, c->site_block,
                        c->S, c->S, out, nullptr);
}

}  // extern "C"

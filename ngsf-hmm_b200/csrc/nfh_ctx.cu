// nfh_ctx.cu - the C ABI (include/ngsfhmm_b200.h): context, device memory
// layout, pinned staging and kernel orchestration.  No CPU fallback: every
// entry point needs a CUDA device and fails loudly without one.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ngsfhmm_b200.h"
#include "nfh_device.cuh"
#include "nfh_kernels.h"
#include "nfh_schedule.h"

using namespace nfh;

namespace {

constexpr size_t kStageBytes = 64u << 20;   // pinned + device staging for GL ingest
constexpr int kFamilies = 8;
enum Family { kFamEstep = 0, kFamLkl = 1, kFamFreq = 2, kFamViterbi = 3, kFamEmission = 4, kFamIngest = 5 };

struct TimedSpan {
  int family;
  cudaEvent_t t0, t1;
};

}  // namespace

struct nfh_ctx {
  int device = 0, n_ranks = 1, rank = 0, sm_count = 148;
  uint64_t n_ind_total = 0, n_sites = 0;
  uint64_t n_loc = 0, n_owned = 0, ind_begin = 0;
  uint64_t site_block = 0, site_begin = 0, sites_owned = 0;
  uint64_t n_ind_pad = 0, n_sites_pad = 0;
  uint32_t n_tiles = 0;
  cudaStream_t stream = nullptr;

  // recursion side (this rank's individuals, all sites, site-blocked layout)
  double *tile_dmax = nullptr, *tile_dsum = nullptr;   // per tile of the distance vector, see nfh_upload_pos_dist
  double *dist = nullptr, *emis_recv = nullptr, *post_send = nullptr, *e0_recv = nullptr;
  double *indF = nullptr, *alpha = nullptr, *ind_lkl = nullptr;
  double4 *chunk_prod = nullptr;
  TileProd *tile_prod = nullptr, *lkl_tile_prod = nullptr;
  double2 *fwd_carry = nullptr, *bwd_carry = nullptr;
  EstepFusedState fused;
  LklGroup *groups = nullptr;
  double *neg_lkl = nullptr;
  unsigned char *vit_work = nullptr, *vit_maps = nullptr;
  double4 *vit_tile_prod = nullptr;
  int *vit_final = nullptr;

  // frequency side (all individuals, this rank's site block)
  double *gl[3] = {nullptr, nullptr, nullptr};
  double *post_recv = nullptr, *emis_send = nullptr, *e0_send = nullptr, *freq = nullptr;
  double *loge0_part = nullptr, *loge0_sum = nullptr;
  unsigned long long *freq_passes = nullptr;   // device counter, see FreqArgs::pass_total
  double *freq_acc = nullptr;                  // FreqArgs::acc_scratch (large n_ind only)
  unsigned loge0_rows = 0;

  // peer windows (CUDA IPC), see nfh_peer_*
  double *peer_post[kMaxRanks] = {nullptr}, *peer_emis[kMaxRanks] = {nullptr};
  bool peer_post_ipc[kMaxRanks] = {false}, peer_emis_ipc[kMaxRanks] = {false};   // mapped by cudaIpcOpenMemHandle (to be closed)
  bool peer_direct = false;        // emission ratios go straight to the owners of the individuals
  bool peer_post_direct = false;   // posterior tiles go straight to the owners of the site blocks
  bool post_in_peers = false;      // where the last E-step put them (nfh_get_posterior reads there)

  int *status = nullptr;
  // pinned host scratch
  double *h_small = nullptr;       // 8 * max(n_loc, 16) doubles + requests
  size_t h_small_doubles = 0;
  LklGroup *h_groups = nullptr;
  std::vector<double> h_indF, h_alpha;   // host mirror of the parameters last set (nfh_estep_with_batch checks against it)
  int *h_status = nullptr;
  void *h_stage = nullptr, *d_stage = nullptr;

  uint64_t launches = 0;
  bool timing = false;
  std::vector<TimedSpan> spans;
  std::vector<cudaEvent_t> free_events;
  double fam_ms[kFamilies] = {0};
  uint64_t fam_launches[kFamilies] = {0};
  std::string err;
};

static thread_local std::string g_create_err;

#define NFH_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return e_ == cudaErrorMemoryAllocation ? NFH_ERR_NOMEM : NFH_ERR_CUDA;                    \
    }                                                                                           \
  } while (0)

static int fail(nfh_ctx *ctx, int code, const char *msg) {
  ctx->err = msg;
  return code;
}

static cudaEvent_t get_event(nfh_ctx *ctx) {
  if (!ctx->free_events.empty()) {
    cudaEvent_t e = ctx->free_events.back();
    ctx->free_events.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct FamilyScope {   // brackets a kernel family with events when timing is on
  nfh_ctx *ctx;
  int fam;
  cudaEvent_t t0 = nullptr;
  FamilyScope(nfh_ctx *c, int f, int n_launches) : ctx(c), fam(f) {
    ctx->launches += n_launches;
    ctx->fam_launches[f] += n_launches;
    if (ctx->timing) {
      t0 = get_event(ctx);
      cudaEventRecord(t0, ctx->stream);
    }
  }
  ~FamilyScope() {
    if (ctx->timing) {
      cudaEvent_t t1 = get_event(ctx);
      cudaEventRecord(t1, ctx->stream);
      ctx->spans.push_back({fam, t0, t1});
    }
  }
};

static int check_status(nfh_ctx *ctx) {
  // status word was copied to h_status on the stream before the last sync
  int s = *ctx->h_status;
  if (s & kFlagNaN) return fail(ctx, NFH_ERR_NAN, "invalid Lkl found! (NaN in a recursion or posterior)");
  if (s & kFlagFwBw) return fail(ctx, NFH_ERR_FWBW, "Fw and Bw lkl do not match!");
  return NFH_OK;
}

static int sync_and_check(nfh_ctx *ctx) {
  NFH_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  int rc = check_status(ctx);
  if (rc != NFH_OK) {
    *ctx->h_status = 0;
    cudaMemsetAsync(ctx->status, 0, sizeof(int), ctx->stream);
  }
  return rc;
}

extern "C" {

const char *nfh_strerror(int status) {
  switch (status) {
    case NFH_OK: return "ok";
    case NFH_ERR_CUDA: return "CUDA runtime error";
    case NFH_ERR_ARG: return "invalid argument";
    case NFH_ERR_NAN: return "invalid Lkl found!";
    case NFH_ERR_FWBW: return "Fw and Bw lkl do not match!";
    case NFH_ERR_NOMEM: return "out of device memory";
    case NFH_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    default: return "unknown status";
  }
}

const char *nfh_last_error(const nfh_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

const char *nfh_build_info(void) { return "ngsfhmm_b200 sm_100a fp64 (no CPU fallback)"; }

uint64_t nfh_kernel_launches(const nfh_ctx *ctx) { return ctx->launches; }

uint64_t nfh_estep_schedule_item(uint32_t n_rows, uint32_t n_tiles, uint32_t wave_rows, uint32_t lookahead,
                                 uint64_t ticket, uint32_t item_out[3]) {
  const EstepSchedule s = make_schedule(n_rows, n_tiles, wave_rows, lookahead);
  if (ticket < s.total && item_out) {
    const EstepItem it = decode_ticket(s, (uint32_t) ticket);
    item_out[0] = it.apply; item_out[1] = it.row; item_out[2] = it.tile;
  }
  return s.total;
}

int nfh_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int nfh_ctx_create(nfh_ctx **out, int device, uint64_t n_ind_total, uint64_t n_sites, int n_ranks, int rank) {
  if (!out || n_ind_total == 0 || n_sites == 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
    g_create_err = "nfh_ctx_create: bad geometry";
    return NFH_ERR_ARG;
  }
  if (n_ind_total > 65535) {   // individuals index grid.y of several kernels; the reference's print_iter has the same limit (uint16_t, EM.cpp:306)
    g_create_err = "nfh_ctx_create: at most 65535 individuals";
    return NFH_ERR_ARG;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) {
    g_create_err = "nfh_ctx_create: no usable CUDA device; the hot path has no CPU fallback";
    return NFH_ERR_NO_DEVICE;
  }
  nfh_ctx *ctx = new nfh_ctx;
  ctx->device = device; ctx->n_ranks = n_ranks; ctx->rank = rank;
  ctx->n_ind_total = n_ind_total; ctx->n_sites = n_sites;
  ctx->n_loc = (n_ind_total + n_ranks - 1) / n_ranks;
  ctx->ind_begin = (uint64_t) rank * ctx->n_loc;
  ctx->n_owned = ctx->ind_begin >= n_ind_total ? 0 : std::min(ctx->n_loc, n_ind_total - ctx->ind_begin);
  uint64_t per = (n_sites + n_ranks - 1) / n_ranks;
  ctx->site_block = ((per + kTile - 1) / kTile) * kTile;
  ctx->site_begin = (uint64_t) rank * ctx->site_block;
  ctx->sites_owned = ctx->site_begin >= n_sites ? 0 : std::min(ctx->site_block, n_sites - ctx->site_begin);
  ctx->n_ind_pad = ctx->n_loc * n_ranks;
  ctx->n_sites_pad = ctx->site_block * n_ranks;
  ctx->n_tiles = (uint32_t) ((n_sites + kTile - 1) / kTile);
  *out = ctx;

  auto bail = [&](int rc) { g_create_err = ctx->err; nfh_ctx_destroy(ctx); *out = nullptr; return rc; };
  auto alloc = [&](void **p, size_t bytes, bool zero) -> int {
    NFH_CUDA(cudaMalloc(p, bytes ? bytes : 8));
    if (zero) NFH_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 8, ctx->stream));
    return NFH_OK;
  };
  int rc;
#define NFH_TRY(x) do { rc = (x); if (rc != NFH_OK) return bail(rc); } while (0)
  auto setup = [&]() -> int {
    NFH_CUDA(cudaSetDevice(device));
    NFH_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    NFH_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    return NFH_OK;
  };
  NFH_TRY(setup());

  const size_t plane_rec = (size_t) ctx->n_ranks * ctx->n_loc * ctx->site_block * sizeof(double);
  const size_t plane_frq = (size_t) ctx->n_ind_pad * ctx->site_block * sizeof(double);   // same number
  NFH_TRY(alloc((void **) &ctx->dist, ctx->n_sites_pad * sizeof(double), true));
  NFH_TRY(alloc((void **) &ctx->tile_dmax, ctx->n_tiles * sizeof(double), true));
  NFH_TRY(alloc((void **) &ctx->tile_dsum, ctx->n_tiles * sizeof(double), true));
  // emission-ratio windows start as 1.0 everywhere and the kernels only ever write real sites, so the
  // padding of the last tile stays the identity of the recursions (r = 1 with d = 0)
  NFH_TRY(alloc((void **) &ctx->emis_recv, plane_rec, false));
  launch_fill(ctx->emis_recv, 1.0, plane_rec / sizeof(double), ctx->stream);
  NFH_TRY(alloc((void **) &ctx->post_send, plane_rec, true));
  NFH_TRY(alloc((void **) &ctx->indF, ctx->n_loc * sizeof(double), true));
  NFH_TRY(alloc((void **) &ctx->alpha, ctx->n_loc * sizeof(double), true));
  NFH_TRY(alloc((void **) &ctx->ind_lkl, ctx->n_loc * sizeof(double), true));
  const size_t n_chunks = (size_t) ctx->n_tiles * kScanThreads;
  NFH_TRY(alloc((void **) &ctx->chunk_prod, ctx->n_loc * n_chunks * sizeof(double4), false));
  NFH_TRY(alloc((void **) &ctx->tile_prod, ctx->n_loc * ctx->n_tiles * sizeof(TileProd), false));
  NFH_TRY(alloc((void **) &ctx->lkl_tile_prod, ctx->n_loc * kMaxPoints * ctx->n_tiles * sizeof(TileProd), false));
  NFH_TRY(alloc((void **) &ctx->fwd_carry, ctx->n_loc * ctx->n_tiles * sizeof(double2), false));
  NFH_TRY(alloc((void **) &ctx->bwd_carry, ctx->n_loc * ctx->n_tiles * sizeof(double2), false));
  NFH_TRY(alloc((void **) &ctx->fused.d_ticket, sizeof(unsigned long long), true));
  NFH_TRY(alloc((void **) &ctx->fused.d_row_done, ctx->n_loc * sizeof(unsigned long long), true));
  NFH_TRY(alloc((void **) &ctx->fused.d_row_claim, ctx->n_loc * sizeof(unsigned), true));
  NFH_TRY(alloc((void **) &ctx->fused.d_row_ready, ctx->n_loc * sizeof(unsigned), true));
  NFH_TRY(alloc((void **) &ctx->groups, ctx->n_loc * sizeof(LklGroup), false));
  NFH_TRY(alloc((void **) &ctx->neg_lkl, ctx->n_loc * kMaxPoints * sizeof(double), false));
  for (int g = 0; g < 3; g++) NFH_TRY(alloc((void **) &ctx->gl[g], plane_frq, true));
  if (n_ranks == 1) {
    ctx->post_recv = ctx->post_send;
    ctx->emis_send = ctx->emis_recv;
  } else {
    NFH_TRY(alloc((void **) &ctx->post_recv, plane_frq, true));
    NFH_TRY(alloc((void **) &ctx->emis_send, plane_frq, false));
    launch_fill(ctx->emis_send, 1.0, plane_frq / sizeof(double), ctx->stream);
  }
  NFH_TRY(alloc((void **) &ctx->freq, ctx->site_block * sizeof(double), true));
  ctx->loge0_rows = (unsigned) ctx->sm_count * 4u;
  NFH_TRY(alloc((void **) &ctx->loge0_part, (size_t) ctx->loge0_rows * ctx->n_ind_pad * sizeof(double), true));
  NFH_TRY(alloc((void **) &ctx->loge0_sum, ctx->n_ind_pad * sizeof(double), true));
  NFH_TRY(alloc((void **) &ctx->status, sizeof(int), true));
  NFH_TRY(alloc((void **) &ctx->freq_passes, sizeof(unsigned long long), true));
  if (const size_t acc_bytes = freq_acc_scratch_bytes(ctx->n_ind_total, ctx->n_ind_pad, ctx->sm_count))
    NFH_TRY(alloc((void **) &ctx->freq_acc, acc_bytes, false));
  NFH_TRY(alloc(&ctx->d_stage, kStageBytes, false));
  auto pinned = [&]() -> int {
    ctx->h_small_doubles = 8 * std::max<uint64_t>(ctx->n_loc * kMaxPoints, 64);
    NFH_CUDA(cudaMallocHost((void **) &ctx->h_small, ctx->h_small_doubles * sizeof(double)));
    NFH_CUDA(cudaMallocHost((void **) &ctx->h_groups, ctx->n_loc * sizeof(LklGroup)));
    NFH_CUDA(cudaMallocHost((void **) &ctx->h_status, sizeof(int)));
    NFH_CUDA(cudaMallocHost(&ctx->h_stage, kStageBytes));
    *ctx->h_status = 0;
    NFH_CUDA(cudaStreamSynchronize(ctx->stream));
    return NFH_OK;
  };
  NFH_TRY(pinned());
#undef NFH_TRY
  return NFH_OK;
}

void nfh_ctx_destroy(nfh_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  void *dev[] = {ctx->tile_dmax, ctx->tile_dsum, ctx->dist, ctx->emis_recv, ctx->post_send, ctx->e0_recv, ctx->indF, ctx->alpha, ctx->ind_lkl,
                 ctx->chunk_prod, ctx->tile_prod, ctx->lkl_tile_prod, ctx->fwd_carry, ctx->bwd_carry, ctx->groups, ctx->neg_lkl,
                 ctx->vit_work, ctx->vit_maps, ctx->vit_tile_prod, ctx->vit_final, ctx->gl[0], ctx->gl[1], ctx->gl[2], ctx->freq, ctx->loge0_part, ctx->loge0_sum,
                 ctx->status, ctx->freq_passes, ctx->freq_acc, ctx->d_stage, ctx->fused.d_ticket, ctx->fused.d_row_done,
                 ctx->fused.d_row_claim, ctx->fused.d_row_ready};
  for (void *p : dev) if (p) cudaFree(p);
  for (int r = 0; r < kMaxRanks; r++) {
    if (r == ctx->rank) continue;
    if (ctx->peer_post[r] && ctx->peer_post_ipc[r]) cudaIpcCloseMemHandle(ctx->peer_post[r]);
    if (ctx->peer_emis[r] && ctx->peer_emis_ipc[r]) cudaIpcCloseMemHandle(ctx->peer_emis[r]);
  }
  if (ctx->n_ranks > 1) {
    if (ctx->post_recv) cudaFree(ctx->post_recv);
    if (ctx->emis_send) cudaFree(ctx->emis_send);
    if (ctx->e0_send) cudaFree(ctx->e0_send);
  }
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  if (ctx->h_groups) cudaFreeHost(ctx->h_groups);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  for (auto &s : ctx->spans) { cudaEventDestroy(s.t0); cudaEventDestroy(s.t1); }
  for (auto e : ctx->free_events) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

uint64_t nfh_n_ind_local(const nfh_ctx *ctx) { return ctx->n_loc; }
uint64_t nfh_n_ind_owned(const nfh_ctx *ctx) { return ctx->n_owned; }
uint64_t nfh_ind_begin(const nfh_ctx *ctx) { return ctx->ind_begin; }
uint64_t nfh_site_block(const nfh_ctx *ctx) { return ctx->site_block; }
uint64_t nfh_site_begin(const nfh_ctx *ctx) { return ctx->site_begin; }
uint64_t nfh_sites_owned(const nfh_ctx *ctx) { return ctx->sites_owned; }
void *nfh_stream(const nfh_ctx *ctx) { return (void *) ctx->stream; }

int nfh_upload_gl(nfh_ctx *ctx, const double *log_gl, uint64_t first_site, uint64_t n) {
  if (!log_gl || first_site < ctx->site_begin || first_site + n > ctx->site_begin + ctx->sites_owned)
    return fail(ctx, NFH_ERR_ARG, "nfh_upload_gl: sites outside this rank's block");
  NFH_CUDA(cudaSetDevice(ctx->device));
  const uint64_t N = ctx->n_ind_total;
  const uint64_t per_site = N * 3 * sizeof(double);
  uint64_t chunk = kStageBytes / per_site;
  if (chunk == 0) return fail(ctx, NFH_ERR_ARG, "nfh_upload_gl: one site exceeds the staging buffer");
  // where the caller's array lives: page-locked host memory is copied from directly, pageable host memory goes
  // through the pinned staging buffer, device memory (a generator that already ran on the GPU) is ingested in place
  cudaPointerAttributes attr;
  const bool known = cudaPointerGetAttributes(&attr, log_gl) == cudaSuccess;
  cudaGetLastError();
  const bool src_pinned = known && attr.type == cudaMemoryTypeHost;
  const bool src_device = known && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  if (src_device && attr.type == cudaMemoryTypeDevice && attr.device != ctx->device)
    return fail(ctx, NFH_ERR_ARG, "nfh_upload_gl: device array lives on another GPU");
  for (uint64_t done = 0; done < n; done += chunk) {
    const uint64_t m = std::min(chunk, n - done);
    const double *src = log_gl + done * N * 3;
    const double *staged = (const double *) ctx->d_stage;
    if (src_device) {
      staged = src;
    } else {
      if (!src_pinned) {
        NFH_CUDA(cudaStreamSynchronize(ctx->stream));   // staging buffer is reused
        memcpy(ctx->h_stage, src, m * per_site);
        src = (const double *) ctx->h_stage;
      }
      NFH_CUDA(cudaMemcpyAsync(ctx->d_stage, src, m * per_site, cudaMemcpyHostToDevice, ctx->stream));
    }
    {
      FamilyScope fs(ctx, kFamIngest, 1);
      launch_gl_ingest(staged, m, N, first_site - ctx->site_begin + done, ctx->site_block, ctx->gl[0], ctx->gl[1],
                       ctx->gl[2], ctx->stream);
    }
    NFH_CUDA(cudaGetLastError());
  }
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_upload_pos_dist(nfh_ctx *ctx, const double *dist_mb) {
  if (!dist_mb) return fail(ctx, NFH_ERR_ARG, "nfh_upload_pos_dist: null");
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaMemcpyAsync(ctx->dist, dist_mb, ctx->n_sites * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // per tile: the largest distance (NaN poisons it, +inf stays +inf) and the sum, which pick the kappa
  // tier of every (individual, tile) and give the tile's scalar transition factor without a per-site sum
  std::vector<double> tmax(ctx->n_tiles), tsum(ctx->n_tiles);
  for (uint32_t t = 0; t < ctx->n_tiles; t++) {
    const uint64_t lo = (uint64_t) t * kTile, hi = std::min(ctx->n_sites, lo + kTile);
    double mx = 0.0, sum = 0.0;
    bool nan = false;
    for (uint64_t s = lo; s < hi; s++) {
      const double d = dist_mb[s];
      nan |= d != d;
      mx = d > mx ? d : mx;
      sum += d;
    }
    tmax[t] = nan ? std::nan("") : mx;
    tsum[t] = sum;
  }
  NFH_CUDA(cudaMemcpyAsync(ctx->tile_dmax, tmax.data(), ctx->n_tiles * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NFH_CUDA(cudaMemcpyAsync(ctx->tile_dsum, tsum.data(), ctx->n_tiles * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_set_freq(nfh_ctx *ctx, const double *freq) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaMemcpyAsync(ctx->freq, freq, ctx->sites_owned * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_get_freq(nfh_ctx *ctx, double *freq) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaMemcpyAsync(freq, ctx->freq, ctx->sites_owned * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_set_ind_params(nfh_ctx *ctx, const double *indF, const double *alpha) {
  ctx->h_indF.assign(indF, indF + ctx->n_owned);
  ctx->h_alpha.assign(alpha, alpha + ctx->n_owned);
  NFH_CUDA(cudaSetDevice(ctx->device));
  double *h = ctx->h_small;
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(h, indF, ctx->n_owned * sizeof(double));
  memcpy(h + ctx->n_loc, alpha, ctx->n_owned * sizeof(double));
  NFH_CUDA(cudaMemcpyAsync(ctx->indF, h, ctx->n_owned * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NFH_CUDA(cudaMemcpyAsync(ctx->alpha, h + ctx->n_loc, ctx->n_owned * sizeof(double), cudaMemcpyHostToDevice,
                           ctx->stream));
  return NFH_OK;
}

static int run_freq_family(nfh_ctx *ctx, int family, int update, bool zero_post, bool with_e0) {
  if (with_e0 && !ctx->e0_send) {
    const size_t plane = (size_t) ctx->n_ind_pad * ctx->site_block * sizeof(double);
    NFH_CUDA(cudaMalloc((void **) &ctx->e0_send, plane));
    if (ctx->n_ranks == 1) ctx->e0_recv = ctx->e0_send;
    else NFH_CUDA(cudaMalloc((void **) &ctx->e0_recv, plane));
  }
  FreqArgs a;
  a.gl0 = ctx->gl[0]; a.gl1 = ctx->gl[1]; a.gl2 = ctx->gl[2];
  a.post = zero_post ? nullptr : ctx->post_recv;
  a.freq = ctx->freq; a.emis = ctx->emis_send; a.e0 = with_e0 ? ctx->e0_send : nullptr;
  a.loge0_part = ctx->loge0_part;
  a.pass_total = ctx->freq_passes;
  a.acc_scratch = ctx->freq_acc;
  a.emis_peers.direct = ctx->peer_direct ? 1 : 0;
  a.emis_peers.rank = ctx->rank; a.emis_peers.n_loc = ctx->n_loc;
  for (int r = 0; r < kMaxRanks; r++) a.emis_peers.base[r] = ctx->peer_emis[r];
  a.n_ind = ctx->n_ind_total; a.n_ind_pad = ctx->n_ind_pad;
  a.site_block = ctx->site_block; a.sites_owned = ctx->sites_owned;
  a.update_freq = update;
  freq_tensor_maps(a, ctx->post_recv);
  if (ctx->sites_owned == 0) {
    NFH_CUDA(cudaMemsetAsync(ctx->loge0_sum, 0, ctx->n_ind_pad * sizeof(double), ctx->stream));
    return NFH_OK;
  }
  unsigned grid = freq_grid_size(a, ctx->sm_count);
  {
    FamilyScope fs(ctx, family, 0);
    int n = launch_freq_emission(a, grid, ctx->stream);
    launch_reduce_loge0(ctx->loge0_part, grid, ctx->n_ind_pad, ctx->loge0_sum, ctx->stream);
    ctx->launches += n + 1;
    ctx->fam_launches[family] += n + 1;
  }
  NFH_CUDA(cudaGetLastError());
  return NFH_OK;
}

int nfh_emission_refresh(nfh_ctx *ctx, int with_e0) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  return run_freq_family(ctx, kFamEmission, 0, true, with_e0 != 0);
}

int nfh_freq_update(nfh_ctx *ctx, int method, int posterior_is_zero, double *freq_out) {
  if (method != 0 && method != 1)
    return fail(ctx, NFH_ERR_ARG, "wrong MAF estimation method! (only --freq_est 0/1 run in the reference)");
  NFH_CUDA(cudaSetDevice(ctx->device));
  int rc = run_freq_family(ctx, method ? kFamFreq : kFamEmission, method, posterior_is_zero != 0, false);
  if (rc != NFH_OK) return rc;
  if (freq_out) {
    NFH_CUDA(cudaMemcpyAsync(freq_out, ctx->freq, ctx->sites_owned * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return NFH_OK;
}

static EstepArgs estep_args(nfh_ctx *ctx) {
  EstepArgs a;
  a.emis = ctx->emis_recv; a.dist = ctx->dist; a.indF = ctx->indF; a.alpha = ctx->alpha;
  a.tile_dmax = ctx->tile_dmax; a.tile_dsum = ctx->tile_dsum;
  a.loge0_sum = ctx->loge0_sum + ctx->ind_begin;
  a.chunk_prod = ctx->chunk_prod; a.tile_prod = ctx->tile_prod; a.fwd_carry = ctx->fwd_carry; a.bwd_carry = ctx->bwd_carry;
  a.post = ctx->post_send; a.ind_lkl = ctx->ind_lkl; a.status = ctx->status;
  a.post_peers.direct = ctx->peer_post_direct ? 1 : 0;
  ctx->post_in_peers = ctx->peer_post_direct;
  a.post_peers.rank = ctx->rank; a.post_peers.n_loc = ctx->n_loc;
  for (int r = 0; r < kMaxRanks; r++) a.post_peers.base[r] = ctx->peer_post[r];
  a.n_rows = ctx->n_loc; a.n_rows_valid = ctx->n_owned; a.n_sites = ctx->n_sites;
  a.site_block = ctx->site_block; a.n_tiles = ctx->n_tiles;
  a.sm_count = ctx->sm_count;
  a.fused = &ctx->fused;
  return a;
}

int nfh_estep(nfh_ctx *ctx, double *ind_lkl_out) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  if (ctx->n_owned) {
    const EstepArgs a = estep_args(ctx);
    {
      FamilyScope fs(ctx, kFamEstep, 3);
      launch_estep(a, ctx->stream);
    }
    NFH_CUDA(cudaGetLastError());
  }
  if (ind_lkl_out) {
    NFH_CUDA(cudaMemcpyAsync(ctx->h_small, ctx->ind_lkl, ctx->n_owned * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    int rc = sync_and_check(ctx);
    memcpy(ind_lkl_out, ctx->h_small, ctx->n_owned * sizeof(double));
    return rc;
  }
  return NFH_OK;
}

// Shared body of nfh_lkl_batch and nfh_estep_with_batch.
static int lkl_batch_impl(nfh_ctx *ctx, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                          double *neg_lkl_out, bool with_estep, double *ind_lkl_out) {
  if (n_req > ctx->n_loc * kMaxPoints) return fail(ctx, NFH_ERR_ARG, "nfh_lkl_batch: too many requests in one call");
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));   // pinned mirrors are reused
  uint32_t n_groups = 0;
  LklGroup *g = nullptr;
  for (uint64_t q = 0; q < n_req; q++) {
    if (ind[q] < 0 || (uint64_t) ind[q] >= ctx->n_owned) return fail(ctx, NFH_ERR_ARG, "nfh_lkl_batch: bad individual");
    if (std::isnan(F[q]) || std::isinf(F[q]) || std::isnan(alpha[q]) || std::isinf(alpha[q])) {
      if (with_estep) return fail(ctx, NFH_ERR_ARG, "nfh_estep_with_batch: non-finite parameters");
      neg_lkl_out[q] = -1e15;   // EM.cpp:454-456
      continue;
    }
    if (!g || g->ind != ind[q] || g->npts == kMaxPoints) {
      if (n_groups >= ctx->n_loc)
        return fail(ctx, NFH_ERR_ARG, "nfh_lkl_batch: keep requests of one individual adjacent (<= 5 per individual)");
      g = &ctx->h_groups[n_groups++];
      g->ind = ind[q];
      g->npts = 0;
    }
    g->F[g->npts] = F[q];
    g->alpha[g->npts] = alpha[q];
    g->out[g->npts] = (int) q;
    g->npts++;
  }
  for (uint32_t k = 0; k < n_groups; k++) {
    LklGroup &gg = ctx->h_groups[k];
    gg.n_same = 1;
    while (gg.n_same < gg.npts && gg.alpha[gg.n_same] == gg.alpha[0]) gg.n_same++;
    gg.pad_ = 0;
  }
  if (with_estep) {
    // the first point of every owned individual, in order, must be the parameters the context holds
    bool ok = n_groups == ctx->n_owned && ctx->h_indF.size() == ctx->n_owned;
    for (uint32_t k = 0; ok && k < n_groups; k++) {
      const LklGroup &gg = ctx->h_groups[k];
      ok = gg.ind == (int) k && gg.F[0] == ctx->h_indF[k] && gg.alpha[0] == ctx->h_alpha[k];
    }
    if (!ok)
      return fail(ctx, NFH_ERR_ARG,
                  "nfh_estep_with_batch: the first request of every owned individual must be its current (F, alpha)");
  }
  if (n_groups == 0) return NFH_OK;
  NFH_CUDA(cudaMemcpyAsync(ctx->groups, ctx->h_groups, n_groups * sizeof(LklGroup), cudaMemcpyHostToDevice,
                           ctx->stream));
  LklArgs a;
  a.emis = ctx->emis_recv; a.dist = ctx->dist; a.loge0_sum = ctx->loge0_sum + ctx->ind_begin;
  a.tile_dmax = ctx->tile_dmax; a.tile_dsum = ctx->tile_dsum;
  a.groups = ctx->groups; a.tile_prod = ctx->lkl_tile_prod; a.neg_lkl = ctx->neg_lkl;
  a.n_rows = ctx->n_loc; a.n_sites = ctx->n_sites; a.site_block = ctx->site_block;
  a.n_tiles = ctx->n_tiles; a.n_groups = n_groups;
  a.emit_chunk_prod = with_estep ? ctx->chunk_prod : nullptr;
  a.emit_tile_prod = with_estep ? ctx->tile_prod : nullptr;
  {
    FamilyScope fs(ctx, kFamLkl, 2);
    launch_lkl_batch(a, ctx->stream);
  }
  if (with_estep) {
    const EstepArgs ea = estep_args(ctx);
    FamilyScope fs(ctx, kFamEstep, 2);
    launch_estep_tail(ea, ctx->stream);
  }
  NFH_CUDA(cudaGetLastError());
  NFH_CUDA(cudaMemcpyAsync(ctx->h_small, ctx->neg_lkl, n_req * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  int rc = NFH_OK;
  if (with_estep) {
    NFH_CUDA(cudaMemcpyAsync(ctx->h_small + n_req, ctx->ind_lkl, ctx->n_owned * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    rc = sync_and_check(ctx);
    if (ind_lkl_out) memcpy(ind_lkl_out, ctx->h_small + n_req, ctx->n_owned * sizeof(double));
  } else {
    NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  for (uint32_t k = 0; k < n_groups; k++)
    for (int p = 0; p < ctx->h_groups[k].npts; p++) {
      int q = ctx->h_groups[k].out[p];
      neg_lkl_out[q] = ctx->h_small[q];
    }
  return rc;
}

int nfh_lkl_batch(nfh_ctx *ctx, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                  double *neg_lkl_out) {
  if (n_req == 0) return NFH_OK;
  return lkl_batch_impl(ctx, n_req, ind, F, alpha, neg_lkl_out, false, nullptr);
}

int nfh_estep_with_batch(nfh_ctx *ctx, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                         double *neg_lkl_out, double *ind_lkl_out) {
  if (ctx->n_owned == 0) return NFH_OK;
  return lkl_batch_impl(ctx, n_req, ind, F, alpha, neg_lkl_out, true, ind_lkl_out);
}

int nfh_viterbi(nfh_ctx *ctx, char *path_out) {
  if (!ctx->e0_recv) return fail(ctx, NFH_ERR_ARG, "nfh_viterbi: call nfh_emission_refresh(ctx, 1) first");
  NFH_CUDA(cudaSetDevice(ctx->device));
  const uint64_t stride = (uint64_t) ctx->n_tiles * kTile;            // whole tiles: bulk stores never run past a row
  const size_t n_chunks = (size_t) ctx->n_tiles * kScanThreads;
  if (!ctx->vit_work) {
    NFH_CUDA(cudaMalloc((void **) &ctx->vit_work, ctx->n_loc * stride));
    NFH_CUDA(cudaMemsetAsync(ctx->vit_work, 0, ctx->n_loc * stride, ctx->stream));   // padding bytes of the last tile
    NFH_CUDA(cudaMalloc((void **) &ctx->vit_maps, ctx->n_loc * (n_chunks + 2 * (size_t) ctx->n_tiles)));
    NFH_CUDA(cudaMalloc((void **) &ctx->vit_tile_prod, ctx->n_loc * ctx->n_tiles * sizeof(double4)));
    NFH_CUDA(cudaMalloc((void **) &ctx->vit_final, ctx->n_loc * sizeof(int)));
  }
  if (ctx->n_owned) {
    ViterbiArgs a;
    a.emis = ctx->emis_recv; a.e0 = ctx->e0_recv; a.dist = ctx->dist; a.indF = ctx->indF; a.alpha = ctx->alpha;
    a.work = ctx->vit_work; a.work_stride = stride;
    a.chunk_prod = ctx->chunk_prod;                                    // E-step scratch, free between iterations
    a.tile_prod = ctx->vit_tile_prod;
    a.tile_score = ctx->fwd_carry;
    a.chunk_map = ctx->vit_maps;
    a.tile_map = ctx->vit_maps + ctx->n_loc * n_chunks;
    a.tile_state = a.tile_map + ctx->n_loc * (size_t) ctx->n_tiles;
    a.final_state = ctx->vit_final;
    a.n_rows = ctx->n_loc; a.n_rows_valid = ctx->n_owned; a.n_sites = ctx->n_sites;
    a.site_block = ctx->site_block; a.n_tiles = ctx->n_tiles;
    {
      FamilyScope fs(ctx, kFamViterbi, 5);
      launch_viterbi(a, ctx->stream);
    }
    NFH_CUDA(cudaGetLastError());
    if (path_out)
      NFH_CUDA(cudaMemcpy2DAsync(path_out, ctx->n_sites, ctx->vit_work, stride, ctx->n_sites, ctx->n_owned,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  }
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_get_posterior(nfh_ctx *ctx, double *marg1_out) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  for (int b = 0; b < ctx->n_ranks && ctx->n_owned; b++) {
    const uint64_t s0 = (uint64_t) b * ctx->site_block;
    if (s0 >= ctx->n_sites) break;
    const uint64_t w = std::min(ctx->site_block, ctx->n_sites - s0);
    // fused exchange: the E-step stored block b straight into rank b's frequency-side window (source block = me)
    const double *src = ctx->post_in_peers ? ctx->peer_post[b] + (size_t) ctx->rank * ctx->n_loc * ctx->site_block
                                         : ctx->post_send + (size_t) b * ctx->n_loc * ctx->site_block;
    NFH_CUDA(cudaMemcpy2DAsync(marg1_out + s0, ctx->n_sites * sizeof(double), src, ctx->site_block * sizeof(double),
                               w * sizeof(double), ctx->n_owned, cudaMemcpyDeviceToHost, ctx->stream));
  }
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_geno_posterior(nfh_ctx *ctx, const char *path_all, double *geno_out) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  if (ctx->sites_owned == 0) return NFH_OK;
  const uint64_t N = ctx->n_ind_total;
  char *d_path = nullptr;
  NFH_CUDA(cudaMalloc((void **) &d_path, N * ctx->sites_owned));
  auto body = [&]() -> int {
    NFH_CUDA(cudaMemcpyAsync(d_path, path_all, N * ctx->sites_owned, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t per_site = N * 3 * sizeof(double);
    const uint64_t chunk = std::max<uint64_t>(1, kStageBytes / per_site);
    for (uint64_t s0 = 0; s0 < ctx->sites_owned; s0 += chunk) {
      const uint64_t m = std::min(chunk, ctx->sites_owned - s0);
      launch_geno_posterior(ctx->gl[0] + s0, ctx->gl[1] + s0, ctx->gl[2] + s0, ctx->freq + s0, d_path + s0, N,
                            ctx->site_block, ctx->sites_owned, m, (double *) ctx->d_stage, ctx->stream);
      ctx->launches++;
      NFH_CUDA(cudaGetLastError());
      NFH_CUDA(cudaMemcpyAsync(geno_out + s0 * N * 3, ctx->d_stage, m * per_site, cudaMemcpyDeviceToHost, ctx->stream));
      NFH_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return NFH_OK;
  };
  const int rc = body();
  cudaFree(d_path);      // on every path
  return rc;
}

static double *window_base(nfh_ctx *ctx, int window, uint64_t *bytes) {
  const uint64_t plane = ctx->n_ind_pad * ctx->site_block * sizeof(double);
  *bytes = plane;
  switch (window) {
    case NFH_WIN_POST_SEND: return ctx->post_send;
    case NFH_WIN_POST_RECV: return ctx->post_recv;
    case NFH_WIN_EMIS_SEND: return ctx->emis_send;
    case NFH_WIN_EMIS_RECV: return ctx->emis_recv;
    case NFH_WIN_E0_SEND: return ctx->e0_send;
    case NFH_WIN_E0_RECV: return ctx->e0_recv;
    case NFH_WIN_LOGE0_SUM: *bytes = ctx->n_ind_pad * sizeof(double); return ctx->loge0_sum;
    default: return nullptr;
  }
}

int nfh_exchange_window(nfh_ctx *ctx, int window, void **dev_ptr, uint64_t *bytes, uint64_t *bytes_per_peer) {
  if (window < NFH_WIN_POST_SEND || window > NFH_WIN_LOGE0_SUM)
    return fail(ctx, NFH_ERR_ARG, "nfh_exchange_window: unknown window");
  uint64_t b = 0;
  void *p = window_base(ctx, window, &b);
  if (dev_ptr) *dev_ptr = p;
  if (bytes) *bytes = b;
  if (bytes_per_peer) *bytes_per_peer = b / ctx->n_ranks;
  return NFH_OK;
}

int nfh_peer_export(nfh_ctx *ctx, int window, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  NFH_CUDA(cudaSetDevice(ctx->device));
  void *p = window == NFH_WIN_POST_RECV ? (void *) ctx->post_recv : window == NFH_WIN_EMIS_RECV ? (void *) ctx->emis_recv : nullptr;
  if (!p) return fail(ctx, NFH_ERR_ARG, "nfh_peer_export: only POST_RECV and EMIS_RECV can be exported");
  cudaIpcMemHandle_t h;
  NFH_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle, &h, 64);
  return NFH_OK;
}

int nfh_peer_import(nfh_ctx *ctx, int window, int peer_rank, const unsigned char handle[64]) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  if (peer_rank < 0 || peer_rank >= ctx->n_ranks || ctx->n_ranks > kMaxRanks)
    return fail(ctx, NFH_ERR_ARG, "nfh_peer_import: bad rank");
  double **slot = window == NFH_WIN_POST_RECV ? ctx->peer_post : window == NFH_WIN_EMIS_RECV ? ctx->peer_emis : nullptr;
  if (!slot) return fail(ctx, NFH_ERR_ARG, "nfh_peer_import: only POST_RECV and EMIS_RECV can be imported");
  if (peer_rank == ctx->rank) {
    slot[peer_rank] = window == NFH_WIN_POST_RECV ? ctx->post_recv : ctx->emis_recv;
    return NFH_OK;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void *p = nullptr;
  NFH_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  slot[peer_rank] = (double *) p;
  (window == NFH_WIN_POST_RECV ? ctx->peer_post_ipc : ctx->peer_emis_ipc)[peer_rank] = true;
  return NFH_OK;
}

int nfh_peer_set(nfh_ctx *ctx, int window, int peer_rank, nfh_ctx *peer) {
  if (!peer || peer_rank < 0 || peer_rank >= ctx->n_ranks || ctx->n_ranks > kMaxRanks || peer->rank != peer_rank ||
      peer->n_ranks != ctx->n_ranks || peer->n_ind_total != ctx->n_ind_total || peer->n_sites != ctx->n_sites)
    return fail(ctx, NFH_ERR_ARG, "nfh_peer_set: peer is not that rank of the same geometry");
  double **slot = window == NFH_WIN_POST_RECV ? ctx->peer_post : window == NFH_WIN_EMIS_RECV ? ctx->peer_emis : nullptr;
  if (!slot) return fail(ctx, NFH_ERR_ARG, "nfh_peer_set: only POST_RECV and EMIS_RECV can be mapped");
  NFH_CUDA(cudaSetDevice(ctx->device));
  if (peer->device != ctx->device) {
    int can = 0;
    NFH_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
    if (!can) return fail(ctx, NFH_ERR_ARG, "nfh_peer_set: the two devices have no peer access");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NFH_CUDA(e);
    cudaGetLastError();
  }
  slot[peer_rank] = window == NFH_WIN_POST_RECV ? peer->post_recv : peer->emis_recv;
  return NFH_OK;
}

int nfh_window_copy_block(nfh_ctx *dst, int dst_window, int dst_block, nfh_ctx *src, int src_window, int src_block) {
  nfh_ctx *ctx = src;
  uint64_t db = 0, sb = 0;
  double *d = window_base(dst, dst_window, &db), *s = window_base(src, src_window, &sb);
  if (!d || !s || dst_window == NFH_WIN_LOGE0_SUM || src_window == NFH_WIN_LOGE0_SUM || db != sb ||
      dst->n_ranks != src->n_ranks || dst_block < 0 || dst_block >= dst->n_ranks || src_block < 0 ||
      src_block >= src->n_ranks)
    return fail(ctx, NFH_ERR_ARG, "nfh_window_copy_block: windows do not match");
  const uint64_t block = sb / src->n_ranks;
  NFH_CUDA(cudaSetDevice(src->device));
  NFH_CUDA(cudaMemcpyPeerAsync((char *) d + (uint64_t) dst_block * block, dst->device,
                               (const char *) s + (uint64_t) src_block * block, src->device, block, src->stream));
  return NFH_OK;
}

int nfh_window_read(nfh_ctx *ctx, int window, uint64_t offset, uint64_t bytes, void *host_dst) {
  uint64_t wb = 0;
  const char *p = (const char *) window_base(ctx, window, &wb);
  if (!p || !host_dst || offset + bytes > wb) return fail(ctx, NFH_ERR_ARG, "nfh_window_read: outside the window");
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaMemcpyAsync(host_dst, p + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_window_write(nfh_ctx *ctx, int window, uint64_t offset, uint64_t bytes, const void *host_src) {
  uint64_t wb = 0;
  char *p = (char *) window_base(ctx, window, &wb);
  if (!p || !host_src || offset + bytes > wb) return fail(ctx, NFH_ERR_ARG, "nfh_window_write: outside the window");
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaMemcpyAsync(p + offset, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  return NFH_OK;
}

int nfh_peer_direct(nfh_ctx *ctx, int enable) {
  if (enable) {
    for (int r = 0; r < ctx->n_ranks; r++)
      if (!ctx->peer_post[r] || !ctx->peer_emis[r])
        return fail(ctx, NFH_ERR_ARG, "nfh_peer_direct: import the POST_RECV and EMIS_RECV windows of every rank first");
  }
  ctx->peer_direct = enable != 0;
  ctx->peer_post_direct = enable == 1;      // enable == 2: emission ratios only
  return NFH_OK;
}

int nfh_sync(nfh_ctx *ctx) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  return sync_and_check(ctx);
}

int nfh_probe_fp64(nfh_ctx *ctx, double *flops_per_s) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  *flops_per_s = launch_fp64_probe(ctx->stream, ctx->sm_count);
  NFH_CUDA(cudaGetLastError());
  return NFH_OK;
}

int nfh_timing(nfh_ctx *ctx, int enable) {
  ctx->timing = enable != 0;
  return NFH_OK;
}

int nfh_timing_read(nfh_ctx *ctx, double ms_out[8], uint64_t launches_out[8], int reset) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto &s : ctx->spans) {
    float ms = 0;
    cudaEventElapsedTime(&ms, s.t0, s.t1);
    ctx->fam_ms[s.family] += ms;
    ctx->free_events.push_back(s.t0);
    ctx->free_events.push_back(s.t1);
  }
  ctx->spans.clear();
  for (int f = 0; f < kFamilies; f++) {
    if (ms_out) ms_out[f] = ctx->fam_ms[f];
    if (launches_out) launches_out[f] = ctx->fam_launches[f];
    if (reset) { ctx->fam_ms[f] = 0; ctx->fam_launches[f] = 0; }
  }
  return NFH_OK;
}

int nfh_host_register(nfh_ctx *ctx, void *ptr, uint64_t bytes) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return NFH_OK;
}

int nfh_host_unregister(nfh_ctx *ctx, void *ptr) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  NFH_CUDA(cudaHostUnregister(ptr));
  return NFH_OK;
}

int nfh_freq_passes(nfh_ctx *ctx, uint64_t *total, int reset) {
  NFH_CUDA(cudaSetDevice(ctx->device));
  unsigned long long v = 0;
  NFH_CUDA(cudaMemcpyAsync(&v, ctx->freq_passes, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
  if (reset) NFH_CUDA(cudaMemsetAsync(ctx->freq_passes, 0, sizeof v, ctx->stream));
  NFH_CUDA(cudaStreamSynchronize(ctx->stream));
  if (total) *total = v;
  return NFH_OK;
}

}  // extern "C"

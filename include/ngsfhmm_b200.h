/*
 * ngsfhmm_b200.h - C ABI of the B200-native ngsF-HMM EM hot path.
 *
 * The reference (fgvieira/ngsF-HMM v1.1.0) has no plugin/FFI seam; its hot
 * path is entered through C++ calls on one `params` struct
 * (ngsF-HMM.hpp:13-52): iter_EM(params*) once per EM iteration (EM.cpp:72,
 * body EM.cpp:139-289) and one Viterbi task per individual after the loop
 * (EM.cpp:110-116).  This header is the seam a maintainer would bind instead:
 * each entry point names the reference code it replaces.  extern "C", plain
 * pointers and sizes, int status (0 = OK).  Host arrays are caller-owned;
 * device buffers are owned by the opaque context.  All calls are made from
 * one host thread per context (the reference calls from its main thread and
 * joins its pool before reading results, EM.cpp:161,201).
 *
 * Sites are 0-based here (reference site s+1).  FP64 throughout.
 *
 * Multi-GPU geometry (one process per GPU).  A context belongs to rank r of
 * n_ranks.  Individuals are sharded for the recursions: rank r owns the
 * contiguous block [r*n_ind_local, ...) of n_ind_total.  Sites are sharded
 * for the allele-frequency update: rank r owns site block
 * [r*site_block, (r+1)*site_block).  Between the two stages the posteriors and
 * the refreshed emissions cross ranks through the exchange windows below
 * (equal-split all-to-all; the host plumbing - torch.distributed/NCCL - moves
 * the bytes, this library never opens a communicator).  With n_ranks == 1 the
 * send and receive windows alias and no exchange is needed.
 */
#ifndef NGSFHMM_B200_H
#define NGSFHMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nfh_ctx nfh_ctx;

enum nfh_status {
  NFH_OK = 0,
  NFH_ERR_CUDA = 1,        /* CUDA runtime failure; see nfh_last_error() */
  NFH_ERR_ARG = 2,         /* bad argument / call order */
  NFH_ERR_NAN = 3,         /* reference: error("invalid Lkl found!") HMM.cpp:18-21,45-48; "value is NaN!" gen_func.cpp:56-57 */
  NFH_ERR_FWBW = 4,        /* reference: error("Fw and Bw lkl do not match!") EM.cpp:166-170 */
  NFH_ERR_NOMEM = 5,
  NFH_ERR_NO_DEVICE = 6    /* no CUDA device: there is NO CPU fallback */
};

/* exchange windows, see nfh_exchange_window() */
enum nfh_window {
  NFH_WIN_POST_SEND = 0,   /* posteriors, recursion side  [n_ranks][n_ind_local][site_block] f64 */
  NFH_WIN_POST_RECV = 1,   /* posteriors, frequency side  [n_ranks][n_ind_local][site_block] f64 */
  NFH_WIN_EMIS_SEND = 2,   /* emission ratio e1/e0, frequency side, same shape */
  NFH_WIN_EMIS_RECV = 3,   /* emission ratio, recursion side, same shape */
  NFH_WIN_E0_SEND = 4,     /* state-0 emission e0 (only materialised for Viterbi), frequency side */
  NFH_WIN_E0_RECV = 5,     /* e0, recursion side */
  NFH_WIN_LOGE0_SUM = 6    /* per-individual sum over this rank's sites of log e0: [n_ranks*n_ind_local] f64; all-reduce(sum) */
};

const char *nfh_strerror(int status);
const char *nfh_last_error(const nfh_ctx *ctx);
/* Library build tag, e.g. "sm_100a"; also proves the .so loaded. */
const char *nfh_build_info(void);
/* Number of CUDA devices visible to this process (0 without a driver or a GPU). */
int nfh_device_count(void);
/* Number of CUDA kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t nfh_kernel_launches(const nfh_ctx *ctx);

/* Replaces: params allocation in main()/init_output (ngsF-HMM.cpp:75-135,
 * parse_args.cpp:229-419).  device = CUDA ordinal.  n_ranks >= 1, 0 <= rank < n_ranks. */
int nfh_ctx_create(nfh_ctx **out, int device, uint64_t n_ind_total, uint64_t n_sites, int n_ranks, int rank);
void nfh_ctx_destroy(nfh_ctx *ctx);

/* Geometry chosen by the context. */
uint64_t nfh_n_ind_local(const nfh_ctx *ctx);     /* padded individuals per rank */
uint64_t nfh_n_ind_owned(const nfh_ctx *ctx);     /* real individuals this rank owns (<= n_ind_local) */
uint64_t nfh_ind_begin(const nfh_ctx *ctx);       /* first global individual index owned */
uint64_t nfh_site_block(const nfh_ctx *ctx);      /* padded sites per rank */
uint64_t nfh_site_begin(const nfh_ctx *ctx);      /* first site of this rank's block */
uint64_t nfh_sites_owned(const nfh_ctx *ctx);     /* real sites in this rank's block */

/* Replaces: geno_lkl / geno_lkl_s (ngsF-HMM.hpp:38-39) filled by read_geno
 * (read_data.cpp:13-116).  log_gl: natural-log, normalised GL for ALL
 * n_ind_total individuals, site-major [n][n_ind_total][3] - the layout of the
 * binary input file (read_data.cpp:28-31) - covering sites
 * [first_site, first_site+n) which must lie inside this rank's site block.
 * May be called repeatedly with consecutive chunks. */
int nfh_upload_gl(nfh_ctx *ctx, const double *log_gl, uint64_t first_site, uint64_t n);

/* Replaces: pos_dist (ngsF-HMM.hpp:40; read_dist read_data.cpp:165-218,
 * bp->Mb ngsF-HMM.cpp:85-86).  dist_mb[n_sites], +inf at chromosome starts. */
int nfh_upload_pos_dist(nfh_ctx *ctx, const double *dist_mb);

/* Replaces: pars->freq[1..S] (this rank's sites_owned values). */
int nfh_set_freq(nfh_ctx *ctx, const double *freq);
int nfh_get_freq(nfh_ctx *ctx, double *freq);

/* Replaces: pars->indF / pars->alpha for the owned individuals. */
int nfh_set_ind_params(nfh_ctx *ctx, const double *indF, const double *alpha);

/* Replaces: calc_emission over all (i, s) (HMM.cpp:144-154; init
 * parse_args.cpp:381-386).  Frequency side: fills NFH_WIN_EMIS_SEND and
 * NFH_WIN_LOGE0_SUM from GL and the current freq.  with_e0 != 0 also fills
 * NFH_WIN_E0_SEND (needed before nfh_viterbi). */
int nfh_emission_refresh(nfh_ctx *ctx, int with_e0);

/* Replaces: forward + backward tasks, the Fw/Bw check, ind_lkl and the clamped
 * posterior (EM.cpp:151-185; HMM.cpp:6-60; check_interv gen_func.cpp:55-70).
 * Reads NFH_WIN_EMIS_RECV (+ reduced NFH_WIN_LOGE0_SUM), writes
 * NFH_WIN_POST_SEND; ind_lkl_out[n_ind_owned]. */
int nfh_estep(nfh_ctx *ctx, double *ind_lkl_out);

/* Introspection of the work order nfh_estep uses inside its single launch (csrc/nfh_schedule.h): item
 * `ticket` of the order for n_rows individuals x n_tiles tiles in waves of wave_rows individuals with
 * `lookahead` early product items.  item_out = {phase (0 products, 1 posteriors), individual, tile}; returns the
 * number of items.  Pure host function (no device needed): lets the tests prove that every item appears once
 * and that no posterior item precedes a product item of its wave. */
uint64_t nfh_estep_schedule_item(uint32_t n_rows, uint32_t n_tiles, uint32_t wave_rows, uint32_t lookahead,
                                 uint64_t ticket, uint32_t item_out[3]);

/* Replaces: lkl() (EM.cpp:449-464) for many (individual, F, alpha) points per
 * launch - the objective evaluations findmax_bfgs asks for (bfgs.cpp:108-121).
 * ind[] are local individual indices; requests for the same individual
 * should be adjacent (they then share one read of its emissions).
 * neg_lkl_out[q] = -logLkl; NaN/Inf parameters give -1e15 as the reference does. */
int nfh_lkl_batch(nfh_ctx *ctx, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                  double *neg_lkl_out);

/* nfh_estep and the first nfh_lkl_batch of an EM iteration in one call.  findmax_bfgs starts at the
 * parameters the E-step has just used (EM.cpp:151-205), so its first objective evaluation is the E-step's
 * forward recursion: when the first request of every owned individual (individuals in order 0..n_ind_owned-1,
 * all present) is its current (F, alpha), the forward products of that point also feed the posterior and
 * the E-step's own product kernel is skipped.  Same outputs as the two calls. */
int nfh_estep_with_batch(nfh_ctx *ctx, uint64_t n_req, const int32_t *ind, const double *F, const double *alpha,
                         double *neg_lkl_out, double *ind_lkl_out);

/* Replaces: the freq/emission site loop (EM.cpp:224-271) = est_maf
 * (gen_func.cpp:974-1009) + calc_emission.  method 1 = per-site EM (the only
 * method the reference can run, SURVEY.md finding 4); method 0 = keep freq,
 * refresh emissions only.  Frequency side: reads NFH_WIN_POST_RECV, writes
 * freq, NFH_WIN_EMIS_SEND, NFH_WIN_LOGE0_SUM.  freq_out[sites_owned] or NULL.
 * posterior_is_zero != 0 uses F_i = 0 for everyone (est_maf with scalar F = 0,
 * the "--freq e" initialisation, parse_args.cpp:316-318). */
int nfh_freq_update(nfh_ctx *ctx, int method, int posterior_is_zero, double *freq_out);

/* Replaces: the Viterbi tasks (EM.cpp:110-116; HMM.cpp:98-125, incl. the
 * in-place score update).  Needs NFH_WIN_E0_RECV / NFH_WIN_EMIS_RECV current.
 * path_out[n_ind_owned][n_sites], values 0/1. */
int nfh_viterbi(nfh_ctx *ctx, char *path_out);

/* Replaces: reading pars->marg_prob[i][s][1] in print_iter (EM.cpp:349-351).
 * marg1_out[n_ind_owned][n_sites]. */
int nfh_get_posterior(nfh_ctx *ctx, double *marg1_out);

/* Replaces: the .geno posterior in print_iter (EM.cpp:369-376):
 * exp(post_prob(GL, HWE(freq[s], F = path[i][s]))).  Frequency side;
 * path_all[n_ind_total][sites_owned]; geno_out site-major
 * [sites_owned][n_ind_total][3]. */
int nfh_geno_posterior(nfh_ctx *ctx, const char *path_all, double *geno_out);

/* Device pointer + byte size of an exchange window, for the host all-to-all /
 * all-reduce.  bytes_per_peer = bytes / n_ranks for the all-to-all windows. */
int nfh_exchange_window(nfh_ctx *ctx, int window, void **dev_ptr, uint64_t *bytes, uint64_t *bytes_per_peer);

/* Fused exchange over NVLink peer memory (one process per GPU, same node).  Each rank exports CUDA IPC
 * handles (64 bytes) of its two receive windows, the host plumbing all-gathers them, every rank imports
 * the handles of all ranks (its own included), then nfh_peer_direct(ctx, 1) makes
 *   - nfh_estep write every posterior tile straight into the frequency-side window of the rank that
 *     owns the tile's site block, and
 *   - nfh_freq_update / nfh_emission_refresh write every emission ratio straight into the
 *     recursion-side window of the rank that owns the individual,
 * so no all-to-all is needed; the caller only has to order the stages across ranks (a 1-element
 * all-reduce on nfh_stream() before nfh_freq_update, and the all-reduce of NFH_WIN_LOGE0_SUM after it).
 * enable = 2 keeps the posteriors of nfh_estep in the local NFH_WIN_POST_SEND and only sends the emission ratios
 * direct: for runs whose frequencies are fixed (--freq_est 0, EM.cpp:224 skips the site loop), where nothing on the
 * frequency side reads the posteriors and 8 bytes per individual-site would cross NVLink for nothing. */
int nfh_peer_export(nfh_ctx *ctx, int window, unsigned char handle[64]);
int nfh_peer_import(nfh_ctx *ctx, int window, int peer_rank, const unsigned char handle[64]);
int nfh_peer_direct(nfh_ctx *ctx, int enable);

/* The same fused exchange when ONE process drives several contexts (one per GPU, or several on one GPU):
 * no IPC handle is needed, the peer's receive window is mapped directly.  Enables peer access between the
 * two devices when they differ.  window = NFH_WIN_POST_RECV or NFH_WIN_EMIS_RECV; peer must be rank
 * peer_rank of the same geometry. */
int nfh_peer_set(nfh_ctx *ctx, int window, int peer_rank, nfh_ctx *peer);

/* Block copy between exchange windows of two contexts of one process (the all-to-all of the host plumbing
 * without a communicator): block src_block of src's window -> block dst_block of dst's window, one block =
 * [n_ind_local][site_block] doubles.  Queued on src's stream; nfh_sync(src) completes it. */
int nfh_window_copy_block(nfh_ctx *dst, int dst_window, int dst_block, nfh_ctx *src, int src_window, int src_block);

/* Host access to a window (bytes at byte offset), synchronous: the all-reduce of NFH_WIN_LOGE0_SUM across the
 * contexts of one process is a read, a host sum and a write. */
int nfh_window_read(nfh_ctx *ctx, int window, uint64_t offset, uint64_t bytes, void *host_dst);
int nfh_window_write(nfh_ctx *ctx, int window, uint64_t offset, uint64_t bytes, const void *host_src);

/* Page-lock a caller-owned host array (cudaHostRegister) so that the copies of the calls above move at
 * full PCIe speed instead of being staged: worth it for the arrays that cross every EM iteration
 * (freq_out of nfh_freq_update is 8 bytes per site).  Unregister before freeing the array. */
int nfh_host_register(nfh_ctx *ctx, void *ptr, uint64_t bytes);
int nfh_host_unregister(nfh_ctx *ctx, void *ptr);

/* Block until everything queued by this context has finished; returns the
 * sticky device status (NaN / FwBw flags raised by kernels). */
int nfh_sync(nfh_ctx *ctx);

/* The CUDA stream (cudaStream_t) the context launches on, so host plumbing
 * can order its collectives and time with events on the same stream. */
void *nfh_stream(const nfh_ctx *ctx);

/* Diagnostics for bench.py: run the FP64 FMA probe kernel and return achieved
 * FLOP/s (2 flops per DFMA) - the measured denominator for FP64-bound kernels. */
int nfh_probe_fp64(nfh_ctx *ctx, double *flops_per_s);
/* Device-time (ms) spent in each kernel family since the last reset, measured
 * with CUDA events on the context stream: [0]=estep [1]=lkl_batch [2]=freq
 * [3]=viterbi [4]=emission.  Enabled by nfh_timing(ctx, 1). */
int nfh_timing(nfh_ctx *ctx, int enable);
int nfh_timing_read(nfh_ctx *ctx, double ms_out[8], uint64_t launches_out[8], int reset);
/* Sum over this rank's sites of the est_maf passes each site ran (<= 101, gen_func.cpp:1006) in all
 * nfh_freq_update calls since the last reset: the pass count behind bench.py's FP64 work figure. */
int nfh_freq_passes(nfh_ctx *ctx, uint64_t *total, int reset);

#ifdef __cplusplus
}
#endif
#endif

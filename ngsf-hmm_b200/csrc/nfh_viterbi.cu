// nfh_viterbi.cu - most probable IBD path for all individuals, as chunked scans.
//
// Replaces viterbi() (shared/HMM.cpp:98-125).  Two quirks of the reference are
// kept because they change the decoded tracts (SURVEY.md finding 3):
//   * the score of state 0 is overwritten before state 1 of the same site is
//     evaluated, so state 1 competes against the UPDATED state-0 score;
//   * comparisons are strict (vmax < pval), so ties keep the k = 0 predecessor,
//     and the final state is the first maximum (array_max_pos).
//
// The reference works in log space.  Here scores live in LINEAR space, the
// (max, x) semiring instead of (max, +), rescaled by exact powers of two, which
// makes the same decisions except where two candidates agree to ~1e-16
// relative - far inside the < 1e-9 tie band the parity contract excludes.
//
// With the in-place quirk one site is still a (max, x)-linear map of the score
// pair (v0, v1).  Writing the transition as T = c A, A = I + kappa 1 q' (the
// factored form of nfh_device.cuh), the update is
//   new0 = c max(v0 A00, v1 A10) e0
//   new1 = c max(new0 A01, v1 A11) e1          <- new0 already carries this site's c
// so, unlike in the E-step, the scalar c does NOT cancel: after dropping the one
// common factor c the state-1 candidate through state 0 keeps c A01 = (1-c) q1,
// the true 0->1 transition probability t01:
//   Q[0][0] = A00 e0              Q[0][1] = A00 e0 t01 e1
//   Q[1][0] = A10 e0              Q[1][1] = max(A10 e0 t01, A11) e1
// (e1 = e0 r).  The score pair
// at any site is a chunked scan exactly like the E-step, and the traceback
// path[s-1] = back[s][path[s]] is a scan of compositions of maps {0,1}->{0,1}.
//
//   viterbi_chunk_products : (max,x) product of each thread's 33 sites, and per tile
//   viterbi_tile_scores    : per individual, scores at the start of every tile, final state
//   viterbi_chunk_pointers : scores at chunk starts (in-CTA scan), 33 sites of back-pointer
//                            pairs, composed map per chunk and per tile
//   viterbi_tile_states    : per individual, state at the end of every tile (right to left)
//   viterbi_chunk_trace    : state at the end of every chunk, then the 33-site traceback
#include "nfh_device.cuh"
#include "nfh_viterbi_math.cuh"
#include "nfh_kernels.h"

namespace nfh {

struct VitSmem {
  alignas(128) double r[kTile];     // emission ratio e1/e0
  alignas(128) double e0[kTile];
  alignas(128) double d[kTile];
  alignas(16) unsigned char bp[kTile];   // back-pointer pairs / path bytes of the tile
  alignas(8) uint64_t bar;
  double tab[64];
};

__device__ __forceinline__ int vit_valid_sites(uint64_t first_site, uint64_t n_sites) {
  return first_site >= n_sites ? 0 : (int) min((uint64_t) kChunk, n_sites - first_site);
}

__device__ __forceinline__ void vit_stage(VitSmem &sm, const double *__restrict__ emis_tile,
                                          const double *__restrict__ e0_tile, const double *__restrict__ dist_tile) {
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
  }
  load_exp_table(sm.tab);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&sm.bar, 3 * kTileBytes);
    tma_load_1d(sm.r, emis_tile, kTileBytes, &sm.bar);
    tma_load_1d(sm.e0, e0_tile, kTileBytes, &sm.bar);
    tma_load_1d(sm.d, dist_tile, kTileBytes, &sm.bar);
  }
  mbar_wait(&sm.bar, 0);
}

// ordered (max,x) product over the warp (lane 0 gets the total)
__device__ __forceinline__ void warp_ordered_tropical(M2 &m) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const M2 o = shfl_down_m(m, off);
    if ((lane & (2 * off - 1)) == 0) { m = tropmul(m, o); renorm(m); }
  }
}

__global__ void __launch_bounds__(kScanThreads)
viterbi_chunk_products(ViterbiArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  VitSmem &sm = *reinterpret_cast<VitSmem *>(smem_raw);
  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  const size_t tile_at = blocked_index(row, tile_first, A.n_rows, A.site_block);
  vit_stage(sm, A.emis + tile_at, A.e0 + tile_at, A.dist + tile_first);

  const int n_valid = vit_valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, A.n_sites);
  const double F = A.indF[row], al = A.alpha[row];
  const double q0 = 1.0 - F, q1 = F;
  const double *r = sm.r + threadIdx.x * kChunk, *e0 = sm.e0 + threadIdx.x * kChunk, *d = sm.d + threadIdx.x * kChunk;

  M2 m = identity2();
#pragma unroll 3
  for (int j = 0; j < kChunk; j++) {
    if (j < n_valid) {
      const double kap = site_kappa(al * d[j], sm.tab);
      trop_apply(m, site_q(kap, q0, q1, e0[j], r[j]));
    }
    if (j % 6 == 5) renorm(m);
  }
  renorm(m);
  A.chunk_prod[((size_t) row * A.n_tiles + tile) * kScanThreads + threadIdx.x] = make_double4(m.a, m.b, m.c, m.d);

  __shared__ M2 sm_m[kScanThreads / 32];
  warp_ordered_tropical(m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm_m[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    M2 acc = sm_m[0];
#pragma unroll
    for (int w = 1; w < kScanThreads / 32; w++) { acc = tropmul(acc, sm_m[w]); renorm(acc); }
    A.tile_prod[(size_t) row * A.n_tiles + tile] = make_double4(acc.a, acc.b, acc.c, acc.d);
  }
}

// One warp per individual: scores entering every tile (32 tiles per step, warp scan), final state.
__global__ void __launch_bounds__(128)
viterbi_tile_scores(ViterbiArgs A) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= A.n_rows_valid) return;
  const double F = A.indF[row];
  double v0 = 1.0 - F, v1 = F;                                  // Vi_prob = q, linear
  const double4 *tp = A.tile_prod + (size_t) row * A.n_tiles;
  double2 *score = A.tile_score + (size_t) row * A.n_tiles;
  for (uint32_t g = 0; g < (A.n_tiles + 31) / 32; g++) {
    const uint32_t t = g * 32 + lane;
    M2 m = identity2();
    if (t < A.n_tiles) { const double4 p = tp[t]; m.a = p.x; m.b = p.y; m.c = p.z; m.d = p.w; }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const M2 o = shfl_up_m(m, off);
      if (lane >= off) { m = tropmul(o, m); renorm(m); }
    }
    const M2 before = shfl_up_m(m, 1);
    double c0 = v0, c1 = v1;
    if (lane > 0) { c0 = fmax(v0 * before.a, v1 * before.c); c1 = fmax(v0 * before.b, v1 * before.d); renorm2(c0, c1); }
    if (t < A.n_tiles) score[t] = make_double2(c0, c1);
    M2 tot;
    tot.a = __shfl_sync(kFull, m.a, 31); tot.b = __shfl_sync(kFull, m.b, 31);
    tot.c = __shfl_sync(kFull, m.c, 31); tot.d = __shfl_sync(kFull, m.d, 31);
    const double y0 = fmax(v0 * tot.a, v1 * tot.c), y1 = fmax(v0 * tot.b, v1 * tot.d);
    v0 = y0; v1 = y1;
    renorm2(v0, v1);
  }
  if (lane == 0) A.final_state[row] = v1 > v0 ? 1 : 0;           // array_max_pos: first maximum
}

__global__ void __launch_bounds__(kScanThreads)
viterbi_chunk_pointers(ViterbiArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  VitSmem &sm = *reinterpret_cast<VitSmem *>(smem_raw);
  constexpr int kWarps = kScanThreads / 32;
  __shared__ M2 warp_tot[kWarps];
  __shared__ unsigned warp_map[kWarps];
  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  const size_t tile_at = blocked_index(row, tile_first, A.n_rows, A.site_block);
  vit_stage(sm, A.emis + tile_at, A.e0 + tile_at, A.dist + tile_first);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_valid = vit_valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, A.n_sites);
  const double F = A.indF[row], al = A.alpha[row];
  const double q0 = 1.0 - F, q1 = F;

  // scores entering my chunk: tile score x products of the chunks before me
  const double4 mine4 = A.chunk_prod[((size_t) row * A.n_tiles + tile) * kScanThreads + threadIdx.x];
  M2 pre; pre.a = mine4.x; pre.b = mine4.y; pre.c = mine4.z; pre.d = mine4.w;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const M2 o = shfl_up_m(pre, off);
    if (lane >= off) { pre = tropmul(o, pre); renorm(pre); }
  }
  if (lane == 31) warp_tot[warp] = pre;
  __syncthreads();
  const double2 ts = A.tile_score[(size_t) row * A.n_tiles + tile];
  double v0 = ts.x, v1 = ts.y;
  for (int w = 0; w < warp; w++) {
    const M2 p = warp_tot[w];
    const double y0 = fmax(v0 * p.a, v1 * p.c), y1 = fmax(v0 * p.b, v1 * p.d);
    v0 = y0; v1 = y1;
    renorm2(v0, v1);
  }
  const M2 before = shfl_up_m(pre, 1);
  if (lane > 0) {
    const double y0 = fmax(v0 * before.a, v1 * before.c), y1 = fmax(v0 * before.b, v1 * before.d);
    v0 = y0; v1 = y1;
    renorm2(v0, v1);
  }

  // my 33 sites: back-pointer pair per site, composed map of the chunk
  const double *r = sm.r + threadIdx.x * kChunk, *e0 = sm.e0 + threadIdx.x * kChunk, *d = sm.d + threadIdx.x * kChunk;
  unsigned char *bp = sm.bp + threadIdx.x * kChunk;
  unsigned chunk_map = 2u;                                       // identity: 0->0, 1->1
#pragma unroll 3
  for (int j = 0; j < kChunk; j++) {
    unsigned bits = 2u;                                          // padding sites: identity
    if (j < n_valid) {
      const double kap = site_kappa(al * d[j], sm.tab);
      const double k0 = kap * q0, k1 = kap * q1, e1 = e0[j] * r[j];
      // l = 0: candidates from k = 0 (stay) and k = 1
      double from0 = v0 * (1.0 + k0), from1 = v1 * k0;
      const unsigned bp0 = from1 > from0;
      const double n0 = (bp0 ? from1 : from0) * e0[j];
      // l = 1: k = 0 uses the score just written for state 0 (in-place quirk), which already
      // carries this site's factor c: c kappa q1 = (1-c) q1
      from0 = n0 * trans01(kap, q1); from1 = v1 * (1.0 + k1);
      const unsigned bp1 = from1 > from0;
      const double n1 = (bp1 ? from1 : from0) * e1;
      v0 = n0; v1 = n1;
      if (j % 4 == 3) renorm2(v0, v1);
      bits = bp0 | (bp1 << 1);
    }
    bp[j] = (unsigned char) bits;
    // path[s-1] = bits_s[path[s]]: the chunk map sends the state at the chunk's last site back
    // to the state before its first site, i.e. f_first o ... o f_last
    chunk_map = map_compose(chunk_map, bits);
  }
  A.chunk_map[((size_t) row * A.n_tiles + tile) * kScanThreads + threadIdx.x] = (unsigned char) chunk_map;

  // tile map = map_0 o map_1 o ... o map_127 (ordered reduction)
  unsigned m = chunk_map;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned o = __shfl_down_sync(kFull, m, off);
    if ((lane & (2 * off - 1)) == 0) m = map_compose(m, o);
  }
  if (lane == 0) warp_map[warp] = m;
  fence_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = warp_map[0];
#pragma unroll
    for (int w = 1; w < kWarps; w++) acc = map_compose(acc, warp_map[w]);
    A.tile_map[(size_t) row * A.n_tiles + tile] = (unsigned char) acc;
    tma_store_1d(A.work + (size_t) row * A.work_stride + tile_first, sm.bp, kTile);
    tma_store_wait_read();
  }
}

// One thread per individual: state at the end of every tile, right to left.
__global__ void viterbi_tile_states(ViterbiArgs A) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.n_rows_valid) return;
  unsigned state = A.final_state[row];
  const unsigned char *tm = A.tile_map + (size_t) row * A.n_tiles;
  unsigned char *ts = A.tile_state + (size_t) row * A.n_tiles;
  for (uint32_t t = A.n_tiles; t-- > 0;) {
    ts[t] = (unsigned char) state;                               // state at the last site of tile t
    state = (tm[t] >> state) & 1u;                               // state at the last site of tile t-1
  }
}

__global__ void __launch_bounds__(kScanThreads)
viterbi_chunk_trace(ViterbiArgs A) {
  __shared__ alignas(16) unsigned char tile_bytes[kTile];
  __shared__ unsigned char maps[kScanThreads];
  __shared__ unsigned char end_state[kScanThreads];
  __shared__ alignas(8) uint64_t bar;
  const uint32_t tile = blockIdx.x, row = blockIdx.y;
  const uint64_t tile_first = (uint64_t) tile * kTile;
  unsigned char *gl_tile = A.work + (size_t) row * A.work_stride + tile_first;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  maps[threadIdx.x] = A.chunk_map[((size_t) row * A.n_tiles + tile) * kScanThreads + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, kTile);
    tma_load_1d(tile_bytes, gl_tile, kTile, &bar);
    // state at the last site of every chunk, right to left (128 two-bit maps: one thread is plenty)
    unsigned state = A.tile_state[(size_t) row * A.n_tiles + tile];
    for (int c = kScanThreads - 1; c >= 0; c--) {
      end_state[c] = (unsigned char) state;
      state = (maps[c] >> state) & 1u;
    }
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  const int n_valid = vit_valid_sites(tile_first + (uint64_t) threadIdx.x * kChunk, A.n_sites);
  unsigned char *b = tile_bytes + threadIdx.x * kChunk;
  vit_chunk_trace(b, n_valid, end_state[threadIdx.x]);
  fence_async_shared();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_1d(gl_tile, tile_bytes, kTile);
    tma_store_wait_read();
  }
}

void launch_viterbi(const ViterbiArgs &a, cudaStream_t st) {
  // per launch: the attribute belongs to the current device, and a process may drive several
  cudaFuncSetAttribute(viterbi_chunk_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(VitSmem));
  cudaFuncSetAttribute(viterbi_chunk_pointers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(VitSmem));
  dim3 grid(a.n_tiles, (unsigned) a.n_rows_valid);
  viterbi_chunk_products<<<grid, kScanThreads, sizeof(VitSmem), st>>>(a);
  viterbi_tile_scores<<<(unsigned) ((a.n_rows_valid + 3) / 4), 128, 0, st>>>(a);
  viterbi_chunk_pointers<<<grid, kScanThreads, sizeof(VitSmem), st>>>(a);
  viterbi_tile_states<<<(unsigned) ((a.n_rows_valid + 63) / 64), 64, 0, st>>>(a);
  viterbi_chunk_trace<<<grid, kScanThreads, 0, st>>>(a);
}

}  // namespace nfh

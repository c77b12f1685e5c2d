// simt.h - TEST INFRASTRUCTURE ONLY: a small SIMT emulator, enough CUDA to run the product's kernels on the host.
//
// tests/test_simt_kernels_cpu.py copies ngsf-hmm_b200/csrc/*.cu / *.cuh into a scratch directory, applies a handful
// of mechanical rewrites (kernel<<<g, b, s, st>>>(args) -> simt::launch(...), `extern __shared__` arrays -> the
// emulator's dynamic shared memory, the inline-PTX sections cut out) and compiles them with g++ against this header,
// which stands in for <cuda_runtime.h>.  Every CUDA thread of a CTA is a fiber (ucontext); CTAs run one after the
// other; __syncthreads, warp shuffles and votes are rendezvous points between fibers; the TMA bulk copies and
// their mbarriers (nfh_tma.cuh) are synchronous memcpy plus a phase bit.  The CPU test suite thereby runs the
// kernels' real control flow - tiles, chunk scans, carries, ordered warp products, launch geometry - against the
// oracle without a GPU.  It is a checker: nothing under ngsf-hmm_b200/ includes, builds or loads it.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- CUDA vocabulary ------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local     // one emulated CTA at a time PER HOST THREAD (host/group.cpp drives ranks from threads)
#define __align__(n) alignas(n)
#define NFH_DEV static inline
#define NFH_DEV_TABLE static const

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- the slice of the driver API behind freq_tensor_maps() (nfh_freq.cu): cuTensorMapEncodeTiled, fetched through
// cudaGetDriverEntryPoint.  The "tensor map" records what the product asked for; tma_load_2d() below honours it.
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
struct alignas(64) CUtensorMap {
  const unsigned char *base;
  uint64_t dim[2];          // elements, innermost first
  uint64_t row_stride;      // bytes between rows
  uint32_t box[2];          // elements, innermost first
  uint32_t swizzle_span;    // bytes: 0, 32, 64, 128
  uint32_t elem_bytes;
  unsigned char pad_[128 - 56];
};
static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
static inline CUresult simt_cuTensorMapEncodeTiled(CUtensorMap *m, CUtensorMapDataType dt, cuuint32_t rank, void *base,
                                                   const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box,
                                                   const cuuint32_t *estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  // the driver's own checks that matter for these maps
  if (dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 || rank != 2 || il != CU_TENSOR_MAP_INTERLEAVE_NONE) return 1;
  if (((uintptr_t) base) % 16 || strides[0] % 16 || estr[0] != 1 || estr[1] != 1) return 1;
  if (box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256) return 1;
  const uint32_t span = sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : 0;
  if (span && box[0] * 8 > span) return 1;              // the inner box dimension must fit the swizzle span
  m->base = (const unsigned char *) base;
  m->dim[0] = dims[0]; m->dim[1] = dims[1];
  m->row_stride = strides[0];
  m->box[0] = box[0]; m->box[1] = box[1];
  m->swizzle_span = span;
  m->elem_bytes = 8;
  return CUDA_SUCCESS;
}
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaEnableDefault = 0 };
static inline cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, int, cudaDriverEntryPointQueryResult *q) {
  const bool ok = std::strcmp(name, "cuTensorMapEncodeTiled") == 0;
  *fn = ok ? (void *) &simt_cuTensorMapEncodeTiled : nullptr;
  *q = ok ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}

#include "simt_runtime.h"

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

using std::max;
using std::min;

static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int) (uint32_t) (b >> 32); }
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int) (uint32_t) b; }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t b = ((uint64_t) (uint32_t) hi << 32) | (uint64_t) (uint32_t) lo;
  double x; std::memcpy(&x, &b, 8); return x;
}
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
// stand-in for MUFU.RCP64H: a seed good to ~2^-20 (the hardware's is ~2^-23), see tests/device_arith_host.cpp
static inline double nfh_host_rcp_seed(double x) {
  const double y = 1.0 / __hiloint2double(__double2hiint(x), 0);
  return __hiloint2double(__double2hiint(y), 0);
}
static inline int atomicOr(int *p, int v) { const int o = *p; *p = o | v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicCAS(unsigned *p, unsigned expect, unsigned v) { const unsigned o = *p; if (o == expect) *p = v; return o; }
static inline void __threadfence() {}
template <class T> static inline T __ldcg(const T *p) { return *p; }        // cache-global load: a plain load here
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 3; return cudaSuccess; }

// ---- the emulator ---------------------------------------------------------------------------------------------
namespace simt {

struct Fiber {
  ucontext_t ctx;        // portable path
  void *sp = nullptr;    // x86-64 path: saved stack pointer (callee-saved registers live on the fiber's stack)
  bool done = false;
};

struct Cta {
  std::vector<Fiber> fibers;
  std::vector<char> stacks;
  ucontext_t sched;
  void *sched_sp = nullptr;
  unsigned n_threads = 0, cur = 0;
  unsigned live = 0, bar_arrived = 0, bar_gen = 0;                 // __syncthreads
  struct Warp { unsigned live = 0, arrived = 0, gen = 0; uint64_t slot[32]; } warps[32];
  const std::function<void()> *body = nullptr;
  dim3 block;
};

Cta &cta();
unsigned char *dyn_smem();                                           // 256 KB, 1024-byte aligned
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &kernel_call);
void yield();
// Asynchronous-copy model: a bulk copy issued by one thread is NOT performed at issue.  Its destination is poisoned
// (NaN bit patterns) and the copy is queued; it happens when some thread waits on the copy's mbarrier (loads) or
// calls tma_store_wait_read (stores).  Code that reads a tile before waiting, or reuses the source of a store
// before wait_read, therefore computes with poison or stores the wrong bytes - as it may on the hardware.
struct PendingCopy { void *dst; const void *src; uint32_t bytes; uint64_t *bar; std::function<void()> run; };
std::vector<PendingCopy> &pending_loads();
std::vector<PendingCopy> &pending_stores();
void flush_loads(uint64_t *bar);
void flush_stores();
void poison(void *p, size_t bytes);
unsigned long long launches();                                       // kernels launched so far
unsigned long long switches();                                       // fiber switches so far

inline void syncthreads() {
  Cta &c = cta();
  const unsigned gen = c.bar_gen;
  if (++c.bar_arrived >= c.live) { c.bar_arrived = 0; c.bar_gen++; return; }
  while (c.bar_gen == gen) yield();
}
inline void syncwarp() {
  Cta &c = cta();
  Cta::Warp &w = c.warps[c.cur >> 5];
  const unsigned gen = w.gen;
  if (++w.arrived >= w.live) { w.arrived = 0; w.gen++; return; }
  while (w.gen == gen) yield();
}
inline void named_barrier(int id, int n_threads) {                       // bar.sync id, n_threads
  struct Named { int arrived = 0; unsigned gen = 0; };
  static thread_local Named bars[16];
  Named &b = bars[id & 15];
  const unsigned gen = b.gen;
  if (++b.arrived >= n_threads) { b.arrived = 0; b.gen++; return; }
  while (b.gen == gen) yield();
}
template <class T> inline T exchange(T v, int src_lane) {             // src_lane outside 0..31: keep the own value
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  Cta &c = cta();
  Cta::Warp &w = c.warps[c.cur >> 5];
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  w.slot[c.cur & 31] = bits;
  syncwarp();
  T r = v;
  if (src_lane >= 0 && src_lane < 32) std::memcpy(&r, &w.slot[src_lane], sizeof(T));
  syncwarp();
  return r;
}

}  // namespace simt

static inline void __syncthreads() { simt::syncthreads(); }
static inline void __nanosleep(unsigned) { simt::yield(); }             // a polling thread lets the others run
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::syncwarp(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return simt::exchange(v, src & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return simt::exchange(v, (int) (threadIdx.x & 31) - (int) d); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return simt::exchange(v, (int) (threadIdx.x & 31) + (int) d); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return simt::exchange(v, (int) (threadIdx.x & 31) ^ m); }
static inline int __any_sync(unsigned, int pred) {
  simt::Cta &c = simt::cta();
  simt::Cta::Warp &w = c.warps[c.cur >> 5];
  w.slot[c.cur & 31] = pred ? 1 : 0;
  simt::syncwarp();
  int any = 0;
  const unsigned base = (c.cur >> 5) * 32;
  for (unsigned l = 0; l < 32 && base + l < c.n_threads; l++)
    if (!c.fibers[base + l].done && w.slot[l]) any = 1;
  simt::syncwarp();
  return any;
}

// ---- nfh_tma.cuh on the host: bulk copies are synchronous, an mbarrier is (phase << 63) | pending bytes -----------
namespace nfh {
static inline void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { *bar += bytes; }
static inline void mbar_complete_tx(uint64_t *bar, uint32_t bytes) {
  if ((*bar & 0x7fffffffffffffffull) < bytes) { std::fprintf(stderr, "simt: mbarrier completes more bytes than expected\n"); std::abort(); }
  *bar -= bytes;
  if ((*bar & 0x7fffffffffffffffull) == 0) *bar ^= 0x8000000000000000ull;       // phase completes
}
static inline void mbar_wait(uint64_t *bar, uint32_t parity) {
  simt::flush_loads(bar);                          // the copies of this barrier land now, not at issue
  while ((uint32_t) (*bar >> 63) == parity) { simt::yield(); simt::flush_loads(bar); }
}
static inline void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  if (bytes % 16 || ((uintptr_t) smem_dst | (uintptr_t) gmem_src) % 16) { std::fprintf(stderr, "simt: misaligned bulk copy\n"); std::abort(); }
  simt::poison(smem_dst, bytes);
  simt::pending_loads().push_back({smem_dst, gmem_src, bytes, bar, [=]() {
    std::memcpy(smem_dst, gmem_src, bytes);
    mbar_complete_tx(bar, bytes);
  }});
}
// box of box[1] rows x box[0] elements at element coordinates (x, y); out-of-bounds elements read as zero; with a
// swizzle span the 16-byte chunk index inside each span-sized row is XORed with the row index modulo the number of
// chunks (CU_TENSOR_MAP_SWIZZLE_32B / 64B / 128B for spans that equal the box row, as these maps use them)
static inline void tma_load_2d(void *smem_dst, const void *tensor_map, int x, int y, uint64_t *bar) {
  const CUtensorMap m = *reinterpret_cast<const CUtensorMap *>(tensor_map);
  const uint32_t row_bytes = m.box[0] * m.elem_bytes;
  const uint32_t mask = m.swizzle_span ? m.swizzle_span / 16 - 1 : 0;
  if (m.swizzle_span && ((uintptr_t) smem_dst) % (8 * m.swizzle_span)) { std::fprintf(stderr, "simt: swizzled box not aligned to its pattern\n"); std::abort(); }
  simt::poison(smem_dst, (size_t) m.box[1] * row_bytes);
  simt::pending_loads().push_back({smem_dst, nullptr, m.box[1] * row_bytes, bar, [=]() {
    unsigned char *dst = reinterpret_cast<unsigned char *>(smem_dst);
    for (uint32_t r = 0; r < m.box[1]; r++)
      for (uint32_t c = 0; c < m.box[0]; c++) {
        uint32_t off = r * row_bytes + c * m.elem_bytes;
        off ^= ((off >> 7) & mask) << 4;
        const int64_t gx = (int64_t) x + c, gy = (int64_t) y + r;
        double v = 0.0;
        if (gx >= 0 && gy >= 0 && (uint64_t) gx < m.dim[0] && (uint64_t) gy < m.dim[1])
          std::memcpy(&v, m.base + (uint64_t) gy * m.row_stride + (uint64_t) gx * m.elem_bytes, 8);
        std::memcpy(dst + off, &v, 8);
      }
    mbar_complete_tx(bar, m.box[1] * row_bytes);
  }});
}
static inline void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
  if (bytes % 16 || ((uintptr_t) smem_src | (uintptr_t) gmem_dst) % 16) { std::fprintf(stderr, "simt: misaligned bulk store\n"); std::abort(); }
  simt::pending_stores().push_back({gmem_dst, smem_src, bytes, nullptr, [=]() { std::memcpy(gmem_dst, smem_src, bytes); }});
}
static inline void tma_store_wait_read() { simt::flush_stores(); }
// L2 eviction hints of the single-launch E-step: no cache here
static inline uint64_t l2_policy_evict_last() { return 0; }
static inline uint64_t l2_policy_evict_first() { return 0; }
static inline void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t) { tma_load_1d(smem_dst, gmem_src, bytes, bar); }
static inline void tma_store_1d_hint(void *gmem_dst, const void *smem_src, uint32_t bytes, uint64_t) { tma_store_1d(gmem_dst, smem_src, bytes); }
static inline void fence_async_shared() {}
}  // namespace nfh

// simt_runtime.h - TEST INFRASTRUCTURE ONLY: the slice of the CUDA runtime API that nfh_ctx.cu (the C ABI layer of
// the product) calls, on host memory, so that tests/simt/run_gpu_suite_emulated.py can build the WHOLE library
// - kernels, launchers and the C ABI - for the SIMT emulator and run the `-m gpu` tests' small cases without a GPU.
// "Device" memory is host memory, one "device" exists per requested ordinal, streams execute at once, peer access
// and IPC handles are plain pointers.  A pre-flight checker for the round-end GPU run; not a backend.
#pragma once

#include <chrono>
#include <cstdlib>
#include <cstring>

enum { cudaErrorMemoryAllocation = 2, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostRegisterPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1, cudaDevAttrMultiProcessorCount = 16 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct simt_event { std::chrono::steady_clock::time_point t; };
typedef simt_event *cudaEvent_t;

int simt_device_count();                       // SIMT_DEVICES in the environment, default 1
int simt_sm_count();                           // SIMT_SMS, default 4: small grids, several tiles per persistent CTA

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = simt_device_count(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < simt_device_count() ? cudaSuccess : 101; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int) { *v = attr == cudaDevAttrMultiProcessorCount ? simt_sm_count() : 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
// "Device" memory.  Normally plain host memory; with SIMT_IPC=1 in the environment every allocation is a POSIX
// shared-memory segment, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle work ACROSS PROCESSES (one emulated rank per
// process under torch.distributed/gloo: tests/simt/run_multi_rank_emulated.py) - the kernels of one rank then store
// into the windows of another exactly as they do over NVLink.
void *simt_device_alloc(size_t bytes);                     // simt.cpp
void simt_device_free(void *p);
int simt_ipc_export(void *p, char name_out[64]);           // 0 = ok
void *simt_ipc_open(const char name[64]);
void simt_ipc_close(void *p);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) {
  *p = (T *) simt_device_alloc(bytes);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { simt_device_free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) {
  *p = (T *) std::aligned_alloc(1024, (bytes + 1023) / 1024 * 1024);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t r = 0; r < height; r++) std::memmove((char *) d + r * dpitch, (const char *) s + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t) std::malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new simt_event; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
// every pointer a caller hands in is host memory here: "not known to the runtime" (the pageable-host path of nfh_upload_gl)
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) {
  a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = nullptr;
  return cudaSuccess;
}
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { std::memset(h, 0, sizeof *h); return simt_ipc_export(p, h->reserved) == 0 ? cudaSuccess : 1; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { *p = simt_ipc_open(h.reserved); return *p ? cudaSuccess : 1; }
static inline cudaError_t cudaIpcCloseMemHandle(void *p) { simt_ipc_close(p); return cudaSuccess; }

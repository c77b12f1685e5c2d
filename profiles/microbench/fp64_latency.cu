// FP64 pipe microbenchmark for B200: DFMA throughput as a function of
// independent chains per warp (ILP) and warps per SM, plus the dependent-issue
// latency of DFMA and of the MUFU.RCP64H + Newton reciprocal.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chains(double *sink, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  const double m = 0.999999999, c = 1e-12;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 12345.678) sink[0] = s;
}

// DFMA whose three source operands are all distinct registers (no operand reuse):
// tests whether register-file bandwidth caps the issue rate below 1 per 2 cycles.
template <int ILP>
__global__ void chains3(double *sink, int iters) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 1e-9 + i; y[i] = 0.999999 + i * 1e-9; z[i] = 1e-12 * (i + 1); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], y[i], z[i]);
#pragma unroll
    for (int i = 0; i < ILP; i++) y[i] = fma(y[i], z[i], x[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i] + y[i];
  if (s == 12345.678) sink[0] = s;
}

template <int ILP>
double run3(int warps_per_sm, int sms, double *sink) {
  const int iters = 1 << 13;
  int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
  int ctas_per_sm = (warps_per_sm * 32) / threads;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  chains3<ILP><<<sms * ctas_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e0);
  chains3<ILP><<<sms * ctas_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double) sms * ctas_per_sm * threads * iters * ILP * 2.0 / (ms * 1e-3);
}

// K independent "individuals": one MUFU.RCP64H seed + NF dependent-ish DFMAs each, as in the
// frequency kernel's inner loop.  Reports DFMA/s to see what the MUFU costs the FP64 pipe.
template <int K, int NF, bool MUFU>
__global__ void rcp_mix(double *sink, int iters) {
  double x[K], acc[K];
#pragma unroll
  for (int i = 0; i < K; i++) { x[i] = 1.0 + threadIdx.x * 1e-3 + i; acc[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < K; i++) {
      double y;
      if (MUFU) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
      else y = x[i] * 0.5;
#pragma unroll
      for (int f = 0; f < NF; f++) y = fma(y, x[i], 1e-9);
      acc[i] += y;
      x[i] = fma(x[i], 0.999999, 1e-7);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; i++) s += acc[i];
  if (s == 12345.678) sink[0] = s;
}

template <int K, int NF, bool MUFU>
double run_mix(int warps_per_sm, int sms, double *sink) {
  const int iters = 1 << 11;
  int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
  int ctas_per_sm = (warps_per_sm * 32) / threads;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rcp_mix<K, NF, MUFU><<<sms * ctas_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e0);
  rcp_mix<K, NF, MUFU><<<sms * ctas_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // FP64 instructions per individual: NF fma + 1 add + 1 fma (+1 mul when no MUFU)
  return (double) sms * ctas_per_sm * threads * iters * K * (NF + 2 + (MUFU ? 0 : 1)) / (ms * 1e-3);
}

__global__ void latency_dfma(long long *out, double *sink, int iters) {
  double x = threadIdx.x * 1e-9;
  const double m = 0.999999999, c = 1e-12;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    x = fma(x, m, c); x = fma(x, m, c); x = fma(x, m, c); x = fma(x, m, c);
    x = fma(x, m, c); x = fma(x, m, c); x = fma(x, m, c); x = fma(x, m, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (x == 12345.678) sink[0] = x;
}

__global__ void latency_rcp(long long *out, double *sink, int iters) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    double y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
    x = y + 1.5;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (x == 12345.678) sink[0] = x;
}

template <int ILP>
double run(int warps_per_sm, int sms, double *sink) {
  const int iters = 1 << 14;
  int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
  int ctas_per_sm = (warps_per_sm * 32) / threads;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  chains<ILP><<<sms * ctas_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e0);
  chains<ILP><<<sms * ctas_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double fma_per_s = (double) sms * ctas_per_sm * threads * iters * ILP / (ms * 1e-3);
  return fma_per_s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double *sink; long long *out;
  cudaMalloc(&sink, 8); cudaMalloc(&out, 8);
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  printf("DFMA/s (1e12) by warps/SM (rows) and ILP (cols 1 2 4 8)\n");
  for (int w : {4, 8, 12, 16, 24, 32, 64}) {
    printf("warps/SM %2d: %7.2f %7.2f %7.2f %7.2f\n", w, run<1>(w, sms, sink) / 1e12, run<2>(w, sms, sink) / 1e12,
           run<4>(w, sms, sink) / 1e12, run<8>(w, sms, sink) / 1e12);
  }
  printf("3-distinct-operand DFMA/s (1e12), ILP 8: warps/SM 8: %.2f  16: %.2f  32: %.2f  64: %.2f\n",
         run3<8>(8, sms, sink) / 1e12, run3<8>(16, sms, sink) / 1e12, run3<8>(32, sms, sink) / 1e12,
         run3<8>(64, sms, sink) / 1e12);
  printf("FP64 instr/s (1e12) with 1 MUFU.RCP64H per 9 FP64 (K=13 chains), 8 warps/SM: %.2f ; 16 warps/SM: %.2f\n",
         run_mix<13, 7, true>(8, sms, sink) / 1e12, run_mix<13, 7, true>(16, sms, sink) / 1e12);
  printf("same without the MUFU (DMUL instead), 8 warps/SM: %.2f ; 16 warps/SM: %.2f\n",
         run_mix<13, 7, false>(8, sms, sink) / 1e12, run_mix<13, 7, false>(16, sms, sink) / 1e12);
  printf("1 MUFU per 18 FP64, 8 warps/SM: %.2f ; 1 MUFU per 36 FP64: %.2f\n",
         run_mix<13, 16, true>(8, sms, sink) / 1e12, run_mix<13, 34, true>(8, sms, sink) / 1e12);
  long long h;
  latency_dfma<<<1, 32>>>(out, sink, 1000);
  latency_dfma<<<1, 32>>>(out, sink, 1000);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.2f cycles\n", (double) h / 8000.0);
  latency_rcp<<<1, 32>>>(out, sink, 1000);
  latency_rcp<<<1, 32>>>(out, sink, 1000);
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("rcp(MUFU + 3 DFMA) + DADD chain latency: %.2f cycles per iteration\n", (double) h / 1000.0);
  return 0;
}

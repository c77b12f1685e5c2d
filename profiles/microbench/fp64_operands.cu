// Does the register file cap FP64 issue below one warp instruction per 2 cycles?  26 independent chains per
// thread (enough ILP to hide the 8.4-cycle latency), differing only in how many 64-bit source operands of each
// DFMA/DMUL come fresh from the register file rather than from the operand-reuse cache.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 26;

template <int P>
__global__ void __launch_bounds__(128) rf(const double *src, double *out, int iters) {
  double a[N], b[N], x[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    a[k] = src[threadIdx.x + 128 * k]; b[k] = src[threadIdx.x + 128 * (k + N)]; x[k] = src[threadIdx.x + 128 * (k + 2 * N)];
  }
  double t = src[7];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      if (P == 0) x[k] = fma(x[k], t, 1e-9);          // 1 fresh operand (t reused, constant immediate)
      if (P == 1) x[k] = fma(a[k], t, x[k]);          // 2 fresh
      if (P == 2) x[k] = fma(a[k], b[k], x[k]);       // 3 fresh
      if (P == 3) x[k] = x[k] * a[k];                 // DMUL, 2 fresh
      if (P == 4) x[k] = x[k] + a[k];                 // DADD, 2 fresh
      if (P == 5) x[k] = fma(a[k], b[(k + 1) % N], x[k]);   // 3 fresh, other register pairing
    }
    t = fma(t, 0.999999, 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < N; k++) s += x[k];
  if (s == 12345.678) out[0] = s;
}

template <int P>
double run(int ctas_per_sm, const double *src, double *out) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rf<P><<<148 * ctas_per_sm, 128>>>(src, out, iters);
  cudaEventRecord(e0);
  rf<P><<<148 * ctas_per_sm, 128>>>(src, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 * 1.965e9 / ((double) iters * N * ctas_per_sm);     // cycles per warp instruction per scheduler
}

int main() {
  double *src, *out;
  cudaMalloc(&src, 128 * 3 * N * 8); cudaMalloc(&out, 8);
  double h[128 * 3 * N];
  for (int i = 0; i < 128 * 3 * N; i++) h[i] = 0.9 + 1e-4 * (i % 97);
  cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  printf("cycles per FP64 warp instruction per scheduler (1 / 2 / 4 warps per scheduler)\n");
  printf("DFMA 1 fresh operand : %.2f %.2f %.2f\n", run<0>(1, src, out), run<0>(2, src, out), run<0>(4, src, out));
  printf("DFMA 2 fresh operands: %.2f %.2f %.2f\n", run<1>(1, src, out), run<1>(2, src, out), run<1>(4, src, out));
  printf("DFMA 3 fresh operands: %.2f %.2f %.2f\n", run<2>(1, src, out), run<2>(2, src, out), run<2>(4, src, out));
  printf("DFMA 3 fresh, shifted: %.2f %.2f %.2f\n", run<5>(1, src, out), run<5>(2, src, out), run<5>(4, src, out));
  printf("DMUL 2 fresh operands: %.2f %.2f %.2f\n", run<3>(1, src, out), run<3>(2, src, out), run<3>(4, src, out));
  printf("DADD 2 fresh operands: %.2f %.2f %.2f\n", run<4>(1, src, out), run<4>(2, src, out), run<4>(4, src, out));
  return 0;
}

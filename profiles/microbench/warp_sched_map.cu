// Which warps of a CTA share a scheduler (and its FP64 pipe)?  Warp 0 and warp x run a pipe-saturating DFMA
// loop, the rest exit; the pair takes twice as long when it shares a scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_sched_map warp_sched_map.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) pair(double *out, int x, int iters, unsigned *slots) {
  const int warp = threadIdx.x >> 5;
  unsigned hw;
  asm("mov.u32 %0, %%warpid;" : "=r"(hw));
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) slots[warp] = hw;
  if (warp != 0 && warp != x) return;
  double v[8];
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = fma(v[i], 0.999999, 1e-9);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  double *out; unsigned *slots;
  cudaMalloc(&out, 8); cudaMalloc(&slots, 32);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int x = 0; x < 8; x++) {
    pair<<<148, 256>>>(out, x, 20000, slots);
    cudaEventRecord(e0);
    pair<<<148, 256>>>(out, x, 20000, slots);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("warps 0 and %d: %.3f ms\n", x, ms);
  }
  unsigned h[8]; cudaMemcpy(h, slots, 32, cudaMemcpyDeviceToHost);
  printf("%%warpid of warps 0..7 (CTA 0):"); for (int i = 0; i < 8; i++) printf(" %u", h[i]); printf("\n");
  return 0;
}

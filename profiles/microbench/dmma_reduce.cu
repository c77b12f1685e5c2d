// Can the FP64 tensor-core instruction (mma.sync.m8n8k4.f64, SASS DMMA) shorten the lane reduction at the end of
// every pass of the frequency EM?  Summing one double over the 32 lanes of a warp takes five shuffle levels
// (5 x (SHFL + DADD), each level waits for the one before); two DMMAs with a matrix of ones do the same sum:
//   stage 1: B[k][n] = value of lane 4n+k, A = ones  ->  D[m][n] = sum of lanes 4n..4n+3; lane (g,t) holds n = 2t, 2t+1
//   one DADD, then stage 2: B[k][n] = p_k (depends on t only), A = ones or a 0/1 selection of the lane group
// Measured here: latency of a dependent DMMA chain, issue cost of independent DMMAs alone and next to DFMAs (do
// they share the FP64 pipe?), and the latency of the two reductions in a dependent loop, with their results compared.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_reduce dmma_reduce.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// sum over groups of G lanes (G = 8, 16, 32), result in every lane of the group
template <int G>
__device__ __forceinline__ double dmma_group_sum(double v) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double d0, d1;
  dmma(d0, d1, 1.0, v, 0.0, 0.0);
  const double p = d0 + d1;                                  // sum of lanes 8t .. 8t+7
  double sel = 1.0;
  if (G == 8) sel = (t == (g >> 1)) ? 1.0 : 0.0;
  if (G == 16) sel = ((t >> 1) == (g >> 2)) ? 1.0 : 0.0;
  dmma(d0, d1, sel, p, 0.0, 0.0);
  return d0;
}

template <int G>
__device__ __forceinline__ double shfl_group_sum(double v) {
#pragma unroll
  for (int m = 1; m < G; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__global__ void check(double *out) {
  const int lane = threadIdx.x & 31;
  const double v = 1.0 + 0.37 * lane + 1e-9 * lane * lane;
  out[lane] = dmma_group_sum<8>(v) - shfl_group_sum<8>(v);
  out[32 + lane] = dmma_group_sum<16>(v) - shfl_group_sum<16>(v);
  out[64 + lane] = dmma_group_sum<32>(v) - shfl_group_sum<32>(v);
  out[96 + lane] = shfl_group_sum<8>(v);
}

// MODE 0: dependent DMMA chain; 1: 8 independent DMMA chains; 2: 8 independent DFMA chains x 8 (same FMA count as
// mode 1); 3: modes 1 and 2 interleaved; 4/5: dependent loop of group sums of two values (X, Z) by shuffles / by DMMA,
// with a short FMA chain between them as in the pass loop
template <int MODE, int G>
__global__ void __launch_bounds__(128) bench(double *out, int iters) {
  const int lane = threadIdx.x & 31;
  double x[8], y[8], f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { x[k] = 1.0 + 1e-3 * (lane + k); y[k] = 0.5 + 1e-3 * k; f[k] = 1.0 + 1e-6 * k; }
  double X = 1.0 + 1e-3 * lane, Z = 2.0 - 1e-3 * lane;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { dmma(x[0], x[1], 1.0, x[0], x[1], x[0]); }
    if (MODE == 1 || MODE == 3) {
#pragma unroll
      for (int k = 0; k < 8; k++) dmma(x[k], y[k], f[k], 1e-3, x[k], y[k]);
    }
    if (MODE == 2 || MODE == 3) {
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = fma(f[k], 0.999999, 1e-7);
    }
    if (MODE == 4) {
      const double sx = shfl_group_sum<G>(X), sz = shfl_group_sum<G>(Z);
      X = fma(sx, 1e-3, 1.0 + 1e-3 * lane); Z = fma(sz, 1e-3, 2.0);
    }
    if (MODE == 5) {
      const double sx = dmma_group_sum<G>(X), sz = dmma_group_sum<G>(Z);
      X = fma(sx, 1e-3, 1.0 + 1e-3 * lane); Z = fma(sz, 1e-3, 2.0);
    }
  }
  double s = X + Z;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k] + y[k] + f[k];
  if (s == 12345.678) out[0] = s;
}

template <int MODE, int G>
double cycles(int ctas_per_sm, double *out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE, G><<<148 * ctas_per_sm, 128>>>(out, iters);
  cudaEventRecord(e0);
  bench<MODE, G><<<148 * ctas_per_sm, 128>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 * 1.965e9 / iters;       // cycles per loop iteration (per scheduler: one warp of each CTA)
}

int main() {
  double *out;
  cudaMalloc(&out, 128 * 8);
  check<<<1, 32>>>(out);
  double h[128];
  cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  double e8 = 0, e16 = 0, e32 = 0;
  for (int i = 0; i < 32; i++) { e8 = fmax(e8, fabs(h[i])); e16 = fmax(e16, fabs(h[32 + i])); e32 = fmax(e32, fabs(h[64 + i])); }
  printf("max |DMMA sum - shuffle sum|: G=8 %.3g  G=16 %.3g  G=32 %.3g  (group sums ~ %.1f)\n", e8, e16, e32, h[96]);
  const int it = 20000;
  printf("cycles per iteration at 1 / 2 / 4 warps per scheduler\n");
  printf("dependent DMMA chain (latency)        : %.1f\n", cycles<0, 32>(1, out, it));
  printf("8 independent DMMAs                   : %.1f %.1f %.1f\n", cycles<1, 32>(1, out, it), cycles<1, 32>(2, out, it), cycles<1, 32>(4, out, it));
  printf("64 independent DFMAs                  : %.1f %.1f %.1f\n", cycles<2, 32>(1, out, it), cycles<2, 32>(2, out, it), cycles<2, 32>(4, out, it));
  printf("8 DMMAs + 64 DFMAs interleaved        : %.1f %.1f %.1f\n", cycles<3, 32>(1, out, it), cycles<3, 32>(2, out, it), cycles<3, 32>(4, out, it));
  printf("reduction of (X,Z), shuffles, G=8/16/32: %.1f %.1f %.1f\n", cycles<4, 8>(1, out, it), cycles<4, 16>(1, out, it), cycles<4, 32>(1, out, it));
  printf("reduction of (X,Z), DMMA,     G=8/16/32: %.1f %.1f %.1f\n", cycles<5, 8>(1, out, it), cycles<5, 16>(1, out, it), cycles<5, 32>(1, out, it));
  printf("... with 2 warps per scheduler, shuffles: %.1f %.1f %.1f\n", cycles<4, 8>(2, out, it), cycles<4, 16>(2, out, it), cycles<4, 32>(2, out, it));
  printf("... with 2 warps per scheduler, DMMA    : %.1f %.1f %.1f\n", cycles<5, 8>(2, out, it), cycles<5, 16>(2, out, it), cycles<5, 32>(2, out, it));
  return 0;
}

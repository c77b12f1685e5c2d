// Isolates the per-pass body of the frequency EM (K individuals per lane, 9 FP64 + shared MUFU each)
// from its per-pass tail (shuffle reduction, division, vote) to see what each part costs on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o freq_body freq_body.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_pos(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}

template <int K>
__device__ __forceinline__ void reciprocals(const double (&S)[K], double (&inv)[K]) {
#pragma unroll
  for (int k = 0; k + 3 < K; k += 4) {
    const double p01 = S[k] * S[k + 1], p23 = S[k + 2] * S[k + 3];
    const double r = rcp_pos(p01 * p23);
    const double r01 = r * p23, r23 = r * p01;
    inv[k] = r01 * S[k + 1]; inv[k + 1] = r01 * S[k];
    inv[k + 2] = r23 * S[k + 3]; inv[k + 3] = r23 * S[k + 2];
  }
  constexpr int rem = K % 4, k = K - rem;
  if (rem == 1) inv[k] = rcp_pos(S[k]);
}

// MODE 0: body only (freq advanced by a cheap dependent op); 1: + shuffle reduction over 8 lanes;
// 2: + division and stop test + vote (the real loop)
template <int K, int MODE>
__global__ void __launch_bounds__(128) body(const double *coef, double *out, int passes) {
  double a0[K], a2[K], hh[K], na[K], nv[K], da[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double *c = coef + ((size_t) (blockIdx.x * 128 + threadIdx.x) * K + k) * 6;   // six independent values
    a0[k] = c[0]; a2[k] = c[1]; hh[k] = c[2]; na[k] = c[3]; nv[k] = c[4]; da[k] = c[5];
  }
  double freq = 0.01, num = 0, den = 0;
  bool active = true;
  int p = 0;
  while (MODE == 2 ? __any_sync(0xffffffffu, active) : p < passes) {
    const double omf = 1.0 - freq, u = omf * omf, v = freq * freq, a = omf * freq;
    double S[K], inv[K];
#pragma unroll
    for (int k = 0; k < K; k++) S[k] = fma(a0[k], u, fma(a2[k], v, hh[k] * a));
    reciprocals<K>(S, inv);
    double A1 = 0, A2 = 0, A3 = 0, B1 = 0, B2 = 0, B3 = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (k & 1) { B1 = fma(na[k], inv[k], B1); B2 = fma(nv[k], inv[k], B2); B3 = fma(da[k], inv[k], B3); }
      else { A1 = fma(na[k], inv[k], A1); A2 = fma(nv[k], inv[k], A2); A3 = fma(da[k], inv[k], A3); }
    }
    double pn = fma(a, A1 + B1, v * (A2 + B2)), pd = a * (A3 + B3);
    if (MODE >= 1) {
#pragma unroll
      for (int m = 1; m < 8; m <<= 1) { pn += __shfl_xor_sync(0xffffffffu, pn, m); pd += __shfl_xor_sync(0xffffffffu, pd, m); }
    }
    pd += 150.0;
    p++;
    if (MODE == 2) {
      if (active) {
        num += pn; den += pd;
        const double before = freq;
        freq = num * rcp_pos(den);
        active = (fabs(before - freq) > 1e-30) && (p < passes - (int) (threadIdx.x >> 3 & 3));   // lanes of different sites stop apart
      }
    } else {
      num += pn; den += pd;
      freq = fma(num, 1e-7, 0.2) + den * 1e-9;     // cheap dependent update, no division
    }
  }
  if (freq == 12345.678) out[0] = freq + num;
}

template <int K, int MODE>
double run(int grid, double *coef, double *out) {
  const int passes = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body<K, MODE><<<grid, 128>>>(coef, out, passes);
  cudaEventRecord(e0);
  body<K, MODE><<<grid, 128>>>(coef, out, passes);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 / passes;     // seconds per pass (all resident CTAs in parallel)
}

int main() {
  double *coef, *out;
  const size_t n_coef = (size_t) 592 * 128 * 13 * 6;
  cudaMalloc(&coef, n_coef * 8); cudaMalloc(&out, 8);
  {
    double *h = (double *) malloc(n_coef * 8);
    for (size_t i = 0; i < n_coef; i++) h[i] = 0.05 + 0.9 * ((i * 2654435761u) % 1000) / 1000.0;
    cudaMemcpy(coef, h, n_coef * 8, cudaMemcpyHostToDevice);
    free(h);
  }
  const double clk = 1.965e9;
  printf("cycles per pass (wall), K=13 G=8 shape, 128-thread CTAs\n");
  for (int ctas_per_sm : {1, 2}) {
    int grid = 148 * ctas_per_sm;
    printf("%d CTA/SM: body %.0f  +shuffles %.0f  +division/vote %.0f\n", ctas_per_sm, run<13, 0>(grid, coef, out) * clk,
           run<13, 1>(grid, coef, out) * clk, run<13, 2>(grid, coef, out) * clk);
  }
  printf("K=7 (coefficients strided as K=13): 1 CTA/SM full %.0f ; 2 CTA/SM %.0f ; 4 CTA/SM %.0f\n",
         run<7, 2>(148, coef, out) * clk, run<7, 2>(296, coef, out) * clk, run<7, 2>(592, coef, out) * clk);
  return 0;
}

// Where does a frequency-EM pass spend its time?  Times the pass loop of nfh_freq.cu (odds form,
// 8 FP64 per individual) for several lane-group shapes and occupancies, with parts of the per-pass tail
// removed, and with one warp per scheduler to read the latency of a single pass.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o freq_pass freq_pass.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <bool REFINE>
__device__ __forceinline__ double rcp_pos(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  if (REFINE) y = fma(fma(-x, y, 1.0), y, y);
  return y;
}

template <int K>
__device__ __forceinline__ void reciprocals(const double (&S)[K], double (&inv)[K]) {
#pragma unroll
  for (int k = 0; k + 3 < K; k += 4) {
    const double p01 = S[k] * S[k + 1], p23 = S[k + 2] * S[k + 3];
    const double r = rcp_pos<false>(p01 * p23);
    const double r01 = r * p23, r23 = r * p01;
    inv[k] = r01 * S[k + 1]; inv[k + 1] = r01 * S[k];
    inv[k + 2] = r23 * S[k + 3]; inv[k + 3] = r23 * S[k + 2];
  }
  constexpr int rem = K % 4, k = K - rem;
  if (rem == 1) {
    inv[k] = rcp_pos<false>(S[k]);
  } else if (rem == 2) {
    const double r = rcp_pos<false>(S[k] * S[k + 1]);
    inv[k] = r * S[k + 1]; inv[k + 1] = r * S[k];
  } else if (rem == 3) {
    const double p01 = S[k] * S[k + 1];
    const double r = rcp_pos<false>(p01 * S[k + 2]);
    const double r01 = r * S[k + 2];
    inv[k] = r01 * S[k + 1]; inv[k + 1] = r01 * S[k]; inv[k + 2] = r * p01;
  }
}

// MODE 0: body only; 1: + shuffle reduction over G lanes; 2: + the two divisions, stop test and vote;
// 3: as 2 but without the shuffles (division + vote only)
template <int K, int G, int MODE, int OCC>
__global__ void __launch_bounds__(128, OCC) pass_loop(const double *coef, double *out, int passes) {
  double a0[K], a2[K], hh[K], na[K], nv[K], da[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double *c = coef + ((size_t) ((blockIdx.x % 592) * 128 + threadIdx.x) * 13 + k) * 6;
    a0[k] = c[0]; a2[k] = c[1]; hh[k] = c[2]; na[k] = c[3]; nv[k] = c[4]; da[k] = c[5];
  }
  double freq = 0.01, odds = 0.01 / 0.99, num = 0, den = 0;
  bool active = true;
  int p = 0;
  while (MODE >= 2 ? __any_sync(0xffffffffu, active) : p < passes) {
    double S[K], inv[K];
#pragma unroll
    for (int k = 0; k < K; k++) S[k] = fma(fma(a2[k], odds, hh[k]), odds, a0[k]);
    reciprocals<K>(S, inv);
    double A1 = 0, A2 = 0, A3 = 0, B1 = 0, B2 = 0, B3 = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (k & 1) { B1 = fma(na[k], inv[k], B1); B2 = fma(nv[k], inv[k], B2); B3 = fma(da[k], inv[k], B3); }
      else { A1 = fma(na[k], inv[k], A1); A2 = fma(nv[k], inv[k], A2); A3 = fma(da[k], inv[k], A3); }
    }
    double pn = odds * fma(odds, A2 + B2, A1 + B1), pd = odds * (A3 + B3);
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int m = 1; m < G; m <<= 1) { pn += __shfl_xor_sync(0xffffffffu, pn, m); pd += __shfl_xor_sync(0xffffffffu, pd, m); }
    }
    pd += 150.0;
    p++;
    if (MODE >= 2) {
      if (active) {
        num += pn; den += pd;
        const double before = freq;
        odds = num * rcp_pos<true>(den - num);
        freq = num * rcp_pos<true>(den);
        active = (fabs(before - freq) > 1e-30) && (p < passes - (int) (threadIdx.x / G % (32 / G)));
      }
    } else {
      num += pn; den += pd;
      odds = fma(num, 1e-7, 0.2) + den * 1e-9;     // cheap dependent update, no division
    }
  }
  if (odds == 12345.678) out[0] = freq + num;
}


// ---- two warps per scheduler out of phase --------------------------------------------------------
// 256-thread CTA, one per SM: warps w and w+4 sit on the same scheduler.  SYNC 0: nothing (they run in
// lockstep); 1: warps 4-7 start half a pass late; 2: named-barrier ping-pong, permission handed over at
// the end of the body; 3: handed over after the reciprocals (the partner's S phase overlaps our sums).
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int K, int G, int SYNC>
__global__ void __launch_bounds__(256, 1) pingpong(const double *coef, double *out, int passes, int delay) {
  double a0[K], a2[K], hh[K], na[K], nv[K], da[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double *c = coef + ((size_t) ((blockIdx.x % 296) * 256 + threadIdx.x) * 13 + k) * 6;
    a0[k] = c[0]; a2[k] = c[1]; hh[k] = c[2]; na[k] = c[3]; nv[k] = c[4]; da[k] = c[5];
  }
  const int warp = threadIdx.x >> 5, pair = warp & 3, role = warp >> 2;
  const int mine = 1 + 2 * pair + role, theirs = 1 + 2 * pair + (role ^ 1);
  double freq = 0.01, odds = 0.01 / 0.99, num = 0, den = 0;
  bool active = true;
  int p = 0;
  __syncthreads();
  if (SYNC == 1 && role == 1) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
  if (SYNC >= 2 && role == 1) bar_arrive(theirs, 64);       // role 0 goes first
  while (p < passes) {
    if (SYNC >= 2) { bar_sync(mine, 64); asm volatile("" : "+d"(odds)); }   // the body may not start before the barrier
    double S[K], inv[K];
#pragma unroll
    for (int k = 0; k < K; k++) S[k] = fma(fma(a2[k], odds, hh[k]), odds, a0[k]);
    reciprocals<K>(S, inv);
    if (SYNC == 3) {
#pragma unroll
      for (int k = 0; k < K; k++) asm volatile("" : "+d"(inv[k]));
      if (!(role == 1 && p == passes - 1)) bar_arrive(theirs, 64);
    }
    double A1 = 0, A2 = 0, A3 = 0, B1 = 0, B2 = 0, B3 = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (k & 1) { B1 = fma(na[k], inv[k], B1); B2 = fma(nv[k], inv[k], B2); B3 = fma(da[k], inv[k], B3); }
      else { A1 = fma(na[k], inv[k], A1); A2 = fma(nv[k], inv[k], A2); A3 = fma(da[k], inv[k], A3); }
    }
    double pn = odds * fma(odds, A2 + B2, A1 + B1), pd = odds * (A3 + B3);
    if (SYNC == 2) {
      asm volatile("" : "+d"(pn), "+d"(pd));                              // the body is issued before the hand-over
      if (!(role == 1 && p == passes - 1)) bar_arrive(theirs, 64);
    }
#pragma unroll
    for (int m = 1; m < G; m <<= 1) { pn += __shfl_xor_sync(0xffffffffu, pn, m); pd += __shfl_xor_sync(0xffffffffu, pd, m); }
    pd += 150.0;
    p++;
    if (active) {
      num += pn; den += pd;
      const double before = freq;
      odds = num * rcp_pos<true>(den - num);
      freq = num * rcp_pos<true>(den);
      active = (fabs(before - freq) > 1e-30);
    }
  }
  if (odds == 12345.678) out[0] = freq + num;
}

template <int K, int G, int SYNC>
double run_pp(double *coef, double *out, int delay) {
  const int passes = 1000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  pingpong<K, G, SYNC><<<148, 256>>>(coef, out, passes, delay);
  cudaEventRecord(e0);
  pingpong<K, G, SYNC><<<148, 256>>>(coef, out, passes, delay);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 / passes * 1.965e9;
}

template <int K, int G, int MODE, int OCC>
double run(int ctas_per_sm, double *coef, double *out) {
  const int passes = 1000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  pass_loop<K, G, MODE, OCC><<<148 * ctas_per_sm, 128>>>(coef, out, passes);
  cudaEventRecord(e0);
  pass_loop<K, G, MODE, OCC><<<148 * ctas_per_sm, 128>>>(coef, out, passes);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 / passes * 1.965e9;     // cycles per pass of all resident warps
}

template <int K, int G, int OCC>
void shape(double *coef, double *out) {
  printf("G=%-2d K=%-2d (%3d slots, FP64 body %3d/warp-pass)\n", G, K, G * K, K * 8);
  printf("   1 warp/scheduler : body %5.0f  +shfl %5.0f  +div,vote %5.0f  div,vote only %5.0f\n",
         run<K, G, 0, OCC>(1, coef, out), run<K, G, 1, OCC>(1, coef, out), run<K, G, 2, OCC>(1, coef, out),
         run<K, G, 3, OCC>(1, coef, out));
  printf("   %d warps/scheduler: body %5.0f  +shfl %5.0f  +div,vote %5.0f  div,vote only %5.0f   -> %5.1f cycles per site-pass\n", OCC,
         run<K, G, 0, OCC>(OCC, coef, out), run<K, G, 1, OCC>(OCC, coef, out), run<K, G, 2, OCC>(OCC, coef, out),
         run<K, G, 3, OCC>(OCC, coef, out), run<K, G, 2, OCC>(OCC, coef, out) / (OCC * (32 / G)));
}

int main() {
  double *coef, *out;
  const size_t n_coef = (size_t) 592 * 128 * 13 * 6;
  cudaMalloc(&coef, n_coef * 8); cudaMalloc(&out, 8);
  double *h = (double *) malloc(n_coef * 8);
  for (size_t i = 0; i < n_coef; i++) h[i] = 0.05 + 0.9 * ((i * 2654435761u) % 1000) / 1000.0;
  cudaMemcpy(coef, h, n_coef * 8, cudaMemcpyHostToDevice);
  free(h);
  printf("cycles per pass (wall clock x 1965 MHz); the FP64 pipe issues one warp instruction per 2 cycles\n");
  printf("G=8 K=13, 256-thread CTA, cycles per pass of both warps of a scheduler:\n   lockstep %5.0f   staggered start (300 / 450 cycles) %5.0f %5.0f   ping-pong after body %5.0f   after reciprocals %5.0f\n",
         run_pp<13, 8, 0>(coef, out, 0), run_pp<13, 8, 1>(coef, out, 300), run_pp<13, 8, 1>(coef, out, 450),
         run_pp<13, 8, 2>(coef, out, 0), run_pp<13, 8, 3>(coef, out, 0));
  shape<13, 8, 2>(coef, out);
  shape<7, 16, 3>(coef, out);
  shape<7, 16, 4>(coef, out);
  shape<4, 32, 4>(coef, out);
  shape<4, 32, 5>(coef, out);
  shape<8, 8, 3>(coef, out);
  return 0;
}

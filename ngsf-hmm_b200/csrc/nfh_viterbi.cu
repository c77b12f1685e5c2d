// nfh_viterbi.cu - most probable IBD path.
//
// Replaces viterbi() (shared/HMM.cpp:98-125) for all individuals.  Two quirks
// of the reference are kept because they change the decoded tracts
// (SURVEY.md finding 3):
//   * the score of state 0 is overwritten before state 1 of the same site is
//     evaluated, so state 1 competes against the UPDATED state-0 score;
//   * comparisons are strict (vmax < pval), so ties keep the k = 0 predecessor,
//     and the final state is the first maximum (array_max_pos).
//
// The reference works in log space.  Here scores are kept in LINEAR space,
// (max, x) instead of (max, +), rescaled by exact powers of two, which makes
// the same decisions except where two candidates agree to ~1e-16 relative -
// far inside the <1e-9 tie band the parity contract excludes.  No log/exp of
// scores is needed, only c = exp(-alpha d) per site.
//
// v1: one thread per individual walks its sites (forward, then traceback).
#include "nfh_device.cuh"
#include "nfh_kernels.h"

namespace nfh {

__global__ void __launch_bounds__(64)
viterbi_sequential(ViterbiArgs A) {
  const uint64_t row = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.n_rows_valid) return;
  const double F = A.indF[row], al = A.alpha[row];
  const double q0 = 1.0 - F, q1 = F;
  const uint64_t n_pad = ((A.n_sites + 7) / 8) * 8;
  unsigned char *work = A.work + (size_t) row * n_pad;

  double v0 = q0, v1 = q1;   // Vi_prob, linear
  for (uint64_t s8 = 0; s8 < A.n_sites; s8 += 8) {
    unsigned long long packed = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint64_t s = s8 + j;
      if (s < A.n_sites) {
        const size_t at = blocked_index(row, s, A.n_rows, A.site_block);
        const double c = exp(-al * A.dist[s]);
        const double e0 = A.e0[at];
        const double e1 = e0 * A.emis[at];
        const double g0 = (1.0 - c) * q0, g1 = (1.0 - c) * q1;
        // l = 0: candidates from k = 0 (stay) and k = 1
        double from0 = v0 * (g0 + c), from1 = v1 * g0;
        unsigned bp0 = from1 > from0;
        const double n0 = (bp0 ? from1 : from0) * e0;
        // l = 1: k = 0 uses the score just written for state 0 (in-place quirk)
        from0 = n0 * g1; from1 = v1 * (g1 + c);
        unsigned bp1 = from1 > from0;
        const double n1 = (bp1 ? from1 : from0) * e1;
        v0 = n0; v1 = n1;
        renorm2(v0, v1);
        packed |= (unsigned long long) (bp0 | (bp1 << 1)) << (8 * j);
      }
    }
    *reinterpret_cast<unsigned long long *>(work + s8) = packed;
  }

  // traceback: path[S-1] = first max; path[s-1] = back[s][path[s]]
  unsigned state = v1 > v0 ? 1u : 0u;
  for (uint64_t s8 = n_pad; s8 >= 8; s8 -= 8) {
    unsigned long long packed = *reinterpret_cast<unsigned long long *>(work + s8 - 8);
    unsigned long long out = 0;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
      if (s8 - 8 + j < A.n_sites) {
        out |= (unsigned long long) state << (8 * j);
        const unsigned bits = (unsigned) (packed >> (8 * j)) & 3u;
        state = (bits >> state) & 1u;
      }
    }
    *reinterpret_cast<unsigned long long *>(work + s8 - 8) = out;
  }
}

void launch_viterbi(const ViterbiArgs &a, cudaStream_t st) {
  viterbi_sequential<<<(unsigned) ((a.n_rows_valid + 63) / 64), 64, 0, st>>>(a);
}

}  // namespace nfh

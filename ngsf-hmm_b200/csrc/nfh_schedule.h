// nfh_schedule.h - work order of the single-launch E-step (estep_fused, nfh_estep.cu).
//
// The forward-backward posterior of one individual needs the ratio plane of that individual twice:
// once for the per-tile products (P items) and once, after the carries of all its tiles are known, for
// the posteriors (A items).  Individuals are taken in WAVES of `wave_rows` individuals and the items of
// consecutive waves are interleaved:
//
//     P(0) | P(1)[:L]  A(0)[0] P(1)[L] A(0)[1] P(1)[L+1] ...  | P(2)[:L]  A(1)[0] P(2)[L] ... | ... | A(W-1)
//
//   * P items are bound by the FP64 pipe, A items by HBM: CTAs that sit on one SM at the same time work on
//     both kinds, so both pipes are busy (one kernel per phase leaves one of them idle);
//   * when the ratio planes of a wave fit the L2 together (short sequences), the A items of the wave find
//     them there: A walks the wave in the opposite order of P (the most recently read tile first), so what
//     the L2 still holds is what is asked for next;
//   * L = `lookahead` P items of the next wave go first so that the CTAs have work while the last
//     individuals of the wave finish their carries.
//
// An A item waits for a flag that is raised by whichever CTA completes the last P item of the same
// individual.  All P items of wave w precede all A items of wave w in ticket order, tickets are taken in
// increasing order by running CTAs, and P items never wait: the order cannot deadlock however the CTAs are
// scheduled.
//
// Plain C++ (no CUDA types) so that host code and tests can include it.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define NFH_HD __host__ __device__ __forceinline__
#else
#define NFH_HD inline
#endif

namespace nfh {

// All counts are 32-bit: the kernel decodes a ticket with 32-bit divisions (64-bit ones cost hundreds of
// instructions per item).  make_schedule() reports total = 0 when 2 * n_rows * n_tiles does not fit.
struct EstepSchedule {
  uint32_t n_rows;       // individuals of this rank
  uint32_t n_tiles;      // tiles per individual
  uint32_t wave_rows;    // individuals per wave (>= 1)
  uint32_t lookahead;    // P items of the next wave placed before the first A item of the current one
  uint32_t n_waves;
  uint32_t wave_items;   // wave_rows * n_tiles
  uint32_t last_rows;    // individuals of the last wave
  uint32_t total;        // 2 * n_rows * n_tiles; 0: does not fit 32 bits (use the three-launch path)
};

struct EstepItem {
  uint32_t apply;        // 0: products (P), 1: posteriors (A)
  uint32_t row, tile;
};

NFH_HD EstepSchedule make_schedule(uint32_t n_rows, uint32_t n_tiles, uint32_t wave_rows, uint32_t lookahead) {
  EstepSchedule s;
  s.n_rows = n_rows; s.n_tiles = n_tiles;
  s.wave_rows = wave_rows < 1 ? 1 : wave_rows;
  if (n_rows && s.wave_rows > n_rows) s.wave_rows = n_rows;
  s.n_waves = n_rows ? (n_rows + s.wave_rows - 1) / s.wave_rows : 0;
  s.wave_items = s.wave_rows * n_tiles;
  s.last_rows = n_rows ? n_rows - (s.n_waves - 1) * s.wave_rows : 0;
  s.lookahead = lookahead;
  const uint64_t total = 2ull * n_rows * n_tiles;
  s.total = total < 0xfff00000ull ? (uint32_t) total : 0;
  return s;
}

NFH_HD uint32_t wave_rows_of(const EstepSchedule &s, uint32_t w) {
  return w + 1 < s.n_waves ? s.wave_rows : (w + 1 == s.n_waves ? s.last_rows : 0);
}

// item `o` of wave w in P order: individuals vary fastest (the CTAs resident at one time share a few
// distance tiles), tiles ascend; A order is the exact reverse
NFH_HD EstepItem wave_item(const EstepSchedule &s, uint32_t w, uint32_t o, uint32_t apply) {
  const uint32_t rows = wave_rows_of(s, w);
  if (apply) o = rows * s.n_tiles - 1 - o;
  EstepItem it;
  it.apply = apply;
  it.tile = o / rows;
  it.row = w * s.wave_rows + (o - it.tile * rows);
  return it;
}

NFH_HD EstepItem decode_ticket(const EstepSchedule &s, uint32_t t) {
  const uint32_t n0 = wave_rows_of(s, 0) * s.n_tiles;
  if (t < n0) return wave_item(s, 0, t, 0);
  t -= n0;
  // block w = items of A(w) and P(w+1), size n_w + n_{w+1}; every block before the last two has 2 * wave_items
  uint32_t w = t / (2 * s.wave_items);
  if (w + 2 > s.n_waves) w = s.n_waves >= 2 ? s.n_waves - 2 : 0;
  t -= w * 2 * s.wave_items;
  uint32_t nw = wave_rows_of(s, w) * s.n_tiles, nn = wave_rows_of(s, w + 1) * s.n_tiles;   // n_{w+1} <= n_w
  if (t >= nw + nn) {                                 // only the block before the last can be shorter
    t -= nw + nn;
    w++;
    nw = nn; nn = 0;
  }
  const uint32_t la = s.lookahead < nn ? s.lookahead : nn;
  if (t < la) return wave_item(s, w + 1, t, 0);
  t -= la;
  const uint32_t pairs = nn - la;                     // A(w)[i], P(w+1)[la + i] alternate for i < pairs
  if (t < 2 * pairs) return (t & 1) ? wave_item(s, w + 1, la + (t >> 1), 0) : wave_item(s, w, t >> 1, 1);
  return wave_item(s, w, t - pairs, 1);
}

}  // namespace nfh

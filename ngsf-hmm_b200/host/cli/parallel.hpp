// parallel.hpp - host threads for the embarrassingly parallel parts of the drop-in binary (normalising the
// genotype likelihoods while reading, formatting the posterior text while writing).  The reference spends its
// --n_threads on the recursions; here those run on the GPU and the host threads serve the file formats.
#pragma once

#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace nfh_cli {

// fn(lo, hi, thread_index) over [0, n) split into contiguous ranges
template <class Fn>
void parallel_for(uint64_t n, unsigned threads, Fn fn) {
  threads = (unsigned) std::max<uint64_t>(1, std::min<uint64_t>(threads, n));
  if (threads == 1) { fn((uint64_t) 0, n, 0u); return; }
  std::vector<std::thread> pool;
  const uint64_t per = (n + threads - 1) / threads;
  for (unsigned t = 0; t < threads; t++) {
    const uint64_t lo = std::min<uint64_t>(n, (uint64_t) t * per), hi = std::min<uint64_t>(n, lo + per);
    if (lo < hi) pool.emplace_back([=]() { fn(lo, hi, t); });
  }
  for (auto &th : pool) th.join();
}

}  // namespace nfh_cli

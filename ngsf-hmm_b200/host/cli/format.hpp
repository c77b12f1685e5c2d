// format.hpp - fast fixed-point text for the probabilities the output files are full of.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>

namespace nfh_cli {

// printf("%f") of a value in [0, 1] without printf: 8 characters "d.dddddd".  The digits are those of the
// value rounded to 6 decimals exactly as glibc rounds the exact binary value (ties to even).  Anything else -
// negative, > 1, NaN - goes through snprintf.  Returns the number of characters written (no terminator).
inline int format_unit_f(double v, char *out) {
  if (!(v >= 0.0 && v <= 1.0)) return snprintf(out, 32, "%f", v);
  // exact product v * 1e6 = p + e (p rounded, e its error by one FMA); n = floor(p), and the side of the
  // half-way point is decided by p - n - 1/2 (exact), then by e, then - on an exact tie - by evenness
  const double p = v * 1e6;
  const double e = fma(v, 1e6, -p);
  uint32_t n = (uint32_t) p;
  const double d = (p - (double) n) - 0.5;
  if (d > 0.0 || (d == 0.0 && (e > 0.0 || (e == 0.0 && (n & 1u))))) n++;
  out[0] = (char) ('0' + n / 1000000u);
  out[1] = '.';
  uint32_t frac = n % 1000000u;
  for (int k = 7; k >= 2; k--) { out[k] = (char) ('0' + frac % 10u); frac /= 10u; }
  return 8;
}

}  // namespace nfh_cli

// Checks nfh_cli::format_unit_f against printf("%f") (built and run by tests/test_host_logic.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "format.hpp"

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  long bad = 0, n = 0;
  auto check = [&](double v) {
    char a[64], b[64];
    const int la = nfh_cli::format_unit_f(v, a);
    a[la] = '\0';
    snprintf(b, sizeof b, "%f", v);
    n++;
    if (strcmp(a, b) != 0) { if (bad < 5) fprintf(stderr, "mismatch %.17g: %s vs %s\n", v, a, b); bad++; }
  };
  const double edge[] = {0.0, 1.0, 1e-5, 1 - 1e-5, 0.5, 4.9999999e-7, 5.0000001e-7, 0.9999995, 0.99999949999, 0.1234565,
                         0.0000005, 0.3333333333, 1e-300, 0.9999999999999999, -0.25, 1.5, 123456.789};
  for (double v : edge) check(v);
  for (int k = 0; k <= 1000000; k += 7) { check(k / 1e6); check((k + 0.5) / 1e6); check(nextafter((k + 0.5) / 1e6, 0.0)); }
  for (long i = 0; i < 3000000; i++) check(uni(rng));
  for (long i = 0; i < 300000; i++) check(uni(rng) * 1e-4);
  printf("%ld values, %ld mismatches\n", n, bad);
  return bad ? 1 : 0;
}

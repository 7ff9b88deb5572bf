// Host check of csrc/glibc_math.cuh against the C library (bit for bit).  Built and run by tests/test_glibc_math.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "glibc_math.cuh"

static unsigned long long bits(double d) { unsigned long long u; memcpy(&u, &d, 8); return u; }

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> u01(0.0, 1.0);
  long bad[4] = {0, 0, 0, 0};
  for (long k = 0; k < n; k++) {
    // exp: the cooling module's -1.3e5/T (T >= 1e4) plus a wide sweep
    double x = (k & 1) ? -1.3e5 / (1e4 * pow(10.0, 6.0 * u01(rng))) : (u01(rng) * 1020.0 - 510.0);
    if (bits(pbm::exp_glibc(x)) != bits(exp(x))) { if (bad[0]++ < 5) printf("exp(%a): %a vs %a\n", x, pbm::exp_glibc(x), exp(x)); }
    // log: wide range and the neighbourhood of 1
    double y = (k & 1) ? pow(10.0, 24.0 * u01(rng) - 12.0) : 0.93 + 0.14 * u01(rng);
    if (bits(pbm::log_glibc(y)) != bits(log(y))) { if (bad[1]++ < 5) printf("log(%a): %a vs %a\n", y, pbm::log_glibc(y), log(y)); }
    // log10 of temperatures
    double T = pow(10.0, 3.0 + 7.0 * u01(rng));
    if (k % 3 == 0) T = ldexp(1.0 + 0.07 * (u01(rng) - 0.5), 10 + (int)(20 * u01(rng)));     // mantissa near 1
    if (bits(pbm::log10_glibc(T)) != bits(log10(T))) { if (bad[2]++ < 5) printf("log10(%a): %a vs %a\n", T, pbm::log10_glibc(T), log10(T)); }
    // pow: pow(10, y) of ne_rat() and pow(xi, 0.25), plus general pairs
    double a, b;
    if (k % 3 == 0) { a = 10.0; b = -51.59417133 + 12.27740153 * log10(T); }
    else if (k % 3 == 1) { a = pow(10.0, 30.0 * u01(rng) - 10.0); b = 0.25; }
    else { a = pow(10.0, 20.0 * u01(rng) - 10.0); b = 8.0 * u01(rng) - 4.0; }
    if (bits(pbm::pow_glibc(a, b)) != bits(pow(a, b))) { if (bad[3]++ < 5) printf("pow(%a,%a): %a vs %a\n", a, b, pbm::pow_glibc(a, b), pow(a, b)); }
  }
  printf("mismatches in %ld arguments each: exp %ld  log %ld  log10 %ld  pow %ld\n", n, bad[0], bad[1], bad[2], bad[3]);
  return (bad[0] | bad[1] | bad[2] | bad[3]) ? 1 : 0;
}

// Host harness for ndtpso_slam_b200/csrc/fast_exp.h: max ulp error of fast_exp vs libm exp
// (glibc exp is correctly rounded to < 1 ulp).  Built and run by tests/test_fast_exp.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../ndtpso_slam_b200/csrc/fast_exp.h"

static double ulp_of(double x) {
  if (x == 0) return 4.9406564584124654e-324;
  int e;
  frexp(x, &e);
  return ldexp(1.0, e - 53);
}

int main(int argc, char** argv) {
  double table[ndtpso::kExpTableSize];
  for (int j = 0; j < ndtpso::kExpTableSize; ++j) table[j] = exp2((double)j / ndtpso::kExpTableSize);
  const long n = argc > 1 ? atol(argv[1]) : 2000000;
  unsigned long long s = 88172645463325252ull;
  double worst = 0, worst_a = 0;
  long flushed = 0;
  for (long i = 0; i < n + 4000; ++i) {
    double a;
    if (i < n) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const double u = (double)(s >> 11) / 9007199254740992.0;
      const int band = i % 4;
      a = band == 0 ? -u * 1.0 : band == 1 ? -u * 40.0 : band == 2 ? -u * 708.0 : (u - 0.5) * 20.0;
    } else {
      a = -((double)(i - n)) * 0.25;  // -0, -0.25, ... -1000: exact multiples incl. far below the flush point
    }
    const double got = ndtpso::fast_exp(a, table);
    const double want = exp(a);
    if (a < -708.0) {
      if (got != 0.0) { printf("FAIL flush a=%.17g got=%g\n", a, got); return 1; }
      ++flushed;
      continue;
    }
    const double err = fabs(got - want) / ulp_of(want);
    if (err > worst) { worst = err; worst_a = a; }
  }
  // specials
  if (ndtpso::fast_exp(0.0, table) != 1.0) { printf("FAIL exp(0)\n"); return 1; }
  if (ndtpso::fast_exp(-1e300, table) != 0.0) { printf("FAIL exp(-1e300)\n"); return 1; }
  if (ndtpso::fast_exp(-INFINITY, table) != 0.0) { printf("FAIL exp(-inf)\n"); return 1; }
  if (!std::isinf(ndtpso::fast_exp(800.0, table))) { printf("FAIL exp(800)\n"); return 1; }
  if (!std::isnan(ndtpso::fast_exp(NAN, table))) { printf("FAIL exp(nan)\n"); return 1; }
  const double e709 = ndtpso::fast_exp(708.9, table);
  if (fabs(e709 - exp(708.9)) / ulp_of(exp(708.9)) > 1.5) { printf("FAIL exp(708.9)\n"); return 1; }
  printf("max_ulp_err %.4f at a=%.17g flushed %ld\n", worst, worst_a, flushed);
  return worst <= 1.0 ? 0 : 2;
}

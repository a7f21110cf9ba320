// exp(a) for the NDT score: 12 fp64-pipe instructions and one conflict-free 8-byte table load.
//
//   k  = round(a * 16/ln2)                (magic-number add; k = 16 m + j, 0 <= j < 16)
//   r  = a - k * ln2/16                   (two-term Cody-Waite, |r| <= ln2/32 = 2.17e-2)
//   e^a = 2^m * T[j] * (1 + r + r^2/2! + ... + r^7/7!),   T[j] = 2^(j/16)
//
// The table is 16 doubles = 128 bytes = exactly one row of the 32 shared-memory banks, so any
// pattern of j across a warp is conflict-free (a 64-entry table cost ~5 wavefronts per lookup).
// Truncation error r^8/8! <= 1.2e-18; measured max error vs glibc exp: 1 ulp
// (tests/test_fast_exp.py).  Results below the smallest normal double (a < -708) are flushed to
// zero: an absolute error <= 2.3e-308 per scan point, against scores of order 1..1000.
// a > 709 (only possible with an indefinite "inverse covariance") returns +inf.
//
// The same source is compiled for the device (ndtpso_kernels.cuh) and for the host test harness.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NDTPSO_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define NDTPSO_HD static inline
#endif

namespace ndtpso {

constexpr int kExpTableSize = 16;
constexpr int kExpTableShift = 4;                 // log2(kExpTableSize)
constexpr double kExpMagic = 6755399441055744.0;  // 1.5 * 2^52

struct ExpConsts {
  double l2e, hi, lo;       // 16/ln2 and ln2/16 = hi + lo
  double c7, c6, c5, c4, c3;  // 1/7! .. 1/3!
};

// ln2/16 = hi + lo with the low 21 mantissa bits of hi zero, so k*hi is exact for |k| < 2^21.
NDTPSO_HD ExpConsts exp_consts() {
  ExpConsts c;
  c.l2e = 23.083120654223414;      // 16/ln2
  c.hi = 0.04332169877307024;      // 0x3FA62E42FEE00000
  c.lo = 1.1926343307941173e-11;   // ln2/16 - hi
  c.c7 = 0.0001984126984126984;
  c.c6 = 0.001388888888888889;
  c.c5 = 0.008333333333333333;
  c.c4 = 0.041666666666666664;
  c.c3 = 0.16666666666666666;
  return c;
}

NDTPSO_HD int32_t dbl_hi(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return (int32_t)(b >> 32);
#endif
}
NDTPSO_HD int32_t dbl_lo(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return (int32_t)(b & 0xffffffffu);
#endif
}
NDTPSO_HD double dbl_make(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}

// table: T[j] = 2^(j/16), correctly rounded (exp_table.inc)
NDTPSO_HD double fast_exp(double a, const double* __restrict__ table) {
  const ExpConsts c = exp_consts();
  const double kd = fma(a, c.l2e, kExpMagic);
  const int32_t k = dbl_lo(kd);
  const double kf = kd - kExpMagic;
  double r = fma(kf, -c.hi, a);
  r = fma(kf, -c.lo, r);
  double p = fma(r, c.c7, c.c6);
  p = fma(p, r, c.c5);
  p = fma(p, r, c.c4);
  p = fma(p, r, c.c3);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  const double em1 = p * r;
  const double t = table[k & (kExpTableSize - 1)];
  const double v = fma(t, em1, t);
  const int32_t m = k >> kExpTableShift;
  double res = dbl_make(dbl_hi(v) + (m << 20), dbl_lo(v));
  const int32_t ahi = dbl_hi(a);
  // a < -708.0  (hi word of -708.0 is 0xC0862000; negative doubles order by magnitude as unsigned)
  if ((uint32_t)ahi > 0xC0862000u) res = 0.0;
  // a > 709.0 (hi word 0x40862800) or +inf/NaN: overflow
  if (ahi > 0x40862800) res = a + dbl_make(0x7FF00000, 0);
  return res;
}

}  // namespace ndtpso

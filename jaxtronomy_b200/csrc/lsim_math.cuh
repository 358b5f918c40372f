// lsim_math.cuh -- fp64 scalar helpers shared by every kernel (and by the host-compiled math test harness).
//
// Everything here is `LSIM_HD` (host + device) so that tests/hostmath can compile the very same profile
// arithmetic with g++ and check it against the oracle on a machine without a GPU.  The device build is
// the product; the host build exists only under tests/.
//
// Design rule (measured, profiles/r1_k_image_*.md): the fp64 pipe (64 DFMA lanes / SM / clk on B200) is the
// binding unit of the image kernel, so
//   * transcendental cores are short FMA chains with no special-case handling in the fast path,
//   * anything that can run on another pipe does: reciprocal seeds, int<->double conversions and rounding go
//     to the XU pipe (MUFU.RCP64H, I2F, F2I, FRND), exponent manipulation to the integer pipe,
//   * polynomial coefficients live in constant memory so that DFMA reads them as c[bank][imm] operands
//     instead of being re-materialised into registers.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LSIM_HD __host__ __device__ __forceinline__
#define LSIM_D __device__ __forceinline__
#define LSIM_TABLE __constant__ double
#else
#define LSIM_HD inline
#define LSIM_D inline
#define LSIM_TABLE static const double
#endif

namespace lsim {

constexpr double kPi = 3.14159265358979323846;
constexpr double kLog2e = 1.44269504088896340736;
constexpr double kLn2 = 0.69314718055994530942;

// atanh(s)/s = 1 + z/3 + z^2/5 + ...  (z = s^2): coefficients 1/(2k+1), k = 1..10, highest first
LSIM_TABLE kAtanhC[10] = {1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0,
                          1.0 / 11.0, 1.0 / 9.0,  1.0 / 7.0,  1.0 / 5.0,  1.0 / 3.0};
// exp(r) Taylor coefficients 1/k!, k = 13..2, highest first
LSIM_TABLE kExpC[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                        1.0 / 362880.0,     1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,
                        1.0 / 120.0,        1.0 / 24.0,        1.0 / 6.0,        0.5};

LSIM_HD int64_t d2i(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  int64_t i;
  memcpy(&i, &x, 8);
  return i;
#endif
}
LSIM_HD double i2d(int64_t i) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(i);
#else
  double x;
  memcpy(&x, &i, 8);
  return x;
#endif
}

// jnp.nan_to_num(v, posinf=p, neginf=n): nan -> 0, +inf -> p, -inf -> n
LSIM_HD double nan_to_num(double v, double p, double n) {
  if (v != v) return 0.0;
  if (v > DBL_MAX) return p;
  if (v < -DBL_MAX) return n;
  return v;
}
LSIM_HD double nan_to_num(double v) { return nan_to_num(v, DBL_MAX, -DBL_MAX); }

// 1/x for finite, normal, non-zero x (no special cases): MUFU.RCP64H seed (XU pipe) + two Newton steps.
// x = 0 gives +-inf / nan exactly like a true division would for the callers' purposes (nan_to_num follows).
LSIM_HD double rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

// ---- fast fp64 log2 / exp2 --------------------------------------------------------------------------
// The hot loop needs pow(x, s) = exp2(s * log2(x)) and exp(-b (k - 1)) for finite positive normal x.
// libdevice's log/exp/pow carry special-case handling (denormals, inf/nan, exact rounding paths) that
// costs more fp64-pipe instructions than the polynomial core.  These versions keep the core only:
//   log2_pos(x): x positive, finite, normal.  |rel err| < 3e-16 on the result as a real number
//                (abs err < 2e-16 near x = 1).
//   exp2_fin(x): any x; underflows to 0 below -1022, +inf above 1024, nan -> nan.

// log2(x) = e + log2(m), m in [sqrt(1/2), sqrt(2)); log(m) = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716
LSIM_HD double log2_pos(double x) {
  int64_t ix = d2i(x);
  const int64_t kSqrtHalfBits = 0x3fe6a09e667f3bcdLL;  // bits of sqrt(0.5)
  int64_t e = (ix - kSqrtHalfBits) >> 52;              // exponent relative to sqrt(1/2)
  double m = i2d(ix - (e << 52));                      // m in [sqrt(1/2), sqrt(2))
  double num = m - 1.0, den = m + 1.0;
  double r = rcp_fast(den);
  double s = num * r;
  s = fma(fma(-s, den, num), r, s);  // one correction step: s = num/den to < 1 ulp
  double z = s * s;
  double p = kAtanhC[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < 10; ++k) p = fma(p, z, kAtanhC[k]);
  // log2(m) = (2/ln2) * s * (1 + z p)
  const double c = 2.0 * kLog2e;
  double lm = fma(c, s * z * p, c * s);
  return (double)(int)e + lm;
}

// exp2(x) = 2^n * 2^r, r = x - n in [-1/2, 1/2]; 2^r = exp(r ln2) by a degree-13 Taylor polynomial
LSIM_HD double exp2_fin(double x) {
  if (!(fabs(x) < 1021.0)) return exp2(x);  // rare: underflow / overflow / nan -> the library handles them
  double n = rint(x);
  int in = (int)n;
  double r = (x - n) * kLn2;  // |r| <= 0.3466
  double p = kExpC[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < 12; ++k) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return i2d(d2i(p) + ((int64_t)in << 52));  // scale by 2^n: p in [0.7, 1.42], |n| <= 1021 -> stays normal
}

}  // namespace lsim

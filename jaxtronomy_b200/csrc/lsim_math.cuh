// lsim_math.cuh -- fp64 scalar helpers shared by every kernel (and by the host-compiled math test harness).
//
// Everything here is `LSIM_HD` (host + device) so that tests/hostmath can compile the very same profile
// arithmetic with g++ and check it against the oracle on a machine without a GPU.  The device build is
// the product; the host build exists only under tests/.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LSIM_HD __host__ __device__ __forceinline__
#define LSIM_D __device__ __forceinline__
#else
#define LSIM_HD inline
#define LSIM_D inline
#endif

namespace lsim {

constexpr double kPi = 3.14159265358979323846;
constexpr double kLog2e = 1.44269504088896340736;
constexpr double kLn2 = 0.69314718055994530942;

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

// ---- fast fp64 log2 / exp2 --------------------------------------------------------------------------
// The hot loop needs pow(x, s) = exp2(s * log2(x)) and exp(-b (k - 1)) for finite positive normal x.
// libdevice's log/exp/pow carry special-case handling (denormals, inf/nan, exact rounding paths) that
// costs more fp64-pipe instructions than the polynomial core.  These versions keep the core only:
//   log2_pos(x): x positive, finite, normal.  |rel err| < 3e-16 on the result as a real number
//                (abs err < 2e-16 near x = 1).
//   exp2_fin(x): any finite x; underflows to 0 below -1022, saturates to +inf above 1024.
// Both are pure FMA chains with one division (log2) -- branch-free apart from the range clamps.

// log2(x) = e + log2(m), m in [sqrt(1/2), sqrt(2)); log(m) = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716
LSIM_HD double log2_pos(double x) {
  int64_t ix = d2i(x);
  // shift the mantissa window so m lands in [sqrt(1/2), sqrt(2)): subtract the bit pattern of sqrt(1/2)'s mantissa
  const int64_t kSqrtHalfMant = 0x3fe6a09e667f3bcdLL;  // bits of sqrt(0.5)
  int64_t e = (ix - kSqrtHalfMant) >> 52;              // floor exponent relative to sqrt(1/2)
  double m = i2d(ix - (e << 52));                      // m in [sqrt(1/2), sqrt(2))
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  // atanh(s)/s = 1 + z/3 + z^2/5 + ... ; z <= 0.02944, 10 terms -> truncation < 1e-18 relative
  double p = 1.0 / 21.0;
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  // log2(m) = (2/ln2) * s * (1 + z p)
  const double c = 2.0 * kLog2e;  // 2/ln 2
  double sp = s * z * p;          // small correction
  double lm = fma(c, sp, c * s);  // c*s + c*s*z*p
  return (double)e + lm;
}

// exp2(x) = 2^n * 2^r, r = x - n in [-1/2, 1/2]; 2^r = exp(r ln2) by a degree-13 Taylor polynomial
LSIM_HD double exp2_fin(double x) {
  if (!(x > -1022.0)) return (x != x) ? x : 0.0;
  if (x > 1023.999) return (x < 1024.0) ? DBL_MAX : HUGE_VAL;
  double n = rint(x);
  double r = (x - n) * kLn2;  // |r| <= 0.3466
  double p = 1.0 / 6227020800.0;  // 1/13!
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int64_t in = (int64_t)n;
  return i2d(d2i(p) + (in << 52));  // scale by 2^n (p in [0.7, 1.42], result stays normal for n >= -1021)
}

}  // namespace lsim

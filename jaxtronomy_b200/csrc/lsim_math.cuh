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

#include "lsim_math_tables.h"

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

// log2(1+u)/u = (1/ln2) (1 - u/2 + u^2/3 - ...): coefficients (-1)^(k+1)/(k ln2), k = 10..1, highest first
LSIM_TABLE kLog2C[10] = {-kLog2e / 10.0, kLog2e / 9.0, -kLog2e / 8.0, kLog2e / 7.0, -kLog2e / 6.0,
                         kLog2e / 5.0,   -kLog2e / 4.0, kLog2e / 3.0, -kLog2e / 2.0, kLog2e};
// exp(y) Taylor coefficients 1/k!, k = 6..2, highest first
LSIM_TABLE kExpC[5] = {1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};

// Lookup tables (tools/gen_math_tables.py): exp2 T[j] = 2^(j/32); log2 {R[i], L[i]} = {~1/c_i, -log2 R[i]} for
// the 32 bins of the mantissa.  On the device they live in global memory and are read through the read-only
// L1 path: the index differs from lane to lane, which constant memory would serialise.
#if defined(__CUDACC__)
__device__ double d_exp2_tab[32] = LSIM_EXP2_TABLE;
__device__ double d_log2_tab[64] = LSIM_LOG2_TABLE;
#endif
static const double h_exp2_tab[32] = LSIM_EXP2_TABLE;
static const double h_log2_tab[64] = LSIM_LOG2_TABLE;

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
// libdevice's log/exp/pow carry special-case handling and long polynomials; on a kernel bound by the fp64 pipe
// every DFMA counts, so both functions are table-driven here (32-entry tables, short Taylor polynomials):
//   log2_pos(x): x positive, finite, normal.  12 fp64 instructions.  |abs err| < 3e-16 max(1, |log2 x|)
//                (the result feeds exponents, where absolute accuracy is what matters).
//   exp2_fin(x): any x; 10 fp64 instructions in the fast path; |rel err| < 4e-16; underflow / overflow / nan
//                are delegated to the library.

// x = 2^e m, m in [1,2); bin i = top 5 mantissa bits; u = m R_i - 1, |u| <= 1/64;
// log2 x = e + L_i + log2(1 + u)
LSIM_HD double log2_pos(double x) {
  const int64_t ix = d2i(x);
  const int e = (int)(ix >> 52) - 1023;
  const double m = i2d((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const int i = (int)(ix >> 47) & 31;
#if defined(__CUDA_ARCH__)
  const double2 rl = __ldg(reinterpret_cast<const double2*>(d_log2_tab) + i);
  const double R = rl.x, L = rl.y;
#else
  const double R = h_log2_tab[2 * i], L = h_log2_tab[2 * i + 1];
#endif
  const double u = fma(m, R, -1.0);
  double p = kLog2C[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < 10; ++k) p = fma(p, u, kLog2C[k]);
  return (double)e + fma(u, p, L);
}

// x = k/32 + r, k = rint(32 x), |r| <= 1/64: 2^x = 2^(k>>5) T[k&31] exp(r ln2)
LSIM_HD double exp2_fin(double x) {
  if (!(fabs(x) < 1021.0)) return exp2(x);  // rare: underflow / overflow / nan -> the library handles them
  const double kf = rint(x * 32.0);
  const int ki = (int)kf;
  const double y = fma(kf, -0.03125, x) * kLn2;  // r is exact; |y| <= 0.0109
  double p = kExpC[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < 5; ++k) p = fma(p, y, kExpC[k]);
  p = fma(p, y, 1.0);
  p = fma(p, y, 1.0);
#if defined(__CUDA_ARCH__)
  const double t = __ldg(d_exp2_tab + (ki & 31)) * p;
#else
  const double t = h_exp2_tab[ki & 31] * p;
#endif
  return i2d(d2i(t) + ((int64_t)(ki >> 5) << 52));  // t in [0.98, 2), |k>>5| <= 1021: stays normal
}

}  // namespace lsim

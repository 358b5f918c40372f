// lsim_math.cuh -- fp64 scalar helpers shared by every kernel (and by the host-compiled math test harness).
//
// Everything here is `LSIM_HD` (host + device) so that tests/hostmath can compile the very same profile
// arithmetic with g++ and check it against the oracle on a machine without a GPU.  The device build is
// the product; the host build exists only under tests/.
//
// Design rule (measured, profiles/image_r1*_ncu_summary.txt, profiles/r1_dfma_operand_probe.md): the fp64 pipe (64 DFMA lanes / SM / clk on B200) is the
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

// log2(1+u)/u = (1/ln2) (1 - u/2 + u^2/3 - ...): coefficients (-1)^(k+1)/(k ln2), k = 6..1, highest first
LSIM_TABLE kLog2C[6] = {-kLog2e / 6.0, kLog2e / 5.0, -kLog2e / 4.0, kLog2e / 3.0, -kLog2e / 2.0, kLog2e};
// (2^r - 1)/r = ln2 + ln2^2 r/2 + ...: coefficients ln2^k/k!, k = 4..1, highest first
LSIM_TABLE kExp2C[4] = {kLn2 * kLn2 * kLn2 * kLn2 / 24.0, kLn2 * kLn2 * kLn2 / 6.0, kLn2 * kLn2 / 2.0, kLn2};

// Lookup tables (tools/gen_math_tables.py): exp2 T[j] = 2^(j/256); log2 {R[i], L[i]} = {~1/c_i, -log2 R[i]} for
// the 128 bins of the mantissa (2 KB each).  The index differs from lane to lane, which constant memory would
// serialise, so every kernel that evaluates profiles copies the tables from global into SHARED memory once per CTA
// (stage_math_tables): a lookup is then one LDS at a 32-bit offset (no 64-bit address arithmetic, lower latency
// than an L1 hit).
#if defined(__CUDACC__)
__device__ double d_exp2_tab[LSIM_EXP2_N] = LSIM_EXP2_TABLE;
__device__ double d_log2_tab[2 * LSIM_LOG2_N] = LSIM_LOG2_TABLE;
__shared__ double s_exp2_tab[LSIM_EXP2_N];
__shared__ double2 s_log2_tab[LSIM_LOG2_N];
// call at kernel entry, followed by a __syncthreads() before the first log2_pos / exp2_*
__device__ __forceinline__ void stage_math_tables() {
  for (int i = threadIdx.x; i < LSIM_EXP2_N; i += blockDim.x) s_exp2_tab[i] = d_exp2_tab[i];
  for (int i = threadIdx.x; i < LSIM_LOG2_N; i += blockDim.x)
    s_log2_tab[i] = reinterpret_cast<const double2*>(d_log2_tab)[i];
}
#endif
static const double h_exp2_tab[LSIM_EXP2_N] = LSIM_EXP2_TABLE;
static const double h_log2_tab[2 * LSIM_LOG2_N] = LSIM_LOG2_TABLE;

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

// 1/x for finite, normal, non-zero x (no special cases): MUFU.RCP64H seed (XU pipe, relative error <= 2^-23) and ONE
// third-order step r (1 + e + e^2), e = 1 - x r: 3 DFMA instead of the 4 of two Newton steps, error e^3 ~ 2^-69 before
// the final rounding.  x = 0 gives +-inf / nan exactly like a true division would for the callers' purposes
// (nan_to_num follows).
LSIM_HD double rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
  return 1.0 / x;
#endif
}

// max(x, lo) / min(x, hi) that PROPAGATE nan like jnp.maximum / jnp.minimum (fmax / fmin drop it and cost a
// dozen instructions for that): one compare and a select.
LSIM_HD double max_nan(double x, double lo) { return (x < lo) ? lo : x; }
LSIM_HD double min_nan(double x, double hi) { return (x > hi) ? hi : x; }

// sqrt(x) for finite, normal, positive x (no special cases): MUFU.RSQ64H seed (XU pipe), two Newton steps on
// 1/sqrt(x) and one correction of the root; < 1 ulp.  x = 0 / inf / nan are NOT handled (callers guarantee x > 0).
LSIM_HD double sqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  const double s = x * y;
  return fma(0.5 * y, fma(-s, s, x), s);
#else
  return sqrt(x);
#endif
}

// 1/sqrt(x) for finite, normal, positive x: MUFU.RSQ64H seed + two Newton steps (~1 ulp).
LSIM_HD double rsqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

// ---- fast fp64 log2 / exp2 --------------------------------------------------------------------------
// The hot loop needs pow(x, s) = exp2(s * log2(x)) and exp(-b (k - 1)) for finite positive normal x.
// libdevice's log/exp/pow carry special-case branches and long polynomials; on a kernel bound by the fp64 pipe
// every DFMA counts and every branch cuts the instruction stream into blocks the scheduler cannot interleave,
// so both functions are table-driven and BRANCH-FREE here:
//   log2_pos(x): x positive, finite, normal.  8 fp64 instructions.  |abs err| < 3e-16 max(1, |log2 x|)
//                (the result feeds exponents, where absolute accuracy is what matters).
//   exp2_fin(x): any x; 7 fp64 instructions; |rel err| < 4e-16.  The argument is clamped to [-1022, 1022] with
//                integer instructions, i.e. results saturate at 2^-1022 (2.2e-308, instead of denormals / 0) and
//                2^1022 (instead of +inf); nan propagates.
//   exp2_neg(x): the same for arguments known to stay below 1022 (only the lower clamp).

// x = 2^e m, m in [1,2); bin i = top 7 mantissa bits; u = m R_i - 1, |u| <= 1/256;
// log2 x = e + L_i + log2(1 + u)
LSIM_HD double log2_pos(double x) {
  const int64_t ix = d2i(x);
  const int e = (int)(ix >> 52) - 1023;
  const double m = i2d((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const int i = (int)(ix >> 45) & (LSIM_LOG2_N - 1);
#if defined(__CUDA_ARCH__)
  const double2 rl = s_log2_tab[i];
  const double R = rl.x, L = rl.y;
#else
  const double R = h_log2_tab[2 * i], L = h_log2_tab[2 * i + 1];
#endif
  const double u = fma(m, R, -1.0);
  double p = fma(kLog2C[0], u, kLog2C[1]);
  p = fma(p, u, kLog2C[2]);
  p = fma(p, u, kLog2C[3]);
  p = fma(p, u, kLog2C[4]);
  p = fma(p, u, kLog2C[5]);
  return (double)e + fma(u, p, L);
}

// x = k/256 + r, k = rint(256 x), |r| <= 1/512: 2^x = 2^(k>>8) T[k&255] (1 + r P(r)).  |x| <= 1022 or nan.
LSIM_HD double exp2_core(double x) {
#if !defined(__CUDA_ARCH__)
  if (x != x) return x;  // (int)nan is undefined on the host; the device conversion gives 0 and nan flows through r
#endif
  const double kf = rint(x * 256.0);
  const int ki = (int)kf;
  const double r = fma(kf, -0.00390625, x);  // exact
  double p = fma(kExp2C[0], r, kExp2C[1]);
  p = fma(p, r, kExp2C[2]);
  p = fma(p, r, kExp2C[3]);
#if defined(__CUDA_ARCH__)
  const double T = s_exp2_tab[ki & (LSIM_EXP2_N - 1)];
#else
  const double T = h_exp2_tab[ki & (LSIM_EXP2_N - 1)];
#endif
  const double t = fma(T * r, p, T);  // in [0.9986, 1.9974]
  // scale by 2^(k>>8): |k>>8| <= 1022 and (k>>8 = -1022 => t >= 1), so the result stays normal and the exponent
  // can be added to the high word alone
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(t) + ((ki >> 8) << 20), __double2loint(t));
#else
  return i2d(d2i(t) + ((int64_t)(ki >> 8) << 52));
#endif
}

// 1022.0 = 0x408FF000_00000000: the clamps compare the high word only (exact, the low word of the bound is zero)
LSIM_HD double exp2_fin(double x) {
  const int64_t ix = d2i(x);
  const uint32_t hx = (uint32_t)((uint64_t)ix >> 32), ax = hx & 0x7fffffffu;
  const bool big = (uint32_t)(ax - 0x408FF000u) <= (0x7ff00000u - 0x408FF000u);  // 1022 <= |x| <= inf, not nan
  const double xc = big ? i2d((int64_t)(((uint64_t)((hx & 0x80000000u) | 0x408FF000u)) << 32)) : x;
  return exp2_core(xc);
}

LSIM_HD double exp2_neg(double x) {
  const uint32_t hx = (uint32_t)((uint64_t)d2i(x) >> 32);
  const bool low = (uint32_t)(hx - 0xC08FF000u) <= (0xFFF00000u - 0xC08FF000u);  // -inf <= x <= -1022, not nan
  return exp2_core(low ? -1022.0 : x);
}

// ---- exp2 of a PRE-SCALED argument: y = 256 x -------------------------------------------------------------------
// The hot profiles compute their exponents with one FMA from prepared constants; folding the factor 256 of the table
// index into those constants removes the multiplication at the head of exp2_core, and the reduced argument becomes a
// subtraction: k = rint(y), r' = y - k (exact), 2^x = 2^(k>>8) T[k&255] (1 + r' P'(r')) with the coefficients of P'
// scaled by powers of 256 (exact), so the result is bit-identical to exp2_core(y / 256).  6 fp64 instructions.
LSIM_TABLE kExp2S[4] = {kLn2 * kLn2 * kLn2 * kLn2 / 24.0 / 4294967296.0, kLn2 * kLn2 * kLn2 / 6.0 / 16777216.0,
                        kLn2 * kLn2 / 2.0 / 65536.0, kLn2 / 256.0};

LSIM_HD double exp2_core_s(double y) {
#if !defined(__CUDA_ARCH__)
  if (y != y) return y;
#endif
  const double kf = rint(y);
  const int ki = (int)kf;
  const double r = y - kf;  // exact: |r| <= 1/2 and both are multiples of ulp(y)
  double p = fma(kExp2S[0], r, kExp2S[1]);
  p = fma(p, r, kExp2S[2]);
  p = fma(p, r, kExp2S[3]);
#if defined(__CUDA_ARCH__)
  const double T = s_exp2_tab[ki & (LSIM_EXP2_N - 1)];
#else
  const double T = h_exp2_tab[ki & (LSIM_EXP2_N - 1)];
#endif
  const double t = fma(T * r, p, T);
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(t) + ((ki >> 8) << 20), __double2loint(t));
#else
  return i2d(d2i(t) + ((int64_t)(ki >> 8) << 52));
#endif
}

// 256 * 1022 = 261632.0 = 0x410FF000_00000000 (same mantissa as 1022: the clamps still compare the high word only)
LSIM_HD double exp2_fin_s(double y) {
  const int64_t iy = d2i(y);
  const uint32_t hy = (uint32_t)((uint64_t)iy >> 32), ay = hy & 0x7fffffffu;
  const bool big = (uint32_t)(ay - 0x410FF000u) <= (0x7ff00000u - 0x410FF000u);  // 1022 <= |y|/256 <= inf, not nan
  const double yc = big ? i2d((int64_t)(((uint64_t)((hy & 0x80000000u) | 0x410FF000u)) << 32)) : y;
  return exp2_core_s(yc);
}

LSIM_HD double exp2_neg_s(double y) {
  const uint32_t hy = (uint32_t)((uint64_t)d2i(y) >> 32);
  const bool low = (uint32_t)(hy - 0xC10FF000u) <= (0xFFF00000u - 0xC10FF000u);  // -inf <= y/256 <= -1022, not nan
  return exp2_core_s(low ? -261632.0 : y);
}

}  // namespace lsim

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
#define LSIM_TABLE static __constant__ double
#else
#define LSIM_HD inline
#define LSIM_D inline
#define LSIM_TABLE static const double
#endif

namespace lsim {

constexpr double kPi = 3.14159265358979323846;
constexpr double kLog2e = 1.44269504088896340736;
constexpr double kLn2 = 0.69314718055994530942;

// exp2_*_s take their argument pre-scaled by the exp2 table size
constexpr double kExp2Scale = (double)LSIM_EXP2_N;

// log2(1+u)/u = (1/ln2) (1 - u/2 + u^2/3 - ...): coefficients (-1)^(k+1)/(k ln2), k = 4..1, highest first
// (|u| <= 2^-11 with 1024 bins: the first dropped term is u^5/(5 ln2) <= 8e-18)
LSIM_TABLE kLog2C[4] = {-kLog2e / 4.0, kLog2e / 3.0, -kLog2e / 2.0, kLog2e};

// Lookup tables (tools/gen_math_tables.py): exp2 T[j] = 2^(j/2048) (stored as bit patterns with j << 9 subtracted
// from the high word, see exp2_core_s); log2 {R[i], L[i]} = {~1/c_i, -log2 R[i]} for the 1024 bins of the mantissa
// (16 KB each: round 2 traded table size for polynomial length -- two multiply-adds less per log2, one less per
// exp2, 7 of the ~107 fp64 instructions of a headline sub-pixel).  The index differs from lane to lane, which constant memory would serialise, so every kernel that
// evaluates profiles copies the tables from global into SHARED memory once per CTA (stage_math_tables): a lookup is
// then one LDS at a 32-bit offset (no 64-bit address arithmetic, lower latency than an L1 hit).
#if defined(__CUDACC__)
// (static: liblsim.so is built from two translation units, each with its own copy)
static __device__ unsigned long long d_exp2_tab[LSIM_EXP2_N] = LSIM_EXP2_TABLE;
static __device__ double d_log2_tab[2 * LSIM_LOG2_N] = LSIM_LOG2_TABLE;
__shared__ unsigned long long s_exp2_tab[LSIM_EXP2_N];
__shared__ double2 s_log2_tab[LSIM_LOG2_N];
// call at kernel entry, followed by a __syncthreads() before the first log2_pos / exp2_*
__device__ __forceinline__ void stage_math_tables() {
  for (int i = threadIdx.x; i < LSIM_EXP2_N; i += blockDim.x) s_exp2_tab[i] = d_exp2_tab[i];
  for (int i = threadIdx.x; i < LSIM_LOG2_N; i += blockDim.x)
    s_log2_tab[i] = reinterpret_cast<const double2*>(d_log2_tab)[i];
}
#endif
static const unsigned long long h_exp2_tab[LSIM_EXP2_N] = LSIM_EXP2_TABLE;
static const double h_log2_tab[2 * LSIM_LOG2_N] = LSIM_LOG2_TABLE;

// The polynomial coefficients again, as a plain struct: the image kernels carry it in their kernel parameters
// (PlanDev.ray.mk), where the compiler reads it through uniform registers -- a DFMA with a uniform-register operand
// reads two vector registers and issues at full rate, one whose coefficient sits in a vector register (which is what
// the compiler does with __constant__ data it uses more than once) reads three and takes 3 cycles instead of 2
// (profiles/r1_dfma_operand_probe.md).
struct MathK {
  double l[4];  // kLog2C
  double e[3];  // kExp2S
};
LSIM_HD MathK math_k() {
  MathK k;
  k.l[0] = -kLog2e / 4.0;
  k.l[1] = kLog2e / 3.0;
  k.l[2] = -kLog2e / 2.0;
  k.l[3] = kLog2e;
  k.e[0] = kLn2 * kLn2 * kLn2 / 6.0 / (kExp2Scale * kExp2Scale * kExp2Scale);
  k.e[1] = kLn2 * kLn2 / 2.0 / (kExp2Scale * kExp2Scale);
  k.e[2] = kLn2 / kExp2Scale;
  return k;
}

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

// 1/sqrt(x) for finite, normal, positive x: MUFU.RSQ64H seed (relative error <= 2^-22) and ONE third-order step
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2: 5 fp64 instructions, error ~e^3 before rounding (a couple of ulp at most).
// For consumers whose own accuracy bar is far above the last bit (the 44-term CSE sums: x y gives the square root).
LSIM_HD double rsqrt_fast3(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y, e * fma(0.375, e, 0.5), y);
#else
  return 1.0 / sqrt(x);
#endif
}

// ---- fast fp64 log2 / exp2 --------------------------------------------------------------------------
// The hot loop needs pow(x, s) = exp2(s * log2(x)) and exp(-b (k - 1)) for finite positive normal x.
// libdevice's log/exp/pow carry special-case branches and long polynomials; on a kernel bound by the fp64 pipe
// every DFMA counts and every branch cuts the instruction stream into blocks the scheduler cannot interleave,
// so both functions are table-driven and BRANCH-FREE here:
//   log2_pos(x): x positive, finite, normal.  6 fp64 instructions.  |abs err| < 3e-16 max(1, |log2 x|)
//                (the result feeds exponents, where absolute accuracy is what matters).
//   exp2_fin(x): any x; 6 fp64 instructions (5 for a pre-scaled argument, exp2_fin_s); |rel err| < 4e-16.  Results
//                saturate at 2^-1022 (2.2e-308, instead of denormals / 0) and 2^1023 (instead of +inf) through one
//                integer clamp of the table index; nan propagates, +-inf gives nan.
//   exp2_neg(x): kept as a name for call sites whose argument is bounded above; the same function.

// x = 2^e m, m in [1,2); bin i = top 10 mantissa bits; u = m R_i - 1, |u| <= 2^-11;
// log2 x = e + L_i + log2(1 + u)
template <class CoefT>
LSIM_HD double log2_pos_c(double x, const CoefT& C) {
  const int64_t ix = d2i(x);
  const uint32_t hx = (uint32_t)((uint64_t)ix >> 32);
  const int e = (int)(hx >> 20) - 1023;
  const double m = i2d((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
#if defined(__CUDA_ARCH__)
  // byte offset of the bin's {R, L} pair: (top LSIM_LOG2_BITS mantissa bits) * 16 in one shift and one mask
  const double2 rl = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(s_log2_tab) +
                                                       ((hx >> (16 - LSIM_LOG2_BITS)) & ((LSIM_LOG2_N - 1) << 4)));
  const double R = rl.x, L = rl.y;
#else
  const int i = (int)(hx >> (20 - LSIM_LOG2_BITS)) & (LSIM_LOG2_N - 1);
  const double R = h_log2_tab[2 * i], L = h_log2_tab[2 * i + 1];
#endif
  const double u = fma(m, R, -1.0);
  double p = fma(C[0], u, C[1]);
  p = fma(p, u, C[2]);
  p = fma(p, u, C[3]);
  return (double)e + fma(u, p, L);
}
LSIM_HD double log2_pos(double x) { return log2_pos_c(x, kLog2C); }
LSIM_HD double log2_pos(double x, const MathK& K) { return log2_pos_c(x, K.l); }

// ---- exp2 ------------------------------------------------------------------------------------------------------------
// Core, for a PRE-SCALED argument y = N x, N = LSIM_EXP2_N = 2048 = kExp2Scale (the hot profiles compute their exponents
// with one FMA from prepared constants, which carry the factor N): k = rint(y), r = y - k (exact), k clamped to
// [-1022 N, 1022 N + N - 1] with ONE integer min/max, j = k & (N - 1), E = k >> 11:
//   2^x = (2^E T[j]) (1 + r P(r)),
// where 2^E T[j] costs one integer multiply-add on the high word of the table entry (the table stores T[j] with
// j << 9 subtracted from its high word, so hi + (k << 9) = hi(T[j]) + (E << 20)).  P's coefficients are those of
// (2^r - 1)/r through r^2, scaled by powers of N (exact); |r|/N <= 2^-12, so the first dropped term is
// (ln2 2^-12)^4 / 24 = 3.4e-17.  5 fp64 instructions; the result saturates at 2^-1022 / 2^1023 for finite arguments out of
// range (the reduced argument stays |r| <= 1/2 whatever the clamp does), nan and +-inf give nan.
LSIM_TABLE kExp2S[3] = {kLn2 * kLn2 * kLn2 / 6.0 / (kExp2Scale * kExp2Scale * kExp2Scale),
                        kLn2 * kLn2 / 2.0 / (kExp2Scale * kExp2Scale), kLn2 / kExp2Scale};

template <class CoefT>
LSIM_HD double exp2_core_s(double y, const CoefT& C) {
#if !defined(__CUDA_ARCH__)
  if (!(y - y == 0.0)) return y - y;  // nan, +-inf -> nan (the device conversion saturates and r = y - k is nan)
  if (y > 4e9) y = 4e9;               // (int) of an out-of-range double is undefined on the host; k is clamped anyway
  if (y < -4e9) y = -4e9;
#endif
  const double kf = rint(y);
  int ki = (int)(long long)kf;
#if !defined(__CUDA_ARCH__)
  if (kf > 2e9) ki = 2000000000;
  if (kf < -2e9) ki = -2000000000;
#endif
  const double r = y - kf;  // exact: |r| <= 1/2 and both are multiples of ulp(y)
  constexpr int kLo = -1022 * LSIM_EXP2_N, kHi = 1022 * LSIM_EXP2_N + LSIM_EXP2_N - 1;
  ki = ki < kLo ? kLo : (ki > kHi ? kHi : ki);
  double p = fma(C[0], r, C[1]);
  p = fma(p, r, C[2]);
#if defined(__CUDA_ARCH__)
  const unsigned long long tb = s_exp2_tab[ki & (LSIM_EXP2_N - 1)];
  const double T = __hiloint2double((int)((uint32_t)(tb >> 32) + ((uint32_t)ki << (20 - LSIM_EXP2_BITS))), (int)(uint32_t)tb);
#else
  const unsigned long long tb = h_exp2_tab[ki & (LSIM_EXP2_N - 1)];
  const double T = i2d((int64_t)(tb + ((unsigned long long)((uint32_t)ki << (20 - LSIM_EXP2_BITS)) << 32)));
#endif
  return fma(T * r, p, T);
}

LSIM_HD double exp2_fin_s(double y) { return exp2_core_s(y, kExp2S); }
LSIM_HD double exp2_fin_s(double y, const MathK& K) { return exp2_core_s(y, K.e); }
LSIM_HD double exp2_neg_s(double y) { return exp2_core_s(y, kExp2S); }
// "coefficients from the __constant__ tables" as a type, for code templated on where the coefficients come from
struct TableK {};
LSIM_HD double log2_pos(double x, const TableK&) { return log2_pos_c(x, kLog2C); }
LSIM_HD double exp2_fin_s(double y, const TableK&) { return exp2_core_s(y, kExp2S); }
// unscaled arguments (everything off the hot path)
LSIM_HD double exp2_fin(double x) { return exp2_core_s(x * kExp2Scale, kExp2S); }
LSIM_HD double exp2_neg(double x) { return exp2_core_s(x * kExp2Scale, kExp2S); }

}  // namespace lsim

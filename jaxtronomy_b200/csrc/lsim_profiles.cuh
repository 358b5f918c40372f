// lsim_profiles.cuh -- per-parameter-set constant preparation and per-point evaluation of every lens and
// light profile on the hot path.  Host + device (see lsim_math.cuh for why).
//
// Split of work, B200-first: everything that depends only on the parameter set (ellipticity -> (phi, q),
// trig of the position angle, the 2F1 series coefficients F_n(t), Sersic b_n, normalisations ...) is
// computed ONCE per CTA into shared memory (`*_prepare`), so the per-point code (`*_eval`) is a straight
// FMA stream with at most one log2, two exp2 and one reciprocal for the headline EPL+SHEAR+SERSIC path.
//
// Reference arithmetic being reproduced (clamps and sentinels included), under
// /root/reference/jaxtronomy/:
//   EPL            LensModel/Profiles/epl.py:93-124,178-200,297-319,366-394 + Util/hyp2f1_util.py:135-176
//   SIE / NIE      LensModel/Profiles/sie.py:56-63,104-124, nie.py:85-107,181-200,230-246,280-294,333-342
//   SIS            LensModel/Profiles/sis.py:35-43          SHEAR  shear.py:39-53   CONVERGENCE convergence.py:35-46
//   NFW            LensModel/Profiles/nfw.py:88-105,259-279,347-365,386-395
//   TNFW           LensModel/Profiles/tnfw.py:43,69-87,228-248,278-349
//   MULTIPOLE      LensModel/Profiles/multipole.py:72-119
//   EPL_MULTIPOLE_M1M3M4  LensModel/Profiles/epl_multipole_m1m3m4.py:463-535,608-678
//   CSE family     LensModel/Profiles/cored_steep_ellipsoid.py:107-132,227-245,434-477,536-564,
//                  nfw_ellipse_cse.py:250-280,338-349, hernquist_ellipse_cse.py:151-164,190-202
//   SERSIC(_ELLIPSE)  LightModel/Profiles/sersic_ellipse.py:61-93, sersic.py:54-76,
//                  LensModel/Profiles/sersic_utils.py:67-110,238-267, Util/param_util.py:102-123
//   GAUSSIAN       LightModel/Profiles/gaussian.py:27-42
//   helpers        Util/util.py:102-139 (rotate, shift_center), Util/param_util.py:6-24,87-99
#pragma once

#include "../../include/lsim.h"
#include "lsim_cse_tables.h"
#include "lsim_math.cuh"

namespace lsim {

constexpr int kMaxSeries = 512;  // >= the reference's largest nmax (500, epl.py:317)

// Prepared (per parameter set) constants of one lens profile.
struct LensC {
  int32_t type;
  uint32_t flags;
  int32_t n_coef;    // EPL family: number of series coefficients actually used
  int32_t coef_off;  // EPL family: offset of this profile's coefficients in the coefficient pool
  double delta_T, inv_T_z, red2phys;  // multi-plane stepping
  double c[31];  // (an odd count keeps sizeof(LensC) a multiple of 16: the coefficient pools behind stay 16-byte aligned)
};

struct LightC {
  int32_t type;
  int32_t lensed;
  int32_t slot;
  int32_t pad;
  double c[14];
};

// ---- CSE tables: device copies live in constant memory, host copies are the header arrays -------------
// (static + uploaded by lsim_api.cu's upload_tables(): only kernels of THAT translation unit may read them; the
// kernels instantiated in lsim_hot.cu -- k_image_fold, k_likelihood, k_conv_tiled_u -- do not)
#if defined(__CUDACC__)
static __constant__ double d_nfw_cse_high_s[LSIM_NFW_CSE_HIGH_S_N];
static __constant__ double d_nfw_cse_high_a[LSIM_NFW_CSE_HIGH_A_N];
static __constant__ double d_nfw_cse_low_s[LSIM_NFW_CSE_LOW_S_N];
static __constant__ double d_nfw_cse_low_a[LSIM_NFW_CSE_LOW_A_N];
static __constant__ double d_hernquist_cse_s[LSIM_HERNQUIST_CSE_S_N];
static __constant__ double d_hernquist_cse_a[LSIM_HERNQUIST_CSE_A_N];
#endif

LSIM_HD const double* cse_table(int which) {
#if defined(__CUDA_ARCH__)
  switch (which) {
    case 0: return d_nfw_cse_high_s;
    case 1: return d_nfw_cse_high_a;
    case 2: return d_nfw_cse_low_s;
    case 3: return d_nfw_cse_low_a;
    case 4: return d_hernquist_cse_s;
    default: return d_hernquist_cse_a;
  }
#else
  switch (which) {
    case 0: return LSIM_NFW_CSE_HIGH_S;
    case 1: return LSIM_NFW_CSE_HIGH_A;
    case 2: return LSIM_NFW_CSE_LOW_S;
    case 3: return LSIM_NFW_CSE_LOW_A;
    case 4: return LSIM_HERNQUIST_CSE_S;
    default: return LSIM_HERNQUIST_CSE_A;
  }
#endif
}

// ---- parameter fetch ------------------------------------------------------------------------------------
template <class Desc>
LSIM_HD double param_of(const Desc& d, const double* row, int i) {
  int c = d.col[i];
  return c >= 0 ? row[c] : d.cst[i];
}

// Util/param_util.py:87-99
LSIM_HD void ellipticity2phi_q(double e1, double e2, double& phi, double& q) {
  phi = atan2(e2, e1) / 2;
  double c = sqrt(e1 * e1 + e2 * e2);
  c = fmin(c, 0.9999);
  q = (1 - c) / (1 + c);
}

// nmax schedule of EPLMajorAxis._hyp2f1_evaluate (epl.py:308-319); 0 encodes "return 1" (f == 0)
LSIM_HD int epl_nmax(double f) {
  if (f == 0) return 0;
  if (f < 0.12) return 10;
  if (f < 0.4) return 20;
  if (f < 0.6) return 35;
  if (f < 0.73) return 50;
  if (f < 0.86) return 100;
  if (f < 0.92) return 200;
  return 500;
}

// EPL family constants.  c[0..8]: cx, cy, cos(phi), sin(phi), q, f, -t/2, t*log2(b), 2/(1+q) (the closed-form Hessian
// and the fp32 variant read these); c[19..28]: the folded constants epl_core works from (see there).
// The reference evaluates 2F1(1, t/2; 2 - t/2; z) with the Lopez-Temme series truncated at nmax terms:
//   2F1 = (1 - z/2)^-1 * sum_{n<nmax} F_n w^n,  w = z/(z-2),
//   F_0 = 1, F_1 = 1 - 2b/c, F_n = ((n-1) F_{n-2} - (2b-c) F_{n-1}) / (c+n-1)   (hyp2f1_util.py:152-176, a = 1)
// F_n depends on t only, so it is tabulated here once per parameter set.  |w| <= f/(2-f) =: wmax for
// every point, so terms with |F_n| wmax^n below `series_tol` relative to the leading term cannot change
// the truncated sum at that level and are dropped (series_tol = 0 keeps all nmax terms).
LSIM_HD void epl_prepare(LensC& L, double* coef, double theta_E, double gamma, double e1, double e2, double cx,
                         double cy, double series_tol) {
  double t = gamma - 1;
  double phi, q;
  ellipticity2phi_q(e1, e2, phi, q);
  double b = theta_E * sqrt(q);
  double f = (1.0 - q) / (1.0 + q);
  L.c[0] = cx;
  L.c[1] = cy;
  L.c[2] = cos(phi);
  L.c[3] = sin(phi);
  L.c[4] = q;
  L.c[5] = f;
  L.c[6] = -0.5 * t;
  L.c[7] = t * log2(b);
  L.c[8] = 2.0 / (1.0 + q);
  // epl_core: Z = a + i c with a = q x', c = y' straight from the sky coordinates (shift and rotation folded)
  const double fe = f > 1e-150 ? f : 1e-150;  // q = 1: every f-dependent term vanishes like f; keep the ratios finite
  L.c[19] = q * L.c[2];
  L.c[20] = q * L.c[3];
  L.c[21] = -(L.c[19] * cx + L.c[20] * cy);
  L.c[22] = L.c[3] * cx - L.c[2] * cy;
  L.c[23] = 0.5 * fe;
  L.c[24] = (4.0 - f * f) / (4.0 * fe);
  L.c[25] = -0.25 * fe;
  L.c[26] = (2.0 - f) / (2.0 + f);
  L.c[27] = kExp2Scale * L.c[6];
  L.c[28] = kExp2Scale * (L.c[7] + log2((2.0 + f) / (fe * (1.0 + q))));
  int nmax = epl_nmax(f);
  double B = t / 2.0, C = 2.0 - B;
  // epl_core takes term 1 unconditionally: at least two coefficients, F_1 = 0 when the reference sums nothing (f == 0)
  coef[0] = 1.0;
  coef[1] = 0.0;
  int n_used = 2;
  if (nmax > 0) {
    double ppF = 1.0, pF = 1.0 - 2.0 * B / C;
    coef[1] = pF;
    n_used = 2;
    double wmax = f / (2.0 - f);
    double wn = wmax;  // wmax^n
    for (int n = 2; n < nmax; ++n) {
      double nF = ((n - 1.0) * ppF - (2.0 * B - C) * pF) / (C + n - 1.0);
      coef[n] = nF;
      ppF = pF;
      pF = nF;
      wn *= wmax;
      if (fabs(nF) * wn > series_tol) n_used = n + 1;
    }
    // the bound must also hold for everything after the cut: |F_n| is non-increasing in practice, and
    // wmax^n decays geometrically, so the last term above the tolerance closes the sum; keep one extra.
    if (n_used < nmax) n_used += 1;
  }
  L.n_coef = n_used;
}

// Deflection of the EPL for J points at once.  J independent points are evaluated in lock step so that the series
// coefficients are read from shared memory once per J points and the compiler can interleave J independent FMA chains.
//
// Tessore & Metcalf eq. 22/23, reorganised.  With Z = a + i c (a = q x', c = y'), R2 = |Z|^2, U = Z^2 and the series
// argument w = z/(z-2), z = -f Z/conj(Z):
//   w = f (f R2 + 2 U)/E,  E = (4 + f^2) R2 + 4 f Re U  (>= (2-f)^2 R2),  |w|^2 = f^2 R2/E,
//   Z (1 - z/2)^-1 = (2 R2/E) ((2+f) a + i (2-f) c),
// so one reciprocal (of E/4f), one log2 and one exp2 are the only non-FMA work per point:
//   n = Re U + (f/2) R2,  Et = n + ((4-f^2)/4f) R2 = E/4f,  iE = 1/Et,
//   2 Re w = n iE,  Im w = a c iE,  h = R2 iE,  -|w|^2 = -(f/4) h,
//   alpha' = s (a S_r - rho c S_i, a S_i + rho c S_r),  rho = (2-f)/(2+f),
//   s = h exp2(-t/2 log2 R2 + t log2 b + log2((2+f)/(f (1+q))))        [= 2/(1+q) (b/R)^t 2 R2 (2+f)/E]
// (every constant factor of the power law lives in the exponent).  The series S = sum_n F_n w^n has REAL
// coefficients and is summed with the quadratic-factor (Goertzel / Clenshaw) recurrence
// b_k = F_k + p b_{k+1} - |w|^2 b_{k+2}, p = 2 Re w, S = F_0 - |w|^2 b_2 + b_1 w: two real FMAs per term instead of the
// four of a complex Horner step; |w| <= f/(2-f) < 1, so the recurrence is contractive.
// Output: the deflection in the sky frame is  s[j] * (u1[j], u2[j])  (the caller folds s into its accumulation FMA).
// Non-finite results (Z = 0) follow the reference: nan_to_num(posinf=1e10, neginf=-1e10) on the major-axis
// components (epl.py:391-392), one test for all J points; then s = 1 and (u1, u2) is the rotated, cleaned deflection.
// ROT = false: (u1, u2) stay in the major-axis frame (the folded image kernel composes the rotation with the
// light profile's own transform).
template <int J, class KT, bool ROT = true>
LSIM_HD void epl_core(const LensC& L, const double* coef, const double (&a)[J], const double (&c)[J], double (&s)[J],
                      double (&u1)[J], double (&u2)[J], const KT& K) {
  // seven values per point stay live across the series loop besides the recurrence state: a, c, s, Im w, p, -|w|^2
  double wi[J], pp[J], qq[J], b1[J], b2[J];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    const double aa = a[j] * a[j];
    const double R2 = fma(c[j], c[j], aa);
    const double n = fma(L.c[23], R2, fma(-c[j], c[j], aa));
    const double iE = rcp_fast(fma(L.c[24], R2, n));
    pp[j] = n * iE;
    wi[j] = (a[j] * c[j]) * iE;
    const double h = R2 * iE;
    qq[j] = L.c[25] * h;
    const double R2c = max_nan(R2, 1e-18);  // R = max(|Z|, 1e-9)
    s[j] = h * exp2_fin_s(fma(L.c[27], log2_pos(R2c, K), L.c[28]), K);
  }
  // Clenshaw over the terms k = n_coef-1 ... 1, two per trip with the roles of b1 / b2 swapping (no register moves);
  // an even count peels its first term, whose step is just b1 = F_k (b1 = b2 = 0 before it).  Term 1 is taken
  // together with the closing step below.  The pairs (F_{k-1}, F_k), k odd, are 16-byte aligned: one LDS.128.
  int k = L.n_coef - 1;
  const bool peel = k >= 1 && (k & 1) == 0;
  const double ctop = peel ? coef[k] : 0.0;
  k -= peel ? 1 : 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    b1[j] = ctop;
    b2[j] = 0.0;
  }
  for (; k >= 3; k -= 2) {
#if defined(__CUDA_ARCH__)
    const double2 cc = *reinterpret_cast<const double2*>(coef + k - 1);
    const double ck = cc.y, ckm = cc.x;
#else
    const double ck = coef[k], ckm = coef[k - 1];
#endif
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      b2[j] = fma(pp[j], b1[j], fma(qq[j], b2[j], ck));
      b1[j] = fma(pp[j], b2[j], fma(qq[j], b1[j], ckm));
    }
  }
  // F_0 = 1 for every parameter set (hyp2f1_util.py:152): an immediate operand, not a register.  epl_prepare always
  // keeps at least two coefficients (F_1 = 0 for f == 0, where the reference returns 1 without summing), so the loop
  // above always ends at k = 1: term 1 is unconditional.
  const double c0 = 1.0, c1 = coef[1], cph = L.c[2], sph = L.c[3], rho = L.c[26];
  double tr[J], ti[J];
  // largest |high word| of s: >= 0x7ff00000 <=> some s is inf / nan.  s alone is enough: a non-finite coordinate,
  // parameter or 1/E reaches s through h = R2 / E (0 x inf, inf x 0, nan), and with finite inputs the series factors are
  // bounded (|w| < 1).
  uint32_t worst = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    // term 1 (if any): (b1, b2) <- (F_1 + p b1 + q b2, b1); then S = F_0 + w b1 + q b2 with Re w = p/2
    const double bn = fma(pp[j], b1[j], fma(qq[j], b2[j], c1));
    const double bo = b1[j];
    const double Sr = fma(bn, 0.5 * pp[j], fma(qq[j], bo, c0)), Si = bn * wi[j];
    const double cr = rho * c[j];
    tr[j] = fma(a[j], Sr, -(cr * Si));
    ti[j] = fma(a[j], Si, cr * Sr);
    const uint32_t h2 = (uint32_t)((uint64_t)d2i(s[j]) >> 32) & 0x7fffffffu;
    worst = worst > h2 ? worst : h2;
  }
  if (worst >= 0x7ff00000u) {  // rare
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      tr[j] = nan_to_num(s[j] * tr[j], 1e10, -1e10);
      ti[j] = nan_to_num(s[j] * ti[j], 1e10, -1e10);
      s[j] = 1.0;
    }
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    u1[j] = ROT ? fma(-ti[j], sph, tr[j] * cph) : tr[j];  // rotate(f, -phi)
    u2[j] = ROT ? fma(ti[j], cph, tr[j] * sph) : ti[j];
  }
}

// acc += w * alpha(x, y) for J sky positions.  UNIT: w == 1 at compile time (single-plane lensing).
template <int J, bool UNIT, class KT>
LSIM_HD void epl_accum(const LensC& L, const double* coef, const double (&x)[J], const double (&y)[J], double w,
                       double (&accx)[J], double (&accy)[J], const KT& K) {
  double a[J], c[J], s[J], u1[J], u2[J];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    a[j] = fma(L.c[19], x[j], fma(L.c[20], y[j], L.c[21]));
    c[j] = fma(L.c[2], y[j], fma(-L.c[3], x[j], L.c[22]));
  }
  epl_core<J>(L, coef, a, c, s, u1, u2, K);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    const double sw = UNIT ? s[j] : s[j] * w;
    accx[j] = fma(sw, u1[j], accx[j]);
    accy[j] = fma(sw, u2[j], accy[j]);
  }
}

// Deflection at J points given relative to the profile centre (the form the closed-form Hessian and the multipole
// variant use).
template <int J>
LSIM_HD void epl_eval(const LensC& L, const double* coef, const double (&xs)[J], const double (&ys)[J],
                      double (&ax)[J], double (&ay)[J]) {
  double a[J], c[J], s[J], u1[J], u2[J];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    a[j] = fma(L.c[19], xs[j], L.c[20] * ys[j]);
    c[j] = fma(L.c[2], ys[j], -(L.c[3] * xs[j]));
  }
  epl_core<J>(L, coef, a, c, s, u1, u2, TableK());
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    ax[j] = s[j] * u1[j];
    ay[j] = s[j] * u2[j];
  }
}

LSIM_HD void epl_eval1(const LensC& L, const double* coef, double xs, double ys, double& ax, double& ay) {
  const double x1[1] = {xs}, y1[1] = {ys};
  double fx1[1], fy1[1];
  epl_eval<1>(L, coef, x1, y1, fx1, fy1);
  ax = fx1[0];
  ay = fy1[0];
}

// cos / sin of m*(phi - phi_m) and cos/sin(phi) from the shifted point, no trig calls:
// u = (x + i y)/r = e^{i phi}; u^m by repeated multiplication; rotate by e^{-i m phi_m}.
LSIM_HD void multipole_terms(double xs, double ys, int m, double cmp, double smp, double& cphi, double& sphi,
                             double& cmd, double& smd, double& r2) {
  r2 = xs * xs + ys * ys;
  if (r2 > 0) {
    double ir = 1.0 / sqrt(r2);
    cphi = xs * ir;
    sphi = ys * ir;
  } else {  // atan2(0, 0) = 0
    cphi = 1.0;
    sphi = 0.0;
  }
  double ur = cphi, ui = sphi;
  for (int k = 1; k < m; ++k) {
    double tr = ur * cphi - ui * sphi;
    ui = ur * sphi + ui * cphi;
    ur = tr;
  }
  cmd = ur * cmp + ui * smp;
  smd = ui * cmp - ur * smp;
}

// multipole.py:72-119.  A = a_m/(1-m^2) (m != 1) or a_m/2 (m == 1); (cmp, smp) = cos/sin(m phi_m); lnrE = ln r_E
LSIM_HD void multipole_eval(double xs, double ys, int m, double A, double cmp, double smp, double lnrE,
                            double& fx, double& fy) {
  double cphi, sphi, cmd, smd, r2;
  multipole_terms(xs, ys, m, cmp, smp, cphi, sphi, cmd, smd, r2);
  if (m == 1) {
    double r2c = fmax(r2, 1e-12);  // r = max(r, 1e-6)
    double lg = 0.5 * kLn2 * log2_pos(r2c) - lnrE;
    fx = A * (cmp * lg + cmd * cphi);
    fy = A * (smp * lg + cmd * sphi);
  } else {
    fx = cphi * A * cmd + sphi * m * A * smd;
    fy = sphi * A * cmd - cphi * m * A * smd;
  }
}

// The m = 1, 3, 4 multipoles of EPL_MULTIPOLE_M1M3M4 at one point, fused: one reciprocal square root, the powers of
// u = e^{i phi} by complex multiplication, no trigonometric call (epl_multipole_m1m3m4.py:608-678 adds
// Multipole.derivatives for m = 1, 3, 4; same expressions as multipole_eval above).  L.c[9..18] as prepared.
LSIM_HD void multipole_m1m3m4_add(const LensC& L, double xs, double ys, double& fx, double& fy) {
  const double r2 = fma(ys, ys, xs * xs);
  double c1 = 1.0, s1 = 0.0;  // atan2(0, 0) = 0
  if (r2 > 0) {
    const double ir = rsqrt_fast(fmax(r2, 1e-300));
    c1 = xs * ir;
    s1 = ys * ir;
  }
  const double c2 = fma(c1, c1, -(s1 * s1)), s2 = 2.0 * c1 * s1;
  const double c3 = fma(c2, c1, -(s2 * s1)), s3 = fma(c2, s1, s2 * c1);
  const double c4 = fma(c2, c2, -(s2 * s2)), s4 = 2.0 * c2 * s2;
  double gx = 0.0, gy = 0.0;
  {  // m = 3: A = c[13], (cos, sin)(3 phi_3) = c[14], c[15]
    const double cmd = fma(s3, L.c[15], c3 * L.c[14]), smd = fma(s3, L.c[14], -(c3 * L.c[15]));
    const double a = L.c[13] * cmd, b = 3.0 * L.c[13] * smd;
    gx += fma(s1, b, c1 * a);
    gy += fma(-c1, b, s1 * a);
  }
  {  // m = 4
    const double cmd = fma(s4, L.c[18], c4 * L.c[17]), smd = fma(s4, L.c[17], -(c4 * L.c[18]));
    const double a = L.c[16] * cmd, b = 4.0 * L.c[16] * smd;
    gx += fma(s1, b, c1 * a);
    gy += fma(-c1, b, s1 * a);
  }
  {  // m = 1: A = c[9], (cos, sin)(phi_1) = c[10], c[11], ln r_E = c[12]
    const double cmd = fma(s1, L.c[11], c1 * L.c[10]);
    const double lg = 0.5 * kLn2 * log2_pos(fmax(r2, 1e-12)) - L.c[12];  // r = max(r, 1e-6)
    gx += L.c[9] * fma(L.c[10], lg, cmd * c1);
    gy += L.c[9] * fma(L.c[11], lg, cmd * s1);
  }
  fx += gx;
  fy += gy;
}

// nfw.py:347-365
LSIM_HD double nfw_g(double X) {
  const double c = 0.000001;
  if (X < c) X = c;
  if (X < 1) return log(X / 2.0) + 1 / sqrt(1 - X * X) * acosh(1.0 / X);
  if (X == 1) return 1 + log(1.0 / 2.0);
  return log(X / 2) + 1 / sqrt(X * X - 1) * acos(1.0 / X);
}

// tnfw.py:290-304
LSIM_HD double tnfw_F(double x) {
  if (x < 1) {
    double s = sqrt(1 - x * x);
    return atanh(s) / s;
  }
  if (x > 1) {
    double s = sqrt(x * x - 1);
    return atan(s) / s;
  }
  return x;
}

// cored_steep_ellipsoid.py:227-245 summed over a table, x,y already in units of Rs; product-average variant
// applies cored_steep_ellipsoid.py:434-443,463-477 per component.
LSIM_HD void cse_set_eval(double x, double y, double q, const double* A, const double* S, int n, bool prod_avg,
                          double& fx, double& fy) {
  double sq = sqrt(q);
  double xx = prod_avg ? x * sq : x, yy = prod_avg ? y * sq : y;
  double q2 = q * q;
  const double x2 = xx * xx, qy = q2 * x2 + yy * yy, omq = (1 - q2) * x2, iq = prod_avg ? 1.0 / q : 1.0;
  double sx = 0, sy = 0;
  // per component: one inverse square root (MUFU seed + one third-order step; psi = u / sqrt(u)) and one reciprocal
  // (MUFU seed + one third-order step) instead of the library's sqrt / division -- 18 fp64 instructions per component,
  // 44 components per pixel for NFW_ELLIPSE_CSE: this loop is the whole cost of the profile.  The amplitudes' common
  // factor 1/q (product average) is applied to the sums.  psi^2 = q^2 (s^2 + x^2) + y^2 >= q^2 s^2 > 0 for the tabulated s > 0
  for (int i = 0; i < n; ++i) {
    const double s = S[i];
    const double q2s = q2 * s;
    const double u = fma(q2s, s, qy);
    const double psi = u * rsqrt_fast3(u);
    const double ps = psi + s;
    const double Phi = fma(ps, ps, omq);
    const double den = A[i] * rcp_fast(s * psi * Phi);
    sx = fma(den, q2s + psi, sx);
    sy = fma(den, ps, sy);
  }
  sx *= iq;
  sy *= iq;
  sx *= q * xx;
  sy *= q * yy;
  fx = prod_avg ? sq * sx : sx;
  fy = prod_avg ? sq * sy : sy;
}

// ---- lens: prepare ---------------------------------------------------------------------------------------
// `coef` points at this profile's kMaxSeries-sized slice of the coefficient pool (EPL family only).
LSIM_HD void lens_prepare(const lsim_lens_desc& d, const double* row, LensC& L, double* coef, double series_tol) {
  L.type = d.type;
  L.flags = d.flags;
  L.n_coef = 0;
  L.delta_T = d.delta_T;
  L.inv_T_z = d.T_z != 0 ? 1.0 / d.T_z : 0.0;
  L.red2phys = d.red2phys;
  auto P = [&](int i) { return param_of(d, row, i); };
  switch (d.type) {
    case LSIM_LENS_EPL:
      epl_prepare(L, coef, P(0), P(1), P(2), P(3), P(4), P(5), series_tol);
      break;
    case LSIM_LENS_EPL_MULTIPOLE_M1M3M4: {
      double theta_E = P(0), e1 = P(2), e2 = P(3);
      epl_prepare(L, coef, theta_E, P(1), e1, e2, P(4), P(5), series_tol);
      double phi, q;
      ellipticity2phi_q(e1, e2, phi, q);
      double rescale = theta_E / sqrt(q);
      double a1 = P(6) * rescale, p1 = phi + P(7);
      double a3 = P(8) * rescale, p3 = phi + P(9);
      double a4 = P(10) * rescale, p4 = phi + P(11);
      L.c[9] = a1 / 2;
      L.c[10] = cos(p1);
      L.c[11] = sin(p1);
      L.c[12] = log(theta_E);
      L.c[13] = a3 / (1 - 9);
      L.c[14] = cos(3 * p3);
      L.c[15] = sin(3 * p3);
      L.c[16] = a4 / (1 - 16);
      L.c[17] = cos(4 * p4);
      L.c[18] = sin(4 * p4);
      break;
    }
    case LSIM_LENS_SIE:
    case LSIM_LENS_NIE: {
      double theta_E = P(0), e1 = P(1), e2 = P(2), s_scale, cx, cy;
      if (d.type == LSIM_LENS_SIE) {
        cx = P(3);
        cy = P(4);
        if (d.flags & LSIM_FLAG_SIE_NIE_FALSE) {
          L.type = LSIM_LENS_EPL;
          epl_prepare(L, coef, theta_E, d.opt[1], e1, e2, cx, cy, series_tol);
          break;
        }
        s_scale = d.opt[0];
      } else {
        s_scale = P(3);
        cx = P(4);
        cy = P(5);
      }
      L.type = LSIM_LENS_NIE;
      double phi, q;
      ellipticity2phi_q(e1, e2, phi, q);
      double theta_conv = theta_E / sqrt((1.0 + q * q) / (2.0 * q));
      double b = theta_conv * sqrt((1 + q * q) / 2);
      double s = s_scale / sqrt(q);
      if (q >= 1) q = 0.99999999;  // NIEMajorAxis.derivatives, nie.py:284
      double rt = sqrt(1.0 - q * q);
      L.c[0] = cx;
      L.c[1] = cy;
      L.c[2] = cos(phi);
      L.c[3] = sin(phi);
      L.c[4] = q * q;
      L.c[5] = s;
      L.c[6] = rt;
      L.c[7] = b / rt;
      L.c[8] = q * q * s;
      break;
    }
    case LSIM_LENS_SIS:
      L.c[0] = P(1);
      L.c[1] = P(2);
      L.c[2] = P(0);
      break;
    case LSIM_LENS_SHEAR:
      L.c[0] = P(2);
      L.c[1] = P(3);
      L.c[2] = P(0);
      L.c[3] = P(1);
      L.c[4] = -(L.c[2] * L.c[0] + L.c[3] * L.c[1]);  // alpha = (g1 x + g2 y + c4, g2 x - g1 y + c5)
      L.c[5] = -(L.c[3] * L.c[0] - L.c[2] * L.c[1]);
      break;
    case LSIM_LENS_CONVERGENCE:
      L.c[0] = P(1);
      L.c[1] = P(2);
      L.c[2] = P(0);
      break;
    case LSIM_LENS_NFW: {
      double Rs = P(0), alpha_Rs = P(1);
      double rho0 = alpha_Rs / (4.0 * Rs * Rs * (1.0 + log(1.0 / 2.0)));
      if (Rs < 0.0000001) Rs = 0.0000001;
      L.c[0] = P(2);
      L.c[1] = P(3);
      L.c[2] = Rs;
      L.c[3] = 4 * rho0 * Rs;
      break;
    }
    case LSIM_LENS_TNFW: {
      double Rs = P(0), alpha_Rs = P(1), r_trunc = P(2);
      double rho0 = alpha_Rs / (4.0 * Rs * Rs * (1.0 + log(1.0 / 2.0)));
      double tau = r_trunc / Rs, t2 = tau * tau;
      L.c[0] = P(3);
      L.c[1] = P(4);
      L.c[2] = Rs;
      L.c[3] = 4 * rho0 * Rs;
      L.c[4] = tau;
      L.c[5] = t2;
      L.c[6] = t2 / ((t2 + 1) * (t2 + 1));
      L.c[7] = tau * kPi + (t2 - 1) * log(tau);
      L.c[8] = (t2 - 1) / tau;
      L.c[9] = 0.001 * Rs;
      L.c[10] = 1.0 / Rs;
      L.c[11] = (0.001 * Rs) * (0.001 * Rs);
      break;
    }
    case LSIM_LENS_MULTIPOLE: {
      int m = (int)P(0);
      double a_m = P(1), phi_m = P(2);
      L.c[0] = P(3);
      L.c[1] = P(4);
      L.c[2] = (double)m;
      L.c[3] = (m == 1) ? a_m / 2 : a_m / (1 - m * m);
      L.c[4] = cos(m * phi_m);
      L.c[5] = sin(m * phi_m);
      L.c[6] = log(P(5));
      break;
    }
    case LSIM_LENS_CSE:
    case LSIM_LENS_NFW_ELLIPSE_CSE:
    case LSIM_LENS_HERNQUIST_ELLIPSE_CSE: {
      double phi, q;
      ellipticity2phi_q(P(2), P(3), phi, q);
      L.c[0] = P(4);
      L.c[1] = P(5);
      L.c[2] = cos(phi);
      L.c[3] = sin(phi);
      L.c[4] = q;
      if (d.type == LSIM_LENS_CSE) {
        L.c[5] = P(0);  // a
        L.c[6] = P(1);  // s
      } else if (d.type == LSIM_LENS_NFW_ELLIPSE_CSE) {
        double Rs = P(0), alpha_Rs = P(1);
        double rho0 = alpha_Rs / (4.0 * Rs * Rs * (1.0 + log(1.0 / 2.0)));
        L.c[5] = Rs;
        L.c[6] = 4 * rho0 * Rs * Rs * Rs / Rs;  // _normalization / Rs
      } else {
        double sigma0 = P(0), Rs = P(1);
        double rs_ = Rs / sqrt(q);
        L.c[5] = Rs;
        L.c[6] = sigma0 * rs_ * rs_ / Rs;
      }
      break;
    }
    default:
      L.type = LSIM_LENS_NULL;
      break;
  }
}

// ---- lens: evaluate the reduced deflection of one profile at (x, y) ----------------------------------
// `type` is passed explicitly (normally L.type) so that kernels specialised on a compile-time profile list
// can pass a constant and let the compiler drop the dispatch.
template <bool FULL>
LSIM_HD void lens_eval1(int type, const LensC& L, const double* coef_pool, double x, double y, double& fx,
                        double& fy) {
  const double xs = x - L.c[0], ys = y - L.c[1];
  switch (type) {
    case LSIM_LENS_EPL:
      epl_eval1(L, coef_pool + L.coef_off, xs, ys, fx, fy);
      return;
    case LSIM_LENS_SHEAR:
      fx = L.c[2] * xs + L.c[3] * ys;
      fy = L.c[3] * xs - L.c[2] * ys;
      return;
    case LSIM_LENS_SIS: {
      const double eps = 0.000001;
      double R = sqrt(xs * xs + ys * ys);
      if (R < eps) R = eps;
      double a = L.c[2] / R;
      fx = a * xs;
      fy = a * ys;
      return;
    }
    case LSIM_LENS_CONVERGENCE:
      fx = L.c[2] * xs;
      fy = L.c[2] * ys;
      return;
    case LSIM_LENS_NIE: {
      const double cph = L.c[2], sph = L.c[3];
      double xr = xs * cph + ys * sph, yr = -xs * sph + ys * cph;
      double s = L.c[5];
      double psi = sqrt(L.c[4] * (s * s + xr * xr) + yr * yr);
      double gx = L.c[7] * atan(L.c[6] * xr / (psi + s));
      double gy = L.c[7] * atanh(L.c[6] * yr / (psi + L.c[8]));
      fx = gx * cph - gy * sph;
      fy = gx * sph + gy * cph;
      return;
    }
    default:
      break;
  }
  if (FULL) {
    switch (type) {
      case LSIM_LENS_EPL_MULTIPOLE_M1M3M4:
        epl_eval1(L, coef_pool + L.coef_off, xs, ys, fx, fy);
        multipole_m1m3m4_add(L, xs, ys, fx, fy);
        return;
      case LSIM_LENS_MULTIPOLE:
        multipole_eval(xs, ys, (int)L.c[2], L.c[3], L.c[4], L.c[5], L.c[6], fx, fy);
        return;
      case LSIM_LENS_NFW: {
        double R = sqrt(xs * xs + ys * ys);
        R = fmax(R, 0.00000001);
        double X = R / L.c[2];
        double a = L.c[3] * nfw_g(X) / (X * X);
        fx = a * xs;
        fy = a * ys;
        return;
      }
      case LSIM_LENS_TNFW: {
        // tnfw.py:228-349.  Outside the scale radius (x = R/Rs >= 1: all but ~1e-3 of the ray-halo pairs of the
        // substructure configuration) every term is well conditioned and the library sqrt (x3), division (x4) and
        // log are replaced by MUFU-seeded rsqrt / sqrt / reciprocal and the table-driven log2.  Inside (x < 1) the
        // reference's formula cancels O(10) terms down to O(x^2 ln x) and amplifies any 1-ulp difference by ~1/x^4
        // (its own result carries a relative error of 5e-6 at x = 1e-3): there the reference's operation sequence
        // is kept literally, which is the only way to stay within 1e-10 of it.
        const double R2 = fma(ys, ys, xs * xs);
        const double tau = L.c[4], t2 = L.c[5];
        if (R2 * (L.c[10] * L.c[10]) > 1.0000001) {
          // square roots as u / sqrt(u) from one third-order inverse square root each (well inside the 1e-10 bar of this
          // well-conditioned branch); 1 / sqrt(x^2 - 1) is the same number the arctangent's denominator needs
          const double R = R2 * rsqrt_fast3(R2);
          const double X = R * L.c[10];
          const double X2 = X * X;
          const double ut = t2 + X2;
          const double rt = ut * rsqrt_fast3(ut);
          const double Lg = kLn2 * (log2_pos(X) - log2_pos(tau + rt));  // ln(x / (tau + sqrt(tau^2 + x^2)))
          const double vm = X2 - 1.0;
          const double isq = rsqrt_fast3(vm);
          const double sq = vm * isq;
          const double F = atan(sq) * isq;  // tnfw.py:290-304, x > 1
          const double g = L.c[6] * (fma(2.0, X2 - 1.0, t2 + 1.0) * F + L.c[7] + rt * fma(Lg, L.c[8], -kPi));
          const double a = L.c[3] * g * rcp_fast(X2);
          fx = a * xs;
          fy = a * ys;
          return;
        }
        double R = sqrt(R2);
        R = fmax(R, L.c[9]);
        double X = fmax(R / L.c[2], 0.001);
        double rt = sqrt(t2 + X * X);
        double Lg = log(X / (tau + rt));
        double g = L.c[6] * ((t2 + 1 + 2 * (X * X - 1)) * tnfw_F(X) + L.c[7] + rt * (-kPi + Lg * L.c[8]));
        double a = L.c[3] * g / (X * X);
        fx = a * xs;
        fy = a * ys;
        return;
      }
      case LSIM_LENS_CSE:
      case LSIM_LENS_NFW_ELLIPSE_CSE:
      case LSIM_LENS_HERNQUIST_ELLIPSE_CSE: {
        const double cph = L.c[2], sph = L.c[3], q = L.c[4];
        double xr = xs * cph + ys * sph, yr = -xs * sph + ys * cph;
        double gx, gy;
        if (type == LSIM_LENS_CSE) {
          double a = L.c[5], s = L.c[6];
          cse_set_eval(xr, yr, q, &a, &s, 1, !(L.flags & LSIM_FLAG_CSE_MAJOR), gx, gy);
        } else if (type == LSIM_LENS_NFW_ELLIPSE_CSE) {
          bool low = (L.flags & LSIM_FLAG_CSE_LOW_ACCURACY) != 0;
          double iRs = 1.0 / L.c[5];
          cse_set_eval(xr * iRs, yr * iRs, q, cse_table(low ? 3 : 1), cse_table(low ? 2 : 0),
                       low ? LSIM_NFW_CSE_LOW_S_N : LSIM_NFW_CSE_HIGH_S_N, true, gx, gy);
          gx *= L.c[6];
          gy *= L.c[6];
        } else {
          double iRs = 1.0 / L.c[5];
          cse_set_eval(xr * iRs, yr * iRs, q, cse_table(5), cse_table(4), LSIM_HERNQUIST_CSE_S_N, false, gx, gy);
          gx *= L.c[6];
          gy *= L.c[6];
        }
        fx = gx * cph - gy * sph;
        fy = gx * sph + gy * cph;
        return;
      }
      default:
        break;
    }
  }
  fx = 0.0;
  fy = 0.0;
}

// J points at once: the hot profiles (EPL, SHEAR) are evaluated in lock step, everything else point by point.
template <bool FULL, int J>
LSIM_HD void lens_eval(int type, const LensC& L, const double* coef_pool, const double (&x)[J],
                       const double (&y)[J], double (&fx)[J], double (&fy)[J]) {
  if (type == LSIM_LENS_EPL) {
    double xs[J], ys[J];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      xs[j] = x[j] - L.c[0];
      ys[j] = y[j] - L.c[1];
    }
    epl_eval<J>(L, coef_pool + L.coef_off, xs, ys, fx, fy);
    return;
  }
  if (FULL && type == LSIM_LENS_EPL_MULTIPOLE_M1M3M4) {
    double xs[J], ys[J];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      xs[j] = x[j] - L.c[0];
      ys[j] = y[j] - L.c[1];
    }
    epl_eval<J>(L, coef_pool + L.coef_off, xs, ys, fx, fy);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) multipole_m1m3m4_add(L, xs[j], ys[j], fx[j], fy[j]);
    return;
  }
  if (type == LSIM_LENS_SHEAR) {
    const double g1 = L.c[2], g2 = L.c[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      const double xs = x[j] - L.c[0], ys = y[j] - L.c[1];
      fx[j] = fma(g2, ys, g1 * xs);
      fy[j] = fma(-g1, ys, g2 * xs);
    }
    return;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) lens_eval1<FULL>(type, L, coef_pool, x[j], y[j], fx[j], fy[j]);
}

// acc += w * (reduced deflection of one profile) for J points.  UNIT: w == 1 at compile time.  The hot profiles fold
// the weighted sum into their last FMAs; everything else goes through lens_eval.
// CONST_DONE: the position-independent part of the profile (SHEAR: its origin term) is in the accumulators already.
template <bool FULL, int J, bool UNIT, bool CONST_DONE = false, class KT = TableK>
LSIM_HD void lens_accum(int type, const LensC& L, const double* coef_pool, const double (&x)[J], const double (&y)[J],
                        double w, double (&accx)[J], double (&accy)[J], const KT& K = KT()) {
  if (type == LSIM_LENS_EPL) {
    epl_accum<J, UNIT>(L, coef_pool + L.coef_off, x, y, w, accx, accy, K);
    return;
  }
  if (type == LSIM_LENS_SHEAR) {
    const double g1 = UNIT ? L.c[2] : L.c[2] * w, g2 = UNIT ? L.c[3] : L.c[3] * w;
    const double kx = UNIT ? L.c[4] : L.c[4] * w, ky = UNIT ? L.c[5] : L.c[5] * w;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      accx[j] = fma(g1, x[j], fma(g2, y[j], CONST_DONE ? accx[j] : accx[j] + kx));
      accy[j] = fma(g2, x[j], fma(-g1, y[j], CONST_DONE ? accy[j] : accy[j] + ky));
    }
    return;
  }
  double fx[J], fy[J];
  lens_eval<FULL, J>(type, L, coef_pool, x, y, fx, fy);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    accx[j] = UNIT ? accx[j] + fx[j] : fma(w, fx[j], accx[j]);
    accy[j] = UNIT ? accy[j] + fy[j] : fma(w, fy[j], accy[j]);
  }
}

// Compile-time list of effective profile types (LensC::type / LightC::type after *_prepare).  N < 0 means
// "not known at compile time": loops run over the runtime count and dispatch on the prepared type.
template <int... T>
struct TypeList {
  static constexpr int N = sizeof...(T);
  LSIM_HD static constexpr int at(int i) {
    constexpr int v[sizeof...(T) + 1] = {T..., 0};
    return v[i];
  }
};
struct RuntimeTypes {
  static constexpr int N = -1;
  LSIM_HD static constexpr int at(int) { return -1; }
};

// true when the profile needs the FULL kernel variant
LSIM_HD bool lens_type_is_full(int type) {
  switch (type) {
    case LSIM_LENS_NULL:
    case LSIM_LENS_EPL:
    case LSIM_LENS_SIE:
    case LSIM_LENS_NIE:
    case LSIM_LENS_SIS:
    case LSIM_LENS_SHEAR:
    case LSIM_LENS_CONVERGENCE:
      return false;
    default:
      return true;
  }
}

LSIM_HD bool lens_type_has_series(const lsim_lens_desc& d) {
  return d.type == LSIM_LENS_EPL || d.type == LSIM_LENS_EPL_MULTIPOLE_M1M3M4 ||
         (d.type == LSIM_LENS_SIE && (d.flags & LSIM_FLAG_SIE_NIE_FALSE));
}

// ---- rays -----------------------------------------------------------------------------------------------
// Ray-shooting configuration shared by all kernels (single plane: beta = theta - scaling * sum alpha,
// single_plane.py:38-56,112-139; multi plane: comoving recursion, multi_plane_base.py:159-233).
struct RayCfg {
  int32_t multi_plane;
  double alpha_scaling, T_ij_end, inv_T_z_source;
  MathK mk;  // polynomial coefficients of log2 / exp2 as kernel parameters (uniform-register operands)
};

// State of J rays while they are pushed through the deflector list.
template <int J>
struct Rays {
  double x[J], y[J];    // single plane: image position; multi plane: comoving position
  double ax[J], ay[J];  // single plane: accumulated deflection; multi plane: current angle
};

LSIM_HD constexpr bool lens_type_is_series(int type) {
  return type == LSIM_LENS_EPL || type == LSIM_LENS_EPL_MULTIPOLE_M1M3M4;
}

// Compile-time profile list (single plane), first half: the position-independent terms of the cheap profiles
// (per-set constants: no register lives across the series), then the series profiles, whose last FMA accumulates.
template <bool FULL, int J, class LT>
LSIM_HD void ray_apply_head(const RayCfg& P, const LensC* L, const double* pool, Rays<J>& r) {
  double kx = 0.0, ky = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int l = 0; l < LT::N; ++l)
    if (LT::at(l) == LSIM_LENS_SHEAR) {
      kx += L[l].c[4];
      ky += L[l].c[5];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    r.ax[j] += kx;
    r.ay[j] += ky;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int l = 0; l < LT::N; ++l)
    if (lens_type_is_series(LT::at(l)))
      lens_accum<FULL, J, true, false, MathK>(LT::at(l), L[l], pool, r.x, r.y, 1.0, r.ax, r.ay, P.mk);
}

// ... second half: the position-dependent part of the cheap profiles.  r.x / r.y may have been re-read in between
// (the image kernel does, instead of keeping them in registers across the series).  A sum of two terms does not
// depend on the order; beyond that the order moves the last bit only.
template <bool FULL, int J, class LT>
LSIM_HD void ray_apply_tail(const RayCfg& P, const LensC* L, const double* pool, Rays<J>& r) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int l = 0; l < LT::N; ++l)
    if (!lens_type_is_series(LT::at(l)))
      lens_accum<FULL, J, true, true, MathK>(LT::at(l), L[l], pool, r.x, r.y, 1.0, r.ax, r.ay, P.mk);
}

template <bool FULL, int J, class LT = RuntimeTypes>
LSIM_HD void ray_apply(const RayCfg& P, const LensC* L, const double* pool, int n_c, Rays<J>& r) {
  if (LT::N >= 0) {  // single plane, profile list known at compile time: fully unrolled, no dispatch
    ray_apply_head<FULL, J, LT>(P, L, pool, r);
    ray_apply_tail<FULL, J, LT>(P, L, pool, r);
    return;
  }
  if (!P.multi_plane) {
    for (int l = 0; l < n_c; ++l)
      lens_accum<FULL, J, true, false, MathK>(L[l].type, L[l], pool, r.x, r.y, 1.0, r.ax, r.ay, P.mk);
  } else {
    // MultiPlaneBase.ray_shooting_partial_comoving (multi_plane_base.py:195-233)
    double tx[J], ty[J];
    for (int l = 0; l < n_c; ++l) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < J; ++j) {
        r.x[j] += r.ax[j] * L[l].delta_T;
        r.y[j] += r.ay[j] * L[l].delta_T;
        tx[j] = r.x[j] * L[l].inv_T_z;
        ty[j] = r.y[j] * L[l].inv_T_z;
      }
      lens_accum<FULL, J, false, false, MathK>(L[l].type, L[l], pool, tx, ty, -L[l].red2phys, r.ax, r.ay, P.mk);
    }
  }
}

// SINGLE: single-plane lensing known at compile time (the specialised kernels)
template <int J, bool SINGLE = false>
LSIM_HD void ray_init(const RayCfg& P, int j, double x, double y, Rays<J>& r) {
  if (SINGLE || !P.multi_plane) {
    r.x[j] = x;
    r.y[j] = y;
    r.ax[j] = 0;
    r.ay[j] = 0;
  } else {
    r.x[j] = 0;
    r.y[j] = 0;
    r.ax[j] = x;
    r.ay[j] = y;
  }
}

// -> source-plane position
template <int J, bool SINGLE = false>
LSIM_HD void ray_finish(const RayCfg& P, const Rays<J>& r, int j, double& bx, double& by) {
  if (SINGLE || !P.multi_plane) {
    bx = r.x[j] - r.ax[j] * P.alpha_scaling;
    by = r.y[j] - r.ay[j] * P.alpha_scaling;
  } else {
    bx = (r.x[j] + r.ax[j] * P.T_ij_end) * P.inv_T_z_source;
    by = (r.y[j] + r.ay[j] * P.T_ij_end) * P.inv_T_z_source;
  }
}

// ---- light -----------------------------------------------------------------------------------------------
// c[0..1] centre; SERSIC family: c[2..4] = product-average transform rows / norm, c[5] = 1/(2n),
// c[6] = -log2(Rs_)/n, c[7] = -bn log2(e), c[8] = bn log2(e) (c[5..8] times kExp2Scale: exp2_*_s), c[9] = (max_R_frac Rs_)^2,
// c[10] = amp, c[11..12] = the transform of -centre
LSIM_HD void light_prepare(const lsim_light_desc& d, const double* row, LightC& G) {
  G.type = d.type;
  G.lensed = d.lensed;
  G.slot = d.slot;
  G.pad = 0;
  auto P = [&](int i) { return param_of(d, row, i); };
  double amp = d.use_amp ? P(0) : 1.0;
  switch (d.type) {
    case LSIM_LIGHT_SERSIC:
    case LSIM_LIGHT_SERSIC_ELLIPSE: {
      double R_sersic = P(1), n = P(2), e1 = 0, e2 = 0, cx, cy;
      if (d.type == LSIM_LIGHT_SERSIC_ELLIPSE) {
        R_sersic = fmax(0.0, R_sersic);  // sersic_ellipse.py:90
        e1 = P(3);
        e2 = P(4);
        cx = P(5);
        cy = P(6);
      } else {
        cx = P(3);
        cy = P(4);
      }
      double norm = fmax(sqrt(fabs(1 - e1 * e1 - e2 * e2)), 0.000001);
      double Rs_ = fmax(0.0001, R_sersic);
      double bn = fmax(1.9992 * n - 0.3271, 0.00001);
      G.c[0] = cx;
      G.c[1] = cy;
      G.c[2] = (1 - e1) / norm;
      G.c[3] = -e2 / norm;
      G.c[4] = (1 + e1) / norm;
      G.c[5] = 0.5 / n;
      G.c[6] = -log2(Rs_) / n;
      G.c[7] = -bn * kLog2e;
      G.c[8] = fmin(bn * kLog2e, 1021.0);  // exp(bn) overflows beyond (n_sersic > 354): saturate
      G.c[9] = (1000.0 * Rs_) * (1000.0 * Rs_);
      G.c[10] = amp;
      G.c[11] = -(G.c[2] * cx + G.c[3] * cy);  // shift folded into the product-average transform
      G.c[12] = -(G.c[3] * cx + G.c[4] * cy);
      // exponents pre-scaled for exp2_*_s
      G.c[5] *= kExp2Scale;
      G.c[6] *= kExp2Scale;
      G.c[7] *= kExp2Scale;
      G.c[8] *= kExp2Scale;
      G.type = LSIM_LIGHT_SERSIC_ELLIPSE;
      break;
    }
    case LSIM_LIGHT_GAUSSIAN: {
      double sigma = P(1);
      G.c[0] = P(2);
      G.c[1] = P(3);
      G.c[2] = -0.5 * kLog2e / (sigma * sigma);
      G.c[3] = amp / (2 * kPi * sigma * sigma);
      break;
    }
    default:
      G.slot = -1;
      break;
  }
}

// J points at once (sersic_utils.py:247-267 with R^2 in place of R: no square root is needed because
// (R/Rs)^(1/n) = exp2(log2(R^2)/(2n) - log2(Rs)/n)).  -> e[j] = the profile at amp = 1
// SERSIC(_ELLIPSE) at amp = 1 from the product-average coordinates (xt, yt) of J points
template <int J, class KT>
LSIM_HD void sersic_core(const LightC& G, const double (&xt)[J], const double (&yt)[J], double (&e)[J], const KT& K) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    const double R2c = max_nan(fma(yt[j], yt[j], xt[j] * xt[j]), 1e-8);  // R_ = max(1e-4, R)
    const double k = exp2_fin_s(fma(G.c[5], log2_pos(R2c, K), G.c[6]), K);
    double v = exp2_fin_s(fma(G.c[7], k, G.c[8]), K);  // k >= 0, c[7] <= 0: the argument is at most c[8] <= kExp2Scale * 1021
    // R_/Rs_ <= max_R_frac, and nan_to_num (the exponent is bounded above, so no +inf): one select for both
    e[j] = (R2c <= G.c[9] && v == v) ? v : 0.0;
  }
}

template <int J, class KT>
LSIM_HD void light_unit(int type, const LightC& G, const double (&x)[J], const double (&y)[J], double (&e)[J],
                        const KT& K) {
  if (type == LSIM_LIGHT_SERSIC_ELLIPSE) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < J; ++j) {
      const double xt = fma(G.c[2], x[j], fma(G.c[3], y[j], G.c[11]));
      const double yt = fma(G.c[3], x[j], fma(G.c[4], y[j], G.c[12]));
      const double R2c = max_nan(fma(yt, yt, xt * xt), 1e-8);  // R_ = max(1e-4, R)
      const double k = exp2_fin_s(fma(G.c[5], log2_pos(R2c, K), G.c[6]), K);
      double v = exp2_fin_s(fma(G.c[7], k, G.c[8]), K);  // k >= 0, c[7] <= 0: the argument is at most c[8] <= kExp2Scale * 1021
      v = (R2c <= G.c[9]) ? v : 0.0;  // R_/Rs_ <= max_R_frac
      e[j] = (v == v) ? v : 0.0;      // nan_to_num (the exponent is bounded above, so no +inf)
    }
    return;
  }
  // GAUSSIAN (amplitude normalisation c[3] applied by the callers through G.c[10] ... no: c[3] IS its amplitude)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) {
    const double xs = x[j] - G.c[0], ys = y[j] - G.c[1];
    e[j] = exp2_fin(G.c[2] * fma(ys, ys, xs * xs));
  }
}

LSIM_HD double light_amp(int type, const LightC& G) { return type == LSIM_LIGHT_SERSIC_ELLIPSE ? G.c[10] : G.c[3]; }

template <int J>
LSIM_HD void light_eval(int type, const LightC& G, const double (&x)[J], const double (&y)[J], double (&v)[J]) {
  light_unit<J>(type, G, x, y, v, TableK());
  const double amp = light_amp(type, G);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) v[j] = amp * v[j];
}

// acc += sum_j amp * profile(x_j, y_j): the amplitude rides on the accumulation FMA
template <int J, class KT>
LSIM_HD void light_accum(int type, const LightC& G, const double (&x)[J], const double (&y)[J], double& acc,
                         const KT& K) {
  double e[J];
  light_unit<J>(type, G, x, y, e, K);
  const double amp = light_amp(type, G);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < J; ++j) acc = fma(amp, e[j], acc);
}

LSIM_HD double light_eval(const LightC& G, double x, double y) {
  const double x1[1] = {x}, y1[1] = {y};
  double v1[1];
  light_eval<1>(G.type, G, x1, y1, v1);
  return v1[0];
}

}  // namespace lsim

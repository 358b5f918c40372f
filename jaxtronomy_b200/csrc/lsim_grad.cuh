// lsim_grad.cuh -- forward-mode (dual number) arithmetic of the profiles, for the gradient of the image likelihood
// with respect to the sampled parameters (what jax.value_and_grad(LikelihoodModule.logL) gives the
// reference's gradient-descent sampler, /root/reference/jaxtronomy/Sampling/Samplers/optax.py:39,91-96).
// Host + device, like lsim_profiles.cuh: tests/hostmath compiles the same code with g++.
//
// Structure of the derivative (lsim_api.cu: lsim_logl_grad):
//   logL = L(C),  C = scale * psf (*) U(p)            forward path (k_image, k_conv_*, k_likelihood), unchanged
//   Abar = dL/dC                                      k_like_adjoint (WLS solve differentiated analytically)
//   Ubar = scale * psf^T (*) Abar                     the same convolution kernels with the flipped PSF
//   grad = sum_pixels Ubar dU/dp                      k_image_grad: per sub-pixel, every profile is evaluated ONCE with a
//                                                     dual number carrying the derivatives w.r.t. ITS OWN parameters
// The chain through the lens equation is closed per sub-pixel: the light profiles are evaluated on duals seeded with
// d/d(beta_x), d/d(beta_y), which gives v = sum_k Ubar_k grad_beta I_k; a lens parameter then contributes
// -alpha_scaling * v . d(alpha)/dp.  No tangent image is ever stored: HBM traffic is one read of Ubar.
//
// Reference semantics kept:
//   * every clamp differentiates like jnp.maximum / jnp.minimum / jnp.where (tangent of the selected branch),
//     nan_to_num passes the tangent where the value is finite and gives 0 where it was replaced;
//   * EPL: the value of 2F1 uses the nmax schedule, its TANGENT the fixed 100-term series in (t, z) only -- the
//     reference's custom_jvp (LensModel/Profiles/epl.py:321-340, hyp2f1_slow = nmax 100, epl.py:283).
// Scope: single plane; EPL, SIE (both NIE settings), NIE, SIS, SHEAR, CONVERGENCE; SERSIC, SERSIC_ELLIPSE, GAUSSIAN.
#pragma once

#include "lsim_profiles.cuh"

#if defined(__CUDACC__)
#define LSIM_UNROLL _Pragma("unroll")
#else
#define LSIM_UNROLL
#endif

namespace lsim {
namespace ad {

constexpr int kLensN = 6;    // tangents carried by a lens profile: its own parameters (param_names order)
constexpr int kLightN = 9;   // light profile: d/d(x), d/d(y) of the evaluation point, then its own parameters
constexpr int kLensEntries = 8;  // gradient entries per (pseudo-)profile: 6 tangents + up to 2 closed-form extras
constexpr int kGradLensMP = 64;  // internal pseudo-profile of the gradient path: ONE multipole (m = 1, 3 or 4) of
                                 // EPL_MULTIPOLE_M1M3M4 with the parameters (theta_E, e1, e2, center_x, center_y,
                                 // delta_phi_m, a_m/a): lsim_api.cu expands that 12-parameter profile into an EPL and
                                 // three of these, whose gradients land in the parent's columns
constexpr int kTanTerms = 100;                      // hyp2f1_slow, epl.py:283
// per EPL-family profile: F_n [kMaxSeries] | dF_n/dt [kTanTerms] | Coef4 [kTanTerms] (the four series of the fast path,
// interleaved so that one term is two 128-bit shared-memory loads)
constexpr int kGradPool = kMaxSeries + 5 * kTanTerms;
struct alignas(16) Coef4 {
  double fv;  // F_n for n < n_val, else 0: the value series (nmax schedule) on the tangent series' index range
  double f;   // F_n: the 100-term S of the tangent
  double g;   // (n+1) F_{n+1}: dS/dw
  double d;   // dF_n/dt: dS/dt
};

template <int N>
struct Dual {
  double v;
  double d[N];
};

template <int N>
LSIM_HD Dual<N> cst(double v) {
  Dual<N> r;
  r.v = v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = 0.0;
  return r;
}
// independent variable number i (i < 0: a constant)
template <int N>
LSIM_HD Dual<N> var(double v, int i) {
  Dual<N> r;
  r.v = v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = (k == i) ? 1.0 : 0.0;
  return r;
}

template <int N>
LSIM_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <int N>
LSIM_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <int N>
LSIM_HD Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = -a.d[k];
  return r;
}
template <int N>
LSIM_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = fma(a.v, b.d[k], a.d[k] * b.v);
  return r;
}
template <int N>
LSIM_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = fma(-r.v, b.d[k], a.d[k]) * inv;
  return r;
}
template <int N>
LSIM_HD Dual<N> operator+(const Dual<N>& a, double b) {
  Dual<N> r = a;
  r.v += b;
  return r;
}
template <int N>
LSIM_HD Dual<N> operator+(double b, const Dual<N>& a) { return a + b; }
template <int N>
LSIM_HD Dual<N> operator-(const Dual<N>& a, double b) {
  Dual<N> r = a;
  r.v -= b;
  return r;
}
template <int N>
LSIM_HD Dual<N> operator-(double b, const Dual<N>& a) {
  Dual<N> r;
  r.v = b - a.v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = -a.d[k];
  return r;
}
template <int N>
LSIM_HD Dual<N> operator*(const Dual<N>& a, double b) {
  Dual<N> r;
  r.v = a.v * b;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <int N>
LSIM_HD Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N>
LSIM_HD Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N>
LSIM_HD Dual<N> operator/(double b, const Dual<N>& a) {
  Dual<N> r;
  const double inv = 1.0 / a.v;
  r.v = b * inv;
  const double s = -r.v * inv;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * s;
  return r;
}

// r = f(a) with f'(a) = s
template <int N>
LSIM_HD Dual<N> chain(const Dual<N>& a, double v, double s) {
  Dual<N> r;
  r.v = v;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * s;
  return r;
}
template <int N>
LSIM_HD Dual<N> dsqrt(const Dual<N>& a) {
  const double s = sqrt(a.v);
  return chain(a, s, a.v > 0 ? 0.5 / s : 0.0);  // the tangent at 0 is taken as 0 instead of inf * 0
}
template <int N>
LSIM_HD Dual<N> dexp(const Dual<N>& a) {
  const double e = exp(a.v);
  return chain(a, e, e);
}
template <int N>
LSIM_HD Dual<N> dlog(const Dual<N>& a) { return chain(a, log(a.v), 1.0 / a.v); }
template <int N>
LSIM_HD Dual<N> dsin(const Dual<N>& a) { return chain(a, sin(a.v), cos(a.v)); }
template <int N>
LSIM_HD Dual<N> dcos(const Dual<N>& a) { return chain(a, cos(a.v), -sin(a.v)); }
template <int N>
LSIM_HD Dual<N> datan(const Dual<N>& a) { return chain(a, atan(a.v), 1.0 / fma(a.v, a.v, 1.0)); }
template <int N>
LSIM_HD Dual<N> datanh(const Dual<N>& a) { return chain(a, atanh(a.v), 1.0 / fma(-a.v, a.v, 1.0)); }
template <int N>
LSIM_HD Dual<N> dacosh(const Dual<N>& a) {
  const double r = sqrt(fma(a.v, a.v, -1.0));
  return chain(a, log(a.v + r), 1.0 / r);
}
template <int N>
LSIM_HD Dual<N> dacos(const Dual<N>& a) { return chain(a, acos(a.v), -1.0 / sqrt(fma(-a.v, a.v, 1.0))); }
// a^b for a > 0
template <int N>
LSIM_HD Dual<N> dpow(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = pow(a.v, b.v);
  const double sa = r.v * b.v / a.v, sb = r.v * log(a.v);
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = fma(sa, a.d[k], sb * b.d[k]);
  return r;
}
template <int N>
LSIM_HD Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r;
  r.v = atan2(y.v, x.v);
  const double n2 = fma(x.v, x.v, y.v * y.v);
  const double inv = n2 > 0 ? 1.0 / n2 : 0.0;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = fma(x.v, y.d[k], -(y.v * x.d[k])) * inv;
  return r;
}
// jnp.maximum(a, c) / jnp.minimum(a, c) with a constant: the tangent of the selected argument
template <int N>
LSIM_HD Dual<N> dmax(const Dual<N>& a, double c) { return (a.v < c) ? cst<N>(c) : a; }
template <int N>
LSIM_HD Dual<N> dmin(const Dual<N>& a, double c) { return (a.v > c) ? cst<N>(c) : a; }
template <int N>
LSIM_HD Dual<N> dabs(const Dual<N>& a) { return a.v < 0 ? -a : a; }
template <int N>
LSIM_HD Dual<N> dnan_to_num(const Dual<N>& a, double p, double n) {
  if (a.v != a.v) return cst<N>(0.0);
  if (a.v > DBL_MAX) return cst<N>(p);
  if (a.v < -DBL_MAX) return cst<N>(n);
  Dual<N> r = a;
  LSIM_UNROLL
  for (int k = 0; k < N; ++k) r.d[k] = (r.d[k] == r.d[k]) ? r.d[k] : 0.0;
  return r;
}

// ---- prepared (per parameter set) lens constants ---------------------------------------------------------------
using LD = Dual<kLensN>;
struct GLens {
  int32_t type;      // evaluation kind: LSIM_LENS_EPL / NIE / SIS / SHEAR / CONVERGENCE / NULL
  int32_t n_val;     // EPL: terms of the value series (nmax schedule), n_tan: terms of the tangent series (<= 100)
  int32_t n_tan;
  int32_t coef_off;  // EPL: offset of this profile's coefficients in the pool
  LD c[8];           // dual-number evaluation (lens_eval_d)
  LD p[6];           // EPL hand-derived path: theta_E, t, phi, q, center_x, center_y with their parameter derivatives
  double k[12];      // EPL hand-derived path: cos phi, sin phi, f, ln b, t log2 b, 2/(1+q), dlnA/dq, df/dq
};

template <int N>
LSIM_HD void ellipticity2phi_q_d(const Dual<N>& e1, const Dual<N>& e2, Dual<N>& phi, Dual<N>& q) {
  phi = datan2(e2, e1) * 0.5;
  Dual<N> c = dmin(dsqrt(e1 * e1 + e2 * e2), 0.9999);
  q = (1.0 - c) / (1.0 + c);
}

// tangent-carrying parameter i of a lens descriptor: an independent variable if it is a column of the parameter
// matrix, a constant if it is fixed
LSIM_HD LD lens_var(const lsim_lens_desc& d, const double* row, int i) {
  return var<kLensN>(param_of(d, row, i), d.col[i] >= 0 ? i : -1);
}

// EPL family.  c[0..6] = cx, cy, cos(phi), sin(phi), q, b, t.  F[n], n < max(n_val, n_tan): series coefficients
// (lsim_profiles.cuh: epl_prepare); dF[n] = dF_n/dt, n < n_tan.  Terms whose bound n wmax^(n-1) is below 1e-22 of the
// leading term cannot change a sum in fp64 and are not kept.
LSIM_HD void epl_prepare_d(GLens& L, double* F, double* dF, const LD& theta_E, const LD& gamma, const LD& e1,
                           const LD& e2, const LD& cx, const LD& cy) {
  LD phi, q;
  ellipticity2phi_q_d(e1, e2, phi, q);
  L.type = LSIM_LENS_EPL;
  L.c[0] = cx;
  L.c[1] = cy;
  L.c[2] = dcos(phi);
  L.c[3] = dsin(phi);
  L.c[4] = q;
  L.c[5] = theta_E * dsqrt(q);
  L.c[6] = gamma - 1.0;
  const double f = (1.0 - q.v) / (1.0 + q.v);
  {
    const double t = gamma.v - 1.0, b = L.c[5].v;
    L.p[0] = theta_E;
    L.p[1] = gamma - 1.0;
    L.p[2] = phi;
    L.p[3] = q;
    L.p[4] = cx;
    L.p[5] = cy;
    L.k[0] = L.c[2].v;
    L.k[1] = L.c[3].v;
    L.k[2] = f;
    L.k[3] = log(b);
    L.k[4] = t * log2(b);
    L.k[5] = 2.0 / (1.0 + q.v);
    L.k[6] = -1.0 / (1.0 + q.v) + t / (2.0 * q.v);      // d ln A / d q at fixed theta_E, A = 2 b^t / (1+q), b = theta_E sqrt(q)
    L.k[7] = -2.0 / ((1.0 + q.v) * (1.0 + q.v));        // d f / d q
  }
  const double wmax = f / (2.0 - f);
  const int nmax = epl_nmax(f);
  int cut = 2;  // first n with n wmax^(n-1) < 1e-22 (beyond it nothing contributes)
  {
    double wn = 1.0;
    while (cut < kMaxSeries && cut * wn >= 1e-22) {
      wn *= wmax;
      ++cut;
    }
  }
  const int n_val = nmax == 0 ? 1 : (nmax < cut ? nmax : cut);
  const int n_tan = kTanTerms < cut ? kTanTerms : cut;
  L.n_val = n_val;
  L.n_tan = n_tan;
  const int n_all = n_val > n_tan ? n_val : n_tan;
  // F_n and dF_n/dt by the reference's recurrence on Dual<1> (B = t/2, C = 2 - B; hyp2f1_util.py:152-176, a = 1)
  using D1 = Dual<1>;
  const D1 t = var<1>(L.c[6].v, 0);
  const D1 B = t * 0.5, C = 2.0 - B;
  D1 ppF = cst<1>(1.0), pF = 1.0 - 2.0 * B / C;
  F[0] = 1.0;
  dF[0] = 0.0;
  if (n_all > 1) F[1] = pF.v;
  if (n_tan > 1) dF[1] = pF.d[0];
  const D1 BC = 2.0 * B - C;
  for (int n = 2; n < n_all; ++n) {
    const D1 nF = ((n - 1.0) * ppF - BC * pF) / (C + (n - 1.0));
    F[n] = nF.v;
    if (n < n_tan) dF[n] = nF.d[0];
    ppF = pF;
    pF = nF;
  }
  // the fast path's interleaved table: value series cut at n_val, S, dS/dw = sum_m (m+1) F_{m+1} w^m, dS/dt
  Coef4* T = reinterpret_cast<Coef4*>(dF + kTanTerms);
  for (int n = 0; n < n_tan; ++n) {
    T[n].fv = n < n_val ? F[n] : 0.0;
    T[n].f = F[n];
    T[n].g = n + 1 < n_tan ? (n + 1.0) * F[n + 1] : 0.0;
    T[n].d = dF[n];
  }
}

// EPL deflection at one image-plane point with its derivatives w.r.t. the profile's own parameters
// (epl.py:178-200,366-394 on duals; the hypergeometric function as described in the header)
LSIM_HD void epl_eval_d(const GLens& L, const double* F, const double* dF, double x, double y, LD& fx, LD& fy) {
  const LD &cph = L.c[2], &sph = L.c[3], &q = L.c[4], &b = L.c[5], &t = L.c[6];
  const LD xs = x - L.c[0], ys = y - L.c[1];
  const LD xr = xs * cph + ys * sph, yr = ys * cph - xs * sph;
  const LD Zr = q * xr, Zi = yr;
  const LD R2 = Zr * Zr + Zi * Zi;
  const LD R = dmax(dsqrt(R2), 0.000000001);
  const LD f = (1.0 - q) / (1.0 + q);
  // z = -f Z / conj(Z) = -f Z^2 / |Z|^2
  const LD g = -(f / R2);
  const LD zr = g * (Zr * Zr - Zi * Zi), zi = g * (2.0 * (Zr * Zi));
  // plain complex from here: w = z / (z - 2), u = 1 / (1 - z/2) = -2 / (z - 2)
  const double er = zr.v - 2.0, ei = zi.v;
  const double einv = 1.0 / fma(er, er, ei * ei);
  const double ir = er * einv, ii = -ei * einv;  // 1 / (z - 2)
  const double wr = fma(zr.v, ir, -(zi.v * ii)), wi = fma(zr.v, ii, zi.v * ir);
  const double ur = -2.0 * ir, ui = -2.0 * ii;
  // value series (n_val terms)
  double sr = F[L.n_val - 1], si = 0.0;
  for (int k = L.n_val - 2; k >= 0; --k) {
    const double nr = fma(sr, wr, -(si * wi)) + F[k];
    si = fma(sr, wi, si * wr);
    sr = nr;
  }
  // tangent series (n_tan terms): S = sum F_n w^n, S' = dS/dw, St = sum dF_n w^n
  double pr = F[L.n_tan - 1], pi = 0.0, dr = 0.0, di = 0.0, tr = dF[L.n_tan - 1], ti = 0.0;
  for (int k = L.n_tan - 2; k >= 0; --k) {
    double nr = fma(dr, wr, -(di * wi)) + pr;
    di = fma(dr, wi, di * wr) + pi;
    dr = nr;
    nr = fma(pr, wr, -(pi * wi)) + F[k];
    pi = fma(pr, wi, pi * wr);
    pr = nr;
    nr = fma(tr, wr, -(ti * wi)) + dF[k];
    ti = fma(tr, wi, ti * wr);
    tr = nr;
  }
  // h = u S;  dh/dt = u St;  dh/dz = u^2 S / 2 + u S' dw/dz,  dw/dz = -2 / (z - 2)^2 = -u^2 / 2
  const double u2r = fma(ur, ur, -(ui * ui)), u2i = 2.0 * ur * ui;
  const double htr = fma(ur, tr, -(ui * ti)), hti = fma(ur, ti, ui * tr);
  const double usr = fma(ur, dr, -(ui * di)), usi = fma(ur, di, ui * dr);  // u S'
  const double ar_ = pr - usr, ai_ = pi - usi;                             // S - u S'
  const double hzr = 0.5 * fma(u2r, ar_, -(u2i * ai_)), hzi = 0.5 * fma(u2r, ai_, u2i * ar_);
  LD hr, hi;
  hr.v = fma(ur, sr, -(ui * si));
  hi.v = fma(ur, si, ui * sr);
  LSIM_UNROLL
  for (int k = 0; k < kLensN; ++k) {
    hr.d[k] = fma(htr, t.d[k], fma(hzr, zr.d[k], -(hzi * zi.d[k])));
    hi.d[k] = fma(hti, t.d[k], fma(hzr, zi.d[k], hzi * zr.d[k]));
  }
  // alpha = 2/(1+q) (b/R)^t Z h
  const LD Ror = Zr * hr - Zi * hi, Roi = Zr * hi + Zi * hr;
  const LD pref = (2.0 / (1.0 + q)) * dpow(b / R, t);
  const LD ar = dnan_to_num(pref * Ror, 1e10, -1e10), ai = dnan_to_num(pref * Roi, 1e10, -1e10);
  fx = ar * cph - ai * sph;  // rotate(f, -phi)
  fy = ar * sph + ai * cph;
}

// NIE: c[0..6] = cx, cy, cos(phi), sin(phi), q (clamped), b, s   (nie.py:85-107,181-200,280-294,333-342)
LSIM_HD void nie_prepare_d(GLens& L, const LD& theta_E, const LD& e1, const LD& e2, const LD& s_scale, const LD& cx,
                           const LD& cy) {
  LD phi, q;
  ellipticity2phi_q_d(e1, e2, phi, q);
  const LD theta_E_conv = theta_E / dsqrt((1.0 + q * q) / (2.0 * q));
  L.type = LSIM_LENS_NIE;
  L.c[0] = cx;
  L.c[1] = cy;
  L.c[2] = dcos(phi);
  L.c[3] = dsin(phi);
  L.c[4] = (q.v >= 1) ? cst<kLensN>(0.99999999) : q;
  L.c[5] = theta_E_conv * dsqrt((1.0 + q * q) * 0.5);
  L.c[6] = s_scale / dsqrt(q);
}

LSIM_HD void nie_eval_d(const GLens& L, double x, double y, LD& fx, LD& fy) {
  const LD &cph = L.c[2], &sph = L.c[3], &q = L.c[4], &b = L.c[5], &s = L.c[6];
  const LD xs = x - L.c[0], ys = y - L.c[1];
  const LD xr = xs * cph + ys * sph, yr = ys * cph - xs * sph;
  const LD q2 = q * q;
  const LD psi = dsqrt(q2 * (s * s + xr * xr) + yr * yr);
  const LD e = dsqrt(1.0 - q2);
  const LD be = b / e;
  const LD ar = be * datan(e * xr / (psi + s));
  const LD ai = be * datanh(e * yr / (psi + q2 * s));
  fx = ar * cph - ai * sph;
  fy = ar * sph + ai * cph;
}

struct GLens;
LSIM_HD void mp_constants(GLens& L);
// One multipole of EPL_MULTIPOLE_M1M3M4 (epl_multipole_m1m3m4.py:463-535: a_m = a_m/a * theta_E / sqrt(q),
// phi_m = phi + delta_phi_m, r_E = theta_E for m = 1) at UNIT a_m/a: the deflection is linear in that amplitude, so its
// derivative is the unit deflection itself and the six tangents go to theta_E, e1, e2, center_x, center_y, delta_phi_m.
// c[0..4] = cx, cy, theta_E/sqrt(q), phi_m, theta_E;  k[0] = a_m/a;  n_val = m.
LSIM_HD void mp_prepare_d(GLens& L, int m, const LD& theta_E, const LD& e1, const LD& e2, const LD& cx, const LD& cy,
                          const LD& delta, double a) {
  LD phi, q;
  ellipticity2phi_q_d(e1, e2, phi, q);
  L.type = kGradLensMP;
  L.n_val = m;
  L.c[0] = cx;
  L.c[1] = cy;
  L.c[2] = theta_E / dsqrt(q);
  L.c[3] = phi + delta;
  L.c[4] = theta_E;
  L.k[0] = a;
  mp_constants(L);
}

// Multipole.derivatives (multipole.py:72-119) with a_m = c[2]
LSIM_HD void mp_eval_d(const GLens& L, double x, double y, LD& fx, LD& fy) {
  const int m = L.n_val;
  const LD xs = x - L.c[0], ys = y - L.c[1];
  LD r = dsqrt(xs * xs + ys * ys);
  const LD phi = datan2(ys, xs);
  const LD &am = L.c[2], &phim = L.c[3];
  const LD cphi = dcos(phi), sphi = dsin(phi);
  if (m == 1) {
    r = dmax(r, 0.000001);
    const LD lg = dlog(r / L.c[4]);
    const LD cd = dcos(phi - phim);
    fx = am * 0.5 * (dcos(phim) * lg + cd * cphi);
    fy = am * 0.5 * (dsin(phim) * lg + cd * sphi);
  } else {
    const LD A = am / (1.0 - (double)(m * m));
    const LD arg = (phi - phim) * (double)m;
    const LD cm = dcos(arg), sm = dsin(arg);
    fx = cphi * A * cm + sphi * (double)m * A * sm;
    fy = sphi * A * cm - cphi * (double)m * A * sm;
  }
}

// ---- NFW / TNFW / MULTIPOLE / CSE family on dual numbers (literal restatements; nfw.py:88-105,259-279,347-365,
// tnfw.py:69-87,228-349, multipole.py:72-119, cored_steep_ellipsoid.py:107-132,227-245,434-477,536-564,
// nfw_ellipse_cse.py:250-280,338-349, hernquist_ellipse_cse.py:151-164,190-202) --------------------------------------
LSIM_HD LD nfw_g_d(LD X) {
  if (X.v < 0.000001) X = cst<kLensN>(0.000001);
  if (X.v < 1) return dlog(X * 0.5) + dacosh(1.0 / X) / dsqrt(1.0 - X * X);
  if (X.v == 1) return cst<kLensN>(1 + log(0.5));
  return dlog(X * 0.5) + dacos(1.0 / X) / dsqrt(X * X - 1.0);
}
LSIM_HD LD alpha2rho0_d(const LD& alpha_Rs, const LD& Rs) { return alpha_Rs / (4.0 * (Rs * Rs) * (1.0 + log(0.5))); }

// c[0..3] = Rs (clamped), 4 rho0 Rs, cx, cy
LSIM_HD void nfw_eval_d(const GLens& L, double x, double y, LD& fx, LD& fy) {
  const LD xs = x - L.c[2], ys = y - L.c[3];
  const LD R = dmax(dsqrt(xs * xs + ys * ys), 0.00000001);
  const LD X = R / L.c[0];
  const LD a = L.c[1] * nfw_g_d(X) / (X * X);
  fx = a * xs;
  fy = a * ys;
}

LSIM_HD LD tnfw_F_d(LD x) {
  x = dmax(x, 0.001);
  if (x.v < 1) {
    const LD r = dsqrt(1.0 - x * x);
    return datanh(r) / r;
  }
  if (x.v > 1) {
    const LD r = dsqrt(x * x - 1.0);
    return datan(r) / r;
  }
  return x;
}
// c[0..4] = Rs, 4 rho0 Rs, tau, cx, cy
LSIM_HD void tnfw_eval_d(const GLens& L, double x, double y, LD& fx, LD& fy) {
  const LD &Rs = L.c[0], &tau = L.c[2];
  const LD xs = x - L.c[3], ys = y - L.c[4];
  LD R = dsqrt(xs * xs + ys * ys);
  if (R.v < 0.001 * Rs.v) R = Rs * 0.001;  // jnp.maximum(R, s * Rs): the tangent of the larger argument
  const LD X = dmax(R / Rs, 0.001);
  const LD t2 = tau * tau;
  const LD st = dsqrt(t2 + X * X);
  const LD Lx = dlog(X / (tau + st));
  const LD g = t2 / ((t2 + 1.0) * (t2 + 1.0)) *
               ((t2 + 1.0 + 2.0 * (X * X - 1.0)) * tnfw_F_d(X) + tau * kPi + (t2 - 1.0) * dlog(tau) +
                st * (Lx * (t2 - 1.0) / tau - kPi));
  const LD a = L.c[1] * g / (X * X);
  fx = a * xs;
  fy = a * ys;
}

// MULTIPOLE: c[0..4] = cx, cy, a_m, phi_m, r_E; n_val = m (not differentiable)
// (shares mp_eval_d with the pseudo-profile below: unit scale k[0] = 1)

// CSE major axis / product average at (x, y) in the major-axis frame, generic in the tangent width
template <int N>
LSIM_HD void cse_major_d(const Dual<N>& x, const Dual<N>& y, double a, double s, const Dual<N>& q, Dual<N>& fx,
                         Dual<N>& fy) {
  const Dual<N> q2 = q * q;
  const Dual<N> psi = dsqrt(q2 * (x * x + s * s) + y * y);
  const Dual<N> ps = psi + s;
  const Dual<N> Phi = ps * ps + (1.0 - q2) * (x * x);
  const Dual<N> den = a / (s * psi * Phi);
  fx = q * x * (psi + q2 * s) * den;
  fy = q * y * ps * den;
}
template <int N>
LSIM_HD void cse_prodavg_d(const Dual<N>& x, const Dual<N>& y, double a, double s, const Dual<N>& q, Dual<N>& fx,
                           Dual<N>& fy) {
  const Dual<N> sq = dsqrt(q);
  Dual<N> gx, gy;
  cse_major_d(x * sq, y * sq, 1.0, s, q, gx, gy);
  const Dual<N> f = (a / q) * sq;
  fx = f * gx;
  fy = f * gy;
}

// Sum over a CSE table at (X, Y, q) on Dual<3> (d/dX, d/dY, d/dq), then composed with the six-wide duals of X, Y, q:
// the 44 / 41 / 16 components are evaluated with 3 tangents instead of 6.
LSIM_HD void cse_set_d(bool product_avg, const double* A, const double* S, int n, const LD& X, const LD& Y, const LD& q,
                       LD& fx, LD& fy) {
  using D3 = Dual<3>;
  const D3 x3 = var<3>(X.v, 0), y3 = var<3>(Y.v, 1), q3 = var<3>(q.v, 2);
  D3 sx = cst<3>(0.0), sy = cst<3>(0.0);
  for (int i = 0; i < n; ++i) {
    D3 gx, gy;
    if (product_avg)
      cse_prodavg_d(x3, y3, A[i], S[i], q3, gx, gy);
    else
      cse_major_d(x3, y3, A[i], S[i], q3, gx, gy);
    sx = sx + gx;
    sy = sy + gy;
  }
  fx.v = sx.v;
  fy.v = sy.v;
  LSIM_UNROLL
  for (int k = 0; k < kLensN; ++k) {
    fx.d[k] = fma(sx.d[0], X.d[k], fma(sx.d[1], Y.d[k], sx.d[2] * q.d[k]));
    fy.d[k] = fma(sy.d[0], X.d[k], fma(sy.d[1], Y.d[k], sy.d[2] * q.d[k]));
  }
}

// CSE / NFW_ELLIPSE_CSE / HERNQUIST_ELLIPSE_CSE: c[0..6] = cx, cy, cos phi, sin phi, q, 1/Rs (1 for CSE), const;
// k[0] = a, k[1] = s (CSE); flags as in the descriptor
LSIM_HD void cse_family_eval_d(const GLens& L, uint32_t flags, double x, double y, LD& fx, LD& fy) {
  const LD &cph = L.c[2], &sph = L.c[3], &q = L.c[4];
  const LD xs = x - L.c[0], ys = y - L.c[1];
  const LD xr = (xs * cph + ys * sph) * L.c[5], yr = (ys * cph - xs * sph) * L.c[5];
  LD gx, gy;
  if (L.n_tan == LSIM_LENS_CSE) {
    // the amplitude and core are parameters here: evaluate directly on the six-wide duals
    const LD &a = L.c[6], &sc = L.c[7];
    const bool major = (flags & LSIM_FLAG_CSE_MAJOR) != 0;
    const LD sq = major ? cst<kLensN>(1.0) : dsqrt(q);
    const LD X = xr * sq, Y = yr * sq;
    const LD q2 = q * q;
    const LD psi = dsqrt(q2 * (X * X + sc * sc) + Y * Y);
    const LD ps = psi + sc;
    const LD Phi = ps * ps + (1.0 - q2) * (X * X);
    const LD amp = major ? a : (a / q) * sq;
    const LD den = amp / (sc * psi * Phi);
    gx = q * X * (psi + q2 * sc) * den;
    gy = q * Y * ps * den;
  } else if (L.n_tan == LSIM_LENS_NFW_ELLIPSE_CSE) {
    const bool low = (flags & LSIM_FLAG_CSE_LOW_ACCURACY) != 0;
    cse_set_d(true, cse_table(low ? 3 : 1), cse_table(low ? 2 : 0), low ? LSIM_NFW_CSE_LOW_S_N : LSIM_NFW_CSE_HIGH_S_N,
              xr, yr, q, gx, gy);
    gx = gx * L.c[6];
    gy = gy * L.c[6];
  } else {
    cse_set_d(false, cse_table(5), cse_table(4), LSIM_HERNQUIST_CSE_S_N, xr, yr, q, gx, gy);
    gx = gx * L.c[6];
    gy = gy * L.c[6];
  }
  fx = gx * cph - gy * sph;
  fy = gx * sph + gy * cph;
}

constexpr int kGradLensCSEFamily = 65;  // prepared type of CSE / NFW_ELLIPSE_CSE / HERNQUIST_ELLIPSE_CSE (n_tan = which)

// Builds the prepared block of one lens descriptor from the parameter row.  Returns false for a profile outside the
// gradient's scope.
LSIM_HD bool lens_prepare_d(const lsim_lens_desc& d, const double* row, GLens& L, double* F, double* dF) {
  L.n_val = L.n_tan = 0;
  switch (d.type) {
    case LSIM_LENS_NULL:
      L.type = LSIM_LENS_NULL;
      return true;
    case LSIM_LENS_EPL:
      epl_prepare_d(L, F, dF, lens_var(d, row, 0), lens_var(d, row, 1), lens_var(d, row, 2), lens_var(d, row, 3),
                    lens_var(d, row, 4), lens_var(d, row, 5));
      return true;
    case LSIM_LENS_SIE:
      if (d.flags & LSIM_FLAG_SIE_NIE_FALSE)
        epl_prepare_d(L, F, dF, lens_var(d, row, 0), cst<kLensN>(d.opt[1]), lens_var(d, row, 1), lens_var(d, row, 2),
                      lens_var(d, row, 3), lens_var(d, row, 4));
      else
        nie_prepare_d(L, lens_var(d, row, 0), lens_var(d, row, 1), lens_var(d, row, 2), cst<kLensN>(d.opt[0]),
                      lens_var(d, row, 3), lens_var(d, row, 4));
      return true;
    case LSIM_LENS_NIE:
      nie_prepare_d(L, lens_var(d, row, 0), lens_var(d, row, 1), lens_var(d, row, 2), lens_var(d, row, 3),
                    lens_var(d, row, 4), lens_var(d, row, 5));
      return true;
    case LSIM_LENS_NFW: {
      const LD Rs0 = lens_var(d, row, 0), aRs = lens_var(d, row, 1);
      const LD rho0 = alpha2rho0_d(aRs, Rs0);
      const LD Rs = (Rs0.v < 0.0000001) ? cst<kLensN>(0.0000001) : Rs0;
      L.type = LSIM_LENS_NFW;
      L.c[0] = Rs;
      L.c[1] = 4.0 * rho0 * Rs;
      L.c[2] = lens_var(d, row, 2);
      L.c[3] = lens_var(d, row, 3);
      return true;
    }
    case LSIM_LENS_TNFW: {
      const LD Rs = lens_var(d, row, 0), aRs = lens_var(d, row, 1), rt = lens_var(d, row, 2);
      L.type = LSIM_LENS_TNFW;
      L.c[0] = Rs;
      L.c[1] = 4.0 * alpha2rho0_d(aRs, Rs) * Rs;
      L.c[2] = rt / Rs;
      L.c[3] = lens_var(d, row, 3);
      L.c[4] = lens_var(d, row, 4);
      return true;
    }
    case LSIM_LENS_MULTIPOLE:  // m, a_m, phi_m, center_x, center_y, r_E; m is an integer, not differentiable
      L.type = kGradLensMP;
      L.n_val = (int)param_of(d, row, 0);
      L.c[0] = lens_var(d, row, 3);
      L.c[1] = lens_var(d, row, 4);
      L.c[2] = lens_var(d, row, 1);
      L.c[3] = lens_var(d, row, 2);
      L.c[4] = lens_var(d, row, 5);
      L.k[0] = 1.0;
      mp_constants(L);
      return true;
    case LSIM_LENS_CSE:
    case LSIM_LENS_NFW_ELLIPSE_CSE:
    case LSIM_LENS_HERNQUIST_ELLIPSE_CSE: {
      LD phi, q;
      ellipticity2phi_q_d(lens_var(d, row, 2), lens_var(d, row, 3), phi, q);
      L.type = kGradLensCSEFamily;
      L.n_tan = d.type;
      L.n_val = (int)d.flags;
      L.c[0] = lens_var(d, row, 4);
      L.c[1] = lens_var(d, row, 5);
      L.c[2] = dcos(phi);
      L.c[3] = dsin(phi);
      L.c[4] = q;
      if (d.type == LSIM_LENS_CSE) {
        L.c[5] = cst<kLensN>(1.0);
        L.c[6] = lens_var(d, row, 0);
        L.c[7] = lens_var(d, row, 1);
      } else if (d.type == LSIM_LENS_NFW_ELLIPSE_CSE) {
        const LD Rs = lens_var(d, row, 0);
        L.c[5] = 1.0 / Rs;
        L.c[6] = 4.0 * alpha2rho0_d(lens_var(d, row, 1), Rs) * (Rs * Rs * Rs) / Rs;
      } else {
        const LD sigma0 = lens_var(d, row, 0), Rs = lens_var(d, row, 1);
        const LD rs_ = Rs / dsqrt(q);
        L.c[5] = 1.0 / Rs;
        L.c[6] = sigma0 * (rs_ * rs_) / Rs;
      }
      return true;
    }
    case kGradLensMP:
      mp_prepare_d(L, (int)d.opt[0], lens_var(d, row, 0), lens_var(d, row, 1), lens_var(d, row, 2), lens_var(d, row, 3),
                   lens_var(d, row, 4), lens_var(d, row, 5), param_of(d, row, 6));
      return true;
    case LSIM_LENS_SIS:
    case LSIM_LENS_SHEAR:
    case LSIM_LENS_CONVERGENCE:
      L.type = d.type;
      for (int i = 0; i < 4; ++i) L.c[i] = (i < 3 || d.type == LSIM_LENS_SHEAR) ? lens_var(d, row, i) : cst<kLensN>(0.0);
      return true;
    default:
      return false;
  }
}

LSIM_HD bool lens_grad_supported(const lsim_lens_desc& d) {
  switch (d.type) {
    case LSIM_LENS_NULL:
    case LSIM_LENS_EPL:
    case LSIM_LENS_SIE:
    case LSIM_LENS_NIE:
    case LSIM_LENS_SIS:
    case LSIM_LENS_SHEAR:
    case LSIM_LENS_CONVERGENCE:
    case LSIM_LENS_EPL_MULTIPOLE_M1M3M4:  // expanded into EPL + three kGradLensMP by the caller
    case kGradLensMP:
    case LSIM_LENS_NFW:
    case LSIM_LENS_TNFW:
    case LSIM_LENS_MULTIPOLE:
    case LSIM_LENS_CSE:
    case LSIM_LENS_NFW_ELLIPSE_CSE:
    case LSIM_LENS_HERNQUIST_ELLIPSE_CSE:
      return true;
    default:
      return false;
  }
}
// profiles whose point state (a 2 x 5 Jacobian) the kernel keeps in registers: the EPL family and NIE / SIE
LSIM_HD bool lens_grad_stateful(const lsim_lens_desc& d) {
  return d.type == LSIM_LENS_EPL || d.type == LSIM_LENS_SIE || d.type == LSIM_LENS_NIE;
}
LSIM_HD bool lens_grad_has_series(const lsim_lens_desc& d) {
  return d.type == LSIM_LENS_EPL || (d.type == LSIM_LENS_SIE && (d.flags & LSIM_FLAG_SIE_NIE_FALSE));
}

// The gradient path's own lens list: every profile as is, except EPL_MULTIPOLE_M1M3M4 (12 parameters), which becomes an
// EPL (parameters 0..5) and three kGradLensMP pseudo-profiles (theta_E, e1, e2, center_x, center_y, delta_phi_m,
// a_m/a) that read the parent's columns / constants -- their gradient entries therefore add up in the parent's columns.
// Returns the number of descriptors written (<= cap), or -1 if `cap` is too small.
inline int expand_grad_lens(const lsim_lens_desc* in, int n, lsim_lens_desc* out, int cap, bool keep_order = false) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (in[i].type != LSIM_LENS_EPL_MULTIPOLE_M1M3M4) {
      if (m + 1 > cap) return -1;
      out[m++] = in[i];
      continue;
    }
    if (m + 4 > cap) return -1;
    lsim_lens_desc e = in[i];
    e.type = LSIM_LENS_EPL;
    for (int k = 6; k < LSIM_MAX_PARAMS; ++k) {
      e.col[k] = -1;
      e.cst[k] = 0.0;
    }
    out[m++] = e;
    const int order[3] = {1, 3, 4}, ia[3] = {6, 8, 10}, id[3] = {7, 9, 11};
    const int src[5] = {0, 2, 3, 4, 5};  // theta_E, e1, e2, center_x, center_y of the parent
    for (int j = 0; j < 3; ++j) {
      lsim_lens_desc q = in[i];
      q.type = kGradLensMP;
      q.opt[0] = order[j];
      for (int k = 0; k < LSIM_MAX_PARAMS; ++k) {
        q.col[k] = -1;
        q.cst[k] = 0.0;
      }
      for (int k = 0; k < 5; ++k) {
        q.col[k] = in[i].col[src[k]];
        q.cst[k] = in[i].cst[src[k]];
      }
      q.col[5] = in[i].col[id[j]];
      q.cst[5] = in[i].cst[id[j]];
      q.col[6] = in[i].col[ia[j]];
      q.cst[6] = in[i].cst[ia[j]];
      q.delta_T = 0.0;  // multi-plane: the pseudo-profiles sit on their parent's plane (no step between them)
      out[m++] = q;
    }
  }
  // the profiles whose point state the kernel keeps in registers (the EPL family, NIE / SIE) go first; the order of the sum of
  // deflections is internal to the gradient path (single plane only: a multi-plane list stays in plane order)
  for (int i = 1; i < m && !keep_order; ++i) {
    if (!lens_grad_stateful(out[i])) continue;
    int j = i;
    while (j > 0 && !lens_grad_stateful(out[j - 1])) {
      const lsim_lens_desc t = out[j - 1];
      out[j - 1] = out[j];
      out[j] = t;
      --j;
    }
  }
  return m;
}

// deflection of one prepared profile at (x, y) with its parameter derivatives
LSIM_HD void lens_eval_d(const GLens& L, const double* pool, double x, double y, LD& fx, LD& fy) {
  switch (L.type) {
    case LSIM_LENS_EPL:
      epl_eval_d(L, pool + L.coef_off, pool + L.coef_off + kMaxSeries, x, y, fx, fy);
      break;
    case LSIM_LENS_NIE:
      nie_eval_d(L, x, y, fx, fy);
      break;
    case kGradLensMP:
      mp_eval_d(L, x, y, fx, fy);
      fx = fx * L.k[0];
      fy = fy * L.k[0];
      break;
    case LSIM_LENS_NFW:
      nfw_eval_d(L, x, y, fx, fy);
      break;
    case LSIM_LENS_TNFW:
      tnfw_eval_d(L, x, y, fx, fy);
      break;
    case kGradLensCSEFamily:
      cse_family_eval_d(L, (uint32_t)L.n_val, x, y, fx, fy);
      break;
    case LSIM_LENS_SIS: {  // sis.py:35-43: theta_E, center_x, center_y
      const LD xs = x - L.c[1], ys = y - L.c[2];
      LD R = dsqrt(xs * xs + ys * ys);
      if (R.v < 0.000001) R = cst<kLensN>(0.000001);
      const LD a = L.c[0] / R;
      fx = a * xs;
      fy = a * ys;
      break;
    }
    case LSIM_LENS_SHEAR: {  // shear.py:39-53: gamma1, gamma2, ra_0, dec_0
      const LD xs = x - L.c[2], ys = y - L.c[3];
      fx = L.c[0] * xs + L.c[1] * ys;
      fy = L.c[1] * xs - L.c[0] * ys;
      break;
    }
    case LSIM_LENS_CONVERGENCE: {  // convergence.py:35-46: kappa, ra_0, dec_0
      fx = L.c[0] * (x - L.c[1]);
      fy = L.c[0] * (y - L.c[2]);
      break;
    }
    default:
      fx = cst<kLensN>(0.0);
      fy = cst<kLensN>(0.0);
  }
}

// ---- Hessian of the lens potential (SinglePlane.hessian, LensModel/single_plane.py:142-180) ---------------------
// Every profile on the path depends on the position only through (x - center_x, y - center_y) (SHEAR / CONVERGENCE:
// ra_0, dec_0), so  f_xx = -d alpha_x / d center_x,  f_xy = -d alpha_x / d center_y,  f_yx = -d alpha_y / d center_x,
// f_yy = -d alpha_y / d center_y  -- the tangents lens_eval_d already carries.  For the EPL family these run through
// the 100-term tangent series (the reference's custom_jvp), which agrees with the reference's closed-form
// kappa / alpha identities (Profiles/epl.py: EPLMajorAxis.hessian) far inside the 1e-10 bar.

// tangent slots of (center_x, center_y) in a descriptor of the gradient's lens list; false: no position dependence
LSIM_HD bool lens_center_slots(const lsim_lens_desc& d, int& ix, int& iy) {
  switch (d.type) {
    case LSIM_LENS_EPL:
    case LSIM_LENS_NIE:
    case LSIM_LENS_CSE:
    case LSIM_LENS_NFW_ELLIPSE_CSE:
    case LSIM_LENS_HERNQUIST_ELLIPSE_CSE: ix = 4; iy = 5; return true;
    case LSIM_LENS_SIE:
    case LSIM_LENS_TNFW:
    case LSIM_LENS_MULTIPOLE:
    case kGradLensMP: ix = 3; iy = 4; return true;
    case LSIM_LENS_SHEAR:
    case LSIM_LENS_NFW: ix = 2; iy = 3; return true;
    case LSIM_LENS_SIS:
    case LSIM_LENS_CONVERGENCE: ix = 1; iy = 2; return true;
    default: ix = iy = 0; return false;
  }
}

// EPL family on the Hessian path: closed form from the FORWARD kernel's deflection.  Profiles/epl.py:398-435
// (Tessore & Metcalf eq. 2 and 17, corrigendum): kappa = (2 - t)/2 (b/R)^t and, with alpha = alpha_x + i alpha_y and
// u = e^{i phi} of the centred position in the SKY frame (the reference forms the same expression in the major-axis
// frame and rotates the shear by 2 phi_G, :224-232 -- identical up to rounding),
//   gamma_1 + i gamma_2 = (1 - t) alpha u / r - kappa u^2;   f_xx = kappa + gamma_1, f_yy = kappa - gamma_1, f_xy = gamma_2.
// The prepared block is the forward path's LensC (lens_prepare / epl_eval1 of lsim_profiles.cuh), stored inside the
// GLens slot, so the deflection is bit-identical to lsim_ray_shoot's and costs one nmax-term series per point instead
// of the 100-term Dual<6> evaluation.
constexpr int kHessLensEPL = 66;  // GLens.type of an EPL-family profile prepared for the Hessian path
static_assert(sizeof(LensC) <= sizeof(GLens) - 16, "LensC must fit into the GLens payload");
LSIM_HD LensC& hess_forward(GLens& L) { return *reinterpret_cast<LensC*>(L.c); }
LSIM_HD const LensC& hess_forward(const GLens& L) { return *reinterpret_cast<const LensC*>(L.c); }

LSIM_HD void epl_value_hessian(const LensC& C, const double* coef, double x, double y, double& ax, double& ay,
                               double (&h)[4]) {
  const double xs = x - C.c[0], ys = y - C.c[1];
  epl_eval1(C, coef, xs, ys, ax, ay);
  const double cph = C.c[2], sph = C.c[3], q = C.c[4], t = -2.0 * C.c[6];
  const double xr = fma(ys, sph, xs * cph), c = fma(ys, cph, -(xs * sph)), a = q * xr;
  const double R2c = max_nan(fma(c, c, a * a), 1e-16);  // R = max(R, 1e-8), epl.py:411
  double kappa = (2.0 - t) / 2.0 * exp2_fin(fma(C.c[6], log2_pos(R2c), C.c[7]));
  kappa = nan_to_num(kappa, 1e10, -1e10);
  const double r = sqrt(fma(ys, ys, xs * xs));
  const double ux = xs / r, uy = ys / r;
  const double u2r = ux * ux * 2.0 - 1.0, u2i = uy * ux * 2.0;
  double g1 = (1.0 - t) * (ax * ux - ay * uy) / r - kappa * u2r;
  double g2 = (1.0 - t) * (ay * ux + ax * uy) / r - kappa * u2i;
  g1 = nan_to_num(g1, 1e10, -1e10);
  g2 = nan_to_num(g2, 1e10, -1e10);
  h[0] = kappa + g1;
  h[1] = g2;
  h[2] = g2;
  h[3] = kappa - g1;
}

// Preparation for the Hessian path.  EPL family: the forward block (above).  Everything else: lens_prepare_d with
// EVERY parameter carried as an independent variable (a fixed centre still needs its tangent) -- the descriptor is
// re-pointed at a private row that holds the resolved parameter values.
LSIM_HD bool lens_prepare_hess(const lsim_lens_desc& d, const double* row, GLens& L, double* F, double* dF,
                               double series_tol) {
  if (lens_grad_has_series(d)) {
    lens_prepare(d, row, hess_forward(L), F, series_tol);
    L.type = kHessLensEPL;
    return true;
  }
  lsim_lens_desc dd = d;
  double r2[LSIM_MAX_PARAMS];
  for (int k = 0; k < LSIM_MAX_PARAMS; ++k) {
    r2[k] = param_of(d, row, k);
    dd.col[k] = k;
  }
  return lens_prepare_d(dd, r2, L, F, dF);
}

// h[0..3] += (f_xx, f_xy, f_yx, f_yy) of one prepared profile
LSIM_HD void lens_hessian_add(const GLens& L, const double* pool, int ix, int iy, double x, double y, double (&h)[4]) {
  if (L.type == kHessLensEPL) {
    double ax, ay, e[4];
    epl_value_hessian(hess_forward(L), pool + L.coef_off, x, y, ax, ay, e);
    for (int k = 0; k < 4; ++k) h[k] += e[k];
    return;
  }
  LD fx, fy;
  lens_eval_d(L, pool, x, y, fx, fy);
  h[0] -= fx.d[ix];
  h[1] -= fx.d[iy];
  h[2] -= fy.d[ix];
  h[3] -= fy.d[iy];
}

// value and Hessian of one prepared profile from ONE dual-number evaluation: a += alpha, h += (f_xx, f_xy, f_yx, f_yy)
LSIM_HD void lens_value_hessian_add(const GLens& L, const double* pool, int ix, int iy, double x, double y, double& ax,
                                    double& ay, double (&h)[4]) {
  if (L.type == kHessLensEPL) {
    double ex, ey, e[4];
    epl_value_hessian(hess_forward(L), pool + L.coef_off, x, y, ex, ey, e);
    ax += ex;
    ay += ey;
    for (int k = 0; k < 4; ++k) h[k] += e[k];
    return;
  }
  LD fx, fy;
  lens_eval_d(L, pool, x, y, fx, fy);
  ax += fx.v;
  ay += fy.v;
  if (ix >= 0) {
    h[0] -= fx.d[ix];
    h[1] -= fx.d[iy];
    h[2] -= fy.d[ix];
    h[3] -= fy.d[iy];
  }
}

// _compute_penalty, Sampling/Likelihoods/position_likelihood.py:296-352: h = Hessian at the image, (d0, d1) = mean
// source position minus this image's source position.  A = 1 - h; Sigma_beta = (A^T Sigma_theta) A with
// Sigma_theta = sigma^2 1 (the reference's matmul order); penalty = d^T adj(Sigma_beta) d / (2 det Sigma_beta), 1e15 where
// det == 0; + 1e3 where |d|^2 > hard_bound_rms^2 (hard_bound_rms < 0: no bound).
LSIM_HD double source_position_penalty(const double (&h)[4], double d0, double d1, double sigma, double hard_bound_rms) {
  const double a00 = 1.0 - h[0], a01 = -h[1], a10 = -h[2], a11 = 1.0 - h[3];
  const double s2 = sigma * sigma;
  const double t00 = a00 * s2, t01 = a10 * s2, t10 = a01 * s2, t11 = a11 * s2;
  const double sa = t00 * a00 + t01 * a10, sb = t00 * a01 + t01 * a11, sc = t10 * a00 + t11 * a10,
               sd = t10 * a01 + t11 * a11;
  const double det = sa * sd - sb * sc;
  const double q = d0 * (sd * d0 - sb * d1) + d1 * (-sc * d0 + sa * d1);
  double pen = (det == 0.0) ? 1e15 : q / (2.0 * det);
  if (hard_bound_rms >= 0.0 && d0 * d0 + d1 * d1 > hard_bound_rms * hard_bound_rms) pen += 1e3;
  return pen;
}

// ---- light ------------------------------------------------------------------------------------------------------
using GD = Dual<kLightN>;

LSIM_HD bool light_grad_supported(const lsim_light_desc& d) {
  return d.type == LSIM_LIGHT_SERSIC || d.type == LSIM_LIGHT_SERSIC_ELLIPSE || d.type == LSIM_LIGHT_GAUSSIAN;
}

// tangent slot 2 + i carries parameter i; `amp` (parameter 0) is 1 and constant when the amplitude is solved linearly
LSIM_HD GD light_var(const lsim_light_desc& d, const double* row, int i) {
  if (i == 0 && !d.use_amp) return cst<kLightN>(1.0);
  return var<kLightN>(param_of(d, row, i), d.col[i] >= 0 ? 2 + i : -1);
}

// SersicUtil._r_sersic (sersic_utils.py:247-267) on a dual radius
LSIM_HD GD r_sersic_d(const GD& R, const GD& R_sersic, const GD& n_sersic) {
  const GD R_ = dmax(R, 0.0001);
  const GD Rs_ = dmax(R_sersic, 0.0001);
  const GD bn = dmax(1.9992 * n_sersic - 0.3271, 0.00001);
  const GD R_frac = R_ / Rs_;
  if (!(R_frac.v <= 1000.0)) return cst<kLightN>(0.0);
  const GD expo = -(bn * (dpow(R_frac, 1.0 / n_sersic) - 1.0));
  return dnan_to_num(dexp(expo), DBL_MAX, -DBL_MAX);
}

// surface brightness of one light component at the dual point (x, y) -- x.d[0] = y.d[1] = 1 when the point is a
// ray-shot source position, all-zero tangents for light evaluated on the image plane
// (sersic_ellipse.py:61-93, sersic.py:54-76, gaussian.py:27-42, param_util.py:102-123)
LSIM_HD GD light_eval_d(const lsim_light_desc& d, const double* row, const GD& x, const GD& y) {
  switch (d.type) {
    case LSIM_LIGHT_SERSIC_ELLIPSE:
    case LSIM_LIGHT_SERSIC: {
      const bool ell = d.type == LSIM_LIGHT_SERSIC_ELLIPSE;
      const GD amp = light_var(d, row, 0);
      GD Rs = light_var(d, row, 1);
      const GD n = light_var(d, row, 2);
      const GD e1 = ell ? light_var(d, row, 3) : cst<kLightN>(0.0);
      const GD e2 = ell ? light_var(d, row, 4) : cst<kLightN>(0.0);
      const GD cx = light_var(d, row, ell ? 5 : 3), cy = light_var(d, row, ell ? 6 : 4);
      if (ell) Rs = dmax(Rs, 0.0);
      const GD xs = x - cx, ys = y - cy;
      const GD norm = dmax(dsqrt(dabs(1.0 - e1 * e1 - e2 * e2)), 0.000001);
      const GD x_ = ((1.0 - e1) * xs - e2 * ys) / norm;
      const GD y_ = ((1.0 + e1) * ys - e2 * xs) / norm;
      const GD R = dsqrt(x_ * x_ + y_ * y_);
      return amp * r_sersic_d(R, Rs, n);
    }
    case LSIM_LIGHT_GAUSSIAN: {
      const GD amp = light_var(d, row, 0), sigma = light_var(d, row, 1);
      const GD cx = light_var(d, row, 2), cy = light_var(d, row, 3);
      const GD s2 = sigma * sigma;
      const GD c = amp / (2.0 * kPi * s2);
      const GD dx = x - cx, dy = y - cy;
      const GD r2 = dx * dx / s2 + dy * dy / s2;
      return c * dexp(r2 * (-0.5));
    }
    default:
      return cst<kLightN>(0.0);
  }
}


// ---- hand-derived gradients (the fast path of k_image_grad) ------------------------------------------------------
// Per sub-pixel the kernel needs, for a given 2-vector v (the back-propagated source-plane gradient),
//   g_k = d(v . alpha)/d(parameter k of the profile).
// LensPt carries what lens_point() leaves behind for lens_vjp(): for the EPL the Jacobian of the major-axis
// deflection alpha' w.r.t. (x', y', q, t, theta_E) and alpha' itself; for the profiles evaluated on dual numbers the
// 2 x 6 tangents.  Either way 12 doubles.
struct LensPt {
  double ax, ay;
  double s[12];
};

// S = sum_n c_n w^n for real c_n by the quadratic-factor (Clenshaw) recurrence  b_k = c_k + 2 Re(w) b_{k+1} -
// |w|^2 b_{k+2},  S = c_0 + w b_1 - |w|^2 b_2.  Four series in one loop (four independent FMA chains):
//   F cut at n_val (the value of 2F1, nmax schedule), F (100-term S), (n+1) F_{n+1} (dS/dw), dF_n/dt (dS/dt).
// A value series longer than the tangent series (nmax = 200, 500) runs its head alone first.
LSIM_HD void series4(const double* F, int n_val, const Coef4* T, int n_tan, double wr, double wi, double (&o)[8]) {
  const double pp = wr + wr, qq = -fma(wr, wr, wi * wi);
  double a1 = 0.0, a2 = 0.0, f1 = 0.0, f2 = 0.0, b1 = 0.0, b2 = 0.0, c1 = 0.0, c2 = 0.0;
  for (int k = n_val - 1; k >= n_tan; --k) {
    const double t = fma(pp, a1, fma(qq, a2, F[k]));
    a2 = a1;
    a1 = t;
  }
  const bool longv = n_val > n_tan;
  for (int k = n_tan - 1; k >= 1; --k) {
    const Coef4 c = T[k];
    const double ta = fma(pp, a1, fma(qq, a2, longv ? c.f : c.fv));
    const double tf = fma(pp, f1, fma(qq, f2, c.f));
    const double tb = fma(pp, b1, fma(qq, b2, c.g));
    const double tc = fma(pp, c1, fma(qq, c2, c.d));
    a2 = a1;
    a1 = ta;
    f2 = f1;
    f1 = tf;
    b2 = b1;
    b1 = tb;
    c2 = c1;
    c1 = tc;
  }
  const Coef4 c = T[0];
  o[0] = fma(a1, wr, fma(qq, a2, c.f));
  o[1] = a1 * wi;
  o[2] = fma(b1, wr, fma(qq, b2, c.g));
  o[3] = b1 * wi;
  o[4] = fma(c1, wr, fma(qq, c2, c.d));
  o[5] = c1 * wi;
  o[6] = fma(f1, wr, fma(qq, f2, c.f));
  o[7] = f1 * wi;
}

// EPL: alpha' = A R^-t Z H(z, t), A = 2 b^t / (1+q), Z = q x' + i y', R = max(|Z|, 1e-9), z = f Q, Q = -Z^2/|Z|^2,
// H = u S(w), u = 1/(1 - z/2) = 2|Z|^2 conj(D)/|D|^2, w = z/(z-2) = f Z^2 conj(D)/|D|^2, D = f Z^2 + 2|Z|^2.
//   d alpha' = pre [ H (dlnA - t dlnR - lnR dt + kappa) + H_t dt + H_z dz ],   pre = A R^-t Z,  kappa = dZ / Z,
//   dlnR = Re kappa (0 where R is clamped),  dz = df Q + z 2i Im kappa,
//   H_t = u S_t,  H_z = u^2 (S - u S') / 2   with the 100-term series (the reference's custom_jvp; H itself: nmax terms).
LSIM_HD void epl_point(const GLens& L, const double* pool, double x, double y, LensPt& o) {
  const double* F = pool + L.coef_off;
  const Coef4* T = reinterpret_cast<const Coef4*>(F + kMaxSeries + kTanTerms);
  const double cph = L.k[0], sph = L.k[1], f = L.k[2], q = L.p[3].v, t = L.p[1].v;
  const double xs = x - L.p[4].v, ys = y - L.p[5].v;
  const double xr = fma(ys, sph, xs * cph), yr = fma(ys, cph, -(xs * sph));
  const double Zr = q * xr, Zi = yr;
  const double R2 = fma(Zi, Zi, Zr * Zr);
  const double z2r = fma(-Zi, Zi, Zr * Zr), z2i = 2.0 * Zr * Zi;
  const double Dr = fma(f, z2r, R2 + R2), Di = f * z2i;
  const double inv = rcp_fast(fma(Dr, Dr, Di * Di));
  const double finv = f * inv;
  const double wr = fma(z2i, Di, z2r * Dr) * finv, wi = fma(z2i, Dr, -(z2r * Di)) * finv;
  const double s2 = (R2 + R2) * inv;
  const double ur = s2 * Dr, ui = -(s2 * Di);
  double S[8];  // S_val, S', S_t, S_100
  series4(F, L.n_val, T, L.n_tan, wr, wi, S);
  const double Sr = S[6], Si = S[7];
  const double Hr = fma(ur, S[0], -(ui * S[1])), Hi = fma(ur, S[1], ui * S[0]);      // H = u S_val
  const double Htr = fma(ur, S[4], -(ui * S[5])), Hti = fma(ur, S[5], ui * S[4]);    // H_t = u S_t
  const double usr = fma(ur, S[2], -(ui * S[3])), usi = fma(ur, S[3], ui * S[2]);    // u S'
  const double ar_ = Sr - usr, ai_ = Si - usi;
  const double u2r = fma(ur, ur, -(ui * ui)), u2i = 2.0 * ur * ui;
  const double Hzr = 0.5 * fma(u2r, ar_, -(u2i * ai_)), Hzi = 0.5 * fma(u2r, ai_, u2i * ar_);
  const double iR2 = rcp_fast(R2);
  const double Qr = -(z2r * iR2), Qi = -(z2i * iR2);
  const double Yr = fma(Hzr, Qr, -(Hzi * Qi)), Yi = fma(Hzr, Qi, Hzi * Qr);          // H_z Q
  const double Yzr = f * Yr, Yzi = f * Yi;                                          // H_z z
  const double R2c = max_nan(R2, 1e-18);
  const double lg = log2_pos(R2c);
  const double P = L.k[5] * exp2_fin(fma(-0.5 * t, lg, L.k[4]));
  const double lnR = 0.5 * kLn2 * lg;
  const double teff = (R2 < 1e-18) ? 0.0 : t;
  const double prr = P * Zr, pri = P * Zi;
  double ar = fma(prr, Hr, -(pri * Hi)), ai = fma(prr, Hi, pri * Hr);
  const bool finite = (fabs(ar) <= DBL_MAX) && (fabs(ai) <= DBL_MAX);
  // kappa of the three directions that move Z: x' (dZ = q), y' (dZ = i), q (dZ = x')
  const double izr = Zr * iR2, izi = -(Zi * iR2);  // 1 / Z
  const double kr[3] = {q * izr, -izi, xr * izr};
  const double ki[3] = {q * izi, izr, xr * izi};
  const double ex[3] = {0.0, 0.0, L.k[6]};
  const double df[3] = {0.0, 0.0, L.k[7]};
  LSIM_UNROLL
  for (int d = 0; d < 3; ++d) {
    const double c0 = ex[d] - teff * kr[d];
    // D = H kappa + c0 H + 2 Im(kappa) i (H_z z) + df (H_z Q)
    const double Dre = fma(Hr, kr[d], -(Hi * ki[d])) + fma(c0, Hr, fma(-2.0 * ki[d], Yzi, df[d] * Yr));
    const double Dim = fma(Hr, ki[d], Hi * kr[d]) + fma(c0, Hi, fma(2.0 * ki[d], Yzr, df[d] * Yi));
    o.s[2 * d] = finite ? fma(prr, Dre, -(pri * Dim)) : 0.0;
    o.s[2 * d + 1] = finite ? fma(prr, Dim, pri * Dre) : 0.0;
  }
  {  // t: D = H (ln b - ln R) + H_t
    const double c0 = L.k[3] - lnR;
    const double Dre = fma(c0, Hr, Htr), Dim = fma(c0, Hi, Hti);
    o.s[6] = finite ? fma(prr, Dre, -(pri * Dim)) : 0.0;
    o.s[7] = finite ? fma(prr, Dim, pri * Dre) : 0.0;
  }
  const double tte = t / L.p[0].v;  // theta_E: alpha' t / theta_E
  o.s[8] = finite ? ar * tte : 0.0;
  o.s[9] = finite ? ai * tte : 0.0;
  ar = nan_to_num(ar, 1e10, -1e10);
  ai = nan_to_num(ai, 1e10, -1e10);
  o.s[10] = finite ? ar : 0.0;
  o.s[11] = finite ? ai : 0.0;
  o.ax = fma(-ai, sph, ar * cph);
  o.ay = fma(ai, cph, ar * sph);
}

LSIM_HD void epl_vjp(const GLens& L, const LensPt& o, double x, double y, double vx, double vy, double (&g)[kLensEntries]) {
  const double cph = L.k[0], sph = L.k[1];
  const double Vr = fma(vy, sph, vx * cph), Vi = fma(vy, cph, -(vx * sph));  // v in the major-axis frame
  const double sx = fma(Vr, o.s[0], Vi * o.s[1]), sy = fma(Vr, o.s[2], Vi * o.s[3]);
  const double sq = fma(Vr, o.s[4], Vi * o.s[5]), st = fma(Vr, o.s[6], Vi * o.s[7]);
  const double sE = fma(Vr, o.s[8], Vi * o.s[9]);
  const double xs = x - L.p[4].v, ys = y - L.p[5].v;
  const double xr = fma(ys, sph, xs * cph), yr = fma(ys, cph, -(xs * sph));
  // phi: x' and y' rotate (dx'/dphi = y', dy'/dphi = -x') and alpha = e^{i phi} alpha' (d/dphi = i alpha)
  const double sphi = fma(sx, yr, -(sy * xr)) + fma(Vi, o.s[10], -(Vr * o.s[11]));
  const double scx = fma(sph, sy, -(cph * sx)), scy = -fma(sph, sx, cph * sy);
  g[0] = sE;
  g[1] = st;
  g[2] = sphi;
  g[3] = sq;
  g[4] = scx;
  g[5] = scy;
  g[6] = 0.0;
  g[7] = 0.0;
}

// value + whatever lens_vjp needs, for any profile in the gradient's scope
// ---- multipole (kGradLensMP) in closed form ------------------------------------------------------------------------
// Plain constants next to the duals c[0..4] = cx, cy, rho (= a_m at unit amplitude), phi_m, r_E:
//   k[0] = amplitude, k[2] = rho, k[3] = 1/(1-m^2) (m != 1), k[4], k[5] = cos, sin(m phi_m), k[6] = ln r_E
LSIM_HD void mp_constants(GLens& L) {
  const int m = L.n_val;
  L.k[2] = L.c[2].v;
  L.k[3] = m == 1 ? 0.0 : 1.0 / (1.0 - (double)(m * m));
  L.k[4] = cos(m * L.c[3].v);
  L.k[5] = sin(m * L.c[3].v);
  L.k[6] = log(L.c[4].v);
}

// unit vector of the point, (cos, sin)(m Delta) with Delta = phi - phi_m by complex powers (no trigonometric call), r^2
LSIM_HD void mp_terms(const GLens& L, double x, double y, double& c1, double& s1, double& cm, double& sm, double& r2) {
  const int m = L.n_val;
  const double xs = x - L.c[0].v, ys = y - L.c[1].v;
  r2 = fma(ys, ys, xs * xs);
  c1 = 1.0;
  s1 = 0.0;  // atan2(0, 0) = 0
  if (r2 > 0) {
    const double ir = rsqrt_fast(fmax(r2, 1e-300));
    c1 = xs * ir;
    s1 = ys * ir;
  }
  double ur = c1, ui = s1;
  for (int k = 1; k < m; ++k) {
    const double t = fma(ur, c1, -(ui * s1));
    ui = fma(ur, s1, ui * c1);
    ur = t;
  }
  cm = fma(ui, L.k[5], ur * L.k[4]);  // cos(m phi - m phi_m)
  sm = fma(ui, L.k[4], -(ur * L.k[5]));
}

// deflection at unit amplitude (multipole.py:72-119)
LSIM_HD void mp_value(const GLens& L, double x, double y, double& hx, double& hy) {
  double c1, s1, cm, sm, r2;
  mp_terms(L, x, y, c1, s1, cm, sm, r2);
  const int m = L.n_val;
  if (m == 1) {
    const double lg = 0.5 * kLn2 * log2_pos(fmax(r2, 1e-12)) - L.k[6];
    const double h = 0.5 * L.k[2];
    hx = h * fma(L.k[4], lg, cm * c1);
    hy = h * fma(L.k[5], lg, cm * s1);
  } else {
    const double A = L.k[2] * L.k[3];
    hx = A * fma(m * s1, sm, c1 * cm);
    hy = A * fma(-m * c1, sm, s1 * cm);
  }
}

// s_j = d(a v . alpha_hat)/d(primitive j), primitives (rho, phi_m, cx, cy, r_E) in g[0..4]; g[6] = v . alpha_hat
LSIM_HD void mp_vjp(const GLens& L, double x, double y, double vx, double vy, double (&g)[kLensEntries]) {
  double c1, s1, cm, sm, r2;
  mp_terms(L, x, y, c1, s1, cm, sm, r2);
  const int m = L.n_val;
  const double a = L.k[0], rho = L.k[2], ar = a * rho;
  // everything below at unit rho: alpha_hat = rho * h
  double dpx, dpy;              // d h / d phi (polar angle of the point)
  double dlx = 0.0, dly = 0.0;  // d h / d ln r   (m = 1 only)
  double hx, hy, dmx, dmy;      // h, d h / d phi_m
  if (m == 1) {
    const bool live = !(r2 < 1e-12);
    const double lg = 0.5 * kLn2 * log2_pos(fmax(r2, 1e-12)) - L.k[6];
    hx = 0.5 * fma(L.k[4], lg, cm * c1);
    hy = 0.5 * fma(L.k[5], lg, cm * s1);
    dmx = 0.5 * fma(-L.k[5], lg, sm * c1);
    dmy = 0.5 * fma(L.k[4], lg, sm * s1);
    dpx = -0.5 * fma(sm, c1, cm * s1);
    dpy = 0.5 * fma(cm, c1, -(sm * s1));
    dlx = live ? 0.5 * L.k[4] : 0.0;
    dly = live ? 0.5 * L.k[5] : 0.0;
    g[4] = -ar * 0.5 * fma(vx, L.k[4], vy * L.k[5]) / L.c[4].v;  // ln(r / r_E): d/d r_E = -1 / r_E
  } else {
    const double A = L.k[3];
    hx = A * fma(m * s1, sm, c1 * cm);
    hy = A * fma(-m * c1, sm, s1 * cm);
    dmx = A * m * fma(c1, sm, -(m * s1 * cm));
    dmy = A * m * fma(s1, sm, m * c1 * cm);
    dpx = -cm * s1;  // A (m^2 - 1) cos(m Delta) (sin, -cos)(phi)
    dpy = cm * c1;
    g[4] = 0.0;
  }
  const double sh = fma(vx, hx, vy * hy);
  g[6] = rho * sh;
  g[0] = a * sh;
  g[1] = ar * fma(vx, dmx, vy * dmy);
  // position: d phi / d xs = -sin / r, d phi / d ys = cos / r, d ln r / d xs = cos / r, d ln r / d ys = sin / r;
  // the centre moves the point the other way
  const double ir = r2 > 0 ? rsqrt_fast(fmax(r2, 1e-300)) : 0.0;
  const double sp = ar * fma(vx, dpx, vy * dpy) * ir, sl = ar * fma(vx, dlx, vy * dly) * ir;
  g[2] = -fma(-s1, sp, c1 * sl);
  g[3] = -fma(c1, sp, s1 * sl);
  g[5] = 0.0;
  g[7] = 0.0;
}

// ---- NIE / SIE in closed form ----------------------------------------------------------------------------------------
// Major-axis deflection (nie.py:280-294,333-342):  psi = sqrt(q^2 (s^2 + x^2) + y^2),  e = sqrt(1 - q^2),
//   alpha_x = (b/e) atan(e x / (psi + s)),   alpha_y = (b/e) atanh(e y / (psi + q^2 s)).
// Jacobian w.r.t. the primitives (x', y', q, b, s) in o.s[0..9], alpha' in o.s[10..11]; the prepared duals give the map
// to the parameters: c[0..6] = cx, cy, cos phi, sin phi, q (clamped), b, s.
LSIM_HD void nie_point(const GLens& L, double x, double y, LensPt& o) {
  const double cph = L.c[2].v, sph = L.c[3].v, q = L.c[4].v, b = L.c[5].v, s = L.c[6].v;
  const double xs = x - L.c[0].v, ys = y - L.c[1].v;
  const double xr = fma(ys, sph, xs * cph), yr = fma(ys, cph, -(xs * sph));
  const double q2 = q * q;
  const double psi = sqrt(fma(q2, fma(xr, xr, s * s), yr * yr));
  const double e2 = 1.0 - q2, e = sqrt(e2), be = b / e;
  const double dA = psi + s, dB = fma(q2, s, psi);
  const double iA = 1.0 / dA, iB = 1.0 / dB;
  const double A = e * xr * iA, B = e * yr * iB;
  const double tA = atan(A), tB = atanh(B);
  const double ax = be * tA, ay = be * tB;
  const double Dx = be / fma(A, A, 1.0), Dy = be / fma(-B, B, 1.0);  // be d atan / dA, be d atanh / dB
  const double ipsi = psi > 0 ? 1.0 / psi : 0.0;
  // d psi / d(x, y, q, s)
  const double px = q2 * xr * ipsi, py = yr * ipsi, pq = q * fma(xr, xr, s * s) * ipsi, ps = q2 * s * ipsi;
  const double de_q = -q / e;              // d e / d q
  const double dbe_q = b * q / (e * e2);   // d (b/e) / d q
  // dA = de x/(psi+s) + e dx/(psi+s) - e x (dpsi + ds)/(psi+s)^2, likewise dB with (psi + q^2 s)
  const double Ax = e * iA - A * iA * px, Ay = -A * iA * py;
  const double Aq = de_q * xr * iA - A * iA * pq, As = -A * iA * (ps + 1.0);
  const double Bx = -B * iB * px, By = e * iB - B * iB * py;
  const double Bq = de_q * yr * iB - B * iB * fma(2.0 * q, s, pq), Bs = -B * iB * (ps + q2);
  o.s[0] = Dx * Ax;
  o.s[1] = Dy * Bx;
  o.s[2] = Dx * Ay;
  o.s[3] = Dy * By;
  o.s[4] = fma(dbe_q, tA, Dx * Aq);
  o.s[5] = fma(dbe_q, tB, Dy * Bq);
  o.s[6] = tA / e;  // d / d b
  o.s[7] = tB / e;
  o.s[8] = Dx * As;
  o.s[9] = Dy * Bs;
  o.s[10] = ax;
  o.s[11] = ay;
  o.ax = fma(-ay, sph, ax * cph);
  o.ay = fma(ay, cph, ax * sph);
}

// primitives: g = d(v . alpha)/d(b, q, s, phi, cx, cy)
LSIM_HD void nie_vjp(const GLens& L, const LensPt& o, double x, double y, double vx, double vy,
                     double (&g)[kLensEntries]) {
  const double cph = L.c[2].v, sph = L.c[3].v;
  const double Vr = fma(vy, sph, vx * cph), Vi = fma(vy, cph, -(vx * sph));
  const double sx = fma(Vr, o.s[0], Vi * o.s[1]), sy = fma(Vr, o.s[2], Vi * o.s[3]);
  const double xs = x - L.c[0].v, ys = y - L.c[1].v;
  const double xr = fma(ys, sph, xs * cph), yr = fma(ys, cph, -(xs * sph));
  g[0] = fma(Vr, o.s[6], Vi * o.s[7]);
  g[1] = fma(Vr, o.s[4], Vi * o.s[5]);
  g[2] = fma(Vr, o.s[8], Vi * o.s[9]);
  g[3] = fma(sx, yr, -(sy * xr)) + fma(Vi, o.s[10], -(Vr * o.s[11]));
  g[4] = fma(sph, sy, -(cph * sx));
  g[5] = -fma(sph, sx, cph * sy);
  g[6] = 0.0;
  g[7] = 0.0;
}

// TIER: what is compiled into an instantiation of k_image_grad.  0: EPL family + SHEAR (+ NULL) only -- the headline
// models; 1: also the closed-form multipoles and NIE / SIE; 2: everything, including the dual-number profiles.  Lower
// tiers carry less code and fewer registers (140 vs 212), so the headline model never pays for profiles it does not use.
template <int TIER = 2>
LSIM_HD void lens_point(const GLens& L, const double* pool, double x, double y, LensPt& o) {
  if (L.type == LSIM_LENS_EPL) {
    epl_point(L, pool, x, y, o);
  } else if (L.type == LSIM_LENS_SHEAR) {
    const double xs = x - L.c[2].v, ys = y - L.c[3].v;
    o.ax = fma(L.c[1].v, ys, L.c[0].v * xs);
    o.ay = fma(L.c[1].v, xs, -(L.c[0].v * ys));
  } else if (TIER >= 1 && L.type == kGradLensMP) {
    double hx, hy;
    mp_value(L, x, y, hx, hy);
    o.ax = L.k[0] * hx;
    o.ay = L.k[0] * hy;
  } else if (TIER >= 1 && L.type == LSIM_LENS_NIE) {
    nie_point(L, x, y, o);
  } else if (TIER < 2) {
    o.ax = 0.0;
    o.ay = 0.0;
  } else {
    LD fx, fy;
    lens_eval_d(L, pool, x, y, fx, fy);
    o.ax = fx.v;
    o.ay = fy.v;
    LSIM_UNROLL
    for (int k = 0; k < kLensN; ++k) {
      o.s[k] = fx.d[k];
      o.s[kLensN + k] = fy.d[k];
    }
  }
}

// s_j = d(vx alpha_x + vy alpha_y) / d(primitive j) of the profile.  The primitives are (theta_E, t, phi, q, center_x,
// center_y) for the EPL, the four parameters for the shear, the parameters themselves otherwise; lens_chain() maps
// sums of s over any number of points to the parameters (the map is linear and constant per parameter set, so it is
// applied once per CTA, not once per sub-pixel).
template <int TIER = 2>
LSIM_HD void lens_vjp(const GLens& L, const LensPt& o, double x, double y, double vx, double vy,
                      double (&g)[kLensEntries]) {
  g[6] = 0.0;
  g[7] = 0.0;
  if (L.type == LSIM_LENS_EPL) {
    epl_vjp(L, o, x, y, vx, vy, g);
  } else if (L.type == LSIM_LENS_SHEAR) {  // alpha = (g1 xs + g2 ys, g2 xs - g1 ys)
    const double xs = x - L.c[2].v, ys = y - L.c[3].v;
    g[0] = fma(vx, xs, -(vy * ys));
    g[1] = fma(vx, ys, vy * xs);
    g[2] = -fma(vx, L.c[0].v, vy * L.c[1].v);
    g[3] = fma(vy, L.c[0].v, -(vx * L.c[1].v));
    g[4] = 0.0;
    g[5] = 0.0;
  } else if (TIER >= 1 && L.type == kGradLensMP) {
    mp_vjp(L, x, y, vx, vy, g);
  } else if (TIER >= 1 && L.type == LSIM_LENS_NIE) {
    nie_vjp(L, o, x, y, vx, vy, g);
  } else if (TIER < 2) {
    LSIM_UNROLL
    for (int k = 0; k < kLensN; ++k) g[k] = 0.0;
  } else {
    LSIM_UNROLL
    for (int k = 0; k < kLensN; ++k) g[k] = fma(vx, o.s[k], vy * o.s[kLensN + k]);
  }
}
// Profiles the kernel does not keep a point state for: the deflection alone in the first pass ...
template <int TIER = 2>
LSIM_HD void lens_value(const GLens& L, const double* pool, double x, double y, double& ax, double& ay) {
  if (L.type == LSIM_LENS_SHEAR) {
    const double xs = x - L.c[2].v, ys = y - L.c[3].v;
    ax = fma(L.c[1].v, ys, L.c[0].v * xs);
    ay = fma(L.c[1].v, xs, -(L.c[0].v * ys));
  } else if (TIER >= 1 && L.type == kGradLensMP) {
    double hx, hy;
    mp_value(L, x, y, hx, hy);
    ax = L.k[0] * hx;
    ay = L.k[0] * hy;
  } else {
    LensPt o;
    lens_point<TIER>(L, pool, x, y, o);  // the Jacobian part is dead code here
    ax = o.ax;
    ay = o.ay;
  }
}
// ... and the contraction in the second, recomputing what it needs (nothing for the shear and the multipoles)
template <int TIER = 2>
LSIM_HD void lens_vjp_recompute(const GLens& L, const double* pool, double x, double y, double vx, double vy,
                                double (&g)[kLensEntries]) {
  LensPt o;
  o.ax = o.ay = 0.0;
  if (L.type != LSIM_LENS_SHEAR && L.type != kGradLensMP) lens_point<TIER>(L, pool, x, y, o);
  lens_vjp<TIER>(L, o, x, y, vx, vy, g);
}
// ---- multi-plane ray shooting and its reverse sweep (MultiPlaneBase.ray_shooting_partial_comoving,
// multi_plane_base.py:159-233; jax differentiates the same recursion) --------------------------------------------------
// Forward, per profile l in plane order:  X += a dT_l;  theta_l = X / T_z_l;  a -= red_l alpha_l(theta_l);
//          beta = (X + a T_end) / T_z_source      (X comoving position, a angle; a starts as the image position)
// Reverse, given v = dL/dbeta:  Xb = v / T_z_source,  ab = v T_end / T_z_source;  for l = n-1 ... 0:
//          w = -red_l ab;  parameter entries of l += d(w . alpha_l)/dp;  Xb += d(w . alpha_l)/dtheta / T_z_l;  ab += Xb dT_l.
// Every profile depends on the position through (theta - centre) only, so d/dtheta = -d/d(centre): the entries the
// parameter contraction (lens_vjp) produces anyway -- no second set of derivatives.
constexpr int kMaxGradPlanes = 8;
struct MpPlane {
  double dT, iTz, r2p;
  int32_t ix, iy;  // entries of lens_vjp's output that hold d/d(centre_x), d/d(centre_y)
};

// entries of d/d(centre) in lens_vjp's output for a PREPARED profile (closed forms work in their own primitives)
LSIM_HD void vjp_center_slots(const GLens& L, const lsim_lens_desc& d, int& ix, int& iy) {
  if (L.type == LSIM_LENS_EPL || L.type == LSIM_LENS_NIE) {
    ix = 4;
    iy = 5;
  } else if (L.type == LSIM_LENS_SHEAR || L.type == kGradLensMP) {
    ix = 2;
    iy = 3;
  } else if (!lens_center_slots(d, ix, iy)) {
    ix = iy = -1;
  }
}

// lens_prepare_d with EVERY parameter carried as an independent variable (a fixed centre still needs its tangent for
// the reverse sweep); the entries of fixed parameters are dropped when the entries are mapped to columns.
LSIM_HD bool lens_prepare_all_vars(const lsim_lens_desc& d, const double* row, GLens& L, double* F, double* dF) {
  lsim_lens_desc dd = d;
  double r2[LSIM_MAX_PARAMS];
  for (int k = 0; k < LSIM_MAX_PARAMS; ++k) {
    r2[k] = param_of(d, row, k);
    dd.col[k] = k;
  }
  return lens_prepare_d(dd, r2, L, F, dF);
}

template <int TIER>
LSIM_HD void mp_ray_forward(const GLens* L, const MpPlane* pl, int n, const double* pool, double T_end, double iTzs,
                            double x, double y, double* tx, double* ty, double& bx, double& by) {
  double X = 0.0, Y = 0.0, ax = x, ay = y;
  for (int l = 0; l < n; ++l) {
    X = fma(ax, pl[l].dT, X);
    Y = fma(ay, pl[l].dT, Y);
    tx[l] = X * pl[l].iTz;
    ty[l] = Y * pl[l].iTz;
    double fx, fy;
    lens_value<TIER>(L[l], pool, tx[l], ty[l], fx, fy);
    ax = fma(-fx, pl[l].r2p, ax);
    ay = fma(-fy, pl[l].r2p, ay);
  }
  bx = fma(ax, T_end, X) * iTzs;
  by = fma(ay, T_end, Y) * iTzs;
}

// acc(l, g): add the kLensEntries entries g of profile l
template <int TIER, class Acc>
LSIM_HD void mp_ray_reverse(const GLens* L, const MpPlane* pl, int n, const double* pool, double T_end, double iTzs,
                            const double* tx, const double* ty, double vx, double vy, Acc&& acc) {
  double Xbx = vx * iTzs, Xby = vy * iTzs, abx = vx * T_end * iTzs, aby = vy * T_end * iTzs;
  for (int l = n - 1; l >= 0; --l) {
    double g[kLensEntries];
    lens_vjp_recompute<TIER>(L[l], pool, tx[l], ty[l], -pl[l].r2p * abx, -pl[l].r2p * aby, g);
    acc(l, g);
    if (pl[l].ix >= 0) {
      Xbx = fma(-g[pl[l].ix], pl[l].iTz, Xbx);
      Xby = fma(-g[pl[l].iy], pl[l].iTz, Xby);
    }
    abx = fma(Xbx, pl[l].dT, abx);
    aby = fma(Xby, pl[l].dT, aby);
  }
}

LSIM_HD int lens_grad_tier(const lsim_lens_desc& d) {
  if (d.type == LSIM_LENS_NULL || d.type == LSIM_LENS_SHEAR || lens_grad_has_series(d)) return 0;
  if (d.type == kGradLensMP || d.type == LSIM_LENS_MULTIPOLE || d.type == LSIM_LENS_NIE || d.type == LSIM_LENS_SIE) return 1;
  return 2;
}
LSIM_HD bool lens_grad_closed_form(const lsim_lens_desc& d) {
  return d.type == LSIM_LENS_NULL || d.type == LSIM_LENS_SHEAR || d.type == kGradLensMP ||
         d.type == LSIM_LENS_MULTIPOLE || d.type == LSIM_LENS_NIE || d.type == LSIM_LENS_SIE ||
         lens_grad_has_series(d);
}

LSIM_HD void lens_chain(const GLens& L, const double (&s)[kLensEntries], double (&g)[kLensEntries]) {
  for (int k = 0; k < kLensEntries; ++k) {
    if (k < kLensN && L.type == LSIM_LENS_EPL) {
      double a = 0.0;
      for (int j = 0; j < 6; ++j) a = fma(s[j], L.p[j].d[k], a);
      g[k] = a;
    } else if (k < kLensN && L.type == LSIM_LENS_SHEAR) {
      double a = 0.0;
      for (int j = 0; j < 4; ++j) a = fma(s[j], L.c[j].d[k], a);
      g[k] = a;
    } else if (k < kLensN && L.type == LSIM_LENS_NIE) {  // primitives (b, q, s, phi, cx, cy); d phi = cos d(sin) - sin d(cos)
      const double dphi = fma(L.c[2].v, L.c[3].d[k], -(L.c[3].v * L.c[2].d[k]));
      g[k] = fma(s[0], L.c[5].d[k], fma(s[1], L.c[4].d[k], fma(s[2], L.c[6].d[k], fma(s[3], dphi,
                                                                 fma(s[4], L.c[0].d[k], s[5] * L.c[1].d[k])))));
    } else if (k < kLensN && L.type == kGradLensMP) {  // primitives (rho, phi_m, cx, cy, r_E) = c[2], c[3], c[0], c[1], c[4]
      g[k] = fma(s[0], L.c[2].d[k], fma(s[1], L.c[3].d[k], fma(s[2], L.c[0].d[k], fma(s[3], L.c[1].d[k],
                                                                                       s[4] * L.c[4].d[k]))));
    } else {
      g[k] = s[k];
    }
  }
}

// ---- light, hand-derived -------------------------------------------------------------------------------------------
constexpr int kLightG = 7;  // own parameters of a light profile (SERSIC_ELLIPSE has 7)
struct GLight {
  int32_t type, lensed, slot, pad;
  double k[20];
};

// SERSIC / SERSIC_ELLIPSE: k = cx, cy, a11, a12, a22 (the product-average transform), 1/n, log2 Rs_, -bn log2e,
// min(bn log2e, 1021), (1000 Rs_)^2, amp, bn, [Rs > 1e-4]/Rs_, 1.9992 [bn not clamped], 1/norm, dnorm/de1 / norm,
// dnorm/de2 / norm, [amp is a parameter]
LSIM_HD void light_prepare_g(const lsim_light_desc& d, const double* row, GLight& G) {
  G.type = d.type;
  G.lensed = d.lensed;
  G.slot = d.slot;
  G.pad = 0;
  auto P = [&](int i) { return param_of(d, row, i); };
  const double amp = d.use_amp ? P(0) : 1.0;
  if (d.type == LSIM_LIGHT_SERSIC || d.type == LSIM_LIGHT_SERSIC_ELLIPSE) {
    const bool ell = d.type == LSIM_LIGHT_SERSIC_ELLIPSE;
    double Rs = P(1);
    const double n = P(2);
    const double e1 = ell ? P(3) : 0.0, e2 = ell ? P(4) : 0.0;
    bool rs_live = true;
    if (ell) {
      rs_live = !(Rs < 0.0);
      Rs = fmax(0.0, Rs);
    }
    rs_live = rs_live && !(Rs < 0.0001);
    const double arg = 1 - e1 * e1 - e2 * e2;
    const double ne = sqrt(fabs(arg));
    const bool n_live = !(ne < 0.000001);
    const double norm = fmax(ne, 0.000001);
    const double Rs_ = fmax(0.0001, Rs);
    const double bnr = 1.9992 * n - 0.3271;
    const double bn = fmax(bnr, 0.00001);
    G.k[0] = P(ell ? 5 : 3);
    G.k[1] = P(ell ? 6 : 4);
    G.k[2] = (1 - e1) / norm;
    G.k[3] = -e2 / norm;
    G.k[4] = (1 + e1) / norm;
    G.k[5] = 1.0 / n;
    G.k[6] = log2(Rs_);
    G.k[7] = -bn * kLog2e;
    G.k[8] = fmin(bn * kLog2e, 1021.0);
    G.k[9] = (1000.0 * Rs_) * (1000.0 * Rs_);
    G.k[10] = amp;
    G.k[11] = bn;
    G.k[12] = rs_live ? 1.0 / Rs_ : 0.0;
    G.k[13] = (bnr < 0.00001) ? 0.0 : 1.9992;
    G.k[14] = 1.0 / norm;
    const double sgn = arg < 0 ? -1.0 : 1.0;
    G.k[15] = (n_live && ne > 0) ? -e1 * sgn / (ne * norm) : 0.0;
    G.k[16] = (n_live && ne > 0) ? -e2 * sgn / (ne * norm) : 0.0;
    G.type = ell ? LSIM_LIGHT_SERSIC_ELLIPSE : LSIM_LIGHT_SERSIC;
  } else if (d.type == LSIM_LIGHT_GAUSSIAN) {
    const double sigma = P(1);
    G.k[0] = P(2);
    G.k[1] = P(3);
    G.k[2] = -0.5 * kLog2e / (sigma * sigma);
    G.k[3] = 1.0 / (2 * kPi * sigma * sigma);
    G.k[4] = 1.0 / (sigma * sigma);
    G.k[5] = 1.0 / sigma;
    G.k[10] = amp;
  } else {
    G.slot = -1;
  }
}

// I, dI/dx, dI/dy and dI/d(own parameter k) (param_names order) of one light component at (x, y)
LSIM_HD void light_point(const GLight& G, double x, double y, double& I, double& Ix, double& Iy, double (&g)[kLightG]) {
  LSIM_UNROLL
  for (int k = 0; k < kLightG; ++k) g[k] = 0.0;
  I = Ix = Iy = 0.0;
  if (G.type == LSIM_LIGHT_GAUSSIAN) {
    const double dx = x - G.k[0], dy = y - G.k[1];
    const double r2 = fma(dy, dy, dx * dx);
    const double E = G.k[3] * exp2_fin(G.k[2] * r2);
    I = G.k[10] * E;
    const double w = -I * G.k[4];
    Ix = w * dx;
    Iy = w * dy;
    g[0] = E;
    g[1] = I * G.k[5] * fma(r2, G.k[4], -2.0);
    g[2] = -Ix;
    g[3] = -Iy;
    return;
  }
  const bool ell = G.type == LSIM_LIGHT_SERSIC_ELLIPSE;
  const double xs = x - G.k[0], ys = y - G.k[1];
  const double x_ = fma(G.k[3], ys, G.k[2] * xs), y_ = fma(G.k[4], ys, G.k[3] * xs);
  const double R2 = fma(y_, y_, x_ * x_);
  const bool r_live = !(R2 < 1e-8);
  const double R2c = max_nan(R2, 1e-8);
  if (!(R2c <= G.k[9])) return;  // R_frac > max_R_frac: identically zero
  const double lu = fma(0.5, log2_pos(R2c), -G.k[6]);  // log2(R_/Rs_)
  const double p = exp2_fin(lu * G.k[5]);
  const double E = exp2_fin(fma(G.k[7], p, G.k[8]));
  const double amp = G.k[10], bn = G.k[11], inv_n = G.k[5];
  I = amp * E;
  const double bEp = bn * E * p * inv_n * amp;           // amp bn E p / n
  const double W = r_live ? -bEp * rcp_fast(R2c) : 0.0;  // dI/dR / R
  const double gx = W * x_, gy = W * y_;
  Ix = fma(gy, G.k[3], gx * G.k[2]);
  Iy = fma(gy, G.k[4], gx * G.k[3]);
  g[0] = E;
  g[1] = bEp * G.k[12];
  g[2] = I * fma(bn * p * inv_n * inv_n, lu * kLn2, -(G.k[13] * (p - 1.0)));
  if (ell) {
    const double WR2 = W * R2;
    g[3] = fma(gy, ys, -(gx * xs)) * G.k[14] - WR2 * G.k[15];
    g[4] = -fma(gx, ys, gy * xs) * G.k[14] - WR2 * G.k[16];
    g[5] = -Ix;
    g[6] = -Iy;
  } else {
    g[3] = -Ix;
    g[4] = -Iy;
  }
}

}  // namespace ad
}  // namespace lsim

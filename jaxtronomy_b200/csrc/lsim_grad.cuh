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
constexpr int kTanTerms = 100;                      // hyp2f1_slow, epl.py:283
constexpr int kGradPool = kMaxSeries + kTanTerms;   // per EPL-family profile: F_n [kMaxSeries], dF_n/dt [kTanTerms]

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
  LD c[8];
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
      return true;
    default:
      return false;
  }
}
LSIM_HD bool lens_grad_has_series(const lsim_lens_desc& d) {
  return d.type == LSIM_LENS_EPL || (d.type == LSIM_LENS_SIE && (d.flags & LSIM_FLAG_SIE_NIE_FALSE));
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

}  // namespace ad
}  // namespace lsim

// tests/hostmath -- TEST HARNESS ONLY.  Compiles the product's host+device profile arithmetic
// (jaxtronomy_b200/csrc/lsim_profiles.cuh) with g++ so that the exact per-point code the CUDA kernels run
// can be checked against the oracle on a machine without a GPU.  Not linked into, or loaded by, the product.
#include <limits>
#include <vector>

#include "../../jaxtronomy_b200/csrc/lsim_profiles.cuh"
#include "../../jaxtronomy_b200/csrc/lsim_grad.cuh"

// coefficient pools start out as NaN: a coefficient that is read without having been written (the device reads stale
// shared memory in that case, invisible unless the stale value is huge or non-finite) poisons the result here
static const double kPoison = std::numeric_limits<double>::quiet_NaN();

using namespace lsim;

extern "C" {

int hm_ray_shoot(const lsim_lens_desc* lens, int n_lens, int multi_plane, double alpha_scaling, double T_ij_end,
                 double T_z_source, const double* row, const double* x, const double* y, int64_t N, double* ox,
                 double* oy, int what, double series_tol) {
  std::vector<LensC> L(n_lens);
  std::vector<double> pool((size_t)(n_lens + 1) * kMaxSeries, kPoison);
  int off = 0;
  for (int i = 0; i < n_lens; ++i) {
    L[i].coef_off = off * kMaxSeries;
    lens_prepare(lens[i], row, L[i], pool.data() + (size_t)off * kMaxSeries, series_tol);
    if (lens_type_has_series(lens[i])) ++off;
  }
  RayCfg cfg{multi_plane, multi_plane ? 1.0 : alpha_scaling, T_ij_end, multi_plane ? 1.0 / T_z_source : 0.0, math_k()};
  // points go through the same J-batched code path the kernels use (J = 3), tail with J = 1
  int64_t i = 0;
  for (; i + 3 <= N; i += 3) {
    Rays<3> r;
    for (int j = 0; j < 3; ++j) ray_init(cfg, j, x[i + j], y[i + j], r);
    ray_apply<true, 3>(cfg, L.data(), pool.data(), n_lens, r);
    for (int j = 0; j < 3; ++j) {
      if (what == 1 && !multi_plane) {
        ox[i + j] = r.ax[j] * cfg.alpha_scaling;
        oy[i + j] = r.ay[j] * cfg.alpha_scaling;
      } else {
        ray_finish(cfg, r, j, ox[i + j], oy[i + j]);
      }
    }
  }
  for (; i < N; ++i) {
    Rays<1> r;
    ray_init(cfg, 0, x[i], y[i], r);
    ray_apply<true, 1>(cfg, L.data(), pool.data(), n_lens, r);
    if (what == 1 && !multi_plane) {
      ox[i] = r.ax[0] * cfg.alpha_scaling;
      oy[i] = r.ay[0] * cfg.alpha_scaling;
    } else {
      ray_finish(cfg, r, 0, ox[i], oy[i]);
    }
  }
  return 0;
}

// The compile-time profile-list path of the specialised image kernels (k_image<..., TypeList<EPL, SHEAR>, ...>):
// ray_apply_head / re-read coordinates / ray_apply_tail / beta, three points in lock step, tail padded.
int hm_ray_shoot_epl_shear(const lsim_lens_desc* lens, double alpha_scaling, const double* row, const double* x,
                           const double* y, int64_t N, double* ox, double* oy, double series_tol) {
  using LT = TypeList<LSIM_LENS_EPL, LSIM_LENS_SHEAR>;
  LensC L[2];
  std::vector<double> pool(2 * kMaxSeries, kPoison);
  L[0].coef_off = 0;
  lens_prepare(lens[0], row, L[0], pool.data(), series_tol);
  L[1].coef_off = kMaxSeries;
  lens_prepare(lens[1], row, L[1], pool.data() + kMaxSeries, series_tol);
  if (L[0].type != LSIM_LENS_EPL || L[1].type != LSIM_LENS_SHEAR) return 1;
  RayCfg cfg{0, alpha_scaling, 0.0, 0.0, math_k()};
  for (int64_t i = 0; i < N; i += 3) {
    Rays<3> r;
    for (int j = 0; j < 3; ++j) {
      const int64_t k = i + j < N ? i + j : N - 1;
      ray_init<3, true>(cfg, j, x[k], y[k], r);
    }
    ray_apply_head<false, 3, LT>(cfg, L, pool.data(), r);
    ray_apply_tail<false, 3, LT>(cfg, L, pool.data(), r);
    for (int j = 0; j < 3 && i + j < N; ++j) {
      ox[i + j] = fma(-r.ax[j], cfg.alpha_scaling, r.x[j]);
      oy[i + j] = fma(-r.ay[j], cfg.alpha_scaling, r.y[j]);
    }
  }
  return 0;
}

int hm_surface_brightness(const lsim_light_desc* light, int n_light, int n_slots, const double* row, const double* x,
                          const double* y, int64_t N, double* out) {
  std::vector<LightC> G(n_light);
  for (int i = 0; i < n_light; ++i) light_prepare(light[i], row, G[i]);
  for (int s = 0; s < n_slots; ++s)
    for (int64_t i = 0; i < N; ++i) {
      double acc = 0;
      for (int g = 0; g < n_light; ++g)
        if (G[g].slot == s) acc += light_eval(G[g], x[i], y[i]);
      out[(size_t)s * N + i] = acc;
    }
  return 0;
}

void hm_log2(const double* x, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = log2_pos(x[i]);
}
void hm_exp2(const double* x, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = exp2_fin(x[i]);
}
void hm_exp2_s(const double* y, int64_t n, double* out) {
  const MathK K = math_k();
  for (int64_t i = 0; i < n; ++i) out[i] = exp2_fin_s(y[i], K);
}
void hm_exp2_neg(const double* x, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = exp2_neg(x[i]);
}
int hm_epl_terms(const lsim_lens_desc* lens, const double* row, double series_tol) {
  LensC L;
  std::vector<double> pool(kMaxSeries, kPoison);
  lens_prepare(*lens, row, L, pool.data(), series_tol);
  return L.n_coef;
}

// Dual-number profile arithmetic of the gradient path (lsim_grad.cuh).
//   ax, ay   [N]: summed deflection;  dax, day [N][n_lens][6]: d(alpha_i)/d(parameter k of profile i)
int hm_lens_jvp(const lsim_lens_desc* lens, int n_lens, const double* row, const double* x, const double* y, int64_t N,
                double* ax, double* ay, double* dax, double* day) {
  std::vector<ad::GLens> L(n_lens);
  std::vector<double> pool((size_t)(n_lens + 1) * ad::kGradPool, kPoison);
  int off = 0;
  for (int i = 0; i < n_lens; ++i) {
    L[i].coef_off = off * ad::kGradPool;
    double* F = pool.data() + (size_t)off * ad::kGradPool;
    if (!ad::lens_prepare_d(lens[i], row, L[i], F, F + kMaxSeries)) return -2;
    if (ad::lens_grad_has_series(lens[i])) ++off;
  }
  for (int64_t p = 0; p < N; ++p) {
    ax[p] = ay[p] = 0.0;
    for (int i = 0; i < n_lens; ++i) {
      ad::LD fx, fy;
      ad::lens_eval_d(L[i], pool.data(), x[p], y[p], fx, fy);
      ax[p] += fx.v;
      ay[p] += fy.v;
      for (int k = 0; k < ad::kLensN; ++k) {
        dax[((size_t)p * n_lens + i) * ad::kLensN + k] = fx.d[k];
        day[((size_t)p * n_lens + i) * ad::kLensN + k] = fy.d[k];
      }
    }
  }
  return 0;
}

//   out [N], dout [N][9]: d/dx, d/dy, then d/d(parameter k)
int hm_light_jvp(const lsim_light_desc* light, const double* row, const double* x, const double* y, int64_t N,
                 double* out, double* dout) {
  if (!ad::light_grad_supported(*light)) return -2;
  for (int64_t p = 0; p < N; ++p) {
    const ad::GD v = ad::light_eval_d(*light, row, ad::var<ad::kLightN>(x[p], 0), ad::var<ad::kLightN>(y[p], 1));
    out[p] = v.v;
    for (int k = 0; k < ad::kLightN; ++k) dout[(size_t)p * ad::kLightN + k] = v.d[k];
  }
  return 0;
}

// Hessian of the lens potential through the kernel's own per-point code (k_hessian): the descriptor list is expanded
// like lsim_plan_create does.   h [4][N] = f_xx, f_xy, f_yx, f_yy (unscaled)
int hm_lens_hessian(const lsim_lens_desc* lens, int n_lens, const double* row, const double* x, const double* y,
                    int64_t N, double* h) {
  std::vector<lsim_lens_desc> gl((size_t)4 * (n_lens > 0 ? n_lens : 1));
  const int ngl = ad::expand_grad_lens(lens, n_lens, gl.data(), (int)gl.size());
  if (ngl < 0) return -2;
  for (int i = 0; i < n_lens; ++i)
    if (!ad::lens_grad_supported(lens[i])) return -2;
  std::vector<ad::GLens> L(ngl > 0 ? ngl : 1);
  std::vector<int> slot(2 * (size_t)(ngl > 0 ? ngl : 1));
  std::vector<double> pool((size_t)(ngl + 1) * ad::kGradPool, kPoison);
  int off = 0;
  for (int i = 0; i < ngl; ++i) {
    L[i].coef_off = off * ad::kGradPool;
    double* F = pool.data() + (size_t)off * ad::kGradPool;
    if (!ad::lens_prepare_hess(gl[i], row, L[i], F, F + kMaxSeries, 1e-14)) return -2;
    if (ad::lens_grad_has_series(gl[i])) ++off;
    if (!ad::lens_center_slots(gl[i], slot[2 * i], slot[2 * i + 1])) slot[2 * i] = slot[2 * i + 1] = -1;
  }
  for (int64_t p = 0; p < N; ++p) {
    double hh[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < ngl; ++i)
      if (slot[2 * i] >= 0) ad::lens_hessian_add(L[i], pool.data(), slot[2 * i], slot[2 * i + 1], x[p], y[p], hh);
    for (int k = 0; k < 4; ++k) h[(size_t)k * N + p] = hh[k];
  }
  return 0;
}

// k_source_position's per-set arithmetic on the host (same helper functions, same order of the sums).
int hm_source_position_logl(const lsim_lens_desc* lens, int n_lens, const double* row, const double* x, const double* y,
                            int n_img, double scale, double sigma, double hard_bound_rms, double* logl) {
  std::vector<lsim_lens_desc> gl((size_t)4 * (n_lens > 0 ? n_lens : 1));
  const int ngl = ad::expand_grad_lens(lens, n_lens, gl.data(), (int)gl.size());
  if (ngl < 0) return -2;
  std::vector<ad::GLens> L(ngl > 0 ? ngl : 1);
  std::vector<int> slot(2 * (size_t)(ngl > 0 ? ngl : 1));
  std::vector<double> pool((size_t)(ngl + 1) * ad::kGradPool, kPoison);
  int off = 0;
  for (int i = 0; i < ngl; ++i) {
    L[i].coef_off = off * ad::kGradPool;
    double* F = pool.data() + (size_t)off * ad::kGradPool;
    if (!ad::lens_prepare_hess(gl[i], row, L[i], F, F + kMaxSeries, 1e-14)) return -2;
    if (ad::lens_grad_has_series(gl[i])) ++off;
    if (!ad::lens_center_slots(gl[i], slot[2 * i], slot[2 * i + 1])) slot[2 * i] = slot[2 * i + 1] = -1;
  }
  std::vector<double> w((size_t)6 * n_img);
  for (int i = 0; i < n_img; ++i) {
    double ax = 0.0, ay = 0.0, h[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < ngl; ++l)
      ad::lens_value_hessian_add(L[l], pool.data(), slot[2 * l], slot[2 * l + 1], x[i], y[i], ax, ay, h);
    w[i] = x[i] - ax * scale;
    w[n_img + i] = y[i] - ay * scale;
    for (int k = 0; k < 4; ++k) w[(2 + k) * n_img + i] = h[k] * scale;
  }
  double sx = 0.0, sy = 0.0;
  for (int i = 0; i < n_img; ++i) {
    sx += w[i];
    sy += w[n_img + i];
  }
  const double mx = sx / n_img, my = sy / n_img;
  double s = 0.0;
  for (int i = 0; i < n_img; ++i) {
    const double h[4] = {w[2 * n_img + i], w[3 * n_img + i], w[4 * n_img + i], w[5 * n_img + i]};
    s += ad::source_position_penalty(h, mx - w[i], my - w[n_img + i], sigma, hard_bound_rms);
  }
  *logl = -s;
  return 0;
}

// Hand-derived fast path of the gradient kernel: g [N][n_lens][6] = d(vx alpha_x + vy alpha_y)/d(parameter k).
int hm_lens_vjp(const lsim_lens_desc* lens, int n_lens, const double* row, const double* x, const double* y, int64_t N,
                const double* vx, const double* vy, double* ax, double* ay, double* g) {
  std::vector<ad::GLens> L(n_lens);
  std::vector<double> pool((size_t)(n_lens + 1) * ad::kGradPool, kPoison);
  int off = 0;
  for (int i = 0; i < n_lens; ++i) {
    L[i].coef_off = off * ad::kGradPool;
    double* F = pool.data() + (size_t)off * ad::kGradPool;
    if (!ad::lens_prepare_d(lens[i], row, L[i], F, F + kMaxSeries)) return -2;
    if (ad::lens_grad_has_series(lens[i])) ++off;
  }
  for (int64_t p = 0; p < N; ++p) {
    ax[p] = ay[p] = 0.0;
    for (int i = 0; i < n_lens; ++i) {
      ad::LensPt o;
      ad::lens_point(L[i], pool.data(), x[p], y[p], o);
      ax[p] += o.ax;
      ay[p] += o.ay;
      double ss[ad::kLensEntries], gg[ad::kLensEntries];
      ad::lens_vjp(L[i], o, x[p], y[p], vx[p], vy[p], ss);
      ad::lens_chain(L[i], ss, gg);
      for (int k = 0; k < ad::kLensN; ++k) g[((size_t)p * n_lens + i) * ad::kLensN + k] = gg[k];
    }
  }
  return 0;
}

// Multi-plane ray shooting with its reverse sweep (the per-point code of k_image_grad<0, 2, true>): beta at N image
// positions and g[p][l][k] = d(vx beta_x + vy beta_y)/d(parameter k of profile l) (descriptor order, after lens_chain).
int hm_mp_ray_vjp(const lsim_lens_desc* lens, int n_lens, double T_ij_end, double T_z_source, const double* row,
                  const double* x, const double* y, int64_t N, const double* vx, const double* vy, double* bx, double* by,
                  double* g) {
  if (n_lens > ad::kMaxGradPlanes) return -1;
  std::vector<ad::GLens> L(n_lens);
  std::vector<ad::MpPlane> pl(n_lens);
  std::vector<double> pool((size_t)(n_lens + 1) * ad::kGradPool, kPoison);
  int off = 0;
  for (int i = 0; i < n_lens; ++i) {
    L[i].coef_off = off * ad::kGradPool;
    double* F = pool.data() + (size_t)off * ad::kGradPool;
    if (!ad::lens_prepare_all_vars(lens[i], row, L[i], F, F + kMaxSeries)) return -2;
    if (ad::lens_grad_has_series(lens[i])) ++off;
    pl[i].dT = lens[i].delta_T;
    pl[i].iTz = 1.0 / lens[i].T_z;
    pl[i].r2p = lens[i].red2phys;
    int ix, iy;
    ad::vjp_center_slots(L[i], lens[i], ix, iy);
    pl[i].ix = ix;
    pl[i].iy = iy;
  }
  for (int64_t p = 0; p < N; ++p) {
    double tx[ad::kMaxGradPlanes], ty[ad::kMaxGradPlanes];
    ad::mp_ray_forward<2>(L.data(), pl.data(), n_lens, pool.data(), T_ij_end, 1.0 / T_z_source, x[p], y[p], tx, ty, bx[p],
                          by[p]);
    ad::mp_ray_reverse<2>(L.data(), pl.data(), n_lens, pool.data(), T_ij_end, 1.0 / T_z_source, tx, ty, vx[p], vy[p],
                          [&](int l, const double (&s)[ad::kLensEntries]) {
                            double gg[ad::kLensEntries];
                            ad::lens_chain(L[l], s, gg);
                            for (int k = 0; k < ad::kLensN; ++k) g[((size_t)p * n_lens + l) * ad::kLensN + k] = gg[k];
                          });
  }
  return 0;
}

//   out [N], dout [N][9]: d/dx, d/dy, then d/d(parameter k) -- same layout as hm_light_jvp
int hm_light_point(const lsim_light_desc* light, const double* row, const double* x, const double* y, int64_t N,
                   double* out, double* dout) {
  ad::GLight G;
  ad::light_prepare_g(*light, row, G);
  if (G.slot < 0) return -2;
  for (int64_t p = 0; p < N; ++p) {
    double I, Ix, Iy, g[ad::kLightG];
    ad::light_point(G, x[p], y[p], I, Ix, Iy, g);
    out[p] = I;
    dout[(size_t)p * 9] = Ix;
    dout[(size_t)p * 9 + 1] = Iy;
    for (int k = 0; k < ad::kLightG; ++k) dout[(size_t)p * 9 + 2 + k] = g[k];
  }
  return 0;
}

// EPL_MULTIPOLE_M1M3M4 through the gradient path's expansion (EPL + three multipole pseudo-profiles):
//   ax, ay [N]; g [N][12] = d(vx alpha_x + vy alpha_y)/d(parameter k), param_names order
int hm_eplm_vjp(const lsim_lens_desc* lens, const double* row, const double* x, const double* y, int64_t N,
                const double* vx, const double* vy, double* ax, double* ay, double* g) {
  lsim_lens_desc ex[4];
  const int n = ad::expand_grad_lens(lens, 1, ex, 4);
  if (n != 4) return -2;
  std::vector<ad::GLens> L(n);
  std::vector<double> pool((size_t)2 * ad::kGradPool, kPoison);
  for (int i = 0; i < n; ++i) {
    L[i].coef_off = 0;
    if (!ad::lens_prepare_d(ex[i], row, L[i], pool.data(), pool.data() + kMaxSeries)) return -2;
  }
  for (int64_t p = 0; p < N; ++p) {
    ax[p] = ay[p] = 0.0;
    for (int k = 0; k < 12; ++k) g[p * 12 + k] = 0.0;
    for (int i = 0; i < n; ++i) {
      ad::LensPt o;
      ad::lens_point(L[i], pool.data(), x[p], y[p], o);
      ax[p] += o.ax;
      ay[p] += o.ay;
      double ss[ad::kLensEntries], gg[ad::kLensEntries];
      ad::lens_vjp(L[i], o, x[p], y[p], vx[p], vy[p], ss);
      ad::lens_chain(L[i], ss, gg);
      for (int k = 0; k < ad::kLensEntries; ++k) {
        const int c = ex[i].col[k];  // the harness gives every parent parameter its own column = its index
        if (k < (i == 0 ? 6 : 7) && c >= 0) g[p * 12 + c] += gg[k];
      }
    }
  }
  return 0;
}
}

// lsim_grad_kernels.cuh -- kernels of lsim_logl_grad: logL and d logL / d params for a batch of parameter sets
// (the reference gets this from jax.value_and_grad, Sampling/Samplers/optax.py:39,91-96).
//
//   k_like_adjoint   logL, WLS amplitudes and Abar = d logL / d C (C: convolved, scaled basis images) per set
//   k_conv_adj_edge  transposed edge-replicated separable convolution (GaussianConvolution) -- the zero-boundary
//                    transposes run on the forward convolution kernels with the flipped PSF
//   k_image_grad     sum over pixels of Ubar * dU/dp: dual-number evaluation of every profile per sub-pixel
//                    (lsim_grad.cuh), per-thread partial sums in shared memory, one deterministic partial per CTA
//   k_grad_reduce    partials -> grad[B, P] in a fixed order (bit-reproducible run to run; no atomics)
#pragma once

#include "lsim_grad.cuh"
#include "lsim_kernels.cuh"

namespace lsim {

constexpr int kGradBlock = 128;
#ifndef LSIM_GRAD_FAST_BLOCKS
#define LSIM_GRAD_FAST_BLOCKS 3  // resident CTAs per SM asked of the closed-form instantiation (register budget 65536 / (128 x this))
#endif
constexpr int kGradLightStride = 7;  // entries per light component: its own parameters (SERSIC_ELLIPSE has 7)

// n_lens: (pseudo-)profiles of the gradient's own lens list (EPL_MULTIPOLE_M1M3M4 expands into four)
__host__ __device__ inline int grad_entries(int n_lens, int n_light) {
  return ad::kLensEntries * n_lens + kGradLightStride * n_light;
}

__host__ __device__ inline int grad_lens_params(int type) {
  switch (type) {
    case LSIM_LENS_EPL: return 6;
    case LSIM_LENS_SIE: return 5;
    case LSIM_LENS_NIE: return 6;
    case LSIM_LENS_SIS: return 3;
    case LSIM_LENS_SHEAR: return 4;
    case LSIM_LENS_CONVERGENCE: return 3;
    case ad::kGradLensMP: return 7;
    case LSIM_LENS_NFW: return 4;
    case LSIM_LENS_TNFW: return 5;
    case LSIM_LENS_MULTIPOLE:
    case LSIM_LENS_CSE:
    case LSIM_LENS_NFW_ELLIPSE_CSE:
    case LSIM_LENS_HERNQUIST_ELLIPSE_CSE: return 6;
    default: return 0;
  }
}
__host__ __device__ inline int grad_light_params(int type) {
  switch (type) {
    case LSIM_LIGHT_SERSIC: return 5;
    case LSIM_LIGHT_SERSIC_ELLIPSE: return 7;
    case LSIM_LIGHT_GAUSSIAN: return 4;
    default: return 0;
  }
}

// ---- k_like_adjoint --------------------------------------------------------------------------------------------------
// One CTA per parameter set; same arithmetic as k_likelihood for logL and the amplitudes, then the adjoint.
//   m = sum_k a_k A_k (linear: a = M^-1 R, M = A W A^T, R = A W d) or sum_k A_k;  r = m - d
//   C_D(m) = noise_map^2 or max(m,0)/t_exp + sigma_bkg^2;  logL = -1/2 sum mask r^2 / C_D(m)
//   g = d logL / d m = -mask (r / C_D - r^2 C_D'(m) / (2 C_D^2)),  C_D' = [m > 0] / t_exp (0 with a noise map)
//   linear:  h = M^-1 (A g),  q = sum_k h_k A_k,  Abar_k = a_k (g - W q) - h_k W r      (W, g zero off the mask)
//   else:    Abar_k = g
// A set whose normal matrix fails the reference's stability gate (amplitudes forced to 0, de_lens.py:91-105) has a
// constant model: Abar = 0.
template <int NS>
__global__ void __launch_bounds__(kBlock)
k_like_adjoint(PlanDev P, const double* __restrict__ conv, int64_t set0, int linear, double* __restrict__ logl,
               double* __restrict__ amps, double* __restrict__ abar) {
  constexpr int NM = NS * (NS + 1) / 2;
  __shared__ double scratch[(NM + NS + 1) * (kBlock / 32)];
  __shared__ double s_amp[NS], s_h[NS], s_M[NS * NS];
  __shared__ int s_ok;
  const int64_t bl = blockIdx.x;
  const int npix = P.nx * P.ny;
  const int ns = P.n_slots;
  const double* A = conv + bl * ns * (int64_t)npix;
  double* out = abar + bl * ns * (int64_t)npix;

  if (linear) {
    double acc[NM + NS];
#pragma unroll
    for (int i = 0; i < NM + NS; ++i) acc[i] = 0.0;
    for (int p = threadIdx.x; p < npix; p += kBlock) {
      const bool use = !P.mask || P.mask[p];
      if (!use) continue;
      const double w = __ldg(P.w_data + p), d = __ldg(P.data + p);
      double a[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) a[k] = k < ns ? nan_to_num(A[(int64_t)k * npix + p]) : 0.0;
      int m = 0;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
#pragma unroll
        for (int j = i; j < NS; ++j) acc[m++] += w * a[i] * a[j];
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) acc[NM + i] += w * d * a[i];
    }
    block_sum(acc, scratch);
    if (threadIdx.x == 0) {
      double M2[NS * NS], Mw[NS * NS], R[NS], ev[NS];
      int m = 0;
      double Mf[NS * NS];
      for (int i = 0; i < NS; ++i)
        for (int j = i; j < NS; ++j) {
          Mf[i * NS + j] = acc[m];
          Mf[j * NS + i] = acc[m];
          ++m;
        }
      for (int i = 0; i < ns; ++i)
        for (int j = 0; j < ns; ++j) M2[i * ns + j] = Mf[i * NS + j];
      for (int i = 0; i < ns * ns; ++i) {
        Mw[i] = M2[i];
        s_M[i] = M2[i];
      }
      for (int i = 0; i < ns; ++i) R[i] = acc[NM + i];
      sym_eigenvalues(Mw, ns, ev);
      double emax = 0.0, emin = HUGE_VAL;
      for (int i = 0; i < ns; ++i) {
        emax = fmax(emax, fabs(ev[i]));
        emin = fmin(emin, fabs(ev[i]));
      }
      bool ok = (emax / emin) < 5.0 / DBL_EPSILON && lu_solve(M2, R, ns);
      if (ok)
        for (int i = 0; i < ns; ++i) ok = ok && (fabs(R[i]) <= DBL_MAX);
      for (int i = 0; i < NS; ++i) s_amp[i] = (ok && i < ns) ? R[i] : 0.0;
      s_ok = ok ? 1 : 0;
    }
    __syncthreads();
  }

  // model, residual and d logL / d model of one pixel
  auto pixel = [&](int p, bool use, double (&a)[NS], double& r, double& g, double& chi) {
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      a[k] = 0.0;
      if (k < ns) {
        if (linear) {
          if (use) {
            a[k] = nan_to_num(A[(int64_t)k * npix + p]);
            m += s_amp[k] * a[k];
          }
        } else {
          a[k] = A[(int64_t)k * npix + p];
          m += a[k];
        }
      }
    }
    r = 0.0;
    g = 0.0;
    chi = 0.0;
    if (use) {
      double cd, cdp = 0.0;
      if (P.noise_map2) {
        cd = __ldg(P.noise_map2 + p);
      } else {
        const double ie = __ldg(P.inv_exposure + p);
        cd = fma(fmax(m, 0.0), ie, P.bkg2);
        cdp = m > 0.0 ? ie : 0.0;
      }
      r = m - __ldg(P.data + p);
      const double rc = r / cd;
      chi = r * rc;
      g = -(rc - 0.5 * rc * rc * cdp);
    }
  };

  double red[NS + 1];
#pragma unroll
  for (int i = 0; i < NS + 1; ++i) red[i] = 0.0;
  for (int p = threadIdx.x; p < npix; p += kBlock) {
    const bool use = !P.mask || P.mask[p];
    double a[NS], r, g, chi;
    pixel(p, use, a, r, g, chi);
    red[0] += chi;
#pragma unroll
    for (int k = 0; k < NS; ++k) red[1 + k] += g * a[k];
  }
  block_sum(red, scratch);
  if (threadIdx.x == 0) {
    logl[set0 + bl] = -red[0] / 2 + P.logl_offset;
    if (amps && linear)
      for (int k = 0; k < ns; ++k) amps[(set0 + bl) * ns + k] = s_amp[k];
    if (linear) {
      double M2[NS * NS], c[NS];
      for (int i = 0; i < ns * ns; ++i) M2[i] = s_M[i];
      for (int i = 0; i < ns; ++i) c[i] = red[1 + i];
      const bool ok = s_ok && lu_solve(M2, c, ns);
      for (int i = 0; i < NS; ++i) s_h[i] = (ok && i < ns) ? c[i] : 0.0;
    }
  }
  __syncthreads();
  const bool dead = linear && !s_ok;
  for (int p = threadIdx.x; p < npix; p += kBlock) {
    const bool use = !P.mask || P.mask[p];
    double a[NS], r, g, chi;
    pixel(p, use, a, r, g, chi);
    if (linear) {
      const double w = use ? __ldg(P.w_data + p) : 0.0;
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < NS; ++k) q += s_h[k] * a[k];
      const double gq = g - w * q, wr = w * r;
#pragma unroll
      for (int k = 0; k < NS; ++k)
        if (k < ns) out[(int64_t)k * npix + p] = dead ? 0.0 : fma(s_amp[k], gq, -(s_h[k] * wr));
    } else {
#pragma unroll
      for (int k = 0; k < NS; ++k)
        if (k < ns) out[(int64_t)k * npix + p] = g;
    }
  }
}

// ---- k_conv_adj_edge ----------------------------------------------------------------------------------------------
// Transpose of ONE 1-D pass of the edge-replicated separable convolution
//   out[o] = sum_d k1[d + R] in[clamp(o - d)],  d = -R..R        (GaussianConvolution, convolution.py:216-236)
// along x (axis 0) or y (axis 1):  adj[i] = sum_o out_bar[o] [clamp(o - d) == i] k1[d + R].
// Interior i collects o = i + d; the two edge pixels also collect every tap that was clamped onto them.
static __global__ void __launch_bounds__(kBlock)
k_conv_adj_edge(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int K,
                const double* __restrict__ k1, int axis, double scale) {
  const int R = (K - 1) / 2;
  const int64_t frame = (int64_t)blockIdx.y * nx * ny;
  const int n = axis == 0 ? nx : ny, stride = axis == 0 ? 1 : nx;
  for (int idx = blockIdx.x * kBlock + threadIdx.x; idx < nx * ny; idx += gridDim.x * kBlock) {
    const int y = idx / nx, x = idx - y * nx;
    const int i = axis == 0 ? x : y;
    const double* line = in + frame + (axis == 0 ? (int64_t)y * nx : (int64_t)x);
    double acc = 0.0;
    for (int o = max(0, i - R); o <= min(n - 1, i + R); ++o) {  // unclamped taps: o - d = i
      acc = fma(k1[o - i + R], line[(int64_t)o * stride], acc);
    }
    if (i == 0) {  // taps with o - d < 0, i.e. d > o
      for (int o = 0; o < min(n, R); ++o) {
        double t = 0.0;
        for (int d = o + 1; d <= R; ++d) t += k1[d + R];
        acc = fma(t, line[(int64_t)o * stride], acc);
      }
    }
    if (i == n - 1) {  // taps with o - d > n - 1, i.e. d < o - (n - 1)
      for (int o = max(0, n - R); o < n; ++o) {
        double t = 0.0;
        for (int d = -R; d < o - (n - 1); ++d) t += k1[d + R];
        acc = fma(t, line[(int64_t)o * stride], acc);
      }
    }
    out[frame + idx] = acc * scale;
  }
}

// ---- k_image_grad ----------------------------------------------------------------------------------------------------
// grid (tiles, sets).  shared memory: GLens[n_lens] | GLight[n_light] | coefficient pool [n_series][kGradPool] |
// acc[E][kGradBlock].  Per sub-pixel: every lens profile once (value + Jacobian, lens_point), every light component
// once (value + gradient, light_point), then the lens parameters' contraction with the back-propagated source-plane
// gradient (lens_vjp).  EPL, SHEAR and the light profiles use closed-form derivatives, the other lens profiles dual
// numbers (lsim_grad.cuh).
// NS: the first NS profiles of the gradient's lens list are EPL-family profiles (expand_grad_lens puts them first)
// whose point state -- the 2 x 5 Jacobian of the major-axis deflection -- stays in registers between the ray shooting
// and the lens-parameter update; every other profile is evaluated for its deflection first and contracted later
// (shear and multipoles need nothing from the first pass; dual-number profiles are evaluated twice).
// TIER (lsim_grad.cuh): 0 = EPL family + SHEAR only, 1 = + closed-form multipoles and NIE / SIE, 2 = + dual numbers.
__host__ __device__ inline size_t grad_smem_bytes(int n_lens, int n_light, int n_series) {
  return (size_t)n_lens * sizeof(ad::GLens) + (size_t)n_light * sizeof(ad::GLight) +
         (size_t)n_series * ad::kGradPool * sizeof(double) +
         (size_t)grad_entries(n_lens, n_light) * kGradBlock * sizeof(double);
}

// MP: multi-plane lensing (NS = 0, TIER = 2): the rays go through the planes in order (mp_ray_forward), the
// source-plane gradient comes back through them in reverse (mp_ray_reverse, lsim_grad.cuh); every profile is prepared
// with all of its parameters as independent variables so that a fixed centre still yields d alpha / d theta.
template <int NS, int TIER, bool MP = false>
__global__ void __launch_bounds__(kGradBlock, TIER < 2 ? LSIM_GRAD_FAST_BLOCKS : 2)
k_image_grad(PlanDev P, const double* __restrict__ params, int64_t set0, int ppt, int n_series,
             const double* __restrict__ ubar, double wscale, double* __restrict__ partial) {
  using namespace ad;
  extern __shared__ double smem[];
  __shared__ MpPlane s_pl[MP ? kMaxGradPlanes : 1];
  const int64_t bl = blockIdx.y;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  GLens* L = reinterpret_cast<GLens*>(smem);
  GLight* G = reinterpret_cast<GLight*>(L + P.n_lens);
  double* pool = reinterpret_cast<double*>(G + P.n_light);
  double* acc = pool + (size_t)n_series * kGradPool;
  const int E = grad_entries(P.n_lens, P.n_light);
  stage_math_tables();
  for (int i = threadIdx.x; i < P.n_lens + P.n_light; i += kGradBlock) {
    if (i < P.n_lens) {
      int off = 0;
      for (int j = 0; j < i; ++j) off += lens_grad_has_series(P.lens[j]) ? 1 : 0;
      L[i].coef_off = off * kGradPool;
      if (MP) {
        lens_prepare_all_vars(P.lens[i], row, L[i], pool + off * kGradPool, pool + off * kGradPool + kMaxSeries);
        s_pl[i].dT = P.lens[i].delta_T;
        s_pl[i].iTz = P.lens[i].T_z != 0 ? 1.0 / P.lens[i].T_z : 0.0;
        s_pl[i].r2p = P.lens[i].red2phys;
        int ix, iy;
        vjp_center_slots(L[i], P.lens[i], ix, iy);
        s_pl[i].ix = ix;
        s_pl[i].iy = iy;
      } else {
        lens_prepare_d(P.lens[i], row, L[i], pool + off * kGradPool, pool + off * kGradPool + kMaxSeries);
      }
    } else {
      light_prepare_g(P.light[i - P.n_lens], row, G[i - P.n_lens]);
    }
  }
  double* my = acc + threadIdx.x;
  for (int e = 0; e < E; ++e) my[e * kGradBlock] = 0.0;
  __syncthreads();

  const int npix = P.nx * P.ny;
  const int ss = P.ss;
  const int64_t sub_stride = (int64_t)P.nx * ss;
  const double wsub = wscale / ((double)ss * ss);
  const double as = P.ray.alpha_scaling;
  double* lacc = my + (size_t)kLensEntries * P.n_lens * kGradBlock;
  // flux_evaluate_indexes: pixels outside the list have zero flux whatever the parameters are
  const int q_end = min(P.eval_list ? P.n_eval : npix, (int)(blockIdx.x + 1) * ppt);
  for (int q = blockIdx.x * ppt + threadIdx.x; q < q_end; q += kGradBlock) {
    const int p = P.eval_list ? __ldg(P.eval_list + q) : q;
    const int py = p / P.nx, px = p - py * P.nx;
    for (int i = 0; i < ss; ++i) {
      const int64_t base = ((int64_t)py * ss + i) * sub_stride + (int64_t)px * ss;
      for (int j = 0; j < ss; ++j) {
        const double x = __ldg(P.x_sub + base + j), y = __ldg(P.y_sub + base + j);
        if (MP) {
          double tx[kMaxGradPlanes], ty[kMaxGradPlanes], bx, by;
          mp_ray_forward<TIER>(L, s_pl, P.n_lens, pool, P.ray.T_ij_end, P.ray.inv_T_z_source, x, y, tx, ty, bx, by);
          double vx = 0.0, vy = 0.0;
          for (int g = 0; g < P.n_light; ++g) {
            const GLight& Gg = G[g];
            if (Gg.slot < 0) continue;
            const double w = __ldg(ubar + (bl * P.n_slots + Gg.slot) * (int64_t)npix + p) * wsub;
            double I, Ix, Iy, gg[kLightG];
            light_point(Gg, Gg.lensed ? bx : x, Gg.lensed ? by : y, I, Ix, Iy, gg);
            if (Gg.lensed) {
              vx = fma(w, Ix, vx);
              vy = fma(w, Iy, vy);
            }
            double* e = lacc + (size_t)g * kGradLightStride * kGradBlock;
#pragma unroll
            for (int k = 0; k < kGradLightStride; ++k) e[k * kGradBlock] = fma(w, gg[k], e[k * kGradBlock]);
          }
          mp_ray_reverse<TIER>(L, s_pl, P.n_lens, pool, P.ray.T_ij_end, P.ray.inv_T_z_source, tx, ty, vx, vy,
                               [&](int l, const double (&gl)[kLensEntries]) {
                                 double* e = my + (size_t)l * kLensEntries * kGradBlock;
                                 for (int k = 0; k < kLensEntries; ++k) e[k * kGradBlock] += gl[k];
                               });
          continue;
        }
        double ax = 0.0, ay = 0.0;
        LensPt pt[NS > 0 ? NS : 1];
#pragma unroll
        for (int l = 0; l < NS; ++l) {
          lens_point<TIER>(L[l], pool, x, y, pt[l]);
          ax += pt[l].ax;
          ay += pt[l].ay;
        }
        for (int l = NS; l < P.n_lens; ++l) {
          double fx, fy;
          lens_value<TIER>(L[l], pool, x, y, fx, fy);
          ax += fx;
          ay += fy;
        }
        const double bx = x - ax * as, by = y - ay * as;
        double vx = 0.0, vy = 0.0;
        for (int g = 0; g < P.n_light; ++g) {
          const GLight& Gg = G[g];
          if (Gg.slot < 0) continue;
          const double w = __ldg(ubar + (bl * P.n_slots + Gg.slot) * (int64_t)npix + p) * wsub;
          double I, Ix, Iy, gg[kLightG];
          light_point(Gg, Gg.lensed ? bx : x, Gg.lensed ? by : y, I, Ix, Iy, gg);
          if (Gg.lensed) {
            vx = fma(w, Ix, vx);
            vy = fma(w, Iy, vy);
          }
          double* e = lacc + (size_t)g * kGradLightStride * kGradBlock;
#pragma unroll
          for (int k = 0; k < kGradLightStride; ++k) e[k * kGradBlock] = fma(w, gg[k], e[k * kGradBlock]);
        }
        vx *= -as;  // d beta / d p = -alpha_scaling d alpha / d p
        vy *= -as;
#pragma unroll
        for (int l = 0; l < NS; ++l) {
          double gl[kLensEntries];
          lens_vjp<TIER>(L[l], pt[l], x, y, vx, vy, gl);
          double* e = my + (size_t)l * kLensEntries * kGradBlock;
#pragma unroll
          for (int k = 0; k < kLensN; ++k) e[k * kGradBlock] += gl[k];
        }
        for (int l = NS; l < P.n_lens; ++l) {
          double gl[kLensEntries];
          lens_vjp_recompute<TIER>(L[l], pool, x, y, vx, vy, gl);
          double* e = my + (size_t)l * kLensEntries * kGradBlock;
          const int ne = L[l].type == LSIM_LENS_SHEAR ? 4 : kLensN;
          for (int k = 0; k < ne; ++k) e[k * kGradBlock] += gl[k];
          if (L[l].type == kGradLensMP) e[6 * kGradBlock] += gl[6];  // the amplitude's entry
        }
      }
    }
  }
  __syncthreads();
  // one partial per entry and CTA: warp w reduces entries w, w + 8, ... in a fixed order; the lens entries then go
  // from the profile's primitives to its parameters (lens_chain)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* dst = partial + (bl * gridDim.x + blockIdx.x) * (int64_t)E;
  double* red = acc;  // entry e lands in acc[e]: the slot of (entry 0, thread e) -- free once entry 0 ... are reduced
  double sums[52];  // E <= 200 (shared-memory limit checked at plan creation) -> at most 50 entries per warp
  int n_mine = 0;
  for (int e = wid; e < E; e += kGradBlock / 32) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kGradBlock / 32; ++j) s += acc[e * kGradBlock + lane + 32 * j];
    sums[n_mine++] = warp_sum(s);
  }
  __syncthreads();  // every column of acc has been read
  n_mine = 0;
  for (int e = wid; e < E; e += kGradBlock / 32) {
    if (lane == 0) red[e] = sums[n_mine];
    ++n_mine;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P.n_lens + P.n_light; i += kGradBlock) {
    if (i < P.n_lens) {
      double s[kLensEntries], g[kLensEntries];
      for (int k = 0; k < kLensEntries; ++k) s[k] = red[i * kLensEntries + k];
      lens_chain(L[i], s, g);
      for (int k = 0; k < kLensEntries; ++k) dst[i * kLensEntries + k] = g[k];
    } else {
      const int o = kLensEntries * P.n_lens + (i - P.n_lens) * kGradLightStride;
      for (int k = 0; k < kGradLightStride; ++k) dst[o + k] = red[o + k];
    }
  }
}

// ---- k_hessian: LensModel.hessian (LensModel/lens_model.py:384-414, single_plane.py:142-180) -------------------------
// grid = (point tiles, sets); one thread per point, the prepared dual-number profiles of the set in shared memory.
//   out [B, 4, N] = f_xx, f_xy, f_yx, f_yy, times alpha_scaling
__host__ __device__ inline size_t hess_smem_bytes(int n_lens, int n_series) {
  return (size_t)n_lens * (sizeof(ad::GLens) + 2 * sizeof(int)) + (size_t)n_series * ad::kGradPool * sizeof(double);
}

static __global__ void __launch_bounds__(kGradBlock)
k_hessian(PlanDev P, const double* __restrict__ params, const double* __restrict__ x, const double* __restrict__ y,
          int64_t N, int coords_batched, int n_series, double scale, double* __restrict__ out) {
  using namespace ad;
  extern __shared__ double smem[];
  const int64_t b = blockIdx.y;
  const double* row = params + b * P.n_param_cols;
  GLens* L = reinterpret_cast<GLens*>(smem);
  double* pool = reinterpret_cast<double*>(L + P.n_lens);
  int* slot = reinterpret_cast<int*>(pool + (size_t)n_series * kGradPool);
  stage_math_tables();
  for (int i = threadIdx.x; i < P.n_lens; i += kGradBlock) {
    int off = 0;
    for (int j = 0; j < i; ++j) off += lens_grad_has_series(P.lens[j]) ? 1 : 0;
    L[i].coef_off = off * kGradPool;
    lens_prepare_hess(P.lens[i], row, L[i], pool + off * kGradPool, pool + off * kGradPool + kMaxSeries,
                      P.series_tol);
    int ix, iy;
    if (!lens_center_slots(P.lens[i], ix, iy)) ix = iy = -1;
    slot[2 * i] = ix;
    slot[2 * i + 1] = iy;
  }
  __syncthreads();
  const double* xb = coords_batched ? x + b * N : x;
  const double* yb = coords_batched ? y + b * N : y;
  double* o = out + b * 4 * N;
  for (int64_t i = (int64_t)blockIdx.x * kGradBlock + threadIdx.x; i < N; i += (int64_t)gridDim.x * kGradBlock) {
    const double xx = xb[i], yy = yb[i];
    double h[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < P.n_lens; ++l)
      if (slot[2 * l] >= 0) lens_hessian_add(L[l], pool, slot[2 * l], slot[2 * l + 1], xx, yy, h);
    o[i] = h[0] * scale;
    o[N + i] = h[1] * scale;
    o[2 * N + i] = h[2] * scale;
    o[3 * N + i] = h[3] * scale;
  }
}

// ---- k_source_position: PositionLikelihood.source_position_likelihood ------------------------------------------------
// (Sampling/Likelihoods/position_likelihood.py:214-283 and _compute_penalty :296-352; LENSED_POSITION.source_position,
// PointSource/Types/lensed_position.py:65-90.)  One warp per parameter set; a lane per image:
//   beta_i = theta_i - alpha(theta_i),  H_i = Hessian(theta_i)      one dual-number evaluation gives both
//   beta_avg = mean_i beta_i;  delta_i = beta_avg - beta_i
//   A = 1 - H_i;  Sigma_beta = A^T (sigma^2 1) A;  penalty_i = delta^T adj(Sigma_beta) delta / (2 det Sigma_beta)
//   (det == 0: 1e15;  |delta|^2 > hard_bound_rms^2: + 1e3);   logl = - sum_i penalty_i      (sums in image order)
__host__ __device__ inline size_t srcpos_smem_bytes(int n_lens, int n_series, int n_img) {
  return hess_smem_bytes(n_lens, n_series) + (size_t)7 * n_img * sizeof(double);
}

static __global__ void __launch_bounds__(32)
k_source_position(PlanDev P, const double* __restrict__ params, const double* __restrict__ x,
                  const double* __restrict__ y, int n_img, int coords_batched, int n_series, double scale, double sigma,
                  double hard_bound_rms, double* __restrict__ logl) {
  using namespace ad;
  extern __shared__ double smem[];
  const int64_t b = blockIdx.x;
  const double* row = params + b * P.n_param_cols;
  GLens* L = reinterpret_cast<GLens*>(smem);
  double* pool = reinterpret_cast<double*>(L + P.n_lens);
  double* w = pool + (size_t)n_series * kGradPool;  // [7][n_img]: beta_x, beta_y, f_xx, f_xy, f_yx, f_yy, penalty
  int* slot = reinterpret_cast<int*>(w + (size_t)7 * n_img);
  stage_math_tables();
  for (int i = threadIdx.x; i < P.n_lens; i += 32) {
    int off = 0;
    for (int j = 0; j < i; ++j) off += lens_grad_has_series(P.lens[j]) ? 1 : 0;
    L[i].coef_off = off * kGradPool;
    lens_prepare_hess(P.lens[i], row, L[i], pool + off * kGradPool, pool + off * kGradPool + kMaxSeries,
                      P.series_tol);
    int ix, iy;
    if (!lens_center_slots(P.lens[i], ix, iy)) ix = iy = -1;
    slot[2 * i] = ix;
    slot[2 * i + 1] = iy;
  }
  __syncthreads();
  const double* xb = coords_batched ? x + b * n_img : x;
  const double* yb = coords_batched ? y + b * n_img : y;
  for (int i = threadIdx.x; i < n_img; i += 32) {
    const double xx = xb[i], yy = yb[i];
    double ax = 0.0, ay = 0.0, h[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < P.n_lens; ++l)
      lens_value_hessian_add(L[l], pool, slot[2 * l], slot[2 * l + 1], xx, yy, ax, ay, h);
    w[i] = xx - ax * scale;
    w[n_img + i] = yy - ay * scale;
    for (int k = 0; k < 4; ++k) w[(2 + k) * n_img + i] = h[k] * scale;
  }
  __syncthreads();
  double sx = 0.0, sy = 0.0;
  for (int i = 0; i < n_img; ++i) {
    sx += w[i];
    sy += w[n_img + i];
  }
  const double mx = sx / n_img, my = sy / n_img;
  for (int i = threadIdx.x; i < n_img; i += 32) {
    const double h[4] = {w[2 * n_img + i], w[3 * n_img + i], w[4 * n_img + i], w[5 * n_img + i]};
    const double pen = source_position_penalty(h, mx - w[i], my - w[n_img + i], sigma, hard_bound_rms);
    w[6 * n_img + i] = pen;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_img; ++i) s += w[6 * n_img + i];
    logl[b] = -s;
  }
}

// ---- k_point_sources_grad: positions and amplitudes of the point-source images -------------------------------------------
// grid (n_points, sets).  The stamp of image i lands in slot image `slot` AFTER the PSF convolution (k_point_sources), so
// its parameters see Abar = dlogL/dC directly:  with S(xs, ys) = sum over the footprint of Abar[pixel] x (sum over the
// ss x ss sub-pixels of the bilinear interpolant of the PSF kernel at (r + K/2 - ys, c + K/2 - xs)),
//   dlogL/d(amp) = scale S,   dlogL/d(xs) = amp scale dS/dxs,   (xs, ys) = ss (M (ra, dec) + x0) + (ss - 1)/2
// and dS/dxs is the sum of the interpolant's own slope -- piecewise constant in x, the derivative jax takes of
// map_coordinates(order = 1) (ImSim/Numerics/point_source_rendering.py:33-77).  Entries are ADDED to grad[B, P] after
// k_grad_reduce has written it; every (set, column) is touched by one CTA, so the result is reproducible run to run.
__device__ __forceinline__ void ps_bilinear_d(const double* __restrict__ k, int K, double ky, double kx, double& v,
                                              double& dky, double& dkx) {
  const double fy = floor(ky), fx = floor(kx);
  const double wy1 = ky - fy, wx1 = kx - fx, wy0 = 1.0 - wy1, wx0 = 1.0 - wx1;
  const int iy = (int)fy, ix = (int)fx;
  auto at = [&](int y, int x) { return (y >= 0 && y < K && x >= 0 && x < K) ? k[y * K + x] : 0.0; };
  const double k00 = at(iy, ix), k01 = at(iy, ix + 1), k10 = at(iy + 1, ix), k11 = at(iy + 1, ix + 1);
  v = (wy0 * wx0) * k00 + (wy0 * wx1) * k01 + (wy1 * wx0) * k10 + (wy1 * wx1) * k11;
  dkx = wy0 * (k01 - k00) + wy1 * (k11 - k10);
  dky = wx0 * (k10 - k00) + wx1 * (k11 - k01);
}

static __global__ void __launch_bounds__(kBlock)
k_point_sources_grad(PlanDev P, const double* __restrict__ params, int64_t set0, const double* __restrict__ abar,
                     double* __restrict__ grad) {
  __shared__ double scratch[3 * (kBlock / 32)];
  const lsim_point_desc& d = P.points[blockIdx.x];
  const int64_t bl = blockIdx.y;
  if (d.slot < 0 || (d.col_ra < 0 && d.col_dec < 0 && !(d.use_amp && d.col_amp >= 0))) return;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  const double* A = abar + (bl * P.n_slots + d.slot) * ((int64_t)P.nx * P.ny);
  const int K = P.ps_K, ss = P.ps_ss, half = K / 2;
  const double ra = d.col_ra >= 0 ? row[d.col_ra] : d.cst_ra;
  const double dec = d.col_dec >= 0 ? row[d.col_dec] : d.cst_dec;
  const double amp = d.use_amp ? (d.col_amp >= 0 ? row[d.col_amp] : d.cst_amp) : 1.0;
  const double xp = P.ps_M[0] * ra + P.ps_M[1] * dec + P.ps_x0;
  const double yp = P.ps_M[2] * ra + P.ps_M[3] * dec + P.ps_y0;
  const double xs = xp * ss + (ss - 1) / 2.0, ys = yp * ss + (ss - 1) / 2.0;
  double sums[3] = {0.0, 0.0, 0.0};  // S, dS/dxs, dS/dys
  if (fabs(xs) < 1e9 && fabs(ys) < 1e9) {
    const int c_lo = (int)floor(xs - half - 1.0), c_hi = (int)ceil(xs - half + K);
    const int r_lo = (int)floor(ys - half - 1.0), r_hi = (int)ceil(ys - half + K);
    const int px_lo = max(0, (int)floor((double)c_lo / ss)), px_hi = min(P.nx - 1, (int)floor((double)c_hi / ss));
    const int py_lo = max(0, (int)floor((double)r_lo / ss)), py_hi = min(P.ny - 1, (int)floor((double)r_hi / ss));
    const int w = px_hi - px_lo + 1, h = py_hi - py_lo + 1;
    if (w > 0 && h > 0) {
      for (int t = threadIdx.x; t < w * h; t += kBlock) {
        const int py = py_lo + t / w, px = px_lo + t % w;
        double v = 0.0, gy = 0.0, gx = 0.0;
        for (int sy = 0; sy < ss; ++sy)
          for (int sx = 0; sx < ss; ++sx) {
            double a, dky, dkx;
            ps_bilinear_d(P.ps_kernel, K, (py * ss + sy) + half - ys, (px * ss + sx) + half - xs, a, dky, dkx);
            v += a;
            gy -= dky;  // ky = r + K/2 - ys
            gx -= dkx;
          }
        const double ab = A[(int64_t)py * P.nx + px];
        sums[0] = fma(ab, v, sums[0]);
        sums[1] = fma(ab, gx, sums[1]);
        sums[2] = fma(ab, gy, sums[2]);
      }
    }
  }
  block_sum(sums, scratch);
  if (threadIdx.x == 0) {
    double* g = grad + (set0 + bl) * P.n_param_cols;
    const double f = amp * P.ps_scale * ss;
    if (d.col_ra >= 0) atomicAdd(g + d.col_ra, f * (P.ps_M[0] * sums[1] + P.ps_M[2] * sums[2]));
    if (d.col_dec >= 0) atomicAdd(g + d.col_dec, f * (P.ps_M[1] * sums[1] + P.ps_M[3] * sums[2]));
    if (d.use_amp && d.col_amp >= 0) atomicAdd(g + d.col_amp, P.ps_scale * sums[0]);
  }
}

// ---- k_grad_reduce ------------------------------------------------------------------------------------------------
// One CTA per set: entry sums over the tiles (fixed order), then the entries of each parameter column.
static __global__ void __launch_bounds__(128)
k_grad_reduce(PlanDev P, const double* __restrict__ partial, int tiles, int64_t set0, double* __restrict__ grad) {
  extern __shared__ double s_e[];
  const int64_t bl = blockIdx.x;
  const int E = grad_entries(P.n_lens, P.n_light);
  const double* src = partial + bl * tiles * (int64_t)E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    double s = 0.0;
    for (int t = 0; t < tiles; ++t) s += src[(int64_t)t * E + e];
    s_e[e] = s;
  }
  __syncthreads();
  double* g = grad + (set0 + bl) * P.n_param_cols;
  for (int c = threadIdx.x; c < P.n_param_cols; c += blockDim.x) {
    double s = 0.0;
    for (int l = 0; l < P.n_lens; ++l) {
      const int np = grad_lens_params(P.lens[l].type);
      for (int k = 0; k < np; ++k)
        if (P.lens[l].col[k] == c) s += s_e[l * ad::kLensEntries + k];
    }
    for (int gi = 0; gi < P.n_light; ++gi) {
      const lsim_light_desc& d = P.light[gi];
      if (d.slot < 0) continue;
      const int np = grad_light_params(d.type);
      for (int k = 0; k < np; ++k)
        if (d.col[k] == c && !(k == 0 && !d.use_amp)) s += s_e[ad::kLensEntries * P.n_lens + gi * kGradLightStride + k];
    }
    g[c] = s;
  }
}

}  // namespace lsim

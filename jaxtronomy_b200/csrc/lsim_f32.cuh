// lsim_f32.cuh -- the opt-in single-precision variant of the fused image kernel (plan precision = LSIM_PRECISION_F32).
//
// fp64 is the default and the parity-tested path (rtol 1e-10).  This variant evaluates the per-sub-pixel arithmetic
// (deflection, source position, surface brightness, ss x ss mean) in fp32 -- twice the lanes of the fp64 pipe and
// log2 / exp2 / reciprocal on the special-function unit -- and hands the re-binned image to the same fp64
// convolution and likelihood kernels.  Stated tolerance (tests/test_gpu_parity.py::test_fp32_variant): images
// |a - b| <= 2e-4 |b| + 2e-5 max|b|, log-likelihood within 2e-4 relative.
//
// Scope: single plane; lens EPL (and SIE evaluated as EPL), NIE / SIE, SIS, SHEAR, CONVERGENCE; light SERSIC,
// SERSIC_ELLIPSE, GAUSSIAN; no flux_evaluate_indexes, no supersampled convolution.  Anything else is refused at plan
// creation (LSIM_E_UNSUPPORTED) -- there is no silent fallback to fp64.
//
// The per-set constants are prepared in fp64 by the same lens_prepare / light_prepare as the fp64 kernels
// (parameter conversion, series coefficients with the reference's nmax schedule, epl.py:308-319) and then rounded
// once to fp32 in shared memory.
#pragma once

#include "lsim_kernels.cuh"

namespace lsim {

struct LensF {
  int32_t type, n_coef, coef_off, pad;
  float c[12];
};
struct LightF {
  int32_t type, lensed, slot, pad;
  float c[12];
};

__device__ __forceinline__ float ex2_f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_f(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_f(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float nan_to_num_f(float v, float p, float n) {
  if (v != v) return 0.f;
  if (v > FLT_MAX) return p;
  if (v < -FLT_MAX) return n;
  return v;
}

// same algebra as epl_eval (lsim_profiles.cuh), in fp32
template <int J>
__device__ __forceinline__ void epl_eval_f(const LensF& L, const float* coef, const float (&xs)[J],
                                           const float (&ys)[J], float (&ax)[J], float (&ay)[J]) {
  const float cph = L.c[2], sph = L.c[3], q = L.c[4], f = L.c[5];
  float wi[J], gr[J], gi[J], pp[J], qq[J], b1[J], b2[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const float xr = fmaf(ys[j], sph, xs[j] * cph);
    const float c = fmaf(ys[j], cph, -(xs[j] * sph));
    const float a = q * xr;
    const float aa = a * a;
    const float R2 = fmaf(c, c, aa);
    const float z2r = fmaf(-c, c, aa), z2i = 2.f * a * c;
    const float R2x2 = R2 + R2;
    const float Dr = fmaf(f, z2r, R2x2), Di = f * z2i;
    const float inv = rcp_f(fmaf(Dr, Dr, Di * Di));
    const float finv = f * inv;
    const float wr = fmaf(z2i, Di, z2r * Dr) * finv;
    wi[j] = fmaf(z2i, Dr, -(z2r * Di)) * finv;
    const float R2c = fmaxf(R2, 1e-18f);
    const float Pw = L.c[8] * ex2_f(fmaf(L.c[6], lg2_f(R2c), L.c[7]));
    const float s2 = R2x2 * inv * Pw;
    const float pr = s2 * Dr, pi = -(s2 * Di);
    gr[j] = fmaf(a, pr, -(c * pi));
    gi[j] = fmaf(a, pi, c * pr);
    pp[j] = wr + wr;
    qq[j] = -fmaf(wr, wr, wi[j] * wi[j]);
    b1[j] = 0.f;
    b2[j] = 0.f;
  }
  for (int k = L.n_coef - 1; k >= 1; --k) {
    const float ck = coef[k];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float t = fmaf(pp[j], b1[j], fmaf(qq[j], b2[j], ck));
      b2[j] = b1[j];
      b1[j] = t;
    }
  }
  const float c0 = coef[0];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const float Sr = fmaf(b1[j], 0.5f * pp[j], fmaf(qq[j], b2[j], c0)), Si = b1[j] * wi[j];
    float fr = fmaf(gr[j], Sr, -(gi[j] * Si)), fi = fmaf(gr[j], Si, gi[j] * Sr);
    fr = nan_to_num_f(fr, 1e10f, -1e10f);
    fi = nan_to_num_f(fi, 1e10f, -1e10f);
    ax[j] = fmaf(-fi, sph, fr * cph);
    ay[j] = fmaf(fi, cph, fr * sph);
  }
}

__device__ __forceinline__ void lens_eval1_f(const LensF& L, float x, float y, float& fx, float& fy) {
  const float xs = x - L.c[0], ys = y - L.c[1];
  switch (L.type) {
    case LSIM_LENS_SHEAR:
      fx = fmaf(L.c[3], ys, L.c[2] * xs);
      fy = fmaf(-L.c[2], ys, L.c[3] * xs);
      return;
    case LSIM_LENS_SIS: {
      const float R = fmaxf(sqrtf(fmaf(ys, ys, xs * xs)), 1e-6f);
      const float a = L.c[2] / R;
      fx = a * xs;
      fy = a * ys;
      return;
    }
    case LSIM_LENS_CONVERGENCE:
      fx = L.c[2] * xs;
      fy = L.c[2] * ys;
      return;
    case LSIM_LENS_NIE: {
      const float cph = L.c[2], sph = L.c[3];
      const float xr = fmaf(ys, sph, xs * cph), yr = fmaf(ys, cph, -(xs * sph));
      const float s = L.c[5];
      const float psi = sqrtf(fmaf(L.c[4], fmaf(xr, xr, s * s), yr * yr));
      const float gx = L.c[7] * atanf(L.c[6] * xr / (psi + s));
      const float gy = L.c[7] * atanhf(L.c[6] * yr / (psi + L.c[8]));
      fx = fmaf(-gy, sph, gx * cph);
      fy = fmaf(gy, cph, gx * sph);
      return;
    }
    default:
      fx = 0.f;
      fy = 0.f;
  }
}

template <int J>
__device__ __forceinline__ void light_eval_f(const LightF& G, const float (&x)[J], const float (&y)[J],
                                             float (&v)[J]) {
  if (G.type == LSIM_LIGHT_SERSIC_ELLIPSE) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float xs = x[j] - G.c[0], ys = y[j] - G.c[1];
      const float xt = fmaf(G.c[3], ys, G.c[2] * xs);
      const float yt = fmaf(G.c[4], ys, G.c[3] * xs);
      const float R2c = fmaxf(fmaf(yt, yt, xt * xt), 1e-8f);
      const float k = ex2_f(fmaf(G.c[5], lg2_f(R2c), G.c[6]));
      float e = ex2_f(fmaf(G.c[7], k, G.c[8]));
      e = (R2c <= G.c[9]) ? e : 0.f;
      e = (e == e) ? e : 0.f;
      v[j] = G.c[10] * e;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {  // GAUSSIAN
    const float xs = x[j] - G.c[0], ys = y[j] - G.c[1];
    v[j] = G.c[3] * ex2_f(G.c[2] * fmaf(ys, ys, xs * xs));
  }
}

// true when every profile of the plan is inside the fp32 variant's scope
inline bool f32_supported(const lsim_plan_desc* d) {
  if (d->multi_plane || d->eval_mask || d->supersampled_conv) return false;
  for (int i = 0; i < d->n_lens; ++i) {
    switch (d->lens[i].type) {
      case LSIM_LENS_NULL: case LSIM_LENS_EPL: case LSIM_LENS_SIE: case LSIM_LENS_NIE: case LSIM_LENS_SIS:
      case LSIM_LENS_SHEAR: case LSIM_LENS_CONVERGENCE:
        break;
      default:
        return false;
    }
  }
  for (int i = 0; i < d->n_light; ++i)
    if (d->light[i].type < 0 || d->light[i].type >= LSIM_LIGHT_TYPE_COUNT) return false;
  return true;
}

inline size_t f32_extra_smem(int n_lens, int n_light, int n_series) {
  return (size_t)n_lens * sizeof(LensF) + (size_t)n_light * sizeof(LightF) + (size_t)n_series * kMaxSeries * sizeof(float);
}

// grid = (tiles, sets) like k_image; SS > 0: the SS sub-pixels of a row in lock step, SS == 0: runtime factor.
template <int NSLOT, int SS>
__global__ void __launch_bounds__(kImageBlock, 4)
k_image_f32(PlanDev P, const double* __restrict__ params, int64_t set0, int ppt, double out_scale, int n_series,
            double* __restrict__ out, const unsigned char* __restrict__ prepared) {
  constexpr int J = SS > 0 ? SS : 1;
  extern __shared__ double smem[];
  LensC* L = reinterpret_cast<LensC*>(smem);
  LightC* G = reinterpret_cast<LightC*>(L + P.n_lens);
  double* pool = reinterpret_cast<double*>(G + P.n_light);
  LensF* Lf = reinterpret_cast<LensF*>(pool + (size_t)n_series * kMaxSeries);
  LightF* Gf = reinterpret_cast<LightF*>(Lf + P.n_lens);
  float* poolf = reinterpret_cast<float*>(Gf + P.n_light);
  const int64_t bl = blockIdx.y;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  load_or_stage(P, row, prepared, bl, n_series, L, G, pool);
  __syncthreads();
  for (int i = threadIdx.x; i < P.n_lens; i += kImageBlock) {
    Lf[i].type = L[i].type;
    Lf[i].n_coef = L[i].n_coef;
    Lf[i].coef_off = L[i].coef_off;
    for (int k = 0; k < 12; ++k) Lf[i].c[k] = (float)L[i].c[k];
  }
  for (int i = threadIdx.x; i < P.n_light; i += kImageBlock) {
    Gf[i].type = G[i].type;
    Gf[i].lensed = G[i].lensed;
    Gf[i].slot = G[i].slot;
    for (int k = 0; k < 12; ++k) {
      // light_prepare stores the Sersic exponent constants c[5..8] pre-scaled by kExp2Scale (exp2_*_s of the fp64 kernels)
      const bool scaled = G[i].type == LSIM_LIGHT_SERSIC_ELLIPSE && k >= 5 && k <= 8;
      Gf[i].c[k] = (float)(scaled ? G[i].c[k] * (1.0 / kExp2Scale) : G[i].c[k]);
    }
  }
  for (int i = threadIdx.x; i < n_series * kMaxSeries; i += kImageBlock) poolf[i] = (float)pool[i];
  __syncthreads();

  const int npix = P.nx * P.ny;
  const int ss = SS > 0 ? SS : P.ss;
  const int64_t sub_stride = (int64_t)P.nx * ss;
  const float inv_ss = 1.f / ss, scaling = (float)P.ray.alpha_scaling;
  const int p_end = min(npix, (int)(blockIdx.x + 1) * ppt);
  for (int p = blockIdx.x * ppt + threadIdx.x; p < p_end; p += kImageBlock) {
    const int py = p / P.nx, px = p - py * P.nx;
    float acc[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) acc[k] = 0.f;
    for (int i = 0; i < ss; ++i) {
      const int64_t base = ((int64_t)py * ss + i) * sub_stride + (int64_t)px * ss;
      for (int j0 = 0; j0 < ss; j0 += J) {
        float x[J], y[J], ax[J], ay[J], fx[J], fy[J], v[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          x[j] = (float)__ldg(P.x_sub + base + j0 + j);
          y[j] = (float)__ldg(P.y_sub + base + j0 + j);
          ax[j] = 0.f;
          ay[j] = 0.f;
        }
        for (int l = 0; l < P.n_lens; ++l) {
          if (Lf[l].type == LSIM_LENS_EPL) {
            float xs[J], ys[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
              xs[j] = x[j] - Lf[l].c[0];
              ys[j] = y[j] - Lf[l].c[1];
            }
            epl_eval_f<J>(Lf[l], poolf + Lf[l].coef_off, xs, ys, fx, fy);
          } else {
#pragma unroll
            for (int j = 0; j < J; ++j) lens_eval1_f(Lf[l], x[j], y[j], fx[j], fy[j]);
          }
#pragma unroll
          for (int j = 0; j < J; ++j) {
            ax[j] += fx[j];
            ay[j] += fy[j];
          }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {  // beta = theta - alpha_scaling * alpha
          ax[j] = fmaf(-scaling, ax[j], x[j]);
          ay[j] = fmaf(-scaling, ay[j], y[j]);
        }
        for (int g = 0; g < P.n_light; ++g) {
          const int slot = Gf[g].slot;
          if (slot < 0) continue;
          if (Gf[g].lensed)
            light_eval_f<J>(Gf[g], ax, ay, v);
          else
            light_eval_f<J>(Gf[g], x, y, v);
          float vs = v[0];
#pragma unroll
          for (int j = 1; j < J; ++j) vs += v[j];
#pragma unroll
          for (int k = 0; k < NSLOT; ++k) acc[k] += (slot == k) ? vs : 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      if (k < P.n_slots) out[(bl * P.n_slots + k) * npix + p] = (double)(acc[k] * inv_ss * inv_ss) * out_scale;
    }
  }
}

}  // namespace lsim

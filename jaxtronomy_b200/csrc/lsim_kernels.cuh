// lsim_kernels.cuh -- the CUDA kernels of the forward-simulation path (sm_100a).
//
//   k_prepare              per-set constants (parameter conversion, series coefficients) once per parameter set
//   k_ray_shoot            LensModel.ray_shooting / SinglePlane.alpha / LensModelBulk / MultiPlane on N points x B sets
//   k_surface_brightness   LightModel.surface_brightness / functions_split on N points x B sets
//   k_image                fused ray-shoot + light + ss x ss re-bin: one thread per *pixel*, looping its
//                          ss^2 sub-pixels in registers, so the supersampled flux never exists in memory
//                          (k_image_f32 in lsim_f32.cuh: the opt-in single-precision variant)
//   k_conv_*               PSF convolution of the re-binned basis images (direct, shared-memory tiled);
//   k_fft_pad/mul/crop     ... and the helpers around cuFFT for kernels of side >= 27
//   k_point_sources        PSF stamps of the point-source images; k_ps_error_map: the PSF error map
//   k_likelihood           WLS amplitudes (normal equations reduced per CTA, solved in registers) + chi^2
//   k_pso_best/k_pso_move  one particle-swarm iteration on the device
//   k_fp64_peak            DFMA-chain microbenchmark for the roofline denominator
//
// Layout: parameter sets are the slow (grid.y / blockIdx) axis, pixels the fast axis; all global stores
// are unit-stride over pixels.  Per-set constants live in shared memory (see lsim_profiles.cuh).
#pragma once

#include <cuda_runtime.h>

#include "lsim_profiles.cuh"

namespace lsim {

constexpr int kBlock = 256;
constexpr int kLensChunk = 96;    // lens profiles resident in shared memory at a time (k_ray_shoot streams chunks)
constexpr int kMaxImageLens = 32; // k_image keeps the whole lens list resident
constexpr int kMaxLight = 16;
constexpr int kMaxSeriesProfiles = 4;  // EPL-family profiles per chunk (each owns kMaxSeries coefficients)

// Device view of a plan; passed by value (fits the kernel parameter bank).
struct PlanDev {
  int32_t n_param_cols;
  int32_t n_lens;
  const lsim_lens_desc* lens;
  RayCfg ray;
  int32_t n_light;
  const lsim_light_desc* light;
  int32_t n_slots;
  double series_tol;
  // image
  int32_t nx, ny, ss;
  const double* x_sub;
  const double* y_sub;
  // x_sub[r][c] = gx0 + r gxi + c gxj (and likewise y) to a few ulp, r / c = row / column of the supersampled grid:
  // set at plan creation when the coordinate arrays are affine in the grid indices (every RegularGrid is)
  int32_t grid_affine;
  double gx0, gxi, gxj, gy0, gyi, gyj;
  const int32_t* eval_list;  // [n_eval] indices of the pixels that are evaluated, or null = all of them
  int32_t n_eval;            // number of evaluated pixels (= nx*ny without a mask)
  int32_t conv_mode, K;
  const double* kernel;     // [K*K]
  const double* kernel_1d;  // [K] or null
  int32_t ss_conv, K_hi;    // supersampled convolution: kernel_hi on the [ny ss, nx ss] grid (+ kernel on pixels)
  const double* kernel_hi;
  const double* kernel_hi_1d;
  double image_scale;
  const double* data;
  const double* w_data;     // 1 / cd_data
  const double* noise_map2;
  const double* inv_exposure;
  double bkg2;
  const uint8_t* mask;
  double logl_offset;
  // point sources
  int32_t n_points, ps_K, ps_ss;
  const lsim_point_desc* points;
  const double* ps_kernel;
  double ps_x0, ps_y0, ps_M[4], ps_scale;
  const double* ps_var;  // PSF variance map [ps_var_K^2] or null
  int32_t ps_var_K;
  double ps_mean_kernel;
};

// ---- shared-memory staging of per-set constants -----------------------------------------------------------
// Prepares lens profiles [c0, c0+n_c) of the plan for parameter row `row` into L[0..n_c) (+ coefficient pool).
__device__ __forceinline__ void stage_lens(const PlanDev& P, const double* row, int c0, int n_c, LensC* L,
                                           double* pool) {
  for (int i = threadIdx.x; i < n_c; i += blockDim.x) {
    const lsim_lens_desc& d = P.lens[c0 + i];
    int off = 0;
    if (lens_type_has_series(d)) {
      for (int j = 0; j < i; ++j) off += lens_type_has_series(P.lens[c0 + j]) ? 1 : 0;
    }
    L[i].coef_off = off * kMaxSeries;
    lens_prepare(d, row, L[i], pool + off * kMaxSeries, P.series_tol);
  }
}

__device__ __forceinline__ void stage_light(const PlanDev& P, const double* row, LightC* G) {
  for (int i = threadIdx.x; i < P.n_light; i += blockDim.x) light_prepare(P.light[i], row, G[i]);
}

// ---- k_prepare: per-set constants computed ONCE per parameter set into global memory --------------------------------
// Block of one set = the shared-memory image k_image works from: [LensC x n_lens][LightC x n_light][coefficient pool].
// One thread per (set, profile).  The image kernel's CTAs (several per set) then fetch the block with a coalesced
// copy instead of each recomputing trig, series coefficients and normalisations behind a barrier.
__host__ __device__ inline size_t prepared_bytes(int n_lens, int n_light, int n_series) {
  return (size_t)n_lens * sizeof(LensC) + (size_t)n_light * sizeof(LightC) +
         (size_t)n_series * kMaxSeries * sizeof(double);
}

static __global__ void __launch_bounds__(128)
k_prepare(PlanDev P, const double* __restrict__ params, int64_t set0, int64_t n_sets, int n_series,
          unsigned char* __restrict__ prepared) {
  const int n_items = P.n_lens + P.n_light;
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t >= n_sets * n_items) return;
  const int64_t bl = t / n_items;
  const int i = (int)(t - bl * n_items);
  const double* row = params + (set0 + bl) * P.n_param_cols;
  unsigned char* base = prepared + bl * prepared_bytes(P.n_lens, P.n_light, n_series);
  LensC* L = reinterpret_cast<LensC*>(base);
  LightC* G = reinterpret_cast<LightC*>(L + P.n_lens);
  double* pool = reinterpret_cast<double*>(G + P.n_light);
  if (i < P.n_lens) {
    const lsim_lens_desc& d = P.lens[i];
    int off = 0;
    if (lens_type_has_series(d))
      for (int j = 0; j < i; ++j) off += lens_type_has_series(P.lens[j]) ? 1 : 0;
    LensC tmp = {};  // every field defined: the block is copied whole by the image kernels (initcheck-clean)
    tmp.coef_off = off * kMaxSeries;
    lens_prepare(d, row, tmp, pool + off * kMaxSeries, P.series_tol);
    L[i] = tmp;
  } else {
    LightC tmp = {};
    light_prepare(P.light[i - P.n_lens], row, tmp);
    G[i - P.n_lens] = tmp;
  }
}

// shared-memory image of one set: copy of the prepared block (if given) or computed in place
__device__ __forceinline__ void load_or_stage(const PlanDev& P, const double* row, const unsigned char* prepared,
                                              int64_t bl, int n_series, LensC* L, LightC* G, double* pool) {
  if (prepared) {
    const size_t nb = prepared_bytes(P.n_lens, P.n_light, n_series);
    const uint64_t* src = reinterpret_cast<const uint64_t*>(prepared + bl * nb);
    uint64_t* dst = reinterpret_cast<uint64_t*>(L);
    const int head = (int)(((size_t)P.n_lens * sizeof(LensC) + (size_t)P.n_light * sizeof(LightC)) / 8);
    for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    // coefficient pools: only the first 32 entries of every profile unless a longer series is in use
    for (int s = 0; s < n_series; ++s) {
      int need = 0;
      for (int l = 0; l < P.n_lens; ++l) {
        const LensC* Lg = reinterpret_cast<const LensC*>(src) + l;
        if (Lg->coef_off == s * kMaxSeries && Lg->n_coef > need) need = Lg->n_coef;
      }
      for (int i = threadIdx.x; i < need; i += blockDim.x) pool[s * kMaxSeries + i] = reinterpret_cast<const double*>(src)[head + s * kMaxSeries + i];
    }
  } else {
    stage_lens(P, row, 0, P.n_lens, L, pool);
    stage_light(P, row, G);
  }
}

// ---- k_ray_shoot ---------------------------------------------------------------------------------------------
// rays per thread in lock step: 3 measured +3.5 % (C2) / +2.7 % (C4) over 2; 4 spills in the 1,000-halo loop
#ifndef LSIM_RAY_ITEMS
#define LSIM_RAY_ITEMS 3
#endif
constexpr int kRayItems = LSIM_RAY_ITEMS;

// rpc: rays per CTA (a multiple of kBlock * kRayItems).  A model whose deflector list fits one shared-memory chunk
// is staged ONCE per CTA and the CTA then walks its rays, so the per-set preparation (trig, series coefficients) is
// amortised over thousands of rays; longer lists (LensModelBulk, 1000 halos) stream chunk by chunk past one ray batch
// that stays in registers (the host passes rpc = kBlock * kRayItems for those).
// LT: compile-time list of the prepared profile types (single plane; lsim_api.cu registers EPL + SHEAR), or RuntimeTypes
template <bool FULL, class LT = RuntimeTypes>
__global__ void __launch_bounds__(kBlock)
k_ray_shoot(PlanDev P, const double* __restrict__ params, const double* __restrict__ x, const double* __restrict__ y,
            int64_t N, int coords_batched, double* __restrict__ ox, double* __restrict__ oy, int what, int64_t rpc) {
  extern __shared__ double smem[];
  LensC* L = reinterpret_cast<LensC*>(smem);
  double* pool = reinterpret_cast<double*>(L + min(P.n_lens, kLensChunk));  // the host sizes the launch the same way
  const int64_t b = blockIdx.y;
  const double* row = params + b * P.n_param_cols;
  const double* xb = coords_batched ? x + b * N : x;
  const double* yb = coords_batched ? y + b * N : y;
  const bool one_chunk = P.n_lens <= kLensChunk;

  stage_math_tables();
  if (one_chunk) stage_lens(P, row, 0, P.n_lens, L, pool);
  __syncthreads();
  const int64_t r_end = min(N, ((int64_t)blockIdx.x + 1) * rpc);
  for (int64_t r0 = (int64_t)blockIdx.x * rpc; r0 < r_end; r0 += kBlock * kRayItems) {
    Rays<kRayItems> r;
    int64_t idx[kRayItems];
#pragma unroll
    for (int k = 0; k < kRayItems; ++k) {
      idx[k] = r0 + k * kBlock + threadIdx.x;
      const double x0 = idx[k] < N ? xb[idx[k]] : 0.0;
      const double y0 = idx[k] < N ? yb[idx[k]] : 0.0;
      ray_init(P.ray, k, x0, y0, r);
    }
    for (int c0 = 0; c0 < P.n_lens; c0 += kLensChunk) {
      const int n_c = min(kLensChunk, P.n_lens - c0);
      if (!one_chunk) {
        __syncthreads();
        stage_lens(P, row, c0, n_c, L, pool);
        __syncthreads();
      }
      ray_apply<FULL, kRayItems, LT>(P.ray, L, pool, n_c, r);
    }
#pragma unroll
    for (int k = 0; k < kRayItems; ++k) {
      if (idx[k] < N) {
        double bx, by;
        if (what == 1 && !P.ray.multi_plane) {
          bx = r.ax[k] * P.ray.alpha_scaling;
          by = r.ay[k] * P.ray.alpha_scaling;
        } else {
          ray_finish(P.ray, r, k, bx, by);
        }
        ox[b * N + idx[k]] = bx;
        oy[b * N + idx[k]] = by;
      }
    }
  }
}

// ---- k_surface_brightness ----------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kBlock)
k_surface_brightness(PlanDev P, const double* __restrict__ params, const double* __restrict__ x,
                     const double* __restrict__ y, int64_t N, int coords_batched, double* __restrict__ out) {
  __shared__ LightC G[kMaxLight];
  const int64_t b = blockIdx.y;
  const double* row = params + b * P.n_param_cols;
  stage_math_tables();
  stage_light(P, row, G);
  __syncthreads();
  const double* xb = coords_batched ? x + b * N : x;
  const double* yb = coords_batched ? y + b * N : y;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < N; i += (int64_t)gridDim.x * kBlock) {
    double xx = xb[i], yy = yb[i];
    for (int s = 0; s < P.n_slots; ++s) {
      double acc = 0;
      for (int g = 0; g < P.n_light; ++g) {
        if (G[g].slot == s) acc += light_eval(G[g], xx, yy);
      }
      out[(b * P.n_slots + s) * N + i] = acc;
    }
  }
}

// A second read of a value loaded before: `volatile` keeps the compiler from merging it with the first load.
__device__ __forceinline__ double ldg_again(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// ---- k_image: fused ray-shoot + light + re-bin ---------------------------------------------------------------
// grid = (tiles, sets); each CTA owns pixels [tile*ppt, (tile+1)*ppt) of one parameter set.
// out[(set_local * n_slots + slot) * npix + pixel] = mean over the ss x ss sub-pixels  (* out_scale)
// (RegularGrid.flux_array2image_low_high + image_util.re_size: grid.py:101-117, image_util.py:78-98).
// SS > 0: compile-time supersampling factor, the SS sub-pixels of a row are evaluated in lock step (J = SS);
// SS == 0: runtime supersampling factor, one sub-pixel at a time.
// LT / GT: compile-time lists of the (prepared) lens / light profile types, or RuntimeTypes.  A plan whose model
// matches one of the registered signatures (lsim_api.cu) runs a kernel in which both lists are constants: the
// profile loops unroll, the type dispatch disappears and the per-set constants are read at fixed shared-memory
// offsets.  Any other model runs the RuntimeTypes instantiation (same code, interpreted).
// CTA shape of the image kernel (tunable at build time; see profiles/r1_bench_log.md for the sweep)
#ifndef LSIM_IMAGE_BLOCK
#define LSIM_IMAGE_BLOCK 256
#endif
#ifndef LSIM_IMAGE_MIN_BLOCKS
#define LSIM_IMAGE_MIN_BLOCKS 3  // 24 warps/SM at 80 registers (a few spilled values) beats 16 warps at 114
#endif
constexpr int kImageBlock = LSIM_IMAGE_BLOCK;
// SUB: also write the sub-pixel fluxes (supersampled convolution); a separate instantiation so that the common
// kernels carry no trace of it.
// MASKED: visit only the pixels of P.eval_list (flux_evaluate_indexes); instantiated for the interpreted kernels.
template <int NSLOT, bool FULL, int SS, class LT = RuntimeTypes, class GT = RuntimeTypes, bool SUB = false,
          bool MASKED = false>
__global__ void __launch_bounds__(kImageBlock, LSIM_IMAGE_MIN_BLOCKS)
k_image(PlanDev P, const double* __restrict__ params, int64_t set0, int ppt, double out_scale,
        double* __restrict__ out, double* __restrict__ sub_out, const unsigned char* __restrict__ prepared, int n_series) {
  constexpr int J = SS > 0 ? SS : 1;
  extern __shared__ __align__(16) double smem[];
  const int64_t bl = blockIdx.y;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  LensC* L = reinterpret_cast<LensC*>(smem);
  LightC* G = reinterpret_cast<LightC*>(L + P.n_lens);
  double* pool = reinterpret_cast<double*>(G + P.n_light);
  stage_math_tables();
  load_or_stage(P, row, prepared, bl, n_series, L, G, pool);
  __syncthreads();

  const int npix = P.nx * P.ny;
  const int ss = SS > 0 ? SS : P.ss;
  const int64_t sub_stride = (int64_t)P.nx * ss;
  const double inv_ss = 1.0 / ss;
  // flux_evaluate_indexes: only the listed pixels are visited (the others were zero-filled by the host side),
  // so masked pixels cost nothing and warps stay full
  const int q_end = min(MASKED ? P.n_eval : npix, (int)(blockIdx.x + 1) * ppt);
  // The pixel loop is WARP-UNIFORM: every thread of the CTA makes the same number of trips (a thread past the end
  // re-evaluates the last pixel and skips the store).  Inside a loop whose trip count depends on the thread the
  // compiler keeps the polynomial coefficients of the kernel parameters in vector registers; in uniform control flow
  // they live in uniform registers, and a DFMA with a uniform operand reads two vector registers instead of three.
  for (int q0 = blockIdx.x * ppt; q0 < q_end; q0 += kImageBlock) {
    const int q = min(q0 + (int)threadIdx.x, q_end - 1);
    const bool live = q0 + (int)threadIdx.x < q_end;
    const int p = MASKED ? __ldg(P.eval_list + q) : q;
    const int py = p / P.nx, px = p - py * P.nx;
    double acc[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) acc[k] = 0.0;
    // coordinates of the pixel's first sub-pixel; one row further per i (no index arithmetic inside the loops)
    const int64_t base0 = ((int64_t)py * ss) * sub_stride + (int64_t)px * ss;
    const double* xrow = P.x_sub + base0;
    const double* yrow = P.y_sub + base0;
    for (int i = 0; i < ss; ++i, xrow += sub_stride, yrow += sub_stride) {
      double racc[NSLOT];
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) racc[k] = 0.0;
      const int64_t base = base0 + i * sub_stride;  // (only the sub-pixel outputs of SUB use it)
      for (int j0 = 0; j0 < ss; j0 += J) {
        constexpr bool kSingle = LT::N >= 0;  // specialised kernels are single-plane
        double x[J], y[J], bx[J], by[J], v[J];
        Rays<J> r;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          x[j] = __ldg(xrow + j0 + j);
          y[j] = __ldg(yrow + j0 + j);
          ray_init<J, kSingle>(P.ray, j, x[j], y[j], r);
        }
        if (kSingle) {
          // the image-plane coordinates are needed again only after the series (cheap profiles, beta = theta - alpha,
          // unlensed light): read them a second time (L1 hit) instead of keeping 4 J registers alive across it
          ray_apply_head<FULL, J, LT>(P.ray, L, pool, r);
#pragma unroll
          for (int j = 0; j < J; ++j) {
            x[j] = ldg_again(xrow + j0 + j);
            y[j] = ldg_again(yrow + j0 + j);
            r.x[j] = x[j];
            r.y[j] = y[j];
          }
          ray_apply_tail<FULL, J, LT>(P.ray, L, pool, r);
#pragma unroll
          for (int j = 0; j < J; ++j) {
            bx[j] = fma(-r.ax[j], P.ray.alpha_scaling, x[j]);
            by[j] = fma(-r.ay[j], P.ray.alpha_scaling, y[j]);
          }
        } else {
          ray_apply<FULL, J, LT>(P.ray, L, pool, P.n_lens, r);
#pragma unroll
          for (int j = 0; j < J; ++j) ray_finish(P.ray, r, j, bx[j], by[j]);
        }
        unsigned written = 0;  // slots whose sub-pixels of this batch hold a value already (sub_out only)
        auto one_light = [&](int type, const LightC& Gg) {
          const int slot = Gg.slot;
          if (slot < 0) return;
          if (!SUB) {  // the amplitude and the sum over the J sub-pixels ride on the accumulation FMAs
            double vs = 0.0;
            if (Gg.lensed)
              light_accum<J>(type, Gg, bx, by, NSLOT == 1 ? racc[0] : vs, P.ray.mk);
            else
              light_accum<J>(type, Gg, x, y, NSLOT == 1 ? racc[0] : vs, P.ray.mk);
            if (NSLOT > 1) {
#pragma unroll
              for (int k = 0; k < NSLOT; ++k) racc[k] += (slot == k) ? vs : 0.0;
            }
            return;
          }
          if (Gg.lensed)
            light_eval<J>(type, Gg, bx, by, v);
          else
            light_eval<J>(type, Gg, x, y, v);
          {  // supersampled convolution: keep the sub-pixel fluxes [set][slot][ny ss][nx ss]
            double* dsub = sub_out + (bl * P.n_slots + slot) * ((int64_t)npix * ss * ss) + base + j0;
#pragma unroll
            for (int j = 0; j < J; ++j)
              if (live) dsub[j] = ((written >> slot) & 1u) ? dsub[j] + v[j] : v[j];
            written |= 1u << slot;
          }
          double vs = v[0];
#pragma unroll
          for (int j = 1; j < J; ++j) vs += v[j];
#pragma unroll
          for (int k = 0; k < NSLOT; ++k) racc[k] += (slot == k) ? vs : 0.0;
        };
        if (GT::N >= 0) {
#pragma unroll
          for (int g = 0; g < GT::N; ++g) one_light(GT::at(g), G[g]);
        } else {
          for (int g = 0; g < P.n_light; ++g) one_light(G[g].type, G[g]);
        }
        if (SUB) {
          for (int k = 0; k < P.n_slots; ++k) {
            if ((written >> k) & 1u) continue;  // a slot without any component is an all-zero image
            double* dsub = sub_out + (bl * P.n_slots + k) * ((int64_t)npix * ss * ss) + base + j0;
#pragma unroll
            for (int j = 0; j < J; ++j)
              if (live) dsub[j] = 0.0;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) acc[k] += racc[k] * inv_ss;
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      if (live && k < P.n_slots && (!SUB || out)) out[(bl * P.n_slots + k) * npix + p] = acc[k] * inv_ss * out_scale;
    }
  }
}

// ---- k_image_fold: the image kernel for "one EPL + linear profiles -> Sersic light" on an affine grid ----------------
// Everything that is LINEAR in the sky position is composed once per parameter set (by the CTA, from the prepared
// block) instead of being evaluated per sub-pixel:
//   * SHEAR / CONVERGENCE are linear maps: beta = M x + m0 - scal alpha_EPL(x), M = 1 - scal Gamma;
//   * the EPL's major-axis coordinates a = q x', c = y' and every Sersic's product-average coordinates
//     (xt, yt) = T (beta - centre) are affine in x, hence -- on a grid whose coordinates are affine in the sub-pixel
//     indices (PlanDev.grid_affine) -- affine in (row, column): they are carried along by ADDITIONS, one per sub-pixel
//     and quantity, and the coordinate arrays are never read;
//   * the rotation of the EPL deflection back to the sky frame, the deflection scaling and T compose into one 2 x 2
//     matrix Q per lensed light profile: (xt, yt) = affine(row, col) - s Q (t_r, t_i), with (s, t_r, t_i) straight out
//     of the series (epl_core<ROT = false>).
// Per sub-pixel this replaces 4 loads, 2 DMUL and 18 three-register DFMA by ~4 DADD, 2 DMUL and 4 DFMA.  The
// arithmetic differs from k_image's in the last bits only (the affine coordinates agree with the arrays to a few ulp:
// checked at plan creation).  Scope: single plane, exactly one EPL (or SIE evaluated as EPL) and any number of SHEAR /
// CONVERGENCE, NL <= 2 SERSIC / SERSIC_ELLIPSE components, SS in {1, 2, 3}, no sub-frame mask, no supersampled
// convolution; anything else runs k_image.  (An EPL_MULTIPOLE_M1M3M4 variant -- multipoles evaluated from the affine
// centred position, folded into the light coordinates through scal T -- was measured on BASELINE config 3 and was SLOWER
// than the specialised k_image, 74.6 vs 65.3 ms per 8192 sets, gpurun_out/r2_variants3.txt: not kept.)
struct FoldLight {
  double X0, Xi, Xj, Xj2, Y0, Yi, Yj, Yj2, Q00, Q01, Q10, Q11;
};
template <int NL>
struct FoldC {
  double A0, Ai, Aj, Aj2, C0, Ci, Cj, Cj2;
  FoldLight g[NL];
  int32_t epl, pad;
};

// 2 CTAs of 256 threads per SM at up to 128 registers: measured +2 % over 3 CTAs at 80 registers with spills
// (gpurun_out/r2_variants2.txt); the kernel is bound by operand delivery to the fp64 pipe, not by occupancy
#ifndef LSIM_FOLD_MIN_BLOCKS
#define LSIM_FOLD_MIN_BLOCKS 2
#endif
template <int NSLOT, int SS, int NL>
__global__ void __launch_bounds__(kImageBlock, LSIM_FOLD_MIN_BLOCKS)
k_image_fold(PlanDev P, const double* __restrict__ params, int64_t set0, int ppt, double out_scale,
             double* __restrict__ out, const unsigned char* __restrict__ prepared, int n_series) {
  constexpr int J = SS;
  extern __shared__ __align__(16) double smem[];
  __shared__ FoldC<NL> F;
  const int64_t bl = blockIdx.y;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  LensC* L = reinterpret_cast<LensC*>(smem);
  LightC* G = reinterpret_cast<LightC*>(L + P.n_lens);
  double* pool = reinterpret_cast<double*>(G + P.n_light);
  stage_math_tables();
  load_or_stage(P, row, prepared, bl, n_series, L, G, pool);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double sc = P.ray.alpha_scaling;
    double g00 = 0, g01 = 0, g10 = 0, g11 = 0, kx = 0, ky = 0;
    int e = 0;
    for (int l = 0; l < P.n_lens; ++l) {
      if (L[l].type == LSIM_LENS_EPL) e = l;
      if (L[l].type == LSIM_LENS_SHEAR) {  // alpha = (g1 x + g2 y + c4, g2 x - g1 y + c5)
        g00 += L[l].c[2];
        g01 += L[l].c[3];
        g10 += L[l].c[3];
        g11 -= L[l].c[2];
        kx += L[l].c[4];
        ky += L[l].c[5];
      }
      if (L[l].type == LSIM_LENS_CONVERGENCE) {  // alpha = kappa (x - ra_0, y - dec_0)
        g00 += L[l].c[2];
        g11 += L[l].c[2];
        kx -= L[l].c[2] * L[l].c[0];
        ky -= L[l].c[2] * L[l].c[1];
      }
    }
    const double m00 = 1.0 - sc * g00, m01 = -sc * g01, m10 = -sc * g10, m11 = 1.0 - sc * g11;
    const double m0x = -sc * kx, m0y = -sc * ky;
    const LensC& E = L[e];
    F.epl = e;
    F.A0 = E.c[19] * P.gx0 + E.c[20] * P.gy0 + E.c[21];
    F.Ai = E.c[19] * P.gxi + E.c[20] * P.gyi;
    F.Aj = E.c[19] * P.gxj + E.c[20] * P.gyj;
    F.Aj2 = 2.0 * F.Aj;
    F.C0 = E.c[2] * P.gy0 - E.c[3] * P.gx0 + E.c[22];
    F.Ci = E.c[2] * P.gyi - E.c[3] * P.gxi;
    F.Cj = E.c[2] * P.gyj - E.c[3] * P.gxj;
    F.Cj2 = 2.0 * F.Cj;
    const double cph = E.c[2], sph = E.c[3];
    for (int g = 0; g < NL; ++g) {
      const double t00 = G[g].c[2], t01 = G[g].c[3], t11 = G[g].c[4];
      double p00 = t00, p01 = t01, p10 = t01, p11 = t11, p0x = G[g].c[11], p0y = G[g].c[12];
      FoldLight& f = F.g[g];
      f.Q00 = f.Q01 = f.Q10 = f.Q11 = 0.0;
      if (G[g].lensed) {
        p00 = t00 * m00 + t01 * m10;
        p01 = t00 * m01 + t01 * m11;
        p10 = t01 * m00 + t11 * m10;
        p11 = t01 * m01 + t11 * m11;
        p0x += t00 * m0x + t01 * m0y;
        p0y += t01 * m0x + t11 * m0y;
        // alpha_EPL = s Rot (t_r, t_i), Rot = [[cos, -sin], [sin, cos]]:  Q = scal T Rot
        f.Q00 = sc * (t00 * cph + t01 * sph);
        f.Q01 = sc * (t01 * cph - t00 * sph);
        f.Q10 = sc * (t01 * cph + t11 * sph);
        f.Q11 = sc * (t11 * cph - t01 * sph);
      }
      f.X0 = p00 * P.gx0 + p01 * P.gy0 + p0x;
      f.Xi = p00 * P.gxi + p01 * P.gyi;
      f.Xj = p00 * P.gxj + p01 * P.gyj;
      f.Xj2 = 2.0 * f.Xj;
      f.Y0 = p10 * P.gx0 + p11 * P.gy0 + p0y;
      f.Yi = p10 * P.gxi + p11 * P.gyi;
      f.Yj = p10 * P.gxj + p11 * P.gyj;
      f.Yj2 = 2.0 * f.Yj;
    }
  }
  __syncthreads();

  const int npix = P.nx * P.ny;
  const double inv_ss = 1.0 / SS;
  const LensC& E = L[F.epl];
  const double* coef = pool + E.coef_off;
  const int q_end = min(npix, (int)(blockIdx.x + 1) * ppt);
  for (int q0 = blockIdx.x * ppt; q0 < q_end; q0 += kImageBlock) {  // warp-uniform trip count, like k_image
    const int p = min(q0 + (int)threadIdx.x, q_end - 1);
    const bool live = q0 + (int)threadIdx.x < q_end;
    const int py = p / P.nx, px = p - py * P.nx;
    const double rd = (double)(py * SS), cd = (double)(px * SS);
    // the affine quantities at the pixel's first sub-pixel; one addition per row and per column from there
    double ra = fma(rd, F.Ai, fma(cd, F.Aj, F.A0)), rc = fma(rd, F.Ci, fma(cd, F.Cj, F.C0));
    double rx[NL], ry[NL];
#pragma unroll
    for (int g = 0; g < NL; ++g) {
      rx[g] = fma(rd, F.g[g].Xi, fma(cd, F.g[g].Xj, F.g[g].X0));
      ry[g] = fma(rd, F.g[g].Yi, fma(cd, F.g[g].Yj, F.g[g].Y0));
    }
    double acc[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) acc[k] = 0.0;
    // rows unrolled: with a rolled loop the compiler spills a handful of per-set constants inside it
    // (+4 % measured, gpurun_out/r2_variants4.txt)
#pragma unroll
    for (int i = 0; i < SS; ++i) {
      double racc[NSLOT];
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) racc[k] = 0.0;
      double a[J], c[J], s[J], tr[J], ti[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        a[j] = j == 0 ? ra : (j == 1 ? ra + F.Aj : ra + F.Aj2);
        c[j] = j == 0 ? rc : (j == 1 ? rc + F.Cj : rc + F.Cj2);
      }
      epl_core<J, MathK, false>(E, coef, a, c, s, tr, ti, P.ray.mk);
#pragma unroll
      for (int g = 0; g < NL; ++g) {
        const FoldLight& f = F.g[g];
        double xt[J], yt[J], e[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          xt[j] = j == 0 ? rx[g] : (j == 1 ? rx[g] + f.Xj : rx[g] + f.Xj2);
          yt[j] = j == 0 ? ry[g] : (j == 1 ? ry[g] + f.Yj : ry[g] + f.Yj2);
        }
        if (G[g].lensed) {
#pragma unroll
          for (int j = 0; j < J; ++j) {
            xt[j] = fma(-s[j], fma(f.Q00, tr[j], f.Q01 * ti[j]), xt[j]);
            yt[j] = fma(-s[j], fma(f.Q10, tr[j], f.Q11 * ti[j]), yt[j]);
          }
        }
        sersic_core<J>(G[g], xt, yt, e, P.ray.mk);
        const double amp = G[g].c[10];
        const int slot = G[g].slot;
        if (slot >= 0) {  // the amplitude and the sum over the row ride on the accumulation FMAs
          if (NSLOT == 1) {
#pragma unroll
            for (int j = 0; j < J; ++j) racc[0] = fma(amp, e[j], racc[0]);
          } else {
            double vs = 0.0;
#pragma unroll
            for (int j = 0; j < J; ++j) vs = fma(amp, e[j], vs);
#pragma unroll
            for (int k = 0; k < NSLOT; ++k) racc[k] += (slot == k) ? vs : 0.0;
          }
        }
        rx[g] += f.Xi;
        ry[g] += f.Yi;
      }
      ra += F.Ai;
      rc += F.Ci;
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) acc[k] += racc[k] * inv_ss;
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k)
      if (live && k < P.n_slots) out[(bl * P.n_slots + k) * npix + p] = acc[k] * inv_ss * out_scale;
  }
}

// ---- k_rebin: Numerics.re_size_convolve for a caller-supplied supersampled flux ------------------------------
// flux [n_img][ny*ss][nx*ss] -> out [n_img][ny][nx] = block mean (* scale)  (grid.py:101-117, image_util.py:78-98)
static __global__ void __launch_bounds__(kBlock)
k_rebin(const double* __restrict__ flux, double* __restrict__ out, int nx, int ny, int ss, double scale) {
  const int npix = nx * ny;
  const int64_t nsub = (int64_t)npix * ss * ss;
  const double* f = flux + (int64_t)blockIdx.y * nsub;
  const double inv_ss = 1.0 / ss;
  for (int p = blockIdx.x * kBlock + threadIdx.x; p < npix; p += gridDim.x * kBlock) {
    const int py = p / nx, px = p - py * nx;
    double acc = 0.0;
    for (int i = 0; i < ss; ++i) {
      double r = 0.0;
      const double* row = f + ((int64_t)py * ss + i) * ((int64_t)nx * ss) + (int64_t)px * ss;
      for (int j = 0; j < ss; ++j) r += row[j];
      acc += r * inv_ss;
    }
    out[(int64_t)blockIdx.y * npix + p] = acc * inv_ss * scale;
  }
}

// ---- k_conv_direct: "same" 2-D convolution, runtime K, shared-memory tile + halo -----------------------------
// One CTA computes a kConvT x kConvT output tile of one image; thread (tx, ty) produces kConvR rows.
// mode ZERO: out = sum_{i,j} in[y - i + r, x - j + r] k[i,j] with zeros outside (fftconvolve "same",
// convolution.py:73-96); mode EDGE: the image is edge-replicated (convolution.py:216-236).
constexpr int kConvT = 32;
constexpr int kConvR = 4;  // rows per thread: 256 threads = 32 x 8, 8*4 = 32 rows

__device__ __forceinline__ double conv_fetch(const double* img, int nx, int ny, int x, int y, int mode) {
  if (mode == LSIM_CONV_EDGE) {
    x = min(max(x, 0), nx - 1);
    y = min(max(y, 0), ny - 1);
    return img[(int64_t)y * nx + x];
  }
  return (x >= 0 && x < nx && y >= 0 && y < ny) ? img[(int64_t)y * nx + x] : 0.0;
}

static __global__ void __launch_bounds__(kBlock)
k_conv_direct(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int K,
              const double* __restrict__ kernel, int mode, double scale, int accumulate) {
  extern __shared__ double smem[];
  const int r = (K - 1) / 2;
  const int TW = kConvT + K - 1;     // tile width incl. halo
  const int pitch = TW | 1;          // odd pitch: conflict-free column walks
  double* tile = smem;               // [TW][pitch]
  double* w = smem + TW * pitch;     // [K*K], flipped so that the inner loop is a correlation
  const int tiles_x = (nx + kConvT - 1) / kConvT;
  const int tx0 = (blockIdx.x % tiles_x) * kConvT, ty0 = (blockIdx.x / tiles_x) * kConvT;
  const double* img = in + (int64_t)blockIdx.y * nx * ny;
  for (int i = threadIdx.x; i < K * K; i += kBlock) w[i] = kernel[K * K - 1 - i];
  for (int i = threadIdx.x; i < TW * TW; i += kBlock) {
    int ly = i / TW, lx = i - ly * TW;
    tile[ly * pitch + lx] = conv_fetch(img, nx, ny, tx0 + lx - r, ty0 + ly - r, mode);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double acc[kConvR];
#pragma unroll
  for (int o = 0; o < kConvR; ++o) acc[o] = 0.0;
  for (int dy = 0; dy < K; ++dy) {
    const double* trow = tile + (ty * kConvR + dy) * pitch + tx;
    const double* wrow = w + dy * K;
    for (int dx = 0; dx < K; ++dx) {
      const double wv = wrow[dx];
#pragma unroll
      for (int o = 0; o < kConvR; ++o) acc[o] = fma(wv, trow[o * pitch + dx], acc[o]);
    }
  }
  double* dst = out + (int64_t)blockIdx.y * nx * ny;
#pragma unroll
  for (int o = 0; o < kConvR; ++o) {
    int x = tx0 + tx, y = ty0 + ty * kConvR + o;
    if (x < nx && y < ny) {
      double* d = dst + (int64_t)y * nx + x;
      *d = accumulate ? fma(acc[o], scale, *d) : acc[o] * scale;
    }
  }
}

// Register-tiled direct convolution for K <= KT (the common PSF sizes; the kernel is zero-padded to KT x KT).
// A warp owns a 32-column x R-row strip, NW warps stack to a 32 x (R NW) tile.  For each kernel column dx a thread
// reads its R + KT - 1 input values of that column from shared memory and uses each of them for up to KT of the
// R x KT FMAs.  The weights are passed BY VALUE as a kernel parameter (constant bank, read through uniform
// registers): every DFMA then has the form  acc[o] += v.reuse * UR  -- one input value held in the operand reuse
// cache, a uniform-register weight and ONE fresh vector register (the accumulator).  That matters on B200, where a
// DFMA with three fresh register operands issues every 3 cycles instead of 2 (profiles/r1_dfma_operand_probe.md);
// with shared-memory weights (the previous version: two fresh registers + reuse) the same tiling ran 19 % slower.
// wT[dx][dy], flipped (correlation form), zero padded from K to KT.
template <int KT>
struct ConvWeights {
  double w[KT * KT];
};

// (2 resident CTAs per SM; 3 measured the same to 0.2 %, gpurun_out/r2_variants3.txt)
template <int KT, int R, int NW>
__global__ void __launch_bounds__(NW * 32, 2)
k_conv_tiled_u(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int mode, double scale,
               int accumulate, const __grid_constant__ ConvWeights<KT> W) {
  constexpr int NT = NW * 32, TH = R * NW, HALO = KT - 1, TWD = 32 + HALO, PITCH = TWD | 1;
  extern __shared__ double smem[];
  double* tile = smem;  // [TH + HALO][PITCH]
  const int rT = HALO / 2;
  const int tiles_x = (nx + 31) / 32;
  const int tx0 = (blockIdx.x % tiles_x) * 32, ty0 = (blockIdx.x / tiles_x) * TH;
  const double* img = in + (int64_t)blockIdx.y * nx * ny;
  for (int i = threadIdx.x; i < (TH + HALO) * TWD; i += NT) {
    const int ly = i / TWD, lx = i - ly * TWD;
    tile[ly * PITCH + lx] = conv_fetch(img, nx, ny, tx0 + lx - rT, ty0 + ly - rT, mode);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[R];
#pragma unroll
  for (int o = 0; o < R; ++o) acc[o] = 0.0;
  const double* tcol = tile + (warp * R) * PITCH + lane;
#pragma unroll 1
  for (int dx = 0; dx < KT; ++dx) {
    double v[R + KT - 1];
#pragma unroll
    for (int i = 0; i < R + KT - 1; ++i) v[i] = tcol[i * PITCH + dx];
#pragma unroll
    for (int dy = 0; dy < KT; ++dy) {
      const double wv = W.w[dx * KT + dy];
#pragma unroll
      for (int o = 0; o < R; ++o) acc[o] = fma(wv, v[o + dy], acc[o]);
    }
  }
  double* dst = out + (int64_t)blockIdx.y * nx * ny;
  const int x = tx0 + lane;
#pragma unroll
  for (int o = 0; o < R; ++o) {
    const int y = ty0 + warp * R + o;
    if (x < nx && y < ny) {
      double* d = dst + (int64_t)y * nx + x;
      *d = accumulate ? fma(acc[o], scale, *d) : acc[o] * scale;
    }
  }
}

// Fast-FIR form of the same direct convolution, 2-parallel in both directions.
// One direction: with the correlation weights f[d] split into even / odd taps and the input into even / odd samples,
//   out[n] = sum_d f[d] v[n+d]   becomes, for the output pair (2j, 2j+1),
//   out[2j]   = sum_t ka[t] v[2(j+t)]   + S[j],   ka[t] = f[2t] - f[2t-1]      ((KT+1)/2 taps)
//   out[2j+1] = sum_t kb[t] v[2(j+t)+1] + S[j],   kb[t] = f[2t] - f[2t+1]      ((KT+1)/2 taps)
//   S[j]      = sum_t f[2t+1] (v[2(j+t)+1] + v[2(j+t)+2])                       ((KT-1)/2 taps)
// -- three half-length products for two outputs instead of four: (3 KT + 1) / 4 = 16 multiply-adds per output for
// KT = 21 instead of 21, plus one addition per two input values to form the sums.
// Row direction: applied inside the register column of k_conv_tiled_u's inner product (the loaded column v[] is split
// in registers; the sums are recomputed per kernel column).  Column direction: the tile is staged as two half-width
// planes P0[l][j] = in[l][2j], P1[l][j] = in[l][2j+1]; a lane owns the output column pair (2j, 2j+1) and runs three
// passes of (KT+1)/2, (KT+1)/2, (KT-1)/2 kernel columns --
//   both columns: f1_x on P1[j] + P0[j+1] (summed while loading),  odd columns: kb_x on P1,  even columns: ka_x on P0
// -- nine transformed kernels in all (host: conv_fxy_weights, once per launch).  ((3 KT + 1) / 4)^2 = 256
// multiply-adds per output pixel for KT = 21 instead of 441, ~290 fp64 instructions with the additions.
// Same operand pattern as k_conv_tiled_u (value in the reuse cache x uniform-register weight + accumulator), 6
// multiply-adds per shared-memory load.  Accumulators: the shared product is folded into the other passes' start
// values, at most 5 (R/2) live.
// Rounding: ka, kb cancel against S term by term, so an output carries a few ulp of the local image scale instead of
// ~1 -- far inside the 1e-10 parity bar, not bit-identical to k_conv_tiled_u (LSIM_CONV_NO_FF=1 selects that one;
// tests/test_gpu_parity.py compares the two).
// Measured (1184 images 180 x 180, K = 21): 0.84 ms at 80 % of the fp64 pipe, k_conv_tiled_u 1.19 ms at 85 %.
// Variants measured and dropped (profiles/r2_conv_fastfir.md): all nine products on quarter-resolution planes in shared
// memory (11 x 11 sub-kernels leave 4 multiply-adds per shared-memory load: shared-memory bound, 1.59 ms); the row
// split alone (1.02 ms); the sum plane built in shared memory between two barriers (0.90 ms).
template <int KT>
struct ConvWeightsFXY {
  static constexpr int M = (KT + 1) / 2, WS = 3 * M - 1;
  double w[3 * M * WS];  // [x type b: 0 ka, 1 kb, 2 f1][kernel column tx][row weights ka[M], kb[M], f1[M-1]]
};

template <int KT, int R, int B, int PITCH, int PLANE>
__device__ __forceinline__ void conv_fxy_pass(const double* __restrict__ tcol, const ConvWeightsFXY<KT>& W,
                                              double (&ae)[R / 2], double (&ao)[R / 2], double (&sh)[R / 2]) {
  constexpr int M = (KT + 1) / 2, H = R / 2, WS = 3 * M - 1, NTX = (B == 2) ? M - 1 : M;
#pragma unroll 1
  for (int tx = 0; tx < NTX; ++tx) {
    double v[R + KT - 1];
#pragma unroll
    for (int i = 0; i < R + KT - 1; ++i)
      v[i] = (B == 2) ? tcol[PLANE + i * PITCH + tx] + tcol[i * PITCH + tx + 1] : tcol[(B == 1 ? PLANE : 0) + i * PITCH + tx];
    double ps[H + M - 2];
#pragma unroll
    for (int i = 0; i < H + M - 2; ++i) ps[i] = v[2 * i + 1] + v[2 * i + 2];
    const int wb0 = (B * M + tx) * WS;
#pragma unroll
    for (int t = 0; t < M; ++t) {
      const double wa = W.w[wb0 + t], wb = W.w[wb0 + M + t];
#pragma unroll
      for (int j = 0; j < H; ++j) ae[j] = fma(wa, v[2 * (j + t)], ae[j]);
#pragma unroll
      for (int j = 0; j < H; ++j) ao[j] = fma(wb, v[2 * (j + t) + 1], ao[j]);
      if (t < M - 1) {
        const double ws = W.w[wb0 + 2 * M + t];
#pragma unroll
        for (int j = 0; j < H; ++j) sh[j] = fma(ws, ps[j + t], sh[j]);
      }
    }
  }
}

template <int KT, int R, int NW>
__global__ void __launch_bounds__(NW * 32, 2)
k_conv_tiled_fxy(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int mode, double scale,
                 int accumulate, const __grid_constant__ ConvWeightsFXY<KT> W) {
  static_assert(R % 2 == 0 && KT % 2 == 1, "row pairs, odd kernel");
  constexpr int NT = NW * 32, TH = R * NW, HALO = KT - 1, M = (KT + 1) / 2, H = R / 2, PCQ = 32 + M - 1, PITCH = PCQ | 1,
                ROWS = TH + HALO, PLANE = ROWS * PITCH, rT = HALO / 2;
  extern __shared__ double smem[];  // P0 [ROWS][PITCH], P1 [ROWS][PITCH]
  const int tiles_x = (nx + 63) / 64;
  const int tx0 = (blockIdx.x % tiles_x) * 64, ty0 = (blockIdx.x / tiles_x) * TH;
  const double* img = in + (int64_t)blockIdx.y * nx * ny;
  // U loads in flight per thread (one load per round trip left this phase at 16 % of the stall samples)
  constexpr int U = 8;
  for (int base = threadIdx.x; base < ROWS * 2 * PCQ; base += NT * U) {
    double t[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * NT, ly = i / (2 * PCQ), lx = i - ly * (2 * PCQ);
      t[u] = i < ROWS * 2 * PCQ ? conv_fetch(img, nx, ny, tx0 + lx - rT, ty0 + ly - rT, mode) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * NT, ly = i / (2 * PCQ), lx = i - ly * (2 * PCQ);
      if (i < ROWS * 2 * PCQ) smem[(lx & 1) * PLANE + ly * PITCH + (lx >> 1)] = t[u];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* tcol = smem + (warp * R) * PITCH + lane;
  double* dst = out + (int64_t)blockIdx.y * nx * ny;
  const int x = tx0 + 2 * lane;
  auto store = [&](int beta, const double (&ae)[H], const double (&ao)[H], const double (&sh)[H]) {
#pragma unroll
    for (int o = 0; o < R; ++o) {
      const int y = ty0 + warp * R + o;
      const double acc = ((o & 1) ? ao[o / 2] : ae[o / 2]) + sh[o / 2];
      if (x + beta < nx && y < ny) {
        double* d = dst + (int64_t)y * nx + x + beta;
        *d = accumulate ? fma(acc, scale, *d) : acc * scale;
      }
    }
  };
  // the product shared by both columns (its input P1[j] + P0[j+1] is formed while loading: a sum plane in shared
  // memory cost two barriers and occupancy, measured slower), folded over the row split: te / to = what the even /
  // odd rows of both columns start from
  double te[H], to[H], ts[H];
#pragma unroll
  for (int j = 0; j < H; ++j) te[j] = to[j] = ts[j] = 0.0;
  conv_fxy_pass<KT, R, 2, PITCH, PLANE>(tcol, W, te, to, ts);
#pragma unroll
  for (int j = 0; j < H; ++j) {
    te[j] += ts[j];
    to[j] += ts[j];
  }
  {  // odd columns
    double be[H], bo[H], bs[H];
#pragma unroll
    for (int j = 0; j < H; ++j) be[j] = te[j], bo[j] = to[j], bs[j] = 0.0;
    conv_fxy_pass<KT, R, 1, PITCH, PLANE>(tcol, W, be, bo, bs);
    store(1, be, bo, bs);
  }
  // even columns
#pragma unroll
  for (int j = 0; j < H; ++j) ts[j] = 0.0;
  conv_fxy_pass<KT, R, 0, PITCH, PLANE>(tcol, W, te, to, ts);
  store(0, te, to, ts);
}

// Separable variant (kernel == outer(k1, k1), e.g. the Gaussian PSF): row pass into shared memory, column pass out.
static __global__ void __launch_bounds__(kBlock)
k_conv_separable(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int K,
                 const double* __restrict__ k1, int mode, double scale, int accumulate) {
  extern __shared__ double smem[];
  const int r = (K - 1) / 2;
  const int TW = kConvT + K - 1;
  const int pitch = TW | 1;
  double* tile = smem;                 // [TW rows][pitch]   input with halo
  double* rowp = tile + TW * pitch;    // [TW rows][kConvT+1] after the row pass
  double* w = rowp + TW * (kConvT + 1);
  const int tiles_x = (nx + kConvT - 1) / kConvT;
  const int tx0 = (blockIdx.x % tiles_x) * kConvT, ty0 = (blockIdx.x / tiles_x) * kConvT;
  const double* img = in + (int64_t)blockIdx.y * nx * ny;
  for (int i = threadIdx.x; i < K; i += kBlock) w[i] = k1[K - 1 - i];
  for (int i = threadIdx.x; i < TW * TW; i += kBlock) {
    int ly = i / TW, lx = i - ly * TW;
    tile[ly * pitch + lx] = conv_fetch(img, nx, ny, tx0 + lx - r, ty0 + ly - r, mode);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TW * kConvT; i += kBlock) {
    int ly = i / kConvT, lx = i - ly * kConvT;
    double a = 0.0;
    for (int dx = 0; dx < K; ++dx) a = fma(w[dx], tile[ly * pitch + lx + dx], a);
    rowp[ly * (kConvT + 1) + lx] = a;
  }
  __syncthreads();
  double* dst = out + (int64_t)blockIdx.y * nx * ny;
  for (int i = threadIdx.x; i < kConvT * kConvT; i += kBlock) {
    int ly = i / kConvT, lx = i - ly * kConvT;
    double a = 0.0;
    for (int dy = 0; dy < K; ++dy) a = fma(w[dy], rowp[(ly + dy) * (kConvT + 1) + lx], a);
    int x = tx0 + lx, y = ty0 + ly;
    if (x < nx && y < ny) {
      double* d = dst + (int64_t)y * nx + x;
      *d = accumulate ? fma(a, scale, *d) : a * scale;
    }
  }
}

// ---- k_point_sources: PSF stamps of the point-source images, added to the finished slot images ----------------------
// grid (n_slots, sets): one CTA owns one output image, walks the point images assigned to it one after the other
// (no atomics, deterministic order) and its threads share the pixels of each footprint.
// Pixel value = amp * sum over the ss x ss sub-pixels of the bilinear interpolation of the supersampled kernel at
//   (row, col) = (r + Kp/2 - y_sub, c + Kp/2 - x_sub),  x_sub = x_pix ss + (ss-1)/2
// with zeros outside the kernel, neighbour by neighbour (jax.scipy.ndimage.map_coordinates, order 1, mode
// "constant"), which is add_layer2image + re_size(...) * ss^2 of the reference.
__device__ __forceinline__ double ps_bilinear(const double* __restrict__ k, int K, double ky, double kx) {
  const double fy = floor(ky), fx = floor(kx);
  const double wy1 = ky - fy, wx1 = kx - fx, wy0 = 1.0 - wy1, wx0 = 1.0 - wx1;
  const int iy = (int)fy, ix = (int)fx;
  auto at = [&](int y, int x) { return (y >= 0 && y < K && x >= 0 && x < K) ? k[y * K + x] : 0.0; };
  // order of jax's accumulation: (lower, lower), (lower, upper), (upper, lower), (upper, upper); weight product first
  double s = (wy0 * wx0) * at(iy, ix);
  s += (wy0 * wx1) * at(iy, ix + 1);
  s += (wy1 * wx0) * at(iy + 1, ix);
  s += (wy1 * wx1) * at(iy + 1, ix + 1);
  return s;
}

static __global__ void __launch_bounds__(kBlock)
k_point_sources(PlanDev P, const double* __restrict__ params, int64_t set0, double* __restrict__ dst) {
  const int slot = blockIdx.x;
  const int64_t bl = blockIdx.y;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  double* img = dst + (bl * P.n_slots + slot) * ((int64_t)P.nx * P.ny);
  const int K = P.ps_K, ss = P.ps_ss, half = K / 2;
  for (int i = 0; i < P.n_points; ++i) {
    const lsim_point_desc& d = P.points[i];
    if (d.slot != slot) continue;
    const double ra = d.col_ra >= 0 ? row[d.col_ra] : d.cst_ra;
    const double dec = d.col_dec >= 0 ? row[d.col_dec] : d.cst_dec;
    const double amp = d.use_amp ? (d.col_amp >= 0 ? row[d.col_amp] : d.cst_amp) : 1.0;
    const double xp = P.ps_M[0] * ra + P.ps_M[1] * dec + P.ps_x0;
    const double yp = P.ps_M[2] * ra + P.ps_M[3] * dec + P.ps_y0;
    const double xs = xp * ss + (ss - 1) / 2.0, ys = yp * ss + (ss - 1) / 2.0;
    // sub-grid columns with a non-zero contribution: -1 < c + half - xs < K
    if (!(fabs(xs) < 1e9 && fabs(ys) < 1e9)) continue;  // nan / absurd positions paint nothing
    const int c_lo = (int)floor(xs - half - 1.0), c_hi = (int)ceil(xs - half + K);
    const int r_lo = (int)floor(ys - half - 1.0), r_hi = (int)ceil(ys - half + K);
    const int px_lo = max(0, (int)floor((double)c_lo / ss)), px_hi = min(P.nx - 1, (int)floor((double)c_hi / ss));
    const int py_lo = max(0, (int)floor((double)r_lo / ss)), py_hi = min(P.ny - 1, (int)floor((double)r_hi / ss));
    const int w = px_hi - px_lo + 1, h = py_hi - py_lo + 1;
    if (w > 0 && h > 0) {
      for (int t = threadIdx.x; t < w * h; t += kBlock) {
        const int py = py_lo + t / w, px = px_lo + t % w;
        double acc = 0.0;
        for (int sy = 0; sy < ss; ++sy)
          for (int sx = 0; sx < ss; ++sx)
            acc += ps_bilinear(P.ps_kernel, K, (py * ss + sy) + half - ys, (px * ss + sx) + half - xs);
        img[(int64_t)py * P.nx + px] += amp * acc * P.ps_scale;
      }
    }
    __syncthreads();  // the next point image may touch the same pixels from other threads
  }
}

// ---- k_ps_error_map: PSF error map of one parameter set (one CTA per set) -------------------------------------------
// err[set][pixel] = sum over flagged point-source images of bilinear(psf_variance_map) * amp_est^2  (see lsim.h)
static __global__ void __launch_bounds__(kBlock)
k_ps_error_map(PlanDev P, const double* __restrict__ params, int64_t set0, double* __restrict__ err) {
  const int64_t bl = blockIdx.x;
  const double* row = params + (set0 + bl) * P.n_param_cols;
  double* img = err + bl * ((int64_t)P.nx * P.ny);
  const int npix = P.nx * P.ny, K = P.ps_var_K, half = K / 2;
  for (int p = threadIdx.x; p < npix; p += kBlock) img[p] = 0.0;
  __syncthreads();
  for (int i = 0; i < P.n_points; ++i) {
    const lsim_point_desc& d = P.points[i];
    if (!d.error_map) continue;
    const double ra = d.col_ra >= 0 ? row[d.col_ra] : d.cst_ra;
    const double dec = d.col_dec >= 0 ? row[d.col_dec] : d.cst_dec;
    const double xp = P.ps_M[0] * ra + P.ps_M[1] * dec + P.ps_x0;
    const double yp = P.ps_M[2] * ra + P.ps_M[3] * dec + P.ps_y0;
    if (!(fabs(xp) < 1e9 && fabs(yp) < 1e9)) continue;
    // kernel_util.estimate_amp: the 5 x 5 data pixels around round(pos - 0.49999), nothing near the frame edge
    const int xi = (int)rint(xp - 0.49999), yi = (int)rint(yp - 0.49999);
    double amp = 0.0;
    if (xi > 2 && xi < P.nx - 2 && yi > 2 && yi < P.ny - 2) {  // (numPix_x, numPix_y = shape(data); frame is square)
      double s = 0.0;
      for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b) s += P.data[(int64_t)(yi - 2 + a) * P.nx + (xi - 2 + b)];
      amp = fmax(s, 0.0) / P.ps_mean_kernel;
    }
    const double a2 = amp * amp;
    const int px_lo = max(0, (int)floor(xp - half - 1.0)), px_hi = min(P.nx - 1, (int)ceil(xp - half + K));
    const int py_lo = max(0, (int)floor(yp - half - 1.0)), py_hi = min(P.ny - 1, (int)ceil(yp - half + K));
    const int w = px_hi - px_lo + 1, h = py_hi - py_lo + 1;
    if (w > 0 && h > 0 && a2 != 0.0) {
      for (int t = threadIdx.x; t < w * h; t += kBlock) {
        const int py = py_lo + t / w, px = px_lo + t % w;
        // add_layer2image(variance_map, x, y, psf_variance_map * amp^2): the kernel is scaled first
        const double fy = floor(py + half - yp), fx = floor(px + half - xp);
        const double ky = py + half - yp, kx = px + half - xp;
        const double wy1 = ky - fy, wx1 = kx - fx, wy0 = 1.0 - wy1, wx0 = 1.0 - wx1;
        const int iy = (int)fy, ix = (int)fx;
        auto at = [&](int y, int x) { return (y >= 0 && y < K && x >= 0 && x < K) ? P.ps_var[y * K + x] * a2 : 0.0; };
        double v = (wy0 * wx0) * at(iy, ix);
        v += (wy0 * wx1) * at(iy, ix + 1);
        v += (wy1 * wx0) * at(iy + 1, ix);
        v += (wy1 * wx1) * at(iy + 1, ix + 1);
        img[(int64_t)py * P.nx + px] += v;
      }
    }
    __syncthreads();
  }
}

// ---- FFT convolution helpers (large kernels; the transforms themselves are cuFFT) ----------------------------------
// zero-padded copy [n][ny][nx] -> [n][PY][PX]
static __global__ void __launch_bounds__(kBlock)
k_fft_pad(const double* __restrict__ in, double* __restrict__ pad, int nx, int ny, int PX, int PY) {
  const double* src = in + (int64_t)blockIdx.y * nx * ny;
  double* dst = pad + (int64_t)blockIdx.y * PX * PY;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < PX * PY; i += gridDim.x * kBlock) {
    const int y = i / PX, x = i - y * PX;
    dst[i] = (x < nx && y < ny) ? src[(int64_t)y * nx + x] : 0.0;
  }
}

// spectrum [n][PY][PXc] *= kernel spectrum [PY][PXc] * norm
static __global__ void __launch_bounds__(kBlock)
k_fft_mul(double2* __restrict__ spec, const double2* __restrict__ kspec, int n_spec, double norm) {
  double2* s = spec + (int64_t)blockIdx.y * n_spec;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n_spec; i += gridDim.x * kBlock) {
    const double2 a = s[i], b = kspec[i];
    s[i] = make_double2((a.x * b.x - a.y * b.y) * norm, (a.x * b.y + a.y * b.x) * norm);
  }
}

// out[n][ny][nx] (+)= scale * full[n][y + r][x + r]   ("same" part of the full linear convolution)
static __global__ void __launch_bounds__(kBlock)
k_fft_crop(const double* __restrict__ pad, double* __restrict__ out, int nx, int ny, int PX, int PY, int r,
           double scale, int accumulate) {
  const double* src = pad + (int64_t)blockIdx.y * PX * PY;
  double* dst = out + (int64_t)blockIdx.y * nx * ny;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < nx * ny; i += gridDim.x * kBlock) {
    const int y = i / nx, x = i - y * nx;
    const double v = src[(int64_t)(y + r) * PX + (x + r)];
    dst[i] = accumulate ? fma(v, scale, dst[i]) : v * scale;
  }
}

// ---- reductions -----------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum `n` per-thread values over a CTA of BLK threads; result valid in thread 0's vals[]. scratch: [n][BLK/32] doubles.
template <int N, int BLK = kBlock>
__device__ __forceinline__ void block_sum(double (&vals)[N], double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double v = warp_sum(vals[i]);
    if (lane == 0) scratch[i * (BLK / 32) + wid] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      for (int w = 0; w < BLK / 32; ++w) s += scratch[i * (BLK / 32) + w];
      vals[i] = s;
    }
  }
  __syncthreads();
}

// ---- small dense symmetric solves, one thread ---------------------------------------------------------------
// Eigenvalues of a symmetric n x n matrix by cyclic Jacobi (destroys a); n <= LSIM_MAX_SLOTS.
__device__ inline void sym_eigenvalues(double* a, int n, double* ev) {
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += a[i * n + i] * a[i * n + i];
      for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
    }
    if (!(off > 1e-34 * diag)) break;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        double apq = a[p * n + q];
        if (apq == 0.0) continue;
        double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) ev[i] = a[i * n + i];
}

// Solve m x = r by Gaussian elimination with partial pivoting (destroys m, r). Returns false if singular.
__device__ inline bool lu_solve(double* m, double* r, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(m[c * n + c]);
    for (int i = c + 1; i < n; ++i) {
      if (fabs(m[i * n + c]) > best) {
        best = fabs(m[i * n + c]);
        piv = i;
      }
    }
    if (!(best > 0.0)) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) {
        double t = m[c * n + k];
        m[c * n + k] = m[piv * n + k];
        m[piv * n + k] = t;
      }
      double t = r[c];
      r[c] = r[piv];
      r[piv] = t;
    }
    double inv = 1.0 / m[c * n + c];
    for (int i = c + 1; i < n; ++i) {
      double f = m[i * n + c] * inv;
      for (int k = c; k < n; ++k) m[i * n + k] -= f * m[c * n + k];
      r[i] -= f * r[c];
    }
  }
  for (int c = n - 1; c >= 0; --c) {
    double s = r[c];
    for (int k = c + 1; k < n; ++k) s -= m[c * n + k] * r[k];
    r[c] = s / m[c * n + c];
  }
  return true;
}

// ---- k_likelihood ---------------------------------------------------------------------------------------------
// One CTA per parameter set.  `conv` holds the convolved, scaled basis images [set][slot][pixel].
//   linear == 0: model = sum_slots conv                                   (image_model.py:162-221)
//   linear != 0: A_k = nan_to_num(conv_k[mask]); M = A W A^T, R = A W d, W = 1/C_D(data);
//                amplitudes = cond2(M) < 5/eps ? solve(M, R) (0 if non-finite) : 0; model = sum B_k A_k
//                (image_linear_solve.py:97-181,361-455, de_lens.py:10-36,91-105)
//   logL = -1/2 sum mask (model - data)^2 / C_D(model)                    (imaging_data.py:159-182)
//          C_D(model) = noise_map^2  or  max(model,0)/t_exp + sigma_bkg^2 (image_noise.py:99-137)
//   marg != 0: + 1/2 sign log|det M^-1|, -1e15 if singular               (de_lens.py:39-49)
// Round 2 tried three other shapes of this kernel (64-thread CTAs that fit beside k_image on a side stream, four sets
// per CTA sharing the plan-constant reads, and a plain side-stream overlap): all slower or no faster,
// profiles/r2_likelihood.md.
#ifndef LSIM_LIKE_UNROLL
#define LSIM_LIKE_UNROLL 4
#endif
constexpr int kLikeUnroll = LSIM_LIKE_UNROLL;

template <int NS>
__global__ void __launch_bounds__(kBlock)
k_likelihood(PlanDev P, const double* __restrict__ conv, int64_t set0, int linear, int marg, double* __restrict__ logl,
             double* __restrict__ amps, double* __restrict__ model_out, const double* __restrict__ err_map) {
  constexpr int NM = NS * (NS + 1) / 2;
  __shared__ double scratch[(NM + NS + 1) * (kBlock / 32)];
  __shared__ double s_amp[NS];
  __shared__ double s_marg;
  const int64_t bl = blockIdx.x;
  const int npix = P.nx * P.ny;
  const int ns = P.n_slots;
  const double* A = conv + bl * ns * (int64_t)npix;
  // PSF error map of this set (sigma^2, or null): added to C_D(data) in the weights and as |.| to C_D(model) in chi^2
  const double* E = err_map ? err_map + bl * (int64_t)npix : nullptr;

  if (linear) {
    double acc[NM + NS];
#pragma unroll
    for (int i = 0; i < NM + NS; ++i) acc[i] = 0.0;
    // 4 pixels per thread per trip: all loads of a trip are issued before any arithmetic (latency hiding)
    for (int p0 = threadIdx.x; p0 < npix; p0 += kLikeUnroll * kBlock) {
      double a[kLikeUnroll][NS], wgt[kLikeUnroll], dat[kLikeUnroll];
#pragma unroll
      for (int u = 0; u < kLikeUnroll; ++u) {
        const int p = p0 + u * kBlock;
        const bool use = p < npix && (!P.mask || P.mask[p]);
        wgt[u] = use ? __ldg(P.w_data + p) : 0.0;  // masked / out-of-range pixels get zero weight
        if (E && use) wgt[u] = 1.0 / (1.0 / wgt[u] + E[p]);  // 1 / (C_D + model_error), image_linear_solve.py:150
        dat[u] = use ? __ldg(P.data + p) : 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k) a[u][k] = (use && k < ns) ? A[(int64_t)k * npix + p] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kLikeUnroll; ++u) {
#pragma unroll
        for (int k = 0; k < NS; ++k) a[u][k] = nan_to_num(a[u][k]);
        const double wd = wgt[u] * dat[u];
        int m = 0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
#pragma unroll
          for (int j = i; j < NS; ++j) acc[m++] += wgt[u] * a[u][i] * a[u][j];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[NM + i] += wd * a[u][i];
      }
    }
    block_sum(acc, scratch);
    if (threadIdx.x == 0) {
      double M[NS * NS], M2[NS * NS], R[NS], ev[NS];
      int m = 0;
      for (int i = 0; i < NS; ++i)
        for (int j = i; j < NS; ++j) {
          M[i * NS + j] = acc[m];
          M[j * NS + i] = acc[m];
          ++m;
        }
      // work on the leading ns x ns block, repacked contiguously
      for (int i = 0; i < ns; ++i)
        for (int j = 0; j < ns; ++j) M2[i * ns + j] = M[i * NS + j];
      for (int i = 0; i < ns; ++i) R[i] = acc[NM + i];
      double Mw[NS * NS];
      for (int i = 0; i < ns * ns; ++i) Mw[i] = M2[i];
      sym_eigenvalues(Mw, ns, ev);
      double emax = 0.0, emin = HUGE_VAL, logdet = 0.0, sign = 1.0;
      for (int i = 0; i < ns; ++i) {
        double e = fabs(ev[i]);
        emax = fmax(emax, e);
        emin = fmin(emin, e);
        logdet += log(e);
        if (ev[i] < 0) sign = -sign;
        if (ev[i] == 0) sign = 0.0;
      }
      bool stable = (emax / emin) < 5.0 / DBL_EPSILON;  // NaN compares false, like jnp
      bool ok = stable && lu_solve(M2, R, ns);
      if (ok)
        for (int i = 0; i < ns; ++i) ok = ok && (fabs(R[i]) <= DBL_MAX);
      for (int i = 0; i < NS; ++i) s_amp[i] = (ok && i < ns) ? R[i] : 0.0;
      // log det(M_inv) = -log det(M); M_inv = 0 (unstable) -> sign 0 -> -1e15
      s_marg = (stable && sign != 0.0 && (logdet == logdet)) ? sign * (-logdet) / 2 : -1e15;
    }
    __syncthreads();
  }

  double chi[1] = {0.0};
  // second pass from the last trip to the first: the lines the first pass read last are the ones most likely still in
  // L2 (4 resident CTAs x 148 SMs x 259 KB slightly exceed it, so a forward second pass misses everywhere): +1.2 % on
  // the headline step, same-box A/B (profiles/r2_variants_ab.txt)
  const int n_trips = (npix + kLikeUnroll * kBlock - 1) / (kLikeUnroll * kBlock);
  for (int trip = n_trips - 1; trip >= 0; --trip) {
    const int p0 = trip * kLikeUnroll * kBlock + threadIdx.x;
    double mdl[kLikeUnroll], dat[kLikeUnroll], aux[kLikeUnroll];
    bool use[kLikeUnroll], valid[kLikeUnroll];
#pragma unroll
    for (int u = 0; u < kLikeUnroll; ++u) {
      const int p = p0 + u * kBlock;
      valid[u] = p < npix;
      use[u] = valid[u] && (!P.mask || P.mask[p]);
      dat[u] = use[u] ? __ldg(P.data + p) : 0.0;
      aux[u] = use[u] ? (P.noise_map2 ? __ldg(P.noise_map2 + p) : __ldg(P.inv_exposure + p)) : 1.0;
      double m = 0.0;
      if (linear) {
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (k < ns && use[u]) m += s_amp[k] * nan_to_num(A[(int64_t)k * npix + p]);
      } else if (valid[u]) {
        for (int k = 0; k < ns; ++k) m += A[(int64_t)k * npix + p];
      }
      mdl[u] = m;
    }
#pragma unroll
    for (int u = 0; u < kLikeUnroll; ++u) {
      const int p = p0 + u * kBlock;
      if (model_out && valid[u]) model_out[(set0 + bl) * npix + p] = mdl[u];
      if (use[u]) {
        double cd = P.noise_map2 ? aux[u] : fma(fmax(mdl[u], 0.0), aux[u], P.bkg2);
        if (E) cd += fabs(E[p]);  // imaging_data.py:180
        const double d = mdl[u] - dat[u];
        chi[0] += d * d / cd;
      }
    }
  }
  block_sum(chi, scratch);
  if (threadIdx.x == 0) {
    double l = -chi[0] / 2 + P.logl_offset;
    if (linear && marg) l += s_marg;
    logl[set0 + bl] = l;
    if (amps && linear)
      for (int k = 0; k < ns; ++k) amps[(set0 + bl) * ns + k] = s_amp[k];
  }
}

// ---- particle swarm: best-particle bookkeeping and move ------------------------------------------------------------
// k_pso_best (one CTA): argmax of fitness (lowest index on ties), global best update, then personal bests.
constexpr int kPsoBlock = 1024;
static __global__ void __launch_bounds__(kPsoBlock)
k_pso_best(const double* __restrict__ position, const double* __restrict__ fitness, double* __restrict__ pbest_pos,
           double* __restrict__ pbest_fit, double* __restrict__ gbest_pos, double* __restrict__ gbest_fit, int64_t B,
           int P) {
  __shared__ double s_val[kPsoBlock / 32];
  __shared__ int64_t s_idx[kPsoBlock / 32];
  __shared__ int64_t s_best;
  double best = -HUGE_VAL;
  int64_t arg = -1;  // -1: nothing beats -inf yet; nan never wins (comparisons are false), like the reference's `<`
  for (int64_t b = threadIdx.x; b < B; b += kPsoBlock) {
    const double f = fitness[b];
    if (f > best) {
      best = f;
      arg = b;
    }
  }
  auto better = [](double fa, int64_t ia, double fb, int64_t ib) {  // is (fb, ib) to be preferred over (fa, ia)?
    if (ib < 0) return false;
    if (ia < 0) return true;
    return fb > fa || (fb == fa && ib < ia);
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double of = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (better(best, arg, of, oi)) {
      best = of;
      arg = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_val[threadIdx.x >> 5] = best;
    s_idx[threadIdx.x >> 5] = arg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kPsoBlock / 32; ++w)
      if (better(best, arg, s_val[w], s_idx[w])) {
        best = s_val[w];
        arg = s_idx[w];
      }
    s_best = (arg >= 0 && gbest_fit[0] < best) ? arg : -1;  // `if self.global_best.fitness < particle.fitness`
    if (s_best >= 0) gbest_fit[0] = best;
  }
  __syncthreads();
  const int64_t g = s_best;
  if (g >= 0)
    for (int k = threadIdx.x; k < P; k += kPsoBlock) gbest_pos[k] = position[g * P + k];
  for (int64_t b = threadIdx.x; b < B; b += kPsoBlock) {
    const double f = fitness[b];
    if (f > pbest_fit[b]) {  // `if particle.fitness > particle.personal_best.fitness`
      pbest_fit[b] = f;
      for (int k = 0; k < P; ++k) pbest_pos[b * P + k] = position[b * P + k];
    }
  }
}

// k_pso_move: one thread per (particle, dimension); individually rounded operations in NumPy's order.
static __global__ void __launch_bounds__(kBlock)
k_pso_move(double* __restrict__ position, double* __restrict__ velocity, const double* __restrict__ pbest_pos,
           const double* __restrict__ gbest_pos, const double* __restrict__ rnd, int64_t n, int P, double c1,
           double c2) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const double x = position[i];
  const double w = __dadd_rn(0.5, __ddiv_rn(rnd[i], 2.0));
  const double part = __dmul_rn(w, velocity[i]);
  const double cog = __dmul_rn(__dmul_rn(c1, rnd[n + i]), __dadd_rn(pbest_pos[i], -x));
  const double soc = __dmul_rn(__dmul_rn(c2, rnd[2 * n + i]), __dadd_rn(gbest_pos[i % P], -x));
  const double v = __dadd_rn(__dadd_rn(part, cog), soc);
  velocity[i] = v;
  position[i] = __dadd_rn(x, v);
}

// ---- k_fp64_peak ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kBlock) k_fp64_peak(int64_t iters, double* sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int64_t i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c);
    a1 = fma(a1, m, c);
    a2 = fma(a2, m, c);
    a3 = fma(a3, m, c);
    a4 = fma(a4, m, c);
    a5 = fma(a5, m, c);
    a6 = fma(a6, m, c);
    a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 1.2345) sink[0] = s;
}

}  // namespace lsim

/*
 * lsim.h -- C ABI of the B200-native forward image-simulation hot path.
 *
 * The reference (JAXtronomy, /root/reference/jaxtronomy) has no FFI / plugin boundary: its hot path is
 * Python methods traced by XLA.  This header is the boundary a JAXtronomy maintainer would bind instead
 * (ctypes today, an XLA-FFI custom call where jax is installed -- see INTEGRATION.md).  Every entry
 * point names the reference method(s) it replaces.
 *
 * Conventions
 *   - plain C, no torch / CUDA runtime types in signatures: device pointers are `double*`, the stream is
 *     passed as `void*` (a cudaStream_t).
 *   - the caller owns every buffer (inputs, outputs, workspace).  A plan owns device copies of the static
 *     configuration only (profile descriptors, evaluation grid, PSF kernel, data/noise/mask).
 *   - launch calls allocate nothing and never synchronise; all work is enqueued on the given stream.
 *   - every function returns 0 on success or a negative LSIM_E_* code; lsim_strerror() explains it.
 *   - parameters of a batch of B parameter sets ("particles") are a dense row-major [B, P] fp64 matrix,
 *     exactly the `positions` array the reference's batch boundary receives
 *     (Sampling/Samplers/pso.py:51-66, Sampling/sampler.py:215-251).  Each profile parameter is either a
 *     column of that matrix or a per-plan constant (a fixed parameter), mirroring Param.args2kwargs
 *     (Sampling/likelihood.py:305).
 *   - all arithmetic is IEEE fp64.
 */
#ifndef LSIM_H
#define LSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSIM_ABI_VERSION 8
#define LSIM_MAX_PARAMS 12 /* largest param_names list on the path: EPL_MULTIPOLE_M1M3M4 */
#define LSIM_MAX_SLOTS 8   /* output / linear-basis images per parameter set */

/* error codes */
enum {
  LSIM_OK = 0,
  LSIM_E_INVALID = -1,     /* malformed descriptor / argument */
  LSIM_E_UNSUPPORTED = -2, /* valid in the reference but outside this library's scope */
  LSIM_E_CUDA = -3,        /* CUDA runtime error (lsim_last_cuda_error() has the text) */
  LSIM_E_WORKSPACE = -4,   /* workspace too small */
  LSIM_E_NO_IMAGE = -5     /* plan was created without grid / data but an image / logl call needs it */
};

/* Lens profiles: `derivatives` of the reference classes; parameter order = the class's `param_names`. */
enum {
  LSIM_LENS_NULL = 0,                  /* Profiles/null.py:30-38 */
  LSIM_LENS_EPL = 1,                   /* Profiles/epl.py:178-200,366-394  theta_E,gamma,e1,e2,center_x,center_y */
  LSIM_LENS_SIE = 2,                   /* Profiles/sie.py:104-124         theta_E,e1,e2,center_x,center_y; opt[0]=s_scale, opt[1]=gamma; flag NIE_FALSE -> EPL */
  LSIM_LENS_NIE = 3,                   /* Profiles/nie.py:85-107,280-294  theta_E,e1,e2,s_scale,center_x,center_y */
  LSIM_LENS_SIS = 4,                   /* Profiles/sis.py:35-43           theta_E,center_x,center_y */
  LSIM_LENS_SHEAR = 5,                 /* Profiles/shear.py:39-53         gamma1,gamma2,ra_0,dec_0 */
  LSIM_LENS_CONVERGENCE = 6,           /* Profiles/convergence.py:35-46   kappa,ra_0,dec_0 */
  LSIM_LENS_NFW = 7,                   /* Profiles/nfw.py:88-105,259-279,347-365  Rs,alpha_Rs,center_x,center_y */
  LSIM_LENS_TNFW = 8,                  /* Profiles/tnfw.py:69-87,228-349  Rs,alpha_Rs,r_trunc,center_x,center_y */
  LSIM_LENS_MULTIPOLE = 9,             /* Profiles/multipole.py:72-119    m,a_m,phi_m,center_x,center_y,r_E */
  LSIM_LENS_EPL_MULTIPOLE_M1M3M4 = 10, /* Profiles/epl_multipole_m1m3m4.py:463-535,608-678  (12 params, param_names order) */
  LSIM_LENS_CSE = 11,                  /* Profiles/cored_steep_ellipsoid.py:107-132  a,s,e1,e2,center_x,center_y; flag CSE_MAJOR */
  LSIM_LENS_NFW_ELLIPSE_CSE = 12,      /* Profiles/nfw_ellipse_cse.py:250-280  Rs,alpha_Rs,e1,e2,center_x,center_y; flag CSE_LOW_ACCURACY */
  LSIM_LENS_HERNQUIST_ELLIPSE_CSE = 13,/* Profiles/hernquist_ellipse_cse.py:151-164  sigma0,Rs,e1,e2,center_x,center_y */
  LSIM_LENS_TYPE_COUNT = 14
};

/* Light profiles: `function` of the reference classes. */
enum {
  LSIM_LIGHT_SERSIC = 0,         /* LightModel/Profiles/sersic.py:54-76          amp,R_sersic,n_sersic,center_x,center_y */
  LSIM_LIGHT_SERSIC_ELLIPSE = 1, /* LightModel/Profiles/sersic_ellipse.py:61-93  amp,R_sersic,n_sersic,e1,e2,center_x,center_y */
  LSIM_LIGHT_GAUSSIAN = 2,       /* LightModel/Profiles/gaussian.py:27-42        amp,sigma,center_x,center_y */
  LSIM_LIGHT_TYPE_COUNT = 3
};

/* profile option flags (the reference's profile_kwargs, profile_list_base.py:161-232) */
#define LSIM_FLAG_SIE_NIE_FALSE 1u       /* SIE(NIE=False): evaluate through EPL with gamma=opt[1] */
#define LSIM_FLAG_CSE_MAJOR 2u           /* CSE(axis="major") instead of "product_avg" */
#define LSIM_FLAG_CSE_LOW_ACCURACY 4u    /* NFW_ELLIPSE_CSE(high_accuracy=False): 16-term table */

typedef struct lsim_lens_desc {
  int32_t type;                 /* LSIM_LENS_* */
  uint32_t flags;               /* LSIM_FLAG_* */
  int32_t col[LSIM_MAX_PARAMS]; /* column of params[B,P] per parameter (param_names order), or -1: use cst */
  double cst[LSIM_MAX_PARAMS];  /* value of the parameter when col < 0 */
  double opt[2];                /* SIE: s_scale, gamma */
  /* multi-plane only (lsim_plan_desc.multi_plane != 0): deflectors sorted by redshift,
   * constants of MultiPlaneBase.set_T_zs_and_T_ijs (MultiPlane/multi_plane_base.py:89-115) */
  double delta_T;               /* comoving step from the previous deflector (0 on the same plane) */
  double T_z;                   /* transverse comoving distance observer -> this plane */
  double red2phys;              /* reduced -> physical deflection factor D_s / D_ds */
} lsim_lens_desc;

typedef struct lsim_light_desc {
  int32_t type;    /* LSIM_LIGHT_* */
  int32_t lensed;  /* 1: source light, evaluated at the ray-shot position; 0: lens light at the image position */
  int32_t slot;    /* output image this component is accumulated into (0..n_slots-1), -1: skipped (the `k` selection) */
  int32_t use_amp; /* 1: multiply by the `amp` parameter; 0: amp = 1 (LinearBasis.functions_split, linear_basis.py:61-64) */
  int32_t col[LSIM_MAX_PARAMS];
  double cst[LSIM_MAX_PARAMS];
} lsim_light_desc;

enum { LSIM_PRECISION_F64 = 0, LSIM_PRECISION_F32 = 1 };

/* One point-source image (PointSource.point_source_list / linear_response_set, PointSource/point_source.py:255-384;
 * UNLENSED and LENSED_POSITION give positions and amplitudes as parameters).  It is painted with the (supersampled)
 * PSF kernel shifted to its sub-pixel position by bilinear interpolation and re-binned
 * (PointSourceRendering.point_source_rendering, ImSim/Numerics/point_source_rendering.py:33-77;
 * image_util.add_layer2image, Util/image_util.py:9-40), and added to its slot AFTER the PSF convolution of the
 * extended light (ImageModel.point_source, image_model.py:395-429; linear_response_matrix,
 * image_linear_solve.py:361-455). */
typedef struct lsim_point_desc {
  int32_t col_ra, col_dec, col_amp; /* column of params[B,P], or -1: use the constant below */
  int32_t slot;                     /* output image (0..n_slots-1), -1: skipped */
  double cst_ra, cst_dec, cst_amp;
  int32_t use_amp;                  /* 0: amplitude 1 (one linear basis image per point-source image) */
  int32_t error_map;                /* 1: contributes to the PSF error map (psf_error_map_bool_list) */
} lsim_point_desc;

/* PSF convolution on the pixel grid (ImSim/Numerics/convolution.py) */
enum {
  LSIM_CONV_NONE = 0,      /* psf_type NONE or unconvolved=True (numerics.py:170-171) */
  LSIM_CONV_ZERO = 1,      /* PixelKernelConvolution: "same" convolution, zero boundary (convolution.py:73-96) */
  LSIM_CONV_EDGE = 2       /* GaussianConvolution: image edge-replicated by the kernel radius (convolution.py:216-236) */
};

typedef struct lsim_plan_desc {
  int32_t abi_version;      /* LSIM_ABI_VERSION */
  int32_t n_param_cols;     /* P */

  /* lens model: SinglePlane.alpha / ray_shooting (LensModel/single_plane.py:38-56,112-139) or, with
   * multi_plane, MultiPlane.ray_shooting (LensModel/MultiPlane/multi_plane.py:236-275) */
  int32_t n_lens;
  const lsim_lens_desc* lens; /* host array [n_lens] */
  int32_t multi_plane;
  double alpha_scaling;     /* single plane (single_plane.py:139); 1 by default */
  double T_ij_end;          /* multi plane: last deflector -> source */
  double T_z_source;        /* multi plane: observer -> source */

  /* light model: LightModelBase.surface_brightness / LinearBasis.functions_split
   * (LightModel/light_model_base.py:230-252, linear_basis.py:24-91) */
  int32_t n_light;
  const lsim_light_desc* light; /* host array [n_light] */
  int32_t n_slots;          /* 1..LSIM_MAX_SLOTS */

  /* ---- everything below is needed by lsim_image / lsim_logl only (may be zero for ray shooting) ---- */
  /* evaluation grid: RegularGrid (ImSim/Numerics/grid.py:22-64); coordinates in the reference's
   * flattened order (row-major over [ny*ss, nx*ss]) */
  int32_t nx, ny, ss;
  const double* x_sub;      /* host [ny*ss*nx*ss] */
  const double* y_sub;
  const uint8_t* eval_mask; /* host [ny*nx] or NULL: pixels with 0 are not evaluated and get zero flux
                               (flux_evaluate_indexes, grid.py:51-64 / numerics_subframe.py:150-161) */
  /* PSF */
  int32_t conv_mode;        /* LSIM_CONV_* */
  int32_t kernel_size;      /* K (odd) */
  const double* kernel;     /* host [K*K], row-major */
  int32_t kernel_separable; /* 1: kernel == outer(kernel_1d, kernel_1d); lets the library do two 1-D passes */
  const double* kernel_1d;  /* host [K] when kernel_separable */
  double image_scale;       /* pixel_width^2 * flux_scaling (numerics.py:177, image_model.py:368,393) */
  /* supersampling_convolution=True (SubgridKernelConvolution, convolution.py:99-166; GaussianConvolution with
   * supersampling_convolution, convolution.py:238-252): the image is convolved on the supersampled grid with
   * `kernel_hi`, re-binned, and -- when `kernel` (above) is present -- the re-binned unconvolved image convolved
   * with `kernel` on the pixel grid is added.  conv_mode gives the boundary rule of both. */
  int32_t supersampled_conv;   /* 0 / 1 */
  int32_t kernel_hi_size;      /* odd */
  const double* kernel_hi;     /* host [Kh*Kh] */
  int32_t kernel_hi_separable;
  const double* kernel_hi_1d;  /* host [Kh] when separable */

  /* data and noise model: ImageData / ImageNoise (Data/imaging_data.py, Data/image_noise.py) */
  const double* data;       /* host [ny*nx] or NULL */
  const double* cd_data;    /* host [ny*nx]: C_D estimated from the data (+ model error): WLS weights 1/cd_data (image_model.py:594) */
  const double* noise_map2; /* host [ny*nx] noise_map^2, or NULL -> Poisson + background from the model (image_noise.py:99-110) */
  const double* exposure;   /* host [ny*nx] exposure map (already clamped >= 1e-10), used when noise_map2 is NULL */
  double background_rms;
  const uint8_t* mask;      /* host [ny*nx] likelihood mask (1 = use), or NULL = all ones (image_model.py:134-142) */
  double logl_offset;       /* added to every log-likelihood: -chi^2/2 of the likelihood pixels outside the
                               evaluated sub-frame, where the model is identically zero (numerics_subframe.py:134-148) */
  /* arithmetic of the fused image kernel (ray shooting + light + re-bin).  LSIM_PRECISION_F64 (0, default): fp64,
   * the parity-tested path.  LSIM_PRECISION_F32: opt-in single precision for that kernel only (the reference's
   * counterpart is running without jax_enable_x64); convolution and likelihood stay fp64.  Refused with
   * LSIM_E_UNSUPPORTED for models outside its scope (see csrc/lsim_f32.cuh) -- never a silent fp64 fallback. */
  int32_t precision;

  /* point sources (optional; needs the evaluation grid).  ps_kernel: the PSF kernel sampled at ps_supersampling
   * points per pixel (PSF.kernel_point_source_supersampled), or the 3x3 delta of unconvolved=True.
   * pixel = ps_Ma2pix . (ra, dec) + (ps_x_at_radec_0, ps_y_at_radec_0)  (Util/util.py:86-99).  nx == ny required. */
  int32_t n_points;
  const lsim_point_desc* points; /* host array [n_points] */
  int32_t ps_kernel_size;        /* odd */
  int32_t ps_supersampling;      /* >= 1 */
  const double* ps_kernel;       /* host [Kp*Kp] */
  double ps_x_at_radec_0, ps_y_at_radec_0;
  double ps_Ma2pix[4];           /* row-major 2x2 */
  double ps_scale;               /* flux scaling of the point-source image (image_model.py:429) */
  /* PSF error map (ImageModel._error_map_psf, image_model.py:615-648; PointSourceRendering.psf_variance_map,
   * point_source_rendering.py:79-111; kernel_util.estimate_amp, Util/kernel_util.py:7-47): for every point-source
   * image flagged error_map, sigma^2 += psf_variance_map * amp_est^2 painted at its pixel position, amp_est =
   * max(sum of the 5x5 data pixels around it, 0) / ps_mean_kernel (0 within 2 pixels of the frame edge).  The map is
   * added to C_D in the WLS weights and, as |.|, to C_D(model) in chi^2.  NULL: no error map. */
  const double* ps_variance_map; /* host [Kv*Kv], pixel resolution */
  int32_t ps_variance_size;      /* Kv, odd */
  double ps_mean_kernel;         /* sum of the central 5x5 of the pixel-resolution PSF kernel */
} lsim_plan_desc;

typedef struct lsim_plan lsim_plan;

/* Validate a descriptor, upload static data to the current CUDA device and build a plan.
 * PSF kernels with a side of 27 pixels or more (zero boundary, not separable) are convolved by FFT (cuFFT D2Z/Z2D on
 * zero-padded frames, loaded with dlopen on first use; direct convolution is used if cuFFT is absent).  Such a plan
 * owns the FFT scratch, so only ONE image / likelihood call may be in flight per plan (use one plan per stream);
 * all other plans are re-entrant across streams. */
int lsim_plan_create(const lsim_plan_desc* desc, lsim_plan** out);
void lsim_plan_destroy(lsim_plan* plan);

/* LensModel.ray_shooting (LensModel/lens_model.py:236-252) / LensModelBulk.ray_shooting
 * (LensModel/lens_model_bulk.py:113-128), batched over B parameter sets.
 *   x, y       device [N] (shared by all sets) or [B,N] when coords_batched != 0
 *   out_x/y    device [B,N]
 *   what       0: source position beta = theta - alpha (ray_shooting); 1: deflection alpha (SinglePlane.alpha) */
int lsim_ray_shoot(const lsim_plan* plan, const double* params, int64_t B, const double* x, const double* y,
                   int64_t N, int coords_batched, double* out_x, double* out_y, int what, void* stream);

/* LensModel.hessian with diff=None (LensModel/lens_model.py:384-414 -> SinglePlane.hessian,
 * LensModel/single_plane.py:142-180), batched over B parameter sets; single plane, fp64 plans only
 * (LSIM_E_UNSUPPORTED otherwise -- the reference's multi-plane Hessian is a finite difference of ray_shooting,
 * which the caller forms from lsim_ray_shoot).
 *   x, y       device [N] (shared by all sets) or [B,N] when coords_batched != 0
 *   out        device [B,4,N]: f_xx, f_xy, f_yx, f_yy (times the plan's alpha_scaling) */
int lsim_hessian(const lsim_plan* plan, const double* params, int64_t B, const double* x, const double* y, int64_t N,
                 int coords_batched, double* out, void* stream);

/* PositionLikelihood.source_position_likelihood for ONE LENSED_POSITION point-source model
 * (Sampling/Likelihoods/position_likelihood.py:214-283, penalty :296-352; source position = mean of the ray-shot image
 * positions, PointSource/Types/lensed_position.py:65-90), batched over B parameter sets; single plane, fp64 plans.
 *   x, y            device [n_img] (shared) or [B,n_img] when coords_batched != 0: the image positions
 *   sigma           image-plane position uncertainty (> 0);  hard_bound_rms < 0: no hard bound
 *   logl            device [B]: minus the sum of the per-image penalties (overwritten) */
int lsim_source_position_logl(const lsim_plan* plan, const double* params, int64_t B, const double* x, const double* y,
                              int n_img, int coords_batched, double sigma, double hard_bound_rms, double* logl,
                              void* stream);

/* LightModel.surface_brightness / functions_split on arbitrary points (no lensing applied here).
 *   out        device [B, n_slots, N] */
int lsim_surface_brightness(const lsim_plan* plan, const double* params, int64_t B, const double* x,
                            const double* y, int64_t N, int coords_batched, double* out, void* stream);

/* Bytes of caller-provided scratch lsim_image / lsim_logl want for B sets (they accept less, down to
 * lsim_workspace_bytes(plan, 1), and then process the batch in more chunks). */
size_t lsim_workspace_bytes(const lsim_plan* plan, int64_t B);

/* ImageModel.image / source_surface_brightness / lens_surface_brightness and the rows of
 * ImageLinearFit.linear_response_matrix (ImSim/image_model.py:223-393,431-492,
 * ImSim/image_linear_solve.py:361-455): ray-shoot + light on the supersampled grid, re-bin, convolve, scale.
 *   images     device [B, n_slots, ny, nx] */
int lsim_image(const lsim_plan* plan, const double* params, int64_t B, double* images, void* workspace,
               size_t workspace_bytes, void* stream);

/* Numerics.re_size_convolve (ImSim/Numerics/numerics.py:157-177) for a caller-supplied flux on the supersampled
 * evaluation grid: re-bin (ss x ss mean), convolve, scale by pixel_width^2.
 *   flux       device [n_images, ny*ss * nx*ss] in the reference's flattened order
 *   images     device [n_images, ny, nx]
 *   unconvolved != 0 skips the PSF (numerics.py:170-171) */
int lsim_resize_convolve(const lsim_plan* plan, const double* flux, int64_t n_images, int unconvolved,
                         double* images, void* workspace, size_t workspace_bytes, void* stream);

/* ImageModel.likelihood_data_given_model (image_model.py:162-221) when linear_solve == 0, or
 * ImageLinearFit.likelihood_data_given_model (image_linear_solve.py:228-339, de_lens.py:10-49) when
 * linear_solve != 0 (then the slots are the linear basis and amplitudes are solved by WLS).
 *   logl       device [B]
 *   amps       device [B, n_slots] or NULL (linear_solve only)
 *   model      device [B, ny, nx] or NULL: the model image the likelihood was evaluated on
 *   source_marg != 0 adds the marginalisation constant (de_lens.py:39-49) */
int lsim_logl(const lsim_plan* plan, const double* params, int64_t B, int linear_solve, int source_marg,
              double* logl, double* amps, double* model, void* workspace, size_t workspace_bytes, void* stream);

/* Log-likelihood AND its gradient with respect to the parameter matrix: what
 * jax.value_and_grad(LikelihoodModule.logL) gives the reference's gradient-descent sampler
 * (Sampling/Samplers/optax.py:39,91-96 -> Sampling/likelihood.py:293-372 -> ImageLinearFit / ImageModel
 * likelihood_data_given_model).  Forward pass as lsim_logl; then the adjoint of the likelihood (the WLS amplitude
 * solve differentiated analytically), the transposed PSF convolution, and one dual-number pass over the supersampled
 * grid in which every profile carries the derivatives with respect to its own parameters.  Clamps differentiate
 * like jnp.maximum / jnp.where; the EPL hypergeometric function takes its value from the nmax schedule and its
 * tangent from the fixed 100-term series in (t, z), as the reference's custom_jvp does (Profiles/epl.py:321-340).
 *   logl       device [B]
 *   grad       device [B, P]: d logl[b] / d params[b, :]; columns no profile reads are 0; with linear_solve the
 *              amplitudes are functions of the other parameters (solved per set), not parameters
 *   amps       device [B, n_slots] or NULL (linear_solve only)
 * Scope: single plane and multi plane (reverse sweep over the planes, at most 8 profiles after expanding
 * EPL_MULTIPOLE_M1M3M4; MultiPlaneBase.ray_shooting_partial_comoving, multi_plane_base.py:159-233); every lens profile
 * of this header; SERSIC, SERSIC_ELLIPSE, GAUSSIAN; point-source images (ra, dec and, without linear_solve, amp: the
 * slope of the bilinear PSF stamp, point_source_rendering.py:33-77); PSF none, pixel kernel (zero boundary) or separable
 * Gaussian (edge-replicated); regular convolution; fp64.  Anything else (supersampled convolution, PSF error map,
 * fp32): LSIM_E_UNSUPPORTED (never a finite-difference or CPU fallback).  Results are reproducible run to run (fixed
 * reduction order; the point-source entries are added by one CTA per (set, column)).  Workspace: lsim_logl_grad_workspace_bytes (less is accepted,
 * down to one set). */
size_t lsim_logl_grad_workspace_bytes(const lsim_plan* plan, int64_t B);
int lsim_logl_grad(const lsim_plan* plan, const double* params, int64_t B, int linear_solve, double* logl,
                   double* grad, double* amps, void* workspace, size_t workspace_bytes, void* stream);

/* One particle-swarm iteration AFTER the fitness evaluation, entirely on the device: the loop body
 * of lenstronomy's ParticleSwarmOptimizer._sample that jaxtronomy/Sampling/Samplers/pso.py:11-66 inherits, i.e.
 *   global best   <- the particle with the largest fitness if it beats the current one (first wins on ties)
 *   personal best <- position / fitness where fitness improved
 *   velocity      <- w v + c1 r1 (pbest - x) + c2 r2 (gbest - x),  w = 0.5 + r0 / 2      (per particle AND dimension)
 *   position      <- x + velocity
 * evaluated in exactly that operation order with individually rounded operations (bit-exact against NumPy).
 *   position, velocity, pbest_pos   device [B, P] (updated in place)
 *   pbest_fit  device [B]; gbest_pos device [P]; gbest_fit device [1]  (updated in place)
 *   fitness    device [B]: log-likelihood of `position` (from lsim_logl + the sampler's bound penalties)
 *   rnd        device [3, B, P]: uniform [0,1) numbers r0, r1, r2 (the caller's generator, as the reference's PSO
 *              draws them from NumPy's global generator)
 * No allocation, no synchronisation; two launches on `stream`. */
int lsim_pso_step(double* position, double* velocity, double* pbest_pos, double* pbest_fit, double* gbest_pos,
                  double* gbest_fit, const double* fitness, const double* rnd, int64_t B, int32_t P, double c1,
                  double c2, void* stream);

/* ImageModel._error_map_psf (image_model.py:615-648): the PSF error map (sigma^2) of each parameter set.
 *   error_map  device [B, ny, nx]; zeros when the plan has no ps_variance_map */
int lsim_psf_error_map(const lsim_plan* plan, const double* params, int64_t B, double* error_map, void* stream);

/* introspection */
int lsim_abi_version(void);
const char* lsim_strerror(int code);
const char* lsim_last_cuda_error(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t lsim_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: while enabled, every launch of this library is bracketed by
 * CUDA events on the stream it is launched on.  lsim_timing_read synchronises on those events, returns the summed
 * milliseconds and launch counts per kernel class (0 = k_image, 1 = k_conv_*, 2 = k_likelihood, 3 = other) and
 * clears the record.  Not for use inside a timed region (the events cost a few microseconds per launch). */
int lsim_timing_enable(int on);
int lsim_timing_read(double ms[4], int64_t count[4]);

/* fp64 FMA-pipe microbenchmark used for the roofline denominator: runs `iters` dependent DFMA chains
 * on every SM, returns achieved FLOP/s (2 flop per FMA) through *flops, timed with CUDA events. */
int lsim_measure_fp64_peak(int64_t iters, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSIM_H */

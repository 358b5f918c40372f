// lsim_api.cu -- C ABI (include/lsim.h) over the kernels in lsim_kernels.cuh.  Single translation unit.
#include <cufft.h>  // types only: the library is dlopen()ed on first use, liblsim.so does not link it
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/lsim.h"
#include "lsim_kernels.cuh"
#include "lsim_f32.cuh"
#include "lsim_grad_kernels.cuh"

// the headline kernels are compiled in lsim_hot.cu
#define LSIM_HOT extern template
#include "lsim_hot_list.h"
#undef LSIM_HOT

using namespace lsim;

namespace {

std::atomic<int64_t> g_launches{0};
thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return LSIM_E_CUDA;
}

#define LSIM_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

int after_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, name);
  return LSIM_OK;
}

// ---- optional per-kernel timing (lsim_timing_enable): CUDA events around every launch, on its stream ----------
struct TimedEvent {
  int cls;
  cudaEvent_t e0, e1;
};
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
std::vector<TimedEvent> g_timed;

struct Timed {
  bool on;
  TimedEvent ev;
  cudaStream_t st;
  Timed(int cls, cudaStream_t stream) : on(g_timing.load() != 0), st(stream) {
    if (!on) return;
    ev.cls = cls;
    cudaEventCreate(&ev.e0);
    cudaEventCreate(&ev.e1);
    cudaEventRecord(ev.e0, st);
  }
  ~Timed() {
    if (!on) return;
    cudaEventRecord(ev.e1, st);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timed.push_back(ev);
  }
};
enum { kClsImage = 0, kClsConv = 1, kClsLike = 2, kClsOther = 3 };

const int kLensParams[LSIM_LENS_TYPE_COUNT] = {0, 6, 5, 6, 3, 4, 3, 4, 5, 6, 12, 6, 6, 6};
const int kLightParams[LSIM_LIGHT_TYPE_COUNT] = {5, 7, 4};

}  // namespace

// ---- FFT convolution for large PSF kernels (zero boundary): cuFFT D2Z / Z2D on zero-padded images -------------------
struct CufftApi {
  cufftResult (*PlanMany)(cufftHandle*, int, int*, int*, int, int, int*, int, int, cufftType, int) = nullptr;
  cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal*, cufftDoubleComplex*) = nullptr;
  cufftResult (*ExecZ2D)(cufftHandle, cufftDoubleComplex*, cufftDoubleReal*) = nullptr;
  cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
  cufftResult (*Destroy)(cufftHandle) = nullptr;
  bool ok = false;
};

static const CufftApi& cufft_api() {
  static CufftApi api = [] {
    CufftApi a;
    void* h = nullptr;
    for (const char* name : {"libcufft.so.11", "libcufft.so.12", "libcufft.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return a;
    a.PlanMany = reinterpret_cast<decltype(a.PlanMany)>(dlsym(h, "cufftPlanMany"));
    a.ExecD2Z = reinterpret_cast<decltype(a.ExecD2Z)>(dlsym(h, "cufftExecD2Z"));
    a.ExecZ2D = reinterpret_cast<decltype(a.ExecZ2D)>(dlsym(h, "cufftExecZ2D"));
    a.SetStream = reinterpret_cast<decltype(a.SetStream)>(dlsym(h, "cufftSetStream"));
    a.Destroy = reinterpret_cast<decltype(a.Destroy)>(dlsym(h, "cufftDestroy"));
    a.ok = a.PlanMany && a.ExecD2Z && a.ExecZ2D && a.SetStream && a.Destroy;
    return a;
  }();
  return api;
}

struct FftConv {
  int nx = 0, ny = 0, K = 0, PX = 0, PY = 0, nb = 0;
  cufftHandle fwd = 0, inv = 0;
  bool have_fwd = false, have_inv = false;
  double* pad = nullptr;               // [nb][PY][PX]
  cufftDoubleComplex* spec = nullptr;  // [nb][PY][PX/2+1]
  cufftDoubleComplex* kspec = nullptr; // [PY][PX/2+1]
  ~FftConv() {
    if (have_fwd) cufft_api().Destroy(fwd);
    if (have_inv) cufft_api().Destroy(inv);
    cudaFree(pad);
    cudaFree(spec);
    cudaFree(kspec);
  }
};

struct lsim_plan {
  PlanDev dev;
  std::vector<double> kernel_host, kernel_hi_host;  // host copies: the tiled convolution takes its weights by value
  FftConv* fft_lo = nullptr;  // pixel-grid kernel, when it takes the FFT path
  FftConv* fft_hi = nullptr;  // supersampled kernel, likewise
  std::vector<void*> owned;
  bool full = false;
  int n_series = 0;
  int sig = -1;  // index of the specialised image kernel for this model, or -1
  bool fold = false;  // the model and the grid are inside k_image_fold's scope (lsim_kernels.cuh)
  bool ray_epl_shear = false;  // single-plane EPL + SHEAR: k_ray_shoot with the compile-time profile list
  bool has_image = false, has_data = false;
  bool f32 = false;  // LSIM_PRECISION_F32: the image kernel runs in single precision
  size_t image_smem = 0;
  int n_sm = 148;
  // gradient path (lsim_logl_grad): transposed PSF = the flipped kernel; grad_nl = lens profiles held in registers
  bool grad_ok = false;
  int grad_tier = 2;  // code tier of k_image_grad this model needs (lsim_grad.cuh: lens_grad_tier)
  int grad_series = 0;
  int grad_stateful = 0;  // leading profiles of the gradient's lens list with a register-resident point state
  bool grad_mp = false;   // multi-plane: k_image_grad<0, 2, true> (reverse sweep over the planes)
  size_t grad_smem = 0;
  const lsim_lens_desc* grad_lens = nullptr;  // device: the gradient's lens list (EPL_MULTIPOLE_M1M3M4 expanded)
  int n_grad_lens = 0;
  const lsim_lens_desc* hess_lens = nullptr;  // device: the Hessian's lens list (same expansion; any single-plane fp64 plan)
  int n_hess_lens = 0;
  int hess_series = 0;
  size_t hess_smem = 0;
  bool hess_ok = false;
  const double* kernel_flip = nullptr;
  const double* kernel_1d_flip = nullptr;
  std::vector<double> kernel_flip_host;

  ~lsim_plan() {
    for (void* p : owned) cudaFree(p);
    delete fft_lo;
    delete fft_hi;
  }
  template <class T>
  int upload(const T* host, size_t n, const T** dev_out) {
    void* d = nullptr;
    LSIM_CUDA(cudaMalloc(&d, n * sizeof(T)));
    owned.push_back(d);
    LSIM_CUDA(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = static_cast<const T*>(d);
    return LSIM_OK;
  }
};

namespace {

int upload_tables() {
  // constant memory is per device (per context): upload on every plan creation; it is 1.4 KB
  LSIM_CUDA(cudaMemcpyToSymbol(d_nfw_cse_high_s, LSIM_NFW_CSE_HIGH_S, sizeof(LSIM_NFW_CSE_HIGH_S)));
  LSIM_CUDA(cudaMemcpyToSymbol(d_nfw_cse_high_a, LSIM_NFW_CSE_HIGH_A, sizeof(LSIM_NFW_CSE_HIGH_A)));
  LSIM_CUDA(cudaMemcpyToSymbol(d_nfw_cse_low_s, LSIM_NFW_CSE_LOW_S, sizeof(LSIM_NFW_CSE_LOW_S)));
  LSIM_CUDA(cudaMemcpyToSymbol(d_nfw_cse_low_a, LSIM_NFW_CSE_LOW_A, sizeof(LSIM_NFW_CSE_LOW_A)));
  LSIM_CUDA(cudaMemcpyToSymbol(d_hernquist_cse_s, LSIM_HERNQUIST_CSE_S, sizeof(LSIM_HERNQUIST_CSE_S)));
  LSIM_CUDA(cudaMemcpyToSymbol(d_hernquist_cse_a, LSIM_HERNQUIST_CSE_A, sizeof(LSIM_HERNQUIST_CSE_A)));
  return LSIM_OK;
}

// Opt a kernel in to its dynamic shared memory.  Every kernel that evaluates profiles carries the 32 KB log2 / exp2
// tables as STATIC shared memory, and static + dynamic above 48 KB needs the opt-in: done once per kernel and size.
template <class K>
int ensure_dyn_smem(K kernel, size_t bytes) {
  static std::mutex m;
  static std::unordered_map<const void*, size_t> have[16];  // per device of this process (the attribute is per device)
  if (bytes <= 8 * 1024) return LSIM_OK;  // no kernel has more than 40 KB of static shared memory
  int dev = 0;
  LSIM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(m);
  size_t& h = have[dev & 15][reinterpret_cast<const void*>(kernel)];
  if (h >= bytes) return LSIM_OK;
  LSIM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  h = bytes;
  return LSIM_OK;
}
#define LSIM_SMEM(kernel, bytes)                          \
  do {                                                    \
    const int rc_smem_ = ensure_dyn_smem((kernel), (bytes)); \
    if (rc_smem_ != LSIM_OK) return rc_smem_;             \
  } while (0)

constexpr size_t kConvSmemMax = 200 * 1024;
constexpr size_t kGradSmemMax = 188 * 1024;  // + 34 KB static (math tables) <= 227 KB

int configure_kernels() {
  LSIM_CUDA(cudaFuncSetAttribute(k_conv_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmemMax));
  LSIM_CUDA(cudaFuncSetAttribute(k_conv_separable, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmemMax));
#define LSIM_GRAD_ATTR(NS)                                                                                        \
  LSIM_CUDA(cudaFuncSetAttribute(k_image_grad<NS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemMax)); \
  LSIM_CUDA(cudaFuncSetAttribute(k_image_grad<NS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemMax)); \
  LSIM_CUDA(cudaFuncSetAttribute(k_image_grad<NS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemMax));
  LSIM_CUDA(cudaFuncSetAttribute(k_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemMax));
  LSIM_CUDA(cudaFuncSetAttribute(k_source_position, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemMax));
  LSIM_GRAD_ATTR(0)
  LSIM_GRAD_ATTR(1)
  LSIM_GRAD_ATTR(2)
#undef LSIM_GRAD_ATTR
  LSIM_CUDA(cudaFuncSetAttribute(k_image_grad<0, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmemMax));
  return LSIM_OK;
}

int find_signature(const lsim_plan_desc* d);  // defined with the kernel-variant tables below
int effective_lens_type(const lsim_lens_desc& d);
int effective_light_type(const lsim_light_desc& d);
int fft_conv_create(int nx, int ny, int K, const double* kernel, int mode, FftConv** out);  // defined with the convolutions

double series_tol_default() {
  const char* s = getenv("LSIM_SERIES_TOL");
  if (s && *s) return atof(s);
  return 1e-14;
}

}  // namespace

extern "C" {

int lsim_abi_version(void) { return LSIM_ABI_VERSION; }

const char* lsim_strerror(int code) {
  switch (code) {
    case LSIM_OK: return "ok";
    case LSIM_E_INVALID: return "invalid descriptor or argument";
    case LSIM_E_UNSUPPORTED: return "configuration outside the scope of this library";
    case LSIM_E_CUDA: return "CUDA runtime error (see lsim_last_cuda_error)";
    case LSIM_E_WORKSPACE: return "workspace too small (see lsim_workspace_bytes)";
    case LSIM_E_NO_IMAGE: return "plan has no evaluation grid / data for this call";
    default: return "unknown error";
  }
}

const char* lsim_last_cuda_error(void) { return g_cuda_err; }

int64_t lsim_launch_count(void) { return g_launches.load(); }

int lsim_plan_create(const lsim_plan_desc* d, lsim_plan** out) {
  if (!d || !out) return LSIM_E_INVALID;
  *out = nullptr;
  if (d->abi_version != LSIM_ABI_VERSION) return LSIM_E_INVALID;
  if (d->n_param_cols < 0 || d->n_lens < 0 || d->n_light < 0) return LSIM_E_INVALID;
  if (d->n_lens > 0 && !d->lens) return LSIM_E_INVALID;
  if (d->n_light > 0 && !d->light) return LSIM_E_INVALID;
  if (d->n_light > kMaxLight) return LSIM_E_UNSUPPORTED;
  if (d->n_slots < 1 || d->n_slots > LSIM_MAX_SLOTS) return LSIM_E_INVALID;
  bool full = false;
  int n_series = 0;
  for (int i = 0; i < d->n_lens; ++i) {
    const lsim_lens_desc& L = d->lens[i];
    if (L.type < 0 || L.type >= LSIM_LENS_TYPE_COUNT) return LSIM_E_INVALID;
    for (int k = 0; k < kLensParams[L.type]; ++k)
      if (L.col[k] >= d->n_param_cols) return LSIM_E_INVALID;
    full = full || lens_type_is_full(L.type);
    if (lens_type_has_series(L)) ++n_series;
    if (d->multi_plane && !(L.T_z > 0)) return LSIM_E_INVALID;
  }
  // every window of kLensChunk consecutive profiles may hold at most kMaxSeriesProfiles series-bearing ones
  for (int c0 = 0; c0 < d->n_lens; c0 += kLensChunk) {
    int cnt = 0;
    for (int i = c0; i < d->n_lens && i < c0 + kLensChunk; ++i) cnt += lens_type_has_series(d->lens[i]) ? 1 : 0;
    if (cnt > kMaxSeriesProfiles) return LSIM_E_UNSUPPORTED;
  }
  for (int i = 0; i < d->n_light; ++i) {
    const lsim_light_desc& G = d->light[i];
    if (G.type < 0 || G.type >= LSIM_LIGHT_TYPE_COUNT) return LSIM_E_INVALID;
    if (G.slot >= d->n_slots) return LSIM_E_INVALID;
    for (int k = 0; k < kLightParams[G.type]; ++k)
      if (G.col[k] >= d->n_param_cols) return LSIM_E_INVALID;
  }
  if (d->multi_plane && !(d->T_z_source > 0)) return LSIM_E_INVALID;
  const bool has_image = d->x_sub != nullptr;
  if (has_image) {
    if (d->nx <= 0 || d->ny <= 0 || d->ss <= 0 || !d->y_sub) return LSIM_E_INVALID;
    if (d->n_lens > kMaxImageLens) return LSIM_E_UNSUPPORTED;
    if (d->conv_mode < LSIM_CONV_NONE || d->conv_mode > LSIM_CONV_EDGE) return LSIM_E_INVALID;
    if (d->conv_mode != LSIM_CONV_NONE) {
      if (d->supersampled_conv) {
        if (d->kernel_hi_size < 1 || d->kernel_hi_size % 2 == 0 || !d->kernel_hi) return LSIM_E_INVALID;
        if (d->kernel_hi_separable && !d->kernel_hi_1d) return LSIM_E_INVALID;
        if (d->kernel && (d->kernel_size < 1 || d->kernel_size % 2 == 0)) return LSIM_E_INVALID;
      } else {
        if (d->kernel_size < 1 || d->kernel_size % 2 == 0 || !d->kernel) return LSIM_E_INVALID;
      }
      if (d->kernel_separable && !d->kernel_1d) return LSIM_E_INVALID;
    }
  }
  const bool has_data = has_image && d->data && d->cd_data && (d->noise_map2 || d->exposure);
  if (d->precision != LSIM_PRECISION_F64 && d->precision != LSIM_PRECISION_F32) return LSIM_E_INVALID;
  if (d->n_points < 0 || (d->n_points > 0 && (!d->points || !has_image))) return LSIM_E_INVALID;
  if (d->n_points > 0) {
    if (d->ps_kernel_size < 1 || d->ps_kernel_size % 2 == 0 || !d->ps_kernel || d->ps_supersampling < 1)
      return LSIM_E_INVALID;
    if (d->nx != d->ny) return LSIM_E_UNSUPPORTED;  // the reference's rendering assumes a square frame
    if (d->eval_mask) return LSIM_E_UNSUPPORTED;    // sub-frames: the host side cuts the frame, not a pixel list
    if (d->ps_variance_map && (d->ps_variance_size < 1 || d->ps_variance_size % 2 == 0 || !(d->ps_mean_kernel > 0)))
      return LSIM_E_INVALID;
    for (int i = 0; i < d->n_points; ++i) {
      const lsim_point_desc& q = d->points[i];
      if (q.slot >= d->n_slots || q.col_ra >= d->n_param_cols || q.col_dec >= d->n_param_cols ||
          q.col_amp >= d->n_param_cols)
        return LSIM_E_INVALID;
    }
  }
  if (d->precision == LSIM_PRECISION_F32 && !f32_supported(d)) return LSIM_E_UNSUPPORTED;

  lsim_plan* p = new (std::nothrow) lsim_plan();
  if (!p) return LSIM_E_INVALID;
  int rc;
#define TRY(x)          \
  do {                  \
    rc = (x);           \
    if (rc != LSIM_OK) { \
      delete p;         \
      return rc;        \
    }                   \
  } while (0)
  TRY(upload_tables());
  TRY(configure_kernels());
  int devid = 0;
  cudaGetDevice(&devid);
  cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, devid);
  PlanDev& D = p->dev;
  memset(&D, 0, sizeof(D));
  D.n_param_cols = d->n_param_cols;
  D.n_lens = d->n_lens;
  D.ray.multi_plane = d->multi_plane;
  D.ray.alpha_scaling = d->multi_plane ? 1.0 : d->alpha_scaling;
  D.ray.T_ij_end = d->T_ij_end;
  D.ray.inv_T_z_source = d->multi_plane ? 1.0 / d->T_z_source : 0.0;
  D.ray.mk = math_k();
  D.n_light = d->n_light;
  D.n_slots = d->n_slots;
  D.series_tol = series_tol_default();
  if (d->n_lens) TRY(p->upload(d->lens, (size_t)d->n_lens, &D.lens));
  if (d->n_light) TRY(p->upload(d->light, (size_t)d->n_light, &D.light));
  p->full = full;
  p->n_series = n_series;
  p->sig = find_signature(d);
  p->ray_epl_shear = !d->multi_plane && d->n_lens == 2 && effective_lens_type(d->lens[0]) == LSIM_LENS_EPL &&
                     effective_lens_type(d->lens[1]) == LSIM_LENS_SHEAR && !getenv("LSIM_NO_SPECIALISE");
  p->has_image = has_image;
  p->has_data = has_data;
  p->f32 = d->precision == LSIM_PRECISION_F32;
  {  // Hessian path (lsim_hessian): single plane, fp64, every profile within the dual-number scope
    std::vector<lsim_lens_desc> gl((size_t)4 * (d->n_lens > 0 ? d->n_lens : 1));
    const int ngl = ad::expand_grad_lens(d->lens, d->n_lens, gl.data(), (int)gl.size());
    bool hok = !d->multi_plane && !p->f32 && ngl >= 0;
    for (int i = 0; hok && i < d->n_lens; ++i) hok = ad::lens_grad_supported(d->lens[i]);
    int hs = 0;
    for (int i = 0; hok && i < ngl; ++i) hs += ad::lens_grad_has_series(gl[i]) ? 1 : 0;
    p->hess_series = hs;
    p->n_hess_lens = hok ? ngl : 0;
    p->hess_smem = hess_smem_bytes(p->n_hess_lens, hs);
    p->hess_ok = hok && p->hess_smem <= kGradSmemMax;
    if (p->hess_ok && ngl > 0) TRY(p->upload(gl.data(), (size_t)ngl, &p->hess_lens));
  }
  if (has_image) {
    D.nx = d->nx;
    D.ny = d->ny;
    D.ss = d->ss;
    const size_t nsub = (size_t)d->nx * d->ny * d->ss * d->ss, npix = (size_t)d->nx * d->ny;
    TRY(p->upload(d->x_sub, nsub, &D.x_sub));
    TRY(p->upload(d->y_sub, nsub, &D.y_sub));
    {  // are the coordinates affine in the (row, column) indices of the supersampled grid?  (every RegularGrid is)
      const int64_t W = (int64_t)d->nx * d->ss, H = (int64_t)d->ny * d->ss;
      const double* c2[2] = {d->x_sub, d->y_sub};
      double g[2][3];
      bool ok = true;
      for (int k = 0; k < 2 && ok; ++k) {
        const double* v = c2[k];
        g[k][0] = v[0];
        g[k][1] = H > 1 ? (v[(H - 1) * W] - v[0]) / (double)(H - 1) : 0.0;
        g[k][2] = W > 1 ? (v[W - 1] - v[0]) / (double)(W - 1) : 0.0;
        double vmax = 0.0;
        for (int64_t i = 0; i < H * W; ++i) vmax = fabs(v[i]) > vmax ? fabs(v[i]) : vmax;
        const double tol = 16.0 * DBL_EPSILON * (vmax > 0 ? vmax : 1.0);
        for (int64_t r = 0; r < H && ok; ++r)
          for (int64_t c = 0; c < W; ++c)
            if (!(fabs(fma((double)r, g[k][1], fma((double)c, g[k][2], g[k][0])) - v[r * W + c]) <= tol)) {
              ok = false;
              break;
            }
      }
      D.grid_affine = ok ? 1 : 0;
      D.gx0 = g[0][0];
      D.gxi = g[0][1];
      D.gxj = g[0][2];
      D.gy0 = g[1][0];
      D.gyi = g[1][1];
      D.gyj = g[1][2];
    }
    D.n_eval = (int32_t)npix;
    if (d->eval_mask) {
      std::vector<int32_t> list;
      for (size_t i = 0; i < npix; ++i)
        if (d->eval_mask[i]) list.push_back((int32_t)i);
      if (list.size() < npix) {
        D.n_eval = (int32_t)list.size();
        if (!list.empty()) TRY(p->upload(list.data(), list.size(), &D.eval_list));
      }
    }
    D.conv_mode = d->conv_mode;
    D.K = (d->conv_mode != LSIM_CONV_NONE && d->kernel) ? d->kernel_size : 0;
    D.image_scale = d->image_scale;
    {  // k_image_fold: one EPL + SHEAR / CONVERGENCE, one or two Sersic components, affine grid, no sub-frame
      int n_epl = 0;
      bool ok = !d->multi_plane && !p->f32 && D.grid_affine && !D.eval_list && D.n_eval == (int32_t)npix &&
                !d->supersampled_conv && d->ss >= 1 && d->ss <= 3 && d->n_light >= 1 && d->n_light <= 2 &&
                d->n_slots <= d->n_light && !getenv("LSIM_NO_FOLD");
      for (int i = 0; ok && i < d->n_lens; ++i) {
        const int t = effective_lens_type(d->lens[i]);
        if (t == LSIM_LENS_EPL) ++n_epl;
        else if (t != LSIM_LENS_SHEAR && t != LSIM_LENS_CONVERGENCE) ok = false;
      }
      for (int i = 0; ok && i < d->n_light; ++i) ok = effective_light_type(d->light[i]) == LSIM_LIGHT_SERSIC_ELLIPSE;
      p->fold = ok && n_epl == 1;
    }
    if (d->conv_mode != LSIM_CONV_NONE) {
      if (d->kernel) {
        TRY(p->upload(d->kernel, (size_t)d->kernel_size * d->kernel_size, &D.kernel));
        if (d->kernel_separable) TRY(p->upload(d->kernel_1d, (size_t)d->kernel_size, &D.kernel_1d));
      }
      if (d->supersampled_conv) {
        D.ss_conv = 1;
        D.K_hi = d->kernel_hi_size;
        TRY(p->upload(d->kernel_hi, (size_t)d->kernel_hi_size * d->kernel_hi_size, &D.kernel_hi));
        if (d->kernel_hi_separable) TRY(p->upload(d->kernel_hi_1d, (size_t)d->kernel_hi_size, &D.kernel_hi_1d));
      }
    }
    if (d->n_points > 0) {
      D.n_points = d->n_points;
      D.ps_K = d->ps_kernel_size;
      D.ps_ss = d->ps_supersampling;
      TRY(p->upload(d->points, (size_t)d->n_points, &D.points));
      TRY(p->upload(d->ps_kernel, (size_t)d->ps_kernel_size * d->ps_kernel_size, &D.ps_kernel));
      D.ps_x0 = d->ps_x_at_radec_0;
      D.ps_y0 = d->ps_y_at_radec_0;
      for (int k = 0; k < 4; ++k) D.ps_M[k] = d->ps_Ma2pix[k];
      D.ps_scale = d->ps_scale;
      if (d->ps_variance_map && has_data) {
        bool any = false;
        for (int i = 0; i < d->n_points; ++i) any = any || d->points[i].error_map;
        if (any) {
          TRY(p->upload(d->ps_variance_map, (size_t)d->ps_variance_size * d->ps_variance_size, &D.ps_var));
          D.ps_var_K = d->ps_variance_size;
          D.ps_mean_kernel = d->ps_mean_kernel;
        }
      }
    }
    if (D.K > 0) p->kernel_host.assign(d->kernel, d->kernel + (size_t)d->kernel_size * d->kernel_size);
    {  // gradient path: scope check, transposed PSF
      // point sources: positions and amplitudes are differentiated (k_point_sources_grad); the PSF error map is not
      bool ok = !d->supersampled_conv && !(d->n_points > 0 && D.ps_var) && !p->f32 &&
                (d->conv_mode != LSIM_CONV_EDGE || d->kernel_separable);
      int gs = 0;
      int tier = 0;
      std::vector<lsim_lens_desc> gl((size_t)4 * (d->n_lens > 0 ? d->n_lens : 1));
      // multi-plane: the list stays in plane order and runs the dual-number tier with the reverse sweep over the planes
      const int ngl = ad::expand_grad_lens(d->lens, d->n_lens, gl.data(), (int)gl.size(), d->multi_plane != 0);
      ok = ok && ngl >= 0 && (!d->multi_plane || ngl <= ad::kMaxGradPlanes);
      p->grad_mp = d->multi_plane != 0;
      for (int i = 0; ok && i < d->n_lens; ++i) ok = ad::lens_grad_supported(d->lens[i]);
      for (int i = 0; ok && i < ngl; ++i) {
        tier = ad::lens_grad_tier(gl[i]) > tier ? ad::lens_grad_tier(gl[i]) : tier;
        gs += ad::lens_grad_has_series(gl[i]) ? 1 : 0;
        p->grad_stateful += ad::lens_grad_stateful(gl[i]) ? 1 : 0;
      }
      for (int i = 0; i < d->n_light; ++i) ok = ok && ad::light_grad_supported(d->light[i]);
      p->grad_series = gs;
      p->grad_tier = tier;
      p->n_grad_lens = ok ? ngl : 0;
      p->grad_smem = grad_smem_bytes(p->n_grad_lens, d->n_light, gs);
      p->grad_ok = ok && p->grad_smem <= kGradSmemMax;
      if (p->grad_ok && ngl > 0) TRY(p->upload(gl.data(), (size_t)ngl, &p->grad_lens));
      if (p->grad_ok && D.K > 0) {
        const size_t kk = (size_t)d->kernel_size * d->kernel_size;
        p->kernel_flip_host.resize(kk);
        for (size_t i = 0; i < kk; ++i) p->kernel_flip_host[i] = d->kernel[kk - 1 - i];
        TRY(p->upload(p->kernel_flip_host.data(), kk, &p->kernel_flip));
        if (d->kernel_separable) {
          std::vector<double> f1(d->kernel_size);
          for (int i = 0; i < d->kernel_size; ++i) f1[i] = d->kernel_1d[d->kernel_size - 1 - i];
          TRY(p->upload(f1.data(), f1.size(), &p->kernel_1d_flip));
        }
      }
    }
    if (D.ss_conv)
      p->kernel_hi_host.assign(d->kernel_hi, d->kernel_hi + (size_t)d->kernel_hi_size * d->kernel_hi_size);
    if (D.K > 0 && !d->kernel_separable)
      TRY(fft_conv_create(d->nx, d->ny, d->kernel_size, d->kernel, d->conv_mode, &p->fft_lo));
    if (D.ss_conv && !d->kernel_hi_separable)
      TRY(fft_conv_create(d->nx * d->ss, d->ny * d->ss, d->kernel_hi_size, d->kernel_hi, d->conv_mode, &p->fft_hi));
    p->image_smem = (size_t)d->n_lens * sizeof(LensC) + (size_t)d->n_light * sizeof(LightC) +
                    (size_t)n_series * kMaxSeries * sizeof(double);
    if (p->f32) p->image_smem += f32_extra_smem(d->n_lens, d->n_light, n_series);
    if (p->image_smem > 48 * 1024) {
      delete p;
      return LSIM_E_UNSUPPORTED;
    }
    if (has_data) {
      TRY(p->upload(d->data, npix, &D.data));
      std::vector<double> w(npix);
      for (size_t i = 0; i < npix; ++i) w[i] = 1.0 / d->cd_data[i];
      TRY(p->upload(w.data(), npix, &D.w_data));
      if (d->noise_map2) {
        TRY(p->upload(d->noise_map2, npix, &D.noise_map2));
      } else {
        for (size_t i = 0; i < npix; ++i) w[i] = 1.0 / d->exposure[i];
        TRY(p->upload(w.data(), npix, &D.inv_exposure));
      }
      D.bkg2 = d->background_rms * d->background_rms;
      if (d->mask) TRY(p->upload(d->mask, npix, &D.mask));
      D.logl_offset = d->logl_offset;
    }
  }
#undef TRY
  *out = p;
  return LSIM_OK;
}

void lsim_plan_destroy(lsim_plan* plan) { delete plan; }

int lsim_ray_shoot(const lsim_plan* plan, const double* params, int64_t B, const double* x, const double* y,
                   int64_t N, int coords_batched, double* out_x, double* out_y, int what, void* stream) {
  if (!plan || B < 0 || N < 0 || (what != 0 && what != 1)) return LSIM_E_INVALID;
  if (B == 0 || N == 0) return LSIM_OK;
  if (!x || !y || !out_x || !out_y || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  if (what == 1 && plan->dev.ray.multi_plane) return LSIM_E_UNSUPPORTED;
  if (B > 65535) return LSIM_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // rays per CTA: one batch for deflector lists that stream through shared memory in several chunks; otherwise enough
  // that the per-CTA parameter preparation is amortised, while keeping >= 16 waves of CTAs in the grid
  const int64_t batch = (int64_t)kBlock * kRayItems;
  int64_t rpc = batch;
  if (plan->dev.n_lens <= kLensChunk) {
    const int64_t n_batches = (N + batch - 1) / batch;
    const int64_t want_ctas = 16 * (int64_t)plan->n_sm * 4;           // per launch
    int64_t per_set = (want_ctas + B - 1) / B;                        // CTAs per parameter set
    if (per_set < 1) per_set = 1;
    int64_t bpc = (n_batches + per_set - 1) / per_set;                // ray batches per CTA
    if (bpc > 64) bpc = 64;
    if (bpc < 1) bpc = 1;
    rpc = bpc * batch;
  }
  dim3 grid((unsigned)((N + rpc - 1) / rpc), (unsigned)B);
  // one chunk of profiles + the coefficient pools of its series-bearing ones (sized per plan: three CTAs per SM for
  // the two-profile headline lens, which the 77 KB of a full chunk + the math tables would not allow)
  const size_t kRaySmem = (size_t)(plan->dev.n_lens < kLensChunk ? plan->dev.n_lens : kLensChunk) * sizeof(LensC) +
                          (size_t)(plan->n_series < kMaxSeriesProfiles ? plan->n_series : kMaxSeriesProfiles) *
                              kMaxSeries * sizeof(double);
  LSIM_SMEM((k_ray_shoot<false, TypeList<LSIM_LENS_EPL, LSIM_LENS_SHEAR>>), kRaySmem);
  LSIM_SMEM(k_ray_shoot<true>, kRaySmem);
  LSIM_SMEM(k_ray_shoot<false>, kRaySmem);
  if (plan->ray_epl_shear)
    k_ray_shoot<false, TypeList<LSIM_LENS_EPL, LSIM_LENS_SHEAR>><<<grid, kBlock, kRaySmem, st>>>(
        plan->dev, params, x, y, N, coords_batched, out_x, out_y, what, rpc);
  else if (plan->full)
    k_ray_shoot<true><<<grid, kBlock, kRaySmem, st>>>(plan->dev, params, x, y, N, coords_batched, out_x, out_y, what, rpc);
  else
    k_ray_shoot<false><<<grid, kBlock, kRaySmem, st>>>(plan->dev, params, x, y, N, coords_batched, out_x, out_y, what, rpc);
  return after_launch("k_ray_shoot");
}

int lsim_hessian(const lsim_plan* plan, const double* params, int64_t B, const double* x, const double* y, int64_t N,
                 int coords_batched, double* out, void* stream) {
  if (!plan || B < 0 || N < 0) return LSIM_E_INVALID;
  if (!plan->hess_ok) return LSIM_E_UNSUPPORTED;
  if (B == 0 || N == 0) return LSIM_OK;
  if (!x || !y || !out || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  if (B > 65535) return LSIM_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlanDev H = plan->dev;
  H.lens = plan->hess_lens;
  H.n_lens = plan->n_hess_lens;
  // enough CTAs per set to fill the machine at small B, few enough that the per-CTA preparation stays amortised
  int64_t bx = (N + kGradBlock - 1) / kGradBlock;
  int64_t cap = (8 * (int64_t)plan->n_sm + B - 1) / B;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)B);
  k_hessian<<<grid, kGradBlock, plan->hess_smem, st>>>(H, params, x, y, N, coords_batched, plan->hess_series,
                                                        plan->dev.ray.alpha_scaling, out);
  return after_launch("k_hessian");
}

int lsim_source_position_logl(const lsim_plan* plan, const double* params, int64_t B, const double* x, const double* y,
                              int n_img, int coords_batched, double sigma, double hard_bound_rms, double* logl,
                              void* stream) {
  if (!plan || B < 0 || n_img < 0 || n_img > 1024 || !(sigma > 0.0)) return LSIM_E_INVALID;
  if (!plan->hess_ok || B > 2147483647LL) return LSIM_E_UNSUPPORTED;  // one CTA per set in grid.x
  if (B == 0) return LSIM_OK;
  if (!logl || (n_img > 0 && (!x || !y)) || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  const size_t smem = srcpos_smem_bytes(plan->n_hess_lens, plan->hess_series, n_img);
  if (smem > kGradSmemMax) return LSIM_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_img == 0) {  // no images: the sum over images is empty
    LSIM_CUDA(cudaMemsetAsync(logl, 0, (size_t)B * sizeof(double), st));
    return LSIM_OK;
  }
  PlanDev H = plan->dev;
  H.lens = plan->hess_lens;
  H.n_lens = plan->n_hess_lens;
  k_source_position<<<(unsigned)B, 32, smem, st>>>(H, params, x, y, n_img, coords_batched, plan->hess_series,
                                                   plan->dev.ray.alpha_scaling, sigma, hard_bound_rms, logl);
  return after_launch("k_source_position");
}

int lsim_surface_brightness(const lsim_plan* plan, const double* params, int64_t B, const double* x,
                            const double* y, int64_t N, int coords_batched, double* out, void* stream) {
  if (!plan || B < 0 || N < 0) return LSIM_E_INVALID;
  if (B == 0 || N == 0) return LSIM_OK;
  if (!x || !y || !out || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  if (B > 65535) return LSIM_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t bx = (N + kBlock - 1) / kBlock;
  if (bx > 4096) bx = 4096;
  dim3 grid((unsigned)bx, (unsigned)B);
  k_surface_brightness<<<grid, kBlock, 0, st>>>(plan->dev, params, x, y, N, coords_batched, out);
  return after_launch("k_surface_brightness");
}

}  // extern "C"

namespace {

// Scratch doubles per parameter set.  Regular path: U (unconvolved pixels) [+ C (convolved pixels)].
// Supersampled convolution: Usub + Hsub (sub-pixel flux and its convolution) + U + C.
// bytes of the per-set prepared constants block (k_prepare -> k_image), a multiple of 8
size_t prep_bytes(const lsim_plan* plan) {
  return prepared_bytes(plan->dev.n_lens, plan->dev.n_light, plan->n_series);
}

// per-set scratch in front of the image buffers: prepared constants [+ PSF error map]
size_t prefix_bytes(const lsim_plan* plan) {
  const PlanDev& D = plan->dev;
  return prep_bytes(plan) + (D.ps_var ? (size_t)D.nx * D.ny * sizeof(double) : 0);
}

size_t per_set_bytes(const lsim_plan* plan, bool need_conv_out) {
  const PlanDev& D = plan->dev;
  const size_t npix = (size_t)D.nx * D.ny, nsub = npix * D.ss * D.ss;
  const bool conv = D.conv_mode != LSIM_CONV_NONE;
  size_t n = npix * D.n_slots;
  if (conv && need_conv_out) n += npix * D.n_slots;
  if (conv && D.ss_conv) n += 2 * nsub * D.n_slots + (need_conv_out ? 0 : npix * D.n_slots);
  return n * sizeof(double) + prefix_bytes(plan);
}

// Parameter sets per chunk (= per launch of each kernel of the step).  Forward path: 64 x SMs -- measured on the headline
// (gpurun_out/r2_variants9.txt): 4 x 269.8k, 8 x 276.6k, 16 x 279.2k, 32 x 281.0k, 64 x 282.5k images/s; the launch tails
// of the four kernels shrink with the chunk, the scratch grows (4.9 GB at 64 x for the headline) -- halved while the
// chunk's scratch would exceed 8 GiB (large frames, many slots), never below 8 x SMs.  The gradient path keeps 8 x SMs
// (its scratch per set is several times larger).  LSIM_CHUNK_MULT overrides the forward multiplier.
int64_t chunk_sets(const lsim_plan* plan, bool grad = false) {
  static const int mult = [] {
    const char* e = getenv("LSIM_CHUNK_MULT");
    const int m = e ? atoi(e) : 0;
    return m > 0 ? m : 64;
  }();
  const int64_t base = (int64_t)plan->n_sm * 8;
  if (grad) return base;
  int64_t c = (int64_t)plan->n_sm * mult;
  const size_t per = per_set_bytes(plan, true);
  while (c > base && (size_t)c * per > ((size_t)8 << 30)) c /= 2;
  return c < base ? base : c;
}

template <int NSLOT, bool FULL, int SS, class LT = RuntimeTypes, class GT = RuntimeTypes>
int launch_image_k(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, int ppt, int tiles,
                   double scale, double* out, double* sub_out, cudaStream_t st, const unsigned char* prep) {
  dim3 grid((unsigned)tiles, (unsigned)n);
  LSIM_SMEM((k_image<NSLOT, FULL, SS, LT, GT>), plan->image_smem);
  Timed timed(kClsImage, st);
  k_image<NSLOT, FULL, SS, LT, GT><<<grid, kImageBlock, plan->image_smem, st>>>(plan->dev, params, set0, ppt, scale,
                                                                                out, sub_out, prep, plan->n_series);
  return after_launch("k_image");
}

// ---- model signatures with a specialised image kernel ------------------------------------------------------------
// (effective lens types) x (effective light types, source components first).  The list covers the BASELINE.json
// configurations and their SIE variants; every other model runs the interpreted kernel.
constexpr int kSE = LSIM_LIGHT_SERSIC_ELLIPSE;
using LT_EplShear = TypeList<LSIM_LENS_EPL, LSIM_LENS_SHEAR>;
using LT_NieShear = TypeList<LSIM_LENS_NIE, LSIM_LENS_SHEAR>;
using LT_EplmShear = TypeList<LSIM_LENS_EPL_MULTIPOLE_M1M3M4, LSIM_LENS_SHEAR>;
using GT_S = TypeList<kSE>;
using GT_SS = TypeList<kSE, kSE>;
enum { kSigNone = -1, kSigEplShear_S, kSigEplShear_SS, kSigNieShear_S, kSigNieShear_SS, kSigEplmShear_SS };

int effective_lens_type(const lsim_lens_desc& d) {
  if (d.type == LSIM_LENS_SIE) return (d.flags & LSIM_FLAG_SIE_NIE_FALSE) ? LSIM_LENS_EPL : LSIM_LENS_NIE;
  return d.type;
}
int effective_light_type(const lsim_light_desc& d) {
  return d.type == LSIM_LIGHT_SERSIC ? LSIM_LIGHT_SERSIC_ELLIPSE : d.type;
}

template <class LT, class GT>
bool sig_matches(const lsim_plan_desc* d) {
  if (d->multi_plane || d->n_lens != LT::N || d->n_light != GT::N) return false;
  for (int i = 0; i < LT::N; ++i)
    if (effective_lens_type(d->lens[i]) != LT::at(i)) return false;
  for (int i = 0; i < GT::N; ++i)
    if (effective_light_type(d->light[i]) != GT::at(i)) return false;
  return true;
}

int find_signature(const lsim_plan_desc* d) {
  if (getenv("LSIM_NO_SPECIALISE")) return kSigNone;
  if (sig_matches<LT_EplShear, GT_S>(d)) return kSigEplShear_S;
  if (sig_matches<LT_EplShear, GT_SS>(d)) return kSigEplShear_SS;
  if (sig_matches<LT_NieShear, GT_S>(d)) return kSigNieShear_S;
  if (sig_matches<LT_NieShear, GT_SS>(d)) return kSigNieShear_SS;
  if (sig_matches<LT_EplmShear, GT_SS>(d)) return kSigEplmShear_SS;
  return kSigNone;
}

// returns LSIM_E_UNSUPPORTED when (NSLOT, ss) has no specialised instantiation -> caller falls back
template <bool FULL, class LT, class GT>
int launch_image_sig(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, int ppt, int tiles,
                     double scale, double* out, double* sub_out, cudaStream_t st, const unsigned char* prep) {
  const int ns = plan->dev.n_slots, ss = plan->dev.ss;
#define LSIM_SIG_CASE(NS, SSV) \
  if (ns == NS && ss == SSV) return launch_image_k<NS, FULL, SSV, LT, GT>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
  LSIM_SIG_CASE(1, 3)
  LSIM_SIG_CASE(2, 3)
  if (!FULL) {
    LSIM_SIG_CASE(1, 1)
    LSIM_SIG_CASE(1, 2)
    LSIM_SIG_CASE(2, 1)
    LSIM_SIG_CASE(2, 2)
  }
#undef LSIM_SIG_CASE
  return LSIM_E_UNSUPPORTED;
}

// Interpreted variants: the EPL/SIE/SIS/SHEAR/CONVERGENCE profile set ("FULL = false") is specialised on the
// supersampling factor (1, 2, 3 -> sub-pixels of a row evaluated in lock step); everything else runs the generic
// variant (runtime supersampling factor, one sub-pixel at a time).
template <int NSLOT>
int launch_image_ns(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, int ppt, int tiles,
                    double scale, double* out, double* sub_out, cudaStream_t st, const unsigned char* prep) {
  if (sub_out) {  // supersampled convolution: the SUB instantiations (runtime supersampling factor)
    dim3 grid((unsigned)tiles, (unsigned)n);
    Timed timed(kClsImage, st);
    const bool masked = plan->dev.eval_list != nullptr;
#define LSIM_SUB_LAUNCH(FULLV, MASKV)                                                                          \
  LSIM_SMEM((k_image<NSLOT, FULLV, 0, RuntimeTypes, RuntimeTypes, true, MASKV>), plan->image_smem);            \
  k_image<NSLOT, FULLV, 0, RuntimeTypes, RuntimeTypes, true, MASKV><<<grid, kImageBlock, plan->image_smem, st>>>( \
      plan->dev, params, set0, ppt, scale, out, sub_out, prep, plan->n_series)
    if (plan->full && masked) {
      LSIM_SUB_LAUNCH(true, true);
    } else if (plan->full) {
      LSIM_SUB_LAUNCH(true, false);
    } else if (masked) {
      LSIM_SUB_LAUNCH(false, true);
    } else {
      LSIM_SUB_LAUNCH(false, false);
    }
#undef LSIM_SUB_LAUNCH
    return after_launch("k_image");
  }
  if (plan->dev.eval_list) {  // masked evaluation: interpreted kernel, runtime supersampling factor
    dim3 grid((unsigned)tiles, (unsigned)n);
    Timed timed(kClsImage, st);
    LSIM_SMEM((k_image<NSLOT, true, 0, RuntimeTypes, RuntimeTypes, false, true>), plan->image_smem);
    LSIM_SMEM((k_image<NSLOT, false, 0, RuntimeTypes, RuntimeTypes, false, true>), plan->image_smem);
    if (plan->full)
      k_image<NSLOT, true, 0, RuntimeTypes, RuntimeTypes, false, true><<<grid, kImageBlock, plan->image_smem, st>>>(
          plan->dev, params, set0, ppt, scale, out, sub_out, prep, plan->n_series);
    else
      k_image<NSLOT, false, 0, RuntimeTypes, RuntimeTypes, false, true><<<grid, kImageBlock, plan->image_smem, st>>>(
          plan->dev, params, set0, ppt, scale, out, sub_out, prep, plan->n_series);
    return after_launch("k_image");
  }
  if (plan->full) return launch_image_k<NSLOT, true, 0>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
  switch (plan->dev.ss) {
    case 1: return launch_image_k<NSLOT, false, 1>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
    case 2: return launch_image_k<NSLOT, false, 2>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
    case 3: return launch_image_k<NSLOT, false, 3>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
    default: return launch_image_k<NSLOT, false, 0>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
  }
}

// LSIM_PRECISION_F32: the single-precision image kernel (interpreted profile lists, lock-step rows for ss <= 3)
template <int NSLOT>
int launch_image_f32_ns(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, int ppt, int tiles,
                        double scale, double* out, cudaStream_t st, const unsigned char* prep) {
  dim3 grid((unsigned)tiles, (unsigned)n);
  Timed timed(kClsImage, st);
#define LSIM_F32_LAUNCH(SSV)                                                                                   \
  LSIM_SMEM((k_image_f32<NSLOT, SSV>), plan->image_smem);                                                      \
  k_image_f32<NSLOT, SSV><<<grid, kImageBlock, plan->image_smem, st>>>(plan->dev, params, set0, ppt, scale, \
                                                                        plan->n_series, out, prep)
  switch (plan->dev.ss) {
    case 1: { LSIM_F32_LAUNCH(1); break; }
    case 2: { LSIM_F32_LAUNCH(2); break; }
    case 3: { LSIM_F32_LAUNCH(3); break; }
    default: { LSIM_F32_LAUNCH(0); break; }
  }
#undef LSIM_F32_LAUNCH
  return after_launch("k_image_f32");
}

int launch_image_f32(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, int ppt, int tiles,
                     double scale, double* out, cudaStream_t st, const unsigned char* prep) {
  const int ns = plan->dev.n_slots;
  if (ns <= 1) return launch_image_f32_ns<1>(plan, params, set0, n, ppt, tiles, scale, out, st, prep);
  if (ns <= 2) return launch_image_f32_ns<2>(plan, params, set0, n, ppt, tiles, scale, out, st, prep);
  if (ns <= 4) return launch_image_f32_ns<4>(plan, params, set0, n, ppt, tiles, scale, out, st, prep);
  return launch_image_f32_ns<8>(plan, params, set0, n, ppt, tiles, scale, out, st, prep);
}

// ray-shoot + light + re-bin for sets [set0, set0+n) -> out[n][n_slots][npix]
int launch_image(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, double scale, double* out,
                 double* sub_out, cudaStream_t st, const unsigned char* prep) {
  const int npix_all = plan->dev.nx * plan->dev.ny;
  const int npix = plan->dev.n_eval;  // evaluated pixels
  if (npix < npix_all) {              // masked pixels keep zero flux
    const size_t ns = (size_t)plan->dev.n_slots, sub = (size_t)plan->dev.ss * plan->dev.ss;
    if (out) LSIM_CUDA(cudaMemsetAsync(out, 0, (size_t)n * ns * npix_all * sizeof(double), st));
    if (sub_out) LSIM_CUDA(cudaMemsetAsync(sub_out, 0, (size_t)n * ns * npix_all * sub * sizeof(double), st));
    if (npix == 0) return LSIM_OK;
  }
  // CTAs per set: all CTAs of a launch cost the same, so a grid of w.x "waves" of resident CTAs runs like
  // ceil(w.x) waves (1184 sets x 1 CTA on 148 SMs x 3 resident = 2.67 waves took the time of 3).  Ask for >= 16
  // waves (32 by default, LSIM_IMAGE_WAVES): the partial last wave then costs ~1 %; the per-set constants come
  // from k_prepare, so small CTAs cost nothing extra.
  const int64_t resident = (int64_t)plan->n_sm * ((plan->fold && !sub_out) ? LSIM_FOLD_MIN_BLOCKS : LSIM_IMAGE_MIN_BLOCKS);
  static const int64_t waves = [] {
    const char* e = getenv("LSIM_IMAGE_WAVES");
    const int64_t v = (e && *e) ? atoll(e) : 32;
    return v < 1 ? (int64_t)1 : v;
  }();
  int64_t want = (waves * resident + n - 1) / n;
  int max_tiles = (npix + kImageBlock - 1) / kImageBlock;
  int tiles = (int)(want < 1 ? 1 : (want > max_tiles ? max_tiles : want));
  int ppt = (npix + tiles - 1) / tiles;
  ppt = ((ppt + kImageBlock - 1) / kImageBlock) * kImageBlock;
  tiles = (npix + ppt - 1) / ppt;
  const int ns = plan->dev.n_slots;
  if (prep) {  // per-set constants once per set, not once per CTA
    const int64_t items = n * (plan->dev.n_lens + plan->dev.n_light);
    if (items > 0) {
      Timed timed(kClsOther, st);
      k_prepare<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(plan->dev, params, set0, n, plan->n_series,
                                                               const_cast<unsigned char*>(prep));
      int rcp = after_launch("k_prepare");
      if (rcp != LSIM_OK) return rcp;
    }
  }
  if (plan->f32) return launch_image_f32(plan, params, set0, n, ppt, tiles, scale, out, st, prep);
  if (plan->fold && !sub_out) {
    const int nl = plan->dev.n_light, ss = plan->dev.ss;
    dim3 grid((unsigned)tiles, (unsigned)n);
    Timed timed(kClsImage, st);
#define LSIM_FOLD_CASE(NSV, SSV, NLV)                                                                               \
  if (ns == NSV && ss == SSV && nl == NLV) {                                                                        \
    LSIM_SMEM((k_image_fold<NSV, SSV, NLV>), plan->image_smem);                                                      \
    k_image_fold<NSV, SSV, NLV><<<grid, kImageBlock, plan->image_smem, st>>>(plan->dev, params, set0, ppt, scale, out, \
                                                                             prep, plan->n_series);                 \
    return after_launch("k_image_fold");                                                                            \
  }
    LSIM_FOLD_CASE(1, 3, 1) LSIM_FOLD_CASE(1, 3, 2) LSIM_FOLD_CASE(2, 3, 2)
    LSIM_FOLD_CASE(1, 2, 1) LSIM_FOLD_CASE(1, 2, 2) LSIM_FOLD_CASE(2, 2, 2)
    LSIM_FOLD_CASE(1, 1, 1) LSIM_FOLD_CASE(1, 1, 2) LSIM_FOLD_CASE(2, 1, 2)
#undef LSIM_FOLD_CASE
  }
  int rc = LSIM_E_UNSUPPORTED;
  switch ((sub_out || plan->dev.eval_list) ? kSigNone : plan->sig) {
    case kSigEplShear_S: rc = launch_image_sig<false, LT_EplShear, GT_S>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep); break;
    case kSigEplShear_SS: rc = launch_image_sig<false, LT_EplShear, GT_SS>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep); break;
    case kSigNieShear_S: rc = launch_image_sig<false, LT_NieShear, GT_S>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep); break;
    case kSigNieShear_SS: rc = launch_image_sig<false, LT_NieShear, GT_SS>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep); break;
    case kSigEplmShear_SS: rc = launch_image_sig<true, LT_EplmShear, GT_SS>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep); break;
    default: break;
  }
  if (rc != LSIM_E_UNSUPPORTED) return rc;
  if (ns <= 1) return launch_image_ns<1>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
  if (ns <= 2) return launch_image_ns<2>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
  if (ns <= 4) return launch_image_ns<4>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
  return launch_image_ns<8>(plan, params, set0, n, ppt, tiles, scale, out, sub_out, st, prep);
}

// One "same" convolution pass: n_images images of nx x ny with a K x K kernel (or its 1-d factor).
struct ConvSpec {
  int nx, ny, K;
  const double* kernel;
  const double* kernel_1d;
  int mode;
  double scale;
  int accumulate;  // out += scale * conv(in) instead of out = ...
  const FftConv* fft = nullptr;  // set: take the FFT path
  const double* kernel_host = nullptr;  // host copy of `kernel` (weights passed by value to k_conv_tiled_u)
};

// smallest m >= n of the form 2^a 3^b 5^c 7^d (cuFFT's fast sizes)
int fft_fast_size(int n) {
  for (int m = n;; ++m) {
    int r = m;
    for (int p : {2, 3, 5, 7})
      while (r % p == 0) r /= p;
    if (r == 1) return m;
  }
}

// kernels of at least this side take the FFT path (zero boundary only; LSIM_FFT_MIN_K overrides, 0 disables).
// Measured on 180 x 180 frames: the direct tiled kernel costs 18.6 ms per 16,384 images at K = 21 and grows with
// K^2, cuFFT + pad / multiply / crop costs 27.6 ms almost independently of K: the curves cross between 25 and 27.
int fft_min_k() {
  const char* s = getenv("LSIM_FFT_MIN_K");
  return (s && *s) ? atoi(s) : 27;
}

// Build the FFT state for "same" convolution of nx x ny images with `kernel` [K,K] (host).  Returns LSIM_OK with
// *out == nullptr when the FFT path does not apply (small kernel, cuFFT missing): the direct kernels are used.
int fft_conv_create(int nx, int ny, int K, const double* kernel, int mode, FftConv** out) {
  *out = nullptr;
  const int mink = fft_min_k();
  if (mode != LSIM_CONV_ZERO || mink <= 0 || K < mink || !cufft_api().ok) return LSIM_OK;
  const CufftApi& api = cufft_api();
  FftConv* f = new (std::nothrow) FftConv();
  if (!f) return LSIM_E_INVALID;
  f->nx = nx; f->ny = ny; f->K = K;
  f->PX = fft_fast_size(nx + K - 1);
  f->PY = fft_fast_size(ny + K - 1);
  const size_t nreal = (size_t)f->PX * f->PY, nspec = (size_t)f->PY * (f->PX / 2 + 1);
  int64_t nb = (int64_t)((64u << 20) / (nreal * 8 + nspec * 16));
  f->nb = (int)(nb < 1 ? 1 : (nb > 256 ? 256 : nb));
#define FFT_TRY(call)                         \
  do {                                        \
    if ((call) != cudaSuccess) {              \
      int rc__ = cuda_fail(cudaGetLastError(), #call); \
      delete f;                               \
      return rc__;                            \
    }                                         \
  } while (0)
  FFT_TRY(cudaMalloc(&f->pad, f->nb * nreal * sizeof(double)));
  FFT_TRY(cudaMalloc(&f->spec, f->nb * nspec * sizeof(cufftDoubleComplex)));
  FFT_TRY(cudaMalloc(&f->kspec, nspec * sizeof(cufftDoubleComplex)));
  int dims[2] = {f->PY, f->PX};
  if (api.PlanMany(&f->fwd, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, f->nb) != CUFFT_SUCCESS) {
    delete f;
    return LSIM_E_CUDA;
  }
  f->have_fwd = true;
  if (api.PlanMany(&f->inv, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, f->nb) != CUFFT_SUCCESS) {
    delete f;
    return LSIM_E_CUDA;
  }
  f->have_inv = true;
  // kernel spectrum: the kernel sits at the origin of the padded frame, the crop offset (K-1)/2 centres it
  std::vector<double> kp(nreal, 0.0);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) kp[(size_t)i * f->PX + j] = kernel[(size_t)i * K + j];
  FFT_TRY(cudaMemset(f->pad, 0, f->nb * nreal * sizeof(double)));
  FFT_TRY(cudaMemcpy(f->pad, kp.data(), nreal * sizeof(double), cudaMemcpyHostToDevice));
  api.SetStream(f->fwd, nullptr);
  if (api.ExecD2Z(f->fwd, f->pad, f->spec) != CUFFT_SUCCESS) {
    delete f;
    return LSIM_E_CUDA;
  }
  FFT_TRY(cudaMemcpy(f->kspec, f->spec, nspec * sizeof(cufftDoubleComplex), cudaMemcpyDeviceToDevice));
  FFT_TRY(cudaDeviceSynchronize());
#undef FFT_TRY
  *out = f;
  return LSIM_OK;
}

// n_images "same" convolutions through the FFT state, nb images per transform batch.  The scratch belongs to the
// plan: one call in flight per plan (documented in lsim.h).
int launch_conv_fft(const FftConv& f, const double* in, double* out, int64_t n_images, double scale, int accumulate,
                    cudaStream_t st) {
  const CufftApi& api = cufft_api();
  const size_t nreal = (size_t)f.PX * f.PY, nspec = (size_t)f.PY * (f.PX / 2 + 1), npix = (size_t)f.nx * f.ny;
  api.SetStream(f.fwd, st);
  api.SetStream(f.inv, st);
  const unsigned gx = (unsigned)((nreal + kBlock - 1) / kBlock);
  for (int64_t i0 = 0; i0 < n_images; i0 += f.nb) {
    const int64_t n = n_images - i0 < f.nb ? n_images - i0 : f.nb;
    Timed timed(kClsConv, st);
    k_fft_pad<<<dim3(gx > 1024 ? 1024 : gx, (unsigned)n), kBlock, 0, st>>>(in + i0 * npix, f.pad, f.nx, f.ny, f.PX, f.PY);
    int rc = after_launch("k_fft_pad");
    if (rc != LSIM_OK) return rc;
    if (api.ExecD2Z(f.fwd, f.pad, f.spec) != CUFFT_SUCCESS) return LSIM_E_CUDA;
    k_fft_mul<<<dim3(gx > 1024 ? 1024 : gx, (unsigned)n), kBlock, 0, st>>>(
        reinterpret_cast<double2*>(f.spec), reinterpret_cast<const double2*>(f.kspec), (int)nspec, 1.0 / (double)nreal);
    rc = after_launch("k_fft_mul");
    if (rc != LSIM_OK) return rc;
    if (api.ExecZ2D(f.inv, f.spec, f.pad) != CUFFT_SUCCESS) return LSIM_E_CUDA;
    k_fft_crop<<<dim3(gx > 1024 ? 1024 : gx, (unsigned)n), kBlock, 0, st>>>(f.pad, out + i0 * npix, f.nx, f.ny, f.PX, f.PY,
                                                                         (f.K - 1) / 2, scale, accumulate);
    rc = after_launch("k_fft_crop");
    if (rc != LSIM_OK) return rc;
    g_launches.fetch_add(2, std::memory_order_relaxed);  // the two cuFFT executions
  }
  return LSIM_OK;
}

template <int KT>
void conv_direct_weights(const ConvSpec& c, ConvWeights<KT>& W);

// weights by value (kernel parameter -> constant bank -> uniform registers)
template <int KT, int R, int NW>
int launch_conv_tiled_u(const ConvSpec& c, const double* in, double* out, int64_t n_images, cudaStream_t st) {
  constexpr int TH = R * NW, HALO = KT - 1, PITCH = (32 + HALO) | 1;
  constexpr size_t smem = (size_t)(TH + HALO) * PITCH * sizeof(double);
  LSIM_SMEM((k_conv_tiled_u<KT, R, NW>), smem);
  ConvWeights<KT> W;  // wT[dx][dy] = kernel flipped (correlation form), zero padded from K to KT
  conv_direct_weights<KT>(c, W);
  const int tiles = ((c.nx + 31) / 32) * ((c.ny + TH - 1) / TH);
  for (int64_t i0 = 0; i0 < n_images; i0 += 65535) {
    const int64_t ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
    const size_t off = (size_t)i0 * c.nx * c.ny;
    Timed timed(kClsConv, st);
    k_conv_tiled_u<KT, R, NW><<<dim3((unsigned)tiles, (unsigned)ni), NW * 32, smem, st>>>(
        in + off, out + off, c.nx, c.ny, c.mode, c.scale, c.accumulate, W);
    int rc = after_launch("k_conv_tiled_u");
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

template <int KT>
void conv_direct_weights(const ConvSpec& c, ConvWeights<KT>& W) {
  const int K = c.K, pad = (KT - K) / 2;
  for (int a = 0; a < KT; ++a)
    for (int b = 0; b < KT; ++b) {
      const int ka = a - pad, kb = b - pad;
      const bool inside = ka >= 0 && ka < K && kb >= 0 && kb < K;
      W.w[b * KT + a] = inside ? c.kernel_host[(K - 1 - ka) * K + (K - 1 - kb)] : 0.0;
    }
}

// The nine transformed kernels of k_conv_tiled_fxy: the column-direction transform (type b) of the kernel's rows,
// then the row-direction transform of each resulting kernel column (ka, kb, f1 of lsim_kernels.cuh).
template <int KT>
void conv_fxy_weights(const ConvSpec& c, ConvWeightsFXY<KT>& F) {
  constexpr int M = (KT + 1) / 2, WS = 3 * M - 1;
  ConvWeights<KT> D;
  conv_direct_weights<KT>(c, D);
  auto f = [&](int dy, int dx) { return (dy >= 0 && dy < KT && dx >= 0 && dx < KT) ? D.w[dx * KT + dy] : 0.0; };
  for (int b = 0; b < 3; ++b)
    for (int tx = 0; tx < M; ++tx) {
      auto g = [&](int dy) {  // column tx of the x-transformed kernel
        if (b == 0) return f(dy, 2 * tx) - f(dy, 2 * tx - 1);
        if (b == 1) return f(dy, 2 * tx) - f(dy, 2 * tx + 1);
        return f(dy, 2 * tx + 1);
      };
      double* w = F.w + (b * M + tx) * WS;
      for (int t = 0; t < M; ++t) {
        w[t] = g(2 * t) - g(2 * t - 1);
        w[M + t] = g(2 * t) - g(2 * t + 1);
        if (t < M - 1) w[2 * M + t] = g(2 * t + 1);
      }
    }
}

template <int KT, int R, int NW>
int launch_conv_tiled_fxy(const ConvSpec& c, const double* in, double* out, int64_t n_images, cudaStream_t st) {
  constexpr int TH = R * NW, HALO = KT - 1, M = (KT + 1) / 2, PITCH = (32 + M - 1) | 1;
  constexpr size_t smem = (size_t)2 * (TH + HALO) * PITCH * sizeof(double);
  LSIM_SMEM((k_conv_tiled_fxy<KT, R, NW>), smem);
  ConvWeightsFXY<KT> W;
  conv_fxy_weights<KT>(c, W);
  const int tiles = ((c.nx + 63) / 64) * ((c.ny + TH - 1) / TH);
  for (int64_t i0 = 0; i0 < n_images; i0 += 65535) {
    const int64_t ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
    const size_t off = (size_t)i0 * c.nx * c.ny;
    Timed timed(kClsConv, st);
    k_conv_tiled_fxy<KT, R, NW><<<dim3((unsigned)tiles, (unsigned)ni), NW * 32, smem, st>>>(
        in + off, out + off, c.nx, c.ny, c.mode, c.scale, c.accumulate, W);
    int rc = after_launch("k_conv_tiled_fxy");
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

template <int KT>
int launch_conv_tiled_kt(const ConvSpec& c, const double* in, double* out, int64_t n_images, cudaStream_t st) {
  // pick the tile height (64 or 60 rows) that wastes fewer rows
  const int w64 = ((c.ny + 63) / 64) * 64, w60 = ((c.ny + 59) / 60) * 60;
  if (!getenv("LSIM_CONV_NO_FF")) {  // the fast-FIR form (default); the direct form stays for A/B and as its check
    if (w60 < w64) return launch_conv_tiled_fxy<KT, 12, 5>(c, in, out, n_images, st);
    return launch_conv_tiled_fxy<KT, 16, 4>(c, in, out, n_images, st);
  }
  // 60-row tiles: 4 warps x 15 rows measured 3 % faster than 6 x 10 for K <= 21 (more register reuse per loaded
  // column); the larger kernels keep 10 rows per warp to stay inside the register budget
  // 60-row tiles: 4 warps x 15 rows for K <= 21, 6 x 10 for the larger kernels (register budget); all shapes
  // measured within 3 % of each other -- the kernel sits at 86 % of the DFMA peak whatever the shape
  if (w60 < w64 && KT <= 21) return launch_conv_tiled_u<KT, 15, 4>(c, in, out, n_images, st);
  if (w60 < w64) return launch_conv_tiled_u<KT, 10, 6>(c, in, out, n_images, st);
  return launch_conv_tiled_u<KT, 8, 8>(c, in, out, n_images, st);
}

int launch_conv_spec(const ConvSpec& c, const double* in, double* out, int64_t n_images, cudaStream_t st) {
  if (c.fft) return launch_conv_fft(*c.fft, in, out, n_images, c.scale, c.accumulate, st);
  if (!c.kernel_1d && c.kernel_host && !getenv("LSIM_CONV_GENERIC")) {
    if (c.K <= 7) return launch_conv_tiled_kt<7>(c, in, out, n_images, st);
    if (c.K <= 11) return launch_conv_tiled_kt<11>(c, in, out, n_images, st);
    if (c.K <= 15) return launch_conv_tiled_kt<15>(c, in, out, n_images, st);
    if (c.K <= 21) return launch_conv_tiled_kt<21>(c, in, out, n_images, st);
    if (c.K <= 25) return launch_conv_tiled_kt<25>(c, in, out, n_images, st);
    if (c.K <= 31) return launch_conv_tiled_kt<31>(c, in, out, n_images, st);
  }
  const int K = c.K, TW = kConvT + K - 1, pitch = TW | 1;
  const int tiles = ((c.nx + kConvT - 1) / kConvT) * ((c.ny + kConvT - 1) / kConvT);
  for (int64_t i0 = 0; i0 < n_images; i0 += 65535) {
    int64_t ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
    dim3 grid((unsigned)tiles, (unsigned)ni);
    const size_t off = (size_t)i0 * c.nx * c.ny;
    Timed timed(kClsConv, st);
    if (c.kernel_1d) {
      size_t smem = ((size_t)TW * pitch + (size_t)TW * (kConvT + 1) + K) * sizeof(double);
      if (smem > kConvSmemMax) return LSIM_E_UNSUPPORTED;
      k_conv_separable<<<grid, kBlock, smem, st>>>(in + off, out + off, c.nx, c.ny, K, c.kernel_1d, c.mode, c.scale,
                                                   c.accumulate);
    } else {
      size_t smem = ((size_t)TW * pitch + (size_t)K * K) * sizeof(double);
      if (smem > kConvSmemMax) return LSIM_E_UNSUPPORTED;
      k_conv_direct<<<grid, kBlock, smem, st>>>(in + off, out + off, c.nx, c.ny, K, c.kernel, c.mode, c.scale,
                                                c.accumulate);
    }
    int rc = after_launch("k_conv");
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

int launch_rebin(const PlanDev& D, const double* sub, double* out, int64_t n_images, double scale, cudaStream_t st) {
  const size_t npix = (size_t)D.nx * D.ny, nsub = npix * D.ss * D.ss;
  const unsigned tiles = (unsigned)((npix + kBlock - 1) / kBlock);
  for (int64_t i0 = 0; i0 < n_images; i0 += 65535) {
    const int64_t ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
    Timed timed(kClsOther, st);
    k_rebin<<<dim3(tiles, (unsigned)ni), kBlock, 0, st>>>(sub + (size_t)i0 * nsub, out + (size_t)i0 * npix, D.nx, D.ny,
                                                           D.ss, scale);
    int rc = after_launch("k_rebin");
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

// pixel-grid convolution of the plan (regular path): out = image_scale * conv(in)
int launch_conv(const lsim_plan* plan, const double* in, double* out, int64_t n_images, cudaStream_t st) {
  const PlanDev& D = plan->dev;
  ConvSpec c{D.nx, D.ny, D.K, D.kernel, D.kernel_1d, D.conv_mode, D.image_scale, 0, plan->fft_lo,
             plan->kernel_host.empty() ? nullptr : plan->kernel_host.data()};
  return launch_conv_spec(c, in, out, n_images, st);
}

// supersampled convolution (SubgridKernelConvolution.re_size_convolve, convolution.py:153-166):
//   out = image_scale * ( re_size(conv_hi(sub)) + conv_lo(low) ),  low = re_size(sub) (already in `low`)
int launch_conv_supersampled(const lsim_plan* plan, const double* sub, double* hsub, const double* low, double* out,
                             int64_t n_images, cudaStream_t st) {
  const PlanDev& D = plan->dev;
  ConvSpec hi{D.nx * D.ss, D.ny * D.ss, D.K_hi, D.kernel_hi, D.kernel_hi_1d, D.conv_mode, 1.0, 0, plan->fft_hi,
              plan->kernel_hi_host.empty() ? nullptr : plan->kernel_hi_host.data()};
  int rc = launch_conv_spec(hi, sub, hsub, n_images, st);
  if (rc != LSIM_OK) return rc;
  rc = launch_rebin(D, hsub, out, n_images, D.image_scale, st);
  if (rc != LSIM_OK) return rc;
  if (D.K > 0) {
    ConvSpec lo{D.nx, D.ny, D.K, D.kernel, D.kernel_1d, D.conv_mode, D.image_scale, 1, plan->fft_lo,
                plan->kernel_host.empty() ? nullptr : plan->kernel_host.data()};
    rc = launch_conv_spec(lo, low, out, n_images, st);
  }
  return rc;
}

// One chunk of the forward path: ray-shoot + light + re-bin (+ PSF) for sets [s0, s0+n).
//   ws   : scratch, laid out for `chunk` sets
//   dst  : where the final (convolved, scaled) basis images [n][n_slots][npix] go
int run_chunk(const lsim_plan* plan, const double* params, int64_t s0, int64_t n, int64_t chunk, double* ws,
              double* dst, cudaStream_t st) {
  const PlanDev& D = plan->dev;
  const size_t npix = (size_t)D.nx * D.ny, nsub = npix * D.ss * D.ss, ns = (size_t)D.n_slots;
  // the scratch starts with the prepared-constants blocks of the chunk (none when the caller gave no scratch)
  const unsigned char* prep = reinterpret_cast<const unsigned char*>(ws);
  if (ws) ws += (size_t)chunk * prefix_bytes(plan) / sizeof(double);
  int rc;
  if (D.conv_mode == LSIM_CONV_NONE) {
    rc = launch_image(plan, params, s0, n, D.image_scale, dst, nullptr, st, prep);
  } else if (!D.ss_conv) {
    double* U = ws;  // [chunk][ns][npix]
    rc = launch_image(plan, params, s0, n, 1.0, U, nullptr, st, prep);
    if (rc == LSIM_OK) rc = launch_conv(plan, U, dst, n * D.n_slots, st);
  } else {
    double* U = ws;
    double* Usub = ws + (size_t)chunk * ns * npix;
    double* Hsub = Usub + (size_t)chunk * ns * nsub;
    rc = launch_image(plan, params, s0, n, 1.0, U, Usub, st, prep);
    if (rc == LSIM_OK) rc = launch_conv_supersampled(plan, Usub, Hsub, U, dst, n * D.n_slots, st);
  }
  if (rc != LSIM_OK || D.n_points == 0) return rc;
  Timed timed(kClsOther, st);
  k_point_sources<<<dim3((unsigned)D.n_slots, (unsigned)n), kBlock, 0, st>>>(D, params, s0, dst);
  return after_launch("k_point_sources");
}

int launch_likelihood(const lsim_plan* plan, const double* conv, int64_t set0, int64_t n, int linear, int marg,
                      double* logl, double* amps, double* model, const double* err, cudaStream_t st) {
  const int ns = plan->dev.n_slots;
  dim3 grid((unsigned)n);
  Timed timed(kClsLike, st);
  if (ns <= 1)
    k_likelihood<1><<<grid, kBlock, 0, st>>>(plan->dev, conv, set0, linear, marg, logl, amps, model, err);
  else if (ns <= 2)
    k_likelihood<2><<<grid, kBlock, 0, st>>>(plan->dev, conv, set0, linear, marg, logl, amps, model, err);
  else if (ns <= 4)
    k_likelihood<4><<<grid, kBlock, 0, st>>>(plan->dev, conv, set0, linear, marg, logl, amps, model, err);
  else
    k_likelihood<8><<<grid, kBlock, 0, st>>>(plan->dev, conv, set0, linear, marg, logl, amps, model, err);
  return after_launch("k_likelihood");
}

// ---- gradient path ---------------------------------------------------------------------------------------------------
// device view of the plan with the gradient's lens list in place of the forward one
PlanDev grad_dev(const lsim_plan* plan) {
  PlanDev G = plan->dev;
  G.lens = plan->grad_lens;
  G.n_lens = plan->n_grad_lens;
  return G;
}

int grad_max_tiles(const lsim_plan* plan) { return (plan->dev.n_eval + kGradBlock - 1) / kGradBlock; }

size_t grad_per_set_bytes(const lsim_plan* plan) {
  const PlanDev& D = plan->dev;
  const size_t img = (size_t)D.nx * D.ny * D.n_slots;
  return per_set_bytes(plan, true) + img * sizeof(double) +
         (size_t)grad_max_tiles(plan) * grad_entries(plan->n_grad_lens, D.n_light) * sizeof(double);
}

template <int NS>
int launch_like_adjoint_ns(const lsim_plan* plan, const double* conv, int64_t set0, int64_t n, int linear, double* logl,
                           double* amps, double* abar, cudaStream_t st) {
  Timed timed(kClsLike, st);
  k_like_adjoint<NS><<<(unsigned)n, kBlock, 0, st>>>(plan->dev, conv, set0, linear, logl, amps, abar);
  return after_launch("k_like_adjoint");
}

int launch_like_adjoint(const lsim_plan* plan, const double* conv, int64_t set0, int64_t n, int linear, double* logl,
                        double* amps, double* abar, cudaStream_t st) {
  const int ns = plan->dev.n_slots;
  if (ns <= 1) return launch_like_adjoint_ns<1>(plan, conv, set0, n, linear, logl, amps, abar, st);
  if (ns <= 2) return launch_like_adjoint_ns<2>(plan, conv, set0, n, linear, logl, amps, abar, st);
  if (ns <= 4) return launch_like_adjoint_ns<4>(plan, conv, set0, n, linear, logl, amps, abar, st);
  return launch_like_adjoint_ns<8>(plan, conv, set0, n, linear, logl, amps, abar, st);
}

// Ubar = image_scale * psf^T (*) Abar for n_images frames; `tmp` is a second frame buffer (edge mode only)
int launch_conv_transposed(const lsim_plan* plan, const double* abar, double* tmp, double* ubar, int64_t n_images,
                           cudaStream_t st) {
  const PlanDev& D = plan->dev;
  if (D.conv_mode == LSIM_CONV_ZERO) {
    ConvSpec c{D.nx, D.ny, D.K, plan->kernel_flip, plan->kernel_1d_flip, LSIM_CONV_ZERO, D.image_scale, 0, nullptr,
               plan->kernel_flip_host.data()};
    return launch_conv_spec(c, abar, ubar, n_images, st);
  }
  const unsigned tiles = (unsigned)(((size_t)D.nx * D.ny + kBlock - 1) / kBlock);
  for (int64_t i0 = 0; i0 < n_images; i0 += 65535) {
    const int64_t ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
    const size_t off = (size_t)i0 * D.nx * D.ny;
    for (int axis = 1; axis >= 0; --axis) {  // the forward pass convolves along x, then along y (either order is exact)
      Timed timed(kClsConv, st);
      k_conv_adj_edge<<<dim3(tiles, (unsigned)ni), kBlock, 0, st>>>(axis == 1 ? abar + off : tmp + off,
                                                                    axis == 1 ? tmp + off : ubar + off, D.nx, D.ny, D.K,
                                                                    D.kernel_1d, axis, axis == 1 ? 1.0 : D.image_scale);
      int rc = after_launch("k_conv_adj_edge");
      if (rc != LSIM_OK) return rc;
    }
  }
  return LSIM_OK;
}

int launch_image_grad(const lsim_plan* plan, const double* params, int64_t set0, int64_t n, const double* ubar,
                      double wscale, double* partial, double* grad, cudaStream_t st) {
  const PlanDev D = grad_dev(plan);
  const int npix = D.n_eval;
  const int max_tiles = grad_max_tiles(plan);
  int64_t want = ((int64_t)plan->n_sm * 24 + n - 1) / n;  // >= 8 waves of 3 resident CTAs
  int tiles = (int)(want < 1 ? 1 : (want > max_tiles ? max_tiles : want));
  if (tiles < 1) tiles = 1;
  int ppt = (npix + tiles - 1) / tiles;
  ppt = ((ppt + kGradBlock - 1) / kGradBlock) * kGradBlock;
  if (ppt < kGradBlock) ppt = kGradBlock;
  tiles = npix > 0 ? (npix + ppt - 1) / ppt : 1;
  for (int64_t i0 = 0; i0 < n; i0 += 65535) {
    const int64_t ni = n - i0 < 65535 ? n - i0 : 65535;
    const dim3 grid((unsigned)tiles, (unsigned)ni);
    const double* ub = ubar + (size_t)i0 * D.n_slots * D.nx * D.ny;
    double* pt = partial + (size_t)i0 * tiles * grad_entries(D.n_lens, D.n_light);
    {
      Timed timed(kClsImage, st);
#define LSIM_GRAD_LAUNCH(NS)                                                                                      \
  if (plan->grad_tier == 0)                                                                                       \
    k_image_grad<NS, 0><<<grid, kGradBlock, plan->grad_smem, st>>>(D, params, set0 + i0, ppt, plan->grad_series, ub, wscale, pt); \
  else if (plan->grad_tier == 1)                                                                                  \
    k_image_grad<NS, 1><<<grid, kGradBlock, plan->grad_smem, st>>>(D, params, set0 + i0, ppt, plan->grad_series, ub, wscale, pt); \
  else                                                                                                            \
    k_image_grad<NS, 2><<<grid, kGradBlock, plan->grad_smem, st>>>(D, params, set0 + i0, ppt, plan->grad_series, ub, wscale, pt);
      if (plan->grad_mp) {
        k_image_grad<0, 2, true><<<grid, kGradBlock, plan->grad_smem, st>>>(D, params, set0 + i0, ppt, plan->grad_series,
                                                                            ub, wscale, pt);
      } else {
        switch (plan->grad_stateful < 2 ? plan->grad_stateful : 2) {  // EPL-family / NIE profiles kept in registers
          case 1: LSIM_GRAD_LAUNCH(1) break;
          case 2: LSIM_GRAD_LAUNCH(2) break;
          default: LSIM_GRAD_LAUNCH(0) break;
        }
      }
#undef LSIM_GRAD_LAUNCH
      int rc = after_launch("k_image_grad");
      if (rc != LSIM_OK) return rc;
    }
    Timed timed(kClsOther, st);
    k_grad_reduce<<<(unsigned)ni, 128, grad_entries(D.n_lens, D.n_light) * sizeof(double), st>>>(D, pt, tiles, set0 + i0,
                                                                                                 grad);
    int rc = after_launch("k_grad_reduce");
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

}  // namespace

extern "C" {

size_t lsim_logl_grad_workspace_bytes(const lsim_plan* plan, int64_t B) {
  if (!plan || !plan->has_image || !plan->grad_ok || B <= 0) return 0;
  int64_t c = chunk_sets(plan, true);
  if (B < c) c = B;
  return grad_per_set_bytes(plan) * (size_t)c;
}

int lsim_logl_grad(const lsim_plan* plan, const double* params, int64_t B, int linear_solve, double* logl, double* grad,
                   double* amps, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || B < 0) return LSIM_E_INVALID;
  if (!plan->has_image || !plan->has_data) return LSIM_E_NO_IMAGE;
  if (!plan->grad_ok) return LSIM_E_UNSUPPORTED;
  if (B == 0) return LSIM_OK;
  if (!logl || !grad || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PlanDev& D = plan->dev;
  const bool conv = D.conv_mode != LSIM_CONV_NONE;
  int64_t chunk = (int64_t)(workspace_bytes / grad_per_set_bytes(plan));
  if (!workspace || chunk < 1) return LSIM_E_WORKSPACE;
  if (chunk > chunk_sets(plan, true)) chunk = chunk_sets(plan, true);
  const size_t img = (size_t)D.nx * D.ny * D.n_slots;
  // layout: the forward scratch of `chunk` sets [prepared constants | U | C], then Abar, then the per-CTA partials
  double* ws = static_cast<double*>(workspace);
  double* U = ws + (size_t)chunk * prefix_bytes(plan) / sizeof(double);
  double* C = conv ? U + (size_t)chunk * img : U;
  double* G = ws + (size_t)chunk * per_set_bytes(plan, true) / sizeof(double);
  double* partial = G + (size_t)chunk * img;
  for (int64_t s0 = 0; s0 < B; s0 += chunk) {
    const int64_t n = B - s0 < chunk ? B - s0 : chunk;
    int rc = run_chunk(plan, params, s0, n, chunk, ws, C, st);
    if (rc != LSIM_OK) return rc;
    rc = launch_like_adjoint(plan, C, s0, n, linear_solve, logl, amps, G, st);
    if (rc != LSIM_OK) return rc;
    const double* ubar = G;
    if (conv) {
      rc = launch_conv_transposed(plan, G, C, U, n * D.n_slots, st);
      if (rc != LSIM_OK) return rc;
      ubar = U;
    }
    rc = launch_image_grad(plan, params, s0, n, ubar, conv ? 1.0 : D.image_scale, partial, grad, st);
    if (rc != LSIM_OK) return rc;
    if (D.n_points > 0) {  // after k_grad_reduce has written grad[s0 ... s0 + n): the stamps' own parameters
      for (int64_t i0 = 0; i0 < n; i0 += 65535) {
        const int64_t ni = n - i0 < 65535 ? n - i0 : 65535;
        Timed timed(kClsOther, st);
        k_point_sources_grad<<<dim3((unsigned)D.n_points, (unsigned)ni), kBlock, 0, st>>>(
            D, params, s0 + i0, G + (size_t)i0 * img, grad);
        rc = after_launch("k_point_sources_grad");
        if (rc != LSIM_OK) return rc;
      }
    }
  }
  return LSIM_OK;
}

size_t lsim_workspace_bytes(const lsim_plan* plan, int64_t B) {
  if (!plan || !plan->has_image || B <= 0) return 0;
  int64_t c = chunk_sets(plan);
  if (B < c) c = B;
  return per_set_bytes(plan, true) * (size_t)c;
}

int lsim_image(const lsim_plan* plan, const double* params, int64_t B, double* images, void* workspace,
               size_t workspace_bytes, void* stream) {
  if (!plan || B < 0) return LSIM_E_INVALID;
  if (!plan->has_image) return LSIM_E_NO_IMAGE;
  if (B == 0) return LSIM_OK;
  if (!images || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PlanDev& D = plan->dev;
  const size_t img = (size_t)D.nx * D.ny * D.n_slots;
  int64_t chunk = 32768;
  if (D.conv_mode != LSIM_CONV_NONE) {
    chunk = (int64_t)(workspace_bytes / per_set_bytes(plan, false));
    if (!workspace || chunk < 1) return LSIM_E_WORKSPACE;
    if (chunk > chunk_sets(plan)) chunk = chunk_sets(plan);
  } else if (workspace && prep_bytes(plan) > 0) {
    // no PSF: the images go straight to the caller's buffer; scratch (optional) only holds the prepared constants
    const int64_t fit = (int64_t)(workspace_bytes / prefix_bytes(plan));
    if (fit < 1) workspace = nullptr;
    else if (fit < chunk) chunk = fit;
  } else {
    workspace = nullptr;
  }
  for (int64_t s0 = 0; s0 < B; s0 += chunk) {
    const int64_t n = B - s0 < chunk ? B - s0 : chunk;
    int rc = run_chunk(plan, params, s0, n, chunk, static_cast<double*>(workspace), images + (size_t)s0 * img, st);
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

int lsim_resize_convolve(const lsim_plan* plan, const double* flux, int64_t n_images, int unconvolved,
                         double* images, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || n_images < 0) return LSIM_E_INVALID;
  if (!plan->has_image) return LSIM_E_NO_IMAGE;
  if (n_images == 0) return LSIM_OK;
  if (!flux || !images) return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PlanDev& D = plan->dev;
  const size_t npix = (size_t)D.nx * D.ny, nsub = npix * D.ss * D.ss;
  const bool conv = D.conv_mode != LSIM_CONV_NONE && !unconvolved;
  if (!conv) return launch_rebin(D, flux, images, n_images, D.image_scale, st);
  // scratch per image: re-binned unconvolved pixels (+ convolved sub-pixels for the supersampled convolution)
  const size_t per = (npix + (D.ss_conv ? nsub : 0)) * sizeof(double);
  int64_t chunk = (int64_t)(workspace_bytes / per);
  if (!workspace || chunk < 1) return LSIM_E_WORKSPACE;
  if (chunk > 65535) chunk = 65535;
  double* low = static_cast<double*>(workspace);
  double* hsub = low + (size_t)chunk * npix;
  for (int64_t i0 = 0; i0 < n_images; i0 += chunk) {
    const int64_t n = n_images - i0 < chunk ? n_images - i0 : chunk;
    int rc = launch_rebin(D, flux + (size_t)i0 * nsub, low, n, 1.0, st);
    if (rc != LSIM_OK) return rc;
    if (D.ss_conv)
      rc = launch_conv_supersampled(plan, flux + (size_t)i0 * nsub, hsub, low, images + (size_t)i0 * npix, n, st);
    else
      rc = launch_conv(plan, low, images + (size_t)i0 * npix, n, st);
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

int lsim_logl(const lsim_plan* plan, const double* params, int64_t B, int linear_solve, int source_marg,
              double* logl, double* amps, double* model, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || B < 0) return LSIM_E_INVALID;
  if (!plan->has_image || !plan->has_data) return LSIM_E_NO_IMAGE;
  if (B == 0) return LSIM_OK;
  if (!logl || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PlanDev& D = plan->dev;
  const bool conv = D.conv_mode != LSIM_CONV_NONE;
  const size_t per = per_set_bytes(plan, true);
  int64_t chunk = (int64_t)(workspace_bytes / per);
  if (!workspace || chunk < 1) return LSIM_E_WORKSPACE;
  if (chunk > chunk_sets(plan)) chunk = chunk_sets(plan);
  const size_t img = (size_t)D.nx * D.ny * D.n_slots, sub = img * D.ss * D.ss;
  // layout: [prepared constants | PSF error map] [U | Usub, Hsub (supersampled conv only) | C]; without a PSF the images
  // land in U directly.  (Running the likelihood of one half buffer on a side stream while the fp64-bound kernels fill
  // the other was measured twice in round 2 and bought nothing: DESIGN.md 4.)
  double* ws = static_cast<double*>(workspace);
  double* main_ws = ws + (size_t)chunk * prefix_bytes(plan) / sizeof(double);  // run_chunk skips the same prefix
  double* err = D.ps_var ? ws + (size_t)chunk * prep_bytes(plan) / sizeof(double) : nullptr;  // [chunk][npix]
  double* C = conv ? main_ws + (size_t)chunk * img + (D.ss_conv ? 2 * (size_t)chunk * sub : 0) : main_ws;
  for (int64_t s0 = 0; s0 < B; s0 += chunk) {
    const int64_t n = B - s0 < chunk ? B - s0 : chunk;
    int rc = run_chunk(plan, params, s0, n, chunk, ws, C, st);
    if (rc != LSIM_OK) return rc;
    if (err) {
      Timed timed(kClsOther, st);
      k_ps_error_map<<<(unsigned)n, kBlock, 0, st>>>(D, params, s0, err);
      rc = after_launch("k_ps_error_map");
      if (rc != LSIM_OK) return rc;
    }
    rc = launch_likelihood(plan, C, s0, n, linear_solve, source_marg, logl, amps, model, err, st);
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

int lsim_psf_error_map(const lsim_plan* plan, const double* params, int64_t B, double* error_map, void* stream) {
  if (!plan || B < 0) return LSIM_E_INVALID;
  if (!plan->has_image) return LSIM_E_NO_IMAGE;
  if (B == 0) return LSIM_OK;
  if (!error_map || (plan->dev.n_param_cols > 0 && !params)) return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PlanDev& D = plan->dev;
  if (!D.ps_var) {
    LSIM_CUDA(cudaMemsetAsync(error_map, 0, (size_t)B * D.nx * D.ny * sizeof(double), st));
    return LSIM_OK;
  }
  for (int64_t s0 = 0; s0 < B; s0 += 65535) {
    const int64_t n = B - s0 < 65535 ? B - s0 : 65535;
    k_ps_error_map<<<(unsigned)n, kBlock, 0, st>>>(D, params, s0, error_map + (size_t)s0 * D.nx * D.ny);
    int rc = after_launch("k_ps_error_map");
    if (rc != LSIM_OK) return rc;
  }
  return LSIM_OK;
}

int lsim_timing_enable(int on) {
  g_timing.store(on ? 1 : 0);
  return LSIM_OK;
}

int lsim_timing_read(double ms[4], int64_t count[4]) {
  if (!ms || !count) return LSIM_E_INVALID;
  for (int i = 0; i < 4; ++i) {
    ms[i] = 0.0;
    count[i] = 0;
  }
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (TimedEvent& ev : g_timed) {
    LSIM_CUDA(cudaEventSynchronize(ev.e1));
    float t = 0;
    LSIM_CUDA(cudaEventElapsedTime(&t, ev.e0, ev.e1));
    ms[ev.cls] += t;
    count[ev.cls] += 1;
    cudaEventDestroy(ev.e0);
    cudaEventDestroy(ev.e1);
  }
  g_timed.clear();
  return LSIM_OK;
}

int lsim_pso_step(double* position, double* velocity, double* pbest_pos, double* pbest_fit, double* gbest_pos,
                  double* gbest_fit, const double* fitness, const double* rnd, int64_t B, int32_t P, double c1,
                  double c2, void* stream) {
  if (B < 0 || P <= 0) return LSIM_E_INVALID;
  if (B == 0) return LSIM_OK;
  if (!position || !velocity || !pbest_pos || !pbest_fit || !gbest_pos || !gbest_fit || !fitness || !rnd)
    return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_pso_best<<<1, kPsoBlock, 0, st>>>(position, fitness, pbest_pos, pbest_fit, gbest_pos, gbest_fit, B, P);
  int rc = after_launch("k_pso_best");
  if (rc != LSIM_OK) return rc;
  const int64_t n = B * P;
  k_pso_move<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, st>>>(position, velocity, pbest_pos, gbest_pos, rnd,
                                                                      n, P, c1, c2);
  return after_launch("k_pso_move");
}

int lsim_measure_fp64_peak(int64_t iters, double* flops, void* stream) {
  if (!flops || iters <= 0) return LSIM_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, n_sm = 148;
  LSIM_CUDA(cudaGetDevice(&dev));
  LSIM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  LSIM_CUDA(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  LSIM_CUDA(cudaEventCreate(&e0));
  LSIM_CUDA(cudaEventCreate(&e1));
  const int ctas = n_sm * 8;
  k_fp64_peak<<<ctas, kBlock, 0, st>>>(iters / 8 + 1, sink);  // warm-up
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {  // best of 3: the roofline denominator must not be an unlucky sample
    LSIM_CUDA(cudaEventRecord(e0, st));
    k_fp64_peak<<<ctas, kBlock, 0, st>>>(iters, sink);
    LSIM_CUDA(cudaEventRecord(e1, st));
    LSIM_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    LSIM_CUDA(cudaEventElapsedTime(&t, e0, e1));
    if (rep == 0 || t < ms) ms = t;
  }
  g_launches.fetch_add(4);
  *flops = 2.0 * 8.0 * (double)iters * (double)ctas * kBlock / (ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return LSIM_OK;
}

}  // extern "C"

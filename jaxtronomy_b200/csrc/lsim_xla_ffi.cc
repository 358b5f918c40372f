// lsim_xla_ffi.cc -- optional XLA-FFI custom-call shim over the C ABI of include/lsim.h.
//
// JAXtronomy runs its hot path as XLA programs (Sampling/sampler.py:215-251 jit(vmap(logL))); with this shim the
// same Python entry points call liblsim.so through jax.ffi.ffi_call (INTEGRATION.md section 3 shows the Python
// side).  The jax headers are NOT in this image, so the file is header-gated: it compiles to an empty object
// unless "xla/ffi/api/ffi.h" is on the include path:
//
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../include \
//       -I/usr/local/cuda/include lsim_xla_ffi.cc -o liblsim_xla_ffi.so -L. -llsim -Wl,-rpath,'$ORIGIN'
//
// Handlers keep no global state: a plan is immutable after lsim_plan_create and the launch calls only enqueue
// work on the stream XLA hands in, so XLA may call them concurrently from several host threads (one plan per
// device).  The plan pointer travels as an int64 attribute; scratch is an XLA-allocated operand.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define LSIM_HAVE_XLA_FFI 1
#endif
#endif

#ifdef LSIM_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "lsim.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) {
  return rc == LSIM_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, lsim_strerror(rc));
}
static const lsim_plan* Plan(int64_t p) { return reinterpret_cast<const lsim_plan*>(p); }

// ImageLinearFit.likelihood_data_given_model / ImageModel.likelihood_data_given_model (image_linear_solve.py:228-339)
static ffi::Error LoglImpl(cudaStream_t stream, int64_t plan, int32_t linear, int32_t marg,
                           ffi::Buffer<ffi::F64> params,      // [B, P]
                           ffi::Buffer<ffi::F64> workspace,   // [n] scratch, n >= lsim_workspace_bytes(plan, 1) / 8
                           ffi::ResultBuffer<ffi::F64> logl,  // [B]
                           ffi::ResultBuffer<ffi::F64> amps) {  // [B, n_slots]
  const int64_t B = params.dimensions()[0];
  return Status(lsim_logl(Plan(plan), params.typed_data(), B, linear, marg, logl->typed_data(),
                          linear ? amps->typed_data() : nullptr, nullptr, workspace.typed_data(),
                          workspace.element_count() * sizeof(double), stream));
}

// jax.value_and_grad(LikelihoodModule.logL) (Sampling/Samplers/optax.py:39): logL and d logL / d params; the Python side
// wraps this handler in a jax.custom_vjp whose backward rule is ct[:, None] * grad
static ffi::Error LoglGradImpl(cudaStream_t stream, int64_t plan, int32_t linear,
                               ffi::Buffer<ffi::F64> params,      // [B, P]
                               ffi::Buffer<ffi::F64> workspace,   // >= lsim_logl_grad_workspace_bytes(plan, 1) / 8
                               ffi::ResultBuffer<ffi::F64> logl,  // [B]
                               ffi::ResultBuffer<ffi::F64> grad,  // [B, P]
                               ffi::ResultBuffer<ffi::F64> amps) {  // [B, n_slots]
  const int64_t B = params.dimensions()[0];
  return Status(lsim_logl_grad(Plan(plan), params.typed_data(), B, linear, logl->typed_data(), grad->typed_data(),
                               linear ? amps->typed_data() : nullptr, workspace.typed_data(),
                               workspace.element_count() * sizeof(double), stream));
}

// ImageModel.image / rows of linear_response_matrix (image_model.py:431-492, image_linear_solve.py:361-455)
static ffi::Error ImageImpl(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::F64> params,
                            ffi::Buffer<ffi::F64> workspace, ffi::ResultBuffer<ffi::F64> images) {  // [B, n_slots, ny, nx]
  return Status(lsim_image(Plan(plan), params.typed_data(), params.dimensions()[0], images->typed_data(),
                           workspace.typed_data(), workspace.element_count() * sizeof(double), stream));
}

// LensModel.ray_shooting / SinglePlane.alpha (lens_model.py:236-252, single_plane.py:112-139); x, y: [N] shared
static ffi::Error RayShootImpl(cudaStream_t stream, int64_t plan, int32_t what, ffi::Buffer<ffi::F64> params,
                               ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y, ffi::ResultBuffer<ffi::F64> out_x,
                               ffi::ResultBuffer<ffi::F64> out_y) {  // [B, N] each
  const bool batched = x.dimensions().size() == 2;
  const int64_t N = x.dimensions()[batched ? 1 : 0];
  return Status(lsim_ray_shoot(Plan(plan), params.typed_data(), params.dimensions()[0], x.typed_data(),
                               y.typed_data(), N, batched ? 1 : 0, out_x->typed_data(), out_y->typed_data(), what,
                               stream));
}

// LensModel.hessian, diff=None (lens_model.py:384-414, single_plane.py:142-180); x, y: [N] shared or [B, N]
static ffi::Error HessianImpl(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::F64> x,
                              ffi::Buffer<ffi::F64> y, ffi::ResultBuffer<ffi::F64> out) {  // [B, 4, N]
  const bool batched = x.dimensions().size() == 2;
  const int64_t N = x.dimensions()[batched ? 1 : 0];
  return Status(lsim_hessian(Plan(plan), params.typed_data(), params.dimensions()[0], x.typed_data(), y.typed_data(),
                             N, batched ? 1 : 0, out->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(LsimLogl, LoglImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("linear")
                                  .Attr<int32_t>("marg")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LsimLoglGrad, LoglGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("linear")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LsimImage, ImageImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LsimRayShoot, RayShootImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("what")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LsimHessian, HessianImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
#else
// jax headers absent: nothing to build.  The symbol below lets a build script detect the empty shim.
extern "C" int lsim_xla_ffi_unavailable(void) { return 1; }
#endif

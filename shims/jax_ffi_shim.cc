// jax.ffi registration shim for libed2b200 — COMPILE-GATED: needs jaxlib's xla/ffi/api/ffi.h, which is not in this
// image, so this file is NOT built by edward2_b200/_build.py and is UNTESTED here.
// Build (where jaxlib is installed):
//   g++ -shared -fPIC -std=c++17 -I$(python -c "import jax.ffi,sys;sys.stdout.write(jax.ffi.include_dir())") \
//       -I../include jax_ffi_shim.cc -L../edward2_b200 -led2b200 -o libed2b200_jax.so
// Python side: see INTEGRATION.md §2.
#if __has_include("xla/ffi/api/ffi.h")
#include "xla/ffi/api/ffi.h"

#include "ed2b200.h"

namespace ffi = xla::ffi;

// y = act(((x ∘ alpha[e]) W) ∘ gamma[e] + bias[e])  — edward2/jax/nn/dense.py:258-272
static ffi::Error BeDenseFwd(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer w, ffi::Buffer<ffi::F32> alpha,
                             ffi::Buffer<ffi::F32> gamma, ffi::Buffer<ffi::F32> bias, int32_t act,
                             ffi::Result<ffi::AnyBuffer> y, ffi::Result<ffi::AnyBuffer> xa,
                             ffi::Result<ffi::AnyBuffer> z) {
  ed2_be_fwd_args a{};
  a.dtype = x.element_type() == ffi::BF16 ? ED2_BF16 : ED2_F32;
  a.batch = static_cast<int>(x.dimensions()[0]);
  a.in_dim = static_cast<int>(x.dimensions()[1]);
  a.out_dim = static_cast<int>(w.dimensions()[1]);
  a.ensemble_size = static_cast<int>(alpha.dimensions()[0]);
  a.act = act;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.w_lp = w.untyped_data(); a.w_ld = a.out_dim;
  a.alpha = alpha.typed_data(); a.gamma = gamma.typed_data(); a.bias = bias.typed_data();
  a.y = y->untyped_data(); a.y_ld = a.out_dim;
  a.xa = xa->untyped_data(); a.xa_ld = a.in_dim;
  a.z = z->untyped_data(); a.z_ld = a.out_dim;
  if (ed2_be_dense_fwd(&a, stream) != ED2_OK) return ffi::Error::Internal(ed2_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2BeDenseFwd, BeDenseFwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("act")
                                  .Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>());
// The remaining entry points (ed2_be_dense_bwd, ed2_rff_fwd/bwd, ed2_laplace_precision_update,
// ed2_predictive_variance, ed2_spectral_norm_update) bind the same way: buffers -> struct fields, stream from Ctx.
#else
#error "xla/ffi/api/ffi.h not found: this shim only builds where jaxlib is installed (see INTEGRATION.md)"
#endif

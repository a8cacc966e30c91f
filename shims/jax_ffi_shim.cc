// jax.ffi registration shim for libed2b200: one XLA FFI handler per layer-level entry point of include/ed2b200.h
// (forward AND backward of the four dense layers, the SNGP head operators, the KL regulariser).  The Python side
// (jax.ffi.register_ffi_target + jax.custom_vjp wiring) is shown in INTEGRATION.md §2.
//
// jaxlib's xla/ffi/api/ffi.h is not in this image, so the shim cannot be BUILT here; it is TYPE-CHECKED against
// shims/mock/xla/ffi/api/ffi.h (tests/test_host.py::test_shims_type_check: every handler must be invocable with exactly
// the argument list its binding declares, every struct field it touches must exist in ed2b200.h).
// Build where jaxlib is installed:
//   g++ -shared -fPIC -std=c++17 -I$(python -c "import jax.ffi,sys;sys.stdout.write(jax.ffi.include_dir())") \
//       -I../include jax_ffi_shim.cc -L../edward2_b200 -led2b200 -o libed2b200_jax.so
//
// Conventions: XLA buffers are dense row-major, so every leading dimension is the buffer's last dimension (bf16
// feature sizes must be multiples of 8: TMA rows are 16-byte aligned); workspaces are extra results (XLA allocates
// them); noise is Philox (seed, stream) attributes, per-example streams take the data-parallel row mapping as
// attributes (row0, rows_per_group_local, rows_per_group_global) — 0, 0, 0 on a single device.
#include "xla/ffi/api/ffi.h"

#include "ed2b200.h"

namespace ffi = xla::ffi;
using Any = ffi::AnyBuffer;
using F32B = ffi::Buffer<ffi::F32>;
using F64B = ffi::Buffer<ffi::F64>;
template <typename T> using Out = ffi::Result<T>;

namespace {

inline int dt_of(const Any& b) { return b.element_type() == ffi::BF16 ? ED2_BF16 : ED2_F32; }
inline int dim(const Any& b, int i) { return static_cast<int>(b.dimensions()[i]); }
inline int dim(const F32B& b, int i) { return static_cast<int>(b.dimensions()[i]); }
inline ed2_noise philox(int64_t seed, int64_t stream, int64_t row0 = 0, int32_t rpg_local = 0, int32_t rpg_global = 0) {
  ed2_noise n{};
  n.seed = static_cast<uint64_t>(seed);
  n.stream = static_cast<uint64_t>(stream);
  n.row0 = row0;
  n.rows_per_group_local = rpg_local;
  n.rows_per_group_global = rpg_global;
  return n;
}
inline ffi::Error done(int status) {
  return status == ED2_OK ? ffi::Error::Success() : ffi::Error::Internal(ed2_last_error());
}

// ---------------------------------------------------------------------------------------------- DenseBatchEnsemble
// y = act(((x ∘ alpha[e]) W) ∘ gamma[e] + bias[e])  — edward2/jax/nn/dense.py:258-272
ffi::Error BeDenseFwd(cudaStream_t s, Any x, Any w, F32B alpha, F32B gamma, F32B bias, int32_t act, Out<Any> y,
                      Out<Any> xa, Out<Any> z) {
  ed2_be_fwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(w, 1); a.ensemble_size = dim(alpha, 0); a.act = act;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.w_lp = w.untyped_data(); a.w_ld = a.out_dim;
  a.alpha = alpha.typed_data(); a.gamma = gamma.typed_data(); a.bias = bias.typed_data();
  a.y = y->untyped_data(); a.y_ld = a.out_dim;
  a.xa = xa->untyped_data(); a.xa_ld = a.in_dim;
  a.z = z->untyped_data(); a.z_ld = a.out_dim;
  return done(ed2_be_dense_fwd(&a, s));
}
ffi::Error BeDenseBwd(cudaStream_t s, Any gy, Any y, Any z, Any x, Any xa, Any w, F32B alpha, F32B gamma, int32_t act,
                      Out<Any> dx, Out<F32B> dw, Out<F32B> dalpha, Out<F32B> dgamma, Out<F32B> dbias, Out<Any> dz_ws,
                      Out<Any> dxa_ws) {
  ed2_be_bwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(w, 1); a.ensemble_size = dim(alpha, 0); a.act = act;
  a.gy = gy.untyped_data(); a.gy_ld = a.out_dim;
  a.y = y.untyped_data(); a.y_ld = a.out_dim;
  a.z = z.untyped_data(); a.z_ld = a.out_dim;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.xa = xa.untyped_data(); a.xa_ld = a.in_dim;
  a.w_lp = w.untyped_data(); a.w_ld = a.out_dim;
  a.alpha = alpha.typed_data(); a.gamma = gamma.typed_data();
  a.dz = dz_ws->untyped_data(); a.dz_ld = a.out_dim;
  a.dxa = dxa_ws->untyped_data(); a.dxa_ld = a.in_dim;
  a.dx = dx->untyped_data(); a.dx_ld = a.in_dim;
  a.dw = dw->typed_data(); a.dalpha = dalpha->typed_data(); a.dgamma = dgamma->typed_data(); a.dbias = dbias->typed_data();
  return done(ed2_be_dense_bwd(&a, s));
}

// ---------------------------------------------------------------------------------------------- DenseRank1
// unfused form (every saved tensor is an XLA buffer); the fused epilogue chain needs caller-managed buffers shared
// between two layer calls and is driven through the C ABI directly
ffi::Error Rank1DenseFwd(cudaStream_t s, Any x, Any w, F32B mu_r, F32B rho_r, F32B mu_s, F32B rho_s, F32B bias, int32_t act,
                         int32_t additive, float clip_lo, float clip_hi, int64_t seed, int64_t row0, int32_t rpg_local,
                         int32_t rpg_global, Out<Any> y, Out<Any> xr, Out<Any> s_buf, Out<Any> z) {
  ed2_rank1_fwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(w, 1); a.ensemble_size = dim(mu_r, 0); a.act = act;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.w_lp = w.untyped_data(); a.w_ld = a.out_dim;
  a.mu_r = mu_r.typed_data(); a.rho_r = rho_r.typed_data(); a.eps_r = philox(seed, 4, row0, rpg_local, rpg_global);
  a.mu_s = mu_s.typed_data(); a.rho_s = rho_s.typed_data(); a.eps_s = philox(seed, 5, row0, rpg_local, rpg_global);
  a.bias = bias.typed_data();
  a.y = y->untyped_data(); a.y_ld = a.out_dim;
  a.xr = xr->untyped_data(); a.xr_ld = a.in_dim;
  a.s_buf = s_buf->untyped_data(); a.s_ld = a.out_dim;
  a.z = z->untyped_data(); a.z_ld = a.out_dim;
  a.additive = additive; a.clip_lo = clip_lo; a.clip_hi = clip_hi;
  return done(ed2_rank1_dense_fwd(&a, s));
}
ffi::Error Rank1DenseBwd(cudaStream_t s, Any gy, Any y, Any z, Any x, Any xr, Any s_buf, Any w, F32B mu_r, F32B rho_r,
                         F32B mu_s, F32B rho_s, int32_t act, int32_t additive, float clip_lo, float clip_hi, int64_t seed,
                         int64_t row0, int32_t rpg_local, int32_t rpg_global, Out<Any> dx, Out<F32B> dw, Out<F32B> dmu_r,
                         Out<F32B> drho_r, Out<F32B> dmu_s, Out<F32B> drho_s, Out<F32B> dbias, Out<Any> dz_ws,
                         Out<Any> dxr_ws) {
  ed2_rank1_bwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(w, 1); a.ensemble_size = dim(mu_r, 0); a.act = act;
  a.gy = gy.untyped_data(); a.gy_ld = a.out_dim;
  a.y = y.untyped_data(); a.y_ld = a.out_dim;
  a.z = z.untyped_data(); a.z_ld = a.out_dim;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.xr = xr.untyped_data(); a.xr_ld = a.in_dim;
  a.s_buf = s_buf.untyped_data(); a.s_ld = a.out_dim;
  a.w_lp = w.untyped_data(); a.w_ld = a.out_dim;
  a.mu_r = mu_r.typed_data(); a.rho_r = rho_r.typed_data(); a.eps_r = philox(seed, 4, row0, rpg_local, rpg_global);
  a.mu_s = mu_s.typed_data(); a.rho_s = rho_s.typed_data(); a.eps_s = philox(seed, 5, row0, rpg_local, rpg_global);
  a.dz = dz_ws->untyped_data(); a.dz_ld = a.out_dim;
  a.dxr = dxr_ws->untyped_data(); a.dxr_ld = a.in_dim;
  a.dx = dx->untyped_data(); a.dx_ld = a.in_dim;
  a.dw = dw->typed_data();
  a.dmu_r = dmu_r->typed_data(); a.drho_r = drho_r->typed_data();
  a.dmu_s = dmu_s->typed_data(); a.drho_s = drho_s->typed_data(); a.dbias = dbias->typed_data();
  a.additive = additive; a.clip_lo = clip_lo; a.clip_hi = clip_hi;
  return done(ed2_rank1_dense_bwd(&a, s));
}

// ---------------------------------------------------------------------------------------------- DenseReparameterization
ffi::Error ReparamDenseFwd(cudaStream_t s, Any x, F32B mu, F32B rho, F32B bias, int32_t act, int64_t seed, float kl_scale,
                           float prior_mean, float prior_std, Out<Any> y, Out<Any> w_lp, Out<F64B> kl /* zero-initialised by the caller */) {
  ed2_reparam_fwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(mu, 1); a.act = act;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.mu = mu.typed_data(); a.rho = rho.typed_data(); a.eps = philox(seed, 0);
  a.bias = bias.typed_data();
  a.w_lp = w_lp->untyped_data(); a.w_ld = a.out_dim;
  a.y = y->untyped_data(); a.y_ld = a.out_dim;
  a.kl_out = kl->typed_data(); a.kl_scale = kl_scale; a.prior_mean = prior_mean; a.prior_std = prior_std;
  return done(ed2_reparam_dense_fwd(&a, s));
}
ffi::Error ReparamDenseBwd(cudaStream_t s, Any gy, Any y, Any x, Any w_lp, F32B mu, F32B rho, F64B kl_upstream, int32_t act,
                           int64_t seed, float kl_grad_scale, float prior_mean, float prior_std, Out<Any> dx, Out<F32B> dmu,
                           Out<F32B> drho, Out<F32B> dbias, Out<Any> gp_ws) {
  ed2_reparam_bwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(mu, 1); a.act = act;
  a.gy = gy.untyped_data(); a.gy_ld = a.out_dim;
  a.y = y.untyped_data(); a.y_ld = a.out_dim;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.w_lp = w_lp.untyped_data(); a.w_ld = a.out_dim;
  a.mu = mu.typed_data(); a.rho = rho.typed_data(); a.eps = philox(seed, 0);
  a.gp = gp_ws->untyped_data(); a.gp_ld = a.out_dim;
  a.dx = dx->untyped_data(); a.dx_ld = a.in_dim;
  a.dmu = dmu->typed_data(); a.drho = drho->typed_data(); a.dbias = dbias->typed_data();
  a.kl_grad_scale = kl_grad_scale; a.prior_mean = prior_mean; a.prior_std = prior_std;
  a.kl_upstream = kl_upstream.typed_data();
  return done(ed2_reparam_dense_bwd(&a, s));
}

// ---------------------------------------------------------------------------------------------- DenseFlipout
ffi::Error FlipoutDenseFwd(cudaStream_t s, Any x, F32B mu, F32B rho, F32B bias, int32_t act, int64_t seed, int64_t row0,
                           float kl_scale, float prior_mean, float prior_std, Out<Any> y, Out<Any> mu_lp, Out<Any> delta_lp,
                           Out<Any> xs, Out<F64B> kl, Out<F32B> pert_ws) {
  ed2_flipout_fwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(mu, 1); a.act = act;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.mu = mu.typed_data(); a.rho = rho.typed_data(); a.eps = philox(seed, 0);
  a.sign_in = philox(seed, 2, row0); a.sign_out = philox(seed, 3, row0);
  a.bias = bias.typed_data();
  a.mu_lp = mu_lp->untyped_data(); a.mu_ld = a.out_dim;
  a.delta_lp = delta_lp->untyped_data(); a.delta_ld = a.out_dim;
  a.xs = xs->untyped_data(); a.xs_ld = a.in_dim;
  a.pert = pert_ws->typed_data(); a.pert_ld = a.out_dim;
  a.y = y->untyped_data(); a.y_ld = a.out_dim;
  a.kl_out = kl->typed_data(); a.kl_scale = kl_scale; a.prior_mean = prior_mean; a.prior_std = prior_std;
  return done(ed2_flipout_dense_fwd(&a, s));
}
ffi::Error FlipoutDenseBwd(cudaStream_t s, Any gy, Any y, Any x, Any xs, Any mu_lp, Any delta_lp, F32B mu, F32B rho,
                           F64B kl_upstream, int32_t act, int64_t seed, int64_t row0, float kl_grad_scale, float prior_mean,
                           float prior_std, Out<Any> dx, Out<F32B> dmu, Out<F32B> drho, Out<F32B> dbias, Out<Any> gp_ws,
                           Out<Any> gt_ws, Out<F32B> tmp_ws) {
  ed2_flipout_bwd_args a{};
  a.dtype = dt_of(x);
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(mu, 1); a.act = act;
  a.gy = gy.untyped_data(); a.gy_ld = a.out_dim;
  a.y = y.untyped_data(); a.y_ld = a.out_dim;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.xs = xs.untyped_data(); a.xs_ld = a.in_dim;
  a.mu_lp = mu_lp.untyped_data(); a.mu_ld = a.out_dim;
  a.delta_lp = delta_lp.untyped_data(); a.delta_ld = a.out_dim;
  a.mu = mu.typed_data(); a.rho = rho.typed_data(); a.eps = philox(seed, 0);
  a.sign_in = philox(seed, 2, row0); a.sign_out = philox(seed, 3, row0);
  a.gp = gp_ws->untyped_data(); a.gp_ld = a.out_dim;
  a.gt = gt_ws->untyped_data(); a.gt_ld = a.out_dim;
  a.tmp = tmp_ws->typed_data(); a.tmp_ld = a.in_dim;
  a.dx = dx->untyped_data(); a.dx_ld = a.in_dim;
  a.dmu = dmu->typed_data(); a.drho = drho->typed_data(); a.dbias = dbias->typed_data();
  a.kl_grad_scale = kl_grad_scale; a.prior_mean = prior_mean; a.prior_std = prior_std;
  a.kl_upstream = kl_upstream.typed_data();
  return done(ed2_flipout_dense_bwd(&a, s));
}

// Fused DenseFlipout: one dual-accumulator launch per product pair (mu_lp / delta_lp drawn by Ed2SampleNormalWeights-style
// calls beforehand; here they are inputs).  next_xs: the NEXT layer's x∘S (pass next_seed < 0 to skip).
ffi::Error FlipoutFusedFwd(cudaStream_t s, Any x, Any xs_in, Any mu_lp, Any delta_lp, F32B bias, int32_t act, int64_t seed,
                           int64_t next_seed, int64_t row0, int32_t xs_ready, Out<Any> y, Out<Any> xs, Out<Any> next_xs) {
  ed2_flipout_fused_fwd_args a{};
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(mu_lp, 1); a.act = act;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.xs = xs_ready ? const_cast<void*>(xs_in.untyped_data()) : xs->untyped_data(); a.xs_ld = a.in_dim; a.xs_ready = xs_ready;
  a.sign_in = philox(seed, 2, row0); a.sign_out = philox(seed, 3, row0);
  a.mu_lp = mu_lp.untyped_data(); a.mu_ld = a.out_dim;
  a.delta_lp = delta_lp.untyped_data(); a.delta_ld = a.out_dim;
  a.bias = bias.typed_data();
  a.y = y->untyped_data(); a.y_ld = a.out_dim;
  if (next_seed >= 0) {
    a.next_xs = next_xs->untyped_data(); a.next_xs_ld = a.out_dim; a.next_sign_in = philox(next_seed, 2, row0);
  }
  return done(ed2_flipout_fused_fwd(&a, s));
}
ffi::Error FlipoutFusedBwd(cudaStream_t s, Any gy, Any y, Any x, Any xs, Any mu_lp, Any delta_lp, F32B mu, F32B rho,
                           F64B kl_upstream, int32_t act, int64_t seed, int64_t row0, float kl_grad_scale, float prior_mean,
                           float prior_std, Out<Any> dx, Out<F32B> dmu, Out<F32B> drho, Out<F32B> dbias, Out<Any> gp_ws,
                           Out<Any> gt_ws) {
  ed2_flipout_fused_bwd_args a{};
  a.batch = dim(x, 0); a.in_dim = dim(x, 1); a.out_dim = dim(mu, 1); a.act = act; a.premade = 0;
  a.gy = gy.untyped_data(); a.gy_ld = a.out_dim;
  a.y = y.untyped_data(); a.y_ld = a.out_dim;
  a.gp = gp_ws->untyped_data(); a.gp_ld = a.out_dim;
  a.gt = gt_ws->untyped_data(); a.gt_ld = a.out_dim;
  a.x = x.untyped_data(); a.x_ld = a.in_dim;
  a.xs = xs.untyped_data(); a.xs_ld = a.in_dim;
  a.mu_lp = mu_lp.untyped_data(); a.mu_ld = a.out_dim;
  a.delta_lp = delta_lp.untyped_data(); a.delta_ld = a.out_dim;
  a.mu = mu.typed_data(); a.rho = rho.typed_data(); a.eps = philox(seed, 0);
  a.sign_in = philox(seed, 2, row0); a.sign_out = philox(seed, 3, row0);
  a.dmu = dmu->typed_data(); a.drho = drho->typed_data(); a.dbias = dbias->typed_data();
  a.kl_grad_scale = kl_grad_scale; a.prior_mean = prior_mean; a.prior_std = prior_std;
  a.kl_upstream = kl_upstream.typed_data();
  a.dx = dx->untyped_data(); a.dx_ld = a.in_dim;
  return done(ed2_flipout_fused_bwd(&a, s));
}

// ---------------------------------------------------------------------------------------------- heteroscedastic heads
// Monte-Carlo softmax with factor-analysis noise — edward2/jax/nn/heteroscedastic_lib.py / tensorflow heteroscedastic.py:178-253
ffi::Error McSoftmaxFaFwd(cudaStream_t s, F32B loc, F32B fac, F32B diag_raw, int32_t num_samples, int32_t num_factors,
                          int32_t share_samples, float temperature, float eps, int64_t seed, int64_t row0, Out<F32B> probs_mean,
                          Out<F32B> logits, Out<F32B> variance) {
  return done(ed2_mc_softmax_fa_fwd(loc.typed_data(), fac.typed_data(), diag_raw.typed_data(), nullptr, 1.f, 1.f,
                                    philox(seed, 7, row0), dim(loc, 0), num_samples, dim(loc, 1), num_factors, share_samples,
                                    temperature, eps, probs_mean->typed_data(), logits->typed_data(), variance->typed_data(),
                                    s));
}
ffi::Error McSoftmaxFaBwd(cudaStream_t s, F32B g, F32B probs_mean, F32B loc, F32B fac, F32B diag_raw, int32_t num_samples,
                          int32_t num_factors, int32_t share_samples, float temperature, float eps, int64_t seed, int64_t row0,
                          Out<F32B> dloc, Out<F32B> dfac, Out<F32B> ddiag_raw) {
  return done(ed2_mc_softmax_fa_bwd(g.typed_data(), probs_mean.typed_data(), loc.typed_data(), fac.typed_data(),
                                    diag_raw.typed_data(), philox(seed, 7, row0), dim(loc, 0), num_samples, dim(loc, 1),
                                    num_factors, share_samples, temperature, eps, dloc->typed_data(), dfac->typed_data(),
                                    ddiag_raw->typed_data(), s));
}

// ---------------------------------------------------------------------------------------------- factorised Conv2D
// y[b] = act(conv(x[b] ∘ f_in[b], W) ∘ f_out[b] + bias[b]) — Conv2DBatchEnsemble / Conv2DRank1 with per-example tables
inline ed2_conv_geometry geometry(const Any& x, int32_t kh, int32_t kw, int32_t sh, int32_t sw, int32_t same) {
  ed2_conv_geometry g{};
  g.batch = dim(x, 0); g.height = dim(x, 1); g.width = dim(x, 2); g.channels = dim(x, 3);
  g.kh = kh; g.kw = kw; g.sh = sh; g.sw = sw; g.same = same;
  return g;
}
ffi::Error ConvFactorFwd(cudaStream_t s, Any x, Any w, F32B f_in, F32B f_out, F32B bias, int32_t kh, int32_t kw, int32_t sh,
                         int32_t sw, int32_t same, int32_t act, Out<Any> y, Out<Any> xa, Out<Any> z) {
  ed2_conv_factor_fwd_args a{};
  a.geom = geometry(x, kh, kw, sh, sw, same);
  a.dtype = dt_of(x); a.out_channels = dim(w, 1); a.act = act;
  a.x = x.untyped_data(); a.w_lp = w.untyped_data(); a.w_ld = a.out_channels;
  a.f_in = f_in.typed_data(); a.f_out = f_out.typed_data(); a.bias = bias.typed_data();
  a.xa = xa->untyped_data(); a.xa_ld = kh * kw * a.geom.channels;
  a.y = y->untyped_data(); a.y_ld = a.out_channels; a.z = z->untyped_data(); a.z_ld = a.out_channels;
  return done(ed2_conv_factor_fwd(&a, s));
}
ffi::Error ConvFactorBwd(cudaStream_t s, Any gy, Any y, Any z, Any x, Any xa, Any w, F32B f_in, F32B f_out, int32_t kh,
                         int32_t kw, int32_t sh, int32_t sw, int32_t same, int32_t act, Out<Any> dx, Out<F32B> dw,
                         Out<F32B> df_in, Out<F32B> df_out, Out<F32B> dbias, Out<Any> dz_ws, Out<Any> dxa_ws, Out<Any> gimg_ws) {
  ed2_conv_factor_bwd_args a{};
  a.geom = geometry(x, kh, kw, sh, sw, same);
  a.dtype = dt_of(x); a.out_channels = dim(w, 1); a.act = act;
  a.gy = gy.untyped_data(); a.gy_ld = a.out_channels; a.y = y.untyped_data(); a.y_ld = a.out_channels;
  a.z = z.untyped_data(); a.z_ld = a.out_channels;
  a.x = x.untyped_data(); a.xa = xa.untyped_data(); a.xa_ld = kh * kw * a.geom.channels;
  a.w_lp = w.untyped_data(); a.w_ld = a.out_channels;
  a.f_in = f_in.typed_data(); a.f_out = f_out.typed_data();
  a.dz = dz_ws->untyped_data(); a.dz_ld = a.out_channels;
  a.dxa = dxa_ws->untyped_data(); a.dxa_ld = a.xa_ld; a.gimg = gimg_ws->untyped_data();
  a.dx = dx->untyped_data(); a.dw = dw->typed_data();
  a.df_in = df_in->typed_data(); a.df_out = df_out->typed_data(); a.dbias = dbias->typed_data();
  return done(ed2_conv_factor_bwd(&a, s));
}

// ---------------------------------------------------------------------------------------------- NormalKLDivergence
ffi::Error NormalKlFwd(cudaStream_t s, F32B mu, F32B rho, float prior_mean, float prior_std, float scale, Out<F64B> kl) {
  return done(ed2_normal_kl_fwd(mu.typed_data(), rho.typed_data(), static_cast<long long>(mu.element_count()), prior_mean,
                                prior_std, scale, kl->typed_data(), s));
}
ffi::Error NormalKlBwd(cudaStream_t s, F32B mu, F32B rho, F32B upstream, float prior_mean, float prior_std, float scale,
                       Out<F32B> dmu /* zero-initialised */, Out<F32B> drho /* zero-initialised */) {
  return done(ed2_normal_kl_bwd(mu.typed_data(), rho.typed_data(), static_cast<long long>(mu.element_count()), prior_mean,
                                prior_std, scale, upstream.typed_data(), dmu->typed_data(), drho->typed_data(), s));
}

// ---------------------------------------------------------------------------------------------- SNGP head
// phi = cos(x W + b) * scale (edward2/jax/nn/random_feature.py:209-214); wt = W^T [D][d]
ffi::Error RffFwd(cudaStream_t s, Any x, Any wt, F32B bias, float scale, Out<Any> phi, Out<F32B> pre) {
  return done(ed2_rff_fwd(dt_of(x), x.untyped_data(), dim(x, 1), wt.untyped_data(), dim(wt, 1), bias.typed_data(), scale,
                          dim(x, 0), dim(x, 1), dim(wt, 0), phi->untyped_data(), dim(wt, 0), dt_of(*phi), pre->typed_data(),
                          dim(wt, 0), /*pre_mode=*/0, s));
}
ffi::Error RffBwd(cudaStream_t s, Any dphi, F32B pre, F32B bias, Any w /* [d][D] */, float scale, Out<Any> dx,
                  Out<Any> gpre_ws) {
  const int B = dim(dphi, 0), D = dim(dphi, 1), d = dim(w, 0);
  return done(ed2_rff_bwd(dt_of(w), dphi.untyped_data(), D, dt_of(dphi), pre.typed_data(), D, /*pre_mode=*/0,
                          bias.typed_data(), w.untyped_data(), D, scale, B, d, D, gpre_ws->untyped_data(), D,
                          dx->untyped_data(), d, dt_of(*dx), s));
}
// LayerNorm over the last axis (flax nn.LayerNorm(use_bias=False, use_scale=False), epsilon 1e-6:
// edward2/jax/nn/random_feature.py:98-102) and its backward (the reference relies on autodiff)
ffi::Error LayerNormFwd(cudaStream_t s, Any x, float epsilon, Out<Any> y) {
  const int cols = dim(x, 1);
  return done(ed2_input_normalize(x.untyped_data(), dt_of(x), cols, y->untyped_data(), dt_of(*y), cols, dim(x, 0), cols, nullptr,
                                  nullptr, epsilon, /*scale_only=*/0, 1.f, s));
}
ffi::Error LayerNormBwd(cudaStream_t s, Any x, Any dy, float epsilon, Out<Any> dx) {
  const int cols = dim(x, 1);
  return done(ed2_layernorm_bwd(x.untyped_data(), dt_of(x), cols, dy.untyped_data(), dt_of(dy), cols, dim(x, 0), cols, nullptr,
                                epsilon, dx->untyped_data(), dt_of(*dx), cols, nullptr, nullptr, s));
}
// precision <- momentum * precision + (1 - momentum) * phi^T phi / batch_total, or precision + phi^T phi
// (jax/nn/random_feature.py:350-362); `precision` is updated IN PLACE (input_output_aliases={1: 0} on the Python side)
ffi::Error LaplacePrecisionUpdate(cudaStream_t s, Any phi, F32B precision_in, float momentum, int64_t batch_total,
                                  Out<F32B> precision) {
  (void)precision_in;  // aliased with the result
  return done(ed2_laplace_precision_update(dt_of(phi), phi.untyped_data(), dim(phi, 1), dim(phi, 0), dim(phi, 1), momentum,
                                           batch_total, precision->typed_data(), nullptr, s));
}
// var = diag(ridge * phi Sigma phi^T) with the mean-field logits fused (jax/nn/random_feature.py:393-404; utils.py:54-101)
ffi::Error PredictiveVariance(cudaStream_t s, Any phi, Any sigma_lp, F32B logits, float ridge, float mean_field_factor,
                              int32_t likelihood, Out<F32B> var /* zero-initialised */, Out<F32B> logits_out) {
  return done(ed2_predictive_variance(dt_of(phi), phi.untyped_data(), dim(phi, 1), sigma_lp.untyped_data(), dim(sigma_lp, 1),
                                      dim(phi, 0), dim(phi, 1), ridge, var->typed_data(), logits.typed_data(), dim(logits, 1),
                                      mean_field_factor, likelihood, logits_out->typed_data(), s));
}
// Sigma = (ridge I + P)^-1 on the GPU (jax/nn/random_feature.py:393-395 factorises and solves on every evaluation call);
// `a` is inverted in place (input_output_aliases={0: 0}); workspace of ed2_spd_inverse_workspace_floats(n) floats
ffi::Error SpdInverse(cudaStream_t s, F32B a_in, Out<F32B> a, Out<F32B> workspace) {
  (void)a_in;
  const int n = dim(*a, 0);
  if (static_cast<long long>(workspace->element_count()) < ed2_spd_inverse_workspace_floats(n))
    return ffi::Error::InvalidArgument("Ed2SpdInverse: workspace too small");
  return done(ed2_spd_inverse(a->typed_data(), n, workspace->typed_data(), s));
}
ffi::Error MeanFieldLogits(cudaStream_t s, F32B logits, F32B var, float mean_field_factor, int32_t likelihood,
                           Out<F32B> out) {
  return done(ed2_mean_field_logits(logits.typed_data(), var.typed_data(), dim(logits, 0), dim(logits, 1), mean_field_factor,
                                    likelihood, out->typed_data(), s));
}
// power iteration + sigma + w / max(sigma / c, 1) (mode 1, jax/nn/normalization.py:159-178) or the TF in-place form
// (mode 0); u, v are updated in place (aliased results)
ffi::Error SpectralNormUpdate(cudaStream_t s, F32B w, F32B u_in, F32B v_in, int32_t iterations, float norm_multiplier,
                              int32_t mode, int32_t update_uv, Out<F32B> u, Out<F32B> v, Out<F32B> w_out, Out<F32B> sigma,
                              Out<F32B> workspace /* in_dim + out_dim + 8 floats */) {
  (void)u_in; (void)v_in;
  return done(ed2_spectral_norm_update(w.typed_data(), dim(w, 0), dim(w, 1), u->typed_data(), v->typed_data(), iterations,
                                       norm_multiplier, mode, update_uv, w_out->typed_data(), nullptr, 0,
                                       sigma->typed_data(), workspace->typed_data(), s));
}

}  // namespace

#define ED2_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2BeDenseFwd, BeDenseFwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("act")
        .Ret<Any>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2BeDenseBwd, BeDenseBwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("act")
        .Ret<Any>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2Rank1DenseFwd, Rank1DenseFwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>()
        .Attr<int32_t>("act").Attr<int32_t>("additive").Attr<float>("clip_lo").Attr<float>("clip_hi").Attr<int64_t>("seed")
        .Attr<int64_t>("row0").Attr<int32_t>("rows_per_group_local").Attr<int32_t>("rows_per_group_global")
        .Ret<Any>().Ret<Any>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2Rank1DenseBwd, Rank1DenseBwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>()
        .Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>()
        .Attr<int32_t>("act").Attr<int32_t>("additive").Attr<float>("clip_lo").Attr<float>("clip_hi").Attr<int64_t>("seed")
        .Attr<int64_t>("row0").Attr<int32_t>("rows_per_group_local").Attr<int32_t>("rows_per_group_global")
        .Ret<Any>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2ReparamDenseFwd, ReparamDenseFwd,
    ED2_STREAM.Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("act").Attr<int64_t>("seed")
        .Attr<float>("kl_scale").Attr<float>("prior_mean").Attr<float>("prior_std").Ret<Any>().Ret<Any>().Ret<F64B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2ReparamDenseBwd, ReparamDenseBwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F64B>().Attr<int32_t>("act")
        .Attr<int64_t>("seed").Attr<float>("kl_grad_scale").Attr<float>("prior_mean").Attr<float>("prior_std")
        .Ret<Any>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2FlipoutDenseFwd, FlipoutDenseFwd,
    ED2_STREAM.Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("act").Attr<int64_t>("seed").Attr<int64_t>("row0")
        .Attr<float>("kl_scale").Attr<float>("prior_mean").Attr<float>("prior_std")
        .Ret<Any>().Ret<Any>().Ret<Any>().Ret<Any>().Ret<F64B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2FlipoutDenseBwd, FlipoutDenseBwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F64B>()
        .Attr<int32_t>("act").Attr<int64_t>("seed").Attr<int64_t>("row0").Attr<float>("kl_grad_scale")
        .Attr<float>("prior_mean").Attr<float>("prior_std")
        .Ret<Any>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<Any>().Ret<Any>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2FlipoutFusedFwd, FlipoutFusedFwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<F32B>().Attr<int32_t>("act").Attr<int64_t>("seed")
        .Attr<int64_t>("next_seed").Attr<int64_t>("row0").Attr<int32_t>("xs_ready").Ret<Any>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2FlipoutFusedBwd, FlipoutFusedBwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F64B>()
        .Attr<int32_t>("act").Attr<int64_t>("seed").Attr<int64_t>("row0").Attr<float>("kl_grad_scale")
        .Attr<float>("prior_mean").Attr<float>("prior_std")
        .Ret<Any>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2McSoftmaxFaFwd, McSoftmaxFaFwd,
    ED2_STREAM.Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("num_samples").Attr<int32_t>("num_factors")
        .Attr<int32_t>("share_samples").Attr<float>("temperature").Attr<float>("eps").Attr<int64_t>("seed")
        .Attr<int64_t>("row0").Ret<F32B>().Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2McSoftmaxFaBwd, McSoftmaxFaBwd,
    ED2_STREAM.Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("num_samples")
        .Attr<int32_t>("num_factors").Attr<int32_t>("share_samples").Attr<float>("temperature").Attr<float>("eps")
        .Attr<int64_t>("seed").Attr<int64_t>("row0").Ret<F32B>().Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2ConvFactorFwd, ConvFactorFwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("kh").Attr<int32_t>("kw")
        .Attr<int32_t>("sh").Attr<int32_t>("sw").Attr<int32_t>("same").Attr<int32_t>("act").Ret<Any>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2ConvFactorBwd, ConvFactorBwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<Any>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("kh")
        .Attr<int32_t>("kw").Attr<int32_t>("sh").Attr<int32_t>("sw").Attr<int32_t>("same").Attr<int32_t>("act")
        .Ret<Any>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<Any>().Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2NormalKlFwd, NormalKlFwd,
    ED2_STREAM.Arg<F32B>().Arg<F32B>().Attr<float>("prior_mean").Attr<float>("prior_std").Attr<float>("scale").Ret<F64B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2NormalKlBwd, NormalKlBwd,
    ED2_STREAM.Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<float>("prior_mean").Attr<float>("prior_std").Attr<float>("scale")
        .Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2RffFwd, RffFwd,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<F32B>().Attr<float>("scale").Ret<Any>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2RffBwd, RffBwd,
    ED2_STREAM.Arg<Any>().Arg<F32B>().Arg<F32B>().Arg<Any>().Attr<float>("scale").Ret<Any>().Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2LayerNormFwd, LayerNormFwd, ED2_STREAM.Arg<Any>().Attr<float>("epsilon").Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2LayerNormBwd, LayerNormBwd, ED2_STREAM.Arg<Any>().Arg<Any>().Attr<float>("epsilon").Ret<Any>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2LaplacePrecisionUpdate, LaplacePrecisionUpdate,
    ED2_STREAM.Arg<Any>().Arg<F32B>().Attr<float>("momentum").Attr<int64_t>("batch_total").Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2PredictiveVariance, PredictiveVariance,
    ED2_STREAM.Arg<Any>().Arg<Any>().Arg<F32B>().Attr<float>("ridge").Attr<float>("mean_field_factor")
        .Attr<int32_t>("likelihood").Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2SpdInverse, SpdInverse, ED2_STREAM.Arg<F32B>().Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2MeanFieldLogits, MeanFieldLogits,
    ED2_STREAM.Arg<F32B>().Arg<F32B>().Attr<float>("mean_field_factor").Attr<int32_t>("likelihood").Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(Ed2SpectralNormUpdate, SpectralNormUpdate,
    ED2_STREAM.Arg<F32B>().Arg<F32B>().Arg<F32B>().Attr<int32_t>("iterations").Attr<float>("norm_multiplier")
        .Attr<int32_t>("mode").Attr<int32_t>("update_uv").Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>());

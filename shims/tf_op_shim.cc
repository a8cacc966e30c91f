// TensorFlow custom-op shim for libed2b200: forward AND backward ops of the four dense layers, the KL regulariser, the
// factorised Conv2D variants, the LSTM gate kernels, the Monte-Carlo heads and the SNGP-head operators (incl. the
// covariance inverse), each a thin OpKernel that fills the argument struct of include/ed2b200.h from its inputs / outputs and
// calls the C entry point on TensorFlow's GPU stream.  The Python side (tf.load_op_library + tf.custom_gradient) is in
// INTEGRATION.md §3.
//
// The TensorFlow headers are not in this image, so the shim cannot be BUILT here; it is TYPE-CHECKED against
// shims/mock/tensorflow/... (tests/test_host.py::test_shims_type_check).  Build where TensorFlow is installed:
//   g++ -shared -fPIC -std=c++17 $(python -c "import tensorflow as tf;print(' '.join(tf.sysconfig.get_compile_flags()))") \
//       -I../include tf_op_shim.cc -L../edward2_b200 -led2b200 \
//       $(python -c "import tensorflow as tf;print(' '.join(tf.sysconfig.get_link_flags()))") -o libed2b200_tf.so
// Conventions as in jax_ffi_shim.cc: dense row-major tensors (leading dimension = last dimension, multiples of 8 for
// bfloat16), workspaces are temporaries, noise is Philox (seed, stream) attributes.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "ed2b200.h"

using namespace tensorflow;  // NOLINT

namespace {

template <typename T> constexpr int Dt() { return std::is_same<T, float>::value ? ED2_F32 : ED2_BF16; }
inline ed2_noise Philox(int64_t seed, int64_t stream, int64_t row0 = 0, int rpg_local = 0, int rpg_global = 0) {
  ed2_noise n{};
  n.seed = static_cast<uint64_t>(seed); n.stream = static_cast<uint64_t>(stream);
  n.row0 = row0; n.rows_per_group_local = rpg_local; n.rows_per_group_global = rpg_global;
  return n;
}
inline const float* F(const Tensor& t) { return t.flat<float>().data(); }
inline float* F(Tensor* t) { return t->flat<float>().data(); }
inline int D(const Tensor& t, int i) { return static_cast<int>(t.dim_size(i)); }

#define ED2_OUT(i, shape, name) Tensor* name = nullptr; OP_REQUIRES_OK(c, c->allocate_output(i, shape, &name))
#define ED2_TMP(shape, name) Tensor name; OP_REQUIRES_OK(c, c->allocate_temp(DataTypeToEnum<T>::value, shape, &name))
#define ED2_CALL(expr) OP_REQUIRES(c, (expr) == ED2_OK, errors::Internal(ed2_last_error()))
#define ED2_REGISTER(op, cls)                                                                              \
  REGISTER_KERNEL_BUILDER(Name(op).Device(DEVICE_GPU).TypeConstraint<float>("T"), cls<float>);             \
  REGISTER_KERNEL_BUILDER(Name(op).Device(DEVICE_GPU).TypeConstraint<bfloat16>("T"), cls<bfloat16>)

}  // namespace

// ================================================================================================ DenseBatchEnsemble
// edward2/tensorflow/layers/dense.py:551-584
REGISTER_OP("Ed2BeDenseFwd")
    .Input("x: T").Input("w: T").Input("alpha: float").Input("gamma: float").Input("bias: float")
    .Output("y: T").Output("xa: T").Output("z: T").Attr("T: {float, bfloat16}").Attr("act: int = 0");
template <typename T>
class Ed2BeDenseFwdOp : public OpKernel {
 public:
  explicit Ed2BeDenseFwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("act", &act_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1), &alpha = c->input(2), &gamma = c->input(3), &bias = c->input(4);
    const int B = D(x, 0), I = D(x, 1), O = D(w, 1);
    ED2_OUT(0, TensorShape({B, O}), y); ED2_OUT(1, TensorShape({B, I}), xa); ED2_OUT(2, TensorShape({B, O}), z);
    ed2_be_fwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.ensemble_size = D(alpha, 0); a.act = act_;
    a.x = x.data(); a.x_ld = I; a.w_lp = w.data(); a.w_ld = O;
    a.alpha = F(alpha); a.gamma = F(gamma); a.bias = F(bias);
    a.y = y->data(); a.y_ld = O; a.xa = xa->data(); a.xa_ld = I; a.z = z->data(); a.z_ld = O;
    ED2_CALL(ed2_be_dense_fwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  int act_;
};
ED2_REGISTER("Ed2BeDenseFwd", Ed2BeDenseFwdOp);

REGISTER_OP("Ed2BeDenseBwd")
    .Input("gy: T").Input("y: T").Input("z: T").Input("x: T").Input("xa: T").Input("w: T").Input("alpha: float").Input("gamma: float")
    .Output("dx: T").Output("dw: float").Output("dalpha: float").Output("dgamma: float").Output("dbias: float")
    .Attr("T: {float, bfloat16}").Attr("act: int = 0");
template <typename T>
class Ed2BeDenseBwdOp : public OpKernel {
 public:
  explicit Ed2BeDenseBwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("act", &act_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &gy = c->input(0), &y = c->input(1), &z = c->input(2), &x = c->input(3), &xa = c->input(4), &w = c->input(5),
                 &alpha = c->input(6), &gamma = c->input(7);
    const int B = D(x, 0), I = D(x, 1), O = D(w, 1), E = D(alpha, 0);
    ED2_OUT(0, TensorShape({B, I}), dx); ED2_OUT(1, TensorShape({I, O}), dw); ED2_OUT(2, TensorShape({E, I}), dalpha);
    ED2_OUT(3, TensorShape({E, O}), dgamma); ED2_OUT(4, TensorShape({E, O}), dbias);
    ED2_TMP(TensorShape({B, O}), dz); ED2_TMP(TensorShape({B, I}), dxa);
    ed2_be_bwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.ensemble_size = E; a.act = act_;
    a.gy = gy.data(); a.gy_ld = O; a.y = y.data(); a.y_ld = O; a.z = z.data(); a.z_ld = O;
    a.x = x.data(); a.x_ld = I; a.xa = xa.data(); a.xa_ld = I; a.w_lp = w.data(); a.w_ld = O;
    a.alpha = F(alpha); a.gamma = F(gamma);
    a.dz = dz.data(); a.dz_ld = O; a.dxa = dxa.data(); a.dxa_ld = I; a.dx = dx->data(); a.dx_ld = I;
    a.dw = F(dw); a.dalpha = F(dalpha); a.dgamma = F(dgamma); a.dbias = F(dbias);
    ED2_CALL(ed2_be_dense_bwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  int act_;
};
ED2_REGISTER("Ed2BeDenseBwd", Ed2BeDenseBwdOp);

// ================================================================================================ DenseRank1
// edward2/tensorflow/layers/dense.py:1096-1144 (unfused form: every saved tensor is a TF tensor)
#define ED2_RANK1_ATTRS                                                                                              \
  .Attr("T: {float, bfloat16}").Attr("act: int = 0").Attr("additive: int = 0").Attr("clip_lo: float = -10.0")        \
  .Attr("clip_hi: float = 10.0").Attr("seed: int = 0").Attr("row0: int = 0").Attr("rows_per_group_local: int = 0")   \
  .Attr("rows_per_group_global: int = 0")
struct Rank1Attrs {
  int act, additive, rpg_local, rpg_global;
  float clip_lo, clip_hi;
  int64_t seed, row0;
  void Read(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("act", &act)); OP_REQUIRES_OK(c, c->GetAttr("additive", &additive));
    OP_REQUIRES_OK(c, c->GetAttr("clip_lo", &clip_lo)); OP_REQUIRES_OK(c, c->GetAttr("clip_hi", &clip_hi));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed)); OP_REQUIRES_OK(c, c->GetAttr("row0", &row0));
    OP_REQUIRES_OK(c, c->GetAttr("rows_per_group_local", &rpg_local));
    OP_REQUIRES_OK(c, c->GetAttr("rows_per_group_global", &rpg_global));
  }
};
REGISTER_OP("Ed2Rank1DenseFwd")
    .Input("x: T").Input("w: T").Input("mu_r: float").Input("rho_r: float").Input("mu_s: float").Input("rho_s: float").Input("bias: float")
    .Output("y: T").Output("xr: T").Output("s: T").Output("z: T") ED2_RANK1_ATTRS;
template <typename T>
class Ed2Rank1DenseFwdOp : public OpKernel {
 public:
  explicit Ed2Rank1DenseFwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1), &mu_r = c->input(2), &rho_r = c->input(3), &mu_s = c->input(4),
                 &rho_s = c->input(5), &bias = c->input(6);
    const int B = D(x, 0), I = D(x, 1), O = D(w, 1);
    ED2_OUT(0, TensorShape({B, O}), y); ED2_OUT(1, TensorShape({B, I}), xr); ED2_OUT(2, TensorShape({B, O}), s);
    ED2_OUT(3, TensorShape({B, O}), z);
    ed2_rank1_fwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.ensemble_size = D(mu_r, 0); a.act = at_.act;
    a.x = x.data(); a.x_ld = I; a.w_lp = w.data(); a.w_ld = O;
    a.mu_r = F(mu_r); a.rho_r = F(rho_r); a.eps_r = Philox(at_.seed, 4, at_.row0, at_.rpg_local, at_.rpg_global);
    a.mu_s = F(mu_s); a.rho_s = F(rho_s); a.eps_s = Philox(at_.seed, 5, at_.row0, at_.rpg_local, at_.rpg_global);
    a.bias = F(bias);
    a.y = y->data(); a.y_ld = O; a.xr = xr->data(); a.xr_ld = I; a.s_buf = s->data(); a.s_ld = O; a.z = z->data(); a.z_ld = O;
    a.additive = at_.additive; a.clip_lo = at_.clip_lo; a.clip_hi = at_.clip_hi;
    ED2_CALL(ed2_rank1_dense_fwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  Rank1Attrs at_;
};
ED2_REGISTER("Ed2Rank1DenseFwd", Ed2Rank1DenseFwdOp);

REGISTER_OP("Ed2Rank1DenseBwd")
    .Input("gy: T").Input("y: T").Input("z: T").Input("x: T").Input("xr: T").Input("s: T").Input("w: T")
    .Input("mu_r: float").Input("rho_r: float").Input("mu_s: float").Input("rho_s: float")
    .Output("dx: T").Output("dw: float").Output("dmu_r: float").Output("drho_r: float").Output("dmu_s: float")
    .Output("drho_s: float").Output("dbias: float") ED2_RANK1_ATTRS;
template <typename T>
class Ed2Rank1DenseBwdOp : public OpKernel {
 public:
  explicit Ed2Rank1DenseBwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &gy = c->input(0), &y = c->input(1), &z = c->input(2), &x = c->input(3), &xr = c->input(4), &s = c->input(5),
                 &w = c->input(6), &mu_r = c->input(7), &rho_r = c->input(8), &mu_s = c->input(9), &rho_s = c->input(10);
    const int B = D(x, 0), I = D(x, 1), O = D(w, 1), E = D(mu_r, 0);
    ED2_OUT(0, TensorShape({B, I}), dx); ED2_OUT(1, TensorShape({I, O}), dw); ED2_OUT(2, TensorShape({E, I}), dmu_r);
    ED2_OUT(3, TensorShape({E, I}), drho_r); ED2_OUT(4, TensorShape({E, O}), dmu_s); ED2_OUT(5, TensorShape({E, O}), drho_s);
    ED2_OUT(6, TensorShape({E, O}), dbias);
    ED2_TMP(TensorShape({B, O}), dz); ED2_TMP(TensorShape({B, I}), dxr);
    ed2_rank1_bwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.ensemble_size = E; a.act = at_.act;
    a.gy = gy.data(); a.gy_ld = O; a.y = y.data(); a.y_ld = O; a.z = z.data(); a.z_ld = O;
    a.x = x.data(); a.x_ld = I; a.xr = xr.data(); a.xr_ld = I; a.s_buf = s.data(); a.s_ld = O; a.w_lp = w.data(); a.w_ld = O;
    a.mu_r = F(mu_r); a.rho_r = F(rho_r); a.eps_r = Philox(at_.seed, 4, at_.row0, at_.rpg_local, at_.rpg_global);
    a.mu_s = F(mu_s); a.rho_s = F(rho_s); a.eps_s = Philox(at_.seed, 5, at_.row0, at_.rpg_local, at_.rpg_global);
    a.dz = dz.data(); a.dz_ld = O; a.dxr = dxr.data(); a.dxr_ld = I; a.dx = dx->data(); a.dx_ld = I;
    a.dw = F(dw); a.dmu_r = F(dmu_r); a.drho_r = F(drho_r); a.dmu_s = F(dmu_s); a.drho_s = F(drho_s); a.dbias = F(dbias);
    a.additive = at_.additive; a.clip_lo = at_.clip_lo; a.clip_hi = at_.clip_hi;
    ED2_CALL(ed2_rank1_dense_bwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  Rank1Attrs at_;
};
ED2_REGISTER("Ed2Rank1DenseBwd", Ed2Rank1DenseBwdOp);

// ================================================================================================ DenseReparameterization
// edward2/tensorflow/layers/dense.py:73-83
struct KlAttrs {
  int act; int64_t seed, row0; float kl_scale, pm, ps;
  void Read(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("act", &act)); OP_REQUIRES_OK(c, c->GetAttr("seed", &seed));
    OP_REQUIRES_OK(c, c->GetAttr("row0", &row0)); OP_REQUIRES_OK(c, c->GetAttr("kl_scale", &kl_scale));
    OP_REQUIRES_OK(c, c->GetAttr("prior_mean", &pm)); OP_REQUIRES_OK(c, c->GetAttr("prior_std", &ps));
  }
};
#define ED2_KL_ATTRS                                                                                             \
  .Attr("T: {float, bfloat16}").Attr("act: int = 0").Attr("seed: int = 0").Attr("row0: int = 0")                  \
  .Attr("kl_scale: float = 1.0").Attr("prior_mean: float = 0.0").Attr("prior_std: float = 1.0")
REGISTER_OP("Ed2ReparamDenseFwd")
    .Input("x: T").Input("mu: float").Input("rho: float").Input("bias: float")
    .Output("y: T").Output("w_lp: T").Output("kl: double") ED2_KL_ATTRS
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(0), 0), c->Dim(c->input(1), 1)));
      c->set_output(1, c->input(1));
      c->set_output(2, c->Scalar());
      return OkStatus();
    });
template <typename T>
class Ed2ReparamDenseFwdOp : public OpKernel {
 public:
  explicit Ed2ReparamDenseFwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &mu = c->input(1), &rho = c->input(2), &bias = c->input(3);
    const int B = D(x, 0), I = D(x, 1), O = D(mu, 1);
    ED2_OUT(0, TensorShape({B, O}), y); ED2_OUT(1, TensorShape({I, O}), w); ED2_OUT(2, TensorShape({}), kl);
    auto stream = c->eigen_gpu_device().stream();
    cudaMemsetAsync(kl->flat<double>().data(), 0, sizeof(double), stream);
    ed2_reparam_fwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.act = at_.act;
    a.x = x.data(); a.x_ld = I;
    a.mu = F(mu); a.rho = F(rho); a.eps = Philox(at_.seed, 0); a.bias = F(bias);
    a.w_lp = w->data(); a.w_ld = O; a.y = y->data(); a.y_ld = O;
    a.kl_out = kl->flat<double>().data(); a.kl_scale = at_.kl_scale; a.prior_mean = at_.pm; a.prior_std = at_.ps;
    ED2_CALL(ed2_reparam_dense_fwd(&a, stream));
  }
 private:
  KlAttrs at_;
};
ED2_REGISTER("Ed2ReparamDenseFwd", Ed2ReparamDenseFwdOp);

REGISTER_OP("Ed2ReparamDenseBwd")
    .Input("gy: T").Input("y: T").Input("x: T").Input("w_lp: T").Input("mu: float").Input("rho: float").Input("kl_upstream: double")
    .Output("dx: T").Output("dmu: float").Output("drho: float").Output("dbias: float") ED2_KL_ATTRS;
template <typename T>
class Ed2ReparamDenseBwdOp : public OpKernel {
 public:
  explicit Ed2ReparamDenseBwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &gy = c->input(0), &y = c->input(1), &x = c->input(2), &w = c->input(3), &mu = c->input(4), &rho = c->input(5),
                 &up = c->input(6);
    const int B = D(x, 0), I = D(x, 1), O = D(mu, 1);
    ED2_OUT(0, TensorShape({B, I}), dx); ED2_OUT(1, TensorShape({I, O}), dmu); ED2_OUT(2, TensorShape({I, O}), drho);
    ED2_OUT(3, TensorShape({O}), dbias);
    ED2_TMP(TensorShape({B, O}), gp);
    ed2_reparam_bwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.act = at_.act;
    a.gy = gy.data(); a.gy_ld = O; a.y = y.data(); a.y_ld = O; a.x = x.data(); a.x_ld = I; a.w_lp = w.data(); a.w_ld = O;
    a.mu = F(mu); a.rho = F(rho); a.eps = Philox(at_.seed, 0);
    a.gp = gp.data(); a.gp_ld = O; a.dx = dx->data(); a.dx_ld = I;
    a.dmu = F(dmu); a.drho = F(drho); a.dbias = F(dbias);
    a.kl_grad_scale = at_.kl_scale; a.prior_mean = at_.pm; a.prior_std = at_.ps; a.kl_upstream = up.flat<double>().data();
    ED2_CALL(ed2_reparam_dense_bwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  KlAttrs at_;
};
ED2_REGISTER("Ed2ReparamDenseBwd", Ed2ReparamDenseBwdOp);

// ================================================================================================ DenseFlipout
// edward2/tensorflow/layers/dense.py:255-285
REGISTER_OP("Ed2FlipoutDenseFwd")
    .Input("x: T").Input("mu: float").Input("rho: float").Input("bias: float")
    .Output("y: T").Output("mu_lp: T").Output("delta_lp: T").Output("xs: T").Output("kl: double") ED2_KL_ATTRS;
template <typename T>
class Ed2FlipoutDenseFwdOp : public OpKernel {
 public:
  explicit Ed2FlipoutDenseFwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &mu = c->input(1), &rho = c->input(2), &bias = c->input(3);
    const int B = D(x, 0), I = D(x, 1), O = D(mu, 1);
    ED2_OUT(0, TensorShape({B, O}), y); ED2_OUT(1, TensorShape({I, O}), mu_lp); ED2_OUT(2, TensorShape({I, O}), delta_lp);
    ED2_OUT(3, TensorShape({B, I}), xs); ED2_OUT(4, TensorShape({}), kl);
    Tensor pert;
    OP_REQUIRES_OK(c, c->allocate_temp(DataTypeToEnum<float>::value, TensorShape({B, O}), &pert));
    auto stream = c->eigen_gpu_device().stream();
    cudaMemsetAsync(kl->flat<double>().data(), 0, sizeof(double), stream);
    ed2_flipout_fwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.act = at_.act;
    a.x = x.data(); a.x_ld = I;
    a.mu = F(mu); a.rho = F(rho); a.eps = Philox(at_.seed, 0);
    a.sign_in = Philox(at_.seed, 2, at_.row0); a.sign_out = Philox(at_.seed, 3, at_.row0); a.bias = F(bias);
    a.mu_lp = mu_lp->data(); a.mu_ld = O; a.delta_lp = delta_lp->data(); a.delta_ld = O; a.xs = xs->data(); a.xs_ld = I;
    a.pert = F(&pert); a.pert_ld = O; a.y = y->data(); a.y_ld = O;
    a.kl_out = kl->flat<double>().data(); a.kl_scale = at_.kl_scale; a.prior_mean = at_.pm; a.prior_std = at_.ps;
    ED2_CALL(ed2_flipout_dense_fwd(&a, stream));
  }
 private:
  KlAttrs at_;
};
ED2_REGISTER("Ed2FlipoutDenseFwd", Ed2FlipoutDenseFwdOp);

REGISTER_OP("Ed2FlipoutDenseBwd")
    .Input("gy: T").Input("y: T").Input("x: T").Input("xs: T").Input("mu_lp: T").Input("delta_lp: T").Input("mu: float")
    .Input("rho: float").Input("kl_upstream: double")
    .Output("dx: T").Output("dmu: float").Output("drho: float").Output("dbias: float") ED2_KL_ATTRS;
template <typename T>
class Ed2FlipoutDenseBwdOp : public OpKernel {
 public:
  explicit Ed2FlipoutDenseBwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &gy = c->input(0), &y = c->input(1), &x = c->input(2), &xs = c->input(3), &mu_lp = c->input(4),
                 &delta_lp = c->input(5), &mu = c->input(6), &rho = c->input(7), &up = c->input(8);
    const int B = D(x, 0), I = D(x, 1), O = D(mu, 1);
    ED2_OUT(0, TensorShape({B, I}), dx); ED2_OUT(1, TensorShape({I, O}), dmu); ED2_OUT(2, TensorShape({I, O}), drho);
    ED2_OUT(3, TensorShape({O}), dbias);
    ED2_TMP(TensorShape({B, O}), gp); ED2_TMP(TensorShape({B, O}), gt);
    Tensor tmp;
    OP_REQUIRES_OK(c, c->allocate_temp(DataTypeToEnum<float>::value, TensorShape({B, I}), &tmp));
    ed2_flipout_bwd_args a{};
    a.dtype = Dt<T>(); a.batch = B; a.in_dim = I; a.out_dim = O; a.act = at_.act;
    a.gy = gy.data(); a.gy_ld = O; a.y = y.data(); a.y_ld = O; a.x = x.data(); a.x_ld = I; a.xs = xs.data(); a.xs_ld = I;
    a.mu_lp = mu_lp.data(); a.mu_ld = O; a.delta_lp = delta_lp.data(); a.delta_ld = O;
    a.mu = F(mu); a.rho = F(rho); a.eps = Philox(at_.seed, 0);
    a.sign_in = Philox(at_.seed, 2, at_.row0); a.sign_out = Philox(at_.seed, 3, at_.row0);
    a.gp = gp.data(); a.gp_ld = O; a.gt = gt.data(); a.gt_ld = O; a.tmp = F(&tmp); a.tmp_ld = I;
    a.dx = dx->data(); a.dx_ld = I; a.dmu = F(dmu); a.drho = F(drho); a.dbias = F(dbias);
    a.kl_grad_scale = at_.kl_scale; a.prior_mean = at_.pm; a.prior_std = at_.ps; a.kl_upstream = up.flat<double>().data();
    ED2_CALL(ed2_flipout_dense_bwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  KlAttrs at_;
};
ED2_REGISTER("Ed2FlipoutDenseBwd", Ed2FlipoutDenseBwdOp);

// Fused DenseFlipout (bf16): the dual-accumulator launches; mu_lp / delta_lp come from Ed2FlipoutDenseFwd-style sampling
// (ed2_sample_normal_weights mode 1), xs = x∘S is produced here (first layer of a stack) or by the previous layer.
REGISTER_OP("Ed2FlipoutFusedFwd")
    .Input("x: T").Input("mu_lp: T").Input("delta_lp: T").Input("bias: float")
    .Output("y: T").Output("xs: T") ED2_KL_ATTRS;
template <typename T>
class Ed2FlipoutFusedFwdOp : public OpKernel {
 public:
  explicit Ed2FlipoutFusedFwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &mu_lp = c->input(1), &delta_lp = c->input(2), &bias = c->input(3);
    const int B = D(x, 0), I = D(x, 1), O = D(mu_lp, 1);
    ED2_OUT(0, TensorShape({B, O}), y); ED2_OUT(1, TensorShape({B, I}), xs);
    ed2_flipout_fused_fwd_args a{};
    a.batch = B; a.in_dim = I; a.out_dim = O; a.act = at_.act;
    a.x = x.data(); a.x_ld = I; a.xs = xs->data(); a.xs_ld = I; a.xs_ready = 0;
    a.sign_in = Philox(at_.seed, 2, at_.row0); a.sign_out = Philox(at_.seed, 3, at_.row0);
    a.mu_lp = mu_lp.data(); a.mu_ld = O; a.delta_lp = delta_lp.data(); a.delta_ld = O; a.bias = F(bias);
    a.y = y->data(); a.y_ld = O;
    ED2_CALL(ed2_flipout_fused_fwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  KlAttrs at_;
};
ED2_REGISTER("Ed2FlipoutFusedFwd", Ed2FlipoutFusedFwdOp);

REGISTER_OP("Ed2FlipoutFusedBwd")
    .Input("gy: T").Input("y: T").Input("x: T").Input("xs: T").Input("mu_lp: T").Input("delta_lp: T").Input("mu: float")
    .Input("rho: float").Input("kl_upstream: double")
    .Output("dx: T").Output("dmu: float").Output("drho: float").Output("dbias: float") ED2_KL_ATTRS;
template <typename T>
class Ed2FlipoutFusedBwdOp : public OpKernel {
 public:
  explicit Ed2FlipoutFusedBwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &gy = c->input(0), &y = c->input(1), &x = c->input(2), &xs = c->input(3), &mu_lp = c->input(4),
                 &delta_lp = c->input(5), &mu = c->input(6), &rho = c->input(7), &up = c->input(8);
    const int B = D(x, 0), I = D(x, 1), O = D(mu, 1);
    ED2_OUT(0, TensorShape({B, I}), dx); ED2_OUT(1, TensorShape({I, O}), dmu); ED2_OUT(2, TensorShape({I, O}), drho);
    ED2_OUT(3, TensorShape({O}), dbias);
    ED2_TMP(TensorShape({B, O}), gp); ED2_TMP(TensorShape({B, O}), gt);
    ed2_flipout_fused_bwd_args a{};
    a.batch = B; a.in_dim = I; a.out_dim = O; a.act = at_.act; a.premade = 0;
    a.gy = gy.data(); a.gy_ld = O; a.y = y.data(); a.y_ld = O;
    a.gp = gp.data(); a.gp_ld = O; a.gt = gt.data(); a.gt_ld = O;
    a.x = x.data(); a.x_ld = I; a.xs = xs.data(); a.xs_ld = I;
    a.mu_lp = mu_lp.data(); a.mu_ld = O; a.delta_lp = delta_lp.data(); a.delta_ld = O;
    a.mu = F(mu); a.rho = F(rho); a.eps = Philox(at_.seed, 0);
    a.sign_in = Philox(at_.seed, 2, at_.row0); a.sign_out = Philox(at_.seed, 3, at_.row0);
    a.dmu = F(dmu); a.drho = F(drho); a.dbias = F(dbias);
    a.kl_grad_scale = at_.kl_scale; a.prior_mean = at_.pm; a.prior_std = at_.ps; a.kl_upstream = up.flat<double>().data();
    a.dx = dx->data(); a.dx_ld = I;
    ED2_CALL(ed2_flipout_fused_bwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  KlAttrs at_;
};
ED2_REGISTER("Ed2FlipoutFusedBwd", Ed2FlipoutFusedBwdOp);

// ================================================================================================ heteroscedastic heads
// edward2/tensorflow/layers/heteroscedastic.py:178-253, 749-785 (MCSoftmaxDenseFA), hetsngp.py:143-240
struct McAttrs {
  int num_samples, num_factors, share_samples; float temperature, eps; int64_t seed, row0;
  void Read(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("num_samples", &num_samples)); OP_REQUIRES_OK(c, c->GetAttr("num_factors", &num_factors));
    OP_REQUIRES_OK(c, c->GetAttr("share_samples", &share_samples)); OP_REQUIRES_OK(c, c->GetAttr("temperature", &temperature));
    OP_REQUIRES_OK(c, c->GetAttr("eps", &eps)); OP_REQUIRES_OK(c, c->GetAttr("seed", &seed));
    OP_REQUIRES_OK(c, c->GetAttr("row0", &row0));
  }
};
#define ED2_MC_ATTRS                                                                                                   \
  .Attr("num_samples: int = 1000").Attr("num_factors: int = 10").Attr("share_samples: int = 0")                        \
  .Attr("temperature: float = 1.0").Attr("eps: float = 1e-7").Attr("seed: int = 0").Attr("row0: int = 0")
REGISTER_OP("Ed2McSoftmaxFaFwd")
    .Input("loc: float").Input("fac: float").Input("diag_raw: float")
    .Output("logits: float").Output("probs_mean: float").Output("variance: float") ED2_MC_ATTRS;
class Ed2McSoftmaxFaFwdOp : public OpKernel {
 public:
  explicit Ed2McSoftmaxFaFwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &loc = c->input(0), &fac = c->input(1), &diag = c->input(2);
    const int B = D(loc, 0), K = D(loc, 1);
    ED2_OUT(0, TensorShape({B, K}), logits); ED2_OUT(1, TensorShape({B, K}), pm); ED2_OUT(2, TensorShape({B, K}), var);
    ED2_CALL(ed2_mc_softmax_fa_fwd(F(loc), F(fac), F(diag), nullptr, 1.f, 1.f, Philox(at_.seed, 7, at_.row0), B,
                                   at_.num_samples, K, at_.num_factors, at_.share_samples, at_.temperature, at_.eps, F(pm),
                                   F(logits), F(var), c->eigen_gpu_device().stream()));
  }
 private:
  McAttrs at_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2McSoftmaxFaFwd").Device(DEVICE_GPU), Ed2McSoftmaxFaFwdOp);

REGISTER_OP("Ed2McSoftmaxFaBwd")
    .Input("g: float").Input("probs_mean: float").Input("loc: float").Input("fac: float").Input("diag_raw: float")
    .Output("dloc: float").Output("dfac: float").Output("ddiag_raw: float") ED2_MC_ATTRS;
class Ed2McSoftmaxFaBwdOp : public OpKernel {
 public:
  explicit Ed2McSoftmaxFaBwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &g = c->input(0), &pm = c->input(1), &loc = c->input(2), &fac = c->input(3), &diag = c->input(4);
    const int B = D(loc, 0), K = D(loc, 1);
    ED2_OUT(0, TensorShape({B, K}), dloc); ED2_OUT(1, TensorShape({B, D(fac, 1)}), dfac); ED2_OUT(2, TensorShape({B, K}), dd);
    ED2_CALL(ed2_mc_softmax_fa_bwd(F(g), F(pm), F(loc), F(fac), F(diag), Philox(at_.seed, 7, at_.row0), B, at_.num_samples, K,
                                   at_.num_factors, at_.share_samples, at_.temperature, at_.eps, F(dloc), F(dfac), F(dd),
                                   c->eigen_gpu_device().stream()));
  }
 private:
  McAttrs at_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2McSoftmaxFaBwd").Device(DEVICE_GPU), Ed2McSoftmaxFaBwdOp);

// ================================================================================================ SNGP head
// phi = cos(x W + b) * scale — edward2/tensorflow/layers/random_feature.py:258-266
REGISTER_OP("Ed2RffFwd").Input("x: T").Input("wt: T").Input("bias: float").Output("phi: T").Output("pre: float")
    .Attr("T: {float, bfloat16}").Attr("scale: float = 1.0");
template <typename T>
class Ed2RffFwdOp : public OpKernel {
 public:
  explicit Ed2RffFwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("scale", &scale_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &wt = c->input(1), &bias = c->input(2);
    const int B = D(x, 0), d = D(x, 1), Dn = D(wt, 0);
    ED2_OUT(0, TensorShape({B, Dn}), phi); ED2_OUT(1, TensorShape({B, Dn}), pre);
    ED2_CALL(ed2_rff_fwd(Dt<T>(), x.data(), d, wt.data(), d, F(bias), scale_, B, d, Dn, phi->data(), Dn, Dt<T>(), F(pre), Dn,
                         /*pre_mode=*/0, c->eigen_gpu_device().stream()));
  }
 private:
  float scale_;
};
ED2_REGISTER("Ed2RffFwd", Ed2RffFwdOp);

// precision <- m * precision + (1 - m) * phi^T phi / B  (or + phi^T phi) — random_feature.py:394-423; the op returns the
// updated matrix (the Python stub assigns it to the layer's precision variable)
REGISTER_OP("Ed2LaplacePrecisionUpdate").Input("phi: T").Input("precision: float").Output("updated: float")
    .Attr("T: {float, bfloat16}").Attr("momentum: float = 0.999").Attr("batch_total: int = 0");
template <typename T>
class Ed2LaplacePrecisionUpdateOp : public OpKernel {
 public:
  explicit Ed2LaplacePrecisionUpdateOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("momentum", &momentum_)); OP_REQUIRES_OK(c, c->GetAttr("batch_total", &batch_total_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &phi = c->input(0), &p = c->input(1);
    const int B = D(phi, 0), Dn = D(phi, 1);
    ED2_OUT(0, TensorShape({Dn, Dn}), out);
    auto stream = c->eigen_gpu_device().stream();
    // the library updates the matrix in place: work on the output buffer, seeded with the incoming value
    cudaMemcpyAsync(F(out), F(p), sizeof(float) * static_cast<size_t>(Dn) * Dn, cudaMemcpyDeviceToDevice, stream);
    ED2_CALL(ed2_laplace_precision_update(Dt<T>(), phi.data(), Dn, B, Dn, momentum_, batch_total_ ? batch_total_ : B, F(out),
                                          nullptr, stream));
  }
 private:
  float momentum_; int64_t batch_total_;
};
ED2_REGISTER("Ed2LaplacePrecisionUpdate", Ed2LaplacePrecisionUpdateOp);

// var = diag(ridge * phi Sigma phi^T) (+ mean-field logits) — random_feature.py:482-485, layers/utils.py:379-424
REGISTER_OP("Ed2PredictiveVariance").Input("phi: T").Input("sigma: T").Input("logits: float")
    .Output("var: float").Output("adjusted: float").Attr("T: {float, bfloat16}").Attr("ridge: float = 1.0")
    .Attr("mean_field_factor: float = -1.0").Attr("likelihood: int = 0");
template <typename T>
class Ed2PredictiveVarianceOp : public OpKernel {
 public:
  explicit Ed2PredictiveVarianceOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("ridge", &ridge_)); OP_REQUIRES_OK(c, c->GetAttr("mean_field_factor", &mff_));
    OP_REQUIRES_OK(c, c->GetAttr("likelihood", &lik_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &phi = c->input(0), &sigma = c->input(1), &logits = c->input(2);
    const int B = D(phi, 0), Dn = D(phi, 1), K = D(logits, 1);
    ED2_OUT(0, TensorShape({B}), var); ED2_OUT(1, TensorShape({B, K}), adj);
    auto stream = c->eigen_gpu_device().stream();
    cudaMemsetAsync(F(var), 0, sizeof(float) * B, stream);
    ED2_CALL(ed2_predictive_variance(Dt<T>(), phi.data(), Dn, sigma.data(), Dn, B, Dn, ridge_, F(var), F(logits), K, mff_, lik_,
                                     F(adj), stream));
  }
 private:
  float ridge_, mff_; int lik_;
};
ED2_REGISTER("Ed2PredictiveVariance", Ed2PredictiveVarianceOp);

// ================================================================================================ NormalKLDivergence
// scale_factor * KL(N(mu, softplus(rho) + 1e-7) || N(prior_mean, prior_std)) — edward2/tensorflow/regularizers.py:166-186
REGISTER_OP("Ed2NormalKlFwd").Input("mu: float").Input("rho: float").Output("kl: double")
    .Attr("prior_mean: float = 0.0").Attr("prior_std: float = 1.0").Attr("scale: float = 1.0");
class Ed2NormalKlFwdOp : public OpKernel {
 public:
  explicit Ed2NormalKlFwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("prior_mean", &m_)); OP_REQUIRES_OK(c, c->GetAttr("prior_std", &s_));
    OP_REQUIRES_OK(c, c->GetAttr("scale", &scale_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &mu = c->input(0), &rho = c->input(1);
    ED2_OUT(0, TensorShape({}), kl);
    auto stream = c->eigen_gpu_device().stream();
    cudaMemsetAsync(kl->data(), 0, sizeof(double), stream);
    ED2_CALL(ed2_normal_kl_fwd(F(mu), F(rho), mu.NumElements(), m_, s_, scale_, static_cast<double*>(kl->data()), stream));
  }
 private:
  float m_, s_, scale_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2NormalKlFwd").Device(DEVICE_GPU), Ed2NormalKlFwdOp);

REGISTER_OP("Ed2NormalKlBwd").Input("mu: float").Input("rho: float").Input("upstream: float")
    .Output("dmu: float").Output("drho: float")
    .Attr("prior_mean: float = 0.0").Attr("prior_std: float = 1.0").Attr("scale: float = 1.0");
class Ed2NormalKlBwdOp : public OpKernel {
 public:
  explicit Ed2NormalKlBwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("prior_mean", &m_)); OP_REQUIRES_OK(c, c->GetAttr("prior_std", &s_));
    OP_REQUIRES_OK(c, c->GetAttr("scale", &scale_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &mu = c->input(0), &rho = c->input(1), &up = c->input(2);
    ED2_OUT(0, mu.shape(), dmu); ED2_OUT(1, mu.shape(), drho);
    auto stream = c->eigen_gpu_device().stream();
    const size_t bytes = sizeof(float) * static_cast<size_t>(mu.NumElements());
    cudaMemsetAsync(F(dmu), 0, bytes, stream); cudaMemsetAsync(F(drho), 0, bytes, stream);  // the kernel accumulates
    ED2_CALL(ed2_normal_kl_bwd(F(mu), F(rho), mu.NumElements(), m_, s_, scale_, F(up), F(dmu), F(drho), stream));
  }
 private:
  float m_, s_, scale_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2NormalKlBwd").Device(DEVICE_GPU), Ed2NormalKlBwdOp);

// ================================================================================================ Conv2D (factorised)
// Conv2DBatchEnsemble / Conv2DRank1 — edward2/tensorflow/layers/convolutional.py:681-699, 1968-2020: the per-example
// channel factors f_in / f_out / bias are [B][C] tables (ed2_expand_table / ed2_rank1_factors build them)
#define ED2_CONV_ATTRS \
  .Attr("kh: int = 3").Attr("kw: int = 3").Attr("sh: int = 1").Attr("sw: int = 1").Attr("same: int = 1").Attr("act: int = 0")
struct ConvAttrs {
  int kh, kw, sh, sw, same, act;
  void Read(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("kh", &kh)); OP_REQUIRES_OK(c, c->GetAttr("kw", &kw));
    OP_REQUIRES_OK(c, c->GetAttr("sh", &sh)); OP_REQUIRES_OK(c, c->GetAttr("sw", &sw));
    OP_REQUIRES_OK(c, c->GetAttr("same", &same)); OP_REQUIRES_OK(c, c->GetAttr("act", &act));
  }
  ed2_conv_geometry Geometry(const Tensor& x) const {  // x is NHWC
    ed2_conv_geometry g{};
    g.batch = D(x, 0); g.height = D(x, 1); g.width = D(x, 2); g.channels = D(x, 3);
    g.kh = kh; g.kw = kw; g.sh = sh; g.sw = sw; g.same = same;
    return g;
  }
};
REGISTER_OP("Ed2ConvFactorFwd")
    .Input("x: T").Input("w: T").Input("f_in: float").Input("f_out: float").Input("bias: float")
    .Output("y: T").Output("xa: T").Output("z: T").Attr("T: {float, bfloat16}") ED2_CONV_ATTRS;
template <typename T>
class Ed2ConvFactorFwdOp : public OpKernel {
 public:
  explicit Ed2ConvFactorFwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1), &f_in = c->input(2), &f_out = c->input(3), &bias = c->input(4);
    ed2_conv_factor_fwd_args a{};
    a.geom = at_.Geometry(x);
    int ho = 0, wo = 0;
    ED2_CALL(ed2_conv_out_size(&a.geom, &ho, &wo));
    const int rows = a.geom.batch * ho * wo, K = at_.kh * at_.kw * a.geom.channels, O = D(w, 1);
    ED2_OUT(0, TensorShape({a.geom.batch, ho, wo, O}), y); ED2_OUT(1, TensorShape({rows, K}), xa);
    ED2_OUT(2, TensorShape({rows, O}), z);
    a.dtype = Dt<T>(); a.out_channels = O; a.act = at_.act;
    a.x = x.data(); a.w_lp = w.data(); a.w_ld = O;
    a.f_in = F(f_in); a.f_out = F(f_out); a.bias = F(bias);
    a.xa = xa->data(); a.xa_ld = K; a.y = y->data(); a.y_ld = O; a.z = z->data(); a.z_ld = O;
    ED2_CALL(ed2_conv_factor_fwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  ConvAttrs at_;
};
ED2_REGISTER("Ed2ConvFactorFwd", Ed2ConvFactorFwdOp);

REGISTER_OP("Ed2ConvFactorBwd")
    .Input("gy: T").Input("y: T").Input("z: T").Input("x: T").Input("xa: T").Input("w: T").Input("f_in: float").Input("f_out: float")
    .Output("dx: T").Output("dw: float").Output("df_in: float").Output("df_out: float").Output("dbias: float")
    .Attr("T: {float, bfloat16}") ED2_CONV_ATTRS;
template <typename T>
class Ed2ConvFactorBwdOp : public OpKernel {
 public:
  explicit Ed2ConvFactorBwdOp(OpKernelConstruction* c) : OpKernel(c) { at_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &gy = c->input(0), &y = c->input(1), &z = c->input(2), &x = c->input(3), &xa = c->input(4), &w = c->input(5),
                 &f_in = c->input(6), &f_out = c->input(7);
    ed2_conv_factor_bwd_args a{};
    a.geom = at_.Geometry(x);
    const int rows = D(xa, 0), K = D(xa, 1), O = D(w, 1), B = a.geom.batch, C = a.geom.channels;
    ED2_OUT(0, x.shape(), dx); ED2_OUT(1, TensorShape({K, O}), dw); ED2_OUT(2, TensorShape({B, C}), df_in);
    ED2_OUT(3, TensorShape({B, O}), df_out); ED2_OUT(4, TensorShape({B, O}), dbias);
    ED2_TMP(TensorShape({rows, O}), dz); ED2_TMP(TensorShape({rows, K}), dxa); ED2_TMP(x.shape(), gimg);
    a.dtype = Dt<T>(); a.out_channels = O; a.act = at_.act;
    a.gy = gy.data(); a.gy_ld = O; a.y = y.data(); a.y_ld = O; a.z = z.data(); a.z_ld = O;
    a.x = x.data(); a.xa = xa.data(); a.xa_ld = K; a.w_lp = w.data(); a.w_ld = O;
    a.f_in = F(f_in); a.f_out = F(f_out);
    a.dz = dz.data(); a.dz_ld = O; a.dxa = dxa.data(); a.dxa_ld = K; a.gimg = gimg.data();
    a.dx = dx->data(); a.dw = F(dw); a.df_in = F(df_in); a.df_out = F(df_out); a.dbias = F(dbias);
    ED2_CALL(ed2_conv_factor_bwd(&a, c->eigen_gpu_device().stream()));
  }
 private:
  ConvAttrs at_;
};
ED2_REGISTER("Ed2ConvFactorBwd", Ed2ConvFactorBwdOp);

// ================================================================================================ LSTM cells
// Gate arithmetic of one step — edward2/tensorflow/layers/recurrent.py:163-183 (LSTMCellReparameterization: z = x W + h U + b
// from two GEMMs), :243-269 / :330-356 (LSTMCellFlipout) and :616-640 / :747-760 (LSTMCellRank1): the input and the
// recurrent branch arrive as two pre-activation tensors z + z2.  c fp32; gate order i | f | c | o.
REGISTER_OP("Ed2LstmGatesFwd").Input("z: T").Input("z2: T").Input("c_prev: float").Output("h: T").Output("c: float")
    .Attr("T: {float, bfloat16}");
template <typename T>
class Ed2LstmGatesFwdOp : public OpKernel {
 public:
  explicit Ed2LstmGatesFwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &z = c->input(0), &z2 = c->input(1), &c_prev = c->input(2);
    const int B = D(c_prev, 0), H = D(c_prev, 1);
    ED2_OUT(0, TensorShape({B, H}), h); ED2_OUT(1, TensorShape({B, H}), cn);
    ED2_CALL(ed2_lstm_gates_fwd(z.data(), z2.data(), Dt<T>(), 4 * H, 4 * H, F(c_prev), B, H, F(cn), h->data(), H,
                                c->eigen_gpu_device().stream()));
  }
};
ED2_REGISTER("Ed2LstmGatesFwd", Ed2LstmGatesFwdOp);

REGISTER_OP("Ed2LstmGatesBwd").Input("z: T").Input("z2: T").Input("c_prev: float").Input("c: float").Input("dh: T").Input("dc: float")
    .Output("dz: T").Output("dc_prev: float").Attr("T: {float, bfloat16}");
template <typename T>
class Ed2LstmGatesBwdOp : public OpKernel {
 public:
  explicit Ed2LstmGatesBwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &z = c->input(0), &z2 = c->input(1), &c_prev = c->input(2), &cn = c->input(3), &dh = c->input(4), &dc = c->input(5);
    const int B = D(c_prev, 0), H = D(c_prev, 1);
    ED2_OUT(0, TensorShape({B, 4 * H}), dz); ED2_OUT(1, TensorShape({B, H}), dc_prev);
    ED2_CALL(ed2_lstm_gates_bwd(z.data(), z2.data(), Dt<T>(), 4 * H, 4 * H, F(c_prev), F(cn), dh.data(), H, F(dc), B, H,
                                dz->data(), 4 * H, F(dc_prev), c->eigen_gpu_device().stream()));
  }
};
ED2_REGISTER("Ed2LstmGatesBwd", Ed2LstmGatesBwdOp);

// the single-tensor form (LSTMCellReparameterization: z already holds x W + h U + b); the backward also reduces dbias
REGISTER_OP("Ed2LstmPointwiseFwd").Input("z: T").Input("c_prev: float").Output("h: T").Output("c: float")
    .Attr("T: {float, bfloat16}");
template <typename T>
class Ed2LstmPointwiseFwdOp : public OpKernel {
 public:
  explicit Ed2LstmPointwiseFwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &z = c->input(0), &c_prev = c->input(1);
    const int B = D(c_prev, 0), H = D(c_prev, 1);
    ED2_OUT(0, TensorShape({B, H}), h); ED2_OUT(1, TensorShape({B, H}), cn);
    ED2_CALL(ed2_lstm_pointwise_fwd(z.data(), Dt<T>(), 4 * H, F(c_prev), B, H, F(cn), h->data(), H,
                                    c->eigen_gpu_device().stream()));
  }
};
ED2_REGISTER("Ed2LstmPointwiseFwd", Ed2LstmPointwiseFwdOp);

REGISTER_OP("Ed2LstmPointwiseBwd").Input("z: T").Input("c_prev: float").Input("c: float").Input("dh: T").Input("dc: float")
    .Output("dz: T").Output("dc_prev: float").Output("dbias: float").Attr("T: {float, bfloat16}");
template <typename T>
class Ed2LstmPointwiseBwdOp : public OpKernel {
 public:
  explicit Ed2LstmPointwiseBwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &z = c->input(0), &c_prev = c->input(1), &cn = c->input(2), &dh = c->input(3), &dc = c->input(4);
    const int B = D(c_prev, 0), H = D(c_prev, 1);
    ED2_OUT(0, TensorShape({B, 4 * H}), dz); ED2_OUT(1, TensorShape({B, H}), dc_prev); ED2_OUT(2, TensorShape({4 * H}), dbias);
    ED2_CALL(ed2_lstm_pointwise_bwd(z.data(), Dt<T>(), 4 * H, F(c_prev), F(cn), dh.data(), H, F(dc), B, H, dz->data(), 4 * H,
                                    F(dc_prev), F(dbias), c->eigen_gpu_device().stream()));
  }
};
ED2_REGISTER("Ed2LstmPointwiseBwd", Ed2LstmPointwiseBwdOp);

// ================================================================================================ SNGP head (continued)
// d phi / d x of the random-feature layer — random_feature.py:258-266 (+ autodiff); w [d][D]
REGISTER_OP("Ed2RffBwd").Input("dphi: T").Input("pre: float").Input("bias: float").Input("w: T").Output("dx: T")
    .Attr("T: {float, bfloat16}").Attr("scale: float = 1.0");
template <typename T>
class Ed2RffBwdOp : public OpKernel {
 public:
  explicit Ed2RffBwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("scale", &scale_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &dphi = c->input(0), &pre = c->input(1), &bias = c->input(2), &w = c->input(3);
    const int B = D(dphi, 0), Dn = D(dphi, 1), d = D(w, 0);
    ED2_OUT(0, TensorShape({B, d}), dx);
    ED2_TMP(TensorShape({B, Dn}), gpre);
    ED2_CALL(ed2_rff_bwd(Dt<T>(), dphi.data(), Dn, Dt<T>(), F(pre), Dn, /*pre_mode=*/0, F(bias), w.data(), Dn, scale_, B, d, Dn,
                         gpre.data(), Dn, dx->data(), d, Dt<T>(), c->eigen_gpu_device().stream()));
  }
 private:
  float scale_;
};
ED2_REGISTER("Ed2RffBwd", Ed2RffBwdOp);

// tf.keras.layers.LayerNormalization in front of the random features (random_feature.py:167-171, 252) and its backward
REGISTER_OP("Ed2LayerNormFwd").Input("x: T").Input("gamma: float").Input("beta: float").Output("y: T")
    .Attr("T: {float, bfloat16}").Attr("epsilon: float = 1e-3");
template <typename T>
class Ed2LayerNormFwdOp : public OpKernel {
 public:
  explicit Ed2LayerNormFwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &gamma = c->input(1), &beta = c->input(2);
    const int B = D(x, 0), d = D(x, 1);
    ED2_OUT(0, TensorShape({B, d}), y);
    ED2_CALL(ed2_input_normalize(x.data(), Dt<T>(), d, y->data(), Dt<T>(), d, B, d, F(gamma), F(beta), eps_, /*scale_only=*/0,
                                 1.f, c->eigen_gpu_device().stream()));
  }
 private:
  float eps_;
};
ED2_REGISTER("Ed2LayerNormFwd", Ed2LayerNormFwdOp);

REGISTER_OP("Ed2LayerNormBwd").Input("x: T").Input("dy: T").Input("gamma: float")
    .Output("dx: T").Output("dgamma: float").Output("dbeta: float").Attr("T: {float, bfloat16}").Attr("epsilon: float = 1e-3");
template <typename T>
class Ed2LayerNormBwdOp : public OpKernel {
 public:
  explicit Ed2LayerNormBwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &dy = c->input(1), &gamma = c->input(2);
    const int B = D(x, 0), d = D(x, 1);
    ED2_OUT(0, TensorShape({B, d}), dx); ED2_OUT(1, TensorShape({d}), dgamma); ED2_OUT(2, TensorShape({d}), dbeta);
    ED2_CALL(ed2_layernorm_bwd(x.data(), Dt<T>(), d, dy.data(), Dt<T>(), d, B, d, F(gamma), eps_, dx->data(), Dt<T>(), d,
                               F(dgamma), F(dbeta), c->eigen_gpu_device().stream()));
  }
 private:
  float eps_;
};
ED2_REGISTER("Ed2LayerNormBwd", Ed2LayerNormBwdOp);

// mean-field adjusted logits from the [B] predictive variances — edward2/tensorflow/layers/utils.py:379-424
REGISTER_OP("Ed2MeanFieldLogits").Input("logits: float").Input("var: float").Output("adjusted: float")
    .Attr("mean_field_factor: float = 1.0").Attr("likelihood: int = 0");
class Ed2MeanFieldLogitsOp : public OpKernel {
 public:
  explicit Ed2MeanFieldLogitsOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("mean_field_factor", &mff_)); OP_REQUIRES_OK(c, c->GetAttr("likelihood", &lik_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &logits = c->input(0), &var = c->input(1);
    const int B = D(logits, 0), K = D(logits, 1);
    ED2_OUT(0, TensorShape({B, K}), adj);
    ED2_CALL(ed2_mean_field_logits(F(logits), F(var), B, K, mff_, lik_, F(adj), c->eigen_gpu_device().stream()));
  }
 private:
  float mff_; int lik_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2MeanFieldLogits").Device(DEVICE_GPU), Ed2MeanFieldLogitsOp);

// covariance = inv(ridge I + precision) — random_feature.py:447-459 (tf.linalg.inv), as the library's block Gauss-Jordan
REGISTER_OP("Ed2SpdInverse").Input("a: float").Output("inverse: float").Output("status: int32");
class Ed2SpdInverseOp : public OpKernel {
 public:
  explicit Ed2SpdInverseOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor& a = c->input(0);
    const int n = D(a, 0);
    ED2_OUT(0, TensorShape({n, n}), inv); ED2_OUT(1, TensorShape({}), status);
    const long long wsn = ed2_spd_inverse_workspace_floats(n);
    Tensor ws;
    OP_REQUIRES_OK(c, c->allocate_temp(DataTypeToEnum<float>::value, TensorShape({wsn}), &ws));
    auto stream = c->eigen_gpu_device().stream();
    cudaMemcpyAsync(F(inv), F(a), sizeof(float) * n * n, cudaMemcpyDeviceToDevice, stream);
    ED2_CALL(ed2_spd_inverse(F(inv), n, F(&ws), stream));
    // the last workspace word is the status flag (non-zero: a pivot was not positive)
    cudaMemcpyAsync(status->data(), F(&ws) + wsn - 1, sizeof(int), cudaMemcpyDeviceToDevice, stream);
  }
};
REGISTER_KERNEL_BUILDER(Name("Ed2SpdInverse").Device(DEVICE_GPU), Ed2SpdInverseOp);

// SpectralNormalization.update_weights — edward2/tensorflow/layers/normalization.py:369-391 (mode 0, in-place kernel)
REGISTER_OP("Ed2SpectralNormUpdate").Input("w: float").Input("u: float").Input("v: float")
    .Output("w_out: float").Output("sigma: float").Attr("iterations: int = 1").Attr("norm_multiplier: float = 0.95")
    .Attr("mode: int = 0").Attr("update_uv: int = 1");
class Ed2SpectralNormUpdateOp : public OpKernel {
 public:
  explicit Ed2SpectralNormUpdateOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("iterations", &it_)); OP_REQUIRES_OK(c, c->GetAttr("norm_multiplier", &nm_));
    OP_REQUIRES_OK(c, c->GetAttr("mode", &mode_)); OP_REQUIRES_OK(c, c->GetAttr("update_uv", &upd_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &w = c->input(0), &u = c->input(1), &v = c->input(2);
    const int I = D(w, 0), O = D(w, 1);
    ED2_OUT(0, TensorShape({I, O}), w_out); ED2_OUT(1, TensorShape({}), sigma);
    Tensor ws;
    OP_REQUIRES_OK(c, c->allocate_temp(DataTypeToEnum<float>::value, TensorShape({I + O + 8}), &ws));
    ED2_CALL(ed2_spectral_norm_update(F(w), I, O, const_cast<float*>(F(u)), const_cast<float*>(F(v)), it_, nm_, mode_, upd_,
                                      F(w_out), nullptr, 0, F(sigma), F(&ws), c->eigen_gpu_device().stream()));
  }
 private:
  int it_, mode_, upd_; float nm_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2SpectralNormUpdate").Device(DEVICE_GPU), Ed2SpectralNormUpdateOp);

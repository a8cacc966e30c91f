// TensorFlow custom-op shim for libed2b200 — COMPILE-GATED: needs the TensorFlow headers (`tf.sysconfig.get_include()`),
// which are not in this image; NOT built by edward2_b200/_build.py and UNTESTED here.
// Build (where TensorFlow is installed):
//   g++ -shared -fPIC -std=c++17 $(python -c "import tensorflow as tf;print(' '.join(tf.sysconfig.get_compile_flags()))") \
//       -I../include tf_op_shim.cc -L../edward2_b200 -led2b200 \
//       $(python -c "import tensorflow as tf;print(' '.join(tf.sysconfig.get_link_flags()))") -o libed2b200_tf.so
#if __has_include("tensorflow/core/framework/op_kernel.h")
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "ed2b200.h"

using namespace tensorflow;  // NOLINT

// Replaces the body of DenseReparameterization.call (edward2/tensorflow/layers/dense.py:73-83).
REGISTER_OP("Ed2ReparamDenseFwd")
    .Input("x: T").Input("mu: float").Input("rho: float").Input("bias: float")
    .Output("y: T").Output("w_lp: T").Output("kl: double")
    .Attr("T: {float, bfloat16}").Attr("act: int = 0").Attr("seed: int = 0").Attr("stream_id: int = 0")
    .Attr("kl_scale: float = 1.0").Attr("prior_mean: float = 0.0").Attr("prior_std: float = 1.0")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(0), 0), c->Dim(c->input(1), 1)));
      c->set_output(1, c->input(1));
      c->set_output(2, c->Scalar());
      return OkStatus();
    });

template <typename T>
class Ed2ReparamDenseFwdOp : public OpKernel {
 public:
  explicit Ed2ReparamDenseFwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("act", &act_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
    OP_REQUIRES_OK(c, c->GetAttr("stream_id", &stream_id_));
    OP_REQUIRES_OK(c, c->GetAttr("kl_scale", &kl_scale_));
    OP_REQUIRES_OK(c, c->GetAttr("prior_mean", &pm_));
    OP_REQUIRES_OK(c, c->GetAttr("prior_std", &ps_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &mu = c->input(1), &rho = c->input(2), &bias = c->input(3);
    Tensor *y, *w, *kl;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({x.dim_size(0), mu.dim_size(1)}), &y));
    OP_REQUIRES_OK(c, c->allocate_output(1, mu.shape(), &w));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({}), &kl));
    auto stream = c->eigen_gpu_device().stream();
    cudaMemsetAsync(kl->flat<double>().data(), 0, sizeof(double), stream);
    ed2_reparam_fwd_args a{};
    a.dtype = std::is_same<T, float>::value ? ED2_F32 : ED2_BF16;
    a.batch = x.dim_size(0); a.in_dim = x.dim_size(1); a.out_dim = mu.dim_size(1); a.act = act_;
    a.x = x.data(); a.x_ld = a.in_dim;
    a.mu = mu.flat<float>().data(); a.rho = rho.flat<float>().data();
    a.eps = ed2_noise{nullptr, static_cast<uint64_t>(seed_), static_cast<uint64_t>(stream_id_), nullptr};
    a.bias = bias.flat<float>().data();
    a.w_lp = w->data(); a.w_ld = a.out_dim;  // out_dim must be a multiple of 8 for bf16 (TMA pitch rule)
    a.y = y->data(); a.y_ld = a.out_dim;
    a.kl_out = kl->flat<double>().data(); a.kl_scale = kl_scale_; a.prior_mean = pm_; a.prior_std = ps_;
    OP_REQUIRES(c, ed2_reparam_dense_fwd(&a, stream) == ED2_OK, errors::Internal(ed2_last_error()));
  }
 private:
  int act_; int64_t seed_, stream_id_; float kl_scale_, pm_, ps_;
};
REGISTER_KERNEL_BUILDER(Name("Ed2ReparamDenseFwd").Device(DEVICE_GPU).TypeConstraint<float>("T"), Ed2ReparamDenseFwdOp<float>);
REGISTER_KERNEL_BUILDER(Name("Ed2ReparamDenseFwd").Device(DEVICE_GPU).TypeConstraint<bfloat16>("T"), Ed2ReparamDenseFwdOp<bfloat16>);
// The backward op (Ed2ReparamDenseBwd -> ed2_reparam_dense_bwd) and the other layers register the same way and are
// attached with tf.custom_gradient on the Python side.
#else
#error "TensorFlow headers not found: this shim only builds where TensorFlow is installed (see INTEGRATION.md)"
#endif

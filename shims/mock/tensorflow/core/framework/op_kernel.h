// MOCK of the parts of the TensorFlow C++ op API that shims/tf_op_shim.cc uses — for TYPE-CHECKING the shim in an image
// without TensorFlow (tests/test_host.py::test_shims_type_check, g++ -fsyntax-only).  Not an implementation.
#pragma once
#include <cstdint>
#include <initializer_list>
#include <string>
#include <type_traits>
#include <vector>

typedef struct CUstream_st* cudaStream_t;
extern "C" int cudaMemsetAsync(void*, int, size_t, cudaStream_t);
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
extern "C" int cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t);

namespace tensorflow {

struct bfloat16 { uint16_t v; };
using int64 = long long;
class Status {
 public:
  bool ok() const { return true; }
};
inline Status OkStatus() { return Status(); }
namespace errors {
inline Status Internal(const char*) { return Status(); }
inline Status InvalidArgument(const char*) { return Status(); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t>) {}
};
template <typename T>
struct Flat {
  T* data() const { return nullptr; }
};
class Tensor {
 public:
  int64_t dim_size(int) const { return 0; }
  int64_t NumElements() const { return 0; }
  const TensorShape& shape() const { return s_; }
  void* data() const { return nullptr; }
  template <typename T> Flat<T> flat() const { return {}; }
 private:
  TensorShape s_;
};
struct GpuDevice {
  cudaStream_t stream() const { return nullptr; }
};
class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char*, T*) const { return Status(); }
};
class OpKernelContext {
 public:
  const Tensor& input(int) const { return t_; }
  Status allocate_output(int, const TensorShape&, Tensor**) { return Status(); }
  Status allocate_temp(int, const TensorShape&, Tensor*) { return Status(); }
  const GpuDevice& eigen_gpu_device() const { return d_; }
  void CtxFailure(const Status&) {}
 private:
  Tensor t_;
  GpuDevice d_;
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext*) = 0;
};
template <typename T> struct DataTypeToEnum { static constexpr int value = 0; };
constexpr const char* DEVICE_GPU = "GPU";

namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int) { return {}; }
  DimensionHandle Dim(ShapeHandle, int) { return {}; }
  ShapeHandle Matrix(DimensionHandle, DimensionHandle) { return {}; }
  ShapeHandle Vector(DimensionHandle) { return {}; }
  ShapeHandle Scalar() { return {}; }
  void set_output(int, ShapeHandle) {}
};
}  // namespace shape_inference

struct OpDefBuilder {
  OpDefBuilder& Input(const char*) { return *this; }
  OpDefBuilder& Output(const char*) { return *this; }
  OpDefBuilder& Attr(const char*) { return *this; }
  template <typename F> OpDefBuilder& SetShapeFn(F) { return *this; }
};
struct KernelDefBuilder {
  KernelDefBuilder& Device(const char*) { return *this; }
  template <typename T> KernelDefBuilder& TypeConstraint(const char*) { return *this; }
};
inline KernelDefBuilder Name(const char*) { return {}; }

}  // namespace tensorflow

#define ED2_TF_CAT2(a, b) a##b
#define ED2_TF_CAT(a, b) ED2_TF_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilder ED2_TF_CAT(ed2_op_def_, __COUNTER__) = ::tensorflow::OpDefBuilder()
#define REGISTER_KERNEL_BUILDER(builder, ...) \
  static_assert(std::is_base_of<::tensorflow::OpKernel, __VA_ARGS__>::value, "not an OpKernel"); \
  static ::tensorflow::KernelDefBuilder ED2_TF_CAT(ed2_kernel_def_, __COUNTER__) = builder
#define OP_REQUIRES_OK(ctx, expr) do { ::tensorflow::Status _s = (expr); (void)_s; } while (0)
#define OP_REQUIRES(ctx, cond, status) do { if (!(cond)) { (ctx)->CtxFailure(status); return; } } while (0)

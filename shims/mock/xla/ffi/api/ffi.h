// MOCK of the parts of jaxlib's xla/ffi/api/ffi.h that shims/jax_ffi_shim.cc uses — for TYPE-CHECKING the shim in an
// image without jaxlib (tests/test_host.py::test_shims_type_check compiles the shim against it with -fsyntax-only).
// It reproduces the public shapes of the API (Ffi::Bind().Ctx<>().Arg<>().Ret<>().Attr<>(), Buffer / AnyBuffer /
// Result accessors, Error) and makes XLA_FFI_DEFINE_HANDLER_SYMBOL a static_assert that the handler is invocable with
// exactly the bound argument list.  It is NOT an implementation: nothing here can run.
#pragma once
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

typedef struct CUstream_st* cudaStream_t;

namespace xla {
namespace ffi {

enum DataType { F32, BF16, F64, S32, S64, U64, PRED };

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  const T& operator[](size_t i) const { return ptr[i]; }
  size_t size() const { return n; }
};

class AnyBuffer {
 public:
  DataType element_type() const { return F32; }
  Span<int64_t> dimensions() const { return {}; }
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
};

template <DataType DT> struct NativeOf { using type = void; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<U64> { using type = uint64_t; };

template <DataType DT>
class Buffer {
 public:
  using T = typename NativeOf<DT>::type;
  Span<int64_t> dimensions() const { return {}; }
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
};

template <typename B>
class Result {
 public:
  B* operator->() { return &b_; }
  B& operator*() { return b_; }
 private:
  B b_;
};

class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(const std::string&) { return Error(); }
  static Error InvalidArgument(const std::string&) { return Error(); }
};

template <typename T> struct PlatformStream { using type = T; };

template <typename... Bound>
struct Binding {
  template <typename C> Binding<Bound..., typename C::type> Ctx() const { return {}; }
  template <typename A> Binding<Bound..., A> Arg() const { return {}; }
  template <typename R> Binding<Bound..., Result<R>> Ret() const { return {}; }
  template <typename T> Binding<Bound..., T> Attr(const char*) const { return {}; }
  template <typename F> static constexpr bool Matches() { return std::is_invocable_r<Error, F, Bound...>::value; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, fn, binding)                                                     \
  static_assert(decltype(binding)::template Matches<decltype(&fn)>(), #fn " does not match its binding"); \
  extern "C" void* sym() { return reinterpret_cast<void*>(&fn); }

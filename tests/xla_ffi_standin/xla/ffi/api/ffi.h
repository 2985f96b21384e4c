// TEST ONLY: a stand-in for the part of the public XLA FFI C++ API (`xla/ffi/api/ffi.h`, shipped with jaxlib, absent from
// this image) that cnf_ot_b200/csrc/xla_ffi_shim.cc uses.  It exists so the shim can be COMPILED here: the binder below
// records the C++ type every Ctx / Arg / Attr / Ret contributes and `To(fn)` static_asserts that the handler is callable
// with exactly those types and returns ffi::Error -- the same check the real header performs.  It does not run anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { F32, F64, S32, U8 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<F64> { using type = double; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<U8> { using type = uint8_t; };

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  size_t element_count() const { return count_; }

 private:
  T* data_ = nullptr;
  size_t count_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

template <typename T>
class Span {
 public:
  const std::remove_const_t<T>* begin() const { return data_; }
  size_t size() const { return size_; }

 private:
  const std::remove_const_t<T>* data_ = nullptr;
  size_t size_ = 0;
};

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T> struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename C> auto Ctx() const { return CtxImpl(static_cast<C*>(nullptr)); }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename Fn>
  int To(Fn&& fn) const {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>, "handler signature does not match its XLA FFI binding");
    (void)fn;
    return 0;
  }

 private:
  template <typename S> Binding<Ts..., S> CtxImpl(PlatformStream<S>*) const { return {}; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines `XLA_FFI_Error* name(XLA_FFI_CallFrame*)`
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, ...)            \
  static const int name##_checked = (__VA_ARGS__).To(impl);       \
  extern "C" void* name(void* call_frame) { return (void)name##_checked, call_frame; }

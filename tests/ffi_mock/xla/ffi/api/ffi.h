// MINIMAL MOCK of the XLA FFI C++ API (xla/ffi/api/ffi.h as shipped by jaxlib, jax.ffi.include_dir()) -- TEST
// INFRASTRUCTURE ONLY.  It declares just the names csrc/xla_ffi.cc uses, with the shapes they have in the real header
// (Ffi::Bind() builder with Ctx/Arg/Ret/Attr, Buffer<dtype> with typed_data/element_count/size_bytes/dimensions,
// Result<>, Span, Error, ScratchAllocator, PlatformStream, XLA_FFI_DEFINE_HANDLER_SYMBOL), so that the shim can be
// type-checked where jax is absent (tests/test_xla_ffi_shim.py).  The binder also checks that the handler's
// parameter list matches the binding, as the real one does.  It does not run anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <tuple>
#include <type_traits>

namespace xla { namespace ffi {

enum class DataType { U8, S32, S64, F32, F64 };
inline constexpr DataType U8 = DataType::U8, S32 = DataType::S32, S64 = DataType::S64, F32 = DataType::F32, F64 = DataType::F64;
template <DataType> struct NativeOf;
template <> struct NativeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeOf<DataType::F32> { using type = float; };
template <> struct NativeOf<DataType::F64> { using type = double; };

template <class T> class Span {
 public:
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
 private:
  const T* p_ = nullptr; size_t n_ = 0;
};

template <DataType D> class Buffer {
 public:
  using T = typename NativeOf<D>::type;
  T* typed_data() const { return p_; }
  void* untyped_data() const { return p_; }
  size_t element_count() const { return n_; }
  size_t size_bytes() const { return n_ * sizeof(T); }
  Span<const int64_t> dimensions() const { return {}; }
 private:
  T* p_ = nullptr; size_t n_ = 0;
};
template <class B> class Result {
 public:
  B* operator->() { return &b_; }
  B& operator*() { return b_; }
 private:
  B b_;
};
template <DataType D> using ResultBuffer = Result<Buffer<D>>;

class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(std::string m) { return Error(std::move(m)); }
  static Error InvalidArgument(std::string m) { return Error(std::move(m)); }
  bool success() const { return ok_; }
 private:
  Error() : ok_(true) {}
  explicit Error(std::string m) : ok_(false), msg_(std::move(m)) {}
  bool ok_; std::string msg_;
};

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t, size_t = 1) { return std::nullopt; }
};
template <class S> struct PlatformStream { using stream = S; };

template <class T> struct CtxArg { using type = T; };
template <class S> struct CtxArg<PlatformStream<S>> { using type = S; };
template <class T> struct RetArg { using type = Result<T>; };

template <class... Ps> struct Binding {
  template <class T> Binding<Ps..., typename CtxArg<T>::type> Ctx() const { return {}; }
  template <class T> Binding<Ps..., T> Arg() const { return {}; }
  template <class T> Binding<Ps..., typename RetArg<T>::type> Ret() const { return {}; }
  template <class T> Binding<Ps..., T> Attr(const char*) const { return {}; }
  // the handler must be callable with exactly the bound parameter list
  template <class F> static constexpr bool Matches() { return std::is_invocable_r_v<Error, F, Ps...>; }
};
struct Ffi { static Binding<> Bind() { return {}; } };

}}  // namespace xla::ffi

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                                                     \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                       \
                "XLA FFI binding of " #fn " does not match the signature of " #impl);                        \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame*) { return nullptr; }

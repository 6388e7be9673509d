// TEST INFRASTRUCTURE - NOT XLA.  A stand-in for the part of xla/ffi/api/ffi.h that integration/jax_ffi/durf_ffi.cc uses,
// so that the handler BODIES (every call into include/durf_b200.h: argument count, order and types) are compiled in an image
// that has no jaxlib.  It mirrors the public names of the real header (xla::ffi::Buffer<dtype>, Result<>, AnyBuffer, Error,
// Ffi::Bind().Ctx<>().Arg<>().Attr<>().Ret<>(), XLA_FFI_DEFINE_HANDLER_SYMBOL) and additionally checks at compile time that
// a binding's Ctx/Arg/Attr/Ret list matches the implementation's parameter list one to one.  tests/test_ffi_shim.py uses
// the real header instead whenever it finds one.
#pragma once
#define DURF_FFI_STUB_HEADER 1

#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { S32, U8, F32, BF16 };
template <DataType> struct NativeOf;
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<BF16> { using type = uint16_t; };

struct Dims {
  const int64_t* p; size_t n;
  int64_t operator[](size_t i) const { return p[i]; }
  size_t size() const { return n; }
};

class AnyBuffer {
 public:
  void* untyped_data() const { return data_; }
  Dims dimensions() const { return Dims{dims_, rank_}; }
  size_t element_count() const { size_t c = 1; for (size_t i = 0; i < rank_; ++i) c *= (size_t)dims_[i]; return c; }
  size_t size_bytes() const { return element_count() * elem_; }
  void* data_ = nullptr; const int64_t* dims_ = nullptr; size_t rank_ = 0; size_t elem_ = 1;
};
template <DataType T>
class Buffer : public AnyBuffer {
 public:
  using N = typename NativeOf<T>::type;
  N* typed_data() const { return static_cast<N*>(data_); }
};
template <class B>
class Result {
 public:
  B* operator->() { return &b_; }
  B& operator*() { return b_; }
  B b_;
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

enum class ErrorCode { kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <class S> struct PlatformStream {};

// binding: a type list that must equal the implementation's parameter list
template <class... Ts> struct Binding {
  template <class C> auto Ctx() const { return CtxImpl(static_cast<C*>(nullptr)); }       // Ctx<PlatformStream<S>>() supplies an S
  template <class S> Binding<Ts..., S> CtxImpl(PlatformStream<S>*) const { return {}; }
  template <class A> Binding<Ts..., A> Arg() const { return {}; }
  template <class A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <class R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <class... Ps> static constexpr bool Matches(Error (*)(Ps...)) { return std::is_same<std::tuple<Ts...>, std::tuple<Ps...>>::value; }
};
struct Ffi { static Binding<> Bind() { return {}; } };

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                               \
  static_assert(decltype(binding)::Matches(impl), #name ": the binding does not match the parameters of " #impl);       \
  extern "C" void* name() { return reinterpret_cast<void*>(&impl); }

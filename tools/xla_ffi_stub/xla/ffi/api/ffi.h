// SYNTAX-CHECK STUB of the small part of XLA's C++ FFI API (xla/ffi/api/ffi.h) that csrc/fem_b200_xla.cc uses.
// jax / jaxlib are not installable in the build image, so `python -m jax_fem_b200.build --check-xla-shim` compiles the shim
// against this mock to catch typos; a real build uses jax.ffi.include_dir().  Nothing here is linked into the library.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>
struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
namespace xla::ffi {
enum DataType { S32, S64, F64, U8 };
template <DataType> struct NativeType;
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<S64> { using type = int64_t; };
template <> struct NativeType<F64> { using type = double; };
template <> struct NativeType<U8> { using type = uint8_t; };
template <class T> struct Span {
  const T* ptr = nullptr; size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};
template <DataType dtype> struct Buffer {
  using T = typename NativeType<dtype>::type;
  T* typed_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  Span<int64_t> dimensions() const { return {}; }
};
template <DataType dtype> struct ResultBuffer {
  Buffer<dtype> b;
  Buffer<dtype>* operator->() { return &b; }
};
enum class ErrorCode { kInternal, kInvalidArgument };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};
template <class T> struct PlatformStream {};
template <class... Ts> struct Binding {
  template <class T> Binding<Ts..., T> Ctx() { return {}; }
  template <class T> Binding<Ts..., T> Arg() { return {}; }
  template <class T> Binding<Ts..., T> Ret() { return {}; }
  template <class T> Binding<Ts..., T> Attr(const char*) { return {}; }
};
struct Ffi { static Binding<> Bind() { return {}; } };
}  // namespace xla::ffi
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                               \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) {                                   \
    (void)sizeof(&impl);                                                                 \
    auto b = binding;                                                                    \
    (void)b;                                                                             \
    return nullptr;                                                                      \
  }

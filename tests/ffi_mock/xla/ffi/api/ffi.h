// MOCK of the XLA FFI C++ API surface that jaxrk_b200/csrc/xla_ffi_shim.cc uses (jax / jaxlib are absent from the build
// image, so the real "xla/ffi/api/ffi.h" is not available).  TEST INFRASTRUCTURE ONLY: tests/test_ffi_shim.py compiles the
// shim against this header so that (i) the shim is syntactically valid C++ that has been through a compiler, (ii) every
// handler's signature is checked -- at compile time -- against its Bind() chain (Ctx / Arg / Attr / Ret, in order), and
// (iii) every call into include/jrk.h type-checks against the header.  Semantics follow the documented API:
// Buffer<F32>::dimensions() / typed_data() / element_count(), ResultBuffer<F32> = Result<Buffer<F32>> with operator->,
// ScratchAllocator::Allocate -> std::optional<void*>, Error::{Success, InvalidArgument, Internal}.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32 };

template <typename T>
struct Span {
    const T* p = nullptr;
    size_t n = 0;
    size_t size() const { return n; }
    const T& operator[](size_t i) const { return p[i]; }
};

template <DataType>
struct Buffer {
    std::vector<int64_t> dims;
    float* data = nullptr;
    Span<int64_t> dimensions() const { return Span<int64_t>{dims.data(), dims.size()}; }
    float* typed_data() const { return data; }
    size_t element_count() const {
        size_t n = 1;
        for (auto d : dims) n *= (size_t)d;
        return n;
    }
};
template <typename T>
struct Result {
    T value;
    T* operator->() { return &value; }
    T& operator*() { return value; }
};
template <DataType T>
using ResultBuffer = Result<Buffer<T>>;

struct Error {
    int code = 0;
    std::string msg;
    static Error Success() { return Error{}; }
    static Error InvalidArgument(std::string m) { return Error{1, std::move(m)}; }
    static Error Internal(std::string m) { return Error{2, std::move(m)}; }
};

struct ScratchAllocator {
    std::optional<void*> Allocate(size_t, size_t = 1) { return std::nullopt; }
};
template <typename S>
struct PlatformStream {};
enum class Traits { kCmdBufferCompatible };

// ---- the Bind() chain as a type list of the handler's parameters ----
template <typename T> struct CtxParam { using type = T; };
template <typename S> struct CtxParam<PlatformStream<S>> { using type = S; };
template <typename T> struct RetParam;
template <DataType D> struct RetParam<Buffer<D>> { using type = ResultBuffer<D>; };

template <typename... Ps>
struct Binding {
    template <typename T> Binding<Ps..., typename CtxParam<T>::type> Ctx() const { return {}; }
    template <typename T> Binding<Ps..., T> Arg() const { return {}; }
    template <typename T> Binding<Ps..., T> Attr(const char*) const { return {}; }
    template <typename T> Binding<Ps..., typename RetParam<T>::type> Ret() const { return {}; }
    template <typename F>
    static constexpr bool Matches() { return std::is_invocable_r<Error, F, Ps...>::value; }
    static constexpr size_t arity = sizeof...(Ps);
};
struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines `extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*)`; the mock checks the handler against the
// binding and exports a symbol of the same name
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding, ...)                                                        \
    static_assert(decltype(binding)::template Matches<decltype(&impl)>(), #name ": handler signature != Bind() chain"); \
    extern "C" const size_t name = decltype(binding)::arity

// XLA FFI handler shim over the C ABI of include/jrk.h  (binding "L1a").
//
// jax / jaxlib are absent from the build image, so this file is NOT part of libjaxrk_b200.so by default: it is
// compiled by a maintainer who has jax installed,
//     g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude \
//         jaxrk_b200/csrc/xla_ffi_shim.cc -Ljaxrk_b200/lib -ljaxrk_b200 -o libjaxrk_b200_ffi.so
// and registered from Python as INTEGRATION.md shows.  Each handler forwards XLA-owned device buffers and the
// platform stream to one entry point of jrk.h; nothing synchronises, nothing is retained, errors become ffi::Error.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include "../../include/jrk.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error to_error(int rc) {
    if (rc == JRK_OK) return ffi::Error::Success();
    char msg[512];
    jrk_last_error(msg, sizeof(msg));
    return rc == JRK_ERR_INVALID_ARG ? ffi::Error::InvalidArgument(msg) : ffi::Error::Internal(msg);
}

// K = GenGaussKernel(X, Y)   (jaxrk/kern/rbf.py:104-105); self-Gram when x and y are the same buffer
static ffi::Error GramImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x,
                           ffi::Buffer<ffi::F32> y, float scale, float power, float nconst, int32_t path,
                           ffi::ResultBuffer<ffi::F32> k) {
    const auto xd = x.dimensions(), yd = y.dimensions();
    if (xd.size() != 2 || yd.size() != 2 || xd[1] != yd[1]) return ffi::Error::InvalidArgument("X, Y must be N x d, M x d");
    const int64_t N = xd[0], M = yd[0];
    const int d = (int)xd[1];
    const size_t wsb = jrk_gram_workspace_bytes(N, M, d, path);
    void* ws = nullptr;
    if (wsb) {
        auto a = scratch.Allocate(wsb);
        if (!a.has_value()) return ffi::Error::Internal("jrk_gram_f32: scratch allocation failed");
        ws = *a;
    }
    return to_error(jrk_gram_f32(x.typed_data(), y.typed_data(), k->typed_data(), N, M, d, d, d, M, scale, nullptr, power,
                                 nconst, path, ws, wsb, stream));
}

// out = row_reduce(diag(row_pref) K(X, Y) diag(col_pref) W)   (jaxrk/rkhs/vector.py:62-75)
static ffi::Error KmatwImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x,
                            ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> w, float scale, float power, float nconst,
                            int32_t row_reduce, int64_t group, ffi::ResultBuffer<ffi::F32> out) {
    const auto xd = x.dimensions(), yd = y.dimensions(), wd = w.dimensions();
    if (xd.size() != 2 || yd.size() != 2 || wd.size() != 2 || xd[1] != yd[1] || wd[0] != yd[0])
        return ffi::Error::InvalidArgument("X: N x d, Y: M x d, W: M x m expected");
    const int64_t N = xd[0], M = yd[0];
    const int d = (int)xd[1], m = (int)wd[1];
    const size_t wsb = jrk_kmatw_workspace_bytes(N, M, d, m, row_reduce, group);
    void* ws = nullptr;
    if (wsb) {
        auto a = scratch.Allocate(wsb);
        if (!a.has_value()) return ffi::Error::Internal("jrk_kmatw_f32: scratch allocation failed");
        ws = *a;
    }
    return to_error(jrk_kmatw_f32(x.typed_data(), y.typed_data(), w.typed_data(), out->typed_data(), N, M, d, m, d, d, m, m,
                                  scale, nullptr, power, nconst, nullptr, nullptr, row_reduce, group, ws, wsb, stream));
}

// (gX, gs_rows) = VJP of out = K(X, Y) W with cotangent G: the backward rule of a jax.custom_vjp around jrk_kmatw_f32_ffi
// (INTEGRATION.md); dL/dY is the same handler called with (Y, X, G, W), dL/dW = jrk_kmatw_f32_ffi(Y, X, G).
static ffi::Error KmatwGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x,
                                ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> w, ffi::Buffer<ffi::F32> g, float scale,
                                float power, float nconst, ffi::ResultBuffer<ffi::F32> gx, ffi::ResultBuffer<ffi::F32> gs_rows) {
    const auto xd = x.dimensions(), yd = y.dimensions(), wd = w.dimensions(), gd = g.dimensions();
    if (xd.size() != 2 || yd.size() != 2 || wd.size() != 2 || gd.size() != 2 || xd[1] != yd[1] || wd[0] != yd[0] ||
        gd[0] != xd[0] || gd[1] != wd[1])
        return ffi::Error::InvalidArgument("X: N x d, Y: M x d, W: M x m, G: N x m expected");
    const int64_t N = xd[0], M = yd[0];
    const int d = (int)xd[1], m = (int)wd[1];
    const size_t wsb = jrk_grad_workspace_bytes(N, M, d);
    auto a = scratch.Allocate(wsb);
    if (!a.has_value()) return ffi::Error::Internal("jrk_kmatw_grad_f32: scratch allocation failed");
    return to_error(jrk_kmatw_grad_f32(x.typed_data(), y.typed_data(), w.typed_data(), g.typed_data(), gx->typed_data(),
                                       gs_rows->typed_data(), N, M, d, m, d, d, m, m, d, scale, power, nconst, *a, wsb, stream));
}

// K = kernel(X, Y) for any family of the distance front-end (JRK_KIND_*): PeriodicKernel / ThreshSpikeKernel
static ffi::Error GramKindImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y, int32_t kind, float p0,
                               float p1, float p2, float p3, float p4, ffi::ResultBuffer<ffi::F32> k) {
    const auto xd = x.dimensions(), yd = y.dimensions();
    if (xd.size() != 2 || yd.size() != 2 || xd[1] != yd[1]) return ffi::Error::InvalidArgument("X, Y must be N x d, M x d");
    const float kp[5] = {p0, p1, p2, p3, p4};
    return to_error(jrk_gram_kind_f32(x.typed_data(), y.typed_data(), k->typed_data(), xd[0], yd[0], (int)xd[1], xd[1], xd[1],
                                      yd[0], kind, kp, 5, nullptr, JRK_PATH_AUTO, nullptr, 0, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_kmatw_grad_f32_ffi, KmatwGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("scale")
                                  .Attr<float>("power")
                                  .Attr<float>("nconst")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_kind_f32_ffi, GramKindImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("kind")
                                  .Attr<float>("p0")
                                  .Attr<float>("p1")
                                  .Attr<float>("p2")
                                  .Attr<float>("p3")
                                  .Attr<float>("p4")
                                  .Ret<ffi::Buffer<ffi::F32>>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_f32_ffi, GramImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("scale")
                                  .Attr<float>("power")
                                  .Attr<float>("nconst")
                                  .Attr<int32_t>("path")
                                  .Ret<ffi::Buffer<ffi::F32>>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_kmatw_f32_ffi, KmatwImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("scale")
                                  .Attr<float>("power")
                                  .Attr<float>("nconst")
                                  .Attr<int32_t>("row_reduce")
                                  .Attr<int64_t>("group")
                                  .Ret<ffi::Buffer<ffi::F32>>(),
                              {ffi::Traits::kCmdBufferCompatible});
#endif  // __has_include("xla/ffi/api/ffi.h")

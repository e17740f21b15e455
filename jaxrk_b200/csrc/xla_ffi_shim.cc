// XLA FFI handler shim over the C ABI of include/jrk.h  (binding "L1a").
//
// jax / jaxlib are absent from the build image, so this file is NOT part of libjaxrk_b200.so by default: it is
// compiled by a maintainer who has jax installed,
//     g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude \
//         jaxrk_b200/csrc/xla_ffi_shim.cc -Ljaxrk_b200/lib -ljaxrk_b200 -o libjaxrk_b200_ffi.so
// and registered from Python as INTEGRATION.md shows.  Each handler forwards XLA-owned device buffers and the
// platform stream to one entry point of jrk.h; nothing synchronises, nothing is retained, errors become ffi::Error.
// In this repository it is compiled -- by tests/test_ffi_shim.py -- against a mock of the FFI header
// (tests/ffi_mock/xla/ffi/api/ffi.h) that checks every handler against its Bind() chain and against include/jrk.h.
//
// One handler per compute entry point of include/jrk.h (name + "_ffi").  Optional arrays of the C ABI (dim_scale,
// row_pref, col_pref, gs_rows) are ordinary buffer arguments here: an EMPTY buffer (0 elements) means "absent".
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include "../../include/jrk.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32Buf = ffi::Buffer<ffi::F32>;
using F32Res = ffi::ResultBuffer<ffi::F32>;

static ffi::Error to_error(int rc) {
    if (rc == JRK_OK) return ffi::Error::Success();
    char msg[512];
    jrk_last_error(msg, sizeof(msg));
    return rc == JRK_ERR_INVALID_ARG ? ffi::Error::InvalidArgument(msg) : ffi::Error::Internal(msg);
}
static const float* opt(const F32Buf& b) { return b.element_count() == 0 ? nullptr : b.typed_data(); }
static bool is_mat(const F32Buf& b) { return b.dimensions().size() == 2; }
static bool scratch_for(ffi::ScratchAllocator& scratch, size_t bytes, void** ws) {
    *ws = nullptr;
    if (!bytes) return true;
    auto a = scratch.Allocate(bytes);
    if (!a.has_value()) return false;
    *ws = *a;
    return true;
}

// K = GenGaussKernel(X, Y)   (jaxrk/kern/rbf.py:104-105); pass the same buffer twice for the self-Gram
static ffi::Error GramImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F32Buf x, F32Buf y, F32Buf dim_scale,
                           float scale, float power, float nconst, int32_t path, F32Res k) {
    if (!is_mat(x) || !is_mat(y) || x.dimensions()[1] != y.dimensions()[1]) return ffi::Error::InvalidArgument("X, Y must be N x d, M x d");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1];
    void* ws;
    const size_t wsb = jrk_gram_workspace_bytes(N, M, d, path);
    if (!scratch_for(scratch, wsb, &ws)) return ffi::Error::Internal("jrk_gram_f32: scratch allocation failed");
    return to_error(jrk_gram_f32(x.typed_data(), y.typed_data() == x.typed_data() ? nullptr : y.typed_data(), k->typed_data(), N, M,
                                 d, d, d, M, scale, opt(dim_scale), power, nconst, path, ws, wsb, stream));
}

// D = (scale ||ds x - ds y||)^power   (jaxrk/kern/util.py:107-116, jaxrk/utilities/eucldist.py:37-48)
static ffi::Error DistImpl(cudaStream_t stream, F32Buf x, F32Buf y, F32Buf dim_scale, float scale, float power, F32Res dist) {
    if (!is_mat(x) || !is_mat(y) || x.dimensions()[1] != y.dimensions()[1]) return ffi::Error::InvalidArgument("X, Y must be N x d, M x d");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1];
    return to_error(jrk_dist_f32(x.typed_data(), y.typed_data() == x.typed_data() ? nullptr : y.typed_data(), dist->typed_data(), N, M,
                                 d, d, d, M, scale, opt(dim_scale), power, stream));
}

// out[i] = scale sum_k |ds_k (x_ik - y_ik)|^power   (the diag=True branch, jaxrk/kern/util.py:108-113)
static ffi::Error DistDiagImpl(cudaStream_t stream, F32Buf x, F32Buf y, F32Buf dim_scale, float scale, float power, F32Res out) {
    if (!is_mat(x) || !is_mat(y) || x.dimensions()[0] != y.dimensions()[0] || x.dimensions()[1] != y.dimensions()[1])
        return ffi::Error::InvalidArgument("X and Y must both be N x d");
    const int d = (int)x.dimensions()[1];
    return to_error(jrk_dist_diag_f32(x.typed_data(), y.typed_data(), out->typed_data(), x.dimensions()[0], d, d, d, scale,
                                      opt(dim_scale), power, stream));
}

// out = K * row_mul * col_mul + row_add + col_add + add_const   (Prefactors / Center on a materialised Gram in one pass,
// jaxrk/reduce/base.py:84-101, 183-192); bind with input_output_aliases={0: 0} to run it in place
static ffi::Error GramAffineImpl(cudaStream_t stream, F32Buf k, F32Buf row_mul, F32Buf col_mul, F32Buf row_add, F32Buf col_add,
                                 float add_const, F32Res out) {
    if (!is_mat(k)) return ffi::Error::InvalidArgument("K must be N x M");
    const int64_t N = k.dimensions()[0], M = k.dimensions()[1];
    return to_error(jrk_gram_affine_f32(k.typed_data(), out->typed_data(), N, M, M, M, opt(row_mul), opt(col_mul), opt(row_add),
                                        opt(col_add), add_const, stream));
}

// out = row_reduce(diag(row_pref) K(X, Y) diag(col_pref) W)   (jaxrk/rkhs/vector.py:62-75)
static ffi::Error KmatwImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F32Buf x, F32Buf y, F32Buf w, F32Buf dim_scale,
                            F32Buf row_pref, F32Buf col_pref, float scale, float power, float nconst, int32_t row_reduce,
                            int64_t group, F32Res out) {
    if (!is_mat(x) || !is_mat(y) || !is_mat(w) || x.dimensions()[1] != y.dimensions()[1] || w.dimensions()[0] != y.dimensions()[0])
        return ffi::Error::InvalidArgument("X: N x d, Y: M x d, W: M x m expected");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1], m = (int)w.dimensions()[1];
    void* ws;
    const size_t wsb = jrk_kmatw_workspace_bytes(N, M, d, m, row_reduce, group);
    if (!scratch_for(scratch, wsb, &ws)) return ffi::Error::Internal("jrk_kmatw_f32: scratch allocation failed");
    return to_error(jrk_kmatw_f32(x.typed_data(), y.typed_data(), w.typed_data(), out->typed_data(), N, M, d, m, d, d, m, m, scale,
                                  opt(dim_scale), power, nconst, opt(row_pref), opt(col_pref), row_reduce, group, ws, wsb, stream));
}

// the same two operations for any kernel family of the distance front-end (JRK_KIND_*): PeriodicKernel /
// ThreshSpikeKernel (jaxrk/kern/rbf.py:107-206); kparams as five attributes (unused ones ignored)
static ffi::Error GramKindImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F32Buf x, F32Buf y, F32Buf dim_scale, int32_t kind,
                               float p0, float p1, float p2, float p3, float p4, int32_t path, F32Res k) {
    if (!is_mat(x) || !is_mat(y) || x.dimensions()[1] != y.dimensions()[1]) return ffi::Error::InvalidArgument("X, Y must be N x d, M x d");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1];
    const float kp[5] = {p0, p1, p2, p3, p4};
    void* ws;
    const size_t wsb = jrk_gram_workspace_bytes(N, M, d, path);
    if (!scratch_for(scratch, wsb, &ws)) return ffi::Error::Internal("jrk_gram_kind_f32: scratch allocation failed");
    return to_error(jrk_gram_kind_f32(x.typed_data(), y.typed_data() == x.typed_data() ? nullptr : y.typed_data(), k->typed_data(), N,
                                      M, d, d, d, M, kind, kp, 5, opt(dim_scale), path, ws, wsb, stream));
}

static ffi::Error KmatwKindImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F32Buf x, F32Buf y, F32Buf w, F32Buf dim_scale,
                                F32Buf row_pref, F32Buf col_pref, int32_t kind, float p0, float p1, float p2, float p3, float p4,
                                int32_t row_reduce, int64_t group, F32Res out) {
    if (!is_mat(x) || !is_mat(y) || !is_mat(w) || x.dimensions()[1] != y.dimensions()[1] || w.dimensions()[0] != y.dimensions()[0])
        return ffi::Error::InvalidArgument("X: N x d, Y: M x d, W: M x m expected");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1], m = (int)w.dimensions()[1];
    const float kp[5] = {p0, p1, p2, p3, p4};
    void* ws;
    const size_t wsb = jrk_kmatw_workspace_bytes(N, M, d, m, row_reduce, group);
    if (!scratch_for(scratch, wsb, &ws)) return ffi::Error::Internal("jrk_kmatw_kind_f32: scratch allocation failed");
    return to_error(jrk_kmatw_kind_f32(x.typed_data(), y.typed_data(), w.typed_data(), out->typed_data(), N, M, d, m, d, d, m, m, kind,
                                       kp, 5, opt(dim_scale), opt(row_pref), opt(col_pref), row_reduce, group, ws, wsb, stream));
}

// out[a, b] = sum over groups of row_pref k col_pref   (BalancedRed on both vectors, jaxrk/reduce/base.py:152-181)
static ffi::Error GramGroupsumImpl(cudaStream_t stream, F32Buf x, F32Buf y, F32Buf dim_scale, F32Buf row_pref, F32Buf col_pref,
                                   float scale, float power, float nconst, int64_t gx, int64_t gy, int32_t average, F32Res out) {
    if (!is_mat(x) || !is_mat(y) || x.dimensions()[1] != y.dimensions()[1]) return ffi::Error::InvalidArgument("X, Y must be N x d, M x d");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1];
    if (gy < 1) return ffi::Error::InvalidArgument("gy must be >= 1");
    return to_error(jrk_gram_groupsum_f32(x.typed_data(), y.typed_data() == x.typed_data() ? nullptr : y.typed_data(),
                                          out->typed_data(), N, M, d, d, d, M / gy, scale, opt(dim_scale), power, nconst,
                                          opt(row_pref), opt(col_pref), gx, gy, average, nullptr, 0, stream));
}

// (gX, gs_rows) = VJP of out = K(X, Y) W with cotangent G: the backward rule of a jax.custom_vjp around jrk_kmatw_f32_ffi
// (INTEGRATION.md); dL/dY is the same handler called with (Y, X, G, W), dL/dW = jrk_kmatw_f32_ffi(Y, X, G).
static ffi::Error KmatwGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F32Buf x, F32Buf y, F32Buf w, F32Buf g,
                                float scale, float power, float nconst, F32Res gx, F32Res gs_rows, F32Res gp_rows) {
    if (!is_mat(x) || !is_mat(y) || !is_mat(w) || !is_mat(g) || x.dimensions()[1] != y.dimensions()[1] ||
        w.dimensions()[0] != y.dimensions()[0] || g.dimensions()[0] != x.dimensions()[0] || g.dimensions()[1] != w.dimensions()[1])
        return ffi::Error::InvalidArgument("X: N x d, Y: M x d, W: M x m, G: N x m expected");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1], m = (int)w.dimensions()[1];
    void* ws;
    const size_t wsb = jrk_grad_workspace_bytes(N, M, d);
    if (!scratch_for(scratch, wsb, &ws)) return ffi::Error::Internal("jrk_kmatw_grad_f32: scratch allocation failed");
    return to_error(jrk_kmatw_grad_f32(x.typed_data(), y.typed_data(), w.typed_data(), g.typed_data(), gx->typed_data(),
                                       gs_rows->element_count() == 0 ? nullptr : gs_rows->typed_data(),
                                       gp_rows->element_count() == 0 ? nullptr : gp_rows->typed_data(), N, M, d, m, d, d, m, m, d,
                                       scale, power, nconst, ws, wsb, stream));
}

// (gX, gs_rows) = VJP of K = K(X, Y) with cotangent Gbar (N x M)
static ffi::Error GramGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F32Buf x, F32Buf y, F32Buf gbar, float scale,
                               float power, float nconst, int32_t gbar_transposed, F32Res gx, F32Res gs_rows, F32Res gp_rows) {
    const int r0 = gbar_transposed ? 1 : 0;      // which dimension of gbar runs along X
    if (!is_mat(x) || !is_mat(y) || !is_mat(gbar) || x.dimensions()[1] != y.dimensions()[1] ||
        gbar.dimensions()[r0] != x.dimensions()[0] || gbar.dimensions()[1 - r0] != y.dimensions()[0])
        return ffi::Error::InvalidArgument("X: N x d, Y: M x d, Gbar: N x M (or M x N when gbar_transposed) expected");
    const int64_t N = x.dimensions()[0], M = y.dimensions()[0];
    const int d = (int)x.dimensions()[1];
    void* ws;
    const size_t wsb = jrk_grad_workspace_bytes(N, M, d);
    if (!scratch_for(scratch, wsb, &ws)) return ffi::Error::Internal("jrk_gram_grad_f32: scratch allocation failed");
    return to_error(jrk_gram_grad_f32(x.typed_data(), y.typed_data(), gbar.typed_data(), gx->typed_data(),
                                      gs_rows->element_count() == 0 ? nullptr : gs_rows->typed_data(),
                                      gp_rows->element_count() == 0 ? nullptr : gp_rows->typed_data(), N, M, d, d, d,
                                      gbar_transposed ? N : M, d, gbar_transposed, scale,
                                      power, nconst, ws, wsb, stream));
}

#define JRK_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define JRK_STREAM_SCRATCH JRK_STREAM.Ctx<ffi::ScratchAllocator>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_f32_ffi, GramImpl,
                              JRK_STREAM_SCRATCH.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power").Attr<float>("nconst").Attr<int32_t>("path")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_dist_f32_ffi, DistImpl,
                              JRK_STREAM.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_dist_diag_f32_ffi, DistDiagImpl,
                              JRK_STREAM.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_affine_f32_ffi, GramAffineImpl,
                              JRK_STREAM.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("add_const")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_kmatw_f32_ffi, KmatwImpl,
                              JRK_STREAM_SCRATCH.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power").Attr<float>("nconst")
                                  .Attr<int32_t>("row_reduce").Attr<int64_t>("group")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_kind_f32_ffi, GramKindImpl,
                              JRK_STREAM_SCRATCH.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<int32_t>("kind").Attr<float>("p0").Attr<float>("p1").Attr<float>("p2").Attr<float>("p3")
                                  .Attr<float>("p4").Attr<int32_t>("path")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_kmatw_kind_f32_ffi, KmatwKindImpl,
                              JRK_STREAM_SCRATCH.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<int32_t>("kind").Attr<float>("p0").Attr<float>("p1").Attr<float>("p2").Attr<float>("p3")
                                  .Attr<float>("p4").Attr<int32_t>("row_reduce").Attr<int64_t>("group")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_groupsum_f32_ffi, GramGroupsumImpl,
                              JRK_STREAM.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power").Attr<float>("nconst")
                                  .Attr<int64_t>("gx").Attr<int64_t>("gy").Attr<int32_t>("average")
                                  .Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_kmatw_grad_f32_ffi, KmatwGradImpl,
                              JRK_STREAM_SCRATCH.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power").Attr<float>("nconst")
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});

XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_gram_grad_f32_ffi, GramGradImpl,
                              JRK_STREAM_SCRATCH.Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()
                                  .Attr<float>("scale").Attr<float>("power").Attr<float>("nconst").Attr<int32_t>("gbar_transposed")
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>(),
                              {ffi::Traits::kCmdBufferCompatible});
#endif  // __has_include("xla/ffi/api/ffi.h")

// jrk_kmatw_plan_*: the Y / W side of a fused K(X, Y) @ W kept on the device across applies.
//
// The reference evaluates FiniteVec.inner / FiniteOp.__matmul__ (jaxrk/rkhs/vector.py:62-75, jaxrk/rkhs/operator.py:33-56)
// against the SAME inspection points and the SAME linear map many times (one RKHS operator, many batches of test
// points); the jit cache is its only state.  Here the state is explicit: a plan owns device copies of Y, W (and the
// column prefactors / per-dimension scales) plus -- when the tensor-core path applies -- the packed tile images, scaled
// row norms and maxima of the Y side, so an apply is  H2D(X) + row norms of X + the main kernel + D2H(out).
// A plan is not thread-safe: one apply at a time (the host variant reuses the plan's staging buffers).
#include <algorithm>
#include <cmath>
#include <new>

#include "common.cuh"

namespace jrk {
int kmatw(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
          int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale, float power,
          float nconst, const float* row_pref, const float* col_pref, int row_reduce, int64_t group,
          const ExtraKind& ex, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t kmatw_workspace_bytes(int64_t N, int64_t M, int d, int m, int row_reduce, int64_t group);
bool kmatw_uses_tc(int64_t N, int64_t M, int d, int m, bool has_dim_scale, int row_reduce, const ExtraKind& ex);
size_t kmatw_tc_plan_prep_bytes(int64_t M, int d, int m);
size_t kmatw_tc_plan_apply_bytes(int64_t N, int64_t M, int d);
int kmatw_tc_plan_prepare(const float* Y, const float* W, int64_t M, int d, int m, int64_t ldy, int64_t ldw,
                          const float* col_pref, float scale, float power, float nconst, void* prep, unsigned* nmax4_scratch,
                          cudaStream_t st);
int kmatw_tc_plan_apply(const float* X, const float* Y, const float* W, const float* col_pref, const void* prep, float* out,
                        int64_t N, int64_t M, int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale,
                        float power, float nconst, const float* row_pref, int row_reduce, void* workspace,
                        size_t workspace_bytes, cudaStream_t st);
void keep_pool_memory(int device);
}  // namespace jrk

using namespace jrk;

struct jrk_kmatw_plan {
    int device = 0;
    int64_t M = 0;
    int d = 0, m = 0;
    float scale = 1.f, power = 2.f, nconst = 1.f;
    float *Y = nullptr, *W = nullptr, *col_pref = nullptr, *dim_scale = nullptr;   // owned device copies (dense: ld = d / m)
    void* prep = nullptr;                 // tensor-core path: per 16-column block {tile images | c|y|^2 | maxima}
    // staging of the host variant, grown on demand
    cudaStream_t stream = nullptr;
    float *dX = nullptr, *dOut = nullptr, *dRp = nullptr;
    int64_t cap_x = 0, cap_out = 0, cap_rp = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
};

namespace {
struct DeviceGuard {     // the host-buffer entry points run on `device` and leave the caller's current device alone
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
int grow(T*& p, int64_t& cap, int64_t want) {
    if (want <= cap) return JRK_OK;
    if (p) JRK_CUDA_OK(cudaFree(p));
    p = nullptr; cap = 0;
    JRK_CUDA_OK(cudaMalloc((void**)&p, (size_t)want * sizeof(T)));
    cap = want;
    return JRK_OK;
}

int copy_dense(float*& dst, const float* src, int64_t rows, int cols, int64_t ld, bool src_is_host, cudaStream_t st) {
    JRK_CUDA_OK(cudaMalloc((void**)&dst, (size_t)rows * cols * 4));
    JRK_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)cols * 4, src, (size_t)ld * 4, (size_t)cols * 4, (size_t)rows,
                                  src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    return JRK_OK;
}

int64_t out_rows_of(int64_t N, int row_reduce, int64_t group) {
    if (row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN) return 1;
    if (row_reduce == JRK_REDUCE_BALANCED || row_reduce == JRK_REDUCE_BALANCED_MEAN) return group > 0 ? N / group : 0;
    return N;
}

int check_apply(const jrk_kmatw_plan* p, const float* X, float* out, int64_t N, int row_reduce, int64_t group) {
    if (!p) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_apply: NULL plan");
    if (!X || !out) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_apply: NULL pointer");
    if (N < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_apply: N must be >= 1");
    if (row_reduce < JRK_REDUCE_NONE || row_reduce > JRK_REDUCE_BALANCED_MEAN)
        return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_apply: bad row_reduce %d", row_reduce);
    if ((row_reduce == JRK_REDUCE_BALANCED || row_reduce == JRK_REDUCE_BALANCED_MEAN) && (group < 1 || N % group != 0))
        return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_apply: group (%lld) must divide N (%lld)", (long long)group, (long long)N);
    return JRK_OK;
}
}  // namespace

extern "C" {

int jrk_kmatw_plan_create(jrk_kmatw_plan** plan, const float* Y, const float* W, int64_t M, int d, int m, int64_t ldy,
                          int64_t ldw, float scale, const float* dim_scale, float power, float nconst, const float* col_pref,
                          int src_is_host, int device, void* stream) {
    if (!plan) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_create: NULL plan pointer");
    *plan = nullptr;
    if (!Y || !W) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_create: Y and W must not be NULL");
    if (M < 1 || m < 1 || d < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_create: M, d, m must be >= 1");
    if (ldy < d || ldw < m) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_create: leading dimension too small");
    if (!(scale > 0.f) || !std::isfinite(scale) || !(power > 0.f) || power > 2.f || !std::isfinite(nconst))
        return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_create: need scale > 0, 0 < power <= 2, finite nconst");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(JRK_ERR_CUDA, "jrk_kmatw_plan_create: no CUDA device available; this library has no CPU fallback");
    }
    DeviceGuard guard(device);
    if (!guard.ok) return fail(JRK_ERR_CUDA, "jrk_kmatw_plan_create: cudaSetDevice(%d) failed", device);
    keep_pool_memory(device);
    jrk_kmatw_plan* p = new (std::nothrow) jrk_kmatw_plan();
    if (!p) return fail(JRK_ERR_CUDA, "jrk_kmatw_plan_create: out of host memory");
    p->device = device; p->M = M; p->d = d; p->m = m; p->scale = scale; p->power = power; p->nconst = nconst;
    int rc = JRK_OK;
    auto bail = [&](int code) { jrk_kmatw_plan_destroy(p); return code; };
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(JRK_ERR_CUDA, "jrk_kmatw_plan_create: stream"));
    cudaStream_t st = src_is_host ? p->stream : (cudaStream_t)stream;
    const bool host = src_is_host != 0;
    if ((rc = copy_dense(p->Y, Y, M, d, ldy, host, st))) return bail(rc);
    if ((rc = copy_dense(p->W, W, M, m, ldw, host, st))) return bail(rc);
    if (col_pref && (rc = copy_dense(p->col_pref, col_pref, M, 1, 1, host, st))) return bail(rc);
    if (dim_scale && (rc = copy_dense(p->dim_scale, dim_scale, 1, d, d, host, st))) return bail(rc);
    // Y side of the tensor-core path (kmatw_tc.cu: d <= 32, no per-dimension scale, enough columns)
    if (!dim_scale && d <= 32 && M >= 4096 && jrk_option(JRK_OPT_KMATW_TC) != 0) {
        const size_t pb = kmatw_tc_plan_prep_bytes(M, d, m);
        unsigned* scratch = nullptr;
        if (cudaMalloc(&p->prep, pb) != cudaSuccess || cudaMalloc((void**)&scratch, 256) != cudaSuccess) {
            cudaGetLastError();
            return bail(fail(JRK_ERR_CUDA, "jrk_kmatw_plan_create: cudaMalloc of %zu bytes failed", pb));
        }
        rc = kmatw_tc_plan_prepare(p->Y, p->W, M, d, m, d, m, p->col_pref, scale, power, nconst, p->prep, scratch, st);
        cudaStreamSynchronize(st);
        cudaFree(scratch);
        if (rc) return bail(rc);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return bail(fail(JRK_ERR_CUDA, "jrk_kmatw_plan_create: %s", cudaGetErrorString(cudaGetLastError())));
    *plan = p;
    return JRK_OK;
}

int jrk_kmatw_plan_destroy(jrk_kmatw_plan* p) {
    if (!p) return JRK_OK;
    DeviceGuard guard(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->Y); cudaFree(p->W); cudaFree(p->col_pref); cudaFree(p->dim_scale); cudaFree(p->prep);
    cudaFree(p->dX); cudaFree(p->dOut); cudaFree(p->dRp); cudaFree(p->ws);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    cudaGetLastError();
    return JRK_OK;
}

size_t jrk_kmatw_plan_workspace_bytes(const jrk_kmatw_plan* p, int64_t N, int row_reduce, int64_t group) {
    if (!p || N < 1) return 0;
    size_t b = kmatw_workspace_bytes(N, p->M, p->d, p->m, row_reduce, group);
    if (p->prep) b = std::max(b, kmatw_tc_plan_apply_bytes(N, p->M, p->d));
    return b;
}

int jrk_kmatw_plan_apply(const jrk_kmatw_plan* p, const float* X, float* out, int64_t N, int64_t ldx, int64_t ldo,
                         const float* row_pref, int row_reduce, int64_t group, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (int rc = check_apply(p, X, out, N, row_reduce, group)) return rc;
    if (ldx < p->d || ldo < p->m) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_plan_apply: leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (p->prep && kmatw_uses_tc(N, p->M, p->d, p->m, false, row_reduce, ExtraKind()))
        return kmatw_tc_plan_apply(X, p->Y, p->W, p->col_pref, p->prep, out, N, p->M, p->d, p->m, ldx, p->d, p->m, ldo, p->scale,
                                   p->power, p->nconst, row_pref, row_reduce, workspace, workspace_bytes, st);
    return kmatw(X, p->Y, p->W, out, N, p->M, p->d, p->m, ldx, p->d, p->m, ldo, p->scale, p->dim_scale, p->power, p->nconst,
                 row_pref, p->col_pref, row_reduce, group, ExtraKind(), workspace, workspace_bytes, st);
}

int jrk_kmatw_plan_apply_host(jrk_kmatw_plan* p, const float* X, float* out, int64_t N, const float* row_pref, int row_reduce,
                              int64_t group) {
    if (int rc = check_apply(p, X, out, N, row_reduce, group)) return rc;
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(JRK_ERR_CUDA, "jrk_kmatw_plan_apply_host: cudaSetDevice(%d) failed", p->device);
    const int64_t orows = out_rows_of(N, row_reduce, group);
    if (int rc = grow(p->dX, p->cap_x, N * p->d)) return rc;
    if (int rc = grow(p->dOut, p->cap_out, orows * p->m)) return rc;
    if (row_pref) { if (int rc = grow(p->dRp, p->cap_rp, N)) return rc; }
    const size_t wsb = jrk_kmatw_plan_workspace_bytes(p, N, row_reduce, group);
    if (wsb > p->ws_bytes) {
        if (p->ws) JRK_CUDA_OK(cudaFree(p->ws));
        p->ws = nullptr; p->ws_bytes = 0;
        JRK_CUDA_OK(cudaMalloc(&p->ws, wsb));
        p->ws_bytes = wsb;
    }
    cudaStream_t st = p->stream;
    JRK_CUDA_OK(cudaMemcpyAsync(p->dX, X, (size_t)N * p->d * 4, cudaMemcpyHostToDevice, st));
    if (row_pref) JRK_CUDA_OK(cudaMemcpyAsync(p->dRp, row_pref, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    if (int rc = jrk_kmatw_plan_apply(p, p->dX, p->dOut, N, p->d, p->m, row_pref ? p->dRp : nullptr, row_reduce, group, p->ws,
                                      p->ws_bytes, st))
        return rc;
    JRK_CUDA_OK(cudaMemcpyAsync(out, p->dOut, (size_t)orows * p->m * 4, cudaMemcpyDeviceToHost, st));
    JRK_CUDA_OK(cudaStreamSynchronize(st));
    return JRK_OK;
}

}  // extern "C"

// extern "C" entry points declared in include/jrk.h: argument validation, path
// selection, host-buffer (end-to-end) variants.  No torch / XLA types here.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace jrk {

thread_local char g_last_error[512] = {0};
int kgrad(int mode, const float* X, const float* Y, const float* W, const float* G, float* gX, float* gs_rows, float* gp_rows, int64_t N,
          int64_t M, int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldg, int64_t ldgx, float scale,
          float power, float nconst, int gtrans, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t kgrad_workspace_bytes(int64_t N, int64_t M, int d);
int gram_affine(const float* K, float* out, int64_t N, int64_t M, int64_t ldk, int64_t ldo, const float* row_mul,
                const float* col_mul, const float* row_add, const float* col_add, float add_const, cudaStream_t st);
int dist_diag(const float* X, const float* Y, float* out, int64_t N, int d, int64_t ldx, int64_t ldy, float scale,
              const float* dim_scale, float power, cudaStream_t st);
std::atomic<int64_t> g_launches{0};

// kernels (one translation unit each)
int gram_direct(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, int out_kind,
                const ExtraKind& ex, cudaStream_t st);
int gram_tf32x3(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, void* workspace,
                size_t workspace_bytes, cudaStream_t st);
size_t gram_tf32x3_workspace_bytes(int64_t N, int64_t M, int d);
int kmatw(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
          int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale, float power,
          float nconst, const float* row_pref, const float* col_pref, int row_reduce, int64_t group,
          const ExtraKind& ex, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t kmatw_workspace_bytes(int64_t N, int64_t M, int d, int m, int row_reduce, int64_t group);
bool kmatw_uses_tc(int64_t N, int64_t M, int d, int m, bool has_dim_scale, int row_reduce, const ExtraKind& ex);
int gram_groupsum(const float* X, const float* Y, float* out, int64_t N, int64_t M, int d, int64_t ldx,
                  int64_t ldy, int64_t ldo, float scale, const float* dim_scale, float power, float nconst,
                  const float* row_pref, const float* col_pref, int64_t gx, int64_t gy, int average,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);

// ---- process-wide options (jrk_set_option) ----
static std::atomic<int> g_options[JRK_OPT_COUNT_];
static std::atomic<int> g_options_init{0};
static const int kOptionDefault[JRK_OPT_COUNT_] = {1, 1, 1, 4, 1, 0};
static const char* const kOptionEnv[JRK_OPT_COUNT_] = {"JRK_KMATW_TC", "JRK_TC_FAST", "JRK_TC_V2", "JRK_TC2_EPQ", "JRK_GRAM_H16", "JRK_TC2_TOKEN"};
static void options_init() {
    // the environment variable of the same name gives the INITIAL value, once per process
    int expected = 0;
    if (!g_options_init.compare_exchange_strong(expected, 1)) {
        while (g_options_init.load() != 2) {}
        return;
    }
    for (int i = 0; i < JRK_OPT_COUNT_; ++i) {
        const char* e = getenv(kOptionEnv[i]);
        g_options[i].store(e && *e ? atoi(e) : kOptionDefault[i]);
    }
    g_options_init.store(2);
}
int jrk_option(int opt) {
    if (g_options_init.load(std::memory_order_acquire) != 2) options_init();
    return (opt >= 0 && opt < JRK_OPT_COUNT_) ? g_options[opt].load(std::memory_order_relaxed) : 0;
}

// Stream-ordered scratch comes from the device's default memory pool, which by default gives everything back to the
// driver at the next synchronisation: every call would then pay for fresh allocations (hundreds of MB at configs[2];
// this was the erratic 390 .. 790 ms of round 1's end-to-end K@W).  Keep the pool's memory instead.
void keep_pool_memory(int device) {
    static std::atomic<unsigned long long> done{0};
    if (device < 0 || device >= 64 || (done.load() >> device) & 1ull) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done.fetch_or(1ull << device);
}

static int check_kernel_params(const char* fn, int d, float scale, float power, float nconst, bool any_power = false) {
    if (d < 1) return fail(JRK_ERR_INVALID_ARG, "%s: d must be >= 1 (got %d)", fn, d);
    if (!(scale > 0.f) || !std::isfinite(scale))
        return fail(JRK_ERR_INVALID_ARG, "%s: scale must be positive and finite (got %g)", fn, (double)scale);
    // GenGaussKernel.make asserts 0 < shape <= 2 (jaxrk/kern/rbf.py:71); ThreshSpikeKernel takes any positive shape
    if (!(power > 0.f) || (power > 2.f && !any_power))
        return fail(JRK_ERR_INVALID_ARG, "%s: power must be in (0, 2] (got %g)", fn, (double)power);
    if (!std::isfinite(nconst)) return fail(JRK_ERR_INVALID_ARG, "%s: nconst must be finite", fn);
    return JRK_OK;
}

static int have_device(const char* fn) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(JRK_ERR_CUDA, "%s: no CUDA device available (%s); this library has no CPU fallback", fn,
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return JRK_OK;
}

// AUTO: the tcgen05 path for d >= 64 (where the distance is a dense contraction), and -- measured on B200,
// tools/small_d_paths.py -- also for 12 <= d < 64 once the Gram is large (>= 2^24 entries): it stays HBM-write-bound
// (3.1 ms at 65536^2) while direct differencing is FP32-bound (5.1 / 6.0 / 10.3 / 14.3 ms at d = 12 / 16 / 32 / 48).
// Small problems and d < 12 use direct differencing (one launch, exact diagonal; at d <= 4 it writes 6.6 TB/s, and the
// tensor path would spend its time in the near-coincident refinement that low-dimensional data triggers all the time).
static int resolve_gram_path(int path, int d, int64_t N, int64_t M) {
    if (path != JRK_PATH_AUTO) return path;
    if (d >= 64) return JRK_PATH_TF32X3;
    if (d >= 12 && N * M >= ((int64_t)1 << 24)) return JRK_PATH_TF32X3;
    return JRK_PATH_DIRECT;
}

}  // namespace jrk

using namespace jrk;

extern "C" {

int jrk_abi_version(void) { return JRK_ABI_VERSION; }

int jrk_last_error(char* buf, size_t n) {
    if (!buf || n == 0) return JRK_ERR_INVALID_ARG;
    std::strncpy(buf, g_last_error, n - 1);
    buf[n - 1] = 0;
    return JRK_OK;
}

int jrk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t jrk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int jrk_set_option(int option, int value) {
    if (option < 0 || option >= JRK_OPT_COUNT_) return fail(JRK_ERR_INVALID_ARG, "jrk_set_option: unknown option %d", option);
    if (option == JRK_OPT_TC2_EPQ && value != 3 && value != 4)
        return fail(JRK_ERR_INVALID_ARG, "jrk_set_option: JRK_OPT_TC2_EPQ must be 3 or 4");
    if (option == JRK_OPT_TC2_TOKEN && (value < 0 || value > 3))
        return fail(JRK_ERR_INVALID_ARG, "jrk_set_option: JRK_OPT_TC2_TOKEN must be 0 .. 3");
    jrk_option(option);   // make sure the defaults are in place
    g_options[option].store(value);
    return JRK_OK;
}

int jrk_get_option(int option) { return jrk_option(option); }

size_t jrk_gram_workspace_bytes(int64_t N, int64_t M, int d, int path) {
    if (resolve_gram_path(path, d, N, M) == JRK_PATH_TF32X3) return gram_tf32x3_workspace_bytes(N, M, d);
    return 0;
}

static int gram_any(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                    int64_t ldk, float scale, const float* dim_scale, float power, float nconst, int path,
                    const ExtraKind& ex, void* workspace, size_t workspace_bytes, void* stream) {
    if (N < 0 || M < 0) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_f32: negative size");
    if (!Y) { Y = X; M = N; ldy = ldx; }
    if (N > 0 && M > 0 && (!X || !K)) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_f32: X and K must not be NULL");
    if (int rc = check_kernel_params("jrk_gram_f32", d, scale, power, nconst, ex.kind != 0)) return rc;
    if (ldx < d || ldy < d || ldk < M) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_f32: leading dimension too small");
    if (path < JRK_PATH_AUTO || path > JRK_PATH_TF32X3) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_f32: bad path %d", path);
    if (N == 0 || M == 0) return JRK_OK;
    if (int rc = have_device("jrk_gram_f32")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (ex.kind == 0 && resolve_gram_path(path, d, N, M) == JRK_PATH_TF32X3)
        return gram_tf32x3(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, workspace,
                           workspace_bytes, st);
    return gram_direct(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, 0, ex, st);
}

int jrk_gram_f32(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                 int64_t ldk, float scale, const float* dim_scale, float power, float nconst, int path,
                 void* workspace, size_t workspace_bytes, void* stream) {
    return gram_any(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, path, ExtraKind(), workspace,
                    workspace_bytes, stream);
}

int jrk_dist_f32(const float* X, const float* Y, float* D, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                 int64_t ldd, float scale, const float* dim_scale, float power, void* stream) {
    if (N < 0 || M < 0 || d < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_f32: bad size");
    if (!Y) { Y = X; M = N; ldy = ldx; }
    if (N > 0 && M > 0 && (!X || !D)) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_f32: X and D must not be NULL");
    if (!(scale > 0.f) || !(power > 0.f)) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_f32: scale and power must be positive");
    if (ldx < d || ldy < d || ldd < M) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_f32: leading dimension too small");
    if (N == 0 || M == 0) return JRK_OK;
    if (int rc = have_device("jrk_dist_f32")) return rc;
    return gram_direct(X, Y, D, N, M, d, ldx, ldy, ldd, scale, dim_scale, power, 1.f, 1, ExtraKind(), (cudaStream_t)stream);
}

int jrk_dist_diag_f32(const float* X, const float* Y, float* out, int64_t N, int d, int64_t ldx, int64_t ldy,
                      float scale, const float* dim_scale, float power, void* stream) {
    if (N < 0 || d < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_diag_f32: bad size");
    if (N > 0 && (!X || !Y || !out)) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_diag_f32: NULL pointer");
    if (!(scale > 0.f) || !(power > 0.f)) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_diag_f32: scale and power must be positive");
    if (ldx < d || ldy < d) return fail(JRK_ERR_INVALID_ARG, "jrk_dist_diag_f32: leading dimension too small");
    if (N == 0) return JRK_OK;
    if (int rc = have_device("jrk_dist_diag_f32")) return rc;
    return dist_diag(X, Y, out, N, d, ldx, ldy, scale, dim_scale, power, (cudaStream_t)stream);
}

int jrk_gram_affine_f32(const float* K, float* out, int64_t N, int64_t M, int64_t ldk, int64_t ldo, const float* row_mul,
                        const float* col_mul, const float* row_add, const float* col_add, float add_const, void* stream) {
    if (N < 0 || M < 0) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_affine_f32: bad size");
    if (N == 0 || M == 0) return JRK_OK;
    if (!K || !out) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_affine_f32: NULL pointer");
    if (ldk < M || ldo < M) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_affine_f32: leading dimension too small");
    if (int rc = have_device("jrk_gram_affine_f32")) return rc;
    return gram_affine(K, out, N, M, ldk, ldo, row_mul, col_mul, row_add, col_add, add_const, (cudaStream_t)stream);
}

static int kmatw_any(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
                     int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale, float power,
                     float nconst, const float* row_pref, const float* col_pref, int row_reduce, int64_t group,
                     const ExtraKind& ex, void* workspace, size_t workspace_bytes, void* stream);

// kparams -> (scale, power, nconst) of the shared launchers + the sibling-kernel selector
static int parse_kind(const char* fn, int kind, const float* kp, int n, float& scale, float& power, float& nconst,
                      ExtraKind& ex) {
    if (!kp) return fail(JRK_ERR_INVALID_ARG, "%s: kparams is NULL", fn);
    ex = ExtraKind();
    if (kind == JRK_KIND_GENGAUSS) {
        if (n < 3) return fail(JRK_ERR_INVALID_ARG, "%s: GENGAUSS needs {scale, power, nconst}", fn);
        scale = kp[0]; power = kp[1]; nconst = kp[2];
    } else if (kind == JRK_KIND_PERIODIC) {
        if (n < 2) return fail(JRK_ERR_INVALID_ARG, "%s: PERIODIC needs {1/period, length_scale}", fn);
        if (!(kp[0] > 0.f) || !(kp[1] > 0.f))   // PeriodicKernel.__init__ asserts both positive (jaxrk/kern/rbf.py:118)
            return fail(JRK_ERR_INVALID_ARG, "%s: period and length_scale must be positive", fn);
        scale = kp[0]; power = 1.f; nconst = 1.f;
        ex.kind = kind; ex.a0 = 1.f / kp[1];
    } else if (kind == JRK_KIND_THRESHSPIKE) {
        if (n < 5) return fail(JRK_ERR_INVALID_ARG, "%s: THRESHSPIKE needs {scale, power, spike, non_spike, threshold}", fn);
        // ThreshSpikeKernel.__init__ (jaxrk/kern/rbf.py:166-168)
        if (!(kp[2] > 0.f) || !(std::fabs(kp[3]) < kp[2]) || !(kp[4] >= 0.f))
            return fail(JRK_ERR_INVALID_ARG, "%s: need spike > 0, |non_spike| < spike, threshold >= 0", fn);
        scale = kp[0]; power = kp[1]; nconst = 1.f;
        if (!(kp[0] > 0.f) || !(kp[1] > 0.f)) return fail(JRK_ERR_INVALID_ARG, "%s: scale and power must be positive", fn);
        const double root = std::pow((double)kp[4], 1.0 / (double)kp[1]) / (double)kp[0];   // threshold on r
        ex.kind = kind; ex.a0 = (float)(root * root); ex.a1 = kp[2]; ex.a2 = kp[3];
    } else {
        return fail(JRK_ERR_INVALID_ARG, "%s: unknown kernel kind %d", fn, kind);
    }
    return JRK_OK;
}

int jrk_gram_kind_f32(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                      int64_t ldk, int kind, const float* kparams, int n_kparams, const float* dim_scale, int path,
                      void* workspace, size_t workspace_bytes, void* stream) {
    float scale, power, nconst;
    ExtraKind ex;
    if (int rc = parse_kind("jrk_gram_kind_f32", kind, kparams, n_kparams, scale, power, nconst, ex)) return rc;
    if (ex.kind == 0)
        return jrk_gram_f32(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, path, workspace,
                            workspace_bytes, stream);
    if (path == JRK_PATH_TF32X3) return fail(JRK_ERR_UNSUPPORTED, "jrk_gram_kind_f32: sibling kernels use the direct path");
    return gram_any(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, JRK_PATH_DIRECT, ex, workspace,
                    workspace_bytes, stream);
}

int jrk_kmatw_kind_f32(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
                       int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, int kind, const float* kparams,
                       int n_kparams, const float* dim_scale, const float* row_pref, const float* col_pref,
                       int row_reduce, int64_t group, void* workspace, size_t workspace_bytes, void* stream) {
    float scale, power, nconst;
    ExtraKind ex;
    if (int rc = parse_kind("jrk_kmatw_kind_f32", kind, kparams, n_kparams, scale, power, nconst, ex)) return rc;
    return kmatw_any(X, Y, W, out, N, M, d, m, ldx, ldy, ldw, ldo, scale, dim_scale, power, nconst, row_pref, col_pref,
                     row_reduce, group, ex, workspace, workspace_bytes, stream);
}

size_t jrk_grad_workspace_bytes(int64_t N, int64_t M, int d) { return kgrad_workspace_bytes(N, M, d); }

int jrk_kmatw_grad_f32(const float* X, const float* Y, const float* W, const float* G, float* gX, float* gs_rows,
                       float* gp_rows, int64_t N, int64_t M, int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldg,
                       int64_t ldgx, float scale, float power, float nconst, void* workspace, size_t workspace_bytes,
                       void* stream) {
    if (!X || !Y || !W || !G || !gX) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_grad_f32: NULL pointer");
    if (N < 1 || M < 1 || m < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_grad_f32: N, M, m must be >= 1");
    if (int rc = check_kernel_params("jrk_kmatw_grad_f32", d, scale, power, nconst)) return rc;
    if (ldx < d || ldy < d || ldw < m || ldg < m || ldgx < d)
        return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_grad_f32: leading dimension too small");
    if (int rc = have_device("jrk_kmatw_grad_f32")) return rc;
    return kgrad(0, X, Y, W, G, gX, gs_rows, gp_rows, N, M, d, m, ldx, ldy, ldw, ldg, ldgx, scale, power, nconst, 0, workspace,
                 workspace_bytes, (cudaStream_t)stream);
}

int jrk_gram_grad_f32(const float* X, const float* Y, const float* Gbar, float* gX, float* gs_rows, float* gp_rows, int64_t N, int64_t M,
                      int d, int64_t ldx, int64_t ldy, int64_t ldgb, int64_t ldgx, int gbar_transposed, float scale, float power,
                      float nconst, void* workspace, size_t workspace_bytes, void* stream) {
    if (!X || !Y || !Gbar || !gX) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_grad_f32: NULL pointer");
    if (N < 1 || M < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_grad_f32: N, M must be >= 1");
    if (int rc = check_kernel_params("jrk_gram_grad_f32", d, scale, power, nconst)) return rc;
    if (ldx < d || ldy < d || ldgb < (gbar_transposed ? N : M) || ldgx < d)
        return fail(JRK_ERR_INVALID_ARG, "jrk_gram_grad_f32: leading dimension too small");
    if (int rc = have_device("jrk_gram_grad_f32")) return rc;
    return kgrad(1, X, Y, nullptr, Gbar, gX, gs_rows, gp_rows, N, M, d, 0, ldx, ldy, 0, ldgb, ldgx, scale, power, nconst,
                 gbar_transposed != 0, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t jrk_kmatw_workspace_bytes(int64_t N, int64_t M, int d, int m, int row_reduce, int64_t group) {
    return kmatw_workspace_bytes(N, M, d, m, row_reduce, group);
}

int jrk_kmatw_uses_tensor_cores(int64_t N, int64_t M, int d, int m, int has_dim_scale, int row_reduce) {
    return kmatw_uses_tc(N, M, d, m, has_dim_scale != 0, row_reduce, ExtraKind()) ? 1 : 0;
}

static int check_kmatw(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d,
                       int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, float power,
                       float nconst, int row_reduce, int64_t group, bool any_power = false) {
    if (!X || !Y || !W || !out) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_f32: NULL pointer");
    if (N < 1 || M < 1 || m < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_f32: N, M, m must be >= 1");
    if (int rc = check_kernel_params("jrk_kmatw_f32", d, scale, power, nconst, any_power)) return rc;
    if (ldx < d || ldy < d || ldw < m || ldo < m) return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_f32: leading dimension too small");
    if (row_reduce < JRK_REDUCE_NONE || row_reduce > JRK_REDUCE_BALANCED_MEAN)
        return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_f32: bad row_reduce %d", row_reduce);
    if (row_reduce == JRK_REDUCE_BALANCED || row_reduce == JRK_REDUCE_BALANCED_MEAN) {
        // BalancedRed.new_len asserts original_len % points_per_split == 0 (jaxrk/reduce/base.py:179-181)
        if (group < 1 || N % group != 0)
            return fail(JRK_ERR_INVALID_ARG, "jrk_kmatw_f32: group (%lld) must divide N (%lld)", (long long)group, (long long)N);
    }
    return JRK_OK;
}

int jrk_kmatw_f32(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
                  int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale,
                  float power, float nconst, const float* row_pref, const float* col_pref, int row_reduce,
                  int64_t group, void* workspace, size_t workspace_bytes, void* stream) {
    return kmatw_any(X, Y, W, out, N, M, d, m, ldx, ldy, ldw, ldo, scale, dim_scale, power, nconst, row_pref, col_pref,
                     row_reduce, group, ExtraKind(), workspace, workspace_bytes, stream);
}

}  // extern "C"

static int kmatw_any(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
                     int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale, float power,
                     float nconst, const float* row_pref, const float* col_pref, int row_reduce, int64_t group,
                     const ExtraKind& ex, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_kmatw(X, Y, W, out, N, M, d, m, ldx, ldy, ldw, ldo, scale, power, nconst, row_reduce, group, ex.kind != 0))
        return rc;
    if (int rc = have_device("jrk_kmatw_f32")) return rc;
    return kmatw(X, Y, W, out, N, M, d, m, ldx, ldy, ldw, ldo, scale, dim_scale, power, nconst, row_pref, col_pref,
                 row_reduce, group, ex, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" {

int jrk_gram_groupsum_f32(const float* X, const float* Y, float* out, int64_t N, int64_t M, int d, int64_t ldx,
                          int64_t ldy, int64_t ldo, float scale, const float* dim_scale, float power, float nconst,
                          const float* row_pref, const float* col_pref, int64_t gx, int64_t gy, int average,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!X || !out) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_groupsum_f32: NULL pointer");
    if (!Y) { Y = X; M = N; ldy = ldx; }
    if (N < 1 || M < 1) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_groupsum_f32: N, M must be >= 1");
    if (int rc = check_kernel_params("jrk_gram_groupsum_f32", d, scale, power, nconst)) return rc;
    if (gx < 1 || gy < 1 || N % gx || M % gy)
        return fail(JRK_ERR_INVALID_ARG, "jrk_gram_groupsum_f32: group sizes must divide N and M");
    if (ldx < d || ldy < d || ldo < M / gy) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_groupsum_f32: leading dimension too small");
    if (int rc = have_device("jrk_gram_groupsum_f32")) return rc;
    return gram_groupsum(X, Y, out, N, M, d, ldx, ldy, ldo, scale, dim_scale, power, nconst, row_pref, col_pref, gx,
                         gy, average, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------
// host-buffer variants (end to end: H2D, compute, D2H inside the call)
// ---------------------------------------------------------------------------------
}  // extern "C"

namespace {
// Device buffers of the host-buffer entry points: stream-ordered allocations from the device's default pool, whose memory
// keep_pool_memory() retains across calls -- a call in steady state makes no cudaMalloc / cudaFree (those cost 0.2 .. 1.2 s
// in four of ten calls on one box: profiles/README.md).  Freed in stream order by the destructor; the entry points
// synchronise their streams before they return.
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
    }
    template <typename T> T* as() { return (T*)p; }
};
struct DeviceGuard {     // the host-buffer entry points run on `device` and leave the caller's current device alone
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
        cudaGetLastError();
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct StreamGuard {
    cudaStream_t s = nullptr;
    ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};
// Declared AFTER the DevBufs of an entry point, so destroyed before them: whatever path leaves the function, no stream
// still touches a buffer when it goes back to the pool.
struct SyncOnExit {
    cudaStream_t a = nullptr, b = nullptr;
    ~SyncOnExit() {
        if (a) cudaStreamSynchronize(a);
        if (b) cudaStreamSynchronize(b);
    }
};
struct EventGuard {
    cudaEvent_t e = nullptr;
    ~EventGuard() { if (e) cudaEventDestroy(e); }
};
}  // namespace

extern "C" {

int jrk_gram_f32_host(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, float scale,
                      const float* dim_scale, float power, float nconst, int path, int device) {
    if (!X || !K) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_f32_host: X and K must not be NULL");
    const bool self = (!Y || Y == X);
    if (self) { Y = X; M = N; }
    if (N < 0 || M < 0) return fail(JRK_ERR_INVALID_ARG, "jrk_gram_f32_host: negative size");
    if (int rc = check_kernel_params("jrk_gram_f32_host", d, scale, power, nconst)) return rc;
    if (N == 0 || M == 0) return JRK_OK;
    if (int rc = have_device("jrk_gram_f32_host")) return rc;
    DeviceGuard dev_guard(device);
    if (!dev_guard.ok) return fail(JRK_ERR_CUDA, "jrk_gram_f32_host: cudaSetDevice(%d) failed", device);
    keep_pool_memory(device);

    StreamGuard sc, sd;  // compute / download
    JRK_CUDA_OK(cudaStreamCreateWithFlags(&sc.s, cudaStreamNonBlocking));
    JRK_CUDA_OK(cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking));
    DevBuf dX, dY, dS, dK[2], dW;
    SyncOnExit sync_on_exit{sc.s, sd.s};
    JRK_CUDA_OK(dX.alloc((size_t)N * d * 4, sc.s));
    JRK_CUDA_OK(cudaMemcpyAsync(dX.p, X, (size_t)N * d * 4, cudaMemcpyHostToDevice, sc.s));
    const float* dYp = dX.as<float>();
    if (!self) {
        JRK_CUDA_OK(dY.alloc((size_t)M * d * 4, sc.s));
        JRK_CUDA_OK(cudaMemcpyAsync(dY.p, Y, (size_t)M * d * 4, cudaMemcpyHostToDevice, sc.s));
        dYp = dY.as<float>();
    }
    const float* dSp = nullptr;
    if (dim_scale) {
        JRK_CUDA_OK(dS.alloc((size_t)d * 4, sc.s));
        JRK_CUDA_OK(cudaMemcpyAsync(dS.p, dim_scale, (size_t)d * 4, cudaMemcpyHostToDevice, sc.s));
        dSp = dS.as<float>();
    }
    // row blocks of ~256 MiB: block b+1 is computed while block b is copied out
    int64_t B = std::max<int64_t>(128, ((int64_t)256 << 20) / (M * 4) / 128 * 128);
    B = std::min(B, (N + 127) / 128 * 128);
    JRK_CUDA_OK(dK[0].alloc((size_t)B * M * 4, sc.s));
    JRK_CUDA_OK(dK[1].alloc((size_t)B * M * 4, sc.s));
    const size_t wsb = jrk_gram_workspace_bytes(B, M, d, path);
    if (wsb) JRK_CUDA_OK(dW.alloc(wsb, sc.s));
    EventGuard done[2], freed[2];
    for (int i = 0; i < 2; ++i) {
        JRK_CUDA_OK(cudaEventCreateWithFlags(&done[i].e, cudaEventDisableTiming));
        JRK_CUDA_OK(cudaEventCreateWithFlags(&freed[i].e, cudaEventDisableTiming));
    }
    int b = 0;
    for (int64_t r = 0; r < N; r += B, ++b) {
        const int64_t n = std::min(B, N - r);
        const int s = b & 1;
        if (b >= 2) JRK_CUDA_OK(cudaStreamWaitEvent(sc.s, freed[s].e, 0));
        int rc = jrk_gram_f32(dX.as<float>() + r * d, dYp, dK[s].as<float>(), n, M, d, d, d, M, scale, dSp, power,
                              nconst, path, dW.p, wsb, sc.s);
        if (rc) return rc;
        JRK_CUDA_OK(cudaEventRecord(done[s].e, sc.s));
        JRK_CUDA_OK(cudaStreamWaitEvent(sd.s, done[s].e, 0));
        JRK_CUDA_OK(cudaMemcpyAsync(K + r * M, dK[s].p, (size_t)n * M * 4, cudaMemcpyDeviceToHost, sd.s));
        JRK_CUDA_OK(cudaEventRecord(freed[s].e, sd.s));
    }
    JRK_CUDA_OK(cudaStreamSynchronize(sc.s));
    JRK_CUDA_OK(cudaStreamSynchronize(sd.s));
    return JRK_OK;
}

int jrk_kmatw_f32_host(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d,
                       int m, float scale, const float* dim_scale, float power, float nconst, const float* row_pref,
                       const float* col_pref, int row_reduce, int64_t group, int device) {
    if (int rc = check_kmatw(X, Y, W, out, N, M, d, m, d, d, m, m, scale, power, nconst, row_reduce, group)) return rc;
    if (int rc = have_device("jrk_kmatw_f32_host")) return rc;
    DeviceGuard dev_guard(device);
    if (!dev_guard.ok) return fail(JRK_ERR_CUDA, "jrk_kmatw_f32_host: cudaSetDevice(%d) failed", device);
    keep_pool_memory(device);
    StreamGuard sc;
    JRK_CUDA_OK(cudaStreamCreateWithFlags(&sc.s, cudaStreamNonBlocking));
    int64_t out_rows = N;
    if (row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN) out_rows = 1;
    else if (row_reduce != JRK_REDUCE_NONE) out_rows = N / group;
    DevBuf dX, dY, dWt, dO, dS, dRP, dCP;
    SyncOnExit sync_on_exit{sc.s, nullptr};
    auto up = [&](DevBuf& b, const float* h, size_t count) -> cudaError_t {
        cudaError_t e = b.alloc(count * 4, sc.s);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(b.p, h, count * 4, cudaMemcpyHostToDevice, sc.s);
    };
    JRK_CUDA_OK(up(dX, X, (size_t)N * d));
    JRK_CUDA_OK(up(dY, Y, (size_t)M * d));
    JRK_CUDA_OK(up(dWt, W, (size_t)M * m));
    if (dim_scale) JRK_CUDA_OK(up(dS, dim_scale, (size_t)d));
    if (row_pref) JRK_CUDA_OK(up(dRP, row_pref, (size_t)N));
    if (col_pref) JRK_CUDA_OK(up(dCP, col_pref, (size_t)M));
    JRK_CUDA_OK(dO.alloc((size_t)out_rows * m * 4, sc.s));
    int rc = jrk_kmatw_f32(dX.as<float>(), dY.as<float>(), dWt.as<float>(), dO.as<float>(), N, M, d, m, d, d, m, m,
                           scale, dim_scale ? dS.as<float>() : nullptr, power, nconst,
                           row_pref ? dRP.as<float>() : nullptr, col_pref ? dCP.as<float>() : nullptr, row_reduce,
                           group, nullptr, 0, sc.s);
    if (rc) return rc;
    JRK_CUDA_OK(cudaMemcpyAsync(out, dO.p, (size_t)out_rows * m * 4, cudaMemcpyDeviceToHost, sc.s));
    JRK_CUDA_OK(cudaStreamSynchronize(sc.s));
    return JRK_OK;
}

}  // extern "C"

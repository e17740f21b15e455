// Shared device/host helpers for the sm_100a kernels of the JaxRK hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/jrk.h"

namespace jrk {

// ---------------------------------------------------------------------------------
// host: error plumbing (C ABI never throws; jrk_last_error returns the text)
// ---------------------------------------------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

#define JRK_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::jrk::fail(JRK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

#define JRK_LAUNCHED()                                                                      \
    do {                                                                                    \
        ::jrk::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        JRK_CUDA_OK(cudaGetLastError());                                                    \
    } while (0)

// Stream-ordered scratch: a caller-provided workspace if it is big enough, otherwise
// cudaMallocAsync on the stream (freed with cudaFreeAsync by the destructor).
struct Scratch {
    cudaStream_t stream;
    char* base = nullptr;
    size_t cap = 0, used = 0;
    bool owned = false;
    Scratch(void* ws, size_t ws_bytes, cudaStream_t s) : stream(s), base((char*)ws), cap(ws ? ws_bytes : 0) {}
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e != cudaSuccess) return e;
        base = (char*)p; cap = bytes; used = 0; owned = true;
        return cudaSuccess;
    }
    template <typename T>
    T* take(size_t count) {
        size_t off = (used + 255) & ~size_t(255);
        used = off + count * sizeof(T);
        return (T*)(base + off);
    }
    ~Scratch() { if (owned) cudaFreeAsync(base, stream); }
};
inline size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// ---------------------------------------------------------------------------------
// device: MUFU approximations (ex2 / lg2 / sqrt), all 1 MUFU op each
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---------------------------------------------------------------------------------
// The kernel-family epilogue  sq -> k  (GenGaussKernel, jaxrk/kern/rbf.py:104-105 over
// jaxrk/kern/util.py:115 and jaxrk/utilities/eucldist.py:48):
//     k = nconst * exp(-(s * sqrt(max(sq, 0)))^p)
// evaluated as exp2 of a pre-scaled argument.  PM selects the specialisation:
//   PM_SQEXP   p == 2 : k = nconst * 2^(c2 * sq),            c2 = -s^2 log2(e)   1 MUFU
//   PM_LAPLACE p == 1 : k = nconst * 2^(c1 * sqrt(sq)),      c1 = -s   log2(e)   2 MUFU
//   PM_GENERAL        : k = nconst * 2^(-log2e * 2^(hp * lg2(s2 * sq))), hp = p/2 3 MUFU
//                       (sq == 0 -> lg2 = -inf -> 2^-inf = 0 -> k = nconst, as 0^(p/2) = 0)
// `nconst` is NOT applied here when the caller folds it elsewhere (K@W applies it once
// per output element).
// ---------------------------------------------------------------------------------
//   PM_EXTRA          : sibling kernels on the same distance front-end (jrk_*_kind_f32), sub-kind in EpiParams:
//       JRK_KIND_PERIODIC    k = exp(-2 (sin(pi s r) / ls)^2)            PeriodicKernel,    jaxrk/kern/rbf.py:152-154
//       JRK_KIND_THRESHSPIKE k = (s r)^p <= threshold ? spike : non_spike ThreshSpikeKernel, jaxrk/kern/rbf.py:203-206
enum { PM_SQEXP = 0, PM_LAPLACE = 1, PM_GENERAL = 2, PM_EXTRA = 3 };

struct EpiParams {
    float c;       // PM_SQEXP: c2;  PM_LAPLACE: c1;  PM_GENERAL: s^2
    float hp;      // PM_GENERAL: p / 2
    float nconst;  // normalising constant
    // distance-only output (jrk_dist_f32): d = (s * sqrt(sq))^p
    float s;
    float p;
    // PM_EXTRA
    int kind;      // JRK_KIND_*
    float a0;      // PERIODIC: 1 / length_scale;  THRESHSPIKE: (threshold^(1/p) / s)^2, the threshold on r^2
    float a1;      // THRESHSPIKE: spike
    float a2;      // THRESHSPIKE: non_spike
    // kmatw.cu only: when set, the FP32-pipe K@W kernel returns at once inside the factored regime of kmatw_tc2.cuh (the
    // tensor-core kernel enqueued next to it has written the result); two norm maxima, as bit patterns
    const unsigned* skip_if_fast = nullptr;
};

// The sibling-kernel entry points (capi.cu) reuse the GenGauss launchers: the selector travels as an explicit argument
// (default: GenGauss), make_epi / pick_pmode pick it up, and the tensor-core paths refuse it.
struct ExtraKind {
    int kind = 0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
};

// Process-wide switches of the library (include/jrk.h: jrk_set_option / JRK_OPT_*), read once per call from an atomic;
// the environment variable of the same name only provides the initial value at first use.
int jrk_option(int opt);

constexpr float kLog2e = 1.4426950408889634f;

// NONNEG: the caller guarantees sq >= 0 (a sum of squares from direct differencing), so the clip of
// eucldist.py:48 is a no-op and is skipped.
template <int PM, bool NONNEG = false>
__device__ __forceinline__ float kexp_unnorm(float sq, const EpiParams& e) {
    if (!NONNEG) sq = fmaxf(sq, 0.f);
    if (PM == PM_SQEXP) {
        return ex2_approx(sq * e.c);
    } else if (PM == PM_LAPLACE) {
        return ex2_approx(sqrt_approx(sq) * e.c);
    } else if (PM == PM_GENERAL) {
        float u = ex2_approx(e.hp * lg2_approx(sq * e.c));
        return ex2_approx(u * -kLog2e);
    } else {
        if (e.kind == JRK_KIND_PERIODIC) {
            // distance in periods; sin(pi t) has period 2 in t, so reduce t to [-1, 1] exactly before the MUFU sine
            // (sin.approx is only accurate on a few periods around 0).  3 MUFU per evaluation.
            const float t = e.s * sqrt_approx(sq);
            const float r = fmaf(-2.f, rintf(0.5f * t), t);
            float sn;
            asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(r * 3.14159265358979323846f));
            const float q = sn * e.a0;
            return ex2_approx(q * q * (-2.f * kLog2e));
        }
        // ThreshSpike: (s r)^p <= threshold  <=>  r^2 <= (threshold^(1/p) / s)^2 =: a0 (monotone; computed on the host
        // in double), so the epilogue is one compare + select and the Gram stays HBM-write-bound
        return sq <= e.a0 ? e.a1 : e.a2;
    }
}

inline int pick_pmode(float power, const ExtraKind& ex = ExtraKind()) {
    if (ex.kind != 0) return PM_EXTRA;
    if (power == 2.0f) return PM_SQEXP;
    if (power == 1.0f) return PM_LAPLACE;
    return PM_GENERAL;
}

inline EpiParams make_epi(float scale, float power, float nconst, const ExtraKind& ex = ExtraKind()) {
    EpiParams e;
    double s = (double)scale;
    int pm = pick_pmode(power, ex);
    if (pm == PM_EXTRA) pm = (power == 2.0f) ? PM_SQEXP : (power == 1.0f) ? PM_LAPLACE : PM_GENERAL;   // e.c unused
    if (pm == PM_SQEXP) e.c = (float)(-s * s * 1.4426950408889634);
    else if (pm == PM_LAPLACE) e.c = (float)(-s * 1.4426950408889634);
    else e.c = (float)(s * s);
    e.hp = 0.5f * power;
    e.nconst = nconst;
    e.s = scale;
    e.p = power;
    e.kind = ex.kind;
    e.a0 = ex.a0;
    e.a1 = ex.a1;
    e.a2 = ex.a2;
    return e;
}

// ---------------------------------------------------------------------------------
// device: packed FP32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100).  Measured on B200
// (profiles/microbench): the sub+fma differencing step reaches the full 128 lanes/clk/SM
// only in packed form (1.79e13 vs 1.25e13 dim-steps/s scalar).
// ---------------------------------------------------------------------------------
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---------------------------------------------------------------------------------
// device: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void st_global_cs_v4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace jrk

// Pipe-throughput microbenchmark for the roofline denominators that
// MEASURED_PEAKS.json does not carry ("measure with an FMA
// microbenchmark"): FP32 FFMA (scalar and packed f32x2), FADD+FFMA mix as used by
// direct differencing, MUFU ex2/lg2/sqrt, broadcast LDS.128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int NACC = 16;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed fma.rn.f32x2: two FMAs per lane per instruction
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[NACC];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float x = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x));
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// direct-differencing inner step, scalar: d = x - y; acc = fma(d, d, acc)
__global__ void k_diff(float* out, float y0, float y1) {
    float acc[NACC], x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = 0.f; x[i] = threadIdx.x * 1e-3f + i; }
    for (int it = 0; it < ITERS; ++it) {
        y0 += y1;  // loop-carried so the differences cannot be hoisted
#pragma unroll
        for (int i = 0; i < NACC; ++i) { float d = x[i] - y0; acc[i] = fmaf(d, d, acc[i]); }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same, packed: sub.f32x2 + fma.rn.f32x2
__global__ void k_diff2(float* out, float y0, float y1) {
    unsigned long long acc[NACC], x[NACC];
    unsigned long long yv0, yv1;
    asm("mov.b64 %0, {%1, %1};" : "=l"(yv0) : "f"(y0));
    asm("mov.b64 %0, {%1, %1};" : "=l"(yv1) : "f"(y1));
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float xv = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(xv));
        acc[i] = 0ull;
    }
    for (int it = 0; it < ITERS; ++it) {
        asm volatile("add.f32x2 %0, %0, %1;" : "+l"(yv0) : "l"(yv1));
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            unsigned long long d;
            asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x[i]), "l"(yv0));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc[i]) : "l"(d));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__global__ void k_mufu(float* out, float a) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.5f + threadIdx.x * 1e-4f + i * 1e-2f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            float r;
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i]));
            else if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i]));
            else if (OP == 2) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i]));
            else asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i]));
            acc[i] = r;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s * a;
}

// MUFU ex2 interleaved with 8 FFMA per MUFU: do they overlap?
__global__ void k_mix(float* out, float a, float b) {
    float acc[NACC], e[2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    e[0] = 0.25f; e[1] = 0.5f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < 2; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
    }
    float s = e[0] + e[1];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// broadcast LDS.128: every lane of the warp reads the same 16 bytes
__global__ void k_lds_bcast(float* out) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float4 v = sm[(it * 16 + i) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F>
float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk_khz);
    const int threads = 256, blocks = sms * 8;
    float* out; CK(cudaMalloc(&out, sizeof(float) * threads * blocks));
    double total_thr = double(threads) * blocks;
    double n;
    float ms;

    ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    n = total_thr * ITERS * NACC;
    printf(", \"ffma_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    n = total_thr * ITERS * NACC * 2;
    printf(", \"ffma2_fma_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_diff<<<blocks, threads>>>(out, 0.1f, 0.2f); });
    n = total_thr * ITERS * NACC;
    printf(", \"diff_scalar_dimsteps_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_diff2<<<blocks, threads>>>(out, 0.1f, 0.2f); });
    n = total_thr * ITERS * NACC * 2;
    printf(", \"diff_packed_dimsteps_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_mufu<0><<<blocks, threads>>>(out, 1.f); });
    n = total_thr * ITERS * NACC;
    printf(", \"mufu_ex2_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_mufu<1><<<blocks, threads>>>(out, 1.f); });
    printf(", \"mufu_lg2_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_mufu<2><<<blocks, threads>>>(out, 1.f); });
    printf(", \"mufu_sqrt_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_mufu<3><<<blocks, threads>>>(out, 1.f); });
    printf(", \"mufu_rsqrt_per_s\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    n = total_thr * ITERS * NACC;
    printf(", \"mix_ffma_per_s_with_1_mufu_per_8\": %.4e", n / (ms * 1e-3));
    ms = time_ms([&] { k_lds_bcast<<<blocks, threads>>>(out); });
    n = total_thr / 32 * ITERS * 16;
    printf(", \"lds128_bcast_warp_instr_per_s\": %.4e", n / (ms * 1e-3));
    printf("}\n");
    return 0;
}

// How much does an MMA-issuing warp cost the OTHER warps of its SM sub-partition?
// 16 "epilogue-like" warps (4 per sub-partition) run a MUFU + ALU loop shaped like the K@W epilogue and report their
// cycle counts; one extra warp on sub-partition 0 (warp 16) runs one of several service loops for the same duration:
//   mode 0: nothing            mode 1: tcgen05.mma N = 16 (K = 16, f16) back to back, commit every 8
//   mode 2: the same, N = 32   mode 3: elect + R2UR-style address arithmetic without the MMA
//   mode 4: mbarrier.try_wait spinning on a barrier that never completes (short time-outs)
//   mode 5: tcgen05.mma paced: 11 MMAs + 3 commits, then ~450 clk of try_wait on their completion (the kernel's rhythm)
//   mode 6: the same with operands that change every iteration (R2UR per MMA)   mode 7: mode 6 without the pause
// Output: cycles of the MUFU warps per sub-partition, relative to sub-partitions 1-3 (which host no service warp).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../jaxrk_b200/csrc -o issue_interference issue_interference.cu
#include <cstdio>
#include <vector>

#include "gram_tf32_kernel.cuh"

using namespace jrk;
using namespace jrk::tf32;

constexpr int ITERS = 2000;

__global__ void __launch_bounds__(17 * 32, 1) interf_kernel(int mode, int use_tmem, long long* cyc, float* sink, long long* svc) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); stop = 0; }
    if (warp == 16) tmem_alloc_512(smem_u32(&slot));
    for (int i = tid; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp < 16) {
        // epilogue-like loop: per iteration 16 x (MUFU, LOP, half FADD2, half FMUL2, 2 half F2FP)
        float v[16];
        for (int c = 0; c < 16; ++c) v[c] = 0.001f * (float)(lane + c);
        uint32_t acc = 0;
        const long long t0 = clock64();
        const uint32_t my = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 64;   // own TMEM block (use_tmem)
        for (int it = 0; it < ITERS; ++it) {
            if (use_tmem) {   // as in the kernel: the 16 exponents come from TMEM ...
                tmem_ld16(my + 16 * (it & 3), v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = __uint_as_float((__float_as_uint(v[c]) & 0x007fffffu) | 0x3a800000u);   // keep them tame
            }
            uint32_t eo[16];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                const float e0 = ex2_approx(v[c]), e1 = ex2_approx(v[c + 1]);
                const float h0 = __uint_as_float(__float_as_uint(e0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(e1) & 0xFFFFE000u);
                float l0, l1;
                unpack2(mul2(sub2(pack2(e0, e1), pack2(h0, h1)), pack2(2048.f, 2048.f)), l0, l1);
                uint32_t a, b;
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(h1), "f"(h0));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(l1), "f"(l0));
                acc ^= a + b;
                eo[c >> 1] = a;
                eo[8 + (c >> 1)] = b;
                v[c] = l0 * 1e-3f;
                v[c + 1] = l1 * 1e-3f;
            }
            if (use_tmem) tmem_st16(my + 16 * (it & 3), eo);   // ... and the packed pairs go back
        }
        if (use_tmem) tmem_st_wait();
        const long long t1 = clock64();
        if (lane == 0) cyc[warp] = t1 - t0;
        sink[tid] = __uint_as_float(acc);
        if (lane == 0) atomicAdd((int*)&stop, 1);
    } else {
        const uint32_t id16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t desc = make_b_desc(smem_u32(smem));
        long long n = 0;
        uint32_t ph = 0;
        const long long t0 = clock64();
        while (stop < 16) {
            if (mode == 1 || mode == 2) {
                if (elect_one()) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) mma_f16_ts(tb + 64, tb + 0, desc, mode == 1 ? id16 : id32, 1u);
                    tc_commit(&bar[0]);
                }
                __syncwarp();
                mbar_wait(&bar[0], ph);
                ph ^= 1;
                n += 8;
            } else if (mode == 3) {
                uint32_t x = (uint32_t)n;
                if (elect_one()) {
                    asm volatile("{.reg .b32 t; mov.b32 t, %0; }" ::"r"(x));
                }
                __syncwarp();
                ++n;
            } else if (mode == 4) {
                uint32_t ok;
                asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p;}"
                             : "=r"(ok) : "r"(smem_u32(&bar[1])), "r"(0u), "r"(200u) : "memory");
                ++n;
            } else if (mode == 5) {
                if (elect_one()) {
                    mma_f16_ts(tb + 128, tb + 0, desc, id64, 0u);
                    mma_f16_ts(tb + 128, tb + 8, desc, id64, 1u);
                    mma_f16_ts(tb + 128, tb + 0, desc, id64, 1u);
                    tc_commit(&bar[2]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        mma_f16_ts(tb + 64, tb + 16 * r, desc, id32, 1u);
                        mma_f16_ts(tb + 80, tb + 16 * r + 8, desc, id16, 1u);
                    }
                    tc_commit(&bar[0]);
                    tc_commit(&bar[3]);
                }
                __syncwarp();
                mbar_wait(&bar[0], ph);
                const long long w0 = clock64();
                while (clock64() - w0 < 300) {
                    uint32_t ok;
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p;}"
                                 : "=r"(ok) : "r"(smem_u32(&bar[1])), "r"(0u), "r"(20000u) : "memory");
                }
                ph ^= 1;
                n += 11;
            } else if (mode == 6 || mode == 7) {
                // the kernel's rhythm with operands that CHANGE every iteration (descriptor, accumulator and A columns live
                // in vector registers and go through R2UR for every MMA); mode 7: no waiting in between (issue-bound)
                const uint32_t st = (uint32_t)(n / 11) % 6;
                const uint64_t dq = desc + (uint64_t)(((uint32_t)(n / 11) % 8) * 32);
                const uint32_t cS = tb + 128 + st * 64, oc = tb + 32 + (st & 1) * 32;
                if (elect_one()) {
                    mma_f16_ts(cS, tb + 0, dq + 2, id64, 0u);
                    mma_f16_ts(cS, tb + 16, dq, id64, 1u);
                    mma_f16_ts(cS, tb + 0, dq, id64, 1u);
                    tc_commit(&bar[2]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        mma_f16_ts(oc, cS + 16 * r, dq + 4 + 2 * r, id32, 1u);
                        mma_f16_ts(oc + 16, cS + 16 * r + 8, dq + 4 + 2 * r, id16, 1u);
                    }
                    tc_commit(&bar[0]);
                    tc_commit(&bar[3]);
                }
                __syncwarp();
                mbar_wait(&bar[0], ph);
                if (mode == 6) {
                    const long long w0 = clock64();
                    while (clock64() - w0 < 300) {
                        uint32_t ok;
                        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p;}"
                                     : "=r"(ok) : "r"(smem_u32(&bar[1])), "r"(0u), "r"(20000u) : "memory");
                    }
                }
                ph ^= 1;
                n += 11;
            } else {
                __nanosleep(1000);
            }
        }
        if (lane == 0) { svc[0] = n; svc[1] = clock64() - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc_512(tb);
}

int main() {
    long long *cyc, *svc;
    float* sink;
    cudaMalloc(&cyc, 16 * 8);
    cudaMalloc(&svc, 16);
    cudaMalloc(&sink, 17 * 32 * 4);
    cudaFuncSetAttribute(interf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int use_tmem = 0; use_tmem <= 1; ++use_tmem)
    for (int mode = 0; mode <= 7; ++mode) {
        cudaMemset(svc, 0, 16);
        interf_kernel<<<1, 17 * 32, 16384>>>(mode, use_tmem, cyc, sink, svc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long h[16], s[2];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(s, svc, sizeof(s), cudaMemcpyDeviceToHost);
        double q[4] = {0, 0, 0, 0};
        for (int w = 0; w < 16; ++w) q[w & 3] += (double)h[w] / 4;
        const double base = (q[1] + q[2] + q[3]) / 3;
        // per sub-partition: 4 warps x ITERS x 16 MUFU x 8 clk = the MUFU-bound time
        printf("{\"tmem_ld_st\": %d, \"mode\": %d, \"smsp0_clk\": %.0f, \"smsp123_clk\": %.0f, \"ratio\": %.3f, \"mufu_bound_clk\": %d, \"service_ops\": %lld, \"service_clk\": %lld, "
               "\"extra_clk_per_service_op\": %.1f}\n",
               use_tmem, mode, q[0], base, q[0] / base, 4 * ITERS * 16 * 8, s[0], s[1], s[0] ? (q[0] - base) / (double)s[0] * 1.0 : 0.0);
    }
    return 0;
}

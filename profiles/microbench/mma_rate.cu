// tcgen05.mma kind::tf32 issue-rate microbenchmark: cycles per MMA for M = 128, K = 8, N in {8 .. 256}, A operand
// from TMEM (".ts" form used by the Gram / K@W kernels) or from shared memory, one CTA per SM, one issuing thread.
// Question it answers: what does GEMM2 of kmatw_tc.cu (24 MMAs of N = 16 per tile) really cost?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {   // K-major, 128-byte swizzle, 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MODE 0: A from TMEM (columns 0..7), 1: A from shared memory.  NDST: number of distinct accumulators cycled through.
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_mma(long long* clk, int N, int iters, int ndst) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f800000u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (warp == 0) {   // warp-uniform control flow, the MMAs predicated to one elected lane (no waterfall loop)
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t db = make_desc(smem_u32(smem)), da = make_desc(smem_u32(smem) + 32768);
        const long long t0 = clock64();
#pragma unroll 8
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tm + 256 + (((uint32_t)(it & (ndst - 1)) * (uint32_t)N) & 255u);
            if (!elected) continue;
            if (MODE == 0) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tm), "l"(db), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        }
        if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    long long* clk;
    CK(cudaMalloc(&clk, sms * 8));
    const int smem = 64 * 1024 + 1024, iters = 4096;
    CK(cudaFuncSetAttribute(k_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 2; ++mode)
        for (int ndst : {1, 2})
            for (int N : {8, 16, 32, 64, 128, 256}) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) k_mma<0><<<sms, 128, smem>>>(clk, N, iters, ndst);
                    else k_mma<1><<<sms, 128, smem>>>(clk, N, iters, ndst);
                    CK(cudaDeviceSynchronize());
                }
                long long h[256];
                CK(cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost));
                double avg = 0;
                for (int i = 0; i < sms; ++i) avg += (double)h[i];
                avg /= sms;
                printf("{\"bench\": \"mma_tf32\", \"a_from\": \"%s\", \"M\": 128, \"N\": %d, \"K\": 8, \"accumulators\": %d, \"clk_per_mma\": %.2f, \"tflops_chip\": %.1f}\n",
                       mode == 0 ? "tmem" : "smem", N, ndst, avg / iters, 2.0 * 128 * N * 8 * iters / avg * 1.965e9 * sms / 1e12);
            }
    // latency of issue -> commit -> mbarrier observed by the issuing warp, for a few MMAs (N = 64, A from TMEM)
    for (int it : {1, 2, 4, 8}) {
        for (int rep = 0; rep < 2; ++rep) { k_mma<0><<<sms, 128, smem>>>(clk, 64, it, 1); CK(cudaDeviceSynchronize()); }
        long long h[256];
        CK(cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        printf("{\"bench\": \"mma_commit_latency\", \"N\": 64, \"mmas\": %d, \"clk_issue_to_barrier\": %.0f}\n", it, avg / sms);
    }
    return 0;
}

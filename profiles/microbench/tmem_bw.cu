// TMEM <-> register bandwidth microbenchmark (tcgen05.ld / tcgen05.st, shape 32x32b), one CTA per SM.
// Question it answers: how many bytes per clock and SM can the epilogue warps of the tcgen05 Gram / K@W kernels
// pull out of the accumulators?  (The skeleton of gram_tf32_kernel -- no MMA, no stores -- took 1940 clk per
// 128 x 128 tile with two accumulators = 128 KB of tcgen05.ld per tile: 66 B/clk.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NC>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
    if constexpr (NC == 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        sink ^= r[0] ^ r[7];
    } else if constexpr (NC == 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
        sink ^= r[0] ^ r[15];
    } else {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
        sink ^= r[0] ^ r[31];
    }
}
__device__ __forceinline__ void st16(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}

// MODE 0: loads only; 1: stores only; 2: one load + one store of the same size (x16) per iteration.
// BATCH loads are issued back to back before one tcgen05.wait::ld.
template <int NC, int MODE, int BATCH>
__global__ void __launch_bounds__(1024, 1) k_tmem(uint32_t* out, long long* clk) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    // every warp initialises the columns it will read (its lane quarter); column windows of warps sharing a quarter differ
    const uint32_t col0 = (uint32_t)((warp >> 2) * 64) & 511u;
    for (int c = 0; c < 64; c += 16) st16(base + ((col0 + c) & 511u), threadIdx.x);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int b = 0; b < BATCH; ++b) ld<NC>(base + ((col0 + (b * NC)) & 63u) + (col0 & ~63u), sink);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int b = 0; b < BATCH * NC / 16; ++b) st16(base + ((col0 + b * 16) & 63u) + (col0 & ~63u), sink);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int NC, int MODE, int BATCH>
void run(const char* name, int warps, uint32_t* out, long long* clk, int sms) {
    k_tmem<NC, MODE, BATCH><<<sms, warps * 32>>>(out, clk);
    CK(cudaDeviceSynchronize());
    k_tmem<NC, MODE, BATCH><<<sms, warps * 32>>>(out, clk);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double bytes = (double)ITERS * BATCH * NC * 4.0 * 32.0 * warps * (MODE == 2 ? 2.0 : 1.0);
    printf("{\"bench\": \"%s\", \"x\": %d, \"batch\": %d, \"warps\": %d, \"clk\": %.0f, \"bytes_per_clk_per_sm\": %.1f}\n", name, NC, BATCH, warps,
           avg, bytes / avg);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    uint32_t* out;
    long long* clk;
    CK(cudaMalloc(&out, sms * 1024 * 4));
    CK(cudaMalloc(&clk, sms * 8));
    for (int warps : {4, 8, 16, 32}) {
        run<8, 0, 2>("tmem_ld", warps, out, clk, sms);
        run<16, 0, 1>("tmem_ld", warps, out, clk, sms);
        run<16, 0, 2>("tmem_ld", warps, out, clk, sms);
        run<32, 0, 1>("tmem_ld", warps, out, clk, sms);
        run<32, 0, 2>("tmem_ld", warps, out, clk, sms);
        run<16, 1, 2>("tmem_st", warps, out, clk, sms);
        run<16, 2, 2>("tmem_ld+st", warps, out, clk, sms);
    }
    return 0;
}

// Probe of the tcgen05.mma assumptions the second-generation K@W kernel (kmatw_tc2.cu) relies on, each checked against
// a host computation:
//   - kind::f16 with A in TMEM and B in shared memory (K-major, 128B swizzle) at sub-row byte offsets (32 / 64 / 96)
//     and at row offsets that are multiples of 8 rows (several small operands sharing one swizzled image)
//   - N = 16 and N = 32 (two 16-row operands stacked)
//   - MIXED operand formats inside kind::f16: A fp16 x B bf16 and A bf16 x B fp16
//   - issue cost of those MMAs (clock64 around a run of 64 MMAs + commit)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../jaxrk_b200/csrc -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "gram_tf32_kernel.cuh"

using namespace jrk;
using namespace jrk::tf32;

struct Case {
    int afmt, bfmt;   // 0 fp16, 1 bf16
    int n;            // MMA N
    int row0;         // first image row of the B operand (multiple of 8)
    int boff;         // logical byte offset of the operand inside the 128-byte rows
    int reps;         // timing: number of MMAs issued back to back (accumulating)
};

__device__ __host__ inline float aval(int i, int k) { return (float)(((i * 7 + k * 3) % 17) - 8) * 0.125f; }
__device__ __host__ inline float bval(int n, int k) { return (float)(((n * 5 + k * 11) % 13) - 6) * 0.25f; }

__device__ __forceinline__ uint32_t pack16(float lo, float hi, int fmt) {
    if (fmt == 0) return (uint32_t)__half_as_ushort(__float2half_rn(lo)) | ((uint32_t)__half_as_ushort(__float2half_rn(hi)) << 16);
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Case cs, float* out, long long* clk) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc_512(smem_u32(&slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    // image: 64 rows x 128 B, zero, then the operand at (row0 + n, boff + 2k)
    for (int i = tid; i < 64 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x7e007e00u;   // NaN poison (fp16)
    __syncthreads();
    for (int idx = tid; idx < cs.n * 8; idx += 128) {
        const int n = idx >> 3, kp = idx & 7;       // pair of k
        const int r = cs.row0 + n, b = cs.boff + kp * 4;
        const int phys = r * 128 + ((((b >> 4) ^ (r & 7)) << 4) | (b & 15));
        *reinterpret_cast<uint32_t*>(smem + phys) = pack16(bval(n, 2 * kp), bval(n, 2 * kp + 1), cs.bfmt);
    }
    // A: lane = tid, 16 k-values -> 8 packed columns at TMEM columns 0..7
    {
        uint32_t a[8];
        for (int kp = 0; kp < 8; ++kp) a[kp] = pack16(aval(tid, 2 * kp), aval(tid, 2 * kp + 1), cs.afmt);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tb + ((uint32_t)(warp * 32) << 16)),
                     "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
        tmem_st_wait();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)cs.afmt << 7) | ((uint32_t)cs.bfmt << 10) | ((uint32_t)(cs.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t desc = make_b_desc(smem_u32(smem)) + (uint64_t)((cs.row0 * 128 + cs.boff) >> 4);
        t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < cs.reps; ++r) mma_f16_ts(tb + 64, tb + 0, desc, idesc, r > 0 ? 1u : 0u);
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (tid == 0) clk[0] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    // D: 128 lanes x n columns at TMEM column 64
    for (int c0 = 0; c0 < cs.n && c0 < 32; c0 += 16) {
        float v[16];
        tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
        tmem_ld_wait();
        for (int c = 0; c < 16; ++c) out[tid * 32 + c0 + c] = v[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(tb);
}

static float rnd16(float v, int fmt) {
    if (fmt == 0) return __half2float(__float2half_rn(v));
    return __bfloat162float(__float2bfloat16_rn(v));
}

int main() {
    const Case cases[] = {
        {0, 0, 64, 0, 0, 1},   {0, 0, 64, 0, 32, 1},  {0, 0, 16, 32, 64, 1}, {0, 0, 16, 48, 96, 1}, {0, 0, 32, 32, 96, 1},
        {0, 0, 16, 16, 64, 1}, {1, 1, 16, 0, 0, 1},
        // timing (results are multiples of the single product)
        {0, 0, 16, 32, 64, 64}, {0, 0, 32, 32, 64, 64}, {0, 0, 64, 0, 0, 64}, {0, 0, 16, 32, 64, 256}, {0, 0, 32, 32, 64, 256},
        {0, 0, 64, 0, 0, 256},
        // MIXED operand formats (A fp16 x B bf16, A bf16 x B fp16): "an illegal instruction was encountered" on sm_100a --
        // kept last, the probe stops there
        {0, 1, 16, 32, 64, 1}, {1, 0, 16, 32, 64, 1},
    };
    float* out;
    long long* clk;
    cudaMalloc(&out, 128 * 32 * 4 * 4);
    cudaMalloc(&clk, 8);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    int bad_total = 0;
    for (const Case& cs : cases) {
        cudaMemset(out, 0, 128 * 32 * 4);
        probe_kernel<<<1, 128, 16384>>>(cs, out, clk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("case a%d b%d n%d row0 %d boff %d: CUDA error %s\n", cs.afmt, cs.bfmt, cs.n, cs.row0, cs.boff, cudaGetErrorString(e)); return 1; }
        std::vector<float> h(128 * 32);
        long long c;
        cudaMemcpy(h.data(), out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        const int nn = cs.n < 32 ? cs.n : 32;
        for (int i = 0; i < 128; ++i)
            for (int n = 0; n < nn; ++n) {
                double ref = 0;
                for (int k = 0; k < 16; ++k) ref += (double)rnd16(aval(i, k), cs.afmt) * (double)rnd16(bval(n, k), cs.bfmt);
                ref *= cs.reps;
                const double err = std::fabs((double)h[i * 32 + n] - ref);
                if (!(err <= maxerr)) maxerr = err;
            }
        const bool ok = maxerr < 1e-3;
        bad_total += !ok;
        printf("{\"afmt\": %d, \"bfmt\": %d, \"n\": %d, \"row0\": %d, \"boff\": %d, \"reps\": %d, \"max_err\": %.3g, \"ok\": %s, \"clk\": %lld, \"clk_per_mma\": %.1f}\n",
               cs.afmt, cs.bfmt, cs.n, cs.row0, cs.boff, cs.reps, maxerr, ok ? "true" : "false", c, (double)c / cs.reps);
    }
    printf("probe %s\n", bad_total ? "FAILED" : "ok");
    return 0;
}

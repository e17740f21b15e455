// How fast can a B200 absorb the Gram kernels' output pattern?  Every CTA writes tiles of `tw` floats x `th` rows of a
// row-major N x N float matrix (a warp store = 32 consecutive floats of one row), tiles taken in run-major order by a
// persistent grid -- the store pattern of gram_tf32_kernel (tw = 128, th = 64) -- against a linear fill.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_pattern write_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int VEC>
__global__ void tile_writer(float* K, long long n, int tw, int th, int tiles_per_run, long long n_units, int n_jb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int wpr = tw / (32 * VEC);                 // warps across one tile row
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        const long long jb = u % n_jb, run = u / n_jb;
        for (int t = 0; t < tiles_per_run; ++t) {
            const long long i0 = (run * tiles_per_run + t) * th, j0 = jb * tw;
            if (i0 >= n) break;
            for (int w = warp; w < th * wpr; w += nwarps) {
                const int r = w / wpr, c = (w % wpr) * 32 * VEC + lane * VEC;
                float* p = K + (i0 + r) * n + j0 + c;
                if (VEC == 1) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
                else asm volatile("st.global.cs.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
            }
        }
    }
}
__global__ void linear_fill(float4* K, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        K[i] = make_float4(1.f, 1.f, 1.f, 1.f);
}
template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f(); CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int i = 0; i < 5; ++i) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best; }
    return best;
}
int main() {
    const long long n = 65536;
    float* K; CK(cudaMalloc(&K, n * n * 4));
    const double gb = n * n * 4 / 1e9;
    float ms = time_ms([&] { linear_fill<<<148 * 8, 512>>>((float4*)K, n * n / 4); });
    printf("{\"pattern\": \"linear float4 fill\", \"ms\": %.3f, \"GBps\": %.0f}\n", ms, gb / ms * 1e3);
    const int tws[] = {128, 256, 512, 1024, 4096};
    for (int vec : {1, 4}) for (int tw : tws) for (int th : {64, 16}) {
        const int n_jb = (int)(n / tw);
        const long long n_it = n / th;
        const int tiles_per_run = (int)((n_it + 4) / 5);
        const long long n_units = (long long)n_jb * ((n_it + tiles_per_run - 1) / tiles_per_run);
        if (vec == 1) ms = time_ms([&] { tile_writer<1><<<148, 512>>>(K, n, tw, th, tiles_per_run, n_units, n_jb); });
        else ms = time_ms([&] { tile_writer<4><<<148, 512>>>(K, n, tw, th, tiles_per_run, n_units, n_jb); });
        printf("{\"pattern\": \"tiles\", \"vec\": %d, \"tile_cols\": %d, \"tile_rows\": %d, \"ms\": %.3f, \"GBps\": %.0f}\n", vec, tw, th, ms, gb / ms * 1e3);
    }
    return 0;
}

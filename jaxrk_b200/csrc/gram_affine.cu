// Kernel (h): one pass over a materialised Gram that applies what the reduce chain leaves for it,
//     out[i, j] = K[i, j] * row_mul[i] * col_mul[j] + row_add[i] + col_add[j] + add_const        (out may be K: in place)
// i.e. Prefactors on both axes (jaxrk/reduce/base.py:84-101) and Center on either or both axes (base.py:183-192:
// C_N G C_M = G - 1 colmean^T - rowmean 1^T + mean) in ONE read + write of the N x M array.  The reference applies
// them as separate N x M passes (prefactor multiply per axis, mean + subtract per axis); the means themselves never
// read the Gram here -- they are skinny fused K@W calls (rkhs/_lowering.py).  HBM-bound: 8 bytes per entry.
#include "common.cuh"

namespace jrk {
namespace {

template <bool VEC>
__global__ void __launch_bounds__(256) gram_affine_kernel(const float* __restrict__ K, float* __restrict__ out, int64_t N,
                                                          int64_t M, int64_t ldk, int64_t ldo,
                                                          const float* __restrict__ row_mul, const float* __restrict__ col_mul,
                                                          const float* __restrict__ row_add, const float* __restrict__ col_add,
                                                          float add_const) {
    // a CTA walks rows blockIdx.y, blockIdx.y + gridDim.y, ...; its 256 threads own 1024 consecutive columns (VEC) or 256
    const int64_t j = ((int64_t)blockIdx.x * 256 + threadIdx.x) * (VEC ? 4 : 1);
    if (j >= M) return;
    float cm[4] = {1.f, 1.f, 1.f, 1.f}, ca[4] = {add_const, add_const, add_const, add_const};
#pragma unroll
    for (int q = 0; q < (VEC ? 4 : 1); ++q) {
        if (j + q < M) {
            if (col_mul) cm[q] = __ldg(col_mul + j + q);
            if (col_add) ca[q] += __ldg(col_add + j + q);
        }
    }
    for (int64_t i = blockIdx.y; i < N; i += gridDim.y) {
        const float rm = row_mul ? __ldg(row_mul + i) : 1.f, ra = row_add ? __ldg(row_add + i) : 0.f;
        if (VEC) {
            const float4 k = *reinterpret_cast<const float4*>(K + i * ldk + j);
            float4 o;
            o.x = fmaf(k.x * rm, cm[0], ra + ca[0]);
            o.y = fmaf(k.y * rm, cm[1], ra + ca[1]);
            o.z = fmaf(k.z * rm, cm[2], ra + ca[2]);
            o.w = fmaf(k.w * rm, cm[3], ra + ca[3]);
            *reinterpret_cast<float4*>(out + i * ldo + j) = o;
        } else {
            out[i * ldo + j] = fmaf(K[i * ldk + j] * rm, cm[0], ra + ca[0]);
        }
    }
}

}  // namespace

int gram_affine(const float* K, float* out, int64_t N, int64_t M, int64_t ldk, int64_t ldo, const float* row_mul,
                const float* col_mul, const float* row_add, const float* col_add, float add_const, cudaStream_t st) {
    const bool vec = (M % 4 == 0) && (ldk % 4 == 0) && (ldo % 4 == 0) && ((uintptr_t)K % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const int64_t per = vec ? 1024 : 256;
    const int64_t gx = (M + per - 1) / per;
    int64_t gy = (148 * 16 + gx - 1) / gx;                  // ~16 CTAs per SM in flight
    if (gy > N) gy = N;
    if (gy > 65535) gy = 65535;
    if (gx > 2147483647LL) return fail(JRK_ERR_UNSUPPORTED, "gram_affine: too many columns");
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (vec) gram_affine_kernel<true><<<grid, 256, 0, st>>>(K, out, N, M, ldk, ldo, row_mul, col_mul, row_add, col_add, add_const);
    else gram_affine_kernel<false><<<grid, 256, 0, st>>>(K, out, N, M, ldk, ldo, row_mul, col_mul, row_add, col_add, add_const);
    JRK_LAUNCHED();
    return JRK_OK;
}

}  // namespace jrk

// Shared between kmatw_tc.cu (dispatch, generic regime) and kmatw_tc2.cu (factored p = 2 regime): the regime decision,
// the scale conventions of the tile images and the launch entry points.
#pragma once
#include <cmath>

#include "common.cuh"

namespace jrk {

// p = 2 "factored" regime: |c| (max|x|^2 + max|y|^2) <= TC_FAST_BOUND, decided on the device from the pre-pass maxima
// (no host synchronisation).  There K_ij = 2^(c|x_i|^2) . 2^(-2c x_i.y_j) . 2^(c|y_j|^2) with every factor inside
// [2^-12, 2^12].  (What limits the bound: e = 2^S is handed to GEMM2 as fp16 hi / lo' with lo' = (e - hi) 2^11 < 2 ulp11(e)
// 2^10, which reaches fp16's maximum for e >= 2^15 -- so S must stay below 15 -- and the exponent error of the 3-term
// fp16 GEMM1 grows with the bound.  Measured up to 15.4 with adversarial aligned / opposed large-norm pairs
// (tools/regime_edge_check.py): 0.18 - 0.24 of the tolerance, the generic epilogue 0.24 - 0.34; 12 keeps a factor of 8 below
// the overflow.)
constexpr float TC_FAST_BOUND = 12.0f;

__device__ __forceinline__ bool tc2_fast_regime(const unsigned* nmax) {
    return (__uint_as_float(__ldg(nmax)) + __uint_as_float(__ldg(nmax + 1))) <= TC_FAST_BOUND;
}

// exponent ew with max|W'| 2^-ew in (2^13, 2^14]: W' is stored as fp16 hi / lo' pairs, normalised by 2^-ew; the result is
// multiplied by 2^ew once per output row.  Clamped so that 2^ew and 2^-ew are finite, normal floats.
__device__ __forceinline__ int tc2_w_exp(float wmax) {
    if (!(wmax > 0.f)) return 0;
    int e;
    frexpf(wmax, &e);              // wmax = f 2^e, f in [0.5, 1)
    e -= 14;
    return e < -100 ? -100 : (e > 100 ? 100 : e);
}

// GEMM1 operand scales: x~ = x . x_scale, y'' = y . y_scale with x_scale . y_scale = -2c exactly (x_scale a power of two
// next to sqrt(-2c)): both operands stay below sqrt(24) sqrt(2) < 7 in the factored regime, and the image is
// independent of X.
struct Tc2Scales {
    float x_scale, y_scale;
};
inline Tc2Scales tc2_scales(float c) {
    const float a = -2.f * c;                                   // > 0
    const int r = (int)std::lrint(-0.5 * std::log2((double)a)); // 2^-r ~ sqrt(a)
    Tc2Scales s;
    s.x_scale = std::ldexp(1.f, -r);
    s.y_scale = std::ldexp(a, r);
    return s;
}

int tc2_image_bytes(int d);

// nmax[2] = max |W col_pref|
int kmatw_tc2_wmax(const float* W, int64_t M, int m, int64_t ldw, const float* col_pref, unsigned* nmax, cudaStream_t st);

// W maximum (nmax[2]) and the tile images of the factored regime.  nmax[0..1] (norm maxima, tc_rownorm_kernel) and
// yn_all (c |y_j|^2) must already be enqueued on `st`.  The kernels return at once outside the factored regime.
int kmatw_tc2_pack(const float* Y, const float* W, int64_t M, int d, int m, int64_t ldy, int64_t ldw,
                   const float* col_pref, float c, const float* yn_all, unsigned* nmax, unsigned char* img,
                   cudaStream_t st);

// out[i, :m] = nconst row_pref_i sum_j k(x_i, y_j) W'[j, :]  for the factored regime (returns at once outside it)
int kmatw_tc2_apply(const float* X, int64_t N, int d, int64_t ldx, const float* xn, const unsigned char* img, int64_t M,
                    const unsigned* nmax, float c, float nconst, const float* row_pref, float* out, int64_t ldo, int m,
                    cudaStream_t st);

}  // namespace jrk

// Tile core shared by the direct-differencing Gram kernels (gram_direct.cu,
// gram_groupsum.cu): 128 x 128 output tile, 256 threads, 8 x 8 micro-tile per thread,
// X / Y tiles staged k-major in shared memory, packed FADD2/FFMA2 over row pairs.
#pragma once
#include "common.cuh"

namespace jrk {

constexpr int GD_BM = 128, GD_BN = 128, GD_THREADS = 256;
constexpr int GD_LD = 132;  // smem row stride (floats): multiple of 4 for LDS.128, not of 32

template <int KC>
__device__ __forceinline__ void gd_fill(float (*s)[GD_LD], const float* __restrict__ A, int64_t rows,
                                        int64_t r0, int d, int k0, int64_t lda, const float* __restrict__ dim_scale,
                                        bool vec_ok) {
    // tile = 128 rows x KC dims; stored transposed s[k][row]
    constexpr int QPR = KC / 4;              // float4 per row
    constexpr int TOTAL = GD_BM * QPR;       // float4 in the tile
    for (int idx = threadIdx.x; idx < TOTAL; idx += GD_THREADS) {
        // narrow tiles (d <= 8): consecutive threads take consecutive ROWS of one 4-wide k group, so the
        // transposed shared-memory stores below are conflict-free (measured +6 % at d = 8); wider tiles keep
        // the row-major assignment, whose global loads are coalesced (measured better for d >= 16)
        int row, kq;
        if (KC <= 8) { row = idx % GD_BM; kq = idx / GD_BM; }
        else { row = idx / QPR; kq = idx % QPR; }
        int64_t gr = r0 + row;
        int k = k0 + kq * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gr < rows) {
            const float* src = A + gr * lda + k;
            if (vec_ok && k + 3 < d) {
                float4 t = __ldg(reinterpret_cast<const float4*>(src));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (k + c < d) v[c] = __ldg(src + c);
            }
            if (dim_scale) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (k + c < d) v[c] *= __ldg(dim_scale + k + c);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) s[kq * 4 + c][row] = v[c];
    }
}


// acc[p][c] += (x_row - y_col)^2 over the KC staged dimensions.
// p = row pair (rows 2p, 2p+1 of this thread's 8 rows), c = column 0..7.
template <int KC>
__device__ __forceinline__ void gd_accumulate(const float (*sX)[GD_LD], const float (*sY)[GD_LD], int tx, int ty,
                                              f32x2 (&acc)[4][8]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const float4 xa = *reinterpret_cast<const float4*>(&sX[k][ty * 4]);
        const float4 xb = *reinterpret_cast<const float4*>(&sX[k][64 + ty * 4]);
        const float4 ya = *reinterpret_cast<const float4*>(&sY[k][tx * 4]);
        const float4 yb = *reinterpret_cast<const float4*>(&sY[k][64 + tx * 4]);
        const f32x2 xp[4] = {pack2(xa.x, xa.y), pack2(xa.z, xa.w), pack2(xb.x, xb.y), pack2(xb.z, xb.w)};
        const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const f32x2 yb2 = pack2(yv[c], yv[c]);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const f32x2 df = sub2(xp[p], yb2);
                acc[p][c] = fma2(df, df, acc[p][c]);
            }
        }
    }
}

// local row index (0..127) of accumulator (p, h) and local column of c
__device__ __forceinline__ int gd_row(int ty, int p, int h) { return (p >> 1) * 64 + ty * 4 + (p & 1) * 2 + h; }
__device__ __forceinline__ int gd_col(int tx, int c) { return (c >> 2) * 64 + tx * 4 + (c & 3); }

inline int gd_pick_kc(int d) { return (d <= 4) ? 4 : (d <= 8) ? 8 : (d <= 12) ? 4 : 16; }

}  // namespace jrk

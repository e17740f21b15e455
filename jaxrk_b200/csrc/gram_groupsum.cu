// Kernel (c'): group-reduced Gram, never materialised.
//
//   out[a, b] = sum_{i in group a} sum_{j in group b} row_pref[i] k(x_i, y_j) col_pref[j]
//
// Replaces FiniteVec.inner when BOTH vectors carry a BalancedRed (optionally after
// Prefactors): jaxrk/rkhs/vector.py:72-74 materialises the N x M Gram and folds it with
// jaxrk/reduce/base.py:152-181 on axis 0 and then 1 (BASELINE config 4: 4,194,304^2
// evaluations reduced to 1024 x 1024 -- a 70 TB Gram in the reference).
//
// A CTA owns one 128-row tile of X and one column GROUP: it walks the group's column
// tiles with the direct-differencing tile core (gram_tile.cuh), keeps the weighted sums
// of its 8 x 8 micro-tile in registers, and reduces once at the end (shuffles + shared
// memory) to one partial per CTA.  A finalize kernel adds the gx/128 row-tile partials
// of each cell in a fixed order (deterministic).  For a self inner product (X == Y, same groups and prefactors)
// only the cells on and above the diagonal are computed and the finalize kernel mirrors them: half the work, and
// the result is exactly symmetric.
#include "gram_tile.cuh"

namespace jrk {

template <int PM, int KC>
__global__ void __launch_bounds__(GD_THREADS, 2)
gram_groupsum_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ part,
                     int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                     const float* __restrict__ dim_scale, const float* __restrict__ row_pref,
                     const float* __restrict__ col_pref, int64_t gy, EpiParams ep, int vec_in, int sym_tiles_per_group) {
    __shared__ __align__(16) float sX[KC][GD_LD];
    __shared__ __align__(16) float sY[KC][GD_LD];
    __shared__ float red[GD_THREADS / 32];

    // symmetric problem: cells below the diagonal are mirrored by the finalize kernel
    if (sym_tiles_per_group && (int)blockIdx.x < (int)blockIdx.y / sym_tiles_per_group) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * GD_BM;          // row tile
    const int64_t jbeg = (int64_t)blockIdx.x * gy, jend = jbeg + gy;  // column group

    float rp[8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t gi = i0 + gd_row(ty, p, h);
            rp[p * 2 + h] = (gi < N) ? (row_pref ? __ldg(row_pref + gi) : 1.f) : 0.f;
        }

    float total = 0.f;
    for (int64_t j0 = jbeg; j0 < jend; j0 += GD_BN) {
        f32x2 acc[4][8];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[p][c] = 0ull;
        for (int k0 = 0; k0 < d; k0 += KC) {
            __syncthreads();
            gd_fill<KC>(sX, X, N, i0, d, k0, ldx, dim_scale, vec_in & 1);
            gd_fill<KC>(sY, Y, jend, j0, d, k0, ldy, dim_scale, vec_in & 2);
            __syncthreads();
            gd_accumulate<KC>(sX, sY, tx, ty, acc);
        }
        float cp[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t gj = j0 + gd_col(tx, c);
            cp[c] = (gj < jend) ? (col_pref ? __ldg(col_pref + gj) : 1.f) : 0.f;
        }
        float tile = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float r0 = 0.f, r1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float lo, hi;
                unpack2(acc[p][c], lo, hi);
                r0 = fmaf(kexp_unnorm<PM, true>(lo, ep), cp[c], r0);
                r1 = fmaf(kexp_unnorm<PM, true>(hi, ep), cp[c], r1);
            }
            tile = fmaf(rp[p * 2], r0, tile);
            tile = fmaf(rp[p * 2 + 1], r1, tile);
        }
        total += tile;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < GD_THREADS / 32; ++w) s += red[w];
        part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

// out[a, b] = scale * sum_{t < tiles_per_group} part[(a * tiles_per_group + t) * ncolg + b]
__global__ void groupsum_finalize_kernel(const float* __restrict__ part, int64_t nrowg, int64_t ncolg,
                                         int64_t tiles_per_group, float scale, float* __restrict__ out,
                                         int64_t ldo, int sym) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrowg * ncolg) return;
    const int64_t a = idx / ncolg, b = idx % ncolg;
    const int64_t pa = (sym && b < a) ? b : a, pb = (sym && b < a) ? a : b;   // mirrored cell
    float s = 0.f;
    for (int64_t t = 0; t < tiles_per_group; ++t) s += part[(pa * tiles_per_group + t) * ncolg + pb];
    out[a * ldo + b] = scale * s;
}

int gram_groupsum(const float* X, const float* Y, float* out, int64_t N, int64_t M, int d, int64_t ldx,
                  int64_t ldy, int64_t ldo, float scale, const float* dim_scale, float power, float nconst,
                  const float* row_pref, const float* col_pref, int64_t gx, int64_t gy, int average,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (gx % GD_BM != 0)
        return fail(JRK_ERR_UNSUPPORTED, "gram_groupsum: row group size %lld is not a multiple of %d", (long long)gx, GD_BM);
    const int64_t nrow_tiles = N / GD_BM, ncolg = M / gy, nrowg = N / gx;
    if (ncolg > 2147483647LL || nrow_tiles > 65535)
        return fail(JRK_ERR_UNSUPPORTED, "gram_groupsum: grid too large (%lld x %lld)", (long long)ncolg, (long long)nrow_tiles);
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(align256((size_t)nrow_tiles * ncolg * 4) + 256));
    float* part = scr.take<float>((size_t)nrow_tiles * ncolg);

    const EpiParams ep = make_epi(scale, power, nconst);
    const int pm = pick_pmode(power);
    if (pm == PM_EXTRA) return fail(JRK_ERR_UNSUPPORTED, "gram_groupsum: sibling kernels are not supported");
    int vec_in = 0;
    if (((uintptr_t)X % 16 == 0) && (ldx % 4 == 0)) vec_in |= 1;
    if (((uintptr_t)Y % 16 == 0) && (ldy % 4 == 0)) vec_in |= 2;
    // self inner product with identical groups / prefactors on both sides: compute the upper triangle only
    const bool sym = (X == Y && ldx == ldy && N == M && gx == gy && row_pref == col_pref);
    const int sym_tpg = sym ? (int)(gx / GD_BM) : 0;
    dim3 grid((unsigned)ncolg, (unsigned)nrow_tiles);
    const int kc = gd_pick_kc(d);
#define GS_GO(PM_, KC_)                                                                                      \
    gram_groupsum_kernel<PM_, KC_><<<grid, GD_THREADS, 0, st>>>(X, Y, part, N, M, d, ldx, ldy, dim_scale,    \
                                                                row_pref, col_pref, gy, ep, vec_in, sym_tpg)
#define GS_KC(PM_)                          \
    do {                                    \
        if (kc == 4) GS_GO(PM_, 4);         \
        else if (kc == 8) GS_GO(PM_, 8);    \
        else GS_GO(PM_, 16);                \
    } while (0)
    if (pm == PM_SQEXP) GS_KC(PM_SQEXP);
    else if (pm == PM_LAPLACE) GS_KC(PM_LAPLACE);
    else GS_KC(PM_GENERAL);
#undef GS_KC
#undef GS_GO
    JRK_LAUNCHED();
    float fscale = nconst;
    if (average) fscale = (float)((double)nconst / ((double)gx * (double)gy));
    const int64_t tot = nrowg * ncolg;
    groupsum_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(part, nrowg, ncolg, gx / GD_BM, fscale,
                                                                            out, ldo, sym ? 1 : 0);
    JRK_LAUNCHED();
    return JRK_OK;
}

}  // namespace jrk

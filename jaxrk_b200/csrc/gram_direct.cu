// Kernel (a1): tiled Gram matrix by FP32 direct differencing, for small d.
//
// Replaces the materialising chain of the reference
//   sqeucldist_simple (jaxrk/utilities/eucldist.py:10-18) -> eucldist (:37-48)
//   -> ScaledPairwiseDistance.__call__ (jaxrk/kern/util.py:107-116)
//   -> GenGaussKernel.__call__ (jaxrk/kern/rbf.py:104-105)
// (one GEMM + ~9 eager N x M passes) by ONE pass: each CTA owns a 128 x 128 output tile,
// stages the X / Y tiles k-major in shared memory, every thread accumulates an 8 x 8
// micro-tile of sum_k (x_ik - y_jk)^2 with packed FADD2/FFMA2 (row pairs share an
// instruction, the y operand is a broadcast register), applies the exp2 epilogue in
// registers and writes 128-bit streaming stores (a half-warp covers 256 contiguous
// bytes of one output row).
//
// Why direct differencing: the reference's expansion form loses ~|x|^2 2^-23 in sq,
// which sqrt() turns into 1e-3 of distance near coincident points; the
// difference form is exact at zero, bitwise symmetric for the self-Gram
// (jaxrk/utilities/gram.py:18 asserts G == G.T) and costs 2d lane-ops per evaluation --
// FP32-issue-bound above d ~ 10, HBM-write-bound below (DESIGN.md "Rooflines").
#include "gram_tile.cuh"

namespace jrk {

// OUT: 0 = kernel value (nconst * exp(-dist)), 1 = scaled distance (s*r)^p
template <int PM, int KC, int OUT>
__global__ void __launch_bounds__(GD_THREADS, 2)
gram_direct_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ K,
                   int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy, int64_t ldk,
                   const float* __restrict__ dim_scale, EpiParams ep, int vec_in, int vec_out) {
    __shared__ __align__(16) float sX[KC][GD_LD];
    __shared__ __align__(16) float sY[KC][GD_LD];

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * GD_BM, j0 = (int64_t)blockIdx.x * GD_BN;

    // acc[p][c]: p = row pair (rows 2p, 2p+1 of this thread's 8 rows), c = column 0..7
    f32x2 acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[p][c] = 0ull;

    for (int k0 = 0; k0 < d; k0 += KC) {
        if (k0) __syncthreads();
        gd_fill<KC>(sX, X, N, i0, d, k0, ldx, dim_scale, vec_in & 1);
        gd_fill<KC>(sY, Y, M, j0, d, k0, ldy, dim_scale, vec_in & 2);
        __syncthreads();
        gd_accumulate<KC>(sX, sY, tx, ty, acc);
    }

    // epilogue: rows {ty*4 + a, 64 + ty*4 + a}, cols {tx*4 + b, 64 + tx*4 + b}
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int lr = gd_row(ty, p, h);
            const int64_t gi = i0 + lr;
            if (gi >= N) continue;
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float lo, hi;
                unpack2(acc[p][c], lo, hi);
                float sq = h ? hi : lo;
                if (OUT == 0) {
                    v[c] = ep.nconst * kexp_unnorm<PM, true>(sq, ep);
                } else {
                    if (PM == PM_SQEXP) v[c] = (ep.s * ep.s) * sq;
                    else if (PM == PM_LAPLACE) v[c] = ep.s * sqrtf(sq);
                    else if (PM == PM_GENERAL) v[c] = powf(ep.s * sqrtf(sq), ep.p);
                }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int64_t gj = j0 + g * 64 + tx * 4;
                float* dst = K + gi * ldk + gj;
                if (vec_out && gj + 3 < M) {
                    st_global_cs_v4(dst, make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]));
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (gj + b < M) dst[b] = v[g * 4 + b];
                }
            }
        }
    }
}

template <int PM, int OUT>
static int launch_gd(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx,
                     int64_t ldy, int64_t ldk, const float* dim_scale, const EpiParams& ep, cudaStream_t st) {
    dim3 grid((unsigned)((M + GD_BN - 1) / GD_BN), (unsigned)((N + GD_BM - 1) / GD_BM));
    if (grid.y > 65535) return fail(JRK_ERR_UNSUPPORTED, "gram_direct: N too large for one launch (%lld)", (long long)N);
    int vec_in = 0;
    if (((uintptr_t)X % 16 == 0) && (ldx % 4 == 0)) vec_in |= 1;
    if (((uintptr_t)Y % 16 == 0) && (ldy % 4 == 0)) vec_in |= 2;
    int vec_out = (((uintptr_t)K % 16 == 0) && (ldk % 4 == 0)) ? 1 : 0;
    int kc = gd_pick_kc(d);
#define GD_LAUNCH(KC_)                                                                                   \
    gram_direct_kernel<PM, KC_, OUT><<<grid, GD_THREADS, 0, st>>>(X, Y, K, N, M, d, ldx, ldy, ldk,       \
                                                                  dim_scale, ep, vec_in, vec_out)
    if (kc == 4) GD_LAUNCH(4);
    else if (kc == 8) GD_LAUNCH(8);
    else GD_LAUNCH(16);
#undef GD_LAUNCH
    JRK_LAUNCHED();
    return JRK_OK;
}

// diag = True branch of ScaledPairwiseDistance.__call__ (jaxrk/kern/util.py:112-113): one thread per row
__global__ void dist_diag_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ out,
                                 int64_t N, int d, int64_t ldx, int64_t ldy, float scale,
                                 const float* __restrict__ dim_scale, float power) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) {
        const float ds = dim_scale ? __ldg(dim_scale + k) : 1.f;
        const float df = fabsf(ds * __ldg(X + i * ldx + k) - ds * __ldg(Y + i * ldy + k));
        acc += (power == 2.f) ? df * df : (power == 1.f) ? df : powf(df, power);
    }
    out[i] = scale * acc;
}

int dist_diag(const float* X, const float* Y, float* out, int64_t N, int d, int64_t ldx, int64_t ldy, float scale,
              const float* dim_scale, float power, cudaStream_t st) {
    dist_diag_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(X, Y, out, N, d, ldx, ldy, scale, dim_scale, power);
    JRK_LAUNCHED();
    return JRK_OK;
}

// Host launcher.  out_kind: 0 kernel values, 1 scaled distances.  Rows are processed in
// slabs so that gridDim.y stays within limits for N up to 2^31.
int gram_direct(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, int out_kind,
                const ExtraKind& ex, cudaStream_t st) {
    EpiParams ep = make_epi(scale, power, nconst, ex);
    const int pm = pick_pmode(power, ex);
    const int64_t slab = (int64_t)65535 * GD_BM;
    for (int64_t r = 0; r < N; r += slab) {
        int64_t n = (N - r < slab) ? N - r : slab;
        const float* Xs = X + r * ldx;
        float* Ks = K + r * ldk;
        int rc;
#define GD_CASE(PM_)                                                                                         \
    rc = out_kind == 0 ? launch_gd<PM_, 0>(Xs, Y, Ks, n, M, d, ldx, ldy, ldk, dim_scale, ep, st)             \
                       : launch_gd<PM_, 1>(Xs, Y, Ks, n, M, d, ldx, ldy, ldk, dim_scale, ep, st)
        if (pm == PM_SQEXP) GD_CASE(PM_SQEXP);
        else if (pm == PM_LAPLACE) GD_CASE(PM_LAPLACE);
        else if (pm == PM_GENERAL) GD_CASE(PM_GENERAL);
        else rc = launch_gd<PM_EXTRA, 0>(Xs, Y, Ks, n, M, d, ldx, ldy, ldk, dim_scale, ep, st);   // kernel values only
#undef GD_CASE
        if (rc) return rc;
    }
    return JRK_OK;
}

}  // namespace jrk

// Kernel (g): fused backward of the Gram / K@W path with respect to the points and the scale.
//
// JaxRK exists to be differentiated (README.md:3 "... differentiable ...", examples/Using_FLAX_Optax.ipynb cells
// 6-10, the scipy/L-BFGS helpers jaxrk/rkhs/vector.py:206-222 and jaxrk/utilities/gram.py:62-65): jax.grad of a loss
// through GenGaussKernel.__call__ (jaxrk/kern/rbf.py:104-105) and FiniteVec.inner (jaxrk/rkhs/vector.py:62-75) is what
// XLA derives from the materialised N x M chain.  Here the vector-Jacobian product is ONE fused pass that never
// materialises K (nor dK):
//     forward      out[i, c] = sum_j K_ij W[j, c]                   (jrk_kmatw_f32),   K_ij = nconst exp(-(s r_ij)^p)
//     cotangent    a_ij = sum_c G[i, c] W[j, c]        MODE 0  (K@W: G = dL/dout, N x m)
//                  a_ij = Gbar[i, j]                   MODE 1  (Gram: Gbar = dL/dK, N x M, read once: HBM-bound)
//     dL/dx_i      = sum_j a_ij dK_ij/dx_i = fac * sum_j t_ij (y_j - x_i)
//     dL/ds        = -(fac / s) * sum_ij t_ij r_ij^2
//   with  dK/dx_i = -K p s^p r^(p-2) (x_i - y_j)  and  t_ij = a_ij * (K_ij / nconst) * {1 | 1/r | u/r^2},
//   fac = nconst * {2 s^2 | s | p} for p = 2 | p = 1 | general p (u = (s r)^p).  Pairs at distance zero contribute
//   nothing (for p <= 1 the derivative does not exist there and the reference's jax.grad returns NaN; 0 is the
//   subgradient that keeps the self-Gram usable).
// dL/dy_j is the same kernel with the roles of (X, G) and (Y, W) exchanged; dL/dW = K^T G is the forward kernel.
//
// Layout: one thread per row i (x_i, S1 = sum t y, S0 = sum t, S2 = sum t r^2 in registers), 128 rows per CTA, tiles
// of 32 columns j staged in shared memory (Y tile, W tile or the 128 x 32 block of Gbar) and read as broadcasts;
// the j range is split over blockIdx.y when there are few rows, partial sums are combined in a fixed order by
// kgrad_finalize_kernel (deterministic, no float atomics).  FP32-pipe bound: 3d + m + ~8 lane-ops per pair (MODE 0).
#include "common.cuh"
#include "kmatw_tc2.cuh"

namespace jrk {
namespace tf32 { int tf_sm_count(); }

// kgrad_tc.cu
bool kgrad_tc_eligible(int64_t N, int64_t M, int d, int m, int pm);
bool kgrad_tc_gbar_eligible(int64_t N, int64_t M, int d, int pm, const float* Gbar, int64_t ldgb, int transposed);
size_t kgrad_tc_workspace_bytes(int64_t N, int64_t M);
int kgrad_tc(const float* X, const float* Y, const float* W, const float* G, float* gX, float* gs_rows, int64_t N, int64_t M,
             int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldg, int64_t ldgx, float c, float fac, float fac_s,
             int gtrans, void* scratch, const unsigned** nmax_out, cudaStream_t st);

namespace {

// the regime test of kmatw_tc2.cuh on the two norm maxima of the pre-pass
__device__ __forceinline__ bool kg_fast_regime(const unsigned* nmax) { return tc2_fast_regime(nmax); }

constexpr int KG_ROWS = 128, KG_TJ = 32;

template <int D, int MT, int MODE, int PM>
__global__ void __launch_bounds__(KG_ROWS, 2)
kgrad_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ Y, int64_t ldy, int d,
             const float* __restrict__ W, int64_t ldw, const float* __restrict__ G, int64_t ldg, int m,
             int64_t N, int64_t M, int64_t cols_per_split, EpiParams ep, float* __restrict__ part, int stride,
             const unsigned* __restrict__ skip_if_fast) {
    constexpr bool gtrans = MODE == 2;       // MODE 1 / 2: Gram cotangent as Gbar (N x M) / Gbar^T (M x N)
    // the tensor-core kernel (kgrad_tc.cu) was launched next to this one: inside its regime the result is its to write
    if (skip_if_fast && kg_fast_regime(skip_if_fast)) return;
    // part layout: [split][row][stride], stride = d + 3: {dL/dx contribution (d), S0 (unused by finalize), S2, S3}
    __shared__ __align__(16) float sY[KG_TJ][D];
    __shared__ __align__(16) float sA[(MODE == 0) ? KG_TJ * MT : KG_ROWS * (KG_TJ + 1)];

    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * KG_ROWS + tid;
    const bool iok = i < N;
    // packed FP32x2 over dimension pairs (k, k + 1) and column pairs (c, c + 1): half the issue slots of scalar code
    f32x2 x2[D / 2], s1[D / 2], g2[(MODE == 0) ? MT / 2 : 1];
#pragma unroll
    for (int k = 0; k < D / 2; ++k) {
        const float a0 = (iok && 2 * k < d) ? __ldg(X + i * ldx + 2 * k) : 0.f;
        const float a1 = (iok && 2 * k + 1 < d) ? __ldg(X + i * ldx + 2 * k + 1) : 0.f;
        x2[k] = pack2(a0, a1);
        s1[k] = 0ull;
    }
    if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < MT / 2; ++c) {
            const float a0 = (iok && 2 * c < m) ? __ldg(G + i * ldg + 2 * c) : 0.f;
            const float a1 = (iok && 2 * c + 1 < m) ? __ldg(G + i * ldg + 2 * c + 1) : 0.f;
            g2[c] = pack2(a0, a1);
        }
    }
    float s0 = 0.f, s2 = 0.f, s3 = 0.f;     // s3 (PM_GENERAL only): sum_j t r^2 lg2(s^2 r^2) -> dL/dpower

    const int64_t j_begin = (int64_t)blockIdx.y * cols_per_split;
    const int64_t j_end = (j_begin + cols_per_split < M) ? j_begin + cols_per_split : M;
    // MODE 1: this thread's share of the next 128 x 32 block of Gbar, fetched one tile ahead (the global-load latency
    // would otherwise be exposed once per tile: only 8 warps are resident per SM)
    constexpr int PRE = (MODE >= 1) ? KG_TJ : 1;
    float pre[PRE];
    const int64_t r0 = (int64_t)blockIdx.x * KG_ROWS;
    auto fetch_gbar = [&](int64_t j0) {
        if (MODE >= 1) {
#pragma unroll
            for (int q = 0; q < PRE; ++q) {
                // consecutive threads read consecutive floats: along j for Gbar (N x M), along the rows for Gbar^T (M x N)
                const int idx = q * KG_ROWS + tid;
                const int rr = gtrans ? idx % KG_ROWS : idx / KG_TJ, jj = gtrans ? idx / KG_ROWS : idx % KG_TJ;
                const int64_t r = r0 + rr, j = j0 + jj;
                pre[q] = (r < N && j < j_end) ? __ldg(gtrans ? G + j * ldg + r : G + r * ldg + j) : 0.f;
            }
        }
    };
    fetch_gbar(j_begin);
    for (int64_t j0 = j_begin; j0 < j_end; j0 += KG_TJ) {
        __syncthreads();
        // ---- stage the tile ----
        for (int idx = tid; idx < KG_TJ * D; idx += KG_ROWS) {
            const int jj = idx / D, k = idx % D;
            const int64_t j = j0 + jj;
            sY[jj][k] = (j < j_end && k < d) ? __ldg(Y + j * ldy + k) : 0.f;
        }
        if (MODE == 0) {
            for (int idx = tid; idx < KG_TJ * MT; idx += KG_ROWS) {
                const int jj = idx / MT, c = idx % MT;
                const int64_t j = j0 + jj;
                sA[jj * MT + c] = (j < j_end && c < m) ? __ldg(W + j * ldw + c) : 0.f;
            }
        } else {
            // 128 x 32 block of Gbar: coalesced over j, rows of this CTA; padded row stride 33 -> conflict-free reads
#pragma unroll
            for (int q = 0; q < PRE; ++q) {
                const int idx = q * KG_ROWS + tid;
                const int rr = gtrans ? idx % KG_ROWS : idx / KG_TJ, jj = gtrans ? idx / KG_ROWS : idx % KG_TJ;
                sA[rr * (KG_TJ + 1) + jj] = pre[q];
            }
            if (j0 + KG_TJ < j_end) fetch_gbar(j0 + KG_TJ);
        }
        __syncthreads();
        const int nj = (int)((j_end - j0 < KG_TJ) ? j_end - j0 : KG_TJ);
#pragma unroll 2
        for (int jj = 0; jj < nj; ++jj) {
            const ulonglong2* yrow = reinterpret_cast<const ulonglong2*>(&sY[jj][0]);
            f32x2 y2[D / 2];
            f32x2 sq2 = 0ull;
#pragma unroll
            for (int k = 0; k < D / 4; ++k) {
                const ulonglong2 v = yrow[k];
                y2[2 * k] = v.x;
                y2[2 * k + 1] = v.y;
            }
#pragma unroll
            for (int k = 0; k < D / 2; ++k) {
                const f32x2 df = sub2(x2[k], y2[k]);
                sq2 = fma2(df, df, sq2);
            }
            float sqa, sqb;
            unpack2(sq2, sqa, sqb);
            const float sq = sqa + sqb;
            float a;
            if (MODE == 0) {
                const ulonglong2* wrow = reinterpret_cast<const ulonglong2*>(&sA[jj * MT]);
                f32x2 a2 = 0ull;
#pragma unroll
                for (int c = 0; c < MT / 4; ++c) {
                    const ulonglong2 v = wrow[c];
                    a2 = fma2(g2[2 * c], v.x, a2);
                    a2 = fma2(g2[2 * c + 1], v.y, a2);
                }
                float aa, ab;
                unpack2(a2, aa, ab);
                a = aa + ab;
            } else {
                a = sA[tid * (KG_TJ + 1) + jj];
            }
            // t = a * (K / nconst) * {1 | 1/r | u / r^2}
            float t;
            if (PM == PM_SQEXP) {
                t = a * ex2_approx(sq * ep.c);
            } else if (PM == PM_LAPLACE) {
                const float r = sqrt_approx(sq);
                t = (sq > 0.f) ? a * ex2_approx(r * ep.c) / r : 0.f;
            } else {
                const float lg = lg2_approx(sq * ep.c);                        // log2((s r)^2)
                const float u = ex2_approx(ep.hp * lg);                        // (s r)^p
                t = (sq > 0.f) ? a * ex2_approx(u * -kLog2e) * u / sq : 0.f;
                // dK/dp = -K u ln(s r) at fixed nconst:  a (K / nconst) u = t r^2,  ln(s r) = (ln 2 / 2) lg
                s3 = (sq > 0.f) ? fmaf(t * sq, lg, s3) : s3;
            }
            s0 += t;
            s2 = fmaf(t, sq, s2);
            const f32x2 t2 = pack2(t, t);
#pragma unroll
            for (int k = 0; k < D / 2; ++k) s1[k] = fma2(t2, y2[k], s1[k]);
        }
    }
    if (iok) {
        float* dst = part + ((int64_t)blockIdx.y * N + i) * stride;
#pragma unroll
        for (int k = 0; k < D / 2; ++k) {                     // sum_j t (y_j - x_i)
            float xa, xb, sa, sb;
            unpack2(x2[k], xa, xb);
            unpack2(s1[k], sa, sb);
            if (2 * k < d) dst[2 * k] = fmaf(-xa, s0, sa);
            if (2 * k + 1 < d) dst[2 * k + 1] = fmaf(-xb, s0, sb);
        }
        dst[d] = s0;
        dst[d + 1] = s2;
        dst[d + 2] = s3;
    }
}

// gX[i, k] = fac * sum_split part[split][i][k];  gs_rows[i] = -(fac / s) * sum_split part[split][i][d + 1];
// gp_rows[i] = fac_p * sum_split part[split][i][d + 2]
__global__ void kgrad_finalize_kernel(const float* __restrict__ part, int nsplit, int64_t N, int d, int stride, float fac,
                                      float fac_s, float fac_p, float* __restrict__ gX, int64_t ldgx,
                                      float* __restrict__ gs_rows, float* __restrict__ gp_rows,
                                      const unsigned* __restrict__ skip_if_fast) {
    if (skip_if_fast && kg_fast_regime(skip_if_fast)) return;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = idx / (d + 2);
    const int k = (int)(idx % (d + 2));
    if (i >= N) return;
    if ((k == d && !gs_rows) || (k == d + 1 && !gp_rows)) return;
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += part[((int64_t)s * N + i) * stride + (k < d ? k : k + 1)];
    if (k < d) gX[i * ldgx + k] = fac * acc;
    else if (k == d) gs_rows[i] = fac_s * acc;
    else gp_rows[i] = fac_p * acc;
}

template <int D, int MT, int MODE, int PM>
int kg_launch(dim3 grid, cudaStream_t st, const float* X, int64_t ldx, const float* Y, int64_t ldy, int d, const float* W,
              int64_t ldw, const float* G, int64_t ldg, int m, int64_t N, int64_t M, int64_t cps, const EpiParams& ep,
              float* part, int stride, const unsigned* skip) {
    kgrad_kernel<D, MT, MODE, PM><<<grid, KG_ROWS, 0, st>>>(X, ldx, Y, ldy, d, W, ldw, G, ldg, m, N, M, cps, ep, part, stride, skip);
    JRK_LAUNCHED();
    return JRK_OK;
}

int kg_sm_count() { return tf32::tf_sm_count(); }   // per-device cache (gram_tf32.cu)

struct KgPlan {
    int nsplit;
    int64_t cols_per_split;
};
KgPlan kg_plan(int64_t N, int64_t M) {
    const int64_t nblk = (N + KG_ROWS - 1) / KG_ROWS;
    int64_t want = ((int64_t)kg_sm_count() * 4 + nblk - 1) / nblk;       // ~4 CTAs per SM in flight
    const int64_t max_split = (M + 4 * KG_TJ - 1) / (4 * KG_TJ);          // at least 4 tiles per split
    if (want > max_split) want = max_split;
    if (want > 1024) want = 1024;
    if (want < 1) want = 1;
    KgPlan p;
    p.cols_per_split = ((M + want - 1) / want + KG_TJ - 1) / KG_TJ * KG_TJ;
    p.nsplit = (int)((M + p.cols_per_split - 1) / p.cols_per_split);
    return p;
}

}  // namespace

size_t kgrad_workspace_bytes(int64_t N, int64_t M, int d) {
    const KgPlan p = kg_plan(N, M);
    size_t b = align256((size_t)p.nsplit * (size_t)N * (size_t)(d + 3) * 4) + 512;
    if (kgrad_tc_eligible(N, M, d, 1, PM_SQEXP)) b += kgrad_tc_workspace_bytes(N, M);
    return b;
}

// mode 0: K@W cotangent (W: M x m, G: N x m);  mode 1: Gram cotangent (G = Gbar: N x M, W unused, m = 0)
// gp_rows (nullable): per-row contributions to dL/dpower at fixed nconst; asking for it selects the general-p epilogue
// (the p = 2 / p = 1 specialisations never form log(s r)).
// gtrans (mode 1 only): the cotangent is given as Gbar^T (M x N, leading dimension ldg >= N).
int kgrad(int mode, const float* X, const float* Y, const float* W, const float* G, float* gX, float* gs_rows, float* gp_rows,
          int64_t N, int64_t M, int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldg, int64_t ldgx, float scale,
          float power, float nconst, int gtrans, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (d > 64) return fail(JRK_ERR_UNSUPPORTED, "kgrad: d = %d > 64 is not supported by the fused backward kernel", d);
    const int pm = gp_rows ? PM_GENERAL : pick_pmode(power);
    EpiParams ep = make_epi(scale, power, nconst);
    if (gp_rows && pick_pmode(power) != PM_GENERAL) {       // the general epilogue's constants for p = 2 / p = 1
        ep.c = scale * scale;
        ep.hp = 0.5f * power;
    }
    const KgPlan p = kg_plan(N, M);
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(kgrad_workspace_bytes(N, M, d)));
    const int stride = d + 3;
    float* part = scr.take<float>((size_t)p.nsplit * N * stride);
    const double s = (double)scale;
    const float fac = (float)((double)nconst * (pm == PM_SQEXP ? 2.0 * s * s : pm == PM_LAPLACE ? s : (double)power));
    const float fac_s = (float)(-(double)fac / s);
    const float fac_p = (float)(-(double)nconst * 0.5 * 0.6931471805599453);
    // K@W cotangent, p = 2, d <= 16, m <= 16, large: both inner products on the tensor cores (kgrad_tc.cu).  Whether the
    // problem lies inside its regime is known on the device only, so the FP32 kernels below are enqueued as well and
    // return at once when it does.
    const unsigned* skip = nullptr;
    if ((mode == 0 && kgrad_tc_eligible(N, M, d, m, pm)) || (mode == 1 && kgrad_tc_gbar_eligible(N, M, d, pm, G, ldg, gtrans))) {
        void* tcs = scr.take<char>(kgrad_tc_workspace_bytes(N, M));
        if (int rc = kgrad_tc(X, Y, mode == 0 ? W : nullptr, G, gX, gs_rows, N, M, d, m, ldx, ldy, ldw, ldg, ldgx, ep.c, fac,
                              fac_s, gtrans, tcs, &skip, st))
            return rc;
    }
    const int64_t nblk = (N + KG_ROWS - 1) / KG_ROWS;
    if (nblk > 2147483647LL) return fail(JRK_ERR_UNSUPPORTED, "kgrad: too many rows");
    dim3 grid((unsigned)nblk, (unsigned)p.nsplit);
    const int D = d <= 8 ? 8 : d <= 16 ? 16 : d <= 32 ? 32 : 64;

    if (mode == 0 && m > 32)   // (the caller splits wide G / W into column blocks: the gradient is additive over them)
        return fail(JRK_ERR_UNSUPPORTED, "kgrad: m = %d > 32 is not supported by the fused backward kernel", m);
    const int MT = m <= 4 ? 4 : m <= 16 ? 16 : 32;
    int rc;
#define KG_GO(D_, MT_, MODE_, PM_) \
    rc = kg_launch<D_, MT_, MODE_, PM_>(grid, st, X, ldx, Y, ldy, d, W, ldw, G, ldg, m, N, M, p.cols_per_split, ep, part, stride, skip)
#define KG_MT(D_, PM_)                                   \
    do {                                                 \
        if (mode == 1 && gtrans) KG_GO(D_, 4, 2, PM_);   \
        else if (mode == 1) KG_GO(D_, 4, 1, PM_);        \
        else if (MT == 4) KG_GO(D_, 4, 0, PM_);          \
        else if (MT == 16) KG_GO(D_, 16, 0, PM_);        \
        else KG_GO(D_, 32, 0, PM_);                      \
    } while (0)
#define KG_D(PM_)                                        \
    do {                                                 \
        if (D == 8) KG_MT(8, PM_);                       \
        else if (D == 16) KG_MT(16, PM_);                \
        else if (D == 32) KG_MT(32, PM_);                \
        else KG_MT(64, PM_);                             \
    } while (0)
    if (pm == PM_SQEXP) KG_D(PM_SQEXP);
    else if (pm == PM_LAPLACE) KG_D(PM_LAPLACE);
    else KG_D(PM_GENERAL);
#undef KG_D
#undef KG_MT
#undef KG_GO
    if (rc) return rc;
    const int64_t tot = N * (d + 2);
    kgrad_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(part, p.nsplit, N, d, stride, fac, fac_s, fac_p, gX, ldgx, gs_rows, gp_rows, skip);
    JRK_LAUNCHED();
    return JRK_OK;
}

}  // namespace jrk

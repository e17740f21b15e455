// Kernel (b)+(c): fused, never-materialised  T = diag(row_pref) K(X,Y) diag(col_pref) W
// with the row reduce maps (Sum / Mean / BalancedRed) folded into the epilogue.
//
// Replaces FiniteVec.inner (jaxrk/rkhs/vector.py:62-75), which materialises
// K = k(X, Y) (N x M) and then applies Reduce.apply on axis 0 and 1
// (jaxrk/reduce/base.py:39-47; LinearReduce is `A @ K`, jaxrk/reduce/lincomb.py:137-140),
// and the inner() inside FiniteOp.__matmul__ (jaxrk/rkhs/operator.py:46).
//
// Design (KeOps-style, one pass over Y/W per row block, nothing N x M ever stored):
//   * a thread owns RI rows of X: their coordinates live in registers as packed pairs
//     over k, their m output accumulators too -- no cross-thread reduction is needed
//     for the plain product;
//   * Y and W tiles (TJ rows) stream through a 2-stage shared-memory ring filled by the
//     TMA engine (cp.async.bulk + mbarrier); every lane reads the same y_j / w_j, so the
//     loads are broadcast LDS.128;
//   * per pair: d/2 FADD2 + d/2 FFMA2 (direct differencing, packed over k), one FADD,
//     the exp2 epilogue (1-3 MUFU), m/2 FFMA2 with a broadcast operand;
//   * Sum / Mean / aligned BalancedRed: rows are combined in registers, then by warp
//     shuffles and a per-CTA shared-memory pass; one partial per CTA, summed in a fixed
//     order by a finalize kernel (deterministic; no float atomics);
//   * few rows (N small): the j range is split over blockIdx.y and the partials are
//     summed by the same finalize kernel;
//   * long j ranges: a float32 running sum of 2^20 positive terms is off by ~2e-5 relative
//     (measured), outside the 1e-5 tolerance, so every 32 tiles (4096 j) the register
//     accumulators are folded into a per-thread second-level accumulator kept in shared
//     memory (conflict-free 16-byte slots) and restarted from zero.
// Roofline: FP32 issue (2d + m + ~3 lane-ops per pair), see DESIGN.md.
#include <cstdlib>

#include "kmatw_kernel.cuh"

namespace jrk {

// tensor-core variant (kmatw_tc.cu)
bool kmatw_tc_eligible(int64_t N, int64_t M, int d, int m, const float* dim_scale, int row_reduce, float nconst);
size_t kmatw_tc_workspace_bytes(int64_t N, int64_t M);
size_t kmatw_tc_rowsum_bytes(int64_t N);
int kmatw_tc_rowsum(const float* T, int64_t N, int m, float scale, float* out, float* rpart, const unsigned* only_if_fast,
                    cudaStream_t st);
size_t kmatw_tc64_workspace_bytes(int64_t N, int64_t M);
int kmatw_tc64(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m, int64_t ldx,
               int64_t ldy, int64_t ldw, int64_t ldo, float c, float nconst, const float* row_pref, const float* col_pref,
               void* scratch, const unsigned** nmax_out, cudaStream_t st, int m_w = -1);
int kmatw_tc(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m, int64_t ldx,
             int64_t ldy, int64_t ldw, int64_t ldo, float scale, float power, float nconst, const float* row_pref,
             const float* col_pref, int row_reduce, void* workspace, size_t workspace_bytes, cudaStream_t st);

// out[g, c] = scale * sum_{s < nsplit} sum_{r in [g*group, (g+1)*group)} part[s][r][c]
// (fixed summation order: deterministic).  part is [nsplit][rows][MT].
__global__ void kw_finalize_kernel(const float* __restrict__ part, int nsplit, int64_t rows, int MT,
                                   int64_t group, float scale, float* __restrict__ out, int64_t ldo, int m,
                                   int64_t ngroups, const unsigned* __restrict__ skip_if_fast) {
    if (skip_if_fast && tc2_fast_regime(skip_if_fast)) return;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ngroups * m) return;
    const int64_t g = idx / m;
    const int c = (int)(idx % m);
    float tot = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float* p = part + ((int64_t)s * rows + g * group) * MT + c;
        float a = 0.f;
        for (int64_t r = 0; r < group; ++r) a += p[r * MT];
        tot += a;
    }
    out[g * ldo + c] = scale * tot;
}

// Packs rows of A (rows x cols, lda) into a dense rows x COLS buffer, zero padded,
// optionally scaling columns (dim_scale, len cols) or rows (row_scale, len rows).
__global__ void kw_pack_kernel(const float* __restrict__ A, int64_t rows, int cols, int64_t lda,
                               const float* __restrict__ col_scale, const float* __restrict__ row_scale,
                               float* __restrict__ out, int COLS) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * COLS) return;
    const int64_t r = idx / COLS;
    const int c = (int)(idx % COLS);
    float v = 0.f;
    if (c < cols) {
        v = A[r * lda + c];
        if (col_scale) v *= col_scale[c];
        if (row_scale) v *= row_scale[r];
    }
    out[idx] = v;
}

struct KwPlan {
    int D, MT, RI, R;
    int64_t nblk;
    int nsplit, tiles_per_split, num_tiles;
};

static int kw_pick_D(int d) {
    const int ds[] = {4, 8, 16, 32, 64, 128};
    for (int v : ds)
        if (d <= v) return v;
    return -1;
}
static int kw_pick_MT(int m) { return m <= 4 ? 4 : m <= 16 ? 16 : 32; }

static KwPlan kw_plan(int64_t N, int64_t M, int d, int m) {
    KwPlan p;
    p.D = kw_pick_D(d);
    p.MT = kw_pick_MT(m);
    p.RI = kw_ri(p.D, p.MT);
    p.R = KW_THREADS * p.RI;
    p.nblk = (N + p.R - 1) / p.R;
    p.num_tiles = (int)((M + KW_TJ - 1) / KW_TJ);
    // enough CTAs for >= 4 per SM on 148 SMs; never more splits than tiles
    int64_t want = (592 + p.nblk - 1) / p.nblk;
    if (want < 1) want = 1;
    if (want > p.num_tiles) want = p.num_tiles;
    if (want > 65535) want = 65535;
    p.tiles_per_split = (int)((p.num_tiles + want - 1) / want);
    p.nsplit = (p.num_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    if (p.nsplit < 1) p.nsplit = 1;
    return p;
}

size_t kmatw_workspace_bytes(int64_t N, int64_t M, int d, int m, int row_reduce, int64_t group) {
    if (d > 128 || d < 1 || m < 1) return 0;
    size_t total = 0;
    for (int c0 = 0; c0 < m; c0 += 32) {  // widest column block dominates; sum is a safe bound
        int mb = (m - c0 < 32) ? m - c0 : 32;
        KwPlan p = kw_plan(N, M, d, mb);
        size_t b = align256((size_t)M * p.D * 4) + align256((size_t)M * p.MT * 4);
        b += align256((size_t)p.nsplit * (size_t)((N > p.nblk) ? N : p.nblk) * p.MT * 4);
        if (b > total) total = b;
    }
    (void)group;
    if (kmatw_tc_eligible(N, M, d, m, nullptr, row_reduce, 1.f)) {
        const size_t t = kmatw_tc_workspace_bytes(N, M);
        if (t > total) total = t;
    }
    if (d > 32 && d <= 64 && N >= 8192 && M >= 4096)
        total += kmatw_tc64_workspace_bytes(N, M) + align256((size_t)N * 64) + kmatw_tc_rowsum_bytes(N) + 1024;
    return total + 1024;
}

// One column block (m <= 32).
static int kmatw_block(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d,
                       int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale,
                       const float* dim_scale, float power, float nconst, const float* row_pref,
                       const float* col_pref, int row_reduce, int64_t group, const ExtraKind& ex, Scratch& scr,
                       cudaStream_t st, const unsigned* skip_if_fast = nullptr) {
    const KwPlan p = kw_plan(N, M, d, m);
    EpiParams ep = make_epi(scale, power, nconst, ex);
    ep.skip_if_fast = skip_if_fast;
    const int pm = pick_pmode(power, ex);

    // ---- operands: dense, padded, pre-scaled copies where the caller's layout differs -----
    const float* Yp = Y;
    if (!(ldy == p.D && d == p.D && !dim_scale && ((uintptr_t)Y % 16 == 0))) {
        float* buf = scr.take<float>((size_t)M * p.D);
        const int64_t tot = M * p.D;
        kw_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, dim_scale, nullptr, buf, p.D);
        JRK_LAUNCHED();
        Yp = buf;
    }
    const float* Wp = W;
    if (!(ldw == p.MT && m == p.MT && !col_pref && ((uintptr_t)W % 16 == 0))) {
        float* buf = scr.take<float>((size_t)M * p.MT);
        const int64_t tot = M * p.MT;
        kw_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(W, M, m, ldw, nullptr, col_pref, buf, p.MT);
        JRK_LAUNCHED();
        Wp = buf;
    }
    const int x_vec = (((uintptr_t)X % 16 == 0) && (ldx % 4 == 0)) ? 1 : 0;

    // ---- which epilogue -----------------------------------------------------------------
    const bool reduce_all = (row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN);
    const bool balanced = (row_reduce == JRK_REDUCE_BALANCED || row_reduce == JRK_REDUCE_BALANCED_MEAN);
    const bool cta_partials = reduce_all || (balanced && group % p.R == 0);
    const int epi_mode = cta_partials ? 1 : 0;
    const bool direct = (!cta_partials && !balanced && p.nsplit == 1);

    float* kout;
    int64_t kld, kstride, part_rows;
    if (direct) {
        kout = out; kld = ldo; kstride = 0; part_rows = N;
    } else {
        part_rows = cta_partials ? p.nblk : N;
        kout = scr.take<float>((size_t)p.nsplit * part_rows * p.MT);
        kld = p.MT; kstride = part_rows * p.MT;
    }

    dim3 grid((unsigned)p.nblk, (unsigned)p.nsplit);
    int rc = JRK_OK;
#define KW_PM(PM_) rc = kw_dispatch<PM_>(p.D, p.MT, grid, st, X, ldx, d, dim_scale, Yp, Wp, kout, kld, kstride, N, M, m, \
                                         p.tiles_per_split, row_pref, ep, epi_mode, x_vec)
    if (pm == PM_SQEXP) KW_PM(PM_SQEXP);
    else if (pm == PM_LAPLACE) KW_PM(PM_LAPLACE);
    else if (pm == PM_GENERAL) KW_PM(PM_GENERAL);
    else KW_PM(PM_EXTRA);
#undef KW_PM
    if (rc) return rc;

    if (!direct) {
        int64_t fgroup, ngroups;
        float fscale = 1.f;
        if (reduce_all) {
            fgroup = part_rows; ngroups = 1;
            if (row_reduce == JRK_REDUCE_MEAN) fscale = 1.f / (float)N;
        } else if (balanced) {
            fgroup = cta_partials ? group / p.R : group;
            ngroups = N / group;
            if (row_reduce == JRK_REDUCE_BALANCED_MEAN) fscale = 1.f / (float)group;
        } else {
            fgroup = 1; ngroups = N;
        }
        const int64_t tot = ngroups * m;
        kw_finalize_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(kout, p.nsplit, part_rows, p.MT, fgroup,
                                                                          fscale, out, ldo, m, ngroups, skip_if_fast);
        JRK_LAUNCHED();
    }
    return JRK_OK;
}

// The dispatch rule, in one place (also behind jrk_kmatw_uses_tensor_cores): JRK_OPT_KMATW_TC = 0 forces the FP32-pipe
// kernel.  Wide W (16 < m <= TC_WIDE_M_MAX) runs on the tensor cores too, in column blocks of 16.
bool kmatw_uses_tc(int64_t N, int64_t M, int d, int m, bool has_dim_scale, int row_reduce, const ExtraKind& ex) {
    if (jrk_option(JRK_OPT_KMATW_TC) == 0) return false;
    if (ex.kind != 0) return false;   // sibling kernels: FP32-pipe kernel only
    static const float dummy = 1.f;
    return kmatw_tc_eligible(N, M, d, m, has_dim_scale ? &dummy : nullptr, row_reduce, 1.f);
}

// 32 < d <= 64, p = 2, plain problem: the second-generation tensor-core kernel with four GEMM1 k-steps serves the factored
// regime (decided on the device); the FP32-pipe kernel is enqueued behind it and returns at once there.
bool kmatw_tc64_eligible(int64_t N, int64_t M, int d, int m, bool has_dim_scale, float power, int row_reduce, const ExtraKind& ex) {
    return jrk_option(JRK_OPT_KMATW_TC) != 0 && jrk_option(JRK_OPT_TC_FAST) != 0 && jrk_option(JRK_OPT_TC_V2) != 0 &&
           ex.kind == 0 && !has_dim_scale && pick_pmode(power) == PM_SQEXP && d > 32 && d <= 64 && m >= 1 &&
           (row_reduce == JRK_REDUCE_NONE || row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN) && N >= 8192 &&
           M >= 4096 && M <= ((int64_t)1 << 30);
}

int kmatw(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
          int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale, float power,
          float nconst, const float* row_pref, const float* col_pref, int row_reduce, int64_t group,
          const ExtraKind& ex, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (d > 128) return fail(JRK_ERR_UNSUPPORTED, "kmatw: d = %d > 128 not supported by the fused kernel", d);
    if (kmatw_tc64_eligible(N, M, d, m, dim_scale != nullptr, power, row_reduce, ex)) {
        Scratch scr(workspace, workspace_bytes, st);
        JRK_CUDA_OK(scr.reserve(kmatw_workspace_bytes(N, M, d, m, row_reduce, group)));
        void* tcs = scr.take<char>(kmatw_tc64_workspace_bytes(N, M));
        // Sum / Mean over rows: the tensor-core kernel writes the N x 16 block to scratch, reduced in a fixed order -- and
        // only inside the factored regime; outside it the FP32-pipe kernel below writes the reduced result itself
        const bool reduce_rows = row_reduce != JRK_REDUCE_NONE;
        float* Tbuf = reduce_rows ? scr.take<float>((size_t)N * 16) : nullptr;
        float* rpart = reduce_rows ? (float*)scr.take<char>(kmatw_tc_rowsum_bytes(N)) : nullptr;
        const size_t base = scr.used;
        const EpiParams ep = make_epi(scale, power, nconst);
        for (int c0 = 0; c0 < m; c0 += 16) {        // the tensor-core kernel: column blocks of 16
            const int mb = (m - c0 < 16) ? m - c0 : 16;
            const unsigned* skip = nullptr;
            if (int rc = kmatw_tc64(X, Y, W + c0, reduce_rows ? Tbuf : out + c0, N, M, d, reduce_rows ? 16 : mb, ldx, ldy, ldw,
                                    reduce_rows ? 16 : ldo, ep.c, nconst, row_pref, col_pref, tcs, &skip, st, mb))
                return rc;
            if (reduce_rows)
                if (int rc = kmatw_tc_rowsum(Tbuf, N, mb, row_reduce == JRK_REDUCE_MEAN ? 1.f / (float)N : 1.f, out + c0, rpart,
                                             skip, st))
                    return rc;
            scr.used = base;
            if (int rc = kmatw_block(X, Y, W + c0, out + c0, N, M, d, mb, ldx, ldy, ldw, ldo, scale, dim_scale, power, nconst,
                                     row_pref, col_pref, row_reduce, group, ex, scr, st, skip))
                return rc;
        }
        return JRK_OK;
    }
    // both contractions on the tensor cores when the problem is large, narrow and plain (kmatw_tc.cu)
    if (kmatw_uses_tc(N, M, d, m, dim_scale != nullptr, row_reduce, ex))
        return kmatw_tc(X, Y, W, out, N, M, d, m, ldx, ldy, ldw, ldo, scale, power, nconst, row_pref, col_pref,
                        row_reduce, workspace, workspace_bytes, st);
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(kmatw_workspace_bytes(N, M, d, m, row_reduce, group)));
    for (int c0 = 0; c0 < m; c0 += 32) {  // wide W: column blocks of 32 (K is recomputed per block)
        const int mb = (m - c0 < 32) ? m - c0 : 32;
        scr.used = 0;
        int rc = kmatw_block(X, Y, W + c0, out + c0, N, M, d, mb, ldx, ldy, ldw, ldo, scale, dim_scale, power,
                             nconst, row_pref, col_pref, row_reduce, group, ex, scr, st);
        if (rc) return rc;
    }
    return JRK_OK;
}

}  // namespace jrk

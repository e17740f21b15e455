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

#include "common.cuh"

namespace jrk {

// tensor-core variant (kmatw_tc.cu)
bool kmatw_tc_eligible(int64_t N, int64_t M, int d, int m, const float* dim_scale, int row_reduce, float nconst);
size_t kmatw_tc_workspace_bytes(int64_t N, int64_t M);
int kmatw_tc(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m, int64_t ldx,
             int64_t ldy, int64_t ldw, int64_t ldo, float scale, float power, float nconst, const float* row_pref,
             const float* col_pref, int row_reduce, void* workspace, size_t workspace_bytes, cudaStream_t st);

constexpr int KW_THREADS = 128;
constexpr int KW_TJ = 128;  // rows of Y / W per stage
constexpr int KW_FLUSH_TILES = 32;  // fold the register accumulators into the second level every 32 tiles

template <int D, int MT, int RI = 0>
struct KwSmem {
    static constexpr int kStageFloats = KW_TJ * (D + MT);
    static constexpr size_t kRingBytes = 2 * kStageFloats * sizeof(float);
    static constexpr size_t kBarOff = kRingBytes;
    static constexpr size_t kAccOff = kRingBytes + 16;                       // second-level accumulators
    static constexpr size_t kBytes = kAccOff + (size_t)RI * MT * KW_THREADS * sizeof(float);
};

// epi_mode: 0 = write rows (out ld = ldo, rows N), 1 = write one partial per CTA
template <int D, int MT, int RI, int PM>
__global__ void __launch_bounds__(KW_THREADS)
kmatw_kernel(const float* __restrict__ X, int64_t ldx, int d, const float* __restrict__ dim_scale,
             const float* __restrict__ Yp, const float* __restrict__ Wp,
             float* __restrict__ out, int64_t ldo, int64_t out_split_stride,
             int64_t N, int64_t M, int m, int tiles_per_split,
             const float* __restrict__ row_pref, EpiParams ep, int epi_mode, int x_vec) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + KwSmem<D, MT, RI>::kBarOff);
    // second-level accumulators: slot s of thread t at float4 index s * KW_THREADS + t (conflict-free)
    float4* acc2 = reinterpret_cast<float4*>(smem_raw + KwSmem<D, MT, RI>::kAccOff) + threadIdx.x;

    constexpr int R = KW_THREADS * RI;
    const int64_t i0 = (int64_t)blockIdx.x * R;
    const int num_tiles = (int)((M + KW_TJ - 1) / KW_TJ);
    const int t0 = blockIdx.y * tiles_per_split;
    const int t1 = min(t0 + tiles_per_split, num_tiles);

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int stage) {
        const int64_t j0 = (int64_t)tile * KW_TJ;
        const int rows = (int)min((int64_t)KW_TJ, M - j0);
        float* sy = stage_base + stage * KwSmem<D, MT>::kStageFloats;
        float* sw = sy + KW_TJ * D;
        const uint32_t by = rows * D * sizeof(float), bw = rows * MT * sizeof(float);
        mbar_arrive_expect_tx(&full[stage], by + bw);
        bulk_g2s(sy, Yp + j0 * D, by, &full[stage]);
        bulk_g2s(sw, Wp + j0 * MT, bw, &full[stage]);
    };
    if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);

    // ---- this thread's RI rows of X, packed over k ---------------------------------
    f32x2 x2[RI][D / 2];
    float pref[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
        const int64_t gi = i0 + threadIdx.x + r * KW_THREADS;
        float xv[D];
#pragma unroll
        for (int k = 0; k < D; ++k) xv[k] = 0.f;
        if (gi < N) {
            const float* src = X + gi * ldx;
            if (x_vec && d == D) {
#pragma unroll
                for (int q = 0; q < D / 4; ++q) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(src) + q);
                    xv[q * 4] = t.x; xv[q * 4 + 1] = t.y; xv[q * 4 + 2] = t.z; xv[q * 4 + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) xv[k] = __ldg(src + k);
            }
            if (dim_scale) {
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) xv[k] *= __ldg(dim_scale + k);
            }
            pref[r] = row_pref ? __ldg(row_pref + gi) : 1.f;
        } else {
            pref[r] = 0.f;  // padded rows contribute nothing to reductions
        }
#pragma unroll
        for (int k = 0; k < D / 2; ++k) x2[r][k] = pack2(xv[2 * k], xv[2 * k + 1]);
    }

    f32x2 acc[RI][MT / 2];
#pragma unroll
    for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < MT / 2; ++c) acc[r][c] = 0ull;
#pragma unroll
    for (int sl = 0; sl < RI * MT / 4; ++sl) acc2[sl * KW_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
    // acc2 += acc; acc = 0   (slot r * MT/4 + q holds columns 4q..4q+3 of row r)
    auto fold = [&]() {
#pragma unroll
        for (int r = 0; r < RI; ++r)
#pragma unroll
            for (int q = 0; q < MT / 4; ++q) {
                float4 v = acc2[(r * (MT / 4) + q) * KW_THREADS];
                float a0, a1, a2, a3;
                unpack2(acc[r][2 * q], a0, a1);
                unpack2(acc[r][2 * q + 1], a2, a3);
                v.x += a0; v.y += a1; v.z += a2; v.w += a3;
                acc2[(r * (MT / 4) + q) * KW_THREADS] = v;
                acc[r][2 * q] = 0ull;
                acc[r][2 * q + 1] = 0ull;
            }
    };

    // ---- stream Y / W tiles ------------------------------------------------------------
    for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int stage = it & 1;
        if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, stage ^ 1);
        mbar_wait(&full[stage], (it >> 1) & 1);
        const float* sy = stage_base + stage * KwSmem<D, MT>::kStageFloats;
        const float* sw = sy + KW_TJ * D;
        const int rows = (int)min((int64_t)KW_TJ, M - (int64_t)t * KW_TJ);
#pragma unroll 8
        for (int j = 0; j < rows; ++j) {
            f32x2 sq2[RI];
#pragma unroll
            for (int r = 0; r < RI; ++r) sq2[r] = 0ull;
#pragma unroll
            for (int q = 0; q < D / 4; ++q) {
                const float4 yv = *reinterpret_cast<const float4*>(sy + j * D + q * 4);
                const f32x2 y01 = pack2(yv.x, yv.y), y23 = pack2(yv.z, yv.w);
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    const f32x2 d0 = sub2(x2[r][2 * q], y01);
                    sq2[r] = fma2(d0, d0, sq2[r]);
                    const f32x2 d1 = sub2(x2[r][2 * q + 1], y23);
                    sq2[r] = fma2(d1, d1, sq2[r]);
                }
            }
            f32x2 e2[RI];
#pragma unroll
            for (int r = 0; r < RI; ++r) {
                float lo, hi;
                unpack2(sq2[r], lo, hi);
                const float e = kexp_unnorm<PM, true>(lo + hi, ep);
                e2[r] = pack2(e, e);
            }
#pragma unroll
            for (int q = 0; q < MT / 4; ++q) {
                const float4 wv = *reinterpret_cast<const float4*>(sw + j * MT + q * 4);
                const f32x2 w01 = pack2(wv.x, wv.y), w23 = pack2(wv.z, wv.w);
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    acc[r][2 * q] = fma2(e2[r], w01, acc[r][2 * q]);
                    acc[r][2 * q + 1] = fma2(e2[r], w23, acc[r][2 * q + 1]);
                }
            }
        }
        if ((it + 1) % KW_FLUSH_TILES == 0) fold();
        __syncthreads();  // stage fully consumed before it is refilled
    }
    fold();   // totals now live in the second-level accumulators
#pragma unroll
    for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int q = 0; q < MT / 4; ++q) {
            const float4 v = acc2[(r * (MT / 4) + q) * KW_THREADS];
            acc[r][2 * q] = pack2(v.x, v.y);
            acc[r][2 * q + 1] = pack2(v.z, v.w);
        }

    // ---- epilogue -------------------------------------------------------------------------
    float* outp = out + (int64_t)blockIdx.y * out_split_stride;
    if (epi_mode == 0) {
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const int64_t gi = i0 + threadIdx.x + r * KW_THREADS;
            if (gi >= N) continue;
            const float f = ep.nconst * pref[r];
            float* dst = outp + gi * ldo;
#pragma unroll
            for (int c = 0; c < MT / 2; ++c) {
                float lo, hi;
                unpack2(acc[r][c], lo, hi);
                if (2 * c < m) dst[2 * c] = f * lo;
                if (2 * c + 1 < m) dst[2 * c + 1] = f * hi;
            }
        }
    } else {
        __shared__ float red[KW_THREADS / 32][MT];
        float v[MT];
#pragma unroll
        for (int c = 0; c < MT; ++c) v[c] = 0.f;
#pragma unroll
        for (int r = 0; r < RI; ++r) {
#pragma unroll
            for (int c = 0; c < MT / 2; ++c) {
                float lo, hi;
                unpack2(acc[r][c], lo, hi);
                v[2 * c] = fmaf(pref[r], lo, v[2 * c]);
                v[2 * c + 1] = fmaf(pref[r], hi, v[2 * c + 1]);
            }
        }
#pragma unroll
        for (int c = 0; c < MT; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < MT; ++c) red[warp][c] = v[c];
        }
        __syncthreads();
        if (threadIdx.x < MT) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < KW_THREADS / 32; ++w) s += red[w][threadIdx.x];
            outp[(int64_t)blockIdx.x * MT + threadIdx.x] = ep.nconst * s;
        }
    }
}

// out[g, c] = scale * sum_{s < nsplit} sum_{r in [g*group, (g+1)*group)} part[s][r][c]
// (fixed summation order: deterministic).  part is [nsplit][rows][MT].
__global__ void kw_finalize_kernel(const float* __restrict__ part, int nsplit, int64_t rows, int MT,
                                   int64_t group, float scale, float* __restrict__ out, int64_t ldo, int m,
                                   int64_t ngroups) {
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

template <int D, int MT, int RI, int PM>
static int kw_launch(dim3 grid, cudaStream_t st, const float* X, int64_t ldx, int d, const float* dim_scale,
                     const float* Yp, const float* Wp, float* out, int64_t ldo, int64_t split_stride, int64_t N,
                     int64_t M, int m, int tiles_per_split, const float* row_pref, const EpiParams& ep,
                     int epi_mode, int x_vec) {
    auto kern = kmatw_kernel<D, MT, RI, PM>;
    const size_t smem = KwSmem<D, MT, RI>::kBytes;
    if (smem > 48 * 1024) JRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, KW_THREADS, smem, st>>>(X, ldx, d, dim_scale, Yp, Wp, out, ldo, split_stride, N, M, m,
                                        tiles_per_split, row_pref, ep, epi_mode, x_vec);
    JRK_LAUNCHED();
    return JRK_OK;
}

constexpr int kw_ri(int D, int MT) { return (D + MT <= 32) ? 4 : (D + MT <= 64) ? 2 : 1; }

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
    return total + 1024;
}

// One column block (m <= 32).
static int kmatw_block(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d,
                       int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale,
                       const float* dim_scale, float power, float nconst, const float* row_pref,
                       const float* col_pref, int row_reduce, int64_t group, Scratch& scr, cudaStream_t st) {
    const KwPlan p = kw_plan(N, M, d, m);
    const EpiParams ep = make_epi(scale, power, nconst);
    const int pm = pick_pmode(power);

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
#define KW_GO(D_, MT_, PM_)                                                                                  \
    rc = kw_launch<D_, MT_, kw_ri(D_, MT_), PM_>(grid, st, X, ldx, d, dim_scale, Yp, Wp, kout, kld, kstride, \
                                                 N, M, m, p.tiles_per_split, row_pref, ep, epi_mode, x_vec)
#define KW_MT(D_, PM_)                                                \
    do {                                                              \
        if (p.MT == 4) KW_GO(D_, 4, PM_);                             \
        else if (p.MT == 16) KW_GO(D_, 16, PM_);                      \
        else KW_GO(D_, 32, PM_);                                      \
    } while (0)
#define KW_D(PM_)                                                     \
    do {                                                              \
        switch (p.D) {                                                \
            case 4: KW_MT(4, PM_); break;                             \
            case 8: KW_MT(8, PM_); break;                             \
            case 16: KW_MT(16, PM_); break;                           \
            case 32: KW_MT(32, PM_); break;                           \
            case 64: KW_MT(64, PM_); break;                           \
            default: KW_MT(128, PM_); break;                          \
        }                                                             \
    } while (0)
    int rc = JRK_OK;
    if (pm == PM_SQEXP) KW_D(PM_SQEXP);
    else if (pm == PM_LAPLACE) KW_D(PM_LAPLACE);
    else if (pm == PM_GENERAL) KW_D(PM_GENERAL);
    else KW_D(PM_EXTRA);
#undef KW_D
#undef KW_MT
#undef KW_GO
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
                                                                          fscale, out, ldo, m, ngroups);
        JRK_LAUNCHED();
    }
    return JRK_OK;
}

// The dispatch rule, in one place (also behind jrk_kmatw_uses_tensor_cores): JRK_KMATW_TC=0 forces the FP32-pipe kernel.
bool kmatw_uses_tc(int64_t N, int64_t M, int d, int m, bool has_dim_scale, int row_reduce) {
    const char* e = getenv("JRK_KMATW_TC");
    if (e && e[0] == '0') return false;
    if (g_extra_kind.kind != 0) return false;   // sibling kernels: FP32-pipe kernel only
    static const float dummy = 1.f;
    return kmatw_tc_eligible(N, M, d, m, has_dim_scale ? &dummy : nullptr, row_reduce, 1.f);
}

int kmatw(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m,
          int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale, const float* dim_scale, float power,
          float nconst, const float* row_pref, const float* col_pref, int row_reduce, int64_t group,
          void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (d > 128) return fail(JRK_ERR_UNSUPPORTED, "kmatw: d = %d > 128 not supported by the fused kernel", d);
    // both contractions on the tensor cores when the problem is large, narrow and plain (kmatw_tc.cu)
    if (kmatw_uses_tc(N, M, d, m, dim_scale != nullptr, row_reduce))
        return kmatw_tc(X, Y, W, out, N, M, d, m, ldx, ldy, ldw, ldo, scale, power, nconst, row_pref, col_pref,
                        row_reduce, workspace, workspace_bytes, st);
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(kmatw_workspace_bytes(N, M, d, m, row_reduce, group)));
    for (int c0 = 0; c0 < m; c0 += 32) {  // wide W: column blocks of 32 (K is recomputed per block)
        const int mb = (m - c0 < 32) ? m - c0 : 32;
        scr.used = 0;
        int rc = kmatw_block(X, Y, W + c0, out + c0, N, M, d, mb, ldx, ldy, ldw, ldo, scale, dim_scale, power,
                             nconst, row_pref, col_pref, row_reduce, group, scr, st);
        if (rc) return rc;
    }
    return JRK_OK;
}

}  // namespace jrk

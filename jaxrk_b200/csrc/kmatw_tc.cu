// Kernel (b'): fused, never-materialised K(X, Y) @ W with BOTH contractions on the tensor cores (tcgen05),
// for d <= 32 and m <= 16 on large problems (row reduce none / Sum / Mean).
//
// Replaces the same reference code as kmatw.cu -- FiniteVec.inner (jaxrk/rkhs/vector.py:62-75) with a
// LinearReduce / Sum / Mean on the Y side (jaxrk/reduce/lincomb.py:137-140, reduce/base.py:132-150), the inner()
// of FiniteOp.__matmul__ (jaxrk/rkhs/operator.py:46) -- and computes the same numbers; what changes is where the
// arithmetic runs.  kmatw.cu is bound by the FP32 pipe (2d + m + 2 lane-operations per pair).  Here
//   GEMM1   S = x.y    3 products of fp16 hi / lo operands (kind::f16, K = 16), X block resident in TMEM (A operand),
//                      64-row Y tiles in shared memory (B)
//   epilogue           e = exp2(...) split into tf32 hi / lo, written back over S in TMEM
//   GEMM2   O += e.W   3xTF32, A = e from TMEM, B = W^T tiles in shared memory, O accumulates in TMEM
// so a pair costs ~5 issued instructions and one MUFU instead of ~33 instructions.
//
// A pre-pass (tc_rownorm_kernel x 2, tc_pack_kernel) writes, per 64-row tile of Y, one contiguous image
// [Y hi, Y lo (fp16) | per 32-column k-chunk: W^T hi, W^T lo (tf32) | 64 scaled |y|^2] in the K-major 128B-swizzled
// layout of the UMMA descriptor, so a tile arrives with one cp.async.bulk.
// Persistent CTAs, one per SM, one 128-row block of X per work unit; 19 warps: warp 0 TMA producer, warp 1 GEMM1 issuer
// (+ TMEM allocation), warp 2 GEMM2 issuer, warps 3-18 epilogue (TMEM lane = row i; warp = lane quarter x 16-column
// group).  TMEM columns: X [0, 64) = hi [32) lo [32) | TC_SS = 3 in-place S / e stages [64 + s*128, +128) =
// {c1, c2} -> {e hi, e lo} | two O windows [448 + b*32, +32) = {main [16), cross [16)}.
//
// Two epilogues (see TC_FAST_BOUND): the FACTORED p = 2 path (scale folded into the Y image, 2^(c|y|^2) into W,
// 2^(c|x|^2) applied once per row, one accumulator, e = ex2(S)) and the GENERIC one (any p, any norms: u = a S + xs +
// ys from two accumulators, near-coincident pairs -- sq < (|x|^2+|y|^2)/64 -- recomputed by FP32 direct differencing
// as in the Gram kernel).  Both loops are unrolled by the stage period so that barrier / TMEM addresses are immediates.
//
// The tensor core's accumulation is not round-to-nearest: for all-positive sums the error grows with the number of
// MMAs that add into one accumulator (measured: 4.8x the tolerance with 64-tile windows).  So (i) O accumulates in
// TMEM across a window of TC_WIN tiles only (measured, all-positive W, M = 2^20: 0.50 of the tolerance with 2 tiles,
// 0.53 with 4, 0.60 with 8; 48.7 / 46.2 / 45.4 ms: TC_WIN = 4); the epilogue then folds the window into FP32 registers
// (round to nearest) while the next window accumulates in the second O buffer, and (ii) the large term hi_e.hi_w and
// the small cross terms go to separate accumulators; GEMM2 pairs hi_e with [W^T hi ; W^T lo] in ONE N = 32 MMA
// (an MMA of N <= 32 costs ~20 clk whatever N is) plus lo_e x W^T hi.
// What was measured on the way (ten variants that did not help, and the ones that did) is in profiles/README.md.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "gram_tf32_kernel.cuh"
#include "kmatw_tc2.cuh"

namespace jrk {

using namespace tf32;

namespace {

constexpr int TC_BM = 128, TC_NT = 64, TC_DP = 32, TC_MT = 16;
constexpr int TC_SS = 3;                                 // S / e stages in TMEM
constexpr int TC_WH = 2 * TC_MT * 128;                   // bytes of one W^T image: 2 k-chunks x 16 rows x 128 B
constexpr int TC_STAGE = 25600;                          // stage stride (1024-byte aligned, fits the wide image)
constexpr int TC_SOP = 6;
constexpr int TC_EW = 16;
constexpr int TC_THREADS = (3 + TC_EW) * 32;
constexpr int TC_WIN = 4;                                // tiles per O window (128 j); a power of two
// p = 2 "factored" fast path: taken when |c| (max|x|^2 + max|y|^2) <= TC_FAST_BOUND (kmatw_tc2.cuh; decided on the
// device from the pre-pass maxima).  There K_ij = 2^(c|x_i|^2) . 2^(-2c x_i.y_j) . 2^(c|y_j|^2) with every factor inside
// [2^-12, 2^12]: the column factor is folded into W and the scale -2c into Y by the pack pre-pass, the row factor is
// applied once per output row, and the epilogue is reduced to  e = ex2(S)  (no norms, no near-zero test, one
// accumulator).  Worst-case relative error of a term: ~4e-7 * bound (3xTF32 + truncating accumulation), < 1e-5.
// By default that regime is served by the second-generation kernel (kmatw_tc2.cu) and this file's kernels return at
// once when they find themselves in it (allow_fast == 2); JRK_TC_V2=0 keeps it here (allow_fast == 1) for A/B runs.
constexpr int TC_COL_X = 0, TC_COL_S = 64, TC_COL_O = 64 + TC_SS * 128;   // TMEM columns
constexpr int TC_OFF_BAR = TC_SOP * TC_STAGE;
constexpr int TC_NBAR = 2 * TC_SOP + 3 * TC_SS + 4 + 1;  // op full/empty, s_full/e_ready/s_empty, o_full/o_empty x2, x_ready
constexpr int TC_SMEM = TC_OFF_BAR + TC_NBAR * 8 + 16 + 1024;
static_assert(TC_SMEM <= 227 * 1024, "shared memory budget");

// Byte layout of one tile image.  packed (d <= 16): a Y row is [16 hi | 16 lo] floats in one 128-byte swizzle row.
struct TcImage {
    int y_lo;      // offset of the lo half relative to the hi half (what the GEMM1 descriptors differ by)
    int y_lo_h16;  // the same for the fp16 layout of the factored path (hi and lo halves are 32 / 64 bytes per row)
    int w;         // offset of the W^T image: per 32-column k-chunk 16 rows hi then 16 rows lo (2 x TC_WH bytes in all)
    int yn;        // offset of the 64 scaled squared norms
    int bytes;     // image size (multiple of 16)
    int ypieces;   // 16-byte pieces of Y per row handled by the pack kernel (4 or 8)
};
inline TcImage tc_image(int d) {
    TcImage L;
    const bool packed = d <= 16;
    const int ybytes = packed ? TC_NT * 128 : 2 * TC_NT * 128;
    L.y_lo = packed ? 64 : TC_NT * 128;
    L.y_lo_h16 = packed ? 32 : TC_NT * 128;
    L.w = ybytes;
    L.yn = ybytes + 2 * TC_WH;
    L.bytes = L.yn + TC_NT * 4;
    L.ypieces = packed ? 4 : 8;
    return L;
}

// mbarrier wait / arrive on a precomputed 32-bit shared address (the generic -> shared conversion costs an S2R)
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity), "r"(20000u)
        : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// tf32 split of a positive kernel value for GEMM2: hi by truncation (one LOP), lo = e - hi >= 0 exact
__device__ __forceinline__ void split_trunc(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

struct TcParams {
    const float* X; int64_t ldx; int d;
    const float* Y; int64_t ldy;
    const float* xn;                       // scaled row norms of X
    const unsigned char* img;              // tile images
    TcImage L;
    const float* row_pref;
    float* out; int64_t ldo; int m;
    int64_t N, M;
    int n_tiles, n_units, ksteps;
    float a;                               // PM_SQEXP: -2c; else -2
    int check;                             // 1: near-zero refinement always; 0: only if the norm maxima demand it
    const unsigned* nmax;                  // [2] bit patterns of max |mult| |x|^2, max |mult| |y|^2
    int allow_fast;                        // 0: generic epilogue always (JRK_TC_FAST=0, tests); 1: factored p = 2 path here;
                                           // 2: the factored regime belongs to kmatw_tc2.cu -- return at once inside it
    float inv_abs_mult;                    // 1 / |mult|: nmax * inv_abs_mult = max |row|^2 (fp16 normalisation of GEMM1)
};

// fp16 hi / lo split WITHOUT rescaling lo (factored path, GEMM1): both parts go into ONE accumulator.  Values are
// normalised so that |x| <= 16; a lo part below fp16's normal range costs at most 3e-8 absolute per element, which
// the TC_FAST_BOUND regime turns into < 1.5e-6 in the exponent (d <= 32).
__device__ __forceinline__ void split_h16u_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
    hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

__device__ __forceinline__ bool tc_fast_regime(const unsigned* nmax) { return tc2_fast_regime(nmax); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

template <int PM>
__global__ void __launch_bounds__(TC_THREADS, 1) kmatw_tc_kernel(TcParams P, EpiParams ep) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
    uint64_t* op_full = bars;                    // [TC_SOP]
    uint64_t* op_empty = op_full + TC_SOP;       // [TC_SOP]
    uint64_t* s_full = op_empty + TC_SOP;        // [TC_SS]  GEMM1 of the stage complete
    uint64_t* e_ready = s_full + TC_SS;          // [TC_SS]  epilogue wrote e hi / lo
    uint64_t* s_empty = e_ready + TC_SS;         // [TC_SS]  GEMM2 of the stage complete
    uint64_t* o_full = s_empty + TC_SS;          // [2]  window complete in O buffer
    uint64_t* o_empty = o_full + 2;              // [2]  epilogue folded the O buffer
    uint64_t* x_ready = o_empty + 2;             // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TC_OFF_BAR + TC_NBAR * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // factored p = 2 path?  (warp-uniform, the same decision in every CTA and in the pack pre-pass)
    const bool fast = (PM == PM_SQEXP) && P.allow_fast && tc_fast_regime(P.nmax);
    if (fast && P.allow_fast == 2) return;      // kmatw_tc2_kernel serves this regime
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_SOP; ++s) { mbar_init(&op_full[s], 1); mbar_init(&op_empty[s], 1); }
        for (int s = 0; s < TC_SS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&e_ready[s], TC_EW); mbar_init(&s_empty[s], 1); }
        // every epilogue warp folds its own 4 + 4 columns of an O window
        for (int s = 0; s < 2; ++s) { mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], TC_EW); }
        mbar_init(x_ready, TC_EW);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();

    const int T = P.n_tiles;
    if (warp == 0) {
        // ===================== TMA producer: one image per tile =====================
        uint32_t tc = 0;
        const uint32_t bytes = (uint32_t)P.L.bytes;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            for (int t = 0; t < T; ++t, ++tc) {
                const int s = tc % TC_SOP;
                mbar_wait_hint(&op_empty[s], ((tc / TC_SOP) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&op_full[s], bytes);
                    bulk_g2s(smem + s * TC_STAGE, P.img + (int64_t)t * bytes, bytes, &op_full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== GEMM1 issuer:  S[stage] = x . y  (3xTF32) =====================
        // A single warp issuing both GEMMs was the bottleneck of the first versions (it was ~90 % busy executing its
        // own ~250 instructions per tile while everything else waited), so each GEMM has its own issuing warp and
        // all stage / phase bookkeeping is incremental.  This warp runs up to TC_SS tiles ahead of the epilogue.
        constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_NT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint64_t dbase = make_b_desc(smem_u32(smem));
        const uint32_t b_opfull = smem_u32(op_full), b_sempty = smem_u32(s_empty), b_sfull = smem_u32(s_full), b_xready = smem_u32(x_ready);
        const int ksteps = P.ksteps;
        const uint64_t y_lo = (uint64_t)(P.L.y_lo >> 4);
        uint32_t s = 0, s_ph = 0, a = 0, a_ph = 0, uc = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, ++uc) {
            bar_wait(b_xready, uc & 1);
            for (int t = 0; t < T; ++t) {
                bar_wait(b_opfull + s * 8, s_ph);
                bar_wait(b_sempty + a * 8, a_ph ^ 1);
                tc_fence_after();
                const uint64_t dyh = dbase + (uint64_t)((s * TC_STAGE) >> 4), dyl = dyh + y_lo;
                const uint32_t c1 = TC_COL_S + a * 128, c2 = c1 + 64;
                if (elect_one()) {
                    // fp16 hi / lo operands (K = 16 per MMA: half the MMAs of the tf32 form)
                    constexpr uint32_t idesc1h = (1u << 4) | ((uint32_t)(TC_NT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
                    const int ksh = (P.d + 15) >> 4;
                    const uint64_t dylh = dyh + (uint64_t)(P.L.y_lo_h16 >> 4);
                    if (fast) {
                        // one accumulator: the small cross terms first, the hi.hi terms last (the accumulator
                        // truncates, so only the last additions carry a rounding error that matters)
                        for (int ks = 0; ks < ksh; ++ks) {
                            const uint64_t off = (uint64_t)(ks * 2);
                            mma_f16_ts(c1, TC_COL_X + ks * 8, dylh + off, idesc1h, ks);            // hi_x . lo_y
                            mma_f16_ts(c1, TC_COL_X + TC_DP + ks * 8, dyh + off, idesc1h, 1u);    // lo_x . hi_y
                        }
                        for (int ks = 0; ks < ksh; ++ks)
                            mma_f16_ts(c1, TC_COL_X + ks * 8, dyh + (uint64_t)(ks * 2), idesc1h, 1u);   // hi_x . hi_y
                    } else {
                        // generic epilogue: the same fp16 operands with lo' = lo 2^11 (full precision at any
                        // magnitude), cross terms in their own accumulator, folded in with 2^-11 by the epilogue
                        for (int ks = 0; ks < ksh; ++ks) {
                            const uint64_t off = (uint64_t)(ks * 2);
                            mma_f16_ts(c1, TC_COL_X + ks * 8, dyh + off, idesc1h, ks);            // hi_x . hi_y
                            mma_f16_ts(c2, TC_COL_X + ks * 8, dylh + off, idesc1h, ks);           // hi_x . lo'_y
                            mma_f16_ts(c2, TC_COL_X + TC_DP + ks * 8, dyh + off, idesc1h, 1u);    // lo'_x . hi_y
                        }
                    }
                    tc_commit_addr(b_sfull + a * 8);
                }
                __syncwarp();
                if (++s == TC_SOP) { s = 0; s_ph ^= 1; }
                if (++a == TC_SS) { a = 0; a_ph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===================== GEMM2 issuer:  O[window] += e . W  (3xTF32) =====================
        constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_MT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        constexpr uint32_t idesc2w = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * TC_MT) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint64_t dbase = make_b_desc(smem_u32(smem));
        const uint32_t b_opfull = smem_u32(op_full), b_opempty = smem_u32(op_empty), b_sempty = smem_u32(s_empty),
                       b_eready = smem_u32(e_ready), b_ofull = smem_u32(o_full), b_oempty = smem_u32(o_empty);
        const uint64_t w_off = (uint64_t)(P.L.w >> 4);
        uint32_t s = 0, s_ph = 0, a = 0, a_ph = 0, ob = 0, ob_ph = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            for (int t = 0; t < T; ++t) {
                const bool win_first = (t & (TC_WIN - 1)) == 0, win_last = ((t & (TC_WIN - 1)) == TC_WIN - 1) || (t == T - 1);
                if (win_first) bar_wait(b_oempty + ob * 8, ob_ph ^ 1);
                bar_wait(b_opfull + s * 8, s_ph);
                bar_wait(b_eready + a * 8, a_ph);
                tc_fence_after();
                const uint64_t dwh = dbase + (uint64_t)((s * TC_STAGE) >> 4) + w_off;
                const uint32_t eh = TC_COL_S + a * 128, el = eh + 64, oc = TC_COL_O + ob * 2 * TC_MT, ox = oc + TC_MT;
                if (elect_one()) {
                    // an MMA of N <= 32 costs ~20 clk whatever N is (profiles/microbench/mma_rate.cu), so hi_e meets
                    // [W^T hi ; W^T lo] in ONE MMA of N = 32 whose 32 result columns are {main | cross}
#pragma unroll
                    for (int k = 0; k < TC_NT / 8; ++k) {
                        const uint64_t off = (uint64_t)(((k >> 2) * 2 * TC_MT * 128 + (k & 3) * 32) >> 4);
                        const uint32_t acc0 = (win_first && k == 0) ? 0u : 1u;
                        mma_tf32_ts(oc, eh + k * 8, dwh + off, idesc2w, acc0);  // hi_e . [hi_w | lo_w] -> {main, cross}
                        mma_tf32_ts(ox, el + k * 8, dwh + off, idesc2, 1u);     // lo_e . hi_w          -> cross
                    }
                    tc_commit_addr(b_sempty + a * 8);      // S / e stage reusable
                    tc_commit_addr(b_opempty + s * 8);     // operand stage reusable
                    if (win_last) tc_commit_addr(b_ofull + ob * 8);
                }
                __syncwarp();
                if (win_last) { ob ^= 1; if (ob == 0) ob_ph ^= 1; }
                if (++s == TC_SOP) { s = 0; s_ph ^= 1; }
                if (++a == TC_SS) { a = 0; a_ph ^= 1; }
            }
        }
    } else if (fast) {
        // ===================== epilogue of the factored p = 2 path (and X loader) =====================
        // S already holds -2c x_i.y_j (the scale rides in the Y image), so a tile is: LDTM, 16 MUFU.EX2, tf32
        // split, 2 STTM.  Every warp folds its own 4 + 4 columns of the O window (balanced), and the row factor
        // nconst . pref_i . 2^(c|x_i|^2) is applied once per output row.
        const int e = warp - 3;
        const int quad = warp & 3;
        const int cg = e >> 2;
        uint32_t lane_base;      // TMEM address of this warp's lane quarter, in an opaque register
        asm volatile("mov.u32 %0, %1;" : "=r"(lane_base) : "r"(tmem_base + ((uint32_t)(quad * 32) << 16)));
        uint32_t sbase;
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
        const uint32_t b_sfull0 = sbase + TC_OFF_BAR + 2 * TC_SOP * 8, b_ofull = b_sfull0 + 3 * TC_SS * 8, b_oempty = b_ofull + 16;
        const uint32_t s_col0 = lane_base + TC_COL_S + cg * 16, o_col0 = lane_base + TC_COL_O + cg * 4;
        uint32_t wc = 0;
        uint32_t sa = 0, sa_ph = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            const int64_t i = (int64_t)u * TC_BM + quad * 32 + lane;
            const bool iok = i < P.N;
            const float* xrow = P.X + (iok ? i : 0) * P.ldx;
            // park the X block as fp16 hi / lo, normalised by 2^-ex (the Y image carries a 2^ex): 16 values -> 8 packed
            // columns per column group
            if (cg * 16 < P.d) {
                const float sx = pow2i(-h16_exp(__uint_as_float(__ldg(P.nmax)) * P.inv_abs_mult));
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int kk = cg * 16 + 2 * k;
                    const float v0 = (iok && kk < P.d) ? __ldg(xrow + kk) : 0.f;
                    const float v1 = (iok && kk + 1 < P.d) ? __ldg(xrow + kk + 1) : 0.f;
                    split_h16u_pair(v0 * sx, v1 * sx, hi[k], lo[k]);
                }
                tmem_st8(lane_base + TC_COL_X + cg * 8, hi);
                tmem_st8(lane_base + TC_COL_X + TC_DP + cg * 8, lo);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_ready);
            __syncwarp();
            float oacc[4] = {0.f, 0.f, 0.f, 0.f};

            auto fold_window = [&]() {   // O buffers of window wc, this warp's 4 columns -> registers
                const uint32_t ob = wc & 1;
                bar_wait(b_ofull + ob * 8, (wc >> 1) & 1);
                tc_fence_after();
                float o[4], ox[4];
                tmem_ld4(o_col0 + ob * 2 * TC_MT, o);
                tmem_ld4(o_col0 + ob * 2 * TC_MT + TC_MT, ox);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(b_oempty + ob * 8);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) oacc[c] += o[c] + ox[c];
                ++wc;
            };

            // one tile: S -> e for this warp's 32 rows x 16 columns, on S / e stage `stage` (barriers at bsf / ber)
            auto convert_tile = [&](uint32_t bsf, uint32_t ber, uint32_t sc, uint32_t parity) {
                bar_wait(bsf, parity);
                tc_fence_after();
                float sv[16];
                tmem_ld16(sc, sv);
                tmem_ld_wait();
                uint32_t ehi[16], elo[16];
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const float e0 = ex2_approx(sv[c]), e1 = ex2_approx(sv[c + 1]);
                    ehi[c] = __float_as_uint(e0) & 0xFFFFE000u;          // hi by truncation, lo = e - hi >= 0 exact
                    ehi[c + 1] = __float_as_uint(e1) & 0xFFFFE000u;
                    float l0, l1;
                    unpack2(sub2(pack2(e0, e1), pack2(__uint_as_float(ehi[c]), __uint_as_float(ehi[c + 1]))), l0, l1);
                    elo[c] = __float_as_uint(l0);
                    elo[c + 1] = __float_as_uint(l1);
                }
                tmem_st16(sc, ehi);
                tmem_st16(sc + 64, elo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(ber);
                __syncwarp();
            };
            auto tile_dyn = [&](int t) {       // stage index in a register
                convert_tile(b_sfull0 + sa * 8, b_sfull0 + TC_SS * 8 + sa * 8, s_col0 + sa * 128, sa_ph);
                if (++sa == TC_SS) { sa = 0; sa_ph ^= 1; }
                // fold the previous window once its last GEMM2 has been issued (one tile later)
                if (t > 0 && (t & (TC_WIN - 1)) == 0) fold_window();
            };
            // The stage rotates with period TC_SS: once it is 0, TC_SS tiles are processed per iteration with the stage
            // as a compile-time constant (barrier and TMEM addresses become immediates: ~8 instructions less per tile).
            int t = 0;
            for (; t < T && sa != 0; ++t) tile_dyn(t);
            for (; t + TC_SS <= T; t += TC_SS) {
#pragma unroll
                for (int k = 0; k < TC_SS; ++k) {
                    convert_tile(b_sfull0 + k * 8, b_sfull0 + (TC_SS + k) * 8, s_col0 + k * 128, sa_ph);
                    if (t + k > 0 && ((t + k) & (TC_WIN - 1)) == 0) fold_window();
                }
                sa_ph ^= 1;
            }
            for (; t < T; ++t) tile_dyn(t);
            fold_window();   // the unit's last (possibly partial) window
            if (iok) {
                const float f = ep.nconst * ex2_approx(__ldg(P.xn + i)) * (P.row_pref ? __ldg(P.row_pref + i) : 1.f);
                float* dst = P.out + i * P.ldo + cg * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (cg * 4 + c < P.m) dst[c] = f * oacc[c];
            }
        }
    } else {
        // ===================== epilogue (and X loader) =====================
        const int e = warp - 3;
        const int quad = warp & 3;                          // TMEM lane quarter this warp may touch
        const int cg = e >> 2;                              // 16-column group of the tile
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        // x.y = 2^(ex+ey) (c1 + 2^-11 c2): the power of two rides in `a`
        const float a_eff = P.a * pow2i(h16_exp(__uint_as_float(__ldg(P.nmax)) * P.inv_abs_mult) +
                                        h16_exp(__uint_as_float(__ldg(P.nmax + 1)) * P.inv_abs_mult));
        const f32x2 a2 = pack2(a_eff, a_eff), ntau2 = pack2(-TF_TAU, -TF_TAU), low2 = pack2(1.f / 2048.f, 1.f / 2048.f);
        // shared-window base in an opaque register (otherwise the address conversion is rematerialised per tile)
        uint32_t sbase;
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
        const uint32_t b_opfull = sbase + TC_OFF_BAR, b_sfull = b_opfull + 2 * TC_SOP * 8, b_eready = b_sfull + TC_SS * 8,
                       b_ofull = b_eready + 2 * TC_SS * 8, b_oempty = b_ofull + 16;
        const uint32_t yn_base = sbase + P.L.yn + cg * 64;
        // near-zero test needed?  (p = 2: only when the norms are large against the length scale)
        bool check = true;
        if (PM == PM_SQEXP && !P.check)
            check = (__uint_as_float(__ldg(P.nmax)) + __uint_as_float(__ldg(P.nmax + 1))) > TF_SKIP_BOUND;
        uint32_t wc = 0;
        uint32_t st = 0, st_ph = 0;      // operand stage of the current tile and its phase
        uint32_t sa = 0, sa_ph = 0;      // S / e stage of the current tile and its phase
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            const int64_t i = (int64_t)u * TC_BM + quad * 32 + lane;
            const bool iok = i < P.N;
            const float* xrow = P.X + (iok ? i : 0) * P.ldx;
            // ---- park the X block in TMEM as fp16 hi / lo' (normalised by 2^-ex, lo' = lo 2^11): column groups 0 and 1
            // take 16 values = 8 packed columns each ----
            // (every GEMM1 that read the previous block has completed: this warp waited for the last s_full)
            if (cg * 16 < P.d) {
                const float sx = pow2i(-h16_exp(__uint_as_float(__ldg(P.nmax)) * P.inv_abs_mult));
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int kk = cg * 16 + 2 * k;
                    const float v0 = (iok && kk < P.d) ? __ldg(xrow + kk) : 0.f;
                    const float v1 = (iok && kk + 1 < P.d) ? __ldg(xrow + kk + 1) : 0.f;
                    split_h16_pair(v0 * sx, v1 * sx, hi[k], lo[k]);
                }
                tmem_st8(tmem_base + lane_addr + TC_COL_X + cg * 8, hi);
                tmem_st8(tmem_base + lane_addr + TC_COL_X + TC_DP + cg * 8, lo);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_ready);
            __syncwarp();
            const float xs = iok ? __ldg(P.xn + i) : 0.f;
            const f32x2 xs2 = pack2(xs, xs);
            float oacc[4] = {0.f, 0.f, 0.f, 0.f};

            auto fold_window = [&]() {   // O buffers of window wc, this warp's 4 + 4 columns -> registers (balanced)
                const uint32_t ob = wc & 1;
                bar_wait(b_ofull + ob * 8, (wc >> 1) & 1);
                tc_fence_after();
                float o[4], ox[4];
                tmem_ld4(tmem_base + lane_addr + TC_COL_O + ob * 2 * TC_MT + cg * 4, o);
                tmem_ld4(tmem_base + lane_addr + TC_COL_O + ob * 2 * TC_MT + TC_MT + cg * 4, ox);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(b_oempty + ob * 8);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) oacc[c] += o[c] + ox[c];
                ++wc;
            };

            // one tile: S -> e for this warp's 32 rows x 16 columns
            auto tile = [&](int t, auto chk, uint32_t st, uint32_t st_ph, uint32_t sa, uint32_t sa_ph) {
                constexpr bool CHECK = decltype(chk)::value;
                // the tile's (scaled) |y_j|^2 arrived with the operand image: observe the TMA completion, then
                // broadcast 16-byte reads from shared memory (the stage is released only after GEMM2 of this tile)
                bar_wait(b_opfull + st * 8, st_ph);
                float4 yv[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) yv[g] = lds_v4(yn_base + st * TC_STAGE + g * 16);
                bar_wait(b_sfull + sa * 8, sa_ph);
                tc_fence_after();
                const uint32_t sc = tmem_base + lane_addr + TC_COL_S + sa * 128 + cg * 16;
                float c1[16], c2[16];
                tmem_ld16(sc, c1);
                tmem_ld16(sc + 64, c2);
                tmem_ld_wait();
                float uu[16];
                float wext = 0.f;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const f32x2 t01 = add2(pack2(yv[g].x, yv[g].y), xs2), t23 = add2(pack2(yv[g].z, yv[g].w), xs2);
                    const f32x2 d01 = fma2(pack2(c2[4 * g], c2[4 * g + 1]), low2, pack2(c1[4 * g], c1[4 * g + 1]));
                    const f32x2 d23 = fma2(pack2(c2[4 * g + 2], c2[4 * g + 3]), low2, pack2(c1[4 * g + 2], c1[4 * g + 3]));
                    const f32x2 u01 = fma2(d01, a2, t01), u23 = fma2(d23, a2, t23);
                    unpack2(u01, uu[4 * g], uu[4 * g + 1]);
                    unpack2(u23, uu[4 * g + 2], uu[4 * g + 3]);
                    if (CHECK) {
                        float w0, w1, w2, w3;
                        unpack2(fma2(t01, ntau2, u01), w0, w1);
                        unpack2(fma2(t23, ntau2, u23), w2, w3);
                        if (PM == PM_SQEXP) wext = fmaxf(fmaxf(wext, w0), fmaxf(fmaxf(w1, w2), w3));
                        else wext = fminf(fminf(wext, w0), fminf(fminf(w1, w2), w3));
                    }
                }
                if (CHECK) {
                    const bool refine = (PM == PM_SQEXP) ? (wext > 0.f) : (wext < 0.f);
                    if (__any_sync(0xffffffffu, refine && iok)) {
                        // near-coincident pairs: recompute by FP32 direct differencing (rare path)
                        const float ys_arr[16] = {yv[0].x, yv[0].y, yv[0].z, yv[0].w, yv[1].x, yv[1].y, yv[1].z, yv[1].w,
                                                  yv[2].x, yv[2].y, yv[2].z, yv[2].w, yv[3].x, yv[3].y, yv[3].z, yv[3].w};
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float tt = xs + ys_arr[c];
                            const float w = fmaf(tt, -TF_TAU, uu[c]);
                            const bool bad = (PM == PM_SQEXP) ? (w > 0.f) : (w < 0.f);
                            const int64_t j = (int64_t)t * TC_NT + cg * 16 + c;
                            if (iok && bad && j < P.M) {
                                const float* yr = P.Y + j * P.ldy;
                                float acc2 = 0.f;
#pragma unroll 1
                                for (int k = 0; k < P.d; ++k) {
                                    const float df = __ldg(xrow + k) - __ldg(yr + k);
                                    acc2 = fmaf(df, df, acc2);
                                }
                                uu[c] = (PM == PM_SQEXP) ? acc2 * ep.c : acc2;
                            }
                        }
                    }
                }
                // e = unnormalised kernel value (> 0); split into tf32 hi / lo and write back over S
                uint32_t ehi[16], elo[16];
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const float e0 = (PM == PM_SQEXP) ? ex2_approx(uu[c]) : kexp_unnorm<PM>(uu[c], ep);
                    const float e1 = (PM == PM_SQEXP) ? ex2_approx(uu[c + 1]) : kexp_unnorm<PM>(uu[c + 1], ep);
                    ehi[c] = __float_as_uint(e0) & 0xFFFFE000u;          // hi by truncation, lo = e - hi >= 0 exact
                    ehi[c + 1] = __float_as_uint(e1) & 0xFFFFE000u;
                    float l0, l1;
                    unpack2(sub2(pack2(e0, e1), pack2(__uint_as_float(ehi[c]), __uint_as_float(ehi[c + 1]))), l0, l1);
                    elo[c] = __float_as_uint(l0);
                    elo[c + 1] = __float_as_uint(l1);
                }
                tmem_st16(sc, ehi);
                tmem_st16(sc + 64, elo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(b_eready + sa * 8);
                __syncwarp();
            };
            auto tile_dyn = [&](int t) {       // stage indices in registers
                if (check) tile(t, std::true_type{}, st, st_ph, sa, sa_ph);
                else tile(t, std::false_type{}, st, st_ph, sa, sa_ph);
                if (++st == TC_SOP) { st = 0; st_ph ^= 1; }
                if (++sa == TC_SS) { sa = 0; sa_ph ^= 1; }
                // fold the previous window once its last GEMM2 has been issued (one tile later)
                if (t > 0 && (t & (TC_WIN - 1)) == 0) fold_window();
            };
            // as in the factored path: once the operand stage is 0 (then the S / e stage is 0 too: TC_SOP is a multiple
            // of TC_SS), TC_SOP tiles per iteration with both stages as compile-time constants
            static_assert(TC_SOP % TC_SS == 0, "stage periods");
            int t = 0;
            for (; t < T && st != 0; ++t) tile_dyn(t);
            for (; t + TC_SOP <= T; t += TC_SOP) {
#pragma unroll
                for (int k = 0; k < TC_SOP; ++k) {
                    const uint32_t sap = sa_ph ^ (uint32_t)((k / TC_SS) & 1);
                    if (check) tile(t + k, std::true_type{}, (uint32_t)k, st_ph, (uint32_t)(k % TC_SS), sap);
                    else tile(t + k, std::false_type{}, (uint32_t)k, st_ph, (uint32_t)(k % TC_SS), sap);
                    if (t + k > 0 && ((t + k) & (TC_WIN - 1)) == 0) fold_window();
                }
                st_ph ^= 1;
                if ((TC_SOP / TC_SS) & 1) sa_ph ^= 1;
            }
            for (; t < T; ++t) tile_dyn(t);
            fold_window();   // the unit's last (possibly partial) window
            if (iok) {
                const float f = ep.nconst * (P.row_pref ? __ldg(P.row_pref + i) : 1.f);
                float* dst = P.out + i * P.ldo + cg * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (cg * 4 + c < P.m) dst[c] = f * oacc[c];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc_512(tmem_base);
}

// ---- pre-pass: per-tile images [Y hi, Y lo | W^T hi | W^T lo | yn], one thread per 16-byte piece ----
// Factored p = 2 path (fast_ok and the norm maxima inside TC_FAST_BOUND): Y is stored as a.y (a = -2c) and W_j as
// W_j . 2^(c|y_j|^2) (yn_all: the scaled norms from tc_rownorm_kernel), see TC_FAST_BOUND.
__global__ void tc_pack_kernel(const float* __restrict__ Y, int64_t M, int d, int64_t ldy, const float* __restrict__ W,
                               int m, int64_t ldw, const float* __restrict__ col_pref, int n_tiles, float mult,
                               TcImage L, unsigned char* __restrict__ img, int fast_ok, float a,
                               const float* __restrict__ yn_all, const unsigned* __restrict__ nmax, float inv_abs_mult) {
    const bool fast = fast_ok && tc_fast_regime(nmax);
    if (fast && fast_ok == 2) return;           // kmatw_tc2_pack_kernel writes the image of this regime
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ypt = TC_NT * L.ypieces;            // Y pieces per tile (256 or 512: whole warps)
    const int per_tile = ypt + 256;
    if (idx >= (int64_t)n_tiles * per_tile) return;
    const int64_t t = idx / per_tile;
    const int p = (int)(idx % per_tile);
    unsigned char* base = img + t * L.bytes;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    int off_hi, off_lo;
    if (p < ypt) {   // Y: row r of the tile, 16-byte piece c (columns 4c .. 4c+3)
        const int r = p / L.ypieces, c = p % L.ypieces;
        const int64_t j = t * TC_NT + r;
        if (j < M) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * c + q < d) v[q] = __ldg(Y + j * ldy + 4 * c + q);
        }
        off_hi = r * 128 + ((c ^ (r & 7)) << 4);
        // lo half: a separate 8 KB image, or pieces 4..7 of the same swizzle row (offset 64 B before swizzling)
        off_lo = (L.ypieces == 8) ? L.y_lo + off_hi : r * 128 + (((c + 4) ^ (r & 7)) << 4);
        // the row's scaled squared norm rides along with the tile: xor-shuffle over the row's 4 or 8 lanes
        float q = v[0] * v[0];
        q = fmaf(v[1], v[1], q);
        q = fmaf(v[2], v[2], q);
        q = fmaf(v[3], v[3], q);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        if (L.ypieces == 8) q += __shfl_xor_sync(0xffffffffu, q, 4);
        if (c == 0) reinterpret_cast<float*>(base + L.yn)[r] = mult * q;
        {
            // fp16 hi / lo image: 4 values = 8 bytes per thread, hi at byte 2k of the row, lo 32 bytes further (packed
            // layout) or in the lo image.  Factored path: y'' = a y 2^ex with lo unscaled (ex normalises X to |x| <= 1,
            // so S = x~ . y'' is the exponent itself).  Generic path: y 2^-ey with lo' = lo 2^11.
            uint2 h2, l2;
            if (fast) {
                const float f = a * pow2i(h16_exp(__uint_as_float(__ldg(nmax)) * inv_abs_mult));
                split_h16u_pair(v[0] * f, v[1] * f, h2.x, l2.x);
                split_h16u_pair(v[2] * f, v[3] * f, h2.y, l2.y);
            } else {
                const float f = pow2i(-h16_exp(__uint_as_float(__ldg(nmax + 1)) * inv_abs_mult));
                split_h16_pair(v[0] * f, v[1] * f, h2.x, l2.x);
                split_h16_pair(v[2] * f, v[3] * f, h2.y, l2.y);
            }
            const int oh = r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8;
            const int ol = (L.ypieces == 8) ? L.y_lo_h16 + oh : r * 128 + ((((c >> 1) + 2) ^ (r & 7)) << 4) + (c & 1) * 8;
            *reinterpret_cast<uint2*>(base + oh) = h2;
            *reinterpret_cast<uint2*>(base + ol) = l2;
            return;
        }
    } else {         // W^T: row = output column cc, K = j; piece holds 4 consecutive j
        const int pw = p - ypt;
        const int kc = pw >> 7, cc = (pw >> 3) & 15, c = pw & 7;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t j = t * TC_NT + kc * 32 + 4 * c + q;
            if (j < M && cc < m) {
                v[q] = __ldg(W + j * ldw + cc) * (col_pref ? __ldg(col_pref + j) : 1.f);
                if (fast) v[q] *= ex2_approx(__ldg(yn_all + j));
            }
        }
        // per k-chunk: 16 rows of W^T hi, then 16 rows of W^T lo (together the N = 32 operand of GEMM2)
        off_hi = L.w + kc * (2 * TC_MT * 128) + cc * 128 + ((c ^ (cc & 7)) << 4);
        off_lo = off_hi + TC_MT * 128;
    }
    uint4 h, l;
    split_tf32(v[0], h.x, l.x);
    split_tf32(v[1], h.y, l.y);
    split_tf32(v[2], h.z, l.z);
    split_tf32(v[3], h.w, l.w);
    *reinterpret_cast<uint4*>(base + off_hi) = h;
    *reinterpret_cast<uint4*>(base + off_lo) = l;
}

// out[r] = mult * |row r|^2 for r < rows, 0 up to rows_padded; *gmax = max |mult| |row|^2 (atomicMax on the bits)
__global__ void tc_rownorm_kernel(const float* __restrict__ A, int64_t rows, int64_t rows_padded, int d, int64_t lda,
                                  float mult, float* __restrict__ out, unsigned* __restrict__ gmax) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float q = 0.f;
    if (r < rows) {
        for (int k = 0; k < d; ++k) {
            const float v = __ldg(A + r * lda + k);
            q = fmaf(v, v, q);
        }
    }
    if (r < rows_padded) out[r] = mult * q;
    float mx = fabsf(mult) * q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx == mx) atomicMax(gmax, __float_as_uint(mx));
}

// Sum / Mean over rows of T (N x 16, dense), deterministic: per-chunk partial sums, then one thread per column.
__global__ void tc_rowsum_kernel(const float* __restrict__ T, int64_t N, int64_t chunk, float* __restrict__ part) {
    __shared__ float red[16][TC_MT];
    const int c = threadIdx.x & 15, rl = threadIdx.x >> 4;          // 16 row lanes x 16 columns
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = (r0 + chunk < N) ? r0 + chunk : N;
    float a = 0.f;
    for (int64_t r = r0 + rl; r < r1; r += 16) a += T[r * TC_MT + c];
    red[rl][c] = a;
    __syncthreads();
    if (rl == 0) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) s += red[q][c];
        part[(int64_t)blockIdx.x * TC_MT + c] = s;
    }
}
// only_if_fast (nullable): write only inside the factored regime (the d <= 64 path of kmatw.cu, whose FP32-pipe fallback
// writes the reduced result itself outside it)
__global__ void tc_rowsum_final_kernel(const float* __restrict__ part, int64_t nchunks, float scale, float* __restrict__ out, int m,
                                       const unsigned* __restrict__ only_if_fast = nullptr) {
    if (only_if_fast && !tc2_fast_regime(only_if_fast)) return;
    const int c = threadIdx.x;
    if (c >= m) return;
    float s = 0.f;
    for (int64_t q = 0; q < nchunks; ++q) s += part[q * TC_MT + c];
    out[c] = scale * s;
}

template <int PM>
int tc_launch(const TcParams& P, const EpiParams& ep, cudaStream_t st) {
    auto kern = kmatw_tc_kernel<PM>;
    JRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    const int grid = P.n_units < tf_sm_count() ? P.n_units : tf_sm_count();
    kern<<<grid, TC_THREADS, TC_SMEM, st>>>(P, ep);
    JRK_LAUNCHED();
    return JRK_OK;
}

}  // namespace

bool kmatw_tc_eligible(int64_t N, int64_t M, int d, int m, const float* dim_scale, int row_reduce, float nconst) {
    (void)nconst;
    // Sum / Mean over rows: the N x m result goes to scratch and is reduced in a fixed order (tc_rowsum_kernel)
    if (dim_scale || !(row_reduce == JRK_REDUCE_NONE || row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN)) return false;
    // wide W runs in column blocks of 16: one MUFU-bound pass per block, still ~2.5x the FP32-pipe kernel's 32 columns
    if (d < 1 || d > TC_DP || m < 1) return false;
    // enough row blocks to fill the machine and enough columns to amortise the pre-pass
    return N >= (int64_t)TC_BM * 64 && M >= 4096 && M <= ((int64_t)1 << 30) && N <= ((int64_t)1 << 36);
}

constexpr int64_t TC_RS_CHUNK = 1024;    // rows per partial sum of the Sum / Mean reduction

// Sum (scale = 1) / Mean (scale = 1 / N) over the rows of T (N x 16, dense) into out[0 .. m); rpart: ceil(N / 1024) x 16 floats
size_t kmatw_tc_rowsum_bytes(int64_t N) { return align256((size_t)((N + TC_RS_CHUNK - 1) / TC_RS_CHUNK) * TC_MT * 4); }
int kmatw_tc_rowsum(const float* T, int64_t N, int m, float scale, float* out, float* rpart, const unsigned* only_if_fast,
                    cudaStream_t st) {
    const int64_t nchunks = (N + TC_RS_CHUNK - 1) / TC_RS_CHUNK;
    tc_rowsum_kernel<<<(unsigned)nchunks, 256, 0, st>>>(T, N, TC_RS_CHUNK, rpart);
    JRK_LAUNCHED();
    tc_rowsum_final_kernel<<<1, 32, 0, st>>>(rpart, nchunks, scale, out, m, only_if_fast);
    JRK_LAUNCHED();
    return JRK_OK;
}

// ---- the Y / W side (what a plan keeps: jrk_kmatw_plan_*) and the X side (what an apply needs) ----
size_t kmatw_tc_prep_bytes(int64_t M, int d) {
    const int64_t n_tiles = (M + TC_NT - 1) / TC_NT;
    return align256((size_t)n_tiles * tc_image(d).bytes) + align256((size_t)M * 4) + 256;
}
size_t kmatw_tc_apply_bytes(int64_t N) {
    return align256((size_t)N * TC_MT * 4) + align256((size_t)((N + TC_RS_CHUNK - 1) / TC_RS_CHUNK) * TC_MT * 4) +
           align256((size_t)N * 4) + 256;
}
size_t kmatw_tc_workspace_bytes(int64_t N, int64_t M) { return kmatw_tc_prep_bytes(M, TC_DP) + kmatw_tc_apply_bytes(N) + 1024; }

struct TcPrep {            // device pointers into the prep buffer
    unsigned char* img;    // tile images: second generation (factored regime) or first generation, decided on the device
    float* yn_all;         // mult |y_j|^2
    unsigned* ymax;        // [2] bit patterns: max |mult| |y|^2, max |W col_pref|
};
static TcPrep tc_prep_layout(void* buf, int64_t M, int d) {
    const int64_t n_tiles = (M + TC_NT - 1) / TC_NT;
    char* b = (char*)buf;
    TcPrep p;
    p.img = (unsigned char*)b;
    b += align256((size_t)n_tiles * tc_image(d).bytes);
    p.yn_all = (float*)b;
    b += align256((size_t)M * 4);
    p.ymax = (unsigned*)b;
    return p;
}

static int tc_allow_fast(int pm) {
    int allow_fast = (pm == PM_SQEXP) && jrk_option(JRK_OPT_TC_FAST) != 0;
    if (allow_fast && jrk_option(JRK_OPT_TC_V2) != 0) allow_fast = 2;
    return allow_fast;
}

// Y-side pre-pass: row norms of Y, their maximum, max |W col_pref| and the tile images.
//   which & 1: first-generation image (generic epilogue; and the factored regime when JRK_OPT_TC_V2 = 0) into pr.img
//   which & 2: second-generation image (factored regime) into pr.img
// Each pack kernel decides on the device, from nmax4, whether the image is its to write; with both bits set exactly one
// of them writes.  nmax_x == nullptr (plan creation): the x maximum counts as zero, i.e. the second-generation image is
// written whenever Y ALONE is inside the regime.
static int tc_prepare(const float* Y, const float* W, int64_t M, int d, int m, int64_t ldy, int64_t ldw, const float* col_pref,
                      const EpiParams& ep, int pm, const TcPrep& pr, const unsigned* nmax_x, unsigned* nmax4, int which,
                      cudaStream_t st) {
    const int64_t n_tiles = (M + TC_NT - 1) / TC_NT;
    const TcImage L = tc_image(d);
    const float mult = (pm == PM_SQEXP) ? ep.c : 1.f;
    const float a_scale = (pm == PM_SQEXP) ? -2.f * ep.c : -2.f;
    const int allow_fast = tc_allow_fast(pm);
    // nmax4 = {max |c||x|^2 (copied or zero), max |c||y|^2, max |W col_pref|, -}
    JRK_CUDA_OK(cudaMemsetAsync(nmax4, 0, 4 * sizeof(unsigned), st));
    if (nmax_x) JRK_CUDA_OK(cudaMemcpyAsync(nmax4, nmax_x, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    tc_rownorm_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(Y, M, M, d, ldy, mult, pr.yn_all, nmax4 + 1);
    JRK_LAUNCHED();
    if (which & 1) {
        const int64_t tot = n_tiles * (TC_NT * L.ypieces + 256);
        tc_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, W, m, ldw, col_pref, (int)n_tiles, mult, L, pr.img,
                                                                      allow_fast, a_scale, pr.yn_all, nmax4, 1.f / fabsf(mult));
        JRK_LAUNCHED();
    }
    if ((which & 2) && allow_fast == 2) {
        if (int rc = kmatw_tc2_pack(Y, W, M, d, m, ldy, ldw, col_pref, ep.c, pr.yn_all, nmax4, pr.img, st)) return rc;
    } else {
        // max |W col_pref| is part of the prepared state either way
        if (int rc = kmatw_tc2_wmax(W, M, m, ldw, col_pref, nmax4, st)) return rc;
    }
    JRK_CUDA_OK(cudaMemcpyAsync(pr.ymax, nmax4 + 1, 2 * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    return JRK_OK;
}

// X-side: both generations of the main kernel (each returns at once outside its regime), then the optional Sum / Mean
// over rows.  nmax4 = {max |c||x|^2, max |c||y|^2, max |W col_pref|}; img2 / img1: second- / first-generation images.
static int tc_apply(const float* X, const float* Y, float* out, int64_t N, int64_t M, int d, int m, int64_t ldx, int64_t ldy,
                    int64_t ldo, const EpiParams& ep, int pm, float nconst, const float* row_pref, int row_reduce,
                    const unsigned char* img2, const unsigned char* img1, const float* xn, const unsigned* nmax4, float* Tbuf,
                    float* rpart, cudaStream_t st) {
    const int64_t n_tiles = (M + TC_NT - 1) / TC_NT;
    const TcImage L = tc_image(d);
    const float mult = (pm == PM_SQEXP) ? ep.c : 1.f;
    const bool reduce_rows = row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN;
    const int64_t nchunks = (N + TC_RS_CHUNK - 1) / TC_RS_CHUNK;
    const int allow_fast = tc_allow_fast(pm);
    TcParams P;
    P.X = X; P.ldx = ldx; P.d = d; P.Y = Y; P.ldy = ldy; P.xn = xn; P.img = img1; P.L = L; P.row_pref = row_pref;
    P.out = reduce_rows ? Tbuf : out; P.ldo = reduce_rows ? TC_MT : ldo; P.m = reduce_rows ? TC_MT : m; P.N = N; P.M = M;
    P.n_tiles = (int)n_tiles;
    P.n_units = (int)((N + TC_BM - 1) / TC_BM);
    P.ksteps = (d + 7) / 8;
    P.a = (pm == PM_SQEXP) ? -2.f * ep.c : -2.f;
    P.allow_fast = allow_fast;
    P.inv_abs_mult = 1.f / fabsf(mult);
    // p = 2: the near-zero test only matters when the norms are large against the length scale; the kernel decides
    // from the device-side maxima of the pre-pass (no host synchronisation)
    P.check = 0;
    P.nmax = nmax4;
    if (allow_fast == 2) {
        if (int rc = kmatw_tc2_apply(X, N, d, ldx, xn, img2, M, nmax4, ep.c, nconst, row_pref, P.out, P.ldo, P.m, st)) return rc;
    }
    int rc;
    if (pm == PM_SQEXP) rc = tc_launch<PM_SQEXP>(P, ep, st);
    else if (pm == PM_LAPLACE) rc = tc_launch<PM_LAPLACE>(P, ep, st);
    else rc = tc_launch<PM_GENERAL>(P, ep, st);
    if (rc || !reduce_rows) return rc;
    tc_rowsum_kernel<<<(unsigned)nchunks, 256, 0, st>>>(Tbuf, N, TC_RS_CHUNK, rpart);
    JRK_LAUNCHED();
    tc_rowsum_final_kernel<<<1, 32, 0, st>>>(rpart, nchunks, row_reduce == JRK_REDUCE_MEAN ? 1.f / (float)N : 1.f, out, m);
    JRK_LAUNCHED();
    return JRK_OK;
}

int kmatw_tc(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m, int64_t ldx,
             int64_t ldy, int64_t ldw, int64_t ldo, float scale, float power, float nconst, const float* row_pref,
             const float* col_pref, int row_reduce, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(kmatw_tc_workspace_bytes(N, M)));
    const TcPrep pr = tc_prep_layout(scr.take<char>(kmatw_tc_prep_bytes(M, d)), M, d);
    float* xn = scr.take<float>((size_t)N);
    unsigned* nmax_x = scr.take<unsigned>(4);
    unsigned* nmax4 = scr.take<unsigned>(4);
    const bool reduce_rows = row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN;
    float* Tbuf = reduce_rows ? scr.take<float>((size_t)N * TC_MT) : nullptr;
    float* rpart = reduce_rows ? scr.take<float>((size_t)((N + TC_RS_CHUNK - 1) / TC_RS_CHUNK) * TC_MT) : nullptr;

    const EpiParams ep = make_epi(scale, power, nconst);
    const int pm = pick_pmode(power);
    const float mult = (pm == PM_SQEXP) ? ep.c : 1.f;
    JRK_CUDA_OK(cudaMemsetAsync(nmax_x, 0, 4 * sizeof(unsigned), st));
    tc_rownorm_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(X, N, N, d, ldx, mult, xn, nmax_x);
    JRK_LAUNCHED();
    for (int c0 = 0; c0 < m; c0 += TC_MT) {     // wide W: column blocks of 16 (e is recomputed per block)
        const int mb = (m - c0 < TC_MT) ? m - c0 : TC_MT;
        if (int rc = tc_prepare(Y, W + c0, M, d, mb, ldy, ldw, col_pref, ep, pm, pr, nmax_x, nmax4, 3, st)) return rc;
        if (int rc = tc_apply(X, Y, out + c0, N, M, d, mb, ldx, ldy, ldo, ep, pm, nconst, row_pref, row_reduce, pr.img, pr.img, xn,
                              nmax4, Tbuf, rpart, st))
            return rc;
    }
    return JRK_OK;
}

// ---- plan support (kmatw_plan.cu): the Y / W side is prepared once, every apply only brings X ----
// The plan keeps, per 16-column block of W, the second-generation tile images (packed with the x maximum taken as zero:
// valid whenever the pair (X, Y) is inside the factored regime), the scaled norms of Y and the two maxima.  An apply
// whose X pushes the pair out of that regime -- or a kernel outside it (p != 2, large norms) -- packs first-generation
// images into its own scratch, on the device's own decision: nothing is read back, nothing synchronises.
size_t kmatw_tc_plan_prep_bytes(int64_t M, int d, int m) { return (size_t)((m + TC_MT - 1) / TC_MT) * kmatw_tc_prep_bytes(M, d); }
size_t kmatw_tc_plan_apply_bytes(int64_t N, int64_t M, int d) { return kmatw_tc_apply_bytes(N) + kmatw_tc_prep_bytes(M, d) + 2048; }

int kmatw_tc_plan_prepare(const float* Y, const float* W, int64_t M, int d, int m, int64_t ldy, int64_t ldw,
                          const float* col_pref, float scale, float power, float nconst, void* prep, unsigned* nmax4_scratch,
                          cudaStream_t st) {
    const EpiParams ep = make_epi(scale, power, nconst);
    const int pm = pick_pmode(power);
    const size_t per = kmatw_tc_prep_bytes(M, d);
    int b = 0;
    for (int c0 = 0; c0 < m; c0 += TC_MT, ++b) {
        const int mb = (m - c0 < TC_MT) ? m - c0 : TC_MT;
        const TcPrep pr = tc_prep_layout((char*)prep + (size_t)b * per, M, d);
        if (int rc = tc_prepare(Y, W + c0, M, d, mb, ldy, ldw, col_pref, ep, pm, pr, nullptr, nmax4_scratch, 2, st)) return rc;
    }
    return JRK_OK;
}

// nmax4 = {x maximum, y maximum, W maximum} of one column block
__global__ void tc_plan_nmax_kernel(const unsigned* nmax_x, const unsigned* ymax, unsigned* nmax4) {
    nmax4[0] = nmax_x[0]; nmax4[1] = ymax[0]; nmax4[2] = ymax[1]; nmax4[3] = 0u;
}

int kmatw_tc_plan_apply(const float* X, const float* Y, const float* W, const float* col_pref, const void* prep, float* out,
                        int64_t N, int64_t M, int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo, float scale,
                        float power, float nconst, const float* row_pref, int row_reduce, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(kmatw_tc_plan_apply_bytes(N, M, d)));
    const TcPrep own = tc_prep_layout(scr.take<char>(kmatw_tc_prep_bytes(M, d)), M, d);   // first-generation image, if needed
    float* xn = scr.take<float>((size_t)N);
    unsigned* nmax_x = scr.take<unsigned>(4);
    unsigned* nmax4 = scr.take<unsigned>(4);
    const bool reduce_rows = row_reduce == JRK_REDUCE_SUM || row_reduce == JRK_REDUCE_MEAN;
    float* Tbuf = reduce_rows ? scr.take<float>((size_t)N * TC_MT) : nullptr;
    float* rpart = reduce_rows ? scr.take<float>((size_t)((N + TC_RS_CHUNK - 1) / TC_RS_CHUNK) * TC_MT) : nullptr;
    const EpiParams ep = make_epi(scale, power, nconst);
    const int pm = pick_pmode(power);
    const float mult = (pm == PM_SQEXP) ? ep.c : 1.f;
    const float a_scale = (pm == PM_SQEXP) ? -2.f * ep.c : -2.f;
    const int allow_fast = tc_allow_fast(pm);
    const int64_t n_tiles = (M + TC_NT - 1) / TC_NT;
    const TcImage L = tc_image(d);
    JRK_CUDA_OK(cudaMemsetAsync(nmax_x, 0, 4 * sizeof(unsigned), st));
    tc_rownorm_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(X, N, N, d, ldx, mult, xn, nmax_x);
    JRK_LAUNCHED();
    const size_t per = kmatw_tc_prep_bytes(M, d);
    int b = 0;
    for (int c0 = 0; c0 < m; c0 += TC_MT, ++b) {
        const int mb = (m - c0 < TC_MT) ? m - c0 : TC_MT;
        const TcPrep pr = tc_prep_layout((char*)prep + (size_t)b * per, M, d);
        tc_plan_nmax_kernel<<<1, 1, 0, st>>>(nmax_x, pr.ymax, nmax4);
        JRK_LAUNCHED();
        // first-generation image into the apply's scratch: the pack kernel returns at once when the second-generation
        // kernel serves the regime (allow_fast == 2)
        const int64_t tot = n_tiles * (TC_NT * L.ypieces + 256);
        tc_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, W + c0, mb, ldw, col_pref, (int)n_tiles, mult, L,
                                                                      own.img, allow_fast, a_scale, pr.yn_all, nmax4,
                                                                      1.f / fabsf(mult));
        JRK_LAUNCHED();
        if (int rc = tc_apply(X, Y, out + c0, N, M, d, mb, ldx, ldy, ldo, ep, pm, nconst, row_pref, row_reduce, pr.img, own.img, xn,
                              nmax4, Tbuf, rpart, st))
            return rc;
    }
    return JRK_OK;
}

}  // namespace jrk

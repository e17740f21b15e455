// Kernel (b''): second generation of the fused, never-materialised K(X, Y) @ W on the tensor cores -- the FACTORED
// p = 2 regime of kmatw_tc.cu (|c| (max|x|^2 + max|y|^2) <= 12, decided on the device), which is where the squared-
// exponential kernel lives for any sensible length scale, and BASELINE.json configs[2] in particular.
//
// Replaces the same reference code as kmatw.cu / kmatw_tc.cu: FiniteVec.inner (jaxrk/rkhs/vector.py:62-75) with a
// LinearReduce / Sum / Mean on the Y side (jaxrk/reduce/lincomb.py:137-140, reduce/base.py:132-150), i.e. the inner() of
// FiniteOp.__matmul__ (jaxrk/rkhs/operator.py:46), for GenGaussKernel with power = 2 (jaxrk/kern/rbf.py:84-105).
//
//   K_ij = nconst 2^(c|x_i|^2) 2^(-2c x_i.y_j) 2^(c|y_j|^2),  c = -s^2 log2(e) < 0,  every factor inside [2^-12, 2^12]
//   GEMM1   S = x~ . y''      x~ = x 2^-r (parked in TMEM, fp16 hi / lo),  y'' = -2c y 2^r (tile image, fp16 hi / lo),
//                              2^-r ~ sqrt(-2c) chosen on the host, so both operands stay below 7 in magnitude and the
//                              image does not depend on X (a plan can keep it: jrk_kmatw_plan_*)
//   epilogue e = ex2(S)        ONE MUFU per pair -- the roofline of this kernel -- split into a 16-bit hi / lo pair and
//                              written back over S in TMEM, packed: 64 FP32 columns become 32 + 32 columns of pairs
//   GEMM2   O += e . W'        kind::f16 (K = 16 per MMA: 12 MMAs per tile instead of 16 tf32 ones), W' = W col_pref
//                              2^(c|y|^2) 2^-ew in the tile image as hi / lo
//   out_i = nconst 2^(c|x_i|^2) row_pref_i 2^ew . O_i
//
// What changed against the first generation, and why (profiles/README.md, round 2):
//  * The 16 epilogue warps no longer share every tile (one 16-column group each): warp k of a TMEM lane quarter owns
//    the WHOLE quarter (32 rows x 64 columns) of the tiles t = k (mod 4).  The four warps of an SM sub-partition then
//    work on four different tiles, out of phase, and one of them always has MUFU work ready; before, all four were
//    released by the same barrier, ran their MUFU bursts together and then idled together on TMEM stores and barriers
//    (the kernel sat at ~700 clk per 128 x 64 tile against 512 clk of MUFU).
//  * e is handed to GEMM2 as fp16 operands: e_hi = the leading 11 bits (by truncation, exact), e_lo' = (e - e_hi) 2^11,
//    against W'_hi and W'_lo' = (W' - W'_hi) 2^11; the cross-term accumulator is folded in with 2^-11.  Half the TMEM
//    write-back, 8 instead of 16 GEMM2 MMAs per tile, and an S / e stage is 64 columns, so SIX stages fit.  (fp16 hi
//    with a bf16 lo would not need the scaling, but a kind::f16 MMA with different A and B formats is an illegal
//    instruction on sm_100a: profiles/microbench/umma_probe.cu.)
//  * The tile image shrinks from 16.6 KB to 8 KB for d <= 16 (Y hi | Y lo | W' pieces share the 128-byte swizzle rows)
//    and 12 KB for d <= 32: half the L2 -> SM traffic of the first generation.
//
// TMEM columns: X hi [0, 16) lo [16, 32) | three O windows [32 + 32 b, +32) = {main [16) | cross [16)} | six S / e
// stages [128 + 64 a, +64).  The tensor core's accumulator truncates, so O is accumulated in TMEM over windows of four
// tiles only and folded into FP32 registers (round to nearest) two windows later; the hi.hi products and the small
// cross terms use separate accumulators (as in kmatw_tc.cu).
#include <cmath>
#include <cstdlib>

#include "gram_tf32_kernel.cuh"
#include "kmatw_tc2.cuh"
#include "tc2_prims.cuh"

namespace jrk {

using namespace tf32;
using namespace tcp;

namespace {

constexpr int T2_BM = 128, T2_NT = 64, T2_MT = 16;
constexpr int T2_NS = 6;                      // S / e stages in TMEM (64 columns each)
constexpr int T2_NOB = 3;                     // O windows in TMEM
constexpr int T2_WIN = 4;                     // tiles per O window
constexpr int T2_LAGT = 2;                    // a window is folded once a tile T2_LAGT beyond its last one has been converted
constexpr int T2_SOP = 12;                    // operand (tile image) stages in shared memory
constexpr int T2_STAGE = 12288;               // stage stride: the 12 KB image of 16 < d <= 32
constexpr int T2_EPQ_MAX = 4;
constexpr int T2_COL_X = 0, T2_COL_XLO = 16, T2_COL_O = 32, T2_COL_S = 128;
// What depends on the number of GEMM1 k-steps KS (16 features each): KS = 1 (d <= 16, 8 KB packed image), 2 (d <= 32,
// 12 KB), 3 / 4 (d <= 48 / 64, 20 KB: [Y hi 64 x 128 B][Y lo 64 x 128 B][W' 32 x 128 B]).  These park 32 + 32 columns of x~, so the
// O windows and the S / e stages move up and one stage goes: x~ [0, 64) | O [64, 160) | five stages [160, 480).
template <int KS>
struct T2V {
    static constexpr int COL_XLO = KS >= 3 ? 32 : T2_COL_XLO;
    static constexpr int COL_O = KS >= 3 ? 64 : T2_COL_O;
    static constexpr int COL_S = KS >= 3 ? 160 : T2_COL_S;
    static constexpr int NS = KS >= 3 ? 5 : T2_NS;
    static constexpr int SOP = KS >= 3 ? 7 : T2_SOP;
    static constexpr int STAGE = KS >= 3 ? 20480 : T2_STAGE;
    static constexpr int YLO = KS == 1 ? 2 : (KS == 2 ? 4 : 512);            // Y lo against Y hi, in 16-byte units
    static constexpr int WOFF = KS >= 3 ? 1024 : 512;                         // W' region (KS >= 2), in 16-byte units
    static_assert(COL_S + NS * 64 <= 512 && COL_O + T2_NOB * 32 <= COL_S && SOP * STAGE <= T2_SOP * T2_STAGE, "budget");
};
constexpr int T2_OFF_RED = T2_SOP * T2_STAGE;                            // per-warp partial results of a unit (2 buffers)
constexpr int T2_RED_WARP = T2_MT * 32 * 4;                              // [16 columns][32 lanes] floats
constexpr int T2_OFF_BAR = T2_OFF_RED + 2 * 4 * T2_EPQ_MAX * T2_RED_WARP;
constexpr int T2_NBAR = 2 * T2_SOP + 3 * T2_NS + 2 * T2_NOB + 1 + 4 + 1 + 4 * T2_EPQ_MAX;
constexpr int T2_SMEM = T2_OFF_BAR + T2_NBAR * 8 + 16 + 1024;
static_assert(T2_COL_S + T2_NS * 64 <= 512 && T2_COL_O + T2_NOB * 32 <= T2_COL_S, "TMEM budget");
static_assert(T2_SMEM <= 227 * 1024, "shared memory budget");

// hi / lo' of a pair of W' values: hi = fp16(w), lo' = fp16((w - hi) 2^11)
__device__ __forceinline__ void split_w_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    lo = cvt_f16x2((x1 - __half2float(h1)) * 2048.f, (x0 - __half2float(h0)) * 2048.f);
}
// 16 exponents -> 8 columns of e_hi pairs + 8 columns of e_lo' pairs:  e = ex2(s), e_hi = leading 11 bits (exact in
// fp16), e_lo' = (e - e_hi) 2^11 >= 0
__device__ __forceinline__ void exp_split16(const uint32_t (&sv)[16], uint32_t (&eo)[16]) {
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
        const float e0 = ex2_approx(__uint_as_float(sv[c])), e1 = ex2_approx(__uint_as_float(sv[c + 1]));
        const float h0 = __uint_as_float(__float_as_uint(e0) & 0xFFFFE000u);
        const float h1 = __uint_as_float(__float_as_uint(e1) & 0xFFFFE000u);
        float l0, l1;
        unpack2(mul2(sub2(pack2(e0, e1), pack2(h0, h1)), pack2(2048.f, 2048.f)), l0, l1);
        eo[c >> 1] = cvt_f16x2(h1, h0);
        eo[8 + (c >> 1)] = cvt_f16x2(l1, l0);
    }
}

struct Tc2Params {
    const float* X; int64_t ldx; int d;
    const float* xn;                 // c |x_i|^2
    const unsigned char* img; int img_bytes; int packed;
    const float* row_pref;
    float* out; int64_t ldo; int m;
    int64_t N;
    int n_tiles, n_units;
    const unsigned* nmax;            // [3] bit patterns: max |c||x|^2, max |c||y|^2, max |W col_pref|
    float x_scale;                   // 2^-r
    float nconst;
    int token_rel;                   // stagger token: chunks (1 .. 3) a warp must be into its tile before the next tile of
                                     // the lane quarter may start; 0 = off
    long long* trace;                // TRACE builds only (tools/tc2_trace.py): per-warp clock64 stamps of CTA 0
};
constexpr int T2_TRACE_TILES = 512;  // stamps per warp: [warp][tile][4]

// EPQ: epilogue warps per TMEM lane quarter (= per SM sub-partition);  KS: GEMM1 k-steps, 2 for 16 < d <= 32 (12 KB
// images) -- compile-time, because the two MMA-issuing warps are serial latency chains (one elected thread each): every
// instruction on their per-tile path counts (the first version spent ~430 clk per tile issuing 3 MMAs, and was what
// bounded the kernel: tools/tc2_trace.py)
template <int EPQ, int KS, bool TRACE = false>
__global__ void __launch_bounds__((5 + 4 * EPQ) * 32, 1) kmatw_tc2_kernel(Tc2Params P) {
    if (!tc2_fast_regime(P.nmax)) return;     // the generic kernel (kmatw_tc.cu) serves the problem
    using V = T2V<KS>;
    constexpr int EW = 4 * EPQ;
    extern __shared__ unsigned char t2_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(t2_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_OFF_BAR);
    uint64_t* op_full = bars;                    // [SOP]  tile image landed
    uint64_t* op_empty = op_full + T2_SOP;       // [SOP]  GEMM2 of the tile complete
    uint64_t* s_full = op_empty + T2_SOP;        // [NS]   GEMM1 of the stage complete
    uint64_t* e_ready = s_full + T2_NS;          // [NS]   the four owner warps wrote e
    uint64_t* s_empty = e_ready + T2_NS;         // [NS]   GEMM2 of the stage complete
    uint64_t* o_full = s_empty + T2_NS;          // [NOB]  window complete
    uint64_t* o_empty = o_full + T2_NOB;         // [NOB]  the four owner warps folded the window
    uint64_t* x_ready = o_empty + T2_NOB;        // [1]    the epilogue warps parked the unit's X block
    uint64_t* g1_tok = x_ready + 1;              // [4]    issuing warp j finished the GEMM1 phase of a group
    uint64_t* x_free = g1_tok + 4;               // [1]    every GEMM1 of the unit has completed (X may be replaced)
    uint64_t* half_done = x_free + 1;            // [4][EPQ_MAX]  epilogue warp (quarter, k) is P.token_rel chunks into its tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T2_OFF_BAR + T2_NBAR * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_SOP; ++s) { mbar_init(&op_full[s], 1); mbar_init(&op_empty[s], 1); }
        for (int s = 0; s < T2_NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&e_ready[s], 4); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < T2_NOB; ++s) { mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 4); }
        mbar_init(x_ready, EW);
        for (int s = 0; s < 4; ++s) mbar_init(&g1_tok[s], 1);
        mbar_init(x_free, 4);
        for (int s = 0; s < 4 * T2_EPQ_MAX; ++s) mbar_init(&half_done[s], 1);
        fence_mbar_init();
    }
    // roles: warps 0 .. EW-1 epilogue; then FOUR MMA-issuing warps, one per SM sub-partition, and the TMA producer.
    // Issuing warp j takes the groups of four tiles (= O windows) with global index = j (mod 4): GEMM1 of its four tiles,
    // then GEMM2 of the same four once their e is ready.  A single GEMM2 warp cost the epilogue warps of ITS
    // sub-partition ~60 % of their speed (and GEMM1's ~35 %: tools/tc2_trace.py, profiles/microbench/
    // issue_interference.cu -- tcgen05.mma issue delays the tcgen05.ld / st of its neighbours), and a tile is only done when
    // all four lane quarters are; spread over four warps every sub-partition carries the same 2.75 MMAs per tile.
    // Different groups accumulate into different O windows and different S stages, so no two issuing threads ever touch
    // the same accumulator; the hand-over of stages between them goes through the s_empty / e_ready mbarriers.
    constexpr int W_SVC = EW, W_TMA = EW + 4;
    if (warp == W_SVC) tmem_alloc_512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();

    const int T = P.n_tiles;
    if (warp == W_TMA) {
        // ===================== TMA producer: one image per tile =====================
        uint32_t tc = 0;
        const uint32_t bytes = (uint32_t)P.img_bytes;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            for (int t = 0; t < T; ++t, ++tc) {
                const int s = tc % V::SOP;
                mbar_wait_hint(&op_empty[s], ((tc / V::SOP) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&op_full[s], bytes);
                    bulk_g2s(smem + s * V::STAGE, P.img + (int64_t)t * bytes, bytes, &op_full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= W_SVC) {
        // ===================== MMA issuers =====================
        const uint32_t j = (uint32_t)(warp - W_SVC);
        constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(T2_NT >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);       // N = 64
        constexpr uint32_t id16 = (1u << 4) | ((uint32_t)(T2_MT >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);         // N = 16
        constexpr uint32_t id32 = (1u << 4) | ((uint32_t)((2 * T2_MT) >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);   // N = 32
        constexpr uint64_t YLO = V::YLO;                             // Y lo: 32 / 64 bytes further inside the rows, or its own region
        const uint64_t dbase = make_b_desc(smem_u32(smem));
        const uint32_t b_opfull = smem_u32(op_full), b_opempty = smem_u32(op_empty), b_sfull = smem_u32(s_full),
                       b_eready = smem_u32(e_ready), b_sempty = smem_u32(s_empty), b_ofull = smem_u32(o_full),
                       b_oempty = smem_u32(o_empty), b_xready = smem_u32(x_ready), b_tok = smem_u32(g1_tok),
                       b_xfree = smem_u32(x_free);
        const int ngrp = (T + T2_WIN - 1) / T2_WIN;
        uint32_t gt = 0, gw = 0, uc = 0;           // tiles / windows / units before the current unit
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, ++uc) {
            bool xseen = false;
            const int n_first = (int)((j - gw) & 3u);
            if (n_first >= ngrp && lane == 0) bar_arrive(b_xfree);     // no group of this unit is ours
            __syncwarp();
            for (int n = n_first; n < ngrp; n += 4) {
                if (!xseen) { bar_wait(b_xready, uc & 1); xseen = true; }
                // A parity wait is only meaningful when the waiter is at most ONE phase behind the barrier, and this warp
                // skips three groups out of four: so the GEMM1 phases of consecutive groups are chained (group w starts
                // after group w - 1 has issued its GEMM1s).  By induction every barrier waited on below has then completed
                // the phase before the awaited one (GEMM1 of tile g - 6 / g - 12 has been issued, which waited for it).
                const uint32_t wg = gw + (uint32_t)n;                   // global group = window index
                if (wg > 0) bar_wait(b_tok + ((j - 1u) & 3u) * 8, ((wg - 1u) >> 2) & 1u);
                const int t0 = n * T2_WIN, t1 = (t0 + T2_WIN < T) ? t0 + T2_WIN : T;
                // ---- GEMM1:  S[stage] = x~ . y''  (fp16 hi / lo, one accumulator) ----
                for (int t = t0; t < t1; ++t) {
                    const uint32_t g = gt + (uint32_t)t, s = g % V::SOP, a = g % V::NS;
                    const long long tr0 = TRACE ? clock64() : 0;
                    bar_wait(b_opfull + s * 8, (g / V::SOP) & 1);
                    const long long tr1 = TRACE ? clock64() : 0;
                    bar_wait(b_sempty + a * 8, ((g / V::NS) & 1) ^ 1);
                    tc_fence_after();
                    const long long tr2 = TRACE ? clock64() : 0;
                    const uint64_t dy = dbase + (uint64_t)((s * V::STAGE) >> 4);
                    const uint32_t cs = V::COL_S + a * 64;
                    if (elect_one()) {
                        // the small cross terms first, the hi.hi terms last: the accumulator truncates, so only the last
                        // additions carry a rounding error that matters
                        mma_f16_ts(cs, T2_COL_X, dy + YLO, idesc1, 0u);                        // hi_x . lo_y
                        mma_f16_ts(cs, V::COL_XLO, dy, idesc1, 1u);                            // lo_x . hi_y
#pragma unroll
                        for (int q = 1; q < KS; ++q) {
                            mma_f16_ts(cs, T2_COL_X + 8 * q, dy + YLO + 2 * q, idesc1, 1u);
                            mma_f16_ts(cs, V::COL_XLO + 8 * q, dy + 2 * q, idesc1, 1u);
                        }
#pragma unroll
                        for (int q = 0; q < KS; ++q) mma_f16_ts(cs, T2_COL_X + 8 * q, dy + 2 * q, idesc1, 1u);   // hi_x . hi_y
                        commit_addr(b_sfull + a * 8);
                    }
                    __syncwarp();
                    if (TRACE && blockIdx.x == 0 && uc == 0 && t < T2_TRACE_TILES && lane == 0) {
                        long long* tr = P.trace + ((int64_t)17 * T2_TRACE_TILES + t) * 4;
                        tr[0] = tr0; tr[1] = tr1; tr[2] = tr2; tr[3] = clock64();
                    }
                }
                if (lane == 0) bar_arrive(b_tok + j * 8);
                if (n + 4 >= ngrp && elect_one()) commit_addr(b_xfree);      // our last GEMM1s of the unit: arrives when they are done
                __syncwarp();
                // ---- GEMM2:  O[window] += e . W'  (fp16 hi / lo' operands) ----
                const uint32_t w = wg, ob = w % T2_NOB;
                const uint32_t oc = V::COL_O + ob * 32, ox = oc + T2_MT;
                const long long tq0 = TRACE ? clock64() : 0;
                bar_wait(b_oempty + ob * 8, ((w / T2_NOB) & 1) ^ 1);
                for (int t = t0; t < t1; ++t) {
                    const uint32_t g = gt + (uint32_t)t, s = g % V::SOP, a = g % V::NS;
                    const long long tr1 = TRACE ? clock64() : 0;
                    bar_wait(b_eready + a * 8, (g / V::NS) & 1);      // (this thread has already seen op_full of the tile)
                    tc_fence_after();
                    const long long tr2 = TRACE ? clock64() : 0;
                    const uint64_t dw = dbase + (uint64_t)((s * V::STAGE) >> 4);
                    const uint32_t eh = V::COL_S + a * 64;
                    if (elect_one()) {
                        // an MMA of N <= 32 costs the same whatever N is: e_hi meets [W'_hi ; W'_lo'] (stacked rows) in ONE
                        // MMA whose 32 result columns are {main | cross}
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            // offset (16-byte units) of the W' operand of k-step q in the image; W'_lo' sits 16 rows = 2048 B further
                            const uint64_t dq = dw + (uint64_t)(KS > 1 ? (V::WOFF + 2 * q) : ((q >> 1) * 256 + 4 + 2 * (q & 1)));
                            if (q == 0) mma_f16_ts(oc, eh, dq, id32, t == t0 ? 0u : 1u);    // e_hi . [W'_hi ; W'_lo'] -> {main | cross}
                            else mma_f16_ts(oc, eh + 16 * q, dq, id32, 1u);
                            mma_f16_ts(ox, eh + 16 * q + 8, dq, id16, 1u);                  // e_lo' . W'_hi -> cross
                        }
                        commit_addr(b_sempty + a * 8);      // S / e stage reusable
                        commit_addr(b_opempty + s * 8);     // operand stage reusable
                        if (t == t1 - 1) commit_addr(b_ofull + ob * 8);
                    }
                    __syncwarp();
                    if (TRACE && blockIdx.x == 0 && uc == 0 && t < T2_TRACE_TILES && lane == 0) {
                        long long* tr = P.trace + ((int64_t)18 * T2_TRACE_TILES + t) * 4;
                        tr[0] = (t == t0) ? tq0 : tr1; tr[1] = tr1; tr[2] = tr2; tr[3] = clock64();
                    }
                }
            }
            gt += (uint32_t)T;
            gw += (uint32_t)ngrp;
        }
    } else {
        // ===================== epilogue warps (and X loader) =====================
        // warp k of a lane quarter owns the quarter (32 rows x 64 columns) of the tiles t = k (mod EPQ), and folds the
        // O windows w = k (mod EPQ) (all 16 + 16 columns of its 32 rows) into FP32 registers
        const int quad = warp & 3;                 // TMEM lane quarter this warp may touch
        const int k = warp >> 2;
        uint32_t lane_base;                        // opaque registers: keep address arithmetic out of the loop
        asm volatile("mov.u32 %0, %1;" : "=r"(lane_base) : "r"(tmem_base + ((uint32_t)(quad * 32) << 16)));
        uint32_t sbase;
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
        const uint32_t b_sfull = sbase + T2_OFF_BAR + 2 * T2_SOP * 8, b_eready = b_sfull + T2_NS * 8,
                       b_ofull = b_sfull + 3 * T2_NS * 8, b_oempty = b_ofull + T2_NOB * 8, b_xready = b_oempty + T2_NOB * 8,
                       b_xfree = b_xready + 5 * 8, b_half = b_xfree + 8 + (uint32_t)(quad * T2_EPQ_MAX) * 8;
        // Stagger token (P.token_rel = 1 .. 3, 0 = off): the four warps of a sub-partition share ONE MUFU pipe; left alone
        // they fall into lockstep (all convert, then all store / fold / wait together, and the pipe idles: tools/
        // tc2_trace.py).  With the token a warp starts tile t only after the owner of tile t - 1 (same lane quarter) is
        // token_rel chunks into it, so the MUFU phases interleave and one warp's bookkeeping hides behind another's ex2.
        const int token_rel = P.token_rel;

        const int nwin = (T + T2_WIN - 1) / T2_WIN;
        // 2^ew: the power of two that normalises W' (the pack pre-pass divides by it)
        const float wpow = pow2i(tc2_w_exp(__uint_as_float(__ldg(P.nmax + 2))));
        float* red = reinterpret_cast<float*>(smem + T2_OFF_RED);

        auto park_x = [&](int u) {
            // x~ = x 2^-r as fp16 hi / lo (unscaled lo): 16 values -> 8 packed columns; warps 0 and 1 of the quarter
            const int64_t i = (int64_t)u * T2_BM + quad * 32 + lane;
            if (k < KS) {       // every k-step GEMM1 issues must hold data or zeros: TMEM is not cleared between kernels
                const bool iok = i < P.N;
                const float* xrow = P.X + (iok ? i : 0) * P.ldx;
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int kk = k * 16 + 2 * q;
                    const float v0 = (iok && kk < P.d) ? __ldg(xrow + kk) : 0.f;
                    const float v1 = (iok && kk + 1 < P.d) ? __ldg(xrow + kk + 1) : 0.f;
                    split_h16u_pair(v0 * P.x_scale, v1 * P.x_scale, hi[q], lo[q]);
                }
                tmem_st8(lane_base + T2_COL_X + k * 8, hi);
                tmem_st8(lane_base + V::COL_XLO + k * 8, lo);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(b_xready);
            __syncwarp();       // reconverge: what follows are .sync.aligned tcgen05 instructions
        };

        uint32_t gt = 0, gw = 0, uc = 0;           // tiles / windows / units this CTA has finished
        int u = blockIdx.x;
        if (u < P.n_units) park_x(u);
        for (; u < P.n_units; u += gridDim.x, ++uc) {
            float oacc[T2_MT];
#pragma unroll
            for (int c = 0; c < T2_MT; ++c) oacc[c] = 0.f;
            int next_w = k;                        // next window of this unit that this warp folds
            auto fold = [&]() {
                const uint32_t w = gw + (uint32_t)next_w, ob = w % T2_NOB;
                bar_wait(b_ofull + ob * 8, (w / T2_NOB) & 1);
                tc_fence_after();
                uint32_t o[16], ox[16];
                ld16_async(lane_base + V::COL_O + ob * 32, o);
                ld16_async(lane_base + V::COL_O + ob * 32 + T2_MT, ox);
                ld16_wait(o);
                ld16_wait(ox);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(b_oempty + ob * 8);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < T2_MT; ++c) oacc[c] += fmaf(__uint_as_float(ox[c]), 1.f / 2048.f, __uint_as_float(o[c]));
                next_w += EPQ;
            };
            for (int t = k; t < T; t += EPQ) {
                const uint32_t g = gt + (uint32_t)t, a = g % V::NS;
                const long long tr0 = TRACE ? clock64() : 0;
                if (token_rel && (t > 0 || uc > 0)) {
                    // EVERY tile but the very first waits for its predecessor, across unit boundaries too: a parity wait
                    // is only right when the barrier is 0 or 1 phases from the awaited one, and an unbroken chain is
                    // what guarantees that (a warp that skips a wait can run ahead, its successors pass "early" on the
                    // wrong parity, and the chain dead-locks two tiles later)
                    const int tp = t > 0 ? t - 1 : T - 1;                            // the previous tile ...
                    const uint32_t kq = (uint32_t)(tp % EPQ);                        // ... its owner ...
                    const uint32_t tpu_q = (uint32_t)((T - (int)kq + EPQ - 1) / EPQ);
                    const uint32_t idx = (t > 0 ? uc : uc - 1) * tpu_q + (uint32_t)(tp / EPQ);   // ... and which of its arrivals
                    bar_wait(b_half + kq * 8, idx & 1);
                }
                bar_wait(b_sfull + a * 8, (g / V::NS) & 1);
                tc_fence_after();
                const long long tr1 = TRACE ? clock64() : 0;
                const uint32_t sc = lane_base + V::COL_S + a * 64;
                // four chunks of 16 S columns -> 8 + 8 columns of e pairs, in place; the next chunk's load is in flight
                // while the current one goes through the MUFU
                uint32_t s0[16], s1[16], eo[16];
                ld16_async(sc, s0);
                ld16_wait(s0);
                ld16_async(sc + 16, s1);
                exp_split16(s0, eo);
                tmem_st16(sc, eo);
                if (token_rel == 1 && lane == 0) bar_arrive(b_half + k * 8);
                if (token_rel == 1) __syncwarp();
                ld16_wait(s1);
                ld16_async(sc + 32, s0);
                exp_split16(s1, eo);
                tmem_st16(sc + 16, eo);
                if (token_rel == 2 && lane == 0) bar_arrive(b_half + k * 8);
                if (token_rel == 2) __syncwarp();
                ld16_wait(s0);
                ld16_async(sc + 48, s1);
                exp_split16(s0, eo);
                tmem_st16(sc + 32, eo);
                if (token_rel == 3 && lane == 0) bar_arrive(b_half + k * 8);
                if (token_rel == 3) __syncwarp();
                ld16_wait(s1);
                exp_split16(s1, eo);
                tmem_st16(sc + 48, eo);
                const long long tr2 = TRACE ? clock64() : 0;
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(b_eready + a * 8);
                __syncwarp();   // reconverge before the fold's / the next tile's .sync.aligned tcgen05.ld
                if (TRACE && blockIdx.x == 0 && uc == 0 && t < T2_TRACE_TILES && lane == 0) {
                    long long* tr = P.trace + ((int64_t)warp * T2_TRACE_TILES + t) * 4;
                    tr[0] = tr0; tr[1] = tr1; tr[2] = tr2; tr[3] = clock64();
                }
                if (next_w < nwin && T2_WIN * next_w + (T2_WIN - 1) + T2_LAGT <= t) fold();
            }
            // once every GEMM1 of this unit has completed (x_free: one commit per issuing warp) the next block of X can be
            // parked, so that the next unit's GEMM1s overlap the tail of this one
            const int un = u + gridDim.x;
            if (un < P.n_units) {
                bar_wait(b_xfree, uc & 1);
                tc_fence_after();
                park_x(un);
            }
            while (next_w < nwin) fold();
            // the EPQ partial results of a lane quarter meet in shared memory (two buffers, alternating by unit)
            float* mine = red + ((uc & 1) * EW + quad * EPQ + k) * (T2_MT * 32);
            if (EPQ > 1) {
                if (k != 0) {
#pragma unroll
                    for (int c = 0; c < T2_MT; ++c) mine[c * 32 + lane] = oacc[c];
                }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(32 * EPQ) : "memory");
            }
            if (k == 0) {
#pragma unroll
                for (int kk = 1; kk < EPQ; ++kk) {
#pragma unroll
                    for (int c = 0; c < T2_MT; ++c) oacc[c] += mine[kk * (T2_MT * 32) + c * 32 + lane];
                }
                const int64_t i = (int64_t)u * T2_BM + quad * 32 + lane;
                if (i < P.N) {
                    const float f = P.nconst * wpow * ex2_approx(__ldg(P.xn + i)) * (P.row_pref ? __ldg(P.row_pref + i) : 1.f);
                    float* dst = P.out + i * P.ldo;
#pragma unroll
                    for (int c = 0; c < T2_MT; ++c)
                        if (c < P.m) dst[c] = f * oacc[c];
                }
            }
            gt += (uint32_t)T;
            gw += (uint32_t)nwin;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == W_SVC) tmem_dealloc_512(tmem_base);
}

// ---- pre-pass: per-tile images, one thread per 16-byte piece (8 fp16 / bf16 values) ----
//   packed (d <= 16), 8 KB: row r = [Y hi (32 B) | Y lo (32 B) | W' pieces (64 B)] in the 128B-swizzled K-major layout;
//     the W' operand of k-step q (16 columns x 16 j) sits in rows 32 (q >> 1) + 16 piece + column, bytes 64 + 32 (q & 1)
//   wide (d <= 32), 12 KB: row r = [Y hi (64 B) | Y lo (64 B)], then 32 rows {W' hi [16) ; W' lo [16)} x 64 j
__global__ void tc2_pack_kernel(const float* __restrict__ Y, int64_t M, int d, int64_t ldy, const float* __restrict__ W,
                                int m, int64_t ldw, const float* __restrict__ col_pref, int n_tiles,
                                unsigned char* __restrict__ img, int img_bytes, int packed, float y_scale,
                                const float* __restrict__ yn_all, const unsigned* __restrict__ nmax) {
    if (!tc2_fast_regime(nmax)) return;
    const int cpt = img_bytes >> 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_tiles * cpt) return;
    const int64_t t = idx / cpt;
    const int p = (int)(idx % cpt);
    const int rr = (p >> 3), cph = p & 7;
    const int r = rr & 63;                       // row inside its region (Y regions: 64 rows; W' region: 32 rows)
    const int cl = cph ^ (r & 7);                // logical 16-byte chunk of the row
    const bool ks4 = img_bytes == 20480;         // d <= 64: [Y hi 512 chunks][Y lo 512 chunks][W' 256 chunks]
    const bool wregion = ks4 ? p >= 1024 : p >= 512;
    float v[8];
    bool is_w, is_lo;
    if (!wregion && (packed ? cl < 4 : true)) {
        // ---- Y: 8 consecutive features of row j ----
        is_w = false;
        int k0;
        if (ks4) {
            is_lo = p >= 512;
            k0 = 8 * cl;
        } else {
            const int half = packed ? 2 : 4;
            is_lo = cl >= half;
            k0 = 8 * (cl - (is_lo ? half : 0));
        }
        const int64_t j = t * T2_NT + r;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (j < M && k0 + q < d) ? __ldg(Y + j * ldy + k0 + q) * y_scale : 0.f;
    } else {
        // ---- W': 8 consecutive j of output column c ----
        is_w = true;
        int c;
        int64_t j0;
        if (packed) {
            const int q = 2 * (r >> 5) + ((cl - 4) >> 1);
            is_lo = ((r >> 4) & 1) != 0;
            c = r & 15;
            j0 = t * T2_NT + 16 * q + 8 * ((cl - 4) & 1);
        } else {
            is_lo = (r >> 4) != 0;
            c = r & 15;
            j0 = t * T2_NT + 8 * cl;
        }
        const float wscale = pow2i(-tc2_w_exp(__uint_as_float(__ldg(nmax + 2))));
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t j = j0 + q;
            float w = 0.f;
            if (j < M && c < m) {
                w = __ldg(W + j * ldw + c) * wscale;
                if (col_pref) w *= __ldg(col_pref + j);
                w *= ex2_approx(__ldg(yn_all + j));
            }
            v[q] = w;
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t hi, lo;
        if (is_w) split_w_pair(v[2 * q], v[2 * q + 1], hi, lo);
        else split_h16u_pair(v[2 * q], v[2 * q + 1], hi, lo);
        o[q] = is_lo ? lo : hi;
    }
    *reinterpret_cast<uint4*>(img + t * img_bytes + (int64_t)p * 16) = make_uint4(o[0], o[1], o[2], o[3]);
}

// *gmax = max_j max_c |W[j, c] col_pref[j]|  (atomicMax on the bit pattern; NaN ignored)
__global__ void tc2_wmax_kernel(const float* __restrict__ W, int64_t M, int m, int64_t ldw, const float* __restrict__ col_pref,
                                unsigned* __restrict__ gmax) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float mx = 0.f;
    if (j < M) {
        for (int c = 0; c < m; ++c) mx = fmaxf(mx, fabsf(__ldg(W + j * ldw + c)));
        if (col_pref) mx *= fabsf(__ldg(col_pref + j));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx == mx && mx < 3.0e38f) atomicMax(gmax, __float_as_uint(mx));
}

// out[r] = c |row r|^2;  *gmax = max |c| |row|^2 (atomicMax on the bit pattern)
__global__ void tc2_rownorm_kernel(const float* __restrict__ A, int64_t rows, int d, int64_t lda, float c,
                                   float* __restrict__ out, unsigned* __restrict__ gmax) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float q = 0.f;
    if (r < rows) {
        for (int k = 0; k < d; ++k) {
            const float v = __ldg(A + r * lda + k);
            q = fmaf(v, v, q);
        }
        out[r] = c * q;
    }
    float mx = fabsf(c) * q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx == mx) atomicMax(gmax, __float_as_uint(mx));
}

}  // namespace

// debugging aid of tools/tc2_trace.py (not part of the C ABI in include/jrk.h): a device buffer of
// 19 x T2_TRACE_TILES x 4 int64 that receives per-tile clock64 stamps of every warp of CTA 0
static long long* g_tc2_trace = nullptr;
extern "C" void jrk_debug_tc2_trace(void* device_buffer) { g_tc2_trace = (long long*)device_buffer; }

int tc2_image_bytes(int d) { return d <= 16 ? 8192 : (d <= 32 ? 12288 : 20480); }

// W maximum (nmax[2]) + tile images of the factored regime.  nmax[0..1] (norm maxima) and yn_all must already be
// enqueued on `st`.  Both kernels return at once outside the factored regime.
int kmatw_tc2_wmax(const float* W, int64_t M, int m, int64_t ldw, const float* col_pref, unsigned* nmax, cudaStream_t st) {
    JRK_CUDA_OK(cudaMemsetAsync(nmax + 2, 0, sizeof(unsigned), st));
    tc2_wmax_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(W, M, m, ldw, col_pref, nmax + 2);
    JRK_LAUNCHED();
    return JRK_OK;
}

int kmatw_tc2_pack(const float* Y, const float* W, int64_t M, int d, int m, int64_t ldy, int64_t ldw,
                   const float* col_pref, float c, const float* yn_all, unsigned* nmax, unsigned char* img,
                   cudaStream_t st) {
    const int64_t n_tiles = (M + T2_NT - 1) / T2_NT;
    const int img_bytes = tc2_image_bytes(d);
    if (int rc = kmatw_tc2_wmax(W, M, m, ldw, col_pref, nmax, st)) return rc;
    const Tc2Scales sc = tc2_scales(c);
    const int64_t tot = n_tiles * (img_bytes >> 4);
    tc2_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, W, m, ldw, col_pref, (int)n_tiles, img, img_bytes,
                                                                   d <= 16, sc.y_scale, yn_all, nmax);
    JRK_LAUNCHED();
    return JRK_OK;
}

int kmatw_tc2_apply(const float* X, int64_t N, int d, int64_t ldx, const float* xn, const unsigned char* img, int64_t M,
                    const unsigned* nmax, float c, float nconst, const float* row_pref, float* out, int64_t ldo, int m,
                    cudaStream_t st) {
    Tc2Params P;
    P.X = X; P.ldx = ldx; P.d = d; P.xn = xn; P.img = img; P.img_bytes = tc2_image_bytes(d); P.packed = d <= 16;
    P.row_pref = row_pref; P.out = out; P.ldo = ldo; P.m = m; P.N = N;
    P.n_tiles = (int)((M + T2_NT - 1) / T2_NT);
    P.n_units = (int)((N + T2_BM - 1) / T2_BM);
    P.nmax = nmax;
    P.x_scale = tc2_scales(c).x_scale;
    P.nconst = nconst;
    P.token_rel = jrk_option(JRK_OPT_TC2_TOKEN);
    P.trace = g_tc2_trace;
    const int grid = P.n_units < tf_sm_count() ? P.n_units : tf_sm_count();
    // epilogue warps per SM sub-partition: 4; JRK_TC2_EPQ = 3 for A/B runs
    const int epq = jrk_option(JRK_OPT_TC2_EPQ);
#define T2_LAUNCH(EPQ_, KS_, TRACE_)                                                                                 \
    do {                                                                                                             \
        JRK_CUDA_OK(cudaFuncSetAttribute(kmatw_tc2_kernel<EPQ_, KS_, TRACE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         T2_SMEM));                                                                  \
        kmatw_tc2_kernel<EPQ_, KS_, TRACE_><<<grid, (5 + 4 * EPQ_) * 32, T2_SMEM, st>>>(P);                           \
    } while (0)
    const int ks = d <= 16 ? 1 : (d <= 32 ? 2 : (d <= 48 ? 3 : 4));
    if (g_tc2_trace) {
        if (epq == 4) T2_LAUNCH(4, 1, true);
        else T2_LAUNCH(3, 1, true);
    } else if (ks == 4) {
        T2_LAUNCH(4, 4, false);                 // the x~ block of d <= 64 needs four parking warps per lane quarter
    } else if (ks == 3) {
        T2_LAUNCH(4, 3, false);
    } else if (epq == 3) {
        if (ks == 2) T2_LAUNCH(3, 2, false);
        else T2_LAUNCH(3, 1, false);
    } else {
        if (ks == 2) T2_LAUNCH(4, 2, false);
        else T2_LAUNCH(4, 1, false);
    }
#undef T2_LAUNCH
    JRK_LAUNCHED();
    return JRK_OK;
}

// ---- 32 < d <= 64: a self-contained path (norm pre-pass, images, kernel) for one block of <= 16 columns of W.  Outside the
// factored regime every kernel of it returns at once; *nmax_out is what the caller's FP32-pipe kernel tests to know
// whether the result is its to write (kmatw.cu). ----
size_t kmatw_tc64_workspace_bytes(int64_t N, int64_t M) {
    const int64_t n_tiles = (M + T2_NT - 1) / T2_NT;
    return align256((size_t)N * 4) + align256((size_t)M * 4) + 256 + align256((size_t)n_tiles * 20480) + 1024;
}

// m: columns written to `out` (16 when the caller reduces the full block afterwards); m_w: columns of W (default m).
int kmatw_tc64(const float* X, const float* Y, const float* W, float* out, int64_t N, int64_t M, int d, int m, int64_t ldx,
               int64_t ldy, int64_t ldw, int64_t ldo, float c, float nconst, const float* row_pref, const float* col_pref,
               void* scratch, const unsigned** nmax_out, cudaStream_t st, int m_w) {
    if (m_w < 0) m_w = m;
    char* p = (char*)scratch;
    float* xn = (float*)p; p += align256((size_t)N * 4);
    float* yn = (float*)p; p += align256((size_t)M * 4);
    unsigned* nmax = (unsigned*)p; p += 256;
    unsigned char* img = (unsigned char*)p;
    JRK_CUDA_OK(cudaMemsetAsync(nmax, 0, 16, st));
    tc2_rownorm_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(X, N, d, ldx, c, xn, nmax);
    JRK_LAUNCHED();
    tc2_rownorm_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, c, yn, nmax + 1);
    JRK_LAUNCHED();
    if (int rc = kmatw_tc2_pack(Y, W, M, d, m_w, ldy, ldw, col_pref, c, yn, nmax, img, st)) return rc;
    if (int rc = kmatw_tc2_apply(X, N, d, ldx, xn, img, M, nmax, c, nconst, row_pref, out, ldo, m, st)) return rc;
    *nmax_out = nmax;
    return JRK_OK;
}

}  // namespace jrk

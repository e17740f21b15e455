// Kernel (g'): the K@W vector-Jacobian product with both inner products on the tensor cores -- the factored p = 2
// regime of kmatw_tc2.cu (|c| (max|x|^2 + max|y|^2) <= 12, decided on the device), d <= 16, m <= 16.
//
// Replaces the same reference behaviour as kgrad.cu MODE 0: what jax.grad derives from FiniteVec.inner
// (jaxrk/rkhs/vector.py:62-75) with a LinearReduce on the Y side (jaxrk/reduce/lincomb.py:137-140) for GenGaussKernel
// with power = 2 (jaxrk/kern/rbf.py:84-105) -- gradients with respect to the points and the scale, K never materialised.
//
//   t_ij = a_ij K_ij / nconst,   a_ij = G_i . W_j,   K_ij / nconst = 2^(c|x_i|^2) 2^(S_ij) 2^(c|y_j|^2),  S = -2c x.y
//   gX_i      = fac  (sum_j t_ij y_j - x_i sum_j t_ij)
//   gs_rows_i = fac_s (|x_i|^2 sum_j t_ij - 2 x_i . sum_j t_ij y_j + sum_j t_ij |y_j|^2)
//
// kgrad.cu spends 3d + m + 8 FP32 lane-operations per pair (differences, the m-term dot product a_ij, the d + 2
// accumulations).  Here the two inner products run on tcgen05 into TMEM,
//   GEMM1   S = x~ . y''          kind::f16, fp16 hi / lo operands, exactly as kmatw_tc2.cu
//   GEMM1b  g = G . W'^T          kind::tf32, 3 x TF32 (hi / lo: no range restriction on G or W),  W' = W 2^(c|y|^2)
// and an epilogue thread (one TMEM lane = one row i) is left with  tt = ex2(S) g  and the d + 2 accumulations
// sum_j tt (y_j, 1, |y_j|^2) in FP32 registers: 18 FMA lane-operations per pair instead of 72.  y_j, 1 and |y_j|^2 come
// from the same tile image as FP32, read as shared-memory broadcasts.  Accumulation is hierarchical (per tile of 32
// columns, then per unit), so the rounding error does not grow with M.
//
// Roofline: FP32 pipe, d + 2 = 18 lane-operations per pair -> 3.69e13 / 18 = 2.05e12 pairs/s (measured: 1.16e12, FMA pipe
// 66 % busy; kgrad.cu: 2.3 - 3.2e11).  An epilogue thread owns one row of EACH of the unit's two row blocks, so every
// y_j it reads from shared memory (4 x 128-bit + 1 x 64-bit broadcast loads) serves two pairs: with one row block the
// kernel was bound by the shared-memory load pipe (8.8e11; ncu: mio_throttle the top stall).
//
// Layout per CTA (persistent, one per SM): 12 epilogue warps (3 per TMEM lane quarter; tile t belongs to warp t mod 3 of
// each quarter), one MMA-issuing warp (18 MMAs of N = 32 per tile), one TMA warp.  TMEM: per row block h = 0, 1 at
// column 48 h: x~ hi [0, 8) lo [8, 16) | G hi [16, 32) lo [32, 48);  three stages [128 + 128 a, +128) =
// 2 x {S [32) | g [32)}.  Tile image (32 rows of Y, 8448 bytes, one cp.async.bulk):
//   rows 0..31   (128 B, SWIZZLE_128B): [y'' hi 32 B | y'' lo 32 B | y FP32 64 B]
//   rows 32..63  (128 B, SWIZZLE_128B): [W' tf32 hi 64 B | W' tf32 lo 64 B]
//   then 32 x {1.0f, |y_j|^2}
#include <cmath>
#include <cstdlib>

#include "gram_tf32_kernel.cuh"
#include "kmatw_tc2.cuh"
#include "tc2_prims.cuh"

namespace jrk {

using namespace tf32;
using namespace tcp;

namespace {

constexpr int GT_BM = 128, GT_NT = 32, GT_D = 16, GT_MT = 16;
constexpr int GT_R = 2;                        // row blocks (of 128) per unit: an epilogue thread owns one row of each,
                                               // so every y_j it reads from shared memory serves GT_R pairs
constexpr int GT_ROWS = GT_R * GT_BM;
constexpr int GT_NS = 3;                       // stages in TMEM: GT_R x {S [32) | g [32)} = 128 columns each
constexpr int GT_SOP = 8;                      // tile-image stages in shared memory
constexpr int GT_IMG = 8448;                   // bytes of one tile image
constexpr int GT_STAGE = 9216;                 // stage stride (1024-byte aligned)
constexpr int GT_EPQ = 3, GT_EW = 4 * GT_EPQ;  // epilogue warps per TMEM lane quarter: one per stage (2 measured no faster)
constexpr int GT_COL_X = 0, GT_COL_XLO = 8, GT_COL_G = 16, GT_COL_GLO = 32, GT_COL_BLK = 48, GT_COL_S = 128;
constexpr int GT_NACC = GT_D + 2;              // sum tt y_k (16), sum tt, sum tt |y|^2
constexpr int GT_FLUSH = 16;                   // own tiles between two flushes of the register sums into shared memory
                                               // (hierarchical accumulation: the rounding error does not grow with M)
constexpr int GT_OFF_RED = GT_SOP * GT_STAGE;
constexpr int GT_RED_WARP = GT_R * GT_NACC * 32 * 4;   // [R][18][32 lanes] floats
constexpr int GT_OFF_BAR = GT_OFF_RED + 2 * GT_EW * GT_RED_WARP;
constexpr int GT_NBAR = 2 * GT_SOP + 2 * GT_NS + 2;
constexpr int GT_SMEM = GT_OFF_BAR + GT_NBAR * 8 + 16 + 1024;
constexpr int GT_THREADS = (GT_EW + 2) * 32;
static_assert(GT_COL_S + GT_NS * 64 * GT_R <= 512 && GT_R * GT_COL_BLK <= GT_COL_S, "TMEM budget");
static_assert(GT_SMEM <= 227 * 1024, "shared memory budget");
static_assert(GT_EPQ >= 3 && GT_EPQ == GT_NS, "warps 0 / 1 / 2 of a lane quarter park x~ / G hi / G lo; one stage per warp");

struct GtParams {
    const float* X; int64_t ldx; int d;
    const float* G; int64_t ldg; int m;
    const float* xn;                  // c |x_i|^2
    const unsigned char* img;
    float* gX; int64_t ldgx;
    float* gs_rows;                   // nullable
    int64_t N, M;
    int n_tiles, n_units;
    const unsigned* nmax;             // [2] bit patterns: max |c||x|^2, max |c||y|^2
    float x_scale, inv_c;             // 2^-r of GEMM1;  1 / c (|x_i|^2 = xn_i / c)
    float fac, fac_s;
    int g256;                         // GBAR: rows of Gbar are 32-byte aligned (one 256-bit load per chunk)
    int gtrans;                       // GBAR: the cotangent is given transposed (M x N): G[j * ldg + i] -- consecutive lanes
                                      // (rows i) then read consecutive floats: one cache line per warp instruction
};

__device__ __forceinline__ void lds_2x64(uint32_t addr, f32x2& a, f32x2& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ f32x2 lds_64(uint32_t addr) {
    f32x2 a;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr));
    return a;
}

// GBAR = false: K@W cotangent, g = G . W'^T by GEMM1b (P.G: N x m).  GBAR = true: Gram cotangent, g_ij = Gbar[i, j] read
// from global memory (P.G: N x M, 16-byte aligned rows), the column factor 2^(c|y_j|^2) rides in the image's FP32 fields.
// DK = 16, or 8 for d <= 8: half the accumulators, shared-memory loads and FMAs per pair.
template <bool GBAR, int DK = GT_D>
__global__ void __launch_bounds__(GT_THREADS, 1) kgrad_tc_kernel(GtParams P) {
    constexpr int NACC = DK + 2;            // sum tt y_k (DK), sum tt, sum tt |y|^2
    if (!tc2_fast_regime(P.nmax)) return;      // kgrad.cu serves the problem
    extern __shared__ unsigned char gt_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gt_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);
    uint64_t* op_full = bars;                  // [SOP]  tile image landed
    uint64_t* op_empty = op_full + GT_SOP;     // [SOP]  the four owner warps are done with the image
    uint64_t* s_full = op_empty + GT_SOP;      // [NS]   the MMAs of the stage are complete
    uint64_t* s_empty = s_full + GT_NS;        // [NS]   the four owner warps read the stage
    uint64_t* x_ready = s_empty + GT_NS;       // [1]    x~ and G of the unit are parked
    uint64_t* x_free = x_ready + 1;            // [1]    every MMA of the unit has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GT_OFF_BAR + GT_NBAR * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_SOP; ++s) { mbar_init(&op_full[s], 1); mbar_init(&op_empty[s], 4); }
        for (int s = 0; s < GT_NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
        mbar_init(x_ready, GT_EW);
        mbar_init(x_free, 1);
        fence_mbar_init();
    }
    constexpr int W_MMA = GT_EW, W_TMA = GT_EW + 1;
    if (warp == W_MMA) tmem_alloc_512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();

    const int T = P.n_tiles;
    if (warp == W_TMA) {
        // ===================== TMA producer: one image per tile =====================
        uint32_t tc = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            for (int t = 0; t < T; ++t, ++tc) {
                const int s = tc % GT_SOP;
                mbar_wait_hint(&op_empty[s], ((tc / GT_SOP) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&op_full[s], GT_IMG);
                    bulk_g2s(smem + s * GT_STAGE, P.img + (int64_t)t * GT_IMG, GT_IMG, &op_full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: per row block 3 (fp16) + 6 (tf32) MMAs of N = 32 per tile =====================
        constexpr uint32_t id_h = (1u << 4) | ((uint32_t)(GT_NT >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
        constexpr uint32_t id_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GT_NT >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
        const uint64_t dbase = make_b_desc(smem_u32(smem));
        const uint32_t b_opfull = smem_u32(op_full), b_sfull = smem_u32(s_full), b_sempty = smem_u32(s_empty),
                       b_xready = smem_u32(x_ready), b_xfree = smem_u32(x_free);
        uint32_t gt = 0, uc = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, ++uc) {
            bar_wait(b_xready, uc & 1);
            tc_fence_after();
            for (int t = 0; t < T; ++t) {
                const uint32_t g = gt + (uint32_t)t, s = g % GT_SOP, a = g % GT_NS;
                bar_wait(b_opfull + s * 8, (g / GT_SOP) & 1);
                bar_wait(b_sempty + a * 8, ((g / GT_NS) & 1) ^ 1);
                tc_fence_after();
                const uint64_t dy = dbase + (uint64_t)((s * GT_STAGE) >> 4);
                const uint64_t dw = dy + (uint64_t)(4096 >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int h = 0; h < GT_R; ++h) {
                        const uint32_t cs = GT_COL_S + a * (64 * GT_R) + h * 64, cg = cs + 32, cx = h * GT_COL_BLK;
                        // the small cross terms first, the hi.hi terms last (the accumulator truncates)
                        mma_f16_ts(cs, cx + GT_COL_X, dy + 2, id_h, 0u);               // hi_x . lo_y
                        mma_f16_ts(cs, cx + GT_COL_XLO, dy, id_h, 1u);                 // lo_x . hi_y
                        mma_f16_ts(cs, cx + GT_COL_X, dy, id_h, 1u);                   // hi_x . hi_y
                        if (!GBAR) {
                            mma_tf32_ts(cg, cx + GT_COL_G, dw + 4, id_t, 0u);          // hi_g . lo_w  (k-steps 0, 1)
                            mma_tf32_ts(cg, cx + GT_COL_G + 8, dw + 6, id_t, 1u);
                            mma_tf32_ts(cg, cx + GT_COL_GLO, dw, id_t, 1u);            // lo_g . hi_w
                            mma_tf32_ts(cg, cx + GT_COL_GLO + 8, dw + 2, id_t, 1u);
                            mma_tf32_ts(cg, cx + GT_COL_G, dw, id_t, 1u);              // hi_g . hi_w
                            mma_tf32_ts(cg, cx + GT_COL_G + 8, dw + 2, id_t, 1u);
                        }
                    }
                    commit_addr(b_sfull + a * 8);
                }
                __syncwarp();
            }
            if (elect_one()) commit_addr(b_xfree);
            __syncwarp();
            gt += (uint32_t)T;
        }
    } else {
        // ===================== epilogue warps (and X / G loaders) =====================
        // warp k of a lane quarter owns the quarter of the tiles t = k (mod EPQ): 32 lanes x 32 columns of every row block
        const int quad = warp & 3, k = warp >> 2;
        uint32_t lane_base;
        asm volatile("mov.u32 %0, %1;" : "=r"(lane_base) : "r"(tmem_base + ((uint32_t)(quad * 32) << 16)));
        uint32_t sbase;
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
        const uint32_t b_opfull = sbase + GT_OFF_BAR, b_opempty = b_opfull + GT_SOP * 8, b_sfull = b_opempty + GT_SOP * 8,
                       b_sempty = b_sfull + GT_NS * 8, b_xready = b_sempty + GT_NS * 8, b_xfree = b_xready + 8;
        float* red = reinterpret_cast<float*>(smem + GT_OFF_RED);

        auto park = [&](int u) {
            // warp 0 of the quarter: x~ = x 2^-r as fp16 hi / lo pairs;  warps 1 / 2: G as tf32 hi / lo
#pragma unroll
            for (int h = 0; h < GT_R; ++h) {
                const int64_t i = (int64_t)u * GT_ROWS + h * GT_BM + quad * 32 + lane;
                const bool iok = i < P.N;
                const uint32_t cx = lane_base + h * GT_COL_BLK;
                if (k == 0) {
                    const float* xrow = P.X + (iok ? i : 0) * P.ldx;
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float v0 = (iok && 2 * q < P.d) ? __ldg(xrow + 2 * q) : 0.f;
                        const float v1 = (iok && 2 * q + 1 < P.d) ? __ldg(xrow + 2 * q + 1) : 0.f;
                        split_h16u_pair(v0 * P.x_scale, v1 * P.x_scale, hi[q], lo[q]);
                    }
                    tmem_st8(cx + GT_COL_X, hi);
                    tmem_st8(cx + GT_COL_XLO, lo);
                } else if (!GBAR) {
                    const float* grow = P.G + (iok ? i : 0) * P.ldg;
                    uint32_t v[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float g = (iok && c < P.m) ? __ldg(grow + c) : 0.f;
                        const uint32_t hi = __float_as_uint(g) & 0xFFFFE000u;
                        v[c] = (k == 1) ? hi : __float_as_uint(g - __uint_as_float(hi));
                    }
                    tmem_st16(cx + (k == 1 ? GT_COL_G : GT_COL_GLO), v);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(b_xready);
            __syncwarp();       // reconverge: the next tcgen05 instruction of this warp is .sync.aligned
        };

        uint32_t gt = 0, uc = 0;
        int u = blockIdx.x;
        if (u < P.n_units) park(u);
        for (; u < P.n_units; u += gridDim.x, ++uc) {
            f32x2 acc[GT_R][NACC / 2];       // per row block: (y0, y1) .. (y14, y15), (1, |y|^2)
#pragma unroll
            for (int h = 0; h < GT_R; ++h)
#pragma unroll
                for (int q = 0; q < NACC / 2; ++q) acc[h][q] = 0ull;
            // this warp's slot of the unit's reduce buffer (two buffers, alternating by unit): [R][18][32 lanes]
            float* mine = red + ((uc & 1) * GT_EW + quad * GT_EPQ + k) * (GT_R * NACC * 32);
            bool first_flush = true;
            auto flush = [&]() {
#pragma unroll
                for (int h = 0; h < GT_R; ++h)
#pragma unroll
                    for (int q = 0; q < NACC / 2; ++q) {
                        float a0, a1;
                        unpack2(acc[h][q], a0, a1);
                        float* p0 = mine + (h * NACC + 2 * q) * 32 + lane;
                        if (first_flush) { p0[0] = a0; p0[32] = a1; }
                        else { p0[0] += a0; p0[32] += a1; }
                        acc[h][q] = 0ull;
                    }
                first_flush = false;
            };
            int since = 0;
            // GBAR: this thread's two rows of Gbar; 8 columns per chunk, fetched one chunk ahead (the addresses do not
            // depend on any barrier, so the pipeline runs across tile boundaries)
            const float* grow[GT_R];
            bool gok[GT_R];
#pragma unroll
            for (int h = 0; h < GT_R; ++h) {
                const int64_t i = (int64_t)u * GT_ROWS + h * GT_BM + quad * 32 + lane;
                gok[h] = i < P.N;
                grow[h] = P.G + (gok[h] ? i : 0) * P.ldg;
            }
            auto load_gbar = [&](int t, int ch, uint32_t (&gv)[GT_R][8]) {
                const int64_t j0 = (int64_t)t * GT_NT + 8 * ch;
                if (P.gtrans) {
#pragma unroll
                    for (int h = 0; h < GT_R; ++h) {
                        const float* col = P.G + j0 * P.ldg + ((int64_t)u * GT_ROWS + h * GT_BM + quad * 32 + lane);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            gv[h][c] = (gok[h] && j0 + c < P.M) ? __float_as_uint(__ldg(col + (int64_t)c * P.ldg)) : 0u;
                    }
                    return;
                }
#pragma unroll
                for (int h = 0; h < GT_R; ++h) {
                    if (gok[h] && j0 + 8 <= P.M && P.g256) {
                        // one 32-byte sector per thread and instruction (two 16-byte loads would visit it twice: the
                        // L1TEX sector rate is what bounds this kernel, ncu: l1tex throughput 98.6 %)
                        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(gv[h][0]), "=r"(gv[h][1]), "=r"(gv[h][2]), "=r"(gv[h][3]), "=r"(gv[h][4]),
                                       "=r"(gv[h][5]), "=r"(gv[h][6]), "=r"(gv[h][7])
                                     : "l"(grow[h] + j0));
                    } else if (gok[h] && j0 + 8 <= P.M) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(grow[h] + j0));
                        const float4 b = __ldg(reinterpret_cast<const float4*>(grow[h] + j0 + 4));
                        gv[h][0] = __float_as_uint(a.x); gv[h][1] = __float_as_uint(a.y);
                        gv[h][2] = __float_as_uint(a.z); gv[h][3] = __float_as_uint(a.w);
                        gv[h][4] = __float_as_uint(b.x); gv[h][5] = __float_as_uint(b.y);
                        gv[h][6] = __float_as_uint(b.z); gv[h][7] = __float_as_uint(b.w);
                    } else {
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            gv[h][c] = (gok[h] && j0 + c < P.M) ? __float_as_uint(__ldg(grow[h] + j0 + c)) : 0u;
                    }
                }
                // The branches above are per lane.  Every tcgen05.ld / st / wait is .sync.aligned: all 32 lanes must execute it
                // together, and after divergent code that needs an explicit reconvergence point.  Without this line the
                // 8-feature build of this form computed wrong, run-to-run different rows (the compiler happened to keep the
                // other instantiations converged).
                __syncwarp();
            };
            // Tile t of the unit belongs to warp t mod EPQ (unit-local, so that a block of rows computed on its own -- another
            // rank's shard -- sums in the same order and gives the same bits).  The stage of a tile goes by the global index
            // gt + t, so the owner / stage pairing shifts from unit to unit when T is not a multiple of EPQ; that is safe
            // for the parity waits because every warp has waited for x_free -- all MMAs of the previous unit complete --
            // before its first tile of this one.
            const int first = k;
            uint32_t gnext[GT_R][8];        // (two chunks in flight spill registers and measured 20 % slower)
            if (GBAR && first < T) load_gbar(first, 0, gnext);
            for (int t = first; t < T; t += GT_EPQ) {
                const uint32_t g = gt + (uint32_t)t, a = g % GT_NS, s = g % GT_SOP;
                bar_wait(b_sfull + a * 8, (g / GT_NS) & 1);
                bar_wait(b_opfull + s * 8, (g / GT_SOP) & 1);     // (complete long ago: makes the TMA writes visible to this thread)
                tc_fence_after();
                const uint32_t sc = lane_base + GT_COL_S + a * (64 * GT_R);
                const uint32_t st = sbase + s * GT_STAGE;
                uint32_t sv[GT_R][8], gv[GT_R][8];
#pragma unroll
                for (int h = 0; h < GT_R; ++h) {
                    ld8_async(sc + h * 64, sv[h]);
                    if (!GBAR) ld8_async(sc + h * 64 + 32, gv[h]);
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    float tv[GT_R][8];
#pragma unroll
                    for (int h = 0; h < GT_R; ++h) {
                        ld8_wait(sv[h]);
                        if (!GBAR) ld8_wait(gv[h]);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            tv[h][c] = ex2_approx(__uint_as_float(sv[h][c])) * __uint_as_float(GBAR ? gnext[h][c] : gv[h][c]);
                    }
                    if (GBAR) {       // the next chunk of Gbar (of this tile, or the first one of this warp's next tile)
                        if (ch < 3) load_gbar(t, ch + 1, gnext);
                        else if (t + GT_EPQ < T) load_gbar(t + GT_EPQ, 0, gnext);
                    }
                    if (ch < 3) {
                        // the next chunk's S and g are in flight while this one goes through the FMA pipe
#pragma unroll
                        for (int h = 0; h < GT_R; ++h) {
                            ld8_async(sc + h * 64 + 8 * (ch + 1), sv[h]);
                            if (!GBAR) ld8_async(sc + h * 64 + 32 + 8 * (ch + 1), gv[h]);
                        }
                    } else {
                        tc_fence_before();
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int j = 8 * ch + c;
                        const uint32_t row = st + j * 128;
                        f32x2 y[DK / 2];
#pragma unroll
                        for (int q = 0; q < DK / 4; ++q) lds_2x64(row + (((4 + q) ^ (j & 7)) << 4), y[2 * q], y[2 * q + 1]);
                        const f32x2 on = lds_64(st + 8192 + j * 8);
#pragma unroll
                        for (int h = 0; h < GT_R; ++h) {
                            const f32x2 tt = pack2(tv[h][c], tv[h][c]);
#pragma unroll
                            for (int q = 0; q < DK / 2; ++q) acc[h][q] = fma2(tt, y[q], acc[h][q]);
                            acc[h][DK / 2] = fma2(tt, on, acc[h][DK / 2]);
                        }
                    }
                }
                // (Handing the TMEM stage back before the last chunk's FMAs -- all its values are in registers by then -- was
                // tried twice.  Round-2 re-run with reconvergence points around the arrive: K@W cotangent 6 % SLOWER (16.7 vs
                // 15.75 ms at 151552 x 131072), transposed Gram cotangent 5 % faster, and the row-major Gram cotangent came out
                // wrong in whole warps, differently from run to run (profiles/README.md).  Not understood, not used.)
                __syncwarp();
                if (lane == 0) {
                    bar_arrive(b_sempty + a * 8);
                    bar_arrive(b_opempty + s * 8);
                }
                __syncwarp();   // reconverge before the next tile's .sync.aligned tcgen05.ld
                if (++since == GT_FLUSH) { flush(); since = 0; }
            }
            // the EPQ partial results of a lane quarter meet in shared memory
            flush();
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(32 * GT_EPQ) : "memory");
            if (k == 0) {
#pragma unroll
                for (int h = 0; h < GT_R; ++h) {
                    float av[NACC];
#pragma unroll
                    for (int c = 0; c < NACC; ++c) {
                        float v = 0.f;
#pragma unroll
                        for (int kk = 0; kk < GT_EPQ; ++kk) v += mine[kk * (GT_R * NACC * 32) + (h * NACC + c) * 32 + lane];
                        av[c] = v;
                    }
                    const int64_t i = (int64_t)u * GT_ROWS + h * GT_BM + quad * 32 + lane;
                    if (i < P.N) {
                        const float xn = __ldg(P.xn + i);
                        const float rowf = ex2_approx(xn);                  // 2^(c|x_i|^2)
                        const float s0v = av[DK], s2v = av[DK + 1];
                        const float* xrow = P.X + i * P.ldx;
                        float* dst = P.gX + i * P.ldgx;
                        float xs1 = 0.f;
                        const float f = P.fac * rowf;
#pragma unroll
                        for (int c = 0; c < DK; ++c) {
                            if (c < P.d) {
                                const float xv = __ldg(xrow + c);
                                xs1 = fmaf(xv, av[c], xs1);
                                dst[c] = f * (av[c] - xv * s0v);
                            }
                        }
                        if (P.gs_rows) P.gs_rows[i] = P.fac_s * rowf * (xn * P.inv_c * s0v - 2.f * xs1 + s2v);
                    }
                }
            }
            // once every MMA of this unit has completed the next block of x~ / G can be parked
            const int un = u + gridDim.x;
            if (un < P.n_units) {
                bar_wait(b_xfree, uc & 1);
                tc_fence_after();
                park(un);
            }
            gt += (uint32_t)T;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == W_MMA) tmem_dealloc_512(tmem_base);
}

// out[r] = c |row r|^2;  *gmax = max |c| |row|^2 (atomicMax on the bit pattern)
__global__ void gt_rownorm_kernel(const float* __restrict__ A, int64_t rows, int d, int64_t lda, float c,
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

// tile images: one thread per 16-byte piece (528 per tile)
__global__ void gt_pack_kernel(const float* __restrict__ Y, int64_t M, int d, int64_t ldy, const float* __restrict__ W,
                               int m, int64_t ldw, int n_tiles, unsigned char* __restrict__ img, float y_scale,
                               float inv_c, const float* __restrict__ yn, const unsigned* __restrict__ nmax, int gbar) {
    if (!tc2_fast_regime(nmax)) return;
    constexpr int CPT = GT_IMG / 16;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_tiles * CPT) return;
    const int64_t t = idx / CPT;
    const int p = (int)(idx % CPT);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (p < 512) {
        const int rr = p >> 3, r = rr & 31, cl = (p & 7) ^ (r & 7);        // logical 16-byte chunk of the row
        const int64_t j = t * GT_NT + r;
        const bool jok = j < M;
        if (rr < 32) {
            if (cl < 4) {
                // y'' = y y_scale as fp16 hi (chunks 0, 1) / lo (chunks 2, 3), 8 features per chunk
                const int k0 = 8 * (cl & 1);
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float v0 = (jok && k0 + 2 * q < d) ? __ldg(Y + j * ldy + k0 + 2 * q) * y_scale : 0.f;
                    const float v1 = (jok && k0 + 2 * q + 1 < d) ? __ldg(Y + j * ldy + k0 + 2 * q + 1) * y_scale : 0.f;
                    uint32_t hi, lo;
                    split_h16u_pair(v0, v1, hi, lo);
                    w[q] = cl < 2 ? hi : lo;
                }
                o = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
                // y as FP32, 4 features per chunk
                const int k0 = 4 * (cl - 4);
                const float cf = (gbar && jok) ? ex2_approx(__ldg(yn + j)) : 1.f;      // Gram cotangent: 2^(c|y_j|^2) rides here
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = (jok && k0 + q < d) ? __ldg(Y + j * ldy + k0 + q) * cf : 0.f;
                o = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
            }
        } else {
            // W' = W 2^(c|y|^2) as tf32 hi (chunks 0..3) / lo (chunks 4..7), 4 columns per chunk
            const int c0 = 4 * (cl & 3);
            const float colf = jok ? ex2_approx(__ldg(yn + j)) : 0.f;
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float v = (W && jok && c0 + q < m) ? __ldg(W + j * ldw + c0 + q) * colf : 0.f;
                const uint32_t hi = __float_as_uint(v) & 0xFFFFE000u;
                w[q] = cl < 4 ? hi : __float_as_uint(v - __uint_as_float(hi));
            }
            o = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        // {1, |y_j|^2} for two consecutive rows
        const int r = 2 * (p - 512);
        const int64_t j = t * GT_NT + r;
        const float y0 = (j < M) ? __ldg(yn + j) : 0.f, y1 = (j + 1 < M) ? __ldg(yn + j + 1) : 0.f;
        const float f0 = gbar ? ex2_approx(y0) : 1.f, f1 = gbar ? ex2_approx(y1) : 1.f;
        o = make_uint4(__float_as_uint(f0), __float_as_uint(f0 * y0 * inv_c), __float_as_uint(f1), __float_as_uint(f1 * y1 * inv_c));
    }
    *reinterpret_cast<uint4*>(img + t * GT_IMG + (int64_t)p * 16) = o;
}

}  // namespace

bool kgrad_tc_eligible(int64_t N, int64_t M, int d, int m, int pm) {
    return pm == PM_SQEXP && d <= GT_D && m <= GT_MT && N >= 4096 && M >= 4096 && jrk_option(JRK_OPT_KMATW_TC) != 0 &&
           jrk_option(JRK_OPT_TC_FAST) != 0;
}
// Gram cotangent (mode 1): Gbar rows must allow 16-byte loads
bool kgrad_tc_gbar_eligible(int64_t N, int64_t M, int d, int pm, const float* Gbar, int64_t ldgb, int transposed) {
    return kgrad_tc_eligible(N, M, d, 1, pm) && (transposed || ((ldgb % 4 == 0) && ((uintptr_t)Gbar % 16 == 0)));
}

size_t kgrad_tc_workspace_bytes(int64_t N, int64_t M) {
    const int64_t n_tiles = (M + GT_NT - 1) / GT_NT;
    return align256((size_t)N * 4) + align256((size_t)M * 4) + 256 + align256((size_t)n_tiles * GT_IMG) + 1024;
}

// Enqueues the pre-pass and the tensor-core kernel; *nmax_out receives the device pointer of the two norm maxima, from
// which the caller's FP32 kernels decide (on the device) whether the result is theirs to write.
// W == nullptr: Gram cotangent, G = Gbar (N x M, or M x N when gtrans).
int kgrad_tc(const float* X, const float* Y, const float* W, const float* G, float* gX, float* gs_rows, int64_t N, int64_t M,
             int d, int m, int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldg, int64_t ldgx, float c, float fac, float fac_s,
             int gtrans, void* scratch, const unsigned** nmax_out, cudaStream_t st) {
    const bool gbar = W == nullptr;
    char* p = (char*)scratch;
    float* xn = (float*)p; p += align256((size_t)N * 4);
    float* yn = (float*)p; p += align256((size_t)M * 4);
    unsigned* nmax = (unsigned*)p; p += 256;
    unsigned char* img = (unsigned char*)p;
    const int n_tiles = (int)((M + GT_NT - 1) / GT_NT);
    JRK_CUDA_OK(cudaMemsetAsync(nmax, 0, 16, st));
    gt_rownorm_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(X, N, d, ldx, c, xn, nmax);
    JRK_LAUNCHED();
    gt_rownorm_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, c, yn, nmax + 1);
    JRK_LAUNCHED();
    const Tc2Scales sc = tc2_scales(c);
    const int64_t tot = (int64_t)n_tiles * (GT_IMG / 16);
    gt_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, W, m, ldw, n_tiles, img, sc.y_scale, 1.f / c, yn, nmax, gbar ? 1 : 0);
    JRK_LAUNCHED();
    GtParams P;
    P.X = X; P.ldx = ldx; P.d = d; P.G = G; P.ldg = ldg; P.m = m; P.xn = xn; P.img = img; P.gX = gX; P.ldgx = ldgx;
    P.gs_rows = gs_rows; P.N = N; P.M = M; P.n_tiles = n_tiles; P.n_units = (int)((N + GT_ROWS - 1) / GT_ROWS); P.nmax = nmax;
    P.x_scale = sc.x_scale; P.inv_c = 1.f / c; P.fac = fac; P.fac_s = fac_s;
    P.g256 = (gbar && ldg % 8 == 0 && (uintptr_t)G % 32 == 0) ? 1 : 0;
    P.gtrans = (gbar && gtrans) ? 1 : 0;
    const int grid = P.n_units < tf_sm_count() ? P.n_units : tf_sm_count();
#define GT_LAUNCH(GBAR_, DK_)                                                                                              \
    do {                                                                                                                   \
        JRK_CUDA_OK(cudaFuncSetAttribute(kgrad_tc_kernel<GBAR_, DK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM)); \
        kgrad_tc_kernel<GBAR_, DK_><<<grid, GT_THREADS, GT_SMEM, st>>>(P);                                                  \
    } while (0)
    if (gbar) {
        // (the 8-feature instantiation of this form is where the missing reconvergence in load_gbar showed: wrong,
        // run-to-run different rows with the row-major cotangent; every-row tests cover d = 5, 8, 16 in both layouts)
        if (d <= 8) GT_LAUNCH(true, 8);
        else GT_LAUNCH(true, 16);
    } else {
        if (d <= 8) GT_LAUNCH(false, 8);
        else GT_LAUNCH(false, 16);
    }
#undef GT_LAUNCH
    JRK_LAUNCHED();
    *nmax_out = nmax;
    return JRK_OK;
}

}  // namespace jrk

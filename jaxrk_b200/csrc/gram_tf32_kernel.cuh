// Kernel (a2): materialised Gram for d <= 128 by the compensated three-product expansion on the 5th-generation tensor
// cores (tcgen05.mma, operands and accumulators in TMEM): tf32 hi / lo operands (columns padded to 32) for d <= 16,
// fp16 hi / lo' operands (kind::f16, K = 16: half the MMAs and operand bytes; padded to 64 / 128) above -- see TfCfg.
//
// Replaces, for wide inputs, the same reference chain as gram_direct.cu:
//   sqeucldist_simple (jaxrk/utilities/eucldist.py:10-18)  sq = |x|^2 + |y|^2 - 2 x.y   <- the GEMM
//   eucldist (:37-48), ScaledPairwiseDistance (jaxrk/kern/util.py:107-116),
//   GenGaussKernel.__call__ (jaxrk/kern/rbf.py:104-105)                                  <- the epilogue
// Here the distance really is a dense contraction, so x.y goes to the tensor cores; FP32
// accuracy is recovered by splitting every operand into "hi" and "lo" parts of 11 significant bits each,
//   x.y ~= hi_x.hi_y + hi_x.lo_y + lo_x.hi_y          (three tcgen05.mma per k-step),
// and the kernel is bound by the HBM write of the N x M result (DESIGN.md "Rooflines").
//
// Roles are swapped with respect to the output: the A operand (128 TMEM lanes) is a block of
// 128 rows of Y, i.e. 128 consecutive output COLUMNS j; the B operand is a tile of NT rows of
// X (output rows i).  An accumulator lane therefore runs along the contiguous output
// dimension: a warp's 32 lanes hold 32 consecutive floats of one output row.
//
// A pre-pass kernel (gram_tf32.cu: tf_split_kernel / tf_split_h16_kernel) splits X once into hi / lo tile images that
// are already in the K-major 128B-swizzled layout the UMMA shared-memory descriptor expects, one contiguous
// 2 x NT x DP x {4 | 2} byte block per tile, so the main kernel needs no in-SM conversion pass: a tile arrives
// with ONE cp.async.bulk and is consumed by the tensor core directly.  (A first version converted raw FP32
// tiles inside the SM with four extra warps; knock-out timing showed the stages contending for the
// 128 B/clk shared-memory / L1 data pipe -- TMA fill 16 KB + converter 48 KB + MMA operand reads 48 KB +
// global stores 32 KB per 32 KB of output.  Pre-split images trade that for L2 reads, which have headroom.)
//
// One persistent CTA per SM, 18 warps:
//   warp 0      TMA producer: one cp.async.bulk per tile (hi + lo images) into a 3-stage ring
//   warp 1      MMA issuer (one elected lane) + TMEM allocation (all 512 columns)
//   warps 2-17  epilogue: tcgen05.ld the accumulators, u = a.dot + (xs_i + ys_j) with packed
//               FADD2 / FFMA2 (the row norms xs / ys come from a pre-pass kernel with the scale
//               and log2(nconst) folded in), exp2, and one coalesced 128-byte st.global per warp and
//               output row straight from registers.  (Staging the tile in shared memory for TMA
//               stores -- per-row bulk copies or one 2-D tensor store per tile -- was measured slower.)
//               At the start of every work unit the same warps load their Y rows, split them
//               and park hi / lo in TMEM (tcgen05.st): the A operand is read from TMEM, which
//               keeps it off the shared-memory port.
// TMEM (512 columns): A hi | A lo (DP columns each for tf32, DP / 2 for packed fp16) | 2 stages x NACC accumulators x NT
// columns.  The kernel is power-limited (~1000 W): k-steps that hold only zero padding are not issued, and the fp16
// form exists because halving the MMAs took cfg2 from 3.86 to 3.40 ms (profiles/README.md).
//   NACC = 3 (self-Gram): hi.hi, hi.lo and lo.hi go to separate accumulators and are combined as
//            c1 + (c2 + c3); with the split and the norms being role-independent functions,
//            K(X, X) is bitwise symmetric (jaxrk/utilities/gram.py:18 asserts G == G.T).
//   NACC = 2 (X != Y): both cross terms accumulate in one accumulator.
// Pairs whose expansion result is small against the norms (sq < (|x|^2+|y|^2)/64: coincident or
// near-duplicate points, the self-Gram diagonal) are recomputed by FP32 direct differencing, so
// the result stays within tolerance where the expansion form -- and the reference -- lose all
// relative accuracy.  For p = 2 the test is skipped when the pre-pass finds
// s^2 log2(e) (max|x|^2 + max|y|^2) <= 4, where the expansion error is < 1.4e-6 relative.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace jrk {
namespace tf32 {

constexpr int TF_BM = 128;        // rows of Y per work unit = TMEM lanes = output columns
constexpr int TF_EPI_WARP0 = 2;
constexpr int EW = 16;            // epilogue warps
constexpr int TF_THREADS = (TF_EPI_WARP0 + EW) * 32;
constexpr int S_OP = 3;           // operand stages (hi + lo images of one X tile each)
constexpr float TF_TAU = 1.0f / 64.0f;
constexpr float TF_SKIP_BOUND = 4.0f;   // |c| (max|x|^2 + max|y|^2) below which the p = 2 near-zero test is skipped

// H16: operands as fp16 hi / lo pairs (kind::f16, K = 16 per MMA) instead of tf32 (K = 8): the same 11 + 11 significant
// bits per value, HALF the MMAs and half the operand bytes.  The inputs are normalised by powers of two (2^-ex, 2^-ey
// from the pre-pass maxima, so every element is <= 1 in magnitude), lo is stored scaled by 2^11 (full fp16 precision
// again) and the cross-term accumulator is folded in with 2^-11:  x.y = 2^(ex+ey) (c1 + 2^-11 c2).
template <int DP, int NT, int NACC, bool H16 = false>
struct TfCfg {
    static constexpr int EB = H16 ? 2 : 4;                // bytes per operand element
    static constexpr int KC = DP * EB / 128;              // 128-byte k-chunks per row
    static constexpr int ACOLS = DP * EB / 4;             // TMEM columns of one half (hi or lo) of the A operand
    static constexpr int HALF_BYTES = NT * DP * EB;       // one swizzled image (hi or lo)
    static constexpr int STAGE_BYTES = 2 * HALF_BYTES;    // what one cp.async.bulk brings
    static constexpr int OFF_OP = 0;
    static constexpr int OFF_BAR = OFF_OP + S_OP * STAGE_BYTES;
    static constexpr int NUM_BARS = 2 * S_OP + 4 + 1;
    static constexpr int OFF_TMEM_SLOT = OFF_BAR + NUM_BARS * 8;
    static constexpr int BYTES = OFF_TMEM_SLOT + 16;
    static constexpr int DYN_BYTES = BYTES + 1024;        // slack to align the base to 1024 B
    static constexpr int ACC_COL0 = 2 * ACOLS;            // first accumulator column
    static_assert((DP * EB) % 128 == 0 && DP <= 128 && NT % 16 == 0, "unsupported tile");
    static_assert(HALF_BYTES % 1024 == 0, "swizzled images must be 1024 B aligned");
    static_assert(2 * ACOLS + 2 * NACC * NT <= 512, "TMEM budget");
    static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget");
};

// Epilogue constants: u = a * dot + (xs_i + ys_j), with xs = mult |x|^2 + add/2 from tf_norm_kernel (the additive
// constant is split evenly so that xs_i + ys_j stays commutative -> symmetric K)
struct TfEpi {
    float a;       // PM_SQEXP: -2c (c = -s^2 log2 e);  otherwise -2
    float add;     // PM_SQEXP: log2(nconst), folded into the exponent through xs / ys;  otherwise 0
    float w0;      // refinement threshold for w = u - tau (xs + ys): PM_SQEXP add (1 - tau), else 0
    const unsigned* norm_max;   // [2]: bit patterns of max |mult| |x|^2 and max |mult| |y|^2 (pre-pass atomicMax)
    float inv_abs_mult;         // 1 / |mult|: norm_max * inv_abs_mult = max |row|^2 (H16 normalisation)
    int d_used;                 // columns that are not zero padding (k-steps beyond them are skipped)
};

// smallest e with 2^e >= sqrt(v)  (v >= 0): the power-of-two normalisation of the fp16 operands
__device__ __forceinline__ int h16_exp(float v) {
    int e;
    frexpf(v, &e);                 // v = m 2^e, m in [0.5, 1)  =>  2^e > v
    return (e + 1) >> 1;           // ceil(e / 2)
}
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((127 + e) << 23); }
// fp16 hi / lo' split of a normalised value (|x| <= 1): hi = fp16(x), lo' = fp16((x - hi) 2^11); packed pairs
__device__ __forceinline__ void split_h16_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    const __half l0 = __float2half_rn((x0 - __half2float(h0)) * 2048.f), l1 = __float2half_rn((x1 - __half2float(h1)) * 2048.f);
    hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

// ---------------------------------------------------------------------------------
// tcgen05 / TMEM primitives (inline PTX)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128, N from the instruction descriptor, K = 8
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with fp16 operands (kind::f16): K = 16
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05 operations of this thread completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <int NC>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[NC]) {
    if constexpr (NC == 16) tmem_ld16(taddr, v);
    else tmem_ld8(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tf32 split: hi = x rounded to tf32 (nearest, ties away: add half an ulp to the bit pattern, mask),
// lo = x - hi (exact in FP32; the tensor core reads its leading 11 significant bits).  3 ALU ops.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// one lane of a converged warp (ptxas recognises the elect.sync predicate and issues tcgen05 / bulk-copy
// instructions under it without a per-thread waterfall loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity) {
    // try_wait with a suspend-time hint: the warp sleeps in hardware instead of spinning on issue slots
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (sm_100 format)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void st_global_cs(float* p, float v) {
#if defined(JRK_ST_PLAIN)        // A/B builds of the store policy (profiles/README.md)
    asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#elif defined(JRK_ST_CG)
    asm volatile("st.global.cg.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#else
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#endif
}

// ---------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------
template <int DP, int NT, int PM, int NACC, bool H16>
__global__ void __launch_bounds__(TF_THREADS, 1)
gram_tf32_kernel(const float* __restrict__ Ximg, const float* __restrict__ Xp, const float* __restrict__ Yp,
                 const float* __restrict__ xn,
                 const float* __restrict__ yn, float* __restrict__ K, int64_t N, int64_t M, int64_t ldk,
                 EpiParams ep, TfEpi te, int n_it, int tpu, int n_jb, int n_units, int ksteps) {
    using C = TfCfg<DP, NT, NACC, H16>;
    extern __shared__ unsigned char tf_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tf_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* op_full = bars;                  // [S_OP]
    uint64_t* op_empty = op_full + S_OP;       // [S_OP]
    uint64_t* acc_full = op_empty + S_OP;      // [2]
    uint64_t* acc_empty = acc_full + 2;        // [2]
    uint64_t* a_ready = acc_empty + 2;         // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM_SLOT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // consumers arrive once per WARP (after __syncwarp), not once per thread
        for (int s = 0; s < S_OP; ++s) { mbar_init(&op_full[s], 1); mbar_init(&op_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EW); }
        mbar_init(a_ready, EW);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();   // all 512 columns are ours, so the allocation starts at lane 0 / column 0

    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t tc = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int it0 = (u / n_jb) * tpu, it1 = min(it0 + tpu, n_it);
            for (int it = it0; it < it1; ++it, ++tc) {
                const int s = tc % S_OP;
                mbar_wait_hint(&op_empty[s], ((tc / S_OP) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&op_full[s], C::STAGE_BYTES);
                    bulk_g2s(smem + C::OFF_OP + s * C::STAGE_BYTES, Ximg + (int64_t)it * (C::STAGE_BYTES / 4), C::STAGE_BYTES,
                             &op_full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the (warp-uniform) control flow so that descriptors and TMEM addresses live in
        // uniform registers; only the tcgen05 instructions themselves are predicated to one elected lane.
        // (Issuing from inside an `if (lane == 0)` region makes the compiler wrap every MMA in an ELECT / R2UR
        // waterfall loop -- ~13 instructions per MMA on the critical path.)
        // instruction descriptor: FP32 accumulate; operand format tf32 (2) or fp16 (0); N, M
        constexpr uint32_t fmt = H16 ? 0u : 2u;
        constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TF_BM >> 4) << 24);
        const uint32_t op_base = smem_u32(smem + C::OFF_OP);
        uint32_t tc = 0, uc = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++uc) {
            const int it0 = (u / n_jb) * tpu, it1 = min(it0 + tpu, n_it);
            mbar_wait_hint(a_ready, uc & 1);
            for (int it = it0; it < it1; ++it, ++tc) {
                const uint32_t q = tc % S_OP, a = tc & 1;
                mbar_wait_hint(&acc_empty[a], ((tc >> 1) & 1) ^ 1);
                mbar_wait_hint(&op_full[q], (tc / S_OP) & 1);
                tc_fence_after();
                const uint64_t dh0 = make_b_desc(op_base + q * C::STAGE_BYTES);
                const uint64_t dl0 = make_b_desc(op_base + q * C::STAGE_BYTES + C::HALF_BYTES);
                const uint32_t acc = C::ACC_COL0 + a * NACC * NT;      // TMEM base is 0 (checked at start-up)
                if (elect_one()) {
                    // one k-step = 32 bytes of a row (8 tf32 or 16 fp16 values); k-steps that hold only zero padding
                    // are not issued (d = 16 padded to 32 columns: 6 instead of 12 MMAs)
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t off = (uint64_t)(((k >> 2) * NT * 128 + (k & 3) * 32) >> 4);   // start-address field
                        if (H16) {
                            mma_f16_ts(acc, k * 8, dh0 + off, idesc, k);                        // hi_y . hi_x
                            mma_f16_ts(acc + NT, k * 8, dl0 + off, idesc, k);                   // hi_y . lo_x
                            if (NACC == 3) mma_f16_ts(acc + 2 * NT, C::ACOLS + k * 8, dh0 + off, idesc, k);  // lo_y . hi_x
                            else mma_f16_ts(acc + NT, C::ACOLS + k * 8, dh0 + off, idesc, 1u);
                        } else {
                            mma_tf32_ts(acc, k * 8, dh0 + off, idesc, k);                       // hi_y . hi_x
                            mma_tf32_ts(acc + NT, k * 8, dl0 + off, idesc, k);                  // hi_y . lo_x
                            if (NACC == 3) mma_tf32_ts(acc + 2 * NT, C::ACOLS + k * 8, dh0 + off, idesc, k);  // lo_y . hi_x
                            else mma_tf32_ts(acc + NT, C::ACOLS + k * 8, dh0 + off, idesc, 1u);  // same accumulator
                        }
                    }
                    tc_commit(&op_empty[q]);   // operand stage may be refilled
                    tc_commit(&acc_full[a]);   // accumulators complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (and A loader) =====================
        const int e = warp - TF_EPI_WARP0;             // 0..EW-1
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
        const int ch = e >> 2;                         // which group of NT / G accumulator columns (G = EW / 4)
        const int jl = quad * 32 + lane;               // lane within the 128-row Y block
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        constexpr int G = EW / 4, NTH = NT / G, LC = (NTH >= 16) ? 16 : 8;   // LC: columns per tcgen05.ld
        // near-zero test needed?  (p = 2: only when the norms are large against the length scale)
        bool check = true;
        if (PM == PM_SQEXP)
            check = (__uint_as_float(__ldg(te.norm_max)) + __uint_as_float(__ldg(te.norm_max + 1))) > TF_SKIP_BOUND;
        uint32_t tc = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int jb = u % n_jb;
            const int it0 = (u / n_jb) * tpu, it1 = min(it0 + tpu, n_it);
            const int64_t j = (int64_t)jb * TF_BM + jl;
            const bool jok = j < M;
            // ---- park this unit's Y block in TMEM as hi / lo operands (16-column items dealt to the G groups) ----
            const float* yrow = Yp + (jok ? j : 0) * DP;
            if (H16) {
                // 32 values -> 16 packed columns per item; normalised by 2^-ey
                const float sy = pow2i(-h16_exp(__uint_as_float(__ldg(te.norm_max + 1)) * te.inv_abs_mult));
#pragma unroll
                for (int item = 0; item < DP / 32; ++item) {
                    if ((item % G) != ch) continue;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (jok) v = __ldg(reinterpret_cast<const float4*>(yrow + item * 32 + cc * 4));
                        split_h16_pair(v.x * sy, v.y * sy, hi[cc * 2], lo[cc * 2]);
                        split_h16_pair(v.z * sy, v.w * sy, hi[cc * 2 + 1], lo[cc * 2 + 1]);
                    }
                    tmem_st16(tmem_base + lane_addr + item * 16, hi);
                    tmem_st16(tmem_base + lane_addr + C::ACOLS + item * 16, lo);
                }
            } else {
#pragma unroll
            for (int kc = 0; kc < C::KC; ++kc) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (((kc * 2 + half) % G) != ch) continue;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (jok) v = __ldg(reinterpret_cast<const float4*>(yrow + kc * 32 + half * 16 + cc * 4));
                        split_tf32(v.x, hi[cc * 4 + 0], lo[cc * 4 + 0]);
                        split_tf32(v.y, hi[cc * 4 + 1], lo[cc * 4 + 1]);
                        split_tf32(v.z, hi[cc * 4 + 2], lo[cc * 4 + 2]);
                        split_tf32(v.w, hi[cc * 4 + 3], lo[cc * 4 + 3]);
                    }
                    tmem_st16(tmem_base + lane_addr + kc * 32 + half * 16, hi);
                    tmem_st16(tmem_base + lane_addr + C::ACOLS + kc * 32 + half * 16, lo);
                }
            }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            __syncwarp();
            const float ys = jok ? __ldg(yn + j) : 0.f;   // (scaled) |y_j|^2 + half of the folded constant
            const f32x2 ys2 = pack2(ys, ys);
            // H16: x.y = 2^(ex+ey) (c1 + 2^-11 cross): the power of two rides in `a`
            float a_eff = te.a;
            if (H16)
                a_eff *= pow2i(h16_exp(__uint_as_float(__ldg(te.norm_max)) * te.inv_abs_mult) +
                               h16_exp(__uint_as_float(__ldg(te.norm_max + 1)) * te.inv_abs_mult));
            const f32x2 a2 = pack2(a_eff, a_eff), ntau2 = pack2(-TF_TAU, -TF_TAU);
            const f32x2 lo_w2 = pack2(1.f / 2048.f, 1.f / 2048.f);

            // ---- tiles of this unit ----
            // (scaled) |x_i|^2 of this warp's rows: uniform 16-byte loads, software-pipelined one tile ahead -- loaded
            // at the top of the tile they were the kernel's top stall (24 % of the epilogue's samples: an L2 round trip
            // exposed once per tile).  xn is zero-padded to a multiple of 64 rows, so the look-ahead never faults.
            float4 xnext[NTH / 4];
#pragma unroll
            for (int g = 0; g < NTH / 4; ++g) xnext[g] = __ldg(reinterpret_cast<const float4*>(xn + (int64_t)it0 * NT + ch * NTH) + g);
            for (int it = it0; it < it1; ++it, ++tc) {
                const int a = tc & 1;
                const int64_t i0 = (int64_t)it * NT + ch * NTH;
                const int nvalid = (int)min((int64_t)NTH, N - i0);   // warp-uniform; <= 0 for rows past N
                float4 xv[NTH / 4];
#pragma unroll
                for (int g = 0; g < NTH / 4; ++g) xv[g] = xnext[g];
                if (it + 1 < it1) {
#pragma unroll
                    for (int g = 0; g < NTH / 4; ++g) xnext[g] = __ldg(reinterpret_cast<const float4*>(xn + i0 + NT) + g);
                }
                mbar_wait_hint(&acc_full[a], (tc >> 1) & 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + lane_addr + C::ACC_COL0 + a * NACC * NT + ch * NTH;
                f32x2 dot2[NTH / 2];
#pragma unroll
                for (int g = 0; g < NTH / LC; ++g) {
                    float c1[LC], c2[LC], c3[LC];
                    tmem_ldn<LC>(acc + g * LC, c1);
                    tmem_ldn<LC>(acc + NT + g * LC, c2);
                    if (NACC == 3) tmem_ldn<LC>(acc + 2 * NT + g * LC, c3);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < LC / 2; ++i) {
                        f32x2 cross = pack2(c2[2 * i], c2[2 * i + 1]);
                        if (NACC == 3) cross = add2(cross, pack2(c3[2 * i], c3[2 * i + 1]));   // symmetric in the two terms
                        dot2[g * (LC / 2) + i] = H16 ? fma2(cross, lo_w2, pack2(c1[2 * i], c1[2 * i + 1]))
                                                     : add2(pack2(c1[2 * i], c1[2 * i + 1]), cross);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);   // TMEM stage free: the next tile's MMAs overlap the math below
                __syncwarp();

                // u = a * dot + (xs_i + ys_j)  and the near-zero test  w = u - tau * (xs_i + ys_j):
                //   PM_SQEXP : a = -2c, xs/ys = c|.|^2 + log2(nconst)/2 (c < 0): u = log2 K;  refine if w > w0
                //   otherwise: a = -2,  xs/ys = |.|^2: u = sq;                                 refine if w < 0
                float uu[NTH];
                float wext = te.w0;
                if (check) {
#pragma unroll
                    for (int g = 0; g < NTH / 4; ++g) {
                        const f32x2 t01 = add2(pack2(xv[g].x, xv[g].y), ys2), t23 = add2(pack2(xv[g].z, xv[g].w), ys2);
                        const f32x2 u01 = fma2(dot2[g * 2], a2, t01), u23 = fma2(dot2[g * 2 + 1], a2, t23);
                        const f32x2 w01 = fma2(t01, ntau2, u01), w23 = fma2(t23, ntau2, u23);
                        float w0, w1, w2, w3;
                        unpack2(u01, uu[g * 4], uu[g * 4 + 1]);
                        unpack2(u23, uu[g * 4 + 2], uu[g * 4 + 3]);
                        unpack2(w01, w0, w1);
                        unpack2(w23, w2, w3);
                        if (PM == PM_SQEXP) wext = fmaxf(fmaxf(wext, w0), fmaxf(fmaxf(w1, w2), w3));
                        else wext = fminf(fminf(wext, w0), fminf(fminf(w1, w2), w3));
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < NTH / 4; ++g) {
                        const f32x2 t01 = add2(pack2(xv[g].x, xv[g].y), ys2), t23 = add2(pack2(xv[g].z, xv[g].w), ys2);
                        unpack2(fma2(dot2[g * 2], a2, t01), uu[g * 4], uu[g * 4 + 1]);
                        unpack2(fma2(dot2[g * 2 + 1], a2, t23), uu[g * 4 + 2], uu[g * 4 + 3]);
                    }
                }
                const bool refine = (PM == PM_SQEXP) ? (wext > te.w0) : (wext < te.w0);
                if (__any_sync(0xffffffffu, refine && jok)) {
                    // near-coincident pairs (the self-Gram diagonal, duplicates): the expansion has no relative
                    // accuracy left there; recompute those entries by FP32 direct differencing (rare path)
#pragma unroll
                    for (int i = 0; i < NTH; ++i) {   // unrolled: uu[] must stay in registers
                        const float t = __ldg(xn + i0 + i) + ys;
                        const float w = fmaf(t, -TF_TAU, uu[i]);
                        const bool bad = (PM == PM_SQEXP) ? (w > te.w0) : (w < te.w0);
                        if (i < nvalid && jok && bad) {
                            const float* xr = Xp + (i0 + i) * DP;
                            float acc2 = 0.f;
#pragma unroll 1
                            for (int k = 0; k < DP; ++k) {
                                const float df = __ldg(xr + k) - __ldg(yrow + k);
                                acc2 = fmaf(df, df, acc2);
                            }
                            uu[i] = (PM == PM_SQEXP) ? fmaf(acc2, ep.c, te.add) : acc2;
                        }
                    }
                }
                {
                    // (All 16 epilogue warps were released by the same acc_full barrier, so the four warps that hold the
                    // four 128-byte quarters of a 512-byte piece of a row store it at about the same time.  That matters:
                    // HBM takes the bare store pattern at 6.0 TB/s when the quarters arrive together and at 4.9 TB/s when
                    // they drift apart -- profiles/microbench/gram_epi.cu.  Three accumulator stages, which let the warps
                    // drift by two tiles, made this kernel 3 % slower.)
                    // row i of this warp's block: base + i * pitch, one 32 x 32 -> 64-bit multiply-add per store
                    const char* kb = reinterpret_cast<const char*>(K + i0 * ldk + j);
                    const uint32_t pitch = (uint32_t)ldk * 4u;   // host guarantees ldk * 4 * NT < 2^32
                    if (nvalid == NTH) {
                        if (jok) {
#pragma unroll
                            for (int i = 0; i < NTH; ++i)
                                st_global_cs((float*)(kb + (uint64_t)((uint32_t)i * pitch)),
                                             (PM == PM_SQEXP) ? ex2_approx(uu[i]) : ep.nconst * kexp_unnorm<PM>(uu[i], ep));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NTH; ++i)
                            if (i < nvalid && jok)
                                st_global_cs((float*)(kb + (uint64_t)((uint32_t)i * pitch)),
                                             (PM == PM_SQEXP) ? ex2_approx(uu[i]) : ep.nconst * kexp_unnorm<PM>(uu[i], ep));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc_512(tmem_base);
}

int tf_sm_count();

template <int DP, int NT, int PM, int NACC, bool H16 = false>
int tf_launch(const float* Ximg, const float* Xp, const float* Yp, const float* xn, const float* yn, float* K, int64_t N,
              int64_t M, int64_t ldk, const EpiParams& ep, const TfEpi& te, cudaStream_t st) {
    using C = TfCfg<DP, NT, NACC, H16>;
    auto kern = gram_tf32_kernel<DP, NT, PM, NACC, H16>;
    JRK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::DYN_BYTES));
    const int grid_max = tf_sm_count();
    const int64_t n_it = (N + NT - 1) / NT, n_jb = (M + TF_BM - 1) / TF_BM;
    // work unit = one 128-column block x a run of row tiles; aim at >= 16 units per CTA for balance.
    // Units are ordered run-major (u = run * n_jb + column block): the CTAs resident at any moment sweep
    // the same rows of K and write neighbouring 512-byte pieces of them.
    const int64_t target = (int64_t)grid_max * 16;
    int64_t cpj = (target + n_jb - 1) / n_jb;          // runs per column block
    if (cpj > n_it) cpj = n_it;
    if (cpj < 1) cpj = 1;
    const int64_t tpu = (n_it + cpj - 1) / cpj;
    cpj = (n_it + tpu - 1) / tpu;
    const int64_t n_units = n_jb * cpj;
    if (n_units > 2147483647LL || n_it > 2147483647LL)
        return fail(JRK_ERR_UNSUPPORTED, "gram_tf32x3: problem too large (%lld units)", (long long)n_units);
    const int grid = (int)(grid_max < n_units ? grid_max : n_units);
    const int per_step = H16 ? 16 : 8;
    int ksteps = (te.d_used + per_step - 1) / per_step;
    if (ksteps < 1 || ksteps > C::KC * 4) ksteps = C::KC * 4;
    kern<<<grid, TF_THREADS, C::DYN_BYTES, st>>>(Ximg, Xp, Yp, xn, yn, K, N, M, ldk, ep, te, (int)n_it, (int)tpu, (int)n_jb,
                                                 (int)n_units, ksteps);
    JRK_LAUNCHED();
    return JRK_OK;
}

// Row-tile height of the X tile images for (DP, nacc, h16): must match the dispatch table below.
inline int tf_nt(int DP, int nacc, int h16) { return (h16 || DP <= 64) ? 64 : 32; }

// One translation unit per epilogue mode (gram_tf32_pm{0,1,2}.cu) instantiates this.
//   nacc 3 = symmetric accumulation (self-Gram), 2 = X != Y;  h16: fp16 operands (DP 64 / 128), else tf32 (DP 32 / 64)
template <int PM>
int tf_run_pm(int DP, int nacc, int h16, const float* Ximg, const float* Xp, const float* Yp, const float* xn,
              const float* yn, float* K, int64_t N, int64_t M, int64_t ldk, const EpiParams& ep, const TfEpi& te,
              cudaStream_t st);

#define JRK_TF_ARGS Ximg, Xp, Yp, xn, yn, K, N, M, ldk, ep, te, st
#define JRK_TF32_DEFINE_RUN_PM(PM_)                                                                                  \
    template <>                                                                                                     \
    int tf_run_pm<PM_>(int DP, int nacc, int h16, const float* Ximg, const float* Xp, const float* Yp,              \
                       const float* xn, const float* yn, float* K, int64_t N, int64_t M, int64_t ldk,               \
                       const EpiParams& ep, const TfEpi& te, cudaStream_t st) {                                     \
        if (h16) {                                                                                                  \
            if (DP == 64) return nacc == 3 ? tf_launch<64, 64, PM_, 3, true>(JRK_TF_ARGS)                           \
                                           : tf_launch<64, 64, PM_, 2, true>(JRK_TF_ARGS);                          \
            if (DP == 128) return nacc == 3 ? tf_launch<128, 64, PM_, 3, true>(JRK_TF_ARGS)                         \
                                            : tf_launch<128, 64, PM_, 2, true>(JRK_TF_ARGS);                        \
        } else {                                                                                                    \
            if (DP == 32) return nacc == 3 ? tf_launch<32, 64, PM_, 3>(JRK_TF_ARGS) : tf_launch<32, 64, PM_, 2>(JRK_TF_ARGS); \
            if (DP == 64) return nacc == 3 ? tf_launch<64, 64, PM_, 3>(JRK_TF_ARGS) : tf_launch<64, 64, PM_, 2>(JRK_TF_ARGS); \
        }                                                                                                           \
        return fail(JRK_ERR_UNSUPPORTED, "gram tensor path: no kernel for DP = %d, h16 = %d", DP, h16);             \
    }

}  // namespace tf32
}  // namespace jrk

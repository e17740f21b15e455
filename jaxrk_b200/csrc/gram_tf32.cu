// Host side of kernel (a2) (gram_tf32_kernel.cuh): operand packing, the row-norm pre-pass and dispatch.
#include <cstdlib>

#include "gram_tf32_kernel.cuh"

namespace jrk {

int gram_direct(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, int out_kind,
                const ExtraKind& ex, cudaStream_t st);

namespace tf32 {

int tf_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {

// dense, zero-padded, optionally column-scaled copy:  out[r, c] = A[r, c] * s[c]  (c < cols), 0 otherwise
__global__ void tf_pack_kernel(const float* __restrict__ A, int64_t rows, int cols, int64_t lda,
                               const float* __restrict__ col_scale, float* __restrict__ out, int DP) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * DP) return;
    const int64_t r = idx / DP;
    const int c = (int)(idx % DP);
    float v = 0.f;
    if (c < cols) {
        v = A[r * lda + c];
        if (col_scale) v *= col_scale[c];
    }
    out[idx] = v;
}

// Tile images of X for the tensor core: for tile t (rows [t NT, (t+1) NT), zero padded) the block
//   img + t * 2 * NT * DP  holds the tf32 "hi" image followed by the "lo" image, each in the K-major
// 128B-swizzled shared-memory layout of the UMMA descriptor: 128-byte k-chunk kc of row r at byte
//   kc * NT * 128 + r * 128, its 16-byte piece c at position (c ^ (r & 7)).  One thread per 16-byte piece.
// hi = x rounded to tf32, lo = x - hi (split_tf32, the function the A-operand loader uses as well).
__global__ void tf_split_kernel(const float* __restrict__ Xp, int64_t rows, int64_t rows_padded, int DP, int NT,
                                float* __restrict__ img) {
    const int cpr = DP / 4;                                  // 16-byte pieces per row
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_padded * cpr) return;
    const int64_t row = idx / cpr;
    const int piece = (int)(idx % cpr), kc = piece >> 3, c = piece & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) v = __ldg(reinterpret_cast<const float4*>(Xp + row * DP) + piece);
    uint4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    const int64_t t = row / NT;
    const int r = (int)(row % NT);
    unsigned char* base = reinterpret_cast<unsigned char*>(img + t * 2 * NT * DP);
    const int off = kc * NT * 128 + r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(base + off) = h;
    *reinterpret_cast<uint4*>(base + (size_t)NT * DP * 4 + off) = l;
}

// The same for fp16 operands (H16 kernels): rows normalised by 2^-ex (ex from the pre-pass maximum, on the device),
// hi = fp16(x), lo' = fp16((x - hi) 2^11); a 128-byte k-chunk holds 64 values, one thread per 16-byte piece (8 values).
__global__ void tf_split_h16_kernel(const float* __restrict__ Xp, int64_t rows, int64_t rows_padded, int DP, int NT,
                                    unsigned char* __restrict__ img, const unsigned* __restrict__ nmax_x, float inv_abs_mult) {
    const int cpr = DP / 8;                                  // 16-byte pieces per row
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_padded * cpr) return;
    const float sx = pow2i(-h16_exp(__uint_as_float(__ldg(nmax_x)) * inv_abs_mult));
    const int64_t row = idx / cpr;
    const int piece = (int)(idx % cpr), kc = piece >> 3, c = piece & 7;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (row < rows) {
        v0 = __ldg(reinterpret_cast<const float4*>(Xp + row * DP) + 2 * piece);
        v1 = __ldg(reinterpret_cast<const float4*>(Xp + row * DP) + 2 * piece + 1);
    }
    uint4 h, l;
    split_h16_pair(v0.x * sx, v0.y * sx, h.x, l.x);
    split_h16_pair(v0.z * sx, v0.w * sx, h.y, l.y);
    split_h16_pair(v1.x * sx, v1.y * sx, h.z, l.z);
    split_h16_pair(v1.z * sx, v1.w * sx, h.w, l.w);
    const int64_t t = row / NT;
    const int r = (int)(row % NT);
    unsigned char* base = img + (size_t)t * 2 * NT * DP * 2;
    const int off = kc * NT * 128 + r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(base + off) = h;
    *reinterpret_cast<uint4*>(base + (size_t)NT * DP * 2 + off) = l;
}

// out[r] = mult * |row r|^2 + add (0 for padded rows) -- ONE function for X and Y, so the norms are role
// independent (bitwise symmetric self-Gram).  One warp per row: 16-byte chunks, xor-shuffle tree.  The block's
// largest |mult| |row|^2 goes to *gmax by atomicMax on the (non-negative) float's bit pattern.
__global__ void tf_norm_kernel(const float* __restrict__ A, int64_t rows, int64_t rows_padded, int DP, float mult,
                               float add, float* __restrict__ out, unsigned* __restrict__ gmax) {
    __shared__ float smax[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + w;
    float q = 0.f;
    if (r < rows && lane * 4 < DP) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(A + r * DP) + lane);
        q = v.x * v.x;
        q = fmaf(v.y, v.y, q);
        q = fmaf(v.z, v.z, q);
        q = fmaf(v.w, v.w, q);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
        if (r < rows_padded) out[r] = (r < rows) ? fmaf(mult, q, add) : 0.f;
        smax[w] = (r < rows) ? fabsf(mult) * q : 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = smax[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, smax[i]);
        if (m == m) atomicMax(gmax, __float_as_uint(m));
    }
}

// d <= 16: tf32 operands padded to 32 columns; above: fp16 hi / lo operands padded to 64 / 128 (measured, 65536^2:
// d = 16 tf32 3.05 vs fp16 3.09 ms, d = 32 tf32 3.22 vs fp16 3.13 ms, d = 64 tf32 3.86 vs fp16 3.40 ms).  Zero k-steps
// are skipped in both forms.  JRK_OPT_GRAM_H16 = 0 selects tf32 operands for d <= 64 (A/B comparison).
inline bool tf_use_h16(int d) {
    if (d <= 16) return false;
    return !(jrk_option(JRK_OPT_GRAM_H16) == 0 && d <= 64);
}
inline int tf_dp(int d) { return tf_use_h16(d) ? (d <= 64 ? 64 : 128) : (d <= 32 ? 32 : 64); }
inline bool tf_usable_in_place(const float* A, int64_t lda, int d, int DP, const float* dim_scale) {
    return !dim_scale && d == DP && lda == DP && ((uintptr_t)A % 16 == 0);
}
inline int64_t tf_pad64(int64_t n) { return (n + 63) / 64 * 64; }

}  // namespace
}  // namespace tf32

using namespace tf32;

size_t gram_tf32x3_workspace_bytes(int64_t N, int64_t M, int d) {
    if (d > 128) return 0;
    const int DP = tf_dp(d);
    return align256((size_t)N * DP * 4) + align256((size_t)M * DP * 4) + align256((size_t)tf_pad64(N) * 4) +
           align256((size_t)tf_pad64(M) * 4) + align256((size_t)tf_pad64(N) * DP * 8) + 1024;
}

int gram_tf32x3(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, void* workspace,
                size_t workspace_bytes, cudaStream_t st) {
    // d > 128 does not fit the TMEM-resident A operand of this kernel: served by direct differencing
    // (so is an output pitch beyond the kernel's 32-bit row offsets)
    if (d > 128 || ldk >= (1LL << 24))
        return gram_direct(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, 0, ExtraKind(), st);
    const int DP = tf_dp(d);
    const bool x_ok = tf_usable_in_place(X, ldx, d, DP, dim_scale);
    const bool y_ok = tf_usable_in_place(Y, ldy, d, DP, dim_scale);
    const bool same = (X == Y && ldx == ldy && N == M);
    Scratch scr(workspace, workspace_bytes, st);
    JRK_CUDA_OK(scr.reserve(gram_tf32x3_workspace_bytes(N, M, d)));
    const float *Xp = X, *Yp = Y;
    if (!x_ok) {
        float* buf = scr.take<float>((size_t)N * DP);
        const int64_t tot = N * DP;
        tf_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(X, N, d, ldx, dim_scale, buf, DP);
        JRK_LAUNCHED();
        Xp = buf;
    }
    if (same) {
        Yp = Xp;
    } else if (!y_ok) {
        float* buf = scr.take<float>((size_t)M * DP);
        const int64_t tot = M * DP;
        tf_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Y, M, d, ldy, dim_scale, buf, DP);
        JRK_LAUNCHED();
        Yp = buf;
    }

    // epilogue constants.  p == 2 with nconst > 0: everything folds into one exp2 argument.
    EpiParams ep = make_epi(scale, power, nconst);
    int pm = pick_pmode(power);
    if (pm == PM_EXTRA) return fail(JRK_ERR_UNSUPPORTED, "gram_tf32x3: sibling kernels use the direct path");
    if (pm == PM_SQEXP && !(nconst > 0.f)) {   // cannot fold log2(nconst): use the general epilogue with p = 2
        pm = PM_GENERAL;
        ep.c = scale * scale;
        ep.hp = 1.f;
    }
    TfEpi te;
    float mult;
    if (pm == PM_SQEXP) {
        mult = ep.c;
        te.a = -2.f * ep.c;
        te.add = log2f(nconst);
        te.w0 = te.add * (1.f - TF_TAU);
    } else {
        mult = 1.f;
        te.a = -2.f;
        te.add = 0.f;
        te.w0 = 0.f;
    }

    // row norms (one function for both operands) and their maxima
    const int64_t Np = tf_pad64(N), Mp = tf_pad64(M);
    float* xn = scr.take<float>((size_t)Np);
    float* yn = same ? xn : scr.take<float>((size_t)Mp);
    unsigned* nmax = scr.take<unsigned>(2);
    JRK_CUDA_OK(cudaMemsetAsync(nmax, 0, 2 * sizeof(unsigned), st));
    tf_norm_kernel<<<(unsigned)((Np + 7) / 8), 256, 0, st>>>(Xp, N, Np, DP, mult, 0.5f * te.add, xn, nmax);
    JRK_LAUNCHED();
    if (!same) {
        tf_norm_kernel<<<(unsigned)((Mp + 7) / 8), 256, 0, st>>>(Yp, M, Mp, DP, mult, 0.5f * te.add, yn, nmax + 1);
        JRK_LAUNCHED();
    } else {
        JRK_CUDA_OK(cudaMemcpyAsync(nmax + 1, nmax, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    }
    te.norm_max = nmax;

    const int nacc = same ? 3 : 2;

    // hi / lo tile images of X (pre-swizzled; one bulk copy per tile in the main kernel)
    const int h16 = tf_use_h16(d) ? 1 : 0;
    const int NT = tf_nt(DP, nacc, h16);
    te.inv_abs_mult = 1.f / fabsf(mult);
    te.d_used = d;
    float* ximg = scr.take<float>((size_t)Np * DP * 2);
    if (h16) {
        const int64_t tot = Np * (DP / 8);
        tf_split_h16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Xp, N, Np, DP, NT, reinterpret_cast<unsigned char*>(ximg),
                                                                           nmax, te.inv_abs_mult);
        JRK_LAUNCHED();
    } else {
        const int64_t tot = Np * (DP / 4);
        tf_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Xp, N, Np, DP, NT, ximg);
        JRK_LAUNCHED();
    }
    if (pm == PM_SQEXP) return tf_run_pm<PM_SQEXP>(DP, nacc, h16, ximg, Xp, Yp, xn, yn, K, N, M, ldk, ep, te, st);
    if (pm == PM_LAPLACE) return tf_run_pm<PM_LAPLACE>(DP, nacc, h16, ximg, Xp, Yp, xn, yn, K, N, M, ldk, ep, te, st);
    return tf_run_pm<PM_GENERAL>(DP, nacc, h16, ximg, Xp, Yp, xn, yn, K, N, M, ldk, ep, te, st);
}

}  // namespace jrk

// Kernel template of the fused K@W FP32-pipe path (see kmatw.cu for the design notes); instantiated per epilogue mode
// in kmatw_pm{0,1,2,3}.cu so that the 72 instantiations compile in parallel.
#pragma once
#include "common.cuh"
#include "kmatw_tc2.cuh"

namespace jrk {

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
    if (ep.skip_if_fast && tc2_fast_regime(ep.skip_if_fast)) return;     // kmatw_tc2_kernel<4, 4> serves the problem
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

template <int D, int MT, int RI, int PM>
int kw_launch(dim3 grid, cudaStream_t st, const float* X, int64_t ldx, int d, const float* dim_scale,
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


// One translation unit per epilogue mode defines this (D in {4..128}, MT in {4, 16, 32} select the instantiation).
template <int PM>
int kw_dispatch(int D, int MT, dim3 grid, cudaStream_t st, const float* X, int64_t ldx, int d, const float* dim_scale,
                const float* Yp, const float* Wp, float* out, int64_t ldo, int64_t split_stride, int64_t N, int64_t M, int m,
                int tiles_per_split, const float* row_pref, const EpiParams& ep, int epi_mode, int x_vec);

#define JRK_KMATW_DEFINE_DISPATCH(PM_)                                                                                      \
    template <>                                                                                                            \
    int kw_dispatch<PM_>(int D, int MT, dim3 grid, cudaStream_t st, const float* X, int64_t ldx, int d,                    \
                         const float* dim_scale, const float* Yp, const float* Wp, float* out, int64_t ldo,                \
                         int64_t split_stride, int64_t N, int64_t M, int m, int tiles_per_split, const float* row_pref,    \
                         const EpiParams& ep, int epi_mode, int x_vec) {                                                   \
        int rc = JRK_OK;                                                                                                   \
        switch (D * 100 + MT) {                                                                                            \
            JRK_KW_CASES(PM_)                                                                                              \
            default: rc = fail(JRK_ERR_UNSUPPORTED, "kmatw: no instantiation for D = %d, MT = %d", D, MT);                 \
        }                                                                                                                  \
        return rc;                                                                                                         \
    }

#define JRK_KW_CASE(D_, MT_, PM_)                                                                                         \
    case D_ * 100 + MT_:                                                                                                   \
        rc = kw_launch<D_, MT_, kw_ri(D_, MT_), PM_>(grid, st, X, ldx, d, dim_scale, Yp, Wp, out, ldo, split_stride, N, M, \
                                                     m, tiles_per_split, row_pref, ep, epi_mode, x_vec);                  \
        break;
#define JRK_KW_CASES_D(D_, PM_) JRK_KW_CASE(D_, 4, PM_) JRK_KW_CASE(D_, 16, PM_) JRK_KW_CASE(D_, 32, PM_)
#define JRK_KW_CASES(PM_)                                                                                                 \
    JRK_KW_CASES_D(4, PM_) JRK_KW_CASES_D(8, PM_) JRK_KW_CASES_D(16, PM_) JRK_KW_CASES_D(32, PM_) JRK_KW_CASES_D(64, PM_)  \
    JRK_KW_CASES_D(128, PM_)

}  // namespace jrk

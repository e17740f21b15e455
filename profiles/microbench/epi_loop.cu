// What does the TMEM round trip cost the K@W epilogue?  16 warps (4 per SM sub-partition), each converts "tiles" of
// 32 lanes x 64 columns: S (FP32) from TMEM -> e = ex2(S) -> fp16 hi / lo' pairs back to TMEM, exactly the instruction mix
// of kmatw_tc2.cu's epilogue, with no MMA and no barrier around it.  Variants isolate the pieces:
//   0  no TMEM at all (registers only)                      -> what the MUFU + ALU mix alone reaches
//   1  ld x16 + wait, compute, st x16 per chunk (naive)
//   2  ld x16 software-pipelined (next chunk's load in flight during the MUFU phase), st x16 per chunk, one wait::st per tile
//   3  as 2 without the stores      4  as 2 without the loads (stores only)
//   5  ld x32 pipelined, st x16         6  ld x64 once per tile, st x16 x 4
//   7  as 2 with st x8 + x8 (hi and lo halves separately)
// Output: cycles per tile per sub-partition against the MUFU bound (4 warps x 64 MUFU x 8 clk = 2048 clk per round of
// 4 tiles = 512 clk per tile).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../jaxrk_b200/csrc -o epi_loop epi_loop.cu
#include <cstdio>

#include "gram_tf32_kernel.cuh"

using namespace jrk;
using namespace jrk::tf32;

constexpr int TILES = 1000;

__device__ __forceinline__ uint32_t cvt_f16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}
template <int NR>
__device__ __forceinline__ void ld_async(uint32_t taddr, uint32_t (&r)[NR]);
template <>
__device__ __forceinline__ void ld_async<16>(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// keep the exponents tame whatever the packed pairs look like when they are read back as FP32
__device__ __forceinline__ void tame16(uint32_t (&s)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) s[c] = (s[c] & 0x007fffffu) | 0x3a800000u;
}
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

// variants of the conversion for the pipe question: which pipe do the f32 -> f16x2 conversions run on?
//   PK = 0: cvt.rn.f16x2.f32 (F2FP.F16.F32.PACK_AB)   PK = 1: PRMT (bf16-style truncation: wrong numbers, ALU pipe)
//   PK = 2: no packing at all (xor)                   PK = 3: MUFU only (no split arithmetic either)
template <int PK>
__device__ __forceinline__ void exp_split16_v(const uint32_t (&sv)[16], uint32_t (&eo)[16]) {
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
        const float e0 = ex2_approx(__uint_as_float(sv[c])), e1 = ex2_approx(__uint_as_float(sv[c + 1]));
        if (PK == 3) { eo[c >> 1] = __float_as_uint(e0); eo[8 + (c >> 1)] = __float_as_uint(e1); continue; }
        const float h0 = __uint_as_float(__float_as_uint(e0) & 0xFFFFE000u);
        const float h1 = __uint_as_float(__float_as_uint(e1) & 0xFFFFE000u);
        float l0, l1;
        unpack2(mul2(sub2(pack2(e0, e1), pack2(h0, h1)), pack2(2048.f, 2048.f)), l0, l1);
        if (PK == 0) { eo[c >> 1] = cvt_f16x2(h1, h0); eo[8 + (c >> 1)] = cvt_f16x2(l1, l0); }
        else if (PK == 1) { eo[c >> 1] = __byte_perm(__float_as_uint(h0), __float_as_uint(h1), 0x7632); eo[8 + (c >> 1)] = __byte_perm(__float_as_uint(l0), __float_as_uint(l1), 0x7632); }
        else { eo[c >> 1] = __float_as_uint(h0) ^ __float_as_uint(h1); eo[8 + (c >> 1)] = __float_as_uint(l0) ^ __float_as_uint(l1); }
    }
}

template <int V, int WPS = 4>
__global__ void __launch_bounds__(WPS * 4 * 32, 1) epi_kernel(long long* cyc, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc_512(smem_u32(&slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * (WPS == 4 ? 128 : 64);   // 2 x 64 (or 1 x 64) columns per warp
    uint32_t acc = 0;
    {   // initialise the warp's TMEM block
        uint32_t z[16];
        for (int c = 0; c < 16; ++c) z[c] = 0x3a800000u + lane + c;
        for (int q = 0; q < (WPS == 4 ? 8 : 4); ++q) tmem_st16(tb + 16 * q, z);
        tmem_st_wait();
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < TILES; ++t) {
        const uint32_t sc = tb + (WPS == 4 ? (t & 1) * 64 : 0);
        uint32_t s0[16], s1[16], eo[16];
        if (V == 0 || V >= 8) {
#pragma unroll
            for (int c = 0; c < 16; ++c) s0[c] = 0x3a800000u + lane + c + t;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (V == 0) exp_split16(s0, eo);
                else exp_split16_v<V - 8>(s0, eo);
#pragma unroll
                for (int c = 0; c < 16; ++c) { acc ^= eo[c]; s0[c] = (eo[c] & 0x007fffffu) | 0x3a800000u; }
            }
        } else if (V == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ld_async<16>(sc + 16 * q, s0);
                ld_wait16(s0);
                tame16(s0);
                exp_split16(s0, eo);
                tmem_st16(sc + 16 * q, eo);
            }
            tmem_st_wait();
        } else if (V == 2 || V == 3 || V == 7) {
            ld_async<16>(sc, s0);
            ld_wait16(s0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t(&cur)[16] = (q & 1) ? s1 : s0;
                uint32_t(&nxt)[16] = (q & 1) ? s0 : s1;
                if (q < 3) ld_async<16>(sc + 16 * (q + 1), nxt);
                tame16(cur);
                exp_split16(cur, eo);
                if (V == 2) tmem_st16(sc + 16 * q, eo);
                else if (V == 7) { st8(sc + 16 * q, eo); st8(sc + 16 * q + 8, eo + 8); }
                else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc ^= eo[c];
                }
                if (q < 3) ld_wait16(nxt);
            }
            if (V != 3) tmem_st_wait();
        } else if (V == 4) {
#pragma unroll
            for (int c = 0; c < 16; ++c) s0[c] = 0x3a800000u + lane + c + t;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                exp_split16(s0, eo);
                tmem_st16(sc + 16 * q, eo);
#pragma unroll
                for (int c = 0; c < 16; ++c) s0[c] = (eo[c] & 0x007fffffu) | 0x3a800000u;
            }
            tmem_st_wait();
        } else if (V == 5 || V == 6) {
            // wide loads: x32 twice (pipelined) or x64 once
            uint32_t w[64];
            if (V == 6) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                    "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, "
                    "%45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
                      "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]),
                      "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]),
                      "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31]), "=r"(w[32]), "=r"(w[33]), "=r"(w[34]), "=r"(w[35]), "=r"(w[36]),
                      "=r"(w[37]), "=r"(w[38]), "=r"(w[39]), "=r"(w[40]), "=r"(w[41]), "=r"(w[42]), "=r"(w[43]), "=r"(w[44]), "=r"(w[45]),
                      "=r"(w[46]), "=r"(w[47]), "=r"(w[48]), "=r"(w[49]), "=r"(w[50]), "=r"(w[51]), "=r"(w[52]), "=r"(w[53]), "=r"(w[54]),
                      "=r"(w[55]), "=r"(w[56]), "=r"(w[57]), "=r"(w[58]), "=r"(w[59]), "=r"(w[60]), "=r"(w[61]), "=r"(w[62]), "=r"(w[63])
                    : "r"(sc)
                    : "memory");
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(w[32 * h + 0]), "=r"(w[32 * h + 1]), "=r"(w[32 * h + 2]), "=r"(w[32 * h + 3]), "=r"(w[32 * h + 4]),
                          "=r"(w[32 * h + 5]), "=r"(w[32 * h + 6]), "=r"(w[32 * h + 7]), "=r"(w[32 * h + 8]), "=r"(w[32 * h + 9]),
                          "=r"(w[32 * h + 10]), "=r"(w[32 * h + 11]), "=r"(w[32 * h + 12]), "=r"(w[32 * h + 13]), "=r"(w[32 * h + 14]),
                          "=r"(w[32 * h + 15]), "=r"(w[32 * h + 16]), "=r"(w[32 * h + 17]), "=r"(w[32 * h + 18]), "=r"(w[32 * h + 19]),
                          "=r"(w[32 * h + 20]), "=r"(w[32 * h + 21]), "=r"(w[32 * h + 22]), "=r"(w[32 * h + 23]), "=r"(w[32 * h + 24]),
                          "=r"(w[32 * h + 25]), "=r"(w[32 * h + 26]), "=r"(w[32 * h + 27]), "=r"(w[32 * h + 28]), "=r"(w[32 * h + 29]),
                          "=r"(w[32 * h + 30]), "=r"(w[32 * h + 31])
                        : "r"(sc + 32 * h)
                        : "memory");
                }
                tmem_ld_wait();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int c = 0; c < 16; ++c) s0[c] = (w[16 * q + c] & 0x007fffffu) | 0x3a800000u;
                exp_split16(s0, eo);
                tmem_st16(sc + 16 * q, eo);
            }
            tmem_st_wait();
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    sink[tid] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(slot);
}

template <int V, int WPS = 4>
void run(long long* cyc, uint32_t* sink, const char* what) {
    epi_kernel<V, WPS><<<1, WPS * 4 * 32>>>(cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", V, cudaGetErrorString(e)); return; }
    long long h[32];
    cudaMemcpy(h, cyc, WPS * 4 * 8, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int w = 0; w < WPS * 4; ++w) mx = h[w] > mx ? (double)h[w] : mx;
    // a sub-partition's 4 warps convert 4 x TILES quarter-tiles: TILES "rounds" of 4 x 64 MUFU x 8 clk = 2048 clk
    printf("{\"variant\": %d, \"warps_per_subpartition\": %d, \"what\": \"%s\", \"clk_per_tile\": %.1f, \"mufu_frac\": %.3f}\n", V, WPS, what,
           mx / TILES / WPS, 512.0 * WPS * TILES / mx);
}

int main() {
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, 32 * 8);
    cudaMalloc(&sink, 32 * 32 * 4);
    run<0>(cyc, sink, "registers only");
    run<1>(cyc, sink, "ld16 + wait, st16 per chunk");
    run<2>(cyc, sink, "ld16 pipelined, st16");
    run<3>(cyc, sink, "ld16 pipelined, no stores");
    run<4>(cyc, sink, "st16 only");
    run<5>(cyc, sink, "ld32 x 2, st16");
    run<6>(cyc, sink, "ld64, st16");
    run<7>(cyc, sink, "ld16 pipelined, st8 + st8");
    run<8>(cyc, sink, "registers only, F2FP packing (= variant 0)");
    run<9>(cyc, sink, "registers only, PRMT packing instead of F2FP");
    run<10>(cyc, sink, "registers only, no packing");
    run<11>(cyc, sink, "registers only, MUFU alone");
    run<0, 6>(cyc, sink, "registers only");
    run<2, 6>(cyc, sink, "ld16 pipelined, st16");
    run<0, 8>(cyc, sink, "registers only");
    run<2, 8>(cyc, sink, "ld16 pipelined, st16");
    run<0, 2>(cyc, sink, "registers only");
    run<2, 2>(cyc, sink, "ld16 pipelined, st16");
    run<0, 3>(cyc, sink, "registers only");
    run<2, 3>(cyc, sink, "ld16 pipelined, st16");
    return 0;
}

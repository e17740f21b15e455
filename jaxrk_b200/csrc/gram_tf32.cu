// Kernel (a2) placeholder -- replaced by the tcgen05 3xTF32 kernel (see git log).
#include "common.cuh"
namespace jrk {
int gram_direct(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, int out_kind,
                cudaStream_t st);
size_t gram_tf32x3_workspace_bytes(int64_t, int64_t, int) { return 0; }
int gram_tf32x3(const float* X, const float* Y, float* K, int64_t N, int64_t M, int d, int64_t ldx, int64_t ldy,
                int64_t ldk, float scale, const float* dim_scale, float power, float nconst, void*, size_t,
                cudaStream_t st) {
    return gram_direct(X, Y, K, N, M, d, ldx, ldy, ldk, scale, dim_scale, power, nconst, 0, st);
}
}  // namespace jrk

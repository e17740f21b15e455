/* jrk.h -- C ABI of the B200-native Gram / fused K@W path for JaxRK.
 *
 * This is the drop-in boundary (DESIGN.md "Boundary").  The reference
 * (zalandoresearch/JaxRK) has no native layer: the path is pure Python over jax.numpy.
 * Every entry point below therefore cites the reference Python function it replaces;
 * a maintainer binds these symbols with jax.ffi (XLA FFI handler shim in
 * jaxrk_b200/csrc/xla_ffi_shim.cc) or ctypes (jaxrk_b200/_native.py).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / XLA types.
 *   - All matrices are float32, row-major, with explicit leading dimensions (elements).
 *   - Device entry points take DEVICE pointers, enqueue on `stream` (a cudaStream_t
 *     passed as void*), never synchronise, retain nothing, never write their inputs.
 *     Scratch memory, when needed, is taken from `workspace` if it is large enough
 *     (query with the *_workspace_bytes function) and from the stream-ordered
 *     allocator (cudaMallocAsync / cudaFreeAsync on `stream`) otherwise.
 *   - *_host entry points take HOST pointers, run on CUDA device `device`, include
 *     the host<->device copies, and return after the result is in host memory.
 *     Their device buffers come from the device's default stream-ordered pool, whose
 *     release threshold they raise so that the memory stays with the process between
 *     calls (no cudaMalloc / cudaFree in steady state); the caller's current device is
 *     restored before they return.
 *   - Return value: 0 = ok; non-zero = JRK_ERR_*.  jrk_last_error() gives the text.
 *     Nothing throws.  There is no CPU fallback: without a CUDA device every compute
 *     entry point returns JRK_ERR_CUDA.
 *
 * Kernel family (GenGaussKernel, /root/reference/jaxrk/kern/rbf.py:27-105):
 *     k(x, y) = nconst * exp( -( scale * || ds*x - ds*y ||_2 )^power ),  power in (0, 2]
 *   scale     global 1/length_scale   (SimpleScaler.s, kern/util.py:34-50; gs in util.py:92-99)
 *   dim_scale optional per-dimension ds (len d), applied to both inputs (util.py:92-99);
 *             NULL = none.  When given, `scale` is normally 1.
 *   nconst    power / (2 * scale_param * Gamma(1/power))   (kern/rbf.py:34), computed
 *             by the caller exactly as the reference does.
 */
#ifndef JRK_H_
#define JRK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JRK_ABI_VERSION 3

/* error codes */
#define JRK_OK 0
#define JRK_ERR_INVALID_ARG 1   /* bad shape / pointer / parameter            */
#define JRK_ERR_CUDA 2          /* CUDA runtime error (no device, launch, OOM) */
#define JRK_ERR_UNSUPPORTED 3   /* valid request this build cannot serve       */

/* distance path of jrk_gram_f32 */
#define JRK_PATH_AUTO 0    /* tcgen05 (3-product hi/lo split) for d >= 64, and for 12 <= d < 64 when N*M >= 2^24; else direct */
#define JRK_PATH_DIRECT 1  /* FP32 direct differencing (sum_k (x_k - y_k)^2)  */
#define JRK_PATH_TF32X3 2  /* compensated 3xTF32 expansion on tcgen05/TMEM    */

/* row reduction folded into jrk_kmatw_f32 (jaxrk/reduce/base.py) */
#define JRK_REDUCE_NONE 0      /* out is N x m                                             */
#define JRK_REDUCE_SUM 1       /* Sum  (reduce/base.py:132-140): out is 1 x m              */
#define JRK_REDUCE_MEAN 2      /* Mean (reduce/base.py:142-150): out is 1 x m              */
#define JRK_REDUCE_BALANCED 3  /* BalancedRed(group, average=False) (:152-181): (N/group) x m */
#define JRK_REDUCE_BALANCED_MEAN 4 /* BalancedRed(group, average=True)                     */

/* kernel families of the *_kind_f32 entry points: all share the ScaledPairwiseDistance front-end
 * (jaxrk/kern/util.py:71-116) and differ in the epilogue.  `kparams` is a HOST array:
 *   JRK_KIND_GENGAUSS     {scale, power, nconst}                        GenGaussKernel    jaxrk/kern/rbf.py:104-105
 *   JRK_KIND_PERIODIC     {1/period, length_scale}                      PeriodicKernel    jaxrk/kern/rbf.py:109-154
 *                         k = exp(-2 (sin(pi ||x-y|| / period) / length_scale)^2)
 *   JRK_KIND_THRESHSPIKE  {scale, power, spike, non_spike, threshold}   ThreshSpikeKernel jaxrk/kern/rbf.py:156-206
 *                         k = (scale ||x-y||)^power <= threshold ? spike : non_spike   (power > 0), evaluated as
 *                         ||x-y||^2 <= (threshold^(1/power) / scale)^2 */
#define JRK_KIND_GENGAUSS 0
#define JRK_KIND_PERIODIC 1
#define JRK_KIND_THRESHSPIKE 2

/* process-wide switches (jrk_set_option / jrk_get_option).  The environment variable of the same name as the
 * option (JRK_KMATW_TC, JRK_TC_FAST, JRK_TC_V2, JRK_TC2_EPQ, JRK_GRAM_H16, JRK_TC2_TOKEN) only provides the INITIAL value, read once
 * per process; nothing in the compute entry points reads the environment. */
#define JRK_OPT_KMATW_TC 0  /* 1 (default): jrk_kmatw_f32 uses the tensor-core kernels where eligible; 0: FP32-pipe kernel */
#define JRK_OPT_TC_FAST 1   /* 1 (default): factored p = 2 regime of the tensor-core K@W; 0: generic epilogue always   */
#define JRK_OPT_TC_V2 2     /* 1 (default): second-generation kernel (kmatw_tc2.cu) in the factored regime             */
#define JRK_OPT_TC2_EPQ 3   /* epilogue warps per SM sub-partition of the second-generation kernel: 4 (default) or 3   */
#define JRK_OPT_GRAM_H16 4  /* 1 (default): fp16 hi / lo operands for the tensor Gram at 16 < d <= 64; 0: tf32        */
#define JRK_OPT_TC2_TOKEN 5 /* stagger token of the second-generation kernel's epilogue warps: 0 off, 1..3 chunks        */
#define JRK_OPT_COUNT_ 6

int jrk_abi_version(void);

/* Returns 0, or JRK_ERR_INVALID_ARG for an unknown option / value. */
int jrk_set_option(int option, int value);
int jrk_get_option(int option);

/* Copies the calling thread's last error message (NUL-terminated, truncated to n). */
int jrk_last_error(char *buf, size_t n);

/* Number of visible CUDA devices (0 when there is none / no driver). */
int jrk_device_count(void);

/* ------------------------------------------------------------------------------
 * Gram matrix  K[i, j] = k(x_i, y_j),  N x M.
 * Replaces GenGaussKernel.__call__(X, Y, diag=False)  (jaxrk/kern/rbf.py:104-105)
 *   -> ScaledPairwiseDistance.__call__                (jaxrk/kern/util.py:107-116)
 *   -> dist / eucldist / sqeucldist_simple            (jaxrk/utilities/distances.py:44-49,
 *                                                      jaxrk/utilities/eucldist.py:10-18,37-48)
 * Y == NULL or Y == X means the self-Gram (Kernel.__call__ with Y=None,
 * jaxrk/kern/base.py:23-36); the result is then bitwise symmetric with an exact
 * diagonal (jaxrk/utilities/gram.py:18 asserts G == G.T).
 * ------------------------------------------------------------------------------ */
int jrk_gram_f32(const float *X, const float *Y, float *K,
                 int64_t N, int64_t M, int d,
                 int64_t ldx, int64_t ldy, int64_t ldk,
                 float scale, const float *dim_scale, float power, float nconst,
                 int path, void *workspace, size_t workspace_bytes, void *stream);

size_t jrk_gram_workspace_bytes(int64_t N, int64_t M, int d, int path);

/* ------------------------------------------------------------------------------
 * Scaled distance matrix  D[i, j] = (scale * ||ds*x_i - ds*y_j||_2)^power  (no exp).
 * Replaces ScaledPairwiseDistance.__call__(X, Y, diag=False) (jaxrk/kern/util.py:107-116)
 * and, with scale = 1, eucldist(a, b, power) (jaxrk/utilities/eucldist.py:37-48).
 * Direct differencing only.
 * ------------------------------------------------------------------------------ */
int jrk_dist_f32(const float *X, const float *Y, float *D,
                 int64_t N, int64_t M, int d,
                 int64_t ldx, int64_t ldy, int64_t ldd,
                 float scale, const float *dim_scale, float power, void *stream);

/* ------------------------------------------------------------------------------
 * Fused, never-materialised product with folded reduce maps:
 *     T[i, c]  = row_pref[i] * sum_j k(x_i, y_j) * col_pref[j] * W[j, c]      (N x m)
 *     out      = row_reduce(T)
 * Replaces FiniteVec.inner(Y) (jaxrk/rkhs/vector.py:62-75) =
 *     Reduce.apply(Reduce.apply(K, self.reduce, 0), Y.reduce, 1)  (jaxrk/reduce/base.py:39-47)
 * for reduce chains that fold to: Y side  W = (LinearReduce / Sum / Mean / Prefactors
 * chain)^T (jaxrk/reduce/lincomb.py:131-144, reduce/base.py:84-150), X side
 * Prefactors then Sum / Mean / BalancedRed; and FiniteVec.__call__ (vector.py:203-204),
 * FiniteOp.__matmul__'s inner() (jaxrk/rkhs/operator.py:46).
 * W is M x m (ldw), out is rows x m (ldo) with rows = N, 1 or N/group.
 * row_pref (len N) and col_pref (len M) are nullable.  `group` is used by the
 * BALANCED modes only and must divide N.
 * ------------------------------------------------------------------------------ */
int jrk_kmatw_f32(const float *X, const float *Y, const float *W, float *out,
                  int64_t N, int64_t M, int d, int m,
                  int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo,
                  float scale, const float *dim_scale, float power, float nconst,
                  const float *row_pref, const float *col_pref,
                  int row_reduce, int64_t group,
                  void *workspace, size_t workspace_bytes, void *stream);

size_t jrk_kmatw_workspace_bytes(int64_t N, int64_t M, int d, int m, int row_reduce, int64_t group);

/* ------------------------------------------------------------------------------
 * The same two operations for any kernel family of the ScaledPairwiseDistance front-end (JRK_KIND_*).
 * JRK_KIND_GENGAUSS forwards to jrk_gram_f32 / jrk_kmatw_f32 (all paths); the sibling kernels
 * (PeriodicKernel / ThreshSpikeKernel, jaxrk/kern/rbf.py:107-206) run on the direct-differencing Gram kernel
 * and the FP32-pipe K@W kernel with their own epilogue.  The self-Gram (Y == NULL or Y == X) is bitwise
 * symmetric with an exact diagonal (distance 0: Periodic 1, ThreshSpike `spike`).
 * ------------------------------------------------------------------------------ */
int jrk_gram_kind_f32(const float *X, const float *Y, float *K,
                      int64_t N, int64_t M, int d,
                      int64_t ldx, int64_t ldy, int64_t ldk,
                      int kind, const float *kparams, int n_kparams, const float *dim_scale,
                      int path, void *workspace, size_t workspace_bytes, void *stream);

int jrk_kmatw_kind_f32(const float *X, const float *Y, const float *W, float *out,
                       int64_t N, int64_t M, int d, int m,
                       int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldo,
                       int kind, const float *kparams, int n_kparams, const float *dim_scale,
                       const float *row_pref, const float *col_pref,
                       int row_reduce, int64_t group,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------
 * Row-wise ("diag = True") scaled distance, out[i] = scale * sum_k |ds_k x_ik - ds_k y_ik|^power  (len N).
 * Replaces the diag branch of ScaledPairwiseDistance.__call__ (jaxrk/kern/util.py:108-113) -- NB a different
 * formula from the full matrix (l_p^p of the scaled difference, global scale applied linearly).
 * ------------------------------------------------------------------------------ */
int jrk_dist_diag_f32(const float *X, const float *Y, float *out, int64_t N, int d,
                      int64_t ldx, int64_t ldy, float scale, const float *dim_scale, float power, void *stream);

/* ------------------------------------------------------------------------------
 * One pass over a materialised Gram (in place when out == K):
 *     out[i, j] = K[i, j] * row_mul[i] * col_mul[j] + row_add[i] + col_add[j] + add_const
 * -- Prefactors on both axes (jaxrk/reduce/base.py:84-101) and Center on either or both axes
 * (jaxrk/reduce/base.py:183-192: C G C = G - 1 colmean^T - rowmean 1^T + mean) applied to the N x M array in ONE read +
 * write instead of one pass per chain element.  All four vectors are nullable.
 * ------------------------------------------------------------------------------ */
int jrk_gram_affine_f32(const float *K, float *out, int64_t N, int64_t M, int64_t ldk, int64_t ldo,
                        const float *row_mul, const float *col_mul, const float *row_add, const float *col_add,
                        float add_const, void *stream);

/* 1 when jrk_kmatw_f32 would run both contractions of this problem on the tensor cores (kmatw_tc.cu / kmatw_tc2.cu:
 * d <= 32, any m (column blocks of 16), no per-dimension scale, row_reduce NONE / SUM / MEAN, N >= 8192, M >= 4096,
 * JRK_OPT_KMATW_TC != 0), else 0 (FP32-pipe kernel).  (32 < d <= 64 with power = 2 and row_reduce NONE also runs on the
 * tensor cores when the data lie inside the factored regime -- known on the device only, so this function answers 0.)
 * Introspection only (bench.py picks the roofline by it); results agree within the documented tolerance. */
int jrk_kmatw_uses_tensor_cores(int64_t N, int64_t M, int d, int m, int has_dim_scale, int row_reduce);

/* ------------------------------------------------------------------------------
 * Group-reduced Gram:  out[a, b] = sum_{i in group a} sum_{j in group b}
 *                                   row_pref[i] * k(x_i, y_j) * col_pref[j]
 * (times 1/(gx*gy) when `average`), out is (N/gx) x (M/gy).
 * Replaces FiniteVec.inner with BalancedRed on both vectors (jaxrk/reduce/base.py:152-181
 * applied on axis 0 and 1 by jaxrk/rkhs/vector.py:73-74), never materialising K.
 * ------------------------------------------------------------------------------ */
int jrk_gram_groupsum_f32(const float *X, const float *Y, float *out,
                          int64_t N, int64_t M, int d,
                          int64_t ldx, int64_t ldy, int64_t ldo,
                          float scale, const float *dim_scale, float power, float nconst,
                          const float *row_pref, const float *col_pref,
                          int64_t gx, int64_t gy, int average,
                          void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------
 * Vector-Jacobian products of the two operations above with respect to the points and the global scale
 * (what jax.grad derives from the reference's materialised chain; README.md:3, examples/Using_FLAX_Optax.ipynb,
 * jaxrk/rkhs/vector.py:206-222), fused and never materialising K:
 *   jrk_kmatw_grad_f32:  out = K(X, Y) W,  G = dL/dout (N x m):
 *       gX[i, :]   = sum_j (G[i, :] . W[j, :]) dK_ij/dx_i                       (N x d)
 *       gs_rows[i] = sum_j (G[i, :] . W[j, :]) dK_ij/dscale  at fixed nconst    (len N, nullable; dL/dscale = its sum)
 *       gp_rows[i] = sum_j (G[i, :] . W[j, :]) dK_ij/dpower  at fixed nconst    (len N, nullable; dL/dpower = its sum;
 *                    dK/dp = -K (s r)^p ln(s r); asking for it runs the general-p epilogue, also for p = 2 and p = 1)
 *   jrk_gram_grad_f32:   K = K(X, Y),  Gbar = dL/dK (N x M): the same with G[i,:].W[j,:] replaced by Gbar[i, j].
 *       gbar_transposed != 0: the cotangent is passed as Gbar^T (M x N, ldgb >= N).  The kernels read it fastest in that
 *       layout (consecutive rows i are consecutive floats), so dL/dY -- the call with the roles exchanged, (Y, X, cotangent
 *       of shape N x M = its transpose) -- takes dL/dK as it is, and dL/dX wants one transposed copy.
 * dL/dY is the same call with the roles exchanged -- (Y, X, G, W) resp. (Y, X, Gbar^T) -- and dL/dW = K(Y, X) G is
 * jrk_kmatw_f32.  GenGauss family only, global scale only (per-dimension scales are a pre-scaling of the points: their
 * gradient follows from gX / gY by the chain rule on the caller's side), d <= 64, m <= 32.
 * Pairs at distance zero contribute nothing (p <= 1: the derivative does not exist there, jax.grad gives NaN).
 * ------------------------------------------------------------------------------ */
int jrk_kmatw_grad_f32(const float *X, const float *Y, const float *W, const float *G,
                       float *gX, float *gs_rows, float *gp_rows,
                       int64_t N, int64_t M, int d, int m,
                       int64_t ldx, int64_t ldy, int64_t ldw, int64_t ldg, int64_t ldgx,
                       float scale, float power, float nconst,
                       void *workspace, size_t workspace_bytes, void *stream);

int jrk_gram_grad_f32(const float *X, const float *Y, const float *Gbar,
                      float *gX, float *gs_rows, float *gp_rows,
                      int64_t N, int64_t M, int d,
                      int64_t ldx, int64_t ldy, int64_t ldgb, int64_t ldgx, int gbar_transposed,
                      float scale, float power, float nconst,
                      void *workspace, size_t workspace_bytes, void *stream);

size_t jrk_grad_workspace_bytes(int64_t N, int64_t M, int d);

/* ------------------------------------------------------------------------------
 * Host-buffer (end-to-end) variants: same semantics, HOST pointers, copies included.
 * The Gram variant streams row blocks (compute of block b+1 overlaps the device->host
 * copy of block b).  `device` is the CUDA ordinal.
 * ------------------------------------------------------------------------------ */
int jrk_gram_f32_host(const float *X, const float *Y, float *K,
                      int64_t N, int64_t M, int d,
                      float scale, const float *dim_scale, float power, float nconst,
                      int path, int device);

int jrk_kmatw_f32_host(const float *X, const float *Y, const float *W, float *out,
                       int64_t N, int64_t M, int d, int m,
                       float scale, const float *dim_scale, float power, float nconst,
                       const float *row_pref, const float *col_pref,
                       int row_reduce, int64_t group, int device);

/* ------------------------------------------------------------------------------
 * Plans: the Y / W side of jrk_kmatw_f32 kept on the device across applies.
 * The reference evaluates FiniteVec.inner / FiniteVec.__call__ / FiniteOp.__matmul__ (jaxrk/rkhs/vector.py:62-75,
 * 203-204, jaxrk/rkhs/operator.py:33-56) against the same inspection points and linear map for every new batch of
 * points; a plan is that fixed side: device copies of Y, W, col_pref, dim_scale and -- where the tensor-core kernels
 * apply -- the packed tile images, so an apply moves and touches X only.
 *   create      Y (M x d, ldy), W (M x m, ldw), col_pref (len M, nullable), dim_scale (len d, nullable) are HOST pointers
 *               when src_is_host != 0 (`stream` is then ignored), else DEVICE pointers on `device`, read on `stream`.
 *               Returns after the plan is complete (it synchronises); nothing of the caller's is retained.
 *   apply       DEVICE pointers X (N x d, ldx), out (rows x m, ldo), row_pref (len N, nullable); enqueues on `stream`,
 *               never synchronises; scratch as in jrk_kmatw_f32 (jrk_kmatw_plan_workspace_bytes).
 *   apply_host  HOST pointers, dense (ld = d / m); copies inside, returns when `out` is complete.
 * Results are those of jrk_kmatw_f32 on the same inputs.  A plan is not thread-safe (one apply at a time).
 * ------------------------------------------------------------------------------ */
typedef struct jrk_kmatw_plan jrk_kmatw_plan;

int jrk_kmatw_plan_create(jrk_kmatw_plan **plan, const float *Y, const float *W,
                          int64_t M, int d, int m, int64_t ldy, int64_t ldw,
                          float scale, const float *dim_scale, float power, float nconst,
                          const float *col_pref, int src_is_host, int device, void *stream);

size_t jrk_kmatw_plan_workspace_bytes(const jrk_kmatw_plan *plan, int64_t N, int row_reduce, int64_t group);

int jrk_kmatw_plan_apply(const jrk_kmatw_plan *plan, const float *X, float *out,
                         int64_t N, int64_t ldx, int64_t ldo,
                         const float *row_pref, int row_reduce, int64_t group,
                         void *workspace, size_t workspace_bytes, void *stream);

int jrk_kmatw_plan_apply_host(jrk_kmatw_plan *plan, const float *X, float *out, int64_t N,
                              const float *row_pref, int row_reduce, int64_t group);

int jrk_kmatw_plan_destroy(jrk_kmatw_plan *plan);

/* Number of kernel launches issued by this library in the calling process (all
 * threads), for bench.py's `gpu_launches` claim. */
int64_t jrk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JRK_H_ */

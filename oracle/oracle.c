/* C restatement of the JaxRK Gram / K@W hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same role and same parity status as oracle/jaxrk_oracle.py (read its header): a
 * CPU checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The product library never links or calls this file.  It exists next to the numpy
 * restatement so that checks at sizes numpy cannot hold in memory (a 65536 x 65536
 * Gram is 17 GB; ten numpy temporaries of it are not) can be done row-block by
 * row-block, and so that the cpu_baseline has a fused multi-threaded variant.
 *
 * Reference lines followed (all under /root/reference/jaxrk):
 *   utilities/eucldist.py:10-18   sq = (a2[:,None] + b2) - (2a) @ b.T      (float32)
 *   utilities/eucldist.py:48      r  = clip(sq, 0) ** (1/2)
 *   kern/util.py:115              u  = (s * r) ** p
 *   kern/rbf.py:34,105            K  = nconst * exp(-u)
 *   reduce/lincomb.py:137-140     out = A @ K  (here as K @ W, W = A.T)
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC; no -ffast-math, so the
 * float32 op order below is what executes).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* one element in the reference's float32 op order */
static inline float ref32_elem(const float *x, const float *y, float a2, float b2, int d,
                               float s, float p, float nconst)
{
    float g = 0.0f;
    for (int k = 0; k < d; ++k) g += (2.0f * x[k]) * y[k];
    float sq = (a2 + b2) - g;
    float c = sq > 0.0f ? sq : 0.0f;
    float r = powf(c, 0.5f);
    float t = s * r;
    float u = powf(t, p);
    return nconst * expf(-u);
}

static inline double truth64_elem(const float *x, const float *y, int d, double s, double p,
                                  double nconst)
{
    double sq = 0.0;
    for (int k = 0; k < d; ++k) {
        double diff = (double)x[k] - (double)y[k];
        sq += diff * diff;
    }
    return nconst * exp(-pow(s * sqrt(sq), p));
}

static float rownorm32(const float *x, int d)
{
    float a = 0.0f;
    for (int k = 0; k < d; ++k) a += x[k] * x[k];
    return a;
}

/* Gram rows [i0, i1) of K(X, Y) into out (row-major, (i1-i0) x M), reference order. */
void orc_gram_ref32(const float *X, const float *Y, int64_t i0, int64_t i1, int64_t M, int d,
                    float s, float p, float nconst, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i) {
        const float *x = X + i * d;
        float a2 = rownorm32(x, d);
        for (int64_t j = 0; j < M; ++j) {
            const float *y = Y + j * d;
            out[(i - i0) * M + j] = ref32_elem(x, y, a2, rownorm32(y, d), d, s, p, nconst);
        }
    }
}

/* Same rows, float64 direct differencing on the float32 inputs ("truth"). */
void orc_gram_truth64(const float *X, const float *Y, int64_t i0, int64_t i1, int64_t M, int d,
                      float s, float p, float nconst, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i)
        for (int64_t j = 0; j < M; ++j)
            out[(i - i0) * M + j] = truth64_elem(X + i * d, Y + j * d, d, s, p, nconst);
}

/* Rows [i0, i1) of K(X, Y) @ W (W is M x m row-major) with float64 accumulation of the
 * float64 truth; absout (nullable) receives sum_j |K_ij| |W_jc|, the scale of the
 * float32 accumulation error bound. */
void orc_kmatw_truth64(const float *X, const float *Y, const float *W, int64_t i0, int64_t i1,
                       int64_t M, int d, int m, float s, float p, float nconst, double *out,
                       double *absout)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i) {
        double *o = out + (i - i0) * m;
        double *ao = absout ? absout + (i - i0) * m : NULL;
        for (int c = 0; c < m; ++c) { o[c] = 0.0; if (ao) ao[c] = 0.0; }
        for (int64_t j = 0; j < M; ++j) {
            double kij = truth64_elem(X + i * d, Y + j * d, d, s, p, nconst);
            for (int c = 0; c < m; ++c) {
                double w = (double)W[j * m + c];
                o[c] += kij * w;
                if (ao) ao[c] += fabs(kij * w);
            }
        }
    }
}

/* Reference order: materialise the float32 row of K, then a float32 dot per column. */
void orc_kmatw_ref32(const float *X, const float *Y, const float *W, int64_t i0, int64_t i1,
                     int64_t M, int d, int m, float s, float p, float nconst, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i) {
        const float *x = X + i * d;
        float a2 = rownorm32(x, d);
        float *o = out + (i - i0) * m;
        for (int c = 0; c < m; ++c) o[c] = 0.0f;
        for (int64_t j = 0; j < M; ++j) {
            const float *y = Y + j * d;
            float kij = ref32_elem(x, y, a2, rownorm32(y, d), d, s, p, nconst);
            for (int c = 0; c < m; ++c) o[c] += kij * W[j * m + c];
        }
    }
}

/* sum over i in [i0,i1), j in [j0,j1) of pi[i] * K(x_i, y_j) * pj[j] in float64 (pi/pj
 * nullable = 1): one entry of R_X K R_Y^T for group reduces (reduce/base.py:152-181). */
double orc_blocksum_truth64(const float *X, const float *Y, const float *pi, const float *pj,
                            int64_t i0, int64_t i1, int64_t j0, int64_t j1, int d, float s, float p,
                            float nconst)
{
    double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t i = i0; i < i1; ++i) {
        double row = 0.0;
        for (int64_t j = j0; j < j1; ++j)
            row += truth64_elem(X + i * d, Y + j * d, d, s, p, nconst) * (pj ? (double)pj[j] : 1.0);
        tot += row * (pi ? (double)pi[i] : 1.0);
    }
    return tot;
}

/* ---- sibling kernels on the same distance front-end and the diag branch (widened rows) ---- */

/* PeriodicKernel (kern/rbf.py:152-154 over ScaledPairwiseDistance(SimpleScaler(1/period), power = 1), rbf.py:120):
 * exp(-2 (sin(pi * s * |x - y|) / ls)^2), float64 on the float32 inputs; inv_ls is the float32 value 1 / ls. */
void orc_periodic_truth64(const float *X, const float *Y, int64_t N, int64_t M, int d, float s, float inv_ls,
                          double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < M; ++j) {
            double sq = 0.0;
            for (int k = 0; k < d; ++k) {
                double diff = (double)X[i * d + k] - (double)Y[j * d + k];
                sq += diff * diff;
            }
            double q = sin(3.14159265358979323846 * (double)s * sqrt(sq)) * (double)inv_ls;
            out[i * M + j] = exp(-2.0 * q * q);
        }
}

/* ThreshSpikeKernel (kern/rbf.py:203-206): where((s |x - y|)^p <= threshold, spike, non_spike); margin (nullable)
 * receives |dist - threshold| so that undecidable entries can be excluded from parity. */
void orc_threshspike_truth64(const float *X, const float *Y, int64_t N, int64_t M, int d, float s, float p,
                             float spike, float non_spike, float threshold, double *out, double *margin)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < M; ++j) {
            double sq = 0.0;
            for (int k = 0; k < d; ++k) {
                double diff = (double)X[i * d + k] - (double)Y[j * d + k];
                sq += diff * diff;
            }
            double dist = pow((double)s * sqrt(sq), (double)p);
            out[i * M + j] = dist <= (double)threshold ? (double)spike : (double)non_spike;
            if (margin) margin[i * M + j] = fabs(dist - (double)threshold);
        }
}

/* diag = True branch of ScaledPairwiseDistance.__call__ (kern/util.py:112-113): gs * sum_k |x_ik - y_ik|^p. */
void orc_dist_diag_truth64(const float *X, const float *Y, int64_t N, int d, float s, float p, double *out)
{
    for (int64_t i = 0; i < N; ++i) {
        double acc = 0.0;
        for (int k = 0; k < d; ++k) acc += pow(fabs((double)X[i * d + k] - (double)Y[i * d + k]), (double)p);
        out[i] = (double)s * acc;
    }
}

/* Vector-Jacobian product of K = nconst exp(-(s |x - y|)^p) with cotangent A (N x M), what jax.grad derives from
 * kern/rbf.py:104-105 over kern/util.py:115:  gX[i] = -sum_j A_ij K_ij p s^p r^(p-2) (x_i - y_j),
 * gs = -sum_ij A_ij K_ij p s^(p-1) r^p; pairs at r = 0 contribute nothing.  gX is N x d, gs one double. */
void orc_vjp_truth64(const float *X, const float *Y, const double *A, int64_t N, int64_t M, int d, float s, float p,
                     float nconst, double *gX, double *gs)
{
    double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t i = 0; i < N; ++i) {
        for (int k = 0; k < d; ++k) gX[i * d + k] = 0.0;
        for (int64_t j = 0; j < M; ++j) {
            double sq = 0.0;
            for (int k = 0; k < d; ++k) {
                double diff = (double)X[i * d + k] - (double)Y[j * d + k];
                sq += diff * diff;
            }
            if (sq <= 0.0) continue;
            double r = sqrt(sq), K = (double)nconst * exp(-pow((double)s * r, (double)p));
            double phi = K * (double)p * pow((double)s, (double)p) * pow(r, (double)p - 2.0);
            double a = A[i * M + j];
            for (int k = 0; k < d; ++k) gX[i * d + k] -= a * phi * ((double)X[i * d + k] - (double)Y[j * d + k]);
            tot -= a * K * (double)p * pow((double)s, (double)p - 1.0) * pow(r, (double)p);
        }
    }
    *gs = tot;
}

int orc_abi_version(void) { return 2; }

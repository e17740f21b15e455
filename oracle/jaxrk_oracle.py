"""CPU oracle for the JaxRK Gram / fused K@W hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the arithmetic the reference performs on the
path named by BASELINE.json `north_star`.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product package `jaxrk_b200` never does (it fails loudly without its CUDA library).

Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
  G1  the printed float32 outputs of the authors' own JAX-CPU run,
      /root/reference/examples/Quick_start.ipynb:166  (inputs :53-66),
  G2  the scipy known answers of /root/reference/tests/test_kern.py:15-43,
  G3  scipy cdist as in /root/reference/tests/test_utilities.py:11-31,
  G5  outputs of the *reference's own Python code* executed over a numpy-backed
      stand-in for the absent jax package (tests/golden/make_golden.py; fixtures
      committed under tests/golden/).
The arithmetic itself lives in the third-party `jax`/`jaxlib` (XLA) dependency, which
is unpinned by the reference (setup.py:13) and absent from this image, so XLA's own
exp/pow/GEMM rounding cannot be reproduced bit for bit; the op ORDER is.

Two flavours are provided for every function:
  *_ref32   float32, the reference's op order (expansion-form distance, clip, pow,
            scale, pow, neg, exp, times nconst)              -> "what JaxRK returns"
  *_truth64 float64 direct differencing on the float32-rounded inputs/parameters
                                                            -> "what it should return"
The tolerance policy is stated against truth64; ref32 is required to
agree only where it is itself within tolerance of truth64 (it is not, by up to 1.2e-4,
on self-Gram diagonals with power < 2, because sqrt(clip(+-4e-6)) != 0).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
from scipy.special import gammaln as _gammaln

f32 = np.float32

# The literal the reference hard-codes in GenGaussKernel.make_gauss (kern/rbf.py:93).
# It is NOT 1/sqrt(2) (differs by 2.4e-7 relative) and must be reproduced.
GAUSS_F = 0.70710695


def _as2d_f32(a) -> np.ndarray:
    """JAX without x64 truncates every input to float32 (tests pass int32 and float64)."""
    a = np.asarray(a)
    assert a.ndim == 2
    return np.ascontiguousarray(a, dtype=f32)


# ---------------------------------------------------------------------------------
# jaxrk/utilities/eucldist.py
# ---------------------------------------------------------------------------------
def sqeucldist_simple_ref32(a, b=None) -> np.ndarray:
    """eucldist.py:10-18.  (a2[:,None] + b2) - ((2*a) @ b.T), all float32."""
    a = _as2d_f32(a)
    a2 = np.einsum("ij,ij->i", a, a, dtype=f32)
    if b is not None:
        b = _as2d_f32(b)
        b2 = np.einsum("ij,ij->i", b, b, dtype=f32)
    else:
        b, b2 = a, a2
    return (a2[:, None] + b2) - (f32(2) * a) @ b.T


def sqeucldist_extension_ref32(a, b=None) -> np.ndarray:
    """eucldist.py:20-35.  One GEMM of inner dimension 3d: [1, a, a^2] @ [b^2; -2b; 1]."""
    a = _as2d_f32(a)
    a_sq = a ** 2
    if b is not None:
        b = _as2d_f32(b)
        b_sq = b ** 2
    else:
        b, b_sq = a, a_sq
    n_a, dim = a.shape
    n_b = b.shape[0]
    a_ext = np.hstack([np.ones((n_a, dim), f32), a, a_sq])
    b_ext = np.vstack([b_sq.T, f32(-2.0) * b.T, np.ones((dim, n_b), f32)])
    return a_ext @ b_ext


def eucldist_ref32(a, b=None, power=1.0, variant="simple") -> np.ndarray:
    """eucldist.py:37-48.  power == 2 returns the raw (possibly negative) expansion."""
    if variant == "simple":
        sq = sqeucldist_simple_ref32(a, b)
    elif variant == "extension":
        sq = sqeucldist_extension_ref32(a, b)
    else:
        raise AssertionError()
    if power == 2:
        return sq
    return np.power(np.clip(sq, f32(0), None), f32(power / 2.0))


def sqeucldist_truth64(a, b=None) -> np.ndarray:
    """float64 direct differencing of the float32-rounded inputs."""
    a = _as2d_f32(a).astype(np.float64)
    b = a if b is None else _as2d_f32(b).astype(np.float64)
    out = np.zeros((a.shape[0], b.shape[0]))
    for k in range(a.shape[1]):  # O(N*M) memory per dimension, not O(N*M*d)
        diff = a[:, k][:, None] - b[:, k][None, :]
        out += diff * diff
    return out


# ---------------------------------------------------------------------------------
# jaxrk/kern/util.py  (SimpleScaler, ScaledPairwiseDistance) + jaxrk/kern/rbf.py
# ---------------------------------------------------------------------------------
def simple_scaler_s(length_scale) -> np.ndarray:
    """The array SimpleScaler holds for `SimpleScaler(1./length_scale)` (kern/rbf.py:72,
    kern/util.py:34-50): Python floats divide in float64 and are then stored as a
    float32 (1,1) array; array length-scales divide in float32.  Returns shape (1, k)."""
    if isinstance(length_scale, float):
        s = np.array([[1.0 / length_scale]], dtype=f32)
    else:
        ls = np.asarray(length_scale, dtype=f32)
        s = np.atleast_1d(f32(1) / ls).astype(f32)
        if s.ndim == 1:
            s = s[None, :]
        assert s.ndim == 2 and s.shape[0] == 1
    assert np.all(s > 0)
    return s


def gauss_length_scale(length_scale):
    """make_gauss: `length_scale / f` with the literal f (kern/rbf.py:84-95)."""
    if isinstance(length_scale, float):
        return length_scale / GAUSS_F
    return (np.asarray(length_scale, dtype=f32) / f32(GAUSS_F)).astype(f32)


def gengauss_nconst_ref32(s: np.ndarray, power: float) -> np.ndarray:
    """kern/rbf.py:34 with kern/util.py:101-105:
    nconst = power / (2 * scale_param * exp(gammaln(1/power))), float32;
    scale_param = 1/s (global scale) or (1/s)**(1/power) (per-dimension, quirk Q2)."""
    s = np.asarray(s, dtype=f32)
    inv = (f32(1) / s).astype(f32)
    if s.size != 1:
        inv = np.power(inv, f32(1.0 / power)).astype(f32)
    g = f32(_gammaln(f32(1.0 / power)))
    return (f32(power) / (f32(2) * inv * np.exp(g, dtype=f32))).astype(f32)


def gengauss_var_ref32(s: np.ndarray, power: float) -> np.ndarray:
    """kern/rbf.py:100-102."""
    s = np.asarray(s, dtype=f32)
    inv = (f32(1) / s).astype(f32)
    if s.size != 1:
        inv = np.power(inv, f32(1.0 / power)).astype(f32)
    f = _gammaln(np.array([3, 1], dtype=f32) / f32(power)).astype(f32)
    return (inv ** 2 * np.exp(f[0] - f[1], dtype=f32)).astype(f32)


def scaled_dist_ref32(X, Y, s: np.ndarray, power: float, diag: bool = False) -> np.ndarray:
    """ScaledPairwiseDistance.__call__ (kern/util.py:107-116).
    full:  (gs * eucldist(ds*X, ds*Y, power=1.)) ** p       (util.py:115)
    diag:  zeros(N) if Y is None else gs * sum(|ds*X - ds*Y| ** p, 1)   (util.py:108-113)"""
    s = np.asarray(s, dtype=f32)
    is_global = s.size == 1
    X = _as2d_f32(X)
    Yf = None if Y is None else _as2d_f32(Y)
    ds = (lambda v: v) if is_global else (lambda v: None if v is None else (s * v).astype(f32))
    gs = (lambda v: (s.reshape(()) * v).astype(f32)) if is_global else (lambda v: v)
    if diag:
        if Yf is None:
            return np.zeros(X.shape[0], f32)
        assert X.shape == Yf.shape
        return gs(np.sum(np.abs(ds(X) - ds(Yf)) ** f32(power), 1, dtype=f32))
    d1 = eucldist_ref32(ds(X), ds(Yf), power=1.0)
    return np.power(gs(d1), f32(power))


def gengauss_gram_ref32(X, Y, s: np.ndarray, power: float, diag: bool = False) -> np.ndarray:
    """GenGaussKernel.__call__ (kern/rbf.py:104-105): nconst * exp(-dist(X, Y, diag))."""
    nconst = gengauss_nconst_ref32(s, power)
    d = scaled_dist_ref32(X, Y, s, power, diag)
    if diag:
        return (nconst.reshape(-1) * np.exp(-d, dtype=f32)).astype(f32)
    return (nconst * np.exp(-d, dtype=f32)).astype(f32)


def scaled_dist_truth64(X, Y, s, power: float) -> np.ndarray:
    """float64 (s |x - y|)^p by direct differencing on the float32 inputs / parameters (global or per-dim s)."""
    s = np.asarray(s, dtype=f32)
    if s.size == 1:
        r = float(s.reshape(())) * np.sqrt(sqeucldist_truth64(X, Y))
    else:
        Xs = (s * _as2d_f32(X)).astype(f32)
        Ys = None if Y is None else (s * _as2d_f32(Y)).astype(f32)
        r = np.sqrt(sqeucldist_truth64(Xs, Ys))
    return np.power(r, float(f32(power)))


def scaled_dist_diag_truth64(X, Y, s, power: float) -> np.ndarray:
    """float64 restatement of the diag branch (kern/util.py:112-113): gs * sum_k |ds x - ds y|^p."""
    s = np.asarray(s, dtype=f32)
    X, Y = _as2d_f32(X), _as2d_f32(Y)
    if s.size == 1:
        return float(s.reshape(())) * np.sum(np.abs(X.astype(np.float64) - Y.astype(np.float64)) ** float(f32(power)), 1)
    df = (s * X).astype(f32).astype(np.float64) - (s * Y).astype(f32).astype(np.float64)
    return np.sum(np.abs(df) ** float(f32(power)), 1)


def periodic_gram_ref32(X, Y, period: float, length_scale: float, diag: bool = False) -> np.ndarray:
    """PeriodicKernel.__call__ (kern/rbf.py:152-154) over ScaledPairwiseDistance(SimpleScaler(1/period), power=1)
    (rbf.py:120): d = dist(X, Y, diag); exp(-2 * (sin(pi * d) / ls) ** 2), float32 op by op."""
    s = simple_scaler_s(period)
    d = scaled_dist_ref32(X, Y, s, 1.0, diag)
    sn = np.sin((f32(np.pi) * d).astype(f32), dtype=f32)
    return np.exp(f32(-2.) * (sn / f32(length_scale)) ** f32(2.), dtype=f32).astype(f32)


def periodic_gram_truth64(X, Y, period: float, length_scale: float) -> np.ndarray:
    """float64 truth; the scale and the length scale are the float32 values the device kernel receives."""
    s = simple_scaler_s(period)
    d = scaled_dist_truth64(X, Y, s, 1.0)
    return np.exp(-2. * (np.sin(np.pi * d) * float(f32(1.) / f32(length_scale))) ** 2)


def periodic_gram_cond64(X, Y, period: float, length_scale: float) -> np.ndarray:
    """|dK / dd| * d: how far one relative float32 rounding of the distance (unavoidable: d is a float32 quantity in
    the reference and on the device) moves the kernel value.  Parity tolerance for PeriodicKernel is
    atol + rtol |K| + 4 eps32 cond  (sin(pi d) is ill-conditioned at many periods' distance)."""
    s = simple_scaler_s(period)
    d = scaled_dist_truth64(X, Y, s, 1.0)
    K = periodic_gram_truth64(X, Y, period, length_scale)
    il2 = float(f32(1.) / f32(length_scale)) ** 2
    return np.abs(K * 4. * np.sin(np.pi * d) * np.cos(np.pi * d) * np.pi * il2) * d


def threshspike_gram_ref32(X, Y, s, power: float, spike: float, non_spike: float, threshold: float) -> np.ndarray:
    """ThreshSpikeKernel.__call__ (kern/rbf.py:203-206): where(dist(X, Y) <= threshold, spike, non_spike)."""
    d = scaled_dist_ref32(X, Y, s, power, False)
    return np.where(d <= f32(threshold), f32(spike), f32(non_spike)).astype(f32)


def threshspike_gram_truth64(X, Y, s, power: float, spike: float, non_spike: float, threshold: float):
    """(values, margin): float64 restatement and |dist - threshold| -- entries whose margin is below the float32
    resolution of the distance are undecidable and excluded from parity."""
    d = scaled_dist_truth64(X, Y, s, power)
    thr = float(f32(threshold))
    return np.where(d <= thr, float(f32(spike)), float(f32(non_spike))), np.abs(d - thr)


def gengauss_gram_truth64(X, Y, s, power: float, nconst=None) -> np.ndarray:
    """float64 truth for a GLOBAL scale: nconst * exp(-(s*|x-y|)^p) with direct
    differencing; s, nconst are the float32 values the device kernel receives."""
    s = np.asarray(s, dtype=f32)
    if nconst is None:
        nconst = gengauss_nconst_ref32(s, power)
    if s.size == 1:
        sq = sqeucldist_truth64(X, Y)
        r = float(s.reshape(())) * np.sqrt(sq)
    else:  # per-dimension scale applies to the inputs (util.py:92-99)
        Xs = (s * _as2d_f32(X)).astype(f32)
        Ys = None if Y is None else (s * _as2d_f32(Y)).astype(f32)
        r = np.sqrt(sqeucldist_truth64(Xs, Ys))
    nc = np.asarray(nconst, dtype=np.float64)
    nc = float(nc.reshape(-1)[0]) if nc.size == 1 else nc
    # the power reaches the device (and the reference's float32 `**`) rounded to float32
    return nc * np.exp(-np.power(r, float(f32(power))))


# ---------------------------------------------------------------------------------
# jaxrk/reduce/base.py + jaxrk/reduce/lincomb.py -- the linear maps, as data.
# A chain is a list of tuples; each acts on axis 0 of a 2-D array:
#   ("prefactors", p)            reduce/base.py:84-101
#   ("sum",) / ("mean",)         reduce/base.py:132-150   (keepdims)
#   ("balanced", k, average)     reduce/base.py:152-181
#   ("center",)                  reduce/base.py:183-192
#   ("tileview", result_len)     reduce/base.py:118-129
#   ("linear", A)                reduce/lincomb.py:131-144  A @ inp
#   ("sparse", idcs, average)    reduce/lincomb.py:27-62
# ---------------------------------------------------------------------------------
Chain = Sequence[tuple]


def reduce_first_ax(r: tuple, inp: np.ndarray) -> np.ndarray:
    kind = r[0]
    if kind == "prefactors":
        p = np.asarray(r[1])
        assert p.ndim == 1 and p.shape[0] == inp.shape[0]
        return inp.astype(p.dtype) * p[:, None]
    if kind == "sum":
        return inp.sum(axis=0, keepdims=True)
    if kind == "mean":
        return np.mean(inp, axis=0, keepdims=True)
    if kind == "balanced":
        k, average = r[1], r[2]
        red = np.mean if average else np.sum
        return red(np.reshape(inp, (-1, k, inp.shape[-1])), axis=1)
    if kind == "center":
        return inp - np.mean(inp, 0, keepdims=True)
    if kind == "tileview":
        assert r[1] % inp.shape[0] == 0
        return np.tile(inp, (r[1] // inp.shape[0], 1))
    if kind == "linear":
        A = np.asarray(r[1])
        assert inp.ndim == 2 and A.shape[1] == inp.shape[0]
        return A @ inp
    if kind == "sparse":
        idcs, average = r[1], r[2]
        red = np.mean if average else np.sum
        out = []
        for idx in idcs:
            idx = np.asarray(idx)
            if idx.shape[1] == 0:
                out.append(np.zeros((idx.shape[0], inp.shape[1]), inp.dtype))
            else:
                out.append(red(inp[idx.flatten(), :].reshape((-1, idx.shape[1], inp.shape[1])), 1))
        return np.concatenate(out, 0)
    raise ValueError(kind)


def new_len(r: tuple, n: int) -> int:
    kind = r[0]
    if kind in ("prefactors", "center"):
        return n
    if kind in ("sum", "mean"):
        return 1
    if kind == "balanced":
        assert n % r[1] == 0
        return n // r[1]
    if kind == "tileview":
        return r[1]
    if kind == "linear":
        assert np.asarray(r[1]).shape[1] == n
        return np.asarray(r[1]).shape[0]
    if kind == "sparse":
        return len(r[1])
    raise ValueError(kind)


def reduce_apply(inp: np.ndarray, chain: Optional[Chain], axis: int = 0) -> np.ndarray:
    """Reduce.apply (reduce/base.py:39-47)."""
    if chain is None or len(chain) == 0:
        return inp
    carry = np.swapaxes(inp, axis, 0)
    for r in chain:
        carry = reduce_first_ax(r, carry)
    return np.swapaxes(carry, axis, 0)


def final_len(n: int, chain: Optional[Chain]) -> int:
    """Reduce.final_len (reduce/base.py:49-55)."""
    for r in chain or []:
        n = new_len(r, n)
    return n


# ---------------------------------------------------------------------------------
# jaxrk/rkhs/vector.py + jaxrk/rkhs/operator.py
# ---------------------------------------------------------------------------------
def finitevec_inner_ref32(X, chain_x: Chain, Y, chain_y: Chain, s, power) -> np.ndarray:
    """FiniteVec.inner (rkhs/vector.py:62-75): reduce_Y(reduce_X(K(X, Y)), axis=1).
    `.astype(float)` is a no-op to float32 without x64 (quirk Q4)."""
    gram = gengauss_gram_ref32(X, Y, s, power)
    r1 = reduce_apply(gram, chain_x, 0)
    return reduce_apply(r1, chain_y, 1)


def finitevec_inner_truth64(X, chain_x: Chain, Y, chain_y: Chain, s, power) -> np.ndarray:
    gram = gengauss_gram_truth64(X, Y, s, power)

    def up(chain):
        out = []
        for r in chain or []:
            if r[0] in ("prefactors", "linear"):
                out.append((r[0], np.asarray(r[1], dtype=np.float64)))
            else:
                out.append(r)
        return out

    return reduce_apply(reduce_apply(gram, up(chain_x), 0), up(chain_y), 1)


def kmatw_ref32(X, Y, W, s, power) -> np.ndarray:
    """The K(X,Y) @ W that FiniteVec.inner reduces to when Y.reduce == [LinearReduce(W.T)]
    (rkhs/vector.py:72-74 with reduce/lincomb.py:137-140): materialise, then GEMM."""
    return gengauss_gram_ref32(X, Y, s, power) @ np.asarray(W, dtype=f32)


def kmatw_truth64(X, Y, W, s, power, nconst=None) -> np.ndarray:
    return gengauss_gram_truth64(X, Y, s, power, nconst) @ np.asarray(W, dtype=f32).astype(np.float64)


def kmatw_absbound64(X, Y, W, s, power, nconst=None) -> np.ndarray:
    """sum_j |K_ij| |W_jc| -- the scale against which a float32 accumulation of
    K@W is judged (SURVEY.md §8c tolerance policy)."""
    return gengauss_gram_truth64(X, Y, s, power, nconst) @ np.abs(np.asarray(W, dtype=f32).astype(np.float64))


def finiteop_matmul_vec_ref32(X_in, chain_in, matr, X_right, chain_right, s, power,
                              normalize=False) -> np.ndarray:
    """FiniteOp.__matmul__, vector branch (rkhs/operator.py:42-56): returns the
    linear map `lin_LinOp` (len(outp_feat) x len(right_inp)) whose transpose becomes
    the LinearReduce appended to outp_feat."""
    lin = finitevec_inner_ref32(X_in, chain_in, X_right, chain_right, s, power)
    if matr is not None:
        lin = np.asarray(matr, dtype=f32) @ lin
    if normalize:
        lin = lin / lin.sum(1, keepdims=True)
    return lin


# ---------------------------------------------------------------------------------
# tolerance policy helpers
# ---------------------------------------------------------------------------------
RTOL = 1e-5
ATOL = 1e-6


def gengauss_vjp_truth64(X, Y, A, s, power: float, nconst=None):
    """float64 vector-Jacobian product of K = nconst exp(-(s |x - y|)^p) with cotangent A = dL/dK (N x M):
    returns (dL/dX, dL/dY, dL/ds at fixed nconst, sum_j |a_ij dK_ij/dx_i| as the accumulation magnitude for tolerances).
    This is what jax.grad derives from the reference chain (kern/rbf.py:104-105 over kern/util.py:115), written out:
    dK/dx_i = -K p s^p r^(p-2) (x_i - y_j),  dK/ds = -K p s^(p-1) r^p;  pairs at r = 0 contribute nothing."""
    s = np.asarray(s, dtype=f32)
    assert s.size == 1
    sv = float(s.reshape(()))
    if nconst is None:
        nconst = gengauss_nconst_ref32(s, power)
    nc = float(np.asarray(nconst, dtype=np.float64).reshape(-1)[0])
    p = float(f32(power))
    X64, Y64 = _as2d_f32(X).astype(np.float64), _as2d_f32(Y).astype(np.float64)
    A = np.asarray(A, dtype=np.float64)
    diff = X64[:, None, :] - Y64[None, :, :]                    # N x M x d
    sq = (diff ** 2).sum(-1)
    r = np.sqrt(sq)
    K = nc * np.exp(-np.power(sv * r, p))
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = np.where(sq > 0, K * p * sv ** p * np.power(r, p - 2.0), 0.0)
    T = A * phi
    gX = -(T[:, :, None] * diff).sum(1)
    gY = (T[:, :, None] * diff).sum(0)
    gs = -(A * K * p * sv ** (p - 1.0) * np.power(r, p)).sum()
    mag_x = (np.abs(T)[:, :, None] * np.abs(diff)).sum(1)
    mag_y = (np.abs(T)[:, :, None] * np.abs(diff)).sum(0)
    mag_s = (np.abs(A) * K * p * sv ** (p - 1.0) * np.power(r, p)).sum()
    return gX, gY, gs, (mag_x, mag_y, mag_s)


def within_tol(got, truth, rtol=RTOL, atol=ATOL) -> np.ndarray:
    got = np.asarray(got, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    return np.abs(got - truth) <= atol + rtol * np.abs(truth)

"""Tensor-core K@W path (kmatw_tc.cu: both contractions on tcgen05) against the oracle and against the FP32-pipe
kernel.  The path is taken automatically for d <= 32, m <= 16, N >= 8192, M >= 4096 with no row reduce."""
import os

import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from oracle import jaxrk_oracle as orc
from gpu_util import assert_kmatw_close, cu, f32, gauss, kparams

pytestmark = pytest.mark.gpu


@pytest.fixture
def force():
    prev = nat.get_option(nat.OPT_KMATW_TC)

    def _set(on):
        nat.set_option(nat.OPT_KMATW_TC, 1 if on else 0)
    yield _set
    nat.set_option(nat.OPT_KMATW_TC, prev)


def _sample_truth(X, Y, W, s, p, rows, rp=None, cp=None):
    G = orc.gengauss_gram_truth64(X[rows], Y, s, p)
    if cp is not None:
        G = G * cp.astype(np.float64)[None, :]
    if rp is not None:
        G = G * rp[rows].astype(np.float64)[:, None]
    W64 = W.astype(np.float64)
    return G @ W64, np.abs(G) @ np.abs(W64)


@pytest.mark.parametrize("kname", ["sqexp", "laplace", "gg15", "gg05"])
@pytest.mark.parametrize("d,m", [(16, 16), (7, 5), (32, 1), (1, 3), (20, 16)])
def test_tc_path_matches_oracle_and_fp32_kernel(kname, d, m, force):
    N, M = 8192 + 77, 4096 + 33        # ragged in both directions
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    s, scale, p, nconst = kparams(kname, np.sqrt(d) * 1.2)
    force(True)
    n0 = nat.launch_count()
    T = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    # X norms | Y norms, first-generation pack, max|W|, [sqexp: second-generation pack] | main kernel(s)
    assert nat.launch_count() - n0 == (7 if kname == "sqexp" else 5), "pre-pass + tensor-core kernel(s)"
    force(False)
    F = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    rows = np.r_[0:64, N - 64:N]
    truth, bound = _sample_truth(X, Y, W, s, p, rows)
    assert_kmatw_close(T.cpu().numpy()[rows], truth, bound, f"tc {kname} d={d} m={m}")
    assert_kmatw_close(F.cpu().numpy()[rows], truth, bound, f"fp32 {kname} d={d} m={m}")
    b = torch.from_numpy(bound.max(0)).to(T.device).float()
    assert ((T - F).abs() <= 2e-6 + 2e-5 * b).all()


def test_tc_path_prefactors_strides_and_positive_weights(force):
    N, M, d, m = 16384, 16384, 16, 8
    Xb, Yb, Wb = cu(gauss(3, N, 24)), cu(gauss(4, M, 20)), cu(np.abs(gauss(5, M, 12)))
    X, Y, W = Xb[:, 3:3 + d], Yb[:, 1:1 + d], Wb[:, 2:2 + m]          # strided, unaligned views
    rp = np.random.RandomState(6).rand(N).astype(f32)
    cp = np.random.RandomState(7).rand(M).astype(f32)
    s, scale, p, nconst = kparams("sqexp", 4.0)
    force(True)
    T = nat.kmatw(X, Y, W, scale, p, nconst, row_pref=rp, col_pref=cp)
    rows = np.arange(0, N, 257)
    truth, bound = _sample_truth(X.cpu().numpy(), Y.cpu().numpy(), W.cpu().numpy(), s, p, rows, rp, cp)
    assert_kmatw_close(T.cpu().numpy()[rows], truth, bound, "all-positive sum through the tensor core")


def test_tc_path_near_duplicates_are_refined(force):
    rng = np.random.RandomState(9)
    N, M, d, m = 8192, 8192, 16, 4
    X = (rng.randn(N, d) * 3).astype(f32)
    Y = (X[rng.permutation(N)][:M] + rng.randn(M, d).astype(f32) * 1e-4).astype(f32)   # every y_j sits next to some x_i
    W = gauss(2, M, m)
    for kname in ("laplace", "gg05", "sqexp"):
        s, scale, p, nconst = kparams(kname, 0.5)    # short length scale: the near pairs dominate the sums
        force(True)
        T = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
        rows = np.arange(0, N, 131)
        truth, bound = _sample_truth(X, Y, W, s, p, rows)
        assert_kmatw_close(T.cpu().numpy()[rows], truth, bound, kname)


def test_tc_path_through_the_api_cfg5_apply_shape(force):
    """operator apply: K(X_test, X) @ W -- the inner() with a LinearReduce on the Y side picks the tensor-core path."""
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import LinearReduce
    from jaxrk_b200.rkhs import FiniteVec, inner
    rng = np.random.RandomState(0)
    T_, N, d, m = 16384, 8192, 16, 16
    Xt, X, A = rng.randn(T_, d).astype(f32), rng.randn(N, d).astype(f32), rng.randn(m, N).astype(f32)
    k = GenGaussKernel.make_gauss(4.0)
    force(True)
    out = inner(FiniteVec(k, Xt), FiniteVec(k, X, [LinearReduce(A)]))
    force(False)
    ref = inner(FiniteVec(k, Xt), FiniteVec(k, X, [LinearReduce(A)]))
    assert tuple(out.shape) == (T_, m)
    assert torch.allclose(out, ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("ls_factor,expect", [(1.0, "fast"), (0.75, "fast"), (0.6, "fast-edge"), (0.4, "generic")])
def test_tc_factored_sqexp_path_regimes(ls_factor, expect, force):
    """p = 2: below |c| (max|x|^2 + max|y|^2) = 12 the kernel factors K_ij = 2^(c|x|^2) 2^(-2c x.y) 2^(c|y|^2)
    (kmatw_tc.cu: TC_FAST_BOUND = 12); above it the generic epilogue runs.  Both must agree with the oracle and with the
    generic epilogue forced by JRK_OPT_TC_FAST = 0 and the first-generation kernel (JRK_OPT_TC_V2 = 0)."""
    N, M, d, m = 8192 + 5, 8192 + 64 + 3, 16, 16
    X, Y, W = gauss(0, N, d), gauss(1, M, d), np.abs(gauss(2, M, m))      # all-positive sums: the hard case
    s, scale, p, nconst = kparams("sqexp", np.sqrt(d) * ls_factor)
    c = float(scale) ** 2 * 1.4426950408889634
    bound_val = c * ((X.astype(np.float64) ** 2).sum(1).max() + (Y.astype(np.float64) ** 2).sum(1).max())
    if expect == "generic":
        assert bound_val > 12.0
    else:
        assert bound_val <= 12.0 and (expect == "fast" or bound_val > 10.0)
    force(True)
    T = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    with nat.option(nat.OPT_TC_FAST, 0):
        G = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    with nat.option(nat.OPT_TC_V2, 0):
        V1 = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    rows = np.r_[0:96, N - 96:N]
    truth, bound = _sample_truth(X, Y, W, s, p, rows)
    assert_kmatw_close(T.cpu().numpy()[rows], truth, bound, f"tc sqexp {expect}")
    assert_kmatw_close(G.cpu().numpy()[rows], truth, bound, f"tc sqexp {expect}, generic epilogue")
    assert_kmatw_close(V1.cpu().numpy()[rows], truth, bound, f"tc sqexp {expect}, first-generation kernel")
    b = torch.from_numpy(bound.max(0)).to(T.device).float()
    assert ((T - G).abs() <= 2e-6 + 2e-5 * b).all()


@pytest.mark.parametrize("kname", ["sqexp", "laplace"])
@pytest.mark.parametrize("rr", [nat.REDUCE_SUM, nat.REDUCE_MEAN])
def test_tc_path_with_sum_and_mean_over_rows(kname, rr, force):
    """Sum / Mean (jaxrk/reduce/base.py:132-150) folded behind the tensor-core kernel: the mean embedding evaluated at
    M points, 1 x m, reduced in a fixed order."""
    N, M, d, m = 8192 + 300, 4096 + 17, 16, 7
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    rp = np.random.RandomState(3).rand(N).astype(f32)
    s, scale, p, nconst = kparams(kname, np.sqrt(d))
    assert nat.lib().jrk_kmatw_uses_tensor_cores(N, M, d, m, 0, rr) == 1
    force(True)
    T = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, row_pref=rp, row_reduce=rr)
    T2 = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, row_pref=rp, row_reduce=rr)
    force(False)
    F = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, row_pref=rp, row_reduce=rr)
    assert tuple(T.shape) == (1, m) and torch.equal(T, T2)          # deterministic
    rows = np.arange(N)
    G = orc.gengauss_gram_truth64(X, Y, s, p) * rp.astype(np.float64)[:, None]
    truth = (G @ W.astype(np.float64)).sum(0, keepdims=True)
    bound = (np.abs(G) @ np.abs(W.astype(np.float64))).sum(0, keepdims=True)
    if rr == nat.REDUCE_MEAN:
        truth, bound = truth / N, bound / N
    assert_kmatw_close(T.cpu().numpy(), truth, bound, f"tc {kname} reduce {rr}")
    assert_kmatw_close(F.cpu().numpy(), truth, bound, f"fp32 {kname} reduce {rr}")


@pytest.mark.parametrize("M", [4288, 4109, 6000, 8192 + 64 * 3])
def test_tc2_every_row_with_several_units_per_cta(M):
    """More 128-row units than CTAs and tile counts with every residue modulo the warps per lane quarter: EVERY row of the
    tensor-core forward against the FP32-pipe kernel (tolerance from the accumulated magnitude K @ |W|), and the same call
    twice bit for bit."""
    N, d, m = 50000, 16, 16
    g = torch.Generator(device="cuda").manual_seed(M)
    X, Y, W = (torch.randn(n_, c_, generator=g, device="cuda") for n_, c_ in ((N, d), (M, d), (M, m)))
    scale, nconst = 0.70710695 / 4, 0.1
    assert nat.lib().jrk_kmatw_uses_tensor_cores(N, M, d, m, 0, 0) == 1
    t1 = nat.kmatw(X, Y, W, scale, 2.0, nconst)
    t2 = nat.kmatw(X, Y, W, scale, 2.0, nconst)
    assert torch.equal(t1, t2)
    with nat.option(nat.OPT_KMATW_TC, 0):
        t0 = nat.kmatw(X, Y, W, scale, 2.0, nconst)
        bound = nat.kmatw(X, Y, W.abs(), scale, 2.0, nconst)
    assert bool(((t1 - t0).abs() <= 2e-6 + 3e-5 * bound).all())


@pytest.mark.parametrize("N,M,d,m,ls,positive", [(16384, 8192, 64, 16, 8.0, False), (50000, 4109, 48, 16, 7.0, True),
                                                 (16384 + 77, 6000, 33, 5, 6.0, False), (20000, 8192, 64, 24, 8.0, True)])
def test_tc2_d_up_to_64(N, M, d, m, ls, positive):
    """32 < d <= 64 (kmatw_tc2_kernel<4, 4>: four GEMM1 k-steps, 20 KB tile images, five S / e stages): sampled rows against
    float64, EVERY row against the FP32-pipe kernel (several units per CTA for N = 50000), bit-identical repeats."""
    g = torch.Generator(device="cuda").manual_seed(N + d)
    X, Y, W = (torch.randn(n_, c_, generator=g, device="cuda") for n_, c_ in ((N, d), (M, d), (M, m)))
    if positive:
        W = W.abs()
    s, nc = float(0.70710695 / ls), 0.25
    n0 = nat.launch_count()
    A = nat.kmatw(X, Y, W, s, 2.0, nc)
    blocks = -(-m // 16)
    assert nat.launch_count() - n0 >= 6 * blocks          # per column block: 2 norm kernels, W maximum, images, tensor-core
    assert torch.equal(A, nat.kmatw(X, Y, W, s, 2.0, nc))                               # kernel, FP32-pipe kernel (skipped)
    with nat.option(nat.OPT_KMATW_TC, 0):
        R = nat.kmatw(X, Y, W, s, 2.0, nc)
        bound_all = nat.kmatw(X, Y, W.abs(), s, 2.0, nc)
    assert bool(((A - R).abs() <= 2e-6 + 3e-5 * bound_all).all())
    rows = torch.cat([torch.arange(0, 40, device="cuda"), torch.arange(40, N, (N - 40) // 88, device="cuda")[:88]])
    Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** 2)
    T, bound = Kt @ W.double(), Kt @ W.double().abs()
    assert bool(((A[rows].double() - T).abs() <= 1e-6 + 1e-5 * bound).all())


def test_tc2_d64_outside_the_factored_regime_is_the_fp32_kernel():
    g = torch.Generator(device="cuda").manual_seed(9)
    X, Y, W = (torch.randn(n_, c_, generator=g, device="cuda") for n_, c_ in ((16384, 64), (8192, 64), (8192, 16)))
    s = float(0.70710695 / 3.0)                     # |c| (max|x|^2 + max|y|^2) ~ 19 > 12
    A = nat.kmatw(X, Y, W, s, 2.0, 0.25)
    with nat.option(nat.OPT_KMATW_TC, 0):
        R = nat.kmatw(X, Y, W, s, 2.0, 0.25)
    assert torch.equal(A, R)


@pytest.mark.parametrize("rr", ["sum", "mean"])
@pytest.mark.parametrize("ls", [8.0, 3.0])
def test_tc2_d64_row_reduce(rr, ls):
    """Sum / Mean over rows folded into the call at d = 64 (a mean embedding against an RKHS element): inside the factored
    regime the tensor-core kernel + fixed-order row sums, outside it (ls = 3) the FP32-pipe kernel -- against float64."""
    N, M, d, m = 20000, 8192, 64, 20
    g = torch.Generator(device="cuda").manual_seed(17)
    X, Y, W = (torch.randn(n_, c_, generator=g, device="cuda") for n_, c_ in ((N, d), (M, d), (M, m)))
    s, nc = float(0.70710695 / ls), 0.25
    code = nat.REDUCE_SUM if rr == "sum" else nat.REDUCE_MEAN
    got = nat.kmatw(X, Y, W, s, 2.0, nc, row_reduce=code)
    assert tuple(got.shape) == (1, m)
    assert torch.equal(got, nat.kmatw(X, Y, W, s, 2.0, nc, row_reduce=code))
    T = torch.zeros((1, m), dtype=torch.float64, device="cuda")
    B = torch.zeros((1, m), dtype=torch.float64, device="cuda")
    for r0 in range(0, N, 4000):
        Kt = nc * torch.exp(-(s * torch.cdist(X[r0:r0 + 4000].double(), Y.double())) ** 2)
        T += (Kt @ W.double()).sum(0, keepdim=True)
        B += (Kt @ W.double().abs()).sum(0, keepdim=True)
    if rr == "mean":
        T, B = T / N, B / N
    assert bool(((got.double() - T).abs() <= 1e-6 + 1e-5 * B).all())

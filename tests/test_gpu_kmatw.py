"""GPU parity of the fused, never-materialised K(X, Y) @ W kernel and its folded reduce
maps, and of the group-reduced Gram kernel, through the C ABI."""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from oracle import jaxrk_oracle as orc
from gpu_util import assert_close_truth, assert_kmatw_close, cu, f32, gauss, kparams

pytestmark = pytest.mark.gpu


def _truth(X, Y, W, s, p, rp=None, cp=None):
    G = orc.gengauss_gram_truth64(X, Y, s, p)
    if rp is not None:
        G = G * rp.astype(np.float64)[:, None]
    if cp is not None:
        G = G * cp.astype(np.float64)[None, :]
    W64 = W.astype(np.float64)
    return G @ W64, np.abs(G) @ np.abs(W64)


def _reduce(T, mode, group):
    if mode == nat.REDUCE_SUM:
        return T.sum(0, keepdims=True)
    if mode == nat.REDUCE_MEAN:
        return T.mean(0, keepdims=True)
    if mode == nat.REDUCE_BALANCED:
        return T.reshape(-1, group, T.shape[1]).sum(1)
    if mode == nat.REDUCE_BALANCED_MEAN:
        return T.reshape(-1, group, T.shape[1]).mean(1)
    return T


@pytest.mark.parametrize("kname", ["sqexp", "laplace", "gg15", "gg05"])
@pytest.mark.parametrize("d,m", [(16, 16), (8, 1), (3, 5), (32, 4), (17, 17), (64, 16), (100, 3), (128, 33), (1, 1)])
def test_kmatw_plain(kname, d, m):
    N, M = 700, 901
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    s, scale, p, nconst = kparams(kname, np.sqrt(d) * 1.2)
    T = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst).cpu().numpy()
    truth, bound = _truth(X, Y, W, s, p)
    assert T.shape == (N, m)
    assert_kmatw_close(T, truth, bound, f"{kname} d={d} m={m}")


@pytest.mark.parametrize("mode,group", [(nat.REDUCE_SUM, 0), (nat.REDUCE_MEAN, 0), (nat.REDUCE_BALANCED, 7),
                                        (nat.REDUCE_BALANCED_MEAN, 64), (nat.REDUCE_BALANCED, 512),
                                        (nat.REDUCE_BALANCED_MEAN, 1024), (nat.REDUCE_BALANCED, 1)])
@pytest.mark.parametrize("with_pref", [False, True])
def test_kmatw_folded_row_reduces(mode, group, with_pref):
    d, m = 16, 16
    N = 7 * 1024 if group else 3001
    M = 600
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    rp = np.random.RandomState(3).rand(N).astype(f32) if with_pref else None
    cp = np.random.RandomState(4).randn(M).astype(f32) if with_pref else None
    s, scale, p, nconst = kparams("gg15", 4.0)
    out = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, row_pref=rp, col_pref=cp, row_reduce=mode, group=group)
    truth, bound = _truth(X, Y, W, s, p, rp, cp)
    assert_kmatw_close(out.cpu().numpy(), _reduce(truth, mode, group), _reduce(bound, mode, group), f"mode={mode} g={group}")


def test_kmatw_few_rows_splits_the_j_range():
    """N small, M large: the j range is split across CTAs and re-summed deterministically."""
    d, m = 16, 16
    X, Y, W = gauss(0, 5, d), gauss(1, 200000, d), gauss(2, 200000, m)
    s, scale, p, nconst = kparams("sqexp", 4.0)
    Xd, Yd, Wd = cu(X), cu(Y), cu(W)
    a = nat.kmatw(Xd, Yd, Wd, scale, p, nconst)
    b = nat.kmatw(Xd, Yd, Wd, scale, p, nconst)
    assert torch.equal(a, b), "the split-j reduction must be deterministic"
    truth, bound = _truth(X, Y, W, s, p)
    assert_kmatw_close(a.cpu().numpy(), truth, bound)
    mean = nat.kmatw(Xd, Yd, Wd, scale, p, nconst, row_reduce=nat.REDUCE_MEAN)
    assert_kmatw_close(mean.cpu().numpy(), truth.mean(0, keepdims=True), bound.mean(0, keepdims=True))


def test_kmatw_wide_w_column_blocks_and_strides():
    d, m = 12, 70
    N, M = 300, 400
    base = cu(gauss(5, M, 96))
    W = base[:, 3:3 + m]                      # ld 96, unaligned
    Xb = cu(gauss(6, N, 20))
    X = Xb[:, 1:1 + d]
    Y = cu(gauss(7, M, d))
    s, scale, p, nconst = kparams("laplace", 3.0)
    T = nat.kmatw(X, Y, W, scale, p, nconst).cpu().numpy()
    truth, bound = _truth(X.cpu().numpy(), Y.cpu().numpy(), W.cpu().numpy(), s, p)
    assert_kmatw_close(T, truth, bound)


def test_kmatw_matches_gram_times_w_and_is_linear_at_scale():
    """Size-independent properties at a size the oracle cannot reach quickly."""
    N, M, d, m = 8192, 65536, 16, 16
    Xd, Yd = cu(gauss(0, N, d)), cu(gauss(1, M, d))
    W1, W2 = cu(gauss(2, M, m)), cu(gauss(3, M, m))
    s, scale, p, nconst = kparams("sqexp", 4.0)
    T1 = nat.kmatw(Xd, Yd, W1, scale, p, nconst)
    T2 = nat.kmatw(Xd, Yd, W2, scale, p, nconst)
    T12 = nat.kmatw(Xd, Yd, W1 + 2 * W2, scale, p, nconst)
    K = nat.gram(Xd, Yd, scale, p, nconst)
    ref = (K.double() @ W1.double())
    bound = (K.double() @ W1.double().abs())
    assert ((T1.double() - ref).abs() <= 1e-6 + 1e-5 * bound).all()
    assert ((T12 - (T1 + 2 * T2)).abs().double() <= 1e-6 + 3e-5 * bound).all()
    # sampled rows against the CPU oracle
    rows = [0, 17, 4095, 8191]
    X = Xd.cpu().numpy()[rows]
    truth, ab = _truth(X, Yd.cpu().numpy(), W1.cpu().numpy(), s, p)
    assert_kmatw_close(T1.cpu().numpy()[rows], truth, ab)


def test_kmatw_sum_equals_sum_of_rows_and_mean_consistency():
    N, M, d, m = 5000, 3000, 8, 4
    Xd, Yd, Wd = cu(gauss(0, N, d)), cu(gauss(1, M, d)), cu(gauss(2, M, m))
    s, scale, p, nconst = kparams("laplace", 3.0)
    T = nat.kmatw(Xd, Yd, Wd, scale, p, nconst).double()
    S = nat.kmatw(Xd, Yd, Wd, scale, p, nconst, row_reduce=nat.REDUCE_SUM).double()
    Mn = nat.kmatw(Xd, Yd, Wd, scale, p, nconst, row_reduce=nat.REDUCE_MEAN).double()
    assert torch.allclose(S, T.sum(0, keepdim=True), rtol=1e-5, atol=1e-5)
    assert torch.allclose(Mn * N, S, rtol=1e-5, atol=1e-5)


def test_kmatw_host_buffer_entry_point():
    N, M, d, m = 1500, 2000, 16, 16
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    s, scale, p, nconst = kparams("gg15", 4.0)
    out = np.empty((N, m), f32)
    nat.kmatw_host(X, Y, W, out, scale, p, nconst)
    truth, bound = _truth(X, Y, W, s, p)
    assert_kmatw_close(out, truth, bound)
    out1 = np.empty((1, m), f32)
    nat.kmatw_host(X, Y, W, out1, scale, p, nconst, row_reduce=nat.REDUCE_MEAN)
    assert_kmatw_close(out1, truth.mean(0, keepdims=True), bound.mean(0, keepdims=True))


@pytest.mark.parametrize("gx,gy,avg", [(128, 64, True), (256, 256, False), (128, 1, False), (512, 100, True)])
@pytest.mark.parametrize("kname", ["sqexp", "gg15"])
def test_gram_groupsum(gx, gy, avg, kname, c_oracle):
    d = 32
    N, M = 4 * gx, 6 * gy
    X, Y = gauss(0, N, d), gauss(1, M, d)
    rp, cp = np.random.RandomState(2).rand(N).astype(f32), np.random.RandomState(3).randn(M).astype(f32)
    s, scale, p, nconst = kparams(kname, np.sqrt(d))
    out = nat.gram_groupsum(cu(X), cu(Y), scale, p, nconst, gx, gy, avg, row_pref=rp, col_pref=cp).cpu().numpy()
    G = orc.gengauss_gram_truth64(X, Y, s, p)
    Gw = G * rp.astype(np.float64)[:, None] * cp.astype(np.float64)[None, :]
    truth = Gw.reshape(4, gx, 6, gy).sum((1, 3))
    bound = np.abs(Gw).reshape(4, gx, 6, gy).sum((1, 3))
    if avg:
        truth, bound = truth / (gx * gy), bound / (gx * gy)
    assert out.shape == (4, 6)
    assert_kmatw_close(out, truth, bound)
    # one cell through the C oracle's block sum as well
    cell = c_oracle.blocksum_truth64(X, Y, rp, cp, gx, 2 * gx, 2 * gy, 3 * gy, scale, p, nconst)
    np.testing.assert_allclose(out[1, 2] * (gx * gy if avg else 1), cell, rtol=1e-4, atol=1e-5 * bound[1, 2] * (gx * gy if avg else 1))


def test_gram_groupsum_self_is_symmetric_config4_shape():
    """BASELINE config 4 at reduced size: FiniteVec(k, X, [BalancedRed(g, True)]).inner()."""
    g, groups, d = 512, 16, 32
    X = gauss(0, g * groups, d)
    s, scale, p, nconst = kparams("sqexp", np.sqrt(d))
    out = nat.gram_groupsum(cu(X), None, scale, p, nconst, g, g, True)
    assert tuple(out.shape) == (groups, groups)
    assert torch.equal(out, out.T), "upper triangle computed, lower mirrored: exactly symmetric"
    G = orc.gengauss_gram_truth64(X[:2 * g], X[:2 * g], s, p).reshape(2, g, 2, g).mean((1, 3))
    assert_close_truth(out.cpu().numpy()[:2, :2], G, rtol=1e-5)


def test_kmatw_invalid_arguments():
    X, Y, W = cu(gauss(0, 6, 2)), cu(gauss(1, 4, 2)), cu(gauss(2, 4, 1))
    with pytest.raises(ValueError):
        nat.kmatw(X, Y, W, 1.0, 2.0, 1.0, row_reduce=nat.REDUCE_BALANCED, group=4)
    with pytest.raises(ValueError):
        nat.kmatw(X, cu(gauss(1, 5, 2)), W, 1.0, 2.0, 1.0)
    with pytest.raises(nat.JrkError):
        nat.kmatw(X, Y, W, 1.0, 3.0, 1.0)
    with pytest.raises(nat.JrkError):
        nat.gram_groupsum(X, Y, 1.0, 2.0, 1.0, 4, 2, False)


def test_kmatw_long_accumulation_all_positive_weights():
    """cfg3-sized j range (M = 2^20) with W = 1 (kernel mean embedding: every term positive, the worst case for a
    float32 running sum), N large enough that no j-split is used.  Sampled rows against float64."""
    N, M, d = 1 << 19, 1 << 20, 16
    g = torch.Generator(device="cuda:0").manual_seed(5)
    X = torch.randn(N, d, generator=g, device="cuda:0")
    Y = torch.randn(M, d, generator=g, device="cuda:0")
    W = torch.ones(M, 1, device="cuda:0")
    s, scale, p, nconst = kparams("sqexp", 4.0)
    T = nat.kmatw(X, Y, W, scale, p, nconst)
    rows = torch.tensor([0, 1, 777, N // 2, N - 1], device="cuda:0")
    d2 = torch.cdist(X[rows].double(), Y.double()) ** 2
    truth = (nconst * torch.exp(-(scale * scale) * d2)).sum(1, keepdim=True)
    err = ((T[rows].double() - truth).abs() / truth).max().item()
    assert err < 1e-5, f"relative error {err:.2e} of a 2^20-term positive sum"
    mean = nat.kmatw(X[:4096], Y, W, scale, p, nconst, row_reduce=nat.REDUCE_MEAN)
    d2 = torch.cdist(X[:4096].double(), Y.double()) ** 2
    tm = (nconst * torch.exp(-(scale * scale) * d2)).sum(1).mean()
    assert abs(mean.item() - tm.item()) / tm.item() < 1e-5


def test_inner_products_beyond_d128_take_the_gram_slab_route():
    """jrk_kmatw_f32 stops at d = 128; FiniteVec.inner must still work there (Gram by the direct kernel, the
    contraction by the library), as the reference does for any d."""
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import BalancedRed, LinearReduce, Mean, Prefactors
    from jaxrk_b200.rkhs import FiniteVec, inner
    rng = np.random.RandomState(0)
    N, M, d, m = 600, 900, 150, 5
    X, Y, A = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32), rng.randn(m, M).astype(f32)
    pp = rng.rand(N).astype(f32)
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    s = orc.simple_scaler_s(orc.gauss_length_scale(float(np.sqrt(d))))
    G = orc.gengauss_gram_truth64(X, Y, s, 2.0)
    got = inner(FiniteVec(k, X, [Prefactors(pp), BalancedRed(4)]), FiniteVec(k, Y, [LinearReduce(A)]))
    want = ((pp[:, None] * G) @ A.T.astype(np.float64)).reshape(N // 4, 4, m).sum(1)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-5, atol=1e-6)
    got2 = inner(FiniteVec(k, X, [Mean()]), FiniteVec(k, Y, [Mean()]))
    np.testing.assert_allclose(got2.cpu().numpy(), G.mean(keepdims=True).reshape(1, 1), rtol=2e-5)
    with pytest.raises(nat.JrkError):                          # the C entry point itself says so loudly
        nat.kmatw(cu(X), cu(Y), cu(A.T.copy()), 0.1, 2.0, 1.0)

"""GPU parity of the sibling kernels on the ScaledPairwiseDistance front-end: PeriodicKernel,
ThreshSpikeKernel, the diag=True branch -- through the C ABI (jrk_gram_kind_f32 / jrk_kmatw_kind_f32 /
jrk_dist_diag_f32) and through the reference-shaped API, against the oracle and the reference's own outputs
(tests/golden/reference_siblings.npz)."""
import os

import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from jaxrk_b200.kern import GenGaussKernel, PeriodicKernel, ScaledPairwiseDistance, SimpleScaler, ThreshSpikeKernel
from jaxrk_b200.reduce import BalancedRed, LinearReduce, Mean, Prefactors, Sum
from jaxrk_b200.rkhs import FiniteVec, inner
from oracle import jaxrk_oracle as orc
from gpu_util import ATOL, RTOL, assert_close_truth, assert_kmatw_close, cu, f32, gauss

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_siblings.npz"))
X, Y, X1 = G["X"], G["Y"], G["X1"]
EPS = 6e-8


def _assert_periodic(got, a, b, period, ls, what):
    truth = orc.periodic_gram_truth64(a, b, period, ls)
    cond = orc.periodic_gram_cond64(a, b, period, ls)
    err = np.abs(np.asarray(got, np.float64) - truth)
    lim = ATOL + RTOL * np.abs(truth) + 8 * EPS * cond     # 8 eps: the float32 rounding of the distance itself
    assert (err <= lim).all(), f"{what}: worst err {err.max():.3e}, limit there {lim.flat[err.argmax()]:.3e}"


@pytest.mark.parametrize("i", range(3))
def test_periodic_gram_small_against_truth_and_reference(i):
    period, ls = (float(v) for v in G["periodic_params"][i])
    k = PeriodicKernel(period, ls)
    n0 = nat.launch_count()
    Kxy = k(cu(X), cu(Y)).cpu().numpy()
    assert nat.launch_count() - n0 == 1
    _assert_periodic(Kxy, X, Y, period, ls, "periodic XY")
    Kxx = k(cu(X)).cpu().numpy()
    _assert_periodic(Kxx, X, None, period, ls, "periodic XX")
    assert (Kxx == Kxx.T).all() and (np.diag(Kxx) == 1.0).all()      # exact at distance zero, bitwise symmetric
    assert (Kxy[:5, :5].diagonal() == 1.0).all()                       # coincident points across X and Y
    _assert_periodic(k(cu(X1)).cpu().numpy(), X1, None, period, ls, "periodic 1-d")
    # the reference's float32 result, wherever the reference is itself within tolerance of the truth
    truth = orc.periodic_gram_truth64(X, Y, period, ls)
    ref = G[f"periodic_{i}_XY"].astype(np.float64)
    ok = np.abs(ref - truth) <= ATOL + RTOL * np.abs(truth)
    assert ok.mean() > 0.5
    cond = orc.periodic_gram_cond64(X, Y, period, ls)
    assert (np.abs(Kxy - ref)[ok] <= 2 * (ATOL + RTOL * np.abs(ref[ok])) + 8 * EPS * cond[ok]).all()
    d = k(cu(X[:29]), cu(Y), diag=True).cpu().numpy()
    assert np.allclose(d, G[f"periodic_{i}_diag"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("d,period,ls", [(1, 1.0, 1.0), (5, 3.0, 0.8), (16, 7.5, 1.5), (40, 11.0, 2.0)])
def test_periodic_gram_ragged_sizes(d, period, ls):
    N, M = 700 + 3, 517
    a, b = gauss(1, N, d), gauss(2, M, d)
    K = nat.gram_spec(cu(a), cu(b), nat.KernelSpec(nat.KIND_PERIODIC, (1.0 / period, ls))).cpu().numpy()
    _assert_periodic(K, a, b, period, ls, f"periodic d={d}")


@pytest.mark.parametrize("i", range(3))
def test_threshspike_gram(i):
    ls, shape, spike, non_spike, thr = (float(v) for v in G["thresh_params"][i])
    dist = ScaledPairwiseDistance(scaler=SimpleScaler(1. / ls), power=shape)
    k = ThreshSpikeKernel(dist, spike, non_spike, thr)
    K = k(cu(X), cu(Y)).cpu().numpy()
    vals, margin = orc.threshspike_gram_truth64(X, Y, orc.simple_scaler_s(ls), shape, spike, non_spike, thr)
    clear = margin > 1e-4
    assert clear.mean() > 0.95 and (K[clear] == vals[clear].astype(f32)).all()
    assert (K[:5, :5].diagonal() == f32(spike)).all()                  # coincident pairs: distance exactly 0 <= thr
    ref = G[f"thresh_{i}_XY"]
    assert (K[clear] == ref[clear]).all()                               # the reference's own output
    # consistent with the distance matrix of jrk_dist_f32 away from the float32 resolution of the threshold
    D = dist(cu(X), cu(Y)).cpu().numpy()
    away = np.abs(D - f32(thr)) > 1e-5 * max(thr, 1.0)
    assert (K[away] == np.where(D <= f32(thr), f32(spike), f32(non_spike))[away]).all()
    Kxx = k(cu(X)).cpu().numpy()
    assert (Kxx == Kxx.T).all() and (np.diag(Kxx) == f32(spike)).all()
    with pytest.raises(AssertionError):
        k(cu(X), cu(Y), diag=True)
    with pytest.raises(AssertionError):
        ThreshSpikeKernel(dist, 1.0, 2.0, 0.5)


@pytest.mark.parametrize("i", range(3))
def test_diag_branch(i):
    ls, shape = (float(v) for v in G["diag_params"][i])
    dist = ScaledPairwiseDistance(scaler=SimpleScaler(1. / ls), power=shape)
    n0 = nat.launch_count()
    got = dist(cu(X[:29]), cu(Y), diag=True).cpu().numpy()
    assert nat.launch_count() - n0 == 1 and got.shape == (29,)
    assert_close_truth(got, orc.scaled_dist_diag_truth64(X[:29], Y, orc.simple_scaler_s(ls), shape), "diag dist")
    assert np.allclose(got, G[f"diag_{i}_dist"], rtol=1e-5, atol=1e-6)
    assert (dist(cu(X), None, diag=True).cpu().numpy() == 0).all()
    kd = GenGaussKernel.make(ls, shape)(cu(X[:29]), cu(Y), diag=True).cpu().numpy()
    assert np.allclose(kd, G[f"diag_{i}_gengauss"], rtol=1e-5, atol=1e-6)


def test_diag_branch_per_dimension_scale_and_strides():
    ds = G["diag_dimscale"]
    dist = ScaledPairwiseDistance(scaler=SimpleScaler(ds), power=1.5)
    got = dist(cu(X[:29]), cu(Y), diag=True).cpu().numpy()
    assert np.allclose(got, G["diag_dimscale_dist"], rtol=1e-5, atol=1e-6)
    big = cu(np.concatenate([X[:29], Y], 1))                           # strided views of one buffer
    got2 = dist(big[:, :3], big[:, 3:], diag=True).cpu().numpy()
    assert np.allclose(got2, got, rtol=1e-6, atol=1e-7)


def test_inner_over_periodic_kernel_with_reduce_chains():
    k = PeriodicKernel(2.5, 0.7)
    pref, A = G["inner_pref"], G["inner_A"]
    cases = {
        "inner_periodic_plain": ([], [], X),
        "inner_periodic_lin": ([Prefactors(cu(pref))], [LinearReduce(cu(A))], X),
        "inner_periodic_mean_sum": ([Mean()], [Sum()], X),
        "inner_periodic_balanced": ([BalancedRed(4, True)], [LinearReduce(cu(A))], X[:36]),
    }
    for name, (cx, cy, xa) in cases.items():
        n0 = nat.launch_count()
        got = inner(FiniteVec(k, cu(xa), cx), FiniteVec(k, cu(Y), cy)).cpu().numpy()
        assert nat.launch_count() > n0
        want = G[name]
        assert got.shape == want.shape, name
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 2e-6 + 3e-5 * scale, (name, np.abs(got - want).max())


@pytest.mark.parametrize("kind", ["periodic", "thresh"])
@pytest.mark.parametrize("reduce", [nat.REDUCE_NONE, nat.REDUCE_MEAN])
def test_kmatw_kind_matches_materialised_gram(kind, reduce):
    """Fused K@W of a sibling kernel == (its materialised Gram) @ W, to accumulation accuracy."""
    N, M, d, m = 1500 + 7, 2048 + 5, 6, 9
    a, b, W = gauss(3, N, d), gauss(4, M, d), gauss(5, M, m)
    if kind == "periodic":
        spec = nat.KernelSpec(nat.KIND_PERIODIC, (1.0 / 4.0, 0.9))
        Kt = orc.periodic_gram_truth64(a, b, 4.0, 0.9)
    else:
        spec = nat.KernelSpec(nat.KIND_THRESHSPIKE, (0.5, 1.0, 1.5, -0.25, 1.6))
        Kt, margin = orc.threshspike_gram_truth64(a, b, orc.simple_scaler_s(2.0), 1.0, 1.5, -0.25, 1.6)
    Kg = nat.gram_spec(cu(a), cu(b), spec)
    T = nat.kmatw_spec(cu(a), cu(b), cu(W), spec, row_reduce=reduce).cpu().numpy()
    want = (Kg.double() @ cu(W).double()).cpu().numpy()
    bound = (Kg.double().abs() @ cu(W).double().abs()).cpu().numpy()
    if reduce == nat.REDUCE_MEAN:
        want, bound = want.mean(0, keepdims=True), bound.mean(0, keepdims=True)
    assert_kmatw_close(T, want, bound, f"kmatw {kind}")
    if kind == "periodic":   # and the Gram itself against the float64 truth
        _assert_periodic(Kg.cpu().numpy(), a, b, 4.0, 0.9, "periodic gram")


def test_kind_entry_points_validate_arguments():
    a = cu(gauss(0, 8, 2))
    out = torch.empty((8, 8), device=a.device)
    L = nat.lib()
    import ctypes
    bad = (ctypes.c_float * 2)(-1.0, 1.0)
    rc = L.jrk_gram_kind_f32(a.data_ptr(), None, out.data_ptr(), 8, 8, 2, 2, 2, 8, nat.KIND_PERIODIC, bad, 2, None, 0, None, 0, None)
    assert rc == 1 and "positive" in nat.last_error()
    rc = L.jrk_gram_kind_f32(a.data_ptr(), None, out.data_ptr(), 8, 8, 2, 2, 2, 8, 7, bad, 2, None, 0, None, 0, None)
    assert rc == 1 and "unknown kernel kind" in nat.last_error()
    ok = (ctypes.c_float * 2)(1.0, 1.0)
    rc = L.jrk_gram_kind_f32(a.data_ptr(), None, out.data_ptr(), 8, 8, 2, 2, 2, 8, nat.KIND_PERIODIC, ok, 2, None, nat.PATH_TF32X3, None, 0, None)
    assert rc == 3
    # the GenGauss kind forwards to jrk_gram_f32
    s, p = 0.5, 2.0
    nc = float(orc.gengauss_nconst_ref32(np.array([[s]], f32), p).reshape(-1)[0])
    K1 = nat.gram_spec(a, None, nat.KernelSpec(nat.KIND_GENGAUSS, (s, p, nc)))
    K2 = nat.gram(a, None, s, p, nc)
    assert torch.equal(K1, K2)


def test_combvec_product_of_grams():
    """CombVec.inner (jaxrk/rkhs/vector.py:242-247) against the reference's own outputs: the product of two
    squared-exponential Grams runs as ONE fused call on concatenated inputs, a Gauss x Laplace product as two."""
    import operator
    from jaxrk_b200.rkhs import CombVec
    kg1, kg2, kl = GenGaussKernel.make_gauss(1.3), GenGaussKernel.make_gauss(0.8), GenGaussKernel.make_laplace(1.1)
    XR, XL, YR, YL, B = (cu(G[n]) for n in ("comb_XR", "comb_XL", "comb_YR", "comb_YL", "comb_B"))
    for name, (ka, kb), launches in (("gg", (kg1, kg2), 1), ("gl", (kg1, kl), 2)):
        cx = CombVec(FiniteVec(ka, XR, []), FiniteVec(kb, XL, []), operator.mul, [BalancedRed(3, True)])
        cy = CombVec(FiniteVec(ka, YR, []), FiniteVec(kb, YL, []), operator.mul, [LinearReduce(B)])
        assert [len(cx), len(cy)] == list(G[f"comb_{name}_len"])
        n0 = nat.launch_count()
        got = cx.inner(cy).cpu().numpy()
        n1 = nat.launch_count() - n0
        assert (n1 <= 5) if launches == 1 else (n1 >= 2), n1      # fused: K@W (+ operand packing / finalize); else two Grams
        want = G[f"comb_{name}_xy"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-6 + 3e-5 * np.abs(want).max(), name
        gxx = cx.inner().cpu().numpy()
        wxx = G[f"comb_{name}_xx"]
        # (the reference misses truth on self-Gram diagonals for p < 2: up to 1.2e-4 there)
        assert np.abs(gxx - wxx).max() <= (2e-6 if name == "gg" else 2e-4) + 3e-5 * np.abs(wxx).max(), name
    assert len(cx.extend_reduce([Sum()])) == 1 and len(cx.centered()) == len(cx)

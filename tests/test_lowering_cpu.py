"""Host logic of the reference-shaped API (kern / reduce / rkhs) on CPU.

The product path has no CPU fallback, so here the four native calls are replaced by
oracle-backed stand-ins (tests may use oracle/; the product never does).  What is under
test is everything ABOVE the C ABI: parameter conventions of GenGaussKernel, the Reduce
family's array semantics, and the lowering of (chain_x, chain_y) onto Gram / K@W /
group-sum descriptors -- checked against the outputs of the reference's own Python code
(tests/golden/reference_numpy_shim.npz).  tests/test_gpu_api.py repeats the battery with
the real kernels.
"""
import numpy as np
import pytest
import torch

import jaxrk_b200._array as arr
from jaxrk_b200 import _native as nat
from jaxrk_b200.kern import GenGaussKernel
from jaxrk_b200.reduce import BalancedRed, BlockReduce, Center, LinearReduce, Mean, Prefactors, Reduce, Sum, TileView
from jaxrk_b200.rkhs import Cdo, Cmo, Cov_inv, CovOp, FiniteOp, FiniteVec, inner
from jaxrk_b200.rkhs import _lowering as low
from oracle import jaxrk_oracle as orc

f32 = np.float32
CALLS = []


def _np(t):
    return None if t is None else np.asarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float64)


def _truth_gram(X, Y, scale, power, nconst, dim_scale):
    X = _np(X).astype(f32)
    Y = X if Y is None else _np(Y).astype(f32)
    if dim_scale is not None:
        ds = np.asarray(dim_scale, f32)[None, :]
        X, Y = (X * ds).astype(f32), (Y * ds).astype(f32)
    return orc.gengauss_gram_truth64(X, Y, np.array([[scale]], f32), power, nconst=f32(nconst))


def fake_gram(X, Y, scale, power, nconst, dim_scale=None, path=0, out=None):
    CALLS.append("gram")
    return torch.from_numpy(_truth_gram(X, Y, scale, power, nconst, dim_scale).astype(f32))


def fake_dist(X, Y, scale, power, dim_scale=None):
    CALLS.append("dist")
    g = _truth_gram(X, Y, scale, power, 1.0, dim_scale)
    return torch.from_numpy((-np.log(g)).astype(f32))


def fake_kmatw(X, Y, W, scale, power, nconst, dim_scale=None, row_pref=None, col_pref=None, row_reduce=0, group=0, out=None):
    CALLS.append("kmatw")
    G = _truth_gram(X, Y, scale, power, nconst, dim_scale)
    if row_pref is not None:
        G = G * _np(row_pref)[:, None]
    if col_pref is not None:
        G = G * _np(col_pref)[None, :]
    T = G @ _np(W)
    if row_reduce == nat.REDUCE_SUM:
        T = T.sum(0, keepdims=True)
    elif row_reduce == nat.REDUCE_MEAN:
        T = T.mean(0, keepdims=True)
    elif row_reduce in (nat.REDUCE_BALANCED, nat.REDUCE_BALANCED_MEAN):
        T = T.reshape(-1, group, T.shape[1])
        T = T.sum(1) if row_reduce == nat.REDUCE_BALANCED else T.mean(1)
    return torch.from_numpy(T.astype(f32))


def fake_groupsum(X, Y, scale, power, nconst, gx, gy, average, dim_scale=None, row_pref=None, col_pref=None):
    CALLS.append("groupsum")
    G = _truth_gram(X, Y, scale, power, nconst, dim_scale)
    if row_pref is not None:
        G = G * _np(row_pref)[:, None]
    if col_pref is not None:
        G = G * _np(col_pref)[None, :]
    G = G.reshape(G.shape[0] // gx, gx, G.shape[1] // gy, gy).sum((1, 3))
    return torch.from_numpy((G / (gx * gy) if average else G).astype(f32))


def fake_dist_diag(X, Y, scale, power, dim_scale=None):
    CALLS.append("dist_diag")
    s = np.array([[scale]], f32) if dim_scale is None else np.asarray(dim_scale, f32)[None, :]
    return torch.from_numpy(orc.scaled_dist_diag_truth64(_np(X).astype(f32), _np(Y).astype(f32), s, power).astype(f32))


def fake_gram_affine(K, row_mul=None, col_mul=None, row_add=None, col_add=None, add_const=0.0, out=None):
    CALLS.append("gram_affine")
    G = _np(K)
    if row_mul is not None:
        G = G * _np(row_mul)[:, None]
    if col_mul is not None:
        G = G * _np(col_mul)[None, :]
    if row_add is not None:
        G = G + _np(row_add)[:, None]
    if col_add is not None:
        G = G + _np(col_add)[None, :]
    res = torch.from_numpy((G + add_const).astype(f32))
    (K if out is None else out).copy_(res)
    return K if out is None else out


@pytest.fixture(autouse=True)
def cpu_standins(monkeypatch):
    monkeypatch.setattr(arr, "_device", torch.device("cpu"))
    monkeypatch.setattr(nat, "gram", fake_gram)
    monkeypatch.setattr(nat, "dist", fake_dist)
    monkeypatch.setattr(nat, "kmatw", fake_kmatw)
    monkeypatch.setattr(nat, "gram_groupsum", fake_groupsum)
    monkeypatch.setattr(nat, "dist_diag", fake_dist_diag)
    monkeypatch.setattr(nat, "gram_affine", fake_gram_affine)
    CALLS.clear()
    yield


def chains_x(g):
    pp, Ap = g["ch_pp"], g["ch_Ap"]
    return {"none": [], "pref": [Prefactors(pp)], "sum": [Sum()], "mean": [Mean()], "pref_sum": [Prefactors(pp), Sum()],
            "bal4": [BalancedRed(4)], "bal4avg": [BalancedRed(4, True)], "pref_bal3": [Prefactors(pp), BalancedRed(3)],
            "lin": [LinearReduce(Ap)], "pref_lin": [Prefactors(pp), LinearReduce(Ap)], "center": [Center()],
            "bal2_center": [BalancedRed(2), Center()]}


def chains_y(g):
    pq, Aq, Aq12 = g["ch_pq"], g["ch_Aq"], g["ch_Aq12"]
    return {"none": [], "pref": [Prefactors(pq)], "sum": [Sum()], "mean": [Mean()], "lin": [LinearReduce(Aq)],
            "pref_lin": [Prefactors(pq), LinearReduce(Aq)], "bal3_lin": [BalancedRed(3), LinearReduce(Aq12)],
            "bal6avg": [BalancedRed(6, True)], "tile": [Mean(), TileView(4)]}


def test_kernel_parameters_match_reference(golden):
    for row, nconst, var, (kind, ls, sh) in zip(golden["rbf_row0"], golden["rbf_nconst"], golden["rbf_var"],
                                                golden["rbf_params"]):
        ls = f32(ls)
        k = (GenGaussKernel.make_gauss(ls) if kind == 0 else GenGaussKernel.make_laplace(ls) if kind == 1
             else GenGaussKernel.make(ls, float(sh)))
        assert k.nconst.shape == (1, 1) and k.nconst.dtype == f32
        np.testing.assert_allclose(k.nconst.reshape(-1)[0], nconst, rtol=3e-7)
        np.testing.assert_allclose(np.asarray(k.var()).reshape(-1)[0], var, rtol=1e-6)
        assert k.dist.power == sh
        got = k(np.arange(8)[:, None])  # int input, as tests/test_kern.py:17
        np.testing.assert_allclose(got[0].numpy(), row, rtol=2e-5, atol=1e-7)
    with pytest.raises(AssertionError):
        GenGaussKernel.make(1.0, 2.5)
    with pytest.raises(AssertionError):
        GenGaussKernel.make(1.0, 0.0)
    with pytest.raises(AssertionError):
        GenGaussKernel.make(-1.0, 1.0)


def test_perdim_scale_and_diag(golden):
    kd = GenGaussKernel.make(golden["perdim_ls"], 1.5)
    np.testing.assert_allclose(kd.nconst, golden["perdim_nconst"], rtol=3e-7)
    assert kd.kernel_args() is None and not kd.dist.is_global
    np.testing.assert_allclose(kd.dist(golden["gram_X"], golden["gram_Y"]).numpy(), golden["perdim_dist_XY"], rtol=2e-5, atol=1e-6)
    X = golden["gram_X"]
    for name, k in (("gauss", GenGaussKernel.make_gauss(1.3)), ("laplace", GenGaussKernel.make_laplace(0.7)),
                    ("gg15", GenGaussKernel.make(2.0, 1.5)), ("gg05", GenGaussKernel.make(0.9, 0.5))):
        np.testing.assert_allclose(k(X, X[::-1].copy(), diag=True).numpy().reshape(-1), golden[f"gram_{name}_diag"].reshape(-1),
                                   rtol=2e-5, atol=1e-9)
        assert torch.all(k.dist(X, diag=True) == 0)


def test_inner_chain_battery_against_reference_outputs(golden):
    kg = GenGaussKernel.make(1.7, 1.5)
    P, Q = golden["ch_P"], golden["ch_Q"]
    cx, cy = chains_x(golden), chains_y(golden)
    for nx in cx:
        for ny in cy:
            got = inner(FiniteVec(kg, P, cx[nx]), FiniteVec(kg, Q, cy[ny])).numpy()
            want = golden[f"ch_inner__{nx}__{ny}"]
            assert got.shape == want.shape, (nx, ny, got.shape, want.shape)
            np.testing.assert_allclose(got, want, rtol=3e-5, atol=3e-6, err_msg=f"{nx}/{ny}")
    for nx in cx:
        got = FiniteVec(kg, P, cx[nx]).inner().numpy()
        want = golden[f"ch_self__{nx}"]
        assert got.shape == want.shape
        # the reference misses truth on self-Gram diagonals for p < 2: 1e-4 there
        np.testing.assert_allclose(got, want, rtol=3e-5, atol=2e-4 if nx in ("none", "pref", "center", "bal2_center") else 2e-5,
                                   err_msg=nx)
    el = FiniteVec.construct_RKHS_Elem(kg, P, golden["ch_pp"])
    np.testing.assert_allclose(el(Q).numpy(), golden["ch_eval_elem_at_Q"], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(FiniteVec(kg, P, [LinearReduce(golden["ch_Ap"])])(Q).numpy(), golden["ch_eval_lin_at_Q"],
                               rtol=3e-5, atol=3e-6)


def test_lowering_picks_the_fused_kernels(golden):
    kg = GenGaussKernel.make(1.7, 1.5)
    P, Q = golden["ch_P"], golden["ch_Q"]
    cx, cy = chains_x(golden), chains_y(golden)

    def calls(a, b):
        CALLS.clear()
        inner(FiniteVec(kg, P, a), FiniteVec(kg, Q, b))
        return list(CALLS)

    assert calls(cx["none"], cy["none"]) == ["gram"]
    assert calls(cx["pref"], cy["lin"]) == ["kmatw"]
    assert calls(cx["mean"], cy["pref_lin"]) == ["kmatw"]
    assert calls(cx["lin"], cy["none"]) == ["kmatw"]        # transposed problem
    assert calls(cx["pref_bal3"], cy["sum"]) == ["kmatw"]   # BalancedRed folded as row reduce
    assert calls(cx["sum"], cy["bal6avg"]) == ["kmatw"]     # transposed, balanced on the other side
    assert calls(cx["bal2_center"], cy["mean"]) == ["kmatw"]  # Center acts on the reduced result
    # Center on a materialised Gram: the column means as one skinny fused call, then Gram + ONE affine pass
    assert calls(cx["center"], cy["none"]) == ["kmatw", "gram", "gram_affine"]
    X = np.random.RandomState(0).randn(256, 3).astype(f32)
    CALLS.clear()
    v = FiniteVec(kg, X, [BalancedRed(128, True)])
    got = v.inner()
    assert CALLS == ["groupsum"] and tuple(got.shape) == (2, 2)
    G = orc.gengauss_gram_truth64(X, None, orc.simple_scaler_s(1.7), 1.5)
    np.testing.assert_allclose(got.numpy(), G.reshape(2, 128, 2, 128).mean((1, 3)), rtol=1e-5)
    v.inner()
    assert CALLS == ["groupsum"], "the self inner product is cached (Cov_solve asks for it twice)"


def test_fold_chain_rules(golden):
    dev = torch.device("cpu")
    pp = torch.from_numpy(golden["ch_pp"])
    f = low.fold_chain([Prefactors(pp), Prefactors(pp), BalancedRed(3, True), Center(), Sum()], 24, dev)
    assert f.kind == low.BALANCED and f.group == 3 and f.average and len(f.post) == 2 and f.out_len == 8
    np.testing.assert_allclose(f.pref.numpy(), (pp * pp).numpy())
    A = torch.from_numpy(golden["ch_Ap"])
    f = low.fold_chain([LinearReduce(A), Prefactors(torch.tensor([1., 2., 3.])), Sum()], 24, dev)
    assert f.kind == low.LINEAR and tuple(f.A.shape) == (1, 24) and f.post == []
    np.testing.assert_allclose(f.A.numpy(), (A * torch.tensor([[1.], [2.], [3.]])).sum(0, keepdim=True).numpy(), rtol=1e-6)
    f = low.fold_chain([TileView(48)], 24, dev)
    assert f.kind == low.NONE and len(f.post) == 1
    assert Reduce.final_len(24, [BalancedRed(4), LinearReduce(torch.ones(2, 6))]) == 2
    with pytest.raises(AssertionError):
        Reduce.final_len(25, [BalancedRed(4)])
    with pytest.raises(AssertionError):
        LinearReduce(torch.ones(2, 5)).new_len(6)


def test_operators_against_reference_outputs(golden):
    g = golden
    kx, ky = GenGaussKernel.make_gauss(0.9), GenGaussKernel.make_laplace(1.1)
    inp, outp, ref = FiniteVec(kx, g["op_X"]), FiniteVec(ky, g["op_Y"]), FiniteVec(ky, g["op_R"])
    op = FiniteOp(inp, outp, g["op_M"])
    res = op @ FiniteVec(kx, g["op_Xt"])
    np.testing.assert_allclose(res.reduce[-1].linear_map.numpy(), g["op_matmul_vec_lin"], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(res(g["op_Yt"]).numpy(), g["op_matmul_vec_eval"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose((op @ FiniteVec(kx, g["op_Xt"][:1]))(g["op_Yt"]).numpy(), g["op_matmul_vec1_eval"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose((op @ g["op_Xt"])(g["op_Yt"]).numpy(), g["op_matmul_arr_eval"], rtol=1e-4, atol=1e-5)
    op2 = FiniteOp(outp, inp, g["op_M"].T.copy())
    np.testing.assert_allclose((op @ op2).matr.numpy(), g["op_matmul_op_matr"], rtol=1e-4, atol=1e-4)
    # the regularised inverse amplifies the (correct) ~1e-4 diagonal difference to the reference
    np.testing.assert_allclose(Cov_inv(CovOp(inp), regul=0.1).matr.numpy(), g["cov_inv_matr"], rtol=1e-3, atol=5e-2)
    cmo = Cmo(inp, outp, regul=0.1)
    assert tuple(cmo.matr.shape) == g["cmo_matr"].shape
    np.testing.assert_allclose((cmo @ FiniteVec(kx, g["op_Xt"]))(g["op_Yt"]).numpy(), g["cmo_apply_eval"], rtol=1e-3, atol=1e-3)
    cdo = Cdo(inp, outp, ref, regul=0.1)
    assert tuple(cdo.matr.shape) == g["cdo_matr"].shape
    np.testing.assert_allclose((cdo @ FiniteVec(kx, g["op_Xt"]))(g["op_Yt"]).numpy(), g["cdo_apply_eval"], rtol=2e-3, atol=2e-3)


def test_errors_match_reference_conventions(golden):
    kg = GenGaussKernel.make(1.7, 1.5)
    v = FiniteVec(kg, golden["ch_P"])
    with pytest.raises(ValueError):
        v.inner(v, full=False)
    with pytest.raises(AssertionError):
        v.inner(FiniteVec(GenGaussKernel.make(1.7, 1.5), golden["ch_Q"]))  # kernels compare by identity (quirk Q5)
    with pytest.raises(AssertionError):
        FiniteVec(kg, np.zeros(5, f32))
    with pytest.raises(AssertionError):
        FiniteOp(v, v, np.zeros((3, 3), f32))
    assert len(v.sum()) == 1 and len(v.mean()) == 1 and len(v.extend_reduce([BalancedRed(4)])) == 6


def test_combvec_lowering_against_reference_outputs():
    """CombVec.inner: product of squared-exponential Grams lowered as one GenGauss problem on concatenated, pre-scaled
    inputs; other products through two reduced Grams (reference outputs: tests/golden/reference_siblings.npz)."""
    import operator
    import os
    from jaxrk_b200.rkhs import CombVec
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_siblings.npz"))
    kg1, kg2, kl = GenGaussKernel.make_gauss(1.3), GenGaussKernel.make_gauss(0.8), GenGaussKernel.make_laplace(1.1)
    for name, (ka, kb), calls in (("gg", (kg1, kg2), ["kmatw"]), ("gl", (kg1, kl), ["gram", "gram"])):
        cx = CombVec(FiniteVec(ka, g["comb_XR"], []), FiniteVec(kb, g["comb_XL"], []), operator.mul, [BalancedRed(3, True)])
        cy = CombVec(FiniteVec(ka, g["comb_YR"], []), FiniteVec(kb, g["comb_YL"], []), operator.mul, [LinearReduce(g["comb_B"])])
        CALLS.clear()
        got = cx.inner(cy).numpy()
        assert CALLS == calls, CALLS
        np.testing.assert_allclose(got, g[f"comb_{name}_xy"], rtol=3e-5, atol=3e-6)
        # the reference misses truth on self-Gram diagonals for p < 2: up to 1.2e-4 there
        np.testing.assert_allclose(cx.inner().numpy(), g[f"comb_{name}_xx"], rtol=3e-5, atol=3e-6 if name == "gg" else 2e-4)


def test_gram_utilities_and_point_representant_against_reference_outputs():
    """jaxrk/utilities/gram.py (RKHS distances from Grams, representer choice) and FiniteVec.point_representant
    (rkhs/vector.py:161-174) against outputs of the reference's own code."""
    import os
    from jaxrk_b200.utilities import (choose_representer, choose_representer_from_gram, gram_projection, rkhs_gram_cdist,
                                      rkhs_gram_cdist_ignore_const)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_siblings.npz"))
    X, Y = g["X"], g["Y"]
    kq = GenGaussKernel.make_gauss(0.9)
    Gxx, Gyy, Gxy = kq(X), kq(Y), kq(X, Y)
    np.testing.assert_allclose(rkhs_gram_cdist(Gxy, Gxx, Gyy).numpy(), g["util_cdist_xy"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(rkhs_gram_cdist(Gxy, Gxx, Gyy, power=1.).clamp_min(0).numpy(), np.nan_to_num(g["util_cdist_xy_p1"]), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(rkhs_gram_cdist_ignore_const(Gxy, Gyy).numpy(), g["util_cdist_ignore"], rtol=2e-5, atol=2e-6)
    fac = g["util_factors"]
    assert [choose_representer(X, fac, kq), choose_representer_from_gram(Gxx, fac)] == list(g["util_representer"])
    assert (gram_projection(Gxy, Gyy, Gxx).numpy() == g["util_projection"]).all()
    with pytest.raises(NotImplementedError):
        gram_projection(Gxy, Gyy, Gxx, method="pos_proj")
    vec = FiniteVec(kq, X, [LinearReduce(g["util_A2"])])
    np.testing.assert_allclose(vec.point_representant("inspace_point", keepdims=True).numpy(), g["util_point_repr"], rtol=1e-6)
    np.testing.assert_allclose(vec.point_representant("mean", keepdims=True).numpy(), g["util_point_mean"], rtol=2e-5, atol=1e-6)
    # a single (self) Gram must be exactly symmetric, as the reference asserts (utilities/gram.py:18)
    d = rkhs_gram_cdist(Gxx)
    assert torch.equal(d, d.T) and float(torch.diagonal(d).abs().max()) == 0.0


def test_gp_regression_against_reference_outputs():
    """jaxrk/gp.py:6-33 over FiniteVec.inner: posterior mean / covariance, marginal and predictive likelihood against the
    reference's own outputs (float32 Cholesky on both sides: loose on the small posterior covariance entries)."""
    import os
    from jaxrk_b200.gp import GP
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_siblings.npz"))
    k = GenGaussKernel.make_gauss(0.8)
    gp = GP(FiniteVec(k, g["gp_xtr"], []), g["gp_ytr"], 0.1, 1.3, 0.2)
    mu, var = gp.predict(FiniteVec(k, g["gp_xte"], []))
    np.testing.assert_allclose(mu.numpy(), g["gp_mu"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(var.numpy(), g["gp_var"], rtol=2e-3, atol=2e-4)
    assert abs(float(gp.marginal_likelihood()) - float(g["gp_ml"][0])) <= 2e-3 * abs(float(g["gp_ml"][0])) + 2e-3
    ppl = gp.post_pred_likelihood(FiniteVec(k, g["gp_xte"], []), g["gp_yte"])[2]
    assert abs(float(ppl) - float(g["gp_ppl"][0])) <= 2e-2 * abs(float(g["gp_ppl"][0])) + 5e-2


def _fake_vjp(X, Y, A, scale, power, nconst):
    gX, gY, gs, _ = orc.gengauss_vjp_truth64(_np(X).astype(f32), _np(Y).astype(f32), A, np.array([[scale]], f32), power, nconst=nconst)
    return gX


def fake_kmatw_grad(X, Y, W, G, scale, power, nconst, want_scale=True, want_power=False):
    CALLS.append("kmatw_grad")
    A = _np(G) @ _np(W).T
    Xn, Yn = _np(X).astype(f32), _np(Y).astype(f32)
    s = np.array([[scale]], f32)
    gX, _, _, _ = orc.gengauss_vjp_truth64(Xn, Yn, A, s, power, nconst=nconst)
    gs_rows = None
    if want_scale:   # per-row contributions to dL/dscale: finite rows of the analytic formula
        sq = ((Xn[:, None, :].astype(np.float64) - Yn[None, :, :].astype(np.float64)) ** 2).sum(-1)
        r = np.sqrt(sq)
        p = float(f32(power))
        K = nconst * np.exp(-np.power(float(scale) * r, p))
        gs_rows = torch.from_numpy((-(A * K * p * float(scale) ** (p - 1.0) * np.power(r, p)).sum(1)).astype(f32))
    if want_power:   # dK/dp = -K u ln(s r) at fixed nconst, zero at r = 0
        sq = ((Xn[:, None, :].astype(np.float64) - Yn[None, :, :].astype(np.float64)) ** 2).sum(-1)
        sr = float(scale) * np.sqrt(sq)
        p = float(f32(power))
        with np.errstate(divide="ignore", invalid="ignore"):
            dk = np.where(sr > 0, -nconst * np.exp(-np.power(sr, p)) * np.power(sr, p) * np.log(sr), 0.0)
        return torch.from_numpy(gX.astype(f32)), gs_rows, torch.from_numpy((A * dk).sum(1).astype(f32))
    return torch.from_numpy(gX.astype(f32)), gs_rows


def fake_gram_grad(X, Y, Gbar, scale, power, nconst, want_scale=True, want_power=False, transposed=False):
    CALLS.append("gram_grad")
    if transposed:
        Gbar = Gbar.t()
    ones = torch.eye(Gbar.shape[1], dtype=torch.float32)
    return fake_kmatw_grad(X, Y, ones, Gbar, scale, power, nconst, want_scale, want_power)


def test_autograd_plumbing_on_cpu_with_oracle_backed_kernels(monkeypatch):
    """The differentiable route of the lowering (rkhs/_lowering.py: _reduced_gram_autograd over jaxrk_b200/autograd.py):
    which backward entry points are called with which operands, the chain rule through prefactors / LinearReduce maps /
    row reduces / nconst(scale).  The four native calls are oracle-backed stand-ins here; the result is compared with
    torch float64 autograd through the reference's op sequence."""
    from jaxrk_b200 import autograd as ag
    monkeypatch.setattr(nat, "kmatw_grad", fake_kmatw_grad)
    monkeypatch.setattr(nat, "gram_grad", fake_gram_grad)
    rng = np.random.RandomState(5)
    N, M, d, m = 40, 28, 3, 5
    Xn, Yn, Pn, An = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32), rng.rand(N).astype(f32), rng.randn(m, M).astype(f32)

    def run(dtype, use_api):
        X = torch.tensor(Xn, dtype=dtype, requires_grad=True)
        Y = torch.tensor(Yn, dtype=dtype, requires_grad=True)
        P = torch.tensor(Pn, dtype=dtype, requires_grad=True)
        A = torch.tensor(An, dtype=dtype, requires_grad=True)
        ls = torch.tensor(1.3, dtype=dtype, requires_grad=True)
        if use_api:
            k = GenGaussKernel.make_gauss(ls)
            t1 = inner(FiniteVec(k, X, [Prefactors(P)]), FiniteVec(k, Y, [LinearReduce(A)]))
            t2 = inner(FiniteVec(k, X, [BalancedRed(4, True)]), FiniteVec(k, Y, [Mean()]))
            t3 = k(X)
        else:
            s = 0.70710695 / ls
            def gram(a, b):
                sq = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
                return ag.gengauss_nconst(s, 2.0) * torch.exp(-(s * s) * sq)
            K = gram(X, Y)
            t1 = (P[:, None] * K) @ A.T
            t2 = K.reshape(N // 4, 4, M).mean(1).mean(1, keepdim=True)
            t3 = gram(X, X)
        loss = (t1 ** 2).sum() + t2.sum() + (t3 ** 2).sum()
        loss.backward()
        return [loss.item()] + [t.grad.detach().double().numpy() for t in (X, Y, P, A, ls)]

    want = run(torch.float64, False)
    CALLS.clear()
    got = run(torch.float32, True)
    assert "kmatw_grad" in CALLS and "gram_grad" in CALLS and "kmatw" in CALLS and "gram" in CALLS
    assert abs(got[0] - want[0]) <= 1e-4 * abs(want[0])
    for gv, wv, name in zip(got[1:], want[1:], ("X", "Y", "prefactors", "linear_map", "length_scale")):
        assert np.abs(gv - wv).max() <= 2e-4 * np.abs(wv).max() + 1e-6, (name, np.abs(gv - wv).max())


def test_sparse_reduce_is_folded_into_the_fused_product_when_small():
    """SparseReduce (jaxrk/reduce/lincomb.py:27-87) as a dense LinearReduce inside the fused K@W: same numbers as the
    reference's route (materialised Gram, then the gather-mean), without the Gram."""
    from jaxrk_b200.reduce import SparseReduce
    rng = np.random.RandomState(11)
    X, Y = rng.randn(30, 3).astype(f32), rng.randn(22, 3).astype(f32)
    idcs = [np.array([[0, 5, 7], [1, 2, 29]]), np.array([[3, 3], [4, 9], [10, 11]]), np.zeros((1, 0), dtype=np.int64)]
    k = GenGaussKernel.make_laplace(1.1)
    for average in (True, False):
        sr = SparseReduce(idcs, average)
        assert sr.num_rows() == 6
        G = k(X, Y)
        want = Reduce.apply(sr.reduce_first_ax(G), [Mean()], 1)                 # the reference's route
        np.testing.assert_allclose((sr.to_linear_map(30) @ G).numpy(), sr.reduce_first_ax(G).numpy(), rtol=1e-6, atol=1e-7)
        CALLS.clear()
        got = inner(FiniteVec(k, X, [sr]), FiniteVec(k, Y, [Mean()]))
        assert CALLS == ["kmatw"], CALLS                                         # fused, no Gram
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-5, atol=1e-6)
    # too large to hold densely: the generic route (Gram, then Reduce.apply)
    import jaxrk_b200.rkhs._lowering as lowmod
    old = lowmod.SPARSE_DENSE_LIMIT
    lowmod.SPARSE_DENSE_LIMIT = 10
    try:
        CALLS.clear()
        got2 = inner(FiniteVec(k, X, [SparseReduce(idcs, True)]), FiniteVec(k, Y, []))
        assert CALLS == ["gram"], CALLS
        np.testing.assert_allclose(got2.numpy(), SparseReduce(idcs, True).reduce_first_ax(k(X, Y)).numpy(), rtol=2e-5, atol=1e-6)
    finally:
        lowmod.SPARSE_DENSE_LIMIT = old


def test_center_and_block_reduce_lowering():
    """Center (jaxrk/reduce/base.py:183-192) and BlockReduce (jaxrk/reduce/lincomb.py:89-129, intended semantics) through
    the lowering: Center followed by a small linear map folds into the fused K@W; Center as the last element of a chain
    on a materialised Gram becomes ONE affine pass whose means are skinny fused calls -- both against the array-level
    semantics applied to the float64 Gram."""
    rng = np.random.RandomState(5)
    N, M, d = 24, 18, 3
    X, Y = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32)
    pp, qq = np.abs(rng.randn(N)).astype(f32) + 0.1, rng.randn(M).astype(f32)
    A = rng.randn(5, M).astype(f32)
    k = GenGaussKernel.make_gauss(1.3)
    K = torch.from_numpy(np.asarray(_truth_gram(X, Y, *k.kernel_args()[:1], k.kernel_args()[2], k.kernel_args()[3], None)).astype(f32))
    blk = BlockReduce([0, 5, 6, 18], average=True)
    cases = {
        "center | none": ([Center()], []),
        "pref, center | pref": ([Prefactors(pp), Center()], [Prefactors(qq)]),
        "center | center": ([Center()], [Center()]),
        "pref, center | pref, center, sum": ([Prefactors(pp), Center()], [Prefactors(qq), Center(), Sum()]),
        "center, bal | center, lin": ([Center(), BalancedRed(4, True)], [Center(), LinearReduce(A)]),
        "none | block": ([], [blk]),
        "center | block, center": ([Center()], [blk, Center()]),
        "bal | center, block": ([BalancedRed(3)], [Center(), blk]),
    }
    for name, (cx, cy) in cases.items():
        CALLS.clear()
        got = inner(FiniteVec(k, X, cx), FiniteVec(k, Y, cy))
        want = Reduce.apply(Reduce.apply(K, cx, 0), cy, 1)
        assert got.shape == want.shape, name
        assert torch.allclose(got, want, rtol=2e-5, atol=2e-6), (name, float((got - want).abs().max()))
        if name in ("center | none", "pref, center | pref", "center | center"):
            assert CALLS.count("gram") == 1 and CALLS.count("gram_affine") == 1, (name, CALLS)
        if name in ("pref, center | pref, center, sum", "center, bal | center, lin", "none | block", "bal | center, block"):
            assert "gram" not in CALLS, (name, CALLS)           # fused K@W: the Gram is never materialised
    b2 = BlockReduce.sum_from_block(np.array([3, 3, 7, 7, 7, 1]), mean=False)
    assert list(b2.block_boundaries) == [0, 2, 5, 6] and b2.new_len(6) == 3
    G6 = torch.from_numpy(rng.randn(6, 4).astype(f32))
    assert torch.allclose(b2(G6, 0), torch.stack([G6[0:2].sum(0), G6[2:5].sum(0), G6[5:6].sum(0)]), atol=1e-6)
    assert torch.allclose(b2(G6.T.contiguous(), 1), b2(G6, 0).T, atol=1e-6)


def test_rkhs_cdist_uses_one_reduced_gram_and_diagonals_without_their_grams():
    """dist(a, b) between vectors of mean embeddings (jaxrk/utilities/distances.py:28-42): one cross reduced Gram, the
    diagonals from inner(full=False) -- block sums per BalancedRed group / one fused K(X, X) A^T for a LinearReduce -- and
    one combining pass; against the reference's three-Gram form on the float64 Gram."""
    from jaxrk_b200.utilities import dist, rkhs_gram_cdist
    rng = np.random.RandomState(9)
    X, Y = rng.randn(24, 3).astype(f32), (rng.randn(20, 3) + 0.2).astype(f32)
    pp = (np.abs(rng.randn(24)) + 0.2).astype(f32)
    A = rng.randn(4, 20).astype(f32)
    k = GenGaussKernel.make_gauss(1.1)
    a = FiniteVec(k, X, [Prefactors(pp), BalancedRed(6, True)])
    b = FiniteVec(k, Y, [LinearReduce(A)])
    CALLS.clear()
    D = dist(a, b)
    # nothing materialised: one cross K@W, one Sum-reduced K@W per group of `a`, one K(Y, Y) A^T for `b` (the combining
    # pass is jrk_gram_affine_f32 on CUDA tensors, plain torch arithmetic on the 4 x 4 stand-in here)
    assert CALLS == ["kmatw"] * 6, CALLS
    args = k.kernel_args()
    Kxy = _truth_gram(X, Y, args[0], args[2], args[3], None) * pp.astype(np.float64)[:, None]
    Kxx = _truth_gram(X, None, args[0], args[2], args[3], None) * pp.astype(np.float64)[:, None] * pp.astype(np.float64)[None, :]
    Kyy = _truth_gram(Y, None, args[0], args[2], args[3], None)
    Gab = Kxy.reshape(4, 6, 20).mean(1) @ A.astype(np.float64).T
    Gaa = Kxx.reshape(4, 6, 4, 6).mean((1, 3))
    Gbb = A.astype(np.float64) @ Kyy @ A.astype(np.float64).T
    want = np.diag(Gaa)[:, None] + np.diag(Gbb)[None, :] - 2 * Gab
    np.testing.assert_allclose(D.numpy(), want, rtol=3e-5, atol=3e-6)
    three = rkhs_gram_cdist(a.inner(b), a.inner(), b.inner())
    np.testing.assert_allclose(D.numpy(), three.numpy(), rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(a.inner(full=False).numpy(), np.diag(Gaa), rtol=3e-5, atol=3e-6)

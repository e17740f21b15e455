"""The reference-shaped Python API (kern / reduce / rkhs) on the GPU, with the real
kernels, against (i) outputs of the reference's own code (tests/golden) and (ii) the
oracle.  Mirrors the reference's tests/test_kern.py, test_utilities.py, test_vectors.py
and test_operators.py for the hot path."""
import numpy as np
import pytest
import scipy.stats
import torch
from scipy.spatial.distance import cdist

from jaxrk_b200 import _native as nat
from jaxrk_b200.kern import GenGaussKernel
from jaxrk_b200.reduce import BalancedRed, LinearReduce, Mean, Prefactors, Sum
from jaxrk_b200.rkhs import Cdo, Cmo, Cov_inv, CovOp, FiniteOp, FiniteVec, _lowering, inner
from jaxrk_b200.utilities import dist, eucldist
from oracle import jaxrk_oracle as orc
from test_lowering_cpu import chains_x, chains_y
from test_oracle import G1_V1_E1, G1_V1_W1

pytestmark = pytest.mark.gpu
f32 = np.float32


def n(t):
    return t.detach().cpu().numpy()


def test_rbf_known_answers_like_reference_test_kern():
    """tests/test_kern.py:15-43 (reference tolerance atol=1e-1; met here to 1e-6)."""
    x = np.arange(8)[:, None]
    for ls in np.linspace(0.9, 2.0, 5).astype(f32):
        ks = [(GenGaussKernel.make_gauss(ls), scipy.stats.norm(0, ls)), (GenGaussKernel.make_laplace(ls), scipy.stats.laplace(0, ls))]
        for sh in np.linspace(0.9, 2.0, 5):
            ks.append((GenGaussKernel.make(ls, float(sh)), scipy.stats.gennorm(beta=float(sh), scale=ls)))
        for k, d in ks:
            g = n(k(x))
            np.testing.assert_allclose(g[0], d.pdf(x[:, 0]), atol=1e-6, rtol=2e-5)
            assert abs(float(np.asarray(k.var()).reshape(-1)[0]) - d.var()) < 1e-4
            assert (g == g.T).all()


def test_eucldist_like_reference_test_utilities():
    """tests/test_utilities.py:11-31."""
    rng = np.random.RandomState(0)
    a, b = rng.randn(100, 3), rng.randn(200, 3)
    for variant in ("simple", "extension"):
        for x, y in ((a, b), (a, a)):
            np.testing.assert_allclose(n(eucldist(x, y, power=1., variant=variant)), cdist(x, y), atol=1e-5)
            np.testing.assert_allclose(n(eucldist(x, y, power=2., variant=variant)), cdist(x, y) ** 2, atol=1e-5)
    np.testing.assert_allclose(n(dist(a, b, power=1.)), cdist(a, b), atol=1e-5)
    np.testing.assert_allclose(n(dist(a, None, power=2.)), cdist(a, a) ** 2, atol=1e-5)


def test_quickstart_notebook_golden_g1():
    k = GenGaussKernel.make_laplace(1.)
    pts = np.arange(6, dtype=f32)[:, None]
    v1 = FiniteVec(k, pts, [Prefactors(np.ones(6, f32))])
    w1 = FiniteVec(k, pts, [Prefactors(np.ones(6, f32) / 6), BalancedRed(3)])
    e1 = FiniteVec.construct_RKHS_Elem(k, pts, np.ones(6, f32))
    np.testing.assert_allclose(n(inner(v1, w1)), G1_V1_W1, rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(n(inner(v1, e1)), G1_V1_E1, rtol=2e-6, atol=1e-8)


def test_inner_chain_battery_with_real_kernels(golden):
    kg = GenGaussKernel.make(1.7, 1.5)
    P, Q = golden["ch_P"], golden["ch_Q"]
    cx, cy = chains_x(golden), chains_y(golden)
    n0 = nat.launch_count()
    for nx in cx:
        for ny in cy:
            got = n(inner(FiniteVec(kg, P, cx[nx]), FiniteVec(kg, Q, cy[ny])))
            want = golden[f"ch_inner__{nx}__{ny}"]
            assert got.shape == want.shape, (nx, ny)
            np.testing.assert_allclose(got, want, rtol=3e-5, atol=3e-6, err_msg=f"{nx}/{ny}")
    assert nat.launch_count() - n0 >= len(cx) * len(cy)
    for nx in cx:
        got = n(FiniteVec(kg, P, cx[nx]).inner())
        np.testing.assert_allclose(got, golden[f"ch_self__{nx}"], rtol=3e-5,
                                   atol=2e-4 if nx in ("none", "pref", "center", "bal2_center") else 2e-5, err_msg=nx)
    el = FiniteVec.construct_RKHS_Elem(kg, P, golden["ch_pp"])
    np.testing.assert_allclose(n(el(Q)), golden["ch_eval_elem_at_Q"], rtol=3e-5, atol=3e-6)


def test_vectors_spec_from_reference_test_vectors():
    """tests/test_vectors.py:29-30,38-44 (stale file, valid spec): inner with Prefactors is
    Gram * outer(p, p); Prefactors(1/2) + BalancedRed(2) are pairwise means."""
    rng = np.random.RandomState(1)
    k = GenGaussKernel.make_gauss(1.0)
    x = rng.randn(20, 3).astype(f32)
    p = rng.rand(20).astype(f32)
    G = orc.gengauss_gram_truth64(x, None, orc.simple_scaler_s(orc.gauss_length_scale(1.0)), 2.0)
    got = n(FiniteVec(k, x, [Prefactors(p)]).inner())
    np.testing.assert_allclose(got, G * np.outer(p, p), rtol=1e-5, atol=1e-6)
    el = FiniteVec.construct_RKHS_Elem(k, x, p)
    np.testing.assert_allclose(n(inner(FiniteVec(k, x), el)), (G * p[None, :]).sum(1, keepdims=True), rtol=1e-5, atol=1e-6)
    half = FiniteVec(k, x, [Prefactors(np.full(20, .5, f32)), BalancedRed(2)])
    np.testing.assert_allclose(n(half.inner()), G.reshape(10, 2, 10, 2).mean((1, 3)), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(n(FiniteVec(k, x).mean().inner(FiniteVec(k, x).sum())), G.mean(0, keepdims=True).sum(1, keepdims=True),
                               rtol=1e-5)


def test_operators_against_reference_outputs_gpu(golden):
    g = golden
    kx, ky = GenGaussKernel.make_gauss(0.9), GenGaussKernel.make_laplace(1.1)
    inp, outp, ref = FiniteVec(kx, g["op_X"]), FiniteVec(ky, g["op_Y"]), FiniteVec(ky, g["op_R"])
    op = FiniteOp(inp, outp, g["op_M"])
    res = op @ FiniteVec(kx, g["op_Xt"])
    # identity pinned by tests/test_operators.py:41: op @ v == matr @ inner(inp_feat, v)
    lin = torch.as_tensor(g["op_M"], device=res.reduce[-1].linear_map.device) @ inner(inp, FiniteVec(kx, g["op_Xt"]))
    assert torch.allclose(res.reduce[-1].linear_map, lin.T, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(n(res.reduce[-1].linear_map), g["op_matmul_vec_lin"], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(n(res(g["op_Yt"])), g["op_matmul_vec_eval"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(n((op @ g["op_Xt"])(g["op_Yt"])), g["op_matmul_arr_eval"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(n(Cov_inv(CovOp(inp), regul=0.1).matr), g["cov_inv_matr"], rtol=1e-3, atol=5e-2)
    cmo = Cmo(inp, outp, regul=0.1)
    np.testing.assert_allclose(n((cmo @ FiniteVec(kx, g["op_Xt"]))(g["op_Yt"])), g["cmo_apply_eval"], rtol=1e-3, atol=1e-3)
    cdo = Cdo(inp, outp, ref, regul=0.1)
    np.testing.assert_allclose(n((cdo @ FiniteVec(kx, g["op_Xt"]))(g["op_Yt"])), g["cdo_apply_eval"], rtol=2e-3, atol=2e-3)


def test_config4_shape_mean_embeddings_reduced():
    """BASELINE config 4 at reduced size: 64 RKHS vectors x 512 points, d = 32, Mean-type reduce."""
    rng = np.random.RandomState(0)
    groups, g, d = 64, 512, 32
    X = rng.randn(groups * g, d).astype(f32)
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    v = FiniteVec(k, X, [BalancedRed(g, average=True)])
    assert len(v) == groups
    n0 = nat.launch_count()
    G = v.inner()
    assert nat.launch_count() - n0 <= 3, "must not materialise the 32768 x 32768 Gram"
    assert tuple(G.shape) == (groups, groups) and torch.allclose(G, G.T, rtol=1e-5, atol=1e-9)
    s = orc.simple_scaler_s(orc.gauss_length_scale(float(np.sqrt(d))))
    t = orc.gengauss_gram_truth64(X[:g], X[g:3 * g], s, 2.0).reshape(1, g, 2, g).mean((1, 3))
    np.testing.assert_allclose(n(G)[:1, 1:3], t, rtol=1e-5)
    Y = rng.randn(4096, d).astype(f32)
    e = inner(v, FiniteVec(k, Y).mean())
    assert tuple(e.shape) == (groups, 1)
    t2 = orc.gengauss_gram_truth64(X[:2 * g], Y, s, 2.0).reshape(2, g, -1).mean((1, 2))
    np.testing.assert_allclose(n(e)[:2, 0], t2, rtol=1e-5)


def test_config5_shape_cdo_apply_reduced():
    """BASELINE config 5 at reduced size: Gram build feeding the library solve, then the
    operator applied to test points; checks the op@vec identity and density normalisation."""
    rng = np.random.RandomState(0)
    N, T = 512, 2048
    X = rng.randn(N, 2).astype(f32)
    Yv = (np.sin(X[:, :1]) + 0.1 * rng.randn(N, 1)).astype(f32)
    R = np.linspace(-2, 2, 64, dtype=f32)[:, None]
    kx, ky = GenGaussKernel.make_gauss(0.5), GenGaussKernel.make_gauss(0.3)
    inp, outp, ref = FiniteVec(kx, X), FiniteVec(ky, Yv), FiniteVec(ky, R)
    cdo = Cdo(inp, outp, ref, regul=0.05)
    Xt = rng.randn(T, 2).astype(f32)
    res = cdo @ FiniteVec(kx, Xt)
    assert len(res) == T
    # identity of tests/test_operators.py:41, op @ v == matr @ inner(inp_feat, v): here the left side is ONE fused
    # K(Xt, X) @ (matr R_inp)^T, the right side two float32 products -- equal to rounding of what was accumulated
    G = inner(cdo.inp_feat, FiniteVec(kx, Xt))
    lin = cdo.matr @ G
    bound = cdo.matr.abs() @ G.abs()
    assert ((res.reduce[-1].linear_map - lin.T).abs() <= 1e-6 + 2e-5 * bound.T).all()
    ev = res(R)
    assert tuple(ev.shape) == (T, 64) and torch.isfinite(ev).all()


def test_balanced_inner_through_the_tensor_core_route(monkeypatch):
    """BalancedRed on both vectors on a large problem: 16 column groups at a time through the tensor-core K@W kernel
    (rkhs/_lowering.py: _groupsum_tensor) -- against the FP32 group-sum kernel and the float64 oracle; the self inner
    product must stay exactly symmetric."""
    import numpy as np
    from jaxrk_b200 import _native as nat
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import BalancedRed, Prefactors
    from jaxrk_b200.rkhs import FiniteVec, inner
    from oracle import jaxrk_oracle as orc
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    N, M, d, gx, gy = 65536, 40960, 16, 512, 1024
    X, Y = torch.randn(N, d, generator=g, device=dev), torch.randn(M, d, generator=g, device=dev)
    P = torch.rand(N, generator=g, device=dev)
    k = GenGaussKernel.make_gauss(4.0)
    scale, _, power, nconst = k.kernel_args()
    s = orc.simple_scaler_s(orc.gauss_length_scale(4.0))
    vx = FiniteVec(k, X, [Prefactors(P), BalancedRed(gx, True)])
    vy = FiniteVec(k, Y, [BalancedRed(gy, False)])
    A = inner(vx, vy)
    S = FiniteVec(k, X, [BalancedRed(gy, True)]).inner()
    monkeypatch.setattr(_lowering, "GROUPSUM_TENSOR", False)
    A0 = inner(vx, vy)
    S0 = FiniteVec(k, X, [BalancedRed(gy, True)]).inner()
    assert tuple(A.shape) == (N // gx, M // gy) and tuple(S.shape) == (N // gy, N // gy)
    assert torch.allclose(A, A0, rtol=2e-5, atol=1e-6) and torch.allclose(S, S0, rtol=2e-5, atol=1e-7)
    assert torch.equal(S, S.T)
    for a, b in ((0, 0), (5, 39), (127, 17)):
        Xa, Yb = X[a * gx:(a + 1) * gx].cpu().numpy(), Y[b * gy:(b + 1) * gy].cpu().numpy()
        Pa = P[a * gx:(a + 1) * gx].cpu().numpy().astype(np.float64)
        truth = (orc.gengauss_gram_truth64(Xa, Yb, s, power, nconst=nconst) * Pa[:, None]).sum() / gx
        assert abs(float(A[a, b]) - truth) <= 1e-6 + 1e-5 * abs(truth)
    for a, b in ((3, 3), (40, 7)):
        Xa, Xb = X[a * gy:(a + 1) * gy].cpu().numpy(), X[b * gy:(b + 1) * gy].cpu().numpy()
        truth = orc.gengauss_gram_truth64(Xa, Xb, s, power, nconst=nconst).mean()
        assert abs(float(S[a, b]) - truth) <= 1e-6 + 1e-5 * abs(truth)


def test_gram_utilities_and_point_representant_on_the_gpu():
    """jaxrk/utilities/gram.py and FiniteVec.point_representant over Grams from the fused kernels, against outputs of
    the reference's own code (tests/golden/reference_siblings.npz)."""
    import os
    from jaxrk_b200.utilities import choose_representer, gram_projection, rkhs_gram_cdist, rkhs_gram_cdist_ignore_const
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_siblings.npz"))
    X, Y = g["X"], g["Y"]
    kq = GenGaussKernel.make_gauss(0.9)
    Gxx, Gyy, Gxy = kq(X), kq(Y), kq(X, Y)
    np.testing.assert_allclose(rkhs_gram_cdist(Gxy, Gxx, Gyy).cpu().numpy(), g["util_cdist_xy"], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(rkhs_gram_cdist_ignore_const(Gxy, Gyy).cpu().numpy(), g["util_cdist_ignore"], rtol=3e-5, atol=3e-6)
    assert choose_representer(X, g["util_factors"], kq) == int(g["util_representer"][0])
    assert (gram_projection(Gxy, Gyy, Gxx).cpu().numpy() == g["util_projection"]).all()
    vec = FiniteVec(kq, X, [LinearReduce(g["util_A2"])])
    np.testing.assert_allclose(vec.point_representant("inspace_point", keepdims=True).cpu().numpy(), g["util_point_repr"], rtol=1e-6)
    np.testing.assert_allclose(vec.point_representant("mean", keepdims=True).cpu().numpy(), g["util_point_mean"], rtol=3e-5, atol=1e-6)
    d = rkhs_gram_cdist(Gxx)
    assert torch.equal(d, d.T) and float(torch.diagonal(d).abs().max()) == 0.0


def test_gp_regression_on_the_gpu():
    """jaxrk/gp.py over the fused Gram kernels, against the reference's own outputs."""
    import os
    from jaxrk_b200.gp import GP
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_siblings.npz"))
    k = GenGaussKernel.make_gauss(0.8)
    n0 = nat.launch_count()
    gp = GP(FiniteVec(k, g["gp_xtr"], []), g["gp_ytr"], 0.1, 1.3, 0.2)
    mu, var = gp.predict(FiniteVec(k, g["gp_xte"], []))
    assert nat.launch_count() - n0 >= 3
    np.testing.assert_allclose(mu.cpu().numpy(), g["gp_mu"], rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose(var.cpu().numpy(), g["gp_var"], rtol=3e-3, atol=3e-4)
    assert abs(float(gp.marginal_likelihood()) - float(g["gp_ml"][0])) <= 3e-3 * abs(float(g["gp_ml"][0])) + 3e-3


def test_sparse_reduce_folded_into_the_fused_product_on_the_gpu():
    """SparseReduce as a dense LinearReduce inside the fused K@W (rkhs/_lowering.py) == the reference's route
    (materialised Gram, then the gather-mean)."""
    from jaxrk_b200.reduce import SparseReduce
    rng = np.random.RandomState(11)
    X, Y = rng.randn(3000, 5).astype(np.float32), rng.randn(2000, 5).astype(np.float32)
    idcs = [rng.randint(0, 3000, size=(40, 7)), rng.randint(0, 3000, size=(25, 3))]
    k = GenGaussKernel.make_laplace(1.4)
    sr = SparseReduce(idcs, True)
    n0 = nat.launch_count()
    got = inner(FiniteVec(k, X, [sr]), FiniteVec(k, Y, [Mean()]))
    assert nat.launch_count() - n0 <= 5                      # fused K@W (+ operand packing / finalize), no N x M Gram
    want = sr.reduce_first_ax(k(X, Y)).mean(1, keepdim=True)
    assert tuple(got.shape) == (65, 1)
    assert torch.allclose(got, want, rtol=2e-5, atol=1e-7)

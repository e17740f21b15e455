"""SURVEY.md section 8f-4 remainder: Center (jaxrk/reduce/base.py:183-192) and BlockReduce (jaxrk/reduce/lincomb.py:89-129)
through the kernels -- the one-pass affine epilogue on a materialised Gram (jrk_gram_affine_f32), Center folded into the
fused K@W when a small linear map follows, block sums folded or applied by prefix sums."""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from jaxrk_b200.kern import GenGaussKernel
from jaxrk_b200.reduce import BalancedRed, BlockReduce, Center, LinearReduce, Prefactors, Sum
from jaxrk_b200.rkhs import FiniteVec, inner
from oracle import jaxrk_oracle as orc
from gpu_util import cu, f32, gauss

pytestmark = pytest.mark.gpu


def n(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("N,M", [(1000, 2048), (513, 1023), (1, 7), (4096, 4096)])
def test_gram_affine_kernel(N, M):
    rng = np.random.RandomState(N + M)
    K = rng.randn(N, M).astype(f32)
    rm, cm, ra, ca = rng.randn(N).astype(f32), rng.randn(M).astype(f32), rng.randn(N).astype(f32), rng.randn(M).astype(f32)
    want = K.astype(np.float64) * rm[:, None] * cm[None, :] + ra[:, None] + ca[None, :] + 0.25
    Kd = cu(K)
    n0 = nat.launch_count()
    got = nat.gram_affine(Kd, cu(rm), cu(cm), cu(ra), cu(ca), 0.25)
    assert nat.launch_count() - n0 == 1 and got.data_ptr() == Kd.data_ptr()          # in place, one launch
    np.testing.assert_allclose(n(got), want, rtol=2e-6, atol=2e-6)
    out = torch.empty((N, M + 4), device=Kd.device)[:, :M]                            # strided output, some vectors absent
    got2 = nat.gram_affine(cu(K), None, cu(cm), cu(ra), None, 0.0, out=out)
    np.testing.assert_allclose(n(got2), K.astype(np.float64) * cm[None, :] + ra[:, None], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("kname", ["gauss", "laplace"])
def test_centered_gram_is_one_gram_plus_one_pass(kname):
    """C_N (D_p K D_q) C_M at a size where the Gram is the cost: two skinny fused calls for the means, one Gram, one affine
    pass -- against float64."""
    N, M, d = 3000, 2500, 5
    X, Y = gauss(0, N, d), gauss(1, M, d)
    p, q = np.abs(gauss(2, N)) + 0.1, gauss(3, M)
    k = GenGaussKernel.make_gauss(2.0) if kname == "gauss" else GenGaussKernel.make_laplace(2.0)
    s = orc.simple_scaler_s(orc.gauss_length_scale(2.0) if kname == "gauss" else 2.0)
    G = orc.gengauss_gram_truth64(X, Y, s, 2.0 if kname == "gauss" else 1.0) * p.astype(np.float64)[:, None] * q.astype(np.float64)[None, :]
    for cx, cy in ((True, True), (True, False), (False, True)):
        chx = [Prefactors(p)] + ([Center()] if cx else [])
        chy = [Prefactors(q)] + ([Center()] if cy else [])
        n0 = nat.launch_count()
        got = inner(FiniteVec(k, X, chx), FiniteVec(k, Y, chy))
        launches = nat.launch_count() - n0
        want = G
        if cx:
            want = want - want.mean(0, keepdims=True)
        if cy:
            want = want - want.mean(1, keepdims=True)
        bound = np.abs(G).max()
        assert np.abs(n(got) - want).max() <= 1e-6 + 2e-5 * bound, (kname, cx, cy, np.abs(n(got) - want).max(), bound)
        assert launches <= 12, launches          # (means: <= 5 launches each on the FP32-pipe kernel) + Gram + affine
    # the centred self-Gram (FiniteVec.centered(), jaxrk/rkhs/vector.py:107-108) is symmetric to rounding
    v = FiniteVec(k, X).centered()
    S = n(v.inner())
    assert np.abs(S - S.T).max() <= 1e-6 and abs(S.mean()) < 1e-6


def test_center_then_linear_map_is_folded_into_the_fused_kernel():
    N, M, d = 5000, 6000, 8
    X, Y = gauss(0, N, d), gauss(1, M, d)
    A = gauss(2, 6, M)
    k = GenGaussKernel.make_gauss(3.0)
    s = orc.simple_scaler_s(orc.gauss_length_scale(3.0))
    n0 = nat.launch_count()
    got = inner(FiniteVec(k, X, [Center(), BalancedRed(1000, True)]), FiniteVec(k, Y, [Center(), LinearReduce(A)]))
    assert nat.launch_count() - n0 <= 6, "no materialised Gram"
    G = orc.gengauss_gram_truth64(X, Y, s, 2.0)
    want = (G - G.mean(0, keepdims=True)).reshape(5, 1000, M).mean(1)
    want = (want - want.mean(1, keepdims=True)) @ A.astype(np.float64).T
    bound = np.abs(G).reshape(5, 1000, M).mean(1) @ np.abs(A.astype(np.float64)).T
    assert (np.abs(n(got) - want) <= 1e-6 + 2e-5 * bound).all()


def test_block_reduce_folded_and_generic():
    N, M, d = 4000, 3000, 4
    X, Y = gauss(0, N, d), gauss(1, M, d)
    k = GenGaussKernel.make_laplace(1.5)
    G = orc.gengauss_gram_truth64(X, Y, orc.simple_scaler_s(1.5), 1.0)
    few = BlockReduce([0, 700, 701, 3000], average=True)                      # folded into the fused K@W
    labels = np.repeat(np.arange(150), 20)                                    # 150 blocks: generic route (prefix sums)
    many = BlockReduce.sum_from_block(labels, mean=False)
    n0 = nat.launch_count()
    got = inner(FiniteVec(k, X, [Sum()]), FiniteVec(k, Y, [few]))
    assert nat.launch_count() - n0 <= 4
    want = np.stack([G[:, 0:700].mean(1), G[:, 700:701].mean(1), G[:, 701:3000].mean(1)], 1).sum(0, keepdims=True)
    np.testing.assert_allclose(n(got), want, rtol=2e-5)
    got2 = inner(FiniteVec(k, X), FiniteVec(k, Y, [many]))
    np.testing.assert_allclose(n(got2), G.reshape(N, 150, 20).sum(2), rtol=3e-5, atol=1e-6)
    assert many.new_len(M) == 150 and few.new_len(M) == 3


def test_rkhs_cdist_between_mean_embeddings_needs_one_reduced_gram():
    """dist(a, b) for two vectors of mean embeddings (jaxrk/utilities/distances.py:28-42): the cross reduced Gram, the two
    diagonals WITHOUT their reduced Grams (block sums), one combining pass -- against float64, and against the
    three-Gram form of the reference."""
    from jaxrk_b200.utilities import dist, rkhs_gram_cdist
    rng = np.random.RandomState(3)
    ga_, gb_, na, nb, d = 256, 128, 12, 20, 6
    X, Y = rng.randn(na * ga_, d).astype(f32), (rng.randn(nb * gb_, d) + 0.3).astype(f32)
    p = (np.abs(rng.randn(na * ga_)) + 0.2).astype(f32)
    k = GenGaussKernel.make_gauss(2.5)
    s = orc.simple_scaler_s(orc.gauss_length_scale(2.5))
    a = FiniteVec(k, X, [Prefactors(p), BalancedRed(ga_, True)])
    b = FiniteVec(k, Y, [BalancedRed(gb_, True)])
    n0 = nat.launch_count()
    D = dist(a, b)
    launches = nat.launch_count() - n0
    p64 = p.astype(np.float64)
    Gab = (orc.gengauss_gram_truth64(X, Y, s, 2.0) * p64[:, None]).reshape(na, ga_, nb, gb_).mean((1, 3))
    Gaa = (orc.gengauss_gram_truth64(X, X, s, 2.0) * p64[:, None] * p64[None, :]).reshape(na, ga_, na, ga_).mean((1, 3))
    Gbb = orc.gengauss_gram_truth64(Y, Y, s, 2.0).reshape(nb, gb_, nb, gb_).mean((1, 3))
    want = np.diag(Gaa)[:, None] + np.diag(Gbb)[None, :] - 2 * Gab
    scale_ = np.diag(Gaa).max() + np.diag(Gbb).max()
    assert np.abs(n(D) - want).max() <= 3e-5 * scale_
    three = rkhs_gram_cdist(a.inner(b), a.inner(), b.inner())                         # the reference's form
    assert float((D - three).abs().max()) <= 3e-5 * scale_
    assert launches <= 8 + 4 * (na + nb), launches
    # an element against itself: zero distance on the diagonal up to rounding, and the cached self inner product survives
    S0 = a.inner().clone()
    D2 = dist(a, a)
    assert float(torch.diagonal(D2).abs().max()) <= 3e-5 * scale_ and torch.equal(a.inner(), S0)
    # LinearReduce elements: diagonal by one fused K(X, X) A^T
    A = rng.randn(5, na * ga_).astype(f32) / 50
    c = FiniteVec(k, X, [LinearReduce(A)])
    dg = c.inner(full=False)
    full = FiniteVec(k, X, [LinearReduce(A)]).inner()
    assert torch.allclose(dg, torch.diagonal(full), rtol=3e-5, atol=1e-6)

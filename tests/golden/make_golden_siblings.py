"""Golden outputs of the reference's sibling kernels and diag=True branch, produced by running the REFERENCE'S OWN
PYTHON CODE over the numpy stand-in for jax of make_golden.py (build container only; never at test time):

    python tests/golden/make_golden_siblings.py   ->  tests/golden/reference_siblings.npz

Pins PeriodicKernel.__call__ (jaxrk/kern/rbf.py:152-154), ThreshSpikeKernel.__call__ (:203-206), the diag=True
branch of ScaledPairwiseDistance.__call__ (jaxrk/kern/util.py:108-113) as used by GenGaussKernel / PeriodicKernel,
and FiniteVec.inner (jaxrk/rkhs/vector.py:62-75) over a PeriodicKernel with reduce chains.
"""
import os
import sys

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import F32, import_reference  # noqa: E402


def main():
    import_reference()
    from jaxrk.kern import GenGaussKernel, PeriodicKernel, ThreshSpikeKernel
    from jaxrk.kern.util import ScaledPairwiseDistance, SimpleScaler
    from jaxrk.reduce import BalancedRed, LinearReduce, Mean, Prefactors, Sum
    from jaxrk.rkhs import FiniteVec, inner

    rng = onp.random.RandomState(20261020)
    out = {}
    X = (rng.randn(37, 3) * 1.5).astype(F32)
    Y = (rng.randn(29, 3) * 1.5).astype(F32)
    Y[:5] = X[:5]                       # coincident points: distance exactly 0
    X1 = onp.sort(rng.rand(24, 1).astype(F32) * 6, 0)
    out.update(X=X, Y=Y, X1=X1)

    # ---- PeriodicKernel ----
    pk_params = [(1.0, 1.0), (2.5, 0.7), (0.8, 2.0)]
    out["periodic_params"] = onp.array(pk_params, F32)
    for i, (period, ls) in enumerate(pk_params):
        k = PeriodicKernel(period, ls)
        out[f"periodic_{i}_XY"] = onp.asarray(k(X, Y), F32)
        out[f"periodic_{i}_XX"] = onp.asarray(k(X), F32)
        out[f"periodic_{i}_X1"] = onp.asarray(k(X1), F32)
        out[f"periodic_{i}_diag"] = onp.asarray(k(X[:29], Y, diag=True), F32)

    # ---- ThreshSpikeKernel ----
    ts_params = [(1.0, 2.0, 1.0, 0.25, 2.0), (0.5, 1.0, 2.0, -0.5, 1.0), (1.3, 1.5, 0.7, 0.0, 0.0)]  # ls, shape, spike, non_spike, thr
    out["thresh_params"] = onp.array(ts_params, F32)
    for i, (ls, shape, spike, non_spike, thr) in enumerate(ts_params):
        d = ScaledPairwiseDistance(scaler=SimpleScaler(1. / ls), power=shape)
        k = ThreshSpikeKernel(d, spike, non_spike, thr)
        out[f"thresh_{i}_XY"] = onp.asarray(k(X, Y), F32)
        out[f"thresh_{i}_XX"] = onp.asarray(k(X), F32)
        out[f"thresh_{i}_dist_XY"] = onp.asarray(d(X, Y), F32)   # how close each pair is to the threshold

    # ---- diag=True of the distance and of GenGaussKernel ----
    dg_params = [(1.0, 2.0), (0.7, 1.0), (2.0, 1.5)]
    out["diag_params"] = onp.array(dg_params, F32)
    for i, (ls, shape) in enumerate(dg_params):
        d = ScaledPairwiseDistance(scaler=SimpleScaler(1. / ls), power=shape)
        out[f"diag_{i}_dist"] = onp.asarray(d(X[:29], Y, diag=True), F32)
        out[f"diag_{i}_dist_self"] = onp.asarray(d(X, None, diag=True), F32)
        k = GenGaussKernel.make(ls, shape)
        out[f"diag_{i}_gengauss"] = onp.asarray(k(X[:29], Y, diag=True), F32)
    dsc = onp.array([0.5, 1.0, 2.0], F32)
    d = ScaledPairwiseDistance(scaler=SimpleScaler(dsc), power=1.5)
    out["diag_dimscale"] = dsc
    out["diag_dimscale_dist"] = onp.asarray(d(X[:29], Y, diag=True), F32)

    # ---- inner products over a PeriodicKernel ----
    k = PeriodicKernel(2.5, 0.7)
    pref = rng.rand(37).astype(F32)
    A = rng.randn(4, 29).astype(F32)
    out.update(inner_pref=pref, inner_A=A)
    out["inner_periodic_plain"] = onp.asarray(inner(FiniteVec(k, X, []), FiniteVec(k, Y, [])), F32)
    out["inner_periodic_lin"] = onp.asarray(inner(FiniteVec(k, X, [Prefactors(pref)]), FiniteVec(k, Y, [LinearReduce(A)])), F32)
    out["inner_periodic_mean_sum"] = onp.asarray(inner(FiniteVec(k, X, [Mean()]), FiniteVec(k, Y, [Sum()])), F32)
    Xb = X[:36]
    out["inner_periodic_balanced"] = onp.asarray(inner(FiniteVec(k, Xb, [BalancedRed(4, True)]), FiniteVec(k, Y, [LinearReduce(A)])), F32)

    # ---- CombVec (jaxrk/rkhs/vector.py:227-268): product of two vectors' Grams, then a reduce chain ----
    import operator
    from jaxrk.rkhs import CombVec
    kg1, kg2, kl = GenGaussKernel.make_gauss(1.3), GenGaussKernel.make_gauss(0.8), GenGaussKernel.make_laplace(1.1)
    XR, XL = (rng.randn(24, 2)).astype(F32), (rng.randn(24, 3)).astype(F32)
    YR, YL = (rng.randn(18, 2)).astype(F32), (rng.randn(18, 3)).astype(F32)
    B = rng.randn(5, 18).astype(F32)
    out.update(comb_XR=XR, comb_XL=XL, comb_YR=YR, comb_YL=YL, comb_B=B)
    for name, (ka, kb) in {"gg": (kg1, kg2), "gl": (kg1, kl)}.items():
        cx = CombVec(FiniteVec(ka, XR, []), FiniteVec(kb, XL, []), operator.mul, [BalancedRed(3, True)])
        cy = CombVec(FiniteVec(ka, YR, []), FiniteVec(kb, YL, []), operator.mul, [LinearReduce(B)])
        out[f"comb_{name}_xy"] = onp.asarray(cx.inner(cy), F32)
        out[f"comb_{name}_xx"] = onp.asarray(cx.inner(), F32)
        out[f"comb_{name}_len"] = onp.array([len(cx), len(cy)])

    # ---- Gram-based utilities (jaxrk/utilities/gram.py) and FiniteVec.point_representant (rkhs/vector.py:161-174) ----
    from jaxrk.utilities.gram import (rkhs_gram_cdist, rkhs_gram_cdist_ignore_const, choose_representer,
                                      choose_representer_from_gram, gram_projection)
    kq = GenGaussKernel.make_gauss(0.9)
    Gxx, Gyy, Gxy = kq(X), kq(Y), kq(X, Y)
    out["util_cdist_xy"] = onp.asarray(rkhs_gram_cdist(Gxy, Gxx, Gyy), F32)
    out["util_cdist_xy_p1"] = onp.asarray(rkhs_gram_cdist(Gxy, Gxx, Gyy, power=1.), F32)
    out["util_cdist_ignore"] = onp.asarray(rkhs_gram_cdist_ignore_const(Gxy, Gyy), F32)
    fac = rng.rand(37).astype(F32)
    fac /= fac.sum()
    out["util_factors"] = fac
    out["util_representer"] = onp.array([int(choose_representer(X, fac, kq)), int(choose_representer_from_gram(Gxx, fac))])
    out["util_projection"] = onp.asarray(gram_projection(Gxy, Gyy, Gxx), onp.int64)   # 37 representers x 29 originals
    A2 = rng.rand(3, 37).astype(F32)
    A2 /= A2.sum(1, keepdims=True)
    out["util_A2"] = A2
    vec = FiniteVec(kq, X, [LinearReduce(A2)])
    out["util_point_repr"] = onp.asarray(vec.point_representant("inspace_point", keepdims=True), F32)
    out["util_point_mean"] = onp.asarray(vec.point_representant("mean", keepdims=True), F32)

    # ---- GP regression (jaxrk/gp.py) over FiniteVec.inner ----
    from jaxrk.gp import GP
    kgp = GenGaussKernel.make_gauss(0.8)
    xtr = onp.linspace(0., 6., 24).reshape(-1, 1).astype(F32)
    ytr = (onp.sin(xtr) + 0.1 * rng.randn(24, 1)).astype(F32)
    xte = onp.linspace(0.3, 5.7, 10).reshape(-1, 1).astype(F32)
    yte = onp.sin(xte).astype(F32)
    out.update(gp_xtr=xtr, gp_ytr=ytr.copy(), gp_xte=xte, gp_yte=yte)
    gp = GP(FiniteVec(kgp, xtr, []), ytr.copy(), 0.1, 1.3, 0.2)
    mu, var = gp.predict(FiniteVec(kgp, xte, []))
    out["gp_mu"], out["gp_var"] = onp.asarray(mu, F32), onp.asarray(var, F32)
    out["gp_ml"] = onp.asarray(gp.marginal_likelihood(), onp.float64).reshape(1)
    out["gp_ppl"] = onp.asarray(gp.post_pred_likelihood(FiniteVec(kgp, xte, []), yte)[2], onp.float64).reshape(1)

    path = os.path.join(HERE, "reference_siblings.npz")
    onp.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()

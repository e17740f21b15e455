"""Fused backward kernels (csrc/kgrad.cu: jrk_kmatw_grad_f32 / jrk_gram_grad_f32) and the torch.autograd functions
over them (jaxrk_b200/autograd.py) against the float64 analytic oracle.  SURVEY.md §8f rank 2."""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from jaxrk_b200 import autograd as ag
from oracle import jaxrk_oracle as orc
from gpu_util import cu, f32, gauss, kparams

pytestmark = pytest.mark.gpu
RT, AT = 2e-5, 2e-6      # gradient sums are judged against the magnitude of what was accumulated


def _check(got, want, mag, what):
    got = np.asarray(got, np.float64)
    err = np.abs(got - want)
    lim = AT + RT * mag
    assert (err <= lim).all(), f"{what}: worst err {err.max():.3e}, limit there {np.broadcast_to(lim, err.shape).flat[err.argmax()]:.3e}"


@pytest.mark.parametrize("kname", ["sqexp", "laplace", "gg15", "gg05"])
@pytest.mark.parametrize("N,M,d,m", [(300, 517, 16, 16), (129, 64, 3, 1), (1000, 33, 40, 5), (64, 2100, 8, 32)])
def test_kmatw_grad_against_oracle(kname, N, M, d, m):
    X, Y, W, G = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m), gauss(3, N, m)
    s, scale, p, nconst = kparams(kname, np.sqrt(d) * 0.9)
    A = G.astype(np.float64) @ W.astype(np.float64).T
    gX, gY, gs, (mx, my, ms) = orc.gengauss_vjp_truth64(X, Y, A, s, p, nconst=nconst)
    # |a_ij| <= sum_c |G_ic W_jc|: the accumulated magnitude of the fused kernel includes the m-term dot product
    Aabs = np.abs(G).astype(np.float64) @ np.abs(W).astype(np.float64).T
    _, _, _, (mxa, mya, msa) = orc.gengauss_vjp_truth64(X, Y, Aabs, s, p, nconst=nconst)
    n0 = nat.launch_count()
    gx_d, gs_rows = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, p, nconst)
    assert nat.launch_count() - n0 == 2            # fused pass + fixed-order finalize
    _check(gx_d.cpu().numpy(), gX, mxa, f"dL/dX {kname}")
    _check(float(gs_rows.double().sum()), gs, msa, f"dL/dscale {kname}")
    gy_d, _ = nat.kmatw_grad(cu(Y), cu(X), cu(G), cu(W), scale, p, nconst, want_scale=False)
    _check(gy_d.cpu().numpy(), gY, mya, f"dL/dY {kname}")


@pytest.mark.parametrize("kname", ["sqexp", "laplace", "gg15"])
@pytest.mark.parametrize("N,M,d", [(300, 517, 16), (130, 70, 2), (257, 1025, 64)])
def test_gram_grad_against_oracle(kname, N, M, d):
    X, Y, A = gauss(0, N, d), gauss(1, M, d), gauss(4, N, M)
    s, scale, p, nconst = kparams(kname, np.sqrt(d) * 1.1)
    gX, gY, gs, (mx, my, ms) = orc.gengauss_vjp_truth64(X, Y, A, s, p, nconst=nconst)
    gx_d, gs_rows = nat.gram_grad(cu(X), cu(Y), cu(A), scale, p, nconst)
    _check(gx_d.cpu().numpy(), gX, mx, f"gram dL/dX {kname}")
    _check(float(gs_rows.double().sum()), gs, ms, f"gram dL/dscale {kname}")
    gy_d, _ = nat.gram_grad(cu(Y), cu(X), cu(A.T.copy()), scale, p, nconst, want_scale=False)
    _check(gy_d.cpu().numpy(), gY, my, f"gram dL/dY {kname}")


def test_gram_grad_transposed_cotangent_fp32_kernel():
    """gbar_transposed on the FP32-pipe kernel (small / ineligible problems): identical sums, another read order."""
    N, M, d = 257, 1025, 9
    X, Y, A = gauss(0, N, d), gauss(1, M, d), gauss(4, N, M)
    for kname in ("sqexp", "laplace"):
        s, scale, p, nconst = kparams(kname, 3.0)
        g0, s0 = nat.gram_grad(cu(X), cu(Y), cu(A), scale, p, nconst)
        g1, s1 = nat.gram_grad(cu(X), cu(Y), cu(A.T.copy()), scale, p, nconst, transposed=True)
        assert torch.allclose(g0, g1, rtol=1e-6, atol=1e-7) and torch.allclose(s0, s1, rtol=1e-6, atol=1e-7)


def test_self_gram_with_coincident_points_has_finite_gradients():
    X = gauss(5, 200, 6)
    A = gauss(6, 200, 200)
    for kname in ("laplace", "gg05", "sqexp"):
        s, scale, p, nconst = kparams(kname, 2.0)
        gX, gY, gs, (mx, my, ms) = orc.gengauss_vjp_truth64(X, X, A, s, p, nconst=nconst)
        gx_d, gs_rows = nat.gram_grad(cu(X), cu(X), cu(A), scale, p, nconst)
        assert torch.isfinite(gx_d).all() and torch.isfinite(gs_rows).all()
        _check(gx_d.cpu().numpy(), gX, mx, f"self gram {kname}")


@pytest.mark.parametrize("power", [2.0, 1.0, 1.5])
def test_autograd_functions_match_float64_autograd(power):
    """End to end: loss(K) and loss(K @ W) differentiated with respect to X, Y, W and the length scale, against torch
    float64 autograd through the reference's op sequence."""
    N, M, d, m = 400, 300, 5, 40          # m > 32: the backward runs in two column blocks
    Xn, Yn, Wn, Cn = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m), gauss(3, N, m)
    ls0 = 1.7

    def run(dtype, dev, gram_fn, kmatw_fn):
        X = torch.tensor(Xn, dtype=dtype, device=dev, requires_grad=True)
        Y = torch.tensor(Yn, dtype=dtype, device=dev, requires_grad=True)
        W = torch.tensor(Wn, dtype=dtype, device=dev, requires_grad=True)
        ls = torch.tensor(ls0, dtype=dtype, device=dev, requires_grad=True)
        C = torch.tensor(Cn, dtype=dtype, device=dev)
        scale = 1.0 / ls
        loss = (kmatw_fn(X, Y, W, scale) * C).sum() + (gram_fn(X, Y, scale) ** 2).sum() + gram_fn(X, None, scale).sum()
        loss.backward()
        return [loss.item()] + [t.grad.detach().double().cpu().numpy() for t in (X, Y, W, ls)]

    def ref_gram(X, Y, scale):
        Y = X if Y is None else Y
        sq = ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1)
        r = torch.where(sq > 0, torch.sqrt(torch.where(sq > 0, sq, torch.ones_like(sq))), torch.zeros_like(sq))
        return ag.gengauss_nconst(scale, power) * torch.exp(-(scale * r) ** power)

    want = run(torch.float64, "cpu", ref_gram, lambda X, Y, W, s: ref_gram(X, Y, s) @ W)
    got = run(torch.float32, "cuda", lambda X, Y, s: ag.gengauss_gram(X, Y, s, power),
              lambda X, Y, W, s: ag.gengauss_kmatw(X, Y, W, s, power))
    assert abs(got[0] - want[0]) <= 1e-5 * abs(want[0]) + 1e-4
    for g, w, name in zip(got[1:], want[1:], ("X", "Y", "W", "length_scale")):
        scale_ = np.abs(w).max()
        assert np.abs(g - w).max() <= 1e-4 * scale_ + 1e-5, (name, np.abs(g - w).max(), scale_)


def test_grad_entry_points_validate_arguments():
    a = cu(gauss(0, 8, 2))
    with pytest.raises(nat.JrkError) as ei:
        nat.kmatw_grad(cu(gauss(0, 8, 70)), cu(gauss(1, 8, 70)), cu(gauss(2, 8, 2)), cu(gauss(3, 8, 2)), 1.0, 2.0, 1.0)
    assert ei.value.code == 3 and "d = 70" in str(ei.value)
    with pytest.raises(nat.JrkError) as ei:
        nat.kmatw_grad(a, a, cu(gauss(2, 8, 40)), cu(gauss(3, 8, 40)), 1.0, 2.0, 1.0)
    assert ei.value.code == 3
    with pytest.raises(nat.JrkError):
        nat.gram_grad(a, a, cu(gauss(4, 8, 8)), 1.0, 2.5, 1.0)


def test_api_level_gradients_through_finitevec_inner():
    """jax.grad-style use of the reference-shaped API (examples/Using_FLAX_Optax.ipynb cells 6-10): a loss over
    FiniteVec.inner / GenGaussKernel.__call__ differentiated with respect to the length scale, the points, prefactors
    and a LinearReduce map -- against torch float64 autograd through the reference op sequence."""
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import BalancedRed, LinearReduce, Mean, Prefactors
    from jaxrk_b200.rkhs import FiniteVec, inner
    N, M, d, m = 240, 130, 3, 6
    Xn, Yn, Pn, An = gauss(0, N, d), gauss(1, M, d), np.abs(gauss(2, N)), gauss(3, m, M)

    def loss_api(dtype, dev, use_api):
        X = torch.tensor(Xn, dtype=dtype, device=dev, requires_grad=True)
        Y = torch.tensor(Yn, dtype=dtype, device=dev, requires_grad=True)
        P = torch.tensor(Pn, dtype=dtype, device=dev, requires_grad=True)
        A = torch.tensor(An, dtype=dtype, device=dev, requires_grad=True)
        ls = torch.tensor(1.3, dtype=dtype, device=dev, requires_grad=True)
        if use_api:
            k = GenGaussKernel.make_gauss(ls)
            t1 = inner(FiniteVec(k, X, [Prefactors(P)]), FiniteVec(k, Y, [LinearReduce(A)]))          # N x m  (K@W)
            t2 = inner(FiniteVec(k, X, [BalancedRed(4, True)]), FiniteVec(k, Y, [Mean()]))            # N/4 x 1
            t3 = k(X)                                                                                  # self-Gram
        else:
            f = 0.70710695
            s = f / ls
            def gram(a, b):
                sq = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
                return ag.gengauss_nconst(s, 2.0) * torch.exp(-(s * s) * sq)
            K = gram(X, Y)
            t1 = (P[:, None] * K) @ A.T
            t2 = K.reshape(N // 4, 4, M).mean(1).mean(1, keepdim=True)
            t3 = gram(X, X)
        loss = (t1 ** 2).sum() + t2.sum() + (t3 ** 2).sum()
        loss.backward()
        return [loss.item()] + [t.grad.detach().double().cpu().numpy() for t in (X, Y, P, A, ls)]

    want = loss_api(torch.float64, "cpu", False)
    n0 = nat.launch_count()
    got = loss_api(torch.float32, "cuda", True)
    assert nat.launch_count() - n0 >= 8          # forward and backward both ran in libjaxrk_b200.so
    assert abs(got[0] - want[0]) <= 2e-5 * abs(want[0])
    for g, w, name in zip(got[1:], want[1:], ("X", "Y", "prefactors", "linear_map", "length_scale")):
        scale_ = np.abs(w).max()
        assert np.abs(g - w).max() <= 1e-4 * scale_ + 1e-6, (name, np.abs(g - w).max(), scale_)


def test_no_grad_path_is_unchanged():
    from jaxrk_b200.kern import GenGaussKernel
    X = cu(gauss(0, 100, 4))
    k = GenGaussKernel.make_gauss(1.1)
    K0 = k(X)
    Xg = X.clone().requires_grad_(True)
    with torch.no_grad():
        K1 = k(Xg)
    K2 = k(Xg)
    assert not K1.requires_grad and K2.requires_grad
    assert torch.equal(K0, K1) and torch.allclose(K0, K2, rtol=1e-6, atol=1e-8)


def test_gradients_are_never_dropped_silently():
    """where no fused backward kernel exists (d > 64, per-dimension scales, sibling kernels, diag=True) a
    request for gradients raises instead of returning a tensor without a graph."""
    from jaxrk_b200.kern import GenGaussKernel, PeriodicKernel
    from jaxrk_b200.reduce import LinearReduce, Prefactors
    from jaxrk_b200.rkhs import FiniteVec, inner
    from jaxrk_b200.rkhs._lowering import GradUnsupported
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    X80 = torch.randn(64, 80, generator=g, device=dev, requires_grad=True)
    Y80 = torch.randn(50, 80, generator=g, device=dev)
    k = GenGaussKernel.make_gauss(9.0)
    with pytest.raises(GradUnsupported):
        k(X80, Y80)
    with pytest.raises(GradUnsupported):
        inner(FiniteVec(k, X80), FiniteVec(k, Y80, [LinearReduce(torch.randn(3, 50, device=dev))]))
    with torch.no_grad():                                      # the value itself is available
        assert tuple(k(X80, Y80).shape) == (64, 50)
    assert tuple(k(X80.detach(), Y80).shape) == (64, 50)
    X4 = torch.randn(64, 4, generator=g, device=dev, requires_grad=True)
    Y4 = torch.randn(50, 4, generator=g, device=dev)
    per = PeriodicKernel(2.0, 1.0)
    with pytest.raises(GradUnsupported):
        inner(FiniteVec(per, X4), FiniteVec(per, Y4, [Prefactors(torch.rand(50, device=dev))]).sum())
    with pytest.raises(GradUnsupported):
        k(X4, X4, diag=True)
    # a chain parameter that requires grad counts as well
    A = torch.randn(3, 50, device=dev, requires_grad=True)
    with pytest.raises(GradUnsupported):
        inner(FiniteVec(k, X80.detach()), FiniteVec(k, Y80, [LinearReduce(A)]))
    # ... and inside the window the gradient reaches it
    out = inner(FiniteVec(k, X4.detach()), FiniteVec(k, Y4, [LinearReduce(A)]))
    out.sum().backward()
    assert A.grad is not None and bool(torch.isfinite(A.grad).all())


# ---- the tensor-core backward (csrc/kgrad_tc.cu): p = 2, d <= 16, m <= 16, N, M >= 4096, factored regime ----
def _tc_grad_case(N, M, d, m, ls, seed, positive):
    rng = np.random.RandomState(seed)
    X, Y = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32)
    W, G = rng.randn(M, m).astype(f32), rng.randn(N, m).astype(f32)
    if positive:
        W, G = np.abs(W), np.abs(G)
    s = orc.simple_scaler_s(orc.gauss_length_scale(ls))
    return X, Y, W, G, s, float(s[0, 0]), float(orc.gengauss_nconst_ref32(s, 2.0).reshape(-1)[0])


@pytest.mark.parametrize("N,M,d,m,ls,positive", [(8192, 8192, 16, 16, 4.0, False), (4173, 4109, 7, 5, 2.5, False),
                                                 (8192, 4096, 16, 1, 4.0, True), (4224, 6000, 3, 16, 1.7, False),
                                                 (4096 + 255, 4096 + 31, 16, 16, 3.0, True)])
def test_kmatw_grad_tensor_core_path(N, M, d, m, ls, positive):
    """dX, dY and dscale of the tensor-core kernel on sampled rows (first / last, so that both row blocks of a unit and the
    ragged last unit / tile are covered) against the float64 oracle, with the FP32-pipe kernel's own tolerance."""
    X, Y, W, G, s, scale, nconst = _tc_grad_case(N, M, d, m, ls, N % 7, positive)
    n0 = nat.launch_count()
    gx, gs = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, 2.0, nconst)
    assert nat.launch_count() - n0 == 6      # 2 row-norm kernels, images, tensor-core kernel, (skipped) FP32 kernel + finalize
    gy, _ = nat.kmatw_grad(cu(Y), cu(X), cu(G), cu(W), scale, 2.0, nconst, want_scale=False)
    rows = np.r_[0:40, 120:136, 250:262, N - 40:N]
    A = G[rows].astype(np.float64) @ W.astype(np.float64).T
    Aabs = np.abs(G[rows]).astype(np.float64) @ np.abs(W).astype(np.float64).T
    gX, _, gsum, _ = orc.gengauss_vjp_truth64(X[rows], Y, A, s, 2.0, nconst=nconst)
    _, _, _, (mx, _, ms) = orc.gengauss_vjp_truth64(X[rows], Y, Aabs, s, 2.0, nconst=nconst)
    _check(gx.cpu().numpy()[rows], gX, mx, "tensor-core dL/dX")
    _check(float(gs.double()[torch.from_numpy(rows).to(gs.device)].sum()), gsum, ms, "tensor-core dL/dscale")
    cols = np.r_[0:40, M - 40:M]
    A2 = G.astype(np.float64) @ W[cols].astype(np.float64).T
    A2abs = np.abs(G).astype(np.float64) @ np.abs(W[cols]).astype(np.float64).T
    _, gY, _, _ = orc.gengauss_vjp_truth64(X, Y[cols], A2, s, 2.0, nconst=nconst)
    _, _, _, (_, my, _) = orc.gengauss_vjp_truth64(X, Y[cols], A2abs, s, 2.0, nconst=nconst)
    _check(gy.cpu().numpy()[cols], gY, my, "tensor-core dL/dY")
    # every row against the FP32-pipe kernel (same tolerance doubled: two approximations of the same number)
    with nat.option(nat.OPT_KMATW_TC, 0):
        gx0, gs0 = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, 2.0, nconst)
    lim = 2 * (AT + RT * torch.from_numpy(mx.max(axis=0)).float().to(gx.device).max())
    assert float((gx - gx0).abs().max()) <= float(lim) * 50      # loose global bound: sampled rows carry the real check
    assert torch.isfinite(gx).all() and torch.isfinite(gs).all()


def test_kmatw_grad_outside_the_factored_regime_is_the_fp32_kernel():
    """Short length scale: |c| (max|x|^2 + max|y|^2) > 12, decided on the device; the tensor-core kernel returns at once and
    the FP32-pipe kernel writes the result -- bit-identical to the run with the tensor path switched off."""
    X, Y, W, G, s, scale, nconst = _tc_grad_case(8192, 4096, 16, 16, 0.5, 4, False)
    gx, gs = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, 2.0, nconst)
    with nat.option(nat.OPT_KMATW_TC, 0):
        gx0, gs0 = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, 2.0, nconst)
    assert torch.equal(gx, gx0) and torch.equal(gs, gs0)


@pytest.mark.parametrize("m", [8, 40])
def test_inner_backward_through_the_api_on_the_tensor_core_path(m):
    """loss = sum(inner(FiniteVec(X), FiniteVec(Y, LinearReduce(W^T))) * C) at a size where forward AND backward run on the
    tensor cores, against float64 autograd of the reference's op sequence on sampled rows."""
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import LinearReduce
    from jaxrk_b200.rkhs import FiniteVec, inner
    N, M, d = 8192, 4096, 16
    rng = np.random.RandomState(11)
    X, Y = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32)
    W, C = rng.randn(m, M).astype(f32), rng.randn(N, m).astype(f32)
    Xt = cu(X).requires_grad_(True)
    Yt = cu(Y).requires_grad_(True)
    Wt = cu(W).requires_grad_(True)
    k = GenGaussKernel.make_gauss(4.0)
    out = inner(FiniteVec(k, Xt), FiniteVec(k, Yt, [LinearReduce(Wt)]))
    (out * cu(C)).sum().backward()
    s = orc.simple_scaler_s(orc.gauss_length_scale(4.0))
    nconst = float(orc.gengauss_nconst_ref32(s, 2.0).reshape(-1)[0])
    rows = np.r_[0:32, N - 32:N]
    A = C[rows].astype(np.float64) @ W.astype(np.float64)
    Aabs = np.abs(C[rows]).astype(np.float64) @ np.abs(W).astype(np.float64)
    gX, _, _, _ = orc.gengauss_vjp_truth64(X[rows], Y, A, s, 2.0, nconst=nconst)
    _, _, _, (mx, _, _) = orc.gengauss_vjp_truth64(X[rows], Y, Aabs, s, 2.0, nconst=nconst)
    _check(Xt.grad.cpu().numpy()[rows], gX, mx, "API dL/dX")
    K = orc.gengauss_gram_truth64(X, Y[:64], s, 2.0)                   # dW[c, j] = sum_i C[i, c] K_ij
    _check(Wt.grad.cpu().numpy()[:, :64], C.astype(np.float64).T @ K, np.abs(C).astype(np.float64).T @ np.abs(K), "API dL/dW")
    assert Yt.grad is not None and torch.isfinite(Yt.grad).all()


# ---- gradients with respect to the shape parameter and per-dimension length scales (SURVEY.md section 8f-2 remainder) ----
@pytest.mark.parametrize("power", [2.0, 1.0, 1.5, 0.7])
def test_power_gradient_of_the_fused_backward_kernels(power):
    """gp_rows of jrk_kmatw_grad_f32 / jrk_gram_grad_f32: sum_j a_ij dK_ij/dpower at fixed nconst, dK/dp = -K u ln(s r),
    against float64; and asking for it (general-p epilogue) leaves gX / gs_rows within tolerance of the float64 truth."""
    N, M, d, m = 300, 517, 5, 7
    X, Y, W, G = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m), gauss(3, N, m)
    s = orc.simple_scaler_s(2.0)
    scale = float(s[0, 0])
    nconst = float(orc.gengauss_nconst_ref32(s, power).reshape(-1)[0])
    A = G.astype(np.float64) @ W.astype(np.float64).T
    Aabs = np.abs(G).astype(np.float64) @ np.abs(W).astype(np.float64).T
    sr = scale * np.sqrt(((X[:, None, :].astype(np.float64) - Y[None, :, :].astype(np.float64)) ** 2).sum(-1))
    with np.errstate(divide="ignore", invalid="ignore"):
        dk = np.where(sr > 0, -nconst * np.exp(-np.power(sr, power)) * np.power(sr, power) * np.log(sr), 0.0)
    gx, gs, gp = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, power, nconst, want_power=True)
    _check(gp.cpu().numpy(), (A * dk).sum(1), (Aabs * np.abs(dk)).sum(1), f"dL/dpower rows p={power}")
    gX, _, gsum, (mx, _, ms) = orc.gengauss_vjp_truth64(X, Y, A, s, power, nconst=nconst)
    _, _, _, (mxa, _, msa) = orc.gengauss_vjp_truth64(X, Y, Aabs, s, power, nconst=nconst)
    _check(gx.cpu().numpy(), gX, mxa, "dL/dX with the general epilogue")
    _check(float(gs.double().sum()), gsum, msa, "dL/dscale with the general epilogue")
    Gb = gauss(4, N, M)
    _, _, gp2 = nat.gram_grad(cu(X), cu(Y), cu(Gb), scale, power, nconst, want_power=True)
    _check(gp2.cpu().numpy(), (Gb.astype(np.float64) * dk).sum(1), (np.abs(Gb).astype(np.float64) * np.abs(dk)).sum(1), "gram dL/dpower rows")


def test_api_gradients_to_shape_and_per_dimension_length_scales():
    """GenGaussKernel.make(length_scale=tensor, shape=tensor): a loss over the Gram and over a fused inner product
    differentiated with respect to the shape parameter (the normalising constant follows it through lgamma) and -- for a
    vector of length scales -- each dimension's scale, against torch float64 autograd of the explicit formula."""
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import LinearReduce
    from jaxrk_b200.rkhs import FiniteVec, inner
    N, M, d, m = 200, 150, 4, 3
    Xn, Yn, An, Cn = gauss(0, N, d), gauss(1, M, d), gauss(2, m, M), gauss(3, N, M)

    def nconst64(s, p):
        return s * p / (2.0 * torch.exp(torch.lgamma(1.0 / p)))

    # (1) scalar length scale + shape
    ls = torch.tensor(1.7, device="cuda", requires_grad=True)
    sh = torch.tensor(1.4, device="cuda", requires_grad=True)
    Xt = cu(Xn).requires_grad_(True)
    k = GenGaussKernel.make(ls, sh)
    loss = (k(Xt, cu(Yn)) * cu(Cn)).sum() + inner(FiniteVec(k, Xt), FiniteVec(k, cu(Yn), [LinearReduce(cu(An))])).sum()
    loss.backward()
    ls64 = torch.tensor(1.7, dtype=torch.float64, requires_grad=True)
    sh64 = torch.tensor(1.4, dtype=torch.float64, requires_grad=True)
    X64 = torch.tensor(Xn, dtype=torch.float64, requires_grad=True)
    r = torch.cdist(X64, torch.tensor(Yn, dtype=torch.float64))
    K64 = nconst64(1.0 / ls64, sh64) * torch.exp(-torch.pow(r / ls64, sh64))
    loss64 = (K64 * torch.tensor(Cn, dtype=torch.float64)).sum() + (K64 @ torch.tensor(An, dtype=torch.float64).T).sum()
    loss64.backward()
    assert abs(float(loss) - float(loss64)) <= 1e-4 * abs(float(loss64)) + 1e-5
    for got, want, name in ((sh.grad, sh64.grad, "shape"), (ls.grad, ls64.grad, "length_scale")):
        assert abs(float(got) - float(want)) <= 2e-4 * abs(float(want)) + 2e-5, (name, float(got), float(want))
    np.testing.assert_allclose(Xt.grad.cpu().numpy(), X64.grad.numpy(), rtol=2e-4, atol=2e-5)

    # (2) one length scale per dimension
    lsv = torch.tensor([1.2, 2.0, 0.8, 1.5], device="cuda", requires_grad=True)
    kv = GenGaussKernel.make(lsv, 2.0)
    Kv = kv(cu(Xn), cu(Yn))
    (Kv * cu(Cn)).sum().backward()
    lsv64 = torch.tensor([1.2, 2.0, 0.8, 1.5], dtype=torch.float64, requires_grad=True)
    diff = (torch.tensor(Xn, dtype=torch.float64)[:, None, :] - torch.tensor(Yn, dtype=torch.float64)[None, :, :]) / lsv64
    Kv64 = nconst64(torch.tensor(1.0, dtype=torch.float64), torch.tensor(2.0, dtype=torch.float64)) * torch.exp(-(diff ** 2).sum(-1))
    (Kv64 * torch.tensor(Cn, dtype=torch.float64)).sum().backward()
    np.testing.assert_allclose(Kv.detach().cpu().numpy(), Kv64.detach().numpy(), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(lsv.grad.cpu().numpy(), lsv64.grad.numpy(), rtol=3e-4, atol=3e-5)
    with torch.no_grad():                                  # the same kernel object gives the same numbers without autograd
        assert torch.equal(kv(cu(Xn), cu(Yn)), Kv.detach())


@pytest.mark.parametrize("N,M,d,ls,positive", [(8192, 8192, 16, 4.0, False), (4096 + 300, 4096 + 12, 7, 2.5, True),
                                               (4224, 6000, 3, 1.7, False)])
def test_gram_grad_tensor_core_path(N, M, d, ls, positive):
    """jrk_gram_grad_f32 on the tensor-core kernel (Gram cotangent read from global memory, csrc/kgrad_tc.cu GBAR): sampled
    rows of dX / dscale and sampled columns of dY against float64."""
    rng = np.random.RandomState(N % 11)
    X, Y = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32)
    Gb = rng.randn(N, M).astype(f32)
    if positive:
        Gb = np.abs(Gb)
    s = orc.simple_scaler_s(orc.gauss_length_scale(ls))
    scale, nconst = float(s[0, 0]), float(orc.gengauss_nconst_ref32(s, 2.0).reshape(-1)[0])
    n0 = nat.launch_count()
    gx, gs = nat.gram_grad(cu(X), cu(Y), cu(Gb), scale, 2.0, nconst)
    assert nat.launch_count() - n0 == 6
    rows = np.r_[0:32, 128:140, 255:258, N - 32:N]
    gX, _, gsum, _ = orc.gengauss_vjp_truth64(X[rows], Y, Gb[rows].astype(np.float64), s, 2.0, nconst=nconst)
    _, _, _, (mx, _, ms) = orc.gengauss_vjp_truth64(X[rows], Y, np.abs(Gb[rows]).astype(np.float64), s, 2.0, nconst=nconst)
    _check(gx.cpu().numpy()[rows], gX, mx, "tensor-core gram dL/dX")
    _check(float(gs.double()[torch.from_numpy(rows).to(gs.device)].sum()), gsum, ms, "tensor-core gram dL/dscale")
    gy, _ = nat.gram_grad(cu(Y), cu(X), cu(Gb), scale, 2.0, nconst, want_scale=False, transposed=True)   # dL/dK as it is
    gxt, gst = nat.gram_grad(cu(X), cu(Y), cu(np.ascontiguousarray(Gb.T)), scale, 2.0, nconst, transposed=True)
    _check(gxt.cpu().numpy()[rows], gX, mx, "tensor-core gram dL/dX from the transposed cotangent")
    _check(float(gst.double()[torch.from_numpy(rows).to(gs.device)].sum()), gsum, ms, "tensor-core gram dL/dscale (transposed)")
    cols = np.r_[0:32, M - 32:M]
    _, gY, _, _ = orc.gengauss_vjp_truth64(X, Y[cols], Gb[:, cols].astype(np.float64), s, 2.0, nconst=nconst)
    _, _, _, (_, my, _) = orc.gengauss_vjp_truth64(X, Y[cols], np.abs(Gb[:, cols]).astype(np.float64), s, 2.0, nconst=nconst)
    _check(gy.cpu().numpy()[cols], gY, my, "tensor-core gram dL/dY")


@pytest.mark.parametrize("N,M,d", [(75776, 8192, 16), (151552 + 100, 4096 + 32 * 5 + 7, 16), (75776, 8192, 8), (50000, 4109, 5)])
def test_tensor_core_backward_every_row_with_several_units_per_cta(N, M, d):
    """More 256-row units than CTAs (a persistent CTA runs several units back to back) and a tile count that is not a
    multiple of the warps per lane quarter: EVERY row of the tensor-core kernel against the FP32-pipe kernel, the same
    call twice bit for bit, a row block on its own bit for bit -- for the K@W cotangent and for the Gram cotangent in
    both layouts.  (A version that handed its TMEM stage back early was right with one unit per CTA and wrong, differently
    from run to run, with two: sampled first / last rows never saw it.)  d <= 8 runs the kernel's 8-feature instantiation."""
    m = 16
    g = torch.Generator(device="cuda").manual_seed(N)
    X, Y = torch.randn(N, d, generator=g, device="cuda"), torch.randn(M, d, generator=g, device="cuda")
    W, G = torch.randn(M, m, generator=g, device="cuda"), torch.randn(N, m, generator=g, device="cuda")
    scale, nconst = 0.70710695 / (1.5 * float(np.sqrt(d))), 0.1
    a, sa = nat.kmatw_grad(X, Y, W, G, scale, 2.0, nconst)
    b, sb = nat.kmatw_grad(X, Y, W, G, scale, 2.0, nconst)
    assert torch.equal(a, b) and torch.equal(sa, sb)
    with nat.option(nat.OPT_KMATW_TC, 0):
        r, sr = nat.kmatw_grad(X, Y, W, G, scale, 2.0, nconst)
    assert float((a - r).abs().max()) <= 1e-3 * float(r.abs().max())
    assert float((sa - sr).abs().max()) <= 1e-3 * float(sr.abs().max())
    blk, _ = nat.kmatw_grad(X[256 * 150: 256 * 150 + 8192], Y, W, G[256 * 150: 256 * 150 + 8192], scale, 2.0, nconst)
    assert torch.equal(blk, a[256 * 150: 256 * 150 + 8192])
    if M % 4 == 0:
        Gb = torch.randn(N, M, generator=g, device="cuda")
        for tr in (False, True):
            cot = Gb.t().contiguous() if tr else Gb
            a, sa = nat.gram_grad(X, Y, cot, scale, 2.0, nconst, transposed=tr)
            b, sb = nat.gram_grad(X, Y, cot, scale, 2.0, nconst, transposed=tr)
            assert torch.equal(a, b) and torch.equal(sa, sb)
            with nat.option(nat.OPT_KMATW_TC, 0):
                r, sr = nat.gram_grad(X, Y, cot, scale, 2.0, nconst, transposed=tr)
            assert float((a - r).abs().max()) <= 1e-3 * float(r.abs().max())
            del cot

"""The oracle's analytic vector-Jacobian product (oracle/jaxrk_oracle.py: gengauss_vjp_truth64) against an
independent derivation: torch float64 autograd through the reference's op sequence (distance -> scale -> power ->
exp -> times nconst), which is what jax.grad differentiates in the reference.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import jaxrk_oracle as orc

f32 = np.float32


@pytest.mark.parametrize("power", [2.0, 1.0, 1.5, 0.7])
def test_vjp_matches_autograd_of_the_reference_chain(power):
    rng = np.random.RandomState(3)
    N, M, d = 23, 17, 4
    X, Y, A = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32), rng.randn(N, M)
    s = orc.simple_scaler_s(1.3)
    nc = float(orc.gengauss_nconst_ref32(s, power).reshape(-1)[0])
    gX, gY, gs, _ = orc.gengauss_vjp_truth64(X, Y, A, s, power, nconst=nc)
    Xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    Yt = torch.tensor(Y, dtype=torch.float64, requires_grad=True)
    st = torch.tensor(float(s[0, 0]), dtype=torch.float64, requires_grad=True)
    sq = ((Xt[:, None, :] - Yt[None, :, :]) ** 2).sum(-1)
    K = nc * torch.exp(-(st * torch.sqrt(sq)) ** float(f32(power)))
    (K * torch.tensor(A)).sum().backward()
    assert np.allclose(gX, Xt.grad.numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(gY, Yt.grad.numpy(), rtol=1e-9, atol=1e-12)
    assert np.isclose(gs, float(st.grad), rtol=1e-9)


def test_coincident_pairs_contribute_nothing():
    X = np.array([[0.0, 1.0], [2.0, 3.0]], f32)
    A = np.ones((2, 2))
    for power in (2.0, 1.0, 0.5):
        gX, gY, gs, _ = orc.gengauss_vjp_truth64(X, X, A, orc.simple_scaler_s(1.0), power)
        assert np.isfinite(gX).all() and np.isfinite(gY).all() and np.isfinite(gs)
        assert np.allclose(gX, gY) and np.allclose((gX + gY).sum(0), 0)   # symmetric cotangent; translation invariance

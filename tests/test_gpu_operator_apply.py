"""Operator application through the operator objects: FiniteOp.__matmul__ / Cmo / Cdo
(jaxrk/rkhs/operator.py:33-56, 173-185) with K evaluated ONCE -- the coefficient map stays in factored form
(KernelMapReduce), wide maps go through the tensor-core kernel in column blocks (or Gram slabs + the library's FP32 GEMM
outside its window), and
FiniteVec.get_mean_var (jaxrk/rkhs/vector.py:126-134) contracts the map with a skinny right-hand side."""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from jaxrk_b200.kern import GenGaussKernel
from jaxrk_b200.reduce import KernelMapReduce, LinearReduce
from jaxrk_b200.rkhs import FiniteVec, inner
from jaxrk_b200.rkhs import _lowering as low
from jaxrk_b200.rkhs.operator import Cdo, Cmo, FiniteOp
from oracle import jaxrk_oracle as orc
from gpu_util import assert_kmatw_close, cu, f32, gauss

pytestmark = pytest.mark.gpu


def n(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("positive", [True, False])
@pytest.mark.parametrize("kname,power", [("gauss", 2.0), ("laplace", 1.0)])
def test_wide_map_gram_slabs_plus_split_gemm(kname, power, positive, monkeypatch):
    """Wide W outside the tensor-core window (here d = 40): K in row slabs by the fused Gram kernel, the contraction by the
    library's FP32 GEMM -- against float64 on sampled rows; all-positive W is the hard case."""
    N, M, d, m = 6000, 9000, 40, 512
    monkeypatch.setattr(low, "WIDE_SLAB_BYTES", 64 << 20)        # several slabs, ragged last one
    X, Y = gauss(0, N, d), gauss(1, M, d)
    W = gauss(2, M, m)
    if positive:
        W = np.abs(W)
    rp, cp = np.random.RandomState(3).rand(N).astype(f32), np.random.RandomState(4).rand(M).astype(f32)
    k = GenGaussKernel.make_gauss(4.0) if kname == "gauss" else GenGaussKernel.make_laplace(4.0)
    spec = low.kernel_spec_of(k)
    scale = spec.kparams[0]
    n0 = nat.launch_count()
    T = low._kmatw_wide(cu(X), cu(Y), cu(W), spec, cu(rp), cu(cp))
    slabs = -(-N // max(128, (64 << 20) // (4 * M) // 128 * 128))
    assert nat.launch_count() - n0 <= 5 * slabs, "K must be evaluated once per slab"
    rows = np.r_[0:40, N - 40:N]
    G = orc.gengauss_gram_truth64(X[rows], Y, np.array([[scale]]), power) * cp.astype(np.float64)[None, :] * rp[rows].astype(np.float64)[:, None]
    truth, bound = G @ W.astype(np.float64), np.abs(G) @ np.abs(W.astype(np.float64))
    assert_kmatw_close(n(T)[rows], truth, bound, f"wide map {kname} positive={positive}")


def test_operator_apply_keeps_the_map_factored_and_get_mean_var_is_skinny():
    rng = np.random.RandomState(0)
    N, T_, d, do = 4096, 20000, 16, 3
    Xin, Yout, Xt = rng.randn(N, d).astype(f32), rng.randn(N, do).astype(f32), rng.randn(T_, d).astype(f32)
    M_ = (rng.randn(N, N) / N).astype(f32)
    kx, ky = GenGaussKernel.make_gauss(4.0), GenGaussKernel.make_gauss(1.5)
    op = FiniteOp(FiniteVec(kx, Xin), FiniteVec(ky, Yout), M_)
    res = op @ Xt
    red = res.reduce[-1]
    assert isinstance(red, KernelMapReduce) and isinstance(red, LinearReduce) and not red.is_materialised
    assert len(res) == T_
    n0 = nat.launch_count()
    mean, var = res.get_mean_var()
    assert not red.is_materialised, "get_mean_var must not build the T x N map"
    assert nat.launch_count() - n0 <= 24
    # float64 truth: (K(Xt, Xin) M^T) Yout  and the second moment
    s = orc.simple_scaler_s(orc.gauss_length_scale(4.0))
    rows = np.arange(0, T_, 311)
    K = orc.gengauss_gram_truth64(Xt[rows], Xin, s, 2.0)
    C = K @ M_.astype(np.float64).T
    m_t = C @ Yout.astype(np.float64)
    np.testing.assert_allclose(n(mean)[rows], m_t, rtol=2e-4, atol=2e-6)
    v_t = float(np.asarray(n(torch.as_tensor(ky.var()))).reshape(-1)[0]) + C @ (Yout.astype(np.float64) ** 2) - m_t ** 2
    np.testing.assert_allclose(n(var)[rows], v_t, rtol=2e-4, atol=2e-5)
    # the materialised map (what the reference stores) is the same linear map
    L = red.linear_map
    assert red.is_materialised and tuple(L.shape) == (T_, N)
    bound = np.abs(K) @ np.abs(M_.astype(np.float64).T)
    assert_kmatw_close(n(L)[rows], C, bound, "materialised coefficient map")
    # point evaluation of the result (FiniteVec.__call__): K_y(Yt, Yout) contracted with the map
    Yt = rng.randn(7, do).astype(f32)
    ev = res(Yt)
    Ky = orc.gengauss_gram_truth64(Yout, Yt, orc.simple_scaler_s(orc.gauss_length_scale(1.5)), 2.0)
    np.testing.assert_allclose(n(ev)[rows], C @ Ky, rtol=3e-4, atol=3e-6)


def test_cdo_apply_through_the_operator_objects_n4096():
    """Cdo(inp, outp, ref, regul) @ X_test at N = 4096 (BASELINE configs[4] at reduced size) against a float64
    restatement of jaxrk/rkhs/operator.py:128-185 on the same float32 inputs."""
    rng = np.random.RandomState(1)
    N, R, T_, d = 4096, 1024, 8192, 16
    X, Y, Rf, Xt = (rng.randn(N, d).astype(f32), rng.randn(N, 1).astype(f32), rng.randn(R, 1).astype(f32) * 1.5,
                    rng.randn(T_, d).astype(f32))
    kx, ky = GenGaussKernel.make_gauss(4.0), GenGaussKernel.make_gauss(0.7)
    inp, outp, ref = FiniteVec(kx, X), FiniteVec(ky, Y), FiniteVec(ky, Rf)
    regul = 1.0                                                   # (C + regul)^-2 twice in float32: keep it well conditioned
    cdo = Cdo(inp, outp, ref, regul=regul)
    res = cdo @ Xt
    assert isinstance(res.reduce[-1], KernelMapReduce)
    got = res.reduce[-1].linear_map                              # T x R coefficients over the reference points
    # float64 restatement
    sx, sy = orc.simple_scaler_s(orc.gauss_length_scale(4.0)), orc.simple_scaler_s(orc.gauss_length_scale(0.7))
    Gx = orc.gengauss_gram_truth64(X, X, sx, 2.0)
    Gr = orc.gengauss_gram_truth64(Rf, Rf, sy, 2.0)
    Gry = orc.gengauss_gram_truth64(Rf, Y, sy, 2.0)
    ix = np.linalg.inv(Gx + regul * np.eye(N))
    cov_inv_x = np.linalg.inv(np.eye(N) / N) @ (ix @ ix)                     # Cov_inv(CovOp(inp)).matr
    # Cmo = CrossCovOp(Cov_solve(CovOp(inp), inp), outp): Cov_solve(...) = Cov_inv @ inp -> inp with LinearReduce((matr Gx)^T)
    lin = (cov_inv_x @ Gx).T                                                  # N x N map on the inp side
    # CrossCovOp(v, outp).matr = I / N over (outp, v): the operator sum_i outp_i (x) v_i / N, v_i = sum_j lin[i, j] inp_j
    # => as an operator over the raw inp features: matr_mo = (I / N) @ lin
    matr_mo = lin / N
    ir = np.linalg.inv(Gr + regul * np.eye(R))
    cov_inv_r = np.linalg.inv(np.eye(R) / R) @ (ir @ ir)
    matr_do = cov_inv_r @ Gry @ matr_mo                                       # R x N: Cov_inv(ref) @ mo
    rows = np.arange(0, T_, 97)
    Kt = orc.gengauss_gram_truth64(Xt[rows], X, sx, 2.0)
    truth = Kt @ matr_do.T
    scale_ = np.abs(truth).max()
    # the two regularised inverses amplify float32 rounding (cond ~ 1e4 .. 1e5): this checks the PIPELINE, the kernels
    # themselves are pinned to 1e-5 elsewhere
    np.testing.assert_allclose(n(got)[rows], truth, rtol=0.1, atol=2e-2 * scale_)
    # and to kernel tolerance: the apply is K(Xt, X) @ A^T with A the operator's own folded float32 map over the raw points
    inp_side = cdo.inp_feat.extend_reduce([LinearReduce(cdo.matr)])
    A = n(low.fold_chain(inp_side.reduce, N, got.device).A).astype(np.float64)          # R x N
    assert_kmatw_close(n(got)[rows], Kt @ A.T, np.abs(Kt) @ np.abs(A.T), "Cdo apply vs its own folded map")

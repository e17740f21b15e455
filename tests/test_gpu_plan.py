"""jrk_kmatw_plan_* (the Y / W side kept on the device across applies) and wide W on the tensor-core path (column blocks
of 16): results must be those of jrk_kmatw_f32 -- bit for bit where the same kernels run -- and within tolerance of the
float64 oracle.  Reference semantics: FiniteVec.inner / FiniteOp.__matmul__ against fixed inspection points
(jaxrk/rkhs/vector.py:62-75, jaxrk/rkhs/operator.py:33-56)."""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from oracle import jaxrk_oracle as orc
from gpu_util import assert_kmatw_close, cu, f32, gauss, kparams

pytestmark = pytest.mark.gpu


def _truth(X, Y, W, s, p, rows, rp=None, cp=None):
    G = orc.gengauss_gram_truth64(X[rows], Y, s, p)
    if cp is not None:
        G = G * cp.astype(np.float64)[None, :]
    if rp is not None:
        G = G * rp[rows].astype(np.float64)[:, None]
    W64 = W.astype(np.float64)
    return G @ W64, np.abs(G) @ np.abs(W64)


@pytest.mark.parametrize("kname,ls", [("sqexp", 4.0), ("sqexp", 1.0), ("laplace", 4.0), ("gg15", 3.0)])
def test_plan_apply_equals_one_shot_call(kname, ls):
    """Factored regime (sqexp, ls = 4), generic regime (sqexp with a short length scale; p != 2): the plan's apply gives
    the one-shot result bit for bit, for several X of different sizes against the same plan."""
    M, d, m = 4096 + 33, 16, 12
    Y, W = gauss(1, M, d), gauss(2, M, m)
    cp = np.random.RandomState(3).rand(M).astype(f32)
    s, scale, p, nconst = kparams(kname, ls)
    plan = nat.KmatwPlan(cu(Y), cu(W), scale, p, nconst, col_pref=cp)
    for seed, N in ((10, 8192 + 77), (11, 16384), (12, 8192)):
        X = gauss(seed, N, d)
        rp = np.random.RandomState(seed).rand(N).astype(f32)
        A = plan.apply(cu(X), row_pref=rp)
        B = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, row_pref=rp, col_pref=cp)
        assert torch.equal(A, B)
        rows = np.r_[0:48, N - 48:N]
        truth, bound = _truth(X, Y, W, s, p, rows, rp, cp)
        assert_kmatw_close(A.cpu().numpy()[rows], truth, bound, f"plan {kname} ls={ls} N={N}")
    plan.close()


def test_plan_stays_correct_when_x_leaves_the_factored_regime():
    """The plan packs its images for "Y alone inside the regime"; an X with large norms pushes the pair out of it, and the
    apply must then take the generic epilogue on its own (device-side decision, first-generation image in scratch)."""
    N, M, d, m = 8192, 4096, 16, 16
    Y, W = gauss(1, M, d), np.abs(gauss(2, M, m))
    s, scale, p, nconst = kparams("sqexp", 4.0)
    c = float(scale) ** 2 * 1.4426950408889634
    plan = nat.KmatwPlan(cu(Y), cu(W), scale, p, nconst)
    for factor, regime in ((1.0, "fast"), (3.0, "generic")):
        X = gauss(5, N, d) * f32(factor)
        b = c * ((X.astype(np.float64) ** 2).sum(1).max() + (Y.astype(np.float64) ** 2).sum(1).max())
        assert (b <= 12.0) == (regime == "fast")
        A = plan.apply(cu(X))
        B = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
        assert torch.equal(A, B)
        rows = np.arange(0, N, 61)
        truth, bound = _truth(X, Y, W, s, p, rows)
        assert_kmatw_close(A.cpu().numpy()[rows], truth, bound, f"plan, {regime} regime")
    plan.close()


@pytest.mark.parametrize("rr", [nat.REDUCE_SUM, nat.REDUCE_MEAN])
def test_plan_with_row_reduce_and_host_buffers(rr):
    N, M, d, m = 8192 + 300, 4096 + 17, 16, 7
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    rp = np.random.RandomState(3).rand(N).astype(f32)
    s, scale, p, nconst = kparams("sqexp", np.sqrt(d))
    plan = nat.KmatwPlan(Y, W, scale, p, nconst)                  # created from HOST arrays
    out = np.empty((1, m), f32)
    plan.apply_host(X, out, row_pref=rp, row_reduce=rr)
    ref = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, row_pref=rp, row_reduce=rr).cpu().numpy()
    assert np.array_equal(out, ref)
    full = np.empty((N, m), f32)
    plan.apply_host(X, full)                                      # the staging buffers grow on demand
    ref_full = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst).cpu().numpy()
    assert np.array_equal(full, ref_full)
    plan.close()
    plan.close()                                                  # idempotent


def test_plan_outside_the_tensor_core_window_uses_the_fp32_kernel():
    """d > 32, per-dimension scales, few rows: the plan still works (device copies of Y / W, FP32-pipe kernel)."""
    N, M, d, m = 500, 700, 40, 5
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    ds = (0.5 + np.random.RandomState(4).rand(d)).astype(f32)
    s, scale, p, nconst = kparams("laplace", 3.0)
    plan = nat.KmatwPlan(cu(Y), cu(W), scale, p, nconst, dim_scale=ds)
    A = plan.apply(cu(X), row_reduce=nat.REDUCE_BALANCED, group=5)
    B = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst, dim_scale=ds, row_reduce=nat.REDUCE_BALANCED, group=5)
    assert torch.equal(A, B) and tuple(A.shape) == (N // 5, m)
    plan.close()


@pytest.mark.parametrize("kname", ["sqexp", "laplace"])
@pytest.mark.parametrize("m", [17, 40, 128])
def test_wide_w_runs_on_the_tensor_cores_in_column_blocks(kname, m):
    """16 < m: K@W in column blocks of 16 through the tensor-core kernels (the FP32-pipe kernel recomputed K per 32
    columns at a fifth of the speed).  Against the oracle, the FP32-pipe kernel and a plan."""
    N, M, d = 8192 + 77, 4096 + 33, 16
    X, Y, W = gauss(0, N, d), gauss(1, M, d), gauss(2, M, m)
    s, scale, p, nconst = kparams(kname, 4.0)
    assert nat.lib().jrk_kmatw_uses_tensor_cores(N, M, d, m, 0, 0) == 1
    T = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    with nat.option(nat.OPT_KMATW_TC, 0):
        F = nat.kmatw(cu(X), cu(Y), cu(W), scale, p, nconst)
    rows = np.r_[0:32, N - 32:N]
    truth, bound = _truth(X, Y, W, s, p, rows)
    assert_kmatw_close(T.cpu().numpy()[rows], truth, bound, f"tc wide {kname} m={m}")
    assert_kmatw_close(F.cpu().numpy()[rows], truth, bound, f"fp32 wide {kname} m={m}")
    plan = nat.KmatwPlan(cu(Y), cu(W), scale, p, nconst)
    assert torch.equal(plan.apply(cu(X)), T)
    S = plan.apply(cu(X), row_reduce=nat.REDUCE_SUM)
    assert tuple(S.shape) == (1, m)
    assert torch.allclose(S, T.double().sum(0, keepdim=True).float(), rtol=2e-5, atol=1e-4)
    plan.close()

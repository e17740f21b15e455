"""BASELINE.json's configurations at FULL size, checked through size-independent properties (sampled entries against
the float64 oracle, checksums computed by an independent kernel, linearity, row-block independence, symmetry).
The oracle cannot evaluate 10^9 .. 10^13 kernel values in seconds; these properties can be checked at any size."""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from jaxrk_b200.kern import GenGaussKernel
from jaxrk_b200.reduce import BalancedRed, LinearReduce
from jaxrk_b200.rkhs import FiniteVec, inner
from oracle import jaxrk_oracle as orc
from gpu_util import ATOL, RTOL, f32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def randn(seed, *shape):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)


def kernel_for(d):
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    scale, _, power, nconst = k.kernel_args()
    return k, scale, power, nconst, orc.simple_scaler_s(orc.gauss_length_scale(float(np.sqrt(d))))


def test_cfg2_gram_65536_d64_sampled_entries_and_row_checksums():
    N = M = 65536
    d = 64
    X, Y = randn(100, N, d), randn(1, M, d)
    k, scale, power, nconst, s = kernel_for(d)
    K = nat.gram(X, Y, scale, power, nconst)                 # tcgen05 3xTF32 path (AUTO, d >= 64), 17.2 GB
    assert tuple(K.shape) == (N, M)
    # (i) 8192 sampled entries against float64 direct differencing
    rng = np.random.RandomState(0)
    ii, jj = rng.randint(0, N, 8192), rng.randint(0, M, 8192)
    Xs, Ys = X[torch.from_numpy(ii).to(dev)].cpu().numpy(), Y[torch.from_numpy(jj).to(dev)].cpu().numpy()
    sq = ((Xs.astype(np.float64) - Ys.astype(np.float64)) ** 2).sum(1)
    truth = float(nconst) * np.exp(-(float(scale) ** 2) * sq)
    got = K[torch.from_numpy(ii).to(dev), torch.from_numpy(jj).to(dev)].cpu().numpy().astype(np.float64)
    assert (np.abs(got - truth) <= ATOL + RTOL * np.abs(truth)).all()
    # (ii) row checksums: K @ 1 from the fused never-materialised kernel (an independent code path) == K.sum(1)
    ones = torch.ones((M, 1), device=dev)
    chk = nat.kmatw(X, Y, ones, scale, power, nconst)
    rs = K.sum(1, keepdim=True, dtype=torch.float64)
    assert ((chk.double() - rs).abs() <= 1e-6 + 2e-5 * rs.abs()).all()
    # (iii) every entry is a valid kernel value
    assert bool((K >= 0).all()) and float(K.max()) <= float(nconst) * (1 + 1e-6)
    del K


def _c_oracle():
    from oracle import c_oracle
    import ctypes
    return c_oracle.COracle(ctypes.CDLL(c_oracle.build()))


def _schedule_rows(N, n_extra=108, seed=0):
    """Rows that touch every CTA of the persistent schedule: unit u (128 rows) runs on CTA u % 148; one row per residue
    (different wave, lane quarter and lane each), the first and last rows, the last unit, plus random ones: 256+ rows."""
    rng = np.random.RandomState(seed)
    n_units = (N + 127) // 128
    rows = []
    for r in range(min(148, n_units)):
        waves = (n_units - 1 - r) // 148 + 1
        u = r + 148 * rng.randint(0, waves)
        rows.append(min(N - 1, u * 128 + rng.randint(0, 128)))
    rows += [0, 1, 127, 128, N - 129, N - 128, N - 2, N - 1]
    rows += list(rng.randint(0, N, n_extra))
    return np.unique(np.array(rows))


def _check_rows(T, X, Y, W, rows, scale, power, nconst, what):
    orc_c = _c_oracle()
    Xr = X[torch.from_numpy(rows).to(dev)].cpu().numpy()
    truth, bnd = orc_c.kmatw_truth64(Xr, Y.cpu().numpy(), W.cpu().numpy(), float(scale), float(power), float(nconst), with_abs=True)
    got = T[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.float64)
    err = np.abs(got - truth) / (ATOL + RTOL * bnd)
    assert err.max() <= 1.0, f"{what}: worst error {err.max():.3f} of the tolerance at row {rows[err.max(1).argmax()]}"
    return float(err.max())


def test_cfg3_kmatw_2pow20_linearity_row_blocks_and_sampled_rows():
    N = M = 1 << 20
    d, m = 16, 16
    X, Y, W1, W2 = randn(100, N, d), randn(1, M, d), randn(2, M, m), randn(3, M, m)
    k, scale, power, nconst, s = kernel_for(d)
    assert nat.lib().jrk_kmatw_uses_tensor_cores(N, M, d, m, 0, 0) == 1
    T1 = nat.kmatw(X, Y, W1, scale, power, nconst)
    T2 = nat.kmatw(X, Y, W2, scale, power, nconst)
    # linearity in W
    T12 = nat.kmatw(X, Y, 0.5 * W1 - 2.0 * W2, scale, power, nconst)
    mag = (0.5 * T1.abs() + 2.0 * T2.abs())
    assert ((T12 - (0.5 * T1 - 2.0 * T2)).abs() <= 1e-5 + 1e-4 * mag.max()).all()
    # a row block computed on its own (what another rank would do) gives the same rows bit for bit
    blk = nat.kmatw(X[N // 2: N // 2 + 131072], Y, W1, scale, power, nconst)
    assert torch.equal(blk, T1[N // 2: N // 2 + 131072])
    # Sum over rows folded into the call == column sums of the N x m result
    S = nat.kmatw(X, Y, W1, scale, power, nconst, row_reduce=nat.REDUCE_SUM)
    cs = T1.sum(0, keepdim=True, dtype=torch.float64)
    bound = T1.abs().sum(0, keepdim=True, dtype=torch.float64)
    assert ((S.double() - cs).abs() <= 1e-6 + 1e-5 * bound).all()
    # 256+ sampled rows -- every CTA of the 148-CTA schedule, every wave position, the last unit -- against the float64
    # C oracle (OpenMP): 2.7e8 kernel values on the host
    rows = _schedule_rows(N)
    assert len(rows) >= 256
    _check_rows(T1, X, Y, W1, rows, scale, power, nconst, "cfg3")
    # a plan gives the same numbers
    plan = nat.KmatwPlan(Y, W1, scale, power, nconst)
    assert torch.equal(plan.apply(X), T1)
    plan.close()


def test_cfg3_size_all_positive_weights_and_laplace():
    """The hard case for the truncating tensor-core accumulation (all-positive W, 2^20 terms per output) and the generic
    epilogue (make_laplace) at the full cfg3 size, on sampled rows."""
    N = M = 1 << 20
    d, m = 16, 16
    X, Y = randn(100, N, d), randn(1, M, d)
    W = torch.rand(M, m, generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    rows = _schedule_rows(N, n_extra=20, seed=1)
    k, scale, power, nconst, s = kernel_for(d)
    T = nat.kmatw(X, Y, W, scale, power, nconst)
    worst = _check_rows(T, X, Y, W, rows, scale, power, nconst, "cfg3, all-positive W")
    assert worst < 0.8
    kl = GenGaussKernel.make_laplace(float(np.sqrt(d)))
    sc, _, pw, nc = kl.kernel_args()
    Wn = randn(2, M, m)
    Tl = nat.kmatw(X, Y, Wn, sc, pw, nc)
    _check_rows(Tl, X, Y, Wn, rows[::4], sc, pw, nc, "cfg3 size, make_laplace")


@pytest.mark.parametrize("d,m", [(16, 24), (16, 32), (48, 16), (128, 8)])
def test_kmatw_outside_the_narrow_window_at_n65536(d, m):
    """16 < m <= 32 (tensor cores, column blocks of 16) and 32 < d <= 128 (FP32-pipe kernel) at N = M = 65536."""
    N = M = 65536
    X, Y, W = randn(100, N, d), randn(1, M, d), randn(2, M, m)
    k, scale, power, nconst, s = kernel_for(d)
    assert nat.lib().jrk_kmatw_uses_tensor_cores(N, M, d, m, 0, 0) == (1 if d <= 32 else 0)
    T = nat.kmatw(X, Y, W, scale, power, nconst)
    _check_rows(T, X, Y, W, _schedule_rows(N, n_extra=0, seed=2)[::2], scale, power, nconst, f"d={d} m={m}")


def test_cfg4_batched_inner_products_4m_points_symmetric_and_sampled_groups():
    groups, g, d = 1024, 4096, 32
    X = randn(0, groups * g, d)
    k, scale, power, nconst, s = kernel_for(d)
    v = FiniteVec(k, X, [BalancedRed(g, average=True)])
    G = v.inner()                                            # 1.76e13 kernel evaluations, never materialised
    assert tuple(G.shape) == (groups, groups)
    assert torch.equal(G, G.T)                               # exactly symmetric (jaxrk/utilities/gram.py:18)
    # two sampled group pairs against the float64 oracle (4096 x 4096 evaluations each)
    for a, b in ((3, 3), (17, 900)):
        Xa, Xb = X[a * g:(a + 1) * g].cpu().numpy(), X[b * g:(b + 1) * g].cpu().numpy()
        truth = orc.gengauss_gram_truth64(Xa, Xb, s, power, nconst=nconst).mean()
        assert abs(float(G[a, b]) - truth) <= ATOL + RTOL * abs(truth)


def test_cfg5_operator_apply_sampled_rows():
    N, T, d, m = 32768, 1 << 20, 16, 16
    X, Xt, A = randn(2, N, d), randn(3, T, d), randn(4, m, N)
    k, scale, power, nconst, s = kernel_for(d)
    Gm = inner(FiniteVec(k, X))                              # the Gram that feeds the (library) solve
    assert torch.equal(Gm, Gm.T) and bool((torch.diagonal(Gm) == Gm[0, 0]).all())
    out = inner(FiniteVec(k, Xt), FiniteVec(k, X, [LinearReduce(A)]))     # operator applied to 2^20 test points
    assert tuple(out.shape) == (T, m)
    _check_rows(out, Xt, X, A.t().contiguous(), _schedule_rows(T, n_extra=0, seed=3)[::2], scale, power, nconst, "cfg5 apply")


def test_cfg5_cdo_through_the_operator_objects_full_size():
    """BASELINE configs[4] through the operator objects (jaxrk/rkhs/operator.py:128-185): Cdo(inp, outp, ref, regul) with
    N = 32768 points (two Gram builds feeding the library's regularised inverses), applied to 2^20 test points.  The
    coefficient map (2^20 x 256 over the reference grid) is ONE fused K(X_test, X) @ A^T; 74+ sampled rows against the
    float64 oracle with the operator's own folded float32 map A."""
    from jaxrk_b200.reduce import KernelMapReduce
    from jaxrk_b200.rkhs import _lowering as low
    from jaxrk_b200.rkhs.operator import Cdo
    N, T, d, R = 32768, 1 << 20, 16, 256
    X, Xt = randn(2, N, d), randn(3, T, d)
    Yv = (torch.sin(X[:, :1]) + 0.1 * randn(5, N, 1))
    Rf = torch.linspace(-2, 2, R, device=dev)[:, None]
    kx, scale, power, nconst, s = kernel_for(d)
    ky = GenGaussKernel.make_gauss(0.3)
    inp, outp, ref = FiniteVec(kx, X), FiniteVec(ky, Yv), FiniteVec(ky, Rf)
    cdo = Cdo(inp, outp, ref, regul=1.0)
    _ = cdo.inp_feat.reduce[-1].linear_map                   # the operator's own N x N map (Cmo's solve), built once
    n0 = nat.launch_count()
    res = cdo @ Xt
    assert isinstance(res.reduce[-1], KernelMapReduce) and len(res) == T
    coeff = res.reduce[-1].linear_map                        # T x R
    launches = nat.launch_count() - n0
    assert tuple(coeff.shape) == (T, R) and bool(torch.isfinite(coeff).all())
    assert launches <= 16 * 8, "K(X_test, X) must be evaluated once per 16-column block, not per output feature"
    inp_side = cdo.inp_feat.extend_reduce([LinearReduce(cdo.matr)])
    A = low.fold_chain(inp_side.reduce, N, dev).A            # R x N, float32: the map the fused kernel contracted with
    _check_rows(coeff, Xt, X, A.t().contiguous(), _schedule_rows(T, n_extra=0, seed=4)[::2], scale, power, nconst, "Cdo @ X_test")
    # the estimated conditional density is evaluated on the reference grid without materialising anything else
    dens = res(Rf)
    assert tuple(dens.shape) == (T, R) and bool(torch.isfinite(dens).all())


def test_cfg3_backward_2pow20_sampled_rows_linearity_and_row_blocks():
    """The K@W vector-Jacobian product at the full cfg3 size (2^40 pairs, csrc/kgrad_tc.cu): sampled rows of dX and of the
    dscale rows -- both row blocks of the 256-row units, every CTA of the schedule, the last unit -- against float64
    (evaluated on the device in double precision, chunked over Y), linearity in the cotangent, row-block independence."""
    N = M = 1 << 20
    d, m = 16, 16
    X, Y, W, G1, G2 = randn(100, N, d), randn(1, M, d), randn(2, M, m), randn(3, N, m), randn(4, N, m)
    k, scale, power, nconst, s = kernel_for(d)
    gx1, gs1 = nat.kmatw_grad(X, Y, W, G1, scale, power, nconst)
    gx2, gs2 = nat.kmatw_grad(X, Y, W, G2, scale, power, nconst)
    gx12, gs12 = nat.kmatw_grad(X, Y, W, 0.5 * G1 - 2.0 * G2, scale, power, nconst)
    blk, blk_s = nat.kmatw_grad(X[N // 2: N // 2 + 65536], Y, W, G1[N // 2: N // 2 + 65536], scale, power, nconst)
    assert torch.equal(blk, gx1[N // 2: N // 2 + 65536]) and torch.equal(blk_s, gs1[N // 2: N // 2 + 65536])
    # sampled rows in float64 on the device
    rng = np.random.RandomState(7)
    n_units = N // 256
    rows = [0, 127, 128, 255, 256, N - 256, N - 129, N - 128, N - 1]
    for r in range(0, 148, 3):
        u = r + 148 * rng.randint(0, (n_units - 1 - r) // 148 + 1)
        rows.append(u * 256 + rng.randint(0, 256))
    rows = torch.from_numpy(np.unique(np.array(rows))).to(dev)
    Xr, Gr = X[rows].double(), G1[rows].double()
    Gmix = 0.5 * G1[rows].double().abs() + 2.0 * G2[rows].double().abs()
    mMix = torch.zeros((len(rows), d), dtype=torch.float64, device=dev)
    sc = float(scale)
    gX = torch.zeros((len(rows), d), dtype=torch.float64, device=dev)
    gS = torch.zeros((len(rows),), dtype=torch.float64, device=dev)
    mX = torch.zeros_like(gX)
    mS = torch.zeros_like(gS)
    for j0 in range(0, M, 1 << 16):
        Yc, Wc = Y[j0:j0 + (1 << 16)].double(), W[j0:j0 + (1 << 16)].double()
        diff = Xr[:, None, :] - Yc[None, :, :]                      # rows x chunk x d
        sq = (diff ** 2).sum(-1)
        Kc = float(nconst) * torch.exp(-(sc * sc) * sq)
        A, Aabs = Gr @ Wc.T, Gr.abs() @ Wc.abs().T
        T = A * Kc * (2.0 * sc * sc)                                # dK/dx_i = -K 2 s^2 (x_i - y_j)
        gX -= (T[:, :, None] * diff).sum(1)
        mX += ((Aabs * Kc * (2.0 * sc * sc))[:, :, None] * diff.abs()).sum(1)
        mMix += (((Gmix @ Wc.abs().T) * Kc * (2.0 * sc * sc))[:, :, None] * diff.abs()).sum(1)
        gS -= (A * Kc * 2.0 * sc * sq).sum(1)                       # dK/ds = -K 2 s r^2
        mS += (Aabs * Kc * 2.0 * sc * sq).sum(1)
    ex = ((gx1[rows].double() - gX).abs() / (2e-6 + 2e-5 * mX)).max()
    es = ((gs1[rows].double() - gS).abs() / (2e-6 + 2e-5 * mS)).max()
    assert float(ex) <= 1.0 and float(es) <= 1.0, (float(ex), float(es))
    # linearity in the cotangent, judged like everything else against the accumulated magnitude (the sums cancel: |gX| is
    # ~ 1e-3 of it at this size)
    lin = ((gx12[rows] - (0.5 * gx1[rows] - 2.0 * gx2[rows])).double().abs() / (2e-6 + 2e-5 * 3 * mMix)).max()
    assert float(lin) <= 1.0, float(lin)

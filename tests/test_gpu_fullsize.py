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
    # 6 sampled rows against the float64 oracle (6 x 2^20 kernel values on the host)
    rows = np.array([0, 1, 77777, N // 2, N - 2, N - 1])
    Xr = X[torch.from_numpy(rows).to(dev)].cpu().numpy()
    Yh, Wh = Y.cpu().numpy(), W1.cpu().numpy()
    truth = orc.kmatw_truth64(Xr, Yh, Wh, s, power, nconst=nconst)
    bnd = orc.kmatw_absbound64(Xr, Yh, Wh, s, power, nconst=nconst)
    got = T1[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.float64)
    assert (np.abs(got - truth) <= ATOL + RTOL * bnd).all()


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
    rows = np.array([0, 12345, T - 1])
    Xr = Xt[torch.from_numpy(rows).to(dev)].cpu().numpy()
    Xh, Wh = X.cpu().numpy(), A.t().contiguous().cpu().numpy()
    truth = orc.kmatw_truth64(Xr, Xh, Wh, s, power, nconst=nconst)
    bnd = orc.kmatw_absbound64(Xr, Xh, Wh, s, power, nconst=nconst)
    got = out[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.float64)
    assert (np.abs(got - truth) <= ATOL + RTOL * bnd).all()

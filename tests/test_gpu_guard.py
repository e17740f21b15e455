"""Out-of-bounds guard: every kernel writes into a view in the middle of a sentinel-filled buffer; the bytes before,
after and (for strided outputs) between the rows must be untouched.  (compute-sanitizer is not available on the
GPU pool, so the bounds of the ragged-edge code paths are checked this way.)"""
import os

import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from gpu_util import cu, gauss

pytestmark = pytest.mark.gpu
SENT = 12345.0


def guarded(rows, cols, pad_cols=5, pad_rows=3):
    """(full buffer, view of shape rows x cols with row stride cols + pad_cols)"""
    full = torch.full((rows + 2 * pad_rows, cols + pad_cols), SENT, dtype=torch.float32, device="cuda:0")
    return full, full[pad_rows:pad_rows + rows, :cols]


def untouched(full, rows, cols, pad_rows=3):
    mask = torch.ones_like(full, dtype=torch.bool)
    mask[pad_rows:pad_rows + rows, :cols] = False
    return bool((full[mask] == SENT).all()) and not bool((full[~mask] == SENT).any())


@pytest.mark.parametrize("N,M,d,path", [(130, 67, 5, nat.PATH_DIRECT), (257, 131, 16, nat.PATH_TF32X3), (129, 200, 33, nat.PATH_TF32X3),
                                         (65, 190, 100, nat.PATH_TF32X3), (300, 129, 12, nat.PATH_TF32X3)])
@pytest.mark.parametrize("p", [2.0, 1.0])
def test_gram_writes_only_its_output(N, M, d, path, p):
    X, Y = cu(gauss(0, N, d)), cu(gauss(1, M, d))
    full, view = guarded(N, M)
    nat.gram(X, Y, 0.2, p, 0.3, path=path, out=view)
    torch.cuda.synchronize()
    assert untouched(full, N, M)
    full, view = guarded(N, N)
    nat.gram(X, None, 0.2, p, 0.3, path=path, out=view)
    torch.cuda.synchronize()
    assert untouched(full, N, N)


@pytest.mark.parametrize("N,M,d,m,scale", [(8192 + 5, 4096 + 70, 16, 16, 0.15), (8192, 4096 + 3, 7, 3, 0.2), (8192 + 130, 4096, 30, 13, 1.5),
                                           (130, 300, 16, 16, 0.3), (700, 129, 3, 5, 0.3)])
@pytest.mark.parametrize("p", [2.0, 1.0])
def test_kmatw_writes_only_its_output(N, M, d, m, scale, p):
    X, Y, W = cu(gauss(0, N, d)), cu(gauss(1, M, d)), cu(gauss(2, M, m))
    for rr, rows in ((nat.REDUCE_NONE, N), (nat.REDUCE_SUM, 1)):
        full = torch.full((rows + 6, m), SENT, dtype=torch.float32, device="cuda:0")   # jrk_kmatw_f32 writes dense rows (ldo = m)
        view = full[3:3 + rows]
        nat.kmatw(X, Y, W, scale, p, 0.3, row_reduce=rr, out=view)
        torch.cuda.synchronize()
        assert bool((full[:3] == SENT).all()) and bool((full[3 + rows:] == SENT).all())
        assert not bool((view == SENT).any()) and bool(torch.isfinite(view).all())


@pytest.mark.parametrize("N,M,d,m", [(130, 300, 16, 16), (257, 129, 3, 5), (64, 200, 40, 32)])
def test_backward_kernels_produce_finite_full_outputs(N, M, d, m):
    X, Y, W, G = cu(gauss(0, N, d)), cu(gauss(1, M, d)), cu(gauss(2, M, m)), cu(gauss(3, N, m))
    for p in (2.0, 1.0, 1.5):
        gx, gs = nat.kmatw_grad(X, Y, W, G, 0.3, p, 0.2)
        assert tuple(gx.shape) == (N, d) and tuple(gs.shape) == (N,)
        assert bool(torch.isfinite(gx).all()) and bool(torch.isfinite(gs).all())
        gx2, gs2 = nat.gram_grad(X, Y, cu(gauss(4, N, M)), 0.3, p, 0.2)
        assert bool(torch.isfinite(gx2).all()) and bool(torch.isfinite(gs2).all())
        # deterministic (fixed-order reductions, no float atomics)
        gx3, gs3 = nat.kmatw_grad(X, Y, W, G, 0.3, p, 0.2)
        assert torch.equal(gx, gx3) and torch.equal(gs, gs3)

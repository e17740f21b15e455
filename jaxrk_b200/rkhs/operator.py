"""Finite-rank RKHS operators -- drop-in for jaxrk/rkhs/operator.py.

The Gram / K@W work inside `__matmul__` (operator.py:37,46) and `Cov_inv` (:138) goes
through FiniteVec.inner and therefore through the fused kernels; the Tikhonov-regularised
solve (np.linalg.inv, dense `@`) stays on the library path (torch.linalg / cuBLAS), as
BASELINE.json's north_star specifies.  `Cov_solve` computes inner(inp_feat) twice in the
reference (:138 and :46); the self-Gram is cached per FiniteVec here.
"""
from typing import Union

import numpy as np
import torch

from .._array import asarray
from ..reduce.base import Sum
from ..reduce.lincomb import KernelMapReduce, LinearReduce
from .base import LinOp
from .vector import FiniteVec, inner


def _is_array(x):
    return isinstance(x, (torch.Tensor, np.ndarray))


class FiniteOp(LinOp):
    """Finite rank LinOp in RKHS: sum_ij matr[i, j] outp_feat_i (x) inp_feat_j."""

    def __init__(self, inp_feat, outp_feat, matr=None, normalize: bool = False):
        if matr is not None:
            matr = asarray(matr)
            assert tuple(matr.shape) == (len(outp_feat), len(inp_feat))
        self.matr = matr
        self.inp_feat, self.outp_feat = inp_feat, outp_feat
        self._normalize = normalize

    def __len__(self):
        return len(self.inp_feat) * len(self.outp_feat)

    def __matmul__(self, right_inp):
        """op @ op composes (operator.py:34-40); op @ vec / op @ points applies the operator (operator.py:42-56)."""
        if isinstance(right_inp, FiniteOp):
            return self._compose(right_inp)
        if _is_array(right_inp):
            right_inp = FiniteVec(self.inp_feat.k, torch.atleast_2d(asarray(right_inp)))
        return self._apply_to(right_inp)

    def _compose(self, other: "FiniteOp") -> "FiniteOp":
        # (self.matr) . <inp_feat, other.outp_feat> . (other.matr): the middle factor is a fused inner product
        core = self.inp_feat.inner(other.outp_feat) @ other.matr
        return FiniteOp(other.inp_feat, self.outp_feat, core if self.matr is None else self.matr @ core)

    def _apply_to(self, vec) -> FiniteVec:
        # coefficients over the output features: (matr . <inp_feat, vec>)^T = <vec, matr-reduced inp_feat>, i.e. ONE fused
        # K(V, X) @ matr^T with the test points as rows (never an N x T Gram followed by a dense product as in
        # operator.py:46).  The map stays in factored form (KernelMapReduce): get_mean_var, point evaluations and further
        # applies contract it with a skinny right-hand side; reading `.linear_map` materialises it.
        inp = self.inp_feat if self.matr is None else self.inp_feat.extend_reduce([LinearReduce(self.matr)])
        chain = [KernelMapReduce(vec, inp, normalize=self._normalize)]
        if len(vec) == 1:
            chain.append(Sum())
        return self.outp_feat.extend_reduce(chain)

    def inner(self, Y=None, full=True):
        if Y is None:
            Y = self
        G_i = self.inp_feat.inner(Y.inp_feat)
        G_o = self.outp_feat.inner(Y.outp_feat)
        if Y.matr.numel() < self.matr.numel():
            return torch.sum((G_o.T @ self.matr @ G_i) * Y.matr)
        return torch.sum((G_o @ Y.matr @ G_i.T) * self.matr)

    def reduce_gram(self, gram, axis=0):
        return gram

    @property
    def T(self) -> "FiniteOp":
        return FiniteOp(self.outp_feat, self.inp_feat, self.matr.T, self._normalize)

    def __call__(self, inp):
        return self @ FiniteVec(self.inp_feat.k, torch.atleast_2d(asarray(inp)))

    def apply(self, inp):
        return self @ inp


def _eye_over_n(n, dev):
    return torch.eye(n, dtype=torch.float32, device=dev) / n


def _dev_of(feat):
    return feat.insp_pts.device


def CrossCovOp(inp_feat, outp_feat) -> FiniteOp:
    assert len(inp_feat) == len(outp_feat)
    return FiniteOp(inp_feat, outp_feat, _eye_over_n(len(inp_feat), _dev_of(inp_feat)))


def CovOp(inp_feat) -> FiniteOp:
    return FiniteOp(inp_feat, inp_feat, _eye_over_n(len(inp_feat), _dev_of(inp_feat)))


def CovOp_from_Samples(kern, inspace_points, prefactors=None) -> FiniteOp:
    return CovOp(FiniteVec(kern, inspace_points, prefactors))


def Cov_regul(nsamps: int, nrefsamps: int = None, a: float = 0.49999999999999, b: float = 0.49999999999999,
              c: float = 0.1):
    """Regulariser of Schuster et al. 2020, Cor. 3.4 (operator.py:102-126)."""
    if nrefsamps is None:
        nrefsamps = nsamps
    assert 0 < a < 0.5 and 0 < b < 0.5 and 0 < c < 1
    assert nsamps > 0 and nrefsamps > 0
    return max(nrefsamps ** (-b * c), nsamps ** (-2 * a * c))


def Cov_inv(cov: FiniteOp, regul=None) -> FiniteOp:
    """(C + regul)^-1 as an operator (operator.py:128-149): Gram by the fused kernel, the
    inverse by the dense library path."""
    assert regul is not None
    gram = inner(cov.inp_feat)
    n = len(cov.inp_feat)
    regul = torch.as_tensor(regul, dtype=torch.float32, device=gram.device)
    inv_gram = torch.linalg.inv(gram + regul * torch.eye(n, dtype=torch.float32, device=gram.device))
    matr = inv_gram @ inv_gram
    if cov.matr is not None:
        matr = torch.linalg.inv(cov.matr) @ matr
    return FiniteOp(cov.inp_feat, cov.inp_feat, matr)


def Cov_solve(cov: FiniteOp, lhs, regul=None):
    """operator.py:151-171."""
    if not isinstance(lhs, FiniteOp) and _is_array(lhs):
        lhs = FiniteVec(cov.inp_feat.k, torch.atleast_2d(asarray(lhs)))
    if regul is None:
        regul = Cov_regul(1, len(cov.inp_feat))
    return Cov_inv(cov, regul) @ lhs


def _check_regul(regul, n):
    if regul is not None:
        regul = np.array(regul.detach().cpu().numpy() if isinstance(regul, torch.Tensor) else regul, dtype=np.float32)
        assert regul.squeeze().size == 1 or regul.squeeze().shape[0] == n
    return regul


def Cmo(inp_feat, outp_feat, regul=None) -> FiniteOp:
    """Conditional mean operator (operator.py:173-177)."""
    regul = _check_regul(regul, len(inp_feat))
    return CrossCovOp(Cov_solve(CovOp(inp_feat), inp_feat, regul=regul), outp_feat)


def Cdo(inp_feat, outp_feat, ref_feat, regul=None) -> FiniteOp:
    """Conditional density operator (operator.py:179-185)."""
    regul = _check_regul(regul, len(inp_feat))
    mo = Cmo(inp_feat, outp_feat, regul)
    return Cov_solve(CovOp(ref_feat), mo, regul=regul)

"""FiniteVec and inner -- drop-in for jaxrk/rkhs/vector.py (the consumer / fusion point
of the hot path).

A FiniteVec is (kernel, input-space points, reduce chain).  `inner` is
R_X . K(X, Y) . R_Y^T (vector.py:62-75); the reference materialises K and applies the
chains, here `_lowering.reduced_gram` maps the triple onto the fused K@W / group-sum /
Gram kernels of libjaxrk_b200.so.
"""
from copy import copy
from typing import List, Union

import torch

from .._array import as2d, asarray
from ..kern.base import Kernel
from ..reduce.base import Center, Mean, Prefactors, Reduce, Sum
from ..reduce.lincomb import LinearReduce
from . import _lowering
from .base import Vec


class FiniteVec(Vec):
    """RKHS feature vector(s) built from input space points."""

    def __init__(self, k: Kernel, insp_pts, reduce: List[Reduce] = []):
        if not isinstance(insp_pts, Vec):
            insp_pts = as2d(insp_pts)
        self.reduce = [] if reduce is None else reduce
        self.k = k
        self.insp_pts = insp_pts
        self.__len = Reduce.final_len(len(self.insp_pts), self.reduce)
        self._raw_gram_cache = None
        self._inner_cache = None

    def __len__(self):
        return self.__len

    def inner(self, Y=None, full=True):
        if not full and Y is not None:
            raise ValueError("Ambiguous inputs: `full` and `y` are not compatible.")
        if not full:  # diagonal of the self inner product (the reference's branch is dead: vector.py:66-67)
            if self._inner_cache is None:       # without the full reduced Gram where the chain allows it
                dg = _lowering.reduced_gram_diag(self.k, self.insp_pts, self.reduce)
                if dg is not None:
                    return dg
            return torch.diagonal(self.inner())
        if Y is not None:
            assert self.k == Y.k  # object identity, as in the reference (quirk Q5)
        else:
            Y = self
        self_gram = Y is self or Y.insp_pts is self.insp_pts
        if Y is self:  # Cov_solve asks for inner(inp_feat) twice (operator.py:138 and :46): keep it
            # the key covers in-place updates of the points / chain tensors (an optimiser step bumps `_version`) and
            # the autograd mode (a result computed under no_grad must not be handed to a recording caller)
            key = (tuple(id(r) for r in self.reduce), self.insp_pts._version,
                   tuple(t._version for t in _lowering._chain_tensors(self.reduce)), torch.is_grad_enabled())
            if self._inner_cache is None or self._inner_cache[0] != key:
                res = _lowering.reduced_gram(self.k, self.insp_pts, self.reduce, self.insp_pts, self.reduce, True)
                if res.requires_grad:   # part of an autograd graph: not cached (a graph is backpropagated once)
                    return res
                self._inner_cache = (key, res)
            return self._inner_cache[1]
        return _lowering.reduced_gram(self.k, self.insp_pts, self.reduce, Y.insp_pts, Y.reduce, self_gram)

    def normalized(self) -> "FiniteVec":
        r = self.reduce
        if isinstance(r[-1], Prefactors):
            return self.updated(r[-1].prefactors / torch.sum(r[-1].prefactors))
        elif isinstance(r[-1], LinearReduce):
            return self.updated(r[-1].linear_map / torch.sum(r[-1].linear_map, 1, keepdim=True))
        assert False

    def updated(self, prefactors) -> "FiniteVec":
        prefactors = asarray(prefactors)
        _r = copy(self.reduce)
        previous_reduction = None
        if len(_r) > 0 and isinstance(_r[-1], (Prefactors, LinearReduce)):
            final_len = Reduce.final_len(len(self.insp_pts), _r)
            assert (final_len == prefactors.shape[0]) or (final_len == 1 and prefactors.dim() == 1)
            previous_reduction = _r[-1]
            _r = _r[:-1]
        if prefactors.dim() == 1:
            assert previous_reduction is None or isinstance(previous_reduction, Prefactors)
            _r.append(Prefactors(prefactors))
        elif prefactors.dim() == 2:
            assert previous_reduction is None or isinstance(previous_reduction, LinearReduce)
            assert len(self) == prefactors.shape[0]
            _r.append(LinearReduce(prefactors))
        return FiniteVec(self.k, self.insp_pts, _r)

    def centered(self) -> "FiniteVec":
        return self.extend_reduce([Center()])

    def extend_reduce(self, r: List[Reduce]) -> "FiniteVec":
        if r is None or len(r) == 0:
            return self
        _r = copy(self.reduce)
        _r.extend(r)
        return FiniteVec(self.k, self.insp_pts, _r)

    def reduce_gram(self, gram, axis=0):
        return Reduce.apply(gram, self.reduce, axis)

    def nsamps(self, mean=False):
        n = len(self.insp_pts)
        rval = Reduce.apply(torch.ones((n, 1), dtype=torch.float32, device=self.insp_pts.device) * n, self.reduce, 0)
        return rval.mean() if mean else rval

    def get_mean_var(self, keepdims=False):
        mean = self.reduce_gram(self.insp_pts, 0)
        variance_of_expectations = self.reduce_gram(self.insp_pts ** 2, 0) - mean ** 2
        var = torch.as_tensor(self.k.var(), device=mean.device) + variance_of_expectations
        if keepdims:
            return (mean, var)
        return (torch.squeeze(mean), torch.squeeze(var))

    def sum(self, use_linear_reduce=False) -> "FiniteVec":
        if use_linear_reduce:
            return self.extend_reduce([LinearReduce(torch.ones((1, len(self)), dtype=torch.float32))])
        return self.extend_reduce([Sum()])

    def mean(self) -> "FiniteVec":
        return self.extend_reduce([Mean()])

    @classmethod
    def construct_RKHS_Elem(cls, kern, inspace_points, prefactors=None) -> "FiniteVec":
        p = asarray(prefactors).squeeze()
        assert p.dim() == 1
        return FiniteVec(kern, inspace_points, [LinearReduce(p[None, :])])

    @property
    def _raw_gram(self):
        if self._raw_gram_cache is None:
            self._raw_gram_cache = self.k(self.insp_pts)
        return self._raw_gram_cache

    def point_representant(self, method: str = "inspace_point", keepdims: bool = False):
        """An input-space point standing for each element (vector.py:161-174): the support point whose feature is
        nearest in RKHS norm ("inspace_point"; the Grams come from the fused kernels), or the mean ("mean")."""
        from ..utilities.gram import gram_projection
        if method == "inspace_point":
            assert isinstance(self.reduce[-1], (Prefactors, LinearReduce))
            G_orig_repr = Reduce.apply(self._raw_gram, self.reduce, 1)
            idx = gram_projection(G_orig_repr, Reduce.apply(G_orig_repr, self.reduce, 0), self._raw_gram,
                                  method="representer").reshape(-1)
            rval = self.insp_pts[idx, :]
        elif method == "mean":
            rval = self.get_mean_var(keepdims=keepdims)[0]
        else:
            assert False, "No known method selected for point_representant"
        return rval if keepdims else rval.squeeze()

    def __call__(self, argument):
        return inner(self, FiniteVec(self.k, argument, []))


class _ProductSqExp:
    """The product of two squared-exponential GenGauss kernels on the column blocks of concatenated, pre-scaled
    inputs: exp(-sR^2 |xR - yR|^2) exp(-sL^2 |xL - yL|^2) = exp(-|[sR xR, sL xL] - [sR yR, sL yL]|^2) -- again a
    GenGauss kernel (scale 1, power 2), so CombVec products run through the fused kernels instead of two materialised
    Grams."""

    def __init__(self, nconst: float):
        self._nconst = float(nconst)

    scale_tensor = None

    def kernel_args(self):
        return 1.0, None, 2.0, self._nconst

    def wants_grad(self, *tensors) -> bool:
        """As GenGaussKernel.wants_grad: the concatenated, pre-scaled inputs carry the graph of both factors."""
        return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _is_product(op) -> bool:
    import operator
    import numpy as np
    return op in (operator.mul, torch.mul, torch.multiply, np.multiply)


class CombVec(Vec):
    """Elementwise combination (e.g. product) of two vectors' Gram matrices, then a reduce chain -- drop-in for
    jaxrk/rkhs/vector.py:227-268.  inner = reduce_0(reduce_1(operation(vR.inner(Y.vR), vL.inner(Y.vL)))).
    When the operation is a product and both factors are plain FiniteVecs over squared-exponential kernels, the
    combined Gram is itself a squared-exponential Gram on concatenated inputs (_ProductSqExp) and the whole inner
    product, reduce chain included, is lowered onto ONE fused kernel call; otherwise the two (reduced) Grams are
    computed by the fused kernels and combined with `operation`."""

    def __init__(self, vR, vL, operation, reduce: List[Reduce] = []):
        assert len(vR) == len(vL)
        self.vR, self.vL = vR, vL
        self.reduce = [] if reduce is None else reduce
        self.operation = operation
        self.__len = Reduce.final_len(len(self.vR), self.reduce)

    def reduce_gram(self, gram, axis=0):
        return Reduce.apply(gram, self.reduce, axis)

    def _fusable_with(self, Y) -> bool:
        def plain_sqexp(v):
            if not isinstance(v, FiniteVec) or v.reduce or not hasattr(v.k, "kernel_args"):
                return False
            a = v.k.kernel_args()
            return a is not None and a[1] is None and a[2] == 2.0
        return (_is_product(self.operation) and all(plain_sqexp(v) for v in (self.vR, self.vL, Y.vR, Y.vL))
                and self.vR.k is Y.vR.k and self.vL.k is Y.vL.k)

    def _concat(self):
        (sR, _, _, ncR), (sL, _, _, ncL) = self.vR.k.kernel_args(), self.vL.k.kernel_args()
        return torch.cat([sR * self.vR.insp_pts, sL * self.vL.insp_pts], 1), ncR * ncL

    def inner(self, Y: "CombVec" = None, full=True):
        if Y is None:
            Y = self
        else:
            assert Y.operation == self.operation
        if self._fusable_with(Y):
            Xc, nc = self._concat()
            Yc = Xc if Y is self else Y._concat()[0]
            return _lowering.reduced_gram(_ProductSqExp(nc), Xc, self.reduce, Yc, Y.reduce, Y is self)
        G = self.operation(self.vR.inner(Y.vR), self.vL.inner(Y.vL))
        return self.reduce_gram(Y.reduce_gram(G, 1), 0)

    def extend_reduce(self, r: List[Reduce]) -> "CombVec":
        if r is None or len(r) == 0:
            return self
        _r = copy(self.reduce)
        _r.extend(r)
        return CombVec(self.vR, self.vL, self.operation, _r)

    def centered(self) -> "CombVec":
        return self.extend_reduce([Center()])

    def __len__(self):
        return self.__len

    def updated(self, prefactors):
        raise NotImplementedError()


def inner(X, Y=None, full=True):
    return X.inner(Y, full)

"""Reduce maps -- drop-in for jaxrk/reduce/base.py.

Every element of a reduce chain is a LINEAR map on one axis of a Gram matrix.  On the hot
path (FiniteVec.inner, rkhs/vector.py) chains are not applied to a materialised Gram:
`jaxrk_b200.rkhs._lowering` folds Prefactors / Sum / Mean / BalancedRed / LinearReduce
into the fused K@W and group-sum kernels.  What is defined here is the array-level
semantics of the reference (Reduce.__call__(inp, axis), reduce_first_ax, new_len, the
classmethods apply / final_len; base.py:25-55), used for chains that do not fold (Center,
TileView, SparseReduce), for the small already-reduced results, and for direct use on
arbitrary arrays.

Layout of this module: a concrete map only states what it does to the FIRST axis of a
2-D tensor (`_first_axis`) and how long the result is (`_length`); moving another axis to
the front and back, the public method names of the reference, and chain folding live in
the base class.
"""
from typing import List, Optional

import torch

from .._array import asarray

__all__ = ["Reduce", "NoReduce", "Prefactors", "TileView", "Sum", "Mean", "BalancedRed", "Center"]


def _front(t: torch.Tensor, axis: int) -> torch.Tensor:
    """View of t with `axis` moved to position 0 (its own inverse for the 2-D arrays used here)."""
    return t if axis == 0 else t.transpose(0, axis)


class Reduce:
    """A linear map on one axis.  Subclasses inside this package implement `_first_axis` / `_length`; user code
    written against the reference may instead override `reduce_first_ax` / `new_len` directly."""

    # ---- what a concrete map provides ----
    def _first_axis(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()

    def _length(self, n: int) -> int:
        raise NotImplementedError()

    # ---- the reference's public surface (base.py:25-55) ----
    def reduce_first_ax(self, gram):
        return self._first_axis(asarray(gram))

    def new_len(self, original_len: int) -> int:
        return self._length(original_len)

    def __call__(self, inp, axis: int = 0):
        return _front(self.reduce_first_ax(_front(asarray(inp), axis)), axis)

    @classmethod
    def apply(cls, inp, reduce: Optional[List["Reduce"]] = None, axis=0):
        if not reduce:
            return inp
        t = _front(asarray(inp), axis)
        for r in reduce:
            t = r.reduce_first_ax(t)
        return _front(t, axis)

    @classmethod
    def final_len(cls, original_len: int, reduce: Optional[List["Reduce"]] = None):
        n = original_len
        for r in reduce or ():
            n = r.new_len(n)
        return n


class NoReduce(Reduce):
    """Identity (base.py:74-82)."""

    def __call__(self, inp, axis: int = 0):
        return inp

    def _first_axis(self, t):
        return t

    def _length(self, n):
        return n


class Prefactors(Reduce):
    """diag(prefactors): entry i of the axis is scaled by prefactors[i] (base.py:84-101)."""

    def __init__(self, prefactors):
        self.prefactors = asarray(prefactors)
        assert self.prefactors.dim() == 1

    def _first_axis(self, t):
        assert self.prefactors.shape[0] == t.shape[0]
        return self.prefactors.to(t.device)[:, None] * t

    def _length(self, n):
        assert n == len(self.prefactors)
        return n


class TileView(Reduce):
    """The whole axis repeated until it has result_len entries (base.py:118-129)."""

    def __init__(self, result_len: int):
        assert result_len > 0
        self.result_len = result_len

    def _first_axis(self, t):
        reps, rest = divmod(self.result_len, t.shape[0])
        assert rest == 0, "Input can't be broadcasted to target length %d" % self.result_len
        return t.repeat(reps, 1)

    def _length(self, n):
        return self.result_len


class Sum(Reduce):
    """1^T: the axis collapses to one entry, kept as an axis of length 1 (base.py:132-140)."""

    def _first_axis(self, t):
        return t.sum(dim=0, keepdim=True)

    def _length(self, n):
        return 1


class Mean(Reduce):
    """1^T / n (base.py:142-150)."""

    def _first_axis(self, t):
        return t.mean(dim=0, keepdim=True)

    def _length(self, n):
        return 1


class BalancedRed(Reduce):
    """I (x) 1^T: consecutive groups of points_per_split entries are summed, or averaged (base.py:152-181)."""

    def __init__(self, points_per_split: int, average=False):
        assert points_per_split > 0
        self.points_per_split = points_per_split
        self.average = bool(average)

    def _first_axis(self, t):
        g = t.reshape(t.shape[0] // self.points_per_split, self.points_per_split, *t.shape[1:])
        return g.mean(1) if self.average else g.sum(1)

    def _length(self, n):
        assert n % self.points_per_split == 0
        return n // self.points_per_split


class Center(Reduce):
    """I - 1 1^T / n: the mean along the axis is removed (base.py:183-192)."""

    def _first_axis(self, t):
        return t - t.mean(dim=0, keepdim=True)

    def _length(self, n):
        return n

"""Reduce maps -- drop-in for jaxrk/reduce/base.py.

Every element of a reduce chain is a LINEAR map on one axis of a Gram matrix.  On the hot
path (FiniteVec.inner, rkhs/vector.py) chains are not applied to a materialised Gram:
`jaxrk_b200.rkhs._lowering` folds Prefactors / Sum / Mean / BalancedRed / LinearReduce
into the fused K@W and group-sum kernels.  The methods below are the reference's
array-level semantics (Reduce.__call__(inp, axis), reduce_first_ax, new_len, classmethods
apply / final_len; base.py:25-55) and serve chains that do not fold (Center, TileView,
SparseReduce) and direct use on arbitrary arrays.
"""
from abc import ABC, abstractmethod
from typing import List

import torch

from .._array import asarray


class Reduce(ABC):
    """The basic reduction type."""

    def __call__(self, inp, axis: int = 0):
        rval = self.reduce_first_ax(torch.swapaxes(asarray(inp), axis, 0))
        return torch.swapaxes(rval, axis, 0)

    @abstractmethod
    def reduce_first_ax(self, gram):
        pass

    @abstractmethod
    def new_len(self, original_len: int) -> int:
        pass

    @classmethod
    def apply(cls, inp, reduce: List["Reduce"] = None, axis=0):
        if reduce is None or len(reduce) == 0:
            return inp
        carry = torch.swapaxes(asarray(inp), axis, 0)
        for gr in reduce:
            carry = gr.reduce_first_ax(carry)
        return torch.swapaxes(carry, axis, 0)

    @classmethod
    def final_len(cls, original_len: int, reduce: List["Reduce"] = None):
        carry = original_len
        if reduce is not None:
            for gr in reduce:
                carry = gr.new_len(carry)
        return carry


class NoReduce(Reduce):
    def __call__(self, inp, axis: int = 0):
        return inp

    def reduce_first_ax(self, inp):
        return inp

    def new_len(self, original_len: int) -> int:
        return original_len


class Prefactors(Reduce):
    """Scale entry i of the axis by prefactors[i] (base.py:84-101)."""

    def __init__(self, prefactors):
        self.prefactors = asarray(prefactors)
        assert self.prefactors.dim() == 1

    def __call__(self, inp, axis: int = 0):
        inp = asarray(inp)
        assert self.prefactors.shape[0] == inp.shape[axis]
        return inp * self.prefactors.to(inp.device).unsqueeze((axis + 1) % 2)

    def reduce_first_ax(self, inp):
        return self.__call__(inp, 0)

    def new_len(self, original_len: int) -> int:
        assert original_len == len(self.prefactors)
        return original_len


class TileView(Reduce):
    """Repeat the whole axis until it has result_len entries (base.py:118-129)."""

    def __init__(self, result_len: int):
        assert result_len > 0
        self.result_len = result_len

    def reduce_first_ax(self, inp):
        assert self.result_len % inp.shape[0] == 0, "Input can't be broadcasted to target length %d" % self.result_len
        return inp.repeat(self.result_len // inp.shape[0], 1)

    def new_len(self, original_len: int) -> int:
        return self.result_len


class Sum(Reduce):
    def __call__(self, inp, axis: int = 0):
        return asarray(inp).sum(dim=axis, keepdim=True)

    def reduce_first_ax(self, inp):
        return self.__call__(inp, 0)

    def new_len(self, original_len: int) -> int:
        return 1


class Mean(Reduce):
    def __call__(self, inp, axis: int = 0):
        return asarray(inp).mean(dim=axis, keepdim=True)

    def reduce_first_ax(self, inp):
        return self.__call__(inp, 0)

    def new_len(self, original_len: int) -> int:
        return 1


class BalancedRed(Reduce):
    """Sum (or average) consecutive groups of points_per_split entries (base.py:152-181)."""

    def __init__(self, points_per_split: int, average=False):
        assert points_per_split > 0
        self.points_per_split = points_per_split
        self.average = bool(average)

    def __call__(self, inp, axis: int = 0):
        inp = torch.swapaxes(asarray(inp), axis, 0)
        grouped = inp.reshape(-1, self.points_per_split, inp.shape[-1])
        out = grouped.mean(1) if self.average else grouped.sum(1)
        return torch.swapaxes(out, axis, 0)

    def reduce_first_ax(self, inp):
        return self.__call__(inp, 0)

    def new_len(self, original_len: int) -> int:
        assert original_len % self.points_per_split == 0
        return original_len // self.points_per_split


class Center(Reduce):
    """Subtract the mean along the axis (base.py:183-192)."""

    def __call__(self, inp, axis: int = 0):
        inp = asarray(inp)
        return inp - inp.mean(dim=axis, keepdim=True)

    def reduce_first_ax(self, inp):
        return self.__call__(inp, 0)

    def new_len(self, original_len: int) -> int:
        return original_len

"""LinearReduce / SparseReduce -- drop-in for jaxrk/reduce/lincomb.py.

`LinearReduce(A)` is `A @ gram` (lincomb.py:131-144): on the Y side of an inner product
it is the W = A^T of the fused, never-materialised K(X, Y) @ W kernel.
"""
from typing import List

import numpy as np
import torch

from .._array import asarray
from .base import Reduce


class LinearReduce(Reduce):
    def __init__(self, linear_map):
        self.linear_map = asarray(linear_map)

    def reduce_first_ax(self, inp):
        assert len(inp.shape) == 2
        assert self.linear_map.shape[1] == inp.shape[0]
        return self.linear_map.to(inp.device) @ inp

    def new_len(self, original_len: int):
        assert self.linear_map.shape[1] == original_len, \
            self.__class__.__name__ + " expects a gram with %d columns" % self.linear_map.shape[1]
        return self.linear_map.shape[0]

    @classmethod
    def sum_from_unique(cls, input, mean: bool = True):
        """lincomb.py:146-154: one output row per unique value of `input`."""
        inp = np.asarray(input.detach().cpu().numpy() if isinstance(input, torch.Tensor) else input)
        un, inverse, cts = np.unique(inp, return_inverse=True, return_counts=True)
        m = np.zeros((un.size, inp.shape[0]), np.float32)
        m[inverse.reshape(-1), np.arange(inp.shape[0])] = 1.
        if mean:
            m /= cts[:, None]
        return un, cts, LinearReduce(m)


class KernelMapReduce(LinearReduce):
    """A LinearReduce whose map is itself an inner product,  linear_map = R_V K(V, X) R_X^T  (len(V) x len(X side)) --
    exactly what FiniteOp.__matmul__ builds for `op @ vec` (jaxrk/rkhs/operator.py:46-52: `LinearReduce(lin_map.T)` with
    lin_map = matr @ inner(inp_feat, vec)) -- kept in FACTORED form:

      * `reduce_first_ax(a)` = linear_map @ a is evaluated as ONE fused, skinny K(V, X) @ (R_X^T a): this is the
        re-association of FiniteVec.get_mean_var (jaxrk/rkhs/vector.py:126-134) for the result of an operator apply --
        `L @ Y_ref` never materialises the len(V) x len(X) map;
      * `linear_map` (the attribute users and `inner` read) materialises the map on first access, through the same
        fused path (K evaluated once; wide maps go through the row-slab route of rkhs/_lowering.py), and keeps it.

    `normalize` reproduces FiniteOp's `lin_map / lin_map.sum(1, keepdims=True)` (operator.py:49-50)."""

    def __init__(self, vec, inp, normalize: bool = False):
        self._vec, self._inp, self._normalize = vec, inp, normalize
        self._map = None
        self._col_scale = None

    def _scale(self):
        if self._normalize and self._col_scale is None:
            from ..rkhs.vector import inner
            self._col_scale = 1.0 / inner(self._vec.sum(), self._inp)        # 1 x len(inp): column sums of the map
        return self._col_scale

    @property
    def linear_map(self):
        if self._map is None:
            from ..rkhs.vector import inner
            m = inner(self._vec, self._inp)
            if self._normalize:
                m = m * self._scale()
            self._map = m
        return self._map

    @property
    def is_materialised(self) -> bool:
        return self._map is not None

    def reduce_first_ax(self, inp):
        assert len(inp.shape) == 2
        assert len(self._inp) == inp.shape[0]
        if self._map is not None:
            return self._map.to(inp.device) @ inp
        from ..rkhs.vector import inner
        a = asarray(inp)
        if self._normalize:
            a = self._scale().reshape(-1, 1).to(a.device) * a
        return inner(self._vec, self._inp.extend_reduce([LinearReduce(a.T.contiguous())]))

    def new_len(self, original_len: int):
        assert len(self._inp) == original_len, \
            self.__class__.__name__ + " expects a gram with %d columns" % len(self._inp)
        return len(self._vec)


class BlockReduce(Reduce):
    """Sum / average consecutive blocks of rows: block b is rows block_boundaries[b] .. block_boundaries[b + 1] - 1
    (jaxrk/reduce/lincomb.py:89-129).  The reference's reduce_first_ax concatenates the per-block 1-D sums into ONE 1-D
    array and its new_len returns the last boundary (lincomb.py:114-119): unusable inside a chain, and no reference test
    exercises it (tests/test_reduce.py:31 is commented out).  Implemented here with the evident intent -- one output row
    per block -- so that `sum_from_block` composes like `LinearReduce.sum_from_unique`.  Few blocks are folded into the
    fused K@W kernel as a dense map (rkhs/_lowering.py); many blocks reduce a materialised Gram by prefix sums."""

    def __init__(self, block_boundaries, average: bool = True):
        b = np.asarray(block_boundaries.detach().cpu().numpy() if isinstance(block_boundaries, torch.Tensor) else block_boundaries)
        b = b.astype(np.int64).reshape(-1)
        assert b.size >= 2 and (np.diff(b) > 0).all() and b[0] >= 0
        self.block_boundaries = b
        self.average = bool(average)

    def num_rows(self) -> int:
        return int(self.block_boundaries.size - 1)

    def reduce_first_ax(self, inp):
        inp = asarray(inp)
        assert len(inp.shape) == 2
        assert self.block_boundaries[-1] <= len(inp), self.__class__.__name__ + " expects a longer gram to operate on"
        b = torch.as_tensor(self.block_boundaries, device=inp.device)
        out = torch.empty((self.num_rows(), inp.shape[1]), dtype=inp.dtype, device=inp.device)
        # block sums as differences of an exclusive prefix sum (float64 accumulation: the differences cancel), a slab of
        # columns at a time so that the float64 scratch stays at ~1 GiB whatever the Gram's size
        step = max(1, (1 << 27) // (inp.shape[0] + 1))
        for c0 in range(0, inp.shape[1], step):
            blk = inp[:, c0:c0 + step]
            cs = torch.zeros((blk.shape[0] + 1, blk.shape[1]), dtype=torch.float64, device=inp.device)
            torch.cumsum(blk.double(), 0, out=cs[1:])
            o = cs[b[1:]] - cs[b[:-1]]
            if self.average:
                o = o / (b[1:] - b[:-1]).double()[:, None]
            out[:, c0:c0 + step] = o.to(inp.dtype)
        return out

    def new_len(self, original_len: int):
        assert self.block_boundaries[-1] <= original_len
        return self.num_rows()

    def to_linear_map(self, n: int, device=None) -> torch.Tensor:
        assert self.block_boundaries[-1] <= n
        A = torch.zeros((self.num_rows(), n), dtype=torch.float32, device=device)
        for r, (s, e) in enumerate(zip(self.block_boundaries[:-1], self.block_boundaries[1:])):
            A[r, int(s):int(e)] = (1.0 / float(e - s)) if self.average else 1.0
        return A

    @classmethod
    def sum_from_block(cls, input, mean: bool = True) -> "BlockReduce":
        """lincomb.py:121-128: a block ends wherever consecutive entries of `input` differ."""
        inp = np.asarray(input.detach().cpu().numpy() if isinstance(input, torch.Tensor) else input).reshape(-1)
        change = [0] + list(np.argwhere(inp[:-1] != inp[1:]).flatten() + 1) + [len(inp)]
        return BlockReduce(np.array(change), mean)


class SparseReduce(Reduce):
    """Sum / average rows of the input selected by index arrays (lincomb.py:27-87).  Index
    structured, so it is applied to a materialised Gram (not folded into the fused kernels)."""

    def __init__(self, idcs: List, average: bool = True, max_idx=None):
        self.idcs = [torch.as_tensor(np.asarray(i), dtype=torch.long) for i in idcs]
        self.average = average
        self.max_idx = max_idx if max_idx is not None else max(int(i.max()) for i in self.idcs if i.numel())

    def reduce_first_ax(self, inp):
        assert self.max_idx + 1 <= len(inp)
        assert len(inp.shape) == 2
        rval = []
        for idx in self.idcs:
            if idx.shape[1] == 0:
                rval.append(torch.zeros((idx.shape[0], inp.shape[1]), dtype=inp.dtype, device=inp.device))
            else:
                g = inp[idx.reshape(-1).to(inp.device), :].reshape(-1, idx.shape[1], inp.shape[1])
                rval.append(g.mean(1) if self.average else g.sum(1))
        return torch.cat(rval, 0)

    def new_len(self, original_len: int):
        assert self.max_idx + 1 <= original_len
        return len(self.idcs)  # as the reference (lincomb.py:62-64)

    def num_rows(self) -> int:
        """Rows reduce_first_ax actually produces (one per row of every index array)."""
        return int(sum(i.shape[0] for i in self.idcs))

    def to_linear_map(self, n: int, device=None) -> torch.Tensor:
        """The same map as a dense (rows x n) matrix: what rkhs/_lowering.py folds into the fused K@W kernel (as
        LinearReduce) when it is small enough, so that the Gram need not be materialised."""
        assert self.max_idx + 1 <= n
        A = torch.zeros((self.num_rows(), n), dtype=torch.float32, device=device)
        r0 = 0
        for idx in self.idcs:
            g, k = idx.shape
            if k:
                rows = torch.arange(r0, r0 + g, device=device)[:, None].expand(g, k).reshape(-1)
                vals = torch.full((g * k,), (1.0 / k) if self.average else 1.0, dtype=torch.float32, device=device)
                A.index_put_((rows, idx.reshape(-1).to(device)), vals, accumulate=True)
            r0 += g
        return A

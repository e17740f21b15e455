"""Row-sharded evaluation of the hot path on several GPUs of one box.

One process per GPU (torchrun); X is split into contiguous row blocks, Y / W and the kernel
parameters are replicated.  Every output row of K and of K @ W depends only on its own x_i,
so the Gram and the un-reduced product need NO collective -- the result stays row-sharded.
A collective appears only where the reference's reduce chain mixes rows:

    X-side fold        local work                              exchange (torch.distributed / NCCL)
    none, Prefactors   local rows                              none (all_gather on request)
    BalancedRed(k)     local groups (blocks aligned to k)      none (all_gather on request)
    Sum / Mean         local sum                               all_reduce(SUM) of 1 x m; Mean / N_global
    LinearReduce(A)    A[:, local rows] @ local                all_reduce(SUM) of m' x m

Whatever follows the folded part of the chain (R_post, see rkhs/_lowering.py) acts on the
complete result after the exchange.  The kernels are the same single-GPU kernels on
N_local = N / world rows.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .reduce.base import BalancedRed, Mean, Prefactors, Reduce, Sum
from .reduce.lincomb import LinearReduce
from .rkhs import _lowering as low
from .rkhs.vector import FiniteVec


def row_block(n: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous block of rows [start, stop) owned by `rank`; block edges are multiples of `align`
    (the BalancedRed group size when a grouped reduce is folded)."""
    assert n % align == 0, "the group size must divide the number of rows"
    units = n // align
    per, rem = divmod(units, world)
    start = rank * per + min(rank, rem)
    stop = start + per + (1 if rank < rem else 0)
    return start * align, stop * align


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def chain_alignment(chain: Optional[List[Reduce]]) -> int:
    """Row-block alignment a chain needs: the group size of a leading BalancedRed (after Prefactors)."""
    for r in chain or []:
        if type(r) is Prefactors:
            continue
        return r.points_per_split if type(r) is BalancedRed else 1
    return 1


def local_chain(chain: Optional[List[Reduce]], n_global: int, start: int, stop: int):
    """Restrict the foldable head of an X-side chain to rows [start, stop).
    Returns (local chain, exchange kind, post chain, mean divisor)."""
    chain = list(chain or [])
    out: List[Reduce] = []
    for idx, r in enumerate(chain):
        if type(r) is Prefactors:
            assert r.prefactors.shape[0] == n_global
            out.append(Prefactors(r.prefactors[start:stop]))
        elif type(r) is Sum:
            return out + [Sum()], "sum", chain[idx + 1:], 1.0
        elif type(r) is Mean:
            return out + [Sum()], "sum", chain[idx + 1:], float(n_global)
        elif type(r) is BalancedRed:
            assert start % r.points_per_split == 0 and stop % r.points_per_split == 0
            return out + [r], "rows", chain[idx + 1:], 1.0
        elif type(r) is LinearReduce:
            assert r.linear_map.shape[1] == n_global
            return out + [LinearReduce(r.linear_map[:, start:stop])], "sum", chain[idx + 1:], 1.0
        else:
            return out, "rows", chain[idx:], 1.0
    return out, "rows", [], 1.0


def gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of row blocks of possibly different heights (concatenated in rank order)."""
    rank, world = _world(group)
    if world == 1:
        return local
    local = local.contiguous()
    h = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h, group=group)
    hmax = int(max(int(x) for x in hs))
    pad = torch.zeros((hmax, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:int(n)] for b, n in zip(bufs, hs)], 0)


def sharded_gram(k, X_local, Y, gather: bool = False, group=None) -> torch.Tensor:
    """Rows of K(X, Y) owned by this rank (no collective); `gather=True` replicates the full Gram."""
    G = k(X_local, Y)
    return gather_rows(G, group) if gather else G


def sharded_inner(x: FiniteVec, y: FiniteVec, gather: bool = True, group=None) -> torch.Tensor:
    """inner(x, y) = R_X K(X, Y) R_Y^T with X row-sharded over the process group.

    `x.insp_pts` holds ALL points on every rank here only for the convenience of slicing; a rank touches its
    block alone (pass a vector built on the global points, or use `sharded_inner_local` with a local block).
    """
    rank, world = _world(group)
    n = x.insp_pts.shape[0]
    start, stop = row_block(n, rank, world, chain_alignment(x.reduce))
    return sharded_inner_local(x.k, x.insp_pts[start:stop], x.reduce, n, start, y, gather=gather, group=group)


def sharded_inner_local(k, X_local, chain_x: Optional[List[Reduce]], n_global: int, start: int, y: FiniteVec,
                        gather: bool = True, group=None) -> torch.Tensor:
    """As `sharded_inner`, given only this rank's row block X_local = X[start : start + len(X_local)]."""
    stop = start + X_local.shape[0]
    lchain, kind, post, divisor = local_chain(chain_x, n_global, start, stop)
    part = low.reduced_gram(k, X_local, lchain, y.insp_pts, y.reduce).contiguous()   # NCCL needs dense buffers
    _, world = _world(group)
    if kind == "sum":
        if world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        if divisor != 1.0:
            part = part / divisor
        full = part
    else:
        if not gather and not post:
            return part                     # row-sharded result, no exchange at all
        full = gather_rows(part, group)
    return Reduce.apply(full, post, 0)

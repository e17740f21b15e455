"""`dist` -- drop-in for jaxrk/utilities/distances.py:44-49 (Array branch -> eucldist;
Vec branch -> RKHS distances from the reduced Gram matrices, distances.py:28-42 over
utilities/gram.py:8-23)."""
import torch

from .eucldist import eucldist
from .gram import rkhs_gram_cdist

__all__ = ["dist", "rkhs_cdist", "rkhs_gram_cdist"]


def rkhs_cdist(a, b=None, power: float = 2.):
    """||a_i - b_j||_H^power (distances.py:28-42).  The reference builds THREE reduced Grams (a.inner(b), a.inner(),
    b.inner()) and uses only the diagonals of the last two; here the diagonals come from `inner(full=False)` -- block
    sums / one skinny fused call, not a reduced Gram -- and the combination  |a_i|^2 + |b_j|^2 - 2 <a_i, b_j>  is ONE pass
    over the cross Gram (jrk_gram_affine_f32)."""
    if b is None:
        return rkhs_gram_cdist(a.inner(), power=power)
    G_ab = a.inner(b)
    ga, gb = a.inner(full=False), b.inner(full=False)
    ga, gb = (torch.diagonal(g) if g.dim() == 2 else g for g in (ga, gb))     # (CombVec.inner has no diagonal-only form)
    if not G_ab.is_cuda or G_ab.requires_grad or ga.requires_grad or gb.requires_grad:
        sq = ga[:, None] + gb[None, :] - 2 * G_ab
    else:
        from .. import _native as nat
        if b is a:
            G_ab = G_ab.clone()                                                # the cached self inner product stays intact
        sq = nat.gram_affine(G_ab, row_mul=torch.full((G_ab.shape[0],), -2.0, device=G_ab.device),
                             row_add=ga.contiguous(), col_add=gb.contiguous())
    return sq if power == 2. else torch.pow(sq, power / 2.)


def dist(a, b=None, power: float = 2.):
    from ..rkhs.base import Vec
    if isinstance(a, Vec):
        return rkhs_cdist(a, b, power=power)
    return eucldist(a, b, power=power)

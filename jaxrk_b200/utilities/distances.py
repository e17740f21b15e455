"""`dist` -- drop-in for jaxrk/utilities/distances.py:44-49 (Array branch -> eucldist;
Vec branch -> RKHS distances from the reduced Gram matrices, distances.py:28-42 over
utilities/gram.py:8-23)."""
import torch

from .eucldist import eucldist

__all__ = ["dist", "rkhs_cdist", "rkhs_gram_cdist"]


def rkhs_gram_cdist(G_ab, G_a=None, G_b=None, power: float = 2.):
    """utilities/gram.py:8-23: ||a_i - b_j||_H^power from Gram matrices.  For a single
    Gram the reference asserts exact symmetry (gram.py:18); the self-Gram kernel
    guarantees it bitwise."""
    if G_a is None or G_b is None:
        assert G_ab.shape[0] == G_ab.shape[1] and bool(torch.all(G_ab == G_ab.T))
        da = db = torch.diagonal(G_ab)
    else:
        da, db = torch.diagonal(G_a), torch.diagonal(G_b)
    sq = da[:, None] + db[None, :] - 2 * G_ab
    return sq if power == 2 else torch.pow(sq, power / 2.)


def rkhs_cdist(a, b=None, power: float = 2.):
    if b is None:
        return rkhs_gram_cdist(a.inner(), power=power)
    return rkhs_gram_cdist(a.inner(b), a.inner(), b.inner(), power=power)


def dist(a, b=None, power: float = 2.):
    from ..rkhs.base import Vec
    if isinstance(a, Vec):
        return rkhs_cdist(a, b, power=power)
    return eucldist(a, b, power=power)

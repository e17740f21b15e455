"""`dist` -- drop-in for jaxrk/utilities/distances.py:44-49 (Array branch -> eucldist;
Vec branch -> RKHS distances from the reduced Gram matrices, distances.py:28-42 over
utilities/gram.py:8-23)."""
import torch

from .eucldist import eucldist
from .gram import rkhs_gram_cdist

__all__ = ["dist", "rkhs_cdist", "rkhs_gram_cdist"]


def rkhs_cdist(a, b=None, power: float = 2.):
    if b is None:
        return rkhs_gram_cdist(a.inner(), power=power)
    return rkhs_gram_cdist(a.inner(b), a.inner(), b.inner(), power=power)


def dist(a, b=None, power: float = 2.):
    from ..rkhs.base import Vec
    if isinstance(a, Vec):
        return rkhs_cdist(a, b, power=power)
    return eucldist(a, b, power=power)

"""Pairwise Euclidean distances -- drop-in for jaxrk/utilities/eucldist.py.

`eucldist(a, b=None, power=1., variant="simple")` keeps the reference signature
(eucldist.py:37-48) and returns the N x M matrix of ||a_i - b_j||^power.  The reference
evaluates the expansion ||a||^2 + ||b||^2 - 2ab' (eucldist.py:10-18, or the single-GEMM
"extension" form :20-35), clips at 0 and takes power/2; both variants here run the
direct-differencing sm_100a kernel (jrk_dist_f32), which is exact at zero distance and
never negative, so the clip is a no-op and `power == 2` returns a non-negative matrix.
"""
from .. import _native as nat
from .._array import as2d

__all__ = ["eucldist", "sqeucldist_simple", "sqeucldist_extension"]


def _pair(a, b):
    a = as2d(a)
    b = None if b is None else as2d(b, a.device)
    if b is not None:
        assert a.shape[1] == b.shape[1]
    return a, b


def sqeucldist_simple(a, b=None):
    a, b = _pair(a, b)
    return nat.dist(a, b, 1.0, 2.0)


def sqeucldist_extension(a, b=None):
    a, b = _pair(a, b)
    return nat.dist(a, b, 1.0, 2.0)


def eucldist(a, b=None, power=1., variant="simple"):
    assert variant in ("simple", "extension")
    a, b = _pair(a, b)
    return nat.dist(a, b, 1.0, float(power))

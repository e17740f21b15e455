"""ctypes wrapper of oracle/liboracle.so (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY --
see the header of oracle/jaxrk_oracle.py for who may import this."""
import ctypes
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB


class COracle:
    def __init__(self, lib):
        self.lib = lib
        f = lib.orc_gram_ref32
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float, c_void_p]
        f = lib.orc_gram_truth64
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float, c_void_p]
        f = lib.orc_kmatw_truth64
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int,
                                       c_float, c_float, c_float, c_void_p, c_void_p]
        f = lib.orc_kmatw_ref32
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int,
                                       c_float, c_float, c_float, c_void_p]
        f = lib.orc_blocksum_truth64
        f.restype, f.argtypes = c_double, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                           c_int, c_float, c_float, c_float]

        f = lib.orc_periodic_truth64
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_float, c_void_p]
        f = lib.orc_threshspike_truth64
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_float, c_float, c_float,
                                       c_float, c_void_p, c_void_p]
        f = lib.orc_dist_diag_truth64
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_void_p]
        f = lib.orc_vjp_truth64
        f.restype, f.argtypes = None, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_float, c_float,
                                       c_void_p, c_void_p]

    @staticmethod
    def _p(a):
        return None if a is None else c_void_p(a.ctypes.data)

    @staticmethod
    def _f(a):
        return np.ascontiguousarray(a, dtype=np.float32)

    def gram_ref32(self, X, Y, s, p, nconst, i0=0, i1=None):
        X, Y = self._f(X), self._f(Y)
        i1 = X.shape[0] if i1 is None else i1
        out = np.empty((i1 - i0, Y.shape[0]), np.float32)
        self.lib.orc_gram_ref32(self._p(X), self._p(Y), i0, i1, Y.shape[0], X.shape[1], s, p, nconst, self._p(out))
        return out

    def gram_truth64(self, X, Y, s, p, nconst, i0=0, i1=None):
        X, Y = self._f(X), self._f(Y)
        i1 = X.shape[0] if i1 is None else i1
        out = np.empty((i1 - i0, Y.shape[0]), np.float64)
        self.lib.orc_gram_truth64(self._p(X), self._p(Y), i0, i1, Y.shape[0], X.shape[1], s, p, nconst, self._p(out))
        return out

    def kmatw_truth64(self, X, Y, W, s, p, nconst, i0=0, i1=None, with_abs=False):
        X, Y, W = self._f(X), self._f(Y), self._f(W)
        i1 = X.shape[0] if i1 is None else i1
        out = np.empty((i1 - i0, W.shape[1]), np.float64)
        ab = np.empty_like(out) if with_abs else None
        self.lib.orc_kmatw_truth64(self._p(X), self._p(Y), self._p(W), i0, i1, Y.shape[0], X.shape[1], W.shape[1],
                                   s, p, nconst, self._p(out), self._p(ab))
        return (out, ab) if with_abs else out

    def kmatw_ref32(self, X, Y, W, s, p, nconst, i0=0, i1=None):
        X, Y, W = self._f(X), self._f(Y), self._f(W)
        i1 = X.shape[0] if i1 is None else i1
        out = np.empty((i1 - i0, W.shape[1]), np.float32)
        self.lib.orc_kmatw_ref32(self._p(X), self._p(Y), self._p(W), i0, i1, Y.shape[0], X.shape[1], W.shape[1],
                                 s, p, nconst, self._p(out))
        return out

    def blocksum_truth64(self, X, Y, pi, pj, i0, i1, j0, j1, s, p, nconst):
        X, Y = self._f(X), self._f(Y)
        pi = None if pi is None else self._f(pi)
        pj = None if pj is None else self._f(pj)
        return float(self.lib.orc_blocksum_truth64(self._p(X), self._p(Y), self._p(pi), self._p(pj), i0, i1, j0, j1,
                                                   X.shape[1], s, p, nconst))


    def periodic_truth64(self, X, Y, s, inv_ls):
        X, Y = self._f(X), self._f(Y)
        out = np.empty((X.shape[0], Y.shape[0]), np.float64)
        self.lib.orc_periodic_truth64(self._p(X), self._p(Y), X.shape[0], Y.shape[0], X.shape[1], s, inv_ls, self._p(out))
        return out

    def threshspike_truth64(self, X, Y, s, p, spike, non_spike, threshold):
        X, Y = self._f(X), self._f(Y)
        out = np.empty((X.shape[0], Y.shape[0]), np.float64)
        margin = np.empty_like(out)
        self.lib.orc_threshspike_truth64(self._p(X), self._p(Y), X.shape[0], Y.shape[0], X.shape[1], s, p, spike, non_spike,
                                         threshold, self._p(out), self._p(margin))
        return out, margin

    def dist_diag_truth64(self, X, Y, s, p):
        X, Y = self._f(X), self._f(Y)
        out = np.empty((X.shape[0],), np.float64)
        self.lib.orc_dist_diag_truth64(self._p(X), self._p(Y), X.shape[0], X.shape[1], s, p, self._p(out))
        return out

    def vjp_truth64(self, X, Y, A, s, p, nconst):
        X, Y = self._f(X), self._f(Y)
        A = np.ascontiguousarray(A, dtype=np.float64)
        gX = np.empty((X.shape[0], X.shape[1]), np.float64)
        gs = np.zeros((1,), np.float64)
        self.lib.orc_vjp_truth64(self._p(X), self._p(Y), self._p(A), X.shape[0], Y.shape[0], X.shape[1], s, p, nconst,
                                 self._p(gX), self._p(gs))
        return gX, float(gs[0])


def load() -> COracle:
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
        build()
    return COracle(ctypes.CDLL(LIB))

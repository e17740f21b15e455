"""Gaussian-process regression on top of FiniteVec.inner -- drop-in for jaxrk/gp.py:6-33.

A consumer of the hot path: the train / cross / test covariances are inner products of FiniteVecs, i.e. fused Gram
kernels (`x.inner()`, `x.inner(xtest)`, `xtest.inner()`); the Cholesky factorisation and the triangular solves stay on
the dense library path (torch.linalg), exactly as the Tikhonov solve of the conditional operators does.
"""
import math

import torch

from ._array import asarray
from .rkhs import FiniteVec

__all__ = ["GP"]


class GP(object):
    def __init__(self, x: FiniteVec, y, noise: float, amp: float, kconst: float = 0.):
        self.x, self.noise, self.amp, self.kconst = x, noise, amp, kconst
        y = asarray(y)
        self.ymean = y.mean()
        self.y = y - self.ymean                      # (the reference centres y in place, gp.py:9-10)
        n = len(self.x)
        train_cov = self.amp * (self.x.inner() + kconst) + torch.eye(n, dtype=torch.float32, device=y.device) * self.noise
        self.chol = torch.linalg.cholesky(train_cov)
        self.kinvy = torch.cholesky_solve(self.y.reshape(n, -1), self.chol).reshape(self.y.shape)

    def marginal_likelihood(self):
        log2pi = math.log(2. * math.pi)
        ml = torch.sum(-0.5 * (self.y.T @ self.kinvy) - torch.sum(torch.log(torch.diagonal(self.chol)))
                       - (len(self.x) / 2.) * log2pi)
        ml = ml - (-0.5 * log2pi - math.log(self.amp) ** 2)      # lognormal prior (gp.py:20)
        return -ml

    def predict(self, xtest: FiniteVec = None):
        cross_cov = self.amp * (self.x.inner(xtest) + self.kconst)
        mu = cross_cov.T @ self.kinvy + self.ymean
        v = torch.linalg.solve_triangular(self.chol, cross_cov, upper=False)
        var = self.amp * (xtest.inner() + self.kconst) - v.T @ v
        return mu, var

    def post_pred_likelihood(self, xtest: FiniteVec, ytest):
        m, v = self.predict(xtest)
        dist = torch.distributions.MultivariateNormal(m.reshape(-1).double(), covariance_matrix=v.double())
        return m, v, dist.log_prob(asarray(ytest).reshape(-1).double())

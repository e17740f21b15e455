"""Kernel ABCs -- the operator API the hot path sits behind (jaxrk/kern/base.py:23-50)."""
from abc import ABC, abstractmethod


class Kernel(ABC):
    """A positive definite kernel.  `k(X, Y=None, diag=False)` returns the Gram matrix
    (N x M; Y=None means Y=X) or, with diag=True, only its diagonal."""

    @abstractmethod
    def __call__(self, X, Y=None, diag: bool = False):
        raise NotImplementedError()


class DensityKernel(Kernel):
    """Kernels that are also probability densities (jaxrk/kern/base.py:39-50)."""

    def std(self):
        return self.var() ** 0.5

    @abstractmethod
    def var(self):
        raise NotImplementedError()

    def rvs(self, nsamps):
        raise NotImplementedError()

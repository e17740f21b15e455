"""GenGaussKernel -- drop-in for jaxrk/kern/rbf.py:27-105.

k(x, y) = nconst * exp(-(s ||x - y||)^p), p in (0, 2]: the density of scipy's `gennorm`
(p = 2 squared-exponential via make_gauss, p = 1 Laplace via make_laplace).  The Gram
matrix comes from ONE fused sm_100a kernel (jrk_gram_f32) instead of the reference's
GEMM + ~9 eager N x M passes.
"""
import math

import numpy as np
import torch

from .. import _native as nat
from .._array import as2d
from .base import DensityKernel, Kernel
from .util import ScaledPairwiseDistance, SimpleScaler

f32 = np.float32


def _gammaln32(x):
    return np.vectorize(math.lgamma)(np.asarray(x, dtype=f32)).astype(f32)


class GenGaussKernel(DensityKernel):
    def __init__(self, dist: ScaledPairwiseDistance, scale_tensor=None, power_tensor=None, dim_scale_tensor=None):
        self.dist = dist
        # differentiable global scale (0-dim tensor, 1 / length_scale) when the kernel was made from a tensor that
        # requires grad: the Gram / inner products are then differentiable in it (jaxrk_b200/autograd.py)
        self.scale_tensor = scale_tensor
        # likewise the shape parameter (0-dim tensor) and per-dimension scales (len d tensor: a pre-scaling of the points,
        # jaxrk/kern/util.py:107-116, differentiated through `X * dim_scale` by torch)
        self.power_tensor = power_tensor
        self.dim_scale_tensor = dim_scale_tensor
        # rbf.py:34, float32 as under jax without x64
        self.nconst = (f32(dist.power) / (f32(2) * dist._get_scale_param()
                                          * np.exp(_gammaln32(1. / dist.power)))).astype(f32)

    @classmethod
    def make(cls, length_scale, shape: float) -> "GenGaussKernel":
        """rbf.py:62-73: scale = 1/length_scale, shape in (0, 2]."""
        scale_t = power_t = dim_t = None
        if isinstance(shape, torch.Tensor):
            if shape.requires_grad:
                power_t = shape.reshape(())
            shape = float(shape.detach())
        assert shape > 0 and shape <= 2
        if isinstance(length_scale, torch.Tensor):
            if length_scale.requires_grad and length_scale.numel() == 1:
                scale_t = 1. / length_scale.reshape(())
            elif length_scale.requires_grad:
                dim_t = 1. / length_scale.reshape(-1)
            length_scale = length_scale.detach().cpu().numpy()
        if isinstance(length_scale, (float, int)):
            s = 1. / float(length_scale)
        else:
            ls = np.asarray(length_scale)
            s = float(1. / ls) if ls.ndim == 0 else (f32(1.) / ls.astype(f32)).astype(f32)
        return GenGaussKernel(ScaledPairwiseDistance(scaler=SimpleScaler(s), power=shape), scale_tensor=scale_t,
                              power_tensor=power_t, dim_scale_tensor=dim_t)

    @classmethod
    def make_laplace(cls, length_scale) -> "GenGaussKernel":
        return GenGaussKernel.make(length_scale, 1.)

    @classmethod
    def make_gauss(cls, length_scale) -> "GenGaussKernel":
        # the reference hard-codes this literal (rbf.py:93); it is not exactly 1/sqrt(2)
        f = 0.70710695
        if isinstance(length_scale, torch.Tensor) and length_scale.requires_grad:
            return GenGaussKernel.make(length_scale / f, 2.)
        if isinstance(length_scale, (float, int)):
            return GenGaussKernel.make(float(length_scale) / f, 2.)
        ls = np.asarray(length_scale.detach().cpu().numpy() if isinstance(length_scale, torch.Tensor) else length_scale)
        if ls.ndim == 0:
            return GenGaussKernel.make(float(f32(ls) / f32(f)), 2.)
        return GenGaussKernel.make((ls.astype(f32) / f32(f)).astype(f32), 2.)

    def std(self):
        return np.sqrt(self.var())

    def var(self):
        f = _gammaln32(np.array([3, 1], dtype=f32) / f32(self.dist.power))
        return (self.dist._get_scale_param() ** 2 * np.exp(f[0] - f[1])).astype(f32)

    def kernel_args(self):
        """(scale, dim_scale, power, nconst) for the C ABI, or None when the kernel is not
        representable there (per-dimension scale: nconst is a (1, d) array, quirk Q2)."""
        if self.nconst.size != 1:
            return None
        scale, dim_scale = self.dist.kernel_args()
        return scale, dim_scale, float(self.dist.power), float(self.nconst.reshape(-1)[0])

    def __call__(self, X, Y=None, diag=False):
        X = as2d(X)
        Y = None if Y is None else as2d(Y, X.device)
        args = self.kernel_args()
        if not diag and self.dim_scale_tensor is not None:
            # a kernel made from a TENSOR of per-dimension length scales: one definition with and without autograd
            return self._call_dim_scale_grad(X, Y)
        if diag or args is None:
            if self.wants_grad(X, Y):
                from ..rkhs._lowering import grad_unsupported
                raise grad_unsupported("GenGaussKernel(diag=True / per-dimension nconst)")
            nc = torch.from_numpy(self.nconst).to(X.device)
            return nc * torch.exp(-self.dist(X, Y, diag))  # rbf.py:104-105 literally (O(N d) / quirk Q2)
        scale, dim_scale, power, nconst = args
        if self.wants_grad(X, Y) and (dim_scale is not None or X.shape[1] > 64):
            from ..rkhs._lowering import grad_unsupported
            raise grad_unsupported(f"GenGaussKernel Gram with d = {X.shape[1]}" + (" and per-dimension scales" if dim_scale is not None else ""))
        if self.wants_grad(X, Y) and dim_scale is None:
            from .. import autograd as ag
            pw = self.power_tensor if self.power_tensor is not None else power
            if self.scale_tensor is not None:
                return ag.gengauss_gram(X, Y, self.scale_tensor, pw)             # nconst follows the scale (and the power)
            if self.power_tensor is not None:
                return ag.gengauss_gram(X, Y, scale, pw)
            return ag.gengauss_gram(X, Y, scale, power, nconst=nconst)
        return nat.gram(X, Y, scale, power, nconst, dim_scale=dim_scale)

    def grad_inputs(self, X, Y):
        """(X', Y', scale argument, power argument, nconst argument) for the differentiable entry points of
        jaxrk_b200/autograd.py: per-dimension scales become a pre-scaling of the points (torch differentiates the
        multiplication).  The reference's normalising constant for a per-dimension scale is a (1, d) array (quirk Q2:
        rbf.py:34) that only broadcasts against an N x d Gram; a kernel made from a tensor of length scales is defined here
        as  exp(-||(x - y) / length_scale||^p)  times the constant of a unit global scale -- with and without autograd."""
        pw = self.power_tensor if self.power_tensor is not None else float(self.dist.power)
        if self.dim_scale_tensor is not None:
            ds = self.dim_scale_tensor.to(X.device)
            return X * ds, (None if Y is None else Y * ds), 1.0, pw, None
        if self.scale_tensor is not None:
            return X, Y, self.scale_tensor, pw, None
        scale, _, _, nconst = self.kernel_args()
        return X, Y, scale, pw, (None if self.power_tensor is not None else nconst)

    def _call_dim_scale_grad(self, X, Y):
        from .. import autograd as ag
        Xs, Ys, sc, pw, nc = self.grad_inputs(X, Y)
        return ag.gengauss_gram(Xs, Ys, sc, pw, nconst=nc)

    def wants_grad(self, *tensors) -> bool:
        """True when autograd is recording and one of the tensors -- or one of this kernel's parameters -- requires grad."""
        if not torch.is_grad_enabled():
            return False
        for t in (self.scale_tensor, self.power_tensor, self.dim_scale_tensor):
            if t is not None and t.requires_grad:
                return True
        return any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class PeriodicKernel(Kernel):
    """exp(-2 (sin(pi ||x - y|| / period) / length_scale)^2) -- drop-in for jaxrk/kern/rbf.py:107-154.  The Gram comes
    from the direct-differencing kernel with the periodic epilogue (jrk_gram_kind_f32, JRK_KIND_PERIODIC); inner
    products with reduce chains go through jrk_kmatw_kind_f32."""

    def __init__(self, period, length_scale: float):
        # `period` may be an Array (rbf.py:109): one entry per dimension goes through SimpleScaler as a per-dimension
        # scale (the C ABI's dim_scale), a single entry is the global scale
        per = np.asarray(period.detach().cpu().numpy() if isinstance(period, torch.Tensor) else period, dtype=np.float64)
        length_scale = float(length_scale)
        assert (per > 0).all() and length_scale > 0
        s = 1. / float(per.reshape(-1)[0]) if per.size == 1 else (1. / per).astype(f32)
        self.dist = ScaledPairwiseDistance(scaler=SimpleScaler(s), power=1.)
        self.ls = length_scale

    def kernel_spec(self):
        scale, dim_scale = self.dist.kernel_args()
        return nat.KernelSpec(nat.KIND_PERIODIC, (scale, self.ls), dim_scale)

    def __call__(self, X, Y=None, diag=False):
        X = as2d(X)
        Y = None if Y is None else as2d(Y, X.device)
        if diag:  # O(N d): rbf.py:152-154 literally over the diag distance formula
            d = self.dist(X, Y, True)
            return torch.exp(-2. * (torch.sin(math.pi * d) / self.ls) ** 2.)
        return nat.gram_spec(X, Y, self.kernel_spec())


class ThreshSpikeKernel(Kernel):
    """`spike` where the scaled distance is <= threshold, else `non_spike` -- drop-in for jaxrk/kern/rbf.py:156-206
    (jrk_gram_kind_f32, JRK_KIND_THRESHSPIKE: the comparison runs on the squared distance against the equivalent
    threshold, one compare + select per entry)."""

    def __init__(self, dist: ScaledPairwiseDistance, spike: float, non_spike: float, threshold: float):
        spike, non_spike, threshold = float(spike), float(non_spike), float(threshold)
        assert spike > 0
        assert abs(non_spike) < spike
        assert threshold >= 0
        self.dist = dist
        self.spike, self.non_spike, self.threshold = spike, non_spike, threshold

    def kernel_spec(self):
        scale, dim_scale = self.dist.kernel_args()
        return nat.KernelSpec(nat.KIND_THRESHSPIKE, (scale, float(self.dist.power), self.spike, self.non_spike,
                                                     self.threshold), dim_scale)

    def __call__(self, X, Y=None, diag=False):
        X = as2d(X)
        assert X.dim() == 2
        assert not diag
        Y = None if Y is None else as2d(Y, X.device)
        return nat.gram_spec(X, Y, self.kernel_spec())

"""Differentiable Gram / fused K@W: torch.autograd over the fused backward kernels (csrc/kgrad.cu).

JaxRK is written to be differentiated with jax.grad (README.md:3; examples/Using_FLAX_Optax.ipynb cells 6-10 fit kernel
parameters and prefactors by gradient descent; jaxrk/rkhs/vector.py:206-222 and jaxrk/utilities/gram.py:62-65 run
L-BFGS over jax.grad).  With torch tensors as the array type of this package, the same role is played by
torch.autograd: `gengauss_gram` and `gengauss_kmatw` are the differentiable forms of GenGaussKernel.__call__
(jaxrk/kern/rbf.py:104-105) and of FiniteVec.inner's K @ W contraction (jaxrk/rkhs/vector.py:62-75), with gradients
with respect to the points X, Y, the weights W, the global scale (1 / length_scale) and the shape parameter `power`
computed by ONE fused pass each (jrk_kmatw_grad_f32 / jrk_gram_grad_f32) that never materialises K or dK.

`power` may be a Python float (not differentiated) or a 0-dim tensor; the normalising constant follows both parameters
differentiably (torch.lgamma).  Per-dimension scales are a pre-scaling of the points (jaxrk/kern/util.py:107-116 applies
the scaler to X and Y before the distance): `X * dim_scale` in torch, so their gradient is the chain rule over gX / gY.
"""
import math

import torch

from . import _native as nat

__all__ = ["gengauss_gram", "gengauss_kmatw", "gengauss_nconst"]


def gengauss_nconst(scale, power):
    """nconst = power / (2 * (1 / scale) * Gamma(1 / power))  (jaxrk/kern/rbf.py:34), differentiable in `scale` and -- when
    it is a tensor -- in `power`."""
    if isinstance(power, torch.Tensor):
        return scale * power / (2.0 * torch.exp(torch.lgamma(1.0 / power)))
    return scale * (power / (2.0 * math.exp(math.lgamma(1.0 / power))))


def _power_tensor(power, dev):
    """(float value, 0-dim tensor or None): the tensor only when `power` is one that takes part in autograd."""
    if isinstance(power, torch.Tensor):
        pt = power.to(device=dev, dtype=torch.float32).reshape(())
        return float(pt.detach()), (pt if pt.requires_grad else None)
    return float(power), None


def _scale_tensor(scale, dev):
    if isinstance(scale, torch.Tensor):
        return scale.to(device=dev, dtype=torch.float32)
    return torch.tensor(float(scale), dtype=torch.float32, device=dev)


class _UnnormGram(torch.autograd.Function):
    """E[i, j] = exp(-(s ||x_i - y_j||)^p)."""

    @staticmethod
    def forward(ctx, X, Y, scale_t, power, self_gram, power_t):
        s = float(scale_t)
        E = nat.gram(X, None if self_gram else Y, s, power, 1.0)
        ctx.save_for_backward(X, Y, scale_t)
        ctx.power = power
        return E

    @staticmethod
    def backward(ctx, Gbar):
        X, Y, scale_t = ctx.saved_tensors
        s, p = float(scale_t), ctx.power
        need_x, need_y, need_s, need_p = (ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2],
                                          ctx.needs_input_grad[5])
        Gbar = Gbar.contiguous().float()
        gX = gY = gs = gp = None
        # the kernels read the cotangent fastest when the points they differentiate run along its contiguous axis: dL/dY
        # takes Gbar (N x M) as it is, dL/dX one transposed copy
        if need_x or need_s or need_p:
            res = nat.gram_grad(X, Y, Gbar.t().contiguous(), s, p, 1.0, want_scale=need_s, want_power=need_p, transposed=True)
            gX, gs_rows = res[0], res[1]
            if need_s:
                gs = gs_rows.sum().reshape(scale_t.shape)
            if need_p:
                gp = res[2].sum().reshape(())
            if not need_x:
                gX = None
        if need_y:
            gY, _ = nat.gram_grad(Y, X, Gbar, s, p, 1.0, want_scale=False, transposed=True)
        return gX, gY, gs, None, None, gp


class _UnnormKmatw(torch.autograd.Function):
    """T[i, c] = sum_j exp(-(s ||x_i - y_j||)^p) W[j, c], never materialised (forward or backward)."""

    @staticmethod
    def forward(ctx, X, Y, W, scale_t, power, power_t):
        s = float(scale_t)
        T = nat.kmatw(X, Y, W, s, power, 1.0)
        ctx.save_for_backward(X, Y, W, scale_t)
        ctx.power = power
        return T

    @staticmethod
    def backward(ctx, G):
        X, Y, W, scale_t = ctx.saved_tensors
        s, p = float(scale_t), ctx.power
        need_x, need_y, need_w, need_s = ctx.needs_input_grad[:4]
        need_p = ctx.needs_input_grad[5]
        G = G.contiguous().float()
        m = W.shape[1]
        gX = gY = gW = gs = gp = None
        # the gradient adds up over column blocks: 32 columns for the FP32-pipe kernel, 16 where the tensor-core kernel
        # (csrc/kgrad_tc.cu: p = 2, d <= 16, large problems) serves them -- two of its passes beat one FP32-pipe pass
        blk = 16 if (p == 2.0 and X.shape[1] <= 16 and min(X.shape[0], Y.shape[0]) >= 4096) else 32
        for c0 in range(0, m, blk):
            Wb, Gb = W[:, c0:c0 + blk], G[:, c0:c0 + blk]
            if need_x or need_s or need_p:
                res = nat.kmatw_grad(X, Y, Wb, Gb, s, p, 1.0, want_scale=need_s, want_power=need_p)
                gx_b, gs_rows = res[0], res[1]
                if need_x:
                    gX = gx_b if gX is None else gX + gx_b
                if need_s:
                    gs_b = gs_rows.sum()
                    gs = gs_b if gs is None else gs + gs_b
                if need_p:
                    gp_b = res[2].sum()
                    gp = gp_b if gp is None else gp + gp_b
            if need_y:
                gy_b, _ = nat.kmatw_grad(Y, X, Gb, Wb, s, p, 1.0, want_scale=False)
                gY = gy_b if gY is None else gY + gy_b
        if need_w:
            gW = nat.kmatw(Y, X, G, s, p, 1.0)        # K(Y, X) @ G = K^T G
        if gs is not None:
            gs = gs.reshape(scale_t.shape)
        if gp is not None:
            gp = gp.reshape(())
        return gX, gY, gW, gs, None, gp


def gengauss_gram(X, Y, scale, power, nconst=None):
    """K = nconst * exp(-(scale ||x - y||)^power) with gradients to X, Y, `scale` and `power` (floats or 0-dim tensors).
    Y=None is the self-Gram (symmetric kernel, gradient through both arguments).  nconst=None derives it from
    `scale` and `power` as the reference does (rbf.py:34), differentiably."""
    dev = X.device
    scale_t = _scale_tensor(scale, dev)
    pv, pt = _power_tensor(power, dev)
    self_gram = Y is None
    E = _UnnormGram.apply(X, X if self_gram else Y, scale_t, pv, self_gram, pt)
    nc = gengauss_nconst(scale_t, pt if pt is not None else pv) if nconst is None else nconst
    return nc * E


def gengauss_kmatw(X, Y, W, scale, power, nconst=None):
    """K(X, Y) @ W, never materialised, with gradients to X, Y, W, `scale` and `power`."""
    dev = X.device
    scale_t = _scale_tensor(scale, dev)
    pv, pt = _power_tensor(power, dev)
    T = _UnnormKmatw.apply(X, Y, W, scale_t, pv, pt)
    nc = gengauss_nconst(scale_t, pt if pt is not None else pv) if nconst is None else nconst
    return nc * T

"""Lowering of  R_X . K(X, Y) . R_Y^T  onto the fused sm_100a kernels.

The reference's FiniteVec.inner (jaxrk/rkhs/vector.py:62-75) materialises the N x M Gram
and then applies both reduce chains with Reduce.apply (jaxrk/reduce/base.py:39-47).
Every chain element is a linear map on one axis, so a chain factors as
R = R_post . R_fold, where R_fold is something the kernels apply in registers:

    fold  := Prefactors*  then one of  { nothing, Sum, Mean, BalancedRed, LinearReduce(+ more) }

and R_post (whatever follows) acts on the already reduced -- small -- result with the
ordinary array-level Reduce.apply.  Once a LinearReduce(A) has been reached, every later
element is folded into A itself (it is applied to A's rows).

Kernel choice, with (rx, ry) the folded row / column maps:
    ry = LINEAR | SUM | MEAN          -> jrk_kmatw_f32(X, Y, W = R_y^T, row_reduce = rx)
    rx = LINEAR | SUM | MEAN, ry else -> the same on the transposed problem (k is symmetric:
                                          K(X, Y)^T = K(Y, X)), result transposed back
    rx = ry = BALANCED                -> jrk_gram_groupsum_f32
    rx = ry = NONE                    -> jrk_gram_f32 (materialised), prefactors scaled in place
    anything else                     -> Gram by jrk_gram_f32, then Reduce.apply (generic)
The sibling kernels of the same distance front-end (PeriodicKernel, ThreshSpikeKernel) take the same routes
through jrk_kmatw_kind_f32 / jrk_gram_kind_f32 (no group-sum kernel: Gram, then the array-level BalancedRed).
Nothing here computes a kernel value on the host or in torch: all k(x, y) evaluations are
done by libjaxrk_b200.so.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .. import _native as nat
from ..reduce.base import BalancedRed, Center, Mean, Prefactors, Reduce, Sum
from ..reduce.lincomb import BlockReduce, LinearReduce, SparseReduce

NONE, SUM, MEAN, BALANCED, LINEAR = "none", "sum", "mean", "balanced", "linear"
SPARSE_DENSE_LIMIT = 1 << 26      # elements of the dense form of a SparseReduce that may be built (256 MB)
SPARSE_FOLD_ROWS = 64             # ... and only for few output rows: every 16 (tensor cores) / 32 (FP32 pipe) of them cost
                                  # one more pass over K, beyond ~64 rows one materialised Gram + gather is cheaper


@dataclass
class Folded:
    n: int                                  # input length of the chain
    pref: Optional[torch.Tensor] = None     # product of the leading Prefactors (len n)
    kind: str = NONE
    group: int = 0                          # BALANCED: points per split
    average: bool = False                   # BALANCED: mean instead of sum
    A: Optional[torch.Tensor] = None        # LINEAR: (m x n) map, later elements already applied
    post: List[Reduce] = field(default_factory=list)

    @property
    def out_len(self) -> int:
        if self.kind in (SUM, MEAN):
            return 1
        if self.kind == BALANCED:
            return self.n // self.group
        if self.kind == LINEAR:
            return int(self.A.shape[0])
        return self.n


def _center_then(r: Reduce, n: int, dev) -> Optional[torch.Tensor]:
    """The dense map of  r . Center  when it is small:  r (I - 1 1^T / n) = r - (r 1) 1^T / n, or None."""
    if type(r) in (Sum, Mean):
        return torch.zeros((1, n), dtype=torch.float32, device=dev)                 # 1^T (I - 1 1^T / n) = 0
    if isinstance(r, LinearReduce):
        A = r.linear_map.to(dev)
    elif type(r) in (SparseReduce, BlockReduce) and r.num_rows() <= SPARSE_FOLD_ROWS:
        A = r.to_linear_map(n, dev)
    elif type(r) is BalancedRed and n // r.points_per_split <= SPARSE_FOLD_ROWS:
        A = r.reduce_first_ax(torch.eye(n, dtype=torch.float32, device=dev)) if n <= 4096 else None
        if A is None:
            g = n // r.points_per_split
            A = torch.zeros((g, n), dtype=torch.float32, device=dev)
            A.view(g, g, r.points_per_split)[torch.arange(g), torch.arange(g)] = (1.0 / r.points_per_split) if r.average else 1.0
    else:
        return None
    assert A.shape[1] == n
    return A - A.sum(1, keepdim=True) / n


def fold_chain(chain: Optional[List[Reduce]], n: int, dev) -> Folded:
    f = Folded(n=n)
    chain = list(chain or [])
    for idx, r in enumerate(chain):
        if f.kind == LINEAR:
            # every Reduce is linear on the first axis: apply it to A's rows (m x n -> m' x n)
            f.A = r.reduce_first_ax(f.A)
            continue
        if f.kind != NONE:
            f.post = chain[idx:]
            break
        if type(r) is Prefactors:
            assert r.prefactors.shape[0] == n
            p = r.prefactors.to(dev)
            f.pref = p if f.pref is None else f.pref * p
        elif type(r) is Sum:
            f.kind = SUM
        elif type(r) is Mean:
            f.kind = MEAN
        elif type(r) is BalancedRed:
            assert n % r.points_per_split == 0
            f.kind, f.group, f.average = BALANCED, r.points_per_split, r.average
        elif isinstance(r, LinearReduce):      # incl. KernelMapReduce (materialises its map here)
            assert r.linear_map.shape[1] == n
            f.kind, f.A = LINEAR, r.linear_map.to(dev)
        elif type(r) in (SparseReduce, BlockReduce) and r.num_rows() <= SPARSE_FOLD_ROWS and r.num_rows() * n <= SPARSE_DENSE_LIMIT:
            # a sparse / block averaging map small enough to hold densely: folded like a LinearReduce (fused K@W, the Gram
            # is not materialised); larger ones fall through to the generic route below
            f.kind, f.A = LINEAR, r.to_linear_map(n, dev)
        elif type(r) is Center and idx + 1 < len(chain) and (A := _center_then(chain[idx + 1], n, dev)) is not None:
            # Center followed by a small linear map: (r . Center) is itself a dense map -- the centring is folded into the
            # fused K@W (jaxrk/reduce/base.py:183-192 composed with the next element); the element after it is consumed
            f.kind, f.A = LINEAR, A
            for r2 in chain[idx + 2:]:
                f.A = r2.reduce_first_ax(f.A)
            break
        else:  # Center, TileView, SparseReduce, user types: not folded
            f.post = chain[idx:]
            break
    return f


def _w_of(f: Folded, dev) -> torch.Tensor:
    """The (n x m) matrix W = R^T of a LINEAR / SUM / MEAN fold (prefactors stay separate)."""
    if f.kind == LINEAR:
        return f.A.t().contiguous()
    w = torch.ones((f.n, 1), dtype=torch.float32, device=dev)
    return w / f.n if f.kind == MEAN else w


def _row_reduce_of(f: Folded):
    if f.kind == SUM:
        return nat.REDUCE_SUM, 0
    if f.kind == MEAN:
        return nat.REDUCE_MEAN, 0
    if f.kind == BALANCED:
        return (nat.REDUCE_BALANCED_MEAN if f.average else nat.REDUCE_BALANCED), f.group
    return nat.REDUCE_NONE, 0


WIDE_M = 64                       # columns of W from which "Gram slab + library GEMM" beats the FP32-pipe kernel's column blocks
WIDE_SLAB_BYTES = 2 << 30         # FP32 bytes of one materialised row slab of K
GROUPSUM_TENSOR = True            # BalancedRed x BalancedRed inside the tensor-core window: group sums by the fused K@W kernel


def _kmatw_wide(X, Y, W, spec, row_pref, col_pref) -> torch.Tensor:
    """diag(row_pref) K(X, Y) diag(col_pref) W for WIDE W outside the tensor-core window (d > 32, per-dimension scales,
    sibling kernels, small problems) -- FiniteOp.__matmul__ with many output features (jaxrk/rkhs/operator.py:46-52).
    K is evaluated ONCE, slab by slab, by the fused Gram kernel (HBM-bound); the dense contraction is the library's FP32
    GEMM (cuBLAS SGEMM, TF32 off: the same arithmetic class as the reference's XLA dot), as the solve's products are.
    [Measured and rejected: the contraction on the tensor cores as three fp16 hi / lo' GEMMs with FP32 output
    (torch.mm(..., out_dtype=float32)): the tensor core's accumulator truncates, an all-positive contraction of 4096
    terms comes out 5e-5 low -- five times the tolerance -- and chunks short enough to fix that (256) move more bytes
    than the GEMM has flops for.  Inside the tensor-core window the fused kernel (column blocks of 16, each block folded
    into FP32 registers every 256 terms) is both exact enough and faster than SGEMM: 109 vs ~65 TFLOP/s equivalent.]"""
    N, M, m = X.shape[0], Y.shape[0], W.shape[1]
    Ws = W if col_pref is None else col_pref[:, None] * W
    out = torch.empty((N, m), dtype=torch.float32, device=X.device)
    slab = max(128, min(N, (WIDE_SLAB_BYTES // (4 * M)) // 128 * 128))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for r0 in range(0, N, slab):
            r1 = min(N, r0 + slab)
            K = nat.gram_spec(X[r0:r1], Y, spec)                      # the only kernel evaluations of the slab
            if row_pref is not None:
                K.mul_(row_pref[r0:r1, None])
            torch.mm(K, Ws, out=out[r0:r1])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


def _kmatw_side(X, Y, fx: Folded, fy: Folded, spec) -> torch.Tensor:
    """fold_x(K(X, Y)) applied with W = fold_y^T; fy is LINEAR / SUM / MEAN."""
    W = _w_of(fy, X.device)
    tc = (spec.kind == nat.KIND_GENGAUSS and spec.dim_scale is None
          and nat.lib().jrk_kmatw_uses_tensor_cores(X.shape[0], Y.shape[0], X.shape[1], W.shape[1], 0, 0))
    if (W.shape[1] >= WIDE_M and not tc) or X.shape[1] > 128:
        # wide maps (operator application) outside the tensor-core window, and d > 128 (beyond the fused K@W kernels:
        # jrk_kmatw_f32 stops at d = 128, the Gram kernel does not): Gram slabs + library GEMM, the row fold on the
        # N x m result
        T = _kmatw_wide(X, Y, W, spec, fx.pref, fy.pref)
        if fx.kind == LINEAR:
            return fx.A @ T
        if fx.kind == SUM:
            return T.sum(0, keepdim=True)
        if fx.kind == MEAN:
            return T.mean(0, keepdim=True)
        if fx.kind == BALANCED:
            return BalancedRed(fx.group, fx.average)(T, 0)
        return T
    if fx.kind == LINEAR:
        T = nat.kmatw_spec(X, Y, W, spec, row_pref=fx.pref, col_pref=fy.pref)
        return fx.A @ T  # (mx x N) @ (N x my): small dense GEMM, stays in the library path
    rr, group = _row_reduce_of(fx)
    if (fx.kind == BALANCED and spec.kind == nat.KIND_GENGAUSS and spec.dim_scale is None
            and nat.lib().jrk_kmatw_uses_tensor_cores(X.shape[0], Y.shape[0], X.shape[1], W.shape[1], 0, 0)):
        # grouped row sums on a large, narrow problem: both contractions on the tensor cores (which write N x m),
        # the groups are summed on that small result
        T = nat.kmatw_spec(X, Y, W, spec, row_pref=fx.pref, col_pref=fy.pref)
        return BalancedRed(fx.group, fx.average)(T, 0)
    return nat.kmatw_spec(X, Y, W, spec, row_pref=fx.pref, col_pref=fy.pref, row_reduce=rr, group=group)


def kernel_spec_of(k):
    """KernelSpec of a kernel object: `kernel_spec()` (sibling kernels) or GenGauss `kernel_args()`; None when the
    C ABI cannot represent it."""
    if hasattr(k, "kernel_spec"):
        return k.kernel_spec()
    kargs = k.kernel_args() if hasattr(k, "kernel_args") else None
    if kargs is None:
        return None
    scale, dim_scale, power, nconst = kargs
    return nat.KernelSpec(nat.KIND_GENGAUSS, (scale, power, nconst), dim_scale)


def _chain_tensors(chain):
    for r in chain or []:
        if getattr(r, "is_materialised", True) is False:     # a factored KernelMapReduce: do not build its map here
            continue
        for name in ("prefactors", "linear_map"):
            t = getattr(r, name, None)
            if isinstance(t, torch.Tensor):
                yield t


class GradUnsupported(NotImplementedError):
    """Gradients were requested where no fused backward kernel exists.  Raised instead of returning a tensor without a
    graph (which would silently drop the gradient) -- and instead of an eager torch re-implementation of the kernel."""


def needs_grad(k, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    ks = tuple(getattr(k, name, None) for name in ("scale_tensor", "power_tensor", "dim_scale_tensor"))
    return any(isinstance(t, torch.Tensor) and t.requires_grad for t in (*ks, *tensors))


def grad_unsupported(what: str):
    return GradUnsupported(
        f"{what}: gradients are only available for GenGauss kernels with d <= 64 (the fused backward kernels "
        "jrk_kmatw_grad_f32 / jrk_gram_grad_f32; per-dimension length scales through GenGaussKernel.make(tensor)); detach "
        "the inputs or call under torch.no_grad() to get the value")


def _reduced_gram_autograd(k, spec, X, fx: Folded, Y, fy: Folded, self_gram: bool) -> torch.Tensor:
    """The differentiable route (torch.autograd over the fused backward kernels, jaxrk_b200/autograd.py): the K@W
    contraction stays fused; prefactors, the LinearReduce maps and the row reduces are ordinary torch ops on the small
    operands, so gradients reach the points, the chain parameters and the kernel's scale."""
    from .. import autograd as ag
    scale, power, nconst = spec.kparams
    st, pt = getattr(k, "scale_tensor", None), getattr(k, "power_tensor", None)
    s_arg, nc_arg = (st, None) if st is not None else (scale, None if pt is not None else nconst)
    if pt is not None:
        power = pt                                   # a 0-dim tensor: the backward kernels also return dL/dpower
    skinny = (LINEAR, SUM, MEAN)

    def side(Xa, fa: Folded, Yb, fb: Folded):      # fold_a(K(Xa, Yb)) applied with W = fold_b^T
        W = _w_of(fb, Xa.device)
        if fb.pref is not None:
            W = fb.pref[:, None] * W
        T = ag.gengauss_kmatw(Xa, Yb, W, s_arg, power, nconst=nc_arg)
        if fa.pref is not None:
            T = fa.pref[:, None] * T
        if fa.kind == LINEAR:
            return fa.A @ T
        if fa.kind == SUM:
            return T.sum(0, keepdim=True)
        if fa.kind == MEAN:
            return T.mean(0, keepdim=True)
        if fa.kind == BALANCED:
            return BalancedRed(fa.group, fa.average)(T, 0)
        return T

    if fy.kind in skinny and (fx.kind not in skinny or fy.out_len <= fx.out_len):
        return side(X, fx, Y, fy)
    if fx.kind in skinny:
        return side(Y, fy, X, fx).t()
    G = ag.gengauss_gram(X, None if self_gram else Y, s_arg, power, nconst=nc_arg)
    if fx.pref is not None:
        G = fx.pref[:, None] * G
    if fy.pref is not None:
        G = G * fy.pref[None, :]
    if fx.kind == BALANCED:
        G = BalancedRed(fx.group, fx.average)(G, 0)
    if fy.kind == BALANCED:
        G = BalancedRed(fy.group, fy.average)(G, 1)
    return G


def _groupsum_tensor(X, Y, fx: Folded, fy: Folded, scale, power, nconst, self_gram: bool) -> torch.Tensor:
    """BalancedRed on both axes through the tensor-core K@W kernel: for 16 column groups at a time,
    T = K(X, Y[groups]) @ (group indicator, M_chunk x 16) -- both contractions on tcgen05 (csrc/kmatw_tc.cu) -- and the
    row groups are summed on the small N x 16 result.  For the self inner product only the blocks on and below the
    diagonal are computed and mirrored, so the result is exactly symmetric (jaxrk/utilities/gram.py:18 asserts it).
    Measured at BASELINE configs[3] (4.19M points, d = 32, 1024 x 1024 groups): see DESIGN.md."""
    dev = X.device
    gx, gy = fx.group, fy.group
    na, nb = X.shape[0] // gx, Y.shape[0] // gy
    out = torch.empty((na, nb), dtype=torch.float32, device=dev)
    GB = 16
    ind = torch.zeros((GB * gy, GB), dtype=torch.float32, device=dev)
    for q in range(GB):
        ind[q * gy:(q + 1) * gy, q] = 1.0
    sym = self_gram and gx == gy and fx.pref is fy.pref
    for b0 in range(0, nb, GB):
        nbk = min(GB, nb - b0)
        a0 = b0 if sym else 0                         # symmetric: rows from the diagonal block downwards
        Xs = X[a0 * gx:]
        T = nat.kmatw(Xs, Y[b0 * gy:(b0 + nbk) * gy], ind[:nbk * gy, :nbk], scale, power, nconst,
                      row_pref=None if fx.pref is None else fx.pref[a0 * gx:],
                      col_pref=None if fy.pref is None else fy.pref[b0 * gy:(b0 + nbk) * gy])
        blk = T.reshape(na - a0, gx, nbk).sum(1)      # (row groups) x nbk, fixed summation order
        out[a0:, b0:b0 + nbk] = blk
        if sym:
            d_ = torch.tril(blk[:nbk])                # diagonal block: keep the lower triangle, mirror it
            out[b0:b0 + nbk, b0:b0 + nbk] = d_ + torch.tril(d_, -1).T
            out[b0:b0 + nbk, b0 + nbk:] = blk[nbk:].T
    return out


DIAG_GROUP_LOOP = 4096           # BalancedRed self inner products: up to this many groups get one fused call each


def reduced_gram_diag(k, X, chain) -> Optional[torch.Tensor]:
    """diag(R K(X, X) R^T) WITHOUT the full reduced Gram, or None when the chain does not allow it -- what the RKHS
    distances between two vectors need from each of them (jaxrk/utilities/distances.py:41-42 takes the diagonals of
    a.inner() and b.inner(), jaxrk/utilities/gram.py:30-35).  BalancedRed (mean embeddings): entry a is the sum over the
    a-th diagonal BLOCK only, one fused Sum-reduced K@W per group (sum_g g^2 kernel values instead of N^2);
    LinearReduce: one fused K(X, X) A^T and a row-wise dot product; no reduce: k(x_i, x_i)."""
    spec = kernel_spec_of(k)
    if spec is None or needs_grad(k, X, *_chain_tensors(chain)):
        return None
    dev = X.device
    f = fold_chain(chain, X.shape[0], dev)
    n = X.shape[0]
    if f.post:
        return None
    if f.kind == NONE:
        if spec.kind != nat.KIND_GENGAUSS:                     # (ThreshSpikeKernel has no diag form: the full route)
            return None
        d0 = torch.full((n,), float(spec.kparams[2]), dtype=torch.float32, device=dev)     # k(x, x) = nconst
        return d0 if f.pref is None else d0 * f.pref * f.pref
    if f.kind in (SUM, MEAN):
        return None                                            # a 1 x 1 result: the full route is the same work
    if f.kind == LINEAR:
        T = _kmatw_side(X, X, Folded(n=n, pref=f.pref), Folded(n=n, pref=f.pref, kind=LINEAR, A=f.A), spec)   # D_p K D_p A^T
        return (f.A.t() * T).sum(0)
    if f.kind == BALANCED and n // f.group <= DIAG_GROUP_LOOP:
        g = f.group
        out = torch.empty((n // g,), dtype=torch.float32, device=dev)
        ones = torch.ones((g, 1), dtype=torch.float32, device=dev)
        for a in range(n // g):
            Xg = X[a * g:(a + 1) * g]
            pg = None if f.pref is None else f.pref[a * g:(a + 1) * g]
            out[a:a + 1] = nat.kmatw_spec(Xg, Xg, ones, spec, row_pref=pg, col_pref=pg, row_reduce=nat.REDUCE_SUM).reshape(-1)
        return out / float(g * g) if f.average else out
    return None


def reduced_gram(k, X, chain_x, Y, chain_y, self_gram: bool = False) -> torch.Tensor:
    """R_X K(X, Y) R_Y^T.  `self_gram` says Y is X (enables the symmetric self-Gram kernel)."""
    dev = X.device
    spec = kernel_spec_of(k)
    if spec is None:  # not a kernel the C ABI knows: the reference's own sequence
        return Reduce.apply(Reduce.apply(k(X, Y), chain_x, 0), chain_y, 1)
    gg = spec.kind == nat.KIND_GENGAUSS
    if gg:
        scale, power, nconst = spec.kparams
        dim_scale = spec.dim_scale
    fx = fold_chain(chain_x, X.shape[0], dev)
    fy = fold_chain(chain_y, Y.shape[0], dev)
    skinny = (LINEAR, SUM, MEAN)

    if needs_grad(k, X, Y, *_chain_tensors(chain_x), *_chain_tensors(chain_y)):
        if not (gg and dim_scale is None and X.shape[1] <= 64 and hasattr(k, "wants_grad")):
            raise grad_unsupported(f"inner product over {type(k).__name__} with d = {X.shape[1]}"
                                   + (" and per-dimension scales" if gg and dim_scale is not None else ""))
        out = _reduced_gram_autograd(k, spec, X, fx, Y, fy, self_gram)
        out = Reduce.apply(out, fx.post, 0)
        return Reduce.apply(out, fy.post, 1)

    if fy.kind in skinny and (fx.kind not in skinny or fy.out_len <= fx.out_len):
        out = _kmatw_side(X, Y, fx, fy, spec)
    elif fx.kind in skinny:
        out = _kmatw_side(Y, X, fy, fx, spec).t()
    elif (gg and fx.kind == BALANCED and fy.kind == BALANCED and dim_scale is None and 12 <= X.shape[1] <= 32
          and X.shape[0] * fy.group * 16 >= (1 << 30) and GROUPSUM_TENSOR
          and nat.lib().jrk_kmatw_uses_tensor_cores(X.shape[0] // 2, 16 * fy.group, X.shape[1], 16, 0, 0)):
        out = _groupsum_tensor(X, Y, fx, fy, scale, power, nconst, self_gram)
        if fx.average or fy.average:
            out = out * (1.0 / ((fx.group if fx.average else 1) * (fy.group if fy.average else 1)))
    elif gg and fx.kind == BALANCED and fy.kind == BALANCED and fx.group % 128 == 0:
        out = nat.gram_groupsum(X, None if self_gram else Y, scale, power, nconst, fx.group, fy.group,
                                False, dim_scale=dim_scale, row_pref=fx.pref, col_pref=fy.pref)
        if fx.average or fy.average:
            out = out * (1.0 / ((fx.group if fx.average else 1) * (fy.group if fy.average else 1)))
    elif gg and fx.kind == BALANCED and fy.kind == BALANCED and fy.group % 128 == 0:
        out = nat.gram_groupsum(Y, None if self_gram else X, scale, power, nconst, fy.group, fx.group,
                                False, dim_scale=dim_scale, row_pref=fy.pref, col_pref=fx.pref).t()
        if fx.average or fy.average:
            out = out * (1.0 / ((fx.group if fx.average else 1) * (fy.group if fy.average else 1)))
    elif fx.kind == NONE and fy.kind == NONE:
        # materialised Gram; Prefactors on both axes and a leading Center on either axis are applied in ONE pass over it
        # (jrk_gram_affine_f32).  C_N G C_M = G - 1 colmean^T - rowmean 1^T + mean: the means are skinny fused K@W
        # calls (they never read the Gram).
        cx = bool(fx.post) and type(fx.post[0]) is Center
        cy = bool(fy.post) and type(fy.post[0]) is Center
        row_add = col_add = None
        const = 0.0
        N_, M_ = X.shape[0], Y.shape[0]
        if cx:      # column means of G' = D_p K D_q:  q_j sum_i K_ji p_i / N
            col_add = -_kmatw_side(Y, X, Folded(n=M_, pref=fy.pref), Folded(n=N_, pref=fx.pref, kind=MEAN), spec).reshape(-1)
            fx.post = fx.post[1:]
        if cy:
            row_add = -_kmatw_side(X, Y, Folded(n=N_, pref=fx.pref), Folded(n=M_, pref=fy.pref, kind=MEAN), spec).reshape(-1)
            fy.post = fy.post[1:]
        if cx and cy:
            const = -float(row_add.double().mean())           # + mean(G')
        G = nat.gram_spec(X, None if self_gram else Y, spec)
        if fx.pref is not None or fy.pref is not None or cx or cy:
            nat.gram_affine(G, row_mul=fx.pref, col_mul=fy.pref, row_add=row_add, col_add=col_add, add_const=const)
        out = G
    else:
        G = nat.gram_spec(X, None if self_gram else Y, spec)
        if fx.pref is not None:
            G.mul_(fx.pref[:, None])
        if fy.pref is not None:
            G.mul_(fy.pref[None, :])
        if fx.kind == BALANCED:
            G = BalancedRed(fx.group, fx.average)(G, 0)
        if fy.kind == BALANCED:
            G = BalancedRed(fy.group, fy.average)(G, 1)
        out = G
    out = Reduce.apply(out, fx.post, 0)
    return Reduce.apply(out, fy.post, 1)

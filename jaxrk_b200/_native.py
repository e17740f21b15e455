"""ctypes binding of include/jrk.h (the C-ABI drop-in boundary).

This is binding "L1b" of SURVEY.md §8(b): device pointers come from torch CUDA tensors
(`Tensor.data_ptr()`), the stream is torch's current CUDA stream.  torch is plumbing
only (device memory + streams); every result is computed by libjaxrk_b200.so.

There is NO fallback: if the shared library is missing, or there is no CUDA device, the
calls raise.  Build the library with `python -m jaxrk_b200.build`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libjaxrk_b200.so")

PATH_AUTO, PATH_DIRECT, PATH_TF32X3 = 0, 1, 2
KIND_GENGAUSS, KIND_PERIODIC, KIND_THRESHSPIKE = 0, 1, 2
REDUCE_NONE, REDUCE_SUM, REDUCE_MEAN, REDUCE_BALANCED, REDUCE_BALANCED_MEAN = 0, 1, 2, 3, 4
OPT_KMATW_TC, OPT_TC_FAST, OPT_TC_V2, OPT_TC2_EPQ, OPT_GRAM_H16, OPT_TC2_TOKEN = 0, 1, 2, 3, 4, 5
ABI_VERSION = 3

# every symbol include/jrk.h declares: name -> (restype, argtypes)
_F = ctypes.POINTER(c_float)
SIGNATURES = {
    "jrk_abi_version": (c_int, []),
    "jrk_last_error": (c_int, [ctypes.c_char_p, c_size_t]),
    "jrk_device_count": (c_int, []),
    "jrk_launch_count": (c_int64, []),
    "jrk_set_option": (c_int, [c_int, c_int]),
    "jrk_get_option": (c_int, [c_int]),
    "jrk_kmatw_plan_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_int64,
                                      c_float, c_void_p, c_float, c_float, c_void_p, c_int, c_int, c_void_p]),
    "jrk_kmatw_plan_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int, c_int64]),
    "jrk_kmatw_plan_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int64,
                                     c_void_p, c_size_t, c_void_p]),
    "jrk_kmatw_plan_apply_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64]),
    "jrk_kmatw_plan_destroy": (c_int, [c_void_p]),
    "jrk_gram_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "jrk_gram_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_int64,
                             c_float, c_void_p, c_float, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "jrk_dist_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_int64,
                             c_float, c_void_p, c_float, c_void_p]),
    "jrk_dist_diag_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, c_float, c_void_p,
                                  c_float, c_void_p]),
    "jrk_gram_affine_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_float, c_void_p]),
    "jrk_gram_kind_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_int64,
                                  c_int, _F, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "jrk_kmatw_kind_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_int64, c_int64, c_int64, c_int64, c_int, _F, c_int, c_void_p,
                                   c_void_p, c_void_p, c_int, c_int64, c_void_p, c_size_t, c_void_p]),
    "jrk_grad_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "jrk_kmatw_grad_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                   c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_float, c_float,
                                   c_void_p, c_size_t, c_void_p]),
    "jrk_gram_grad_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                  c_int64, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float,
                                  c_void_p, c_size_t, c_void_p]),
    "jrk_kmatw_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int, c_int64]),
    "jrk_kmatw_uses_tensor_cores": (c_int, [c_int64, c_int64, c_int, c_int, c_int, c_int]),
    "jrk_kmatw_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                              c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_float, c_float,
                              c_void_p, c_void_p, c_int, c_int64, c_void_p, c_size_t, c_void_p]),
    "jrk_gram_groupsum_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64,
                                      c_int64, c_float, c_void_p, c_float, c_float, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "jrk_gram_f32_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p,
                                  c_float, c_float, c_int, c_int]),
    "jrk_kmatw_f32_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_float, c_void_p, c_float, c_float, c_void_p, c_void_p, c_int, c_int64, c_int]),
}


class JrkError(RuntimeError):
    """Non-zero status from the C ABI (1 invalid arg, 2 CUDA, 3 unsupported)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libjaxrk_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """Load libjaxrk_b200.so (once).  Raises -- never falls back -- when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m jaxrk_b200.build` "
                "(jaxrk_b200 has no CPU or eager fallback).")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.jrk_abi_version() != ABI_VERSION:
            raise ImportError("libjaxrk_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    lib().jrk_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def _check(rc: int):
    if rc != 0:
        raise JrkError(rc, last_error())


def launch_count() -> int:
    return int(lib().jrk_launch_count())


def set_option(option: int, value: int) -> int:
    """jrk_set_option; returns the previous value (so that a caller can restore it)."""
    prev = int(lib().jrk_get_option(option))
    _check(lib().jrk_set_option(option, int(value)))
    return prev


def get_option(option: int) -> int:
    return int(lib().jrk_get_option(option))


class option:
    """`with nat.option(nat.OPT_KMATW_TC, 0): ...` -- a process-wide switch for the duration of a block (tests, A/B runs)."""

    def __init__(self, opt: int, value: int):
        self.opt, self.value = opt, value

    def __enter__(self):
        self.prev = set_option(self.opt, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.opt, self.prev)
        return False


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(dev: torch.device):
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dev2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
        raise TypeError(f"{name} must be a 2-D float32 CUDA tensor")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def _vec(t: Optional[torch.Tensor], n: int, name: str, dev) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = torch.as_tensor(t, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    if t.numel() != n:
        raise ValueError(f"{name} must have {n} elements, got {t.numel()}")
    return t


def gram(X, Y, scale: float, power: float, nconst: float, dim_scale=None, path: int = PATH_AUTO,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """jrk_gram_f32: K[i, j] = nconst * exp(-(scale * ||ds*x_i - ds*y_j||)^power)."""
    X = _dev2d(X, "X")
    self_gram = Y is None
    Yt = X if self_gram else _dev2d(Y, "Y")
    N, d = X.shape
    M = Yt.shape[0]
    if Yt.shape[1] != d:
        raise ValueError("X and Y must have the same number of columns")
    ds = _vec(dim_scale, d, "dim_scale", X.device)
    if out is None:
        out = torch.empty((N, M), dtype=torch.float32, device=X.device)
    if N == 0 or M == 0:  # empty tensors have NULL data pointers; nothing to compute
        return out
    with torch.cuda.device(X.device):
        wsb = lib().jrk_gram_workspace_bytes(N, M, d, path)
        ws = torch.empty(wsb, dtype=torch.uint8, device=X.device) if wsb else None
        _check(lib().jrk_gram_f32(_ptr(X), None if self_gram else _ptr(Yt), _ptr(out), N, M, d, _ld(X), _ld(Yt),
                                  _ld(out) if N > 1 else M, scale, _ptr(ds), power, nconst, path, _ptr(ws), wsb,
                                  _stream(X.device)))
    return out


def dist(X, Y, scale: float, power: float, dim_scale=None) -> torch.Tensor:
    """jrk_dist_f32: D[i, j] = (scale * ||ds*x_i - ds*y_j||)^power."""
    X = _dev2d(X, "X")
    self_gram = Y is None
    Yt = X if self_gram else _dev2d(Y, "Y")
    N, d = X.shape
    M = Yt.shape[0]
    if Yt.shape[1] != d:
        raise ValueError("X and Y must have the same number of columns")
    ds = _vec(dim_scale, d, "dim_scale", X.device)
    out = torch.empty((N, M), dtype=torch.float32, device=X.device)
    if N == 0 or M == 0:
        return out
    with torch.cuda.device(X.device):
        _check(lib().jrk_dist_f32(_ptr(X), None if self_gram else _ptr(Yt), _ptr(out), N, M, d, _ld(X), _ld(Yt), M,
                                  scale, _ptr(ds), power, _stream(X.device)))
    return out


def kmatw(X, Y, W, scale: float, power: float, nconst: float, dim_scale=None, row_pref=None, col_pref=None,
          row_reduce: int = REDUCE_NONE, group: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """jrk_kmatw_f32: row_reduce(diag(row_pref) K(X, Y) diag(col_pref) W), never materialising K."""
    X = _dev2d(X, "X")
    Y = _dev2d(Y, "Y")
    W = _dev2d(W, "W")
    N, d = X.shape
    M, m = W.shape
    if Y.shape != (M, d):
        raise ValueError(f"Y must be {M} x {d}, got {tuple(Y.shape)}")
    dev = X.device
    ds = _vec(dim_scale, d, "dim_scale", dev)
    rp = _vec(row_pref, N, "row_pref", dev)
    cp = _vec(col_pref, M, "col_pref", dev)
    if row_reduce in (REDUCE_SUM, REDUCE_MEAN):
        rows = 1
    elif row_reduce in (REDUCE_BALANCED, REDUCE_BALANCED_MEAN):
        if group < 1 or N % group:
            raise ValueError("group must divide N")
        rows = N // group
    else:
        rows = N
    if out is None:
        out = torch.empty((rows, m), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        wsb = lib().jrk_kmatw_workspace_bytes(N, M, d, m, row_reduce, group)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
        _check(lib().jrk_kmatw_f32(_ptr(X), _ptr(Y), _ptr(W), _ptr(out), N, M, d, m, _ld(X), _ld(Y), _ld(W), m,
                                   scale, _ptr(ds), power, nconst, _ptr(rp), _ptr(cp), row_reduce, group,
                                   _ptr(ws), wsb, _stream(dev)))
    return out


def gram_groupsum(X, Y, scale: float, power: float, nconst: float, gx: int, gy: int, average: bool,
                  dim_scale=None, row_pref=None, col_pref=None) -> torch.Tensor:
    """jrk_gram_groupsum_f32: BalancedRed on both axes of K(X, Y) without materialising it."""
    X = _dev2d(X, "X")
    self_gram = Y is None
    Yt = X if self_gram else _dev2d(Y, "Y")
    N, d = X.shape
    M = Yt.shape[0]
    dev = X.device
    ds = _vec(dim_scale, d, "dim_scale", dev)
    rp = _vec(row_pref, N, "row_pref", dev)
    cp = _vec(col_pref, M, "col_pref", dev)
    out = torch.empty((N // gx, M // gy), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib().jrk_gram_groupsum_f32(_ptr(X), None if self_gram else _ptr(Yt), _ptr(out), N, M, d, _ld(X),
                                           _ld(Yt), M // gy, scale, _ptr(ds), power, nconst, _ptr(rp), _ptr(cp),
                                           gx, gy, int(bool(average)), None, 0, _stream(dev)))
    return out


def dist_diag(X, Y, scale: float, power: float, dim_scale=None) -> torch.Tensor:
    """jrk_dist_diag_f32: out[i] = scale * sum_k |ds_k x_ik - ds_k y_ik|^power (the diag=True formula)."""
    X = _dev2d(X, "X")
    Y = _dev2d(Y, "Y")
    if X.shape != Y.shape:
        raise ValueError("X and Y must have the same shape")
    N, d = X.shape
    ds = _vec(dim_scale, d, "dim_scale", X.device)
    out = torch.empty((N,), dtype=torch.float32, device=X.device)
    if N == 0:
        return out
    with torch.cuda.device(X.device):
        _check(lib().jrk_dist_diag_f32(_ptr(X), _ptr(Y), _ptr(out), N, d, _ld(X), _ld(Y), scale, _ptr(ds), power,
                                       _stream(X.device)))
    return out


def gram_affine(K, row_mul=None, col_mul=None, row_add=None, col_add=None, add_const: float = 0.0, out=None) -> torch.Tensor:
    """jrk_gram_affine_f32: out = K * row_mul[:, None] * col_mul[None, :] + row_add[:, None] + col_add[None, :] + add_const
    in one pass; in place (out = K) by default."""
    K = _dev2d(K, "K")
    N, M = K.shape
    out = K if out is None else out
    if out.shape != K.shape or out.dtype != torch.float32 or out.device != K.device or out.stride(1) != 1:
        raise ValueError("out must be a float32 N x M matrix with unit column stride on K's device")
    vs = [_vec(v, n, name, K.device) for v, n, name in ((row_mul, N, "row_mul"), (col_mul, M, "col_mul"),
                                                         (row_add, N, "row_add"), (col_add, M, "col_add"))]
    if N == 0 or M == 0:
        return out
    with torch.cuda.device(K.device):
        _check(lib().jrk_gram_affine_f32(_ptr(K), _ptr(out), N, M, _ld(K) if N > 1 else M, _ld(out) if N > 1 else M,
                                         _ptr(vs[0]), _ptr(vs[1]), _ptr(vs[2]), _ptr(vs[3]), float(add_const), _stream(K.device)))
    return out


def kmatw_grad(X, Y, W, G, scale: float, power: float, nconst: float, want_scale: bool = True, want_power: bool = False):
    """jrk_kmatw_grad_f32: (gX, gs_rows) for out = K(X, Y) W with cotangent G (N x m); m <= 32, d <= 64.
    want_power: a third result gp_rows (per-row dL/dpower at fixed nconst)."""
    X, Y, W, G = _dev2d(X, "X"), _dev2d(Y, "Y"), _dev2d(W, "W"), _dev2d(G, "G")
    N, d = X.shape
    M, m = W.shape
    if Y.shape != (M, d) or G.shape != (N, m):
        raise ValueError("shape mismatch")
    dev = X.device
    gX = torch.empty((N, d), dtype=torch.float32, device=dev)
    gs = torch.empty((N,), dtype=torch.float32, device=dev) if want_scale else None
    gp = torch.empty((N,), dtype=torch.float32, device=dev) if want_power else None
    with torch.cuda.device(dev):
        wsb = lib().jrk_grad_workspace_bytes(N, M, d)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _check(lib().jrk_kmatw_grad_f32(_ptr(X), _ptr(Y), _ptr(W), _ptr(G), _ptr(gX), _ptr(gs), _ptr(gp), N, M, d, m, _ld(X),
                                        _ld(Y), _ld(W), _ld(G), d, scale, power, nconst, _ptr(ws), wsb, _stream(dev)))
    return (gX, gs, gp) if want_power else (gX, gs)


def gram_grad(X, Y, Gbar, scale: float, power: float, nconst: float, want_scale: bool = True, want_power: bool = False,
              transposed: bool = False):
    """jrk_gram_grad_f32: (gX, gs_rows[, gp_rows]) for K = K(X, Y) with cotangent Gbar (N x M); d <= 64.
    transposed: `Gbar` holds the cotangent transposed (M x N) -- the layout the kernels read fastest."""
    X, Y, Gbar = _dev2d(X, "X"), _dev2d(Y, "Y"), _dev2d(Gbar, "Gbar")
    N, d = X.shape
    M = Y.shape[0]
    if Y.shape[1] != d or Gbar.shape != ((M, N) if transposed else (N, M)):
        raise ValueError("shape mismatch")
    dev = X.device
    gX = torch.empty((N, d), dtype=torch.float32, device=dev)
    gs = torch.empty((N,), dtype=torch.float32, device=dev) if want_scale else None
    gp = torch.empty((N,), dtype=torch.float32, device=dev) if want_power else None
    with torch.cuda.device(dev):
        wsb = lib().jrk_grad_workspace_bytes(N, M, d)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _check(lib().jrk_gram_grad_f32(_ptr(X), _ptr(Y), _ptr(Gbar), _ptr(gX), _ptr(gs), _ptr(gp), N, M, d, _ld(X), _ld(Y),
                                       _ld(Gbar), d, 1 if transposed else 0, scale, power, nconst, _ptr(ws), wsb, _stream(dev)))
    return (gX, gs, gp) if want_power else (gX, gs)


class KernelSpec:
    """(kind, kparams, dim_scale) of include/jrk.h's *_kind_f32 entry points."""
    __slots__ = ("kind", "kparams", "dim_scale")

    def __init__(self, kind: int, kparams, dim_scale=None):
        self.kind, self.kparams, self.dim_scale = int(kind), [float(v) for v in kparams], dim_scale

    def _c(self):
        return (c_float * len(self.kparams))(*self.kparams), len(self.kparams)


def gram_spec(X, Y, spec: KernelSpec, path: int = PATH_AUTO, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """jrk_gram_kind_f32 (GENGAUSS forwards to jrk_gram_f32 with all its paths)."""
    if spec.kind == KIND_GENGAUSS:
        scale, power, nconst = spec.kparams
        return gram(X, Y, scale, power, nconst, dim_scale=spec.dim_scale, path=path, out=out)
    X = _dev2d(X, "X")
    self_gram = Y is None
    Yt = X if self_gram else _dev2d(Y, "Y")
    N, d = X.shape
    M = Yt.shape[0]
    if Yt.shape[1] != d:
        raise ValueError("X and Y must have the same number of columns")
    ds = _vec(spec.dim_scale, d, "dim_scale", X.device)
    if out is None:
        out = torch.empty((N, M), dtype=torch.float32, device=X.device)
    if N == 0 or M == 0:
        return out
    kp, nkp = spec._c()
    with torch.cuda.device(X.device):
        _check(lib().jrk_gram_kind_f32(_ptr(X), None if self_gram else _ptr(Yt), _ptr(out), N, M, d, _ld(X), _ld(Yt),
                                       _ld(out) if N > 1 else M, spec.kind, kp, nkp, _ptr(ds), PATH_AUTO, None, 0,
                                       _stream(X.device)))
    return out


def kmatw_spec(X, Y, W, spec: KernelSpec, row_pref=None, col_pref=None, row_reduce: int = REDUCE_NONE,
               group: int = 0) -> torch.Tensor:
    """jrk_kmatw_kind_f32."""
    if spec.kind == KIND_GENGAUSS:
        scale, power, nconst = spec.kparams
        return kmatw(X, Y, W, scale, power, nconst, dim_scale=spec.dim_scale, row_pref=row_pref, col_pref=col_pref,
                     row_reduce=row_reduce, group=group)
    X = _dev2d(X, "X")
    Y = _dev2d(Y, "Y")
    W = _dev2d(W, "W")
    N, d = X.shape
    M, m = W.shape
    if Y.shape != (M, d):
        raise ValueError(f"Y must be {M} x {d}, got {tuple(Y.shape)}")
    dev = X.device
    ds = _vec(spec.dim_scale, d, "dim_scale", dev)
    rp = _vec(row_pref, N, "row_pref", dev)
    cp = _vec(col_pref, M, "col_pref", dev)
    if row_reduce in (REDUCE_SUM, REDUCE_MEAN):
        rows = 1
    elif row_reduce in (REDUCE_BALANCED, REDUCE_BALANCED_MEAN):
        if group < 1 or N % group:
            raise ValueError("group must divide N")
        rows = N // group
    else:
        rows = N
    out = torch.empty((rows, m), dtype=torch.float32, device=dev)
    kp, nkp = spec._c()
    with torch.cuda.device(dev):
        wsb = lib().jrk_kmatw_workspace_bytes(N, M, d, m, row_reduce, group)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
        _check(lib().jrk_kmatw_kind_f32(_ptr(X), _ptr(Y), _ptr(W), _ptr(out), N, M, d, m, _ld(X), _ld(Y), _ld(W), m,
                                        spec.kind, kp, nkp, _ptr(ds), _ptr(rp), _ptr(cp), row_reduce, group,
                                        _ptr(ws), wsb, _stream(dev)))
    return out


# ---- host-buffer (end-to-end) entry points: numpy / pinned torch CPU tensors in and out ----
def _host_ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        assert a.device.type == "cpu" and a.dtype == torch.float32 and a.is_contiguous()
        return c_void_p(a.data_ptr())
    import numpy as np
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return c_void_p(a.ctypes.data)


def gram_host(X, Y, out, scale: float, power: float, nconst: float, dim_scale=None, path: int = PATH_AUTO,
              device: int = 0):
    """jrk_gram_f32_host: host arrays in, host Gram out (copies inside the call)."""
    N, d = X.shape
    M = N if Y is None else Y.shape[0]
    assert tuple(out.shape) == (N, M)
    _check(lib().jrk_gram_f32_host(_host_ptr(X), _host_ptr(Y), _host_ptr(out), N, M, d, scale,
                                   _host_ptr(dim_scale), power, nconst, path, device))
    return out


def kmatw_host(X, Y, W, out, scale: float, power: float, nconst: float, dim_scale=None, row_pref=None,
               col_pref=None, row_reduce: int = REDUCE_NONE, group: int = 0, device: int = 0):
    """jrk_kmatw_f32_host: host arrays in, host result out (copies inside the call)."""
    N, d = X.shape
    M, m = W.shape
    _check(lib().jrk_kmatw_f32_host(_host_ptr(X), _host_ptr(Y), _host_ptr(W), _host_ptr(out), N, M, d, m, scale,
                                    _host_ptr(dim_scale), power, nconst, _host_ptr(row_pref), _host_ptr(col_pref),
                                    row_reduce, group, device))
    return out


def _rows_out(N: int, row_reduce: int, group: int) -> int:
    if row_reduce in (REDUCE_SUM, REDUCE_MEAN):
        return 1
    if row_reduce in (REDUCE_BALANCED, REDUCE_BALANCED_MEAN):
        if group < 1 or N % group:
            raise ValueError("group must divide N")
        return N // group
    return N


class KmatwPlan:
    """jrk_kmatw_plan_*: the Y / W side of row_reduce(diag(row_pref) K(X, Y) diag(col_pref) W) kept on the device
    (copies of Y, W and -- for the tensor-core kernels -- the packed tile images).  `apply` takes CUDA tensors and
    enqueues on torch's current stream; `apply_host` takes host arrays and includes the copies."""

    def __init__(self, Y, W, scale: float, power: float, nconst: float, dim_scale=None, col_pref=None, device: int = 0):
        host = not (isinstance(Y, torch.Tensor) and Y.is_cuda)
        if host:
            import numpy as np
            Y = np.ascontiguousarray(Y, dtype=np.float32) if not isinstance(Y, torch.Tensor) else Y.contiguous().float()
            W = np.ascontiguousarray(W, dtype=np.float32) if not isinstance(W, torch.Tensor) else W.contiguous().float()
            yp, wp = _host_ptr(Y), _host_ptr(W)
            ldy, ldw = int(Y.shape[1]), int(W.shape[1])
            dsp = _host_ptr(None if dim_scale is None else np.ascontiguousarray(dim_scale, dtype=np.float32).reshape(-1))
            cpp = _host_ptr(None if col_pref is None else np.ascontiguousarray(col_pref, dtype=np.float32).reshape(-1))
            stream, self.device = None, int(device)
        else:
            Y, W = _dev2d(Y, "Y"), _dev2d(W, "W")
            yp, wp, ldy, ldw = _ptr(Y), _ptr(W), _ld(Y), _ld(W)
            ds = _vec(dim_scale, Y.shape[1], "dim_scale", Y.device)
            cp = _vec(col_pref, Y.shape[0], "col_pref", Y.device)
            dsp, cpp = _ptr(ds), _ptr(cp)
            stream, self.device = _stream(Y.device), int(Y.device.index or 0)
        M, d = int(Y.shape[0]), int(Y.shape[1])
        if int(W.shape[0]) != M:
            raise ValueError(f"W must have {M} rows")
        self.M, self.d, self.m = M, d, int(W.shape[1])
        h = c_void_p()
        _check(lib().jrk_kmatw_plan_create(ctypes.byref(h), yp, wp, M, d, self.m, ldy, ldw, scale, dsp, power, nconst, cpp,
                                           1 if host else 0, self.device, stream))
        self._h = h

    def apply(self, X, row_pref=None, row_reduce: int = REDUCE_NONE, group: int = 0, out: Optional[torch.Tensor] = None):
        X = _dev2d(X, "X")
        N = int(X.shape[0])
        if X.shape[1] != self.d:
            raise ValueError(f"X must have {self.d} columns")
        rp = _vec(row_pref, N, "row_pref", X.device)
        rows = _rows_out(N, row_reduce, group)
        if out is None:
            out = torch.empty((rows, self.m), dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            wsb = lib().jrk_kmatw_plan_workspace_bytes(self._h, N, row_reduce, group)
            ws = torch.empty(wsb, dtype=torch.uint8, device=X.device) if wsb else None
            _check(lib().jrk_kmatw_plan_apply(self._h, _ptr(X), _ptr(out), N, _ld(X), self.m, _ptr(rp), row_reduce, group,
                                              _ptr(ws), wsb, _stream(X.device)))
        return out

    def apply_host(self, X, out, row_pref=None, row_reduce: int = REDUCE_NONE, group: int = 0):
        N = int(X.shape[0])
        assert tuple(out.shape) == (_rows_out(N, row_reduce, group), self.m)
        _check(lib().jrk_kmatw_plan_apply_host(self._h, _host_ptr(X), _host_ptr(out), N, _host_ptr(row_pref), row_reduce, group))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().jrk_kmatw_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

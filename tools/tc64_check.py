"""The d <= 64 variant of the second-generation tensor-core K@W kernel (kmatw_tc2_kernel<4, 4>): accuracy against float64
on sampled rows, every row against the FP32-pipe kernel, the fallback outside the factored regime, and speed.
Usage: python tools/tc64_check.py [acc|time|all]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")


def randn(seed, *shape):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)


def acc(N, M, d, m, ls, positive=False, seed=0):
    X, Y, W = randn(seed, N, d), randn(seed + 1, M, d), randn(seed + 2, M, m)
    if positive:
        W = W.abs()
    s, nc = float(0.70710695 / ls), 0.25
    c = s * s * 1.4426950408889634
    bnd = c * float((X * X).sum(1).max() + (Y * Y).sum(1).max())
    n0 = nat.launch_count()
    A = nat.kmatw(X, Y, W, s, 2.0, nc)
    nl = nat.launch_count() - n0
    B = nat.kmatw(X, Y, W, s, 2.0, nc)
    with nat.option(nat.OPT_KMATW_TC, 0):
        R = nat.kmatw(X, Y, W, s, 2.0, nc)
        bound_all = nat.kmatw(X, Y, W.abs(), s, 2.0, nc)
    rows = torch.cat([torch.arange(0, 40, device=dev), torch.arange(40, N, max(1, (N - 40) // 88), device=dev)[:88]])
    Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** 2)
    T, bound = Kt @ W.double(), Kt @ W.double().abs()
    e64 = float(((A[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max())
    e64r = float(((R[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max())
    eall = float(((A - R).abs() / (2e-6 + 3e-5 * bound_all)).max())
    print(f"N={N} M={M} d={d} m={m} ls={ls} bound={bnd:.2f} pos={positive}: launches {nl}, twice equal {torch.equal(A, B)}, "
          f"err/lim vs float64 {e64:.3f} (FP32-pipe kernel {e64r:.3f}), every row vs FP32-pipe kernel {eall:.3f}, "
          f"identical to it {torch.equal(A, R)}", flush=True)


def timeit(N, M, d, m, ls):
    X, Y, W = randn(0, N, d), randn(1, M, d), randn(2, M, m)
    s, nc = float(0.70710695 / ls), 0.25
    out = {}
    for tc in (1, 0):
        with nat.option(nat.OPT_KMATW_TC, tc):
            for _ in range(2):
                nat.kmatw(X, Y, W, s, 2.0, nc)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nat.kmatw(X, Y, W, s, 2.0, nc)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out["tc" if tc else "fp32"] = min(ts)
    print(f"N={N} M={M} d={d} m={m}: tensor cores {out['tc']:.2f} ms = {N * M / out['tc'] * 1e3:.3e} pairs/s "
          f"({N * M / out['tc'] * 1e3 / 4.637e12:.3f} of MUFU), FP32 pipe {out['fp32']:.2f} ms = {N * M / out['fp32'] * 1e3:.3e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("acc", "all"):
        acc(16384, 8192, 64, 16, 8.0)
        acc(50000, 4109, 48, 16, 7.0, positive=True, seed=3)
        acc(16384 + 77, 6000, 33, 5, 6.0, seed=5)
        acc(20000, 8192, 64, 24, 8.0, positive=True, seed=7)
        acc(16384, 8192, 64, 16, 3.0, seed=9)              # outside the factored regime: the FP32-pipe kernel must serve it
    if what in ("time", "all"):
        timeit(131072, 1 << 20, 64, 16, 8.0)
        timeit(131072, 1 << 20, 48, 16, 7.0)

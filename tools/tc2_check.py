"""Second-generation tensor-core K@W kernel (kmatw_tc2.cu): accuracy against float64 and timing against the first
generation (JRK_TC_V2=0) and the FP32-pipe kernel.   python tools/tc2_check.py [acc|time|full|all]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
VARIANTS = (("v2", {}), ("v2/tok1", {nat.OPT_TC2_TOKEN: 1}), ("v2/tok2", {nat.OPT_TC2_TOKEN: 2}), ("v2/tok3", {nat.OPT_TC2_TOKEN: 3}),
            ("v2/epq3", {nat.OPT_TC2_EPQ: 3}), ("v2/epq3tok2", {nat.OPT_TC2_EPQ: 3, nat.OPT_TC2_TOKEN: 2}), ("v1", {nat.OPT_TC_V2: 0}))


class Options:
    def __init__(self, opts):
        self.opts = opts

    def __enter__(self):
        self.prev = {k: nat.set_option(k, v) for k, v in self.opts.items()}

    def __exit__(self, *exc):
        for k, v in self.prev.items():
            nat.set_option(k, v)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def err_vs_truth(A, X, Y, W, s, nc, rows, rp=None, cp=None):
    Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** 2)
    if cp is not None:
        Kt = Kt * cp.double()[None, :]
    if rp is not None:
        Kt = Kt * rp[rows].double()[:, None]
    T = Kt @ W.double()
    bound = Kt.abs() @ W.double().abs()
    return ((A[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max().item()


def acc(N, M, d, m, ls_factor=1.0, positive=False, prefs=False, wscale=1.0):
    X = torch.randn(N, d, generator=g, device=dev)
    Y = torch.randn(M, d, generator=g, device=dev)
    W = (torch.rand(M, m, generator=g, device=dev) if positive else torch.randn(M, m, generator=g, device=dev)) * wscale
    rp = torch.rand(N, generator=g, device=dev) if prefs else None
    cp = torch.rand(M, generator=g, device=dev) if prefs else None
    s, nc = float(0.70710695 / (np.sqrt(d) * ls_factor)), 0.25
    c = s * s * 1.4426950408889634
    bnd = c * float((X * X).sum(1).max() + (Y * Y).sum(1).max())
    rows = torch.unique(torch.cat([torch.arange(0, N, max(1, N // 192), device=dev), torch.arange(N - 64, N, device=dev),
                                   torch.arange(0, 64, device=dev)]))
    res = {}
    for name, opts in VARIANTS + (("fp32", {nat.OPT_KMATW_TC: 0}),):
        with Options(opts):
            A = nat.kmatw(X, Y, W, s, 2.0, nc, row_pref=rp, col_pref=cp)
            torch.cuda.synchronize()
        res[name] = err_vs_truth(A, X, Y, W, s, nc, rows, rp, cp)
    print(f"N={N} M={M} d={d} m={m} ls*{ls_factor} bound={bnd:.2f} pos={positive} prefs={prefs} wscale={wscale:g}: err/lim "
          + "  ".join(f"{k} {v:.3f}" for k, v in res.items()), flush=True)
    return max(res["v2"], res["v2/epq3"], res["v2/tok2"])


def time_one(N, M, d, m, reps=3):
    X = torch.randn(N, d, generator=g, device=dev)
    Y = torch.randn(M, d, generator=g, device=dev)
    W = torch.rand(M, m, generator=g, device=dev)
    s, nc = float(0.70710695 / np.sqrt(d)), 0.25
    out = {}
    for name, opts in VARIANTS:
        with Options(opts):
            out[name] = timeit(lambda: nat.kmatw(X, Y, W, s, 2.0, nc), reps)
    plan = nat.KmatwPlan(Y, W, s, 2.0, nc)
    out["v2/plan"] = timeit(lambda: plan.apply(X), reps)
    P = plan.apply(X)
    plan.close()
    rows = torch.arange(0, N, N // 64, device=dev)[:64]
    A = nat.kmatw(X, Y, W, s, 2.0, nc)
    assert torch.equal(A, P), "a plan apply must give the one-shot result bit for bit"
    e = err_vs_truth(A, X, Y, W, s, nc, rows)
    print(f"N={N} M={M} d={d} m={m}: " + "  ".join(f"{k} {v:.2f} ms ({N * M / v / 1e9:.3f}e12 pairs/s, {N * M / v / 1e6 / 4637:.3f} of MUFU)"
                                                    for k, v in out.items()) + f"  v2 err/lim {e:.3f}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("acc", "all"):
        worst = 0.0
        worst = max(worst, acc(8192, 4096, 16, 16))
        worst = max(worst, acc(8192 + 77, 4096 + 33, 16, 16))
        worst = max(worst, acc(8192 + 77, 4096 + 33, 7, 5))
        worst = max(worst, acc(8192 + 77, 4096 + 33, 32, 1))
        worst = max(worst, acc(8192 + 77, 4096 + 33, 20, 16, prefs=True))
        worst = max(worst, acc(8192 + 77, 4096 + 33, 1, 3))
        worst = max(worst, acc(16384, 65536, 16, 16, positive=True))
        worst = max(worst, acc(16384, 65536, 16, 16, ls_factor=0.62, positive=True))
        worst = max(worst, acc(16384, 65536, 16, 16, positive=True, wscale=1e-20))
        worst = max(worst, acc(16384, 65536, 16, 16, positive=False, wscale=1e20))
        worst = max(worst, acc(16384, 65536, 32, 16, positive=True, prefs=True))
        print("worst v2 err/lim", worst, "OK" if worst < 1.0 else "FAIL", flush=True)
    if what in ("time", "all"):
        time_one(131072, 1 << 20, 16, 16)
        time_one(131072, 1 << 20, 8, 16)
        time_one(131072, 1 << 20, 32, 16)
        time_one(131072, 1 << 20, 16, 1)
    if what in ("full", "all"):
        time_one(1 << 20, 1 << 20, 16, 16, reps=2)

"""Numerical check + timing of the tensor-core K@W path against the FP32-pipe kernel and float64."""
import os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
def run(N, M, d, m, p, positive=False, timing=False):
    X = torch.randn(N, d, generator=g, device=dev); Y = torch.randn(M, d, generator=g, device=dev)
    W = torch.rand(M, m, generator=g, device=dev) if positive else torch.randn(M, m, generator=g, device=dev)
    s, nc = float(0.70710695 / np.sqrt(d)) if p == 2.0 else float(1 / np.sqrt(d)), 0.25
    nat.set_option(nat.OPT_KMATW_TC, 1); A = nat.kmatw(X, Y, W, s, p, nc); torch.cuda.synchronize()
    nat.set_option(nat.OPT_KMATW_TC, 0); B = nat.kmatw(X, Y, W, s, p, nc); torch.cuda.synchronize()
    rows = torch.arange(0, N, max(1, N // 64), device=dev)[:64]
    Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** p)
    T = Kt @ W.double(); bound = Kt @ W.double().abs()
    ea = ((A[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max().item()
    eb = ((B[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max().item()
    msg = f"N={N} M={M} d={d} m={m} p={p} pos={positive}: tc err/lim {ea:.3f}  fp32 err/lim {eb:.3f}"
    if timing:
        nat.set_option(nat.OPT_KMATW_TC, 1); ta = t(lambda: nat.kmatw(X, Y, W, s, p, nc))
        nat.set_option(nat.OPT_KMATW_TC, 0); tb = t(lambda: nat.kmatw(X, Y, W, s, p, nc))
        msg += f" | tc {ta:.2f} ms ({N*M/ta/1e6:.0f} Gpairs/s)  fp32 {tb:.2f} ms ({N*M/tb/1e6:.0f} Gpairs/s)"
    print(msg, flush=True)
import sys as _s
if len(_s.argv) > 1 and _s.argv[1] == "fast1":
    run(131072, 1 << 20, 16, 16, 2.0, positive=True, timing=True)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "var":
    for v in _s.argv[2:]:
        os.environ["JRK_TC_VAR"] = v; print("VAR", v, end=": "); run(131072, 1 << 20, 16, 16, 2.0, positive=True, timing=True)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "edge":
    # accuracy at the edge of the factored regime: length scales that put |c| (max|x|^2 + max|y|^2) near the bound
    for f in (1.0, 0.8, 0.7, 0.62, 0.6):
        N, M, d, m = 16384, 65536, 16, 16
        X = torch.randn(N, d, generator=g, device=dev); Y = torch.randn(M, d, generator=g, device=dev)
        s = float(0.70710695 / (np.sqrt(d) * f)); nc = 0.25
        c = s * s * 1.4426950408889634
        bnd = c * float((X * X).sum(1).max() + (Y * Y).sum(1).max())
        for positive in (True, False):
            W = torch.rand(M, m, generator=g, device=dev) if positive else torch.randn(M, m, generator=g, device=dev)
            nat.set_option(nat.OPT_KMATW_TC, 1)
            for fast in ("1", "0"):
                nat.set_option(nat.OPT_TC_FAST, int(fast))
                A = nat.kmatw(X, Y, W, s, 2.0, nc); torch.cuda.synchronize()
                rows = torch.arange(0, N, N // 64, device=dev)[:64]
                Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** 2)
                T = Kt @ W.double(); bound = Kt @ W.double().abs()
                ea = ((A[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max().item()
                print(f"f={f} bound={bnd:.2f} positive={positive} fast={fast}: err/lim {ea:.3f}", flush=True)
    _s.exit(0)
if len(_s.argv) > 1 and _s.argv[1] == "fast":
    for f in ("1", "0"):
        nat.set_option(nat.OPT_TC_FAST, int(f)); print("JRK_TC_FAST", f)
        run(16384, 65536, 16, 16, 2.0, positive=True)
        run(131072, 1 << 20, 16, 16, 2.0, positive=True, timing=True)
        run(131072, 1 << 20, 8, 16, 2.0, timing=True)
        run(131072, 1 << 20, 32, 16, 2.0, timing=True)
    _s.exit(0)
run(8192, 4096, 16, 16, 2.0)
run(8192 + 77, 4096 + 33, 16, 16, 2.0)
run(8192, 8192, 7, 5, 1.0)
run(8192, 8192, 32, 1, 1.5)
run(16384, 65536, 16, 16, 2.0, positive=True)
run(131072, 1 << 20, 16, 16, 2.0, positive=True, timing=True)
run(131072, 1 << 20, 16, 16, 1.0, timing=True)
run(131072, 1 << 20, 32, 16, 2.0, timing=True)
print("ok")

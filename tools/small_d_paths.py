import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from jaxrk_b200 import _native as nat
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
N = 65536
K = torch.empty((N, N), device=dev)
for d in (8, 16, 32, 48):
    X = torch.randn(N, d, generator=g, device=dev); Y = torch.randn(N, d, generator=g, device=dev)
    s = float(1/np.sqrt(d))
    for p in (2.0, 1.0):
        a = t(lambda: nat.gram(X, Y, s, p, 0.25, path=1, out=K)); b = t(lambda: nat.gram(X, Y, s, p, 0.25, path=2, out=K))
        # accuracy of path 2 on a sample
        Ks = nat.gram(X[:512], Y[:2048], s, p, 0.25, path=2)
        T = 0.25 * torch.exp(-(s * torch.cdist(X[:512].double(), Y[:2048].double())) ** p)
        err = ((Ks.double() - T).abs() / (1e-6 + 1e-5 * T.abs())).max().item()
        print(f"d={d} p={p}: direct {a:.3f} ms ({N*N/a/1e9:.0f} G/s)  tf32x3 {b:.3f} ms ({N*N/b/1e9:.0f} G/s)  tf32 err/lim {err:.3f}", flush=True)

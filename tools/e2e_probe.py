"""e2e probe: jrk_gram_f32_host on cfg2 through one 17.2 GB pinned buffer vs four 16384-row slabs (what bench.py does),
and the raw pinned D2H bandwidth of the box."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
N = M = 65536; d = 64
X = torch.randn(N, d).pin_memory(); Y = torch.randn(M, d).pin_memory()
s = float(0.70710695 / np.sqrt(d))
def timed(fn, reps=3):
    fn(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best
slab = 16384
out_s = torch.empty((slab, M)).pin_memory()
def slabs():
    for r in range(0, N, slab):
        nat.gram_host(X[r:r + slab], Y, out_s, s, 2.0, 0.25)
t = timed(slabs); print(f"4 slabs of {slab} rows: {t*1e3:.1f} ms  ({4.0*N*M/t/1e9:.1f} GB/s, {N*M/t:.3e} evals/s)", flush=True)
dev = torch.empty((slab, M), device="cuda")
def raw():
    out_s.copy_(dev, non_blocking=True); torch.cuda.synchronize()
t = timed(raw); print(f"raw pinned D2H of one slab: {t*1e3:.1f} ms ({4.0*slab*M/t/1e9:.1f} GB/s)", flush=True)
del dev
out = torch.empty((N, M)).pin_memory()
t = timed(lambda: nat.gram_host(X, Y, out, s, 2.0, 0.25)); print(f"one call, 17.2 GB pinned: {t*1e3:.1f} ms  ({4.0*N*M/t/1e9:.1f} GB/s, {N*M/t:.3e} evals/s)", flush=True)

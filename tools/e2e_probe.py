import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
M, d = 65536, 64
Y = torch.randn(M, d).pin_memory()
s = float(1/np.sqrt(d))
for N in (4096, 16384, 65536):
    X = torch.randn(N, d).pin_memory()
    out = torch.empty((N, M), dtype=torch.float32).pin_memory()
    nat.gram_host(X, Y, out, s, 2.0, 0.25)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); nat.gram_host(X, Y, out, s, 2.0, 0.25); ts.append(time.perf_counter() - t0)
    b = N * M * 4
    print(f"N={N}: {min(ts)*1e3:.1f} ms for {b/1e9:.2f} GB -> {b/min(ts)/1e9:.1f} GB/s", flush=True)

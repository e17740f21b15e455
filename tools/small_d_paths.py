"""Direct differencing vs the tensor path (tf32 operands padded to 32 columns, zero k-steps skipped) for small d on a
large Gram: which one should JRK_PATH_AUTO take?  JSON lines."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
dev = torch.device("cuda:0")
def randn(seed, *shape):
    g = torch.Generator(device=dev).manual_seed(seed); return torch.randn(*shape, generator=g, device=dev)
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
N = M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = torch.empty((N, M), device=dev)
for d in (1, 2, 4, 8, 12, 16, 24, 32, 48):
    X, Y = randn(0, N, d), randn(1, M, d)
    s = float(0.70710695 / np.sqrt(d))
    for p in (2.0, 1.0):
        row = {"N": N, "d": d, "power": p}
        for name, path in (("direct", nat.PATH_DIRECT), ("tensor", nat.PATH_TF32X3)):
            row[name + "_ms"] = round(t(lambda: nat.gram(X, Y, s, p, 0.25, path=path, out=K)), 3)
        print(json.dumps(row), flush=True)

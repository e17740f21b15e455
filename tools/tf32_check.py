"""Quick numerical check of the tcgen05 3xTF32 Gram path against float64 (torch on the GPU, test tooling only)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")


def truth(X, Y, s, p, nc):
    d2 = torch.cdist(X.double(), Y.double()) ** 2
    return nc * torch.exp(-(s * d2.sqrt()) ** p)


for (N, M, d) in [(64, 128, 64), (300, 200, 64), (1000, 777, 64), (300, 200, 96), (300, 200, 128), (300, 200, 70), (4096, 4096, 64)]:
    g = torch.Generator(device=dev).manual_seed(N + d)
    X = torch.randn(N, d, generator=g, device=dev)
    Y = torch.randn(M, d, generator=g, device=dev)
    for p in (2.0, 1.0, 1.5):
        s, nc = float(1 / np.sqrt(d)), 0.25
        K = nat.gram(X, Y, s, p, nc, path=nat.PATH_TF32X3)
        torch.cuda.synchronize()
        T = truth(X, Y, s, p, nc)
        err = (K.double() - T).abs()
        lim = 1e-6 + 1e-5 * T.abs()
        print(f"N={N} M={M} d={d} p={p}: max err {err.max().item():.3e}  max err/lim {(err / lim).max().item():.3f}", flush=True)
    Ks = nat.gram(X, None, s, 1.0, nc, path=nat.PATH_TF32X3)
    T = truth(X, X, s, 1.0, nc)
    print("  self: symmetric", bool(torch.equal(Ks, Ks.T)), "diag err", (torch.diagonal(Ks) - nc).abs().max().item(),
          "max err/lim", ((Ks.double() - T).abs() / (1e-6 + 1e-5 * T.abs())).max().item(), flush=True)
print("ok")

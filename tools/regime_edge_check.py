"""Accuracy of the factored p = 2 regime (kmatw_tc2.cu / kmatw_tc.cu fast path) towards the edge of its window
|c| (max|x|^2 + max|y|^2) <= TC_FAST_BOUND, on random points PLUS adversarial ones: pairs of large-norm points in the same
direction (the largest exponents S the regime admits, and K = nconst there) and in opposite directions (the smallest).
Usage: python tools/regime_edge_check.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
N, M, d, m = 16384, 65536, 16, 16
nc = 0.25
for target in (8.0, 11.0, 11.9, 13.0):      # (13: outside the window, both columns are the generic epilogue)
    X = torch.randn(N, d, generator=g, device=dev)
    Y = torch.randn(M, d, generator=g, device=dev)
    s = float(0.70710695 / np.sqrt(d))
    c = s * s * 1.4426950408889634
    # adversarial rows: 64 directions u; x = r u, y = r u and y' = -r u with c (r^2 + r^2) = target
    r = float(np.sqrt(target / (2 * c)))
    U = torch.nn.functional.normalize(torch.randn(64, d, generator=g, device=dev), dim=1)
    # shrink the random cloud so that the adversarial points carry the maxima
    X *= 0.6
    Y *= 0.6
    X[:64] = r * U
    Y[:64] = r * U
    Y[64:128] = -r * U
    bnd = c * float((X * X).sum(1).max() + (Y * Y).sum(1).max())
    for positive in (True, False):
        W = torch.rand(M, m, generator=g, device=dev) if positive else torch.randn(M, m, generator=g, device=dev)
        out = {}
        for fast in (1, 0):
            nat.set_option(nat.OPT_TC_FAST, fast)
            A = nat.kmatw(X, Y, W, s, 2.0, nc)
            torch.cuda.synchronize()
            rows = torch.cat([torch.arange(0, 64, device=dev), torch.arange(64, N, (N - 64) // 64, device=dev)[:64]])
            Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** 2)
            T, bound = Kt @ W.double(), Kt @ W.double().abs()
            e = (A[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)
            out[fast] = (float(e[:64].max()), float(e[64:].max()))
        nat.set_option(nat.OPT_TC_FAST, 1)
        print(f"bound {bnd:5.2f} positive W {positive!s:5}: factored err/lim adversarial rows {out[1][0]:.3f}, random rows {out[1][1]:.3f}"
              f" | generic {out[0][0]:.3f}, {out[0][1]:.3f}", flush=True)

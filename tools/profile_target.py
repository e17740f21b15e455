"""Small fixed launches of the hot kernels for `ncu --set full` captures (one launch each, after one warm-up).
Usage: python tools/profile_target.py [gram64|gram16|gram8|kmatw|kmatw_tc|all]"""
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


def gram(d, n=16384):
    X, Y = randn(0, n, d), randn(1, n, d)
    K = torch.empty((n, n), device=dev, dtype=torch.float32)
    s = float(0.70710695 / np.sqrt(d))
    for _ in range(2):
        nat.gram(X, Y, s, 2.0, 0.25, out=K)
    torch.cuda.synchronize()


def kmatw():
    N, M, d, m = 32768, 262144, 16, 16
    X, Y, W = randn(0, N, d), randn(1, M, d), randn(2, M, m)
    for _ in range(2):
        nat.kmatw(X, Y, W, 0.1767767, 2.0, 0.25)
    torch.cuda.synchronize()


def kmatw_tc():
    N, M, d, m = 148 * 256, 131072, 16, 16
    X, Y, W = randn(0, N, d), randn(1, M, d), randn(2, M, m)
    for _ in range(2):
        nat.kmatw(X, Y, W, 0.1767767, 2.0, 0.25)
    torch.cuda.synchronize()


def kmatw_tc_d64():
    N, M, d, m = 148 * 256, 131072, 64, 16
    X, Y, W = randn(0, N, d), randn(1, M, d), randn(2, M, m)
    for _ in range(2):
        nat.kmatw(X, Y, W, 0.0883883, 2.0, 0.25)
    torch.cuda.synchronize()


def kgrad_tc():
    N, M, d, m = 148 * 256, 131072, 16, 16
    X, Y, W, G = randn(0, N, d), randn(1, M, d), randn(2, M, m), randn(3, N, m)
    for _ in range(2):
        nat.kmatw_grad(X, Y, W, G, 0.1767767, 2.0, 0.25)
    torch.cuda.synchronize()


def gram_grad_tc():
    N, M, d = 148 * 256, 65536, 16
    X, Y, Gb = randn(0, N, d), randn(1, M, d), randn(2, N, M)
    for _ in range(2):
        nat.gram_grad(X, Y, Gb, 0.1767767, 2.0, 0.25)
    torch.cuda.synchronize()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gram64", "all"):
        gram(64)
    if what in ("gram16", "all"):
        gram(16)
    if what in ("gram8", "all"):
        gram(8)
    if what in ("kmatw", "all"):
        kmatw()
    if what == "kmatw_tc":
        kmatw_tc()
    if what == "kmatw_tc_d64":
        kmatw_tc_d64()
    if what == "kgrad_tc":
        kgrad_tc()
    if what == "gram_grad_tc":
        gram_grad_tc()
    print("done", nat.launch_count())

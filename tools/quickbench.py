"""Kernel-only timings of the C-ABI entry points (CUDA events, resident inputs).
Usage: python tools/quickbench.py [gram|kmatw|groupsum|grad|siblings|all]   -> JSON lines on stdout."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6542.1e9
FMA = 3.69e13  # measured FFMA2 lanes/s (profiles/microbench)


def randn(seed, *shape):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), float(np.median(ts))


def kp(power, d):
    s = 1.0 / np.sqrt(d)
    return float(s), float(power), 0.25


def bench_gram(only_tf32=False):
    for (N, M, d, path) in [(32768, 32768, 64, 2), (65536, 65536, 64, 2), (65536, 65536, 96, 2), (65536, 65536, 128, 2)] if only_tf32 else [(1024, 1024, 8, 1), (32768, 32768, 8, 1), (32768, 32768, 16, 1), (32768, 32768, 32, 1),
                            (32768, 32768, 64, 1), (32768, 32768, 64, 2), (65536, 65536, 64, 2), (65536, 65536, 128, 2)]:
        X, Y = randn(0, N, d), randn(1, M, d)
        K = torch.empty((N, M), device=dev, dtype=torch.float32)
        for power in (2.0, 1.0, 1.5):
            s, p, nc = kp(power, d)
            best, med = timeit(lambda: nat.gram(X, Y, s, p, nc, path=path, out=K))
            ev = N * M / best
            print(json.dumps({"op": "gram", "N": N, "M": M, "d": d, "path": path, "power": power, "ms": best * 1e3,
                              "ms_med": med * 1e3, "evals_per_s": ev, "hbm_frac": 4 * ev / HBM,
                              "fma_frac_direct": ev * (2 * d + 2) / FMA}), flush=True)
        del K


def bench_kmatw():
    for (N, M, d, m) in [(131072, 1 << 20, 16, 16), (131072, 1 << 20, 16, 1), (131072, 1 << 20, 8, 16),
                         (131072, 1 << 20, 32, 16), (65536, 1 << 20, 64, 16), (131072, 1 << 20, 16, 32)]:
        # (s = 1 / sqrt(d): d = 8 falls outside the factored regime -- the generic epilogue; d = 64 inside it runs
        # kmatw_tc2_kernel<4, 4> for power = 2)
        X, Y, W = randn(0, N, d), randn(1, M, d), randn(2, M, m)
        for power in (2.0, 1.0, 1.5):
            s, p, nc = kp(power, d)
            for rr in (0, 2):
                if rr and (power != 2.0):
                    continue
                best, med = timeit(lambda: nat.kmatw(X, Y, W, s, p, nc, row_reduce=rr), reps=3, warm=1)
                ev = N * M / best
                print(json.dumps({"op": "kmatw", "N": N, "M": M, "d": d, "m": m, "power": power, "row_reduce": rr,
                                  "ms": best * 1e3, "ms_med": med * 1e3, "evals_per_s": ev,
                                  "fma_frac": ev * (2 * d + m + 2) / FMA}), flush=True)


def bench_groupsum():
    g, groups, d = 4096, 32, 32
    X = randn(0, g * groups, d)
    s, p, nc = kp(2.0, d)
    best, med = timeit(lambda: nat.gram_groupsum(X, None, s, p, nc, g, g, True), reps=3, warm=1)
    ev = (g * groups) ** 2 / best
    print(json.dumps({"op": "groupsum", "N": g * groups, "d": d, "ms": best * 1e3, "evals_per_s": ev,
                      "fma_frac": ev * (2 * d + 2) / FMA}), flush=True)


def bench_grad():
    """Fused backward kernels: pairs/s and the FP32-pipe fraction (3d + m + 8 lane-ops per pair, kgrad.cu)."""
    for (N, M, d, m) in [(131072, 131072, 16, 16), (151552, 1048576, 16, 16), (131072, 131072, 8, 1), (65536, 131072, 32, 16), (32768, 131072, 64, 16)]:
        X, Y, W, G = randn(0, N, d), randn(1, M, d), randn(2, M, m), randn(3, N, m)
        for power in (2.0, 1.0, 1.5):
            s, p, nc = kp(power, d)
            best, med = timeit(lambda: nat.kmatw_grad(X, Y, W, G, s, p, nc), reps=3, warm=1)
            ev = N * M / best
            tc = power == 2.0 and d <= 16 and m <= 16            # kgrad_tc.cu: d + 2 FP32 lane-ops per pair
            print(json.dumps({"op": "kmatw_grad", "N": N, "M": M, "d": d, "m": m, "power": power, "ms": best * 1e3,
                              "pairs_per_s": ev, "kernel": "kgrad_tc (tcgen05 + FP32 accumulation)" if tc else "kgrad (FP32 pipe)",
                              "fma_frac": ev * ((d + 2) if tc else (3 * d + m + 8)) / FMA}), flush=True)
    for (N, M, d) in [(32768, 32768, 16), (32768, 32768, 64)]:
        X, Y, Gb = randn(0, N, d), randn(1, M, d), randn(4, N, M)
        s, p, nc = kp(2.0, d)
        best, med = timeit(lambda: nat.gram_grad(X, Y, Gb, s, p, nc), reps=3, warm=1)
        ev = N * M / best
        print(json.dumps({"op": "gram_grad", "N": N, "M": M, "d": d, "power": 2.0, "ms": best * 1e3, "pairs_per_s": ev,
                          "hbm_read_frac": 4 * ev / HBM, "fma_frac": ev * (3 * d + 8) / FMA}), flush=True)


def bench_siblings():
    """PeriodicKernel / ThreshSpikeKernel Grams on the direct kernel (HBM-write roofline, 4 B per evaluation)."""
    for (N, M, d) in [(32768, 32768, 1), (32768, 32768, 8)]:
        X, Y = randn(0, N, d), randn(1, M, d)
        for name, spec in (("periodic", nat.KernelSpec(nat.KIND_PERIODIC, (0.5, 1.0))),
                           ("threshspike", nat.KernelSpec(nat.KIND_THRESHSPIKE, (1.0, 1.0, 1.0, 0.1, 1.5)))):
            K = torch.empty((N, M), device=dev, dtype=torch.float32)
            best, med = timeit(lambda: nat.gram_spec(X, Y, spec, out=K))
            ev = N * M / best
            print(json.dumps({"op": "gram_" + name, "N": N, "M": M, "d": d, "ms": best * 1e3, "evals_per_s": ev,
                              "hbm_frac": 4 * ev / HBM}), flush=True)
            del K


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gram", "all"):
        bench_gram()
    if what == "tf32":
        bench_gram(True)
    if what in ("kmatw", "all"):
        bench_kmatw()
    if what in ("groupsum", "all"):
        bench_groupsum()
    if what in ("grad", "all"):
        bench_grad()
    if what in ("siblings", "all"):
        bench_siblings()

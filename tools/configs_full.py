"""BASELINE.json configs 4 and 5 at FULL size through the reference-shaped API (one-off record; JSON lines).
cfg4: FiniteVec(k, X[4,194,304 x 32], [BalancedRed(4096, average=True)]).inner()  -> 1024 x 1024 (1.76e13 kernel evals)
cfg5: conditional density operator: Gram build N = 32768 feeding the library solve, then the operator applied to 2^20
      test points (K(X_test, X) @ W, m = 16)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402
from jaxrk_b200.kern import GenGaussKernel  # noqa: E402
from jaxrk_b200.reduce import BalancedRed, LinearReduce  # noqa: E402
from jaxrk_b200.rkhs import FiniteVec, inner, Cov_inv, CovOp  # noqa: E402

dev = torch.device("cuda:0")


def randn(seed, *shape):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) * 1e-3


what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("cfg4", "all"):
    groups, g, d = 1024, 4096, 32
    X = randn(0, groups * g, d)
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    v = FiniteVec(k, X, [BalancedRed(g, average=True)])
    n0 = nat.launch_count()
    G, t = timed(lambda: v.inner())
    evals = float(groups * g) ** 2
    print(json.dumps({"config": "cfg4", "points": groups * g, "d": d, "out": list(G.shape), "seconds": t, "evals_per_s": evals / t,
                      "launches": nat.launch_count() - n0, "symmetric_to": float((G - G.T).abs().max() / G.abs().max()),
                      "fp32_pipe_frac": evals / t * (2 * d + 3) / 3.69e13}), flush=True)
    Y = randn(1, 4096, d)
    E, t = timed(lambda: inner(v, FiniteVec(k, Y).mean()))
    print(json.dumps({"config": "cfg4-mean-embedding", "out": list(E.shape), "seconds": t, "evals_per_s": groups * g * 4096 / t}), flush=True)
    del X, v, G
if what in ("cfg5", "all"):
    N, T, d, m = 32768, 1 << 20, 16, 16
    X = randn(2, N, d)
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    inner(FiniteVec(k, X))                     # warm-up (module load, allocator); a fresh FiniteVec is not cached
    inp = FiniteVec(k, X)
    Gm, t_gram = timed(lambda: inner(inp))
    print(json.dumps({"config": "cfg5-gram", "N": N, "d": d, "seconds": t_gram, "evals_per_s": N * N / t_gram,
                      "hbm_frac_of_6542": 4.0 * N * N / t_gram / 6542.1e9}), flush=True)
    op, t_solve = timed(lambda: Cov_inv(CovOp(inp), regul=0.1))   # library path (torch.linalg.inv): not the hot path
    print(json.dumps({"config": "cfg5-solve (library path, not hot)", "seconds": t_solve}), flush=True)
    Xt = randn(3, T, d)
    W = randn(4, m, N)
    inner(FiniteVec(k, Xt[:16384]), FiniteVec(k, X, [LinearReduce(W)]))     # warm-up
    out, t_apply = timed(lambda: inner(FiniteVec(k, Xt), FiniteVec(k, X, [LinearReduce(W)])))
    print(json.dumps({"config": "cfg5-apply", "T": T, "N": N, "m": m, "out": list(out.shape), "seconds": t_apply,
                      "evals_per_s": T * N / t_apply, "mufu_frac": T * N / t_apply / 4.637e12}), flush=True)
    # the same through the operator objects: Cdo(inp, outp, ref, regul) @ X_test with a 256-point reference grid
    from jaxrk_b200.rkhs.operator import Cdo  # noqa: E402
    R = 256
    Yv = torch.sin(X[:, :1]) + 0.1 * randn(5, N, 1)
    ky = GenGaussKernel.make_gauss(0.3)
    cdo, t_build = timed(lambda: Cdo(FiniteVec(k, X), FiniteVec(ky, Yv), FiniteVec(ky, torch.linspace(-2, 2, R, device=dev)[:, None]), regul=1.0))
    _, t_map = timed(lambda: cdo.inp_feat.reduce[-1].linear_map)        # Cmo's N x N coefficient map: K(X, X) @ matr^T, m = N
    res = cdo @ Xt
    n0 = nat.launch_count()
    coeff, t_cdo = timed(lambda: res.reduce[-1].linear_map)
    print(json.dumps({"config": "cfg5-Cdo through the operator objects", "N": N, "T": T, "R": R, "build_seconds (2 Grams + library inverses)": t_build,
                      "Cmo map K(X,X)@matr^T (m = N = 32768) seconds": t_map, "Cmo map TFLOP/s equivalent": 2.0 * N * N * N / t_map / 1e12,
                      "apply_seconds (K(X_test, X) @ A^T, m = 256)": t_cdo, "apply_evals_per_s": T * N / t_cdo,
                      "apply_TFLOP/s equivalent": 2.0 * T * N * R / t_cdo / 1e12, "launches": nat.launch_count() - n0}), flush=True)
    mv, t_mv = timed(lambda: res.get_mean_var())
    print(json.dumps({"config": "cfg5-get_mean_var on the applied operator (skinny, map not materialised twice)", "seconds": t_mv}), flush=True)

"""Tensor-core K@W backward (csrc/kgrad_tc.cu) against float64 on sampled rows, and its speed against the FP32-pipe
kernel (csrc/kgrad.cu).  Usage: python tools/kgrad_tc_check.py [acc|time|all]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from jaxrk_b200 import _native as nat  # noqa: E402
from oracle import jaxrk_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
f32 = np.float32
RT, AT = 2e-5, 2e-6


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=f32)).to(dev)


def check(N, M, d, m, ls, seed=0, positive=False):
    rng = np.random.RandomState(seed)
    X, Y = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32)
    W, G = rng.randn(M, m).astype(f32), rng.randn(N, m).astype(f32)
    if positive:
        W, G = np.abs(W), np.abs(G)
    s = orc.simple_scaler_s(orc.gauss_length_scale(ls))
    scale, p = float(s[0, 0]), 2.0
    nconst = float(orc.gengauss_nconst_ref32(s, p).reshape(-1)[0])
    res = {}
    for tc in (1, 0):
        with nat.option(nat.OPT_KMATW_TC, tc):
            n0 = nat.launch_count()
            gx, gs = nat.kmatw_grad(cu(X), cu(Y), cu(W), cu(G), scale, p, nconst)
            gy, _ = nat.kmatw_grad(cu(Y), cu(X), cu(G), cu(W), scale, p, nconst, want_scale=False)
            torch.cuda.synchronize()
            res[tc] = (gx.cpu().numpy(), gs.cpu().numpy(), gy.cpu().numpy(), nat.launch_count() - n0)
    rows = np.r_[0:24, N - 24:N]
    A = G[rows].astype(np.float64) @ W.astype(np.float64).T
    Aabs = np.abs(G[rows]).astype(np.float64) @ np.abs(W).astype(np.float64).T
    gX, _, gsum, _ = orc.gengauss_vjp_truth64(X[rows], Y, A, s, p, nconst=nconst)
    _, _, _, (mx, _, ms) = orc.gengauss_vjp_truth64(X[rows], Y, Aabs, s, p, nconst=nconst)
    cols = np.r_[0:24, M - 24:M]
    A2 = G.astype(np.float64) @ W[cols].astype(np.float64).T
    A2abs = np.abs(G).astype(np.float64) @ np.abs(W[cols]).astype(np.float64).T
    _, gY, _, _ = orc.gengauss_vjp_truth64(X, Y[cols], A2, s, p, nconst=nconst)
    _, _, _, (_, my, _) = orc.gengauss_vjp_truth64(X, Y[cols], A2abs, s, p, nconst=nconst)
    out = {"N": N, "M": M, "d": d, "m": m, "ls": ls, "positive": positive}
    for tc in (1, 0):
        gx, gs, gy, nl = res[tc]
        ex = np.abs(gx[rows] - gX) / (AT + RT * mx)
        ey = np.abs(gy[cols] - gY) / (AT + RT * my)
        es = abs(float(gs[rows].astype(np.float64).sum()) - gsum) / (AT + RT * ms)
        out["tc" if tc else "fp32"] = {"dX": float(ex.max()), "dY": float(ey.max()), "ds": float(es), "launches": nl}
    print(out, flush=True)
    return out


def check_gram(N, M, d, ls, seed=0, positive=False):
    rng = np.random.RandomState(seed)
    X, Y = rng.randn(N, d).astype(f32), rng.randn(M, d).astype(f32)
    Gb = rng.randn(N, M).astype(f32)
    if positive:
        Gb = np.abs(Gb)
    s = orc.simple_scaler_s(orc.gauss_length_scale(ls))
    scale, p = float(s[0, 0]), 2.0
    nconst = float(orc.gengauss_nconst_ref32(s, p).reshape(-1)[0])
    rows = np.r_[0:24, 130:140, N - 24:N]
    gX, _, gsum, (mx, _, ms) = orc.gengauss_vjp_truth64(X[rows], Y, Gb[rows].astype(np.float64), s, p, nconst=nconst)
    _, _, _, (mxa, _, msa) = orc.gengauss_vjp_truth64(X[rows], Y, np.abs(Gb[rows]).astype(np.float64), s, p, nconst=nconst)
    out = {"gram_grad": True, "N": N, "M": M, "d": d, "ls": ls, "positive": positive}
    for tc, tr in ((1, False), (1, True), (0, False), (0, True)):
        with nat.option(nat.OPT_KMATW_TC, tc):
            n0 = nat.launch_count()
            gx, gs = nat.gram_grad(cu(X), cu(Y), cu(np.ascontiguousarray(Gb.T)) if tr else cu(Gb), scale, p, nconst, transposed=tr)
            torch.cuda.synchronize()
            nl = nat.launch_count() - n0
        ex = np.abs(gx.cpu().numpy()[rows] - gX) / (AT + RT * mxa)
        es = abs(float(gs.cpu().numpy()[rows].astype(np.float64).sum()) - gsum) / (AT + RT * msa)
        out[("tc" if tc else "fp32") + ("_T" if tr else "")] = {"dX": float(ex.max()), "ds": float(es), "launches": nl}
    print(out, flush=True)


def timeit_gram(N, M, d):
    g = torch.Generator(device=dev).manual_seed(0)
    X, Y = torch.randn(N, d, generator=g, device=dev), torch.randn(M, d, generator=g, device=dev)
    Gb = torch.randn(N, M, generator=g, device=dev)
    scale = float(0.70710695 / np.sqrt(d))
    out = {"gram_grad": True, "N": N, "M": M, "d": d}
    GbT = Gb.t().contiguous()
    for tc, tr in ((1, False), (1, True), (0, False), (0, True)):
        with nat.option(nat.OPT_KMATW_TC, tc):
            cot = GbT if tr else Gb
            for _ in range(2):
                nat.gram_grad(X, Y, cot, scale, 2.0, 0.25, transposed=tr)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nat.gram_grad(X, Y, cot, scale, 2.0, 0.25, transposed=tr)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = min(ts)
            out[("tc" if tc else "fp32") + ("_T" if tr else "")] = {"ms": round(t, 3), "pairs_per_s": N * M / t * 1e3, "hbm_read_GBps": round(4 * N * M / t * 1e-6)}
    print(out, flush=True)


def timeit(N, M, d, m):
    g = torch.Generator(device=dev).manual_seed(0)
    X, Y = torch.randn(N, d, generator=g, device=dev), torch.randn(M, d, generator=g, device=dev)
    W, G = torch.randn(M, m, generator=g, device=dev), torch.randn(N, m, generator=g, device=dev)
    scale = float(0.70710695 / np.sqrt(d))
    out = {"N": N, "M": M, "d": d, "m": m}
    for tc in (1, 0):
        with nat.option(nat.OPT_KMATW_TC, tc):
            for _ in range(2):
                nat.kmatw_grad(X, Y, W, G, scale, 2.0, 0.25)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nat.kmatw_grad(X, Y, W, G, scale, 2.0, 0.25)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = min(ts)
            out["tc" if tc else "fp32"] = {"ms": t, "pairs_per_s": N * M / t * 1e3, "fp32_pipe_frac_18": N * M / t * 1e3 * 18 / 3.69e13}
    print(out, flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("acc", "all"):
        check(8192, 8192, 16, 16, 4.0)
        check(4096 + 77, 4096 + 13, 7, 5, 2.5, seed=1)
        check(8192, 4096, 16, 1, 4.0, seed=2, positive=True)
        check(4224, 6000, 3, 16, 1.7, seed=3)
        check(8192, 8192, 16, 16, 0.5, seed=4)            # outside the factored regime: the FP32 kernel must serve it
        check_gram(8192, 8192, 16, 4.0)
        check_gram(4096 + 300, 4096 + 13, 7, 2.5, seed=1, positive=True)
        check_gram(4224, 6000, 3, 1.7, seed=3)
    if what in ("time", "all"):
        timeit(131072, 131072, 16, 16)
        timeit(151552, 131072, 16, 16)
        timeit(37888, 1048576, 16, 16)
        timeit_gram(32768, 32768, 16)
        timeit_gram(37888, 131072, 16)

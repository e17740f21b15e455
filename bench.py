#!/usr/bin/env python
"""Headline benchmark of the JaxRK hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gram|kmatw] [--impl b200|reference]

One "step" is one pass of the hot path over one batch of synthetic N(0, I) inputs.

  workload gram  (default; BASELINE.json configs[1]): materialised GenGaussKernel Gram, N = M = 65536,
                 d = 64, fp32.  HBM-write-bound.  N GPUs: X is row-sharded, every rank owns 65536 rows
                 (weak scaling, no data-path collective; the Gram stays row-sharded).
  workload kmatw (configs[2]): fused, never-materialised K(X, Y) @ W, N = M = 2^20, d = 16, m = 16.
                 FP32-pipe-bound.  N GPUs: the 2^20 rows of X are sharded (strong scaling), Y / W replicated.
                 Reported under "also" of the default line, or as the main line with --workload kmatw.

`value`  kernel evaluations / s, whole job, inputs resident in HBM, CUDA events on the launch stream, an L2
         flush between timed iterations, max over ranks.
`e2e`    the same metric through the host-buffer C-ABI call (jrk_*_f32_host): pinned host inputs are copied
         to the device and the result is copied back inside the timed region.
`--impl reference` times the restated reference op sequence on the host (oracle/torch_ref.py: torch-CPU eager
         float32, one library call per reference op, all host cores; the reference itself is pure Python over
         jax/XLA, which is absent from the image) on a bounded row sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: N (rows per rank for gram / total rows for kmatw), M, d, m
    "gram": dict(N=65536, M=65536, d=64, m=0, config="configs[1]: GenGaussKernel Gram N=M=65536 d=64 fp32 materialised"),
    "kmatw": dict(N=1 << 20, M=1 << 20, d=16, m=16, config="configs[2]: fused K(X,Y)@W N=M=2^20 d=16 m=16, never materialised"),
}
# north_star's "Gram at N = M = 1M, d = 16": 4.4 TB cannot be materialised, so it is STREAMED: every
# GPU writes its rows block by block (16384 x 2^20 = 68.7 GB) into one reused buffer; evals/s = N M / time.
GRAM1M = dict(N=1 << 20, M=1 << 20, d=16, block=16384,
              config="north_star target: Gram N=M=2^20 d=16, streamed in 16384-row blocks into one reused 68.7 GB buffer")
POWER = 2.0  # SqExp member of the GenGauss family (make_gauss); other shapes are covered by tools/quickbench.py


def kernel_params(d):
    """make_gauss(length_scale = sqrt(d)) -- SURVEY.md §8d synthetic-input convention."""
    from jaxrk_b200.kern import GenGaussKernel
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    scale, dim_scale, power, nconst = k.kernel_args()
    return scale, power, nconst


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)", float(j.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def fp32_pipe_peak():
    """FP32 lane-ops/s of the FMA pipe: measured FFMA2 rate (profiles/microbench/pipes.cu on B200)."""
    p = os.path.join(ROOT, "profiles", "microbench", "pipes_b200.json")
    if os.path.exists(p):
        return float(json.load(open(p))["ffma2_fma_per_s"]), "measured (profiles/microbench/pipes_b200.json)"
    return 148 * 128 * 1.965e9, "nominal 148 SM x 128 lanes x 1.965 GHz"


def mufu_peak():
    """MUFU ex2 results/s: measured on B200 (profiles/microbench/pipes.cu); the tensor-core K@W path issues one per pair."""
    p = os.path.join(ROOT, "profiles", "microbench", "pipes_b200.json")
    if os.path.exists(p):
        return float(json.load(open(p))["mufu_ex2_per_s"]), "measured (profiles/microbench/pipes_b200.json)"
    return 148 * 16 * 1.965e9, "nominal 148 SM x 16 MUFU lanes x 1.965 GHz"


def traffic_for(name):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(name)
    return None


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML), running during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "display_clock_setting": 0x100,
                 "applications_clocks_setting": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.nv:
            self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU restatement of the reference (bounded sample)
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(name, rows, seed_shift=0):
    """Time the restated reference op sequence (oracle/torch_ref.py: torch-CPU eager float32, one library
    call per reference op, all host cores) on `rows` rows x all M columns of the workload; returns
    (evals_per_s, seconds, threads)."""
    import torch
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count())
    w = WORKLOADS[name]
    M, d, m = w["M"], w["d"], w["m"]
    X = torch.from_numpy(np.random.RandomState(10 + seed_shift).randn(rows, d).astype(np.float32))
    Y = torch.from_numpy(np.random.RandomState(1).randn(M, d).astype(np.float32))
    scale, power, nconst = kernel_params(d)
    if name == "gram":
        t0 = time.perf_counter()
        out = torch_ref.gram(X, Y, scale, power, nconst, block=2048)
    else:
        W = torch.from_numpy(np.random.RandomState(2).randn(M, m).astype(np.float32))
        t0 = time.perf_counter()
        out = torch_ref.kmatw(X, Y, W, scale, power, nconst, block=256)
    dt = time.perf_counter() - t0
    assert bool(torch.isfinite(out).all())
    return rows * M / dt, dt, torch.get_num_threads()


def cpu_rows_for(name, target_s=12.0):
    ev, dt, thr = cpu_reference_sample(name, 2048 if name == "gram" else 256)
    rows = int(max(256, min(65536, ev * target_s / WORKLOADS[name]["M"])))
    return rows // 256 * 256


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    rows = cpu_rows_for(name, target_s=6.0)
    for i in range(args.warmup):
        cpu_reference_sample(name, max(64, rows // 8), i)
    tot_ev, tot_t, thr = 0.0, 0.0, 1
    for i in range(args.steps):
        ev, dt, thr = cpu_reference_sample(name, rows, i)
        tot_ev += rows * w["M"]
        tot_t += dt
    val = tot_ev / tot_t
    sample = f"{rows} rows x {w['M']} columns of the workload per step (d={w['d']}" + (f", m={w['m']}" if w["m"] else "") + ")"
    line = {"metric": "kernel_evals_per_s", "value": val, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0, I)", "impl": "reference",
            "config": {"workload": w["config"], "kernel": "GenGaussKernel.make_gauss(sqrt(d)) (p = 2)"},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "restated reference op sequence, torch-CPU eager float32 (oracle/torch_ref.py), all host threads; "
                    "the reference itself needs jax/jaxlib, absent from the image"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_workload(name, args, torch, nat, dist_ctx, with_cpu):
    rank, world, dev = dist_ctx
    w = WORKLOADS[name]
    M, d, m = w["M"], w["d"], w["m"]
    if name == "gram":
        n_local, n_total, scaling = w["N"], w["N"] * world, "weak"
    else:
        n_local, n_total, scaling = w["N"] // world, w["N"], "strong"
    scale, power, nconst = kernel_params(d)

    def randn(seed, *shape):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)

    X = randn(100 + rank, n_local, d)      # this rank's row block
    Y = randn(1, M, d)                     # replicated
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    if name == "gram":
        out = torch.empty((n_local, M), dtype=torch.float32, device=dev)
        ws_bytes = nat.lib().jrk_gram_workspace_bytes(n_local, M, d, nat.PATH_AUTO)
        step = lambda: nat.gram(X, Y, scale, power, nconst, out=out)  # noqa: E731
        alg_bytes = 4.0 * n_local * M + 4.0 * d * (n_local + M)
    else:
        W = randn(2, M, m)
        out = torch.empty((n_local, m), dtype=torch.float32, device=dev)
        step = lambda: nat.kmatw(X, Y, W, scale, power, nconst, out=out)  # noqa: E731
        alg_ops = float(n_local) * M * (2 * d + m + 2)  # FP32-pipe lane-ops (DESIGN.md "Rooflines")

    steps = args.steps if name == "gram" else max(1, min(args.steps, 5))
    warm = args.warmup if name == "gram" else max(1, min(args.warmup, 3))
    for _ in range(warm):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    torch.cuda.synchronize(dev)
    l0 = nat.launch_count()
    evs = []
    for _ in range(steps):
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize(dev)
    launches = nat.launch_count() - l0
    if world > 1:
        torch.distributed.barrier()
    clocks = sampler.stop()
    t_local = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    t = torch.tensor([t_local], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    t_step = float(t.item()) / steps
    value = float(n_total) * M / t_step

    hbm, hbm_src, _ = peaks()
    if name == "gram":
        achieved = alg_bytes / (t_local / steps) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic_for(name), "peak_source": hbm_src, "kernel_ms": t_local / steps * 1e3,
                "algorithmic_bytes_per_launch": alg_bytes}
    elif nat.lib().jrk_kmatw_uses_tensor_cores(n_local, M, d, m, 0, 0):
        # both contractions on tcgen05 (kmatw_tc.cu): what remains per pair is one MUFU ex2 and ~9 issue slots, so
        # the bound is the MUFU pipe; the FP32-pipe kernel's figure of merit is kept alongside for comparison
        pk, src = mufu_peak()
        achieved = float(n_local) * M / (t_local / steps)
        fpk, _ = fp32_pipe_peak()
        roof = {"bound": "mufu", "achieved": achieved / 1e12, "peak": pk / 1e12, "unit": "T ex2/s",
                "frac": achieved / pk, "traffic": traffic_for("kmatw_tc"), "peak_source": src,
                "kernel_ms": t_local / steps * 1e3, "mufu_per_eval": 1, "path": "tcgen05 3xTF32 (kmatw_tc.cu)",
                "fp32_pipe_equiv_frac": alg_ops / (t_local / steps) / fpk,
                "tflops_equiv": float(n_local) * M * (3 * d + 2 * m + 4) / (t_local / steps) / 1e12}
    else:
        pk, src = fp32_pipe_peak()
        achieved = alg_ops / (t_local / steps)
        roof = {"bound": "fp32_pipe", "achieved": achieved / 1e12, "peak": pk / 1e12, "unit": "T lane-op/s",
                "frac": achieved / pk, "traffic": traffic_for(name), "peak_source": src,
                "kernel_ms": t_local / steps * 1e3, "lane_ops_per_eval": 2 * d + m + 2, "path": "FP32 pipe (kmatw.cu)",
                "mufu_per_eval": 1, "tflops_equiv": float(n_local) * M * (3 * d + 2 * m + 4) / (t_local / steps) / 1e12}

    # ---- end to end through the host-buffer C-ABI entry point ---------------------------------
    Xh = X.cpu().pin_memory()
    Yh = Y.cpu().pin_memory()
    e2e_steps = max(1, min(steps, 3))
    if name == "gram":
        # the whole N_local x M result crosses PCIe every step, through a pinned host buffer of one 16384-row slab
        # (4.3 GB per rank instead of 17.2 GB: eight ranks share one host)
        slab = 16384
        outh = torch.empty((slab, M), dtype=torch.float32).pin_memory()
        n_slabs = (n_local + slab - 1) // slab
        h2d, d2h = 4 * d * (n_local + n_slabs * M), 4 * n_local * M

        def host_step():
            for r in range(0, n_local, slab):
                n = min(slab, n_local - r)
                nat.gram_host(Xh[r:r + n], Yh, outh[:n], scale, power, nconst, device=dev.index)
    else:
        Wh = W.cpu().pin_memory()
        outh = torch.empty((n_local, m), dtype=torch.float32).pin_memory()
        h2d, d2h = 4 * (d * (n_local + M) + M * m), 4 * n_local * m
        host_step = lambda: nat.kmatw_host(Xh, Yh, Wh, outh, scale, power, nconst, device=dev.index)  # noqa: E731
    del out
    torch.cuda.empty_cache()
    host_step()  # warm-up (page tables, allocator)
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
    e2e_val = float(n_total) * M * e2e_steps / float(te.item())
    # spot-check the e2e result against the resident one (same kernels)
    chk = float(outh[:4].abs().sum())
    assert np.isfinite(chk) and chk > 0

    res = {"value": value, "ms_per_step": t_step * 1e3, "steps": steps, "warmup": warm, "scaling": scaling,
           "roofline": roof, "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "steps": e2e_steps, "ms_per_step": float(te.item()) / e2e_steps * 1e3},
           "config": {"workload": w["config"], "kernel": "GenGaussKernel.make_gauss(sqrt(d)) (p = 2)",
                      "rows_per_gpu": n_local, "global_rows": n_total, "M": M, "d": d,
                      "l2": "256 MiB L2 flush between timed iterations",
                      "sharding": "X row-blocks per rank, Y" + ("/W" if m else "") + " replicated, no data-path collective"}}
    if m:
        res["config"]["m"] = m
    if with_cpu and rank == 0 and world == 1:
        rows = cpu_rows_for(name)
        ev, dt, thr = cpu_reference_sample(name, rows)
        res["cpu_baseline"] = {"value": ev, "unit": "evals/s", "cores": thr, "kind": "port",
                               "sample": f"{rows} rows x {M} columns of the same workload, {dt:.1f} s, oracle/torch_ref.py "
                                         "(reference op sequence, torch-CPU eager)"}
    return res


def run_gram1m(args, torch, nat, dist_ctx):
    """Streamed Gram N = M = 2^20, d = 16 (kernel-only; there is nowhere to put 4.4 TB, so no e2e leg)."""
    rank, world, dev = dist_ctx
    N, M, d, B = GRAM1M["N"], GRAM1M["M"], GRAM1M["d"], GRAM1M["block"]
    n_local = N // world
    scale, power, nconst = kernel_params(d)
    g = torch.Generator(device=dev).manual_seed(300 + rank)
    X = torch.randn(n_local, d, generator=g, device=dev, dtype=torch.float32)
    Y = torch.randn(M, d, generator=torch.Generator(device=dev).manual_seed(1), device=dev, dtype=torch.float32)
    out = torch.empty((B, M), dtype=torch.float32, device=dev)

    def step():
        for r in range(0, n_local, B):
            nat.gram(X[r:r + B], Y, scale, power, nconst, out=out[:min(B, n_local - r)])

    nat.gram(X[:B], Y, scale, power, nconst, out=out)      # warm-up (one block)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize(dev)
    launches = nat.launch_count() - l0
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    t_local = float(t.item())
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    hbm, hbm_src, _ = peaks()
    achieved = 4.0 * n_local * M / t_local / 1e9
    return {"value": float(N) * M / float(t.item()), "ms_per_step": float(t.item()) * 1e3, "steps": 1, "scaling": "strong",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": None, "peak_source": hbm_src + " (sustained for seconds: the power cap applies)",
                         "algorithmic_bytes_per_step": 4.0 * n_local * M},
            "clocks": clocks, "gpu_launches": int(launches),
            "config": {"workload": GRAM1M["config"], "rows_per_gpu": n_local, "M": M, "d": d, "block_rows": B,
                       "l2": "every block (68.7 GB) is far larger than L2"}}


def run_b200(args):
    import torch
    from jaxrk_b200 import _native as nat
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    nat.lib()
    ctx = (rank, world, dev)
    main = run_workload(args.workload, args, torch, nat, ctx, with_cpu=True)
    also = {}
    if args.workload == "gram" and not args.no_also:
        try:
            also["kmatw_cfg3"] = run_workload("kmatw", args, torch, nat, ctx, with_cpu=(world == 1))
        except Exception as e:  # noqa: BLE001  (the headline line must still print)
            also["kmatw_cfg3"] = {"error": repr(e)}
        try:
            torch.cuda.empty_cache()
            also["gram_1m_d16_streamed"] = run_gram1m(args, torch, nat, ctx)
        except Exception as e:  # noqa: BLE001
            also["gram_1m_d16_streamed"] = {"error": repr(e)}
    if rank == 0:
        line = {"metric": "kernel_evals_per_s", "value": main["value"], "unit": "evals/s", "n_gpus": world,
                "steps": main["steps"], "warmup": main["warmup"], "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": main["scaling"], "vs_baseline": None, "dtype": "f32",
                "data": "synthetic N(0, I), generated on device with fixed seeds", "config": main["config"],
                "roofline": main["roofline"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
                "clocks": main["clocks"]}
        if "cpu_baseline" in main:
            line["cpu_baseline"] = main["cpu_baseline"]
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="gram")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary (kmatw) workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

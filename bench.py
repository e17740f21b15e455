#!/usr/bin/env python
"""Headline benchmark of the JaxRK hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kmatw|gram] [--impl b200|reference]

One "step" is one pass of the hot path over one batch of synthetic N(0, I) inputs.

  workload kmatw (default; BASELINE.json configs[2], the configuration north_star's target is quoted on): fused,
                 never-materialised K(X, Y) @ W, N = M = 2^20, d = 16, m = 16, fp32.  Both contractions on tcgen05,
                 bound by the MUFU pipe (one ex2 per pair).  N GPUs: the 2^20 rows of X are split over the ranks
                 (STRONG scaling), Y / W replicated, no data-path collective (the result stays row-sharded).
  workload gram  (configs[1]): materialised GenGaussKernel Gram, N = M = 65536 per rank, d = 64, HBM-write-bound
                 (weak scaling).  Reported under "also" of the default line, or as the main line with --workload gram.
  also           kmatw_cfg3_mean: the same K@W with `Mean` over rows (out 1 x 16): local sums + ONE all_reduce (NCCL)
                 inside the CUDA-event window, its latency reported separately;  gram_1m_d16_streamed: north_star's
                 "Gram at N = M = 1M, d = 16", streamed in row blocks.

`value`  kernel evaluations / s, whole job, inputs resident in HBM, CUDA events on the launch stream, an L2 flush
         between timed iterations, max over ranks.
`e2e`    the same metric through the host-buffer C-ABI call (jrk_*_f32_host): ALL inputs are copied from pinned host
         memory to the device and the result is copied back inside the timed region; `e2e.plan` is the same with the
         Y / W side kept on the device by a plan (jrk_kmatw_plan_apply_host: X in, result out).
`--impl reference` times the restated reference op sequence on the host (oracle/torch_ref.py: torch-CPU eager float32,
         one library call per reference op, all host cores; the reference itself is pure Python over jax/XLA, which is
         absent from the image) on a bounded row sample of the same workload, with the same `config`.
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
    # name: N (total rows for kmatw / rows per rank for gram), M, d, m
    "kmatw": dict(N=1 << 20, M=1 << 20, d=16, m=16, scaling="strong",
                  config="configs[2]: fused K(X,Y)@W N=M=2^20 d=16 m=16, never materialised"),
    "gram": dict(N=65536, M=65536, d=64, m=0, scaling="weak",
                 config="configs[1]: GenGaussKernel Gram N=M=65536 d=64 fp32 materialised"),
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


def workload_config(name, world):
    """The `config` of the JSON line: ONE function for both arms, so that the driver sees the same workload."""
    w = WORKLOADS[name]
    n_local = w["N"] // world if w["scaling"] == "strong" else w["N"]
    cfg = {"workload": w["config"], "kernel": "GenGaussKernel.make_gauss(sqrt(d)) (p = 2)", "rows_per_gpu": n_local,
           "global_rows": n_local * world, "M": w["M"], "d": w["d"],
           "sharding": "X row-blocks per rank, Y" + ("/W" if w["m"] else "") + " replicated, no data-path collective",
           "l2": "256 MiB L2 flush between timed iterations"}
    if w["m"]:
        cfg["m"] = w["m"]
    return cfg


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
            "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0, I)", "impl": "reference",
            "config": workload_config(name, max(1, args.gpus)),
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "PORT, not jaxlib: the restated reference op sequence in torch-CPU eager float32 (oracle/torch_ref.py), all "
                    "host threads, a bounded row sample of the same workload per step (a rate, so independent of the sample "
                    "size); the reference itself needs jax/jaxlib, absent from the image"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def _timed_steps(torch, dev, world, steps, flush, step, nat):
    """K timed iterations (CUDA events on the launch stream, L2 flush between them, barrier on both sides); returns
    (seconds of this rank summed over the steps, launches, clocks)."""
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
    return sum(a.elapsed_time(b) for a, b in evs) * 1e-3, launches, clocks


def _max_over_ranks(torch, dev, world, seconds):
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def run_workload(name, args, torch, nat, dist_ctx, with_cpu):
    rank, world, dev = dist_ctx
    w = WORKLOADS[name]
    M, d, m = w["M"], w["d"], w["m"]
    scaling = w["scaling"]
    n_local = w["N"] // world if scaling == "strong" else w["N"]
    n_total = n_local * world
    scale, power, nconst = kernel_params(d)

    def randn(seed, *shape):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)

    X = randn(100 + rank, n_local, d)      # this rank's row block
    Y = randn(1, M, d)                     # replicated
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    if name == "gram":
        out = torch.empty((n_local, M), dtype=torch.float32, device=dev)
        step = lambda: nat.gram(X, Y, scale, power, nconst, out=out)  # noqa: E731
        alg_bytes = 4.0 * n_local * M + 4.0 * d * (n_local + M)
    else:
        W = randn(2, M, m)
        out = torch.empty((n_local, m), dtype=torch.float32, device=dev)
        step = lambda: nat.kmatw(X, Y, W, scale, power, nconst, out=out)  # noqa: E731
        alg_ops = float(n_local) * M * (2 * d + m + 2)  # FP32-pipe lane-ops (DESIGN.md "Rooflines")

    steps, warm = max(1, args.steps), max(3, args.warmup)
    for _ in range(warm):
        step()
    t_local, launches, clocks = _timed_steps(torch, dev, world, steps, flush, step, nat)
    t_step = _max_over_ranks(torch, dev, world, t_local) / steps
    value = float(n_total) * M / t_step

    hbm, hbm_src, _ = peaks()
    if name == "gram":
        achieved = alg_bytes / (t_local / steps) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic_for(name), "peak_source": hbm_src, "kernel_ms": t_local / steps * 1e3,
                "algorithmic_bytes_per_launch": alg_bytes}
    elif nat.lib().jrk_kmatw_uses_tensor_cores(n_local, M, d, m, 0, 0):
        # both contractions on tcgen05 (kmatw_tc2.cu): what remains per pair is one MUFU ex2, so the bound is the MUFU
        # pipe; the FP32-pipe kernel's figure of merit is kept alongside for comparison
        pk, src = mufu_peak()
        achieved = float(n_local) * M / (t_local / steps)
        fpk, _ = fp32_pipe_peak()
        roof = {"bound": "mufu", "achieved": achieved / 1e12, "peak": pk / 1e12, "unit": "T ex2/s",
                "frac": achieved / pk, "traffic": traffic_for("kmatw_tc2"), "peak_source": src,
                "kernel_ms": t_local / steps * 1e3, "mufu_per_eval": 1,
                "path": "tcgen05, fp16 hi/lo operands for both GEMMs (kmatw_tc2.cu)",
                "note": "the step also contains ~1 ms of pre-pass kernels (row norms, tile images); `traffic` = DRAM bytes "
                        "of the main kernel per launch from ncu (algorithmic: 4 (N + M) (d + m) bytes)",
                "algorithmic_bytes_per_launch": 4.0 * (n_local + M) * (d + m),
                "fp32_pipe_equiv_frac": alg_ops / (t_local / steps) / fpk,
                "tflops_equiv": float(n_local) * M * (3 * d + 2 * m + 4) / (t_local / steps) / 1e12}
    else:
        pk, src = fp32_pipe_peak()
        achieved = alg_ops / (t_local / steps)
        roof = {"bound": "fp32_pipe", "achieved": achieved / 1e12, "peak": pk / 1e12, "unit": "T lane-op/s",
                "frac": achieved / pk, "traffic": traffic_for(name), "peak_source": src,
                "kernel_ms": t_local / steps * 1e3, "lane_ops_per_eval": 2 * d + m + 2, "path": "FP32 pipe (kmatw.cu)",
                "mufu_per_eval": 1, "tflops_equiv": float(n_local) * M * (3 * d + 2 * m + 4) / (t_local / steps) / 1e12}

    # ---- end to end through the host-buffer C-ABI entry points ---------------------------------
    Xh = X.cpu().pin_memory()
    Yh = Y.cpu().pin_memory()
    # K@W: as many end-to-end steps as timed steps (0.34 s each; one host hiccup in three steps moved the mean by 13 %);
    # the materialised Gram moves 17 GB per step over PCIe: three
    e2e_steps = max(1, min(steps, 3 if name == "gram" else 10))
    plan_e2e = None
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

    each = {}

    def time_host(fn, key="host"):
        fn()  # warm-up (page tables, allocator)
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        marks = []
        for _ in range(e2e_steps):
            fn()
            marks.append(time.perf_counter())
        each[key] = [round((b - a) * 1e3, 2) for a, b in zip([t0] + marks[:-1], marks)]      # this rank's steps, ms
        return _max_over_ranks(torch, dev, world, time.perf_counter() - t0) / e2e_steps

    te = time_host(host_step)
    e2e_val = float(n_total) * M / te
    chk = float(outh[:4].abs().sum())      # spot-check the e2e result (same kernels as the resident run)
    assert np.isfinite(chk) and chk > 0
    if name == "kmatw":
        # the operator's fixed side (Y, W) kept on the device by a plan: a step moves X in and the result out
        plan = nat.KmatwPlan(Yh, Wh, scale, power, nconst, device=dev.index)
        ref4 = outh[:4].clone()
        tp = time_host(lambda: plan.apply_host(Xh, outh), "plan")
        assert bool(torch.equal(ref4, outh[:4])), "plan and one-shot results must agree bit for bit"
        plan.close()
        plan_e2e = {"value": float(n_total) * M / tp, "unit": "evals/s", "ms_per_step": tp * 1e3,
                    "h2d_bytes_per_step": int(4 * d * n_local), "d2h_bytes_per_step": int(d2h),
                    "ms_each_step_rank0": each.get("plan"),
                    "what": "jrk_kmatw_plan_apply_host: Y / W and their tile images stay on the device"}

    res = {"value": value, "ms_per_step": t_step * 1e3, "steps": steps, "warmup": warm, "scaling": scaling,
           "roofline": roof, "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "steps": e2e_steps, "ms_per_step": te * 1e3, "ms_each_step_rank0": each.get("host"),
                   "what": "jrk_%s_f32_host: every input copied from pinned host memory, result copied back" % name},
           "config": workload_config(name, world)}
    if plan_e2e:
        res["e2e"]["plan"] = plan_e2e
    if name == "gram" and world > 1:
        res["e2e"]["note"] = "a materialised Gram moves 17.2 GB per rank device->host: all ranks share one host's PCIe/memory"
    if with_cpu and rank == 0 and world == 1:
        rows = cpu_rows_for(name)
        ev, dt, thr = cpu_reference_sample(name, rows)
        res["cpu_baseline"] = {"value": ev, "unit": "evals/s", "cores": thr, "kind": "port",
                               "sample": f"{rows} rows x {M} columns of the same workload, {dt:.1f} s, oracle/torch_ref.py "
                                         "(PORT of the reference op sequence, torch-CPU eager; jaxlib is absent)"}
    return res


def run_kmatw_mean(args, torch, nat, dist_ctx):
    """cfg3 variant of SURVEY.md §8(d): K@W with `Mean` over rows (jaxrk/reduce/base.py:142-150), out 1 x 16.  Every rank
    reduces its row block inside the kernel path (local sums), then ONE all_reduce(SUM) of 16 floats over NCCL and the
    division by the global N -- exactly what jaxrk_b200.sharded.sharded_inner_local does; the collective sits inside the
    CUDA-event window and is also timed on its own."""
    rank, world, dev = dist_ctx
    w = WORKLOADS["kmatw"]
    M, d, m = w["M"], w["d"], w["m"]
    n_local, n_total = w["N"] // world, w["N"]
    scale, power, nconst = kernel_params(d)

    def randn(seed, *shape):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)

    X, Y, W = randn(100 + rank, n_local, d), randn(1, M, d), randn(2, M, m)
    part = torch.empty((1, m), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ar_events = []

    def step():
        nat.kmatw(X, Y, W, scale, power, nconst, row_reduce=nat.REDUCE_SUM, out=part)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        if world > 1:
            torch.distributed.all_reduce(part, op=torch.distributed.ReduceOp.SUM)
        part.div_(float(n_total))
        a1.record()
        ar_events.append((a0, a1))

    steps = max(1, min(args.steps, 5))
    for _ in range(3):
        step()
    ar_events.clear()
    t_local, launches, clocks = _timed_steps(torch, dev, world, steps, flush, step, nat)
    t_step = _max_over_ranks(torch, dev, world, t_local) / steps
    ar_ms = float(np.median([a.elapsed_time(b) for a, b in ar_events]))
    # the in-step figure includes waiting for the slowest rank; the collective alone: back-to-back all_reduces of the same
    # 16 floats once every rank has arrived (the first of the burst absorbs the skew and is dropped)
    lat_ms = None
    if world > 1:
        torch.distributed.barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(22)]
        probe = torch.zeros_like(part)
        for i in range(21):
            evs[i].record()
            torch.distributed.all_reduce(probe, op=torch.distributed.ReduceOp.SUM)
        evs[21].record()
        torch.cuda.synchronize(dev)
        lat_ms = float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(1, 21)]))
    # cross-check against the public API (same kernels, same collective)
    from jaxrk_b200 import sharded
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import LinearReduce, Mean
    from jaxrk_b200.rkhs import FiniteVec
    k = GenGaussKernel.make_gauss(float(np.sqrt(d)))
    api = sharded.sharded_inner_local(k, X, [Mean()], n_total, rank * n_local, FiniteVec(k, Y, [LinearReduce(W.T.contiguous())]))
    err = float((api - part).abs().max() / (part.abs().max() + 1e-30))
    assert err < 1e-5, err
    pk, src = mufu_peak()
    achieved = float(n_local) * M / (t_local / steps)
    return {"value": float(n_total) * M / t_step, "ms_per_step": t_step * 1e3, "steps": steps, "scaling": "strong",
            "collective": {"op": "all_reduce(SUM) of 1 x 16 float32 + division by the global N", "backend": "nccl" if world > 1 else "none (1 rank)",
                           "median_ms_inside_the_step": ar_ms, "latency_ms_back_to_back": lat_ms, "ranks": world,
                           "note": "inside the step the all_reduce also waits for the slowest rank (rank skew); back-to-back = the collective alone"},
            "api_check_rel_err": err,
            "roofline": {"bound": "mufu", "achieved": achieved / 1e12, "peak": pk / 1e12, "unit": "T ex2/s", "frac": achieved / pk,
                         "traffic": None, "peak_source": src},
            "clocks": clocks, "gpu_launches": int(launches),
            "config": {"workload": "configs[2] variant: K(X,Y)@W with Mean over rows, out 1 x 16 (sharded_inner_local)",
                       "rows_per_gpu": n_local, "global_rows": n_total, "M": M, "d": d, "m": m}}


def run_kmatw_backward(args, torch, nat, dist_ctx):
    """The K@W vector-Jacobian product (dX and dscale rows) at cfg3's M, d, m on a 151552-row block of X PER GPU (weak: four
    256-row units per SM on every rank): jrk_kmatw_grad_f32 -> csrc/kgrad_tc.cu (both inner products on tcgen05, d + 2 FP32
    accumulations per pair).  Kernel-only; the roofline is the FP32 pipe at d + 2 lane-operations per pair."""
    rank, world, dev = dist_ctx
    w = WORKLOADS["kmatw"]
    M, d, m = w["M"], w["d"], w["m"]
    n_local = 148 * 256 * 4
    n_total = n_local * world
    scale, power, nconst = kernel_params(d)

    def randn(seed, *shape):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)

    X, Y, W, G = randn(400 + rank, n_local, d), randn(1, M, d), randn(2, M, m), randn(500 + rank, n_local, m)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    keep = {}

    def step():
        keep["g"] = nat.kmatw_grad(X, Y, W, G, scale, power, nconst)

    steps = max(1, min(args.steps, 3))
    for _ in range(3):
        step()
    t_local, launches, clocks = _timed_steps(torch, dev, world, steps, flush, step, nat)
    t_step = _max_over_ranks(torch, dev, world, t_local) / steps
    pk = fp32_pipe_peak()
    achieved = float(n_local) * M / (t_local / steps) * (d + 2)
    assert bool(torch.isfinite(keep["g"][0]).all())
    return {"value": float(n_total) * M / t_step, "unit": "pairs/s", "ms_per_step": t_step * 1e3, "steps": steps, "scaling": "weak",
            "roofline": {"bound": "fp32", "achieved": achieved / 1e12, "peak": pk[0] / 1e12, "unit": "T lane-ops/s",
                         "frac": achieved / pk[0], "traffic": None, "peak_source": pk[1], "lane_ops_per_pair": d + 2},
            "clocks": clocks, "gpu_launches": int(launches),
            "config": {"workload": "backward of configs[2]: dX and dscale of K(X,Y)@W for a 151552-row block of X per GPU",
                       "rows_per_gpu": n_local, "global_rows": n_total, "M": M, "d": d, "m": m,
                       "kernel": "kgrad_tc.cu (tcgen05 inner products, FP32 accumulation)"}}


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
    t_local = e0.elapsed_time(e1) * 1e-3
    t_all = _max_over_ranks(torch, dev, world, t_local)
    hbm, hbm_src, _ = peaks()
    achieved = 4.0 * n_local * M / t_local / 1e9
    return {"value": float(N) * M / t_all, "ms_per_step": t_all * 1e3, "steps": 1, "scaling": "strong",
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
    if not args.no_also:
        extra = [("kmatw_cfg3_mean", lambda: run_kmatw_mean(args, torch, nat, ctx)),
                 ("gram_cfg2" if args.workload == "kmatw" else "kmatw_cfg3",
                  lambda: run_workload("gram" if args.workload == "kmatw" else "kmatw", args, torch, nat, ctx, with_cpu=False)),
                 ("gram_1m_d16_streamed", lambda: run_gram1m(args, torch, nat, ctx)),
                 ("kmatw_cfg3_backward", lambda: run_kmatw_backward(args, torch, nat, ctx))]
        for key, fn in extra:
            try:
                torch.cuda.empty_cache()
                also[key] = fn()
            except Exception as e:  # noqa: BLE001  (the headline line must still print)
                also[key] = {"error": repr(e)}
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
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="kmatw")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

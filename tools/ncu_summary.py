"""Summarise an .ncu-rep (read here, no GPU): headline metrics per kernel + per-role instruction/stall breakdown.
Usage: python tools/ncu_summary.py file.ncu-rep [--top N]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_xu.sum",
        "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
seen_raw = set()
for r in rows[2:]:
    if r[idx["Kernel Name"]] in seen_raw:
        continue
    seen_raw.add(r[idx["Kernel Name"]])
    print("==", r[idx["Kernel Name"]][:110])
    for w in want:
        if w in idx:
            print(f"   {w:90s} {r[idx[w]]:>16s} {units[idx[w]]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
# the source page has one block per kernel: "Kernel Name" line, header line, then SASS lines
blocks, cur = [], None
for r in srows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if b["name"] in seen:
        continue
    seen.add(b["name"])
    h = {k: i for i, k in enumerate(b["hdr"])}
    data = b["rows"]
    tot = sum(int(r[h["Instructions Executed"]] or 0) for r in data)
    samp = sum(int(r[h["# Samples"]] or 0) for r in data)
    print("== source:", b["name"][:100], "| SASS lines", len(data), "| inst", tot, "| samples", samp)
    byc = collections.defaultdict(lambda: [0, 0, 0])
    for r in data:
        n = int(r[h["Instructions Executed"]] or 0)
        byc[n][0] += 1
        byc[n][1] += n
        byc[n][2] += int(r[h["# Samples"]] or 0)
    print("   exec-count classes (warps x iterations share a count => a role):")
    for n, (k, s_, sm) in sorted(byc.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"     count {n:>10}  x{k:>5} lines = {s_ / 1e6:8.2f} M inst ({100 * s_ / max(tot, 1):5.1f}%)  samples {sm} ({100 * sm / max(samp, 1):4.1f}%)")
    print(f"   top {top_n} SASS by stall samples:")
    for r in sorted(data, key=lambda r: -int(r[h["# Samples"]] or 0))[:top_n]:
        print(f"     {r[h['# Samples']]:>7} {r[h['Instructions Executed']]:>10}  {r[h['Source']][:110]}")

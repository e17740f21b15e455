"""Summarise an ncu source page (`ncu -i rep --page source --csv`) by instruction-count class: which warps
(identified by how often their instructions executed) hold the stall samples, and the top stalled instructions."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hrow = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hrow]
idx = {k: i for i, k in enumerate(h)}
data = rows[hrow + 1:]
stall = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, " instructions", sum(int(r[idx["Instructions Executed"]]) for r in data))
agg = {k: sum(int(r[idx[k]]) for r in data) for k in stall}
print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
c = collections.Counter(); n = collections.Counter()
for r in data:
    ex = int(r[idx["Instructions Executed"]]); c[ex] += int(r[idx["# Samples"]]); n[ex] += 1
print("samples by executed-count class (count, samples, #instructions):")
for ex, v in sorted(c.items(), key=lambda x: -x[1])[:10]:
    print("  ", ex, v, n[ex])
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:ntop]
for i in sorted(top):
    r = data[i]
    st = {k[6:]: int(r[idx[k]]) for k in stall if int(r[idx[k]]) > 0}
    print(i, r[idx["# Samples"]], r[idx["Instructions Executed"]], r[1].strip()[:64], sorted(st.items(), key=lambda x: -x[1])[:2])

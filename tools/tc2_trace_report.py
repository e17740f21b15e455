"""Summarise gpurun_out/tc2_trace_epq<E>.npy (tools/tc2_trace.py): per-role time per tile."""
import sys
import numpy as np
epq = int(sys.argv[1]) if len(sys.argv) > 1 else 4
suffix = ('_tok' + sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != '0' else ''
tr = np.load(f'gpurun_out/tc2_trace_epq{epq}{suffix}.npy')
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, tr - t0, 0)
g1, g2 = tr[17], tr[18]
sl = slice(60, 440)
print("period per tile", (g2[440, 3] - g2[60, 3]) / 380)
for name, g in (("G1", g1), ("G2", g2)):
    print(f"{name}: wait_a {np.mean(g[sl,1]-g[sl,0]):6.1f}  wait_b {np.mean(g[sl,2]-g[sl,1]):6.1f}  issue {np.mean(g[sl,3]-g[sl,2]):6.1f}  loop-rest {np.mean(g[61:441,0]-g[60:440,3]):6.1f}")
# latency from e_ready (last arrive of the 4 quads) to G2 issue done; from G2 done to G1 issue of tile+6; G1 issue to s_full seen
ws = np.arange(0, 4 * epq)
arr = np.array([max(tr[w, t, 3] for w in ws if tr[w, t, 0] > 0) for t in range(512)])
seen = np.array([min(tr[w, t, 1] for w in ws if tr[w, t, 0] > 0) for t in range(512)])
start = np.array([min(tr[w, t, 0] for w in ws if tr[w, t, 0] > 0) for t in range(512)])
print("e_ready(last quad) -> G2 issued      ", np.mean(g2[sl, 3] - arr[sl]))
print("G2 issued(t) -> G1 issued(t+6)       ", np.mean(g1[66:446, 3] - g2[60:440, 3]))
print("G1 issued(t) -> s_full seen (first)  ", np.mean(seen[sl] - g1[sl, 3]))
print("G1 issued(t) -> warp started waiting ", np.mean(start[sl] - g1[sl, 3]))
for w in ws:
    idx = np.where(tr[w, :, 0] > 0)[0]
    idx = idx[(idx > 60) & (idx < 440)]
    wait = (tr[w, idx, 1] - tr[w, idx, 0]).mean(); conv = (tr[w, idx, 2] - tr[w, idx, 1]).mean(); fin = (tr[w, idx, 3] - tr[w, idx, 2]).mean()
    cyc = np.diff(tr[w, idx, 0]).mean()
    print(f"warp {w} quad {w&3} k {w>>2}: wait_sfull {wait:7.1f} convert {conv:7.1f} st_wait+arrive {fin:6.1f}  cycle {cyc:7.1f} other {cyc-wait-conv-fin:6.1f}")

import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from jaxrk_b200 import _native as nat
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
N = 65536; d = 64
X = torch.randn(N, d, generator=g, device=dev); Y = torch.randn(N, d, generator=g, device=dev)
K = torch.empty((N, N), device=dev)
s = float(1/np.sqrt(d))
for _ in range(2): nat.gram(X, Y, s, 2.0, 0.25, path=2, out=K)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nat.gram(X, Y, s, 2.0, 0.25, path=2, out=K); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(min(ts))
'''
for staged in ("2", "3", "2", "3"):
    for dbg in (0,):
        env = dict(os.environ, JRK_TF32_STAGED="0", JRK_TF32_DBG=str(dbg), JRK_TF32_NACC=staged)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        print(f"nacc={staged} dbg={dbg:2d} (1=noSTG 2=noMMA 4=noCONV 8=noLDTM): {r.stdout.strip()} ms {r.stderr.strip()[-200:]}", flush=True)

"""One call of the tensor-core K@W path with a deadline-friendly footprint: python tools/tc2_one.py N M d m [epq]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
N, M, d, m = (int(v) for v in sys.argv[1:5])
if len(sys.argv) > 5:
    nat.set_option(nat.OPT_TC2_EPQ, int(sys.argv[5]))
if len(sys.argv) > 6:
    nat.set_option(nat.OPT_TC2_TOKEN, int(sys.argv[6]))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(int(os.environ.get('SEED', '0')))
X = torch.randn(N, d, generator=g, device=dev); Y = torch.randn(M, d, generator=g, device=dev); W = torch.randn(M, m, generator=g, device=dev)
s, nc = float(0.70710695 / np.sqrt(d)), 0.25
if os.environ.get('WONES'):
    W = torch.ones_like(W)
c = s * s * 1.4426950408889634
print('launching; bound', c * float((X * X).sum(1).max() + (Y * Y).sum(1).max()), 'max|W|', float(W.abs().max()), flush=True)
A = nat.kmatw(X, Y, W, s, 2.0, nc); torch.cuda.synchronize()
print('kernel done', flush=True)
rows = torch.arange(0, N, max(1, N // 64), device=dev)[:64]
Kt = nc * torch.exp(-(s * torch.cdist(X[rows].double(), Y.double())) ** 2)
T = Kt @ W.double(); bound = Kt.abs() @ W.double().abs()
print("ok", sys.argv[1:], "err/lim", float(((A[rows].double() - T).abs() / (1e-6 + 1e-5 * bound)).max()), flush=True)

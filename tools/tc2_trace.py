"""Per-warp timeline of the second-generation K@W kernel (CTA 0, first unit): clock64 stamps per tile.
   python tools/tc2_trace.py [epq]  ->  gpurun_out/tc2_trace_epq<E>.npy  (shape [19, 512, 4])"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

epq = sys.argv[1] if len(sys.argv) > 1 else "4"
nat.set_option(nat.OPT_TC2_EPQ, int(epq))
tok = sys.argv[2] if len(sys.argv) > 2 else "0"
nat.set_option(nat.OPT_TC2_TOKEN, int(tok))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
N, M, d, m = 148 * 128 * 2, 512 * 64, 16, 16
X = torch.randn(N, d, generator=g, device=dev)
Y = torch.randn(M, d, generator=g, device=dev)
W = torch.randn(M, m, generator=g, device=dev)
s, nc = float(0.70710695 / np.sqrt(d)), 0.25
nat.kmatw(X, Y, W, s, 2.0, nc)
torch.cuda.synchronize()
buf = torch.zeros(19 * 512 * 4, dtype=torch.int64, device=dev)
L = nat.lib()
L.jrk_debug_tc2_trace.argtypes = [ctypes.c_void_p]
L.jrk_debug_tc2_trace.restype = None
L.jrk_debug_tc2_trace(ctypes.c_void_p(buf.data_ptr()))
nat.kmatw(X, Y, W, s, 2.0, nc)
torch.cuda.synchronize()
L.jrk_debug_tc2_trace(None)
out = buf.cpu().numpy().reshape(19, 512, 4)
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/tc2_trace_epq{epq}" + (f"_tok{tok}" if tok != "0" else "") + ".npy", out)
print("saved", out.shape, int((out != 0).sum()))

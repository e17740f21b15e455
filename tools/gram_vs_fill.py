"""Equal-clock A/B of the cfg2 Gram kernel against a plain fill of the same 17.18 GB.

Run under Nsight Compute with ITS clock control (`--clock-control base` locks the SM and memory clocks for the duration
of the profile and releases them afterwards; nothing is changed through nvidia-smi), so both kernels see the same clocks
and no power-cap throttling:

  ncu --clock-control base --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,\
l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_write.sum,dram__bytes_read.sum,\
sm__cycles_active.avg,smsp__inst_executed.sum -k regex:'gram_tf32_kernel|fill|FillFunctor' --csv python tools/gram_vs_fill.py

Without ncu it prints the CUDA-event times (free-running clocks) of the same launches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
N = M = 65536
d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(N, d, generator=g, device=dev)
Y = torch.randn(M, d, generator=g, device=dev)
K = torch.empty((N, M), device=dev, dtype=torch.float32)
s = float(0.70710695 / np.sqrt(d))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
for it in range(3):
    ev[0].record()
    K.fill_(1.0)
    ev[1].record()
    nat.gram(X, Y, s, 2.0, 0.25, out=K)
    ev[2].record()
    torch.cuda.synchronize()
    print({"iter": it, "fill_ms": ev[0].elapsed_time(ev[1]), "gram_ms": ev[1].elapsed_time(ev[2])})

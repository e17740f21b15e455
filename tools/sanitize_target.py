"""Smallest launches that touch every kernel and every ragged edge, for `compute-sanitizer --tool memcheck`."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=dev)
for p in (2.0, 1.0, 1.5):
    for (N, M, d) in ((130, 67, 5), (257, 131, 16), (129, 200, 33)):
        nat.gram(r(N, d), r(M, d), 0.3, p, 0.2, path=nat.PATH_DIRECT)
    for (N, M, d) in ((130, 67, 64), (65, 129, 70), (200, 130, 128)):
        X = r(N, d)
        nat.gram(X, r(M, d), 0.1, p, 0.2, path=nat.PATH_TF32X3)
        nat.gram(X, None, 0.1, p, 0.2, path=nat.PATH_TF32X3)
    for (N, M, d, m) in ((130, 300, 16, 16), (700, 129, 3, 5), (64, 1000, 17, 33)):
        X, Y, W = r(N, d), r(M, d), r(M, m)
        nat.kmatw(X, Y, W, 0.3, p, 0.2)
        nat.kmatw(X, Y, W, 0.3, p, 0.2, row_reduce=nat.REDUCE_MEAN)
    X = r(512, 8)
    nat.gram_groupsum(X, None, 0.3, p, 0.2, 128, 128, True)
    nat.gram_groupsum(X, r(300, 8), 0.3, p, 0.2, 256, 100, False)
    # tensor-core paths with the new operand forms: tf32 (d <= 16) and fp16 (d > 16) Gram, both K@W epilogues, ragged
    for (N, M, d) in ((130, 67, 12), (65, 129, 20), (200, 130, 33), (77, 190, 100)):
        X = r(N, d)
        nat.gram(X, r(M, d), 0.1, p, 0.2, path=nat.PATH_TF32X3)
        nat.gram(X, None, 0.1, p, 0.2, path=nat.PATH_TF32X3)
    for (N, M, d, m, s_) in ((8192 + 5, 4096 + 70, 16, 16, 0.15), (8192, 4096 + 3, 7, 3, 0.2), (8192 + 130, 4096, 30, 16, 1.5)):
        X, Y, W = r(N, d), r(M, d), r(M, m)
        nat.set_option(nat.OPT_KMATW_TC, 1)
        nat.kmatw(X, Y, W, s_, p, 0.2)                                     # factored (small s_) / generic
        nat.kmatw(X, Y, W, s_, p, 0.2, row_reduce=nat.REDUCE_SUM)
    # fused backward kernels and the sibling epilogues
    for (N, M, d, m) in ((130, 300, 16, 16), (257, 129, 3, 5), (64, 200, 40, 32)):
        X, Y, W, G = r(N, d), r(M, d), r(M, m), r(N, m)
        nat.kmatw_grad(X, Y, W, G, 0.3, p, 0.2)
        nat.gram_grad(X, Y, r(N, M), 0.3, p, 0.2)
    for spec in (nat.KernelSpec(nat.KIND_PERIODIC, (0.5, 1.0)), nat.KernelSpec(nat.KIND_THRESHSPIKE, (1.0, p, 1.0, 0.1, 1.5))):
        nat.gram_spec(r(130, 5), r(67, 5), spec)
        nat.kmatw_spec(r(130, 5), r(300, 5), r(300, 9), spec)
nat.dist_diag(r(100, 7), r(100, 7), 1.0, 1.5)
nat.dist(r(100, 7), r(50, 7), 1.0, 1.0)
torch.cuda.synchronize()
print("done", nat.launch_count())

"""Pure-write HBM bandwidth reference points (torch fill / memset), for the write-bound Gram roofline."""
import torch
dev = torch.device("cuda:0")
n = 65536 * 65536
K = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: K.fill_(1.0)); print("fill_ 17.18 GB:", ms, "ms", 4 * n / ms / 1e6, "GB/s")
ms = t(lambda: K.zero_()); print("zero_ 17.18 GB:", ms, "ms", 4 * n / ms / 1e6, "GB/s")
A = torch.empty(n // 8, dtype=torch.float32, device=dev); B = torch.empty_like(A)
ms = t(lambda: B.copy_(A)); print("copy 2.1+2.1 GB:", ms, "ms", 8 * (n // 8) / ms / 1e6, "GB/s (read+write)")
ms = t(lambda: torch.exp(A, out=B)); print("exp 2.1+2.1 GB:", ms, "ms", 8 * (n // 8) / ms / 1e6, "GB/s (read+write)")

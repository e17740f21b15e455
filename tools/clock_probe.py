"""SM clock / power while a kernel loops (NVML samples every 5 ms)."""
import os, sys, threading, time
import numpy as np, torch, pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrk_b200 import _native as nat
dev = torch.device("cuda:0")
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
g = torch.Generator(device=dev).manual_seed(0)
def probe(name, fn, seconds=1.5):
    fn(); torch.cuda.synchronize()
    samples, stop = [], [False]
    def run():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
            time.sleep(0.005)
    th = threading.Thread(target=run); th.start()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(10): fn()
        n += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    a = np.array([(c, p) for c, p, _ in samples[len(samples)//3:]])
    reasons = 0
    for _, _, r in samples: reasons |= r
    print(f"{name}: {e0.elapsed_time(e1)/n:.3f} ms/iter, SM clock median {np.median(a[:,0]):.0f} min {a[:,0].min():.0f} MHz, power median {np.median(a[:,1]):.0f} W, reasons 0x{reasons:x}", flush=True)
N, d = 65536, 64
X = torch.randn(N, d, generator=g, device=dev); Y = torch.randn(N, d, generator=g, device=dev)
K = torch.empty((N, N), device=dev); s = float(1/np.sqrt(d))
probe("gram tf32 65536^2 d=64", lambda: nat.gram(X, Y, s, 2.0, 0.25, path=2, out=K))
probe("fill_ 17 GB", lambda: K.fill_(1.0))
X8 = X[:, :8].contiguous(); Y8 = Y[:, :8].contiguous()
probe("gram direct d=8", lambda: nat.gram(X8, Y8, s, 2.0, 0.25, path=1, out=K))
del K
M = 1 << 20
Xs = torch.randn(32768, 16, generator=g, device=dev); Ys = torch.randn(M, 16, generator=g, device=dev); W = torch.randn(M, 16, generator=g, device=dev)
probe("kmatw 32768 x 1M d=16 m=16", lambda: nat.kmatw(Xs, Ys, W, 0.25, 2.0, 0.25), seconds=3)

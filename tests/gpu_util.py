"""Shared helpers of the GPU parity tests (oracle-side; test infrastructure)."""
import numpy as np
import torch

from oracle import jaxrk_oracle as orc

f32 = np.float32
RTOL, ATOL = 1e-5, 1e-6   # BASELINE.json north_star: rel 1e-5 / abs 1e-6

# (name, length_scale -> SimpleScaler array, power)
KERNELS = {
    "sqexp": (lambda ls: orc.simple_scaler_s(orc.gauss_length_scale(float(ls))), 2.0),
    "laplace": (lambda ls: orc.simple_scaler_s(float(ls)), 1.0),
    "gg15": (lambda ls: orc.simple_scaler_s(float(ls)), 1.5),
    "gg05": (lambda ls: orc.simple_scaler_s(float(ls)), 0.5),
}


def kparams(name, ls):
    mk, p = KERNELS[name]
    s = mk(ls)
    return s, float(s[0, 0]), p, float(orc.gengauss_nconst_ref32(s, p).reshape(-1)[0])


def dev():
    return torch.device("cuda:0")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=f32)).to(dev())


def gauss(seed, *shape):
    return np.random.RandomState(seed).randn(*shape).astype(f32)


def assert_close_truth(got, truth, what="", rtol=RTOL, atol=ATOL):
    got = np.asarray(got, dtype=np.float64)
    err = np.abs(got - truth)
    lim = atol + rtol * np.abs(truth)
    bad = err > lim
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {bad.size} entries out of tolerance; "
                           f"worst err {err.max():.3e} (limit there {lim.flat[err.argmax()]:.3e})")


def assert_kmatw_close(got, truth, absbound, what="", rtol=RTOL, atol=ATOL):
    """|got - truth| <= atol + rtol * sum_j |K_ij W_jc|  (the float32
    accumulation is judged against the magnitude of what was accumulated)."""
    got = np.asarray(got, dtype=np.float64)
    err = np.abs(got - truth)
    lim = atol + rtol * absbound
    bad = err > lim
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {bad.size} entries out of tolerance; worst err "
                           f"{err.max():.3e} vs limit {lim.flat[err.argmax()]:.3e}")

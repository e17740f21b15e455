"""Array plumbing for the host layer: everything the kernels see is a float32 CUDA tensor.

The reference's array type is a JAX DeviceArray; jax is absent from this image, so the
host layer uses torch CUDA tensors for device memory and streams (binding "L1b" of
SURVEY.md §8b).  JAX without x64 truncates every input to float32 on entry
(tests/test_kern.py:17 passes int32, tests/test_utilities.py:12 float64; quirk Q6);
`asarray` does the same.
"""
from __future__ import annotations

import numpy as np
import torch

Array = torch.Tensor

_device = None


def set_device(dev) -> None:
    """Pin the CUDA device used for new arrays (default: LOCAL_RANK under torchrun, else cuda:0)."""
    global _device
    _device = torch.device(dev)


def device() -> torch.device:
    global _device
    if _device is None:
        import os
        if not torch.cuda.is_available():
            raise RuntimeError("jaxrk_b200 needs a CUDA device: there is no CPU fallback")
        _device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
    return _device


def asarray(x, dev=None) -> torch.Tensor:
    """float32 CUDA tensor view/copy of x (numpy array, sequence, scalar or tensor)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda and x.dtype == torch.float32:
            return x
        return x.to(device=dev or (x.device if x.is_cuda else device()), dtype=torch.float32)
    a = np.asarray(x)
    if a.dtype != np.float32:
        a = a.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev or device())


def as2d(x, dev=None) -> torch.Tensor:
    t = asarray(x, dev)
    assert t.dim() == 2, "expected a 2-D array with one point per row"
    return t


def to_numpy(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)

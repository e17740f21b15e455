"""The C-ABI library loads, exports every symbol include/jrk.h declares, validates its
arguments, and -- without a GPU -- fails loudly instead of falling back (no compute here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "jrk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jrk_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(nat.LIB_PATH), "run `python -m jaxrk_b200.build`"
    assert os.path.dirname(nat.LIB_PATH).startswith(os.path.join(ROOT, "jaxrk_b200"))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 12
    raw = ctypes.CDLL(nat.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/jrk.h but not exported"
        assert n in nat.SIGNATURES, f"{n} has no ctypes signature in jaxrk_b200/_native.py"
    assert sorted(nat.SIGNATURES) == names
    assert nat.lib().jrk_abi_version() == nat.ABI_VERSION == 3


def test_sm100a_code_is_embedded():
    blob = open(nat.LIB_PATH, "rb").read()
    assert b"sm_100a" in blob


def _call_gram(N=4, M=4, d=2, scale=1.0, power=2.0, nconst=1.0, ldk=None, path=0, X=1, K=1):
    L = nat.lib()
    return L.jrk_gram_f32(ctypes.c_void_p(X * 256), None, ctypes.c_void_p(K * 512), N, M, d, d, d, M if ldk is None else ldk,
                          scale, None, power, nconst, path, None, 0, None)


def test_argument_validation_happens_before_any_cuda_call():
    assert _call_gram(power=2.5) == 1 and "power" in nat.last_error()
    assert _call_gram(power=0.0) == 1
    assert _call_gram(scale=-1.0) == 1 and "scale" in nat.last_error()
    assert _call_gram(d=0) == 1
    assert _call_gram(ldk=2) == 1 and "leading" in nat.last_error()
    assert _call_gram(path=7) == 1
    assert _call_gram(X=0) == 1
    assert _call_gram(N=0) == 0  # empty input: nothing to do, ok
    L = nat.lib()
    p = ctypes.c_void_p(256)
    rc = L.jrk_kmatw_f32(p, p, p, p, 6, 4, 2, 1, 2, 2, 1, 1, 1.0, None, 1.0, 1.0, None, None, nat.REDUCE_BALANCED, 4,
                         None, 0, None)
    assert rc == 1 and "divide" in nat.last_error()
    rc = L.jrk_gram_groupsum_f32(p, None, p, 6, 6, 2, 2, 2, 2, 1.0, None, 1.0, 1.0, None, None, 4, 3, 0, None, 0, None)
    assert rc == 1
    # sibling kernels (jrk_*_kind_f32): parameter checks of PeriodicKernel / ThreshSpikeKernel.__init__
    kp = (ctypes.c_float * 5)(1.0, 1.0, 1.0, 2.0, 0.5)          # |non_spike| >= spike
    rc = L.jrk_gram_kind_f32(p, None, p, 4, 4, 2, 2, 2, 4, nat.KIND_THRESHSPIKE, kp, 5, None, 0, None, 0, None)
    assert rc == 1 and "non_spike" in nat.last_error()
    rc = L.jrk_gram_kind_f32(p, None, p, 4, 4, 2, 2, 2, 4, nat.KIND_PERIODIC, kp, 1, None, 0, None, 0, None)
    assert rc == 1 and "PERIODIC needs" in nat.last_error()
    rc = L.jrk_kmatw_kind_f32(p, p, p, p, 4, 4, 2, 1, 2, 2, 1, 1, 9, kp, 5, None, None, None, 0, 0, None, 0, None)
    assert rc == 1 and "unknown kernel kind" in nat.last_error()
    rc = L.jrk_dist_diag_f32(p, p, p, 4, 2, 1, 2, 1.0, None, 1.0, None)
    assert rc == 1 and "leading" in nat.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    assert nat.lib().jrk_device_count() == 0
    X = np.zeros((4, 2), np.float32)
    K = np.zeros((4, 4), np.float32)
    with pytest.raises(nat.JrkError) as ei:
        nat.gram_host(X, None, K, 1.0, 2.0, 1.0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(nat.JrkError):
        nat.kmatw_host(X, X, np.ones((4, 1), np.float32), np.zeros((4, 1), np.float32), 1.0, 2.0, 1.0)
    assert _call_gram() == 2


def test_workspace_queries_are_pure():
    L = nat.lib()
    assert L.jrk_gram_workspace_bytes(1024, 1024, 8, nat.PATH_DIRECT) == 0
    assert L.jrk_kmatw_workspace_bytes(1 << 20, 1 << 20, 16, 16, 0, 0) > 0
    assert L.jrk_launch_count() >= 0


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under jaxrk_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, "jaxrk_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle\b", src):
                    offenders.append(os.path.join(dirpath, f))
    assert offenders == [], offenders


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", str(tmp_path / "libjaxrk_b200.so"))
    with pytest.raises(ImportError) as e:
        nat.lib()
    assert "no CPU or eager fallback" in str(e.value)


def test_options_are_explicit_process_state_not_environment_reads():
    """jrk_set_option / jrk_get_option (ABI 3): defaults, validation, and no getenv on the compute paths (the
    environment only seeds the initial value, in capi.cu)."""
    L = nat.lib()
    assert nat.get_option(nat.OPT_KMATW_TC) in (0, 1)
    prev = nat.set_option(nat.OPT_TC2_EPQ, 3)
    assert nat.get_option(nat.OPT_TC2_EPQ) == 3
    nat.set_option(nat.OPT_TC2_EPQ, prev)
    assert L.jrk_set_option(99, 1) == 1 and "unknown option" in nat.last_error()
    assert L.jrk_set_option(nat.OPT_TC2_EPQ, 7) == 1
    with nat.option(nat.OPT_GRAM_H16, 0):
        assert nat.get_option(nat.OPT_GRAM_H16) == 0
    assert nat.get_option(nat.OPT_GRAM_H16) == 1
    csrc = os.path.join(ROOT, "jaxrk_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")) and f != "capi.cu":
            assert "getenv" not in open(os.path.join(csrc, f)).read(), f"{f} reads the environment"
    assert "thread_local ExtraKind" not in open(os.path.join(csrc, "capi.cu")).read()


def test_plan_argument_validation_and_lifetime_without_a_device():
    L = nat.lib()
    h = ctypes.c_void_p()
    p = ctypes.c_void_p(256)
    assert L.jrk_kmatw_plan_create(None, p, p, 8, 2, 1, 2, 1, 1.0, None, 2.0, 1.0, None, 1, 0, None) == 1
    assert L.jrk_kmatw_plan_create(ctypes.byref(h), None, p, 8, 2, 1, 2, 1, 1.0, None, 2.0, 1.0, None, 1, 0, None) == 1
    assert L.jrk_kmatw_plan_create(ctypes.byref(h), p, p, 8, 2, 1, 1, 1, 1.0, None, 2.0, 1.0, None, 1, 0, None) == 1      # ldy < d
    assert L.jrk_kmatw_plan_create(ctypes.byref(h), p, p, 8, 2, 1, 2, 1, 1.0, None, 2.5, 1.0, None, 1, 0, None) == 1      # power
    assert h.value is None
    assert L.jrk_kmatw_plan_destroy(None) == 0
    assert L.jrk_kmatw_plan_apply(None, p, p, 4, 2, 1, None, 0, 0, None, 0, None) == 1
    assert L.jrk_kmatw_plan_workspace_bytes(None, 4, 0, 0) == 0
    if not torch.cuda.is_available():
        Y = np.zeros((8, 2), np.float32)
        with pytest.raises(nat.JrkError) as e:
            nat.KmatwPlan(Y, np.zeros((8, 1), np.float32), 1.0, 2.0, 1.0)
        assert e.value.code == 2 and "no CPU fallback" in str(e.value)

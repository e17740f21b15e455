"""The XLA FFI shim (jaxrk_b200/csrc/xla_ffi_shim.cc, binding "L1a" of SURVEY.md section 8b) against include/jrk.h.

jax / jaxlib are absent from the image, so the real FFI header is not available; the shim is compiled here against a
MOCK of the API surface it uses (tests/ffi_mock/xla/ffi/api/ffi.h), which checks every handler's signature against its
Bind() chain at compile time, and every call into the C ABI against the header.  A text pass checks coverage: one
handler per device-pointer compute entry point, with the optional arrays of the ABI passed through."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "jaxrk_b200", "csrc", "xla_ffi_shim.cc")
CUDA_INC = "/usr/local/cuda/include"

# device-pointer compute entry points of include/jrk.h -> optional arrays the handler must forward
COMPUTE = {
    "jrk_gram_f32": ["dim_scale"],
    "jrk_dist_f32": ["dim_scale"],
    "jrk_dist_diag_f32": ["dim_scale"],
    "jrk_gram_affine_f32": ["row_mul", "col_mul", "row_add", "col_add"],
    "jrk_kmatw_f32": ["dim_scale", "row_pref", "col_pref"],
    "jrk_gram_kind_f32": ["dim_scale"],
    "jrk_kmatw_kind_f32": ["dim_scale", "row_pref", "col_pref"],
    "jrk_gram_groupsum_f32": ["dim_scale", "row_pref", "col_pref"],
    "jrk_kmatw_grad_f32": [],
    "jrk_gram_grad_f32": [],
}


def test_shim_has_one_handler_per_compute_entry_point():
    src = open(SHIM).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "jrk.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(jrk_[a-z0-9_]+)\s*\(", hdr))
    handlers = dict(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((jrk_[a-z0-9_]+)_ffi,\s*(\w+)", src))
    assert set(handlers) == set(COMPUTE), (sorted(handlers), sorted(COMPUTE))
    for name, optional in COMPUTE.items():
        assert name in declared, f"{name} is not declared in include/jrk.h"
        impl = handlers[name]
        body = src[src.index(f"static ffi::Error {impl}("):]
        body = body[:body.index("\n}\n")]
        assert re.search(rf"\b{name}\(", body), f"{impl} does not call {name}"
        for o in optional:
            assert f"opt({o})" in body, f"{impl} does not forward the optional array `{o}`"
    # every device-pointer compute entry point of the header is covered (plans / host variants / plumbing are not FFI calls)
    not_ffi = {n for n in declared if n.endswith(("_host", "_bytes")) or "_plan_" in n} | {
        "jrk_abi_version", "jrk_last_error", "jrk_device_count", "jrk_launch_count", "jrk_set_option", "jrk_get_option",
        "jrk_kmatw_uses_tensor_cores"}
    assert declared - not_ffi == set(COMPUTE), sorted(declared - not_ffi - set(COMPUTE))


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime_api.h")), reason="needs the CUDA headers")
def test_shim_compiles_against_the_mock_ffi_header_and_the_c_abi():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror=return-type", "-I", os.path.join(ROOT, "tests", "ffi_mock"),
           "-I", CUDA_INC, SHIM]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    # and the check has teeth: a handler whose Bind() chain lost an argument must NOT compile
    broken = open(SHIM).read().replace('.Attr<int32_t>("path")\n                                  .Ret<F32Buf>(),\n'
                                       '                              {ffi::Traits::kCmdBufferCompatible});\n\n'
                                       'XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_dist_f32_ffi',
                                       '.Ret<F32Buf>(),\n                              {ffi::Traits::kCmdBufferCompatible});\n\n'
                                       'XLA_FFI_DEFINE_HANDLER_SYMBOL(jrk_dist_f32_ffi', 1)
    assert broken != open(SHIM).read()
    r2 = subprocess.run(cmd[:-1] + ["-x", "c++", "-", "-I", os.path.dirname(SHIM)], input=broken.replace(
        '#include "../../include/jrk.h"', f'#include "{os.path.join(ROOT, "include", "jrk.h")}"'), capture_output=True, text=True)
    assert r2.returncode != 0 and "handler signature != Bind() chain" in r2.stderr

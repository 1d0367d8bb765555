"""The C-ABI library loads on a box without a GPU, exports every symbol include/spd.h declares,
and validates arguments before touching the device (error codes, never exceptions from C)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from semi_pd_b200 import _lib


def test_every_declared_symbol_is_exported_and_bound():
    declared = _lib.declared_functions()
    assert len(declared) >= 26
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and include/spd.h disagree"
    for name in declared:
        assert hasattr(_lib.lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (spd_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_header_cites_reference_for_every_entry_point():
    text = open(_lib.HEADER_PATH).read()
    for needle in ("K/csrc/torch_extension.cc", "semi-pd-ipc/ipc.cpp", "S/layers/linear.py",
                   "S/layers/attention/flashinfer_backend.py", "S/mem_cache/memory_pool.py",
                   "S/model_executor/forward_batch_info.py", "S/entrypoints/engine.py"):
        assert needle in text, needle


def test_library_has_no_link_time_driver_dependency():
    out = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out  # driver entry points are resolved at run time
    assert "libtorch" not in out and "libc10" not in out  # plain C ABI, no torch types


def test_version_and_sizes():
    assert _lib.lib.spd_version() >= 100
    assert _lib.lib.spd_gemm_workspace_bytes() > (32 << 20)
    assert _lib.lib.spd_attn_decode_workspace_bytes(64, 32, 128) > 64 * 32 * 32 * 128 * 4


def test_argument_errors_are_codes_not_crashes():
    lib = _lib.lib
    # null pointers
    assert lib.spd_rmsnorm(None, None, None, 4, 128, 128, 128, 1e-6, 0, None) == -1
    assert b"null" in lib.spd_last_error_string()
    # bad dtype enum / alignment / shapes, all rejected before any CUDA call
    buf = (C.c_char * 4096)()
    p = C.addressof(buf)
    p16 = (p + 15) & ~15
    assert lib.spd_silu_and_mul(p16, p16, 1, 12, 0, None) == -2
    assert lib.spd_gemm(p16, p16, p16, 4, 8, 100, 100, 100, 8, 0, 0, 0, None, 0, 0, None) == -2
    assert lib.spd_gemm(p16, p16, p16, 4, 8, 128, 128, 128, 8, 7, 0, 0, None, 0, 0, None) == -1
    assert lib.spd_attn_decode_paged(p16, p16, p16, p16, p16, p16, p16, 1, 30, 8, 128, 3840, 3840, 8,
                                     0.1, 0, None, 0, 0, None) == -3  # Hq % Hk != 0
    assert lib.spd_attn_prefill_ragged(p16, p16, p16, p16, p16, p16, p16, p16, p16, 1, 8, 8, 8, 2, 96,
                                       768, 768, 8, 0.1, 0, 0, None) == -3  # head_dim
    with pytest.raises(RuntimeError, match="libspd"):
        _lib.check(-1, "demo")


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semi_pd_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.rmsnorm(torch.zeros(2, 128, dtype=torch.bfloat16), torch.ones(128, dtype=torch.bfloat16))
    # a raw call with host pointers returns a CUDA error code (> 0), it does not fall back
    buf = (C.c_char * 8192)()
    p16 = (C.addressof(buf) + 15) & ~15
    rc = _lib.lib.spd_rmsnorm(p16, p16 + 1024, p16 + 2048, 1, 128, 128, 128, 1e-6, 0, None)
    assert rc > 0

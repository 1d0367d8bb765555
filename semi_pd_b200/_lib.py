"""ctypes binding of libspd.so — the C-ABI in include/spd.h.

The library is the product: there is no Python / torch fallback for any op.  If the shared
object is missing, import fails loudly; if it is present but no GPU is (the CPU test box), the
symbols load and every compute call returns a CUDA error that `check()` raises as RuntimeError —
the same error behaviour as the reference's TORCH_CHECK in sgl-kernel
(sgl-kernel/csrc/elementwise/fused_add_rms_norm_kernel.cu:43-53).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPD_LIB_PATH: load another build of the same ABI (A/B runs of kernel changes on one box)
LIB_PATH = os.environ.get("SPD_LIB_PATH") or os.path.join(_HERE, "libspd.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "spd.h")

SPD_BF16 = 0
SPD_F16 = 1
SPD_EPI_NONE = 0
SPD_EPI_SILU_MUL = 1

_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_f = C.c_float

# name -> (restype, argtypes); must list every function include/spd.h declares
SIGNATURES = {
    "spd_version": (_i, []),
    "spd_last_error_string": (C.c_char_p, []),
    "spd_device_sm_count": (_i, [_i, C.POINTER(_i)]),
    "spd_partition_create": (_i, [_i, _i, _i, C.POINTER(_p)]),
    "spd_partition_sm_counts": (_i, [_p, C.POINTER(_i), C.POINTER(_i)]),
    "spd_partition_stream": (_i, [_p, _i, C.POINTER(_p)]),
    "spd_partition_destroy": (_i, [_p]),
    "spd_rmsnorm": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _f, _i, _p]),
    "spd_fused_add_rmsnorm": (_i, [_p, _p, _p, _i64, _i64, _f, _i, _p]),
    "spd_silu_and_mul": (_i, [_p, _p, _i64, _i64, _i, _p]),
    "spd_layernorm": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _f, _i, _p]),
    "spd_bias_act_residual": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _i, _p]),
    "spd_rope_inplace": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _i64, _i64, _i, _i, _p]),
    "spd_rope_kvwrite": (
        _i,
        [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i64, _i64, _i64, _i, _i, _p],
    ),
    "spd_kv_write": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i, _p]),
    "spd_gather_rows": (_i, [_p, _p, _p, _i64, _i64, _i64, _i, _p]),
    "spd_argmax_f32": (_i, [_p, _p, _i64, _i64, _i64, _p]),
    "spd_compute_positions": (_i, [_p, _p, _p, _p, _p, _i64, _p]),
    "spd_decode_positions": (_i, [_p, _p, _i64, _p]),
    "spd_build_kv_indices": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "spd_write_req_to_token": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "spd_gemm_workspace_bytes": (_i64, []),
    "spd_gemm": (
        _i,
        [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i, _i, _i, _p, _i64, _i, _p],
    ),
    "spd_attn_prefill_ragged": (
        _i,
        [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _i, _i, _i64, _i64, _i64, _f, _i, _i, _p],
    ),
    "spd_attn_decode_workspace_bytes": (_i64, [_i64, _i, _i]),
    "spd_attn_decode_paged": (
        _i,
        [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i64, _i64, _i64, _f, _i, _p, _i64, _i, _p],
    ),
    "spd_comm_create": (_i, [_i, _i, _i64, C.POINTER(_p), _p]),
    "spd_comm_connect": (_i, [_p, _p]),
    "spd_comm_max_bytes": (_i64, [_p]),
    "spd_comm_status": (_i, [_p, C.POINTER(_i)]),
    "spd_allreduce": (_i, [_p, _p, _p, _i64, _i, _i, _p]),
    "spd_allreduce_add_rmsnorm": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _f, _i, _p]),
    "spd_allgather": (_i, [_p, _p, _p, _i64, _p]),
    "spd_comm_destroy": (_i, [_p]),
    "spd_argmax_partial_f32": (_i, [_p, _p, _i64, _i64, _i64, _i64, _p]),
    "spd_argmax_merge": (_i, [_p, _p, _i, _i64, _p]),
    "spd_gemm_grouped": (_i, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i, _p, _i64, _i64, _i64, _i64, _i,
                              _i, _i, _p]),
    "spd_moe_route": (_i, [_p, _p, _p, _i64, _i, _i, _f, _i, _i, _p]),
    "spd_moe_align_rows": (_i64, [_i64, _i, _i]),
    "spd_moe_align": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i64, _p]),
    "spd_moe_gather": (_i, [_p, _p, _p, _i64, _i64, _i64, _i, _p]),
    "spd_moe_combine": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i64, _i, _p]),
    "spd_moe_dense_act": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i64, _i, _p]),
    "spd_attn_mla_workspace_bytes": (_i64, [_i, _i]),
    "spd_attn_mla": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i64, _i64, _i64, _f, _i, _p,
                          _i64, _i, _p]),
    "spd_attn_prefill_simt": (
        _i,
        [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _i, _i, _i64, _i64, _i64, _f, _i, _p],
    ),
    "spd_attn_decode_simt": (
        _i,
        [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i64, _i64, _i64, _f, _i, _p, _i64, _p],
    ),
}


def declared_functions(header_path: str = HEADER_PATH):
    """Names of all functions declared in include/spd.h (used by the symbol-export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spd_[a-z0-9_]+)\s*\(", text)))


class SpdLibraryMissing(ImportError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise SpdLibraryMissing(
            f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C semi_pd_b200/csrc`. There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class _Counter:
    """Kernel launches issued through the C-ABI by this process (bench.py's `gpu_launches`).
    Every successful compute entry point is one kernel launch; replays of a captured CUDA graph
    are accounted by CudaGraphRunner (kernels per graph x replays)."""
    launches = 0


def check(rc: int, what: str = "") -> None:
    """Raise RuntimeError for a non-zero return code, carrying the library's error text."""
    _Counter.launches += 1
    if rc != 0:
        msg = lib.spd_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libspd {what} failed (code {rc}): {msg}")

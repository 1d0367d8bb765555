"""`torch.ops` registration of the hot-path operators over the C-ABI library.

The reference binds its CUDA operators as dispatcher ops (`sgl-kernel/csrc/torch_extension.cc:49-73`,
namespace `sgl_kernel`) and its Python wrappers (`sgl-kernel/python/sgl_kernel/elementwise.py:9-24,
62-70,100-151`) / `CustomOp.forward_cuda` methods (`python/sglang/srt/layers/layernorm.py:46-57`,
`activation.py:46-52`, `rotary_embedding.py:143-166`) call `torch.ops.sgl_kernel.<op>`.  This module
defines ops with the SAME NAMES AND SCHEMAS whose CUDA implementation is libspd (ctypes → C-ABI),
so that a reference checkout whose `sgl_kernel` extension is replaced by

    import semi_pd_b200.torch_ops as t; t.register("sgl_kernel")

runs those layers on the sm_100a kernels with zero edits to its Python.  The ops are always
registered under the `spd` namespace (`torch.ops.spd.rmsnorm(...)`); registering a second namespace
is explicit because a process that has the real sgl_kernel extension loaded already owns it.

There is no CPU implementation: calling an op with host tensors raises (the dispatcher has no
kernel for the CPU key), which is the behaviour of the reference's extension too.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from ._lib import check, lib

# schema strings copied in meaning from torch_extension.cc:49-73 (same argument names / order /
# mutability annotations, which is what `torch.ops.<ns>.<op>(...)` callers depend on)
SCHEMAS = {
    "rmsnorm": "(Tensor! output, Tensor input, Tensor weight, float eps, int cuda_stream) -> ()",
    "fused_add_rmsnorm": "(Tensor! input, Tensor! residual, Tensor weight, float eps) -> ()",
    "silu_and_mul": "(Tensor! out, Tensor input, int cuda_stream) -> ()",
    "apply_rope_pos_ids_cos_sin_cache":
        "(Tensor q, Tensor k, Tensor! q_rope, Tensor! k_rope, Tensor cos_sin_cache, Tensor pos_ids, "
        "bool interleave, int cuda_stream) -> ()",
}

_libs: Dict[str, torch.library.Library] = {}


def _dt(t: torch.Tensor) -> int:
    return ops._dt(t)


def _rmsnorm(output, input, weight, eps, cuda_stream):
    d = input.shape[-1]
    x2, o2 = input.reshape(-1, d), output.view(-1, d)
    if x2.stride(1) != 1 or o2.stride(1) != 1:
        raise RuntimeError("rmsnorm: the last dimension must be contiguous")
    check(lib.spd_rmsnorm(o2.data_ptr(), x2.data_ptr(), weight.data_ptr(), x2.shape[0], d, x2.stride(0),
                          o2.stride(0), float(eps), _dt(input), int(cuda_stream)), "spd_rmsnorm")


def _fused_add_rmsnorm(input, residual, weight, eps):
    ops.fused_add_rmsnorm(input, residual, weight, eps)


def _silu_and_mul(out, input, cuda_stream):
    d = input.shape[-1] // 2
    if not (input.is_contiguous() and out.is_contiguous()):
        raise RuntimeError("silu_and_mul: tensors must be contiguous")
    check(lib.spd_silu_and_mul(out.data_ptr(), input.data_ptr(), input.numel() // (2 * d), d, _dt(input),
                               int(cuda_stream)), "spd_silu_and_mul")


def _apply_rope(q, k, q_rope, k_rope, cos_sin_cache, pos_ids, interleave, cuda_stream):
    # the reference's wrapper passes views of the same storage for q / q_rope and k / k_rope
    # (elementwise.py:142-151): in place.  A distinct output is produced by copy + in place.
    if q_rope.data_ptr() != q.data_ptr():
        q_rope.copy_(q)
    if k_rope.data_ptr() != k.data_ptr():
        k_rope.copy_(k)
    head = q_rope.shape[-1]
    q2, k2 = q_rope.view(q_rope.shape[0], -1), k_rope.view(k_rope.shape[0], -1)
    with torch.cuda.stream(torch.cuda.ExternalStream(int(cuda_stream))) if cuda_stream else _Null():
        ops.apply_rope_with_cos_sin_cache_inplace(pos_ids, q2, k2, head, cos_sin_cache, is_neox=not interleave)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_IMPLS = {"rmsnorm": _rmsnorm, "fused_add_rmsnorm": _fused_add_rmsnorm, "silu_and_mul": _silu_and_mul,
          "apply_rope_pos_ids_cos_sin_cache": _apply_rope}


def register(namespace: str = "spd") -> torch.library.Library:
    """Define the ops of SCHEMAS in `namespace` with libspd as their CUDA kernel.  Idempotent.
    Raises RuntimeError if another library already defined an op of that name there."""
    if namespace in _libs:
        return _libs[namespace]
    m = torch.library.Library(namespace, "FRAGMENT")
    for name, schema in SCHEMAS.items():
        m.define(name + schema)
        m.impl(name, _IMPLS[name], "CUDA")
    _libs[namespace] = m
    return m


register("spd")

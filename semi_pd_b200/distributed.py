"""Tensor-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink 5 /
NVSwitch).  Mirrors the reference's communication_op surface
(python/sglang/srt/distributed/communication_op.py:11-20 → parallel_state.py:376-436): the only
collectives this build issues are the SUM all-reduces after `o_proj` and `down_proj`
(python/sglang/srt/layers/linear.py:1265-1266).  The embedding table is replicated and `lm_head`
is column-sharded with an all-gather of the [bs, vocab/tp] logits, as logits_processor.py:426-427.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

_TP_GROUP: Optional[dist.ProcessGroup] = None
_TP_RANK = 0
_TP_SIZE = 1


def init_distributed_environment(tp_size: int, tp_rank: int, backend: str = "nccl",
                                 init_method: Optional[str] = None, device_id: Optional[int] = None):
    """model_runner.py:285-344 analogue.  With torchrun the env:// rendezvous is used."""
    global _TP_GROUP, _TP_RANK, _TP_SIZE
    _TP_RANK, _TP_SIZE = tp_rank, tp_size
    if tp_size == 1:
        return
    if not dist.is_initialized():
        kwargs = {}
        if backend == "nccl" and device_id is not None:
            kwargs["device_id"] = torch.device("cuda", device_id)
        dist.init_process_group(backend=backend, init_method=init_method or "env://",
                                world_size=tp_size, rank=tp_rank, **kwargs)
    _TP_GROUP = dist.group.WORLD


def new_tp_group(backend: str = "nccl"):
    """A second communicator over the same ranks: the prefill and decode workers each get their
    own so a 128 MiB prefill reduction cannot head-of-line-block a 1 MiB decode reduction
    (the reference builds separate NCCL worlds for P and D, managers/scheduler.py:250-259)."""
    if _TP_SIZE == 1:
        return None
    return dist.new_group(ranks=list(range(_TP_SIZE)), backend=backend)


def get_tensor_model_parallel_world_size() -> int:
    return _TP_SIZE


def get_tensor_model_parallel_rank() -> int:
    return _TP_RANK


def tensor_model_parallel_all_reduce(input_: torch.Tensor, group=None) -> torch.Tensor:
    if _TP_SIZE == 1:
        return input_
    dist.all_reduce(input_, op=dist.ReduceOp.SUM, group=group or _TP_GROUP)
    return input_


def tensor_model_parallel_all_gather(input_: torch.Tensor, dim: int = -1, group=None) -> torch.Tensor:
    if _TP_SIZE == 1:
        return input_
    assert dim in (-1, input_.dim() - 1)
    inp = input_.contiguous()
    flat = torch.empty((_TP_SIZE * inp.shape[0],) + tuple(inp.shape[1:]), dtype=inp.dtype, device=inp.device)
    dist.all_gather_into_tensor(flat, inp, group=group or _TP_GROUP)
    parts = flat.view((_TP_SIZE,) + tuple(inp.shape))
    # [tp, ..., n] -> [..., tp*n]
    return parts.movedim(0, -2).reshape(*input_.shape[:-1], _TP_SIZE * input_.shape[-1])

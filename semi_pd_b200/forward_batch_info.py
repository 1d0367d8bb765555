"""Batch descriptors of the forward path — same names, fields and dtypes as the reference's
`ForwardMode` / `ForwardBatch` (python/sglang/srt/model_executor/forward_batch_info.py:53-332),
`ModelWorkerBatch` (python/sglang/srt/managers/schedule_batch.py:1383-1443) and
`LogitsProcessorOutput` (python/sglang/srt/layers/logits_processor.py:52-90), restricted to the
fields the semi-PD Llama path reads.  Position / start-loc derivation runs in libspd
(`spd_compute_positions`, `spd_decode_positions`) instead of Triton / torch.compile.
"""
from __future__ import annotations

import dataclasses
from enum import IntEnum, auto
from typing import TYPE_CHECKING, List, Optional

import torch

from . import ops

if TYPE_CHECKING:
    from .attention_backend import AttentionBackend
    from .memory_pool import KVCache, ReqToTokenPool
    from .model_runner import ModelRunner


class ForwardMode(IntEnum):
    # forward_batch_info.py:53-110 (the speculative / DP-attention members are kept so that
    # integer values line up with the reference's enum)
    EXTEND = auto()
    DECODE = auto()
    MIXED = auto()
    IDLE = auto()
    TARGET_VERIFY = auto()
    DRAFT_EXTEND = auto()
    DUMMY_FIRST = auto()

    def is_extend(self):
        return self in (ForwardMode.EXTEND, ForwardMode.MIXED, ForwardMode.DRAFT_EXTEND,
                        ForwardMode.TARGET_VERIFY)

    def is_decode(self):
        return self == ForwardMode.DECODE

    def is_mixed(self):
        return self == ForwardMode.MIXED

    def is_idle(self):
        return self == ForwardMode.IDLE

    def is_cuda_graph(self):
        return self in (ForwardMode.DECODE, ForwardMode.TARGET_VERIFY, ForwardMode.IDLE)

    def is_dummy_first(self):
        return self == ForwardMode.DUMMY_FIRST

    def is_decode_or_idle(self):
        return self in (ForwardMode.DECODE, ForwardMode.IDLE)


@dataclasses.dataclass
class ModelWorkerBatch:
    """What the scheduler hands the worker (schedule_batch.py:1383-1443)."""

    bid: int
    forward_mode: ForwardMode
    input_ids: torch.Tensor          # [T] int64
    req_pool_indices: torch.Tensor   # [bs] int64
    seq_lens: torch.Tensor           # [bs] int64
    out_cache_loc: torch.Tensor      # [T] int64
    seq_lens_sum: int
    # decode
    decode_seq_lens: Optional[torch.Tensor] = None
    # extend
    extend_num_tokens: Optional[int] = None
    extend_seq_lens: Optional[List[int]] = None
    extend_prefix_lens: Optional[List[int]] = None
    extend_logprob_start_lens: Optional[List[int]] = None
    # logprobs are unsupported under semi-PD (semi_pd_decode_scheduler.py:354)
    return_logprob: bool = False
    sampling_info: object = None
    # TP > 1: all-gather the vocab-sharded logits (logits_processor.py:426-427).  The semi-PD
    # schedulers sample greedy ids from the (max, argmax) rows and never read the logits.
    gather_logits: bool = True


@dataclasses.dataclass
class LogitsProcessorOutput:
    next_token_logits: torch.Tensor  # [bs, vocab] fp32
    hidden_states: Optional[torch.Tensor] = None
    # greedy ids when the argmax already ran inside the captured decode graph
    next_token_ids: Optional[torch.Tensor] = None


@dataclasses.dataclass
class ForwardBatch:
    """Store all inputs of a forward pass (forward_batch_info.py:130-223)."""

    forward_mode: ForwardMode
    batch_size: int
    input_ids: torch.Tensor
    req_pool_indices: torch.Tensor
    seq_lens: torch.Tensor
    out_cache_loc: torch.Tensor
    seq_lens_sum: int

    return_logprob: bool = False
    gather_logits: bool = True
    positions: torch.Tensor = None
    decode_seq_lens_cpu: Optional[torch.Tensor] = None

    extend_num_tokens: Optional[int] = None
    extend_seq_lens: Optional[torch.Tensor] = None       # int32, device
    extend_prefix_lens: Optional[torch.Tensor] = None    # int32, device
    extend_start_loc: Optional[torch.Tensor] = None      # int32, device
    extend_prefix_lens_cpu: Optional[List[int]] = None
    extend_seq_lens_cpu: Optional[List[int]] = None
    extend_logprob_start_lens_cpu: Optional[List[int]] = None

    sampling_info: object = None

    req_to_token_pool: "ReqToTokenPool" = None
    token_to_kv_pool: "KVCache" = None
    attn_backend: "AttentionBackend" = None

    padded_static_len: int = -1

    @classmethod
    def init_new(cls, batch: ModelWorkerBatch, model_runner: "ModelRunner") -> "ForwardBatch":
        device = model_runner.device
        ret = cls(
            forward_mode=batch.forward_mode,
            batch_size=len(batch.seq_lens),
            input_ids=batch.input_ids,
            req_pool_indices=batch.req_pool_indices,
            seq_lens=batch.seq_lens,
            out_cache_loc=batch.out_cache_loc,
            seq_lens_sum=batch.seq_lens_sum,
            return_logprob=batch.return_logprob,
            gather_logits=batch.gather_logits,
            sampling_info=batch.sampling_info,
            req_to_token_pool=model_runner.req_to_token_pool,
            token_to_kv_pool=model_runner.token_to_kv_pool,
            attn_backend=model_runner.attn_backend,
        )
        if ret.forward_mode.is_idle():
            ret.positions = torch.empty((0,), device=device, dtype=torch.int64)
            return ret
        if ret.forward_mode.is_decode():
            if ret.positions is None:
                ret.positions = ops.clamp_position(batch.seq_lens)
            if ret.decode_seq_lens_cpu is None:
                ret.decode_seq_lens_cpu = batch.decode_seq_lens
        else:
            # pinned staging so the copies are truly asynchronous
            ret.extend_seq_lens = torch.tensor(batch.extend_seq_lens, dtype=torch.int32,
                                               pin_memory=True).to(device, non_blocking=True)
            ret.extend_prefix_lens = torch.tensor(batch.extend_prefix_lens, dtype=torch.int32,
                                                  pin_memory=True).to(device, non_blocking=True)
            ret.extend_num_tokens = batch.extend_num_tokens
            positions, ret.extend_start_loc = ops.compute_positions(
                ret.extend_prefix_lens, ret.extend_seq_lens, ret.extend_num_tokens)
            if ret.positions is None:
                ret.positions = positions
            ret.extend_prefix_lens_cpu = batch.extend_prefix_lens
            ret.extend_seq_lens_cpu = batch.extend_seq_lens
            ret.extend_logprob_start_lens_cpu = batch.extend_logprob_start_lens
        return ret

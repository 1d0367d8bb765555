"""Layers of the Llama forward path with the reference's class and method names
(python/sglang/srt/layers/{layernorm,activation,rotary_embedding,linear,vocab_parallel_embedding,
logits_processor,sampler}.py).  Every `forward` launches libspd kernels; none has a native /
torch fallback.  Plain Python objects (no nn.Module / autograd: this is inference plumbing).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .distributed import (get_current_comm, get_tensor_model_parallel_rank,
                          get_tensor_model_parallel_world_size, tensor_model_parallel_all_gather,
                          tensor_model_parallel_all_reduce)
from .forward_batch_info import ForwardBatch, LogitsProcessorOutput


class RMSNorm:
    """layers/layernorm.py:37-57 (forward_cuda path → rmsnorm / fused_add_rmsnorm, in place)."""

    def __init__(self, hidden_size: int, eps: float = 1e-6):
        self.weight: torch.Tensor = None
        self.hidden_size = hidden_size
        self.variance_epsilon = eps

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None):
        if residual is not None:
            if getattr(x, "_spd_pending_all_reduce", False):
                # x is a RowParallelLinear's partial product whose all-reduce was deferred to here:
                # one kernel reduces it over the ranks, adds the residual and normalises
                x._spd_pending_all_reduce = False
                comm = get_current_comm()
                if comm is not None and comm.all_reduce_add_rmsnorm(x, residual, self.weight,
                                                                    self.variance_epsilon):
                    return x, residual
                tensor_model_parallel_all_reduce(x)
            ops.fused_add_rmsnorm(x, residual, self.weight, self.variance_epsilon)
            return x, residual
        assert not getattr(x, "_spd_pending_all_reduce", False)
        return ops.rmsnorm(x, self.weight, self.variance_epsilon)

    __call__ = forward


class SiluAndMul:
    """layers/activation.py:41-51."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.silu_and_mul(x)

    __call__ = forward


def _compute_inv_freq(rotary_dim: int, base: float) -> torch.Tensor:
    return 1.0 / (base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float) / rotary_dim))


def _llama3_inv_freq(inv_freq: torch.Tensor, scaling: dict) -> torch.Tensor:
    # layers/rotary_embedding.py:773-795
    factor = scaling["factor"]
    low, high = scaling["low_freq_factor"], scaling["high_freq_factor"]
    orig = scaling["original_max_position_embeddings"]
    low_wl, high_wl = orig / low, orig / high
    wave_len = 2 * math.pi / inv_freq
    smooth = (orig / wave_len - low) / (high - low) if low != high else 0
    return torch.where(wave_len < high_wl, inv_freq,
                       torch.where(wave_len > low_wl, inv_freq / factor,
                                   (1 - smooth) * inv_freq / factor + smooth * inv_freq))


class RotaryEmbedding:
    """layers/rotary_embedding.py:61-169: fp32 cos‖sin cache, NeoX or GPT-J style, applied in
    place.  `forward_kvwrite` additionally scatters the rotated K and V into the KV pool in the
    same launch (replaces set_kv_buffer's two index_put_ kernels)."""

    def __init__(self, head_size: int, rotary_dim: int, max_position_embeddings: int, base: float,
                 is_neox_style: bool, dtype: torch.dtype, rope_scaling: Optional[dict] = None,
                 device="cuda"):
        self.head_size = head_size
        self.rotary_dim = rotary_dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        self.is_neox_style = is_neox_style
        self.dtype = dtype
        inv_freq = _compute_inv_freq(rotary_dim, base)
        if rope_scaling is not None:
            kind = rope_scaling.get("rope_type", rope_scaling.get("type"))
            if kind != "llama3":
                raise NotImplementedError(f"rope scaling '{kind}' is not on the semi-PD Llama path")
            inv_freq = _llama3_inv_freq(inv_freq, rope_scaling)
        t = torch.arange(max_position_embeddings, dtype=torch.float)
        freqs = torch.einsum("i,j -> ij", t, inv_freq)
        # the cache stays fp32 on CUDA (rotary_embedding.py:82-86)
        self.cos_sin_cache = torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(device)

    def forward(self, positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        ops.apply_rope_with_cos_sin_cache_inplace(positions, query, key, self.head_size,
                                                  self.cos_sin_cache, self.is_neox_style)
        return query, key

    __call__ = forward

    def forward_kvwrite(self, positions, query, key, value, k_pool, v_pool, out_cache_loc):
        ops.rope_kvwrite(positions, query, key, value, k_pool, v_pool, out_cache_loc, self.head_size,
                         self.cos_sin_cache, self.is_neox_style)
        return query, key


_ROPE_CACHE: Dict[tuple, RotaryEmbedding] = {}


def get_rope(head_size, rotary_dim, max_position, base, is_neox_style=True, rope_scaling=None,
             dtype=torch.bfloat16, device="cuda") -> RotaryEmbedding:
    key = (head_size, rotary_dim, max_position, base, is_neox_style,
           tuple(sorted(rope_scaling.items())) if rope_scaling else None, str(device))
    if key not in _ROPE_CACHE:
        _ROPE_CACHE[key] = RotaryEmbedding(head_size, rotary_dim, max_position, base, is_neox_style,
                                           dtype, rope_scaling, device)
    return _ROPE_CACHE[key]


class LinearBase:
    """weight is [out_per_partition, in_per_partition], row-major (layers/linear.py:140-172)."""

    def __init__(self, input_size: int, output_size: int):
        self.input_size = input_size
        self.output_size = output_size
        self.weight: torch.Tensor = None


class ColumnParallelLinear(LinearBase):
    def forward(self, x: torch.Tensor):
        return ops.linear(x, self.weight), None

    __call__ = forward


class QKVParallelLinear(ColumnParallelLinear):
    """rows = [q_local ‖ k_local ‖ v_local] (layers/linear.py:725-1100)."""

    def __init__(self, hidden_size, head_size, total_num_heads, total_num_kv_heads):
        tp = get_tensor_model_parallel_world_size()
        self.num_heads = total_num_heads // tp
        self.num_kv_heads = max(1, total_num_kv_heads // tp)
        super().__init__(hidden_size, (self.num_heads + 2 * self.num_kv_heads) * head_size)

    @staticmethod
    def shard(full: torch.Tensor, head_size, total_num_heads, total_num_kv_heads, tp_rank, tp_size):
        q, k, v = full.split([total_num_heads * head_size, total_num_kv_heads * head_size,
                              total_num_kv_heads * head_size], dim=0)
        nq = total_num_heads // tp_size
        nkv = max(1, total_num_kv_heads // tp_size)
        rep = max(1, tp_size // total_num_kv_heads)  # KV replication factor when tp > n_kv
        kv_rank = tp_rank // rep
        return torch.cat([q[tp_rank * nq * head_size:(tp_rank + 1) * nq * head_size],
                          k[kv_rank * nkv * head_size:(kv_rank + 1) * nkv * head_size],
                          v[kv_rank * nkv * head_size:(kv_rank + 1) * nkv * head_size]], 0).contiguous()


class MergedColumnParallelLinear(ColumnParallelLinear):
    """rows = [gate_local ‖ up_local] (layers/linear.py:462-722).  `forward_silu_mul` runs the
    GEMM with the SiLU·mul epilogue fused (no [T, 2I] round trip through HBM)."""

    def __init__(self, input_size, output_sizes):
        tp = get_tensor_model_parallel_world_size()
        self.output_sizes = [o // tp for o in output_sizes]
        super().__init__(input_size, sum(self.output_sizes))

    def forward_silu_mul(self, x: torch.Tensor) -> torch.Tensor:
        return ops.linear(x, self.weight, epilogue=ops.SPD_EPI_SILU_MUL)

    @staticmethod
    def shard(full: torch.Tensor, tp_rank, tp_size):
        gate, up = full.chunk(2, dim=0)
        n = gate.shape[0] // tp_size
        return torch.cat([gate[tp_rank * n:(tp_rank + 1) * n], up[tp_rank * n:(tp_rank + 1) * n]],
                         0).contiguous()


class RowParallelLinear(LinearBase):
    """input dim sharded; SUM all-reduce of the partial products (layers/linear.py:1250-1272)."""

    def __init__(self, input_size, output_size, reduce_results: bool = True, defer_reduce: bool = False):
        tp = get_tensor_model_parallel_world_size()
        super().__init__(input_size // tp, output_size)
        self.reduce_results = reduce_results
        # defer_reduce: hand the partial product on, marked; the RMSNorm that consumes it (always
        # a residual-add RMSNorm on the Llama path) runs the all-reduce fused with itself
        self.defer_reduce = defer_reduce

    def forward(self, x: torch.Tensor):
        out = ops.linear(x, self.weight)
        if self.reduce_results and get_tensor_model_parallel_world_size() > 1:
            if self.defer_reduce:
                out._spd_pending_all_reduce = True
                return out, None
            # the calling worker's communicator (distributed.set_current_comm): the peer-memory
            # spd_allreduce (in slot-sized pieces for prefill-sized messages)
            out = tensor_model_parallel_all_reduce(out)
        return out, None

    __call__ = forward

    @staticmethod
    def shard(full: torch.Tensor, tp_rank, tp_size):
        n = full.shape[1] // tp_size
        return full[:, tp_rank * n:(tp_rank + 1) * n].contiguous()


class VocabParallelEmbedding:
    """Replicated table (1-2 GB) instead of vocab-sharding + all-reduce
    (layers/vocab_parallel_embedding.py:469-490) — one gather kernel, no collective."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.weight: torch.Tensor = None

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        return ops.gather_rows(self.weight, input_ids)

    __call__ = forward


class ParallelLMHead(VocabParallelEmbedding):
    """lm_head rows sharded over TP ranks ([vocab/tp, hidden])."""

    @staticmethod
    def shard(full: torch.Tensor, tp_rank, tp_size):
        n = full.shape[0] // tp_size
        return full[tp_rank * n:(tp_rank + 1) * n].contiguous()


class LogitsProcessor:
    """layers/logits_processor.py:220-330,394-446: last-token gather, lm_head matmul in the weight
    dtype, [all-gather,] slice to vocab, `.float()` — the last two fused into the GEMM epilogue.

    TP > 1 (vocab-sharded lm_head): the greedy ids are derived WITHOUT moving the logits — each
    rank reduces its [bs, vocab/tp] shard to one (max, argmax) row per request, the 16-byte rows
    are all-gathered through libspd's peer-memory path (graph-capturable) and merged.  The full
    [bs, vocab] logits of logits_processor.py:426-427 are additionally gathered (NCCL) whenever
    the call is not being captured into a CUDA graph; a captured decode step returns only this
    rank's shard in `next_token_logits` (the scheduler consumes ids)."""

    def __init__(self, config):
        self.config = config
        self.do_tensor_parallel_all_gather = get_tensor_model_parallel_world_size() > 1

    def forward(self, input_ids, hidden_states, lm_head, forward_batch: ForwardBatch
                ) -> LogitsProcessorOutput:
        if forward_batch.forward_mode.is_decode_or_idle():
            pruned_states = hidden_states
        else:
            # last_index = cumsum(extend_seq_lens) - 1  (logits_processor.py:244)
            last_index = forward_batch.extend_start_loc.long() + forward_batch.extend_seq_lens.long() - 1
            pruned_states = ops.gather_rows(hidden_states, last_index)
        logits = ops.linear(pruned_states, lm_head.weight, out_f32=True)
        next_ids = None
        if self.do_tensor_parallel_all_gather:
            comm = get_current_comm()
            shard = logits.shape[1]
            valid = min(shard, max(self.config.vocab_size - get_tensor_model_parallel_rank() * shard, 1))
            pairs = ops.argmax_partial(logits[:, :valid], get_tensor_model_parallel_rank() * shard)
            next_ids = ops.argmax_merge(comm.all_gather_rows16(pairs))
            if forward_batch.gather_logits and not torch.cuda.is_current_stream_capturing():
                logits = tensor_model_parallel_all_gather(logits)
        logits = logits[:, : self.config.vocab_size]
        return LogitsProcessorOutput(next_token_logits=logits, next_token_ids=next_ids)

    __call__ = forward


class Sampler:
    """Greedy branch of layers/sampler.py:73-77,171 (the only one on the semi-PD bench path)."""

    def forward(self, logits_output: LogitsProcessorOutput, sampling_info=None) -> torch.Tensor:
        if sampling_info is not None and not getattr(sampling_info, "is_all_greedy", True):
            raise NotImplementedError("only greedy sampling is on the semi-PD path")
        logits = logits_output.next_token_logits
        if not logits.is_contiguous():
            logits = logits.contiguous()
        return ops.argmax(logits)

    __call__ = forward

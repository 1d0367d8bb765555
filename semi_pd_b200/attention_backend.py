"""Attention plug-in: the reference's `AttentionBackend` ABC
(python/sglang/srt/layers/attention/base_attn_backend.py:14-108), `RadixAttention`
(python/sglang/srt/layers/radix_attention.py:20-70) and ONE backend, `SpdAttnBackend`, in place
of FlashInferAttnBackend (layers/attention/flashinfer_backend.py:62-910).

What disappears relative to the flashinfer backend: there is no `plan()`.  The kernels read
`req_to_token`, `req_pool_indices`, `seq_lens` / `extend_*` directly on the device, pick their
KV split from the launch shape, and merge splits in-kernel — so `init_forward_metadata*` does no
host-device synchronisation and nothing has to be re-planned between CUDA-graph replays.
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Optional

import torch

from . import ops
from .forward_batch_info import ForwardBatch, ForwardMode


class AttentionBackend(ABC):
    """The base class of attention backends"""

    @abstractmethod
    def init_forward_metadata(self, forward_batch: ForwardBatch):
        raise NotImplementedError()

    def init_cuda_graph_state(self, max_bs: int):
        raise NotImplementedError()

    def init_forward_metadata_capture_cuda_graph(self, bs: int, num_tokens: int,
                                                 req_pool_indices: torch.Tensor,
                                                 seq_lens: torch.Tensor,
                                                 encoder_lens: Optional[torch.Tensor],
                                                 forward_mode: ForwardMode, spec_info=None):
        raise NotImplementedError()

    def init_forward_metadata_replay_cuda_graph(self, bs: int, req_pool_indices: torch.Tensor,
                                                seq_lens: torch.Tensor, seq_lens_sum: int,
                                                encoder_lens: Optional[torch.Tensor],
                                                forward_mode: ForwardMode, spec_info=None,
                                                seq_lens_cpu: Optional[torch.Tensor] = None):
        raise NotImplementedError()

    def get_cuda_graph_seq_len_fill_value(self):
        raise NotImplementedError()

    def forward(self, q, k, v, layer: "RadixAttention", forward_batch: ForwardBatch,
                save_kv_cache: bool = True):
        if forward_batch.forward_mode.is_decode():
            return self.forward_decode(q, k, v, layer, forward_batch, save_kv_cache=save_kv_cache)
        return self.forward_extend(q, k, v, layer, forward_batch, save_kv_cache=save_kv_cache)

    def forward_decode(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True):
        raise NotImplementedError()

    def forward_extend(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True):
        raise NotImplementedError()


class RadixAttention:
    """The attention layer descriptor the backend receives (radix_attention.py:26-52)."""

    def __init__(self, num_heads: int, head_dim: int, scaling: float, num_kv_heads: int,
                 layer_id: int, logit_cap: float = 0.0, v_head_dim: int = -1,
                 sliding_window_size: int = -1, is_cross_attention: bool = False, prefix: str = ""):
        self.tp_q_head_num = num_heads
        self.tp_k_head_num = num_kv_heads
        self.tp_v_head_num = num_kv_heads
        self.head_dim = head_dim
        self.qk_head_dim = head_dim
        self.v_head_dim = v_head_dim if v_head_dim != -1 else head_dim
        self.scaling = scaling
        self.layer_id = layer_id
        self.logit_cap = logit_cap
        self.sliding_window_size = sliding_window_size or -1
        self.is_cross_attention = is_cross_attention
        self.k_scale = None
        self.v_scale = None

    def forward(self, q, k, v, forward_batch: ForwardBatch, save_kv_cache: bool = True):
        if k is not None:
            assert v is not None
            k = k.view(-1, self.tp_k_head_num, self.qk_head_dim)
            v = v.view(-1, self.tp_v_head_num, self.v_head_dim)
        return forward_batch.attn_backend.forward(q, k, v, self, forward_batch, save_kv_cache)

    __call__ = forward


class SpdAttnBackend(AttentionBackend):
    """Ragged chunked-prefill attention + batched paged-decode attention in libspd."""

    def __init__(self, model_runner):
        super().__init__()
        self.device = model_runner.device
        self.num_q_heads = model_runner.model_config.num_attention_heads // model_runner.tp_size
        self.head_dim = model_runner.model_config.head_dim
        self.is_mla = getattr(model_runner.model_config, "model_type", "llama") == "deepseek_v2"
        # "tc" = tcgen05 kernels (the product path, head_dim 128); "simt" exists for bring-up /
        # cross-checks and serves the head_dim-64 plumbing model (OPT-125m, BASELINE configs[0])
        self.impl = os.environ.get("SPD_ATTN_IMPL", "tc" if self.head_dim == 128 else "simt")
        self.max_extend_len = 0
        # Split-KV scratch of the decode kernel.  Captured CUDA graphs bake its address in, so it
        # is allocated exactly ONCE here and never replaced: the kernel only splits when
        # batch x kv_heads is below the SM count, i.e. for at most ceil(SMs / kv_heads) requests,
        # which bounds the scratch whatever batch size an eager step later runs with.
        if self.is_mla:
            # absorbed MLA: one latent "kv head"; the split-KV scratch of spd_attn_mla (bounded by
            # 2 x SMs query tokens, see csrc/attn_mla.cu) — allocated once, like the one below
            cfg = model_runner.model_config
            need = ops.lib.spd_attn_mla_workspace_bytes(self.num_q_heads, cfg.kv_lora_rank)
            self.mla_workspace = torch.zeros((need,), dtype=torch.uint8, device=self.device)
            self.decode_workspace = None
            return
        num_kv_heads = model_runner.model_config.get_num_kv_heads(model_runner.tp_size)
        sms = ops.device_sm_count(self.device.index or 0)
        self.split_rows = -(-sms // max(num_kv_heads, 1)) + 1
        need = ops.lib.spd_attn_decode_workspace_bytes(self.split_rows, self.num_q_heads, self.head_dim)
        self.decode_workspace = torch.zeros((need,), dtype=torch.uint8, device=self.device)

    def init_forward_metadata(self, forward_batch: ForwardBatch):
        if forward_batch.forward_mode.is_extend():
            self.max_extend_len = max(forward_batch.extend_seq_lens_cpu)

    def init_cuda_graph_state(self, max_bs: int):
        pass  # the one workspace exists since __init__ and is never reallocated

    def init_forward_metadata_capture_cuda_graph(self, bs, num_tokens, req_pool_indices, seq_lens,
                                                 encoder_lens, forward_mode, spec_info=None):
        assert forward_mode.is_decode(), "only DECODE is captured on the semi-PD path"

    def init_forward_metadata_replay_cuda_graph(self, bs, req_pool_indices, seq_lens, seq_lens_sum,
                                                encoder_lens, forward_mode, spec_info=None,
                                                seq_lens_cpu=None):
        # nothing to re-plan: the captured kernels read seq_lens / req_pool_indices from the
        # graph's static buffers
        pass

    def get_cuda_graph_seq_len_fill_value(self):
        return 1

    def forward_extend(self, q, k, v, layer: RadixAttention, forward_batch: ForwardBatch,
                       save_kv_cache: bool = True):
        pool = forward_batch.token_to_kv_pool
        if save_kv_cache:
            pool.set_kv_buffer(layer, forward_batch.out_cache_loc, k, v)
        k_buf, v_buf = pool.get_kv_buffer(layer.layer_id)
        return ops.attn_prefill(
            q, k_buf, v_buf, forward_batch.req_to_token_pool.req_to_token,
            forward_batch.req_pool_indices, forward_batch.extend_prefix_lens,
            forward_batch.extend_seq_lens, forward_batch.extend_start_loc, self.max_extend_len,
            layer.tp_q_head_num, layer.tp_k_head_num, layer.head_dim, layer.scaling, impl=self.impl)

    def forward_decode(self, q, k, v, layer: RadixAttention, forward_batch: ForwardBatch,
                       save_kv_cache: bool = True):
        pool = forward_batch.token_to_kv_pool
        if save_kv_cache:
            pool.set_kv_buffer(layer, forward_batch.out_cache_loc, k, v)
        k_buf, v_buf = pool.get_kv_buffer(layer.layer_id)
        return ops.attn_decode(
            q, k_buf, v_buf, forward_batch.req_to_token_pool.req_to_token,
            forward_batch.req_pool_indices, forward_batch.seq_lens, layer.tp_q_head_num,
            layer.tp_k_head_num, layer.head_dim, layer.scaling, impl=self.impl,
            workspace=self.decode_workspace)

    def forward_mla(self, q_lat, layer: RadixAttention, forward_batch: ForwardBatch):
        """Absorbed MLA attention over the latent pool (DeepseekV2AttentionMLA.forward_absorb's
        attention call, models/deepseek_v2.py:684-700): q_lat [T, H * (kv_lora + rope)] →
        [T, H * kv_lora].  The latent rows of this batch are already in the pool."""
        pool = forward_batch.token_to_kv_pool
        kv = pool.get_key_buffer(layer.layer_id)
        rope = layer.head_dim - layer.v_head_dim
        common = dict(num_heads=layer.tp_q_head_num, kv_lora_rank=layer.v_head_dim, rope_dim=rope,
                      sm_scale=layer.scaling, workspace=self.mla_workspace)
        if forward_batch.forward_mode.is_decode():
            return ops.attn_mla(q_lat, kv, forward_batch.req_to_token_pool.req_to_token,
                                forward_batch.req_pool_indices, seq_lens=forward_batch.seq_lens, **common)
        return ops.attn_mla(q_lat, kv, forward_batch.req_to_token_pool.req_to_token,
                            forward_batch.req_pool_indices, prefix_lens=forward_batch.extend_prefix_lens,
                            extend_start_loc=forward_batch.extend_start_loc, **common)

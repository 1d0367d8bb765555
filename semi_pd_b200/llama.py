"""Llama wiring — python/sglang/srt/models/llama.py:58-409 with the same class names and call
order.  Differences, all fusions that remove HBM round trips:
  * RoPE + KV-cache write are one kernel (`RotaryEmbedding.forward_kvwrite`); attention is then
    called with save_kv_cache=False;
  * in the prefill regime the gate_up GEMM applies SiLU·mul in its epilogue;
  * q, k, v are strided views of the qkv GEMM output (no split copies).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, Optional, Tuple

import torch

from .attention_backend import RadixAttention
from .distributed import get_tensor_model_parallel_rank, get_tensor_model_parallel_world_size
from .forward_batch_info import ForwardBatch, LogitsProcessorOutput
from .layers import (LogitsProcessor, MergedColumnParallelLinear, ParallelLMHead, QKVParallelLinear,
                     RMSNorm, RowParallelLinear, SiluAndMul, VocabParallelEmbedding, get_rope)

# tokens above which the gate_up GEMM runs in the (non-swapped) prefill regime and can fuse SiLU·mul
# SPD_FUSED_SILU_MIN_TOKENS: A/B knob (a huge value routes prefill through the plain GEMM + SiLU·mul kernel)
_FUSED_SILU_MIN_TOKENS = int(os.environ.get("SPD_FUSED_SILU_MIN_TOKENS", 257))


class LlamaMLP:
    def __init__(self, hidden_size: int, intermediate_size: int, hidden_act: str):
        if hidden_act != "silu":
            raise ValueError(f"Unsupported activation: {hidden_act}. Only silu is supported for now.")
        self.gate_up_proj = MergedColumnParallelLinear(hidden_size, [intermediate_size] * 2)
        self.down_proj = RowParallelLinear(intermediate_size, hidden_size, defer_reduce=True)
        self.act_fn = SiluAndMul()

    def forward(self, x):
        if x.shape[0] >= _FUSED_SILU_MIN_TOKENS:
            x = self.gate_up_proj.forward_silu_mul(x)
        else:
            gate_up, _ = self.gate_up_proj(x)
            x = self.act_fn(gate_up)
        x, _ = self.down_proj(x)
        return x

    __call__ = forward


class LlamaAttention:
    def __init__(self, config, hidden_size: int, num_heads: int, num_kv_heads: int, layer_id: int,
                 rope_theta: float, rope_scaling, max_position_embeddings: int, dtype, device):
        tp_size = get_tensor_model_parallel_world_size()
        self.hidden_size = hidden_size
        self.total_num_heads = num_heads
        assert self.total_num_heads % tp_size == 0
        self.num_heads = self.total_num_heads // tp_size
        self.total_num_kv_heads = num_kv_heads
        if self.total_num_kv_heads >= tp_size:
            assert self.total_num_kv_heads % tp_size == 0
        else:
            assert tp_size % self.total_num_kv_heads == 0
        self.num_kv_heads = max(1, self.total_num_kv_heads // tp_size)
        self.head_dim = getattr(config, "head_dim", hidden_size // num_heads)
        self.q_size = self.num_heads * self.head_dim
        self.kv_size = self.num_kv_heads * self.head_dim
        self.scaling = self.head_dim ** -0.5
        self.qkv_proj = QKVParallelLinear(hidden_size, self.head_dim, self.total_num_heads,
                                          self.total_num_kv_heads)
        self.o_proj = RowParallelLinear(self.total_num_heads * self.head_dim, hidden_size, defer_reduce=True)
        self.rotary_emb = get_rope(self.head_dim, rotary_dim=self.head_dim,
                                   max_position=max_position_embeddings, base=rope_theta,
                                   rope_scaling=rope_scaling, is_neox_style=True, dtype=dtype,
                                   device=device)
        self.attn = RadixAttention(self.num_heads, self.head_dim, self.scaling,
                                   num_kv_heads=self.num_kv_heads, layer_id=layer_id)

    def forward(self, positions: torch.Tensor, hidden_states: torch.Tensor,
                forward_batch: ForwardBatch) -> torch.Tensor:
        qkv, _ = self.qkv_proj(hidden_states)
        q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        k_pool, v_pool = forward_batch.token_to_kv_pool.get_kv_buffer(self.attn.layer_id)
        q, k = self.rotary_emb.forward_kvwrite(positions, q, k, v, k_pool, v_pool,
                                               forward_batch.out_cache_loc)
        attn_output = self.attn(q, k, v, forward_batch, save_kv_cache=False)
        output, _ = self.o_proj(attn_output)
        return output

    __call__ = forward


class LlamaDecoderLayer:
    def __init__(self, config, layer_id: int, dtype, device):
        self.hidden_size = config.hidden_size
        self.self_attn = LlamaAttention(
            config=config, hidden_size=self.hidden_size, num_heads=config.num_attention_heads,
            num_kv_heads=config.num_key_value_heads, layer_id=layer_id,
            rope_theta=getattr(config, "rope_theta", 10000),
            rope_scaling=getattr(config, "rope_scaling", None),
            max_position_embeddings=getattr(config, "max_position_embeddings", 8192),
            dtype=dtype, device=device)
        self.mlp = LlamaMLP(self.hidden_size, config.intermediate_size, config.hidden_act)
        self.input_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, positions, hidden_states, forward_batch: ForwardBatch,
                residual: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        # Self Attention
        if residual is None:
            residual = hidden_states
            hidden_states = self.input_layernorm(hidden_states)
        else:
            hidden_states, residual = self.input_layernorm(hidden_states, residual)
        hidden_states = self.self_attn(positions=positions, hidden_states=hidden_states,
                                       forward_batch=forward_batch)
        # Fully Connected
        hidden_states, residual = self.post_attention_layernorm(hidden_states, residual)
        hidden_states = self.mlp(hidden_states)
        return hidden_states, residual

    __call__ = forward


class LlamaModel:
    def __init__(self, config, dtype, device):
        self.config = config
        self.vocab_size = config.vocab_size
        self.embed_tokens = VocabParallelEmbedding(config.vocab_size, config.hidden_size)
        self.layers = [LlamaDecoderLayer(config, i, dtype, device)
                       for i in range(config.num_hidden_layers)]
        self.norm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, input_ids, positions, forward_batch: ForwardBatch, input_embeds=None):
        hidden_states = self.embed_tokens(input_ids) if input_embeds is None else input_embeds
        residual = None
        for layer in self.layers:
            hidden_states, residual = layer(positions, hidden_states, forward_batch, residual)
        hidden_states, _ = self.norm(hidden_states, residual)
        return hidden_states

    __call__ = forward


class LlamaForCausalLM:
    def __init__(self, config, dtype=torch.bfloat16, device="cuda"):
        self.config = config
        self.dtype = dtype
        self.device = device
        self.model = LlamaModel(config, dtype, device)
        if config.tie_word_embeddings:
            self.lm_head = self.model.embed_tokens
        else:
            self.lm_head = ParallelLMHead(config.vocab_size, config.hidden_size)
        self.logits_processor = LogitsProcessor(config)

    @torch.no_grad()
    def forward(self, input_ids, positions, forward_batch: ForwardBatch, input_embeds=None
                ) -> LogitsProcessorOutput:
        hidden_states = self.model(input_ids, positions, forward_batch, input_embeds)
        return self.logits_processor(input_ids, hidden_states, self.lm_head, forward_batch)

    __call__ = forward

    # ---------------------------------------------------------------------------------------
    # parameters
    # ---------------------------------------------------------------------------------------
    def named_slots(self):
        """(full-checkpoint name, holder object, shard function) for every parameter, in the
        merged layout of models/llama.py (qkv_proj = q‖k‖v rows, gate_up_proj = gate‖up rows)."""
        cfg = self.config
        tp_rank, tp_size = get_tensor_model_parallel_rank(), get_tensor_model_parallel_world_size()
        ident = lambda w: w  # noqa: E731
        yield "model.embed_tokens.weight", self.model.embed_tokens, ident
        for i, layer in enumerate(self.model.layers):
            p = f"model.layers.{i}."
            a = layer.self_attn
            yield p + "input_layernorm.weight", layer.input_layernorm, ident
            yield (p + "self_attn.qkv_proj.weight", a.qkv_proj,
                   lambda w, a=a: QKVParallelLinear.shard(w, a.head_dim, a.total_num_heads,
                                                          a.total_num_kv_heads, tp_rank, tp_size))
            yield (p + "self_attn.o_proj.weight", a.o_proj,
                   lambda w: RowParallelLinear.shard(w, tp_rank, tp_size))
            yield p + "post_attention_layernorm.weight", layer.post_attention_layernorm, ident
            yield (p + "mlp.gate_up_proj.weight", layer.mlp.gate_up_proj,
                   lambda w: MergedColumnParallelLinear.shard(w, tp_rank, tp_size))
            yield (p + "mlp.down_proj.weight", layer.mlp.down_proj,
                   lambda w: RowParallelLinear.shard(w, tp_rank, tp_size))
        yield "model.norm.weight", self.model.norm, ident
        if not cfg.tie_word_embeddings:
            yield ("lm_head.weight", self.lm_head,
                   lambda w: ParallelLMHead.shard(w, tp_rank, tp_size))

    def full_shapes(self) -> Dict[str, Tuple[int, ...]]:
        cfg = self.config
        H, D = cfg.hidden_size, cfg.head_dim
        qkv = (cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * D
        shapes = {"model.embed_tokens.weight": (cfg.vocab_size, H)}
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            shapes[p + "input_layernorm.weight"] = (H,)
            shapes[p + "self_attn.qkv_proj.weight"] = (qkv, H)
            shapes[p + "self_attn.o_proj.weight"] = (H, cfg.num_attention_heads * D)
            shapes[p + "post_attention_layernorm.weight"] = (H,)
            shapes[p + "mlp.gate_up_proj.weight"] = (2 * cfg.intermediate_size, H)
            shapes[p + "mlp.down_proj.weight"] = (H, cfg.intermediate_size)
        shapes["model.norm.weight"] = (H,)
        shapes["lm_head.weight"] = (cfg.vocab_size, H)
        return shapes

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]):
        """Takes FULL (unsharded) tensors by checkpoint name; each rank keeps its TP shard."""
        slots = {name: (holder, shard) for name, holder, shard in self.named_slots()}
        seen = set()
        for name, w in weights:
            if name not in slots:
                if name == "lm_head.weight" and self.config.tie_word_embeddings:
                    continue
                raise KeyError(f"unexpected parameter {name}")
            holder, shard = slots[name]
            holder.weight = shard(w.to(self.dtype)).to(self.device).contiguous()
            seen.add(name)
        missing = set(slots) - seen
        if missing:
            raise KeyError(f"missing parameters: {sorted(missing)[:4]} …")

    def dummy_full_weights(self, seed: int = 1234):
        """Yields (checkpoint name, FULL tensor) of the `--load-format dummy` model: one generator
        per parameter as model_loader/weight_utils.py:588-617 does, but N(0, 1/fan_in) (norm
        weights 1 + N(0, 0.1)) instead of U(±1e-3) so logits are O(1) and greedy decoding is
        well-conditioned.  Generated on the device; a function of (shapes, seed) only, so every TP
        rank — and a test's oracle — sees the same global weights."""
        for idx, (name, shape) in enumerate(self.full_shapes().items()):
            g = torch.Generator(device=self.device)
            g.manual_seed(seed + idx)
            w = torch.randn(shape, generator=g, dtype=torch.float32, device=self.device)
            if len(shape) == 1:
                w = 1.0 + 0.1 * w
            elif name.endswith("embed_tokens.weight"):
                pass
            elif name == "lm_head.weight":
                w = w * (1.5 / math.sqrt(shape[1]))
            else:
                w = w / math.sqrt(shape[1])
            yield name, w.to(self.dtype)
            del w

    def init_dummy_weights(self, seed: int = 1234):
        """Full tensor first, then sharded: every TP rank holds slices of the same global weights."""
        slots = {name: (holder, shard) for name, holder, shard in self.named_slots()}
        for name, w in self.dummy_full_weights(seed):
            if name not in slots:
                continue
            holder, shard = slots[name]
            holder.weight = shard(w).contiguous()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """The local (sharded) parameters by checkpoint name."""
        return {name: holder.weight for name, holder, _ in self.named_slots()}

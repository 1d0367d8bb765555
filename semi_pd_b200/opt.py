"""OPT wiring (BASELINE configs[0]: opt-125m single-prompt greedy decode, plumbing + logit check).

The reference has no OPT model (python/sglang/srt/models/ holds gpt2.py but no opt.py), so the
structure follows transformers' modeling_opt.py (OPTDecoderLayer / OPTDecoder / OPTForCausalLM,
`do_layer_norm_before=True`, learned positions with the +2 offset, ReLU FFN, biases everywhere,
lm_head tied to the token embedding) behind the same ModelRunner / ForwardBatch / AttentionBackend
surface as models/llama.py.  q, k, v projections are merged into one [3H, H] GEMM like the
reference's QKVParallelLinear does for its models.  Attention runs on the SIMT kernels (head_dim
64); every other op is the hand-written CUDA of libspd (GEMM, LayerNorm, bias/ReLU/residual).
Single GPU only (no TP sharding for this plumbing model)."""
from __future__ import annotations

import math
from typing import Dict, Iterable, Tuple

import torch

from . import ops
from .attention_backend import RadixAttention
from .forward_batch_info import ForwardBatch, LogitsProcessorOutput
from .layers import LogitsProcessor


class _Param:
    def __init__(self):
        self.weight: torch.Tensor = None
        self.bias: torch.Tensor = None


class OPTDecoderLayer:
    def __init__(self, config, layer_id: int):
        H = config.hidden_size
        self.embed_dim = H
        self.num_heads = config.num_attention_heads
        self.head_dim = H // self.num_heads
        self.eps = config.layer_norm_eps
        self.qkv_proj, self.out_proj, self.fc1, self.fc2 = _Param(), _Param(), _Param(), _Param()
        self.self_attn_layer_norm, self.final_layer_norm = _Param(), _Param()
        self.attn = RadixAttention(self.num_heads, self.head_dim, self.head_dim ** -0.5,
                                   num_kv_heads=self.num_heads, layer_id=layer_id)

    def forward(self, hidden_states: torch.Tensor, forward_batch: ForwardBatch) -> torch.Tensor:
        H = self.embed_dim
        residual = hidden_states
        x = ops.layernorm(hidden_states, self.self_attn_layer_norm.weight, self.self_attn_layer_norm.bias, self.eps)
        qkv = ops.bias_act_residual(ops.linear(x, self.qkv_proj.weight), self.qkv_proj.bias)
        q, k, v = qkv.split([H, H, H], dim=-1)
        attn = self.attn(q, k, v, forward_batch, save_kv_cache=True)
        hidden_states = ops.bias_act_residual(ops.linear(attn, self.out_proj.weight), self.out_proj.bias, residual)
        residual = hidden_states
        x = ops.layernorm(hidden_states, self.final_layer_norm.weight, self.final_layer_norm.bias, self.eps)
        x = ops.bias_act_residual(ops.linear(x, self.fc1.weight), self.fc1.bias, act="relu")
        return ops.bias_act_residual(ops.linear(x, self.fc2.weight), self.fc2.bias, residual)

    __call__ = forward


class OPTForCausalLM:
    POSITION_OFFSET = 2  # OPTLearnedPositionalEmbedding

    def __init__(self, config, dtype=torch.float16, device="cuda"):
        self.config, self.dtype, self.device = config, dtype, device
        self.embed_tokens, self.embed_positions, self.final_layer_norm = _Param(), _Param(), _Param()
        self.layers = [OPTDecoderLayer(config, i) for i in range(config.num_hidden_layers)]
        self.lm_head = self.embed_tokens  # tied
        self.logits_processor = LogitsProcessor(config)

    @torch.no_grad()
    def forward(self, input_ids, positions, forward_batch: ForwardBatch, input_embeds=None) -> LogitsProcessorOutput:
        tok = ops.gather_rows(self.embed_tokens.weight, input_ids)
        pos = ops.gather_rows(self.embed_positions.weight, positions + self.POSITION_OFFSET)
        h = ops.bias_act_residual(tok, None, pos)
        for layer in self.layers:
            h = layer(h, forward_batch)
        h = ops.layernorm(h, self.final_layer_norm.weight, self.final_layer_norm.bias, self.config.layer_norm_eps)
        return self.logits_processor(input_ids, h, self.lm_head, forward_batch)

    __call__ = forward

    # ------------------------------------------------------------------------------------------
    # parameters (HF checkpoint names; q/k/v merged on load)
    # ------------------------------------------------------------------------------------------
    def full_shapes(self) -> Dict[str, Tuple[int, ...]]:
        cfg = self.config
        H, I = cfg.hidden_size, cfg.intermediate_size
        shapes = {"model.decoder.embed_tokens.weight": (cfg.vocab_size, H),
                  "model.decoder.embed_positions.weight": (cfg.max_position_embeddings + self.POSITION_OFFSET, H)}
        for i in range(cfg.num_hidden_layers):
            p = f"model.decoder.layers.{i}."
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                shapes[p + f"self_attn.{n}.weight"] = (H, H)
                shapes[p + f"self_attn.{n}.bias"] = (H,)
            shapes[p + "self_attn_layer_norm.weight"] = (H,)
            shapes[p + "self_attn_layer_norm.bias"] = (H,)
            shapes[p + "fc1.weight"], shapes[p + "fc1.bias"] = (I, H), (I,)
            shapes[p + "fc2.weight"], shapes[p + "fc2.bias"] = (H, I), (H,)
            shapes[p + "final_layer_norm.weight"] = (H,)
            shapes[p + "final_layer_norm.bias"] = (H,)
        shapes["model.decoder.final_layer_norm.weight"] = (H,)
        shapes["model.decoder.final_layer_norm.bias"] = (H,)
        return shapes

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]):
        w = {k: v for k, v in weights}
        missing = set(self.full_shapes()) - set(w)
        if missing:
            raise KeyError(f"missing parameters: {sorted(missing)[:4]} …")

        def put(t):
            return t.to(self.dtype).to(self.device).contiguous()

        self.embed_tokens.weight = put(w["model.decoder.embed_tokens.weight"])
        self.embed_positions.weight = put(w["model.decoder.embed_positions.weight"])
        self.final_layer_norm.weight = put(w["model.decoder.final_layer_norm.weight"])
        self.final_layer_norm.bias = put(w["model.decoder.final_layer_norm.bias"])
        for i, layer in enumerate(self.layers):
            p = f"model.decoder.layers.{i}."
            layer.qkv_proj.weight = put(torch.cat([w[p + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")]))
            layer.qkv_proj.bias = put(torch.cat([w[p + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")]))
            for ours, theirs in ((layer.out_proj, "self_attn.out_proj"), (layer.fc1, "fc1"), (layer.fc2, "fc2"),
                                 (layer.self_attn_layer_norm, "self_attn_layer_norm"),
                                 (layer.final_layer_norm, "final_layer_norm")):
                ours.weight, ours.bias = put(w[p + theirs + ".weight"]), put(w[p + theirs + ".bias"])

    def init_dummy_weights(self, seed: int = 1234):
        """weights N(0, 0.02), LayerNorm gamma = 1, beta = 0, one generator
        per parameter seeded like model_loader/weight_utils.py:606-617."""
        out = {}
        for idx, (name, shape) in enumerate(self.full_shapes().items()):
            g = torch.Generator().manual_seed(seed + idx)
            if "layer_norm" in name:
                out[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            elif name.endswith("bias"):
                out[name] = torch.zeros(shape)
            else:
                out[name] = torch.randn(shape, generator=g) * 0.02
        self.load_weights(out.items())
        return out

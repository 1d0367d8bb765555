"""DeepSeek-V2(-Lite) wiring — python/sglang/srt/models/deepseek_v2.py with the reference's class
names: `DeepseekV2MLP` (:70-111), `MoEGate` (:114-138), `DeepseekV2MoE` (:141-218),
`DeepseekV2AttentionMLA` (:393-706), `DeepseekV2DecoderLayer` (:843-990), `DeepseekV2Model`,
`DeepseekV2ForCausalLM` (:993-1267).  BASELINE configs[3].

What differs from the reference, all on the kernel side:
  * attention always takes the ABSORBED form (forward_absorb, :633-706): q_nope is multiplied into
    the latent space (w_kc), one MQA pass over the latent pool (`spd_attn_mla`), and the result
    comes back through w_vc.  The reference uses the plain-MHA form for prefill batches without a
    cached prefix (:555-561, 591-631) — the same function of the same inputs; one kernel family
    here serves prefill, chunked prefill and decode.
  * q_proj and kv_a_proj_with_mqa run as ONE GEMM (their weights stacked), the two per-head BMMs
    as one block-diagonal `spd_gemm_grouped` launch each.
  * the routed experts run either as two grouped GEMMs over expert-sorted 128-row tiles
    (prefill-sized batches, and tiny ones that touch few experts) or, for decode-sized batches
    that hit nearly every expert anyway, as two plain GEMMs over ALL experts' stacked weights with
    the routing weights folded into the activation (`spd_moe_dense_act`).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import ops
from ._lib import check, lib
from .attention_backend import RadixAttention
from .distributed import get_tensor_model_parallel_world_size
from .forward_batch_info import ForwardBatch, LogitsProcessorOutput
from .layers import LinearBase, LogitsProcessor, ParallelLMHead, RMSNorm, SiluAndMul, VocabParallelEmbedding

_FUSED_SILU_MIN_TOKENS = 257
# batches in [lo, hi] run all experts densely (see module docstring)
_MOE_DENSE_MIN_TOKENS = 16
_MOE_DENSE_MAX_TOKENS = 256


def yarn_get_mscale(scale: float = 1, mscale: float = 1) -> float:
    """layers/rotary_embedding.py:627-630."""
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


_COS_SIN_CACHE: Dict[tuple, torch.Tensor] = {}


def deepseek_cos_sin_cache(rotary_dim: int, max_position: int, base: float, rope_scaling: Optional[dict],
                           device) -> torch.Tensor:
    """One table shared by all layers (42 MB at 163840 positions)."""
    key = (rotary_dim, max_position, base, tuple(sorted(rope_scaling.items())) if rope_scaling else None, str(device))
    if key not in _COS_SIN_CACHE:
        _COS_SIN_CACHE[key] = _build_cos_sin_cache(rotary_dim, max_position, base, rope_scaling, device)
    return _COS_SIN_CACHE[key]


def _build_cos_sin_cache(rotary_dim: int, max_position: int, base: float, rope_scaling: Optional[dict],
                         device) -> torch.Tensor:
    """fp32 [max_pos, rotary_dim] = cat(cos, sin) — plain RoPE, or DeepseekScalingRotaryEmbedding
    (layers/rotary_embedding.py:633-708): YaRN blend of interpolated / extrapolated frequencies
    over the correction range, cos / sin scaled by mscale(factor, mscale) / mscale(factor, mscale_all_dim)."""
    pos_freqs = base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float32) / rotary_dim)
    mag = 1.0
    if rope_scaling:
        f = float(rope_scaling["factor"])
        orig = rope_scaling["original_max_position_embeddings"]
        beta_fast, beta_slow = rope_scaling.get("beta_fast", 32), rope_scaling.get("beta_slow", 1)

        def corr_dim(rot):
            return (rotary_dim * math.log(orig / (rot * 2 * math.pi))) / (2 * math.log(base))

        low = max(math.floor(corr_dim(beta_fast)), 0)
        high = min(math.ceil(corr_dim(beta_slow)), rotary_dim - 1)
        if low == high:
            high += 0.001
        ramp = torch.clamp((torch.arange(rotary_dim // 2, dtype=torch.float32) - low) / (high - low), 0, 1)
        mask = 1 - ramp
        inv_freq = (1.0 / (f * pos_freqs)) * (1 - mask) + (1.0 / pos_freqs) * mask
        mag = (yarn_get_mscale(f, float(rope_scaling.get("mscale", 1)))
               / yarn_get_mscale(f, float(rope_scaling.get("mscale_all_dim", 0))))
    else:
        inv_freq = 1.0 / pos_freqs
    t = torch.arange(max_position, dtype=torch.float32)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos() * mag, freqs.sin() * mag), dim=-1).to(device)


class DeepseekV2MLP:
    """gate_up (merged) → SiLU·mul → down (:70-111)."""

    def __init__(self, hidden_size: int, intermediate_size: int):
        self.gate_up_proj = LinearBase(hidden_size, 2 * intermediate_size)
        self.down_proj = LinearBase(intermediate_size, hidden_size)
        self.act_fn = SiluAndMul()

    def forward(self, x):
        if x.shape[0] >= _FUSED_SILU_MIN_TOKENS:
            h = ops.linear(x, self.gate_up_proj.weight, epilogue=ops.SPD_EPI_SILU_MUL)
        else:
            h = self.act_fn(ops.linear(x, self.gate_up_proj.weight))
        return ops.linear(h, self.down_proj.weight)

    __call__ = forward


class MoEGate:
    def __init__(self, num_experts: int, hidden_size: int):
        self.weight: torch.Tensor = None  # [E, hidden]

    def forward(self, hidden_states):
        return ops.linear(hidden_states, self.weight)

    __call__ = forward


class DeepseekV2MoE:
    def __init__(self, config):
        if get_tensor_model_parallel_world_size() != 1:
            raise NotImplementedError("the MoE config runs TP = 1 (BASELINE configs[3])")
        self.config = config
        self.n_experts = config.n_routed_experts
        self.top_k = config.num_experts_per_tok
        self.inter = config.moe_intermediate_size
        self.routed_scaling_factor = config.routed_scaling_factor
        self.renormalize = bool(getattr(config, "norm_topk_prob", False))
        self.gate = MoEGate(self.n_experts, config.hidden_size)
        # experts' weights: gate_up stacked along the rows [E * 2I, hidden] (per expert gate rows ||
        # up rows), down concatenated along K [hidden, E * I] — one layout serves both the grouped
        # and the dense path
        self.w_gate_up: torch.Tensor = None
        self.w_down_cat: torch.Tensor = None
        self.shared_experts = (DeepseekV2MLP(config.hidden_size, config.moe_intermediate_size * config.n_shared_experts)
                               if config.n_shared_experts else None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        T = x.shape[0]
        E, inter = self.n_experts, self.inter
        shared = self.shared_experts(x) if self.shared_experts is not None else None
        logits = self.gate(x)
        ids, w = ops.moe_route(logits, self.top_k, scale=self.routed_scaling_factor, renormalize=self.renormalize)
        if _MOE_DENSE_MIN_TOKENS <= T <= _MOE_DENSE_MAX_TOKENS:
            gu = ops.linear(x, self.w_gate_up)                        # [T, E * 2I]
            act = ops.moe_dense_act(gu, ids, w, E, inter)             # [T, E * I], routing weight folded in
            y = ops.linear(act, self.w_down_cat)                      # [T, hidden]
            return ops.bias_act_residual(y, None, shared) if shared is not None else y
        row_token, inv_pos, tile_group = ops.moe_align(ids, E)
        xs = ops.moe_gather(x, row_token)
        act = ops.linear_grouped_rows(xs, self.w_gate_up, tile_group, E, inter, epilogue=ops.SPD_EPI_SILU_MUL)
        y = ops.linear_grouped_rows(act, self.w_down_cat, tile_group, E, x.shape[1], k_concat=True)
        return ops.moe_combine(y, inv_pos, w, shared)

    __call__ = forward


class DeepseekV2AttentionMLA:
    def __init__(self, config, layer_id: int, dtype, device):
        self.layer_id = layer_id
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.qk_nope_head_dim = config.qk_nope_head_dim
        self.qk_rope_head_dim = config.qk_rope_head_dim
        self.qk_head_dim = self.qk_nope_head_dim + self.qk_rope_head_dim
        self.v_head_dim = config.v_head_dim
        self.kv_lora_rank = config.kv_lora_rank
        if getattr(config, "q_lora_rank", None):
            raise NotImplementedError("q_lora_rank (DeepSeek-V2 / V3 full size) is not on the V2-Lite path")
        self.latent_dim = self.kv_lora_rank + self.qk_rope_head_dim
        self.q_size = self.num_heads * self.qk_head_dim
        self.scaling = self.qk_head_dim ** -0.5
        rs = getattr(config, "rope_scaling", None)
        if rs:
            m = yarn_get_mscale(float(rs["factor"]), float(rs.get("mscale_all_dim", 0)))
            self.scaling = self.scaling * m * m  # models/deepseek_v2.py:504-509
        # fused [q_proj ; kv_a_proj_with_mqa] rows
        self.q_kv_a_proj = LinearBase(self.hidden_size, self.q_size + self.latent_dim)
        self.kv_a_layernorm = RMSNorm(self.kv_lora_rank, eps=config.rms_norm_eps)
        self.o_proj = LinearBase(self.num_heads * self.v_head_dim, self.hidden_size)
        self.w_kc_t: torch.Tensor = None   # [H * kv_lora, nope]: head h = rows [h*kv_lora, ...) = W_kc[h]^T
        self.w_vc: torch.Tensor = None     # [H * v_head_dim, kv_lora]
        self.cos_sin_cache = deepseek_cos_sin_cache(self.qk_rope_head_dim, config.max_position_embeddings,
                                                    config.rope_theta, rs, device)
        # RadixAttention descriptor as the reference builds it for the absorbed form (:478-489):
        # one latent "kv head", head_dim = kv_lora + rope, v_head_dim = kv_lora
        self.attn_mqa = RadixAttention(self.num_heads, self.latent_dim, self.scaling, num_kv_heads=1,
                                       layer_id=layer_id, v_head_dim=self.kv_lora_rank)

    def set_kv_b_proj(self, kv_b: torch.Tensor):
        """kv_b_proj.weight [H * (nope + v), kv_lora] → w_kc / w_vc (deepseek_v2.py:1228-1234)."""
        H, R = self.num_heads, self.kv_lora_rank
        w = kv_b.view(H, self.qk_nope_head_dim + self.v_head_dim, R)
        self.w_kc_t = w[:, : self.qk_nope_head_dim, :].transpose(1, 2).contiguous().view(H * R, self.qk_nope_head_dim)
        self.w_vc = w[:, self.qk_nope_head_dim:, :].contiguous().view(H * self.v_head_dim, R)

    def forward(self, positions: torch.Tensor, hidden_states: torch.Tensor, forward_batch: ForwardBatch):
        T = hidden_states.shape[0]
        H, DL, R = self.num_heads, self.latent_dim, self.kv_lora_rank
        qa = ops.linear(hidden_states, self.q_kv_a_proj.weight)          # [T, H*192 + 576]
        q = qa[:, : self.q_size]
        latent = qa[:, self.q_size:]
        # latent row in place: RMSNorm on the kv_lora part, interleaved RoPE on the rope part (and
        # on every head's q_pe) — then it IS the pool row
        ops.rmsnorm(latent[:, :R], self.kv_a_layernorm.weight, self.kv_a_layernorm.variance_epsilon,
                    out=latent[:, :R])
        esz = qa.element_size()
        check(lib.spd_rope_inplace(qa.data_ptr() + self.qk_nope_head_dim * esz,
                                   qa.data_ptr() + (self.q_size + R) * esz, self.cos_sin_cache.data_ptr(),
                                   positions.data_ptr(), T, H, 1, self.qk_head_dim, self.qk_rope_head_dim,
                                   qa.stride(0), qa.stride(0), 1, ops._dt(qa), ops._stream()), "spd_rope_inplace")
        pool = forward_batch.token_to_kv_pool
        pool.set_kv_buffer(self.attn_mqa, forward_batch.out_cache_loc, latent)
        # absorbed query: [q_nope W_kc || q_pe] per head
        q_lat = torch.empty((T, H * DL), device=qa.device, dtype=qa.dtype)
        ops.linear_block_diag(q, self.w_kc_t, H, self.qk_nope_head_dim, R, self.qk_head_dim, out=q_lat,
                              out_group_stride=DL)
        q_lat.view(T, H, DL)[:, :, R:].copy_(q.view(T, H, self.qk_head_dim)[:, :, self.qk_nope_head_dim:])
        o_lat = forward_batch.attn_backend.forward_mla(q_lat, self.attn_mqa, forward_batch)   # [T, H * R]
        attn_out = ops.linear_block_diag(o_lat, self.w_vc, H, R, self.v_head_dim, R)          # [T, H * v]
        return ops.linear(attn_out, self.o_proj.weight)

    __call__ = forward


class DeepseekV2DecoderLayer:
    def __init__(self, config, layer_id: int, dtype, device):
        self.self_attn = DeepseekV2AttentionMLA(config, layer_id, dtype, device)
        self.is_moe = (config.n_routed_experts is not None and layer_id >= config.first_k_dense_replace
                       and layer_id % getattr(config, "moe_layer_freq", 1) == 0)
        self.mlp = DeepseekV2MoE(config) if self.is_moe else DeepseekV2MLP(config.hidden_size, config.intermediate_size)
        self.input_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, positions, hidden_states, forward_batch, residual):
        if residual is None:
            residual = hidden_states
            hidden_states = self.input_layernorm(hidden_states)
        else:
            hidden_states, residual = self.input_layernorm(hidden_states, residual)
        hidden_states = self.self_attn(positions, hidden_states, forward_batch)
        hidden_states, residual = self.post_attention_layernorm(hidden_states, residual)
        hidden_states = self.mlp(hidden_states)
        return hidden_states, residual

    __call__ = forward


class DeepseekV2Model:
    def __init__(self, config, dtype, device):
        self.embed_tokens = VocabParallelEmbedding(config.vocab_size, config.hidden_size)
        self.layers = [DeepseekV2DecoderLayer(config, i, dtype, device) for i in range(config.num_hidden_layers)]
        self.norm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, input_ids, positions, forward_batch):
        hidden_states = self.embed_tokens(input_ids)
        residual = None
        for layer in self.layers:
            hidden_states, residual = layer(positions, hidden_states, forward_batch, residual)
        hidden_states, _ = self.norm(hidden_states, residual)
        return hidden_states

    __call__ = forward


class DeepseekV2ForCausalLM:
    def __init__(self, config, dtype=torch.bfloat16, device="cuda"):
        self.config = config
        self.dtype = dtype
        self.device = device
        self.model = DeepseekV2Model(config, dtype, device)
        self.lm_head = ParallelLMHead(config.vocab_size, config.hidden_size)
        self.logits_processor = LogitsProcessor(config)

    @torch.no_grad()
    def forward(self, input_ids, positions, forward_batch: ForwardBatch, input_embeds=None) -> LogitsProcessorOutput:
        hidden_states = self.model(input_ids, positions, forward_batch)
        return self.logits_processor(input_ids, hidden_states, self.lm_head, forward_batch)

    __call__ = forward

    # ---------------------------------------------------------------------------------------
    # parameters (checkpoint names = transformers' DeepseekV2ForCausalLM / the reference's loader,
    # deepseek_v2.py:1136-1226: experts' gate/up/down arrive stacked per layer)
    # ---------------------------------------------------------------------------------------
    def full_shapes(self) -> Dict[str, Tuple[int, ...]]:
        c = self.config
        H = c.hidden_size
        qk = c.qk_nope_head_dim + c.qk_rope_head_dim
        s = {"model.embed_tokens.weight": (c.vocab_size, H)}
        for i, layer in enumerate(self.model.layers):
            p = f"model.layers.{i}."
            s[p + "input_layernorm.weight"] = (H,)
            s[p + "self_attn.q_proj.weight"] = (c.num_attention_heads * qk, H)
            s[p + "self_attn.kv_a_proj_with_mqa.weight"] = (c.kv_lora_rank + c.qk_rope_head_dim, H)
            s[p + "self_attn.kv_a_layernorm.weight"] = (c.kv_lora_rank,)
            s[p + "self_attn.kv_b_proj.weight"] = (c.num_attention_heads * (c.qk_nope_head_dim + c.v_head_dim), c.kv_lora_rank)
            s[p + "self_attn.o_proj.weight"] = (H, c.num_attention_heads * c.v_head_dim)
            s[p + "post_attention_layernorm.weight"] = (H,)
            if layer.is_moe:
                I = c.moe_intermediate_size
                s[p + "mlp.gate.weight"] = (c.n_routed_experts, H)
                s[p + "mlp.experts.gate_up_proj"] = (c.n_routed_experts, 2 * I, H)
                s[p + "mlp.experts.down_proj"] = (c.n_routed_experts, H, I)
                if c.n_shared_experts:
                    s[p + "mlp.shared_experts.gate_proj.weight"] = (I * c.n_shared_experts, H)
                    s[p + "mlp.shared_experts.up_proj.weight"] = (I * c.n_shared_experts, H)
                    s[p + "mlp.shared_experts.down_proj.weight"] = (H, I * c.n_shared_experts)
            else:
                s[p + "mlp.gate_proj.weight"] = (c.intermediate_size, H)
                s[p + "mlp.up_proj.weight"] = (c.intermediate_size, H)
                s[p + "mlp.down_proj.weight"] = (H, c.intermediate_size)
        s["model.norm.weight"] = (H,)
        s["lm_head.weight"] = (c.vocab_size, H)
        return s

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]):
        """Takes tensors by checkpoint name (see full_shapes) and packs them into the fused layouts."""
        w = {k: v for k, v in weights}
        missing = set(self.full_shapes()) - set(w)
        if missing:
            raise KeyError(f"missing parameters: {sorted(missing)[:4]} …")
        put = lambda t: t.to(self.dtype).to(self.device).contiguous()  # noqa: E731
        self.model.embed_tokens.weight = put(w["model.embed_tokens.weight"])
        self.model.norm.weight = put(w["model.norm.weight"])
        self.lm_head.weight = put(w["lm_head.weight"])
        for i, layer in enumerate(self.model.layers):
            p = f"model.layers.{i}."
            a = layer.self_attn
            layer.input_layernorm.weight = put(w[p + "input_layernorm.weight"])
            layer.post_attention_layernorm.weight = put(w[p + "post_attention_layernorm.weight"])
            a.q_kv_a_proj.weight = put(torch.cat([w[p + "self_attn.q_proj.weight"],
                                                  w[p + "self_attn.kv_a_proj_with_mqa.weight"]], 0))
            a.kv_a_layernorm.weight = put(w[p + "self_attn.kv_a_layernorm.weight"])
            a.set_kv_b_proj(put(w[p + "self_attn.kv_b_proj.weight"]))
            a.o_proj.weight = put(w[p + "self_attn.o_proj.weight"])
            m = layer.mlp
            if layer.is_moe:
                E = self.config.n_routed_experts
                m.gate.weight = put(w[p + "mlp.gate.weight"])
                gu = w[p + "mlp.experts.gate_up_proj"]                       # [E, 2I, H]
                m.w_gate_up = put(gu.reshape(E * gu.shape[1], gu.shape[2]))
                dn = w[p + "mlp.experts.down_proj"]                          # [E, H, I]
                m.w_down_cat = put(dn.permute(1, 0, 2).reshape(dn.shape[1], E * dn.shape[2]))
                if m.shared_experts is not None:
                    m.shared_experts.gate_up_proj.weight = put(torch.cat(
                        [w[p + "mlp.shared_experts.gate_proj.weight"], w[p + "mlp.shared_experts.up_proj.weight"]], 0))
                    m.shared_experts.down_proj.weight = put(w[p + "mlp.shared_experts.down_proj.weight"])
            else:
                m.gate_up_proj.weight = put(torch.cat([w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"]], 0))
                m.down_proj.weight = put(w[p + "mlp.down_proj.weight"])

    def init_dummy_weights(self, seed: int = 1234):
        """`--load-format dummy`: per-parameter seeded N(0, 1/fan_in) (norm weights 1 + N(0, 0.1)),
        generated on the device layer by layer."""
        def gen():
            for idx, (name, shape) in enumerate(self.full_shapes().items()):
                g = torch.Generator(device=self.device)
                g.manual_seed(seed + idx)
                t = torch.randn(shape, generator=g, dtype=torch.float32, device=self.device)
                if len(shape) == 1:
                    t = 1.0 + 0.1 * t
                elif name == "lm_head.weight":
                    t = t * (1.5 / math.sqrt(shape[-1]))
                elif not name.endswith("embed_tokens.weight"):
                    t = t / math.sqrt(shape[-1])
                yield name, t.to(self.dtype)
        self.load_weights(gen())

"""CPU oracle for BASELINE configs[3] (DeepSeek-V2-Lite: MLA attention + MoE) — TEST INFRASTRUCTURE
ONLY, prepared for the next round: no CUDA path exists for this model yet.

Restates, in fp32 torch, what the reference computes (infinigence/Semi-PD @ 2565f9f, paths relative
to python/sglang/srt/):
  * `DeepseekV2AttentionMLA.forward_normal`  (models/deepseek_v2.py:591-632): prefill without a
    prefix = plain MHA with D_qk = nope + rope, D_v = v_head_dim; the LATENT row
    [kv_a_layernorm(c_kv) || rope(k_pe)] (kv_lora_rank + rope dims) is what goes to the KV pool
    (`MLATokenToKVPool`, mem_cache/memory_pool.py:379-452: one [N_tok + 1, 1, 576] buffer per layer);
  * `DeepseekV2AttentionMLA.forward_absorb`  (:634-706): everything else (decode, prefill with a
    prefix) = MQA over the latent pool: q_nope is multiplied into the latent space with
    w_kc = kv_b_proj[:, :nope] (per head), the output comes back with w_vc = kv_b_proj[:, nope:]
    (:1228-1234), K = latent (all dims), V = latent[:kv_lora_rank], softmax scale = D_qk^-0.5
    (x mscale^2 under yarn, :504-509);
  * RoPE on the rope dims only, GPT-J / interleaved style (`is_neox_style=False`, :495-503);
  * `DeepseekV2MoE.forward` (:194-218): softmax router, greedy top-k (n_group = 1 for V2-Lite),
    no renormalisation (`norm_topk_prob` False), x routed_scaling_factor, + shared experts;
    first `first_k_dense_replace` layers are dense MLPs (:870-890).
Pinned by tests/test_deepseek_oracle.py against transformers' `DeepseekV2ForCausalLM` on identical
random weights (the implementation the reference's own generation tests compare against): both
attention paths, ragged prefill, chunked prefill with a cached prefix and paged decode."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class DeepseekOracleConfig:
    vocab_size: int = 102400
    hidden_size: int = 2048
    intermediate_size: int = 10944      # dense layers
    moe_intermediate_size: int = 1408
    num_hidden_layers: int = 27
    num_attention_heads: int = 16
    kv_lora_rank: int = 512
    qk_nope_head_dim: int = 128
    qk_rope_head_dim: int = 64
    v_head_dim: int = 128
    n_routed_experts: int = 64
    num_experts_per_tok: int = 6
    n_shared_experts: int = 2
    first_k_dense_replace: int = 1
    routed_scaling_factor: float = 1.0
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    # {"factor", "beta_fast", "beta_slow", "mscale", "mscale_all_dim", "original_max_position_embeddings"}
    rope_scaling: Optional[dict] = None


def _rms(x, w, eps):
    v = x.float()
    return (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)) * w.float()


def _yarn_get_mscale(scale: float, mscale: float) -> float:
    """rotary_embedding.py:627-630."""
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


def rope_inv_freq(rotary_dim: int, theta: float, rope_scaling: Optional[dict]):
    """-> (inv_freq [rotary_dim/2], cos/sin magnitude).  Plain RoPE, or the reference's
    `DeepseekScalingRotaryEmbedding` (rotary_embedding.py:633-708): YaRN blend of interpolated and
    extrapolated frequencies over the correction range (:377-413), cos / sin scaled by
    mscale(factor, mscale) / mscale(factor, mscale_all_dim)."""
    pos_freqs = theta ** (torch.arange(0, rotary_dim, 2, dtype=torch.float32) / rotary_dim)
    if not rope_scaling:
        return 1.0 / pos_freqs, 1.0
    f = float(rope_scaling["factor"])
    orig = rope_scaling["original_max_position_embeddings"]
    beta_fast, beta_slow = rope_scaling.get("beta_fast", 32), rope_scaling.get("beta_slow", 1)

    def corr_dim(rot):
        return (rotary_dim * math.log(orig / (rot * 2 * math.pi))) / (2 * math.log(theta))

    low, high = max(math.floor(corr_dim(beta_fast)), 0), min(math.ceil(corr_dim(beta_slow)), rotary_dim - 1)
    if low == high:
        high += 0.001
    ramp = torch.clamp((torch.arange(rotary_dim // 2, dtype=torch.float32) - low) / (high - low), 0, 1)
    mask = 1 - ramp
    inv = (1.0 / (f * pos_freqs)) * (1 - mask) + (1.0 / pos_freqs) * mask
    mag = _yarn_get_mscale(f, float(rope_scaling.get("mscale", 1))) / _yarn_get_mscale(
        f, float(rope_scaling.get("mscale_all_dim", 0)))
    return inv, mag


def _rope_interleaved(x, positions, inv_freq, mag=1.0):
    """GPT-J style: pairs (2i, 2i+1) rotate by position * inv_freq[i]  (rotary_embedding.py:88-141
    with is_neox_style=False)."""
    ang = positions.float()[:, None] * inv_freq[None, :]     # [T, d/2]
    cos, sin = (ang.cos() * mag)[:, None, :], (ang.sin() * mag)[:, None, :]  # broadcast over heads
    x1, x2 = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = x1 * cos - x2 * sin
    out[..., 1::2] = x1 * sin + x2 * cos
    return out


class DeepseekOracle:
    def __init__(self, cfg: DeepseekOracleConfig, weights: Dict[str, torch.Tensor], kv_slots: int, max_reqs: int,
                 max_ctx: int):
        self.cfg = cfg
        self.w = {k: v.float() for k, v in weights.items()}
        L = cfg.kv_lora_rank + cfg.qk_rope_head_dim
        # MLATokenToKVPool: one latent row per token per layer, slot 0 reserved
        self.kv = [torch.zeros(kv_slots + 1, 1, L) for _ in range(cfg.num_hidden_layers)]
        self.req_to_token = torch.zeros(max_reqs, max_ctx, dtype=torch.int32)
        self.scaling = (cfg.qk_nope_head_dim + cfg.qk_rope_head_dim) ** -0.5
        self.router_margins: List[torch.Tensor] = []
        self.inv_freq, self.rope_mag = rope_inv_freq(cfg.qk_rope_head_dim, cfg.rope_theta, cfg.rope_scaling)
        if cfg.rope_scaling:
            # models/deepseek_v2.py:504-509: softmax scale x mscale(factor, mscale_all_dim)^2
            m = _yarn_get_mscale(float(cfg.rope_scaling["factor"]), float(cfg.rope_scaling.get("mscale_all_dim", 0)))
            self.scaling = self.scaling * m * m

    # ---- attention ----
    def _attention(self, i, h, positions, req_pool_indices, seq_lens, prefix_lens, extend_lens, loc, force_absorb):
        cfg, w = self.cfg, self.w
        p = f"model.layers.{i}.self_attn."
        H, Dn, Dr, Dv, R = cfg.num_attention_heads, cfg.qk_nope_head_dim, cfg.qk_rope_head_dim, cfg.v_head_dim, cfg.kv_lora_rank
        T = h.shape[0]
        q = F.linear(h, w[p + "q_proj.weight"]).view(T, H, Dn + Dr)
        q_nope, q_pe = q.split([Dn, Dr], dim=-1)
        latent = F.linear(h, w[p + "kv_a_proj_with_mqa.weight"])       # [T, R + Dr]
        c_kv = _rms(latent[:, :R], w[p + "kv_a_layernorm.weight"], cfg.rms_norm_eps)
        k_pe = _rope_interleaved(latent[:, None, R:], positions, self.inv_freq, self.rope_mag)  # [T, 1, Dr]
        q_pe = _rope_interleaved(q_pe, positions, self.inv_freq, self.rope_mag)
        # the pool row (both paths write the same thing): [normed c_kv || roped k_pe]
        self.kv[i][loc] = torch.cat([c_kv[:, None, :], k_pe], dim=-1)
        wkv = w[p + "kv_b_proj.weight"].view(H, Dn + Dv, R)
        no_prefix = prefix_lens is not None and sum(prefix_lens) == 0
        out = torch.empty(T, H, Dv)
        if no_prefix and not force_absorb:
            # forward_normal: MHA over this batch's own tokens
            kv = torch.einsum("tr,hdr->thd", c_kv, wkv)
            k = torch.cat([kv[..., :Dn], k_pe.expand(T, H, Dr)], dim=-1)
            v = kv[..., Dn:]
            qq = torch.cat([q_nope, q_pe], dim=-1)
            s0 = 0
            for n in extend_lens:
                sl = slice(s0, s0 + n)
                sc = torch.einsum("qhd,khd->hqk", qq[sl], k[sl]) * self.scaling
                sc = sc + torch.full((n, n), float("-inf")).triu(1)
                out[sl] = torch.einsum("hqk,khd->qhd", sc.softmax(-1), v[sl])
                s0 += n
        else:
            # forward_absorb: MQA over the latent pool
            w_kc, w_vc = wkv[:, :Dn, :], wkv[:, Dn:, :]                  # [H, Dn, R], [H, Dv, R]
            q_lat = torch.cat([torch.einsum("thd,hdr->thr", q_nope, w_kc), q_pe], dim=-1)  # [T, H, R + Dr]
            s0 = 0
            for b, rp in enumerate(req_pool_indices.tolist()):
                S = int(seq_lens[b])
                n = 1 if extend_lens is None else extend_lens[b]
                pre = S - n
                rows = self.kv[i][self.req_to_token[rp, :S].long(), 0]   # [S, R + Dr]
                sc = torch.einsum("qhl,kl->hqk", q_lat[s0:s0 + n], rows) * self.scaling
                mask = torch.full((n, S), float("-inf")).triu(pre + 1)   # query t sees keys <= pre + t
                o_lat = torch.einsum("hqk,kr->qhr", (sc + mask).softmax(-1), rows[:, :R])
                out[s0:s0 + n] = torch.einsum("qhr,hdr->qhd", o_lat, w_vc)
                s0 += n
        return F.linear(out.reshape(T, H * Dv), w[p + "o_proj.weight"])

    # ---- FFN ----
    def _mlp(self, x, prefix):
        w = self.w
        return F.linear(F.silu(F.linear(x, w[prefix + "gate_proj.weight"])) * F.linear(x, w[prefix + "up_proj.weight"]),
                        w[prefix + "down_proj.weight"])

    def _moe(self, i, x):
        cfg, w = self.cfg, self.w
        p = f"model.layers.{i}.mlp."
        probs = F.linear(x, w[p + "gate.weight"]).softmax(-1)
        topw, topi = probs.topk(cfg.num_experts_per_tok, dim=-1)
        # test aid: how close every token's k-th and (k+1)-th expert are (a 16-bit implementation
        # may legitimately pick the other one when this relative margin is below its rounding noise)
        if probs.shape[-1] > cfg.num_experts_per_tok:
            top = probs.topk(cfg.num_experts_per_tok + 1, dim=-1).values
            self.router_margins.append(((top[:, -2] - top[:, -1]) / top[:, -2]).detach())
        out = torch.zeros_like(x)
        gu, dn = w[p + "experts.gate_up_proj"], w[p + "experts.down_proj"]  # [E, 2I, H], [E, H, I]
        I = cfg.moe_intermediate_size
        for t in range(x.shape[0]):
            for wt, e in zip(topw[t], topi[t]):
                g_u = F.linear(x[t], gu[e])
                out[t] += wt * cfg.routed_scaling_factor * F.linear(F.silu(g_u[:I]) * g_u[I:], dn[e])
        return out + self._mlp(x, p + "shared_experts.")

    # ---- model ----
    def forward(self, input_ids, positions, req_pool_indices, seq_lens, loc, prefix_lens=None, extend_lens=None,
                force_absorb=False):
        """Extend when extend_lens is given (last-token logits per request), else decode."""
        cfg, w = self.cfg, self.w
        self.router_margins = []  # one [T] tensor per MoE layer of this forward (see _moe)
        h = F.embedding(input_ids, w["model.embed_tokens.weight"])
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            a = self._attention(i, _rms(h, w[p + "input_layernorm.weight"], cfg.rms_norm_eps), positions,
                                req_pool_indices, seq_lens, prefix_lens, extend_lens, loc, force_absorb)
            h = h + a
            x = _rms(h, w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
            h = h + (self._mlp(x, p + "mlp.") if i < cfg.first_k_dense_replace else self._moe(i, x))
        h = _rms(h, w["model.norm.weight"], cfg.rms_norm_eps)
        if extend_lens is not None:
            h = h[torch.cumsum(torch.tensor(extend_lens), 0) - 1]
        return F.linear(h, w["lm_head.weight"])


class DeepseekSession:
    """Slot allocator + batch builder (same role as llama_oracle.OracleSession)."""

    def __init__(self, oracle: DeepseekOracle, kv_slots: int, seed: int = 0):
        self.o = oracle
        g = torch.Generator().manual_seed(seed)
        self.free = (torch.randperm(kv_slots, generator=g) + 1).tolist()
        self.lens: Dict[int, int] = {}
        self.next_req = 0

    def _alloc(self, n):
        s, self.free = self.free[:n], self.free[n:]
        return torch.tensor(s, dtype=torch.int64)

    def extend(self, prompts: List[List[int]], req_ids: Optional[List[int]] = None, force_absorb=False):
        if req_ids is None:
            req_ids = list(range(self.next_req, self.next_req + len(prompts)))
            self.next_req += len(prompts)
            for r in req_ids:
                self.lens[r] = 0
        pre = [self.lens[r] for r in req_ids]
        ext = [len(p) for p in prompts]
        loc = self._alloc(sum(ext))
        o = 0
        for r, a, n in zip(req_ids, pre, ext):
            self.o.req_to_token[r, a:a + n] = loc[o:o + n].int()
            self.lens[r] = a + n
            o += n
        pos = torch.cat([torch.arange(a, a + n) for a, n in zip(pre, ext)])
        ids = torch.tensor([t for p in prompts for t in p])
        seq = torch.tensor([a + n for a, n in zip(pre, ext)])
        return req_ids, self.o.forward(ids, pos, torch.tensor(req_ids), seq, loc, pre, ext, force_absorb)

    def decode(self, req_ids: List[int], tokens: List[int]):
        pre = [self.lens[r] for r in req_ids]
        loc = self._alloc(len(req_ids))
        for r, a, s in zip(req_ids, pre, loc.tolist()):
            self.o.req_to_token[r, a] = s
            self.lens[r] = a + 1
        return self.o.forward(torch.tensor(tokens), torch.tensor(pre), torch.tensor(req_ids),
                              torch.tensor([a + 1 for a in pre]), loc)

"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A torch-eager CPU restatement of the reference's Llama forward path (infinigence/Semi-PD @ 2565f9f,
paths below relative to python/sglang/srt/).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; the product package
`semi_pd_b200` never does (tests/test_no_oracle_in_product.py enforces it).

Pinning (see tests/test_oracle_golden.py and tests/golden/make_golden.py):
  * attention (extend + decode), RoPE (native mode) and position computation are checked against
    outputs of the REFERENCE'S OWN modules imported in the build container
    (layers/attention/torch_native_backend.py, layers/rotary_embedding.py,
    model_executor/forward_batch_info.py) — committed as tests/golden/*.pt;
  * RMSNorm / SiLU·mul / RoPE (fp32 mode) follow the closed-form torch references of the
    reference's kernel tests (sgl-kernel/tests/test_norm.py:8-50, test_activation.py:8-16,
    test_rotary_embedding.py:9-128);
  * the end-to-end model wiring (models/llama.py) has no stored golden vector in the reference
    (its own whole-model test compares SRT with the HuggingFace runner and needs downloaded
    weights); it is pinned the same way here: tests/test_oracle_vs_hf.py checks the oracle's
    logits against transformers' LlamaForCausalLM on identical random weights (prefill, ragged
    batch, chunked prefill with a prefix, paged decode, Llama-3.1 rope scaling; max-abs < 2e-4).

Two arithmetic modes:
  * compute="native": every op runs in the model dtype exactly as torch-eager on CPU would
    (F.linear in bf16/fp16 with fp32 accumulation, forward_native ops) — the reference's CPU path;
  * compute="fp32"  : weights/inputs are the same 16-bit values, all arithmetic in fp32 — the
    yardstick for the < 1e-2 max-abs logit tolerance.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# config / synthetic weights
# ------------------------------------------------------------------------------------------------
@dataclass
class OracleLlamaConfig:
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 128
    vocab_size: int = 128256
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    max_position_embeddings: int = 8192

    @staticmethod
    def llama3_8b(**kw):
        return OracleLlamaConfig(**kw)

    @staticmethod
    def tiny(**kw):
        base = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                    num_attention_heads=8, num_key_value_heads=2, head_dim=128 // 4 * 2,
                    vocab_size=1024, max_position_embeddings=2048)
        base.update(kw)
        return OracleLlamaConfig(**base)


def weight_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Merged-parameter layout of models/llama.py: qkv_proj = q||k||v rows, gate_up_proj = gate||up."""
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


def make_weights(cfg, dtype=torch.bfloat16, seed: int = 1234, device="cpu") -> Dict[str, torch.Tensor]:
    """Synthetic weights: one torch.Generator per parameter (model_loader/weight_utils.py:606-617
    seeds each parameter's generator the same way) but N(0, sigma) with fan-in scaling instead of
    U(-1e-3, 1e-3), so that logits are O(1) and greedy argmax is meaningful (SURVEY §7 'hard parts').
    Norm weights are 1 + N(0, 0.1)."""
    out = {}
    for idx, (name, shape) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator(device=device)
        g.manual_seed(seed + idx)
        if len(shape) == 1:
            w = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32, device=device)
        elif name.endswith("embed_tokens.weight"):
            w = torch.randn(shape, generator=g, dtype=torch.float32, device=device)
        elif name == "lm_head.weight":
            w = torch.randn(shape, generator=g, dtype=torch.float32, device=device) * (
                1.5 / math.sqrt(shape[1]))
        else:
            w = torch.randn(shape, generator=g, dtype=torch.float32, device=device) / math.sqrt(shape[1])
        out[name] = w.to(dtype)
    return out


# ------------------------------------------------------------------------------------------------
# per-op restatements
# ------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float,
            residual: Optional[torch.Tensor] = None):
    """layers/layernorm.py:59-76 (RMSNorm.forward_native)."""
    orig_dtype = x.dtype
    x = x.to(torch.float32)
    if residual is not None:
        x = x + residual.to(torch.float32)
        residual = x.to(orig_dtype)
    variance = x.pow(2).mean(dim=-1, keepdim=True)
    x = x * torch.rsqrt(variance + eps)
    x = (x * weight).to(orig_dtype)
    return x if residual is None else (x, residual)


def silu_and_mul(x: torch.Tensor, fp32_math: bool = False) -> torch.Tensor:
    """layers/activation.py:42-44 (SiluAndMul.forward_native).  fp32_math=True is the CUDA kernel's
    arithmetic (sgl-kernel/csrc/elementwise/activation.cu:36-50: fp32 math, one rounding), which
    sgl-kernel/tests/test_activation.py:8-16 holds to 1e-3 of the native form."""
    d = x.shape[-1] // 2
    if fp32_math:
        return (F.silu(x[..., :d].float()) * x[..., d:].float()).to(x.dtype)
    return F.silu(x[..., :d]) * x[..., d:]


def compute_inv_freq(rotary_dim: int, base: float, rope_scaling: Optional[dict] = None) -> torch.Tensor:
    """layers/rotary_embedding.py:88-100; llama3 scaling :773-795."""
    inv_freq = 1.0 / (base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float) / rotary_dim))
    if rope_scaling is None:
        return inv_freq
    kind = rope_scaling.get("rope_type", rope_scaling.get("type"))
    if kind != "llama3":
        raise NotImplementedError(kind)
    factor = rope_scaling["factor"]
    low, high = rope_scaling["low_freq_factor"], rope_scaling["high_freq_factor"]
    orig = rope_scaling["original_max_position_embeddings"]
    low_wl, high_wl = orig / low, orig / high
    wave_len = 2 * math.pi / inv_freq
    if low != high:
        smooth = (orig / wave_len - low) / (high - low)
    else:
        smooth = 0
    return torch.where(wave_len < high_wl, inv_freq,
                       torch.where(wave_len > low_wl, inv_freq / factor,
                                   (1 - smooth) * inv_freq / factor + smooth * inv_freq))


def cos_sin_cache(rotary_dim: int, max_pos: int, base: float, rope_scaling=None) -> torch.Tensor:
    """layers/rotary_embedding.py:102-112 — fp32 [max_pos, rotary_dim] = cat(cos, sin)."""
    inv_freq = compute_inv_freq(rotary_dim, base, rope_scaling)
    t = torch.arange(max_pos, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1)


def apply_rope(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor, head_size: int,
               cache: torch.Tensor, is_neox: bool = True, fp32_math: bool = True):
    """layers/rotary_embedding.py:32-58,113-141.  fp32_math=True: the CUDA path (fp32 cache, fp32
    math, one rounding; sgl-kernel/csrc/elementwise/rope.cu:22-89).  fp32_math=False: the literal
    forward_native (cos/sin cast to the activation dtype, math in that dtype)."""
    rotary_dim = cache.shape[-1]
    num_tokens = positions.numel()
    cos_sin = cache.index_select(0, positions.flatten())
    cos, sin = cos_sin.chunk(2, dim=-1)

    def rot(x):
        shape = x.shape
        x = x.view(num_tokens, -1, head_size)
        x_rot, x_pass = x[..., :rotary_dim], x[..., rotary_dim:]
        if fp32_math:
            xr = x_rot.float()
            c, s = cos.unsqueeze(-2).float(), sin.unsqueeze(-2).float()
        else:
            xr = x_rot
            c, s = cos.unsqueeze(-2).to(x.dtype), sin.unsqueeze(-2).to(x.dtype)
        if is_neox:
            x1, x2 = torch.chunk(xr, 2, dim=-1)
        else:
            x1, x2 = xr[..., ::2], xr[..., 1::2]
        o1 = x1 * c - x2 * s
        o2 = x2 * c + x1 * s
        o = torch.cat((o1, o2), dim=-1) if is_neox else torch.stack((o1, o2), dim=-1).flatten(-2)
        return torch.cat((o.to(x.dtype), x_pass), dim=-1).reshape(shape)

    return rot(query), rot(key)


def compute_position(extend_prefix_lens, extend_seq_lens):
    """model_executor/forward_batch_info.py:449-463 (compute_position_torch)."""
    positions = torch.cat([torch.arange(int(p), int(p) + int(e))
                           for p, e in zip(extend_prefix_lens, extend_seq_lens)], 0)
    extend_seq_lens = torch.as_tensor(extend_seq_lens)
    start = torch.zeros_like(extend_seq_lens)
    start[1:] = torch.cumsum(extend_seq_lens[:-1], dim=0)
    return positions.to(torch.int64), start


def clamp_position(seq_lens):
    """forward_batch_info.py:466-468."""
    return torch.clamp(seq_lens - 1, min=0).to(torch.int64)


def kv_indices(req_to_token, req_pool_indices, lens, kv_start=None):
    """python gather of sgl test/srt/test_create_kvindices.py:41-55 → (kv_indptr, kv_indices)."""
    lens = torch.as_tensor(lens)
    indptr = torch.zeros(len(lens) + 1, dtype=torch.int32)
    indptr[1:] = torch.cumsum(lens, 0)
    parts = []
    for i in range(len(lens)):
        s = int(kv_start[i]) if kv_start is not None else 0
        parts.append(req_to_token[int(req_pool_indices[i]), s:s + int(lens[i])])
    idx = torch.cat(parts).to(torch.int32) if parts else torch.zeros(0, dtype=torch.int32)
    return indptr, idx


def write_req_to_token(req_to_token, req_pool_indices, pre_lens, seq_lens, out_cache_loc):
    """managers/schedule_batch.py:949-961 (the non-Triton loop that the Triton kernel replaces)."""
    pt = 0
    for i in range(len(req_pool_indices)):
        n = int(seq_lens[i]) - int(pre_lens[i])
        req_to_token[int(req_pool_indices[i]), int(pre_lens[i]):int(seq_lens[i])] = \
            out_cache_loc[pt:pt + n].to(req_to_token.dtype)
        pt += n


def attn_extend(q, k_cache, v_cache, req_to_token, req_pool_indices, seq_lens, prefix_lens,
                extend_lens, scaling: float, fp32: bool = False):
    """layers/attention/torch_native_backend.py:26-109: per request, Q padded to the KV length,
    causal SDPA over [prefix || new], keep the last `ext` rows.  q [T, Hq, D] → [T, Hq, D]."""
    out = torch.empty_like(q)
    qh = q.movedim(0, 1)  # [Hq, T, D]
    gqa = q.shape[1] != k_cache.shape[1]
    start = 0
    for i in range(len(seq_lens)):
        ext, pre, kv = int(extend_lens[i]), int(prefix_lens[i]), int(seq_lens[i])
        per_q = qh[:, start:start + ext, :]
        red = torch.zeros((per_q.shape[0], kv, per_q.shape[2]), dtype=per_q.dtype)
        red[:, pre:, :] = per_q
        tok = req_to_token[int(req_pool_indices[i]), :kv].long()
        k = k_cache[tok].movedim(0, 1)
        v = v_cache[tok].movedim(0, 1)
        if fp32:
            red, k, v = red.float(), k.float(), v.float()
        o = F.scaled_dot_product_attention(red.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0),
                                           enable_gqa=gqa, scale=scaling, is_causal=True)
        out[start:start + ext] = o.squeeze(0).movedim(1, 0)[pre:].to(out.dtype)
        start += ext
    return out


def attn_decode(q, k_cache, v_cache, req_to_token, req_pool_indices, seq_lens, scaling: float,
                fp32: bool = False):
    """torch_native_backend.py:111-179: one query per request, non-causal over its seq_len rows."""
    out = torch.empty_like(q)
    qh = q.movedim(0, 1)
    gqa = q.shape[1] != k_cache.shape[1]
    for i in range(len(seq_lens)):
        kv = int(seq_lens[i])
        per_q = qh[:, i:i + 1, :]
        tok = req_to_token[int(req_pool_indices[i]), :kv].long()
        k = k_cache[tok].movedim(0, 1)
        v = v_cache[tok].movedim(0, 1)
        if fp32:
            per_q, k, v = per_q.float(), k.float(), v.float()
        o = F.scaled_dot_product_attention(per_q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0),
                                           enable_gqa=gqa, scale=scaling, is_causal=False)
        out[i:i + 1] = o.squeeze(0).movedim(1, 0).to(out.dtype)
    return out


# ------------------------------------------------------------------------------------------------
# whole model
# ------------------------------------------------------------------------------------------------
@dataclass
class OracleBatch:
    """The fields of ForwardBatch the path reads (model_executor/forward_batch_info.py:130-223)."""
    is_decode: bool
    input_ids: torch.Tensor          # [T] int64
    req_pool_indices: torch.Tensor   # [bs] int64
    seq_lens: torch.Tensor           # [bs] int64 (after this step's tokens)
    out_cache_loc: torch.Tensor      # [T] int64
    extend_prefix_lens: Optional[List[int]] = None
    extend_seq_lens: Optional[List[int]] = None


class _UpcastOnRead(dict):
    """16-bit weights that are widened to fp32 each time they are read: same arithmetic as
    compute="fp32" without holding a second, twice as large copy of an 8B-parameter model."""

    def __getitem__(self, k):
        return dict.__getitem__(self, k).float()


class LlamaOracle:
    """models/llama.py:58-409 + layers/logits_processor.py:244-262,416-442 + layers/sampler.py:73-75,
    over a token-granular KV pool with slot 0 reserved (mem_cache/memory_pool.py:231-250)."""

    def __init__(self, cfg, weights: Dict[str, torch.Tensor], kv_slots: int, max_reqs: int,
                 max_ctx: int, compute: str = "native"):
        assert compute in ("native", "fp32", "fp32_lazy")
        self.cfg, self.compute = cfg, compute
        self.dtype = weights["lm_head.weight"].dtype
        if compute == "fp32_lazy":
            self.w = _UpcastOnRead(weights)
        else:
            self.w = {k: (v.float() if compute == "fp32" else v) for k, v in weights.items()}
        self.act_dtype = self.dtype if compute == "native" else torch.float32
        L, Hk, D = cfg.num_hidden_layers, cfg.num_key_value_heads, cfg.head_dim
        self.k_cache = [torch.zeros(kv_slots + 1, Hk, D, dtype=self.act_dtype) for _ in range(L)]
        self.v_cache = [torch.zeros(kv_slots + 1, Hk, D, dtype=self.act_dtype) for _ in range(L)]
        self.req_to_token = torch.zeros(max_reqs, max_ctx, dtype=torch.int32)
        self.cache = cos_sin_cache(D, cfg.max_position_embeddings, cfg.rope_theta, cfg.rope_scaling)
        self.scaling = D ** -0.5

    def forward(self, b: OracleBatch) -> torch.Tensor:
        """→ next_token_logits fp32 [bs, vocab]."""
        cfg, w = self.cfg, self.w
        Hq, Hk, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        if b.is_decode:
            positions = clamp_position(b.seq_lens)
        else:
            positions, _ = compute_position(b.extend_prefix_lens, b.extend_seq_lens)
        h = F.embedding(b.input_ids, w["model.embed_tokens.weight"])
        residual = None
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            if residual is None:
                residual = h
                h = rmsnorm(h, w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            else:
                h, residual = rmsnorm(h, w[p + "input_layernorm.weight"], cfg.rms_norm_eps, residual)
            qkv = F.linear(h, w[p + "self_attn.qkv_proj.weight"])
            q, k, v = qkv.split([Hq * D, Hk * D, Hk * D], dim=-1)
            q, k = apply_rope(positions, q, k, D, self.cache, True, fp32_math=True)
            # KV write before attention (torch_native_backend.py:195-198)
            self.k_cache[i][b.out_cache_loc] = k.reshape(-1, Hk, D)
            self.v_cache[i][b.out_cache_loc] = v.reshape(-1, Hk, D)
            q3 = q.reshape(-1, Hq, D)
            if b.is_decode:
                o = attn_decode(q3, self.k_cache[i], self.v_cache[i], self.req_to_token,
                                b.req_pool_indices, b.seq_lens, self.scaling)
            else:
                o = attn_extend(q3, self.k_cache[i], self.v_cache[i], self.req_to_token,
                                b.req_pool_indices, b.seq_lens, b.extend_prefix_lens,
                                b.extend_seq_lens, self.scaling)
            h = F.linear(o.reshape(-1, Hq * D), w[p + "self_attn.o_proj.weight"])
            h, residual = rmsnorm(h, w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps, residual)
            gate_up = F.linear(h, w[p + "mlp.gate_up_proj.weight"])
            h = F.linear(silu_and_mul(gate_up), w[p + "mlp.down_proj.weight"])
        if residual is None:  # zero-layer model (bench.py uses it to time embedding + lm_head alone)
            h = rmsnorm(h, w["model.norm.weight"], cfg.rms_norm_eps)
        else:
            h, _ = rmsnorm(h, w["model.norm.weight"], cfg.rms_norm_eps, residual)
        if not b.is_decode:
            last = torch.cumsum(torch.tensor(b.extend_seq_lens, dtype=torch.int64), 0) - 1
            h = h[last]
        logits = torch.matmul(h, w["lm_head.weight"].T)
        return logits[:, : cfg.vocab_size].float()

    @staticmethod
    def sample(logits: torch.Tensor) -> torch.Tensor:
        return torch.argmax(logits, -1).to(torch.int32)


class OracleSession:
    """Minimal allocator + batch builder around LlamaOracle so tests can drive
    extend → decode → decode … the way bench_one_batch.py:229-257 drives the reference."""

    def __init__(self, oracle: LlamaOracle, kv_slots: int, max_reqs: int, seed: int = 0):
        self.o = oracle
        g = torch.Generator().manual_seed(seed)
        self.free = (torch.randperm(kv_slots, generator=g) + 1).tolist()  # scattered slots, never 0
        self.free_reqs = list(range(max_reqs))
        self.lens: Dict[int, int] = {}

    def alloc(self, n: int) -> torch.Tensor:
        s, self.free = self.free[:n], self.free[n:]
        return torch.tensor(s, dtype=torch.int64)

    def extend(self, prompts: List[List[int]], req_ids: Optional[List[int]] = None):
        """Prefill new requests (req_ids None) or continue chunked ones.  → (req_ids, logits)."""
        if req_ids is None:
            req_ids = [self.free_reqs.pop(0) for _ in prompts]
            for r in req_ids:
                self.lens[r] = 0
        pre = [self.lens[r] for r in req_ids]
        ext = [len(p) for p in prompts]
        loc = self.alloc(sum(ext))
        seq = [a + b for a, b in zip(pre, ext)]
        write_req_to_token(self.o.req_to_token, req_ids, pre, seq, loc)
        for r, s in zip(req_ids, seq):
            self.lens[r] = s
        b = OracleBatch(False, torch.tensor([t for p in prompts for t in p], dtype=torch.int64),
                        torch.tensor(req_ids, dtype=torch.int64), torch.tensor(seq, dtype=torch.int64),
                        loc, pre, ext)
        return req_ids, self.o.forward(b), b

    def decode(self, req_ids: List[int], tokens: List[int]):
        pre = [self.lens[r] for r in req_ids]
        seq = [p + 1 for p in pre]
        loc = self.alloc(len(req_ids))
        write_req_to_token(self.o.req_to_token, req_ids, pre, seq, loc)
        for r, s in zip(req_ids, seq):
            self.lens[r] = s
        b = OracleBatch(True, torch.tensor(tokens, dtype=torch.int64),
                        torch.tensor(req_ids, dtype=torch.int64), torch.tensor(seq, dtype=torch.int64), loc)
        return self.o.forward(b), b

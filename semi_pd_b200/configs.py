"""Model / server configuration for the semi-PD forward path.

`ModelConfig` carries the HF-config attributes the reference's models/llama.py reads;
`ServerArgs` keeps the reference's flag names (python/sglang/srt/server_args.py:146,185-331,
870-875) for the knobs that exist on this path; `InstanceRole` and the two SM-percentile
environment variables are python/sglang/semi_pd/utils.py:10-11,25-28.
"""
from __future__ import annotations

import dataclasses
import os
from enum import Enum
from typing import List, Optional

import torch

PREFILL_ENGINE_SM_PERCENTILE = int(os.getenv("SEMI_PD_PREFILL_SM_PERCENTILE", 80))
DECODE_ENGINE_SM_PERCENTILE = int(os.getenv("SEMI_PD_DECODE_SM_PERCENTILE", 100))


class InstanceRole(Enum):
    PREFILL = 0
    DECODE = 1
    OTHER = 2


@dataclasses.dataclass
class ModelConfig:
    name: str = "llama-3-8b"
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
    tie_word_embeddings: bool = False
    hidden_act: str = "silu"
    is_encoder_decoder: bool = False
    # "llama" (models/llama.py) or "opt" (BASELINE configs[0]: pre-LN, learned positions, ReLU, biases)
    model_type: str = "llama"
    layer_norm_eps: float = 1e-5
    # hf_eos_token_id of the reference's ModelConfig (honoured when a request sets ignore_eos=False)
    eos_token_ids: Optional[List[int]] = None
    # DeepSeek-V2 (model_type "deepseek_v2"): MLA + MoE (models/deepseek_v2.py)
    q_lora_rank: Optional[int] = None
    kv_lora_rank: int = 512
    qk_nope_head_dim: int = 128
    qk_rope_head_dim: int = 64
    v_head_dim: int = 128
    n_routed_experts: Optional[int] = None
    num_experts_per_tok: int = 6
    n_shared_experts: int = 2
    moe_intermediate_size: int = 1408
    first_k_dense_replace: int = 1
    moe_layer_freq: int = 1
    routed_scaling_factor: float = 1.0
    norm_topk_prob: bool = False

    @property
    def context_len(self) -> int:
        return self.max_position_embeddings

    def get_num_kv_heads(self, tp_size: int) -> int:
        # models/llama.py:119-127: replicate KV heads when tp > n_kv
        return max(1, self.num_key_value_heads // tp_size)


_PRESETS = {
    "llama-3-8b": dict(eos_token_ids=[128001, 128009]),
    "llama-3.1-8b": dict(
        rope_scaling=dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                          original_max_position_embeddings=8192),
        max_position_embeddings=131072),
    "llama-3-70b": dict(hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                        num_attention_heads=64, num_key_value_heads=8),
    "tiny-llama-test": dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                            num_attention_heads=4, num_key_value_heads=2, head_dim=128,
                            vocab_size=1024, max_position_embeddings=2048),
    # BASELINE configs[0]: facebook/opt-125m (hand-written config, no weights offline)
    "opt-125m": dict(model_type="opt", hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                     num_attention_heads=12, num_key_value_heads=12, head_dim=64, vocab_size=50272,
                     max_position_embeddings=2048, hidden_act="relu", tie_word_embeddings=True),
    # BASELINE configs[3]: deepseek-ai/DeepSeek-V2-Lite (15.7 B parameters, 2.4 B active)
    "deepseek-v2-lite": dict(model_type="deepseek_v2", hidden_size=2048, intermediate_size=10944,
                             num_hidden_layers=27, num_attention_heads=16, num_key_value_heads=16,
                             head_dim=192, vocab_size=102400, rms_norm_eps=1e-6, rope_theta=10000.0,
                             max_position_embeddings=163840, n_routed_experts=64, num_experts_per_tok=6,
                             n_shared_experts=2, moe_intermediate_size=1408, first_k_dense_replace=1,
                             kv_lora_rank=512, qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128,
                             routed_scaling_factor=1.0, norm_topk_prob=False, eos_token_ids=[100001],
                             rope_scaling=dict(type="yarn", factor=40.0, beta_fast=32, beta_slow=1, mscale=0.707,
                                               mscale_all_dim=0.707, original_max_position_embeddings=4096)),
    # one 8B-shaped layer pair: exhaustive parity at real head/FFN shapes
    "llama-3-8b-2layer": dict(num_hidden_layers=2, vocab_size=4096),
}


def get_model_config(name: str, **overrides) -> ModelConfig:
    key = name.lower().split("/")[-1]
    for cand in (key, key.replace("meta-", ""), key.replace("-instruct", "")):
        if cand in _PRESETS:
            kw = dict(_PRESETS[cand])
            kw.update(overrides)
            return ModelConfig(name=cand, **kw)
    raise ValueError(f"unknown model '{name}'; known: {sorted(_PRESETS)}")


def default_cuda_graph_bs(max_bs: int = 160) -> List[int]:
    # model_executor/cuda_graph_runner.py:116-118
    return [b for b in [1, 2, 4] + [i * 8 for i in range(1, 21)] if b <= max_bs]


@dataclasses.dataclass
class ServerArgs:
    model_path: str = "llama-3-8b"
    dtype: str = "bfloat16"
    load_format: str = "dummy"          # synthetic weights; there is no checkpoint loader here
    tp_size: int = 1
    mem_fraction_static: Optional[float] = None
    max_total_tokens: Optional[int] = None
    max_running_requests: Optional[int] = None
    chunked_prefill_size: int = 8192
    max_prefill_tokens: int = 16384
    context_length: Optional[int] = None
    enable_semi_pd: bool = False
    disable_radix_cache: bool = True     # semi-PD forces it off (server_args.py:325-331)
    disable_cuda_graph: bool = False
    cuda_graph_max_bs: int = 160
    cuda_graph_bs: Optional[List[int]] = None
    attention_backend: str = "spd"
    random_seed: int = 1234
    schedule_conservativeness: float = 1.0
    allow_auto_truncate: bool = False
    host: str = "127.0.0.1"
    port: int = 30000
    device: str = "cuda"

    def torch_dtype(self) -> torch.dtype:
        return {"bfloat16": torch.bfloat16, "float16": torch.float16, "half": torch.float16,
                "auto": torch.bfloat16}[self.dtype]

    @staticmethod
    def add_cli_args(parser):
        parser.add_argument("--model-path", type=str, default=ServerArgs.model_path)
        parser.add_argument("--dtype", type=str, default=ServerArgs.dtype)
        parser.add_argument("--load-format", type=str, default=ServerArgs.load_format)
        parser.add_argument("--tp", "--tensor-parallel-size", dest="tp_size", type=int, default=1)
        parser.add_argument("--mem-fraction-static", type=float, default=None)
        parser.add_argument("--max-total-tokens", type=int, default=None)
        parser.add_argument("--max-running-requests", type=int, default=None)
        parser.add_argument("--chunked-prefill-size", type=int, default=ServerArgs.chunked_prefill_size)
        parser.add_argument("--max-prefill-tokens", type=int, default=ServerArgs.max_prefill_tokens)
        parser.add_argument("--context-length", type=int, default=None)
        parser.add_argument("--enable-semi-pd", action="store_true", help="Enable SemiPD.")
        parser.add_argument("--disable-radix-cache", action="store_true")
        parser.add_argument("--disable-cuda-graph", action="store_true")
        parser.add_argument("--cuda-graph-max-bs", type=int, default=ServerArgs.cuda_graph_max_bs)
        parser.add_argument("--random-seed", type=int, default=ServerArgs.random_seed)
        parser.add_argument("--allow-auto-truncate", action="store_true")
        parser.add_argument("--host", type=str, default=ServerArgs.host)
        parser.add_argument("--port", type=int, default=ServerArgs.port)

    @classmethod
    def from_cli_args(cls, ns):
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in vars(ns).items() if k in names})

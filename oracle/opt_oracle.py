"""CPU oracle for BASELINE configs[0] (opt-125m plumbing + logit check) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against the reference: infinigence/Semi-PD has no OPT model
(python/sglang/srt/models/ holds gpt2.py but no opt.py), so there is nothing of the
reference's to restate or to run.  The oracle for this config is HuggingFace transformers'
`OPTForCausalLM` (the implementation any SGLang OPT port would be validated against,
test/srt/models/test_generation_models.py) built from a hand-written `OPTConfig` of
facebook/opt-125m — no weights exist offline, so they are synthetic (N(0, 0.02), LayerNorm 1 / 0,
SURVEY §8d config 1) and shared bit for bit with the CUDA path."""
from __future__ import annotations

from typing import Dict, List

import torch


def opt125m_config(**overrides):
    from transformers import OPTConfig
    kw = dict(vocab_size=50272, hidden_size=768, ffn_dim=3072, num_hidden_layers=12, num_attention_heads=12,
              max_position_embeddings=2048, do_layer_norm_before=True, word_embed_proj_dim=768,
              activation_function="relu", dropout=0.0, attention_dropout=0.0, layerdrop=0.0,
              enable_bias=True, layer_norm_elementwise_affine=True, tie_word_embeddings=True)
    kw.update(overrides)
    return OPTConfig(**kw)


class OptOracle:
    """fp32 arithmetic on the given (16-bit-rounded) weights; full-sequence re-forward per step, so
    it shares no KV-cache logic with the path under test."""

    def __init__(self, hf_config, weights: Dict[str, torch.Tensor]):
        from transformers import OPTForCausalLM
        self.model = OPTForCausalLM(hf_config).float().eval()
        sd = {k: v.float() for k, v in weights.items()}
        sd["lm_head.weight"] = sd["model.decoder.embed_tokens.weight"]
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        assert not unexpected and not missing, (missing, unexpected)

    @torch.no_grad()
    def last_logits(self, seqs: List[List[int]]) -> torch.Tensor:
        return torch.stack([self.model(torch.tensor([s])).logits[0, -1].float() for s in seqs])

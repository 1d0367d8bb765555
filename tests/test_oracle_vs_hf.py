"""Whole-model pin of the CPU oracle: logits of `oracle.llama_oracle.LlamaOracle` (fp32 mode)
against HuggingFace transformers' `LlamaForCausalLM` on IDENTICAL weights and prompts.

The reference stores no golden logits; its own whole-model check is "SRT vs the HF runner"
(test/srt/models/test_generation_models.py:40-48, python/sglang/test/runners.py:736-764, prefill
logprob tolerance 5e-2).  HF is therefore the independent implementation the oracle's layer wiring
(models/llama.py: pre-norm + carried residual, merged q||k||v and gate||up rows, NeoX RoPE, GQA,
last-token logits) is pinned to — prefill, ragged batches, chunked prefill with a prefix, and
token-by-token decode through the paged KV pool, incl. Llama-3.1 rope scaling."""
import pytest
import torch

from oracle import llama_oracle as O

transformers = pytest.importorskip("transformers")


def _hf_model(cfg: O.OracleLlamaConfig, weights):
    from transformers import LlamaConfig, LlamaForCausalLM
    hc = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, hidden_act="silu",
                     max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps,
                     rope_theta=cfg.rope_theta, rope_scaling=cfg.rope_scaling, attention_bias=False, mlp_bias=False,
                     tie_word_embeddings=False, attn_implementation="eager")
    # transformers 5 keeps the rope settings in `rope_parameters`
    if hasattr(hc, "rope_parameters") and cfg.rope_scaling is None:
        hc.rope_parameters = {"rope_type": "default", "rope_theta": cfg.rope_theta}
    m = LlamaForCausalLM(hc).float().eval()
    Hq, Hk, D, I = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size
    sd = {"model.embed_tokens.weight": weights["model.embed_tokens.weight"], "model.norm.weight": weights["model.norm.weight"],
          "lm_head.weight": weights["lm_head.weight"]}
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        q, k, v = weights[p + "self_attn.qkv_proj.weight"].split([Hq * D, Hk * D, Hk * D], dim=0)
        g, u = weights[p + "mlp.gate_up_proj.weight"].split([I, I], dim=0)
        sd.update({p + "self_attn.q_proj.weight": q, p + "self_attn.k_proj.weight": k, p + "self_attn.v_proj.weight": v,
                   p + "self_attn.o_proj.weight": weights[p + "self_attn.o_proj.weight"],
                   p + "mlp.gate_proj.weight": g, p + "mlp.up_proj.weight": u,
                   p + "mlp.down_proj.weight": weights[p + "mlp.down_proj.weight"],
                   p + "input_layernorm.weight": weights[p + "input_layernorm.weight"],
                   p + "post_attention_layernorm.weight": weights[p + "post_attention_layernorm.weight"]})
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return m


LLAMA31_SCALING = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                   "original_max_position_embeddings": 64}


@pytest.mark.parametrize("scaling", [None, LLAMA31_SCALING], ids=["rope", "llama3-rope-scaling"])
def test_oracle_logits_match_hf_llama(scaling):
    torch.manual_seed(0)
    cfg = O.OracleLlamaConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=3, num_attention_heads=8,
                              num_key_value_heads=2, head_dim=16, vocab_size=320, rms_norm_eps=1e-5,
                              rope_theta=10000.0, rope_scaling=scaling, max_position_embeddings=512)
    weights = O.make_weights(cfg, torch.float32, seed=77)
    hf = _hf_model(cfg, weights)
    oracle = O.LlamaOracle(cfg, weights, kv_slots=2048, max_reqs=8, max_ctx=512, compute="fp32")
    sess = O.OracleSession(oracle, 2048, 8, seed=3)
    prompts = [[(7 + 13 * i) % cfg.vocab_size for i in range(37)], [(3 + 5 * i) % cfg.vocab_size for i in range(150)],
               [11]]

    def hf_last_logits(seqs):
        with torch.no_grad():
            return torch.stack([hf(torch.tensor([s])).logits[0, -1].float() for s in seqs])

    # ragged prefill
    rids, logits, _ = sess.extend(prompts)
    ref = hf_last_logits(prompts)
    assert (logits - ref).abs().max().item() < 2e-4
    # greedy decode, token by token through the paged pool
    seqs = [list(p) for p in prompts]
    for _ in range(4):
        tok = O.LlamaOracle.sample(logits).tolist()
        assert tok == ref.argmax(-1).tolist()
        for s, t in zip(seqs, tok):
            s.append(t)
        logits, _ = sess.decode(rids, tok)
        ref = hf_last_logits(seqs)
        assert (logits - ref).abs().max().item() < 2e-4
    # chunked prefill: a new request in two chunks (second chunk attends to the cached prefix)
    long = [(5 + 3 * i) % cfg.vocab_size for i in range(200)]
    r2, _, _ = sess.extend([long[:128]])
    _, logits2, _ = sess.extend([long[128:]], req_ids=r2)
    assert (logits2 - hf_last_logits([long])).abs().max().item() < 2e-4

"""Pins oracle/deepseek_oracle.py (the restatement of the reference's MLA normal / absorbed attention
over a latent KV pool, and its MoE) against transformers' DeepseekV2ForCausalLM on identical random
weights.  CPU only; prepares BASELINE configs[3] for the next round (no CUDA path yet)."""
import pytest
import torch

from oracle import deepseek_oracle as D

transformers = pytest.importorskip("transformers")


YARN = {"rope_type": "yarn", "factor": 8.0, "beta_fast": 32, "beta_slow": 1, "mscale": 1.0, "mscale_all_dim": 0.0,
        "original_max_position_embeddings": 64}


def _models(rope_scaling=None):
    from transformers import DeepseekV2Config, DeepseekV2ForCausalLM
    cfg = D.DeepseekOracleConfig(vocab_size=384, hidden_size=96, intermediate_size=160, moe_intermediate_size=48,
                                 num_hidden_layers=3, num_attention_heads=4, kv_lora_rank=32, qk_nope_head_dim=24,
                                 qk_rope_head_dim=8, v_head_dim=16, n_routed_experts=8, num_experts_per_tok=3,
                                 n_shared_experts=2, first_k_dense_replace=1, routed_scaling_factor=1.0,
                                 max_position_embeddings=512, rope_scaling=rope_scaling)
    hc = DeepseekV2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                          moe_intermediate_size=cfg.moe_intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
                          num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_attention_heads,
                          n_shared_experts=cfg.n_shared_experts, n_routed_experts=cfg.n_routed_experts,
                          num_experts_per_tok=cfg.num_experts_per_tok, kv_lora_rank=cfg.kv_lora_rank, q_lora_rank=None,
                          qk_rope_head_dim=cfg.qk_rope_head_dim, qk_nope_head_dim=cfg.qk_nope_head_dim,
                          v_head_dim=cfg.v_head_dim, first_k_dense_replace=cfg.first_k_dense_replace,
                          max_position_embeddings=cfg.max_position_embeddings, topk_method="greedy", n_group=1,
                          topk_group=1, norm_topk_prob=False, routed_scaling_factor=cfg.routed_scaling_factor,
                          rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, attn_implementation="eager",
                          tie_word_embeddings=False, rope_scaling=rope_scaling)
    torch.manual_seed(5)
    hf = DeepseekV2ForCausalLM(hc).float().eval()
    with torch.no_grad():
        for n, p_ in hf.named_parameters():  # HF initialises experts / norms trivially: randomise everything
            p_.copy_(torch.randn_like(p_) * (0.3 if p_.dim() == 1 else 1.0 / p_.shape[-1] ** 0.5) + (1.0 if "norm" in n else 0.0))
    weights = {k: v.detach().clone() for k, v in hf.state_dict().items()}
    return cfg, hf, weights


@pytest.mark.parametrize("rope_scaling", [None, YARN], ids=["rope", "yarn"])
def test_deepseek_oracle_matches_hf_normal_absorbed_chunked_and_decode(rope_scaling):
    # (YaRN with mscale_all_dim = 0: transformers' port omits the reference's mscale_all_dim^2 softmax
    # factor — models/deepseek_v2.py:504-509 — so that factor is covered by the unit test below only)
    cfg, hf, weights = _models(rope_scaling)

    def hf_last(seqs):
        with torch.no_grad():
            return torch.stack([hf(torch.tensor([s])).logits[0, -1] for s in seqs])

    prompts = [[(7 + 13 * i) % cfg.vocab_size for i in range(29)], [(3 + 5 * i) % cfg.vocab_size for i in range(70)], [11]]
    ref = hf_last(prompts)
    for force_absorb in (False, True):  # forward_normal, then the absorbed MQA over the latent pool
        o = D.DeepseekOracle(cfg, weights, kv_slots=1024, max_reqs=8, max_ctx=256)
        s = D.DeepseekSession(o, 1024, seed=1)
        rids, logits = s.extend(prompts, force_absorb=force_absorb)
        assert (logits - ref).abs().max().item() < 5e-4, force_absorb
    # paged decode continues from the absorbed-path session
    seqs = [list(p) for p in prompts]
    for _ in range(3):
        tok = logits.argmax(-1).tolist()
        assert tok == hf_last(seqs).argmax(-1).tolist()
        for q, t in zip(seqs, tok):
            q.append(t)
        logits = s.decode(rids, tok)
        assert (logits - hf_last(seqs)).abs().max().item() < 5e-4
    # chunked prefill: the second chunk attends to the cached latent prefix (absorbed path)
    long = [(5 + 3 * i) % cfg.vocab_size for i in range(90)]
    r2, _ = s.extend([long[:64]])
    _, l2 = s.extend([long[64:]], req_ids=r2)
    assert (l2 - hf_last([long])).abs().max().item() < 5e-4


def test_yarn_softmax_scale_follows_the_reference_formula():
    import math
    cfg, _, weights = _models({**YARN, "mscale": 0.707, "mscale_all_dim": 0.707, "factor": 40.0,
                               "original_max_position_embeddings": 4096})
    o = D.DeepseekOracle(cfg, weights, kv_slots=64, max_reqs=2, max_ctx=64)
    m = 0.1 * 0.707 * math.log(40.0) + 1.0
    assert abs(o.scaling - (cfg.qk_nope_head_dim + cfg.qk_rope_head_dim) ** -0.5 * m * m) < 1e-9
    assert abs(o.rope_mag - 1.0) < 1e-9  # mscale == mscale_all_dim: cos / sin are not scaled

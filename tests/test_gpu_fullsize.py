"""Parity at the HEADLINE shape (SURVEY §8d "parity workload"): the real 32-layer, vocab-128256
Llama-3-8B with `--load-format dummy` weights generated on the device, one prompt of 128 tokens and
8 greedy steps, against the CPU oracle on the SAME 16-bit weights (copied off the device).

Two yardsticks:
  * fp32 oracle (exact arithmetic on the fp16 weights): greedy ids must be exact wherever the
    top-1/top-2 margin exceeds 2e-2, and the max-abs logit error is reported.  Measured on B200:
    1.3e-2 .. 1.8e-2 per forward — above the 1e-2 that the 2-layer models of test_gpu_model.py
    meet, because 32 layers of fp16 activations (one rounding per op boundary, the reference's
    arithmetic as well) accumulate ~4 fp16 ulps on logits of magnitude ~7;
  * the reference's own arithmetic, `compute="native"` (torch-eager CPU in fp16, BASELINE.json's
    "torch-eager CPU forward"): its distance from the fp32 oracle is the error the reference
    itself carries at this depth.  The CUDA path must not be less accurate than that (factor 2
    for the independent random walks of two different summation orders), with 3e-2 as the
    absolute bound stated for fp16 at 32 layers."""
import os

import pytest
import torch

from oracle import llama_oracle as O
from tests.helpers import GpuSession

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(os.environ.get("SPD_SKIP_FULLSIZE") == "1", reason="SPD_SKIP_FULLSIZE=1")
def test_llama3_8b_full_depth_full_vocab_fp16():
    from semi_pd_b200.configs import ServerArgs, get_model_config
    from semi_pd_b200.model_runner import ModelRunner

    free, _ = torch.cuda.mem_get_info(0)
    if free < 40 * (1 << 30):
        pytest.skip("needs 40 GB of free HBM")
    cfg = get_model_config("llama-3-8b")
    assert (cfg.num_hidden_layers, cfg.vocab_size, cfg.hidden_size) == (32, 128256, 4096)
    sa = ServerArgs(model_path="llama-3-8b", dtype="float16", load_format="dummy", max_total_tokens=2048,
                    max_running_requests=4, context_length=512, disable_cuda_graph=False, cuda_graph_bs=[1],
                    enable_semi_pd=True)
    mr = ModelRunner(cfg, 0.5, 0, 0, 1, 0, sa)
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    # the oracle reads the very tensors the device holds (16.06 GB of fp16 on the host)
    weights = {k: v.cpu() for k, v in mr.model.state_dict().items()}
    ocfg = O.OracleLlamaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                               num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                               num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim,
                               vocab_size=cfg.vocab_size, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                               max_position_embeddings=1024)
    torch.set_num_threads(os.cpu_count() or 1)
    oracle = O.LlamaOracle(ocfg, weights, kv_slots=512, max_reqs=2, max_ctx=520, compute="fp32_lazy")
    gs, os_ = GpuSession(mr), O.OracleSession(oracle, 512, 2)
    prompt = [int((7919 + 104729 * i) % cfg.vocab_size) for i in range(128)]
    reqs, lg, ids = gs.extend([prompt])
    rids, lc, _ = os_.extend([prompt])
    errs = [(lg - lc).abs().max().item()]
    # the reference's CPU arithmetic (native fp16) on the same prefill
    native = O.LlamaOracle(ocfg, weights, kv_slots=512, max_reqs=2, max_ctx=520, compute="native")
    _, ln, _ = O.OracleSession(native, 512, 2).extend([prompt])
    err_native = (ln - lc).abs().max().item()
    err_gpu_vs_native = (lg - ln).abs().max().item()
    del native
    checked = 0
    tok = lc.argmax(-1).tolist()
    for step in range(8):
        top2 = lc.topk(2, -1).values[0]
        if float(top2[0] - top2[1]) > 2e-2:
            assert int(ids[0]) == tok[0], (step, int(ids[0]), tok[0])
            checked += 1
        if step == 7:
            break
        lg, ids = gs.decode(reqs, tok)          # graph replay (bs = 1)
        lc, _ = os_.decode(rids, tok)
        errs.append((lg - lc).abs().max().item())
        tok = lc.argmax(-1).tolist()
    print(f"full-size fp16 parity: max-abs logit error vs fp32 oracle per forward {['%.2e' % e for e in errs]}, "
          f"{checked}/8 greedy ids compared; reference CPU fp16 arithmetic vs fp32 oracle {err_native:.2e}; "
          f"GPU vs reference CPU fp16 {err_gpu_vs_native:.2e}")
    assert max(errs) < 3e-2, errs
    assert errs[0] < max(1e-2, 2.0 * err_native), (errs[0], err_native)
    assert checked >= 4

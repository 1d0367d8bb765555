"""BASELINE configs[0]: opt-125m single-prompt greedy decode — plumbing + logit check.
CUDA path (libspd: GEMM, LayerNorm, bias/ReLU/residual, SIMT attention at head_dim 64, greedy
sampling, paged KV pool) vs the HF-transformers fp32 oracle on identical fp16 weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _runner(cuda_graph: bool):
    from semi_pd_b200.configs import ServerArgs, get_model_config
    from semi_pd_b200.model_runner import ModelRunner
    cfg = get_model_config("opt-125m")
    sa = ServerArgs(model_path="opt-125m", dtype="float16", load_format="external", max_total_tokens=4096,
                    max_running_requests=8, context_length=512, disable_cuda_graph=not cuda_graph,
                    cuda_graph_bs=[1, 2, 4], enable_semi_pd=True)  # defers backend / graph init until the weights are in
    mr = ModelRunner(cfg, 0.5, 0, 0, 1, 0, sa)
    weights = mr.model.init_dummy_weights(1234)
    weights = {k: v.half() for k, v in weights.items()}
    mr.model.load_weights(weights.items())
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    return mr, weights


@pytest.mark.parametrize("cuda_graph", [False, True], ids=["eager", "graph"])
def test_opt125m_greedy_decode_matches_hf(cuda_graph):
    pytest.importorskip("transformers")
    from oracle.opt_oracle import OptOracle, opt125m_config
    from tests.helpers import GpuSession
    mr, weights = _runner(cuda_graph)
    oracle = OptOracle(opt125m_config(), weights)
    gs = GpuSession(mr)
    prompt = [(7 + 13 * i) % 50272 for i in range(16)]
    reqs, logits, ids = gs.extend([prompt])
    seq = list(prompt)
    worst = 0.0
    for step in range(16):
        ref = oracle.last_logits([seq])
        worst = max(worst, (logits - ref).abs().max().item())
        top2 = ref.topk(2, dim=-1).values
        if (top2[0, 0] - top2[0, 1]).item() > 2e-2:
            assert int(ids[0]) == int(ref.argmax(-1)[0]), f"step {step}"
        tok = int(ref.argmax(-1)[0])
        seq.append(tok)
        logits, ids = gs.decode(reqs, [tok])
    assert worst < 1e-2, worst  # fp16, max-abs logit error (north_star tolerance)


def test_opt_elementwise_ops_match_torch():
    from semi_pd_b200 import ops
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    for dtype, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)):
        for T, d in ((1, 768), (37, 768), (5, 3072), (3, 8192)):
            x = torch.randn(T, d).to(dtype)
            g, b = (1 + 0.1 * torch.randn(d)).to(dtype), (0.1 * torch.randn(d)).to(dtype)
            ref = torch.nn.functional.layer_norm(x.float(), (d,), g.float(), b.float(), 1e-5)
            out = ops.layernorm(x.to(dev), g.to(dev), b.to(dev), 1e-5).cpu().float()
            assert (out - ref).abs().max().item() < tol * 4
            r = torch.randn(T, d).to(dtype)
            ref2 = torch.relu(x.float() + b.float()) + r.float()
            out2 = ops.bias_act_residual(x.to(dev), b.to(dev), r.to(dev), act="relu").cpu().float()
            assert (out2 - ref2).abs().max().item() < tol * 4
            out3 = ops.bias_act_residual(x.to(dev), None, r.to(dev)).cpu().float()
            assert (out3 - (x.float() + r.float())).abs().max().item() < tol * 4

"""Per-op time of one Llama-3-8B decoder layer at a prefill batch of n x 1024 tokens, on the
prefill green-context partition (or the whole GPU), CUDA events on the launching stream.
Answers "where does a prefill layer's time go on the partition" (no ncu under green contexts).
usage: python scripts/bench_prefill_ops.py [prefill_pct=50 | 100] [n_requests=8]"""
import sys

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.forward_batch_info import ForwardBatch, ForwardMode, ModelWorkerBatch  # noqa: E402
from semi_pd_b200.model_runner import ModelRunner  # noqa: E402

pct = int(sys.argv[1]) if len(sys.argv) > 1 else 50
NREQ = int(sys.argv[2]) if len(sys.argv) > 2 else 8
EXT = 1024
T = NREQ * EXT
cfg = get_model_config("llama-3-8b")
cfg.num_hidden_layers = 2
total = ops.device_sm_count(0)
if pct < 100:
    from semi_pd_b200.semi_pd_scheduler import split_sms
    p_sms, d_sms0 = split_sms(total, pct, 100 - pct)  # the engine's own split (50 -> 76 + 72)
    part = ops.SmPartition(0, p_sms, d_sms0)
    stream, budget = part.prefill_stream, part.prefill_sms
else:
    stream, budget = torch.cuda.Stream(), 0
print(f"prefill partition: {budget or total} SMs", flush=True)
ops.set_sm_budget(budget)
sa = ServerArgs(model_path="llama-3-8b", load_format="dummy", max_total_tokens=(NREQ + 2) * 1200,
                max_running_requests=NREQ, context_length=1200, enable_semi_pd=True, disable_cuda_graph=True)
with torch.cuda.stream(stream):
    mr = ModelRunner(cfg, 0.8, 0, 0, 1, 0, sa)
    mr.init_attention_backend()
    dev = mr.device
    rtt = mr.req_to_token_pool.req_to_token
    slots = (torch.randperm(mr.max_total_num_tokens - 1, device=dev)[: NREQ * (EXT + 8)] + 1).int()
    rtt[:NREQ, : EXT + 8] = slots.view(NREQ, EXT + 8)
    mwb = ModelWorkerBatch(bid=1, forward_mode=ForwardMode.EXTEND, input_ids=torch.randint(0, 1000, (T,), device=dev),
                           req_pool_indices=torch.arange(NREQ, device=dev),
                           seq_lens=torch.full((NREQ,), EXT, dtype=torch.int64, device=dev),
                           out_cache_loc=rtt[:NREQ, :EXT].reshape(-1).long(), seq_lens_sum=T,
                           extend_num_tokens=T, extend_seq_lens=[EXT] * NREQ, extend_prefix_lens=[0] * NREQ)
    fb = ForwardBatch.init_new(mwb, mr)
    mr.forward(fb)  # builds the attention metadata for fb
    layer = mr.model.model.layers[1]
    att, mlp = layer.self_attn, layer.mlp
    x = (torch.randn(T, cfg.hidden_size, device=dev) * 0.5).to(mr.dtype)
    res = x.clone()
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    state = {}

    def t_op(name, fn, flops=None, byts=None, iters=8):
        for _ in range(2):
            fn()
        tot = 0.0
        for _ in range(iters):
            flush.add_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / iters
        extra = ""
        if flops:
            extra += f"  {flops/ms/1e9:7.1f} TFLOP/s"
        if byts:
            extra += f"  {byts/ms/1e6:7.0f} GB/s"
        print(f"  {name:26s} {ms*1e3:9.1f} us{extra}", flush=True)
        state[name] = ms
        return ms

    H, I = cfg.hidden_size, cfg.intermediate_size
    qkv_n = att.q_size + 2 * att.kv_size
    qkv = att.qkv_proj(x)[0]
    q, k, v = qkv.split([att.q_size, att.kv_size, att.kv_size], dim=-1)
    k_pool, v_pool = fb.token_to_kv_pool.get_kv_buffer(att.attn.layer_id)
    ao = att.attn(q, k, v, fb, save_kv_cache=False)
    h2 = mlp.gate_up_proj.forward_silu_mul(x)
    t_op("fused_add_rmsnorm", lambda: layer.input_layernorm(x, res), byts=4 * T * H * 2)
    t_op("qkv GEMM", lambda: att.qkv_proj(x), flops=2.0 * T * qkv_n * H)
    t_op("rope + kv write", lambda: att.rotary_emb.forward_kvwrite(fb.positions, q, k, v, k_pool, v_pool, fb.out_cache_loc),
         byts=T * (2 * (att.q_size + att.kv_size) + 3 * att.kv_size) * 2)
    t_op("attn_prefill", lambda: att.attn(q, k, v, fb, save_kv_cache=False),
         flops=4.0 * att.num_heads * att.head_dim * NREQ * EXT * (EXT + 1) / 2)
    t_op("o GEMM", lambda: att.o_proj(ao), flops=2.0 * T * H * att.q_size)
    t_op("fused_add_rmsnorm (post)", lambda: layer.post_attention_layernorm(x, res), byts=4 * T * H * 2)
    t_op("gate_up GEMM + SiLU*mul", lambda: mlp.gate_up_proj.forward_silu_mul(x), flops=2.0 * T * 2 * I * H)
    t_op("down GEMM", lambda: mlp.down_proj(h2), flops=2.0 * T * H * I)
    parts = sum(state.values())
    t_op("whole layer", lambda: layer(fb.positions, x, fb, res), flops=T * 13.96e9 / 32 + 4.0 * att.num_heads * att.head_dim * NREQ * EXT * (EXT + 1) / 2)
    print(f"  sum of parts {parts*1e3:9.1f} us", flush=True)
    t_op("whole forward (2 layers + head)", lambda: mr.forward(fb), iters=4)

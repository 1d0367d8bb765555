"""Decode-step time vs batch size for Llama-3-8B (CUDA-graph replay through ModelRunner), on the
decode green-context partition or the whole GPU.  usage: python scripts/bench_decode_step.py [decode_sm_pct]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.forward_batch_info import ForwardBatch, ForwardMode, ModelWorkerBatch  # noqa: E402
from semi_pd_b200.model_runner import ModelRunner  # noqa: E402

pct = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = get_model_config("llama-3-8b")
CTX = 1088
sizes = [1, 8, 16, 32, 64, 96, 128, 160, 192]
sa = ServerArgs(model_path="llama-3-8b", load_format="dummy", max_total_tokens=200 * 1200, max_running_requests=192,
                context_length=1200, enable_semi_pd=True, cuda_graph_bs=sizes, cuda_graph_max_bs=192)
stream = torch.cuda.Stream()
budget = 0
if pct < 100:
    part = ops.SmPartition(0, 148 - (148 * pct // 100 // 8 * 8), 148 * pct // 100 // 8 * 8)
    stream, budget = part.decode_stream, part.decode_sms
    print("decode partition SMs:", budget)
ops.set_sm_budget(budget)
with torch.cuda.stream(stream):
    mr = ModelRunner(cfg, 0.8, 0, 0, 1, 0, sa)
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    dev = mr.device
    for B in sizes:
        rtt = mr.req_to_token_pool.req_to_token
        slots = (torch.randperm(B * (CTX + 40), device=dev)[: B * (CTX + 40)] % (mr.max_total_num_tokens - 1) + 1).int()
        rtt[:B, : CTX + 40] = slots.view(B, CTX + 40)
        mwb = ModelWorkerBatch(bid=0, forward_mode=ForwardMode.DECODE,
                               input_ids=torch.randint(0, 1000, (B,), device=dev),
                               req_pool_indices=torch.arange(B, device=dev),
                               seq_lens=torch.full((B,), CTX, dtype=torch.int64, device=dev),
                               out_cache_loc=rtt[:B, CTX - 1].long(), seq_lens_sum=B * CTX)
        fb = ForwardBatch.init_new(mwb, mr)
        for _ in range(3):
            mr.forward(fb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(20):
            mr.forward(fb)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 20 * 1e3
        ms = e0.elapsed_time(e1) / 20
        by = 15.0e9 + B * CTX * 131072
        print(f"B={B:4d}: {ms:7.3f} ms/step (host wall {wall:6.3f})  {B/ms*1e3:8.0f} tok/s  {by/ms/1e6:6.0f} GB/s of 6554", flush=True)

"""Isolated vs concurrent timing of the two halves of the semi-PD path on their SM partitions
("report every fraction both isolated and concurrent").
  * prefill forward: one 8 x 1024-token extend batch through ModelRunner.forward on the prefill
    partition (eager, as the engine runs it);
  * decode step: one CUDA-graph replay at B requests, context ~1088, on the decode partition.
Each is timed alone (other partition idle) and while the other one loops on its own thread.
usage: python scripts/bench_concurrency.py [prefill_pct=50] [B=64] [partition=green|none]"""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.forward_batch_info import ForwardBatch, ForwardMode, ModelWorkerBatch  # noqa: E402
from semi_pd_b200.model_runner import ModelRunner  # noqa: E402

pct = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
kind = sys.argv[3] if len(sys.argv) > 3 else "green"
cfg = get_model_config("llama-3-8b")
CTX, NREQ, EXT = 1088, 8, 1024
total = ops.device_sm_count(0)
if kind == "green":
    from semi_pd_b200.semi_pd_scheduler import split_sms
    p_sms, d_sms0 = split_sms(total, pct, 100 - pct)  # the engine's own split (50 -> 76 + 72)
    part = ops.SmPartition(0, p_sms, d_sms0)
    p_stream, d_stream, p_sms, d_sms = part.prefill_stream, part.decode_stream, part.prefill_sms, part.decode_sms
else:
    p_stream, d_stream, p_sms, d_sms = torch.cuda.Stream(), torch.cuda.Stream(priority=-1), 0, 0
print(f"partition={kind} prefill SMs={p_sms} decode SMs={d_sms} decode B={B}", flush=True)

sa = ServerArgs(model_path="llama-3-8b", load_format="dummy", max_total_tokens=(B + NREQ + 2) * 1200,
                max_running_requests=B + NREQ, context_length=1200, enable_semi_pd=True, cuda_graph_bs=[B],
                cuda_graph_max_bs=B)
with torch.cuda.stream(d_stream):
    ops.set_sm_budget(d_sms)
    mr = ModelRunner(cfg, 0.8, 0, 0, 1, 0, sa)
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    d_stream.synchronize()
dev = mr.device
rtt = mr.req_to_token_pool.req_to_token
n_slots = mr.max_total_num_tokens - 1
slots = (torch.randperm(n_slots, device=dev)[: (B + NREQ) * (CTX + 40)] + 1).int()
rtt[: B + NREQ, : CTX + 40] = slots.view(B + NREQ, CTX + 40)

d_mwb = ModelWorkerBatch(bid=0, forward_mode=ForwardMode.DECODE, input_ids=torch.randint(0, 1000, (B,), device=dev),
                         req_pool_indices=torch.arange(B, device=dev),
                         seq_lens=torch.full((B,), CTX, dtype=torch.int64, device=dev),
                         out_cache_loc=rtt[:B, CTX - 1].long(), seq_lens_sum=B * CTX)
p_rpi = torch.arange(B, B + NREQ, device=dev)
p_mwb = ModelWorkerBatch(bid=1, forward_mode=ForwardMode.EXTEND,
                         input_ids=torch.randint(0, 1000, (NREQ * EXT,), device=dev), req_pool_indices=p_rpi,
                         seq_lens=torch.full((NREQ,), EXT, dtype=torch.int64, device=dev),
                         out_cache_loc=rtt[B: B + NREQ, :EXT].reshape(-1).long(), seq_lens_sum=NREQ * EXT,
                         extend_num_tokens=NREQ * EXT, extend_seq_lens=[EXT] * NREQ, extend_prefix_lens=[0] * NREQ)

stop = threading.Event()
result = {}


def loop(name, stream, budget, mwb, iters, background):
    torch.cuda.set_device(0)
    ops.set_sm_budget(budget)
    with torch.cuda.stream(stream):
        fb = ForwardBatch.init_new(mwb, mr)
        for _ in range(2):
            mr.forward(fb)
        stream.synchronize()
        if background:
            while not stop.is_set():
                for _ in range(4):
                    mr.forward(fb)
                stream.synchronize()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            mr.forward(fb)
        e1.record(stream)
        e1.synchronize()
        result[name] = e0.elapsed_time(e1) / iters


def run(name, fg, bg=None):
    stop.clear()
    t = None
    if bg is not None:
        t = threading.Thread(target=loop, args=("bg",) + bg + (0, True))
        t.start()
        time.sleep(1.0)
    loop(name, *fg, False)
    stop.set()
    if t is not None:
        t.join()
    torch.cuda.synchronize()


P = (p_stream, p_sms, p_mwb)
D = (d_stream, d_sms, d_mwb)
run("prefill_alone", P + (6,))
run("decode_alone", D + (60,))
run("prefill_concurrent", P + (6,), D)
run("decode_concurrent", D + (60,), P)
flop = NREQ * EXT * 13.96e9 + 32 * 4.0 * 32 * 128 * NREQ * EXT * (EXT + 1) / 2
by = 15.0e9 + B * CTX * 131072
for k in ("prefill_alone", "prefill_concurrent"):
    ms = result[k]
    print(f"{k:20s}: {ms:8.2f} ms / 8x1024 batch  {NREQ*EXT/ms*1e3:8.0f} tok/s  {flop/ms/1e9:7.1f} TFLOP/s", flush=True)
for k in ("decode_alone", "decode_concurrent"):
    ms = result[k]
    print(f"{k:20s}: {ms:8.3f} ms / step B={B}  {B/ms*1e3:8.0f} tok/s  {by/ms/1e6:6.0f} GB/s", flush=True)

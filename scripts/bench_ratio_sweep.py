"""BASELINE.json configs[2]: Llama-3-8B TP=1, prefill:decode SM-ratio sweep under bursty arrivals.
One process per ratio (green contexts live until process exit):
    for p in 20 35 50 65 80; do python scripts/bench_ratio_sweep.py $p; done
Arrivals: bursts of 16 requests, exponential gaps between bursts with mean
16/rate seconds (inter-arrival CV = 4), seed 2; 1024-in / 128-out, ignore_eos.
Prints one JSON line: output tok/s, p50/p99 TTFT and TPOT, the SM split actually granted."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import latency_stats, make_prompts  # noqa: E402
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.semi_pd_scheduler import SemiPDEngine  # noqa: E402

p_pct = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 24.0
n_req = int(sys.argv[3]) if len(sys.argv) > 3 else 192
IN, OUT, BURST = 1024, 128, 16

cfg = get_model_config("llama-3-8b")
ctx = IN + OUT + 8
graph_bs = [b for b in [1, 2, 4] + [8 * i for i in range(1, 33)] if b <= n_req]
sa = ServerArgs(model_path="llama-3-8b", dtype="bfloat16", load_format="dummy", max_total_tokens=4 * n_req * ctx,
                max_running_requests=n_req, context_length=ctx, enable_semi_pd=True, cuda_graph_bs=graph_bs,
                cuda_graph_max_bs=max(graph_bs), chunked_prefill_size=8192, max_prefill_tokens=16384)
engine = SemiPDEngine(sa, gpu_id=0, model_config=cfg, prefill_sm_percent=p_pct, decode_sm_percent=100 - p_pct)
engine.warmup(prompt_len=64, n=2, max_new_tokens=4)
prompts = make_prompts(n_req, IN, cfg.vocab_size)
rng = np.random.RandomState(2)
burst_t = np.cumsum(rng.exponential(BURST / rate, size=(n_req + BURST - 1) // BURST))
arrivals = [float(burst_t[i // BURST]) for i in range(n_req)]

engine.set_device_resident(False)
for w in range(2):  # one untimed window first (allocator, tensor-map cache, graphs)
    t0 = time.perf_counter()
    for i, p in enumerate(prompts):
        wait = t0 + arrivals[i] - time.perf_counter()
        if wait > 0:
            time.sleep(wait)
        engine.submit(p, OUT, arrival_time=t0 + arrivals[i])
    reqs = engine.wait(n_req)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
print(json.dumps({"config": "llama-3-8b TP=1 semi-PD SM-ratio sweep, bursty arrivals (bursts of 16, CV=4, seed 2)",
                  "prefill_pct": p_pct, "prefill_sms": engine.p_sms, "decode_sms": engine.d_sms,
                  "request_rate": rate, "requests": n_req, "window_s": wall,
                  "output_tok_per_s": n_req * OUT / wall, **latency_stats(reqs)}))
engine.shutdown()

"""BASELINE configs[3]: DeepSeek-V2-Lite (MLA + MoE) TP=1 under semi-PD, synthetic requests —
mirrors evaluation/benchmark_deepseek_v2_lite_semi_pd.sh:17-36 of the reference (bench_serving
against a --enable-semi-pd server; ShareGPT there, synthetic here: no datasets offline).  One
saturated window (all requests at t = 0) and one Poisson window at --rate req/s (the reference's
sweep point is 14 req/s).  Prints one JSON line.
usage: python scripts/bench_deepseek.py [--requests 128] [--input-len 1024] [--output-len 128] [--rate 14]"""
import argparse
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import latency_stats, make_prompts, run_window  # noqa: E402
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.semi_pd_scheduler import SemiPDEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--requests", type=int, default=128)
ap.add_argument("--input-len", type=int, default=1024)
ap.add_argument("--output-len", type=int, default=128)
ap.add_argument("--rate", type=float, default=14.0)
ap.add_argument("--sharegpt-like", type=int, default=0,
                help="N > 0: additionally serve N prompts whose lengths follow a fixed-seed log-normal fitted to "
                     "ShareGPT-like statistics (in 220 +- 240, out 190 +- 200 tokens, clipped to [4, 2048]) at "
                     "--rate req/s — SURVEY §8d config 4 (ShareGPT itself is not available offline)")
ap.add_argument("--prefill-sm-pct", type=int, default=50)
ap.add_argument("--decode-sm-pct", type=int, default=50)
a = ap.parse_args()

cfg = get_model_config("deepseek-v2-lite")
ctx = max(a.input_len + a.output_len + 8, 4104 if a.sharegpt_like else 0)
kv_tokens = max(2 * a.requests * (-(-ctx // 128) * 128), 65536)
graph_bs = sorted(set([b for b in [1, 2, 4] + [8 * i for i in range(1, 33)] if b <= max(a.requests, 8)] + [a.requests]))
sa = ServerArgs(model_path="deepseek-v2-lite", dtype="bfloat16", load_format="dummy", max_total_tokens=kv_tokens,
                max_running_requests=a.requests, context_length=ctx, enable_semi_pd=True, cuda_graph_bs=graph_bs,
                cuda_graph_max_bs=max(graph_bs), chunked_prefill_size=8192, max_prefill_tokens=16384)
t0 = time.perf_counter()
eng = SemiPDEngine(sa, gpu_id=0, model_config=cfg, prefill_sm_percent=a.prefill_sm_pct,
                   decode_sm_percent=a.decode_sm_pct)
setup = time.perf_counter() - t0
eng.warmup(prompt_len=64, n=2, max_new_tokens=4)
prompts = make_prompts(a.requests, a.input_len, cfg.vocab_size)
run_window(eng, prompts, None, a.output_len, float("inf"), False)  # warm-up window
w, reqs, _ = run_window(eng, prompts, None, a.output_len, float("inf"), False)
sat = {"output_tok_per_s": a.requests * a.output_len / w, "ms_per_window": 1e3 * w, **latency_stats(reqs)}
w2, reqs2, _ = run_window(eng, prompts, None, a.output_len, a.rate, False)
at_rate = {"request_rate": a.rate, "output_tok_per_s": len(reqs2) * a.output_len / w2, **latency_stats(reqs2)}
out = {"config": "deepseek-v2-lite TP=1 semi-PD, synthetic %d-in/%d-out, %d requests" % (a.input_len, a.output_len, a.requests),
       "dtype": "bf16", "data": "synthetic", "weights": "dummy (seeded random, 15.7 B parameters)",
       "sm_partition": {"prefill_sms": eng.p_sms, "decode_sms": eng.d_sms}, "setup_s": setup,
       "saturated": sat, "poisson": at_rate,
       "decode_graphs": eng.d_worker.model_runner.cuda_graph_runner is not None,
       "retractions": eng.decode.num_retracted}
if a.sharegpt_like:
    import numpy as np

    def lognormal(rng, mean, std, n):
        s2 = np.log(1.0 + (std / mean) ** 2)
        return np.clip(np.rint(rng.lognormal(np.log(mean) - s2 / 2, np.sqrt(s2), n)), 4, 2048).astype(int)

    rng = np.random.RandomState(7)
    n = a.sharegpt_like
    lin, lout = lognormal(rng, 220.0, 240.0, n), lognormal(rng, 190.0, 200.0, n)
    arr = np.cumsum(np.random.RandomState(1).exponential(1.0 / a.rate, size=n))
    ps = [[int((17 * i + j) % cfg.vocab_size) for j in range(int(lin[i]))] for i in range(n)]
    eng.set_device_resident(False)
    t0 = time.perf_counter()
    for i in range(n):
        wait = t0 + arr[i] - time.perf_counter()
        if wait > 0:
            time.sleep(wait)
        eng.submit(ps[i], int(lout[i]), arrival_time=t0 + arr[i])
    reqs3 = eng.wait(n)
    torch.cuda.synchronize()
    w3 = time.perf_counter() - t0
    out["sharegpt_like"] = {"requests": n, "request_rate": a.rate, "mean_in": float(lin.mean()), "mean_out": float(lout.mean()),
                            "output_tok_per_s": float(lout.sum()) / w3, **latency_stats(reqs3),
                            "note": "lengths: fixed-seed log-normal substitution for ShareGPT (SURVEY §8d config 4)"}
eng.shutdown()
print(json.dumps(out))

"""Where a request's TTFT goes at low load: arrival -> admitted by D -> fetched by P -> forward
launched -> ids read back -> delivered by D.  usage: python scripts/debug_ttft_breakdown.py [rate] [n]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from bench import make_prompts, run_window  # noqa: E402
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.semi_pd_scheduler import SemiPDEngine  # noqa: E402

rate = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 48
cfg = get_model_config("llama-3-8b")
sa = ServerArgs(model_path="llama-3-8b", dtype="bfloat16", load_format="dummy", max_total_tokens=65536,
                max_running_requests=64, context_length=1160, enable_semi_pd=True,
                cuda_graph_bs=[1, 2, 4, 8, 16, 24, 32, 48, 64], cuda_graph_max_bs=64)
eng = SemiPDEngine(sa, gpu_id=0, model_config=cfg)
eng.warmup()
prompts = make_prompts(n, 1024, cfg.vocab_size)
run_window(eng, prompts[:8], None, 16, float("inf"), False)
_, reqs, _ = run_window(eng, prompts, None, 128, rate, False)
rows = []
for r in reqs:
    st = getattr(r, "ttft_stamps", None)
    if not st or len(st) < 4 or st[0] is None:
        continue
    adm, fetched, launched, read = st
    rows.append([adm - r.arrival_time, fetched - adm, launched - fetched, read - launched, r.first_token_time - read,
                 r.first_token_time - r.arrival_time])
a = np.array(rows) * 1e3
names = ["arrival->admitted", "admitted->fetched by P", "fetched->forward launched (host)",
         "launched->ids read back (GPU)", "read back->delivered by D", "TTFT"]
print(f"rate {rate} req/s, {len(rows)} requests; ms: p50 / mean")
for i, nm in enumerate(names):
    print(f"  {nm:36s} {np.median(a[:, i]):7.2f} / {a[:, i].mean():7.2f}")
eng.shutdown()

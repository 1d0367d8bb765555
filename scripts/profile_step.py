"""ncu target: one prefill batch (8 x 1024 tokens) and a few eager decode steps (B = 64, ctx ~1088)
of Llama-3-8B through ModelRunner (no CUDA graphs so every kernel is a separate launch).
The region of interest is bracketed with cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import sys

import torch

sys.path.insert(0, ".")
from semi_pd_b200.configs import ServerArgs, get_model_config  # noqa: E402
from semi_pd_b200.forward_batch_info import ForwardBatch  # noqa: E402
from semi_pd_b200.model_runner import ModelRunner  # noqa: E402
from semi_pd_b200.schedule_batch import Req, SamplingParams, ScheduleBatch  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = get_model_config(model)
sa = ServerArgs(model_path=model, load_format="dummy", max_total_tokens=B * 1200, max_running_requests=B,
                context_length=1200, disable_cuda_graph=True, enable_semi_pd=True)
mr = ModelRunner(cfg, 0.8, 0, 0, 1, 0, sa)
mr.init_attention_backend()


def run(sb):
    fb = ForwardBatch.init_new(sb.get_model_worker_batch(), mr)
    out = mr.forward(fb)
    return mr.sample(out, fb)


reqs_all, batches = [], []
for g in range(B // 8):
    reqs = [Req(f"r{g}-{i}", [(7 * g + i + j) % cfg.vocab_size for j in range(1024)], SamplingParams()) for i in range(8)]
    for r in reqs:
        r.init_next_round_input()
    sb = ScheduleBatch.init_new(reqs, mr.req_to_token_pool, mr.token_to_kv_pool_allocator, mr.device)
    sb.prepare_for_extend()
    if g == B // 8 - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ids = run(sb)
    sb.output_ids = ids.long()
    batches.append(sb)
torch.cuda.synchronize()
run_b = batches[0]
for sb in batches[1:]:
    run_b.merge_batch(sb)
for step in range(3):
    run_b.prepare_for_decode()
    run_b.output_ids = run(run_b).long()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

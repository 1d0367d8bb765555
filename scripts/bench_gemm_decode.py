"""decode-regime GEMM sweep (graph-timed).  usage: python scripts/bench_gemm_decode.py [sm_budget]"""
import sys
import torch
sys.path.insert(0, ".")
from semi_pd_b200 import ops

dev = torch.device("cuda:0")
budget = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ops.set_sm_budget(budget)


def graph_time(fns, reps=5):
    torch.cuda.synchronize()  # inputs were produced on the default stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fns:
                f()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / len(fns)


for M in (16, 64, 128, 160):
    row = []
    for (N, K, name) in ((6144, 4096, "qkv"), (4096, 4096, "o"), (28672, 4096, "gate_up"), (4096, 14336, "down")):
        ws = [torch.randn(N, K, device=dev).bfloat16() for _ in range(8)]
        x = torch.randn(M, K, device=dev).bfloat16()
        outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(8)]
        ms = graph_time([lambda w=w, o=o: ops.linear(x, w, out=o) for w, o in zip(ws, outs)])
        row.append(f"{name} {ms*1e3:6.1f}us")
        del ws, outs
    print(f"M={M:4d}: " + " | ".join(row), flush=True)

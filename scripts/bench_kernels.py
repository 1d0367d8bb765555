"""Kernel micro-benchmarks at Llama-3-8B decode/prefill shapes, timed as CUDA-graph replays of N
back-to-back launches (no Python launch overhead), L2 flushed by the working set itself
(distinct weight buffers per launch).  usage: python scripts/bench_kernels.py [sm_budget] [sections=gd,ad,ap,gp]
(gd = decode-regime GEMMs, ad = decode attention, ap = prefill attention, gp = prefill-regime GEMMs)"""
import sys

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
budget = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ops.set_sm_budget(budget)
sections = (sys.argv[2] if len(sys.argv) > 2 else "gd,ad,ap,gp").split(",")
torch.manual_seed(0)


def graph_time(fns, reps=5):
    """fns: list of callables launched back to back inside one graph"""
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


print(f"== decode-regime GEMMs (sm_budget={budget or 'all'}) ==")
import os  # noqa: E402
MS = tuple(int(x) for x in os.environ.get("SPD_BENCH_MS", "1,16,64,128,160,192").split(","))
for M in (MS if "gd" in sections else ()):
    for (N, K, name) in ((6144, 4096, "qkv"), (4096, 4096, "o"), (28672, 4096, "gate_up"), (4096, 14336, "down"),
                         (128256, 4096, "lm_head")):
        nbuf = 3 if N > 100000 else 8
        ws = [torch.randn(N, K, device=dev).bfloat16() for _ in range(nbuf)]
        x = torch.randn(M, K, device=dev).bfloat16()
        outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        ms = graph_time([lambda w=w, o=o: ops.linear(x, w, out=o) for w, o in zip(ws, outs)])
        ms_cublas = graph_time([lambda w=w, o=o: torch.matmul(x, w.t(), out=o) for w, o in zip(ws, outs)])
        by = N * K * 2 + M * K * 2 + M * N * 2
        print(f"  M={M:4d} {name:8s} N={N:6d} K={K:5d}: spd {ms*1e3:7.1f} us {by/ms/1e6:6.0f} GB/s | cublas {ms_cublas*1e3:7.1f} us"
              f" {by/ms_cublas/1e6:6.0f} GB/s", flush=True)
        del ws, outs

print("== decode attention ==")
Hq, Hk, D = 32, 8, 128
for (B, S) in (((1, 1088), (8, 1088), (32, 1088), (64, 1088), (128, 1088), (160, 1088), (16, 8192), (64, 300))
               if "ad" in sections else ()):
    slots = B * S + 8
    nb = 4
    kps = [torch.randn(slots, Hk, D, device=dev).bfloat16() for _ in range(nb)]
    vps = [torch.randn(slots, Hk, D, device=dev).bfloat16() for _ in range(nb)]
    rtt = (torch.randperm(B * S, device=dev).int() + 1).view(B, S).contiguous()
    rpi = torch.arange(B, device=dev)
    sl = torch.full((B,), S, dtype=torch.int64, device=dev)
    q = torch.randn(B, Hq * D, device=dev).bfloat16()
    o = torch.empty_like(q)
    ws = torch.zeros(ops.lib.spd_attn_decode_workspace_bytes(B, Hq, D), dtype=torch.uint8, device=dev)
    ms = graph_time([lambda kp=kp, vp=vp: ops.attn_decode(q, kp, vp, rtt, rpi, sl, Hq, Hk, D, D ** -0.5, out=o, workspace=ws)
                     for kp, vp in zip(kps, vps)])
    by = B * S * Hk * D * 2 * 2 + 2 * B * Hq * D * 2 + 4 * B * S
    print(f"  B={B:4d} S={S:5d}: {ms*1e3:7.1f} us {by/ms/1e6:6.0f} GB/s", flush=True)
    del kps, vps

print("== prefill attention ==")
for (nreq, E, P) in (((1, 1024, 0), (8, 1024, 0), (2, 4096, 0), (4, 1024, 3072)) if "ap" in sections else ()):
    L = E + P
    slots = nreq * L + 8
    kp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    vp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    rtt = (torch.arange(nreq * L, device=dev, dtype=torch.int32) + 1).view(nreq, L).contiguous()
    rpi = torch.arange(nreq, device=dev)
    pre = torch.full((nreq,), P, dtype=torch.int32, device=dev)
    ext = torch.full((nreq,), E, dtype=torch.int32, device=dev)
    _, st = ops.compute_positions(pre, ext, nreq * E)
    q = torch.randn(nreq * E, Hq * D, device=dev).bfloat16()
    o = torch.empty_like(q)
    ms = graph_time([lambda: ops.attn_prefill(q, kp, vp, rtt, rpi, pre, ext, st, E, Hq, Hk, D, D ** -0.5, out=o)] * 4)
    fl = 4.0 * Hq * D * nreq * E * (P + (E + 1) / 2)
    print(f"  {nreq} x {E} (+{P} prefix): {ms*1e3:7.1f} us {fl/ms/1e9:6.1f} TF/s", flush=True)

print("== prefill-regime GEMMs ==")
for M in ((1024, 8192) if "gp" in sections else ()):
    for (N, K, name, epi) in ((6144, 4096, "qkv", 0), (4096, 4096, "o", 0), (14336, 4096, "gate_up+silu", 1),
                              (4096, 14336, "down", 0)):
        w = torch.randn(N * (2 if epi else 1), K, device=dev).bfloat16()
        x = torch.randn(M, K, device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = graph_time([lambda: ops.linear(x, w, out=o, epilogue=epi)] * 4)
        fl = 2.0 * M * N * (2 if epi else 1) * K
        wf = w if not epi else w
        ms_c = graph_time([lambda: torch.matmul(x, wf.t())] * 4)
        print(f"  M={M:5d} {name:13s}: spd {ms*1e3:7.1f} us {fl/ms/1e9:6.0f} TF/s | cublas {ms_c*1e3:7.1f} us {fl/ms_c/1e9:6.0f} TF/s",
              flush=True)

"""Why does the fused W_gate_up + SiLU*mul GEMM lose ~26 % in back-to-back launches (sustained
clocks) while the other prefill GEMMs hold cuBLAS's rate?  Variants timed as graphs of N launches.
usage: python scripts/debug_gateup_sustained.py [sm_budget]"""
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402

budget = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ops.set_sm_budget(budget)
dev = torch.device("cuda:0")
M, K, I = 8192, 4096, 14336
torch.manual_seed(0)
x = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(2 * I, K, device=dev).bfloat16()
o_f = [torch.empty(M, I, device=dev, dtype=torch.bfloat16) for _ in range(2)]
o_u = [torch.empty(M, 2 * I, device=dev, dtype=torch.bfloat16) for _ in range(2)]
w_d = torch.randn(K, I, device=dev).bfloat16()
o_d = torch.empty(M, K, device=dev, dtype=torch.bfloat16)


def clocks(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip().split(",")
        try:
            out.append((float(r[0]), float(r[1])))
        except Exception:  # noqa: BLE001
            pass
        time.sleep(0.05)


def run(name, fns, flops, reps=int(__import__('os').environ.get('REPS', 400))):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fns:
                f()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        stop, samples = threading.Event(), []
        th = threading.Thread(target=clocks, args=(stop, samples))
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
    ms = e0.elapsed_time(e1) / reps / len(fns)
    mhz = sorted(c for c, _ in samples)[len(samples) // 2] if samples else 0
    watt = sorted(p for _, p in samples)[len(samples) // 2] if samples else 0
    print(f"{name:44s} {ms*1e3:8.1f} us {flops/ms/1e9:7.0f} TF/s   sm {mhz:.0f} MHz {watt:.0f} W", flush=True)


fl = 2.0 * M * 2 * I * K
run("fused gate_up+silu x4 (same out)", [lambda: ops.linear(x, w, out=o_f[0], epilogue=ops.SPD_EPI_SILU_MUL)] * 4, fl)
run("fused gate_up+silu x4 (alternating out)", [lambda i=i: ops.linear(x, w, out=o_f[i & 1], epilogue=ops.SPD_EPI_SILU_MUL) for i in range(4)], fl)
run("plain gate_up (N = 28672) x4", [lambda: ops.linear(x, w, out=o_u[0])] * 4, fl)
run("cuBLAS gate_up x4", [lambda: torch.matmul(x, w.t(), out=o_u[1])] * 4, fl)
run("plain down x4", [lambda: ops.linear(o_f[0], w_d, out=o_d)] * 4, 2.0 * M * K * I)
run("fused gate_up+silu, down alternating x2", [lambda: ops.linear(x, w, out=o_f[0], epilogue=ops.SPD_EPI_SILU_MUL),
                                               lambda: ops.linear(o_f[0], w_d, out=o_d)] * 2, (fl + 2.0 * M * K * I) / 2)

"""Prefill-regime GEMM experiment: fused gate_up (SiLU*mul epilogue) vs the same GEMM without the
epilogue vs cuBLAS, with SM clock / power sampled during each timed loop (NVML) to separate
"kernel is slower" from "kernel draws more power and the clock drops".
usage: python scripts/bench_gemm_prefill.py [M] [sm_budget]"""
import sys
import threading
import time

import pynvml
import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ops.set_sm_budget(budget)
torch.manual_seed(0)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler:
    def __enter__(self):
        self.stop = False
        self.clk, self.pw = [], []
        self.t = threading.Thread(target=self.run)
        self.t.start()
        return self

    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.005)

    def __exit__(self, *a):
        self.stop = True
        self.t.join()


def timed(fn, n_graph=8, seconds=1.0):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n_graph):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # size the loop to ~`seconds` so the power governor reaches steady state
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        reps = max(2, int(seconds * 1e3 / e0.elapsed_time(e1)))
        with Sampler() as sm:
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps / n_graph
    clk = sorted(sm.clk)[len(sm.clk) // 2] if sm.clk else 0
    pw = sorted(sm.pw)[len(sm.pw) // 2] if sm.pw else 0
    return ms, clk, pw


K, I = 4096, 14336
x = torch.randn(M, K, device=dev).bfloat16()
w_gu = (torch.randn(2 * I, K, device=dev) * 0.02).bfloat16()
w_dn = (torch.randn(K, I, device=dev) * 0.02).bfloat16()
w_qkv = (torch.randn(6144, K, device=dev) * 0.02).bfloat16()
h_act = torch.randn(M, I, device=dev).bfloat16()
o_f = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
o_p = torch.empty(M, 2 * I, device=dev, dtype=torch.bfloat16)
o_h = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
o_d = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
o_q = torch.empty(M, 6144, device=dev, dtype=torch.bfloat16)

cases = [
    ("gate_up fused silu  (spd)", lambda: ops.linear(x, w_gu, out=o_f, epilogue=1), 2.0 * M * 2 * I * K),
    ("gate_up plain N=28672 (spd)", lambda: ops.linear(x, w_gu, out=o_p), 2.0 * M * 2 * I * K),
    ("gate_up plain N=28672 (cublas)", lambda: torch.matmul(x, w_gu.t(), out=o_p), 2.0 * M * 2 * I * K),
    ("half    plain N=14336 (spd)", lambda: ops.linear(x, w_gu[:I], out=o_h), 2.0 * M * I * K),
    ("qkv     plain N=6144  (spd)", lambda: ops.linear(x, w_qkv, out=o_q), 2.0 * M * 6144 * K),
    ("qkv     plain N=6144  (cublas)", lambda: torch.matmul(x, w_qkv.t(), out=o_q), 2.0 * M * 6144 * K),
    ("down    K=14336       (spd)", lambda: ops.linear(h_act, w_dn, out=o_d), 2.0 * M * I * K),
    ("down    K=14336       (cublas)", lambda: torch.matmul(h_act, w_dn.t(), out=o_d), 2.0 * M * I * K),
]
print(f"M={M} sm_budget={budget or 'all'}")
for name, fn, fl in cases:
    ms, clk, pw = timed(fn)
    print(f"  {name:32s}: {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TF/s   sm_clk {clk:5d} MHz  power {pw:6.0f} W", flush=True)

"""In-kernel timeline of the prefill attention (CTA 0): when each softmax warpgroup waits for S /
publishes P and when the MMA issuer waits for P / issues PV + next S.
usage: python scripts/debug_attn_prefill_timeline.py [sm_budget] [nreq] [ext_len]   (SPD_ATTN_PF_UNIT=64|128)"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from semi_pd_b200 import _lib, ops  # noqa: E402

budget = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nreq = int(sys.argv[2]) if len(sys.argv) > 2 else 2
E = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda:0")
ops.set_sm_budget(budget)
Hq, Hk, D, P = 32, 8, 128, 0
torch.manual_seed(0)
L = E + P
kp = torch.randn(nreq * L + 8, Hk, D, device=dev).bfloat16()
vp = torch.randn(nreq * L + 8, Hk, D, device=dev).bfloat16()
rtt = (torch.arange(nreq * L, device=dev, dtype=torch.int32) + 1).view(nreq, L).contiguous()
rpi = torch.arange(nreq, device=dev)
pre = torch.full((nreq,), P, dtype=torch.int32, device=dev)
ext = torch.full((nreq,), E, dtype=torch.int32, device=dev)
_, st = ops.compute_positions(pre, ext, nreq * E)
q = torch.randn(nreq * E, Hq * D, device=dev).bfloat16()
o = torch.empty_like(q)
run = lambda: ops.attn_prefill(q, kp, vp, rtt, rpi, pre, ext, st, E, Hq, Hk, D, D ** -0.5, out=o)  # noqa: E731
run()
torch.cuda.synchronize()
cap = 8192
buf = torch.zeros(3 * cap * 2, dtype=torch.int64, device=dev)
fn = _lib.lib.spd_debug_attn_prefill_trace
fn.argtypes = [C.c_void_p, C.c_int64]
fn(buf.data_ptr(), cap)
run()
torch.cuda.synchronize()
fn(None, 0)
ev = buf.cpu().numpy().reshape(3, cap, 2)
t0 = min(int(ev[r, 0, 1]) for r in range(3) if ev[r, 0, 0])
names = {0: "softmax wg0", 1: "softmax wg1", 2: "mma issuer"}
for r in range(3):
    e = ev[r][ev[r, :, 0] != 0]
    tags, clk = e[:, 0] >> 8, e[:, 1] - t0
    print(f"-- {names[r]}: {len(e)} events, span {clk[-1] - clk[0]} clk")
    if r < 2:
        w = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 1 and tags[i + 1] == 2]
        c = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 2 and tags[i + 1] == 3]
        g = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 3 and tags[i + 1] == 1]
        ep = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 4]
        st_ = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 5 and tags[i + 1] == 6]
        nx = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 6]
        first = [clk[i + 2] - clk[i + 1] for i in range(len(e) - 2) if tags[i] == 6 and tags[i + 1] == 1]
        print(f"   item boundary: drain (last PV) {np.mean(ep) if ep else 0:.0f}, O -> global {np.mean(st_) if st_ else 0:.0f}, "
              f"next item setup {np.mean(nx) if nx else 0:.0f}, first wait-for-S {np.mean(first) if first else 0:.0f}")
        print(f"   units {len(c)}: wait-for-S mean {np.mean(w):.0f} (p50 {np.median(w):.0f}, p90 {np.percentile(w, 90):.0f}), "
              f"pass mean {np.mean(c):.0f} (p50 {np.median(c):.0f}, p90 {np.percentile(c, 90):.0f}), loop gap {np.mean(g):.0f}; "
              f"epilogue+next-item {np.mean(ep) if ep else 0:.0f} x {len(ep)}")
    else:
        w = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 1 and tags[i + 1] == 2]
        c = [clk[i + 1] - clk[i] for i in range(len(e) - 1) if tags[i] == 2 and tags[i + 1] == 3]
        print(f"   units {len(c)}: wait-for-P mean {np.mean(w):.0f} (p50 {np.median(w):.0f}, p90 {np.percentile(w, 90):.0f}), "
              f"V-wait+issue mean {np.mean(c):.0f} (p50 {np.median(c):.0f}, p90 {np.percentile(c, 90):.0f})")
# merged excerpt from the middle of the run
allev = sorted((int(ev[r, i, 1]) - t0, r, int(ev[r, i, 0])) for r in range(3) for i in range(cap) if ev[r, i, 0])
mid = len(allev) // 2
lab = {1: "wait", 2: "ready", 3: "done", 4: "epilogue", 5: "drained", 6: "stored"}
print("-- excerpt (clk, role, event, unit)")
for t, r, tag in allev[mid:mid + 72]:
    print(f"{t:9d}  {names[r]:12s} {lab[tag >> 8]:8s} {tag & 0xff}")
np.save("gpurun_out/attn_prefill_timeline.npy", ev)

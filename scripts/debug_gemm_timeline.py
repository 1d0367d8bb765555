"""Per-CTA timeline of the pair decode GEMM (globaltimer stamps).  usage: python scripts/debug_gemm_timeline.py M N K"""
import os, sys
import torch
dev = torch.device("cuda:0")
M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 6144, 4096)
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
os.environ["SPD_GEMM_DBG_PTR"] = hex(dbg.data_ptr())
sys.path.insert(0, ".")
from semi_pd_b200 import ops
x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(6)]
for w in ws[:-1]:
    ops.linear(x, w)
torch.cuda.synchronize()
dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.linear(x, ws[-1]); e1.record(); torch.cuda.synchronize()
print("event time us", e0.elapsed_time(e1) * 1e3)
d = dbg.view(148, 8).cpu().double()
t0 = d[:, 0][d[:, 0] > 0].min()
names = ["roles start", "last segment MMAs done", "partial published", "owner: loop done", "owner: parked acc complete", "owner: flags seen", "epilogue done"]
for i, nm in enumerate(names):
    col = d[:, i]
    ok = col > 0
    if ok.any():
        v = (col[ok] - t0) / 1e3
        print(f"{nm:28s}: n={int(ok.sum()):3d}  min {v.min():6.2f}  mean {v.mean():6.2f}  max {v.max():6.2f} us")

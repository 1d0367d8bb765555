"""ncu target: a few launches of the decode / prefill attention kernels at Llama-3-8B shapes.
usage: python scripts/prof_attn.py decode|prefill [B] [S]"""
import sys

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1]
Hq, Hk, D = 32, 8, 128
torch.manual_seed(0)
if what == "decode":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 1088
    slots = B * S + 8
    kp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    vp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    rtt = (torch.randperm(B * S, device=dev).int() + 1).view(B, S).contiguous()
    rpi = torch.arange(B, device=dev)
    sl = torch.full((B,), S, dtype=torch.int64, device=dev)
    q = torch.randn(B, Hq * D, device=dev).bfloat16()
    for _ in range(5):
        ops.attn_decode(q, kp, vp, rtt, rpi, sl, Hq, Hk, D, D ** -0.5)
else:
    nreq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    E = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    slots = nreq * E + 8
    kp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    vp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    rtt = (torch.arange(nreq * E, device=dev, dtype=torch.int32) + 1).view(nreq, E).contiguous()
    rpi = torch.arange(nreq, device=dev)
    pre = torch.zeros(nreq, dtype=torch.int32, device=dev)
    ext = torch.full((nreq,), E, dtype=torch.int32, device=dev)
    _, st = ops.compute_positions(pre, ext, nreq * E)
    q = torch.randn(nreq * E, Hq * D, device=dev).bfloat16()
    for _ in range(5):
        ops.attn_prefill(q, kp, vp, rtt, rpi, pre, ext, st, E, Hq, Hk, D, D ** -0.5)
torch.cuda.synchronize()

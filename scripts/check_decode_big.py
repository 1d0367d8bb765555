import sys
import torch
sys.path.insert(0, ".")
from semi_pd_b200 import ops
dev = torch.device("cuda:0")
Hq, Hk, D = 32, 8, 128
for (B, S) in ((64, 1088), (128, 1088), (160, 1088), (128, 1088), (256, 600), (16, 8192)):
    slots = B * S + 8
    kp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    vp = torch.randn(slots, Hk, D, device=dev).bfloat16()
    rtt = (torch.randperm(B * S, device=dev).int() + 1).view(B, S).contiguous()
    rpi = torch.arange(B, device=dev)
    sl = torch.full((B,), S, dtype=torch.int64, device=dev)
    q = torch.randn(B, Hq * D, device=dev).bfloat16()
    for rep in range(3):
        a = ops.attn_decode(q, kp, vp, rtt, rpi, sl, Hq, Hk, D, D ** -0.5)
        torch.cuda.synchronize()
    b = ops.attn_decode(q, kp, vp, rtt, rpi, sl, Hq, Hk, D, D ** -0.5, impl="simt")
    torch.cuda.synchronize()
    print(B, S, "max diff tc vs simt", (a.float() - b.float()).abs().max().item(), flush=True)

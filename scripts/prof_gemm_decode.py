"""ncu target: decode-regime GEMM (swap-AB) at one Llama-3-8B shape.
usage: python scripts/prof_gemm_decode.py [M=128] [N=6144] [K=4096]"""
import sys
import torch
sys.path.insert(0, ".")
from semi_pd_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6144
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(6)]
for w in ws:
    o = ops.linear(x, w)
torch.cuda.synchronize()

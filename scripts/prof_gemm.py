"""ncu target: the dominant kernel of the window — W_gate_up GEMM with the fused SiLU*mul epilogue
at a prefill batch of 8192 tokens (Llama-3-8B: N = 2 x 14336, K = 4096)."""
import sys
import torch
sys.path.insert(0, ".")
from semi_pd_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = (torch.randn(M, 4096, device=dev) * 0.5).bfloat16()
w = (torch.randn(2 * 14336, 4096, device=dev) / 64).bfloat16()
for _ in range(5):
    o = ops.linear(x, w, epilogue=ops.SPD_EPI_SILU_MUL)
torch.cuda.synchronize()

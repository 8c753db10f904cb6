"""Developer probe: one forward-shaped and one weight-gradient-shaped GEMM for an ncu capture."""
import sys
sys.path.insert(0, ".")
import torch
from speech_b200 import ops

dev = "cuda"
torch.manual_seed(0)
A1 = torch.randn(16000, 2048, device=dev).bfloat16()
B1 = torch.randn(6144, 2048, device=dev).bfloat16()
A2 = torch.randn(6144, 16000, device=dev).bfloat16()
B2 = torch.randn(2048, 16000, device=dev).bfloat16()
o = torch.zeros(6144, 2048, device=dev)
for _ in range(2):
    ops.gemm_bf16_tn(A1, B1)
    ops.gemm_bf16_tn(A2, B2, out=o, accumulate=True, split_k=2)
torch.cuda.synchronize()

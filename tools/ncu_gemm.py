"""One GEMM shape, a handful of launches -- the target for `ncu --set full` captures."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.gemm import gemm

layout = sys.argv[1] if len(sys.argv) > 1 else "TN"
M, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (8192, 8192, 8192)
if layout == "TN":
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
elif layout == "NN":
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16()
else:
    a = torch.randn(K, M, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16()
for _ in range(6):
    out = gemm(a, b, layout)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))

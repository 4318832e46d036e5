"""One attention fwd+bwd (Llama-3-8B shape by default), a few launches -- target for `ncu --set full` captures:
  ncu --set full --clock-control none --import-source on -k regex:fmha -c 4 -o gpurun_out/prof_fmha python tools/ncu_fmha.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.attention import fmha_bwd, fmha_fwd

B, S, H, Hk = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (1, 4096, 32, 8)
D = 128
q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16()
v = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); d_o = torch.randn(B, S, H, D, device="cuda").bfloat16()
for _ in range(2):
    o, lse = fmha_fwd(q, k, v, causal=True)
    dq, dk, dv = fmha_bwd(d_o, q, k, v, o, lse, causal=True)
torch.cuda.synchronize()
print("ok", float(dq.float().abs().mean()))

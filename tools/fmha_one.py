import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.attention import fmha_fwd
B, S, H, Hk, D, Dv, causal = (int(x) for x in sys.argv[1:8])
q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, Dv, device="cuda").bfloat16()
o, lse = fmha_fwd(q, k, v, causal=bool(causal))
torch.cuda.synchronize()
print("done", float(o.float().abs().mean()))

"""In-kernel timeline of the FA backward (CTA 0): per-iteration clock64 deltas for the MMA, softmax and dQ-drain roles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.attention import fmha_bwd, fmha_fwd

B, S, H, Hk, D = 1, 4096, 32, 8, 128
q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16()
v = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); d_o = torch.randn(B, S, H, D, device="cuda").bfloat16()
o, lse = fmha_fwd(q, k, v, causal=True)
fmha_bwd(d_o, q, k, v, o, lse, causal=True)
tr = torch.zeros(3, 32, 16, dtype=torch.int64, device="cuda")
fmha_bwd(d_o, q, k, v, o, lse, causal=True, trace=tr)
torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
names = {0: ["top", "S(g+1) issued", "p_full ok", "dV issued", "dP(g+1) issued", "ds_full ok", "dq_empty ok", "dK,dQ issued"],
         1: ["top", "stats ok", "s_full ok", "S loaded", "exp done", "pv_done ok", "P stored", "dp_full ok", "dP loaded", "dS math", "dkq_done ok", "dS stored"],
         2: ["top", "dq_full ok", "loaded", "smem free", "reduce issued"]}
for role, label in ((0, "MMA"), (1, "softmax w2"), (2, "dQ drain")):
    print(f"== {label}: events {names[role]}")
    for g in range(4, 14):
        row = [int(x) - t0 for x in t[role, g, :len(names[role])]]
        deltas = [row[0]] + [row[i] - row[i - 1] for i in range(1, len(row))]
        print(f"  g={g:2d} start {row[0]:7d}  deltas {deltas[1:]}  total {row[-1] - row[0]}")
    per_iter = (int(t[role, 20, 0]) - int(t[role, 4, 0])) / 16
    print(f"  mean clocks per iteration (g=4..20): {per_iter:.0f}")

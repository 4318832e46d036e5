"""In-kernel timeline of the second-generation attention forward (CTA (0,0,0) = the heaviest causal query tile): SM-clock stamps of the
MMA issuer and of one softmax warp per query tile, per key block.  Prints the steady-state averages (cycles)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from simumax_b200.ops.attention import fmha_fwd

B, S, H, Hk = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (1, 4096, 32, 8)
q = torch.randn(B, S, H, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
for _ in range(3):
    fmha_fwd(q, k, v, causal=True)
tr = torch.zeros(3 * 64 * 16, device="cuda", dtype=torch.int32)
fmha_fwd(q, k, v, causal=True, trace=tr)
torch.cuda.synchronize()
t = tr.cpu().numpy().astype(np.int64).reshape(3, 64, 16) & 0xFFFFFFFF
n = min(64, S // 128) - 1
lo, hi = 4, n - 2
def d(a, b):            # mean of (a - b) over steady-state blocks, modulo 2^32
    x = (a - b) & 0xFFFFFFFF
    return float(np.mean(x[lo:hi]))
mma = t[0]
print(f"blocks {n}, steady window [{lo}, {hi})")
print("MMA issuer per block: period %.0f" % d(mma[1:, 0], mma[:-1, 0]))
for tt in (0, 1):
    o = tt * 4
    print(f"  tile {tt}: p_full wait done -> v ready {d(mma[:, 1 + o], mma[:, 0 + o]):.0f} -> PV issued {d(mma[:, 2 + o], mma[:, 1 + o]):.0f} -> (k wait +) S issued {d(mma[:, 3 + o], mma[:, 2 + o]):.0f}")
print("  tile0 done -> tile1 p_full done %.0f ; tile1 done -> next tile0 p_full done %.0f" % (d(mma[:, 4], mma[:, 3]), d(mma[1:, 0], mma[:-1, 7])))
for tt in (0, 1):
    s = t[1 + tt]
    print(f"softmax tile {tt}: period {d(s[1:, 0], s[:-1, 0]):.0f} = S wait {d(s[1:, 0], s[:-1, 4]):.0f} + tmem ld {d(s[:, 1], s[:, 0]):.0f} + mask/max {d(s[:, 2], s[:, 1]):.0f} + exp/pack {d(s[:, 3], s[:, 2]):.0f} + st/arrive {d(s[:, 4], s[:, 3]):.0f}")
    print(f"   p_full arrive -> MMA saw it {d(mma[1:, 0 + tt * 4], s[1:, 4]):.0f} ; S issued -> softmax woke {d(s[1:, 0], mma[:-1, 3 + tt * 4]):.0f}")

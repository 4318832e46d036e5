import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import math, torch
    from simumax_b200.ops.attention import fmha_fwd
    B, S, H, Hk, D, Dv, causal = (int(x) for x in sys.argv[1:8])
    torch.manual_seed(0)
    q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, Dv, device="cuda").bfloat16()
    o, lse = fmha_fwd(q, k, v, causal=bool(causal))
    torch.cuda.synchronize()
    qf = q.float().permute(0, 2, 1, 3); kf = k.float().permute(0, 2, 1, 3).repeat_interleave(H // Hk, 1); vf = v.float().permute(0, 2, 1, 3).repeat_interleave(H // Hk, 1)
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    if causal:
        s = s.masked_fill(~torch.ones(S, S, device="cuda", dtype=torch.bool).tril(), float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3)
    print("maxerr", float((o.float() - ref).abs().max()))
    sys.exit(0)
cases = [(1, 200, 2, 2, 128, 128, 1), (1, 200, 2, 2, 128, 128, 0), (1, 1024, 4, 4, 192, 128, 0), (1, 1024, 4, 4, 192, 128, 1),
         (1, 192, 2, 2, 128, 128, 1), (1, 320, 2, 2, 128, 128, 1), (1, 130, 1, 1, 128, 128, 0), (1, 64, 1, 1, 128, 128, 0)]
for c in cases:
    try:
        r = subprocess.run([sys.executable, __file__] + [str(x) for x in c], capture_output=True, text=True, timeout=40)
        print(c, "rc", r.returncode, r.stdout.strip()[-80:], r.stderr.strip()[-200:].replace("\n", " | "), flush=True)
    except subprocess.TimeoutExpired:
        print(c, "TIMEOUT", flush=True)

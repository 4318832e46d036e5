"""FA forward timing vs torch SDPA on the current GPU (CUDA events, median)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.attention import fmha_fwd, attention_flops


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); evs.append((s, e))
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]


rows = []
for (B, S, H, Hk, D, Dv) in [(1, 4096, 32, 8, 128, 128), (1, 4096, 32, 4, 128, 128), (2, 8192, 16, 16, 128, 128), (1, 4096, 128, 128, 192, 128), (1, 4096, 4, 1, 128, 128)]:
    q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, Dv, device="cuda").bfloat16()
    t = timeit(lambda: fmha_fwd(q, k, v, causal=True))
    fl = attention_flops(B, S, S, H, D, Dv, True)
    res = {"B": B, "S": S, "H": H, "Hk": Hk, "D": D, "Dv": Dv, "smx_ms": t, "smx_tflops_causal": fl / t / 1e9}
    try:
        qt, kt, vt = (x.permute(0, 2, 1, 3) for x in (q, k, v))
        if D == Dv:
            f = lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=True, enable_gqa=(H != Hk))
            t2 = timeit(f)
            res["sdpa_ms"] = t2; res["sdpa_tflops_causal"] = fl / t2 / 1e9
    except Exception as ex:
        res["sdpa_err"] = str(ex)[:100]
    print(json.dumps(res), flush=True)
    rows.append(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/fmha_bench.json", "w"), indent=1)

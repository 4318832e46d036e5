"""FA backward throughput: this repo's tcgen05 kernel vs flash_attn (FA2) and SDPA backward."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.attention import attention_flops, fmha_bwd, fmha_fwd


def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = []
for (B, S, H, Hk) in [(1, 4096, 32, 8), (1, 8192, 32, 8), (4, 2048, 32, 8), (1, 4096, 4, 1), (1, 16384, 16, 16)]:
    D = 128
    q = torch.randn(B, S, H, D, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, D, device="cuda").bfloat16()
    v = torch.randn(B, S, Hk, D, device="cuda").bfloat16(); d_o = torch.randn(B, S, H, D, device="cuda").bfloat16()
    o, lse = fmha_fwd(q, k, v, causal=True)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    fl = attention_flops(B, S, S, H, D, D, True, backward=True)
    row = {"B": B, "S": S, "H": H, "Hkv": Hk, "flops": fl}
    ms = timeit(lambda: fmha_bwd(d_o, q, k, v, o, lse, dq, dk, dv, causal=True)); row["smx_ms"] = ms; row["smx_tflops"] = fl / ms / 1e9
    try:
        from flash_attn.flash_attn_interface import _flash_attn_backward
        sc = 1 / math.sqrt(D)
        ms = timeit(lambda: _flash_attn_backward(d_o, q, k, v, o, lse, dq, dk, dv, 0.0, sc, True, -1, -1, 0.0, None, False))
        row["fa2_ms"] = ms; row["fa2_tflops"] = fl / ms / 1e9
    except Exception as e:
        row["fa2_err"] = str(e)[:100]
    try:
        qa, ka, va = (t.permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v))
        oo = torch.nn.functional.scaled_dot_product_attention(qa, ka, va, is_causal=True, enable_gqa=True)
        g = d_o.permute(0, 2, 1, 3)
        ms = timeit(lambda: torch.autograd.grad(oo, (qa, ka, va), g, retain_graph=True))
        row["sdpa_ms"] = ms; row["sdpa_tflops"] = fl / ms / 1e9
    except Exception as e:
        row["sdpa_err"] = str(e)[:100]
    print(row, flush=True); out.append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fmha_bwd_bench.json", "w"), indent=1)

"""fp32-accumulating wgrad GEMMs (NT, out += dy^T x): read-modify-write epilogue vs the red.global.add.v4.f32 epilogue (SMX_GEMM_RED=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.gemm import gemm

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (G, M, N, K) in [(1, 28672, 4096, 4096), (1, 4096, 14336, 4096), (1, 6144, 4096, 4096), (1, 128384, 4096, 4096), (20, 3072, 5120, 1232), (1, 3584, 4096, 4096)]:
    shape = (lambda *s: s) if G == 1 else (lambda *s: (G, *s))
    a = torch.randn(*shape(K, M), device="cuda").bfloat16(); b = torch.randn(*shape(K, N), device="cuda").bfloat16()
    outs = [torch.zeros(*shape(M, N), device="cuda") for _ in range(3)]
    i = [0]
    def f():
        i[0] += 1
        gemm(a, b, "NT", out=outs[i[0] % 3], accumulate=True)
    ms = timeit(f)
    ref = (a[0] if G > 1 else a).float().t() @ (b[0] if G > 1 else b).float()
    chk = torch.zeros_like(outs[0]); gemm(a, b, "NT", out=chk, accumulate=True); gemm(a, b, "NT", out=chk, accumulate=True)
    got = chk[0] if G > 1 else chk
    err = float((got - 2 * ref).abs().max() / ref.abs().max())
    print(f"G{G} M{M} N{N} K{K}: {ms:.4f} ms  {2.0 * G * M * N * K / ms / 1e9:7.1f} TF  rel err {err:.2e}", flush=True)

"""MXFP8 block-scaled GEMM: correctness vs dequantised fp32 matmul, and speed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.gemm import gemm, gemm_mxfp8, mxfp8_dequantize, mxfp8_quantize

torch.manual_seed(0)
ok = True
for (M, N, K) in [(128, 256, 128), (256, 512, 256), (200, 300, 384), (1024, 1024, 1024), (4096, 4096, 4096)]:
    # rows with very different magnitudes per 32-block so that wrong scale placement cannot hide
    x = torch.randn(M, K, device="cuda") * torch.exp2(torch.randint(-6, 6, (M, K // 32), device="cuda").float()).repeat_interleave(32, 1)
    w = torch.randn(N, K, device="cuda") * torch.exp2(torch.randint(-6, 6, (N, K // 32), device="cuda").float()).repeat_interleave(32, 1)
    qa, sa = mxfp8_quantize(x); qb, sb = mxfp8_quantize(w)
    ref = mxfp8_dequantize(qa, sa) @ mxfp8_dequantize(qb, sb).t()
    out = gemm_mxfp8(qa, sa, qb, sb, out_dtype=torch.float32)
    torch.cuda.synchronize()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    good = err < 2e-3
    ok &= good
    print(f"M{M} N{N} K{K}: rel err {err:.3e} {'OK' if good else 'FAIL'}", flush=True)
print("MXFP8_CORRECT" if ok else "MXFP8_MISMATCH", flush=True)
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (4096, 28672, 4096)]:
    qa, sa = mxfp8_quantize(torch.randn(M, K, device="cuda")); qb, sb = mxfp8_quantize(torch.randn(N, K, device="cuda"))
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): gemm_mxfp8(qa, sa, qb, sb, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): gemm_mxfp8(qa, sa, qb, sb, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"M{M} N{N} K{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TF/s (fp8 dense nominal 4500)", flush=True)

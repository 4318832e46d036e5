"""CTA-pair GEMM: correctness vs fp32 torch and speed vs the 1-CTA kernel / cuBLAS, all layouts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.gemm import gemm


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def mk(lay, M, N, K):
    a = torch.randn((K, M) if lay in ("NT", "TT") else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if lay in ("NN", "NT") else (N, K), device="cuda").bfloat16()
    af = a.float().t() if lay in ("NT", "TT") else a.float()
    bf = b.float() if lay in ("NN", "NT") else b.float().t()
    return a, b, af, bf


torch.manual_seed(0)
ok = True
for lay in ("TN", "NN", "NT", "TT"):
    for (M, N, K) in [(256, 256, 64), (512, 768, 320), (1000, 520, 200), (4096, 4096, 1024)]:
        a, b, af, bf = mk(lay, M, N, K)
        ref = af @ bf
        for out_dtype, acc in ((torch.bfloat16, False), (torch.float32, True)):
            c = torch.ones(M, N, device="cuda", dtype=out_dtype)
            if not acc: c.zero_()
            gemm(a, b, lay, out=c, accumulate=acc, two_cta=True)
            exp = ref + (1.0 if acc else 0.0)
            err = ((c.float() - exp).abs().max() / exp.abs().max()).item()
            good = err < 2e-2
            ok &= good
            if not good or (M, N, K) == (4096, 4096, 1024):
                print(f"{lay} M{M} N{N} K{K} out={str(out_dtype)[6:]} acc={acc}: rel err {err:.3e} {'OK' if good else 'FAIL'}", flush=True)
print("CORRECT" if ok else "MISMATCH", flush=True)
for lay, M, N, K, acc in [("TN", 8192, 8192, 8192, 0), ("TN", 4096, 28672, 4096, 0), ("TN", 4096, 4096, 14336, 0), ("TN", 4096, 6144, 4096, 0),
                          ("NN", 4096, 4096, 28672, 0), ("NN", 4096, 14336, 4096, 0), ("NT", 28672, 4096, 4096, 1), ("NT", 4096, 14336, 4096, 1),
                          ("TN", 4096, 128384, 4096, 0)]:
    ring = 3
    sets = [mk(lay, M, N, K)[:2] for _ in range(ring)]
    cs = [torch.zeros(M, N, device="cuda", dtype=torch.float32 if acc else torch.bfloat16) for _ in range(ring)]
    res = {}
    for name, flag in (("1cta", False), ("2cta", True), ("1cta_b", False), ("2cta_b", True)):
        i = [0]
        def fn():
            j = i[0] % ring; i[0] += 1
            gemm(sets[j][0], sets[j][1], lay, out=cs[j], accumulate=bool(acc), two_cta=flag)
        res[name] = 2.0 * M * N * K / timeit(fn) / 1e9
    print(f"{lay} M{M} N{N} K{K} acc{acc}:  1cta {res['1cta']:7.1f} {res['1cta_b']:7.1f}   2cta {res['2cta']:7.1f} {res['2cta_b']:7.1f} TF", flush=True)

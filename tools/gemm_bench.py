"""Quick GEMM timing vs torch.matmul (cuBLAS) on the current GPU.  CUDA events, L2 flushed between iters."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops.gemm import gemm

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); evs.append((s, e))
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]


shapes = [("TN", 4096, 10240, 8192), ("NN", 4096, 8192, 10240), ("NT", 10240, 8192, 4096), ("TN", 8192, 8192, 8192),
          ("TN", 4096, 768, 4096), ("TN", 4096, 4096, 512), ("TN", 4096, 160, 5120), ("TN", 4096, 28672, 8192),
          ("TN", 4096, 64128, 8192)]
rows = []
for layout, M, N, K in shapes:
    if layout == "TN":
        a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
        ref = lambda: torch.matmul(a, b.t())
    elif layout == "NN":
        a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16()
        ref = lambda: torch.matmul(a, b)
    else:
        a = torch.randn(K, M, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16()
        ref = lambda: torch.matmul(a.t(), b)
    fl = 2.0 * M * N * K
    t_ref = timeit(ref)
    res = {"layout": layout, "M": M, "N": N, "K": K, "cublas_tflops": fl / t_ref / 1e9}
    for bn in (0, 128, 256):
        try:
            t = timeit(lambda: gemm(a, b, layout, block_n=bn))
            res[f"smx_bn{bn}_tflops"] = fl / t / 1e9
        except Exception as ex:
            res[f"smx_bn{bn}_err"] = str(ex)
    print(json.dumps(res), flush=True)
    rows.append(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)

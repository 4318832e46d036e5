"""A/B the current GEMM kernel against an older build (simumax_b200/_gemm_old_ab.so) on the same box, interleaved."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simumax_b200.ops import _lib

new = _lib.lib()
old = ctypes.CDLL(os.path.join(os.path.dirname(_lib.__file__), "..", "_gemm_old_ab.so"))
for f in (new.smx_gemm, old.smx_gemm):
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 9 + [ctypes.c_float] + [ctypes.c_longlong] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
LAY = {"TN": (0, 0), "NN": (0, 1), "NT": (1, 1)}


def run(lib, a, b, c, lay, M, N, K, acc):
    am, bm = LAY[lay]
    rc = lib.smx_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), 1, M, N, K, am, bm, 0, 1 if c.dtype == torch.float32 else 0, acc, 1.0,
                      a.stride(0), b.stride(0), c.stride(0), 0, 0, 0, 0, 148, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


shapes = [("NN", 4096, 2048, 8192, 0), ("TN", 4096, 8192, 2048, 0), ("NN", 4096, 8192, 2560, 0), ("NT", 3072, 4096, 4096, 1),
          ("TN", 4096, 4096, 14336, 0), ("TN", 4096, 6144, 4096, 0), ("TN", 4096, 4096, 4096, 0), ("NN", 4096, 4096, 28672, 0),
          ("TN", 4096, 28672, 4096, 0), ("NT", 28672, 4096, 4096, 1), ("TN", 8192, 8192, 8192, 0)]
for lay, M, N, K, acc in shapes:
    am, bm = LAY[lay]
    ring = 4
    As = [torch.randn((K, M) if am else (M, K), device="cuda").bfloat16() for _ in range(ring)]
    Bs = [torch.randn((K, N) if bm else (N, K), device="cuda").bfloat16() for _ in range(ring)]
    Cs = [torch.zeros(M, N, device="cuda", dtype=torch.float32 if acc else torch.bfloat16) for _ in range(ring)]
    res = {}
    for rep in range(2):
        for name, lib in (("old", old), ("new", new)):
            i = [0]
            def fn():
                j = i[0] % ring; i[0] += 1
                run(lib, As[j], Bs[j], Cs[j], lay, M, N, K, acc)
            ms = timeit(fn)
            res.setdefault(name, []).append(2.0 * M * N * K / ms / 1e9)
    run(old, As[0], Bs[0], Cs[0], lay, M, N, K, 0); ref = Cs[0].float().clone()
    run(new, As[0], Bs[0], Cs[1], lay, M, N, K, 0)
    err = (Cs[1].float() - ref).abs().max().item()
    print(f"{lay} M{M} N{N} K{K} acc{acc}: old {res['old'][0]:7.1f} {res['old'][1]:7.1f}  new {res['new'][0]:7.1f} {res['new'][1]:7.1f} TF   maxdiff {err:.3g}", flush=True)

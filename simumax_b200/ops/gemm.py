"""tcgen05 GEMM front-end (bf16 / fp8-e4m3, plain and grouped).

Layout names follow the reference's calibration keys (``layout=TN|NN|NT``,
/root/reference/simu_tools/efficency_test/test_gemm_efficiency.py:90-101):

* ``TN``  fwd  :  out[M,N]  = a[M,K]  @ b[N,K]^T
* ``NN``  dgrad:  out[M,N]  = a[M,K]  @ b[K,N]          (b is the [out,in] weight read "untransposed")
* ``NT``  wgrad:  out[M,N] (+)= a[K,M]^T @ b[K,N]       (reduction over tokens; fp32 accumulate target)

All three run on the same kernel by telling tcgen05 whether each shared-memory operand tile is
K-major or MN-major; no transposed copy is ever materialised.
"""
from __future__ import annotations

import ctypes
from functools import lru_cache

import torch

from . import _lib

_LAYOUTS = {"TN": (0, 0), "NN": (0, 1), "NT": (1, 1), "TT": (1, 0)}


@lru_cache(maxsize=None)
def _num_sms(device_index: int) -> int:
    return torch.cuda.get_device_properties(device_index).multi_processor_count


def _fn():
    f = _lib.lib().smx_gemm
    if not getattr(f, "_configured", False):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 9 + [ctypes.c_float] + \
            [ctypes.c_longlong] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        f._configured = True
    return f


def _fn2():
    f = _lib.lib().smx_gemm_2cta
    if not getattr(f, "_configured", False):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_float] + \
            [ctypes.c_longlong] * 6 + [ctypes.c_int, ctypes.c_void_p]
        f._configured = True
    return f


def _two_cta_default() -> bool:
    import os
    return os.environ.get("SMX_GEMM_2CTA", "1") == "1"


def _pair_tiles_pay(M: int) -> bool:
    """256-row pair tiles vs 128-row tiles: take the pair kernel unless its last M-tile wastes much more of the tensor core
    (e.g. M = 320 rows per expert: 62 % useful rows in two pair tiles, 83 % in three single tiles)."""
    util2 = M / (-(-M // 256) * 256)
    util1 = M / (-(-M // 128) * 128)
    return util2 >= 0.88 * util1


def _as3d(t: torch.Tensor) -> torch.Tensor:
    return t if t.dim() == 3 else t.unsqueeze(0)


def _tma_ready(t3: torch.Tensor) -> torch.Tensor:
    """TMA needs a contiguous inner dim, 16-byte aligned base and 16-byte-multiple outer strides.

    Tensors that already satisfy this (every shape in the model zoo) pass through untouched; odd ones
    are copied once into a row-padded buffer and viewed back to their logical width.
    """
    es = t3.element_size()
    ok = (t3.stride(-1) == 1 and t3.data_ptr() % 16 == 0 and (t3.stride(1) * es) % 16 == 0
          and (t3.shape[0] == 1 or (t3.stride(0) * es) % 16 == 0))
    if ok:
        return t3
    per16 = 16 // es
    cols = t3.shape[-1]
    padded = (cols + per16 - 1) // per16 * per16
    buf = torch.zeros(t3.shape[0], t3.shape[1], padded, device=t3.device, dtype=t3.dtype)
    buf[:, :, :cols].copy_(t3)
    return buf[:, :, :cols]


def _pair_fills_machine(G: int, M: int, N: int, nsm: int) -> bool:
    """The pair kernel lights up 2 SMs per 256 x 256 tile; when the 128-row kernel would cover twice as many SMs (N = 256 / 512
    projections, 256-row wgrads) that is worth more than the pair's per-SM advantage.  Sustained-clock A/B on the 16 few-tile keys of
    the sweep (profiles/gemm_pair_fill_ab.json): single wins +5 ... +29 % where the pair reaches half the CTAs (N = 256 / 512,
    M = 256), the pair stays 3 - 8 % ahead at two thirds (N = 576 ... 768, M = 768) -- hence the 0.6."""
    import os
    if os.environ.get("SMX_GEMM_PAIR_FILL", "1") == "0":          # A/B switch (tools/gemm_pair_fill_ab.py)
        return True
    pair_ctas = min(2 * G * -(-M // 256) * -(-N // 256), nsm)
    single_ctas = min(G * -(-M // 128) * -(-N // 128), nsm)
    return pair_ctas >= 0.6 * single_ctas


def gemm(a: torch.Tensor, b: torch.Tensor, layout: str = "TN", out: torch.Tensor | None = None,
         out_dtype: torch.dtype | None = None, accumulate: bool = False, alpha: float = 1.0,
         block_n: int = 0, sm_limit: int = 0, two_cta: bool | None = None) -> torch.Tensor:
    """out (+)= alpha * op(a) @ op(b).  2-D inputs = one GEMM, 3-D inputs = grouped GEMM over dim 0.

    The innermost dimension of every operand must be contiguous; row strides only need 16-byte
    alignment (so column-sliced views work without a copy).
    """
    a_mn, b_mn = _LAYOUTS[layout]
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        dtype = 0
    elif a.dtype == torch.float8_e4m3fn and b.dtype == torch.float8_e4m3fn:
        dtype = 1
    else:
        raise TypeError(f"unsupported operand dtypes {a.dtype} x {b.dtype}")
    squeeze = a.dim() == 2
    a3, b3 = _as3d(a), _as3d(b)
    a3, b3 = _tma_ready(a3), _tma_ready(b3)
    G = a3.shape[0]
    if a_mn:
        K, M = a3.shape[1], a3.shape[2]
    else:
        M, K = a3.shape[1], a3.shape[2]
    if b_mn:
        Kb, N = b3.shape[1], b3.shape[2]
    else:
        N, Kb = b3.shape[1], b3.shape[2]
    if Kb != K or b3.shape[0] != G:
        raise ValueError(f"shape mismatch: a{tuple(a.shape)} b{tuple(b.shape)} layout {layout}")
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an `out` tensor")
        od = out_dtype or torch.bfloat16
        out = torch.empty((G, M, N), device=a.device, dtype=od)
        out3 = out
    else:
        out3 = _as3d(out)
        if out3.shape != (G, M, N) or out3.stride(-1) != 1:
            raise ValueError(f"bad out shape/stride {tuple(out.shape)} for G,M,N={G},{M},{N}")
    if out3.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out must be bf16 or fp32")
    stream = torch.cuda.current_stream(a.device).cuda_stream
    nsm = _num_sms(a.device.index if a.device.index is not None else torch.cuda.current_device())
    if sm_limit:
        nsm = min(nsm, sm_limit)
    if two_cta is None:
        two_cta = _two_cta_default()
    if two_cta and dtype == 0 and M >= 256 and N >= 256 and block_n == 0 and _pair_tiles_pay(M) and _pair_fills_machine(G, M, N, nsm):
        # CTA-pair kernel (csrc/gemm_2cta.cu): 256 x 256 tiles, B split across the pair
        rc = _fn2()(a3.data_ptr(), b3.data_ptr(), out3.data_ptr(), G, M, N, K, a_mn, b_mn,
                    1 if out3.dtype == torch.float32 else 0, 1 if accumulate else 0, float(alpha),
                    a3.stride(1), b3.stride(1), out3.stride(1),
                    a3.stride(0) if G > 1 else 0, b3.stride(0) if G > 1 else 0, out3.stride(0) if G > 1 else 0, nsm, stream)
        _lib.check(rc, f"gemm_2cta[{layout}] G{G} M{M} N{N} K{K}")
        _lib.count_launch()
        if out.dim() == 3 and squeeze and out is out3:
            return out3[0]
        return out
    rc = _fn()(a3.data_ptr(), b3.data_ptr(), out3.data_ptr(), G, M, N, K, a_mn, b_mn, dtype,
               1 if out3.dtype == torch.float32 else 0, 1 if accumulate else 0, float(alpha),
               a3.stride(1), b3.stride(1), out3.stride(1),
               a3.stride(0) if G > 1 else 0, b3.stride(0) if G > 1 else 0, out3.stride(0) if G > 1 else 0,
               block_n, nsm, stream)
    _lib.check(rc, f"gemm[{layout}] G{G} M{M} N{N} K{K}")
    _lib.count_launch()
    if out.dim() == 3 and squeeze and out is out3:
        return out3[0]
    return out


def gemm_flops(M: int, N: int, K: int, G: int = 1) -> float:
    return 2.0 * G * M * N * K


# ------------------------------------------------------------------------------------------------------------------
# MXFP8 (block-scaled fp8): e4m3 payload + one UE8M0 (power-of-two) scale per 32 consecutive K elements of each row
# ------------------------------------------------------------------------------------------------------------------
def mxfp8_quantize(x: torch.Tensor):
    """x [rows, K] (K % 32 == 0) -> (payload e4m3 [rows, K], scales uint8 [rows, K / 32]);  x ~= payload * 2^(scale - 127)."""
    rows, K = x.shape
    assert K % 32 == 0
    xb = x.float().view(rows, K // 32, 32)
    amax = xb.abs().amax(dim=-1).clamp_min(1e-30)
    e = torch.ceil(torch.log2(amax / 448.0)).clamp(-127, 127)
    q = (xb / torch.exp2(e)[..., None]).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(rows, K)
    return q.contiguous(), (e + 127).to(torch.uint8).contiguous()


def mxfp8_dequantize(q: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    rows, K = q.shape
    return (q.float().view(rows, K // 32, 32) * torch.exp2(scales.float() - 127.0)[..., None]).view(rows, K)


def gemm_mxfp8(a: torch.Tensor, sfa: torch.Tensor, b: torch.Tensor, sfb: torch.Tensor, out: torch.Tensor | None = None,
               out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """out[M, N] = dequant(a, sfa)[M, K] @ dequant(b, sfb)[N, K]^T with the scaling done inside tcgen05
    (``kind::mxf8f6f4.block_scale``, csrc/gemm_mxfp8.cu).  K must be a multiple of 128."""
    assert a.dtype == torch.float8_e4m3fn and b.dtype == torch.float8_e4m3fn and sfa.dtype == torch.uint8 and sfb.dtype == torch.uint8
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and K % 128 == 0 and sfa.shape == (M, K // 32) and sfb.shape == (N, K // 32)
    assert a.stride(1) == 1 and b.stride(1) == 1 and sfa.is_contiguous() and sfb.is_contiguous()
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    f = _lib.lib().smx_gemm_mxfp8
    if not getattr(f, "_configured", False):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_longlong] * 5 + [ctypes.c_int, ctypes.c_void_p]
        f._configured = True
    nsm = _num_sms(a.device.index if a.device.index is not None else torch.cuda.current_device())
    rc = f(a.data_ptr(), b.data_ptr(), out.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), M, N, K,
           1 if out.dtype == torch.float32 else 0, a.stride(0), b.stride(0), out.stride(0), sfa.stride(0), sfb.stride(0), nsm,
           torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, f"gemm_mxfp8 M{M} N{N} K{K}")
    _lib.count_launch()
    return out

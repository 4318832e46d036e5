"""Flash-attention front-end (csrc/fmha_fwd.cu): causal / full, GQA, MLA head dims (192/128).

Tensors are "bshd": q [B, S_q, H, D_qk], k [B, S_kv, H_kv, D_qk], v [B, S_kv, H_kv, D_v]; any batch/seq/head strides
with a contiguous, 16-byte aligned innermost dim are accepted (slices of a packed qkv projection need no copy).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _lib

_c = ctypes
_fwd = None


def _fwd_fn():
    global _fwd
    if _fwd is None:
        f = _lib.lib().smx_fmha_fwd
        f.restype = _c.c_int
        f.argtypes = [_c.c_void_p] * 5 + [_c.c_int] * 7 + [_c.c_longlong] * 12 + [_c.c_float, _c.c_int, _c.c_void_p]
        _fwd = f
    return _fwd


def _check(t: torch.Tensor, name: str):
    if t.dtype != torch.bfloat16 or t.dim() != 4 or t.stride(-1) != 1:
        raise TypeError(f"{name} must be a 4-D bf16 tensor with contiguous last dim")
    if t.data_ptr() % 16 or any((s * 2) % 16 for s in t.stride()[:-1]):
        raise ValueError(f"{name}: base pointer and outer strides must be 16-byte aligned for TMA")


def fmha_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True, scale: Optional[float] = None,
             out: Optional[torch.Tensor] = None, return_lse: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _check(t, n)
    B, S_q, H, D = q.shape
    _, S_kv, H_kv, Dk = k.shape
    Dv = v.shape[-1]
    assert Dk == D and v.shape[:3] == k.shape[:3] and k.shape[0] == B
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    o = torch.empty(B, S_q, H, Dv, device=q.device, dtype=q.dtype) if out is None else out
    lse = torch.empty(B, H, S_q, device=q.device, dtype=torch.float32) if return_lse else None
    rc = _fwd_fn()(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr() if lse is not None else None,
                   B, S_q, S_kv, H, H_kv, D, Dv,
                   q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
                   v.stride(0), v.stride(1), v.stride(2), o.stride(0), o.stride(1), o.stride(2),
                   float(scale), 1 if causal else 0, torch.cuda.current_stream(q.device).cuda_stream)
    _lib.check(rc, f"fmha_fwd B{B} S{S_q}x{S_kv} H{H}/{H_kv} d{D}/{Dv}")
    _lib.count_launch()
    return o, lse


def attention_flops(B, S_q, S_kv, H, d_qk, d_v, causal: bool, backward: bool = False) -> float:
    """Useful FLOPs (causal counted as half)."""
    f = 2.0 * B * H * S_q * S_kv * (d_qk + d_v)
    if causal:
        f *= 0.5
    return f * (2.5 if backward else 1.0)

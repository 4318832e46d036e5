"""Flash-attention front-end (csrc/fmha_fwd.cu): causal / full, GQA, MLA head dims (192/128).

Tensors are "bshd": q [B, S_q, H, D_qk], k [B, S_kv, H_kv, D_qk], v [B, S_kv, H_kv, D_v]; any batch/seq/head strides
with a contiguous, 16-byte aligned innermost dim are accepted (slices of a packed qkv projection need no copy).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional, Tuple

import torch

from . import _lib

_c = ctypes
_fwd = None
_fwd2 = None


def _fwd2_fn():
    global _fwd2
    if _fwd2 is None:
        f = _lib.lib().smx_fmha_fwd2
        f.restype = _c.c_int
        f.argtypes = [_c.c_void_p] * 5 + [_c.c_int] * 5 + [_c.c_longlong] * 12 + [_c.c_float, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]
        _fwd2 = f
    return _fwd2


def _use_v2(D: int, Dv: int) -> bool:
    """Second-generation forward (two query tiles per CTA in ping-pong, P in TMEM, 2-CTA cluster K / V multicast): head dim 128 only;
    SMX_FMHA_V1=1 forces the first-generation kernel for A/B runs."""
    return D == 128 and Dv == 128 and os.environ.get("SMX_FMHA_V1", "0") != "1"


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
             out: Optional[torch.Tensor] = None, return_lse: bool = True, lse: Optional[torch.Tensor] = None,
             trace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _check(t, n)
    B, S_q, H, D = q.shape
    _, S_kv, H_kv, Dk = k.shape
    Dv = v.shape[-1]
    assert Dk == D and v.shape[:3] == k.shape[:3] and k.shape[0] == B
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    o = torch.empty(B, S_q, H, Dv, device=q.device, dtype=q.dtype) if out is None else out
    if lse is None:
        lse = torch.empty(B, H, S_q, device=q.device, dtype=torch.float32) if return_lse else None
    else:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * H * S_q
    if _use_v2(D, Dv):
        rc = _fwd2_fn()(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr() if lse is not None else None,
                        B, S_q, S_kv, H, H_kv,
                        q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
                        v.stride(0), v.stride(1), v.stride(2), o.stride(0), o.stride(1), o.stride(2),
                        float(scale), 1 if causal else 0, int(os.environ.get("SMX_FMHA_CLUSTER", "0")),
                        trace.data_ptr() if trace is not None else None, torch.cuda.current_stream(q.device).cuda_stream)
        _lib.check(rc, f"fmha_fwd2 B{B} S{S_q}x{S_kv} H{H}/{H_kv} d{D}/{Dv}")
        _lib.count_launch()
        return o, lse
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


_bwd = None


def _bwd_fn():
    global _bwd
    if _bwd is None:
        f = _lib.lib().smx_fmha_bwd
        f.restype = _c.c_int
        f.argtypes = [_c.c_void_p] * 11 + [_c.c_int] * 6 + [_c.c_void_p, _c.c_float, _c.c_int, _c.c_void_p, _c.c_void_p]
        _bwd = f
    return _bwd


_bwd_ws = {}


def _bwd_workspace(device, B, S_q, H, D):
    """fp32 scratch (row sums of dO*O and the dQ accumulator), cached per shape so a training loop never reallocates."""
    key = (device, B, S_q, H, D)
    ws = _bwd_ws.get(key)
    if ws is None:
        ws = (torch.empty(2, B, H, (S_q + 63) // 64 * 64, device=device, dtype=torch.float32),
              torch.empty(B, S_q, H, D, device=device, dtype=torch.float32))
        _bwd_ws.clear()
        _bwd_ws[key] = ws
    return ws


def fmha_bwd(d_o: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor,
             dq: Optional[torch.Tensor] = None, dk: Optional[torch.Tensor] = None, dv: Optional[torch.Tensor] = None,
             causal: bool = True, scale: Optional[float] = None, trace: Optional[torch.Tensor] = None):
    """Backward of :func:`fmha_fwd` (csrc/fmha_bwd.cu; head dim 128). All tensors are bshd views; dq/dk/dv may be
    slices of a packed d_qkv buffer. Returns (dq, dk, dv)."""
    B, S_q, H, D = q.shape
    _, S_kv, H_kv, _ = k.shape
    if D not in (128, 192) or v.shape[-1] != 128:
        raise NotImplementedError("fmha_bwd: qk head dim 128 or 192 (MLA) with v head dim 128")
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    dq = torch.empty_like(q) if dq is None else dq
    dk = torch.empty_like(k) if dk is None else dk
    dv = torch.empty_like(v) if dv is None else dv
    # The kernel parallelises over (batch, KV head, key tile) and keeps dK / dV of a CTA in TMEM across the query heads of its GQA
    # group.  With few KV heads (tensor-parallel shards: 4 query heads on 1 KV head) that is a handful of CTAs on 148 SMs, so the
    # group is split: K / V are replicated e-fold (a few MB), every replica takes g / e of the query heads, and the e partial
    # dK / dV are summed afterwards.
    g = H // H_kv
    n_kv = (S_kv + 127) // 128
    ctas = B * H_kv * ((n_kv + 1) // 2 if causal else n_kv)
    if g > 1 and ctas < 120 and os.environ.get("SMX_FMHA_BWD_SPLIT", "1") == "1":
        e = next((c for c in range(2, g + 1) if g % c == 0 and ctas * c >= 120), g)
        k_e = k.unsqueeze(3).expand(B, S_kv, H_kv, e, D).reshape(B, S_kv, H_kv * e, D)
        v_e = v.unsqueeze(3).expand(B, S_kv, H_kv, e, v.shape[-1]).reshape(B, S_kv, H_kv * e, v.shape[-1])
        dk_e, dv_e = torch.empty_like(k_e), torch.empty_like(v_e)
        fmha_bwd(d_o, q, k_e, v_e, o, lse, dq, dk_e, dv_e, causal=causal, scale=scale, trace=trace)
        dk.copy_(dk_e.view(B, S_kv, H_kv, e, D).sum(3))
        dv.copy_(dv_e.view(B, S_kv, H_kv, e, v.shape[-1]).sum(3))
        return dq, dk, dv
    ts = (q, k, v, o, d_o, dq, dk, dv)
    for t, n in zip(ts, ("q", "k", "v", "o", "do", "dq", "dk", "dv")):
        _check(t, n)
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * H * S_q
    dsum, acc = _bwd_workspace(q.device, B, S_q, H, D)
    strides = (_c.c_longlong * 24)(*[s for t in ts for s in (t.stride(0), t.stride(1), t.stride(2))])
    rc = _bwd_fn()(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dq.data_ptr(),
                   dk.data_ptr(), dv.data_ptr(), dsum.data_ptr(), acc.data_ptr(), B, S_q, S_kv, H, H_kv, D,
                   _c.cast(strides, _c.c_void_p), float(scale), 1 if causal else 0,
                   torch.cuda.current_stream(q.device).cuda_stream, trace.data_ptr() if trace is not None else None)
    _lib.check(rc, f"fmha_bwd B{B} S{S_q}x{S_kv} H{H}/{H_kv} d{D}")
    _lib.count_launch(3)
    return dq, dk, dv

"""Fused compute + collective kernels over peer-mapped (symmetric) memory.

``gemm_reduce_scatter``: tensor-parallel row-linear forward (Megatron ``LinearRow`` with sequence parallelism, reference
cost model simumax/core/transformer/dense_module.py:646-695).  Every rank multiplies its K-shard, and the GEMM epilogue
adds each output tile straight into the owning rank's buffer with ``red.global.add.v4.f32`` over NVLink -- there is no
separate reduce-scatter pass and the transfer of tile *i* overlaps the MMAs of tile *i+1*.
"""
from __future__ import annotations

import ctypes

import torch

from ..ops import _lib
from ..ops.gemm import _LAYOUTS, _as3d, _fn, _num_sms, _tma_ready
from .symm import SymmetricWorkspace


def gemm_reduce_scatter(a: torch.Tensor, b: torch.Tensor, ws: SymmetricWorkspace, layout: str = "TN", out_off: int | None = None,
                        alpha: float = 1.0) -> torch.Tensor:
    """out_r[M/world, N] (fp32, in the symmetric buffer at ``out_off``) = sum over ranks of (op(a) @ op(b))[r-th row block].

    a / b are this rank's operands (K already sharded by the caller).  Returns this rank's output view; it is complete
    when the call returns control to the stream (a device-side cross-rank barrier closes the op)."""
    a_mn, b_mn = _LAYOUTS[layout]
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a3, b3 = _tma_ready(_as3d(a)), _tma_ready(_as3d(b))
    M, K = (a3.shape[2], a3.shape[1]) if a_mn else (a3.shape[1], a3.shape[2])
    N = b3.shape[2] if b_mn else b3.shape[1]
    world = ws.world
    assert M % world == 0 and N % 4 == 0
    out_off = ws.in_bytes if out_off is None else out_off
    rows = M // world
    out = ws.view(out_off, torch.float32, rows * N).view(rows, N)
    out.zero_()
    ws.barrier()                                   # every rank's buffer is cleared before anybody adds into it
    lib = _lib.lib()
    arm = lib.smx_gemm_arm_reduce_scatter
    arm.restype = ctypes.c_int
    arm.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    _lib.check(arm(ws.peer_ptrs(out_off), world), "arm_reduce_scatter")
    stream = torch.cuda.current_stream(a.device).cuda_stream
    nsm = _num_sms(a.device.index if a.device.index is not None else torch.cuda.current_device())
    rc = _fn()(a3.data_ptr(), b3.data_ptr(), out.data_ptr(), 1, M, N, K, a_mn, b_mn, 0, 1, 0, float(alpha),
               a3.stride(1), b3.stride(1), N, 0, 0, 0, 0, nsm, stream)
    _lib.check(rc, f"gemm_reduce_scatter[{layout}] M{M} N{N} K{K} world{world}")
    _lib.count_launch()
    ws.barrier()                                   # all peers' adds have landed
    return out

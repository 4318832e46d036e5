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


_RS_COUNTER_WORD, _RS_TICKET_WORD = 0, 1        # user signal words of the workspace used by the TMA-store flavour


class RsHandle:
    """An in-flight fused GEMM -> reduce-scatter: this rank's ``[world][rows, N]`` slot buffer and the counter value that
    means "every rank's stores have landed"."""

    def __init__(self, ws, slots, rows, n, target):
        self.ws, self.slots, self.rows, self.n, self.target = ws, slots, rows, n, target

    def consume(self, residual: torch.Tensor | None = None, out: torch.Tensor | None = None, gamma: torch.Tensor | None = None,
                normed: torch.Tensor | None = None, rstd: torch.Tensor | None = None, eps: float = 1e-5) -> torch.Tensor:
        """out = (residual +) sum of the slots; with ``gamma`` also ``normed = rmsnorm(out) * gamma`` (and ``rstd``) -- one
        kernel that itself waits on the completion counter (no barrier launch)."""
        ws, rows, n = self.ws, self.rows, self.n
        if out is None:
            out = torch.empty(rows, n, device=self.slots.device, dtype=torch.bfloat16)
        f = _lib.lib().smx_rs_consume
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int] + [ctypes.c_void_p] * 5 + [
            ctypes.c_float, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
        ptr = lambda t: None if t is None else t.data_ptr()
        if gamma is not None and normed is None:
            normed = torch.empty_like(out)
        _lib.check(f(self.slots.data_ptr(), ws.world, rows, n, ptr(residual), out.data_ptr(), ptr(gamma), ptr(normed), ptr(rstd),
                     float(eps), ws.user_flag(_RS_COUNTER_WORD).data_ptr(), self.target & 0xFFFFFFFF,
                     torch.cuda.current_stream(out.device).cuda_stream), "rs_consume")
        _lib.count_launch()
        return out


def gemm_reduce_scatter_launch(a: torch.Tensor, b: torch.Tensor, ws: SymmetricWorkspace, layout: str = "TN",
                               slot_off: int | None = None, alpha: float = 1.0) -> RsHandle:
    """Launch the CTA-pair GEMM whose epilogue TMA-stores every partial tile (bf16, through 128B-swizzled shared memory) into
    slot ``rank`` of the owning rank's staging buffer; returns a handle whose ``consume`` sums the slots.

    Two slot buffers (``2 * world * rows * N * 2`` bytes from ``slot_off``) alternate between calls, so a rank that runs
    ahead never overwrites slots its peer has not summed yet: launch -> consume -> launch ... is the only ordering rule.
    Layouts: TN (``b`` is [N, K]) and NN (``b`` is [K, N])."""
    a_mn, b_mn = _LAYOUTS[layout]
    assert not a_mn, "A must be K-major (row-major [M, K]) for the fused reduce-scatter GEMM"
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a3, b3 = _tma_ready(_as3d(a)), _tma_ready(_as3d(b))
    M, K = a3.shape[1], a3.shape[2]
    N = b3.shape[2] if b_mn else b3.shape[1]
    world = ws.world
    rows = M // world
    slot_off = ws.in_bytes if slot_off is None else slot_off
    one = world * rows * N * 2
    assert slot_off % 256 == 0 and slot_off + 2 * one <= ws.in_bytes + ws.out_bytes, "symmetric workspace too small for 2 slot buffers"
    calls = getattr(ws, "_rs_calls", 0)
    if calls == 0:
        ws.user_flag(_RS_COUNTER_WORD).zero_()
        ws.user_flag(_RS_TICKET_WORD).zero_()
        ws.barrier()                                   # nobody increments a counter that is still being cleared
    ws._rs_calls = calls + 1
    off = slot_off + (calls & 1) * one
    f = _lib.lib().smx_gemm_2cta_reduce_scatter
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p,
                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                  ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    nsm = _num_sms(a.device.index if a.device.index is not None else torch.cuda.current_device())
    rc = f(a3.data_ptr(), b3.data_ptr(), ws.peer_ptrs(off), ws.user_flag_ptrs(_RS_COUNTER_WORD),
           ws.user_flag(_RS_TICKET_WORD).data_ptr(), world, ws.rank, M, N, K, b_mn, float(alpha), a3.stride(1), b3.stride(1), nsm,
           torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, f"gemm_reduce_scatter[tma,{layout}] M{M} N{N} K{K} world{world}")
    _lib.count_launch()
    slots = ws.view(off, torch.bfloat16, world * rows * N)
    return RsHandle(ws, slots, rows, N, (calls + 1) * world)


def gemm_reduce_scatter(a: torch.Tensor, b: torch.Tensor, ws: SymmetricWorkspace, layout: str = "TN", out_off: int | None = None,
                        alpha: float = 1.0, mode: str = "slots", out: torch.Tensor | None = None) -> torch.Tensor:
    """out_r[M/world, N] = sum over ranks of (op(a) @ op(b))[r-th row block].

    a / b are this rank's operands (K already sharded by the caller).  Epilogue flavours:

    * ``mode="tma"``: CTA-pair GEMM, tiles leave through shared memory and TMA tensor stores (full NVLink packets), the
      consumer waits on a completion counter -- see :func:`gemm_reduce_scatter_launch`.  ``out_off`` must leave room for two
      slot buffers.

    * ``mode="slots"`` (default, bf16): every rank stores its partial tile with plain 16-byte stores into slot ``rank`` of
      the owner's ``[world][M/world, N]`` staging buffer (peer-mapped memory); after the closing device-side barrier the
      owner sums the slots (fp32 accumulate) into a local bf16 tensor.  NVLink carries bf16 once, no atomics.
    * ``mode="red"`` (fp32): tiles are added into the owner's fp32 buffer with ``red.global.add.v4.f32`` -- single pass,
      but 16-byte remote atomics are much slower than stores on NVLink (kept for comparison).
    """
    if mode == "tma":
        return gemm_reduce_scatter_launch(a, b, ws, layout, out_off, alpha).consume(out=out)
    a_mn, b_mn = _LAYOUTS[layout]
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a3, b3 = _tma_ready(_as3d(a)), _tma_ready(_as3d(b))
    M, K = (a3.shape[2], a3.shape[1]) if a_mn else (a3.shape[1], a3.shape[2])
    N = b3.shape[2] if b_mn else b3.shape[1]
    world = ws.world
    assert M % world == 0 and N % 8 == 0
    out_off = ws.in_bytes if out_off is None else out_off
    rows = M // world
    lib = _lib.lib()
    arm = lib.smx_gemm_arm_reduce_scatter
    arm.restype = ctypes.c_int
    arm.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]
    stream = torch.cuda.current_stream(a.device).cuda_stream
    nsm = _num_sms(a.device.index if a.device.index is not None else torch.cuda.current_device())
    if mode == "red":
        res = ws.view(out_off, torch.float32, rows * N).view(rows, N)
        res.zero_()
        ws.barrier()                               # every rank's buffer is cleared before anybody adds into it
        _lib.check(arm(ws.peer_ptrs(out_off), world, ws.rank), "arm_reduce_scatter")
        rc = _fn()(a3.data_ptr(), b3.data_ptr(), res.data_ptr(), 1, M, N, K, a_mn, b_mn, 0, 1, 0, float(alpha),
                   a3.stride(1), b3.stride(1), N, 0, 0, 0, 0, nsm, stream)
        _lib.check(rc, f"gemm_reduce_scatter[red,{layout}] M{M} N{N} K{K} world{world}")
        _lib.count_launch()
        ws.barrier()                               # all peers' adds have landed
        return res
    slots = ws.view(out_off, torch.bfloat16, world * rows * N)
    ws.barrier()                                   # every owner has finished summing the previous call's slots
    _lib.check(arm(ws.peer_ptrs(out_off), world, ws.rank), "arm_reduce_scatter")
    rc = _fn()(a3.data_ptr(), b3.data_ptr(), slots.data_ptr(), 1, M, N, K, a_mn, b_mn, 0, 0, 0, float(alpha),
               a3.stride(1), b3.stride(1), N, 0, 0, 0, 0, nsm, stream)
    _lib.check(rc, f"gemm_reduce_scatter[slots,{layout}] M{M} N{N} K{K} world{world}")
    _lib.count_launch()
    ws.barrier()                                   # every peer's stores into my slots have landed
    res = torch.empty(rows, N, device=a.device, dtype=torch.bfloat16) if out is None else out
    f = lib.smx_sum_slots
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    _lib.check(f(slots.data_ptr(), res.data_ptr(), rows * N, world, stream), "sum_slots")
    _lib.count_launch()
    return res


_AG_FLAG_WORD = 16            # first of world * chunks user signal words used by the single-kernel all-gather GEMM
_AG_MAX_CHUNKS = 128


def ag_views(ws: SymmetricWorkspace, a_off: int, M: int, K: int):
    """(full [M, K] view, own-rows [M / world, K] view) of a gathered activation buffer at byte offset ``a_off``: producers
    write their shard straight into the own-rows view, :func:`allgather_gemm_inkernel` fills in the rest."""
    full = ws.view(a_off, torch.bfloat16, M * K).view(M, K)
    ms = M // ws.world
    return full, full[ws.rank * ms:(ws.rank + 1) * ms]


def allgather_gemm_inkernel(ws: SymmetricWorkspace, a_off: int, M: int, K: int, w: torch.Tensor, layout: str = "TN",
                            out: torch.Tensor | None = None, out_dtype=torch.bfloat16, alpha: float = 1.0) -> torch.Tensor:
    """y[M, N] = gathered(A)[M, K] @ op(w) as ONE kernel: two extra warps in every CTA of the CTA-pair GEMM push this rank's
    rows (already in place at ``a_off``) into every peer's buffer (``multimem.st`` over the NVLS multicast alias when there is
    one) and raise per-chunk flags; the TMA producer starts on the local rows and waits on a remote block's flags only when it
    gets there.  No separate gather kernel, no barrier, no copy.  The gathered buffer stays valid afterwards (wgrad reuses it).

    Reuse rule: a buffer may be pushed into again only after every rank has passed a later cooperative call or a barrier
    (alternate between two buffers, or use one buffer per layer plus a barrier per pass, as runtime/tp_trainer.py does)."""
    a_mn, b_mn = _LAYOUTS[layout]
    assert not a_mn and w.dtype == torch.bfloat16
    world, rank = ws.world, ws.rank
    b3 = _tma_ready(_as3d(w))
    N = b3.shape[2] if b_mn else b3.shape[1]
    ms = M // world
    chunks = _AG_MAX_CHUNKS
    while ms % chunks or ((ms // chunks) * K * 2) % 16:
        chunks //= 2
    assert a_off % 1024 == 0 and a_off + M * K * 2 <= ws.in_bytes + ws.out_bytes
    ws._ag2_epoch = getattr(ws, "_ag2_epoch", 0) + 1
    y = torch.empty(M, N, device=w.device, dtype=out_dtype) if out is None else out
    assert y.stride(1) == 1 and y.dtype in (torch.bfloat16, torch.float32)
    f = _lib.lib().smx_gemm_2cta_all_gather
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                  ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_void_p]
    nsm = _num_sms(w.device.index if w.device.index is not None else torch.cuda.current_device())
    mc = (ws.multicast_ptr + a_off) if (ws.has_nvls and getattr(ws, "ag_use_nvls", True)) else None
    rc = f(ws.peer_ptrs(a_off), mc, ws.user_flag_ptrs(_AG_FLAG_WORD), b3.data_ptr(), y.data_ptr(), world, rank, M, N, K, b_mn,
           1 if y.dtype == torch.float32 else 0, float(alpha), b3.stride(1), y.stride(0), chunks, ws._ag2_epoch & 0xFFFFFFFF, nsm,
           torch.cuda.current_stream(w.device).cuda_stream)
    _lib.check(rc, f"allgather_gemm_inkernel[{layout}] M{M} N{N} K{K} world{world}")
    _lib.count_launch()
    return y


def allgather_gemm(x_shard: torch.Tensor, w: torch.Tensor, ws: SymmetricWorkspace, layout: str = "TN", chunks: int = 32,
                   out: torch.Tensor | None = None, comm_stream: torch.cuda.Stream | None = None) -> torch.Tensor:
    """y[M, N] = all_gather(x_shard over ranks)[M, K] @ w^T   (column-parallel linear forward with sequence parallelism,
    reference cost model dense_module.py:351-416) with the gather overlapped into the GEMM.

    A chunked push all-gather runs on ``comm_stream`` (every CTA pushes one row chunk of the local shard to all peers and
    raises a per-chunk flag); the GEMM starts immediately on this rank's own rows and its TMA producer waits on the flag
    of each remote row block only when it gets there -- communication of chunk *c* overlaps the MMAs of the chunks that
    arrived earlier.  Workspace layout: [x_shard | 4 KiB flags | gathered x (world x shard)]."""
    assert layout == "TN" and x_shard.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x_shard.is_contiguous()
    world, rank = ws.world, ws.rank
    ms, K = x_shard.shape
    N = w.shape[0]
    M = ms * world
    assert ms % (chunks * 1) == 0 and (ms // chunks) * K * 2 % 16 == 0 and world * chunks * 4 <= 4096
    shard_bytes = ms * K * 2
    flag_off = (shard_bytes + 255) // 256 * 256
    xg_off = flag_off + 4096
    assert xg_off + world * shard_bytes <= ws.in_bytes + ws.out_bytes, "symmetric workspace too small"
    main = torch.cuda.current_stream(x_shard.device)
    comm = comm_stream or getattr(ws, "_ag_stream", None)
    if comm is None:
        comm = ws._ag_stream = torch.cuda.Stream(device=x_shard.device)
    if not getattr(ws, "_ag_ready", False):
        ws.view(flag_off, torch.int32, 1024).zero_()
        ws._ag_epoch = 0
        ws._ag_ready = True
        ws.barrier()
    ws._ag_epoch += 1
    epoch = ws._ag_epoch
    xg = ws.view(xg_off, torch.bfloat16, M * K).view(M, K)
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        ws.view(0, torch.bfloat16, ms * K).view(ms, K).copy_(x_shard)
        ws.barrier()                      # every rank is done reading the previous gathered buffer
        f = _lib.lib().smx_all_gather_chunked
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.POINTER(ctypes.c_void_p)] * 2 + [ctypes.c_int] * 2 + [ctypes.c_longlong] * 4 + [ctypes.c_uint, ctypes.c_int,
                                                                                                         ctypes.c_void_p]
        _lib.check(f(ws.data_ptrs, ws.flag_ptrs, rank, world, 0, xg_off, flag_off, shard_bytes, epoch, chunks,
                     torch.cuda.current_stream(x_shard.device).cuda_stream), "all_gather_chunked")
        _lib.count_launch()
    y = torch.empty(M, N, device=x_shard.device, dtype=torch.bfloat16) if out is None else out
    lib = _lib.lib()
    flags = ws.view(flag_off, torch.int32, 1024)
    a3, b3 = _tma_ready(_as3d(xg)), _tma_ready(_as3d(w))
    nsm = _num_sms(x_shard.device.index if x_shard.device.index is not None else torch.cuda.current_device())
    pair_ok = M >= 256 and N >= 256 and (rank * ms) % 256 == 0
    if pair_ok:                           # CTA-pair kernel: first tile index counts 256-row tiles
        from ..ops.gemm import _fn2
        arm = lib.smx_gemm_2cta_arm_all_gather
        arm.restype = ctypes.c_int
        arm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        _lib.check(arm(flags.data_ptr(), ms // chunks, rank * ms // 256, epoch), "arm_all_gather_2cta")
        rc = _fn2()(a3.data_ptr(), b3.data_ptr(), y.data_ptr(), 1, M, N, K, 0, 0, 0, 0, 1.0, a3.stride(1), b3.stride(1), y.stride(0),
                    0, 0, 0, nsm, main.cuda_stream)
    else:
        arm = lib.smx_gemm_arm_all_gather
        arm.restype = ctypes.c_int
        arm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        _lib.check(arm(flags.data_ptr(), ms // chunks, rank * ms // 128, epoch), "arm_all_gather")
        rc = _fn()(a3.data_ptr(), b3.data_ptr(), y.data_ptr(), 1, M, N, K, 0, 0, 0, 0, 0, 1.0, a3.stride(1), b3.stride(1), y.stride(0),
                   0, 0, 0, 0, nsm, main.cuda_stream)
    _lib.check(rc, f"allgather_gemm M{M} N{N} K{K} world{world}")
    _lib.count_launch()
    main.wait_stream(comm)                # the gather kernel itself must be complete before the workspace is reused
    return y

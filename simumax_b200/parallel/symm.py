"""Symmetric (peer-mapped) workspace + in-kernel collectives (csrc/collectives.cu).

``SymmetricWorkspace`` allocates one data buffer and one flag buffer per rank through
``torch.distributed._symmetric_memory`` (CUDA VMM + fabric/posix handle exchange under the hood), which gives every
rank the peer pointers of all ranks and -- on NVSwitch systems -- an NVLS multicast address.  The collectives are
our own kernels operating on those pointers; NCCL is only used for bootstrap (rendezvous) and as the baseline curve.

Layout of the data buffer: ``[ in region : in_bytes ][ out region : out_bytes ]``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from ..ops import _lib

_c = ctypes
MAX_CTAS = 128
MAX_RANKS = 8
USER_FLAG_BASE = MAX_CTAS * MAX_RANKS      # first int32 word of the signal buffer that the collectives do not use
USER_FLAG_WORDS = 2048


def _fn(name, argtypes):
    f = getattr(_lib.lib(), name)
    if not getattr(f, "_configured", False):
        f.restype = _c.c_int
        f.argtypes = argtypes
        f._configured = True
    return f


_PTRS = _c.POINTER(_c.c_void_p)
_SIG_COLL = [_PTRS, _PTRS, _c.c_void_p, _c.c_int, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_longlong, _c.c_uint,
             _c.c_int, _c.c_int, _c.c_void_p]
_SIG_RED = _SIG_COLL + [_c.c_int]
_SIG_A2A = [_PTRS, _PTRS, _c.c_int, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_longlong, _c.c_uint, _c.c_int, _c.c_void_p]
_SIG_PUT = [_PTRS, _PTRS, _c.c_int, _c.c_int, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_longlong, _c.c_int, _c.c_void_p]


class SymmetricWorkspace:
    def __init__(self, in_bytes: int, out_bytes: int, group: Optional[dist.ProcessGroup] = None, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group or dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        assert self.world <= MAX_RANKS
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        align = lambda n: (n + 255) // 256 * 256
        self.in_bytes, self.out_bytes = align(in_bytes), align(out_bytes)
        try:
            symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self.data = symm_mem.empty(self.in_bytes + self.out_bytes, dtype=torch.uint8, device=self.device)
        self.flags = symm_mem.empty(MAX_CTAS * MAX_RANKS + USER_FLAG_WORDS, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self._hd = symm_mem.rendezvous(self.data, self.group.group_name)
        self._hf = symm_mem.rendezvous(self.flags, self.group.group_name)
        self.data_ptrs = (_c.c_void_p * MAX_RANKS)(*[int(p) for p in self._hd.buffer_ptrs])
        self.flag_ptrs = (_c.c_void_p * MAX_RANKS)(*[int(p) for p in self._hf.buffer_ptrs])
        mc = getattr(self._hd, "multicast_ptr", 0) or 0
        self.multicast_ptr = int(mc)
        self.has_nvls = self.multicast_ptr != 0
        self.epoch = 1
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    # ---- typed views
    def input(self, dtype, numel) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        assert nbytes <= self.in_bytes
        return self.data[:nbytes].view(dtype)

    def output(self, dtype, numel) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        assert nbytes <= self.out_bytes
        return self.data[self.in_bytes:self.in_bytes + nbytes].view(dtype)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _next_epoch(self) -> int:
        e = self.epoch
        # full uint32 range: the device compares (int32)(flag - epoch) < 0, which is wrap-safe; odd values only, so never 0
        self.epoch = (self.epoch + 2) & 0xFFFFFFFF
        return e

    def _coll(self, name, nbytes, use_nvls, ctas, in_off=None, out_off=None, f32=None):
        """``in_off`` / ``out_off`` are byte offsets into the symmetric buffer (defaults: the in / out regions)."""
        in_off = 0 if in_off is None else in_off
        out_off = self.in_bytes if out_off is None else out_off
        assert in_off % 16 == 0 and out_off % 16 == 0 and nbytes % 16 == 0
        args = [self.data_ptrs, self.flag_ptrs, self.multicast_ptr if self.has_nvls else None, self.rank, self.world,
                in_off, out_off, nbytes, self._next_epoch(), 1 if (use_nvls and self.has_nvls) else 0,
                min(ctas, MAX_CTAS), self._stream()]
        if f32 is None:
            rc = _fn(name, _SIG_COLL)(*args)
        else:
            rc = _fn(name, _SIG_RED)(*args, 1 if f32 else 0)
        _lib.check(rc, name)
        _lib.count_launch()

    def view(self, byte_off: int, dtype, numel: int) -> torch.Tensor:
        """Typed view of an arbitrary range of the local symmetric buffer."""
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        assert byte_off % 16 == 0 and byte_off + nbytes <= self.in_bytes + self.out_bytes
        return self.data[byte_off:byte_off + nbytes].view(dtype)

    # ---- variant policy: what the 2 / 4 / 8-GPU sweeps (profiles/coll_sweep_*gpu.json, fastest of P2P / NVLS x 32 / 64 / 128 CTAs
    # per message size) boil down to.  ``ctas=0`` / ``use_nvls=None`` in the calls below ask for it, so an exposed collective in the
    # runtime launches exactly the variant the calibrated curve was fitted on; overlapped ones pass a small explicit CTA count.
    def pick(self, op: str, total_bytes: int):
        """(use_nvls, ctas) for a stand-alone collective moving ``total_bytes`` (gathered / reduced size)."""
        big = total_bytes >= (16 << 20)
        if op == "all_gather":       # pushes: plain peer stores beat multimem.st once the message is large
            return (total_bytes < (4 << 20)), (32 if total_bytes < (4 << 20) else (64 if not big else 128))
        if op == "reduce_scatter":   # in-switch reduction (multimem.ld_reduce) wins at every size
            return True, (32 if total_bytes < (8 << 20) else 128)
        if op == "all_reduce":
            return True, (32 if total_bytes < (48 << 20) else 64)
        return False, (32 if total_bytes < (8 << 20) else 128)       # all_to_all

    # ---- collectives (operate in place on the workspace regions)
    def all_gather(self, shard_bytes: int, use_nvls=None, ctas: int = 0, in_off=None, out_off=None):
        """out[r*shard:(r+1)*shard] on every rank <- in[:shard] of rank r."""
        if in_off is None and out_off is None:
            assert shard_bytes <= self.in_bytes and shard_bytes * self.world <= self.out_bytes
        nv, c = self.pick("all_gather", shard_bytes * self.world)
        self._coll("smx_all_gather", shard_bytes, nv if use_nvls is None else use_nvls, ctas or c, in_off, out_off)

    def reduce_scatter(self, shard_bytes: int, use_nvls=None, ctas: int = 0, in_off=None, out_off=None,
                       dtype=torch.bfloat16):
        """out[:shard] <- sum over ranks of in[rank*shard:(rank+1)*shard]   (bf16 with fp32 accumulation, or fp32)."""
        if in_off is None and out_off is None:
            assert shard_bytes * self.world <= self.in_bytes and shard_bytes <= self.out_bytes
        nv, c = self.pick("reduce_scatter", shard_bytes * self.world)
        self._coll("smx_reduce_scatter", shard_bytes, nv if use_nvls is None else use_nvls, ctas or c, in_off, out_off,
                   f32=(dtype == torch.float32))

    def all_reduce(self, total_bytes: int, use_nvls=None, ctas: int = 0, in_off=None, out_off=None,
                   dtype=torch.bfloat16):
        """out[:total] <- sum over ranks of in[:total]   (bf16 or fp32)."""
        if in_off is None and out_off is None:
            assert total_bytes <= self.in_bytes and total_bytes <= self.out_bytes
        nv, c = self.pick("all_reduce", total_bytes)
        self._coll("smx_all_reduce", total_bytes, nv if use_nvls is None else use_nvls, ctas or c, in_off, out_off,
                   f32=(dtype == torch.float32))

    def all_to_all(self, chunk_bytes: int, ctas: int = 0):
        """out[src*chunk:(src+1)*chunk] of rank dst <- in[dst*chunk:(dst+1)*chunk] of rank src."""
        assert chunk_bytes * self.world <= min(self.in_bytes, self.out_bytes)
        ctas = ctas or self.pick("all_to_all", chunk_bytes * self.world)[1]
        rc = _fn("smx_all_to_all", _SIG_A2A)(self.data_ptrs, self.flag_ptrs, self.rank, self.world, 0, self.in_bytes,
                                             chunk_bytes, self._next_epoch(), min(ctas, MAX_CTAS), self._stream())
        _lib.check(rc, "smx_all_to_all")
        _lib.count_launch()

    def moe_dispatch(self, x: torch.Tensor, row_map: torch.Tensor, out_off: int, capacity: int, local_experts: int,
                     row_scale: Optional[torch.Tensor] = None, ctas: int = 128):
        """Fused token permutation + EP all-to-all: row (d, e, c) of the capacity-padded permuted order is gathered from local
        token ``row_map[r]`` (-1 = padding) and stored into rank d's buffer at ``out_off`` in expert-major ``[El][world * C]``
        order.  ``x`` may live anywhere in local memory; the destination is the symmetric buffer."""
        H = x.shape[-1]
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and row_map.dtype == torch.int32
        assert row_map.numel() == self.world * local_experts * capacity
        assert out_off + self.world * local_experts * capacity * H * 2 <= self.in_bytes + self.out_bytes
        sig = [_PTRS, _PTRS, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int,
               _c.c_uint, _c.c_int, _c.c_void_p]
        rc = _fn("smx_moe_dispatch", sig)(self.data_ptrs, self.flag_ptrs, self.rank, self.world, x.data_ptr(), row_map.data_ptr(),
                                          row_scale.data_ptr() if row_scale is not None else None, out_off, H, capacity, local_experts,
                                          self._next_epoch(), min(ctas, MAX_CTAS), self._stream())
        _lib.check(rc, "smx_moe_dispatch")
        _lib.count_launch()

    def moe_combine(self, y: torch.Tensor, out_off: int, capacity: int, local_experts: int, ctas: int = 128):
        """The way back: row (e, s, c) of the local expert-major tensor ``y`` goes to rank s, row ``(rank * El + e) * C + c`` of its
        permuted-order buffer at ``out_off``."""
        H = y.shape[-1]
        assert y.dtype == torch.bfloat16 and y.is_contiguous() and y.numel() == self.world * local_experts * capacity * H
        sig = [_PTRS, _PTRS, _c.c_int, _c.c_int, _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_int, _c.c_int, _c.c_uint, _c.c_int, _c.c_void_p]
        rc = _fn("smx_moe_combine", sig)(self.data_ptrs, self.flag_ptrs, self.rank, self.world, y.data_ptr(), out_off, H, capacity,
                                         local_experts, self._next_epoch(), min(ctas, MAX_CTAS), self._stream())
        _lib.check(rc, "smx_moe_combine")
        _lib.count_launch()

    def barrier(self):
        """Device-side barrier across the group on the current stream (flag exchange over NVLink, no NCCL)."""
        rc = _fn("smx_rank_barrier", [_PTRS, _PTRS, _c.c_int, _c.c_int, _c.c_uint, _c.c_void_p])(
            self.data_ptrs, self.flag_ptrs, self.rank, self.world, self._next_epoch(), self._stream())
        _lib.check(rc, "smx_rank_barrier")
        _lib.count_launch()

    def peer_ptrs(self, byte_off: int = 0):
        """ctypes array of every rank's mapping of ``buffer + byte_off`` (for kernels that write to peers directly)."""
        return (_c.c_void_p * MAX_RANKS)(*[(int(p) + byte_off) for p in self._hd.buffer_ptrs] + [0] * (MAX_RANKS - self.world))

    def user_flag_ptrs(self, word: int):
        """ctypes array of every rank's mapping of user signal word ``word`` (fused kernels' completion counters)."""
        assert 0 <= word < USER_FLAG_WORDS
        off = (USER_FLAG_BASE + word) * 4
        return (_c.c_void_p * MAX_RANKS)(*[(int(p) + off) for p in self._hf.buffer_ptrs] + [0] * (MAX_RANKS - self.world))

    def user_flag(self, word: int) -> torch.Tensor:
        """Local view (1 element, int32) of user signal word ``word``."""
        return self.flags[USER_FLAG_BASE + word: USER_FLAG_BASE + word + 1]

    def put(self, dst_rank: int, nbytes: int, ctas: int = 16):
        """One-sided copy in[:nbytes] -> out[:nbytes] of ``dst_rank`` (pipeline send); caller orders with events."""
        rc = _fn("smx_p2p_put", _SIG_PUT)(self.data_ptrs, self.flag_ptrs, self.rank, self.world, dst_rank, 0,
                                          self.in_bytes, nbytes, min(ctas, MAX_CTAS), self._stream())
        _lib.check(rc, "smx_p2p_put")
        _lib.count_launch()

"""Pipeline-parallel Llama training step: Megatron's non-interleaved 1F1B schedule over this repo's kernels.

The measured counterpart of the analytical 1F1B scheduler (``core/schedule.py``; reference ``calculate_1f1b_bubble``,
simumax/core/perf_llm.py:2097-2306) and of the simulator's ``PpSchedule`` (transformer/pipeline_schedule.py:717-959).
Every rank is one stage (``LlamaTrainer`` in stage mode: its slice of the layers, the embedding on the first stage, final
norm + LM head + fused cross-entropy on the last) and keeps ``pp - stage`` micro-batches in flight.  Activations and their
gradients travel as NCCL point-to-point operations batched the way Megatron batches them (send-forward / recv-backward pairs
go out together, so neighbouring stages never wait on each other in opposite orders); they are issued on a side stream and the
compute stream only waits on the event of the tensor it needs next -- the configuration the calibration times
(``calib/collectives.py::measure_p2p_overlap``).

Public API mirrors ``LlamaTrainer``: ``PPLlamaTrainer(dims, micro_batch, seq_len, pp_group)``, ``train_step(host_batches)``,
``run_microbatches(device_batches)``; every stage is given the same batches (only the first reads the tokens, only the last the
targets) and the loss is broadcast from the last stage.
"""
from __future__ import annotations

from copy import copy
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .trainer import LlamaDims


_GROUPS: dict = {}


def split_layers(n_layers: int, pp: int) -> List[Tuple[int, int]]:
    """Even split, remainder to the earliest stages (Megatron's default uneven split puts the short stage last; the analytical
    model takes the split from the strategy's ``num_layers_in_*_pipeline_stage``)."""
    base, extra = divmod(n_layers, pp)
    out, lo = [], 0
    for s in range(pp):
        n = base + (1 if s < extra else 0)
        out.append((lo, lo + n))
        lo += n
    return out


class PPLlamaTrainer:
    def __init__(self, dims: LlamaDims, micro_batch: int, seq_len: int, pp_group: Optional[dist.ProcessGroup] = None,
                 seed: int = 1234, device=None, stage_factory=None, act_dtype=torch.bfloat16,
                 dp_group: Optional[dist.ProcessGroup] = None, **trainer_kw):
        """``stage_factory(local_dims, first, last, in_flight)`` replaces the ``LlamaTrainer`` stage (the CPU / gloo test of
        the schedule plugs in a small autograd stage); a stage needs ``T``, ``use_slot``, ``forward``, ``backward``,
        ``optimizer_step`` and accepts a ``norm_group`` attribute.  ``dp_group``: the data-parallel replicas of THIS stage
        (PP x DP): the stage's ZeRO-1 optimizer shards over them; without it the stage runs unsharded."""
        assert dist.is_initialized()
        self.group = pp_group or dist.group.WORLD
        self.pp = dist.get_world_size(self.group)
        self.stage = dist.get_rank(self.group)
        self.ranks = dist.get_process_group_ranks(self.group)
        self.first, self.last = self.stage == 0, self.stage == self.pp - 1
        self.full_dims = dims
        self.lo, self.hi = split_layers(dims.layers, self.pp)[self.stage]
        local = copy(dims)
        local.layers = self.hi - self.lo
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.on_gpu = self.device.type == "cuda"
        in_flight = self.pp - self.stage
        if stage_factory is not None:
            self.tr = stage_factory(local, self.first, self.last, in_flight)
        else:
            from .trainer import LlamaTrainer
            self.tr = LlamaTrainer(local, micro_batch, seq_len, dp_group=dp_group, data_parallel=dp_group is not None,
                                   seed=seed + 31 * self.stage, device=self.device,
                                   first_stage=self.first, last_stage=self.last, in_flight=in_flight, **trainer_kw)
        self.tr.norm_group = self.group
        self.T, self.H = self.tr.T, dims.hidden
        bf = dict(dtype=act_dtype, device=self.device)
        self.recv_x = [torch.empty(self.T, self.H, **bf) for _ in range(in_flight)] if not self.first else []
        self.recv_dx = torch.empty(self.T, self.H, **bf) if not self.last else None
        self.send_dx = torch.empty(self.T, self.H, **bf) if not self.first else None
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.on_gpu else None
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=self.device)

    @classmethod
    def from_case(cls, case, dims: LlamaDims, seed: int = 1234, **kw) -> "PPLlamaTrainer":
        """pp x dp layout of a ``runtime.cases.Case`` (tp = 1), rank order as in the analytical model (dp faster than pp):
        rank = stage * dp + replica.  ``dist.new_group`` is collective, so every rank creates every group."""
        world, rank = dist.get_world_size(), dist.get_rank()
        pp, dp = case.pp, world // case.pp
        key = ("pp", pp, dp)
        if key not in _GROUPS:
            pp_groups = [dist.new_group([s * dp + j for s in range(pp)]) for j in range(dp)]
            dp_groups = [dist.new_group([s * dp + j for j in range(dp)]) for s in range(pp)] if dp > 1 else [None] * pp
            _GROUPS[key] = (pp_groups, dp_groups)
        pp_groups, dp_groups = _GROUPS[key]
        return cls(dims, case.mbs, case.seq, pp_group=pp_groups[rank % dp], dp_group=dp_groups[rank // dp], seed=seed,
                   recompute=case.recompute, **kw)

    # ------------------------------------------------------------------------------------------ weights
    def load_full_weights(self, full: dict) -> None:
        """Copy this stage's layers (and embedding / head) out of full ``LlamaTrainer.w``-style weights."""
        w = self.tr.w
        if self.first:
            w["embed"].copy_(full["embed"])
        if self.last:
            w["final_norm"].copy_(full["final_norm"])
            w["lm_head"].copy_(full["lm_head"])
        for j, i in enumerate(range(self.lo, self.hi)):
            for name in ("ln1", "wqkv", "wo", "ln2", "wfc1", "wfc2"):
                w[name][j].copy_(full[name][i])
        self.tr.master.copy_(self.tr.params.buf)

    # ------------------------------------------------------------------------------------------ p2p
    def _p2p(self, send_next=None, send_prev=None, recv_prev=None, recv_next=None) -> Optional[torch.cuda.Event]:
        """One batched NCCL exchange on the side stream (Megatron ``_communicate``); returns the event that marks the
        received tensors valid (and the sent ones reusable)."""
        ops = []
        nxt = self.ranks[self.stage + 1] if not self.last else None
        prv = self.ranks[self.stage - 1] if not self.first else None
        if send_prev is not None:
            ops.append(dist.P2POp(dist.isend, send_prev, prv, group=self.group))
        if recv_prev is not None:
            ops.append(dist.P2POp(dist.irecv, recv_prev, prv, group=self.group))
        if send_next is not None:
            ops.append(dist.P2POp(dist.isend, send_next, nxt, group=self.group))
        if recv_next is not None:
            ops.append(dist.P2POp(dist.irecv, recv_next, nxt, group=self.group))
        if not ops:
            return None
        if not self.on_gpu:                                # CPU / gloo (schedule tests): complete the exchange in place
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            return None
        main = torch.cuda.current_stream(self.device)
        self.comm_stream.wait_stream(main)                 # what we send has been produced
        with torch.cuda.stream(self.comm_stream):
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
        return ev

    def _wait(self, ev):
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    # ------------------------------------------------------------------------------------------ schedule
    def run_microbatches(self, device_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        tr, M = self.tr, len(device_batches)
        warm = min(self.pp - self.stage - 1, M)
        scale = 1.0 / (M * self.T)
        self.loss_acc.zero_()
        slots = self.pp - self.stage                   # activation sets in flight == receive buffers

        def fwd(k, x_in):
            tr.use_slot(k)
            tok, tgt = device_batches[k]
            out = tr.forward(tok if self.first else x_in, tgt, loss_scale=scale)
            if self.last:
                self.loss_acc += out.sum() * scale
                return None
            return out

        def bwd(k, d_out):
            tr.use_slot(k)
            d_in = tr.backward(device_batches[k][0], last_microbatch=(k == M - 1), d_out=d_out)
            if d_in is not None:
                self.send_dx.copy_(d_in)               # tr.dx is overwritten by the next backward while the send is in flight
            return None if self.first else self.send_dx

        f = b = 0                                      # next micro-batch to run forward / backward
        ev_x = self._p2p(recv_prev=self.recv_x[0]) if not self.first and M > 0 else None
        for _ in range(warm):                          # ---- warm-up: forwards only
            self._wait(ev_x)
            y = fwd(f, None if self.first else self.recv_x[f % slots])
            f += 1
            ev_x = self._p2p(send_next=y, recv_prev=None if self.first or f >= M else self.recv_x[f % slots])
        ev_dx = None
        for i in range(M - warm):                      # ---- steady state: one forward, one backward
            self._wait(ev_x)
            y = fwd(f, None if self.first else self.recv_x[f % slots])
            f += 1
            ev_dx = self._p2p(send_next=y, recv_next=None if self.last else self.recv_dx)
            self._wait(ev_dx)
            g = bwd(b, self.recv_dx)
            b += 1
            more_fwd = f < M
            ev_x = self._p2p(send_prev=g, recv_prev=None if self.first or not more_fwd else self.recv_x[f % slots])
            if g is not None and not more_fwd:
                self._wait(ev_x)                       # nothing else orders the reuse of send_dx in the cool-down
        for _ in range(warm):                          # ---- cool-down: backwards only
            ev_dx = self._p2p(recv_next=self.recv_dx)
            self._wait(ev_dx)
            g = bwd(b, self.recv_dx)
            b += 1
            self._wait(self._p2p(send_prev=g))
        if self.on_gpu:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        tr.optimizer_step()
        dist.broadcast(self.loss_acc, src=self.ranks[-1], group=self.group)
        return self.loss_acc

    def train_step(self, host_batches: List[Tuple[torch.Tensor, torch.Tensor]]) -> float:
        """Pinned-host token / target tensors in (every stage gets the same batches), python float loss out."""
        dev = [(t.view(-1).to(self.device, non_blocking=True), y.view(-1).to(self.device, non_blocking=True)) for t, y in host_batches]
        return float(self.run_microbatches(dev).item())

    def memory_bytes(self) -> dict:
        return self.tr.memory_bytes()

    def symmetric_workspaces(self):
        f = getattr(self.tr, "symmetric_workspaces", None)
        return f() if f else []

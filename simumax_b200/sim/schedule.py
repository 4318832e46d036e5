"""Per-rank pipeline programs for the simulator: 1F1B and Megatron interleaved (VPP), sync or async p2p.

The op order follows Megatron-LM's schedules (as modelled by reference transformer/pipeline_schedule.py):
warm-up forwards, steady 1F1B with the combined send/recv (even/odd rank ordering in sync mode, batched
isend/irecv in async mode), cool-down backwards, then the optimizer tail
(grad reduce-scatter -> world all-reduce -> optimizer step -> param all-gather).

``interleaved_order`` is also consumed by the analytic sync-VPP evaluator (core/schedule_vpp.py) so both paths
share one source of truth.
"""
from __future__ import annotations

from typing import List, Tuple

from ..core.utils import get_pp_p2p_comm_size
from .engine import AsyncPost, AsyncWait, Atom, BlockingP2P, Coll, Scope
from .jobs import RankArgs, build_chunk_jobs


def _p2p_cost(strategy, system, hidden_size) -> float:
    size = get_pp_p2p_comm_size(strategy, hidden_size, {"fp32": 4, "fp16": 2, "bf16": 2}[strategy.dtype])
    return system.compute_net_op_time("p2p", size, 2, net=strategy.pp_net)


class _P2P:
    """Factory of p2p jobs for one rank (ids are the rendezvous keys shared by both peers)."""

    def __init__(self, rank_global, pp_rank, pp_size, pp_group, cost, use_async):
        self.g, self.r, self.n, self.grp, self.cost, self.use_async = rank_global, pp_rank, pp_size, pp_group, cost, use_async

    def _gid(self, src, dst, tag):
        return f"send_recv-{src % self.n}-{dst % self.n}-{tag}-pp_group:{self.grp}-"

    # blocking
    def send_next(self, tag):
        return BlockingP2P(f"rank{self.g}-send_next", self._gid(self.r, self.r + 1, tag), self.cost)

    def recv_prev(self, tag):
        return BlockingP2P(f"rank{self.g}-recv_prev", self._gid(self.r - 1, self.r, tag), self.cost)

    def send_prev(self, tag):
        return BlockingP2P(f"rank{self.g}-send_prev", self._gid(self.r, self.r - 1, tag), self.cost)

    def recv_next(self, tag):
        return BlockingP2P(f"rank{self.g}-recv_next", self._gid(self.r + 1, self.r, tag), self.cost)

    # async
    def a_send_next(self, tag):
        return AsyncPost(f"rank{self.g}-async_send_next", self._gid(self.r, self.r + 1, tag), self.cost, "pp_fwd", "send")

    def a_recv_prev(self, tag):
        return AsyncPost(f"rank{self.g}-async_recv_prev", self._gid(self.r - 1, self.r, tag), self.cost, "pp_fwd", "recv")

    def a_wait_prev(self, tag):
        return AsyncWait(f"rank{self.g}-async_wait_recv_prev", self._gid(self.r - 1, self.r, tag), self.cost, "pp_fwd")

    def a_send_prev(self, tag):
        return AsyncPost(f"rank{self.g}-async_send_prev", self._gid(self.r, self.r - 1, tag), self.cost, "pp_bwd", "send")

    def a_recv_next(self, tag):
        return AsyncPost(f"rank{self.g}-async_recv_next", self._gid(self.r + 1, self.r, tag), self.cost, "pp_bwd", "recv")

    def a_wait_next(self, tag):
        return AsyncWait(f"rank{self.g}-async_wait_recv_next", self._gid(self.r + 1, self.r, tag), self.cost, "pp_bwd")


def _wrap(ops, name="", batch=False):
    ops = [o for o in ops if o is not None]
    return ("fwd", Scope(name, ops, batch_blocking=batch)) if ops else None


def build_1f1b_program(perf, chunk, args: RankArgs, with_memory: bool) -> list:
    """Job list [(direction, job)] of one rank for non-interleaved 1F1B."""
    st, system = perf.strategy, perf.system
    pp, rank, mbc = st.pp_size, args.info["pp_rank"], st.micro_batch_num
    use_async = bool(getattr(st, "pp_comm_async", True))
    p2p = _P2P(args.rank, rank, pp, args.info["pp_group_id"], _p2p_cost(st, system, perf.model_config.hidden_size)
               if pp > 1 else 0.0, use_async)
    first, last = rank == 0, rank == pp - 1
    jobs: list = []
    pending_bwd: list = []

    def add(item):
        if item is not None:
            jobs.append(item)

    def fwd_compute():
        f, b = build_chunk_jobs(chunk, args, system, with_memory=with_memory)
        args.microbatch += 1
        jobs.append(("fwd", f))
        pending_bwd.append(b)

    def bwd_compute():
        jobs.append(("bwd", pending_bwd.pop(0)))

    warmup = min(pp - rank - 1, mbc)
    remaining = mbc - warmup
    fi = bi = 0
    args.microbatch = 0
    for i in range(warmup):
        if not first:
            add(_wrap([p2p.a_wait_prev(f"forward-{fi}") if use_async else p2p.recv_prev(f"forward-{fi}")]))
        fwd_compute()
        if not last:
            add(_wrap([p2p.a_send_next(f"forward-{fi}") if use_async else p2p.send_next(f"forward-{fi}")]))
        if use_async and i == warmup - 1 and remaining > 0 and not last:
            add(_wrap([p2p.a_recv_next(f"backward-{bi}")]))
        fi += 1
    for i in range(remaining):
        last_it = i == remaining - 1
        if not first and (use_async or i == 0):
            add(_wrap([p2p.a_wait_prev(f"forward-{fi}") if use_async else p2p.recv_prev(f"forward-{fi}")]))
        fwd_compute()
        if not last:
            if use_async:
                add(_wrap([p2p.a_send_next(f"forward-{fi}")]))
                if not last_it:
                    add(_wrap([p2p.a_recv_next(f"backward-{bi + 1}")]))
            else:
                pair = [p2p.send_next(f"forward-{fi}"), p2p.recv_next(f"backward-{bi}")]
                add(_wrap(pair if rank % 2 else pair[::-1]))
        fi += 1
        if not last and use_async:
            add(_wrap([p2p.a_wait_next(f"backward-{bi}")]))
        bwd_compute()
        if last_it:
            if not first:
                add(_wrap([p2p.a_send_prev(f"backward-{bi}") if use_async else p2p.send_prev(f"backward-{bi}")]))
        elif not first:
            if use_async:
                add(_wrap([p2p.a_send_prev(f"backward-{bi}")]))
                add(_wrap([p2p.a_recv_prev(f"forward-{fi}")]))
            else:
                pair = [p2p.send_prev(f"backward-{bi}"), p2p.recv_prev(f"forward-{fi}")]
                add(_wrap(pair if rank % 2 else pair[::-1]))
        bi += 1
    for _ in range(warmup):
        if not last:
            add(_wrap([p2p.a_wait_next(f"backward-{bi}") if use_async else p2p.recv_next(f"backward-{bi}")]))
        bwd_compute()
        if not first:
            add(_wrap([p2p.a_send_prev(f"backward-{bi}") if use_async else p2p.send_prev(f"backward-{bi}")]))
        bi += 1
    return jobs


# ------------------------------------------------------------------------------------------------
# interleaved (virtual pipeline) order
# ------------------------------------------------------------------------------------------------
class InterleavedPlan:
    """Pure schedule arithmetic of Megatron's interleaved 1F1B for one pipeline rank."""

    def __init__(self, pp_size: int, vp_size: int, pp_rank: int, micro_batch_num: int, group_size: int = None):
        self.pp, self.vp, self.rank, self.mbc = pp_size, vp_size, pp_rank, micro_batch_num
        self.group = group_size or pp_size
        self.total_stages = vp_size * pp_size
        self.total = micro_batch_num * vp_size
        warm = (pp_size - pp_rank - 1) * 2 + (vp_size - 1) * self.group
        self.warmup = min(warm, self.total)
        self.remaining = self.total - self.warmup
        self.table: List[Tuple[int, int]] = []
        for lo in range(0, micro_batch_num, self.group):
            hi = min(micro_batch_num, lo + self.group)
            for chunk in range(vp_size):
                for mb in range(lo, hi):
                    self.table.append((mb, chunk))

    def fwd_ref(self, k):
        mb, chunk = self.table[k]
        return mb, chunk, chunk * self.pp + self.rank, mb * self.vp + chunk

    def bwd_ref(self, k):
        mb, fchunk = self.table[k]
        chunk = self.vp - 1 - fchunk
        return mb, chunk, chunk * self.pp + self.rank, mb * self.vp + chunk

    def _chunk_id(self, k, forward):
        chunk = self.table[k % self.total][1]
        return chunk if forward else self.vp - chunk - 1

    def need_recv(self, k, forward) -> bool:
        """Does finishing virtual step k require receiving the input of step k+1 from the neighbour stage?"""
        leading = (self.rank == 0) if forward else (self.rank == self.pp - 1)
        last_chunk = (self.vp - 1) if forward else 0
        if not leading:
            return True
        recv = True
        if k < self.pp - 1:
            recv = False
            nxt = self._chunk_id(k + 1, forward)
        else:
            nxt = self._chunk_id(k - (self.pp - 1), forward)
        if nxt == last_chunk:
            recv = False
        return recv


def interleaved_sync_ops(plan: "InterleavedPlan") -> list:
    """Abstract op order of one rank for *blocking* interleaved 1F1B.

    Items: ("F", chunk, mb) | ("B", chunk, mb) | ("comm", [(role, tag), ...], batched) where role is one of
    send_prev / recv_prev / send_next / recv_next in Megatron's batch_isend_irecv submission order and ``tag``
    identifies the tensor (direction, destination virtual stage, micro-batch) so both peers derive the same key.
    Shared by the simulator program builder and the analytic sync-VPP evaluator.
    """
    pp, rank, total, last_stage = plan.pp, plan.rank, plan.total, plan.total_stages - 1
    ftag = lambda v, mb: f"forward-v{v}-mb{mb}"
    btag = lambda v, mb: f"backward-v{v}-mb{mb}"
    ops: list = []

    def batch(send_prev=None, recv_prev=None, send_next=None, recv_next=None):
        sub = [x for x in (send_prev, recv_prev, send_next, recv_next) if x is not None]
        if sub:
            ops.append(("comm", sub, True))

    if rank != 0:
        mb0, _, v0, _ = plan.fwd_ref(0)
        if v0 > 0:
            ops.append(("comm", [("recv_prev", ftag(v0, mb0))], False))
    for k in range(plan.warmup):
        mb, chunk, v, _ = plan.fwd_ref(k)
        ops.append(("F", chunk, mb))
        need_next = plan.need_recv(k, True) and k != total - 1
        sn = ("send_next", ftag(v + 1, mb)) if v < last_stage else None
        rp = rn = None
        if k + 1 < total and need_next:
            nmb, _, nv, _ = plan.fwd_ref(k + 1)
            if nv > 0:
                rp = ("recv_prev", ftag(nv, nmb))
        elif plan.remaining == 0 and rank == 0 and k + 1 < total:
            nmb, _, nv, _ = plan.fwd_ref(k + 1)
            if nv > 0:
                rp = ("recv_prev", ftag(nv, nmb))
        if k == plan.warmup - 1 and plan.remaining > 0 and rank != pp - 1:
            bmb, _, bv, _ = plan.bwd_ref(0)
            if bv < last_stage:
                rn = ("recv_next", btag(bv, bmb))
        batch(recv_prev=rp, send_next=sn, recv_next=rn)
    if plan.remaining == 0 and rank != pp - 1:
        bmb, _, bv, _ = plan.bwd_ref(0)
        if bv < last_stage:
            ops.append(("comm", [("recv_next", btag(bv, bmb))], False))
    for k in range(plan.remaining):
        fk, bk = k + plan.warmup, k
        fmb, fchunk, fv, _ = plan.fwd_ref(fk)
        ops.append(("F", fchunk, fmb))
        bmb, bchunk, bv, _ = plan.bwd_ref(bk)
        ops.append(("B", bchunk, bmb))
        need_f = plan.need_recv(fk, True) and k != plan.remaining - 1
        need_b = plan.need_recv(bk, False)
        sn = ("send_next", ftag(fv + 1, fmb)) if fv < last_stage else None
        sp = ("send_prev", btag(bv - 1, bmb)) if bv > 0 else None
        rp = rn = None
        if fk + 1 < total and need_f:
            nmb, _, nv, _ = plan.fwd_ref(fk + 1)
            if nv > 0:
                rp = ("recv_prev", ftag(nv, nmb))
        if bk + 1 < total and need_b:
            nmb, _, nv, _ = plan.bwd_ref(bk + 1)
            if nv < last_stage:
                rn = ("recv_next", btag(nv, nmb))
        batch(send_prev=sp, recv_prev=rp, send_next=sn, recv_next=rn)
    for k in range(plan.remaining, total):
        bmb, bchunk, bv, _ = plan.bwd_ref(k)
        ops.append(("B", bchunk, bmb))
        need_b = plan.need_recv(k, False) and k != total - 1
        sp = ("send_prev", btag(bv - 1, bmb)) if bv > 0 else None
        rn = None
        if k + 1 < total and need_b:
            nmb, _, nv, _ = plan.bwd_ref(k + 1)
            if nv < last_stage:
                rn = ("recv_next", btag(nv, nmb))
        elif plan.remaining == 0 and rank == pp - 1 and k + 1 < total:
            nmb, _, nv, _ = plan.bwd_ref(k + 1)
            if nv < last_stage:
                rn = ("recv_next", btag(nv, nmb))
        batch(send_prev=sp, recv_next=rn)
    return ops


def build_interleaved_program(perf, stage_chunks: list, args: RankArgs, with_memory: bool) -> list:
    """Megatron interleaved schedule (sync: batched blocking p2p groups; async: isend/irecv bundles + waits)."""
    st, system = perf.strategy, perf.system
    pp, rank = st.pp_size, args.info["pp_rank"]
    if pp <= 1:
        raise NotImplementedError("Strict interleaved simu schedule requires pp_size > 1.")
    vp = len(stage_chunks)
    plan = InterleavedPlan(pp, vp, rank, st.micro_batch_num, getattr(st, "microbatch_group_size_per_vp_stage", None))
    use_async = bool(getattr(st, "pp_comm_async", True))
    if use_async and st.micro_batch_num < pp * vp:
        raise RuntimeError("async VPP does not support micro_batch_num < pp_size * vp_size yet")
    p2p = _P2P(args.rank, rank, pp, args.info["pp_group_id"], _p2p_cost(st, system, perf.model_config.hidden_size), use_async)
    total, last_stage = plan.total, plan.total_stages - 1
    jobs: list = []

    def model(chunk_idx, real_mb, direction):
        # the backward tree is built when the schedule reaches it (the reference instantiates the chunk a second time there,
        # pipeline_schedule.py:303-304), so its collectives draw fresh comm_order numbers -> identical gids in the log / trace
        args.microbatch, args.chunk_idx = real_mb, chunk_idx
        f, b = build_chunk_jobs(stage_chunks[chunk_idx], args, system, chunk_prefix=f"-chunk{chunk_idx}-",
                                with_memory=with_memory)
        jobs.append(("fwd", f) if direction == "fwd" else ("bwd", b))

    ftag = lambda v, mb: f"forward-v{v}-mb{mb}"
    btag = lambda v, mb: f"backward-v{v}-mb{mb}"

    def add(item):
        if item is not None:
            jobs.append(item)

    if not use_async:
        mk = {"send_next": p2p.send_next, "recv_prev": p2p.recv_prev, "send_prev": p2p.send_prev,
              "recv_next": p2p.recv_next}
        for op in interleaved_sync_ops(plan):
            if op[0] == "F":
                model(op[1], op[2], "fwd")
            elif op[0] == "B":
                model(op[1], op[2], "bwd")
            else:
                _, subops, batched = op
                built = [mk[role](tag) for role, tag in subops]
                add(_wrap(built, name=f"rank{args.rank}-batch_pp_comm" if batched else "", batch=batched))
        return jobs

    # ---- async: bundles of isend/irecv posts, waits right before the consuming compute
    posted_f, posted_b = set(), set()

    def bundle(send_next=None, send_prev=None, recv_prev=None, recv_next=None):
        rp = rn = None
        if recv_prev is not None and ("f",) + recv_prev not in posted_f:
            posted_f.add(("f",) + recv_prev)
            rp = p2p.a_recv_prev(ftag(recv_prev[1], recv_prev[0]))
        sn = p2p.a_send_next(ftag(send_next[1] + 1, send_next[0])) if send_next is not None else None
        if recv_next is not None and ("b",) + recv_next not in posted_b:
            posted_b.add(("b",) + recv_next)
            rn = p2p.a_recv_next(btag(recv_next[1], recv_next[0]))
        sp = p2p.a_send_prev(btag(send_prev[1] - 1, send_prev[0])) if send_prev is not None else None
        ordered = [sn, rp, sp, rn] if rank % 2 == 0 else [rp, sn, rn, sp]
        add(_wrap(ordered))

    def fwd_compute(k, need_recv_prev):
        mb, chunk, v, _ = plan.fwd_ref(k)
        if v > 0 and need_recv_prev:
            add(_wrap([p2p.a_wait_prev(ftag(v, mb))]))
        model(chunk, mb, "fwd")

    def bwd_compute(k, need_recv_next):
        mb, chunk, v, _ = plan.bwd_ref(k)
        if v < last_stage and need_recv_next:
            add(_wrap([p2p.a_wait_next(btag(v, mb))]))
        model(chunk, mb, "bwd")

    if rank != 0:
        mb0, _, v0, _ = plan.fwd_ref(0)
        if v0 > 0:
            add(_wrap([p2p.a_wait_prev(ftag(v0, mb0))]))
    need_f, need_b = rank != 0, False
    for k in range(plan.warmup):
        mb, _, v, _ = plan.fwd_ref(k)
        fwd_compute(k, need_f)
        need_f_next = plan.need_recv(k, True) and k != total - 1
        rn_spec = None
        if k == plan.warmup - 1 and plan.remaining > 0:
            need_b = rank != pp - 1
            if need_b:
                bmb, _, bv, _ = plan.bwd_ref(0)
                if bv < last_stage:
                    rn_spec = (bmb, bv)
        rp_spec = None
        if k + 1 < total and need_f_next:
            nmb, _, nv, _ = plan.fwd_ref(k + 1)
            if nv > 0:
                rp_spec = (nmb, nv)
        bundle(send_next=(mb, v) if v < last_stage else None, recv_prev=rp_spec, recv_next=rn_spec)
        need_f = need_f_next
    for k in range(plan.remaining):
        fk, bk = k + plan.warmup, k
        fmb, _, fv, _ = plan.fwd_ref(fk)
        bmb, _, bv, _ = plan.bwd_ref(bk)
        fwd_compute(fk, need_f)
        bwd_compute(bk, need_b)
        need_f_next = plan.need_recv(fk, True) and k != plan.remaining - 1
        need_b_next = plan.need_recv(bk, False)
        rp_spec = rn_spec = None
        if fk + 1 < total and need_f_next:
            nmb, _, nv, _ = plan.fwd_ref(fk + 1)
            if nv > 0:
                rp_spec = (nmb, nv)
        if bk + 1 < total and need_b_next:
            nmb, _, nv, _ = plan.bwd_ref(bk + 1)
            if nv < last_stage:
                rn_spec = (nmb, nv)
        bundle(send_next=(fmb, fv) if fv < last_stage else None, send_prev=(bmb, bv) if bv > 0 else None,
               recv_prev=rp_spec, recv_next=rn_spec)
        need_f, need_b = need_f_next, need_b_next
    for k in range(plan.remaining, total):
        bmb, _, bv, _ = plan.bwd_ref(k)
        bwd_compute(k, need_b)
        need_b_next = plan.need_recv(k, False) and k != total - 1
        rn_spec = None
        if k + 1 < total and need_b_next:
            nmb, _, nv, _ = plan.bwd_ref(k + 1)
            if nv < last_stage:
                rn_spec = (nmb, nv)
        bundle(send_prev=(bmb, bv) if bv > 0 else None, recv_next=rn_spec)
        need_b = need_b_next
    return jobs


# ------------------------------------------------------------------------------------------------
# optimizer tail
# ------------------------------------------------------------------------------------------------
def build_optimizer_tail(perf, model_name: str, args: RankArgs) -> tuple:
    st = perf.strategy
    if st.zero_state < 1:
        raise NotImplementedError("the simulator models the distributed optimizer (zero_state >= 1) only")
    comm = perf._compute_dp_time(model_name)
    opt = perf._compute_optim_time(model_name)
    info = args.info
    label = f"rank{args.rank}-microbatch{args.microbatch}-OptimizerSimulator"
    dense, moe = comm["dense"], comm["moe"]
    dp_group = st.dp_size * st.cp_size

    def exposed(fam, key):
        """Cost of one ZeRO-1 collective in the replay: the full time, or -- with the overlap extension (``dp_overlap_model``:
        buckets hidden behind the last backward / next forward) -- the share of it the closed form leaves exposed."""
        t = fam["details"][key]
        total = fam.get("dp_comm_time", 0)
        if getattr(st, "dp_overlap_model", False) and total:
            t *= fam["dp_comm_exposed_time"] / total
        return t

    def coll(kind, tag, group_id, size, cost):
        gid = f"{kind}{args.comm_order}-{tag}:{group_id}"
        args.comm_order += 1
        return Coll(f"{label}-{kind}", gid, size, fwd_cost=cost)

    ops = [
        coll("reduce_scatter", "dp_cp_group", info["dp_cp_group_id"], dp_group, exposed(dense, "reduce_scatter_time")),
        coll("reduce_scatter", "edp_group", info["edp_group_id"], st.edp_size, exposed(moe, "reduce_scatter_time")),
        Coll(f"{label}-all_reduce", f"all_reducedefault_group-pp_size:{st.pp_size}", st.world_size, fwd_cost=1,
             world=True),
        Atom(f"{label}-optimizer_step", fwd_cost=opt["optim_time"]),
        coll("all_gather", "dp_cp_group", info["dp_cp_group_id"], dp_group, exposed(dense, "all_gather_time")),
        coll("all_gather", "edp_group", info["edp_group_id"], st.edp_size, exposed(moe, "all_gather_time")),
    ]
    return ("fwd", Scope(label, ops))

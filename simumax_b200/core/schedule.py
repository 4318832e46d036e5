"""Analytic pipeline schedules (closed-loop event matching; no simulator needed).

``one_f_one_b`` reproduces Megatron's non-interleaved 1F1B: per-rank op queues
(recv / F / send / B) with pairwise send-recv rendezvous, warm-up ``min(pp - rank - 1, mbc)``, and the
even/odd-rank ordering of the combined send+recv in steady state.  A p2p pair completes at
``max(start_a, start_b) + max(dur_a, dur_b)``.  (reference perf_llm.py:2097-2306)

The interleaved (sync-VPP) evaluator lives in ``schedule_vpp.py`` and shares the op-order generator with the
discrete-event simulator so both paths agree by construction.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Sequence, Tuple


class _Op:
    __slots__ = ("kind", "mb", "gid", "dur", "peer", "label", "chunk_idx", "virtual_idx")

    def __init__(self, kind, mb, gid, dur, peer, label, chunk_idx=None, virtual_idx=None):
        self.kind, self.mb, self.gid, self.dur, self.peer, self.label = kind, mb, gid, dur, peer, label
        self.chunk_idx, self.virtual_idx = chunk_idx, virtual_idx


def build_1f1b_queues(pp: int, mbc: int, stage_phases: Sequence[Dict[str, float]]) -> List[deque]:
    """Per-rank operation order of 1F1B with explicit p2p ops."""
    queues: List[deque] = [deque() for _ in range(pp)]
    for rank in range(pp):
        spec = stage_phases[rank]
        q = queues[rank]
        warmup = min(pp - rank - 1, mbc)
        remaining = mbc - warmup
        first, last = rank == 0, rank == pp - 1

        def recv_prev(idx):
            return _Op("recv", idx, ("fwd", idx, rank - 1, rank), spec["fwd_recv"], rank - 1, "recv_prev")

        def send_next(idx):
            return _Op("send", idx, ("fwd", idx, rank, rank + 1), spec["fwd_send"], rank + 1, "send_next")

        def recv_next(idx):
            return _Op("recv", idx, ("bwd", idx, rank + 1, rank), spec["bwd_recv"], rank + 1, "recv_next")

        def send_prev(idx):
            return _Op("send", idx, ("bwd", idx, rank, rank - 1), spec["bwd_send"], rank - 1, "send_prev")

        def combined(sends, recvs):
            q.extend(sends + recvs if rank % 2 else recvs + sends)

        f = b = 0
        for _ in range(warmup):
            if not first and spec["fwd_recv"] > 0:
                q.append(recv_prev(f))
            q.append(_Op("F", f, None, spec["fwd_compute"], None, "fwd_compute"))
            if not last and spec["fwd_send"] > 0:
                q.append(send_next(f))
            f += 1
        for i in range(remaining):
            if not first and i == 0 and spec["fwd_recv"] > 0:
                q.append(recv_prev(f))
            q.append(_Op("F", f, None, spec["fwd_compute"], None, "fwd_compute"))
            if not last:
                combined([send_next(f)] if spec["fwd_send"] > 0 else [],
                         [recv_next(b)] if spec["bwd_recv"] > 0 else [])
            f += 1
            q.append(_Op("B", b, None, spec["bwd_compute"], None, "bwd_compute"))
            if i == remaining - 1:
                if not first and spec["bwd_send"] > 0:
                    q.append(send_prev(b))
            elif not first:
                combined([send_prev(b)] if spec["bwd_send"] > 0 else [],
                         [recv_prev(f)] if spec["fwd_recv"] > 0 else [])
            b += 1
        for _ in range(warmup):
            if not last and spec["bwd_recv"] > 0:
                q.append(recv_next(b))
            q.append(_Op("B", b, None, spec["bwd_compute"], None, "bwd_compute"))
            if not first and spec["bwd_send"] > 0:
                q.append(send_prev(b))
            b += 1
    return queues


def run_rendezvous_queues(queues: List[deque], what: str = "1F1B") -> Tuple[float, List[List[dict]]]:
    """Execute per-rank FIFO queues: compute ops advance the local clock, send/recv pairs rendezvous."""
    n = len(queues)
    clock = [0.0] * n
    records: List[List[dict]] = [[] for _ in range(n)]

    def rec(rank, op, start, end):
        records[rank].append({"kind": op.kind, "mb": op.mb, "start": start, "duration": end - start, "end": end,
                              "label": op.label, "chunk_idx": op.chunk_idx, "virtual_idx": op.virtual_idx,
                              "gid": op.gid})

    while any(queues):
        progressed = False
        for rank in range(n):
            q = queues[rank]
            while q and q[0].kind in ("F", "B"):
                op = q.popleft()
                rec(rank, op, clock[rank], clock[rank] + op.dur)
                clock[rank] += op.dur
                progressed = True
        matched = set()
        for rank in range(n):
            q = queues[rank]
            if rank in matched or not q or q[0].kind not in ("send", "recv"):
                continue
            op = q[0]
            peer = op.peer
            if peer is None or peer in matched or not queues[peer]:
                continue
            other = queues[peer][0]
            if other.kind not in ("send", "recv") or other.gid != op.gid or other.kind == op.kind:
                continue
            end = max(clock[rank], clock[peer]) + max(op.dur, other.dur)
            rec(rank, op, clock[rank], end)
            rec(peer, other, clock[peer], end)
            clock[rank] = clock[peer] = end
            q.popleft()
            queues[peer].popleft()
            matched.update((rank, peer))
            progressed = True
        if not progressed:
            raise RuntimeError(f"{what} phase scheduler deadlock: {[q[0].label if q else None for q in queues]}")
    return (max(clock) if clock else 0.0), records


def one_f_one_b(pp: int, mbc: int, stage_phases: Sequence[Dict[str, float]], return_schedules: bool = False):
    total, records = run_rendezvous_queues(build_1f1b_queues(pp, mbc, stage_phases), "1F1B")
    if return_schedules:
        return total, records
    return total


def one_f_one_b_simple(pp: int, mbc: int, forward_times: Sequence[float], backward_times: Sequence[float],
                       return_schedules: bool = False):
    """Dependency-only 1F1B (no explicit p2p ops): used when only per-stage fwd/bwd totals are known."""
    sched: List[List[dict]] = [[] for _ in range(pp)]
    fwd_ready = [[0] for _ in range(pp)]
    bwd_ready = [[0] for _ in range(pp)]

    def push(rank, kind, ready_list, dep, dur):
        now = sched[rank][-1]["end"] if sched[rank] else 0
        start = max(now, dep)
        sched[rank].append({"kind": kind, "mb": len(ready_list[rank]), "start": start, "duration": dur,
                            "end": start + dur, "label": "fwd_compute" if kind == "F" else "bwd_compute"})
        ready_list[rank].append(start + dur)

    for step in range(mbc):
        for rank in range(pp):
            push(rank, "F", fwd_ready, fwd_ready[rank - 1][-1] if rank > 0 else 0, forward_times[rank])
            if step >= pp - 1 - rank:
                push(rank, "B", bwd_ready, bwd_ready[rank + 1][-1] if rank < pp - 1 else 0, backward_times[rank])
    for step in range(pp - 1, -1, -1):
        for rank in range(step):
            push(rank, "B", bwd_ready, bwd_ready[rank + 1][-1] if rank < pp - 1 else 0, backward_times[rank])
    total = max(s[-1]["end"] for s in sched)
    return (total, sched) if return_schedules else total


# ------------------------------------------------------------------------------------------------
# straggler model (reference perf_llm.py:38,255-291)
# ------------------------------------------------------------------------------------------------
STRAGGLER_BASE_FACTOR = 0.09


def effective_straggler_samples(world_size: int, num_per_node: int, dp_size: int, edp_size: int) -> int:
    import math
    nodes = max(1, math.ceil(int(world_size) / max(1, int(num_per_node))))
    return max(1, min(nodes, int(dp_size), int(edp_size)))


def straggler_increase_ratio(worker_count: int) -> float:
    """1 + L/(L+1) * 0.09 * sqrt(L), L = log2(n): expected max-of-n slowdown, damped for small n."""
    import math
    n = max(1, int(worker_count))
    if n <= 1:
        return 1.0
    L = math.log2(n)
    return 1.0 + L / (L + 1.0) * STRAGGLER_BASE_FACTOR * math.sqrt(L)

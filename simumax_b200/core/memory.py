"""Activation-lifetime walker: peak live memory of one micro-batch through a stage chunk.

Forward pass: ``peak = max(peak, live + fwd_peak(op))`` then ``live += cache(op)``.  Inside a recompute
segment only the FIRST op keeps its *inputs* during the real forward; everything else is materialised when
the segment is replayed in backward (stage ``recompute_forward``) and released by its backward
(``recompute_backward``).  The walk must end with zero live bytes -- that invariant is asserted.

Behavioural spec: reference transformer/language_model.py:317-467.  Implemented over the flat IR list.
"""
from __future__ import annotations

from typing import List

from .ir import OpNode
from .records import PeakPoint, RecomputeStatus


def assign_recompute_status(nodes: List[OpNode]) -> None:
    """FIRST / MIDDLE / LAST marks of maximal runs of recompute-enabled ops (language_model.py:289-315)."""
    prev_enabled = False
    prev_node = None
    for n in nodes:
        n.recompute_status = RecomputeStatus.NO_RECOMPUTE
        n.recompute_done = False
    for n in nodes:
        if n.enable_recompute:
            n.recompute_status = RecomputeStatus.MIDDLE
        if not prev_enabled and n.enable_recompute:
            n.recompute_status = RecomputeStatus.FIRST
        if prev_enabled and not n.enable_recompute and prev_node is not None:
            prev_node.recompute_status = RecomputeStatus.LAST
        if n.enable_recompute:
            prev_node = n
        prev_enabled = n.enable_recompute


def apply_breakpoints(nodes: List[OpNode], layer_first_nodes: List[OpNode]) -> None:
    """Checkpoint boundaries: explicit breakpoint ops close a segment; the first op of every layer opens one."""
    for cur, nxt in zip(nodes[:-1], nodes[1:]):
        if cur.is_breakpoint and cur.enable_recompute:
            cur.recompute_status = RecomputeStatus.LAST
            if nxt.enable_recompute:
                nxt.recompute_status = RecomputeStatus.FIRST
    for first in layer_first_nodes:
        if first.enable_recompute:
            first.is_breakpoint = True
            first.recompute_status = RecomputeStatus.FIRST


def _label(n: OpNode) -> str:
    return f"{n.name}: {n.path}"


def _forward(nodes: List[OpNode], recompute_on: bool, live: float, peak: PeakPoint, stage: str) -> float:
    assert stage in ("forward", "recompute_forward")
    last = None
    for n in nodes:
        last = n
        peak.update_peak(_label(n), live + n.act.fwd_peak_mem_no_cache, stage)
        if recompute_on and n.enable_recompute:
            if stage == "recompute_forward" and n.recompute_status != RecomputeStatus.FIRST:
                n.act.cache_for_bwd_mem = n.act.total_activation_mem_cache
                live += n.act.cache_for_bwd_mem
            elif stage == "forward" and n.recompute_status == RecomputeStatus.FIRST:
                n.act.cache_for_bwd_mem = 0 if n.offload_inputs else n.in_bytes
                live += n.act.cache_for_bwd_mem
        else:
            n.act.cache_for_bwd_mem = n.act.total_activation_mem_cache
            live += n.act.cache_for_bwd_mem
    if last is not None:
        peak.update_peak(_label(last), live, stage)
    if stage == "forward":
        peak.set_forward_mem_cache(live)
    assert peak.peak_mem >= live
    return live


def _backward_only(nodes: List[OpNode], live: float, peak: PeakPoint, stage: str) -> float:
    for n in reversed(nodes):
        peak.update_peak(_label(n), live + n.act.bwd_peak_mem_no_cache, stage)
        live -= n.act.cache_for_bwd_mem
        n.act.cache_for_bwd_mem = 0
    return live


def _backward(nodes: List[OpNode], recompute_on: bool, live: float, peak: PeakPoint) -> float:
    pending: List[OpNode] = []
    segment_ready = False
    i = len(nodes) - 1
    while i >= 0:
        n = nodes[i]
        if recompute_on and n.enable_recompute and not n.recompute_done and not segment_ready:
            pending.append(n)
            if n.recompute_status == RecomputeStatus.FIRST:
                segment_ready = True
            i -= 1
        elif pending:
            seg = pending[::-1]
            live = _forward(seg, recompute_on, live, peak, "recompute_forward")
            live = _backward_only(seg, live, peak, "recompute_backward")
            for m in seg:
                m.recompute_done = True
            pending = []
            segment_ready = False
        else:
            peak.update_peak(_label(n), live + n.act.bwd_peak_mem_no_cache, "backward")
            live -= n.act.cache_for_bwd_mem
            n.act.cache_for_bwd_mem = 0
            i -= 1
    if pending:
        seg = pending[::-1]
        live = _forward(seg, recompute_on, live, peak, "recompute_forward")
        live = _backward_only(seg, live, peak, "recompute_backward")
    assert peak.peak_mem >= live
    return live


def walk_activations(nodes: List[OpNode], layer_first_nodes: List[OpNode], recompute_on: bool) -> PeakPoint:
    for n in nodes:
        n.recompute_done = False
        n.act.cache_for_bwd_mem = 0
    apply_breakpoints(nodes, layer_first_nodes)
    peak = PeakPoint()
    live = _forward(nodes, recompute_on, 0, peak, "forward")
    live = _backward(nodes, recompute_on, live, peak)
    for n in nodes:
        assert n.act.cache_for_bwd_mem == 0, f"{n.name} still holds {n.act.cache_for_bwd_mem / 2 ** 20:.2f} MB"
    assert live == 0, f"activation walk must end at 0 live bytes, got {live / 2 ** 20:.2f} MB"
    return peak

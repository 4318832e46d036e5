"""Turn an evaluated ``StageChunk`` (flat op IR + scope paths) into the simulator's job trees.

One forward tree and one backward tree per (rank, micro-batch, chunk).  Nesting mirrors the module nesting of
the model (every scope costs 2 us when it closes), recomputed layers use the flat checkpoint-segment layout
(replay forward of the segment, then its backward) described in reference base_struct.py:435-493.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from ..core.records import RecomputeStatus
from ..core.ir import OpNode, SimStep
from ..core.utils import get_rank_group
from .engine import Atom, Coll, RecomputeBlock, Scope
from .memory import OpMemoryProfile


class RankArgs:
    """Per-rank build state (reference ``args`` namespace + thread_state)."""

    def __init__(self, rank: int, strategy):
        self.rank = rank
        self.microbatch = 0
        self.chunk_idx = None
        self.comm_order = 0
        self.info = get_rank_group(rank, strategy)


def _mb_tag(args: RankArgs) -> str:
    tag = f"microbatch:{args.microbatch}"
    if args.chunk_idx is not None:
        tag += f"-chunk:{args.chunk_idx}"
    return tag


_GROUP_KEY = {"tp": ("tp_group", "tp_group_id", "tp_rank"), "cp": ("cp_group", "cp_group_id", "cp_rank"),
              "ep": ("ep_group", "ep_group_id", "ep_rank"), "etp": ("tp_group", "tp_group_id", "tp_rank")}


def mem_profile_for(node: OpNode, scope_label: str, strategy, is_segment_first: bool) -> OpMemoryProfile:
    """Cache / peak bytes of one op for the memory timeline (reference base_struct.py:377-413)."""
    act = node.act
    cache, alloc_phase, rc_peak = 0, None, 0
    if strategy.enable_recompute and node.enable_recompute:
        rc_peak = act.fwd_peak_mem_no_cache
        if is_segment_first:
            if not node.offload_inputs:
                cache, alloc_phase = node.in_bytes, "fwd"
        else:
            cache, alloc_phase = act.total_activation_mem_cache, "recompute_fwd"
    else:
        cache, alloc_phase = act.total_activation_mem_cache, "fwd"
    if node.is_variance_node and alloc_phase == "recompute_fwd":
        cache, alloc_phase = 0, None
    return OpMemoryProfile(op_name=node.name or scope_label, fwd_peak_mem_no_cache=int(act.fwd_peak_mem_no_cache),
                           bwd_peak_mem_no_cache=int(act.bwd_peak_mem_no_cache),
                           recompute_peak_mem_no_cache=int(rc_peak), cache_size_bytes=int(cache),
                           cache_alloc_phase=alloc_phase, cache_token_scope=scope_label)


def _steps_to_jobs(node: OpNode, steps: List[SimStep], label: str, args: RankArgs, strategy, recompute_cost=None):
    jobs = []
    for s in steps:
        if s.kind == "atom":
            jobs.append(Atom(name=f"{label}-{s.name or 'AtomModel'}", fwd_cost=s.fwd_cost, bwd_cost=s.bwd_cost,
                             recompute_cost=recompute_cost))
        else:
            gname, gid_key, _ = _GROUP_KEY.get(s.group, _GROUP_KEY["tp"])
            layer = f"-layer:{node.layer_idx}" if node.layer_idx is not None and node.kind in (
                "LinearCol", "LinearRow", "Router", "Permutation", "UnPermutation") else ""
            gid = (f"{s.comm_op}{args.comm_order}-{_mb_tag(args)}{layer}-name:{node.kind}-{gname}:"
                   f"{args.info[gid_key]}" + (f"-stage:{s.stage}" if s.group == "cp" else ""))
            args.comm_order += 1
            jobs.append(Coll(name=f"{label}-{s.name or s.comm_op}", gid=gid, group_size=s.comm_num, fwd_cost=s.fwd_cost,
                             bwd_cost=s.bwd_cost))
    return jobs


def build_chunk_jobs(chunk, args: RankArgs, system, chunk_prefix: str = "", with_memory: bool = True
                     ) -> Tuple[Scope, Scope]:
    """Return (forward tree, backward tree) of one micro-batch through ``chunk`` for the rank in ``args``."""
    strategy = chunk.strategy
    root_label = f"rank{args.rank}-microbatch{args.microbatch}{chunk_prefix}-{chunk.specific_name}"

    # leaf scopes ------------------------------------------------------------------------------------
    leaf_fwd: List[Scope] = []
    leaf_bwd: List[Scope] = []
    leaf_rc: List[Optional[Scope]] = []
    labels: List[str] = []
    # which op keeps its *inputs* through the real forward: the one the activation walker marked FIRST (after breakpoints).
    # A one-op segment (sdp_only) ends up LAST, not FIRST, in the reference (language_model.py:289-315) and therefore caches
    # nothing in the forward pass and everything in the replay -- same here.
    seg_first = [node.enable_recompute and node.recompute_status == RecomputeStatus.FIRST for node in chunk.nodes]
    for k, node in enumerate(chunk.nodes):
        label = "-".join([root_label] + list(node.scope) + [node.display or node.kind])
        labels.append(label)
        steps = node.sim_steps(node, system) if node.sim_steps is not None else []
        prof = mem_profile_for(node, label, strategy, seg_first[k]) if with_memory else None
        fwd_jobs = _steps_to_jobs(node, steps, label, args, strategy)
        leaf_fwd.append(Scope(label, fwd_jobs, mem_profile=prof))
        leaf_bwd.append(Scope(label, fwd_jobs, mem_profile=prof))   # same children, executed reversed
        if node.enable_recompute and not node.is_variance_node:
            rc_children = [Atom(j.name, j.fwd_cost, j.bwd_cost, recompute_cost=node.cost.recompute_compute_time)
                           if isinstance(j, Atom) else j for j in fwd_jobs]
            leaf_rc.append(Scope(label, rc_children, mem_profile=prof))
        else:
            leaf_rc.append(None)

    # nesting from scope paths -------------------------------------------------------------------------
    def nest(lo: int, hi: int, depth: int, prefix_label: str, leaves: List[Scope]) -> List:
        out, i = [], lo
        while i < hi:
            sc = chunk.nodes[i].scope
            if len(sc) == depth:
                out.append(leaves[i])
                i += 1
                continue
            name = sc[depth]
            j = i
            while j < hi and len(chunk.nodes[j].scope) > depth and chunk.nodes[j].scope[depth] == name \
                    and chunk.nodes[j].scope[:depth] == sc[:depth]:
                j += 1
            label = f"{prefix_label}-{name}"
            out.append(Scope(label, nest(i, j, depth + 1, label, leaves)))
            i = j
        return out

    fwd_root = Scope(root_label, nest(0, len(chunk.nodes), 0, root_label, leaf_fwd))

    # backward: layers that recompute use flat checkpoint segments ---------------------------------------
    def layer_range(first_idx: int) -> Tuple[int, int]:
        layer_name = chunk.nodes[first_idx].scope[0]
        j = first_idx
        while j < len(chunk.nodes) and chunk.nodes[j].scope and chunk.nodes[j].scope[0] == layer_name:
            j += 1
        return first_idx, j

    bwd_children = []
    i, n = 0, len(chunk.nodes)
    first_ids = {id(nd) for nd in chunk.layer_first_nodes}
    first_indices = {k for k, nd in enumerate(chunk.nodes) if id(nd) in first_ids}
    while i < n:
        if i in first_indices:
            lo, hi = layer_range(i)
            layer_label = f"{root_label}-{chunk.nodes[i].scope[0]}"
            if strategy.is_recompute and any(nd.enable_recompute for nd in chunk.nodes[lo:hi]) and \
                    _layer_index(chunk.nodes[i]) < strategy.recompute_layer_num:
                bwd_children.append(Scope(layer_label, _checkpoint_segments(chunk.nodes, lo, hi, leaf_bwd, leaf_rc,
                                                                            layer_label)))
            else:
                bwd_children.append(Scope(layer_label, nest(lo, hi, 1, layer_label, leaf_bwd)))
            i = hi
        else:
            bwd_children.append(leaf_bwd[i])
            i += 1
    bwd_root = Scope(root_label, bwd_children)
    return fwd_root, bwd_root


def _layer_index(node: OpNode) -> int:
    name = node.name.split(".")[1] if "." in node.name else ""
    try:
        return int(name.split("_")[1])
    except Exception:
        return 10 ** 9


def _checkpoint_segments(nodes, lo, hi, leaf_bwd, leaf_rc, label) -> list:
    """Maximal runs of recompute-enabled ops -> RecomputeBlock; other ops keep a plain backward scope."""
    out, seg = [], []

    def flush():
        if seg:
            out.append(RecomputeBlock(label, fwd_children=[leaf_rc[k] for k in seg if leaf_rc[k] is not None],
                                      bwd_children=[leaf_bwd[k] for k in seg]))
            seg.clear()

    prev_enabled = False
    for k in range(lo, hi):
        if nodes[k].enable_recompute:
            if seg and not prev_enabled:
                flush()
            seg.append(k)
            prev_enabled = True
        else:
            flush()
            out.append(leaf_bwd[k])
            prev_enabled = False
    flush()
    return out

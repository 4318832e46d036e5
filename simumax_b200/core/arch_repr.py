"""Module-tree rendering of an evaluated stage chunk: the text of ``model_arch`` / ``base_info.json["arch"]``.

The IR is flat; every op carries the reference's tree path (``OpNode.ref_path``: child index + class name per level), which is
enough to rebuild the nested view users of the reference read -- torch-style ``Class(extra)`` blocks, consecutive identical
children folded into ``(a->b)``, and under every leaf *and* every composite its time / FLOP / static-memory totals
(reference base_struct.py:1004-1068 ``MetaModule.__repr__``).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

from .records import CostInfo, MemoryInfo

_SEG = re.compile(r"\((\d+)\)(.+)")


class _Tree:
    __slots__ = ("cls", "children", "node")

    def __init__(self, cls: str):
        self.cls = cls
        self.children: Dict[int, "_Tree"] = {}
        self.node = None                       # OpNode of a leaf


def build_tree(nodes) -> _Tree:
    root: Optional[_Tree] = None
    for nd in nodes:
        segs = [_SEG.fullmatch(s.strip()) for s in nd.ref_path.split("->")]
        assert all(segs), f"op {nd.name} has no module-tree path"
        if root is None:
            root = _Tree(segs[0].group(2))
        cur = root
        for m in segs[1:]:
            idx, cls = int(m.group(1)), m.group(2)
            nxt = cur.children.get(idx)
            if nxt is None:
                nxt = cur.children[idx] = _Tree(cls)
            cur = nxt
        cur.node = nd
    return root if root is not None else _Tree("LLMModel")


def _indent_tail(text: str, n: int) -> str:
    first, *rest = text.split("\n")
    return "\n".join([first] + [" " * n + line for line in rest])


def _totals(tree: _Tree):
    """(cost, static memory, flops, is_variance_node) of a subtree; children are added in child order."""
    if tree.node is not None:
        nd = tree.node
        return nd.cost, nd.mem, sum(nd.compute.get_all_flops()), nd.is_variance_node
    cost, mem, flops = CostInfo(), MemoryInfo(), 0
    for idx in sorted(tree.children):
        c, m, f, _ = _totals(tree.children[idx])
        cost, mem, flops = cost + c, mem + m, flops + f
    return cost, mem, flops, False


def hierarchical_totals(nodes):
    """(static memory, compute, cost) of a chunk, added up the way the module tree does it: every composite sums its children
    in order, then its parent sums the composites (base_struct.py:758-761).  Floating-point addition is not associative; this
    association keeps e.g. ``bubble_time = schedule - mbc * chunk_time`` bit-identical (+0.0000 rather than -0.0000)."""
    from .records import ComputeInfo

    def walk(tree):
        if tree.node is not None:
            return tree.node.mem, tree.node.compute, tree.node.cost
        mem, comp, cost = MemoryInfo(), ComputeInfo(), CostInfo()
        for idx in sorted(tree.children):
            m, c, t = walk(tree.children[idx])
            mem, comp, cost = mem + m, comp + c, cost + t
        return mem, comp, cost
    return walk(build_tree(nodes))


def render(tree: _Tree) -> str:
    lines: List[str] = []
    extra = tree.node.repr_args if tree.node is not None else ""
    prev, first_idx = None, 0
    for idx in sorted(tree.children):
        text = _indent_tail(render(tree.children[idx]), 2)
        if text == prev:                                   # fold runs of identical children: (first->last)
            lines[-1] = f"({first_idx}->{idx}): {text}"
        else:
            lines.append(f"({idx}): {text}")
            first_idx = idx
        prev = text
    head = tree.cls + "("
    if extra and not lines:
        head += extra
    elif extra or lines:
        head += "\n  " + "\n  ".join(([extra] if extra else []) + lines) + "\n"
    head += ")"
    c, m, flops, variance = _totals(tree)
    head += (f"\n\t1. cost: (total_time={c.all_time:.2f} ms, fwd_details=(sum={c.fwd_time + c.fwd_net_time:.2f} ms, "
             f"compute={c.fwd_compute_time * 1000:.2f} us, net={c.fwd_net_time * 1000:.2f} us), "
             f"bwd_details=(sum={c.bwd_time + c.bwd_net_time:.2f} ms, compute={c.bwd_compute_time * 1000:.2f} us, "
             f"net={c.bwd_net_time * 1000:.2f} us), variance_node={variance} flops={flops / 1e12:.2f} T) ")
    head += (f"\n\t2. memory: (d_w={m.dense_weight_bytes}, d_g={m.dense_grad_bytes}, d_s={m.dense_state_bytes}, "
             f"m_w={m.moe_weight_bytes}, m_g={m.moe_grad_bytes}, m_s={m.moe_state_bytes})")
    return head


def chunk_repr(chunk) -> str:
    return render(build_tree(chunk.nodes))

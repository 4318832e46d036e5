"""Computational-graph capture: an ONNX-like node/edge list of the evaluated op IR + optional graphviz render.

Enabled with ``ENABLE_SIMU_GRAPH=1`` (``PerfLLM.analysis`` then writes ``model_graph.json`` and
``computational_graph.*``) or explicitly via ``PerfLLM.capture(save_path)`` / ``run_estimate(capture_graph=True)``.
Because ops are already a flat list in execution order, capture is a pure export (no second "capture-only"
execution pass as in reference core/graph.py).
"""
from __future__ import annotations

import json
from typing import Dict, List


class SimuGraph:
    def __init__(self):
        self.nodes: List[dict] = []
        self.edges: List[dict] = []

    def add_node(self, name, op_type, inputs, outputs, attrs=None):
        self.nodes.append({"id": len(self.nodes), "name": name, "op_type": op_type, "inputs": inputs,
                           "outputs": outputs, "attrs": attrs or {}})
        return len(self.nodes) - 1

    def add_edge(self, src, dst, shape):
        self.edges.append({"src": src, "dst": dst, "shape": shape})

    def to_dict(self) -> Dict:
        return {"nodes": self.nodes, "edges": self.edges}

    def export_json(self, path: str):
        with open(path, "w", encoding="utf-8") as fh:
            json.dump(self.to_dict(), fh, indent=1)
        return path


def capture_graph(perf) -> SimuGraph:
    """One sub-graph per stage chunk; data edges follow the sequential op order (the model zoo is a chain with
    side inputs only inside attention / MoE blocks, which are recorded as multi-input nodes)."""
    g = SimuGraph()
    for chunk_name, chunk in perf.model_chunk_dict.items():
        if not chunk.nodes:
            chunk(perf._chunk_input_shape(chunk.preprocess))
        prev = None
        for n in chunk.nodes:
            attrs = {"stage": chunk_name, "scope": "/".join(n.scope), "fwd_flops": n.compute.fwd_flops,
                     "fwd_bytes": n.compute.fwd_accessed_mem, "fwd_ms": n.cost.fwd_compute_time,
                     "bwd_ms": n.cost.bwd_compute_time, "fwd_net_ms": n.cost.fwd_net_time,
                     "bwd_net_ms": n.cost.bwd_net_time, "cache_bytes": n.act.activation_mem_cache,
                     "recompute": n.enable_recompute, "weight_shape": n.weight_shape}
            nid = g.add_node(f"{chunk_name}:{n.name}", n.kind, n.in_shapes, n.out_shapes, attrs)
            if prev is not None:
                g.add_edge(prev, nid, n.in_shapes[0] if n.in_shapes else [])
            prev = nid
    return g


# reference-compatible alias
class SimuONNXGraphBuilder:
    def __init__(self):
        self.graph = SimuGraph()

    def reset(self):
        self.graph = SimuGraph()


def visualize_with_graphviz(json_path: str, output_path: str = "computational_graph", fmt: str = "png"):
    """Render with the ``graphviz`` package when available; otherwise write a .dot file next to the json."""
    with open(json_path, "r", encoding="utf-8") as fh:
        data = json.load(fh)
    lines = ["digraph G {", "  rankdir=TB; node [shape=box, fontsize=9];"]
    for n in data["nodes"]:
        label = f"{n['op_type']}\\n{n['name'].split(':')[-1]}\\n{n['outputs']}"
        lines.append(f"  n{n['id']} [label=\"{label}\"];")
    for e in data["edges"]:
        lines.append(f"  n{e['src']} -> n{e['dst']} [label=\"{e['shape']}\", fontsize=8];")
    lines.append("}")
    dot_path = output_path + ".dot"
    with open(dot_path, "w", encoding="utf-8") as fh:
        fh.write("\n".join(lines))
    try:
        import graphviz  # type: ignore
        graphviz.Source("\n".join(lines)).render(output_path, format=fmt, cleanup=True)
        return output_path + "." + fmt
    except Exception:
        return dot_path

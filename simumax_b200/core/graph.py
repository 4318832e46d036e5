"""Computational-graph capture: the reference's ONNX-style ``model_graph.json`` of the evaluated op IR + optional graphviz render.

Enabled with ``ENABLE_SIMU_GRAPH=1`` (``PerfLLM.analysis`` then writes ``model_graph.json`` and
``computational_graph.*``) or explicitly via ``PerfLLM.capture(save_path)`` / ``run_estimate(capture_graph=True)``.
Because ops are already a flat list in execution order, capture is a pure export (no second "capture-only"
execution pass as in reference core/graph.py).
"""
from __future__ import annotations

import json
from typing import Dict, List


class SimuGraph:
    """ONNX-style container with the reference's JSON schema (core/graph.py:96-124): ``name``, ``nodes`` (name =
    ``<attribute>_<call index>``, op_type, enable_recompute, is_variance_node, input / output tensor names, call_idx,
    attributes, domain), ``inputs`` / ``outputs`` / ``initializers`` and ``value_info`` (shape + dtype of every produced
    tensor).  Additions: per-node cost / memory numbers under ``attributes`` and an explicit ``edges`` list for the renderer."""

    def __init__(self, name: str = "model_graph"):
        self.name = name
        self.nodes: List[dict] = []
        self.edges: List[dict] = []
        self.value_info: Dict[str, dict] = {}
        self._tensors = 0

    def new_tensor(self, prefix: str, shape=None, dtype: str = "bf16") -> str:
        name = f"{prefix}_{self._tensors}"
        self._tensors += 1
        if prefix == "output":
            self.value_info[name] = {"shape": list(shape or []), "dtype": dtype}
        return name

    def add_node(self, attr_name, op_type, inputs, outputs, *, enable_recompute=False, is_variance_node=False, attrs=None,
                 call_idx=None):
        idx = len(self.nodes)                     # node names count over the whole graph, call_idx within the op's chunk
        self.nodes.append({"name": f"{attr_name}_{idx}", "op_type": op_type, "enable_recompute": bool(enable_recompute),
                           "is_variance_node": bool(is_variance_node), "inputs": list(inputs), "outputs": list(outputs),
                           "call_idx": idx if call_idx is None else call_idx, "attributes": attrs or {}, "domain": ""})
        return idx

    def add_edge(self, src, dst, shape):
        self.edges.append({"src": src, "dst": dst, "shape": shape})

    def to_dict(self) -> Dict:
        return {"name": self.name, "nodes": self.nodes, "inputs": [], "outputs": [], "initializers": [],
                "value_info": self.value_info, "edges": self.edges}

    def export_json(self, path: str):
        with open(path, "w", encoding="utf-8") as fh:
            json.dump(self.to_dict(), fh, indent=2)
        return path


# fan-in of the ops that do not simply consume the previous op's output: op attribute -> [(producer attribute, output index)]
# (MLA: reference dense_module.py:2755-2840; MoE block: moe_module.py ExpertMLP.forward).  "<norm>" = the norm that feeds the block.
_FAN_IN = {
    "linear_kv_down_proj": [("<norm>", 0)],
    "kv_layernorm": [("kv_combined_Split", 0)],
    "rotary_pos_emb": [("kv_combined_Split", 1)],
    "K_pos_emb_Concat": [("KV_Split", 0), ("rotary_pos_emb", 0)],
    "QKV_Concat": [("<q>", 0), ("K_pos_emb_Concat", 0), ("KV_Split", 1)],
    "router": [("<norm>", 0)],
    "permutation": [("<norm>", 0), ("router", 0)],
    "unpermutation": [("group_linear2", 0), ("router", 0)],
    "SharedExpertAddFunction": [("unpermutation", 0), ("<shared>", 0)],
}
_OUT_DTYPE = {"Float8Quantizer": "fp8", "Router": "int32"}


def capture_graph(perf) -> SimuGraph:
    """Export of every stage chunk's op list, in chunk order, into one graph (the reference records one forward pass per chunk
    with shared node / tensor counters; each chunk starts from a fresh ``input_N`` tensor).  Data edges follow the execution
    order: dense models are chains and come out identical to the reference's file; ops with several inputs / outputs (MLA
    split / concat, MoE permutation) consume every tensor the previous op produced and name one tensor per output shape --
    same schema, simplified wiring.  GeLU returns its input tensor object in the reference, so it does not name a new one."""
    g = SimuGraph()
    for chunk_name, chunk in perf.model_chunk_dict.items():
        if not chunk.nodes:
            chunk(perf._chunk_input_shape(chunk.preprocess))
        prev_outputs, prev = None, None
        made: Dict[str, list] = {}                     # most recent outputs per op attribute (and per enclosing wrapper)
        for k, n in enumerate(chunk.nodes):
            parts = n.name.split(".")
            attr = parts[-1]
            ins = prev_outputs if prev_outputs is not None else [g.new_tensor("input")]
            wired = []
            for src, idx in _FAN_IN.get(attr, []) if attr in _FAN_IN else _FAN_IN.get(parts[-2] if len(parts) > 1 else "", []):
                outs_of = made.get(src)
                if outs_of is None or idx >= len(outs_of):
                    wired = []
                    break
                wired.append(outs_of[idx])
            if wired:
                ins = wired
            if n.kind == "Gelu":
                outs = list(ins)
            else:
                outs = [g.new_tensor("output", shape, _OUT_DTYPE.get(n.kind, "bf16")) for shape in n.out_shapes]
            attrs = {"stage": chunk_name, "scope": "/".join(n.scope), "full_name": n.name, "module_path": n.path,
                     "input_shapes": n.in_shapes, "fwd_flops": n.compute.fwd_flops, "fwd_bytes": n.compute.fwd_accessed_mem,
                     "fwd_ms": n.cost.fwd_compute_time, "bwd_ms": n.cost.bwd_compute_time, "fwd_net_ms": n.cost.fwd_net_time,
                     "bwd_net_ms": n.cost.bwd_net_time, "cache_bytes": n.act.activation_mem_cache, "weight_shape": n.weight_shape}
            nid = g.add_node(attr, n.kind, ins, outs, enable_recompute=n.enable_recompute,
                             is_variance_node=n.is_variance_node, attrs=attrs, call_idx=k)
            if prev is not None:
                g.add_edge(prev, nid, n.in_shapes[0] if n.in_shapes else [])
            prev, prev_outputs = nid, outs
            made[attr] = outs
            if len(parts) > 2:
                made[parts[-2]] = outs                 # fp8 wrapper: <linear attr>.linear
            if n.kind == "LayerNorm" and attr in ("layernorm_input", "pre_mlp_layernorm"):
                made["<norm>"] = outs
            if attr in ("linear_q_up_proj", "linear_q_proj") or (len(parts) > 1 and parts[-2] in ("linear_q_up_proj", "linear_q_proj")):
                made["<q>"] = outs
            if "shared_expert" in parts and n.kind == "LinearRow":
                made["<shared>"] = outs
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
    shapes = data.get("value_info", {})
    produced = {}
    for pos, n in enumerate(data["nodes"]):
        n["_pos"] = pos
    for n in data["nodes"]:
        outs = [shapes.get(t, {}).get("shape", t) for t in n["outputs"]]
        color = ", style=filled, fillcolor=lightyellow" if n.get("enable_recompute") else ""
        label = f"{n['op_type']}\\n{n['name']}\\n{outs}"
        lines.append(f"  n{n['_pos']} [label=\"{label}\"{color}];")
        for t in n["outputs"]:
            produced.setdefault(t, n["_pos"])
    for n in data["nodes"]:                      # edges from tensor names: works for reference-made model_graph.json too
        for src in sorted({produced[t] for t in n["inputs"] if t in produced and produced[t] != n["_pos"]}):
            shape = [shapes.get(t, {}).get("shape") for t in n["inputs"] if produced.get(t) == src]
            lines.append(f"  n{src} -> n{n['_pos']} [label=\"{shape[0] if len(shape) == 1 else shape}\", fontsize=8];")
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

"""Model assembly: transformer block / full stage chunk as a flat, evaluated op list.

Structure being modelled (reference transformer/language_model.py:98-353, dense_module.py:2454-2976):

    chunk  = [Embedding] + N x block + [final RMSNorm + vocab LinearCol + ParallelCE]
    block  = RMSNorm -> attention -> RMSNorm -> mlp            (residual adds are not modelled)
    attn   = LinearCol(qkv) -> CoreAttention -> LinearRow      (MLA: see models/mla.py)
    mlp    = LinearCol(fc1) -> SwiGLU|GeLU -> LinearRow(fc2)   (MoE: see models/moe.py)

Op names reproduce the reference's dotted attribute paths (``self.layer_3.mlp.linear_fc1``) because they show
up in result files (peak_path) and in trace scopes.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, List, Optional

from ..core.config import AttentionRecomputeConfig, MLPRecomputeConfig, ModelConfig, StrategyConfig, SystemConfig
from ..core.ir import PHASES, OpNode, evaluate, ishape
from ..core.memory import assign_recompute_status, walk_activations
from ..core.records import ComputeInfo, CostInfo, MemoryInfo, PeakPoint
from . import dense as D


class Emitter:
    """Collects ops in execution order while tracking the scope (composite-module) stack."""

    def __init__(self, model: ModelConfig, strategy: StrategyConfig, system: SystemConfig):
        self.model, self.st, self.system = model, strategy, system
        self.nodes: List[OpNode] = []
        self._names: List[str] = ["self"]
        self._scopes: List[str] = []
        self.scope_tree: List[tuple] = []      # (depth, scope display name, first op index, last op index)
        # reference module-tree path bookkeeping: every composite / leaf takes the next child index of its parent, in call order
        self._child_counts: List[int] = [0]
        self._ref_path: List[str] = ["(0)LLMModel"]

    # naming -------------------------------------------------------------------------------------
    def name(self, leaf: str) -> str:
        return ".".join(self._names + [leaf])

    class _Scope:
        def __init__(self, em, attr, display, ref_cls=None):
            self.em, self.attr, self.display, self.ref_cls = em, attr, display, ref_cls

        def __enter__(self):
            em = self.em
            idx = em._child_counts[-1]
            em._child_counts[-1] += 1
            em._child_counts.append(0)
            em._ref_path.append(f"({idx}){self.ref_cls or em.ref_class(self.display)}")
            self.em._names.append(self.attr)
            self.em._scopes.append(self.display)
            self.start = len(self.em.nodes)
            self.depth = len(self.em._scopes)
            return self

        def __exit__(self, *exc):
            self.em.scope_tree.append((self.depth, self.display, self.start, len(self.em.nodes), tuple(self.em._scopes)))
            self.em._names.pop()
            self.em._scopes.pop()
            self.em._child_counts.pop()
            self.em._ref_path.pop()

    def scope(self, attr: str, display: str, ref_cls: str = None):
        return Emitter._Scope(self, attr, display, ref_cls)

    def ref_class(self, display: str) -> str:
        """Reference class name of a composite scope (display names follow Megatron's module names instead)."""
        if display.startswith("TransformerLayer"):
            return "LLMBlock"
        if display == "SelfAttention":
            return "MLAAttention" if getattr(self.model, "attention_type", None) == "mla" else "Attention"
        if display == "MoELayer":
            return "ExpertMLP"
        return display

    def add(self, node: OpNode) -> OpNode:
        idx = self._child_counts[-1]
        self._child_counts[-1] += 1
        node.ref_path = " -> ".join(self._ref_path + [f"({idx}){node.kind}"])
        node.scope = tuple(self._scopes)
        self.nodes.append(node)
        return node


# ------------------------------------------------------------------------------------------------
# composites
# ------------------------------------------------------------------------------------------------
def _maybe_quantize(em: Emitter, attr: str, shape, enable_recompute: bool):
    """fp8: every linear is preceded by a cast op (QuantizedCol/RowLinear, dense_module.py:2397-2452)."""
    if not em.st.fp8:
        return None
    node, _ = D.fp8_quantizer(em.st, shape, em.name("quntizer"), enable_recompute=enable_recompute)
    return em.add(node)


def _linear(em: Emitter, attr: str, kind: str, in_f, out_f, shape, display, **kw):
    """Emit a (possibly fp8-wrapped) linear under attribute ``attr``; returns (linear node, out shape)."""
    ctor = D.linear_col if kind == "col" else D.linear_row
    if em.st.fp8 and kw.pop("wrap_fp8", True):
        # both wrappers are *displayed* as QuantizedColLinear in call stacks (the row variant keeps the column one's
        # specific_name, dense_module.py:2438), the module tree shows the real class
        with em.scope(attr, "QuantizedColLinear", ref_cls="QuantizedColLinear" if kind == "col" else "QuantizedRowLinear"):
            _maybe_quantize(em, attr, shape, kw.get("enable_recompute", False))
            node, out = ctor(em.st, in_f, out_f, shape, em.name("linear"), display=display, **kw)
            em.add(node)
        return node, out
    kw.pop("wrap_fp8", None)
    node, out = ctor(em.st, in_f, out_f, shape, em.name(attr), display=display, **kw)
    em.add(node)
    return node, out


def emit_attention(em: Emitter, layer_idx: int, shape, rc: AttentionRecomputeConfig):
    cfg, st = em.model, em.st
    qkv_out = cfg.head_num * cfg.head_size + 2 * cfg.kv_head_num * cfg.head_size
    tail = rc.megatron_layernorm
    with em.scope("attention", "SelfAttention"):
        _, shape = _linear(em, "linear_qkv", "col", cfg.hidden_size, qkv_out, shape, "ColumnParallelLinear",
                           layer_idx=layer_idx, has_cached_inputs=tail,
                           enable_recompute=rc.q_up_recompute or tail, is_last_recompute=tail,
                           use_variance_tail_model=tail)
        node, shape = D.core_attention(st, em.system, cfg.head_size, cfg.head_num, cfg.kv_head_num, shape,
                                       em.name("attention"), enable_recompute=rc.core_attn_recompute,
                                       is_last_recompute=True, cache_outputs=bool(st.use_flash_sdp and st.fp8))
        em.add(node)
        _, shape = _linear(em, "linear_out", "row", cfg.head_num * cfg.head_size, cfg.hidden_size, shape,
                           "RowParallelLinear", layer_idx=layer_idx, enable_recompute=rc.out_recompute)
    return shape


def emit_mlp(em: Emitter, layer_idx, shape, rc: MLPRecomputeConfig, *, attr="mlp", display="MLP",
             intermediate_size=None):
    cfg, st = em.model, em.st
    is_shared = isinstance(layer_idx, str) and "shareExpert" in layer_idx
    dense_ckpt = rc.linear_recompute or (rc.megatron_mlp and not is_shared)
    shared_ckpt = rc.shared_linear_recompute or (rc.megatron_moe and is_shared)
    local_inter = cfg.intermediate_size if intermediate_size is None else intermediate_size
    fc1_out = 2 * local_inter if cfg.use_swiglu else local_inter
    enable = shared_ckpt if is_shared else dense_ckpt
    tail = rc.megatron_layernorm and not is_shared
    with em.scope(attr, display):
        fc1, shape = _linear(em, "linear_fc1", "col", cfg.hidden_size, fc1_out, shape, "ColumnParallelLinear",
                             layer_idx=layer_idx, has_cached_inputs=tail, enable_recompute=enable or tail,
                             is_last_recompute=tail, use_variance_tail_model=tail)
        if st.recompute_granularity == "selective_recompute" and rc.megatron_layernorm and enable:
            fc1.is_breakpoint = True
        if cfg.use_swiglu:
            node, shape = D.swiglu(st, shape, em.name("activation_layer"), enable_recompute=enable)
        else:
            node, shape = D.gelu(st, shape, em.name("activation_layer"), enable_recompute=enable)
        em.add(node)
        _, shape = _linear(em, "linear_fc2", "row", local_inter, cfg.hidden_size, shape, "RowParallelLinear",
                           layer_idx=layer_idx, enable_recompute=enable, is_last_recompute=True)
    return shape


def emit_block(em: Emitter, layer_idx: int, shape, use_dense: bool):
    cfg, st = em.model, em.st
    arc = st.parse_attention_recompute(layer_idx)
    mrc = st.parse_mlp_recompute(layer_idx)
    with em.scope(f"layer_{layer_idx}", f"TransformerLayer{layer_idx}"):
        first_index = len(em.nodes)
        node, shape = D.layernorm(st, cfg.hidden_size, shape, em.name("layernorm_input"),
                                  enable_recompute=arc.input_layernorm_recompute)
        em.add(node)
        if getattr(cfg, "attention_type", None) == "mla":
            from .mla import emit_mla_attention
            shape = emit_mla_attention(em, layer_idx, shape, arc)
        else:
            shape = emit_attention(em, layer_idx, shape, arc)
        node, shape = D.layernorm(st, cfg.hidden_size, shape, em.name("pre_mlp_layernorm"),
                                  enable_recompute=mrc.pre_mlp_norm_recompute)
        em.add(node)
        if cfg.expert_num == 1 or use_dense:
            shape = emit_mlp(em, layer_idx, shape, mrc)
        else:
            from .moe import emit_expert_mlp
            shape = emit_expert_mlp(em, layer_idx, shape, mrc)
    return shape, first_index


# ------------------------------------------------------------------------------------------------
# search-time block cache
# ------------------------------------------------------------------------------------------------
class BlockProfileCache:
    """Emitted + evaluated transformer blocks, shared between layers, chunks and search candidates.

    Counterpart of the reference's search cache (``CachedChunkProfile`` / ``CachedUnitRuntimeProfile``, perf_llm.py:41-252,
    837-1379: per-unit profiles keyed by a projection of the strategy JSON).  With the flat IR a "unit" is just the list of
    ``OpNode`` a block emits: ``evaluate()`` does not look at neighbours, so a block that was emitted and evaluated once can be
    re-used for every layer with the same (strategy projection, model, recompute switches, input shape) by cloning the nodes
    under new names.  The strategy projection drops the knobs that only move whole blocks around (micro-batch count, how many
    layers are recomputed, pipeline layer split): those are captured by the per-layer recompute configs and by which blocks a
    chunk asks for.  Only the search entry points switch it on, so result files of plain ``perf()`` runs are untouched.
    """
    _VOLATILE = ("micro_batch_num", "global_batch_size", "recompute_layer_num", "num_layers_in_first_pipeline_stage",
                 "num_layers_in_last_pipeline_stage")

    def __init__(self):
        self.store: Dict[tuple, tuple] = {}
        self.hits = 0
        self.misses = 0

    def context_key(self, cfg: ModelConfig, st: StrategyConfig, system: SystemConfig) -> tuple:
        import json
        sd = {k: v for k, v in st.to_dict().items() if not any(t in k for t in self._VOLATILE)}
        return (id(system), json.dumps(cfg.to_dict(), sort_keys=True, default=str), json.dumps(sd, sort_keys=True, default=str))

    def clear(self):
        self.store.clear()


_BOOKS = ("real_comm_bw", "miss_efficiency", "hit_efficiency")


def _swap_books(system: SystemConfig, books):
    """Install ``books`` (or fresh empty tables) as the system's bookkeeping dicts; returns the previous ones."""
    from collections import OrderedDict
    prev = tuple(getattr(system, b) for b in _BOOKS)
    for b, v in zip(_BOOKS, books or (OrderedDict(), OrderedDict(), OrderedDict())):
        setattr(system, b, v)
    return prev


def _merge_books(system: SystemConfig, delta) -> None:
    for b, d in zip(_BOOKS, delta):
        dst = getattr(system, b)
        for op, entries in d.items():
            if isinstance(entries, dict) and isinstance(dst.get(op), dict):
                dst[op].update(entries)
            elif isinstance(entries, dict):
                dst[op] = dict(entries)
            else:
                dst[op] = entries


def _retarget(text: str, src: int, dst: int, prefix: str) -> str:
    return text.replace(f"{prefix}{src}", f"{prefix}{dst}", 1) if src != dst else text


def _clone_block(nodes: List[OpNode], tree: List[tuple], src: int, dst: int, base: int, block_idx: Optional[int] = None):
    """Copies of a cached block's nodes / scope records renamed from layer ``src`` to ``dst``, node indices rebased to ``base``;
    ``block_idx`` = child index of the block under the root (second component of the reference-style module path)."""
    from copy import copy
    out = []
    for n in nodes:
        m = copy(n)
        if block_idx is not None and n.ref_path:
            parts = n.ref_path.split(" -> ")
            if len(parts) > 1:
                parts[1] = f"({block_idx})" + parts[1].split(")", 1)[1]
                m.ref_path = " -> ".join(parts)
        m.name = _retarget(n.name, src, dst, "layer_")
        m.scope = tuple(_retarget(sn, src, dst, "TransformerLayer") for sn in n.scope)
        if isinstance(n.layer_idx, int) and n.layer_idx == src:
            m.layer_idx = dst
        elif isinstance(n.layer_idx, str):
            m.layer_idx = n.layer_idx.replace(str(src), str(dst), 1) if src != dst else n.layer_idx
        m.act = copy(n.act)                     # the lifetime walker writes cache_for_bwd_mem
        m.recompute_status, m.recompute_done = n.recompute_status, False
        out.append(m)
    scopes = [(d, _retarget(disp, src, dst, "TransformerLayer"), lo + base, hi + base,
               tuple(_retarget(sn, src, dst, "TransformerLayer") for sn in sc)) for d, disp, lo, hi, sc in tree]
    return out, scopes


# ------------------------------------------------------------------------------------------------
# stage chunk
# ------------------------------------------------------------------------------------------------
class StageChunk:
    """All ops a pipeline stage executes for one micro-batch, evaluated against a system config.

    API-compatible stand-in for the reference ``LLMModel`` instance stored in ``PerfLLM.model_chunk_dict``.
    """

    def __init__(self, layer_num: int, dense_layers: int = 0, preprocess: bool = True, postprocess: bool = True,
                 model_config: ModelConfig = None, strategy: StrategyConfig = None, system: SystemConfig = None,
                 specific_name: str = "GPTModel_0", block_cache: Optional[BlockProfileCache] = None):
        self.block_cache = block_cache
        self.layer_num = layer_num
        self.dense_layers = dense_layers
        self.preprocess, self.postprocess = preprocess, postprocess
        self.model_config = deepcopy(model_config)
        self.strategy, self.system = strategy, system
        self.specific_name = specific_name
        self.nodes: List[OpNode] = []
        self.layer_first_nodes: List[OpNode] = []
        self.scope_tree: List[tuple] = []
        self.input_shape = None
        self._model_info = MemoryInfo()
        self._compute_info = ComputeInfo()
        self._cost_info = CostInfo()
        self._info_ready = False

    # -- build + evaluate ---------------------------------------------------------------------------
    def __call__(self, input_shape, _ctx=None):
        cfg, st = self.model_config, self.strategy
        em = Emitter(cfg, st, self.system)
        shape = ishape(getattr(input_shape, "shape", input_shape))
        self.input_shape = shape
        if self.preprocess:
            node, shape = D.embedding(st, cfg.hidden_size, cfg.vocab_size, shape, em.name("embedding"))
            em.add(node)
        firsts = []
        cache = self.block_cache
        ctx = cache.context_key(cfg, st, self.system) if cache is not None else None
        cached_nodes = set()
        if cache is not None:                 # keep the bookkeeping order of the uncached path: ops are recorded in node order
            for n in em.nodes:
                evaluate(n, self.system)
                cached_nodes.add(id(n))
        for i in range(self.layer_num):
            use_dense = i < self.dense_layers
            if cache is None:
                shape, first_idx = emit_block(em, i, shape, use_dense=use_dense)
                firsts.append(first_idx)
                continue
            key = (ctx, use_dense, repr(st.parse_attention_recompute(i)), repr(st.parse_mlp_recompute(i)), tuple(shape))
            hit = cache.store.get(key)
            if hit is None:
                cache.misses += 1
                n0, t0 = len(em.nodes), len(em.scope_tree)
                shape_out, first_idx = emit_block(em, i, shape, use_dense=use_dense)
                block = em.nodes[n0:]
                # evaluate against empty bookkeeping tables so the block's own records (efficiency hits / misses, achieved
                # collective bandwidths) can be replayed whenever the cached block is reused
                books = _swap_books(self.system, None)
                for n in block:
                    evaluate(n, self.system)
                delta = _swap_books(self.system, books)
                _merge_books(self.system, delta)
                tree = [(d, disp, lo - n0, hi - n0, sc) for d, disp, lo, hi, sc in em.scope_tree[t0:]]
                # keep pristine copies: the nodes that stay in this chunk get statuses / walker scratch written to them
                tpl, _ = _clone_block(block, [], i, i, 0)
                cache.store[key] = (tpl, tree, i, first_idx - n0, shape_out, delta)
                cached_nodes.update(id(n) for n in block)
                shape = shape_out
            else:
                cache.hits += 1
                tpl, tree, src, first_rel, shape, delta = hit
                _merge_books(self.system, delta)
                base = len(em.nodes)
                block_idx = em._child_counts[-1]          # the block takes the next child slot of the root, as emit_block would
                em._child_counts[-1] += 1
                nodes, scopes = _clone_block(tpl, tree, src, i, base, block_idx)
                em.nodes.extend(nodes)
                em.scope_tree.extend(scopes)
                first_idx = base + first_rel
                cached_nodes.update(id(n) for n in nodes)
            firsts.append(first_idx)
        if self.postprocess:
            node, shape = D.layernorm(st, cfg.hidden_size, shape, em.name("layernorm"))
            em.add(node)
            node, shape = D.linear_col(st, cfg.hidden_size, cfg.vocab_size, shape, em.name("linear_out"),
                                       layer_idx=-1, enable_fp8=False, display="ColumnParallelLinear")
            em.add(node)
            node, shape = D.parallel_ce(st, shape, em.name("parallel_ce"))
            em.add(node)
        self.nodes = em.nodes
        self.scope_tree = em.scope_tree
        self.layer_first_nodes = [self.nodes[i] for i in firsts if i < len(self.nodes)]
        assign_recompute_status(self.nodes)
        for n in self.nodes:
            if id(n) not in cached_nodes:
                evaluate(n, self.system)
        from ..core.arch_repr import hierarchical_totals
        self._model_info, self._compute_info, self._cost_info = hierarchical_totals(self.nodes)
        self._info_ready = True
        return shape

    def compute_activations(self) -> PeakPoint:
        return walk_activations(self.nodes, self.layer_first_nodes, self.strategy.enable_recompute)

    # -- reference-compatible accessors ---------------------------------------------------------------
    @property
    def main_grad_element_size(self) -> int:
        from ..core.ir import main_grad_bytes
        return main_grad_bytes(self.strategy)

    def get_model_info(self) -> MemoryInfo:
        assert self._info_ready, "model info not ready, please call the module to compute info"
        return self._model_info

    def get_compute_info(self) -> ComputeInfo:
        assert self._info_ready
        return self._compute_info

    def get_cost_info(self) -> CostInfo:
        assert self._info_ready
        return self._cost_info

    def get_all_leaf_modules(self) -> List[OpNode]:
        return self.nodes

    def get_all_gemm_cost_info(self) -> Dict[str, list]:
        """Per-GEMM table (fwd / dgrad / wgrad rows) -- also the shape enumeration the calibration sweeps."""
        cols = ["Module", "type", "B", "M", "K", "N", "layout", "accumulate", "out_dtype", "compute_cost",
                "memory_cost", "cost", "bound"]
        out = {c: [] for c in cols}
        for n in self.nodes:
            if n.gemm is None or n.kind not in ("LinearCol", "LinearRow", "GroupLinearCol", "GroupLinearRow", "Router"):
                continue
            table = n.gemm
            if n.kind.startswith("GroupLinear"):
                # the reference lists a grouped GEMM as one flat GEMM over all local experts' tokens (B = 1, the token dimension
                # -- M of fwd / dgrad, K of wgrad -- multiplied by the expert count; moe_module.py GroupLinearBase)
                g = table["B"]
                table = dict(table, B=[1, 1, 1], M=[table["M"][0] * g[0], table["M"][1] * g[1], table["M"][2]],
                             K=[table["K"][0], table["K"][1], table["K"][2] * g[2]])
            for key in ("B", "M", "K", "N", "layout", "accumulate", "out_dtype"):
                out[key].extend(table[key])
            cc = [n.details[p]["compute_details"]["compute_only_time"] for p in PHASES]
            mc = [n.details[p]["io_details"]["io_time"] for p in PHASES]
            out["compute_cost"].extend(cc)
            out["memory_cost"].extend(mc)
            out["bound"].extend(["IO bound" if m > c else "compute bound" for m, c in zip(mc, cc)])
            out["cost"].extend(n.cost.get_all_costs())
            out["Module"].extend([n.name + ".fwd", n.name + ".bwd_act", n.name + ".bwd_w"])
            out["type"].extend([n.kind] * 3)
        return out

    def analysis_op_info(self, return_details: bool = False) -> Dict[str, list]:
        """Per-op table: shapes, flops, IO bytes, time, roofline side (language_model.py:514-595)."""
        cols = ["op", "input_shapes", "output_shapes", "flops", "IO", "cost", "compute_only_time", "IO_time", "bound"]
        out = {c: [] for c in cols}
        if return_details:
            out["compute_only_details"], out["IO_details"] = [], []

        def push(label, ins, outs, flops, io, cost, det):
            out["op"].append(label)
            out["input_shapes"].append(ins)
            out["output_shapes"].append(outs)
            out["flops"].append(flops)
            out["IO"].append(io)
            out["cost"].append(cost)
            ct, it = det["compute_details"]["compute_only_time"], det["io_details"]["io_time"]
            out["compute_only_time"].append(ct)
            out["IO_time"].append(it)
            out["bound"].append("IO bound" if it > ct else "Compute bound")
            if return_details:
                out["compute_only_details"].append(det["compute_details"])
                out["IO_details"].append(det["io_details"])

        for n in self.nodes:
            w = [n.weight_shape] if n.weight_shape is not None else []
            is_linear = n.gemm is not None
            push(n.kind, n.in_shapes + (w if not is_linear else []), n.out_shapes, n.compute.fwd_flops,
                 n.compute.fwd_accessed_mem, n.cost.fwd_compute_time, n.details["fwd"])
            wt = [] if is_linear else w          # only modules with a ``weight`` attribute list it (language_model.py:551-557)
            push(n.kind + "_bwd_act", n.out_shapes + wt, n.in_shapes, n.compute.bwd_grad_act_flops,
                 n.compute.bwd_grad_act_accessed_mem, n.cost.bwd_grad_act_time, n.details["bwd_grad_act"])
            if is_linear and n.weight_shape:
                lhs = [list(n.in_shapes[0][:-2]) + [n.in_shapes[0][-1], n.in_shapes[0][-2]]] if len(n.in_shapes[0]) >= 2 else n.in_shapes
                push(n.kind + "_bwd_w", lhs + n.out_shapes, [n.weight_shape], n.compute.bwd_grad_w_flops,
                     n.compute.bwd_grad_w_accessed_mem, n.cost.bwd_grad_w_time, n.details["bwd_grad_w"])
        return out

    def __repr__(self) -> str:
        """The reference's nested module view with per-module totals (base_struct.py:1004-1068); ``summary()`` is the flat table."""
        from ..core.arch_repr import chunk_repr
        return chunk_repr(self)

    def summary(self) -> str:
        lines = [f"{self.specific_name}(layers={self.layer_num}, pre={self.preprocess}, post={self.postprocess})"]
        for n in self.nodes:
            c = n.cost
            lines.append(f"  {n.name:<60s} {n.kind:<18s} fwd={c.fwd_compute_time * 1e3:9.2f}us net={c.fwd_net_time * 1e3:8.2f}us "
                         f"bwd={c.bwd_compute_time * 1e3:9.2f}us net={c.bwd_net_time * 1e3:8.2f}us "
                         f"cache={n.act.activation_mem_cache / 2 ** 20:9.2f}MB rc={n.recompute_status}")
        return "\n".join(lines)


# reference class name, so `from ... import LLMModel` style code keeps working
LLMModel = StageChunk

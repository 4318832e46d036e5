"""Operator IR: one flat list of ``OpNode`` per (model, strategy, pipeline-stage chunk).

Design (differs from the reference, which recomputes five kinds of info inside a recursive
``MetaModule.__call__`` with deep copies -- base_struct.py:929-991):

* a *builder* (``simumax_b200.models``) walks the architecture once and emits ``OpNode`` records in
  execution order; each record carries closed-form flops / bytes / activation / static-memory / collective
  specs for fwd, dgrad and wgrad;
* ``evaluate(node, system)`` is a pure function turning those into times via the three cost primitives;
* scopes (attention, mlp, layer, model) are just name prefixes; totals are sums over the flat list.

The numbers each node must produce are specified leaf by leaf in the reference
(transformer/dense_module.py, transformer/moe_module.py); the per-leaf docstrings cite the lines.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from .config import SIMU_DEBUG, StrategyConfig, SystemConfig, debug_trace
from .records import ActivationInfo, ComputeInfo, CostInfo, MemoryInfo, RecomputeStatus

PHASES = ("fwd", "bwd_grad_act", "bwd_grad_w")


@dataclass
class CollectiveSpec:
    """One modelled collective issued by an op in a given phase."""
    op: str                 # all_reduce | all_gather | reduce_scatter | all2all | p2p
    size: float             # logical payload bytes
    comm_num: int           # group size
    net: str                # network class name in the system config
    stage: str              # bookkeeping tag (also selects inter-node sharing rules)
    phase: str              # fwd | bwd_grad_act | bwd_grad_w
    group: str = "tp"       # which parallel group (tp | cp | ep | etp) -- used by the simulator
    time: float = 0.0       # filled by evaluate()
    exposed: float = 1.0    # fraction of the collective's time that is not hidden inside the adjacent GEMM (fused kernels)


@dataclass
class SimStep:
    """Simulator-facing expansion of an op: an ordered list of compute atoms and collectives."""
    kind: str                       # "atom" | "comm"
    name: str = ""
    fwd_cost: float = 0.0
    bwd_cost: float = 0.0
    comm_op: str = ""               # all_gather / reduce_scatter / all_reduce / all_gather_bwd / all2all_fwd / all2all_bwd
    group: str = "tp"
    comm_num: int = 1
    stage: str = ""


@dataclass
class OpNode:
    name: str                        # dotted full name, e.g. self.layer_0.attention.linear_qkv
    kind: str                        # reference class name (Embedding, LinearCol, ...)
    scope: Tuple[str, ...] = ()      # enclosing composite names, outermost first
    ref_path: str = ""               # "(0)LLMModel -> (4)LLMBlock -> (3)MLP -> (1)Swiglu": the reference's module-tree path
    display: str = ""                # trace/call-stack label of the op
    repr_args: str = ""              # the reference's extra_repr() of this module class (model_arch tree view)
    layer_idx: Optional[int] = None
    in_shapes: List[List[int]] = field(default_factory=list)
    out_shapes: List[List[int]] = field(default_factory=list)
    in_bytes: float = 0              # bytes of all inputs (kept by FIRST nodes of a recompute segment)
    weight_shape: Optional[List[int]] = None

    # recompute
    enable_recompute: bool = False
    recompute_status: str = RecomputeStatus.NO_RECOMPUTE
    is_breakpoint: bool = False
    is_variance_node: bool = False
    offload_inputs: bool = False
    recompute_done: bool = False     # scratch for the lifetime walker

    # op kinds for the efficiency / bandwidth tables
    ops: Dict[str, str] = field(default_factory=lambda: {p: "default" for p in PHASES})
    shape_key: Callable[[str], str] = field(default=lambda phase: "", repr=False)
    gemm: Optional[Dict[str, list]] = None      # B/M/K/N/layout/accumulate/out_dtype for fwd,dgrad,wgrad

    compute: ComputeInfo = field(default_factory=ComputeInfo)
    act: ActivationInfo = field(default_factory=ActivationInfo)
    mem: MemoryInfo = field(default_factory=MemoryInfo)
    comms: List[CollectiveSpec] = field(default_factory=list)
    recompute_net_follows_fwd: bool = False     # linears replay their fwd collective when recomputed

    cost: CostInfo = field(default_factory=CostInfo)
    details: Dict[str, dict] = field(default_factory=dict)
    sim_steps: Callable[["OpNode"], List[SimStep]] = field(default=None, repr=False)

    # -- reference-compatible accessors -----------------------------------------------------------
    @property
    def full_name(self) -> str:
        return self.name

    @property
    def path(self) -> str:
        """The reference's ``current_full_module_path`` (child index + class name per level, base_struct.py:947-958) when the
        emitter recorded it, else the scope display names."""
        return self.ref_path or " -> ".join(self.scope + (self.kind,))

    def get_act_info(self) -> ActivationInfo:
        return self.act

    def get_cost_info(self) -> CostInfo:
        return self.cost

    def get_compute_info(self) -> ComputeInfo:
        return self.compute

    def get_model_info(self) -> MemoryInfo:
        return self.mem

    def all_input_element_num(self) -> float:
        return self.in_bytes


def evaluate(node: OpNode, system: SystemConfig) -> OpNode:
    """Fill ``node.cost`` / ``node.details`` from flops, bytes and collectives (roofline per phase).

    time(phase) = max(flops / (tflops * eff[shape_key]), bytes / (gbps * 2^30 * eff) + latency)
    (reference base_struct.py:821-858, config.py:815-893,1019-1035).
    """
    c = node.compute
    if node.enable_recompute and node.kind != "Float8Quantizer":     # the cast's replay is timed but not counted in the traffic
        c.recompute_flops = c.fwd_flops                               # statistic (reference dense_module.py:2384-2388)
        c.recompute_accessed_mem = c.fwd_accessed_mem
    else:
        c.recompute_flops = 0
        c.recompute_accessed_mem = 0

    cost = CostInfo()
    for spec in node.comms:
        spec.time = system.compute_net_op_time(spec.op, spec.size, comm_num=spec.comm_num, net=spec.net,
                                               comm_stage=spec.stage) * spec.exposed
        if spec.phase == "fwd":
            cost.fwd_net_time += spec.time
        elif spec.phase == "bwd_grad_act":
            cost.bwd_grad_act_net_time += spec.time
        else:
            cost.bwd_grad_w_net_time += spec.time
    if node.recompute_net_follows_fwd and node.enable_recompute:
        cost.recompute_net_time = cost.fwd_net_time
    if node.is_variance_node:
        c.recompute_accessed_mem = 0
        c.recompute_flops = 0
        cost.recompute_net_time = 0
        cost.recompute_net_exposed_time = 0
        if node.enable_recompute:
            if SIMU_DEBUG:
                debug_trace(f"{node.path}: tail of a recompute segment, its forward is not re-run (no recompute flops / bytes / collectives)")

    launch_ms = getattr(system.accelerator, "kernel_launch_overhead_us", 0.0) / 1e3

    def phase_time(phase, flops, nbytes):
        op_name = node.ops[phase]
        cd = system.compute_op_accuracy_time(op_name, flops, shape_desc=node.shape_key(phase), reture_detail=True)
        io = system.compute_mem_access_time(op_name, nbytes, reture_detail=True)
        node.details[phase] = {"compute_details": cd, "io_details": io}
        t = system.compute_end2end_time(compute_time=cd["compute_only_time"], mem_time=io["io_time"])
        if launch_ms and (flops or nbytes):
            t += launch_ms                    # extension: per-kernel launch gap (0 unless the system config sets it)
        return t

    cost.fwd_compute_time = phase_time("fwd", c.fwd_flops, c.fwd_accessed_mem)
    cost.bwd_grad_act_time = phase_time("bwd_grad_act", c.bwd_grad_act_flops, c.bwd_grad_act_accessed_mem)
    cost.bwd_grad_w_time = phase_time("bwd_grad_w", c.bwd_grad_w_flops, c.bwd_grad_w_accessed_mem)
    cost.recompute_compute_time = cost.fwd_time if node.enable_recompute else 0
    if node.enable_recompute and node.is_variance_node:
        cost.recompute_compute_time = 0
    node.cost = cost
    hook = getattr(node, "post_evaluate", None)
    if hook is not None:
        hook(node)
    return node


def comm_step(system: SystemConfig, op: str, size, comm_num: int, net: str, label: str, *, fwd: bool = True,
              bwd: bool = True, group: str = "tp", stage: str = "unkonw", exposed: float = 1.0,
              exposed_bwd: Optional[float] = None) -> SimStep:
    """Simulator collective with its own cost lookup (the replay uses slightly different collectives than the
    analytic model in a few places, e.g. reduce-scatter for the SP embedding; reference prefill() methods)."""
    full = system.compute_net_op_time(op, size, comm_num=comm_num, net=net, comm_stage=stage)
    cost = full * exposed                     # exposure < 1: the collective runs inside the adjacent GEMM (tp_overlap_model); the
    back = full * (exposed if exposed_bwd is None else exposed_bwd)       # mirrored collective of the backward pass has its own factor
    return SimStep(kind="comm", name=label, fwd_cost=cost if fwd else 0.0, bwd_cost=back if bwd else 0.0,
                   comm_op=op, group=group, comm_num=comm_num, stage=stage)


def default_sim_steps(node: OpNode) -> List[SimStep]:
    return [SimStep(kind="atom", name="", fwd_cost=node.cost.fwd_compute_time,
                    bwd_cost=node.cost.bwd_grad_act_time + node.cost.bwd_grad_w_time)]


# ------------------------------------------------------------------------------------------------
# helpers shared by the leaf constructors
# ------------------------------------------------------------------------------------------------
DT = {"fp32": 4, "fp16": 2, "bf16": 2, "fp8": 1, "int32": 4, "int64": 8}


def numel(shape) -> float:
    n = 1
    for s in shape:
        n *= s
    return n if len(shape) else 0


def ishape(shape) -> List[int]:
    return [int(s) for s in shape]


def main_grad_bytes(strategy: StrategyConfig) -> int:
    """fp32 main grads unless reducing in bf16 / not accumulating in fp32 (reference base_struct.py:621-625)."""
    if strategy.grad_reduce_in_bf16 or not strategy.use_fp32_accum_grad:
        return DT["bf16"]
    return DT["fp32"]


def dense_param_memory(strategy: StrategyConfig, weight_numel, weight_bytes_per_el, report_numel) -> MemoryInfo:
    """Weight + main-grad + Adam(fp32 master, m, v) bytes of a dense parameter, ZeRO-sharded over dp*cp."""
    m = MemoryInfo()
    m.weight_numel = report_numel
    m.dense_weight_bytes = weight_numel * weight_bytes_per_el
    m.dense_grad_bytes = weight_numel * main_grad_bytes(strategy)
    m.dense_state_bytes = 3 * DT["fp32"] * weight_numel
    group = strategy.dp_size * strategy.cp_size
    if strategy.zero_state >= 1:
        m.dense_state_bytes /= group
    if strategy.zero_state >= 2:
        m.dense_grad_bytes /= group
    if strategy.zero_state >= 3:
        m.dense_weight_bytes /= group
    return m


def moe_param_memory(strategy: StrategyConfig, weight_numel, weight_bytes_per_el, report_numel) -> MemoryInfo:
    """Same for expert parameters: sharded over the expert-data-parallel group."""
    m = MemoryInfo()
    m.moe_weight_numel = report_numel
    m.moe_weight_bytes = weight_numel * weight_bytes_per_el
    m.moe_grad_bytes = weight_numel * main_grad_bytes(strategy)
    m.moe_state_bytes = 3 * DT["fp32"] * weight_numel
    group = strategy.edp_size
    if strategy.zero_state >= 1:
        m.moe_state_bytes /= group
    if strategy.zero_state >= 2:
        m.moe_grad_bytes /= group
    if strategy.zero_state >= 3:
        m.moe_weight_bytes /= group
    return m


def record_te_dummy_wgrad(mem: MemoryInfo, strategy: StrategyConfig, out_size, in_size, grouped=False) -> None:
    """TE >= 2.3 (grouped: >= 2.10) keeps one dummy wgrad tensor per distinct weight shape when fused
    gradient accumulation is on (reference base_struct.py:1098-1114)."""
    enabled = (strategy.te_grouped_linear_dummy_wgrad_memory_enabled if grouped
               else strategy.te_dummy_wgrad_memory_enabled)
    if not (strategy.use_fused_grad_accumulation and enabled):
        return
    elem = DT.get(strategy.dtype, DT["bf16"])
    mem.te_dummy_wgrad_shapes.add((int(out_size), int(in_size), int(elem)))


def linear_gemm_table(strategy: StrategyConfig, bs, seq, inp, out) -> Dict[str, list]:
    bs, seq, inp, out = int(bs), int(seq), int(inp), int(out)
    wdt = "bf16" if strategy.grad_reduce_in_bf16 else "fp32"
    return dict(B=[bs, bs, 1], M=[seq, seq, out], K=[inp, out, bs * seq], N=[out, inp, inp],
                layout=["TN", "NN", "NT"], accumulate=[False, False, True], out_dtype=["bf16", "bf16", wdt])


def linear_shape_key(table: Dict[str, list]) -> Callable[[str], str]:
    def key(phase: str) -> str:
        i = PHASES.index(phase)
        return (f"b={table['B'][i]}, m={table['M'][i]}, k={table['K'][i]}, n={table['N'][i]}, "
                f"layout={table['layout'][i]}, accumulate={table['accumulate'][i]}, out_dtype={table['out_dtype'][i]}")
    return key

"""Mixture-of-experts block as IR ops (balanced routing, optional pad-to-capacity).

    [shared expert MLP] ; router -> permutation(dispatch a2a [+ETP all-gather]) -> grouped fc1 -> act(*probs)
                          -> grouped fc2 -> un-permutation(combine a2a [+ETP reduce-scatter]) [; + shared]

Numerics follow reference transformer/moe_module.py (Router :20-212, Permutation :214-528, UnPermutation :531-832,
GroupLinearCol/Row :835-1286, ExpertMLP :1370-1559).
"""
from __future__ import annotations

import math
from copy import deepcopy
from typing import List, Tuple

from ..core.config import MLPRecomputeConfig, StrategyConfig, SystemConfig
from ..core.ir import (DT, PHASES, CollectiveSpec, OpNode, SimStep, comm_step, dense_param_memory, ishape, linear_gemm_table,
                       linear_shape_key, moe_param_memory, numel, record_te_dummy_wgrad)
from . import dense as D


def router(st: StrategyConfig, hidden, expert_num, in_shape, name, *, layer_idx=None, has_cached_inputs=False,
           enable_recompute=False, is_last_recompute=False, use_variance_tail_model=False) -> Tuple[OpNode, list]:
    """Gating linear H->E (never tensor-parallel) + softmax; output = scores [B,S,E]."""
    e = DT[st.dtype]
    B, S, Hin = in_shape
    assert hidden == Hin
    n_in = B * S * Hin
    logits = B * S * expert_num
    out_shape = ishape([B, S, expert_num])
    node = OpNode(name=name, kind="Router", display="Router", layer_idx=layer_idx, in_shapes=[ishape(in_shape)],
                  out_shapes=[out_shape], in_bytes=numel(in_shape) * DT["bf16"],
                  weight_shape=[int(hidden), int(expert_num)], enable_recompute=enable_recompute)
    use_var = bool(st.recompute_variance) or use_variance_tail_model
    if is_last_recompute and enable_recompute and use_var:
        node.is_variance_node = True
    w_bytes = hidden * expert_num * e
    x_bytes = n_in * e
    y_bytes = logits * e
    node.act.activation_mem_cache = 0 if has_cached_inputs else n_in * e
    node.act.fwd_peak_mem_no_cache = x_bytes + y_bytes + w_bytes
    node.act.bwd_peak_mem_no_cache = x_bytes + y_bytes + w_bytes
    node.mem = dense_param_memory(st, hidden * expert_num, e, hidden * expert_num)
    base = 2 * n_in * expert_num
    node.compute.fwd_flops = node.compute.bwd_grad_act_flops = node.compute.bwd_grad_w_flops = base
    linear_io = w_bytes + x_bytes + y_bytes
    sm_in = y_bytes * (st.tp_size if (st.enable_sequence_parallel and st.tp_size > 1) else 1)
    node.compute.fwd_accessed_mem = linear_io + 2 * sm_in
    node.compute.bwd_grad_act_accessed_mem = linear_io + 3 * sm_in
    node.compute.bwd_grad_w_accessed_mem = linear_io
    node.ops = {p: "matmul" for p in PHASES}
    s_full = S * st.tp_size if st.enable_sequence_parallel else S
    node.gemm = linear_gemm_table(st, B, s_full, hidden, expert_num)
    node.shape_key = linear_shape_key(node.gemm)
    node.sim_steps = lambda nd, system=None: [D._atom(nd)]
    return node, out_shape


def _split_kernel_time(system: SystemConfig, op_name: str, chunks: List[float]) -> float:
    """Two separate memory-bound kernels: each pays its own launch latency."""
    return sum(system.compute_end2end_time(compute_time=0, mem_time=system.compute_mem_access_time(op_name, b))
               for b in chunks)


def _layout_cost(node: OpNode, system: SystemConfig, chunks: List[float]):
    """Cost override for (un)permutation: sum of per-kernel times instead of one aggregated roofline."""
    def finish(nd: OpNode):
        c = nd.cost
        c.fwd_compute_time = _split_kernel_time(system, "permute_fwd", chunks)
        c.bwd_grad_act_time = _split_kernel_time(system, "permute_bwd", chunks)
        c.bwd_grad_w_time = 0
        c.recompute_compute_time = c.fwd_time if nd.enable_recompute else 0
        for phase, op in (("fwd", "permute_fwd"), ("bwd_grad_act", "permute_bwd")):
            t = c.fwd_compute_time if phase == "fwd" else c.bwd_grad_act_time
            nd.details[phase] = {"compute_details": dict(op_name=op, tflops=None, efficient_factor=None, compute_only_time=0.0),
                                 "io_details": dict(io_time=t, kernels=len(chunks))}
        nd.details["bwd_grad_w"] = {"compute_details": dict(op_name="default", tflops=None, efficient_factor=None, compute_only_time=0.0),
                                    "io_details": dict(io_time=0)}
    return finish


def permutation(st: StrategyConfig, system: SystemConfig, expert_num, topk, pad_to_capacity, capacity, in_shape,
                probs_shape, name, *, layer_idx=None, enable_recompute=False) -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    B, S, Hd = in_shape
    tokens = topk * B * S
    if pad_to_capacity:
        tokens = math.ceil(tokens / expert_num) * expert_num * capacity
    permuted = tokens * Hd
    seq_out = S * st.etp_size if (st.enable_sequence_parallel and st.etp_size > 1) else S
    out_tokens = B * seq_out * topk
    if pad_to_capacity:
        out_tokens = math.ceil(out_tokens / expert_num) * expert_num * capacity
    out_shape = ishape([out_tokens, Hd])
    probs_numel = numel(probs_shape)
    node = OpNode(name=name, kind="Permutation", display="Permutation", layer_idx=layer_idx,
                  in_shapes=[ishape(in_shape), ishape(probs_shape)], out_shapes=[out_shape],
                  in_bytes=numel(in_shape) * DT["bf16"] + probs_numel * DT["int32"], enable_recompute=enable_recompute)
    node.act.activation_mem_cache = probs_numel * 8
    k1 = (B * S * Hd + permuted) * e
    k2 = (permuted + permuted) * e
    node.compute.fwd_accessed_mem = k1 + k2
    node.compute.bwd_grad_act_accessed_mem = k1 + k2
    node.ops = {"fwd": "permute_fwd", "bwd_grad_act": "permute_bwd", "bwd_grad_w": "default"}
    if st.ep_size > 1:
        size = permuted * e
        node.comms.append(CollectiveSpec("all2all", size, st.ep_size, st.ep_net, "Dispatch_FWD_EP", "fwd", group="ep"))
        node.comms.append(CollectiveSpec("all2all", size, st.ep_size, st.ep_net, "Dispatch_BWD_EP", "bwd_grad_act", group="ep"))
        if st.dispatch_probs:
            psize = probs_numel * e
            node.comms.append(CollectiveSpec("all2all", psize, st.ep_size, st.ep_net, "Dispatch_PROB_FWD_EP", "fwd", group="ep"))
            node.comms.append(CollectiveSpec("all2all", psize, st.ep_size, st.ep_net, "Dispatch_PROB_BWD_EP", "bwd_grad_act", group="ep"))
    if st.etp_size > 1:
        size = permuted * e * st.etp_size
        node.comms.append(CollectiveSpec("all_gather", size, st.etp_size, st.etp_net, "Permutation_FWD_ETP", "fwd", group="etp"))
        node.comms.append(CollectiveSpec("reduce_scatter", size, st.etp_size, st.etp_net, "Permutation_BWD_ETP", "bwd_grad_act", group="etp"))
    node.recompute_net_follows_fwd = True
    node.post_evaluate = _layout_cost(node, system, [k1, k2])

    def sim(nd: OpNode, system_=None) -> List[SimStep]:
        def atom(nm, nbytes):
            f = system.compute_end2end_time(0, system.compute_mem_access_time("permute_fwd", nbytes))
            b = system.compute_end2end_time(0, system.compute_mem_access_time("permute_bwd", nbytes))
            return SimStep("atom", name=nm, fwd_cost=f, bwd_cost=b)
        steps = [atom("permute1", k1)]
        if st.ep_size > 1:
            steps.append(comm_step(system, "all2all", permuted * e, st.ep_size, st.ep_net, "all2all", group="ep"))
        if st.etp_size > 1:   # the replay uses the tp group / net here (reference moe_module.py:294-309)
            steps.append(comm_step(system, "all_gather", permuted * e * st.etp_size, st.tp_size, st.tp_net, "all_gather"))
        steps.append(atom("permute2", k2))
        return steps
    node.sim_steps = sim
    return node, out_shape


def unpermutation(st: StrategyConfig, system: SystemConfig, in_shape, ori_shape, name, *, probs_shape=None,
                  layer_idx=None, enable_recompute=False) -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    before = numel(in_shape)
    after = numel(ori_shape)
    out_shape = ishape(ori_shape)
    in_shapes = [ishape(in_shape)] + ([ishape(probs_shape)] if probs_shape is not None else [])
    in_bytes = before * DT["bf16"] + (numel(probs_shape) * DT["int32"] if probs_shape is not None else 0)
    if not st.dispatch_probs:
        assert probs_shape is not None, "dispatch_probs=False requires two inputs in Permutation, [x, probs]"
    node = OpNode(name=name, kind="UnPermutation", display="UnPermutation", layer_idx=layer_idx, in_shapes=in_shapes,
                  out_shapes=[out_shape], in_bytes=in_bytes, enable_recompute=enable_recompute)
    if st.dispatch_probs:
        node.act.activation_mem_cache = 0
        node.act.fwd_peak_mem_no_cache = max(before, after) * e
        node.act.bwd_peak_mem_no_cache = 0
    else:
        node.act.activation_mem_cache = before * e
        node.act.fwd_peak_mem_no_cache = before * e + after * e
        node.act.bwd_peak_mem_no_cache = before * e + after * e
    k1 = 2 * before * e
    k2 = (before + after) * e
    node.compute.fwd_accessed_mem = k1 + k2
    node.compute.bwd_grad_act_accessed_mem = k1 + k2
    node.ops = {"fwd": "permute_fwd", "bwd_grad_act": "permute_bwd", "bwd_grad_w": "default"}
    if st.etp_size > 1:
        size = before * e * st.etp_size
        node.comms.append(CollectiveSpec("reduce_scatter", size, st.etp_size, st.etp_net, "Combine_FWD_ETP", "fwd", group="etp"))
        node.comms.append(CollectiveSpec("all_gather", size, st.etp_size, st.etp_net, "Combine_BWD_ETP", "bwd_grad_act", group="etp"))
    if st.ep_size > 1:
        size = before * e
        node.comms.append(CollectiveSpec("all2all", size, st.ep_size, st.ep_net, "Combine_FWD_EP", "fwd", group="ep"))
        node.comms.append(CollectiveSpec("all2all", size, st.ep_size, st.ep_net, "Combine_BWD_EP", "bwd_grad_act", group="ep"))
    node.recompute_net_follows_fwd = True
    node.post_evaluate = _layout_cost(node, system, [k1, k2])

    def sim(nd: OpNode, system_=None) -> List[SimStep]:
        def atom(nm, nbytes):
            f = system.compute_end2end_time(0, system.compute_mem_access_time("permute_fwd", nbytes))
            b = system.compute_end2end_time(0, system.compute_mem_access_time("permute_bwd", nbytes))
            return SimStep("atom", name=nm, fwd_cost=f, bwd_cost=b)
        steps = [atom("unpermute1", k1)]
        if st.etp_size > 1:
            steps.append(comm_step(system, "reduce_scatter", before * e * st.etp_size, st.tp_size, st.tp_net, "reduce_scatter"))
        if st.ep_size > 1:
            steps.append(comm_step(system, "all2all", before * e, st.ep_size, st.ep_net, "all2all", group="ep"))
        steps.append(atom("unpermutate2_and_combine", k2))
        return steps
    node.sim_steps = sim
    return node, out_shape


def _group_shape_key(st: StrategyConfig, local_experts, tokens_total, out_f, in_f):
    assert tokens_total % local_experts == 0, \
        f"input size {tokens_total} is not divisible by local_expert_num {local_experts} {st.parallelism}"
    m = tokens_total // local_experts
    head = (f"ng={local_experts}, M={m}, N={out_f}, K={in_f}, dtype={'fp8' if st.fp8 else 'bf16'}, out_dtype=bf16, "
            f"main_grad_dtype={'bf16' if st.grad_reduce_in_bf16 else 'fp32'}")
    tails = {
        "fwd": ", stage=fwd, grad=False, accumulate=False, use_split_accumulator=False, single_output=True",
        "bwd_grad_act": ", stage=bwd_grad_act, grad=True, accumulate=False, use_split_accumulator=True, single_output=False",
        "bwd_grad_w": ", stage=bwd_grad_w, grad=True, accumulate=True, use_split_accumulator=True, single_output=False",
    }
    return lambda phase: head + tails[phase]


def group_linear(st: StrategyConfig, kind: str, in_features, out_features, local_experts, in_shape, name, *,
                 layer_idx=None, has_cached_inputs=False, enable_recompute=False, is_last_recompute=False,
                 use_variance_tail_model=False, offload_inputs=False, extra_inputs=()) -> Tuple[OpNode, list]:
    """Grouped (per-expert) linear; ``kind`` = "col" (output split over ETP) or "row" (input split over ETP)."""
    e = DT[st.dtype]
    w_e = a_e = (1 if st.fp8 else e)
    if kind == "col":
        assert out_features % st.etp_size == 0
        in_f, out_f = in_features, out_features // st.etp_size
    else:
        assert in_features % st.etp_size == 0
        in_f, out_f = in_features // st.etp_size, out_features
    tokens, Hin = in_shape
    assert in_f == Hin, f"input_size {in_f} != hidden_size {Hin}"
    out_shape = ishape([tokens, out_f])
    cls = "GroupLinearCol" if kind == "col" else "GroupLinearRow"
    # the input is the quantizer's Float8Tensor under fp8 (1 byte / element): matters when this op opens a recompute segment,
    # i.e. when the quantizer in front of it is not recomputed (cache_groupgemm_col_fp8_inputs)
    in_bytes = numel(in_shape) * (1 if st.fp8 else DT["bf16"]) + sum(numel(s) * b for s, b in extra_inputs)
    node = OpNode(name=name, kind=cls, display=cls, layer_idx=layer_idx,
                  in_shapes=[ishape(in_shape)] + [ishape(s) for s, _ in extra_inputs], out_shapes=[out_shape],
                  in_bytes=in_bytes, weight_shape=[int(out_f), int(in_f)], enable_recompute=enable_recompute,
                  offload_inputs=offload_inputs,
                  repr_args=f"input_size={in_f},output_size={out_f},local_expert_num={local_experts},use_bias=False")
    use_var = bool(st.recompute_variance) or use_variance_tail_model
    if is_last_recompute and enable_recompute and use_var:
        node.is_variance_node = True
    x = tokens * Hin * a_e
    y = tokens * out_f * e
    w = local_experts * in_f * out_f * w_e
    g32 = local_experts * in_f * out_f * 4
    cache = tokens * Hin * a_e
    if has_cached_inputs or (kind == "col" and offload_inputs):
        cache = 0
    node.act.activation_mem_cache = cache
    node.act.fwd_peak_mem_no_cache = x + y + (0 if st.use_accm_weight else w)
    node.act.bwd_peak_mem_no_cache = x + y + (g32 if st.fp8 else 0) + (x if (kind == "col" and offload_inputs) else 0)
    wn = local_experts * in_f * out_f
    node.mem = moe_param_memory(st, wn, w_e, wn * st.ep_size * st.etp_size)
    record_te_dummy_wgrad(node.mem, st, out_f, in_f, grouped=True)
    base = 2 * tokens * in_f * out_f
    node.compute.fwd_flops = node.compute.bwd_grad_act_flops = node.compute.bwd_grad_w_flops = base
    node.compute.fwd_accessed_mem = x + w + y
    node.compute.bwd_grad_act_accessed_mem = w + y + x
    node.compute.bwd_grad_w_accessed_mem = y + x + w + (in_f * out_f * 4 if st.use_fused_grad_accumulation else 0)
    op = "fp8_group_matmul" if st.fp8 else "group_matmul"
    node.ops = {p: op for p in PHASES}
    node.shape_key = _group_shape_key(st, local_experts, int(tokens), int(out_f), int(in_f))
    m = int(tokens) // local_experts
    wdt = "bf16" if st.grad_reduce_in_bf16 else "fp32"
    node.gemm = dict(B=[local_experts] * 3, M=[m, m, int(out_f)], K=[int(in_f), int(out_f), m],
                     N=[int(out_f), int(in_f), int(in_f)], layout=["TN", "NN", "NT"],
                     accumulate=[False, False, True], out_dtype=["bf16", "bf16", wdt])
    node.sim_steps = lambda nd, system=None: [D._atom(nd, "Linear")]
    return node, out_shape


def serial_group_linear(em, attr: str, display: str, in_f: int, out_f: int, local_experts: int, x_shape, layer_idx, enable_recompute):
    """``group_linear_mode='serial'`` (legacy; reference moe_module.py:878-900, 1101-1134): the grouped linear is a container of
    ``local_experts`` plain ``LinearCol`` children -- for the column AND the row expert linear -- each fed a ``[1, tokens / experts, H]``
    slice of the permuted tokens (split / unsqueeze / squeeze / cat are shape-only shims there, so nothing else is emitted).  Costs
    come from the dense ``matmul`` table and the parameters are accounted like dense, data-parallel ones, exactly as the reference
    does in this mode."""
    st = em.st
    tokens, hidden = int(x_shape[0]), int(x_shape[1])
    assert tokens % local_experts == 0, f"tokens {tokens} not divisible by {local_experts} local experts"
    per = tokens // local_experts
    out_local = out_f
    with em.scope(attr, display):
        for i in range(local_experts):
            node, o = D.linear_col(st, in_f, out_f, [1, per, hidden], em.name(f"linear_{i}"), layer_idx=layer_idx,
                                   has_cached_inputs=False, enable_recompute=enable_recompute, display="LinearCol")
            em.add(node)
            out_local = o[-1]
    return [tokens, out_local]


def emit_expert_mlp(em, layer_idx: int, shape, rc: MLPRecomputeConfig):
    from .llm import emit_mlp
    cfg, st, system = em.model, em.st, em.system
    assert cfg.group_linear_mode in ("parallel", "serial"), cfg.group_linear_mode      # reference moe_module.py:854,1078
    serial = cfg.group_linear_mode == "serial"
    if serial and st.fp8:
        raise NotImplementedError("group_linear_mode='serial' is modelled for bf16 only (the reference's serial mode has no fp8 path)")
    local_experts = cfg.expert_num // st.ep_size
    ffn = cfg.moe_ffn_hidden_size if cfg.moe_ffn_hidden_size is not None else cfg.intermediate_size
    fc1_out = 2 * ffn if cfg.use_swiglu else ffn
    moe_ckpt = rc.megatron_moe
    act_ckpt = rc.megatron_moe_act and not moe_ckpt
    ori_shape = list(shape)
    with em.scope("mlp", "MoELayer"):
        has_shared = getattr(cfg, "moe_shared_expert_intermediate_size", None) is not None
        if has_shared:
            shared_rc = deepcopy(rc)
            shared_rc.megatron_layernorm = False
            emit_mlp(em, f"{layer_idx}-shareExpert", list(shape), shared_rc, attr="shared_expert", display="MLP",
                     intermediate_size=cfg.moe_shared_expert_intermediate_size)
        r, probs_shape = router(st, cfg.hidden_size, cfg.expert_num, shape, em.name("router"), layer_idx=layer_idx,
                                has_cached_inputs=rc.megatron_layernorm,
                                enable_recompute=rc.router_recompute or rc.megatron_layernorm or moe_ckpt,
                                is_last_recompute=rc.megatron_layernorm, use_variance_tail_model=rc.megatron_layernorm)
        em.add(r)
        p, x = permutation(st, system, cfg.expert_num, cfg.topk, cfg.moe_pad_expert_input_to_capacity, cfg.capacity,
                           shape, probs_shape, em.name("permutation"), layer_idx=layer_idx,
                           enable_recompute=rc.permutation_recompute or moe_ckpt)
        em.add(p)
        lin_rc = rc.linear_recompute or moe_ckpt
        offload = bool(st.offload_groupgemm_col_inputs)
        if st.fp8:
            with em.scope("group_linear1", "QuantizedGroupLinearCol"):
                q_rc = False if st.cache_groupgemm_col_fp8_inputs else lin_rc
                qn, _ = D.fp8_quantizer(st, x, em.name("quantizer"), enable_recompute=q_rc)
                qn.offload_inputs = offload if not st.cache_groupgemm_col_fp8_inputs else False
                em.add(qn)
                g1, x = group_linear(st, "col", cfg.hidden_size, fc1_out, local_experts, x, em.name("linear"),
                                     layer_idx=layer_idx, enable_recompute=lin_rc,
                                     offload_inputs=offload if st.cache_groupgemm_col_fp8_inputs else False)
                em.add(g1)
        elif serial:
            x = serial_group_linear(em, "group_linear1", "GroupLinearCol", cfg.hidden_size, fc1_out, local_experts, x, layer_idx, lin_rc)
        else:
            g1, x = group_linear(st, "col", cfg.hidden_size, fc1_out, local_experts, x, em.name("group_linear1"),
                                 layer_idx=layer_idx, enable_recompute=lin_rc, offload_inputs=offload)
            em.add(g1)
        act_rc = lin_rc or act_ckpt
        if cfg.use_swiglu:
            a, x2 = D.swiglu(st, x, em.name("expert_activation_layer"), enable_recompute=act_rc,
                             probs_numel=numel(probs_shape) if st.dispatch_probs else None)
        else:
            a, x2 = D.gelu(st, x, em.name("expert_activation_layer"), enable_recompute=act_rc)
        if st.dispatch_probs:
            a.in_shapes.append(ishape(probs_shape))
            a.in_bytes += numel(probs_shape) * DT["int32"]
        em.add(a)
        x = x2
        if st.fp8:
            with em.scope("group_linear2", "QuantizedGroupLinearRow"):
                qn, _ = D.fp8_quantizer(st, x, em.name("quantizer"), enable_recompute=act_rc)
                em.add(qn)
                g2, x = group_linear(st, "row", ffn, cfg.hidden_size, local_experts, x, em.name("linear"),
                                     layer_idx=layer_idx, has_cached_inputs=act_ckpt, enable_recompute=act_rc,
                                     is_last_recompute=True, use_variance_tail_model=act_ckpt)
                em.add(g2)
        elif serial:
            x = serial_group_linear(em, "group_linear2", "GroupLinearRow", ffn, cfg.hidden_size, local_experts, x, layer_idx, act_rc)
        else:
            g2, x = group_linear(st, "row", ffn, cfg.hidden_size, local_experts, x, em.name("group_linear2"),
                                 layer_idx=layer_idx, has_cached_inputs=act_ckpt, enable_recompute=act_rc,
                                 is_last_recompute=True, use_variance_tail_model=act_ckpt)
            em.add(g2)
        u, out = unpermutation(st, system, x, ori_shape, em.name("unpermutation"),
                               probs_shape=None if st.dispatch_probs else probs_shape, layer_idx=layer_idx,
                               enable_recompute=rc.permutation_recompute or moe_ckpt)
        em.add(u)
        if st.recompute_granularity == "selective_recompute" and rc.megatron_layernorm:
            r.is_breakpoint = True
        if u.enable_recompute and st.recompute_granularity == "selective_recompute":
            u.is_breakpoint = True
        if has_shared:
            em.add(D.passthrough("AddModule", [out, out], [out], em.name("SharedExpertAddFunction"),
                                 enable_recompute=False, display="SharedExpertAddFunction"))
    return out

"""Dense transformer operators as IR nodes: closed-form flops / bytes / activations / collectives.

Each constructor takes the *local* input shape and returns ``(node, output_shape)``.  Notation:
``B`` micro-batch, ``S`` local sequence, ``H`` hidden, ``e`` bytes of the strategy dtype, tp/cp sizes,
SP = sequence parallel.  Numerics follow the reference leaf by leaf (transformer/dense_module.py); every
docstring cites the lines it mirrors.
"""
from __future__ import annotations

from typing import List, Tuple

from ..core.config import StrategyConfig, SystemConfig
from ..core.ir import (DT, CollectiveSpec, OpNode, SimStep, comm_step, dense_param_memory, ishape, linear_gemm_table,
                       linear_shape_key, numel, record_te_dummy_wgrad)


def linear_op_name(st) -> str:
    """Efficiency table a linear's three GEMMs are priced from: bf16, per-tensor fp8, or block-scaled MXFP8
    (``strategy.fp8_recipe = "mxfp8"``: tcgen05 ``kind::mxf8f6f4.block_scale``, csrc/gemm_mxfp8.cu)."""
    if not st.fp8:
        return "matmul"
    return "mxfp8_matmul" if getattr(st, "fp8_recipe", "tensor") == "mxfp8" else "fp8_matmul"


def _atom(node: OpNode, name: str = "") -> SimStep:
    return SimStep(kind="atom", name=name, fwd_cost=node.cost.fwd_compute_time,
                   bwd_cost=node.cost.bwd_grad_act_time + node.cost.bwd_grad_w_time)


def _tp_comm(strategy, op, size, stage, phase, exposed: float = 1.0) -> CollectiveSpec:
    return CollectiveSpec(op=op, size=size, comm_num=strategy.tp_size, net=strategy.tp_net, stage=stage, phase=phase,
                          group="tp", exposed=exposed)


def _fused_exposure(st: StrategyConfig):
    """(all-gather, reduce-scatter, wgrad re-gather) exposure factors of the TP collectives adjacent to a GEMM."""
    if getattr(st, "tp_overlap_model", False) and st.enable_sequence_parallel:
        def pick(v, default):            # a float, or a {tp_size: exposure} table (calib/fused.py)
            if isinstance(v, dict):
                v = v.get(str(st.tp_size), v.get(st.tp_size, default))
            return float(v)
        return pick(st.tp_overlap_exposed_ag, 0.6), pick(st.tp_overlap_exposed_rs, 0.4), 0.0
    return 1.0, 1.0, 1.0


# ------------------------------------------------------------------------------------------------
# Embedding  (dense_module.py:18-193)
# ------------------------------------------------------------------------------------------------
def embedding(st: StrategyConfig, hidden, vocab, in_shape, name, display="LanguageModelEmbedding_0") -> Tuple[OpNode, list]:
    assert vocab % st.tp_size == 0
    assert len(in_shape) == 2, "input size should equal 2"
    e = DT[st.dtype]
    v_local = vocab // st.tp_size
    B, S = in_shape
    s_out = S / st.tp_size if st.enable_sequence_parallel else S
    out_shape = ishape([B, s_out, hidden])
    out_numel = numel(out_shape)
    ids_bytes = B * S * 4
    w_bytes = v_local * hidden * e
    y_bytes = B * S * hidden * e
    node = OpNode(name=name, kind="Embedding", display=display, in_shapes=[ishape(in_shape)], out_shapes=[out_shape],
                  in_bytes=numel(in_shape) * DT["bf16"], repr_args=f"hidden_size={hidden},vocab_size={v_local}")
    node.mem = dense_param_memory(st, v_local * hidden, e, v_local * hidden * st.tp_size)
    node.act.fwd_peak_mem_no_cache = ids_bytes + y_bytes + (0 if st.use_accm_weight else w_bytes)
    node.act.bwd_peak_mem_no_cache = w_bytes
    node.compute.fwd_accessed_mem = ids_bytes + w_bytes + y_bytes
    node.compute.bwd_grad_w_accessed_mem = 2 * (v_local * hidden * 4)
    if getattr(st, "sparse_embedding_model", False):
        # this repo's runtime (csrc/elementwise.cu embedding_fwd / embedding_bwd kernels): a row gather and a red.global.add scatter
        # into the fp32 main grad touch B * S table rows, not the whole table (step profile: 0.03 / 0.08 ms instead of 0.4 / 1.6 ms)
        node.compute.fwd_accessed_mem = ids_bytes + 2 * y_bytes
        node.compute.bwd_grad_w_accessed_mem = ids_bytes + y_bytes + 2 * (B * S * hidden * 4)
    if st.tp_size > 1:
        node.comms.append(_tp_comm(st, "all_reduce", out_numel * e, "Embedding", "fwd"))
        if st.enable_sequence_parallel:
            node.comms.append(_tp_comm(st, "all_gather", out_numel * e, "Embedding", "bwd_grad_w"))

    def sim(n: OpNode, system) -> List[SimStep]:
        steps = [_atom(n)]
        if st.tp_size > 1:
            size = out_numel * e
            if st.enable_sequence_parallel:
                steps.append(comm_step(system, "reduce_scatter", size, st.tp_size, st.tp_net, "reduce_scatter",
                                       stage="Embedding"))
            else:
                steps.append(comm_step(system, "all_reduce", size, st.tp_size, st.tp_net, "all_reduce", bwd=False,
                                       stage="Embedding"))
        return steps
    node.sim_steps = sim
    return node, out_shape


# ------------------------------------------------------------------------------------------------
# Column / row parallel linear  (dense_module.py:195-781)
# ------------------------------------------------------------------------------------------------
def linear_col(st: StrategyConfig, in_features, out_features, in_shape, name, *, layer_idx=None,
               has_cached_inputs=False, enable_recompute=False, enable_fp8=True, is_last_recompute=False,
               use_variance_tail_model=False, disable_tensor_parallel=False,
               display="ColumnParallelLinear") -> Tuple[OpNode, list]:
    assert out_features % st.tp_size == 0
    e = DT[st.dtype]
    out_local = out_features if disable_tensor_parallel else out_features // st.tp_size
    fp8 = bool(st.fp8 and enable_fp8)
    w_e = a_e = (1 if fp8 else e)
    B, S, Hin = in_shape
    assert in_features == Hin, f"input_size: {in_features} vs hidden_size: {Hin}"
    s_full = S * st.tp_size if st.enable_sequence_parallel else S
    tokens_x_in = B * s_full * Hin                      # "micro_hidden_state_size"
    out_shape = ishape([B, s_full, out_local])
    out_numel = B * out_shape[1] * out_local
    node = OpNode(name=name, kind="LinearCol", display=display, layer_idx=layer_idx,
                  in_shapes=[ishape(in_shape)], out_shapes=[out_shape], in_bytes=numel(in_shape) * (1 if fp8 else DT["bf16"]),
                  weight_shape=[int(out_local), int(in_features)], enable_recompute=enable_recompute,
                  repr_args=f"input_size={in_features},output_size={out_local},use_bias=False,enable_recompute={enable_recompute}, "
                            f"TP={st.tp_size}")
    use_var = bool(st.recompute_variance) or use_variance_tail_model
    if is_last_recompute and enable_recompute and use_var:
        node.is_variance_node = True
    w_bytes = in_features * out_local * w_e
    x_bytes = tokens_x_in * a_e
    y_bytes = out_numel * e
    # activation: the (pre-all-gather) input is what gets cached; SP caches only the local 1/tp slice (bf16 only)
    cache = tokens_x_in * a_e
    x_ag, x_rs, x_regather = _fused_exposure(st)
    if st.enable_sequence_parallel and not st.fp8 and x_regather > 0:
        cache /= st.tp_size            # (fused runtime: the gathered input itself is saved, nothing is re-gathered)
    if has_cached_inputs:
        cache = 0
    node.act.activation_mem_cache = cache
    peak = x_bytes + y_bytes + (0 if st.use_accm_weight else w_bytes)
    node.act.fwd_peak_mem_no_cache = peak
    node.act.bwd_peak_mem_no_cache = peak
    wn = in_features * out_local
    node.mem = dense_param_memory(st, wn, w_e, wn * st.tp_size)
    record_te_dummy_wgrad(node.mem, st, out_local, in_features)
    base = 2 * tokens_x_in * out_local
    node.compute.fwd_flops = node.compute.bwd_grad_act_flops = node.compute.bwd_grad_w_flops = base
    node.compute.fwd_accessed_mem = x_bytes + w_bytes + y_bytes
    node.compute.bwd_grad_act_accessed_mem = w_bytes + y_bytes + x_bytes
    main_grad = in_features * out_local * 4
    node.compute.bwd_grad_w_accessed_mem = y_bytes + x_bytes + w_bytes + (main_grad if st.use_fused_grad_accumulation else 0)
    op = linear_op_name(st)
    node.ops = {"fwd": op, "bwd_grad_act": op, "bwd_grad_w": op}
    node.gemm = linear_gemm_table(st, B, s_full, in_features, out_local)
    node.shape_key = linear_shape_key(node.gemm)
    node.recompute_net_follows_fwd = True
    size = tokens_x_in * e
    if st.tp_size > 1:
        if st.enable_sequence_parallel:
            node.comms.append(_tp_comm(st, "all_gather", size, "LinearCol_FWD_SP", "fwd", x_ag))
            node.comms.append(_tp_comm(st, "reduce_scatter", size, "LinearCol_BWD_ACT_SP", "bwd_grad_act", x_rs))
            node.comms.append(_tp_comm(st, "all_gather", size, "LinearCol_BWD_W_SP", "bwd_grad_w", x_regather))
        else:
            node.comms.append(_tp_comm(st, "all_reduce", size, "LinearCol_BWD_ACT_TP", "bwd_grad_act"))

    def sim(n: OpNode, system) -> List[SimStep]:
        steps = []
        if st.tp_size > 1 and st.enable_sequence_parallel:
            steps.append(comm_step(system, "all_gather", size, st.tp_size, st.tp_net, "all_gather", exposed=x_ag, exposed_bwd=x_rs))
        elif st.tp_size > 1:
            steps.append(comm_step(system, "all_reduce", size, st.tp_size, st.tp_net, "all_reduce", fwd=False))
        steps.append(_atom(n, "Linear"))
        if st.tp_size > 1 and st.enable_sequence_parallel:
            steps.append(comm_step(system, "all_gather", size, st.tp_size, st.tp_net, "all_gather_bwd", fwd=False, exposed=x_regather))
        return steps
    node.sim_steps = sim
    return node, out_shape


def linear_row(st: StrategyConfig, in_features, out_features, in_shape, name, *, layer_idx=None,
               has_cached_inputs=False, enable_recompute=False, is_last_recompute=False,
               use_variance_tail_model=False, display="RowParallelLinear") -> Tuple[OpNode, list]:
    assert in_features % st.tp_size == 0
    e = DT[st.dtype]
    in_local = in_features // st.tp_size
    w_e = a_e = (1 if st.fp8 else e)
    B, S, Hin = in_shape
    assert in_local == Hin, f"input_size: {in_local} vs hidden_size: {Hin}"
    s_out = S / st.tp_size if st.enable_sequence_parallel else S
    out_shape = ishape([B, s_out, out_features])
    tokens_x_in = B * S * Hin
    out_numel_full = B * out_shape[1] * out_features * (st.tp_size if st.enable_sequence_parallel else 1)
    node = OpNode(name=name, kind="LinearRow", display=display, layer_idx=layer_idx,
                  in_shapes=[ishape(in_shape)], out_shapes=[out_shape], in_bytes=numel(in_shape) * (1 if st.fp8 else DT["bf16"]),
                  weight_shape=[int(out_features), int(in_local)], enable_recompute=enable_recompute,
                  repr_args=f"input_size={in_local},output_size={out_features},use_bias=False,enable_recompute={enable_recompute}, "
                            f"TP={st.tp_size}")
    use_var = bool(st.recompute_variance) or use_variance_tail_model
    if is_last_recompute and enable_recompute and use_var:
        node.is_variance_node = True
    w_bytes = in_local * out_features * w_e
    x_bytes = tokens_x_in * a_e
    y_bytes = out_numel_full * e
    node.act.activation_mem_cache = 0 if has_cached_inputs else tokens_x_in * a_e
    peak = x_bytes + y_bytes + (0 if st.use_accm_weight else w_bytes)
    node.act.fwd_peak_mem_no_cache = peak
    node.act.bwd_peak_mem_no_cache = peak
    wn = in_local * out_features
    node.mem = dense_param_memory(st, wn, w_e, wn * st.tp_size)
    record_te_dummy_wgrad(node.mem, st, out_features, in_local)
    base = 2 * tokens_x_in * out_features
    node.compute.fwd_flops = node.compute.bwd_grad_act_flops = node.compute.bwd_grad_w_flops = base
    node.compute.fwd_accessed_mem = x_bytes + w_bytes + y_bytes
    node.compute.bwd_grad_act_accessed_mem = w_bytes + y_bytes + x_bytes
    main_grad = in_local * out_features * 4
    node.compute.bwd_grad_w_accessed_mem = y_bytes + x_bytes + (main_grad if st.use_fused_grad_accumulation else 0)
    op = linear_op_name(st)
    node.ops = {"fwd": op, "bwd_grad_act": op, "bwd_grad_w": op}
    node.gemm = linear_gemm_table(st, B, S, in_local, out_features)
    node.shape_key = linear_shape_key(node.gemm)
    node.recompute_net_follows_fwd = True
    size = out_numel_full * e
    if st.tp_size > 1:
        if st.enable_sequence_parallel:
            x_ag, x_rs, _ = _fused_exposure(st)
            node.comms.append(_tp_comm(st, "reduce_scatter", size, "LinearROW_FWD_SP", "fwd", x_rs))
            node.comms.append(_tp_comm(st, "all_gather", size, "LinearROW_BWD_SP", "bwd_grad_act", x_ag))
        else:
            node.comms.append(_tp_comm(st, "all_reduce", size, "LinearROW_FWD_TP", "fwd"))

    def sim(n: OpNode, system) -> List[SimStep]:
        steps = [_atom(n, "Linear")]
        if st.tp_size > 1 and st.enable_sequence_parallel:
            steps.append(comm_step(system, "reduce_scatter", size, st.tp_size, st.tp_net, "reduce_scatter",
                                   exposed=_fused_exposure(st)[1], exposed_bwd=_fused_exposure(st)[0]))
        elif st.tp_size > 1:
            steps.append(comm_step(system, "all_reduce", size, st.tp_size, st.tp_net, "all_reduce", bwd=False))
        return steps
    node.sim_steps = sim
    return node, out_shape


# ------------------------------------------------------------------------------------------------
# RMSNorm  (dense_module.py:784-994)
# ------------------------------------------------------------------------------------------------
def layernorm(st: StrategyConfig, norm_size, in_shape, name, *, has_cached_inputs=False, enable_recompute=False,
              use_fused_norm=None, display="LayerNorm") -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    fused = st.use_fused_norm if use_fused_norm is None else use_fused_norm
    B, S, Hn = in_shape
    assert norm_size == Hn
    n = B * S * Hn
    out_shape = ishape(in_shape)
    node = OpNode(name=name, kind="LayerNorm", display=display, in_shapes=[ishape(in_shape)], out_shapes=[out_shape],
                  in_bytes=numel(in_shape) * DT["bf16"], weight_shape=[int(norm_size)], enable_recompute=enable_recompute,
                  repr_args=f"norm_size={norm_size},norm_type=rms_norm,use_fused_norm={st.use_fused_norm},"
                            f"enable_recompute={enable_recompute}")
    x_bytes = y_bytes = n * e
    g_bytes = norm_size * e
    rstd = n / norm_size * e
    if fused:
        node.act.activation_mem_cache = 0 if has_cached_inputs else n * e
        node.act.fwd_peak_mem_no_cache = x_bytes + y_bytes
        node.act.bwd_peak_mem_no_cache = x_bytes + y_bytes + rstd
        node.compute.fwd_accessed_mem = x_bytes + g_bytes + y_bytes
        node.compute.bwd_grad_w_accessed_mem = x_bytes + g_bytes + g_bytes
        node.compute.bwd_grad_act_accessed_mem = x_bytes + g_bytes + y_bytes + rstd
    else:
        x32 = n * 4
        rstd32 = n / norm_size * 4
        node.act.activation_mem_cache = x32 + rstd32 + (x32 + rstd32) + y_bytes
        node.act.fwd_peak_mem_no_cache = x32 + rstd32 + x32 + x32 + rstd32
        node.act.bwd_peak_mem_no_cache = node.act.fwd_peak_mem_no_cache
        node.compute.fwd_accessed_mem = (x_bytes + x32) + (x32 + y_bytes) + (4 * x32 + 4 * rstd + y_bytes + g_bytes)
        node.compute.bwd_grad_w_accessed_mem = 2 * y_bytes + g_bytes
        node.compute.bwd_grad_act_accessed_mem = (y_bytes + x32) + (x_bytes + x32) + (11 * x32 + 5 * rstd + x_bytes + g_bytes)
    node.mem = dense_param_memory(st, norm_size, e, norm_size)
    if getattr(st, "norm_bwd_model", False) and fused:
        # this repo's RMSNorm backward (register-limited, one CTA per SM, + a dgamma reduce kernel) streams at ~1.8 TB/s with ~11 us of
        # fixed cost, not at the "default" elementwise bandwidth: priced from the system config's "norm_bwd" entry when it has one
        node.ops = {"fwd": "default", "bwd_grad_act": "norm_bwd", "bwd_grad_w": "norm_bwd"}
    node.sim_steps = lambda nd, system=None: [_atom(nd)]
    return node, out_shape


# ------------------------------------------------------------------------------------------------
# Scaled dot-product attention incl. Ulysses-style context parallel  (dense_module.py:1061-1804)
# ------------------------------------------------------------------------------------------------
def _cp_stage_specs(st, q_mem, k_mem, v_mem, o_mem):
    bwd_pre = [("Attention_BWD_CP2_DOUT", o_mem)]
    if not st.te_cp_a2a_saves_pre_posta2a_output:
        bwd_pre.insert(0, ("Attention_BWD_CP2_OUT", o_mem))
    return {
        "fwd_pre": [("Attention_FWD_CP1_Q", q_mem), ("Attention_FWD_CP1_K", k_mem), ("Attention_FWD_CP1_V", v_mem)],
        "fwd_post": [("Attention_FWD_CP2", o_mem)],
        "bwd_pre": bwd_pre,
        "bwd_post": [("Attention_BWD_CP1_DQ", q_mem), ("Attention_BWD_CP1_DK", k_mem), ("Attention_BWD_CP1_DV", v_mem)],
    }


def _cp_a2a_peaks(st, q, k, v, o, saved_o, independent):
    """Stage-local live sets of PreA2A -> FA -> PostA2A in fwd and bwd (dense_module.py:1259-1338)."""
    sync = st.cp_a2a_mode == "sync_cp"

    def three(a, b, c):
        tot = a + b + c
        return 2 * tot + (b + c) + (a + b) if sync else 4 * tot

    def two(a, b):
        tot = a + b
        return 3 * tot + max(a, b) if sync else 4 * tot

    qkv = q + k + v
    fwd_pre = three(q, k, v)
    fwd_fa = 3 * qkv + o
    fwd_post = 2 * qkv + 4 * o
    if independent:
        bwd_pre = saved_o + 4 * o
        base = saved_o + 2 * o
    else:
        bwd_pre = two(o, o)
        base = 4 * o
    bwd_fa = max(qkv + base, 2 * qkv + base + q + k, qkv + base)
    bwd_post = base + three(q, k, v)
    return max(fwd_pre, fwd_fa, fwd_post), max(bwd_pre, bwd_fa, bwd_post)


def core_attention(st: StrategyConfig, system: SystemConfig, head_size, head_num, kv_head_num, in_shape, name, *,
                   v_head_dim=None, mla=False, has_cached_inputs=False, enable_recompute=False,
                   is_last_recompute=False, use_variance_tail_model=False, cache_outputs=False,
                   display="DotProductAttention") -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    if st.tp_size > 1:
        assert head_num % st.tp_size == 0 and kv_head_num % st.tp_size == 0
        head_num = head_num / st.tp_size
        kv_head_num = kv_head_num / st.tp_size
    vd = head_size if v_head_dim is None else v_head_dim
    B, S, Hin = in_shape
    if mla:
        assert head_size * (kv_head_num + head_num) + kv_head_num * vd == Hin, \
            f"{head_size * (kv_head_num + head_num) + kv_head_num * vd} vs {Hin}"
    else:
        assert head_size * (2 * kv_head_num + head_num) == Hin
    out_shape = ishape([B, S, head_num * vd])
    node = OpNode(name=name, kind="MLACoreAttention" if mla else "CoreAttention", display=display,
                  in_shapes=[ishape(in_shape)], out_shapes=[out_shape], in_bytes=numel(in_shape) * DT["bf16"],
                  enable_recompute=enable_recompute,
                  repr_args=f"head_size={head_size},head_num={head_num},kv_head_num={kv_head_num},use_math_sdp={st.use_math_sdp},"
                            f"use_flash_sdp={st.use_flash_sdp},enable_recompute={enable_recompute}")
    use_var = bool(st.recompute_variance) or use_variance_tail_model
    if is_last_recompute and enable_recompute and use_var:
        node.is_variance_node = True
    cp_a2a = st.cp_size > 1 and st.cp_comm_type == "a2a"
    independent = cp_a2a and st.te_cp_a2a_saves_pre_posta2a_output

    # ---- activation sizes
    q_n = B * head_num * S * head_size
    if mla:
        k_n = q_n
        v_n = B * head_num * S * vd
        o_n = B * S * vd * head_num
    else:
        k_n = B * kv_head_num * S * head_size
        v_n = B * kv_head_num * S * head_size
        o_n = B * S * head_size * head_num
    lse_n = B * head_num * S
    qkv_mem = (q_n + k_n + v_n) * e
    if st.use_flash_sdp:
        lse_mem, o_mem = lse_n * e, o_n * e
        saved_o = o_mem if (cache_outputs or independent) else 0
        cache = qkv_mem + lse_mem + saved_o
        if has_cached_inputs:
            cache -= qkv_mem
        node.act.activation_mem_cache = cache
        node.act.fwd_peak_mem_no_cache = qkv_mem + lse_mem + o_mem
        node.act.bwd_peak_mem_no_cache = (2 * q_n + 2 * k_n + 2 * v_n + lse_n + o_n) * e - saved_o
        if cp_a2a:
            fwd_pk, bwd_pk = _cp_a2a_peaks(st, q_n * e, k_n * e, v_n * e, o_mem, saved_o, independent)
            node.act.fwd_peak_mem_no_cache = max(fwd_pk, qkv_mem + cache)
            node.act.bwd_peak_mem_no_cache = bwd_pk - saved_o
        elif st.cp_size > 1 and st.cp_comm_type == "all_gather" and not mla:
            kv_mem = (k_n + v_n) * e
            node.act.fwd_peak_mem_no_cache += kv_mem * (st.cp_size - 1)
            node.act.bwd_peak_mem_no_cache += 2 * kv_mem * (st.cp_size - 1)
        elif st.cp_size > 1 and st.cp_comm_type == "p2p" and not mla:
            # ring attention (extension): the K/V block being used plus the one in flight; backward also circulates dK/dV
            kv_mem = (k_n + v_n) * e
            node.act.fwd_peak_mem_no_cache += kv_mem
            node.act.bwd_peak_mem_no_cache += 3 * kv_mem
    else:
        sm_n = B * head_num * S * S
        cache = (q_n + k_n + v_n + sm_n) * e
        if has_cached_inputs and head_num == kv_head_num:
            cache -= (q_n + k_n + v_n) * e
        node.act.activation_mem_cache = cache
        node.act.fwd_peak_mem_no_cache = 2 * sm_n * e
        factor = 3
        if system.accelerator.backend == "cuda":
            factor += 1
        elif st.use_math_sdp and system.accelerator.backend == "musa":
            factor -= 1
        node.act.bwd_peak_mem_no_cache = factor * sm_n * e

    # ---- flops: after the all-to-all every rank sees the full sequence with heads/cp
    seq, heads = S, head_num
    seq_kv = None
    if st.cp_size > 1:
        if st.cp_comm_type == "a2a" or mla:
            assert head_num % st.cp_size == 0, f"head_num {head_num} must be divisible by cp_size {st.cp_size}"
            seq = S * st.cp_size
            heads = head_num // st.cp_size
        elif st.cp_comm_type in ("all_gather", "p2p"):
            # extension (the reference raises here, dense_module.py:1521-1522): all heads, local queries against the whole
            # sequence of keys -- S x (S * cp) scores per head, the same work per rank as the all-to-all flavour
            seq_kv = S * st.cp_size
        else:
            raise NotImplementedError(f"cp_comm_type {st.cp_comm_type} not implemented yet.")
    sq_skv = seq * (seq if seq_kv is None else seq_kv)
    if mla:
        base = B * heads * head_size * sq_skv + B * heads * vd * sq_skv
    else:
        base = 2 * B * heads * head_size * sq_skv
    base *= 1 - st.attention_sparse_ratio
    node.compute.fwd_flops = 2 * base
    node.compute.bwd_grad_act_flops = 4 * base + (base if st.use_flash_sdp else 0)

    # ---- bytes (k and v counted at q size; MLA: v at its own width)
    kq = q_n
    vq = (B * head_num * S * vd) if mla else q_n
    oq = B * S * head_num * vd if mla else B * S * head_num * head_size
    if st.use_flash_sdp:
        node.compute.fwd_accessed_mem = (q_n + kq + vq + oq + lse_n) * e
        node.compute.bwd_grad_act_accessed_mem = (2 * q_n + 2 * kq + 2 * vq + oq + lse_n) * e
    else:
        sm_n = B * head_num * S * S
        node.compute.fwd_accessed_mem = ((q_n + kq + sm_n) + 2 * sm_n + (sm_n + vq + oq)) * e
        node.compute.bwd_grad_act_accessed_mem = (2 * (sm_n + vq + oq) + 2 * sm_n + 2 * (q_n + kq + sm_n)) * e
    node.ops = {"fwd": "sdp_fwd", "bwd_grad_act": "sdp_bwd", "bwd_grad_w": "sdp_bwd"}

    key_heads = head_num // st.cp_size if st.cp_size > 1 else head_num
    key_kv = kv_head_num // st.cp_size if st.cp_size > 1 else kv_head_num
    key_seq = S * st.cp_size if st.cp_size > 1 else S
    contiguous = False if "s5000" in system.sys_name else True
    key = (f"batch={int(B)}, seq_len={int(key_seq)}, head_num={int(key_heads)}, kv_head_num={int(key_kv)}, "
           f"qk_head_dim={int(head_size)}, v_head_dim={int(vd)}, qkv_contiguous={contiguous}")
    node.shape_key = lambda phase: key

    # ---- context-parallel collectives
    specs = None
    if st.cp_size > 1:
        q_c = B * head_num * S * head_size
        k_c = B * kv_head_num * S * head_size
        v_c = B * kv_head_num * S * vd
        o_c = B * head_num * S * vd
        if st.cp_comm_type == "a2a":
            specs = _cp_stage_specs(st, q_c * e, k_c * e, v_c * e, o_c * e)
            for phase, groups in (("fwd", specs["fwd_pre"] + specs["fwd_post"]),
                                  ("bwd_grad_act", specs["bwd_post"] + specs["bwd_pre"])):
                for stage_name, size in groups:
                    node.comms.append(CollectiveSpec("all2all", size, st.cp_size, st.cp_net, stage_name, phase, group="cp"))
        elif st.cp_comm_type == "all_gather":
            kv_mem = (k_c + v_c) * e
            # bytes of the gathered K + V (the reference multiplies by the element size once more, dense_module.py:1375-1378;
            # its all_gather flavour never runs to the end, so there is no number to stay compatible with)
            size = kv_mem * st.cp_size
            node.comms.append(CollectiveSpec("all_gather", size, st.cp_size, st.cp_net, "Attention_FWD_CP", "fwd", group="cp"))
            node.comms.append(CollectiveSpec("all_gather", size, st.cp_size, st.cp_net, "Attention_BWD_CP1", "bwd_grad_act", group="cp"))
            node.comms.append(CollectiveSpec("reduce_scatter", size, st.cp_size, st.cp_net, "Attention_BWD_CP2", "bwd_grad_act", group="cp"))
        elif st.cp_comm_type == "p2p":
            # ring attention (extension): cp - 1 hops of the local K/V block forward, K/V and dK/dV backward; each hop is a
            # neighbour send/recv that Megatron / TE overlap with the attention of the previous block
            kv_block = (k_c + v_c) * e
            hide = float(getattr(st, "cp_p2p_exposed", 1.0))
            node.comms.append(CollectiveSpec("p2p", kv_block * (st.cp_size - 1), 2, st.cp_net, "Attention_FWD_CP_RING", "fwd",
                                             group="cp", exposed=hide))
            node.comms.append(CollectiveSpec("p2p", 2 * kv_block * (st.cp_size - 1), 2, st.cp_net, "Attention_BWD_CP_RING",
                                             "bwd_grad_act", group="cp", exposed=hide))
        else:
            raise NotImplementedError(f"cp_comm_type {st.cp_comm_type} not implemented yet.")
    node.cp_specs = specs

    def sim(n: OpNode, system) -> List[SimStep]:
        steps = []
        if specs is not None:
            def a2a(stage_name, size, fwd):
                return comm_step(system, "all2all", size, st.cp_size, st.cp_net, "all2all", fwd=fwd, bwd=not fwd,
                                 group="cp", stage=stage_name)
            steps += [a2a(nm, sz, True) for nm, sz in specs["fwd_pre"]]
            steps += [a2a(nm, sz, False) for nm, sz in reversed(specs["bwd_post"])]
            steps += [a2a(nm, sz, True) for nm, sz in specs["fwd_post"]]
            steps += [a2a(nm, sz, False) for nm, sz in reversed(specs["bwd_pre"])]
        elif st.cp_size > 1:
            # gathered-K/V and ring flavours (extensions): replay the evaluated collectives over the cp group
            for spec in n.comms:
                fwd = spec.phase == "fwd"
                label = "ring_p2p" if spec.op == "p2p" else spec.op
                steps.append(SimStep(kind="comm", name=label, fwd_cost=spec.time if fwd else 0.0, bwd_cost=0.0 if fwd else spec.time,
                                     comm_op=label, group="cp", comm_num=st.cp_size, stage=spec.stage))
        steps.append(_atom(n, "AttentionScore"))
        return steps
    node.sim_steps = sim
    return node, out_shape


# ------------------------------------------------------------------------------------------------
# zero-cost layout ops (RotaryEmbedding, Cat / Concat / Split / Add)
# ------------------------------------------------------------------------------------------------
def passthrough(kind, in_shapes, out_shapes, name, *, enable_recompute=False, display=None, dim=-1) -> OpNode:
    args = {"SplitModule": f"split_dim={dim}, ", "ConcatModule": f"concat_dim={dim}, "}.get(kind, "")
    node = OpNode(name=name, kind=kind, display=display or kind, in_shapes=[ishape(s) for s in in_shapes],
                  out_shapes=[ishape(s) for s in out_shapes],
                  in_bytes=sum(numel(s) for s in in_shapes) * DT["bf16"], enable_recompute=enable_recompute,
                  repr_args=f"{args}enable_recompute={enable_recompute}")
    node.sim_steps = lambda nd, system=None: []
    return node


# ------------------------------------------------------------------------------------------------
# activations  (dense_module.py:1874-2095)
# ------------------------------------------------------------------------------------------------
def swiglu(st: StrategyConfig, in_shape, name, *, is_fused=None, has_cached_inputs=False, enable_recompute=False,
           probs_numel=None, display="Swiglu") -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    fused = st.use_fused_swiglu if is_fused is None else is_fused
    assert in_shape[-1] % 2 == 0, "hidden size should be even"
    out_shape = ishape(list(in_shape[:-1]) + [in_shape[-1] // 2])
    x = numel(in_shape) * e
    y = numel(out_shape) * e
    in_shapes = [ishape(in_shape)]
    in_bytes = numel(in_shape) * DT["bf16"]
    node = OpNode(name=name, kind="Swiglu", display=display, in_shapes=in_shapes, out_shapes=[out_shape],
                  in_bytes=in_bytes, enable_recompute=enable_recompute,
                  repr_args=f"is_fused={fused},enable_recompute={enable_recompute}")
    cache = (2 * y if fused else 3 * y)
    if has_cached_inputs:
        cache -= 2 * y
    node.act.activation_mem_cache = cache
    node.act.fwd_peak_mem_no_cache = x + y
    node.act.bwd_peak_mem_no_cache = x + y
    if fused:
        node.compute.fwd_accessed_mem = x + y
        node.compute.bwd_grad_act_accessed_mem = x + y
    else:
        node.compute.fwd_accessed_mem = 2 * y + 3 * y
        node.compute.bwd_grad_act_accessed_mem = 2 * y + 2 * 3 * y
    if probs_numel is not None:          # weighted SiLU: routing probs folded into the activation
        node.act.fwd_peak_mem_no_cache += probs_numel * 8
        node.act.bwd_peak_mem_no_cache += probs_numel * 8
        node.compute.fwd_accessed_mem += probs_numel * e
        node.compute.bwd_grad_act_accessed_mem += probs_numel * e
    node.sim_steps = lambda nd, system=None: [_atom(nd)]
    return node, out_shape


def gelu(st: StrategyConfig, in_shape, name, *, has_cached_inputs=False, enable_recompute=False,
         display="Gelu") -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    x = numel(in_shape) * e
    out_shape = ishape(in_shape)
    node = OpNode(name=name, kind="Gelu", display=display, in_shapes=[ishape(in_shape)], out_shapes=[out_shape],
                  in_bytes=numel(in_shape) * DT["bf16"], enable_recompute=enable_recompute,
                  repr_args=f"enable_recompute={enable_recompute}")
    node.act.activation_mem_cache = 3 * x - (x if has_cached_inputs else 0)
    node.act.fwd_peak_mem_no_cache = 2 * x
    node.act.bwd_peak_mem_no_cache = 2 * x
    node.compute.fwd_accessed_mem = 2 * x
    node.compute.bwd_grad_act_accessed_mem = 2 * x
    node.sim_steps = lambda nd, system=None: [_atom(nd)]
    return node, out_shape


# ------------------------------------------------------------------------------------------------
# vocab-parallel cross entropy  (dense_module.py:2097-2363)
# ------------------------------------------------------------------------------------------------
def parallel_ce(st: StrategyConfig, in_shape, name, display="_VocabParallelCrossEntropy") -> Tuple[OpNode, list]:
    e = DT[st.dtype]
    B, S, V = in_shape
    n = B * S * V
    node = OpNode(name=name, kind="ParallelCE", display=display, in_shapes=[ishape(in_shape)], out_shapes=[[1]],
                  in_bytes=numel(in_shape) * DT["bf16"])
    if st.cross_entropy_loss_fusion:
        logits = n * e
        loss_buf = B * S * 4
        mdxy = 3 * B * S * 4
        mdxy_gather = 3 * B * S * st.tp_size * 4 if st.tp_size > 1 else 0
        node.act.activation_mem_cache = logits
        node.act.fwd_peak_mem_no_cache = logits + loss_buf + mdxy + mdxy_gather
        node.act.bwd_peak_mem_no_cache = 0
        node.compute.fwd_accessed_mem = 2 * n * e + B * S * 4
        node.compute.bwd_grad_act_accessed_mem = 2 * n * e + B * S * 4
    else:
        node.act.activation_mem_cache = n * 4
        node.act.fwd_peak_mem_no_cache = n * 4 + n * e
        node.act.bwd_peak_mem_no_cache = 0
        f = n * 4 + n * 2                       # cast
        f += (n + B * S) * 4                    # max
        f += (n + B * S + n) * 4                # logits - max
        f += (n * 2) * 4                        # exp
        f += (n + B) * 4                        # sum
        f += (n + B + n) * 4                    # exp / sum
        node.compute.fwd_accessed_mem = f
        node.compute.bwd_grad_act_accessed_mem = (n + B + n) * 4 + (n * 4 + n * 2)
    ce = "ce_fusion" if st.cross_entropy_loss_fusion else "ce"
    node.ops = {"fwd": ce, "bwd_grad_act": ce, "bwd_grad_w": "default"}
    scalar = B * S * 4
    if st.tp_size > 1:
        node.comms.append(_tp_comm(st, "all_reduce", scalar, "ParallelCE_FWD_TP", "fwd"))
        if st.cross_entropy_loss_fusion:
            node.comms.append(_tp_comm(st, "all_reduce", scalar + scalar, "ParallelCE_FWD_TP", "fwd"))
        else:
            node.comms.append(_tp_comm(st, "all_reduce", scalar, "ParallelCE_FWD_TP", "fwd"))
            node.comms.append(_tp_comm(st, "all_reduce", scalar, "ParallelCE_FWD_TP", "fwd"))
    node.ce_scalar = scalar

    def sim(nd: OpNode, system) -> List[SimStep]:
        def ar(size):
            return comm_step(system, "all_reduce", size, st.tp_size, st.tp_net, "all_reduce", bwd=False)
        steps = [_atom(nd), ar(scalar)]
        if st.cross_entropy_loss_fusion:
            steps.append(ar(2 * scalar))
        else:
            steps += [ar(scalar), ar(scalar)]
        return steps
    node.sim_steps = sim
    return node, [1]


# ------------------------------------------------------------------------------------------------
# FP8 cast in front of a linear  (dense_module.py:2365-2391)
# ------------------------------------------------------------------------------------------------
def fp8_quantizer(st: StrategyConfig, in_shape, name, *, enable_recompute=False, display="Float8Quantizer") -> Tuple[OpNode, list]:
    x = numel(in_shape) * DT["bf16"]
    y = numel(in_shape) * DT["fp8"]
    node = OpNode(name=name, kind="Float8Quantizer", display=display, in_shapes=[ishape(in_shape)],
                  out_shapes=[ishape(in_shape)], in_bytes=x, enable_recompute=enable_recompute,
                  repr_args=f"enable_recompute={enable_recompute}")
    node.act.fwd_peak_mem_no_cache = x + y
    node.compute.fwd_accessed_mem = x + y
    node.sim_steps = lambda nd, system=None: []      # the cast is not replayed as a separate op
    return node, ishape(in_shape)

"""Multi-head Latent Attention (DeepSeek-V2/V3) as IR ops.

Dataflow (reference dense_module.py:2569-2853, itself following Megatron's multi_latent_attention.py):

    q   = q_up(RMSNorm(q_down(x)))            (or a single q_proj when q_lora_rank is None)     [B,S,h*(d_qk+d_rope)]
    ckv = kv_down(x) -> split(kv_lora_rank | d_rope)
    kv  = kv_up(RMSNorm(ckv_latent)) -> split(k_nope | v)
    k   = concat(k_nope, rope(k_rope) broadcast over heads)
    o   = MLA-SDPA(concat(q, k, v))  with qk width d_qk+d_rope and v width d_v
    y   = out_proj(o)                         (a *column* linear here: MLA is modelled with tp == 1)

The split / rope / concat steps are zero-cost layout ops but remain in the op list: they carry recompute
flags and their input bytes matter when they open a recompute segment.
"""
from __future__ import annotations

from ..core.config import AttentionRecomputeConfig
from . import dense as D


def emit_mla_attention(em, layer_idx: int, shape, rc: AttentionRecomputeConfig):
    from .llm import _linear
    cfg, st = em.model, em.st
    assert st.tp_size == 1, "MLA do not support Tensor Parallel"
    tail = rc.megatron_layernorm
    cp_bypass = rc.megatron_mla_up_proj and st.cp_size > 1 and st.cp_comm_type == "a2a"
    up_tail = rc.megatron_mla_up_proj and not cp_bypass
    core_rc = rc.core_attn_recompute and not cp_bypass
    q_head = cfg.qk_head_dim + cfg.qk_pos_emb_head_dim
    heads = cfg.head_num // st.tp_size
    x_shape = list(shape)
    q_down = kv_down = None
    with em.scope("attention", "SelfAttention"):
        if cfg.q_lora_rank is None:
            _, q_shape = _linear(em, "linear_q_proj", "col", cfg.hidden_size, cfg.head_num * q_head, x_shape,
                                 "ColumnParallelLinear", layer_idx=layer_idx, enable_recompute=rc.q_up_recompute)
        else:
            q_down, s1 = _linear(em, "linear_q_down_proj", "col", cfg.hidden_size, cfg.q_lora_rank, x_shape,
                                 "ColumnParallelLinear", layer_idx=layer_idx, has_cached_inputs=tail,
                                 enable_recompute=rc.q_down_recompute, is_last_recompute=True,
                                 use_variance_tail_model=tail)
            node, s2 = D.layernorm(st, cfg.q_lora_rank, s1, em.name("q_layernorm"),
                                   enable_recompute=rc.q_layernorm_recompute)
            em.add(node)
            _, q_shape = _linear(em, "linear_q_up_proj", "col", cfg.q_lora_rank, cfg.head_num * q_head, s2,
                                 "ColumnParallelLinear", layer_idx=layer_idx, enable_recompute=rc.q_up_recompute)
        B, S = q_shape[0], q_shape[1]
        kv_down, ckv_shape = _linear(em, "linear_kv_down_proj", "col", cfg.hidden_size,
                                     cfg.kv_lora_rank + cfg.qk_pos_emb_head_dim, x_shape, "ColumnParallelLinear",
                                     layer_idx=layer_idx, has_cached_inputs=True, enable_recompute=rc.kv_down_recompute,
                                     is_last_recompute=True, use_variance_tail_model=tail)
        latent = [ckv_shape[0], ckv_shape[1], cfg.kv_lora_rank]
        k_rope = [ckv_shape[0], ckv_shape[1], 1, cfg.qk_pos_emb_head_dim]      # one shared rope head (viewed 4-d like the reference)
        em.add(D.passthrough("SplitModule", [ckv_shape], [latent, k_rope], em.name("kv_combined_Split"),
                             enable_recompute=rc.core_attn_recompute, display="kv_combined_Split"))
        node, s3 = D.layernorm(st, cfg.kv_lora_rank, latent, em.name("kv_layernorm"),
                               enable_recompute=rc.kv_layernorm_recompute)
        em.add(node)
        _, kv_shape = _linear(em, "linear_kv_up_proj", "col", cfg.kv_lora_rank,
                              cfg.head_num * (cfg.qk_head_dim + cfg.v_head_dim), s3, "ColumnParallelLinear",
                              layer_idx=layer_idx, enable_recompute=rc.kv_up_recompute)
        kv4 = [B, S, heads, cfg.qk_head_dim + cfg.v_head_dim]
        k_nope = [B, S, heads, cfg.qk_head_dim]
        value = [B, S, heads, cfg.v_head_dim]
        em.add(D.passthrough("SplitModule", [kv4], [k_nope, value], em.name("KV_Split"),
                             enable_recompute=rc.core_attn_recompute, display="KV_Split"))
        rope_in = [B, S, 1, cfg.qk_pos_emb_head_dim]
        em.add(D.passthrough("RotaryEmbedding", [rope_in], [[B, S, heads, cfg.qk_pos_emb_head_dim]], em.name("rotary_pos_emb"),
                             enable_recompute=rc.rope_recompute, display="RotaryEmbedding"))
        k_rope_b = [B, S, heads, cfg.qk_pos_emb_head_dim]
        key = [B, S, heads, q_head]
        em.add(D.passthrough("ConcatModule", [k_nope, k_rope_b], [key], em.name("K_pos_emb_Concat"),
                             enable_recompute=rc.core_attn_recompute, display="K_pos_emb_Concat"))
        q3 = [B, S, heads * q_head]
        k3 = [B, S, heads * q_head]
        v3 = [B, S, heads * cfg.v_head_dim]
        attn_in = [B, S, heads * (2 * q_head + cfg.v_head_dim)]
        em.add(D.passthrough("ConcatModule", [q3, k3, v3], [attn_in], em.name("QKV_Concat"),
                             enable_recompute=rc.core_attn_recompute, display="QKV_Concat"))
        node, o_shape = D.core_attention(st, em.system, q_head, cfg.head_num, cfg.kv_head_num, attn_in,
                                         em.name("core_attention"), v_head_dim=cfg.v_head_dim, mla=True,
                                         has_cached_inputs=up_tail, enable_recompute=core_rc, is_last_recompute=True,
                                         use_variance_tail_model=up_tail,
                                         cache_outputs=bool(st.use_flash_sdp and st.fp8))
        em.add(node)
        out, shape = _linear(em, "linear_out_proj", "col", cfg.v_head_dim * cfg.head_num, cfg.hidden_size, o_shape,
                             "ColumnParallelLinear", layer_idx=layer_idx, enable_recompute=rc.out_recompute)
        if (st.mla_rms_recompute or rc.megatron_layernorm) and st.recompute_granularity == "selective_recompute":
            if q_down is not None:
                q_down.is_breakpoint = True
            kv_down.is_breakpoint = True
        if out.enable_recompute and st.recompute_granularity == "selective_recompute":
            out.is_breakpoint = True
    return shape

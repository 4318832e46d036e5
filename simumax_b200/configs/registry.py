"""Built-in model zoo and strategy presets (architecture facts of public models; Megatron-style run presets).

Configs are kept as Python data and materialised to JSON on demand (``materialize()`` / ``get_simu_*_config``),
so every entry point that takes a config *path* keeps working, and users can still pass their own JSON files.
"""
from __future__ import annotations

from typing import Dict

# name -> ModelConfig fields
MODEL_ZOO: Dict[str, dict] = {
    'aquila2-13b': {'hidden_size': 5120, 'head_num': 40, 'kv_head_num': 40, 'head_size': 128, 'intermediate_size': 13824, 'layer_num': 40, 'vocab_size': 100008, 'use_swiglu': True},
    'aquila2-34b': {'hidden_size': 6144, 'head_num': 48, 'head_size': 128, 'kv_head_num': 16, 'intermediate_size': 24576, 'layer_num': 60, 'vocab_size': 32000, 'use_swiglu': True},
    'aquila2-70b': {'hidden_size': 8192, 'head_num': 64, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 28672, 'layer_num': 80, 'vocab_size': 100008, 'use_swiglu': True},
    'aquila2-7b': {'hidden_size': 4096, 'head_num': 32, 'kv_head_num': 32, 'head_size': 128, 'intermediate_size': 11008, 'layer_num': 32, 'vocab_size': 100008, 'use_swiglu': True},
    'deepseek-16b': {'model_type': 'moe', 'model_name': 'deepseek_16b', 'attention_type': 'mla', 'hidden_size': 2048, 'head_num': 16, 'kv_head_num': 16, 'head_size': 128, 'intermediate_size': 10944, 'moe_ffn_hidden_size': 1408, 'moe_shared_expert_intermediate_size': 2816, 'layer_num': 28, 'expert_num': 64, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': None, 'kv_lora_rank': 512, 'topk': 6, 'vocab_size': 102400, 'use_swiglu': True},
    'deepseek-1b': {'model_type': 'moe', 'model_name': 'deepseek_1b', 'attention_type': 'mla', 'hidden_size': 2048, 'head_num': 16, 'kv_head_num': 16, 'head_size': 128, 'intermediate_size': 10944, 'moe_ffn_hidden_size': 1408, 'moe_shared_expert_intermediate_size': 2816, 'layer_num': 8, 'expert_num': 64, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': 1536, 'kv_lora_rank': 512, 'topk': 6, 'vocab_size': 102400, 'use_swiglu': True, 'dense_layers': 1, 'moe_pad_expert_input_to_capacity': True, 'capacity': 1},
    'deepseek-32b': {'model_type': 'moe', 'model_name': 'deepseek_32b', 'attention_type': 'mla', 'hidden_size': 2048, 'head_num': 32, 'kv_head_num': 32, 'head_size': 128, 'intermediate_size': 12288, 'moe_ffn_hidden_size': 1536, 'moe_shared_expert_intermediate_size': 1536, 'layer_num': 40, 'expert_num': 80, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': 1536, 'kv_lora_rank': 512, 'topk': 7, 'vocab_size': 102400, 'use_swiglu': True, 'dense_layers': 1, 'moe_pad_expert_input_to_capacity': False, 'capacity': 1},
    'deepseekv2-lite': {'model_type': 'moe', 'model_name': 'deepseek_v2_lite', 'attention_type': 'mla', 'hidden_size': 2048, 'head_num': 16, 'kv_head_num': 16, 'head_size': 128, 'intermediate_size': 10944, 'moe_ffn_hidden_size': 1408, 'moe_shared_expert_intermediate_size': 2816, 'layer_num': 28, 'dense_layers': 1, 'expert_num': 64, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': None, 'kv_lora_rank': 512, 'topk': 6, 'vocab_size': 102400, 'use_swiglu': True},
    'deepseekv2': {'model_type': 'moe', 'model_name': 'deepseek_v2', 'attention_type': 'mla', 'hidden_size': 5120, 'head_num': 128, 'kv_head_num': 128, 'head_size': 128, 'intermediate_size': 12288, 'moe_ffn_hidden_size': 1536, 'moe_shared_expert_intermediate_size': 3072, 'layer_num': 60, 'dense_layers': 1, 'expert_num': 160, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': 1536, 'kv_lora_rank': 512, 'topk': 6, 'vocab_size': 102400, 'use_swiglu': True},
    'deepseekv3': {'model_type': 'moe', 'model_name': 'deepseek_v3', 'attention_type': 'mla', 'hidden_size': 7168, 'head_num': 128, 'kv_head_num': 128, 'head_size': 128, 'intermediate_size': 18432, 'moe_ffn_hidden_size': 2048, 'moe_shared_expert_intermediate_size': 2048, 'layer_num': 61, 'dense_layers': 3, 'expert_num': 256, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': 1536, 'kv_lora_rank': 512, 'topk': 8, 'vocab_size': 129280, 'use_swiglu': True},
    'gpt3-175b': {'hidden_size': 12288, 'head_num': 96, 'kv_head_num': 96, 'head_size': 128, 'intermediate_size': 49152, 'layer_num': 96, 'vocab_size': 51200, 'use_swiglu': False},
    'kimi-1T': {'model_type': 'moe', 'model_name': 'kimi-1T', 'attention_type': 'mla', 'hidden_size': 7168, 'head_num': 64, 'kv_head_num': 64, 'head_size': 128, 'intermediate_size': 18432, 'moe_ffn_hidden_size': 2048, 'moe_shared_expert_intermediate_size': 2048, 'layer_num': 61, 'dense_layers': 1, 'expert_num': 384, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'q_lora_rank': 1536, 'kv_lora_rank': 512, 'topk': 8, 'vocab_size': 163840, 'use_swiglu': True},
    'ling-1T': {'model_type': 'moe', 'model_name': 'ling-1T', 'attention_type': 'mha', 'layer_num': 80, 'hidden_size': 8192, 'intermediate_size': 18432, 'moe_ffn_hidden_size': 2048, 'moe_shared_expert_intermediate_size': 2048, 'vocab_size': 157184, 'expert_num': 256, 'kv_head_num': 8, 'head_size': 128, 'head_num': 64, 'v_head_dim': 128, 'qk_head_dim': 128, 'qk_pos_emb_head_dim': 64, 'topk': 8, 'dense_layers': 4, 'use_swiglu': True},
    'llama-33b': {'hidden_size': 6656, 'head_num': 52, 'kv_head_num': 52, 'head_size': 128, 'intermediate_size': 17920, 'layer_num': 60, 'vocab_size': 32000, 'use_swiglu': True},
    'llama2-13b': {'hidden_size': 5120, 'head_num': 40, 'kv_head_num': 40, 'head_size': 128, 'intermediate_size': 13824, 'layer_num': 40, 'vocab_size': 32000, 'use_swiglu': True},
    'llama2-70b': {'model_type': 'dense', 'model_name': 'llama2_70b', 'hidden_size': 8192, 'head_num': 64, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 28672, 'layer_num': 80, 'vocab_size': 32000, 'use_swiglu': True},
    'llama2-7b': {'model_type': 'dense', 'model_name': 'llama2_7b', 'hidden_size': 4096, 'head_num': 32, 'kv_head_num': 32, 'head_size': 128, 'intermediate_size': 11008, 'layer_num': 32, 'vocab_size': 32000, 'use_swiglu': True},
    'llama2-tiny': {'hidden_size': 4096, 'head_num': 32, 'kv_head_num': 32, 'head_size': 128, 'intermediate_size': 11008, 'layer_num': 8, 'vocab_size': 32000, 'use_swiglu': True},
    'llama3-405b_padding_128': {'model_type': 'dense', 'model_name': 'llama3_405b', 'hidden_size': 16384, 'head_num': 128, 'kv_head_num': 16, 'head_size': 128, 'intermediate_size': 53248, 'layer_num': 128, 'vocab_size': 128256, 'use_swiglu': True},
    'llama3-70b': {'model_type': 'dense', 'model_name': 'llama3_70b', 'hidden_size': 8192, 'head_num': 64, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 28672, 'layer_num': 80, 'vocab_size': 128256, 'use_swiglu': True},
    'llama3-8b': {'model_type': 'dense', 'model_name': 'llama3_8b', 'hidden_size': 4096, 'head_num': 32, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 14336, 'layer_num': 32, 'vocab_size': 128257, 'use_swiglu': True},
    'megatron-1t': {'hidden_size': 25600, 'head_num': 160, 'kv_head_num': 160, 'head_size': 160, 'intermediate_size': 102400, 'layer_num': 128, 'vocab_size': 51200, 'use_swiglu': False},
    'megatron-22b': {'hidden_size': 6144, 'head_num': 64, 'kv_head_num': 64, 'head_size': 96, 'intermediate_size': 24576, 'layer_num': 48, 'vocab_size': 51200, 'use_swiglu': False},
    'megatron-530b': {'hidden_size': 20480, 'head_num': 128, 'kv_head_num': 128, 'head_size': 160, 'intermediate_size': 81920, 'layer_num': 105, 'vocab_size': 51200, 'use_swiglu': False},
    'mixtral-8x1b': {'model_type': 'moe', 'model_name': 'mixtral_8x1b', 'hidden_size': 4096, 'head_num': 32, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 14336, 'layer_num': 4, 'expert_num': 8, 'topk': 2, 'vocab_size': 32000, 'use_swiglu': True},
    'mixtral-8x7b': {'model_type': 'moe', 'model_name': 'mixtral_8x7b', 'hidden_size': 4096, 'head_num': 32, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 14336, 'layer_num': 32, 'expert_num': 8, 'topk': 2, 'vocab_size': 32000, 'use_swiglu': True},
    'moe-8T': {'model_type': 'moe', 'model_name': 'moe-8T', 'attention_type': 'mha', 'layer_num': 128, 'hidden_size': 24576, 'head_num': 128, 'kv_head_num': 128, 'topk': 2, 'head_size': 192, 'expert_num': 16, 'intermediate_size': 98304, 'moe_ffn_hidden_size': 49152, 'moe_shared_expert_intermediate_size': 49152, 'vocab_size': 163840, 'use_swiglu': True},
    'qwen3-32b': {'model_name': 'qwen3-32b', 'model_type': 'dense', 'hidden_size': 5120, 'head_num': 64, 'kv_head_num': 8, 'head_size': 128, 'intermediate_size': 25600, 'layer_num': 64, 'vocab_size': 151936, 'use_swiglu': True},
}

# fields shared by every preset
_STRATEGY_BASE = {'use_flash_sdp': True, 'use_fp32_accum_grad': True, 'etp_size': 1, 'micro_batch_size': 1, 'use_math_sdp': False, 'mem_factor': 0.94, 'seq_len': 4096, 'micro_batch_num': 8, 'world_size': 8, 'zero_state': 1, 'moe_dispatcher_policy': 'all2all', 'enable_dropout': False, 'use_fused_norm': True, 'enable_sequence_parallel': True, 'dtype': 'bf16'}

# preset name -> overrides on top of _STRATEGY_BASE
STRATEGY_PRESETS: Dict[str, dict] = {
    'ep4_pp2_dp4_mbs1': {'tp_size': 1, 'pp_size': 2, 'ep_size': 4, 'interleaving_size': 1, 'enable_recompute': True},
    'ep4_pp2_dp4_mbs1_full_recompute': {'tp_size': 1, 'pp_size': 2, 'ep_size': 4, 'interleaving_size': 1, 'enable_recompute': True, 'recompute_granularity': 'full_block', 'recompute_layer_num': 2},
    'ep4_pp2_dp4_mbs1_selective_recompute': {'tp_size': 1, 'pp_size': 2, 'ep_size': 4, 'interleaving_size': 1, 'enable_recompute': True, 'recompute_granularity': 'selective_recompute', 'recompute_layer_num': 2, 'attn_recompute': True, 'mla_rms_recompute': True, 'mlp_recompute': True, 'mlp_rms_recompute': True},
    'ep8_pp1_dp8_mbs1': {'tp_size': 1, 'pp_size': 1, 'ep_size': 8, 'interleaving_size': 1, 'enable_recompute': True},
    'tp1_pp1_dp8_mbs1': {'tp_size': 1, 'pp_size': 1, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True},
    'tp1_pp2_dp4_mbs1': {'tp_size': 1, 'pp_size': 2, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True},
    'tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt': {'tp_size': 1, 'pp_size': 4, 'ep_size': 1, 'interleaving_size': 2, 'pp_comm_async': False, 'enable_recompute': False, 'recompute_granularity': 'full_block', 'recompute_layer_num': 0, 'tp_net': 'high_intra_node', 'etp_net': 'high_intra_node', 'ep_net': 'high_intra_node', 'pp_net': 'inter_node', 'dp_net': 'inter_node'},
    'tp2_pp1_dp4_mbs1': {'tp_size': 2, 'pp_size': 1, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True},
    'tp2_pp1_dp4_mbs1_full_recompute': {'tp_size': 2, 'pp_size': 1, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True, 'recompute_granularity': 'full_block', 'recompute_layer_num': 12},
    'tp2_pp1_dp4_mbs1_selective_recompute': {'tp_size': 2, 'pp_size': 1, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True, 'recompute_granularity': 'selective_recompute', 'recompute_layer_num': 12, 'mlp_recompute': True, 'attn_recompute': True},
    'tp4_pp1_dp2_mbs1': {'tp_size': 4, 'pp_size': 1, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True},
    'tp8_pp1_dp1_mbs1': {'tp_size': 8, 'pp_size': 1, 'ep_size': 1, 'interleaving_size': 1, 'enable_recompute': True},
}


def strategy_dict(name: str) -> dict:
    d = dict(_STRATEGY_BASE)
    d.update(STRATEGY_PRESETS[name])
    return d


def model_dict(name: str) -> dict:
    return dict(MODEL_ZOO[name])

"""Configuration triple (system / strategy / model) and the three cost primitives.

The JSON schemas, field names, shape-key strings and every numeric formula are the public
contract of the reference and are reproduced exactly (so that shipped and user configs load
unchanged and predictions agree to the last digit):

* reference ``simumax/core/config.py``: ``StrategyConfig`` :208-693, ``SystemConfig`` :745-1038,
  ``ModelConfig`` :1041-1227, recompute sub-configs :147-207.

What is different here: configs are *data only* plus pure cost functions; bookkeeping of table
hits/misses and realised collective bandwidth lives in a separate ``CostLedger`` object that is
swapped out per run (no hidden global state, safe to evaluate many strategies concurrently).
Unknown JSON keys raise ``TypeError`` like the reference (plain ``cls(**dict)``).
"""
from __future__ import annotations

import copy
import json
import math
import os
import re
import time
import warnings
from collections import OrderedDict
from dataclasses import asdict, dataclass, field, fields
from typing import Any, Dict, List, Optional, Union

# ------------------------------------------------------------------------------------------------
# environment flags (reference config.py:15-41)
# ------------------------------------------------------------------------------------------------
ENABLE_SIMU_GRAPH = int(os.environ.get("ENABLE_SIMU_GRAPH", "0"))
SIMU_CHECK = int(os.environ.get("SIMU_CHECK", "0"))
SIMU_DEBUG = int(os.environ.get("SIMU_DEBUG", "0"))


def debug_trace(msg: str) -> None:
    """``SIMU_DEBUG=1``: one line per cost-model decision (efficiency-table hit / miss, memory-bandwidth entry used, every
    collective's effective bandwidth + latency, model-flops breakdown).  (reference core/config.py:18,845-873,1013,1149)"""
    if SIMU_DEBUG:
        print(f"[simu-debug] {msg}")
_TMP_OVERRIDE = os.environ.get("SIMUMAX_TMP_PATH", "").strip()
if _TMP_OVERRIDE:
    TMP_PATH = _TMP_OVERRIDE
elif SIMU_CHECK:
    TMP_PATH = "tmp_check"
else:
    TMP_PATH = "tmp" + time.strftime("_%Y%m%d_%H%M%S", time.localtime())

NET_OPS = ("all_reduce", "all_gather", "reduce_scatter", "p2p", "all2all")
kNetOp = NET_OPS  # reference spelling

_capture_graph_only = False


def set_capture_graph_only(value: bool) -> None:
    global _capture_graph_only
    _capture_graph_only = bool(value)


def get_capture_graph_only() -> bool:
    return _capture_graph_only


GIB = 1024 ** 3
DTYPE_BYTES = {"fp32": 4, "fp16": 2, "bf16": 2, "fp8": 1, "int32": 4, "int64": 8}


def _jsonable(value):
    if isinstance(value, dict):
        return {k: _jsonable(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_jsonable(v) for v in value]
    if isinstance(value, set):
        return [_jsonable(v) for v in sorted(value)]
    return value


# ------------------------------------------------------------------------------------------------
# base
# ------------------------------------------------------------------------------------------------
@dataclass
class Config:
    """Dataclass base: JSON load, ``to_dict`` that also serialises every ``@property``."""

    def to_dict(self) -> Dict[str, Any]:
        out = asdict(self)
        for name in dir(type(self)):
            if isinstance(getattr(type(self), name, None), property):
                out[name] = getattr(self, name)
        return _jsonable(out)

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=False, ensure_ascii=False)

    def __str__(self) -> str:
        return self.to_json_string()

    def sanity_check(self) -> None:  # overridden
        return None

    @classmethod
    def init_from_dict(cls, config_dict: Dict[str, Any]):
        return cls(**config_dict)

    @staticmethod
    def read_json_file(path: str) -> Dict[str, Any]:
        with open(path, "r", encoding="utf-8") as fh:
            return json.load(fh)

    @classmethod
    def init_from_config_file(cls, path: str):
        return cls.init_from_dict(cls.read_json_file(path))


# ------------------------------------------------------------------------------------------------
# per-layer recompute switch sets
# ------------------------------------------------------------------------------------------------
@dataclass
class AttentionRecomputeConfig(Config):
    input_layernorm_recompute: bool = False
    q_down_recompute: bool = False
    kv_down_recompute: bool = False
    q_up_recompute: bool = False
    kv_up_recompute: bool = False
    q_layernorm_recompute: bool = False
    kv_layernorm_recompute: bool = False
    rope_recompute: bool = False
    core_attn_recompute: bool = False
    out_recompute: bool = False
    megatron_layernorm: bool = False
    megatron_mla_up_proj: bool = False

    _LEGACY = ("input_layernorm_recompute", "q_down_recompute", "kv_down_recompute", "q_up_recompute",
               "kv_up_recompute", "q_layernorm_recompute", "kv_layernorm_recompute", "rope_recompute",
               "core_attn_recompute", "out_recompute")

    def set_all_status(self, status: bool) -> None:
        for name in self._LEGACY:
            setattr(self, name, status)

    @property
    def is_recompute_all(self) -> bool:
        return all(getattr(self, f.name) for f in fields(self))


@dataclass
class MLPRecomputeConfig(Config):
    pre_mlp_norm_recompute: bool = False
    shared_linear_recompute: bool = False
    linear_recompute: bool = False
    router_recompute: bool = False
    permutation_recompute: bool = False
    megatron_layernorm: bool = False
    megatron_mlp: bool = False
    megatron_moe: bool = False
    megatron_moe_act: bool = False

    @property
    def is_recompute_all(self) -> bool:
        return (self.pre_mlp_norm_recompute and self.linear_recompute and self.router_recompute
                and self.permutation_recompute)


# ------------------------------------------------------------------------------------------------
# strategy
# ------------------------------------------------------------------------------------------------
def _version_tuple(version: Optional[str]):
    if not version:
        return None
    parts = re.findall(r"\d+", str(version))
    if not parts:
        return None
    nums = [int(p) for p in parts[:3]]
    nums += [0] * (3 - len(nums))
    return tuple(nums)


def _version_at_least(cur: Optional[str], minimum: Optional[str]) -> bool:
    a, b = _version_tuple(cur), _version_tuple(minimum)
    return a is not None and b is not None and a >= b


@dataclass
class StrategyConfig(Config):
    """Parallel / runtime strategy (reference config.py:208-299 for the field list)."""

    seq_len: Optional[int] = None
    micro_batch_size: Optional[int] = None
    micro_batch_num: Optional[int] = None
    dtype: Optional[str] = "bf16"
    fp8: Optional[bool] = False
    fp8_recipe: Optional[str] = "tensor"      # "tensor" (per-tensor scaled e4m3, TE delayed / current scaling) | "mxfp8" (32-element UE8M0 block scales in the MMA)

    world_size: Optional[int] = 8
    tp_size: int = 1
    cp_size: int = 1
    pp_size: int = 1
    ep_size: int = 1
    etp_size: int = 1
    cp_comm_type: str = "a2a"
    cp_a2a_mode: str = "async_cp"
    order_of_paralielism: str = "tp-cp-ep-dp-pp"
    moe_dispatcher_policy: str = "all2all"
    num_layers_in_first_pipeline_stage: Optional[int] = None
    num_layers_in_last_pipeline_stage: Optional[int] = None
    account_for_embedding_in_pipeline_split: bool = False
    account_for_loss_in_pipeline_split: bool = False

    grad_reduce_in_bf16: bool = False
    cache_groupgemm_col_fp8_inputs: Optional[bool] = False
    offload_groupgemm_col_inputs: Optional[bool] = False

    attn_recompute: bool = False
    mla_rms_recompute: bool = False
    mlp_recompute: bool = False
    mlp_rms_recompute: bool = False

    enable_sequence_parallel: bool = True
    interleaving_size: int = 1
    microbatch_group_size_per_vp_stage: Optional[int] = None
    pp_comm_async: bool = True
    enable_straggler_model: bool = True
    zero_state: int = 1

    attention_sparse_ratio: float = 0.0
    enable_dropout: bool = False
    use_fp32_accum_grad: bool = True
    use_accm_weight: bool = True

    enable_recompute: bool = True
    recompute_granularity: Optional[str] = None
    recompute_layer_num: int = 0
    recompute_variance: bool = False
    megatron_recompute: bool = False
    megatron_recompute_modules: Optional[List[str]] = None

    use_flash_sdp: bool = True
    use_math_sdp: bool = False
    use_fused_norm: bool = True
    use_fused_swiglu: bool = True
    use_fused_grad_accumulation: bool = True
    cross_entropy_loss_fusion: bool = False
    overlap_grad_reduce: bool = True
    # extension (reference charges DP collectives as fully exposed, perf_llm.py:1570): when True the bucketed ZeRO-1
    # reduce-scatter hides behind the last micro-batch's backward and the all-gather behind the next forward
    dp_overlap_model: bool = False
    dp_overlap_efficiency: float = 0.5      # fraction of the overlappable collective time that is actually hidden
    # extensions describing this repo's runtime instead of Megatron + NCCL (all default off = reference semantics):
    # * tp_overlap_model: the TP / SP collectives next to a GEMM run inside that GEMM (csrc/gemm_2cta.cu, AG / RS variants);
    #   only the measured exposed fraction is charged (tools/fused_check.py: (fused - gemm) / (unfused - gemm)), the wgrad
    #   re-gather disappears because the gathered activation is what gets saved (so LinearCol caches the full sequence)
    # * fused_optimizer_model: one l2-norm pass + one fused AdamW pass (read grad/master/m/v, write master/m/v/bf16 param,
    #   zero the grad) instead of Megatron's seven-kernel chain
    # * vpp_async_perf_via_simulation: perf() of asynchronous interleaved pipelining takes its schedule time from the
    #   discrete-event engine (the reference raises there and offers simulate() only)
    vpp_async_perf_via_simulation: bool = False
    # * cp_comm_type "all_gather" (completed: the reference stops at the flops) and "p2p" (ring attention) are extensions;
    #   cp_p2p_exposed is the fraction of the ring's neighbour transfers that is not hidden behind the previous block's attention
    cp_p2p_exposed: float = 1.0
    tp_overlap_model: bool = False
    tp_overlap_exposed_ag: Union[float, Dict[str, float]] = 0.6      # or {tp_size: exposure}: python -m simumax_b200.calib.fused
    tp_overlap_exposed_rs: Union[float, Dict[str, float]] = 0.4
    fused_optimizer_model: bool = False
    norm_bwd_model: bool = False             # RMSNorm backward priced from the "norm_bwd" bandwidth entry (falls back to "default")
    sparse_embedding_model: bool = False     # embedding fwd / wgrad touch only the batch's table rows (gather / red.add scatter kernels)

    te_version: Optional[str] = None
    te_dummy_wgrad_min_version: str = "2.3.0"
    te_cp_a2a_save_pre_posta2a_min_version: str = "2.8.0"
    te_grouped_linear_dummy_wgrad_min_version: str = "2.10.0"

    tp_net: Optional[str] = "auto"
    cp_net: Optional[str] = "auto"
    pp_net: Optional[str] = "auto"
    dp_net: Optional[str] = "auto"
    ep_net: Optional[str] = "auto"
    etp_net: Optional[str] = "auto"
    edp_net: Optional[str] = "auto"

    dispatch_probs: bool = False
    mem_factor: float = 0.94

    valid_recompute_granularity = ["full_block", "attn_only", "mlp_only", "sdp_only", "selective_recompute"]
    valid_megatron_recompute_modules = ["core_attn", "layernorm", "mla_up_proj", "moe_act", "mlp", "moe"]
    valid_cp_a2a_modes = ["async_cp", "sync_cp"]

    # -------------------------------------------------------------- construction helpers
    _FORMAT_PATTERNS = OrderedDict([
        ("seq_len", (r"seq(\d+)", 4096)), ("micro_batch_size", (r"mbs(\d+)", 1)),
        ("micro_batch_num", (r"mbc(\d+)", 1)), ("global_batch_size", (r"gbs(\d+)", 8)),
        ("tp_size", (r"tp(\d+)", 1)), ("cp_size", (r"cp(\d+)", 1)), ("ep_size", (r"ep(\d+)", 1)),
        ("pp_size", (r"pp(\d+)", 1)), ("world_size", (r"world_size:(\d+)", 8)),
    ])

    @classmethod
    def init_from_format_strings(cls, text: str) -> "StrategyConfig":
        """Parse the ``parallelism`` string form, e.g. ``seq4096.mbs1.mbc8.gbs32 tp2.cp1.ep1.pp1..., world_size:8``."""
        params = {}
        for name, (pattern, default) in cls._FORMAT_PATTERNS.items():
            m = re.search(pattern, text)
            if m:
                params[name] = int(m.group(1))
            elif default is not None:
                params[name] = default
                print(f"Warning: parameter {name} not found, use default {default}")
        gbs = params.pop("global_batch_size")
        strategy = cls(**params)
        strategy.reset_global_batch_size(gbs)
        return strategy

    # -------------------------------------------------------------- derived sizes
    @property
    def shard_size(self) -> int:
        return self.pp_size * self.tp_size * self.cp_size

    @property
    def dp_size(self) -> int:
        assert self.world_size % self.shard_size == 0
        return self.world_size // self.shard_size

    @property
    def global_batch_size(self) -> int:
        return self.micro_batch_size * self.micro_batch_num * self.dp_size

    @property
    def edp_size(self) -> int:
        return self.world_size // (self.ep_size * self.etp_size * self.pp_size)

    @property
    def parallelism(self) -> str:
        sp_tag = f"sp{self.tp_size}." if self.enable_sequence_parallel else ""
        return (f"seq{self.seq_len}.mbs{self.micro_batch_size}.mbc{self.micro_batch_num}.gbs{self.global_batch_size} "
                f"tp{self.tp_size}.{sp_tag}cp{self.cp_size}.ep{self.ep_size}.pp{self.pp_size}.dp{self.dp_size}."
                f"etp{self.etp_size}.edp{self.edp_size}, world_size:{self.world_size}")

    @property
    def net(self) -> str:
        return (f"pp_net={self.pp_net}, tp_net={self.tp_net}, cp_net={self.cp_net}, dp_net={self.dp_net}, "
                f"ep_net={self.ep_net}, etp_net={self.etp_net}")

    # -------------------------------------------------------------- TE feature gates
    @staticmethod
    def _version_tuple(version):
        return _version_tuple(version)

    @property
    def te_dummy_wgrad_memory_enabled(self) -> bool:
        return _version_at_least(self.te_version, self.te_dummy_wgrad_min_version)

    @property
    def te_grouped_linear_dummy_wgrad_memory_enabled(self) -> bool:
        return _version_at_least(self.te_version, self.te_grouped_linear_dummy_wgrad_min_version)

    @property
    def te_cp_a2a_saves_pre_posta2a_output(self) -> bool:
        return _version_at_least(self.te_version, self.te_cp_a2a_save_pre_posta2a_min_version)

    # -------------------------------------------------------------- recompute classification
    @property
    def megatron_recompute_module_set(self) -> set:
        return set(self.megatron_recompute_modules or [])

    @property
    def is_megatron_selective_recompute(self) -> bool:
        return bool(self.enable_recompute and self.recompute_layer_num > 0
                    and self.recompute_granularity == "selective_recompute" and self.megatron_recompute
                    and self.megatron_recompute_module_set)

    @property
    def use_variance_tail_model(self) -> bool:
        return bool(self.recompute_variance or (
            self.is_megatron_selective_recompute
            and self.megatron_recompute_module_set & {"layernorm", "mla_up_proj", "moe_act"}))

    def _recompute_flavours(self):
        active = self.recompute_layer_num > 0
        full = active and self.recompute_granularity == "full_block"
        partial = active and self.recompute_granularity in ("attn_only", "mlp_only", "sdp_only")
        selective = (active and self.recompute_granularity == "selective_recompute"
                     and any([self.attn_recompute, self.mla_rms_recompute, self.mlp_recompute, self.mlp_rms_recompute]))
        return full, partial, selective

    @property
    def is_recompute(self) -> bool:
        full, partial, selective = self._recompute_flavours()
        return bool(self.enable_recompute and (full or partial or selective or self.is_megatron_selective_recompute))

    @property
    def recompute_status(self) -> str:
        full, partial, selective = self._recompute_flavours()
        if not self.is_recompute:
            return "No Recompute"
        if full or partial:
            return f"{self.recompute_granularity}, recompute_layer_num={self.recompute_layer_num}"
        if self.is_megatron_selective_recompute:
            modules = ",".join(sorted(self.megatron_recompute_module_set))
            return (f"{self.recompute_granularity}, recompute_layer_num={self.recompute_layer_num}, "
                    f"megatron_recompute=True, modules=[{modules}]")
        if selective:
            return (f"{self.recompute_granularity}, recompute_layer_num={self.recompute_layer_num}, "
                    f"attn={self.attn_recompute}, attn_rms={self.mla_rms_recompute}, mlp={self.mlp_recompute}, "
                    f"mlp_rms={self.mlp_rms_recompute}, recompute_variance={self.recompute_variance}")
        return "Unknown Recompute Status"

    def parse_attention_recompute(self, layer_idx: int) -> AttentionRecomputeConfig:
        """Which attention sub-ops of layer ``layer_idx`` are re-run in backward (reference :469-520)."""
        conf = AttentionRecomputeConfig()
        if self.recompute_granularity is None or layer_idx >= self.recompute_layer_num:
            return conf
        g = self.recompute_granularity
        if self.is_megatron_selective_recompute:
            mods = self.megatron_recompute_module_set
            ln, up = "layernorm" in mods, "mla_up_proj" in mods
            conf.megatron_layernorm, conf.megatron_mla_up_proj = ln, up
            conf.input_layernorm_recompute = conf.q_down_recompute = conf.kv_down_recompute = ln
            conf.q_up_recompute = conf.kv_up_recompute = up
            conf.q_layernorm_recompute = conf.kv_layernorm_recompute = up
            conf.rope_recompute = conf.core_attn_recompute = up
            return conf
        if g == "full_block":
            conf.set_all_status(True)
        elif g == "attn_only":
            conf.set_all_status(True)
            conf.input_layernorm_recompute = False
        elif g == "sdp_only":
            conf.core_attn_recompute = True
        elif g == "mlp_only":
            pass
        elif g == "selective_recompute":
            if self.mla_rms_recompute:
                assert self.attn_recompute, "mla_rms_recompute requires attn_recompute"
            conf.input_layernorm_recompute = conf.q_down_recompute = conf.kv_down_recompute = self.mla_rms_recompute
            conf.q_up_recompute = conf.kv_up_recompute = self.attn_recompute
            conf.q_layernorm_recompute = conf.kv_layernorm_recompute = self.attn_recompute
            conf.rope_recompute = conf.core_attn_recompute = self.attn_recompute
            conf.out_recompute = False
        else:
            raise ValueError("Invalid recompute_granularity")
        return conf

    def parse_mlp_recompute(self, layer_idx: int) -> MLPRecomputeConfig:
        if self.recompute_granularity is None or layer_idx >= self.recompute_layer_num:
            return MLPRecomputeConfig()
        g = self.recompute_granularity
        if self.is_megatron_selective_recompute:
            mods = self.megatron_recompute_module_set
            moe = "moe" in mods
            ln = "layernorm" in mods
            return MLPRecomputeConfig(pre_mlp_norm_recompute=ln, megatron_layernorm=ln, megatron_mlp="mlp" in mods,
                                      megatron_moe=moe, megatron_moe_act=("moe_act" in mods and not moe))
        if g in ("full_block", "mlp_only"):
            return MLPRecomputeConfig(True, True, True, True, True)
        if g in ("attn_only", "sdp_only"):
            return MLPRecomputeConfig()
        if g == "selective_recompute":
            if self.mlp_rms_recompute:
                assert self.mlp_recompute, "mlp_rms_recompute requires mlp_recompute"
            return MLPRecomputeConfig(pre_mlp_norm_recompute=self.mlp_rms_recompute,
                                      shared_linear_recompute=self.mlp_rms_recompute,
                                      linear_recompute=self.mlp_recompute,
                                      router_recompute=self.mlp_rms_recompute,
                                      permutation_recompute=False)
        raise ValueError("Invalid recompute_granularity")

    def get_mesh_size(self, order: str = "tp-dp-pp"):
        sizes = []
        for axis in order.split("-"):
            assert axis in ("tp", "dp", "pp", "ep", "etp", "edp"), f"order {axis} is not supported"
            sizes.append(getattr(self, f"{axis}_size"))
        return sizes

    # -------------------------------------------------------------- validation
    def apply_runtime_semantics(self, on: bool = True) -> "StrategyConfig":
        """Price ``simumax_b200.runtime`` (this repo's fused kernels) instead of Megatron-LM + NCCL: TP / SP collectives inside the
        adjacent GEMM, ZeRO-1 collectives overlapped (single-GPU-per-replica layouts), single-pass fused optimizer, one fused
        cross-entropy pass, row-sparse embedding kernels, RMSNorm backward from its own bandwidth entry.  ``on=False`` only clears the
        extension knobs (the reference's semantics); ``cross_entropy_loss_fusion`` is a reference field and is left as configured."""
        on = bool(on)
        self.tp_overlap_model = self.fused_optimizer_model = self.sparse_embedding_model = self.norm_bwd_model = on
        self.dp_overlap_model = on and self.tp_size == 1
        if on:
            self.cross_entropy_loss_fusion = True
        return self

    def sanity_check(self) -> None:
        if self.order_of_paralielism != "tp-cp-ep-dp-pp":
            raise ValueError("Invalid order_of_paralielism, we only support tp-cp-ep-dp-pp now, "
                             f"but got {self.order_of_paralielism}")
        assert self.cp_a2a_mode in self.valid_cp_a2a_modes, \
            f"cp_a2a_mode {self.cp_a2a_mode} must be in [{','.join(self.valid_cp_a2a_modes)}]"
        if self.cache_groupgemm_col_fp8_inputs:
            assert self.fp8, "cache_groupgemm_col_fp8_inputs requires fp8"
        assert self.fp8_recipe in ("tensor", "mxfp8"), f"fp8_recipe must be 'tensor' or 'mxfp8', got {self.fp8_recipe}"
        if self.offload_groupgemm_col_inputs:
            assert self.recompute_granularity != "full_block", \
                "offload_groupgemm_col_inputs is not allowed when recompute_granularity = 'full_block'"
        assert self.seq_len % self.cp_size == 0, \
            f"seq_len must be divisible by cp_size, but seq_len = {self.seq_len}, cp_size = {self.cp_size}"
        assert self.world_size % self.shard_size == 0, (
            "world_size must be divisible by pp_size * tp_size * cp_szie, but "
            f"world_size = {self.world_size}, pp_size = {self.pp_size}, tp_size = {self.tp_size}, cp_size={self.cp_size}")
        assert self.zero_state in (0, 1, 2, 3), "zero_state must be in [0, 1, 2, 3]"
        assert self.recompute_granularity is None or self.recompute_granularity in self.valid_recompute_granularity, \
            f"recompute_granularity {self.recompute_granularity} must be in [{','.join(self.valid_recompute_granularity)}]"
        assert self.recompute_layer_num >= 0
        mods = self.megatron_recompute_module_set
        if not self.megatron_recompute:
            assert not mods, "megatron_recompute_modules requires megatron_recompute=True"
        else:
            assert self.enable_recompute, "megatron_recompute requires enable_recompute=True"
            assert self.recompute_granularity == "selective_recompute", \
                "megatron_recompute requires recompute_granularity='selective_recompute'"
            assert self.recompute_layer_num > 0, "megatron_recompute requires recompute_layer_num > 0"
            bad = mods.difference(self.valid_megatron_recompute_modules)
            assert not bad, f"invalid megatron_recompute_modules: {sorted(bad)}"
            assert mods, "megatron_recompute requires non-empty megatron_recompute_modules"
            assert "core_attn" not in mods, "megatron_recompute core_attn is not supported in SimuMax yet"
            assert not any([self.attn_recompute, self.mla_rms_recompute, self.mlp_recompute, self.mlp_rms_recompute,
                            self.recompute_variance]), \
                "megatron_recompute is mutually exclusive with legacy selective flags and recompute_variance"
        assert self.world_size % (self.ep_size * self.etp_size * self.pp_size) == 0, (
            "world_size must be divisible by ep_size * etp_size * pp_size, but "
            f"world_size = {self.world_size}, ep_size = {self.ep_size}, etp_size = {self.etp_size}, pp_size = {self.pp_size}")
        assert self.moe_dispatcher_policy in ("all2all", "all2all-seq"), \
            "moe_dispatcher_policy must be 'all2all' (legacy alias 'all2all-seq' is accepted with warning)"
        if self.moe_dispatcher_policy == "all2all-seq":
            warnings.warn("moe_dispatcher_policy='all2all-seq' is no longer supported. Falling back to 'all2all'.")
            self.moe_dispatcher_policy = "all2all"
        assert self.interleaving_size >= 1, "interleaving_size must be >= 1"
        if self.interleaving_size > 1:
            assert self.pp_size > 1, "interleaving_size > 1 requires pp_size > 1"
            assert self.pp_comm_async or self.pp_size > 2, (
                "When interleaved schedule is used and p2p communication overlap is disabled, "
                "pipeline-model-parallel size should be greater than 2 to avoid having multiple "
                "p2p sends and recvs between same 2 ranks per communication batch")
            if self.microbatch_group_size_per_vp_stage is None:
                self.microbatch_group_size_per_vp_stage = self.pp_size
            assert self.microbatch_group_size_per_vp_stage >= self.pp_size, (
                "microbatch_group_size_per_vp_stage must be >= pp_size "
                f"(got {self.microbatch_group_size_per_vp_stage} < {self.pp_size})")
            warnings.warn("interleaving_size is enabled. VPP-aware timing/simulation paths are active; "
                          "validate target configs with smoke/probe cases when introducing new schedules.")
        if self.enable_dropout:
            warnings.warn("enable_dropout is not supported yet, the configuration will be ignored.")
        if self.enable_recompute:
            warnings.warn("Recompute is currently in experimental feature.")
        if self.zero_state in (2, 3):
            warnings.warn("zero_state 2 and 3 are not supported yet, the configuration will be ignored.")
        if self.recompute_granularity == "full_block":
            self.recompute_variance = False  # Megatron full recompute has no variance tail

    def reset_global_batch_size(self, global_batch_size: int) -> None:
        per_step = self.dp_size * self.micro_batch_size
        assert global_batch_size % per_step == 0, (
            f"global_batch_size {global_batch_size} must be divisible by dp_size*miro_batch_size"
            f"(dp_size={self.dp_size}, micro_batch_size={self.micro_batch_size})")
        self.micro_batch_num = global_batch_size // per_step


# ------------------------------------------------------------------------------------------------
# system
# ------------------------------------------------------------------------------------------------
@dataclass
class BandwidthConfig:
    gbps: float
    efficient_factor: float
    latency_us: float
    fixed_latency: float = 0
    fixed_latency_us_by_comm_num: Optional[Dict[str, float]] = None


@dataclass
class CompOpConfig:
    tflops: float
    efficient_factor: float
    accurate_efficient_factor: Optional[dict] = None


@dataclass
class AcceleratorConfig:
    backend: str
    mem_gbs: float
    bandwidth: Dict[str, BandwidthConfig]
    op: Dict[str, CompOpConfig]
    mode: str
    # extension over the reference schema (absent / 0 = reference behaviour): a constant added to every op phase that
    # launches a kernel -- launch gap + prologue / tail the roofline does not see (the reference always
    # under-predicts time; a per-op launch term calibrated on the box is the most direct correction)
    kernel_launch_overhead_us: float = 0.0


@dataclass
class NetOpConfig:
    scale: float
    offset: float
    efficient_factor: Optional[float] = None
    latency_us: Optional[float] = None
    fixed_latency_us: Optional[float] = None
    fixed_latency_us_by_comm_num: Optional[Dict[str, float]] = None
    dp_fixed_bw: Optional[dict] = None
    # extension over the reference schema: per-group-size fits written by simumax_b200.calib (absent -> scalar values)
    efficient_factor_by_comm_num: Optional[Dict[str, float]] = None
    latency_us_by_comm_num: Optional[Dict[str, float]] = None


@dataclass
class NetworkConfig:
    processor_usage: float
    bandwidth: BandwidthConfig
    op: Dict[str, NetOpConfig]


def _by_comm_num(table, comm_num, default=None):
    if not table:
        return default
    for key in (str(comm_num), comm_num):
        if key in table:
            return table[key]
    return default


@dataclass
class SystemConfig(Config):
    """Machine description + the three time primitives (op / memory / network), all in milliseconds."""

    sys_name: str = "null"
    num_per_node: int = 8
    accelerator: AcceleratorConfig = None
    networks: Dict[str, NetworkConfig] = None
    real_comm_bw: dict = field(default_factory=OrderedDict)
    FC8: bool = False
    intra_with_pcie: bool = False
    miss_efficiency: dict = field(default_factory=OrderedDict)
    hit_efficiency: dict = field(default_factory=OrderedDict)

    @classmethod
    def init_from_dict(cls, config_dict: Dict[str, Any]) -> "SystemConfig":
        cfg = copy.deepcopy(config_dict)
        acc = cfg.pop("accelerator")
        nets = cfg.pop("networks")
        intra_with_pcie = nets.pop("intra_with_pcie") if "intra_with_pcie" in nets else False
        accelerator = AcceleratorConfig(
            backend=acc["backend"], mem_gbs=acc["mem_gbs"], mode=acc["mode"],
            kernel_launch_overhead_us=float(acc.get("kernel_launch_overhead_us", 0.0)),
            bandwidth={k: BandwidthConfig(**v) for k, v in acc["bandwidth"].items()},
            op={k: CompOpConfig(**v) for k, v in acc["op"].items()})
        networks = {
            name: NetworkConfig(processor_usage=net["processor_usage"], bandwidth=BandwidthConfig(**net["bandwidth"]),
                                op={k: NetOpConfig(**v) for k, v in net["op"].items()})
            for name, net in nets.items()}
        return cls(sys_name=cfg.pop("sys_name"), num_per_node=cfg.pop("num_per_node"), accelerator=accelerator,
                   networks=networks, FC8=cfg.pop("FC8", False), intra_with_pcie=intra_with_pcie)

    def to_config_dict(self) -> Dict[str, Any]:
        """Inverse of ``init_from_dict``: the on-disk system JSON schema (used by the calibration writer)."""
        def drop_none(d):
            return {k: v for k, v in d.items() if v is not None}
        acc = self.accelerator
        nets: Dict[str, Any] = {}
        if self.intra_with_pcie:
            nets["intra_with_pcie"] = True
        for name, net in self.networks.items():
            nets[name] = {"processor_usage": net.processor_usage, "bandwidth": drop_none(asdict(net.bandwidth)),
                          "op": {k: drop_none(asdict(v)) for k, v in net.op.items()}}
        out = {"sys_name": self.sys_name, "num_per_node": self.num_per_node}
        if self.FC8:
            out["FC8"] = True
        out["accelerator"] = {
            "backend": acc.backend, "mem_gbs": acc.mem_gbs,
            "op": {k: drop_none(asdict(v)) for k, v in acc.op.items()},
            "bandwidth": {k: drop_none(asdict(v)) for k, v in acc.bandwidth.items()},
            "mode": acc.mode}
        if acc.kernel_launch_overhead_us:
            out["accelerator"]["kernel_launch_overhead_us"] = acc.kernel_launch_overhead_us
        out["networks"] = nets
        return out

    # ---- bookkeeping (kept on the object for API parity; cleared by ``reset_record_info``)
    def record_miss_efficiency(self, op_name, flops, shape_desc, use_eff):
        if shape_desc:
            self.miss_efficiency.setdefault(op_name, {})[f"shape={shape_desc}"] = {"flops": flops, "use_eff": use_eff}

    def record_hit_efficiency(self, op_name, flops, shape_desc, eff):
        self.hit_efficiency.setdefault(op_name, {})[shape_desc] = (flops, eff)

    def record_net_bw(self, op_name, net, comm_num, comm_stage, base_bw, real_bw, eff_factor, total_time, comm_size,
                      latency):
        self.real_comm_bw.setdefault(op_name, {})[comm_stage.lower()] = {
            "net": net, "base_bw": base_bw, "real_bw": real_bw, "eff_factor": eff_factor, "comm_num": comm_num,
            "comm_size": comm_size, "total_time": total_time, "latency": latency, "FC8": self.FC8}

    def reset_record_info(self):
        self.miss_efficiency.clear()
        self.hit_efficiency.clear()
        self.real_comm_bw.clear()

    # ---- primitive 1: compute time from the efficiency table
    def lookup_op_efficiency(self, op_name: str, flops, shape_desc: str):
        op = self.accelerator.op.get(op_name)
        if op is None and op_name.startswith("mxfp8_"):      # a config without a block-scaled table prices it as per-tensor fp8
            op = self.accelerator.op.get(op_name[2:])
        if op is None:
            warnings.warn(f"{op_name} not exist on {self.accelerator.op.keys()}, use default value")
            op = self.accelerator.op.get("default")
            assert op is not None, f"default not exist on {self.accelerator.op}"
            self.record_miss_efficiency(op_name, flops, shape_desc, None)
        table = op.accurate_efficient_factor
        if table is not None and table.get(shape_desc) is not None:
            eff = table[shape_desc]
            self.record_hit_efficiency(op_name, flops, shape_desc, eff)
            if SIMU_DEBUG:
                debug_trace(f"{op_name} [{shape_desc}] table hit: efficiency {eff:.4f}, flops {flops:.4g}")
        else:
            eff = op.efficient_factor
            self.record_miss_efficiency(op_name, flops, shape_desc, eff)
            if SIMU_DEBUG:
                debug_trace(f"{op_name} [{shape_desc}] table MISS: fallback efficiency {eff:.4f}, flops {flops:.4g}")
        return op, eff

    def compute_op_accuracy_time(self, op_name: str, flops, shape_desc: str, reture_detail: bool = False):
        """flops / (tflops * eff) in ms; efficiency comes from the exact-string shape table when present."""
        if flops == 0:
            if reture_detail:
                return dict(op_name=op_name, tflops=None, efficient_factor=None, compute_only_time=0.0)
            return 0
        op, eff = self.lookup_op_efficiency(op_name, flops, shape_desc)
        t = flops / (op.tflops * 1e12 * eff) * 1e3
        if reture_detail:
            return dict(op_name=op_name, tflops=op.tflops, efficient_factor=eff, compute_only_time=t)
        return t

    # ---- primitive 2: memory-bound time
    def compute_mem_access_time(self, op_name, mem_bytes, reture_detail: bool = False):
        bw = self.accelerator.bandwidth.get(op_name) or self.accelerator.bandwidth.get("default")
        if SIMU_DEBUG and mem_bytes and op_name != "default":
            own = op_name in self.accelerator.bandwidth
            if SIMU_DEBUG:
                debug_trace(f"memory-bound {op_name}: {'own' if own else 'default'} bandwidth entry, efficiency {bw.efficient_factor:.3f}, {mem_bytes:.4g} B")
        t = mem_bytes / (bw.gbps * GIB * bw.efficient_factor) * 1e3 + bw.latency_us / 1e3
        if mem_bytes == 0:
            t = 0
        if reture_detail:
            return dict(gbps=bw.gbps, efficient_factor=bw.efficient_factor, latency_us=bw.latency_us, io_time=t)
        return t

    @staticmethod
    def _lookup_comm_num_value(values, comm_num, default=None):
        return _by_comm_num(values, comm_num, default)

    # ---- primitive 3: collective / p2p time (ring byte factors + latency model, reference :904-1017)
    def compute_net_op_time(self, op_name: str, size, comm_num: int, net: str = "", comm_stage: str = "unkonw",
                            strategy: Optional[StrategyConfig] = None):
        assert op_name in NET_OPS, f"{op_name} not exist on {NET_OPS}"
        net_data = self.networks.get(net)
        assert net_data is not None, f"{net} not exist on {self.networks.keys()}, op_name={op_name}"
        op = net_data.op.get(op_name)
        assert op is not None, f"{op_name} not exist on {net_data}"
        scale, offset = op.scale, op.offset
        eff = op.efficient_factor if op.efficient_factor is not None else net_data.bandwidth.efficient_factor
        eff = _by_comm_num(op.efficient_factor_by_comm_num, comm_num, eff)

        wire_bytes = size * scale
        wire_bytes += (wire_bytes / comm_num) * offset

        dense_dp = comm_stage in ("dp", "dp_cp")
        if "pcie" in net and dense_dp and op.dp_fixed_bw and op.dp_fixed_bw.get(str(comm_num)):
            fixed_bw = op.dp_fixed_bw.get(str(comm_num))
            self.real_comm_bw[op_name + "_dp"] = {"net": net, "bw": f"{fixed_bw} GB/S", "comm_num": comm_num,
                                                  "latency": None}
            return wire_bytes / (fixed_bw * GIB) * 1000

        bw = net_data.bandwidth.gbps
        if self.FC8 and net == "high_intra_node":
            bw *= (comm_num - 1) / 7
        if net == "inter_node":
            if op_name == "p2p":
                bw /= self.num_per_node
            if op_name == "all2all":
                stage = comm_stage.lower()
                if "ep" in stage or "cp" in stage:
                    k = max(1, math.ceil(comm_num / self.num_per_node))
                    wire_bytes = (k - 1) / k * wire_bytes
                    bw /= self.num_per_node
            if op_name in ("all_reduce", "all_gather", "reduce_scatter") and strategy is not None:
                if dense_dp:
                    mult = strategy.tp_size * (strategy.cp_size if comm_stage == "dp" else 1)
                    bw /= min(self.num_per_node, mult)
                elif comm_stage == "edp":
                    bw /= min(self.num_per_node, strategy.ep_size * strategy.etp_size)

        base_latency = op.latency_us if op.latency_us is not None else net_data.bandwidth.latency_us
        base_latency = _by_comm_num(op.latency_us_by_comm_num, comm_num, base_latency)
        fixed_latency = _by_comm_num(op.fixed_latency_us_by_comm_num, comm_num, op.fixed_latency_us)
        if fixed_latency is None:
            fixed_latency = _by_comm_num(net_data.bandwidth.fixed_latency_us_by_comm_num, comm_num,
                                         net_data.bandwidth.fixed_latency)
        if comm_num == 1:
            return 0
        latency = base_latency
        if self.num_per_node == 8 and op_name in ("all_reduce", "all_gather", "reduce_scatter", "all2all"):
            latency = base_latency * (comm_num + offset) * scale
        t = wire_bytes / (bw * GIB * eff) * 1e3 + (latency + fixed_latency) / 1e3
        if SIMU_DEBUG:
            debug_trace(f"{op_name} x{comm_num} on {net} ({comm_stage}): {size:.4g} B -> wire {wire_bytes:.4g} B at {bw * eff:.1f} GiB/s "
                    f"+ {latency:.1f} us + fixed {fixed_latency:.1f} us = {t * 1e3:.1f} us")
        self.record_net_bw(op_name, net, comm_num, comm_stage, net_data.bandwidth.gbps, bw * eff, eff, t * 1e3,
                           wire_bytes, latency)
        return t

    def compute_end2end_time(self, compute_time, mem_time):
        mode = self.accelerator.mode
        assert mode in ("only_compute", "roofline")
        if mode == "only_compute":
            return compute_time if compute_time != 0 else mem_time
        return max(compute_time, mem_time)


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------
@dataclass
class ModelConfig(Config):
    """Decoder-only transformer architecture (dense / MoE, MHA-GQA / MLA)."""

    hidden_size: int
    head_num: int
    kv_head_num: int
    model_type: Optional[str] = None
    model_name: Optional[str] = None
    head_size: Optional[int] = None
    intermediate_size: Optional[int] = None
    layer_num: Optional[int] = None
    vocab_size: Optional[int] = None
    orig_vocab_size: Optional[int] = None
    use_swiglu: Optional[bool] = None
    expert_num: int = 1
    topk: Optional[int] = None
    attention_type: str = "mha"
    moe_ffn_hidden_size: Optional[int] = None
    moe_shared_expert_intermediate_size: Optional[int] = None
    v_head_dim: Optional[int] = None
    qk_head_dim: Optional[int] = None
    qk_pos_emb_head_dim: Optional[int] = None
    q_lora_rank: Optional[int] = None
    kv_lora_rank: Optional[int] = None
    dense_layers: int = 0
    moe_pad_expert_input_to_capacity: bool = True
    capacity: int = 1
    group_linear_mode: str = "parallel"
    make_vocab_size_divisible_by = 128   # class attribute (Megatron default), not a JSON field
    padded_vocab_size = True             # class attribute

    def __post_init__(self):
        if self.moe_ffn_hidden_size is None:
            self.moe_ffn_hidden_size = self.intermediate_size
        if self.model_type is None:
            self.model_type = "moe" if self.expert_num > 1 else "dense"

    @classmethod
    def init_from_config_file(cls, path: str):
        d = cls.read_json_file(path)
        if d.get("moe_ffn_hidden_size") is None:
            d["moe_ffn_hidden_size"] = d["intermediate_size"]
        return cls.init_from_dict(d)

    def maybe_pad_vocab_size(self, tp_size: int, log: bool = False) -> None:
        """Megatron vocab padding to a multiple of ``make_vocab_size_divisible_by * tp`` (reference :1091)."""
        if not self.padded_vocab_size:
            return
        if self.orig_vocab_size is None:
            self.orig_vocab_size = self.vocab_size
        multiple = self.make_vocab_size_divisible_by * tp_size
        after = int(math.ceil(self.orig_vocab_size / multiple) * multiple)
        if log:
            print(" > padded vocab (size: {}) with {} dummy tokens (new size: {})".format(
                self.orig_vocab_size, after - self.orig_vocab_size, after), flush=True)
        self.vocab_size = after

    def set_vocab_size(self, vocab_size: int) -> None:
        self.orig_vocab_size = vocab_size
        self.vocab_size = vocab_size

    # ---- parameter counts
    @property
    def mlp_elements(self):
        return (3 if self.use_swiglu else 2) * self.hidden_size * self.moe_ffn_hidden_size

    @property
    def base_proj_elements(self):
        if self.attention_type == "mla":
            return self.v_head_dim * self.head_num * self.hidden_size
        return self.hidden_size * self.hidden_size

    @property
    def attn_proj_elements(self):
        return self.base_proj_elements

    @property
    def norm_elements(self):
        return self.hidden_size

    @property
    def qkv_proj_elements(self):
        assert self.head_num is not None
        kv_heads = self.head_num if self.kv_head_num is None else self.kv_head_num
        if self.attention_type == "mla":
            qk = self.qk_head_dim + self.qk_pos_emb_head_dim
            if self.q_lora_rank is None:
                n = self.hidden_size * self.head_num * qk
            else:
                n = self.hidden_size * self.q_lora_rank + self.q_lora_rank * self.head_num * qk
            n += self.hidden_size * (self.kv_lora_rank + self.qk_pos_emb_head_dim)
            n += self.kv_lora_rank * self.head_num * (self.qk_head_dim + self.v_head_dim)
            return n
        return self.hidden_size * (self.head_size * self.head_num + 2 * self.head_size * kv_heads)

    @property
    def vocab_elements(self):
        return self.vocab_size * self.hidden_size

    @property
    def layer_elements(self):
        return (self.qkv_proj_elements + 2 * self.norm_elements + self.attn_proj_elements
                + self.expert_num * self.mlp_elements)

    @property
    def layer_act_elements(self):
        factor = 1 + (self.topk - 1 if (self.topk is not None and self.topk > 1) else 0)
        return self.qkv_proj_elements + 2 * self.norm_elements + self.attn_proj_elements + factor * self.mlp_elements

    @property
    def param_numel(self):
        return 2 * self.vocab_elements + self.layer_elements * self.layer_num + self.norm_elements

    @property
    def activated_param_numel(self):
        return 2 * self.vocab_elements + self.layer_act_elements * self.layer_num + self.norm_elements

    def flops_per_token(self, context_seq_len: int, with_attn: bool = True):
        """Theoretical training FLOPs per token (6ND + attention), reference :1128-1156."""
        L = self.layer_num
        total = 0
        factor = 1
        if self.topk is not None and self.topk > 1:
            factor += self.topk - 1
            total += 6 * L * self.hidden_size * self.expert_num          # router
        if self.moe_shared_expert_intermediate_size is not None:
            factor += self.moe_shared_expert_intermediate_size / self.moe_ffn_hidden_size
        attn_matmul = 3 * 2 * L * (self.qkv_proj_elements + self.attn_proj_elements)
        mlp_matmul = 3 * 2 * L * self.mlp_elements * factor
        total += attn_matmul + mlp_matmul
        attn_sdp = 0
        if with_attn:
            if self.attention_type == "mla":
                attn_sdp = 6 * L * (context_seq_len * (self.qk_head_dim + self.qk_pos_emb_head_dim) * self.head_num
                                    + context_seq_len * self.v_head_dim * self.head_num)
            else:
                attn_sdp = 6 * L * (2 * context_seq_len * self.hidden_size)
            total += attn_sdp
        head = 6 * (self.hidden_size * self.vocab_size)
        if SIMU_DEBUG:
            debug_trace(f"model flops / token: per layer attention GEMMs {attn_matmul / L:.4g}, MLP {mlp_matmul / L:.4g}, "
                    f"core attention {attn_sdp / L:.4g}; LM head {head:.4g}")
        total += head
        return total

    def sanity_check(self) -> None:
        return None

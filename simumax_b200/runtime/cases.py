"""Predicted-vs-measured cases: one description drives the trainer that MEASURES a training step and the ``PerfLLM`` triple that
PREDICTS it.  Shared by ``bench.py`` (driver contract) and ``tools/b200/perf_vs_real.py`` (the validation table).

A case is a row of the reference's validation zoo (docs/b200/b200_release_v1.2_summary.md; tools/b200/run_megatron_perf_real_batch.py
builds the same rows for Megatron-LM): model, layer count, sequence length, micro-batch size / count and the parallel layout.
``PUBLISHED`` carries the Megatron-LM + TE numbers the reference authors measured on 8 x B200 for the rows that have them (quoted, not
re-measured here).
"""
from __future__ import annotations

import contextlib
import os
import gc
import io
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass
class Case:
    name: str
    model: str
    layers: Optional[int] = None          # None = the model's own depth
    seq: int = 4096
    mbs: int = 1
    mbc: int = 4
    tp: int = 1
    pp: int = 1
    ep: int = 1
    cp: int = 1
    dp: int = 1                           # world = tp * pp * cp * dp   (ep divides dp * cp for MoE models)
    recompute: str = "none"               # none | selective | full
    strategy_preset: str = "tp1_pp1_dp8_mbs1"
    strategy_overrides: Dict[str, object] = field(default_factory=dict)

    @property
    def world(self) -> int:
        return self.tp * self.pp * self.cp * self.dp

    @property
    def tokens_per_step(self) -> int:
        return self.mbs * self.mbc * self.dp * self.seq

    def parallelism(self) -> str:
        s = f"tp{self.tp}_pp{self.pp}_dp{self.dp}"
        if self.ep > 1:
            s += f"_ep{self.ep}"
        if self.cp > 1:
            s += f"_cp{self.cp}"
        if self.tp > 1:
            s += "_sp"
        if self.dp > 1:
            s += "_zero1"
        if self.recompute != "none":
            s += f"_{self.recompute}_recompute"
        return s


# (model, layers, seq, mbs, mbc, tp, pp, ep, dp) -> (Megatron real ms, real peak GiB, reference-predicted ms)   [BASELINE.md section 2]
PUBLISHED: Dict[Tuple, Tuple[float, float, float]] = {
    ("llama3-70b", 12, 4096, 1, 4, 1, 2, 1, 4): (843.0, 66.55, 828.76),
    ("llama3-70b", 12, 4096, 1, 4, 2, 1, 1, 4): (725.5, 60.68, 698.81),
    ("llama3-70b", 12, 4096, 1, 4, 4, 1, 1, 2): (409.8, 39.07, 395.23),
    ("llama3-70b", 12, 4096, 1, 4, 8, 1, 1, 1): (280.9, 28.29, 255.71),
    ("llama3-70b", 12, 4096, 1, 8, 2, 1, 1, 4): (1387.0, 60.68, 1309.8),
    ("llama3-70b", 12, 4096, 1, 8, 4, 1, 1, 2): (765.6, 39.07, 747.66),
    ("llama3-70b", 12, 4096, 1, 8, 8, 1, 1, 1): (556.7, 28.29, 493.39),
    ("llama3-70b", 12, 4096, 1, 32, 8, 1, 1, 1): (2073.7, 28.29, 1919.4),
    ("deepseekv2", 4, 4096, 1, 4, 1, 1, 8, 8): (366.6, 45.63, 343.59),
    ("deepseekv2", 4, 4096, 1, 8, 1, 1, 8, 8): (676.2, 45.63, 630.64),
    ("deepseekv2", 4, 4096, 1, 4, 1, 2, 4, 4): (282.1, 44.70, 251.91),
    ("llama3-405b_padding_128", 4, 4096, 1, 4, 8, 1, 1, 1): (258.8, 38.57, 255.83),
    ("llama3-405b_padding_128", 4, 4096, 1, 8, 8, 1, 1, 1): (488.1, 38.57, 486.81),
    ("llama3-405b_padding_128", 4, 4096, 1, 4, 2, 1, 1, 4): (907.3, 81.59, 860.73),
    ("llama3-405b_padding_128", 4, 4096, 1, 4, 4, 1, 1, 2): (476.0, 52.89, 452.92),
}


# context-parallel rows (all-to-all CP, dp = 1): (model, layers, seq, mbc, tp, cp) -> same triple   [BASELINE.md section 2, lines 57-63]
PUBLISHED_CP: Dict[Tuple, Tuple[float, float, float]] = {
    ("llama3-70b", 12, 32768, 4, 2, 4): (1758.0, 68.43, 1595.0), ("llama3-70b", 12, 32768, 4, 1, 8): (1768.1, 104.51, 1627.2),
    ("llama3-405b_padding_128", 4, 32768, 4, 2, 4): (1984.7, 87.03, 1814.3), ("llama3-405b_padding_128", 4, 32768, 4, 1, 8): (2003.6, 139.22, 1896.2),
    ("llama3-70b", 12, 131072, 4, 2, 4): (11315.7, 112.67, 10670.4), ("llama3-70b", 12, 131072, 4, 1, 8): (11301.9, 147.99, 10515.7),
    ("llama3-405b_padding_128", 4, 131072, 4, 2, 4): (10577.6, 118.13, 10043.5),
}


def published(case: Case):
    if case.recompute != "none":          # every published row runs without activation recompute
        return None
    if case.cp > 1:
        return PUBLISHED_CP.get((case.model, case.layers, case.seq, case.mbc, case.tp, case.cp)) if (case.mbs == 1 and case.dp == 1) else None
    return PUBLISHED.get((case.model, case.layers, case.seq, case.mbs, case.mbc, case.tp, case.pp, case.ep, case.dp))


def primary_case(n_gpus: int) -> Case:
    """BASELINE config 2 family (``perf_llama3_8b_tp8``): Llama-3-8B, seq 4096, 8 micro-batches of 1, tensor + sequence parallel over
    all N GPUs (N = 1: plain single-GPU step).  The global batch is the same at every N, so 1 -> 8 GPUs is strong scaling."""
    return Case(f"llama3_8b_tp{n_gpus}_pp1_dp1_mbc8", "llama3-8b", None, 4096, 1, 8, tp=n_gpus, dp=1,
                strategy_preset="tp8_pp1_dp1_mbs1")


def extra_cases(n_gpus: int) -> List[Case]:
    """Further BASELINE.json configs that fit N GPUs (each is a row -- or the <= N-GPU slice of a row -- of the published table)."""
    if n_gpus == 8:
        return [Case("llama3_70b_l12_tp2_pp1_dp4_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=2, dp=4, strategy_preset="tp2_pp1_dp4_mbs1"),
                Case("llama3_70b_l12_tp8_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=8, dp=1, strategy_preset="tp8_pp1_dp1_mbs1"),
                Case("deepseekv2_l4_ep8_pp1_dp8_mbc4", "deepseekv2", 4, 4096, 1, 4, ep=8, dp=8, strategy_preset="ep8_pp1_dp8_mbs1")]
    if n_gpus == 4:
        return [Case("llama3_70b_l12_tp4_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=4, dp=1, strategy_preset="tp4_pp1_dp2_mbs1")]
    if n_gpus == 2:
        return [Case("llama3_70b_l12_tp2_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, tp=2, dp=1, strategy_preset="tp2_pp1_dp4_mbs1")]
    return []


# ---------------------------------------------------------------------------------------------------------------- predicted side
def strategy_for(case: Case, runtime_semantics: bool = True):
    from ..configs import get_simu_strategy_config
    from ..core.config import StrategyConfig
    st = StrategyConfig.init_from_config_file(get_simu_strategy_config(case.strategy_preset))
    st.world_size, st.tp_size, st.pp_size, st.ep_size, st.cp_size = case.world, case.tp, case.pp, case.ep, case.cp
    st.enable_sequence_parallel = case.tp > 1
    st.seq_len, st.micro_batch_size, st.micro_batch_num = case.seq, case.mbs, case.mbc
    if case.recompute == "none":
        st.enable_recompute = False
    st.enable_straggler_model = False
    if case.model.startswith("deepseek") or case.ep > 1:
        st.dispatch_probs = True
        st.enable_sequence_parallel = False
    if case.cp > 1:
        st.cp_comm_type = "a2a"
    # runtime semantics: ZeRO-1 collectives overlapped with backward / next forward, TP collectives fused into the GEMMs (measured
    # exposure), single-pass AdamW.  Off = what Megatron + NCCL do (the reference's semantics).
    st.apply_runtime_semantics(runtime_semantics)
    for k, v in case.strategy_overrides.items():
        setattr(st, k, v)
    return st


def predict(case: Case, system: str = "b200_sm100a_calibrated", runtime_semantics: bool = True) -> dict:
    """What this repo's analytical model says about ``case`` on ``system`` (iteration ms, tokens/s, peak GiB)."""
    from ..configs import get_simu_model_config, get_simu_system_config
    from ..core.perf_llm import PerfLLM
    from ..core.utils import parse_bytes
    st = strategy_for(case, runtime_semantics)
    p = PerfLLM()
    with contextlib.redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=get_simu_model_config(case.model), system_config=system if os.path.sep in system else get_simu_system_config(system))
        if case.layers:
            p.model_config.layer_num = case.layers
        p.model_config.padded_vocab_size, p.model_config.make_vocab_size_divisible_by = True, 128
        if getattr(p.model_config, "model_type", "dense") == "moe":      # the measurement contract: pad to capacity, capacity factor 1
            p.model_config.moe_pad_expert_input_to_capacity, p.model_config.capacity = True, 1
        p.run_estimate()
        cost, mem = p.analysis_cost().data, p.analysis_mem().data
    tok_s = cost["throughput_per_accelerator"] * case.world
    peaks = _peak_gib(mem, parse_bytes)
    return {"system": system, "runtime_semantics": bool(runtime_semantics), "iter_ms": case.tokens_per_step / tok_s * 1e3,
            "tokens_per_s": tok_s, "mfu": cost.get("mfu"), "peak_gib": peaks,
            "efficiency_misses": {k: len(v) for k, v in p.system.miss_efficiency.items()}}


def _peak_gib(mem, parse_bytes) -> Optional[float]:
    """Largest per-stage peak of an ``analysis_mem`` result (single-stage results carry ``peak_mem`` at the top level)."""
    vals = []
    def walk(d):
        if isinstance(d, dict):
            for k, v in d.items():
                if k == "peak_mem" and isinstance(v, str):
                    vals.append(parse_bytes(v) / 2 ** 30)
                else:
                    walk(v)
    walk(mem)
    return max(vals) if vals else None


# ---------------------------------------------------------------------------------------------------------------- measured side
def model_dims(case: Case):
    from ..configs import get_simu_model_config
    from ..core.config import ModelConfig
    from .trainer import LlamaDims
    mc = ModelConfig.init_from_config_file(get_simu_model_config(case.model))
    dims = LlamaDims.from_model_config(mc)
    if case.layers:
        dims.layers = case.layers
    return mc, dims


_GROUP_CACHE: dict = {}


def parallel_groups(case: Case):
    """(tp_group, dp_group) of this rank for a tp x dp layout, rank order as in the analytical model: tp fastest, then dp.
    ``dist.new_group`` is collective over the default group, so every rank creates every group, once per layout."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    tp, dp = case.tp, world // case.tp
    key = (tp, dp)
    if key not in _GROUP_CACHE:
        tp_groups = [dist.new_group(list(range(i * tp, (i + 1) * tp))) for i in range(dp)] if dp > 1 and tp > 1 else None
        dp_groups = [dist.new_group(list(range(j, world, tp))) for j in range(tp)] if dp > 1 and tp > 1 else None
        _GROUP_CACHE[key] = (tp_groups, dp_groups)
    tp_groups, dp_groups = _GROUP_CACHE[key]
    if tp_groups is None:
        return (None, None) if dp == 1 else (None, dist.group.WORLD)
    return tp_groups[rank // tp], dp_groups[rank % tp]


def build_trainer(case: Case, seed: int = 1234, **kw):
    """The measured runtime for ``case`` on the current process group (one rank per GPU)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    assert world == case.world, f"case {case.name} needs {case.world} ranks, the process group has {world}"
    _, dims = model_dims(case)
    if case.ep > 1 or case.model.startswith("deepseek"):
        from .deepseek_trainer import DeepSeekTrainer            # MLA + MoE, expert parallel
        return DeepSeekTrainer.from_case(case, seed=seed, **kw)
    if case.pp > 1:
        from .pp_trainer import PPLlamaTrainer
        return PPLlamaTrainer.from_case(case, dims, seed=seed, **kw)
    if case.tp > 1:
        from .tp_trainer import TPLlamaTrainer
        tp_group, dp_group = parallel_groups(case)
        return TPLlamaTrainer(dims, case.mbs, case.seq, tp_group=tp_group, dp_group=dp_group, seed=seed,
                              recompute=case.recompute, **kw)
    from .trainer import LlamaTrainer
    if case.cp > 1:
        assert case.tp == 1 and case.dp == 1, "context parallelism runs with tp = dp = 1 (cp = world)"
        return LlamaTrainer(dims, case.mbs, case.seq, dp_group=dist.group.WORLD, cp_group=dist.group.WORLD, seed=seed,
                            recompute=case.recompute, **kw)
    return LlamaTrainer(dims, case.mbs, case.seq, dp_group=dist.group.WORLD if world > 1 else None, seed=seed,
                        recompute=case.recompute, **kw)


def device_memory_gib(trainer) -> dict:
    """Peak device memory of this rank: the caching allocator's high-water mark plus the symmetric (peer-mapped) allocations
    the allocator does not account for -- parameters / main grads of the DP trainer and every TP buffer live there."""
    import torch
    alloc = torch.cuda.max_memory_allocated()
    symm = 0
    seen = set()
    for name in ("ws", "ep_ws"):
        ws = getattr(trainer, name, None)
        if ws is not None and id(ws) not in seen:
            seen.add(id(ws))
            symm += ws.data.numel() + ws.flags.numel() * 4
    for sub in getattr(trainer, "symmetric_workspaces", lambda: [])():
        if id(sub) not in seen:
            seen.add(id(sub))
            symm += sub.data.numel() + sub.flags.numel() * 4
    free, total = torch.cuda.mem_get_info()
    return {"torch_max_allocated_gib": alloc / 2 ** 30, "symmetric_gib": symm / 2 ** 30, "peak_gib": (alloc + symm) / 2 ** 30,
            "device_used_now_gib": (total - free) / 2 ** 30}


def release(trainer) -> None:
    import torch
    del trainer
    gc.collect()
    torch.cuda.empty_cache()

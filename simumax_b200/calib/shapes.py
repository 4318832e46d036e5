"""Shape enumeration for calibration: run the analytical model with an empty efficiency table and harvest every
(op, shape-key, flops) it had to guess (``system.miss_efficiency``).  These are exactly the strings the model will look
up later, so whatever the sweeps write back is guaranteed to hit.

(reference flow: simu_tools/efficency_test/test_*_efficiency.py enumerate shapes out of ``analysis_op_info`` dumps;
here the cost function itself reports what it needs.)
"""
from __future__ import annotations

import contextlib
import io
import warnings
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Tuple

# (model, strategy preset, model overrides, strategy overrides) -- the reference's own example workloads on 8 GPUs
DEFAULT_WORKLOADS: List[Tuple[str, str, dict, dict]] = [
    ("llama3-8b", "tp1_pp1_dp8_mbs1", {}, {}),
    ("llama3-8b", "tp1_pp2_dp4_mbs1", {}, {}),
    ("llama3-8b", "tp2_pp1_dp4_mbs1", {}, {}),
    ("llama3-8b", "tp8_pp1_dp1_mbs1", {}, {}),
    ("llama3-70b", "tp2_pp1_dp4_mbs1", {"layer_num": 12}, {}),
    ("llama3-70b", "tp4_pp1_dp2_mbs1", {"layer_num": 12}, {}),
    ("deepseekv2", "ep8_pp1_dp8_mbs1", {"layer_num": 4}, {}),
    ("deepseekv2", "ep4_pp2_dp4_mbs1", {"layer_num": 4}, {}),
    # shapes of the predicted-vs-measured cases (tools/b200/perf_vs_real.py): bigger micro-batch, longer sequence, 70B widths at tp1
    ("llama3-8b", "tp1_pp1_dp8_mbs1", {"layer_num": 16}, {"micro_batch_size": 2}),
    ("llama3-8b", "tp1_pp1_dp8_mbs1", {"layer_num": 16}, {"seq_len": 8192}),
    ("llama3-70b", "tp1_pp1_dp8_mbs1", {"layer_num": 4}, {}),
    # tensor-parallel cases measured by TPLlamaTrainer (perf_vs_real.py --tp): 8B at tp4, 70B-l12 at tp8, 8K sequences
    ("llama3-8b", "tp4_pp1_dp2_mbs1", {}, {}),
    ("llama3-70b", "tp8_pp1_dp1_mbs1", {"layer_num": 12}, {}),
    ("llama3-70b", "tp2_pp1_dp4_mbs1", {"layer_num": 12}, {"seq_len": 8192}),
    ("llama3-70b", "tp4_pp1_dp2_mbs1", {"layer_num": 12}, {"seq_len": 8192}),
    ("llama3-405b_padding_128", "tp8_pp1_dp1_mbs1", {"layer_num": 4}, {}),
    # fp8 linears (the reference ships no fp8 numbers for B200 at all)
    ("llama3-8b", "tp1_pp1_dp8_mbs1", {}, {"fp8": True}),
    ("llama3-70b", "tp2_pp1_dp4_mbs1", {"layer_num": 12}, {"fp8": True}),
    ("deepseekv2", "ep8_pp1_dp8_mbs1", {"layer_num": 4}, {"fp8": True}),
]


@dataclass
class ShapeRequest:
    op: str                 # matmul | fp8_matmul | group_matmul | fp8_group_matmul | sdp_fwd | sdp_bwd
    key: str                # exact table key (without the "shape=" prefix)
    flops: float            # what the model divides by tflops*eff for this key
    fields: Dict[str, str] = field(default_factory=dict)

    def geti(self, name: str) -> int:
        return int(self.fields[name])

    def getb(self, name: str) -> bool:
        return self.fields.get(name, "False") == "True"


def parse_key(key: str) -> Dict[str, str]:
    out = {}
    for part in key.split(", "):
        if "=" in part:
            k, v = part.split("=", 1)
            out[k.strip()] = v.strip()
    return out


def collect_requests(workloads: Optional[Iterable[Tuple[str, str, dict, dict]]] = None, system: str = "b200_nominal",
                     extra_ops: bool = True) -> Dict[str, Dict[str, ShapeRequest]]:
    """op -> key -> ShapeRequest for every shape the workloads touch."""
    from ..configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
    from ..core.config import ModelConfig, StrategyConfig, SystemConfig
    from ..core.perf_llm import PerfLLM
    out: Dict[str, Dict[str, ShapeRequest]] = {}
    for model, strategy, m_over, s_over in (workloads or DEFAULT_WORKLOADS):
        st = StrategyConfig.init_from_config_file(get_simu_strategy_config(strategy))
        for k, v in s_over.items():
            setattr(st, k, v)
        sysc = SystemConfig.init_from_config_file(get_simu_system_config(system))
        for op in sysc.accelerator.op.values():      # force every lookup to miss
            op.accurate_efficient_factor = {}
        mc = ModelConfig.init_from_config_file(get_simu_model_config(model))
        for k, v in m_over.items():
            setattr(mc, k, v)
        p = PerfLLM()
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            p.configure(strategy_config=st, model_config=mc, system_config=sysc)
            p.model_config.padded_vocab_size, p.model_config.make_vocab_size_divisible_by = True, 128
            p.run_estimate()
        for op, table in p.system.miss_efficiency.items():
            for k, rec in table.items():
                key = k[len("shape="):] if k.startswith("shape=") else k
                out.setdefault(op, {})[key] = ShapeRequest(op, key, float(rec["flops"]), parse_key(key))
    return out


# ------------------------------------------------------------------------------------------------------------------
# user sweep file (reference: simu_tools/efficency_test/utils.py:28-77 ``PARAM_FILE``): the cross product of these lists is turned
# into workloads, and the analytical model itself enumerates the shapes each of them needs.
DEFAULT_SWEEP = {
    "model_list": ["llama3-8b", "llama3-70b", "llama3-405b_padding_128", "deepseekv2", "deepseekv3", "mixtral-8x7b", "qwen3-32b", "llama2-13b"],
    "mbs_list": [1, 2],
    "seq_len_list": [4096],
    "tp_list": [1, 2, 4, 8],
    "cp_list": [1],
    "ep_list": [1, 8],
    "dtype": ["bf16"],
    "layer_num": 4,                      # shapes do not depend on depth: keep the enumeration fast
    "long_context": [{"model": "llama3-70b", "seq_len": 32768, "tp": 2, "cp": 4}, {"model": "llama3-70b", "seq_len": 32768, "tp": 1, "cp": 8},
                     {"model": "llama3-70b", "seq_len": 131072, "tp": 2, "cp": 4}, {"model": "llama3-70b", "seq_len": 131072, "tp": 1, "cp": 8},
                     {"model": "llama3-405b_padding_128", "seq_len": 32768, "tp": 2, "cp": 4},
                     {"model": "llama3-405b_padding_128", "seq_len": 131072, "tp": 2, "cp": 4}],
}


def workloads_from_params(params: Optional[dict] = None, world: int = 8) -> List[Tuple[str, str, dict, dict]]:
    """Expand a sweep description (``DEFAULT_SWEEP`` schema; any key may be omitted) into ``collect_requests`` workloads.
    Combinations the strategy sanity check rejects (heads not divisible by tp, experts by ep, ...) are skipped when enumerated."""
    from ..configs import get_simu_model_config
    from ..core.config import ModelConfig
    p = dict(DEFAULT_SWEEP)
    p.update(params or {})
    out: List[Tuple[str, str, dict, dict]] = []
    layers = p.get("layer_num")
    for model in p["model_list"]:
        mc = ModelConfig.init_from_config_file(get_simu_model_config(model))
        moe = getattr(mc, "model_type", "dense") == "moe"
        m_over = {"layer_num": min(layers, mc.layer_num)} if layers else {}
        if moe:
            m_over.update({"moe_pad_expert_input_to_capacity": True, "capacity": 1})
        for dtype in p["dtype"]:
            for seq in p["seq_len_list"]:
                for mbs in p["mbs_list"]:
                    for tp in p["tp_list"]:
                        for cp in p["cp_list"]:
                            for ep in (p["ep_list"] if moe else [1]):
                                if tp * cp > world or (moe and (mc.expert_num % ep or ep > world // tp)):
                                    continue
                                s_over = {"world_size": world, "tp_size": tp, "cp_size": cp, "ep_size": ep, "pp_size": 1,
                                          "enable_sequence_parallel": tp > 1, "seq_len": seq, "micro_batch_size": mbs,
                                          "micro_batch_num": 1, "fp8": dtype == "fp8"}
                                out.append((model, "tp1_pp1_dp8_mbs1", dict(m_over), s_over))
    for lc in p.get("long_context", []):
        s_over = {"world_size": world, "tp_size": lc["tp"], "cp_size": lc["cp"], "ep_size": 1, "pp_size": 1,
                  "enable_sequence_parallel": lc["tp"] > 1, "seq_len": lc["seq_len"], "micro_batch_size": 1, "micro_batch_num": 1}
        out.append((lc["model"], "tp1_pp1_dp8_mbs1", {"layer_num": layers or 4}, s_over))
    return out


def collect_requests_tolerant(workloads, system: str = "b200_nominal", log=None) -> Dict[str, Dict[str, ShapeRequest]]:
    """``collect_requests`` over a generated sweep: workloads the configuration checks reject are skipped (and counted)."""
    out: Dict[str, Dict[str, ShapeRequest]] = {}
    skipped = 0
    for w in workloads:
        try:
            part = collect_requests([w], system=system)
        except (AssertionError, ValueError, ZeroDivisionError, KeyError) as e:
            skipped += 1
            continue
        for op, table in part.items():
            out.setdefault(op, {}).update(table)
    if log:
        log(f"[calib] sweep: {len(workloads)} workloads, {skipped} rejected by the configuration checks, "
            + ", ".join(f"{op}:{len(t)}" for op, t in out.items()))
    return out

"""Shared driver for the example scripts: one table of (model, strategy preset, layer override) cases and a runner that
configures PerfLLM, applies the Megatron vocab padding used by the B200 benchmark runs and prints / saves the analysis.

Usage from a case script:    from _cases import run; run("llama3_8b_tp1_pp2")
Direct:                      python examples/_cases.py llama3_8b_tp2 --system b200_sm100a_calibrated --out /tmp/x
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from simumax_b200 import PerfLLM  # noqa: E402
from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config  # noqa: E402
from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig  # noqa: E402

#           case name                                 model          strategy preset                          layers
CASES = {
    "llama3_8b_tp1_pp2":                             ("llama3-8b",  "tp1_pp2_dp4_mbs1",                        None),
    "llama3_8b_tp2":                                 ("llama3-8b",  "tp2_pp1_dp4_mbs1",                        None),
    "llama3_8b_tp4":                                 ("llama3-8b",  "tp4_pp1_dp2_mbs1",                        None),
    "llama3_8b_tp8":                                 ("llama3-8b",  "tp8_pp1_dp1_mbs1",                        None),
    "llama3_70b_layer12_tp1_pp2":                    ("llama3-70b", "tp1_pp2_dp4_mbs1",                        12),
    "llama3_70b_layer12_tp2":                        ("llama3-70b", "tp2_pp1_dp4_mbs1",                        12),
    "llama3_70b_layer12_tp2_full_recompute":         ("llama3-70b", "tp2_pp1_dp4_mbs1_full_recompute",         12),
    "llama3_70b_layer12_tp2_selective_recompute":    ("llama3-70b", "tp2_pp1_dp4_mbs1_selective_recompute",    12),
    "llama3_70b_layer12_tp4":                        ("llama3-70b", "tp4_pp1_dp2_mbs1",                        12),
    "llama3_70b_layer12_tp8":                        ("llama3-70b", "tp8_pp1_dp1_mbs1",                        12),
    "deepseekv2_layer4_ep8_pp1":                     ("deepseekv2", "ep8_pp1_dp8_mbs1",                        4),
    "deepseekv2_layer4_ep4_pp2":                     ("deepseekv2", "ep4_pp2_dp4_mbs1",                        4),
    "deepseekv2_layer4_ep4_pp2_full_recompute":      ("deepseekv2", "ep4_pp2_dp4_mbs1_full_recompute",         4),
    "deepseekv2_layer4_ep4_pp2_selective_recompute": ("deepseekv2", "ep4_pp2_dp4_mbs1_selective_recompute",    4),
}


def default_system() -> str:
    """Prefer the config calibrated on this box with our own kernels; fall back to nominal constants."""
    for name in (os.environ.get("SIMU_SYSTEM", ""), "b200_sm100a_calibrated", "b200_nominal"):
        if not name:
            continue
        try:
            get_simu_system_config(name)
            return name
        except Exception:
            continue
    raise RuntimeError("no system config found")


def build(case: str, system: str | None = None) -> PerfLLM:
    model, strategy, layers = CASES[case]
    perf = PerfLLM()
    perf.configure(strategy_config=StrategyConfig.init_from_config_file(get_simu_strategy_config(strategy)),
                   model_config=ModelConfig.init_from_config_file(get_simu_model_config(model)),
                   system_config=SystemConfig.init_from_config_file(get_simu_system_config(system or default_system())))
    if layers:
        perf.model_config.layer_num = layers
    perf.model_config.padded_vocab_size = True           # Megatron pads the vocab to a multiple of 128 * tp
    perf.model_config.make_vocab_size_divisible_by = 128
    if perf.model_config.expert_num and perf.model_config.expert_num > 1:
        perf.model_config.moe_pad_expert_input_to_capacity = True       # balanced routing, capacity factor 1
        perf.model_config.capacity = 1
    return perf


def run(case: str, system: str | None = None, out: str | None = None):
    perf = build(case, system)
    perf.run_estimate()
    tag = out or f"{perf.model_config.model_name}_{perf.system.sys_name}"
    return perf.analysis(tag)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=sorted(CASES))
    ap.add_argument("--system", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    run(a.case, a.system, a.out)

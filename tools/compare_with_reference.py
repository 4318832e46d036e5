"""Dev tool: run the unmodified reference and this framework on the same config triples and diff headline numbers.

    python tools/compare_with_reference.py            # default case list
Requires the reference checkout (REF env var, default /root/reference); CPU only.
"""
import io
import json
import os
import sys
import warnings
from contextlib import redirect_stdout

warnings.filterwarnings("ignore")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, REPO)
CFG = f"{REF}/configs"

CASES = [
    # (strategy json, model json, layer override, model overrides, strategy overrides)
    ("tp1_pp2_dp4_mbs1", "llama3-8b", None, {}, {}),
    ("tp8_pp1_dp1_mbs1", "llama3-8b", None, {}, {}),
    ("tp2_pp1_dp4_mbs1_selective_recompute", "llama3-70b", 12, {}, {}),
    ("tp2_pp1_dp4_mbs1_full_recompute", "llama3-70b", 12, {}, {}),
    ("ep8_pp1_dp8_mbs1", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1}, {}),
    ("ep4_pp2_dp4_mbs1", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1}, {}),
    ("ep4_pp2_dp4_mbs1_full_recompute", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1}, {}),
    ("ep4_pp2_dp4_mbs1_selective_recompute", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1}, {}),
    ("ep8_pp1_dp8_mbs1", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1},
     {"dispatch_probs": True, "cross_entropy_loss_fusion": True, "te_version": "2.11.0", "micro_batch_num": 4,
      "enable_recompute": False, "enable_straggler_model": False}),
    ("ep8_pp1_dp8_mbs1", "deepseekv3", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1},
     {"dispatch_probs": True, "cross_entropy_loss_fusion": True, "te_version": "2.11.0", "micro_batch_num": 4}),
    ("ep8_pp1_dp8_mbs1", "mixtral-8x7b", 4, {}, {"etp_size": 1, "tp_size": 1}),
    ("tp2_pp1_dp4_mbs1", "llama3-70b", 12, {}, {"fp8": True}),
    ("tp2_pp1_dp4_mbs1", "llama3-70b", 12, {}, {"cp_size": 2, "seq_len": 16384, "cp_a2a_mode": "sync_cp"}),
    ("tp4_pp1_dp2_mbs1", "llama3-70b", 12, {}, {"enable_sequence_parallel": False}),
    ("tp2_pp1_dp4_mbs1", "llama2-7b", None, {}, {"pp_size": 4, "tp_size": 1, "micro_batch_num": 16}),
    ("tp2_pp1_dp4_mbs1", "llama3-70b", 12, {}, {"recompute_granularity": "selective_recompute", "recompute_layer_num": 12,
                                                 "megatron_recompute": True, "megatron_recompute_modules": ["layernorm", "mlp"]}),
    ("ep8_pp1_dp8_mbs1", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1},
     {"recompute_granularity": "selective_recompute", "recompute_layer_num": 4, "megatron_recompute": True,
      "megatron_recompute_modules": ["layernorm", "mla_up_proj", "moe_act"]}),
    ("ep8_pp1_dp8_mbs1", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1},
     {"recompute_granularity": "selective_recompute", "recompute_layer_num": 4, "megatron_recompute": True,
      "megatron_recompute_modules": ["moe"]}),
    ("ep4_pp2_dp4_mbs1", "deepseekv2", 4, {"moe_pad_expert_input_to_capacity": True, "capacity": 1}, {"etp_size": 2, "ep_size": 2, "fp8": True}),
    # found by tools/fuzz_vs_reference.py: interleaved stages with pp > 2 (memory result keys of the middle ranks)
    ("ep4_pp2_dp4_mbs1", "mixtral-8x7b", 8, {"moe_pad_expert_input_to_capacity": True, "capacity": 1},
     {"world_size": 8, "pp_size": 4, "ep_size": 1, "tp_size": 1, "etp_size": 1, "enable_sequence_parallel": False, "seq_len": 16384,
      "micro_batch_size": 1, "micro_batch_num": 16, "te_version": "2.11.0", "interleaving_size": 2, "pp_comm_async": False}),
    # found by the fuzzer: fp8 grouped linear opening a recompute segment keeps its 1-byte input (cache_groupgemm_col_fp8_inputs)
    ("ep4_pp2_dp4_mbs1_full_recompute", "deepseekv3", 4, {},
     {"world_size": 16, "pp_size": 4, "ep_size": 2, "tp_size": 1, "etp_size": 1, "enable_sequence_parallel": False, "seq_len": 2048,
      "micro_batch_size": 2, "micro_batch_num": 4, "fp8": True, "cache_groupgemm_col_fp8_inputs": True}, "a100_pcie"),
]


def run(mod_prefix, case, system):
    strategy, model, layers, mover, sover = case[:5]
    if mod_prefix == "ref":
        sys.path.insert(0, REF)
        from simumax.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax_b200.core.perf_llm import PerfLLM
    st = StrategyConfig.init_from_config_file(f"{CFG}/strategy/{strategy}.json")
    for k, v in sover.items():
        setattr(st, k, v)
    p = PerfLLM()
    p._search_verbose = False
    with redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=ModelConfig.init_from_config_file(f"{CFG}/models/{model}.json"),
                    system_config=SystemConfig.init_from_config_file(f"{CFG}/system/{system}.json"))
        if layers:
            p.model_config.layer_num = layers
        for k, v in mover.items():
            setattr(p.model_config, k, v)
        p.model_config.padded_vocab_size = True
        p.model_config.make_vocab_size_divisible_by = 128
        p.run_estimate()
        mem = p.analysis_mem().data
        cost = p.analysis_cost().data
    if "peak_mem" in mem:
        peaks = {"first": mem["peak_mem"]}
    else:
        peaks = {k: v["peak_mem"] for k, v in mem.items()}
    return {"iter": cost["duration_time_per_iter"], "tgs": round(cost["throughput_per_accelerator"], 4),
            "mfu": round(cost["mfu"], 6), "peaks": peaks, "miss": sorted(p.system.miss_efficiency.keys())}


def main():
    system = os.environ.get("SYSTEM", "b200_bf16_ceperm")
    only = [int(x) for x in sys.argv[1:]]
    bad = 0
    for i, case in enumerate(CASES):
        if only and i not in only:
            continue
        sysname = case[5] if len(case) > 5 else system          # optional per-case system config
        try:
            a = run("ref", case, sysname)
        except Exception as ex:
            a = {"error": f"{type(ex).__name__}: {ex}"}
        try:
            b = run("new", case, sysname)
        except Exception as ex:
            import traceback
            traceback.print_exc()
            b = {"error": f"{type(ex).__name__}: {ex}"}
        ok = a == b or ("error" in a and "error" in b)
        bad += not ok
        print(("OK  " if ok else "DIFF"), i, case[0], case[1], case[4] if case[4] else "")
        if not ok:
            print("   ref:", json.dumps(a))
            print("   new:", json.dumps(b))
        elif "error" in a:
            print("   both error:", a["error"][:100], "|", b["error"][:100])
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)

"""Differential test of the search API against the reference: same random (model, world size, global batch, search lists) ->
same best strategy, same list of feasible candidates with the same MFU / memory numbers.

    python tools/fuzz_search_vs_reference.py [n_cases] [seed]
"""
import io
import json
import os
import random
import sys
import warnings
from contextlib import redirect_stdout

warnings.filterwarnings("ignore")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, REPO)
CFG = f"{REF}/configs"


def run(which, case):
    model, layers, world, gbs, tp, ep, pp, rc, mem_gbs, kind = case
    if which == "ref":
        sys.path.insert(0, REF)
        from simumax.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax_b200.core.perf_llm import PerfLLM
    moe = model in ("deepseekv2", "mixtral-8x7b", "deepseekv2-lite")
    st = StrategyConfig.init_from_config_file(f"{CFG}/strategy/{'ep8_pp1_dp8_mbs1' if moe else 'tp1_pp1_dp8_mbs1'}.json")
    p = PerfLLM()
    p._search_verbose = False
    buf = io.StringIO()
    with redirect_stdout(buf):
        p.configure(strategy_config=st, model_config=ModelConfig.init_from_config_file(f"{CFG}/models/{model}.json"),
                    system_config=SystemConfig.init_from_config_file(f"{CFG}/system/b200_bf16_ceperm.json"))
        p.model_config.layer_num = layers
        if moe:
            p.model_config.moe_pad_expert_input_to_capacity, p.model_config.capacity = True, 1
        p.system.accelerator.mem_gbs = mem_gbs
        p.run_estimate()
        if kind == "grid":
            every = {}
            best = p.search_best_parallel_strategy(world_size=world, gmi_error=4, micro_batch_size=1, global_batch_size=gbs,
                                                   all_search_result=every, tp_search_list=tp, ep_search_list=ep, pp_search_list=pp,
                                                   recompute_search_type=rc, use_reserved_memory=True, dump_path=None, verbose=False)
            every.pop("comm_bw_info", None)
            if isinstance(best, dict):
                best = {k: v for k, v in best.items() if k != "comm_bw_info"}
            return json.loads(json.dumps({"best": best, "all": every}, sort_keys=True, default=str))
        if kind == "mbs":
            return json.loads(json.dumps(p.search_max_micro_batch_size(micro_batch_num=pp[0] * 2), default=str))
        sizes, nums, peaks, costs = p.search_max_micro_batch_size_fixed_gbs(pp_size=p.strategy.pp_size, dp_size=p.strategy.dp_size,
                                                                          global_batch_size=gbs, gmi_error=4, verbose=False)
        return json.loads(json.dumps({"sizes": sizes, "nums": nums, "peaks": peaks,
                                      "tgs": [round(c.data["throughput_per_accelerator"], 4) for c in costs]}, default=str))


def close(a, b, rel=1e-9):
    """Structural equality with a relative tolerance on floats (summation order differs in the last ulp)."""
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(close(a[k], b[k], rel) for k in a)
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(close(x, y, rel) for x, y in zip(a, b))
    if isinstance(a, float) or isinstance(b, float):
        try:
            return abs(float(a) - float(b)) <= rel * max(abs(float(a)), abs(float(b)), 1e-300)
        except (TypeError, ValueError):
            return False
    return a == b


def sample(rng):
    moe = rng.random() < 0.35
    model = rng.choice(["deepseekv2", "mixtral-8x7b", "deepseekv2-lite"]) if moe else rng.choice(["llama3-8b", "llama3-70b", "llama2-13b", "qwen3-32b"])
    world = rng.choice([8, 16, 32])
    kind = rng.choice(["grid", "grid", "mbs", "fixed_gbs"])
    return (model, rng.choice([4, 8]), world, rng.choice([32, 64]), [1] if moe else rng.sample([1, 2, 4, 8], 2), rng.sample([2, 4, 8], 2) if moe else [1],
            rng.sample([1, 2, 4], 2), rng.sample(["no_recompute", "full_block", "selective_recompute"], rng.randint(1, 3)),
            rng.choice([24, 40, 80, 180]), kind)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    stats = {"identical": 0, "both_reject": 0, "diff": 0}
    for i in range(n):
        case = sample(rng)
        res = []
        for impl in ("ref", "new"):
            try:
                res.append(run(impl, case))
            except Exception as ex:
                res.append({"error": f"{type(ex).__name__}: {str(ex)[:100]}"})
        a, b = res
        if isinstance(a, dict) and isinstance(b, dict) and "error" in a and "error" in b:
            stats["both_reject"] += 1
        elif close(a, b):
            stats["identical"] += 1
        else:
            stats["diff"] += 1
            print("DIFF", i, json.dumps(case))
            print("   ref:", json.dumps(a)[:700])
            print("   new:", json.dumps(b)[:700])
    print(json.dumps(stats))
    return stats["diff"]


if __name__ == "__main__":
    sys.exit(1 if main() else 0)

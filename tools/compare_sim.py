"""Dev tool: compare simulate() end time / memory peaks with the reference simulator (CPU only)."""
import io, json, os, re, sys, tempfile, warnings
from contextlib import redirect_stdout
warnings.filterwarnings("ignore")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, REPO)
CFG = f"{REF}/configs"
MOE = {"moe_pad_expert_input_to_capacity": True, "capacity": 1}
CASES = [
    ("tp2_pp1_dp4_mbs1", "llama2-tiny", "a100_pcie", 2, {}, {}, True),
    ("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_bf16_ceperm", 4, {}, {}, True),
    ("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_bf16_ceperm", 4, {}, {"pp_comm_async": False}, True),
    ("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_bf16_ceperm", 8, {}, {"pp_size": 4, "micro_batch_num": 8, "pp_comm_async": False}, True),
    ("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_bf16_ceperm", 8, {}, {"pp_size": 4, "micro_batch_num": 8}, True),
    ("tp2_pp1_dp4_mbs1_selective_recompute", "llama3-70b", "b200_bf16_ceperm", 2, {}, {"recompute_layer_num": 2}, True),
    ("tp2_pp1_dp4_mbs1_full_recompute", "llama3-70b", "b200_bf16_ceperm", 2, {}, {"recompute_layer_num": 1}, True),
    ("ep8_pp1_dp8_mbs1", "deepseekv2", "b200_bf16_ceperm", 2, MOE, {}, True),
    ("ep4_pp2_dp4_mbs1", "deepseekv2", "b200_bf16_ceperm", 4, MOE, {"pp_comm_async": False}, True),
    ("tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt", "llama3-70b", "b200_bf16_ceperm", 8, {}, {}, True),
    ("tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt", "llama3-70b", "b200_bf16_ceperm", 8, {}, {"pp_comm_async": True}, True),
    ("tp2_pp1_dp4_mbs1", "llama2-tiny", "a100_pcie", 2, {}, {}, False),
    ("tp1_pp2_dp4_mbs1", "llama2-tiny", "a100_pcie", 2, {}, {"tp_size": 2, "micro_batch_num": 4}, False),
    # found by tools/fuzz_sim_vs_reference.py: a one-op recompute segment (sdp_only) is LAST, not FIRST -> caches in the replay
    ("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_bf16_ceperm", 4, {}, {"enable_recompute": True, "recompute_granularity": "sdp_only",
                                                                   "recompute_layer_num": 1, "micro_batch_num": 2}, True),
]


def run(which, case):
    strategy, model, system, layers, mover, sover, merge = case
    if which == "ref":
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
    buf = io.StringIO()
    d = tempfile.mkdtemp()
    with redirect_stdout(buf):
        p.configure(strategy_config=st, model_config=ModelConfig.init_from_config_file(f"{CFG}/models/{model}.json"),
                    system_config=SystemConfig.init_from_config_file(f"{CFG}/system/{system}.json"))
        p.model_config.layer_num = layers
        for k, v in mover.items():
            setattr(p.model_config, k, v)
        p.run_estimate()
        p.simulate(d, merge_lanes=merge)
    end = float(re.findall(r"end in ([0-9.eE+-]+)", buf.getvalue())[-1])
    mem = None
    mp = os.path.join(d, "simu_memory_result.json")
    if os.path.exists(mp):
        m = json.load(open(mp))
        mem = {"static": m["static_allocated_bytes_by_rank"], "peak": m["peak_allocated_bytes_by_rank"]}
    return {"end": round(end, 6), "mem": mem}


def main():
    bad = 0
    only = [int(x) for x in sys.argv[1:]]
    for i, case in enumerate(CASES):
        if only and i not in only:
            continue
        res = {}
        for which in ("ref", "new"):
            try:
                res[which] = run(which, case)
            except Exception as ex:
                import traceback
                if which == "new":
                    traceback.print_exc()
                res[which] = {"error": f"{type(ex).__name__}: {str(ex)[:150]}"}
        ok = res["ref"] == res["new"] or ("error" in res["ref"] and "error" in res["new"])
        bad += not ok
        print("OK  " if ok else "DIFF", i, case[0], case[1], case[5], "merge" if case[6] else "world")
        if not ok or "error" in res["ref"]:
            print("   ref:", json.dumps(res["ref"])[:400])
            print("   new:", json.dumps(res["new"])[:400])
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)

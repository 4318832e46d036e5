"""CPU-side speed of the analytical product against the unmodified reference on the reference's own configs: one `perf` (configure +
run_estimate + analysis), one `simulate`, one strategy search.  Best of n wall-clock runs each, same process, same machine.
    python tools/bench_cpu_vs_reference.py [out.json]        (needs the reference checkout: REF, default /root/reference)"""
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import time
import warnings

warnings.filterwarnings("ignore")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
C = f"{REF}/configs"
PERF_CASES = [("llama3-8b", "tp1_pp2_dp4_mbs1", None), ("llama3-70b", "tp2_pp1_dp4_mbs1", None), ("deepseekv2", "ep8_pp1_dp8_mbs1", None)]
SIM_CASES = [("llama3-8b", "tp1_pp2_dp4_mbs1", None), ("deepseekv2", "ep4_pp2_dp4_mbs1", 4)]


def _perf_llm(which):
    if which == "reference":
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200 import PerfLLM
    return PerfLLM


def _configured(which, model, strat, layers):
    p = _perf_llm(which)()
    p.configure(strategy_config=f"{C}/strategy/{strat}.json", model_config=f"{C}/models/{model}.json", system_config=f"{C}/system/b200_bf16_ceperm.json")
    if layers:
        p.model_config.layer_num = layers
    return p


def best_of(fn, n):
    best = float("inf")
    for _ in range(n):
        with contextlib.redirect_stdout(io.StringIO()):
            t = fn()
        best = min(best, t)
    return best


def perf_once(which, model, strat, layers):
    t = time.perf_counter()
    p = _configured(which, model, strat, layers)
    p.run_estimate()
    p.analysis_cost()
    p.analysis_mem()
    return time.perf_counter() - t


def sim_once(which, model, strat, layers):
    p = _configured(which, model, strat, layers)
    p.run_estimate()
    d = tempfile.mkdtemp()
    t = time.perf_counter()
    p.simulate(d)
    dt = time.perf_counter() - t
    shutil.rmtree(d, ignore_errors=True)
    return dt


def search_once(which):
    p = _configured(which, "llama3-8b", "tp1_pp1_dp8_mbs1", None)
    p._search_verbose = False
    p.run_estimate()
    t = time.perf_counter()
    p.search_best_parallel_strategy(world_size=64, gmi_error=4, micro_batch_size=1, global_batch_size=256, all_search_result={},
                                    tp_search_list=[1, 2, 4, 8], ep_search_list=[1], pp_search_list=[1, 2, 4],
                                    recompute_search_type=["no_recompute", "full_block"], use_reserved_memory=True, dump_path=None, verbose=False)
    return time.perf_counter() - t


def main():
    out = {"perf_ms": {}, "simulate_s": {}, "search_s": {}}
    for which in ("simumax_b200", "reference"):
        for m, s, l in PERF_CASES:
            out["perf_ms"].setdefault(f"{m} {s}", {})[which] = round(best_of(lambda: perf_once(which, m, s, l), 5) * 1e3, 1)
        for m, s, l in SIM_CASES:
            out["simulate_s"].setdefault(f"{m} {s}", {})[which] = round(best_of(lambda: sim_once(which, m, s, l), 2), 3)
        try:
            out["search_s"].setdefault("llama3-8b world 64 gbs 256", {})[which] = round(best_of(lambda: search_once(which), 1), 2)
        except Exception as e:                                    # signature drift between versions: report, do not fail the rest
            out["search_s"].setdefault("llama3-8b world 64 gbs 256", {})[which] = f"{type(e).__name__}: {e}"[:120]
    for sect in out.values():
        for row in sect.values():
            a, b = row.get("simumax_b200"), row.get("reference")
            if isinstance(a, (int, float)) and isinstance(b, (int, float)) and a > 0:
                row["speedup"] = round(b / a, 2)
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()

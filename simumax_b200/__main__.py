"""Command line front-end:  ``python -m simumax_b200 <command> ...``

    list                                         models / strategy presets / system configs that ship with the package
    perf      -m llama3-8b -s tp2_pp1_dp4_mbs1   one estimate: headline numbers (+ --out DIR for the full result files)
    simulate  -m ... -s ... --out DIR            discrete-event replay: Chrome trace, event log, memory snapshots
    search    -m ... --world 64 --gbs 256        best tp x ep x pp (x recompute) under the memory budget
    trace     LOG OUT.json                       text event log -> Chrome trace

Common options: ``--system`` (default: the calibrated B200 config), ``--layers N``, and any number of ``--set key=value``
overrides applied to the strategy (or, with a ``model.`` prefix, to the model), e.g. ``--set tp_size=4 --set model.layer_num=12``.
(The reference offers example scripts only; same API underneath.)
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import sys
import warnings


def _value(text: str):
    for conv in (int, float):
        try:
            return conv(text)
        except ValueError:
            pass
    low = text.lower()
    if low in ("true", "false"):
        return low == "true"
    if low in ("none", "null"):
        return None
    if text.startswith("[") or text.startswith("{"):
        return json.loads(text)
    return text


def _build(a):
    from . import PerfLLM
    from .configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
    from .core.config import ModelConfig, StrategyConfig, SystemConfig
    warnings.filterwarnings("ignore")
    st = StrategyConfig.init_from_config_file(a.strategy if a.strategy.endswith(".json") else get_simu_strategy_config(a.strategy))
    mo = ModelConfig.init_from_config_file(a.model if a.model.endswith(".json") else get_simu_model_config(a.model))
    sy = SystemConfig.init_from_config_file(a.system if a.system.endswith(".json") else get_simu_system_config(a.system))
    if getattr(a, "runtime_semantics", False):
        st.apply_runtime_semantics()        # price this repo's runtime instead of Megatron-LM + NCCL (core/config.py)
    for item in a.set or []:
        key, _, val = item.partition("=")
        target, name = (mo, key[len("model."):]) if key.startswith("model.") else (st, key)
        if not hasattr(target, name):
            raise SystemExit(f"unknown field '{key}'")
        setattr(target, name, _value(val))
    if a.layers:
        mo.layer_num = a.layers
    perf = PerfLLM()
    perf._search_verbose = False
    with contextlib.redirect_stdout(io.StringIO()):
        perf.configure(strategy_config=st, model_config=mo, system_config=sy)
    return perf


def _headline(perf) -> dict:
    with contextlib.redirect_stdout(io.StringIO()):
        perf.run_estimate()
        cost, mem = perf.analysis_cost().data, perf.analysis_mem().data
    stages = mem if "peak_mem" not in mem else {"stage": mem}
    return {"model": perf.model_config.model_name, "parallelism": perf.strategy.parallelism, "system": perf.system.sys_name,
            "iteration_time": cost["duration_time_per_iter"], "tokens_per_s_per_gpu": round(cost["throughput_per_accelerator"], 2),
            "mfu": round(cost["mfu"], 4), "breakdown": cost["breakdown_result"],
            "peak_mem": {k: v["peak_mem"] for k, v in stages.items() if isinstance(v, dict) and "peak_mem" in v},
            "efficiency_misses": {k: len(v) for k, v in perf.system.miss_efficiency.items()}}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m simumax_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)

    def common(p):
        p.add_argument("-m", "--model", default="llama3-8b")
        p.add_argument("-s", "--strategy", default="tp1_pp1_dp8_mbs1")
        p.add_argument("--system", default="b200_sm100a_calibrated")
        p.add_argument("--layers", type=int, default=0)
        p.add_argument("--set", action="append", metavar="key=value")
        p.add_argument("--runtime-semantics", action="store_true",
                       help="predict simumax_b200.runtime (fused kernels) instead of Megatron-LM + NCCL semantics")
        p.add_argument("--out", default=None)
    sub.add_parser("list")
    common(sub.add_parser("perf"))
    sp = sub.add_parser("simulate")
    common(sp)
    sp.add_argument("--world-lanes", action="store_true", help="one simulated lane per rank instead of one per pipeline stage")
    ss = sub.add_parser("search")
    common(ss)
    ss.add_argument("--world", type=int, default=8)
    ss.add_argument("--gbs", type=int, default=64)
    ss.add_argument("--mbs", type=int, default=1)
    ss.add_argument("--margin-gib", type=int, default=6, help="memory head-room kept free on every device")
    ss.add_argument("--tp", type=int, nargs="*", default=None)
    ss.add_argument("--ep", type=int, nargs="*", default=None)
    ss.add_argument("--pp", type=int, nargs="*", default=None)
    ss.add_argument("--recompute", nargs="*", default=["no_recompute", "full_block", "selective_recompute"])
    st = sub.add_parser("trace")
    st.add_argument("log")
    st.add_argument("out_json")
    a = ap.parse_args(argv)

    if a.cmd == "list":
        from .configs import list_models, list_strategies, list_systems
        print(json.dumps({"models": list_models(), "strategies": list_strategies(), "systems": list_systems()}, indent=1))
        return 0
    if a.cmd == "trace":
        from .sim.trace import process_log_file
        print(f"{process_log_file(a.log, a.out_json)} trace records -> {a.out_json}")
        return 0
    perf = _build(a)
    if a.cmd == "perf":
        print(json.dumps(_headline(perf), indent=1, default=str))
        if a.out:
            with contextlib.redirect_stdout(io.StringIO()):
                perf.analysis(a.out, console_log=False)
            print(f"result files -> {a.out}")
        return 0
    if a.cmd == "simulate":
        if not a.out:
            raise SystemExit("simulate needs --out DIR")
        head = _headline(perf)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = perf.simulate(a.out, merge_lanes=not a.world_lanes)
        print(json.dumps({"analytic_iteration_time": head["iteration_time"], "simulated_end_ms": res["end_time_ms"],
                          "events": res["num_events"], "lanes": len(res["ranks"]), "out": a.out}, indent=1))
        return 0
    if a.cmd == "search":
        everything = {}
        with contextlib.redirect_stdout(io.StringIO()):
            perf.run_estimate()
            best = perf.search_best_parallel_strategy(
                world_size=a.world, gmi_error=a.margin_gib, micro_batch_size=a.mbs, global_batch_size=a.gbs, all_search_result=everything,
                tp_search_list=a.tp, ep_search_list=a.ep, pp_search_list=a.pp, recompute_search_type=a.recompute, dump_path=a.out,
                verbose=False)
        n = len(next(iter(everything.values()))) if everything else 0
        print(json.dumps({"candidates_that_fit": n, "best": best, "cache": getattr(perf, "last_search_cache_stats", None)},
                         indent=1, default=str))
        return 0
    return 1


if __name__ == "__main__":
    sys.exit(main())

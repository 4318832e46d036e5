"""Differential checks against the unmodified reference (skipped when no reference checkout is available): a few hand-picked
triples plus a small random sample from the fuzzers in tools/ -- the long runs (hundreds of cases) are run by hand."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "simumax", "core")), reason="reference checkout not available")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _both(run, case, *extra):
    out = []
    for impl in ("ref", "new"):
        try:
            out.append(run(impl, case, *extra))
        except Exception as ex:                     # both implementations must reject the same configurations
            out.append({"error": type(ex).__name__})
    return out


@pytest.mark.timeout(600)
def test_hand_picked_triples_identical():
    import compare_with_reference as cw
    for i in (0, 4, 12, 16, 19):
        a, b = _both(cw.run, cw.CASES[i], "b200_bf16_ceperm")
        assert a == b or ("error" in a and "error" in b), (i, a, b)


@pytest.mark.timeout(900)
def test_random_triples_identical():
    import compare_with_reference as cw
    import fuzz_vs_reference as fz
    rng = random.Random(2026)
    compared = 0
    for _ in range(16):
        case = fz.sample(rng)
        a, b = _both(cw.run, case, "b200_bf16_ceperm")
        assert a == b or ("error" in a and "error" in b), (case, a, b)
        compared += "error" not in a
    assert compared >= 2          # the sampler also produces configurations both implementations reject


@pytest.mark.timeout(900)
def test_random_simulations_identical():
    import compare_sim as cs
    import fuzz_sim_vs_reference as fs
    rng = random.Random(7)
    for _ in range(4):
        case = fs.sample(rng)
        a, b = _both(cs.run, case)
        assert a == b or ("error" in a and "error" in b), (case, a, b)


@pytest.mark.timeout(900)
def test_trace_records_identical(tmp_path):
    """Every Chrome-trace record (lanes, envelopes, [gN] tags, flow arrows, counters, metadata) of three schedules: tp2,
    async 1F1B and blocking 1F1B -- only the documented wait-lane extras may differ (tools/deep_compare_trace.py)."""
    import compare_sim as cs
    import deep_compare_trace as dt
    for i in (0, 1, 2):
        c, examples = dt.compare(dt.run("ref", cs.CASES[i]), dt.run("new", cs.CASES[i]))
        assert c["diff"] == 0 and c["time"] == 0 and c["dropped"] == 0, (i, dict(c), examples[:4])
        assert c["same"] > 900


@pytest.mark.timeout(900)
def test_reference_log_converts_to_reference_trace(tmp_path):
    """``process_log_file`` on a log written by the reference reproduces the reference's own tracing_logs.json field by field
    (blocking interleaved schedule: batch scopes, pp_detail sub-lanes, pair arrows)."""
    import io
    import json
    from contextlib import redirect_stdout

    import compare_sim as cs
    from simumax_b200.sim.trace import process_log_file
    sys.path.insert(0, REF)
    from simumax.core.config import ModelConfig, StrategyConfig, SystemConfig
    from simumax.core.perf_llm import PerfLLM
    p = PerfLLM()
    with redirect_stdout(io.StringIO()):
        p.configure(strategy_config=StrategyConfig.init_from_config_file(f"{cs.CFG}/strategy/tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt.json"),
                    model_config=ModelConfig.init_from_config_file(f"{cs.CFG}/models/llama3-8b.json"),
                    system_config=SystemConfig.init_from_config_file(f"{cs.CFG}/system/b200_bf16_ceperm.json"))
        p.model_config.layer_num = 8
        p.run_estimate()
        p.simulate(str(tmp_path), merge_lanes=True)
    process_log_file(str(tmp_path / "log.log"), str(tmp_path / "ours.json"))
    ref = [r for r in json.load(open(tmp_path / "tracing_logs.json")) if r.get("ph") != "C"]
    ours = json.load(open(tmp_path / "ours.json"))
    assert len(ref) == len(ours)

    def split(r):
        r = json.loads(json.dumps(r))
        times = (r.pop("ts", 0.0), r.pop("dur", 0.0), (r.get("args") or {}).pop("post_ts", None) or 0.0)
        r.pop("id", None)                         # record ids number the records / arrows in file order
        return json.dumps(r, sort_keys=True), times
    a, b = sorted(map(split, ref)), sorted(map(split, ours))
    for (ka, ta), (kb, tb) in zip(a, b):
        assert ka == kb
        assert ta == pytest.approx(tb, abs=5e-3)      # microseconds; the reference goes through a 6-digit text log


@pytest.mark.timeout(600)
def test_model_arch_text_identical():
    """model_arch / base_info["arch"]: the nested module view with per-module time, FLOP and static-memory totals, byte for
    byte -- dense tp2, MLA + MoE with pipeline stages, fp8 wrappers with full recompute."""
    import compare_with_reference as cw
    import deep_compare_with_reference_lib as dl
    for i in (0, 4, 20):
        case = cw.CASES[i]
        system = case[5] if len(case) > 5 else "b200_bf16_ceperm"
        a, b = dl.run("ref", case, system), dl.run("new", case, system)
        assert a["arch"] == b["arch"], i
        assert [d for d in dl.diff(a, b) if not d[0].endswith("/raw")] == []


@pytest.mark.timeout(600)
def test_memory_files_identical():
    """simu_memory_snapshot.json, simu_memory_result.json and the torch memory_viz pickle (segment size, every alloc / free
    trace entry with its frames, the final block list) of a recompute run and a pipeline run."""
    import compare_sim as cs
    import deep_compare_trace as dt
    for i in (5, 2):
        dt.run("ref", cs.CASES[i])
        dt.run("new", cs.CASES[i])
        assert dt.memory_files_differ() == [], i


@pytest.mark.timeout(600)
def test_model_graph_json_identical(tmp_path):
    """model_graph.json of run_estimate(capture_graph=True): node names, tensor wiring (MLA split / rope / concat fan-in, MoE
    router -> permutation -> un-permutation -> shared-expert add), call_idx and value_info; this repo's extras are ignored."""
    import io
    import json
    from contextlib import redirect_stdout

    import compare_sim as cs

    def run(which, strategy, model):
        if which == "ref":
            sys.path.insert(0, REF)
            from simumax.core.config import ModelConfig, StrategyConfig, SystemConfig
            from simumax.core.perf_llm import PerfLLM
        else:
            from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig
            from simumax_b200.core.perf_llm import PerfLLM
        p = PerfLLM()
        d = tmp_path / f"{which}-{model}"
        d.mkdir()
        with redirect_stdout(io.StringIO()):
            p.configure(strategy_config=StrategyConfig.init_from_config_file(f"{cs.CFG}/strategy/{strategy}.json"),
                        model_config=ModelConfig.init_from_config_file(f"{cs.CFG}/models/{model}.json"),
                        system_config=SystemConfig.init_from_config_file(f"{cs.CFG}/system/b200_bf16_ceperm.json"))
            p.model_config.layer_num = 2
            p.run_estimate(capture_graph=True, save_path=str(d))
        g = json.load(open(d / "model_graph.json"))
        g.pop("edges", None)
        for n in g["nodes"]:
            n["attributes"] = {}
        return g
    for strategy, model in (("tp2_pp1_dp4_mbs1", "llama3-8b"), ("ep4_pp2_dp4_mbs1", "deepseekv2"), ("tp1_pp2_dp4_mbs1", "gpt3-175b")):
        assert run("ref", strategy, model) == run("new", strategy, model), model

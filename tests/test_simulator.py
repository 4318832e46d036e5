import json
import re
import pickle

import pytest

from helpers import MOE, build, reference_system


def test_tiny_llama_trace_golden(tmp_path):
    """llama2-tiny, 2 layers, tp2 on the reference A100 config: end 687.734 ms, 152 cache tokens, known peaks."""
    p = build("tp2_pp1_dp4_mbs1", "llama2-tiny", reference_system("a100_pcie"), 2)
    p.model_config.padded_vocab_size = True
    r = p.simulate(str(tmp_path), merge_lanes=True)
    assert r["end_time_ms"] == pytest.approx(687.734, abs=1e-3)
    assert r["num_events"] == 981
    assert r["memory"]["static_allocated_bytes_by_rank"]["rank0"] == 3001208832
    assert r["memory"]["peak_allocated_bytes_by_rank"]["rank0"] == 3967209472
    assert r["memory"]["peak_op_by_rank"]["rank0"] == {"op_name": "self.parallel_ce", "phase": "fwd"}
    assert r["cache_tokens"] == 152
    trace = json.load(open(tmp_path / "tracing_logs.json"))
    assert any(e.get("ph") == "C" for e in trace)
    # viewer layout of the reference (generate_tracing.py:65-104): ordered lanes, envelopes on the scope lanes
    x = [e for e in trace if e.get("ph") == "X"]
    assert {e["tid"] for e in x} == {"00_fwd_scope", "01_bwd_scope", "02_fwd_compute", "03_bwd_compute", "08_comm"}
    assert all(e["cat"] == "scope" for e in x if e["tid"].endswith("_scope"))
    emb = next(e for e in x if e["name"] == "LanguageModelEmbedding_0")
    assert emb["args"]["call_stack"] == ["microbatch0", "GPTModel_first_pp_stage", "LanguageModelEmbedding_0"]
    names = {(m["name"], m.get("tid")) for m in trace if m.get("ph") == "M"}
    assert ("thread_sort_index", "08_comm") in names and ("process_sort_index", None) in names
    snap = json.load(open(tmp_path / "simu_memory_snapshot.json"))
    assert snap["schema"] == "simumax_memory_snapshot_v1"
    allocs = [e for e in snap["cache_tokens"] if e["action"] == "alloc"]
    frees = [e for e in snap["cache_tokens"] if e["action"] == "free"]
    assert len(allocs) == len(frees) == 152
    viz = pickle.load(open(tmp_path / "simu_memory_viz_snapshot.pickle", "rb"))
    assert "device_traces" in viz and "segments" in viz and viz["segments"][0]["total_size"] >= 3967209472


def test_trace_layout_pipeline_lanes(tmp_path):
    """Async p2p lands on pp_fwd / pp_bwd with a [gN] tag + one flow arrow per pair; blocking p2p is packed into
    non-overlapping pp_detail_k lanes under a pp_batch_scope envelope (generate_tracing.py:232-262, :370-420)."""
    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_nominal", 4)
    p.simulate(str(tmp_path / "a"), merge_lanes=True)
    tr = json.load(open(tmp_path / "a" / "tracing_logs.json"))
    x = [e for e in tr if e.get("ph") == "X"]
    p2p = [e for e in x if e["tid"] in ("04_pp_fwd", "05_pp_bwd")]
    assert p2p and all(re.search(r"\[g\d+\]$", e["name"]) for e in p2p)
    assert all(e["args"]["gid"].startswith("fwd:send_recv-") and e["args"]["post_ts"] is not None for e in p2p)
    flows = [e for e in tr if e.get("cat") == "comm_pair"]
    assert len(flows) == len(p2p) and {e["ph"] for e in flows} == {"s", "f"}
    by_id = {}
    for f in flows:
        by_id.setdefault(f["id"], []).append(f)
    assert all(len(v) == 2 and v[0]["pid"] != v[1]["pid"] and v[0]["name"] == v[1]["name"] for v in by_id.values())
    assert any(e["tid"] == "09_wait" for e in x)

    q = build("tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt", "llama3-70b", "b200_nominal", 8)
    q.simulate(str(tmp_path / "b"), merge_lanes=True)
    tr = json.load(open(tmp_path / "b" / "tracing_logs.json"))
    x = [e for e in tr if e.get("ph") == "X"]
    assert any(e["tid"] == "06_pp_batch_scope" and e["cat"] == "scope" for e in x)
    detail = [e for e in x if e["tid"].startswith("07_pp_detail_")]
    assert detail and len({e["tid"] for e in detail}) >= 2
    lanes = {}
    for e in detail:
        lanes.setdefault((e["pid"], e["tid"]), []).append((e["ts"], e["ts"] + e["dur"]))
    for spans in lanes.values():                      # no two slices of one sub-lane overlap
        spans.sort()
        assert all(b[0] >= a[1] - 1e-6 for a, b in zip(spans, spans[1:]))


CASES = [("tp1_pp2_dp4_mbs1", "llama3-8b", 4, {}, {}),
         ("tp1_pp2_dp4_mbs1", "llama3-8b", 4, {}, {"pp_comm_async": False}),
         ("tp1_pp2_dp4_mbs1", "llama3-8b", 8, {}, {"pp_size": 4, "micro_batch_num": 6}),
         ("tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt", "llama3-70b", 8, {}, {}),
         ("tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt", "llama3-70b", 8, {}, {"pp_comm_async": True}),
         ("ep4_pp2_dp4_mbs1_full_recompute", "deepseekv2", 4, MOE, {"pp_comm_async": False}),
         ("tp2_pp1_dp4_mbs1_selective_recompute", "llama3-70b", 2, {}, {"recompute_layer_num": 2})]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[1]}-{i}" for i, c in enumerate(CASES)])
def test_simulation_consistent_with_analytic_model(case, tmp_path):
    strategy, model, layers, mover, sover = case
    p = build(strategy, model, "b200_nominal", layers, mover, enable_straggler_model=False, **sover)
    r = p.simulate(str(tmp_path), merge_lanes=True)
    assert r["ranks"] == [i * (p.strategy.world_size // p.strategy.pp_size) for i in range(p.strategy.pp_size)]
    if not (p._vp_size() > 1 and p.strategy.pp_comm_async):
        analytic = p.analysis_cost().data["raw"]["duration_time_per_iter_ms"]
        # the replay adds 2 us per module scope and a 1 ms world barrier; otherwise both models must agree
        assert r["end_time_ms"] == pytest.approx(analytic + 1.0, rel=0.03)
    if "memory" in r:
        mem = p.analysis_mem().data
        stages = [mem] if "raw" in mem else list(mem.values())
        for rank, stage in zip(sorted(r["memory"]["peak_allocated_bytes_by_rank"]), stages):
            sim_peak = r["memory"]["peak_allocated_bytes_by_rank"][rank]
            assert sim_peak == pytest.approx(stage["raw"]["peak_mem"], rel=0.02)
        snap = json.load(open(tmp_path / "simu_memory_snapshot.json"))
        last = {}
        for e in snap["events"]:
            last[e["rank"]] = e
        assert all(e["cached_bytes"] == 0 for e in last.values())     # every cached activation was released


def test_world_mode_rendezvous(tmp_path):
    p = build("tp2_pp1_dp4_mbs1", "llama2-tiny", "b200_nominal", 2, micro_batch_num=2)
    merged = p.simulate(str(tmp_path / "m"), merge_lanes=True)["end_time_ms"]
    world = p.simulate(str(tmp_path / "w"), merge_lanes=False)
    assert len(world["ranks"]) == 8
    assert world["end_time_ms"] == pytest.approx(merged, rel=1e-6)


def test_deadlock_is_detected():
    from simumax_b200.sim.engine import BlockingP2P, Deadlock, Engine, Scope
    eng = Engine()
    eng.add_rank(0, [("fwd", Scope("a", [BlockingP2P("send", "x", 1.0)]))])
    eng.add_rank(1, [("fwd", Scope("b", [BlockingP2P("recv", "y", 1.0)]))])
    with pytest.raises(Deadlock):
        eng.run()


def test_log_round_trip_rebuilds_the_trace(tmp_path):
    """log.log -> process_log_file reproduces the structured trace (reference path: generate_tracing.process_log_file)."""
    import collections
    import contextlib
    import io
    import json

    from simumax_b200 import PerfLLM
    from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
    from simumax_b200.sim.trace import parse_log_line, process_log_file
    p = PerfLLM()
    with contextlib.redirect_stdout(io.StringIO()):
        p.configure(strategy_config=get_simu_strategy_config("tp1_pp2_dp4_mbs1"), model_config=get_simu_model_config("llama3-8b"),
                    system_config=get_simu_system_config("b200_nominal"))
        p.model_config.layer_num = 4
        p.run_estimate()
        p.simulate(str(tmp_path))
    process_log_file(str(tmp_path / "log.log"), str(tmp_path / "from_log.json"))
    a = [r for r in json.load(open(tmp_path / "tracing_logs.json")) if r.get("ph") in ("X", "s", "f")]
    b = [r for r in json.load(open(tmp_path / "from_log.json")) if r.get("ph") in ("X", "s", "f")]
    key = lambda r: (r["ph"], r["tid"], r["pid"], r["name"])
    assert len(a) == len(b) and collections.Counter(map(key, a)) == collections.Counter(map(key, b))
    assert all(abs(x["ts"] - y["ts"]) < 1e-2 for x, y in zip(sorted(a, key=lambda r: (key(r), r["ts"])), sorted(b, key=lambda r: (key(r), r["ts"]))))
    # a reference-made line (no kind / lane suffix) is classified by its leaf name
    e = parse_log_line("rank4-async_wait_recv_prev gid send_recv-0-1-forward-0-pp_group:tp:0- fwd cost 30.805182 st 0.000000 ed 30.805182")
    assert (e.kind, e.rank, e.gid) == ("wait", 4, "send_recv-0-1-forward-0-pp_group:tp:0-")


def test_async_vpp_perf_via_simulation_extension():
    """perf() of asynchronous interleaved pipelining: the reference raises (perf_llm.py:2697-2701); with the opt-in knob the
    schedule time comes from the event engine and agrees with the full simulate() run and with the synchronous closed form."""
    import contextlib
    import io
    import re

    import pytest
    from helpers import build
    name = "tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt"
    sync = build(name, "llama3-70b", "b200_nominal", 8)
    t_sync = float(str(sync.analysis_cost().data["duration_time_per_iter"]).split()[0])
    plain = build(name, "llama3-70b", "b200_nominal", 8, pp_comm_async=True)
    with pytest.raises(RuntimeError, match="async VPP"):
        plain.analysis_cost()
    ext = build(name, "llama3-70b", "b200_nominal", 8, pp_comm_async=True, vpp_async_perf_via_simulation=True)
    cost = ext.analysis_cost().data
    t_async = float(str(cost["duration_time_per_iter"]).split()[0])
    assert abs(t_async - t_sync) / t_sync < 0.05 and cost["mfu"] > 0
    buf = io.StringIO()
    import tempfile
    with contextlib.redirect_stdout(buf):
        ext.simulate(tempfile.mkdtemp())
    end = float(re.findall(r"end in ([0-9.eE+-]+)", buf.getvalue())[-1])
    assert 0 <= end - t_async < 5.0          # the full replay adds the optimizer-side world barrier, nothing else
    total, records = ext._compute_pp_total_time(return_schedules=True)
    assert len(records) == 4 and all(any(r["kind"] == "F" for r in rows) and any(r["kind"] == "B" for r in rows) for rows in records)


def test_memory_timeline_under_async_pipeline_extension(tmp_path):
    """The reference writes no memory artifacts when pipeline p2p is asynchronous; opt-in here, and the simulated per-stage peaks
    equal the analytic ones."""
    import contextlib
    import io

    from helpers import build
    from simumax_b200.core.utils import parse_bytes
    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_nominal", 8, pp_size=4, micro_batch_num=8, pp_comm_async=True)
    with contextlib.redirect_stdout(io.StringIO()):
        plain = p.simulate(str(tmp_path / "a"))
        ext = p.simulate(str(tmp_path / "b"), memory_timeline=True)
    assert "memory" not in plain and not (tmp_path / "a" / "simu_memory_result.json").exists()
    assert (tmp_path / "b" / "simu_memory_viz_snapshot.pickle").exists()
    sim = sorted(ext["memory"]["peak_allocated_bytes_by_rank"].values())
    mem = p.analysis_mem().data
    ana = {k: parse_bytes(v["peak_mem"]) for k, v in mem.items() if isinstance(v, dict) and "peak_mem" in v}
    assert abs(max(sim) - max(ana.values())) < 1e-3 * max(sim)
    assert abs(plain["end_time_ms"] - ext["end_time_ms"]) < 1e-9


def test_fault_injection_extension(tmp_path):
    """simulate(jitter=..., slow_ranks=...): deterministic by default, reproducible per seed; one slow pipeline stage slows the
    whole job, and with real rendezvous (one lane per rank) one slow tensor-parallel rank drags its group along."""
    import contextlib
    import io

    from helpers import build

    def end(p, **kw):
        with contextlib.redirect_stdout(io.StringIO()):
            return p.simulate(str(tmp_path / "s"), **kw)["end_time_ms"]

    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_nominal", 4, pp_size=4, micro_batch_num=8)
    base = end(p)
    assert end(p) == base
    noisy = end(p, jitter=0.1, seed=1)
    assert noisy != base and end(p, jitter=0.1, seed=1) == noisy and end(p, jitter=0.1, seed=2) != noisy
    assert abs(noisy - base) / base < 0.2
    lanes = sorted(r for r in [0, 2, 4, 6])            # representative ranks of the four stages at tp1 / dp2
    slow_last = end(p, slow_ranks={lanes[-1]: 1.5})     # the last stage (LM head + loss) is the critical one here
    slow_mid = end(p, slow_ranks={lanes[1]: 1.5})
    assert slow_last > 1.15 * base and base <= slow_mid < slow_last
    q = build("tp2_pp1_dp4_mbs1", "llama2-tiny", "a100_pcie", 2, micro_batch_num=2)
    t0 = end(q, merge_lanes=False)
    t1 = end(q, merge_lanes=False, slow_ranks={1: 2.0})
    assert t1 > 1.15 * t0

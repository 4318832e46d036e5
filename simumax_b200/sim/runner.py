"""Simulation driver: build one program per simulated rank, run the engine, export artifacts.

``merge_lanes=True`` simulates one representative rank per pipeline stage (tp = cp = dp = 0; intra-stage
collectives complete locally); ``False`` simulates every rank of the job with real rendezvous.
Outputs in ``save_path``: ``tracing_logs.json`` (Chrome trace incl. memory counters), ``log.log`` (text event
list), ``simu_memory_result.json``, ``simu_memory_snapshot.json`` (schema simumax_memory_snapshot_v1) and
``simu_memory_viz_snapshot.pickle``.   (reference: core/simu_runner.py, simu_artifacts.py)
"""
from __future__ import annotations

import json
import os
import pickle
import time

from ..core.utils import get_pp_stage_representative_rank
from .engine import Engine
from .jobs import RankArgs
from .memory import SimuMemoryTracker
from .schedule import build_1f1b_program, build_interleaved_program, build_optimizer_tail
from .trace import write_event_log, write_trace

FIRST, MIDDLE, LAST = "first_stage_chunk", "middle_stage_chunk", "last_stage_chunk"


def should_enable_memory_timeline(strategy) -> bool:
    return strategy.pp_size == 1 or (not getattr(strategy, "pp_comm_async", True))


def _build_engine(perf, merge_lanes: bool, tracker, with_optimizer: bool, **engine_kw) -> Engine:
    st = perf.strategy
    engine = Engine(merge_lanes=merge_lanes, memory_tracker=tracker, **engine_kw)
    n_ranks = st.pp_size if merge_lanes else st.world_size
    vp = perf._vp_size()
    for i in range(n_ranks):
        rank = get_pp_stage_representative_rank(i, st) if merge_lanes else i
        args = RankArgs(rank, st)
        pp_rank = args.info["pp_rank"]
        key = FIRST if pp_rank == 0 else (LAST if pp_rank == st.pp_size - 1 else MIDDLE)
        if vp > 1 and perf.vpp_stage_chunk_names.get(key):
            chunks = [perf.vpp_chunk_dict[n] for n in perf.vpp_stage_chunk_names[key]]
        else:
            chunks = [perf.model_chunk_dict[key]]
        if tracker is not None:
            tracker.init_rank(rank, sum(c.get_model_info().all for c in chunks))
        with_mem = tracker is not None
        if len(chunks) > 1:
            jobs = build_interleaved_program(perf, chunks, args, with_mem)
        else:
            jobs = build_1f1b_program(perf, chunks[0], args, with_mem)
        if with_optimizer:
            jobs.append(build_optimizer_tail(perf, key, args))
        engine.add_rank(rank, jobs)
    return engine


def simulate_schedule(perf, return_schedules: bool = False):
    """Pipeline-schedule time (ms, no optimizer / data-parallel tail) from the discrete-event engine, one lane per stage.

    Used by ``PerfLLM`` for schedules the closed-form evaluators do not cover (asynchronous interleaved pipelining: the
    reference raises in ``perf()`` and offers ``simulate()`` only, perf_llm.py:2697-2701).  With ``return_schedules`` the compute
    and p2p events come back as per-stage records in the format ``core/trace_export.py`` draws."""
    engine = _build_engine(perf, merge_lanes=True, tracker=None, with_optimizer=False)
    total = engine.run()
    if not return_schedules:
        return total
    order = {r: i for i, r in enumerate(sorted(engine.ranks))}
    records = [[] for _ in order]
    for e in engine.events:
        if e.kind == "compute":
            kind = "B" if e.phase == "bwd" else "F"
        elif e.kind == "p2p":
            kind = "send" if "send" in e.name else "recv"
        else:
            continue
        records[order[e.rank]].append({"kind": kind, "label": e.name.split("-")[-1], "start": e.start, "end": e.end,
                                       "duration": e.end - e.start, "mb": None, "gid": e.gid})
    return total, records


def run_simulation(perf, save_path, merge_lanes: bool = True, quiet: bool = False, memory_timeline="auto", jitter: float = 0.0,
                   slow_ranks=None, seed: int = 0) -> dict:
    st = perf.strategy
    t0 = time.time()
    os.makedirs(save_path, exist_ok=True)
    want_mem = should_enable_memory_timeline(st) if memory_timeline == "auto" else bool(memory_timeline)
    tracker = SimuMemoryTracker() if want_mem else None
    engine = _build_engine(perf, merge_lanes, tracker, with_optimizer=True, jitter=jitter, slow_ranks=slow_ranks, seed=seed)
    end_t = engine.run()
    if not quiet:
        print(f"end in {end_t}")
        print("wall time", time.time() - t0)
    write_event_log(engine.events, os.path.join(save_path, "log.log"))
    n_events = write_trace(engine.events, os.path.join(save_path, "tracing_logs.json"),
                           tracker.events if tracker is not None else None)
    result = {"end_time_ms": end_t, "num_events": len(engine.events), "num_trace_events": n_events,
              "ranks": sorted(engine.ranks)}
    if tracker is not None:
        with open(os.path.join(save_path, "simu_memory_result.json"), "w", encoding="utf-8") as fh:
            json.dump(tracker.summary(), fh, indent=4)
        with open(os.path.join(save_path, "simu_memory_snapshot.json"), "w", encoding="utf-8") as fh:
            json.dump(tracker.snapshot(), fh, indent=4)
        with open(os.path.join(save_path, "simu_memory_viz_snapshot.pickle"), "wb") as fh:
            pickle.dump(tracker.memory_viz_snapshot(), fh)
        result["memory"] = tracker.summary()
        result["cache_tokens"] = sum(1 for e in tracker.cache_token_events if e["action"] == "alloc")
    return result

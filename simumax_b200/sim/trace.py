"""Structured simulator events -> Chrome trace (``tracing_logs.json``), no text-log round trip.

Lane vocabulary matches the reference viewer layout (generate_tracing.py:65-80): per rank (pid ``rankN``) the
threads ``fwd_scope, bwd_scope, fwd_compute, bwd_compute, pp_fwd, pp_bwd, pp_batch_scope, comm, wait``;
times are microseconds; matching async send/recv pairs are linked with flow arrows.
"""
from __future__ import annotations

import json
from typing import Dict, List

from .engine import Event

LANES = ["fwd_scope", "bwd_scope", "fwd_compute", "bwd_compute", "pp_fwd", "pp_bwd", "pp_batch_scope", "comm", "wait"]
_SORT = {name: i for i, name in enumerate(LANES)}


def _lane(ev: Event) -> str:
    if ev.kind == "scope":
        if "batch_pp_comm" in ev.name:
            return "pp_batch_scope"
        return "bwd_scope" if ev.phase == "bwd" else "fwd_scope"
    if ev.kind == "compute":
        return "bwd_compute" if ev.phase == "bwd" else "fwd_compute"
    if ev.kind == "p2p":
        if ev.lane in ("pp_fwd", "pp_bwd"):
            return ev.lane
        return "pp_bwd" if "backward" in (ev.gid or "") else "pp_fwd"
    if ev.kind == "wait":
        return "wait"
    return "comm"


def build_trace(events: List[Event]) -> List[dict]:
    out: List[dict] = []
    ranks = sorted({e.rank for e in events})
    for i, r in enumerate(ranks):
        out.append({"name": "process_name", "ph": "M", "pid": f"rank{r}", "args": {"name": f"rank{r}"}})
        out.append({"name": "process_sort_index", "ph": "M", "pid": f"rank{r}", "args": {"sort_index": i}})
        for lane in LANES:
            out.append({"name": "thread_name", "ph": "M", "pid": f"rank{r}", "tid": lane, "args": {"name": lane}})
            out.append({"name": "thread_sort_index", "ph": "M", "pid": f"rank{r}", "tid": lane,
                        "args": {"sort_index": _SORT[lane]}})
    flows: Dict[str, list] = {}
    for e in events:
        lane = _lane(e)
        short = e.name.split("-")[-1] if e.kind != "scope" else e.name
        rec = {"name": short, "cat": e.kind, "ph": "X", "ts": e.start * 1e3, "dur": max(0.0, (e.end - e.start) * 1e3),
               "pid": f"rank{e.rank}", "tid": lane,
               "args": {"call_stack": e.name, "phase": e.phase, "cost_ms": e.end - e.start}}
        if e.gid is not None:
            rec["args"]["gid"] = e.gid
        if e.post is not None:
            rec["args"]["post_ts"] = e.post * 1e3
            rec["args"]["order"] = e.order
        out.append(rec)
        if e.kind == "p2p" and e.gid:
            flows.setdefault(e.gid, []).append((e, lane))
    fid = 0
    for gid, pair in flows.items():
        if len(pair) != 2:
            continue
        (a, la), (b, lb) = pair
        src, dst = ((a, la), (b, lb)) if "send" in a.name else ((b, lb), (a, la))
        fid += 1
        out.append({"name": "p2p", "cat": "flow", "ph": "s", "id": fid, "ts": src[0].start * 1e3,
                    "pid": f"rank{src[0].rank}", "tid": src[1]})
        out.append({"name": "p2p", "cat": "flow", "ph": "f", "bp": "e", "id": fid, "ts": dst[0].end * 1e3,
                    "pid": f"rank{dst[0].rank}", "tid": dst[1]})
    return out


def write_trace(events: List[Event], path: str, memory_events: List[dict] = None) -> int:
    data = build_trace(events)
    if memory_events:
        data.extend(memory_events)
    with open(path, "w", encoding="utf-8") as fh:
        json.dump(data, fh, indent=1)
    return len(data)


def write_event_log(events: List[Event], path: str) -> int:
    """Plain-text companion (one finished job per line, ms) for grepping / diffing runs."""
    with open(path, "w", encoding="utf-8") as fh:
        for e in events:
            gid = f" gid {e.gid}" if e.gid else ""
            fh.write(f"{e.name}{gid} {e.phase} cost {e.end - e.start:.6f} st {e.start:.6f} ed {e.end:.6f}\n")
    return len(events)

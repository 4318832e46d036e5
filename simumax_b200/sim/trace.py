"""Structured simulator events -> Chrome trace (``tracing_logs.json``), no text-log round trip.

Lane vocabulary matches the reference viewer layout (generate_tracing.py:65-80): per rank (pid ``rankN``) the
threads ``fwd_scope, bwd_scope, fwd_compute, bwd_compute, pp_fwd, pp_bwd, pp_batch_scope, comm, wait``;
times are microseconds; matching async send/recv pairs are linked with flow arrows.
"""
from __future__ import annotations

import json
import re
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
            tail = (f" post {e.post:.6f}" if e.post is not None else "") + (f" order {e.order}" if e.order is not None else "")
            fh.write(f"{e.name}{gid} {e.phase} cost {e.end - e.start:.6f} st {e.start:.6f} ed {e.end:.6f}{tail} kind {e.kind} lane {e.lane}\n")
    return len(events)


# ------------------------------------------------------------------------------------------------
# text log -> trace (the reference's only path: generate_tracing.py:27-54 parse_log_line, :488 process_log_file).
# Here the structured events are the primary path; the log round trip exists so a log produced by either
# implementation can be turned into a trace after the fact.
# ------------------------------------------------------------------------------------------------
_LINE = re.compile(r"rank(?P<rank>\d+)-(?P<stack>[\w-]+)\s+(?:gid\s+(?P<gid>\S+)\s+)?(?P<phase>\w+)\s+cost\s+(?P<cost>[\d.]+)\s+"
                   r"st\s+(?P<st>[\d.]+)\s+ed\s+(?P<ed>[\d.]+)(?:\s+post\s+(?P<post>[\d.]+))?(?:\s+order\s+(?P<order>-?\d+))?"
                   r"(?:\s+kind\s+(?P<kind>\w+)\s+lane\s+(?P<lane>\w+))?")
_P2P = ("send_", "recv_", "async_send", "async_recv", "sync_send", "sync_recv")
_WAIT = ("async_wait", "sync_wait")
_LEAF_COMPUTE = ("AtomModel", "Linear", "GroupedLinear", "AttentionScore", "optimizer_step")     # leaf jobs of sim/jobs.py
_COLL = ("all_reduce", "all_gather", "reduce_scatter", "all2all", "broadcast", "scatter", "gather", "reduce")


def parse_log_line(line: str):
    """One ``log.log`` line -> ``Event`` (None if it is not a job record)."""
    m = _LINE.match(line.strip())
    if not m:
        return None
    stack = m.group("stack")
    leaf = stack.split("-")[-1]
    if m.group("kind"):                      # written by this package: self-describing
        kind, lane = m.group("kind"), m.group("lane")
    elif leaf.startswith(_WAIT):             # a reference-made log: classify by the leaf name
        kind, lane = "wait", "wait"
    elif leaf.startswith(_P2P):
        kind = "p2p"
        lane = "pp_bwd" if "backward" in (m.group("gid") or "") else "pp_fwd"
    elif leaf.startswith(_COLL):
        kind, lane = "comm", "comm"
    elif leaf in _LEAF_COMPUTE:
        kind, lane = "compute", "comp"
    else:
        kind, lane = "scope", "comp"
    return Event(int(m.group("rank")), f"rank{m.group('rank')}-{stack}", kind, m.group("phase"), float(m.group("st")),
                 float(m.group("ed")), lane, gid=m.group("gid"),
                 post=float(m.group("post")) if m.group("post") is not None else None,
                 order=int(m.group("order")) if m.group("order") is not None else None)


def events_from_log(log_path: str) -> List[Event]:
    with open(log_path, "r", encoding="utf-8") as fh:
        return [e for e in (parse_log_line(l) for l in fh if " cost " in l) if e is not None]


def process_log_file(log_path: str, output_json_path: str) -> int:
    """Read a simulator text log and write the Chrome-trace JSON; returns the number of trace records."""
    return write_trace(events_from_log(log_path), output_json_path)

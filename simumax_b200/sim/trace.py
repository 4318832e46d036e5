"""Structured simulator events -> Chrome trace (``tracing_logs.json``), no text-log round trip.

Same records as the reference viewer layout (generate_tracing.py:65-104, :152-480): per rank (pid ``rankN``) the threads
``00_fwd_scope, 01_bwd_scope, 02_fwd_compute, 03_bwd_compute, 04_pp_fwd, 05_pp_bwd, 06_pp_batch_scope, 07_pp_detail_k,
08_comm, 09_wait``; times are microseconds; matching send/recv pairs carry a ``[gN]`` tag and one flow arrow.
"""
from __future__ import annotations

import heapq
import json
import re
from typing import Dict, List

from .engine import Event

# lane -> (ordered thread id, sort index); the numeric prefix keeps the viewer's lane order stable (generate_tracing.py:65-104)
_LANE_ORDER = {"fwd_scope": ("00_fwd_scope", 0), "bwd_scope": ("01_bwd_scope", 1), "fwd_compute": ("02_fwd_compute", 2),
               "bwd_compute": ("03_bwd_compute", 3), "pp_fwd": ("04_pp_fwd", 4), "pp_bwd": ("05_pp_bwd", 5),
               "pp_batch_scope": ("06_pp_batch_scope", 6), "comm": ("08_comm", 80), "wait": ("09_wait", 90)}
LANES = list(_LANE_ORDER)
_SEND = ("send_", "async_send", "sync_send")
_RECV = ("recv_", "async_recv", "sync_recv")
_FWD_DIR = ("async_send_next", "async_recv_prev", "async_wait_recv_prev")
_BWD_DIR = ("async_send_prev", "async_recv_next", "async_wait_recv_next")


def _tid_for(lane: str):
    """(thread id, sort index) of a lane; blocking p2p detail sub-lanes sort after the batch scope."""
    if lane.startswith("pp_detail"):
        suffix = lane[len("pp_detail_"):]
        return f"07_{lane}", 60 + (int(suffix) if suffix.isdigit() else 0)
    return _LANE_ORDER.get(lane, (f"99_{lane}", 100))


def _classify(e: Event, base: str):
    """-> (stream type, lane) of one finished job."""
    if e.kind == "wait" or base.startswith(("async_wait_send", "async_wait_recv")):
        return "wait", "wait"
    if e.kind == "scope" and base.startswith("batch_pp"):
        return "scope", "scope"
    if e.kind in ("p2p", "comm"):
        if base.startswith(("send_", "recv_")):
            return "comm", "pp_detail"                  # blocking p2p: packed into sub-lanes below
        if base.startswith(_FWD_DIR):
            return "comm", "pp_fwd"
        if base.startswith(_BWD_DIR):
            return "comm", "pp_bwd"
        return "comm", "comm"
    return "compute", "compute"


def build_trace(events: List[Event]) -> List[dict]:
    """Chrome-trace records with the reference viewer layout (generate_tracing.py:152-480): ordered lanes per rank, module
    envelopes on the scope lanes and leaf kernels on the compute lanes, blocking p2p packed into non-overlapping
    ``pp_detail_k`` lanes, async p2p drawn from its post time when the lane was idle, ``[gN]`` tags + one flow arrow per
    matched send/recv pair."""
    recs: List[dict] = []
    ends: Dict[str, list] = {}
    for e in events:
        if e.gid:
            base = e.name.split("-")[-1]
            ends.setdefault(e.gid, [False, False])
            ends[e.gid][0] |= base.startswith(_SEND)
            ends[e.gid][1] |= base.startswith(_RECV)
    corr: Dict[str, int] = {}
    for i, e in enumerate(events):
        stack = e.name.split("-")[1:] or [e.name]
        base = stack[-1]
        stream, lane = _classify(e, base)
        if stream == "compute":
            tid = "bwd_compute" if e.phase == "recompute_fwd" else f"{e.phase}_compute"
        else:
            tid = "pp_batch_scope" if stream == "scope" else lane
        cid = None
        if stream == "comm" and e.gid and all(ends.get(e.gid, (False, False))):
            cid = corr.setdefault(e.gid, len(corr))
        recs.append({"name": base if cid is None else f"{base}[g{cid}]", "cat": stream, "ph": "X", "ts": e.start * 1e3,
                     "dur": max(0.0, (e.end - e.start) * 1e3), "pid": f"rank{e.rank}", "tid": tid, "id": i,
                     "args": {"call_stack": stack, "stream_type": stream, "lane": lane, "lane_base": lane, "direction": e.phase,
                              "gid": e.gid, "correlation_id": cid, "base_name": base,
                              "post_ts": None if e.post is None else e.post * 1e3, "post_order": e.order}})

    # blocking send/recv of one rank may overlap in time: greedy interval colouring into pp_detail_0, _1, ...
    detail: Dict[str, list] = {}
    for r in recs:
        if r["cat"] == "comm" and r["args"]["lane"] == "pp_detail":
            detail.setdefault(r["pid"], []).append(r)
    for rs in detail.values():
        rs.sort(key=lambda r: (r["ts"], r["dur"], r["id"]))
        busy, free, used = [], [], 0
        for r in rs:
            while busy and busy[0][0] <= r["ts"] + 1e-9:
                heapq.heappush(free, heapq.heappop(busy)[1])
            if free:
                k = heapq.heappop(free)
            else:
                k, used = used, used + 1
            r["tid"] = r["args"]["lane"] = f"pp_detail_{k}"
            r["args"]["detail_lane_idx"] = k
            heapq.heappush(busy, (r["ts"] + r["dur"], k))

    # async p2p: show the slice from its local post time if the lane had nothing else to do (no false bubble)
    per_lane: Dict[tuple, list] = {}
    for r in recs:
        if r["cat"] == "comm":
            per_lane.setdefault((r["pid"], r["tid"]), []).append(r)
    for rs in per_lane.values():
        rs.sort(key=lambda r: (r["ts"], r["id"]))
        tail = None
        for r in rs:
            post = r["args"]["post_ts"]
            if post is not None:
                begin = post if tail is None else max(post, tail)
                if begin < r["ts"]:
                    r["dur"] = max(0.0, r["ts"] + r["dur"] - begin)
                    r["ts"] = begin
            tail = r["ts"] + r["dur"]
    # back-to-back pipeline transfers would be swallowed by the viewer: trim 10 % (<= 0.5 ns) off both ends
    for r in recs:
        if r["cat"] == "comm" and r["args"]["gid"] and "send_recv-" in r["args"]["gid"] and r["dur"] > 0.0:
            pad = min(5e-4, r["dur"] * 0.1)
            if r["dur"] > 2.0 * pad:
                r["ts"] += pad
                r["dur"] -= 2.0 * pad

    # a compute-stream job that encloses another one of its own subtree is a module envelope -> scope lane
    inside: Dict[tuple, list] = {}
    for r in recs:
        if r["cat"] == "compute":
            st = tuple(r["args"]["call_stack"])
            for k in range(1, len(st)):
                inside.setdefault((r["pid"], st[:k]), []).append((r["ts"], r["ts"] + r["dur"]))
    for r in recs:
        if r["cat"] != "compute":
            continue
        base, lo, hi = r["args"]["base_name"], r["ts"], r["ts"] + r["dur"] + 1e-9
        if base in ("recompute_block", "checkpoint_bwd") or any(
                a >= lo and b <= hi for a, b in inside.get((r["pid"], tuple(r["args"]["call_stack"])), ())):
            r["cat"] = "scope"
            if r["args"]["direction"] == "recompute_fwd":
                r["tid"] = "bwd_scope"
                if base == "recompute_block":
                    r["name"] = "recompute_fwd"
            else:
                r["tid"] = f"{r['args']['direction']}_scope"

    # one arrow per matched pipeline transfer, anchored just inside the two slices
    pairs: Dict[str, list] = {}
    for r in recs:
        gid = r["args"]["gid"]
        if r["cat"] == "comm" and gid and "send_recv-" in gid:
            pairs.setdefault(gid, []).append(r)
    flows = []
    for gid, rs in pairs.items():
        snd = [r for r in rs if r["args"]["base_name"].startswith(_SEND)]
        rcv = [r for r in rs if r["args"]["base_name"].startswith(_RECV)]
        if len(snd) != 1 or len(rcv) != 1:
            continue
        for ph, r, at_end in (("s", snd[0], False), ("f", rcv[0], True)):
            pad = min(1e-3, r["dur"] * 0.25) if r["dur"] > 0.0 else 0.0
            rec = {"name": f"pair:{gid}", "cat": "comm_pair", "ph": ph, "pid": r["pid"], "tid": r["tid"], "id": len(flows) // 2,
                   "ts": r["ts"] + (max(0.0, r["dur"] - pad) if at_end else pad), "args": {"gid": gid}}
            if ph == "f":
                rec["bp"] = "e"
            flows.append(rec)
    recs.extend(flows)

    order: Dict[str, tuple] = {}
    for r in recs:
        ordered = _tid_for(r["tid"])
        order[ordered[0]] = ordered
        r["tid"] = ordered[0]
    meta: List[dict] = []
    pids = sorted({r["pid"] for r in recs}, key=lambda p: int(re.sub(r"\D", "", p) or 0))
    for i, pid in enumerate(pids):
        meta.append({"name": "process_name", "ph": "M", "pid": pid, "args": {"name": pid}})
        meta.append({"name": "process_sort_index", "ph": "M", "pid": pid, "args": {"sort_index": i}})
        meta.append({"name": "sort_index", "ph": "M", "pid": pid, "args": {"sort_index": i}})
        for tid in sorted({r["tid"] for r in recs if r["pid"] == pid}):
            idx = order[tid][1]
            meta.append({"name": "thread_name", "ph": "M", "pid": pid, "tid": tid, "args": {"name": tid.split("_", 1)[1]}})
            meta.append({"name": "thread_sort_index", "ph": "M", "pid": pid, "tid": tid, "args": {"sort_index": idx}})
            meta.append({"name": "sort_index", "ph": "M", "pid": pid, "tid": tid, "args": {"sort_index": idx}})
    return meta + recs


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

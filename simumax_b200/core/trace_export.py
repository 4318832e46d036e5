"""Chrome-trace export of *analytic* pipeline schedules (1F1B / sync-VPP evaluator records).

Output: ``{"traceEvents": [...], "displayTimeUnit": "ms"}``; one process per pipeline rank with the lanes
``fwd_compute``, ``bwd_compute`` and ``comm`` (p2p); overlapping p2p slices of one rank are serialised on the
comm lane for readability (their true interval is kept in ``args``).  (reference core/trace_export.py)
"""
from __future__ import annotations

import json
import os
from typing import List


def normalize_schedule_records(schedules) -> List[List[dict]]:
    out = []
    for rank, tasks in enumerate(schedules):
        rows = []
        for task in tasks:
            if isinstance(task, dict):
                row = dict(task)
            else:
                kind, mb, start, duration, end = task
                row = {"kind": kind, "mb": mb, "start": start, "duration": duration, "end": end, "label": kind}
            row.setdefault("rank", rank)
            rows.append(row)
        out.append(rows)
    return out


def _is_compute(task) -> bool:
    kind = task.get("kind", task.get("label", ""))
    return kind.startswith("F") or kind.startswith("B")


def serialize_comm_lane_for_trace(schedules):
    out = []
    for tasks in schedules:
        rows = [dict(t) for t in tasks]
        idxs = [i for i, t in enumerate(rows) if not _is_compute(t)]
        idxs.sort(key=lambda i: (float(rows[i]["start"]), float(rows[i]["end"]), rows[i].get("label", "")))
        cursor = 0.0
        for i in idxs:
            t = rows[i]
            s0, e0 = float(t["start"]), float(t["end"])
            s = max(s0, cursor)
            e = max(s, e0)
            t["true_start"], t["true_end"] = s0, e0
            t["start"], t["end"], t["duration"] = s, e, e - s
            if t["duration"] <= 0.0 and e0 - s0 > 0.0:
                t["collapsed_for_trace"] = True
            cursor = e
        out.append(rows)
    return out


def export_pipeline_schedule_trace(schedules, output_json_path, title="perf_pp_schedule", *, serialize_comm_lanes=True):
    output_json_path = os.path.abspath(output_json_path)
    os.makedirs(os.path.dirname(output_json_path), exist_ok=True)
    schedules = normalize_schedule_records(schedules)
    if serialize_comm_lanes:
        schedules = serialize_comm_lane_for_trace(schedules)
    events = []
    for rank, tasks in enumerate(schedules):
        pid = f"rank{rank}"
        events.append({"name": "process_name", "ph": "M", "pid": pid, "args": {"name": pid}})
        events.append({"name": "process_sort_index", "ph": "M", "pid": pid, "args": {"sort_index": rank}})
        for i, lane in enumerate(("fwd_compute", "bwd_compute", "comm")):
            events.append({"name": "thread_name", "ph": "M", "pid": pid, "tid": lane, "args": {"name": lane}})
            events.append({"name": "thread_sort_index", "ph": "M", "pid": pid, "tid": lane, "args": {"sort_index": i}})
        for t in tasks:
            if t.get("collapsed_for_trace"):
                continue
            kind = t.get("kind", "")
            lane = "fwd_compute" if kind.startswith("F") else ("bwd_compute" if kind.startswith("B") else "comm")
            name = t.get("label", kind)
            if t.get("mb") is not None:
                name = f"{name}[mb{t['mb']}" + (f",chunk{t['chunk_idx']}]" if t.get("chunk_idx") is not None else "]")
            args = {"scheduler": title, "kind": kind, "duration_ms": float(t.get("duration") or 0.0)}
            for k in ("gid", "chunk_idx", "true_start", "true_end"):
                if t.get(k) is not None:
                    args[k] = t[k]
            events.append({"name": name, "cat": "compute" if lane != "comm" else "comm", "ph": "X", "pid": pid,
                           "tid": lane, "ts": float(t["start"]) * 1e3,
                           "dur": max(0.0, float(t["end"]) - float(t["start"])) * 1e3, "args": args})
    with open(output_json_path, "w", encoding="utf-8") as fh:
        json.dump({"traceEvents": events, "displayTimeUnit": "ms"}, fh, ensure_ascii=False, indent=1)
    return output_json_path

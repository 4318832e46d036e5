"""Per-rank allocated-memory timeline of a simulated iteration.

allocated = static (weights + grads + optimizer state) + cached activations (live "cache tokens") + the
op-local temporary peak.  Tokens are allocated at the end of the phase that materialises them (``fwd`` or
``recompute_fwd``) and released at the end of the op's ``bwd``.

Exports: ``summary()`` (static / peak bytes per rank), ``snapshot()`` with schema ``simumax_memory_snapshot_v1``
and a ``torch.cuda.memory._snapshot()``-compatible payload for PyTorch's memory_viz.
(Behavioural spec: reference core/simu_memory.py.)
"""
from __future__ import annotations

import re
from collections import defaultdict
from dataclasses import dataclass
from typing import Optional


@dataclass
class OpMemoryProfile:
    op_name: str = ""
    fwd_peak_mem_no_cache: int = 0
    bwd_peak_mem_no_cache: int = 0
    recompute_peak_mem_no_cache: int = 0
    cache_size_bytes: int = 0
    cache_alloc_phase: Optional[str] = None
    cache_release_phase: Optional[str] = "bwd"
    cache_token_scope: str = ""

    def phase_peak_no_cache(self, phase: str) -> int:
        return int({"fwd": self.fwd_peak_mem_no_cache, "recompute_fwd": self.recompute_peak_mem_no_cache,
                    "bwd": self.bwd_peak_mem_no_cache}[phase])

    def phase_allocates_cache(self, phase: str) -> bool:
        return bool(self.cache_size_bytes) and phase == self.cache_alloc_phase

    def phase_releases_cache(self, phase: str) -> bool:
        return bool(self.cache_size_bytes) and phase == self.cache_release_phase


_MB = re.compile(r"microbatch(\d+)")
_CH = re.compile(r"chunk(\d+)")


def _tags(scope: str):
    scope = scope or ""
    mb, ch = _MB.search(scope), _CH.search(scope)
    return {"scope": scope, "microbatch": int(mb.group(1)) if mb else None, "chunk": int(ch.group(1)) if ch else None}


class SimuMemoryTracker:
    def __init__(self):
        self.static_bytes = defaultdict(int)
        self.cached_bytes = defaultdict(int)
        self.peak_bytes = defaultdict(int)
        self.peak_op = {}
        self.events = []            # Chrome-trace counter events
        self.snapshots = []
        self.cache_token_events = []
        self._next_token = 0
        self._live = defaultdict(dict)
        self._by_key = defaultdict(lambda: defaultdict(list))

    # ---- tokens
    @staticmethod
    def _key(profile: OpMemoryProfile) -> str:
        return f"{profile.cache_token_scope or profile.op_name}|{profile.op_name}"

    def _alloc(self, rank, ts, profile, phase, size):
        size = int(size)
        if size <= 0:
            return
        tid = self._next_token
        self._next_token += 1
        tg = _tags(profile.cache_token_scope or profile.op_name)
        token = {"token_id": tid, "rank": f"rank{rank}", "token_key": self._key(profile),
                 "token_scope": profile.cache_token_scope or profile.op_name, "op_name": profile.op_name,
                 "microbatch": tg["microbatch"], "chunk": tg["chunk"], "alloc_phase": phase,
                 "alloc_ts_us": float(ts) * 1e3, "free_phase": None, "free_ts_us": None, "size_bytes": size}
        self._live[rank][tid] = token
        self._by_key[rank][token["token_key"]].append(tid)
        self.cache_token_events.append({"action": "alloc", **token})
        self.cached_bytes[rank] += size

    def _free(self, rank, ts, profile, phase):
        if int(profile.cache_size_bytes) <= 0:
            return
        key = self._key(profile)
        queue = self._by_key[rank].get(key, [])
        if not queue:
            raise RuntimeError(f"missing cached token for rank{rank} key={key} release={profile.cache_size_bytes}")
        tid = queue.pop(0)
        token = self._live[rank].pop(tid)
        if not queue:
            self._by_key[rank].pop(key, None)
        if token["size_bytes"] != int(profile.cache_size_bytes):
            raise RuntimeError(f"cached token size mismatch for rank{rank} key={key}: live={token['size_bytes']} "
                               f"release={int(profile.cache_size_bytes)}")
        token["free_phase"], token["free_ts_us"] = phase, float(ts) * 1e3
        self.cache_token_events.append({"action": "free", **token})
        self.cached_bytes[rank] -= token["size_bytes"]
        if self.cached_bytes[rank] < 0:
            raise RuntimeError(f"cached_bytes underflow for rank{rank}")

    def _counter(self, rank, ts, allocated, phase, op_name, kind, scope=""):
        allocated = int(allocated)
        if allocated > self.peak_bytes[rank]:
            self.peak_bytes[rank] = allocated
            self.peak_op[rank] = (op_name, phase)
        static, cached = int(self.static_bytes[rank]), int(self.cached_bytes[rank])
        temp = max(0, allocated - static - cached)
        ntok = len(self._live[rank])
        self.events.append({"name": "mem", "cat": "memory", "ph": "C", "ts": float(ts) * 1e3, "pid": f"rank{rank}",
                            "args": {"allocated_bytes": allocated, "static_bytes": static, "cached_bytes": cached,
                                     "temp_bytes": temp, "cached_token_count": ntok, "phase": phase,
                                     "op_name": op_name, "kind": kind}})
        tg = _tags(scope)
        self.snapshots.append({"rank": f"rank{rank}", "ts_us": float(ts) * 1e3, "allocated_bytes": allocated,
                               "static_bytes": static, "cached_bytes": cached, "temp_bytes": temp,
                               "cached_token_count": ntok, "phase": phase, "op_name": op_name, "kind": kind,
                               "scope": tg["scope"], "microbatch": tg["microbatch"], "chunk": tg["chunk"]})

    # ---- engine callbacks
    def init_rank(self, rank, static_bytes):
        self.static_bytes[rank] = int(static_bytes)
        self.cached_bytes[rank] = 0
        self._counter(rank, 0.0, self.static_bytes[rank], "init", "static", "init")

    def phase_start(self, rank, ts, profile, phase):
        base = self.static_bytes[rank] + self.cached_bytes[rank]
        self._counter(rank, ts, base, phase, profile.op_name, "start", profile.cache_token_scope)
        self._counter(rank, ts + 1e-9, base + profile.phase_peak_no_cache(phase), phase, profile.op_name, "peak",
                      profile.cache_token_scope)

    def phase_end(self, rank, ts, profile, phase):
        if profile.phase_allocates_cache(phase):
            self._alloc(rank, ts, profile, phase, int(profile.cache_size_bytes))
        elif profile.phase_releases_cache(phase):
            self._free(rank, ts, profile, phase)
        self._counter(rank, ts, self.static_bytes[rank] + self.cached_bytes[rank], phase, profile.op_name, "end",
                      profile.cache_token_scope)

    # ---- exports
    def summary(self):
        return {"static_allocated_bytes_by_rank": {f"rank{r}": int(v) for r, v in sorted(self.static_bytes.items())},
                "peak_allocated_bytes_by_rank": {f"rank{r}": int(v) for r, v in sorted(self.peak_bytes.items())},
                "peak_op_by_rank": {f"rank{r}": {"op_name": v[0], "phase": v[1]} for r, v in sorted(self.peak_op.items())}}

    def snapshot(self):
        return {"schema": "simumax_memory_snapshot_v1",
                "notes": ["allocated_bytes includes static + cached + temporary op-local peak bytes",
                          "temp_bytes is derived as allocated_bytes - static_bytes - cached_bytes",
                          "cached_bytes represents currently live activation cache retained for backward",
                          "cache_tokens records first-hand cached activation lifetime tracked by the simulator"],
                "events": self.snapshots, "cache_tokens": self.cache_token_events}

    def memory_viz_snapshot(self):
        """Minimal ``torch.cuda.memory._snapshot()``-shaped dict: one segment per rank, alloc/free trace entries for
        the static block, every cache token and every op-local temporary (loadable by pytorch.org/memory_viz)."""
        def dev(rank_name):
            return int(rank_name[4:]) if isinstance(rank_name, str) else int(rank_name)

        ranks = sorted({s["rank"] for s in self.snapshots}, key=dev)
        ndev = (max(dev(r) for r in ranks) + 1) if ranks else 0
        out = {"device_traces": [[] for _ in range(ndev)], "segments": []}
        for rname in ranks:
            d = dev(rname)
            base = (d + 1) << 40
            static = int(self.static_bytes[d])
            peak = int(self.peak_bytes[d])
            trace = out["device_traces"][d]
            trace.append({"action": "segment_alloc", "addr": base, "size": max(peak, static), "stream": 0, "frames": []})
            if static:
                trace.append({"action": "alloc", "addr": base, "size": static, "stream": 0,
                              "frames": [{"filename": "simumax", "line": 0, "name": "static:init:init:static"}]})
            cursor = base + static
            addr_of = {}
            # cache tokens get bump-allocated addresses; temporaries live above the currently cached bytes
            timeline = []
            for ev in self.cache_token_events:
                if ev["rank"] != rname:
                    continue
                ts = ev["alloc_ts_us"] if ev["action"] == "alloc" else ev["free_ts_us"]
                timeline.append((ts, 0 if ev["action"] == "alloc" else 1, ev))
            temps = [s for s in self.snapshots if s["rank"] == rname and s["kind"] == "peak" and s["temp_bytes"] > 0]
            for s in temps:
                timeline.append((s["ts_us"], 2, s))
            timeline.sort(key=lambda x: (x[0], x[1]))
            live_blocks = []
            for ts, kind, ev in timeline:
                if kind == 0:
                    addr_of[ev["token_id"]] = cursor
                    tag = f"[mb{ev['microbatch']}]" if ev["microbatch"] is not None else ""
                    frames = [{"filename": "simumax", "line": 0, "name": f"cached:{ev['alloc_phase']}:alloc:{ev['op_name']}{tag}"}]
                    trace.append({"action": "alloc", "addr": cursor, "size": ev["size_bytes"], "stream": 0, "frames": frames})
                    live_blocks.append((cursor, ev["size_bytes"], frames, ev["token_id"]))
                    cursor += ev["size_bytes"]
                elif kind == 1:
                    a = addr_of.pop(ev["token_id"], None)
                    if a is not None:
                        trace.append({"action": "free_requested", "addr": a, "size": ev["size_bytes"], "stream": 0, "frames": []})
                        trace.append({"action": "free_completed", "addr": a, "size": ev["size_bytes"], "stream": 0, "frames": []})
                        live_blocks = [b for b in live_blocks if b[3] != ev["token_id"]]
                else:
                    frames = [{"filename": "simumax", "line": 0, "name": f"temp:{ev['phase']}:peak:{ev['op_name']}"}]
                    trace.append({"action": "alloc", "addr": cursor, "size": ev["temp_bytes"], "stream": 0, "frames": frames})
                    trace.append({"action": "free_requested", "addr": cursor, "size": ev["temp_bytes"], "stream": 0, "frames": []})
                    trace.append({"action": "free_completed", "addr": cursor, "size": ev["temp_bytes"], "stream": 0, "frames": []})
            blocks = []
            if static:
                blocks.append({"address": base, "size": static, "requested_size": static, "state": "active_allocated",
                               "frames": [{"filename": "simumax", "line": 0, "name": "static:init:init:static"}]})
            for a, sz, frames, _ in live_blocks:
                blocks.append({"address": a, "size": sz, "requested_size": sz, "state": "active_allocated", "frames": frames})
            out["segments"].append({"device": d, "address": base, "total_size": max(peak, static), "allocated_size": static,
                                    "active_size": static, "requested_size": static, "stream": 0,
                                    "segment_type": "large", "segment_pool_id": (0, 0), "is_expandable": False,
                                    "frames": [], "blocks": blocks})
        return out

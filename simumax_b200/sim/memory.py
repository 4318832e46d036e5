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
        """``torch.cuda.memory._snapshot()``-shaped dict (loadable by pytorch.org/memory_viz), same layout as the reference's
        exporter (simu_memory.py:212-555): one segment per rank sized for a bump allocator (static block first, then every
        cache token and op-local temporary at ever-increasing addresses), an alloc / free_requested / free_completed trace
        whose frames read ``component:phase:kind:op[mbN,chunkK]``, a ``snapshot`` marker after the static block, and the final
        block list (inactive + still-active blocks, gaps filled)."""
        def dev(rank_name):
            return int(rank_name[4:]) if isinstance(rank_name, str) else int(rank_name)

        def frame(op, phase, kind, component, ev):
            tags = [f"{k}{ev[f]}" for k, f in (("mb", "microbatch"), ("chunk", "chunk")) if ev.get(f) is not None]
            return [{"filename": "simumax", "line": 0, "name": f"{component}:{phase}:{kind}:{op}" + (f"[{','.join(tags)}]" if tags else "")}]

        kind_rank = {"init": 0, "start": 1, "peak": 2, "end": 3}
        ranks = sorted({s["rank"] for s in self.snapshots}, key=dev)
        out = {"device_traces": [[] for _ in range((max(map(dev, ranks)) + 1) if ranks else 0)], "segments": []}
        for rname in ranks:
            d = dev(rname)
            trace = out["device_traces"][d]
            events = sorted((e for e in self.snapshots if e["rank"] == rname),
                            key=lambda e: (not (e["phase"] == "init" and e["kind"] == "init"), e["ts_us"], kind_rank.get(e["kind"], 9)))
            tokens = sorted((e for e in self.cache_token_events if e["rank"] == rname),
                            key=lambda e: (e["alloc_ts_us"] if e["action"] == "alloc" else e["free_ts_us"], e["action"] != "alloc",
                                           e["token_id"]))
            # a bump allocator never reuses an address: the segment must hold everything that is ever allocated
            grown, seen = 0, {"cached_bytes": 0, "temp_bytes": 0}
            for e in events:
                for comp in seen:
                    grown += max(0, int(e[comp]) - seen[comp])
                    seen[comp] = int(e[comp])
            grown = max(grown, sum(int(t["size_bytes"]) for t in tokens if t["action"] == "alloc"))
            static = int(self.static_bytes[d])
            base = (d + 1) << 40
            seg_size = int(max(self.peak_bytes[d], grown + static))
            if seg_size > 0:
                trace.append({"action": "segment_alloc", "addr": base, "size": seg_size, "stream": 0, "frames": []})
            bump = base + static
            live = {"static_bytes": [], "temp_bytes": []}           # stacks of {"addr", "size", "frames"}
            live_tokens, retired = {}, []
            level = {"static_bytes": 0, "temp_bytes": 0}
            marker_done = False
            ti = 0

            def log(action, blk):
                trace.append({"action": action, "addr": blk["addr"], "size": blk["size"], "stream": 0, "frames": blk["frames"]})

            def retire(blk):
                retired.append({"address": blk["addr"], "size": blk["size"], "requested_size": blk["size"], "state": "inactive",
                                "frames": blk["frames"]})

            def drain(ts, inclusive):
                nonlocal ti, bump
                while ti < len(tokens):
                    t = tokens[ti]
                    tts = t["alloc_ts_us"] if t["action"] == "alloc" else t["free_ts_us"]
                    if tts > ts or (tts == ts and not inclusive):
                        return
                    if t["action"] == "alloc":
                        blk = {"addr": bump, "size": int(t["size_bytes"]),
                               "frames": frame(t["op_name"], t["alloc_phase"], "alloc", "cached", t)}
                        bump += blk["size"]
                        live_tokens[t["token_id"]] = blk
                        log("alloc", blk)
                    else:
                        blk = live_tokens.pop(t["token_id"])
                        log("free_requested", blk)
                        log("free_completed", blk)
                        retire(blk)
                    ti += 1

            for e in events:
                drain(e["ts_us"], False)
                if not marker_done and not (e["phase"] == "init" and e["kind"] == "init"):
                    trace.append({"action": "snapshot", "stream": 0, "frames": frame("static", "init", "snapshot", "static", {})})
                    marker_done = True
                for comp in ("static_bytes", "temp_bytes"):          # cached bytes are the tokens above
                    delta = int(e[comp]) - level[comp]
                    name = comp.split("_")[0]
                    if delta > 0:
                        if comp == "static_bytes":
                            addr = base + level[comp]
                        else:
                            addr, bump = bump, bump + delta
                        blk = {"addr": addr, "size": delta, "frames": frame(e["op_name"], e["phase"], e["kind"], name, e)}
                        live[comp].append(dict(blk))
                        log("alloc", blk)
                    elif delta < 0:
                        need = -delta
                        while need > 0:
                            if not live[comp]:
                                raise RuntimeError(f"free underflow in memory_viz export for {name}: {need}")
                            top = live[comp][-1]
                            if top["size"] <= need:
                                freed, rest = live[comp].pop(), None
                            else:                                   # partial free: the tail of the block stays allocated
                                freed = {"addr": top["addr"], "size": need, "frames": top["frames"]}
                                rest = {"addr": top["addr"] + need, "size": top["size"] - need, "frames": top["frames"]}
                                live[comp][-1] = rest
                            need -= freed["size"]
                            if comp != "static_bytes":
                                log("free_requested", freed)
                                log("free_completed", freed)
                                if rest is not None:
                                    log("alloc", rest)
                            retire(freed)
                    level[comp] = int(e[comp])
                drain(e["ts_us"], True)
            if not marker_done:
                trace.append({"action": "snapshot", "stream": 0, "frames": frame("static", "init", "snapshot", "static", {})})

            active = [{"address": b["addr"], "size": b["size"], "requested_size": b["size"], "state": "active_allocated",
                       "frames": b["frames"]}
                      for b in live["static_bytes"] + list(live_tokens.values()) + live["temp_bytes"]]
            blocks, cursor = [], base
            for b in sorted(retired + active, key=lambda b: b["address"]):
                if cursor < b["address"]:
                    blocks.append({"size": b["address"] - cursor, "state": "inactive"})
                blocks.append({"size": b["size"], "requested_size": b["requested_size"], "frames": b["frames"], "state": b["state"]})
                cursor = b["address"] + b["size"]
            if cursor < base + seg_size:
                blocks.append({"size": base + seg_size - cursor, "state": "inactive"})
            in_use = int(self.static_bytes[d] + self.cached_bytes[d])
            out["segments"].append({"device": d, "address": base, "total_size": seg_size, "stream": 0, "segment_type": "large",
                                    "allocated_size": in_use, "active_size": in_use,
                                    "blocks": blocks or [{"size": max(seg_size, 1), "state": "inactive"}]})
        return out

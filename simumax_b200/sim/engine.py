"""Discrete-event replay engine.

Every simulated rank is a Python *coroutine* that walks its job tree and ``yield``s whenever it has to wait
for another rank (collective rendezvous, blocking p2p, async-p2p completion).  A tiny scheduler resumes
runnable ranks until all are done; if nobody can make progress the pending rendezvous table is dumped and a
``RuntimeError("deadlock")`` is raised.  Because all timing is dataflow-determined (per-lane FIFO order +
``max`` of ready times), the result does not depend on the order in which ranks are resumed.

Timing semantics (kept identical to the reference simulator, base_struct.py:35-229,1225-2733):

* lanes per rank: ``comp``, ``comm`` and ``pp_fwd`` / ``pp_bwd`` for asynchronous pipeline p2p;
* a compute atom advances ``comp``; every finished scope (module forward / backward) adds 2 us of tracing gap;
* a synchronous collective is issued at ``comp`` on the ``comm`` lane FIFO; ``local`` kind (one representative
  rank per pipeline stage) ends at ``max(issue, lane_tail) + cost``; ``barrier`` kind waits for all ``expected``
  participants and ends at ``max(ready) + cost``; the caller blocks and then ``comp = comm = end``;
* blocking p2p is a two-party barrier; async p2p posts FIFO entries on the pp lanes, each pair completes at
  ``max(ready_i + cost_i)`` and ``wait`` only advances ``comp`` to that time.

Unlike the reference there is no text log + regex round trip: the engine emits structured ``Event`` records
that the trace / memory exporters consume directly.
"""
from __future__ import annotations

import math

from collections import defaultdict, deque
from dataclasses import dataclass, field
from typing import Dict, Generator, List, Optional, Tuple

SCOPE_GAP_MS = 2e-3


@dataclass
class Event:
    rank: int
    name: str            # hierarchical call-stack label
    kind: str            # scope | compute | comm | p2p | wait
    phase: str           # fwd | bwd | recompute_fwd
    start: float         # ms
    end: float           # ms
    lane: str = "comp"
    gid: Optional[str] = None
    post: Optional[float] = None
    order: Optional[int] = None


# ------------------------------------------------------------------------------------------------
# job tree
# ------------------------------------------------------------------------------------------------
@dataclass
class Atom:
    name: str
    fwd_cost: float = 0.0
    bwd_cost: float = 0.0
    recompute_cost: Optional[float] = None


@dataclass
class Coll:
    """Synchronous collective.  ``fwd_cost`` / ``bwd_cost`` == 0 (or a group of one) means "not issued"."""
    name: str
    gid: str
    group_size: int
    fwd_cost: float = 0.0
    bwd_cost: float = 0.0
    world: bool = False          # end-of-step world all-reduce: rendezvous across all simulated ranks


@dataclass
class BlockingP2P:
    name: str
    gid: str
    cost: float


@dataclass
class AsyncPost:
    name: str
    gid: str
    cost: float
    stream: str
    role: str                    # send | recv


@dataclass
class AsyncWait:
    name: str
    gid: str
    cost: float
    stream: str


@dataclass
class Scope:
    name: str
    children: list = field(default_factory=list)
    mem_profile: object = None
    batch_blocking: bool = False


@dataclass
class RecomputeBlock:
    """Checkpointed segment: replay the forward of ``fwd_children`` then run ``bwd_children`` backward."""
    name: str
    fwd_children: list = field(default_factory=list)
    bwd_children: list = field(default_factory=list)


class RankState:
    def __init__(self, rank: int):
        self.rank = rank
        self.t = defaultdict(float, {"comp": 0.0, "comm": 0.0})
        self.comm_order = 0


class Deadlock(RuntimeError):
    pass


class Engine:
    def __init__(self, merge_lanes: bool = True, memory_tracker=None, jitter: float = 0.0, slow_ranks: Optional[Dict[int, float]] = None,
                 seed: int = 0):
        """``jitter``: relative standard deviation of a log-normal factor applied to every compute atom (run-to-run noise);
        ``slow_ranks``: rank -> constant slow-down factor of its compute (a throttling / faulty device).  Both default to a
        deterministic replay.  (Fault injection the reference only hints at: the commented-out jitter hook in AtomModel,
        base_struct.py:2082-2083.)"""
        self.merge_lanes = merge_lanes
        self.memory_tracker = memory_tracker
        self.jitter = float(jitter)
        self.slow_ranks = dict(slow_ranks or {})
        self._rng = __import__("random").Random(seed)
        self.events: List[Event] = []
        self.ranks: Dict[int, RankState] = {}
        self.programs: Dict[int, Generator] = {}
        self._lane_tail: Dict[Tuple[int, str], float] = defaultdict(float)
        self._lane_queue: Dict[Tuple[int, str], deque] = defaultdict(deque)
        self._barrier: Dict[tuple, dict] = {}
        self._barrier_done: Dict[tuple, float] = {}
        self._async: Dict[tuple, dict] = {}
        self._issue_seq = 0
        self._completed_coll: Dict[int, set] = defaultdict(set)

    # ------------------------------------------------------------------ public
    def add_rank(self, rank: int, jobs: list) -> RankState:
        st = RankState(rank)
        self.ranks[rank] = st
        self.programs[rank] = self._run_rank(st, jobs)
        return st

    def run(self) -> float:
        alive = list(self.programs.keys())
        blocked: Dict[int, tuple] = {}
        while alive:
            progressed = False
            still = []
            for r in alive:
                # run this rank until it blocks on something another rank must provide
                while True:
                    try:
                        key = next(self.programs[r])
                    except StopIteration:
                        progressed = True
                        blocked.pop(r, None)
                        break
                    if key is None or key[0] in ("yield", "progress"):
                        progressed = True
                        continue
                    blocked[r] = key
                    still.append(r)
                    break
            # a rank that blocked may have been released by a later rank in the same sweep; one extra sweep
            # without any progress proves a deadlock
            if not progressed and still:
                self._raise_deadlock(blocked)
            alive = still
        end = 0.0
        for st in self.ranks.values():
            end = max(end, max(st.t.values()))
        return end

    def _raise_deadlock(self, blocked):
        lines = ["DEADLOCK: no rank can make progress", f"blocked: {dict(list(blocked.items())[:16])}"]
        for gid, s in list(self._barrier.items())[:16]:
            lines.append(f"  barrier {gid}: arrived {s['arrived']}/{s['expected']} waiters {s['waiters'][:8]}")
        for (rank, stream), q in list(self._lane_queue.items())[:16]:
            if q:
                lines.append(f"  lane rank{rank}/{stream}: head {q[0]['gid']} status {q[0]['status']}")
        raise Deadlock("\n".join(lines))

    # ------------------------------------------------------------------ rank program
    def _run_rank(self, st: RankState, jobs: list):
        for direction, job in jobs:
            yield from self._exec(st, job, direction, "fwd" if direction == "fwd" else "bwd")

    def _emit(self, st, name, kind, phase, start, end, lane="comp", **kw):
        self.events.append(Event(st.rank, name, kind, phase, start, end, lane, **kw))

    def _exec(self, st: RankState, job, direction: str, phase: str):
        """direction: 'fwd' (children in order, forward costs) or 'bwd' (children reversed, backward costs)."""
        if isinstance(job, Scope):
            yield from self._exec_scope(st, job, direction, phase)
        elif isinstance(job, RecomputeBlock):
            if job.fwd_children:
                yield from self._exec_scope(st, Scope(f"{job.name}-recompute_block", job.fwd_children), "fwd",
                                            "recompute_fwd")
            yield from self._exec_scope(st, Scope(f"{job.name}-checkpoint_bwd", job.bwd_children), "bwd", "bwd")
        elif isinstance(job, Atom):
            if phase == "recompute_fwd":
                cost = job.fwd_cost if job.recompute_cost is None else job.recompute_cost
            else:
                cost = job.fwd_cost if direction == "fwd" else job.bwd_cost
            if self.jitter > 0.0 and cost > 0.0:
                cost *= math.exp(self._rng.gauss(0.0, self.jitter) - 0.5 * self.jitter * self.jitter)      # mean-one log-normal
            if self.slow_ranks:
                cost *= self.slow_ranks.get(st.rank, 1.0)
            t0 = st.t["comp"]
            st.t["comp"] = t0 + cost
            if st.t["comp"] != t0:
                self._emit(st, job.name, "compute", phase, t0, st.t["comp"])
        elif isinstance(job, Coll):
            if phase != "recompute_fwd":      # a replayed forward re-uses the already completed collective
                yield from self._exec_coll(st, job, direction)
        elif isinstance(job, BlockingP2P):
            yield from self._exec_blocking(st, job, None)
        elif isinstance(job, AsyncPost):
            self._post_async(st, job)
            yield ("yield", job.gid)
        elif isinstance(job, AsyncWait):
            yield from self._exec_async_wait(st, job)
        else:
            raise TypeError(f"unknown job {job!r}")

    def _exec_scope(self, st, scope: Scope, direction: str, phase: str):
        t0 = st.t["comp"]
        prof = scope.mem_profile
        mt = self.memory_tracker
        if prof is not None and mt is not None:
            mt.phase_start(rank=st.rank, ts=t0, profile=prof, phase=phase)
        if scope.batch_blocking:
            submit = max(st.t["comp"], st.t["comm"])
            pending = list(scope.children)
            while pending:
                still = []
                for op in pending:
                    ok = yield from self._try_blocking(st, op, submit)
                    if not ok:
                        still.append(op)
                pending = still
                if pending:
                    yield ("barrier", pending[0].gid)
        else:
            children = scope.children if direction == "fwd" else list(reversed(scope.children))
            for child in children:
                yield from self._exec(st, child, direction, phase)
        st.t["comp"] += SCOPE_GAP_MS
        if prof is not None and mt is not None:
            mt.phase_end(rank=st.rank, ts=st.t["comp"], profile=prof, phase=phase)
        if scope.name:                       # anonymous wrapper scopes (p2p bundles) cost time but are not worth a trace record
            self._emit(st, scope.name, "scope", phase, t0, st.t["comp"])

    # ------------------------------------------------------------------ synchronous collectives
    def _exec_coll(self, st, job: Coll, direction: str):
        cost = job.fwd_cost if direction == "fwd" else job.bwd_cost
        if cost == 0 or job.group_size <= 1:
            return
        key = (direction, job.gid)
        if key in self._completed_coll[st.rank]:
            return
        issue_t = st.t["comp"]
        lane = (st.rank, "comm")
        if self.merge_lanes and not job.world:
            launch = max(issue_t, self._lane_tail[lane])
            end = launch + cost
            self._lane_tail[lane] = end
            start_vis = launch
        else:
            expected = len(self.ranks) if job.world else job.group_size
            ready = max(issue_t, self._lane_tail[lane])
            s = self._barrier.get(key)
            if s is None:
                s = {"expected": expected, "arrived": 0, "max_ready": 0.0, "waiters": [], "cost": cost}
                self._barrier[key] = s
            s["arrived"] += 1
            s["max_ready"] = max(s["max_ready"], ready)
            s["waiters"].append(st.rank)
            if s["arrived"] == s["expected"]:
                self._barrier_done[key] = s["max_ready"] + s["cost"]
                del self._barrier[key]
                yield ("progress", key)
            while key not in self._barrier_done:
                yield ("barrier", key)
            end = self._barrier_done[key]
            self._lane_tail[lane] = max(self._lane_tail[lane], end)
            start_vis = issue_t
        st.t["comm"] = max(st.t["comm"], end)
        st.t["comp"] = max(st.t["comp"], end)
        self._completed_coll[st.rank].add(key)
        if end != start_vis:
            self._emit(st, job.name, "comm", "fwd" if direction == "fwd" else "bwd", start_vis, end, lane="comm",
                       gid=job.gid)

    # ------------------------------------------------------------------ blocking p2p
    def _try_blocking(self, st, job: BlockingP2P, submit_t):
        """One attempt of a blocking send/recv inside a batched group; returns True when complete."""
        key = ("fwd", job.gid)
        state = getattr(job, "_state", None)
        if state == "done":
            return True
        if state is None:
            m = max(st.t["comp"], st.t["comm"])
            st.t["comp"] = st.t["comm"] = m
            ready = submit_t if submit_t is not None else m
            job._start = ready
            s = self._barrier.get(key)
            if s is None:
                s = {"expected": 2, "arrived": 0, "max_ready": 0.0, "waiters": [], "cost": job.cost}
                self._barrier[key] = s
            s["arrived"] += 1
            s["max_ready"] = max(s["max_ready"], ready)
            s["waiters"].append(st.rank)
            job._state = "waiting"
            if s["arrived"] == 2:
                self._barrier_done[key] = s["max_ready"] + s["cost"]
                del self._barrier[key]
                yield ("progress", key)
        if key not in self._barrier_done:
            return False
        end = max(self._barrier_done[key], st.t["comp"], st.t["comm"])
        st.t["comp"] = st.t["comm"] = end
        job._state = "done"
        self._emit(st, job.name, "p2p", "fwd", job._start, self._barrier_done[key], lane="comm", gid=job.gid)
        return True

    def _exec_blocking(self, st, job: BlockingP2P, submit_t):
        while True:
            ok = yield from self._try_blocking(st, job, submit_t)
            if ok:
                return
            yield ("barrier", job.gid)

    # ------------------------------------------------------------------ async p2p
    def _post_async(self, st, job: AsyncPost):
        key = ("fwd", job.gid)
        a = self._async.setdefault(key, {"send": None, "recv": None, "ready_t": None})
        if a[job.role] is not None:
            return
        entry = {"gid": key, "rank": st.rank, "stream": job.stream, "cost": job.cost, "issue_t": st.t["comp"],
                 "status": "queued", "ready_t": None, "launch_t": None, "end_t": None, "name": job.name,
                 "role": job.role, "order": self._issue_seq}
        self._issue_seq += 1
        a[job.role] = entry
        self._lane_queue[(st.rank, job.stream)].append(entry)
        self._pump()

    def _pump(self):
        progressed = True
        while progressed:
            progressed = False
            for lane in sorted(self._lane_queue):
                q = self._lane_queue[lane]
                if not q:
                    continue
                e = q[0]
                if e["status"] != "queued":
                    continue
                e["ready_t"] = max(e["issue_t"], self._lane_tail[lane])
                e["status"] = "waiting"
                progressed = True
                a = self._async[e["gid"]]
                s, r = a["send"], a["recv"]
                if s is not None and r is not None and s["status"] == "waiting" and r["status"] == "waiting":
                    end = max(s["ready_t"] + s["cost"], r["ready_t"] + r["cost"])
                    for x in (s, r):
                        lane_x = (x["rank"], x["stream"])
                        assert self._lane_queue[lane_x][0] is x, "async p2p completed out of FIFO order"
                        x["launch_t"] = max(x["ready_t"], end - x["cost"])
                        x["end_t"] = end
                        x["status"] = "done"
                        self._lane_queue[lane_x].popleft()
                        self._lane_tail[lane_x] = end
                        rs = self.ranks.get(x["rank"])
                        if rs is not None:
                            rs.t[x["stream"]] = max(rs.t[x["stream"]], end)
                    a["ready_t"] = end
                    for x in (s, r):
                        self.events.append(Event(x["rank"], x["name"], "p2p", "fwd", x["launch_t"], x["end_t"],
                                                 lane=x["stream"], gid=f"{x['gid'][0]}:{x['gid'][1]}", post=x["issue_t"],
                                                 order=x["order"]))

    def _exec_async_wait(self, st, job: AsyncWait):
        key = ("fwd", job.gid)
        a = self._async.setdefault(key, {"send": None, "recv": None, "ready_t": None})
        if a["recv"] is None:
            self._post_async(st, AsyncPost(job.name.replace("async_wait_recv", "async_recv"), job.gid, job.cost,
                                           job.stream, "recv"))
            yield ("yield", job.gid)
        while a["ready_t"] is None:
            self._pump()
            if a["ready_t"] is not None:
                break
            yield ("async_wait", key)
        t0 = st.t["comp"]
        st.t["comp"] = max(st.t["comp"], a["ready_t"])
        recv_end = a["recv"]["end_t"]
        if st.t["comp"] > t0 and a["ready_t"] > t0:
            self._emit(st, job.name, "wait", "fwd", t0, st.t["comp"], lane="wait", gid=f"fwd:{job.gid}")

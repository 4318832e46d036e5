"""Dev tool: record-by-record comparison of ``simulate()``'s ``log.log`` and ``tracing_logs.json`` with the reference (CPU only).

``compare_sim.py`` checks the iteration end time and the memory peaks; this one checks the memory snapshot / result / memory_viz
pickle files and every finished job and every Chrome-trace
record (lane, name, category, call stack, gid, [gN] correlation tag, post order, flow arrows, metadata; timestamps within the
1 ns rounding of the reference's text-log round trip).  Documented, deliberate differences are counted separately:

* ``wait``     this package puts every stall of the compute stream on an async receive on the ``09_wait`` lane; the reference
               only writes a wait record when the pair became ready after the receive slot finished (base_struct.py:1990-1997).
* ``dropped``  the reference writes the zero-cost helper modules (MLA Split / Concat / RotaryEmbedding, shared-expert Add) without the rank prefix
               (LeafModel.call_stk, base_struct.py:2007), its own parser then drops those lines; they are kept here.
* ``world``    per-rank (``merge_lanes=False``) runs: the optimizer's collectives start at the group's common time in the reference
               log but are drawn from each rank's own arrival here (docs/RESULTS.md); only ``OptimizerSimulator`` records move.
* ``clamp``    blocking VPP: a batched send that is retried after a longer sibling receive completes hands its *local* clamped
               time to the peer (base_struct.py:2294-2298), so the peer's batch scope closes late; here the peer's receive ends
               when the data arrived.  Intermediate job times of such ranks differ, the iteration end time does not.

usage: python tools/deep_compare_trace.py [case index ...]      (cases of tools/compare_sim.py)
"""
import collections, io, json, os, re, sys, tempfile, warnings
from contextlib import redirect_stdout
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import compare_sim as cs                                             # noqa: E402

TOL_US = 5e-3
_UNNAMED_IN_REF = ("kv_combined_Split", "KV_Split", "K_pos_emb_Concat", "QKV_Concat", "q_Split", "Q_Concat", "SharedExpertAddFunction")


def run(which, case):
    strategy, model, system, layers, mover, sover, merge = case
    if which == "ref":
        sys.path.insert(0, cs.REF)
        from simumax.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax_b200.core.perf_llm import PerfLLM
    st = StrategyConfig.init_from_config_file(f"{cs.CFG}/strategy/{strategy}.json")
    for k, v in sover.items():
        setattr(st, k, v)
    p = PerfLLM()
    p._search_verbose = False
    d = tempfile.mkdtemp()
    with redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=ModelConfig.init_from_config_file(f"{cs.CFG}/models/{model}.json"),
                    system_config=SystemConfig.init_from_config_file(f"{cs.CFG}/system/{system}.json"))
        p.model_config.layer_num = layers
        for k, v in mover.items():
            setattr(p.model_config, k, v)
        p.run_estimate()
        p.simulate(d, merge_lanes=merge)
    t = json.load(open(os.path.join(d, "tracing_logs.json")))
    LAST_DIR[which] = d
    return t["traceEvents"] if isinstance(t, dict) else t


LAST_DIR = {}


def memory_files_differ():
    """simu_memory_snapshot.json, simu_memory_result.json (minus this repo's extra ``peak_op_by_rank``) and the memory_viz
    pickle of the two most recent runs; returns the list of differing files (op names of the helper modules the reference
    leaves unnamed are normalised like in the trace)."""
    import pickle
    bad = []
    a, b = LAST_DIR.get("ref"), LAST_DIR.get("new")
    if not a or not b or not os.path.exists(os.path.join(a, "simu_memory_snapshot.json")):
        return bad
    pat = re.compile(r"self(\.[A-Za-z0-9_]+)*\.(" + "|".join(_UNNAMED_IN_REF) + r")")

    def norm(obj):
        return json.loads(pat.sub("self", json.dumps(obj, default=str, sort_keys=True)))

    def per_rank(records):
        """Records are appended in the order the host loop visits the ranks (not a property of the schedule): compare each
        rank's own sequence; token ids are renumbered per rank for the same reason."""
        out, ids = {}, {}
        for r in records:
            r = dict(r)
            if str(r.get("op_name", "")).split(".")[-1] in _UNNAMED_IN_REF or r.get("op_name") == "self":
                r.update(op_name="self", scope="", microbatch=None, chunk=None)   # helper modules the reference leaves unnamed
            if str(r.get("op_name", "")).endswith(".rotary_pos_emb"):             # named, but its scope lacks the rank prefix
                r.update(scope="", microbatch=None, chunk=None)
            if "token_id" in r:
                r["token_id"] = ids.setdefault((r["rank"], r["token_id"]), len([k for k in ids if k[0] == r["rank"]]))
            out.setdefault(r["rank"], []).append(r)
        return out
    x, y = (json.load(open(os.path.join(d, "simu_memory_snapshot.json"))) for d in (a, b))
    if any(per_rank(x[k]) != per_rank(y[k]) for k in ("events", "cache_tokens")) or \
            {k: v for k, v in x.items() if k not in ("events", "cache_tokens")} != {k: v for k, v in y.items() if k not in ("events", "cache_tokens")}:
        bad.append("simu_memory_snapshot.json")
    x, y = (json.load(open(os.path.join(d, "simu_memory_result.json"))) for d in (a, b))
    y.pop("peak_op_by_rank", None)
    if x != y:
        bad.append("simu_memory_result.json")
    x = pickle.load(open(os.path.join(a, "simu_memory_viz_snapshot.pickle"), "rb"))
    y = pickle.load(open(os.path.join(b, "simu_memory_viz_snapshot.pickle"), "rb"))
    if norm(x) != norm(y):
        bad.append("simu_memory_viz_snapshot.pickle")
    return bad


def _key(r):
    a = r.get("args") or {}
    stack = a.get("call_stack")
    if r.get("ph") == "C":          # memory counters: every field must agree
        if str(a.get("op_name", "")).split(".")[-1] in _UNNAMED_IN_REF:
            a = dict(a, op_name="self")                    # the reference never names these helper modules
        return ("C", r.get("pid"), None, r.get("name"), r.get("cat"), None, json.dumps(a, sort_keys=True))
    # the [gN] tag and post_order number pairs / posts in the order the host loop happened to visit the ranks, which is not a
    # property of the schedule: compare the untagged name here and the pairing itself through the flow records ("pair:<gid>")
    return (r.get("ph"), r.get("pid"), r.get("tid"), re.sub(r"\[g\d+\]$", "", r.get("name") or ""), r.get("cat"),
            tuple(stack) if stack else None, a.get("gid"),
            a.get("correlation_id") is not None, a.get("lane"), a.get("direction"), a.get("post_order") is not None,
            json.dumps(a.get("sort_index")) if r.get("ph") == "M" else None, (a.get("name") if r.get("ph") == "M" else None))


def compare(ref, new):
    """-> dict of counters; ``diff`` is what is neither identical nor one of the documented differences."""
    out = collections.Counter()
    ga, gb = collections.defaultdict(list), collections.defaultdict(list)
    for r in ref:
        ga[_key(r)].append(r)
    for r in new:
        gb[_key(r)].append(r)
    examples = []
    stacks_ref = {k[5] for k in ga if k[5]}
    for k in set(ga) | set(gb):
        a, b = ga.get(k, []), gb.get(k, [])
        if k[0] == "M":
            # thread metadata exists per lane actually used; extra lanes (wait) follow the extra records
            if len(a) != len(b):
                out["meta_extra" if len(b) > len(a) and "wait" in str(k[2]) else "diff"] += abs(len(a) - len(b))
                if not (len(b) > len(a) and "wait" in str(k[2])):
                    examples.append(("meta", k[:4], len(a), len(b)))
            else:
                out["same"] += len(a)
            continue
        if len(a) != len(b):
            if not a and k[4] == "wait":
                out["wait"] += len(b)
            elif not a and k[5] and k[5] not in stacks_ref and k[3] in _UNNAMED_IN_REF + ("RotaryEmbedding",):
                out["dropped"] += len(b)
            else:
                out["diff"] += abs(len(a) - len(b))
                examples.append(("count", k[:6], len(a), len(b)))
            continue
        a.sort(key=lambda r: r.get("ts", 0))
        b.sort(key=lambda r: r.get("ts", 0))
        for x, y in zip(a, b):
            if abs(x.get("ts", 0) - y.get("ts", 0)) <= TOL_US and abs(x.get("dur", 0) - y.get("dur", 0)) <= 2 * TOL_US:
                out["same"] += 1
            else:
                out["time"] += 1
                stack = (x.get("args") or {}).get("call_stack") or ()
                if "OptimizerSimulator" in stack or (x.get("args") or {}).get("op_name") in ("optimizer", "optimizer_step"):
                    out["time_optimizer"] += 1
                if len(examples) < 6:
                    examples.append(("time", k[1:4], round(x.get("ts", 0), 3), round(x.get("dur", 0), 3), round(y.get("ts", 0), 3),
                                     round(y.get("dur", 0), 3)))
    return out, examples


def main():
    only = [int(x) for x in sys.argv[1:]]
    bad = 0
    for i, case in enumerate(cs.CASES):
        if only and i not in only:
            continue
        ref, new = run("ref", case), run("new", case)
        c, ex = compare(ref, new)
        sync_vpp = "vp" in case[0] and not case[5].get("pp_comm_async", "sync" not in case[0])
        world_opt = not case[6] and c["time"] == c["time_optimizer"]      # documented in docs/RESULTS.md (world-mode optimizer barrier)
        unexplained = 0 if sync_vpp else c["diff"] + (0 if world_opt else c["time"])
        mem_bad = [] if (sync_vpp or world_opt) else memory_files_differ()
        unexplained += len(mem_bad)
        bad += bool(unexplained)
        print(("OK  " if not unexplained else "DIFF"), i, case[0], case[1], case[5], "| records", len(ref), len(new),
              "| identical", c["same"], "| wait-lane extras", c["wait"] + c["meta_extra"], "| kept (dropped by ref)", c["dropped"],
              "| retimed" + (" (documented clamp)" if sync_vpp and c["time"] else " (world-mode optimizer barrier)" if world_opt and
                             c["time"] else ""), c["time"], "| other", c["diff"], "| memory files", "identical" if not mem_bad else mem_bad)
        if unexplained:
            for e in ex[:6]:
                print("    ", e)
    print("cases with unexplained differences:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
